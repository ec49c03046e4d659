"""Pins oracle/cv_warp.py (the restatement of OpenCV's 8-bit fixed-point warps that the device
kernels of gen6d_b200/csrc/warp.cu follow) against cv2 itself, bit for bit."""
import os
import sys

import cv2
import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cv_warp  # noqa: E402
from gen6d_b200 import geometry as G  # noqa: E402


def random_homography(rng, strength=1.0):
    H = np.eye(3) + rng.randn(3, 3) * np.array([[0.3, 0.3, 40], [0.3, 0.3, 40], [1e-3, 1e-3, 0]]) * strength
    return H


@pytest.mark.parametrize('seed', range(6))
def test_perspective_matches_cv2(seed):
    rng = np.random.RandomState(seed)
    src = (rng.rand(97 + 13 * seed, 160 - 7 * seed, 3) * 255).astype(np.uint8)
    H = random_homography(rng)
    if seed % 2:
        H = H.astype(np.float32)
    for dsize in ((128, 128), (70, 45), (200, 9)):
        ref = cv2.warpPerspective(src, H, dsize, flags=cv2.INTER_LINEAR)
        np.testing.assert_array_equal(cv_warp.warp_perspective_u8(src, H, dsize), ref)


@pytest.mark.parametrize('seed', range(6))
def test_affine_matches_cv2(seed):
    rng = np.random.RandomState(100 + seed)
    src = (rng.rand(120, 160, 3) * 255).astype(np.uint8)
    ang, s = rng.rand() * 6.28, 0.4 + rng.rand() * 2
    M = np.array([[s * np.cos(ang), -s * np.sin(ang), rng.randn() * 40 + 30],
                  [s * np.sin(ang), s * np.cos(ang), rng.randn() * 40 + 30]], np.float32)
    for dsize in ((128, 128), (50, 77)):
        ref = cv2.warpAffine(src, M, dsize, flags=cv2.INTER_LINEAR)
        np.testing.assert_array_equal(cv_warp.warp_affine_u8(src, M, dsize), ref)


def test_host_inverses_match_oracle():
    """geometry.perspective_dst_to_src / affine_dst_to_src hand the kernels the very matrices OpenCV
    iterates with."""
    rng = np.random.RandomState(7)
    for _ in range(20):
        H = random_homography(rng)
        np.testing.assert_array_equal(G.perspective_dst_to_src(H), cv_warp.invert_3x3(H).reshape(9))
        M = rng.randn(2, 3).astype(np.float32)
        np.testing.assert_array_equal(G.affine_dst_to_src(M)[:6], cv_warp.invert_2x3(M))


def test_warp_job_layout():
    """numpy record == struct g6d_warp_job of include/gen6d_b200.h (pointer, 2 ints, 9 doubles)."""
    assert G.WARP_JOB.itemsize == 88
    assert [G.WARP_JOB.fields[k][1] for k in ('src', 'rows', 'cols', 'M')] == [0, 8, 12, 16]


def test_refine_problem_homographies_reproduce_crops():
    """refine_problem(warp=False) returns the homographies whose warps are the crops of warp=True."""
    from gen6d_b200.database import SyntheticObjectDatabase
    db = SyntheticObjectDatabase(n_views=24, seed=3)
    ids = db.get_img_ids()
    q = ids[5]
    full = G.refine_problem(db, ids, db.get_image(q), db.get_K(q), db.get_pose(ids[6]), 128, 6, True)
    lean = G.refine_problem(db, ids, None, db.get_K(q), db.get_pose(ids[6]), 128, 6, True, warp=False)
    assert lean['que_img'] is None and lean['ref_imgs'] is None
    np.testing.assert_array_equal(full['ref_ids'], lean['ref_ids'])
    np.testing.assert_array_equal(cv_warp.warp_perspective_u8(db.get_image(q), lean['que_H'], (128, 128)), full['que_img'])
    for k, i in enumerate(lean['ref_ids']):
        np.testing.assert_array_equal(cv_warp.warp_perspective_u8(db.get_image(i), lean['ref_Hs'][k], (128, 128)),
                                      full['ref_imgs'][k])
