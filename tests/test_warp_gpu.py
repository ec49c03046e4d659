"""Device warps (gen6d_b200/csrc/warp.cu) against OpenCV itself: the kernels
must return the very bytes cv2.warpPerspective / cv2.warpAffine return, so that moving the
between-stage crops onto the GPU changes nothing downstream."""
import os
import sys

import cv2
import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gen6d_b200 import geometry as G, ops  # noqa: E402
from oracle import cv_warp  # noqa: E402

pytestmark = pytest.mark.gpu


def _jobs(srcs, mats):
    return torch.from_numpy(G.pack_warp_jobs(srcs, mats)).cuda()


def test_perspective_bit_exact_vs_cv2():
    rng = np.random.RandomState(0)
    imgs = [(rng.rand(100 + 37 * k, 200 - 21 * k, 3) * 255).astype(np.uint8) for k in range(5)]
    dev = [torch.from_numpy(i).cuda() for i in imgs]
    for dsize in ((128, 128), (70, 45), (300, 9)):
        Hs = []
        for k in range(len(imgs)):
            H = np.eye(3) + rng.randn(3, 3) * np.array([[0.3, 0.3, 40], [0.3, 0.3, 40], [1e-3, 1e-3, 0]])
            Hs.append(H.astype(np.float32) if k % 2 else H)
        out = ops.warp_perspective_u8(_jobs(dev, [G.perspective_dst_to_src(H) for H in Hs]), len(imgs),
                                      dsize[1], dsize[0]).cpu().numpy()
        for k, img in enumerate(imgs):
            ref = cv2.warpPerspective(img, Hs[k], dsize, flags=cv2.INTER_LINEAR)
            np.testing.assert_array_equal(out[k], ref)
            np.testing.assert_array_equal(out[k], cv_warp.warp_perspective_u8(img, Hs[k], dsize))


def test_perspective_many_random_pixels():
    """~2.6M output pixels over 160 homographies: no coordinate-rounding disagreement with OpenCV."""
    rng = np.random.RandomState(1)
    img = (rng.rand(480, 640, 3) * 255).astype(np.uint8)
    dev = torch.from_numpy(img).cuda()
    Hs = [np.eye(3) + rng.randn(3, 3) * np.array([[0.5, 0.5, 200], [0.5, 0.5, 200], [1e-3, 1e-3, 0]]) for _ in range(160)]
    out = ops.warp_perspective_u8(_jobs([dev] * len(Hs), [G.perspective_dst_to_src(H) for H in Hs]), len(Hs),
                                  128, 128).cpu().numpy()
    for k, H in enumerate(Hs):
        np.testing.assert_array_equal(out[k], cv2.warpPerspective(img, H, (128, 128), flags=cv2.INTER_LINEAR))


def test_affine_bit_exact_vs_cv2():
    rng = np.random.RandomState(2)
    img = (rng.rand(480, 640, 3) * 255).astype(np.uint8)
    dev = torch.from_numpy(img).cuda()
    Ms = []
    for _ in range(64):
        ang, s = rng.rand() * 6.28, 0.2 + rng.rand() * 3
        Ms.append(np.array([[s * np.cos(ang), -s * np.sin(ang), rng.randn() * 200 + 60],
                            [s * np.sin(ang), s * np.cos(ang), rng.randn() * 200 + 60]], np.float32))
    for dsize in ((128, 128), (50, 77)):
        out = ops.warp_affine_u8(_jobs([dev] * len(Ms), [G.affine_dst_to_src(M) for M in Ms]), len(Ms),
                                 dsize[1], dsize[0]).cpu().numpy()
        for k, M in enumerate(Ms):
            np.testing.assert_array_equal(out[k], cv2.warpAffine(img, M, dsize, flags=cv2.INTER_LINEAR))


@pytest.fixture(scope='module')
def est():
    from gen6d_b200.synthetic import build_estimator
    return build_estimator()


def test_refine_device_warps_equal_host_warps(est):
    """refine_que_imgs with the crops cut on the device == with OpenCV crops on the host."""
    e, db = est
    ids = db.get_img_ids()
    img, K = db.get_image(ids[3]), db.get_K(ids[3])
    init = db.get_pose(ids[4])
    a = e.refiner.refine_que_imgs(img, K, init, 128, 6, True, host_warps=True)
    b = e.refiner.refine_que_imgs(img, K, init, 128, 6, True)
    np.testing.assert_array_equal(a, b)


def test_predict_device_warps_equal_host_warps(est):
    """Whole predict(): same detection crop bytes, same selection, same poses either way."""
    e, db = est
    ids = db.get_img_ids()
    img, K = db.get_image(ids[7]), db.get_K(ids[7])
    e.cfg['host_warps'] = True
    pose_h, inter_h = e.predict(img, K)
    e.cfg['host_warps'] = False
    pose_d, inter_d = e.predict(img, K)
    np.testing.assert_array_equal(inter_d['det_que_img'], inter_h['det_que_img'])
    assert int(inter_d['sel_ref_idx']) == int(inter_h['sel_ref_idx'])
    np.testing.assert_array_equal(np.stack(inter_d['refine_poses']), np.stack(inter_h['refine_poses']))
    np.testing.assert_array_equal(pose_d, pose_h)
