"""Host geometry (gen6d_b200/geometry.py, database.py) against golden vectors produced by the
reference's utils/*.py and refiner host code (tests/golden/make_golden_estimator.py).  CPU-only."""
import os

import numpy as np
import pytest

from golden import cases
from gen6d_b200 import geometry as G
from gen6d_b200.database import SyntheticObjectDatabase

HERE = os.path.dirname(os.path.abspath(__file__))
E = np.load(os.path.join(HERE, 'golden', 'est_golden.npz'))


@pytest.fixture(scope='module')
def db():
    return SyntheticObjectDatabase(**cases.estimator_case()['db'])


def img_close(a, b, frac=0.01):
    """Warped uint8 images: identical up to interpolation rounding on a few pixels."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32))
    assert d.mean() < 0.5 and (d > 8).mean() < frac, (d.mean(), (d > 8).mean())


def test_fps_view_selection(db):
    ids = G.select_views_fps(db, db.get_img_ids(), 64)
    assert [int(i) for i in ids] == E['geo.fps_ids'].tolist()


def test_normalize_reference_views(db):
    ids = [str(i) for i in E['geo.fps_ids'][:6]]
    imgs, Ks, poses, Hs = G.normalize_reference_views(db, ids, 128, 0.05)
    np.testing.assert_allclose(Ks, E['geo.norm.Ks'], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(poses, E['geo.norm.poses'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(Hs, E['geo.norm.Hs'], rtol=1e-4, atol=1e-4)
    for a, b in zip(imgs[:3], E['geo.norm.imgs']):
        img_close(a, b)


def test_crop_similarity(db):
    crop, M = G.crop_similarity(db.get_image('3'), np.array([300.5, 260.25], np.float32), 1 / 1.37, 0, 128)
    np.testing.assert_allclose(M, E['geo.crop.M'], atol=1e-4)
    img_close(crop, E['geo.crop.img'])


def test_pose_from_similarity():
    pose = G.pose_from_similarity(np.array([310.0, 225.0], np.float32), np.float32(1.21), np.float32(0.33),
                                  E['geo.norm.poses'][2], E['geo.norm.Ks'][2],
                                  SyntheticObjectDatabase(n_views=1).K, np.zeros(3, np.float32))
    np.testing.assert_allclose(pose, E['geo.sim_pose'], rtol=1e-5, atol=1e-4)


def test_refine_problem_and_pose_update(db):
    in_pose = E['geo.refine.in_pose']
    prob = G.refine_problem(db, db.get_img_ids(), db.get_image('5'), db.K, in_pose, 128, 6, True)
    np.testing.assert_allclose(prob['que_K'], E['geo.refine.que_K'], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(prob['que_pose'], E['geo.refine.que_pose'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(prob['ref_Ks'], E['geo.refine.ref_Ks'], rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(prob['ref_poses'], E['geo.refine.ref_poses'], rtol=1e-4, atol=1e-4)
    img_close(prob['que_img'], E['geo.refine.que_img'])
    img_close(prob['ref_imgs'][0], E['geo.refine.ref_img0'])
    f = E['geo.refine.fixed']
    pose = G.apply_refinement(prob, quat=f[:4], offset=f[4:6], scale=2.0 ** f[6])
    np.testing.assert_allclose(pose, E['geo.refine.pose_out'], rtol=1e-4, atol=1e-4)


def test_batched_refinement_geometry_equals_per_frame(db):
    """refine_problems / apply_refinements (one pass of stacked numpy operations for a batch of frames) give every
    frame bit for bit what refine_problem / apply_refinement give it alone, for float32 and float64 input poses."""
    ids = db.get_img_ids()
    rng = np.random.RandomState(3)
    poses = []
    for i in range(12):
        p = db.get_pose(ids[rng.randint(len(ids))]).astype(np.float64).copy()
        w = rng.randn(3) * 0.05
        Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        p[:, :3] = (np.eye(3) + Wx + Wx @ Wx / 2) @ p[:, :3]        # a small (not exactly orthonormal) rotation is fine here
        p[:, 3] += rng.randn(3) * 0.02
        poses.append(p)
    outs = (rng.randn(12, 7) * 0.05).astype(np.float32)
    outs[:, 0] += 1
    keys = ('que_K', 'que_pose', 'pose_rect', 'que_H', 'ref_Ks', 'ref_poses', 'ref_Hs')
    for dt in (np.float32, np.float64):
        ps = [p.astype(dt) for p in poses]
        batch = G.refine_problems(db, ids, [db.K] * len(ps), ps, 128, 6, True)
        new = G.apply_refinements(batch, outs[:, :4], outs[:, 4:6], [2.0 ** o[6] for o in outs])
        assert new.dtype == np.float32 and new.shape == (12, 3, 4)
        for i, p in enumerate(ps):
            one = G.refine_problem(db, ids, None, db.K, p, 128, 6, True, warp=False)
            assert [str(v) for v in one['ref_ids']] == [str(v) for v in batch['ref_ids'][i]]
            for k in keys:
                assert one[k].dtype == batch[k].dtype and np.array_equal(one[k], batch[k][i]), (k, i)
            o = outs[i]
            assert np.array_equal(new[i], G.apply_refinement(one, quat=o[:4], offset=o[4:6], scale=2.0 ** o[6]))
    # nearest views without the FPS re-spread (ref_even=False) go through the generic per-frame selection
    b2 = G.refine_problems(db, ids[:40], [db.K] * 3, poses[:3], 128, 6, False)
    for i in range(3):
        one = G.refine_problem(db, ids[:40], None, db.K, poses[i], 128, 6, False, warp=False)
        assert np.array_equal(one['ref_poses'], b2['ref_poses'][i]) and np.array_equal(one['que_H'], b2['que_H'][i])
