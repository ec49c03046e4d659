"""The device-side camera algebra (csrc/glue_math.cuh, g6d_glue_*) through its *_host entry points -- the same
__host__ __device__ code the kernels run -- against the numpy restatement of the reference's host geometry
(gen6d_b200/geometry.py, itself pinned on reference goldens in test_geometry.py).  CPU-only."""
import numpy as np
import pytest

from golden import cases
from gen6d_b200 import geometry as G, glue
from gen6d_b200.database import SyntheticObjectDatabase


@pytest.fixture(scope='module')
def db():
    return SyntheticObjectDatabase(**cases.estimator_case()['db'])


def perturbed_poses(db, n, seed):
    ids = db.get_img_ids()
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        p = db.get_pose(ids[rng.randint(len(ids))]).astype(np.float64).copy()
        w = rng.randn(3) * 0.05
        Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        U, _, Vt = np.linalg.svd((np.eye(3) + Wx) @ p[:, :3])
        p[:, :3] = U @ Vt
        p[:, 3] += rng.randn(3) * 0.02
        out.append(p)
    return out


def test_detection_crop_jobs_equal_numpy():
    rng = np.random.RandomState(0)
    det = np.stack([rng.rand(16) * 640, rng.rand(16) * 480, 0.5 + rng.rand(16) * 1.5, rng.rand(16)], 1).astype(np.float32)
    jobs = glue.host_detection_jobs(det, 480, 640, 128, frame_ptr=4096)
    for i in range(16):
        M = G.crop_similarity(None, det[i, :2], 1 / det[i, 2], 0, 128)[1]
        np.testing.assert_array_equal(jobs['M'][i], G.affine_dst_to_src(M))          # bit for bit
        assert jobs['src'][i] == 4096 + i * 480 * 640 * 3 and jobs['rows'][i] == 480 and jobs['cols'][i] == 640


def test_initial_poses_equal_numpy(db):
    ids = [str(i) for i in G.select_views_fps(db, db.get_img_ids(), 16)]
    _, ref_Ks, ref_poses, _ = G.normalize_reference_views(db, ids, 128, 0.05, warp=False)
    info = {'poses': ref_poses, 'Ks': ref_Ks, 'center': db.object_center()}
    rng = np.random.RandomState(1)
    n = 24
    det = np.stack([rng.rand(n) * 640, rng.rand(n) * 480, 0.5 + rng.rand(n) * 1.5, rng.rand(n)], 1).astype(np.float32)
    idx = rng.randint(0, len(ids), n)
    sel = np.stack([rng.randn(n) * 0.8, rng.randn(n)], 1).astype(np.float32)
    for K in (db.K, db.K.astype(np.float64)):
        Ks = np.stack([K] * n, 0)
        want = G.poses_from_similarity(det[:, :2], det[:, 2], sel[:, 0], ref_poses[idx], ref_Ks[idx], Ks, info['center'])
        got = glue.host_initial_poses(det, idx, sel, glue.selector_refs(info), glue.cameras(Ks))
        # float64 throughout, except cos / sin of the float32 in-plane angle: numpy evaluates them in float32 with a
        # vectorised routine that is not always correctly rounded (1 ulp = 6e-8 on the rotation, 1e-7 relative on t)
        d = np.abs(got - want)
        assert np.median(d) < 1e-14
        np.testing.assert_allclose(got[:, :, :3], want[:, :, :3], rtol=0, atol=1e-7)
        np.testing.assert_allclose(got[:, :, 3], want[:, :, 3], rtol=3e-7, atol=3e-7)


def test_refine_problems_and_updates_equal_numpy(db):
    ids = db.get_img_ids()
    tables = glue.refiner_views(db, ids, 128, 6)
    poses = perturbed_poses(db, 20, 5)
    rng = np.random.RandomState(6)
    net = (rng.randn(20, 7) * 0.05).astype(np.float32)
    net[:, 0] += 1
    net[:, :4] /= np.linalg.norm(net[:, :4], axis=1, keepdims=True)
    src = 1000 + 7 * np.arange(len(ids))
    for dt in (np.float64, np.float32):
        ps = np.stack(poses, 0).astype(dt)
        Ks = np.stack([db.K] * len(ps), 0)
        want = G.refine_problems(db, ids, Ks, ps, 128, 6, True)
        got = glue.host_refine_problems(tables, glue.cameras(Ks), ps, dt == np.float32, 480, 640, frame_ptr=64, src=src,
                                        img_rows=np.full(len(ids), 480), img_cols=np.full(len(ids), 640))
        assert [[tables['ids'][r] for r in row] for row in got['ref_rows']] == [[str(v) for v in row] for row in want['ref_ids']]
        ulp = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
        # float32 outputs: identical up to the summation order of three-term float32 dot products (<= 2 ulp)
        for k in ('que_K', 'que_pose', 'pose_rect', 'ref_Ks', 'ref_poses'):
            assert ulp(got[k], want[k]) < 3e-7, (k, ulp(got[k], want[k]))
        jobs = got['jobs'].reshape(len(ps), 7)
        for i in range(len(ps)):
            Hs = [want['que_H'][i]] + list(want['ref_Hs'][i])
            for j, H in enumerate(Hs):
                np.testing.assert_allclose(jobs['M'][i, j], G.perspective_dst_to_src(H), rtol=2e-6, atol=1e-9)
            assert jobs['src'][i, 0] == 64 + i * 480 * 640 * 3
            assert list(jobs['src'][i, 1:]) == [src[r] for r in got['ref_rows'][i]]
        # the pose update, from the SAME problem arrays
        prob = {k: (want[k] if k in ('view', 'center') else np.ascontiguousarray(want[k])) for k in want}
        new = G.apply_refinements(prob, net[:, :4], net[:, 4:6], [2.0 ** o[6] for o in net])
        mine = glue.host_apply_refinements(tables, prob, net)
        assert mine.dtype == np.float64
        np.testing.assert_allclose(mine, new, rtol=0, atol=3e-7)
        np.testing.assert_array_equal(mine, mine.astype(np.float32).astype(np.float64))      # float32 values


def test_small_reference_set_and_float64_intrinsics():
    """Fewer database views than the 128-view FPS subset, intrinsics given as float64."""
    small = SyntheticObjectDatabase(**{**cases.estimator_case()['db'], 'n_views': 40})
    ids = small.get_img_ids()
    tables = glue.refiner_views(small, ids, 128, 6)
    assert len(ids) == 40 and 6 <= len(tables['even_idx']) <= 40        # the FPS re-spread of the whole (small) set
    poses = np.stack(perturbed_poses(small, 8, 9), 0)
    Ks = np.stack([small.K.astype(np.float64)] * len(poses), 0)
    want = G.refine_problems(small, ids, Ks, poses, 128, 6, True)
    got = glue.host_refine_problems(tables, glue.cameras(Ks), poses, False, 480, 640)
    assert [[tables['ids'][r] for r in row] for row in got['ref_rows']] == [[str(v) for v in row] for row in want['ref_ids']]
    for k in ('que_K', 'que_pose', 'pose_rect', 'ref_Ks', 'ref_poses'):
        assert float(np.abs(got[k] - want[k]).max() / np.abs(want[k]).max()) < 3e-7, k
