"""End-to-end: Gen6DEstimator.build + predict on the B200 networks vs the golden run of the
unmodified reference estimator (CPU) on the same synthetic database and seeded checkpoints."""
import os

import numpy as np
import pytest

from golden import cases

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
E = np.load(os.path.join(HERE, 'golden', 'est_golden.npz'))


@pytest.fixture(scope='module')
def est():
    from gen6d_b200.synthetic import build_estimator
    return build_estimator()


def test_build_matches_reference(est):
    e, db = est
    d = np.abs(e.ref_info['imgs'][:2].astype(np.int32) - E['est.ref_imgs0'].astype(np.int32))
    assert d.mean() < 0.5
    np.testing.assert_allclose(e.ref_info['poses'], E['est.ref_poses'], rtol=1e-4, atol=1e-4)


def test_predict_matches_reference(est):
    e, db = est
    q = str(int(E['est.query_id']))
    pose, inter = e.predict(db.get_image(q), db.get_K(q))
    s = np.sort(E['est.sel_scores'])
    print('reference selector margin', s[-1] - s[-2])
    np.testing.assert_allclose(inter['det_position'], E['est.det_position'], atol=0.05)      # pixels
    np.testing.assert_allclose(inter['det_scale_r2q'], E['est.det_scale'], rtol=5e-3)
    assert int(inter['sel_ref_idx']) == int(E['est.sel_ref_idx'])                               # bit-exact viewpoint
    np.testing.assert_allclose(inter['sel_angle_r2q'], E['est.sel_angle'], atol=2e-2)
    # The whole chain: the pose handed to the refiner (detection + selection + similarity->pose), the
    # pose after each of the three refinement iterations and the returned pose, against the golden
    # run of the unmodified reference (est.det_scale = 0.93: the seeded heads keep the chain well-posed).
    got, want = np.stack(inter['refine_poses'], 0), E['est.refine_poses']
    assert got.shape == want.shape == (4, 3, 4)
    err_r = np.abs(got[:, :, :3] - want[:, :, :3]).reshape(len(got), -1).max(1)
    err_t = np.abs(got[:, :, 3] - want[:, :, 3]).max(1) / np.linalg.norm(want[:, :, 3], axis=1)
    print('full chain, per-iteration max |dR|', err_r, 'relative |dt|', err_t)
    np.testing.assert_allclose(got[0][:, :3], want[0][:, :3], atol=5e-3)
    np.testing.assert_allclose(got[0][:, 3], want[0][:, 3], rtol=5e-3, atol=5e-2)
    np.testing.assert_allclose(got[:, :, :3], want[:, :, :3], atol=1e-2)
    assert (err_t < 1e-2).all()
    np.testing.assert_allclose(pose[:, :3], E['est.pose'][:, :3], atol=1e-2)
    np.testing.assert_allclose(pose[:, 3], E['est.pose'][:, 3], rtol=1e-2, atol=5e-2)
    np.testing.assert_array_equal(pose, inter['refine_poses'][-1])


def test_tracking_refinement_matches_reference(est):
    """predict(pose_init=...) = three refinement iterations from a perturbed ground-truth pose."""
    e, db = est
    q = str(int(E['est.query_id']))
    pose, inter = e.predict(db.get_image(q), db.get_K(q), pose_init=E['est.track_init'])
    got, want = np.stack(inter['refine_poses'], 0), E['est.track_poses']
    err_r = np.abs(got[:, :, :3] - want[:, :, :3]).reshape(len(got), -1).max(1)
    err_t = np.abs(got[:, :, 3] - want[:, :, 3]).max(1) / np.linalg.norm(want[:, :, 3], axis=1)
    print('per-iteration max |dR|', err_r, 'relative |dt|', err_t)
    np.testing.assert_allclose(got[:, :, :3], want[:, :, :3], atol=1e-2)
    assert (err_t < 1e-2).all()


def test_predict_many_equals_predict(est):
    """The pipelined throughput API returns exactly what per-frame predict() returns."""
    e, db = est
    ids = db.get_img_ids()[:4]
    imgs, Ks = [db.get_image(i) for i in ids], [db.get_K(i) for i in ids]
    seq = [e.predict(im, K)[0] for im, K in zip(imgs, Ks)]
    par = e.predict_many(imgs, Ks, workers=2)
    for a, (b, _) in zip(seq, par):
        np.testing.assert_allclose(a, b, atol=1e-5)


def test_predict_batch_equals_predict(est):
    """Row f3: qn frames through one detect / one select / one refine stage per iteration give the poses
    of per-frame predict() (the kernels are batch-independent up to the split-K summation order, which
    follows M), and the pipelined API on top of it agrees too."""
    e, db = est
    ids = db.get_img_ids()[:5]
    imgs, Ks = [db.get_image(i) for i in ids], [db.get_K(i) for i in ids]
    seq = [e.predict(im, K) for im, K in zip(imgs, Ks)]
    poses, inter = e.predict_batch(imgs, Ks)
    assert poses.shape == (5, 3, 4) and len(inter['refine_poses']) == e.cfg['refine_iter'] + 1
    assert inter['sel_ref_idx'].tolist() == [int(s[1]['sel_ref_idx']) for s in seq]        # bit-exact selections
    np.testing.assert_allclose(inter['det_position'], np.stack([s[1]['det_position'] for s in seq]), atol=1e-3)
    worst = max(float(np.abs(a[0] - b).max()) for a, b in zip(seq, poses))
    print('predict_batch vs predict, max |dpose|', worst)
    for (a, _), b in zip(seq, poses):
        np.testing.assert_allclose(a, b, atol=2e-4)
    many = e.predict_many(imgs, Ks, workers=2, batch=2)
    for (a, _), (b, one) in zip(seq, many):
        np.testing.assert_allclose(a, b, atol=2e-4)
        assert len(one['refine_poses']) == e.cfg['refine_iter'] + 1


def test_device_build_equals_host_build(est):
    """Gen6DEstimator.build with cfg['device_build'] cuts the same reference crops as the OpenCV path (row f2)."""
    from gen6d_b200.synthetic import build_estimator
    host, _ = est
    dev, _ = build_estimator(device_build=True)
    np.testing.assert_array_equal(dev.ref_info['imgs'], host.ref_info['imgs'])
    np.testing.assert_array_equal(dev.ref_info['ref_imgs'], host.ref_info['ref_imgs'])


def test_add_and_prj_match_reference_over_20_frames(est):
    """north_star acceptance: "matched ADD-0.1d on synthetic inputs".  20 frames through predict(); the
    per-frame ADD / projection errors (g6d_pose_errors, row f4) and the ADD-0.1d / Prj-5 rates against the
    database's ground truth must equal those of the poses the unmodified reference estimator produced
    (tests/golden/make_golden_add.py -> add_golden.npz, scored by the reference's utils/pose_utils.py:149-215)."""
    from gen6d_b200 import metrics as M
    e, db = est
    A = np.load(os.path.join(HERE, 'golden', 'add_golden.npz'))
    ids = [str(int(i)) for i in A['frame_ids']]
    assert ids == cases.add_frame_ids(db)
    poses, sel = [], []
    for fid in ids:
        pose, inter = e.predict(db.get_image(fid), db.get_K(fid))
        poses.append(pose)
        sel.append(int(inter['sel_ref_idx']))
    poses = np.stack(poses, 0)
    print('reference selector margins', A['sel_margin'])
    assert sel == A['sel_ref_idx'].tolist()                                     # bit-exact viewpoints, every frame
    pts, diameter = db.object_point_cloud.astype(np.float32), float(A['diameter'])
    assert diameter == db.object_diameter()
    # the reference's own poses through our metric kernel reproduce the reference's metric values ...
    err_ref = M.pose_errors(pts, A['poses_pr'], A['poses_gt'], A['Ks']).cpu().numpy()
    np.testing.assert_allclose(err_ref[:, 0], A['prj_err'], rtol=1e-4)
    np.testing.assert_allclose(err_ref[:, 1], A['obj_err'], rtol=1e-4)
    # ... and our poses score the same, frame by frame and in the rates
    err = M.pose_errors(pts, poses, A['poses_gt'], A['Ks']).cpu().numpy()
    d_add = np.abs(err[:, 1] - A['obj_err']) / A['obj_err']
    d_prj = np.abs(err[:, 0] - A['prj_err']) / A['prj_err']
    print('per-frame relative |dADD|', d_add.round(4), 'relative |dPrj|', d_prj.round(4))
    # three refinement iterations of a random-weight network amplify fp32 summation-order differences
    # (measured on the reference itself: tests/golden/make_golden_sensitivity.py); the per-frame errors
    # (22-38 x 0.1 d, 85-450 px with untrained weights) agree to a few per cent, the rates exactly
    assert d_add.max() < 0.05 and d_prj.max() < 0.05
    got = M.compute_metrics_impl(pts, diameter, list(A['poses_gt']), list(poses), list(A['Ks']))
    assert float(got['add-0.1d']) == float(A['res.add-0.1d']) and float(got['prj-5']) == float(A['res.prj-5'])
    # the thresholds are not vacuous for this kernel: ground-truth poses score 100 %, and the reference's
    # per-frame pass/fail pattern at a loose threshold is reproduced exactly
    perfect = M.compute_metrics_impl(pts, diameter, list(A['poses_gt']), list(A['poses_gt']), list(A['Ks']))
    assert float(perfect['add-0.1d']) == 1.0 and float(perfect['prj-5']) == 1.0
    thr = float(np.median(A['obj_err']))
    clear = np.abs(A['obj_err'] - thr) > 0.05 * thr                    # frames not within 5 % of the threshold
    assert ((err[:, 1] < thr) == (A['obj_err'] < thr))[clear].all()
