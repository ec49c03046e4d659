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
    # The pose handed to the refiner (detection + selection + similarity->pose) must agree.  The
    # refined poses of THIS chain are not compared: with random weights the detector's scale is
    # ~0.07, the implied object distance is ~40x off, the refiner then looks at featureless
    # background and its InstanceNorms (1/sqrt(var + 1e-5) on near-constant channels) amplify fp32
    # summation-order noise without bound.  Refinement parity is checked on a well-posed input in
    # test_tracking_refinement_matches_reference and at tensor level in test_networks_gpu.py.
    got, want = inter['refine_poses'][0], E['est.refine_poses'][0]
    np.testing.assert_allclose(got[:, :3], want[:, :3], atol=5e-3)
    np.testing.assert_allclose(got[:, 3], want[:, 3], rtol=5e-3, atol=5e-2)


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
