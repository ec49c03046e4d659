"""Row f4 on the device: g6d_pose_errors / gen6d_b200.metrics (ADD-0.1d, Prj-5, ADD-S) against the
goldens of the unmodified reference (utils/pose_utils.py:149-215) and against the numpy oracle."""
import os

import numpy as np
import pytest

from golden import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pose_error_metrics_match_reference():
    from gen6d_b200 import metrics as M
    from oracle import metrics as OM
    c = cases.metrics_case()
    Gm = np.load(os.path.join(ROOT, 'tests', 'golden', 'metrics_golden.npz'))
    err = M.pose_errors(c['pts'], c['pr'], c['gt'], c['Ks'], symmetric=True).cpu().numpy()
    np.testing.assert_allclose(err[:, 0], Gm['prj_err'], rtol=2e-5, atol=1e-4)       # pixels
    np.testing.assert_allclose(err[:, 1], Gm['obj_err'], rtol=2e-5, atol=1e-6)
    np.testing.assert_allclose(err[:, 2], Gm['obj_err_sym'], rtol=2e-5, atol=1e-6)
    assert np.isnan(M.pose_errors(c['pts'], c['pr'], c['gt'], c['Ks'], symmetric=False).cpu().numpy()[:, 2]).all()
    for scale in (1.0, 0.5):
        for symmetric in (False, True):
            got = M.compute_metrics_impl(c['pts'], c['diameter'], list(c['gt']), list(c['pr']), list(c['Ks']), scale, symmetric)
            want = OM.compute_metrics_impl(c['pts'], c['diameter'], list(c['gt']), list(c['pr']), list(c['Ks']), scale, symmetric)
            assert {k: float(v) for k, v in got.items()} == {k: float(v) for k, v in want.items()}
            for k, v in got.items():
                assert float(v) == float(Gm[f'res.{scale}.{int(symmetric)}.{k}'])


def test_pose_errors_accept_device_tensors():
    """Poses straight from the refiner stay on the device: no host round trip before scoring."""
    import torch
    from gen6d_b200 import metrics as M
    c = cases.metrics_case()
    host = M.pose_errors(c['pts'], c['pr'], c['gt'], c['Ks']).cpu().numpy()
    dev = M.pose_errors(torch.from_numpy(c['pts']).cuda(), torch.from_numpy(c['pr']).cuda(), torch.from_numpy(c['gt']).cuda(),
                        torch.from_numpy(c['Ks']).cuda()).cpu().numpy()
    np.testing.assert_array_equal(host, dev)
