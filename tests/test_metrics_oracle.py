"""oracle/metrics.py (row f4: ADD-0.1d / Prj-5 / ADD-S) against the goldens of the unmodified reference."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from golden import cases  # noqa: E402
from oracle import metrics as OM  # noqa: E402

G = np.load(os.path.join(HERE, 'golden', 'metrics_golden.npz'))


def test_per_pose_errors_match_reference():
    c = cases.metrics_case()
    got = [OM.pose_errors(c['pts'], pr, gt, K, True) for pr, gt, K in zip(c['pr'], c['gt'], c['Ks'])]
    np.testing.assert_allclose([g[0] for g in got], G['prj_err'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose([g[1] for g in got], G['obj_err'], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose([g[2] for g in got], G['obj_err_sym'], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize('scale', [1.0, 0.5])
@pytest.mark.parametrize('symmetric', [False, True])
def test_metric_dict_matches_reference(scale, symmetric):
    c = cases.metrics_case()
    res = OM.compute_metrics_impl(c['pts'], c['diameter'], list(c['gt']), list(c['pr']), list(c['Ks']), scale, symmetric)
    want = {k.split('.', 4)[-1]: float(G[k]) for k in G.files if k.startswith(f'res.{scale}.{int(symmetric)}.')}
    assert set(res) == set(want)
    for k in want:
        assert float(res[k]) == want[k], (k, res[k], want[k])      # rates of booleans: exact


def test_depth_clamp_quirk():
    """0 < |depth| < 1e-4 -> +1e-4 regardless of sign; exactly 0 stays 0 (division by zero like the reference)."""
    pts = np.array([[0, 0, 5e-5], [0, 0, -5e-5], [0, 0, 1.0]], np.float32)
    pose = np.concatenate([np.eye(3), np.zeros((3, 1))], 1).astype(np.float32)
    _, d = OM.project_points(pts, pose, np.eye(3, dtype=np.float32))
    np.testing.assert_allclose(d, [1e-4, 1e-4, 1.0])
