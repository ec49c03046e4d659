"""TEST INFRASTRUCTURE ONLY - never imported by the product path (gen6d_b200/).

CPU restatement (numpy) of the reference's pose-evaluation metrics:
/root/reference/utils/pose_utils.py:149-171 compute_pose_errors (projection error, ADD error),
:187-215 compute_metrics_impl (ADD-0.1d, Prj-5, ADD-S for symmetric objects) and the helpers they
call, utils/base_utils.py:256-265 project_points and :390-394 transform_points_pose.
Pinned by tests/test_metrics_oracle.py against tests/golden/metrics_golden.npz, which
tests/golden/make_golden_metrics.py produced by running the unmodified reference functions.

The rotation / translation error pair of compute_pose_errors (:160-170) is not restated: the
reference computes it but compute_metrics_impl never reads it.
"""
import numpy as np


def transform_points(pts, pose):
    """base_utils.py:390-394."""
    return pts @ pose[:, :3].T + pose[:, 3][None, :]


def project_points(pts, pose, K):
    """base_utils.py:256-265, including its depth clamp as written: depths with 0 < |d| < 1e-4 become
    +1e-4 (whatever their sign); the second mask of the reference (|d| > -1e-4 and |d| < 0) is empty."""
    p = pts @ pose[:, :3].T + pose[:, 3:].T
    p = p @ K.T
    d = p[:, 2].copy()
    tiny = (np.abs(d) < 1e-4) & (np.abs(d) > 0)
    d[tiny] = 1e-4
    return p[:, :2] / d[:, None], d


def pose_errors(pts, pose_pr, pose_gt, K, symmetric=False):
    """(prj_err, obj_err, obj_err_sym or None): pose_utils.py:149-158 and :192-196."""
    prj = np.mean(np.linalg.norm(project_points(pts, pose_pr, K)[0] - project_points(pts, pose_gt, K)[0], 2, 1))
    a, b = transform_points(pts, pose_pr), transform_points(pts, pose_gt)
    obj = np.mean(np.linalg.norm(a - b, 2, 1))
    sym = None
    if symmetric:
        sym = np.mean(np.min(np.linalg.norm(a[:, None] - b[None, :], 2, 2), 1))
    return prj, obj, sym


def compute_metrics_impl(object_pts, diameter, pose_gt_list, pose_pr_list, Ks, scale=1.0, symmetric=False):
    """pose_utils.py:187-215 (same arguments, same result dict)."""
    prj, obj, sym = [], [], []
    for gt, pr, K in zip(pose_gt_list, pose_pr_list, Ks):
        p, o, s = pose_errors(object_pts, pr, gt, K, symmetric)
        prj.append(p)
        obj.append(o * scale)
        if symmetric:
            sym.append(s * scale)
    res = {'add-0.1d': np.mean(np.asarray(obj) < diameter * 0.1), 'prj-5': np.mean(np.asarray(prj) < 5)}
    if symmetric:
        res['add-0.1d-sym'] = np.mean(np.asarray(sym) < diameter * 0.1)
    return res
