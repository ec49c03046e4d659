"""Pose-evaluation metrics on the device: the reference's
utils/pose_utils.py:187-215 `compute_metrics_impl` (ADD-0.1d, Prj-5 and, for symmetric objects,
ADD-S) with the per-pose errors of :149-158 computed by g6d_pose_errors.  Poses may be numpy arrays
or CUDA tensors (e.g. straight from the refiner), so an evaluation loop needs one small D2H of
[q,3] errors instead of one per stage.  No CPU fallback: needs libgen6d_b200.so and a GPU.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, ops


def _dev(x, shape_tail):
    t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(x, np.float32)))
    t = t.to(device='cuda', dtype=torch.float32).contiguous()
    if tuple(t.shape[-len(shape_tail):]) != shape_tail:
        raise ValueError(f'expected trailing shape {shape_tail}, got {tuple(t.shape)}')
    return t


def pose_errors(object_pts, poses_pr, poses_gt, Ks, symmetric=False):
    """object_pts [n,3]; poses_pr / poses_gt [q,3,4]; Ks [q,3,3] -> float32 tensor [q,3] on the device:
    (reprojection error, ADD error, ADD-S error or NaN)."""
    ops.require_cuda()
    pts, pr, gt, K = _dev(object_pts, (3,)), _dev(poses_pr, (3, 4)), _dev(poses_gt, (3, 4)), _dev(Ks, (3, 3))
    n, q = pts.shape[0], pr.shape[0]
    if gt.shape[0] != q or K.shape[0] != q:
        raise ValueError('poses_pr, poses_gt and Ks must have the same leading dimension')
    out = torch.empty(q, 3, device=pts.device, dtype=torch.float32)
    nbytes = _lib.lib().g6d_pose_errors_workspace_bytes(n, q)
    if nbytes < 0:
        _lib.check(-1, 'g6d_pose_errors_workspace_bytes')
    ws = torch.empty(nbytes // 8, device=pts.device, dtype=torch.float64)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().g6d_pose_errors(ptr(pts), n, ptr(pr), ptr(gt), ptr(K), q, int(bool(symmetric)), ptr(out), ptr(ws),
                                          stream), 'g6d_pose_errors')
    return out


def compute_metrics_impl(object_pts, diameter, pose_gt_list, pose_pr_list, Ks, scale=1.0, symmetric=False):
    """Same arguments and result dict as the reference's compute_metrics_impl (pose_utils.py:187-215)."""
    stack = lambda xs: xs if isinstance(xs, torch.Tensor) else np.stack([np.asarray(x, np.float32) for x in xs], 0)
    err = pose_errors(object_pts, stack(pose_pr_list), stack(pose_gt_list), stack(Ks), symmetric).cpu().numpy().astype(np.float64)
    res = {'add-0.1d': np.mean(err[:, 1] * scale < diameter * 0.1), 'prj-5': np.mean(err[:, 0] < 5)}
    if symmetric:
        res['add-0.1d-sym'] = np.mean(err[:, 2] * scale < diameter * 0.1)
    return res
