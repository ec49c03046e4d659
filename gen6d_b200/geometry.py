"""Host-side geometry of the Gen6D estimator (numpy + OpenCV), restated for this package.

These are the small closed-form steps that sit between the three networks: look-at crops,
in-plane alignment of reference views, farthest-point view selection, similarity -> rigid pose
conversion.  Everything is float64 internally and cast to float32 at the same points the
reference does.  Reference call sites: utils/base_utils.py:256-266,502-524,558-666,
utils/pose_utils.py:12-111,217-244, utils/database_utils.py:8-139, dataset/database.py:400-410,
667-694, network/refiner.py:275-341.

2-D similarity / affine transforms are 3x3 homogeneous matrices here (the reference composes
2x3 blocks); poses are [R|t] 3x4 world->camera.
"""
import cv2
import numpy as np


def usable_cpus():
    """CPUs this process may really use: affinity mask capped by the cgroup quota."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    try:
        quota, period = open('/sys/fs/cgroup/cpu.max').read().split()
        if quota != 'max':
            n = min(n, max(1, int(float(quota) / float(period))))
    except Exception:
        pass
    return max(1, n)


def configure_host_threads(n=None):
    """The per-frame host work is a handful of 128x128 warps, 3x3 linear algebra and sub-megabyte
    staging copies.  OpenCV / OpenMP pools sized for every visible core (128 on the GPU hosts, with
    a 16-CPU quota) make each of those calls slower by an order of magnitude, so the estimator caps
    them (default: min(8, usable CPUs))."""
    import torch
    n = n or min(8, usable_cpus())
    cv2.setNumThreads(n)
    torch.set_num_threads(n)
    return n


# ------------------------------------------------------------------------------------------ poses
def pose_inverse(pose):
    Rt = pose[:, :3].T
    return np.concatenate([Rt, -Rt @ pose[:, 3:]], 1)


def pose_compose(first, second):
    """Apply `first`, then `second` (base_utils.py:512-521)."""
    return np.concatenate([second[:, :3] @ first[:, :3], second[:, :3] @ first[:, 3:] + second[:, 3:]], 1)


def pose_apply(pose, pts):
    return pts @ pose[:, :3].T + pose[:, 3]


def camera_center(pose):
    return -pose[:, :3].T @ pose[:, 3]


def project(pts, pose, K):
    """World points -> pixels, depths (base_utils.py:256-266).  Depths with 0 < |d| < 1e-4 are
    clamped to 1e-4, as the reference does."""
    p = (pts @ pose[:, :3].T + pose[:, 3]) @ K.T
    d = p[:, 2].copy()
    tiny = (np.abs(d) < 1e-4) & (np.abs(d) > 0)
    d[tiny] = 1e-4
    return p[:, :2] / d[:, None], d


def quat_to_matrix(q):
    """(w, x, y, z) -> 3x3, the transforms3d.quaternions.quat2mat convention (pose_utils.py:239)."""
    w, x, y, z = [float(v) for v in q]
    n = w * w + x * x + y * y + z * z
    if n < np.finfo(np.float64).eps:
        return np.eye(3)
    s = 2.0 / n
    return np.array([[1 - s * (y * y + z * z), s * (x * y - w * z), s * (x * z + w * y)],
                     [s * (x * y + w * z), 1 - s * (x * x + z * z), s * (y * z - w * x)],
                     [s * (x * z - w * y), s * (y * z + w * x), 1 - s * (x * x + y * y)]])


def rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def look_at_rotation(xy):
    """Rotation taking the ray through normalised image point (x, y, 1) onto the optical axis:
    first about y by -atan(x), then about x by +atan(y) (base_utils.py:657-666)."""
    x, y = float(xy[0]), float(xy[1])
    a, b = -np.arctan2(x, 1.0), np.arctan2(y, 1.0)
    ry = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
    rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    return rx @ ry


def look_at_pixel(center_px, K):
    """(R, f): rotation that centres pixel `center_px`, and the focal length along that ray
    (pose_utils.py:52-58 let_me_look_at_2d)."""
    f = (K[0, 0] + K[1, 1]) / 2
    c = np.asarray(center_px, np.float64) - K[:2, 2]
    return look_at_rotation(c / f), np.sqrt(c @ c + f * f)


def look_at_point(pose, K, point):
    px, _ = project(np.asarray(point, np.float64)[None], pose, K)
    return look_at_pixel(px[0], K)


def inplane_angle_between(ref_pose, ref_K, que_pose, que_K, center):
    """In-plane rotation (about the optical axis) that takes the look-at-rectified reference
    view to the look-at-rectified query view: first angle of the static z-y-x Euler split of
    R_que R_ref^T (pose_utils.py:60-102; only the angle output is used on the path)."""
    Rr = look_at_point(ref_pose, ref_K, center)[0] @ ref_pose[:, :3]
    Rq = look_at_point(que_pose, que_K, center)[0] @ que_pose[:, :3]
    rel = Rq @ Rr.T
    # R = Rx(c) Ry(b) Rz(a)  =>  first row = [cos b cos a, -cos b sin a, sin b]
    return float(np.arctan2(-rel[0, 1], rel[0, 0]))


# ------------------------------------------------------------------------------------------ warps
def look_at_crop(img, K, pose, position, angle, scale, h, w):
    """Rotate the camera to look at pixel `position`, spin by `angle`, zoom by `scale`, and cut
    an (h, w) window (database_utils.py:8-25).  Returns img, K_new, pose_new, pose_rect, H."""
    R, f = look_at_pixel(position, K)
    R = rot_z(angle).astype(np.float32) @ R           # reference builds R_z in float32
    f = f * scale
    K_new = np.array([[f, 0, w / 2], [0, f, h / 2], [0, 0, 1]], np.float32)
    H = K_new @ R @ np.linalg.inv(K)
    out = cv2.warpPerspective(img, H, (w, h), flags=cv2.INTER_LINEAR) if img is not None else None
    rect = np.concatenate([R, np.zeros((3, 1))], 1).astype(np.float32)
    return out, K_new, pose_compose(pose, rect), rect, H


def similarity_2d(position, scale, angle, target):
    """3x3 matrix: translate `position` to the origin, scale, rotate, translate to `target`."""
    c, s = np.cos(angle), np.sin(angle)
    A = scale * np.array([[c, -s], [s, c]])
    M = np.eye(3)
    M[:2, :2] = A
    M[:2, 2] = np.asarray(target, np.float64) - A @ np.asarray(position, np.float64)
    return M


def crop_similarity(img, position, scale, angle, size):
    """base_utils.py:646-655 transformation_crop (float32 2x3 matrix, cv2.warpAffine).  img=None:
    only the matrix (the caller warps on the device with g6d_warp_affine_u8)."""
    # the reference accumulates the 2x3 blocks in float32; keep the same rounding
    M = np.asarray([[1, 0, -position[0]], [0, 1, -position[1]]], np.float32)
    S = np.asarray([[scale, 0, 0], [0, scale, 0]], np.float32)
    Rm = np.asarray([[np.cos(angle), -np.sin(angle), 0], [np.sin(angle), np.cos(angle), 0]], np.float32)
    T = np.asarray([[1, 0, size / 2], [0, 1, size / 2]], np.float32)
    for nxt in (S, Rm, T):
        M = np.concatenate([nxt[:, :2] @ M[:, :2], (nxt[:, :2] @ M[:, 2] + nxt[:, 2])[:, None]], 1)
    return (cv2.warpAffine(img, M, (size, size), flags=cv2.INTER_LINEAR) if img is not None else None), M


# ------------------------------------------------------------------------------------------ device warps
WARP_JOB = np.dtype([('src', '<u8'), ('rows', '<i4'), ('cols', '<i4'), ('M', '<f8', (9,))])   # g6d_warp_job


def perspective_dst_to_src(H):
    """The matrix cv2.warpPerspective actually iterates with: cv::invert (LU) of H taken to double."""
    ok, inv = cv2.invert(np.ascontiguousarray(H, dtype=np.float64))
    return inv.reshape(9)


def affine_dst_to_src(M):
    """The 2x3 matrix cv2.warpAffine iterates with: OpenCV's closed-form inverse, evaluated in
    double in OpenCV's operation order (so that the fixed-point coordinates round identically)."""
    m = [float(v) for v in np.asarray(M, np.float64).reshape(6)]
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0], m[1], m[3], m[4] = a11, m[1] * -d, m[3] * -d, a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return np.asarray(m + [0.0, 0.0, 1.0])


def warp_source(t):
    """(device pointer, rows, cols) of a contiguous CUDA uint8 [rows, cols, 3] tensor: the source half of a
    g6d_warp_job.  Long-lived sources (the resident database images) are described once and reused."""
    if t.dim() != 3 or t.shape[2] != 3 or not t.is_contiguous() or not t.is_cuda or t.element_size() != 1:
        raise ValueError('pack_warp_jobs: sources must be contiguous CUDA uint8 [rows, cols, 3] tensors')
    return (t.data_ptr(), t.shape[0], t.shape[1])


def pack_warp_jobs(srcs, mats):
    """srcs: device uint8 tensors [rows, cols, 3] (contiguous) or their warp_source() triples; mats: dst->src
    matrices [9].  -> uint8 [n*88] host array, the g6d_warp_job records of include/gen6d_b200.h."""
    desc = [s if isinstance(s, tuple) else warp_source(s) for s in srcs]
    jobs = np.zeros(len(desc), WARP_JOB)
    if desc:
        ptr, rows, cols = zip(*desc)
        jobs['src'], jobs['rows'], jobs['cols'] = ptr, rows, cols
        jobs['M'] = np.asarray(mats, np.float64).reshape(len(desc), 9)
    return jobs.view(np.uint8)


# ------------------------------------------------------------------------------------------ view selection
def farthest_point_indices(points, count):
    """Farthest-point sampling seeded at the centroid (the centroid itself is not returned):
    base_utils.py:558-586 with init_center=True, index_model=True -> count-1 indices."""
    points = np.asarray(points)
    count = min(points.shape[0], count)
    cur = points.mean(0)
    dist = np.full(points.shape[0], 1e8)
    picked = []
    for _ in range(min(count - 1, points.shape[0] - 1)):
        dist = np.minimum(dist, np.linalg.norm(cur[None] - points, 2, 1))
        i = int(np.argmax(dist))
        picked.append(i)
        cur = points[i]
    return np.asarray(picked, dtype=np.int64)


def select_views_fps(database, ids, count):
    """database_utils.py:112-123 (random_fps=False): FPS over camera centres relative to the object."""
    center = database.object_center()
    cams = np.asarray([camera_center(database.get_pose(i)) - center for i in ids])
    return np.asarray(ids)[farthest_point_indices(cams, count + 1)]


def select_views_near_pose(database, center, ids, pose, count=6, even=False, even_count=128):
    """database_utils.py:125-139: optionally re-spread with FPS, then the `count` views whose
    viewing direction is closest (largest cosine) to that of `pose`."""
    unit = lambda v: v / np.linalg.norm(v, 2, -1, keepdims=True)
    if even and hasattr(database, 'even_subset'):
        ids, poses, dirs = database.even_subset(ids, even_count, center)     # pose-independent: cached per reference set
    else:
        ids = np.asarray(ids)
        poses = np.asarray([database.get_pose(i) for i in ids])
        if even:
            keep = farthest_point_indices(np.asarray([camera_center(p) for p in poses]), even_count + 1)
            ids, poses = ids[keep], poses[keep]
        dirs = unit(np.asarray([camera_center(p) for p in poses]) - center[None])
    q = unit(camera_center(pose) - center)
    return ids[np.argsort(-(dirs @ q))[:count]]


def _look_at_batch(cen_px, Ks):
    """Vectorised look_at_pixel: cen_px [n,2], Ks [n,3,3] -> (R [n,3,3], f [n])."""
    f = (Ks[:, 0, 0] + Ks[:, 1, 1]) / 2
    c = cen_px - Ks[:, :2, 2]
    xy = c / f[:, None]
    a, b = -np.arctan2(xy[:, 0], 1.0), np.arctan2(xy[:, 1], 1.0)
    ca, sa, cb, sb = np.cos(a), np.sin(a), np.cos(b), np.sin(b)
    # rx @ ry written out (every element is a single product: no rounding difference to the matrix product,
    # and no np.stack calls -- they were a third of refine_problem's time)
    R = np.zeros((len(a), 3, 3))
    R[:, 0, 0], R[:, 0, 2] = ca, sa
    R[:, 1, 0], R[:, 1, 1], R[:, 1, 2] = sb * sa, cb, -(sb * ca)
    R[:, 2, 0], R[:, 2, 1], R[:, 2, 2] = -(cb * sa), sb, cb * ca
    return R, np.sqrt(np.sum(c * c, 1) + f * f)


def _project_center_batch(center, poses, Ks):
    p = (poses[:, :, :3] @ center + poses[:, :, 3])[:, None, :] @ np.transpose(Ks, (0, 2, 1))
    p = p[:, 0]
    d = p[:, 2].copy()
    tiny = (np.abs(d) < 1e-4) & (np.abs(d) > 0)
    d[tiny] = 1e-4
    return p[:, :2] / d[:, None]


def reference_view_table(database, ids, size, margin):
    """The part of normalize_reference_views that does not depend on the alignment pose, batched over
    the views (float64): look-at rotation of every view, its product with the view's rotation, the
    focal length of the normalised crop, K^-1."""
    center = database.object_center().astype(np.float64)
    diameter = database.object_diameter()
    poses = np.stack([database.get_pose(i) for i in ids], 0).astype(np.float64)
    Ks = np.stack([database.get_K(i) for i in ids], 0).astype(np.float64)
    cen_px = _project_center_batch(center, poses, Ks)
    cams = -np.einsum('nji,nj->ni', poses[:, :, :3], poses[:, :, 3])
    dist = np.linalg.norm(cams - center[None], axis=1)
    R_look, f_look = _look_at_batch(cen_px, Ks)
    scale = size * (1 - margin) / diameter * dist / f_look
    return {'poses': poses, 'R_look': R_look, 'RlookR': R_look @ poses[:, :, :3], 'f': f_look * scale,
            'Kinv': np.linalg.inv(Ks)}


def _views_at_angle(tab, angle, size):
    """Second half of normalize_reference_views for table rows `tab` [n] and in-plane angles [n]:
    (K_new [n,3,3] f32, poses_new [n,3,4], Hs [n,3,3] f64).  Every view is independent of the others."""
    poses, n = tab['poses'], len(angle)
    ca, sa = np.cos(angle), np.sin(angle)
    Rz = np.zeros((n, 3, 3), np.float32)                       # reference builds R_z in float32
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = ca, -sa, sa, ca, 1
    R = Rz @ tab['R_look']
    K_new = np.zeros((n, 3, 3), np.float32)
    K_new[:, 0, 0] = K_new[:, 1, 1] = tab['f']
    K_new[:, 0, 2], K_new[:, 1, 2], K_new[:, 2, 2] = size / 2, size / 2, 1
    Hs = K_new @ R @ tab['Kinv']
    rect = R.astype(np.float32)
    poses_new = np.concatenate([rect @ poses[:, :, :3], rect @ poses[:, :, 3:]], 2)
    return K_new, poses_new, Hs


def normalize_reference_views(database, ids, size, margin, align_pose=None, align_K=None, warp=True):
    """database_utils.py:54-110 (rectify_rot=True, no extra rotations): every reference view is
    re-rendered as a look-at crop of the object at a common apparent size, with the in-plane
    orientation either 'object-up' (build time) or aligned to a given pose (refinement).
    Returns imgs [n,size,size,3] u8, Ks, poses, Hs (masks are not used on the inference path).
    The camera algebra is batched over the views (float64); only the warps loop.  Databases that
    offer `view_table` (NormalizedView) serve the pose-independent half from a per-object cache."""
    tab = database.view_table(ids, size, margin) if hasattr(database, 'view_table') else reference_view_table(database, ids, size, margin)
    poses, R_look = tab['poses'], tab['R_look']
    n = len(ids)
    if align_pose is not None:
        center = database.object_center().astype(np.float64)
        ap, aK = align_pose.astype(np.float64), align_K.astype(np.float64)
        Rq = look_at_point(ap, aK, center)[0] @ ap[:, :3]
        rel = Rq[None] @ np.transpose(tab['RlookR'], (0, 2, 1))
        # R = Rx(c) Ry(b) Rz(a)  =>  first row = [cos b cos a, -cos b sin a, sin b]
        angle = np.arctan2(-rel[:, 0, 1], rel[:, 0, 0])
    else:
        v = (poses[:, :, :3] @ database.object_vert().astype(np.float64))[:, :2]
        small = np.linalg.norm(v, axis=1) < 1e-5
        v[small] += 1e-5 * np.sign(v[small])
        angle = -np.arctan2(v[:, 1], v[:, 0]) - np.pi / 2
    K_new, poses_new, Hs = _views_at_angle(tab, angle, size)
    imgs = np.stack([cv2.warpPerspective(database.get_image(i), Hs[k], (size, size), flags=cv2.INTER_LINEAR)
                     for k, i in enumerate(ids)], 0) if warp else None   # warp=False: the caller warps on the device
    return imgs, K_new, poses_new, Hs


# ------------------------------------------------------------------------------------------ pose from detection + selection
def poses_from_similarity(positions, scales_r2q, angles_r2q, ref_poses, ref_Ks, que_Ks, center):
    """pose_from_similarity for n independent detections in one pass of stacked operations: positions [n,2],
    scales / angles [n], ref_poses [n,3,4], ref_Ks / que_Ks [n,3,3] -> poses [n,3,4] float64.  Slices are
    independent: a detection gets the same pose alone or inside a batch."""
    center = np.asarray(center, np.float64)
    ref_poses, ref_Ks, que_Ks = np.asarray(ref_poses), np.asarray(ref_Ks), np.asarray(que_Ks)
    n = len(ref_poses)
    ref_cen = _project_center_batch(center, ref_poses, ref_Ks)
    # query -> reference similarity (similarity_2d), then inverted (reference -> query)
    sc, ang = 1.0 / np.asarray(scales_r2q), -np.asarray(angles_r2q)
    c, s_ = np.cos(ang), np.sin(ang)
    M = np.zeros((n, 3, 3))
    M[:, 0, 0], M[:, 0, 1], M[:, 1, 0], M[:, 1, 1], M[:, 2, 2] = sc * c, sc * -s_, sc * s_, sc * c, 1
    M[:, :2, 2] = ref_cen - (M[:, :2, :2] @ np.asarray(positions, np.float64)[:, :, None])[:, :, 0]
    M_r2q = np.linalg.inv(M)
    que_cen = (M_r2q[:, :2, :2] @ ref_cen[:, :, None])[:, :, 0] + M_r2q[:, :2, 2]
    bearing = (np.linalg.inv(que_Ks) @ np.concatenate([que_cen, np.ones((n, 1))], 1)[:, :, None])[:, :, 0]
    bearing_xy = bearing[:, :2] / bearing[:, 2:3]
    scale = np.sqrt(np.linalg.det(M_r2q[:, :2, :2]))
    rotation = np.arctan2(M_r2q[:, 1, 0], M_r2q[:, 0, 0])
    que_f, ref_f = (que_Ks[:, 0, 0] + que_Ks[:, 1, 1]) / 2, (ref_Ks[:, 0, 0] + ref_Ks[:, 1, 1]) / 2
    que_f_ray = np.sqrt(que_f ** 2 + np.linalg.norm(bearing_xy * que_f[:, None], axis=1) ** 2)
    ref_dist = np.asarray([np.linalg.norm(camera_center(p) - center) for p in ref_poses])
    que_dist = ref_dist * que_f_ray / ref_f / scale
    ray = np.concatenate([bearing_xy, np.ones((n, 1))], 1)
    cen3d = ray / np.linalg.norm(ray, axis=1)[:, None] * que_dist[:, None]
    # look_at_rotation(bearing_xy).T @ (rot_z(rotation) @ R_ref)
    R_look = _look_at_batch(bearing_xy, np.broadcast_to(np.eye(3), (n, 3, 3)))[0]
    cr, sr = np.cos(rotation), np.sin(rotation)
    Rz = np.zeros((n, 3, 3))
    Rz[:, 0, 0], Rz[:, 0, 1], Rz[:, 1, 0], Rz[:, 1, 1], Rz[:, 2, 2] = cr, -sr, sr, cr, 1
    R = np.transpose(R_look, (0, 2, 1)) @ (Rz @ ref_poses[:, :, :3])
    return np.concatenate([R, (cen3d - R @ center)[:, :, None]], 2)


def pose_from_similarity(position, scale_r2q, angle_r2q, ref_pose, ref_K, que_K, center):
    """estimate_pose_from_similarity_transform_compose (pose_utils.py:104-111 -> :12-46): the 2-D
    similarity (detected position / scale, selected in-plane angle) that maps the reference crop
    into the query image fixes the object's bearing, depth and in-plane rotation."""
    return poses_from_similarity(np.asarray(position)[None], np.asarray(scale_r2q)[None], np.asarray(angle_r2q)[None],
                                 np.asarray(ref_pose)[None], np.asarray(ref_K)[None], np.asarray(que_K)[None], center)[0]


# ------------------------------------------------------------------------------------------ refinement bookkeeping
class NormalizedView:
    """The database seen through 'object in the unit sphere at the origin' coordinates
    (dataset/database.py:667-694 NormalizedDatabase; normalize_pose :400-404)."""

    def __init__(self, database):
        self.db = database
        self.scale = 2.0 / database.object_diameter()
        self.offset = -self.scale * database.object_center()
        self._poses = {}        # normalised poses, cached (pose-independent of the query)
        self._even = {}         # FPS-resampled reference subsets per (ids, count)
        self._tables = {}       # pose-independent halves of normalize_reference_views per (size, margin)

    def normalize_pose(self, pose):
        R, t = pose[:3, :3], pose[:3, 3]
        return np.concatenate([R, (R @ -self.offset + self.scale * t)[:, None]], -1).astype(np.float32)

    def denormalize_pose(self, pose):
        R, t = pose[:3, :3], pose[:3, 3]
        return np.concatenate([R, (R @ self.offset / self.scale + t / self.scale)[:, None]], -1).astype(np.float32)

    def normalize_poses(self, poses):
        """normalize_pose over a stack [n,3,4] (each pose independent of the others)."""
        R, t = poses[:, :, :3], poses[:, :, 3]
        return np.concatenate([R, (R @ -self.offset + self.scale * t)[:, :, None]], -1).astype(np.float32)

    def denormalize_poses(self, poses):
        R, t = poses[:, :, :3], poses[:, :, 3]
        return np.concatenate([R, (R @ self.offset / self.scale + t / self.scale)[:, :, None]], -1).astype(np.float32)

    def get_pose(self, i):
        if i not in self._poses:
            self._poses[i] = self.normalize_pose(self.db.get_pose(i))
        return self._poses[i]

    def even_subset(self, ids, count, center):
        """The FPS re-spread of database_utils.py:129-134 and the unit viewing directions depend only
        on the reference set, not on the query pose: computed once per (ids, count) instead of once
        per refinement iteration."""
        key = (tuple(ids), count)
        if key not in self._even:
            ids_a = np.asarray(ids)
            poses = np.asarray([self.get_pose(i) for i in ids_a])
            keep = farthest_point_indices(np.asarray([camera_center(p) for p in poses]), count + 1)
            poses = poses[keep]
            cams = np.asarray([camera_center(p) for p in poses]) - np.asarray(center)[None]
            self._even[key] = (ids_a[keep], poses, cams / np.linalg.norm(cams, 2, -1, keepdims=True))
        return self._even[key]

    def view_table(self, ids, size, margin, prefill=()):
        """Rows of reference_view_table for `ids`, from a table built once per (size, margin) over the
        reference set `prefill` (per-view quantities do not depend on which other views are in the batch)."""
        key = (size, margin)
        tab, index = self._tables.get(key, (None, {}))
        missing = [i for i in ids if i not in index]
        if missing:
            all_ids = list(index) + [i for i in dict.fromkeys(list(prefill) + missing) if i not in index]
            tab = reference_view_table(self, all_ids, size, margin)
            index = {i: k for k, i in enumerate(all_ids)}
            self._tables[key] = (tab, index)
        rows = np.asarray([index[i] for i in ids])
        return {k: v[rows] for k, v in tab.items()}

    def get_K(self, i):
        return self.db.get_K(i)

    def get_image(self, i):
        return self.db.get_image(i)

    def object_center(self):
        return np.zeros(3, np.float32)

    def object_diameter(self):
        return 2.0

    def object_vert(self):
        return self.db.object_vert()


_VIEW_CACHE = {}


def _normalized_view(database):
    key = id(database)
    if key not in _VIEW_CACHE or _VIEW_CACHE[key].db is not database:
        _VIEW_CACHE[key] = NormalizedView(database)
    return _VIEW_CACHE[key]


def refine_problems(database, ref_ids, que_Ks, in_poses, size=128, ref_num=6, ref_even=False, margin=0.05):
    """refine_problem(warp=False) for qn independent frames in one pass of stacked numpy operations (the
    per-frame version costs ~60 small numpy calls; a batch of 10 frames three times per prediction made the
    host the bottleneck of the batched stages).  Every frame's slice is computed independently of the others,
    so a frame gets the same numbers alone or inside any batch: refine_problem is the qn = 1 case.
    Returns stacked arrays: 'que_K' [qn,3,3] f32, 'que_pose' [qn,3,4] f32, 'pose_rect' [qn,3,4] f32, 'que_H'
    [qn,3,3], 'ref_ids' [qn,ref_num], 'ref_Ks' [qn,ref_num,3,3] f32, 'ref_poses' [qn,ref_num,3,4] f32, 'ref_Hs'
    [qn,ref_num,3,3], plus 'view' / 'center'."""
    view = _normalized_view(database)
    Ks, pose_n = np.asarray(que_Ks), view.normalize_poses(np.asarray(in_poses))
    qn = len(pose_n)
    center = view.object_center()
    c64 = center.astype(np.float64)
    R, t = pose_n[:, :, :3], pose_n[:, :, 3]
    # look at the projected object centre (pose_utils.py:52-58), zoom so that the object fills the crop
    cen_px = _project_center_batch(c64, pose_n, Ks)
    R_look, f_look = _look_at_batch(cen_px, Ks)
    # float32 reductions: per frame with the operations the single-frame path has always used (a stacked matmul /
    # axis-norm may sum in another order: 1 ulp of the crop's focal length)
    dist = np.asarray([np.linalg.norm(camera_center(p) - center) for p in pose_n])
    scale = size * (1 - margin) / view.object_diameter() * dist / f_look
    # look_at_crop with angle 0 (database_utils.py:8-25)
    K_warp = np.zeros((qn, 3, 3), np.float32)
    K_warp[:, 0, 0] = K_warp[:, 1, 1] = f_look * scale
    K_warp[:, 0, 2], K_warp[:, 1, 2], K_warp[:, 2, 2] = size / 2, size / 2, 1
    que_H = K_warp @ R_look @ np.linalg.inv(Ks)
    rect = np.concatenate([R_look, np.zeros((qn, 3, 1))], 2).astype(np.float32)
    pose_warp = np.concatenate([rect[:, :, :3] @ R, rect[:, :, :3] @ pose_n[:, :, 3:] + rect[:, :, 3:]], 2)
    # the ref_num reference views closest in viewing direction (database_utils.py:125-139)
    if ref_even and hasattr(view, 'even_subset'):
        ids_e, _, dirs = view.even_subset(ref_ids, min(128, len(ref_ids)), center)
        cam_w = -(np.transpose(pose_warp[:, :, :3], (0, 2, 1)) @ pose_warp[:, :, 3:])[:, :, 0] - center[None]
        qdir = cam_w / np.linalg.norm(cam_w, 2, -1, keepdims=True)
        order = np.asarray([np.argsort(-(dirs @ qdir[i]))[:ref_num] for i in range(qn)])
        ids = ids_e[order]
    else:
        ids = np.stack([select_views_near_pose(view, center, ref_ids, pose_warp[i], ref_num, ref_even, min(128, len(ref_ids)))
                        for i in range(qn)], 0)
    # reference views re-rendered with their in-plane orientation aligned to the query's (database_utils.py:54-110)
    tab = view.view_table(list(ids.reshape(-1)), size, margin, prefill=ref_ids)
    ap, aK = pose_warp.astype(np.float64), K_warp.astype(np.float64)
    Rq = _look_at_batch(_project_center_batch(c64, ap, aK), aK)[0] @ ap[:, :, :3]
    rel = np.repeat(Rq, ref_num, 0) @ np.transpose(tab['RlookR'], (0, 2, 1))
    # R = Rx(c) Ry(b) Rz(a)  =>  first row = [cos b cos a, -cos b sin a, sin b]
    ref_Ks, ref_poses, ref_Hs = _views_at_angle(tab, np.arctan2(-rel[:, 0, 1], rel[:, 0, 0]), size)
    return {'view': view, 'center': center, 'que_K': K_warp, 'que_pose': pose_warp.astype(np.float32), 'pose_rect': rect,
            'que_H': que_H, 'ref_ids': ids, 'ref_Ks': ref_Ks.reshape(qn, ref_num, 3, 3).astype(np.float32),
            'ref_poses': ref_poses.reshape(qn, ref_num, 3, 4).astype(np.float32), 'ref_Hs': ref_Hs.reshape(qn, ref_num, 3, 3)}


def refine_problem(database, ref_ids, que_img, que_K, in_pose, size=128, ref_num=6, ref_even=False, margin=0.05,
                   warp=True):
    """Everything refiner.py:285-325 prepares on the host for one refinement step: the query
    look-at crop at the input pose and the `ref_num` nearest reference views re-rendered with
    their in-plane orientation aligned to it.  warp=False skips the OpenCV warps (que_img may be
    None) and only returns their homographies 'que_H' / 'ref_Hs' for g6d_warp_perspective_u8."""
    b = refine_problems(database, ref_ids, [que_K], [in_pose], size, ref_num, ref_even, margin)
    prob = {k: (v if k in ('view', 'center') else v[0]) for k, v in b.items()}
    prob['que_img'] = prob['ref_imgs'] = None
    if warp:
        prob['que_img'] = cv2.warpPerspective(que_img, prob['que_H'], (size, size), flags=cv2.INTER_LINEAR)
        prob['ref_imgs'] = np.stack([cv2.warpPerspective(prob['view'].get_image(i), H, (size, size), flags=cv2.INTER_LINEAR)
                                     for i, H in zip(prob['ref_ids'], prob['ref_Hs'])], 0)
    return prob


def _quat_to_matrix_batch(q):
    """quat_to_matrix over [n,4] (w, x, y, z), float64."""
    q = np.asarray(q, np.float64)
    w, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    n = w * w + x * x + y * y + z * z
    ok = n >= np.finfo(np.float64).eps
    sc = 2.0 / np.where(ok, n, 1.0)
    m = np.stack([np.stack([1 - sc * (y * y + z * z), sc * (x * y - w * z), sc * (x * z + w * y)], -1),
                  np.stack([sc * (x * y + w * z), 1 - sc * (x * x + z * z), sc * (y * z - w * x)], -1),
                  np.stack([sc * (x * z - w * y), sc * (y * z + w * x), 1 - sc * (x * x + y * y)], -1)], 1)
    m[~ok] = np.eye(3)
    return m


def apply_refinements(probs, quats, offsets, scales):
    """apply_refinement over the stacked problems of refine_problems: quats [qn,4], offsets [qn,2], scales [qn]
    -> poses [qn,3,4] float32.  Frame slices are independent; apply_refinement is the qn = 1 case."""
    pose_in, K = probs['que_pose'].astype(np.float64), probs['que_K'].astype(np.float64)
    center = probs['center'].astype(np.float64)
    Rin = pose_in[:, :, :3]
    cen_in = Rin @ center + pose_in[:, :, 3]
    A = np.asarray(scales, np.float64)[:, None, None] * _quat_to_matrix_batch(quats)
    off = np.asarray(offsets, np.float64)
    cen_que = cen_in + np.concatenate([off, np.zeros((len(off), 1))], 1)
    A_cen = (A @ cen_in[:, :, None])[:, :, 0]
    sim_t = cen_que - A_cen
    # similarity -> rigid: keep the rotation part, move the centre along its new ray to the depth
    # implied by the scale change
    U, S, Vt = np.linalg.svd(A)
    Rdelta = U @ Vt
    f = (K[:, 0, 0] + K[:, 1, 1]) / 2
    depth = cen_in[:, 2] / np.mean(np.abs(S), axis=1) * f / f
    cen_sim = A_cen + sim_t
    cen_new = cen_sim / cen_sim[:, 2:3] * depth[:, None]
    Rn = Rdelta @ Rin
    pose = np.concatenate([Rn, (cen_new - Rn @ center)[:, :, None]], 2)
    # undo the look-at rectification: then apply the inverse of pose_rect
    rect = probs['pose_rect']
    Rt = np.transpose(rect[:, :, :3], (0, 2, 1))
    inv = np.concatenate([Rt, -Rt @ rect[:, :, 3:]], 2)
    pose = np.concatenate([inv[:, :, :3] @ pose[:, :, :3], inv[:, :, :3] @ pose[:, :, 3:] + inv[:, :, 3:]], 2)
    return probs['view'].denormalize_poses(pose)


def apply_refinement(prob, quat, offset, scale):
    """refiner.py:333-340: (scale, quaternion, 2-D offset) -> similarity transform about the object
    centre -> rigid pose at the matching depth -> undo the look-at rectification -> undo the
    unit-sphere normalisation (pose_utils.py:217-244)."""
    probs = {k: (v if k in ('view', 'center') else v[None]) for k, v in prob.items() if k in
             ('view', 'center', 'que_pose', 'que_K', 'pose_rect')}
    return apply_refinements(probs, np.asarray(quat, np.float64)[None], np.asarray(offset, np.float64)[None],
                             [float(np.asarray(scale).reshape(-1)[0])])[0]
