"""TEST INFRASTRUCTURE ONLY - never imported by the product path (gen6d_b200/).

CPU restatement (numpy) of the two OpenCV image warps that sit between the network stages of the
Gen6D estimator, for uint8 3-channel images, INTER_LINEAR, BORDER_CONSTANT(0):

  * cv2.warpPerspective - look_at_crop, /root/reference/utils/database_utils.py:8-25, called by
    network/refiner.py:300 (query crop) and through normalize_reference_views
    (database_utils.py:95) by network/refiner.py:309-313 (reference crops);
  * cv2.warpAffine - transformation_crop, /root/reference/utils/base_utils.py:646-655, called by
    estimator.py:184 (detection crop).

The algorithm lives in a third-party dependency that is not in /root/reference: OpenCV
(opencv-python 4.13.0 in this image; the reference's requirements do not pin a version).  What is
restated here is OpenCV's published 8-bit fixed-point path (modules/imgproc/src/imgwarp.cpp:
WarpPerspectiveInvoker / WarpAffineInvoker feeding remapBilinear with INTER_BITS = 5,
INTER_REMAP_COEF_BITS = 15, AB_BITS = 10).  Pinning: tests/test_warp_oracle.py checks this file
bit-for-bit against cv2 itself, which is installed here and on the GPU box, on seeded random
homographies / similarities including out-of-image borders.

Everything is evaluated in IEEE double without fused multiply-add, in OpenCV's operation order,
because the rounding of the scaled source coordinate to an integer decides which taps are read.
"""
import numpy as np

INTER_BITS = 5
TAB = 1 << INTER_BITS            # 32 sub-pixel positions per axis
COEF_BITS = 15                   # weights sum to 1 << 15
AB_BITS = 10


def _sat_round(v):
    """cv::saturate_cast<int>(double): clamp, then round half to even (cvRound)."""
    return np.rint(np.clip(v, -2147483648.0, 2147483647.0)).astype(np.int64)


def _fixed_bilinear(src, X, Y):
    """remapBilinear for 8UC3 with a zero constant border.  X, Y: source coordinates * 32, int64 [h,w]."""
    rows, cols = src.shape[:2]
    sx = np.clip(X >> INTER_BITS, -32768, 32767)
    sy = np.clip(Y >> INTER_BITS, -32768, 32767)
    a, b = X & (TAB - 1), Y & (TAB - 1)
    acc = np.zeros(X.shape + (src.shape[2],), np.int64)
    for dy, dx, wgt in ((0, 0, (TAB - b) * (TAB - a)), (0, 1, (TAB - b) * a), (1, 0, b * (TAB - a)), (1, 1, b * a)):
        yy, xx = sy + dy, sx + dx
        ok = (yy >= 0) & (yy < rows) & (xx >= 0) & (xx < cols)
        tap = src[np.clip(yy, 0, rows - 1), np.clip(xx, 0, cols - 1)].astype(np.int64)
        acc += np.where(ok, wgt * (1 << (COEF_BITS - 2 * INTER_BITS)), 0)[..., None] * tap
    return ((acc + (1 << (COEF_BITS - 1))) >> COEF_BITS).astype(np.uint8)


def invert_3x3(H):
    """cv::invert(DECOMP_LU) of a 3x3 double matrix: OpenCV special-cases n = 3 with the adjugate
    formula scaled by 1/det (modules/core/src/lapack.cpp)."""
    s = np.asarray(H, np.float64)
    d = (s[0, 0] * (s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1]) - s[0, 1] * (s[1, 0] * s[2, 2] - s[1, 2] * s[2, 0])
         + s[0, 2] * (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0]))
    d = 1.0 / d
    t = np.empty((3, 3))
    t[0, 0] = (s[1, 1] * s[2, 2] - s[1, 2] * s[2, 1]) * d
    t[0, 1] = (s[0, 2] * s[2, 1] - s[0, 1] * s[2, 2]) * d
    t[0, 2] = (s[0, 1] * s[1, 2] - s[0, 2] * s[1, 1]) * d
    t[1, 0] = (s[1, 2] * s[2, 0] - s[1, 0] * s[2, 2]) * d
    t[1, 1] = (s[0, 0] * s[2, 2] - s[0, 2] * s[2, 0]) * d
    t[1, 2] = (s[0, 2] * s[1, 0] - s[0, 0] * s[1, 2]) * d
    t[2, 0] = (s[1, 0] * s[2, 1] - s[1, 1] * s[2, 0]) * d
    t[2, 1] = (s[0, 1] * s[2, 0] - s[0, 0] * s[2, 1]) * d
    t[2, 2] = (s[0, 0] * s[1, 1] - s[0, 1] * s[1, 0]) * d
    return t


def invert_2x3(M):
    """The closed-form inverse cv::warpAffine applies to its matrix (imgwarp.cpp, !WARP_INVERSE_MAP)."""
    m = [float(v) for v in np.asarray(M, np.float64).reshape(6)]
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0], m[1], m[3], m[4] = a11, m[1] * -d, m[3] * -d, a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return np.asarray(m)


def warp_perspective_u8(src, H, dsize):
    """cv2.warpPerspective(src, H, dsize=(w, h), flags=cv2.INTER_LINEAR)."""
    w, h = dsize
    M = invert_3x3(H).reshape(9)
    bh0 = min(16, h)
    bw0 = min(1024 // bh0, w)                     # column-block width of WarpPerspectiveInvoker
    x = np.arange(w)
    xb, x1 = ((x // bw0) * bw0).astype(np.float64)[None, :], (x % bw0).astype(np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    X0 = M[0] * xb + M[1] * y + M[2]
    Y0 = M[3] * xb + M[4] * y + M[5]
    W0 = M[6] * xb + M[7] * y + M[8]
    W = W0 + M[6] * x1
    with np.errstate(divide='ignore'):
        W = np.where(W != 0, TAB / np.where(W != 0, W, 1.0), 0.0)
    X = _sat_round((X0 + M[0] * x1) * W)
    Y = _sat_round((Y0 + M[3] * x1) * W)
    return _fixed_bilinear(src, X, Y)


def warp_affine_u8(src, M, dsize):
    """cv2.warpAffine(src, M, dsize=(w, h), flags=cv2.INTER_LINEAR)."""
    w, h = dsize
    m = invert_2x3(M)
    scale = float(1 << AB_BITS)
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    adelta, bdelta = _sat_round(m[0] * x * scale), _sat_round(m[3] * x * scale)
    round_delta = (1 << AB_BITS) // TAB // 2
    X0 = _sat_round((m[1] * y + m[2]) * scale) + round_delta
    Y0 = _sat_round((m[4] * y + m[5]) * scale) + round_delta
    return _fixed_bilinear(src, (X0 + adelta) >> (AB_BITS - INTER_BITS), (Y0 + bdelta) >> (AB_BITS - INTER_BITS))
