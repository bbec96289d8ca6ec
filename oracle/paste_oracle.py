"""oracle/paste_oracle.py -- CPU checker for the paste-back kernels.  TEST INFRASTRUCTURE ONLY.

The oracle for resize + paste is OpenCV itself (lipreal.py:211-214 calls cv2.resize / ndarray
assignment; cv2 is present on the GPU box), plus a numpy restatement of OpenCV's fixed-point
INTER_LINEAR used to document the arithmetic the CUDA kernel follows
(opencv/modules/imgproc/src/resize.cpp, HResizeLinear / VResizeLinear for uchar)."""
import copy

import numpy as np


def paste_cv2(frame, face_u8, bbox):
    """lipreal.py:207-214 verbatim semantics"""
    import cv2
    y1, y2, x1, x2 = bbox
    combine = copy.deepcopy(frame)
    combine[y1:y2, x1:x2] = cv2.resize(face_u8.astype(np.uint8), (x2 - x1, y2 - y1))
    return combine


def _coef(dsize, ssize, vertical=False):
    """x axis: the fractional offset is zeroed when the tap falls off either end (xofs/alpha set-up);
    y axis: OpenCV keeps the fraction and clamps the two ROW INDICES instead (resizeGeneric_Invoker
    `clip(sy0 - ksize2 + 1 + k, 0, ssize.height)`), which rounds differently on the border rows."""
    scale = float(ssize) / dsize
    d = np.arange(dsize)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo], s[lo] = 0, 0
        hi = s >= ssize - 1
        f[hi], s[hi] = 0, ssize - 1
    a0 = np.clip(np.rint((np.float32(1) - f) * np.float32(2048)), -32768, 32767).astype(np.int64)
    a1 = np.clip(np.rint(f * np.float32(2048)), -32768, 32767).astype(np.int64)
    return np.clip(s, 0, ssize - 1), np.clip(s + 1, 0, ssize - 1), a0, a1


def resize_linear_u8(img, dw, dh):
    """numpy restatement of cv2.resize(img_u8, (dw, dh), interpolation=INTER_LINEAR)"""
    S_h, S_w = img.shape[:2]
    if S_w == 2 * dw and S_h == 2 * dh:   # INTER_AREA fast path
        i = img.astype(np.int64)
        return ((i[0::2, 0::2] + i[0::2, 1::2] + i[1::2, 0::2] + i[1::2, 1::2] + 2) >> 2).astype(np.uint8)
    sx0, sx1, ax0, ax1 = _coef(dw, S_w)
    sy0, sy1, by0, by1 = _coef(dh, S_h, vertical=True)
    i = img.astype(np.int64)
    rows = i[:, sx0] * ax0[None, :, None] + i[:, sx1] * ax1[None, :, None]       # [S_h, dw, 3]
    r0, r1 = rows[sy0], rows[sy1]
    v = (((by0[:, None, None] * (r0 >> 4)) >> 16) + ((by1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(v, 0, 255).astype(np.uint8)


def blend_cv2(frame, face_u8, bbox, mask_bgr, crop_box):
    """musereal.py:240-248 + musetalk/utils/blending.py:103-125 verbatim semantics (bbox = (x1, y1, x2, y2),
    crop_box = (x_s, y_s, x_e, y_e)); OpenCV itself is the oracle"""
    import cv2
    body = copy.deepcopy(frame)
    x, y, x1, y1 = bbox
    face = cv2.resize(face_u8.astype(np.uint8), (x1 - x, y1 - y))
    x_s, y_s, x_e, y_e = crop_box
    face_large = copy.deepcopy(body[y_s:y_e, x_s:x_e])
    face_large[y - y_s:y1 - y_s, x - x_s:x1 - x_s] = face
    mask_image = cv2.cvtColor(mask_bgr, cv2.COLOR_BGR2GRAY)
    mask_image = (mask_image / 255).astype(np.float32)
    body[y_s:y_e, x_s:x_e] = cv2.blendLinear(face_large, body[y_s:y_e, x_s:x_e], mask_image, 1 - mask_image)
    return body


def blend_numpy(frame, face_u8, bbox, mask_bgr, crop_box):
    """the arithmetic the CUDA kernel restates: 15-bit fixed-point BGR2GRAY, float64 `/ 255` narrowed to fp32, fp32 blend
    with separately rounded products, round-half-even"""
    body = frame.copy()
    x, y, x1, y1 = bbox
    x_s, y_s, x_e, y_e = crop_box
    fl = body[y_s:y_e, x_s:x_e].copy()
    fl[y - y_s:y1 - y_s, x - x_s:x1 - x_s] = resize_linear_u8(face_u8, x1 - x, y1 - y)
    m = mask_bgr.astype(np.int64)
    gray = (m[..., 0] * 3735 + m[..., 1] * 19235 + m[..., 2] * 9798 + (1 << 14)) >> 15
    w1 = (gray / 255).astype(np.float32)[..., None]
    w2 = np.float32(1) - w1
    den = (w1 + w2) + np.float32(1e-5)
    num = fl.astype(np.float32) * w1 + body[y_s:y_e, x_s:x_e].astype(np.float32) * w2
    body[y_s:y_e, x_s:x_e] = np.clip(np.rint(num / den), 0, 255).astype(np.uint8)
    return body
