"""ORACLE (test infrastructure, not product code) -- image pyramid pre-processing, restating
``lib/utils/test_utils.py:8-46``, ``lib/utils/blob.py:16-32`` and ``lib/test.py:30-38,131-137``.

The reference resizes with OpenCV (``cv2.resize(im, None, None, fx=s, fy=s, INTER_LINEAR)``), a
third-party dependency that is not under /root/reference (``requirements.txt:1``, unpinned).  This
image ships opencv-python-headless 4.13; ``resize_linear_cv2`` calls it and ``resize_linear`` restates
its algorithm in NumPy (bilinear, half-pixel centres, border-clamped taps; see ``resize_linear``):
tests require the two to agree to <=1 float32 ulp, which pins the restatement the CUDA kernel
follows.

Note the dtype chain in ``_get_image_blob``: ``im.astype(float32) - np.array(PIXEL_MEANS)`` is a
float64 array (PIXEL_MEANS is a python-float list), so the bilinear blend runs in float64 and only ``im_list_to_blob`` rounds to float32.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
PIXEL_MEANS = np.array([[[102.9801, 115.9465, 122.7717]]])       # configs/default.toml:6
MAX_RESOLUTION = 16                                              # configs/default.toml:4
TEST_SCALES = (100, 300, 600, 1000, 1400)                        # configs/default.toml:137
PYRAMID_BASE_SIZE = (800, 1200)                                  # configs/default.toml:135


def compute_scaling_factor(im_shape, target_size, max_size):
    """``test_utils.py:8-26``."""
    im_size_min = np.min(im_shape[0:2])
    im_size_max = np.max(im_shape[0:2])
    im_scale = float(target_size) / float(im_size_min)
    if np.round(im_scale * im_size_max) > max_size:
        im_scale = float(max_size) / float(im_size_max)
    return im_scale


def pyramid_scales(im_shape, scales=TEST_SCALES, base=PYRAMID_BASE_SIZE):
    """``lib/test.py:131-137``."""
    base_scale = compute_scaling_factor(im_shape, base[0], base[1])
    return [float(s) / base[0] * base_scale for s in scales]


def _linear_coeffs(dst, src, inv_scale):
    """Source tap + fraction per destination index: ``f = (d+0.5)*(1/inv_scale) - 0.5`` in double
    (scale is 1/fx, NOT src/dst, because only fx/fy are given -- the two differ by up to 0.18 px at
    1024->1867), taps clamped to the image with the fraction zeroed at the borders."""
    d = np.arange(dst, dtype=np.float64)
    f = (d + 0.5) * (1.0 / inv_scale) - 0.5
    s = np.floor(f).astype(np.int64)
    a = f - s
    lo = s < 0
    a = np.where(lo, 0.0, a); s = np.where(lo, 0, s)
    hi = s >= src - 1
    a = np.where(hi, 0.0, a); s = np.where(hi, src - 1, s)
    return s, np.minimum(s + 1, src - 1), a


def resize_linear(img, fx, fy):
    """NumPy restatement of ``cv2.resize(img, None, None, fx, fy, INTER_LINEAR)`` for float64 HWC
    input as opencv 4.13 computes it: output size ``rint(src*f)``, double coefficients, lerp form
    ``S0 + (S1-S0)*a``, horizontal then vertical.  Agrees with cv2 4.13 to <1e-8 in float64 and to
    <=1 float32 ulp after the blob cast (tests/test_oracle_preprocess.py).  OpenCV builds of the
    reference's era (2018, 3.4/4.0) used float32 coefficients instead (up to 0.012 grey levels
    away); the reference pins neither, so the library actually installed is the oracle."""
    img = np.asarray(img, dtype=np.float64)
    h, w = img.shape[:2]
    dw = int(np.rint(w * fx))              # dsize = saturate_cast<int>(src * fx): round half to even
    dh = int(np.rint(h * fy))
    sx, sx1, ax = _linear_coeffs(dw, w, fx)
    sy, sy1, ay = _linear_coeffs(dh, h, fy)
    t = img[:, sx] + (img[:, sx1] - img[:, sx]) * ax[None, :, None]
    return t[sy] + (t[sy1] - t[sy]) * ay[:, None, None]


def resize_linear_cv2(img, fx, fy):
    import cv2
    return cv2.resize(img, None, None, fx=fx, fy=fy, interpolation=cv2.INTER_LINEAR)


def get_image_blobs(im, scales, use_cv2=True):
    """``test_utils.py:29-46`` + ``blob.py:16-32``: list of (1,3,h,w) float32 blobs."""
    im_copy = im.astype(F32, copy=True) - PIXEL_MEANS                   # float64
    blobs = []
    for s in scales:
        if s == 1.0:
            r = im_copy
        else:
            r = resize_linear_cv2(im_copy, s, s) if use_cv2 else resize_linear(im_copy, s, s)
        blob = np.zeros((1, r.shape[0], r.shape[1], 3), dtype=F32)
        blob[0] = r
        blobs.append(np.ascontiguousarray(blob.transpose(0, 3, 1, 2)))
    return blobs


def pad_to_multiple(data, mult=MAX_RESOLUTION):
    """``lib/test.py:30-38``: im_info from the unpadded blob, zero-pad bottom/right."""
    h, w = data.shape[2:]
    nh = int(np.ceil(1.0 * h / mult) * mult)
    nw = int(np.ceil(1.0 * w / mult) * mult)
    return np.pad(data, ((0, 0), (0, 0), (0, nh - h), (0, nw - w)), "constant")
