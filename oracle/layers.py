"""ORACLE (test infrastructure, not product code) -- fp32 CPU restatement of the Caffe layers
the deploy nets use.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package.

Every function follows the cited reference file and computes in float32 NCHW like Caffe's CPU
mode.  ``conv`` has two engines: ``'sgemm'`` is the literal reference algorithm (im2col into a
column buffer + one BLAS sgemm per image, ``base_conv_layer.cpp:255-279``; NumPy dispatches the
matmul to OpenBLAS, the BLAS family ``caffe/Makefile.config:50`` selects) and ``'torch'`` is a
faster best-effort CPU conv (oneDNN) used only where noted.

Pinned against the reference's own known-answer tests in tests/test_oracle_kat.py
(``caffe/src/caffe/test/test_{convolution,pooling,deconvolution,softmax,concat,reshape}_layer.cpp``).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def conv_out_size(size, k, p, s, d):
    """``conv_layer.cpp:8-28`` compute_output_shape."""
    return (size + 2 * p - (d * (k - 1) + 1)) // s + 1


def im2col(x, kh, kw, ph, pw, sh, sw, dh, dw):
    """``util/im2col.cpp:19-55``: (C,H,W) -> (C*kh*kw, Ho*Wo), rows ordered (c, kr, kc),
    out-of-image taps read 0."""
    c, h, w = x.shape
    ho = conv_out_size(h, kh, ph, sh, dh)
    wo = conv_out_size(w, kw, pw, sw, dw)
    xp = np.zeros((c, h + 2 * ph, w + 2 * pw), dtype=x.dtype)
    xp[:, ph:ph + h, pw:pw + w] = x
    s0, s1, s2 = xp.strides
    view = np.lib.stride_tricks.as_strided(
        xp, shape=(c, kh, kw, ho, wo), strides=(s0, s1 * dh, s2 * dw, s1 * sh, s2 * sw), writeable=False)
    return np.ascontiguousarray(view).reshape(c * kh * kw, ho * wo), ho, wo


def conv(x, w, b=None, pad=(0, 0), stride=(1, 1), dilation=(1, 1), group=1, engine="sgemm",
         max_col_bytes=1 << 30):
    """Convolution forward, ``base_conv_layer.cpp:255-279`` forward_cpu_gemm + forward_cpu_bias:
    per image, per group: ``Y = W[Co,K] @ col[K,HoWo]`` then ``Y += b * 1^T``."""
    x = np.ascontiguousarray(x, dtype=F32)
    w = np.ascontiguousarray(w, dtype=F32)
    n, c, h, wd = x.shape
    co, cg, kh, kw = w.shape
    assert c == cg * group and co % group == 0, "channel/group mismatch"
    if engine == "torch":
        import torch
        with torch.no_grad():
            y = torch.nn.functional.conv2d(torch.from_numpy(x), torch.from_numpy(w),
                                           None if b is None else torch.from_numpy(np.asarray(b, F32)),
                                           stride=stride, padding=pad, dilation=dilation, groups=group)
        return y.numpy()
    ho = conv_out_size(h, kh, pad[0], stride[0], dilation[0])
    wo = conv_out_size(wd, kw, pad[1], stride[1], dilation[1])
    y = np.empty((n, co, ho, wo), dtype=F32)
    cog = co // group
    k = cg * kh * kw
    is_1x1 = kh == 1 and kw == 1 and pad == (0, 0) and stride == (1, 1)      # base_conv_layer.cpp:108-115
    # the column buffer of a 1408^2 conv1_2 is 4.5 GB; build it in row bands of the output so the
    # oracle stays within RAM (each output element is still one length-K sgemm dot product)
    rows_per_band = max(1, int(max_col_bytes // max(1, 4 * k * wo)))
    for i in range(n):
        for g in range(group):
            xg = x[i, g * cg:(g + 1) * cg]
            wg = w[g * cog:(g + 1) * cog].reshape(cog, k)
            if is_1x1:
                yg = wg @ xg.reshape(cg, h * wd)
            elif rows_per_band >= ho:
                col, _, _ = im2col(xg, kh, kw, pad[0], pad[1], stride[0], stride[1], dilation[0], dilation[1])
                yg = wg @ col
            else:
                yg = np.empty((cog, ho * wo), dtype=F32)
                xp = np.zeros((cg, h + 2 * pad[0], wd + 2 * pad[1]), dtype=F32)
                xp[:, pad[0]:pad[0] + h, pad[1]:pad[1] + wd] = xg
                for r0 in range(0, ho, rows_per_band):
                    r1 = min(ho, r0 + rows_per_band)
                    in0 = r0 * stride[0]
                    in1 = (r1 - 1) * stride[0] + dilation[0] * (kh - 1) + 1
                    col, bh, bw = im2col(xp[:, in0:in1], kh, kw, 0, 0, stride[0], stride[1],
                                         dilation[0], dilation[1])
                    assert bh == r1 - r0 and bw == wo
                    yg[:, r0 * wo:r1 * wo] = wg @ col
            y[i, g * cog:(g + 1) * cog] = yg.reshape(cog, ho, wo)
        if b is not None:
            y[i] += np.asarray(b, F32)[:, None, None]
    return y


def conv_naive(x, w, b=None, pad=(0, 0), stride=(1, 1), dilation=(1, 1), group=1):
    """The reference test-suite's 7-loop convolution (``test_convolution_layer.cpp:21-139``
    caffe_conv), used to pin ``conv`` itself.  Small inputs only."""
    n, c, h, wd = x.shape
    co, cg, kh, kw = w.shape
    ho = conv_out_size(h, kh, pad[0], stride[0], dilation[0])
    wo = conv_out_size(wd, kw, pad[1], stride[1], dilation[1])
    y = np.zeros((n, co, ho, wo), dtype=F32)
    og, kg = co // group, c // group
    for i in range(n):
        for g in range(group):
            for o in range(og):
                for kc in range(kg):
                    for yy in range(ho):
                        for xx in range(wo):
                            acc = y[i, g * og + o, yy, xx]
                            for p in range(kh):
                                for q in range(kw):
                                    iy = yy * stride[0] - pad[0] + p * dilation[0]
                                    ix = xx * stride[1] - pad[1] + q * dilation[1]
                                    if 0 <= iy < h and 0 <= ix < wd:
                                        acc = F32(acc + x[i, g * kg + kc, iy, ix] * w[g * og + o, kc, p, q])
                            y[i, g * og + o, yy, xx] = acc
    if b is not None:
        y += np.asarray(b, F32)[None, :, None, None]
    return y


def relu(x, negative_slope=0.0):
    """``relu_layer.cpp:9-19``: max(x,0) + slope*min(x,0)."""
    x = np.asarray(x, F32)
    if negative_slope == 0.0:
        return np.maximum(x, F32(0))
    return np.maximum(x, F32(0)) + F32(negative_slope) * np.minimum(x, F32(0))


def pool_out_size(size, k, p, s):
    """``pooling_layer.cpp:91-106``: ceil mode, last window must start inside image+pad."""
    o = int(np.ceil(F32(size + 2 * p - k) / F32(s))) + 1
    if p and (o - 1) * s >= size + p:
        o -= 1
    return o


def max_pool(x, k=(2, 2), stride=(2, 2), pad=(0, 0), return_mask=False):
    """``pooling_layer.cpp:140-187`` MAX: windows clipped to the image, init -FLT_MAX, first
    maximum wins (strict ``>``), mask = flat h*W+w index."""
    x = np.asarray(x, F32)
    n, c, h, w = x.shape
    ho = pool_out_size(h, k[0], pad[0], stride[0])
    wo = pool_out_size(w, k[1], pad[1], stride[1])
    y = np.full((n, c, ho, wo), -np.finfo(F32).max, dtype=F32)
    mask = np.full((n, c, ho, wo), -1, dtype=np.int32)
    for py in range(ho):
        hs = py * stride[0] - pad[0]
        he = min(hs + k[0], h)
        hs = max(hs, 0)
        for px in range(wo):
            ws = px * stride[1] - pad[1]
            we = min(ws + k[1], w)
            ws = max(ws, 0)
            for yy in range(hs, he):
                for xx in range(ws, we):
                    v = x[:, :, yy, xx]
                    better = v > y[:, :, py, px]
                    y[:, :, py, px] = np.where(better, v, y[:, :, py, px])
                    mask[:, :, py, px] = np.where(better, yy * w + xx, mask[:, :, py, px])
    return (y, mask) if return_mask else y


def max_pool_2x2_fast(x):
    """2x2/2 max pool for even H, W (what every hot-path level has: H%16==0) -- same result as
    ``max_pool`` without the python loops; used at full benchmark sizes."""
    n, c, h, w = x.shape
    assert h % 2 == 0 and w % 2 == 0
    v = np.asarray(x, F32).reshape(n, c, h // 2, 2, w // 2, 2)
    return v.max(axis=(3, 5))


def col2im(col, c, h, w, kh, kw, ph, pw, sh, sw, dh, dw):
    """``util/im2col.cpp:163-197``: scatter-add columns back into the (C,H,W) image."""
    ho = conv_out_size(h, kh, ph, sh, dh)
    wo = conv_out_size(w, kw, pw, sw, dw)
    img = np.zeros((c, h + 2 * ph, w + 2 * pw), dtype=F32)
    col = col.reshape(c, kh, kw, ho, wo)
    for p in range(kh):
        for q in range(kw):
            img[:, p * dh:p * dh + sh * ho:sh, q * dw:q * dw + sw * wo:sw] += col[:, p, q]
    return img[:, ph:ph + h, pw:pw + w]


def deconv(x, w, b=None, pad=(0, 0), stride=(1, 1), dilation=(1, 1), group=1):
    """Deconvolution forward = conv backward-data: ``deconv_layer.cpp:30-46`` ->
    ``base_conv_layer.cpp:281-297`` backward_cpu_gemm: ``col = W^T @ x`` then col2im.
    Weight blob is (Cin, Cout/group, kh, kw)."""
    x = np.ascontiguousarray(x, dtype=F32)
    w = np.ascontiguousarray(w, dtype=F32)
    n, c, h, wd = x.shape
    cin, cog, kh, kw = w.shape
    assert cin == c
    co = cog * group
    ho = stride[0] * (h - 1) + (dilation[0] * (kh - 1) + 1) - 2 * pad[0]
    wo = stride[1] * (wd - 1) + (dilation[1] * (kw - 1) + 1) - 2 * pad[1]
    cg = c // group
    y = np.empty((n, co, ho, wo), dtype=F32)
    for i in range(n):
        for g in range(group):
            wg = w[g * cg:(g + 1) * cg].reshape(cg, cog * kh * kw)
            col = wg.T @ x[i, g * cg:(g + 1) * cg].reshape(cg, h * wd)
            y[i, g * cog:(g + 1) * cog] = col2im(col, cog, ho, wo, kh, kw, pad[0], pad[1],
                                                 stride[0], stride[1], dilation[0], dilation[1])
        if b is not None:
            y[i] += np.asarray(b, F32)[:, None, None]
    return y


def deconv_depthwise_fast(x, w, pad, stride):
    """group == channels, one output channel per group (the net's ``conv5_256_up``): same sums as
    ``deconv`` without 256 tiny GEMMs; used at full benchmark sizes."""
    n, c, h, wd = x.shape
    _, _, kh, kw = w.shape
    ho = stride[0] * (h - 1) + kh - 2 * pad[0]
    wo = stride[1] * (wd - 1) + kw - 2 * pad[1]
    full = np.zeros((n, c, stride[0] * (h - 1) + kh, stride[1] * (wd - 1) + kw), dtype=F32)
    for p in range(kh):
        for q in range(kw):
            full[:, :, p:p + stride[0] * h:stride[0], q:q + stride[1] * wd:stride[1]] += \
                x * w[None, :, 0, p, q][:, :, None, None]
    return np.ascontiguousarray(full[:, :, pad[0]:pad[0] + ho, pad[1]:pad[1] + wo])


def concat(xs, axis=1):
    """``concat_layer.cpp:47-74``."""
    return np.concatenate([np.asarray(x, F32) for x in xs], axis=axis)


def softmax(x, axis=1):
    """``softmax_layer.cpp:27-60``: subtract channel max, exp, divide by channel sum (fp32)."""
    x = np.asarray(x, F32)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m, dtype=F32)
    return (e / e.sum(axis=axis, keepdims=True, dtype=F32)).astype(F32)


def batch_norm(x, mean, var, factor, eps=1e-5):
    """BatchNormLayer forward with use_global_stats (TEST), ``batch_norm_layer.cpp:98-104,112-117,147-165``:
    scale_factor = 0 if blobs[2][0] == 0 else 1 / blobs[2][0]; mean_ = blobs[0]*scale_factor; variance_ = blobs[1]*scale_factor;
    top = (x - mean_) / sqrt(variance_ + eps), all in float32."""
    x = np.asarray(x, dtype=F32)
    f = F32(np.asarray(factor, dtype=F32).reshape(-1)[0])
    sf = F32(0) if f == 0 else F32(1) / f
    m = np.asarray(mean, F32) * sf
    v = np.sqrt(np.asarray(var, F32) * sf + F32(eps))
    shp = (1, -1) + (1,) * (x.ndim - 2)
    return ((x - m.reshape(shp)) / v.reshape(shp)).astype(F32, copy=False)


def scale(x, gamma, beta=None):
    """ScaleLayer forward, per-channel form (axis 1, num_axes 1; ``scale_layer.cpp:117-147``): top = x * gamma[c] (+ beta[c])."""
    x = np.asarray(x, dtype=F32)
    shp = (1, -1) + (1,) * (x.ndim - 2)
    y = x * np.asarray(gamma, F32).reshape(shp)
    if beta is not None:
        y = y + np.asarray(beta, F32).reshape(shp)
    return y.astype(F32, copy=False)


def bilinear_filler(shape):
    """``filler.hpp:244-262`` BilinearFiller: f = ceil(k/2), c = (2f-1-f%2)/(2f),
    w[x,y] = (1-|x/f-c|)(1-|y/f-c|), identical for every (n,c) plane."""
    assert len(shape) == 4 and shape[2] == shape[3]
    k = shape[3]
    f = int(np.ceil(k / 2.0))
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    out = np.empty(shape, dtype=F32)
    flat = out.reshape(-1)
    for i in range(flat.size):
        x = i % k
        y = (i // k) % k
        flat[i] = (1 - abs(x / f - c)) * (1 - abs(y / f - c))
    return out


def eltwise_sum(xs, coeffs=None):
    """EltwiseLayer SUM, ``eltwise_layer.cpp:52-57``: ``top = coeff0 * bottom0`` (``caffe_set`` + ``caffe_axpy`` in the
    reference, i.e. fp32 multiply-adds in bottom order)."""
    coeffs = [1.0] * len(xs) if not coeffs else list(coeffs)
    y = np.zeros_like(xs[0], dtype=F32)
    for c, x in zip(coeffs, xs):
        y = (y + F32(c) * x.astype(F32)).astype(F32)
    return y
