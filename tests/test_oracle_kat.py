"""Known-answer tests re-expressed from the vendored Caffe's gtest suite:
they pin oracle/layers.py, which in turn is what the CUDA kernels are checked against."""
import numpy as np
import pytest

from oracle import layers as L

F32 = np.float32


def test_pooling_square_kat():                  # test_pooling_layer.cpp:49-119
    plane = np.array([[1, 2, 5, 2, 3], [9, 4, 1, 4, 8], [1, 2, 5, 2, 3]], F32)
    x = np.tile(plane, (2, 2, 1, 1))
    y, mask = L.max_pool(x, (2, 2), (1, 1), return_mask=True)
    assert y.shape == (2, 2, 2, 4)
    assert np.array_equal(y[1, 1], [[9, 5, 5, 8], [9, 5, 5, 8]])
    assert np.array_equal(mask[0, 0], [[5, 2, 2, 9], [5, 12, 12, 9]])


def test_pooling_rect_kernels_kat():            # test_pooling_layer.cpp:121-245 (RectHigh 3x2), :247-375 (RectWide 2x3)
    magic6 = np.array([[35, 1, 6, 26, 19, 24], [3, 32, 7, 21, 23, 25], [31, 9, 2, 22, 27, 20],
                       [8, 28, 33, 17, 10, 15], [30, 5, 34, 12, 14, 16], [4, 36, 29, 13, 18, 11]], F32)
    x = np.tile(magic6, (2, 2, 1, 1))
    y, mask = L.max_pool(x, (3, 2), (1, 1), return_mask=True)
    assert y.shape == (2, 2, 4, 5)
    assert np.array_equal(y[1, 0], [[35, 32, 26, 27, 27], [32, 33, 33, 27, 27], [31, 34, 34, 27, 27], [36, 36, 34, 18, 18]])
    assert np.array_equal(mask[0, 1], [[0, 7, 3, 16, 16], [7, 20, 20, 16, 16], [12, 26, 26, 16, 16], [31, 31, 26, 34, 34]])
    y, mask = L.max_pool(x, (2, 3), (1, 1), return_mask=True)
    assert y.shape == (2, 2, 5, 4)
    assert np.array_equal(y[0, 1], [[35, 32, 26, 26], [32, 32, 27, 27], [33, 33, 33, 27], [34, 34, 34, 17], [36, 36, 34, 18]])
    assert np.array_equal(mask[1, 1], [[0, 7, 3, 3], [7, 7, 16, 16], [20, 20, 20, 16], [26, 26, 26, 21], [31, 31, 26, 34]])


def test_pooling_padded_kat():                  # test_pooling_layer.cpp:478-521 TestForwardMaxPadded
    plane = np.array([[1, 2, 4], [2, 3, 2], [4, 2, 1]], F32)
    y = L.max_pool(plane[None, None], (3, 3), (2, 2), (2, 2))
    assert y.shape == (1, 1, 3, 3)
    assert np.array_equal(y[0, 0], [[1, 4, 4], [4, 4, 4], [4, 4, 1]])


def test_pooling_ceil_mode_and_fast_path():
    assert L.pool_out_size(5, 2, 0, 2) == 3 and L.pool_out_size(1408, 2, 0, 2) == 704
    x = np.random.RandomState(0).randn(1, 3, 16, 32).astype(F32)
    assert np.array_equal(L.max_pool(x), L.max_pool_2x2_fast(x))
    y = L.max_pool(x[:, :, :5, :7])             # odd sizes: last window clipped to the image
    assert y.shape == (1, 3, 3, 4) and y[0, 0, 2, 3] == x[0, 0, 4, 6]


def test_deconv_constant_kat():                 # test_deconvolution_layer.cpp:91-137
    x = np.ones((2, 3, 6, 4), F32)
    w = np.ones((3, 4, 3, 3), F32)
    y = L.deconv(x, w, np.full(4, 0.1, F32), stride=(2, 2))
    assert y.shape == (2, 4, 13, 9)
    H, W = y.shape[2:]
    for h in range(H):
        for w_ in range(W):
            e = 3.1
            ho = h % 2 == 0 and 0 < h < H - 1
            wo = w_ % 2 == 0 and 0 < w_ < W - 1
            e += 9 if (ho and wo) else (3 if (ho or wo) else 0)
            assert np.allclose(y[:, :, h, w_], e, atol=1e-4)


def test_deconv_depthwise_bilinear_upsample():
    """The net's conv5_256_up: group=C, k4 s2 p1, bilinear weights -> 2x upsample whose interior is the
    1/4-3/4 blend and whose border ring is attenuated."""
    c = 5
    x = np.random.RandomState(1).rand(1, c, 7, 6).astype(F32)
    w = L.bilinear_filler((c, 1, 4, 4))
    assert np.allclose(w[0, 0], np.outer([.25, .75, .75, .25], [.25, .75, .75, .25]))   # test_filler.cpp:241-280
    y = L.deconv(x, w, None, pad=(1, 1), stride=(2, 2), group=c)
    assert y.shape == (1, c, 14, 12)
    assert np.allclose(y, L.deconv_depthwise_fast(x, w, (1, 1), (2, 2)), atol=1e-6)
    i, j = 3, 2
    exp = .75 * .75 * x[0, :, i, j] + .75 * .25 * x[0, :, i, j + 1] + .25 * .75 * x[0, :, i + 1, j] + .25 * .25 * x[0, :, i + 1, j + 1]
    assert np.allclose(y[0, :, 2 * i + 1, 2 * j + 1], exp, atol=1e-6)
    assert np.allclose(y[0, :, 0, 0], .75 * .75 * x[0, :, 0, 0], atol=1e-6)               # corner sees one tap


@pytest.mark.parametrize("kw", [
    dict(pad=(0, 0), stride=(2, 2), dilation=(1, 1), group=1),          # TestSimpleConvolution :231-265
    dict(pad=(0, 0), stride=(1, 1), dilation=(2, 2), group=1),          # TestDilatedConvolution :267-309
    dict(pad=(1, 1), stride=(1, 1), dilation=(1, 1), group=1),
    dict(pad=(4, 4), stride=(1, 1), dilation=(4, 4), group=1),
    dict(pad=(0, 0), stride=(2, 2), dilation=(1, 1), group=2),          # TestSimpleConvolutionGroup
])
def test_conv_vs_naive_reference_loop(kw):
    rng = np.random.RandomState(1701)
    x = rng.randn(2, 4, 9, 8).astype(F32)
    w = rng.randn(4, 4 // kw["group"], 3, 3).astype(F32)
    b = rng.randn(4).astype(F32)
    ref = L.conv_naive(x, w, b, **kw)
    for eng in ("sgemm", "torch"):
        assert np.allclose(L.conv(x, w, b, engine=eng, **kw), ref, atol=1e-4)
    assert np.allclose(L.conv(x, w, b, max_col_bytes=2000, **kw), ref, atol=1e-4)   # banded column buffer


def test_conv_1x1():                            # test_convolution_layer.cpp:443-468
    rng = np.random.RandomState(2)
    x = rng.randn(2, 6, 5, 4).astype(F32)
    w = rng.randn(3, 6, 1, 1).astype(F32)
    b = rng.randn(3).astype(F32)
    assert np.allclose(L.conv(x, w, b), np.einsum("oc,nchw->nohw", w[:, :, 0, 0], x) + b[None, :, None, None], atol=1e-5)


def test_conv_sobel_separable():                # test_convolution_layer.cpp:498-589
    rng = np.random.RandomState(3)
    x = rng.randn(1, 1, 8, 7).astype(F32)
    sob = np.array([[-1, 0, 1], [-2, 0, 2], [-1, 0, 1]], F32)[None, None]
    full = L.conv(x, sob)
    col = L.conv(x, np.array([[1], [2], [1]], F32)[None, None])
    sep = L.conv(col, np.array([[-1, 0, 1]], F32)[None, None])
    assert np.allclose(full, sep, atol=1e-4)


def test_im2col_layout():                       # util/im2col.cpp:19-55 row order (c, kr, kc)
    x = np.arange(2 * 3 * 3, dtype=F32).reshape(2, 3, 3)
    col, ho, wo = L.im2col(x, 2, 2, 0, 0, 1, 1, 1, 1)
    assert (ho, wo) == (2, 2) and col.shape == (8, 4)
    assert col[0].tolist() == [0, 1, 3, 4] and col[3].tolist() == [4, 5, 7, 8] and col[4].tolist() == [9, 10, 12, 13]
    col, _, _ = L.im2col(x[:1], 3, 3, 1, 1, 1, 1, 1, 1)
    assert col[0].tolist() == [0, 0, 0, 0, 0, 1, 0, 3, 4]             # top-left tap reads the zero pad


def test_softmax_identity():                    # test_softmax_layer.cpp:43-75
    x = np.random.RandomState(4).randn(2, 10, 2, 3).astype(F32) * 5
    y = L.softmax(x, 1)
    assert np.allclose(y.sum(1), 1, atol=1e-5)
    e = np.exp(x.astype(np.float64))
    assert np.allclose(y, e / e.sum(1, keepdims=True), atol=1e-4)
    big = np.array([[[[1000.0]], [[1001.0]]]], F32)                    # max-subtract keeps it finite
    assert np.allclose(L.softmax(big, 1).ravel(), [0.26894143, 0.7310586], atol=1e-6)


def test_concat_values():                       # test_concat_layer.cpp:114-167
    a = np.random.rand(2, 3, 6, 5).astype(F32)
    b = np.random.rand(2, 2, 6, 5).astype(F32)
    y = L.concat([a, b], 1)
    assert y.shape == (2, 5, 6, 5) and np.array_equal(y[:, :3], a) and np.array_equal(y[:, 3:], b)
    c = L.concat([a[:, :, :2], a[:, :, 2:]], 2)
    assert np.array_equal(c, a)


def test_relu():
    x = np.array([-2.0, 0.0, 3.0], F32)
    assert L.relu(x).tolist() == [0, 0, 3]
    assert np.allclose(L.relu(x, 0.1), [-0.2, 0, 3])


def test_batchnorm_global_stats_and_scale():
    """batch_norm_layer.cpp:98-104,147-165 with use_global_stats: the stored sums are divided by the moving-average
    factor (0 -> 0), then y = (x - mean) / sqrt(var + eps); scale_layer.cpp: y = x * gamma + beta per channel."""
    rng = np.random.RandomState(0)
    x = rng.randn(2, 3, 4, 5).astype(np.float32)
    mean_sum = np.array([1.0, -2.0, 0.5], np.float32) * 4
    var_sum = np.array([4.0, 0.25, 1.0], np.float32) * 4
    y = L.batch_norm(x, mean_sum, var_sum, np.array([4.0], np.float32), eps=1e-5)
    for c, (m, v) in enumerate([(1.0, 4.0), (-2.0, 0.25), (0.5, 1.0)]):
        assert np.allclose(y[:, c], (x[:, c] - m) / np.sqrt(v + 1e-5), rtol=1e-6, atol=1e-6)
    y0 = L.batch_norm(x, mean_sum, var_sum, np.array([0.0], np.float32), eps=1e-5)      # factor 0: mean = var = 0
    assert np.allclose(y0, x / np.sqrt(np.float32(1e-5)), rtol=1e-6)
    g, b = np.array([2.0, -1.0, 0.5], np.float32), np.array([0.1, 0.2, -0.3], np.float32)
    z = L.scale(x, g, b)
    assert np.allclose(z[:, 1], -x[:, 1] + 0.2) and np.allclose(L.scale(x, g)[:, 2], 0.5 * x[:, 2])

