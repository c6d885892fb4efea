"""Per-kernel parity on the B200: every CUDA entry point of libshf_b200.so against the CPU oracle
(oracle/*.py) on the same seeded inputs, called through the C ABI (smallhardface_b200.lib)."""
import ctypes as C
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():          # collected on the CPU box too; every test here needs the device
    pytest.skip("no CUDA device", allow_module_level=True)

from oracle import layers as OL
from oracle import postprocess as OP
from oracle import preprocess as OPRE
from oracle import proposal as OPROP
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv_weights, split_h2_np

DEV = torch.device("cuda:0")
F32 = np.float32


_KEEP = []          # device tensors passed by raw pointer must outlive the asynchronous launch


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    t = t.to(DEV)
    _KEEP.append(t)
    if len(_KEEP) > 64:
        torch.cuda.synchronize()
        del _KEEP[:-32]
    return t


def h2_roundtrip_np(x):
    hi, lo = split_h2_np(x)
    return hi.astype(F32) + lo.astype(F32)


def relerr(a, b):
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


def test_library_loads_and_reports_b200():
    lib = L.load()
    assert lib.shf_abi_version() == L.ABI_VERSION
    sm, maj, mnr, mem = C.c_int(), C.c_int(), C.c_int(), C.c_longlong()
    L.call("shf_device_info", 0, C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(mem))
    assert maj.value == 10, "built for sm_100a only"
    assert sm.value >= 100


def test_h2_roundtrip():
    x = (np.random.RandomState(0).randn(2, 24, 9, 13) * 50).astype(F32)
    t = H2.from_nchw(dev(x))
    back = t.to_nchw().cpu().numpy()
    assert np.array_equal(back, h2_roundtrip_np(x))
    assert relerr(back, x) < 2 ** -21


def test_conv1_c3_matches_oracle():
    rng = np.random.RandomState(1)
    x = (rng.rand(1, 3, 37, 53) * 255 - 110).astype(F32)
    w = (rng.randn(64, 3, 3, 3) * 0.27).astype(F32)
    b = (rng.randn(64) * 0.05).astype(F32)
    out = H2.empty(1, 37, 53, 64, DEV)
    L.call("shf_conv1_c3", _ptr(dev(x)), _ptr(dev(w)), _ptr(dev(b)), _ptr(out.t), 1, 37, 53, 64, 1, 0, _stream())
    got = out.to_nchw().cpu().numpy()
    ref = OL.relu(OL.conv(x, w, b, pad=(1, 1)))
    assert relerr(got, ref) < 2e-6


@pytest.mark.parametrize("H,W,fmt", [(37, 53, 0), (16, 8, 0), (40, 24, 1), (5, 3, 0)])
def test_conv1_tensor_core_matches_oracle(H, W, fmt):
    """conv1_1 as six tcgen05 MMAs per tile (the product path) vs the oracle and vs its SIMT twin."""
    from smallhardface_b200.engine import pack_conv1_weights
    rng = np.random.RandomState(H + W)
    x = (rng.rand(2, 3, H, W) * 255 - 110).astype(F32)
    w = (rng.randn(64, 3, 3, 3) * 0.27).astype(F32)
    b = (rng.randn(64) * 0.05).astype(F32)
    packed, k = pack_conv1_weights(w)
    out = H2(torch.zeros((2, 2, H, W, 64), dtype=torch.float16, device=DEV), fmt=fmt)
    L.call("shf_conv1_tc", _ptr(dev(x)), _ptr(dev(packed)), _ptr(dev(b)), _ptr(out.t), 2, H, W, 64, float(2.0 ** -k), 1, fmt,
           None, _stream())
    got = out.to_nchw().cpu().numpy()
    ref = OL.relu(OL.conv(x, w, b, pad=(1, 1)))
    assert relerr(got, ref) < (2e-6 if fmt == 0 else 2 ** -14)
    simt = H2.empty(2, H, W, 64, DEV, fmt=fmt)
    L.call("shf_conv1_c3", _ptr(dev(x)), _ptr(dev(w)), _ptr(dev(b)), _ptr(simt.t), 2, H, W, 64, 1, fmt, _stream())
    assert relerr(got, simt.to_nchw().cpu().numpy()) < (2e-6 if fmt == 0 else 2 ** -13)
    # the two-threads-per-pixel instantiation of csrc/conv_first_tc.cu: same operand rows, same MMAs -> the same bits
    from smallhardface_b200.engine import pack_conv_first_tc_weights
    packed2, k2 = pack_conv_first_tc_weights(w)
    assert k2 == k
    out2 = H2(torch.zeros((2, 2, H, W, 64), dtype=torch.float16, device=DEV), fmt=fmt)
    L.call("shf_conv_first_tc", _ptr(dev(x)), _ptr(dev(packed2)), _ptr(dev(b)), _ptr(out2.t), 2, H, W, 64, 3, 1, 1, float(2.0 ** -k2), 1,
           fmt, None, _stream())
    got2 = out2.to_nchw().cpu().numpy()
    assert relerr(got2, ref) < (2e-6 if fmt == 0 else 2 ** -14)
    print("conv1_1 pair kernel vs single-thread kernel: max abs difference %.3g" % float(np.abs(got2 - got).max()))


CONV_CASES = [
    # cin, cout, H, W, k, dil, ctot, coff, relu
    (64, 64, 24, 40, 3, 1, 64, 0, 1),
    (64, 128, 21, 35, 3, 1, 128, 0, 1),       # ragged tiles right/bottom
    (128, 128, 30, 30, 3, 2, 128, 0, 1),      # dilated head
    (128, 128, 30, 30, 3, 4, 128, 0, 1),
    (512, 256, 11, 14, 1, 1, 512, 256, 1),    # 1x1 into a concat window
    (256, 512, 16, 16, 3, 1, 512, 0, 0),      # no ReLU, long K
    (512, 512, 8, 8, 3, 1, 512, 0, 1),        # narrow level (W <= 8 tile shape)
    (64, 64, 5, 7, 3, 1, 64, 0, 1),           # smaller than one tile
]


@pytest.mark.parametrize("cin,cout,H,W,k,dil,ctot,coff,relu", CONV_CASES)
def test_conv_igemm_matches_oracle(cin, cout, H, W, k, dil, ctot, coff, relu):
    rng = np.random.RandomState(cin + cout + H + W + k + dil)
    x = h2_roundtrip_np((np.abs(rng.randn(1, cin, H, W)) * 40 * (rng.rand(1, cin, H, W) > 0.4)).astype(F32))
    w = (rng.randn(cout, cin, k, k) * np.sqrt(2.0 / (cin * k * k))).astype(F32)
    b = (rng.randn(cout) * 0.05).astype(F32)
    packed, kexp = pack_conv_weights(w)
    w_eff = (packed[0].astype(F32) + packed[1].astype(F32)) * F32(2.0 ** -kexp)          # what the kernel multiplies by
    w_eff = w_eff.reshape(k, k, cout, cin).transpose(2, 3, 0, 1)
    assert relerr(w_eff, w) < 2 ** -20
    xin = H2.from_nchw(dev(x))
    out = H2(torch.zeros((2, 1, H, W, ctot), dtype=torch.float16, device=DEV), coff, cout)
    L.call("shf_conv_igemm", _ptr(xin.t), _ptr(dev(packed)), _ptr(dev(b)), _ptr(out.t), 1, H, W, cin, cout, k, dil,
           ctot, coff, float(2.0 ** -kexp), relu, 0, 0, None, _stream())
    torch.cuda.synchronize()
    got = out.to_nchw().cpu().numpy()
    pad = dil if k == 3 else 0
    ref = OL.conv(x, w_eff, b, pad=(pad, pad), dilation=(dil, dil))
    if relu:
        ref = OL.relu(ref)
    err = relerr(got, ref)
    assert err < 3e-6, "tcgen05 conv rel err %.3e" % err
    if ctot != cout:                                   # channels outside the window stay untouched
        full = H2(out.t).to_nchw().cpu().numpy()
        assert np.all(full[:, :coff] == 0) and np.all(full[:, coff + cout:] == 0)
    # the validation kernel agrees too (it is what big-size tests lean on)
    out2 = H2.empty(1, H, W, cout, DEV)
    L.call("shf_debug_conv_direct", _ptr(xin.t), _ptr(dev(w_eff)), _ptr(dev(b)), _ptr(out2.t), 1, H, W, cin, cout, k,
           dil, pad, cout, 0, relu, 0, 0, _stream())
    assert relerr(out2.to_nchw().cpu().numpy(), ref) < 3e-6


@pytest.mark.parametrize("cin,cout,H,W,write_full", [(64, 64, 32, 48, 0), (128, 128, 18, 22, 1), (256, 256, 16, 8, 0)])
def test_conv_relu_pool_fused(cin, cout, H, W, write_full):
    """Convolution + ReLU + MAX 2x2/2 pooling in one launch vs the three oracle layers."""
    rng = np.random.RandomState(cin + H)
    x = h2_roundtrip_np(np.abs(rng.randn(2, cin, H, W) * 20).astype(F32))
    w = (rng.randn(cout, cin, 3, 3) * np.sqrt(2.0 / (cin * 9))).astype(F32)
    b = (rng.randn(cout) * 0.5).astype(F32)
    packed, kexp = pack_conv_weights(w)
    w_eff = ((packed[0].astype(F32) + packed[1].astype(F32)) * F32(2.0 ** -kexp)).reshape(3, 3, cout, cin).transpose(2, 3, 0, 1)
    xin = H2.from_nchw(dev(x))
    full = H2.empty(2, H, W, cout, DEV) if write_full else None
    pooled = H2(torch.zeros((2, 2, H // 2, W // 2, cout + 64), dtype=torch.float16, device=DEV), 64, cout)
    L.call("shf_conv_igemm_pool", _ptr(xin.t), _ptr(dev(packed)), _ptr(dev(b)), _ptr(full.t if full else None),
           _ptr(pooled.t), 2, H, W, cin, cout, 3, 1, cout, 0, cout + 64, 64, float(2.0 ** -kexp), 1, 0, 0, None, _stream())
    ref = OL.relu(OL.conv(x, w_eff, b, pad=(1, 1)))
    refp = OL.max_pool(ref)
    got = pooled.to_nchw().cpu().numpy()
    assert got.shape == refp.shape and relerr(got, refp) < 3e-6
    assert torch.all(pooled.t[..., :64] == 0)
    if full is not None:
        assert relerr(full.to_nchw().cpu().numpy(), ref) < 3e-6


def test_conv_igemm_batch2():
    rng = np.random.RandomState(5)
    x = h2_roundtrip_np(np.abs(rng.randn(2, 64, 18, 20) * 10).astype(F32))
    w = (rng.randn(64, 64, 3, 3) * 0.06).astype(F32)
    packed, kexp = pack_conv_weights(w)
    w_eff = ((packed[0].astype(F32) + packed[1].astype(F32)) * F32(2.0 ** -kexp)).reshape(3, 3, 64, 64).transpose(2, 3, 0, 1)
    xin = H2.from_nchw(dev(x))
    out = H2.empty(2, 18, 20, 64, DEV)
    L.call("shf_conv_igemm", _ptr(xin.t), _ptr(dev(packed)), C.c_void_p(0), _ptr(out.t), 2, 18, 20, 64, 64, 3, 1, 64, 0,
           float(2.0 ** -kexp), 0, 0, 0, None, _stream())
    assert relerr(out.to_nchw().cpu().numpy(), OL.conv(x, w_eff, None, pad=(1, 1))) < 3e-6


@pytest.mark.parametrize("H,W", [(16, 32), (7, 9)])
def test_maxpool(H, W):
    x = h2_roundtrip_np((np.random.RandomState(2).randn(1, 64, H, W) * 30).astype(F32))
    xin = H2.from_nchw(dev(x))
    out = H2.empty(1, (H + 1) // 2, (W + 1) // 2, 64, DEV)
    L.call("shf_maxpool2x2", _ptr(xin.t), _ptr(out.t), 1, H, W, 64, 0, _stream())
    assert np.array_equal(out.to_nchw().cpu().numpy(), OL.max_pool(x))


def test_deconv_depthwise_into_concat_window():
    rng = np.random.RandomState(3)
    c, H, W = 256, 9, 11
    x = h2_roundtrip_np(np.abs(rng.randn(1, c, H, W) * 20).astype(F32))
    w = OL.bilinear_filler((c, 1, 4, 4)) * (1 + 0.1 * rng.rand(c, 1, 1, 1)).astype(F32)
    xin = H2.from_nchw(dev(x))
    dst = torch.zeros((2, 1, 2 * H, 2 * W, 512), dtype=torch.float16, device=DEV)
    L.call("shf_deconv_depthwise", _ptr(xin.t), _ptr(dev(w)), _ptr(dst), 1, H, W, c, 4, 2, 1, 512, 0, 0, 0, None, _stream())
    got = H2(dst, 0, c).to_nchw().cpu().numpy()
    ref = OL.deconv(x, w, None, pad=(1, 1), stride=(2, 2), group=c)
    assert relerr(got, ref) < 1e-6
    assert torch.all(dst[..., c:] == 0)


# ---------------------------------------------------------------------------------------------------------------
# fast operand format "hf8" (fp16 hi plane + 8-bit-float correction plane, include/shf_b200.h)
# ---------------------------------------------------------------------------------------------------------------
def _e4m3(a):
    """float32 -> e4m3 -> float32 the way cvt.rn.satfinite.e4m3x2.f32 does it (round to nearest even, saturate at 448)."""
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=F32)).clamp(-448.0, 448.0)
    return t.to(torch.float8_e4m3fn).to(torch.float32).numpy()


def _e4m3_bytes_to_f32(b):
    return torch.from_numpy(np.ascontiguousarray(b)).view(torch.float8_e4m3fn).to(torch.float32).numpy()


def hf8_planes_np(x):
    """The three operand planes the tensor core sees: ah (fp16), al8 = e4m3((x - ah) * 2^6), ah8 = e4m3(ah * 2^-5)."""
    x = np.asarray(x, F32)
    ah = x.astype(np.float16).astype(F32)
    al8 = _e4m3((x - ah) * F32(64.0))
    ah8 = _e4m3(ah * F32(0.03125))
    return ah, al8, ah8


def hf8_roundtrip_np(x):
    ah, al8, _ = hf8_planes_np(x)
    return (ah.astype(np.float64) + al8.astype(np.float64) / 64.0).astype(F32)


def hf8_conv_model(x, packed8, kexp, cout, cin, k, pad, dil, bias, relu):
    """What the fast conv computes, in float64: ah*wh + (al8*wh8 + ah8*wl8), all at scale 2^kexp (al8 carries 2^6 and
    wh8 2^-6, ah8 2^-5 and wl8 2^5: the products are at the scale of the main term)."""
    ah, al8, ah8 = hf8_planes_np(x)
    wh = packed8[0].astype(np.float64).reshape(k, k, cout, cin).transpose(2, 3, 0, 1)
    p1 = _e4m3_bytes_to_f32(packed8[1].view(np.uint8)).reshape(k * k, cout, cin // 64, 2, 64)
    wh8 = p1[:, :, :, 0].reshape(k, k, cout, cin).transpose(2, 3, 0, 1).astype(np.float64)
    wl8 = p1[:, :, :, 1].reshape(k, k, cout, cin).transpose(2, 3, 0, 1).astype(np.float64)
    cv = lambda a, w: torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)),
                                                 torch.from_numpy(np.ascontiguousarray(w)), None, padding=pad,
                                                 dilation=dil).numpy()
    y = (cv(ah, wh) + cv(al8, wh8) + cv(ah8, wl8)) * 2.0 ** -kexp
    if bias is not None:
        y = y + bias.astype(np.float64)[None, :, None, None]
    if relu:
        y = np.maximum(y, 0)
    return y


def test_hf8_roundtrip():
    x = (np.random.RandomState(0).randn(2, 128, 9, 13) * 50).astype(F32)
    x[0, :, 0, 0] = np.float32(2.0) ** np.arange(-64, 64)[:128].clip(-24, 15)       # wide exponent range
    t = H2.from_nchw(dev(x), fmt=1)
    back = t.to_nchw().cpu().numpy()
    assert np.array_equal(back, hf8_roundtrip_np(x))
    assert relerr(back, x) < 2 ** -14
    # plane 1 really holds [64 x e4m3(lo * 2^6) | 64 x e4m3(hi * 2^-5)] per pixel and 64-channel block
    p1 = t.t[1].view(torch.uint8).reshape(2, 9, 13, 2, 2, 64).cpu()
    ah, al8, ah8 = hf8_planes_np(x)
    got_al8 = p1[:, :, :, :, 0].contiguous().view(torch.float8_e4m3fn).to(torch.float32).numpy().reshape(2, 9, 13, 128).transpose(0, 3, 1, 2)
    got_ah8 = p1[:, :, :, :, 1].contiguous().view(torch.float8_e4m3fn).to(torch.float32).numpy().reshape(2, 9, 13, 128).transpose(0, 3, 1, 2)
    assert np.array_equal(got_al8, al8) and np.array_equal(got_ah8, ah8)


HF8_CONV_CASES = [
    # cin, cout, H, W, k, dil, ctot, coff, relu, out_fmt
    (64, 64, 24, 40, 3, 1, 64, 0, 1, 1),
    (64, 128, 21, 35, 3, 1, 128, 0, 1, 0),
    (128, 128, 30, 30, 3, 2, 128, 0, 1, 0),      # dilated head: reads hf8, writes h2 for the detection tail
    (128, 128, 30, 30, 3, 4, 128, 0, 1, 1),
    (512, 256, 11, 14, 1, 1, 512, 256, 1, 1),    # 1x1 into a concat window
    (256, 512, 16, 16, 3, 1, 512, 0, 0, 1),
    (512, 512, 8, 8, 3, 1, 512, 0, 1, 1),
    (64, 64, 5, 7, 3, 1, 64, 0, 1, 1),
]


@pytest.mark.parametrize("impl", [8, 7])
@pytest.mark.parametrize("cin,cout,H,W,k,dil,ctot,coff,relu,ofmt", HF8_CONV_CASES)
def test_conv_igemm_hf8_matches_operand_model(cin, cout, H, W, k, dil, ctot, coff, relu, ofmt, impl):
    """The fast conv (1 fp16 + 1 fp8 MMA per 16 channels) equals its float64 operand model to accumulation accuracy,
    and stays within ~1e-4 of the fp32 convolution (2^-15-class operands)."""
    from smallhardface_b200.engine import pack_conv_weights_hf8
    rng = np.random.RandomState(cin + cout + H + W + k + dil)
    x = (np.abs(rng.randn(1, cin, H, W)) * 40 * (rng.rand(1, cin, H, W) > 0.4)).astype(F32)
    w = (rng.randn(cout, cin, k, k) * np.sqrt(2.0 / (cin * k * k))).astype(F32)
    b = (rng.randn(cout) * 0.05).astype(F32)
    packed8, kexp = pack_conv_weights_hf8(w)
    pad = dil if k == 3 else 0
    model = hf8_conv_model(x, packed8, kexp, cout, cin, k, pad, dil, b, relu)
    xin = H2.from_nchw(dev(x), fmt=1)
    out = H2(torch.zeros((2, 1, H, W, ctot), dtype=torch.float16, device=DEV), coff, cout, fmt=ofmt)
    L.call("shf_set_conv_impl", impl)
    try:
        L.call("shf_conv_igemm", _ptr(xin.t), _ptr(dev(packed8)), _ptr(dev(b)), _ptr(out.t), 1, H, W, cin, cout, k, dil,
               ctot, coff, float(2.0 ** -kexp), relu, 1, ofmt, None, _stream())
        torch.cuda.synchronize()
    finally:
        L.call("shf_set_conv_impl", 8)
    got = out.to_nchw().cpu().numpy()
    want = model.astype(F32) if ofmt == 0 else hf8_roundtrip_np(model.astype(F32))
    tol = 3e-6 if ofmt == 0 else 2 ** -13         # an hf8 destination re-quantises (a 1-ulp flip of hi moves lo's grid)
    assert relerr(got, want) < tol, "rel err vs operand model %.3e" % relerr(got, want)
    ref = OL.conv(x, w, b, pad=(pad, pad), dilation=(dil, dil))
    if relu:
        ref = OL.relu(ref)
    assert relerr(got, ref) < 2e-4, "rel err vs fp32 conv %.3e" % relerr(got, ref)
    if ctot != cout:
        full = H2(out.t, fmt=ofmt).to_nchw().cpu().numpy()
        assert np.all(full[:, :coff] == 0) and np.all(full[:, coff + cout:] == 0)


def test_conv_relu_pool_fused_hf8():
    from smallhardface_b200.engine import pack_conv_weights_hf8
    cin, cout, H, W = 128, 128, 18, 22
    rng = np.random.RandomState(7)
    x = np.abs(rng.randn(2, cin, H, W) * 20).astype(F32)
    w = (rng.randn(cout, cin, 3, 3) * np.sqrt(2.0 / (cin * 9))).astype(F32)
    b = (rng.randn(cout) * 0.5).astype(F32)
    packed8, kexp = pack_conv_weights_hf8(w)
    model = hf8_conv_model(x, packed8, kexp, cout, cin, 3, 1, 1, b, 1).astype(F32)
    xin = H2.from_nchw(dev(x), fmt=1)
    full = H2.empty(2, H, W, cout, DEV, fmt=1)
    pooled = H2(torch.zeros((2, 2, H // 2, W // 2, cout + 64), dtype=torch.float16, device=DEV), 64, cout, fmt=1)
    L.call("shf_conv_igemm_pool", _ptr(xin.t), _ptr(dev(packed8)), _ptr(dev(b)), _ptr(full.t), _ptr(pooled.t), 2, H, W,
           cin, cout, 3, 1, cout, 0, cout + 64, 64, float(2.0 ** -kexp), 1, 1, 1, None, _stream())
    assert relerr(full.to_nchw().cpu().numpy(), model) < 2 ** -13
    assert relerr(pooled.to_nchw().cpu().numpy(), OL.max_pool(model)) < 2 ** -13
    assert torch.all(pooled.t[0][..., :64] == 0)


def test_simt_layers_hf8():
    """conv1_1 / max pool / depthwise deconv reading and writing the fast format."""
    rng = np.random.RandomState(11)
    x = (rng.rand(1, 3, 20, 28) * 255 - 110).astype(F32)
    w = (rng.randn(64, 3, 3, 3) * 0.27).astype(F32)
    b = (rng.randn(64) * 0.05).astype(F32)
    out = H2.empty(1, 20, 28, 64, DEV, fmt=1)
    L.call("shf_conv1_c3", _ptr(dev(x)), _ptr(dev(w)), _ptr(dev(b)), _ptr(out.t), 1, 20, 28, 64, 1, 1, _stream())
    c1 = out.to_nchw().cpu().numpy()
    ref = OL.relu(OL.conv(x, w, b, pad=(1, 1)))
    assert relerr(c1, ref) < 2 ** -14
    pooled = H2.empty(1, 10, 14, 64, DEV, fmt=1)
    L.call("shf_maxpool2x2", _ptr(out.t), _ptr(pooled.t), 1, 20, 28, 64, 1, _stream())
    assert np.array_equal(pooled.to_nchw().cpu().numpy(), OL.max_pool(c1))      # max of representable values, re-split exactly
    c, H, W = 256, 9, 11
    xd = hf8_roundtrip_np(np.abs(rng.randn(1, c, H, W) * 20).astype(F32))
    wd = OL.bilinear_filler((c, 1, 4, 4)) * (1 + 0.1 * rng.rand(c, 1, 1, 1)).astype(F32)
    xin = H2.from_nchw(dev(xd), fmt=1)
    dst = torch.zeros((2, 1, 2 * H, 2 * W, 512), dtype=torch.float16, device=DEV)
    L.call("shf_deconv_depthwise", _ptr(xin.t), _ptr(dev(wd)), _ptr(dst), 1, H, W, c, 4, 2, 1, 512, 256, 1, 1, None, _stream())
    got = H2(dst, 256, c, fmt=1).to_nchw().cpu().numpy()
    assert relerr(got, OL.deconv(xd, wd, None, pad=(1, 1), stride=(2, 2), group=c)) < 2 ** -14
    assert torch.all(dst[0][..., :256] == 0)


@pytest.mark.parametrize("hw,scale,flip", [((224, 224), 1.0, 0), ((224, 224), 1.3671875, 1), ((96, 130), 0.29296875, 0),
                                           ((96, 130), 2.678571428571429, 1), ((300, 200), 0.5859375, 1)])
def test_preprocess_level_matches_cv2(hw, scale, flip):
    from smallhardface_b200.detector import level_geometry
    im = np.random.RandomState(3).randint(0, 256, hw + (3,)).astype(np.uint8)
    blob = OPRE.get_image_blobs(im, [scale])[0]
    if flip:
        blob = np.ascontiguousarray(blob[..., ::-1])
    ref = OPRE.pad_to_multiple(blob)
    oh, ow, hp, wp = level_geometry(hw[0], hw[1], scale)
    assert (oh, ow) == blob.shape[2:] and (hp, wp) == ref.shape[2:]
    out = torch.empty((1, 3, hp, wp), dtype=torch.float32, device=DEV)
    means = (C.c_double * 3)(*OPRE.PIXEL_MEANS.ravel())
    L.call("shf_preprocess_level", _ptr(dev(im)), hw[0], hw[1], _ptr(out), oh, ow, hp, wp, float(scale), flip, means,
           _stream())
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-5          # <= 1 float32 ulp of a grey level
    assert (got != ref).mean() < 1e-3


def _run_tail(feat_nchw, wc, bc, wb, bb, im_info, topn=10000, score_thresh=0.002):
    """feat list (A) of (1,C,H,W) -> boxes (R,5), probs (R,2) via head_decode + sort + gather."""
    A = len(feat_nchw)
    _, Cc, H, W = feat_nchw[0].shape
    n = H * W * A
    feats = [H2.from_nchw(dev(f)) for f in feat_nchw]
    fp = (C.c_void_p * A)(*[f.t.data_ptr() for f in feats])
    anchors = np.ascontiguousarray(OPROP.generate_anchors(16, (1,), (1, 2, 4), (0,), (8, 8, 8)), F32)
    prob = torch.empty((2 * A, H, W), dtype=torch.float32, device=DEV)
    delta = torch.empty((4 * A, H, W), dtype=torch.float32, device=DEV)
    boxes = torch.empty((n, 4), dtype=torch.float32, device=DEV)
    keys = torch.empty((n,), dtype=torch.int64, device=DEV)
    skeys = torch.empty((n,), dtype=torch.int64, device=DEV)
    meta = torch.zeros((4,), dtype=torch.int64, device=DEV)
    ws_bytes = int(L.load().shf_sort_keys_workspace(n))
    ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=DEV)
    ob = torch.empty((min(topn, n), 5), dtype=torch.float32, device=DEV)
    op = torch.empty((min(topn, n), 2), dtype=torch.float32, device=DEV)
    cptr, rptr, bptr = (C.c_void_p(meta.data_ptr() + o) for o in (0, 4, 8))
    L.call("shf_head_decode", fp, 0, A, _ptr(dev(wc)), _ptr(dev(bc)), _ptr(dev(wb)), _ptr(dev(bb)),
           anchors.ctypes.data_as(C.POINTER(C.c_float)), H, W, Cc, 8, float(im_info[0]), float(im_info[1]), 0.0,
           float(F32(score_thresh)), _ptr(prob), _ptr(delta), _ptr(boxes), _ptr(keys), cptr, bptr, _stream())
    L.call("shf_sort_keys", _ptr(keys), _ptr(skeys), 1, n, None, n, 32, _ptr(ws), ws_bytes, _stream())   # stable: ties keep row order
    valid = keys[keys != -1]                                     # sentinels (~0) are dropped by the sort
    assert torch.equal(skeys[:valid.numel()], torch.sort(valid).values), \
        "stable sort of the score field must equal a full 64-bit sort of (desc score, row) keys"
    assert bool((skeys[valid.numel():] == -1).all())
    L.call("shf_proposal_gather", _ptr(skeys), cptr, bptr, _ptr(prob), _ptr(boxes), A, H * W, min(topn, n), _ptr(ob),
           _ptr(op), rptr, C.c_void_p(0), C.c_void_p(0), 0, 0, 0, 0.0, 1.0, 0.05, _stream())
    R = int(meta.view(torch.int32)[1].item())
    return ob[:R].cpu().numpy(), op[:R].cpu().numpy(), prob.cpu().numpy(), delta.cpu().numpy()


@pytest.mark.parametrize("shared_feature,shift", [(False, -3.0), (True, -3.0), (False, -14.0)])
def test_head_decode_sort_gather_matches_oracle(shared_feature, shift):
    rng = np.random.RandomState(11)
    A, Cc, H, W = 3, 128, 13, 17
    feats = [h2_roundtrip_np(np.abs(rng.randn(1, Cc, H, W)).astype(F32)) for _ in range(1 if shared_feature else A)]
    feats = feats * A if shared_feature else feats
    wc = (rng.randn(A, 2, Cc) * 0.12).astype(F32)
    bc = np.tile(np.array([0.0, shift], F32), (A, 1))
    wb = (rng.randn(A, 4, Cc) * 0.02).astype(F32)
    bb = (rng.randn(A, 4) * 0.01).astype(F32)
    im_info = np.array([[H * 8 - 5, W * 8 - 3, 1.0]], F32)
    boxes, probs, prob_map, delta_map = _run_tail(feats, wc, bc, wb, bb, im_info[0])
    # oracle: 1x1 convs -> cls (1,2A,H,W) [bg.., fg..] / bbox (1,4A,H,W), softmax over (bg,fg), proposal
    cls = np.zeros((1, 2 * A, H, W), F32)
    bbx = np.zeros((1, 4 * A, H, W), F32)
    for a in range(A):
        s = OL.conv(feats[a], wc[a][:, :, None, None], bc[a])
        cls[0, a], cls[0, A + a] = s[0, 0], s[0, 1]
        bbx[0, 4 * a:4 * a + 4] = OL.conv(feats[a], wb[a][:, :, None, None], bb[a])[0]
    sm = OL.softmax(cls.reshape(1, 2, A * H, W), 1).reshape(1, 2 * A, H, W)
    assert np.abs(prob_map - sm[0]).max() < 2e-6
    assert np.abs(delta_map - bbx[0]).max() < 2e-6
    # feed the DEVICE's own probabilities/deltas to the oracle decode: isolates decode+sort+gather, which must
    # then agree to the last bit except for expf ulps in the box sizes
    ref_boxes, ref_probs, _ = OPROP.proposal_forward(prob_map[None], delta_map[None], im_info)
    assert boxes.shape == ref_boxes.shape and probs.shape == ref_probs.shape
    assert np.array_equal(probs, ref_probs)
    assert np.abs(boxes - ref_boxes).max() < 1e-3
    if shift < -10:
        assert boxes.shape[0] == 1                    # nothing above SCORE_THRESH: keep the single best row


def test_sort_ties_lower_index_first():
    """Equal scores must come out in ascending anchor order (the oracle's stated tie rule)."""
    A, Cc, H, W = 3, 128, 4, 5
    feats = [np.zeros((1, Cc, H, W), F32)] * A                      # all logits equal -> all scores 0.5
    z = np.zeros
    boxes, probs, _, _ = _run_tail(feats, z((A, 2, Cc), F32), z((A, 2), F32), z((A, 4, Cc), F32), z((A, 4), F32),
                                   np.array([64, 64, 1.0], F32))
    ref_boxes, ref_probs, order = OPROP.proposal_forward(
        np.full((1, 6, H, W), 0.5, F32), z((1, 12, H, W), F32), np.array([[64, 64, 1.0]], F32))
    assert np.array_equal(order, np.arange(A * H * W))
    assert np.array_equal(boxes, ref_boxes) and np.array_equal(probs, ref_probs)


def _postprocess(dets_list, method, thresh=0.4, mode=0, out_cap=None):
    B = len(dets_list)
    cap = max(1, max(len(d) for d in dets_list))
    out_cap = out_cap or cap
    flat = np.zeros((B, cap, 5), F32)
    for i, d in enumerate(dets_list):
        flat[i, :len(d)] = d
    seg_b = np.arange(B, dtype=np.int32) * cap
    seg_e = seg_b + np.array([len(d) for d in dets_list], np.int32)
    ws_bytes = int(L.load().shf_postprocess_workspace(B, cap))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=DEV)
    oi = torch.empty((B, out_cap), dtype=torch.int32, device=DEV)
    od = torch.empty((B, out_cap, 5), dtype=torch.float32, device=DEV)
    oc = torch.zeros((B,), dtype=torch.int32, device=DEV)
    L.call("shf_postprocess", _ptr(dev(flat)), _ptr(dev(seg_b)), _ptr(dev(seg_e)), B, cap, float(thresh), method, mode,
           _ptr(oi), _ptr(od), _ptr(oc), out_cap, _ptr(ws), ws_bytes, _stream())
    cnt = oc.cpu().numpy()
    assert (cnt <= out_cap).all(), "out_count is the TRUE count; the caller sized out_cap too small"
    if method == 0:
        return [oi[i, :cnt[i]].cpu().numpy().tolist() for i in range(B)]
    return [od[i, :cnt[i]].cpu().numpy() for i in range(B)]


def test_nms_indices_bit_exact_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms.npz"))
    tags = ["n300", "n1", "n1500", "grid"]
    for thr in (0.4, 0.7):
        got = _postprocess([g[t + "_dets"] for t in tags], 0, thr, mode=0)
        for t, k in zip(tags, got):
            assert k == g["%s_cpu_%g" % (t, thr)].tolist(), (t, thr)
    # reference GPU kernel semantics (IoU > thresh) against the oracle restatement of nms_kernel.cu
    got = _postprocess([g["grid_dets"], g["n300_dets"]], 0, 0.4, mode=1)
    assert got[0] == OP.nms(g["grid_dets"], 0.4, OP.NMS_GPU) and got[1] == OP.nms(g["n300_dets"], 0.4, OP.NMS_GPU)
    assert _postprocess([np.zeros((0, 5), F32)], 0)[0] == []


def _clustered_dets(n, seed, centers=None, score_ties=False):
    """n detections jittered around a few hundred face-like centres (so greedy clusters have 1..dozens of members)."""
    rng = np.random.RandomState(seed)
    k = centers or max(1, n // 12)
    c = rng.rand(k, 2) * 900 + 40
    size = rng.rand(k) * 60 + 12
    which = rng.randint(0, k, n)
    ctr = c[which] + rng.randn(n, 2) * (size[which, None] * 0.15)
    wh = size[which, None] * (1 + rng.randn(n, 2) * 0.12)
    d = np.hstack([ctr - wh / 2, ctr + wh / 2, rng.rand(n, 1) * 0.94 + 0.051]).astype(F32)
    if score_ties:
        d[:, 4] = np.round(d[:, 4] * 20) / 20 + F32(0.001)           # many exactly equal scores
    return d


@pytest.mark.parametrize("method", [0, 1], ids=["nms", "vote"])
def test_postprocess_ragged_batch_mask_and_serial_paths_vs_oracle(method):
    """One batched call over images with 0 .. 20000 rows: the small ones take the IoU bit-mask kernels (block edges at
    63/64/65/128/129 rows), the ones above 16384 rows the serial sweep; every image must equal the oracle (NMS keep lists
    bit-exact, voted boxes to float32 sum order)."""
    sizes = [0, 1, 2, 63, 64, 65, 128, 129, 1000, 5000, 16384, 16385, 20000]
    dets = [_clustered_dets(n, 100 + i) if n else np.zeros((0, 5), F32) for i, n in enumerate(sizes)]
    got = _postprocess(dets, method, 0.4, mode=0)
    for n, d, g in zip(sizes, dets, got):
        if method == 0:
            assert g == OP.nms(d, 0.4, OP.NMS_CPU), n
        else:
            ref = OP.bbox_vote(d.copy(), 0.4)
            assert g.shape == ref.shape, (n, g.shape, ref.shape)
            assert np.abs(g[:, :4] - ref[:, :4]).max() < 2e-3 and np.abs(g[:, 4] - ref[:, 4]).max() < 1e-6, n


def test_postprocess_score_ties_and_signed_scores():
    """Ties resolve to the lower row (the oracle's stated order) through sort, NMS and voting; scores of any sign sort
    correctly (the nms drop-ins accept arbitrary user scores, e.g. logits)."""
    d = _clustered_dets(3000, 7, score_ties=True)
    assert len(np.unique(d[:, 4])) < 40
    assert _postprocess([d], 0, 0.4, mode=0)[0] == OP.nms(d, 0.4, OP.NMS_CPU)
    ref = OP.bbox_vote(d.copy(), 0.4)
    got = _postprocess([d], 1, 0.4)[0]
    assert got.shape == ref.shape and np.abs(got - ref).max() < 2e-3
    neg = _clustered_dets(2000, 8)
    neg[:, 4] = np.random.RandomState(9).randn(2000).astype(F32) * 3          # logits: both signs, wide range
    for mode, omode in ((0, OP.NMS_CPU), (1, OP.NMS_GPU)):
        assert _postprocess([neg], 0, 0.5, mode=mode)[0] == OP.nms(neg, 0.5, omode)


def test_segmented_sort_ragged_segments_and_sentinels():
    """shf_sort_keys directly: several segments of different lengths in one launch, sentinels dropped, the 32 score bits
    sorted stably (tile edges at 4096 keys; begin_bit 32 and the batched layout's 27)."""
    rng = np.random.RandomState(5)
    stride = 3 * 4096 + 77
    lens = np.array([0, 1, 31, 4096, 4097, stride, 9000], np.int32)
    S = len(lens)
    for begin_bit in (32, 27):
        row_bits = begin_bit
        score = rng.randint(0, 1 << 20, (S, stride)).astype(np.uint64) << np.uint64(7)       # many duplicate fields
        score[rng.rand(S, stride) < 0.6] = np.uint64(0xffffffff)                              # sentinels
        rows = np.tile(np.arange(stride, dtype=np.uint64), (S, 1))
        keys = (score << np.uint64(row_bits)) | rows
        if begin_bit == 27:
            keys |= (np.arange(S, dtype=np.uint64)[:, None] & np.uint64(31)) << np.uint64(59)
        kin = dev(keys.view(np.int64))
        kout = torch.zeros_like(kin)
        ws_bytes = int(L.load().shf_sort_keys_workspace(S * stride))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=DEV)
        L.call("shf_sort_keys", _ptr(kin), _ptr(kout), S, stride, _ptr(dev(lens)), stride, begin_bit, _ptr(ws), ws_bytes,
               _stream())
        out = kout.cpu().numpy().view(np.uint64)
        assert torch.equal(kin, dev(keys.view(np.int64)))                                     # input untouched
        for sgm in range(S):
            k = keys[sgm, :lens[sgm]]
            k = k[((k >> np.uint64(begin_bit)) & np.uint64(0xffffffff)) != np.uint64(0xffffffff)]
            assert np.array_equal(out[sgm, :len(k)], np.sort(k, kind="stable")), (begin_bit, sgm)
            assert (out[sgm, len(k):lens[sgm]] == np.uint64(0xffffffffffffffff)).all()


def test_nms_host_abi_symbol(golden_dir):
    """`_nms` with the reference's host-pointer contract (lib/nms/gpu_nms.hpp:1-2)."""
    g = np.load(os.path.join(golden_dir, "nms.npz"))
    d = g["n300_dets"]
    order = np.argsort(-d[:, 4], kind="stable")
    sd = np.ascontiguousarray(d[order])
    keep = np.zeros(len(sd), np.int32)
    num = C.c_int(0)
    L.load()._nms(keep.ctypes.data_as(C.POINTER(C.c_int)), C.byref(num), sd.ctypes.data_as(C.POINTER(C.c_float)),
                  len(sd), 5, 0.4, 0)
    assert order[keep[:num.value]].tolist() == OP.nms(d, 0.4, OP.NMS_GPU)


def test_preprocess_level_batched_equals_per_image_kernel():
    """One launch for a stacked (N, h, w, 3) batch x {plain, mirrored} gives, slot for slot, the single-image kernel's bits."""
    from smallhardface_b200.detector import level_geometry
    rng = np.random.RandomState(8)
    n, h, w, scale = 3, 70, 101, 1.3671875
    ims = rng.randint(0, 256, (n, h, w, 3)).astype(np.uint8)
    oh, ow, hp, wp = level_geometry(h, w, scale)
    means = (C.c_double * 3)(*OPRE.PIXEL_MEANS.ravel())
    stacked = dev(ims)
    for passes in (1, 2):
        out = torch.empty((n * passes, 3, hp, wp), dtype=torch.float32, device=DEV)
        L.call("shf_preprocess_level_batched", _ptr(stacked), n, h, w, _ptr(out), oh, ow, hp, wp, float(scale), passes,
               means, _stream())
        for j in range(n):
            for f in range(passes):
                one = torch.empty((1, 3, hp, wp), dtype=torch.float32, device=DEV)
                L.call("shf_preprocess_level", _ptr(stacked[j]), h, w, _ptr(one), oh, ow, hp, wp, float(scale), f, means,
                       _stream())
                assert torch.equal(out[j * passes + f], one[0]), (passes, j, f)


def test_bbox_vote_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "bbox_vote.npz"))
    tags = ["n0", "n1", "n2far", "n400", "n3000"]
    got = _postprocess([g[t + "_dets"] for t in tags], 1, 0.4)
    for t, r in zip(tags, got):
        ref = g[t + "_vote"]
        assert r.shape == ref.shape, t
        assert np.abs(r[:, :4] - ref[:, :4]).max() < 1e-2        # px; measured ~1e-4 (fp32 sum order)
        assert np.abs(r[:, 4] - ref[:, 4]).max() < 1e-6


def test_bbox_overlaps_vs_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "bbox_overlaps.npz"))
    b, q = g["boxes"], g["query"]
    for kind, name in enumerate(["iou", "ioa", "itself"]):
        out = torch.empty((len(b), len(q)), dtype=torch.float64, device=DEV)
        L.call("shf_bbox_overlaps", _ptr(dev(b)), _ptr(dev(q)), len(b), len(q), kind, _ptr(out), _stream())
        assert np.array_equal(out.cpu().numpy(), g[name]), name
