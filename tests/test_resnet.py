"""ResNet-style backbone support (BASELINE north_star: "the ResNet/VGG backbone conv stack"): Eltwise SUM, strided 1x1
convolutions, MAX pooling with any kernel / stride / pad, the 7x7 stride-2 first convolution, BatchNorm / Scale folding --
per kernel and as a whole ResNet-50-through-res3 detection net, against the CPU oracle."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

torch = pytest.importorskip("torch")
from oracle import layers as OL
from oracle.indep_net import IndepNet
from oracle.net import OracleNet
from smallhardface_b200 import caffe_proto as cp
from smallhardface_b200 import deploy
from smallhardface_b200.graph import NetSpec, load_weights

F32 = np.float32
HAVE_GPU = torch.cuda.is_available()
needs_gpu = pytest.mark.gpu


def _data(h, w, seed=3):
    im = np.random.RandomState(seed).randint(0, 256, (h, w, 3)).astype(F32)
    return np.ascontiguousarray((im - np.array([[[102.9801, 115.9465, 122.7717]]], F32)).transpose(2, 0, 1)[None])


def test_resnet_deployment_both_oracle_readings_agree(tmp_path):
    """The graph-based reading (shared with the product) and the independent reading agree blob for blob on the ResNet net:
    in-place BatchNorm / Scale / ReLU, Eltwise wiring, ceil-mode 3x3 pooling, strided convs."""
    proto, model = deploy.write_synthetic_resnet_deployment(str(tmp_path), blocks=(2, 2), input_hw=(64, 96))
    spec = NetSpec(cp.read_net_text(proto))
    assert [l.type for l in spec.layers].count("Eltwise") == 4
    shapes = spec.infer_shapes({})
    assert shapes["conv1"] == (1, 64, 32, 48) and shapes["pool1"] == (1, 64, 16, 24) and shapes["res3b"] == (1, 512, 8, 12)
    a = OracleNet(proto, model, engine="torch")
    b = IndepNet(proto, model, engine="torch")
    data, info = _data(64, 96), np.array([[64, 96, 1.0]], F32)
    ra, rb = a.forward(data=data, im_info=info), b.forward(data=data, im_info=info)
    for nm in ("conv1", "pool1", "res2a_branch1", "res2a", "res2b", "res3a_branch2a", "res3a", "res3b", "head"):
        assert np.array_equal(a.blobs[nm], b.blobs[nm]), nm
    assert np.array_equal(ra["boxes"], rb["boxes"]) and len(ra["boxes"]) > 10
    # odd sizes: 3x3/2 ceil-mode pooling keeps a partial last window (pooling_layer.cpp:91-94)
    assert NetSpec(cp.read_net_text(proto)).infer_shapes({"data": (1, 3, 72, 104)})["pool1"] == (1, 64, 18, 26)


def test_eltwise_oracle_kat():
    """test_eltwise_layer.cpp TestSum / TestSumCoeff: a + b + c and a - 0.5 b + 2 c."""
    rng = np.random.RandomState(0)
    a, b, c = (rng.randn(2, 3, 4, 5).astype(F32) for _ in range(3))
    assert np.allclose(OL.eltwise_sum([a, b, c]), a + b + c, atol=1e-6)
    assert np.allclose(OL.eltwise_sum([a, b, c], [1, -0.5, 2]), a - 0.5 * b + 2 * c, atol=1e-5)


def test_eltwise_spec_errors():
    from smallhardface_b200.models import build_resnet_test_net
    net = build_resnet_test_net(blocks=(1, 1))
    for l in net.layer:
        if l.type == "Eltwise":
            l.eltwise_param = cp.Msg("EltwiseParameter", coeff=[1.0])
            break
    with pytest.raises(ValueError, match="one coefficient per bottom"):
        NetSpec(net)


# ------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------
if HAVE_GPU:
    from smallhardface_b200 import lib as L
    from smallhardface_b200.engine import H2, GpuNet, _ptr, _stream, pack_conv_weights, pack_conv_weights_hf8
    DEV = torch.device("cuda:0")
_KEEP = []


def dev(a):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    _KEEP.append(t)
    return t


def relerr(a, b):
    return float(np.abs(a - b).max() / max(1e-30, np.abs(b).max()))


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
@pytest.mark.parametrize("fmt,n_in,relu,coeffs", [(0, 2, 1, None), (1, 2, 1, None), (0, 3, 0, [1.0, -0.5, 2.0]), (1, 1, 0, [0.25])])
def test_eltwise_sum_kernel(fmt, n_in, relu, coeffs):
    rng = np.random.RandomState(n_in + fmt)
    xs = [(rng.randn(2, 64, 9, 13) * 40).astype(F32) for _ in range(n_in)]
    hs = [H2.from_nchw(dev(x), fmt) for x in xs]
    rounded = [h.to_nchw().cpu().numpy() for h in hs]                 # what the kernel actually reads
    out = H2.empty(2, 9, 13, 64, DEV, fmt)
    guard = torch.zeros(1, dtype=torch.int32, device=DEV)
    ptrs = (C.c_void_p * n_in)(*[h.t.data_ptr() for h in hs])
    cf = (C.c_float * n_in)(*coeffs) if coeffs else None
    L.call("shf_eltwise_sum", ptrs, cf, n_in, _ptr(out.t), 2 * 9 * 13, 64, 64, 0, relu, fmt, fmt, _ptr(guard), _stream())
    ref = OL.eltwise_sum(rounded, coeffs)
    if relu:
        ref = OL.relu(ref)
    got = out.to_nchw().cpu().numpy()
    assert relerr(got, ref) < (2e-6 if fmt == 0 else 2 ** -14)
    assert abs(float(guard.cpu().numpy().view(F32)[0]) - np.abs(ref).max()) < 1e-3 * np.abs(ref).max()


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
@pytest.mark.parametrize("H,W,k,s,p", [(112, 112, 3, 2, 0), (35, 51, 3, 2, 0), (17, 20, 3, 2, 1), (12, 12, 2, 2, 0), (9, 14, 3, 1, 1),
                                       (10, 11, 4, 3, 0)])
def test_maxpool_general_bit_exact(H, W, k, s, p):
    x = (np.random.RandomState(H * W).randn(2, 64, H, W) * 30).astype(F32)
    for fmt in (0, 1):
        h = H2.from_nchw(dev(x), fmt)
        src = h.to_nchw().cpu().numpy()
        ref = OL.max_pool(src, (k, k), (s, s), (p, p))
        ho, wo = ref.shape[2:]
        lib = L.load()
        assert (lib.shf_pool_out_size(H, k, s, p, int(p > 0)), lib.shf_pool_out_size(W, k, s, p, int(p > 0))) == (ho, wo)
        out = H2.empty(2, ho, wo, 64, DEV, fmt)
        L.call("shf_maxpool", _ptr(h.t), _ptr(out.t), 2, H, W, 64, k, k, s, s, p, p, fmt, _stream())
        assert np.array_equal(out.to_nchw().cpu().numpy(), ref)        # a maximum re-splits into the planes it came from


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
@pytest.mark.parametrize("H,W,k,s,p,fmt", [(64, 96, 7, 2, 3, 0), (37, 53, 7, 2, 3, 1), (20, 24, 3, 1, 1, 0), (33, 31, 5, 2, 2, 0),
                                           (9, 9, 7, 2, 3, 0)])
def test_conv_first_matches_oracle(H, W, k, s, p, fmt):
    rng = np.random.RandomState(H + W + k)
    x = (rng.rand(2, 3, H, W) * 255 - 110).astype(F32)
    w = (rng.randn(64, 3, k, k) * np.sqrt(2.0 / (3 * k * k))).astype(F32)
    b = (rng.randn(64) * 0.05).astype(F32)
    ref = OL.relu(OL.conv(x, w, b, pad=(p, p), stride=(s, s)))
    out = H2.empty(2, ref.shape[2], ref.shape[3], 64, DEV, fmt)
    L.call("shf_conv_first", _ptr(dev(x)), _ptr(dev(w)), _ptr(dev(b)), _ptr(out.t), 2, H, W, 64, k, s, p, 1, fmt, None, _stream())
    assert relerr(out.to_nchw().cpu().numpy(), ref) < (3e-6 if fmt == 0 else 2 ** -14)


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
@pytest.mark.parametrize("H,W,pad,fmt,batch", [(64, 96, 3, 0, 2), (37, 53, 3, 1, 2), (9, 9, 3, 0, 1), (224, 224, 3, 1, 1), (40, 31, 0, 0, 3),
                                               (17, 300, 2, 0, 1)])
def test_conv7_tensor_core_matches_oracle_and_simt_twin(H, W, pad, fmt, batch):
    """shf_conv_first_tc, 7x7 stride 2 (on tcgen05: 147-tap split-fp16 operand rows, 30 MMAs per tile) vs the oracle and vs its
    fp32 SIMT twin shf_conv_first."""
    from smallhardface_b200.engine import pack_conv_first_tc_weights
    rng = np.random.RandomState(H * 7 + W)
    x = (rng.rand(batch, 3, H, W) * 255 - 110).astype(F32)
    w = (rng.randn(64, 3, 7, 7) * np.sqrt(2.0 / 147)).astype(F32)
    b = (rng.randn(64) * 0.05).astype(F32)
    ref = OL.relu(OL.conv(x, w, b, pad=(pad, pad), stride=(2, 2)))
    packed, e = pack_conv_first_tc_weights(w)
    out = H2(torch.zeros((2, batch, ref.shape[2], ref.shape[3], 64), dtype=torch.float16, device=DEV), fmt=fmt)
    guard = torch.zeros(1, dtype=torch.int32, device=DEV)
    L.call("shf_conv_first_tc", _ptr(dev(x)), _ptr(dev(packed)), _ptr(dev(b)), _ptr(out.t), batch, H, W, 64, 7, 2, pad, float(2.0 ** -e),
           1, fmt, _ptr(guard), _stream())
    got = out.to_nchw().cpu().numpy()
    assert got.shape == ref.shape and relerr(got, ref) < (3e-6 if fmt == 0 else 2 ** -14)
    assert abs(float(guard.cpu().numpy().view(F32)[0]) - np.abs(ref).max()) < 1e-3 * np.abs(ref).max()
    simt = H2.empty(batch, ref.shape[2], ref.shape[3], 64, DEV, fmt)
    L.call("shf_conv_first", _ptr(dev(x)), _ptr(dev(w)), _ptr(dev(b)), _ptr(simt.t), batch, H, W, 64, 7, 2, pad, 1, fmt, None, _stream())
    assert relerr(got, simt.to_nchw().cpu().numpy()) < (3e-6 if fmt == 0 else 2 ** -13)


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
@pytest.mark.parametrize("cin,cout,H,W,s,fmt", [(256, 128, 28, 28, 2, 0), (256, 512, 31, 45, 2, 0), (64, 64, 16, 24, 2, 1), (128, 256, 13, 9, 3, 0)])
def test_strided_1x1_conv_matches_oracle(cin, cout, H, W, s, fmt):
    rng = np.random.RandomState(cin + H)
    x = (rng.randn(2, cin, H, W) * 20).astype(F32)
    w = (rng.randn(cout, cin, 1, 1) * np.sqrt(2.0 / cin)).astype(F32)
    b = (rng.randn(cout) * 0.05).astype(F32)
    h = H2.from_nchw(dev(x), fmt)
    src = h.to_nchw().cpu().numpy()
    ref = OL.conv(src, w, b, stride=(s, s))
    packed, kexp = (pack_conv_weights_hf8 if fmt else pack_conv_weights)(w)
    out = H2.empty(2, ref.shape[2], ref.shape[3], cout, DEV, fmt)
    L.call("shf_conv_igemm_strided", _ptr(h.t), _ptr(dev(packed)), _ptr(dev(b)), _ptr(out.t), 2, H, W, s, cin, cout, cout, 0,
           float(2.0 ** -kexp), 0, fmt, fmt, None, _stream())
    assert relerr(out.to_nchw().cpu().numpy(), ref) < (3e-6 if fmt == 0 else 3e-4)


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
@pytest.mark.parametrize("cin,cout,H,W,fmt,batch", [(128, 128, 28, 28, 0, 2), (64, 64, 17, 23, 0, 1), (256, 256, 31, 45, 1, 2), (128, 256, 56, 56, 1, 1),
                                                    (64, 128, 2, 2, 0, 1), (512, 512, 9, 14, 0, 1)])
def test_conv3x3_stride2_matches_oracle(cin, cout, H, W, fmt, batch):
    """shf_conv3x3_s2: the 3x3 stride-2 pad-1 convolution on the four parity views of the input (one TMA load per tap and
    64-channel chunk) against the oracle's strided convolution -- even / odd sizes, both tile widths, both formats."""
    rng = np.random.RandomState(cin + H + W)
    x = (rng.randn(batch, cin, H, W) * 20).astype(F32)
    w = (rng.randn(cout, cin, 3, 3) * np.sqrt(2.0 / (9 * cin))).astype(F32)
    b = (rng.randn(cout) * 0.05).astype(F32)
    h = H2.from_nchw(dev(x), fmt)
    ref = OL.relu(OL.conv(h.to_nchw().cpu().numpy(), w, b, pad=(1, 1), stride=(2, 2)))
    packed, kexp = (pack_conv_weights_hf8 if fmt else pack_conv_weights)(w)
    out = H2.empty(batch, ref.shape[2], ref.shape[3], cout, DEV, fmt)
    L.call("shf_conv3x3_s2", _ptr(h.t), _ptr(dev(packed)), _ptr(dev(b)), _ptr(out.t), batch, H, W, cin, cout, cout, 0,
           float(2.0 ** -kexp), 1, fmt, fmt, None, _stream())
    got = out.to_nchw().cpu().numpy()
    assert got.shape == ref.shape and relerr(got, ref) < (3e-6 if fmt == 0 else 3e-4)


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
def test_resnet_with_the_stride_on_the_3x3_convolution(tmp_path):
    """The torchvision / "v1.5" bottleneck (stride on the 3x3 convolution of the stage transition) as a whole net."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_net import BOX_TOL, SCORE_TOL, match_rows
    hw = (150, 202)
    proto, model = deploy.write_synthetic_resnet_deployment(str(tmp_path), blocks=(2, 3), input_hw=hw, stride_on_3x3=True)
    spec = NetSpec(cp.read_net_text(proto))
    gnet = GpuNet(spec, load_weights(spec, cp.read_net_binary(model)), "cuda:0", fast_min_scale=None)
    assert sum(1 for k, _, s in gnet.ops if k == "conv" and s["stride"] == 2 and s["k"] == 3) == 1
    onet = IndepNet(proto, model, engine="torch")
    data, info = _data(*hw), np.array([[hw[0], hw[1], 1.0]], F32)
    ref = onet.forward(data=data, im_info=info)
    boxes, probs, rows = gnet.forward(torch.from_numpy(data).to(DEV), info[0])
    torch.cuda.synchronize()
    R = int(rows.item())
    for nm in ("res3a_branch2b", "res3a", "res3c", "head"):
        assert relerr(gnet.blob_nchw(nm).cpu().numpy(), onet.blobs[nm]) < 2e-5, nm
    ws, wb = match_rows(boxes[:R].cpu().numpy()[:, 1:], probs[:R].cpu().numpy()[:, 1], ref["boxes"][:, 1:], ref["cls_prob"][:, 1])
    print("resnet (stride on the 3x3) rows %d (ref %d): worst score err %.2e, worst box err %.2e px" % (R, len(ref["boxes"]), ws, wb))
    assert R > 20 and ws < SCORE_TOL and wb < BOX_TOL


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
@pytest.mark.parametrize("cin,cout,k,H,W,fmt,relu", [(64, 256, 1, 20, 28, 0, 1), (512, 128, 1, 9, 13, 1, 1), (128, 128, 3, 17, 24, 0, 0),
                                                    (256, 64, 1, 16, 16, 1, 1)])
def test_conv_with_fused_residual_matches_oracle(cin, cout, k, H, W, fmt, relu):
    """shf_conv_igemm_res: conv * scale + bias + shortcut (+ ReLU) in one launch vs conv -> Eltwise SUM -> ReLU of the oracle."""
    rng = np.random.RandomState(cin + cout + H)
    x = (rng.randn(2, cin, H, W) * 20).astype(F32)
    w = (rng.randn(cout, cin, k, k) * np.sqrt(2.0 / (cin * k * k))).astype(F32)
    b = (rng.randn(cout) * 0.05).astype(F32)
    res = (rng.randn(2, cout, H, W) * 30).astype(F32)
    hx, hr = H2.from_nchw(dev(x), fmt), H2.from_nchw(dev(res), fmt)
    ref = OL.eltwise_sum([hr.to_nchw().cpu().numpy(), OL.conv(hx.to_nchw().cpu().numpy(), w, b, pad=(k // 2, k // 2))])
    if relu:
        ref = OL.relu(ref)
    packed, kexp = (pack_conv_weights_hf8 if fmt else pack_conv_weights)(w)
    out = H2.empty(2, H, W, cout, DEV, fmt)
    L.call("shf_conv_igemm_res", _ptr(hx.t), _ptr(dev(packed)), _ptr(dev(b)), _ptr(hr.t), _ptr(out.t), 2, H, W, cin, cout, k, 1,
           cout, 0, cout, 0, float(2.0 ** -kexp), relu, fmt, fmt, fmt, None, _stream())
    assert relerr(out.to_nchw().cpu().numpy(), ref) < (3e-6 if fmt == 0 else 3e-4)


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
@pytest.mark.parametrize("hw,fuse", [((224, 224), True), ((150, 202), True), ((150, 202), False)])
def test_resnet_net_forward_matches_oracle(tmp_path, hw, fuse):
    """ResNet-50 through res3 + the standard detection head: every residual stage and the outputs against the independent
    oracle reading; BASELINE tolerances (scores 1e-3, boxes 1e-2 px).  fuse: residual adds inside the branch2c conv
    launches (the default) or as Eltwise launches of their own."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_net import BOX_TOL, SCORE_TOL, match_rows
    proto, model = deploy.write_synthetic_resnet_deployment(str(tmp_path), blocks=(3, 4), input_hw=hw)
    spec = NetSpec(cp.read_net_text(proto))
    gnet = GpuNet(spec, load_weights(spec, cp.read_net_binary(model)), "cuda:0", fast_min_scale=None, fuse_pool=fuse)
    kinds = [k for k, _, _ in gnet.ops]
    assert kinds.count("eltwise") == (0 if fuse else 7) and sum(1 for _, _, s in gnet.ops if "residual" in s) == (7 if fuse else 0)
    assert kinds.count("conv_first") == 1 and "pool" in kinds
    assert sum(1 for k, _, s in gnet.ops if k == "conv" and s["stride"] == 2) == 2          # res3a_branch1 / _branch2a
    onet = IndepNet(proto, model, engine="torch")
    data, info = _data(*hw), np.array([[hw[0], hw[1], 1.0]], F32)
    ref = onet.forward(data=data, im_info=info)
    boxes, probs, rows = gnet.forward(torch.from_numpy(data).to(DEV), info[0])
    torch.cuda.synchronize()
    assert not gnet.check_ranges()
    R = int(rows.item())
    report = {}
    for nm in ("conv1", "pool1", "res2a", "res2c", "res3a_branch1", "res3a", "res3d", "head"):
        got, want = gnet.blob_nchw(nm).cpu().numpy(), onet.blobs[nm]
        assert got.shape == want.shape, nm
        report[nm] = relerr(got, want)
    print("resnet per-blob max rel err:", {k: "%.2e" % v for k, v in report.items()})
    assert max(report.values()) < 2e-5, report
    gb, gp = boxes[:R].cpu().numpy(), probs[:R].cpu().numpy()
    ws, wb = match_rows(gb[:, 1:], gp[:, 1], ref["boxes"][:, 1:], ref["cls_prob"][:, 1])
    print("resnet %s rows %d (ref %d): worst score err %.2e, worst box err %.2e px" % (hw, R, len(ref["boxes"]), ws, wb))
    assert R > 50 and ws < SCORE_TOL and wb < BOX_TOL


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
def test_resnet_detector_pyramid_fast_format(tmp_path):
    """The same net through Detector (pyramid + flip + voting) with the default operand policy (hf8 on levels >= 0.9)."""
    from oracle import detect as OD
    from smallhardface_b200.detector import DetectConfig, Detector
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_net import BOX_TOL, SCORE_TOL, match_rows
    proto, model = deploy.write_synthetic_resnet_deployment(str(tmp_path), blocks=(3, 4))
    cfg = DetectConfig(scales=(300, 800))
    det = Detector(proto, model, "cuda:0", cfg)
    im = deploy.synthetic_image(4, (160, 208))
    b = det.detect_device(det.upload([im]))
    got = det.download(b, 1)[0]
    raw = det.raw_detections(b, 0)
    onet = IndepNet(proto, model, engine="torch")
    probs, boxes = OD.detect_raw(onet, im, scales=cfg.scales, flip=True)
    ref = OD.threshold_dets(probs, boxes, 0.05)
    ws, wb = match_rows(raw[:, :4], raw[:, 4], ref[:, :4], ref[:, 4])
    print("resnet detector: %d raw rows (ref %d), %d voted; worst score err %.2e, box err %.2e px" % (len(raw), len(ref), len(got), ws, wb))
    assert len(ref) > 20 and ws < SCORE_TOL and wb < BOX_TOL


@needs_gpu
@pytest.mark.skipif(not HAVE_GPU, reason="no CUDA device")
def test_resnet_through_the_caffe_net_surface(tmp_path):
    """The ResNet deployment through the drop-in `caffe.Net` (plugin path: CUDA graph per level shape -- first forward
    captures, second replays), driven like lib/test.py:21-66, both passes of two pyramid levels against the oracle."""
    from oracle import detect as OD
    from oracle import preprocess as OPRE
    from smallhardface_b200 import compat
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_net import BOX_TOL, SCORE_TOL, match_rows
    compat.install()
    import caffe
    proto, model = deploy.write_synthetic_resnet_deployment(str(tmp_path), blocks=(3, 4))
    caffe.set_mode_gpu()
    caffe.set_device(0)
    net = caffe.Net(str(proto), str(model), caffe.TEST)
    assert net.outputs == ["boxes", "cls_prob"] and "res3d" in net.blobs and net.params["conv1"][0].data.shape == (64, 3, 7, 7)
    onet = IndepNet(proto, model, engine="torch")
    im = deploy.synthetic_image(9, (120, 168))
    scales = OPRE.pyramid_scales(im.shape, (300, 800))
    worst = 0.0
    for rep in range(2):                                   # 0: graph capture, 1: replay
        for blob, s in zip(OPRE.get_image_blobs(im, scales), scales):
            for flip in (False, True):
                d = np.ascontiguousarray(blob[..., ::-1]) if flip else blob
                h, w = d.shape[2:]
                nh, nw = -(-h // 16) * 16, -(-w // 16) * 16
                data = np.pad(d, ((0, 0), (0, 0), (0, nh - h), (0, nw - w)), "constant").astype(F32)
                info = np.array([[h, w, s]], F32)
                net.blobs["data"].reshape(*data.shape)
                net.blobs["im_info"].reshape(*info.shape)
                out = net.forward(data=data, im_info=info)
                if flip:
                    out["boxes"][:, [1, 3]] = w - out["boxes"][:, [3, 1]]
                boxes = net.blobs["boxes"].data[:, 1:5] / s
                probs = net.blobs["cls_prob"].data
                rp, rb = OD.forward_level(onet, d, s, flip)
                ws, wb = match_rows(boxes, probs[:, 1], rb[:, :4], rp[:, 1])
                assert ws < SCORE_TOL and wb < BOX_TOL, (rep, s, flip, ws, wb)
                worst = max(worst, wb)
    got = net.blobs["res2c"].data                           # an intermediate blob through the lazy device->host sync
    assert got.shape == onet.blobs["res2c"].shape and relerr(got, onet.blobs["res2c"]) < 5e-4
    print("resnet through caffe.Net: worst box error %.2e raw px" % worst)
