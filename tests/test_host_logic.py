"""CPU-side checks of the product's host logic and of the drop-in boundary: the C-ABI library loads and
exports every symbol include/shf_b200.h declares (no compute calls without a GPU), the `caffe` /
`caffe.proto.caffe_pb2` shims behave, the py2 source hook handles the reference's files, and the host scalar
logic (pyramid scales, level geometry, anchors, shard ranges) equals the oracle's restatement."""
import os
import re
import sys

import numpy as np
import pytest

from smallhardface_b200 import caffe_proto as cp
from smallhardface_b200 import lib as L
from smallhardface_b200.models import build_test_net, splice_dim_red

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "shf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(shf_\w+|_nms)\s*\(", hdr))
    assert len(declared) >= 20
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    if not os.path.exists(L.LIB_PATH):
        L.build()
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.shf_abi_version() == L.ABI_VERSION == 8                    # a plain host function: safe without a GPU
    assert lib.shf_last_error() is not None


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(L.ShfError, match="no CPU fallback"):
        L.load()


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "smallhardface_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(f)
    assert not bad, bad


def test_pyramid_and_geometry_match_oracle():
    from oracle import preprocess as P
    from smallhardface_b200.detector import DetectConfig, compute_scaling_factor, level_geometry, pyramid_scales
    cfg = DetectConfig()
    for shape in [(1024, 1024, 3), (768, 1024, 3), (1024, 683, 3), (500, 1024, 3), (224, 224, 3), (97, 1301, 3)]:
        assert pyramid_scales(shape, cfg) == P.pyramid_scales(shape)
        assert compute_scaling_factor(shape, 800, 1200) == P.compute_scaling_factor(shape, 800, 1200)
        for s in pyramid_scales(shape, cfg):
            oh, ow, hp, wp = level_geometry(shape[0], shape[1], s)
            assert (oh, ow) == (int(np.rint(shape[0] * s)), int(np.rint(shape[1] * s)))
            assert hp % 16 == 0 and wp % 16 == 0 and 0 <= hp - oh < 16 and 0 <= wp - ow < 16
    assert [level_geometry(1024, 1024, s)[2] for s in pyramid_scales((1024, 1024, 3), cfg)] == [112, 304, 608, 1008, 1408]


def test_anchors_match_reference_golden(golden_dir):
    from smallhardface_b200.anchors import generate_anchors
    g = np.load(os.path.join(golden_dir, "anchors.npz"))
    assert np.array_equal(generate_anchors(16, (1,), (1, 2, 4), (0,), (8, 8, 8)), g["anchors"])
    assert np.array_equal(generate_anchors(strides=(16, 16, 16)), g["default"])


def test_shard_range_is_the_reference_partition():
    from smallhardface_b200.parallel import shard_range
    for n, g in [(3226, 8), (3226, 4), (10, 4), (3, 8), (8, 8)]:
        per = int(np.ceil(1. * n / g))                                   # lib/test.py:329
        got = [shard_range(n, g, r) for r in range(g)]
        assert got == [(min(per * r, n), min(per * (r + 1), n)) for r in range(g)]
        assert sum(b - a for a, b in got) == n


def test_weight_packing_is_exact_to_22_bits():
    from smallhardface_b200.engine import pack_conv_weights
    w = (np.random.RandomState(0).randn(128, 64, 3, 3) * 0.03).astype(np.float32)
    packed, k = pack_conv_weights(w)
    assert packed.shape == (2, 9, 128, 64) and packed.dtype == np.float16
    back = (packed[0].astype(np.float32) + packed[1].astype(np.float32)) * np.float32(2.0 ** -k)
    back = back.reshape(3, 3, 128, 64).transpose(2, 3, 0, 1)
    assert np.abs(back - w).max() <= 2.0 ** -21 * np.abs(w).max()
    assert np.isfinite(packed.astype(np.float32)).all()


def test_fast_format_weight_packing_layout_and_precision():
    """hf8 weights: plane 0 = fp16 hi of w*2^k; plane 1 = per (tap, Cout, 64-channel block) 64 bytes e4m3(hi*2^-6) then
    64 bytes e4m3(lo*2^5).  hi + lo8 reconstructs w to ~2^-15; the 8-bit hi copy is the 4-bit-mantissa image of hi."""
    import torch
    from smallhardface_b200.engine import pack_conv_weights, pack_conv_weights_hf8
    rng = np.random.RandomState(5)
    w = (rng.randn(128, 192, 3, 3) * 0.03).astype(np.float32)
    p8, k8 = pack_conv_weights_hf8(w)
    p2, k2 = pack_conv_weights(w)
    assert k8 == k2 and p8.dtype == np.float16 and p8.shape == (2, 9, 128, 192)
    assert np.array_equal(p8[0], p2[0])                              # same fp16 hi plane as the precise format
    f8 = torch.from_numpy(p8[1].view(np.uint8).copy()).view(torch.float8_e4m3fn).to(torch.float32).numpy()
    f8 = f8.reshape(9, 128, 3, 2, 64)
    hi8 = f8[:, :, :, 0].reshape(9, 128, 192) * 64.0
    lo8 = f8[:, :, :, 1].reshape(9, 128, 192) / 32.0
    hi = p8[0].astype(np.float32)
    ws = w.transpose(2, 3, 0, 1).reshape(9, 128, 192) * np.float32(2.0 ** k8)
    assert np.abs(hi8 - hi).max() <= np.abs(hi).max() * 2.0 ** -4           # e4m3: 3 mantissa bits + rounding
    assert np.abs(hi + lo8 - ws).max() <= np.abs(ws).max() * 2.0 ** -15
    assert np.abs(hi - ws).max() > np.abs(hi + lo8 - ws).max() * 4           # the 8-bit lo really refines hi


def test_conv1_tensor_core_weight_packing():
    from smallhardface_b200.engine import pack_conv1_weights
    w = (np.random.RandomState(6).randn(64, 3, 3, 3) * 0.27).astype(np.float32)
    p, k = pack_conv1_weights(w)
    assert p.shape == (2, 64, 64) and p.dtype == np.float16
    hi, lo = p[0].astype(np.float64), p[1].astype(np.float64)
    assert np.array_equal(hi[:, :27], hi[:, 32:59]) and not hi[:, 27:32].any() and not hi[:, 59:].any() and not lo[:, 27:].any()
    rec = (hi[:, :27] + lo[:, :27]) * 2.0 ** -k
    assert np.abs(rec - w.reshape(64, 27)).max() <= np.abs(w).max() * 2.0 ** -21


BN_NET = """
name: "bn_test"
input: "data"
input_shape { dim: 1 dim: 3 dim: 24 dim: 40 }
layer { name: "c1" type: "Convolution" bottom: "data" top: "c1" convolution_param { num_output: 64 kernel_size: 3 pad: 1 bias_term: false } }
layer { name: "c1_bn" type: "BatchNorm" bottom: "c1" top: "c1" batch_norm_param { use_global_stats: true } }
layer { name: "c1_scale" type: "Scale" bottom: "c1" top: "c1" scale_param { bias_term: true } }
layer { name: "c1_relu" type: "ReLU" bottom: "c1" top: "c1" }
layer { name: "c2" type: "Convolution" bottom: "c1" top: "c2" convolution_param { num_output: 64 kernel_size: 3 pad: 1 } }
layer { name: "c2_bn" type: "BatchNorm" bottom: "c2" top: "c2_bn" batch_norm_param { eps: 0.001 } }
layer { name: "c2_scale" type: "Scale" bottom: "c2_bn" top: "c2_out" }
layer { name: "c2_relu" type: "ReLU" bottom: "c2_out" top: "c2_out" }
"""


def bn_net_params(spec, seed=0):
    """Random but well-conditioned parameters for BN_NET, keyed like graph.load_weights."""
    rng = np.random.RandomState(seed)
    spec.infer_shapes({})
    params = {}
    for l in spec.layers:
        for i, key in enumerate(l.param_keys):
            shp = spec.param_shapes[key]
            if l.type == "Convolution":
                fan = shp[1] * shp[2] * shp[3] if len(shp) == 4 else 1
                params[key] = (rng.randn(*shp) * (np.sqrt(2.0 / fan) if len(shp) == 4 else 0.1)).astype(np.float32)
            elif l.type == "BatchNorm":
                params[key] = [rng.randn(*shp) * 3, rng.rand(*shp) * 4 + 0.5, np.full(shp, 1.7)][i].astype(np.float32)
            else:
                params[key] = ([rng.rand(*shp) + 0.5, rng.randn(*shp) * 0.2][i]).astype(np.float32)
    return params


def test_batchnorm_scale_fold_equals_layerwise_oracle():
    """conv -> BatchNorm -> Scale (-> ReLU) as ONE convolution with folded weights equals the layer-by-layer oracle."""
    from oracle import layers as OL
    from oracle.net import OracleNet
    from smallhardface_b200 import caffe_proto as cp
    from smallhardface_b200.graph import NetSpec, fold_batchnorm_scale
    net = cp.parse_text(BN_NET, "NetParameter")
    onet = OracleNet(net, None, engine="torch", fast=True)
    params = bn_net_params(onet.spec)
    onet.params.update(params)
    x = (np.random.RandomState(1).rand(1, 3, 24, 40) * 255 - 110).astype(np.float32)
    out = onet.forward(data=x)
    assert list(out) == ["c2_out"]
    by = {l.name: l for l in onet.spec.layers}
    k = lambda n, i: params[by[n].param_keys[i]]
    assert by["c2_bn"].p == dict(use_global_stats=True, eps=np.float32(0.001)) or abs(by["c2_bn"].p["eps"] - 0.001) < 1e-9
    w1, b1 = fold_batchnorm_scale(k("c1", 0), None, [("BatchNorm", k("c1_bn", 0), k("c1_bn", 1), k("c1_bn", 2), 1e-5),
                                                     ("Scale", k("c1_scale", 0), k("c1_scale", 1))])
    y1 = OL.relu(OL.conv(x, w1, b1, pad=(1, 1), engine="torch"))
    assert np.abs(y1 - onet.blobs["c1"]).max() <= 2e-5 * np.abs(y1).max()
    w2, b2 = fold_batchnorm_scale(k("c2", 0), k("c2", 1), [("BatchNorm", k("c2_bn", 0), k("c2_bn", 1), k("c2_bn", 2), by["c2_bn"].p["eps"]),
                                                           ("Scale", k("c2_scale", 0), None)])
    y2 = OL.relu(OL.conv(onet.blobs["c1"], w2, b2, pad=(1, 1), engine="torch"))
    assert np.abs(y2 - out["c2_out"]).max() <= 2e-5 * np.abs(y2).max()


def test_precision_model_scheme_equals_the_packed_operands():
    """tools/precision_model.py's `h2f8v2` scheme (the accuracy evidence in DESIGN.md) multiplies exactly the operands that
    engine.pack_conv_weights_hf8 ships and that the kernels' hf8 activation format stores."""
    import importlib.util
    import torch
    from smallhardface_b200.engine import pack_conv_weights_hf8
    spec = importlib.util.spec_from_file_location("precision_model", os.path.join(ROOT, "tools", "precision_model.py"))
    pm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pm)
    rng = np.random.RandomState(9)
    x = (np.abs(rng.randn(1, 64, 6, 7)) * 30).astype(np.float32)
    w = (rng.randn(64, 64, 3, 3) * 0.04).astype(np.float32)
    y_model = pm.make_conv("h2f8v2")(x, w, None, pad=(1, 1))
    packed, k = pack_conv_weights_hf8(w)
    e4 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).clamp(-448, 448).to(torch.float8_e4m3fn).to(torch.float64)
    xt = torch.from_numpy(x).to(torch.float64)
    ah = xt.to(torch.float16).to(torch.float64)
    al8, ah8 = e4((xt - ah) * 64.0), e4(ah / 32.0)
    wh = torch.from_numpy(packed[0].astype(np.float64)).reshape(3, 3, 64, 64).permute(2, 3, 0, 1).contiguous()
    p1 = torch.from_numpy(packed[1].view(np.uint8).copy()).view(torch.float8_e4m3fn).to(torch.float64).reshape(9, 64, 1, 2, 64)
    wh8 = p1[:, :, :, 0].reshape(3, 3, 64, 64).permute(2, 3, 0, 1).contiguous()
    wl8 = p1[:, :, :, 1].reshape(3, 3, 64, 64).permute(2, 3, 0, 1).contiguous()
    cv = lambda a, b: torch.nn.functional.conv2d(a, b, None, padding=1)
    y = ((cv(ah, wh) + cv(al8, wh8) + cv(ah8, wl8)) * 2.0 ** -k).to(torch.float32).numpy()
    assert np.array_equal(y, y_model)


# ---- caffe / caffe_pb2 shims ---------------------------------------------------------------------
def test_caffe_pb2_shim_text_and_wire():
    from smallhardface_b200 import compat
    compat.install()
    import caffe
    from caffe.proto import caffe_pb2
    import google.protobuf.text_format as txtf
    assert (caffe.TRAIN, caffe.TEST) == (0, 1)
    with pytest.raises(RuntimeError):
        caffe.set_mode_cpu()
    txt = cp.format_text(build_test_net(True))
    pb = caffe_pb2.NetParameter()
    txtf.Merge(txt, pb)
    assert len(pb.layer) == 55 and pb.layer[-1].python_param.layer == "ProposalLayer"
    assert pb.layer[0].convolution_param.bias_term is True              # proto2 default survives
    assert pb.SerializeToString() == cp.encode(cp.parse_text(txt))       # same bytes as the hand-written codec
    assert cp.format_text(cp.parse_text(str(pb))) == txt


def test_caffe_pb2_supports_what_manipulate_py_does():
    """lib/prototxt/manipulate.py:98-188 restated call by call on the shim classes."""
    from smallhardface_b200 import compat
    compat.install()
    from caffe.proto import caffe_pb2
    import google.protobuf.text_format as txtf
    pb = caffe_pb2.NetParameter()
    txtf.Merge(cp.format_text(build_test_net(True)), pb)
    split = min(i for i, x in enumerate(pb.layer) if x.name.startswith("head"))
    pb.layer[split - 2].top[0] += "_tmp"
    pb.layer[split - 1].bottom[0] += "_tmp"
    pb.layer[split - 1].top[0] += "_tmp"
    conv = caffe_pb2.LayerParameter()
    conv.name, conv.type = "conv4_fuse_final_dim_red", "Convolution"
    conv.bottom.append("conv4_fuse_final_tmp"); conv.top.append("conv4_fuse_final")
    conv.convolution_param.num_output = 128
    conv.convolution_param.pad.append(1); conv.convolution_param.kernel_size.append(3)
    conv.convolution_param.weight_filler.type = "gaussian"; conv.convolution_param.weight_filler.std = 0.01
    conv.convolution_param.bias_filler.type = "constant"; conv.convolution_param.bias_filler.value = 0.0
    conv.convolution_param.dilation.append(1)
    conv.ClearField("param")
    conv.param.extend([caffe_pb2.ParamSpec()] * 2)
    conv.param[0].lr_mult = 1.0; conv.param[0].decay_mult = 1.0
    conv.param[1].lr_mult = 2.0; conv.param[1].decay_mult = 1.0
    relu = caffe_pb2.LayerParameter()
    relu.name, relu.type = "conv4_fuse_final_dim_red_relu", "ReLU"
    relu.bottom.append("conv4_fuse_final"); relu.top.append("conv4_fuse_final")
    new_layers = pb.layer[:split] + [conv, relu] + pb.layer[split:]
    pb.ClearField("layer")
    pb.layer.extend(new_layers)
    assert cp.format_text(cp.parse_text(str(pb))) == cp.format_text(splice_dim_red(build_test_net(True)))


# ---- py2 hook --------------------------------------------------------------------------------------
def test_py2_transform_units():
    from smallhardface_b200.compat.py2hook import transform_source as T
    assert "print('a', b)" in T("print 'a', b\n")
    assert "print(x, end=' ')" in T("print x,\n")
    assert "print('f: {}'.format(s))" in T("    print 'f: {}'.format(s)\n")
    assert "range(3)" in T("for i in xrange(3): pass\n")
    assert "(k in d)" in T("if d.has_key(k): pass\n")
    assert ".items()" in T("for a, b in d.iteritems(): pass\n")
    assert "import pickle" in T("import cPickle\n")
    assert "'xrange'" in T("s = 'xrange'\n")                          # strings untouched
    assert "except ValueError as e:" in T("try:\n    pass\nexcept ValueError, e:\n    pass\n")
    assert T("from __future__ import print_function\nprint('x', end='')\n").count("print('x', end='')") == 1


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_reference_python_files_import_unmodified_through_the_hook(tmp_path, monkeypatch):
    import glob
    from smallhardface_b200 import compat
    from smallhardface_b200.compat import py2hook
    for f in glob.glob(REF + "/lib/**/*.py", recursive=True) + [REF + "/train_test.py"]:
        compile(py2hook.transform_source(open(f).read(), f), f, "exec")
    # the config module asserts configs/default.toml relative to the CWD (get_config.py:24-27)
    os.symlink(REF + "/configs", tmp_path / "configs")
    os.symlink(REF + "/models", tmp_path / "models")
    monkeypatch.chdir(tmp_path)
    for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k.startswith("lib.") or k == "lib"]:
        del sys.modules[k]
    compat.install(REF)
    try:
        from utils.get_config import cfg
        assert list(cfg.TEST.SCALES) == [100, 300, 600, 1000, 1400] and cfg.TEST.NMS_METHOD == "BBOX_VOTE"
        from lib.layers.generate_anchors import generate_anchors
        a = generate_anchors(scales=np.array([1, 2, 4]), base_size=16, ratios=np.array([1]), shifts=np.array([0]),
                             strides=np.array([8, 8, 8]))
        assert a.tolist() == [[0, 0, 15, 15], [-8, -8, 23, 23], [-24, -24, 39, 39]]
        from utils.test_utils import _compute_scaling_factor
        from smallhardface_b200.detector import compute_scaling_factor
        assert _compute_scaling_factor((768, 1024, 3), 800, 1200) == compute_scaling_factor((768, 1024, 3), 800, 1200)
        import utils.cython_bbox as cb                                    # our drop-in shadows the unbuilt .pyx
        assert cb.__name__.endswith("cython_bbox") and hasattr(cb, "bbox_overlaps_IoA")
        # the reference's own prototxt rewriting runs on the caffe_pb2 shim
        cfg.MODEL.DIFFERENT_DILATION.ENABLE = True
        from lib.prototxt import manipulate
        out = tmp_path / "test.prototxt"
        manipulate.manipulate_test("models/test_template.prototxt", str(out))
        assert cp.format_text(cp.read_net_text(str(out))) == cp.format_text(splice_dim_red(build_test_net(True)))
    finally:
        sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, py2hook.Py2Finder)]
        sys.path[:] = [p for p in sys.path if not p.startswith(REF)]
        for k in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k.startswith("lib.") or k == "lib"]:
            del sys.modules[k]


def test_detection_writer_byte_exact_vs_reference_golden(tmp_path):
    """smallhardface_b200.writers against files written by the reference's own `wider.write_detections`
    (tests/golden/wider_writer/, generated by tests/golden/make_golden.py:make_wider_writer_golden)."""
    import importlib.util
    from smallhardface_b200.writers import write_detections
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    paths, boxes = mg.wider_writer_inputs()
    write_detections(paths, boxes, str(tmp_path))
    gold = os.path.join(ROOT, "tests", "golden", "wider_writer")
    n = 0
    for root, _, files in os.walk(gold):
        for f in files:
            rel = os.path.relpath(os.path.join(root, f), gold)
            assert open(os.path.join(str(tmp_path), rel), "rb").read() == open(os.path.join(root, f), "rb").read(), rel
            n += 1
    assert n == 3
    # the general_* loader's variant: absolute paths with the leading slash stripped (general.py:52-53)
    write_detections(["/data/a/im0.png"], [boxes[2]], str(tmp_path / "g"), extension="png", strip_leading_slash=True)
    txt = open(os.path.join(str(tmp_path / "g"), "data", "a", "im0.txt")).read().splitlines()
    assert txt[0] == "/data/a/im0.png" and txt[1] == "3" and len(txt) == 5 and txt[2].endswith(" ")


def test_py2_hook_reproduces_list_comprehension_variable_leak():
    """Python 2 leaves a list comprehension's loop variable bound in the enclosing scope; `lib/utils/blob.py:21-23` reads
    `im.shape[2]` after `[im.shape for im in ims]`.  The hook reproduces that (and print statements, xrange, has_key ...)
    without touching the files on disk."""
    from smallhardface_b200.compat import py2hook
    src = '''
def shapes(ims):
    m = max([im[0] for im in ims])
    return m, im[1]
def untouched(xs):
    ys = [x + 1 for x in xs]
    return ys
def nested(xs):
    if xs:
        zs = [q * 2 for q in xs]
        return q
    return None
d = {1: 2}
flag = d.has_key(1)
total = 0
for i in xrange(3):
    total += i
print >>__import__("sys").stderr, "py2 print", total
'''
    ns = {}
    exec(py2hook.compile_py2(src, "<py2>"), ns)
    assert ns["shapes"]([(1, 10), (5, 20), (3, 30)]) == (5, 30)
    assert ns["untouched"]([1, 2]) == [2, 3]
    assert ns["nested"]([1, 5]) == 5 and ns["nested"]([]) is None
    assert ns["flag"] is True and ns["total"] == 3
