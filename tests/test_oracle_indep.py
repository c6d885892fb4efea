"""The oracle's two readings of a deployment must agree: oracle/net.py (graph + codecs shared with the product) and
oracle/indep_net.py (own text reader, own wire reader, own wiring).  A wiring error in the shared code -- channel order
of the axis-2 cls concat, shared head weights, the dim_red splice, in-place tops -- would make them differ."""
import os
import re

import numpy as np
import pytest

from oracle.indep_net import IndepNet, parse_prototxt, read_caffemodel, _all, _one
from oracle.net import OracleNet
from smallhardface_b200 import deploy

F32 = np.float32
REF_PROTO = "/root/reference/caffe/src/caffe/proto/caffe.proto"


def _inputs(hw=(96, 128), seed=3):
    im = np.random.RandomState(seed).randint(0, 256, hw + (3,)).astype(np.uint8)
    data = (im.astype(F32) - np.array([[[102.9801, 115.9465, 122.7717]]])).astype(F32).transpose(2, 0, 1)[None]
    return np.ascontiguousarray(data), np.array([[hw[0], hw[1], 1.0]], F32)


@pytest.mark.parametrize("dilation", [True, False], ids=["dilation", "standard"])
def test_both_readings_agree_on_every_blob(tmp_path, dilation):
    proto, model = deploy.write_synthetic_deployment(str(tmp_path), dilation=dilation)
    a = OracleNet(proto, model, engine="torch", fast=True)
    b = IndepNet(proto, model, engine="torch")
    data, info = _inputs()
    ra, rb = a.forward(data=data, im_info=info), b.forward(data=data, im_info=info)
    assert list(ra) == list(rb) == ["boxes", "cls_prob"]
    assert set(a.blobs) == set(b.blobs)
    for k in a.blobs:
        assert a.blobs[k].shape == b.blobs[k].shape and np.array_equal(a.blobs[k], b.blobs[k]), k
    assert len(ra["boxes"]) > 50                     # a non-trivial forward


# hand-written expectation for the deployed dilation net (test_different_dilation_template.prototxt:479-669 after the
# dim_red splice of lib/prototxt/manipulate.py:166-188): layer -> (bottoms, tops[, extra])
TAIL_WIRING = [
    ("conv4_fuse_final", ["conv4_fuse"], ["conv4_fuse_final_tmp"]),                 # manipulate.py:171-172
    ("conv4_fuse_final_relu", ["conv4_fuse_final_tmp"], ["conv4_fuse_final_tmp"]),  # manipulate.py:173-174
    ("conv4_fuse_final_dim_red", ["conv4_fuse_final_tmp"], ["conv4_fuse_final"]),   # manipulate.py:176-183
    ("conv4_fuse_final_dim_red_relu", ["conv4_fuse_final"], ["conv4_fuse_final"]),
    ("head_1", ["conv4_fuse_final"], ["head_1"]),
    ("head_2", ["conv4_fuse_final"], ["head_2"]),
    ("head_4", ["conv4_fuse_final"], ["head_4"]),
    ("cls_score_1", ["head_1"], ["cls_score_1_output"]),
    ("cls_score_2", ["head_2"], ["cls_score_2_output"]),
    ("cls_score_4", ["head_4"], ["cls_score_4_output"]),
    ("bbox_pred_1", ["head_1"], ["bbox_pred_1_output"]),
    ("bbox_pred_2", ["head_2"], ["bbox_pred_2_output"]),
    ("bbox_pred_4", ["head_4"], ["bbox_pred_4_output"]),
    # test_different_dilation_template.prototxt:646-669: cls maps stacked along HEIGHT in anchor order, bbox along channels
    ("cls_score_output_concat", ["cls_score_1_output", "cls_score_2_output", "cls_score_4_output"], ["cls_score_reshape_output"]),
    ("bbox_pred_output_concat", ["bbox_pred_1_output", "bbox_pred_2_output", "bbox_pred_4_output"], ["bbox_pred_output"]),
    ("cls_prob", ["cls_score_reshape_output"], ["cls_prob_output"]),
    ("cls_prob_reshape", ["cls_prob_output"], ["cls_prob_reshape_output"]),
    ("proposal", ["cls_prob_reshape_output", "bbox_pred_output", "im_info"], ["boxes", "cls_prob"]),
]


def test_deployed_dilation_net_wiring_table(tmp_path):
    proto, model = deploy.write_synthetic_deployment(str(tmp_path), dilation=True)
    net = parse_prototxt(open(proto).read())
    layers = {_one(l, "name"): l for l in _all(net, "layer")}
    for name, bottoms, tops in TAIL_WIRING:
        assert _all(layers[name], "bottom") == bottoms and _all(layers[name], "top") == tops, name
    for d, name in ((1, "head_1"), (2, "head_2"), (4, "head_4")):
        c = _one(layers[name], "convolution_param")
        assert (_one(c, "dilation", 1), _one(c, "pad"), _one(c, "kernel_size"), _one(c, "num_output")) == (d, d, 3, 128)
        assert [_one(p, "name") for p in _all(layers[name], "param")] == ["head_w", "head_b"]          # shared weights
    for n in (1, 2, 4):
        assert _one(_one(layers["cls_score_%d" % n], "convolution_param"), "num_output") == 2
        assert _one(_one(layers["bbox_pred_%d" % n], "convolution_param"), "num_output") == 4
    assert _one(_one(layers["cls_score_output_concat"], "concat_param"), "axis") == 2
    assert _one(_one(layers["bbox_pred_output_concat"], "concat_param"), "axis") == 1
    assert [int(d) for d in _all(_one(_one(layers["cls_prob_reshape"], "reshape_param"), "shape"), "dim")] == [0, 6, -1, 0]
    # only head_1 of the sharing set is stored in the caffemodel's sharing set order; all three use the same array
    stored = read_caffemodel(model)
    inet = IndepNet(proto, model)
    w = [inet._param(layers[n], 0, (128, 128, 3, 3)) for n in ("head_1", "head_2", "head_4")]
    assert w[0] is w[1] is w[2] and np.abs(w[0]).max() > 0
    assert "conv4_fuse_final_dim_red" in stored and stored["conv1_1"][0][0] == [64, 3, 3, 3]


@pytest.mark.skipif(not os.path.exists(REF_PROTO), reason="reference tree not present (GPU box)")
def test_wire_field_numbers_match_the_reference_schema():
    """The numbers restated in oracle/indep_net.py against the reference's caffe.proto."""
    text = open(REF_PROTO).read()

    def block(name):
        m = re.search(r"message %s \{(.*?)\n\}" % name, text, re.S)
        return m.group(1)
    assert re.search(r"repeated LayerParameter layer = 100;", block("NetParameter"))
    lp = block("LayerParameter")
    assert re.search(r"optional string name = 1;", lp) and re.search(r"repeated BlobProto blobs = 7;", lp)
    bp = block("BlobProto")
    for pat in (r"optional BlobShape shape = 7;", r"repeated float data = 5 \[packed = true\];", r"optional int32 num = 1",
                r"optional int32 channels = 2", r"optional int32 height = 3", r"optional int32 width = 4"):
        assert re.search(pat, bp), pat
    assert re.search(r"repeated int64 dim = 1 \[packed = true\];", block("BlobShape"))


@pytest.mark.skipif(not os.path.exists("/root/reference/models"), reason="reference tree not present (GPU box)")
def test_independent_reader_parses_the_reference_templates():
    """The schema-less reader sees the same layer list in the reference's own template files as the product parser."""
    from smallhardface_b200 import caffe_proto as cp
    for fn in ("test_template.prototxt", "test_different_dilation_template.prototxt"):
        path = os.path.join("/root/reference/models", fn)
        mine = parse_prototxt(open(path).read())
        theirs = cp.read_net_text(path)
        assert [(_one(l, "name"), _one(l, "type"), _all(l, "bottom"), _all(l, "top")) for l in _all(mine, "layer")] == \
               [(l.name, l.type, list(l.bottom), list(l.top)) for l in theirs.layer]
