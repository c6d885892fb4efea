"""Format codecs: prototxt text + caffemodel wire format (caffe.proto restated in caffe_proto.py)."""
import os

import numpy as np
import pytest

from smallhardface_b200 import caffe_proto as cp
from smallhardface_b200.graph import NetSpec, load_weights, reshape_shape, upgrade_net_input
from smallhardface_b200.models import build_test_net, splice_dim_red
from smallhardface_b200 import deploy

REF = "/root/reference/models"


def test_text_parse_quotes_comments_and_colon_brace():
    txt = """
    name: "face"   # a comment
    input: 'data'
    input_shape { dim: 1 dim: 3 dim: 224 dim: 224 }
    layer { name: 'up' type: "Deconvolution" bottom: "a" top: "b"
      convolution_param { kernel_size: 4 stride: 2 num_output: 256 group: 256 pad: 1
        weight_filler: { type: "bilinear" } bias_term: false }
      param { lr_mult: 0 decay_mult: 0 } }
    """
    net = cp.parse_text(txt)
    assert net.name == "face" and net.input == ["data"]
    assert net.input_shape[0].dim == [1, 3, 224, 224]
    l = net.layer[0]
    assert l.convolution_param.weight_filler.type == "bilinear"
    assert l.convolution_param.bias_term is False
    assert l.convolution_param.group == 256
    assert l.param[0].lr_mult == 0
    # defaults
    assert l.convolution_param.axis == 1 and l.convolution_param.engine == 0


def test_text_unknown_field_is_an_error():
    with pytest.raises(ValueError):
        cp.parse_text("layer { name: 'x' bogus_field: 3 }")


def test_text_roundtrip_both_templates():
    for dil in (False, True):
        net = build_test_net(dil)
        txt = cp.format_text(net)
        assert cp.format_text(cp.parse_text(txt)) == txt


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
@pytest.mark.parametrize("fname,dil", [("test_template.prototxt", False),
                                       ("test_different_dilation_template.prototxt", True)])
def test_builders_equal_reference_templates(fname, dil):
    ref = cp.read_net_text(os.path.join(REF, fname))
    assert cp.format_text(ref) == cp.format_text(build_test_net(dil))


def test_wire_known_bytes():
    # BlobShape{dim:[1,300]} packed varints; hand-encoded: field 1, wt 2, len 3, 01 AC 02
    assert cp.encode(cp.Msg("BlobShape", dim=[1, 300])) == bytes([0x0A, 0x03, 0x01, 0xAC, 0x02])
    bp = cp.blob_from_array(np.array([[1.5, -2.0]], np.float32))
    raw = cp.encode(bp)
    # data=5 packed float (tag 0x2A), shape=7 (tag 0x3A)
    assert raw[0] == 0x2A and raw[1] == 8 and np.frombuffer(raw[2:10], "<f4").tolist() == [1.5, -2.0]
    back = cp.decode(raw, "BlobProto")
    assert cp.array_from_blob(back).tolist() == [[1.5, -2.0]]


def test_wire_roundtrip_negative_and_unknown_fields():
    r = cp.Msg("ReshapeParameter", shape=cp.Msg("BlobShape", dim=[0, 2, -1, 0]), num_axes=-1)
    back = cp.decode(cp.encode(r), "ReshapeParameter")
    assert back.shape.dim == [0, 2, -1, 0] and back.num_axes == -1
    # unknown field 99 (varint) + unknown length-delimited field 98 are skipped
    raw = bytes([0x98, 0x06, 0x05]) + bytes([0x92, 0x06, 0x02, 0x41, 0x42]) + cp.encode(r)
    assert cp.decode(raw, "ReshapeParameter").shape.dim == [0, 2, -1, 0]


def test_protobuf_runtime_agrees_on_wire_format():
    """Cross-check the hand-written encoder against google.protobuf's own varint/packed encoding
    through a dynamically built descriptor of BlobProto."""
    pytest.importorskip("google.protobuf")
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="t.proto", package="t", syntax="proto2")
    m = fd.message_type.add(name="BlobShape")
    f = m.field.add(name="dim", number=1, type=3, label=3)
    f.options.packed = True
    m = fd.message_type.add(name="BlobProto")
    m.field.add(name="shape", number=7, type=11, label=1, type_name=".t.BlobShape")
    f = m.field.add(name="data", number=5, type=2, label=3)
    f.options.packed = True
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    cls = message_factory.GetMessageClass(pool.FindMessageTypeByName("t.BlobProto"))
    arr = np.random.RandomState(0).randn(3, 4).astype(np.float32)
    pb = cls()
    pb.shape.dim.extend(arr.shape)
    pb.data.extend(arr.ravel().tolist())
    assert pb.SerializeToString() == cp.encode(cp.blob_from_array(arr))
    assert np.array_equal(cp.array_from_blob(cp.decode(pb.SerializeToString(), "BlobProto")), arr)


def test_upgrade_legacy_input_and_outputs():
    net = upgrade_net_input(build_test_net(False))
    assert net.layer[0].type == "Input" and net.layer[0].top == ["data", "im_info"]
    spec = NetSpec(build_test_net(False))
    assert spec.inputs == ["data", "im_info"]
    assert spec.outputs == ["boxes", "cls_prob"]
    assert not any(n.startswith("boxes") and n != "boxes" for n in spec.blob_names)


def test_shapes_and_shared_params_dilation():
    spec = NetSpec(splice_dim_red(build_test_net(True)))
    sh = spec.infer_shapes({"data": (1, 3, 1408, 1408)})
    assert sh["conv4_3"] == (1, 512, 176, 176) and sh["conv5_3"] == (1, 512, 88, 88)
    assert sh["conv5_256_up"] == (1, 256, 176, 176) and sh["conv4_fuse"] == (1, 512, 176, 176)
    assert sh["conv4_fuse_final_tmp"] == (1, 512, 176, 176) and sh["conv4_fuse_final"] == (1, 128, 176, 176)
    assert sh["cls_score_reshape_output"] == (1, 2, 3 * 176, 176) and sh["bbox_pred_output"] == (1, 12, 176, 176)
    assert sh["cls_prob_reshape_output"] == (1, 6, 176, 176)
    by = {l.name: l for l in spec.layers}
    assert by["head_1"].param_keys == by["head_2"].param_keys == by["head_4"].param_keys
    assert by["conv5_256_up"].p["group"] == 256 and spec.param_shapes["conv5_256_up/0"] == (256, 1, 4, 4)


def test_concat_mismatch_raises_like_caffe():
    spec = NetSpec(build_test_net(False))
    with pytest.raises(ValueError):
        spec.infer_shapes({"data": (1, 3, 232, 232)})       # 232/16 not integral: 2*14 != 29


def test_reshape_semantics():      # test_reshape_layer.cpp:78-135
    assert reshape_shape((2, 3, 6, 5), [0, -1, 1, 0], 0, -1) == (2, 18, 1, 5)
    assert reshape_shape((2, 3, 6, 5), [0, 3, 10, -1], 0, -1) == (2, 3, 10, 3)
    assert reshape_shape((1, 6, 28, 28), [0, 2, -1, 0], 0, -1) == (1, 2, 84, 28)
    with pytest.raises(ValueError):
        reshape_shape((2, 3, 6, 5), [0, 7, -1, 0], 0, -1)


def test_caffemodel_roundtrip_and_load(tmp_path):
    proto, model = deploy.write_synthetic_deployment(str(tmp_path), dilation=True)
    net = cp.read_net_text(proto)
    spec = NetSpec(net)
    params = load_weights(spec, cp.read_net_binary(model))
    ref = deploy.synthetic_params(NetSpec(net))
    assert set(params) == set(ref)
    for k in ref:
        assert np.array_equal(params[k], ref[k]), k
    assert np.allclose(params["conv5_256_up/0"][7, 0], np.outer([.25, .75, .75, .25], [.25, .75, .75, .25]))


def test_load_shape_mismatch_is_fatal(tmp_path):
    _, model = deploy.write_synthetic_deployment(str(tmp_path), dilation=False)
    m = cp.read_net_binary(model)
    for l in m.layer:
        if l.name == "conv1_1":
            l.blobs[0] = cp.blob_from_array(np.zeros((64, 3, 5, 5), np.float32))
    with pytest.raises(RuntimeError, match="shape mismatch"):
        load_weights(NetSpec(build_test_net(False)), m)


def test_missing_file_message():
    with pytest.raises(RuntimeError, match="Could not open file"):
        cp.read_net_text("/nonexistent/x.prototxt")
