"""pycaffe surface beyond what lib/test.py touches: the generic `caffe.Layer` protocol
(caffe/include/caffe/layers/python_layer.hpp:19-43; KATs of caffe/python/caffe/test/test_python_layer.py:9-57,103-168) and
`net.forward(start=, end=)` (caffe/python/caffe/pycaffe.py:88-134; KAT of caffe/python/caffe/test/test_net.py:74-90)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from smallhardface_b200 import compat, deploy

compat.install()
import caffe                                    # noqa: E402

F32 = np.float32

PYTHON_NET = """name: 'pythonnet' force_backward: true
input: 'data' input_shape { dim: 10 dim: 9 dim: 8 }
layer { type: 'Python' name: 'one' bottom: 'data' top: 'one'
  python_param { module: 'shf_pylayers' layer: 'SimpleLayer' } }
layer { type: 'Python' name: 'two' bottom: 'one' top: 'two'
  python_param { module: 'shf_pylayers' layer: 'SimpleLayer' } }
layer { type: 'Python' name: 'three' bottom: 'two' top: 'three'
  python_param { module: 'shf_pylayers' layer: 'SimpleLayer' } }"""


def _write(tmp_path, name, text):
    p = tmp_path / name
    p.write_text(text)
    return str(p)


def test_python_layer_forward_and_reshape(tmp_path):
    assert "Python" in caffe.layer_type_list()
    net = caffe.Net(_write(tmp_path, "py.prototxt", PYTHON_NET), caffe.TEST)
    net.blobs["data"].data[...] = 8
    out = net.forward()
    assert list(out) == ["three"] and out["three"].shape == (10, 9, 8)
    assert np.all(net.blobs["three"].data == 10 ** 3 * 8) and np.all(net.blobs["one"].data == 80)
    net.blobs["data"].reshape(4, 4, 4, 4)                       # test_python_layer.py:121-127
    net.blobs["data"].data[...] = 1
    net.forward()
    for blob in net.blobs.values():
        assert blob.data.shape == (4, 4, 4, 4)
    assert np.all(net.blobs["three"].data == 1000)


def test_python_layer_exception_phase_param_str_and_blobs(tmp_path):
    exc = PYTHON_NET.split("layer {")[0] + """layer { type: 'Python' name: 'layer' bottom: 'data' top: 'top'
      python_param { module: 'shf_pylayers' layer: 'ExceptionLayer' } }"""
    with pytest.raises(RuntimeError):
        caffe.Net(_write(tmp_path, "exc.prototxt", exc), caffe.TEST)
    phase = """name: 'pythonnet' layer { type: 'Python' name: 'layer' top: 'phase'
      python_param { module: 'shf_pylayers' layer: 'PhaseLayer' } }"""
    net = caffe.Net(_write(tmp_path, "phase.prototxt", phase), caffe.TEST)
    assert net.forward()["phase"] == caffe.TEST
    ps = PYTHON_NET.split("layer {")[0] + """layer { type: 'Python' name: 'add' bottom: 'data' top: 'sum'
      python_param { module: 'shf_pylayers' layer: 'ParamStrLayer' param_str: '2.5' } }"""
    net = caffe.Net(_write(tmp_path, "ps.prototxt", ps), caffe.TEST)
    net.blobs["data"].data[...] = 1
    assert np.all(net.forward()["sum"] == 3.5)
    par = PYTHON_NET.split("layer {")[0] + """layer { type: 'Python' name: 'layer' bottom: 'data' top: 'top'
      python_param { module: 'shf_pylayers' layer: 'ParameterLayer' } }"""
    net = caffe.Net(_write(tmp_path, "par.prototxt", par), caffe.TEST)
    net.forward()
    layer = net._spec.py_layers["layer"]["obj"]
    assert layer.blobs[0].data[0] == 0 and layer.blobs[0].data.shape == (1,)


CONV_PY_NET = """name: 'mixed'
input: 'data' input_shape { dim: 1 dim: 3 dim: 24 dim: 40 }
layer { name: 'c1' type: 'Convolution' bottom: 'data' top: 'c1' convolution_param { num_output: 64 kernel_size: 3 pad: 1 } }
layer { name: 'c1_relu' type: 'ReLU' bottom: 'c1' top: 'c1' }
%s
layer { name: 'c2' type: 'Convolution' bottom: '%s' top: 'c2' convolution_param { num_output: 64 kernel_size: 3 pad: 1 } }
layer { name: 'c2_relu' type: 'ReLU' bottom: 'c2' top: 'c2' }"""


def test_python_layer_between_tensor_core_convs(tmp_path):
    """conv -> Python (x10 on the host) -> conv equals conv -> conv with the second conv's weights x10: the device ->
    host -> device round trip of the generic protocol carries the activations faithfully."""
    from smallhardface_b200 import caffe_proto as cp
    py = "layer { type: 'Python' name: 'x10' bottom: 'c1' top: 'c1x' python_param { module: 'shf_pylayers' layer: 'SimpleLayer' } }"
    with_py = _write(tmp_path, "with.prototxt", CONV_PY_NET % (py, "c1x"))
    without = _write(tmp_path, "without.prototxt", CONV_PY_NET % ("", "c1"))
    rng = np.random.RandomState(4)
    w1, b1 = (rng.randn(64, 3, 3, 3) * 0.2).astype(F32), (rng.randn(64) * 0.1).astype(F32)
    w2, b2 = (rng.randn(64, 64, 3, 3) * 0.05).astype(F32), (rng.randn(64) * 0.1).astype(F32)

    def model(path, w2_):
        net = cp.Msg("NetParameter", name="m")
        for name, blobs in (("c1", [w1, b1]), ("c2", [w2_, b2])):
            lay = cp.Msg("LayerParameter", name=name, type="Convolution")
            lay.blobs = [cp.blob_from_array(a) for a in blobs]
            net.layer.append(lay)
        cp.write_net_binary(path, net)
        return path
    caffe.set_fast_min_scale(None)
    try:
        a = caffe.Net(with_py, model(str(tmp_path / "a.caffemodel"), w2), caffe.TEST)
        b = caffe.Net(without, model(str(tmp_path / "b.caffemodel"), w2 * F32(10)), caffe.TEST)
    finally:
        caffe.set_fast_min_scale(0.9)
    x = (rng.rand(1, 3, 24, 40) * 255 - 110).astype(F32)
    ya, yb = a.forward(data=x)["c2"], b.forward(data=x)["c2"]
    assert ya.shape == yb.shape == (1, 64, 24, 40) and np.abs(yb).max() > 10
    assert np.abs(ya - yb).max() <= 2e-5 * np.abs(yb).max()
    assert np.allclose(a.blobs["c1x"].data, 10 * a.blobs["c1"].data)


def test_forward_start_end_on_a_launch_boundary(tmp_path):
    """test_net.py:74-90 (`forward(start='ip', end='ip')` against a NumPy product) on this net's 1x1 conv4_256 + ReLU: the
    caller writes the bottom blob through `.data`, the range runs alone, the top equals the NumPy result."""
    proto, model = deploy.write_synthetic_deployment(str(tmp_path), dilation=True)
    caffe.set_fast_min_scale(None)
    try:
        net = caffe.Net(proto, model, caffe.TEST)
    finally:
        caffe.set_fast_min_scale(0.9)
    net.blobs["data"].reshape(1, 3, 64, 96)
    net.blobs["im_info"].reshape(1, 3)
    net.forward(data=np.zeros((1, 3, 64, 96), F32), im_info=np.array([[64, 96, 1.0]], F32))
    conv_blob = net.blobs["conv4_3"]
    sample = np.random.RandomState(2).uniform(size=conv_blob.data.shape).astype(F32)
    conv_blob.data[:] = sample
    out = net.forward(start="conv4_256", end="conv4_256_relu")
    assert "conv4_256" in out
    w, b = net.params["conv4_256"][0].data[:, :, 0, 0], net.params["conv4_256"][1].data
    manual = np.maximum(np.einsum("oc,nchw->nohw", w, sample) + b[None, :, None, None], 0)
    np.testing.assert_allclose(net.blobs["conv4_256"].data, manual, rtol=1e-3, atol=1e-5)
    with pytest.raises(Exception, match="fused launch"):
        net.forward(start="conv4_256", end="conv4_256")           # conv and its in-place ReLU are one kernel
    with pytest.raises(ValueError):
        net.forward(start="no_such_layer")
