"""The reference-facing plugin surface on the B200: `caffe.Net` driven exactly the way lib/test.py:21-66 drives
pycaffe, plus the `nms` and `utils.cython_bbox` drop-in modules, against the oracle / reference golden vectors."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from oracle import detect as OD
from oracle import postprocess as OP
from oracle import preprocess as OPRE
from oracle.indep_net import IndepNet
from smallhardface_b200 import compat, deploy
from test_gpu_net import match_rows

compat.install()
import caffe                                    # noqa: E402  (our drop-in)
from nms.cpu_nms import cpu_nms                 # noqa: E402
from nms.gpu_nms import gpu_nms                 # noqa: E402
from nms.nms_wrapper import nms                 # noqa: E402
import smallhardface_b200.compat.cython_bbox as cython_bbox   # noqa: E402

F32 = np.float32


@pytest.fixture(scope="module")
def deployed(tmp_path_factory):
    d = tmp_path_factory.mktemp("deploy")
    return deploy.write_synthetic_deployment(str(d), dilation=True)


def forward_net_like_reference(net, blob, im_scale, flip=False):
    """lib/test.py:21-66, statement for statement (pyramid=True, single 'boxes' module)."""
    blob = dict(blob)
    blob["im_info"] = np.array([[blob["data"].shape[2], blob["data"].shape[3], im_scale]], dtype=np.float32)
    h, w = blob["data"].shape[2:]
    new_h = int(np.ceil(1.0 * h / 16) * 16)
    new_w = int(np.ceil(1.0 * w / 16) * 16)
    data = np.pad(blob["data"], ((0, 0), (0, 0), (0, new_h - h), (0, new_w - w)), "constant")
    net.blobs["data"].reshape(*(data.shape))
    net.blobs["im_info"].reshape(*(blob["im_info"].shape))
    net_args = {"data": data.astype(np.float32, copy=False), "im_info": blob["im_info"].astype(np.float32, copy=False)}
    blobs_out = net.forward(**net_args)
    if flip:
        for i in filter(lambda x: x.startswith("boxes"), blobs_out.keys()):
            blobs_out[i][:, [1, 3]] = w - blobs_out[i][:, [3, 1]]
    assert "boxes" in net.blobs
    cur_boxes = net.blobs["boxes"].data
    cur_boxes = cur_boxes[:, 1:5] / im_scale
    cur_probs = net.blobs["cls_prob"].data
    return cur_probs, np.tile(cur_boxes, (1, cur_probs.shape[1]))


def test_net_surface_and_forward_like_lib_test(deployed):
    proto, model = deployed
    caffe.set_mode_gpu()
    caffe.set_device(0)
    net = caffe.Net(str(proto), str(model), caffe.TEST)
    assert net.inputs == ["data", "im_info"] and net.outputs == ["boxes", "cls_prob"]
    assert list(net.blobs.keys())[:3] == ["data", "im_info", "conv1_1"]
    assert not any(k.startswith("boxes") and k != "boxes" for k in net.blobs)
    assert net.params["head_1"][0].data.shape == (128, 128, 3, 3)
    assert net.blobs["data"].reshape(1, 3, 64, 96) is None and net.blobs["data"].data.shape == (1, 3, 64, 96)
    with pytest.raises(Exception, match="Input blob arguments do not match net inputs"):
        net.forward(data=np.zeros((1, 3, 64, 96), F32))
    with pytest.raises(Exception, match="Input is not batch sized"):
        net.forward(data=np.zeros((2, 3, 64, 96), F32), im_info=np.zeros((1, 3), F32))
    with pytest.raises(RuntimeError, match="Could not open file"):
        caffe.Net("/nonexistent.prototxt", str(model), caffe.TEST)

    onet = IndepNet(proto, model, engine="sgemm")
    im = deploy.synthetic_image(7, (90, 140))
    scales = OPRE.pyramid_scales(im.shape, (300, 600))
    blobs = OPRE.get_image_blobs(im, scales)
    worst = 0.0
    for blob, s in zip(blobs, scales):
        for flip in (False, True):
            d = np.ascontiguousarray(blob[..., ::-1]) if flip else blob
            probs, boxes = forward_net_like_reference(net, {"data": d}, s, flip)
            rp, rb = OD.forward_level(onet, d, s, flip)
            # every reference row has a device row within 1e-3 (score) / 1e-2 raw px (box); rank swaps between
            # near-equal scores and rows straddling SCORE_THRESH are the only freedoms (match_rows)
            assert boxes.shape[0] == probs.shape[0]
            ws, wb = match_rows(boxes[:, :4], probs[:, 1], rb[:, :4], rp[:, 1])
            assert ws < 1e-3 and wb < 1e-2, (s, flip, ws, wb)
            worst = max(worst, wb)
    print("plugin surface: worst box error %.2e raw px" % worst)
    # views alias blob storage: an in-place edit is visible through net.blobs (what the flip fix relies on)
    out = net.forward(data=net.blobs["data"].data.copy(), im_info=net.blobs["im_info"].data.copy())
    out["boxes"][:, 1] = -7
    assert np.all(net.blobs["boxes"].data[:, 1] == -7)
    # an intermediate blob is readable after the forward (lazy device->host sync) and matches the oracle
    got = net.blobs["conv5_3"].data
    want = onet.blobs["conv5_3"]
    # this level has im_scale > 0.9, i.e. it ran on the fast f16+f8 operand format (2^-15-class operands)
    assert float(net.blobs["im_info"].data[0, 2]) > 0.9
    assert got.shape == want.shape and 1e-6 < np.abs(got - want).max() / np.abs(want).max() < 5e-4
    caffe.set_fast_min_scale(None)
    try:
        pnet = caffe.Net(str(proto), str(model), caffe.TEST)
        pnet.blobs["data"].reshape(*net.blobs["data"].data.shape)
        pnet.forward(data=net.blobs["data"].data.copy(), im_info=net.blobs["im_info"].data.copy())
        got = pnet.blobs["conv5_3"].data
        assert np.abs(got - want).max() / np.abs(want).max() < 2e-5          # split-fp16 operands on request
    finally:
        caffe.set_fast_min_scale(0.9)
    with pytest.raises(Exception, match="fused"):
        net.blobs["cls_prob_output"].data


def test_nms_modules_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "nms.npz"))
    for tag in ("n300", "n1500", "grid"):
        d = g[tag + "_dets"]
        assert [int(i) for i in cpu_nms(d, 0.4)] == g[tag + "_cpu_0.4"].tolist()
        assert [int(i) for i in gpu_nms(d, 0.4, device_id=0)] == OP.nms(d, 0.4, OP.NMS_GPU)
        assert [int(i) for i in nms(d, 0.4, force_cpu=True)] == g[tag + "_cpu_0.4"].tolist()
    assert nms(np.zeros((0, 5), F32), 0.4) == []
    assert isinstance(cpu_nms(g["n300_dets"], 0.7), list)
    with pytest.raises(ValueError):
        cpu_nms(g["n300_dets"].astype(np.float64), 0.4)


def test_cython_bbox_module_against_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "bbox_overlaps.npz"))
    b, q = g["boxes"], g["query"]
    assert np.array_equal(cython_bbox.bbox_overlaps(b, q), g["iou"])
    assert np.array_equal(cython_bbox.bbox_overlaps_IoA(b, q), g["ioa"])
    assert np.array_equal(cython_bbox.bbox_overlaps_itself(b, q), g["itself"])
    assert np.array_equal(cython_bbox.bbox_overlaps_IoA(b, b), g["ioa_sq"])
    assert cython_bbox.bbox_overlaps(np.zeros((0, 4)), q).shape == (0, len(q))
    with pytest.raises(ValueError):
        cython_bbox.bbox_overlaps(b.astype(np.float32), q)
