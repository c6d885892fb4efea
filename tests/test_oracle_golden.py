"""Pin the oracle against golden vectors produced by the reference's OWN code
(tests/golden/make_golden.py: proposal_layer.py, generate_anchors.py, bbox_transform.py,
py_cpu_nms.py, cpu_nms.pyx, bbox.pyx, lib/test.py:bbox_vote)."""
import os

import numpy as np
import pytest

from oracle import postprocess as post
from oracle import proposal as prop


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name), allow_pickle=False)


def test_anchors(golden_dir):
    g = _load(golden_dir, "anchors.npz")
    a = prop.generate_anchors(base_size=16, ratios=(1,), scales=(1, 2, 4), shifts=(0,), strides=(8, 8, 8))
    assert np.array_equal(a, g["anchors"])
    assert a.tolist() == [[0, 0, 15, 15], [-8, -8, 23, 23], [-24, -24, 39, 39]]
    assert np.array_equal(prop.generate_anchors(strides=(16, 16, 16)), g["default"])


@pytest.mark.parametrize("tag", ["small", "level", "allbelow", "wide"])
def test_proposal_forward_bit_exact(golden_dir, tag):
    g = _load(golden_dir, "proposal.npz")
    boxes, probs, _ = prop.proposal_forward(g[tag + "_cls"], g[tag + "_deltas"], g[tag + "_im_info"])
    assert boxes.dtype == np.float32 and probs.dtype == np.float32
    assert np.array_equal(boxes, g[tag + "_boxes"])
    assert np.array_equal(probs, g[tag + "_probs"])


def test_proposal_micro_case_zero_deltas():
    """Hand-derived: zero deltas return the anchors (x2/y2 + 1 because decode has no -1), clipped."""
    h, w = 2, 3
    cls = np.zeros((1, 6, h, w), np.float32)
    cls[0, 3:] = np.linspace(0.9, 0.1, 3 * h * w).reshape(3, h, w)
    cls[0, :3] = 1 - cls[0, 3:]
    boxes, probs, order = prop.proposal_forward(cls, np.zeros((1, 12, h, w), np.float32),
                                                np.array([[16, 24, 1.0]], np.float32))
    assert boxes.shape == (18, 5) and order[0] == 0
    assert boxes[0].tolist() == [0, 0, 0, 16, 15]            # anchor 0 at cell (0,0): [0,0,16,16], y2 clipped to h-1
    top_a1 = boxes[np.where(order == 1)[0][0]]                # anchor 1 at cell (0,0): [-8,-8,24,24] clipped
    assert top_a1.tolist() == [0, 0, 0, 23, 15]


@pytest.mark.parametrize("tag", ["n300", "n1", "n1500", "grid"])
def test_nms_matches_reference_cython_and_python(golden_dir, tag):
    g = _load(golden_dir, "nms.npz")
    d = g[tag + "_dets"]
    for thr in ((0.4, 0.5, 0.7) if tag == "grid" else (0.4, 0.7, 0.3)):
        assert post.nms(d, thr, post.NMS_CPU) == g["%s_cpu_%g" % (tag, thr)].tolist()
        assert post.nms(d, thr, post.NMS_PY) == g["%s_py_%g" % (tag, thr)].tolist()


def test_nms_ge_vs_gt_semantics(golden_dir):
    g = _load(golden_dir, "nms.npz")
    d = g["grid_dets"]
    # IoU([0,0,9,9],[0,0,9,3]) = 40/100 = 0.4 exactly in float32 -> float32(0.4) >= 0.4 (double): suppressed by
    # cpu_nms (>=), kept by the GPU kernel's '>' (nms_kernel.cu:82)
    assert 1 not in post.nms(d, 0.4, post.NMS_CPU)
    assert 1 in post.nms(d, 0.4, post.NMS_GPU)
    assert post.nms(np.zeros((0, 5), np.float32), 0.4) == []


@pytest.mark.parametrize("tag", ["n0", "n1", "n2far", "n400", "n3000"])
def test_bbox_vote(golden_dir, tag):
    g = _load(golden_dir, "bbox_vote.npz")
    r = post.bbox_vote(g[tag + "_dets"].copy(), 0.4)
    assert str(r.dtype) == str(g[tag + "_vote_dtype"])
    assert r.shape == g[tag + "_vote"].shape
    assert np.array_equal(r, g[tag + "_vote"])


def test_bbox_vote_drops_singletons(golden_dir):
    g = _load(golden_dir, "bbox_vote.npz")
    # two far-apart boxes: the first (a singleton with rows remaining) is dropped, the last kept
    assert np.array_equal(g["n2far_vote"], g["n2far_dets"][1:2])


def test_bbox_overlaps(golden_dir):
    g = _load(golden_dir, "bbox_overlaps.npz")
    b, q = g["boxes"], g["query"]
    assert np.array_equal(post.bbox_overlaps(b, q, "iou"), g["iou"])
    assert np.array_equal(post.bbox_overlaps(b, q, "ioa"), g["ioa"])
    assert np.array_equal(post.bbox_overlaps(b, q, "itself"), g["itself"])
    assert np.array_equal(post.bbox_overlaps(b, b, "ioa"), g["ioa_sq"])
    with pytest.raises(ValueError):
        post.bbox_overlaps(b.astype(np.float32), q)
