"""ORACLE (test infrastructure, not product code) -- greedy NMS, box voting and the IoU matrix
helpers, restating ``lib/nms/cpu_nms.pyx:17-68``, ``lib/nms/py_cpu_nms.py:10-38``,
``lib/nms/nms_kernel.cu:24-32,45-155`` (+ ``gpu_nms.pyx:11-30``), ``lib/test.py:181-217`` and
``lib/utils/bbox.pyx:14-142``.

The ``*_c`` variants call the plain-C restatement in oracle/c/oracle_post.c (built into
oracle/_build/liboracle.so by ``oracle.build()``); the numpy versions here are the readable form
and the two are checked against each other and against golden vectors made by the reference's own
``py_cpu_nms.py`` / ``bbox_vote`` (tests/golden/make_golden.py).

Tie order: every ``argsort()[::-1]`` of the reference is restated as a stable descending sort
(ties -> lower index first); see oracle/proposal.py.
"""
from __future__ import annotations

import numpy as np

from .proposal import stable_desc_order

F32 = np.float32

NMS_CPU = 0      # cpu_nms.pyx:65   suppress when (double)ovr >= thresh
NMS_GPU = 1      # nms_kernel.cu:82 suppress when ovr > (float)thresh
NMS_PY = 2       # py_cpu_nms.py:35 suppress when ovr > thresh (float32 vs python float -> float32 compare)


def _iou_row(boxes, areas, i, js):
    """float32 arithmetic exactly as ``cpu_nms.pyx:55-64`` / ``lib/test.py:188-197``."""
    xx1 = np.maximum(boxes[i, 0], boxes[js, 0])
    yy1 = np.maximum(boxes[i, 1], boxes[js, 1])
    xx2 = np.minimum(boxes[i, 2], boxes[js, 2])
    yy2 = np.minimum(boxes[i, 3], boxes[js, 3])
    w = np.maximum(F32(0.0), xx2 - xx1 + F32(1))
    h = np.maximum(F32(0.0), yy2 - yy1 + F32(1))
    inter = w * h
    return inter / (areas[i] + areas[js] - inter)


def nms(dets, thresh, mode=NMS_CPU):
    """Greedy NMS over (N,5) float32 ``[x1,y1,x2,y2,score]``; returns indices into ``dets`` in
    kept (descending score) order, like ``cpu_nms`` / ``gpu_nms``."""
    dets = np.ascontiguousarray(dets, dtype=F32)
    n = dets.shape[0]
    if n == 0:
        return []
    areas = (dets[:, 2] - dets[:, 0] + F32(1)) * (dets[:, 3] - dets[:, 1] + F32(1))
    order = stable_desc_order(dets[:, 4])
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        keep.append(int(i))
        js = order[_i + 1:]
        if js.size == 0:
            break
        ovr = _iou_row(dets, areas, i, js)
        if mode == NMS_CPU:
            hit = ovr.astype(np.float64) >= float(thresh)
        elif mode == NMS_GPU:
            hit = ovr > F32(thresh)
        else:
            hit = ovr > F32(thresh)
        suppressed[js[hit]] = True
    return keep


def bbox_vote(det, nms_thresh=0.4):
    """``lib/test.py:181-217``: score-sorted greedy clusters (IoU >= thresh with the current top,
    float32 compare); singleton clusters are dropped unless nothing remains; a cluster becomes its
    score-weighted mean box with the max score.  Output dtype follows NumPy promotion in the
    reference (float64 as soon as one merged cluster is emitted)."""
    det = np.asarray(det)
    order = stable_desc_order(det[:, 4].ravel())
    det = det[order, :]
    dets = None
    if det.shape[0] == 0:
        dets = np.array([[10, 10, 20, 20, 0.0001]])
        det = np.empty(shape=[0, 5])
    thr = F32(nms_thresh) if det.dtype == F32 else nms_thresh
    while det.shape[0] > 0:
        area = (det[:, 2] - det[:, 0] + 1) * (det[:, 3] - det[:, 1] + 1)
        xx1 = np.maximum(det[0, 0], det[:, 0])
        yy1 = np.maximum(det[0, 1], det[:, 1])
        xx2 = np.minimum(det[0, 2], det[:, 2])
        yy2 = np.minimum(det[0, 3], det[:, 3])
        w = np.maximum(0.0, xx2 - xx1 + 1)
        h = np.maximum(0.0, yy2 - yy1 + 1)
        inter = w * h
        o = inter / (area[0] + area[:] - inter)
        merge_index = np.where(o >= thr)[0]
        det_accu = det[merge_index, :]
        det = np.delete(det, merge_index, 0)
        if merge_index.shape[0] <= 1:
            if det.shape[0] == 0:
                dets = det_accu if dets is None else np.vstack((dets, det_accu))
            continue
        det_accu[:, 0:4] = det_accu[:, 0:4] * np.tile(det_accu[:, -1:], (1, 4))
        max_score = np.max(det_accu[:, 4])
        det_accu_sum = np.zeros((1, 5))
        det_accu_sum[:, 0:4] = np.sum(det_accu[:, 0:4], axis=0) / np.sum(det_accu[:, -1:])
        det_accu_sum[:, 4] = max_score
        dets = det_accu_sum if dets is None else np.vstack((dets, det_accu_sum))
    return dets


def bbox_overlaps(boxes, query, kind="iou"):
    """``lib/utils/bbox.pyx``: ``iou`` :14-54, ``ioa`` :56-102 (diagonal zeroed), ``itself`` :106-142.
    float64 (N,4),(K,4) -> (N,K)."""
    boxes = np.asarray(boxes)
    query = np.asarray(query)
    if boxes.dtype != np.float64 or query.dtype != np.float64:
        raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' (float64)")
    n, k = boxes.shape[0], query.shape[0]
    out = np.zeros((n, k), dtype=np.float64)
    if n == 0 or k == 0:
        return out
    qa = (query[:, 2] - query[:, 0] + 1) * (query[:, 3] - query[:, 1] + 1)
    ba = (boxes[:, 2] - boxes[:, 0] + 1) * (boxes[:, 3] - boxes[:, 1] + 1)
    iw = np.minimum(boxes[:, None, 2], query[None, :, 2]) - np.maximum(boxes[:, None, 0], query[None, :, 0]) + 1
    ih = np.minimum(boxes[:, None, 3], query[None, :, 3]) - np.maximum(boxes[:, None, 1], query[None, :, 1]) + 1
    ok = (iw > 0) & (ih > 0)
    inter = iw * ih
    if kind == "iou":
        ua = ba[:, None] + qa[None, :] - inter
        vals = inter / np.where(ok, ua, 1.0)
    else:
        vals = inter / ba[:, None]
    out[ok] = vals[ok]
    if kind == "ioa":
        m = min(n, k)
        out[np.arange(m), np.arange(m)] = 0
    return out
