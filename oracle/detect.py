"""ORACLE (test infrastructure, not product code) -- the per-image driver, restating
``lib/test.py:21-106`` (forward_net) and ``lib/test.py:109-178`` (detect): image pyramid x flip,
un-mirror, unscale, concat, threshold 0.05, then bbox_vote or NMS.
"""
from __future__ import annotations

import numpy as np

from . import postprocess as post
from . import preprocess as pre

F32 = np.float32


def forward_level(net, data, im_scale, flip=False):
    """``lib/test.py:21-66`` for the single-module ('boxes' blob) case with pyramid=True.
    ``data`` is the unpadded (1,3,h,w) float32 blob (already mirrored when flip)."""
    h, w = data.shape[2:]
    im_info = np.array([[h, w, im_scale]], dtype=F32)
    padded = pre.pad_to_multiple(data)
    out = net.forward(data=padded.astype(F32, copy=False), im_info=im_info)
    boxes = out["boxes"]
    if flip:
        boxes[:, [1, 3]] = w - boxes[:, [3, 1]]                    # lib/test.py:52-54 (no -1)
    cur_boxes = boxes[:, 1:5] / im_scale                           # float32 / python float -> float32
    cur_probs = out["cls_prob"]
    return cur_probs.copy(), np.tile(cur_boxes, (1, cur_probs.shape[1]))[:, 0:4]


def detect_raw(net, im, scales=pre.TEST_SCALES, base=pre.PYRAMID_BASE_SIZE, flip=True, use_cv2=True):
    """Concatenated (probs (R,2), boxes (R,4)) over all pyramid passes, ``lib/test.py:125-158``."""
    pscales = pre.pyramid_scales(im.shape, scales, base)
    blobs = pre.get_image_blobs(im, pscales, use_cv2=use_cv2)
    all_p, all_b = [], []
    for blob, s in zip(blobs, pscales):
        p, b = forward_level(net, blob, s)
        all_p.append(p); all_b.append(b)
        if flip:
            p, b = forward_level(net, np.ascontiguousarray(blob[..., ::-1]), s, flip=True)
            all_p.append(p); all_b.append(b)
    return np.concatenate(all_p), np.concatenate(all_b)


def threshold_dets(probs, boxes, thresh=0.05):
    """``lib/test.py:162-167``: strict ``>`` in float32, dets = [boxes | score] float32."""
    inds = np.where(probs[:, 1] > F32(thresh))[0]
    return np.hstack((boxes[inds, :], probs[inds, 1][:, None])).astype(F32, copy=False)


def detect(net, im, thresh=0.05, nms_method="BBOX_VOTE", nms_thresh=0.4, **kw):
    probs, boxes = detect_raw(net, im, **kw)
    dets = threshold_dets(probs, boxes, thresh)
    if nms_method == "BBOX_VOTE":
        return post.bbox_vote(dets, nms_thresh)
    keep = post.nms(dets, nms_thresh, post.NMS_CPU)
    return dets[keep, :]
