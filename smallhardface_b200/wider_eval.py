"""WIDER FACE evaluation -- ``lib/wider_eval_tools/wider_eval.py:10-222`` restated for Python 3 with the
O(detections x faces) IoU work on the GPU.

What the reference does per difficulty setting and image: a Python loop over the detections in descending score order,
each computing its IoU with every ground-truth face (``_boxoverlap``), matching it to ``argmax`` and updating two state
lists (``_image_evaluation``, ``:78-101``); then a loop over 1000 score thresholds with ``np.where`` over the detections
(``_image_pr_info``, ``:104-117``).  Here:

* the (faces x detections) IoU matrix of an image is ONE ``shf_bbox_overlaps`` launch (float64, the ``+1`` pixel
  convention -- the same arithmetic as ``_boxoverlap``: ``inter / (area_a + area_b - inter)``, 0 when the boxes do not
  overlap), shared by the three settings, which differ only in which faces count;
* the matching state machine is closed form: a detection's match depends only on its own IoU column; a counted face turns
  "recalled" at its FIRST matching detection, so ``pred_recall`` is a prefix count of those first indices;
* the 1000-threshold sweep is a suffix-maximum + ``searchsorted`` (no assumption that the scores are sorted).

Bit-compatible quirks that are kept (they change the published numbers): ``mimic_eval_bug`` rounds every IoU to 0 / 1 BEFORE
the argmax (``:90-92``; Python 2 ``round``: halves away from zero, so IoU == 0.5 matches), which makes a detection match
the FIRST face with IoU >= 0.5 rather than the best one; scores are min-max normalised over the whole set
(``_norm_score``); faces outside a setting's ``gt_list`` absorb detections (``proposal_list = -1``) without counting.

The IoU provider is injectable (``overlaps=``): the default is the CUDA kernel through the ``utils.cython_bbox`` drop-in and
raises without a GPU; the CPU tests pass the oracle's NumPy restatement of ``bbox.pyx``.
"""
from __future__ import annotations

import copy
import logging
import os
from functools import reduce
from typing import Callable, List, Optional

import numpy as np

logger = logging.getLogger(__name__)

EVENT_NUM = 61          # wider_eval.py:12,44,137: the WIDER FACE event count is hard-coded
THRESH_NUM = 1000       # :138


def _gpu_overlaps(boxes: np.ndarray, query: np.ndarray) -> np.ndarray:
    from .compat.cython_bbox import bbox_overlaps
    return bbox_overlaps(np.ascontiguousarray(boxes, dtype=np.float64), np.ascontiguousarray(query, dtype=np.float64))


def py2_round(x: np.ndarray) -> np.ndarray:
    """Python 2 ``round(x)`` (C ``round``: halves away from zero) for x >= 0."""
    f = np.floor(x)
    return np.where(x - f >= 0.5, f + 1.0, f)


def load_gt(path: str):
    from scipy import io as sio
    return sio.loadmat(path)


def read_pred(pred_dir: str, gt_data, strict: bool = False):
    """``_read_pred`` (``:10-39``): per event, per image an (n, 5) ``[x, y, w, h, score]`` array sorted by descending score;
    ``None`` for a file that is missing or malformed (the reference logs and carries on; ``strict`` raises instead)."""
    pred_list = [None] * EVENT_NUM
    for i in range(EVENT_NUM):
        img_list = gt_data["file_list"][i][0]
        bbx_list: List[Optional[np.ndarray]] = [None] * img_list.shape[0]
        event = gt_data["event_list"][i][0][0]
        for j in range(img_list.shape[0]):
            fname = "{:s}/{:s}/{:s}.txt".format(pred_dir, event, img_list[j][0][0])
            try:
                with open(fname, "r") as f:
                    tmp = [x.strip() for x in f.readlines()]
                bbx_num = int(tmp[1])
                bbx = np.zeros((bbx_num, 5))
                for k in range(bbx_num):
                    bbx[k] = [float(x) for x in tmp[k + 2].split()]
                bbx_list[j] = bbx[bbx[:, -1].argsort()[::-1]]
            except Exception:
                if strict:
                    raise
                logger.error("Fail to parse the prediction file {:s} {:s}".format(event, img_list[j][0][0]))
        pred_list[i] = bbx_list
    return pred_list


def norm_score(org_pred_list):
    """``_norm_score`` (``:42-57``): min-max normalisation of the scores over the whole set (in place, like the reference)."""
    max_score, min_score = 0.0, np.inf
    for i in range(EVENT_NUM):
        allp = np.vstack(org_pred_list[i])
        max_score = max(max_score, np.max(allp[:, -1]))
        min_score = min(min_score, np.min(allp[:, -1]))
    norm = [None] * EVENT_NUM
    for i in range(EVENT_NUM):
        pred_list_i = copy.copy(org_pred_list[i])
        for j in range(len(pred_list_i)):
            pred_list_i[j][:, -1] -= min_score
            pred_list_i[j][:, -1] /= (max_score - min_score)
        norm[i] = pred_list_i
    return norm


def image_overlaps(pred_info: np.ndarray, gt_bbx: np.ndarray, overlaps: Callable) -> np.ndarray:
    """(faces, detections) IoU of one image; inputs are ``[x, y, w, h, ...]`` rows (``:83-86`` turn them into corners)."""
    p = np.array(pred_info[:, :4], dtype=np.float64)
    g = np.array(gt_bbx[:, :4], dtype=np.float64)
    p[:, 2] += p[:, 0]
    p[:, 3] += p[:, 1]
    g[:, 2] += g[:, 0]
    g[:, 3] += g[:, 1]
    return overlaps(g, p)


def image_evaluation(ov: np.ndarray, ignore: np.ndarray, iou_thresh: float, mimic_eval_bug: bool):
    """``_image_evaluation`` (``:78-101``) from the (faces, detections) IoU matrix.  ``ignore[g] == 1`` marks a face that
    COUNTS in this setting (the reference's naming).  Returns ``pred_recall`` (P, 1) and ``proposal_list`` (P, 1)."""
    G, P = ov.shape
    if mimic_eval_bug:
        ov = py2_round(ov)
    idx = np.argmax(ov, axis=0)                              # first maximum, like np.argmax in the loop
    hit = ov[idx, np.arange(P)] >= iou_thresh
    counted = ignore.reshape(-1)[idx] != 0
    proposal_list = np.ones((P, 1))
    proposal_list[hit & ~counted] = -1
    first = np.full(G, P, dtype=np.int64)                    # first detection matched to each counted face
    sel = np.nonzero(hit & counted)[0]
    np.minimum.at(first, idx[sel], sel)
    pred_recall = np.cumsum(np.bincount(first[first < P], minlength=P)).astype(np.float64).reshape(P, 1)
    return pred_recall, proposal_list


def image_pr_info(thresh_num: int, scores: np.ndarray, proposal_list: np.ndarray, pred_recall: np.ndarray) -> np.ndarray:
    """``_image_pr_info`` (``:104-117``): per threshold ``1 - (t + 1) / thresh_num`` the number of counted proposals and the
    recalled faces among the detections up to the LAST one whose score reaches the threshold."""
    P = scores.shape[0]
    thresh = 1 - (np.arange(thresh_num) + 1.0) / thresh_num
    suffix_max = np.maximum.accumulate(scores[::-1])[::-1]   # non-increasing; #{h: suffix_max[h] >= t} - 1 = last index
    n_ge = P - np.searchsorted(suffix_max[::-1], thresh, side="left")
    r_index = n_ge - 1
    cum_prop = np.cumsum(proposal_list.reshape(-1) == 1)
    info = np.zeros((thresh_num, 2))
    ok = r_index >= 0
    info[ok, 0] = cum_prop[r_index[ok]]
    info[ok, 1] = pred_recall.reshape(-1)[r_index[ok]]
    return info


def dataset_pr_info(thresh_num: int, org_pr_curve: np.ndarray, count_face: int) -> np.ndarray:
    """``_dataset_pr_info`` (``:120-127``); 0/0 stays NaN like the reference's element-wise division."""
    pr = np.zeros((thresh_num, 2))
    with np.errstate(divide="ignore", invalid="ignore"):
        pr[:, 0] = org_pr_curve[:, 1] / org_pr_curve[:, 0]
        pr[:, 1] = org_pr_curve[:, 1] / count_face
    return pr


def voc_ap(rec: np.ndarray, prec: np.ndarray) -> float:
    """``_VOCap`` (``:130-136``)."""
    mrec = np.hstack([0, rec, 1])
    mpre = np.hstack([0, prec, 0])
    for i in range(mpre.shape[0] - 2, -1, -1):
        mpre[i] = max(mpre[i], mpre[i + 1])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))


def evaluation(norm_pred_list, setting_gt, iou_thresh: float, mimic_eval_bug: bool, ov_cache: dict, overlaps: Callable) -> np.ndarray:
    """``_evaluation`` (``:139-176``) for one difficulty setting.  ``ov_cache`` shares the IoU matrices between settings."""
    org_pr_curve = np.zeros((THRESH_NUM, 2))
    count_face = 0
    img_list = np.vstack([_[0] for _ in setting_gt["file_list"]])
    gt_bbx_list = np.vstack([_[0] for _ in setting_gt["face_bbx_list"]])
    pred_list = reduce(lambda x, y: x + y, norm_pred_list)
    sub_gt_list = np.vstack([_[0] for _ in setting_gt["gt_list"]])
    for j in range(img_list.shape[0]):
        gt_bbx = gt_bbx_list[j][0]
        pred_info = pred_list[j]
        keep_index = sub_gt_list[j][0] - 1
        count_face += keep_index.shape[0]
        if gt_bbx.size == 0 or pred_info.size == 0:
            continue
        ignore = np.zeros((gt_bbx.shape[0], 1))
        if keep_index.size > 0:
            ignore[keep_index] = 1
        ov = ov_cache.get(j)
        if ov is None:
            ov = ov_cache[j] = image_overlaps(pred_info, gt_bbx, overlaps)
        pred_recall, proposal_list = image_evaluation(ov, ignore, iou_thresh, mimic_eval_bug)
        org_pr_curve += image_pr_info(THRESH_NUM, pred_info[:, -1], proposal_list, pred_recall)
    return dataset_pr_info(THRESH_NUM, org_pr_curve, count_face)


def wider_eval(pred_dir: str, gt_dir_base: str, mimic_eval_bug: bool = True, IoU_thresh: float = 0.5,
               overlaps: Optional[Callable] = None):
    """``wider_eval`` (``:179-222``): returns ``(ap, pr_curve)`` for the easy / medium / hard validation settings."""
    overlaps = overlaps or _gpu_overlaps
    gt_data = load_gt("{:s}/wider_face_val.mat".format(gt_dir_base))
    norm_pred_list = norm_score(read_pred(pred_dir, gt_data))
    ov_cache: dict = {}
    pr_curve, ap = [], []
    for setting in ("easy_val", "medium_val", "hard_val"):
        setting_gt = load_gt("{:s}/wider_{:s}.mat".format(gt_dir_base, setting))
        pr = evaluation(norm_pred_list, setting_gt, IoU_thresh, mimic_eval_bug, ov_cache, overlaps)
        pr_curve.append(pr)
        ap.append(voc_ap(pr[:, 1], pr[:, 0]))
    return ap, pr_curve


def format_result(ap) -> str:
    """The line ``lib/datasets/wider.py:193`` logs."""
    return "Easy: {:.4f}, Medium: {:.4f}, Hard: {:.4f}".format(*ap)


if __name__ == "__main__":          # python -m smallhardface_b200.wider_eval <pred_dir> <ground_truth dir>
    import sys
    print(format_result(wider_eval(sys.argv[1], sys.argv[2])[0]))
    del os
