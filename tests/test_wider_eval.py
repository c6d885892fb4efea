"""WIDER evaluator (smallhardface_b200/wider_eval.py) against AP / PR curves computed by the reference's own
lib/wider_eval_tools/wider_eval.py on the same synthetic ground truth (tests/golden/make_wider_eval_golden.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytest.importorskip("scipy")
import wider_synth  # noqa: E402
from oracle import postprocess as OP  # noqa: E402
from smallhardface_b200 import wider_eval as W  # noqa: E402

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "wider_eval.npz"))


def _check(ap, pr, key):
    assert np.array_equal(np.array(ap), GOLD[key + "_ap"]), (ap, GOLD[key + "_ap"])
    assert np.array_equal(np.stack(pr), GOLD[key + "_pr"], equal_nan=True)


@pytest.mark.parametrize("seed", [0, 1])
@pytest.mark.parametrize("bug,thr", [(True, 0.5), (False, 0.5), (False, 0.3)])
def test_evaluator_equals_the_reference_bit_for_bit(tmp_path, seed, bug, thr):
    pred_dir, gt_dir = wider_synth.make(str(tmp_path), seed=seed)
    ap, pr = W.wider_eval(pred_dir, gt_dir, mimic_eval_bug=bug, IoU_thresh=thr, overlaps=OP.bbox_overlaps)
    _check(ap, pr, "s%d_bug%d_t%d" % (seed, int(bug), int(thr * 10)))


def test_closed_form_matching_equals_the_sequential_state_machine():
    """image_evaluation / image_pr_info against a literal restatement of wider_eval.py:78-117 on random cases, including
    unsorted scores and duplicate detections."""
    rng = np.random.RandomState(3)
    for _ in range(40):
        G, P = rng.randint(1, 9), rng.randint(1, 30)
        ov = rng.choice([0.0, 0.2, 0.5, 0.49999999999999994, 0.7, 1.0], (G, P))
        ignore = rng.randint(0, 2, (G, 1)).astype(np.float64)
        scores = rng.rand(P)
        for bug in (True, False):
            pred_recall = np.zeros((P, 1)); recall_list = np.zeros((G, 1)); proposal_list = np.ones((P, 1))
            for h in range(P):
                o = ov[:, h]
                if bug:
                    o = W.py2_round(o)
                mx, idx = np.max(o), np.argmax(o)
                if mx >= 0.5:
                    if ignore[idx] == 0:
                        recall_list[idx] = -1; proposal_list[h] = -1
                    elif recall_list[idx] == 0:
                        recall_list[idx] = 1
                pred_recall[h] = len(np.where(recall_list == 1)[0])
            pr, pl = W.image_evaluation(ov, ignore, 0.5, bug)
            assert np.array_equal(pr, pred_recall) and np.array_equal(pl, proposal_list)
            info = np.zeros((50, 2))
            for t in range(50):
                thresh = 1 - (t + 1.) / 50
                w = np.where(scores >= thresh)[0]
                if len(w):
                    r = w[-1]
                    info[t] = [len(np.where(proposal_list[:r + 1] == 1)[0]), pred_recall[r, 0]]
            assert np.array_equal(W.image_pr_info(50, scores, pl, pr), info)
    assert W.py2_round(np.array([0.5, 0.49999999999999994, 0.0, 1.0, 0.75])).tolist() == [1.0, 0.0, 0.0, 1.0, 1.0]


def test_missing_prediction_file_is_reported(tmp_path):
    pred_dir, gt_dir = wider_synth.make(str(tmp_path), seed=0)
    gt = W.load_gt(os.path.join(gt_dir, "wider_face_val.mat"))
    victim = os.path.join(pred_dir, gt["event_list"][3][0][0], gt["file_list"][3][0][0][0][0] + ".txt")
    os.remove(victim)
    assert W.read_pred(pred_dir, gt)[3][0] is None          # the reference logs and carries on (:35-38)
    with pytest.raises(FileNotFoundError):
        W.read_pred(pred_dir, gt, strict=True)


@pytest.mark.gpu
def test_evaluator_on_the_gpu_iou_kernel(tmp_path):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    pred_dir, gt_dir = wider_synth.make(str(tmp_path), seed=1)
    for bug, thr in ((True, 0.5), (False, 0.3)):
        ap, pr = W.wider_eval(pred_dir, gt_dir, mimic_eval_bug=bug, IoU_thresh=thr)      # default: shf_bbox_overlaps
        _check(ap, pr, "s1_bug%d_t%d" % (int(bug), int(thr * 10)))
