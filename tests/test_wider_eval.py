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


def _pred_as_all_boxes(pred_dir, imdb):
    """Detections of the synthetic prediction files as the driver's all_boxes[1] ([x1, y1, x2, y2, score] rows)."""
    out = []
    for rel in imdb.image_paths:
        lines = open(os.path.join(pred_dir, rel[:-4] + ".txt")).read().splitlines()
        d = np.array([[float(v) for v in ln.split()] for ln in lines[2:2 + int(lines[1])]], dtype=np.float64).reshape(-1, 5)
        d[:, 2] += d[:, 0]
        d[:, 3] += d[:, 1]
        out.append(d)
    return out


def test_wider_imdb_lists_images_and_evaluates_like_the_toolbox(tmp_path):
    """run_test.WiderImdb (lib/datasets/wider.py): annotation parsing, image paths, evaluate_detections = writer + evaluator
    + result.tar.gz; the AP equals the evaluator run directly on the same detection files."""
    import tarfile
    from smallhardface_b200 import config as C
    from smallhardface_b200 import run_test as R
    root = str(tmp_path / "WIDER")
    pred_dir, gt_dir = wider_synth.make(root, seed=1, with_images=True)
    cfg = C.load_default(C.write_builtin_tree(str(tmp_path / "tree")))
    C.cfg_from_list(cfg, ["DATA_DIR", root, "TEST.DB", "wider_val"])
    imdb = R.get_imdb(cfg, cfg.TEST.DB)
    imdb.overlaps = OP.bbox_overlaps                     # CPU test: the oracle's IoU instead of the GPU kernel
    gt = W.load_gt(os.path.join(gt_dir, "wider_face_val.mat"))
    n_img = sum(gt["file_list"][i][0].shape[0] for i in range(W.EVENT_NUM))
    assert imdb.name == "wider_val" and len(imdb) == n_img and os.path.isfile(imdb.image_path_at(0))
    assert imdb.image_paths[0] == "0--Event0/0_Event0_0.jpg"
    k = next(i for i, p in enumerate(imdb.image_paths) if imdb.gt_boxes[p])
    ev, nm = imdb.image_paths[k].split("/")
    e = int(ev.split("--")[0])
    j = [str(x[0][0]) for x in gt["file_list"][e][0]].index(nm[:-4])
    g = gt["face_bbx_list"][e][0][j][0][0]
    assert imdb.gt_boxes[imdb.image_paths[k]][0] == [int(g[0]), int(g[1]), int(g[0] + g[2]), int(g[1] + g[3])]
    out = tmp_path / "out"
    out.mkdir()
    msg = imdb.evaluate_detections([[], _pred_as_all_boxes(pred_dir, imdb)], str(out))
    ap, _ = W.wider_eval(pred_dir, gt_dir, mimic_eval_bug=True, IoU_thresh=0.5, overlaps=OP.bbox_overlaps)
    assert msg == W.format_result(ap) and msg.startswith("Easy: 0.0742, Medium: 0.2602, Hard: 0.4050")
    assert np.array_equal(np.array(imdb.ap), GOLD["s1_bug1_t5_ap"])            # = the reference evaluator's numbers
    assert not os.path.exists(out / "detections")
    with tarfile.open(out / "result.tar.gz") as tar:
        names = tar.getnames()
    assert "detections/0--Event0/0_Event0_0.txt" in names and sum(n.endswith(".txt") for n in names) == n_img


@pytest.mark.gpu
def test_native_driver_on_a_wider_shaped_tree(tmp_path):
    """`python -m smallhardface_b200.run_test --amend TEST.DB wider_val ...` end to end on the B200: WIDER directory layout,
    Detector over every image, result files, the evaluation on the GPU IoU kernel, result.tar.gz."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from smallhardface_b200 import config as C
    from smallhardface_b200 import deploy
    from smallhardface_b200 import run_test as R
    root = str(tmp_path / "WIDER")
    wider_synth.make(root, seed=0, with_images=True, max_imgs=2)
    tree = C.write_builtin_tree(str(tmp_path / "tree"))
    proto, model = deploy.write_synthetic_deployment(str(tmp_path / "deploy"), dilation=True)
    out_dir, dets, result = R.main(["--root", tree, "--conf", "configs/smallhardface.toml", "--output", str(tmp_path / "out"), "--amend",
                                    "DATA_DIR", root, "TEST.DB", "wider_val", "TEST.MODEL", model, "TEST.GPU_ID", "[0]",
                                    "TEST.SCALES", "[100, 300]"])
    assert result.startswith("Easy: ") and os.path.isfile(os.path.join(out_dir, "result.tar.gz"))
    assert len(dets[1]) > 61 and sum(len(d) for d in dets[1]) > 0
    print("native driver on a WIDER-shaped tree (%d images, synthetic weights): %s" % (len(dets[1]), result))
