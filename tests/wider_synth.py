"""Synthetic WIDER-FACE-shaped ground truth (.mat files with the nested cell layout ``scipy.io.loadmat`` returns for the
official ``wider_face_val.mat`` / ``wider_{easy,medium,hard}_val.mat``) and a matching directory of detection files, for
the evaluator tests (smallhardface_b200/wider_eval.py vs the reference's lib/wider_eval_tools/wider_eval.py).
Deterministic in ``seed``; integer-valued boxes so that exact IoU == 0.5 ties (the ``mimic_eval_bug`` rounding case) occur."""
import os

import numpy as np

EVENTS = 61


def _cell(items):
    a = np.empty((len(items), 1), dtype=object)
    for i, it in enumerate(items):
        a[i, 0] = it
    return a


def make(root, seed=0, max_imgs=3, with_images=False, image_hw=(48, 64)):
    """``with_images``: also lay out ``WIDER_val/images/<event>/<name>.jpg`` (small random images) and
    ``wider_face_split/wider_face_val_bbx_gt.txt`` in the format lib/datasets/wider.py:46-61 parses, i.e. a complete
    ``DATA_DIR`` for the ``wider_val`` imdb."""
    from scipy import io as sio
    rng = np.random.RandomState(seed)
    gt_dir = os.path.join(root, "ground_truth")
    pred_dir = os.path.join(root, "pred")
    os.makedirs(gt_dir, exist_ok=True)
    anno_lines = []
    img_rng = np.random.RandomState(seed + 1000)        # images draw from their own stream: boxes / detections do not depend on with_images
    events, files, faces = [], [], []
    keep = {"easy": [], "medium": [], "hard": []}
    for e in range(EVENTS):
        ev = "%d--Event%d" % (e, e)
        events.append(np.array([ev]))
        n_img = 1 + rng.randint(max_imgs)
        f_e, b_e, k_e = [], [], {k: [] for k in keep}
        os.makedirs(os.path.join(pred_dir, ev), exist_ok=True)
        n_dets_event = 0
        for j in range(n_img):
            name = "%d_Event%d_%d" % (e, e, j)
            f_e.append(np.array([name]))
            n_face = rng.randint(0, 12) if (e + j) % 7 else 0            # some images without faces
            xy = rng.randint(0, 900, (n_face, 2))
            wh = rng.choice([8, 12, 16, 24, 32, 64, 100], (n_face, 2))
            gt = np.hstack([xy, wh]).astype(np.float64).reshape(n_face, 4)
            b_e.append(gt)
            if with_images:
                import cv2
                d = os.path.join(root, "WIDER_val", "images", ev)
                os.makedirs(d, exist_ok=True)
                cv2.imwrite(os.path.join(d, name + ".jpg"), img_rng.randint(0, 256, image_hw + (3,)).astype(np.uint8))
                anno_lines.append("%s/%s.jpg\n%d\n" % (ev, name, n_face))
                for g in gt:
                    anno_lines.append("%d %d %d %d 0 0 0 0 0 0 \n" % tuple(g))
            sizes = wh.min(axis=1) if n_face else np.zeros(0)
            for k, lim in (("easy", 60), ("medium", 20), ("hard", 0)):
                idx = np.nonzero(sizes >= lim)[0] + 1                     # 1-based, (k, 1) like the official files
                k_e[k].append(idx.reshape(-1, 1).astype(np.int32))
            # detections: jittered / half-overlapping copies of the faces + false positives + duplicates
            dets = []
            for g in gt:
                r = rng.rand()
                if r < 0.15:
                    continue
                d = g.copy()
                if r < 0.45:
                    d[0] += d[2] / 2 if rng.rand() < 0.5 else 0           # IoU exactly 1/3 or 1 ...
                    d[2] = d[2] if rng.rand() < 0.5 else d[2] * 2          # ... or exactly 0.5 (same x, double width)
                else:
                    d[:2] += rng.randint(-3, 4, 2)
                dets.append(np.r_[d, rng.rand()])
                if rng.rand() < 0.2:
                    dets.append(np.r_[d, rng.rand()])                     # duplicate detection of the same face
            for _ in range(rng.randint(0, 6)):
                dets.append(np.r_[rng.randint(0, 900, 2), rng.choice([10, 30, 80], 2), rng.rand() * 0.6])
            if not dets and ((e + j) % 3 == 0 or (j == n_img - 1 and n_dets_event == 0)):
                dets.append(np.r_[5, 5, 20, 20, 0.3])      # (_norm_score fails on an event without any detection, :48)
            n_dets_event += len(dets)
            dets = np.array(dets, dtype=np.float64).reshape(-1, 5)
            with open(os.path.join(pred_dir, ev, name + ".txt"), "w") as f:
                f.write("%s/%s.jpg\n%d\n" % (ev, name, len(dets)))
                for d in dets:
                    f.write("%d %d %d %d %g \n" % (d[0], d[1], d[2], d[3], d[4]))
        files.append(_cell(f_e))
        faces.append(_cell(b_e))
        for k in keep:
            keep[k].append(_cell(k_e[k]))
    if with_images:
        os.makedirs(os.path.join(root, "wider_face_split"), exist_ok=True)
        with open(os.path.join(root, "wider_face_split", "wider_face_val_bbx_gt.txt"), "w") as f:
            f.write("".join(anno_lines))
    base = {"event_list": _cell(events), "file_list": _cell(files), "face_bbx_list": _cell(faces)}
    sio.savemat(os.path.join(gt_dir, "wider_face_val.mat"), base)
    for k in keep:
        sio.savemat(os.path.join(gt_dir, "wider_%s_val.mat" % k), dict(base, gt_list=_cell(keep[k])))
    return pred_dir, gt_dir
