"""``train_test.py --train false`` in Python 3 on the device-resident pipeline: the same command line,
configuration files, deploy-prototxt rewriting, dataset partition, detection cache and result files as the reference's
entry point (``train_test.py:33-137``, ``lib/test.py:290-356``), with ``Detector.detect`` in the place of the
per-pass ``caffe.Net.forward`` loop.

    python -m smallhardface_b200.run_test --root <checkout with configs/ and models/> --conf configs/smallhardface.toml \\
        --amend DATA_DIR <images> TEST.DB general_png TEST.MODEL <caffemodel> TEST.GPU_ID "[0]"

Under ``torchrun`` every rank takes the reference's contiguous shard (``lib/test.py:329-335``) and the boxes are
exchanged with the path's one all-gather (``parallel.py``, after a scalar all-reduce that sizes it); rank 0 writes the files.  The unmodified reference driver
on the drop-in ``caffe`` module is ``tools/run_reference_driver.py``; this module is the variant that keeps images and
detections on the GPU between the layers the reference round-trips through NumPy.

Datasets: the ``general_<ext>`` loader (``lib/datasets/general.py:19-79``: every ``*.<ext>`` under ``DATA_DIR``, results as
text files under the output directory) and ``wider_{train,val,test}`` (``lib/datasets/wider.py``: image list from the
annotation file, result files, the WIDER toolbox evaluation -> ``Easy / Medium / Hard`` AP, ``result.tar.gz``).  The FDDB /
AFW / PASCAL loaders are out of scope (SURVEY section 2 OUT); they run unmodified through ``compat``.
"""
from __future__ import annotations

import argparse
import datetime
import logging
import os
import os.path as osp
import pickle
import sys
from typing import List

import numpy as np

from . import config as C
from . import prototxt
from .writers import write_detections

logger = logging.getLogger(__name__)


class GeneralImdb:
    """``lib/datasets/general.py:19-42``: name ``general_<ext>``, images = every file ending in ``.<ext>`` below ``DATA_DIR``
    in ``os.walk`` order."""

    def __init__(self, data_dir: str, split: str):
        self.name = "general_" + split
        self.extension = split
        self.classes = ["bg", "face"]
        self.image_paths: List[str] = []
        for root, _dirs, files in os.walk(data_dir):
            for f in files:
                if f.endswith(".{}".format(split)):
                    self.image_paths.append(osp.join(root, f))

    num_classes = 2

    def __len__(self):
        return len(self.image_paths)

    def image_path_at(self, i):
        p = self.image_paths[i]
        assert osp.exists(p), "Path does not exist: {}".format(p)
        return p

    def evaluate_detections(self, all_boxes, output_dir="./output/"):
        """``general.py:44-79``: no ground truth -- write the text files."""
        write_detections(self.image_paths, all_boxes[1], output_dir, extension=self.extension, strip_leading_slash=True)
        return "Detection results wrote to {}".format(output_dir)


class WiderImdb:
    """``lib/datasets/wider.py:21-72,143-195``: WIDER FACE ``train`` / ``val`` / ``test``.  Layout under ``DATA_DIR``:
    ``WIDER_<split>/images/<event>/<name>.jpg``, ``wider_face_split/wider_face_<split>_bbx_gt.txt`` (``test``:
    ``wider_face_test_filelist.txt``), ``ground_truth/wider_{face,easy,medium,hard}_val.mat``.  Only what the TEST path reads
    is kept: the image list (in annotation-file order; the reference's Python 2 ``dict.keys()`` order is arbitrary and only
    permutes the result files) and the evaluation."""

    num_classes = 2

    def __init__(self, cfg, split: str, overlaps=None):
        self.name = "wider_" + split
        self.split = split
        self.classes = ["bg", "face"]
        self.cfg = cfg
        self.overlaps = overlaps                      # None: the GPU IoU kernel (wider_eval.py)
        self.imgs_path = osp.join(cfg.DATA_DIR, "WIDER_{}".format(split), "images")
        anno = osp.join(cfg.DATA_DIR, "wider_face_split",
                        "wider_face_test_filelist.txt" if split == "test" else "wider_face_{}_bbx_gt.txt".format(split))
        assert osp.isfile(anno), "Annotation file not found {}".format(anno)
        with open(anno, "r") as f:
            lines = f.readlines()
        self.image_paths: List[str] = []
        self.gt_boxes = {}
        if split == "test":
            self.image_paths = [str(p).rstrip() for p in lines]
        else:
            count = 0
            while count < len(lines):                 # wider.py:46-61
                name = str(lines[count]).rstrip()
                boxes = []
                count += 1
                n_anno = int(lines[count])
                for _ in range(n_anno):
                    count += 1
                    b = [int(round(float(x))) for x in lines[count].split(" ")[0:4]]
                    x1, y1 = max(0, b[0]), max(0, b[1])
                    boxes.append([x1, y1, x1 + b[2], y1 + b[3]])
                count += 1
                if name not in self.gt_boxes:
                    self.image_paths.append(name)
                self.gt_boxes[name] = boxes

    def __len__(self):
        return len(self.image_paths)

    def image_path_at(self, i):
        p = osp.join(self.imgs_path, self.image_paths[i])
        assert osp.exists(p), "Path does not exist: {}".format(p)
        return p

    def evaluate_detections(self, all_boxes, output_dir="./output/"):
        """``wider.py:172-195``: text files under ``<output_dir>/detections``, the WIDER toolbox evaluation, ``result.tar.gz``
        (what the evaluation server takes), the text files removed again.  ``test`` has no public ground truth: the
        archive is written and nothing is evaluated."""
        import shutil
        import tarfile
        from .wider_eval import format_result, wider_eval
        det_dir = osp.join(output_dir, "detections")
        write_detections(self.image_paths, all_boxes[1], det_dir)
        result = "Detections archived for the evaluation server (no public ground truth for the test split)"
        if self.split != "test":
            ap, _ = wider_eval(det_dir, osp.join(self.cfg.DATA_DIR, "ground_truth"), mimic_eval_bug=self.cfg.MISC.MIMIC_EVAL_BUG,
                               IoU_thresh=self.cfg.TEST.IOU_THRESH, overlaps=self.overlaps)
            self.ap = ap
            result = format_result(ap)
        with tarfile.open(osp.join(output_dir, "result.tar.gz"), "w:gz") as tar:
            tar.add(det_dir, arcname=osp.basename(det_dir))
        shutil.rmtree(det_dir)
        return result


def get_imdb(cfg, name: str):
    """``lib/datasets/factory.py:9-34`` for the datasets in scope."""
    if name.startswith("general_") and name[len("general_"):] in ("png", "jpg"):
        return GeneralImdb(cfg.DATA_DIR, name[len("general_"):])
    if name in ("wider_train", "wider_val", "wider_test"):
        return WiderImdb(cfg, name[len("wider_"):])
    if name.split("_")[0] in ("fddb", "pascalface", "afw"):
        raise NotImplementedError("dataset %s: only general_{png,jpg} and wider_{train,val,test} are in scope here" % name)
    raise KeyError("Unknown dataset: {}".format(name))


def inference(cfg, imdb, prototxt_path: str, start: int, end: int, thresh: float = 0.05, batch: int = 8, device="cuda:0",
              detector=None, group_by_size: bool = True):
    """``lib/test.py:220-267`` inference_worker for images [start, end): returns ``all_boxes[class][image]``.
    ``detector``: reuse a built ``Detector`` (weights loaded and packed) instead of constructing one."""
    import cv2
    from .detector import Detector
    det = detector or Detector(prototxt_path, cfg.TEST.MODEL, device, C.detect_config(cfg, thresh=thresh))
    all_boxes = [[[] for _ in range(end - start)] for _ in range(imdb.num_classes)]
    # with the conv stack on the GPU, `cv2.imread` (file read + JPEG / PNG decode on the host, lib/test.py:113)
    # is the next ceiling.  The decode of batch i+1 runs on worker threads (OpenCV releases the GIL) while batch i is on the
    # GPU; images reach the device as uint8 through page-locked staging (Detector.upload) and are resized there.
    from concurrent.futures import ThreadPoolExecutor

    def load(idxs):
        paths = [imdb.image_path_at(i) for i in idxs]
        return paths, list(pool.map(cv2.imread, paths))

    with ThreadPoolExecutor(max_workers=max(1, min(batch, os.cpu_count() or 1))) as pool, \
            ThreadPoolExecutor(max_workers=1) as ahead:
        # Same-sized images are batched per pyramid level (Detector.detect), so the shard is walked in size order -- image
        # headers only (PIL parses them without decoding) -- and the results go back to their own slots.  Results do not
        # depend on the order: every image is processed independently.
        order = list(range(start, end))
        if group_by_size and len(order) > batch:
            try:
                from PIL import Image

                def size_of(i):
                    with Image.open(imdb.image_path_at(i)) as im:
                        return im.size
                sizes = list(pool.map(size_of, order))
                order = [i for _, i in sorted(zip(sizes, order), key=lambda t: (t[0], t[1]))]
            except Exception as e:                                     # unreadable header: keep the list order
                logger.warning("size scan failed (%s); images are taken in list order", e)
        chunks = [order[k:k + batch] for k in range(0, len(order), batch)]
        nxt = ahead.submit(load, chunks[0]) if chunks else None
        done = 0
        for k, idxs in enumerate(chunks):
            paths, images = nxt.result()
            nxt = ahead.submit(load, chunks[k + 1]) if k + 1 < len(chunks) else None
            for p, im in zip(paths, images):
                if im is None:
                    raise IOError("cv2.imread failed on {}".format(p))
            for i, d in zip(idxs, det.detect(images)):
                all_boxes[1][i - start] = d
            done += len(idxs)
            logger.info("im_detect: %d/%d", done, end - start)
    return all_boxes


def test_net(cfg, imdb, output_dir: str, target_test: str, thresh: float = 0.05, no_cache: bool = False, batch: int = 8):
    """``lib/test.py:290-356``.  The detection cache is the reference's ``detections.pkl``."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    det_file = osp.join(output_dir, "detections.pkl")
    dets = None
    if not no_cache and osp.exists(det_file):
        try:
            with open(det_file, "rb") as f:
                dets = pickle.load(f)
            logger.info("Loading detections from cache: %s", det_file)
        except Exception:
            logger.warning("Could not load the cached detections file, detecting from scratch!")
    if dets is None:
        from .parallel import shard_range
        a, b = shard_range(len(imdb), world, rank)
        dev = "cuda:%d" % (int(os.environ.get("LOCAL_RANK", "0")) if world > 1 else _first_gpu(cfg))
        mine = inference(cfg, imdb, target_test, a, b, thresh, batch, dev)
        if world > 1:
            dets = [[[] for _ in range(len(imdb))], gather_all(mine[1], len(imdb), world, torch.device(dev))]
        else:
            dets = mine
        assert len(dets[1]) == len(imdb), "Detection result compromised"
        if not no_cache and rank == 0:
            with open(det_file, "wb") as f:
                pickle.dump(dets, f, pickle.HIGHEST_PROTOCOL)
    result = imdb.evaluate_detections(dets, output_dir) if rank == 0 else None
    if rank == 0:
        logger.info(result)
    del torch
    return dets, result


def gather_all(local_dets, n_images: int, world: int, device):
    """Every rank's per-image (M, 5) results -> the full rank-ordered list on every rank (``lib/test.py:336-344``): a scalar
    all-reduce sizes the payload, then the path's ONE all-gather of packed (count | boxes) blocks (``parallel.py``).  Rows
    travel as float32 -- what the device produced (``Detector.download`` widens to float64 afterwards)."""
    import torch
    import torch.distributed as dist
    from .parallel import gather_detections, merge_gathered
    per = int(np.ceil(1.0 * n_images / world))
    rows = torch.tensor([max([len(d) for d in local_dets] + [1])], dtype=torch.int64, device=device)
    dist.all_reduce(rows, op=dist.ReduceOp.MAX)
    rows = int(rows.item())
    buf = torch.zeros((per, rows, 5), dtype=torch.float32)
    cnt = torch.zeros((per,), dtype=torch.int32)
    for j, d in enumerate(local_dets):
        buf[j, :len(d)] = torch.from_numpy(np.asarray(d, dtype=np.float32).reshape(-1, 5))
        cnt[j] = len(d)
    merged = merge_gathered(gather_detections(buf.to(device), cnt.to(device), world, rows=rows), n_images, world)
    return [m.astype(local_dets[0].dtype) if len(local_dets) else m for m in merged]


def _first_gpu(cfg) -> int:
    g = cfg.TEST.GPU_ID
    return int(g[0] if isinstance(g, (list, tuple)) else g)


def build_cfg(root: str, conf_file: str = "", set_cfgs=None):
    """``train_test.py:52-66``."""
    cfg = C.load_default(root)
    if conf_file:
        cfg_path = conf_file if osp.isabs(conf_file) else osp.join(root, conf_file)
        C.cfg_from_file(cfg, cfg_path)
    cfg.TEST.NO_CACHE = True
    if set_cfgs:
        C.cfg_from_list(cfg, set_cfgs)
    cfg.LOG.CMD = " ".join(sys.argv)
    cfg.LOG.TIME = datetime.datetime.now().strftime("%Y_%m_%d_%H_%M_%S")
    np.random.seed(int(cfg.RNG_SEED))
    return cfg


def main(argv=None):
    ap = argparse.ArgumentParser("Test", description="Give settings")
    ap.add_argument("--root", default=".", help="checkout holding configs/ and models/ (the reference's working directory)")
    ap.add_argument("--conf", dest="conf_file", default="")
    ap.add_argument("--output", default=None, help="base of the output tree (default: <root>/output)")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--train", default="false")
    ap.add_argument("--test", default="true")
    ap.add_argument("--amend", dest="set_cfgs", default=None, nargs=argparse.REMAINDER)
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(levelname)-8s [%(filename)s:%(lineno)d] %(message)s")
    if args.train.lower() == "true":
        raise SystemExit("training is outside this framework's scope (SURVEY section 2): pass --train false")
    cfg = build_cfg(osp.abspath(args.root), args.conf_file, args.set_cfgs)
    if os.environ.get("WORLD_SIZE", "1") != "1":
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        stamp = [cfg.LOG.TIME]                          # one output directory for the job: rank 0's time stamp
        dist.broadcast_object_list(stamp, src=0)
        cfg.LOG.TIME = stamp[0]
    if cfg.TEST.DEMO.ENABLE:
        raise NotImplementedError("TEST.DEMO draws boxes with cv2 on one image; use Detector.detect directly")
    imdb = get_imdb(cfg, cfg.TEST.DB)
    out_base = args.output or "output"
    output_dir = C.get_output_dir(cfg, imdb.name, cfg.NAME + "_" + cfg.LOG.TIME, output_dir=out_base)
    target_test = osp.join(output_dir, "test.prototxt")
    prototxt.manipulate_test(cfg, cfg.TEST.PROTOTXT, target_test)
    with open(osp.join(output_dir, "cfgs.txt"), "w") as f:
        C.cfg_dump({k: cfg[k] for k in cfg if k != "TRAIN"}, f)
    dets, result = test_net(cfg, imdb, output_dir, target_test, no_cache=cfg.TEST.NO_CACHE, batch=args.batch)
    return output_dir, dets, result


if __name__ == "__main__":
    main()
