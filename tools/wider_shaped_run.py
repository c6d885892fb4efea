"""BASELINE.json configs[4]: a WIDER-val-shaped synthetic set (mixed H x W, long side 1024) taken through the whole
pyramid detector, sharded over the ranks exactly like lib/test.py:324-344 (contiguous ranges of ceil(N / n_gpu) images),
with ONE all-gather of the boxes at the end.

    python tools/wider_shaped_run.py --images 256                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/wider_shaped_run.py --images 3226                                       # 8 GPUs, NCCL

The real val list is not in the reference repo; sizes are drawn with a fixed seed from the aspect ratios WIDER FACE is
dominated by (histogram printed).  Prints one JSON line (rank 0): images/s over the whole set incl. upload and download.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from smallhardface_b200 import deploy
from smallhardface_b200.detector import DetectConfig, Detector
from smallhardface_b200.parallel import shard_range

ASPECTS = [(3, 4), (2, 3), (9, 16), (1, 1), (4, 3), (3, 2)]            # h : w -- landscape 4:3 / 3:2 / 16:9, square, portrait
WEIGHTS = [0.45, 0.20, 0.10, 0.07, 0.10, 0.08]


def image_sizes(n, seed=3):
    rng = np.random.RandomState(seed)
    idx = rng.choice(len(ASPECTS), size=n, p=WEIGHTS)
    out = []
    for i in idx:
        ah, aw = ASPECTS[i]
        if ah >= aw:
            h, w = 1024, int(round(1024.0 * aw / ah))
        else:
            h, w = int(round(1024.0 * ah / aw)), 1024
        out.append((h, w))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=3226)
    ap.add_argument("--chunk", type=int, default=16, help="images per Detector.detect call")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic contents per shape (re-used cyclically)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    import tempfile
    proto, model = deploy.write_synthetic_deployment(os.path.join(tempfile.gettempdir(), "shf_b200_deploy"), dilation=True)
    det = Detector(proto, model, "cuda:%d" % local, DetectConfig())
    sizes = image_sizes(args.images)
    a, b = shard_range(args.images, world, rank)
    cache = {}

    def image(i):
        hw = sizes[i]
        key = (hw, i % args.distinct)
        if key not in cache:
            cache[key] = deploy.synthetic_image(3 + key[1], hw)
        return cache[key]

    mine = [image(i) for i in range(a, b)]
    # warm-up: one call per distinct shape (tensor maps, buffers)
    seen = {}
    for im in mine:
        seen.setdefault(im.shape, im)
    det.detect(list(seen.values()))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    results = []
    for c0 in range(0, len(mine), args.chunk):
        results.extend(det.detect(mine[c0:c0 + args.chunk]))
    # the one collective: counts, then boxes padded to the largest shard / detection count
    per = int(np.ceil(1.0 * args.images / world))
    cap = det.cfg.max_dets_out
    counts = torch.zeros((per,), dtype=torch.int32, device="cuda")
    boxes = torch.zeros((per, cap, 5), dtype=torch.float32, device="cuda")
    for i, r in enumerate(results):
        n = min(len(r), cap)
        counts[i] = n
        if n:
            boxes[i, :n] = torch.from_numpy(np.ascontiguousarray(r[:n], dtype=np.float32)).cuda()
    if world > 1:
        all_counts = [torch.empty_like(counts) for _ in range(world)]
        all_boxes = [torch.empty_like(boxes) for _ in range(world)]
        dist.all_gather(all_counts, counts)
        dist.all_gather(all_boxes, boxes)
    else:
        all_counts, all_boxes = [counts], [boxes]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        total = sum(int(c.sum().item()) for c in all_counts)
        hist = {}
        for hw in sizes:
            hist["%dx%d" % hw] = hist.get("%dx%d" % hw, 0) + 1
        print(json.dumps({"metric": "images/sec (WIDER-val-shaped synthetic set, full pyramid + flip + bbox_vote)",
                          "value": args.images / float(t[0]), "unit": "images/s", "n_gpus": world, "images": args.images,
                          "seconds": float(t[0]), "detections_gathered": total, "shape_histogram": hist,
                          "partition": "lib/test.py:329-335 contiguous ceil(N/G) ranges, one all_gather of (counts, boxes)",
                          "includes": "pinned H2D of every uint8 image, device pyramid, box voting, D2H of boxes"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
