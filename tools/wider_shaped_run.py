"""BASELINE.json configs[4]: a WIDER-val-shaped synthetic set (mixed H x W, long side 1024) taken through the whole
pyramid detector, sharded over the ranks exactly like lib/test.py:324-344 (contiguous ranges of ceil(N / n_gpu) images),
with ONE all-gather of the boxes at the end.

    python tools/wider_shaped_run.py --images 256                      # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/wider_shaped_run.py --images 3226 [--partition area_rr]                 # 8 GPUs, NCCL
    python bench.py --gpus N --wider-shaped 3226                                      # the same leg inside the bench line

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


def partition(n_images, sizes, world, rank, kind):
    """Image indices of `rank`.  'reference' = lib/test.py:329-335 (contiguous ceil(N/G) ranges); 'area_rr' = images
    sorted by pixel area, dealt round-robin (work is ~ proportional to H x W), results un-permuted afterwards."""
    if kind == "reference":
        a, b = shard_range(n_images, world, rank)
        return list(range(a, b))
    order = sorted(range(n_images), key=lambda i: (-sizes[i][0] * sizes[i][1], i))
    return sorted(order[rank::world])


def run(n_images=3226, chunk=8, distinct=8, kind="reference", det=None):
    """Returns the result dict on rank 0 (None elsewhere).  torch.distributed must already be initialised when
    WORLD_SIZE > 1."""
    from smallhardface_b200.parallel import gather_detections
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = "cuda:%d" % local
    if det is None:
        import tempfile
        proto, model = deploy.write_synthetic_deployment(os.path.join(tempfile.gettempdir(), "shf_b200_deploy"), dilation=True)
        det = Detector(proto, model, dev, DetectConfig())
    sizes = image_sizes(n_images)
    mine_idx = partition(n_images, sizes, world, rank, kind)
    cache = {}

    def image(i):
        hw = sizes[i]
        key = (hw, i % distinct)
        if key not in cache:
            cache[key] = deploy.synthetic_image(3 + key[1], hw)
        return cache[key]

    mine = [image(i) for i in mine_idx]
    # same-shaped images of the shard are processed together (level batches need one shape); results go back in order
    order = sorted(range(len(mine)), key=lambda k: (mine[k].shape, k))
    seen = {}
    for im in mine:
        seen.setdefault(im.shape, im)
    for im in seen.values():                                  # warm-up: one call per distinct shape (buffers, kernels)
        det.detect([im] * min(chunk, 2))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    results = [None] * len(mine)
    for c0 in range(0, len(order), chunk):
        ks = order[c0:c0 + chunk]
        if any(mine[k].shape != mine[ks[0]].shape for k in ks):      # a chunk never mixes shapes
            groups = {}
            for k in ks:
                groups.setdefault(mine[k].shape, []).append(k)
            parts = list(groups.values())
        else:
            parts = [ks]
        for part in parts:
            for k, r in zip(part, det.detect([mine[k] for k in part])):
                results[k] = r
    ev1.record()
    torch.cuda.synchronize()
    busy = time.perf_counter() - t0                            # this rank's own work, before the collective
    # the ONE collective: (count | boxes) blocks, padded to the largest shard
    per = int(np.ceil(1.0 * n_images / world))
    rows = det.cfg.gather_rows
    counts = torch.zeros((per,), dtype=torch.int32, device=dev)
    boxes = torch.zeros((per, rows, 5), dtype=torch.float32, device=dev)
    host = np.zeros((per, rows, 5), np.float32)
    cnt = np.zeros((per,), np.int32)
    for i, r in enumerate(results):
        cnt[i] = len(r)
        host[i, :min(len(r), rows)] = r[:rows]
    boxes.copy_(torch.from_numpy(host))
    counts.copy_(torch.from_numpy(cnt))
    gathered = gather_detections(boxes, counts, world, rows)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt, busy], dtype=torch.float64, device=dev)
    tmax = t.clone()
    busy_all = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_gather(busy_all, t)
    else:
        busy_all = [t]
    if rank != 0:
        return None
    g = gathered.cpu()
    all_counts = g[:, :, 0, 0].contiguous().view(torch.int32).numpy()
    if int(all_counts.max()) > rows:
        raise RuntimeError("an image produced %d detections, the gather payload carries %d rows" % (all_counts.max(), rows))
    hist = {}
    for hw in sizes:
        hist["%dx%d" % hw] = hist.get("%dx%d" % hw, 0) + 1
    busy_s = [float(b[1]) for b in busy_all]
    px = [sum(sizes[i][0] * sizes[i][1] for i in partition(n_images, sizes, world, r, kind)) for r in range(world)]
    return {"metric": "images/sec (WIDER-val-shaped synthetic set, full pyramid + flip + bbox_vote)",
            "value": n_images / float(tmax[0]), "unit": "images/s", "n_gpus": world, "images": n_images,
            "seconds": float(tmax[0]), "detections_gathered": int(all_counts.sum()), "shape_histogram": hist,
            "partition": ("lib/test.py:329-335 contiguous ceil(N/G) ranges" if kind == "reference"
                          else "area-sorted round-robin (un-permuted afterwards)") + ", ONE all_gather of (count | boxes) blocks",
            "rank_busy_seconds": busy_s, "rank_megapixels": [p / 1e6 for p in px],
            "imbalance": (max(busy_s) / (sum(busy_s) / len(busy_s))) if busy_s else None,
            "gather_payload_bytes_per_rank": int(boxes.numel() * 4 + per * 20),
            "includes": "pinned H2D of every uint8 image, device pyramid, box voting, D2H of boxes, the all-gather"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=3226)
    ap.add_argument("--chunk", type=int, default=8, help="images per Detector.detect call")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic contents per shape (re-used cyclically)")
    ap.add_argument("--partition", default="reference", choices=["reference", "area_rr"])
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
    res = run(args.images, args.chunk, args.distinct, args.partition)
    if res is not None:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
