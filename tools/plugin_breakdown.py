"""Where the GPU time of one `caffe.Net.forward` (plugin path, batch 1) goes, per pyramid level of a 1024x1024 image:
the upload of the page-locked level blob, the CUDA-graph replay, the result download -- each timed alone with CUDA
events (median of 10).  The plugin path runs them back to back on one stream."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

from smallhardface_b200 import deploy  # noqa: E402

proto, model = deploy.write_synthetic_deployment(bench.deploy_dir(), dilation=True)
from smallhardface_b200.detector import DetectConfig  # noqa: E402

cfg = DetectConfig()
plug = bench.PluginDriver(proto, model, cfg, 0)
img = bench.make_images(0)[0]
passes = plug.prepare(img)
for _ in range(2):
    plug.detect(passes)
eng = plug.net._engine
print("%-22s %9s %9s %9s %9s   %s" % ("level (padded)", "MB", "H2D ms", "GB/s", "graph ms", "fmt"))
tot_h = tot_g = 0.0
for key, ent in eng._graphs.items():
    shape, info, fast = key[0], key[1], key[2]
    host = torch.empty(shape, dtype=torch.float32, pin_memory=True)
    th, tg = [], []
    for it in range(10):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ent["x"].copy_(host, non_blocking=True)
        e[1].record()
        ent["graph"].replay()
        e[2].record()
        torch.cuda.synchronize()
        th.append(e[0].elapsed_time(e[1]))
        tg.append(e[1].elapsed_time(e[2]))
    mb = host.numel() * 4 / 1e6
    h, g = float(np.median(th)), float(np.median(tg))
    tot_h += h
    tot_g += g
    print("%-22s %9.2f %9.3f %9.1f %9.3f   %s" % ("%dx%d" % (shape[2], shape[3]), mb, h, mb / h, g, "hf8" if fast else "h2"))
print("per image (x2 for the mirrored passes): upload %.2f ms, graphs %.2f ms" % (2 * tot_h, 2 * tot_g))
