"""CUDA-event timing of conv1_1 (shf_conv1_tc) alone on one pyramid level: algorithmic bytes (12 B read + 256 B written
per pixel) / launch time against the measured HBM peak.  usage: time_conv1.py [H] [batch] [fmt]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv1_weights, pack_conv_first_tc_weights

H = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
fmt = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
x = torch.from_numpy((rng.rand(B, 3, H, H) * 255 - 110).astype(np.float32)).to(dev)
w = (rng.randn(64, 3, 3, 3) * 0.27).astype(np.float32)
packed, k = pack_conv1_weights(w)
pk = torch.from_numpy(packed).to(dev)
bias = torch.from_numpy((rng.randn(64) * 0.05).astype(np.float32)).to(dev)
out = H2(torch.zeros((2, B, H, H, 64), dtype=torch.float16, device=dev), fmt=fmt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
peak = 6458.1
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
packed2, k2 = pack_conv_first_tc_weights(w)
pk2 = torch.from_numpy(packed2).to(dev)
gb = B * H * H * 268 / 1e9
for name in ("single (conv1_tc.cu, one thread per pixel)", "pair (conv_first_tc.cu, two threads per pixel)"):
    ts = []
    for it in range(13):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if name.startswith("single"):
            L.call("shf_conv1_tc", _ptr(x), _ptr(pk), _ptr(bias), _ptr(out.t), B, H, H, 64, float(2.0 ** -k), 1, fmt, None, _stream())
        else:
            L.call("shf_conv_first_tc", _ptr(x), _ptr(pk2), _ptr(bias), _ptr(out.t), B, H, H, 64, 3, 1, 1, float(2.0 ** -k2), 1, fmt,
                   None, _stream())
        e1.record()
        torch.cuda.synchronize()
        if it >= 3:
            ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print("conv1_1 %s %dx%d batch %d fmt %d: %.3f ms median (min %.3f), %.2f GB algorithmic -> %.0f GB/s = %.3f of %.0f GB/s"
          % (name, H, H, B, fmt, ms, min(ts), gb, gb / ms * 1e3, gb / ms * 1e3 / peak, peak))

# context: what a pure WRITE stream reaches on this GPU (conv1_1 writes 256 B for every 12 B it reads; the measured peak in
# MEASURED_PEAKS.json is a copy, i.e. half reads)
big = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for label, fn in (("torch fill_ (1 GiB)", lambda: big.fill_(7)), ("cudaMemsetAsync (1 GiB)", lambda: big.zero_())):
    ts = []
    for it in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print("pure write, %s: %.3f ms -> %.0f GB/s = %.3f of the measured copy peak" % (label, ms, big.numel() / ms / 1e6, big.numel() / ms / 1e6 / peak))
