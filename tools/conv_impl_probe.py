"""Probe the conv operand-staging variants (shf_set_conv_impl 0..4) for correctness against the CPU oracle and for
speed on the 2048-level shapes.  Prints a table; never raises on a numerical mismatch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import layers as OL
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv_weights, split_h2_np

dev = torch.device("cuda:0")
CASES = [(64, 64, 24, 40, 3, 1), (64, 128, 21, 35, 3, 1), (128, 128, 30, 30, 3, 2), (128, 128, 30, 30, 3, 4),
         (512, 256, 11, 14, 1, 1), (256, 512, 16, 16, 3, 1), (64, 64, 5, 7, 3, 1)]
impls = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 1, 2, 3, 4]
for impl in impls:
    L.call("shf_set_conv_impl", impl)
    errs = []
    for cin, cout, H, W, k, dil in CASES:
        rng = np.random.RandomState(cin + cout + H + W + k + dil)
        x = (np.abs(rng.randn(1, cin, H, W)) * 40 * (rng.rand(1, cin, H, W) > 0.4)).astype(np.float32)
        hi, lo = split_h2_np(x); x = hi.astype(np.float32) + lo.astype(np.float32)
        w = (rng.randn(cout, cin, k, k) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
        b = (rng.randn(cout) * 0.05).astype(np.float32)
        packed, kexp = pack_conv_weights(w)
        w_eff = ((packed[0].astype(np.float32) + packed[1].astype(np.float32)) * np.float32(2.0 ** -kexp)).reshape(k, k, cout, cin).transpose(2, 3, 0, 1)
        xin = H2.from_nchw(torch.from_numpy(x).to(dev))
        wd, bd = torch.from_numpy(packed).to(dev), torch.from_numpy(b).to(dev)
        out = H2.empty(1, H, W, cout, dev)
        try:
            L.call("shf_conv_igemm", _ptr(xin.t), _ptr(wd), _ptr(bd), _ptr(out.t), 1, H, W, cin, cout, k, dil, cout, 0,
                   float(2.0 ** -kexp), 1, 0, 0, _stream())
            torch.cuda.synchronize()
            got = out.to_nchw().cpu().numpy()
            pad = dil if k == 3 else 0
            ref = OL.relu(OL.conv(x, w_eff, b, pad=(pad, pad), dilation=(dil, dil)))
            errs.append("%.1e" % (np.abs(got - ref).max() / np.abs(ref).max()))
        except Exception as e:
            errs.append("EXC:" + str(e)[:60])
            break
    print("impl %d correctness (rel err per case):" % impl, errs, flush=True)
    if any(e.startswith("EXC") for e in errs):
        continue
    if all(float(e) < 1e-5 for e in errs):
        for name, cin, cout, H, W, k, dil in [("conv1_2@2048", 64, 64, 2048, 2048, 3, 1), ("conv2_2@2048", 128, 128, 1024, 1024, 3, 1),
                                              ("conv3_2@2048", 256, 256, 512, 512, 3, 1), ("conv4_2@2048", 512, 512, 256, 256, 3, 1),
                                              ("head_2@2048", 128, 128, 256, 256, 3, 2), ("conv4_256@2048", 512, 256, 256, 256, 1, 1)]:
            x = H2(torch.randn((2, 1, H, W, cin), device=dev).abs().half())
            w = (np.random.RandomState(0).randn(cout, cin, k, k) * 0.02).astype(np.float32)
            packed, kexp = pack_conv_weights(w)
            wd = torch.from_numpy(packed).to(dev); bd = torch.zeros(cout, device=dev); out = H2.empty(1, H, W, cout, dev)
            run = lambda: L.call("shf_conv_igemm", _ptr(x.t), _ptr(wd), _ptr(bd), _ptr(out.t), 1, H, W, cin, cout, k, dil, cout, 0, float(2.0 ** -kexp), 1, 0, 0, _stream())
            run(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): run()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            fl = 2.0 * cin * cout * k * k * H * W
            print("   impl %d %-15s %7.3f ms  %6.1f TFLOP/s algorithmic (%6.1f executed)" % (impl, name, ms, fl / ms / 1e9, 3 * fl / ms / 1e9), flush=True)
