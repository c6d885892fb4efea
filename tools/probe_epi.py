"""Epilogue / pipeline elimination probes for the tcgen05 conv (timing only, WRONG results): needs a probe build
(`make -C smallhardface_b200/csrc clean && make -C smallhardface_b200/csrc PROBES=1`).  SHF_PROBE_EPI bits: 1 = no global
stores, 2 = drain only, 4 = no activation loads, 8 = no MMAs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv_weights, pack_conv_weights_hf8
SHAPES = [("conv1_2@2048", 64, 64, 2048, 2048), ("conv2_1@2048", 64, 128, 1024, 1024), ("conv2_2@2048", 128, 128, 1024, 1024),
          ("conv3_3@2048", 256, 256, 512, 512)]
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
for name, cin, cout, H, W in SHAPES:
    w = (rng.randn(cout, cin, 3, 3) * np.sqrt(2.0 / (cin * 9))).astype(np.float32)
    packed, kexp = pack_conv_weights(w)
    wd = [torch.from_numpy(packed).to(dev), torch.from_numpy(pack_conv_weights_hf8(w)[0]).to(dev)]
    xs = [H2.from_nchw(torch.randn((1, cin, H, W), device=dev).abs(), fmt=f) for f in (0, 1)]
    b = torch.zeros(cout, device=dev); out = H2.empty(1, H, W, cout, dev); pooled = H2.empty(1, H // 2, W // 2, cout, dev)
    for fmt in (0, 1):
        for pool in (0, 1):
            for probe in (0, 2):
                os.environ["SHF_PROBE_EPI"] = str(probe)
                if pool:
                    run = lambda: L.call("shf_conv_igemm_pool", _ptr(xs[fmt].t), _ptr(wd[fmt]), _ptr(b), None, _ptr(pooled.t), 1, H, W,
                                         cin, cout, 3, 1, cout, 0, cout, 0, float(2.0 ** -kexp), 1, fmt, fmt, None, _stream())
                else:
                    run = lambda: L.call("shf_conv_igemm", _ptr(xs[fmt].t), _ptr(wd[fmt]), _ptr(b), _ptr(out.t), 1, H, W, cin, cout, 3,
                                         1, cout, 0, float(2.0 ** -kexp), 1, fmt, fmt, None, _stream())
                run(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10): run()
                e1.record(); torch.cuda.synchronize()
                print("%-14s fmt %d pool %d probe %d  %7.3f ms" % (name, fmt, pool, probe, e0.elapsed_time(e1) / 10), flush=True)
os.environ["SHF_PROBE_EPI"] = "0"
