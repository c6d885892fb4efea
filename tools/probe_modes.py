"""Timing probe for the tcgen05 conv issue loop (numerics of modes 1/2 are garbage by design):
mode 0 = production (N=2BN f16 + N=BN f16 per k-step), 1 = two N=BN f16 MMAs, 2 = one N=BN f16 + one N=BN f8f6f4 (K=32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv_weights
SHAPES = [("conv1_2@2048", 64, 64, 2048, 2048, 3, 1), ("conv2_2@2048", 128, 128, 1024, 1024, 3, 1),
          ("conv3_2@2048", 256, 256, 512, 512, 3, 1), ("conv4_2@2048", 512, 512, 256, 256, 3, 1),
          ("conv5_2@2048", 512, 512, 128, 128, 3, 1), ("head_2@2048", 128, 128, 256, 256, 3, 2),
          ("conv4_256@2048", 512, 256, 256, 256, 1, 1)]
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
for name, cin, cout, H, W, k, dil in SHAPES:
    x = H2(torch.randn((2, 1, H, W, cin), device=dev).abs().half())
    w = (rng.randn(cout, cin, k, k) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    packed, kexp = pack_conv_weights(w)
    wd = torch.from_numpy(packed).to(dev); b = torch.zeros(cout, device=dev); out = H2.empty(1, H, W, cout, dev)
    run = lambda: L.call("shf_conv_igemm", _ptr(x.t), _ptr(wd), _ptr(b), _ptr(out.t), 1, H, W, cin, cout, k, dil, cout, 0,
                         float(2.0 ** -kexp), 1, _stream())
    for impl, mode in ((7, 0), (8, 0), (8, 2)):
        L.call("shf_set_conv_impl", impl)
        os.environ["SHF_PROBE_MODE"] = str(mode)
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * cin * cout * k * k * H * W
        print("%-16s impl %d mode %d %7.3f ms  %7.1f TFLOP/s algorithmic" % (name, impl, mode, ms, fl / ms / 1e9), flush=True)
os.environ["SHF_PROBE_MODE"] = "0"
L.call("shf_set_conv_impl", 8)
