"""Event timing of the tcgen05 conv on the 2048-px level shapes: kernel variant (7 = one CTA per tile, 8 = CTA pairs)
x operand format (0 = h2 split fp16, 3 MMAs per 16 channels; 1 = hf8, fp16 + fp8 correction, 2 MMAs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv_weights, pack_conv_weights_hf8
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
    wd = [torch.from_numpy(packed).to(dev), torch.from_numpy(pack_conv_weights_hf8(w)[0]).to(dev)]
    xs = [x, H2.from_nchw(torch.randn((1, cin, H, W), device=dev).abs(), fmt=1)]
    b = torch.zeros(cout, device=dev); out = H2.empty(1, H, W, cout, dev)
    for impl, mode in ((7, 0), (8, 0), (7, 1), (8, 1)):
        L.call("shf_set_conv_impl", impl)
        run = lambda: L.call("shf_conv_igemm", _ptr(xs[mode].t), _ptr(wd[mode]), _ptr(b), _ptr(out.t), 1, H, W, cin, cout, k,
                             dil, cout, 0, float(2.0 ** -kexp), 1, mode, mode, None, _stream())
        run(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 2.0 * cin * cout * k * k * H * W
        print("%-16s impl %d fmt %d %7.3f ms  %7.1f TFLOP/s algorithmic" % (name, impl, mode, ms, fl / ms / 1e9), flush=True)
L.call("shf_set_conv_impl", 8)
