"""Runs a few representative tcgen05 conv launches (for ncu captures and quick event timing):
shapes of the 2048x2048 pyramid level of the dilated-head net (the roofline config)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv_weights

SHAPES = [  # name, cin, cout, H, W, k, dil
    ("conv1_2@2048", 64, 64, 2048, 2048, 3, 1),
    ("conv2_2@2048", 128, 128, 1024, 1024, 3, 1),
    ("conv3_2@2048", 256, 256, 512, 512, 3, 1),
    ("conv4_2@2048", 512, 512, 256, 256, 3, 1),
    ("conv5_2@2048", 512, 512, 128, 128, 3, 1),
    ("head_2@2048", 128, 128, 256, 256, 3, 2),
    ("conv4_256@2048", 512, 256, 256, 256, 1, 1),
]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
only = sys.argv[2] if len(sys.argv) > 2 else None
dev = torch.device("cuda:0")
rng = np.random.RandomState(0)
for name, cin, cout, H, W, k, dil in SHAPES:
    if only and only not in name:
        continue
    x = H2(torch.randn((2, 1, H, W, cin), device=dev).abs().half())
    w = (rng.randn(cout, cin, k, k) * np.sqrt(2.0 / (cin * k * k))).astype(np.float32)
    packed, kexp = pack_conv_weights(w)
    wd = torch.from_numpy(packed).to(dev)
    b = torch.zeros(cout, device=dev)
    out = H2.empty(1, H, W, cout, dev)
    def run():
        L.call("shf_conv_igemm", _ptr(x.t), _ptr(wd), _ptr(b), _ptr(out.t), 1, H, W, cin, cout, k, dil, cout, 0,
               float(2.0 ** -kexp), 1, 0, 0, None, _stream())
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    fl = 2.0 * cin * cout * k * k * H * W
    byts = 4.0 * (cin * H * W + cout * H * W + cout * cin * k * k)
    print("%-16s %7.3f ms  %7.1f TFLOP/s algorithmic (%6.1f executed)  %6.1f GB/s algorithmic" %
          (name, ms, fl / ms / 1e9, 3 * fl / ms / 1e9, byts / ms / 1e6))
