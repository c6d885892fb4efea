"""One pyramid level of the ResNet-50-through-res3 detection net (models.build_resnet_test_net) on the B200: CUDA-event time
per launch kind (first conv, tcgen05 convs, residual adds, pooling), both operand formats.
usage: resnet_level.py [H] [batch] [nofuse]"""
import os
import sys
import tempfile
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smallhardface_b200 import caffe_proto as cp
from smallhardface_b200 import deploy
from smallhardface_b200.engine import GpuNet
from smallhardface_b200.graph import NetSpec, load_weights

H = int(sys.argv[1]) if len(sys.argv) > 1 else 1408
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
FUSE = not (len(sys.argv) > 3 and sys.argv[3] == "nofuse")      # nofuse: residual adds as Eltwise launches of their own
proto, model = deploy.write_synthetic_resnet_deployment(os.path.join(tempfile.gettempdir(), "shf_resnet"), blocks=(3, 4), input_hw=(H, H))
spec = NetSpec(cp.read_net_text(proto))
shapes = spec.infer_shapes({"data": (B, 3, H, H)})
gnet = GpuNet(spec, load_weights(spec, cp.read_net_binary(model)), "cuda:0", fast_min_scale=0.9, fuse_pool=FUSE)
flops = 0.0
for l in spec.layers:
    if l.type == "Convolution":
        n, co, ho, wo = shapes[l.tops[0]]
        ci = shapes[l.bottoms[0]][1]
        flops += 2.0 * ci * co * l.p["kh"] * l.p["kw"] * ho * wo * B
x = torch.from_numpy((np.random.RandomState(0).rand(B, 3, H, H) * 255 - 110).astype(np.float32)).cuda()
orig = gnet._run_op
for fast in (False, True):
    times = OrderedDict()

    def timed(kind, l, s, T, fmt, st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        orig(kind, l, s, T, fmt, st)
        e1.record()
        ev.append((kind if kind != "conv" else "conv %dx%d%s" % (s["k"], s["k"], "/2" if s.get("stride", 1) == 2 else ""), e0, e1))

    for it in range(4):
        ev = []
        gnet._run_op = timed
        gnet.forward_body(x, fast=fast)
        torch.cuda.synchronize()
    for k, e0, e1 in ev:
        times[k] = times.get(k, 0.0) + e0.elapsed_time(e1)
    tot = sum(times.values())
    print("ResNet-50 conv1..res3 + head, %dx%d batch %d, operand format %s: %.3f ms (%.1f GFLOP algorithmic -> %.0f TFLOP/s)"
          % (H, H, B, "hf8" if fast else "h2", tot, flops / 1e9, flops / tot / 1e9))
    for k, v in times.items():
        print("    %-12s %8.3f ms  %5.1f %%" % (k, v, 100 * v / tot))
