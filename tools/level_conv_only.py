"""BASELINE configs[1] + the roofline config of north_star: the CONV STACK of one pyramid level (no decode, no NMS) of the
dilated-head net, timed per layer with CUDA events on the launching stream.

    python tools/level_conv_only.py [size=2048] [reps=5]          # prints a per-layer table for both operand formats
    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:conv_stream --csv --log-file out.csv \
        python tools/level_conv_only.py 2048 1                    # the same launches under ncu (per-conv tensor-pipe % and DRAM bytes)

achieved = algorithmic FLOPs (2*Cin*Cout*k*k*Hout*Wout) / event time; frac = achieved / the MEASURED cuBLAS bf16 burst peak
(MEASURED_PEAKS.json: a kernel timed alone); bytes = 4 B per activation element in and out + packed weights.
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from smallhardface_b200 import caffe_proto as cp
from smallhardface_b200 import deploy
from smallhardface_b200.engine import GpuNet
from smallhardface_b200.graph import NetSpec, load_weights

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
peaks = {"bf16_tflops": 1590.0, "hbm_gbs": 6650.0}
src = "fallback (B200_PROFILING.md)"
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    src = "measured (MEASURED_PEAKS.json)"
proto, model = deploy.write_synthetic_deployment(os.path.join(tempfile.gettempdir(), "shf_b200_deploy"), dilation=True)
spec = NetSpec(cp.read_net_text(proto))
net = GpuNet(spec, load_weights(spec, cp.read_net_binary(model)), "cuda:0")
im = deploy.synthetic_image(3, (size, size)).astype(np.float32) - np.array([[[102.9801, 115.9465, 122.7717]]], np.float32)
data = torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1)[None])).cuda()
shapes = spec.infer_shapes({"data": (1, 3, size, size)})
convs = [(l, st) for k, l, st in net.ops if k == "conv"]
print("conv stack of one %dx%d level, dilated-head net, batch 1, %d reps; peak = %.1f TFLOP/s bf16 burst, %s"
      % (size, size, reps, peaks["bf16_tflops"], src))
for fast in (False, True):
    net.forward_body(data, fast=fast)                       # warm-up (function attributes, allocator)
    torch.cuda.synchronize()
    net.profile, net.events = True, []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.forward_body(data, fast=fast)
    e1.record()
    torch.cuda.synchronize()
    net.profile = False
    per = np.array([a.elapsed_time(b) for a, b in net.events]).reshape(reps, len(convs)).mean(axis=0)
    fmt = "hf8 (f16 + f8 correction, 2 MMAs per 16 channels)" if fast else "h2 (split fp16, 3 MMAs per 16 channels)"
    print("\noperand format %s" % fmt)
    print("%-28s %5s %5s %2s %10s %9s %10s %7s %9s" % ("layer", "Cin", "Cout", "k", "out", "ms", "TFLOP/s", "frac", "GB/s alg"))
    tot_f = tot_t = 0.0
    for (l, st), ms in zip(convs, per):
        _, ci, hi, wi = shapes[l.bottoms[0]]
        _, co, ho, wo = shapes[l.tops[0]]
        fl = 2.0 * st["cin"] * co * st["k"] ** 2 * ho * wo
        out_elems = co * ho * wo if ("pool_top" not in st or st.get("write_full")) else 0
        if "pool_top" in st:
            out_elems += co * (ho // 2) * (wo // 2)
        byts = 4.0 * (ci * hi * wi + out_elems) + 4.0 * st["cin"] * co * st["k"] ** 2
        tf = fl / ms / 1e9
        tot_f += fl; tot_t += ms
        print("%-28s %5d %5d %2d %10s %9.3f %10.1f %7.3f %9.1f" % (l.name + ("+pool" if "pool_top" in st else ""), st["cin"], co,
                                                                st["k"], "%dx%d" % (ho, wo), ms, tf, tf / peaks["bf16_tflops"],
                                                                byts / ms / 1e6))
    print("%-28s %39s %9.3f %10.1f %7.3f   (whole level incl. conv1_1 / deconv / gaps: %.3f ms)"
          % ("ALL tcgen05 convs", "%.1f GFLOP" % (tot_f / 1e9), tot_t, tot_f / tot_t / 1e9, tot_f / tot_t / 1e9 / peaks["bf16_tflops"],
             e0.elapsed_time(e1) / reps))
