"""Box / score error of single pyramid levels of a 1024x1024 image against the CPU oracle, in raw-image px."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smallhardface_b200 import deploy, caffe_proto as cp
from smallhardface_b200.detector import DetectConfig, Detector
from oracle.net import OracleNet
from oracle import detect as OD, preprocess as PRE
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
proto, model = deploy.write_synthetic_deployment(os.path.join(tempfile.gettempdir(), "shf_b200_deploy"), dilation=True)
onet = OracleNet(proto, model, engine="torch", fast=True)
levels = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [100, 300]
kind = sys.argv[2] if len(sys.argv) > 2 else "bench"          # "bench" = the multi-octave image bench.py times, "randint" = white noise
seed = int(sys.argv[3]) if len(sys.argv) > 3 else 3
im = deploy.synthetic_image(seed) if kind == "bench" else np.random.RandomState(seed).randint(0, 256, (1024, 1024, 3)).astype(np.uint8)
print("# image: %s seed %d" % (kind, seed), flush=True)
for lv, fast_min in [(lv, fm) for lv in levels for fm in (None, 0.0)]:
    cfg = DetectConfig(scales=(lv, lv + 1), flip=False, thresh=0.002, fast_min_scale=fast_min)     # two scales -> pyramid mode; use pass 0 only
    det = Detector(proto, model, "cuda:0", cfg)
    b = det.detect_device(det.upload([im]))
    n0 = int(b["offs"][0, 1].item())
    raw = b["dets"][0, :n0].cpu().numpy()
    s = PRE.pyramid_scales(im.shape, (lv, lv + 1))[0]
    if fast_min is None:
        blob = PRE.get_image_blobs(im, [s])[0]
        p, bx = OD.forward_level(onet, blob, s)
        ref = np.hstack([bx, p[:, 1:2]])
        ref = ref[ref[:, 4] > np.float32(0.002)]
    n = min(len(raw), len(ref))
    # rows are in descending score order in both; compare after aligning by nearest score within a window
    worst_b = worst_s = 0.0
    used = np.zeros(len(raw), bool)
    cut = max(raw[:, 4].min(), ref[:, 4].min()) if min(len(raw), len(ref)) >= 10000 else -1.0      # N_DETS_PER_MODULE cut
    for i in range(len(ref)):
        if ref[i, 4] < cut + 1e-3: continue
        cand = np.where((np.abs(raw[:, 4] - ref[i, 4]) < 1e-3) & ~used)[0]
        if not len(cand): continue
        d = np.abs(raw[cand, :4] - ref[i, :4]).max(axis=1)
        j = cand[np.argmin(d)]; used[j] = True
        worst_b = max(worst_b, d.min()); worst_s = max(worst_s, abs(raw[j, 4] - ref[i, 4]))
    print("%s level %d scale %.4f rows %d/%d  worst score err %.2e  worst box err %.2e raw px (%.2e level px)" %
          ("split-f16" if fast_min is None else "f16+f8   ", lv, s, len(raw), len(ref), worst_s, worst_b, worst_b * s), flush=True)
