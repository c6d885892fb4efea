"""SURVEY 8f.1 in numbers: the Python 3 test driver (smallhardface_b200.run_test) over JPEG files on disk -- file read +
libjpeg decode (cv2.imread, on worker threads, overlapped with the GPU) + upload + the device pipeline + result files --
next to the decode alone and to the device pipeline alone.  usage: native_driver_bench.py [n_images] [batch]"""
import os
import sys
import tempfile
import time

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from smallhardface_b200 import config as C
from smallhardface_b200 import deploy
from smallhardface_b200 import run_test as R

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
work = tempfile.mkdtemp(prefix="shf_native_")
imgs = os.path.join(work, "imgs")
os.makedirs(imgs)
for i in range(n):
    cv2.imwrite(os.path.join(imgs, "im%03d.jpg" % i), deploy.synthetic_image(100 + i, (1024, 1024)), [cv2.IMWRITE_JPEG_QUALITY, 90])
tree = C.write_builtin_tree(os.path.join(work, "tree"))
proto, model = deploy.write_synthetic_deployment(os.path.join(work, "deploy"), dilation=True)
paths = sorted(os.path.join(imgs, f) for f in os.listdir(imgs))
t0 = time.perf_counter()
for p in paths:
    cv2.imread(p)
t_dec = time.perf_counter() - t0
print("cv2.imread alone, one thread: %.1f ms per 1024x1024 JPEG (%.0f images/s)" % (1e3 * t_dec / n, n / t_dec))
argv = ["--root", tree, "--conf", "configs/smallhardface.toml", "--output", os.path.join(work, "out"), "--batch", str(batch), "--amend",
        "DATA_DIR", imgs, "TEST.DB", "general_jpg", "TEST.MODEL", model, "TEST.GPU_ID", "[0]"]
cfg = R.build_cfg(tree, "configs/smallhardface.toml", argv[argv.index("--amend") + 1:])
imdb = R.get_imdb(cfg, cfg.TEST.DB)
target = os.path.join(work, "test.prototxt")
from smallhardface_b200 import prototxt
prototxt.manipulate_test(cfg, cfg.TEST.PROTOTXT, target)
from smallhardface_b200.detector import Detector
det = Detector(target, model, "cuda:0", C.detect_config(cfg))                     # weights loaded and packed once, outside the timing
R.inference(cfg, imdb, target, 0, min(2 * batch, n), batch=batch, detector=det)   # warm-up: buffers, page-locked staging
import torch
torch.cuda.synchronize()
t0 = time.perf_counter()
boxes = R.inference(cfg, imdb, target, 0, n, batch=batch, detector=det)
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("run_test.inference over %d JPEG files (decode on worker threads overlapped with the GPU, batch %d): %.1f images/s, "
      "%d detections" % (n, batch, n / t_all, sum(len(b) for b in boxes[1])))
# the device pipeline alone on the same (pre-decoded) images
dec = [cv2.imread(p) for p in paths]
det.detect(dec[:batch])
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(0, n, batch):
    det.detect(dec[i:i + batch])
torch.cuda.synchronize()
t_det = time.perf_counter() - t0
print("Detector.detect on the decoded images (no file I/O, no decode): %.1f images/s" % (n / t_det))
