"""`bbox_vote(det, thresh)` on the GPU with host arrays in and out (reference: ``lib/test.py:181-217``, where it is
a NumPy `while` loop with `np.delete` inside the driver itself -- O(n x clusters), the bulk of the reference's 'misc'
timer).  Not one of the reference's importable modules: a host-buffer twin of ``nms.gpu_nms`` for drivers that keep
the detections on the host (bench.py's plugin-path leg); same semantics, singleton clusters dropped, float64 output."""
import ctypes as C

import numpy as np

from smallhardface_b200 import lib as L
from .cpu_nms import _device


def bbox_vote(det, nms_thresh=0.4):
    det = np.ascontiguousarray(det, dtype=np.float32)
    if det.ndim != 2 or (det.shape[0] and det.shape[1] != 5):
        raise ValueError("det must be (n, 5) [x1, y1, x2, y2, score]")
    n = det.shape[0]
    out = np.empty((n // 2 + 1, 5), dtype=np.float32)
    num = C.c_int(0)
    L.call("shf_bbox_vote_host", out.ctypes.data_as(C.POINTER(C.c_float)), C.byref(num),
           det.ctypes.data_as(C.POINTER(C.c_float)), n, float(nms_thresh), _device())
    return out[:num.value].astype(np.float64)
