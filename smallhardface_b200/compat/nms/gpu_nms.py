"""`nms.gpu_nms` drop-in (reference: lib/nms/gpu_nms.pyx:11-30 over lib/nms/nms_kernel.cu): sort by descending
score on the host, call the C-ABI `_nms` symbol with host pointers, map kept rows back through the order."""
import ctypes as C

import numpy as np

from smallhardface_b200 import lib as L


def gpu_nms(dets, thresh, device_id=0):
    dets = np.asarray(dets)
    if dets.dtype != np.float32 or dets.ndim != 2:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' 2-d array")      # Cython buffer typing
    boxes_num, boxes_dim = dets.shape
    if boxes_num == 0:
        return []
    order = np.argsort(-dets[:, 4], kind="stable")            # ties: lower index first (documented deviation)
    sorted_dets = np.ascontiguousarray(dets[order, :])
    keep = np.zeros(boxes_num, dtype=np.int32)
    num_out = C.c_int(0)
    L.load()._nms(keep.ctypes.data_as(C.POINTER(C.c_int)), C.byref(num_out),
                  sorted_dets.ctypes.data_as(C.POINTER(C.c_float)), boxes_num, boxes_dim, float(thresh), int(device_id))
    if num_out.value < 0:
        raise RuntimeError("gpu_nms failed: " + L.load().shf_last_error().decode())
    return list(order[keep[:num_out.value]])
