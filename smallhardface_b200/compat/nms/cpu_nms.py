"""`nms.cpu_nms` drop-in (reference: lib/nms/cpu_nms.pyx:17-68): same signature, same `ovr >= thresh` rule in
double precision, same return value (python list of indices into `dets`, descending score) -- evaluated by the
CUDA sweep kernel through `shf_nms_host` (mode 0); there is no host implementation in the product."""
import ctypes as C

import numpy as np

from smallhardface_b200 import lib as L


def cpu_nms(dets, thresh):
    dets = np.asarray(dets)
    if dets.dtype != np.float32 or dets.ndim != 2:
        raise ValueError("Buffer dtype mismatch, expected 'float32_t' 2-d array")
    n, dim = dets.shape
    if n == 0:
        return []
    order = np.argsort(-dets[:, 4], kind="stable")
    sorted_dets = np.ascontiguousarray(dets[order, :])
    keep = np.zeros(n, dtype=np.int32)
    num_out = C.c_int(0)
    L.call("shf_nms_host", keep.ctypes.data_as(C.POINTER(C.c_int)), C.byref(num_out),
           sorted_dets.ctypes.data_as(C.POINTER(C.c_float)), n, dim, float(thresh), 0, _device())
    return list(order[keep[:num_out.value]])


def _device():
    try:
        import torch
        return int(torch.cuda.current_device())
    except Exception:
        return 0
