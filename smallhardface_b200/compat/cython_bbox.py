"""`utils.cython_bbox` drop-in (reference: lib/utils/bbox.pyx:14-142): float64 (N,4) x (K,4) -> (N,K)."""
import numpy as np

from smallhardface_b200 import lib as L


def _run(boxes, query_boxes, kind):
    import torch
    from smallhardface_b200.engine import _ptr, _stream
    for a in (boxes, query_boxes):
        if not isinstance(a, np.ndarray) or a.dtype != np.float64 or a.ndim != 2:
            raise ValueError("Buffer dtype mismatch, expected 'DTYPE_t' but got %r" % getattr(a, "dtype", type(a)))
    n, k = boxes.shape[0], query_boxes.shape[0]
    if n == 0 or k == 0:
        return np.zeros((n, k), dtype=np.float64)
    dev = torch.device("cuda", torch.cuda.current_device())
    b = torch.from_numpy(np.ascontiguousarray(boxes[:, :4])).to(dev)
    q = torch.from_numpy(np.ascontiguousarray(query_boxes[:, :4])).to(dev)
    out = torch.empty((n, k), dtype=torch.float64, device=dev)
    L.call("shf_bbox_overlaps", _ptr(b), _ptr(q), n, k, kind, _ptr(out), _stream())
    return out.cpu().numpy()


def bbox_overlaps(boxes, query_boxes):
    return _run(boxes, query_boxes, 0)


def bbox_overlaps_IoA(boxes, query_boxes):
    return _run(boxes, query_boxes, 1)


def bbox_overlaps_itself(boxes, query_boxes):
    return _run(boxes, query_boxes, 2)
