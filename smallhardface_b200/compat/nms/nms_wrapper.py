"""`nms.nms_wrapper` drop-in (reference: lib/nms/nms_wrapper.py:13-21)."""
from nms.cpu_nms import cpu_nms
from nms.gpu_nms import gpu_nms


def _cfg():
    try:
        from utils.get_config import cfg
        return bool(cfg.USE_GPU_NMS), int(cfg.GPU_ID if isinstance(cfg.GPU_ID, int) else cfg.GPU_ID[0])
    except Exception:
        return True, 0


def nms(dets, thresh, force_cpu=False):
    """Dispatch to either CPU- or GPU-semantics NMS (both run on the device here)."""
    if dets.shape[0] == 0:
        return []
    use_gpu, gpu_id = _cfg()
    if use_gpu and not force_cpu:
        return gpu_nms(dets, thresh, device_id=gpu_id)
    return cpu_nms(dets, thresh)
