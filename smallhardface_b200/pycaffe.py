"""The reference's plugin surface: a ``caffe.Net``-compatible object over the CUDA engine.

Mirrors ``caffe/python/caffe/_caffe.cpp`` (Net ctor :109-151, Blob.reshape :244-256, zero-copy ``.data``
views :205-242) and ``caffe/python/caffe/pycaffe.py`` (``blobs``/``inputs``/``outputs``/``params`` :24-85,
``forward`` :88-134) for what ``lib/test.py`` touches, same names, argument meaning and error behaviour:

    net = Net(prototxt, caffemodel, TEST)
    net.blobs['data'].reshape(1, 3, H, W); net.blobs['im_info'].reshape(1, 3)
    out = net.forward(data=ndarray, im_info=ndarray)        # {'boxes': (R,5) view, 'cls_prob': (R,2) view}
    out['boxes'][:, [1, 3]] = w - out['boxes'][:, [3, 1]]     # in-place edits are visible through net.blobs[...].data

``Blob.data`` is a writable float32 C-contiguous ndarray that aliases the blob's host mirror; reading an
intermediate blob copies it from the device on first access after a forward (the lazy sync of
``syncedmem.cpp:39-91``).  The forward itself never leaves the GPU; only ``boxes``/``cls_prob`` (<= 10 000
rows) come back.  There is no CPU mode: ``set_mode_cpu()`` raises.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from . import caffe_proto as cp
from . import lib as L
from .graph import NetSpec, TEST as _TEST, TRAIN as _TRAIN, load_weights

TRAIN, TEST = _TRAIN, _TEST
def _env_fast_min_scale():
    """SHF_FAST_MIN_SCALE: 'none' = split-fp16 operands for every forward, a number = the im_info scale from which
    the fast f16+f8 operand format is used (default 0.9, see engine.GpuNet)."""
    import os
    v = os.environ.get("SHF_FAST_MIN_SCALE", "0.9").strip().lower()
    return None if v in ("none", "off", "") else float(v)


def _now():
    import time
    return time.perf_counter()


_state = {"device": 0, "mode": "gpu", "fast_min_scale": _env_fast_min_scale()}


def set_fast_min_scale(value):
    """Not a pycaffe function: operand-format policy of Nets constructed afterwards (None = always precise)."""
    _state["fast_min_scale"] = None if value is None else float(value)



def set_mode_gpu():
    _state["mode"] = "gpu"


def set_mode_cpu():
    raise RuntimeError("smallhardface_b200 has no CPU mode: the inference path runs on sm_100a kernels only")


def set_device(device_id):
    _state["device"] = int(device_id)
    import torch
    torch.cuda.set_device(int(device_id))


_ref_cfg = [None, False]          # [the reference's cfg object, looked up?]


def _hot_path_cfg():
    """TEST.N_DETS_PER_MODULE / SCORE_THRESH / ANCHOR_MIN_SIZE as the ProposalLayer reads them at forward time
    (``proposal_layer.py:88-92``); defaults of ``configs/default.toml`` when the reference's cfg is not importable.
    The module lookup happens once (a failing import per forward cost ~0.1 ms); the VALUES are read at every forward,
    like the reference does."""
    if not _ref_cfg[1]:
        _ref_cfg[1] = True
        try:
            import sys
            mod = sys.modules.get("utils.get_config")
            if mod is None:
                from utils import get_config as mod    # the reference's own config module, when running inside its tree
            _ref_cfg[0] = mod.cfg
        except Exception:
            _ref_cfg[0] = None
    cfg = _ref_cfg[0]
    if cfg is None:
        return dict(pre_nms_topn=10000, score_thresh=0.002, min_size=0.0)
    try:
        return dict(pre_nms_topn=int(cfg.TEST.N_DETS_PER_MODULE), score_thresh=float(cfg.TEST.SCORE_THRESH),
                    min_size=float(cfg.TEST.ANCHOR_MIN_SIZE))
    except Exception:
        return dict(pre_nms_topn=10000, score_thresh=0.002, min_size=0.0)


class Blob(object):
    """``caffe.Blob`` look-alike: shape bookkeeping + a host mirror exposed through ``.data``."""

    def __init__(self, net, name, shape, pinned=False):
        self._net, self._name = net, name
        self._pinned = bool(pinned)    # net inputs: page-locked storage, so forward() uploads straight from `.data`
        self._cap = None               # torch storage; only ever grows (Blob::Reshape, blob.cpp:46-50)
        self._tview = None
        self._alloc(tuple(int(d) for d in shape))
        self._stale = False            # device copy newer than the host mirror

    def _alloc(self, dims):
        import torch
        n = 1
        for d in dims:
            n *= d
        if self._cap is None or self._cap.numel() < n:
            pin = self._pinned and torch.cuda.is_available()
            self._cap = torch.zeros(max(n, 1), dtype=torch.float32, pin_memory=pin)
        self._tview = self._cap[:n].view(dims) if n else self._cap[:0].view(dims)
        self._host = self._tview.numpy()

    # -- shape protocol (``_caffe.cpp:453-476``) --
    @property
    def shape(self):
        return tuple(self._host.shape)

    @property
    def count(self):
        return int(self._host.size)

    def _legacy(self, i):
        s = self._host.shape
        if len(s) > 4:
            raise ValueError("Cannot use legacy accessors on Blobs with > 4 axes.")        # blob.hpp:141-154
        return int(s[i]) if i < len(s) else 1
    num = property(lambda self: self._legacy(0))
    channels = property(lambda self: self._legacy(1))
    height = property(lambda self: self._legacy(2))
    width = property(lambda self: self._legacy(3))

    def reshape(self, *dims):
        """``Blob::Reshape`` (``blob.cpp:23-51``): new shape; storage only ever grows."""
        if len(dims) > 32:
            raise ValueError("blob shape exceeds kMaxBlobAxes")
        dims = tuple(int(d) for d in dims)
        if any(d < 0 for d in dims):
            raise ValueError("blob dimensions must be non-negative")
        if dims != self._host.shape:
            self._alloc(dims)          # contents are unspecified after a reshape, as in Caffe (no zero fill)
        return None

    @property
    def data(self):
        if self._stale:
            self._net._sync_blob(self)
        return self._host

    @property
    def diff(self):
        raise RuntimeError("inference-only Net: blobs carry no diff")


class Net(object):
    """``caffe.Net(prototxt, caffemodel, phase)`` -- the legacy positional constructor ``lib/test.py:231`` uses."""

    def __init__(self, network_file, weights=None, phase=TEST, level=0, stages=None):
        if isinstance(weights, int) and phase == TEST and not isinstance(weights, bool):
            weights, phase = None, weights                    # Net(file, phase) form
        net_param = cp.read_net_text(str(network_file))          # raises "Could not open file ..." like CheckFile
        self._spec = NetSpec(net_param, int(phase))
        model = cp.read_net_binary(str(weights)) if weights is not None else cp.Msg("NetParameter")
        self._params = load_weights(self._spec, model)
        if int(phase) != TEST:
            raise RuntimeError("only the TEST phase (inference) is implemented")
        from .engine import GpuNet
        self._engine = GpuNet(self._spec, self._params, "cuda:%d" % _state["device"],
                              fast_min_scale=_state["fast_min_scale"], **_hot_path_cfg())
        shapes = self._spec.infer_shapes({})
        self.blobs = OrderedDict((n, Blob(self, n, shapes[n], pinned=n in self._spec.inputs)) for n in self._spec.blob_names)
        self.inputs = list(self._spec.inputs)
        self.outputs = list(self._spec.outputs)
        self._layer_names = [l.name for l in self._spec.layers]
        self.top_names = OrderedDict((l.name, list(l.tops)) for l in self._spec.layers)
        self.bottom_names = OrderedDict((l.name, list(l.bottoms)) for l in self._spec.layers)
        import torch
        self._prof = None
        self._out_host = None                                   # page-locked mirror of the packed result block (grow-only)
        self._guard_host = torch.zeros(tuple(self._engine.guard.shape), dtype=torch.int32,
                                       pin_memory=torch.cuda.is_available())
        self.params = OrderedDict()
        for l in self._spec.layers:
            if l.param_keys:
                self.params[l.name] = [_ParamBlob(self._params[k]) for k in l.param_keys]

    # ------------------------------------------------------------------------------------------
    def forward(self, blobs=None, start=None, end=None, **kwargs):
        """``pycaffe.py:88-134`` _Net_forward (whole-net form).  Host work per call: one copy of the caller's array into
        the page-locked input blob, one upload, one CUDA-graph launch (engine.GpuNet.forward_cached), one download of
        the packed (rows | boxes | cls_prob) block + the range guard, one synchronisation."""
        if start is not None or end is not None:
            return self._forward_range(blobs, start, end, kwargs)
        import torch
        prof = self._prof                                       # bench.py: where a forward's host time goes
        t0 = _now() if prof is not None else 0.0
        if kwargs:
            if set(kwargs.keys()) != set(self.inputs):
                raise Exception("Input blob arguments do not match net inputs.")
            for in_, blob in kwargs.items():
                if blob.shape[0] != self.blobs[in_].shape[0]:
                    raise Exception("Input is not batch sized")
                self._assign(self.blobs[in_], blob)              # blob.data[...] = arr, as pycaffe
        eng = self._engine
        eng.cfg.update(_hot_path_cfg())
        if eng.tail is None or eng.has_python or len(self.inputs) != 2:
            return self._forward_generic(blobs)
        data_blob = self.blobs[self.inputs[0]]
        info = self.blobs[self.inputs[1]]._host.reshape(-1)
        n, c, h, w = data_blob.shape
        if (h % 16) or (w % 16):
            # concat_layer.cpp:40-44 would fail the same way inside Caffe's Reshape
            self._spec.infer_shapes({self.inputs[0]: data_blob.shape})
        info = (float(info[0]), float(info[1]), float(info[2]))
        stream = torch.cuda.current_stream()
        t1 = _now() if prof is not None else 0.0
        for attempt in (0, 1):
            eng.guard.zero_()
            # `.data` of an input blob is page-locked: the upload reads it directly (valid until the sync below)
            pack = eng.forward_cached(data_blob._tview, info)
            self._guard_host.copy_(eng.guard, non_blocking=True)
            if pack is not None:
                if self._out_host is None or self._out_host.numel() < pack.numel():
                    self._out_host = torch.empty(pack.numel(), dtype=torch.float32, pin_memory=True)
                self._out_host[:pack.numel()].copy_(pack, non_blocking=True)
            t2 = _now() if prof is not None else 0.0
            stream.synchronize()                                 # the one synchronisation of the forward
            # a fast-format level outside the format's exponent window: the engine has switched itself to split fp16
            # (sticky); repeat this forward there.  fp16 overflow raises.
            if not eng.check_ranges(self._guard_host):
                break
        for b in self.blobs.values():
            b._stale = True
        for name in self.inputs:
            self.blobs[name]._stale = False
        if pack is not None:
            host = self._out_host.numpy()
            R = int(host[:1].view(np.int32)[0])
            topn = (pack.numel() - 4) // 7
            tops = eng.tail["tops"]
            # views into the page-locked result block, overwritten by the next forward -- what Caffe's top blobs are
            # (lib/test.py:154 copies what it keeps)
            self._set_host(tops[0], host[4:4 + 5 * topn].reshape(topn, 5)[:R])
            if len(tops) > 1:
                self._set_host(tops[1], host[4 + 5 * topn:4 + 7 * topn].reshape(topn, 2)[:R])
        outs = set(self.outputs + list(blobs or []))
        res = {out: self.blobs[out].data for out in outs}
        if prof is not None:
            t3 = _now()
            prof["assign_s"] = prof.get("assign_s", 0.0) + (t1 - t0)      # caller's array -> page-locked input blob
            prof["enqueue_s"] = prof.get("enqueue_s", 0.0) + (t2 - t1)    # upload + graph launch + download, enqueued
            prof["wait_s"] = prof.get("wait_s", 0.0) + (t3 - t2)          # GPU (H2D + kernels + D2H) not hidden by the host
            prof["forwards"] = prof.get("forwards", 0) + 1
        return res

    def _forward_generic(self, blobs=None):
        """Nets without the detection tail, with generic Python layers, or with other input sets: every launch eagerly
        (no CUDA graph: Python layers run on the host in the middle of the net), outputs synchronised lazily by `.data`."""
        import torch
        eng = self._engine
        dev = eng.device
        ins = {name: self.blobs[name]._tview.to(dev, non_blocking=True) for name in self.inputs}
        shapes = self._spec.infer_shapes({name: self.blobs[name].shape for name in self.inputs})
        first = ins.pop(self.inputs[0]) if self.inputs else None
        eng.guard.zero_()
        eng.forward_body(first, fast=False, extra_inputs=ins)
        res = None
        if eng.tail is not None:
            info = self.blobs[eng.tail["info"]]._host.reshape(-1)
            res = eng.run_tail(0, (float(info[0]), float(info[1]), float(info[2])))
        torch.cuda.current_stream().synchronize()
        eng.check_ranges()
        for name, b in self.blobs.items():
            b._stale = name not in self.inputs
        if res is not None:
            boxes, probs, rows = res
            R = int(rows.item())
            self._set_host(eng.tail["tops"][0], boxes[:R].cpu().numpy())
            if len(eng.tail["tops"]) > 1:
                self._set_host(eng.tail["tops"][1], probs[:R].cpu().numpy())
        del shapes
        outs = set(self.outputs + list(blobs or []))
        return {out: self.blobs[out].data for out in outs}

    def _forward_range(self, blobs, start, end, kwargs):
        """``net.forward(start=, end=)`` (pycaffe.py:88-134, test: caffe/python/caffe/test/test_net.py:74-90): run only the
        layers start..end, reading their bottoms from the blobs' CURRENT contents (whatever the caller wrote through
        ``.data``) and returning the tops of ``end``.  The range must fall on launch boundaries of the fused plan."""
        import torch
        eng = self._engine
        if kwargs:
            if set(kwargs.keys()) != set(self.inputs):
                raise Exception("Input blob arguments do not match net inputs.")
            for in_, blob in kwargs.items():
                self._assign(self.blobs[in_], blob)
        if start is not None and start not in self._layer_names:
            raise ValueError("%r is not in list" % start)             # list.index() in pycaffe
        if end is not None and end not in self._layer_names:
            raise ValueError("%r is not in list" % end)
        first_op, last_op = eng.op_span(start, end)
        produced, needed = set(), []
        for kind, l, st in eng.ops[first_op:last_op + 1]:
            for b in l.bottoms:
                if b not in produced and b not in needed:
                    needed.append(b)
            produced.update(l.tops)
            if "top" in st:
                produced.add(st["top"])
            if "pool_top" in st:
                produced.add(st["pool_top"])
        preset = {b: torch.from_numpy(np.ascontiguousarray(self.blobs[b].data)).to(eng.device) for b in needed}
        eng.guard.zero_()
        eng.forward_body(None, fast=False, op_range=(first_op, last_op), preset=preset)
        torch.cuda.current_stream().synchronize()
        eng.check_ranges()
        for name in produced:
            if name in self.blobs and name in eng.tensors:
                self.blobs[name]._stale = True
        tops = self.top_names[end] if end is not None else self.outputs
        outs = set(list(tops) + list(blobs or []))
        return {out: self.blobs[out].data for out in outs}

    @staticmethod
    def _assign(blob, arr):
        """``blob.data[...] = arr`` -- through torch's multi-threaded copy when the array is a plain float32 block of the
        blob's shape (a 24 MB level blob per forward), NumPy's broadcasting assignment otherwise."""
        import torch
        host = blob.data
        if (isinstance(arr, np.ndarray) and arr.dtype == np.float32 and arr.shape == host.shape and arr.flags.c_contiguous
                and arr.size >= (1 << 16) and blob._tview is not None):
            blob._tview.copy_(torch.from_numpy(arr))
        else:
            host[...] = arr

    __call__ = forward

    def _set_host(self, name, arr):
        b = self.blobs[name]
        b._host = arr if (arr.dtype == np.float32 and arr.flags.c_contiguous) else np.ascontiguousarray(arr, dtype=np.float32)
        b._stale = False

    def _sync_blob(self, blob):
        t = self._engine.blob_nchw(blob._name)                   # raises for blobs fused away inside a kernel
        blob._host = np.ascontiguousarray(t.cpu().numpy().reshape(blob._net._current_shape(blob._name, t)))
        blob._stale = False

    def _current_shape(self, name, t):
        return tuple(t.shape)

    @property
    def layer_dict(self):
        return OrderedDict((l.name, l) for l in self._spec.layers)

    def save(self, filename):
        from .deploy import params_to_netparameter
        cp.write_net_binary(str(filename), params_to_netparameter(self._spec, self._params))


class _ParamBlob(object):
    def __init__(self, arr):
        self.data = arr

    @property
    def shape(self):
        return tuple(self.data.shape)


class Layer(object):
    """``caffe.Layer``: base class of Python layers (``caffe/include/caffe/layers/python_layer.hpp:19-43``, exposed by
    ``_caffe.cpp:420-430``).  The net builds ``module.layer()``, sets ``param_str`` and ``phase``, calls
    ``setup(bottom, top)`` once, then ``reshape(bottom, top)`` + ``forward(bottom, top)`` per forward with Blob-like
    bottoms / tops (``smallhardface_b200.graph.PyBlob``: ``.data``, ``.diff``, ``.reshape``, ``.shape``, ``.num`` ...).
    The one Python layer on the test path, ``lib.layers.proposal_layer.ProposalLayer``, is matched by name and runs as
    the fused CUDA tail instead; every other subclass goes through this protocol on host blobs, as in Caffe."""
    param_str = ""
    phase = TEST
    blobs = None

    def setup(self, bottom, top):
        pass

    def reshape(self, bottom, top):
        pass

    def forward(self, bottom, top):
        pass

    def backward(self, top, propagate_down, bottom):
        pass


def layer_type_list():
    """``caffe.layer_type_list()`` (``_caffe.cpp:380``): the layer types this build executes."""
    from .graph import SUPPORTED
    return list(SUPPORTED)


def _training_only(name):
    def _raise(*a, **k):
        raise RuntimeError("caffe.%s belongs to the training path, which is out of scope for this build" % name)
    return _raise


# referenced by lib/train.py at import/definition time only
SGDSolver = _training_only("SGDSolver")
NCCL = _training_only("NCCL")
set_random_seed = lambda seed: None
set_solver_count = _training_only("set_solver_count")
set_solver_rank = _training_only("set_solver_rank")
set_multiprocess = _training_only("set_multiprocess")
init_log = lambda *a, **k: None
log = lambda *a, **k: None
