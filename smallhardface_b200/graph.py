"""Deploy-net graph: NetParameter -> ordered layer specs, blob table, shape inference, weights.

Restates the parts of ``caffe/src/caffe/net.cpp`` the inference path depends on:
legacy ``input:``/``input_shape`` upgrade (``util/upgrade_proto.cpp:966-998``), phase filtering
(``net.cpp:259-287``), in-place tops (``net.cpp:363-370``), named-parameter sharing
(``net.cpp:457-512``), "unconsumed tops are net outputs" (``net.cpp:240-246``) and
load-by-layer-name with exact shape checks (``net.cpp:733-785``).  Split layers are not
materialised: sharing a blob between consumers is free here, so ``net.blobs`` does not contain
the ``*_split_*`` aliases Caffe inserts (nothing on the test path reads them).
"""
from __future__ import annotations

import ast
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import caffe_proto as cp

TRAIN, TEST = 0, 1

SUPPORTED = ("Input", "Convolution", "Deconvolution", "ReLU", "Pooling", "Concat", "Reshape",
             "Softmax", "Python", "Split", "BatchNorm", "Scale", "Eltwise")


@dataclass
class LayerSpec:
    name: str
    type: str
    bottoms: List[str]
    tops: List[str]
    p: dict = field(default_factory=dict)          # normalised layer parameters
    param_keys: List[str] = field(default_factory=list)   # storage keys of weight/bias blobs
    msg: Optional[cp.Msg] = None


PROPOSAL_LAYER = ("lib.layers.proposal_layer", "ProposalLayer")      # executed by the fused CUDA tail, never instantiated


class PyBlob:
    """What a generic ``caffe.Layer`` sees as ``bottom[i]`` / ``top[i]``: the ``caffe.Blob`` attributes exposed by
    ``caffe/python/caffe/_caffe.cpp:453-488`` over a host float32 array."""

    def __init__(self, shape=()):
        self.data = np.zeros(tuple(int(d) for d in shape), dtype=np.float32)
        self._diff = None

    def reshape(self, *dims):
        dims = tuple(int(d) for d in dims)
        if dims != self.data.shape:
            self.data = np.zeros(dims, dtype=np.float32)
            self._diff = None

    @property
    def diff(self):
        if self._diff is None or self._diff.shape != self.data.shape:
            self._diff = np.zeros_like(self.data)
        return self._diff

    shape = property(lambda self: tuple(self.data.shape))
    count = property(lambda self: int(self.data.size))

    def _legacy(self, i):
        if self.data.ndim > 4:
            raise ValueError("Cannot use legacy accessors on Blobs with > 4 axes.")
        return int(self.data.shape[i]) if i < self.data.ndim else 1
    num = property(lambda self: self._legacy(0))
    channels = property(lambda self: self._legacy(1))
    height = property(lambda self: self._legacy(2))
    width = property(lambda self: self._legacy(3))


class _LayerBlobs(list):
    """``layer.blobs`` of a Python layer (``python_layer.hpp`` exposes the Layer's blob vector; ``add_blob(*dims)``)."""

    def add_blob(self, *dims):
        self.append(PyBlob(dims))


def make_python_layer(spec: "LayerSpec", phase: int, bottom_shapes):
    """``PythonLayer::LayerSetUp`` (``caffe/include/caffe/layers/python_layer.hpp:19-32``): import ``module``, build
    ``layer``, hand it ``param_str`` and ``phase``, call ``setup(bottom, top)`` once with shaped bottoms."""
    import importlib
    mod = importlib.import_module(spec.p["module"])
    obj = getattr(mod, spec.p["layer"])()
    obj.param_str = spec.p["param_str"]
    obj.phase = phase
    if not hasattr(obj, "blobs") or obj.blobs is None:
        obj.blobs = _LayerBlobs()
    bottoms = [PyBlob(s) for s in bottom_shapes]
    tops = [PyBlob() for _ in spec.tops]
    obj.setup(bottoms, tops)
    return dict(obj=obj, bottoms=bottoms, tops=tops)


def _pair(rep, h, w, default, what, lname):
    """ConvolutionParameter's repeated/ _h/_w forms -> (h, w) (``base_conv_layer.cpp:33-92``)."""
    if h is not None or w is not None:
        if len(rep):
            raise ValueError("%s: either %s or %s_h/_w, not both" % (lname, what, what))
        return int(h or default), int(w or default)
    if len(rep) == 0:
        return default, default
    if len(rep) == 1:
        return int(rep[0]), int(rep[0])
    if len(rep) == 2:
        return int(rep[0]), int(rep[1])
    raise ValueError("%s: only 2 spatial axes supported" % lname)


def _conv_params(l: cp.Msg) -> dict:
    c = l.convolution_param
    if not c.has("num_output"):
        raise ValueError("%s: num_output missing" % l.name)
    kh, kw = _pair(c.kernel_size, c._f.get("kernel_h"), c._f.get("kernel_w"), None, "kernel_size", l.name)
    if kh is None:
        raise ValueError("%s: kernel_size missing" % l.name)
    ph, pw = _pair(c.pad, c.pad_h if c.has("pad_h") and c.pad_h else None,
                   c.pad_w if c.has("pad_w") and c.pad_w else None, 0, "pad", l.name)
    sh, sw = _pair(c.stride, c._f.get("stride_h"), c._f.get("stride_w"), 1, "stride", l.name)
    dh, dw = _pair(c.dilation, None, None, 1, "dilation", l.name)
    return dict(num_output=int(c.num_output), bias_term=bool(c.bias_term), group=int(c.group),
                kh=kh, kw=kw, ph=ph, pw=pw, sh=sh, sw=sw, dh=dh, dw=dw)


def _pool_params(l: cp.Msg) -> dict:
    q = l.pooling_param
    if q.global_pooling:
        raise ValueError("%s: global pooling not on the hot path" % l.name)
    kh = int(q.kernel_h) if q.has("kernel_h") else int(q.kernel_size)
    kw = int(q.kernel_w) if q.has("kernel_w") else int(q.kernel_size)
    sh = int(q.stride_h) if q.has("stride_h") else int(q.stride)
    sw = int(q.stride_w) if q.has("stride_w") else int(q.stride)
    ph = int(q.pad_h) if q.has("pad_h") and q.pad_h else int(q.pad)
    pw = int(q.pad_w) if q.has("pad_w") and q.pad_w else int(q.pad)
    return dict(pool=int(q.pool), kh=kh, kw=kw, sh=sh, sw=sw, ph=ph, pw=pw)


def _state_meets_rule(phase: int, rule: cp.Msg) -> bool:
    """``net.cpp:289-354`` restricted to the phase test (no levels/stages in the deploy nets)."""
    if rule.has("phase") and rule.phase != phase:
        return False
    return True


def upgrade_net_input(net: cp.Msg) -> cp.Msg:
    """Legacy ``input`` / ``input_shape`` / ``input_dim`` -> one ``Input`` layer at index 0
    (``upgrade_proto.cpp:966-998``)."""
    if not net.has("input"):
        return net
    out = net.copy()
    lay = cp.Msg("LayerParameter", name="input", type="Input")
    has_shape = len(out.input_shape) > 0
    has_dim = len(out.input_dim) > 0
    for i, nm in enumerate(out.input):
        lay.top.append(nm)
        if has_shape:
            lay.input_param.shape.append(out.input_shape[i].copy())
        elif has_dim:
            lay.input_param.shape.append(cp.Msg("BlobShape", dim=[int(d) for d in out.input_dim[4 * i:4 * i + 4]]))
    out.clear("input"); out.clear("input_shape"); out.clear("input_dim")
    out.layer = [lay] + list(out.layer)
    return out


class NetSpec:
    """Static description of a deploy net."""

    def __init__(self, net_param: cp.Msg, phase: int = TEST):
        self.phase = phase
        net_param = upgrade_net_input(net_param)
        self.name = net_param.name
        self.layers: List[LayerSpec] = []
        self.blob_names: List[str] = []
        self.inputs: List[str] = []
        self.input_shapes: Dict[str, Tuple[int, ...]] = {}
        self.py_layers: Dict[str, dict] = {}           # generic Python layers: name -> {obj, bottoms, tops} (lazy, see infer_shapes)
        self.param_owner: Dict[str, str] = {}          # named param -> storage key
        self.param_shapes: Dict[str, Tuple[int, ...]] = {}
        produced: "OrderedDict[str, bool]" = OrderedDict()    # blob -> consumed?
        for l in net_param.layer:
            if l.has("phase") and l.phase != phase and not (l.include or l.exclude):
                continue
            if l.include and not any(_state_meets_rule(phase, r) for r in l.include):
                continue
            if l.exclude and any(_state_meets_rule(phase, r) for r in l.exclude):
                continue
            if l.type not in SUPPORTED:
                raise ValueError("layer %r: type %r is outside the inference hot path "
                                 "(supported: %s)" % (l.name, l.type, ", ".join(SUPPORTED)))
            spec = LayerSpec(l.name, l.type, list(l.bottom), list(l.top), msg=l)
            for b in spec.bottoms:
                if b not in produced:
                    raise ValueError("Unknown bottom blob '%s' (layer '%s')" % (b, l.name))
                produced[b] = True
            for t in spec.tops:
                if t in produced and t not in spec.bottoms:
                    raise ValueError("Top blob '%s' produced by multiple sources." % t)
                if t not in produced:
                    self.blob_names.append(t)
                produced[t] = False
            if l.type == "Input":
                for i, t in enumerate(spec.tops):
                    self.inputs.append(t)
                    shp = l.input_param.shape
                    s = shp[i] if len(shp) > 1 else (shp[0] if len(shp) == 1 else None)
                    self.input_shapes[t] = tuple(int(d) for d in s.dim) if s is not None else ()
            elif l.type in ("Convolution", "Deconvolution"):
                spec.p = _conv_params(l)
            elif l.type == "Pooling":
                spec.p = _pool_params(l)
            elif l.type == "ReLU":
                spec.p = dict(negative_slope=float(l.relu_param.negative_slope))
            elif l.type == "Concat":
                spec.p = dict(axis=int(l.concat_param.axis) if l.concat_param.has("axis") or not l.concat_param.has("concat_dim")
                              else int(l.concat_param.concat_dim))
            elif l.type == "Reshape":
                r = l.reshape_param
                spec.p = dict(dims=[int(d) for d in r.shape.dim], axis=int(r.axis), num_axes=int(r.num_axes))
            elif l.type == "Softmax":
                spec.p = dict(axis=int(l.softmax_param.axis))
            elif l.type == "BatchNorm":                    # batch_norm_layer.cpp:12-21
                q = l.batch_norm_param
                spec.p = dict(use_global_stats=bool(q.use_global_stats) if q.has("use_global_stats") else phase == TEST,
                              eps=float(q.eps))
            elif l.type == "Scale":                        # scale_layer.cpp:12-60 (per-channel form only)
                q = l.scale_param
                if len(spec.bottoms) != 1 or int(q.axis) != 1 or int(q.num_axes) != 1:
                    raise ValueError("%s: only the per-channel Scale (one bottom, axis 1, num_axes 1) is supported" % l.name)
                spec.p = dict(bias_term=bool(q.bias_term))
            elif l.type == "Eltwise":                      # eltwise_layer.cpp:10-35
                q = l.eltwise_param
                coeff = [float(c) for c in q.coeff]
                if coeff and len(coeff) != len(spec.bottoms):
                    raise ValueError("Eltwise Layer takes one coefficient per bottom blob.")
                if coeff and int(q.operation) == 0:
                    raise ValueError("Eltwise layer only takes coefficients for summation.")
                spec.p = dict(operation=int(q.operation), coeff=coeff or [1.0] * len(spec.bottoms))
            elif l.type == "Python":
                q = l.python_param
                spec.p = dict(module=q.module, layer=q.layer, param_str=q.param_str)
            self.layers.append(spec)
        self.outputs = [b for b, used in produced.items() if not used]
        self._assign_params()

    # -- parameters ------------------------------------------------------------------
    def _assign_params(self) -> None:
        for spec in self.layers:
            if spec.type not in ("Convolution", "Deconvolution", "BatchNorm", "Scale"):
                continue
            n = 3 if spec.type == "BatchNorm" else (2 if spec.p["bias_term"] else 1)
            pspecs = spec.msg.param
            for i in range(n):
                pname = pspecs[i].name if i < len(pspecs) and pspecs[i].has("name") else ""
                key = "%s/%d" % (spec.name, i)
                if pname:
                    if pname in self.param_owner:
                        key = self.param_owner[pname]              # share the owner's storage
                    else:
                        self.param_owner[pname] = key
                spec.param_keys.append(key)

    def check_param_shape(self, spec: LayerSpec, cin: int) -> List[Tuple[int, ...]]:
        p = spec.p
        if spec.type in ("BatchNorm", "Scale"):
            # batch_norm_layer.cpp:25-36: mean (C), variance (C), moving-average factor (1); scale_layer.cpp: gamma (C)[, beta (C)]
            shapes = [(cin,), (cin,), (1,)] if spec.type == "BatchNorm" else [(cin,)] * (2 if p["bias_term"] else 1)
            for key, shp in zip(spec.param_keys, shapes):
                prev = self.param_shapes.get(key)
                if prev is not None and prev != shp:
                    raise ValueError("Cannot share param %s: shape mismatch %s vs %s" % (key, prev, shp))
                self.param_shapes[key] = shp
            return shapes
        g = p["group"]
        if spec.type == "Convolution":
            if cin % g or p["num_output"] % g:
                raise ValueError("%s: channels not divisible by group" % spec.name)
            w = (p["num_output"], cin // g, p["kh"], p["kw"])
        else:                                                      # base_conv_layer.cpp:124-139 (reversed)
            w = (cin, p["num_output"] // g, p["kh"], p["kw"])
        shapes = [w] + ([(p["num_output"],)] if p["bias_term"] else [])
        for key, shp in zip(spec.param_keys, shapes):
            prev = self.param_shapes.get(key)
            if prev is not None and prev != shp:                   # net.cpp:489-508 STRICT share mode
                raise ValueError("Cannot share param %s: shape mismatch %s vs %s" % (key, prev, shp))
            self.param_shapes[key] = shp
        return shapes

    # -- shapes ----------------------------------------------------------------------
    def infer_shapes(self, input_shapes: Dict[str, Tuple[int, ...]]) -> "OrderedDict[str, Tuple[int, ...]]":
        """Blob shapes for given input shapes (what ``Net::Reshape`` computes layer by layer)."""
        shapes: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
        for spec in self.layers:
            t = spec.type
            if t == "Input":
                for nm in spec.tops:
                    shapes[nm] = tuple(input_shapes.get(nm, self.input_shapes[nm]))
                continue
            bs = [shapes[b] for b in spec.bottoms]
            if t == "Convolution":
                n, c, h, w = bs[0]
                p = spec.p
                self.check_param_shape(spec, c)
                ho = (h + 2 * p["ph"] - (p["dh"] * (p["kh"] - 1) + 1)) // p["sh"] + 1     # conv_layer.cpp:8-28
                wo = (w + 2 * p["pw"] - (p["dw"] * (p["kw"] - 1) + 1)) // p["sw"] + 1
                out = (n, p["num_output"], ho, wo)
            elif t == "Deconvolution":
                n, c, h, w = bs[0]
                p = spec.p
                self.check_param_shape(spec, c)
                ho = p["sh"] * (h - 1) + (p["dh"] * (p["kh"] - 1) + 1) - 2 * p["ph"]      # deconv_layer.cpp:8-28
                wo = p["sw"] * (w - 1) + (p["dw"] * (p["kw"] - 1) + 1) - 2 * p["pw"]
                out = (n, p["num_output"], ho, wo)
            elif t == "Pooling":
                n, c, h, w = bs[0]
                p = spec.p
                ho = -((h + 2 * p["ph"] - p["kh"]) // -p["sh"]) + 1                       # ceil, pooling_layer.cpp:91-94
                wo = -((w + 2 * p["pw"] - p["kw"]) // -p["sw"]) + 1
                if p["ph"] or p["pw"]:                                                  # pooling_layer.cpp:95-106
                    if (ho - 1) * p["sh"] >= h + p["ph"]:
                        ho -= 1
                    if (wo - 1) * p["sw"] >= w + p["pw"]:
                        wo -= 1
                out = (n, c, ho, wo)
            elif t in ("ReLU", "Softmax"):
                out = bs[0]
            elif t == "Eltwise":
                for b in bs[1:]:
                    if tuple(b) != tuple(bs[0]):                 # eltwise_layer.cpp:27-30
                        raise ValueError("%s: bottom shapes differ (%s vs %s)" % (spec.name, bs[0], b))
                out = bs[0]
            elif t in ("BatchNorm", "Scale"):
                self.check_param_shape(spec, bs[0][1] if len(bs[0]) > 1 else 1)
                out = bs[0]
            elif t == "Split":
                for nm in spec.tops:
                    shapes[nm] = bs[0]
                continue
            elif t == "Concat":
                ax = spec.p["axis"]
                ax = ax + len(bs[0]) if ax < 0 else ax
                for b in bs[1:]:
                    if len(b) != len(bs[0]) or any(b[i] != bs[0][i] for i in range(len(b)) if i != ax):
                        raise ValueError("%s: all inputs must have the same shape, except at concat_axis "
                                         "(%s vs %s)" % (spec.name, bs[0], b))       # concat_layer.cpp:40-44
                out = tuple(sum(b[ax] for b in bs) if i == ax else d for i, d in enumerate(bs[0]))
            elif t == "Reshape":
                out = reshape_shape(bs[0], spec.p["dims"], spec.p["axis"], spec.p["num_axes"], spec.name)
            elif t == "Python":
                if (spec.p["module"], spec.p["layer"]) == PROPOSAL_LAYER:
                    # ProposalLayer.setup: rois (1,5), scores (1,1,1,1) until the first forward
                    for i, nm in enumerate(spec.tops):
                        shapes[nm] = (1, 5) if i == 0 else (1, 2)
                    continue
                # any other Python layer: the generic protocol (python_layer.hpp:19-43) -- instantiate once (setup),
                # then ask its reshape() for the top shapes, exactly what Net::Reshape does
                st = self.py_layers.get(spec.name)
                if st is None:
                    st = self.py_layers[spec.name] = make_python_layer(spec, self.phase, bs)
                for blob, shp in zip(st["bottoms"], bs):
                    blob.reshape(*shp)
                st["obj"].reshape(st["bottoms"], st["tops"])
                for nm, blob in zip(spec.tops, st["tops"]):
                    shapes[nm] = blob.shape
                continue
            else:                                                   # pragma: no cover
                raise AssertionError(t)
            shapes[spec.tops[0]] = out
        return shapes


def fold_batchnorm_scale(w, b, chain):
    """Folds a Convolution's TEST-phase BatchNorm / Scale followers into its weights and bias.
    ``chain``: list of ("BatchNorm", mean, var, factor, eps) / ("Scale", gamma, beta-or-None) in layer order.
    BatchNorm with global stats (batch_norm_layer.cpp:98-104,147-165): s = 0 if factor == 0 else 1 / factor;
    y = (x - mean*s) / sqrt(var*s + eps).  Scale (scale_layer.cpp): y = x * gamma + beta.  Everything is affine per
    output channel, so conv -> BN -> Scale == conv with w' = w * a, b' = b * a + c (float64 arithmetic, rounded once)."""
    w = np.asarray(w, dtype=np.float64)
    co = w.shape[0]
    a = np.ones(co, dtype=np.float64)
    c = np.zeros(co, dtype=np.float64) if b is None else np.asarray(b, dtype=np.float64).copy()
    for item in chain:
        if item[0] == "BatchNorm":
            _, mean, var, factor, eps = item
            f = float(np.asarray(factor).reshape(-1)[0])
            s = 0.0 if f == 0 else 1.0 / f
            inv = 1.0 / np.sqrt(np.asarray(var, np.float64) * s + eps)
            a = a * inv
            c = (c - np.asarray(mean, np.float64) * s) * inv
        elif item[0] == "Scale":
            _, gamma, beta = item
            g = np.asarray(gamma, np.float64)
            a = a * g
            c = c * g + (0.0 if beta is None else np.asarray(beta, np.float64))
        else:                                                       # pragma: no cover
            raise AssertionError(item[0])
    return (w * a[:, None, None, None]).astype(np.float32), c.astype(np.float32)


def reshape_shape(bottom, dims, axis, num_axes, lname="reshape"):
    """``reshape_layer.cpp:32-84``: 0 copies the bottom dim, one -1 is inferred."""
    nb = len(bottom)
    start = axis if axis >= 0 else nb + axis + 1
    end = nb if num_axes == -1 else start + num_axes
    if not (0 <= start <= nb) or end > nb:
        raise ValueError("%s: axis / num_axes out of range" % lname)
    top = list(bottom[:start])
    infer = -1
    const = 1
    for i, d in enumerate(dims):
        if d == 0:
            if start + i >= nb:
                raise ValueError("%s: dimension %d out of range for copy" % (lname, i))
            top.append(bottom[start + i])
        elif d == -1:
            if infer != -1:
                raise ValueError("%s: at most one -1" % lname)
            infer = len(top)
            top.append(-1)
        else:
            top.append(d)
            const *= d
    top += list(bottom[end:])
    if infer >= 0:
        known = 1
        for i, d in enumerate(top):
            if i != infer:
                known *= d
        total = int(np.prod(bottom))
        if known == 0 or total % known:
            raise ValueError("%s: bottom count %d must be divisible by the product of the "
                             "specified dimensions %d" % (lname, total, known))
        top[infer] = total // known
    if int(np.prod(top)) != int(np.prod(bottom)):
        raise ValueError("%s: output count must match input count" % lname)
    return tuple(int(d) for d in top)


def parse_python_param_str(s: str) -> dict:
    """The ProposalLayer's ``param_str`` is YAML flow syntax that is also a python literal
    (``proposal_layer.py:21-24``); parse it without PyYAML's ``Loader`` pitfalls."""
    if not s.strip():
        return {}
    try:
        return dict(ast.literal_eval(s))
    except Exception:
        import yaml
        return dict(yaml.safe_load(s))


def load_weights(spec: NetSpec, model: cp.Msg, input_shapes=None) -> Dict[str, np.ndarray]:
    """``Net::CopyTrainedLayersFrom`` (``net.cpp:733-785``): match by layer name, shapes must be
    identical, source layers unknown to the net are ignored; shared params land in the owner's
    storage (later layers in file order win, as in Caffe where they alias one Blob)."""
    spec.infer_shapes(input_shapes or {})
    by_name = {l.name: l for l in spec.layers}
    params: Dict[str, np.ndarray] = {}
    for src in model.layer:
        tgt = by_name.get(src.name)
        if tgt is None or not tgt.param_keys:
            continue
        if len(src.blobs) != len(tgt.param_keys):
            raise RuntimeError("Incompatible number of blobs for layer %s" % src.name)   # net.cpp:751-752
        for key, bp in zip(tgt.param_keys, src.blobs):
            arr = cp.array_from_blob(bp)
            want = spec.param_shapes[key]
            if tuple(arr.shape) != tuple(want):
                # legacy 4-D bias blobs (1,1,1,N) are accepted by Blob::ShapeEquals (blob.cpp:454-470)
                if bp.has("shape") or arr.size != int(np.prod(want)) or arr.ndim != 4:
                    raise RuntimeError(
                        "Cannot copy param %s weights from layer '%s'; shape mismatch.  Source param "
                        "shape is %s; target param shape is %s." % (key, src.name, arr.shape, want))
                arr = arr.reshape(want)
            params[key] = np.ascontiguousarray(arr, dtype=np.float32)
    for key, shp in spec.param_shapes.items():
        if key not in params:                                        # FillerParameter default: constant 0
            params[key] = np.zeros(shp, dtype=np.float32)
    return params
