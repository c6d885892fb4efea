"""ORACLE (test infrastructure, not product code) -- a SECOND, independent reading of a deployment.

oracle/net.py walks the graph that smallhardface_b200.graph / caffe_proto build, i.e. it shares the
prototxt parser, the wire decoder, the in-place / shared-parameter resolution and the tail wiring with
the product: a wiring mistake there would be common-mode and invisible to every parity test.  This
module shares NONE of that code.  It has

  * its own schema-less protobuf TEXT reader (``name: value`` / ``name { ... }`` -> nested lists),
  * its own protobuf WIRE reader with the handful of field numbers it needs, restated here straight
    from ``caffe/src/caffe/proto/caffe.proto`` (NetParameter.layer = 100 ``:92``, LayerParameter.name = 1
    / .blobs = 7 ``:313,331``, BlobProto.shape = 7 / .data = 5 / legacy num..width = 1..4 ``:10-22``,
    BlobShape.dim = 1 ``:5-8``),
  * its own net walk, restating ``Net::Init`` / ``ForwardFromTo`` (``caffe/src/caffe/net.cpp:44-256,
    516-532``) as a name -> array dict: a top that reuses its bottom's name overwrites it (in-place
    layers, ``net.cpp:363-370``), ``param { name: ".." }`` entries alias the first layer that used the
    name (``net.cpp:457-512``), weights are matched by LAYER NAME (``net.cpp:733-785``), blobs nobody
    consumes are the outputs (``net.cpp:240-246``).

Only the layer ARITHMETIC (oracle/layers.py, pinned by Caffe's own known-answer tests) and the
ProposalLayer math (oracle/proposal.py, pinned bit-exactly by the reference's own output) are reused.
tests/test_oracle_indep.py requires both readings to agree on every blob of both templates; the GPU
parity tests compare the CUDA path against THIS reading.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

from . import layers as L
from .proposal import proposal_forward

F32 = np.float32


# ------------------------------------------------------------------------------------------------
# protobuf text format, schema-less: message -> list of (field, value); value = scalar or message
# ------------------------------------------------------------------------------------------------
def _text_tokens(text):
    i, n = 0, len(text)
    while i < n:
        c = text[i]
        if c.isspace():
            i += 1
        elif c == "#":
            while i < n and text[i] != "\n":
                i += 1
        elif c in "{}:<>[],;":
            yield c
            i += 1
        elif c in "\"'":
            j, out = i + 1, []
            while text[j] != c:
                if text[j] == "\\":
                    j += 1
                    out.append({"n": "\n", "t": "\t", "\\": "\\", "'": "'", '"': '"'}.get(text[j], text[j]))
                else:
                    out.append(text[j])
                j += 1
            yield ("str", "".join(out))
            i = j + 1
        else:
            j = i
            while j < n and not text[j].isspace() and text[j] not in "{}:<>[],;#\"'":
                j += 1
            yield text[i:j]
            i = j


def _scalar(tok):
    if isinstance(tok, tuple):
        return tok[1]
    if tok in ("true", "false"):
        return tok == "true"
    try:
        return int(tok, 0)
    except ValueError:
        pass
    try:
        return float(tok.rstrip("f"))
    except ValueError:
        return tok                     # enum identifier


def parse_prototxt(text):
    toks = list(_text_tokens(text))
    pos = [0]

    def message(closer):
        fields = []
        while pos[0] < len(toks) and toks[pos[0]] != closer:
            name = toks[pos[0]]
            pos[0] += 1
            if toks[pos[0]] == ":":
                pos[0] += 1
            if toks[pos[0]] in ("{", "<"):
                close = "}" if toks[pos[0]] == "{" else ">"
                pos[0] += 1
                fields.append((name, message(close)))
                pos[0] += 1
            elif toks[pos[0]] == "[":
                pos[0] += 1
                while toks[pos[0]] != "]":
                    if toks[pos[0]] != ",":
                        fields.append((name, _scalar(toks[pos[0]])))
                    pos[0] += 1
                pos[0] += 1
            else:
                val = _scalar(toks[pos[0]])
                pos[0] += 1
                # adjacent string literals concatenate
                while isinstance(val, str) and pos[0] < len(toks) and isinstance(toks[pos[0]], tuple):
                    val += toks[pos[0]][1]
                    pos[0] += 1
                fields.append((name, val))
            if pos[0] < len(toks) and toks[pos[0]] in (",", ";"):
                pos[0] += 1
        return fields

    return message(None)


def _all(msg, name):
    return [v for k, v in msg if k == name]


def _one(msg, name, default=None):
    vals = _all(msg, name)
    return vals[-1] if vals else default


# ------------------------------------------------------------------------------------------------
# protobuf wire format: just enough for NetParameter -> {layer name: [blob arrays]}
# ------------------------------------------------------------------------------------------------
def _varint(buf, i):
    shift = val = 0
    while True:
        b = buf[i]
        i += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, i
        shift += 7


def _wire_fields(buf, lo, hi):
    i = lo
    while i < hi:
        tag, i = _varint(buf, i)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, i = _varint(buf, i)
            yield num, wt, v
        elif wt == 1:
            yield num, wt, (i, i + 8)
            i += 8
        elif wt == 2:
            ln, i = _varint(buf, i)
            yield num, wt, (i, i + ln)
            i += ln
        elif wt == 5:
            yield num, wt, (i, i + 4)
            i += 4
        else:
            raise ValueError("wire type %d" % wt)


def _blob(buf, lo, hi):
    dims, legacy, chunks = None, {}, []
    for num, wt, v in _wire_fields(buf, lo, hi):
        if num == 7 and wt == 2:                          # BlobShape { repeated int64 dim = 1 [packed] }
            dims = []
            for n2, w2, v2 in _wire_fields(buf, v[0], v[1]):
                if n2 == 1 and w2 == 2:
                    j = v2[0]
                    while j < v2[1]:
                        d, j = _varint(buf, j)
                        dims.append(d)
                elif n2 == 1 and w2 == 0:
                    dims.append(v2)
        elif num == 5 and wt == 2:                        # packed float data
            chunks.append(np.frombuffer(buf, dtype="<f4", count=(v[1] - v[0]) // 4, offset=v[0]))
        elif num == 5 and wt == 5:
            chunks.append(np.array(struct.unpack_from("<f", buf, v[0]), dtype=F32))
        elif num in (1, 2, 3, 4) and wt == 0:
            legacy[num] = v
    data = np.concatenate(chunks) if chunks else np.zeros(0, F32)
    if dims is None:
        dims = [legacy.get(k, 0) for k in (1, 2, 3, 4)] if legacy else [data.size]
    return dims, data


def read_caffemodel(path):
    with open(path, "rb") as f:
        buf = f.read()
    layers = OrderedDict()
    for num, wt, v in _wire_fields(buf, 0, len(buf)):
        if num != 100 or wt != 2:                         # NetParameter.layer
            continue
        name, blobs = None, []
        for n2, w2, v2 in _wire_fields(buf, v[0], v[1]):
            if n2 == 1 and w2 == 2:
                name = bytes(buf[v2[0]:v2[1]]).decode()
            elif n2 == 7 and w2 == 2:
                blobs.append(_blob(buf, v2[0], v2[1]))
        layers[name] = blobs
    return layers


# ------------------------------------------------------------------------------------------------
# the net
# ------------------------------------------------------------------------------------------------
def _hw(msg, rep, h, w, default):
    """ConvolutionParameter's repeated / _h,_w spellings (``base_conv_layer.cpp:33-92``)."""
    r = _all(msg, rep)
    if _one(msg, h) is not None or _one(msg, w) is not None:
        return int(_one(msg, h, default)), int(_one(msg, w, default))
    if not r:
        return default, default
    return (int(r[0]), int(r[0])) if len(r) == 1 else (int(r[0]), int(r[1]))


class IndepNet:
    """Same call surface as oracle.net.OracleNet (``forward(**inputs)``, ``.blobs``) over the independent reading."""

    def __init__(self, prototxt_path, caffemodel_path=None, engine="torch", pre_nms_topn=10000, score_thresh=0.002,
                 min_size=0):
        with open(prototxt_path) as f:
            self.net = parse_prototxt(f.read())
        self.stored = read_caffemodel(caffemodel_path) if caffemodel_path else {}
        self.engine = engine
        self.cfg = dict(pre_nms_topn=pre_nms_topn, score_thresh=score_thresh, min_size=min_size)
        self.inputs = list(_all(self.net, "input"))
        self.layers = [l for l in _all(self.net, "layer") if self._in_phase(l)]
        for l in self.layers:
            if _one(l, "type") == "Input":
                self.inputs += _all(l, "top")
        self._storage = None
        self._arrays = {}
        self.blobs = OrderedDict()
        self.outputs = []

    @staticmethod
    def _in_phase(l):
        """TEST-phase filter (``net.cpp:259-287``; the deploy nets carry no include/exclude rules, but be exact)."""
        inc = [r for r in _all(l, "include")]
        exc = [r for r in _all(l, "exclude")]
        ph = _one(l, "phase")
        if ph is not None and not inc and not exc and ph != "TEST":
            return False
        if inc and not any(_one(r, "phase", "TEST") == "TEST" for r in inc):
            return False
        if exc and any(_one(r, "phase") == "TEST" for r in exc):
            return False
        return True

    def _storage_key(self, layer, idx):
        """``param { name: "x" }`` entries alias one blob per name (``net.cpp:457-512``); unnamed ones are private."""
        specs = _all(layer, "param")
        pname = _one(specs[idx], "name") if idx < len(specs) else None
        return ("shared", pname) if pname else (_one(layer, "name"), idx)

    def _load(self):
        """``Net::CopyTrainedLayersFrom`` (``net.cpp:733-785``): source layers in FILE order, matched by layer name; a
        shared blob written by several source layers keeps the last one."""
        by_name = {_one(l, "name"): l for l in self.layers}
        storage = {}
        for src, blobs in self.stored.items():
            l = by_name.get(src)
            if l is None:
                continue
            for idx, (dims, data) in enumerate(blobs):
                storage[self._storage_key(l, idx)] = (dims, data)
        return storage

    def _param(self, layer, idx, shape):
        """Blob ``idx`` of ``layer``; absent from the caffemodel = the default filler, constant 0 (``caffe.proto:45-46``)."""
        if self._storage is None:
            self._storage = self._load()
        key = self._storage_key(layer, idx)
        cached = self._arrays.get((key, tuple(shape)))
        if cached is not None:
            return cached
        hit = self._storage.get(key)
        if hit is None:
            arr = np.zeros(shape, F32)
        else:
            dims, data = hit
            if data.size != int(np.prod(shape)) or int(np.prod(dims)) != data.size:
                raise RuntimeError("shape mismatch for %s blob %d: %s vs %s" % (_one(layer, "name"), idx, dims, shape))
            arr = np.array(data, dtype=F32).reshape(shape)
        self._arrays[(key, tuple(shape))] = arr
        return arr

    def forward(self, **inputs):
        if set(inputs) != set(self.inputs):
            raise Exception("Input blob arguments do not match net inputs.")
        b = self.blobs = OrderedDict((k, np.ascontiguousarray(v, F32)) for k, v in inputs.items())
        consumed = set()
        for l in self.layers:
            t, name = _one(l, "type"), _one(l, "name")
            bottoms, tops = _all(l, "bottom"), _all(l, "top")
            consumed.update(bottoms)
            xs = [b[n] for n in bottoms]
            if t == "Input":
                continue
            if t in ("Convolution", "Deconvolution"):
                c = _one(l, "convolution_param")
                co, g = int(_one(c, "num_output")), int(_one(c, "group", 1))
                kh, kw = _hw(c, "kernel_size", "kernel_h", "kernel_w", None)
                ph, pw = _hw(c, "pad", "pad_h", "pad_w", 0)
                sh, sw = _hw(c, "stride", "stride_h", "stride_w", 1)
                d = _all(c, "dilation")
                dh, dw = (int(d[0]), int(d[-1])) if d else (1, 1)
                cin = xs[0].shape[1]
                wshape = (co, cin // g, kh, kw) if t == "Convolution" else (cin, co // g, kh, kw)
                w = self._param(l, 0, wshape)
                bias = self._param(l, 1, (co,)) if _one(c, "bias_term", True) else None
                if t == "Convolution":
                    y = L.conv(xs[0], w, bias, (ph, pw), (sh, sw), (dh, dw), g, engine=self.engine)
                elif g == cin == co and bias is None:
                    y = L.deconv_depthwise_fast(xs[0], w, (ph, pw), (sh, sw))
                else:
                    y = L.deconv(xs[0], w, bias, (ph, pw), (sh, sw), (dh, dw), g)
            elif t == "ReLU":
                r = _one(l, "relu_param", [])
                y = L.relu(xs[0], float(_one(r, "negative_slope", 0.0)))
            elif t == "Pooling":
                q = _one(l, "pooling_param")
                if _one(q, "pool", "MAX") != "MAX":
                    raise ValueError("only MAX pooling is on the hot path")
                k, s, p = int(_one(q, "kernel_size")), int(_one(q, "stride", 1)), int(_one(q, "pad", 0))
                if (k, s, p) == (2, 2, 0) and xs[0].shape[2] % 2 == 0 and xs[0].shape[3] % 2 == 0:
                    y = L.max_pool_2x2_fast(xs[0])
                else:
                    y = L.max_pool(xs[0], (k, k), (s, s), (p, p))
            elif t == "Eltwise":
                q = _one(l, "eltwise_param", [])
                if _one(q, "operation", "SUM") != "SUM":
                    raise ValueError("only Eltwise SUM is on the hot path")
                y = L.eltwise_sum(xs, [float(v) for v in _all(q, "coeff")])
            elif t == "Concat":
                q = _one(l, "concat_param", [])
                y = np.concatenate(xs, axis=int(_one(q, "axis", _one(q, "concat_dim", 1))))
            elif t == "Reshape":
                dims = [int(v) for v in _all(_one(_one(l, "reshape_param"), "shape"), "dim")]
                shp = [xs[0].shape[i] if v == 0 else v for i, v in enumerate(dims)]      # reshape_layer.cpp:32-84
                y = xs[0].reshape(shp)
            elif t == "Softmax":
                q = _one(l, "softmax_param", [])
                y = L.softmax(xs[0], int(_one(q, "axis", 1)))
            elif t == "BatchNorm":
                q = _one(l, "batch_norm_param", [])
                c = xs[0].shape[1]
                y = L.batch_norm(xs[0], self._param(l, 0, (c,)), self._param(l, 1, (c,)), self._param(l, 2, (1,)),
                                 float(_one(q, "eps", 1e-5)))
            elif t == "Scale":
                q = _one(l, "scale_param", [])
                c = xs[0].shape[1]
                y = L.scale(xs[0], self._param(l, 0, (c,)), self._param(l, 1, (c,)) if _one(q, "bias_term", False) else None)
            elif t == "Python":
                q = _one(l, "python_param")
                if (_one(q, "module"), _one(q, "layer")) != ("lib.layers.proposal_layer", "ProposalLayer"):
                    raise ValueError("the oracle only knows the ProposalLayer python layer")
                import yaml
                lp = yaml.safe_load(_one(q, "param_str"))              # proposal_layer.py:21-24
                boxes, probs, _ = proposal_forward(
                    xs[0], xs[1], xs[2], feat_stride=tuple(lp["feat_stride"]), scales=tuple(lp.get("scales", (8, 16, 32))),
                    ratios=tuple(lp.get("ratios", (0.5, 1, 2))), base_size=lp.get("base_size", 16),
                    shifts=tuple(lp.get("shifts", [0])), **self.cfg)
                b[tops[0]] = boxes
                if len(tops) > 1:
                    b[tops[1]] = probs
                continue
            else:
                raise ValueError("layer type %r is outside the inference hot path" % t)
            b[tops[0]] = y
        self.outputs = [k for k in b if k not in consumed]
        return {o: b[o] for o in self.outputs}
