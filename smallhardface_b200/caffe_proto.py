"""Caffe on-disk formats without protoc: text ``.prototxt`` and binary ``.caffemodel``.

The reference reads both through generated protobuf code
(``caffe/src/caffe/util/io.cpp:34-65`` -> ``ReadProtoFromTextFile`` /
``ReadProtoFromBinaryFile``; schema ``caffe/src/caffe/proto/caffe.proto``).  This
image has no ``protoc``, so the subset of the schema the deploy nets use is restated
here as a field table (field numbers / wire types are the format, see SURVEY.md
Appendix C) and both codecs are written directly against it.

Public API
----------
``Msg``                       schema-typed message (attribute access, repeated fields are lists)
``parse_text(s, 'NetParameter')``   protobuf text format -> Msg
``format_text(msg)``                 Msg -> protobuf text format
``decode(buf, 'NetParameter')``      binary wire format -> Msg  (packed float blobs -> numpy)
``encode(msg)``                      Msg -> binary wire format
``read_net_text(path)`` / ``read_net_binary(path)`` / ``write_net_binary(path, msg)``
"""
from __future__ import annotations

import re
import struct
from typing import Any, Dict, List, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# Schema: message -> {field: (number, type, label)}; label: 'o' optional, 'r' repeated,
# 'p' repeated+packed.  Scalar types: int32 int64 uint32 bool float double string enum:<E>.
# Field numbers restate caffe.proto (cited per message).
# --------------------------------------------------------------------------------------
ENUMS: Dict[str, Dict[str, int]] = {
    "Phase": {"TRAIN": 0, "TEST": 1},                                   # caffe.proto:254-257
    "DimCheckMode": {"STRICT": 0, "PERMISSIVE": 1},                      # caffe.proto:294-299
    "Engine": {"DEFAULT": 0, "CAFFE": 1, "CUDNN": 2},                    # caffe.proto:600-604
    "PoolMethod": {"MAX": 0, "AVE": 1, "STOCHASTIC": 2},                 # caffe.proto:919-923
    "VarianceNorm": {"FAN_IN": 0, "FAN_OUT": 1, "AVERAGE": 2},           # caffe.proto:56-60
    "EltwiseOp": {"PROD": 0, "SUM": 1, "MAX": 2},                        # caffe.proto:702-706
    "SolverMode": {"CPU": 0, "GPU": 1},
    "SnapshotFormat": {"HDF5": 0, "BINARYPROTO": 1},
}

SCHEMA: Dict[str, Dict[str, Tuple[int, str, str]]] = {
    "BlobShape": {"dim": (1, "int64", "p")},                             # caffe.proto:6-8
    "BlobProto": {                                                       # caffe.proto:10-22
        "num": (1, "int32", "o"), "channels": (2, "int32", "o"),
        "height": (3, "int32", "o"), "width": (4, "int32", "o"),
        "data": (5, "float", "p"), "diff": (6, "float", "p"),
        "shape": (7, "BlobShape", "o"),
        "double_data": (8, "double", "p"), "double_diff": (9, "double", "p"),
    },
    "FillerParameter": {                                                 # caffe.proto:43-62
        "type": (1, "string", "o"), "value": (2, "float", "o"), "min": (3, "float", "o"),
        "max": (4, "float", "o"), "mean": (5, "float", "o"), "std": (6, "float", "o"),
        "sparse": (7, "int32", "o"), "variance_norm": (8, "enum:VarianceNorm", "o"),
    },
    "NetState": {                                                        # caffe.proto:259-263
        "phase": (1, "enum:Phase", "o"), "level": (2, "int32", "o"), "stage": (3, "string", "r"),
    },
    "NetStateRule": {                                                    # caffe.proto:265-283
        "phase": (1, "enum:Phase", "o"), "min_level": (2, "int32", "o"),
        "max_level": (3, "int32", "o"), "stage": (4, "string", "r"), "not_stage": (5, "string", "r"),
    },
    "ParamSpec": {                                                       # caffe.proto:285-305
        "name": (1, "string", "o"), "share_mode": (2, "enum:DimCheckMode", "o"),
        "lr_mult": (3, "float", "o"), "decay_mult": (4, "float", "o"),
    },
    "NetParameter": {                                                    # caffe.proto:64-95
        "name": (1, "string", "o"), "input": (3, "string", "r"),
        "input_dim": (4, "int32", "r"), "force_backward": (5, "bool", "o"),
        "state": (6, "NetState", "o"), "debug_info": (7, "bool", "o"),
        "input_shape": (8, "BlobShape", "r"), "layer": (100, "LayerParameter", "r"),
    },
    "LayerParameter": {                                                  # caffe.proto:312-410
        "name": (1, "string", "o"), "type": (2, "string", "o"),
        "bottom": (3, "string", "r"), "top": (4, "string", "r"),
        "loss_weight": (5, "float", "r"), "param": (6, "ParamSpec", "r"),
        "blobs": (7, "BlobProto", "r"), "include": (8, "NetStateRule", "r"),
        "exclude": (9, "NetStateRule", "r"), "phase": (10, "enum:Phase", "o"),
        "propagate_down": (11, "bool", "r"),
        "concat_param": (104, "ConcatParameter", "o"),
        "convolution_param": (106, "ConvolutionParameter", "o"),
        "eltwise_param": (110, "EltwiseParameter", "o"),
        "pooling_param": (121, "PoolingParameter", "o"),
        "relu_param": (123, "ReLUParameter", "o"),
        "softmax_param": (125, "SoftmaxParameter", "o"),
        "python_param": (130, "PythonParameter", "o"),
        "reshape_param": (133, "ReshapeParameter", "o"),
        "batch_norm_param": (139, "BatchNormParameter", "o"),
        "scale_param": (142, "ScaleParameter", "o"),
        "input_param": (143, "InputParameter", "o"),
    },
    "BatchNormParameter": {                                              # caffe.proto:507-527
        "use_global_stats": (1, "bool", "o"), "moving_average_fraction": (2, "float", "o"), "eps": (3, "float", "o"),
    },
    "ScaleParameter": {                                                  # caffe.proto:1099-1129
        "axis": (1, "int32", "o"), "num_axes": (2, "int32", "o"), "filler": (3, "FillerParameter", "o"),
        "bias_term": (4, "bool", "o"), "bias_filler": (5, "FillerParameter", "o"),
    },
    "ConcatParameter": {"concat_dim": (1, "uint32", "o"), "axis": (2, "int32", "o")},  # :496-505
    "EltwiseParameter": {                                                # caffe.proto:701-713
        "operation": (1, "enum:EltwiseOp", "o"), "coeff": (2, "float", "r"), "stable_prod_grad": (3, "bool", "o"),
    },
    "ConvolutionParameter": {                                            # caffe.proto:573-624
        "num_output": (1, "uint32", "o"), "bias_term": (2, "bool", "o"),
        "pad": (3, "uint32", "r"), "kernel_size": (4, "uint32", "r"),
        "group": (5, "uint32", "o"), "stride": (6, "uint32", "r"),
        "weight_filler": (7, "FillerParameter", "o"), "bias_filler": (8, "FillerParameter", "o"),
        "pad_h": (9, "uint32", "o"), "pad_w": (10, "uint32", "o"),
        "kernel_h": (11, "uint32", "o"), "kernel_w": (12, "uint32", "o"),
        "stride_h": (13, "uint32", "o"), "stride_w": (14, "uint32", "o"),
        "engine": (15, "enum:Engine", "o"), "axis": (16, "int32", "o"),
        "force_nd_im2col": (17, "bool", "o"), "dilation": (18, "uint32", "r"),
    },
    "InputParameter": {"shape": (1, "BlobShape", "r")},                   # caffe.proto:841-848
    "PoolingParameter": {                                                # caffe.proto:918-945
        "pool": (1, "enum:PoolMethod", "o"), "kernel_size": (2, "uint32", "o"),
        "stride": (3, "uint32", "o"), "pad": (4, "uint32", "o"),
        "kernel_h": (5, "uint32", "o"), "kernel_w": (6, "uint32", "o"),
        "stride_h": (7, "uint32", "o"), "stride_w": (8, "uint32", "o"),
        "pad_h": (9, "uint32", "o"), "pad_w": (10, "uint32", "o"),
        "engine": (11, "enum:Engine", "o"), "global_pooling": (12, "bool", "o"),
    },
    "PythonParameter": {                                                 # caffe.proto:954-965
        "module": (1, "string", "o"), "layer": (2, "string", "o"),
        "param_str": (3, "string", "o"), "share_in_parallel": (4, "bool", "o"),
    },
    "ReLUParameter": {"negative_slope": (1, "float", "o"), "engine": (2, "enum:Engine", "o")},
    "ReshapeParameter": {                                                # caffe.proto:1030-1092
        "shape": (1, "BlobShape", "o"), "axis": (2, "int32", "o"), "num_axes": (3, "int32", "o"),
    },
    "SoftmaxParameter": {"engine": (1, "enum:Engine", "o"), "axis": (2, "int32", "o")},
    # SolverParameter subset that lib/prototxt/manipulate.py:13-32 touches (caffe.proto:102-245)
    "SolverParameter": {
        "train_net": (1, "string", "o"), "test_iter": (3, "int32", "r"),
        "test_interval": (4, "int32", "o"), "base_lr": (5, "float", "o"),
        "display": (6, "int32", "o"), "max_iter": (7, "int32", "o"),
        "lr_policy": (8, "string", "o"), "gamma": (9, "float", "o"), "power": (10, "float", "o"),
        "momentum": (11, "float", "o"), "weight_decay": (12, "float", "o"),
        "stepsize": (13, "int32", "o"), "snapshot": (14, "int32", "o"),
        "snapshot_prefix": (15, "string", "o"), "solver_mode": (17, "enum:SolverMode", "o"),
        "device_id": (18, "int32", "o"), "random_seed": (20, "int64", "o"),
        "net": (24, "string", "o"), "average_loss": (33, "int32", "o"),
        "stepvalue": (34, "int32", "r"), "iter_size": (36, "int32", "o"),
        "snapshot_format": (37, "enum:SnapshotFormat", "o"), "type": (40, "string", "o"),
    },
}

# proto2 defaults the hot path relies on (caffe.proto, same lines as above)
DEFAULTS: Dict[Tuple[str, str], Any] = {
    ("ConvolutionParameter", "bias_term"): True, ("ConvolutionParameter", "group"): 1,
    ("ConvolutionParameter", "pad_h"): 0, ("ConvolutionParameter", "pad_w"): 0,
    ("ConvolutionParameter", "axis"): 1, ("ConvolutionParameter", "engine"): 0,
    ("ConvolutionParameter", "force_nd_im2col"): False,
    ("PoolingParameter", "pool"): 0, ("PoolingParameter", "pad"): 0,
    ("PoolingParameter", "pad_h"): 0, ("PoolingParameter", "pad_w"): 0,
    ("PoolingParameter", "stride"): 1, ("PoolingParameter", "global_pooling"): False,
    ("PoolingParameter", "engine"): 0,
    ("ConcatParameter", "axis"): 1, ("ConcatParameter", "concat_dim"): 1,
    ("EltwiseParameter", "operation"): 1, ("EltwiseParameter", "stable_prod_grad"): True,
    ("SoftmaxParameter", "axis"): 1, ("SoftmaxParameter", "engine"): 0,
    ("ReLUParameter", "negative_slope"): 0.0, ("ReLUParameter", "engine"): 0,
    ("ReshapeParameter", "axis"): 0, ("ReshapeParameter", "num_axes"): -1,
    ("BatchNormParameter", "moving_average_fraction"): 0.999, ("BatchNormParameter", "eps"): 1e-5,
    ("ScaleParameter", "axis"): 1, ("ScaleParameter", "num_axes"): 1, ("ScaleParameter", "bias_term"): False,
    ("PythonParameter", "param_str"): "", ("PythonParameter", "share_in_parallel"): False,
    ("FillerParameter", "type"): "constant", ("FillerParameter", "value"): 0.0,
    ("FillerParameter", "min"): 0.0, ("FillerParameter", "max"): 1.0,
    ("FillerParameter", "mean"): 0.0, ("FillerParameter", "std"): 1.0,
    ("FillerParameter", "sparse"): -1, ("FillerParameter", "variance_norm"): 0,
    ("ParamSpec", "lr_mult"): 1.0, ("ParamSpec", "decay_mult"): 1.0,
    ("NetState", "phase"): 1, ("NetState", "level"): 0,
    ("BlobProto", "num"): 0, ("BlobProto", "channels"): 0,
    ("BlobProto", "height"): 0, ("BlobProto", "width"): 0,
    ("NetParameter", "force_backward"): False, ("NetParameter", "debug_info"): False,
}

_BY_NUMBER: Dict[str, Dict[int, Tuple[str, str, str]]] = {
    m: {num: (name, typ, lab) for name, (num, typ, lab) in fields.items()}
    for m, fields in SCHEMA.items()
}


class Msg:
    """A schema-typed protobuf message.  Repeated fields are python lists (packed float/double
    fields of BlobProto decode to numpy arrays); unset optional fields read as their proto2
    default (or None) and ``has(name)`` tells the difference."""

    __slots__ = ("_type", "_f")

    def __init__(self, type_name: str, **kw):
        if type_name not in SCHEMA:
            raise KeyError("unknown message type %r" % type_name)
        object.__setattr__(self, "_type", type_name)
        object.__setattr__(self, "_f", {})
        for k, v in kw.items():
            setattr(self, k, v)

    # -- attribute protocol
    def __getattr__(self, name):
        sch = SCHEMA[self._type]
        if name not in sch:
            raise AttributeError("%s has no field %r" % (self._type, name))
        f = self._f
        if name in f:
            return f[name]
        _, typ, lab = sch[name]
        if lab in ("r", "p"):
            f[name] = []
            return f[name]
        if typ in SCHEMA:                       # unset sub-message: lazily created (protobuf semantics)
            sub = Msg(typ)
            f[name] = sub
            return sub
        return DEFAULTS.get((self._type, name))

    def __setattr__(self, name, value):
        sch = SCHEMA[self._type]
        if name not in sch:
            raise AttributeError("%s has no field %r" % (self._type, name))
        self._f[name] = value

    def has(self, name: str) -> bool:
        v = self._f.get(name)
        if v is None:
            return False
        if isinstance(v, Msg):
            return True
        if isinstance(v, (list, np.ndarray)):
            return len(v) > 0
        return True

    def clear(self, name: str) -> None:
        self._f.pop(name, None)

    @property
    def type_name(self) -> str:
        return self._type

    def fields(self):
        """(name, value) for set fields in field-number order."""
        sch = SCHEMA[self._type]
        for name in sorted(self._f, key=lambda n: sch[n][0]):
            if self.has(name):
                yield name, self._f[name]

    def copy(self) -> "Msg":
        out = Msg(self._type)
        for k, v in self._f.items():
            if isinstance(v, Msg):
                out._f[k] = v.copy()
            elif isinstance(v, list):
                out._f[k] = [x.copy() if isinstance(x, Msg) else x for x in v]
            elif isinstance(v, np.ndarray):
                out._f[k] = v.copy()
            else:
                out._f[k] = v
        return out

    def __eq__(self, other):
        return isinstance(other, Msg) and format_text(self) == format_text(other)

    def __repr__(self):
        return "<%s %s>" % (self._type, format_text(self)[:120].replace("\n", " "))


# --------------------------------------------------------------------------------------
# Text format
# --------------------------------------------------------------------------------------
_TOKEN = re.compile(
    r"""\s*(?:(\#[^\n]*)            # comment
        |([A-Za-z_][A-Za-z0-9_\.]*)  # identifier
        |("(?:[^"\\]|\\.)*"|'(?:[^'\\]|\\.)*')   # string
        |([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf|nan)f?)  # number
        |([{}<>:\[\],;]))""",
    re.X,
)

_ESC = {"n": "\n", "t": "\t", "r": "\r", "\\": "\\", "'": "'", '"': '"', "0": "\0"}


def _unescape(s: str) -> str:
    out, i = [], 0
    while i < len(s):
        c = s[i]
        if c == "\\" and i + 1 < len(s):
            nxt = s[i + 1]
            if nxt in _ESC:
                out.append(_ESC[nxt]); i += 2; continue
            if nxt == "x":
                out.append(chr(int(s[i + 2:i + 4], 16))); i += 4; continue
            if nxt.isdigit():
                j = i + 1
                while j < len(s) and j < i + 4 and s[j].isdigit():
                    j += 1
                out.append(chr(int(s[i + 1:j], 8))); i = j; continue
        out.append(c); i += 1
    return "".join(out)


def _tokenize(text: str):
    pos, n = 0, len(text)
    while True:
        m = _TOKEN.match(text, pos)
        if m is None:
            if text[pos:].strip() == "":
                return
            line = text.count("\n", 0, pos) + 1
            raise ValueError("prototxt parse error at line %d: %r" % (line, text[pos:pos + 30]))
        pos = m.end()
        if m.group(1) is not None:
            continue
        if m.group(2) is not None:
            yield ("id", m.group(2))
        elif m.group(3) is not None:
            yield ("str", _unescape(m.group(3)[1:-1]))
        elif m.group(4) is not None:
            yield ("num", m.group(4))
        else:
            yield ("sym", m.group(5))
        if pos >= n:
            return


def _scalar_from_token(typ: str, kind: str, tok: str, field: str):
    if typ == "string":
        if kind != "str":
            raise ValueError("field %s expects a string, got %r" % (field, tok))
        return tok
    if typ == "bool":
        if tok in ("true", "True", "1", "t"):
            return True
        if tok in ("false", "False", "0", "f"):
            return False
        raise ValueError("field %s expects a bool, got %r" % (field, tok))
    if typ.startswith("enum:"):
        table = ENUMS[typ[5:]]
        if kind == "id":
            if tok not in table:
                raise ValueError("field %s: unknown enum value %r" % (field, tok))
            return table[tok]
        return int(tok)
    if typ in ("float", "double"):
        t = tok[:-1] if tok.endswith("f") and not tok.endswith("inf") else tok
        return float(t)
    if typ in ("int32", "int64", "uint32", "uint64"):
        v = int(tok, 0) if not re.search(r"[.eE]", tok) else int(float(tok))
        if typ.startswith("u") and v < 0:
            raise ValueError("field %s is unsigned, got %r" % (field, tok))
        return v
    raise ValueError("unsupported scalar type %s" % typ)


def parse_text(text: str, type_name: str = "NetParameter", into: Msg | None = None) -> Msg:
    """protobuf text format -> Msg (``google.protobuf.text_format.Merge`` semantics: repeated
    fields append, optional scalars overwrite, sub-messages merge)."""
    toks = list(_tokenize(text))
    pos = 0

    def parse_message(msg: Msg, closer):
        nonlocal pos
        sch = SCHEMA[msg.type_name]
        while pos < len(toks):
            kind, tok = toks[pos]
            if kind == "sym" and tok == closer:
                pos += 1
                return
            if kind == "sym" and tok in (",", ";"):
                pos += 1
                continue
            if kind != "id":
                raise ValueError("expected field name in %s, got %r" % (msg.type_name, tok))
            pos += 1
            name = tok
            if name not in sch:
                raise ValueError("message %s has no field %r (not in the restated schema subset)"
                                 % (msg.type_name, name))
            _, typ, lab = sch[name]
            if pos < len(toks) and toks[pos] == ("sym", ":"):
                pos += 1
            if typ in SCHEMA:
                k2, t2 = toks[pos]
                if k2 != "sym" or t2 not in ("{", "<"):
                    raise ValueError("field %s expects a message body" % name)
                pos += 1
                if lab == "o":
                    sub = msg._f.get(name)
                    if not isinstance(sub, Msg):
                        sub = Msg(typ)
                        msg._f[name] = sub
                else:
                    sub = Msg(typ)
                    getattr(msg, name).append(sub)
                parse_message(sub, "}" if t2 == "{" else ">")
                continue
            # scalar(s); "[a, b]" list syntax allowed for repeated
            values = []
            k2, t2 = toks[pos]
            if k2 == "sym" and t2 == "[":
                pos += 1
                while toks[pos] != ("sym", "]"):
                    k3, t3 = toks[pos]
                    pos += 1
                    if (k3, t3) == ("sym", ","):
                        continue
                    values.append(_scalar_from_token(typ, k3, t3, name))
                pos += 1
            else:
                pos += 1
                v = _scalar_from_token(typ, k2, t2, name)
                if typ == "string":                      # adjacent string literals concatenate
                    while pos < len(toks) and toks[pos][0] == "str":
                        v += toks[pos][1]
                        pos += 1
                values.append(v)
            if lab == "o":
                msg._f[name] = values[-1]
            else:
                getattr(msg, name).extend(values)
        if closer is not None:
            raise ValueError("unterminated message %s" % msg.type_name)

    root = into if into is not None else Msg(type_name)
    parse_message(root, None)
    return root


def _fmt_float(v: float) -> str:
    f32 = np.float32(v)
    if float(f32) == int(f32) and abs(float(f32)) < 1e15:
        return str(int(f32))
    # shortest repr that round-trips through float32 (what protobuf's text printer emits)
    return np.format_float_positional(f32, unique=True, trim="-") if 1e-4 <= abs(float(f32)) < 1e16 \
        else np.format_float_scientific(f32, unique=True, trim="-")


_CESC = {"\\": "\\\\", "'": "\\'", '"': '\\"', "\n": "\\n", "\r": "\\r", "\t": "\\t"}


def _fmt_scalar(typ: str, v) -> str:
    if typ == "string":
        # protobuf's text printer (text_encoding.CEscape): \\ \' \" \n \r \t, other non-printable bytes as 3-digit octal
        out = []
        for ch in v:
            o = ord(ch)
            if ch in _CESC:
                out.append(_CESC[ch])
            elif o < 0x20 or o == 0x7f:
                out.append("\\%03o" % o)
            else:
                out.append(ch)
        return '"%s"' % "".join(out)
    if typ == "bool":
        return "true" if v else "false"
    if typ.startswith("enum:"):
        for k, n in ENUMS[typ[5:]].items():
            if n == v:
                return k
        return str(int(v))
    if typ == "float":
        return _fmt_float(v)
    if typ == "double":
        return repr(float(v))
    return str(int(v))


def format_text(msg: Msg, indent: int = 0) -> str:
    """Msg -> protobuf text format (field-number order, like ``str(pb)``)."""
    sch = SCHEMA[msg.type_name]
    pad = "  " * indent
    out: List[str] = []
    for name, val in msg.fields():
        _, typ, lab = sch[name]
        vals = val if lab in ("r", "p") else [val]
        for v in vals:
            if typ in SCHEMA:
                out.append("%s%s {\n%s%s}\n" % (pad, name, format_text(v, indent + 1), pad))
            else:
                out.append("%s%s: %s\n" % (pad, name, _fmt_scalar(typ, v)))
    return "".join(out)


# --------------------------------------------------------------------------------------
# Binary wire format
# --------------------------------------------------------------------------------------
def _read_varint(buf, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _write_varint(out: bytearray, v: int) -> None:
    if v < 0:
        v += 1 << 64
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return


def _to_signed(v: int, bits: int) -> int:
    v &= (1 << 64) - 1
    if v >= 1 << 63:
        v -= 1 << 64
    if bits == 32:
        v = ((v + (1 << 31)) % (1 << 32)) - (1 << 31)
    return v


_WT = {"int32": 0, "int64": 0, "uint32": 0, "uint64": 0, "bool": 0, "float": 5, "double": 1, "string": 2}


def decode(buf, type_name: str = "NetParameter") -> Msg:
    """Binary wire format -> Msg.  Unknown fields are skipped (a real VGG ``.caffemodel`` carries
    layers/params outside the restated subset; ``Net::CopyTrainedLayersFrom`` ignores them too,
    ``net.cpp:741-750``)."""
    mv = memoryview(buf) if not isinstance(buf, memoryview) else buf
    return _decode(mv, 0, len(mv), type_name)


def _decode(mv, pos: int, end: int, type_name: str) -> Msg:
    msg = Msg(type_name)
    table = _BY_NUMBER[type_name]
    while pos < end:
        key, pos = _read_varint(mv, pos)
        num, wt = key >> 3, key & 7
        ent = table.get(num)
        if ent is None:                                   # skip unknown
            if wt == 0:
                _, pos = _read_varint(mv, pos)
            elif wt == 1:
                pos += 8
            elif wt == 2:
                ln, pos = _read_varint(mv, pos)
                pos += ln
            elif wt == 5:
                pos += 4
            else:
                raise ValueError("unsupported wire type %d" % wt)
            continue
        name, typ, lab = ent
        if typ in SCHEMA:
            ln, pos = _read_varint(mv, pos)
            sub = _decode(mv, pos, pos + ln, typ)
            pos += ln
            if lab == "o":
                msg._f[name] = sub
            else:
                getattr(msg, name).append(sub)
            continue
        base = "int32" if typ.startswith("enum:") else typ
        if wt == 2 and base != "string":                  # packed repeated scalars
            ln, pos = _read_varint(mv, pos)
            chunk = mv[pos:pos + ln]
            pos += ln
            if base == "float":
                arr = np.frombuffer(chunk, dtype="<f4")
            elif base == "double":
                arr = np.frombuffer(chunk, dtype="<f8")
            else:
                vals, p2 = [], 0
                while p2 < ln:
                    v, p2 = _read_varint(chunk, p2)
                    vals.append(_to_signed(v, 32 if base == "int32" else 64) if base.startswith("int")
                                else (bool(v) if base == "bool" else v))
                arr = vals
            cur = msg._f.get(name)
            if isinstance(arr, np.ndarray):
                msg._f[name] = arr if cur is None or len(cur) == 0 else np.concatenate([np.asarray(cur, arr.dtype), arr])
            else:
                getattr(msg, name).extend(arr)
            continue
        if wt == 0:
            v, pos = _read_varint(mv, pos)
            if base == "bool":
                v = bool(v)
            elif base in ("int32", "int64"):
                v = _to_signed(v, 32 if base == "int32" else 64)
        elif wt == 5:
            v = struct.unpack_from("<f", mv, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<d", mv, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(mv, pos)
            v = bytes(mv[pos:pos + ln]).decode("utf-8")
            pos += ln
        else:
            raise ValueError("unsupported wire type %d" % wt)
        if lab == "o":
            msg._f[name] = v
        else:
            cur = getattr(msg, name)
            if isinstance(cur, np.ndarray):
                msg._f[name] = np.append(cur, v)
            else:
                cur.append(v)
    return msg


def encode(msg: Msg) -> bytes:
    out = bytearray()
    _encode(msg, out)
    return bytes(out)


def _encode(msg: Msg, out: bytearray) -> None:
    sch = SCHEMA[msg.type_name]
    for name, val in msg.fields():
        num, typ, lab = sch[name]
        if typ in SCHEMA:
            for sub in (val if lab != "o" else [val]):
                body = bytearray()
                _encode(sub, body)
                _write_varint(out, (num << 3) | 2)
                _write_varint(out, len(body))
                out += body
            continue
        base = "int32" if typ.startswith("enum:") else typ
        if lab == "p":
            if base == "float":
                body = np.ascontiguousarray(val, dtype="<f4").tobytes()
            elif base == "double":
                body = np.ascontiguousarray(val, dtype="<f8").tobytes()
            else:
                b2 = bytearray()
                for v in val:
                    _write_varint(b2, int(v))
                body = bytes(b2)
            _write_varint(out, (num << 3) | 2)
            _write_varint(out, len(body))
            out += body
            continue
        for v in (val if lab == "r" else [val]):
            wt = _WT[base]
            _write_varint(out, (num << 3) | wt)
            if wt == 0:
                _write_varint(out, int(v))
            elif wt == 5:
                out += struct.pack("<f", float(v))
            elif wt == 1:
                out += struct.pack("<d", float(v))
            else:
                b = v.encode("utf-8") if isinstance(v, str) else bytes(v)
                _write_varint(out, len(b))
                out += b


# --------------------------------------------------------------------------------------
# File helpers (error text mirrors caffe/python/caffe/_caffe.cpp:77-84 CheckFile)
# --------------------------------------------------------------------------------------
def _check_file(path: str) -> None:
    try:
        with open(path, "rb"):
            pass
    except OSError:
        raise RuntimeError("Could not open file " + str(path))


def read_net_text(path: str) -> Msg:
    _check_file(path)
    with open(path, "r") as f:
        return parse_text(f.read(), "NetParameter")


def read_net_binary(path: str) -> Msg:
    _check_file(path)
    with open(path, "rb") as f:
        data = f.read()
    if len(data) >= (1 << 31):                            # io.cpp:22,57 kProtoReadBytesLimit = INT_MAX
        raise RuntimeError("caffemodel larger than the 2 GB protobuf read limit")
    return decode(data, "NetParameter")


def write_net_binary(path: str, net: Msg) -> None:
    with open(path, "wb") as f:
        f.write(encode(net))


def blob_from_array(arr: np.ndarray) -> Msg:
    """ndarray -> BlobProto with ``shape`` + packed float ``data`` (``blob.cpp:522-546`` ToProto)."""
    arr = np.ascontiguousarray(arr, dtype=np.float32)
    return Msg("BlobProto", shape=Msg("BlobShape", dim=[int(d) for d in arr.shape]), data=arr.ravel())


def array_from_blob(bp: Msg) -> np.ndarray:
    """BlobProto -> float32 ndarray; honours legacy num/channels/height/width
    (``blob.cpp:472-506`` FromProto)."""
    if bp.has("shape"):
        shape = tuple(int(d) for d in bp.shape.dim)
    elif bp.has("num") or bp.has("channels") or bp.has("height") or bp.has("width"):
        shape = (bp.num, bp.channels, bp.height, bp.width)
    else:
        shape = (len(bp.data),)
    if bp.has("double_data"):
        data = np.asarray(bp.double_data, dtype=np.float64).astype(np.float32)
    else:
        data = np.asarray(bp.data, dtype=np.float32)
    if data.size != int(np.prod(shape)):
        raise RuntimeError("BlobProto data size %d does not match shape %s" % (data.size, shape))
    return data.reshape(shape)
