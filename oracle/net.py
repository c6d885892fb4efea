"""ORACLE (test infrastructure, not product code) -- runs a deploy net layer by layer on the CPU
the way ``Net::ForwardFromTo`` does (``caffe/src/caffe/net.cpp:516-532``): one fp32 NCHW blob per
top, every layer a separate pass (conv, then in-place ReLU, then pool ...), the Python
ProposalLayer evaluated by oracle/proposal.py.

The graph itself (prototxt parsing, in-place tops, shared params, weight loading by layer name)
comes from smallhardface_b200.graph / caffe_proto, which are format code shared with the product;
all arithmetic is in oracle/layers.py.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np

from smallhardface_b200 import caffe_proto as cp
from smallhardface_b200.graph import NetSpec, TEST, load_weights, parse_python_param_str, reshape_shape

from . import layers as L
from .proposal import proposal_forward

F32 = np.float32


class OracleNet:
    def __init__(self, net_param, model_param=None, engine="sgemm", fast=False,
                 pre_nms_topn=10000, score_thresh=0.002, min_size=0):
        """``engine``: 'sgemm' = the reference algorithm (im2col + BLAS), 'torch' = oneDNN conv.
        ``fast`` swaps the python-loop pool/deconv for their vectorised equals (full-size runs)."""
        if isinstance(net_param, str):
            net_param = cp.read_net_text(net_param)
        if isinstance(model_param, str):
            model_param = cp.read_net_binary(model_param)
        self.spec = NetSpec(net_param, TEST)
        self.params = load_weights(self.spec, model_param if model_param is not None else cp.Msg("NetParameter"))
        self.engine = engine
        self.fast = fast
        self.cfg = dict(pre_nms_topn=pre_nms_topn, score_thresh=score_thresh, min_size=min_size)
        self.blobs: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def forward(self, **inputs):
        spec = self.spec
        if set(inputs) != set(spec.inputs):
            raise Exception("Input blob arguments do not match net inputs.")
        b = self.blobs = OrderedDict()
        for spec_l in spec.layers:
            t = spec_l.type
            if t == "Input":
                for nm in spec_l.tops:
                    b[nm] = np.ascontiguousarray(inputs[nm], dtype=F32)
                continue
            xs = [b[n] for n in spec_l.bottoms]
            p = spec_l.p
            if t == "Convolution":
                w = self.params[spec_l.param_keys[0]]
                bias = self.params[spec_l.param_keys[1]] if p["bias_term"] else None
                y = L.conv(xs[0], w, bias, (p["ph"], p["pw"]), (p["sh"], p["sw"]), (p["dh"], p["dw"]),
                           p["group"], engine=self.engine)
            elif t == "Deconvolution":
                w = self.params[spec_l.param_keys[0]]
                bias = self.params[spec_l.param_keys[1]] if p["bias_term"] else None
                if self.fast and p["group"] == xs[0].shape[1] == p["num_output"] and bias is None:
                    y = L.deconv_depthwise_fast(xs[0], w, (p["ph"], p["pw"]), (p["sh"], p["sw"]))
                else:
                    y = L.deconv(xs[0], w, bias, (p["ph"], p["pw"]), (p["sh"], p["sw"]), (p["dh"], p["dw"]), p["group"])
            elif t == "ReLU":
                y = L.relu(xs[0], p["negative_slope"])
            elif t == "BatchNorm":
                if not p["use_global_stats"]:
                    raise ValueError("BatchNorm %s: batch statistics are a TRAIN-phase mode" % spec_l.name)
                mean, var, factor = (self.params[k] for k in spec_l.param_keys)
                y = L.batch_norm(xs[0], mean, var, factor, p["eps"])
            elif t == "Scale":
                gamma = self.params[spec_l.param_keys[0]]
                beta = self.params[spec_l.param_keys[1]] if p["bias_term"] else None
                y = L.scale(xs[0], gamma, beta)
            elif t == "Pooling":
                if p["pool"] != 0:
                    raise ValueError("only MAX pooling is on the hot path")
                h, w_ = xs[0].shape[2:]
                if self.fast and (p["kh"], p["kw"], p["sh"], p["sw"], p["ph"], p["pw"]) == (2, 2, 2, 2, 0, 0) \
                        and h % 2 == 0 and w_ % 2 == 0:
                    y = L.max_pool_2x2_fast(xs[0])
                else:
                    y = L.max_pool(xs[0], (p["kh"], p["kw"]), (p["sh"], p["sw"]), (p["ph"], p["pw"]))
            elif t == "Eltwise":
                if p["operation"] != 1:
                    raise ValueError("only Eltwise SUM is on the hot path")
                y = L.eltwise_sum(xs, p["coeff"])
            elif t == "Concat":
                y = L.concat(xs, p["axis"])
            elif t == "Reshape":
                y = xs[0].reshape(reshape_shape(xs[0].shape, p["dims"], p["axis"], p["num_axes"], spec_l.name))
            elif t == "Softmax":
                y = L.softmax(xs[0], p["axis"])
            elif t == "Split":
                for nm in spec_l.tops:
                    b[nm] = xs[0]
                continue
            elif t == "Python":
                if (p["module"], p["layer"]) != ("lib.layers.proposal_layer", "ProposalLayer"):
                    raise ValueError("oracle only knows the ProposalLayer python layer")
                lp = parse_python_param_str(p["param_str"])
                boxes, probs, order = proposal_forward(
                    xs[-3], xs[-2], xs[-1], feat_stride=tuple(lp["feat_stride"]),
                    scales=tuple(lp.get("scales", (8, 16, 32))), ratios=tuple(lp.get("ratios", (0.5, 1, 2))),
                    base_size=lp.get("base_size", 16), shifts=tuple(lp.get("shifts", [0])), **self.cfg)
                b[spec_l.tops[0]] = boxes
                if len(spec_l.tops) > 1:
                    b[spec_l.tops[1]] = probs
                self.last_order = order
                continue
            else:                                                  # pragma: no cover
                raise AssertionError(t)
            b[spec_l.tops[0]] = y
        return {o: b[o] for o in spec.outputs}
