"""Synthetic deployment artefacts: the reference ships no weights (``README.md:3`` links to a
SharePoint file) and no network is available, so benchmarks and parity tests run on random-init
weights of the exact architecture, written as a real binary ``NetParameter`` so the loader path
(``Net::CopyTrainedLayersFrom``, ``net.cpp:733-785``) is exercised end to end.

Initialisation (seed ``RNG_SEED`` = 3, ``configs/default.toml:7``):
  * VGG16 + fusion + head 3x3/1x1 convs: He-normal ``std = sqrt(2 / (Cin*k*k))`` so activations keep
    their scale through 17-20 ReLU layers (the templates' own ``gaussian std=0.01`` fillers would
    shrink them to denormals by the heads);  biases ``N(0, 0.05)``.
  * ``conv5_256_up``: the bilinear kernel the template's filler asks for (``filler.hpp:244-262``).
  * ``cls_score*``: Gaussian scaled so the (fg-bg) logit has sigma ~ 1.5 at the reference input
    statistics, with a background bias of +7.3 -- about 2-3 % of anchors clear the 0.05 detection
    threshold, which gives box voting / NMS a realistic few thousand candidates per image.
  * ``bbox_pred*``: Gaussian scaled for delta sigma ~ 0.25.
The two scale constants were calibrated once with tools/calibrate_synthetic.py on the 224x224
parity image and are frozen here so every run (and the committed golden vectors) sees the same file.
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np

from . import caffe_proto as cp
from .graph import NetSpec, TEST
from .models import build_test_net, splice_dim_red

RNG_SEED = 3
CLS_LOGIT_GAIN = 0.019      # multiplies the He-normal cls weights (calibrated, see module docstring)
CLS_LOGIT_GAIN_STD = 0.034
BBOX_DELTA_GAIN = 0.00225
CLS_BG_BIAS = 7.3


def synthetic_image(seed: int, hw=(1024, 1024)) -> np.ndarray:
    """uint8 HxWx3 test image with structure at every scale (sum of 8 octaves of box-upsampled uniform
    noise, equal weights): unlike white noise it keeps its contrast when the pyramid resizes it by
    0.1x..1.4x, so every level produces detections, as natural images do."""
    rng = np.random.RandomState(seed)
    h, w = hw
    acc = np.zeros((h, w, 3), dtype=np.float64)
    for octave in range(8):
        s = 1 << octave
        small = rng.rand(-(-h // s), -(-w // s), 3)
        acc += np.repeat(np.repeat(small, s, axis=0), s, axis=1)[:h, :w]
    acc -= acc.min()
    acc /= max(acc.max(), 1e-9)
    return np.round(acc * 255.0).astype(np.uint8)


def bilinear_kernel(shape) -> np.ndarray:
    """``filler.hpp:244-262``: separable tent, f = ceil(k/2), c = (2f-1-f%2)/(2f)."""
    k = shape[3]
    f = int(np.ceil(k / 2.0))
    c = (2 * f - 1 - f % 2) / (2.0 * f)
    tent = np.array([1 - abs(x / f - c) for x in range(k)], dtype=np.float64)
    kern = np.outer(tent, tent).astype(np.float32)
    return np.broadcast_to(kern, shape).copy()


def synthetic_params(spec: NetSpec, seed: int = RNG_SEED) -> Dict[str, np.ndarray]:
    spec.infer_shapes({})
    rng = np.random.RandomState(seed)
    out: Dict[str, np.ndarray] = {}
    for l in spec.layers:
        if l.type == "BatchNorm":                    # running mean / variance / moving-average factor (batch_norm_layer.cpp:23-45)
            c = spec.param_shapes[l.param_keys[0]][0]
            factor = np.float32(0.999 * 50)           # the blobs hold sums scaled by this factor
            out[l.param_keys[0]] = (rng.standard_normal(c) * 0.1).astype(np.float32) * factor
            out[l.param_keys[1]] = (0.5 + rng.rand(c)).astype(np.float32) * factor
            out[l.param_keys[2]] = np.array([factor], dtype=np.float32)
            continue
        if l.type == "Scale":
            c = spec.param_shapes[l.param_keys[0]][0]
            # the last Scale of a residual branch is damped (as zero-gamma initialisation does) so that activations neither
            # grow nor vanish through a stack of residual adds
            gain = 0.5 if l.name.endswith("_branch2c") else 1.0
            out[l.param_keys[0]] = (gain * (0.8 + 0.4 * rng.rand(c))).astype(np.float32)
            if len(l.param_keys) > 1:
                out[l.param_keys[1]] = (rng.standard_normal(c) * 0.05).astype(np.float32)
            continue
        if l.type not in ("Convolution", "Deconvolution"):
            continue
        for i, key in enumerate(l.param_keys):
            if key in out:                        # shared (head_w / head_b): owner already filled
                continue
            shp = spec.param_shapes[key]
            if l.type == "Deconvolution":
                out[key] = bilinear_kernel(shp)
                continue
            if i == 0:
                fan_in = shp[1] * shp[2] * shp[3]
                w = rng.standard_normal(shp).astype(np.float32) * np.float32(np.sqrt(2.0 / fan_in))
                if l.name == "cls_score":          # standard template: one 6-channel cls conv
                    w *= np.float32(CLS_LOGIT_GAIN_STD)
                elif l.name.startswith("cls_score"):
                    w *= np.float32(CLS_LOGIT_GAIN)
                elif l.name.startswith("bbox_pred"):
                    w *= np.float32(BBOX_DELTA_GAIN)
                out[key] = w
            else:
                b = (rng.standard_normal(shp) * 0.05).astype(np.float32)
                if l.name.startswith("cls_score"):
                    # channel layout: first half background, second half foreground
                    # (test_template.prototxt:528-549 reshape -> softmax over (bg, fg))
                    half = shp[0] // 2
                    b[:half] += np.float32(CLS_BG_BIAS)
                elif l.name.startswith("bbox_pred"):
                    b *= np.float32(0.2)
                out[key] = b
    return out


def params_to_netparameter(spec: NetSpec, params: Dict[str, np.ndarray]) -> cp.Msg:
    """One LayerParameter per parametrised layer with its blobs (what ``Net::ToProto`` saves).
    Layers that share a param each carry a copy, as Caffe snapshots do."""
    net = cp.Msg("NetParameter", name=spec.name or "face")
    for l in spec.layers:
        if not l.param_keys:
            continue
        lay = cp.Msg("LayerParameter", name=l.name, type=l.type)
        lay.blobs = [cp.blob_from_array(params[k]) for k in l.param_keys]
        net.layer.append(lay)
    return net


def write_synthetic_deployment(out_dir: str, dilation: bool = True, seed: int = RNG_SEED, input_hw=(224, 224)):
    """Write ``test.prototxt`` (+ dim_red splice when ``dilation``) and ``synthetic.caffemodel`` into
    ``out_dir``; returns (prototxt_path, caffemodel_path).  Idempotent."""
    os.makedirs(out_dir, exist_ok=True)
    tag = "dil" if dilation else "std"
    proto = os.path.join(out_dir, "test_%s.prototxt" % tag)
    model = os.path.join(out_dir, "synthetic_%s_seed%d.caffemodel" % (tag, seed))
    net = build_test_net(dilation, input_hw)
    if dilation:
        net = splice_dim_red(net)
    tmp = ".tmp%d" % os.getpid()                  # several ranks may race here: each writes its own temp, rename is atomic
    if not os.path.exists(proto):
        with open(proto + tmp, "w") as f:
            f.write(cp.format_text(net))
        os.replace(proto + tmp, proto)
    if not os.path.exists(model):
        spec = NetSpec(net, TEST)
        params = synthetic_params(spec, seed)
        cp.write_net_binary(model + tmp, params_to_netparameter(spec, params))
        os.replace(model + tmp, model)
    return proto, model


def write_synthetic_resnet_deployment(out_dir: str, blocks=(3, 4), seed: int = RNG_SEED, input_hw=(224, 224),
                                      stride_on_3x3: bool = False):
    """``models.build_resnet_test_net`` + synthetic weights (He-normal convs, plausible BatchNorm statistics, Scale gains)
    as ``test_resnet.prototxt`` / ``.caffemodel``; returns the two paths."""
    from .models import build_resnet_test_net
    os.makedirs(out_dir, exist_ok=True)
    tag = "resnet_" + "_".join(str(b) for b in blocks) + ("_s3x3" if stride_on_3x3 else "")
    proto = os.path.join(out_dir, "test_%s.prototxt" % tag)
    model = os.path.join(out_dir, "synthetic_%s_seed%d.caffemodel" % (tag, seed))
    net = build_resnet_test_net(blocks, input_hw, stride_on_3x3=stride_on_3x3)
    tmp = ".tmp%d" % os.getpid()
    if not os.path.exists(proto):
        with open(proto + tmp, "w") as f:
            f.write(cp.format_text(net))
        os.replace(proto + tmp, proto)
    if not os.path.exists(model):
        spec = NetSpec(net, TEST)
        cp.write_net_binary(model + tmp, params_to_netparameter(spec, synthetic_params(spec, seed)))
        os.replace(model + tmp, model)
    return proto, model
