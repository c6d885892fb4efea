"""Programmatic builders for the two deploy nets the reference ships as prototxt templates.

``build_test_net(dilation=False)`` reproduces ``models/test_template.prototxt`` (VGG16 conv1_1..conv5_3,
the conv5/conv4 fusion, one 3x3 ``head`` and 1x1 ``cls_score``/``bbox_pred``), and
``build_test_net(dilation=True)`` reproduces ``models/test_different_dilation_template.prototxt``
(three weight-sharing heads with dilation 1/2/4) -- layer for layer, name for name, so that a
``.caffemodel`` written for one loads into the other.  tests/test_models_match_reference.py parses
the reference's own files (when /root/reference is mounted) and checks equality message by message.

``splice_dim_red`` restates ``lib/prototxt/manipulate.py:166-188`` (the runtime insertion of
``conv4_fuse_final_dim_red`` that every shipped dilation config performs).
"""
from __future__ import annotations

from typing import List

from .caffe_proto import Msg

PROPOSAL_PARAM_STR = "{'feat_stride': [8,8,8],'scales': [1,2,4], 'ratios':[1,]}"   # test_template.prototxt:561


def _param(lr=None, decay=None, name=None) -> Msg:
    m = Msg("ParamSpec")
    if name is not None:
        m.name = name
    if lr is not None:
        m.lr_mult = lr
    if decay is not None:
        m.decay_mult = decay
    return m


def _filler(kind, **kw) -> Msg:
    return Msg("FillerParameter", type=kind, **kw)


def conv_layer(name, bottom, top, num_output, kernel, pad, params, stride=None, dilation=None,
               fillers=True) -> Msg:
    c = Msg("ConvolutionParameter", num_output=num_output)
    c.pad = [pad]
    c.kernel_size = [kernel]
    if stride is not None:
        c.stride = [stride]
    if dilation is not None:
        c.dilation = [dilation]
    if fillers:
        c.weight_filler = _filler("gaussian", std=0.01)
        c.bias_filler = _filler("constant", value=0)
    return Msg("LayerParameter", name=name, type="Convolution", bottom=[bottom], top=[top],
               param=params, convolution_param=c)


def relu_layer(name, blob) -> Msg:
    return Msg("LayerParameter", name=name, type="ReLU", bottom=[blob], top=[blob])


def pool_layer(name, bottom, top) -> Msg:
    return Msg("LayerParameter", name=name, type="Pooling", bottom=[bottom], top=[top],
               pooling_param=Msg("PoolingParameter", pool=0, kernel_size=2, stride=2))


VGG_CFG = [(1, 2, 64), (2, 2, 128), (3, 3, 256), (4, 3, 512), (5, 3, 512)]   # test_template.prototxt:17-367


def _backbone() -> List[Msg]:
    layers, bottom = [], "data"
    for stage, reps, ch in VGG_CFG:
        for r in range(1, reps + 1):
            nm = "conv%d_%d" % (stage, r)
            # conv1_x / conv2_x are frozen (lr 0), conv3_x.. train at lr 1 / 2 -- training metadata,
            # kept so the generated NetParameter equals the template message for message
            ps = [_param(0, 0), _param(0, 0)] if stage <= 2 else [_param(1), _param(2)]
            layers.append(conv_layer(nm, bottom, nm, ch, 3, 1, ps, fillers=False))
            layers.append(relu_layer("relu%d_%d" % (stage, r), nm))
            bottom = nm
        if stage < 5:
            layers.append(pool_layer("pool%d" % stage, bottom, "pool%d" % stage))
            bottom = "pool%d" % stage
    return layers


def _fusion() -> List[Msg]:
    """conv5_256 (+ReLU) -> 2x depthwise bilinear deconv; conv4_256 (+ReLU); concat; 3x3 fuse
    (``test_template.prototxt:369-478``)."""
    p12 = lambda: [_param(1), _param(2)]
    up = Msg("ConvolutionParameter", num_output=256, bias_term=False, group=256)
    up.kernel_size = [4]; up.stride = [2]; up.pad = [1]
    up.weight_filler = _filler("bilinear")
    return [
        conv_layer("conv5_256", "conv5_3", "conv5_256", 256, 1, 0, p12()),
        relu_layer("conv5_256_relu", "conv5_256"),
        Msg("LayerParameter", name="conv5_256_up", type="Deconvolution", bottom=["conv5_256"],
            top=["conv5_256_up"], param=[_param(0, 0)], convolution_param=up),
        conv_layer("conv4_256", "conv4_3", "conv4_256", 256, 1, 0, p12()),
        relu_layer("conv4_256_relu", "conv4_256"),
        Msg("LayerParameter", name="conv4_fuse", type="Concat", bottom=["conv5_256_up", "conv4_256"],
            top=["conv4_fuse"], concat_param=Msg("ConcatParameter", axis=1)),
        conv_layer("conv4_fuse_final", "conv4_fuse", "conv4_fuse_final", 512, 3, 1, p12()),
        relu_layer("conv4_fuse_final_relu", "conv4_fuse_final"),
    ]


def _tail(cls_blob: str) -> List[Msg]:
    """Softmax over (bg,fg), reshape back to 6 channels, Python ProposalLayer
    (``test_template.prototxt:537-563``)."""
    shape6 = Msg("BlobShape", dim=[0, 6, -1, 0])
    return [
        Msg("LayerParameter", name="cls_prob", type="Softmax", bottom=[cls_blob], top=["cls_prob_output"]),
        Msg("LayerParameter", name="cls_prob_reshape", type="Reshape", bottom=["cls_prob_output"],
            top=["cls_prob_reshape_output"], reshape_param=Msg("ReshapeParameter", shape=shape6)),
        Msg("LayerParameter", name="proposal", type="Python",
            bottom=["cls_prob_reshape_output", "bbox_pred_output", "im_info"], top=["boxes", "cls_prob"],
            python_param=Msg("PythonParameter", module="lib.layers.proposal_layer", layer="ProposalLayer",
                             param_str=PROPOSAL_PARAM_STR)),
    ]


def build_test_net(dilation: bool = False, input_hw=(224, 224)) -> Msg:
    net = Msg("NetParameter", name="face")
    net.input = ["data", "im_info"]
    net.input_shape = [Msg("BlobShape", dim=[1, 3, int(input_hw[0]), int(input_hw[1])]),
                       Msg("BlobShape", dim=[1, 3])]
    layers = _backbone() + _fusion()
    hp = lambda: [_param(1.0, 1.0), _param(2.0, 0)]
    if not dilation:                                    # test_template.prototxt:480-534
        layers += [
            conv_layer("head", "conv4_fuse_final", "head", 128, 3, 1, hp(), stride=1),
            relu_layer("head_relu", "head"),
            conv_layer("cls_score", "head", "cls_score_output", 6, 1, 0, hp(), stride=1),
            conv_layer("bbox_pred", "head", "bbox_pred_output", 12, 1, 0, hp(), stride=1),
            Msg("LayerParameter", name="cls_reshape", type="Reshape", bottom=["cls_score_output"],
                top=["cls_score_reshape_output"],
                reshape_param=Msg("ReshapeParameter", shape=Msg("BlobShape", dim=[0, 2, -1, 0]))),
        ]
    else:                                               # test_different_dilation_template.prototxt:479-669
        for d in (1, 2, 4):
            shared = [_param(1.0, 1.0, name="head_w"), _param(2.0, 0, name="head_b")]
            layers.append(conv_layer("head_%d" % d, "conv4_fuse_final", "head_%d" % d, 128, 3, d, shared,
                                     stride=1, dilation=d))
            layers.append(relu_layer("head_%d_relu" % d, "head_%d" % d))
        for d in (1, 2, 4):
            layers.append(conv_layer("cls_score_%d" % d, "head_%d" % d, "cls_score_%d_output" % d, 2, 1, 0, hp(), stride=1))
            layers.append(conv_layer("bbox_pred_%d" % d, "head_%d" % d, "bbox_pred_%d_output" % d, 4, 1, 0, hp(), stride=1))
        layers += [
            Msg("LayerParameter", name="cls_score_output_concat", type="Concat",
                bottom=["cls_score_%d_output" % d for d in (1, 2, 4)], top=["cls_score_reshape_output"],
                concat_param=Msg("ConcatParameter", axis=2)),
            Msg("LayerParameter", name="bbox_pred_output_concat", type="Concat",
                bottom=["bbox_pred_%d_output" % d for d in (1, 2, 4)], top=["bbox_pred_output"],
                concat_param=Msg("ConcatParameter", axis=1)),
        ]
    layers += _tail("cls_score_reshape_output")
    net.layer = layers
    return net


def splice_dim_red(net: Msg) -> Msg:
    """``lib/prototxt/manipulate.py:166-188`` _add_dimension_reduction: rename the 512-ch
    ``conv4_fuse_final`` blob to ``..._tmp`` and insert a 3x3 512->128 conv + ReLU that takes over
    the name ``conv4_fuse_final`` right before the first ``head*`` layer."""
    out = net.copy()
    layers = list(out.layer)
    split = min(i for i, l in enumerate(layers) if l.name.startswith("head"))
    assert layers[split - 2].name == "conv4_fuse_final"
    layers[split - 2].top[0] += "_tmp"
    layers[split - 1].bottom[0] += "_tmp"
    layers[split - 1].top[0] += "_tmp"
    red = conv_layer("conv4_fuse_final_dim_red", "conv4_fuse_final_tmp", "conv4_fuse_final", 128, 3, 1,
                     [_param(1.0, 1.0), _param(2.0, 1.0)], dilation=1)
    red.convolution_param.bias_filler.value = 0.0
    out.layer = layers[:split] + [red, relu_layer("conv4_fuse_final_dim_red_relu", "conv4_fuse_final")] + layers[split:]
    return out


# ---------------------------------------------------------------------------------------------------------------
# ResNet-style backbone (BASELINE north_star names "the ResNet/VGG backbone"; the reference ships no ResNet prototxt,
# SURVEY F1).  Layer naming and structure follow the public Caffe ResNet-50 deploy file (conv1 7x7/2 + bn + scale + relu,
# pool1 MAX 3x3/2, bottleneck blocks res{2,3}{a,b,..}_branch{1,2a,2b,2c} with the stride on the FIRST 1x1 convolutions of
# a stage, in-place BatchNorm(use_global_stats) / Scale(bias_term) / ReLU, Eltwise SUM + ReLU), cut after res3 (stride 8:
# the stride of the reference's detection heads) and followed by the standard single-head tail of test_template.prototxt.
# ---------------------------------------------------------------------------------------------------------------
def _bn_scale(name: str, blob: str) -> List[Msg]:
    return [
        Msg("LayerParameter", name="bn" + name, type="BatchNorm", bottom=[blob], top=[blob],
            batch_norm_param=Msg("BatchNormParameter", use_global_stats=True)),
        Msg("LayerParameter", name="scale" + name, type="Scale", bottom=[blob], top=[blob],
            scale_param=Msg("ScaleParameter", bias_term=True)),
    ]


def _res_conv(name, bottom, num_output, kernel, pad, stride) -> Msg:
    c = Msg("ConvolutionParameter", num_output=num_output, bias_term=False)
    c.pad = [pad]
    c.kernel_size = [kernel]
    c.stride = [stride]
    return Msg("LayerParameter", name=name, type="Convolution", bottom=[bottom], top=[name], convolution_param=c)


def _bottleneck(stage: int, block: str, bottom: str, mid: int, out: int, stride: int, project: bool,
                stride_on_3x3: bool = False) -> List[Msg]:
    pre = "res%d%s" % (stage, block)
    tag = "%d%s" % (stage, block)
    layers: List[Msg] = []
    shortcut = bottom
    if project:
        layers += [_res_conv(pre + "_branch1", bottom, out, 1, 0, stride)] + _bn_scale(tag + "_branch1", pre + "_branch1")
        shortcut = pre + "_branch1"
    s_a, s_b = (1, stride) if stride_on_3x3 else (stride, 1)      # "v1.5" / torchvision put the stride on the 3x3 convolution
    layers += [_res_conv(pre + "_branch2a", bottom, mid, 1, 0, s_a)] + _bn_scale(tag + "_branch2a", pre + "_branch2a")
    layers.append(relu_layer(pre + "_branch2a_relu", pre + "_branch2a"))
    layers += [_res_conv(pre + "_branch2b", pre + "_branch2a", mid, 3, 1, s_b)] + _bn_scale(tag + "_branch2b", pre + "_branch2b")
    layers.append(relu_layer(pre + "_branch2b_relu", pre + "_branch2b"))
    layers += [_res_conv(pre + "_branch2c", pre + "_branch2b", out, 1, 0, 1)] + _bn_scale(tag + "_branch2c", pre + "_branch2c")
    layers.append(Msg("LayerParameter", name=pre, type="Eltwise", bottom=[shortcut, pre + "_branch2c"], top=[pre]))
    layers.append(relu_layer(pre + "_relu", pre))
    return layers


def build_resnet_test_net(blocks=(3, 4), input_hw=(224, 224), stride_on_3x3: bool = False) -> Msg:
    """ResNet-50 through res3 (``blocks`` = bottlenecks per stage) + the standard detection head.  ``stride_on_3x3``: the
    stage transition strides the 3x3 convolution (torchvision / "v1.5") instead of the first 1x1 (the original Caffe model)."""
    net = Msg("NetParameter", name="face_resnet")
    net.input = ["data", "im_info"]
    net.input_shape = [Msg("BlobShape", dim=[1, 3, int(input_hw[0]), int(input_hw[1])]), Msg("BlobShape", dim=[1, 3])]
    layers = [_res_conv("conv1", "data", 64, 7, 3, 2)] + _bn_scale("_conv1", "conv1") + [relu_layer("conv1_relu", "conv1")]
    layers.append(Msg("LayerParameter", name="pool1", type="Pooling", bottom=["conv1"], top=["pool1"],
                      pooling_param=Msg("PoolingParameter", pool=0, kernel_size=3, stride=2)))
    bottom = "pool1"
    for si, n_blocks in enumerate(blocks):
        stage = si + 2
        mid, out = 64 << si, 256 << si
        for bi in range(n_blocks):
            block = "abcdefgh"[bi]
            layers += _bottleneck(stage, block, bottom, mid, out, 2 if (bi == 0 and stage > 2) else 1, project=bi == 0,
                                  stride_on_3x3=stride_on_3x3)
            bottom = "res%d%s" % (stage, block)
    hp = lambda: [_param(1.0, 1.0), _param(2.0, 0)]
    layers += [
        conv_layer("head", bottom, "head", 128, 3, 1, hp(), stride=1),
        relu_layer("head_relu", "head"),
        conv_layer("cls_score", "head", "cls_score_output", 6, 1, 0, hp(), stride=1),
        conv_layer("bbox_pred", "head", "bbox_pred_output", 12, 1, 0, hp(), stride=1),
        Msg("LayerParameter", name="cls_reshape", type="Reshape", bottom=["cls_score_output"],
            top=["cls_score_reshape_output"],
            reshape_param=Msg("ReshapeParameter", shape=Msg("BlobShape", dim=[0, 2, -1, 0]))),
    ]
    layers += _tail("cls_score_reshape_output")
    net.layer = layers
    return net
