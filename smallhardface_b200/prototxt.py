"""Deploy-prototxt rewriting of the test path in Python 3 -- ``lib/prototxt/manipulate.py:64-86,166-188`` without
``caffe_pb2`` / the py2 hook.

``manipulate_test`` does what the reference's function does before every test run: pick the dilation template when
``MODEL.DIFFERENT_DILATION.ENABLE`` is set (``manipulate.py:66-67``: the ``ori`` argument is then IGNORED), parse it,
splice the ``conv4_fuse_final_dim_red`` 3x3 conv + ReLU in front of the heads (``_add_dimension_reduction``; a no-op
unless the dilation switch is on, ``:167-168``) and write the result as text.  The reference also renders a .jpg of the
graph for TensorBoard (``:70-75``); graphviz is not part of this path and is skipped.

The text is produced by this package's own codec (``caffe_proto.format_text``), which prints what protobuf's
``str(message)`` prints for these messages (tests/test_config_native.py compares against the file the reference's own
function wrote through the real protobuf runtime).
"""
from __future__ import annotations

import os

from . import caffe_proto as cp
from .models import splice_dim_red

DILATION_TEMPLATE = "models/test_different_dilation_template.prototxt"       # manipulate.py:67


def add_dimension_reduction(net: cp.Msg, cfg) -> cp.Msg:
    """``manipulate.py:166-188``."""
    if not cfg.MODEL.DIFFERENT_DILATION.ENABLE:
        return net
    return splice_dim_red(net)


def manipulate_test(cfg, ori: str, target_test: str, root_dir: str | None = None) -> cp.Msg:
    """Returns the rewritten NetParameter and writes it to ``target_test``.  Relative template paths are resolved
    against ``root_dir`` (default ``cfg.ROOT_DIR``) -- the reference resolves them against the current directory."""
    root = root_dir or cfg.ROOT_DIR
    if cfg.MODEL.DIFFERENT_DILATION.ENABLE:
        ori = DILATION_TEMPLATE
    path = ori if os.path.isabs(ori) else os.path.join(root, ori)
    net = add_dimension_reduction(cp.read_net_text(path), cfg)
    with open(target_test, "w") as f:
        f.write(cp.format_text(net))
    return net
