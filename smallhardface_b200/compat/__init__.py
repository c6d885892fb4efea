"""Drop-in modules under the reference's own import names.

``install()`` makes ``import caffe``, ``import nms.nms_wrapper`` / ``nms.cpu_nms`` / ``nms.gpu_nms`` and
``import utils.cython_bbox`` resolve to this package (the reference builds those as a Boost.Python module and
three Cython extensions: ``caffe/python/caffe/_caffe.cpp``, ``lib/setup.py:112-143``).  With
``reference_root`` it also puts the reference's ``lib/`` on ``sys.path`` behind a py2->py3 source hook
(smallhardface_b200.compat.py2hook) so its unmodified Python files import on this interpreter.
"""
from __future__ import annotations

import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install(reference_root: str | None = None):
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)                       # `caffe`, `nms` packages live here
    for stale in ("caffe", "nms"):
        mod = sys.modules.get(stale)
        if mod is not None and not getattr(mod, "__file__", "").startswith(_HERE):
            for k in [k for k in sys.modules if k == stale or k.startswith(stale + ".")]:
                del sys.modules[k]
    importlib.import_module("caffe")
    importlib.import_module("nms")
    from . import cython_bbox
    sys.modules["utils.cython_bbox"] = cython_bbox
    if reference_root:
        from . import py2hook
        py2hook.install(reference_root)
        utils = importlib.import_module("utils")         # the reference's lib/utils package
        utils.cython_bbox = cython_bbox
    return True
