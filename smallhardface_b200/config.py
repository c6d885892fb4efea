"""Python 3 configuration plumbing of the test path -- ``lib/utils/get_config.py`` without the py2 import hook.

Same semantics as the reference module: ``configs/default.toml`` is the schema (every key that may be set must exist
there), a ``--conf`` file and ``--amend KEY VALUE ...`` pairs are merged into it with the reference's checks --
unknown key -> ``KeyError('<k> is not a valid config key')`` (``get_config.py:107-108``), type mismatch ->
``ValueError('Type mismatch ...')`` (``:111-123``), the ``LOG`` table is never merged (``:103-105``) -- and dumps are
key-sorted TOML (``_sort_dict`` + ``toml.dumps``, ``:11-21,77-96``).  Differences, all deliberate: the configuration is an
object (``Config``) instead of module state created at import time with an ``assert`` on the current directory, and
``cfg_from_list`` raises ``KeyError`` instead of failing an ``assert`` on an unknown key.

``detect_config`` maps the ``cfg.TEST.*`` keys the hot path reads onto ``detector.DetectConfig`` (the same table as
``pycaffe._hot_path_cfg`` uses when the reference's own ``cfg`` is importable).
"""
from __future__ import annotations

import os
import os.path as osp
from ast import literal_eval
from collections import OrderedDict

import numpy as np

try:                                    # the reference depends on `toml` (requirements.txt); tomllib only parses
    import toml as _toml
except ImportError:                     # pragma: no cover
    _toml = None
    import tomllib


class Config(dict):
    """Attribute-access dictionary (the reference uses ``easydict.EasyDict``); nested tables become ``Config`` too."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, Config):
            v = Config(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(Config(x) if isinstance(x, dict) and not isinstance(x, Config) else x for x in v)
        super().__setitem__(k, v)

    __setattr__ = __setitem__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def has_key(self, k):               # the reference's py2 spelling, kept for drop-in use
        return k in self


# The schema: every key of ``configs/default.toml`` with its default (values restated from that file so that the test path
# runs without a reference checkout; tests/test_config_native.py asserts equality with the file whenever the reference
# tree is present).  TRAIN is listed because the shipped ``--conf`` files amend TRAIN keys and an unknown key is an error.
DEFAULTS = {'DATA_DIR': '/mnt/WIDER_FACE',
 'EPS': 1e-14,
 'EXP_DIR': 'face',
 'MAX_RESOLUTION': 16,
 'NAME': 'face',
 'PIXEL_MEANS': [[[102.9801, 115.9465, 122.7717]]],
 'RNG_SEED': 3,
 'USE_GPU_NMS': True,
 'DEBUG': False,
 'PDB': False,
 'MISC': {'MIMIC_EVAL_BUG': True, 'ACCURACY_THRESHOLD': 0.9},
 'TENSORBOARD': {'ENABLE': False, 'HOSTNAME': 'example.com', 'PORT': 8889},
 'MODEL': {'DIAGNOSE': '', 'DIFFERENT_DILATION': {'ENABLE': False}, 'HACK': {'TRAIN': '', 'TEST': ''}},
 'TRAIN': {'ANCHOR_MIN_SIZE': 4,
           'ANCHOR_N_POST_NMS': 300,
           'ANCHOR_N_PRE_NMS': 1000,
           'ANCHOR_NEGATIVE_OVERLAP': 0.3,
           'ANCHOR_POSITIVE_OVERLAP': 0.5,
           'ANCHOR_REGRESSION_OVERLAP': 0.3,
           'ASPECT_GROUPING': True,
           'BBOX_INSIDE_WEIGHTS': [1, 1, 1, 1],
           'BG_THRESH_HI': 0.5,
           'BG_THRESH_LOW': 0,
           'DB': 'wider_train',
           'IMS_PER_BATCH': 1,
           'ITERS': 60000,
           'ITERSIZE': 2,
           'LR_POLICY': 'STEP',
           'ORIG_SIZE': False,
           'POSITIVE_MINING': True,
           'PRETRAINED': '/mnt/WIDER_FACE/imagenet_models/VGG16.caffemodel',
           'PROTOTXT': 'models/train_template.prototxt',
           'SNAPSHOT': 1000,
           'SNAPSHOT_INFIX': '',
           'SOLVER': 'models/solver_template.prototxt',
           'STEPSIZE': 46000,
           'STEPVALUE': [21000, 42000],
           'WEIGHT_DECAY': 0.00025,
           'USE_FLIPPED': True,
           'GPU_ID': [0, 1, 2, 3],
           'LR': {'BASELR': 0.004, 'BACKBONE_MULT': 2.0, 'HEAD_MULT': 1.0},
           'SCALES': {'MODE': 'SHORT_SIDE', 'SHORT_SIDE': [400, 800, 1200], 'MAX_SIZE': 2000},
           'AUGMENT': {'ENABLE': True,
                       'BRIGHTNESS': {'PROB': 0.5, 'DELTA': 32.0},
                       'CONTRAST': {'PROB': 0.5, 'LOWER': 0.5, 'UPPER': 1.5},
                       'SATURATION': {'PROB': 0.5, 'LOWER': 0.5, 'UPPER': 1.5},
                       'HUE': {'PROB': 0.5, 'DELTA': 18.0},
                       'CROP': {'PROB': 0.5,
                                'LOWER': 0.6,
                                'UPPER': 1.0,
                                'POSITIVE_ENFORCE': True,
                                'MAX_TRIES': 50,
                                'KEEP_ONLY_CENTER_INSIDE': True}},
           'DISABLE_EASY_IMAGE': {'ENABLE': False, 'THRESHOLD': 1.0, 'PROB': 0.5, 'SMOOTH': False},
           'ANCHOR_SAMPLING': {'ANCHORS_PER_BATCH': 256,
                               'ANCHOR_FG_FRACTION': 0.25,
                               'ANCHOR_NUM_METHOD': 'fixed_num',
                               'BATCH_POS_NEG_RATIO': 0.33}},
 'TEST': {'ANCHOR_MIN_SIZE': 0,
          'ANCHOR_N_POST_NMS': -1,
          'DB': 'wider_val',
          'FLIP': True,
          'LEVEL': [],
          'MAX_SIZE': 2000,
          'MODEL': '',
          'NO_CACHE': False,
          'NMS_THRESH': 0.4,
          'NMS_METHOD': 'BBOX_VOTE',
          'N_DETS_PER_MODULE': 10000,
          'ORIG_SIZE': False,
          'PYRAMID_BASE_SIZE': [800, 1200],
          'PROTOTXT': 'models/test_template.prototxt',
          'SCALES': [100, 300, 600, 1000, 1400],
          'SCORE_THRESH': 0.002,
          'GPU_ID': [0, 1, 2, 3],
          'IOU_THRESH': 0.5,
          'DEMO': {'ENABLE': False, 'IMAGE': 'demo/demo.jpg'}}}


def write_builtin_tree(root_dir: str) -> str:
    """A self-contained working directory for ``run_test``: ``configs/default.toml`` from ``DEFAULTS``, the shipped
    ``configs/smallhardface.toml`` amendments and the two deploy templates from ``models.build_test_net``."""
    from . import caffe_proto as cp
    from .models import build_test_net
    os.makedirs(osp.join(root_dir, "configs"), exist_ok=True)
    os.makedirs(osp.join(root_dir, "models"), exist_ok=True)
    with open(osp.join(root_dir, "configs", "default.toml"), "w") as f:
        f.write(_toml.dumps(DEFAULTS))
    with open(osp.join(root_dir, "configs", "smallhardface.toml"), "w") as f:     # the shipped WIDER configuration
        f.write(_toml.dumps({"MODEL": {"DIFFERENT_DILATION": {"ENABLE": True}},
                             "TRAIN": {"DISABLE_EASY_IMAGE": {"ENABLE": True, "THRESHOLD": 0.85, "PROB": 0.7, "SMOOTH": True}}}))
    for name, dil in (("test_template.prototxt", False), ("test_different_dilation_template.prototxt", True)):
        with open(osp.join(root_dir, "models", name), "w") as f:
            f.write(cp.format_text(build_test_net(dilation=dil)))
    return root_dir


def sort_dict(d):
    """``get_config.py:11-21``: keys sorted at every level."""
    res = OrderedDict(sorted(d.items()))
    for k, v in res.items():
        if isinstance(v, dict):
            res[k] = sort_dict(v)
    return res


def _load_toml(path):
    if _toml is not None:
        return _toml.load(path)
    with open(path, "rb") as f:         # pragma: no cover
        return tomllib.load(f)


def load_default(root_dir: str, default_path: str | None = None) -> Config:
    """``get_config.py:24-47``: parse ``configs/default.toml``, add the empty ``LOG`` table, ``ROOT_DIR``, the joined
    ``DATA_DIR`` and ``DEBUG`` from the environment.  ``root_dir`` is the checkout the relative paths refer to."""
    import copy
    path = default_path or osp.join(root_dir, "configs", "default.toml")
    if default_path == "builtin":
        d = copy.deepcopy(DEFAULTS)
    elif not osp.isfile(path):
        raise FileNotFoundError("The default config is not found in {}!".format(path))
    else:
        d = _load_toml(path)
    d.update({"LOG": {}})
    cfg = Config(sort_dict(d))
    cfg.ROOT_DIR = osp.abspath(root_dir)
    cfg.DATA_DIR = osp.join(cfg.ROOT_DIR, cfg.DATA_DIR)
    cfg.DEBUG = os.environ.get("DEBUG") == "1"
    return cfg


def merge_a_into_b(a, b) -> None:
    """``get_config.py:96-133``: merge ``a`` into ``b``, clobbering; ``a`` may only name keys ``b`` has, with equal types."""
    if not isinstance(a, Config):
        return
    for k, v in a.items():
        if k == "LOG":
            continue
        if k not in b:
            raise KeyError("{} is not a valid config key".format(k))
        old_type = type(b[k])
        if old_type is not type(v):
            if isinstance(b[k], np.ndarray):
                v = np.array(v, dtype=b[k].dtype)
            elif isinstance(b[k], str) and isinstance(v, str):
                pass
            else:
                raise ValueError("Type mismatch ({} vs. {}) for config key: {}".format(type(b[k]), type(v), k))
        if isinstance(v, Config):
            try:
                merge_a_into_b(a[k], b[k])
            except Exception:
                print("Error under config key: {}".format(k))
                raise
        else:
            b[k] = v


def cfg_from_file(cfg: Config, filename: str) -> Config:
    """``get_config.py:136-139``."""
    merge_a_into_b(Config(_load_toml(filename)), cfg)
    return cfg


def cfg_from_list(cfg: Config, cfg_list) -> Config:
    """``get_config.py:142-158``: ``['TEST.DB', 'general_png', 'TEST.GPU_ID', '[0]', ...]``; values go through
    ``literal_eval`` and stay strings when that fails.  No type check (the reference has none here)."""
    if len(cfg_list) % 2:
        raise ValueError("--amend takes KEY VALUE pairs")
    for k, v in zip(cfg_list[0::2], cfg_list[1::2]):
        keys = k.split(".")
        d = cfg
        for sub in keys[:-1]:
            if sub not in d:
                raise KeyError("{} is not a valid config key".format(k))
            d = d[sub]
        if keys[-1] not in d:
            raise KeyError("Please put {} in default.toml".format(keys[-1]))
        try:
            value = literal_eval(v)
        except Exception:
            value = v
        d[keys[-1]] = value
    return cfg


def _plain(d):
    return {k: (_plain(v) if isinstance(v, dict) else v) for k, v in d.items()}


def cfg_dumps(cfg) -> str:
    """Key-sorted TOML text (``cfg_print`` / ``cfg_dump``, ``get_config.py:69-78``)."""
    if _toml is None:                   # pragma: no cover
        raise RuntimeError("the `toml` package is needed to write configuration dumps")
    return _toml.dumps(sort_dict(_plain(cfg)))


def cfg_dump(cfg, file) -> None:
    file.write(cfg_dumps(cfg))


def cfg_table(cfg) -> str:
    """``get_config.py:81-93``: the markdown table TensorBoard receives."""
    table = "|key|value|\n|---|---|\n"
    for raw_line in cfg_dumps(cfg).split("\n"):
        line = raw_line.split("=")
        if len(line) == 1 and len(line[0]) > 0:
            table += "|**{}**||\n".format(line[0])
        elif len(line) == 2:
            table += "|{}|{}|\n".format(line[0], line[1])
    return table


def get_output_dir(cfg: Config, imdb_name: str, net_name: str | None = None, output_dir: str = "output", idx: int = -1) -> str:
    """``get_config.py:50-66``: ``<ROOT_DIR>/<output_dir>/<EXP_DIR>/<imdb>[/<net>][/<idx>]``, created on demand."""
    outdir = osp.abspath(osp.join(cfg.ROOT_DIR, output_dir, cfg.EXP_DIR, imdb_name))
    if net_name is not None:
        outdir = osp.join(outdir, net_name)
    if idx >= 0:
        outdir = osp.join(outdir, str(idx))
    os.makedirs(outdir, exist_ok=True)
    return outdir


def detect_config(cfg: Config, **overrides):
    """``cfg.TEST.*`` / ``cfg.*`` -> ``detector.DetectConfig`` (the keys ``lib/test.py`` and ``proposal_layer.py`` read)."""
    from .detector import DetectConfig
    T = cfg.TEST
    kw = dict(scales=tuple(T.SCALES), pyramid_base_size=tuple(T.PYRAMID_BASE_SIZE), max_size=int(T.MAX_SIZE),
              flip=bool(T.FLIP), nms_method=str(T.NMS_METHOD), nms_thresh=float(T.NMS_THRESH),
              n_dets_per_module=int(T.N_DETS_PER_MODULE), score_thresh=float(T.SCORE_THRESH),
              anchor_min_size=float(T.ANCHOR_MIN_SIZE), max_resolution=int(cfg.MAX_RESOLUTION),
              pixel_means=tuple(np.asarray(cfg.PIXEL_MEANS, dtype=np.float64).reshape(-1).tolist()),
              nms_mode=1 if cfg.USE_GPU_NMS else 0)
    kw.update(overrides)
    return DetectConfig(**kw)
