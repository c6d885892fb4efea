"""`import caffe` drop-in (see smallhardface_b200/pycaffe.py; reference: caffe/python/caffe/__init__.py:1-8)."""
from smallhardface_b200.pycaffe import (Net, Blob, Layer, TRAIN, TEST, set_mode_gpu, set_mode_cpu, set_device, set_fast_min_scale,  # noqa: F401
                                        SGDSolver, NCCL, set_random_seed, set_solver_count, set_solver_rank,
                                        set_multiprocess, init_log, log, layer_type_list)
from . import proto  # noqa: F401
from . import draw   # noqa: F401

__version__ = "1.0.0-shf_b200"
