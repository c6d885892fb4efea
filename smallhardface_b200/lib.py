"""ctypes binding of libshf_b200.so (the C ABI in include/shf_b200.h).

The product path has no CPU fallback: if the shared library is missing or a call fails, this module
raises -- it never routes around the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC, "libshf_b200.so")

_lib = None

c_void_p, c_int, c_float, c_double, c_ll = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_longlong

FMT_H2, FMT_HF8 = 0, 1          # SHF_FMT_* of include/shf_b200.h
ABI_VERSION = 8                 # shf_abi_version() of the library these signatures describe

# name -> (restype, argtypes); must list every symbol include/shf_b200.h declares
SIGNATURES = {
    "shf_last_error": (C.c_char_p, []),
    "shf_abi_version": (c_int, []),
    "shf_device_info": (c_int, [c_int, C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_ll)]),
    "shf_conv_igemm": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p]),
    "shf_conv_igemm_pool": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p,
                                    c_void_p]),
    "shf_conv_igemm_strided": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p]),
    "shf_conv_igemm_res": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "shf_conv3x3_s2": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                               c_int, c_int, c_int, c_void_p, c_void_p]),
    "shf_eltwise_sum": (c_int, [C.POINTER(c_void_p), C.POINTER(c_float), c_int, c_void_p, c_ll, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_void_p]),
    "shf_pool_out_size": (c_int, [c_int, c_int, c_int, c_int, c_int]),
    "shf_maxpool": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_void_p]),
    "shf_conv_first": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_void_p, c_void_p]),
    "shf_conv_first_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                                  c_int, c_int, c_void_p, c_void_p]),
    "shf_set_conv_impl": (c_int, [c_int]),
    "shf_conv1_c3": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "shf_conv1_tc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int, c_int,
                             c_void_p, c_void_p]),
    "shf_maxpool2x2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "shf_deconv_depthwise": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "shf_h2_to_nchw": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "shf_nchw_to_h2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "shf_preprocess_level": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_double, c_int,
                                     C.POINTER(c_double), c_void_p]),
    "shf_preprocess_level_batched": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_double,
                                             c_int, C.POINTER(c_double), c_void_p]),
    "shf_head_decode": (c_int, [C.POINTER(c_void_p), c_ll, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                C.POINTER(c_float), c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "shf_head_decode_batched": (c_int, [C.POINTER(c_void_p), c_ll, c_ll, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        C.POINTER(c_float), c_int, c_int, c_int, c_int, c_float, c_float, c_float, c_float,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "shf_gather_dets_batched": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_float, c_float,
                                        c_void_p]),
    "shf_sort_keys_workspace": (c_ll, [c_int]),
    "shf_sort_keys": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "shf_proposal_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_float,
                                    c_float, c_void_p]),
    "shf_postprocess_workspace": (c_ll, [c_int, c_int]),
    "shf_postprocess": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_int, c_int, c_void_p, c_void_p,
                                c_void_p, c_int, c_void_p, c_ll, c_void_p]),
    "_nms": (None, [C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_float), c_int, c_int, c_float, c_int]),
    "shf_nms_host": (c_int, [C.POINTER(c_int), C.POINTER(c_int), C.POINTER(c_float), c_int, c_int, c_double, c_int,
                             c_int]),
    "shf_bbox_vote_host": (c_int, [C.POINTER(c_float), C.POINTER(c_int), C.POINTER(c_float), c_int, c_double, c_int]),
    "shf_bbox_overlaps": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "shf_debug_conv_direct": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
}


class ShfError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libshf_b200.so (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if r.returncode != 0:
        raise ShfError("building libshf_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def load():
    """Load the library (once) and attach signatures.  Raises ShfError if it is missing: there is no
    fallback implementation of the hot path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ShfError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(the CUDA extension is required; there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise ShfError("libshf_b200.so does not export %s (stale build?)" % name)
        fn.restype = res
        fn.argtypes = args
    if lib.shf_abi_version() != ABI_VERSION:
        raise ShfError("%s is a stale build (ABI %d, this package needs %d): run `make -C %s`"
                       % (LIB_PATH, lib.shf_abi_version(), ABI_VERSION, CSRC))
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().shf_last_error()
        raise ShfError("%s failed (%d): %s" % (what or "shf call", rc, msg.decode() if msg else "?"))


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
