"""Generate tests/golden/*.npz by running the REFERENCE's own code in this container.

Run from the repo root where /root/reference is mounted (it is not on the GPU box, which only sees
the committed .npz files):

    python tests/golden/make_golden.py

What runs, and how it is made runnable on python 3.12 / NumPy 2.3 without editing the files:
  * ``lib/layers/generate_anchors.py``, ``lib/utils/bbox_transform.py``, ``lib/layers/proposal_layer.py``,
    ``lib/nms/py_cpu_nms.py`` and the ``bbox_vote`` function of ``lib/test.py`` are read from
    /root/reference, passed through a minimal py2->py3 source transform (print statements, xrange,
    integer ``/`` used as an index, ``np.float``) and exec'd against stub ``caffe`` / ``cfg`` / ``nms``
    modules.  No arithmetic is changed.
  * ``lib/nms/cpu_nms.pyx`` and ``lib/utils/bbox.pyx`` are copied to a temp dir, four NumPy type
    aliases that no longer exist are renamed (``np.float``->double/float64, ``np.int_t``->``np.intp_t``,
    ``np.int``->``np.intp``), and compiled with Cython.  No arithmetic is changed.
Inputs are seeded and tie-free in score so the reference's unstable ``argsort()[::-1]`` is well defined.
"""
import os
import re
import subprocess
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def py2to3(src: str) -> str:
    out = []
    for line in src.splitlines():
        m = re.match(r"^(\s*)print\s+(?!\()(.*)$", line)
        if m:
            line = "%sprint(%s)" % (m.group(1), m.group(2))
        out.append(line)
    s = "\n".join(out) + "\n"
    s = s.replace("xrange(", "range(")
    s = s.replace("scores.shape[1] / (A * self._num_feats)", "scores.shape[1] // (A * self._num_feats)")
    s = s.replace("self._feat_stride[i / len(self._shifts)**", "self._feat_stride[i // len(self._shifts)**")
    s = s.replace("np.float)", "float)")
    s = s.replace("yaml.load(self.param_str_)", "yaml.safe_load(self.param_str_)")
    s = s.replace("yaml.load(self.param_str)", "yaml.safe_load(self.param_str)")
    return s


def load_module(name, path, extra=None):
    mod = types.ModuleType(name)
    mod.__file__ = path
    if extra:
        mod.__dict__.update(extra)
    sys.modules[name] = mod
    exec(compile(py2to3(open(path).read()), path, "exec"), mod.__dict__)
    return mod


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def install_stubs():
    caffe = types.ModuleType("caffe")

    class Layer(object):
        pass
    caffe.Layer = Layer
    sys.modules["caffe"] = caffe
    cfg = _Cfg(TEST=_Cfg(N_DETS_PER_MODULE=10000, SCORE_THRESH=0.002, ANCHOR_MIN_SIZE=0, NMS_THRESH=0.4),
               TRAIN=_Cfg(), USE_GPU_NMS=False, GPU_ID=0)
    for pkg in ("utils", "lib", "lib.layers", "nms"):
        sys.modules.setdefault(pkg, types.ModuleType(pkg))
    gc = types.ModuleType("utils.get_config")
    gc.cfg = cfg
    sys.modules["utils.get_config"] = gc
    nw = types.ModuleType("nms.nms_wrapper")
    nw.nms = lambda dets, thresh, force_cpu=False: list(range(len(dets)))
    sys.modules["nms.nms_wrapper"] = nw
    return cfg


class FakeBlob(object):
    def __init__(self, arr=None):
        self.data = arr

    def reshape(self, *shape):
        self.data = np.zeros(shape, dtype=np.float32)


def softmax2(bg, fg):
    m = np.maximum(bg, fg)
    e0, e1 = np.exp(bg - m), np.exp(fg - m)
    return (e0 / (e0 + e1)).astype(np.float32), (e1 / (e0 + e1)).astype(np.float32)


def clustered_dets(rng, n, span=400.0, nclusters=25):
    c = rng.rand(nclusters, 2) * span
    k = rng.randint(0, nclusters, n)
    ctr = c[k] + rng.randn(n, 2) * 6
    wh = np.exp(rng.randn(n, 2) * 0.3) * (20 + 30 * rng.rand(nclusters)[k])[:, None]
    d = np.empty((n, 5), dtype=np.float32)
    d[:, 0:2] = ctr - wh / 2
    d[:, 2:4] = ctr + wh / 2
    s = rng.permutation(n).astype(np.float64) / n * 0.94 + 0.055 + rng.rand(n) * 1e-4   # tie-free
    d[:, 4] = s
    assert len(np.unique(d[:, 4])) == n
    return d


def main():
    cfg = install_stubs()
    ga = load_module("lib.layers.generate_anchors", REF + "/lib/layers/generate_anchors.py")
    sys.modules["utils.bbox_transform"] = load_module("utils.bbox_transform", REF + "/lib/utils/bbox_transform.py")
    pl = load_module("lib.layers.proposal_layer", REF + "/lib/layers/proposal_layer.py")
    pynms = load_module("nms.py_cpu_nms", REF + "/lib/nms/py_cpu_nms.py")

    # ---- anchors ------------------------------------------------------------------------
    anchors = ga.generate_anchors(scales=np.array([1, 2, 4]), base_size=16, ratios=np.array([1, ]),
                                  shifts=np.array([0]), strides=np.array([8, 8, 8]))
    np.savez(os.path.join(OUT, "anchors.npz"), anchors=anchors,
             default=ga.generate_anchors(strides=np.array([16, 16, 16])))

    # ---- ProposalLayer.forward ------------------------------------------------------------
    rng = np.random.RandomState(3)
    cases = {}
    for tag, (h, w, imh, imw, scale, shift, spread) in {
        "small": (6, 9, 45, 70, 1.0, -1.0, 2.5),
        "level": (28, 28, 220, 217, 0.2734375, -3.0, 2.0),
        "allbelow": (4, 5, 32, 40, 1.0, -12.0, 0.5),
        "wide": (10, 23, 80, 184, 1.3671875, 0.0, 3.0),
    }.items():
        logits_bg = rng.randn(3, h, w) * spread
        logits_fg = rng.randn(3, h, w) * spread + shift
        pbg, pfg = softmax2(logits_bg, logits_fg)
        cls = np.concatenate([pbg, pfg])[None].astype(np.float32)                # (1,6,h,w)
        assert len(np.unique(cls[0, 3:])) == 3 * h * w, "ties in golden scores"
        deltas = (rng.randn(1, 12, h, w) * np.array([0.3, 0.3, 0.4, 0.4] * 3)[None, :, None, None]).astype(np.float32)
        im_info = np.array([[imh, imw, scale]], dtype=np.float32)
        layer = pl.ProposalLayer()
        layer.param_str = "{'feat_stride': [8,8,8],'scales': [1,2,4], 'ratios':[1,]}"
        layer.phase = 1
        bottom = [FakeBlob(cls), FakeBlob(deltas), FakeBlob(im_info)]
        top = [FakeBlob(), FakeBlob()]
        layer.setup(bottom, top)
        layer.forward(bottom, top)
        cases[tag + "_cls"] = cls
        cases[tag + "_deltas"] = deltas
        cases[tag + "_im_info"] = im_info
        cases[tag + "_boxes"] = top[0].data.copy()
        cases[tag + "_probs"] = top[1].data.copy()
        print("proposal", tag, "R =", top[0].data.shape[0])
    np.savez_compressed(os.path.join(OUT, "proposal.npz"), **cases)

    # ---- compile the Cython modules -----------------------------------------------------------
    tmp = tempfile.mkdtemp(prefix="refcy_")
    src = open(REF + "/lib/nms/cpu_nms.pyx").read()
    src = src.replace("np.float thresh", "double thresh").replace("np.int_t", "np.intp_t").replace("dtype=np.int)", "dtype=np.intp)")
    open(tmp + "/ref_cpu_nms.pyx", "w").write(src)
    src = open(REF + "/lib/utils/bbox.pyx").read()
    src = src.replace("DTYPE = np.float\n", "DTYPE = np.float64\n").replace("ctypedef np.float_t DTYPE_t", "ctypedef np.float64_t DTYPE_t")
    open(tmp + "/ref_bbox.pyx", "w").write(src)
    open(tmp + "/setup.py", "w").write(
        "from setuptools import setup, Extension\nfrom Cython.Build import cythonize\nimport numpy as np\n"
        "setup(ext_modules=cythonize([Extension('ref_cpu_nms',['ref_cpu_nms.pyx'],include_dirs=[np.get_include()]),"
        "Extension('ref_bbox',['ref_bbox.pyx'],include_dirs=[np.get_include()])], language_level=2))\n")
    subprocess.check_call([sys.executable, "setup.py", "-q", "build_ext", "--inplace"], cwd=tmp,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sys.path.insert(0, tmp)
    import ref_bbox
    import ref_cpu_nms

    # ---- NMS ------------------------------------------------------------------------------
    out = {}
    for tag, n in (("n300", 300), ("n1", 1), ("n1500", 1500)):
        d = clustered_dets(rng, n)
        out[tag + "_dets"] = d
        for thr in (0.4, 0.7, 0.3):
            out["%s_cpu_%g" % (tag, thr)] = np.array(ref_cpu_nms.cpu_nms(d, thr), dtype=np.int64)
            out["%s_py_%g" % (tag, thr)] = np.array(pynms.py_cpu_nms(d, thr), dtype=np.int64)
    # integer-coordinate boxes: IoU lands exactly on thresholds, exercising >= vs >
    g = np.array([[0, 0, 9, 9, 0.9], [0, 0, 9, 3, 0.8], [0, 5, 9, 9, 0.7], [0, 0, 9, 6, 0.6],
                  [20, 20, 29, 29, 0.5], [20, 20, 29, 23, 0.45]], dtype=np.float32)
    out["grid_dets"] = g
    for thr in (0.4, 0.5, 0.7):
        out["grid_cpu_%g" % thr] = np.array(ref_cpu_nms.cpu_nms(g, thr), dtype=np.int64)
        out["grid_py_%g" % thr] = np.array(pynms.py_cpu_nms(g, thr), dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **out)

    # ---- bbox_vote (lib/test.py:181-217, function source extracted) ------------------------------
    tsrc = open(REF + "/lib/test.py").read()
    start = tsrc.index("def bbox_vote(det):")
    end = tsrc.index("def inference_worker(")
    ns = {"np": np, "cfg": cfg}
    exec(compile(py2to3(tsrc[start:end]), "lib/test.py:bbox_vote", "exec"), ns)
    import warnings
    warnings.simplefilter("ignore")
    out = {}
    for tag, n in (("n0", 0), ("n1", 1), ("n2far", 2), ("n400", 400), ("n3000", 3000)):
        if tag == "n2far":
            d = np.array([[0, 0, 10, 10, 0.9], [100, 100, 120, 130, 0.8]], dtype=np.float32)
        else:
            d = clustered_dets(rng, n) if n else np.zeros((0, 5), dtype=np.float32)
        out[tag + "_dets"] = d
        r = ns["bbox_vote"](d.copy())
        out[tag + "_vote"] = r
        out[tag + "_vote_dtype"] = np.array(str(r.dtype))
        print("vote", tag, d.shape, "->", r.shape, r.dtype)
    np.savez_compressed(os.path.join(OUT, "bbox_vote.npz"), **out)

    # ---- bbox_overlaps x3 -----------------------------------------------------------------------
    b = clustered_dets(rng, 60)[:, :4].astype(np.float64)
    q = clustered_dets(rng, 45)[:, :4].astype(np.float64)
    np.savez_compressed(os.path.join(OUT, "bbox_overlaps.npz"), boxes=b, query=q,
                        iou=ref_bbox.bbox_overlaps(b, q), ioa=ref_bbox.bbox_overlaps_IoA(b, q),
                        itself=ref_bbox.bbox_overlaps_itself(b, q),
                        ioa_sq=ref_bbox.bbox_overlaps_IoA(b, b))
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()


def wider_writer_inputs():
    """Inputs of the detection-writer golden (tests/golden/wider_writer/): three WIDER-style relative paths; float64 rows,
    an empty image, float32 rows; scores 1.0 and 1e-5 exercise '%g'."""
    rng = np.random.RandomState(5)
    paths = ['0--Parade/0_Parade_marchingband_1_465.jpg', '12--Group/12_Group_Group_12_Group_Group_12_10.jpg',
             '2--Demonstration/2_Demonstration_Political_Rally_2_71.jpg']

    def dets(n):
        x = rng.rand(n, 2) * 900
        wh = rng.rand(n, 2) * 200 + 3
        sc = np.concatenate([rng.rand(n - 2), [1.0, 1e-5]]) if n >= 2 else rng.rand(n)
        return np.hstack([x, x + wh, sc[:, None]]).astype(np.float64)
    return paths, [dets(7), np.zeros((0, 5)), dets(3).astype(np.float32)]


def make_wider_writer_golden(reference_root, out_dir):
    """Runs the REFERENCE's own writer (lib/datasets/wider.py:143-170, imported unmodified through the py2 hook, cwd = the
    reference root because lib/utils/get_config.py:25 reads configs/default.toml relatively) on a stub imdb."""
    import shutil
    os.chdir(reference_root)
    from smallhardface_b200 import compat
    compat.install(reference_root=reference_root)
    import datasets.wider as W

    class Stub(object):
        pass
    stub = Stub()
    stub._image_paths, boxes = wider_writer_inputs()
    shutil.rmtree(out_dir, ignore_errors=True)
    W.wider.write_detections(stub, [[[] for _ in boxes], boxes], out_dir)
