"""Python 3 configuration / deploy-prototxt plumbing (smallhardface_b200.config, .prototxt, .run_test)
against golden files written by the reference's own get_config.py / manipulate.py (tests/golden/make_config_golden.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden", "config")
REF = "/root/reference"

from smallhardface_b200 import caffe_proto as cp
from smallhardface_b200 import config as C
from smallhardface_b200 import prototxt

AMEND = ["DATA_DIR", "/data/images", "TEST.DB", "general_png", "TEST.MODEL", "/data/final.caffemodel", "TEST.GPU_ID", "[0]",
         "TEST.SCALES", "[300, 600]", "NAME", "golden"]


def _golden_cfg(root):
    cfg = C.load_default(root)
    C.cfg_from_file(cfg, os.path.join(root, "configs", "smallhardface.toml"))
    cfg.TEST.NO_CACHE = True
    C.cfg_from_list(cfg, AMEND)
    cfg.LOG.CMD = "golden"
    cfg.LOG.TIME = "2026_01_01_00_00_00"
    cfg.ROOT_DIR = "/root/reference"
    return cfg


@pytest.fixture()
def tree(tmp_path):
    return C.write_builtin_tree(str(tmp_path / "tree"))


def test_config_dump_and_table_equal_the_reference_golden(tree):
    cfg = _golden_cfg(tree)
    sub = {k: cfg[k] for k in cfg if k != "TRAIN"}
    assert C.cfg_dumps(sub) == open(os.path.join(GOLD, "cfgs_test.toml")).read()
    assert C.cfg_table(sub) == open(os.path.join(GOLD, "cfg_table.md")).read()
    assert cfg.TEST.SCALES == [300, 600] and cfg.TEST.GPU_ID == [0] and cfg.MODEL.DIFFERENT_DILATION.ENABLE is True
    assert cfg.TRAIN.DISABLE_EASY_IMAGE.THRESHOLD == 0.85


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")
def test_builtin_schema_equals_default_toml_of_the_reference():
    import toml
    assert C.DEFAULTS == toml.load(REF + "/configs/default.toml")
    a, b = C.load_default(REF), C.load_default(REF, default_path="builtin")
    assert dict(a) == dict(b)


def test_merge_errors_like_get_config(tree, tmp_path):
    cfg = C.load_default(tree)
    bad = tmp_path / "bad.toml"
    bad.write_text("NOT_A_KEY = 1\n")
    with pytest.raises(KeyError, match="NOT_A_KEY is not a valid config key"):
        C.cfg_from_file(cfg, str(bad))
    bad.write_text("[TEST]\nNMS_THRESH = \"high\"\n")
    with pytest.raises(ValueError, match="Type mismatch"):
        C.cfg_from_file(cfg, str(bad))
    bad.write_text("[LOG]\nanything = 1\n[TEST]\nFLIP = false\n")           # LOG is skipped, the rest merges
    C.cfg_from_file(cfg, str(bad))
    assert cfg.TEST.FLIP is False and "anything" not in cfg.LOG
    with pytest.raises(KeyError):
        C.cfg_from_list(cfg, ["TEST.NOPE", "1"])
    C.cfg_from_list(cfg, ["TEST.NMS_METHOD", "NMS", "TEST.NMS_THRESH", "0.3"])
    assert cfg.TEST.NMS_METHOD == "NMS" and cfg.TEST.NMS_THRESH == 0.3
    with pytest.raises(ValueError):
        C.cfg_from_list(cfg, ["TEST.FLIP"])


def test_detect_config_reads_the_keys_of_the_hot_path(tree):
    cfg = _golden_cfg(tree)
    d = C.detect_config(cfg)
    assert d.scales == (300, 600) and d.flip and d.nms_method == "BBOX_VOTE" and d.nms_thresh == 0.4
    assert d.n_dets_per_module == 10000 and d.score_thresh == 0.002 and d.pixel_means == (102.9801, 115.9465, 122.7717)
    assert d.nms_mode == 1                                                   # USE_GPU_NMS -> gpu_nms ('>' suppression)


@pytest.mark.parametrize("dilation,gold", [(True, "test_dilation.prototxt"), (False, "test_plain.prototxt")])
def test_manipulate_test_equals_the_reference_output(tree, tmp_path, dilation, gold):
    """Same message as the reference's manipulate_test wrote (parsed with our codec on both sides), and the same TEXT as
    protobuf's str() produced."""
    cfg = C.load_default(tree)
    cfg.MODEL.DIFFERENT_DILATION.ENABLE = dilation
    out = tmp_path / "test.prototxt"
    net = prototxt.manipulate_test(cfg, cfg.TEST.PROTOTXT, str(out))
    ref_text = open(os.path.join(GOLD, gold)).read()
    assert cp.format_text(net) == cp.format_text(cp.parse_text(ref_text))
    assert out.read_text() == ref_text
    names = [l.name for l in net.layer]
    assert ("conv4_fuse_final_dim_red" in names) == dilation


def test_general_imdb_and_output_dir(tree, tmp_path):
    from smallhardface_b200 import run_test as R
    imgs = tmp_path / "imgs" / "sub"
    imgs.mkdir(parents=True)
    for n in ("b.png", "a.png", "c.jpg"):
        (imgs / n).write_bytes(b"x")
    cfg = C.load_default(tree)
    C.cfg_from_list(cfg, ["DATA_DIR", str(tmp_path / "imgs"), "TEST.DB", "general_png"])
    imdb = R.get_imdb(cfg, cfg.TEST.DB)
    assert imdb.name == "general_png" and len(imdb) == 2 and all(p.endswith(".png") for p in imdb.image_paths)
    with pytest.raises(KeyError, match="Unknown dataset"):
        R.get_imdb(cfg, "nothing_val")
    with pytest.raises(NotImplementedError):
        R.get_imdb(cfg, "fddb_val")
    with pytest.raises(AssertionError, match="Annotation file not found"):        # wider.py:38-39
        R.get_imdb(cfg, "wider_val")
    out = C.get_output_dir(cfg, imdb.name, "face_x", output_dir=str(tmp_path / "o"))
    assert out == str(tmp_path / "o" / "face" / "general_png" / "face_x") and os.path.isdir(out)
    dets = [[[], []], [np.array([[1.7, 2.2, 30.9, 40.1, 0.93]]), np.zeros((0, 5))]]
    msg = imdb.evaluate_detections(dets, out)
    assert msg.startswith("Detection results wrote to")
    rel = imdb.image_paths[0][1:-4] + ".txt"
    text = open(os.path.join(out, rel)).read().splitlines()
    assert text[0] == imdb.image_paths[0] and text[1] == "1" and text[2] == "1 2 29 38 0.93 "


def test_inference_walks_images_in_size_order_and_returns_them_in_list_order(tree, tmp_path):
    """run_test.inference: same-sized images are grouped into batches (header scan), decoded on worker threads one batch
    ahead, and every result lands in the slot of ITS image -- checked with a stand-in detector that reports what it saw."""
    import cv2
    from smallhardface_b200 import run_test as R
    imgs = tmp_path / "imgs"
    imgs.mkdir()
    rng = np.random.RandomState(2)
    shapes = [(40, 64), (64, 40), (40, 64), (48, 48), (64, 40), (40, 64), (48, 48)]
    for i, (h, w) in enumerate(shapes):
        cv2.imwrite(str(imgs / ("im%02d.png" % i)), np.full((h, w, 3), 10 + i, np.uint8))
    cfg = C.load_default(tree)
    C.cfg_from_list(cfg, ["DATA_DIR", str(imgs), "TEST.DB", "general_png"])
    imdb = R.get_imdb(cfg, cfg.TEST.DB)
    want = {p: cv2.imread(p) for p in imdb.image_paths}

    class FakeDetector:
        batches = []

        def detect(self, images):
            FakeDetector.batches.append([im.shape[:2] for im in images])
            return [np.array([[im.shape[0], im.shape[1], float(im[0, 0, 0]), 0, 1.0]]) for im in images]

    for grouped in (True, False):
        FakeDetector.batches = []
        boxes = R.inference(cfg, imdb, None, 0, len(imdb), batch=2, detector=FakeDetector(), group_by_size=grouped)
        assert len(boxes[1]) == len(imdb)
        for i, p in enumerate(imdb.image_paths):
            h, w, v = boxes[1][i][0, :3]
            assert (h, w) == want[p].shape[:2] and v == float(want[p][0, 0, 0]), (grouped, i)
        mixed = sum(1 for b in FakeDetector.batches if len(set(b)) > 1)
        assert sum(len(b) for b in FakeDetector.batches) == len(imdb)
        if grouped:
            assert mixed <= 2                    # at most the batches that straddle two size groups
    # a sub-range keeps its own slots
    part = R.inference(cfg, imdb, None, 2, 5, batch=2, detector=FakeDetector())
    assert [tuple(b[0, :2]) for b in part[1]] == [want[p].shape[:2] for p in imdb.image_paths[2:5]]


def test_cli_refuses_training_and_unknown_keys(tree):
    from smallhardface_b200 import run_test as R
    with pytest.raises(SystemExit, match="training is outside"):
        R.main(["--root", tree, "--train", "true"])
    with pytest.raises(KeyError):
        R.build_cfg(tree, "configs/smallhardface.toml", ["TEST.NOT_A_KEY", "1"])
    cfg = R.build_cfg(tree, "configs/smallhardface.toml", ["TEST.GPU_ID", "[2, 3]"])
    assert cfg.TEST.NO_CACHE is True and cfg.TEST.GPU_ID == [2, 3] and R._first_gpu(cfg) == 2 and cfg.LOG.TIME


@pytest.mark.gpu
def test_native_driver_equals_detector_and_caches(tmp_path):
    """`python -m smallhardface_b200.run_test` end to end on the B200: config files, deploy prototxt with the dim_red splice,
    general_png imdb, Detector, result files, detections.pkl cache."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import cv2
    from smallhardface_b200 import deploy
    from smallhardface_b200 import run_test as R
    from smallhardface_b200.detector import Detector
    tree = C.write_builtin_tree(str(tmp_path / "tree"))
    proto, model = deploy.write_synthetic_deployment(str(tmp_path / "deploy"), dilation=True)
    imgs = tmp_path / "imgs"
    imgs.mkdir()
    rng = np.random.RandomState(5)
    shapes = [(96, 128), (128, 96), (96, 128)]
    for i, (h, w) in enumerate(shapes):
        cv2.imwrite(str(imgs / ("im%d.png" % i)), rng.randint(0, 256, (h, w, 3)).astype(np.uint8))
    argv = ["--root", tree, "--conf", "configs/smallhardface.toml", "--output", str(tmp_path / "out"), "--batch", "2", "--amend",
            "DATA_DIR", str(imgs), "TEST.DB", "general_png", "TEST.MODEL", model, "TEST.GPU_ID", "[0]", "TEST.SCALES",
            "[100, 300]", "TEST.NO_CACHE", "False"]
    out_dir, dets, result = R.main(argv)
    assert os.path.isfile(os.path.join(out_dir, "test.prototxt")) and os.path.isfile(os.path.join(out_dir, "cfgs.txt"))
    assert os.path.isfile(os.path.join(out_dir, "detections.pkl"))
    cfg = R.build_cfg(tree, "configs/smallhardface.toml", argv[argv.index("--amend") + 1:])
    imdb = R.get_imdb(cfg, cfg.TEST.DB)
    det = Detector(os.path.join(out_dir, "test.prototxt"), model, "cuda:0", C.detect_config(cfg))
    n_rows = 0
    for i, p in enumerate(imdb.image_paths):
        want = det.detect([cv2.imread(p)])[0]
        got = dets[1][i]
        assert got.shape == want.shape and np.abs(got - want).max(initial=0) < 1e-3
        txt = open(os.path.join(out_dir, p[1:-4] + ".txt")).read().splitlines()
        assert txt[0] == p and int(txt[1]) == len(want) == len(txt) - 2
        n_rows += len(want)
    assert n_rows > 0
    # second run with the cache: no Detector is built (TEST.MODEL is bogus), the pickle is used
    import pickle
    with open(os.path.join(out_dir, "detections.pkl"), "rb") as f:
        cached = pickle.load(f)
    assert len(cached[1]) == len(shapes)
    cfg.TEST.MODEL = "/nonexistent"
    d2, _ = R.test_net(cfg, imdb, out_dir, os.path.join(out_dir, "test.prototxt"), no_cache=False)
    assert all(np.array_equal(a, b) for a, b in zip(d2[1], cached[1]))
