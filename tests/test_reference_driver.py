"""`train_test.py --train false` of the reference, UNMODIFIED, on this framework's drop-in modules (north_star's acceptance
sentence).

The reference tree is not part of this repo and does not exist on the GPU box, so the test looks for it at
$SHF_REFERENCE_ROOT, /root/reference or <repo>/_refdata/reference (a git-ignored copy pushed as DATA for a gpurun call:
`tools/gpu_reference_driver.sh`) and skips otherwise.  It copies the tree to a scratch directory (`get_output_dir` writes
under the tree), writes a few PNGs for the reference's own `general_png` imdb (lib/datasets/general.py:19-79) and a
synthetic caffemodel, and runs tools/run_reference_driver.py.  Everything from `test_net` (lib/test.py:290-356) down to
`net.forward` is then the reference's code.  Checks: the run completes, the reference's detection writer produced one
text file per image in its `%d %d %d %d %g` format, and the detections it pickled equal `Detector.detect` on the same
files (which shares no driver code with lib/test.py) within the north_star tolerances.
"""
import glob
import os
import pickle
import shutil
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.environ.get("SHF_REFERENCE_ROOT", ""), "/root/reference", os.path.join(ROOT, "_refdata", "reference")]
REF = next((c for c in CANDIDATES if c and os.path.exists(os.path.join(c, "train_test.py"))), None)


@pytest.mark.skipif(REF is None, reason="reference tree not present (see tools/gpu_reference_driver.sh)")
def test_train_test_py_runs_unmodified_and_matches_the_detector(tmp_path):
    import cv2
    from smallhardface_b200 import deploy
    from smallhardface_b200.detector import DetectConfig, Detector
    ref = str(tmp_path / "reference")
    shutil.copytree(REF, ref, ignore=shutil.ignore_patterns("output", ".git"))
    os.chmod(ref, 0o755)
    for dirpath, dirnames, _ in os.walk(ref):
        for d in dirnames:
            os.chmod(os.path.join(dirpath, d), 0o755)
    imgs = tmp_path / "imgs"
    imgs.mkdir()
    sizes = [(96, 128), (80, 80), (64, 112), (120, 90)]
    for i, hw in enumerate(sizes):
        cv2.imwrite(str(imgs / ("im%d.png" % i)), deploy.synthetic_image(20 + i, hw))
    proto, model = deploy.write_synthetic_deployment(str(tmp_path / "deploy"), dilation=True)
    cmd = [sys.executable, os.path.join(ROOT, "tools", "run_reference_driver.py"), "--reference-root", ref, "--",
           "--train", "false", "--conf", "configs/smallhardface.toml", "--amend", "DATA_DIR", str(imgs), "TEST.DB",
           "general_png", "TEST.MODEL", model, "TEST.GPU_ID", "[0]", "TEST.NO_CACHE", "False"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    logs = glob.glob(os.path.join(ref, "output", "face", "general_png", "*", "stderr.log"))
    err = open(logs[0]).read() if logs else ""
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:], err[-3000:])
    assert "All Done!" in err                                   # lib/test.py:356 (the driver redirects stderr to this file)
    out_dir = os.path.dirname(logs[0])
    # the deploy prototxt the reference's manipulate_test wrote (dilation template + dim_red splice)
    assert "conv4_fuse_final_dim_red" in open(os.path.join(out_dir, "test.prototxt")).read()
    with open(os.path.join(out_dir, "detections.pkl"), "rb") as f:
        dets = pickle.load(f)                                   # [class][image] -> (n, 5), class 0 = background (empty)
    assert len(dets) == 2 and len(dets[1]) == len(sizes)
    # the reference's imdb order = os.walk order of the PNG directory
    paths = []
    for root, _, files in os.walk(str(imgs)):
        paths += [os.path.join(root, f) for f in files if f.endswith(".png")]
    det = Detector(proto, model, "cuda:0", DetectConfig())
    ours = det.detect([cv2.imread(p) for p in paths])
    worst = 0.0
    for i, p in enumerate(paths):
        got, want = np.asarray(dets[1][i]), ours[i]
        assert got.shape == want.shape, (p, got.shape, want.shape)
        assert np.abs(got[:, 4] - want[:, 4]).max() < 1e-3
        worst = max(worst, float(np.abs(got[:, :4] - want[:, :4]).max()))
        # lib/datasets/general.py:44-69: path, count, then `x y w h score` rows
        txt = glob.glob(os.path.join(out_dir, "**", os.path.basename(p).replace("png", "txt")), recursive=True)
        lines = open(txt[0]).read().splitlines()
        assert lines[0] == p and int(lines[1]) == len(got) == len(lines) - 2
        first = lines[2].split()
        d0 = got[0]
        assert [int(v) for v in first[:4]] == [int(d0[0]), int(d0[1]), int(d0[2]) - int(d0[0]), int(d0[3]) - int(d0[1])]
        assert first[4] == "%g" % d0[4]
    print("unmodified train_test.py vs Detector.detect: %d images, worst box difference %.2e px" % (len(paths), worst))
    assert worst < 1e-2
