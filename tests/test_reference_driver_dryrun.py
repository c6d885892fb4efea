"""CPU dry run of the reference's UNMODIFIED `train_test.py --train false` (row b6 of the coverage table): everything
host-side is the reference's code running through this repo's compat layer; only `caffe.Net` is a CPU stand-in backed by
the oracle (tests/ref_driver_dryrun_boot.py).  The GPU twin, with the real `caffe.Net` drop-in, is
tests/test_reference_driver.py.  Skipped where the reference tree is absent (the GPU box)."""
import glob
import os
import pickle
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "train_test.py")), reason="reference tree not present")
def test_unmodified_driver_dry_run_on_cpu(tmp_path):
    import cv2
    from oracle import detect as OD
    from oracle.indep_net import IndepNet
    from smallhardface_b200 import deploy
    ref = str(tmp_path / "reference")
    shutil.copytree(REF, ref, ignore=shutil.ignore_patterns("output", ".git", "caffe"))
    for dirpath, dirnames, _ in os.walk(ref):
        os.chmod(dirpath, 0o755)
    imgs = tmp_path / "imgs"
    imgs.mkdir()
    ims = [deploy.synthetic_image(30 + i, hw) for i, hw in enumerate([(40, 56), (48, 48)])]
    for i, im in enumerate(ims):
        cv2.imwrite(str(imgs / ("im%d.png" % i)), im)
    proto, model = deploy.write_synthetic_deployment(str(tmp_path / "deploy"), dilation=True)
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_driver_dryrun_boot.py"), ref, "--train", "false", "--conf",
           "configs/smallhardface.toml", "--amend", "DATA_DIR", str(imgs), "TEST.DB", "general_png", "TEST.MODEL", model,
           "TEST.GPU_ID", "[0]", "TEST.NO_CACHE", "False", "TEST.SCALES", "[100,300]"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    logs = glob.glob(os.path.join(ref, "output", "face", "general_png", "*", "stderr.log"))
    err = open(logs[0]).read() if logs else ""
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:], err[-3000:])
    assert "All Done!" in err
    out_dir = os.path.dirname(logs[0])
    dets = pickle.load(open(os.path.join(out_dir, "detections.pkl"), "rb"))
    assert len(dets) == 2 and len(dets[1]) == 2
    # the same images through the oracle's own restatement of detect(): the reference's driver must agree with it
    paths = []
    for root, _, files in os.walk(str(imgs)):
        paths += [os.path.join(root, f) for f in files if f.endswith(".png")]
    onet = IndepNet(proto, model, engine="torch")
    for i, p in enumerate(paths):
        want = OD.detect(onet, cv2.imread(p), scales=(100, 300))
        got = np.asarray(dets[1][i])
        assert got.shape == want.shape and np.abs(got - want).max() < 1e-4, p
        txt = glob.glob(os.path.join(out_dir, "**", os.path.basename(p).replace("png", "txt")), recursive=True)
        assert len(txt) == 1 and int(open(txt[0]).read().splitlines()[1]) == len(got)
