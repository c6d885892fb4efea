"""Golden files for the Python 3 configuration / prototxt plumbing (smallhardface_b200/config.py, prototxt.py):
outputs of the REFERENCE's own `lib/utils/get_config.py` and `lib/prototxt/manipulate.py`, imported unmodified through
the py2 hook (smallhardface_b200.compat) in this container -- `str(NetParameter)` is the real protobuf runtime's text
format.  Usage (needs /root/reference; writes tests/golden/config/):

    python tests/golden/make_config_golden.py
"""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SHF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "config")

AMEND = ["DATA_DIR", "/data/images", "TEST.DB", "general_png", "TEST.MODEL", "/data/final.caffemodel", "TEST.GPU_ID", "[0]",
         "TEST.SCALES", "[300, 600]", "NAME", "golden"]


def main():
    sys.path.insert(0, ROOT)
    work = tempfile.mkdtemp()
    os.symlink(REF + "/configs", os.path.join(work, "configs"))
    os.symlink(REF + "/models", os.path.join(work, "models"))
    os.chdir(work)
    from smallhardface_b200 import compat
    compat.install(REF)
    from utils.get_config import cfg, cfg_from_file, cfg_from_list, cfg_dump, cfg_table
    os.makedirs(OUT, exist_ok=True)
    # the plain template path first (dilation switch still off)
    from lib.prototxt import manipulate
    manipulate.manipulate_test("models/test_template.prototxt", os.path.join(OUT, "test_plain.prototxt"))
    cfg_from_file("configs/smallhardface.toml")
    cfg.TEST.NO_CACHE = True
    cfg_from_list(AMEND)
    cfg.LOG.CMD = "golden"
    cfg.LOG.TIME = "2026_01_01_00_00_00"
    cfg.ROOT_DIR = "/root/reference"
    with open(os.path.join(OUT, "cfgs_test.toml"), "w") as f:
        cfg_dump({i: cfg[i] for i in cfg if i != "TRAIN"}, f)
    with open(os.path.join(OUT, "cfg_table.md"), "w") as f:
        f.write(cfg_table({i: cfg[i] for i in cfg if i != "TRAIN"}))
    manipulate.manipulate_test("models/test_template.prototxt", os.path.join(OUT, "test_dilation.prototxt"))
    for f in os.listdir(OUT):
        if f.endswith(".jpg"):
            os.remove(os.path.join(OUT, f))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
