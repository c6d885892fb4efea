"""Runs the reference's OWN entry point, `train_test.py`, byte-identical, on this framework's drop-in modules.

    python tools/run_reference_driver.py --reference-root <writable copy of the reference tree> -- \
        --train false --conf configs/smallhardface.toml --amend DATA_DIR <dir of .png> TEST.DB general_png \
        TEST.MODEL <caffemodel> TEST.GPU_ID "[0]"

What happens: the working directory becomes the reference root (train_test.py:5-8 and lib/utils/get_config.py:25 use
relative paths), `smallhardface_b200.compat.install(reference_root)` puts `caffe`, `nms.*`, `utils.cython_bbox` in place
and registers the py2 -> py3 import hook for the reference's `lib/` (the files on disk are NOT modified: the hook
transforms the source text at import time), and `train_test.py` is executed as `__main__` through
the same transform.  Everything below `test_net` is then the reference's code: `lib/test.py:290-356` (test_net), `:220-267`
(inference_worker), `:109-178` (detect), `:21-106` (forward_net), `lib/datasets/general.py` (imdb + detection writer),
`lib/prototxt/manipulate.py:63-86,166-188` (deploy prototxt + dim_red splice) -- calling `caffe.Net.forward` of this
repo.  The reference tree must be a writable copy: `get_output_dir` writes under `<root>/output/`.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference-root", required=True)
    ap.add_argument("rest", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    ref = os.path.realpath(args.reference_root)
    rest = args.rest[1:] if args.rest[:1] == ["--"] else args.rest
    os.chdir(ref)
    from smallhardface_b200 import compat
    from smallhardface_b200.compat import py2hook
    compat.install(reference_root=ref)
    script = os.path.join(ref, "train_test.py")
    with open(script) as f:
        src = f.read()
    code = py2hook.compile_py2(src, script)
    sys.argv = [script] + rest
    glb = {"__name__": "__main__", "__file__": script, "__builtins__": __builtins__}
    exec(code, glb)


if __name__ == "__main__":
    main()
