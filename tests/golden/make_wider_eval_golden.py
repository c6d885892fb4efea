"""Golden AP / PR curves for the WIDER evaluator: the REFERENCE's lib/wider_eval_tools/wider_eval.py executed in this
container on the synthetic ground truth of tests/wider_synth.py.  The file is Python 2; it is run with exactly these
substitutions (nothing else is touched): `xrange` -> `range`; `map(...)` wrapped in `list(...)` (:28, the row assignment
needs a sequence); `reduce` taken from functools (:146); and `round` bound to Python 2's semantics (C round, halves away
from zero, :92 -- Python 3's banker's rounding would change which detections match at IoU == 0.5).

    python tests/golden/make_wider_eval_golden.py        # writes tests/golden/wider_eval.npz
"""
import math
import os
import sys
import tempfile
import types
from functools import reduce

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
REF = os.environ.get("SHF_REFERENCE_ROOT", "/root/reference")


def py2_round(x):
    return float(math.floor(x + 0.5)) if (x - math.floor(x)) != 0.5 else float(math.floor(x) + 1.0)


def load_reference_evaluator():
    path = REF + "/lib/wider_eval_tools/wider_eval.py"
    src = open(path).read().replace("xrange(", "range(").replace("raw_info = map(lambda x: float(x), tmp[k + 2].split())",
                                                                  "raw_info = list(map(lambda x: float(x), tmp[k + 2].split()))")
    mod = types.ModuleType("ref_wider_eval")
    mod.__dict__.update(reduce=reduce, round=lambda v: py2_round(float(v)))
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


def main():
    import wider_synth
    ref = load_reference_evaluator()
    out = {}
    for seed in (0, 1):
        root = tempfile.mkdtemp()
        pred_dir, gt_dir = wider_synth.make(root, seed=seed)
        for bug in (True, False):
            for thr in (0.5, 0.3):
                with np.errstate(all="ignore"):
                    ap, pr = ref.wider_eval(pred_dir, gt_dir, parallel=False, mimic_eval_bug=bug, IoU_thresh=thr)
                key = "s%d_bug%d_t%d" % (seed, int(bug), int(thr * 10))
                out[key + "_ap"] = np.array(ap, dtype=np.float64)
                out[key + "_pr"] = np.stack(pr)
                print(key, ap)
    np.savez_compressed(os.path.join(HERE, "wider_eval.npz"), **out)


if __name__ == "__main__":
    main()
