"""Bootstrap of tests/test_reference_driver_dryrun.py (TEST INFRASTRUCTURE): runs the reference's unmodified
train_test.py like tools/run_reference_driver.py does, but with `caffe.Net` replaced by a CPU stand-in backed by the
ORACLE, so that the whole host-side plumbing (py2 hook, cfg, manipulate_test, imdb, detect / forward_net, bbox_vote,
detection writer, detections.pkl) is exercised on a machine without a GPU.  Never used by the product."""
import os
import sys
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _Blob(object):
    def __init__(self):
        self.data = np.zeros((1,), np.float32)

    def reshape(self, *dims):
        if tuple(dims) != self.data.shape:
            self.data = np.zeros(dims, np.float32)


class OracleBackedNet(object):
    def __init__(self, proto, model, phase):
        from oracle.indep_net import IndepNet
        self._net = IndepNet(proto, model, engine="torch")
        self.blobs = OrderedDict((k, _Blob()) for k in ("data", "im_info", "boxes", "cls_prob"))
        self.inputs, self.outputs = ["data", "im_info"], ["boxes", "cls_prob"]

    def forward(self, **kw):
        for k, v in kw.items():
            self.blobs[k].data[...] = v
        out = self._net.forward(**{k: self.blobs[k].data for k in self.inputs})
        for k in self.outputs:
            self.blobs[k].data = np.ascontiguousarray(out[k], np.float32)
        return {k: self.blobs[k].data for k in self.outputs}


def main():
    ref = os.path.realpath(sys.argv[1])
    os.chdir(ref)
    from smallhardface_b200 import compat
    from smallhardface_b200.compat import py2hook
    compat.install(reference_root=ref)
    import caffe
    caffe.Net = OracleBackedNet
    caffe.set_mode_gpu = lambda: None
    caffe.set_device = lambda i: None
    script = os.path.join(ref, "train_test.py")
    sys.argv = [script] + sys.argv[2:]
    exec(py2hook.compile_py2(open(script).read(), script), {"__name__": "__main__", "__file__": script})


if __name__ == "__main__":
    main()
