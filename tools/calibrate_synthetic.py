"""One-off calibration of the synthetic cls/bbox weight gains frozen in smallhardface_b200/deploy.py.
Runs the CPU oracle on the 224x224 parity image and prints the (fg-bg)
logit sigma and the delta sigma so the gains can be set for sigma ~1.5 / ~0.25."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from smallhardface_b200 import deploy, caffe_proto as cp
from smallhardface_b200.graph import NetSpec
from smallhardface_b200.models import build_test_net, splice_dim_red
from oracle.net import OracleNet

for dil in (True, False):
    net = build_test_net(dil)
    if dil:
        net = splice_dim_red(net)
    spec = NetSpec(net)
    params = deploy.synthetic_params(spec)
    model = deploy.params_to_netparameter(spec, params)
    on = OracleNet(net, model, engine="torch", fast=True)
    im = np.random.RandomState(3).randint(0, 256, (224, 224, 3)).astype(np.float32) - np.array([[[102.9801, 115.9465, 122.7717]]])
    data = np.ascontiguousarray(im.transpose(2, 0, 1)[None].astype(np.float32))
    t = time.time()
    out = on.forward(data=data, im_info=np.array([[224, 224, 1.0]], np.float32))
    b = on.blobs
    for k in ["conv1_2", "conv3_3", "conv5_3", "conv4_fuse_final", "head_1" if dil else "head"]:
        print(k, "rms %.3f max %.1f" % (np.sqrt((b[k] ** 2).mean()), b[k].max()))
    cls = b["cls_score_reshape_output"]
    d = cls[0, 1] - cls[0, 0]
    print("dil", dil, "logit diff mean %.3f std %.3f" % (d.mean(), d.std()), "bbox std", b["bbox_pred_output"].std(),
          "R", out["boxes"].shape, "n>0.05", (out["cls_prob"][:, 1] > 0.05).sum(), "time %.1fs" % (time.time() - t))
