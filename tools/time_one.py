import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv_weights
impl = int(sys.argv[1]); cin, cout, H, W, k, dil = [int(a) for a in sys.argv[2:8]]
dev = torch.device("cuda:0")
L.call("shf_set_conv_impl", impl)
x = H2(torch.randn((2, 1, H, W, cin), device=dev).abs().half())
packed, kexp = pack_conv_weights((np.random.RandomState(0).randn(cout, cin, k, k) * 0.02).astype(np.float32))
wd = torch.from_numpy(packed).to(dev); bd = torch.zeros(cout, device=dev); out = H2.empty(1, H, W, cout, dev)
run = lambda: L.call("shf_conv_igemm", _ptr(x.t), _ptr(wd), _ptr(bd), _ptr(out.t), 1, H, W, cin, cout, k, dil, cout, 0, float(2.0 ** -kexp), 1, 0, 0, None, _stream())
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 2.0 * cin * cout * k * k * H * W
print("impl %d %s env[%s]  %.3f ms  %.1f TF alg (%.1f exec)" % (impl, sys.argv[2:8], " ".join("%s=%s" % (k_, v) for k_, v in os.environ.items() if k_.startswith("SHF_PROBE")), ms, fl / ms / 1e9, 3 * fl / ms / 1e9))
