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
for _ in range(3):
    L.call("shf_conv_igemm", _ptr(x.t), _ptr(wd), _ptr(bd), _ptr(out.t), 1, H, W, cin, cout, k, dil, cout, 0, float(2.0 ** -kexp), 1, 0, 0, _stream())
torch.cuda.synchronize()
