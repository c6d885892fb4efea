"""One tcgen05 conv shape, three launches (for `ncu -k regex:conv_stream -s 2 -c 1`):
python tools/ncu_one.py <impl> <cin> <cout> <H> <W> <k> <dil> [fmt]      fmt 0 = h2 (split fp16), 1 = hf8 (fp16 + fp8)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smallhardface_b200 import lib as L
from smallhardface_b200.engine import H2, _ptr, _stream, pack_conv_weights, pack_conv_weights_hf8
impl = int(sys.argv[1]); cin, cout, H, W, k, dil = [int(a) for a in sys.argv[2:8]]
fmt = int(sys.argv[8]) if len(sys.argv) > 8 else 0
dev = torch.device("cuda:0")
L.call("shf_set_conv_impl", impl)
x = H2.from_nchw(torch.randn((1, cin, H, W), device=dev).abs(), fmt=fmt)
w = (np.random.RandomState(0).randn(cout, cin, k, k) * 0.02).astype(np.float32)
packed, kexp = (pack_conv_weights_hf8 if fmt else pack_conv_weights)(w)
wd = torch.from_numpy(packed).to(dev); bd = torch.zeros(cout, device=dev); out = H2.empty(1, H, W, cout, dev, fmt)
for _ in range(3):
    L.call("shf_conv_igemm", _ptr(x.t), _ptr(wd), _ptr(bd), _ptr(out.t), 1, H, W, cin, cout, k, dil, cout, 0, float(2.0 ** -kexp), 1, fmt, fmt, None, _stream())
torch.cuda.synchronize()
