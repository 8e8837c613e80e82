"""One contraction alone (for `ncu --set full`): tcgen05 implicit-GEMM conv cin@DxHxW -> cout, N objects.
Default = the dominant kernel, 224@16^3 -> 224, N=16 (7 per shape step).  Usage: profile_conv.py [cin cout d h w k]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoscene_b200 import _lib  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
cin, cout, d, h, w, k = a if len(a) == 6 else (224, 224, 16, 16, 16, 3)
n = 16
dev = torch.device("cuda:0")
x = torch.randn(n, d, h, w, cin, device=dev)
wt = torch.randn(cout, cin, k, k, k, device=dev) * 0.01
b = torch.zeros(cout, device=dev)
out = torch.empty(n, d, h, w, cout, device=dev)
L = _lib.lib()
for i in range(4):
    _lib.check(L.echo_op_conv3d(x.data_ptr(), n, d, h, w, cin, wt.data_ptr(), b.data_ptr(), cout, k, 1, 1, out.data_ptr(),
                                _lib.PREC_BF16, _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
