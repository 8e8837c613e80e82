"""The dominant kernel alone (for `ncu --set full`): tcgen05 implicit-GEMM conv 224@16^3 -> 224, N=16 (7 per shape step)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoscene_b200 import _lib  # noqa: E402

n, c = 16, 224
dev = torch.device("cuda:0")
x = torch.randn(n, 16, 16, 16, c, device=dev)
w = torch.randn(c, c, 3, 3, 3, device=dev) * 0.01
b = torch.zeros(c, device=dev)
out = torch.empty(n, 16, 16, 16, c, device=dev)
L = _lib.lib()
for i in range(4):
    _lib.check(L.echo_op_conv3d(x.data_ptr(), n, 16, 16, 16, c, w.data_ptr(), b.data_ptr(), c, 3, 1, 1, out.data_ptr(),
                                _lib.PREC_BF16, _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
