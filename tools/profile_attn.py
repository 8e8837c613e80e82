"""The attention operator alone (for ncu): n objects x tokens x heads x dh through echo_op_attention (bf16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoscene_b200 import _lib  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
n, tokens, heads, dh = a if len(a) == 4 else (16, 1024, 8, 56)
qkv = torch.randn(n * tokens, 3 * heads * dh, device="cuda")
out = torch.empty(n * tokens, heads * dh, device="cuda")
for _ in range(3):
    _lib.check(_lib.lib().echo_op_attention(qkv.data_ptr(), n, tokens, heads, dh, out.data_ptr(), _lib.PREC_BF16, _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
