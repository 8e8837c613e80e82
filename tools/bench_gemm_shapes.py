"""Per-shape A/B of the tcgen05 GEMM modes (1 = single CTA tiles, 2 = CTA pairs) on the shape UNet's contraction list."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from echoscene_b200 import _lib  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda:0")
N = 16
shapes = [  # (cin, cout, d, h, w, k)
    (224, 224, 16, 16, 16, 3), (448, 448, 16, 8, 8, 3), (672, 672, 16, 4, 4, 3), (448, 448, 16, 16, 16, 3),
    (448, 224, 16, 16, 16, 3), (672, 224, 16, 16, 16, 3), (1120, 448, 16, 8, 8, 3), (1344, 672, 16, 4, 4, 3),
    (672, 672, 16, 8, 8, 3), (896, 448, 16, 8, 8, 3), (448, 3584, 16, 8, 8, 1), (1792, 448, 16, 8, 8, 1),
    (672, 5376, 16, 4, 4, 1), (448, 448, 16, 8, 8, 1), (672, 672, 16, 4, 4, 1), (448, 1536, 16, 8, 8, 1),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (cin, cout, d, h, w, k) in shapes:
    x = torch.randn(N, d, h, w, cin, device=dev)
    wt = torch.randn(cout, cin, k, k, k, device=dev) * 0.01
    b = torch.zeros(cout, device=dev)
    out = torch.empty(N, d, h, w, cout, device=dev)
    res = {}
    for mode in (1, 2):
        L.echo_debug_set_tc_mode(mode)
        def call():
            _lib.check(L.echo_op_conv3d(x.data_ptr(), N, d, h, w, cin, wt.data_ptr(), b.data_ptr(), cout, k, 1, 1, out.data_ptr(),
                                        _lib.PREC_BF16, _lib.stream_ptr()))
        for _ in range(2):
            call()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); call(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res[mode] = min(ts)
    # subtract the fp32<->bf16 conversions of the op wrapper (same in both modes): report raw and the ratio
    fl = 2.0 * N * d * h * w * cin * cout * k ** 3
    print(f"{cin:5d}->{cout:5d} @{d}x{h}x{w} k{k}: mode1 {res[1]*1e3:8.1f} us  mode2 {res[2]*1e3:8.1f} us  ratio {res[2]/res[1]:.3f}  ({fl/1e9:.0f} GFLOP)")
L.echo_debug_set_tc_mode(0)
