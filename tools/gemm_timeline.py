"""Timeline of CTA 0 of one tcgen05 contraction inside a real shape step (echo_debug_probe_timeline): per tile, when the
TMA producer started it, when the MMA warp saw its first stage, when it committed, when the epilogue got it / finished.
Usage: python tools/gemm_timeline.py ROWS CIN COUT K   (e.g. 16384 448 3584 1 for the GEGLU projection)"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from echoscene_b200 import _lib, synth  # noqa: E402

rows, cin, cout, k = [int(v) for v in sys.argv[1:5]]
dev = torch.device("cuda:0")
m, _ = bench.build_model("bf16", dev)
g = synth.make_scene_graph(16, 64, 2)
tri = g.triples.to(dev)
uc, x = synth.shape_inputs(16, 2, same_noise=True)
uc, x = uc.to(dev), x.to(dev)
for i in range(3):
    x = m.ddim_step(x, uc, tri, 99 - i)
torch.cuda.synchronize()
buf = torch.zeros(8001, dtype=torch.int64, device=dev)
L = _lib.lib()
L.echo_debug_probe_begin(rows, cin, cout, k)
L.echo_debug_probe_timeline(buf.data_ptr())
x = m.ddim_step(x, uc, tri, 96)
torch.cuda.synchronize()
avg = ctypes.c_double(0.0)
n = L.echo_debug_probe_end(ctypes.byref(avg))
b = buf.cpu().tolist()
cnt = b[0]
ev = sorted(((b[2 + 2 * i], b[1 + 2 * i] >> 32, b[1 + 2 * i] & 0xFFFFFFFF) for i in range(min(cnt, 4000))))
t0 = ev[0][0]
names = {1: "setup done", 2: "TMA starts tile", 3: "MMA first stage", 4: "MMA issued all", 5: "epilogue gets tile", 6: "epilogue done", 7: "kernel end"}
print(f"{n} launches of this shape in the step, avg {avg.value * 1e3:.1f} us; timeline of CTA 0 of the first one ({cnt} marks):")
for t, tag, tile in ev:
    print(f"  {(t - t0) / 1e3:8.2f} us  {names.get(tag, tag):20s} tile {tile}")
