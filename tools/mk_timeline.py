"""Reads the per-CTA per-stage clock dump of the persistent layout executor (ECHO_MK_TIMELINE=<file>, csrc/layout.cu) and
prints, per stage, where the time goes: barrier wait, foreground compute, background ops (SM clocks, max / median over CTAs).
Usage: ECHO_MK_TIMELINE=gpurun_out/mk_timeline.bin python tools/profile_step.py --branch layout; python tools/mk_timeline.py gpurun_out/mk_timeline.bin"""
import sys

import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.int64)
G, S = int(raw[0]), int(raw[1])
t = raw[2:2 + G * S * 12].reshape(G, S, 12).astype(np.float64)
meta = raw[2 + G * S * 12:].reshape(S, 8)
wait = t[:, :, 1] - t[:, :, 0]
fg = t[:, :, 2] - t[:, :, 1]
bg = (raw[2:2 + G * S * 12].reshape(G, S, 12)[:, :, 3] & 0xffffffff).astype(np.float64)
bg2 = (raw[2:2 + G * S * 12].reshape(G, S, 12)[:, :, 3] >> 32).astype(np.float64)   # (slot 3 now holds the probed latency of one dependent L2 load)
stage = np.empty((G, S))
stage[:, :-1] = t[:, 1:, 0] - t[:, :-1, 0]
stage[:, -1] = t[:, -1, 2] - t[:, -1, 0]
print(f"{G} CTAs, {S} stages; clocks per stage (CTA-local SM clock)")
busy = t[:, :, 4] > 0   # CTAs that ran a 16-row unit in the stage
def med(x, s):
    v = x[:, s][busy[:, s]]
    return float(np.median(v)) if v.size else 0.0
ahead, feedt = t[:, :, 4] - t[:, :, 8], t[:, :, 10] - t[:, :, 9]
wland, stg, mma, epi = t[:, :, 4] - t[:, :, 1], t[:, :, 5] - t[:, :, 4], t[:, :, 6] - t[:, :, 5], t[:, :, 7] - t[:, :, 6]
print(f"{'st':>3s} {'nA':>2s} {'nB':>2s} {'type':>4s} {'M':>4s} {'K':>5s} {'nout':>5s} {'pro':>3s} {'units':>5s} | {'stage med':>9s} | {'wait min':>8s} {'med':>6s} {'max':>6s} | {'fg med':>6s} {'max':>6s} | {'L2 ld':>6s} {'2nd':>6s} | unit (med): {'w-land':>6s} {'stage':>6s} {'mma':>6s} {'epi':>6s} | {'issued-before-land':>18s} {'feed call':>9s}")
for s in range(S):
    m = meta[s]
    print(f"{s:3d} {m[0]:2d} {m[1]:2d} {m[2]:4d} {m[3]:4d} {m[4]:5d} {m[5]:5d} {m[6]:3d} {m[7]:5d} | {np.median(stage[:, s]):9.0f} | {wait[:, s].min():8.0f} {np.median(wait[:, s]):6.0f} {wait[:, s].max():6.0f} | "
          f"{np.median(fg[:, s]):6.0f} {fg[:, s].max():6.0f} | {np.median(bg[:, s]):6.0f} {np.median(bg2[:, s]):6.0f} |             {med(wland, s):6.0f} {med(stg, s):6.0f} {med(mma, s):6.0f} {med(epi, s):6.0f} | {med(ahead, s):18.0f} {float(np.median(feedt[:, s])):9.0f}")
tot = np.median(stage, axis=0).sum()
print(f"sum of per-stage medians: {tot:.0f} clocks; min-wait sum {wait.min(axis=0).sum():.0f}; max-fg sum {fg.max(axis=0).sum():.0f}; max-bg sum {bg.max(axis=0).sum():.0f}")
