"""TEST INFRASTRUCTURE -- CPU restatement of the reference's constraint metrics (helpers/metrics_3dfront.py:57-306, with
corners_from_box :308-328, box3d_iou :331-364, polygon_clip :390-434, close_dis :10-15, cal_l2_distance :17-18).  Pinned against the
reference's own functions by oracle/gen_golden_metrics.py (tests/golden/metrics.pt).  Only tests import this.

The restatement is table-driven (one predicate per relation) and keeps the reference's numeric types: entries of a float32 box are
numpy float32 scalars, so differences / products / quotients round to float32 before they meet a threshold; the box geometry
(corners, bird's-eye clip, volumes, corner distances) runs in float64 on float32-rounded half extents, as np.dot with a float64
identity makes it in the reference."""
from __future__ import annotations


import numpy as np

KEYS = ["left", "right", "front", "behind", "bigger", "smaller", "taller", "shorter", "standing on", "close by", "symmetrical to"]
NAMES = ["left", "right", "front", "behind", "bigger than", "smaller than", "taller than", "shorter than", "standing on", "close by",
         "symmetrical to"]
f32 = np.float32


def _corners(b):
    hw, hl = b[2] / f32(2), b[0] / f32(2)
    xs = np.array([hw, hw, -hw, -hw, hw, hw, -hw, -hw], dtype=np.float64) + np.float64(b[3])
    ys = np.array([b[1], b[1], b[1], b[1], 0, 0, 0, 0], dtype=np.float64) + np.float64(b[4])
    zs = np.array([hl, -hl, -hl, hl, hl, -hl, -hl, hl], dtype=np.float64) + np.float64(b[5])
    return np.stack([xs, ys, zs], axis=1)


def _clip_area(subject, clip):
    """Sutherland-Hodgman with the reference's strict inside test; area of the clipped polygon (0.0 when nothing is left)"""
    out = [tuple(p) for p in subject]
    a = tuple(clip[-1])
    for b in (tuple(p) for p in clip):
        src, out = out, []
        inside = lambda p: (b[0] - a[0]) * (p[1] - a[1]) > (b[1] - a[1]) * (p[0] - a[0])   # noqa: E731
        s = src[-1]
        for e in src:
            if inside(e) != inside(s):
                dc, dp = (a[0] - b[0], a[1] - b[1]), (s[0] - e[0], s[1] - e[1])
                n1, n2 = a[0] * b[1] - a[1] * b[0], s[0] * e[1] - s[1] * e[0]
                n3 = 1.0 / (dc[0] * dp[1] - dc[1] * dp[0])
                out.append(((n1 * dp[0] - n2 * dc[0]) * n3, (n1 * dp[1] - n2 * dc[1]) * n3))
            if inside(e):
                out.append(e)
            s = e
        a = b
        if not out:
            return 0.0
    area = sum(out[i][0] * out[(i + 1) % len(out)][1] - out[(i + 1) % len(out)][0] * out[i][1] for i in range(len(out)))
    return 0.5 * abs(area)


def iou3d(b1, b2):
    c1, c2 = _corners(b1), _corners(b2)
    inter = _clip_area([(c1[i, 2], c1[i, 0]) for i in range(4)], [(c2[i, 2], c2[i, 0]) for i in range(4)])
    height = max(0.0, min(c1[0, 1], c2[0, 1]) - max(c1[4, 1], c2[4, 1]))
    vol = lambda c: np.linalg.norm(c[0] - c[1]) * np.linalg.norm(c[1] - c[2]) * np.linalg.norm(c[0] - c[4])   # noqa: E731
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(inter * height) / min(vol(c1), vol(c2))


def close_dis(b1, b2):
    c1, c2 = _corners(b1), _corners(b2)
    with np.errstate(invalid="ignore"):
        d = -2.0 * (c1 @ c2.T) + (c1 ** 2).sum(-1)[:, None] + (c2 ** 2).sum(-1)[None, :]
        return np.min(np.sqrt(d))


def _l2(ax, ay, bx, by):
    return np.sqrt((bx - ax) ** 2 + (by - ay) ** 2)          # float32 throughout


def holds(rel: int, s, o, strict=True, thr=0.3) -> bool:
    """does relation `rel` (index into NAMES) hold for subject box s / object box o (float32 arrays)?"""
    with np.errstate(divide="ignore", invalid="ignore"):
        if rel < 4:
            gap = [s[5] - o[5] > f32(-0.05), s[5] - o[5] < f32(0.05), s[3] - o[3] < f32(-0.05), s[3] - o[3] > f32(0.05)][rel]
            return not (bool(gap) or (strict and iou3d(s, o) > thr))
        if rel in (4, 5):
            sv, ov = s[0] * s[1] * s[2], o[0] * o[1] * o[2]
            q = (sv - ov) / sv
            return not bool(q < f32(0.15)) if rel == 4 else not bool(q > f32(-0.15))
        if rel in (6, 7):
            hs, ho = s[4] + s[1], o[4] + o[1]
            q = (hs - ho) / hs
            return not bool(q < f32(0.1)) if rel == 6 else not bool(q > f32(-0.1))
        if rel == 8:
            return bool(np.abs(s[4] - o[4]) < f32(0.04))
        if rel == 9:
            return not bool(close_dis(s, o) > 0.45)
        return bool(_l2(-s[3], -s[5], o[3], o[5]) < f32(0.45) or _l2(-s[3], s[5], o[3], o[5]) < f32(0.45)
                    or _l2(s[3], -s[5], o[3], o[5]) < f32(0.45))


def validate(triples, boxes, keep, pred_names, changes=False, strict=True, thr=0.3):
    """-> accuracy dict like the reference's, from numpy inputs: triples (T,3) int, boxes (N,6|7) float32, keep (N,) or None"""
    acc = {k: [] for k in KEYS + ["total"]}
    boxes = np.asarray(boxes, dtype=np.float32)
    for s, p, o in np.asarray(triples).tolist():
        if keep is not None:
            if changes and not (keep[s] == 0 or keep[o] == 0):
                continue
            if not changes and not (keep[s] == 1 and keep[o] == 1):
                continue
        name = pred_names[p][:-1]
        if name not in NAMES:
            continue
        r = NAMES.index(name)
        ok = int(holds(r, boxes[s], boxes[o], strict, thr))
        acc[KEYS[r]].append(ok)
        acc["total"].append(ok)
    return acc
