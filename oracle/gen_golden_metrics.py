"""Pins oracle/metrics_oracle.py against the reference's own helpers/metrics_3dfront.py (imported in place, build container only)
and writes tests/golden/metrics.pt: seeded boxes / triples / keep vectors and the accuracy lists the reference's
validate_constrains and validate_constrains_changes produce for them.  Usage: python oracle/gen_golden_metrics.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import metrics_oracle as mo, ref_import  # noqa: E402

PRED_NAMES = ["__in_room__\n", "left\n", "right\n", "front\n", "behind\n", "close by\n", "symmetrical to\n", "bigger than\n", "smaller than\n",
              "taller than\n", "shorter than\n", "standing on\n", "above\n", "same style as\n", "same material as\n", "same category as\n"]


def make_case(seed, n=48, t=600, box_dim=6):
    g = torch.Generator().manual_seed(seed)
    size = torch.rand(n, 3, generator=g) * 1.5 + 0.2
    pos = (torch.rand(n, 3, generator=g) - 0.5) * torch.tensor([4.0, 0.6, 4.0])
    boxes = torch.cat([size, pos], dim=1)
    # near-threshold and degenerate situations: stacked, touching, nested, identical and mirrored boxes
    boxes[1] = boxes[0]
    boxes[2, 3:] = boxes[0, 3:] + torch.tensor([0.0, 0.02, 0.0])
    boxes[3, 3:] = torch.tensor([-1.0, 1.0, -1.0]) * boxes[0, 3:]
    boxes[4, :3] = boxes[0, :3] * 0.5
    boxes[4, 3:] = boxes[0, 3:]
    boxes[5, 3] = boxes[0, 3] + (boxes[0, 2] + boxes[5, 2]) / 2
    if box_dim == 7:
        boxes = torch.cat([boxes, torch.rand(n, 1, generator=g) * 6.28], dim=1)
    s = torch.randint(0, n, (t,), generator=g)
    o = torch.randint(0, n, (t,), generator=g)
    p = torch.randint(0, len(PRED_NAMES), (t,), generator=g)
    s[:12], o[:12] = torch.tensor([0, 0, 0, 0, 0, 1, 2, 3, 4, 5, 0, 1]), torch.tensor([1, 2, 3, 4, 5, 0, 0, 0, 0, 0, 0, 1])
    triples = torch.stack([s, p, o], dim=1)
    keep = torch.randint(0, 2, (n,), generator=g)
    return boxes, triples, keep


def main():
    ref_import.install_stubs()
    sys.path.insert(0, ref_import.REF_ROOT)
    ref = importlib.import_module("helpers.metrics_3dfront")
    vocab = {"pred_idx_to_name": PRED_NAMES}
    out = {"pred_names": PRED_NAMES, "cases": []}
    worst = 0
    for seed, dim in ((1, 6), (2, 7), (3, 6)):
        boxes, triples, keep = make_case(seed, box_dim=dim)
        for keep_arg in (None, keep):
            for changes in ((False, True) if dim == 6 else (False,)):   # the _changes variant unpacks six values (metrics_3dfront.py:313 via :199)
                fn = ref.validate_constrains_changes if changes else ref.validate_constrains
                acc = {k: [] for k in mo.KEYS + ["total"]}
                np.seterr(all="ignore")
                fn(triples, boxes, None, keep_arg, vocab, acc)
                mine = mo.validate(triples.numpy(), boxes.numpy(), None if keep_arg is None else keep_arg.numpy(), PRED_NAMES, changes)
                for k in acc:
                    worst += int(acc[k] != mine[k])
                    assert acc[k] == mine[k], (seed, dim, changes, k)
                out["cases"].append({"seed": seed, "box_dim": dim, "use_keep": keep_arg is not None, "changes": changes,
                                     "accuracy": {k: list(map(int, v)) for k, v in acc.items()}})
    torch.save(out, os.path.join(ROOT, "tests", "golden", "metrics.pt"))
    n = sum(len(c["accuracy"]["total"]) for c in out["cases"])
    print(f"metrics oracle pinned to the reference on {len(out['cases'])} cases / {n} evaluated triples: {worst} mismatching lists")


if __name__ == "__main__":
    main()
