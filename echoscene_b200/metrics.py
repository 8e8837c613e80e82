"""Scene-graph constraint metrics on the GPU: ``validate_constrains`` / ``validate_constrains_changes`` with the
reference's signatures (helpers/metrics_3dfront.py:57, :181), as scripts/eval_3dfront.py calls them (:209-210, :305).  The
reference walks the triples in Python and copies two boxes device->host per triple; here one kernel evaluates every triple and one
small copy brings the 0/1 flags back, which are appended to the ``accuracy`` lists in triple order exactly as the reference
appends them.  No CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import torch

from . import _lib
from ._lib import EchoError

RELATIONS = {"left": 0, "right": 1, "front": 2, "behind": 3, "bigger than": 4, "smaller than": 5, "taller than": 6, "shorter than": 7,
             "standing on": 8, "close by": 9, "symmetrical to": 10}
# the keys the reference appends to (metrics_3dfront.py:75-176)
ACCURACY_KEY = ["left", "right", "front", "behind", "bigger", "smaller", "taller", "shorter", "standing on", "close by", "symmetrical to"]


def relation_codes(vocab: dict) -> List[int]:
    """vocab["pred_idx_to_name"][p][:-1] (the names carry a trailing newline in the reference's vocab files) -> relation code or -1"""
    return [RELATIONS.get(name[:-1], -1) for name in vocab["pred_idx_to_name"]]


def _validate(triples, pred_boxes, keep, vocab, accuracy, strict, overlap_threshold, changes_mode):
    _lib.require_cuda(triples, pred_boxes)
    if triples.dim() != 2 or triples.shape[1] != 3 or triples.dtype != torch.int64:
        raise EchoError(f"triples must be (T,3) int64, got {tuple(triples.shape)} {triples.dtype}")
    if pred_boxes.dim() != 2 or pred_boxes.shape[1] not in (6, 7):
        raise EchoError(f"pred_boxes must be (N,6) or (N,7), got {tuple(pred_boxes.shape)}")
    tri, boxes = triples.contiguous(), pred_boxes.detach().float().contiguous()
    keep_t = None
    if keep is not None:
        keep_t = torch.as_tensor(keep, device=boxes.device).reshape(-1).to(torch.int32).contiguous()
        if keep_t.numel() != boxes.shape[0]:
            raise EchoError(f"keep has {keep_t.numel()} entries for {boxes.shape[0]} boxes")
    codes = relation_codes(vocab)
    rel_host = (C.c_int32 * len(codes))(*codes)
    T = tri.shape[0]
    out_rel = torch.empty(T, dtype=torch.int8, device=boxes.device)
    out_ok = torch.empty(T, dtype=torch.int8, device=boxes.device)
    _lib.check(_lib.lib().echo_metrics_validate_constraints(_lib.ptr(tri), T, _lib.ptr(boxes), boxes.shape[0], boxes.shape[1], _lib.ptr(keep_t),
                                                            int(changes_mode), rel_host, len(codes), int(bool(strict)), float(overlap_threshold),
                                                            _lib.ptr(out_rel), _lib.ptr(out_ok), _lib.stream_ptr()))
    rel, ok = out_rel.cpu().tolist(), out_ok.cpu().tolist()          # the only device->host copy
    for r, k in zip(rel, ok):
        if r >= 0:
            accuracy[ACCURACY_KEY[r]].append(k)
            accuracy["total"].append(k)
    return accuracy


def validate_constrains(triples, pred_boxes, pred_angles, keep, vocab, accuracy: Dict[str, list], strict=True, overlap_threshold=0.3):
    """helpers/metrics_3dfront.py:57-178: triples whose two nodes are both kept (or all triples when ``keep`` is None).
    ``pred_angles`` is accepted and unused, as in the reference."""
    return _validate(triples, pred_boxes, keep, vocab, accuracy, strict, overlap_threshold, changes_mode=False)


def validate_constrains_changes(triples, pred_boxes, pred_angles, keep, vocab, accuracy: Dict[str, list], strict=True, overlap_threshold=0.3):
    """helpers/metrics_3dfront.py:181-306: triples with at least one changed node (or all triples when ``keep`` is None)."""
    if pred_boxes.dim() == 2 and pred_boxes.shape[1] != 6:
        # the reference calls box3d_iou without param6 here (:199), which unpacks exactly six values (:313) and raises otherwise
        raise EchoError("validate_constrains_changes takes six-parameter boxes (helpers/metrics_3dfront.py:199, :313)")
    return _validate(triples, pred_boxes, keep, vocab, accuracy, strict, overlap_threshold, changes_mode=True)


def new_accuracy() -> Dict[str, list]:
    """the dict scripts/eval_3dfront.py:241-243 starts from"""
    return {k: [] for k in ACCURACY_KEY + ["total"]}
