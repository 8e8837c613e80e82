"""GPU: echoscene_b200.metrics (csrc/metrics.cu, one thread per triple) against the accuracy lists of the REFERENCE's own
validate_constrains / validate_constrains_changes (tests/golden/metrics.pt): identical 0/1 lists, in the reference's order."""
import pytest
import torch

from echoscene_b200 import metrics
from echoscene_b200._lib import EchoError
from oracle import gen_golden_metrics as gg
from util import gold

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_constraint_metrics_match_the_reference_lists():
    G = gold("metrics.pt")
    vocab = {"pred_idx_to_name": G["pred_names"]}
    n = 0
    for c in G["cases"]:
        boxes, triples, keep = gg.make_case(c["seed"], box_dim=c["box_dim"])
        fn = metrics.validate_constrains_changes if c["changes"] else metrics.validate_constrains
        acc = fn(triples.to(DEV), boxes.to(DEV), None, keep.to(DEV) if c["use_keep"] else None, vocab, metrics.new_accuracy())
        for k, v in c["accuracy"].items():
            assert acc[k] == v, (c["seed"], c["box_dim"], c["use_keep"], c["changes"], k)
        n += len(acc["total"])
    assert n > 2000


def test_constraint_metrics_accumulate_and_refuse_bad_input():
    G = gold("metrics.pt")
    vocab = {"pred_idx_to_name": G["pred_names"]}
    boxes, triples, keep = gg.make_case(1)
    acc = metrics.new_accuracy()
    metrics.validate_constrains(triples.to(DEV), boxes.to(DEV), None, None, vocab, acc)
    first = len(acc["total"])
    metrics.validate_constrains(triples.to(DEV), boxes.to(DEV), None, keep.tolist(), vocab, acc)      # keep as a python list, as eval passes it
    assert len(acc["total"]) > first
    empty = metrics.validate_constrains(triples[:0].to(DEV), boxes.to(DEV), None, None, vocab, metrics.new_accuracy())
    assert empty["total"] == []
    with pytest.raises(EchoError, match="CUDA"):
        metrics.validate_constrains(triples, boxes, None, None, vocab, metrics.new_accuracy())
    with pytest.raises(EchoError, match="six"):
        metrics.validate_constrains_changes(triples.to(DEV), torch.zeros(48, 7, device=DEV), None, None, vocab, metrics.new_accuracy())
    with pytest.raises(EchoError):
        metrics.validate_constrains(triples.to(DEV), boxes[:, :5].to(DEV), None, None, vocab, metrics.new_accuracy())
