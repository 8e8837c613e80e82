"""CPU: the metrics oracle (oracle/metrics_oracle.py) against the accuracy lists the REFERENCE's own validate_constrains /
validate_constrains_changes produced for the seeded cases (tests/golden/metrics.pt, written by oracle/gen_golden_metrics.py)."""
import torch

from oracle import gen_golden_metrics as gg, metrics_oracle as mo
from util import gold


def test_metrics_oracle_reproduces_the_reference_lists():
    G = gold("metrics.pt")
    assert len(G["cases"]) == 10
    evaluated = 0
    for c in G["cases"]:
        boxes, triples, keep = gg.make_case(c["seed"], box_dim=c["box_dim"])
        got = mo.validate(triples.numpy(), boxes.numpy(), keep.numpy() if c["use_keep"] else None, G["pred_names"], c["changes"])
        for k, v in c["accuracy"].items():
            assert got[k] == v, (c["seed"], c["box_dim"], c["use_keep"], c["changes"], k)
        evaluated += len(c["accuracy"]["total"])
    assert evaluated > 2000
    # every relation occurs, with both outcomes, somewhere in the fixtures
    for k in mo.KEYS:
        seen = {x for c in G["cases"] for x in c["accuracy"][k]}
        assert seen == {0, 1}, (k, seen)


def test_relation_codes_strip_the_trailing_newline():
    from echoscene_b200 import metrics
    codes = metrics.relation_codes({"pred_idx_to_name": gg.PRED_NAMES})
    assert codes[0] == -1 and codes[1] == 0 and codes[gg.PRED_NAMES.index("symmetrical to\n")] == 10 and codes[-1] == -1
    assert sorted(c for c in codes if c >= 0) == list(range(11))
