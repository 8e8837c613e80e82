"""TEST INFRASTRUCTURE -- the oracle of the GraphTripleConvNet backward pass.

The reference has no explicit backward: `loss.backward()` (scripts/train_3dfront.py:247) differentiates model/graph.py:124-211 and
model/layers.py:21-38 with torch autograd.  The oracle does the same over its OWN forward restatement
(`echoscene_oracle.graph_triple_conv_net(batch_stats=True)`, plain torch fp32 on the CPU), so the CUDA backward is checked against
derivatives of the restated algorithm, and the restatement is pinned against the reference's autograd by
`oracle/gen_golden_gcn_bwd.py` (tests/golden/gcn_bwd.pt).  Only tests/ may import this."""
from typing import Dict, Optional, Tuple

import torch

from . import echoscene_oracle as orc

Tensor = torch.Tensor


def graph_triple_conv_net_backward(sd: Dict[str, Tensor], obj: Tensor, pred: Tensor, edges: Tensor, d_obj_out: Tensor,
                                   d_pred_out: Optional[Tensor], num_layers: int
                                   ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Dict[str, Tensor], Dict[str, Tensor]]:
    """-> (obj_out, pred_out, d_obj, d_pred, {parameter name: gradient}, {buffer name: value after the forward}).
    The cotangent of an output that does not reach the loss is None (= zeros)."""
    leaf = {}
    for k, v in sd.items():
        t = v.detach().clone()
        if t.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            t.requires_grad_(True)
        leaf[k] = t
    track: Dict[str, Tensor] = {}
    leaf["__running_update__"] = track
    o = obj.detach().clone().requires_grad_(True)
    p = pred.detach().clone().requires_grad_(True)
    obj_out, pred_out = orc.graph_triple_conv_net(leaf, "", o, p, edges, num_layers=num_layers, batch_stats=True)
    loss = (obj_out * d_obj_out).sum()
    if d_pred_out is not None:
        loss = loss + (pred_out * d_pred_out).sum()
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaf.items()
             if torch.is_tensor(v) and v.requires_grad}
    return obj_out.detach(), pred_out.detach(), o.grad, p.grad, grads, track
