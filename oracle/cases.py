"""Seeded parity cases shared by oracle/gen_golden.py and tests/ (TEST INFRASTRUCTURE).

Every case is fully determined by integers: weights come from ``arch.make_state_dict(specs, seed)``,
graphs/inputs from ``echoscene_b200.synth``.  Fixtures under tests/golden/ hold only the *reference's
outputs* for these cases.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from echoscene_b200 import arch, synth

from echoscene_b200.synth import (WEIGHT_SEED_GCN, WEIGHT_SEED_LAYOUT, WEIGHT_SEED_SCENE, WEIGHT_SEED_SHAPE,  # noqa: E402,F401
                                  WEIGHT_SEED_VQVAE)


@dataclass
class GraphCase:
    name: str
    n_nodes: int
    n_triples: int
    seed: int


GCN_CASE = GraphCase("gcn_layout_n8", 8, 32, 1)
LAYOUT_CASE = GraphCase("layout_step_n8", 8, 32, 1)          # BASELINE config 1 graph
LAYOUT_CHAIN_STEPS = 10                                        # config 1: 10 DDPM steps
SHAPE_CASE = GraphCase("shape_step_n4", 4, 8, 4)
SHAPE_CHAIN_STEPS = 3                                          # first 3 iterations of the S=100 DDIM schedule
SHAPE_CHAIN_CASE = GraphCase("shape_chain_n3", 3, 4, 5)


def layout_cfg() -> arch.UNet1DConfig:
    return arch.UNet1DConfig()


def shape_cfg() -> arch.UNet3DConfig:
    return arch.UNet3DConfig()


def gcn_inputs(case: GraphCase, cfg: arch.GCNConfig):
    g = synth.make_scene_graph(case.n_nodes, case.n_triples, case.seed)
    gen = torch.Generator().manual_seed(case.seed + 100)
    obj = torch.randn(case.n_nodes, cfg.input_dim_obj, generator=gen)
    pred = torch.randn(case.n_triples, cfg.input_dim_pred, generator=gen)
    return g, obj, pred


def layout_step_inputs(case: GraphCase, cfg: arch.UNet1DConfig):
    g = synth.make_scene_graph(case.n_nodes, case.n_triples, case.seed)
    obj_embed, x = synth.layout_inputs(case.n_nodes, case.seed + 200, cfg.obj_embed_dim, cfg.in_channels)
    t = torch.tensor([999, 500, 3, 0, 17, 250, 731, 64][: case.n_nodes], dtype=torch.int64)
    return g, obj_embed, x, t


def layout_chain_inputs(case: GraphCase, cfg: arch.UNet1DConfig, steps: int):
    g = synth.make_scene_graph(case.n_nodes, case.n_triples, case.seed)
    obj_embed, x_T = synth.layout_inputs(case.n_nodes, case.seed + 300, cfg.obj_embed_dim, cfg.in_channels)
    gen = torch.Generator().manual_seed(case.seed + 301)
    noises = [torch.randn(case.n_nodes, cfg.in_channels, generator=gen) for _ in range(steps)]
    return g, obj_embed, x_T, noises


def shape_step_inputs(case: GraphCase, cfg: arch.UNet3DConfig, same_noise: bool = False):
    g = synth.make_scene_graph(case.n_nodes, case.n_triples, case.seed)
    uc, x = synth.shape_inputs(case.n_nodes, case.seed + 400, cfg.context_dim, same_noise=same_noise)
    t = torch.tensor([991, 501, 1, 251, 11, 741, 331, 91][: case.n_nodes], dtype=torch.int64)
    return g, uc, x, t


def vqvae_cfg() -> arch.VQVAEConfig:
    return arch.VQVAEConfig()


VQVAE_CASE_OBJECTS = 2


def vqvae_inputs(n: int = VQVAE_CASE_OBJECTS, seed: int = 6):
    """latents as the DDIM chain leaves them: (n, 3, 16, 16, 16), O(1) values"""
    return synth.vqvae_inputs(n, seed)


def scene_cfg() -> arch.SceneEncoderConfig:
    return arch.SceneEncoderConfig()


SCENE_CASE = GraphCase("scene_encode_n8", 8, 32, 7)          # BASELINE config 1 shaped graph


def scene_inputs(case: GraphCase = SCENE_CASE, cfg: "arch.SceneEncoderConfig" = None):
    """dec_objs (N,) i64 class ids (last node = '_scene_' class 0), triples, CLIP-like text / relation features"""
    return synth.scene_inputs(case.n_nodes, case.n_triples, case.seed, cfg)


# glue of sample / sample_with_changes / sample_with_additions (oracle/gen_golden_scene_glue.py)
SCENE_GLUE_NP_SEED = 123                                       # np.random stream of the change flags (EchoScene.py:437, 494)
SCENE_ENC_CASE_SAME = GraphCase("scene_enc_n8", 8, 28, 11)     # the scene before a relationship change: same nodes, other triples
SCENE_ENC_CASE_SMALL = GraphCase("scene_enc_n6", 6, 18, 12)    # the scene before two objects are added
# (case name, Sg2ScDiffModel method, replace_latent)
SCENE_GLUE_CASES = (
    ("sample", "sample", False),
    ("changes", "sample_with_changes", False),
    ("changes_replace_all", "sample_with_changes", True),
    ("additions", "sample_with_additions", False),
    ("additions_replace_all", "sample_with_additions", True),
    # the layout-only model (model/EchoLayout.py:291-400); its additions draw the change flags at nodes_added (:371)
    ("box_sample", "sampleBoxes", False),
    ("box_changes", "sampleBoxes_with_changes", False),
    ("box_additions", "sampleBoxes_with_additions", False),
)


def scene_box_cfg() -> arch.SceneEncoderConfig:
    """the layout-only model (model/EchoLayout.py): predicates of `manipulate` from pred_embeddings_man_dc"""
    return arch.SceneEncoderConfig(man_dc_preds=True)


def scene_box_state_dict():
    """state_dict of the layout-only model's encoder slice: as scene_encoder_specs, plus the second predicate table, without
    rel_s_mlp"""
    sd = arch.make_state_dict(arch.scene_encoder_specs(scene_box_cfg()), WEIGHT_SEED_SCENE + 1)
    return type(sd)((k, v) for k, v in sd.items() if not k.startswith("rel_s_mlp."))


def scene_glue_inputs(name: str):
    """-> (positional tensor arguments of the method, marked nodes or None).  `changes`: nodes 1 and 5 manipulated;
    `additions`: missing_nodes [2, 4] -> rows inserted at [2, 5] of the 6-node encoder scene (8 decoder nodes)."""
    g, objs, text, rel = scene_inputs()
    if name in ("sample", "box_sample"):
        return (objs, g.triples, text, rel), None
    name = name[4:] if name.startswith("box_") else name
    if name.startswith("changes"):
        ge, eobjs, etext, erel = scene_inputs(SCENE_ENC_CASE_SAME)
        return (eobjs, ge.triples, etext, erel, objs, g.triples, text, rel), [5, 1]
    if name.startswith("additions"):
        ge, eobjs, etext, erel = scene_inputs(SCENE_ENC_CASE_SMALL)
        return (eobjs, ge.triples, etext, erel, objs, g.triples, text, rel), [2, 4]
    raise KeyError(name)


def vqvae_sdf_inputs(n: int = 1, seed: int = 8):
    """truncated-SDF-like volumes (N, 1, 64, 64, 64), clamped to +-0.2 as the dataset does (threedfront_dataset.py)"""
    gen = torch.Generator().manual_seed(seed + 700)
    return (torch.randn(n, 1, 64, 64, 64, generator=gen) * 0.15).clamp(-0.2, 0.2)
