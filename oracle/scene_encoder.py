"""TEST INFRASTRUCTURE: modules.SceneEncoder's method surface on the CPU oracle, so that the host glue of
echoscene_b200/scene.py can be exercised without a GPU (tests/test_scene_host.py, oracle/gen_golden_scene_glue.py).
Never imported by the product path."""
from __future__ import annotations

from echoscene_b200 import arch
from oracle import cases, echoscene_oracle as orc


class OracleSceneEncoder:
    def __init__(self, box: bool = False):
        if box:   # the layout-only model: second predicate table for manipulate, no rel_s_mlp (model/EchoLayout.py)
            self.cfg, self.sd = cases.scene_box_cfg(), cases.scene_box_state_dict()
        else:
            self.cfg = cases.scene_cfg()
            self.sd = arch.make_state_dict(arch.scene_encoder_specs(self.cfg), cases.WEIGHT_SEED_SCENE)
        self.embedding_dim = self.cfg.gconv_dim
        self.out_dim_ini_encoder = self.cfg.feat_dim

    def init_encoder(self, objs, triples, text, rel):
        return orc.scene_init_encoder(self.sd, self.cfg, objs, triples, text, rel)

    def manipulate(self, latent_f, objs, triples, text, rel):
        return orc.scene_manipulate(self.sd, self.cfg, latent_f, objs, triples, text, rel)

    def rel_s(self, x):
        return orc.scene_rel_s(self.sd, x)

    def encode(self, objs, triples, text, rel, change=None, shape_cond=True):
        return orc.scene_encode(self.sd, self.cfg, objs, triples, text, rel, change)
