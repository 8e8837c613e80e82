"""GPU: VQVAE.encode_no_quant (Encoder3D -> quant_conv in csrc/vqvae.cu behind echo_vqvae_encode) against the
reference's output (tests/golden/vqvae_encode.pt, made by oracle/gen_golden_vqvae.py from the reference's own VQVAE module)
and the oracle; fp32 contract 1e-3 relative (tests/util.py)."""
import pytest
import torch

from echoscene_b200 import _lib, arch, modules
from oracle import cases, echoscene_oracle as orc
from util import FP32_TOL, assert_close, gold

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _dd(cfg):
    return dict(double_z=False, z_channels=cfg.z_channels, resolution=cfg.resolution, in_channels=1, out_ch=cfg.out_ch, ch=cfg.ch,
                ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0)


def _state(cfg):
    sd = dict(arch.make_state_dict(arch.vqvae_encode_specs(cfg), cases.WEIGHT_SEED_VQVAE + 1))   # as oracle/gen_golden_vqvae.py
    sd.update(arch.make_state_dict(arch.vqvae_decode_specs(cfg), cases.WEIGHT_SEED_VQVAE))
    return sd


@pytest.fixture(scope="module")
def model():
    cfg = cases.vqvae_cfg()
    m = modules.VQVAE(_dd(cfg), cfg.n_embed, cfg.embed_dim, with_encoder=True)
    m.load_state_dict(_state(cfg), strict=True)
    return m.to(DEV)


def test_vqvae_encode_vs_reference_golden(model):
    x = cases.vqvae_sdf_inputs()
    z = model.encode_no_quant(x.to(DEV))
    G = gold("vqvae_encode.pt")
    assert z.shape == (1, 3, 16, 16, 16)
    assert_close(z, G["z"], FP32_TOL, "encode_no_quant fp32")


def test_vqvae_encode_objects_are_independent_and_deterministic(model):
    x = cases.vqvae_sdf_inputs(3, seed=11).to(DEV)
    a = model.encode_no_quant(x)
    assert torch.equal(a, model.encode_no_quant(x)) and torch.isfinite(a).all()
    assert torch.equal(a[2:3], model.encode_no_quant(x[2:3]))
    assert model.encode_no_quant(x[:0]).shape == (0, 3, 16, 16, 16)      # empty batch
    old = model.max_encode_chunk
    model.max_encode_chunk = 2                                            # chunked (3 = 2 + 1) is bit-identical
    try:
        assert torch.equal(model.encode_no_quant(x), a)
    finally:
        model.max_encode_chunk = old
    with torch.no_grad():
        want = orc.vqvae_encode_no_quant({k: v for k, v in _state(cases.vqvae_cfg()).items()}, cases.vqvae_cfg(), x[1:2].cpu())
    assert_close(a[1:2], want, FP32_TOL, "encode_no_quant vs oracle, second object")


def test_vqvae_encode_then_decode_surface(model):
    """the two halves on one module: latents of the encoder feed decode_no_quant (shapes, finiteness; weights are random)."""
    x = cases.vqvae_sdf_inputs(1, seed=12).to(DEV)
    z = model.encode_no_quant(x)
    dec = model.decode_no_quant(z)
    assert dec.shape == (1, 1, 64, 64, 64) and torch.isfinite(dec).all()
    cfg = cases.vqvae_cfg()
    only_dec = modules.VQVAE(_dd(cfg), cfg.n_embed, cfg.embed_dim).to(DEV)
    with pytest.raises(_lib.EchoError, match="with_encoder"):
        only_dec.encode_no_quant(x)
    with pytest.raises(_lib.EchoError):
        model.encode_no_quant(x[:, :, :32])
