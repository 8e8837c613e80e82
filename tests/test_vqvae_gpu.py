"""GPU: VQVAE.decode_no_quant through the C ABI against the reference's output (tests/golden/vqvae_decode.pt,
made by oracle/gen_golden_vqvae.py from the reference's own VQVAE module)."""
import pytest
import torch

from echoscene_b200 import _lib, arch, modules
from oracle import cases
from util import BF16_OP_TOL, BF16_TOL, FP32_TOL, assert_close, gold

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _model(precision):
    cfg = cases.vqvae_cfg()
    dd = dict(double_z=False, z_channels=cfg.z_channels, resolution=cfg.resolution, in_channels=1, out_ch=cfg.out_ch, ch=cfg.ch,
              ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0)
    m = modules.VQVAE(dd, cfg.n_embed, cfg.embed_dim, precision=precision)
    sd = arch.make_state_dict(arch.vqvae_decode_specs(cfg), cases.WEIGHT_SEED_VQVAE)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


def test_vqvae_decode_fp32_vs_reference_golden():
    m = _model("fp32")
    z = cases.vqvae_inputs()
    G = gold("vqvae_decode.pt")
    dec, idx = m.decode_no_quant(z.to(DEV), return_indices=True)
    assert torch.equal(idx.cpu(), G["indices"])                       # integer work: bit-exact (quantizer.py:84)
    assert dec.shape == (z.shape[0], 1, 64, 64, 64)
    assert_close(dec[:, :, ::2, ::2, ::2], G["dec_sub"], FP32_TOL, "decode_no_quant fp32")
    assert abs(float(dec.double().sum()) - float(G["dec_sum"])) < 1e-3 * float(G["dec_abs_sum"])


def test_vqvae_decode_objects_are_independent_and_deterministic():
    m = _model("fp32")
    z = cases.vqvae_inputs(3, seed=9).to(DEV)
    a = m.decode_no_quant(z)
    b = m.decode_no_quant(z)
    assert torch.equal(a, b) and torch.isfinite(a).all()
    one = m.decode_no_quant(z[1:2])
    assert torch.equal(a[1:2], one)
    assert m.decode_no_quant(z[:0]).shape == (0, 1, 64, 64, 64)         # empty batch
    m.max_chunk = 2                                                      # chunked decoding (3 = 2 + 1) is bit-identical
    c, idx = m.decode_no_quant(z, return_indices=True)
    assert torch.equal(c, a) and idx.shape == (3 * 4096,)


def test_vqvae_decode_bf16_vs_reference_golden():
    if not _lib.lib().echo_has_tcgen05():
        pytest.skip("tcgen05 kernels not available")
    m = _model("bf16")
    z = cases.vqvae_inputs()
    G = gold("vqvae_decode.pt")
    dec, idx = m.decode_no_quant(z.to(DEV), return_indices=True)
    assert torch.equal(idx.cpu(), G["indices"])                       # the quantiser stays fp32 in every mode
    assert_close(dec[:, :, ::2, ::2, ::2], G["dec_sub"], BF16_OP_TOL, "decode_no_quant bf16")


def test_vqvae_surface_errors():
    m = _model("fp32")
    with pytest.raises(Exception):
        m.encode(torch.zeros(1, 1, 64, 64, 64, device=DEV))
    with pytest.raises(Exception):
        m.decode_no_quant(torch.zeros(1, 3, 16, 16, 16))                # CPU tensor: no CPU path
