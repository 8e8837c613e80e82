"""CPU: the VQVAE module surface -- state_dict keys / shapes are the reference's, a reference checkpoint
(with encoder and quant_conv entries) loads, and nothing computes without the GPU library."""
import pytest
import torch

from echoscene_b200 import _lib, arch, modules
from oracle import cases


def _module():
    cfg = cases.vqvae_cfg()
    dd = dict(double_z=False, z_channels=cfg.z_channels, resolution=cfg.resolution, in_channels=1, out_ch=cfg.out_ch, ch=cfg.ch,
              ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0)
    return modules.VQVAE(dd, cfg.n_embed, cfg.embed_dim), cfg


def test_state_dict_matches_the_decode_specs():
    m, cfg = _module()
    specs = arch.vqvae_decode_specs(cfg)
    sd = m.state_dict()
    assert list(sd.keys()) == list(specs.keys())
    for k, sp in specs.items():
        assert tuple(sd[k].shape) == tuple(sp.shape), k
    assert arch.count_params(specs) == 14846349          # tests/golden/PINNING.json, counted on the reference module


def test_reference_checkpoint_layout_loads():
    m, cfg = _module()
    sd = arch.make_state_dict(arch.vqvae_decode_specs(cfg), cases.WEIGHT_SEED_VQVAE)
    full = dict(sd)
    full["encoder.conv_in.weight"] = torch.zeros(64, 1, 3, 3, 3)      # present in a reference checkpoint, not on this path
    full["quant_conv.weight"] = torch.zeros(3, 3, 1, 1, 1)
    res = m.load_state_dict(full, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["decoder.conv_out.weight"], sd["decoder.conv_out.weight"])


def test_no_cpu_path():
    m, _ = _module()
    with pytest.raises(Exception):
        m.decode_no_quant(torch.zeros(1, 3, 16, 16, 16))
    with pytest.raises(Exception):
        m.encode(torch.zeros(1, 1, 64, 64, 64))
    with pytest.raises(Exception):
        m(torch.zeros(1, 1, 64, 64, 64))


def test_vqvae_with_encoder_module_surface():
    """VQVAE(with_encoder=True) owns encoder.* and quant_conv.* under the reference's names; nothing computes on CPU;
    the encoder entry points reject bad arguments with codes."""
    import ctypes as C
    cfg = cases.vqvae_cfg()
    dd = dict(double_z=False, z_channels=cfg.z_channels, resolution=cfg.resolution, in_channels=1, out_ch=cfg.out_ch, ch=cfg.ch,
              ch_mult=list(cfg.ch_mult), num_res_blocks=cfg.num_res_blocks, attn_resolutions=[], dropout=0.0)
    m = modules.VQVAE(dd, cfg.n_embed, cfg.embed_dim, with_encoder=True)
    enc, dec = arch.vqvae_encode_specs(cfg), arch.vqvae_decode_specs(cfg)
    sd = m.state_dict()
    assert set(sd) == set(enc) | set(dec) and len(sd) == len(enc) + len(dec)
    for k, s in list(enc.items()) + list(dec.items()):
        assert tuple(sd[k].shape) == tuple(s.shape), k
    full = dict(arch.make_state_dict(enc, 3))
    full.update(arch.make_state_dict(dec, 4))
    m.load_state_dict(full, strict=True)
    with pytest.raises(_lib.EchoError):
        m.encode_no_quant(torch.zeros(1, 1, 64, 64, 64))                # CPU tensor: no fallback
    only_dec = modules.VQVAE(dd, cfg.n_embed, cfg.embed_dim)
    with pytest.raises(_lib.EchoError, match="with_encoder"):
        only_dec.encode_no_quant(torch.zeros(1, 1, 64, 64, 64))
    with pytest.raises(_lib.EchoError):
        modules.VQVAE(dict(dd, double_z=True), cfg.n_embed, cfg.embed_dim, with_encoder=True)
    L = _lib.lib()
    h = C.c_void_p()
    assert L.echo_vqvae_encoder_create(C.byref(h), None, None, 0) == -1
    d = m._desc(1, "bf16")
    assert L.echo_vqvae_encoder_create(C.byref(h), C.byref(d), None, 0) == -4 and b"ECHO_PREC_FP32" in L.echo_last_error()
    assert L.echo_vqvae_encode(None, None, 1, None, None) == -1
