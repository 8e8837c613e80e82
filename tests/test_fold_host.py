"""CPU: the weight fold behind the upsample-folded convolutions (host code of libechoscene_b200, no GPU needed).
nearest-upsample followed by a 3x3x3 conv equals, per output phase, a conv with 2 taps per upsampled axis on the LOW-res
input (openai_model_3d.py:150-157 for x(1,2,2), vqvae_modules.py:35-39 for x2): the test applies the folded taps with
the kernel's offset convention (tap a of phase p reads low-res offset p - 1 + a) in plain torch."""
import pytest
import torch
import torch.nn.functional as F

from echoscene_b200 import _lib


def _fold(w, up_depth):
    cout, cin = w.shape[:2]
    wt = w.permute(0, 2, 3, 4, 1).reshape(cout, 27, cin).contiguous()          # [cout][tap (kd,kh,kw)][cin], as repacked on the device
    out = torch.empty(cout, 64 if up_depth else 48, cin)
    _lib.check(_lib.lib().echo_debug_fold_upsample_weight(wt.data_ptr(), cout, cin, 1 if up_depth else 0, out.data_ptr()))
    return out


@pytest.mark.parametrize("up_depth", [False, True])
def test_folded_taps_equal_upsample_then_conv(up_depth):
    g = torch.Generator().manual_seed(5 + int(up_depth))
    n, cin, cout, D, H, W = 2, 3, 4, 3, 4, 5
    x = torch.randn(n, cin, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64)
    scale = (2.0, 2.0, 2.0) if up_depth else (1.0, 2.0, 2.0)
    want = F.conv3d(F.interpolate(x, scale_factor=scale, mode="nearest"), w, padding=1)
    wf = _fold(w.float(), up_depth).double()
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))                                          # zero padding == the TMA out-of-bounds fill
    got = torch.zeros_like(want)
    if up_depth:
        wf = wf.view(cout, 8, 8, cin)
        for ph in range(8):
            pz, py, px = (ph >> 2) & 1, (ph >> 1) & 1, ph & 1
            acc = torch.zeros(n, cout, D, H, W, dtype=torch.float64)
            for t in range(8):
                dz, dy, dx = pz - 1 + ((t >> 2) & 1), py - 1 + ((t >> 1) & 1), px - 1 + (t & 1)
                sl = xp[:, :, 1 + dz:1 + dz + D, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                acc += torch.einsum("oc,ncdhw->nodhw", wf[:, ph, t], sl)
            got[:, :, pz::2, py::2, px::2] = acc
    else:
        wf = wf.view(cout, 4, 12, cin)
        for ph in range(4):
            py, px = (ph >> 1) & 1, ph & 1
            acc = torch.zeros(n, cout, D, H, W, dtype=torch.float64)
            for t in range(12):
                kd, a, b = t >> 2, (t >> 1) & 1, t & 1
                dz, dy, dx = kd - 1, py - 1 + a, px - 1 + b
                sl = xp[:, :, 1 + dz:1 + dz + D, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
                acc += torch.einsum("oc,ncdhw->nodhw", wf[:, ph, t], sl)
            got[:, :, :, py::2, px::2] = acc
    assert (got - want).abs().max() < 1e-5 * want.abs().max()
