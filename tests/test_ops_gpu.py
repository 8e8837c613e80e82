"""GPU: every kernel family of the hot path, called through the C ABI, against a plain PyTorch fp32 reference of the
same op computed on the CPU (deterministic, no TF32)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from echoscene_b200 import _lib
from util import BF16_OP_TOL, BF16_TOL, FP32_TOL, assert_close

pytestmark = pytest.mark.gpu


def _cl(x):   # NCDHW -> channels-last (n,d,h,w,c)
    return x.permute(0, 2, 3, 4, 1).contiguous()


def _conv(x, w, b, k, stride_hw, prec):
    n, cin, d, h, ww = x.shape
    cout = w.shape[0]
    pad = k // 2
    oh = (h + 2 * pad - k) // stride_hw + 1
    ow = (ww + 2 * pad - k) // stride_hw + 1
    xc = _cl(x).cuda()
    out = torch.empty(n, d, oh, ow, cout, device="cuda")
    wd, bd = w.cuda().contiguous(), (b.cuda() if b is not None else None)
    _lib.check(_lib.lib().echo_op_conv3d(xc.data_ptr(), n, d, h, ww, cin, wd.data_ptr(), _lib.ptr(bd), cout, k, 1,
                                         stride_hw, out.data_ptr(), prec, _lib.stream_ptr()))
    return out.permute(0, 4, 1, 2, 3).cpu()


@pytest.mark.parametrize("cin,cout,dhw,k,stride", [
    (3, 224, (16, 16, 16), 3, 1),       # stem: K = 81, generic loader
    (224, 224, (16, 16, 16), 3, 1),     # the most frequent contraction (Appendix E)
    (448, 672, (16, 4, 4), 3, 1),
    (224, 224, (16, 16, 16), 3, 2),     # Downsample, stride (1,2,2)
    (1120, 448, (16, 8, 8), 1, 1),      # 1x1x1 skip conv
    (224, 3, (16, 16, 16), 3, 1),       # output conv: small-cout kernel
    (32, 64, (8, 8, 8), 3, 1),          # shape_embeddings.2
    (40, 24, (3, 5, 7), 3, 1),          # ragged: nothing divides the tile sizes
])
def test_conv3d_fp32(cin, cout, dhw, k, stride):
    g = torch.Generator().manual_seed(cin * 7 + cout)
    n = 2
    x = torch.randn(n, cin, *dhw, generator=g)
    w = torch.randn(cout, cin, k, k, k, generator=g) / (cin * k ** 3) ** 0.5
    b = torch.randn(cout, generator=g)
    want = F.conv3d(x, w, b, stride=(1, stride, stride), padding=k // 2)
    got = _conv(x, w, b, k, stride, _lib.PREC_FP32)
    assert_close(got, want, 1e-5, f"conv3d {cin}->{cout} k{k} s{stride}")


@pytest.mark.parametrize("rows,cin,cout", [(1, 8, 64), (8, 512, 2048), (16, 2048, 512), (33, 1280, 448), (64, 1664, 256),
                                           (200, 256, 640), (4096, 448, 1344)])
def test_linear_fp32(rows, cin, cout):
    g = torch.Generator().manual_seed(rows + cin)
    x = torch.randn(rows, cin, generator=g)
    w = torch.randn(cout, cin, generator=g) / cin ** 0.5
    b = torch.randn(cout, generator=g)
    want = F.linear(x, w, b)
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    out = torch.empty(rows, cout, device="cuda")
    _lib.check(_lib.lib().echo_op_linear(xd.data_ptr(), rows, cin, wd.data_ptr(), bd.data_ptr(), cout, out.data_ptr(),
                                         _lib.PREC_FP32, _lib.stream_ptr()))
    assert_close(out.cpu(), want, 1e-5, f"linear {rows}x{cin}->{cout}")


@pytest.mark.parametrize("n,voxels,c,eps,silu", [(2, 4096, 224, 1e-5, 1), (3, 256, 1344, 1e-6, 0), (2, 1024, 1120, 1e-5, 1),
                                                 (8, 1, 512, 1e-5, 1), (5, 1, 1024, 1e-6, 0), (1, 37, 96, 1e-5, 0)])
def test_group_norm(n, voxels, c, eps, silu):
    g = torch.Generator().manual_seed(c + voxels)
    x = torch.randn(n, voxels, c, generator=g) * 2 + 0.5
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    want = F.group_norm(x.permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        want = F.silu(want)
    want = want.permute(0, 2, 1)
    xd, gd, bd = x.cuda(), gamma.cuda(), beta.cuda()
    out = torch.empty_like(xd)
    _lib.check(_lib.lib().echo_op_group_norm(xd.data_ptr(), n, voxels, c, 32, gd.data_ptr(), bd.data_ptr(),
                                             eps, silu, out.data_ptr(), _lib.stream_ptr()))
    assert_close(out.cpu(), want, 1e-5, "group_norm")


@pytest.mark.parametrize("rows,c", [(1024, 448), (256, 672), (7, 512)])
def test_layer_norm(rows, c):
    g = torch.Generator().manual_seed(c)
    x = torch.randn(rows, c, generator=g) * 3 - 1
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    want = F.layer_norm(x, (c,), gamma, beta, 1e-5)
    xd, gd, bd = x.cuda(), gamma.cuda(), beta.cuda()
    out = torch.empty_like(xd)
    _lib.check(_lib.lib().echo_op_layer_norm(xd.data_ptr(), rows, c, gd.data_ptr(), bd.data_ptr(), 1e-5,
                                             out.data_ptr(), _lib.stream_ptr()))
    assert_close(out.cpu(), want, 1e-5, "layer_norm")


def _attention_ref(qkv, n, tokens, heads, dh):
    C_ = heads * dh
    q, k, v = qkv.view(n, tokens, 3, heads, dh).permute(2, 0, 3, 1, 4)   # (n, heads, tokens, dh)
    sim = torch.einsum("bhid,bhjd->bhij", q, k) * dh ** -0.5            # attention.py:203
    out = torch.einsum("bhij,bhjd->bhid", sim.softmax(-1), v)
    return out.permute(0, 2, 1, 3).reshape(n * tokens, C_)


@pytest.mark.parametrize("n,c,cout,dhw", [(2, 256, 256, (16, 16, 16)), (1, 128, 128, (32, 32, 32)), (3, 64, 96, (4, 4, 8))])
def test_upsample_x2_conv_folded(n, c, cout, dhw):
    """Upsample(nearest x2 in d, h, w) + Conv3d k3 (vqvae_modules.py:24-39) as eight phase convolutions on the low-res input."""
    _need_tc()
    g = torch.Generator().manual_seed(c + cout + 1)
    x = torch.randn(n, c, *dhw, generator=g).bfloat16().float()
    w = (torch.randn(cout, c, 3, 3, 3, generator=g) / (c * 27) ** 0.5)
    b = torch.randn(cout, generator=g)
    want = F.conv3d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1)
    xc = _cl(x).cuda()
    out = torch.empty(n, 2 * dhw[0], 2 * dhw[1], 2 * dhw[2], cout, device="cuda")
    wd, bd = w.cuda().contiguous(), b.cuda()
    _lib.check(_lib.lib().echo_op_upconv3d_x2(xc.data_ptr(), n, *dhw, c, wd.data_ptr(), bd.data_ptr(), cout, out.data_ptr(),
                                              _lib.PREC_BF16, _lib.stream_ptr()))
    assert_close(out.permute(0, 4, 1, 2, 3).cpu(), want, 1.5e-2, "folded x2 upsample conv")


@pytest.mark.parametrize("n,tokens,heads,dh", [(2, 1024, 8, 56), (2, 256, 8, 84), (1, 50, 3, 20)])
def test_attention_fp32(n, tokens, heads, dh):
    g = torch.Generator().manual_seed(tokens + dh)
    qkv = torch.randn(n * tokens, 3 * heads * dh, generator=g)
    want = _attention_ref(qkv, n, tokens, heads, dh)
    qd = qkv.cuda()
    out = torch.empty(n * tokens, heads * dh, device="cuda")
    _lib.check(_lib.lib().echo_op_attention(qd.data_ptr(), n, tokens, heads, dh, out.data_ptr(), _lib.PREC_FP32, _lib.stream_ptr()))
    assert_close(out.cpu(), want, 1e-5, "attention")


def test_gather_rows_bit_exact():
    g = torch.Generator().manual_seed(3)
    for D in (768, 1408, 10):
        obj = torch.randn(32, D, generator=g)
        idx = torch.randint(0, 32, (128,), generator=g)
        od, idd = obj.cuda(), idx.cuda()
        out = torch.empty(128, D, device="cuda")
        _lib.check(_lib.lib().echo_gather_rows(od.data_ptr(), idd.data_ptr(), 128, 32, D, out.data_ptr(), _lib.stream_ptr()))
        assert torch.equal(out.cpu(), obj[idx])                      # model/graph.py:146-147, bit-exact
    out = torch.empty(0, 8, device="cuda")
    _lib.check(_lib.lib().echo_gather_rows(od.data_ptr(), None, 0, 32, 8, out.data_ptr(), _lib.stream_ptr()))   # empty edge list


def _need_tc():
    if not _lib.lib().echo_has_tcgen05():
        pytest.skip("tcgen05 kernels not available on this device/build")


@pytest.mark.parametrize("cin,cout,dhw,k,stride", [
    (224, 224, (16, 16, 16), 3, 1), (448, 448, (16, 8, 8), 3, 1), (672, 672, (16, 4, 4), 3, 1),
    (1120, 448, (16, 8, 8), 3, 1), (672, 224, (16, 16, 16), 3, 1), (224, 224, (16, 16, 16), 3, 2),
    (448, 448, (16, 8, 8), 3, 2), (1344, 672, (16, 4, 4), 1, 1), (448, 3584, (16, 8, 8), 1, 1),
])
def test_conv3d_tcgen05(cin, cout, dhw, k, stride):
    _need_tc()
    g = torch.Generator().manual_seed(cin + cout + k)
    n = 3
    x = torch.randn(n, cin, *dhw, generator=g).bfloat16().float()
    w = (torch.randn(cout, cin, k, k, k, generator=g) / (cin * k ** 3) ** 0.5).bfloat16().float()
    b = torch.randn(cout, generator=g)
    want = F.conv3d(x, w, b, stride=(1, stride, stride), padding=k // 2)   # exact products of bf16 values, fp32 sums
    got = _conv(x, w, b, k, stride, _lib.PREC_BF16)
    assert_close(got, want, 1e-2, f"tcgen05 conv {cin}->{cout} k{k} s{stride}")   # output rounded to bf16 (2^-9)


@pytest.mark.parametrize("n,c,cout,dhw", [(16, 448, 448, (16, 8, 8)), (4, 672, 672, (16, 4, 4)), (3, 64, 96, (4, 4, 8))])
def test_upsample_conv_folded(n, c, cout, dhw):
    """Upsample(nearest x(1,2,2)) + Conv3d k3 (openai_model_3d.py:150-157) as four phase convolutions on the low-res input."""
    _need_tc()
    g = torch.Generator().manual_seed(c + cout)
    x = torch.randn(n, c, *dhw, generator=g).bfloat16().float()
    w = (torch.randn(cout, c, 3, 3, 3, generator=g) / (c * 27) ** 0.5)
    b = torch.randn(cout, generator=g)
    up = F.interpolate(x, (dhw[0], dhw[1] * 2, dhw[2] * 2), mode="nearest")
    want = F.conv3d(up, w, b, padding=1)
    xc = _cl(x).cuda()
    out = torch.empty(n, dhw[0], 2 * dhw[1], 2 * dhw[2], cout, device="cuda")
    wd, bd = w.cuda().contiguous(), b.cuda()
    _lib.check(_lib.lib().echo_op_upconv3d(xc.data_ptr(), n, *dhw, c, wd.data_ptr(), bd.data_ptr(), cout, out.data_ptr(),
                                           _lib.PREC_BF16, _lib.stream_ptr()))
    # folded weights are rounded to bf16 after the fp32 fold (2^-9 each) and the output is rounded to bf16
    assert_close(out.permute(0, 4, 1, 2, 3).cpu(), want, 1.5e-2, "folded upsample conv")


@pytest.mark.parametrize("n,tokens,heads,dh", [(2, 1024, 8, 56), (2, 256, 8, 84), (16, 1024, 8, 56), (40, 256, 8, 84), (3, 128, 8, 56),
                                               (2, 256, 4, 64), (1, 64, 8, 56)])
def test_attention_bf16(n, tokens, heads, dh):
    _need_tc()
    g = torch.Generator().manual_seed(tokens + dh)
    qkv = torch.randn(n * tokens, 3 * heads * dh, generator=g).bfloat16().float()
    want = _attention_ref(qkv, n, tokens, heads, dh)
    qd = qkv.cuda()
    out = torch.empty(n * tokens, heads * dh, device="cuda")
    _lib.check(_lib.lib().echo_op_attention(qd.data_ptr(), n, tokens, heads, dh, out.data_ptr(), _lib.PREC_BF16, _lib.stream_ptr()))
    assert_close(out.cpu(), want, BF16_OP_TOL, "attention bf16")
