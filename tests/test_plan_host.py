"""CPU: the launch plan of the tcgen05 contraction kernel (host logic of libechoscene_b200, gemm_tc.cu::tc_plan) for the
contraction problems of one shape step at N = 16 on a 148-SM B200 -- the decisions DESIGN.md section 4
describes: two sub-blocks per CTA for the long-K convolutions, split-K over tap groups for the 16x4x4 level, a 256-wide tile
for the GEGLU projection, one sub-block for the short-K token GEMMs."""
import ctypes as C

import pytest

from echoscene_b200 import _lib

SMS = 148


def plan(n, dhw, cin, cout, k, epi=0, up2=0, splitk=1):
    out = (C.c_int32 * 4)()
    _lib.lib().echo_debug_tc_plan(n, dhw[0], dhw[1], dhw[2], cin, cout, k, epi, up2, splitk, SMS, out)
    return {"block_n": out[0], "msub": out[1], "splitk": out[2], "cta2": out[3]}


def test_long_k_convs_share_b_between_two_sub_blocks():
    for cin, cout, dhw in [(224, 224, (16, 16, 16)), (448, 448, (16, 8, 8)), (448, 224, (16, 16, 16)), (1120, 448, (16, 8, 8))]:
        p = plan(16, dhw, cin, cout, 3)
        assert p == {"block_n": 224, "msub": 2, "splitk": 1, "cta2": 1}, (cin, cout, dhw, p)


def test_coarse_level_splits_k_when_a_workspace_is_offered():
    p = plan(16, (16, 4, 4), 672, 672, 3, splitk=1)
    assert p["splitk"] == 3 and p["msub"] == 2 and p["block_n"] == 224 and p["cta2"] == 1
    q = plan(16, (16, 4, 4), 672, 672, 3, splitk=0)
    assert q["splitk"] == 1 and q["block_n"] % 32 == 0


def test_token_gemms():
    assert plan(16, (16, 8, 8), 448, 3584, 1, epi=1) == {"block_n": 256, "msub": 1, "splitk": 1, "cta2": 1}   # GEGLU: [128 a | 128 g] tiles
    p = plan(16, (16, 8, 8), 448, 448, 1)
    assert p["msub"] == 1 and p["splitk"] == 1 and p["block_n"] in (192, 224)
    assert plan(16, (16, 16, 16), 224, 32, 3)["block_n"] == 32                                                    # padded output conv


def test_upsample_folded_convs_and_odd_tile_counts():
    p = plan(16, (16, 8, 8), 448, 448, 3, up2=1)
    assert p["splitk"] == 1 and p["block_n"] == 224
    p = plan(16, (16, 16, 16), 256, 256, 3, up2=2)
    assert p["splitk"] == 1 and p["block_n"] in (128, 256)
    assert plan(1, (3, 5, 7), 64, 96, 3)["cta2"] in (0, 1)          # ragged geometry still yields a plan
    assert plan(1, (3, 5, 7), 64, 96, 3)["block_n"] > 0
