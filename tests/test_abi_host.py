"""CPU: the C-ABI library loads, exports every symbol include/echoscene_b200.h declares, rejects bad arguments with
codes (never crashes), and the host-side mirror keeps the reference's state_dict keys."""
import ctypes as C
import os
import re

import pytest
import torch

from echoscene_b200 import _lib, arch, modules, synth
from oracle import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "echoscene_b200.h")).read()
    return sorted(set(re.findall(r"ECHO_API [^;()]*?\b(echo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = _declared()
    assert len(names) >= 29
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names
    assert L.echo_version() == 1


def test_bad_arguments_return_codes():
    L = _lib.lib()
    assert L.echo_gather_rows(None, None, 4, 4, 16, None, None) == -1
    assert b"gather_rows" in L.echo_last_error()
    h = C.c_void_p()
    assert L.echo_layout_create(C.byref(h), None, None, 0) == -1
    assert L.echo_shape_create(C.byref(h), None, None, 0) == -1
    assert L.echo_gcn_create(C.byref(h), None, None, 0) == -1
    assert L.echo_op_conv3d(None, 1, 1, 1, 1, 16, None, None, 16, 5, 1, 1, None, 0, None) == -1
    with pytest.raises(_lib.EchoError):
        _lib.check(-1)
    # the entry points added in round 2 validate before they touch the device
    assert L.echo_gcn_train_create(C.byref(h), None, None, 0, None, 0) < 0 and b"gcn_train" in L.echo_last_error()
    assert L.echo_gcn_train_forward(None, None, None, None, None, None, None) < 0
    assert L.echo_gcn_train_backward(None, None, None, None, None, None, None) < 0
    for fn in (L.echo_layout_set_batch_stats, L.echo_shape_set_batch_stats, L.echo_scene_set_batch_stats):
        assert fn(None, 1) < 0 and b"set_batch_stats" in L.echo_last_error()
    assert L.echo_mesh_workspace_bytes(1) == -1 and L.echo_mesh_workspace_bytes(1000) == -1 and L.echo_mesh_workspace_bytes(64) > 4 * 4 * 64 ** 3
    assert L.echo_mesh_marching_cubes(None, 64, 0.02, None, 0, None, 0, None, None, 0, None) < 0 and b"marching_cubes" in L.echo_last_error()
    assert L.echo_optimizer_create(None, None, 0) < 0
    L.echo_gcn_train_destroy(None)                                              # destroying nothing is allowed


def test_state_dict_keys_match_reference_specs():
    for cls, kw, specs in (
        (modules.UNet1DModel, dict(in_channels=8, model_channels=512, out_channels=8, num_res_blocks=2,
                                   attention_resolutions=[4, 2], channel_mult=[1, 1, 1, 1], num_heads=8,
                                   use_spatial_transformer=True, concat_dim=1280, crossattn_dim=1280, enable_t_emb=True),
         arch.unet1d_specs(cases.layout_cfg())),
    ):
        m = cls(**kw)
        sd = m.state_dict()
        assert list(sd.keys()) == list(specs.keys())
        for k, s in specs.items():
            assert tuple(sd[k].shape) == tuple(s.shape), k
    g = modules.GraphTripleConvNet(768, 128, num_layers=5, hidden_dim=256, residual=True, pooling="avg",
                                   mlp_normalization="batch", output_dim=1280)
    assert list(g.state_dict().keys()) == list(arch.gcn_specs(cases.layout_cfg().gcn()).keys())
    one = modules.GraphTripleConv(768, 128, output_dim=1280, hidden_dim=256, mlp_normalization="batch")
    assert "net1.0.weight" in one.state_dict() and "linear_projection_pred.bias" in one.state_dict()


def test_no_cpu_fallback():
    g = modules.GraphTripleConvNet(64, 16, num_layers=1, hidden_dim=32, residual=True, mlp_normalization="batch")
    with pytest.raises(_lib.EchoError):
        g(torch.randn(4, 64), torch.randn(3, 16), torch.zeros(3, 2, dtype=torch.int64))
    with pytest.raises(_lib.EchoError):
        modules.GraphTripleConvNet(64, 16, pooling="wAvg")
    g.train()
    with pytest.raises(_lib.EchoError):
        g(torch.randn(4, 64), torch.randn(3, 16), torch.zeros(3, 2, dtype=torch.int64))


def test_synthetic_graph_contract():
    g = synth.make_scene_graph(16, 64, 2)
    t = g.triples
    assert t.shape == (64, 3) and t.dtype == torch.int64
    assert (t[:15, 1] == 0).all() and (t[:15, 2] == 15).all()          # "in" edges to the _scene_ node
    assert (t[15:, 0] != t[15:, 2]).all() and t[15:, 1].min() >= 1 and t[:, 1].max() <= 15
    b = synth.batch_scene_graphs([g, g])
    assert b.n_nodes == 32 and (b.triples[64:, 0] == t[:, 0] + 16).all()
    with pytest.raises(ValueError):
        synth.make_scene_graph(2, 4, 0)


def test_ctypes_structs_match_the_header_layout(tmp_path):
    """Every descriptor struct of the header, compiled as C (the header must stay plain C), has the size and field offsets of
    its ctypes mirror in _lib.py."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    pairs = {"echo_weight_t": _lib.Weight, "echo_gcn_desc_t": _lib.GcnDesc, "echo_layout_desc_t": _lib.LayoutDesc,
             "echo_shape_desc_t": _lib.ShapeDesc, "echo_scene_desc_t": _lib.SceneDesc, "echo_vqvae_desc_t": _lib.VqvaeDesc,
             "echo_opt_tensor_t": _lib.OptTensor}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{os.path.join(ROOT, "include", "echoscene_b200.h")}"',
             'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == len(pairs)
    for line, (cname, cls) in zip(out, pairs.items()):
        got = line.split()
        want = [cname, str(C.sizeof(cls))] + [str(getattr(cls, f).offset) for f, _ in cls._fields_]
        assert got == want, f"{cname}: header {got} vs ctypes {want}"


def test_plain_c_client_links_and_calls(tmp_path):
    """The boundary is a C ABI: a C99 translation unit that includes the header links against the library with no C++/torch in
    sight, calls host-only entry points and gets codes + messages back (no GPU needed for these calls)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    src = tmp_path / "client.c"
    src.write_text(r"""
#include <stdio.h>
#include <string.h>
#include "echoscene_b200.h"
int main(void) {
  float tab[5 * 10];
  int64_t tri[6] = {0, 3, 1, 1, 5, 0};
  int32_t off[3], items[4];
  int64_t range[2];
  echo_gcn_t* h = NULL;
  if (echo_version() != ECHO_ABI_VERSION) return 1;
  if (echo_debug_ddpm_tables(10, 1e-4f, 0.02f, tab) != ECHO_OK) return 2;
  if (echo_debug_graph_csr(tri, 2, 2, off, items, range) != ECHO_OK) return 3;
  if (echo_gcn_create(&h, NULL, NULL, 0) != ECHO_ERR_INVALID || strlen(echo_last_error()) == 0) return 4;
  printf("%d %.6f %d %d %d %d %lld %lld\n", echo_version(), tab[0], off[2], items[0], items[1], items[3], (long long)range[0],
         (long long)range[1]);
  return 0;
}
""")
    exe = tmp_path / "client"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-lechoscene_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    # node 0: subject of t0 (item 0), object of t1 (item 3); node 1: subject of t1 (item 2), object of t0 (item 1)
    assert out[0] == "1" and out[2:] == ["4", "0", "3", "1", "3", "5"]
    assert abs(float(out[1]) - 1.00005) < 1e-4                 # sqrt(1 / alphas_cumprod[0]) = sqrt(1 / (1 - 1e-4))


def test_product_never_imports_the_oracle():
    """oracle/ is the checker: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it."""
    import ast
    pkg = os.path.join(ROOT, "echoscene_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith(".py"):
                continue
            tree = ast.parse(open(os.path.join(dirpath, f)).read())
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), f"{f} imports the oracle"
    # importing the whole product surface must not pull it in either
    import subprocess
    import sys
    code = ("import sys; import echoscene_b200.modules, echoscene_b200.samplers, echoscene_b200.scene, echoscene_b200.sgdiff, "
            "echoscene_b200.integrate, echoscene_b200.shard, echoscene_b200.synth; "
            "sys.exit(1 if any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules) else 0)")
    assert subprocess.run([sys.executable, "-c", code], cwd=ROOT).returncode == 0


def test_missing_library_fails_loudly(monkeypatch):
    """No silent fallback: without the built shared library every entry into the product raises."""
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libechoscene_b200.so")
    with pytest.raises(_lib.EchoError, match="not built"):
        _lib.lib()
    g = modules.GraphTripleConvNet(64, 16, num_layers=1, hidden_dim=32, residual=True, mlp_normalization="batch")
    with pytest.raises(_lib.EchoError):
        g(torch.randn(4, 64), torch.randn(3, 16), torch.zeros(3, 2, dtype=torch.int64))
