"""ctypes binding of libechoscene_b200.so (the C ABI declared in include/echoscene_b200.h).

The product path has no CPU fallback: importing this module without the built library, or calling a compute entry
point without a CUDA device, raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"`` or
``echoscene_b200/csrc/build.sh``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Tuple

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libechoscene_b200.so")

PREC_FP32 = 0
PREC_BF16 = 1
PREC_X3 = 2
INDEX_FROM_DEVICE = -2   # ECHO_INDEX_FROM_DEVICE: the step reads its DDIM index from the handle's device slot


def precision_code(name: str) -> int:
    try:
        return {"fp32": PREC_FP32, "bf16": PREC_BF16, "x3": PREC_X3}[name]
    except KeyError:
        raise EchoError(f"unknown precision {name!r} (fp32 | bf16 | x3)") from None


class EchoError(RuntimeError):
    pass


class Weight(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data", C.c_void_p), ("ndim", C.c_int32), ("dtype", C.c_int32),
                ("shape", C.c_int64 * 6)]


class OptTensor(C.Structure):   # echo_opt_tensor_t
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("numel", C.c_int64), ("clip_group", C.c_int32), ("reserved", C.c_int32)]


class GcnDesc(C.Structure):
    _fields_ = [("input_dim_obj", C.c_int32), ("input_dim_pred", C.c_int32), ("num_layers", C.c_int32),
                ("hidden_dim", C.c_int32), ("output_dim", C.c_int32), ("max_nodes", C.c_int32),
                ("max_triples", C.c_int32), ("bn_eps", C.c_float), ("keep_train_weights", C.c_int32)]


class LayoutDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("model_channels", C.c_int32),
                ("num_levels", C.c_int32), ("channel_mult", C.c_int32 * 8), ("num_res_blocks", C.c_int32),
                ("num_attention_resolutions", C.c_int32), ("attention_resolutions", C.c_int32 * 8),
                ("num_heads", C.c_int32), ("context_dim", C.c_int32), ("obj_embed_dim", C.c_int32),
                ("gconv_dim", C.c_int32), ("enable_t_emb", C.c_int32), ("max_nodes", C.c_int32),
                ("max_triples", C.c_int32), ("precision", C.c_int32), ("time_num", C.c_int32),
                ("beta_start", C.c_float), ("beta_end", C.c_float), ("keep_train_weights", C.c_int32)]


class ShapeDesc(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("out_channels", C.c_int32), ("model_channels", C.c_int32),
                ("num_levels", C.c_int32), ("channel_mult", C.c_int32 * 8), ("num_res_blocks", C.c_int32),
                ("num_attention_resolutions", C.c_int32), ("attention_resolutions", C.c_int32 * 8),
                ("num_heads", C.c_int32), ("context_dim", C.c_int32), ("gconv_dim", C.c_int32),
                ("enable_t_emb", C.c_int32), ("latent_size", C.c_int32), ("max_nodes", C.c_int32),
                ("max_triples", C.c_int32), ("max_local_nodes", C.c_int32), ("precision", C.c_int32),
                ("timesteps", C.c_int32), ("ddim_steps", C.c_int32), ("linear_start", C.c_float),
                ("linear_end", C.c_float), ("keep_train_weights", C.c_int32)]


class SceneDesc(C.Structure):
    _fields_ = [("gconv_dim", C.c_int32), ("add_dim", C.c_int32), ("num_objs", C.c_int32), ("num_preds", C.c_int32),
                ("num_layers", C.c_int32), ("rel_s_hidden", C.c_int32), ("context_dim", C.c_int32),
                ("max_nodes", C.c_int32), ("max_triples", C.c_int32), ("bn_eps", C.c_float),
                ("manipulate_pred_dc", C.c_int32), ("keep_train_weights", C.c_int32)]


class VqvaeDesc(C.Structure):
    _fields_ = [("embed_dim", C.c_int32), ("n_embed", C.c_int32), ("z_channels", C.c_int32), ("latent_size", C.c_int32),
                ("ch", C.c_int32), ("num_levels", C.c_int32), ("ch_mult", C.c_int32 * 8), ("num_res_blocks", C.c_int32),
                ("out_ch", C.c_int32), ("max_objects", C.c_int32), ("precision", C.c_int32)]


_P = C.c_void_p
_I = C.c_int32
_L = C.c_int64

# name -> (restype, argtypes); must list every symbol of include/echoscene_b200.h (tests check this)
PROTOTYPES = {
    "echo_version": (C.c_int, []),
    "echo_last_error": (C.c_char_p, []),
    "echo_has_tcgen05": (C.c_int, []),
    "echo_launch_count": (C.c_int64, []),
    "echo_launch_count_reset": (None, []),
    "echo_debug_set_tc_mode": (None, [C.c_int]),
    "echo_debug_set_layout_mode": (None, [C.c_int]),
    "echo_debug_layout_info": (C.c_int, [_P, _P]),
    "echo_debug_probe_begin": (None, [C.c_int64, C.c_int, C.c_int, C.c_int]),
    "echo_debug_probe_end": (C.c_int, [C.POINTER(C.c_double)]),
    "echo_debug_probe_timeline": (None, [_P]),
    "echo_debug_fold_upsample_weight": (C.c_int, [_P, _I, _I, _I, _P]),
    "echo_debug_tc_plan": (None, [_I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "echo_debug_ddpm_tables": (C.c_int, [_I, C.c_float, C.c_float, _P]),
    "echo_debug_ddim_schedule": (C.c_int, [_I, _I, C.c_float, C.c_float, _I, _P, _P, _P]),
    "echo_debug_graph_csr": (C.c_int, [_P, _I, _I, _P, _P, _P]),
    "echo_graph_create": (C.c_int, [C.POINTER(_P), _P, _I, _I, _P]),
    "echo_graph_destroy": (None, [_P]),
    "echo_gather_rows": (C.c_int, [_P, _P, _L, _L, _L, _P, _P]),
    "echo_gcn_create": (C.c_int, [C.POINTER(_P), C.POINTER(GcnDesc), C.POINTER(Weight), _I]),
    "echo_gcn_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "echo_gcn_forward_train": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "echo_gcn_train_create": (C.c_int, [C.POINTER(_P), C.POINTER(GcnDesc), C.POINTER(Weight), _I, C.POINTER(Weight), _I]),
    "echo_gcn_train_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "echo_gcn_train_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "echo_gcn_train_destroy": (None, [_P]),
    "echo_mesh_workspace_bytes": (C.c_int64, [_I]),
    "echo_mesh_marching_cubes": (C.c_int, [_P, _I, C.c_float, _P, _L, _P, _L, _P, _P, _L, _P]),
    "echo_layout_set_batch_stats": (C.c_int, [_P, _I]),
    "echo_shape_set_batch_stats": (C.c_int, [_P, _I]),
    "echo_scene_set_batch_stats": (C.c_int, [_P, _I]),
    "echo_gcn_destroy": (None, [_P]),
    "echo_layout_create": (C.c_int, [C.POINTER(_P), C.POINTER(LayoutDesc), C.POINTER(Weight), _I]),
    "echo_layout_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "echo_layout_step": (C.c_int, [_P, _P, _P, _P, _I, _P, _P, _P]),
    "echo_layout_destroy": (None, [_P]),
    "echo_layout_schedule": (C.c_int, [_P, _P]),
    "echo_shape_create": (C.c_int, [C.POINTER(_P), C.POINTER(ShapeDesc), C.POINTER(Weight), _I]),
    "echo_shape_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "echo_shape_step": (C.c_int, [_P, _P, _P, _P, _I, _P, _P]),
    "echo_shape_set_index": (C.c_int, [_P, _I, _P]),
    "echo_metrics_validate_constraints": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _I, _P, _I, _P, _I, _I, C.c_float, _P, _P, _P]),
    "echo_train_q_sample": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, _P, _P]),
    "echo_train_mse_rows": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P, _I, _P, _P]),
    "echo_optimizer_create": (C.c_int, [_P, _P, _I]),
    "echo_optimizer_set_tensors": (C.c_int, [_P, _P, _I, _P]),
    "echo_optimizer_step": (C.c_int, [_P, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _P]),
    "echo_optimizer_info": (C.c_int, [_P, _P, _P, _P]),
    "echo_optimizer_destroy": (None, [_P]),
    "echo_shape_embed": (C.c_int, [_P, _P, _I, _P, _P]),
    "echo_shape_trunk": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _P, _I, _P, _P]),
    "echo_shape_trunk_async": (C.c_int, [_P, _P, _P, _I, _I, _P, _P, _P, _I, _P, _P, _P]),
    "echo_shape_latent": (C.c_int, [_P, _I, _P, _P]),
    "echo_shape_destroy": (None, [_P]),
    "echo_shape_schedule": (C.c_int, [_P, _P, _P]),
    "echo_scene_create": (C.c_int, [C.POINTER(_P), C.POINTER(SceneDesc), C.POINTER(Weight), _I]),
    "echo_scene_init_encoder": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "echo_scene_manipulate": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "echo_scene_rel_s": (C.c_int, [_P, _P, _I, _P, _P]),
    "echo_scene_encode": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "echo_scene_destroy": (None, [_P]),
    "echo_vqvae_create": (C.c_int, [C.POINTER(_P), C.POINTER(VqvaeDesc), C.POINTER(Weight), _I]),
    "echo_vqvae_decode": (C.c_int, [_P, _P, _I, _P, _P, _P]),
    "echo_vqvae_encoder_create": (C.c_int, [C.POINTER(_P), C.POINTER(VqvaeDesc), C.POINTER(Weight), _I]),
    "echo_vqvae_encode": (C.c_int, [_P, _P, _I, _P, _P]),
    "echo_vqvae_destroy": (None, [_P]),
    "echo_op_conv3d": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P]),
    "echo_op_upconv3d": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P]),
    "echo_op_upconv3d_x2": (C.c_int, [_P, _I, _I, _I, _I, _I, _P, _P, _I, _P, _I, _P]),
    "echo_op_linear": (C.c_int, [_P, _L, _I, _P, _P, _I, _P, _I, _P]),
    "echo_op_group_norm": (C.c_int, [_P, _I, _L, _I, _I, _P, _P, C.c_float, _I, _P, _P]),
    "echo_op_layer_norm": (C.c_int, [_P, _L, _I, _P, _P, C.c_float, _P, _P]),
    "echo_op_attention": (C.c_int, [_P, _I, _I, _I, _I, _P, _I, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Loads the shared library once.  Raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EchoError(f"{LIB_PATH} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback for the denoiser hot path)")
        # load torch's CUDA runtime first so both sides share the primary context / libcudart
        torch.cuda.is_available()
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(code: int) -> None:
    if code != 0:
        msg = lib().echo_last_error()
        raise EchoError(f"libechoscene_b200 error {code}: {msg.decode() if msg else '?'}")


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise EchoError("echoscene_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on "
                            f"{t.device}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: torch.Tensor) -> int:
    return 0 if t is None else t.data_ptr()


def weights_table(sd: Dict[str, torch.Tensor], extra: Dict[str, torch.Tensor] = None
                  ) -> Tuple[C.Array, int, List]:
    """state_dict (CUDA tensors) -> echo_weight_t[].  Returns (array, n, keepalive)."""
    items = list(sd.items()) + list((extra or {}).items())
    arr = (Weight * len(items))()
    keep = []
    for i, (k, v) in enumerate(items):
        if v.dtype == torch.float32:
            dt = 0
        elif v.dtype == torch.int64:
            dt = 1
        else:
            raise EchoError(f"weight {k}: unsupported dtype {v.dtype} (the reference path is fp32)")
        require_cuda(v)
        v = v.detach().contiguous()
        name = k.encode()
        keep.append((name, v))
        arr[i].name = name
        arr[i].data = v.data_ptr()
        arr[i].ndim = v.dim()
        arr[i].dtype = dt
        for j, s in enumerate(v.shape):
            arr[i].shape[j] = s
    return arr, len(items), keep


def timestep_freqs(dim: int, device) -> torch.Tensor:
    """The frequency table of timestep_embedding, computed exactly as the reference computes it
    (ldm_diffusion_util.py:186-189) so that the CUDA embedding sees bit-identical frequencies."""
    import math
    half = dim // 2
    f = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    return f.to(device)


class Graph:
    """CSR of a scene graph built once from `triples` (T,3) int64 [s,p,o]; edges are constant over a chain."""

    def __init__(self, triples: torch.Tensor, n_nodes: int):
        require_cuda(triples)
        if triples.dtype != torch.int64 or triples.dim() != 2 or triples.shape[1] != 3:
            raise EchoError(f"triples must be (T,3) int64, got {tuple(triples.shape)} {triples.dtype}")
        self.source = triples                 # the cache key is the caller tensor's storage: keep it alive, its address
        self.triples = triples.contiguous()   # cannot be recycled for another graph while this entry exists
        self.n_nodes = int(n_nodes)
        self.n_triples = int(triples.shape[0])
        h = _P()
        check(lib().echo_graph_create(C.byref(h), ptr(self.triples), self.n_triples, self.n_nodes, stream_ptr()))
        self.h = h

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().echo_graph_destroy(self.h)
                self.h = None
        except Exception:
            pass


_graph_cache: Dict[tuple, Graph] = {}


def graph_for(triples: torch.Tensor, n_nodes: int) -> Graph:
    """Graph handles are cached on (storage, shape, version): building one costs a device->host copy."""
    key = (triples.data_ptr(), tuple(triples.shape), triples._version, int(n_nodes), triples.device.index)
    g = _graph_cache.get(key)
    if g is None:
        if len(_graph_cache) > 64:
            _graph_cache.clear()
        g = Graph(triples, n_nodes)
        _graph_cache[key] = g
    return g


_edge_cache: Dict[tuple, Graph] = {}


def graph_for_edges(edges: torch.Tensor, n_nodes: int) -> Graph:
    """Graph handle of an (T,2) [s,o] edge list, cached on the EDGE tensor (a GraphTripleConvNet sees the same `edges` in every
    layer and every step; building a handle costs a device->host copy, a stream sync and five allocations)."""
    key = (edges.data_ptr(), tuple(edges.shape), edges._version, int(n_nodes), edges.device.index)
    g = _edge_cache.get(key)
    if g is None:
        if len(_edge_cache) > 64:
            _edge_cache.clear()
        g = Graph(edges_to_triples(edges), n_nodes)
        g.source = edges   # the key is this tensor's storage
        _edge_cache[key] = g
    return g


def edges_to_triples(edges: torch.Tensor) -> torch.Tensor:
    """edges (T,2) [s,o] -> triples (T,3) [s,0,o] for the graph handle (predicate ids are not used by the GCN)."""
    t = torch.zeros(edges.shape[0], 3, dtype=torch.int64, device=edges.device)
    t[:, 0] = edges[:, 0]
    t[:, 2] = edges[:, 1]
    return t
