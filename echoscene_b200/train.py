"""Training-side pieces of the reference's train step, on the CUDA path:

  * ``q_sample`` and the diffusion losses given a denoiser output (diffusion_ddpm.py:191-201, 451-477; echo2shape.py:254-258,
    297-331), the training tables as the reference's constructors build them (``layout_train_tables`` / ``shape_train_tables``), the
    LambdaLR schedule (``lr_lambda`` / ``learning_rate``) -- what ``scene.Sg2ScDiffModel.forward`` / ``SGDiff.forward_mani`` use to
    compute the training forward's loss VALUES;
  * ``GraphTripleConvNetTrainer``: forward AND backward of a GraphTripleConvNet in training mode (the first backward of this path;
    the backward of the two denoiser trunks is not built), also as one replayed CUDA graph;
  * ``FusedAdamW``: the optimizer step of ``scripts/train_3dfront.py:247-259`` (clip_grad_norm_ of the shape denoiser, the
    per-parameter NaN scrub loop, ``optimizerFULL.step()`` = AdamW) as one fused multi-tensor pass with no host synchronisation; it
    accepts gradients from any producer (torch autograd, or the trainer above).

No CPU fallback: CPU tensors raise ``EchoError``."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Sequence

import torch

from . import _lib
from ._lib import EchoError


def q_sample(x_start: torch.Tensor, t: torch.Tensor, noise: torch.Tensor, sqrt_alphas_cumprod: torch.Tensor,
             sqrt_one_minus_alphas_cumprod: torch.Tensor) -> torch.Tensor:
    """GaussianDiffusion.q_sample (diffusion_ddpm.py:191-201) / EchoToShape.q_sample (echo2shape.py:254-258): one timestep per
    leading-dimension row, any trailing shape (boxes (B, 8), latents (B, 3, 16, 16, 16))."""
    _lib.require_cuda(x_start, t, noise, sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod)
    if noise.shape != x_start.shape or t.shape != (x_start.shape[0],):
        raise EchoError(f"q_sample: x_start {tuple(x_start.shape)}, noise {tuple(noise.shape)}, t {tuple(t.shape)}")
    x0, nz = x_start.float().contiguous(), noise.float().contiguous()
    t = t.to(torch.int64).contiguous()
    a, b = sqrt_alphas_cumprod.float().contiguous(), sqrt_one_minus_alphas_cumprod.float().contiguous()
    out = torch.empty_like(x0)
    rows = x0.shape[0]
    _lib.check(_lib.lib().echo_train_q_sample(_lib.ptr(x0), _lib.ptr(nz), _lib.ptr(t), _lib.ptr(a), _lib.ptr(b), rows,
                                              x0.numel() // max(rows, 1), _lib.ptr(out), _lib.stream_ptr()))
    return out


def mse_rows(pred: torch.Tensor, target: torch.Tensor, ranges: Sequence[Sequence[int]]) -> torch.Tensor:
    """out[r, k] = mean((target - pred)[r, ranges[k][0]:ranges[k][1]] ** 2) over the flattened trailing dimensions."""
    _lib.require_cuda(pred, target)
    if pred.shape != target.shape:
        raise EchoError(f"mse_rows: {tuple(pred.shape)} vs {tuple(target.shape)}")
    p, q = pred.float().contiguous(), target.float().contiguous()
    rows = p.shape[0]
    row_len = p.numel() // max(rows, 1)
    flat = (C.c_int32 * (2 * len(ranges)))(*[int(v) for r in ranges for v in r])
    out = torch.empty(rows, len(ranges), device=p.device)
    _lib.check(_lib.lib().echo_train_mse_rows(_lib.ptr(p), _lib.ptr(q), rows, row_len, flat, len(ranges), _lib.ptr(out), _lib.stream_ptr()))
    return out


def layout_diffusion_loss(denoise_out: torch.Tensor, target: torch.Tensor, size_dim: int = 3, translation_dim: int = 3,
                          angle_dim: int = 2):
    """GaussianDiffusion.diffusion_loss with loss_iou = False (diffusion_ddpm.py:451-477): -> (loss, loss_dict with the reference's
    keys).  Boxes are (B, size | translation | sin, cos)."""
    bbox = size_dim + translation_dim + angle_dim
    D = denoise_out.shape[1]
    parts = mse_rows(denoise_out, target, [(0, size_dim), (size_dim, size_dim + translation_dim), (size_dim + translation_dim, bbox),
                                            (0, bbox), (0, D)])
    m = parts.mean(dim=0)
    zero = torch.zeros((), device=parts.device)
    return m[4] + zero, {"loss.bbox": m[3], "loss.trans": m[1], "loss.size": m[0], "loss.angle": m[2], "loss.liou": zero, "loss.bbox_iou": zero}


def shape_diffusion_loss(model_output: torch.Tensor, target: torch.Tensor, t: torch.Tensor, logvar: torch.Tensor,
                         lvlb_weights: torch.Tensor, l_simple_weight: float = 1.0, original_elbo_weight: float = 0.0):
    """The loss terms of EchoToShape.p_losses (echo2shape.py:297-331) for the eps parameterisation."""
    loss_simple = mse_rows(model_output, target, [(0, model_output[0].numel())])[:, 0]
    logvar_t = logvar.to(loss_simple.device)[t]
    loss = l_simple_weight * (loss_simple / torch.exp(logvar_t) + logvar_t).mean()
    loss_vlb = (lvlb_weights.to(loss_simple.device)[t] * loss_simple).mean()
    total = loss + original_elbo_weight * loss_vlb
    return total, {"loss_simple": loss_simple.mean(), "loss_vlb": loss_vlb, "loss_total": total.detach().clone()}


def layout_train_tables(time_num: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02):
    """The q_sample tables of GaussianDiffusion.__init__ (diffusion_ddpm.py:133-148, linear get_betas :38-40): float64 betas and
    cumulative product, rounded to fp32 BEFORE the square roots.  -> (sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod) fp32."""
    import numpy as np
    betas = np.linspace(beta_start, beta_end, time_num).astype(np.float64)
    ac = torch.from_numpy(np.cumprod(1.0 - betas, axis=0)).float()
    return torch.sqrt(ac).float(), torch.sqrt(1.0 - ac).float()


def shape_train_tables(timesteps: int = 1000, linear_start: float = 0.00085, linear_end: float = 0.012, v_posterior: float = 0.0):
    """The tables EchoToShape.register_schedule builds for training (echo2shape.py:173-226; linear make_beta_schedule,
    ldm_diffusion_util.py:44-47): float64 arithmetic, rounded to fp32 at the end; lvlb_weights for the eps parameterisation
    (computed from the fp32 tables, as the reference does) with entry 0 overwritten by entry 1; logvar = 0 (:167-168).
    -> dict(sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod, lvlb_weights, logvar)."""
    import numpy as np
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, timesteps, dtype=torch.float64) ** 2).numpy()
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    f32 = lambda a: torch.tensor(a, dtype=torch.float32)
    post_var = (1 - v_posterior) * betas * (1.0 - ac_prev) / (1.0 - ac) + v_posterior * betas
    lvlb = f32(betas) ** 2 / (2 * f32(post_var) * f32(alphas) * (1 - f32(ac)))
    lvlb[0] = lvlb[1]
    return {"sqrt_alphas_cumprod": f32(np.sqrt(ac)), "sqrt_one_minus_alphas_cumprod": f32(np.sqrt(1.0 - ac)),
            "lvlb_weights": lvlb, "logvar": torch.full((timesteps,), 0.0)}


def lr_lambda(counter: int, lr_init: float = 1e-4, lr_step: Sequence[int] = (35000, 70000, 140000),
              lr_evo: Sequence[float] = (5e-5, 1e-5, 5e-6)) -> float:
    """Sg2ScDiffModel.lr_lambda (model/EchoScene.py:115-128; hyper.lr_init / lr_step / lr_evo of config/full_mp.yaml): the factor the
    reference's LambdaLR applies to the optimizer's base rate at iteration ``counter``."""
    if counter < lr_step[0]:
        return 1.0
    if counter < lr_step[1]:
        return lr_evo[0] / lr_init
    if counter < lr_step[2]:
        return lr_evo[1] / lr_init
    return lr_evo[2] / lr_init


def learning_rate(counter: int, base_lr: float = 1e-4, **schedule) -> float:
    """The rate ``optimizerFULL`` runs at after ``counter`` scheduler steps: AdamW's base rate (1e-4, EchoScene.py:134) times
    ``lr_lambda(counter)`` -- what ``update_learning_rate`` (EchoScene.py:138-141) leaves in ``param_groups[0]['lr']``; pass it to
    ``FusedAdamW.step(lr=...)``."""
    return base_lr * lr_lambda(counter, **schedule)


class GraphTripleConvNetTrainer:
    """Training-mode forward AND backward of a ``modules.GraphTripleConvNet`` / ``GraphTripleConv`` on the CUDA path
    (``echo_gcn_train_*``, csrc/gcn_train.cu): what the reference gets from autograd over model/graph.py:124-211 under
    ``model.train()`` -- the first block of the backward pass.

        tr = GraphTripleConvNetTrainer(net)                     # net: BatchNorm1d MLPs (mlp_normalization='batch')
        obj_out, pred_out = tr.forward(obj_vecs, pred_vecs, edges)
        d_obj, d_pred = tr.backward(d_obj_out, d_pred_out)     # every net parameter's .grad += its gradient
        FusedAdamW(net.parameters()).step()                     # parameters are read in place: no rebuild after the step

    The parameters' ``.grad`` tensors are views of ONE flat buffer owned by the trainer (``zero_grad`` is a single memset instead of
    a launch per parameter) and are accumulated into, as autograd does; zero them between iterations.  The forward updates
    running_mean / running_var / num_batches_tracked as torch does.  Deterministic: no float atomics anywhere."""

    def __init__(self, net, max_nodes: int = 64, max_triples: int = 256):
        if getattr(net.cfg, "mlp_normalization", None) != "batch":
            raise EchoError("GraphTripleConvNetTrainer needs BatchNorm1d MLPs (mlp_normalization='batch', every construction on the hot path)")
        self.net = net
        self.cap = (int(max_nodes), int(max_triples))
        self._handle, self._key, self._graph = None, None, None
        params = list(net.parameters())
        _lib.require_cuda(*params)
        self._flat = torch.zeros(sum(p.numel() for p in params), device=params[0].device)
        off = 0
        for p in params:       # existing gradients are carried over into the flat buffer
            view = self._flat[off:off + p.numel()].view_as(p)
            if p.grad is not None:
                view.copy_(p.grad)
            p.grad = view
            off += p.numel()

    def _tables(self):
        sd = self.net.state_dict_for_lib()
        for k, v in sd.items():
            if not v.is_contiguous():
                raise EchoError(f"GraphTripleConvNetTrainer: parameter {k} must be contiguous (it is read in place)")
        pre = "gconvs.0." if type(self.net).__name__ == "GraphTripleConv" else ""
        grads = {}
        for k, p in self.net.named_parameters():
            if p.grad is None:
                p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
            if p.grad.dtype != torch.float32 or not p.grad.is_contiguous():
                raise EchoError(f"GraphTripleConvNetTrainer: gradient of {k} must be contiguous fp32")
            grads[pre + k] = p.grad
        return sd, grads

    def _ensure(self, n, t):
        sd, grads = self._tables()
        key = (tuple(v.data_ptr() for v in sd.values()), tuple(g.data_ptr() for g in grads.values()))
        if self._handle is not None and key == self._key and n <= self.cap[0] and t <= self.cap[1]:
            return
        self._destroy()
        self.cap = (max(self.cap[0], n), max(self.cap[1], t))
        c = self.net.cfg
        d = _lib.GcnDesc(c.input_dim_obj, c.input_dim_pred, c.num_layers, c.hidden_dim, c.output_dim or 0, self.cap[0], self.cap[1], 1e-5, 0)
        pa, np_, keep1 = _lib.weights_table(sd)
        ga, ng, keep2 = _lib.weights_table(grads)
        h = C.c_void_p()
        _lib.check(_lib.lib().echo_gcn_train_create(C.byref(h), C.byref(d), pa, np_, ga, ng))
        self._handle, self._key = h, key

    def _destroy(self):
        if self._handle is not None:
            _lib.lib().echo_gcn_train_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    def zero_grad(self):
        self._flat.zero_()

    @torch.no_grad()
    def forward(self, obj_vecs, pred_vecs, edges):
        _lib.require_cuda(obj_vecs, pred_vecs, edges)
        c = self.net.cfg
        obj_vecs, pred_vecs = obj_vecs.float().contiguous(), pred_vecs.float().contiguous()
        n, t = obj_vecs.shape[0], pred_vecs.shape[0]
        if obj_vecs.shape[1] != c.input_dim_obj or pred_vecs.shape[1] != c.input_dim_pred or tuple(edges.shape) != (t, 2):
            raise EchoError(f"GraphTripleConvNetTrainer.forward: obj {tuple(obj_vecs.shape)}, pred {tuple(pred_vecs.shape)}, edges {tuple(edges.shape)}")
        self._ensure(n, t)
        self._graph = _lib.graph_for_edges(edges, n)
        obj_out = torch.empty(n, c.output_dim or c.input_dim_obj, device=obj_vecs.device)
        pred_out = torch.empty(t, c.input_dim_pred, device=obj_vecs.device)
        _lib.check(_lib.lib().echo_gcn_train_forward(self._handle, self._graph.h, _lib.ptr(obj_vecs), _lib.ptr(pred_vecs), _lib.ptr(obj_out),
                                                     _lib.ptr(pred_out), _lib.stream_ptr()))
        return obj_out, pred_out

    def capture(self, obj_vecs, pred_vecs, edges, d_obj_out, d_pred_out=None):
        """One training iteration -- forward + backward on fixed-size inputs -- as ONE replayed CUDA graph (the ~300 launches of an
        iteration are launch-bound for a single scene).  The four tensors are copied into static buffers owned by the returned
        object; ``replay(obj_vecs, pred_vecs, d_obj_out, d_pred_out)`` copies new values in and replays; its results
        (``obj_out, pred_out, d_obj, d_pred``) are static tensors overwritten by the next replay.  Gradients accumulate into ``.grad``
        exactly as with ``forward`` / ``backward``; the graph topology (``edges``) is fixed."""
        _lib.require_cuda(obj_vecs, pred_vecs, edges, d_obj_out, d_pred_out)
        trainer = self

        class Iteration:
            def __init__(it):
                it.obj, it.pred = obj_vecs.float().contiguous().clone(), pred_vecs.float().contiguous().clone()
                it.d_obj = d_obj_out.float().contiguous().clone()
                it.d_pred = None if d_pred_out is None else d_pred_out.float().contiguous().clone()
                it.edges = edges.contiguous().clone()
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):            # warm-up outside the capture: handle, graph CSR, lazy module loads
                    for _ in range(2):
                        trainer.forward(it.obj, it.pred, it.edges)
                        trainer.backward(it.d_obj, it.d_pred)
                torch.cuda.current_stream().wait_stream(side)
                trainer.zero_grad()
                it.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(it.graph):
                    it.obj_out, it.pred_out = trainer.forward(it.obj, it.pred, it.edges)
                    it.d_obj_in, it.d_pred_in = trainer.backward(it.d_obj, it.d_pred)
                trainer.zero_grad()                      # (capturing does not execute: nothing was accumulated, keep it explicit)

            def replay(it, obj=None, pred=None, d_obj=None, d_pred=None):
                for dst, src in ((it.obj, obj), (it.pred, pred), (it.d_obj, d_obj), (it.d_pred, d_pred)):
                    if src is not None:
                        dst.copy_(src)
                it.graph.replay()
                return it.obj_out, it.pred_out, it.d_obj_in, it.d_pred_in

        return Iteration()

    @torch.no_grad()
    def backward(self, d_obj_out, d_pred_out=None, need_input_grads: bool = True):
        """cotangents of ``forward``'s two outputs (``d_pred_out`` None = zeros) -> (d_obj_vecs, d_pred_vecs) (None, None when
        ``need_input_grads`` is False); the parameters' ``.grad`` are accumulated into."""
        if self._graph is None:
            raise EchoError("GraphTripleConvNetTrainer.backward: call forward first")
        _lib.require_cuda(d_obj_out, d_pred_out)
        c, g = self.net.cfg, self._graph
        n, t = d_obj_out.shape[0], (d_pred_out.shape[0] if d_pred_out is not None else None)
        d_obj_out = d_obj_out.float().contiguous()
        if d_obj_out.shape[1] != (c.output_dim or c.input_dim_obj):
            raise EchoError(f"backward: d_obj_out {tuple(d_obj_out.shape)}")
        if d_pred_out is not None:
            d_pred_out = d_pred_out.float().contiguous()
            if d_pred_out.shape[1] != c.input_dim_pred:
                raise EchoError(f"backward: d_pred_out {tuple(d_pred_out.shape)}")
        self._ensure(n, t or 0)   # same tensors as in forward: no rebuild; a rebuilt handle refuses (nothing saved)
        d_obj = d_pred = None
        if need_input_grads:
            d_obj = torch.empty(n, c.input_dim_obj, device=d_obj_out.device)
            d_pred = torch.empty(g.n_triples, c.input_dim_pred, device=d_obj_out.device)
        _lib.check(_lib.lib().echo_gcn_train_backward(self._handle, g.h, _lib.ptr(d_obj_out), _lib.ptr(d_pred_out), _lib.ptr(d_obj),
                                                      _lib.ptr(d_pred), _lib.stream_ptr()))
        return d_obj, d_pred


class FusedAdamW:
    """``optimizerFULL`` of the reference (model/EchoScene.py:130-136: AdamW over the encoders, the layout denoiser and the shape
    denoiser) with the step sequence of scripts/train_3dfront.py:247-259 fused into one pass:

        clip_grad_norm_(clip_params, clip_max_norm)          # :251, the shape denoiser's parameters
        for p in params: p.grad[isnan(p.grad)] = 0           # :252-256, ~700 `.any()` host synchronisations in the reference
        optimizer.step()                                      # :258

    ``params`` must have ``.grad`` set (fp32, CUDA, contiguous); the AdamW state lives in this object (``state_dict`` /
    ``load_state_dict`` use torch.optim.AdamW's layout so the reference's checkpoints carry over)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, clip_params: Optional[Iterable[torch.nn.Parameter]] = None, clip_max_norm: float = 5.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise EchoError("FusedAdamW: no trainable parameters")
        _lib.require_cuda(*self.params)
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.clip_ids = {id(p) for p in (clip_params or [])}
        self.clip_max_norm = float(clip_max_norm) if self.clip_ids else 0.0
        self.exp_avg = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p, memory_format=torch.contiguous_format) for p in self.params]
        self._handle, self._key = None, None
        self.steps = 0

    def _ensure(self):
        for p in self.params:
            if p.grad is None:
                raise EchoError("FusedAdamW.step: a parameter has no gradient (the reference skips such parameters; allocate zeros)")
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                raise EchoError("FusedAdamW: parameters and gradients must be contiguous fp32")
        key = tuple((p.data_ptr(), p.grad.data_ptr()) for p in self.params)
        if self._handle is not None and key == self._key:
            return
        arr = (_lib.OptTensor * len(self.params))()
        for i, p in enumerate(self.params):
            arr[i].param, arr[i].grad = p.data_ptr(), p.grad.data_ptr()
            arr[i].exp_avg, arr[i].exp_avg_sq = self.exp_avg[i].data_ptr(), self.exp_avg_sq[i].data_ptr()
            arr[i].numel, arr[i].clip_group = p.numel(), int(id(p) in self.clip_ids)
        if self._handle is None:
            h = C.c_void_p()
            _lib.check(_lib.lib().echo_optimizer_create(C.byref(h), arr, len(self.params)))
            self._handle = h
        else:   # fresh gradient tensors (zero_grad(set_to_none=True)): same list, new addresses
            _lib.check(_lib.lib().echo_optimizer_set_tensors(self._handle, arr, len(self.params), _lib.stream_ptr()))
        self._key = key

    def _destroy(self):
        if self._handle is not None:
            _lib.lib().echo_optimizer_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._destroy()
        except Exception:
            pass

    @torch.no_grad()
    def step(self, lr: Optional[float] = None):
        """One fused step (three launches; no host synchronisation unless the gradient tensors moved).  ``lr``: this step's learning rate (the reference's LambdaLR
        schedule, EchoScene.py:115-128, is evaluated by the caller)."""
        self._ensure()
        self.steps += 1
        _lib.check(_lib.lib().echo_optimizer_step(self._handle, self.steps, float(self.lr if lr is None else lr), self.betas[0], self.betas[1],
                                                  self.eps, self.weight_decay, self.clip_max_norm, _lib.stream_ptr()))

    def info(self) -> Dict[str, float]:
        """{steps, parameters, chunks, nan_gradients_scrubbed, last_grad_norm, last_clip_coef}; synchronises the stream."""
        self._ensure()
        out, clip = (C.c_int64 * 4)(), (C.c_float * 2)()
        _lib.check(_lib.lib().echo_optimizer_info(self._handle, out, clip, _lib.stream_ptr()))
        return {"steps": out[0], "parameters": out[1], "chunks": out[2], "nan_gradients_scrubbed": out[3], "last_grad_norm": clip[0],
                "last_clip_coef": clip[1]}

    def zero_grad(self):
        for p in self.params:
            if p.grad is not None:
                p.grad.zero_()
