/*
 * echoscene_b200.h — C ABI of libechoscene_b200.so: the B200 (sm_100a) denoiser hot path of EchoScene.
 *
 * The reference (ymxlzgy/echoscene) has no FFI on this path: its seam is the Python class surface
 * (GraphTripleConv / GraphTripleConvNet / UNet1DModel / UNet3DModel / DDIMSampler / DiffusionPoint).
 * Each entry point below names the reference interface it replaces (file:line under the reference root).
 * The only native precedent in the reference is extension/old_chamfer/chamfer_cuda.cpp:17-32 (returns int,
 * printf on error, default stream); this ABI keeps "returns int" and fixes the rest:
 *
 *   - every call returns 0 on success and a negative code on failure; echo_last_error() gives the
 *     thread-local message; nothing prints, aborts or synchronises the device;
 *   - all tensor arguments are DEVICE pointers owned by the caller, dense row-major, fp32 / int64 exactly
 *     as the reference's torch tensors are laid out (NCDHW for the latent);
 *   - every compute call takes an explicit stream (a cudaStream_t passed as void*) and is asynchronous;
 *   - handles own repacked weights and workspace; no allocation happens after *_create;
 *   - a handle is re-entrant but not thread-safe; distinct handles may be used from distinct threads.
 *
 * No torch / libtorch types appear here: the library is loadable with ctypes, cgo, JNI, ...
 */
#ifndef ECHOSCENE_B200_H
#define ECHOSCENE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECHO_ABI_VERSION 1

#if defined(__GNUC__)
#define ECHO_API __attribute__((visibility("default")))
#else
#define ECHO_API
#endif

/* error codes */
#define ECHO_OK 0
#define ECHO_ERR_INVALID (-1)   /* bad argument / shape / missing weight */
#define ECHO_ERR_CUDA (-2)      /* CUDA runtime or driver error */
#define ECHO_ERR_NOMEM (-3)     /* workspace exhausted */
#define ECHO_ERR_UNSUPPORTED (-4)

/* arithmetic of the dense contractions */
#define ECHO_PREC_FP32 0        /* fp32 operands, fp32 FMA: the 1e-3 parity mode */
#define ECHO_PREC_BF16 1        /* bf16 operands on tcgen05 tensor cores, fp32 accumulate, fp32 norms */
#define ECHO_PREC_X3 2          /* shape branch: fp32 activations; every contraction on tcgen05 with both operands split into hi + lo bf16
                                 * halves and three MMAs per k-step into one fp32 TMEM accumulator (hi.hi + lo.hi + hi.lo): fp32-grade
                                 * results (~1e-5) -- the 1e-3 parity contract on the tensor cores */

typedef struct echo_graph echo_graph_t;     /* CSR of one (batched) scene graph; edges are constant over a chain */
typedef struct echo_gcn echo_gcn_t;         /* GraphTripleConvNet */
typedef struct echo_layout echo_layout_t;   /* UNet1DModel + DDPM schedule */
typedef struct echo_shape echo_shape_t;     /* UNet3DModel + DDIM schedule */
typedef struct echo_scene echo_scene_t;     /* once-per-scene encoders of Sg2ScDiffModel */

/* One named parameter / buffer of a reference state_dict (device pointer, fp32 or int64, reference layout). */
typedef struct echo_weight {
  const char* name;     /* state_dict key, e.g. "input_blocks.4.1.transformer_blocks.0.attn1.to_v.weight" */
  const void* data;     /* device pointer */
  int32_t ndim;
  int32_t dtype;        /* 0 = float32, 1 = int64 (BatchNorm num_batches_tracked; ignored) */
  int64_t shape[6];
} echo_weight_t;

/* GraphTripleConvNet(input_dim_obj, input_dim_pred, num_layers, hidden_dim, output_dim) — model/graph.py:214-244 */
typedef struct echo_gcn_desc {
  int32_t input_dim_obj, input_dim_pred, num_layers, hidden_dim;
  int32_t output_dim;   /* <= 0: same as input_dim_obj for every layer */
  int32_t max_nodes, max_triples;
  float bn_eps;         /* 1e-5, model/layers.py:29-30 */
  int32_t keep_train_weights;   /* != 0: also keep the unfolded Linear / BatchNorm1d tensors, for echo_gcn_forward_train */
} echo_gcn_desc_t;

/* UNet1DModel(**denoiser_kwargs) — config/full_mp.yaml:24-39, denoise_net.py:451-756 */
typedef struct echo_layout_desc {
  int32_t in_channels, out_channels, model_channels;
  int32_t num_levels;
  int32_t channel_mult[8];
  int32_t num_res_blocks;
  int32_t num_attention_resolutions;
  int32_t attention_resolutions[8];
  int32_t num_heads;
  int32_t context_dim;          /* crossattn_dim = concat_dim = 1280 */
  int32_t obj_embed_dim;        /* 640 */
  int32_t gconv_dim;            /* 64 */
  int32_t enable_t_emb;
  int32_t max_nodes, max_triples;
  int32_t precision;            /* ECHO_PREC_* */
  /* DDPM schedule — diffusion_ddpm.py:38-40,133-162 */
  int32_t time_num;
  float beta_start, beta_end;
  int32_t keep_train_weights;   /* != 0: the GCN MLPs (and rel_s_mlp) also keep their unfolded tensors, for the batch-statistics forward
                                  (echo_*_set_batch_stats) */
} echo_layout_desc_t;

/* UNet3DModel(**unet.params) — config/sdfusion-txt2shape_mp.yaml:16-41, openai_model_3d.py:452-782 */
typedef struct echo_shape_desc {
  int32_t in_channels, out_channels, model_channels;
  int32_t num_levels;
  int32_t channel_mult[8];
  int32_t num_res_blocks;
  int32_t num_attention_resolutions;
  int32_t attention_resolutions[8];
  int32_t num_heads;
  int32_t context_dim;          /* 1280 */
  int32_t gconv_dim;            /* 64 */
  int32_t enable_t_emb;
  int32_t latent_size;          /* 16: latent is (C,16,16,16) */
  int32_t max_nodes, max_triples;
  int32_t max_local_nodes;      /* objects whose trunk runs on this GPU (== max_nodes on one GPU) */
  int32_t precision;            /* ECHO_PREC_* */
  /* DDIM schedule — ldm_diffusion_util.py:43-47,68-96; samplers/ddim.py:28-57 */
  int32_t timesteps;            /* 1000 */
  int32_t ddim_steps;           /* S */
  float linear_start, linear_end;
  int32_t keep_train_weights;   /* != 0: the GCN MLPs (and rel_s_mlp) also keep their unfolded tensors, for the batch-statistics forward
                                  (echo_*_set_batch_stats) */
} echo_shape_desc_t;

ECHO_API int echo_version(void);
ECHO_API const char* echo_last_error(void);
/* 1 if the library was built with the sm_100a tcgen05/TMA kernels AND the current device can run them. */
ECHO_API int echo_has_tcgen05(void);
/* kernels launched by this library on the calling thread since the last reset (bench.py's gpu_launches). */
ECHO_API int64_t echo_launch_count(void);
ECHO_API void echo_launch_count_reset(void);
/* tuning/profiling knob of the tcgen05 GEMM: 0 = automatic, 1 = one CTA per 128-row tile, 2 = CTA pairs (cta_group::2). */
ECHO_API void echo_debug_set_tc_mode(int mode);
/* Measurement aid (bench.py roofline): between begin and end, every tcgen05 contraction launch with this output row
 * count / cin / cout / kernel size is bracketed by CUDA events on its launching stream; end returns the number of
 * launches seen and their average duration in milliseconds (synchronise the stream first). */
ECHO_API void echo_debug_probe_begin(int64_t rows, int32_t cin, int32_t cout, int32_t ksize);
ECHO_API int32_t echo_debug_probe_end(double* avg_ms);
/* The next probed launch writes the timeline of its first CTA (tag << 32 | tile, globaltimer ns pairs after a count word)
 * into buf_dev (device memory, >= 8001 x 8 bytes, zeroed by the caller): where a tile's time goes (tools/gemm_timeline.py). */
ECHO_API void echo_debug_probe_timeline(void* buf_dev);
/* Host-only: the launch plan of the tcgen05 contraction kernel for an (n, d, h, w, cin) -> cout problem on `sms` SMs:
 * out4 = {tile width, 128-row sub-blocks per CTA, split-K factor, CTA pairs}.  epi = 1: GEGLU epilogue; up2 = 1 / 2: conv after a
 * nearest x(1,2,2) / x2 upsample; allow_splitk: a split-K workspace is offered. */
ECHO_API void echo_debug_tc_plan(int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin, int32_t cout, int32_t ksize, int32_t epi,
                        int32_t up2, int32_t allow_splitk, int32_t sms, int32_t* out4);
/* Host-only (no GPU needed): the weight fold behind the upsample-folded convolutions.  w_host [cout][27 taps (kd,kh,kw)][cin]
 * -> out_host [cout][4 phases (py,px)][12 taps (kd,a,b)][cin] (up_depth = 0, nearest x(1,2,2)) or [cout][8 phases
 * (pz,py,px)][8 taps (a_d,a_h,a_w)][cin] (up_depth = 1, nearest x2); tap a of phase p along an axis reads low-res offset
 * p - 1 + a.  Exported so the fold can be checked against upsample + conv on the CPU (tests/test_fold_host.py). */
ECHO_API int echo_debug_fold_upsample_weight(const float* w_host, int32_t cout, int32_t cin, int32_t up_depth, float* out_host);

/* ---- graph: edges = stack([s, o]) of `triples` (T,3) int64 [s,p,o] — denoise_net.py:759-761, graph.py:142-143 */
/* host-only: the sampler tables exactly as echo_layout_create / echo_shape_create compute them (same layouts as
 * echo_layout_schedule / echo_shape_schedule below), for CPU tests against the reference's buffers.
 * ddpm: host_out 5 x time_num f32.  ddim: host_coef_out n x 4 f32, host_timesteps_out n i32, *n_out = n <= capacity. */
ECHO_API int echo_debug_ddpm_tables(int32_t time_num, float beta_start, float beta_end, float* host_out);
ECHO_API int echo_debug_ddim_schedule(int32_t timesteps, int32_t ddim_steps, float linear_start, float linear_end, int32_t capacity,
                                      float* host_coef_out, int32_t* host_timesteps_out, int32_t* n_out);
/* host-only: the CSR echo_graph_create builds from host triples (T,3) [s,p,o] -- node_off_out (N+1), node_items_out (2T)
 * with item = 2*t + role (0 subject, 1 object) in the order the reference's scatter_add visits them (graph.py:176-177),
 * pred_range_out = {min p, max p} ({0,-1} when T == 0).  For CPU tests of the index work. */
ECHO_API int echo_debug_graph_csr(const int64_t* triples_host, int32_t n_triples, int32_t n_nodes, int32_t* node_off_out,
                                  int32_t* node_items_out, int64_t* pred_range_out);
ECHO_API int echo_graph_create(echo_graph_t** out, const int64_t* triples_dev, int32_t n_triples, int32_t n_nodes, void* stream);
ECHO_API void echo_graph_destroy(echo_graph_t* g);

/* ---- obj_vecs[idx]: the bit-exact edge-index gather — model/graph.py:146-147.  An index outside [0, n_rows) (torch raises an
 * IndexError there) yields a row of NaNs; foreign memory is never read. */
ECHO_API int echo_gather_rows(const float* obj_vecs, const int64_t* idx, int64_t n_idx, int64_t n_rows, int64_t dim,
                     float* out, void* stream);

/* ---- GraphTripleConvNet.forward(obj_vecs, pred_vecs, edges) (eval mode) — model/graph.py:246-250, 124-211.
 * Weight names are state_dict keys relative to the net ("gconvs.0.net1.0.weight", ...). */
ECHO_API int echo_gcn_create(echo_gcn_t** out, const echo_gcn_desc_t* desc, const echo_weight_t* weights, int32_t n_weights);
ECHO_API int echo_gcn_forward(echo_gcn_t* h, const echo_graph_t* g, const float* obj_vecs, const float* pred_vecs,
                     float* obj_out, float* pred_out, void* stream);
/* The same forward as the reference computes it under model.train() (scripts/train_3dfront.py:237): every BatchNorm1d of the
 * build_mlp stacks (model/layers.py:21-38) normalises with the statistics of the batch -- the rows of the call: triples for
 * net1, nodes for net2 -- (biased variance), not with its running statistics.  Forward only: no autograd tape, running
 * statistics are not updated.  Needs echo_gcn_desc_t.keep_train_weights. */
ECHO_API int echo_gcn_forward_train(echo_gcn_t* h, const echo_graph_t* g, const float* obj_vecs, const float* pred_vecs,
                           float* obj_out, float* pred_out, void* stream);
ECHO_API void echo_gcn_destroy(echo_gcn_t* h);

/* ---- GraphTripleConvNet, training executor: the forward under model.train() that keeps what the backward needs, and the backward
 * the reference gets from autograd over model/graph.py:124-211 + model/layers.py:21-38 (loss.backward(), scripts/train_3dfront.py:247).
 * `params` / `grads`: tables under the GraphTripleConvNet state_dict names ("gconvs.0.net1.0.weight", ...; BatchNorm1d buffers
 * running_mean / running_var / num_batches_tracked optional in `params`).  The handle keeps the POINTERS: parameters are read in
 * place on every call (an optimizer step between two iterations needs no rebuild), gradients are ACCUMULATED in place (+=, as
 * autograd accumulates into .grad), the running statistics are updated by the forward as torch does (momentum 0.1, unbiased
 * variance).  desc->keep_train_weights is ignored; BatchNorm1d MLPs (mlp_normalization = 'batch') are required.
 *   forward : obj_vecs (N, din), pred_vecs (T, dp) -> obj_out (N, dout), pred_out (T, dp); T >= 2 and N >= 2 (torch refuses
 *             BatchNorm1d on a single row in training mode)
 *   backward: cotangents d_obj_out (N, dout), d_pred_out (T, dp) or NULL (= zeros) -> d_obj_in (N, din) / d_pred_in (T, dp), each
 *             optional (NULL: not needed); must follow a forward on the same graph.  Deterministic (no float atomics). */
typedef struct echo_gcn_train echo_gcn_train_t;
ECHO_API int echo_gcn_train_create(echo_gcn_train_t** out, const echo_gcn_desc_t* desc, const echo_weight_t* params, int32_t n_params,
                                   const echo_weight_t* grads, int32_t n_grads);
ECHO_API int echo_gcn_train_forward(echo_gcn_train_t* h, const echo_graph_t* g, const float* obj_vecs, const float* pred_vecs,
                                    float* obj_out, float* pred_out, void* stream);
ECHO_API int echo_gcn_train_backward(echo_gcn_train_t* h, const echo_graph_t* g, const float* d_obj_out, const float* d_pred_out,
                                     float* d_obj_in, float* d_pred_in, void* stream);
ECHO_API void echo_gcn_train_destroy(echo_gcn_train_t* h);

/* ---- layout branch.
 * echo_layout_forward == UNet1DModel.forward(box_t, obj_embed, triples, timesteps, context) — denoise_net.py:773-806;
 *   box_t (N,8) f32, obj_embed (N,640) f32, timesteps (N,) i64 -> eps (N,8) f32 (the reference returns (N,8,1)).
 * echo_layout_step == one iteration of GaussianDiffusion.p_sample_loop_sg — diffusion_ddpm.py:220-264,296-309,330-345:
 *   forward at timestep t for all nodes, eps->x0, posterior mean, + [t>0] exp(0.5 logvar) * noise. */
ECHO_API int echo_layout_create(echo_layout_t** out, const echo_layout_desc_t* desc, const echo_weight_t* weights, int32_t n_weights);
ECHO_API int echo_layout_forward(echo_layout_t* h, const echo_graph_t* g, const float* box_t, const float* obj_embed,
                        const int64_t* timesteps, float* eps_out, void* stream);
ECHO_API int echo_layout_step(echo_layout_t* h, const echo_graph_t* g, const float* x_t, const float* obj_embed, int32_t t,
                     const float* noise, float* x_prev, void* stream);
ECHO_API void echo_layout_destroy(echo_layout_t* h);
/* How echo_layout_step executes (no reference counterpart; diagnostics for tests and bench.py).
 * mode 0 (default): graphs of <= 64 nodes / <= 512 triples run as ONE persistent cooperative kernel (csrc/layout_mk.cu), larger
 * batches as a replayed CUDA graph of the per-layer kernels; mode 1: never the persistent kernel.
 * info out6 = {persistent-kernel steps so far, graph replays so far, stages and ops of the current program, kernels inside the
 * replayed graph, CTAs of the persistent kernel (0: unavailable on this device)}. */
ECHO_API void echo_debug_set_layout_mode(int mode);
ECHO_API int echo_debug_layout_info(const echo_layout_t* h, int64_t* out6);

/* ---- shape branch.
 * echo_shape_forward == UNet3DModel.forward(x, obj_embed, triples, timesteps, context) — openai_model_3d.py:816-863;
 *   x (N,3,16,16,16) f32 NCDHW, obj_embed (N,1,1280) f32 (the `uc_s` conditioning), timesteps (N,) i64 -> e_t like x.
 * echo_shape_step == one iteration of DDIMSampler.ddim_sampling / p_sample_ddim with eta = 0 —
 *   samplers/ddim.py:156-181,184-262: forward at ddim_timesteps[index] and the x_prev update (:252-261).
 * The two halves are exported separately so that a per-object shard can all-gather the 64-d shape codes between
 * them: echo_shape_embed == the `shape_embeddings` stack (openai_model_3d.py:757-764, 805-806) on
 * the local objects; echo_shape_trunk == everything else, with the codes of ALL nodes given and the trunk run on
 * objects [obj_begin, obj_begin + n_local). */
ECHO_API int echo_shape_create(echo_shape_t** out, const echo_shape_desc_t* desc, const echo_weight_t* weights, int32_t n_weights);
ECHO_API int echo_shape_forward(echo_shape_t* h, const echo_graph_t* g, const float* x, const float* obj_embed,
                       const int64_t* timesteps, float* eps_out, void* stream);
ECHO_API int echo_shape_step(echo_shape_t* h, const echo_graph_t* g, const float* x_t, const float* obj_embed, int32_t ddim_index,
                    float* x_prev, void* stream);
ECHO_API int echo_shape_embed(echo_shape_t* h, const float* x_local, int32_t n_local, float* codes_out, void* stream);
ECHO_API int echo_shape_trunk(echo_shape_t* h, const echo_graph_t* g, const float* x_local, int32_t obj_begin, int32_t n_local,
                     const float* codes_all, const float* obj_embed_all, const int64_t* timesteps_all,
                     int32_t ddim_index /* < 0: no sampler update, write e_t */, float* out_local, void* stream);
/* Same, for a sharded step whose all-gather runs on its own stream: `codes_all` is produced by work already queued on
 * `codes_stream` (the embed + NCCL all-gather); only the echo chain (GCN -> cross-attention vectors) waits for it, the
 * trunk's first blocks start at once on `stream` and meet the echo chain at the first SpatialTransformer. */
ECHO_API int echo_shape_trunk_async(echo_shape_t* h, const echo_graph_t* g, const float* x_local, int32_t obj_begin, int32_t n_local,
                           const float* codes_all, const float* obj_embed_all, const int64_t* timesteps_all,
                           int32_t ddim_index, float* out_local, void* codes_stream, void* stream);
/* Training-mode forward values.  Under model.train() the reference's BatchNorm1d layers -- the MLPs of every
 * GraphTripleConvNet (the two scene encoders, box_graph_cov, shape_code_graph_cov) and rel_s_mlp -- normalise with the statistics of
 * the batch (model/layers.py:29-30).  With batch statistics switched on, echo_layout_forward / echo_shape_forward / echo_shape_trunk /
 * echo_scene_init_encoder / _manipulate / _rel_s compute exactly that forward (GroupNorm / LayerNorm do not depend on the mode,
 * dropout is 0 in the reference's configs).  Forward values only: nothing records an autograd tape, running statistics are not
 * updated.  The handle must have been created with keep_train_weights; the sampler steps (echo_layout_step, echo_shape_step)
 * refuse to run while the switch is on. */
ECHO_API int echo_layout_set_batch_stats(echo_layout_t* h, int32_t on);
ECHO_API int echo_shape_set_batch_stats(echo_shape_t* h, int32_t on);
ECHO_API int echo_scene_set_batch_stats(echo_scene_t* h, int32_t on);

/* ---- SDF -> triangle mesh: marching cubes over ONE (R, R, R) fp32 volume, the stage the reference runs per object on
 * the CPU with PyMCubes -- mcubes.marching_cubes(sdf_i, level), model/diff_utils/util_3d.py:213-218 (level 0.02), followed by
 * verts / n_cell - 0.5 (:219, left to the caller).  Vertices are shared per grid edge, in index coordinates (x = first array axis),
 * linearly interpolated along the edge; "inside" is value < level; triangles are oriented with their normals pointing from inside
 * to outside.  Case tables: derived from the method's definition (tools/gen_mc_tables.py); PyMCubes itself is not available here,
 * so vertex / triangle ORDER and the diagonals chosen on ambiguous configurations are this library's, not PyMCubes' (DESIGN.md).
 *   verts  (max_verts, 3) f32, faces (max_faces, 3) i32 vertex indices: written up to the capacities given (0 / NULL: count only);
 *   counts (2,) i32 DEVICE: [number of vertices, number of triangles] of the whole mesh -- when they exceed the capacities the call
 *          must be repeated with larger outputs (typical use: a counting call, one host read, then the emitting call);
 *   workspace: echo_mesh_workspace_bytes(R) bytes of device scratch.  Deterministic: output order depends on the volume only. */
ECHO_API int64_t echo_mesh_workspace_bytes(int32_t resolution);
ECHO_API int echo_mesh_marching_cubes(const float* sdf, int32_t resolution, float level, float* verts, int64_t max_verts, int32_t* faces,
                                      int64_t max_faces, int32_t* counts, void* workspace, int64_t workspace_bytes, void* stream);

/* A chain whose launch sequence does not depend on the step: pass ECHO_INDEX_FROM_DEVICE as `ddim_index` to echo_shape_step /
 * echo_shape_trunk / echo_shape_trunk_async and the step reads its DDIM index (timesteps, update coefficients) from a slot on
 * the device, written by echo_shape_set_index (stream-ordered).  Such a step can be captured ONCE into a CUDA graph -- together
 * with the NCCL all-gather of a sharded step -- and replayed for every iteration of DDIMSampler.ddim_sampling
 * (samplers/ddim.py:156-181): set the index, replay. */
#define ECHO_INDEX_FROM_DEVICE (-2)
ECHO_API int echo_shape_set_index(echo_shape_t* h, int32_t ddim_index, void* stream);
/* copies latent_shape_rel of the last forward/trunk call, (n_nodes, context_dim) f32, into out_dev. */
ECHO_API int echo_shape_latent(const echo_shape_t* h, int32_t n_nodes, float* out_dev, void* stream);
ECHO_API void echo_shape_destroy(echo_shape_t* h);

/* ---- training side: what a train_3dfront.py iteration does around the denoisers' forward / backward.
 * echo_train_q_sample == GaussianDiffusion.q_sample (diffusion_ddpm.py:191-201) / EchoToShape.q_sample (echo2shape.py:254-258):
 *   out[r, :] = sqrt_alphas_cumprod[t[r]] * x0[r, :] + sqrt_one_minus_alphas_cumprod[t[r]] * noise[r, :]   (rows x row_len f32, t int64).
 * echo_train_mse_rows == the loss terms of diffusion_loss (diffusion_ddpm.py:451-462: ((target - out)**2).mean(dim=1) over the
 *   size / translation / angle / whole-box column ranges) and of EchoToShape.p_losses (echo2shape.py:313: mse(reduction='none')
 *   .mean([1,2,3,4])): out[r, k] = mean over columns [ranges[2k], ranges[2k+1]) of (target - pred)^2; `ranges` is a HOST array.
 * echo_optimizer_* == the optimizer step of scripts/train_3dfront.py:247-259 in one pass over the parameters, no host
 *   synchronisation: clip_grad_norm_(shape denoiser, max_norm) (:251) on the tensors flagged clip_group, then the
 *   "isnan(grad).any() -> grad[isnan] = 0" loop over every parameter (:252-256), then optimizerFULL.step() = torch.optim.AdamW
 *   (:258).  Gradients are left clipped and scrubbed in place, as the reference leaves them. */
typedef struct echo_optimizer echo_optimizer_t;
typedef struct {
  float* param;        /* n f32, updated in place */
  float* grad;         /* n f32 */
  float* exp_avg;      /* n f32, AdamW state (zero before the first step) */
  float* exp_avg_sq;   /* n f32 */
  int64_t numel;
  int32_t clip_group;  /* != 0: counted in, and scaled by, the gradient-norm clip */
  int32_t reserved;
} echo_opt_tensor_t;
ECHO_API int echo_train_q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_alphas_cumprod,
                        const float* sqrt_one_minus_alphas_cumprod, int64_t rows, int64_t row_len, float* out, void* stream);
ECHO_API int echo_train_mse_rows(const float* pred, const float* target, int64_t rows, int64_t row_len, const int32_t* ranges_host,
                        int32_t n_ranges, float* out, void* stream);
ECHO_API int echo_optimizer_create(echo_optimizer_t** out, const echo_opt_tensor_t* tensors, int32_t n_tensors);
/* the same tensor list at new addresses (torch's zero_grad(set_to_none=True) allocates fresh gradients every iteration) */
ECHO_API int echo_optimizer_set_tensors(echo_optimizer_t* h, const echo_opt_tensor_t* tensors, int32_t n_tensors, void* stream);
/* `step` = the number of this update, starting at 1 (torch.optim's state['step']): it sets the bias corrections.  Hyper-parameters
 * are doubles: torch evaluates 1 - beta2, lr / (1 - beta1^step), ... on python floats and rounds the results to fp32. */
ECHO_API int echo_optimizer_step(echo_optimizer_t* h, int64_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                        double clip_max_norm /* <= 0: no clip */, void* stream);
/* out4 = {steps taken, parameters, chunks, NaN gradients scrubbed so far}; clip2 (may be NULL) = {last total norm, last
 * clip coefficient}.  Synchronises the stream (diagnostics; the step itself never does). */
ECHO_API int echo_optimizer_info(const echo_optimizer_t* h, int64_t* out4, float* clip2, void* stream);
ECHO_API void echo_optimizer_destroy(echo_optimizer_t* h);

/* ---- constraint metrics: validate_constrains / validate_constrains_changes, helpers/metrics_3dfront.py:57-306 -- the
 * Python loop over the triples at the end of scripts/eval_3dfront.py (:209-210, :305) as one kernel.
 *   triples (T,3) int64 [s,p,o]; boxes (N, D) f32, D = 6 or 7: [l, h, w, px, py, pz(, angle)] (denormalised, as the reference passes
 *   them); keep (N) int32 or NULL; rel_of_pred_host[n_preds]: relation code of every predicate id (ECHO_REL_*, -1 = not a checked
 *   relation) -- the caller resolves vocab["pred_idx_to_name"][p][:-1] once; changes_mode 0: a triple counts when keep is NULL or both
 *   nodes are kept (validate_constrains), 1: when keep is NULL or either node changed (validate_constrains_changes).
 *   out_rel (T) int8: ECHO_REL_* of the triple or -1 when it was skipped; out_ok (T) int8: 1 = the constraint holds. */
#define ECHO_REL_LEFT 0
#define ECHO_REL_RIGHT 1
#define ECHO_REL_FRONT 2
#define ECHO_REL_BEHIND 3
#define ECHO_REL_BIGGER 4
#define ECHO_REL_SMALLER 5
#define ECHO_REL_TALLER 6
#define ECHO_REL_SHORTER 7
#define ECHO_REL_STANDING_ON 8
#define ECHO_REL_CLOSE_BY 9
#define ECHO_REL_SYMMETRICAL_TO 10
ECHO_API int echo_metrics_validate_constraints(const int64_t* triples, int64_t n_triples, const float* boxes, int64_t n_nodes, int32_t box_dim,
                                      const int32_t* keep, int32_t changes_mode, const int32_t* rel_of_pred_host, int32_t n_preds,
                                      int32_t strict, float overlap_threshold, int8_t* out_rel, int8_t* out_ok, void* stream);

/* ---- VQ-VAE decode: VQVAE.decode_no_quant, model/networks/vqvae_networks/network.py:95-103 -- what
 * EchoToShape.rel2shape calls on the latents the DDIM chain returns (echo2shape.py:522).
 *   quantize (quantizer.py:68-99: nearest codebook entry per voxel) -> post_quant_conv -> Decoder3D
 *   (vqvae_modules.py:292-409, config/vqvae_snet.yaml).  Weights by the reference's state_dict names:
 *   quantize.embedding.weight, post_quant_conv.*, decoder.*.
 * latents (n, 3, 16, 16, 16) f32 NCDHW -> sdf_out (n, 1, 64, 64, 64) f32; indices_out (n * 4096) int32 or NULL. */
typedef struct echo_vqvae echo_vqvae_t;
typedef struct {
  int32_t embed_dim, n_embed, z_channels, latent_size;   /* 3, 8192, 3, 16 */
  int32_t ch, num_levels, ch_mult[8], num_res_blocks, out_ch;   /* 64, 3, {1,2,4}, 1, 1 */
  int32_t max_objects;
  int32_t precision;                                     /* ECHO_PREC_* */
} echo_vqvae_desc_t;
ECHO_API int echo_vqvae_create(echo_vqvae_t** out, const echo_vqvae_desc_t* desc, const echo_weight_t* weights, int32_t n_weights);
ECHO_API int echo_vqvae_decode(echo_vqvae_t* h, const float* latents, int32_t n, float* sdf_out, int32_t* indices_out, void* stream);
/* ---- VQ-VAE encode: VQVAE.encode_no_quant, model/networks/vqvae_networks/network.py:84-88 -- what the
 * training step calls on the ground-truth SDFs (echo2shape.py:334-364): Encoder3D (vqvae_modules.py:198-289: conv_in,
 * per level ResnetBlock(s) + Downsample [zero pad (0,1) + Conv3d k3 stride 2], mid ResnetBlock / AttnBlock / ResnetBlock,
 * GroupNorm, GELU, conv_out) -> quant_conv 1x1x1, no quantisation.  Same desc as the decoder (in_channels == out_ch == 1,
 * double_z False); weights by the reference's names encoder.*, quant_conv.*.  ECHO_PREC_FP32 only (ECHO_ERR_UNSUPPORTED
 * otherwise).  A handle is either a decoder or an encoder; both are freed by echo_vqvae_destroy.
 * sdf (n, 1, R, R, R) f32, R = latent_size * 2^(num_levels-1) -> latents_out (n, embed_dim, L, L, L) f32 NCDHW. */
ECHO_API int echo_vqvae_encoder_create(echo_vqvae_t** out, const echo_vqvae_desc_t* desc, const echo_weight_t* weights, int32_t n_weights);
ECHO_API int echo_vqvae_encode(echo_vqvae_t* h, const float* sdf, int32_t n, float* latents_out, void* stream);
ECHO_API void echo_vqvae_destroy(echo_vqvae_t* h);

/* ---- once-per-scene encoders of Sg2ScDiffModel.sample (model/EchoScene.py:143-157, 181-195, 388-410): the stage between
 * the dataset's tensors and the two chains.  state_dict keys: obj_embeddings_ec.weight (num_objs, 2*gconv_dim),
 * pred_embeddings_ec.weight (num_preds, 2*gconv_dim), gconv_net_ec.*, gconv_net_manipulation.*, rel_s_mlp.{0,1,3}.*
 * (rel_s_mlp optional: the layout-only model has none).  feat = 2*gconv_dim + add_dim (640). */
typedef struct echo_scene_desc {
  int32_t gconv_dim;            /* embedding_dim = 64 (SGDiff.py:21) */
  int32_t add_dim;              /* 512 with CLIP features (use_clip), else 0 */
  int32_t num_objs;             /* rows of obj_embeddings_ec = len(vocab['object_idx_to_name']) + 1 (EchoScene.py:47) */
  int32_t num_preds;            /* rows of pred_embeddings_ec (EchoScene.py:48) */
  int32_t num_layers;           /* gconv_num_layers = 5 */
  int32_t rel_s_hidden;         /* 960 (EchoScene.py:97-100) */
  int32_t context_dim;          /* 1280 */
  int32_t max_nodes, max_triples;
  float bn_eps;                 /* 1e-5 */
  int32_t manipulate_pred_dc;   /* 0: manipulate embeds predicates with pred_embeddings_ec (Sg2ScDiffModel, EchoScene.py:187);
                                 * 1: with pred_embeddings_man_dc.weight (the layout-only Sg2BoxDiffModel, EchoLayout.py:154) */
  int32_t keep_train_weights;   /* != 0: the GCN MLPs (and rel_s_mlp) also keep their unfolded tensors, for the batch-statistics forward
                                  (echo_*_set_batch_stats) */
} echo_scene_desc_t;
ECHO_API int echo_scene_create(echo_scene_t** out, const echo_scene_desc_t* desc, const echo_weight_t* weights, int32_t n_weights);
/* init_encoder(objs, triples, text_feat, rel_feat) (EchoScene.py:143-157): objs (N) i64 class ids -- the caller guarantees
 * 0 <= objs[i] < num_objs (nn.Embedding would raise; predicate ids are checked against the graph) --, text_feat (N, add_dim),
 * rel_feat (T, add_dim) -> obj_embed (N, feat), pred_embed (T, feat), latent_obj (N, feat).  Outputs may be NULL. */
ECHO_API int echo_scene_init_encoder(echo_scene_t* h, const echo_graph_t* g, const int64_t* objs, const float* text_feat,
                                     const float* rel_feat, float* obj_embed_out, float* pred_embed_out, float* latent_obj_out,
                                     void* stream);
/* manipulate(latent_f, objs, triples, text_feat, rel_feat) (EchoScene.py:181-195): latent_f (N, feat + gconv_dim) =
 * [latent | change flag] -> latent (N, feat), obj_embed (N, feat), pred_embed (T, feat).  Outputs may be NULL. */
ECHO_API int echo_scene_manipulate(echo_scene_t* h, const echo_graph_t* g, const float* latent_f, const int64_t* objs,
                                   const float* text_feat, const float* rel_feat, float* latent_out, float* obj_embed_out,
                                   float* pred_embed_out, void* stream);
/* rel_s_mlp(x) (EchoScene.py:97-100, 405-410): x (rows, feat) -> out (rows, context_dim); rows <= max_nodes */
ECHO_API int echo_scene_rel_s(echo_scene_t* h, const float* x, int32_t rows, float* out, void* stream);
/* The whole encoder stage of Sg2ScDiffModel.sample (EchoScene.py:388-410) as one asynchronous call: init_encoder ->
 * [latent_obj | change | obj_embed] -> manipulate -> rel_s_mlp twice.  change (N, gconv_dim) or NULL = zeros (sample, :393-397).
 * -> obj_embed (N, feat) [layout branch obj_embed], latent (N, feat) [layout branch relation condition], uc_s / c_s
 * (N, context_dim) [shape branch conditionings; either may be NULL when only the layout is sampled]. */
ECHO_API int echo_scene_encode(echo_scene_t* h, const echo_graph_t* g, const int64_t* objs, const float* text_feat,
                               const float* rel_feat, const float* change, float* obj_embed_out, float* latent_out,
                               float* uc_s_out, float* c_s_out, void* stream);
ECHO_API void echo_scene_destroy(echo_scene_t* h);

/* ---- schedule tables, for host-side checks against the reference's buffers.
 * layout: 5 x time_num f32 [sqrt_recip_ac, sqrt_recipm1_ac, post_mean_coef1, post_mean_coef2, post_log_var_clipped]
 * shape : ddim_steps x 4 f32 [sqrt(a_t), sqrt(1-a_t), sqrt(a_prev), sqrt(1-a_prev)] and the int32 DDIM timesteps. */
ECHO_API int echo_layout_schedule(const echo_layout_t* h, float* host_out);
ECHO_API int echo_shape_schedule(const echo_shape_t* h, float* host_coef_out, int32_t* host_timesteps_out);

/* ---- single operators (each is one kernel family of the hot path), exported for per-kernel parity tests and
 * profiling.  Layouts: activations channels-last (n, d, h, w, c) f32; conv weights in the reference layout
 * (cout, cin, kd, kh, kw) f32.  precision selects the contraction arithmetic (ECHO_PREC_*). */
ECHO_API int echo_op_conv3d(const float* x, int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin,
                   const float* weight, const float* bias, int32_t cout, int32_t ksize,
                   int32_t stride_d, int32_t stride_hw, float* out, int32_t precision, void* stream);
/* Upsample (nearest x(1,2,2), openai_model_3d.py:150-155) followed by Conv3d k3 p1 (:156-157), evaluated as four
 * output-phase convolutions with folded 3x2x2 taps on the low-resolution input: x (n,d,h,w,cin) -> out (n,d,2h,2w,cout).
 * ECHO_PREC_BF16 only. */
ECHO_API int echo_op_upconv3d(const float* x, int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin,
                     const float* weight, const float* bias, int32_t cout, float* out, int32_t precision, void* stream);
/* the same with the upsample doubling the depth as well (F.interpolate(scale_factor=2), vqvae_modules.py:35-39): eight output
 * phases x 2x2x2 folded taps; x (n,d,h,w,cin) -> out (n,2d,2h,2w,cout) */
ECHO_API int echo_op_upconv3d_x2(const float* x, int32_t n, int32_t d, int32_t h, int32_t w, int32_t cin,
                        const float* weight, const float* bias, int32_t cout, float* out, int32_t precision, void* stream);
ECHO_API int echo_op_linear(const float* x, int64_t rows, int32_t cin, const float* weight, const float* bias, int32_t cout,
                   float* out, int32_t precision, void* stream);
ECHO_API int echo_op_group_norm(const float* x, int32_t n, int64_t voxels, int32_t c, int32_t groups, const float* gamma,
                       const float* beta, float eps, int32_t silu, float* out, void* stream);
ECHO_API int echo_op_layer_norm(const float* x, int64_t rows, int32_t c, const float* gamma, const float* beta, float eps,
                       float* out, void* stream);
ECHO_API int echo_op_attention(const float* qkv, int32_t n, int32_t tokens, int32_t heads, int32_t dim_head, float* out,
                      int32_t precision, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ECHOSCENE_B200_H */
