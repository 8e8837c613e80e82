// Kernel launchers of the EchoScene denoiser hot path (declarations).  Every launcher is asynchronous on the
// given stream, allocates nothing, and counts its launches (common.cuh).
#pragma once
#include "common.cuh"

namespace echo {

// ---------------------------------------------------------------------------------------------------------------
// Implicit-GEMM contraction:  out[m, n] = act( sum_{tap,c} A[vox(m,tap), c] * W[n, tap*cin + c]
//                                               + bias[n] + rowvec[obj(m), n] + res[m, n] )
// A is a channels-last activation; m enumerates OUTPUT voxels (obj, od, oh, ow).  Linear layers are the 1x1x1 case.
// Strides are in ELEMENTS.  `batch` lets one launch run nb0*nb1 independent problems (attention heads).
// ---------------------------------------------------------------------------------------------------------------
struct GemmArgs {
  const void* A = nullptr;
  DT a_dt = F32;
  int n = 1, d = 1, h = 1, w = 1;   // input grid per object
  int cin = 0;                      // reduction channels per tap
  int64_t lda = 0;                  // elements between consecutive input voxels (>= cin)
  int od = 1, oh = 1, ow = 1;       // output grid per object
  int kd = 1, kh = 1, kw = 1, sd = 1, sh = 1, sw = 1, pd = 0, ph = 0, pw = 0;
  const void* W = nullptr;
  DT w_dt = F32;
  int64_t w_stride_n = 0, w_stride_k = 1;   // W[n, k] at n*w_stride_n + k*w_stride_k
  int cout = 0;
  const float* bias = nullptr;
  const float* rowvec = nullptr;    // [n_obj, ld_rowvec]
  int64_t ld_rowvec = 0;
  const void* res = nullptr;        // [rows_out, ld_res]
  DT res_dt = F32;
  int64_t ld_res = 0;
  void* out = nullptr;
  DT out_dt = F32;
  int64_t ldo = 0;
  int act = 0;                      // 0 none, 1 relu
  int epi = 0;                      // tcgen05 path only: 1 = GEGLU epilogue (W rows tiled [128 a | 128 g]; out width cout/2)
  int out_t = 0;                    // tcgen05 path only: bf16 output stored transposed per object, out[(obj*cout + n)*voxels + voxel]
  int up2 = 0;                      // tcgen05 path only: the conv follows a nearest upsample of A; A is the LOW-resolution tensor
                                    // (d,h,w), W = fold_upsample_weight(): 1 = x(1,2,2) (openai_model_3d.py:150-157), output grid
                                    // (d,2h,2w); 2 = x(2,2,2) (vqvae_modules.py:35-39), output grid (2d,2h,2w)
  float alpha = 1.f;                // scales the accumulator before the epilogue adds
  // batching: problem b = b0*nb1 + b1
  int nb0 = 1, nb1 = 1;
  int64_t a_bs0 = 0, a_bs1 = 0, w_bs0 = 0, w_bs1 = 0, o_bs0 = 0, o_bs1 = 0;
  float* colsum = nullptr;          // tcgen05 path: per-column (sum, sumsq) partials, gemm_tc_colsum_rows() x cout x 2 floats
  void* scratch = nullptr;          // tcgen05 path, stride-2 convs: room for a space-to-depth copy of A (same bytes)
  void* splitk_ws = nullptr;        // tcgen05 path: gemm_tc_splitk_ws_bytes() of scratch lets the plan split K over tap groups
  // tcgen05 path, split precision (ECHO_PREC_X3): A / W are the HIGH bf16 halves, these the LOW halves (x = hi + lo, split_bf16());
  // three MMAs per k-step into one fp32 accumulator.  Stride-2 convs need `scratch` for a space-to-depth copy of BOTH halves.
  const void* A_lo = nullptr;
  const void* W_lo = nullptr;
  int64_t rows_out() const { return (int64_t)n * od * oh * ow; }
  int ktot() const { return kd * kh * kw * cin; }
};

void gemm_simt(const GemmArgs& g, cudaStream_t s);
// tcgen05 + TMA path (gemm_tc.cu).  Requirements are checked by gemm_tc_supported().
bool gemm_tc_supported(const GemmArgs& g);
void gemm_tc(const GemmArgs& g, cudaStream_t s);
bool tc_available();
// timing probe: CUDA events around every gemm_tc launch of one shape while steps run (bench.py roofline)
void tc_probe_begin(long long rows, int cin, int cout, int k);
int tc_probe_end(double* avg_ms);
void tc_probe_timeline(unsigned long long* buf);
// rows of the column-sum partial buffer per OBJECT for this problem (the tcgen05 kernel writes one per 128-voxel tile)
int gemm_tc_colsum_rows_per_obj(const GemmArgs& g);
// floats of one partial row for a tensor with c channels: [c/32 chunks][8 slots][2] -- (sum, sumsq) of the 7-channel blocks
// every 32-column chunk touches (slot s of chunk q <-> block 32q/7 + s); 0 when c is not a multiple of 224 (32 groups x 7)
size_t gemm_tc_colsum_row_floats(int c);
// the launch plan gemm_tc() would pick (host-only): out4 = {block_n, sub-blocks per CTA, split-K, CTA pairs}
void tc_plan_describe(int n, int d, int h, int w, int cin, int cout, int ksize, int epi, int up2, int allow_splitk, int sms, int* out4);
// bytes of fp32 workspace that let gemm_tc() split the reduction of this problem (0: never split)
size_t gemm_tc_splitk_ws_bytes(const GemmArgs& g);
// GroupNorm(+SiLU) with the statistics folded from the producer's column partials inside the apply kernel; `xb` != null:
// the input is the channel concat [xa | xb] (read in place), `cat` != null additionally receives the raw concat.
bool gn_apply_cs_supported(const Act& xa, const Act* xb, const Act& out);
void gn_apply_cs(const Act& xa, const Act* xb, const float* gamma, const float* beta, int groups, float eps, bool silu, const Act& out,
                 const Act* cat, cudaStream_t s);
// precision: ECHO_PREC_*; BF16 falls back to the SIMT kernel (bf16 operands, fp32 accumulate) for shapes the
// tensor-core kernel does not take (tiny channel counts).
void gemm(const GemmArgs& g, int precision, cudaStream_t s);

// ---------------------------------------------------------------------------------------------------------------
// Few-row ("skinny") linear: Y[M, nout] = epi( f(X)[M, K] @ W[nout, K]^T ), M = nodes/edges.  HBM-bound on W.
// ---------------------------------------------------------------------------------------------------------------
enum LinPrologue : int { PRO_NONE = 0, PRO_SILU = 1, PRO_GN = 2, PRO_LN = 3, PRO_GEGLU = 4 };
struct LinArgs {
  const float* X = nullptr;
  int64_t ldx = 0;
  int M = 0, K = 0, nout = 0;
  // optional channel concat: columns [0, K1) come from X, [K1, K) from X2
  const float* X2 = nullptr;
  int64_t ldx2 = 0;
  int K1 = 0;
  // prologue applied to X while it is loaded (see linear.cu)
  int pro = PRO_NONE;
  int pro_act = 0;             // PRO_GN: SiLU after the affine normalisation
  const float* gamma = nullptr;
  const float* beta = nullptr;
  float eps = 1e-5f;
  int cpg = 0;                 // PRO_GN: channels per group
  const float* res2 = nullptr; // second residual [M, ld_res2]
  int64_t ld_res2 = 0;
  const void* W = nullptr;     // [nout, K] row-major, row stride ldw (0 = K)
  DT w_dt = F32;
  int64_t ldw = 0;
  const float* bias = nullptr;
  const float* res = nullptr;  // [M, ld_res] added after activation
  int64_t ld_res = 0;
  float* Y = nullptr;
  int64_t ldy = 0;
  int in_act = 0;              // 0 none, 1 SiLU applied to X on load
  int act = 0;                 // 0 none, 1 relu, 2 silu (before res)
  // linear_auto, more than 64 rows (collated batches): room for the prologue's output [M, K] (after which the contraction runs as a
  // tensor-core GEMM; without it a layer with a prologue stays on the few-row kernel at any M) and, behind it, for the partial tiles
  // of a split reduction
  float* scratch = nullptr;
  size_t scratch_floats = 0;
};
void linear_rows(const LinArgs& a, cudaStream_t s);
// out[m, :] = the input row of the Linear after its prologue (SiLU / GroupNorm(+SiLU) / LayerNorm / GEGLU, concat of X and X2), the
// arithmetic of linear_rows_kernel; out is dense [M, K]
void linear_prologue(const LinArgs& a, float* out, cudaStream_t s);

// fp32-grade GEMM on the tensor cores (3 x TF32 split, sgemm_x3.cu): C[m, n] = sum_k A(m, k) B(k, n) (+ bias[n]) (+ res[m, n]),
// A(m, k) = A[m * sam + k * sak], B(k, n) = B[k * sbk + n * sbn]; one stride of each operand must be 1.
struct SgemmX3Args {
  const float* A = nullptr;
  int64_t sam = 0, sak = 1;
  const float* B = nullptr;
  int64_t sbk = 0, sbn = 1;
  float* C = nullptr;
  int64_t ldc = 0;
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;
  const float* res = nullptr;   // may alias C (accumulate in place)
  int64_t ld_res = 0;
  int act = 0;                  // 0 none, 1 relu, 2 silu: applied after the bias, before res / res2 (LinArgs order)
  const float* res2 = nullptr;
  int64_t ld_res2 = 0;
  // split reduction: slice z of `splits` covers k in [z * chunk, min(K, (z + 1) * chunk)) and writes C + z * c_bs (no bias / res then)
  int splits = 1, chunk = 0;
  int64_t c_bs = 0;
};
bool sgemm_x3_supported(const SgemmX3Args& g);
void sgemm_x3(const SgemmX3Args& g, cudaStream_t s);
void sgemm_x3_auto(const SgemmX3Args& g, float* ws, size_t ws_floats, cudaStream_t s);   // splits the reduction when the grid is small
bool linear_rows_gn_supported(int K, int cpg);

// ---------------------------------------------------------------------------------------------------------------
// normalisations / elementwise
// ---------------------------------------------------------------------------------------------------------------
// GroupNorm over (voxels x C/groups) per (object, group); fp32 statistics (GroupNorm32, ldm_diffusion_util.py:237).
// `stats` gets (mean, rstd) per (object, group); `partial` is scratch of gn_partial_floats(x) floats.
size_t gn_partial_floats(const Act& x, int groups);
void gn_stats(const Act& x, int groups, float eps, float* stats, float* partial, cudaStream_t s);
void gn_apply(const Act& x, const float* stats, const float* gamma, const float* beta, int groups, bool silu,
              const Act& out, cudaStream_t s);
// rows variant for the layout branch (length-1 signals): X [M, C] -> Y [M, C]
void gn_rows(const float* x, int M, int C, int groups, const float* gamma, const float* beta, float eps, bool silu,
             float* y, cudaStream_t s);
void layer_norm(const void* x, DT xdt, int64_t rows, int C, const float* gamma, const float* beta, float eps,
                void* y, DT ydt, cudaStream_t s);
// GEGLU: x [rows, 2F] = [a | g] -> y [rows, F] = a * gelu_erf(g)   (attention.py:39-46)
void geglu(const void* x, DT xdt, int64_t rows, int F, void* y, DT ydt, cudaStream_t s);
void concat_channels(const Act& a, const Act& b, const Act& out, cudaStream_t s);
void upsample_hw2(const Act& x, const Act& out, cudaStream_t s);              // nearest x(1,2,2)
void maxpool3d(const Act& x, int k, int stride, const Act& out, cudaStream_t s);
void ncdhw_to_cl(const float* x, int n, int c, int64_t voxels, void* out, DT odt, cudaStream_t s);
void ncdhw_to_cl_pad16(const float* x, int n, int c, int64_t voxels, __nv_bfloat16* out, cudaStream_t s);   // channels zero-padded to 16
// x channels-last with row stride ld (>= c) -> out (n, c, voxels)
void cl_to_ncdhw(const void* x, DT xdt, int n, int c, int64_t voxels, int ld, float* out, cudaStream_t s);
void convert(const void* x, DT xdt, void* y, DT ydt, int64_t count, cudaStream_t s);
// x (fp32) -> hi = bf16(x), lo = bf16(x - hi): the operand format of the split-precision tcgen05 contraction; count % 4 == 0
void split_bf16(const float* x, int64_t count, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s);
// y = x * sigmoid(x) (the SiLU in front of every ResBlock's emb_layers Linear, applied once to the shared time embedding)
void silu_f32(const float* x, float* y, int64_t count, cudaStream_t s);
// y[r, :] += v[r / rows_per_obj, :]
void add_rowvec(void* y, DT ydt, int64_t rows, int C, const float* v, int64_t ldv, int64_t rows_per_obj, cudaStream_t s);
// in-place row softmax of S [rows, cols] (fp32)
void softmax_rows(float* S, int64_t rows, int cols, cudaStream_t s);
// [cos(t f) | sin(t f)] (ldm_diffusion_util.py:174-194); t int64 device, freqs [dim/2] device
void timestep_embedding_tab(const int64_t* t, const float* freqs, int n, int dim, float* out, cudaStream_t s);
void fill_i64(int64_t* p, int n, int64_t v, cudaStream_t s);
// out[r, col0 : col0+D] = table[idx[r*idx_stride + idx_off], :]
void embedding_rows(const float* table, int D, const int64_t* idx, int64_t idx_stride, int64_t idx_off, int64_t rows,
                    float* out, int64_t ldo, cudaStream_t s);
// out[r, col0:col0+D] = src[r, :D]
void copy_cols(const float* src, int64_t lds, int64_t rows, int D, float* out, int64_t ldo, cudaStream_t s);
// flatten(1) of a channels-last (n, d,h,w, c) tensor in NCDHW order -> [n, c*voxels] f32
void flatten_ncdhw(const Act& x, float* out, cudaStream_t s);

// samplers
// DDPM: x_prev = c1*(a*x - b*eps) + c2*x + [t>0] exp(0.5 lv) noise   (diffusion_ddpm.py:220-309); tab = 5 x T
void ddpm_update(const float* x, const float* eps, const float* noise, const float* tab, int T, int t, int64_t count,
                 float* out, cudaStream_t s);
// DDIM eta=0 on an NCDHW latent, e_t given channels-last (or NCDHW when e_cl == false)
void ddim_update(const float* x_ncdhw, const void* e, DT edt, bool e_cl, int n, int c, int64_t voxels, int e_ld,
                 const float* coef4 /* device, 4 floats */, float* out_ncdhw, cudaStream_t s, const int* slot = nullptr);
// p[i] = table[*slot] (i < n) / *p = v: the DDIM index of a graph-replayed chain lives on the device
void fill_i64_from_slot(int64_t* p, int n, const int32_t* table, const int* slot, cudaStream_t s);
void set_i32(int* p, int v, cudaStream_t s);

// attention (fp32, materialised scores): qkv [n*tokens, 3*heads*dh] -> out [n*tokens, heads*dh]
void attention_f32(const float* qkv, int n, int tokens, int heads, int dh, float* scores_ws, float* out, cudaStream_t s);
size_t attention_f32_ws_floats(int n, int tokens, int heads);
// attention (bf16 flash kernel): qkv bf16 [rows, 3*heads*attention_pad_dh(dh)] (heads zero-padded), out bf16 [rows, heads*dh]
// attention (tcgen05 flash kernel, flash_tc.cu): qk bf16 [rows, 2*heads*64] (q then k, heads zero-padded to 64), vt bf16
// [(n*heads*64), tokens] (V transposed per (object, head)), out bf16 [rows, heads*dh]
bool attention_tc_supported(int tokens, int dh);
void attention_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vt, int n, int tokens, int heads, int dh, __nv_bfloat16* out, cudaStream_t s);
void split_qkv_tc(const float* qkv, int n, int tokens, int heads, int dh, __nv_bfloat16* qk, __nv_bfloat16* vt, cudaStream_t s);
int attention_tc_error();   // 1 if a block ever found its shared-memory window misaligned (results invalid)
int attention_pad_dh(int dh);
bool attention_bf16_supported(int tokens, int dh);
void attention_bf16(const __nv_bfloat16* qkv, int n, int tokens, int heads, int dh, __nv_bfloat16* out, cudaStream_t s);
void pad_qkv(const float* x, int64_t rows, int heads, int dh, int dhp, __nv_bfloat16* y, cudaStream_t s);

// weight preparation
// 3x3x3 conv after nearest x(1,2,2) upsample == 4 output-phase convs with 3x2x2 taps on the low-res input (the two
// high-res rows that map to one low-res row share it, so their weights add): w [cout][27][cin] fp32 (tap-major) ->
// out [cout][4 phases (py,px)][12 taps (kd,a,b)][cin] fp32, host-side (runs once at handle creation)
// up_depth: the upsample also doubles the depth (F.interpolate(scale_factor=2), vqvae_modules.py:36): 8 phases (pz,py,px) x
// 8 taps -> out [cout][8][8][cin]
void fold_upsample_weight(const float* w_host, int cout, int cin, float* out_host, bool up_depth = false);
// conv weight (cout, cin, taps) -> (cout, taps, cin); taps = kd*kh*kw
void repack_conv_weight(const float* w, int cout, int cin, int taps, float* out, cudaStream_t s);
// centre tap of a Conv1d(k=3) weight (cout, cin, 3) -> (cout, cin)
void conv1d_center_tap(const float* w, int cout, int cin, int k, float* out, cudaStream_t s);
// fold eval BatchNorm1d into the preceding Linear: W' = W*g/sqrt(v+eps), b' = (b-m)*g/sqrt(v+eps)+beta
void fold_bn(const float* w, const float* b, const float* gamma, const float* beta, const float* mean, const float* var,
             float eps, int nout, int K, float* w_out, float* b_out, cudaStream_t s);

// graph ops
void gather_rows(const float* src, const int64_t* idx, int64_t n_idx, int64_t n_rows, int64_t D, float* out, cudaStream_t s);   // idx outside [0, n_rows): NaN row

// ---- training side (train.cu) ----
// out = sqrt_ac[t_r] * x0 + sqrt_1mac[t_r] * noise, rows x row_len, one timestep per row
void q_sample(const float* x0, const float* noise, const int64_t* t, const float* sqrt_ac, const float* sqrt_1mac, int64_t rows, int64_t row_len,
              float* out, cudaStream_t s);
// out[r][k] = mean((target - pred)^2 over columns ranges[2k] .. ranges[2k + 1]) -- ranges on the device
void mse_rows(const float* pred, const float* target, int64_t rows, int64_t row_len, const int* ranges_dev, int n_ranges, float* out, cudaStream_t s);

// ---- constraint metrics (metrics.cu): one thread per triple; out_rel = relation code (-1: not evaluated), out_ok = 0 / 1
void validate_constraints(const int64_t* triples, int64_t T, const float* boxes, int64_t N, int D, const int32_t* keep, bool changes_mode,
                          const int32_t* rel_of_pred_dev, int n_preds, bool strict, float overlap_threshold, int8_t* out_rel, int8_t* out_ok,
                          cudaStream_t s);
}  // namespace echo
