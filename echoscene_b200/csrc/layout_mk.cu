// Persistent executor of one layout DDPM iteration: UNet1DModel.forward (denoise_net.py:773-806, incl. box_messsage_passing
// :758-771 and the GraphTripleConvNet of graph.py:124-250) + the posterior update (diffusion_ddpm.py:220-309) as ONE kernel.
//
// Why: the layout step is a chain of ~130 dependent few-row layers (rows = nodes / triples, 8..64 of them) whose cost as
// separate launches is the chain of launch + first-touch latencies (1.56 ms for 461 MB of weights = 4.5 % of the HBM
// roofline, r1).  Here the step is a PROGRAM (layout.cu records it once per node/triple count): stages of independent ops,
// each op cut into units (16 rows x FU output features); one cooperative grid of one CTA per SM walks the stages.
//
//   * weights never wait for a barrier: they are constants, so a producer warp per CTA streams the CTA's future weight
//     slices HBM -> shared memory with cp.async.bulk (1-D TMA) through a 3-slot ring (120 KB in flight per SM), running
//     ahead of the compute by up to three units = usually three stages;
//   * a stage boundary is one arrival counter in L2 (release: bar.sync + fence + atomicAdd; acquire: one polling thread +
//     bar.sync), ~1 us instead of a kernel boundary; activations are read with ld.global.cg (they are rewritten every step by
//     other SMs, L1 must not serve them);
//   * the elementwise op in front of a Linear is a prologue applied while its input rows are staged in shared memory (SiLU,
//     GroupNorm via lane shuffles, LayerNorm per warp-row, GEGLU, the GraphTripleConv edge gather-combine with the node
//     features staged per edge row, and the CSR mean pooling in the reference's summation order), so a stage is exactly one
//     dependent contraction;
//   * the time-embedding path (time MLP, the 22 stacked emb_layers projections = 40 % of the weight bytes) is computed for
//     ONE row (all nodes of a step share t) and runs as background ops in the barrier shadow of the GCN stages.
//
// Contraction mapping inside a unit (8 consumer warps): the staged rows X_s [R][K] and the weight slice W_s [FU][K] are
// both K-contiguous in shared memory; warp w owns K-chunks of 64 (one float2 per lane), keeps R x 4 accumulators per
// feature group in registers, and a halving butterfly + a fixed-order sum over the warps finishes the dot products
// (deterministic: no atomics on data).
#include "layout_mk.cuh"

#include "tc_ptx.cuh"

namespace echo {
namespace {

using namespace ptx;

constexpr int MK_THREADS = 288;   // 8 consumer warps + 1 producer warp
constexpr int MK_SLOTS = 3;
constexpr int MK_MAXG = MK_MAX_FU / 4;
constexpr int SM_W = MK_XCAP * 4;
constexpr int SM_RED = SM_W + MK_SLOTS * MK_SLOT_BYTES;
constexpr int SM_OPS = SM_RED + MK_MAXG * 8 * 64 * 4;
constexpr int SM_BAR = SM_OPS + MK_MAX_STAGE_OPS * 256;
constexpr int SM_TOTAL = SM_BAR + 64;

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float4 ld4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Watchdog: a counter that does not fill within ~2^27 polls (seconds; a stage takes microseconds) means a broken program or a
// lost CTA -- flag it and run on, so that a bug shows up as a wrong result + error word instead of a hung device.
__device__ __forceinline__ void wait_counter(const unsigned* p, unsigned target, unsigned* err) {
  unsigned spins = 0;
  while ((int)(ld_acquire(p) - target) < 0) {
    if (++spins > (1u << 27)) { atomicExch(err, 1u); break; }
  }
}
__device__ __forceinline__ void arrive_counter(unsigned* p) {
  __threadfence();
  atomicAdd(p, 1u);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Halving butterfly over the lanes: on return lane l holds the totals of elements [base, base + N/32) in v[0 .. N/32).
template <int N>
__device__ __forceinline__ int butterfly(float (&v)[N], int lane) {
  int base = 0;
#pragma unroll
  for (int off = 16, n = N; off >= 1; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    if (up) base += n / 2;
  }
  return base;
}

__device__ __forceinline__ const float* resolve_x(const MkOp& op, const MkArgs& a) {
  return op.x_ext == MK_EXT_XT ? a.x_t : op.x_ext == MK_EXT_OBJ ? a.obj_embed : op.X;
}

// rows m0 .. m0+R of the op's input, prologue applied, columns [seg0, seg0 + seg_len) -> Xs [R][KS]; one warp per row
template <int R>
__device__ void stage_input(const MkOp& op, const MkArgs& a, const float* X, float* Xs, int KS, int m0, int seg0, int seg_len, int warp,
                            int lane) {
  const int nq = seg_len >> 2;
  for (int i = warp; i < R; i += 8) {
    const int m = m0 + i;
    float4* xr = reinterpret_cast<float4*>(Xs + i * KS);
    if (m >= op.M) {
      for (int q = lane; q < nq; q += 32) xr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    switch (op.pro) {
      case MK_GN: {   // K % 128 == 0: every lane is in range in every iteration, a group is cpg/4 adjacent lanes
        const int gl = op.cpg >> 2;
        const float inv = 1.f / (float)op.cpg;
        for (int q = lane; q < nq; q += 32) {
          const int k = seg0 + 4 * q;
          float4 v = (op.X2 && k >= op.K1) ? ld4(op.X2 + (long long)m * op.ldx2 + (k - op.K1)) : ld4(X + (long long)m * op.ldx + k);
          const float4 gm = __ldg(reinterpret_cast<const float4*>(op.gamma + k)), bt = __ldg(reinterpret_cast<const float4*>(op.beta + k));
          float s = (v.x + v.y) + (v.z + v.w);
          for (int o = 1; o < gl; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = s * inv;
          const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
          float ss = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          for (int o = 1; o < gl; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
          const float rstd = rsqrtf(ss * inv + op.eps);
          v.x = d0 * rstd * gm.x + bt.x; v.y = d1 * rstd * gm.y + bt.y; v.z = d2 * rstd * gm.z + bt.z; v.w = d3 * rstd * gm.w + bt.w;
          if (op.pro_act) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
          xr[q] = v;
        }
        break;
      }
      case MK_LN: {   // whole row by this warp (K <= 1280): two-pass statistics in registers
        float4 v[10];
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const int q = lane + 32 * j;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < nq) { v[j] = ld4(X + (long long)m * op.ldx + 4 * q); s += (v[j].x + v[j].y) + (v[j].z + v[j].w); }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / (float)op.K;
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          if (lane + 32 * j < nq) {
            const float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
            ss += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float rstd = rsqrtf(ss / (float)op.K + op.eps);
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const int q = lane + 32 * j;
          if (q < nq) {
            const float4 gm = __ldg(reinterpret_cast<const float4*>(op.gamma + 4 * q)), bt = __ldg(reinterpret_cast<const float4*>(op.beta + 4 * q));
            float4 o4;
            o4.x = (v[j].x - mean) * rstd * gm.x + bt.x; o4.y = (v[j].y - mean) * rstd * gm.y + bt.y;
            o4.z = (v[j].z - mean) * rstd * gm.z + bt.z; o4.w = (v[j].w - mean) * rstd * gm.w + bt.w;
            xr[q] = o4;
          }
        }
        break;
      }
      case MK_EDGE: {   // X = [Ps | Po] (N, 2H), aux0 = Pp (T, H), aux1 = folded bias; same association as edge_combine_kernel
        const int H = op.K;
        const float* ps = X + (long long)a.s_idx[m] * 2 * H;
        const float* po = X + (long long)a.o_idx[m] * 2 * H + H;
        const float* pq = op.aux0 + (long long)m * H;
        for (int q = lane; q < nq; q += 32) {
          const int k = seg0 + 4 * q;
          const float4 A = ld4(ps + k), B = ld4(pq + k), C = ld4(po + k), D = __ldg(reinterpret_cast<const float4*>(op.aux1 + k));
          float4 r;
          r.x = fmaxf(((A.x + B.x) + C.x) + D.x, 0.f);
          r.y = fmaxf(((A.y + B.y) + C.y) + D.y, 0.f);
          r.z = fmaxf(((A.z + B.z) + C.z) + D.z, 0.f);
          r.w = fmaxf(((A.w + B.w) + C.w) + D.w, 0.f);
          xr[q] = r;
        }
        break;
      }
      case MK_POOL: {   // CSR order = the reference's scatter_add order (subject roles by ascending t, then object roles)
        const int beg = a.node_off[m], end = a.node_off[m + 1];
        const float cnt = fmaxf((float)(end - beg), 1.f);
        for (int q = lane; q < nq; q += 32) {
          const int k = seg0 + 4 * q;
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int it = beg; it < end; ++it) {
            const int item = a.node_items[it], t = item >> 1, role = item & 1;
            const float4 v = ld4(X + (long long)t * op.ldx + (role ? op.aux_i : 0) + k);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
          }
          acc.x /= cnt; acc.y /= cnt; acc.z /= cnt; acc.w /= cnt;
          xr[q] = acc;
        }
        break;
      }
      case MK_TEMB: {   // [cos(t f) | sin(t f)], the arithmetic of timestep_embedding_kernel (elem.cu)
        const int half = op.K >> 1;
        float* xs = Xs + i * KS;
        for (int k = lane; k < half; k += 32) {
          const float arg = __fmul_rn((float)a.t, __ldg(a.freqs + k));
          xs[k] = cosf(arg);
          xs[half + k] = sinf(arg);
        }
        break;
      }
      default: {   // MK_NONE / MK_SILU / MK_GEGLU, optional channel concat [X | X2]
        for (int q = lane; q < nq; q += 32) {
          const int k = seg0 + 4 * q;
          float4 v = (op.X2 && k >= op.K1) ? ld4(op.X2 + (long long)m * op.ldx2 + (k - op.K1)) : ld4(X + (long long)m * op.ldx + k);
          if (op.pro == MK_SILU) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
          if (op.pro == MK_GEGLU) {
            const float4 g = ld4(X + (long long)m * op.ldx + op.K + k);
            v.x *= gelu_erf(g.x); v.y *= gelu_erf(g.y); v.z *= gelu_erf(g.z); v.w *= gelu_erf(g.w);
          }
          xr[q] = v;
        }
        break;
      }
    }
  }
}

// acc[i*4 + j] += sum_k Xs[i][k] * Ws[(g*4 + j)][k] over this warp's K-chunks of the segment
template <int R>
__device__ __forceinline__ void fma_group(const float* Xs, int KS, const float* Ws, int K, int seg_len, int g, int feats, int warp, int lane,
                                          float (&acc)[R * 4]) {
  const int nch = (seg_len + 63) >> 6;
  for (int c = warp; c < nch; c += 8) {
    const int k = (c << 6) + 2 * lane;
    if (k < seg_len) {
      float2 xv[R];
#pragma unroll
      for (int i = 0; i < R; ++i) xv[i] = *reinterpret_cast<const float2*>(Xs + i * KS + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (g * 4 + j < feats) {
          const float2 wv = *reinterpret_cast<const float2*>(Ws + (size_t)(g * 4 + j) * K + k);
#pragma unroll
          for (int i = 0; i < R; ++i) acc[i * 4 + j] = fmaf(xv[i].x, wv.x, fmaf(xv[i].y, wv.y, acc[i * 4 + j]));
        }
      }
    }
  }
}

template <int R>
__device__ __forceinline__ void reduce_store(float (&acc)[R * 4], float* red, int lane) {
  if constexpr (R * 4 >= 32) {
    constexpr int N = R * 4;
    float v[N];
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] = i < R * 4 ? acc[i] : 0.f;
    const int base = butterfly<N>(v, lane);
#pragma unroll
    for (int e = 0; e < N / 32; ++e) red[base + e] = v[e];
  } else {
#pragma unroll
    for (int i = 0; i < R * 4; ++i) {
      float s = acc[i];
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) red[i] = s;
    }
  }
}

// one unit of a LIN op: rows [m0, m0+R) x features [n0, n0+feats)
template <int R>
__device__ void lin_unit(const MkOp& op, const MkArgs& a, float* Xs, const float* Ws, float* red_s, int m0, int n0, int feats, bool restage,
                         int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const int K = op.K, groups = (feats + 3) >> 2;
  constexpr int NG = R * 4;
  const int seg_max = ((MK_XCAP / R) >> 7) << 7;
  const int nseg = (K + seg_max - 1) / seg_max;
  const float* X = resolve_x(op, a);
  const int rows = min(R, op.M - m0);
  // epilogue operands requested before the contraction (their L2 latency hides under it)
  const int total = groups * NG;
  float pb[2] = {0.f, 0.f}, pr[2] = {0.f, 0.f};
  bool pv[2] = {false, false};
  int pm[2] = {0, 0}, pn[2] = {0, 0};
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int o = tid + 256 * e;
    if (o < total) {
      const int g = o / NG, rem = o - g * NG, i = rem >> 2, j = rem & 3;
      pv[e] = i < rows && g * 4 + j < feats;
      pm[e] = m0 + i;
      pn[e] = n0 + g * 4 + j;
      if (pv[e]) {
        if (op.bias) pb[e] = __ldg(op.bias + pn[e]);
        if (op.res) pr[e] = __ldcg(op.res + (long long)pm[e] * op.ld_res + pn[e]);
        if (op.res2) pr[e] += __ldcg(op.res2 + (long long)pm[e] * op.ld_res2 + pn[e]);
      }
    }
  }
  float acc[NG];
  if (nseg == 1) {
    const int KS = K;
    if (restage) stage_input<R>(op, a, X, Xs, KS, m0, 0, K, warp, lane);
    cons_sync();   // rows staged; also: everyone left the previous unit's epilogue (red_s is about to be rewritten)
    for (int g = 0; g < groups; ++g) {
#pragma unroll
      for (int i = 0; i < NG; ++i) acc[i] = 0.f;
      fma_group<R>(Xs, KS, Ws, K, K, g, feats, warp, lane, acc);
      reduce_store<R>(acc, red_s + (g * 8 + warp) * 64, lane);
    }
  } else {   // long rows (host guarantees one feature group): accumulators persist over the segments
#pragma unroll
    for (int i = 0; i < NG; ++i) acc[i] = 0.f;
    for (int sg = 0; sg < nseg; ++sg) {
      const int seg0 = sg * seg_max, seg_len = min(seg_max, K - seg0);
      if (sg) cons_sync();   // everyone is done reading the previous segment
      stage_input<R>(op, a, X, Xs, seg_max, m0, seg0, seg_len, warp, lane);
      cons_sync();
      fma_group<R>(Xs, seg_max, Ws + seg0, K, seg_len, 0, feats, warp, lane, acc);
    }
    reduce_store<R>(acc, red_s + warp * 64, lane);
  }
  cons_sync();
  float* Y = op.y_ext == MK_EXT_XPREV ? a.x_prev : op.Y;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int o = tid + 256 * e;
    if (o < total && pv[e]) {
      const int g = o / NG, rem = o - g * NG;
      const float* r = red_s + g * 8 * 64 + rem;
      float v = ((r[0] + r[64]) + (r[128] + r[192])) + ((r[256] + r[320]) + (r[384] + r[448]));
      v += pb[e];
      if (op.act == 1) v = fmaxf(v, 0.f);
      else if (op.act == 2) v = silu_f(v);
      v += pr[e];
      const int m = pm[e], n = pn[e];
      if (op.epi == MK_EPI_DDPM) {   // v = eps: x0 = a x - b eps; mean = c1 x0 + c2 x; + [t > 0] exp(0.5 logvar) noise (ddpm_update_kernel)
        const int t = a.t, T = a.T;
        const float ca = __ldg(a.tab + t), cb = __ldg(a.tab + T + t), c1 = __ldg(a.tab + 2 * T + t), c2 = __ldg(a.tab + 3 * T + t),
                    lv = __ldg(a.tab + 4 * T + t);
        const float sig = (t == 0 ? 0.f : 1.f) * expf(0.5f * lv);
        const long long idx = (long long)m * op.nout + n;
        const float x = __ldcg(a.x_t + idx);
        const float x0 = __fsub_rn(__fmul_rn(ca, x), __fmul_rn(cb, v));
        const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x));
        Y[idx] = __fadd_rn(mean, __fmul_rn(sig, __ldcg(a.noise + idx)));
      } else if (op.bcast_rows > 0) {
        for (int rr = 0; rr < op.bcast_rows; ++rr) Y[(long long)rr * op.ldy + n] = v;
      } else {
        Y[(long long)m * op.ldy + n] = v;
      }
    }
  }
}

__device__ __forceinline__ int first_unit(const MkOp& op, int cta, int G) {
  int f = (cta - op.unit_begin % G) % G;
  return f < 0 ? f + G : f;
}

__global__ void __launch_bounds__(MK_THREADS, 1) layout_mk_kernel(const MkArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* Xs = reinterpret_cast<float*>(smem);
  uint8_t* Wslots = smem + SM_W;
  float* red_s = reinterpret_cast<float*>(smem + SM_RED);
  MkOp* ops_s = reinterpret_cast<MkOp*>(smem + SM_OPS);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint64_t* empty_bar = full_bar + MK_SLOTS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  if (tid == 0) {
    for (int i = 0; i < MK_SLOTS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const unsigned epoch = *a.epoch + 1u;
  const unsigned target = epoch * (unsigned)G;

  if (warp == 8) {
    // ================= producer: this CTA's weight slices, in program order, as far ahead as the ring allows =================
    if (lane == 0) {
      unsigned seq = 0;
      for (int s = 0; s < a.n_stages; ++s) {
        const MkStage st = a.stages[s];
        for (int oi = 0; oi < st.n_a + st.n_b; ++oi) {
          const MkOp* op = a.ops + st.op_begin + oi;
          if (op->type != MK_T_LIN) continue;
          const int units = op->units, FU = op->FU, n_slices = op->n_slices, K = op->K, nout = op->nout;
          const float* W = op->W;
          for (int u = first_unit(*op, cta, G); u < units; u += G) {
            const int n0 = (u % n_slices) * FU;
            const int feats = min(FU, nout - n0);
            const unsigned slot = seq % MK_SLOTS, ph = (seq / MK_SLOTS) & 1u;
            mbar_wait(&empty_bar[slot], ph ^ 1u);
            const uint32_t bytes = (uint32_t)feats * (uint32_t)K * 4u;
            mbar_arrive_expect_tx(&full_bar[slot], bytes);
            bulk_g2s(Wslots + slot * MK_SLOT_BYTES, W + (size_t)n0 * K, bytes, &full_bar[slot]);
            ++seq;
          }
        }
      }
    }
    return;
  }

  // ================= consumers =================
  unsigned seq = 0;
  for (int s = 0; s < a.n_stages; ++s) {
    const MkStage st = a.stages[s];
    const int n_ops = st.n_a + st.n_b;
    {   // this stage's op records -> shared memory (constants: fetched while the barrier is still filling)
      const uint4* src = reinterpret_cast<const uint4*>(a.ops + st.op_begin);
      uint4* dst = reinterpret_cast<uint4*>(ops_s);
      for (int i = tid; i < n_ops * 16; i += 256) dst[i] = __ldg(src + i);
    }
    if (tid == 0) {
      if (s > 0) wait_counter(a.bar + s - 1, target, a.err);
      if (st.bg_wait >= 0) wait_counter(a.bg + st.bg_wait, target, a.err);
    }
    cons_sync();
    int staged = -1;
    for (int oi = 0; oi <= n_ops; ++oi) {
      if (oi == st.n_a) {   // the stage's foreground ops are done in this CTA
        cons_sync();
        if (tid == 0) arrive_counter(a.bar + s);
      }
      if (oi == n_ops) break;
      const MkOp& op = ops_s[oi];
      if (op.type == MK_T_LIN) {
        for (int u = first_unit(op, cta, G); u < op.units; u += G) {
          const int slice = u % op.n_slices, rt = u / op.n_slices;
          const int n0 = slice * op.FU, feats = min(op.FU, op.nout - n0);
          const unsigned slot = seq % MK_SLOTS, ph = (seq / MK_SLOTS) & 1u;
          const int key = oi * 64 + rt;
          const bool restage = staged != key;
          mbar_wait(&full_bar[slot], ph);
          const float* Ws = reinterpret_cast<const float*>(Wslots + slot * MK_SLOT_BYTES);
          if (op.rclass == 16) lin_unit<16>(op, a, Xs, Ws, red_s, rt * 16, n0, feats, restage, tid);
          else if (op.rclass == 8) lin_unit<8>(op, a, Xs, Ws, red_s, 0, n0, feats, restage, tid);
          else lin_unit<1>(op, a, Xs, Ws, red_s, 0, n0, feats, restage, tid);
          // every warp passed the barrier in front of the epilogue, i.e. finished reading the slot
          if (lane == 0) mbar_arrive(&empty_bar[slot]);
          const int seg_max = ((MK_XCAP / op.rclass) >> 7) << 7;
          staged = op.K <= seg_max ? key : -1;
          ++seq;
        }
      } else {
        const float* X = resolve_x(op, a);
        for (int u = first_unit(op, cta, G); u < op.units; u += G) {
          const int m0 = u * 16, rows = min(16, op.M - m0), kq = op.K >> 2;
          for (int e = tid; e < rows * kq; e += 256) {
            const int i = e / kq, q = e - i * kq, m = m0 + i;
            float4 v;
            if (op.type == MK_T_COPY) v = ld4(X + (long long)m * op.ldx + 4 * q);
            else v = __ldg(reinterpret_cast<const float4*>(op.aux0 + (long long)a.triples[(long long)m * 3 + 1] * op.K + 4 * q));
            *reinterpret_cast<float4*>(op.Y + (long long)m * op.ldy + 4 * q) = v;
          }
        }
        staged = -1;
      }
    }
    if (st.bg_arrive >= 0) {
      cons_sync();
      if (tid == 0) arrive_counter(a.bg + st.bg_arrive);
    }
    cons_sync();   // ops_s is rewritten by the next stage
  }
  if (cta == 0 && tid == 0) *a.epoch = epoch;
}

struct MkState {
  bool checked = false, ok = false;
  int ctas = 0;
} g_mk;

}  // namespace

bool mk_available(int* ctas_out) {
  if (!g_mk.checked) {
    g_mk.checked = true;
    int dev = 0, coop = 0, sms = 0, occ = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        cudaFuncSetAttribute(layout_mk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL) == cudaSuccess &&
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, layout_mk_kernel, MK_THREADS, SM_TOTAL) == cudaSuccess && occ >= 1) {
      g_mk.ok = true;
      g_mk.ctas = sms;
    } else {
      cudaGetLastError();
    }
  }
  if (ctas_out) *ctas_out = g_mk.ctas;
  return g_mk.ok;
}

void mk_plan_op(MkOp& op, int ctas) {
  if (op.type != MK_T_LIN) {
    ECHO_CHECK(op.K % 4 == 0 && op.ldx % 4 == 0 && op.ldy % 4 == 0, "layout program: copy op needs 16-byte rows");
    op.rclass = 16; op.row_tiles = cdiv(op.M, 16); op.FU = 0; op.n_slices = 1; op.units = op.row_tiles;
    return;
  }
  ECHO_CHECK(op.M >= 1 && op.K >= 4 && op.K % 4 == 0 && op.nout >= 1 && op.ldx % 4 == 0, "layout program: bad linear op (M=%d K=%d nout=%d)", op.M,
             op.K, op.nout);
  op.rclass = op.M == 1 ? 1 : (op.M <= 8 ? 8 : 16);
  op.row_tiles = op.rclass == 16 ? cdiv(op.M, 16) : 1;
  int cap = (MK_SLOT_BYTES / (op.K * 4)) / 4 * 4;
  ECHO_CHECK(cap >= 4, "layout program: K=%d too long for a weight slot", op.K);
  cap = cap > MK_MAX_FU ? MK_MAX_FU : cap;
  const int seg_max = ((MK_XCAP / op.rclass) >> 7) << 7;
  if (op.K > seg_max) {
    ECHO_CHECK(op.pro == MK_NONE || op.pro == MK_SILU || op.pro == MK_GEGLU, "layout program: prologue %d needs the whole row staged (K=%d)", op.pro, op.K);
    cap = 4;
  }
  if (op.pro == MK_GN) ECHO_CHECK(op.K % 128 == 0 && op.cpg >= 4 && op.cpg <= 128 && (op.cpg & (op.cpg - 1)) == 0, "layout program: GroupNorm prologue K=%d cpg=%d", op.K, op.cpg);
  if (op.pro == MK_LN) ECHO_CHECK(op.K <= 1280, "layout program: LayerNorm prologue K=%d", op.K);
  if (op.X2) ECHO_CHECK(op.K1 % 4 == 0 && op.K1 > 0 && op.K1 < op.K && op.ldx2 % 4 == 0 && op.pro != MK_GEGLU && op.pro != MK_LN, "layout program: bad concat input");
  if (op.bcast_rows > 0) ECHO_CHECK(op.M == 1, "layout program: broadcast store needs a one-row op");
  int want = 4 * cdiv((int64_t)op.nout * op.row_tiles, 4 * (int64_t)ctas);
  want = want < 4 ? 4 : want;
  op.FU = want > cap ? cap : want;
  op.n_slices = cdiv(op.nout, op.FU);
  op.units = op.n_slices * op.row_tiles;
}

void mk_launch(const MkArgs& a, int ctas, cudaStream_t s) {
  MkArgs args = a;
  void* params[] = {(void*)&args};
  ECHO_CUDA(cudaLaunchCooperativeKernel((const void*)layout_mk_kernel, dim3(ctas), dim3(MK_THREADS), params, (size_t)SM_TOTAL, s));
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
