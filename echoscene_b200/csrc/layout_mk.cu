// Persistent executor of one layout DDPM iteration: UNet1DModel.forward (denoise_net.py:773-806, incl. box_messsage_passing
// :758-771 and the GraphTripleConvNet of graph.py:124-250) + the posterior update (diffusion_ddpm.py:220-309) as ONE kernel.
//
// Why: the layout step is a chain of ~120 dependent few-row layers (rows = nodes / triples, 8..64 of them) whose cost as
// separate launches is the chain of launch + first-touch latencies (1.56 ms for 461 MB of weights = 4.5 % of the HBM
// roofline, r1).  Here the step is a PROGRAM (layout.cu records it once per node/triple count): stages of independent ops,
// each op cut into units (16 rows x FU output features); one cooperative grid of one CTA per SM walks the stages.
//
//   * weights never wait for a barrier: they are constants, so every CTA streams its future weight slices HBM -> shared
//     memory with cp.async.bulk (1-D TMA, one copy per feature row into a bank-conflict-free padded layout) through a
//     3-slot ring (108 KB in flight per SM), two to three stages ahead of the compute; the ring is refilled by one warp
//     while the CTA waits at the stage barrier (the slot of a finished unit is free by program order: no empty-barriers);
//   * a stage boundary is one arrival counter in L2 (release: bar.sync + fence + atomicAdd; acquire: one polling thread +
//     bar.sync), ~1 us instead of a kernel boundary; activations are read with ld.global.cg (they are rewritten every step by
//     other SMs, L1 must not serve them);
//   * the elementwise op in front of a Linear is a prologue applied while its input rows are staged in shared memory (SiLU,
//     GroupNorm via lane shuffles, LayerNorm per warp-row, the GraphTripleConv edge gather-combine with the node features
//     staged per edge row, and the CSR mean pooling in the reference's summation order); GEGLU is the epilogue of ff1 (each
//     unit carries the value row and the gate row of its features), so a stage is exactly one dependent contraction;
//   * the time-embedding path (time MLP, the 22 stacked emb_layers projections = 40 % of the weight bytes) is computed for
//     ONE row (all nodes of a step share t) and runs as background ops in the barrier shadow of the GCN stages.
//
// Contraction of a unit (16 consumer warps): 16 staged rows X_s [16][K] x weight slice W_s [<= 24][K], both K-contiguous in
// shared memory with a 16-float row pad.  The K range is dealt to the warps in steps of 16; a step is two
// mma.sync.m16n8k8 per 8 features in split precision (3xTF32: x = hi + lo, hi.hi + lo.hi + hi.lo into an fp32 accumulator,
// small terms first -- fp32-grade products, the tensor core does the k-reduction), fed by one LDS.128 per fragment (the k
// index inside a step is permuted identically for A and B so that a thread's four k values are contiguous).  The warps'
// partial tiles meet in shared memory and are summed in fixed order (deterministic: no atomics on data).  One-row ops (the
// time path) are plain fp32 dot products, a warp group per feature.
#include "layout_mk.cuh"

#include "tc_ptx.cuh"

namespace echo {
namespace {

using namespace ptx;

constexpr int MK_CW = 16;                  // warps (all of them compute; the last one also feeds the weight ring)
constexpr int MK_CT = MK_CW * 32;
constexpr int MK_THREADS = MK_CT;          // 512 threads: 128 registers each
constexpr int MK_SLOTS = 3;
constexpr int MK_XSTRIDE = MK_XROW + MK_PAD;             // floats between staged rows
constexpr int SM_X = 16 * MK_XSTRIDE * 4;
constexpr int SM_RED = SM_X + MK_SLOTS * MK_SLOT_BYTES;
constexpr int SM_OPS = SM_RED + MK_CW * 16 * MK_MAX_FU * 4;
constexpr int SM_BAR = SM_OPS + 2 * MK_MAX_STAGE_OPS * 256;   // op records of the current and the next stage
constexpr int SM_FEED = SM_BAR + 64;
constexpr int SM_STG = SM_FEED + 320;                          // every stage record of the program (8 bytes each)
constexpr int SM_TOTAL = SM_STG + MK_MAX_STAGES * 8;
static_assert(SM_TOTAL <= 232448, "shared memory budget of one CTA");

// x * sigmoid(x) with the fast exponential / reciprocal (relative error ~1e-6; the prologue runs redundantly in every CTA)
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// Activations are plain (weak) loads.  They were written by other SMs in an earlier stage; the stage barrier's polling thread
// acquires at gpu scope (LDG.STRONG.GPU + CCTL.IVALL: the SM's L1 is invalidated) and the CTA barrier behind it orders every
// other thread after that acquire, so no stale line can be served -- the pattern of cooperative_groups' grid.sync().
// (ld.global.cg compiles to LDG.STRONG.GPU, which measured ~4x slower for the 32 KB staging burst of a CTA.)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float ld1(const float* p) { return *p; }
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
// release: the CTA's stores (ordered before this thread by the preceding bar.sync) become visible before the count does
__device__ __forceinline__ void arrive_counter(unsigned* p, bool fenced) {
  if (fenced) {
    __threadfence();
    atomicAdd(p, 1u);
  } else {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(1u) : "memory");
  }
}
// weights are read once per step: evict-first keeps the 439 MB stream from pushing the few MB of activations out of L2
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_policy(bool evict_first) {
  uint64_t p;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
// x = hi + lo with hi = x truncated to TF32 (exact subtraction, |lo| < 2^-10 |x|); the tensor core reads the upper 19 bits
// of an operand register, i.e. truncates lo itself.  hi.hi + lo.hi + hi.lo then drops ~2^-20 |x w| per product.
// (cvt.rna.tf32.f32 expands to four instructions on this target; truncation is one.)
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ const float* resolve_x(const MkOp& op, const MkArgs& a) {
  return op.x_ext == MK_EXT_XT ? a.x_t : op.x_ext == MK_EXT_OBJ ? a.obj_embed : op.x_ext == MK_EXT_TNODE ? a.tnode_row : op.X;
}

// One input row of an op, resolved into registers once per unit (the op record itself lives in shared memory, and every
// shared-memory store in between would make the compiler read its fields again)
struct RowIn {
  const float* x;    // row m of X
  const float* x2;   // row m of X2, shifted so that x2[k] is column k of the concatenation (null: no concat)
  int K1;
};
__device__ __forceinline__ RowIn row_in(const MkOp& op, const float* X, int m) {
  RowIn r;
  r.x = X + (long long)m * op.ldx;
  r.x2 = op.X2 ? op.X2 + (long long)m * op.ldx2 - op.K1 : nullptr;
  r.K1 = op.K1;
  return r;
}
__device__ __forceinline__ float4 load_in(const RowIn& r, int k) { return (r.x2 && k >= r.K1) ? ld4(r.x2 + k) : ld4(r.x + k); }

// LayerNorm of one row by one warp, NJ float4 per lane (row length <= 128 * NJ)
template <int NJ>
__device__ __forceinline__ void ln_row(const MkOp& op, const float* X, const float* gs, const float* bs, int m, int nq, float4* xr, int lane) {
  float4 v[NJ];
  float s = 0.f;
  const float* xrow = X + (long long)m * op.ldx;
  const float inv_k = 1.f / (float)op.K, eps = op.eps;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int q = lane + 32 * j;
    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < nq) v[j] = ld4(xrow + 4 * q);
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s * inv_k;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    if (lane + 32 * j < nq) {
      const float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
      ss += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss * inv_k + eps);
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int q = lane + 32 * j;
    if (q < nq) {
      const float4 gm = *reinterpret_cast<const float4*>(gs + 4 * q), bt = *reinterpret_cast<const float4*>(bs + 4 * q);
      float4 o4;
      o4.x = (v[j].x - mean) * rstd * gm.x + bt.x; o4.y = (v[j].y - mean) * rstd * gm.y + bt.y;
      o4.z = (v[j].z - mean) * rstd * gm.z + bt.z; o4.w = (v[j].w - mean) * rstd * gm.w + bt.w;
      xr[q] = o4;
    }
  }
}

// rows m0 .. m0+16 of the op's input, prologue applied, columns [seg0, seg0 + seg_len) -> Xs [16][MK_XSTRIDE]; one warp per
// row, every global load of a row in flight before the first use
// sum over the gl (power of two <= 32) adjacent lanes of a group, same order as a doubling butterfly
__device__ __forceinline__ float group_sum(float s, int gl) {
  if (gl > 1) s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (gl > 2) s += __shfl_xor_sync(0xffffffffu, s, 2);
  if (gl > 4) s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (gl > 8) s += __shfl_xor_sync(0xffffffffu, s, 8);
  if (gl > 16) s += __shfl_xor_sync(0xffffffffu, s, 16);
  return s;
}

// gs / bs: the prologue's scale and shift rows (they arrive with the unit's weight slice)
__device__ void stage_rows(const MkOp& op, const MkArgs& a, const float* X, float* Xs, const float* gs, const float* bs, int m0, int seg0,
                           int seg_len, int warp, int lane) {
  const int nq = seg_len >> 2;
  const int m = m0 + warp;
  float4* xr = reinterpret_cast<float4*>(Xs + warp * MK_XSTRIDE);
  if (m >= op.M) {
    for (int q = lane; q < nq; q += 32) xr[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int pro = op.pro;
  const RowIn in = row_in(op, X, m);
  switch (pro) {
    case MK_GN: {   // K % 128 == 0: every lane is in range in every iteration, a group is cpg/4 adjacent lanes
      const int gl = op.cpg >> 2;
      const float inv = 1.f / (float)op.cpg, eps = op.eps;
      const bool act = op.pro_act != 0;
      for (int q0 = 0; q0 < nq; q0 += 128) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = q0 + lane + 32 * j;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < nq) v[j] = load_in(in, seg0 + 4 * q);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = q0 + lane + 32 * j;
          if (q < nq) {   // warp-uniform (nq % 32 == 0)
            const int k = seg0 + 4 * q;
            const float4 gm = *reinterpret_cast<const float4*>(gs + k), bt = *reinterpret_cast<const float4*>(bs + k);
            float s = (v[j].x + v[j].y) + (v[j].z + v[j].w);
            s = group_sum(s, gl);
            const float mean = s * inv;
            const float d0 = v[j].x - mean, d1 = v[j].y - mean, d2 = v[j].z - mean, d3 = v[j].w - mean;
            float ss = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            ss = group_sum(ss, gl);
            const float rstd = rsqrtf(ss * inv + eps);
            float4 r;
            r.x = d0 * rstd * gm.x + bt.x; r.y = d1 * rstd * gm.y + bt.y; r.z = d2 * rstd * gm.z + bt.z; r.w = d3 * rstd * gm.w + bt.w;
            if (act) { r.x = silu_f(r.x); r.y = silu_f(r.y); r.z = silu_f(r.z); r.w = silu_f(r.w); }
            xr[q] = r;
          }
        }
      }
      break;
    }
    case MK_LN: {   // whole row by this warp: two-pass statistics in registers (K <= 512: four vectors per lane, else up to eight)
      if (nq <= 128) ln_row<4>(op, X, gs, bs, m, nq, xr, lane);
      else ln_row<8>(op, X, gs, bs, m, nq, xr, lane);
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
      const long long ldx = op.ldx;
      const int aux_i = op.aux_i;
      for (int q0 = 0; q0 < nq; q0 += 64) {   // two float4 columns per lane and pass (H = 256: one pass)
        const int qa = q0 + lane, qb = q0 + lane + 32;
        float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
        for (int base = beg; base < end; base += 32) {
          const int mine = base + lane < end ? a.node_items[base + lane] : 0;
          const int nb = min(32, end - base);
#pragma unroll 4
          for (int e = 0; e < nb; ++e) {
            const int item = __shfl_sync(0xffffffffu, mine, e), t = item >> 1, role = item & 1;
            const float* src = X + (long long)t * ldx + (role ? aux_i : 0) + seg0;
            if (qa < nq) { const float4 v = ld4(src + 4 * qa); acc0.x += v.x; acc0.y += v.y; acc0.z += v.z; acc0.w += v.w; }
            if (qb < nq) { const float4 v = ld4(src + 4 * qb); acc1.x += v.x; acc1.y += v.y; acc1.z += v.z; acc1.w += v.w; }
          }
        }
        if (qa < nq) xr[qa] = make_float4(acc0.x / cnt, acc0.y / cnt, acc0.z / cnt, acc0.w / cnt);
        if (qb < nq) xr[qb] = make_float4(acc1.x / cnt, acc1.y / cnt, acc1.z / cnt, acc1.w / cnt);
      }
      break;
    }
    default: {   // MK_NONE / MK_SILU, optional channel concat [X | X2]
      for (int q0 = 0; q0 < nq; q0 += 128) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = q0 + lane + 32 * j;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < nq) v[j] = load_in(in, seg0 + 4 * q);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int q = q0 + lane + 32 * j;
          if (q < nq) {
            if (pro == MK_SILU) { v[j].x = silu_f(v[j].x); v[j].y = silu_f(v[j].y); v[j].z = silu_f(v[j].z); v[j].w = silu_f(v[j].w); }
            xr[q] = v[j];
          }
        }
      }
      break;
    }
  }
}

// the single input row of a one-row op (the time path) -> Xs[0 .. K), all consumer threads
__device__ void stage_row1(const MkOp& op, const MkArgs& a, const float* X, float* Xs, int tid) {
  if (op.pro == MK_TEMB) {   // [cos(t f) | sin(t f)], the arithmetic of timestep_embedding_kernel (elem.cu)
    const int half = op.K >> 1;
    for (int k = tid; k < half; k += MK_CT) {
      const float arg = __fmul_rn((float)a.t, __ldg(a.freqs + k));
      Xs[k] = cosf(arg);
      Xs[half + k] = sinf(arg);
    }
    return;
  }
  const int nq = op.K >> 2;
  for (int q = tid; q < nq; q += MK_CT) {
    float4 v = ld4(X + 4 * q);
    if (op.pro == MK_SILU) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
    reinterpret_cast<float4*>(Xs)[q] = v;
  }
}

// the epilogue of one output element (m, n): v = the dot product; eb = bias, er = residual(s), both requested before the
// contraction so that their L2 latency hides under it
__device__ __forceinline__ void epi_store(const MkOp& op, const MkArgs& a, int em, int en, float eb, float er, float v) {
  v += eb;
  if (op.act == 1) v = fmaxf(v, 0.f);
  else if (op.act == 2) v = v / (1.f + expf(-v));
  v += er;
  float* Y = op.y_ext == MK_EXT_XPREV ? a.x_prev : op.Y;
  if (op.epi == MK_EPI_DDPM) {   // v = eps: x0 = a x - b eps; mean = c1 x0 + c2 x; + [t > 0] exp(0.5 logvar) noise (ddpm_update_kernel)
    const int t = a.t, T = a.T;
    const float ca = __ldg(a.tab + t), cb = __ldg(a.tab + T + t), c1 = __ldg(a.tab + 2 * T + t), c2 = __ldg(a.tab + 3 * T + t),
                lv = __ldg(a.tab + 4 * T + t);
    const float sig = (t == 0 ? 0.f : 1.f) * expf(0.5f * lv);
    const long long idx = (long long)em * op.nout + en;
    const float x = ld1(a.x_t + idx);
    const float x0 = __fsub_rn(__fmul_rn(ca, x), __fmul_rn(cb, v));
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x));
    Y[idx] = __fadd_rn(mean, __fmul_rn(sig, ld1(a.noise + idx)));
  } else if (op.bcast_rows > 0) {
    for (int rr = 0; rr < op.bcast_rows; ++rr) Y[(long long)rr * op.ldy + en] = v;
  } else {
    Y[(long long)em * op.ldy + en] = v;
  }
}

// one unit of a 16-row LIN op: rows [m0, m0+16) x features [n0, n0+feats)
__device__ void lin_unit16(const MkOp& op, const MkArgs& a, float* Xs, const float* Ws, float* red_s, int m0, int n0, int feats, bool restage,
                           int tid, long long* dbg) {
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int K = op.K, ws = K + MK_PAD;
  const bool geglu = op.epi == MK_EPI_GEGLU;
  const int srows = geglu ? 2 * feats : feats;     // weight rows in the slot
  const int nt = (srows + 7) >> 3;                   // 8-feature MMA tiles (<= 3)
  const int nseg = (K + MK_XROW - 1) / MK_XROW;
  const float* X = resolve_x(op, a);
  const int ei = (int)(((float)tid + 0.5f) * __frcp_rn((float)feats)), ej = tid - ei * feats;   // this thread's output element (tid / feats)
  const int em = m0 + ei, en = n0 + ej;
  const bool evalid = tid < 16 * feats && em < op.M;
  float eb = 0.f, eb2 = 0.f, er = 0.f;
  if (evalid) {
    if (op.bias) { eb = __ldg(op.bias + en); if (geglu) eb2 = __ldg(op.bias + op.nout + en); }
    if (op.res_ext) er = __ldg(a.emb_row + op.aux_i + en);
    else if (op.res) er = ld1(op.res + (long long)em * op.ld_res + en);
    if (op.res2) er += ld1(op.res2 + (long long)em * op.ld_res2 + en);
  }
  // two accumulator chains per tile (hi.hi / the two cross terms): half the dependent-MMA depth, and the small terms meet first
  float acc[3][4], acs[3][4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    acs[i][0] = acs[i][1] = acs[i][2] = acs[i][3] = 0.f;
  }
  for (int sg = 0; sg < nseg; ++sg) {
    const int seg0 = sg * MK_XROW, seg_len = min(MK_XROW, K - seg0);
    if (restage || nseg > 1) {
      if (sg) cons_sync();   // everyone is done reading the previous segment
      stage_rows(op, a, X, Xs, Ws + (size_t)srows * ws, Ws + (size_t)(srows + 1) * ws, m0, seg0, seg_len, warp, lane);
      if (dbg && tid == 0) dbg[11] = clock64();
    }
    cons_sync();   // rows staged; also: everyone left the previous unit's epilogue (red_s is about to be rewritten)
    if (dbg && tid == 0) dbg[5] = clock64();
    const int nks = (seg_len + 15) >> 4;
    const float* xa_p = Xs + g * MK_XSTRIDE + 4 * t;
    const float* xb_p = xa_p + 8 * MK_XSTRIDE;
    const float* w_p = Ws + (size_t)g * ws + seg0 + 4 * t;
    for (int ks = warp; ks < nks; ks += MK_CW) {
      const int kk = ks << 4;
      const bool kin = kk + 4 * t < seg_len;   // K % 16 != 0 tail: this thread's four k values are past the row and contribute zeros
      float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;   // (the whole warp stays in the loop: mma.sync is warp-wide)
      if (kin) { xa = *reinterpret_cast<const float4*>(xa_p + kk); xb = *reinterpret_cast<const float4*>(xb_p + kk); }
      uint32_t ah[8], al[8];
      split_tf32(xa.x, ah[0], al[0]); split_tf32(xb.x, ah[1], al[1]); split_tf32(xa.y, ah[2], al[2]); split_tf32(xb.y, ah[3], al[3]);
      split_tf32(xa.z, ah[4], al[4]); split_tf32(xb.z, ah[5], al[5]); split_tf32(xa.w, ah[6], al[6]); split_tf32(xb.w, ah[7], al[7]);
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        if (i < nt) {
          float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kin && i * 8 + g < srows) w = *reinterpret_cast<const float4*>(w_p + (size_t)i * 8 * ws + kk);
          uint32_t bh[4], bl[4];
          split_tf32(w.x, bh[0], bl[0]); split_tf32(w.y, bh[1], bl[1]); split_tf32(w.z, bh[2], bl[2]); split_tf32(w.w, bh[3], bl[3]);
          mma_tf32(acs[i], al[0], al[1], al[2], al[3], bh[0], bh[1]);
          mma_tf32(acc[i], ah[0], ah[1], ah[2], ah[3], bh[0], bh[1]);
          mma_tf32(acs[i], ah[0], ah[1], ah[2], ah[3], bl[0], bl[1]);
          mma_tf32(acc[i], ah[4], ah[5], ah[6], ah[7], bh[2], bh[3]);
          mma_tf32(acs[i], al[4], al[5], al[6], al[7], bh[2], bh[3]);
          mma_tf32(acs[i], ah[4], ah[5], ah[6], ah[7], bl[2], bl[3]);
        }
      }
    }
  }
  // partial tiles -> red_s [warp][16 rows][8 * nt columns]
  {
    const int rw = 8 * nt;
    float* r = red_s + warp * 16 * MK_MAX_FU;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (i < nt) {
        *reinterpret_cast<float2*>(r + g * rw + i * 8 + 2 * t) = make_float2(acc[i][0] + acs[i][0], acc[i][1] + acs[i][1]);
        *reinterpret_cast<float2*>(r + (g + 8) * rw + i * 8 + 2 * t) = make_float2(acc[i][2] + acs[i][2], acc[i][3] + acs[i][3]);
      }
    }
  }
  cons_sync();
  if (dbg && tid == 0) dbg[6] = clock64();
  if (evalid) {
    const int rw = 8 * nt;
    const float* r = red_s + ei * rw + ej;
    float v = 0.f, gt = 0.f;
#pragma unroll
    for (int w = 0; w < MK_CW; ++w) v += r[w * 16 * MK_MAX_FU];
    if (geglu) {
#pragma unroll
      for (int w = 0; w < MK_CW; ++w) gt += r[w * 16 * MK_MAX_FU + feats];
      v = (v + eb) * gelu_erf(gt + eb2);
      eb = 0.f;
    }
    epi_store(op, a, em, en, eb, er, v);
  }
  if (dbg && tid == 0) dbg[7] = clock64();
}

// one unit of a one-row LIN op: features [n0, n0+feats) of the row staged in Xs; feats <= 16, a group of 16 / feats warps per feature
__device__ void lin_unit1(const MkOp& op, const MkArgs& a, float* Xs, const float* Ws, float* red_s, int n0, int feats, bool restage, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  const int K = op.K, ws = K + MK_PAD;
  const float* X = resolve_x(op, a);
  const int en = n0 + tid;
  float eb = 0.f, er = 0.f;
  if (tid < feats) {
    if (op.bias) eb = __ldg(op.bias + en);
    if (op.res) er = ld1(op.res + en);
    if (op.res2) er += ld1(op.res2 + en);
  }
  if (restage) stage_row1(op, a, X, Xs, tid);
  cons_sync();
  const int parts = MK_CW / feats;       // warps per feature
  const int j = warp % feats, part = warp / feats;
  if (part < parts) {
    const int nq = K >> 2;
    const float4* x4 = reinterpret_cast<const float4*>(Xs);
    const float4* w4 = reinterpret_cast<const float4*>(Ws + (size_t)j * ws);
    float s0 = 0.f, s1 = 0.f;
    int q = part * 32 + lane;
    for (; q + parts * 32 < nq; q += 2 * parts * 32) {
      const float4 xa = x4[q], wa = w4[q], xb = x4[q + parts * 32], wb = w4[q + parts * 32];
      s0 = fmaf(xa.x, wa.x, fmaf(xa.y, wa.y, fmaf(xa.z, wa.z, fmaf(xa.w, wa.w, s0))));
      s1 = fmaf(xb.x, wb.x, fmaf(xb.y, wb.y, fmaf(xb.z, wb.z, fmaf(xb.w, wb.w, s1))));
    }
    if (q < nq) {
      const float4 xa = x4[q], wa = w4[q];
      s0 = fmaf(xa.x, wa.x, fmaf(xa.y, wa.y, fmaf(xa.z, wa.z, fmaf(xa.w, wa.w, s0))));
    }
    float s = s0 + s1;
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red_s[part * 16 + j] = s;
  }
  cons_sync();
  if (tid < feats) {
    float v = 0.f;
    for (int p = 0; p < parts; ++p) v += red_s[p * 16 + tid];
    epi_store(op, a, 0, en, eb, er, v);
  }
}

// first unit of an op that CTA `cta` runs: units are dealt round-robin starting where the previous op of the stage stopped
// (ufirst = unit_begin % G, precomputed by the host: no integer division on the device)
__device__ __forceinline__ int first_unit(int ufirst, int cta, int G) {
  const int f = cta - ufirst;
  return f < 0 ? f + G : f;
}

// The weight ring's feeder (shared memory; only warp MK_CW - 1 touches it).  The host lists every CTA's weight fetches in
// consumption order (mk_build_fetch); the next MK_FETCH_AHEAD descriptors sit in shared memory, fetched with cp.async, so
// that issuing a unit costs no global-memory round trip on the CTA's critical path.
constexpr int MK_FETCH_AHEAD = 4;
struct Feeder {
  MkFetch desc[MK_FETCH_AHEAD];
  const MkFetch* list;   // this CTA's descriptors
  unsigned total;        // how many
  unsigned issued;       // units issued so far == sequence number of the next one
  long long issue_clk[8];   // diagnostics: SM clock at which unit q was issued, [q % 8]
};
static_assert(sizeof(MkFetch) == 48, "a fetch descriptor is three 16-byte cp.async transfers");

__device__ __forceinline__ void feeder_prefetch(Feeder* f, unsigned q, int lane) {   // descriptor q -> desc[q % AHEAD]
  if (q < f->total && lane < 3)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(reinterpret_cast<uint8_t*>(&f->desc[q % MK_FETCH_AHEAD]) + 16 * lane)),
                 "l"(reinterpret_cast<const uint8_t*>(f->list + q) + 16 * lane)
                 : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// Issue this CTA's next weight slices until the ring is full: unit q goes to slot q % MK_SLOTS, which is free once unit
// q - MK_SLOTS is complete (`completed` units are: the caller has passed their last shared-memory read in program order).
// One bulk copy per feature row into the padded slot layout; the GroupNorm / LayerNorm scale and shift of the op's prologue ride
// along as two more rows.  Called by the whole feeder warp.
__device__ __noinline__ void feed(Feeder* f, unsigned completed, uint8_t* Wslots, uint64_t* full_bar, int lane, uint64_t policy) {
  unsigned issued = f->issued;
  const unsigned total = f->total;
  if (issued >= completed + MK_SLOTS || issued >= total) return;
  while (issued < completed + MK_SLOTS && issued < total) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    const MkFetch d = f->desc[issued % MK_FETCH_AHEAD];
    const unsigned slot = issued % MK_SLOTS;
    const int srows = d.rows0 + d.rows1, all = srows + d.naux;
    if (lane == 0) mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)all * (uint32_t)d.K * 4u);
    __syncwarp();
    if (lane < all) {
      const float* src = lane < d.rows0 ? d.w + (size_t)lane * d.K
                         : lane < srows ? d.g + (size_t)(lane - d.rows0) * d.K
                         : lane == srows ? d.aux : d.aux2;
      bulk_g2s(Wslots + slot * MK_SLOT_BYTES + (size_t)lane * (d.K + MK_PAD) * 4, src, (uint32_t)d.K * 4u, &full_bar[slot], policy);
    }
    if (lane == 0) f->issue_clk[issued & 7] = clock64();
    __syncwarp();   // every lane has read the descriptor: its place may be refilled
    feeder_prefetch(f, issued + MK_FETCH_AHEAD, lane);
    ++issued;
  }
  __syncwarp();
  if (lane == 0) f->issued = issued;
  __syncwarp();
}

__global__ void __launch_bounds__(MK_THREADS, 1) layout_mk_kernel(const MkArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  float* Xs = reinterpret_cast<float*>(smem);
  uint8_t* Wslots = smem + SM_X;
  float* red_s = reinterpret_cast<float*>(smem + SM_RED);
  MkOp* ops_s = reinterpret_cast<MkOp*>(smem + SM_OPS);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  Feeder* feeder = reinterpret_cast<Feeder*>(smem + SM_FEED);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x, cta = blockIdx.x;
  if (tid == 0) {
    for (int i = 0; i < MK_SLOTS; ++i) mbar_init(&full_bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    feeder->list = a.fetch + a.fetch_off[cta];
    feeder->total = (unsigned)(a.fetch_off[cta + 1] - a.fetch_off[cta]);
    feeder->issued = 0;
  }
  __syncthreads();
  if (warp == MK_CW - 1)
    for (int q = 0; q < MK_FETCH_AHEAD; ++q) feeder_prefetch(feeder, q, lane);
  const unsigned epoch = *a.epoch + 1u;
  const unsigned target = epoch * (unsigned)G;
  const bool is_feeder = warp == MK_CW - 1;
  const uint64_t policy = l2_policy((a.flags & 1) != 0);

  // the program's stage records -> shared memory once; op records of stage s + 1 are fetched (cp.async) while stage s runs
  MkStageLite* stg = reinterpret_cast<MkStageLite*>(smem + SM_STG);
  for (int i = tid; i < a.n_stages; i += MK_CT) stg[i] = a.stages_lite[i];
  auto fetch_ops = [&](int s) {   // threads 0 .. 16 n_ops: one 16-byte piece of stage s's op records each
    if (s < a.n_stages) {
      const MkStageLite st = stg[s];
      const int pieces = (st.n_a + st.n_b) * 16;
      if (tid < pieces)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(reinterpret_cast<uint8_t*>(ops_s + (s & 1) * MK_MAX_STAGE_OPS) + 16 * tid)),
                     "l"(reinterpret_cast<const uint8_t*>(a.ops + st.op_begin) + 16 * tid)
                     : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  __syncthreads();
  fetch_ops(0);

  unsigned seq = 0;          // units this CTA has completed
  unsigned slot = 0, ph = 0;   // ring slot / mbarrier parity of unit `seq`
  for (int s = 0; s < a.n_stages; ++s) {
    const MkStageLite st = stg[s];
    const int n_ops = st.n_a + st.n_b;
    const MkOp* ops_cur = ops_s + (s & 1) * MK_MAX_STAGE_OPS;
    long long* dbg = a.dbg ? a.dbg + ((long long)cta * a.n_stages + s) * 12 : nullptr;
    // weights are constants: the ring is topped up (two to three stages ahead) while the barrier fills
    if (is_feeder) {
      if (dbg && lane == 0) dbg[9] = clock64();
      feed(feeder, seq, Wslots, full_bar, lane, policy);
      if (dbg && lane == 0) dbg[10] = clock64();
    }
    if (tid == 0) {
      if (dbg) dbg[0] = clock64();
      if (s > 0) wait_counter(a.bar + s - 1, target, a.err);
      if (st.bg_wait >= 0) wait_counter(a.bg + st.bg_wait, target, a.err);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");   // this stage's op records (requested a stage ago)
    cons_sync();
    if (dbg && tid == 0) dbg[1] = clock64();
    if (dbg && tid == 32) {   // diagnostics: latency of two dependent L2 loads of freshly written activations, nothing else in flight yet
      const MkOp& o0 = ops_cur[0];
      if (o0.type == MK_T_LIN && o0.x_ext == MK_EXT_NONE && o0.X) {
        long long t0, t1 = 0, t2 = 0;   // the clock reads are control-dependent on the loaded values: they cannot be issued early
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0)::"memory");
        const float v = __ldcg(o0.X + (cta & 15) * 4);
        if (__float_as_int(v) != 0x7fbadbad) asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1)::"memory");
        const float w = __ldcg(o0.X + (o0.M > 8 ? 8 * o0.ldx : 0) + 64 + (__float_as_int(v) & 4));
        if (__float_as_int(w) != 0x7fbadbad) asm volatile("mov.u64 %0, %%clock64;" : "=l"(t2)::"memory");
        dbg[3] = (t1 - t0) | ((t2 - t1) << 32);
      }
    }
    fetch_ops(s + 1);   // into the other buffer: its last readers left it before this barrier
    int staged = -1;
    for (int oi = 0; oi <= n_ops; ++oi) {
      if (oi == st.n_a) {   // the stage's foreground ops are done in this CTA
        cons_sync();
        if (tid == 0) {
          if (dbg) dbg[2] = clock64();
          arrive_counter(a.bar + s, (a.flags & 2) != 0);
        }
      }
      if (oi == n_ops) break;
      const MkOp& op = ops_cur[oi];
      if (op.type == MK_T_LIN) {
        for (int u = first_unit(op.unit_begin, cta, G); u < op.units; u += G) {
          const int rt = op.row_tiles == 1 ? 0 : u / op.n_slices, slice = u - rt * op.n_slices;
          const int n0 = slice * op.FU, feats = min(op.FU, op.nout - n0);
          const int key = oi * 64 + rt;
          const bool restage = staged != key;
          if (is_feeder) feed(feeder, seq, Wslots, full_bar, lane, policy);   // several units per stage: keep the ring moving
          mbar_wait(&full_bar[slot], ph);
          if (dbg && tid == 0) { dbg[4] = clock64(); dbg[8] = feeder->issue_clk[seq & 7]; }
          const float* Ws = reinterpret_cast<const float*>(Wslots + slot * MK_SLOT_BYTES);
          if (op.rclass == 16) lin_unit16(op, a, Xs, Ws, red_s, rt * 16, n0, feats, restage, tid, dbg);
          else lin_unit1(op, a, Xs, Ws, red_s, n0, feats, restage, tid);
          // every warp passed the barrier in front of the epilogue, i.e. finished reading the slot: unit `seq` is complete
          staged = (op.rclass == 1 || op.K <= MK_XROW) ? key : -1;
          ++seq;
          if (++slot == MK_SLOTS) { slot = 0; ph ^= 1u; }
        }
      } else {
        const float* X = resolve_x(op, a);
        for (int u = first_unit(op.unit_begin, cta, G); u < op.units; u += G) {
          const int m0 = u * 16, rows = min(16, op.M - m0), kq = op.K >> 2;
          for (int e = tid; e < rows * kq; e += MK_CT) {
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
      if (tid == 0) arrive_counter(a.bg + st.bg_arrive, (a.flags & 2) != 0);
    }
    // no barrier here: the next stage's records live in the other buffer, and its first unit synchronises before it
    // touches the shared staging areas
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

void mk_plan_op(MkOp& op, int ctas, int min_fu) {
  if (op.type != MK_T_LIN) {
    ECHO_CHECK(op.K % 4 == 0 && op.ldx % 4 == 0 && op.ldy % 4 == 0, "layout program: copy op needs 16-byte rows");
    op.rclass = 16; op.row_tiles = cdiv(op.M, 16); op.FU = 0; op.n_slices = 1; op.units = op.row_tiles;
    return;
  }
  ECHO_CHECK(op.M >= 1 && op.K >= 4 && op.K % 4 == 0 && op.nout >= 1 && op.ldx % 4 == 0, "layout program: bad linear op (M=%d K=%d nout=%d)", op.M,
             op.K, op.nout);
  const bool geglu = op.epi == MK_EPI_GEGLU;
  // one-row ops without a row-wise prologue (the time path; also a one-node graph's plain layers) are fp32 dot products
  const bool plain1 = op.M == 1 && (op.pro == MK_NONE || op.pro == MK_SILU || op.pro == MK_TEMB) && !op.X2 && !geglu && op.epi == MK_EPI_LIN &&
                      op.K + MK_PAD <= 16 * MK_XSTRIDE;
  op.rclass = plain1 ? 1 : 16;
  op.row_tiles = op.rclass == 16 ? cdiv(op.M, 16) : 1;
  const int per_feat = geglu ? 2 : 1;   // weight rows a feature brings into the slot
  const int naux = (op.pro == MK_GN || op.pro == MK_LN) ? 2 : 0;   // scale / shift rows travel in the slot
  int cap = (MK_SLOT_BYTES / ((op.K + MK_PAD) * 4) - naux) / per_feat;
  const int hard = (op.rclass == 1 ? 16 : MK_MAX_FU) / per_feat;
  cap = cap > hard ? hard : cap;
  ECHO_CHECK(cap >= 1, "layout program: K=%d too long for a weight slot", op.K);
  if (op.rclass == 16) {
    ECHO_CHECK(op.pro != MK_TEMB, "layout program: the timestep embedding is a one-row prologue");
    if (op.K > MK_XROW)
      ECHO_CHECK(op.pro == MK_NONE || op.pro == MK_SILU, "layout program: prologue %d needs the whole row staged (K=%d)", op.pro, op.K);
    ECHO_CHECK(op.bcast_rows == 0, "layout program: broadcast store needs a one-row op");
  }
  if (op.pro == MK_GN) ECHO_CHECK(op.K % 128 == 0 && op.cpg >= 4 && op.cpg <= 128 && (op.cpg & (op.cpg - 1)) == 0, "layout program: GroupNorm prologue K=%d cpg=%d", op.K, op.cpg);
  if (op.pro == MK_LN) ECHO_CHECK(op.K <= MK_XROW, "layout program: LayerNorm prologue K=%d", op.K);
  if (op.X2) ECHO_CHECK(op.K1 % 4 == 0 && op.K1 > 0 && op.K1 < op.K && op.ldx2 % 4 == 0 && op.pro != MK_LN && op.pro != MK_EDGE && op.pro != MK_POOL, "layout program: bad concat input");
  if (geglu) ECHO_CHECK(op.act == 0 && !op.res && !op.res2 && op.bias, "layout program: GEGLU epilogue takes bias only");
  int want = cdiv((int64_t)op.nout * op.row_tiles, (int64_t)ctas);
  want = want < min_fu ? min_fu : want;   // min_fu: the stage balancer asks for fatter units (layout.cu flush())
  op.FU = want > cap ? cap : want;
  op.n_slices = cdiv(op.nout, op.FU);
  op.units = op.n_slices * op.row_tiles;
}

// Every CTA's weight fetches in the order its consumer loop meets them (stages -> foreground then background ops -> the op's
// units dealt round-robin from unit_begin): off[c] .. off[c + 1] index CTA c's descriptors in `out`.
void mk_build_fetch(const std::vector<MkOp>& ops, const std::vector<MkStage>& stages, int ctas, std::vector<MkFetch>& out, std::vector<int>& off) {
  out.clear();
  off.assign(ctas + 1, 0);
  for (int c = 0; c < ctas; ++c) {
    off[c] = (int)out.size();
    for (const MkStage& st : stages)
      for (int oi = 0; oi < st.n_a + st.n_b; ++oi) {
        const MkOp& op = ops[st.op_begin + oi];
        if (op.type != MK_T_LIN) continue;
        int first = c - op.unit_begin;
        if (first < 0) first += ctas;
        for (int u = first; u < op.units; u += ctas) {
          const int slice = u % op.n_slices, n0 = slice * op.FU, feats = op.nout - n0 < op.FU ? op.nout - n0 : op.FU;
          MkFetch d;
          memset(&d, 0, sizeof(d));
          d.w = op.W + (size_t)n0 * op.K;
          d.K = op.K;
          d.rows0 = feats;
          if (op.epi == MK_EPI_GEGLU) { d.g = op.W + (size_t)(op.nout + n0) * op.K; d.rows1 = feats; }
          if (op.pro == MK_GN || op.pro == MK_LN) { d.aux = op.gamma; d.aux2 = op.beta; d.naux = 2; }
          ECHO_CHECK(d.rows0 + d.rows1 + d.naux <= 32 && (size_t)(d.rows0 + d.rows1 + d.naux) * (op.K + MK_PAD) * 4 <= (size_t)MK_SLOT_BYTES,
                     "layout program: unit of %d rows x K=%d exceeds a weight slot", d.rows0 + d.rows1 + d.naux, op.K);
          out.push_back(d);
        }
      }
  }
  off[ctas] = (int)out.size();
}

void mk_launch(const MkArgs& a, int ctas, cudaStream_t s) {
  MkArgs args = a;
  void* params[] = {(void*)&args};
  ECHO_CUDA(cudaLaunchCooperativeKernel((const void*)layout_mk_kernel, dim3(ctas), dim3(MK_THREADS), params, (size_t)SM_TOTAL, s));
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
