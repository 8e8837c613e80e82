// Few-row linear layers: Y[M, nout] = epi( pro(X)[M, K] @ W[nout, K]^T ), M = nodes or edges of one scene graph.
//
// These are the layout branch's contractions, the GraphTripleConv MLPs and every per-object vector op of the shape step
// (time MLP, ResBlock emb_layers, attn2 to_v/to_out).  With a few dozen rows they are bound by streaming W once and —
// being tiny — by latency and launch count, so the kernel (a) puts >= ~1200 warps on the chip per launch (RN = 4 output
// features x MT = 8 rows per CTA, K split over the CTA's 1..8 warps) and (b) absorbs the elementwise op that precedes
// the Linear in the network as a PROLOGUE applied while X is loaded, so a ResBlock or transformer block of the layout
// denoiser is 2-6 launches instead of 6-12:
//   PRO_SILU   x -> silu(x)                                   (emb_layers: Linear(SiLU(emb)))
//   PRO_GN     GroupNorm(32 groups)(+SiLU): a group is 16 or 32 consecutive channels = 4 or 8 adjacent lanes of the warp,
//              so its statistics are a 2-3 step shuffle — no pass over X, no extra kernel
//   PRO_LN     LayerNorm: each warp first reduces its rows over the full K (X is tiny and L1-resident)
//   PRO_GEGLU  x = a * gelu_erf(g) with [a | g] the two halves of a 2K-wide row        (attention.py:39-46)
// X may be the channel concat [X1 | X2] of two tensors (skip connections), and the epilogue takes two residuals.
// Reduction: a halving butterfly over the lanes (31 shuffles for 32 accumulators), then a fixed-order sum over the
// warps through shared memory — deterministic, no atomics.
#include "ops.cuh"

namespace echo {
namespace {

constexpr int RN = 4;   // output features per CTA
constexpr int MT = 8;   // rows per CTA

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <class TW>
__device__ __forceinline__ void ldw4(const TW* p, float (&v)[4]);
template <>
__device__ __forceinline__ void ldw4<float>(const float* p, float (&v)[4]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ldw4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}

// Halving butterfly: on return lane l holds the full sums of elements [base, base + N/32) in v[0 .. N/32).
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

// element (m, k..k+3) of the (possibly concatenated) input
__device__ __forceinline__ float4 load_x(const LinArgs& a, int m, int k) {
  if (a.X2 && k >= a.K1) return *reinterpret_cast<const float4*>(a.X2 + (int64_t)m * a.ldx2 + (k - a.K1));
  return *reinterpret_cast<const float4*>(a.X + (int64_t)m * a.ldx + k);
}

template <class TW, int PRO>
__global__ void __launch_bounds__(256) linear_rows_kernel(const LinArgs a, int k_slice) {
  constexpr int N = MT * RN;
  extern __shared__ float part[];   // [WK][N]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, WK = blockDim.x >> 5;
  const int n0 = blockIdx.x * RN;
  const int m0 = blockIdx.y * MT;
  const TW* __restrict__ W = reinterpret_cast<const TW*>(a.W);
  const int64_t ldw = a.ldw ? a.ldw : a.K;
  const int k_beg = warp * k_slice;
  const int k_end = min(a.K, k_beg + k_slice);

  // Programmatic dependent launch: this grid is scheduled while its predecessor still runs.  The weights are constants,
  // so this warp's slice of W is pulled into L2 NOW -- the HBM latency of a few-row layer (its whole cost besides
  // launch) hides under the previous layer; only X has to wait for the predecessor to finish.
  griddep_launch();
  {
    constexpr int EPL = 128 / (int)sizeof(TW);   // elements per 128-byte line
#pragma unroll
    for (int j = 0; j < RN; ++j)
      if (n0 + j < a.nout)
        for (int k = k_beg + lane * EPL; k < k_end; k += 32 * EPL)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(W + (int64_t)(n0 + j) * ldw + k));
  }
  griddep_wait();

  // LayerNorm statistics of this CTA's rows over the whole K (two-pass, fp32)
  float ln_mean[MT], ln_rstd[MT];
  if (PRO == PRO_LN) {
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      ln_mean[i] = 0.f;
      ln_rstd[i] = 0.f;
      if (m0 + i < a.M) {
        float s = 0.f;
        for (int k = lane * 4; k < a.K; k += 128) { const float4 x = load_x(a, m0 + i, k); s += (x.x + x.y) + (x.z + x.w); }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const float mean = s / a.K;
        float ss = 0.f;
        for (int k = lane * 4; k < a.K; k += 128) {
          const float4 x = load_x(a, m0 + i, k);
          const float d0 = x.x - mean, d1 = x.y - mean, d2 = x.z - mean, d3 = x.w - mean;
          ss += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        ln_mean[i] = mean;
        ln_rstd[i] = rsqrtf(ss / a.K + a.eps);
      }
    }
  }

  float acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.f;

#pragma unroll 1
  for (int kb = k_beg; kb < k_end; kb += 128) {
    const int k = kb + lane * 4;
    const bool kin = k < k_end;
    float w[RN][4];
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      if (kin && n0 + j < a.nout) ldw4<TW>(W + (int64_t)(n0 + j) * ldw + k, w[j]);
      else w[j][0] = w[j][1] = w[j][2] = w[j][3] = 0.f;
    }
    float4 gm = make_float4(1.f, 1.f, 1.f, 1.f), bt = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((PRO == PRO_GN || PRO == PRO_LN) && kin) {
      gm = __ldg(reinterpret_cast<const float4*>(a.gamma + k));
      bt = __ldg(reinterpret_cast<const float4*>(a.beta + k));
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      const bool rin = m0 + i < a.M;   // warp-uniform
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (rin && kin) {
        xv = load_x(a, m0 + i, k);
        if (PRO == PRO_GEGLU) {
          const float4 g = *reinterpret_cast<const float4*>(a.X + (int64_t)(m0 + i) * a.ldx + a.K + k);
          xv.x *= gelu_erf(g.x); xv.y *= gelu_erf(g.y); xv.z *= gelu_erf(g.z); xv.w *= gelu_erf(g.w);
        }
      }
      if (PRO == PRO_SILU) { xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w); }
      if (PRO == PRO_GN) {
        if (rin) {   // group = cpg consecutive channels = cpg/4 adjacent lanes (K % 128 == 0: every lane is in range)
          const int gl = a.cpg >> 2;
          float s = (xv.x + xv.y) + (xv.z + xv.w);
          for (int o = 1; o < gl; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          const float mean = s / a.cpg;
          const float d0 = xv.x - mean, d1 = xv.y - mean, d2 = xv.z - mean, d3 = xv.w - mean;
          float ss = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
          for (int o = 1; o < gl; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
          const float rstd = rsqrtf(ss / a.cpg + a.eps);
          xv.x = d0 * rstd * gm.x + bt.x; xv.y = d1 * rstd * gm.y + bt.y;
          xv.z = d2 * rstd * gm.z + bt.z; xv.w = d3 * rstd * gm.w + bt.w;
          if (a.pro_act) { xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w); }
        }
      }
      if (PRO == PRO_LN) {
        if (rin && kin) {
          xv.x = (xv.x - ln_mean[i]) * ln_rstd[i] * gm.x + bt.x; xv.y = (xv.y - ln_mean[i]) * ln_rstd[i] * gm.y + bt.y;
          xv.z = (xv.z - ln_mean[i]) * ln_rstd[i] * gm.z + bt.z; xv.w = (xv.w - ln_mean[i]) * ln_rstd[i] * gm.w + bt.w;
        }
      }
      if (rin) {
#pragma unroll
        for (int j = 0; j < RN; ++j) {
          float s = acc[i * RN + j];
          s = fmaf(xv.x, w[j][0], s);
          s = fmaf(xv.y, w[j][1], s);
          s = fmaf(xv.z, w[j][2], s);
          s = fmaf(xv.w, w[j][3], s);
          acc[i * RN + j] = s;
        }
      }
    }
  }
  const int base = butterfly<N>(acc, lane);
#pragma unroll
  for (int e = 0; e < N / 32; ++e) part[warp * N + base + e] = acc[e];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int e = 0; e < N / 32; ++e) {
      const int idx = lane * (N / 32) + e;   // (i, j) = (idx / RN, idx % RN)
      float v = 0.f;
      for (int wk = 0; wk < WK; ++wk) v += part[wk * N + idx];
      const int m = m0 + idx / RN, n = n0 + idx % RN;
      if (m < a.M && n < a.nout) {
        if (a.bias) v += __ldg(a.bias + n);
        if (a.act == 1) v = fmaxf(v, 0.f);
        else if (a.act == 2) v = silu_f(v);
        if (a.res) v += a.res[(int64_t)m * a.ld_res + n];
        if (a.res2) v += a.res2[(int64_t)m * a.ld_res2 + n];
        a.Y[(int64_t)m * a.ldy + n] = v;
      }
    }
  }
}

// The prologue alone, one warp per row: out[m, :] = what linear_rows_kernel feeds its dot products for row m (same expressions in
// the same order), for the batched path where the contraction runs as a tiled GEMM over the materialised rows.
template <int PRO>
__global__ void __launch_bounds__(256) linear_prologue_kernel(const LinArgs a, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, m = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (m >= a.M) return;
  float ln_mean = 0.f, ln_rstd = 0.f;
  if (PRO == PRO_LN) {
    float s = 0.f;
    for (int k = lane * 4; k < a.K; k += 128) { const float4 x = load_x(a, m, k); s += (x.x + x.y) + (x.z + x.w); }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    ln_mean = s / a.K;
    float ss = 0.f;
    for (int k = lane * 4; k < a.K; k += 128) {
      const float4 x = load_x(a, m, k);
      const float d0 = x.x - ln_mean, d1 = x.y - ln_mean, d2 = x.z - ln_mean, d3 = x.w - ln_mean;
      ss += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    ln_rstd = rsqrtf(ss / a.K + a.eps);
  }
  for (int kb = 0; kb < a.K; kb += 128) {
    const int k = kb + lane * 4;
    const bool kin = k < a.K;
    float4 gm = make_float4(1.f, 1.f, 1.f, 1.f), bt = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((PRO == PRO_GN || PRO == PRO_LN) && kin) {
      gm = __ldg(reinterpret_cast<const float4*>(a.gamma + k));
      bt = __ldg(reinterpret_cast<const float4*>(a.beta + k));
    }
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kin) {
      xv = load_x(a, m, k);
      if (PRO == PRO_GEGLU) {
        const float4 g = *reinterpret_cast<const float4*>(a.X + (int64_t)m * a.ldx + a.K + k);
        xv.x *= gelu_erf(g.x); xv.y *= gelu_erf(g.y); xv.z *= gelu_erf(g.z); xv.w *= gelu_erf(g.w);
      }
    }
    if (PRO == PRO_SILU) { xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w); }
    if (PRO == PRO_GN) {   // K % 128 == 0: every lane is in range
      const int gl = a.cpg >> 2;
      float s = (xv.x + xv.y) + (xv.z + xv.w);
      for (int o = 1; o < gl; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s / a.cpg;
      const float d0 = xv.x - mean, d1 = xv.y - mean, d2 = xv.z - mean, d3 = xv.w - mean;
      float ss = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      for (int o = 1; o < gl; o <<= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
      const float rstd = rsqrtf(ss / a.cpg + a.eps);
      xv.x = d0 * rstd * gm.x + bt.x; xv.y = d1 * rstd * gm.y + bt.y;
      xv.z = d2 * rstd * gm.z + bt.z; xv.w = d3 * rstd * gm.w + bt.w;
      if (a.pro_act) { xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w); }
    }
    if (PRO == PRO_LN && kin) {
      xv.x = (xv.x - ln_mean) * ln_rstd * gm.x + bt.x; xv.y = (xv.y - ln_mean) * ln_rstd * gm.y + bt.y;
      xv.z = (xv.z - ln_mean) * ln_rstd * gm.z + bt.z; xv.w = (xv.w - ln_mean) * ln_rstd * gm.w + bt.w;
    }
    if (kin) *reinterpret_cast<float4*>(out + (int64_t)m * a.K + k) = xv;
  }
}

template <class TW, int PRO>
void launch_pro(const LinArgs& a, dim3 grid, int wk, int k_slice, size_t smem, cudaStream_t s) {
  launch_pdl(linear_rows_kernel<TW, PRO>, grid, dim3(32 * wk), smem, s, a, k_slice);
}

template <class TW>
void launch(const LinArgs& a, cudaStream_t s) {
  const int row_tiles = cdiv(a.M, MT);
  const int base_warps = cdiv(a.nout, RN) * row_tiles;
  int wk = 1;
  while (wk < 8 && base_warps * wk < 1184 && a.K / (wk * 2) >= 128) wk *= 2;
  const int k_slice = cdiv(cdiv(a.K, wk), 128) * 128;
  dim3 grid(cdiv(a.nout, RN), row_tiles);
  const size_t smem = (size_t)wk * MT * RN * sizeof(float);
  switch (a.pro) {
    case PRO_NONE: launch_pro<TW, PRO_NONE>(a, grid, wk, k_slice, smem, s); break;
    case PRO_SILU: launch_pro<TW, PRO_SILU>(a, grid, wk, k_slice, smem, s); break;
    case PRO_GN: launch_pro<TW, PRO_GN>(a, grid, wk, k_slice, smem, s); break;
    case PRO_LN: launch_pro<TW, PRO_LN>(a, grid, wk, k_slice, smem, s); break;
    case PRO_GEGLU: launch_pro<TW, PRO_GEGLU>(a, grid, wk, k_slice, smem, s); break;
    default: fail(ECHO_ERR_INVALID, "linear_rows: unknown prologue %d", a.pro);
  }
}

}  // namespace

bool linear_rows_gn_supported(int K, int cpg) { return K % 128 == 0 && (cpg == 4 || cpg == 8 || cpg == 16 || cpg == 32 || cpg == 64 || cpg == 128); }

static void check_lin_args(const LinArgs& a) {
  ECHO_CHECK(a.X && a.W && a.Y, "linear_rows: null operand");
  ECHO_CHECK(a.ldw % 4 == 0, "linear_rows: ldw %% 4");
  ECHO_CHECK(a.K % 4 == 0 && a.ldx % 4 == 0 && ((uintptr_t)a.X % 16) == 0 && ((uintptr_t)a.W % 16) == 0,
             "linear_rows: K=%d ldx=%lld must be multiples of 4 and 16-byte aligned", a.K, (long long)a.ldx);
  if (a.X2) ECHO_CHECK(a.K1 % 4 == 0 && a.ldx2 % 4 == 0 && ((uintptr_t)a.X2 % 16) == 0 && a.K1 > 0 && a.K1 < a.K && a.pro != PRO_GEGLU,
                       "linear_rows: bad concat input");
  if (a.pro == PRO_GN) ECHO_CHECK(a.gamma && a.beta && linear_rows_gn_supported(a.K, a.cpg), "linear_rows: GroupNorm prologue needs K %% 128 == 0 and a power-of-two group of >= 4 channels (K=%d cpg=%d)", a.K, a.cpg);
  if (a.pro == PRO_LN) ECHO_CHECK(a.gamma && a.beta, "linear_rows: LayerNorm prologue needs gamma/beta");
}

void linear_prologue(const LinArgs& a_in, float* out, cudaStream_t s) {
  LinArgs a = a_in;
  if (a.in_act == 1 && a.pro == PRO_NONE) a.pro = PRO_SILU;
  ECHO_CHECK(out && ((uintptr_t)out % 16) == 0, "linear_prologue: null / unaligned output");
  check_lin_args(a);
  if (a.M == 0) return;
  const dim3 grid(cdiv(a.M, 8));
  switch (a.pro) {
    case PRO_NONE: linear_prologue_kernel<PRO_NONE><<<grid, 256, 0, s>>>(a, out); break;   // (a concat of X and X2)
    case PRO_SILU: linear_prologue_kernel<PRO_SILU><<<grid, 256, 0, s>>>(a, out); break;
    case PRO_GN: linear_prologue_kernel<PRO_GN><<<grid, 256, 0, s>>>(a, out); break;
    case PRO_LN: linear_prologue_kernel<PRO_LN><<<grid, 256, 0, s>>>(a, out); break;
    case PRO_GEGLU: linear_prologue_kernel<PRO_GEGLU><<<grid, 256, 0, s>>>(a, out); break;
    default: fail(ECHO_ERR_INVALID, "linear_prologue: unknown prologue %d", a.pro);
  }
  ECHO_LAUNCH_CHECK();
}

void linear_rows(const LinArgs& a_in, cudaStream_t s) {
  if (dbg_skip("linear_rows")) return;
  LinArgs a = a_in;
  if (a.in_act == 1 && a.pro == PRO_NONE) a.pro = PRO_SILU;   // legacy spelling
  check_lin_args(a);
  if (a.M == 0) return;
  ECHO_CHECK(cdiv(a.M, 8) <= 65535, "linear_rows: too many rows");
  if (a.w_dt == F32) launch<float>(a, s);
  else launch<__nv_bfloat16>(a, s);
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
