// Few-row linear layers: Y[M, nout] = epi( f(X)[M, K] @ W[nout, K]^T ), M = nodes or edges of one scene graph.
//
// These are the layout branch's ~180 contractions per step, the GraphTripleConv MLPs and every per-object vector op of
// the shape step (time MLP, ResBlock emb_layers, attn2 to_v/to_out).  With a few dozen rows they are bound by streaming
// W from HBM once, and — being tiny — by how many bytes are in flight: a launch must put >= ~1200 warps on the chip.
//
// Layout of one CTA: RN = 4 output features x MT rows, K split over the CTA's WK warps (1..8, chosen per launch so that
// small-nout layers still fill the machine).  Inside a warp the 32 lanes split the K slice in 16-byte pieces (512
// contiguous bytes of a weight row per warp-load, all RN x unroll loads issued before the FMAs), X is re-read through L1
// (it is tiny).  Reduction: a halving butterfly over the lanes (62 shuffles for 64 accumulators instead of 320), then a
// fixed-order sum over the WK warps through shared memory — deterministic, no atomics.
#include "ops.cuh"

namespace echo {
namespace {

constexpr int RN = 4;   // output features per CTA

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

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

// (code size matters here: the kernel runs ~1 iteration per warp, so it is instruction-fetch bound when over-unrolled;
//  the SiLU-on-load variant is a separate instantiation and the K loop is not unrolled)
template <class TW, int MT, bool IN_SILU>
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

  float acc[N];
#pragma unroll
  for (int i = 0; i < N; ++i) acc[i] = 0.f;

#pragma unroll 1
  for (int k = k_beg + lane * 4; k < k_end; k += 128) {
    float w[RN][4];
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      if (n0 + j < a.nout) ldw4<TW>(W + (int64_t)(n0 + j) * ldw + k, w[j]);
      else w[j][0] = w[j][1] = w[j][2] = w[j][3] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
      if (m0 + i < a.M) {
        float4 xv = *reinterpret_cast<const float4*>(a.X + (int64_t)(m0 + i) * a.ldx + k);
        if (IN_SILU) { xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w); }
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
        a.Y[(int64_t)m * a.ldy + n] = v;
      }
    }
  }
}

template <class TW>
void launch(const LinArgs& a, cudaStream_t s) {
  constexpr int MT = 8;
  const int row_tiles = cdiv(a.M, MT);
  const int base_warps = cdiv(a.nout, RN) * row_tiles;
  int wk = 1;
  while (wk < 8 && base_warps * wk < 1184 && a.K / (wk * 2) >= 128) wk *= 2;
  const int k_slice = cdiv(cdiv(a.K, wk), 128) * 128;
  dim3 grid(cdiv(a.nout, RN), row_tiles);
  const size_t smem = (size_t)wk * MT * RN * sizeof(float);
  if (a.in_act == 1) linear_rows_kernel<TW, MT, true><<<grid, 32 * wk, smem, s>>>(a, k_slice);
  else linear_rows_kernel<TW, MT, false><<<grid, 32 * wk, smem, s>>>(a, k_slice);
}

}  // namespace

void linear_rows(const LinArgs& a, cudaStream_t s) {
  ECHO_CHECK(a.X && a.W && a.Y, "linear_rows: null operand");
  ECHO_CHECK(a.ldw % 4 == 0, "linear_rows: ldw %% 4");
  ECHO_CHECK(a.K % 4 == 0 && a.ldx % 4 == 0 && ((uintptr_t)a.X % 16) == 0 && ((uintptr_t)a.W % 16) == 0,
             "linear_rows: K=%d ldx=%lld must be multiples of 4 and 16-byte aligned", a.K, (long long)a.ldx);
  if (a.M == 0) return;
  ECHO_CHECK(cdiv(a.M, 8) <= 65535, "linear_rows: too many rows");
  if (a.w_dt == F32) launch<float>(a, s);
  else launch<__nv_bfloat16>(a, s);
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
