// Few-row linear layers: Y[M, nout] = epi( f(X)[M, K] @ W[nout, K]^T ), M = nodes or edges of one scene graph.
//
// These are the layout branch's ~100 contractions per step, the GraphTripleConv MLPs and every per-object vector
// op of the shape step (time MLP, ResBlock emb_layers, attn2 to_v/to_out).  With M <= a few dozen rows they are
// bound by streaming W from HBM once: each warp owns RN output features, the 32 lanes split K in 16-byte pieces
// (512 contiguous bytes of a weight row per warp-load), X is re-read through L1 (it is tiny), and the cross-lane
// reduction is a fixed-order shuffle tree, so results are deterministic.
#include "ops.cuh"

namespace echo {
namespace {

constexpr int MT = 8;   // rows per pass (register tile)
constexpr int RN = 4;   // output features per warp

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

template <class TW>
__global__ void __launch_bounds__(256) linear_rows_kernel(const LinArgs a) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int n0 = warp * RN;
  if (n0 >= a.nout) return;
  const int m0 = blockIdx.y * MT;
  const TW* __restrict__ W = reinterpret_cast<const TW*>(a.W);
  const int64_t ldw = a.ldw ? a.ldw : a.K;

  float acc[MT][RN];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

  for (int k = lane * 4; k < a.K; k += 128) {
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
        if (a.in_act == 1) { xv.x = silu_f(xv.x); xv.y = silu_f(xv.y); xv.z = silu_f(xv.z); xv.w = silu_f(xv.w); }
#pragma unroll
        for (int j = 0; j < RN; ++j) {
          acc[i][j] = fmaf(xv.x, w[j][0], acc[i][j]);
          acc[i][j] = fmaf(xv.y, w[j][1], acc[i][j]);
          acc[i][j] = fmaf(xv.z, w[j][2], acc[i][j]);
          acc[i][j] = fmaf(xv.w, w[j][3], acc[i][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      float v = acc[i][j];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      acc[i][j] = v;
    }
  // lane l writes output (i, j) = (l / RN, l % RN)
  const int i = lane / RN, j = lane % RN;
  if (i < MT) {
    float v = 0.f;
#pragma unroll
    for (int ii = 0; ii < MT; ++ii)
#pragma unroll
      for (int jj = 0; jj < RN; ++jj)
        if (ii == i && jj == j) v = acc[ii][jj];
    const int m = m0 + i, n = n0 + j;
    if (m < a.M && n < a.nout) {
      if (a.bias) v += __ldg(a.bias + n);
      if (a.act == 1) v = fmaxf(v, 0.f);
      else if (a.act == 2) v = silu_f(v);
      if (a.res) v += a.res[(int64_t)m * a.ld_res + n];
      a.Y[(int64_t)m * a.ldy + n] = v;
    }
  }
}

}  // namespace

void linear_rows(const LinArgs& a, cudaStream_t s) {
  ECHO_CHECK(a.X && a.W && a.Y, "linear_rows: null operand");
  ECHO_CHECK(a.ldw % 4 == 0, "linear_rows: ldw %% 4");
  ECHO_CHECK(a.K % 4 == 0 && a.ldx % 4 == 0 && ((uintptr_t)a.X % 16) == 0 && ((uintptr_t)a.W % 16) == 0,
             "linear_rows: K=%d ldx=%lld must be multiples of 4 and 16-byte aligned", a.K, (long long)a.ldx);
  if (a.M == 0) return;
  const int warps = cdiv(a.nout, RN);
  dim3 grid(cdiv((int64_t)warps * 32, 256), cdiv(a.M, MT));
  if (a.w_dt == F32) linear_rows_kernel<float><<<grid, 256, 0, s>>>(a);
  else linear_rows_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a);
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
