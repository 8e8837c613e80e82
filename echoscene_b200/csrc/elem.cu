// Normalisation, activation, layout and sampler kernels of the denoiser step.  All are HBM-bound streaming
// kernels: 16-byte vector accesses along the channel (innermost) dimension, deterministic reductions (no float
// atomics), grid sized from the element count.
#include "ops.cuh"
#include <math.h>

namespace echo {

namespace {

constexpr int kSMs = 148;

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <class T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
template <class T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
template <class T>
__device__ __forceinline__ float load1(const T* p);
template <>
__device__ __forceinline__ float load1<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float load1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <class T>
__device__ __forceinline__ void store1(T* p, float v);
template <>
__device__ __forceinline__ void store1<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

static inline int grid_for(int64_t work_items, int threads) {
  int64_t b = (work_items + threads - 1) / threads;
  int64_t cap = (int64_t)kSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm statistics.  Block (chunk, obj): thread (q, ry) owns channel quad q and rows ry, ry+RY, ... of the chunk.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GN_ROWS = 32;

template <class T>
__global__ void gn_partial_kernel(const T* __restrict__ x, int64_t V, int C, int groups, int nquad, int RY,
                                  float* __restrict__ partial, int nchunks) {
  extern __shared__ float sm[];   // [RY][C] sums, [RY][C] squares
  float* ssum = sm;
  float* ssq = sm + (size_t)RY * C;
  const int chunk = blockIdx.x, obj = blockIdx.y;
  const int q = threadIdx.x % nquad, ry = threadIdx.x / nquad;
  const int64_t r0 = (int64_t)chunk * GN_ROWS;
  const int64_t r1 = min(r0 + (int64_t)GN_ROWS, V);
  float s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  if (ry < RY) {
    const T* base = x + ((int64_t)obj * V) * C + q * 4;
    for (int64_t r = r0 + ry; r < r1; r += RY) {
      float v[4];
      load4<T>(base + r * C, v);
#pragma unroll
      for (int j = 0; j < 4; ++j) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { ssum[ry * C + q * 4 + j] = s[j]; ssq[ry * C + q * 4 + j] = ss[j]; }
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float a = 0.f, b2 = 0.f;
    for (int r = 0; r < RY; ++r)
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += ssum[r * C + c]; b2 += ssq[r * C + c]; }
    float* o = partial + (((int64_t)obj * nchunks + chunk) * groups + g) * 2;
    o[0] = a;
    o[1] = b2;
  }
}

// one warp per (object, group): lanes stride over the chunk partials, fixed-pattern shuffle tree in double
__global__ void gn_finalize_kernel(const float* __restrict__ partial, int n, int nchunks, int groups, double count,
                                   float eps, float* __restrict__ stats) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n * groups) return;
  const int obj = i / groups, g = i % groups;
  double a = 0.0, b = 0.0;
  for (int c = lane; c < nchunks; c += 32) {
    const float2 p = *reinterpret_cast<const float2*>(partial + (((int64_t)obj * nchunks + c) * groups + g) * 2);
    a += (double)p.x;
    b += (double)p.y;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    const double mean = a / count;
    double var = b / count - mean * mean;
    if (var < 0.0) var = 0.0;
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

// Block (row chunk, obj): thread (q, ry) owns channel quad q with its affine terms folded into (scale, shift) once,
// and streams rows ry, ry+RY, ... of the chunk with 16-byte accesses.
constexpr int GN_APPLY_ROWS = 64;
template <class TI, class TO, bool FAST>
__global__ void gn_apply_kernel(const TI* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int64_t V, int C, int groups, int silu, int nquad, int RY,
                                TO* __restrict__ y) {
  const int q = threadIdx.x % nquad, ry = threadIdx.x / nquad;
  if (ry >= RY) return;
  const int obj = blockIdx.y, cpg = C / groups, c0 = q * 4;
  float sc[4], sh[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float* st = stats + ((int64_t)obj * groups + (c0 + j) / cpg) * 2;
    sc[j] = st[1] * __ldg(gamma + c0 + j);
    sh[j] = __ldg(beta + c0 + j) - st[0] * sc[j];
  }
  const int64_t r0 = (int64_t)blockIdx.x * GN_APPLY_ROWS;
  const int64_t r1 = min(r0 + (int64_t)GN_APPLY_ROWS, V);
  const TI* xb = x + ((int64_t)obj * V) * C + c0;
  TO* yb = y + ((int64_t)obj * V) * C + c0;
  for (int64_t r = r0 + ry; r < r1; r += RY) {
    float v[4];
    load4<TI>(xb + r * C, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float t = fmaf(v[j], sc[j], sh[j]);
      v[j] = silu ? (FAST ? __fdividef(t, 1.f + __expf(-t)) : silu_f(t)) : t;
    }
    store4<TO>(yb + r * C, v);
  }
}

// bf16 -> bf16 fast path: 8 channels (16 bytes) per thread, four rows in flight per thread
__global__ void __launch_bounds__(256) gn_apply_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ stats,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              int64_t V, int C, int groups, int silu, int noct, int RY, int rows_per_block,
                                                              __nv_bfloat16* __restrict__ y) {
  const int q = threadIdx.x % noct, ry = threadIdx.x / noct;
  if (ry >= RY) return;
  const int obj = blockIdx.y, cpg = C / groups, c0 = q * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float* st = stats + ((int64_t)obj * groups + (c0 + j) / cpg) * 2;
    sc[j] = st[1] * __ldg(gamma + c0 + j);
    sh[j] = __ldg(beta + c0 + j) - st[0] * sc[j];
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(r0 + (int64_t)rows_per_block, V);
  const __nv_bfloat16* xb = x + ((int64_t)obj * V) * C + c0;
  __nv_bfloat16* yb = y + ((int64_t)obj * V) * C + c0;
  for (int64_t r = r0 + ry; r < r1; r += 4 * RY) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (r + k * RY < r1) u[k] = __ldg(reinterpret_cast<const uint4*>(xb + (r + k * RY) * C));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (r + k * RY >= r1) break;
      uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
        float a = fmaf(__low2float(h2), sc[2 * e], sh[2 * e]);
        float b = fmaf(__high2float(h2), sc[2 * e + 1], sh[2 * e + 1]);
        if (silu) {
          a = __fdividef(a, 1.f + __expf(-a));
          b = __fdividef(b, 1.f + __expf(-b));
        }
        const __nv_bfloat162 o2 = __floats2bfloat162_rn(a, b);
        w[e] = *reinterpret_cast<const uint32_t*>(&o2);
      }
      *reinterpret_cast<uint4*>(yb + (r + k * RY) * C) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// rows variant: one warp per (row, group)
__global__ void gn_rows_kernel(const float* __restrict__ x, int M, int C, int groups, const float* __restrict__ gamma,
                               const float* __restrict__ beta, float eps, int silu, float* __restrict__ y) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M * groups) return;
  const int row = warp / groups, g = warp % groups, cpg = C / groups;
  const float* p = x + (int64_t)row * C + g * cpg;
  float s = 0.f;
  for (int c = lane; c < cpg; c += 32) s += p[c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / cpg;
  float ss = 0.f;
  for (int c = lane; c < cpg; c += 32) { float d = p[c] - mean; ss = fmaf(d, d, ss); }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / cpg + eps);
  for (int c = lane; c < cpg; c += 32) {
    const int ch = g * cpg + c;
    float t = (p[c] - mean) * rstd * gamma[ch] + beta[ch];
    y[(int64_t)row * C + ch] = silu ? silu_f(t) : t;
  }
}

// LayerNorm: one warp per row, two-pass statistics in fp32
template <class TI, class TO>
__global__ void layer_norm_kernel(const TI* __restrict__ x, int64_t rows, int C, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float eps, TO* __restrict__ y) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const TI* p = x + row * C;
  const int nq = C / 4;
  float s = 0.f;
  for (int q = lane; q < nq; q += 32) { float v[4]; load4<TI>(p + q * 4, v); s += (v[0] + v[1]) + (v[2] + v[3]); }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float ss = 0.f;
  for (int q = lane; q < nq; q += 32) {
    float v[4]; load4<TI>(p + q * 4, v);
#pragma unroll
    for (int j = 0; j < 4; ++j) { float d = v[j] - mean; ss = fmaf(d, d, ss); }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / C + eps);
  for (int q = lane; q < nq; q += 32) {
    float v[4]; load4<TI>(p + q * 4, v);
    const float4 gm = *reinterpret_cast<const float4*>(gamma + q * 4), bt = *reinterpret_cast<const float4*>(beta + q * 4);
    v[0] = (v[0] - mean) * rstd * gm.x + bt.x;
    v[1] = (v[1] - mean) * rstd * gm.y + bt.y;
    v[2] = (v[2] - mean) * rstd * gm.z + bt.z;
    v[3] = (v[3] - mean) * rstd * gm.w + bt.w;
    store4<TO>(y + row * C + q * 4, v);
  }
}

// bf16 -> bf16 LayerNorm with the row held in registers (C <= 768): one global read, 16-byte accesses
__global__ void __launch_bounds__(256) layer_norm_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t rows, int C,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                              __nv_bfloat16* __restrict__ y) {
  griddep_launch();
  griddep_wait();
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nch = C >> 3;
  const __nv_bfloat16* p = x + row * C;
  float v[3][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int ch = lane + 32 * k;
    if (ch < nch) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + ch);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
        v[k][2 * e] = __low2float(h2);
        v[k][2 * e + 1] = __high2float(h2);
        s += v[k][2 * e] + v[k][2 * e + 1];
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k)
    if (lane + 32 * k < nch) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { const float d = v[k][e] - mean; ss = fmaf(d, d, ss); }
    }
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / C + eps);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int ch = lane + 32 * k;
    if (ch < nch) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + ch * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + ch * 8 + 4));
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      uint32_t w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 o2 = __floats2bfloat162_rn((v[k][2 * e] - mean) * rstd * gm[2 * e] + bt[2 * e],
                                                        (v[k][2 * e + 1] - mean) * rstd * gm[2 * e + 1] + bt[2 * e + 1]);
        w[e] = *reinterpret_cast<const uint32_t*>(&o2);
      }
      *(reinterpret_cast<uint4*>(y + row * C) + ch) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

template <class TI, class TO>
__global__ void geglu_kernel(const TI* __restrict__ x, int64_t nquads_total, int F, TO* __restrict__ y) {
  const int nq = F / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquads_total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nq;
    const int f0 = (int)(i - row * nq) * 4;
    float a[4], g[4];
    load4<TI>(x + row * 2 * F + f0, a);
    load4<TI>(x + row * 2 * F + F + f0, g);
#pragma unroll
    for (int j = 0; j < 4; ++j) a[j] *= gelu_erf(g[j]);
    store4<TO>(y + row * F + f0, a);
  }
}

template <class T>
__global__ void concat_kernel(const T* __restrict__ a, int Ca, const T* __restrict__ b, int Cb, int64_t nquads_total,
                              T* __restrict__ out) {
  const int C = Ca + Cb, nq = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquads_total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / nq;
    const int c0 = (int)(i - row * nq) * 4;
    float v[4];
    if (c0 < Ca) load4<T>(a + row * Ca + c0, v); else load4<T>(b + row * Cb + (c0 - Ca), v);
    store4<T>(out + row * C + c0, v);
  }
}

template <class T>
__global__ void upsample_hw2_kernel(const T* __restrict__ x, int n, int d, int h, int w, int C, int64_t nquads_total,
                                    T* __restrict__ out) {
  const int nq = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nquads_total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / nq;
    const int c0 = (int)(i - row * nq) * 4;
    const int ow = (int)(row % (2 * w)); int64_t r = row / (2 * w);
    const int oh = (int)(r % (2 * h)); r /= (2 * h);     // r = obj*d + dd
    const int64_t src = (r * h + oh / 2) * w + ow / 2;
    float v[4];
    load4<T>(x + src * C + c0, v);
    store4<T>(out + row * C + c0, v);
  }
}

template <class T>
__global__ void maxpool3d_kernel(const T* __restrict__ x, int n, int d, int h, int w, int C, int k, int stride, int od,
                                 int oh, int ow, T* __restrict__ out) {
  const int64_t total = (int64_t)n * od * oh * ow * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C); int64_t r = i / C;
    const int x_ = (int)(r % ow); r /= ow;
    const int y_ = (int)(r % oh); r /= oh;
    const int z_ = (int)(r % od); const int obj = (int)(r / od);
    float m = -INFINITY;
    for (int a = 0; a < k; ++a)
      for (int b = 0; b < k; ++b)
        for (int e = 0; e < k; ++e) {
          const int zz = z_ * stride + a, yy = y_ * stride + b, xx = x_ * stride + e;
          m = fmaxf(m, load1<T>(x + ((((int64_t)obj * d + zz) * h + yy) * w + xx) * C + c));
        }
    store1<T>(out + i, m);
  }
}

template <class TO>
__global__ void ncdhw_to_cl_kernel(const float* __restrict__ x, int n, int c, int64_t V, TO* __restrict__ out) {
  const int64_t total = (int64_t)n * c * V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ch = (int)(i % c); const int64_t r = i / c;   // r = obj*V + v
    const int64_t obj = r / V, v = r - obj * V;
    store1<TO>(out + i, x[(obj * c + ch) * V + v]);
  }
}

template <class TI>
__global__ void cl_to_ncdhw_kernel(const TI* __restrict__ x, int n, int c, int64_t V, int ld, float* __restrict__ out) {
  const int64_t total = (int64_t)n * c * V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i % V; const int64_t r = i / V;     // r = obj*c + ch
    const int64_t obj = r / c; const int ch = (int)(r - obj * c);
    out[i] = load1<TI>(x + (obj * V + v) * ld + ch);
  }
}

template <class TI, class TO>
__global__ void convert_kernel(const TI* __restrict__ x, TO* __restrict__ y, int64_t count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    store1<TO>(y + i, load1<TI>(x + i));
}

template <class T>
__global__ void add_rowvec_kernel(T* __restrict__ y, int64_t rows, int C, const float* __restrict__ v, int64_t ldv,
                                  int64_t rpo) {
  const int64_t total = rows * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / C; const int c = (int)(i - row * C);
    store1<T>(y + i, load1<T>(y + i) + v[(row / rpo) * ldv + c]);
  }
}

// one warp per row
__global__ void softmax_rows_kernel(float* __restrict__ S, int64_t rows, int cols) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float* p = S + row * cols;
  float m = -INFINITY;
  for (int c = lane; c < cols; c += 32) m = fmaxf(m, p[c]);
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float s = 0.f;
  for (int c = lane; c < cols; c += 32) { float e = expf(p[c] - m); p[c] = e; s += e; }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;
  for (int c = lane; c < cols; c += 32) p[c] *= inv;
}

__global__ void timestep_embedding_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs, int n, int dim,
                                          float* __restrict__ out) {
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, k = i % half;
  const float arg = __fmul_rn((float)t[r], freqs[k]);
  out[(int64_t)r * dim + k] = cosf(arg);
  out[(int64_t)r * dim + half + k] = sinf(arg);
  if ((dim & 1) && k == 0) out[(int64_t)r * dim + dim - 1] = 0.f;
}

__global__ void fill_i64_kernel(int64_t* p, int n, int64_t v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void embedding_rows_kernel(const float* __restrict__ table, int D, const int64_t* __restrict__ idx,
                                      int64_t idx_stride, int64_t idx_off, int64_t rows, float* __restrict__ out, int64_t ldo) {
  const int64_t total = rows * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D; const int c = (int)(i - r * D);
    out[r * ldo + c] = table[idx[r * idx_stride + idx_off] * D + c];
  }
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int64_t lds, int64_t rows, int D, float* __restrict__ out,
                                 int64_t ldo) {
  const int64_t total = rows * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / D; const int c = (int)(i - r * D);
    out[r * ldo + c] = src[r * lds + c];
  }
}

template <class T>
__global__ void flatten_ncdhw_kernel(const T* __restrict__ x, int n, int64_t V, int C, float* __restrict__ out) {
  const int64_t total = (int64_t)n * V * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = i % V; const int64_t r = i / V;
    const int64_t obj = r / C; const int c = (int)(r - obj * C);
    out[i] = load1<T>(x + (obj * V + v) * C + c);
  }
}

// p_mean_variance('eps', 'fixedsmall', clip_denoised=False) + p_sample_sg (diffusion_ddpm.py:220-264, 296-309);
// operation order and rounding follow the reference's fp32 tensor expressions (no FMA contraction).
__global__ void ddpm_update_kernel(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ noise,
                                   const float* __restrict__ tab, int T, int t, int64_t count, float* __restrict__ out) {
  const float a = tab[t], b = tab[T + t], c1 = tab[2 * T + t], c2 = tab[3 * T + t], lv = tab[4 * T + t];
  const float sig = (t == 0 ? 0.f : 1.f) * expf(0.5f * lv);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float x0 = __fsub_rn(__fmul_rn(a, x[i]), __fmul_rn(b, eps[i]));
    const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, x[i]));
    out[i] = __fadd_rn(mean, __fmul_rn(sig, noise[i]));
  }
}

// p_sample_ddim tail, sigma = 0 (samplers/ddim.py:252-261)
template <class TE>
__global__ void ddim_update_kernel(const float* __restrict__ x, const TE* __restrict__ e, int e_cl, int n, int c, int64_t V, int ld,
                                   const float* __restrict__ coef, const int* __restrict__ slot, float* __restrict__ out) {
  if (slot) coef += 4 * *slot;   // the step's DDIM index lives on the device (graph-replayed chains)
  const float sa = coef[0], s1m = coef[1], sap = coef[2], sdir = coef[3];
  const int64_t total = (int64_t)n * c * V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float ev;
    if (e_cl) {
      const int64_t v = i % V; const int64_t r = i / V;
      const int64_t obj = r / c; const int ch = (int)(r - obj * c);
      ev = load1<TE>(e + (obj * V + v) * ld + ch);
    } else {
      ev = load1<TE>(e + i);
    }
    const float px0 = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(s1m, ev)), sa);
    out[i] = __fadd_rn(__fmul_rn(sap, px0), __fmul_rn(sdir, ev));
  }
}

__global__ void repack_conv_kernel(const float* __restrict__ w, int cout, int cin, int taps, float* __restrict__ out) {
  const int64_t total = (int64_t)cout * cin * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin); int64_t r = i / cin;
    const int t = (int)(r % taps); const int64_t o = r / taps;
    out[i] = w[(o * cin + c) * taps + t];
  }
}

__global__ void center_tap_kernel(const float* __restrict__ w, int64_t total, int k, float* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = w[i * k + k / 2];
}

__global__ void fold_bn_kernel(const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ var,
                               float eps, int nout, int K, float* __restrict__ w_out, float* __restrict__ b_out) {
  const int64_t total = (int64_t)nout * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / K);
    const float sc = gamma[o] / sqrtf(var[o] + eps);
    w_out[i] = w[i] * sc;
    if (i % K == 0) b_out[o] = (b[o] - mean[o]) * sc + beta[o];
  }
}

// one warp per gathered row; 16-byte copies when the row length allows
__global__ void gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx, int64_t n_idx, int64_t n_rows, int64_t D,
                                   float* __restrict__ out) {
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= n_idx) return;
  float* o = out + r * D;
  const int64_t i = idx[r];
  if (i < 0 || i >= n_rows) {   // torch's obj_vecs[idx] raises here; a kernel cannot: the row comes back as NaNs, never as foreign memory
    for (int64_t c = lane; c < D; c += 32) o[c] = __int_as_float(0x7fc00000);
    return;
  }
  const float* s = src + i * D;
  if ((D & 3) == 0) {
    for (int64_t q = lane; q < D / 4; q += 32) reinterpret_cast<float4*>(o)[q] = __ldg(reinterpret_cast<const float4*>(s) + q);
  } else {
    for (int64_t c = lane; c < D; c += 32) o[c] = s[c];
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------------
size_t gn_partial_floats(const Act& x, int groups) {
  return (size_t)x.n * cdiv(x.voxels(), GN_ROWS) * groups * 2;
}

void gn_stats(const Act& x, int groups, float eps, float* stats, float* partial, cudaStream_t s) {
  ECHO_CHECK(x.c % groups == 0 && x.c % 4 == 0, "gn_stats: C=%d not divisible", x.c);
  const int64_t V = x.voxels();
  const int nquad = x.c / 4;
  ECHO_CHECK(nquad <= 1024, "gn_stats: C too large");
  int RY = 256 / nquad;
  if (RY < 1) RY = 1;
  if (RY > GN_ROWS) RY = GN_ROWS;
  const int threads = ((nquad * RY + 31) / 32) * 32;
  const int nchunks = cdiv(V, GN_ROWS);
  const size_t smem = (size_t)RY * x.c * 2 * sizeof(float);
  dim3 grid(nchunks, x.n);
  if (x.dt == F32)
    gn_partial_kernel<float><<<grid, threads, smem, s>>>((const float*)x.p, V, x.c, groups, nquad, RY, partial, nchunks);
  else
    gn_partial_kernel<__nv_bfloat16><<<grid, threads, smem, s>>>((const __nv_bfloat16*)x.p, V, x.c, groups, nquad, RY, partial, nchunks);
  ECHO_LAUNCH_CHECK();
  const int tot = x.n * groups;
  gn_finalize_kernel<<<cdiv((int64_t)tot * 32, 256), 256, 0, s>>>(partial, x.n, nchunks, groups, (double)V * (x.c / groups), eps, stats);
  ECHO_LAUNCH_CHECK();
}

void gn_apply(const Act& x, const float* stats, const float* gamma, const float* beta, int groups, bool silu, const Act& out,
              cudaStream_t s) {
  if (dbg_skip("gn_apply")) return;
  if (x.dt == BF16 && out.dt == BF16 && x.c % 8 == 0 && x.c / 8 <= 256) {
    const int noct = x.c / 8;
    const int RY = 256 / noct;
    const int threads = ((noct * RY + 31) / 32) * 32;
    const int64_t V = x.voxels();
    const int rows_per_block = 16 * RY;   // 4 passes of 4 rows in flight per thread
    dim3 grid(cdiv(V, rows_per_block), x.n);
    gn_apply_bf16x8_kernel<<<grid, threads, 0, s>>>((const __nv_bfloat16*)x.p, stats, gamma, beta, V, x.c, groups, silu ? 1 : 0, noct, RY,
                                                    rows_per_block, (__nv_bfloat16*)out.p);
    ECHO_LAUNCH_CHECK();
    return;
  }
  const int nquad = x.c / 4;
  int RY = 256 / nquad;
  if (RY < 1) RY = 1;
  if (RY > GN_APPLY_ROWS) RY = GN_APPLY_ROWS;
  const int threads = ((nquad * RY + 31) / 32) * 32;
  const int64_t V = x.voxels();
  dim3 grid(cdiv(V, GN_APPLY_ROWS), x.n);
#define GA(TI, TO, FAST) gn_apply_kernel<TI, TO, FAST><<<grid, threads, 0, s>>>((const TI*)x.p, stats, gamma, beta, V, x.c, groups, silu ? 1 : 0, nquad, RY, (TO*)out.p)
  if (x.dt == F32 && out.dt == F32) GA(float, float, false);
  else if (x.dt == F32) GA(float, __nv_bfloat16, true);
  else if (out.dt == F32) GA(__nv_bfloat16, float, false);
  else GA(__nv_bfloat16, __nv_bfloat16, true);
#undef GA
  ECHO_LAUNCH_CHECK();
}

void gn_rows(const float* x, int M, int C, int groups, const float* gamma, const float* beta, float eps, bool silu, float* y,
             cudaStream_t s) {
  const int warps = M * groups;
  gn_rows_kernel<<<cdiv((int64_t)warps * 32, 256), 256, 0, s>>>(x, M, C, groups, gamma, beta, eps, silu ? 1 : 0, y);
  ECHO_LAUNCH_CHECK();
}

void layer_norm(const void* x, DT xdt, int64_t rows, int C, const float* gamma, const float* beta, float eps, void* y, DT ydt,
                cudaStream_t s) {
  if (dbg_skip("layer_norm")) return;
  ECHO_CHECK(C % 4 == 0, "layer_norm: C %% 4");
  const int grid = cdiv(rows * 32, 256);
  if (xdt == BF16 && ydt == BF16 && C % 8 == 0 && C <= 768) {
    launch_pdl(layer_norm_bf16_kernel, dim3(grid), dim3(256), 0, s, (const __nv_bfloat16*)x, rows, C, gamma, beta, eps, (__nv_bfloat16*)y);
    ECHO_LAUNCH_CHECK();
    return;
  }
#define LN(TI, TO) layer_norm_kernel<TI, TO><<<grid, 256, 0, s>>>((const TI*)x, rows, C, gamma, beta, eps, (TO*)y)
  if (xdt == F32 && ydt == F32) LN(float, float);
  else if (xdt == F32) LN(float, __nv_bfloat16);
  else if (ydt == F32) LN(__nv_bfloat16, float);
  else LN(__nv_bfloat16, __nv_bfloat16);
#undef LN
  ECHO_LAUNCH_CHECK();
}

void geglu(const void* x, DT xdt, int64_t rows, int F, void* y, DT ydt, cudaStream_t s) {
  ECHO_CHECK(F % 4 == 0, "geglu: F %% 4");
  const int64_t nq = rows * (F / 4);
  const int grid = grid_for(nq, 256);
#define GG(TI, TO) geglu_kernel<TI, TO><<<grid, 256, 0, s>>>((const TI*)x, nq, F, (TO*)y)
  if (xdt == F32 && ydt == F32) GG(float, float);
  else if (xdt == F32) GG(float, __nv_bfloat16);
  else if (ydt == F32) GG(__nv_bfloat16, float);
  else GG(__nv_bfloat16, __nv_bfloat16);
#undef GG
  ECHO_LAUNCH_CHECK();
}

void concat_channels(const Act& a, const Act& b, const Act& out, cudaStream_t s) {
  if (dbg_skip("concat")) return;
  ECHO_CHECK(a.dt == b.dt && a.dt == out.dt && a.rows() == b.rows() && out.c == a.c + b.c && a.c % 4 == 0 && b.c % 4 == 0,
             "concat: mismatch");
  const int64_t nq = a.rows() * (out.c / 4);
  const int grid = grid_for(nq, 256);
  if (a.dt == F32) concat_kernel<float><<<grid, 256, 0, s>>>((const float*)a.p, a.c, (const float*)b.p, b.c, nq, (float*)out.p);
  else concat_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)a.p, a.c, (const __nv_bfloat16*)b.p, b.c, nq, (__nv_bfloat16*)out.p);
  ECHO_LAUNCH_CHECK();
}

void upsample_hw2(const Act& x, const Act& out, cudaStream_t s) {
  if (dbg_skip("upsample")) return;
  ECHO_CHECK(out.h == 2 * x.h && out.w == 2 * x.w && out.d == x.d && out.c == x.c && x.dt == out.dt && x.c % 4 == 0, "upsample: mismatch");
  const int64_t nq = out.rows() * (x.c / 4);
  const int grid = grid_for(nq, 256);
  if (x.dt == F32) upsample_hw2_kernel<float><<<grid, 256, 0, s>>>((const float*)x.p, x.n, x.d, x.h, x.w, x.c, nq, (float*)out.p);
  else upsample_hw2_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x.p, x.n, x.d, x.h, x.w, x.c, nq, (__nv_bfloat16*)out.p);
  ECHO_LAUNCH_CHECK();
}

void maxpool3d(const Act& x, int k, int stride, const Act& out, cudaStream_t s) {
  ECHO_CHECK(x.dt == out.dt && out.c == x.c, "maxpool: mismatch");
  const int64_t total = out.rows() * out.c;
  const int grid = grid_for(total, 256);
  if (x.dt == F32) maxpool3d_kernel<float><<<grid, 256, 0, s>>>((const float*)x.p, x.n, x.d, x.h, x.w, x.c, k, stride, out.d, out.h, out.w, (float*)out.p);
  else maxpool3d_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x.p, x.n, x.d, x.h, x.w, x.c, k, stride, out.d, out.h, out.w, (__nv_bfloat16*)out.p);
  ECHO_LAUNCH_CHECK();
}

void ncdhw_to_cl(const float* x, int n, int c, int64_t V, void* out, DT odt, cudaStream_t s) {
  const int grid = grid_for((int64_t)n * c * V, 256);
  if (odt == F32) ncdhw_to_cl_kernel<float><<<grid, 256, 0, s>>>(x, n, c, V, (float*)out);
  else ncdhw_to_cl_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x, n, c, V, (__nv_bfloat16*)out);
  ECHO_LAUNCH_CHECK();
}

namespace {
// NCDHW fp32 -> channels-last bf16 with the channel dimension zero-padded to cpad (one thread per voxel, 16-byte stores)
__global__ void ncdhw_to_cl_pad16_kernel(const float* __restrict__ x, int n, int c, int64_t V, __nv_bfloat16* __restrict__ out) {
  const int64_t rows = (int64_t)n * V;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const int64_t obj = r / V, v = r - obj * V;
    __align__(16) __nv_bfloat16 t[16];
#pragma unroll
    for (int ch = 0; ch < 16; ++ch) t[ch] = __float2bfloat16(ch < c ? x[(obj * c + ch) * V + v] : 0.f);
    uint4* o = reinterpret_cast<uint4*>(out + r * 16);
    o[0] = reinterpret_cast<const uint4*>(t)[0];
    o[1] = reinterpret_cast<const uint4*>(t)[1];
  }
}
}  // namespace

void ncdhw_to_cl_pad16(const float* x, int n, int c, int64_t V, __nv_bfloat16* out, cudaStream_t s) {
  ECHO_CHECK(c <= 16, "ncdhw_to_cl_pad16: c > 16");
  ncdhw_to_cl_pad16_kernel<<<grid_for((int64_t)n * V, 256), 256, 0, s>>>(x, n, c, V, out);
  ECHO_LAUNCH_CHECK();
}

void cl_to_ncdhw(const void* x, DT xdt, int n, int c, int64_t V, int ld, float* out, cudaStream_t s) {
  const int grid = grid_for((int64_t)n * c * V, 256);
  if (xdt == F32) cl_to_ncdhw_kernel<float><<<grid, 256, 0, s>>>((const float*)x, n, c, V, ld, out);
  else cl_to_ncdhw_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, n, c, V, ld, out);
  ECHO_LAUNCH_CHECK();
}

namespace {
__global__ void silu_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v / (1.f + expf(-v));
  }
}
}  // namespace

void silu_f32(const float* x, float* y, int64_t count, cudaStream_t s) {
  silu_f32_kernel<<<grid_for(count, 256), 256, 0, s>>>(x, y, count);
  ECHO_LAUNCH_CHECK();
}

void convert(const void* x, DT xdt, void* y, DT ydt, int64_t count, cudaStream_t s) {
  const int grid = grid_for(count, 256);
  if (xdt == F32 && ydt == BF16) convert_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>((const float*)x, (__nv_bfloat16*)y, count);
  else if (xdt == BF16 && ydt == F32) convert_kernel<__nv_bfloat16, float><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (float*)y, count);
  else if (xdt == F32) convert_kernel<float, float><<<grid, 256, 0, s>>>((const float*)x, (float*)y, count);
  else convert_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, count);
  ECHO_LAUNCH_CHECK();
}

namespace {
__global__ void split_bf16_kernel(const float4* __restrict__ x, int64_t nvec, uint2* __restrict__ hi, uint2* __restrict__ lo) {
  griddep_launch();
  griddep_wait();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    const __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
    const __nv_bfloat162 c = __floats2bfloat162_rn(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
    const __nv_bfloat162 d = __floats2bfloat162_rn(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
    hi[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    lo[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&c), *reinterpret_cast<const uint32_t*>(&d));
  }
}
}  // namespace

void split_bf16(const float* x, int64_t count, __nv_bfloat16* hi, __nv_bfloat16* lo, cudaStream_t s) {
  ECHO_CHECK(count % 4 == 0 && ((uintptr_t)x % 16) == 0 && ((uintptr_t)hi % 8) == 0 && ((uintptr_t)lo % 8) == 0, "split_bf16: count %% 4 and aligned operands");
  if (count == 0) return;
  int64_t blocks = (count / 4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  launch_pdl(split_bf16_kernel, dim3((int)blocks), dim3(256), 0, s, (const float4*)x, count / 4, (uint2*)hi, (uint2*)lo);
  ECHO_LAUNCH_CHECK();
}

void add_rowvec(void* y, DT ydt, int64_t rows, int C, const float* v, int64_t ldv, int64_t rpo, cudaStream_t s) {
  const int grid = grid_for(rows * C, 256);
  if (ydt == F32) add_rowvec_kernel<float><<<grid, 256, 0, s>>>((float*)y, rows, C, v, ldv, rpo);
  else add_rowvec_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)y, rows, C, v, ldv, rpo);
  ECHO_LAUNCH_CHECK();
}

void softmax_rows(float* S, int64_t rows, int cols, cudaStream_t s) {
  softmax_rows_kernel<<<cdiv(rows * 32, 256), 256, 0, s>>>(S, rows, cols);
  ECHO_LAUNCH_CHECK();
}

// the frequency table is passed in (see model.cu: it is computed once, the way torch computes it)
void timestep_embedding_tab(const int64_t* t, const float* freqs, int n, int dim, float* out, cudaStream_t s) {
  timestep_embedding_kernel<<<cdiv((int64_t)n * (dim / 2), 128), 128, 0, s>>>(t, freqs, n, dim, out);
  ECHO_LAUNCH_CHECK();
}

void fill_i64(int64_t* p, int n, int64_t v, cudaStream_t s) {
  fill_i64_kernel<<<cdiv(n, 128), 128, 0, s>>>(p, n, v);
  ECHO_LAUNCH_CHECK();
}

void embedding_rows(const float* table, int D, const int64_t* idx, int64_t idx_stride, int64_t idx_off, int64_t rows, float* out,
                    int64_t ldo, cudaStream_t s) {
  embedding_rows_kernel<<<grid_for(rows * D, 256), 256, 0, s>>>(table, D, idx, idx_stride, idx_off, rows, out, ldo);
  ECHO_LAUNCH_CHECK();
}

void copy_cols(const float* src, int64_t lds, int64_t rows, int D, float* out, int64_t ldo, cudaStream_t s) {
  copy_cols_kernel<<<grid_for(rows * D, 256), 256, 0, s>>>(src, lds, rows, D, out, ldo);
  ECHO_LAUNCH_CHECK();
}

void flatten_ncdhw(const Act& x, float* out, cudaStream_t s) {
  const int grid = grid_for(x.rows() * x.c, 256);
  if (x.dt == F32) flatten_ncdhw_kernel<float><<<grid, 256, 0, s>>>((const float*)x.p, x.n, x.voxels(), x.c, out);
  else flatten_ncdhw_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)x.p, x.n, x.voxels(), x.c, out);
  ECHO_LAUNCH_CHECK();
}

void ddpm_update(const float* x, const float* eps, const float* noise, const float* tab, int T, int t, int64_t count, float* out,
                 cudaStream_t s) {
  ddpm_update_kernel<<<grid_for(count, 128), 128, 0, s>>>(x, eps, noise, tab, T, t, count, out);
  ECHO_LAUNCH_CHECK();
}

void ddim_update(const float* x, const void* e, DT edt, bool e_cl, int n, int c, int64_t V, int ld, const float* coef4, float* out,
                 cudaStream_t s, const int* slot) {
  const int grid = grid_for((int64_t)n * c * V, 256);
  if (edt == F32) ddim_update_kernel<float><<<grid, 256, 0, s>>>(x, (const float*)e, e_cl ? 1 : 0, n, c, V, ld, coef4, slot, out);
  else ddim_update_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x, (const __nv_bfloat16*)e, e_cl ? 1 : 0, n, c, V, ld, coef4, slot, out);
  ECHO_LAUNCH_CHECK();
}

namespace {
__global__ void fill_i64_slot_kernel(int64_t* p, int n, const int32_t* __restrict__ table, const int* __restrict__ slot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int64_t)table[*slot];
}
__global__ void set_i32_kernel(int* p, int v) { *p = v; }
}  // namespace

void fill_i64_from_slot(int64_t* p, int n, const int32_t* table, const int* slot, cudaStream_t s) {
  fill_i64_slot_kernel<<<cdiv(n, 128), 128, 0, s>>>(p, n, table, slot);
  ECHO_LAUNCH_CHECK();
}

void set_i32(int* p, int v, cudaStream_t s) {
  set_i32_kernel<<<1, 1, 0, s>>>(p, v);
  ECHO_LAUNCH_CHECK();
}

void fold_upsample_weight(const float* w, int cout, int cin, float* out, bool up_depth) {
  // phase p in {0,1} along one axis: folded tap 0 reads low-res offset p-1, tap 1 reads offset p;
  // p = 0: {k=0}, {k=1,2};  p = 1: {k=0,1}, {k=2}
  auto member = [](int p, int a, int k) { return p == 0 ? (a == 0 ? k == 0 : k >= 1) : (a == 0 ? k <= 1 : k == 2); };
  if (!up_depth) {   // x(1,2,2): [cout][4 phases (py,px)][12 taps (kd,a,b)][cin]
    for (int n = 0; n < cout; ++n)
      for (int py = 0; py < 2; ++py)
        for (int px = 0; px < 2; ++px)
          for (int kd = 0; kd < 3; ++kd)
            for (int a = 0; a < 2; ++a)
              for (int b = 0; b < 2; ++b) {
                float* dst = out + (((size_t)n * 4 + py * 2 + px) * 12 + (kd * 2 + a) * 2 + b) * cin;
                for (int c = 0; c < cin; ++c) dst[c] = 0.f;
                for (int kh = 0; kh < 3; ++kh)
                  for (int kw = 0; kw < 3; ++kw) {
                    if (!member(py, a, kh) || !member(px, b, kw)) continue;
                    const float* src = w + ((size_t)n * 27 + kd * 9 + kh * 3 + kw) * cin;
                    for (int c = 0; c < cin; ++c) dst[c] += src[c];
                  }
              }
    return;
  }
  // x(2,2,2): [cout][8 phases (pz,py,px)][8 taps (a_d,a_h,a_w)][cin]
  for (int n = 0; n < cout; ++n)
    for (int ph = 0; ph < 8; ++ph)
      for (int t = 0; t < 8; ++t) {
        const int pz = (ph >> 2) & 1, py = (ph >> 1) & 1, px = ph & 1, ad = (t >> 2) & 1, ah = (t >> 1) & 1, aw = t & 1;
        float* dst = out + (((size_t)n * 8 + ph) * 8 + t) * cin;
        for (int c = 0; c < cin; ++c) dst[c] = 0.f;
        for (int kd = 0; kd < 3; ++kd)
          for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
              if (!member(pz, ad, kd) || !member(py, ah, kh) || !member(px, aw, kw)) continue;
              const float* src = w + ((size_t)n * 27 + kd * 9 + kh * 3 + kw) * cin;
              for (int c = 0; c < cin; ++c) dst[c] += src[c];
            }
      }
}

void repack_conv_weight(const float* w, int cout, int cin, int taps, float* out, cudaStream_t s) {
  repack_conv_kernel<<<grid_for((int64_t)cout * cin * taps, 256), 256, 0, s>>>(w, cout, cin, taps, out);
  ECHO_LAUNCH_CHECK();
}

void conv1d_center_tap(const float* w, int cout, int cin, int k, float* out, cudaStream_t s) {
  center_tap_kernel<<<grid_for((int64_t)cout * cin, 256), 256, 0, s>>>(w, (int64_t)cout * cin, k, out);
  ECHO_LAUNCH_CHECK();
}

void fold_bn(const float* w, const float* b, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
             int nout, int K, float* w_out, float* b_out, cudaStream_t s) {
  fold_bn_kernel<<<grid_for((int64_t)nout * K, 256), 256, 0, s>>>(w, b, gamma, beta, mean, var, eps, nout, K, w_out, b_out);
  ECHO_LAUNCH_CHECK();
}

void gather_rows(const float* src, const int64_t* idx, int64_t n_idx, int64_t n_rows, int64_t D, float* out, cudaStream_t s) {
  if (n_idx == 0) return;
  gather_rows_kernel<<<cdiv(n_idx * 32, 256), 256, 0, s>>>(src, idx, n_idx, n_rows, D, out);
  ECHO_LAUNCH_CHECK();
}

// ---- fp32 attention with materialised scores ---------------------------------------------------------------------
size_t attention_f32_ws_floats(int n, int tokens, int heads) { return (size_t)n * heads * tokens * tokens; }

void attention_f32(const float* qkv, int n, int tokens, int heads, int dh, float* ws, float* out, cudaStream_t s) {
  const int C = heads * dh;
  const int64_t t3c = (int64_t)tokens * 3 * C, tt = (int64_t)tokens * tokens;
  GemmArgs g;  // S = (Q K^T) * dh^-0.5   (attention.py:203)
  g.A = qkv; g.n = 1; g.w = tokens; g.ow = tokens; g.cin = dh; g.lda = 3 * C;
  g.W = qkv + C; g.w_stride_n = 3 * C; g.w_stride_k = 1; g.cout = tokens;
  g.out = ws; g.ldo = tokens; g.alpha = 1.0f / sqrtf((float)dh);
  g.nb0 = n; g.nb1 = heads;
  g.a_bs0 = t3c; g.a_bs1 = dh; g.w_bs0 = t3c; g.w_bs1 = dh; g.o_bs0 = heads * tt; g.o_bs1 = tt;
  gemm_simt(g, s);
  softmax_rows(ws, (int64_t)n * heads * tokens, tokens, s);   // attention.py:215
  GemmArgs p;  // O = P V   (attention.py:217)
  p.A = ws; p.n = 1; p.w = tokens; p.ow = tokens; p.cin = tokens; p.lda = tokens;
  p.W = qkv + 2 * C; p.w_stride_n = 1; p.w_stride_k = 3 * C; p.cout = dh;
  p.out = out; p.ldo = C;
  p.nb0 = n; p.nb1 = heads;
  p.a_bs0 = heads * tt; p.a_bs1 = tt; p.w_bs0 = t3c; p.w_bs1 = dh; p.o_bs0 = (int64_t)tokens * C; p.o_bs1 = dh;
  gemm_simt(p, s);
}

// ---- GroupNorm statistics from the producing GEMM's column partials (gemm_tc.cu epilogue) ---------------------------
namespace {
// GroupNorm(+SiLU) whose statistics come from the partials the producing tcgen05 GEMM left behind
// ([obj][rows_per_obj tiles][C/32 chunks][8 slots][2] = per-tile (sum, sumsq) of 7-channel blocks; every group of this
// network is a whole number of such blocks because the channel counts are multiples of 224): every block first folds the partials of ITS
// object into per-group (mean, rstd) in shared memory (fp64, fixed order), then streams its rows.  No statistics kernel
// and no statistics pass over the activation.  The input may be the channel concat [A | B] of two tensors (skip
// connections, openai_model_3d.py:857-858): both are read in place and the raw concat is written next to the
// normalised output for the ResBlock's 1x1 skip convolution, so the concat is never a pass of its own.
__global__ void __launch_bounds__(256) gn_apply_cs_kernel(const __nv_bfloat16* __restrict__ xa, int CA, const float* __restrict__ csa,
                                                          const __nv_bfloat16* __restrict__ xb, int CB, const float* __restrict__ csb,
                                                          int R, double count, float eps, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, int64_t V, int groups, int silu, int noct, int RY,
                                                          int rows_per_block, __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ ycat) {
  extern __shared__ double gn_sm[];           // [RS][C / 7][2] block sums, then [groups][2] floats
  griddep_launch();
  griddep_wait();
  const int C = CA + CB, obj = blockIdx.y, cpg = C / groups;
  // phase 1: totals of every 7-channel block over the object's R tile partials.  A partial row is [chunk][8 slots][2]:
  // slot s of 32-column chunk q holds the part of block 32q/7 + s that lies inside the chunk (gemm_tc.cu epilogue).
  // The R rows of a block are dealt to RS threads (all their loads in flight together: this prologue is a latency chain in
  // front of the streaming pass of EVERY block of the grid); partial sums meet in shared memory in a fixed order.
  const int nblocks = C / 7;
  const int RS = max(1, min(8, (int)blockDim.x / nblocks));
  double2* part = reinterpret_cast<double2*>(gn_sm);   // [RS][nblocks]
  float* gst = reinterpret_cast<float*>(gn_sm + 2 * nblocks * RS);
  for (int item = threadIdx.x; item < nblocks * RS; item += blockDim.x) {
    const int b = item % nblocks, rs = item / nblocks;
    const bool in_a = b * 7 < CA;
    const int lb = in_a ? b : b - CA / 7;                 // block index inside its producer
    const int nch = (in_a ? CA : CB) >> 5;
    const float* src = (in_a ? csa : csb) + (int64_t)obj * R * nch * 16;
    const int q1 = (lb * 7) >> 5, q2 = (lb * 7 + 6) >> 5;
    const int i1 = (q1 * 8 + (lb - (q1 * 32) / 7)) * 2, i2 = (q2 * 8 + (lb - (q2 * 32) / 7)) * 2;
    const bool two = q2 != q1;
    double a = 0.0, bsum = 0.0;
    for (int r = rs; r < R; r += 4 * RS) {
      float2 v[4], w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int rr = r + k * RS;
        const float* row = src + (int64_t)rr * nch * 16;
        v[k] = rr < R ? *reinterpret_cast<const float2*>(row + i1) : make_float2(0.f, 0.f);
        w[k] = (two && rr < R) ? *reinterpret_cast<const float2*>(row + i2) : make_float2(0.f, 0.f);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        a += (double)v[k].x;
        bsum += (double)v[k].y;
        a += (double)w[k].x;
        bsum += (double)w[k].y;
      }
    }
    part[rs * nblocks + b] = make_double2(a, bsum);
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int bpg = cpg / 7;
    double a = 0.0, b = 0.0;
    for (int k = threadIdx.x * bpg; k < (threadIdx.x + 1) * bpg; ++k)
      for (int rs = 0; rs < RS; ++rs) { a += part[rs * nblocks + k].x; b += part[rs * nblocks + k].y; }
    const double mean = a / count;
    double var = b / count - mean * mean;
    if (var < 0.0) var = 0.0;
    gst[2 * threadIdx.x] = (float)mean;
    gst[2 * threadIdx.x + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int q = threadIdx.x % noct, ry = threadIdx.x / noct;
  if (ry >= RY) return;
  const int c0 = q * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float* st = gst + ((c0 + j) / cpg) * 2;
    sc[j] = st[1] * __ldg(gamma + c0 + j);
    sh[j] = __ldg(beta + c0 + j) - st[0] * sc[j];
  }
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = min(r0 + (int64_t)rows_per_block, V);
  const bool in_a = c0 < CA;
  const int Cs = in_a ? CA : CB;
  const __nv_bfloat16* xs = (in_a ? xa + c0 : xb + (c0 - CA)) + ((int64_t)obj * V) * Cs;
  __nv_bfloat16* yb = y + ((int64_t)obj * V) * C + c0;
  __nv_bfloat16* cb = ycat ? ycat + ((int64_t)obj * V) * C + c0 : nullptr;
  // software pipeline: the four rows of the NEXT pass are requested before this pass is evaluated (a pass is one memory
  // round trip otherwise, and this kernel is then latency-, not bandwidth-bound)
  uint4 un[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (r0 + ry + k * RY < r1) un[k] = __ldg(reinterpret_cast<const uint4*>(xs + (r0 + ry + k * RY) * Cs));
  for (int64_t r = r0 + ry; r < r1; r += 4 * RY) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = un[k];
    const int64_t rn = r + 4 * RY;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (rn + k * RY < r1) un[k] = __ldg(reinterpret_cast<const uint4*>(xs + (rn + k * RY) * Cs));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (r + k * RY >= r1) break;
      if (cb) *reinterpret_cast<uint4*>(cb + (r + k * RY) * C) = u[k];
      uint32_t w[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w[e]);
        float a = fmaf(__low2float(h2), sc[2 * e], sh[2 * e]);
        float b = fmaf(__high2float(h2), sc[2 * e + 1], sh[2 * e + 1]);
        if (silu) {
          a = __fdividef(a, 1.f + __expf(-a));
          b = __fdividef(b, 1.f + __expf(-b));
        }
        const __nv_bfloat162 o2 = __floats2bfloat162_rn(a, b);
        w[e] = *reinterpret_cast<const uint32_t*>(&o2);
      }
      *reinterpret_cast<uint4*>(yb + (r + k * RY) * C) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}
}  // namespace

bool gn_apply_cs_supported(const Act& xa, const Act* xb, const Act& out) {
  const int C = xa.c + (xb ? xb->c : 0);
  if (xa.dt != BF16 || out.dt != BF16 || !xa.colsum || xa.c % 224 || C % 224 || C / 8 > 256 || out.c != C) return false;
  if (xb && (xb->dt != BF16 || !xb->colsum || xb->c % 224 || xb->colsum_rows != xa.colsum_rows || xb->rows() != xa.rows())) return false;
  return true;
}

void gn_apply_cs(const Act& xa, const Act* xb, const float* gamma, const float* beta, int groups, float eps, bool silu, const Act& out,
                 const Act* cat, cudaStream_t s) {
  if (dbg_skip("gn_apply")) return;
  ECHO_CHECK(gn_apply_cs_supported(xa, xb, out), "gn_apply_cs: unsupported operands");
  const int C = out.c, noct = C / 8, RY = 256 / noct;
  const int threads = ((noct * RY + 31) / 32) * 32;
  const int64_t V = xa.voxels();
  // 4 passes of 4 rows in flight per thread, fewer when that would leave the chip with less than ~2 blocks per SM (the
  // coarse levels are latency-, not bandwidth-bound)
  int rows_per_block = 16 * RY;
  while (rows_per_block > 2 * RY && (int64_t)cdiv(V, rows_per_block) * xa.n < 2 * 148) rows_per_block -= 2 * RY;
  {   // never a second, nearly empty wave (464 blocks on 444 resident slots cost 2x): grow the blocks until the grid is resident at once
    static int resident = 0;
    if (!resident) {
      int per_sm = 0, sms = 148;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_apply_cs_kernel, 256, 4096 + 256);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
      resident = (per_sm > 0 ? per_sm : 2) * sms;
    }
    while ((int64_t)cdiv(V, rows_per_block) * xa.n > resident && rows_per_block < V) rows_per_block += 4 * RY;
  }
  dim3 grid(cdiv(V, rows_per_block), xa.n);
  const size_t smem = (size_t)(C / 7 > threads ? C / 7 : threads) * 2 * sizeof(double) + (size_t)groups * 2 * sizeof(float);
  launch_pdl(gn_apply_cs_kernel, grid, dim3(threads), smem, s, (const __nv_bfloat16*)xa.p, xa.c, (const float*)xa.colsum,
             xb ? (const __nv_bfloat16*)xb->p : (const __nv_bfloat16*)nullptr, xb ? xb->c : 0, xb ? (const float*)xb->colsum : (const float*)nullptr,
             xa.colsum_rows, (double)V * (C / groups), eps, gamma, beta, V, groups, silu ? 1 : 0, noct, RY, rows_per_block,
             (__nv_bfloat16*)out.p, cat ? (__nv_bfloat16*)cat->p : (__nv_bfloat16*)nullptr);
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
