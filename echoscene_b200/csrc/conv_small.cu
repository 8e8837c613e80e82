// 3x3x3 convolution with a handful of output channels (the UNet's final 224 -> 3 conv, openai_model_3d.py:730-737).
// A GEMM tile would waste > 95 % of its columns here; the op is a streaming reduction instead: the whole filter
// (cout*27*cin floats) sits in shared memory, one warp per output voxel walks the 27 taps with 16-byte channel
// loads (neighbouring voxels' rows come from L1/L2), and a shuffle tree finishes the dot products.
#include "ops.cuh"

namespace echo {
namespace {

template <class T>
__device__ __forceinline__ void lda4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void lda4<float>(const float* p, float (&v)[4]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void lda4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}

template <class TA, int COUT>
__global__ void __launch_bounds__(256) conv3d_small_cout_kernel(const TA* __restrict__ x, int n, int d, int h, int w, int cin,
                                                                const float* __restrict__ wt, const float* __restrict__ bias,
                                                                float* __restrict__ out) {
  extern __shared__ __align__(16) float sw[];   // [COUT][27][cin]
  const int wtot = COUT * 27 * cin;
  for (int i = threadIdx.x * 4; i < wtot; i += blockDim.x * 4)
    *reinterpret_cast<float4*>(sw + i) = __ldg(reinterpret_cast<const float4*>(wt + i));
  __syncthreads();
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t rows = (int64_t)n * d * h * w;
  for (int64_t m = (int64_t)blockIdx.x * wpb + wib; m < rows; m += (int64_t)gridDim.x * wpb) {
    int64_t r = m;
    const int ow = (int)(r % w); r /= w;
    const int oh = (int)(r % h); r /= h;
    const int od = (int)(r % d); const int64_t obj = r / d;
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
    for (int tap = 0; tap < 27; ++tap) {
      const int id = od + tap / 9 - 1, ih = oh + (tap / 3) % 3 - 1, iw = ow + tap % 3 - 1;
      if ((unsigned)id >= (unsigned)d || (unsigned)ih >= (unsigned)h || (unsigned)iw >= (unsigned)w) continue;
      const TA* p = x + (((obj * d + id) * h + ih) * w + iw) * (int64_t)cin;
      for (int c = lane * 4; c < cin; c += 128) {
        float a[4];
        lda4<TA>(p + c, a);
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
          const float4 ww = *reinterpret_cast<const float4*>(sw + (co * 27 + tap) * cin + c);
          acc[co] = fmaf(a[0], ww.x, acc[co]);
          acc[co] = fmaf(a[1], ww.y, acc[co]);
          acc[co] = fmaf(a[2], ww.z, acc[co]);
          acc[co] = fmaf(a[3], ww.w, acc[co]);
        }
      }
    }
#pragma unroll
    for (int co = 0; co < COUT; ++co) {
      float v = acc[co];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) out[m * COUT + co] = v + (bias ? bias[co] : 0.f);
    }
  }
}

}  // namespace

bool conv3d_small_cout_supported(int cin, int cout, int taps) {
  return cout == 3 && taps == 27 && cin % 4 == 0 && (size_t)cout * 27 * cin * 4 <= 200 * 1024;
}

// x channels-last (n,d,h,w,cin), stride 1, pad 1; weights repacked [cout][27][cin] fp32; out [rows][cout] fp32
void conv3d_small_cout(const Act& x, const float* wt, const float* bias, int cout, float* out, cudaStream_t s) {
  ECHO_CHECK(conv3d_small_cout_supported(x.c, cout, 27), "conv3d_small_cout: unsupported shape");
  const size_t smem = (size_t)cout * 27 * x.c * sizeof(float);
  const int grid = 148 * 2;
  if (x.dt == F32) {
    static bool attr = false;
    if (!attr) { ECHO_CUDA(cudaFuncSetAttribute(conv3d_small_cout_kernel<float, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    conv3d_small_cout_kernel<float, 3><<<grid, 256, smem, s>>>((const float*)x.p, x.n, x.d, x.h, x.w, x.c, wt, bias, out);
  } else {
    static bool attr = false;
    if (!attr) { ECHO_CUDA(cudaFuncSetAttribute(conv3d_small_cout_kernel<__nv_bfloat16, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr = true; }
    conv3d_small_cout_kernel<__nv_bfloat16, 3><<<grid, 256, smem, s>>>((const __nv_bfloat16*)x.p, x.n, x.d, x.h, x.w, x.c, wt, bias, out);
  }
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
