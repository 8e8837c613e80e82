// fp32-accumulate SIMT implicit GEMM (conv3d / linear / batched attention products).
//
// This is the ECHO_PREC_FP32 contraction: every product and sum is an fp32 FMA, which is what the 1e-3 parity
// contract against the reference's fp32 PyTorch path needs (SURVEY.md §7 "Hard parts": single-pass TF32 misses it
// on the shape step).  In ECHO_PREC_BF16 it also serves the few shapes the tcgen05 kernel (gemm_tc.cu) does not
// take (3-channel stem, 3-channel output conv, the 32/64-channel shape_embeddings convs).
//
// Tiling: 128x64 output tile per 256-thread CTA, BK = 16, 8x4 register tile per thread, register-prefetched
// global loads, smem transposed so both operand reads are conflict-free float4.
#include "ops.cuh"

namespace echo {

namespace {

constexpr int BM = 128, BN = 64, BK = 16, NT = 256;

template <class T>
__device__ __forceinline__ float ld1(const T* p);
template <>
__device__ __forceinline__ float ld1<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <class T>
__device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = __ldg(reinterpret_cast<const uint2*>(p));
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}

struct RowInfo {
  int64_t base;      // element offset of (obj, 0,0,0) in A
  int id0, ih0, iw0; // input coordinate of tap (0,0,0)
  bool valid;
};

template <class TA, class TW, bool FASTA, bool FASTB>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const GemmArgs g) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int b = blockIdx.z, b0 = b / g.nb1, b1 = b % g.nb1;
  const TA* __restrict__ A = reinterpret_cast<const TA*>(g.A) + b0 * g.a_bs0 + b1 * g.a_bs1;
  const TW* __restrict__ W = reinterpret_cast<const TW*>(g.W) + b0 * g.w_bs0 + b1 * g.w_bs1;
  const int64_t o_off = b0 * g.o_bs0 + b1 * g.o_bs1;

  const int64_t M = (int64_t)g.n * g.od * g.oh * g.ow;
  const int ovox = g.od * g.oh * g.ow;
  const int Ktot = g.kd * g.kh * g.kw * g.cin;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // ---- A-load mapping: rows tid/4 and tid/4+64, 4 consecutive k at (tid%4)*4 ---------------------------------
  const int a_kq = (tid & 3) * 4;
  RowInfo ri[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int64_t m = m0 + (tid >> 2) + i * 64;
    ri[i].valid = m < M;
    int64_t mm = ri[i].valid ? m : 0;
    int obj = (int)(mm / ovox);
    int r = (int)(mm - (int64_t)obj * ovox);
    int od = r / (g.oh * g.ow);
    r -= od * g.oh * g.ow;
    int oh = r / g.ow, ow = r - oh * g.ow;
    ri[i].base = (int64_t)obj * g.d * g.h * g.w * g.lda;
    ri[i].id0 = od * g.sd - g.pd;
    ri[i].ih0 = oh * g.sh - g.ph;
    ri[i].iw0 = ow * g.sw - g.pw;
  }
  // ---- B-load mapping: col tid/4, 4 consecutive k at (tid%4)*4 -------------------------------------------------
  const int b_col = tid >> 2, b_kq = (tid & 3) * 4;
  const bool b_valid = (n0 + b_col) < g.cout;

  float ra[2][4], rb[4];

  auto load_tiles = [&](int k0) {
    // A
    if (FASTA) {
      const int kk = k0 + a_kq;          // chunk lies inside one tap (cin % BK == 0)
      const int tap = kk / g.cin, c = kk - tap * g.cin;
      const int kd_ = tap / (g.kh * g.kw), rem = tap - kd_ * g.kh * g.kw, kh_ = rem / g.kw, kw_ = rem - kh_ * g.kw;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int id = ri[i].id0 + kd_, ih = ri[i].ih0 + kh_, iw = ri[i].iw0 + kw_;
        const bool ok = ri[i].valid && kk < Ktot && (unsigned)id < (unsigned)g.d && (unsigned)ih < (unsigned)g.h &&
                        (unsigned)iw < (unsigned)g.w;
        if (ok) {
          ld4<TA>(A + ri[i].base + (((int64_t)id * g.h + ih) * g.w + iw) * g.lda + c, ra[i]);
        } else {
          ra[i][0] = ra[i][1] = ra[i][2] = ra[i][3] = 0.f;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k0 + a_kq + j;
        const int tap = kk / g.cin, c = kk - tap * g.cin;
        const int kd_ = tap / (g.kh * g.kw), rem = tap - kd_ * g.kh * g.kw, kh_ = rem / g.kw, kw_ = rem - kh_ * g.kw;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int id = ri[i].id0 + kd_, ih = ri[i].ih0 + kh_, iw = ri[i].iw0 + kw_;
          const bool ok = ri[i].valid && kk < Ktot && (unsigned)id < (unsigned)g.d && (unsigned)ih < (unsigned)g.h &&
                          (unsigned)iw < (unsigned)g.w;
          ra[i][j] = ok ? ld1<TA>(A + ri[i].base + (((int64_t)id * g.h + ih) * g.w + iw) * g.lda + c) : 0.f;
        }
      }
    }
    // B
    if (FASTB) {
      const int kk = k0 + b_kq;
      if (b_valid && kk < Ktot) {
        ld4<TW>(W + (int64_t)(n0 + b_col) * g.w_stride_n + kk, rb);
      } else {
        rb[0] = rb[1] = rb[2] = rb[3] = 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kk = k0 + b_kq + j;
        rb[j] = (b_valid && kk < Ktot) ? ld1<TW>(W + (int64_t)(n0 + b_col) * g.w_stride_n + (int64_t)kk * g.w_stride_k) : 0.f;
      }
    }
  };

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int ty = tid >> 4, tx = tid & 15;

  load_tiles(0);
  for (int k0 = 0; k0 < Ktot; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) As[a_kq + j][(tid >> 2) + i * 64] = ra[i][j];
#pragma unroll
    for (int j = 0; j < 4; ++j) Bs[b_kq + j][b_col] = rb[j];
    __syncthreads();
    if (k0 + BK < Ktot) load_tiles(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 bb = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

  // ---- epilogue ---------------------------------------------------------------------------------------------------
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const int obj = (int)(m / ovox);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.cout) continue;
      float v = acc[i][j] * g.alpha;
      if (g.bias) v += __ldg(g.bias + n);
      if (g.rowvec) v += __ldg(g.rowvec + (int64_t)obj * g.ld_rowvec + n);
      if (g.res) {
        if (g.res_dt == F32) v += reinterpret_cast<const float*>(g.res)[o_off + m * g.ld_res + n];
        else v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g.res)[o_off + m * g.ld_res + n]);
      }
      if (g.act == 1) v = fmaxf(v, 0.f);
      if (g.out_dt == F32) reinterpret_cast<float*>(g.out)[o_off + m * g.ldo + n] = v;
      else reinterpret_cast<__nv_bfloat16*>(g.out)[o_off + m * g.ldo + n] = __float2bfloat16(v);
    }
  }
}

template <class TA, class TW>
void launch(const GemmArgs& g, bool fasta, bool fastb, dim3 grid, cudaStream_t s) {
  if (fasta && fastb) gemm_simt_kernel<TA, TW, true, true><<<grid, NT, 0, s>>>(g);
  else if (fasta) gemm_simt_kernel<TA, TW, true, false><<<grid, NT, 0, s>>>(g);
  else if (fastb) gemm_simt_kernel<TA, TW, false, true><<<grid, NT, 0, s>>>(g);
  else gemm_simt_kernel<TA, TW, false, false><<<grid, NT, 0, s>>>(g);
}

}  // namespace

void gemm_simt(const GemmArgs& g, cudaStream_t s) {
  ECHO_CHECK(g.A && g.W && g.out, "gemm: null operand");
  ECHO_CHECK(g.cin > 0 && g.cout > 0 && g.lda >= g.cin, "gemm: bad dims cin=%d cout=%d lda=%lld", g.cin, g.cout, (long long)g.lda);
  const int64_t M = g.rows_out();
  if (M == 0) return;
  const size_t ea = dt_size(g.a_dt), ew = dt_size(g.w_dt);
  const int Ktot = g.ktot();
  const bool fasta = (g.cin % BK == 0) && (g.lda % 4 == 0) && (((uintptr_t)g.A) % 16 == 0) && (g.a_bs0 % 4 == 0) &&
                     (g.a_bs1 % 4 == 0);
  const bool fastb = (g.w_stride_k == 1) && (Ktot % 4 == 0) && (g.w_stride_n % 4 == 0) && (((uintptr_t)g.W) % 16 == 0) &&
                     (g.w_bs0 % 4 == 0) && (g.w_bs1 % 4 == 0);
  (void)ea; (void)ew;
  dim3 grid(cdiv(M, BM), cdiv(g.cout, BN), g.nb0 * g.nb1);
  ECHO_CHECK(grid.y <= 65535 && grid.z <= 65535, "gemm: grid too large");
  if (g.a_dt == F32 && g.w_dt == F32) launch<float, float>(g, fasta, fastb, grid, s);
  else if (g.a_dt == BF16 && g.w_dt == BF16) launch<__nv_bfloat16, __nv_bfloat16>(g, fasta, fastb, grid, s);
  else if (g.a_dt == BF16 && g.w_dt == F32) launch<__nv_bfloat16, float>(g, fasta, fastb, grid, s);
  else launch<float, __nv_bfloat16>(g, fasta, fastb, grid, s);
  ECHO_LAUNCH_CHECK();
}

void gemm(const GemmArgs& g, int precision, cudaStream_t s) {
  if (precision == ECHO_PREC_BF16 && tc_available() && gemm_tc_supported(g)) gemm_tc(g, s);
  else gemm_simt(g, s);
}

}  // namespace echo
