// fp32-grade GEMM on the tensor cores: C[M, N] (+)= A B with every operand split as x = hi + lo (hi = the TF32 truncation of x, lo the
// exact fp32 remainder) and three mma.sync.m16n8k8.tf32 per product -- hi*hi into one fp32 accumulator, lo*hi + hi*lo into a second
// one, added in the epilogue.  The products dropped (lo*lo) are below 2^-22 of the term, so results agree with an fp32 FMA chain to
// ~1e-6, which the training path's parity contract needs (single-pass TF32 does not: SURVEY section 7, "Hard parts").
//
// Used where rows go into the hundreds or thousands and weights are a few MB (the GraphTripleConvNet forward / dgrad / wgrad of a
// collated batch, gcn_train.cu): operands are addressed by (row stride, column stride) so that Y = X W^T, dX = dY W and dW = dY^T X
// run on the same kernel without transposed copies, and the reduction can be split over blockIdx.z into equal chunks whose partial
// tiles land in a workspace (summed in chunk order by the caller: deterministic).
//
// Tile 128 x 64 x 16 per 256-thread CTA, eight warps of 32 x 32 (2 x 4 MMA tiles), operands staged k-major in shared memory with
// rows padded to 8 mod 32 floats so that every fragment load is conflict-free, register-prefetched global loads (float4 along
// whichever operand dimension is contiguous).
#include "ops.cuh"

#include <algorithm>

namespace echo {
namespace {

constexpr int XM = 128, XN = 64, XK = 16, XT = 256, XA_LD = XM + 8, XB_LD = XN + 8;

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A_KC: A's reduction index is the contiguous one (row-major [M, K]); else its row index is (A read transposed).
// B_NC: B's column index is contiguous (row-major [K, N]); else its reduction index is (B = W^T of a row-major [N, K] weight).
template <bool A_KC, bool B_NC>
__global__ void __launch_bounds__(XT, 2) sgemm_x3_kernel(const SgemmX3Args g) {
  __shared__ __align__(16) float As[XK][XA_LD];
  __shared__ __align__(16) float Bs[XK][XB_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gq = lane >> 2, tq = lane & 3;
  const int m0 = blockIdx.x * XM, n0 = blockIdx.y * XN, z = blockIdx.z;
  const int kbeg = z * g.chunk, kend = min(g.K, kbeg + g.chunk);
  const float* __restrict__ A = g.A;
  const float* __restrict__ B = g.B;

  float4 ra[2], rb;
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + XT * i;
      ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A_KC) {
        const int m = m0 + (idx & 127), k = k0 + (idx >> 7) * 4;
        if (m < g.M && k < kend) ra[i] = __ldg(reinterpret_cast<const float4*>(A + (int64_t)m * g.sam + k));
      } else {
        const int m = m0 + (idx & 31) * 4, k = k0 + (idx >> 5);
        if (m < g.M && k < kend) ra[i] = __ldg(reinterpret_cast<const float4*>(A + (int64_t)k * g.sak + m));
      }
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (B_NC) {
      const int n = n0 + (tid & 15) * 4, k = k0 + (tid >> 4);
      if (n < g.N && k < kend) rb = __ldg(reinterpret_cast<const float4*>(B + (int64_t)k * g.sbk + n));
    } else {
      const int n = n0 + (tid & 63), k = k0 + (tid >> 6) * 4;
      if (n < g.N && k < kend) rb = __ldg(reinterpret_cast<const float4*>(B + (int64_t)n * g.sbn + k));
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = tid + XT * i;
      if (A_KC) {
        const int m = idx & 127, k = (idx >> 7) * 4;
        As[k][m] = ra[i].x; As[k + 1][m] = ra[i].y; As[k + 2][m] = ra[i].z; As[k + 3][m] = ra[i].w;
      } else {
        *reinterpret_cast<float4*>(&As[idx >> 5][(idx & 31) * 4]) = ra[i];
      }
    }
    if (B_NC) {
      *reinterpret_cast<float4*>(&Bs[tid >> 4][(tid & 15) * 4]) = rb;
    } else {
      const int n = tid & 63, k = (tid >> 6) * 4;
      Bs[k][n] = rb.x; Bs[k + 1][n] = rb.y; Bs[k + 2][n] = rb.z; Bs[k + 3][n] = rb.w;
    }
  };

  float acc[2][4][4], acs[2][4][4];   // hi*hi, and the two cross terms
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = acs[i][j][q] = 0.f;

  const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
  load_tiles(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += XK) {
    store_tiles();
    __syncthreads();
    if (k0 + XK < kend) load_tiles(k0 + XK);
#pragma unroll
    for (int ks = 0; ks < XK; ks += 8) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = wm + mt * 16 + gq;
        split_tf32(As[ks + tq][r], ah[mt][0], al[mt][0]);
        split_tf32(As[ks + tq][r + 8], ah[mt][1], al[mt][1]);
        split_tf32(As[ks + tq + 4][r], ah[mt][2], al[mt][2]);
        split_tf32(As[ks + tq + 4][r + 8], ah[mt][3], al[mt][3]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = wn + nt * 8 + gq;
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32(Bs[ks + tq][c], bh0, bl0);
        split_tf32(Bs[ks + tq + 4][c], bh1, bl1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          mma_tf32(acs[mt][nt], al[mt], bh0, bh1);
          mma_tf32(acs[mt][nt], ah[mt], bl0, bl1);
          mma_tf32(acc[mt][nt], ah[mt], bh0, bh1);
        }
      }
    }
    __syncthreads();
  }

  float* C = g.C + (int64_t)z * g.c_bs;   // may alias g.res (accumulate in place): every element is read and written by one thread
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int m = m0 + wm + mt * 16 + gq + half * 8;
      if (m >= g.M) continue;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = n0 + wn + nt * 8 + 2 * tq;
        if (n >= g.N) continue;
        float2 v = make_float2(acc[mt][nt][half * 2] + acs[mt][nt][half * 2], acc[mt][nt][half * 2 + 1] + acs[mt][nt][half * 2 + 1]);
        if (g.bias) { v.x += __ldg(g.bias + n); v.y += __ldg(g.bias + n + 1); }
        if (g.act == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
        else if (g.act == 2) { v.x = v.x / (1.f + expf(-v.x)); v.y = v.y / (1.f + expf(-v.y)); }
        if (g.res) {
          const float2 r = *reinterpret_cast<const float2*>(g.res + (int64_t)m * g.ld_res + n);
          v.x += r.x; v.y += r.y;
        }
        if (g.res2) {
          const float2 r = *reinterpret_cast<const float2*>(g.res2 + (int64_t)m * g.ld_res2 + n);
          v.x += r.x; v.y += r.y;
        }
        *reinterpret_cast<float2*>(C + (int64_t)m * g.ldc + n) = v;
      }
    }
}

// out[m, n] = epilogue(sum over the chunks of a split reduction, in chunk order)
__global__ void sgemm_x3_splitk_out_kernel(const float* __restrict__ ws, int M, int N, int splits, const float* __restrict__ bias, int act,
                                           const float* __restrict__ res, int64_t ld_res, const float* __restrict__ res2, int64_t ld_res2,
                                           float* out, int64_t ldo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, mn = (int64_t)M * N;
  if (i >= mn) return;
  const int m = (int)(i / N), n = (int)(i % N);
  float t = 0.f;
  for (int b = 0; b < splits; ++b) t += ws[(int64_t)b * mn + i];
  if (bias) t += __ldg(bias + n);
  if (act == 1) t = fmaxf(t, 0.f);
  else if (act == 2) t = t / (1.f + expf(-t));
  if (res) t += res[(int64_t)m * ld_res + n];
  if (res2) t += res2[(int64_t)m * ld_res2 + n];
  out[(int64_t)m * ldo + n] = t;
}

}  // namespace

bool sgemm_x3_supported(const SgemmX3Args& g) {
  const bool a_kc = g.sak == 1, a_mc = g.sam == 1, b_nc = g.sbn == 1, b_kc = g.sbk == 1;
  if (!(a_kc || a_mc) || !(b_nc || b_kc)) return false;
  auto al16 = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
  if (!al16(g.A) || !al16(g.B) || ((uintptr_t)g.C % 8) || (g.res && ((uintptr_t)g.res % 8)) || (g.res2 && ((uintptr_t)g.res2 % 8))) return false;
  if ((a_kc ? g.sam : g.sak) % 4 || (b_nc ? g.sbk : g.sbn) % 4 || g.ldc % 2 || g.c_bs % 2 || (g.res && g.ld_res % 2) ||
      (g.res2 && g.ld_res2 % 2)) return false;
  if (g.splits > 1 && (g.bias || g.res || g.res2 || g.act)) return false;
  // the contiguous extent must be a multiple of the float4 a thread moves; a split reduction needs chunks of whole k-tiles
  if (a_kc ? g.K % 4 : g.M % 4) return false;
  if (b_nc ? g.N % 4 : g.K % 4) return false;
  if (g.N % 2) return false;
  if (g.splits > 1 && (g.chunk % XK || (int64_t)(g.splits - 1) * g.chunk >= g.K || (int64_t)g.splits * g.chunk < g.K)) return false;
  return true;
}

void sgemm_x3(const SgemmX3Args& a, cudaStream_t s) {
  SgemmX3Args g = a;
  if (g.M == 0 || g.N == 0) return;
  if (g.splits <= 1) { g.splits = 1; g.chunk = g.K; }
  ECHO_CHECK(g.A && g.B && g.C && g.K > 0, "sgemm_x3: null operand");
  ECHO_CHECK(sgemm_x3_supported(g), "sgemm_x3: operand strides / alignment outside the kernel's contract");
  dim3 grid(cdiv(g.M, XM), cdiv(g.N, XN), g.splits);
  ECHO_CHECK(grid.y <= 65535 && grid.z <= 65535, "sgemm_x3: grid too large");
  const bool a_kc = g.sak == 1, b_nc = g.sbn == 1;
  if (a_kc && b_nc) sgemm_x3_kernel<true, true><<<grid, XT, 0, s>>>(g);
  else if (a_kc) sgemm_x3_kernel<true, false><<<grid, XT, 0, s>>>(g);
  else if (b_nc) sgemm_x3_kernel<false, true><<<grid, XT, 0, s>>>(g);
  else sgemm_x3_kernel<false, false><<<grid, XT, 0, s>>>(g);
  ECHO_LAUNCH_CHECK();
}

// sgemm_x3 with the reduction split over blockIdx.z whenever the tile grid would leave most SMs idle (a few hundred rows against a
// wide reduction) and `ws` has room for the partial tiles; the partials are summed in chunk order, then the epilogue is applied.
void sgemm_x3_auto(const SgemmX3Args& g, float* ws, size_t ws_floats, cudaStream_t s) {
  if (g.M == 0 || g.N == 0) return;
  const int tiles = cdiv(g.M, XM) * cdiv(g.N, XN);
  const size_t mn = (size_t)g.M * g.N;
  int splits = std::min({cdiv(148, tiles), g.K / 64, ws ? (int)(ws_floats / mn) : 1, 16});
  if (splits >= 2) {
    const int chunk = (cdiv(g.K, splits) + XK - 1) / XK * XK;
    splits = cdiv(g.K, chunk);
    SgemmX3Args p = g;
    p.C = ws; p.ldc = g.N; p.bias = nullptr; p.res = nullptr; p.res2 = nullptr; p.act = 0; p.splits = splits; p.chunk = chunk;
    p.c_bs = (int64_t)mn;
    if (splits >= 2 && sgemm_x3_supported(p)) {
      sgemm_x3(p, s);
      sgemm_x3_splitk_out_kernel<<<cdiv((int64_t)mn, 256), 256, 0, s>>>(ws, g.M, g.N, splits, g.bias, g.act, g.res, g.ld_res, g.res2, g.ld_res2,
                                                                        g.C, g.ldc);
      ECHO_LAUNCH_CHECK();
      return;
    }
  }
  sgemm_x3(g, s);
}

}  // namespace echo
