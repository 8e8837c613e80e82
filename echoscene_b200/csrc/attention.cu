// Self-attention of the SpatialTransformer3D blocks (attention.py:174-219) for ECHO_PREC_BF16: a flash-style kernel
// that never materialises the (tokens x tokens) score matrix (the reference allocates (N*8, 1024, 1024) fp32 = 537 MB
// at N = 16).  bf16 operands, fp32 scores / softmax / accumulation.
//
// Layout: the QKV projection writes [rows, 3 * heads * DHP] with every head padded from dh = 56 / 84 to DHP = 64 / 96
// columns (the padding rows of the projection weight are zero, so q.k is unchanged and the padded V columns are
// never stored).  One CTA = 64 queries of one (object, head); 4 warps x 16 query rows; K/V stream through shared
// memory in 64-key tiles (cp.async, double buffered); S = Q K^T and O += P V run on mma.sync.m16n8k16 (the legacy
// warp-level tensor path: attention is ~2 % of the step's FLOPs, SURVEY.md Appendix E — the tcgen05 budget goes to the
// convolutions).  Softmax is the online (running max / running sum) form in the exp2 domain.
#include "ops.cuh"

namespace echo {
namespace {

constexpr int BQ = 64, BKV = 64, NTHREADS = 128;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MT = 16-row query tiles per warp (CTA = 64*MT queries): every K / V fragment pulled out of shared memory with one
// ldmatrix.x4 feeds 2*MT MMAs, which is what the warp-level tensor path needs to stay off the shared-memory port.
template <int DHP, int MT>
__global__ void __launch_bounds__(NTHREADS) flash_attn_kernel(const __nv_bfloat16* __restrict__ qkv, int tokens, int heads, int dh,
                                                              float scale_log2e, __nv_bfloat16* __restrict__ out) {
  constexpr int LDS = DHP + 8;            // padded smem row (elements): conflict-free ldmatrix
  constexpr int KSTEPS = DHP / 16;        // k-steps of Q K^T
  constexpr int ONT = DHP / 8;            // n-tiles of O
  constexpr int CHUNKS = DHP / 8;         // 16-byte chunks per row
  constexpr int BQT = BQ * MT;            // queries per CTA
  static_assert(KSTEPS % 2 == 0 && ONT % 2 == 0, "fragments are fetched in pairs");
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* Ks = Qs + BQT * LDS;     // [2][BKV][LDS]
  __nv_bfloat16* Vs = Ks + 2 * BKV * LDS; // [2][BKV][LDS]

  griddep_launch();
  griddep_wait();
  const int qt = blockIdx.x, head = blockIdx.y, obj = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ld = (int64_t)3 * heads * DHP;
  const __nv_bfloat16* base = qkv + (int64_t)obj * tokens * ld;
  const __nv_bfloat16* qg = base + (int64_t)qt * BQT * ld + head * DHP;
  const __nv_bfloat16* kg = base + (int64_t)heads * DHP + head * DHP;
  const __nv_bfloat16* vg = base + (int64_t)2 * heads * DHP + head * DHP;

  auto load_tile = [&](__nv_bfloat16* dst, const __nv_bfloat16* src, int rows) {
    for (int i = tid; i < rows * CHUNKS; i += NTHREADS) {
      const int r = i / CHUNKS, c = i - r * CHUNKS;
      cp_async16(dst + r * LDS + c * 8, src + (int64_t)r * ld + c * 8);
    }
  };
  load_tile(Qs, qg, BQT);
  load_tile(Ks, kg, BKV);
  load_tile(Vs, vg, BKV);
  cp_async_commit();

  float o[MT][ONT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int i = 0; i < ONT; ++i) o[mt][i][0] = o[mt][i][1] = o[mt][i][2] = o[mt][i][3] = 0.f;
  float m_run[MT][2], l_run[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) { m_run[mt][0] = m_run[mt][1] = -INFINITY; l_run[mt][0] = l_run[mt][1] = 0.f; }
  uint32_t qf[MT][KSTEPS][4];

  const int ntiles = tokens / BKV;
  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) {   // prefetch the next K/V tile into the other buffer
      load_tile(Ks + (buf ^ 1) * BKV * LDS, kg + (int64_t)(t + 1) * BKV * ld, BKV);
      load_tile(Vs + (buf ^ 1) * BKV * LDS, vg + (int64_t)(t + 1) * BKV * ld, BKV);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks)
          ldsm_x4(qf[mt][ks], Qs + ((warp * MT + mt) * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + ks * 16 + (lane >> 4) * 8);
    }
    const __nv_bfloat16* Kt = Ks + buf * BKV * LDS;
    const __nv_bfloat16* Vt = Vs + buf * BKV * LDS;

    // S = Q K^T for this warp's 16*MT rows x 64 keys
    float s[MT][BKV / 8][4];
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) s[mt][j][0] = s[mt][j][1] = s[mt][j][2] = s[mt][j][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KSTEPS; ks += 2) {
        uint32_t bf[4];   // {ks: k 0-7, k 8-15, ks+1: k 0-7, k 8-15} of keys j*8 .. j*8+7
        ldsm_x4(bf, Kt + (j * 8 + (lane & 7)) * LDS + ks * 16 + (lane >> 3) * 8);
        const uint32_t b0[2] = {bf[0], bf[1]}, b1[2] = {bf[2], bf[3]};
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16(s[mt][j], qf[mt][ks], b0);
          mma_bf16(s[mt][j], qf[mt][ks + 1], b1);
        }
      }
    }
    // online softmax (rows lane/4 and lane/4 + 8 of each 16-row tile)
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < BKV / 8; ++j) {
        mx[0] = fmaxf(mx[0], fmaxf(s[mt][j][0], s[mt][j][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[mt][j][2], s[mt][j][3]));
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      }
      float corr[2], msc[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float mn = fmaxf(m_run[mt][r], mx[r]);
        corr[r] = ex2_approx((m_run[mt][r] - mn) * scale_log2e);
        m_run[mt][r] = mn;
        msc[r] = mn * scale_log2e;
        l_run[mt][r] *= corr[r];
      }
      float rs[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < BKV / 8; ++j) {
        s[mt][j][0] = ex2_approx(fmaf(s[mt][j][0], scale_log2e, -msc[0]));
        s[mt][j][1] = ex2_approx(fmaf(s[mt][j][1], scale_log2e, -msc[0]));
        s[mt][j][2] = ex2_approx(fmaf(s[mt][j][2], scale_log2e, -msc[1]));
        s[mt][j][3] = ex2_approx(fmaf(s[mt][j][3], scale_log2e, -msc[1]));
        rs[0] += s[mt][j][0] + s[mt][j][1];
        rs[1] += s[mt][j][2] + s[mt][j][3];
      }
      l_run[mt][0] += rs[0];
      l_run[mt][1] += rs[1];
#pragma unroll
      for (int i = 0; i < ONT; ++i) {
        o[mt][i][0] *= corr[0]; o[mt][i][1] *= corr[0];
        o[mt][i][2] *= corr[1]; o[mt][i][3] *= corr[1];
      }
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t pf[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        pf[mt][0] = pack_bf16(s[mt][2 * kk][0], s[mt][2 * kk][1]);
        pf[mt][1] = pack_bf16(s[mt][2 * kk][2], s[mt][2 * kk][3]);
        pf[mt][2] = pack_bf16(s[mt][2 * kk + 1][0], s[mt][2 * kk + 1][1]);
        pf[mt][3] = pack_bf16(s[mt][2 * kk + 1][2], s[mt][2 * kk + 1][3]);
      }
#pragma unroll
      for (int i = 0; i < ONT; i += 2) {
        uint32_t bf[4];   // {dims i*8..: keys 0-7, keys 8-15, dims (i+1)*8..: keys 0-7, keys 8-15} of keys kk*16 ..
        ldsm_x4_trans(bf, Vt + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + (i + (lane >> 4)) * 8);
        const uint32_t b0[2] = {bf[0], bf[1]}, b1[2] = {bf[2], bf[3]};
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          mma_bf16(o[mt][i], pf[mt], b0);
          mma_bf16(o[mt][i + 1], pf[mt], b1);
        }
      }
    }
    __syncthreads();   // everyone is done with `buf` before it is refilled two iterations later
  }
  // finalise: the row sums live distributed over the 4 lanes of a quad
  const int C = heads * dh;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[mt][r] += __shfl_xor_sync(0xffffffffu, l_run[mt][r], 1);
      l_run[mt][r] += __shfl_xor_sync(0xffffffffu, l_run[mt][r], 2);
    }
    const float inv0 = 1.f / l_run[mt][0], inv1 = 1.f / l_run[mt][1];
    const int64_t row0 = (int64_t)obj * tokens + qt * BQT + (warp * MT + mt) * 16 + (lane >> 2);
    __nv_bfloat16* o0 = out + row0 * C + head * dh;
    __nv_bfloat16* o1 = o0 + (int64_t)8 * C;
#pragma unroll
    for (int i = 0; i < ONT; ++i) {
      const int dcol = i * 8 + (lane & 3) * 2;
      if (dcol < dh) {
        *reinterpret_cast<uint32_t*>(o0 + dcol) = pack_bf16(o[mt][i][0] * inv0, o[mt][i][1] * inv0);
        *reinterpret_cast<uint32_t*>(o1 + dcol) = pack_bf16(o[mt][i][2] * inv1, o[mt][i][3] * inv1);
      }
    }
  }
}

// qkv fp32 [rows, 3*heads*dh] -> bf16 [rows, 3*heads*dhp] zero padded (single-operator entry point only)
__global__ void pad_qkv_kernel(const float* __restrict__ x, int64_t rows, int heads, int dh, int dhp, __nv_bfloat16* __restrict__ y) {
  const int64_t total = rows * 3 * heads * dhp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % dhp); int64_t r = i / dhp;
    const int hh = (int)(r % (3 * heads)); const int64_t row = r / (3 * heads);
    y[i] = __float2bfloat16(d < dh ? x[(row * 3 * heads + hh) * dh + d] : 0.f);
  }
}

}  // namespace

int attention_pad_dh(int dh) { return dh <= 64 ? 64 : (dh <= 96 ? 96 : 0); }

bool attention_bf16_supported(int tokens, int dh) { return tokens % 64 == 0 && tokens > 0 && attention_pad_dh(dh) != 0 && dh % 2 == 0; }

template <int DHP, int MT>
static void launch_flash(const __nv_bfloat16* qkv, int n, int tokens, int heads, int dh, float scale_log2e, __nv_bfloat16* out, cudaStream_t s) {
  dim3 grid(tokens / (BQ * MT), heads, n);
  const size_t smem = (size_t)(BQ * MT + 4 * BKV) * (DHP + 8) * sizeof(__nv_bfloat16);
  static bool attr = false;
  if (!attr) {
    ECHO_CUDA(cudaFuncSetAttribute(flash_attn_kernel<DHP, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  launch_pdl(flash_attn_kernel<DHP, MT>, grid, dim3(NTHREADS), smem, s, qkv, tokens, heads, dh, scale_log2e, out);
}

void attention_bf16(const __nv_bfloat16* qkv, int n, int tokens, int heads, int dh, __nv_bfloat16* out, cudaStream_t s) {
  if (dbg_skip("flash")) return;
  ECHO_CHECK(attention_bf16_supported(tokens, dh), "attention_bf16: tokens=%d dh=%d unsupported", tokens, dh);
  const int dhp = attention_pad_dh(dh);
  const float scale_log2e = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  // 32 query rows per warp when that still leaves >= 2 CTAs per SM worth of blocks, else 16
  const bool wide = tokens % (2 * BQ) == 0 && (int64_t)(tokens / (2 * BQ)) * heads * n >= 2 * 148;
  if (dhp == 64) {
    if (wide) launch_flash<64, 2>(qkv, n, tokens, heads, dh, scale_log2e, out, s);
    else launch_flash<64, 1>(qkv, n, tokens, heads, dh, scale_log2e, out, s);
  } else {
    if (wide) launch_flash<96, 2>(qkv, n, tokens, heads, dh, scale_log2e, out, s);
    else launch_flash<96, 1>(qkv, n, tokens, heads, dh, scale_log2e, out, s);
  }
  ECHO_LAUNCH_CHECK();
}

void pad_qkv(const float* x, int64_t rows, int heads, int dh, int dhp, __nv_bfloat16* y, cudaStream_t s) {
  int64_t total = rows * 3 * heads * dhp;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pad_qkv_kernel<<<(int)blocks, 256, 0, s>>>(x, rows, heads, dh, dhp, y);
  ECHO_LAUNCH_CHECK();
}

}  // namespace echo
