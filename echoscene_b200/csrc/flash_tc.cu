// Self-attention of the 1024-token SpatialTransformer3D blocks (attention.py:174-219) on the 5th-generation tensor
// cores: S = Q K^T and O_j = P_j V_j are tcgen05.mma with accumulators in TMEM, softmax runs in registers straight out
// of TMEM, the (tokens x tokens) score matrix never exists (the reference allocates (N*8, 1024, 1024) fp32).
//
// One CTA = 128 queries of one (object, head), one CTA per SM.
//   warp 0      TMA producer: Q once, then K / V^T blocks of 128 keys through a 4-deep ring
//   warp 1      TMEM allocation (S[2] x 128 + O_j[2] x 64 columns) + MMA issue
//   warps 2-9   softmax, two groups of four warps: one query row per thread (TMEM lane); per key block: row max over S
//               (pass 1), p = 2^((s-m)c)
//               (pass 2) written as bf16 into shared memory in the K-major SWIZZLE_128B layout the second MMA reads as
//               its A operand; O_j is read back and folded into the running output  o = o * corr + O_j  in registers
//               (no TMEM read-modify-write, no rescale hazards).
// Operands are all K-major: Q [q x d], K [key x d], P [q x key], V^T [d x key] -- the last one is why the V projection
// GEMM stores its output transposed (GemmArgs::out_t).  Heads are zero-padded 56 -> 64 by the projection weights.
#include "ops.cuh"
#include "tc_ptx.cuh"

namespace echo {
using namespace ptx;

namespace {

constexpr int FQ = 128, FKV = 128, FD = 64, FTHREADS = 320, KV_STAGES = 4;
constexpr int TILE_BYTES = 128 * 128;                                 // 128 rows x 64 bf16
constexpr int SM_Q = 0, SM_K = TILE_BYTES, SM_V = SM_K + KV_STAGES * TILE_BYTES, SM_P = SM_V + KV_STAGES * TILE_BYTES,
              SM_BAR = SM_P + 4 * TILE_BYTES;
constexpr int FLASH_SMEM = SM_BAR + 256;
constexpr int TM_S = 0, TM_O = 256;                                   // TMEM columns: S[2] x 128, O[2] x 64

struct FlashTcParams {
  int tokens, heads, dh;
  float scale_log2e;
  __nv_bfloat16* out;
  int* err;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Pipeline: key blocks alternate between two softmax groups (even blocks -> group 0, odd -> group 1), each with its own
// S / P / O_j buffers and its own running (max, sum, output): two independent  MMA1 -> softmax -> MMA2 -> accumulate
// chains per CTA, so every SM sub-partition always has a second warp to issue from while one waits on a tensor-core or
// TMEM round trip; the groups are merged once at the end (the usual split-KV combine).
__global__ void __launch_bounds__(FTHREADS, 1)
flash_tc_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_vt, const FlashTcParams p) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BAR);
  uint64_t* q_full = bars;                  // 1
  uint64_t* kv_full = bars + 1;             // [KV_STAGES]
  uint64_t* kv_empty = kv_full + KV_STAGES; // [KV_STAGES]
  uint64_t* s_full = kv_empty + KV_STAGES;  // [2]
  uint64_t* p_full = s_full + 2;            // [2]  (also: S[g] has been read)
  uint64_t* o_full = p_full + 2;            // [2]
  uint64_t* o_free = o_full + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, obj = blockIdx.z;
  const int nblk = p.tokens / FKV;

  griddep_launch();
  if ((smem_u32(sm) & 1023u) != 0) {   // the swizzled layouts below assume a 1024-byte aligned window
    if (threadIdx.x == 0) atomicExch(p.err, 1);
    return;
  }
  if (warp == 0 && elect_one()) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KV_STAGES; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&p_full[i], 128); mbar_init(&o_full[i], 1); mbar_init(&o_free[i], 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      const int row_q = obj * p.tokens + qt * FQ;
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(&map_qk, q_full, sm + SM_Q, head * FD, row_q);
      const int vt_row = (obj * p.heads + head) * FD;
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < nblk; ++j) {
        mbar_wait(&kv_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * TILE_BYTES);
        tma_load_2d(&map_qk, &kv_full[st], sm + SM_K + st * TILE_BYTES, (p.heads + head) * FD, obj * p.tokens + j * FKV);
        tma_load_2d(&map_vt, &kv_full[st], sm + SM_V + st * TILE_BYTES, j * FKV, vt_row);                       // keys 0..63  x 64 d
        tma_load_2d(&map_vt, &kv_full[st], sm + SM_V + st * TILE_BYTES + TILE_BYTES / 2, j * FKV + 64, vt_row);  // keys 64..127
        if (++st == KV_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // instruction descriptors: D = f32, A = B = bf16, K-major; M = 128; N = 128 (scores) / 64 (output block)
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FKV >> 3) << 17) | ((uint32_t)(FQ >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FD >> 3) << 17) | ((uint32_t)(FQ >> 4) << 24);
    const uint64_t dq = make_smem_desc(smem_u32(sm + SM_Q));
    auto mma1 = [&](int j) {   // S[j&1] = Q K_j^T   (S[j&1] was drained: p_full of block j-2 has been waited for by MMA2(j-2))
      const int st = j % KV_STAGES;
      mbar_wait(&kv_full[st], (j / KV_STAGES) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dk = make_smem_desc(smem_u32(sm + SM_K + st * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < FD / 16; ++k)
          umma_bf16(tmem_base + TM_S + (j & 1) * 128, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
        umma_commit(&s_full[j & 1]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mma1(0);
    if (nblk > 1) mma1(1);
    for (int j = 0; j < nblk; ++j) {
      const int g = j & 1, st = j % KV_STAGES;
      const uint32_t gph = (j >> 1) & 1;
      mbar_wait(&p_full[g], gph);          // P[g] written and S[g] drained
      mbar_wait(&o_free[g], gph ^ 1);      // O[g] of block j-2 folded into the running output
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t dp = make_smem_desc(smem_u32(sm + SM_P + (g * 2 + kb) * TILE_BYTES));
          const uint64_t dv = make_smem_desc(smem_u32(sm + SM_V + st * TILE_BYTES + kb * (TILE_BYTES / 2)));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + TM_O + g * 64, dp + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), idesc_o, (kb | k) ? 1u : 0u);
        }
        umma_commit(&o_full[g]);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
      if (j + 2 < nblk) mma1(j + 2);
    }
  } else {
    // ================= softmax / output accumulation: one query row per thread, key blocks j = g, g+2, ... =================
    const int g = (warp - 2) >> 2;                  // softmax group
    const int quarter = warp & 3;                   // TMEM lanes this warp may touch
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t tmem_s = tmem_base + TM_S + g * 128 + lane_addr, tmem_o = tmem_base + TM_O + g * 64 + lane_addr;
    const float c = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f;
    float o[FD];
#pragma unroll
    for (int i = 0; i < FD; ++i) o[i] = 0.f;
    uint8_t* prow = sm + SM_P + g * 2 * TILE_BYTES + row * 128;
    const int sw = row & 7;
    for (int j = g; j < nblk; j += 2) {
      const uint32_t gph = (j >> 1) & 1;
      mbar_wait(&s_full[g], gph);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < FKV / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_s + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float corr = ex2f((m_run - m_new) * c);   // first block: 2^(-inf) = 0
      const float mc = m_new * c;
      float rs = 0.f;
      // P[g] of block j-2 has been consumed: its o_full was waited for below before this iteration started
#pragma unroll
      for (int cc = 0; cc < FKV / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_s + cc * 32, v);
        tmem_ld_wait();
        uint8_t* pk = prow + (cc >> 1) * TILE_BYTES;   // k-block of 64 keys
#pragma unroll
        for (int q = 0; q < 4; ++q) {                 // 16-byte chunks of 8 keys
          float e[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            e[i] = ex2f(fmaf(__uint_as_float(v[q * 8 + i]), c, -mc));
            rs += e[i];
          }
          const int chunk = (cc & 1) * 4 + q;         // chunk index inside the 128-byte row
          *reinterpret_cast<uint4*>(pk + ((chunk ^ sw) << 4)) = make_uint4(pack2(e[0], e[1]), pack2(e[2], e[3]), pack2(e[4], e[5]), pack2(e[6], e[7]));
        }
      }
      tc_fence_before();
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the MMA's async proxy
      mbar_arrive(&p_full[g]);
      l_run = fmaf(l_run, corr, rs);
      m_run = m_new;
      mbar_wait(&o_full[g], gph);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < FD / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_o + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[cc * 32 + i] = fmaf(o[cc * 32 + i], corr, __uint_as_float(v[i]));
      }
      tc_fence_before();
      mbar_arrive(&o_free[g]);
    }
    // ---- merge the two groups (split-KV combine) through shared memory: the P buffers are idle now ----
    float* xch = reinterpret_cast<float*>(sm + SM_P);   // [66][128] floats = 33 KiB, column-major so lanes hit distinct banks
    asm volatile("bar.sync 1, 256;" ::: "memory");     // both groups have seen their last o_full: no MMA still reads P
    if (g == 1) {
      xch[0 * 128 + row] = m_run;
      xch[1 * 128 + row] = l_run;
#pragma unroll
      for (int i = 0; i < FD; ++i) xch[(2 + i) * 128 + row] = o[i];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");     // the eight softmax warps
    if (g == 0) {
      const float m1 = xch[0 * 128 + row], l1 = xch[1 * 128 + row];
      const float m = fmaxf(m_run, m1);
      const float a0 = ex2f((m_run - m) * c), a1 = ex2f((m1 - m) * c);   // a group that saw no block has m = -inf -> weight 0
      const float inv = 1.f / fmaf(l_run, a0, l1 * a1);
      const float w0 = a0 * inv, w1 = a1 * inv;
      const int C = p.heads * p.dh;
      __nv_bfloat16* op = p.out + ((long long)obj * p.tokens + qt * FQ + row) * C + head * p.dh;
#pragma unroll
      for (int q = 0; q < FD / 8; ++q) {
        if (q * 8 < p.dh) {
          float r[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) r[i] = fmaf(o[q * 8 + i], w0, xch[(2 + q * 8 + i) * 128 + row] * w1);
          *reinterpret_cast<uint4*>(op + q * 8) = make_uint4(pack2(r[0], r[1]), pack2(r[2], r[3]), pack2(r[4], r[5]), pack2(r[6], r[7]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

int* g_flash_err = nullptr;

// qkv fp32 [rows, 3*heads*dh] -> qk bf16 [rows, 2*heads*64] (heads zero-padded) and vt bf16 [(n*heads*64), tokens]
// (single-operator entry point only: inside the step the projection GEMMs write these layouts directly)
__global__ void split_qkv_tc_kernel(const float* __restrict__ x, int n, int tokens, int heads, int dh, __nv_bfloat16* __restrict__ qk,
                                    __nv_bfloat16* __restrict__ vt) {
  const long long rows = (long long)n * tokens, total = rows * 3 * heads * 64;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % 64); long long r = i / 64;
    const int hh = (int)(r % (3 * heads)); const long long row = r / (3 * heads);
    const float v = d < dh ? x[(row * 3 * heads + hh) * dh + d] : 0.f;
    if (hh < 2 * heads) qk[(row * 2 * heads + hh) * 64 + d] = __float2bfloat16(v);
    else {
      const long long obj = row / tokens, tok = row - obj * tokens;
      vt[((obj * heads + (hh - 2 * heads)) * 64 + d) * tokens + tok] = __float2bfloat16(v);
    }
  }
}

}  // namespace

bool attention_tc_supported(int tokens, int dh) {
  static const bool off = getenv("ECHO_FLASH_LEGACY") != nullptr;   // A/B: force the mma.sync kernel
  return !off && tc_available() && tokens > 0 && tokens % FKV == 0 && dh <= FD && dh % 8 == 0;
}

void attention_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vt, int n, int tokens, int heads, int dh, __nv_bfloat16* out, cudaStream_t s) {
  if (dbg_skip("flash")) return;
  ECHO_CHECK(attention_tc_supported(tokens, dh), "attention_tc: tokens=%d dh=%d unsupported", tokens, dh);
  static bool init = false;
  if (!init) {
    ECHO_CUDA(cudaFuncSetAttribute(flash_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FLASH_SMEM));
    ECHO_CUDA(cudaMalloc((void**)&g_flash_err, sizeof(int)));
    ECHO_CUDA(cudaMemset(g_flash_err, 0, sizeof(int)));
    init = true;
  }
  CUtensorMap map_qk, map_vt;
  const uint64_t ld = (uint64_t)2 * heads * FD;
  tc_encode_2d_bf16(&map_qk, qk, ld, (uint64_t)n * tokens, ld * 2, FD, 128);
  tc_encode_2d_bf16(&map_vt, vt, (uint64_t)tokens, (uint64_t)n * heads * FD, (uint64_t)tokens * 2, 64, FD);
  FlashTcParams p;
  p.tokens = tokens; p.heads = heads; p.dh = dh;
  p.scale_log2e = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  p.out = out;
  p.err = g_flash_err;
  launch_pdl(flash_tc_kernel, dim3(tokens / FQ, heads, n), dim3(FTHREADS), (size_t)FLASH_SMEM, s, map_qk, map_vt, p);
  ECHO_LAUNCH_CHECK();
}

void split_qkv_tc(const float* qkv, int n, int tokens, int heads, int dh, __nv_bfloat16* qk, __nv_bfloat16* vt, cudaStream_t s) {
  split_qkv_tc_kernel<<<148 * 8, 256, 0, s>>>(qkv, n, tokens, heads, dh, qk, vt);
  ECHO_LAUNCH_CHECK();
}

// host-side check used by tests / the first call: did any block find a misaligned shared-memory window?
int attention_tc_error() {
  if (!g_flash_err) return 0;
  int v = 0;
  cudaMemcpy(&v, g_flash_err, sizeof(int), cudaMemcpyDeviceToHost);
  return v;
}

}  // namespace echo
