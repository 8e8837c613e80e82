// Self-attention of the 1024-token SpatialTransformer3D blocks (attention.py:174-219) on the 5th-generation tensor
// cores: S = Q K^T and O_j = P_j V_j are tcgen05.mma with accumulators in TMEM, softmax runs in registers straight out
// of TMEM, the (tokens x tokens) score matrix never exists (the reference allocates (N*8, 1024, 1024) fp32).
//
// One CTA = 128 queries of one (object, head); two CTAs share an SM so one's softmax overlaps the other's MMAs.
//   warp 0      TMA producer: Q once, then K / V^T blocks of 128 keys through a 2-deep ring
//   warp 1      TMEM allocation (256 columns: S = 128, O_j = 64) + MMA issue
//   warps 2-5   softmax: one query row per thread (TMEM lane); per key block: row max over S (pass 1), p = 2^((s-m)c)
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

constexpr int FQ = 128, FKV = 128, FD = 64, FTHREADS = 192;
constexpr int TILE_BYTES = 128 * 128;                                 // 128 rows x 64 bf16
constexpr int SM_Q = 0, SM_K = TILE_BYTES, SM_V = 3 * TILE_BYTES, SM_P = 5 * TILE_BYTES, SM_BAR = 7 * TILE_BYTES;
constexpr int FLASH_SMEM = SM_BAR + 256;

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

__global__ void __launch_bounds__(FTHREADS, 2)
flash_tc_kernel(const __grid_constant__ CUtensorMap map_qk, const __grid_constant__ CUtensorMap map_vt, const FlashTcParams p) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + SM_BAR);
  uint64_t* q_full = bars;            // 1
  uint64_t* kv_full = bars + 1;       // [2]
  uint64_t* kv_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* o_full = bars + 8;
  uint64_t* o_free = bars + 9;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

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
    for (int i = 0; i < 2; ++i) { mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1); }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    mbar_init(o_free, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;
  griddep_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      const int row_q = obj * p.tokens + qt * FQ;
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(&map_qk, q_full, sm + SM_Q, head * FD, row_q);
      const int vt_row = (obj * p.heads + head) * FD;
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 2 * TILE_BYTES);
        tma_load_2d(&map_qk, &kv_full[st], sm + SM_K + st * TILE_BYTES, (p.heads + head) * FD, obj * p.tokens + j * FKV);
        tma_load_2d(&map_vt, &kv_full[st], sm + SM_V + st * TILE_BYTES, j * FKV, vt_row);                       // keys 0..63  x 64 d
        tma_load_2d(&map_vt, &kv_full[st], sm + SM_V + st * TILE_BYTES + TILE_BYTES / 2, j * FKV + 64, vt_row);  // keys 64..127
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // instruction descriptors: D = f32, A = B = bf16, K-major; M = 128; N = 128 (scores) / 64 (output block)
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FKV >> 3) << 17) | ((uint32_t)(FQ >> 4) << 24);
    const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(FD >> 3) << 17) | ((uint32_t)(FQ >> 4) << 24);
    mbar_wait(q_full, 0);
    for (int j = 0; j < nblk; ++j) {
      const int st = j & 1;
      mbar_wait(&kv_full[st], (j >> 1) & 1);
      mbar_wait(s_free, (j & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t dq = make_smem_desc(smem_u32(sm + SM_Q));
        const uint64_t dk = make_smem_desc(smem_u32(sm + SM_K + st * TILE_BYTES));
#pragma unroll
        for (int k = 0; k < FD / 16; ++k) umma_bf16(tmem_s, dq + (uint64_t)(k * 2), dk + (uint64_t)(k * 2), idesc_s, k ? 1u : 0u);
        umma_commit(s_full);
      }
      __syncwarp();
      mbar_wait(p_full, j & 1);
      mbar_wait(o_free, (j & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          const uint64_t dp = make_smem_desc(smem_u32(sm + SM_P + kb * TILE_BYTES));
          const uint64_t dv = make_smem_desc(smem_u32(sm + SM_V + st * TILE_BYTES + kb * (TILE_BYTES / 2)));
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(tmem_o, dp + (uint64_t)(k * 2), dv + (uint64_t)(k * 2), idesc_o, (kb | k) ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[st]);
      }
      __syncwarp();
    }
  } else {
    // ================= softmax / output accumulation: one query row per thread =================
    const int quarter = warp & 3;                   // TMEM lanes this warp may touch
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float c = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f;
    float o[FD];
#pragma unroll
    for (int i = 0; i < FD; ++i) o[i] = 0.f;
    uint8_t* prow = sm + SM_P + row * 128;
    const int sw = row & 7;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      float mx = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < FKV / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_addr + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      const float m_new = fmaxf(m_run, mx);
      const float corr = ex2f((m_run - m_new) * c);   // first block: 2^(-inf) = 0
      const float mc = m_new * c;
      float rs = 0.f;
      // P of block j-1 was consumed: its o_full was waited for below before this iteration started
#pragma unroll
      for (int cc = 0; cc < FKV / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_s + lane_addr + cc * 32, v);
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
      mbar_arrive(p_full);
      mbar_arrive(s_free);
      l_run = fmaf(l_run, corr, rs);
      m_run = m_new;
      mbar_wait(o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < FD / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem_o + lane_addr + cc * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[cc * 32 + i] = fmaf(o[cc * 32 + i], corr, __uint_as_float(v[i]));
      }
      tc_fence_before();
      mbar_arrive(o_free);
    }
    const float inv = 1.f / l_run;
    const int C = p.heads * p.dh;
    __nv_bfloat16* op = p.out + ((long long)obj * p.tokens + qt * FQ + row) * C + head * p.dh;
#pragma unroll
    for (int q = 0; q < FD / 8; ++q) {
      if (q * 8 < p.dh)
        *reinterpret_cast<uint4*>(op + q * 8) = make_uint4(pack2(o[q * 8] * inv, o[q * 8 + 1] * inv), pack2(o[q * 8 + 2] * inv, o[q * 8 + 3] * inv),
                                                          pack2(o[q * 8 + 4] * inv, o[q * 8 + 5] * inv), pack2(o[q * 8 + 6] * inv, o[q * 8 + 7] * inv));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
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
