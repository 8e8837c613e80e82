// bf16 implicit-GEMM convolution / linear on the 5th-generation tensor cores (sm_100a): TMA -> shared memory ->
// tcgen05.mma (accumulators in TMEM) -> tcgen05.ld epilogue.  This is the contraction kernel of ECHO_PREC_BF16:
// every 3x3x3 / 1x1x1 Conv3d and every token-wise Linear of the shape UNet runs here.
//
//   out[m, n] = act( sum_{tap, c} A[vox(m) + off(tap), c] * W[n, tap*cin + c] + bias[n] + rowvec[obj(m), n] + res[m, n] )
//
// A is a channels-last bf16 activation (obj, d, h, w, c) described to TMA as a 5-D tensor (c, w, h, d, obj).  One CTA
// tile is 128 output voxels forming a (bd, bh, bw) box inside one object; for every filter tap the SAME box shifted by
// the tap offset is fetched with one TMA box copy (64 channels x 128 voxels, 128-byte swizzle), and the hardware's
// out-of-bounds zero fill implements the convolution's zero padding, the channel tail (cin not a multiple of 64) and
// partial tiles.  The box lands in shared memory exactly in the K-major SWIZZLE_128B layout tcgen05.mma consumes, so
// there is no im2col buffer and no register staging.  W is a [cout, taps*cin] bf16 matrix (tap-major K), tiled
// 64 x BLOCK_N by a 2-D TMA map.  Stride-(1,2,2) convolutions read a 4-phase space-to-depth copy of the input (one
// cheap streaming pre-pass), which turns every tap into a unit-stride box of one phase.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over tiles): warp 0 = TMA producer (one elected lane), warp 1 = MMA
// issuer (one elected lane issues tcgen05.mma kind::f16, M=128, N=BLOCK_N, K=16), warp 2 = TMEM allocator,
// warps 4-11 = epilogue: warp w owns TMEM lanes 32*(w%4).. (32 output voxels) and the 32-column chunks of parity
// (w-4)/4, so two warps per SM sub-partition drain one accumulator (the epilogue is instruction-issue bound).  Pipelines: STAGES-deep smem ring (full/empty
// mbarriers, tcgen05.commit releases slots) and a 2-deep TMEM accumulator ring so the epilogue of tile i overlaps
// the main loop of tile i+1.
#include "ops.cuh"
#include "tc_ptx.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <mutex>
#include <utility>
#include <vector>

namespace echo {

using namespace ptx;

namespace {

constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
constexpr int MAX_BLOCK_N = 256;
constexpr int MAX_STAGES = 12;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;       // 16 KiB
constexpr int SMEM_LIMIT = 227 * 1024;                     // opt-in dynamic shared memory per CTA on sm_100
constexpr int NUM_THREADS = 384;
constexpr int CS_SB = 7;                                   // channels per statistics block (see the epilogue): C % 224 == 0
constexpr int CS_SLOTS = 8;                                // block slots per 32-column chunk (at most 6 are live)
constexpr int SMEM_CS_BYTES = 2 * 4 * (MAX_BLOCK_N / 32) * CS_SLOTS * 8;   // [sub-block parity][row-quarter warp][chunk][slot] float2
constexpr int SMEM_ADD_BYTES = 2 * 2 * MAX_BLOCK_N * 4;    // per-column epilogue addend (bias + rowvec): [tile parity][sub-block][column]
constexpr int SMEM_FIXED = 1024 /*align*/ + 512 /*barriers + tmem ptr*/ + SMEM_CS_BYTES + SMEM_ADD_BYTES;
// bytes in flight per SM is what hides the L2 latency of the TMA stream: use every stage that fits
static inline int stages_for(int b_rows, int msub, bool split = false) {
  const int st = (SMEM_LIMIT - SMEM_FIXED) / ((split ? 2 : 1) * (msub * A_STAGE_BYTES + b_rows * BLOCK_K * 2));
  return st > MAX_STAGES ? MAX_STAGES : st;
}
constexpr int MAX_TAPS = 64;   // 27 taps of a 3x3x3 kernel, or phases x folded taps of a conv after a nearest upsample: 4 x 12 (x(1,2,2)), 8 x 8 (x2)

struct TcParams {
  // problem
  int n_obj, od, oh, ow;        // output grid per object
  int bd, bh, bw;               // tile box (bd*bh*bw == 128)
  int tiles_d, tiles_h, tiles_w;
  int num_m_tiles, num_n_tiles, block_n;
  int up;                       // conv after a nearest upsample, evaluated per output phase on the LOW-resolution input:
                                // 0 none, 1 = x(1,2,2) (4 phases (py,px)), 2 = x(2,2,2) (8 phases (pz,py,px))
  int vm_tiles;                 // schedulable 128-row sub-blocks: num_m_tiles, or 4 x num_m_tiles (phase-major) when up
  int splitk;                   // > 1: the taps are cut into `splitk` equal groups, one CTA tile per group; raw fp32 partial
                                // sums go to out + ks * rows * ldo (splitk_reduce_kernel finishes the epilogue)
  long long total_rows;
  int cin, cout, taps, kblocks_per_tap;
  int obj_mul;                  // 4 for the space-to-depth input (obj index = obj*4 + phase), else 1
  int stages;                   // depth of the smem ring
  int8_t tap_d[MAX_TAPS], tap_h[MAX_TAPS], tap_w[MAX_TAPS], tap_p[MAX_TAPS];
  // epilogue
  const float* bias;
  const float* rowvec;
  long long ld_rowvec;
  const void* res;
  int res_bf16;
  long long ld_res;
  void* out;
  int out_bf16;
  long long ldo;
  int relu;
  float* colsum;   // optional [num_m_tiles][cout/32 chunks][8 slots][2]: (sum, sumsq) of the 7-channel blocks each chunk touches
  int out_t;   // bf16 output stored transposed per object: out[(obj * cout + n) * voxels + voxel] (V^T for the tcgen05 attention)
  unsigned long long* dbg;   // optional timeline of CTA 0 (globaltimer ns): see tc_timeline_*
  int geglu;   // epilogue: tile columns are [a (block_n/2) | g (block_n/2)]; out = (a+ba) * gelu_erf(g+bg), out width cout/2
};

// GELU(g) = 0.5 g (1 + erf(g / sqrt 2)) with erf from Abramowitz & Stegun 7.1.26 (|error| < 1.5e-7 + the ulp-level error
// of ex2.approx / rcp.approx): 2 MUFU + ~12 ALU instructions per element instead of erff's branchy polynomial; the
// GEGLU epilogue is instruction-issue bound (attention.py:39-46 uses the exact erf form, F.gelu default).
__device__ __forceinline__ float gelu_erf_fast(float g) {
  const float x = fabsf(g) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, x, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * x * 1.44269504088896340736f));
  const float erf_abs = fmaf(-poly, e, 1.f);          // erf(|g| / sqrt 2)
  const float erf_s = copysignf(erf_abs, g);
  return 0.5f * g * (1.f + erf_s);
}

// Sum of 8 per-lane values over the 32 lanes: three halving stages (8 -> 4 -> 2 -> 1 values) and two plain ones; on return
// v[0] of lane l is the total of element l >> 2 (9 shuffles).
__device__ __forceinline__ void slot_butterfly(float (&v)[8], int lane) {
#pragma unroll
  for (int off = 16, n = 8; off >= 4; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}



// Where the output rows of one 128-row sub-block live: box coordinates inside the object grid.
// debug timeline: CTA 0 appends (tag, globaltimer) pairs; tags: 1 setup done, 2 TMA issued first stage of a tile, 3 MMA saw the
// first full stage of a tile, 4 MMA committed a tile, 5 epilogue got a tile, 6 epilogue finished a tile, 7 kernel end
__device__ __forceinline__ void dbg_mark(const TcParams& p, int tag, int tile) {
  if (p.dbg && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned long long i = atomicAdd(p.dbg, 1ULL);
    if (i < 4000) { p.dbg[1 + 2 * i] = ((unsigned long long)tag << 32) | (unsigned)tile; p.dbg[2 + 2 * i] = t; }
  }
}

struct SubTile {
  int obj, w0, h0, d0, phase, m_blk;
};
__device__ __forceinline__ SubTile sub_tile(const TcParams& p, int vm) {
  SubTile t;
  t.phase = 0;
  t.m_blk = vm;
  if (p.up) { t.phase = vm / p.num_m_tiles; t.m_blk = vm - t.phase * p.num_m_tiles; }
  int r = t.m_blk;
  const int tw = r % p.tiles_w; r /= p.tiles_w;
  const int th = r % p.tiles_h; r /= p.tiles_h;
  const int td = r % p.tiles_d;
  t.obj = r / p.tiles_d;
  t.w0 = tw * p.bw; t.h0 = th * p.bh; t.d0 = td * p.bd;
  return t;
}

// CTA2 = false: one CTA per (MSUB*128) x block_n tile (cta_group::1).
// CTA2 = true : a cluster of two CTAs per (MSUB*256) x block_n tile (cta_group::2, launched with cluster dims {2,1,1});
//               the pair's leader (rank 0) issues the MMAs, both CTAs run their own TMA producer and epilogue.
// MSUB = 2    : every CTA owns TWO 128-row sub-blocks that share each staged B tile (two accumulators of block_n
//               columns fill TMEM, so the accumulator ring is one deep).  The main loop is bound by the ~56 B/clk an SM
//               can pull from L2; sharing B between two A tiles cuts the bytes per MMA from (16 KiB + B/2) / 1 to
//               (32 KiB + B/2) / 2, which is what makes the long-K convolutions MMA-bound.
// SPLIT = true: split-precision contraction (ECHO_PREC_X3).  Both operands arrive as TWO bf16 tensors, x = hi + lo with
//               hi = bf16(x), lo = bf16(x - hi) (16 mantissa bits together); every k-step issues three MMAs into the SAME fp32
//               TMEM accumulator, A_hi B_hi + A_lo B_hi + A_hi B_lo (the dropped lo x lo term is ~2^-16 relative), so the
//               tensor cores reproduce an fp32 contraction to ~1e-5 -- north_star's 1e-3 parity contract on tcgen05.  A stage
//               holds [A_hi | A_lo | B_hi | B_lo]; MSUB = 1 only.
template <bool CTA2, int MSUB, bool SPLIT>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_b2, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  static_assert(!SPLIT || MSUB == 1, "the split-precision mode runs one sub-block per CTA");
  constexpr int A_BYTES = (SPLIT ? 2 : MSUB) * A_STAGE_BYTES;   // per stage: MSUB sub-block tiles back to back (SPLIT: hi tile, lo tile)
  constexpr int ACC_SLOTS = MSUB == 1 ? 2 : 1;
  const int STAGES = p.stages;
  const int B_ROWS = CTA2 ? (p.block_n >> 1) : p.block_n;   // rows of the B tile staged by THIS CTA
  const int B_HALF_BYTES = B_ROWS * BLOCK_K * 2;
  const int B_STAGE_BYTES = (SPLIT ? 2 : 1) * B_HALF_BYTES;   // SPLIT: hi tile, lo tile
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* bars = (uint64_t*)(smem + STAGES * (A_BYTES + B_STAGE_BYTES));
  uint64_t* full_bar = bars;                   // [STAGES]  (CTA2: the leader's copy is the one in use)
  uint64_t* empty_bar = bars + STAGES;         // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;     // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]       (CTA2: the leader's copy is the one in use)
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * STAGES + 4);
  float2* cs_smem = (float2*)((uint8_t*)bars + 512);
  float* add_smem = (float*)((uint8_t*)bars + 512 + SMEM_CS_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  griddep_launch();               // the next kernel may be scheduled as SMs free up (it waits for our completion itself)
  if (CTA2) cluster_sync_all();   // both CTAs resident before the paired TMEM allocation
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], CTA2 ? 2 : 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], CTA2 ? 512 : 256); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  } else if (warp == 2) {
    if (CTA2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  griddep_wait();                 // everything above overlapped the predecessor's tail; from here on we read its output
  if (threadIdx.x == 0) dbg_mark(p, 1, 0);

  // tile schedule: CTA tiles of MSUB sub-blocks; CTA2 walks (n_blk, m_pair) pairs, this CTA owning m index 2*m_pair + rank
  const int cta_m_tiles = p.vm_tiles / MSUB;
  const int sched_m = CTA2 ? (cta_m_tiles >> 1) : cta_m_tiles;
  const int mn_tiles = sched_m * p.num_n_tiles;
  const int num_tiles = mn_tiles * p.splitk;               // tile = ks * mn_tiles + n_blk * sched_m + mm
  const int taps_per_split = p.taps / p.splitk;
  const int tile0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tstride = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int kblocks = taps_per_split * p.kblocks_per_tap;
  const uint32_t stage_bytes = (uint32_t)A_BYTES + (uint32_t)B_STAGE_BYTES;

  if (warp == 0) {
    // ================= TMA producer =================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tstride) {
        const int ks = tile / mn_tiles, tmn = tile - ks * mn_tiles;
        const int n_blk = tmn / sched_m, mm = tmn - n_blk * sched_m;
        const int m_cta = CTA2 ? 2 * mm + (int)rank : mm;
        SubTile st[MSUB];
#pragma unroll
        for (int j = 0; j < MSUB; ++j) st[j] = sub_tile(p, m_cta * MSUB + j);
        const int n_row0 = n_blk * p.block_n + (CTA2 ? (int)rank * B_ROWS : 0);
        // all sub-blocks of a tile share the phase (host checks); split-K walks its own group of taps
        const int tap_begin = st[0].phase * p.taps + ks * taps_per_split, tap_end = tap_begin + taps_per_split;
        dbg_mark(p, 2, tile);
        for (int tap = tap_begin; tap < tap_end; ++tap) {
          for (int kb = 0; kb < p.kblocks_per_tap; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (CTA2) {
              // both CTAs' copies complete on the LEADER's full barrier (2 arrivals + the bytes of both CTAs)
              if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);
              else mbar_arrive_remote(&full_bar[stage], 0);
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
            }
#pragma unroll
            for (int j = 0; j < MSUB; ++j) {
              uint8_t* dst = smem_a + stage * A_BYTES + j * A_STAGE_BYTES;
              const int cw = st[j].w0 + p.tap_w[tap], chh = st[j].h0 + p.tap_h[tap], cd = st[j].d0 + p.tap_d[tap];
              const int cn = st[j].obj * p.obj_mul + p.tap_p[tap];
              if (CTA2) tma_load_5d_2sm(&map_a, &full_bar[stage], dst, kb * BLOCK_K, cw, chh, cd, cn);
              else tma_load_5d(&map_a, &full_bar[stage], dst, kb * BLOCK_K, cw, chh, cd, cn);
            }
            if (CTA2) tma_load_2d_2sm(&map_b, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, tap * p.cin + kb * BLOCK_K, n_row0);
            else tma_load_2d(&map_b, &full_bar[stage], smem_b + stage * B_STAGE_BYTES, tap * p.cin + kb * BLOCK_K, n_row0);
            if (SPLIT) {   // the low halves of both operands
              uint8_t* dst = smem_a + stage * A_BYTES + A_STAGE_BYTES;
              const int cw = st[0].w0 + p.tap_w[tap], chh = st[0].h0 + p.tap_h[tap], cd = st[0].d0 + p.tap_d[tap];
              const int cn = st[0].obj * p.obj_mul + p.tap_p[tap];
              if (CTA2) {
                tma_load_5d_2sm(&map_a2, &full_bar[stage], dst, kb * BLOCK_K, cw, chh, cd, cn);
                tma_load_2d_2sm(&map_b2, &full_bar[stage], smem_b + stage * B_STAGE_BYTES + B_HALF_BYTES, tap * p.cin + kb * BLOCK_K, n_row0);
              } else {
                tma_load_5d(&map_a2, &full_bar[stage], dst, kb * BLOCK_K, cw, chh, cd, cn);
                tma_load_2d(&map_b2, &full_bar[stage], smem_b + stage * B_STAGE_BYTES + B_HALF_BYTES, tap * p.cin + kb * BLOCK_K, n_row0);
              }
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ================= MMA issuer (the pair's leader CTA only) =================
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N = block_n, M = 128 (or 256 across the CTA pair)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.block_n >> 3) << 17) |
                           ((uint32_t)((CTA2 ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int tile = tile0; tile < num_tiles; tile += tstride, ++iter) {
      const int as = iter % ACC_SLOTS;
      const uint32_t aphase = (iter / ACC_SLOTS) & 1;
      mbar_wait(&tmem_empty[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * MAX_BLOCK_N;
      int kb_total = 0;
      for (int tap = 0; tap < taps_per_split; ++tap) {
        for (int kb = 0; kb < p.kblocks_per_tap; ++kb, ++kb_total) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (kb_total == 0 && lane == 0) dbg_mark(p, 3, tile);
          if (elect_one()) {
            const uint64_t db = make_smem_desc(smem_u32(smem_b + stage * B_STAGE_BYTES));
            int nk = (p.cin - kb * BLOCK_K + UMMA_K - 1) / UMMA_K;   // skip the zero-filled channel tail
            nk = nk > BLOCK_K / UMMA_K ? BLOCK_K / UMMA_K : nk;
            for (int k = 0; k < nk; ++k) {
              const uint64_t koff = (uint64_t)(k * UMMA_K * 2 / 16);
              const uint32_t acc = (kb_total | k) != 0 ? 1u : 0u;
              if (SPLIT) {
                const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * A_BYTES));
                const uint64_t dal = make_smem_desc(smem_u32(smem_a + stage * A_BYTES + A_STAGE_BYTES));
                const uint64_t dbl = make_smem_desc(smem_u32(smem_b + stage * B_STAGE_BYTES + B_HALF_BYTES));
                if (CTA2) {
                  umma_bf16_2sm(tmem_d, dal + koff, db + koff, idesc, acc);     // small terms first
                  umma_bf16_2sm(tmem_d, da + koff, dbl + koff, idesc, 1u);
                  umma_bf16_2sm(tmem_d, da + koff, db + koff, idesc, 1u);
                } else {
                  umma_bf16(tmem_d, dal + koff, db + koff, idesc, acc);
                  umma_bf16(tmem_d, da + koff, dbl + koff, idesc, 1u);
                  umma_bf16(tmem_d, da + koff, db + koff, idesc, 1u);
                }
              } else {
#pragma unroll
              for (int j = 0; j < MSUB; ++j) {
                const uint64_t da = make_smem_desc(smem_u32(smem_a + stage * A_BYTES + j * A_STAGE_BYTES));
                if (CTA2) umma_bf16_2sm(tmem_d + j * MAX_BLOCK_N, da + koff, db + koff, idesc, acc);
                else umma_bf16(tmem_d + j * MAX_BLOCK_N, da + koff, db + koff, idesc, acc);
              }
              }
            }
            if (CTA2) {
              umma_commit_2sm(&empty_bar[stage]);                          // frees the slot in BOTH CTAs
              if (kb_total == kblocks - 1) umma_commit_2sm(&tmem_full[as]);   // both epilogues
            } else {
              umma_commit(&empty_bar[stage]);                         // frees the smem slot when these MMAs retire
              if (kb_total == kblocks - 1) umma_commit(&tmem_full[as]);   // accumulator complete -> epilogue
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (lane == 0) dbg_mark(p, 4, tile);
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> global =================
    const int ew = (warp - 4) & 3;                 // TMEM lanes [32*ew, 32*ew+32)
    const int cg = (warp - 4) >> 2;                // column group: 32-wide chunks with (chunk index & 1) == cg
    const int row = ew * 32 + lane;                // tile row = box voxel index
    int cs_par = 0;
    const int ww = row % p.bw, hh = (row / p.bw) % p.bh, dd = row / (p.bw * p.bh);
    int iter = 0;
    for (int tile = tile0; tile < num_tiles; tile += tstride, ++iter) {
      const int as = iter % ACC_SLOTS;
      const uint32_t aphase = (iter / ACC_SLOTS) & 1;
      const int ks = tile / mn_tiles, tmn = tile - ks * mn_tiles;
      const int n_blk = tmn / sched_m, mm = tmn - n_blk * sched_m;
      const int m_cta = CTA2 ? 2 * mm + (int)rank : mm;
      // per-column addend (bias + this object's rowvec) of each sub-block, staged in shared memory while the main loop
      // of this tile is still running: the epilogue proper then has no global round trip per chunk besides the residual
      float* addv = add_smem + (iter & 1) * 2 * MAX_BLOCK_N;
      const bool use_add = !p.geglu && (p.bias || p.rowvec);
      if (use_add) {
        const int t = threadIdx.x - 128;
#pragma unroll
        for (int sub = 0; sub < MSUB; ++sub) {
          if (t < p.block_n) {
            const int n = n_blk * p.block_n + t;
            float a = 0.f;
            if (n < p.cout) {
              if (p.bias) a = __ldg(p.bias + n);
              if (p.rowvec) a += __ldg(p.rowvec + (long long)sub_tile(p, m_cta * MSUB + sub).obj * p.ld_rowvec + n);
            }
            addv[sub * MAX_BLOCK_N + t] = a;
          }
        }
        asm volatile("bar.sync 3, 256;" ::: "memory");   // the eight epilogue warps
      }
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      if (threadIdx.x == 128) dbg_mark(p, 5, tile);
#pragma unroll 1
      for (int sub = 0; sub < MSUB; ++sub) {
      const SubTile stl = sub_tile(p, m_cta * MSUB + sub);
      const int obj = stl.obj;
      const int ow_ = stl.w0 + ww, oh_ = stl.h0 + hh, od_ = stl.d0 + dd;
      const bool valid = ow_ < p.ow && oh_ < p.oh && od_ < p.od;
      const int upd = p.up == 2 ? 1 : 0;   // depth doubled as well
      const long long orow = (p.up ? ((((long long)obj * (p.od << upd) + ((od_ << upd) + (upd ? (stl.phase >> 2) : 0))) * (2 * p.oh) + 2 * oh_ +
                                       ((stl.phase >> 1) & 1)) * (2 * p.ow) + 2 * ow_ + (stl.phase & 1))
                                   : (((long long)obj * p.od + od_) * p.oh + oh_) * p.ow + ow_) + ks * p.total_rows;
      const long long cs_row = p.up ? (long long)stl.m_blk * (p.up == 2 ? 8 : 4) + stl.phase : stl.m_blk;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (as + sub) * MAX_BLOCK_N;
      const int n_base = n_blk * p.block_n;
      if (p.geglu) {
        // GEGLU fused into the producing GEMM (attention.py:39-46): the weight rows were permuted at load time so that
        // one tile holds 128 `a` columns followed by their 128 gate columns
        // This epilogue is what bounds the ff1 contraction (K = C is short): it runs in 16-column pieces, the TMEM read of
        // piece i+1 in flight while piece i is evaluated, biases fetched as float4.  Pieces of this warp: columns
        // 32*cg + 64*k + {0, 16} of the `a` half (block_n = 256 -> four pieces).
        const int half = p.block_n >> 1;   // 128
        uint32_t va[2][16], vg[2][16];
        tmem_ld16(taddr + 32 * cg, va[0]);
        tmem_ld16(taddr + half + 32 * cg, vg[0]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = 32 * cg + 64 * (i >> 1) + 16 * (i & 1);
          tmem_ld_wait();
          if (i + 1 < 4) {
            const int cn = 32 * cg + 64 * ((i + 1) >> 1) + 16 * ((i + 1) & 1);
            tmem_ld16(taddr + cn, va[(i + 1) & 1]);
            tmem_ld16(taddr + half + cn, vg[(i + 1) & 1]);
          }
          if (valid) {
            const float4* ba = reinterpret_cast<const float4*>(p.bias + n_base + c);
            const float4* bg = reinterpret_cast<const float4*>(p.bias + n_base + half + c);
            uint32_t w8[8];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 b1 = __ldg(ba + q), b2 = __ldg(bg + q);
              const float a0 = __uint_as_float(va[i & 1][4 * q]) + b1.x, a1 = __uint_as_float(va[i & 1][4 * q + 1]) + b1.y;
              const float a2 = __uint_as_float(va[i & 1][4 * q + 2]) + b1.z, a3 = __uint_as_float(va[i & 1][4 * q + 3]) + b1.w;
              const float g0 = __uint_as_float(vg[i & 1][4 * q]) + b2.x, g1 = __uint_as_float(vg[i & 1][4 * q + 1]) + b2.y;
              const float g2 = __uint_as_float(vg[i & 1][4 * q + 2]) + b2.z, g3 = __uint_as_float(vg[i & 1][4 * q + 3]) + b2.w;
              __nv_bfloat162 h0 = __floats2bfloat162_rn(a0 * gelu_erf_fast(g0), a1 * gelu_erf_fast(g1));
              __nv_bfloat162 h1 = __floats2bfloat162_rn(a2 * gelu_erf_fast(g2), a3 * gelu_erf_fast(g3));
              w8[2 * q] = *reinterpret_cast<uint32_t*>(&h0);
              w8[2 * q + 1] = *reinterpret_cast<uint32_t*>(&h1);
            }
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + n_blk * half + c);
            op[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
            op[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
          }
        }
      } else {
        // 32-column chunks c = 32*cg + 64*k (k < 4).  Software pipeline: the residual rows of chunk k+1 are requested
        // before chunk k is evaluated (an L2 round trip each), the per-column addends come from shared memory.
        const bool res16 = p.res && p.res_bf16;
        const __nv_bfloat16* res_row = res16 ? reinterpret_cast<const __nv_bfloat16*>(p.res) + orow * p.ld_res + n_base : nullptr;
        float2* csb = cs_smem + (cs_par * 4 + ew) * (MAX_BLOCK_N / 32) * CS_SLOTS;
        uint4 rres[2][4];
        const int c0 = 32 * cg;
        if (c0 < p.block_n) {
          if (res16 && valid && n_base + c0 < p.cout) {
#pragma unroll
            for (int q = 0; q < 4; ++q) rres[0][q] = __ldg(reinterpret_cast<const uint4*>(res_row + c0) + q);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c = 32 * cg + 64 * k;
          if (c >= p.block_n) break;
          const int n0 = n_base + c;
          uint32_t v[32];
          tmem_ld32(taddr + c, v);
          const int cn = c + 64;
          if (k + 1 < 4 && cn < p.block_n) {
            if (res16 && valid && n_base + cn < p.cout) {
#pragma unroll
              for (int q = 0; q < 4; ++q) rres[(k + 1) & 1][q] = __ldg(reinterpret_cast<const uint4*>(res_row + cn) + q);
            }
          }
          tmem_ld_wait();
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = 0.f;
          if (valid && n0 < p.cout) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (use_add) {
              const float4* ap = reinterpret_cast<const float4*>(addv + sub * MAX_BLOCK_N + c);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = ap[j >> 2];
                f[j] += b.x; f[j + 1] += b.y; f[j + 2] += b.z; f[j + 3] += b.w;
              }
            }
            if (p.res) {
              if (p.res_bf16) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const uint4 u = rres[k & 1][q];
                  const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[e]);
                    f[q * 8 + e * 2] += __low2float(h2);
                    f[q * 8 + e * 2 + 1] += __high2float(h2);
                  }
                }
              } else {
                const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res) + orow * p.ld_res + n0);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const float4 u = __ldg(rp + q);
                  f[q * 4] += u.x; f[q * 4 + 1] += u.y; f[q * 4 + 2] += u.z; f[q * 4 + 3] += u.w;
                }
              }
            }
            if (p.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (p.out_t) {
              // lanes are consecutive voxels of one object: each column is a 64-byte run of the transposed tensor
              const long long vox = (long long)p.od * p.oh * p.ow;
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + ((long long)obj * p.cout + n0) * vox + (orow - (long long)obj * vox);
#pragma unroll
              for (int j = 0; j < 32; ++j) op[(long long)j * vox] = __float2bfloat16(f[j]);
            } else if (p.out_bf16) {
              uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + n0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint32_t w4[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(f[q * 8 + e * 2], f[q * 8 + e * 2 + 1]);
                  w4[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                op[q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              }
            } else {
              float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + n0);
#pragma unroll
              for (int q = 0; q < 8; ++q) op[q] = make_float4(f[q * 4], f[q * 4 + 1], f[q * 4 + 2], f[q * 4 + 3]);
            }
          }
          if (p.colsum) {
            // GroupNorm statistics of the NEXT layer, gathered here instead of in a pass over the activation.  Every
            // GroupNorm group of this network is a run of whole 7-channel blocks (channel counts are multiples of 224 =
            // 32 groups x 7), so each thread first folds its row's 32 values into the <= 6 blocks this chunk touches
            // (block b0 + s <-> slot s, b0 = 32q / 7; `off` = position of the chunk's first column inside its block) and
            // only those slot sums cross lanes: 18 shuffles per chunk instead of 62.
            const int off = n0 % CS_SB;
            float tot[5], lo[5], qtot[5], qlo[5];
#pragma unroll
            for (int a = 0; a < 5; ++a) tot[a] = lo[a] = qtot[a] = qlo[a] = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int a = j / CS_SB, r = j % CS_SB;
              tot[a] += f[j];
              qtot[a] = fmaf(f[j], f[j], qtot[a]);
              if (r + off < CS_SB) {   // still inside slot a; otherwise it already belongs to slot a + 1
                lo[a] += f[j];
                qlo[a] = fmaf(f[j], f[j], qlo[a]);
              }
            }
            float ss[8], sq[8];
            ss[0] = lo[0]; sq[0] = qlo[0];
#pragma unroll
            for (int k = 1; k < 5; ++k) { ss[k] = (tot[k - 1] - lo[k - 1]) + lo[k]; sq[k] = (qtot[k - 1] - qlo[k - 1]) + qlo[k]; }
            ss[5] = tot[4] - lo[4]; sq[5] = qtot[4] - qlo[4];
            ss[6] = ss[7] = sq[6] = sq[7] = 0.f;
            slot_butterfly(ss, lane);
            slot_butterfly(sq, lane);
            if ((lane & 3) == 0) csb[(c >> 5) * CS_SLOTS + (lane >> 2)] = make_float2(ss[0], sq[0]);
          }
        }
        if (p.colsum) {
          // -> partial row of this sub-block, [chunk][slot] (sum, sumsq).  One barrier per sub-block for the four warps of
          // this column group, fixed summation order (run-to-run reproducible); buffers alternate between sub-blocks so
          // nobody waits for the readers.
          asm volatile("bar.sync %0, 128;" ::"r"(1 + cg) : "memory");
          const int t = ew * 32 + lane;                              // 0..127 within the column group
          if (t < 4 * CS_SLOTS) {                                    // (chunk k of this group, slot)
            const int c = 32 * cg + 64 * (t / CS_SLOTS), slot = t % CS_SLOTS;
            if (c < p.block_n && n_base + c < p.cout) {
              constexpr int WS = (MAX_BLOCK_N / 32) * CS_SLOTS;        // one warp's region
              const float2* b0 = cs_smem + (cs_par * 4) * WS + (c >> 5) * CS_SLOTS + slot;
              const float2 a0 = b0[0], a1 = b0[WS], a2 = b0[2 * WS], a3 = b0[3 * WS];
              const long long nchunks = p.cout >> 5;
              *reinterpret_cast<float2*>(p.colsum + ((cs_row * nchunks + ((n_base + c) >> 5)) * CS_SLOTS + slot) * 2) =
                  make_float2((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
            }
          }
          cs_par ^= 1;
        }
      }
      }   // sub
      if (threadIdx.x == 128) dbg_mark(p, 6, tile);
      tc_fence_before();
      if (CTA2 && !leader) mbar_arrive_remote(&tmem_empty[as], 0);   // the leader's MMA warp owns the accumulator ring
      else mbar_arrive(&tmem_empty[as]);
    }
  }

  if (threadIdx.x == 0) dbg_mark(p, 7, 0);
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();   // (pair) nobody still signals a barrier or reads TMEM of an exited CTA
  if (warp == 2) {
    tc_fence_after();
    if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// 4-phase space-to-depth of a channels-last tensor: out[(obj*4 + ph*2 + pw), d, h/2, w/2, c] = x[obj, d, 2h'+ph, 2w'+pw, c]
__global__ void s2d_kernel(const __nv_bfloat16* __restrict__ x, int n, int d, int h, int w, int C, long long nvec, __nv_bfloat16* __restrict__ out) {
  griddep_launch();
  griddep_wait();
  const int cv = C / 8, h2 = h / 2, w2 = w / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv); long long r = i / cv;
    const int xw = (int)(r % w); r /= w;
    const int xh = (int)(r % h); r /= h;
    const int xd = (int)(r % d); const long long obj = r / d;
    const int ph = xh & 1, pw = xw & 1;
    const long long dst = ((((obj * 4 + ph * 2 + pw) * d + xd) * h2 + (xh >> 1)) * w2 + (xw >> 1)) * (long long)C + c8 * 8;
    *reinterpret_cast<uint4*>(out + dst) = __ldg(reinterpret_cast<const uint4*>(x + i * 8));
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcState {
  bool checked = false, ok = false;
  EncodeTiledFn encode = nullptr;
  int sms = 148;
};
TcState g_tc;
int g_tc_mode = 0;   // 0 = automatic, 1 = one CTA per tile, 2 = CTA pairs where possible (echo_debug_set_tc_mode)
std::mutex g_tc_mu;

void tc_init() {
  std::lock_guard<std::mutex> lk(g_tc_mu);
  if (g_tc.checked) return;
  g_tc.checked = true;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { cudaGetLastError(); return; }
  if (prop.major != 10) return;   // tcgen05 / TMEM: sm_100a family only
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    cudaGetLastError();
    return;
  }
  g_tc.encode = (EncodeTiledFn)fn;
  g_tc.sms = prop.multiProcessorCount;
  if (cudaFuncSetAttribute(gemm_tc_kernel<false, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tc_kernel<true, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tc_kernel<false, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tc_kernel<true, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tc_kernel<false, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess ||
      cudaFuncSetAttribute(gemm_tc_kernel<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) {
    cudaGetLastError();
    return;
  }
  g_tc.ok = true;
}

int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

// ---- launch plan ------------------------------------------------------------------------------------------------
// block_n : tile width, multiples of 32 up to 256 (a partial last n-tile is fine: TMA zero-fills weight rows >= cout, the
//           epilogue masks whole 32-column chunks)
// msub    : 128-row sub-blocks per CTA (2 = share every staged B tile between two A tiles, one-deep accumulator ring)
// splitk  : tap groups of a 3x3x3 conv run as separate CTA tiles writing fp32 partials (finished by
//           splitk_reduce_kernel) -- for the coarse levels whose few output tiles would leave most SMs idle
// chosen by a small cycle model: a persistent grid runs `waves` rounds of one tile per SM; a k-block stage costs
// max(MMA time, bytes / ~52 B/clk an SM pulls from L2 through TMA); the epilogue overlaps the next tile only when the
// accumulator ring is two deep (msub == 1).
struct TcPlan {
  int block_n = 0, msub = 1, splitk = 1;
  bool cta2 = false;
};

struct TcGeom {
  int bw, bh, bd, tiles_w, tiles_h, tiles_d, num_m_tiles, vm_tiles, taps, kblocks_per_tap;
};

TcGeom tc_geom(const GemmArgs& g) {
  TcGeom t;
  const bool up = g.up2 != 0;
  const int gw = up ? g.w : g.ow, gh = up ? g.h : g.oh, gd = g.up2 == 2 ? g.d : g.od;
  t.bw = pow2_floor(gw) > 128 ? 128 : pow2_floor(gw);
  t.bh = pow2_floor(gh) > 128 / t.bw ? 128 / t.bw : pow2_floor(gh);
  t.bd = 128 / (t.bw * t.bh);
  t.tiles_w = cdiv(gw, t.bw); t.tiles_h = cdiv(gh, t.bh); t.tiles_d = cdiv(gd, t.bd);
  t.num_m_tiles = g.n * t.tiles_d * t.tiles_h * t.tiles_w;
  t.vm_tiles = t.num_m_tiles * (g.up2 == 2 ? 8 : up ? 4 : 1);
  t.taps = g.up2 == 2 ? 8 : up ? 12 : g.kd * g.kh * g.kw;
  t.kblocks_per_tap = cdiv(g.cin, BLOCK_K);
  return t;
}

TcPlan tc_plan(const GemmArgs& g, const TcGeom& t, int sms) {
  static const int mode_env = getenv("ECHO_TC_MODE") ? atoi(getenv("ECHO_TC_MODE")) : 0;   // 1 / 2 force a mode (tests, profiling)
  static const int msub_env = getenv("ECHO_TC_MSUB") ? atoi(getenv("ECHO_TC_MSUB")) : 0;
  static const int split_env = getenv("ECHO_TC_SPLITK") ? atoi(getenv("ECHO_TC_SPLITK")) : 0;   // 1 = never split
  const int mode = g_tc_mode ? g_tc_mode : mode_env;
  TcPlan best;
  double best_cost = 0;
  const bool geglu = g.epi == 1;
  const bool x3 = g.A_lo != nullptr;
  const bool can_split = !x3 && g.splitk_ws && !geglu && !g.up2 && t.taps == 27 && g.sh == 1 && ((int64_t)g.od * g.oh * g.ow) % 128 == 0 &&
                         g.out_dt == BF16 && split_env != 1;
  for (int bn = 256; bn >= 32; bn -= 32) {
    if (geglu && bn != 256) continue;
    if (bn > g.cout && bn != 32 && bn - 32 >= g.cout) continue;   // never wider than needed
    const int n_tiles = cdiv(g.cout, bn);
    const bool cta2 = mode != 1 && t.num_m_tiles % 2 == 0 && bn % 32 == 0;
    for (int msub = 1; msub <= 2; ++msub) {
      if (msub == 2 && (x3 || geglu || t.num_m_tiles % (cta2 ? 4 : 2) != 0)) continue;
      if (!x3 && msub_env && msub != msub_env && !(msub_env == 2 && (geglu || t.num_m_tiles % (cta2 ? 4 : 2) != 0))) continue;
      const int b_rows = cta2 ? bn / 2 : bn;
      if (stages_for(b_rows, msub, x3) < (x3 ? 2 : 3)) continue;
      for (int sk = 1; sk <= (can_split ? 3 : 1); sk += 2) {
        const long long tiles = (long long)(t.vm_tiles / msub) * n_tiles * sk;
        const long long waves = (tiles + sms - 1) / sms;
        const double mma = (x3 ? 3.0 : 1.0) * msub * 4.0 * (bn / 2.0);
        const double feed = (x3 ? 2.0 : 1.0) * (msub * 16384.0 + b_rows * 128.0) / 52.0;
        const double stage = mma > feed ? mma : feed;
        const double mainloop = (double)t.taps * t.kblocks_per_tap / sk * stage;
        const double epi = msub * (1500.0 + 12.0 * bn);
        double cost = msub == 2 ? waves * (mainloop + epi) : waves * mainloop + epi;
        cost += 2500.0;                                         // launch + prologue
        if (sk > 1) cost += 6000.0 + 2500.0;                    // the reduce pass: a second (small) kernel
        if (!best.block_n || cost < best_cost) {
          best.block_n = bn; best.msub = msub; best.splitk = sk; best.cta2 = cta2;
          best_cost = cost;
        }
      }
    }
  }
  return best;
}

// Second half of a split-K contraction: out = sum_ks partial[ks] + bias + rowvec[obj] + res, rounded to bf16, plus the
// per-(128-row group, column) GroupNorm partials the single-pass epilogue would have written.  One block = 128 rows x
// 64 columns; thread = (row lane 0..31, column octet 0..7); rows are consecutive, so a group never straddles two objects
// (voxels % 128 == 0 is required by the plan).
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, int splitk, long long rows, int cout,
                                                            const float* __restrict__ bias, const float* __restrict__ rowvec,
                                                            long long ld_rowvec, long long rows_per_obj,
                                                            const __nv_bfloat16* __restrict__ res, long long ld_res, int relu,
                                                            __nv_bfloat16* __restrict__ out, long long ldo, float* __restrict__ colsum) {
  __shared__ float2 red[32][64];
  griddep_launch();
  griddep_wait();
  const int c8 = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const int col = blockIdx.y * 64 + c8 * 8;
  const long long row0 = (long long)blockIdx.x * 128;
  const bool col_ok = col < cout;
  float sum[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
  float add[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) add[j] = 0.f;
  if (col_ok && bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + col)), b1 = __ldg(reinterpret_cast<const float4*>(bias + col + 4));
    add[0] = b0.x; add[1] = b0.y; add[2] = b0.z; add[3] = b0.w; add[4] = b1.x; add[5] = b1.y; add[6] = b1.z; add[7] = b1.w;
  }
  if (col_ok && rowvec) {
    const float* rv = rowvec + (row0 / rows_per_obj) * ld_rowvec + col;
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(rv)), b1 = __ldg(reinterpret_cast<const float4*>(rv + 4));
    add[0] += b0.x; add[1] += b0.y; add[2] += b0.z; add[3] += b0.w; add[4] += b1.x; add[5] += b1.y; add[6] += b1.z; add[7] += b1.w;
  }
  if (col_ok) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long r = row0 + rl + 32 * i;
      if (r >= rows) break;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = add[j];
      for (int ks = 0; ks < splitk; ++ks) {
        const float* wp = ws + ((long long)ks * rows + r) * cout + col;
        const float4 a0 = *reinterpret_cast<const float4*>(wp), a1 = *reinterpret_cast<const float4*>(wp + 4);
        f[0] += a0.x; f[1] += a0.y; f[2] += a0.z; f[3] += a0.w; f[4] += a1.x; f[5] += a1.y; f[6] += a1.z; f[7] += a1.w;
      }
      if (res) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(res + r * ld_res + col));
        const uint32_t w4[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h2 = *reinterpret_cast<const __nv_bfloat162*>(&w4[e]);
          f[2 * e] += __low2float(h2);
          f[2 * e + 1] += __high2float(h2);
        }
      }
      if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], 0.f);
      }
      uint32_t w4[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
        w4[e] = *reinterpret_cast<uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(out + r * ldo + col) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { sum[j] += f[j]; sq[j] = fmaf(f[j], f[j], sq[j]); }
    }
  }
  if (!colsum) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[rl][c8 * 8 + j] = make_float2(sum[j], sq[j]);
  __syncthreads();
  __shared__ float2 colt[64];
  if (threadIdx.x < 64) {
    float a = 0.f, b = 0.f;
    if (blockIdx.y * 64 + threadIdx.x < cout)
      for (int r = 0; r < 32; ++r) { a += red[r][threadIdx.x].x; b += red[r][threadIdx.x].y; }
    colt[threadIdx.x] = make_float2(a, b);
  }
  __syncthreads();
  // same partial-row format as the single-pass epilogue: [row tile][32-column chunk][slot] = sums of the 7-channel blocks
  if (threadIdx.x < 2 * CS_SLOTS) {
    const int ql = threadIdx.x / CS_SLOTS, slot = threadIdx.x % CS_SLOTS;
    const int q = blockIdx.y * 2 + ql, n0 = 32 * q;
    if (n0 < cout) {
      const int b = n0 / CS_SB + slot;
      const int lo = max(n0, b * CS_SB), hi = min(min(n0 + 32, (b + 1) * CS_SB), cout);
      float a = 0.f, bb = 0.f;
      for (int c = lo; c < hi; ++c) { a += colt[c - blockIdx.y * 64].x; bb += colt[c - blockIdx.y * 64].y; }
      *reinterpret_cast<float2*>(colsum + (((long long)blockIdx.x * (cout >> 5) + q) * CS_SLOTS + slot) * 2) = make_float2(a, bb);
    }
  }
}

}  // namespace

void set_tc_mode(int m) { g_tc_mode = m; }

// ---- in-situ timing probe (bench.py's roofline): CUDA events on the launching stream around every launch of ONE
//      contraction shape while a real step runs, so the kernel is timed warm, between its actual neighbours ----
namespace {
struct TcProbe {
  unsigned long long* timeline = nullptr;   // device buffer handed to the NEXT probed launch (then cleared)
  bool on = false;
  long long rows = 0;
  int cin = 0, cout = 0, k = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t used = 0;
} g_probe;
}  // namespace

void tc_probe_begin(long long rows, int cin, int cout, int k) {
  g_probe.on = true;
  g_probe.rows = rows; g_probe.cin = cin; g_probe.cout = cout; g_probe.k = k;
  g_probe.used = 0;
}

// the next probed launch records the timeline of its CTA 0 into `buf` (device, >= 8001 u64, zeroed by the caller)
void tc_probe_timeline(unsigned long long* buf) { g_probe.timeline = buf; }

// average milliseconds per probed launch and the number of launches seen (call after the stream has been synchronised)
int tc_probe_end(double* avg_ms) {
  g_probe.on = false;
  double tot = 0.0;
  for (size_t i = 0; i < g_probe.used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_probe.ev[i].first, g_probe.ev[i].second) == cudaSuccess) tot += ms;
  }
  const int n = (int)g_probe.used;
  if (avg_ms) *avg_ms = n ? tot / n : 0.0;
  g_probe.used = 0;
  return n;
}

// 2-D bf16 tensor map, SWIZZLE_128B (for the other tcgen05 kernels of the library: flash_tc.cu)
void tc_encode_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner,
                       uint32_t box_outer) {
  ECHO_CHECK(tc_available(), "tcgen05 path unavailable on this device");
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_tc.encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ECHO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: %d", (int)r);
}

bool tc_available() {
  tc_init();
  return g_tc.ok;
}

int gemm_tc_colsum_rows_per_obj(const GemmArgs& g) {
  GemmArgs t = g;
  t.n = 1;
  if (t.up2) { t.h = t.oh / 2; t.w = t.ow / 2; }
  if (t.up2 == 2) t.d = t.od / 2;
  const TcGeom ge = tc_geom(t);
  return ge.vm_tiles;
}

// host-only view of the launch plan for a (n, d, h, w, cin) -> cout contraction on `sms` SMs (tests/test_plan_host.py)
void tc_plan_describe(int n, int d, int h, int w, int cin, int cout, int ksize, int epi, int up2, int allow_splitk, int sms, int* out4) {
  GemmArgs g;
  g.a_dt = BF16; g.w_dt = BF16; g.out_dt = BF16;
  g.n = n; g.d = d; g.h = h; g.w = w; g.cin = cin; g.lda = cin; g.cout = cout;
  g.kd = g.kh = g.kw = ksize; g.pd = g.ph = g.pw = ksize / 2;
  g.up2 = up2; g.epi = epi;
  g.od = up2 == 2 ? 2 * d : d; g.oh = up2 ? 2 * h : h; g.ow = up2 ? 2 * w : w;
  static float dummy;
  g.splitk_ws = allow_splitk ? (void*)&dummy : nullptr;
  const TcGeom ge = tc_geom(g);
  const TcPlan p = tc_plan(g, ge, sms);
  out4[0] = p.block_n; out4[1] = p.msub; out4[2] = p.splitk; out4[3] = p.cta2 ? 1 : 0;
}

// floats of one partial row ([chunk][slot][2]) for `c` channels; 0 if c does not decompose into 7-channel blocks per group
size_t gemm_tc_colsum_row_floats(int c) { return c % (32 * CS_SB) == 0 ? (size_t)(c / 32) * CS_SLOTS * 2 : 0; }

size_t gemm_tc_splitk_ws_bytes(const GemmArgs& g) {
  if (!tc_available() || g.kd != 3 || g.sh != 1 || g.up2 || g.epi) return 0;
  // only problems that cannot fill two waves of 128 x 224 tiles are worth a second pass over fp32 partials; when a
  // workspace is offered the plan decides: size it for the deepest split
  if ((long long)cdiv(g.rows_out(), 128) * cdiv(g.cout, 224) >= 2LL * g_tc.sms) return 0;
  return (size_t)3 * g.rows_out() * g.cout * sizeof(float);
}

bool gemm_tc_supported(const GemmArgs& g) {
  if (g.a_dt != BF16 || g.w_dt != BF16) return false;
  if ((g.A_lo != nullptr) != (g.W_lo != nullptr)) return false;
  if (g.A_lo) {   // split-precision mode: plain convolutions / linears with the standard epilogue
    if (g.epi || g.out_t || g.up2 || g.colsum) return false;
    if (((uintptr_t)g.A_lo % 16) || ((uintptr_t)g.W_lo % 16)) return false;
  }
  if (g.nb0 * g.nb1 != 1 || g.alpha != 1.f || g.act > 1) return false;
  if (g.up2) {   // W = [cout][4 phases][12 folded taps][cin] or [cout][8][8][cin] (fold_upsample_weight)
    if (g.up2 != 1 && g.up2 != 2) return false;
    if (g.kd != 3 || g.kh != 3 || g.kw != 3 || g.sd != 1 || g.sh != 1 || g.sw != 1 || g.oh != 2 * g.h || g.ow != 2 * g.w) return false;
    if (g.od != (g.up2 == 2 ? 2 * g.d : g.d)) return false;
    if (g.cin % 16 != 0 || g.lda != g.cin || g.w_stride_k != 1 || g.w_stride_n != (int64_t)(g.up2 == 2 ? 64 : 48) * g.cin || g.epi || g.res) return false;
  } else
  if (g.cin % 16 != 0 || g.lda != g.cin || g.w_stride_k != 1 || g.w_stride_n != (int64_t)g.ktot()) return false;
  if (g.cout % 32 != 0) return false;
  if (g.out_t && (g.out_dt != BF16 || g.epi || g.up2 || g.splitk_ws || g.colsum)) return false;
  if (g.splitk_ws && ((uintptr_t)g.splitk_ws % 16)) return false;
  if (g.epi == 1 && g.colsum) return false;
  if (g.epi == 1 && (g.cout % 256 != 0 || g.res || g.rowvec || !g.bias || g.out_dt != BF16 || g.act != 0)) return false;
  if (!((g.kd == 1 && g.kh == 1 && g.kw == 1) || (g.kd == 3 && g.kh == 3 && g.kw == 3))) return false;
  if (g.pd != g.kd / 2 || g.ph != g.kh / 2 || g.pw != g.kw / 2 || g.sd != 1 || g.sh != g.sw) return false;
  if (g.sh == 2) {
    if (g.kd != 3 || (g.h & 1) || (g.w & 1) || !g.scratch) return false;
    if (g.oh != g.h / 2 || g.ow != g.w / 2) return false;
  } else if (g.sh == 1) {
    if (!g.up2 && (g.od != g.d || g.oh != g.h || g.ow != g.w)) return false;
  } else {
    return false;
  }
  if (((uintptr_t)g.A % 16) || ((uintptr_t)g.W % 16) || ((uintptr_t)g.out % 16)) return false;
  if (g.res && (((uintptr_t)g.res % 16) || g.ld_res % 8)) return false;
  if (g.ldo % 8) return false;
  if (g.rowvec && (g.ld_rowvec % 4 || ((uintptr_t)g.rowvec % 16))) return false;
  if (g.bias && ((uintptr_t)g.bias % 16)) return false;
  return true;
}

void gemm_tc(const GemmArgs& g, cudaStream_t s) {
  ECHO_CHECK(tc_available(), "gemm_tc: tcgen05 path unavailable on this device");
  ECHO_CHECK(gemm_tc_supported(g), "gemm_tc: unsupported problem");
  if (g.rows_out() == 0) return;
  if (dbg_skip("gemm_tc")) return;
  const bool s2 = g.sh == 2;
  const bool x3 = g.A_lo != nullptr;
  const __nv_bfloat16* a_ptr = (const __nv_bfloat16*)g.A;
  const __nv_bfloat16* a_lo_ptr = (const __nv_bfloat16*)g.A_lo;
  int in_h = g.h, in_w = g.w, in_objs = g.n;
  if (s2) {
    const long long nvec = (long long)g.n * g.d * g.h * g.w * (g.cin / 8);
    long long blocks = (nvec + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    launch_pdl(s2d_kernel, dim3((int)blocks), dim3(256), 0, s, a_ptr, g.n, g.d, g.h, g.w, g.cin, nvec, (__nv_bfloat16*)g.scratch);
    ECHO_LAUNCH_CHECK();
    a_ptr = (const __nv_bfloat16*)g.scratch;
    if (x3) {   // the scratch holds both halves back to back (the caller sizes it for two copies)
      __nv_bfloat16* lo_dst = (__nv_bfloat16*)g.scratch + nvec * 8;
      launch_pdl(s2d_kernel, dim3((int)blocks), dim3(256), 0, s, a_lo_ptr, g.n, g.d, g.h, g.w, g.cin, nvec, lo_dst);
      ECHO_LAUNCH_CHECK();
      a_lo_ptr = lo_dst;
    }
    in_h = g.h / 2;
    in_w = g.w / 2;
    in_objs = g.n * 4;
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  const bool up = g.up2 != 0;
  const TcGeom ge = tc_geom(g);
  const TcPlan plan = tc_plan(g, ge, g_tc.sms);
  ECHO_CHECK(plan.block_n > 0, "gemm_tc: no launch plan");
  if (dbg_trace())
    fprintf(stderr, "[echo-trace] gemm_tc rows=%lld cin=%d cout=%d k=%d stride=%d epi=%d up=%d bn=%d msub=%d splitk=%d cta2=%d x3=%d\n",
            (long long)g.rows_out(), g.cin, g.cout, g.kd, g.sh, g.epi, g.up2, plan.block_n, plan.msub, plan.splitk, plan.cta2 ? 1 : 0, x3 ? 1 : 0);
  p.up = g.up2;
  p.n_obj = g.n; p.od = g.up2 == 2 ? g.d : g.od; p.oh = up ? g.h : g.oh; p.ow = up ? g.w : g.ow;   // the grid the 128-voxel boxes tile
  p.bw = ge.bw; p.bh = ge.bh; p.bd = ge.bd;
  p.tiles_w = ge.tiles_w; p.tiles_h = ge.tiles_h; p.tiles_d = ge.tiles_d;
  p.num_m_tiles = ge.num_m_tiles;
  p.vm_tiles = ge.vm_tiles;
  p.block_n = plan.block_n;
  p.num_n_tiles = cdiv(g.cout, p.block_n);
  p.cin = g.cin; p.cout = g.cout; p.taps = ge.taps;
  p.kblocks_per_tap = ge.kblocks_per_tap;
  p.splitk = plan.splitk;
  p.total_rows = g.rows_out();
  p.obj_mul = s2 ? 4 : 1;
  if (g.up2 == 2) {
    // output voxel (2z+pz, 2y+py, 2x+px) reads low-res {z+pz-1, z+pz} x {y+py-1, y+py} x {x+px-1, x+px}: tap (a_d, a_h, a_w)
    for (int ph = 0; ph < 8; ++ph)
      for (int t = 0; t < 8; ++t) {
        p.tap_d[ph * 8 + t] = (int8_t)(((ph >> 2) & 1) - 1 + ((t >> 2) & 1));
        p.tap_h[ph * 8 + t] = (int8_t)(((ph >> 1) & 1) - 1 + ((t >> 1) & 1));
        p.tap_w[ph * 8 + t] = (int8_t)((ph & 1) - 1 + (t & 1));
        p.tap_p[ph * 8 + t] = 0;
      }
  } else if (up) {
    // output voxel (d, 2y+py, 2x+px) reads low-res rows {y+py-1, y+py}: tap (kd, a, b) of phase (py, px)
    for (int ph = 0; ph < 4; ++ph)
      for (int t = 0; t < 12; ++t) {
        const int kd = t >> 2, a = (t >> 1) & 1, b = t & 1;
        p.tap_d[ph * 12 + t] = (int8_t)(kd - 1);
        p.tap_h[ph * 12 + t] = (int8_t)((ph >> 1) - 1 + a);
        p.tap_w[ph * 12 + t] = (int8_t)((ph & 1) - 1 + b);
        p.tap_p[ph * 12 + t] = 0;
      }
  } else
  for (int t = 0; t < p.taps; ++t) {
    const int kd = t / (g.kh * g.kw), kh = (t / g.kw) % g.kh, kw = t % g.kw;
    p.tap_d[t] = (int8_t)(kd - g.pd);
    if (!s2) {
      p.tap_h[t] = (int8_t)(kh - g.ph);
      p.tap_w[t] = (int8_t)(kw - g.pw);
      p.tap_p[t] = 0;
    } else {   // input row 2*oh + kh - 1: kh=1 -> even phase, offset 0; kh=0 -> odd phase, offset -1; kh=2 -> odd phase, offset 0
      const int ph = (kh == 1) ? 0 : 1, pw = (kw == 1) ? 0 : 1;
      p.tap_h[t] = (int8_t)(kh == 0 ? -1 : 0);
      p.tap_w[t] = (int8_t)(kw == 0 ? -1 : 0);
      p.tap_p[t] = (int8_t)(ph * 2 + pw);
    }
  }
  p.bias = g.bias; p.rowvec = g.rowvec; p.ld_rowvec = g.ld_rowvec;
  p.res = g.res; p.res_bf16 = g.res_dt == BF16; p.ld_res = g.ld_res;
  p.out = g.out; p.out_bf16 = g.out_dt == BF16; p.ldo = g.ldo; p.relu = g.act == 1;
  p.colsum = g.colsum;
  p.out_t = g.out_t;
  p.geglu = g.epi == 1;
  if (plan.splitk > 1) {   // raw fp32 partials into the workspace; splitk_reduce_kernel applies the epilogue terms
    p.bias = nullptr; p.rowvec = nullptr; p.res = nullptr; p.colsum = nullptr; p.relu = 0;
    p.out = g.splitk_ws; p.out_bf16 = 0; p.ldo = g.cout;
  }
  // CTA pairs (cta_group::2) whenever the 128-row tiles pair up: halves the B bytes each SM has to pull from L2
  const bool cta2 = plan.cta2;   // (a pair never straddles two phases: num_m_tiles is even)

  CUtensorMap map_a, map_b, map_a2, map_b2;
  for (int half = 0; half < (x3 ? 2 : 1); ++half) {
    const cuuint64_t dims[5] = {(cuuint64_t)g.cin, (cuuint64_t)in_w, (cuuint64_t)in_h, (cuuint64_t)g.d, (cuuint64_t)in_objs};
    const cuuint64_t strides[4] = {(cuuint64_t)g.cin * 2, (cuuint64_t)g.cin * 2 * in_w, (cuuint64_t)g.cin * 2 * in_w * in_h,
                                   (cuuint64_t)g.cin * 2 * in_w * in_h * g.d};
    const cuuint32_t box[5] = {(cuuint32_t)BLOCK_K, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bd, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_tc.encode(half ? &map_a2 : &map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)(half ? a_lo_ptr : a_ptr), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ECHO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d", (int)r);
  }
  for (int half = 0; half < (x3 ? 2 : 1); ++half) {
    const cuuint64_t ktot = g.up2 == 2 ? (cuuint64_t)64 * g.cin : up ? (cuuint64_t)48 * g.cin : (cuuint64_t)g.ktot();
    const cuuint64_t dims[2] = {ktot, (cuuint64_t)g.cout};
    const cuuint64_t strides[1] = {ktot * 2};
    const cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)(cta2 ? p.block_n / 2 : p.block_n)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = g_tc.encode(half ? &map_b2 : &map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)(half ? g.W_lo : g.W), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ECHO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d", (int)r);
  }
  const int msub = plan.msub;
  const int tiles = p.vm_tiles / msub * p.num_n_tiles * p.splitk;   // CTA tiles
  const int b_rows = cta2 ? p.block_n / 2 : p.block_n;
  if (!x3) { map_a2 = map_a; map_b2 = map_b; }
  p.stages = stages_for(b_rows, msub, x3);
  ECHO_CHECK(p.stages >= 2 && (!x3 || msub == 1), "gemm_tc: tile does not fit shared memory");
  const int smem_bytes = p.stages * (x3 ? 2 : 1) * (msub * A_STAGE_BYTES + b_rows * BLOCK_K * 2) + SMEM_FIXED;
  const bool probed = g_probe.on && g.rows_out() == g_probe.rows && g.cin == g_probe.cin && g.cout == g_probe.cout && g.kd == g_probe.k &&
                      !g.up2 && g.sh == 1;
  if (probed && g_probe.timeline) { p.dbg = g_probe.timeline; g_probe.timeline = nullptr; }
  if (probed) {
    if (g_probe.used == g_probe.ev.size()) {
      cudaEvent_t a, b;
      ECHO_CUDA(cudaEventCreate(&a));
      ECHO_CUDA(cudaEventCreate(&b));
      g_probe.ev.emplace_back(a, b);
    }
    ECHO_CUDA(cudaEventRecord(g_probe.ev[g_probe.used].first, s));
  }
  if (cta2) {
    int grid = tiles < (g_tc.sms & ~1) ? tiles : (g_tc.sms & ~1);   // whole CTA pairs, one CTA per SM
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (x3) ECHO_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, 1, true>, map_a, map_b, map_a2, map_b2, p));
    else if (msub == 2) ECHO_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, 2, false>, map_a, map_b, map_a2, map_b2, p));
    else ECHO_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<true, 1, false>, map_a, map_b, map_a2, map_b2, p));
  } else {
    const int grid = tiles < g_tc.sms ? tiles : g_tc.sms;
    if (x3) launch_pdl(gemm_tc_kernel<false, 1, true>, dim3(grid), dim3(NUM_THREADS), smem_bytes, s, map_a, map_b, map_a2, map_b2, p);
    else if (msub == 2) launch_pdl(gemm_tc_kernel<false, 2, false>, dim3(grid), dim3(NUM_THREADS), smem_bytes, s, map_a, map_b, map_a2, map_b2, p);
    else launch_pdl(gemm_tc_kernel<false, 1, false>, dim3(grid), dim3(NUM_THREADS), smem_bytes, s, map_a, map_b, map_a2, map_b2, p);
  }
  ECHO_LAUNCH_CHECK();
  if (probed) ECHO_CUDA(cudaEventRecord(g_probe.ev[g_probe.used++].second, s));
  if (plan.splitk > 1) {
    dim3 rgrid(cdiv(g.rows_out(), 128), cdiv(g.cout, 64));
    launch_pdl(splitk_reduce_kernel, rgrid, dim3(256), 0, s, (const float*)g.splitk_ws, plan.splitk, (long long)g.rows_out(), g.cout, g.bias,
               g.rowvec, (long long)g.ld_rowvec, (long long)g.od * g.oh * g.ow, (const __nv_bfloat16*)g.res, (long long)g.ld_res,
               g.act == 1 ? 1 : 0, (__nv_bfloat16*)g.out, (long long)g.ldo, g.colsum);
    ECHO_LAUNCH_CHECK();
  }
}

}  // namespace echo

