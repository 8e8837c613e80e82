// SDF -> triangle mesh on the GPU: marching cubes over one (R, R, R) volume -- the stage the reference runs per object
// on the CPU with PyMCubes (mcubes.marching_cubes(sdf_i, level), model/diff_utils/util_3d.py:213-218), the last host stage of
// scripts/eval_3dfront.py after the VQ-VAE decode.
//
//   1. mc_classify_kernel   one thread per grid point: which of its three outgoing grid edges are cut (the two ends on different sides
//                           of the level), and for the cell it anchors the case index and triangle count (tables: mc_tables.inc,
//                           derived by tools/gen_mc_tables.py).
//   2. exclusive scans      over the 3 R^3 edge flags and the R^3 triangle counts (two-level block scan): every cut edge gets ONE
//                           vertex index -- vertices are shared between the up to four cells around an edge, as PyMCubes shares them --
//                           every cell its first triangle slot.  Totals go to the caller's counter pair.
//   3. mc_vertices_kernel   one thread per grid edge: p0 + mu * direction, mu = (level - v0) / (v1 - v0) in fp32, index coordinates
//                           (x = first array axis).
//   4. mc_faces_kernel      one thread per cell: its table row, local edge -> grid edge id -> vertex index.
//
// Output order is a function of the volume only (edge id, cell id), so two runs are identical and the CPU oracle reproduces it exactly.
// HBM traffic: the volume is read ~3 times (1 MB at R = 64), the scans move 16 B per grid point, the mesh is written once.
#include "common.cuh"

#define MC_TABLE_QUAL __device__ __constant__
#include "mc_tables.inc"

namespace echo {
namespace {

constexpr int SCAN_BLOCK = 1024;

__device__ __forceinline__ bool inside(float v, float level) { return v < level; }

__global__ void mc_classify_kernel(const float* __restrict__ sdf, int R, float level, int* __restrict__ eflag, int* __restrict__ tcount,
                                   unsigned char* __restrict__ cidx) {
  const int64_t R3 = (int64_t)R * R * R, p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= R3) return;
  const int k = (int)(p % R), j = (int)((p / R) % R), i = (int)(p / ((int64_t)R * R));
  const bool in0 = inside(sdf[p], level);
  eflag[p] = (i + 1 < R && inside(sdf[p + (int64_t)R * R], level) != in0) ? 1 : 0;
  eflag[R3 + p] = (j + 1 < R && inside(sdf[p + R], level) != in0) ? 1 : 0;
  eflag[2 * R3 + p] = (k + 1 < R && inside(sdf[p + 1], level) != in0) ? 1 : 0;
  int ci = 0, nt = 0;
  if (i + 1 < R && j + 1 < R && k + 1 < R) {
    const int64_t sx = (int64_t)R * R, sy = R;
    // corners v0..v7 = (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)
    ci = (in0 ? 1 : 0) | (inside(sdf[p + sx], level) ? 2 : 0) | (inside(sdf[p + sx + sy], level) ? 4 : 0) |
         (inside(sdf[p + sy], level) ? 8 : 0) | (inside(sdf[p + 1], level) ? 16 : 0) | (inside(sdf[p + sx + 1], level) ? 32 : 0) |
         (inside(sdf[p + sx + sy + 1], level) ? 64 : 0) | (inside(sdf[p + sy + 1], level) ? 128 : 0);
    nt = MC_NUM_TRI[ci];
  }
  tcount[p] = nt;
  cidx[p] = (unsigned char)ci;
}

// ---- exclusive scan of int32, two levels: per-block scan + block totals, scan of the totals by one block, add back
__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
  __shared__ int warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_sums[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  const int before = warp ? warp_sums[warp - 1] : 0;
  *total = warp_sums[31];
  __syncthreads();
  return before + x - v;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_blocks_kernel(int* __restrict__ a, int64_t n, int* __restrict__ bsum) {
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  const int v = i < n ? a[i] : 0;
  int total;
  const int ex = block_exclusive_scan(v, &total);
  if (i < n) a[i] = ex;
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_totals_kernel(int* __restrict__ bsum, int nb, int* __restrict__ grand_total) {
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nb; b0 += SCAN_BLOCK) {
    const int i = b0 + threadIdx.x;
    const int v = i < nb ? bsum[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, &total);
    const int carry = carry_s;
    if (i < nb) bsum[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand_total = carry_s;
}

__global__ void __launch_bounds__(SCAN_BLOCK) scan_add_kernel(int* __restrict__ a, int64_t n, const int* __restrict__ bsum) {
  const int64_t i = (int64_t)blockIdx.x * SCAN_BLOCK + threadIdx.x;
  if (i < n) a[i] += bsum[blockIdx.x];
}

void exclusive_scan(int* a, int64_t n, int* bsum, int* total, cudaStream_t s) {
  const int nb = cdiv(n, SCAN_BLOCK);
  scan_blocks_kernel<<<nb, SCAN_BLOCK, 0, s>>>(a, n, bsum);
  ECHO_LAUNCH_CHECK();
  scan_totals_kernel<<<1, SCAN_BLOCK, 0, s>>>(bsum, nb, total);
  ECHO_LAUNCH_CHECK();
  scan_add_kernel<<<nb, SCAN_BLOCK, 0, s>>>(a, n, bsum);
  ECHO_LAUNCH_CHECK();
}

__global__ void mc_vertices_kernel(const float* __restrict__ sdf, int R, float level, const int* __restrict__ escan, float* __restrict__ verts,
                                   int64_t cap_v) {
  const int64_t R3 = (int64_t)R * R * R, e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 3 * R3) return;
  const int d = (int)(e / R3);
  const int64_t p = e - d * R3;
  const int k = (int)(p % R), j = (int)((p / R) % R), i = (int)(p / ((int64_t)R * R));
  const int c = d == 0 ? i : (d == 1 ? j : k);
  if (c + 1 >= R) return;
  const int64_t q = p + (d == 0 ? (int64_t)R * R : (d == 1 ? R : 1));
  const float v0 = sdf[p], v1 = sdf[q];
  if (inside(v0, level) == inside(v1, level)) return;
  const int64_t idx = escan[e];
  if (idx >= cap_v) return;
  const float mu = (level - v0) / (v1 - v0);
  float x = (float)i, y = (float)j, z = (float)k;
  if (d == 0) x = x + mu;
  else if (d == 1) y = y + mu;
  else z = z + mu;
  verts[idx * 3 + 0] = x;
  verts[idx * 3 + 1] = y;
  verts[idx * 3 + 2] = z;
}

__global__ void mc_faces_kernel(int R, const unsigned char* __restrict__ cidx, const int* __restrict__ tscan, const int* __restrict__ escan,
                                int* __restrict__ faces, int64_t cap_f) {
  const int64_t R3 = (int64_t)R * R * R, p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= R3) return;
  const int ci = cidx[p], nt = MC_NUM_TRI[ci];
  if (nt == 0) return;
  const int64_t base = tscan[p];
  for (int t = 0; t < nt; ++t) {
    if (base + t >= cap_f) return;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int e = MC_TRI_TABLE[ci][3 * t + c];
      const int64_t g = (int64_t)MC_EDGE_BASE[e][3] * R3 + p + ((int64_t)MC_EDGE_BASE[e][0] * R + MC_EDGE_BASE[e][1]) * R + MC_EDGE_BASE[e][2];
      faces[(base + t) * 3 + c] = escan[g];
    }
  }
}

}  // namespace

size_t mesh_workspace_bytes(int R) {
  const size_t R3 = (size_t)R * R * R;
  const size_t nb = cdiv((int64_t)(3 * R3), SCAN_BLOCK) + cdiv((int64_t)R3, SCAN_BLOCK) + 8;
  return sizeof(int) * (3 * R3 + R3 + nb) + R3 + 64;
}

void mesh_marching_cubes(const float* sdf, int R, float level, float* verts, int64_t cap_v, int* faces, int64_t cap_f, int* counts,
                         void* workspace, size_t ws_bytes, cudaStream_t s) {
  ECHO_CHECK(sdf && counts && workspace, "marching_cubes: null argument");
  ECHO_CHECK(R >= 2 && R <= 512, "marching_cubes: resolution %d outside [2, 512]", R);
  ECHO_CHECK(ws_bytes >= mesh_workspace_bytes(R), "marching_cubes: workspace of %zu bytes, need %zu (echo_mesh_workspace_bytes)", ws_bytes,
             mesh_workspace_bytes(R));
  ECHO_CHECK((cap_v == 0 || verts) && (cap_f == 0 || faces) && cap_v >= 0 && cap_f >= 0, "marching_cubes: null output with a capacity");
  ECHO_CHECK(((uintptr_t)workspace % 4) == 0, "marching_cubes: unaligned workspace");
  const int64_t R3 = (int64_t)R * R * R;
  int* escan = (int*)workspace;
  int* tscan = escan + 3 * R3;
  int* bsum_e = tscan + R3;
  int* bsum_t = bsum_e + cdiv(3 * R3, SCAN_BLOCK) + 4;
  unsigned char* cidx = (unsigned char*)(bsum_t + cdiv(R3, SCAN_BLOCK) + 4);
  mc_classify_kernel<<<cdiv(R3, 256), 256, 0, s>>>(sdf, R, level, escan, tscan, cidx);
  ECHO_LAUNCH_CHECK();
  exclusive_scan(escan, 3 * R3, bsum_e, counts, s);
  exclusive_scan(tscan, R3, bsum_t, counts + 1, s);
  if (cap_v > 0) {
    mc_vertices_kernel<<<cdiv(3 * R3, 256), 256, 0, s>>>(sdf, R, level, escan, verts, cap_v);
    ECHO_LAUNCH_CHECK();
  }
  if (cap_f > 0) {
    mc_faces_kernel<<<cdiv(R3, 256), 256, 0, s>>>(R, cidx, tscan, escan, faces, cap_f);
    ECHO_LAUNCH_CHECK();
  }
}

}  // namespace echo
