// Persistent executor of one layout DDPM iteration (layout_mk.cu): the step's ~150 few-row layers recorded once as a
// PROGRAM (stages of ops) and run by ONE cooperative kernel, one CTA per SM, instead of ~180 dependent launches.
#pragma once
#include <stdint.h>

#include <vector>

#include "ops.cuh"

namespace echo {

enum MkType : int { MK_T_LIN = 0, MK_T_COPY = 1, MK_T_EMBROWS = 2 };
// prologue applied to the input rows while they are staged into shared memory
enum MkPro : int {
  MK_NONE = 0,
  MK_SILU = 1,
  MK_GN = 2,      // GroupNorm(32 groups)(+SiLU) over the (concatenated) channels of a row
  MK_LN = 3,      // LayerNorm
  MK_EDGE = 5,    // rows = triples: relu(Ps[s_t] + Pp[t] + Po[o_t] + b1)            (graph.py:146-156, re-associated)
  MK_POOL = 6,    // rows = nodes: mean over the node's CSR items of t2 rows          (graph.py:161-199)
  MK_TEMB = 7,    // row 0 = timestep_embedding(t)                                    (ldm_diffusion_util.py:174-194)
};
enum MkEpi : int {
  MK_EPI_LIN = 0,
  MK_EPI_DDPM = 1,
  MK_EPI_GEGLU = 2,   // W = [2 nout][K]: y[n] = (w[n].x + b[n]) * gelu_erf(w[nout + n].x + b[nout + n])   (attention.py GEGLU)
};
enum MkExt : int { MK_EXT_NONE = 0, MK_EXT_XT = 1, MK_EXT_OBJ = 2, MK_EXT_XPREV = 3, MK_EXT_TNODE = 4 };

struct alignas(16) MkOp {
  const float* X;
  const float* X2;
  const float* W;       // [nout][K] dense fp32
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* aux0;    // EDGE: Pp [T][H];  EMBROWS: table
  const float* aux1;    // EDGE: b1 [H]
  const float* res;
  const float* res2;
  float* Y;
  const void* pad0;
  long long ldx, ldx2, ld_res, ld_res2, ldy, pad1;
  int type, M, K, nout;
  int K1, pro, pro_act, cpg;
  int act, epi, bcast_rows, aux_i;   // aux_i: POOL column offset of the object-role half
  int x_ext, y_ext, FU, n_slices;
  int row_tiles, units, unit_begin, rclass;
  float eps;
  int res_ext;          // 1: the residual is columns [aux_i, aux_i + nout) of the step's row of the time table (one row for all nodes)
  int pad2[6];
};
static_assert(sizeof(MkOp) == 256, "MkOp is copied into shared memory as 16 uint4");

struct alignas(16) MkStage {
  int op_begin, n_a, n_b;   // ops [op_begin, op_begin + n_a): finish before the stage barrier; then n_b background ops
  int bg_wait, bg_arrive;   // background counters (-1: none) waited on before the stage / arrived on after the background ops
  int pad[3];
};

// one unit's weight fetch: rows0 value rows at w, rows1 GEGLU gate rows at g, naux (0 / 2) prologue rows (scale, shift), K floats each
struct alignas(16) MkFetch {
  const float* w;
  const float* g;
  const float* aux;
  const float* aux2;
  int K, rows0, rows1, naux;
};

// what the kernel keeps of every stage in shared memory
struct MkStageLite {
  unsigned short op_begin;
  unsigned char n_a, n_b;
  signed char bg_wait, bg_arrive;
  unsigned short pad;
};
static_assert(sizeof(MkStageLite) == 8, "stage records are copied as 8-byte words");
constexpr int MK_MAX_STAGES = 192;

struct MkArgs {
  const MkStageLite* stages_lite;
  const MkFetch* fetch;    // all CTAs' fetch lists back to back
  const int* fetch_off;    // [ctas + 1]
  const MkOp* ops;
  const MkStage* stages;
  int n_stages;
  unsigned* bar;     // [n_stages] arrival counters, monotonic across launches
  unsigned* bg;      // background counters
  unsigned* epoch;   // launches completed so far (device resident: a captured launch replays correctly)
  unsigned* err;     // set to 1 by the barrier watchdog
  const float* x_t;
  const float* obj_embed;
  const float* noise;
  float* x_prev;
  int t;
  const float* emb_row;     // time table: the stacked emb_layers projections of this step's t (see layout.cu), or null
  const float* tnode_row;   // time table: box_time_emb(emb(t))
  const float* tab;   // DDPM tables 5 x T
  int T;
  const float* freqs;
  int temb_dim;
  const int* s_idx;
  const int* o_idx;
  const int* node_off;
  const int* node_items;
  const long long* triples;
  int H;
  int flags;         // bit 0: stream the weights with the L2 evict-first policy; bit 1: barrier arrival as fence + atomic
  long long* dbg;    // diagnostics (ECHO_MK_TIMELINE): [cta][stage][12] SM clocks: stage entered, barrier passed, foreground ops done, stage left; of the stage's
                     // last 16-row unit: weights landed, rows staged, contraction done, epilogue done, weights issued; feeder call entered / left
};

constexpr int MK_PAD = 16;             // floats of padding behind every staged row (activations and weights): conflict-free LDS.128
constexpr int MK_SLOT_BYTES = 41216;   // one staged weight slice: up to MK_MAX_FU rows x (K + MK_PAD) floats (4 rows of K = 2560)
constexpr int MK_MAX_FU = 24;          // weight rows of a unit (three 8-feature MMA tiles)
constexpr int MK_XROW = 1024;          // staged columns of a 16-row tile per pass (longer rows go in segments)
constexpr int MK_MAX_STAGE_OPS = 8;

// host: fills FU / n_slices / row_tiles / units / rclass of a LIN op for a grid of `ctas`
void mk_plan_op(MkOp& op, int ctas, int min_fu = 4);
bool mk_available(int* ctas_out);
void mk_build_fetch(const std::vector<MkOp>& ops, const std::vector<MkStage>& stages, int ctas, std::vector<MkFetch>& out, std::vector<int>& off);
void mk_launch(const MkArgs& a, int ctas, cudaStream_t s);

}  // namespace echo
