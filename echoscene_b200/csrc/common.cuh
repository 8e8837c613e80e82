// Shared host/device plumbing of libechoscene_b200: error reporting, dtype tags, workspace arena,
// launch counting.  No torch types anywhere in this library.
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <string>
#include <stdexcept>

#include "../../include/echoscene_b200.h"

namespace echo {

enum DT : int { F32 = 0, BF16 = 1 };
static inline size_t dt_size(DT d) { return d == F32 ? 4 : 2; }

// ---- errors: C++ exceptions inside, converted to codes + thread-local message at the C boundary ----------------
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const char* msg);
[[noreturn]] void fail(int code, const char* fmt, ...);

#define ECHO_CUDA(expr)                                                                                   \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess)                                                                                \
      ::echo::fail(ECHO_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define ECHO_CHECK(cond, ...)                                   \
  do {                                                          \
    if (!(cond)) ::echo::fail(ECHO_ERR_INVALID, __VA_ARGS__);   \
  } while (0)

// ---- launch accounting (bench.py reports gpu_launches from this) ------------------------------------------------
extern thread_local int64_t g_launches;
static inline void count_launch(int n = 1) { g_launches += n; }
// after every kernel launch
#define ECHO_LAUNCH_CHECK()                                                                              \
  do {                                                                                                   \
    ::echo::count_launch();                                                                              \
    cudaError_t _e = cudaPeekAtLastError();                                                              \
    if (_e != cudaSuccess) {                                                                             \
      cudaGetLastError();                                                                                \
      ::echo::fail(ECHO_ERR_CUDA, "%s:%d: kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
    }                                                                                                    \
  } while (0)

// ---- workspace: one cudaMalloc at create, stack discipline inside a step (same addresses every step, so a
//      step can be captured in a CUDA graph) ---------------------------------------------------------------------
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, high = 0;
  void init(size_t bytes) {
    ECHO_CUDA(cudaMalloc((void**)&base, bytes));
    cap = bytes;
    off = 0;
  }
  void destroy() {
    if (base) cudaFree(base);
    base = nullptr;
  }
  void* alloc(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    if (a + bytes > cap) fail(ECHO_ERR_NOMEM, "workspace exhausted: need %zu more bytes (cap %zu)", a + bytes - cap, cap);
    off = a + bytes;
    if (off > high) high = off;
    return base + a;
  }
  template <class T>
  T* alloc_n(size_t n) { return (T*)alloc(n * sizeof(T)); }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

// channels-last activation (n, d, h, w, c)
struct Act {
  void* p = nullptr;
  DT dt = F32;
  int n = 0, d = 1, h = 1, w = 1, c = 0;
  // optional per-column (sum, sumsq) partials written by the producing tcgen05 GEMM: [n * colsum_rows][c][2]
  float* colsum = nullptr;
  int colsum_rows = 0;
  int64_t voxels() const { return (int64_t)d * h * w; }
  int64_t rows() const { return (int64_t)n * d * h * w; }
  size_t bytes() const { return (size_t)rows() * c * dt_size(dt); }
};

// debugging aids (runtime.cu): ECHO_SKIP=name,name drops whole kernel classes from a step (WRONG results; in-situ cost
// attribution by difference), ECHO_TRACE=1 prints one line per contraction launch to stderr
bool dbg_skip(const char* name);
bool dbg_trace();

// ---- programmatic dependent launch: the next kernel of the stream is scheduled while this one drains; everything it
//      does before griddep_wait() (barrier init, TMEM allocation, descriptor prefetch, parameter math) overlaps the tail
//      of its predecessor.  Rule: a kernel launched through launch_pdl() MUST execute griddep_wait() before it touches
//      global memory another kernel may write or may still be reading (the wait returns once the predecessor grid has
//      completed and its writes are visible; dependency chains stay transitive because every link waits).
#ifdef __CUDACC__
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();   // ECHO_NO_PDL=1 turns the attribute off (A/B timing)
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  ECHO_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
}

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace echo
