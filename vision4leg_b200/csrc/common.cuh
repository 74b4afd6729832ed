// Shared declarations for libv4l_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "v4l_b200.h"

// one deferred weight-gradient reduction (see tc_gemm.cu: v4l_tc_wgrad with defer = 1)
struct v4l_reduce_job {
  const float* partial;      // [splits][kin_tiles + has_bias][Nmma][128]  (packed-K lane fastest)
  const int32_t* index;      // packing table or NULL
  float* dw;
  float* dbias;              // or NULL
  int splits, kin_tiles, has_bias, Nmma, N_valid, Kp;
  float scale;
  int accumulate;            // dw += instead of dw = (gradient accumulation over micro-batches)
};
constexpr int V4L_MAX_JOBS = 48;

struct v4l_ctx {
  int device;
  int sm_count;
  float* scratch;        // context-owned scratch (split partials)
  size_t scratch_elems;  // elements usable by immediate users = lower half; the upper half holds
                         // the partials of deferred weight-gradient reductions until the flush
  float* defer_base;
  size_t defer_elems, defer_cursor;
  v4l_reduce_job jobs[V4L_MAX_JOBS];
  int n_jobs;
  int early_flush_count;    // how many times that happened since the context was created (v4l_ctx_early_flushes)
  int early_flush;          // a deferred reduction was flushed before the optimiser tail (scratch full): the
                            // tail then takes the gradient norm from the bucket instead of from its own writes
  unsigned int* counters;   // V4L_N_COUNTERS zero-initialised device words: last-CTA arrival counters and
                            // the grid barrier of the fused optimiser tail (each user re-arms its own)
};
constexpr int V4L_N_COUNTERS = 64;

void v4l_set_error(const char* fmt, ...);

#define V4L_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      v4l_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                      \
    }                                                                                 \
  } while (0)

#define V4L_CHECK_LAUNCH()                                                            \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      v4l_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return -3;                                                                      \
    }                                                                                 \
  } while (0)

#define V4L_REQUIRE(cond, ...)                                                        \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      v4l_set_error(__VA_ARGS__);                                                     \
      return -1;                                                                      \
    }                                                                                 \
  } while (0)

// ---- programmatic dependent launch (PDL) -----------------------------------------------------
// Every kernel of the library (a) lets its successor start launching right away and (b) blocks
// until its predecessor has completed and flushed before touching global memory.  Launches go
// through v4l_launch(), which sets programmaticStreamSerialization, so the successor's launch
// latency and prologue (barrier init, TMEM allocation, descriptor prefetch) overlap the tail of
// the predecessor; in a captured CUDA graph these become programmatic dependency edges.
__device__ __forceinline__ void v4l_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void v4l_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void v4l_pdl_enter() { v4l_pdl_trigger(); v4l_pdl_wait(); }

bool v4l_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t v4l_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                     cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = v4l_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define V4L_LAUNCH(kernel, grid, block, smem, stream, ...)                                   \
  do {                                                                                       \
    cudaError_t _le = v4l_launch(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__); \
    if (_le != cudaSuccess) {                                                                \
      v4l_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_le));      \
      return -3;                                                                             \
    }                                                                                        \
  } while (0)

__device__ __forceinline__ long long v4l_row_addr(const v4l_rowmap& rm, int m) {
  int item = m / rm.P;
  int pos = m - item * rm.P;
  if (rm.idx) item = rm.idx[item];
  long long off = rm.pos_off ? (long long)rm.pos_off[pos] : (long long)pos * rm.pos_stride;
  return rm.base + (long long)item * rm.item_stride + off;
}

static inline int v4l_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
