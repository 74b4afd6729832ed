// Context, error reporting and copy helpers of libv4l_b200.so.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void v4l_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool v4l_pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    // measured on B200 (round 1): with the side-stream branches of the captured graph, early-resident
    // dependents (1 CTA/SM, ~200 KB smem each) hold SMs the weight-gradient branch could use:
    // 606 k samples/s with PDL vs 657 k without -> opt-in until the triggers are placed per kernel
    const char* e = getenv("V4L_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
  }
  return on == 1;
}

extern "C" int v4l_version(void) { return V4L_ABI_VERSION; }

extern "C" const char* v4l_last_error(void) { return g_err; }

extern "C" int v4l_ctx_create(v4l_ctx** out, int device, size_t scratch_bytes) {
  V4L_REQUIRE(out != nullptr, "v4l_ctx_create: out is NULL");
  int count = 0;
  V4L_CHECK_CUDA(cudaGetDeviceCount(&count));
  V4L_REQUIRE(device >= 0 && device < count, "v4l_ctx_create: bad device %d (have %d)", device, count);
  V4L_CHECK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  V4L_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  V4L_REQUIRE(prop.major == 10, "libv4l_b200 is built for sm_100a only; device %d is sm_%d%d",
              device, prop.major, prop.minor);
  v4l_ctx* c = new v4l_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  // default 1 GiB: the upper half holds the split-K partials of every weight gradient of one backward pass until the
  // optimiser tail reduces them — at most ~21 jobs x 148 CTAs x 128 x 256 floats = 410 MB however large the minibatch
  if (scratch_bytes == 0) scratch_bytes = (size_t)1 << 30;
  const size_t total_elems = scratch_bytes / sizeof(float);
  c->scratch_elems = total_elems / 2;
  c->defer_elems = total_elems - c->scratch_elems;
  c->defer_cursor = 0;
  c->n_jobs = 0;
  c->early_flush = 0;
  c->early_flush_count = 0;
  cudaError_t e = cudaMalloc(&c->scratch, scratch_bytes);
  if (e != cudaSuccess) {
    v4l_set_error("v4l_ctx_create: cudaMalloc(%zu) -> %s", scratch_bytes, cudaGetErrorString(e));
    delete c;
    return -2;
  }
  c->defer_base = c->scratch + c->scratch_elems;
  c->counters = nullptr;
  e = cudaMalloc(&c->counters, V4L_N_COUNTERS * sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(c->counters, 0, V4L_N_COUNTERS * sizeof(unsigned int));
  if (e != cudaSuccess) {
    v4l_set_error("v4l_ctx_create: counters -> %s", cudaGetErrorString(e));
    cudaFree(c->scratch);
    delete c;
    return -2;
  }
  *out = c;
  return 0;
}

extern "C" int v4l_ctx_destroy(v4l_ctx* ctx) {
  if (!ctx) return 0;
  cudaFree(ctx->scratch);
  cudaFree(ctx->counters);
  delete ctx;
  return 0;
}

extern "C" int v4l_ctx_sm_count(const v4l_ctx* ctx) { return ctx ? ctx->sm_count : -1; }
extern "C" int v4l_ctx_early_flushes(const v4l_ctx* ctx) { return ctx ? ctx->early_flush_count : -1; }

extern "C" int v4l_h2d_2d(void* stream, void* dst, size_t dpitch, const void* h_src, size_t spitch,
                          size_t width, size_t height) {
  V4L_CHECK_CUDA(cudaMemcpy2DAsync(dst, dpitch, h_src, spitch, width, height,
                                   cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return 0;
}

// Rows t of a pinned host matrix -> the same rows of a device matrix (row_bytes each), one call per
// minibatch: sorted, adjacent rows merged, then ONE cudaMemcpyBatchAsync (copy engine; no per-row
// driver call).  Falls back to a loop of cudaMemcpyAsync if the driver rejects the batch call.
extern "C" int v4l_h2d_rows(void* stream, void* dst_base, const void* src_base, const int32_t* rows,
                            int n_rows, size_t row_bytes) {
  if (n_rows <= 0) return 0;
  if (!dst_base || !src_base || !rows) { v4l_set_error("v4l_h2d_rows: NULL argument"); return -1; }
  static thread_local std::vector<int32_t> order;
  static thread_local std::vector<void*> dsts, srcs;
  static thread_local std::vector<size_t> sizes;
  order.assign(rows, rows + n_rows);
  std::sort(order.begin(), order.end());
  dsts.clear(); srcs.clear(); sizes.clear();
  for (int i = 0; i < n_rows;) {
    int j = i + 1;
    while (j < n_rows && order[j] == order[j - 1] + 1) ++j;
    const size_t off = (size_t)order[i] * row_bytes;
    dsts.push_back((char*)dst_base + off);
    srcs.push_back((char*)const_cast<void*>(src_base) + off);
    sizes.push_back((size_t)(j - i) * row_bytes);
    i = j;
  }
  cudaStream_t s = (cudaStream_t)stream;
  static int batch_ok = -1;                 // -1 untested, 0 unsupported, 1 works
  if (batch_ok != 0) {
    cudaMemcpyAttributes attr;
    memset(&attr, 0, sizeof(attr));
    attr.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
    attr.flags = cudaMemcpyFlagPreferOverlapWithCompute;
    size_t attr_idx = 0, fail = 0;
    cudaError_t e = cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), dsts.size(), &attr, &attr_idx, 1,
                                         &fail, s);
    if (e == cudaSuccess) { batch_ok = 1; return 0; }
    if (batch_ok == 1) { v4l_set_error("cudaMemcpyBatchAsync -> %s", cudaGetErrorString(e)); return -2; }
    (void)cudaGetLastError();
    batch_ok = 0;
  }
  for (size_t i = 0; i < dsts.size(); ++i)
    V4L_CHECK_CUDA(cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], cudaMemcpyHostToDevice, s));
  return 0;
}
