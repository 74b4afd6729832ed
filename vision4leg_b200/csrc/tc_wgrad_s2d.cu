// Weight gradient of the first NatureCNN convolution (Conv2d(4, 32, 8, stride 4), reference
// torchrl/networks/base.py:317-318) on the space-to-depth layouts of the tensor-core tier.
//
//   image  X  [N,16,16,64]  4x4 space-to-depth of the 4x64x64 depth stack (channel = (py4*4+px4)*4 + c)
//   dY        [B,8,8,128]   gradient w.r.t. conv1's pre-activation, stored as 2x2 "cells" of the 15x15x32
//                           output map (channel = (py*2+px)*32 + n; pad positions are exact zeros)
//   dW[n][tap=(dy,dx)][c64] = sum_{image, cell (Y,X), sub (py,px)} X[2Y+py+dy, 2X+px+dx, c] * dY[Y,X,(py,px,n)]
//
// The generic v4l_tc_wgrad walks this as 4 sub-positions x 2 K-slices (+ a bias slice) on separate CTAs and
// re-loads every pixel 4x and every dY cell 12x: 230 MB of L2->SM traffic per launch at minibatch 1024,
// which is what bounded it (ncu: 4.8 TB/s, tensor pipe 14 %).  Here ONE CTA owns a range of images and per
// image loads the 9 distinct stride-2 windows (u,v) in {0,1,2}^2 of the image (window (u,v), row Y*8+X =
// pixel (2Y+u, 2X+v); 8 KB each, the TMA unit does the strided gather) plus the two 64-channel atoms of dY
// ONCE (88 KB instead of 224 KB), and issues per 16 cells
//   D[py][v] (128 lanes x 64 cols) += [ window(py, v) ; window(py+1, v) ]^T (M = 128)  x  dY atom py (N = 64)
// for py in {0,1}, v in {0,1,2}: both operands MN-major straight from the TMA tiles.  In every D the lanes
// 0-63 hold tap row dy = 0 and lanes 64-127 dy = 1 (window row u = py + dy), columns 0-31 sub px = 0 and
// 32-63 px = 1, so the four sub-positions of a tap are four TMEM reads of the SAME lane:
//   dW[(dy,dx)][c][n] = D[0][dx][r][n] + D[0][dx+1][r][32+n] + D[1][dx][r][n] + D[1][dx+1][r][32+n],  r = dy*64 + c.
// The bias gradient rides along as two more MMAs whose A operand is the constant 1.  8 x 64 = 512 TMEM
// columns.  Split-K over images: fp32 partials in the context scratch, summed by the optimiser tail.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int W1_THREADS = 192;            // warp 0 = TMA, warp 1 = TMEM alloc + MMA issue, warps 2-5 = epilogue
constexpr int WIN_BYTES = 64 * 128;        // one window / one dY atom: 64 rows x 64 fp16
constexpr int STAGE_BYTES = 11 * WIN_BYTES;
constexpr int N_STAGES = 2;
constexpr int ONES_BYTES = 16 * 128;       // 16 rows of fp16 1.0 (one K step), shared by every MMA that needs it

struct Conv1WgradParams {
  CUtensorMap tmap_x;            // {64, 16, 16, N} traversed with stride 2 in W and H, box = 8 x 8 positions
  CUtensorMap tmap_dy;           // {128, 8, 8, B}, box {64, 8, 8, 1}
  const int32_t* x_idx;          // minibatch row list or NULL
  int B;
  float* partial;                // [gridDim.x][3][32][128]
};

__global__ void __launch_bounds__(W1_THREADS, 1) tc_wgrad_conv1_kernel(const __grid_constant__ Conv1WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[N_STAGES], empty_bar[N_STAGES], tmem_full;
  __shared__ uint32_t tmem_base_slot;

  v4l_pdl_trigger();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  // Images are dealt round-robin: at any moment the CTAs of the grid stream through ONE contiguous window of
  // gridDim.x images.  With a contiguous range per CTA the concurrent streams are 2^k-strided at power-of-two
  // minibatches and collide in the memory system (measured: 101 ns / sample at minibatch 32768 against 17 at 16384,
  // profiles/r2_trace_sizes.txt).
  const int img_lo = blockIdx.x, img_step = gridDim.x;
  const int img_hi = p.B;
  uint8_t* ones = smem + N_STAGES * STAGE_BYTES;
  {
    const uint32_t one2 = 0x3C003C00u;     // two fp16 1.0
    for (int i = threadIdx.x; i < ONES_BYTES / 16; i += W1_THREADS)
      reinterpret_cast<uint4*>(ones)[i] = make_uint4(one2, one2, one2, one2);
  }
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tmap_x);
    tc::tma_prefetch_desc(&p.tmap_dy);
    for (int s = 0; s < N_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_slot, 512);
  tc::fence_proxy_async();                 // the constant tile is read by the async (UMMA) proxy
  v4l_pdl_wait();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  // The TMA and MMA warps run their loops CONVERGED (warp index made provably uniform with a shuffle) and
  // only the issue itself is done by one elected lane: descriptors / coordinates then live in uniform
  // registers.  With the whole loop under `if (lane == 0)` the compiler has to move every operand of every
  // UTCHMMA / UTMALDG into uniform registers through an elect + broadcast loop (~150 cycles per MMA
  // instead of ~48: tools/ubench/mma_rate.cu, profiles/r2_mma_issue_rate.txt).
  if (warp == 0) {
    int stage = 0; uint32_t phase = 0;
    for (int img = img_lo; img < img_hi; img += img_step) {
      int xi = img;
      if (p.x_idx) xi = __shfl_sync(0xffffffffu, p.x_idx[img], 0);
      tc::mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* s = smem + stage * STAGE_BYTES;
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
        for (int w = 0; w < 9; ++w)        // window (u, v) = (w / 3, w % 3): pixel (2Y + u, 2X + v), zero beyond row/col 15
          tc::tma_load_4d(s + w * WIN_BYTES, &p.tmap_x, &full_bar[stage], 0, w % 3, w / 3, xi);
        tc::tma_load_4d(s + 9 * WIN_BYTES, &p.tmap_dy, &full_bar[stage], 0, 0, 0, img);
        tc::tma_load_4d(s + 10 * WIN_BYTES, &p.tmap_dy, &full_bar[stage], 64, 0, 0, img);
      }
      __syncwarp();
      if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    const uint32_t idesc = tc::umma_idesc_f16(128, 64, 1, 1);      // both operands MN-major
    const uint32_t ones_a = tc::smem_u32(ones);
    const uint64_t ones_desc = tc::umma_smem_desc(ones_a, 0, 1024);
    int stage = 0; uint32_t phase = 0;
    uint32_t acc = 0;
    for (int img = img_lo; img < img_hi; img += img_step) {
      tc::mbar_wait(&full_bar[stage], phase);
      tc::tc_fence_after();
      const uint32_t sx = tc::smem_u32(smem + stage * STAGE_BYTES);
      // descriptors of K step 0; a K step (16 cells) further is +2048 B = +128 in the address field
      const uint64_t a0 = tc::umma_smem_desc(sx, 3 * WIN_BYTES, 1024);
      const uint64_t b0 = tc::umma_smem_desc(sx + 9 * WIN_BYTES, WIN_BYTES, 1024);
      if (tc::elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {                                 // 64 cells = 4 K steps of 16 rows
#pragma unroll
          for (int py = 0; py < 2; ++py) {
            const uint64_t bdesc = b0 + (uint64_t)((py * WIN_BYTES + k * 2048) >> 4);
#pragma unroll
            for (int v = 0; v < 3; ++v) {
              // lanes 0-63: window (py, v), lanes 64-127: window (py + 1, v) = 3 windows further
              const uint64_t adesc = a0 + (uint64_t)(((py * 3 + v) * WIN_BYTES + k * 2048) >> 4);
              tc::umma_f16(tmem_base + (py * 3 + v) * 64, adesc, bdesc, idesc, acc);
            }
            // bias: A = ones (the same 16 rows for every K step, both lane halves)
            tc::umma_f16(tmem_base + (6 + py) * 64, ones_desc, bdesc, idesc, acc);
          }
          acc = 1;
        }
        tc::umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      acc = 1;
      if (++stage == N_STAGES) { stage = 0; phase ^= 1; }
    }
    if (tc::elect_one()) tc::umma_commit(&tmem_full);
    __syncwarp();
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;            // TMEM lane: dy = r >> 6, c = r & 63
    const int dy = r >> 6, c = r & 63;
    float* out = p.partial + (long long)blockIdx.x * 3 * 32 * 128;
    if (img_hi > img_lo) {
      tc::mbar_wait(&tmem_full, 0);
      tc::tc_fence_after();
      const uint32_t ta = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        uint32_t a0[32], a1[32], a2[32], a3[32];
        tc::tmem_ld_32x32(ta + (0 * 3 + dx) * 64, a0);              // sub (0,0) x window (dy, dx)
        tc::tmem_ld_32x32(ta + (0 * 3 + dx + 1) * 64 + 32, a1);     // sub (0,1) x window (dy, dx + 1)
        tc::tmem_ld_32x32(ta + (1 * 3 + dx) * 64, a2);              // sub (1,0) x window (1 + dy, dx)
        tc::tmem_ld_32x32(ta + (1 * 3 + dx + 1) * 64 + 32, a3);     // sub (1,1) x window (1 + dy, dx + 1)
        tc::tmem_ld_wait();
        // packed K index kp = (dy*2 + dx)*64 + c  ->  K slice dy, lane dx*64 + c
        float* o = out + ((long long)dy * 32) * 128 + dx * 64 + c;
#pragma unroll
        for (int n = 0; n < 32; ++n)
          o[n * 128] = ((__uint_as_float(a0[n]) + __uint_as_float(a1[n])) + __uint_as_float(a2[n])) + __uint_as_float(a3[n]);
      }
      {
        uint32_t b0[32], b1[32], b2[32], b3[32];
        tc::tmem_ld_32x32(ta + 6 * 64, b0); tc::tmem_ld_32x32(ta + 6 * 64 + 32, b1);
        tc::tmem_ld_32x32(ta + 7 * 64, b2); tc::tmem_ld_32x32(ta + 7 * 64 + 32, b3);
        tc::tmem_ld_wait();
        if (r == 0) {
#pragma unroll
          for (int n = 0; n < 32; ++n)
            out[((long long)2 * 32 + n) * 128] =
              ((__uint_as_float(b0[n]) + __uint_as_float(b1[n])) + __uint_as_float(b2[n])) + __uint_as_float(b3[n]);
        }
      }
    } else {
      for (int dx = 0; dx < 2; ++dx)
        for (int n = 0; n < 32; ++n) out[((long long)dy * 32 + n) * 128 + dx * 64 + c] = 0.f;
      if (r == 0) for (int n = 0; n < 32; ++n) out[((long long)2 * 32 + n) * 128] = 0.f;
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

extern "C" int v4l_tc_wgrad_flush(v4l_ctx* ctx, void* stream);

extern "C" int v4l_tc_wgrad_conv1(v4l_ctx* ctx, void* stream, const void* x_s2d, int64_t n_img, const int32_t* x_idx,
                                  const void* dy_cells, int B, const int32_t* index, float* dw, float* dbias,
                                  float out_scale, int defer, int accumulate) {
  V4L_REQUIRE(ctx && x_s2d && dy_cells && dw && dbias && n_img > 0 && B >= 0, "v4l_tc_wgrad_conv1: bad argument");
  if (B == 0) return 0;
  Conv1WgradParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[4] = {64, 16, 16, (uint64_t)n_img};
    uint64_t str[3] = {64 * 2, 64 * 16 * 2, 64 * 16 * 16 * 2};
    uint32_t box[4] = {64, 16, 16, 1};          // extent in traversed elements: 8 positions with stride 2
    uint32_t estr[4] = {1, 2, 2, 1};
    if (int r = v4l_encode_tmap(&p.tmap_x, x_s2d, 4, dims, str, box, "v4l_tc_wgrad_conv1(X)", estr)) return r;
  }
  {
    uint64_t dims[4] = {128, 8, 8, (uint64_t)B};
    uint64_t str[3] = {128 * 2, 128 * 8 * 2, 128 * 64 * 2};
    uint32_t box[4] = {64, 8, 8, 1};
    if (int r = v4l_encode_tmap(&p.tmap_dy, dy_cells, 4, dims, str, box, "v4l_tc_wgrad_conv1(dY)", nullptr)) return r;
  }
  p.x_idx = x_idx;
  p.B = B;
  const size_t per_split = (size_t)3 * 32 * 128;
  // a deferred job that cannot get its full split count from what is left of the scratch flushes the pending jobs
  // first: running with the few splits that still fit serialises the whole minibatch on a handful of SMs (minibatch
  // 32768 ran this kernel at 1/6 of its speed that way, profiles/r2_trace_sizes.txt)
  const size_t want = (size_t)min(ctx->sm_count, B) * per_split;
  if (defer && ctx->n_jobs > 0 && (ctx->n_jobs == V4L_MAX_JOBS || ctx->defer_elems - ctx->defer_cursor < want)) {
    if (int r = v4l_tc_wgrad_flush(ctx, stream)) return r;
    ctx->early_flush = 1;
    ++ctx->early_flush_count;
  }
  float* region = defer ? ctx->defer_base + ctx->defer_cursor : ctx->scratch;
  const size_t avail = defer ? ctx->defer_elems - ctx->defer_cursor : ctx->scratch_elems;
  int splits = (int)min((size_t)min(ctx->sm_count, B), avail / per_split);
  V4L_REQUIRE(splits >= 1, "v4l_tc_wgrad_conv1: scratch too small");
  p.partial = region;
  static bool attr_set = false;
  const size_t smem = (size_t)N_STAGES * STAGE_BYTES + ONES_BYTES + 1024;
  if (!attr_set) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_conv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  V4L_LAUNCH(tc_wgrad_conv1_kernel, splits, W1_THREADS, smem, (cudaStream_t)stream, p);
  V4L_CHECK_LAUNCH();
  v4l_reduce_job job;
  job.partial = region; job.index = index; job.dw = dw; job.dbias = dbias;
  job.splits = splits; job.kin_tiles = 2; job.has_bias = 1; job.Nmma = 32;
  job.N_valid = 32; job.Kp = 256; job.scale = out_scale != 0.f ? out_scale : 1.f;
  job.accumulate = accumulate ? 1 : 0;
  if (defer) {
    ctx->jobs[ctx->n_jobs++] = job;
    ctx->defer_cursor += ((size_t)splits * per_split + 63) / 64 * 64;
    return 0;
  }
  v4l_reduce_job saved[V4L_MAX_JOBS];
  const int n_saved = ctx->n_jobs;
  const size_t cursor = ctx->defer_cursor;
  for (int i = 0; i < n_saved; ++i) saved[i] = ctx->jobs[i];
  ctx->jobs[0] = job; ctx->n_jobs = 1;
  const int rc = v4l_tc_wgrad_flush(ctx, stream);
  for (int i = 0; i < n_saved; ++i) ctx->jobs[i] = saved[i];
  ctx->n_jobs = n_saved; ctx->defer_cursor = cursor;
  return rc;
}
