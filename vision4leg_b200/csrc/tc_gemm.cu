// Tensor-core tier: tap-shifted TMA + tcgen05 GEMM (f16 operands, fp32 accumulators in TMEM).
//
//   D[128-row tile, N] = sum_{tap, kc} A_tap[rows, 64 k] * W[N, (tap, kc, 64 k)]^T
//
// One kernel covers every forward layer and every data-gradient of the PPO networks:
//   * A is an NHWC f16 activation viewed as a 4-D TMA tensor {C, W, H, B}; a row tile is a TMA
//     box {64, bw, bh, bb} (<= 128 rows).  Convolutions are a sum over taps of the SAME box
//     shifted by (dw, dh) — implicit im2col done by the TMA unit, out-of-range pixels zero-filled
//     by the hardware (this also gives the "full" correlation of the data-gradient with negative
//     shifts).  Plain matrices are the degenerate case W = H = 1, B = M.
//   * W is a packed f16 weight matrix [N, taps*kchunks*64] (K-major), one TMA box {64, N}.
//   * both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes.
// Warp roles (576 threads, 1 CTA/SM, persistent over tiles): warp 0 = TMA producer, warp 1 =
// TMEM allocator + single-thread MMA issuer, then 2 (N > 128) or 4 epilogue warpgroups, one per
// TMEM accumulator stage (TMEM -> registers -> bias / ReLU / ReLU-mask / accumulate -> global).
// 4-8-stage smem ring; the epilogue of tile i overlaps the MMAs and the epilogue of tile i+1.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TC_THREADS = 64 + 4 * 128;          // TMA warp, MMA warp, up to 4 epilogue warpgroups
constexpr int TC_MAX_ACC = 4;
constexpr int TC_MAX_STAGES = 8;                 // TMA ring depth: as many stages as fit in ~196 KB
constexpr int A_TILE_BYTES = 128 * 128;          // 128 rows x 64 f16
constexpr int ACC_COLS = 256;                    // TMEM columns per accumulator stage
constexpr int MAX_TAPS = 16;

struct TcGemmParams {
  CUtensorMap tmap_a;
  CUtensorMap tmap_b;
  int B, Hout, Wout;
  int bw, bh, bb;
  int h_tiles, num_tiles;
  int n_taps, kchunks;
  int tap_dw[MAX_TAPS], tap_dh[MAX_TAPS];
  int N, N_total, N_valid;        // UMMA N per chunk (<=256), padded rows of W, real outputs
  const float* bias;
  void* c;
  v4l_rowmap c_map;
  int c_f32;
  const __half* mask;
  int flags;
  const int32_t* a_idx;
  const __half* res;              // optional residual added to the result (same addressing as c)
  int stages;                     // TMA ring depth
};

__device__ __forceinline__ float h16_lo(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xffffu))); }
__device__ __forceinline__ float h16_hi(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))); }
__device__ __forceinline__ uint32_t pack_h16(float a, float b) {
  __half2 v = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16-B aligned: round up to 1024 (SWIZZLE_128B atoms)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[TC_MAX_STAGES], empty_bar[TC_MAX_STAGES], tmem_full[TC_MAX_ACC], tmem_empty[TC_MAX_ACC];
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_bias[256];

  v4l_pdl_trigger();               // the next kernel may start its prologue now
  // warp index through a shuffle: provably warp-uniform, so the producer / issuer loops below run converged
  // and their descriptors and coordinates stay in uniform registers (see tc_wgrad_s2d.cu)
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const int n_chunk = blockIdx.y;
  const int n0 = n_chunk * p.N;
  const int N = min(p.N, p.N_total - n0);                 // UMMA N of this chunk (multiple of 16)
  const uint32_t b_bytes = static_cast<uint32_t>(N) * 128u;
  const uint32_t stage_bytes = A_TILE_BYTES + static_cast<uint32_t>(p.N) * 128u;
  const int box_rows = p.bw * p.bh * p.bb;
  const uint32_t a_bytes = static_cast<uint32_t>(box_rows) * 128u;
  const int k_iters = p.n_taps * p.kchunks;
  // accumulator stages in TMEM = epilogue warpgroups (narrow layers: the MMA -> epilogue -> MMA
  // hand-over, not the MMAs, bounds a tile, so several tiles must be in flight)
  const int n_acc = p.N <= 128 ? 4 : 2;
  const int acc_cols = 512 / n_acc;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tmap_a);
    tc::tma_prefetch_desc(&p.tmap_b);
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < TC_MAX_ACC; ++s) { tc::mbar_init(&tmem_full[s], 1); tc::mbar_init(&tmem_empty[s], 128); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_slot, 512);
  // prologue done without touching global memory: now wait for the predecessor grid
  v4l_pdl_wait();
  for (int i = threadIdx.x; i < 256; i += TC_THREADS) {
    const int n = n_chunk * p.N + i;
    s_bias[i] = (p.bias && i < p.N && n < p.N_valid) ? p.bias[n] : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    // whole warp converged, one elected lane issues
    {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        int b0, h0;
        if (p.bb == 1) { b0 = tile / p.h_tiles; h0 = (tile - b0 * p.h_tiles) * p.bh; }
        else           { b0 = tile * p.bb; h0 = 0; }
        // minibatch row gather (bb == 1 only)
        int ab0 = b0;
        if (p.a_idx) ab0 = __shfl_sync(0xffffffffu, p.a_idx[b0], 0);
        for (int it = 0; it < k_iters; ++it) {
          const int tap = it / p.kchunks, kc = it - tap * p.kchunks;
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + A_TILE_BYTES;
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&full_bar[stage], a_bytes + b_bytes);
            tc::tma_load_4d(sa, &p.tmap_a, &full_bar[stage], kc * 64, p.tap_dw[tap], h0 + p.tap_dh[tap], ab0);
            tc::tma_load_2d(sb, &p.tmap_b, &full_bar[stage], it * 64, n0);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    {
      const uint32_t idesc = tc::umma_idesc_f16(128, N, 0, 0);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        tc::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_cols;
        for (int it = 0; it < k_iters; ++it) {
          tc::mbar_wait(&full_bar[stage], phase);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc0 = tc::umma_smem_desc(sa, 0, 1024);
          const uint64_t bdesc0 = tc::umma_smem_desc(sa + A_TILE_BYTES, 0, 1024);
          if (tc::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)                        // 4 x UMMA_K(16) = 64; +32 B = +2 in the address field
              tc::umma_f16(d_tmem, adesc0 + 2 * k, bdesc0 + 2 * k, idesc, (it | k) ? 1u : 0u);
            tc::umma_commit(&empty_bar[stage]);                // frees the smem stage when MMAs retire
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (tc::elect_one()) tc::umma_commit(&tmem_full[acc]); // accumulator ready for the epilogue
        __syncwarp();
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (((warp - 2) >> 2) < n_acc) {
    // ============================== epilogue ==================================
    // one epilogue warpgroup per TMEM accumulator stage, tiles round-robin: unloading tile i
    // overlaps the MMAs and the unloading of the following tiles
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;                               // TMEM lane quadrant of this warp
    const int r = quad * 32 + lane;                          // row inside the tile
    const bool relu = p.flags & V4L_RELU, accum = p.flags & V4L_ACCUM;
    const int acc = wg;
    uint32_t acc_phase = 0;
    // row r of the box = (w fastest, then h, then b)
    const int ww = r % p.bw;
    const int t2 = r / p.bw;
    const int hh = t2 % p.bh;
    const int bi = t2 / p.bh;
    // output address of this thread's row in a tile (table look-ups inside): computed one tile
    // ahead so that its global-load latency hides behind the current tile
    auto row_of = [&](int tile, bool& ok) -> long long {
      int b0, h0;
      if (p.bb == 1) { b0 = tile / p.h_tiles; h0 = (tile - b0 * p.h_tiles) * p.bh; }
      else           { b0 = tile * p.bb; h0 = 0; }
      const int b = b0 + bi, h = h0 + hh;
      ok = (r < box_rows) && (b < p.B) && (h < p.Hout);
      return ok ? v4l_row_addr(p.c_map, (b * p.Hout + h) * p.Wout + ww) + n0 : 0;
    };
    const int tile_step = n_acc * gridDim.x;
    int tile = blockIdx.x + wg * gridDim.x;
    bool ok_next = false;
    long long addr_next = tile < p.num_tiles ? row_of(tile, ok_next) : 0;
    for (; tile < p.num_tiles; tile += tile_step) {
      const bool row_ok = ok_next;
      const long long row_addr = addr_next;
      if (tile + tile_step < p.num_tiles) addr_next = row_of(tile + tile_step, ok_next);

      tc::mbar_wait(&tmem_full[acc], acc_phase);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * acc_cols;
      for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        if (N - c0 >= 32) {
          tc::tmem_ld_32x32(taddr + c0, v);
        } else {
          uint32_t v16[16];
          tc::tmem_ld_32x16(taddr + c0, v16);
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = v16[j]; v[16 + j] = 0; }
        }
        tc::tmem_ld_wait();
        if (!row_ok) continue;
        const int ncols = min(32, N - c0);
        const int nabs = n0 + c0;                            // absolute first column of this chunk
        const long long addr = row_addr + c0;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
          if (relu) f[j] = fmaxf(f[j], 0.f);
        }
        if (!p.c_f32 && ncols == 32 && nabs + 32 <= p.N_valid && ((addr & 7) == 0)) {
          // fast path: 32 fp16 outputs = 4 x 16-byte stores per row
          __half* cp = reinterpret_cast<__half*>(p.c) + addr;
          if (p.mask) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 m = __ldg(reinterpret_cast<const uint4*>(p.mask + addr) + q);
              const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (!(h16_lo(mw[j]) > 0.f)) f[8 * q + 2 * j] = 0.f;
                if (!(h16_hi(mw[j]) > 0.f)) f[8 * q + 2 * j + 1] = 0.f;
              }
            }
          }
          if (p.res) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 o = __ldg(reinterpret_cast<const uint4*>(p.res + addr) + q);
              const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) { f[8 * q + 2 * j] += h16_lo(ow[j]); f[8 * q + 2 * j + 1] += h16_hi(ow[j]); }
            }
          }
          if (accum) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 o = *(reinterpret_cast<const uint4*>(cp) + q);
              const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) { f[8 * q + 2 * j] += h16_lo(ow[j]); f[8 * q + 2 * j + 1] += h16_hi(ow[j]); }
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_h16(f[8 * q + 0], f[8 * q + 1]); o.y = pack_h16(f[8 * q + 2], f[8 * q + 3]);
            o.z = pack_h16(f[8 * q + 4], f[8 * q + 5]); o.w = pack_h16(f[8 * q + 6], f[8 * q + 7]);
            *(reinterpret_cast<uint4*>(cp) + q) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (j >= ncols || nabs + j >= p.N_valid) continue;
            float x = f[j];
            if (p.mask && !(__half2float(p.mask[addr + j]) > 0.f)) x = 0.f;
            if (p.res) x += __half2float(p.res[addr + j]);
            if (p.c_f32) {
              float* cp = reinterpret_cast<float*>(p.c) + addr + j;
              if (accum) x += *cp;
              *cp = x;
            } else {
              __half* cp = reinterpret_cast<__half*>(p.c) + addr + j;
              if (accum) x += __half2float(*cp);
              *cp = __float2half(x);
            }
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(&tmem_empty[acc]);
      acc_phase ^= 1;
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

}  // namespace

int v4l_encode_tmap(CUtensorMap* out, const void* gaddr, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const char* who,
                    const uint32_t* elem_strides) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn) {
      v4l_set_error("%s: cuTensorMapEncodeTiled entry point unavailable (%s)", who, cudaGetErrorString(e));
      return -2;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; if (elem_strides) estr[i] = elem_strides[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(gaddr), gd, gs, bx, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    v4l_set_error("%s: cuTensorMapEncodeTiled failed (CUresult %d; rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)",
                  who, (int)r, rank, (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0),
                  (unsigned long long)(rank > 2 ? gd[2] : 0), (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0],
                  rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return -2;
  }
  return 0;
}

extern "C" int v4l_tc_gemm(v4l_ctx* ctx, void* stream, const v4l_tc_gemm_args* a) {
  V4L_REQUIRE(ctx && a && a->a && a->w && a->c, "v4l_tc_gemm: NULL argument");
  V4L_REQUIRE(a->n_taps >= 1 && a->n_taps <= MAX_TAPS && a->kchunks >= 1, "v4l_tc_gemm: bad taps/kchunks");
  V4L_REQUIRE(a->a_C % 8 == 0 && a->kchunks * 64 <= ((a->a_C + 63) / 64) * 64, "v4l_tc_gemm: bad channel count %d", a->a_C);
  V4L_REQUIRE(a->N_pad % 16 == 0 && a->N_pad >= 16, "v4l_tc_gemm: N_pad=%d must be a multiple of 16", a->N_pad);
  V4L_REQUIRE(a->N_pad <= 256 || a->N_pad % 256 == 0, "v4l_tc_gemm: N_pad > 256 must be a multiple of 256");
  V4L_REQUIRE(a->N_pad <= 4096, "v4l_tc_gemm: N_pad too large");
  V4L_REQUIRE(a->N_valid >= 1 && a->N_valid <= a->N_pad, "v4l_tc_gemm: bad N_valid");
  const int rows = a->bw * a->bh * a->bb;
  V4L_REQUIRE(rows >= 1 && rows <= 128 && a->bw == a->Wout && a->bw <= 256 && a->bh <= 256 && a->bb <= 256,
              "v4l_tc_gemm: bad box %dx%dx%d (Wout=%d)", a->bw, a->bh, a->bb, a->Wout);
  V4L_REQUIRE(a->bb == 1 || a->bh == a->Hout, "v4l_tc_gemm: multi-item boxes must cover the whole image");
  V4L_REQUIRE(a->c_map.P > 0, "v4l_tc_gemm: row map with P <= 0");
  if (a->B == 0) return 0;

  TcGemmParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[4] = {(uint64_t)a->a_C, (uint64_t)a->a_W, (uint64_t)a->a_H, (uint64_t)a->a_B};
    uint64_t str[3] = {(uint64_t)a->a_C * 2, (uint64_t)a->a_C * a->a_W * 2, (uint64_t)a->a_C * a->a_W * a->a_H * 2};
    if (a->a_sW) { str[0] = (uint64_t)a->a_sW * 2; str[1] = (uint64_t)a->a_sH * 2; str[2] = (uint64_t)a->a_sB * 2; }
    uint32_t box[4] = {64, (uint32_t)a->bw, (uint32_t)a->bh, (uint32_t)a->bb};
    if (int r = v4l_encode_tmap(&p.tmap_a, a->a, 4, dims, str, box, "v4l_tc_gemm(A)", nullptr)) return r;
  }
  const int Ktot = a->n_taps * a->kchunks * 64;
  // N per CTA: the whole (<=256) width, or 64-wide slices when there are too few row tiles to
  // fill the machine (small-M layers: proprio MLP, heads)
  const int tiles_est = (a->bb == 1) ? a->B * v4l_cdiv(a->Hout, a->bh) : v4l_cdiv(a->B, a->bb);
  int Nchunk = a->N_pad > 256 ? 256 : a->N_pad;
  if (tiles_est * 2 <= ctx->sm_count && a->N_pad % 64 == 0 && a->N_pad > 64) Nchunk = 64;
  {
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)a->N_pad};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)Nchunk};
    if (int r = v4l_encode_tmap(&p.tmap_b, a->w, 2, dims, str, box, "v4l_tc_gemm(W)", nullptr)) return r;
  }
  p.B = a->B; p.Hout = a->Hout; p.Wout = a->Wout;
  p.bw = a->bw; p.bh = a->bh; p.bb = a->bb;
  p.h_tiles = (a->bb == 1) ? v4l_cdiv(a->Hout, a->bh) : 1;
  p.num_tiles = (a->bb == 1) ? a->B * p.h_tiles : v4l_cdiv(a->B, a->bb);
  p.n_taps = a->n_taps; p.kchunks = a->kchunks;
  for (int t = 0; t < a->n_taps; ++t) { p.tap_dw[t] = a->tap_dw[t]; p.tap_dh[t] = a->tap_dh[t]; }
  p.N = Nchunk; p.N_total = a->N_pad; p.N_valid = a->N_valid;
  p.bias = a->bias; p.c = a->c; p.c_map = a->c_map; p.c_f32 = a->c_f32;
  p.mask = reinterpret_cast<const __half*>(a->mask);
  p.flags = a->flags;
  p.a_idx = a->a_idx;
  p.res = reinterpret_cast<const __half*>(a->res);
  V4L_REQUIRE(!a->a_idx || a->bb == 1, "v4l_tc_gemm: a_idx needs single-item boxes (bb == 1)");

  // ring depth: the loads are latency-bound (one 16-48 KB stage per ~1 us of L2/HBM latency), so the
  // bytes in flight per SM set the feed rate of the narrow-N convolution layers
  const size_t stage_bytes = A_TILE_BYTES + (size_t)Nchunk * 128;
  p.stages = (int)max((size_t)2, min((size_t)TC_MAX_STAGES, (size_t)(196 * 1024) / stage_bytes));
  const size_t smem = (size_t)p.stages * stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const int n_chunks = a->N_pad / Nchunk;
  dim3 grid(min(p.num_tiles, max(1, ctx->sm_count / n_chunks)), n_chunks);
  V4L_LAUNCH(tc_gemm_kernel, grid, TC_THREADS, smem, (cudaStream_t)stream, p);
  V4L_CHECK_LAUNCH();
  return 0;
}

// ---- helper kernels of the tier: packing / conversion ------------------------------------------
namespace {
__global__ void pack_f16_kernel(const float* __restrict__ src, const int32_t* __restrict__ index,
                                 __half* __restrict__ dst, long long n) {
  v4l_pdl_enter();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = index ? (long long)index[i] : i;
    dst[i] = __float2half(s >= 0 ? src[s] : 0.f);
  }
}
}  // namespace

extern "C" int v4l_pack_f16(v4l_ctx* ctx, void* stream, const float* src, const int32_t* index, void* dst,
                             int64_t n) {
  V4L_REQUIRE(ctx && src && dst && n >= 0, "v4l_pack_f16: bad argument");
  if (n == 0) return 0;
  const int blocks = (int)min((long long)8 * ctx->sm_count, (long long)((n + 255) / 256));
  V4L_LAUNCH(pack_f16_kernel, blocks, 256, 0, (cudaStream_t)stream, src, index, reinterpret_cast<__half*>(dst), n);
  V4L_CHECK_LAUNCH();
  return 0;
}

// =================================================================================================
// Weight gradient on tensor cores:  D[kin (128 lanes), n] = sum_rows X_tap[row, kin] * dY[row, n]
//
// Both operands are read exactly as they sit in HBM (row = reduction index, 64 contiguous
// channels) by the same tap-shifted TMA boxes as the forward pass and consumed as MN-major
// SWIZZLE_128B operands (instruction-descriptor transpose bits), so no transposed copy of any
// activation or gradient is ever written.  A CTA owns one 128-wide slice of the packed K index
// (tap, c) and one split of the row tiles; fp32 partials go to the context scratch and
// tc_wgrad_reduce_kernel sums the splits in a fixed order and scatters into the reference-layout
// fp32 gradient through the weight-packing index table.
// =================================================================================================
namespace {

constexpr int WG_THREADS = 192;
constexpr int ATOM_BYTES = 128 * 128;           // [<=128 reduction rows][64 channels] f16

struct TcWgradParams {
  CUtensorMap tmap_x;
  CUtensorMap tmap_dy;
  int B, Hout, Wout;
  int bw, bh, bb;
  int h_tiles, num_tiles;
  int n_taps, x_C;
  int tap_dw[MAX_TAPS], tap_dh[MAX_TAPS];
  int n_atoms;                 // dY atoms of 64 channels (UMMA N = 64 * n_atoms)
  int stages;
  float* partial;              // [splits][kin_tiles][128][64 * n_atoms]
  int x_estride;               // traversal stride of X in W and H (1, or 2 for sub-sampled grids)
  int n_sub;                   // sub-iterations per tile (sub-positions of a space-to-depth cell)
  int sub_dw[4], sub_dh[4], sub_dyc[4];
  const int32_t* x_idx;
  int kin_tiles;               // K slices with real X data; blockIdx.y == kin_tiles is the bias CTA
};

__global__ void __launch_bounds__(WG_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ TcWgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[4], empty_bar[4], tmem_full;
  __shared__ uint32_t tmem_base_slot;

  v4l_pdl_trigger();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int split = blockIdx.x, kt = blockIdx.y;
  const int Nmma = 64 * p.n_atoms;
  const int box_rows = p.bw * p.bh * p.bb;
  const int ksteps = (box_rows + 15) / 16;
  const uint32_t stage_bytes = (2 + p.n_atoms) * ATOM_BYTES;
  // row tiles are dealt round-robin over the splits: the CTAs of a launch stream through one contiguous window of
  // tiles at a time (contiguous ranges per split are 2^k-strided concurrent streams at power-of-two minibatches)
  const int tile_lo = split, tile_step = gridDim.x, tile_hi = p.num_tiles;
  const int my_tiles = tile_lo < tile_hi ? (tile_hi - tile_lo + tile_step - 1) / tile_step : 0;

  // Rows of a stage beyond the TMA box (box rows not a multiple of the 16-row K step) must read as zero for
  // the whole kernel: zero just those rows of every atom, once.  The bias slice (an extra K slice whose "X" is
  // the constant 1: D = column sums of dY) reads ONE 16-row block of ones for every K step (LBO = 0: both
  // 64-lane halves see it), placed after the ring.
  const bool bias_cta = (kt == p.kin_tiles);
  uint8_t* ones = smem + p.stages * stage_bytes;
  {
    const int pad_rows = ksteps * 16 - box_rows;
    if (pad_rows > 0) {
      const int per_atom = pad_rows * 8;                          // uint4 per atom
      const int n_atoms_all = p.stages * (2 + p.n_atoms);
      for (int i = threadIdx.x; i < n_atoms_all * per_atom; i += WG_THREADS) {
        const int a = i / per_atom, o = i - a * per_atom;
        reinterpret_cast<uint4*>(smem + (a / (2 + p.n_atoms)) * stage_bytes + (a % (2 + p.n_atoms)) * ATOM_BYTES +
                                 box_rows * 128)[o] = make_uint4(0, 0, 0, 0);
      }
    }
    if (bias_cta) {
      const uint32_t one2 = 0x3C003C00u;       // two fp16 1.0
      for (int i = threadIdx.x; i < 2048 / 16; i += WG_THREADS) reinterpret_cast<uint4*>(ones)[i] = make_uint4(one2, one2, one2, one2);
    }
  }
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tmap_x);
    tc::tma_prefetch_desc(&p.tmap_dy);
    for (int s = 0; s < 4; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_slot, 256);
  tc::fence_proxy_async();          // generic-proxy zero fill -> visible to the async (TMA/UMMA) proxy
  v4l_pdl_wait();                   // smem/TMEM prologue overlapped the predecessor's tail
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // TMA producer: whole warp converged, one elected lane issues
    {
      // the two 64-channel atoms of this CTA's K slice: packed index kp = tap * x_C + c
      int a_tap[2], a_c0[2];
      for (int j = 0; j < 2; ++j) {
        const int kp0 = kt * 128 + j * 64;
        int tap = kp0 / p.x_C, c0 = kp0 - tap * p.x_C;
        if (tap >= p.n_taps) { tap = p.n_taps - 1; c0 = p.x_C; }       // beyond K: all-OOB box -> zeros
        a_tap[j] = tap; a_c0[j] = c0;
      }
      int stage = 0; uint32_t phase = 0;
      for (int tile = tile_lo; tile < tile_hi; tile += tile_step) {
        int b0, h0;
        if (p.bb == 1) { b0 = tile / p.h_tiles; h0 = (tile - b0 * p.h_tiles) * p.bh; }
        else           { b0 = tile * p.bb; h0 = 0; }
        int xb0 = b0;
        if (p.x_idx) xb0 = __shfl_sync(0xffffffffu, p.x_idx[b0], 0);
        for (int sub = 0; sub < p.n_sub; ++sub) {
          tc::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* s = smem + stage * stage_bytes;
          if (tc::elect_one()) {
            tc::mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>((bias_cta ? 0 : 2) + p.n_atoms) * box_rows * 128u);
            for (int j = 0; j < 2 && !bias_cta; ++j)
              tc::tma_load_4d(s + j * ATOM_BYTES, &p.tmap_x, &full_bar[stage], a_c0[j],
                              p.sub_dw[sub] + p.tap_dw[a_tap[j]],
                              h0 * p.x_estride + p.sub_dh[sub] + p.tap_dh[a_tap[j]], xb0);
            for (int j = 0; j < p.n_atoms; ++j)
              tc::tma_load_4d(s + (2 + j) * ATOM_BYTES, &p.tmap_dy, &full_bar[stage], p.sub_dyc[sub] + j * 64, 0,
                              h0, b0);
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = tc::umma_idesc_f16(128, Nmma, 1, 1);     // both operands MN-major
      int stage = 0; uint32_t phase = 0;
      uint32_t first = 1;
      for (int it = my_tiles * p.n_sub; it > 0; --it) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + stage * stage_bytes);
        // MN-major SW128: 16 reduction rows per MMA = 2 groups of 8 rows (SBO = 1024 B);
        // consecutive 64-channel atoms are ATOM_BYTES apart (LBO); a K step is +2048 B = +128 in the address field
        const uint64_t adesc0 = bias_cta ? tc::umma_smem_desc(tc::smem_u32(ones), 0, 1024) : tc::umma_smem_desc(sa, ATOM_BYTES, 1024);
        const uint64_t astep = bias_cta ? 0 : 128;          // the ones block serves every K step
        const uint64_t bdesc0 = tc::umma_smem_desc(sa + 2 * ATOM_BYTES, ATOM_BYTES, 1024);
        const uint32_t accum0 = first ? 0u : 1u;           // uniform: only the very first MMA overwrites
        if (tc::elect_one()) {
          for (int k = 0; k < ksteps; ++k)
            tc::umma_f16(tmem_base, adesc0 + astep * k, bdesc0 + 128 * k, idesc, k ? 1u : accum0);
          tc::umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        first = 0;
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&tmem_full);
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    // partial[split][kt][n][r]: the packed-K lane r is the fastest index, so a warp writes 128 contiguous
    // bytes per column and the reduction reads / scatters along kp (contiguous in the reference layout
    // of Linear weights)
    float* out = p.partial + (static_cast<long long>(split) * gridDim.y + kt) * 128 * Nmma + r;
    if (tile_hi > tile_lo) {
      tc::mbar_wait(&tmem_full, 0);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      for (int c0 = 0; c0 < Nmma; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(taddr + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) out[(c0 + j) * 128] = __uint_as_float(v[j]);
      }
    } else {
      for (int c = 0; c < Nmma; ++c) out[c * 128] = 0.f;
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 256);
  }
}

// The split-K reduction of these partials lives in step_ops.cu (opt_tail_kernel, phase 1): on its own
// (v4l_tc_wgrad_flush) or fused with clip + Adam + weight re-pack (v4l_opt_tail).

// column sums of a row-mapped f16 [M, N] matrix (bias gradients), two deterministic stages
__global__ void __launch_bounds__(256) colsum_f16_kernel(const __half* __restrict__ dy,
                                                          const v4l_rowmap map, int M, int N,
                                                          int rows_per_cta, float* __restrict__ part) {
  v4l_pdl_enter();
  __shared__ float red[8][256];
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  const int c = threadIdx.x & 31, w = threadIdx.x >> 5;     // 8 row-lanes x 32 column-lanes
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int m = r0 + w; m < r1; m += 8) {
    const long long a = v4l_row_addr(map, m);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int n = c + 32 * i;
      if (n < N) acc[i] += __half2float(dy[a + n]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[w][c + 32 * i] = acc[i];
  __syncthreads();
  for (int n = threadIdx.x; n < N; n += 256) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][n];
    part[(long long)blockIdx.x * N + n] = s;
  }
}
// one warp per output column: lanes stride over the partials, fixed-order shuffle reduction
__global__ void colsum_reduce_kernel(const float* __restrict__ part, int nparts, int N, int fold,
                                     float* __restrict__ out, float out_scale) {
  v4l_pdl_enter();
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float s = 0.f;
  const int total = nparts * fold;
  for (int i = lane; i < total; i += 32) {
    const int p = i / fold, f = i - p * fold;
    s += part[(long long)p * N * fold + f * N + n];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[n] = s * out_scale;
}

}  // namespace

extern "C" int v4l_tc_wgrad_flush(v4l_ctx* ctx, void* stream);

extern "C" int v4l_tc_wgrad(v4l_ctx* ctx, void* stream, const v4l_tc_wgrad_args* a) {
  V4L_REQUIRE(ctx && a && a->x && a->dy && a->dw, "v4l_tc_wgrad: NULL argument");
  V4L_REQUIRE(a->n_taps >= 1 && a->n_taps <= MAX_TAPS, "v4l_tc_wgrad: bad taps");
  V4L_REQUIRE(a->x_C % 64 == 0 && a->dy_C % 8 == 0, "v4l_tc_wgrad: x_C must be a multiple of 64, dy_C of 8");
  V4L_REQUIRE(a->N_valid >= 1 && a->N_valid <= a->dy_C && a->dy_C <= 256, "v4l_tc_wgrad: bad N");
  const int rows = a->bw * a->bh * a->bb;
  V4L_REQUIRE(rows >= 1 && rows <= 128 && a->bw == a->Wout, "v4l_tc_wgrad: bad box %dx%dx%d", a->bw, a->bh, a->bb);
  V4L_REQUIRE(a->bb == 1 || a->bh == a->Hout, "v4l_tc_wgrad: multi-item boxes must cover the whole image");
  if (a->B == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;

  TcWgradParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[4] = {(uint64_t)a->x_C, (uint64_t)a->x_W, (uint64_t)a->x_H, (uint64_t)a->x_B};
    uint64_t str[3] = {(uint64_t)a->x_C * 2, (uint64_t)a->x_C * a->x_W * 2, (uint64_t)a->x_C * a->x_W * a->x_H * 2};
    if (a->x_sW) { str[0] = (uint64_t)a->x_sW * 2; str[1] = (uint64_t)a->x_sH * 2; str[2] = (uint64_t)a->x_sB * 2; }
    uint32_t box[4] = {64, (uint32_t)a->bw, (uint32_t)a->bh, (uint32_t)a->bb};
    const uint32_t es = a->x_estride > 1 ? (uint32_t)a->x_estride : 1u;
    uint32_t estr[4] = {1, es, es, 1};
    // with a traversal stride the box extent is given in traversed elements: bw outputs span bw*es inputs
    box[1] *= es; box[2] *= es;
    if (int r = v4l_encode_tmap(&p.tmap_x, a->x, 4, dims, str, box, "v4l_tc_wgrad(X)", estr)) return r;
  }
  {
    uint64_t dims[4] = {(uint64_t)a->dy_C, (uint64_t)a->Wout, (uint64_t)a->Hout, (uint64_t)a->B};
    uint64_t str[3] = {(uint64_t)a->dy_C * 2, (uint64_t)a->dy_C * a->Wout * 2,
                       (uint64_t)a->dy_C * a->Wout * a->Hout * 2};
    if (a->dy_sW) { str[0] = (uint64_t)a->dy_sW * 2; str[1] = (uint64_t)a->dy_sH * 2; str[2] = (uint64_t)a->dy_sB * 2; }
    uint32_t box[4] = {64, (uint32_t)a->bw, (uint32_t)a->bh, (uint32_t)a->bb};
    if (int r = v4l_encode_tmap(&p.tmap_dy, a->dy, 4, dims, str, box, "v4l_tc_wgrad(dY)", nullptr)) return r;
  }
  p.B = a->B; p.Hout = a->Hout; p.Wout = a->Wout;
  p.bw = a->bw; p.bh = a->bh; p.bb = a->bb;
  p.h_tiles = (a->bb == 1) ? v4l_cdiv(a->Hout, a->bh) : 1;
  p.num_tiles = (a->bb == 1) ? a->B * p.h_tiles : v4l_cdiv(a->B, a->bb);
  p.n_taps = a->n_taps; p.x_C = a->x_C;
  for (int t = 0; t < a->n_taps; ++t) { p.tap_dw[t] = a->tap_dw[t]; p.tap_dh[t] = a->tap_dh[t]; }
  p.n_atoms = v4l_cdiv(a->N_valid, 64);
  p.x_estride = a->x_estride > 1 ? a->x_estride : 1;
  p.n_sub = a->n_sub > 0 ? a->n_sub : 1;
  V4L_REQUIRE(p.n_sub <= 4, "v4l_tc_wgrad: at most 4 sub-iterations");
  for (int i = 0; i < p.n_sub; ++i) {
    p.sub_dw[i] = a->n_sub > 0 ? a->sub_dw[i] : 0;
    p.sub_dh[i] = a->n_sub > 0 ? a->sub_dh[i] : 0;
    p.sub_dyc[i] = a->n_sub > 0 ? a->sub_dyc[i] : 0;
  }
  p.x_idx = a->x_idx;
  V4L_REQUIRE(!a->x_idx || a->bb == 1, "v4l_tc_wgrad: x_idx needs single-item boxes (bb == 1)");
  const int Nmma = 64 * p.n_atoms;
  const int Kp = a->n_taps * a->x_C;
  const int kin_tiles = v4l_cdiv(Kp, 128);
  const size_t stage_bytes = (size_t)(2 + p.n_atoms) * ATOM_BYTES;
  p.stages = (int)min((size_t)4, (size_t)(200 * 1024 - 1024 - 2048) / stage_bytes);
  const int has_bias = a->dbias ? 1 : 0;
  const int ytiles = kin_tiles + has_bias;
  p.kin_tiles = kin_tiles;
  // deferred reductions keep their partials in the upper half of the scratch until the flush
  // split-K over the row tiles: >= 8 row tiles per CTA (the partial sums cost 128 x Nmma floats of
  // traffic per CTA, twice), at most one wave; several of these launches run side by side
  int splits = max(1, min(p.num_tiles / 8, min(48, ctx->sm_count / ytiles)));
  // a deferred job that cannot get its split count from what is left of the scratch flushes the pending jobs first
  // (running on the few splits that still fit would serialise the launch on a handful of SMs)
  if (a->defer && ctx->n_jobs > 0 && (ctx->n_jobs == V4L_MAX_JOBS ||
                   ctx->defer_elems - ctx->defer_cursor < (size_t)splits * ytiles * 128 * Nmma)) {
    if (int r = v4l_tc_wgrad_flush(ctx, stream)) return r;
    ctx->early_flush = 1;
    ++ctx->early_flush_count;
  }
  float* region = a->defer ? ctx->defer_base + ctx->defer_cursor : ctx->scratch;
  const size_t avail = a->defer ? ctx->defer_elems - ctx->defer_cursor : ctx->scratch_elems;
  splits = (int)min((size_t)splits, avail / ((size_t)ytiles * 128 * Nmma));
  V4L_REQUIRE(splits >= 1, "v4l_tc_wgrad: scratch too small (flush deferred reductions more often)");
  splits = v4l_cdiv(p.num_tiles, v4l_cdiv(p.num_tiles, splits));     // no split without a tile
  p.partial = region;

  static bool attr_set = false;
  if (!attr_set) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  const size_t smem = (size_t)p.stages * stage_bytes + 2048 + 1024;      // ring + block of ones + alignment slack
  V4L_LAUNCH(tc_wgrad_kernel, dim3(splits, ytiles), WG_THREADS, smem, s, p);
  V4L_CHECK_LAUNCH();
  v4l_reduce_job job;
  job.partial = region; job.index = a->index; job.dw = a->dw; job.dbias = a->dbias;
  job.splits = splits; job.kin_tiles = kin_tiles; job.has_bias = has_bias; job.Nmma = Nmma;
  job.N_valid = a->N_valid; job.Kp = Kp; job.scale = a->out_scale != 0.f ? a->out_scale : 1.f;
  job.accumulate = a->accumulate ? 1 : 0;
  if (a->defer) {
    ctx->jobs[ctx->n_jobs++] = job;
    ctx->defer_cursor += (((size_t)splits * ytiles * 128 * Nmma) + 63) / 64 * 64;
    return 0;
  }
  // immediate reduction: the pending deferred jobs (if any) are set aside around the one-job pass
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

// split-K reduction of every pending (deferred) weight-gradient job in one launch: phase 1 of the
// optimiser tail kernel (step_ops.cu) on its own
extern "C" int v4l_tc_wgrad_flush(v4l_ctx* ctx, void* stream) {
  V4L_REQUIRE(ctx, "v4l_tc_wgrad_flush: NULL ctx");
  if (ctx->n_jobs == 0) return 0;
  v4l_opt_tail_args t;
  memset(&t, 0, sizeof(t));
  t.phases = 1;
  t.norm_slot = -1;
  return v4l_opt_tail(ctx, stream, &t);
}

extern "C" int v4l_colsum_f16(v4l_ctx* ctx, void* stream, const void* dy, const v4l_rowmap* map, int M, int N,
                              int fold, float out_scale, float* out) {
  V4L_REQUIRE(ctx && dy && map && out && map->P > 0, "v4l_colsum_f16: bad argument");
  V4L_REQUIRE(fold >= 1 && N >= 1 && N * fold <= 256 && M >= 1, "v4l_colsum_f16: bad shape M=%d N=%d fold=%d", M, N, fold);
  const int Nout = N;
  N = N * fold;
  cudaStream_t s = (cudaStream_t)stream;
  int ctas = min(ctx->sm_count, v4l_cdiv(M, 64));
  const int rpc = v4l_cdiv(M, ctas);
  ctas = v4l_cdiv(M, rpc);
  // partials live past the region v4l_tc_wgrad uses? no: separate calls are stream-ordered
  float* part = ctx->scratch;
  V4L_REQUIRE((size_t)ctas * N <= ctx->scratch_elems, "v4l_colsum_f16: scratch too small");
  V4L_LAUNCH(colsum_f16_kernel, ctas, 256, 0, s, reinterpret_cast<const __half*>(dy), *map, M, N, rpc, part);
  V4L_CHECK_LAUNCH();
  V4L_LAUNCH(colsum_reduce_kernel, v4l_cdiv(Nout, 8), 256, 0, s, part, ctas, Nout, fold, out, out_scale);
  V4L_CHECK_LAUNCH();
  return 0;
}

// =================================================================================================
// Layout helpers of the tensor-core tier
// =================================================================================================
namespace {

// fp32 CHW [4,64,64] observation image -> f16 4x4 space-to-depth NHWC [16,16,64],
// channel = (py*4 + px)*4 + c for source pixel (4Y+py, 4X+px): the 8x8/4 conv becomes a 2x2/1
// conv with 64-channel (128-byte) rows — exactly one TMA/UMMA swizzle atom per tap.
// fp16 source (the replay buffer's half-precision staging copy of the depth stack): same re-ordering,
// the fp32 -> fp16 rounding already happened on the host (round-to-nearest, identical values)
__global__ void ingest_img_f16_kernel(const __half* __restrict__ img, __half* __restrict__ out, long long n_img,
                                      const int32_t* __restrict__ idx) {
  v4l_pdl_enter();
  const long long total = n_img * 16 * 4 * 16;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(t & 15);
    const int py = (int)((t >> 4) & 3);
    const int Y = (int)((t >> 6) & 15);
    const long long n = idx ? (long long)idx[t >> 10] : (t >> 10);
    const __half* src = img + n * 16384 + (4 * Y + py) * 64 + 4 * X;
    unsigned short h[4][4];                        // [c][px]
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const uint2 v = *reinterpret_cast<const uint2*>(src + c * 4096);
      h[c][0] = v.x & 0xffffu; h[c][1] = v.x >> 16; h[c][2] = v.y & 0xffffu; h[c][3] = v.y >> 16;
    }
    uint32_t w[8];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      w[2 * px] = (uint32_t)h[0][px] | ((uint32_t)h[1][px] << 16);
      w[2 * px + 1] = (uint32_t)h[2][px] | ((uint32_t)h[3][px] << 16);
    }
    uint4* dst = reinterpret_cast<uint4*>(out + ((n * 16 + Y) * 16 + X) * 64 + py * 16);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

__global__ void ingest_img_kernel(const float* __restrict__ img, __half* __restrict__ out, long long n_img,
                                  const int32_t* __restrict__ idx) {
  v4l_pdl_enter();
  const long long total = n_img * 16 * 4 * 16;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(t & 15);
    const int py = (int)((t >> 4) & 3);
    const int Y = (int)((t >> 6) & 15);
    const long long n = idx ? (long long)idx[t >> 10] : (t >> 10);    // optional row list (streamed ingest)
    const float* src = img + n * 16384 + (4 * Y + py) * 64 + 4 * X;
    float4 v[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) v[c] = *reinterpret_cast<const float4*>(src + c * 4096);
    uint32_t w[8];
    const float* f0 = reinterpret_cast<const float*>(&v[0]);
    // out order: px (4) x c (4); v[c] holds px = 0..3 for channel c
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const float a0 = reinterpret_cast<const float*>(&v[0])[px], a1 = reinterpret_cast<const float*>(&v[1])[px];
      const float a2 = reinterpret_cast<const float*>(&v[2])[px], a3 = reinterpret_cast<const float*>(&v[3])[px];
      w[2 * px] = pack_h16(a0, a1);
      w[2 * px + 1] = pack_h16(a2, a3);
    }
    (void)f0;
    uint4* dst = reinterpret_cast<uint4*>(out + ((n * 16 + Y) * 16 + X) * 64 + py * 16);
    dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
    dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

// dst[i, :] = src[idx ? idx[i] : i, :] with fp32 or f16 source, f16 destination, zero padded to dcols
template <typename S>
__global__ void gather_rows_kernel(const S* __restrict__ src, const int32_t* __restrict__ idx,
                                   __half* __restrict__ dst, int rows, int scols, long long sstride,
                                   int dcols, float scale) {
  v4l_pdl_enter();
  const long long total = (long long)rows * dcols;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(e / dcols), c = (int)(e - (long long)r * dcols);
    float v = 0.f;
    if (c < scols) {
      const long long sr = idx ? (long long)idx[r] : (long long)r;
      v = static_cast<float>(src[sr * sstride + c]) * scale;
    }
    dst[e] = __float2half(v);
  }
}

__global__ void relu_bwd_f16_kernel(const __half* __restrict__ dy, const v4l_rowmap dy_map,
                                     const __half* __restrict__ act, const v4l_rowmap act_map,
                                     __half* __restrict__ out, const v4l_rowmap out_map, int M, int N) {
  v4l_pdl_enter();
  const long long total = (long long)M * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / N), n = (int)(e - (long long)m * N);
    const __half g = dy[v4l_row_addr(dy_map, m) + n];
    const float a = __half2float(act[v4l_row_addr(act_map, m) + n]);
    out[v4l_row_addr(out_map, m) + n] = a > 0.f ? g : __float2half(0.f);
  }
}

}  // namespace

extern "C" int v4l_ingest_img(v4l_ctx* ctx, void* stream, const float* img, void* out_s2d, int64_t n_img,
                              const int32_t* idx) {
  V4L_REQUIRE(ctx && img && out_s2d && n_img >= 0, "v4l_ingest_img: bad argument");
  V4L_REQUIRE(((reinterpret_cast<uintptr_t>(img) | reinterpret_cast<uintptr_t>(out_s2d)) & 15) == 0,
              "v4l_ingest_img: img / out_s2d must be 16-byte aligned (vector loads)");
  if (n_img == 0) return 0;
  const long long total = n_img * 1024;
  const int blocks = (int)min((long long)16 * ctx->sm_count, (total + 255) / 256);
  V4L_LAUNCH(ingest_img_kernel, blocks, 256, 0, (cudaStream_t)stream, img, reinterpret_cast<__half*>(out_s2d), n_img, idx);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_ingest_img_f16(v4l_ctx* ctx, void* stream, const void* img_f16, void* out_s2d, int64_t n_img,
                                  const int32_t* idx) {
  V4L_REQUIRE(ctx && img_f16 && out_s2d && n_img >= 0, "v4l_ingest_img_f16: bad argument");
  V4L_REQUIRE(((reinterpret_cast<uintptr_t>(img_f16) | reinterpret_cast<uintptr_t>(out_s2d)) & 15) == 0,
              "v4l_ingest_img_f16: img / out_s2d must be 16-byte aligned (vector loads)");
  if (n_img == 0) return 0;
  const long long total = n_img * 1024;
  const int blocks = (int)min((long long)16 * ctx->sm_count, (total + 255) / 256);
  V4L_LAUNCH(ingest_img_f16_kernel, blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const __half*>(img_f16),
             reinterpret_cast<__half*>(out_s2d), n_img, idx);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_gather_rows_f16(v4l_ctx* ctx, void* stream, const void* src, int src_is_f32,
                                   const int32_t* idx, void* dst, int rows, int src_cols, int64_t src_stride,
                                   int dst_cols, float scale) {
  V4L_REQUIRE(ctx && src && dst && rows >= 0 && src_cols >= 0 && dst_cols >= src_cols, "v4l_gather_rows_f16: bad argument");
  const long long total = (long long)rows * dst_cols;
  if (total == 0) return 0;
  const int blocks = (int)min((long long)8 * ctx->sm_count, (total + 255) / 256);
  cudaStream_t s = (cudaStream_t)stream;
  if (src_is_f32)
    V4L_LAUNCH((gather_rows_kernel<float>), blocks, 256, 0, s, (const float*)src, idx, (__half*)dst, rows, src_cols, src_stride, dst_cols, scale);
  else
    V4L_LAUNCH((gather_rows_kernel<__half>), blocks, 256, 0, s, (const __half*)src, idx, (__half*)dst, rows, src_cols, src_stride, dst_cols, scale);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_relu_bwd_f16(v4l_ctx* ctx, void* stream, const void* dy, const v4l_rowmap* dy_map,
                                 const void* act, const v4l_rowmap* act_map, void* out,
                                 const v4l_rowmap* out_map, int M, int N) {
  V4L_REQUIRE(ctx && dy && dy_map && act && act_map && out && out_map, "v4l_relu_bwd_f16: NULL argument");
  V4L_REQUIRE(dy_map->P > 0 && act_map->P > 0 && out_map->P > 0, "v4l_relu_bwd_f16: row map with P <= 0");
  const long long total = (long long)M * N;
  if (total <= 0) return 0;
  const int blocks = (int)min((long long)8 * ctx->sm_count, (total + 255) / 256);
  V4L_LAUNCH(relu_bwd_f16_kernel, blocks, 256, 0, (cudaStream_t)stream, 
    (const __half*)dy, *dy_map, (const __half*)act, *act_map, (__half*)out, *out_map, M, N);
  V4L_CHECK_LAUNCH();
  return 0;
}

// =================================================================================================
// Zero-copy rollout ingest: one CTA per observation row reads it straight from PINNED HOST memory
// (UVA pointer, PCIe reads issued by the SMs, fully coalesced 4-byte loads so the 4*S-byte offset of
// the image inside the row needs no alignment) and writes the device-side layouts in one pass:
// proprio plane fp32 [N,S], optional fp32 image plane [N,16384] (exact tier) and the fp16
// space-to-depth image [N,16,16,64] (tensor-core tier).  Replaces a 1 GB staging copy + a
// conversion pass, and lets the host->device stream follow the minibatch order of the first
// opt-epoch with ONE launch per minibatch (row list = that minibatch's indices).
// =================================================================================================
namespace {
__global__ void __launch_bounds__(256) ingest_rows_kernel(const float* __restrict__ obs, long long row_stride, int S,
                                                          const int32_t* __restrict__ idx, float* __restrict__ state_out,
                                                          float* __restrict__ img_out, __half* __restrict__ s2d_out) {
  v4l_pdl_enter();
  __shared__ __half sm[16384];
  const long long n = idx ? (long long)idx[blockIdx.x] : (long long)blockIdx.x;
  const float* src = obs + n * row_stride;
  if (state_out)
    for (int i = threadIdx.x; i < S; i += 256) state_out[n * S + i] = src[i];
  const float* im = src + S;
#pragma unroll 8
  for (int j = 0; j < 64; ++j) {
    const int e = j * 256 + threadIdx.x;
    const float v = im[e];
    if (img_out) img_out[n * 16384 + e] = v;
    sm[e] = __float2half(v);
  }
  if (!s2d_out) return;
  __syncthreads();
  // 2048 chunks of 8 halfs: chunk = (Y, X, py, hi) -> pixels px = 2*hi, 2*hi+1, channels 0..3
  __half* dst = s2d_out + n * 16384;
  for (int q = threadIdx.x; q < 2048; q += 256) {
    const int hi = q & 1, py = (q >> 1) & 3, X = (q >> 3) & 15, Y = q >> 7;
    const int base = (4 * Y + py) * 64 + 4 * X + 2 * hi;
    __half h[8];
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int c = 0; c < 4; ++c) h[px * 4 + c] = sm[c * 4096 + base + px];
    *reinterpret_cast<uint4*>(dst + (Y * 16 + X) * 64 + py * 16 + hi * 8) = *reinterpret_cast<const uint4*>(h);
  }
}
}  // namespace

extern "C" int v4l_ingest_rows(v4l_ctx* ctx, void* stream, const float* obs, int64_t row_stride, int S,
                               const int32_t* idx, int64_t n_rows, float* state_out, float* img_out,
                               void* s2d_out) {
  V4L_REQUIRE(ctx && obs && n_rows >= 0 && S >= 0 && row_stride >= S + 16384, "v4l_ingest_rows: bad argument");
  V4L_REQUIRE(S == 0 || state_out, "v4l_ingest_rows: state_out is NULL");
  if (n_rows == 0) return 0;
  V4L_LAUNCH(ingest_rows_kernel, (unsigned)n_rows, 256, 0, (cudaStream_t)stream, obs, (long long)row_stride, S, idx,
             state_out, img_out, reinterpret_cast<__half*>(s2d_out));
  return 0;
}
