// Fused forward of one nn.TransformerEncoderLayer(64, 1 head, ff 256, dropout 0, post-norm, ReLU)
// on tcgen05 — the LocoTransformer "attention block" (reference torchrl/networks/nets.py:949-955,
// 1009-1011; math: SURVEY Appendix A2).  One CTA owns a 128-row tile = 7 samples x 17 tokens (or
// 8 x 16) and runs the whole layer without leaving the SM:
//
//   x --QKV GEMM--> [q|k|v] --S=QK^T--> masked softmax --O=PV--> out-proj --(+x) LN1--> h
//     --FFN1 (ReLU)--> f1 --FFN2--> (+h) LN2 --> y
//
// All six contractions are tcgen05.mma with fp32 accumulators in TMEM; the layer's weights
// (96 KB fp16) are TMA-loaded into shared memory once per CTA; every intermediate operand is
// written by the epilogue warps straight into the 128-byte-swizzled shared-memory layout the next
// MMA consumes (K-major A tiles; V is consumed MN-major).  LayerNorm, softmax, bias, residual and
// ReLU happen in registers (thread = token row).  The tensors the backward pass needs
// (qkv, P, o, z1/stats1, h, f1, z2/stats2) are stored to HBM on the way.
// 160 threads: warp 0 = TMEM alloc + TMA + MMA issue, warps 1-4 = epilogues (TMEM lane quadrants).
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int BK_THREADS = 160;

// phase timeline of CTA 0 (globaltimer ns), [0] forward kernel, [1] data-gradient kernel; read back by
// v4l_tc_block_timeline().  One predicated store per phase: free next to the phases themselves.
__device__ unsigned long long g_timeline[2][32];
__device__ __forceinline__ void stamp(int which, int slot, bool on) {
  if (on) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_timeline[which][slot] = t;
  }
}
constexpr int TB = 128 * 128;              // one [128 rows][64 fp16] swizzled tile

struct BlockParams {
  CUtensorMap tm_x;        // [R, 64]   box {64, rows_per_tile}
  CUtensorMap tm_win;      // [192, 64] box {64, 192}
  CUtensorMap tm_wo;       // [64, 64]  box {64, 64}
  CUtensorMap tm_w1;       // [256, 64] box {64, 256}
  CUtensorMap tm_w2;       // [64, 256] box {64, 64} (4 k-chunks)
  CUtensorMap tm_qkv_o, tm_o_o, tm_h_o, tm_f1_o;   // outputs [R,192] [R,64] [R,64] [R,256], box {64, rows_per_tile}
  int R, T, rows_per_tile;
  float scale, eps;
  const float *b_in, *b_o, *g1, *be1, *b1, *b2, *g2, *be2;
  __half *qkv, *o, *h, *f1, *y;
  float *p, *z1, *st1, *z2, *st2;
  __half *xh1, *xh2;
};

// shared-memory map (bytes)
constexpr int OFF_WIN = 0;                 // 192 x 128 B = 24 KB
constexpr int OFF_WO = OFF_WIN + 192 * 128;        //  8 KB
constexpr int OFF_W1 = OFF_WO + 64 * 128;          // 32 KB
constexpr int OFF_W2 = OFF_W1 + 256 * 128;         // 4 x 8 KB
constexpr int OFF_X = OFF_W2 + 4 * 64 * 128;       // x, later h      16 KB
constexpr int OFF_O = OFF_X + TB;                  // attention output 16 KB
constexpr int OFF_QKV = OFF_O + TB;                // q | k | v tiles  48 KB  \ reused by f1 (4 tiles, 64 KB)
constexpr int OFF_P = OFF_QKV + 3 * TB;            // P, 2 tiles       32 KB  /
constexpr int SMEM_BYTES = OFF_P + 2 * TB;         // 212992
constexpr int OFF_F1 = OFF_QKV;

__device__ __forceinline__ void st_sw(uint8_t* tile, int row, int chunk, uint4 v) {
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}
__device__ __forceinline__ uint4 ld_sw(const uint8_t* tile, int row, int chunk) {
  return *reinterpret_cast<const uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ float ex2_approx(float x) {     // one MUFU.EX2 (results below 2^-126 flush to 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pk2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 up2(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 w;
  w.x = pk2(f[0], f[1]); w.y = pk2(f[2], f[3]); w.z = pk2(f[4], f[5]); w.w = pk2(f[6], f[7]);
  return w;
}

// Walk NCOLS accumulator columns (a multiple of 64) in 32-column chunks with the TMEM load of the next
// chunk in flight while the current one is processed: fn(c0, v) gets the 32 fp32 bit patterns.
template <int NCOLS, class F>
__device__ __forceinline__ void tmem_walk(uint32_t taddr, F&& fn) {
  uint32_t va[32], vb[32];
  tc::tmem_ld_32x32(taddr, va);
#pragma unroll
  for (int c0 = 0; c0 < NCOLS; c0 += 64) {
    tc::tmem_ld_wait();
    tc::tmem_ld_32x32(taddr + c0 + 32, vb);
    fn(c0, va);
    tc::tmem_ld_wait();
    if (c0 + 64 < NCOLS) tc::tmem_ld_32x32(taddr + c0 + 64, va);
    fn(c0 + 32, vb);
  }
}

// Attention-score chunks: of the 128 accumulator columns at taddr, read the 32-column chunks
// [cb, ce) that hold this warp's samples (TMEM loads are warp-wide; next load in flight while the
// current chunk is looked at) and keep, per thread, the chunk holding the first (ownA) and the last
// (ownB, may equal ownA) key column of its own sample.
__device__ __forceinline__ void tmem_own_chunks(uint32_t taddr, int cb, int ce, int ownA, int ownB,
                                                uint32_t (&ka)[32], uint32_t (&kb)[32]) {
  uint32_t va[32], vb[32];
  tc::tmem_ld_32x32(taddr + cb, va);
#pragma unroll
  for (int i = 0; i < 4; i += 2) {
    const int c0 = cb + 32 * i;
    if (c0 < ce) {
      tc::tmem_ld_wait();
      if (c0 + 32 < ce) tc::tmem_ld_32x32(taddr + c0 + 32, vb);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (c0 == ownA) ka[j] = va[j];
        if (c0 == ownB) kb[j] = va[j];
      }
      if (c0 + 32 < ce) {
        tc::tmem_ld_wait();
        if (c0 + 64 < ce) tc::tmem_ld_32x32(taddr + c0 + 64, va);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (c0 + 32 == ownA) ka[j] = vb[j];
          if (c0 + 32 == ownB) kb[j] = vb[j];
        }
      }
    }
  }
}

__global__ void __launch_bounds__(BK_THREADS, 1) tc_block_fwd_kernel(const __grid_constant__ BlockParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_in, bar_w, bar_m[6], bar_e[5];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_par[192 + 64 * 6 + 256];   // b_in | b_o | g1 | be1 | b1(256) | b2 | g2 | be2

  v4l_pdl_trigger();
  const bool tl = blockIdx.x == 0 && threadIdx.x == 32;
  stamp(0, 0, tl);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int row0 = blockIdx.x * p.rows_per_tile;
  {  // zero the activation tiles (padding rows of a tile must be exact zeros / finite)
    uint4* z = reinterpret_cast<uint4*>(sm + OFF_X);
    for (int i = threadIdx.x; i < (SMEM_BYTES - OFF_X) / 16; i += BK_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tm_x); tc::tma_prefetch_desc(&p.tm_win); tc::tma_prefetch_desc(&p.tm_wo);
    tc::tma_prefetch_desc(&p.tm_w1); tc::tma_prefetch_desc(&p.tm_w2);
    tc::tma_prefetch_desc(&p.tm_qkv_o); tc::tma_prefetch_desc(&p.tm_o_o); tc::tma_prefetch_desc(&p.tm_h_o);
    tc::tma_prefetch_desc(&p.tm_f1_o);
    tc::mbar_init(&bar_in, 1); tc::mbar_init(&bar_w, 1);
    for (int i = 0; i < 6; ++i) tc::mbar_init(&bar_m[i], 1);
    for (int i = 0; i < 5; ++i) tc::mbar_init(&bar_e[i], 128);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_proxy_async();
  v4l_pdl_wait();
  // parameters -> smem (after the wait: the optimiser step of the previous launch wrote them)
  for (int i = threadIdx.x; i < 192; i += BK_THREADS) s_par[i] = p.b_in[i];
  for (int i = threadIdx.x; i < 64; i += BK_THREADS) {
    s_par[192 + i] = p.b_o[i]; s_par[256 + i] = p.g1[i]; s_par[320 + i] = p.be1[i];
    s_par[640 + i] = p.b2[i]; s_par[704 + i] = p.g2[i]; s_par[768 + i] = p.be2[i];
  }
  for (int i = threadIdx.x; i < 256; i += BK_THREADS) s_par[384 + i] = p.b1[i];
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  stamp(0, 1, tl);
  const float *sb_in = s_par, *sb_o = s_par + 192, *sg1 = s_par + 256, *sbe1 = s_par + 320, *sb1 = s_par + 384,
              *sb2 = s_par + 640, *sg2 = s_par + 704, *sbe2 = s_par + 768;
  // TMEM columns: qkv [0,192)  S [192,320)  O [320,384)  proj [384,448)  f1 [0,256) (after qkv is consumed)
  //               f2 [448,512)
  constexpr uint32_t C_QKV = 0, C_S = 192, C_O = 320, C_PROJ = 384, C_F1 = 0, C_F2 = 448;

  if (warp == 0) {
    // converged warp, one elected lane issues: operands of UTCHMMA / UTMALDG stay in uniform registers
    if (tc::elect_one()) {
      const uint32_t base = tc::smem_u32(sm);
      tc::mbar_expect_tx(&bar_in, static_cast<uint32_t>(p.rows_per_tile) * 128u + 192u * 128u);
      tc::tma_load_2d(sm + OFF_X, &p.tm_x, &bar_in, 0, row0);
      tc::tma_load_2d(sm + OFF_WIN, &p.tm_win, &bar_in, 0, 0);
      tc::mbar_expect_tx(&bar_w, (64u + 256u + 256u) * 128u);
      tc::tma_load_2d(sm + OFF_WO, &p.tm_wo, &bar_w, 0, 0);
      tc::tma_load_2d(sm + OFF_W1, &p.tm_w1, &bar_w, 0, 0);
      for (int c = 0; c < 4; ++c) tc::tma_load_2d(sm + OFF_W2 + c * 64 * 128, &p.tm_w2, &bar_w, c * 64, 0);
      // (1) qkv = x Win^T
      tc::mbar_wait(&bar_in, 0);
      tc::tc_fence_after();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 192, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_f16(tmem + C_QKV, tc::umma_smem_desc(base + OFF_X + k * 32, 0, 1024),
                       tc::umma_smem_desc(base + OFF_WIN + k * 32, 0, 1024), id, k ? 1u : 0u);
        tc::umma_commit(&bar_m[0]);
      }
      // (2) S = Q K^T   (the q|k|v tiles leave for HBM by TMA meanwhile)
      tc::mbar_wait(&bar_e[0], 0);
      tc::tc_fence_after();
      for (int c = 0; c < 3; ++c) tc::tma_store_2d(&p.tm_qkv_o, sm + OFF_QKV + c * TB, c * 64, row0);
      tc::tma_store_commit();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 128, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_f16(tmem + C_S, tc::umma_smem_desc(base + OFF_QKV + k * 32, 0, 1024),
                       tc::umma_smem_desc(base + OFF_QKV + TB + k * 32, 0, 1024), id, k ? 1u : 0u);
        tc::umma_commit(&bar_m[1]);
      }
      // (3) O = P V (V MN-major)
      tc::mbar_wait(&bar_e[1], 0);
      tc::tc_fence_after();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 64, 0, 1);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          tc::umma_f16(tmem + C_O, tc::umma_smem_desc(base + OFF_P + (k >> 2) * TB + (k & 3) * 32, 0, 1024),
                       tc::umma_smem_desc(base + OFF_QKV + 2 * TB + k * 2048, TB, 1024), id, k ? 1u : 0u);
        tc::umma_commit(&bar_m[2]);
      }
      // (4) proj = O Wo^T
      tc::mbar_wait(&bar_e[2], 0);
      tc::mbar_wait(&bar_w, 0);
      tc::tc_fence_after();
      tc::tma_store_2d(&p.tm_o_o, sm + OFF_O, 0, row0);
      tc::tma_store_commit();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_f16(tmem + C_PROJ, tc::umma_smem_desc(base + OFF_O + k * 32, 0, 1024),
                       tc::umma_smem_desc(base + OFF_WO + k * 32, 0, 1024), id, k ? 1u : 0u);
        tc::umma_commit(&bar_m[3]);
      }
      // (5) f1 = h W1^T   (h lives where x was)
      tc::mbar_wait(&bar_e[3], 0);
      tc::tc_fence_after();
      tc::tma_store_wait_read();            // q|k|v have left: the f1 epilogue may overwrite their tiles
      tc::tma_store_2d(&p.tm_h_o, sm + OFF_X, 0, row0);
      tc::tma_store_commit();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 256, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_f16(tmem + C_F1, tc::umma_smem_desc(base + OFF_X + k * 32, 0, 1024),
                       tc::umma_smem_desc(base + OFF_W1 + k * 32, 0, 1024), id, k ? 1u : 0u);
        tc::umma_commit(&bar_m[4]);
      }
      // (6) f2 = f1 W2^T  (K = 256: 4 tiles of A, 4 k-chunk tiles of W2)
      tc::mbar_wait(&bar_e[4], 0);
      tc::tc_fence_after();
      for (int c = 0; c < 4; ++c) tc::tma_store_2d(&p.tm_f1_o, sm + OFF_F1 + c * TB, c * 64, row0);
      tc::tma_store_commit();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          tc::umma_f16(tmem + C_F2, tc::umma_smem_desc(base + OFF_F1 + (k >> 2) * TB + (k & 3) * 32, 0, 1024),
                       tc::umma_smem_desc(base + OFF_W2 + (k >> 2) * 64 * 128 + (k & 3) * 32, 0, 1024), id,
                       k ? 1u : 0u);
        tc::umma_commit(&bar_m[5]);
      }
      tc::tma_store_wait_all();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int s_loc = r / p.T;
    const int lo = s_loc * p.T, hi = lo + p.T;
    const int grow = row0 + r;
    const bool live = (r < p.rows_per_tile) && (grow < p.R);
    const uint32_t ta = tmem + (static_cast<uint32_t>(quad * 32) << 16);

    // ---- (1) qkv epilogue: + bias -> fp16 -> q/k/v tiles + global
    tc::mbar_wait(&bar_m[0], 0);
    stamp(0, 3, tl);
    tc::tc_fence_after();
    tmem_walk<192>(ta + C_QKV, [&](int c0, const uint32_t (&v)[32]) {
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = live ? __uint_as_float(v[j]) + sb_in[c0 + j] : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + 8 * q;
        st_sw(sm + OFF_QKV + (col >> 6) * TB, r, (col & 63) >> 3, pack8(f + 8 * q));
      }
    });
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[0]);
    stamp(0, 2, tl);

    // ---- (2) masked softmax over the sample's own keys -> P (unnormalised, fp16) tiles
    tc::mbar_wait(&bar_m[1], 0);
    stamp(0, 5, tl);
    tc::tc_fence_after();
    // one pass: the thread's own score chunk(s) are kept in registers; the rest of P stays zero (prologue)
    const int wlo = ((quad * 32) / p.T) * p.T;
    const int whi = ((quad * 32 + 31) / p.T + 1) * p.T;
    const int cb = wlo & ~31, ce = min(128, (whi + 31) & ~31);
    const int ownA = min(lo, 127) & ~31, ownB = min(hi - 1, 127) & ~31;
    float sum = 0.f;
    {
      uint32_t ka[32], kb[32];
      tmem_own_chunks(ta + C_S, cb, ce, ownA, ownB, ka, kb);
      const bool two = ownB != ownA;
      float mx = -3.0e38f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (ownA + j >= lo && ownA + j < hi) mx = fmaxf(mx, __uint_as_float(ka[j]));
        if (two && ownB + j < hi) mx = fmaxf(mx, __uint_as_float(kb[j]));
      }
      const float sc2 = p.scale * 1.4426950408889634f;      // exp(x * scale) = 2^(x * scale * log2 e)
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        // exp unconditionally on a clamped argument, then select: a per-element branch around the
        // MUFU costs far more than the exp itself
        const bool in = live && (ownA + j >= lo) && (ownA + j < hi);
        const float t = ex2_approx(fminf((__uint_as_float(ka[j]) - mx) * sc2, 0.f));
        e[j] = in ? t : 0.f;
        sum += e[j];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = ownA + 8 * q;
        st_sw(sm + OFF_P + (col >> 6) * TB, r, (col & 63) >> 3, pack8(e + 8 * q));
      }
      if (two) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const bool in = live && (ownB + j < hi);
          const float t = ex2_approx(fminf((__uint_as_float(kb[j]) - mx) * sc2, 0.f));
          e[j] = in ? t : 0.f;
          sum += e[j];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = ownB + 8 * q;
          st_sw(sm + OFF_P + (col >> 6) * TB, r, (col & 63) >> 3, pack8(e + 8 * q));
        }
      }
    }
    const float inv = live ? 1.f / sum : 0.f;
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[1]);
    stamp(0, 4, tl);
    if (live) {     // normalised probabilities of the block -> global (backward)
      float* prow = p.p + (long long)grow * p.T;
      for (int col = lo; col < hi; ++col) {
        const __half hv = *reinterpret_cast<const __half*>(sm + OFF_P + (col >> 6) * TB + r * 128 +
                                                           ((((col & 63) >> 3) ^ (r & 7)) << 4) + (col & 7) * 2);
        prow[col - lo] = __half2float(hv) * inv;
      }
    }

    // ---- (3) O epilogue: / sum -> fp16 -> O tile + global
    tc::mbar_wait(&bar_m[2], 0);
    stamp(0, 7, tl);
    tc::tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(ta + C_O + c0, v);
      tc::tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * inv;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        st_sw(sm + OFF_O, r, (c0 >> 3) + q, pack8(f + 8 * q));
      }
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[2]);
    stamp(0, 6, tl);

    // ---- (4) out-proj epilogue: + bias + x -> LayerNorm1 -> h (tile over x, global), z1, stats1
    tc::mbar_wait(&bar_m[3], 0);
    stamp(0, 9, tl);
    tc::tc_fence_after();
    {
      float zrow[64];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(ta + C_PROJ + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 xr = ld_sw(sm + OFF_X, r, (c0 >> 3) + q);
          const uint32_t xw[4] = {xr.x, xr.y, xr.z, xr.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 xf = up2(xw[j]);
            const int c = c0 + 8 * q + 2 * j;
            zrow[c] = __uint_as_float(v[8 * q + 2 * j]) + sb_o[c] + xf.x;
            zrow[c + 1] = __uint_as_float(v[8 * q + 2 * j + 1]) + sb_o[c + 1] + xf.y;
          }
        }
      }
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) s += zrow[c];
      const float mean = s * (1.f / 64.f);
      float vs = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) { const float d = zrow[c] - mean; vs = fmaf(d, d, vs); }
      const float rstd = rsqrtf(vs * (1.f / 64.f) + p.eps);
      if (live) {
        if (p.z1) {
          float4* zg = reinterpret_cast<float4*>(p.z1 + (long long)grow * 64);
#pragma unroll
          for (int c = 0; c < 64; c += 4) zg[c >> 2] = make_float4(zrow[c], zrow[c + 1], zrow[c + 2], zrow[c + 3]);
        }
        *reinterpret_cast<float2*>(p.st1 + (long long)grow * 2) = make_float2(mean, rstd);
      }
#pragma unroll
      for (int c = 0; c < 64; ++c) zrow[c] = (zrow[c] - mean) * rstd;          // normalised row
      if (live && p.xh1) {
#pragma unroll
        for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(p.xh1 + (long long)grow * 64 + 8 * q) = pack8(zrow + 8 * q);
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = 8 * q + j;
          f[j] = live ? zrow[c] * sg1[c] + sbe1[c] : 0.f;
        }
        st_sw(sm + OFF_X, r, q, pack8(f));                // h replaces x (each thread only touches its row)
      }
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[3]);
    stamp(0, 8, tl);

    // ---- (5) FFN1 epilogue: + bias, ReLU -> f1 tiles (over q/k/v/P) + global
    tc::mbar_wait(&bar_m[4], 0);
    stamp(0, 11, tl);
    tc::tc_fence_after();
    tmem_walk<256>(ta + C_F1, [&](int c0, const uint32_t (&v)[32]) {
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = live ? fmaxf(__uint_as_float(v[j]) + sb1[c0 + j], 0.f) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + 8 * q;
        st_sw(sm + OFF_F1 + (col >> 6) * TB, r, (col & 63) >> 3, pack8(f + 8 * q));
      }
    });
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[4]);
    stamp(0, 10, tl);

    // ---- (6) FFN2 epilogue: + bias + h -> LayerNorm2 -> y, z2, stats2
    tc::mbar_wait(&bar_m[5], 0);
    stamp(0, 13, tl);
    tc::tc_fence_after();
    {
      float zrow[64];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(ta + C_F2 + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 hr = ld_sw(sm + OFF_X, r, (c0 >> 3) + q);
          const uint32_t hw[4] = {hr.x, hr.y, hr.z, hr.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 hf = up2(hw[j]);
            const int c = c0 + 8 * q + 2 * j;
            zrow[c] = __uint_as_float(v[8 * q + 2 * j]) + sb2[c] + hf.x;
            zrow[c + 1] = __uint_as_float(v[8 * q + 2 * j + 1]) + sb2[c + 1] + hf.y;
          }
        }
      }
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) s += zrow[c];
      const float mean = s * (1.f / 64.f);
      float vs = 0.f;
#pragma unroll
      for (int c = 0; c < 64; ++c) { const float d = zrow[c] - mean; vs = fmaf(d, d, vs); }
      const float rstd = rsqrtf(vs * (1.f / 64.f) + p.eps);
      if (live) {
        if (p.z2) {
          float4* zg = reinterpret_cast<float4*>(p.z2 + (long long)grow * 64);
#pragma unroll
          for (int c = 0; c < 64; c += 4) zg[c >> 2] = make_float4(zrow[c], zrow[c + 1], zrow[c + 2], zrow[c + 3]);
        }
        *reinterpret_cast<float2*>(p.st2 + (long long)grow * 2) = make_float2(mean, rstd);
#pragma unroll
        for (int c = 0; c < 64; ++c) zrow[c] = (zrow[c] - mean) * rstd;
        if (p.xh2) {
#pragma unroll
          for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(p.xh2 + (long long)grow * 64 + 8 * q) = pack8(zrow + 8 * q);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int c = 8 * q + j;
            f[j] = zrow[c] * sg2[c] + sbe2[c];
          }
          *reinterpret_cast<uint4*>(p.y + (long long)grow * 64 + 8 * q) = pack8(f);
        }
      }
    }
    stamp(0, 20, tl);
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused data-gradient pass of the same layer: given dy = dL/dy it chains
//   LN2 bwd -> dz2 --(W2)--> df1 (ReLU mask) --(W1)--> (+dz2) dh -> LN1 bwd -> dz1 --(Wo)--> do
//   -> attention bwd (dP, dS, dQ/dK/dV) -> dqkv --(Win)--> (+dz1) dx
// through shared memory and TMEM.  The row gradients the weight-gradient GEMMs need (dz2, df1, dh,
// dz1, dqkv) are stored to HBM on the way; those GEMMs (and the LayerNorm affine gradients, which
// are the diagonal of xhat^T dy) run as v4l_tc_wgrad launches beside this kernel.
struct BlockBwdParams {
  CUtensorMap tm_qkv;      // [R,192]   box {64, rows_per_tile}
  CUtensorMap tm_w2d;      // [256,64]  box {64,256}   (W2^T: rows = f1 column)
  CUtensorMap tm_w1d;      // [64,256]  box {64,64}    (W1^T: rows = h column), 4 k-chunks
  CUtensorMap tm_wod;      // [64,64]   box {64,64}
  CUtensorMap tm_wind;     // [64,192]  box {64,64}    (Win^T: rows = x column), 3 k-chunks
  CUtensorMap tm_f1;       // [R,256]   box {64, rows_per_tile}: ReLU gate, loaded where df1 is written
  CUtensorMap tm_dz2_o, tm_df1_o, tm_dz1_o, tm_dqkv_o;     // outputs, box {64, rows_per_tile}
  int R, T, rows_per_tile;
  float scale;
  const float *g1, *g2, *st1, *st2, *p;
  const __half *dy, *xh1, *xh2, *f1;
  __half *dz2, *df1, *dh, *dz1, *dqkv, *dx;
};

constexpr int BO_W = 0;                       // 64 KB: W2^T | W1^T, later Wo^T | Win^T
constexpr int BO_R1 = 64 * 1024;              // dz2, later do            16 KB
constexpr int BO_R2 = BO_R1 + TB;             // df1 (4 tiles), later dS (2) | P (2)   64 KB
constexpr int BO_R3 = BO_R2 + 4 * TB;         // dz1                      16 KB
constexpr int BO_R4 = BO_R3 + TB;             // q | k | v, later dqkv    48 KB
constexpr int BWD_SMEM = BO_R4 + 3 * TB;      // 212992

__device__ __forceinline__ void ld_row64(const __half* g, uint4 (&u)[8]) {
  const uint4* s = reinterpret_cast<const uint4*>(g);
#pragma unroll
  for (int q = 0; q < 8; ++q) u[q] = s[q];
}

// LayerNorm backward on one row held in registers: d[] = dL/d(out) in, dL/d(pre-norm sum) out
__device__ __forceinline__ void ln_bwd_row(float (&d)[64], const uint4 (&xh)[8], float rstd, const float* g) {
  float m1 = 0.f, m2 = 0.f;
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t w[4] = {xh[q].x, xh[q].y, xh[q].z, xh[q].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = up2(w[j]);
      const int c = 8 * q + 2 * j;
      d[c] *= g[c]; d[c + 1] *= g[c + 1];
      m1 += d[c] + d[c + 1];
      m2 = fmaf(d[c], x.x, fmaf(d[c + 1], x.y, m2));
    }
  }
  m1 *= (1.f / 64.f); m2 *= (1.f / 64.f);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint32_t w[4] = {xh[q].x, xh[q].y, xh[q].z, xh[q].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x = up2(w[j]);
      const int c = 8 * q + 2 * j;
      d[c] = rstd * (d[c] - m1 - x.x * m2);
      d[c + 1] = rstd * (d[c + 1] - m1 - x.y * m2);
    }
  }
}

__global__ void __launch_bounds__(BK_THREADS, 1) tc_block_bwd_kernel(const __grid_constant__ BlockBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_qkv, bar_f1, bar_wa, bar_wb, bar_m[6], bar_e[6];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_g[128];                  // g2 | g1

  v4l_pdl_trigger();
  const bool tl = blockIdx.x == 0 && threadIdx.x == 32;
  stamp(1, 0, tl);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int row0 = blockIdx.x * p.rows_per_tile;
  {
    uint4* z = reinterpret_cast<uint4*>(sm + BO_R1);
    for (int i = threadIdx.x; i < (BWD_SMEM - BO_R1) / 16; i += BK_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tm_qkv); tc::tma_prefetch_desc(&p.tm_w2d); tc::tma_prefetch_desc(&p.tm_w1d);
    tc::tma_prefetch_desc(&p.tm_wod); tc::tma_prefetch_desc(&p.tm_wind); tc::tma_prefetch_desc(&p.tm_f1);
    tc::tma_prefetch_desc(&p.tm_dz2_o); tc::tma_prefetch_desc(&p.tm_df1_o); tc::tma_prefetch_desc(&p.tm_dz1_o);
    tc::tma_prefetch_desc(&p.tm_dqkv_o);
    tc::mbar_init(&bar_qkv, 1); tc::mbar_init(&bar_f1, 1); tc::mbar_init(&bar_wa, 1); tc::mbar_init(&bar_wb, 1);
    for (int i = 0; i < 6; ++i) { tc::mbar_init(&bar_m[i], 1); tc::mbar_init(&bar_e[i], 128); }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_proxy_async();
  v4l_pdl_wait();
  for (int i = threadIdx.x; i < 64; i += BK_THREADS) { s_g[i] = p.g2[i]; s_g[64 + i] = p.g1[i]; }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  stamp(1, 1, tl);
  // TMEM columns: df1 [0,256)  dh [256,320)  do [320,384)  dx [384,448);  dP [0,128), dQ|dK|dV [128,320) later
  constexpr uint32_t C_DF1 = 0, C_DH = 256, C_DO = 320, C_DX = 384, C_DP = 0, C_DQKV = 128;

  if (warp == 0) {
    // converged warp, one elected lane issues: operands of UTCHMMA / UTMALDG stay in uniform registers
    if (tc::elect_one()) {
      const uint32_t base = tc::smem_u32(sm);
      tc::mbar_expect_tx(&bar_wa, 64u * 1024u);
      tc::tma_load_2d(sm + BO_W, &p.tm_w2d, &bar_wa, 0, 0);
      for (int c = 0; c < 4; ++c) tc::tma_load_2d(sm + BO_W + 32 * 1024 + c * 8192, &p.tm_w1d, &bar_wa, c * 64, 0);
      tc::mbar_expect_tx(&bar_f1, 4u * p.rows_per_tile * 128u);
      for (int c = 0; c < 4; ++c) tc::tma_load_2d(sm + BO_R2 + c * TB, &p.tm_f1, &bar_f1, c * 64, row0);
      tc::mbar_expect_tx(&bar_qkv, 3u * p.rows_per_tile * 128u);
      for (int c = 0; c < 3; ++c) tc::tma_load_2d(sm + BO_R4 + c * TB, &p.tm_qkv, &bar_qkv, c * 64, row0);
      // (1) df1 = dz2 W2
      tc::mbar_wait(&bar_e[0], 0);
      tc::mbar_wait(&bar_wa, 0);
      tc::tc_fence_after();
      tc::tma_store_2d(&p.tm_dz2_o, sm + BO_R1, 0, row0);
      tc::tma_store_commit();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 256, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_f16(tmem + C_DF1, tc::umma_smem_desc(base + BO_R1 + k * 32, 0, 1024),
                       tc::umma_smem_desc(base + BO_W + k * 32, 0, 1024), id, k ? 1u : 0u);
        tc::umma_commit(&bar_m[0]);
      }
      // (2) dh = df1 W1
      tc::mbar_wait(&bar_e[1], 0);
      tc::tc_fence_after();
      for (int c = 0; c < 4; ++c) tc::tma_store_2d(&p.tm_df1_o, sm + BO_R2 + c * TB, c * 64, row0);
      tc::tma_store_commit();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < 16; ++k)
          tc::umma_f16(tmem + C_DH, tc::umma_smem_desc(base + BO_R2 + (k >> 2) * TB + (k & 3) * 32, 0, 1024),
                       tc::umma_smem_desc(base + BO_W + 32 * 1024 + (k >> 2) * 8192 + (k & 3) * 32, 0, 1024), id,
                       k ? 1u : 0u);
        tc::umma_commit(&bar_m[1]);
      }
      // weights of the second half replace the first two once the MMAs above have read them
      tc::mbar_wait(&bar_m[1], 0);
      tc::mbar_expect_tx(&bar_wb, 32u * 1024u);
      tc::tma_load_2d(sm + BO_W, &p.tm_wod, &bar_wb, 0, 0);
      for (int c = 0; c < 3; ++c) tc::tma_load_2d(sm + BO_W + 8192 + c * 8192, &p.tm_wind, &bar_wb, c * 64, 0);
      // (3) do = dz1 Wo
      tc::mbar_wait(&bar_e[2], 0);
      tc::mbar_wait(&bar_wb, 0);
      tc::tc_fence_after();
      tc::tma_store_wait_read();            // dz2 and df1 have left: do / dS / P may overwrite their tiles
      tc::tma_store_2d(&p.tm_dz1_o, sm + BO_R3, 0, row0);
      tc::tma_store_commit();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_f16(tmem + C_DO, tc::umma_smem_desc(base + BO_R3 + k * 32, 0, 1024),
                       tc::umma_smem_desc(base + BO_W + k * 32, 0, 1024), id, k ? 1u : 0u);
        tc::umma_commit(&bar_m[2]);
      }
      // (4) dP = do V^T
      tc::mbar_wait(&bar_e[3], 0);
      tc::mbar_wait(&bar_qkv, 0);
      tc::tc_fence_after();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 128, 0, 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          tc::umma_f16(tmem + C_DP, tc::umma_smem_desc(base + BO_R1 + k * 32, 0, 1024),
                       tc::umma_smem_desc(base + BO_R4 + 2 * TB + k * 32, 0, 1024), id, k ? 1u : 0u);
        tc::umma_commit(&bar_m[3]);
      }
      // (5) dQ = dS K, dK = dS^T Q, dV = P^T do   (dS | P tiles in R2)
      tc::mbar_wait(&bar_e[4], 0);
      tc::tc_fence_after();
      {
        const uint32_t idKm = tc::umma_idesc_f16(128, 64, 0, 1);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          tc::umma_f16(tmem + C_DQKV, tc::umma_smem_desc(base + BO_R2 + (k >> 2) * TB + (k & 3) * 32, 0, 1024),
                       tc::umma_smem_desc(base + BO_R4 + TB + k * 2048, TB, 1024), idKm, k ? 1u : 0u);
        const uint32_t idMM = tc::umma_idesc_f16(128, 64, 1, 1);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          tc::umma_f16(tmem + C_DQKV + 64, tc::umma_smem_desc(base + BO_R2 + k * 2048, TB, 1024),
                       tc::umma_smem_desc(base + BO_R4 + k * 2048, TB, 1024), idMM, k ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          tc::umma_f16(tmem + C_DQKV + 128, tc::umma_smem_desc(base + BO_R2 + 2 * TB + k * 2048, TB, 1024),
                       tc::umma_smem_desc(base + BO_R1 + k * 2048, TB, 1024), idMM, k ? 1u : 0u);
        tc::umma_commit(&bar_m[4]);
      }
      // (6) dx = dqkv Win
      tc::mbar_wait(&bar_e[5], 0);
      tc::tc_fence_after();
      for (int c = 0; c < 3; ++c) tc::tma_store_2d(&p.tm_dqkv_o, sm + BO_R4 + c * TB, c * 64, row0);
      tc::tma_store_commit();
      {
        const uint32_t id = tc::umma_idesc_f16(128, 64, 0, 0);
#pragma unroll
        for (int k = 0; k < 12; ++k)
          tc::umma_f16(tmem + C_DX, tc::umma_smem_desc(base + BO_R4 + (k >> 2) * TB + (k & 3) * 32, 0, 1024),
                       tc::umma_smem_desc(base + BO_W + 8192 + (k >> 2) * 8192 + (k & 3) * 32, 0, 1024), id,
                       k ? 1u : 0u);
        tc::umma_commit(&bar_m[5]);
      }
      tc::tma_store_wait_all();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int s_loc = r / p.T;
    const int lo = s_loc * p.T, hi = lo + p.T;
    const int grow = row0 + r;
    const bool live = (r < p.rows_per_tile) && (grow < p.R);
    const long long gr = live ? grow : 0;             // safe row for loads of non-live threads
    const uint32_t ta = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const float *sg2 = s_g, *sg1 = s_g + 64;

    // ---- (a) LayerNorm2 backward: dy -> dz2 (tile R1 + global)
    {
      uint4 u[8], xh[8];
      ld_row64(p.dy + gr * 64, u);
      ld_row64(p.xh2 + gr * 64, xh);
      const float rstd = p.st2[gr * 2 + 1];
      float d[64];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t w[4] = {u[q].x, u[q].y, u[q].z, u[q].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = up2(w[j]); d[8 * q + 2 * j] = f.x; d[8 * q + 2 * j + 1] = f.y; }
      }
      ln_bwd_row(d, xh, rstd, sg2);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 w = pack8(d + 8 * q);
        if (!live) w = make_uint4(0, 0, 0, 0);
        st_sw(sm + BO_R1, r, q, w);
      }
    }
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[0]);
    stamp(1, 2, tl);

    // ---- (b) df1 = (dz2 W2) * (f1 > 0) -> tiles R2 + global
    tc::mbar_wait(&bar_m[0], 0);
    stamp(1, 3, tl);
    tc::tc_fence_after();
    tc::mbar_wait(&bar_f1, 0);              // the forward activation f1 sits in the tiles df1 replaces
    tmem_walk<256>(ta + C_DF1, [&](int c0, const uint32_t (&v)[32]) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + 8 * q;
        uint8_t* tile = sm + BO_R2 + (col >> 6) * TB;
        const uint4 m = ld_sw(tile, r, (col & 63) >> 3);
        const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 a = up2(mw[j]);
          f[2 * j] = (live && a.x > 0.f) ? __uint_as_float(v[8 * q + 2 * j]) : 0.f;
          f[2 * j + 1] = (live && a.y > 0.f) ? __uint_as_float(v[8 * q + 2 * j + 1]) : 0.f;
        }
        st_sw(tile, r, (col & 63) >> 3, pack8(f));
      }
    });
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[1]);
    stamp(1, 4, tl);

    // ---- (c) dh = df1 W1 + dz2 -> global (LN1 affine gradients); LayerNorm1 backward -> dz1 (R3 + global)
    uint4 xh1r[8];
    ld_row64(p.xh1 + gr * 64, xh1r);        // fetched ahead of the MMA wait
    const float rstd1 = p.st1[gr * 2 + 1];
    tc::mbar_wait(&bar_m[1], 0);
    stamp(1, 5, tl);
    tc::tc_fence_after();
    {
      float d[64];
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(ta + C_DH + c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 rr = ld_sw(sm + BO_R1, r, (c0 >> 3) + q);
          const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = up2(w[j]);
            d[c0 + 8 * q + 2 * j] = __uint_as_float(v[8 * q + 2 * j]) + f.x;
            d[c0 + 8 * q + 2 * j + 1] = __uint_as_float(v[8 * q + 2 * j + 1]) + f.y;
          }
        }
      }
      if (live) {
#pragma unroll
        for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(p.dh + (long long)grow * 64 + 8 * q) = pack8(d + 8 * q);
      }
      ln_bwd_row(d, xh1r, rstd1, sg1);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 w = pack8(d + 8 * q);
        if (!live) w = make_uint4(0, 0, 0, 0);
        st_sw(sm + BO_R3, r, q, w);
      }
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[2]);
    stamp(1, 6, tl);

    // ---- (d) do = dz1 Wo -> tile R1 (never leaves the SM)
    tc::mbar_wait(&bar_m[2], 0);
    stamp(1, 7, tl);
    tc::tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(ta + C_DO + c0, v);
      tc::tmem_ld_wait();
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = live ? __uint_as_float(v[j]) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) st_sw(sm + BO_R1, r, (c0 >> 3) + q, pack8(f + 8 * q));
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[3]);
    stamp(1, 8, tl);

    // ---- (e) dS = P (dP - rowsum(dP P)) scale and P -> tiles R2 (dS: 0,1; P: 2,3)
    // the saved probabilities of this query row, aligned to its own 32-column chunk(s), are fetched
    // before waiting for dP so that their latency hides behind the MMA
    const int ownA = min(lo, 127) & ~31, ownB = min(hi - 1, 127) & ~31;
    const bool two = ownB != ownA;
    float pa[32], pb[32];
    {
      const float* prow = p.p + gr * p.T;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int ia = ownA + j - lo, ib = ownB + j - lo;
        pa[j] = (live && ia >= 0 && ia < p.T) ? prow[ia] : 0.f;
        pb[j] = (live && two && ib < p.T) ? prow[ib] : 0.f;
      }
    }
    tc::mbar_wait(&bar_m[3], 0);
    stamp(1, 9, tl);
    tc::tc_fence_after();
    {
      const int wlo = ((quad * 32) / p.T) * p.T;
      const int whi = ((quad * 32 + 31) / p.T + 1) * p.T;
      const int cb = wlo & ~31, ce = min(128, (whi + 31) & ~31);     // chunks holding this warp's samples
      uint32_t ka[32], kb[32];
      tmem_own_chunks(ta + C_DP, cb, ce, ownA, ownB, ka, kb);
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        dot = fmaf(__uint_as_float(ka[j]), pa[j], dot);
        if (two) dot = fmaf(__uint_as_float(kb[j]), pb[j], dot);
      }
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {   // every chunk is written: the tiles held df1 before
        float ds[32], pp[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float pv = (c0 == ownA) ? pa[j] : ((c0 == ownB) ? pb[j] : 0.f);
          const float dp = (c0 == ownA) ? __uint_as_float(ka[j]) : __uint_as_float(kb[j]);
          pp[j] = pv;
          ds[j] = pv * (dp - dot) * p.scale;     // pv == 0 off the sample's block
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = c0 + 8 * q;
          st_sw(sm + BO_R2 + (col >> 6) * TB, r, (col & 63) >> 3, pack8(ds + 8 * q));
          st_sw(sm + BO_R2 + (2 + (col >> 6)) * TB, r, (col & 63) >> 3, pack8(pp + 8 * q));
        }
      }
    }
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[4]);
    stamp(1, 10, tl);

    // ---- (f) dQ | dK | dV -> tiles R4 + global dqkv
    tc::mbar_wait(&bar_m[4], 0);
    stamp(1, 11, tl);
    tc::tc_fence_after();
    tmem_walk<192>(ta + C_DQKV, [&](int c0, const uint32_t (&v)[32]) {
      float f[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = live ? __uint_as_float(v[j]) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + 8 * q;
        st_sw(sm + BO_R4 + (col >> 6) * TB, r, (col & 63) >> 3, pack8(f + 8 * q));
      }
    });
    tc::tc_fence_before();
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_e[5]);
    stamp(1, 12, tl);

    // ---- (g) dx = dqkv Win + dz1 -> global
    tc::mbar_wait(&bar_m[5], 0);
    stamp(1, 13, tl);
    tc::tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(ta + C_DX + c0, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 rr = ld_sw(sm + BO_R3, r, (c0 >> 3) + q);
        const uint32_t w[4] = {rr.x, rr.y, rr.z, rr.w};
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 z = up2(w[j]);
          f[2 * j] = __uint_as_float(v[8 * q + 2 * j]) + z.x;
          f[2 * j + 1] = __uint_as_float(v[8 * q + 2 * j + 1]) + z.y;
        }
        if (live) *reinterpret_cast<uint4*>(p.dx + (long long)grow * 64 + c0 + 8 * q) = pack8(f);
      }
    }
    stamp(1, 20, tl);
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

int enc2(CUtensorMap* m, const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows, const char* who) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {cols * 2};
  uint32_t box[2] = {64, box_rows};
  return v4l_encode_tmap(m, ptr, 2, dims, str, box, who, nullptr);
}

}  // namespace

extern "C" int v4l_tc_block_fwd(v4l_ctx* ctx, void* stream, const v4l_tc_block_args* a) {
  V4L_REQUIRE(ctx && a && a->x && a->w_in && a->w_o && a->w_1 && a->w_2 && a->y && a->qkv && a->o && a->h && a->f1 &&
              a->p && a->st1 && a->st2, "v4l_tc_block_fwd: NULL argument");
  V4L_REQUIRE(a->T >= 2 && a->T <= 64 && a->B >= 0, "v4l_tc_block_fwd: bad shape");
  if (a->B == 0) return 0;
  BlockParams p;
  memset(&p, 0, sizeof(p));
  p.T = a->T; p.R = a->B * a->T;
  const int spt = 128 / a->T;
  p.rows_per_tile = spt * a->T;
  p.scale = 0.125f; p.eps = a->eps;
  const char* who = "v4l_tc_block_fwd";
  if (int r = enc2(&p.tm_x, a->x, 64, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_win, a->w_in, 64, 192, 192, who)) return r;
  if (int r = enc2(&p.tm_wo, a->w_o, 64, 64, 64, who)) return r;
  if (int r = enc2(&p.tm_w1, a->w_1, 64, 256, 256, who)) return r;
  if (int r = enc2(&p.tm_w2, a->w_2, 256, 64, 64, who)) return r;
  if (int r = enc2(&p.tm_qkv_o, a->qkv, 192, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_o_o, a->o, 64, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_h_o, a->h, 64, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_f1_o, a->f1, 256, p.R, p.rows_per_tile, who)) return r;
  p.b_in = a->b_in; p.b_o = a->b_o; p.g1 = a->g1; p.be1 = a->be1; p.b1 = a->b1; p.b2 = a->b2; p.g2 = a->g2; p.be2 = a->be2;
  p.qkv = (__half*)a->qkv; p.o = (__half*)a->o; p.h = (__half*)a->h; p.f1 = (__half*)a->f1; p.y = (__half*)a->y;
  p.p = a->p; p.z1 = a->z1; p.st1 = a->st1; p.z2 = a->z2; p.st2 = a->st2;
  p.xh1 = (__half*)a->xh1; p.xh2 = (__half*)a->xh2;
  static bool attr = false;
  if (!attr) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_block_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES + 1024));
    attr = true;
  }
  V4L_LAUNCH(tc_block_fwd_kernel, v4l_cdiv(a->B, spt), BK_THREADS, SMEM_BYTES + 1024, (cudaStream_t)stream, p);
  return 0;
}

extern "C" int v4l_tc_block_bwd(v4l_ctx* ctx, void* stream, const v4l_tc_block_bwd_args* a) {
  V4L_REQUIRE(ctx && a && a->dy && a->qkv && a->p && a->xh1 && a->xh2 && a->st1 && a->st2 && a->f1 && a->w2d && a->w1d &&
              a->wod && a->wind && a->g1 && a->g2 && a->dz2 && a->df1 && a->dh && a->dz1 && a->dqkv && a->dx,
              "v4l_tc_block_bwd: NULL argument");
  V4L_REQUIRE(a->T >= 2 && a->T <= 64 && a->B >= 0, "v4l_tc_block_bwd: bad shape");
  if (a->B == 0) return 0;
  BlockBwdParams p;
  memset(&p, 0, sizeof(p));
  p.T = a->T; p.R = a->B * a->T;
  const int spt = 128 / a->T;
  p.rows_per_tile = spt * a->T;
  p.scale = 0.125f;
  const char* who = "v4l_tc_block_bwd";
  if (int r = enc2(&p.tm_qkv, a->qkv, 192, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_w2d, a->w2d, 64, 256, 256, who)) return r;
  if (int r = enc2(&p.tm_w1d, a->w1d, 256, 64, 64, who)) return r;
  if (int r = enc2(&p.tm_wod, a->wod, 64, 64, 64, who)) return r;
  if (int r = enc2(&p.tm_wind, a->wind, 192, 64, 64, who)) return r;
  if (int r = enc2(&p.tm_f1, a->f1, 256, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_dz2_o, a->dz2, 64, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_df1_o, a->df1, 256, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_dz1_o, a->dz1, 64, p.R, p.rows_per_tile, who)) return r;
  if (int r = enc2(&p.tm_dqkv_o, a->dqkv, 192, p.R, p.rows_per_tile, who)) return r;
  p.g1 = a->g1; p.g2 = a->g2; p.st1 = a->st1; p.st2 = a->st2; p.p = a->p;
  p.dy = (const __half*)a->dy; p.xh1 = (const __half*)a->xh1; p.xh2 = (const __half*)a->xh2; p.f1 = (const __half*)a->f1;
  p.dz2 = (__half*)a->dz2; p.df1 = (__half*)a->df1; p.dh = (__half*)a->dh; p.dz1 = (__half*)a->dz1;
  p.dqkv = (__half*)a->dqkv; p.dx = (__half*)a->dx;
  static bool attr = false;
  if (!attr) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_block_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM + 1024));
    attr = true;
  }
  V4L_LAUNCH(tc_block_bwd_kernel, v4l_cdiv(a->B, spt), BK_THREADS, BWD_SMEM + 1024, (cudaStream_t)stream, p);
  return 0;
}

// diagnostics: globaltimer stamps (ns) of CTA 0's epilogue thread in the most recent forward ([0..31])
// and data-gradient ([32..63]) launch: 0 start, 1 set-up done, 2+2i epilogue i finished,
// 3+2i MMA i result observed, 20 end.  Synchronises the device.
extern "C" int v4l_tc_block_timeline(unsigned long long* host_out) {
  V4L_REQUIRE(host_out, "v4l_tc_block_timeline: NULL argument");
  V4L_CHECK_CUDA(cudaDeviceSynchronize());
  V4L_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, g_timeline, sizeof(unsigned long long) * 64));
  return 0;
}
