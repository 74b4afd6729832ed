// Tensor-core attention core of the LocoTransformer block (single head, d = 64, T <= 32 tokens):
// several samples are packed into one 128-row tile (7 x 17 tokens, or 8 x 16) and the per-sample
// 17x17 attention becomes block-diagonal 128x128 tcgen05 MMAs:
//
//   forward   S = Q K^T (M128 N128 K64)  -> masked softmax in registers (thread = row) ->
//             P (fp16, unnormalised) to swizzled smem -> O = P V (M128 N64 K128, V as MN-major B)
//   backward  dP = dO V^T -> dS = P (dP - rowsum(dP P)) scale in registers -> dS, P to smem ->
//             dQ = dS K (K as MN-major B), dK = dS^T Q and dV = P^T dO (the SAME smem tiles read as
//             MN-major A operands: no transposes)
//
// Off-block entries of P / dS are written as exact zeros, so the cross-sample terms vanish.
// Reference math: nn.MultiheadAttention inside nn.TransformerEncoderLayer
// (torchrl/networks/nets.py:949-955).  One CTA per tile, 160 threads:
// warp 0 = TMEM allocation + TMA + MMA issue, warps 1-4 = softmax / epilogue (one TMEM lane
// quadrant each).
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int AT_THREADS = 160;
constexpr int TILE_BYTES = 128 * 128;        // [128 rows][64 fp16]

struct AttnParams {
  CUtensorMap tmap_qkv;     // 2D [R, 192] fp16, box {64, rows_per_tile}
  CUtensorMap tmap_do;      // 2D [R, 64] fp16 (backward only)
  int R, T, spt, rows_per_tile, B;
  float scale;
  __half* o;                // fwd out [R, 64]
  float* p;                 // [B, T, T] fp32 (fwd: out, bwd: in)
  __half* dqkv;             // bwd out [R, 192]
};

__device__ __forceinline__ void sw128_store16(uint8_t* tile, int row, int chunk, uint4 v) {
  // K-major / MN-major SWIZZLE_128B tile: row pitch 128 B, 16-byte chunk index XOR (row mod 8)
  *reinterpret_cast<uint4*>(tile + row * 128 + ((chunk ^ (row & 7)) << 4)) = v;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS, 1) tc_attn_fwd_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                       // [rows, 64]   A of S, K-major
  uint8_t* sK = smem + TILE_BYTES;          // [rows, 64]   B of S, K-major (N = key row)
  uint8_t* sV = smem + 2 * TILE_BYTES;      // [rows, 64]   B of O, MN-major (K = key row)
  uint8_t* sP = smem + 3 * TILE_BYTES;      // 2 atoms [128 rows i][64 cols j]   A of O, K-major
  __shared__ uint64_t bar_load, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_slot;

  v4l_pdl_trigger();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int row0 = blockIdx.x * p.rows_per_tile;
  {  // zero all tiles: rows beyond the TMA box (tile padding) must contribute exact zeros
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < 5 * TILE_BYTES / 16; i += AT_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tmap_qkv);
    tc::mbar_init(&bar_load, 1); tc::mbar_init(&bar_s, 1); tc::mbar_init(&bar_p, 128); tc::mbar_init(&bar_o, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  tc::fence_proxy_async();
  v4l_pdl_wait();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // converged warp, one elected lane issues: operands of UTCHMMA / UTMALDG stay in uniform registers
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&bar_load, 3u * p.rows_per_tile * 128u);
      tc::tma_load_2d(sQ, &p.tmap_qkv, &bar_load, 0, row0);
      tc::tma_load_2d(sK, &p.tmap_qkv, &bar_load, 64, row0);
      tc::tma_load_2d(sV, &p.tmap_qkv, &bar_load, 128, row0);
      tc::mbar_wait(&bar_load, 0);
      tc::tc_fence_after();
      // S = Q K^T
      const uint32_t idS = tc::umma_idesc_f16(128, 128, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc::umma_f16(tmem, tc::umma_smem_desc(tc::smem_u32(sQ) + k * 32, 0, 1024),
                     tc::umma_smem_desc(tc::smem_u32(sK) + k * 32, 0, 1024), idS, k ? 1u : 0u);
      tc::umma_commit(&bar_s);
      // O = P V  (after the softmax warps have written P)
      tc::mbar_wait(&bar_p, 0);
      tc::tc_fence_after();
      const uint32_t idO = tc::umma_idesc_f16(128, 64, 0, 1);
#pragma unroll
      for (int k = 0; k < 8; ++k) {           // K = 128 key rows, 16 per MMA
        const uint32_t a = tc::smem_u32(sP) + (k >> 2) * TILE_BYTES + (k & 3) * 32;
        const uint32_t b = tc::smem_u32(sV) + k * 2048;      // MN-major: 16 rows x 128 B per K step
        tc::umma_f16(tmem + 128, tc::umma_smem_desc(a, 0, 1024), tc::umma_smem_desc(b, TILE_BYTES, 1024), idO,
                     k ? 1u : 0u);
      }
      tc::umma_commit(&bar_o);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                      // tile row = TMEM lane
    const int s_loc = r / p.T;                           // sample inside the tile
    const int ti = r - s_loc * p.T;
    const int lo = s_loc * p.T, hi = lo + p.T;           // this sample's key columns
    const int grow = row0 + r;
    const bool live = (r < p.rows_per_tile) && (grow < p.R);
    const uint32_t taddr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    tc::mbar_wait(&bar_s, 0);
    tc::tc_fence_after();
    // pass 1: row maximum over the sample's own columns
    float mx = -3.0e38f;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(taddr + c0, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j >= lo && c0 + j < hi) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    // pass 2: e = exp(scale (s - max)) inside the block, exact 0 outside; unnormalised P -> smem
    float sum = 0.f;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(taddr + c0, v);
      tc::tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const bool in = live && (c0 + j >= lo) && (c0 + j < hi);
        e[j] = in ? __expf((__uint_as_float(v[j]) - mx) * p.scale) : 0.f;
        sum += e[j];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack2(e[8 * q + 0], e[8 * q + 1]); w.y = pack2(e[8 * q + 2], e[8 * q + 3]);
        w.z = pack2(e[8 * q + 4], e[8 * q + 5]); w.w = pack2(e[8 * q + 6], e[8 * q + 7]);
        const int col = c0 + 8 * q;
        sw128_store16(sP + (col >> 6) * TILE_BYTES, r, (col & 63) >> 3, w);
      }
    }
    const float inv = live ? 1.f / sum : 0.f;
    tc::fence_proxy_async();                // generic-proxy smem writes -> visible to tcgen05.mma
    tc::mbar_arrive(&bar_p);
    // pass 3: normalised probabilities of the sample's block -> global (saved for the backward)
    if (live) {
      float* prow = p.p + ((long long)(grow / p.T) * p.T + ti) * p.T;
      for (int c0 = 0; c0 < 128; c0 += 32) {
        if (c0 + 32 <= lo || c0 >= hi) continue;          // warp-divergent skip is fine: no collectives inside
        for (int j = 0; j < 32; ++j) {
          const int col = c0 + j;
          if (col >= lo && col < hi) {
            // recompute from the unnormalised fp16 value just written (what the MMA consumes)
            const __half hv = *reinterpret_cast<const __half*>(sP + (col >> 6) * TILE_BYTES + r * 128 +
                                                               ((((col & 63) >> 3) ^ (r & 7)) << 4) + (col & 7) * 2);
            prow[col - lo] = __half2float(hv) * inv;
          }
        }
      }
    }
    // epilogue: O / sum -> fp16 -> global
    tc::mbar_wait(&bar_o, 0);
    tc::tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(taddr + 128 + c0, v);
      tc::tmem_ld_wait();
      if (live) {
        __half* dst = p.o + (long long)grow * 64 + c0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack2(__uint_as_float(v[8 * q + 0]) * inv, __uint_as_float(v[8 * q + 1]) * inv);
          w.y = pack2(__uint_as_float(v[8 * q + 2]) * inv, __uint_as_float(v[8 * q + 3]) * inv);
          w.z = pack2(__uint_as_float(v[8 * q + 4]) * inv, __uint_as_float(v[8 * q + 5]) * inv);
          w.w = pack2(__uint_as_float(v[8 * q + 6]) * inv, __uint_as_float(v[8 * q + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 8 * q) = w;
        }
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(AT_THREADS, 1) tc_attn_bwd_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;
  uint8_t* sV = smem + 2 * TILE_BYTES;
  uint8_t* sG = smem + 3 * TILE_BYTES;      // dO
  uint8_t* sS = smem + 4 * TILE_BYTES;      // dS: 2 atoms [128 i][64 j]
  uint8_t* sP = smem + 6 * TILE_BYTES;      // P : 2 atoms
  __shared__ uint64_t bar_load, bar_dp, bar_w, bar_out;
  __shared__ uint32_t tmem_slot;

  v4l_pdl_trigger();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int row0 = blockIdx.x * p.rows_per_tile;
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (int i = threadIdx.x; i < 8 * TILE_BYTES / 16; i += AT_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tmap_qkv);
    tc::tma_prefetch_desc(&p.tmap_do);
    tc::mbar_init(&bar_load, 1); tc::mbar_init(&bar_dp, 1); tc::mbar_init(&bar_w, 128); tc::mbar_init(&bar_out, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_proxy_async();
  v4l_pdl_wait();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  // TMEM columns: dP [0,128)  dQ [128,192)  dK [192,256)  dV [256,320)

  if (warp == 0) {
    // converged warp, one elected lane issues: operands of UTCHMMA / UTMALDG stay in uniform registers
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&bar_load, 4u * p.rows_per_tile * 128u);
      tc::tma_load_2d(sQ, &p.tmap_qkv, &bar_load, 0, row0);
      tc::tma_load_2d(sK, &p.tmap_qkv, &bar_load, 64, row0);
      tc::tma_load_2d(sV, &p.tmap_qkv, &bar_load, 128, row0);
      tc::tma_load_2d(sG, &p.tmap_do, &bar_load, 0, row0);
      tc::mbar_wait(&bar_load, 0);
      tc::tc_fence_after();
      // dP = dO V^T
      const uint32_t idP = tc::umma_idesc_f16(128, 128, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        tc::umma_f16(tmem, tc::umma_smem_desc(tc::smem_u32(sG) + k * 32, 0, 1024),
                     tc::umma_smem_desc(tc::smem_u32(sV) + k * 32, 0, 1024), idP, k ? 1u : 0u);
      tc::umma_commit(&bar_dp);
      tc::mbar_wait(&bar_w, 0);             // dS and P are in smem
      tc::tc_fence_after();
      // dQ = dS K : A = dS K-major (k = key j), B = K tile MN-major (rows = j)
      const uint32_t idKm = tc::umma_idesc_f16(128, 64, 0, 1);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        tc::umma_f16(tmem + 128,
                     tc::umma_smem_desc(tc::smem_u32(sS) + (k >> 2) * TILE_BYTES + (k & 3) * 32, 0, 1024),
                     tc::umma_smem_desc(tc::smem_u32(sK) + k * 2048, TILE_BYTES, 1024), idKm, k ? 1u : 0u);
      // dK = dS^T Q and dV = P^T dO : A = the same tiles read MN-major (M = key j, K = query i)
      const uint32_t idMM = tc::umma_idesc_f16(128, 64, 1, 1);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        tc::umma_f16(tmem + 192, tc::umma_smem_desc(tc::smem_u32(sS) + k * 2048, TILE_BYTES, 1024),
                     tc::umma_smem_desc(tc::smem_u32(sQ) + k * 2048, TILE_BYTES, 1024), idMM, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        tc::umma_f16(tmem + 256, tc::umma_smem_desc(tc::smem_u32(sP) + k * 2048, TILE_BYTES, 1024),
                     tc::umma_smem_desc(tc::smem_u32(sG) + k * 2048, TILE_BYTES, 1024), idMM, k ? 1u : 0u);
      tc::umma_commit(&bar_out);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int s_loc = r / p.T;
    const int ti = r - s_loc * p.T;
    const int lo = s_loc * p.T, hi = lo + p.T;
    const int grow = row0 + r;
    const bool live = (r < p.rows_per_tile) && (grow < p.R);
    const uint32_t taddr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
    const float* prow = p.p + ((long long)(grow / p.T) * p.T + ti) * p.T;
    tc::mbar_wait(&bar_dp, 0);
    tc::tc_fence_after();
    // pass 1: rowdot = sum_j dP_ij P_ij over the sample's block
    float dot = 0.f;
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(taddr + c0, v);
      tc::tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j >= lo && c0 + j < hi) dot = fmaf(__uint_as_float(v[j]), prow[c0 + j - lo], dot);
      }
    }
    // pass 2: dS = P (dP - rowdot) scale and P, fp16, exact zeros off-block -> smem
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(taddr + c0, v);
      tc::tmem_ld_wait();
      float ds[32], pp[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const bool in = live && (c0 + j >= lo) && (c0 + j < hi);
        const float pv = in ? prow[c0 + j - lo] : 0.f;
        pp[j] = pv;
        ds[j] = in ? pv * (__uint_as_float(v[j]) - dot) * p.scale : 0.f;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int col = c0 + 8 * q;
        uint4 w;
        w.x = pack2(ds[8 * q + 0], ds[8 * q + 1]); w.y = pack2(ds[8 * q + 2], ds[8 * q + 3]);
        w.z = pack2(ds[8 * q + 4], ds[8 * q + 5]); w.w = pack2(ds[8 * q + 6], ds[8 * q + 7]);
        sw128_store16(sS + (col >> 6) * TILE_BYTES, r, (col & 63) >> 3, w);
        w.x = pack2(pp[8 * q + 0], pp[8 * q + 1]); w.y = pack2(pp[8 * q + 2], pp[8 * q + 3]);
        w.z = pack2(pp[8 * q + 4], pp[8 * q + 5]); w.w = pack2(pp[8 * q + 6], pp[8 * q + 7]);
        sw128_store16(sP + (col >> 6) * TILE_BYTES, r, (col & 63) >> 3, w);
      }
    }
    tc::fence_proxy_async();
    tc::mbar_arrive(&bar_w);
    tc::mbar_wait(&bar_out, 0);
    tc::tc_fence_after();
    // dQ | dK | dV -> dqkv[row, 0:64 | 64:128 | 128:192]
#pragma unroll
    for (int part = 0; part < 3; ++part) {
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(taddr + 128 + part * 64 + c0, v);
        tc::tmem_ld_wait();
        if (live) {
          __half* dst = p.dqkv + (long long)grow * 192 + part * 64 + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w;
            w.x = pack2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
            w.y = pack2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
            w.z = pack2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
            w.w = pack2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
            *reinterpret_cast<uint4*>(dst + 8 * q) = w;
          }
        }
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

int setup(AttnParams& p, const void* qkv, const void* d_o, int B, int T, const char* who) {
  memset(&p, 0, sizeof(p));
  p.B = B; p.T = T; p.R = B * T;
  p.spt = 128 / T;
  p.rows_per_tile = p.spt * T;
  p.scale = 0.125f;                                    // 1 / sqrt(64)
  {
    uint64_t dims[2] = {192, (uint64_t)p.R};
    uint64_t str[1] = {192 * 2};
    uint32_t box[2] = {64, (uint32_t)p.rows_per_tile};
    if (int r = v4l_encode_tmap(&p.tmap_qkv, qkv, 2, dims, str, box, who, nullptr)) return r;
  }
  if (d_o) {
    uint64_t dims[2] = {64, (uint64_t)p.R};
    uint64_t str[1] = {64 * 2};
    uint32_t box[2] = {64, (uint32_t)p.rows_per_tile};
    if (int r = v4l_encode_tmap(&p.tmap_do, d_o, 2, dims, str, box, who, nullptr)) return r;
  }
  return 0;
}

}  // namespace

extern "C" int v4l_tc_attn_fwd(v4l_ctx* ctx, void* stream, const void* qkv, void* o, float* p, int B, int T) {
  V4L_REQUIRE(ctx && qkv && o && p, "v4l_tc_attn_fwd: NULL argument");
  V4L_REQUIRE(T >= 2 && T <= 64, "v4l_tc_attn_fwd: T=%d unsupported", T);
  if (B == 0) return 0;
  AttnParams a;
  if (int r = setup(a, qkv, nullptr, B, T, "v4l_tc_attn_fwd")) return r;
  a.o = reinterpret_cast<__half*>(o); a.p = p;
  static bool attr = false;
  if (!attr) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * TILE_BYTES + 1024));
    attr = true;
  }
  V4L_LAUNCH(tc_attn_fwd_kernel, v4l_cdiv(B, a.spt), AT_THREADS, 5 * TILE_BYTES + 1024, (cudaStream_t)stream, a);
  return 0;
}

extern "C" int v4l_tc_attn_bwd(v4l_ctx* ctx, void* stream, const void* qkv, const float* p, const void* d_o,
                               void* d_qkv, int B, int T) {
  V4L_REQUIRE(ctx && qkv && p && d_o && d_qkv, "v4l_tc_attn_bwd: NULL argument");
  V4L_REQUIRE(T >= 2 && T <= 64, "v4l_tc_attn_bwd: T=%d unsupported", T);
  if (B == 0) return 0;
  AttnParams a;
  if (int r = setup(a, qkv, d_o, B, T, "v4l_tc_attn_bwd")) return r;
  a.p = const_cast<float*>(p); a.dqkv = reinterpret_cast<__half*>(d_qkv);
  static bool attr = false;
  if (!attr) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * TILE_BYTES + 1024));
    attr = true;
  }
  V4L_LAUNCH(tc_attn_bwd_kernel, v4l_cdiv(B, a.spt), AT_THREADS, 8 * TILE_BYTES + 1024, (cudaStream_t)stream, a);
  return 0;
}
