// A chain of up to three Linear layers in ONE kernel on tcgen05 (the actor / critic MLP heads
// `append_fcs`, reference torchrl/networks/nets.py:973-992,1036, the proprio MLP + projector,
// base.py:8-44,209-230, and their data-gradient chains):
//
//   h1 = act(x  W1^T + b1) [* (m1 > 0)],  h2 = act(h1 W2^T + b2) [* (m2 > 0)],  y = act(h2 W3^T + b3) [* (m3 > 0)]
//
// A CTA owns a 128-row tile.  Layer 1 streams its A operand (x, TMA) and its weights through a 3-stage
// shared-memory ring, 64 K-columns at a time; the following layers read A from the hidden tile the previous
// epilogue left in shared memory (fp16, 128-byte-swizzled K-major, 4 x [128 x 64]) and stream only weights.
// Accumulators ping-pong between two 256-column TMEM buffers.  Every layer's result can also be stored to
// global memory (the saved activations / row gradients the weight-gradient GEMMs read, or the fp32 output
// through a row map).  With the 3 launches of a head merged, the 128 rows never leave the SM between layers
// and two launch + pipeline-fill latencies disappear from the minibatch's serial chain.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer (both converged, one elected
// lane issues: operands stay in uniform registers), warps 2-9 epilogue: two warps per TMEM lane quadrant,
// each taking every other 32-column chunk.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int ML_THREADS = 64 + 8 * 32;
constexpr int ML_STAGES = 3;
constexpr int ML_A_BYTES = 128 * 128;                // [128 rows][64 f16]
constexpr int ML_W_BYTES = 256 * 128;                // [<=256 rows][64 f16]
constexpr int ML_STAGE = ML_A_BYTES + ML_W_BYTES;
constexpr int ML_H_BYTES = 4 * ML_A_BYTES;           // hidden tile: up to 256 columns
constexpr int ML_SMEM = ML_STAGES * ML_STAGE + ML_H_BYTES;

struct MlpLayerP {
  CUtensorMap tm_w;
  int K, N, N_valid, relu;
  const float* bias;
  const __half* mask; long long mask_ld;
  void* out; int out_f32;
  v4l_rowmap out_map;
};
struct MlpParams {
  CUtensorMap tm_x;
  int M, n_layers;
  MlpLayerP L[3];
};

__device__ __forceinline__ uint32_t ml_pk2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ml_lo(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u & 0xffffu))); }
__device__ __forceinline__ float ml_hi(uint32_t u) { return __half2float(__ushort_as_half(static_cast<unsigned short>(u >> 16))); }

__global__ void __launch_bounds__(ML_THREADS, 1) tc_mlp_chain_kernel(const __grid_constant__ MlpParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[ML_STAGES], empty_bar[ML_STAGES], acc_bar, h_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_bias[3][256];

  v4l_pdl_trigger();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int row0 = blockIdx.x * 128;
  uint8_t* hbuf = smem + ML_STAGES * ML_STAGE;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tm_x);
    for (int l = 0; l < p.n_layers; ++l) tc::tma_prefetch_desc(&p.L[l].tm_w);
    for (int s = 0; s < ML_STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(&acc_bar, 1);
    tc::mbar_init(&h_bar, 256);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_slot, 512);
  v4l_pdl_wait();
  for (int l = 0; l < p.n_layers; ++l)
    for (int i = threadIdx.x; i < 256; i += ML_THREADS)
      s_bias[l][i] = (p.L[l].bias && i < p.L[l].N_valid) ? p.L[l].bias[i] : 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    int stage = 0; uint32_t phase = 0;
    for (int l = 0; l < p.n_layers; ++l) {
      const int kcn = p.L[l].K >> 6;
      const uint32_t bytes = static_cast<uint32_t>(p.L[l].N) * 128u + (l == 0 ? ML_A_BYTES : 0);
      for (int kc = 0; kc < kcn; ++kc) {
        tc::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* s = smem + stage * ML_STAGE;
        if (tc::elect_one()) {
          tc::mbar_expect_tx(&full_bar[stage], bytes);
          if (l == 0) tc::tma_load_2d(s, &p.tm_x, &full_bar[stage], kc * 64, row0);
          tc::tma_load_2d(s + ML_A_BYTES, &p.L[l].tm_w, &full_bar[stage], kc * 64, 0);
        }
        __syncwarp();
        if (++stage == ML_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    int stage = 0; uint32_t phase = 0;
    const uint32_t hb = tc::smem_u32(hbuf);
    for (int l = 0; l < p.n_layers; ++l) {
      if (l > 0) {                                   // the previous epilogue has written the hidden tile
        tc::mbar_wait(&h_bar, (l - 1) & 1);
        tc::tc_fence_after();
      }
      const uint32_t idesc = tc::umma_idesc_f16(128, p.L[l].N, 0, 0);
      const uint32_t d = tmem + (l & 1) * 256;
      const int kcn = p.L[l].K >> 6;
      for (int kc = 0; kc < kcn; ++kc) {
        tc::mbar_wait(&full_bar[stage], phase);
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + stage * ML_STAGE);
        const uint64_t adesc0 = tc::umma_smem_desc(l == 0 ? sa : hb + kc * ML_A_BYTES, 0, 1024);
        const uint64_t bdesc0 = tc::umma_smem_desc(sa + ML_A_BYTES, 0, 1024);
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_f16(d, adesc0 + 2 * k, bdesc0 + 2 * k, idesc, (kc | k) ? 1u : 0u);
          tc::umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == ML_STAGES) { stage = 0; phase ^= 1; }
      }
      if (tc::elect_one()) tc::umma_commit(&acc_bar);
      __syncwarp();
    }
  } else {
    // ============================== epilogue ==================================
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int r = quad * 32 + lane;
    const int grow = row0 + r;
    const bool live = grow < p.M;
    for (int l = 0; l < p.n_layers; ++l) {
      const MlpLayerP& L = p.L[l];
      const bool last = (l + 1 == p.n_layers);
      tc::mbar_wait(&acc_bar, l & 1);
      tc::tc_fence_after();
      const uint32_t ta = tmem + (static_cast<uint32_t>(quad * 32) << 16) + (l & 1) * 256;
      const long long oaddr = (L.out && live) ? v4l_row_addr(L.out_map, grow) : 0;
      const int nchunks = (L.N + 31) >> 5;
      for (int c = half; c < nchunks; c += 2) {
        const int c0 = c * 32;
        uint32_t v[32];
        if (L.N - c0 >= 32) {
          tc::tmem_ld_32x32(ta + c0, v);
        } else {
          uint32_t v16[16];
          tc::tmem_ld_32x16(ta + c0, v16);
#pragma unroll
          for (int j = 0; j < 16; ++j) { v[j] = v16[j]; v[16 + j] = 0; }
        }
        tc::tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __uint_as_float(v[j]) + s_bias[l][c0 + j];
          if (L.relu) f[j] = fmaxf(f[j], 0.f);
          if (!live) f[j] = 0.f;
        }
        if (L.mask && live) {                                    // ReLU gate of the layer this gradient enters
          const __half* mp = L.mask + (long long)grow * L.mask_ld + c0;
          if (c0 + 32 <= L.N_valid && ((reinterpret_cast<uintptr_t>(mp) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 m = __ldg(reinterpret_cast<const uint4*>(mp) + q);
              const uint32_t mw[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (!(ml_lo(mw[j]) > 0.f)) f[8 * q + 2 * j] = 0.f;
                if (!(ml_hi(mw[j]) > 0.f)) f[8 * q + 2 * j + 1] = 0.f;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < L.N_valid && !(__half2float(mp[j]) > 0.f)) f[j] = 0.f;
          }
        }
        if (!last) {
          // next layer's A operand: 128-byte-swizzled K-major tile (col >> 6), 16-byte chunk ((col & 63) >> 3) ^ (row & 7)
          uint8_t* tile = hbuf + (c0 >> 6) * ML_A_BYTES + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 w;
            w.x = ml_pk2(f[8 * q + 0], f[8 * q + 1]); w.y = ml_pk2(f[8 * q + 2], f[8 * q + 3]);
            w.z = ml_pk2(f[8 * q + 4], f[8 * q + 5]); w.w = ml_pk2(f[8 * q + 6], f[8 * q + 7]);
            const int chunk = ((c0 & 63) >> 3) + q;
            *reinterpret_cast<uint4*>(tile + ((chunk ^ (r & 7)) << 4)) = w;
          }
        }
        if (L.out && live) {
          if (!L.out_f32 && c0 + 32 <= L.N_valid && (((oaddr + c0) & 7) == 0)) {
            __half* cp = reinterpret_cast<__half*>(L.out) + oaddr + c0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 w;
              w.x = ml_pk2(f[8 * q + 0], f[8 * q + 1]); w.y = ml_pk2(f[8 * q + 2], f[8 * q + 3]);
              w.z = ml_pk2(f[8 * q + 4], f[8 * q + 5]); w.w = ml_pk2(f[8 * q + 6], f[8 * q + 7]);
              *(reinterpret_cast<uint4*>(cp) + q) = w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (c0 + j >= L.N_valid) continue;
              if (L.out_f32) reinterpret_cast<float*>(L.out)[oaddr + c0 + j] = f[j];
              else reinterpret_cast<__half*>(L.out)[oaddr + c0 + j] = __float2half(f[j]);
            }
          }
        }
      }
      if (!last) {
        // the hidden tile's columns beyond this layer's N (K padding of the next layer) must read as zero
        const int n_next = p.L[l + 1].K;
        for (int c0 = L.N + half * 32; c0 < n_next; c0 += 64) {
          uint8_t* tile = hbuf + (c0 >> 6) * ML_A_BYTES + r * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int chunk = ((c0 & 63) >> 3) + q;
            *reinterpret_cast<uint4*>(tile + ((chunk ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
          }
        }
        tc::fence_proxy_async();             // generic-proxy writes of the hidden tile -> visible to the UMMA proxy
        tc::tc_fence_before();
        tc::mbar_arrive(&h_bar);
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

extern "C" int v4l_tc_mlp_chain(v4l_ctx* ctx, void* stream, const v4l_tc_mlp_chain_args* a) {
  V4L_REQUIRE(ctx && a && a->x, "v4l_tc_mlp_chain: NULL argument");
  V4L_REQUIRE(a->n_layers >= 1 && a->n_layers <= 3, "v4l_tc_mlp_chain: 1..3 layers");
  V4L_REQUIRE(a->M >= 0 && a->x_cols >= 8 && a->x_cols % 8 == 0 && a->x_ld >= a->x_cols && a->x_ld % 8 == 0,
              "v4l_tc_mlp_chain: bad input shape (cols %d, ld %lld)", a->x_cols, (long long)a->x_ld);
  if (a->M == 0) return 0;
  MlpParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[2] = {(uint64_t)a->x_cols, (uint64_t)a->M};
    uint64_t str[1] = {(uint64_t)a->x_ld * 2};
    uint32_t box[2] = {64, 128};
    if (int r = v4l_encode_tmap(&p.tm_x, a->x, 2, dims, str, box, "v4l_tc_mlp_chain(x)", nullptr)) return r;
  }
  p.M = a->M; p.n_layers = a->n_layers;
  for (int l = 0; l < a->n_layers; ++l) {
    const v4l_tc_mlp_layer& s = a->layer[l];
    V4L_REQUIRE(s.w && s.K >= 64 && s.K % 64 == 0 && s.K <= 1024, "v4l_tc_mlp_chain: layer %d K = %d", l, s.K);
    V4L_REQUIRE(s.N_pad >= 16 && s.N_pad % 16 == 0 && s.N_pad <= 256 && s.N_valid >= 1 && s.N_valid <= s.N_pad,
                "v4l_tc_mlp_chain: layer %d N = %d (%d valid)", l, s.N_pad, s.N_valid);
    V4L_REQUIRE(l == 0 || (s.K <= 256 && s.K >= a->layer[l - 1].N_valid),
                "v4l_tc_mlp_chain: layer %d reads %d columns of a %d-column hidden tile", l, s.K, a->layer[l - 1].N_pad);
    V4L_REQUIRE(!s.out || s.out_map.P > 0, "v4l_tc_mlp_chain: layer %d output row map", l);
    uint64_t dims[2] = {(uint64_t)s.K, (uint64_t)s.N_pad};
    uint64_t str[1] = {(uint64_t)s.K * 2};
    uint32_t box[2] = {64, (uint32_t)s.N_pad};
    if (int r = v4l_encode_tmap(&p.L[l].tm_w, s.w, 2, dims, str, box, "v4l_tc_mlp_chain(W)", nullptr)) return r;
    p.L[l].K = s.K; p.L[l].N = s.N_pad; p.L[l].N_valid = s.N_valid; p.L[l].relu = s.relu;
    p.L[l].bias = s.bias;
    p.L[l].mask = reinterpret_cast<const __half*>(s.mask); p.L[l].mask_ld = s.mask_ld;
    p.L[l].out = s.out; p.L[l].out_f32 = s.out_f32; p.L[l].out_map = s.out_map;
  }
  static bool attr_set = false;
  const size_t smem = (size_t)ML_SMEM + 1024;
  if (!attr_set) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  V4L_LAUNCH(tc_mlp_chain_kernel, v4l_cdiv(a->M, 128), ML_THREADS, smem, (cudaStream_t)stream, p);
  V4L_CHECK_LAUNCH();
  return 0;
}
