// "Flat" valid convolution on tcgen05 (NatureCNN trunk forward, reference torchrl/networks/base.py:304-342).
//
// tc_gemm.cu expresses a convolution as a sum over taps of tap-shifted TMA boxes: every tap pulls
// its own copy of the activation tile through L2 (4x for the 2x2 space-to-depth forms, 9x for conv3),
// and that L2->SM traffic, not HBM or the tensor pipe, bounds the narrow-N trunk layers.  Here the
// (image, position) grid is flattened to rows R = img * P + h * Wg + w of a [rows, C] fp16 matrix and
// a 128-row tile [R0, R0 + 128) plus a few slack rows is loaded ONCE; tap (dh, dw) is the same
// shared-memory tile read through a UMMA descriptor whose start is moved by dh * Wg + dw rows.  Rows
// whose window would leave the image (h >= Hout or w >= Wout) compute garbage and are not stored.
//
// The swizzle is a function of the absolute shared-memory address bits, so a descriptor whose start is
// NOT a multiple of the 8-row / 1024-byte atom reads the TMA-written tile correctly with the
// base-offset field left 0 (measured: bit-exact against the tap-box path; setting base offset =
// shift mod 8 is wrong).  `mode` 0 keeps the conservative alternative: one pre-shifted copy of the
// tile per distinct (shift mod 8), every start atom aligned; mode 1 = single copy.
// Warp roles as tc_gemm.cu: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM owner), then up to four
// epilogue warpgroups, one per TMEM accumulator stage, on round-robin tiles.  The packed weights (all taps) stay resident in
// shared memory for the whole kernel.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int CV_THREADS = 64 + 4 * 128;      // TMA warp, MMA warp, up to 4 epilogue warpgroups
constexpr int CV_MAX_ACC = 4;
constexpr int CV_MAX_STAGES = 8;
constexpr int CV_MAX_TAPS = 16;
constexpr int CV_MAX_MMA = CV_MAX_TAPS * 4 * 4;

struct ConvFlatParams {
  CUtensorMap tm_a;            // [rows_total, C] box {64, load_rows}
  CUtensorMap tm_w;            // [N_pad, n_taps * kc * 64] box {64, N}
  int P, Wg, Hout, Wout;
  int kc, n_taps;
  int shift[CV_MAX_TAPS];      // dh * Wg + dw >= 0
  int load_rows;               // 128 + slack, multiple of 8
  int n_copies, copy_res[8];
  int mode;
  int num_tiles, tiles_per_img;
  int N, N_valid;
  long long n_img;
  const int32_t* a_idx;
  const float* bias;
  int flags;
  __half* c;
  v4l_rowmap c_map;
  int stages;
  int group;                   // tiles whose MMAs are issued interleaved (independent accumulators)
};

// MMA issue loop, executed by the WHOLE issuer warp (converged; the warp index is made provably uniform
// with a shuffle) with one elected lane issuing: every descriptor is then computed in uniform registers and
// a tcgen05.mma costs its hardware floor (~48 cycles for N <= 64; tools/ubench/mma_rate.cu).  Round 1 ran
// this loop under `if (lane == 0)` — tabulated descriptors included — where the compiler must move each
// operand of each UTCHMMA into uniform registers through an elect + broadcast loop: ~150 cycles per MMA
// whatever N is, which had been misread as a property of the tensor pipe.  The MMAs of G tiles
// (independent accumulators) are still issued interleaved.
struct IssueCtx {
  uint64_t *full_bar, *empty_bar, *tmem_full, *tmem_empty;
  uint32_t ring_addr, stage_bytes, tmem_base, w_addr, copy_bytes;
  int acc_cols, n_acc;
  uint32_t idesc;
  int stages, num_tiles;
};

// byte offset of tap t's A start inside a stage (mode 0: pre-shifted copy whose residue matches)
__device__ __forceinline__ uint32_t tap_offset(const struct ConvFlatParams& p, int t, uint32_t copy_bytes);

template <int G>
__device__ __forceinline__ void issue_mmas(const struct ConvFlatParams& p, const IssueCtx& c);

__device__ __forceinline__ uint32_t tap_offset(const ConvFlatParams& p, int t, uint32_t copy_bytes) {
  const int sh = p.shift[t];
  if (p.mode != 0) return static_cast<uint32_t>(sh) * 128u;
  int cpy = 0;
  for (int q = 0; q < p.n_copies; ++q) if (p.copy_res[q] == (sh & 7)) cpy = q;
  return cpy * copy_bytes + static_cast<uint32_t>(sh & ~7) * 128u;
}

template <int G>
__device__ __forceinline__ void issue_mmas(const ConvFlatParams& p, const IssueCtx& c) {
  int stage = 0; uint32_t phase = 0;
  int acc = 0; uint32_t acc_phase = 0;
  const int step = static_cast<int>(gridDim.x);
  const uint64_t bdesc0 = tc::umma_smem_desc(c.w_addr, 0, 1024);
  for (int tile = blockIdx.x; tile < c.num_tiles; tile += G * step) {
    bool v[G];
    uint32_t d[G];
    uint64_t ab[G];
    uint64_t *eb[G], *tf[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      v[g] = tile + g * step < c.num_tiles;
      d[g] = 0; ab[g] = 0; eb[g] = nullptr; tf[g] = nullptr;
      if (v[g]) {
        tc::mbar_wait(&c.tmem_empty[acc], acc_phase ^ 1);
        tc::mbar_wait(&c.full_bar[stage], phase);
        d[g] = c.tmem_base + acc * c.acc_cols;
        ab[g] = tc::umma_smem_desc(c.ring_addr + stage * c.stage_bytes, 0, 1024);
        eb[g] = &c.empty_bar[stage]; tf[g] = &c.tmem_full[acc];
        if (++stage == c.stages) { stage = 0; phase ^= 1; }
        if (++acc == c.n_acc) { acc = 0; acc_phase ^= 1; }
      }
    }
    tc::tc_fence_after();
    if (tc::elect_one()) {
      for (int t = 0; t < p.n_taps; ++t) {
        const uint32_t toff = tap_offset(p, t, c.copy_bytes);
        for (int j = 0; j < p.kc; ++j) {
          const uint32_t ao = (toff + j * p.load_rows * 128) >> 4;
          const uint64_t bd = bdesc0 + (((t * p.kc + j) * p.N * 128) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int g = 0; g < G; ++g)
              if (v[g]) tc::umma_f16(d[g], ab[g] + ao + 2 * k, bd + 2 * k, c.idesc, (t | j | k) ? 1u : 0u);
          }
        }
      }
#pragma unroll
      for (int g = 0; g < G; ++g)
        if (v[g]) { tc::umma_commit(eb[g]); tc::umma_commit(tf[g]); }
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(CV_THREADS, 1) tc_conv_flat_kernel(const __grid_constant__ ConvFlatParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[CV_MAX_STAGES], empty_bar[CV_MAX_STAGES], tmem_full[CV_MAX_ACC], tmem_empty[CV_MAX_ACC], w_bar;
  __shared__ uint32_t tmem_base_slot;
  __shared__ float s_bias[256];

  v4l_pdl_trigger();
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;   // provably warp-uniform
  const int N = p.N;
  // accumulator stages in TMEM = epilogue warpgroups: a tile's MMAs are short next to the
  // MMA -> epilogue -> MMA hand-over, so narrow layers need several tiles in flight
  const int n_acc = N <= 128 ? 4 : 2;
  const int acc_cols = 512 / n_acc;
  const int n_chunks = p.n_taps * p.kc;
  const uint32_t w_bytes = static_cast<uint32_t>(n_chunks) * N * 128u;
  const uint32_t copy_bytes = static_cast<uint32_t>(p.kc) * p.load_rows * 128u;     // one copy: kc chunks of [load_rows][64]
  const uint32_t stage_bytes = copy_bytes * p.n_copies;
  uint8_t* w_smem = smem;
  uint8_t* ring = smem + ((w_bytes + 1023u) & ~1023u);

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&p.tm_a);
    tc::tma_prefetch_desc(&p.tm_w);
    for (int s = 0; s < p.stages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < CV_MAX_ACC; ++s) { tc::mbar_init(&tmem_full[s], 1); tc::mbar_init(&tmem_empty[s], 128); }
    tc::mbar_init(&w_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(&tmem_base_slot, 512);
  v4l_pdl_wait();
  for (int i = threadIdx.x; i < 256; i += CV_THREADS) s_bias[i] = (p.bias && i < p.N_valid) ? p.bias[i] : 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    // whole warp converged, one elected lane issues
    if (tc::elect_one()) {
      tc::mbar_expect_tx(&w_bar, w_bytes);
      for (int q = 0; q < n_chunks; ++q) tc::tma_load_2d(w_smem + q * N * 128, &p.tm_w, &w_bar, q * 64, 0);
    }
    __syncwarp();
    int stage = 0; uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      int row0 = tile * 128;                          // first tensor row of the tile
      if (p.a_idx) {
        const int img = tile / p.tiles_per_img;
        row0 = __shfl_sync(0xffffffffu, p.a_idx[img], 0) * p.P + (tile - img * p.tiles_per_img) * 128;
      }
      tc::mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* s = ring + stage * stage_bytes;
      if (tc::elect_one()) {
        tc::mbar_expect_tx(&full_bar[stage], stage_bytes);
        for (int c = 0; c < p.n_copies; ++c)
          for (int j = 0; j < p.kc; ++j)
            tc::tma_load_2d(s + c * copy_bytes + j * p.load_rows * 128, &p.tm_a, &full_bar[stage], j * 64,
                            row0 + p.copy_res[c]);
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    const uint32_t idesc = tc::umma_idesc_f16(128, N, 0, 0);
    tc::mbar_wait(&w_bar, 0);
    IssueCtx c{full_bar, empty_bar, tmem_full, tmem_empty, tc::smem_u32(ring), stage_bytes, tmem_base,
               tc::smem_u32(w_smem), copy_bytes, acc_cols, n_acc, idesc, p.stages, p.num_tiles};
    if (p.group >= 4) issue_mmas<4>(p, c);
    else if (p.group == 2) issue_mmas<2>(p, c);
    else issue_mmas<1>(p, c);
  } else if (((warp - 2) >> 2) < n_acc) {
    // ============================== epilogue ==================================
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const bool relu = p.flags & V4L_RELU;
    const int acc = wg;
    uint32_t acc_phase = 0;
    auto row_of = [&](int tile, bool& ok) -> long long {
      const long long R = (long long)tile * 128 + r;
      const long long img = R / p.P;
      const int pos = static_cast<int>(R - img * p.P);
      const int h = pos / p.Wg, w = pos - h * p.Wg;
      ok = (img < p.n_img) && (h < p.Hout) && (w < p.Wout);
      return ok ? v4l_row_addr(p.c_map, static_cast<int>((img * p.Hout + h) * p.Wout + w)) : 0;
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
        tc::tmem_ld_32x32(taddr + c0, v);
        tc::tmem_ld_wait();
        if (!row_ok) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          f[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
          if (relu) f[j] = fmaxf(f[j], 0.f);
        }
        __half* cp = p.c + row_addr + c0;
        if (c0 + 32 <= p.N_valid) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            __half2 h0 = __floats2half2_rn(f[8 * q + 0], f[8 * q + 1]), h1 = __floats2half2_rn(f[8 * q + 2], f[8 * q + 3]);
            __half2 h2 = __floats2half2_rn(f[8 * q + 4], f[8 * q + 5]), h3 = __floats2half2_rn(f[8 * q + 6], f[8 * q + 7]);
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
            *(reinterpret_cast<uint4*>(cp) + q) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < p.N_valid) cp[j] = __float2half(f[j]);
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


}  // namespace

extern "C" int v4l_tc_conv_flat(v4l_ctx* ctx, void* stream, const v4l_tc_conv_flat_args* a) {
  V4L_REQUIRE(ctx && a && a->x && a->w && a->c, "v4l_tc_conv_flat: NULL argument");
  V4L_REQUIRE(a->C >= 64 && a->C % 64 == 0 && a->C <= 256, "v4l_tc_conv_flat: C=%d must be a multiple of 64", a->C);
  V4L_REQUIRE(a->n_taps >= 1 && a->n_taps <= CV_MAX_TAPS, "v4l_tc_conv_flat: bad taps");
  V4L_REQUIRE(a->N_pad % 32 == 0 && a->N_pad >= 32 && a->N_pad <= 256 && a->N_valid >= 1 && a->N_valid <= a->N_pad &&
              a->N_valid % 8 == 0, "v4l_tc_conv_flat: N_pad=%d N_valid=%d", a->N_pad, a->N_valid);
  V4L_REQUIRE((a->mode & 15) <= 1 && ((a->mode >> 4) & 15) <= 4, "v4l_tc_conv_flat: bad mode");
  V4L_REQUIRE(a->P >= 1 && a->Wg >= 1 && a->Hout >= 1 && a->Wout <= a->Wg && a->Hout * a->Wg <= a->P + a->Wg,
              "v4l_tc_conv_flat: bad grid");
  V4L_REQUIRE(!a->x_idx || a->P % 128 == 0, "v4l_tc_conv_flat: gathered images need P %% 128 == 0");
  if (a->n_img == 0) return 0;
  ConvFlatParams p;
  memset(&p, 0, sizeof(p));
  p.P = a->P; p.Wg = a->Wg; p.Hout = a->Hout; p.Wout = a->Wout;
  p.kc = a->C / 64; p.n_taps = a->n_taps;
  int max_shift = 0;
  bool seen[8] = {false, false, false, false, false, false, false, false};
  for (int t = 0; t < a->n_taps; ++t) {
    V4L_REQUIRE(a->tap_dh[t] >= 0 && a->tap_dw[t] >= 0, "v4l_tc_conv_flat: taps must be non-negative (valid convolution)");
    p.shift[t] = a->tap_dh[t] * a->Wg + a->tap_dw[t];
    max_shift = max(max_shift, p.shift[t]);
    seen[p.shift[t] & 7] = true;
  }
  p.mode = a->mode & 15;
  p.group = ((a->mode >> 4) & 15) ? ((a->mode >> 4) & 15) : 2;      // bits 4-7: interleave group (default 2)
  if (p.mode == 0) {
    for (int q = 0; q < 8; ++q) if (seen[q]) p.copy_res[p.n_copies++] = q;
  } else {
    p.n_copies = 1; p.copy_res[0] = 0;
  }
  p.load_rows = 128 + ((max_shift + 7) / 8) * 8;
  V4L_REQUIRE(p.load_rows <= 256, "v4l_tc_conv_flat: tap reach %d too large", max_shift);
  p.tiles_per_img = a->x_idx ? a->P / 128 : 1;
  p.num_tiles = v4l_cdiv((long long)a->n_img * a->P, 128);
  p.N = a->N_pad; p.N_valid = a->N_valid; p.n_img = a->n_img;
  p.a_idx = a->x_idx; p.bias = a->bias; p.flags = a->flags;
  p.c = reinterpret_cast<__half*>(a->c); p.c_map = a->c_map;
  const char* who = "v4l_tc_conv_flat";
  {
    uint64_t dims[2] = {(uint64_t)a->C, (uint64_t)a->x_rows};
    uint64_t str[1] = {(uint64_t)a->C * 2};
    uint32_t box[2] = {64, (uint32_t)p.load_rows};
    if (int r = v4l_encode_tmap(&p.tm_a, a->x, 2, dims, str, box, who, nullptr)) return r;
  }
  const int Ktot = a->n_taps * a->C;
  {
    uint64_t dims[2] = {(uint64_t)Ktot, (uint64_t)a->N_pad};
    uint64_t str[1] = {(uint64_t)Ktot * 2};
    uint32_t box[2] = {64, (uint32_t)a->N_pad};
    if (int r = v4l_encode_tmap(&p.tm_w, a->w, 2, dims, str, box, who, nullptr)) return r;
  }
  const size_t w_bytes = (((size_t)a->n_taps * p.kc * a->N_pad * 128) + 1023) & ~(size_t)1023;
  const size_t stage_bytes = (size_t)p.n_copies * p.kc * p.load_rows * 128;
  V4L_REQUIRE(w_bytes + 2 * stage_bytes <= 200 * 1024, "v4l_tc_conv_flat: weights + 2 stages exceed shared memory");
  p.stages = (int)min((size_t)CV_MAX_STAGES, (200 * 1024 - w_bytes) / stage_bytes);
  p.group = min(p.group, min(p.stages, a->N_pad <= 128 ? 4 : 2));
  const size_t smem = w_bytes + (size_t)p.stages * stage_bytes + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    V4L_CHECK_CUDA(cudaFuncSetAttribute(tc_conv_flat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 202 * 1024));
    attr_set = true;
  }
  V4L_LAUNCH(tc_conv_flat_kernel, min(p.num_tiles, ctx->sm_count), CV_THREADS, smem, (cudaStream_t)stream, p);
  V4L_CHECK_LAUNCH();
  return 0;
}
