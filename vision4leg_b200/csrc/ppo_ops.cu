// GAE scan, PPO loss epilogues and the fused clip-grad-norm + Adam step.
//
// These are HBM/latency-bound kernels: GAE moves 18 B per transition, the
// loss epilogue ~(3A+6)*4 B per sample, clip+Adam 20 B per parameter.  They read coalesced,
// keep all reductions on-chip (warp shuffles) and never sync with the host: the 18 logged
// statistics of a minibatch (reference ppo.py:77-92,122-123,142-145) land in a device-side
// info row that the host reads once per epoch.
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace {

// =============================================================================================
// GAE: A_t = a_t + b_t * A_{t+1}; (a1,b1) o (a2,b2) = (a1 + b1*a2, b1*b2)   [SURVEY Appendix A4]
// =============================================================================================
struct AB { double a, b; };

__device__ __forceinline__ AB ab_compose(AB x, AB later) {   // x is EARLIER in time
  return AB{fma(x.b, later.a, x.a), x.b * later.b};
}

__device__ __forceinline__ AB ab_shfl_down(AB x, int off) {
  AB r;
  r.a = __shfl_down_sync(0xffffffffu, x.a, off);
  r.b = __shfl_down_sync(0xffffffffu, x.b, off);
  return r;
}

constexpr int GAE_THREADS = 256;

// Inclusive SUFFIX scan over the block (thread i = time offset i inside the tile):
// returns S_i = x_i o x_{i+1} o ... o x_{n-1}.  wsum: shared AB[GAE_THREADS/32 + 1].
__device__ __forceinline__ AB block_suffix_scan(AB x, AB* wsum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NW = GAE_THREADS / 32;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    AB y = ab_shfl_down(x, off);
    if (lane + off < 32) x = ab_compose(x, y);
  }
  if (lane == 0) wsum[warp] = x;          // composite of the whole warp
  __syncthreads();
  if (warp == 0) {
    AB w = (lane < NW) ? wsum[lane] : AB{0.0, 1.0};
#pragma unroll
    for (int off = 1; off < NW; off <<= 1) {
      AB y = ab_shfl_down(w, off);
      if (lane + off < NW) w = ab_compose(w, y);
    }
    if (lane < NW) wsum[lane] = w;        // inclusive suffix over warps
  }
  __syncthreads();
  if (warp + 1 < NW) x = ab_compose(x, wsum[warp + 1]);
  __syncthreads();                        // wsum reusable after return
  return x;
}

struct GaeArgs {
  const float *r, *v, *d, *tl, *last_value;
  long long tl_st, tl_se;
  float *advs, *rets;
  int T, E, chunk_len, n_chunks;
  double gamma, tau;
  int use_tl, mode;
};

__device__ __forceinline__ AB gae_element(const GaeArgs& g, int t, int e) {
  const long long i = (long long)t * g.E + e;
  const double r = g.r[i], v = g.v[i];
  const double nt = 1.0 - (double)g.d[i];
  const double tl = g.use_tl ? (double)g.tl[t * g.tl_st + e * g.tl_se] : 0.0;
  const double ntl = 1.0 - tl;
  AB x;
  if (g.mode == 0) {
    const double vn = (t + 1 < g.T) ? (double)g.v[i + g.E] : (double)g.last_value[e];
    const double delta = r + nt * g.gamma * vn - v;
    x.a = ntl * delta;
    x.b = ntl * nt * g.gamma * g.tau;
  } else {
    x.a = r + tl * v;
    x.b = nt * g.gamma * ntl;
  }
  return x;
}

// phase 1: per (column e, chunk c) composite of the chunk -> agg[e * n_chunks + c]
__global__ void __launch_bounds__(GAE_THREADS) gae_aggregate_kernel(const GaeArgs g, AB* __restrict__ agg) {
  v4l_pdl_enter();
  __shared__ AB wsum[GAE_THREADS / 32 + 1];
  const int e = blockIdx.y, c = blockIdx.x;
  const int t_lo = c * g.chunk_len, t_hi = min(g.T, t_lo + g.chunk_len);
  AB total{0.0, 1.0};                                    // identity
  for (int tile_hi = t_hi; tile_hi > t_lo; tile_hi -= GAE_THREADS) {
    const int tile_lo = max(t_lo, tile_hi - GAE_THREADS);
    const int t = tile_lo + threadIdx.x;
    AB x = (t < tile_hi) ? gae_element(g, t, e) : AB{0.0, 1.0};
    AB s = block_suffix_scan(x, wsum);
    // thread 0 holds the tile composite; this tile is EARLIER than what `total` covers
    if (threadIdx.x == 0) total = ab_compose(s, total);
  }
  if (threadIdx.x == 0) agg[(long long)e * g.n_chunks + c] = total;
}

// phase 2: carry[e][c] = A at the first step AFTER chunk c
__global__ void gae_carry_kernel(const GaeArgs g, const AB* __restrict__ agg, double* __restrict__ carry) {
  v4l_pdl_enter();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= g.E) return;
  double A = (g.mode == 0) ? 0.0 : (double)g.last_value[e];
  for (int c = g.n_chunks - 1; c >= 0; --c) {
    carry[(long long)e * g.n_chunks + c] = A;
    const AB x = agg[(long long)e * g.n_chunks + c];
    A = fma(x.b, A, x.a);
  }
}

// phase 3: scan each chunk with its carry-in and write advs / rets
__global__ void __launch_bounds__(GAE_THREADS) gae_scan_kernel(const GaeArgs g, const double* __restrict__ carry) {
  v4l_pdl_enter();
  __shared__ AB wsum[GAE_THREADS / 32 + 1];
  __shared__ double s_carry;
  const int e = blockIdx.y, c = blockIdx.x;
  const int t_lo = c * g.chunk_len, t_hi = min(g.T, t_lo + g.chunk_len);
  double A_next = carry ? carry[(long long)e * g.n_chunks + c]
                        : ((g.mode == 0) ? 0.0 : (double)g.last_value[e]);
  for (int tile_hi = t_hi; tile_hi > t_lo; tile_hi -= GAE_THREADS) {
    const int tile_lo = max(t_lo, tile_hi - GAE_THREADS);
    const int t = tile_lo + threadIdx.x;
    const bool live = t < tile_hi;
    AB x = live ? gae_element(g, t, e) : AB{0.0, 1.0};
    AB s = block_suffix_scan(x, wsum);
    const double A = fma(s.b, A_next, s.a);
    if (live) {
      const long long i = (long long)t * g.E + e;
      const double v = g.v[i];
      if (g.mode == 0) { g.advs[i] = (float)A; g.rets[i] = (float)(A + v); }
      else             { g.advs[i] = (float)(A - v); g.rets[i] = (float)A; }
    }
    if (threadIdx.x == 0) s_carry = A;     // A at tile_lo = carry for the next (earlier) tile
    __syncthreads();
    A_next = s_carry;
    __syncthreads();
  }
}

// =============================================================================================
// minibatch bookkeeping
// =============================================================================================
__global__ void select_rows_kernel(const int32_t* __restrict__ flat_idx, const int32_t* __restrict__ slot,
                                   int32_t* __restrict__ cur, int n) {
  v4l_pdl_enter();
  const long long base = (long long)(*slot) * n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    cur[i] = flat_idx[base + i];
}

__global__ void slot_advance_kernel(int32_t* slot, int32_t wrap) {
  v4l_pdl_enter();
  int32_t s = *slot + 1;
  if (wrap > 0 && s >= wrap) s = 0;
  *slot = s;
}

// =============================================================================================
// block reductions
// =============================================================================================
template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T ident, T* sh /* [32] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  v = (threadIdx.x < nw) ? sh[threadIdx.x] : ident;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) sh[0] = v;
  }
  __syncthreads();
  v = sh[0];
  return v;
}

struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpAddF { __device__ float operator()(float a, float b) const { return a + b; } };
struct OpMaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct OpMinF { __device__ float operator()(float a, float b) const { return fminf(a, b); } };

// stats = {sum, sumsq, n, max, min} (doubles) — one CTA, deterministic
__global__ void __launch_bounds__(1024) adv_stats_kernel(const float* __restrict__ adv,
                                                         const int32_t* __restrict__ idx, int n,
                                                         double* __restrict__ stats) {
  v4l_pdl_enter();
  __shared__ double shd[32];
  __shared__ float shf[32];
  double s = 0.0, s2 = 0.0;
  float mx = -FLT_MAX, mn = FLT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = adv[idx ? idx[i] : i];
    s += a; s2 += (double)a * a;
    mx = fmaxf(mx, a); mn = fminf(mn, a);
  }
  s = block_reduce(s, OpAddD(), 0.0, shd);
  s2 = block_reduce(s2, OpAddD(), 0.0, shd);
  mx = block_reduce(mx, OpMaxF(), -FLT_MAX, shf);
  mn = block_reduce(mn, OpMinF(), FLT_MAX, shf);
  if (threadIdx.x == 0) {
    stats[0] = s; stats[1] = s2; stats[2] = (double)n; stats[3] = mx; stats[4] = mn;
  }
}

// =============================================================================================
// critic loss (reference ppo.py:94-114)
// =============================================================================================
constexpr int LOSS_THREADS = 256;

// The CTA that finishes last (device-wide arrival counter, reset by it for the next launch) sums the
// per-CTA partials in CTA order: one launch, same fixed summation order as a separate finalize pass.
__device__ __forceinline__ bool last_block_arrives(unsigned int* counter) {
  __shared__ bool s_last;
  __threadfence();                                   // this CTA's partials are visible device-wide
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *counter = 0u;                       // every CTA has arrived: safe to re-arm
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}

__global__ void __launch_bounds__(LOSS_THREADS)
vf_loss_kernel(const float* __restrict__ values, const float* __restrict__ returns,
               const float* __restrict__ old_values, const int32_t* __restrict__ idx,
               float* __restrict__ d_values, __half* __restrict__ d_f16, float scale_f16, int n,
               float inv_global, float inv_local, int clipped, float clip,
               double* part, unsigned int* counter, float* __restrict__ info,
               const int32_t* __restrict__ slot) {
  v4l_pdl_enter();
  __shared__ double shd[32];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = idx ? idx[i] : i;
    const float v = values[i], R = returns[r];
    const float e1 = v - R;
    float l, g;
    if (!clipped) {
      l = e1 * e1;
      g = 2.f * e1 * inv_global;
    } else {
      const float vo = old_values[r];
      const float dv = v - vo;
      const float vc = vo + fminf(fmaxf(dv, -clip), clip);
      const float e2 = vc - R;
      const float l1 = e1 * e1, l2 = e2 * e2;
      const float inside = (dv >= -clip && dv <= clip) ? 1.f : 0.f;
      // 0.5 * max(l1, l2); torch.max splits the gradient evenly on ties
      if (l1 > l2)      { l = 0.5f * l1; g = e1 * inv_global; }
      else if (l2 > l1) { l = 0.5f * l2; g = e2 * inside * inv_global; }
      else              { l = 0.5f * l1; g = 0.5f * (e1 + e2 * inside) * inv_global; }
    }
    d_values[i] = g;
    if (d_f16) {
      // tensor-core tier: the backward pass starts from the loss-scaled fp16 gradient, padded to 16 columns
      uint4 z = make_uint4(0, 0, 0, 0);
      uint4 f = z;
      f.x = (uint32_t)__half_as_ushort(__float2half(g * scale_f16));
      reinterpret_cast<uint4*>(d_f16 + (long long)i * 16)[0] = f;
      reinterpret_cast<uint4*>(d_f16 + (long long)i * 16)[1] = z;
    }
    acc += l;
  }
  acc = block_reduce(acc, OpAddD(), 0.0, shd);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
  if (!last_block_arrives(counter)) return;
  if (threadIdx.x < 32) {
    // lane-strided partial sums + fixed shuffle tree: the loads pipeline instead of forming a serial chain
    double s = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) s += __ldcg(part + i);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) {
      float* row = info + (long long)(slot ? *slot : 0) * V4L_INFO_STRIDE;
      row[V4L_INFO_VF_LOSS] = (float)(s * inv_local);
    }
  }
}

// =============================================================================================
// actor loss (reference ppo.py:42-92, continuous_policy.py:127-146, :486-492)
// part layout per CTA (doubles): [0] sum loss_term, [1] sum lp, [2] sum lp^2, [3] max lp,
// [4] min lp, [5] max ratio, [6] min ratio, [8 .. 8+A) sum d logstd
// =============================================================================================
constexpr int PF_PART = 8 + 64;
constexpr int MAX_A = 64;

__global__ void __launch_bounds__(LOSS_THREADS)
pf_loss_kernel(const float* __restrict__ mean, const float* __restrict__ logstd,
               const float* __restrict__ tmean, const float* __restrict__ tlogstd,
               const float* __restrict__ acts, const float* __restrict__ adv,
               const int32_t* __restrict__ idx, const double* __restrict__ adv_stats,
               float* __restrict__ d_mean, __half* __restrict__ d_f16, float scale_f16, int n, int A,
               float inv_global, float inv_local, float clip, float entropy_coeff, double* part, int t_indexed,
               float* __restrict__ d_logstd, unsigned int* counter, float* __restrict__ info,
               const int32_t* __restrict__ slot, int stats_per_slot) {
  v4l_pdl_enter();
  if (stats_per_slot && slot) adv_stats += (long long)(*slot) * 8;      // per-minibatch table (v4l_adv_stats_epoch)
  __shared__ float s_ls[MAX_A], s_tls[MAX_A], s_ivar[MAX_A], s_tivar[MAX_A];
  __shared__ float s_dls[LOSS_THREADS / 32][MAX_A];
  __shared__ double shd[32];
  __shared__ float shf[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < A) {
    const float ls = fminf(fmaxf(logstd[threadIdx.x], -5.f), 2.f);      // LOG_SIG_MIN/MAX
    const float tls = fminf(fmaxf(tlogstd[threadIdx.x], -5.f), 2.f);
    s_ls[threadIdx.x] = ls; s_tls[threadIdx.x] = tls;
    s_ivar[threadIdx.x] = expf(-2.f * ls); s_tivar[threadIdx.x] = expf(-2.f * tls);
  }
  for (int e = threadIdx.x; e < (LOSS_THREADS / 32) * MAX_A; e += LOSS_THREADS) (&s_dls[0][0])[e] = 0.f;
  __syncthreads();
  const double an = adv_stats[2];
  const double amean = adv_stats[0] / an;
  const double avar = fmax((adv_stats[1] - adv_stats[0] * adv_stats[0] / an) / (an - 1.0), 0.0);
  const float adv_mean = (float)amean;
  const float adv_inv = 1.f / ((float)sqrt(avar) + 1e-5f);
  const float HALF_LOG_2PI = 0.91893853320467274178f;

  double a_loss = 0.0, a_lp = 0.0, a_lp2 = 0.0;
  float lp_max = -FLT_MAX, lp_min = FLT_MAX, r_max = -FLT_MAX, r_min = FLT_MAX;
  // per-thread d logstd accumulators (A <= 16: registers; larger A: shared memory through a warp reduction per
  // sample).  All loads of a sample are issued before any is used (independent, fully unrolled).
  float gacc[16];
#pragma unroll
  for (int a = 0; a < 16; ++a) gacc[a] = 0.f;
  const bool small_A = A <= 16;
  for (int i0 = blockIdx.x * blockDim.x; i0 < n; i0 += gridDim.x * blockDim.x) {
    const int i = i0 + threadIdx.x;
    const bool live = i < n;
    float coef = 0.f;   // dL/dlp for this sample
    int r = 0;
    float dmv[16];
#pragma unroll
    for (int a = 0; a < 16; ++a) dmv[a] = 0.f;
    if (live) {
      r = idx ? idx[i] : i;
      float lp = 0.f, tlp = 0.f;
      if (small_A) {
        float xa[16], ma[16], ta[16];
        const float* pa = acts + (long long)r * A;
        const float* pm = mean + (long long)i * A;
        const float* pt = tmean + (long long)(t_indexed ? r : i) * A;
#pragma unroll
        for (int a = 0; a < 16; ++a) { const bool on = a < A; xa[a] = on ? pa[a] : 0.f; ma[a] = on ? pm[a] : 0.f; ta[a] = on ? pt[a] : 0.f; }
#pragma unroll
        for (int a = 0; a < 16; ++a) {
          if (a < A) {
            const float dm = xa[a] - ma[a];
            const float dt = xa[a] - ta[a];
            dmv[a] = dm;
            lp += -0.5f * dm * dm * s_ivar[a] - s_ls[a] - HALF_LOG_2PI;
            tlp += -0.5f * dt * dt * s_tivar[a] - s_tls[a] - HALF_LOG_2PI;
          }
        }
      } else {
        for (int a = 0; a < A; ++a) {
          const float x = acts[(long long)r * A + a];
          const float dm = x - mean[(long long)i * A + a];
          const float dt = x - tmean[(long long)(t_indexed ? r : i) * A + a];
          lp += -0.5f * dm * dm * s_ivar[a] - s_ls[a] - HALF_LOG_2PI;
          tlp += -0.5f * dt * dt * s_tivar[a] - s_tls[a] - HALF_LOG_2PI;
        }
      }
      const float ratio = expf(lp - tlp);
      const float ah = (adv[r] - adv_mean) * adv_inv;
      const float s1 = ratio * ah;
      const float rc = fminf(fmaxf(ratio, 1.f - clip), 1.f + clip);
      const float s2 = rc * ah;
      const float inside = (ratio >= 1.f - clip && ratio <= 1.f + clip) ? 1.f : 0.f;
      // L = -mean(min(s2, s1)); torch.min splits the gradient evenly on ties
      float dr;
      if (s1 < s2) dr = ah;
      else if (s2 < s1) dr = ah * inside;
      else dr = 0.5f * ah * (1.f + inside);
      coef = -dr * ratio * inv_global;
      a_loss += (double)(-fminf(s1, s2));
      a_lp += lp; a_lp2 += (double)lp * lp;
      lp_max = fmaxf(lp_max, lp); lp_min = fminf(lp_min, lp);
      r_max = fmaxf(r_max, ratio); r_min = fminf(r_min, ratio);
    }
    if (small_A) {
      if (live) {
        float gm[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) {
          gm[a] = coef * dmv[a] * (a < A ? s_ivar[a] : 0.f);
          if (a < A) {
            d_mean[(long long)i * A + a] = gm[a];
            gacc[a] += coef * (dmv[a] * dmv[a] * s_ivar[a] - 1.f);       // d lp / d logstd_a
          }
        }
        if (d_f16) {                                       // 16-column fp16 row, zero padded: two 16-byte stores
          uint4 w0, w1;
          __half2 h;
          h = __floats2half2_rn(gm[0] * scale_f16, gm[1] * scale_f16); w0.x = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2half2_rn(gm[2] * scale_f16, gm[3] * scale_f16); w0.y = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2half2_rn(gm[4] * scale_f16, gm[5] * scale_f16); w0.z = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2half2_rn(gm[6] * scale_f16, gm[7] * scale_f16); w0.w = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2half2_rn(gm[8] * scale_f16, gm[9] * scale_f16); w1.x = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2half2_rn(gm[10] * scale_f16, gm[11] * scale_f16); w1.y = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2half2_rn(gm[12] * scale_f16, gm[13] * scale_f16); w1.z = *reinterpret_cast<uint32_t*>(&h);
          h = __floats2half2_rn(gm[14] * scale_f16, gm[15] * scale_f16); w1.w = *reinterpret_cast<uint32_t*>(&h);
          reinterpret_cast<uint4*>(d_f16 + (long long)i * 16)[0] = w0;
          reinterpret_cast<uint4*>(d_f16 + (long long)i * 16)[1] = w1;
        }
      }
    } else {
      for (int a = 0; a < A; ++a) {
        float g = 0.f;
        if (live) {
          const float dm = acts[(long long)r * A + a] - mean[(long long)i * A + a];
          const float gm = coef * dm * s_ivar[a];
          d_mean[(long long)i * A + a] = gm;
          g = coef * (dm * dm * s_ivar[a] - 1.f);           // d lp / d logstd_a
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
        if (lane == 0) s_dls[warp][a] += g;
      }
    }
  }
  if (small_A) {                                           // one warp reduction per action dimension, once
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      float g = gacc[a];
#pragma unroll
      for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
      if (lane == 0 && a < A) s_dls[warp][a] = g;
    }
  }
  a_loss = block_reduce(a_loss, OpAddD(), 0.0, shd);
  a_lp = block_reduce(a_lp, OpAddD(), 0.0, shd);
  a_lp2 = block_reduce(a_lp2, OpAddD(), 0.0, shd);
  lp_max = block_reduce(lp_max, OpMaxF(), -FLT_MAX, shf);
  lp_min = block_reduce(lp_min, OpMinF(), FLT_MAX, shf);
  r_max = block_reduce(r_max, OpMaxF(), -FLT_MAX, shf);
  r_min = block_reduce(r_min, OpMinF(), FLT_MAX, shf);
  double* p = part + (long long)blockIdx.x * PF_PART;
  if (threadIdx.x == 0) {
    p[0] = a_loss; p[1] = a_lp; p[2] = a_lp2; p[3] = lp_max; p[4] = lp_min; p[5] = r_max; p[6] = r_min;
  }
  __syncthreads();
  if (threadIdx.x < A) {
    double s = 0.0;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) s += s_dls[w][threadIdx.x];
    p[8 + threadIdx.x] = s;
  }
  if (!last_block_arrives(counter)) return;
  if (threadIdx.x >= 32) return;
  // ---- finalize (last CTA, one warp): every lane sums a strided subset of the per-CTA partials (the loads
  //      pipeline), a fixed shuffle tree combines the lanes; then lane a < A owns d_logstd[a], lane 0 the info row
  const int nparts = (int)gridDim.x;
  float* row = info + (long long)(slot ? *slot : 0) * V4L_INFO_STRIDE;
  double loss = 0.0, slp = 0.0, slp2 = 0.0;
  float lpmax = -FLT_MAX, lpmin = FLT_MAX, rmax = -FLT_MAX, rmin = FLT_MAX;
  double dls[16];
#pragma unroll
  for (int a = 0; a < 16; ++a) dls[a] = 0.0;
  double dls_hi = 0.0;                         // A > 16: lane a handles its own column serially below
  for (int pi = lane; pi < nparts; pi += 32) {
    const double* q = part + (long long)pi * PF_PART;
    loss += __ldcg(q + 0); slp += __ldcg(q + 1); slp2 += __ldcg(q + 2);
    lpmax = fmaxf(lpmax, (float)__ldcg(q + 3)); lpmin = fminf(lpmin, (float)__ldcg(q + 4));
    rmax = fmaxf(rmax, (float)__ldcg(q + 5)); rmin = fminf(rmin, (float)__ldcg(q + 6));
#pragma unroll
    for (int a = 0; a < 16; ++a) if (a < A) dls[a] += __ldcg(q + 8 + a);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, o); slp += __shfl_xor_sync(0xffffffffu, slp, o);
    slp2 += __shfl_xor_sync(0xffffffffu, slp2, o);
    lpmax = fmaxf(lpmax, __shfl_xor_sync(0xffffffffu, lpmax, o)); lpmin = fminf(lpmin, __shfl_xor_sync(0xffffffffu, lpmin, o));
    rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, o)); rmin = fminf(rmin, __shfl_xor_sync(0xffffffffu, rmin, o));
#pragma unroll
    for (int a = 0; a < 16; ++a) dls[a] += __shfl_xor_sync(0xffffffffu, dls[a], o);
  }
  for (int a = lane; a < A; a += 32) {
    double sd = 0.0;
    if (a < 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (j == a) sd = dls[j];
    } else {
      for (int pi = 0; pi < nparts; ++pi) dls_hi += __ldcg(part + (long long)pi * PF_PART + 8 + a);
      sd = dls_hi;
    }
    const float raw = logstd[a];
    const float pass = (raw >= -5.f && raw <= 2.f) ? 1.f : 0.f;      // clamp backward
    // entropy = sum_a (0.5 + 0.5 log 2pi + logstd_a): d(-c * mean ent)/d logstd_a = -c over the GLOBAL
    // minibatch; this rank contributes its share n / (n * world) of it (the gradient buckets are
    // SUM-all-reduced across ranks, like the surrogate part which already carries inv_global)
    d_logstd[a] = pass * ((float)sd - entropy_coeff * ((float)n * inv_global));
  }
  if (lane == 0) {
    double ent = 0.0, ls_s = 0.0, ls_s2 = 0.0;
    float ls_max = -FLT_MAX, ls_min = FLT_MAX;
    for (int a = 0; a < A; ++a) {
      const float ls = fminf(fmaxf(logstd[a], -5.f), 2.f);
      ent += 0.5 + 0.91893853320467274178 + (double)ls;
      ls_s += ls; ls_s2 += (double)ls * ls;
      ls_max = fmaxf(ls_max, ls); ls_min = fminf(ls_min, ls);
    }
    const double an = adv_stats[2];
    const double amean = adv_stats[0] / an;
    const double avar = fmax((adv_stats[1] - adv_stats[0] * adv_stats[0] / an) / (an - 1.0), 0.0);
    row[V4L_INFO_ADV_MEAN] = (float)amean;
    row[V4L_INFO_ADV_STD] = (float)sqrt(avar);
    row[V4L_INFO_ADV_MAX] = (float)adv_stats[3];
    row[V4L_INFO_ADV_MIN] = (float)adv_stats[4];
    row[V4L_INFO_POLICY_LOSS] = (float)(loss * inv_local - entropy_coeff * ent);
    const double lpm = slp / n;
    row[V4L_INFO_LP_MEAN] = (float)lpm;
    row[V4L_INFO_LP_STD] = (float)sqrt(fmax((slp2 - slp * slp / n) / (n - 1.0), 0.0));
    row[V4L_INFO_LP_MAX] = lpmax;
    row[V4L_INFO_LP_MIN] = lpmin;
    row[V4L_INFO_LS_MEAN] = (float)(ls_s / A);
    row[V4L_INFO_LS_STD] = (float)sqrt(fmax((ls_s2 - ls_s * ls_s / A) / (A - 1.0), 0.0));
    row[V4L_INFO_LS_MAX] = ls_max;
    row[V4L_INFO_LS_MIN] = ls_min;
    row[V4L_INFO_RATIO_MAX] = rmax;
    row[V4L_INFO_RATIO_MIN] = rmin;
  }
}

// =============================================================================================
// clip_grad_norm_ + Adam over a flat bucket
// =============================================================================================
constexpr int ADAM_THREADS = 256;

// sum of the per-CTA squared-norm partials by one warp: lane-strided partial sums, then a fixed
// shuffle tree (the same order wherever it is called, so every CTA derives the same clip factor)
__device__ __forceinline__ double sum_parts(const double* __restrict__ part, int nparts) {
  double s = 0.0;
  for (int i = threadIdx.x & 31; i < nparts; i += 32) s += part[i];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

__global__ void __launch_bounds__(ADAM_THREADS)
sqnorm_kernel(const float* __restrict__ g, long long n, double* __restrict__ part) {
  v4l_pdl_enter();
  __shared__ double shd[32];
  double s = 0.0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const double x = g[i];
    s += x * x;
  }
  s = block_reduce(s, OpAddD(), 0.0, shd);
  if (threadIdx.x == 0) part[blockIdx.x] = s;
}

__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, long long n, const float* __restrict__ hyper,
            const double* __restrict__ part, int nparts, int vec_ok) {
  v4l_pdl_enter();
  __shared__ float s_coef;
  if (threadIdx.x < 32) {
    const double s = sum_parts(part, nparts);             // same order in every CTA
    if (threadIdx.x == 0) {
      const float total = (float)sqrt(s);
      // a non-finite gradient norm (fp16 overflow in the tensor-core tier) would poison the Adam
      // moments for good: skip this step instead (coef < 0 marks it; the norm is still logged)
      s_coef = isfinite(total) ? fminf(hyper[4] / (total + 1e-6f), 1.f) : -1.f;    // clip_grad_norm_
    }
  }
  __syncthreads();
  const float coef = s_coef;
  if (coef < 0.f) return;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3];
  const float step = hyper[5] + 1.f;
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  // four parameters per thread and iteration (the bucket base is 16-byte aligned); element-wise math
  // identical to the scalar form
  const long long n4 = vec_ok ? (n >> 2) : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i], p4 = reinterpret_cast<float4*>(p)[i];
    float* gm = reinterpret_cast<float*>(&m4); float* gv = reinterpret_cast<float*>(&v4); float* gp = reinterpret_cast<float*>(&p4);
    const float* gg = reinterpret_cast<const float*>(&g4);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gi = gg[j] * coef;
      const float mi = b1 * gm[j] + (1.f - b1) * gi;
      const float vi = b2 * gv[j] + (1.f - b2) * gi * gi;
      gm[j] = mi; gv[j] = vi;
      const float denom = sqrtf(vi) / bc2_sqrt + eps;
      gp[j] -= step_size * (mi / denom);
    }
    reinterpret_cast<float4*>(m)[i] = m4; reinterpret_cast<float4*>(v)[i] = v4; reinterpret_cast<float4*>(p)[i] = p4;
  }
  for (long long i = (n4 << 2) + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}

__global__ void adam_finish_kernel(float* hyper, const double* __restrict__ part, int nparts,
                                   float* info, const int32_t* slot, int norm_slot) {
  v4l_pdl_enter();
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    const double s = sum_parts(part, nparts);
    if (threadIdx.x == 0) {
      if (isfinite((float)sqrt(s))) hyper[5] += 1.f;
      if (info && norm_slot >= 0)
        info[(long long)(slot ? *slot : 0) * V4L_INFO_STRIDE + norm_slot] = (float)sqrt(s);
    }
  }
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" int v4l_gae(v4l_ctx* ctx, void* stream, const float* rewards, const float* values,
                       const float* terminals, const float* time_limits, int64_t tl_st, int64_t tl_se,
                       const float* last_value, float* advs, float* rets, int T, int E,
                       double gamma, double tau, int time_limit_filter, int mode) {
  V4L_REQUIRE(ctx && rewards && values && terminals && last_value && advs && rets, "v4l_gae: NULL argument");
  V4L_REQUIRE(!time_limit_filter || time_limits, "v4l_gae: time_limit_filter set but time_limits is NULL");
  V4L_REQUIRE(T >= 0 && E >= 0 && E <= 65535, "v4l_gae: bad shape T=%d E=%d", T, E);
  V4L_REQUIRE(mode == 0 || mode == 1, "v4l_gae: bad mode %d", mode);
  if (T == 0 || E == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  GaeArgs g;
  g.r = rewards; g.v = values; g.d = terminals; g.tl = time_limits; g.last_value = last_value;
  g.tl_st = tl_st; g.tl_se = tl_se; g.advs = advs; g.rets = rets; g.T = T; g.E = E;
  g.gamma = gamma; g.tau = tau; g.use_tl = time_limit_filter ? 1 : 0; g.mode = mode;
  // enough (column, chunk) CTAs to cover the machine ~2x; chunk length a multiple of the tile
  int want = v4l_cdiv(2LL * ctx->sm_count, E);
  int chunk = v4l_cdiv(T, want);
  chunk = ((chunk + GAE_THREADS - 1) / GAE_THREADS) * GAE_THREADS;
  g.chunk_len = chunk;
  g.n_chunks = v4l_cdiv(T, chunk);
  dim3 grid(g.n_chunks, E);
  if (g.n_chunks == 1) {
    V4L_LAUNCH(gae_scan_kernel, grid, GAE_THREADS, 0, s, g, nullptr);
    V4L_CHECK_LAUNCH();
    return 0;
  }
  const size_t need = ((size_t)E * g.n_chunks * 3) * sizeof(double);
  V4L_REQUIRE(need <= ctx->scratch_elems * sizeof(float), "v4l_gae: scratch too small");
  AB* agg = reinterpret_cast<AB*>(ctx->scratch);
  double* carry = reinterpret_cast<double*>(ctx->scratch) + (size_t)E * g.n_chunks * 2;
  V4L_LAUNCH(gae_aggregate_kernel, grid, GAE_THREADS, 0, s, g, agg);
  V4L_CHECK_LAUNCH();
  V4L_LAUNCH(gae_carry_kernel, v4l_cdiv(E, 128), 128, 0, s, g, agg, carry);
  V4L_CHECK_LAUNCH();
  V4L_LAUNCH(gae_scan_kernel, grid, GAE_THREADS, 0, s, g, carry);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_select_rows(v4l_ctx* ctx, void* stream, const int32_t* flat_idx, const int32_t* slot,
                               int32_t* cur_idx, int n) {
  V4L_REQUIRE(ctx && flat_idx && slot && cur_idx && n >= 0, "v4l_select_rows: bad argument");
  if (n == 0) return 0;
  V4L_LAUNCH(select_rows_kernel, min(v4l_cdiv(n, 256), 4 * ctx->sm_count), 256, 0, (cudaStream_t)stream, flat_idx, slot, cur_idx, n);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_slot_advance(v4l_ctx* ctx, void* stream, int32_t* slot, int32_t wrap) {
  V4L_REQUIRE(ctx && slot, "v4l_slot_advance: NULL argument");
  V4L_LAUNCH(slot_advance_kernel, 1, 1, 0, (cudaStream_t)stream, slot, wrap);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_adv_stats(v4l_ctx* ctx, void* stream, const float* adv, const int32_t* idx, int n,
                             double* stats) {
  V4L_REQUIRE(ctx && adv && stats && n > 0, "v4l_adv_stats: bad argument");
  V4L_LAUNCH(adv_stats_kernel, 1, 1024, 0, (cudaStream_t)stream, adv, idx, n, stats);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_vf_loss(v4l_ctx* ctx, void* stream, const float* values, const float* returns,
                           const float* old_values, const int32_t* idx, float* d_values, int n,
                           float inv_global, float inv_local, int clipped, float clip_para,
                           float* info, const int32_t* slot, void* d_values_f16, float scale_f16) {
  V4L_REQUIRE(ctx && values && returns && d_values && info && n > 0, "v4l_vf_loss: bad argument");
  V4L_REQUIRE(!clipped || old_values, "v4l_vf_loss: clipped loss needs old_values");
  cudaStream_t s = (cudaStream_t)stream;
  const int ctas = min(v4l_cdiv(n, LOSS_THREADS), 2 * ctx->sm_count);
  double* part = reinterpret_cast<double*>(ctx->scratch);
  V4L_LAUNCH(vf_loss_kernel, ctas, LOSS_THREADS, 0, s, values, returns, old_values, idx, d_values,
             reinterpret_cast<__half*>(d_values_f16), scale_f16, n, inv_global, inv_local, clipped, clip_para, part,
             ctx->counters + 0, info, slot);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_pf_loss(v4l_ctx* ctx, void* stream, const float* mean, const float* logstd,
                           const float* target_mean, const float* target_logstd, const float* acts,
                           const float* adv, const int32_t* idx, const double* adv_stats,
                           float* d_mean, float* d_logstd, int n, int A, float inv_global,
                           float inv_local, float clip_para, float entropy_coeff, float* info,
                           const int32_t* slot, int target_indexed, void* d_mean_f16, float scale_f16,
                           int stats_per_slot) {
  V4L_REQUIRE(ctx && mean && logstd && target_mean && target_logstd && acts && adv && adv_stats &&
              d_mean && d_logstd && info, "v4l_pf_loss: NULL argument");
  V4L_REQUIRE(n > 0 && A > 0 && A <= MAX_A, "v4l_pf_loss: bad shape n=%d A=%d (A <= %d)", n, A, MAX_A);
  cudaStream_t s = (cudaStream_t)stream;
  const int ctas = min(v4l_cdiv(n, LOSS_THREADS), 2 * ctx->sm_count);
  double* part = reinterpret_cast<double*>(ctx->scratch);
  V4L_REQUIRE(!d_mean_f16 || A <= 16, "v4l_pf_loss: the fp16 gradient row holds 16 columns (A = %d)", A);
  V4L_LAUNCH(pf_loss_kernel, ctas, LOSS_THREADS, 0, s, mean, logstd, target_mean, target_logstd, acts, adv, idx,
             adv_stats, d_mean, reinterpret_cast<__half*>(d_mean_f16), scale_f16, n, A, inv_global, inv_local,
             clip_para, entropy_coeff, part, (target_indexed && idx) ? 1 : 0, d_logstd, ctx->counters + 1, info, slot,
             stats_per_slot);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_clip_adam(v4l_ctx* ctx, void* stream, float* param, const float* grad, float* m,
                             float* v, int64_t n, float* hyper, float* info, const int32_t* slot,
                             int norm_slot) {
  V4L_REQUIRE(ctx && param && grad && m && v && hyper && n > 0, "v4l_clip_adam: bad argument");
  V4L_REQUIRE(norm_slot < V4L_INFO_STRIDE, "v4l_clip_adam: bad norm_slot");
  cudaStream_t s = (cudaStream_t)stream;
  const int ctas = (int)min((long long)2 * ctx->sm_count, (long long)((n + ADAM_THREADS - 1) / ADAM_THREADS));
  double* part = reinterpret_cast<double*>(ctx->scratch);
  V4L_LAUNCH(sqnorm_kernel, ctas, ADAM_THREADS, 0, s, grad, n, part);
  V4L_CHECK_LAUNCH();
  const int vec_ok = (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  V4L_LAUNCH(adam_kernel, ctas, ADAM_THREADS, 0, s, param, grad, m, v, n, hyper, part, ctas, vec_ok);
  V4L_CHECK_LAUNCH();
  V4L_LAUNCH(adam_finish_kernel, 1, 32, 0, s, hyper, part, ctas, info, slot, norm_slot);
  V4L_CHECK_LAUNCH();
  return 0;
}
