// Minibatch prologue and the fused optimiser tail of the tensor-core tier.
//
// One PPO minibatch (reference torchrl/algo/on_policy/ppo.py:125-153) is a strictly serial chain:
// critic forward -> loss -> backward -> clip + Adam -> actor forward (on the encoder the critic just
// stepped) -> loss -> backward -> clip + Adam.  At the benchmark's minibatch of 1024 the step is bound
// by the LENGTH of that chain, so the small HBM/latency-bound links are merged:
//
//   v4l_mb_begin   row-index selection + advantage statistics + proprio rows -> fp16 (was 3 launches)
//   v4l_opt_tail   two launches (was 5-6): (1) split-K reduction of the weight-gradient partials, which also
//                  accumulates the squared norm of what it writes; (2) global-norm clip + Adam, each updated
//                  weight written straight into the fp16 operand copies the NEXT forward passes read, then
//                  the step / slot counters.  (A single kernel with a device-wide barrier between the two
//                  was measured SLOWER — 36 us against 12: one 1024-thread CTA per SM has too little
//                  memory-level parallelism for the latency-bound reduction.)
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>
#include <string.h>

#include "common.cuh"

namespace {

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce_t(T v, Op op, T ident, T* sh /* [32] */) {
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
struct AddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct MaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct MinF { __device__ float operator()(float a, float b) const { return fminf(a, b); } };

// =================================================================================================
// minibatch prologue
// =================================================================================================
constexpr int MB_THREADS = 256;

__global__ void __launch_bounds__(MB_THREADS)
mb_begin_kernel(const int32_t* __restrict__ flat_idx, const int32_t* __restrict__ slot, int32_t* __restrict__ cur,
                int n, const float* __restrict__ adv, double* __restrict__ stats, double* part,
                unsigned int* counter, const float* __restrict__ state, int S, __half* __restrict__ st16, int Sp) {
  v4l_pdl_enter();
  __shared__ double shd[32];
  __shared__ float shf[32];
  __shared__ bool s_last;
  const long long base = (long long)(*slot) * n;
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
  // (1) this minibatch's row list (reference replay_buffers/on_policy.py:76-89) + advantage partials
  double s = 0.0, s2 = 0.0;
  float mx = -FLT_MAX, mn = FLT_MAX;
  for (int i = gtid; i < n; i += gsz) {
    const int r = flat_idx[base + i];
    cur[i] = r;
    const float a = adv[r];
    s += a; s2 += (double)a * a;
    mx = fmaxf(mx, a); mn = fminf(mn, a);
  }
  // (2) proprio rows -> fp16, zero padded to Sp columns (the first Linear's K is a multiple of 64): one unit =
  //     8 columns of one row -> one index load, 8 independent loads, one 16-byte store
  if (st16) {
    const int cpr = Sp >> 3;                       // 8-column chunks per row (Sp is a multiple of 64)
    const long long units = (long long)n * cpr;
    for (long long u = gtid; u < units; u += gsz) {
      const int r = (int)(u / cpr), c0 = (int)(u - (long long)r * cpr) << 3;
      const float* src = state + (long long)flat_idx[base + r] * S + c0;
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (c0 + j < S) ? src[j] : 0.f;
      uint4 w;
      __half2 h;
      h = __floats2half2_rn(f[0], f[1]); w.x = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2half2_rn(f[2], f[3]); w.y = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2half2_rn(f[4], f[5]); w.z = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2half2_rn(f[6], f[7]); w.w = *reinterpret_cast<uint32_t*>(&h);
      *reinterpret_cast<uint4*>(st16 + (long long)r * Sp + c0) = w;
    }
  }
  // (3) advantage statistics {sum, sumsq, n, max, min}: per-CTA partials, summed in CTA order by the
  //     last CTA to arrive (deterministic).  stats == NULL: they were computed for the whole epoch up front
  //     (v4l_adv_stats_epoch) and nothing is left to do here.
  if (!stats) return;
  s = block_reduce_t(s, AddD(), 0.0, shd);
  s2 = block_reduce_t(s2, AddD(), 0.0, shd);
  mx = block_reduce_t(mx, MaxF(), -FLT_MAX, shf);
  mn = block_reduce_t(mn, MinF(), FLT_MAX, shf);
  if (threadIdx.x == 0) {
    double* p = part + 4 * blockIdx.x;
    p[0] = s; p[1] = s2; p[2] = mx; p[3] = mn;
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  __threadfence();
  // lane-strided sums (the loads pipeline) + fixed shuffle tree
  double S1 = 0.0, S2 = 0.0;
  float MX = -FLT_MAX, MN = FLT_MAX;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) {
    S1 += __ldcg(part + 4 * b); S2 += __ldcg(part + 4 * b + 1);
    MX = fmaxf(MX, (float)__ldcg(part + 4 * b + 2)); MN = fminf(MN, (float)__ldcg(part + 4 * b + 3));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    S1 += __shfl_xor_sync(0xffffffffu, S1, o); S2 += __shfl_xor_sync(0xffffffffu, S2, o);
    MX = fmaxf(MX, __shfl_xor_sync(0xffffffffu, MX, o)); MN = fminf(MN, __shfl_xor_sync(0xffffffffu, MN, o));
  }
  if (threadIdx.x == 0) { stats[0] = S1; stats[1] = S2; stats[2] = (double)n; stats[3] = MX; stats[4] = MN; }
}

// Advantage statistics of EVERY minibatch of an epoch in one launch: the advantages and the row lists are fixed
// once GAE has run and the permutations are drawn, so the per-minibatch statistics (and, data parallel, their
// exchange across ranks) leave the per-minibatch chain.  stats[mb] = {sum, sumsq, n, max, min, -, -, -}.
__global__ void __launch_bounds__(256) adv_stats_epoch_kernel(const int32_t* __restrict__ flat_idx, int n,
                                                              const float* __restrict__ adv, double* __restrict__ stats) {
  v4l_pdl_enter();
  __shared__ double shd[32];
  __shared__ float shf[32];
  const int32_t* idx = flat_idx + (long long)blockIdx.x * n;
  double s = 0.0, s2 = 0.0;
  float mx = -FLT_MAX, mn = FLT_MAX;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float a = adv[idx[i]];
    s += a; s2 += (double)a * a;
    mx = fmaxf(mx, a); mn = fminf(mn, a);
  }
  s = block_reduce_t(s, AddD(), 0.0, shd);
  s2 = block_reduce_t(s2, AddD(), 0.0, shd);
  mx = block_reduce_t(mx, MaxF(), -FLT_MAX, shf);
  mn = block_reduce_t(mn, MinF(), FLT_MAX, shf);
  if (threadIdx.x == 0) {
    double* o = stats + (long long)blockIdx.x * 8;
    o[0] = s; o[1] = s2; o[2] = (double)n; o[3] = mx; o[4] = mn; o[5] = 0.0; o[6] = 0.0; o[7] = 0.0;
  }
}

// =================================================================================================
// fused optimiser tail
// =================================================================================================
constexpr int RED_THREADS = 256;     // reduction kernel: many small CTAs (memory-level parallelism)
constexpr int STEP_THREADS = 256;

struct TailParams {
  v4l_reduce_job jobs[V4L_MAX_JOBS];
  int job_first[V4L_MAX_JOBS + 1];   // prefix sums of lane-items (outputs x lanes, padded to whole warps) per job
  int job_lanes_log2[V4L_MAX_JOBS];  // log2 of the lanes that share one output's split sum
  int n_jobs;
  int phases;                        // bit 0: reduce, bit 1: norm + clip + Adam (+ fp16 operand copies)
  float* p; float* g; float* m; float* v; long long n;
  float* hyper;                      // {lr, b1, b2, eps, max_norm, step, ...}
  float* info; const int32_t* slot; int norm_slot;
  long long extra_lo, extra_n;       // gradient range NOT produced by a reduction job (logstd), for the norm
  const int4* scatter;               // [n] packed fp16 positions of parameter i: (self a, self b, other a, other b), -1 = none
  __half* packed_self; __half* packed_other;
  int32_t* slot_advance;             // optional: minibatch slot counter to increment at the very end
  double* part;                      // [gridDim.x] squared-norm partials
  unsigned int* bar;                 // arrival counter of the step kernel (re-armed by its last CTA)
};

// dw[index[n*Kp + kp]] = scale * sum_split partial[split][kp / 128][n][kp % 128]  (+ bias rows).  Work unit =
// 4 consecutive kp of one n (a float4 of the kp-fastest partial layout; Kp is a multiple of 64) shared by a
// group of 2^lanes_log2 lanes: each lane sums a strided subset of the splits (<= 4 independent 16-byte loads in
// flight), then a fixed shuffle tree combines the lanes.  A bias output is a unit with one live component.
// Returns this thread's sum of squares of the gradient values it wrote (for the global norm).
__device__ __forceinline__ double reduce_phase(const TailParams& P) {
  const int total = P.job_first[P.n_jobs];
  double sq = 0.0;
  for (int base = blockIdx.x * RED_THREADS; base < total; base += gridDim.x * RED_THREADS) {
    const int item = base + threadIdx.x;
    // a job's items are padded to whole warps, so all lanes of a warp work on the same job (same L)
    int j = 0;
    bool live = item < total;
    if (live) {
      int lo = 0, hi = P.n_jobs;                // job_first[lo] <= item < job_first[hi]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (P.job_first[mid] <= item) lo = mid; else hi = mid; }
      j = lo;
    }
    const v4l_reduce_job& J = P.jobs[j];
    const int ll = P.job_lanes_log2[j];
    const int L = 1 << ll;
    const int rel = live ? item - P.job_first[j] : 0;
    const int u = rel >> ll, l = rel & (L - 1);
    const int units_w = (J.N_valid * J.Kp) >> 2;
    live = live && u < units_w + (J.has_bias ? J.N_valid : 0);
    const long long split_stride = (long long)(J.kin_tiles + J.has_bias) * 128 * J.Nmma;
    const float* src = J.partial;
    int4 dst = make_int4(-1, -1, -1, -1);
    float* out = J.dw;
    const bool is_w = u < units_w;
    if (live) {
      if (is_w) {
        const int o = u << 2;
        const int n = o / J.Kp, kp = o - n * J.Kp;
        src = J.partial + ((long long)(kp >> 7) * J.Nmma + n) * 128 + (kp & 127);
        dst = J.index ? __ldg(reinterpret_cast<const int4*>(J.index + o)) : make_int4(o, o + 1, o + 2, o + 3);
      } else {                                  // bias gradient: K slice `kin_tiles`, lane 0 of the row
        const int n = u - units_w;
        src = J.partial + ((long long)J.kin_tiles * J.Nmma + n) * 128;
        dst.x = n; out = J.dbias;
      }
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
      // (the partial loads do not wait for the index load: one memory round trip, not two)
      for (int z0 = l; z0 < J.splits; z0 += 4 * L) {
        float4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int z = z0 + q * L;
          v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (z < J.splits) {
            const float* a = src + (long long)z * split_stride;
            if (is_w) v[q] = __ldcg(reinterpret_cast<const float4*>(a));
            else v[q].x = __ldcg(a);
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
      }
    }
    // combine the L lanes of a unit (L <= 32 divides the warp; same tree for every output)
    for (int off = 1; off < L; off <<= 1) {
      acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
      acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off); acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
    }
    if (l == 0 && live) {
      const float r4[4] = {acc.x, acc.y, acc.z, acc.w};
      const int d4[4] = {dst.x, dst.y, dst.z, dst.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (d4[q] < 0) continue;
        float r = r4[q] * J.scale;
        if (J.accumulate) r += out[d4[q]];
        out[d4[q]] = r;
        sq += (double)r * r;
      }
    }
  }
  return sq;
}

// ---- kernel 1: split-K reduction (+ squared norm of what it writes), or the bucket's squared norm alone
__global__ void __launch_bounds__(RED_THREADS) opt_reduce_kernel(const __grid_constant__ TailParams P) {
  v4l_pdl_enter();
  __shared__ double shd[32];
  const long long gtid = blockIdx.x * (long long)RED_THREADS + threadIdx.x;
  const long long gsz = (long long)gridDim.x * RED_THREADS;
  double sq = 0.0;
  if (P.phases & 1) {
    sq = reduce_phase(P);
    if (P.phases & 2)    // gradient range no reduction job writes (logstd, written by the loss kernel)
      for (long long i = gtid; i < P.extra_n; i += gsz) { const double x = __ldcg(P.g + P.extra_lo + i); sq += x * x; }
  } else {
    for (long long i = gtid; i < P.n; i += gsz) { const double x = __ldcg(P.g + i); sq += x * x; }
  }
  if (P.phases & 2) {
    sq = block_reduce_t(sq, AddD(), 0.0, shd);
    if (threadIdx.x == 0) P.part[blockIdx.x] = sq;
  }
}

// ---- kernel 2: clip_grad_norm_ + Adam + fp16 operand copies + counters
__global__ void __launch_bounds__(STEP_THREADS) opt_step_kernel(const __grid_constant__ TailParams P, int nparts) {
  v4l_pdl_enter();
  __shared__ float s_coef;
  const long long gtid = blockIdx.x * (long long)STEP_THREADS + threadIdx.x;
  const long long gsz = (long long)gridDim.x * STEP_THREADS;
  // every CTA derives the same clip factor from the partials, summed in the same (fixed) order by the whole CTA
  {
    __shared__ double shd[32];
    double t = 0.0;
    for (int i = threadIdx.x; i < nparts; i += STEP_THREADS) t += __ldcg(P.part + i);
    t = block_reduce_t(t, AddD(), 0.0, shd);
    if (threadIdx.x == 0) {
      const float total = (float)sqrt(t);
      s_coef = isfinite(total) ? fminf(P.hyper[4] / (total + 1e-6f), 1.f) : -1.f;
      if (blockIdx.x == 0 && P.info && P.norm_slot >= 0)
        P.info[(long long)(P.slot ? *P.slot : 0) * V4L_INFO_STRIDE + P.norm_slot] = total;
    }
  }
  __syncthreads();
  const float coef = s_coef;
  if (coef >= 0.f) {                       // a non-finite norm skips the step (moments stay clean)
    const float lr = P.hyper[0], b1 = P.hyper[1], b2 = P.hyper[2], eps = P.hyper[3];
    const float step = P.hyper[5] + 1.f;
    const float bc1 = 1.f - powf(b1, step);
    const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
    const float step_size = lr / bc1;
    const long long n4 = P.n >> 2;         // buckets are 16-byte aligned and padded to 4 floats
    for (long long i = gtid; i < n4; i += gsz) {
      const float4 g4 = __ldcg(reinterpret_cast<const float4*>(P.g) + i);
      float4 m4 = reinterpret_cast<float4*>(P.m)[i], v4 = reinterpret_cast<float4*>(P.v)[i];
      float4 p4 = reinterpret_cast<float4*>(P.p)[i];
      float* gm = reinterpret_cast<float*>(&m4); float* gv = reinterpret_cast<float*>(&v4);
      float* gp = reinterpret_cast<float*>(&p4);
      const float* gg = reinterpret_cast<const float*>(&g4);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float gi = gg[q] * coef;
        const float mi = b1 * gm[q] + (1.f - b1) * gi;
        const float vi = b2 * gv[q] + (1.f - b2) * gi * gi;
        gm[q] = mi; gv[q] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        gp[q] -= step_size * (mi / denom);
      }
      reinterpret_cast<float4*>(P.m)[i] = m4; reinterpret_cast<float4*>(P.v)[i] = v4;
      reinterpret_cast<float4*>(P.p)[i] = p4;
      if (P.scatter) {
        // fp16 operand copies of the weights (tap-major forward + data-gradient orientations) the next
        // forward passes read: this network's own and, for shared-encoder weights, the other network's
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int4 pos = __ldg(P.scatter + 4 * i + q);
          const __half h = __float2half(gp[q]);
          if (pos.x >= 0) P.packed_self[pos.x] = h;
          if (pos.y >= 0) P.packed_self[pos.y] = h;
          if (pos.z >= 0) P.packed_other[pos.z] = h;
          if (pos.w >= 0) P.packed_other[pos.w] = h;
        }
      }
    }
    for (long long i = (n4 << 2) + gtid; i < P.n; i += gsz) {
      const float gi = __ldcg(P.g + i) * coef;
      const float mi = b1 * P.m[i] + (1.f - b1) * gi;
      const float vi = b2 * P.v[i] + (1.f - b2) * gi * gi;
      P.m[i] = mi; P.v[i] = vi;
      const float pn = P.p[i] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
      P.p[i] = pn;
      if (P.scatter) {
        const int4 pos = __ldg(P.scatter + i);
        const __half h = __float2half(pn);
        if (pos.x >= 0) P.packed_self[pos.x] = h;
        if (pos.y >= 0) P.packed_self[pos.y] = h;
        if (pos.z >= 0) P.packed_other[pos.z] = h;
        if (pos.w >= 0) P.packed_other[pos.w] = h;
      }
    }
  }
  // the CTA that finishes last moves the counters (every CTA has read the Adam step count by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(P.bar, 1u);
    if (t == gridDim.x - 1) {
      P.bar[0] = 0u;
      if (coef >= 0.f) P.hyper[5] += 1.f;
      if (P.slot_advance) *P.slot_advance += 1;
      __threadfence();
    }
  }
}

}  // namespace

extern "C" int v4l_mb_begin(v4l_ctx* ctx, void* stream, const int32_t* flat_idx, const int32_t* slot,
                            int32_t* cur_idx, int n, const float* adv, double* stats, const float* state,
                            int S, void* state_f16, int Sp) {
  V4L_REQUIRE(ctx && flat_idx && slot && cur_idx && adv && n > 0, "v4l_mb_begin: bad argument");
  V4L_REQUIRE(!state_f16 || (S >= 0 && Sp >= S && (S == 0 || state)), "v4l_mb_begin: bad proprio arguments");
  V4L_REQUIRE(!state_f16 || Sp % 8 == 0, "v4l_mb_begin: Sp must be a multiple of 8");
  const long long work = max((long long)n, state_f16 ? (long long)n * Sp / 8 : 0LL);
  const int ctas = max(1, min(2 * ctx->sm_count, v4l_cdiv(work, MB_THREADS)));
  double* part = reinterpret_cast<double*>(ctx->scratch);
  V4L_LAUNCH(mb_begin_kernel, ctas, MB_THREADS, 0, (cudaStream_t)stream, flat_idx, slot, cur_idx, n, adv, stats, part,
             ctx->counters + 2, state, S, reinterpret_cast<__half*>(state_f16), Sp);
  V4L_CHECK_LAUNCH();
  return 0;
}

// lanes sharing one unit's split sum: <= 4 splits per lane (one round of independent loads)
static int lanes_log2_for(int splits) {
  int ll = 0;
  while ((1 << ll) < 32 && ((splits + (1 << ll) - 1) >> ll) > 4) ++ll;
  return ll;
}

extern "C" int v4l_opt_tail(v4l_ctx* ctx, void* stream, const v4l_opt_tail_args* a) {
  V4L_REQUIRE(ctx && a, "v4l_opt_tail: NULL argument");
  const int phases = a->phases;
  V4L_REQUIRE(phases > 0 && phases < 4, "v4l_opt_tail: bad phases %d", phases);
  V4L_REQUIRE(!(phases & 2) || (a->param && a->grad && a->m && a->v && a->hyper && a->n > 0),
              "v4l_opt_tail: the optimiser phase needs param/grad/m/v/hyper");
  V4L_REQUIRE(!(phases & 2) || ((((uintptr_t)a->param | (uintptr_t)a->grad | (uintptr_t)a->m | (uintptr_t)a->v) & 15) == 0),
              "v4l_opt_tail: buckets must be 16-byte aligned");
  V4L_REQUIRE(a->norm_slot < V4L_INFO_STRIDE, "v4l_opt_tail: bad norm_slot");
  V4L_REQUIRE(!a->scatter || (a->packed_self && a->packed_other && (((uintptr_t)a->scatter) & 15) == 0),
              "v4l_opt_tail: bad scatter arguments");
  V4L_REQUIRE(a->extra_n >= 0 && a->extra_lo >= 0 && a->extra_lo + a->extra_n <= (a->n > 0 ? a->n : 0) + 0,
              "v4l_opt_tail: bad extra range");
  if (phases == 3 && ctx->early_flush) {
    // some gradients were already written by an earlier flush: reduce first, then take the norm from the bucket
    ctx->early_flush = 0;
    v4l_opt_tail_args t = *a;
    t.phases = 1;
    if (int r = v4l_opt_tail(ctx, stream, &t)) return r;
    t.phases = 2;
    return v4l_opt_tail(ctx, stream, &t);
  }
  if (phases & 2) ctx->early_flush = 0;
  if ((phases & 1) && ctx->n_jobs == 0) {        // nothing pending: the norm comes from the bucket
    if (phases == 1) return 0;
    v4l_opt_tail_args t = *a;
    t.phases = 2;
    return v4l_opt_tail(ctx, stream, &t);
  }
  static TailParams P;               // large: keep it off the stack (single host thread per context)
  memset(&P, 0, sizeof(P));
  int total = 0;
  if (phases & 1) {
    for (int i = 0; i < ctx->n_jobs; ++i) {
      const v4l_reduce_job& J = ctx->jobs[i];
      P.jobs[i] = J;
      P.job_first[i] = total;
      const int ll = lanes_log2_for(J.splits);
      P.job_lanes_log2[i] = ll;
      V4L_REQUIRE(J.Kp % 4 == 0, "v4l_opt_tail: packed K (%d) must be a multiple of 4", J.Kp);
      const long long outs = (((long long)J.N_valid * J.Kp) >> 2) + (J.has_bias ? J.N_valid : 0);    // units
      V4L_REQUIRE(((outs << ll) + total + 32) < (1LL << 31), "v4l_opt_tail: too many reduction items");
      total += (int)((((outs << ll) + 31) >> 5) << 5);     // whole warps per job: a warp never straddles two jobs
    }
    P.n_jobs = ctx->n_jobs;
    P.job_first[P.n_jobs] = total;
    ctx->n_jobs = 0;
    ctx->defer_cursor = 0;
  }
  P.phases = phases;
  P.p = a->param; P.g = a->grad; P.m = a->m; P.v = a->v; P.n = a->n;
  P.hyper = a->hyper; P.info = a->info; P.slot = a->slot; P.norm_slot = a->norm_slot;
  P.extra_lo = a->extra_lo; P.extra_n = a->extra_n;
  P.scatter = reinterpret_cast<const int4*>(a->scatter);
  P.packed_self = reinterpret_cast<__half*>(a->packed_self);
  P.packed_other = reinterpret_cast<__half*>(a->packed_other);
  P.slot_advance = a->slot_advance;
  P.part = reinterpret_cast<double*>(ctx->scratch);
  P.bar = ctx->counters + 4;
  cudaStream_t st = (cudaStream_t)stream;
  // kernel 1 (reduction and / or squared norm): enough 256-thread CTAs to keep ~8 resident per SM
  const long long work1 = (phases & 1) ? (long long)total : a->n;
  const long long want1 = (work1 + RED_THREADS - 1) / RED_THREADS, cap1 = (long long)8 * ctx->sm_count;
  const int grid1 = (int)(want1 < 1 ? 1 : (want1 > cap1 ? cap1 : want1));
  if (!(phases & 1) || total > 0 || (phases & 2)) {
    V4L_LAUNCH(opt_reduce_kernel, grid1, RED_THREADS, 0, st, P);
    V4L_CHECK_LAUNCH();
  }
  if (phases & 2) {
    const long long want2 = ((a->n >> 2) + STEP_THREADS - 1) / STEP_THREADS, cap2 = (long long)4 * ctx->sm_count;
    const int grid2 = (int)(want2 < 1 ? 1 : (want2 > cap2 ? cap2 : want2));
    V4L_LAUNCH(opt_step_kernel, grid2, STEP_THREADS, 0, st, P, grid1);
    V4L_CHECK_LAUNCH();
  }
  return 0;
}

extern "C" int v4l_adv_stats_epoch(v4l_ctx* ctx, void* stream, const int32_t* flat_idx, int n_minibatches, int n,
                                   const float* adv, double* stats) {
  V4L_REQUIRE(ctx && flat_idx && adv && stats && n > 0 && n_minibatches > 0, "v4l_adv_stats_epoch: bad argument");
  V4L_LAUNCH(adv_stats_epoch_kernel, n_minibatches, 256, 0, (cudaStream_t)stream, flat_idx, n, adv, stats);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_opt_tail_error(v4l_ctx* ctx) {
  V4L_REQUIRE(ctx, "v4l_opt_tail_error: NULL ctx");
  return 0;                         // no in-kernel barrier any more: nothing can time out
}
