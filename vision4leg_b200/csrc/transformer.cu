// Per-sample attention core, LayerNorm(+residual) and token pooling of the LocoTransformer
// block (reference torchrl/networks/nets.py:949-955, 1009-1034; math: SURVEY Appendix A2).
// fp32 tier: one CTA per sample keeps q/k/v (T<=33 tokens x d<=128) in shared memory.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

// IO element type: float (exact tier) or __half (tensor-core tier); math is fp32 either way
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__half* p, float v) { *p = __float2half(v); }

constexpr int ATT_THREADS = 128;

// smem layout (floats): q[T][dp] k[T][dp] v[T][dp] (dp = d+1 to spread banks), p[nh][T][T]
template <typename T_>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const T_* __restrict__ qkv, T_* __restrict__ o, float* __restrict__ p_out,
                int T, int d, int nh) {
  v4l_pdl_enter();
  extern __shared__ float sm[];
  const int dp = d + 1, hd = d / nh;
  float* q = sm;
  float* k = q + T * dp;
  float* v = k + T * dp;
  float* p = v + T * dp;
  const int b = blockIdx.x, tid = threadIdx.x;
  const T_* src = qkv + (long long)b * T * 3 * d;
  for (int e = tid; e < T * 3 * d; e += ATT_THREADS) {
    const int t = e / (3 * d), c = e - t * 3 * d;
    const float val = ldf(src + e);
    if (c < d) q[t * dp + c] = val;
    else if (c < 2 * d) k[t * dp + c - d] = val;
    else v[t * dp + c - 2 * d] = val;
  }
  __syncthreads();
  const float scale = rsqrtf((float)hd);
  for (int e = tid; e < nh * T * T; e += ATT_THREADS) {
    const int h = e / (T * T), r = e - h * T * T, i = r / T, j = r - i * T;
    const float* qi = q + i * dp + h * hd;
    const float* kj = k + j * dp + h * hd;
    float s = 0.f;
    for (int c = 0; c < hd; ++c) s = fmaf(qi[c], kj[c], s);
    p[e] = s * scale;
  }
  __syncthreads();
  for (int r = tid; r < nh * T; r += ATT_THREADS) {
    float* row = p + r * T;
    float mx = row[0];
    for (int j = 1; j < T; ++j) mx = fmaxf(mx, row[j]);
    float sum = 0.f;
    for (int j = 0; j < T; ++j) { const float ex = expf(row[j] - mx); row[j] = ex; sum += ex; }
    const float inv = 1.f / sum;
    for (int j = 0; j < T; ++j) row[j] *= inv;
  }
  __syncthreads();
  float* pg = p_out + (long long)b * nh * T * T;
  for (int e = tid; e < nh * T * T; e += ATT_THREADS) pg[e] = p[e];
  T_* og = o + (long long)b * T * d;
  for (int e = tid; e < T * d; e += ATT_THREADS) {
    const int i = e / d, c = e - i * d, h = c / hd;
    const float* pr = p + (h * T + i) * T;
    float s = 0.f;
    for (int j = 0; j < T; ++j) s = fmaf(pr[j], v[j * dp + c], s);
    stf(og + e, s);
  }
}

// smem: q,k,v,dO [T][dp] each; p, ds [nh][T][T]
template <typename T_>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_kernel(const T_* __restrict__ qkv, const float* __restrict__ p_in,
                const T_* __restrict__ d_o, T_* __restrict__ d_qkv, int T, int d, int nh) {
  v4l_pdl_enter();
  extern __shared__ float sm[];
  const int dp = d + 1, hd = d / nh;
  float* q = sm;
  float* k = q + T * dp;
  float* v = k + T * dp;
  float* go = v + T * dp;
  float* p = go + T * dp;
  float* ds = p + nh * T * T;
  const int b = blockIdx.x, tid = threadIdx.x;
  const T_* src = qkv + (long long)b * T * 3 * d;
  for (int e = tid; e < T * 3 * d; e += ATT_THREADS) {
    const int t = e / (3 * d), c = e - t * 3 * d;
    const float val = ldf(src + e);
    if (c < d) q[t * dp + c] = val;
    else if (c < 2 * d) k[t * dp + c - d] = val;
    else v[t * dp + c - 2 * d] = val;
  }
  const T_* gsrc = d_o + (long long)b * T * d;
  for (int e = tid; e < T * d; e += ATT_THREADS) {
    const int t = e / d, c = e - t * d;
    go[t * dp + c] = ldf(gsrc + e);
  }
  const float* pg = p_in + (long long)b * nh * T * T;
  for (int e = tid; e < nh * T * T; e += ATT_THREADS) p[e] = pg[e];
  __syncthreads();
  // dP = dO V^T
  for (int e = tid; e < nh * T * T; e += ATT_THREADS) {
    const int h = e / (T * T), r = e - h * T * T, i = r / T, j = r - i * T;
    const float* gi = go + i * dp + h * hd;
    const float* vj = v + j * dp + h * hd;
    float s = 0.f;
    for (int c = 0; c < hd; ++c) s = fmaf(gi[c], vj[c], s);
    ds[e] = s;
  }
  __syncthreads();
  // dS = P * (dP - rowsum(dP * P)) * scale
  const float scale = rsqrtf((float)hd);
  for (int r = tid; r < nh * T; r += ATT_THREADS) {
    float* drow = ds + r * T;
    const float* prow = p + r * T;
    float dot = 0.f;
    for (int j = 0; j < T; ++j) dot = fmaf(drow[j], prow[j], dot);
    for (int j = 0; j < T; ++j) drow[j] = prow[j] * (drow[j] - dot) * scale;
  }
  __syncthreads();
  T_* out = d_qkv + (long long)b * T * 3 * d;
  for (int e = tid; e < T * d; e += ATT_THREADS) {
    const int t = e / d, c = e - t * d, h = c / hd;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < T; ++j) {
      dq = fmaf(ds[(h * T + t) * T + j], k[j * dp + c], dq);   // dQ[t] = sum_j dS[t,j] K[j]
      dk = fmaf(ds[(h * T + j) * T + t], q[j * dp + c], dk);   // dK[t] = sum_i dS[i,t] Q[i]
      dv = fmaf(p[(h * T + j) * T + t], go[j * dp + c], dv);   // dV[t] = sum_i P[i,t] dO[i]
    }
    stf(out + t * 3 * d + c, dq);
    stf(out + t * 3 * d + d + c, dk);
    stf(out + t * 3 * d + 2 * d + c, dv);
  }
}

// ---- LayerNorm over the last dim (d <= 256), one warp per row ---------------------------------
constexpr int LN_MAXPER = 8;

template <typename T_>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const T_* __restrict__ a, const T_* __restrict__ res,
              const float* __restrict__ gamma, const float* __restrict__ beta,
              T_* __restrict__ y, float* __restrict__ z, float* __restrict__ stats,
              int rows, int d, float eps) {
  v4l_pdl_enter();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  const long long off = (long long)row * d;
  float x[LN_MAXPER];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXPER; ++i) {
    const int c = lane + 32 * i;
    x[i] = 0.f;
    if (c < d) {
      x[i] = ldf(a + off + c) + (res ? ldf(res + off + c) : 0.f);
      s += x[i];
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / d;
  float vs = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXPER; ++i) {
    const int c = lane + 32 * i;
    if (c < d) { const float t = x[i] - mean; vs = fmaf(t, t, vs); }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, o);
  const float rstd = rsqrtf(vs / d + eps);
#pragma unroll
  for (int i = 0; i < LN_MAXPER; ++i) {
    const int c = lane + 32 * i;
    if (c < d) {
      stf(y + off + c, (x[i] - mean) * rstd * gamma[c] + beta[c]);
      if (z) z[off + c] = x[i];
    }
  }
  if (stats && lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}

// dz = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat)), dxhat = dy * gamma.
// Per-CTA partial sums of dgamma/dbeta go to part[cta][2][d] (reduced by ln_bwd_reduce_kernel).
template <typename T_>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const T_* __restrict__ dy, const float* __restrict__ z,
              const float* __restrict__ stats, const float* __restrict__ gamma,
              T_* __restrict__ dz, float* __restrict__ part, int rows, int d, int rows_per_cta) {
  v4l_pdl_enter();
  __shared__ float red[8][2][32 * LN_MAXPER];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float dg[LN_MAXPER], db[LN_MAXPER], gm[LN_MAXPER];
#pragma unroll
  for (int i = 0; i < LN_MAXPER; ++i) {
    dg[i] = 0.f; db[i] = 0.f;
    const int c = lane + 32 * i;
    gm[i] = (c < d) ? gamma[c] : 0.f;
  }
  for (int row = r0 + warp; row < r1; row += 8) {
    const long long off = (long long)row * d;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    float xh[LN_MAXPER], dxh[LN_MAXPER];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXPER; ++i) {
      const int c = lane + 32 * i;
      xh[i] = 0.f; dxh[i] = 0.f;
      if (c < d) {
        const float g = ldf(dy + off + c);
        xh[i] = (z[off + c] - mean) * rstd;
        dxh[i] = g * gm[i];
        dg[i] = fmaf(g, xh[i], dg[i]);
        db[i] += g;
        s1 += dxh[i];
        s2 = fmaf(dxh[i], xh[i], s2);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    const float m1 = s1 / d, m2 = s2 / d;
#pragma unroll
    for (int i = 0; i < LN_MAXPER; ++i) {
      const int c = lane + 32 * i;
      if (c < d) stf(dz + off + c, rstd * (dxh[i] - m1 - xh[i] * m2));
    }
  }
#pragma unroll
  for (int i = 0; i < LN_MAXPER; ++i) {
    red[warp][0][lane + 32 * i] = dg[i];
    red[warp][1][lane + 32 * i] = db[i];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * d; e += 256) {
    const int which = e / d, c = e - which * d;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][which][c];
    part[((long long)blockIdx.x * 2 + which) * d + c] = s;
  }
}

// one warp per (gamma|beta, column): lanes stride over the CTA partials
__global__ void ln_bwd_reduce_kernel(const float* __restrict__ part, int nparts, int d,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta, float out_scale) {
  v4l_pdl_enter();
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (e >= 2 * d) return;
  const int which = e / d, c = e - which * d;
  float s = 0.f;
  for (int p = lane; p < nparts; p += 32) s += part[((long long)p * 2 + which) * d + c];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) (which ? dbeta : dgamma)[c] = s * out_scale;
}

// ---- d = 64, fp16 IO fast paths: one 128-byte row per warp load (half2 per lane), grid-stride over rows
__global__ void __launch_bounds__(256)
ln_fwd_h64_kernel(const __half2* __restrict__ a, const __half2* __restrict__ res,
                  const float2* __restrict__ gamma, const float2* __restrict__ beta,
                  __half2* __restrict__ y, float2* __restrict__ z, float2* __restrict__ stats, int rows, float eps) {
  v4l_pdl_enter();
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = gridDim.x * 8;
  const float2 g = gamma[lane], b = beta[lane];
  for (int row = warp; row < rows; row += nwarps) {
    const long long o = (long long)row * 32 + lane;
    float2 x = __half22float2(a[o]);
    if (res) { const float2 r = __half22float2(res[o]); x.x += r.x; x.y += r.y; }
    float s = x.x + x.y;
#pragma unroll
    for (int k = 16; k; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
    const float mean = s * (1.f / 64.f);
    const float d0 = x.x - mean, d1 = x.y - mean;
    float vs = d0 * d0 + d1 * d1;
#pragma unroll
    for (int k = 16; k; k >>= 1) vs += __shfl_xor_sync(0xffffffffu, vs, k);
    const float rstd = rsqrtf(vs * (1.f / 64.f) + eps);
    y[o] = __floats2half2_rn(d0 * rstd * g.x + b.x, d1 * rstd * g.y + b.y);
    if (z) z[o] = x;
    if (stats && lane == 0) stats[row] = make_float2(mean, rstd);
  }
}

__global__ void __launch_bounds__(256)
ln_bwd_h64_kernel(const __half2* __restrict__ dy, const float2* __restrict__ z, const float2* __restrict__ stats,
                  const float2* __restrict__ gamma, __half2* __restrict__ dz, float* __restrict__ part, int rows) {
  v4l_pdl_enter();
  __shared__ float red[8][2][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int warp = blockIdx.x * 8 + w, nwarps = gridDim.x * 8;
  const float2 g = gamma[lane];
  float2 dg = make_float2(0.f, 0.f), db = make_float2(0.f, 0.f);
  for (int row = warp; row < rows; row += nwarps) {
    const long long o = (long long)row * 32 + lane;
    const float2 st = stats[row];
    const float2 gy = __half22float2(dy[o]);
    const float2 zz = z[o];
    const float xh0 = (zz.x - st.x) * st.y, xh1 = (zz.y - st.x) * st.y;
    const float dx0 = gy.x * g.x, dx1 = gy.y * g.y;
    dg.x = fmaf(gy.x, xh0, dg.x); dg.y = fmaf(gy.y, xh1, dg.y);
    db.x += gy.x; db.y += gy.y;
    float s1 = dx0 + dx1, s2 = dx0 * xh0 + dx1 * xh1;
#pragma unroll
    for (int k = 16; k; k >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, k);
      s2 += __shfl_xor_sync(0xffffffffu, s2, k);
    }
    const float m1 = s1 * (1.f / 64.f), m2 = s2 * (1.f / 64.f);
    dz[o] = __floats2half2_rn(st.y * (dx0 - m1 - xh0 * m2), st.y * (dx1 - m1 - xh1 * m2));
  }
  red[w][0][2 * lane] = dg.x; red[w][0][2 * lane + 1] = dg.y;
  red[w][1][2 * lane] = db.x; red[w][1][2 * lane + 1] = db.y;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][which][c];
    part[((long long)blockIdx.x * 2 + which) * 64 + c] = s;
  }
}

template <typename T_>
__global__ void pool_fwd_kernel(const T_* __restrict__ tok, T_* __restrict__ out, int B, int T,
                                int d, int mode) {
  v4l_pdl_enter();
  const int od = mode == 0 ? 2 * d : d;
  const long long total = (long long)B * od;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(e / od), c = (int)(e - (long long)b * od);
    const T_* t = tok + (long long)b * T * d;
    float v;
    if (mode == 0 && c < d) {
      v = ldf(t + c);
    } else {
      const int cc = mode == 0 ? c - d : c;
      const int t0 = mode == 0 ? 1 : 0;
      float s = 0.f;
      for (int i = t0; i < T; ++i) s += ldf(t + i * d + cc);
      v = s / (float)(T - t0);
    }
    stf(out + e, v);
  }
}

template <typename T_>
__global__ void pool_bwd_kernel(const T_* __restrict__ dout, T_* __restrict__ dtok, int B, int T,
                                int d, int mode) {
  v4l_pdl_enter();
  const int od = mode == 0 ? 2 * d : d;
  const long long total = (long long)B * T * d;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % d);
    const long long r = e / d;
    const int t = (int)(r % T), b = (int)(r / T);
    const T_* g = dout + (long long)b * od;
    float v;
    if (mode == 0) v = (t == 0) ? ldf(g + c) : ldf(g + d + c) / (float)(T - 1);
    else v = ldf(g + c) / (float)T;
    stf(dtok + e, v);
  }
}

}  // namespace

static int attn_check(const char* who, int B, int T, int d, int nh) {
  V4L_REQUIRE(B >= 0 && T > 0 && T <= 64 && d > 0 && d <= 256 && nh > 0 && d % nh == 0,
              "%s: unsupported shape B=%d T=%d d=%d n_head=%d", who, B, T, d, nh);
  return 0;
}

typedef __half f16_t;

template <typename T_>
static int attn_fwd_impl(v4l_ctx* ctx, void* stream, const void* qkv, void* o, float* p, int B, int T, int d,
                         int n_head) {
  V4L_REQUIRE(ctx && qkv && o && p, "v4l_attn_fwd: NULL argument");
  if (int r = attn_check("v4l_attn_fwd", B, T, d, n_head)) return r;
  if (B == 0) return 0;
  const size_t smem = sizeof(float) * (3 * T * (d + 1) + n_head * T * T);
  if (smem > 48 * 1024)
    V4L_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  V4L_LAUNCH((attn_fwd_kernel<T_>), B, ATT_THREADS, smem, (cudaStream_t)stream, (const T_*)qkv, (T_*)o, p, T, d, n_head);
  V4L_CHECK_LAUNCH();
  return 0;
}

template <typename T_>
static int attn_bwd_impl(v4l_ctx* ctx, void* stream, const void* qkv, const float* p, const void* d_o,
                         void* d_qkv, int B, int T, int d, int n_head) {
  V4L_REQUIRE(ctx && qkv && p && d_o && d_qkv, "v4l_attn_bwd: NULL argument");
  if (int r = attn_check("v4l_attn_bwd", B, T, d, n_head)) return r;
  if (B == 0) return 0;
  const size_t smem = sizeof(float) * (4 * T * (d + 1) + 2 * n_head * T * T);
  if (smem > 48 * 1024)
    V4L_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<T_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  V4L_LAUNCH((attn_bwd_kernel<T_>), B, ATT_THREADS, smem, (cudaStream_t)stream, (const T_*)qkv, p, (const T_*)d_o, (T_*)d_qkv, T, d, n_head);
  V4L_CHECK_LAUNCH();
  return 0;
}

template <typename T_>
static int ln_fwd_impl(v4l_ctx* ctx, void* stream, const void* a, const void* res, const float* gamma,
                       const float* beta, void* y, float* z, float* stats, int rows, int d, float eps) {
  V4L_REQUIRE(ctx && a && gamma && beta && y, "v4l_ln_fwd: NULL argument");
  V4L_REQUIRE(d > 0 && d <= 32 * LN_MAXPER, "v4l_ln_fwd: d=%d unsupported (max %d)", d, 32 * LN_MAXPER);
  if (rows == 0) return 0;
  if (sizeof(T_) == 2 && d == 64) {
    const int ctas = min(v4l_cdiv(rows, 8), 8 * ctx->sm_count);
    V4L_LAUNCH(ln_fwd_h64_kernel, ctas, 256, 0, (cudaStream_t)stream, (const __half2*)a, (const __half2*)res,
               (const float2*)gamma, (const float2*)beta, (__half2*)y, (float2*)z, (float2*)stats, rows, eps);
    return 0;
  }
  V4L_LAUNCH((ln_fwd_kernel<T_>), v4l_cdiv(rows, 8), 256, 0, (cudaStream_t)stream, (const T_*)a, (const T_*)res, gamma, beta,
                                                                          (T_*)y, z, stats, rows, d, eps);
  V4L_CHECK_LAUNCH();
  return 0;
}

template <typename T_>
static int ln_bwd_impl(v4l_ctx* ctx, void* stream, const void* dy, const float* z, const float* stats,
                       const float* gamma, void* dz, float* dgamma, float* dbeta, int rows, int d,
                       float out_scale) {
  V4L_REQUIRE(ctx && dy && z && stats && gamma && dz && dgamma && dbeta, "v4l_ln_bwd: NULL argument");
  V4L_REQUIRE(d > 0 && d <= 32 * LN_MAXPER, "v4l_ln_bwd: d=%d unsupported", d);
  V4L_REQUIRE(rows > 0, "v4l_ln_bwd: rows must be > 0");
  int ctas = min(2 * ctx->sm_count, v4l_cdiv(rows, 8));
  const int rpc = ((v4l_cdiv(rows, ctas) + 7) / 8) * 8;
  ctas = v4l_cdiv(rows, rpc);
  V4L_REQUIRE((size_t)ctas * 2 * d <= ctx->scratch_elems, "v4l_ln_bwd: scratch too small");
  cudaStream_t s = (cudaStream_t)stream;
  if (sizeof(T_) == 2 && d == 64) {
    ctas = min(ctas, ctx->sm_count);
    V4L_LAUNCH(ln_bwd_h64_kernel, ctas, 256, 0, s, (const __half2*)dy, (const float2*)z, (const float2*)stats,
               (const float2*)gamma, (__half2*)dz, ctx->scratch, rows);
  } else {
    V4L_LAUNCH((ln_bwd_kernel<T_>), ctas, 256, 0, s, (const T_*)dy, z, stats, gamma, (T_*)dz, ctx->scratch, rows, d, rpc);
  }
  V4L_CHECK_LAUNCH();
  V4L_LAUNCH(ln_bwd_reduce_kernel, v4l_cdiv(2 * d, 8), 256, 0, s, ctx->scratch, ctas, d, dgamma, dbeta, out_scale);
  V4L_CHECK_LAUNCH();
  return 0;
}

template <typename T_>
static int pool_impl(v4l_ctx* ctx, void* stream, const void* in, void* out, int B, int T, int d, int mode, bool fwd) {
  V4L_REQUIRE(ctx && in && out, "v4l_pool: NULL argument");
  V4L_REQUIRE((mode == 0 && T >= 2) || (mode == 1 && T >= 1), "v4l_pool: bad mode/T");
  const long long total = fwd ? (long long)B * (mode == 0 ? 2 * d : d) : (long long)B * T * d;
  if (total == 0) return 0;
  const int blocks = (int)min((long long)8 * ctx->sm_count, (total + 255) / 256);
  if (fwd) V4L_LAUNCH((pool_fwd_kernel<T_>), blocks, 256, 0, (cudaStream_t)stream, (const T_*)in, (T_*)out, B, T, d, mode);
  else     V4L_LAUNCH((pool_bwd_kernel<T_>), blocks, 256, 0, (cudaStream_t)stream, (const T_*)in, (T_*)out, B, T, d, mode);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_attn_fwd(v4l_ctx* ctx, void* stream, const float* qkv, float* o, float* p,
                            int B, int T, int d, int n_head) {
  return attn_fwd_impl<float>(ctx, stream, qkv, o, p, B, T, d, n_head);
}
extern "C" int v4l_attn_bwd(v4l_ctx* ctx, void* stream, const float* qkv, const float* p,
                            const float* d_o, float* d_qkv, int B, int T, int d, int n_head) {
  return attn_bwd_impl<float>(ctx, stream, qkv, p, d_o, d_qkv, B, T, d, n_head);
}
extern "C" int v4l_ln_fwd(v4l_ctx* ctx, void* stream, const float* a, const float* res,
                          const float* gamma, const float* beta, float* y, float* z, float* stats,
                          int rows, int d, float eps) {
  return ln_fwd_impl<float>(ctx, stream, a, res, gamma, beta, y, z, stats, rows, d, eps);
}
extern "C" int v4l_ln_bwd(v4l_ctx* ctx, void* stream, const float* dy, const float* z,
                          const float* stats, const float* gamma, float* dz, float* dgamma,
                          float* dbeta, int rows, int d) {
  return ln_bwd_impl<float>(ctx, stream, dy, z, stats, gamma, dz, dgamma, dbeta, rows, d, 1.f);
}
extern "C" int v4l_pool_fwd(v4l_ctx* ctx, void* stream, const float* tok, float* out, int B, int T,
                            int d, int mode) {
  return pool_impl<float>(ctx, stream, tok, out, B, T, d, mode, true);
}
extern "C" int v4l_pool_bwd(v4l_ctx* ctx, void* stream, const float* dout, float* dtok, int B, int T,
                            int d, int mode) {
  return pool_impl<float>(ctx, stream, dout, dtok, B, T, d, mode, false);
}

// ---- f16-IO variants used by the tensor-core tier (fp32 math, fp32 softmax/LayerNorm statistics)
extern "C" int v4l_attn_fwd_f16(v4l_ctx* ctx, void* stream, const void* qkv, void* o, float* p,
                                 int B, int T, int d, int n_head) {
  return attn_fwd_impl<f16_t>(ctx, stream, qkv, o, p, B, T, d, n_head);
}
extern "C" int v4l_attn_bwd_f16(v4l_ctx* ctx, void* stream, const void* qkv, const float* p,
                                 const void* d_o, void* d_qkv, int B, int T, int d, int n_head) {
  return attn_bwd_impl<f16_t>(ctx, stream, qkv, p, d_o, d_qkv, B, T, d, n_head);
}
extern "C" int v4l_ln_fwd_f16(v4l_ctx* ctx, void* stream, const void* a, const void* res,
                               const float* gamma, const float* beta, void* y, float* z, float* stats,
                               int rows, int d, float eps) {
  return ln_fwd_impl<f16_t>(ctx, stream, a, res, gamma, beta, y, z, stats, rows, d, eps);
}
extern "C" int v4l_ln_bwd_f16(v4l_ctx* ctx, void* stream, const void* dy, const float* z,
                               const float* stats, const float* gamma, void* dz, float* dgamma,
                               float* dbeta, int rows, int d, float out_scale) {
  return ln_bwd_impl<f16_t>(ctx, stream, dy, z, stats, gamma, dz, dgamma, dbeta, rows, d, out_scale);
}
extern "C" int v4l_pool_fwd_f16(v4l_ctx* ctx, void* stream, const void* tok, void* out, int B, int T,
                                 int d, int mode) {
  return pool_impl<f16_t>(ctx, stream, tok, out, B, T, d, mode, true);
}
extern "C" int v4l_pool_bwd_f16(v4l_ctx* ctx, void* stream, const void* dout, void* dtok, int B, int T,
                                 int d, int mode) {
  return pool_impl<f16_t>(ctx, stream, dout, dtok, B, T, d, mode, false);
}
