// FP32 CUDA-core GEMMs with gathered operands: the exact-fp32 tier of the PPO update path.
//
//   v4l_gemm_rows   C = act(A_gather * B + bias) [* relu-mask] — forward of every Linear/Conv
//                   layer (im2col on the fly through the row map) and their data-gradients
//   v4l_gemm_wgrad  dW = dY^T * A_gather, dbias = colsum(dY) — deterministic split-M, the bias
//                   gradient rides as a virtual all-ones column K of A
//   v4l_col2im      gather-form col2im for the conv data-gradients
//
// Tensor cores have no true-fp32 mode; this tier is what meets the 1e-3 fp32 parity bar
// (DESIGN.md §Precision tiers).  Tiles: 64 x BN x 16, 256 threads, 4 x (BN/16) outputs/thread.
#include "common.cuh"

namespace {

constexpr int BM = 64;
constexpr int BK = 16;
constexpr int NT = 256;

template <int BN, bool B_NCONTIG>
__global__ void __launch_bounds__(NT) gemm_rows_kernel(const v4l_gemm_args g) {
  v4l_pdl_enter();
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int M = g.M, N = g.N, K = g.K;

  // A loader: thread owns column a_kk of 4 rows (a_r + 16 i)
  const int a_kk = tid & 15, a_r = tid >> 4;
  long long a_row[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + a_r + 16 * i;
    a_row[i] = (m < M) ? v4l_row_addr(g.a_map, m) : -1;
  }
  // B loader
  int b_kk, b_nn, b_kstep, b_nstep;
  if (B_NCONTIG) { b_nn = tid % BN; b_kk = tid / BN; b_kstep = NT / BN; b_nstep = 0; }
  else           { b_kk = tid & 15; b_nn = tid >> 4; b_kstep = 0; b_nstep = 16; }

  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
    {
      const int k = k0 + a_kk;
      const bool kv = k < K;
      const long long koff = kv ? (g.a_koff ? (long long)g.a_koff[k] : (long long)k) : 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        As[a_kk][a_r + 16 * i] = (kv && a_row[i] >= 0) ? g.a[a_row[i] + koff] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < BN / 16; ++j) {
      const int kk = b_kk + j * b_kstep, nn = b_nn + j * b_nstep;
      const int k = k0 + kk, n = n0 + nn;
      Bs[kk][nn] = (k < K && n < N) ? g.b[(long long)k * g.b_sk + (long long)n * g.b_sn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      float b4[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b4[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }

  const bool relu = g.flags & V4L_RELU, accum = g.flags & V4L_ACCUM;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const long long crow = v4l_row_addr(g.c_map, m);
    const long long mrow = g.mask ? v4l_row_addr(g.mask_map, m) : 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= N) continue;
      const long long cn = g.c_koff ? (long long)g.c_koff[n] : (long long)n;
      float v = acc[i][j];
      if (g.bias) v += g.bias[n];
      if (relu) v = fmaxf(v, 0.f);
      if (g.mask) v = (g.mask[mrow + cn] > 0.f) ? v : 0.f;
      if (accum) v += g.c[crow + cn];
      g.c[crow + cn] = v;
    }
  }
}

// ---- weight gradient: partial[z][n][kext] = sum_{m in split z} dY(m,n) * Aext(m,kext) ----------
__global__ void __launch_bounds__(NT) wgrad_kernel(const v4l_wgrad_args g, float* __restrict__ partial,
                                                   int Kext, int rows_per_split) {
  v4l_pdl_enter();
  __shared__ __align__(16) float Ys[BK][64 + 4];
  __shared__ __align__(16) float As[BK][64 + 4];
  __shared__ long long rowY[BK], rowA[BK];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int M = g.M, N = g.N, K = g.K;
  const int mbeg = blockIdx.z * rows_per_split;
  const int mend = min(M, mbeg + rows_per_split);

  const int l_c = tid & 63, l_r = tid >> 6;       // loader: column l_c, rows l_r + 4 j
  const int kcol = k0 + l_c;
  const bool k_real = kcol < K;
  const bool k_ones = (kcol == K) && (Kext > K);
  const long long koff = k_real ? (g.a_koff ? (long long)g.a_koff[kcol] : (long long)kcol) : 0;
  const bool n_ok = (n0 + l_c) < N;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int m0 = mbeg; m0 < mend; m0 += BK) {
    if (tid < BK) {
      const int m = m0 + tid;
      rowY[tid] = (m < mend) ? v4l_row_addr(g.dy_map, m) : -1;
    } else if (tid < 2 * BK) {
      const int m = m0 + tid - BK;
      rowA[tid - BK] = (m < mend) ? v4l_row_addr(g.a_map, m) : -1;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int mm = l_r + 4 * j;
      const long long ry = rowY[mm], ra = rowA[mm];
      Ys[mm][l_c] = (ry >= 0 && n_ok) ? g.dy[ry + n0 + l_c] : 0.f;
      float av = 0.f;
      if (ra >= 0) av = k_real ? g.a[ra + koff] : (k_ones ? 1.f : 0.f);
      As[mm][l_c] = av;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < BK; ++mm) {
      const float4 yv = *reinterpret_cast<const float4*>(&Ys[mm][ty * 4]);
      const float4 av = *reinterpret_cast<const float4*>(&As[mm][tx * 4]);
      const float y4[4] = {yv.x, yv.y, yv.z, yv.w};
      const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(y4[i], a4[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* out = partial + (long long)blockIdx.z * N * Kext;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < Kext) out[(long long)n * Kext + k] = acc[i][j];
    }
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int N, int K,
                                    int Kext, float* __restrict__ dw, long long ldw,
                                    float* __restrict__ dbias) {
  v4l_pdl_enter();
  const long long total = (long long)N * Kext;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += partial[(long long)z * total + e];
    const int n = (int)(e / Kext), k = (int)(e - (long long)n * Kext);
    if (k < K) dw[(long long)n * ldw + k] = s;
    else dbias[n] = s;
  }
}

__global__ void col2im_kernel(const float* __restrict__ dcol, const float* __restrict__ x,
                              float* __restrict__ dx, int B, int Hin, int Win, int C, int KH, int KW,
                              int stride, int Hout, int Wout) {
  v4l_pdl_enter();
  const long long total = (long long)B * Hin * Win * C;
  const int K = C * KH * KW;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    long long r = e / C;
    const int w = (int)(r % Win); r /= Win;
    const int h = (int)(r % Hin);
    const int b = (int)(r / Hin);
    float s = 0.f;
    if (x == nullptr || x[e] > 0.f) {
      for (int kh = 0; kh < KH; ++kh) {
        const int hh = h - kh;
        if (hh < 0 || hh % stride) continue;
        const int oh = hh / stride;
        if (oh >= Hout) continue;
        for (int kw = 0; kw < KW; ++kw) {
          const int ww = w - kw;
          if (ww < 0 || ww % stride) continue;
          const int ow = ww / stride;
          if (ow >= Wout) continue;
          s += dcol[((long long)(b * Hout + oh) * Wout + ow) * K + (c * KH + kh) * KW + kw];
        }
      }
    }
    dx[e] = s;
  }
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, const v4l_rowmap dy_map,
                                const float* __restrict__ act, const v4l_rowmap act_map,
                                float* __restrict__ out, const v4l_rowmap out_map, int M, int N) {
  v4l_pdl_enter();
  const long long total = (long long)M * N;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
       e += (long long)gridDim.x * blockDim.x) {
    const int m = (int)(e / N), n = (int)(e - (long long)m * N);
    const float g = dy[v4l_row_addr(dy_map, m) + n];
    const float a = act[v4l_row_addr(act_map, m) + n];
    out[v4l_row_addr(out_map, m) + n] = a > 0.f ? g : 0.f;
  }
}

}  // namespace

extern "C" int v4l_relu_bwd(v4l_ctx* ctx, void* stream, const float* dy, const v4l_rowmap* dy_map,
                            const float* act, const v4l_rowmap* act_map, float* out,
                            const v4l_rowmap* out_map, int M, int N) {
  V4L_REQUIRE(ctx && dy && dy_map && act && act_map && out && out_map, "v4l_relu_bwd: NULL argument");
  V4L_REQUIRE(dy_map->P > 0 && act_map->P > 0 && out_map->P > 0, "v4l_relu_bwd: row map with P <= 0");
  const long long total = (long long)M * N;
  if (total <= 0) return 0;
  const int blocks = (int)min((long long)8 * ctx->sm_count, (total + 255) / 256);
  V4L_LAUNCH(relu_bwd_kernel, blocks, 256, 0, (cudaStream_t)stream, dy, *dy_map, act, *act_map, out, *out_map, M, N);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_gemm_rows(v4l_ctx* ctx, void* stream, const v4l_gemm_args* a) {
  V4L_REQUIRE(ctx && a, "v4l_gemm_rows: NULL argument");
  V4L_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, "v4l_gemm_rows: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  V4L_REQUIRE(a->a && a->b && a->c, "v4l_gemm_rows: NULL operand");
  V4L_REQUIRE(a->a_map.P > 0 && a->c_map.P > 0 && (!a->mask || a->mask_map.P > 0),
              "v4l_gemm_rows: row map with P <= 0");
  if (a->M == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const bool ncontig = (a->b_sn == 1 && a->b_sk != 1);
  const int bn = (a->N <= 32) ? 32 : 64;
  dim3 grid(v4l_cdiv(a->M, BM), v4l_cdiv(a->N, bn));
  V4L_REQUIRE(grid.y <= 65535, "v4l_gemm_rows: N too large");
  if (bn == 32) {
    if (ncontig) V4L_LAUNCH((gemm_rows_kernel<32, true>), grid, NT, 0, s, *a);
    else         V4L_LAUNCH((gemm_rows_kernel<32, false>), grid, NT, 0, s, *a);
  } else {
    if (ncontig) V4L_LAUNCH((gemm_rows_kernel<64, true>), grid, NT, 0, s, *a);
    else         V4L_LAUNCH((gemm_rows_kernel<64, false>), grid, NT, 0, s, *a);
  }
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_gemm_wgrad(v4l_ctx* ctx, void* stream, const v4l_wgrad_args* a) {
  V4L_REQUIRE(ctx && a, "v4l_gemm_wgrad: NULL argument");
  V4L_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, "v4l_gemm_wgrad: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  V4L_REQUIRE(a->dy && a->a && a->dw, "v4l_gemm_wgrad: NULL operand");
  V4L_REQUIRE(a->dy_map.P > 0 && a->a_map.P > 0, "v4l_gemm_wgrad: row map with P <= 0");
  cudaStream_t s = (cudaStream_t)stream;
  const int Kext = a->K + (a->dbias ? 1 : 0);
  const int gx = v4l_cdiv(Kext, 64), gy = v4l_cdiv(a->N, 64);
  const long long per_split = (long long)a->N * Kext;
  int splits = v4l_cdiv(4LL * ctx->sm_count, (long long)gx * gy);
  splits = max(1, min(splits, v4l_cdiv(a->M, 64)));
  splits = (int)min((long long)splits, (long long)(ctx->scratch_elems / per_split));
  V4L_REQUIRE(splits >= 1, "v4l_gemm_wgrad: scratch too small for N=%d K=%d", a->N, a->K);
  int rps = v4l_cdiv(a->M, splits);
  rps = ((rps + BK - 1) / BK) * BK;
  splits = v4l_cdiv(a->M, rps);
  V4L_REQUIRE(splits <= 65535, "v4l_gemm_wgrad: too many splits");
  dim3 grid(gx, gy, splits);
  V4L_LAUNCH(wgrad_kernel, grid, NT, 0, s, *a, ctx->scratch, Kext, rps);
  V4L_CHECK_LAUNCH();
  const int rblocks = (int)min((long long)4 * ctx->sm_count, (per_split + 255) / 256);
  V4L_LAUNCH(wgrad_reduce_kernel, rblocks, 256, 0, s, ctx->scratch, splits, a->N, a->K, Kext, a->dw, a->ldw, a->dbias);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_col2im(v4l_ctx* ctx, void* stream, const float* dcol, const float* x, float* dx,
                          int B, int Hin, int Win, int C, int KH, int KW, int stride, int Hout, int Wout) {
  V4L_REQUIRE(ctx && dcol && dx, "v4l_col2im: NULL argument");
  V4L_REQUIRE((Hin - KH) / stride + 1 == Hout && (Win - KW) / stride + 1 == Wout,
              "v4l_col2im: inconsistent geometry");
  const long long total = (long long)B * Hin * Win * C;
  if (total == 0) return 0;
  const int blocks = (int)min((long long)16 * ctx->sm_count, (total + 255) / 256);
  V4L_LAUNCH(col2im_kernel, blocks, 256, 0, (cudaStream_t)stream, dcol, x, dx, B, Hin, Win, C, KH, KW, stride, Hout, Wout);
  V4L_CHECK_LAUNCH();
  return 0;
}
