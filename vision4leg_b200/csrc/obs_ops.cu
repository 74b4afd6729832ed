// Observation pipeline on the device — the step on the far side of the collector:
//   * depth frames: OpenGL depth-buffer value z -> metric depth far*near / (far - (far-near) z) -> clip [0.3, 10]
//     -> sqrt(log(d + 1)) [-> (x - 1.25) / 0.425]  (reference vision4leg/envs/
//     locomotion_gym_env_with_rich_information.py:620-633,649-650), kept in a per-env ring of processed frames;
//   * k-frame stacking: the observation's 4 channels are ring slots head - frame_idx[k] (the reference's deque
//     indices, :315-336,549-554,641-648), written straight as the fp16 4x4 space-to-depth image the tensor-core
//     tier consumes (and, optionally, as the fp32 CHW row of the reference's observation vector);
//   * running-mean proprio normaliser: batch mean / variance merged into (mean, var, count) the way the reference
//     does (torchrl/env/base_wrapper.py:44-61), then clip((x - mean) / (sqrt(var) + 1e-4), +-clip) (:88-90,119-122).
// Elementwise / small-reduction HBM-bound work: coalesced loads, one pass.
#include <cuda_fp16.h>
#include <math.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ float depth_feature(float z, float farp, float far_near, float far_minus_near) {
  // numpy evaluates this as separately rounded fp32 operations and the expression cancels badly near z = 1 (the far
  // field): no fused multiply-add here, or the result differs from the reference's by 1e-5
  float d = __fdiv_rn(far_near, __fsub_rn(farp, __fmul_rn(far_minus_near, z)));
  d = fminf(fmaxf(d, 0.3f), 10.f);
  return sqrtf(logf(d + 1.f));
}

// one thread per 4 pixels of one env's new frame (16-byte loads and stores)
__global__ void depth_frame_kernel(const float* __restrict__ z, float* __restrict__ ring, const uint8_t* __restrict__ reset,
                                   int E, int n_slots, int head, float farp, float far_near, float far_minus_near) {
  v4l_pdl_enter();
  const long long total = (long long)E * 1024;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int e = (int)(i >> 10), q = (int)(i & 1023);
    const float4 zz = __ldcs(reinterpret_cast<const float4*>(z) + i);
    const float4 f = make_float4(depth_feature(zz.x, farp, far_near, far_minus_near), depth_feature(zz.y, farp, far_near, far_minus_near),
                                 depth_feature(zz.z, farp, far_near, far_minus_near), depth_feature(zz.w, farp, far_near, far_minus_near));
    float4* mine = reinterpret_cast<float4*>(ring + (long long)e * n_slots * 4096) + q;
    if (reset && reset[e]) {                         // an episode start fills the whole history (reference :635-637)
      for (int s = 0; s < n_slots; ++s) mine[(long long)s * 1024] = f;
    } else {
      mine[(long long)head * 1024] = f;
    }
  }
}

// thread per (env, Y, X, py), py fastest: 4 px x 4 channels = 16 fp16 = 32 bytes of the s2d image, so four
// neighbouring threads fill one 128-byte line of it
__global__ void stack_frames_kernel(const float* __restrict__ ring, const int32_t* __restrict__ slots, int E, int n_slots,
                                    int normalise, __half* __restrict__ s2d, float* __restrict__ chw, long long chw_stride,
                                    int chw_vec) {
  v4l_pdl_enter();
  const long long total = (long long)E * 16 * 16 * 4;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int py = (int)(t & 3), X = (int)((t >> 2) & 15), Y = (int)((t >> 6) & 15), e = (int)(t >> 10);
    const int row = 4 * Y + py, col = 4 * X;
    float v[4][4];                                   // [channel][px]
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int slot = slots[e * 4 + c];
      const float4 f = *reinterpret_cast<const float4*>(ring + ((long long)e * n_slots + slot) * 4096 + row * 64 + col);
      v[c][0] = f.x; v[c][1] = f.y; v[c][2] = f.z; v[c][3] = f.w;
      if (normalise) {
#pragma unroll
        for (int p = 0; p < 4; ++p) v[c][p] = (v[c][p] - 1.25f) / 0.425f;
      }
      if (chw) {
        float* dst = chw + (long long)e * chw_stride + c * 4096 + row * 64 + col;
        if (chw_vec) {
          *reinterpret_cast<float4*>(dst) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        } else {                                     // the image part of an observation row starts at column S: any alignment
          dst[0] = v[c][0]; dst[1] = v[c][1]; dst[2] = v[c][2]; dst[3] = v[c][3];
        }
      }
    }
    if (s2d) {
      uint32_t w[8];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        __half2 a = __floats2half2_rn(v[0][p], v[1][p]), b = __floats2half2_rn(v[2][p], v[3][p]);
        w[2 * p] = *reinterpret_cast<uint32_t*>(&a); w[2 * p + 1] = *reinterpret_cast<uint32_t*>(&b);
      }
      uint4* dst = reinterpret_cast<uint4*>(s2d + (((long long)e * 16 + Y) * 16 + X) * 64 + py * 16);
      dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
      dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
  }
}

// thread per column: statistics of x [n, S] in double, merge into the running (mean, var) with the host-tracked
// count (it is 1e-4 + the rows seen so far: deterministic, so it travels by value), then filter the n rows
__global__ void __launch_bounds__(256) normalizer_kernel(const float* __restrict__ x, int n, int S, double* __restrict__ mean,
                                                         double* __restrict__ var, double count, int update, float clip,
                                                         float* __restrict__ out) {
  v4l_pdl_enter();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= S) return;
  double m = mean[c], v = var[c];
  if (update == 2) {                                         // batch statistics only (the data-parallel merge is the caller's)
    double bs = 0.0;
    for (int i = 0; i < n; ++i) bs += x[(long long)i * S + c];
    const double bm = n > 0 ? bs / n : 0.0;
    double bv = 0.0;
    for (int i = 0; i < n; ++i) { const double d = x[(long long)i * S + c] - bm; bv += d * d; }
    mean[c] = bm; var[c] = n > 0 ? bv / n : 0.0;
    return;
  }
  if (update && n > 0) {
    double bs = 0.0;
    for (int i = 0; i < n; ++i) bs += x[(long long)i * S + c];
    const double bm = bs / n;
    double bv = 0.0;
    for (int i = 0; i < n; ++i) { const double d = x[(long long)i * S + c] - bm; bv += d * d; }
    bv /= n;                                                 // np.var: population variance
    const double tot = count + n, delta = bm - m;
    const double M2 = v * count + bv * n + delta * delta * count * n / tot;
    m = m + delta * n / tot;
    v = M2 / tot;
    mean[c] = m; var[c] = v;
  }
  if (out) {
    const double sd = sqrt(v) + 1e-4;
    for (int i = 0; i < n; ++i) {
      const double f = ((double)x[(long long)i * S + c] - m) / sd;
      out[(long long)i * S + c] = (float)fmin(fmax(f, -(double)clip), (double)clip);
    }
  }
}

}  // namespace

extern "C" int v4l_depth_frame(v4l_ctx* ctx, void* stream, const float* zbuf, float* ring, const uint8_t* reset, int E,
                               int n_slots, int head, float near_plane, float far_plane) {
  V4L_REQUIRE(ctx && zbuf && ring && E > 0 && n_slots > 0 && head >= 0 && head < n_slots, "v4l_depth_frame: bad argument");
  V4L_REQUIRE((((uintptr_t)zbuf | (uintptr_t)ring) & 15) == 0, "v4l_depth_frame: zbuf and ring must be 16-byte aligned");
  const long long total = (long long)E * 1024;
  V4L_LAUNCH(depth_frame_kernel, (int)min((long long)8 * ctx->sm_count, (total + 255) / 256), 256, 0, (cudaStream_t)stream,
             zbuf, ring, reset, E, n_slots, head, far_plane, (float)((double)far_plane * (double)near_plane),
             (float)((double)far_plane - (double)near_plane));   // Python-float scalars, rounded once (NumPy weak scalars)
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_stack_frames(v4l_ctx* ctx, void* stream, const float* ring, const int32_t* slots, int E, int n_slots,
                                int normalise, void* out_s2d, float* out_chw, int64_t chw_stride) {
  V4L_REQUIRE(ctx && ring && slots && E > 0 && n_slots > 0 && (out_s2d || out_chw), "v4l_stack_frames: bad argument");
  V4L_REQUIRE((((uintptr_t)ring | (uintptr_t)out_s2d) & 15) == 0, "v4l_stack_frames: ring and out_s2d must be 16-byte aligned");
  V4L_REQUIRE(!out_chw || chw_stride >= 16384, "v4l_stack_frames: the fp32 CHW row stride must be at least 16384");
  const int chw_vec = out_chw && chw_stride % 4 == 0 && ((uintptr_t)out_chw & 15) == 0;
  const long long total = (long long)E * 1024;
  V4L_LAUNCH(stack_frames_kernel, (int)min((long long)8 * ctx->sm_count, (total + 255) / 256), 256, 0, (cudaStream_t)stream,
             ring, slots, E, n_slots, normalise, reinterpret_cast<__half*>(out_s2d), out_chw, (long long)chw_stride, chw_vec);
  V4L_CHECK_LAUNCH();
  return 0;
}

extern "C" int v4l_normalizer(v4l_ctx* ctx, void* stream, const float* x, int n, int S, double* mean, double* var,
                              double count, int update, float clip, float* out) {
  V4L_REQUIRE(ctx && x && mean && var && n >= 0 && S > 0 && count > 0.0 && update >= 0 && update <= 2,
              "v4l_normalizer: bad argument");
  V4L_REQUIRE(update != 2 || out == nullptr, "v4l_normalizer: update = 2 (batch statistics only) does not filter");
  V4L_LAUNCH(normalizer_kernel, (S + 255) / 256, 256, 0, (cudaStream_t)stream, x, n, S, mean, var, count, update, clip, out);
  V4L_CHECK_LAUNCH();
  return 0;
}
