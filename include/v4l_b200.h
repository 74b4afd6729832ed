/*
 * v4l_b200.h — C ABI of libv4l_b200.so: the B200 (sm_100a) PPO-update hot path of
 * Mehooz/vision4leg.
 *
 * The reference has no FFI for this path: its boundary is a set of Python classes
 * (SURVEY.md §8(b)).  The host-side mirror of those classes lives in vision4leg_b200/ (Python)
 * and calls ONLY the entry points declared here.  Each entry point cites the reference code
 * whose arithmetic it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; v4l_last_error() gives the message;
 *   - all data pointers are CALLER-OWNED DEVICE pointers (tensor.data_ptr()) unless a name
 *     starts with `h_` (host, pinned) — no torch types cross this boundary;
 *   - `stream` is a cudaStream_t passed as void* (the caller's current stream);
 *   - no allocation of user-visible memory; the only library-owned memory is the per-context
 *     scratch (split-K partials), so a context must be used from one stream at a time;
 *   - no exceptions, no global state besides the last-error string.
 */
#ifndef V4L_B200_H_
#define V4L_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V4L_ABI_VERSION 1

typedef struct v4l_ctx v4l_ctx;

/* ---- library / context ------------------------------------------------------------------- */
int         v4l_version(void);
const char* v4l_last_error(void);
/* scratch_bytes: size of the context-owned scratch buffer (0 = default 1 GiB: half for immediate users, half for deferred
 * weight-gradient partials). */
int         v4l_ctx_create(v4l_ctx** out, int device, size_t scratch_bytes);
int         v4l_ctx_destroy(v4l_ctx* ctx);
int         v4l_ctx_sm_count(const v4l_ctx* ctx);
/* How many times a deferred weight-gradient job found the scratch too full for its split count and reduced the
 * pending jobs early, ON ITS OWN STREAM.  That is only ordered after the other pending jobs when they were all
 * launched on that stream: a caller that spreads weight-gradient launches over several streams must size the
 * scratch so that this stays constant (the default, 1 GiB, does for every shipped network at any minibatch). */
int         v4l_ctx_early_flushes(const v4l_ctx* ctx);

/* ---- addressing ---------------------------------------------------------------------------
 * A "row map" addresses logical row m of a matrix that lives inside a larger activation
 * tensor:   item = m / P,  pos = m % P,  item' = idx ? idx[item] : item
 *           addr(m) = base + item' * item_stride + (pos_off ? pos_off[pos] : pos * pos_stride)
 * (all in elements).  It expresses on-the-fly im2col (pos_off = receptive-field origin of an
 * output pixel), minibatch row gathers (idx) and token-slot writes without copies.          */
typedef struct {
  int32_t        P;
  int64_t        item_stride;
  int64_t        pos_stride;
  int64_t        base;
  const int32_t* idx;      /* [M / P] or NULL */
  const int32_t* pos_off;  /* [P] or NULL */
} v4l_rowmap;

enum { V4L_RELU = 1, V4L_ACCUM = 2 };

/* C(m,n) (+)= act( sum_k A(m,k) * B(k,n) + bias[n] ) * (mask(m,n) > 0)
 *   A(m,k) = a[ addr_a(m) + (a_koff ? a_koff[k] : k) ]      (gathered operand)
 *   B(k,n) = b[ k * b_sk + n * b_sn ]                       (weights, either orientation)
 *   C(m,n) = c[ addr_c(m) + cn ],  mask(m,n) = mask[ addr_mask(m) + cn ],
 *            cn = c_koff ? c_koff[n] : n   (column scatter, e.g. (c,p) -> (p,c) of a flatten)
 * Forward of every Linear / Conv2d on the path (reference torchrl/networks/base.py:8-44,
 * 209-230, 317-324, 531, nets.py:973-992, torch.nn.MultiheadAttention projections) and, with
 * B transposed, their data-gradients.                                                       */
typedef struct {
  const float* a;  v4l_rowmap a_map;  const int32_t* a_koff;
  const float* b;  int64_t b_sk;  int64_t b_sn;
  const float* bias;
  float*       c;  v4l_rowmap c_map;  const int32_t* c_koff;
  const float* mask;  v4l_rowmap mask_map;
  int32_t M, N, K;
  int32_t flags;
} v4l_gemm_args;
int v4l_gemm_rows(v4l_ctx* ctx, void* stream, const v4l_gemm_args* args);

/* dW[n*ldw + k] = sum_m dY(m,n) * A(m,k);  dbias[n] = sum_m dY(m,n)   (deterministic split-M)
 * Weight/bias gradients of the same layers.                                                  */
typedef struct {
  const float* dy;  v4l_rowmap dy_map;
  const float* a;   v4l_rowmap a_map;  const int32_t* a_koff;
  float*       dw;  int64_t ldw;
  float*       dbias;          /* or NULL */
  int32_t M, N, K;
} v4l_wgrad_args;
int v4l_gemm_wgrad(v4l_ctx* ctx, void* stream, const v4l_wgrad_args* args);

/* out(m,n) = dy(m,n) * (act(m,n) > 0): ReLU backward between two row-mapped [M,N] views (used
 * where a layer output lands in a token slot, reference base.py:613-615).                   */
int v4l_relu_bwd(v4l_ctx* ctx, void* stream, const float* dy, const v4l_rowmap* dy_map,
                 const float* act, const v4l_rowmap* act_map, float* out,
                 const v4l_rowmap* out_map, int M, int N);

/* dx[b,h,w,c] = (x[b,h,w,c] > 0) * sum_{kh,kw} dcol[(b,oh,ow), (c,kh,kw)]  with
 * oh*stride + kh == h, ow*stride + kw == w.  NHWC activations, (c,kh,kw) column order (the
 * OIHW weight order of torch.nn.Conv2d).  Data-gradient of conv2/conv3
 * (reference torchrl/networks/base.py:320-323).  x may be NULL (no ReLU mask).               */
int v4l_col2im(v4l_ctx* ctx, void* stream, const float* dcol, const float* x, float* dx,
               int B, int Hin, int Win, int C, int KH, int KW, int stride, int Hout, int Wout);

/* ---- transformer block pieces (nn.TransformerEncoderLayer(d, n_head, ff, dropout=0),
 *      instantiated at reference torchrl/networks/nets.py:949-955; math: SURVEY Appendix A2) */
/* qkv [B,T,3d] (q|k|v) -> o [B,T,d], p [B,nh,T,T] (softmax probabilities, saved for bwd)    */
int v4l_attn_fwd(v4l_ctx* ctx, void* stream, const float* qkv, float* o, float* p,
                 int B, int T, int d, int n_head);
int v4l_attn_bwd(v4l_ctx* ctx, void* stream, const float* qkv, const float* p, const float* d_o,
                 float* d_qkv, int B, int T, int d, int n_head);
/* y = LayerNorm(a + res) * gamma + beta; z = a + res and stats = (mean, rstd) per row saved */
int v4l_ln_fwd(v4l_ctx* ctx, void* stream, const float* a, const float* res, const float* gamma,
               const float* beta, float* y, float* z, float* stats, int rows, int d, float eps);
int v4l_ln_bwd(v4l_ctx* ctx, void* stream, const float* dy, const float* z, const float* stats,
               const float* gamma, float* dz, float* dgamma, float* dbeta, int rows, int d);
/* mode 0: out[b] = [ tok[b,0,:] | mean_t>=1 tok[b,t,:] ]  (LocoTransformer, nets.py:1015-1034)
 * mode 1: out[b] = mean_t tok[b,t,:]                       (Transformer, nets.py:887-900)   */
int v4l_pool_fwd(v4l_ctx* ctx, void* stream, const float* tok, float* out, int B, int T, int d, int mode);
int v4l_pool_bwd(v4l_ctx* ctx, void* stream, const float* dout, float* dtok, int B, int T, int d, int mode);

/* f16-IO variants of the block pieces for the tensor-core tier (activations and activation
 * gradients f16; softmax probabilities, LayerNorm inputs z and statistics, dgamma/dbeta fp32) */
int v4l_attn_fwd_f16(v4l_ctx* ctx, void* stream, const void* qkv, void* o, float* p,
                      int B, int T, int d, int n_head);
int v4l_attn_bwd_f16(v4l_ctx* ctx, void* stream, const void* qkv, const float* p, const void* d_o,
                      void* d_qkv, int B, int T, int d, int n_head);
int v4l_ln_fwd_f16(v4l_ctx* ctx, void* stream, const void* a, const void* res, const float* gamma,
                    const float* beta, void* y, float* z, float* stats, int rows, int d, float eps);
int v4l_ln_bwd_f16(v4l_ctx* ctx, void* stream, const void* dy, const float* z, const float* stats,
                   const float* gamma, void* dz, float* dgamma, float* dbeta, int rows, int d,
                   float out_scale);
int v4l_pool_fwd_f16(v4l_ctx* ctx, void* stream, const void* tok, void* out, int B, int T, int d, int mode);
int v4l_pool_bwd_f16(v4l_ctx* ctx, void* stream, const void* dout, void* dtok, int B, int T, int d, int mode);

/* Tensor-core attention core (single head, d = 64): several samples packed per 128-row tile, the
 * per-sample TxT attention as block-diagonal tcgen05 MMAs (vision4leg_b200/csrc/tc_attn.cu).
 * qkv fp16 [B*T,192], o fp16 [B*T,64], p fp32 [B,T,T], d_o fp16 [B*T,64], d_qkv fp16 [B*T,192].  */
int v4l_tc_attn_fwd(v4l_ctx* ctx, void* stream, const void* qkv, void* o, float* p, int B, int T);
int v4l_tc_attn_bwd(v4l_ctx* ctx, void* stream, const void* qkv, const float* p, const void* d_o,
                    void* d_qkv, int B, int T);

/* "Flat" valid convolution (NatureCNN trunk forward, reference torchrl/networks/base.py:304-342):
 * x fp16 [x_rows, C] = images flattened to rows img * P + h * Wg + w (C a multiple of 64); each
 * 128-row tile is loaded once and tap (dh, dw) is the same shared-memory tile read through a UMMA
 * descriptor shifted by dh * Wg + dw rows (vision4leg_b200/csrc/tc_conv.cu).  Output rows with
 * h < Hout, w < Wout are stored through c_map at row index (img * Hout + h) * Wout + w; w = packed
 * fp16 [N_pad, n_taps * C] (tap-major K).  x_idx (optional, P % 128 == 0): image i of the problem is
 * image x_idx[i] of x.  mode bits 0-3: 1 = one copy of the tile, shifted descriptor starts (default);
 * 0 = one pre-shifted copy per (shift mod 8), atom-aligned starts.  Bits 4-7: tiles whose MMAs are
 * issued interleaved (0 = 2).                                                                    */
typedef struct v4l_tc_conv_flat_args {
  const void* x; int64_t x_rows; int32_t C, P, Wg, Hout, Wout;
  int32_t n_taps; int32_t tap_dw[16], tap_dh[16];
  const void* w; int32_t N_pad, N_valid; const float* bias;
  void* c; v4l_rowmap c_map;
  int64_t n_img; const int32_t* x_idx; int32_t flags, mode;
} v4l_tc_conv_flat_args;
int v4l_tc_conv_flat(v4l_ctx* ctx, void* stream, const v4l_tc_conv_flat_args* args);

/* Fused forward of one post-norm TransformerEncoderLayer (d=64, 1 head, FFN 256, ReLU, dropout 0;
 * reference torchrl/networks/nets.py:949-955 -> torch nn.TransformerEncoderLayer) as ONE kernel:
 * QKV GEMM -> block-diagonal attention -> out-proj -> +x, LayerNorm1 -> FFN1+ReLU -> FFN2 -> +h,
 * LayerNorm2, six tcgen05 contractions chained through shared memory
 * (vision4leg_b200/csrc/tc_block.cu).  x/y fp16 [B*T,64]; weights fp16 row-major [out,in]
 * (w_in [192,64], w_o [64,64], w_1 [256,64], w_2 [64,256]); biases / LayerNorm affine fp32.
 * Saved for the backward: qkv fp16 [B*T,192], p fp32 [B,T,T], o fp16 [B*T,64], z1/z2 fp32 [B*T,64]
 * (pre-norm sums), st1/st2 fp32 [B*T,2] (mean, rstd), h fp16 [B*T,64], f1 fp16 [B*T,256].     */
typedef struct v4l_tc_block_args {
  const void* x; int32_t B, T; float eps;
  const void *w_in, *w_o, *w_1, *w_2;
  const float *b_in, *b_o, *g1, *be1, *b1, *b2, *g2, *be2;
  void *qkv, *o, *h, *f1, *y;
  float *p, *z1, *st1, *z2, *st2;    /* z1 / z2 may be NULL */
  void *xh1, *xh2;                   /* fp16 [B*T,64] normalised rows (before the affine), or NULL */
} v4l_tc_block_args;
int v4l_tc_block_fwd(v4l_ctx* ctx, void* stream, const v4l_tc_block_args* args);

/* Fused data-gradient pass of the same layer (torch autograd of nn.TransformerEncoderLayer):
 * dy fp16 [B*T,64] -> dx fp16 [B*T,64], storing the row gradients the weight-gradient GEMMs consume:
 * dz2 [B*T,64], df1 [B*T,256], dh [B*T,64], dz1 [B*T,64], dqkv [B*T,192] (all fp16).
 * Weights in data-gradient orientation, fp16 row-major: w2d [256,64] = W2^T, w1d [64,256] = W1^T,
 * wod [64,64] = Wo^T, wind [64,192] = Win^T.  g1, g2 = LayerNorm weights; xh1, xh2, st1, st2, qkv, p, f1 as saved
 * by v4l_tc_block_fwd.  Weight / bias / LayerNorm-affine gradients: v4l_tc_wgrad on
 * (x,dqkv) (o,dz1) (h,df1) (f1,dz2) (xh1,dh) (xh2,dy).                                         */
typedef struct v4l_tc_block_bwd_args {
  const void* dy; int32_t B, T, pad_;
  const void *qkv, *xh1, *xh2, *f1;
  const float *p, *st1, *st2, *g1, *g2;
  const void *w2d, *w1d, *wod, *wind;
  void *dz2, *df1, *dh, *dz1, *dqkv, *dx;
} v4l_tc_block_bwd_args;
int v4l_tc_block_bwd(v4l_ctx* ctx, void* stream, const v4l_tc_block_bwd_args* args);
/* Profiling hook: per-phase globaltimer stamps (ns) of CTA 0 in the most recent fused forward
 * (host_out[0..31]) and data-gradient (host_out[32..63]) launch; synchronises the device.  */
int v4l_tc_block_timeline(unsigned long long* host_out);

/* ---- GAE / discounted return: reverse segmented scan over the rollout buffer
 *      (reference torchrl/replay_buffers/on_policy.py:17-71; recurrence: SURVEY Appendix A4).
 * rewards/values/terminals/advs/rets: [T,E] fp32; time_limits addressed t*tl_st + e*tl_se
 * ([T,1] -> (1,0), [T,E] -> (E,1)); last_value [E].  Arithmetic is float64 in registers.
 * mode 0 = GAE(gamma,tau), mode 1 = discount_reward(gamma).                                  */
int v4l_gae(v4l_ctx* ctx, void* stream, const float* rewards, const float* values,
            const float* terminals, const float* time_limits, int64_t tl_st, int64_t tl_se,
            const float* last_value, float* advs, float* rets, int T, int E,
            double gamma, double tau, int time_limit_filter, int mode);

/* ---- PPO loss epilogue (reference torchrl/algo/on_policy/ppo.py:42-153,
 *      torchrl/policies/continuous_policy.py:127-146) --------------------------------------
 * `info` rows are float[32]; the row written is info + 32 * (*slot).  Layout: V4L_INFO_*.   */
enum {
  V4L_INFO_ADV_MEAN = 0, V4L_INFO_ADV_STD, V4L_INFO_ADV_MAX, V4L_INFO_ADV_MIN,
  V4L_INFO_VF_LOSS, V4L_INFO_GRAD_NORM_VF, V4L_INFO_POLICY_LOSS,
  V4L_INFO_LP_MEAN, V4L_INFO_LP_STD, V4L_INFO_LP_MAX, V4L_INFO_LP_MIN,
  V4L_INFO_LS_MEAN, V4L_INFO_LS_STD, V4L_INFO_LS_MAX, V4L_INFO_LS_MIN,
  V4L_INFO_RATIO_MAX, V4L_INFO_RATIO_MIN, V4L_INFO_GRAD_NORM_PF,
  V4L_INFO_COUNT = 18, V4L_INFO_STRIDE = 32
};
/* cur_idx[i] = flat_idx[(*slot) * n + i]  — selects this minibatch's rollout rows            */
int v4l_select_rows(v4l_ctx* ctx, void* stream, const int32_t* flat_idx, const int32_t* slot,
                    int32_t* cur_idx, int n);
int v4l_slot_advance(v4l_ctx* ctx, void* stream, int32_t* slot, int32_t wrap);
/* stats (double[8]) = { sum, sumsq, n, max, min } of adv[idx[i]], i<n  (ppo.py:142-148)     */
int v4l_adv_stats(v4l_ctx* ctx, void* stream, const float* adv, const int32_t* idx, int n,
                  double* stats);
/* critic loss + d loss / d value  (ppo.py:94-114).  values [n]; returns/old_values gathered
 * through idx.  inv_global = 1 / (global minibatch), inv_local = 1 / n.                      */
int v4l_vf_loss(v4l_ctx* ctx, void* stream, const float* values, const float* returns,
                const float* old_values, const int32_t* idx, float* d_values, int n,
                float inv_global, float inv_local, int clipped, float clip_para,
                float* info, const int32_t* slot,
                void* d_values_f16 /* optional f16 [n,16]: scale_f16 * d_values in column 0, zero padded:
                                      the tensor-core tier's backward starts from it */, float scale_f16);
/* actor loss: log-prob, ratio, clipped surrogate, entropy bonus and their gradients w.r.t.
 * mean [n,A] and logstd [A]  (ppo.py:42-92).  target_mean/target_logstd come from the frozen
 * target policy; adv is normalised with the (possibly all-reduced) stats of v4l_adv_stats.
 * The target policy is frozen for a whole update_per_epoch (ppo.py:34), so its mean for a rollout row
 * can be computed once (first opt-epoch) into a [N,A] table and re-read afterwards: target_indexed. */
int v4l_pf_loss(v4l_ctx* ctx, void* stream, const float* mean, const float* logstd,
                const float* target_mean, const float* target_logstd, const float* acts,
                const float* adv, const int32_t* idx, const double* adv_stats,
                float* d_mean, float* d_logstd, int n, int A, float inv_global, float inv_local,
                float clip_para, float entropy_coeff, float* info, const int32_t* slot,
                int target_indexed /* 0: target_mean is [n,A] (row i); 1: a per-rollout table read at row idx[i] */,
                void* d_mean_f16 /* optional f16 [n,16] = scale_f16 * d_mean, zero padded (A <= 16) */, float scale_f16,
                int stats_per_slot /* 1: adv_stats is a [minibatches, 8] table (v4l_adv_stats_epoch) read at row *slot */);
/* stats[mb] (double[8]) = { sum, sumsq, n, max, min } of adv[flat_idx[mb*n + i]], i < n, for every minibatch of an
 * epoch in one launch (the row lists and the advantages are fixed once GAE has run)              */
int v4l_adv_stats_epoch(v4l_ctx* ctx, void* stream, const int32_t* flat_idx, int n_minibatches, int n,
                        const float* adv, double* stats);

/* ---- clip_grad_norm_(0.5) + Adam(eps=1e-5) over a flat bucket
 *      (reference ppo.py:71-75,116-120; a2c.py:30-40).
 * hyper (device, float[8]) = { lr, beta1, beta2, eps, max_norm, step (as float), 0, 0 };
 * the kernel reads step, uses step+1 for the bias corrections and stores step+1 back.
 * norm_slot: index into the info row that receives the pre-clip total norm (or -1).          */
int v4l_clip_adam(v4l_ctx* ctx, void* stream, float* param, const float* grad, float* m,
                  float* v, int64_t n, float* hyper, float* info, const int32_t* slot,
                  int norm_slot);

/* ---- tensor-core tier (f16 operands, fp32 accumulate; tcgen05.mma fed by TMA) ---------------
 * D[row tile, N] = sum_{tap,kc} A_tap[rows, 64k] * W[N, (tap,kc,64k)]^T  (+bias, ReLU, ReLU-mask,
 * accumulate), see vision4leg_b200/csrc/tc_gemm.cu.  A: f16 NHWC activation [a_B,a_H,a_W,a_C]
 * (plain matrices: a_H = a_W = 1); a row tile is the TMA box {64, bw, bh, bb} shifted per tap by
 * (tap_dw, tap_dh) with hardware zero fill outside the tensor.  W: packed f16
 * [N_pad, n_taps*kchunks*64].  Output rows are the logical positions (b, h, w) of a
 * [B, Hout, Wout] grid addressed through c_map (+ column n); c is f16 unless c_f32.
 * Same reference layers as v4l_gemm_rows.                                                    */
typedef struct {
  const void* a;  int32_t a_B, a_H, a_W, a_C;
  int64_t a_sW, a_sH, a_sB;      /* element strides of the W/H/B dims; 0,0,0 = packed NHWC       */
  const int32_t* a_idx;          /* optional item gather (minibatch rows); needs bb == 1         */
  int32_t B, Hout, Wout;
  int32_t bw, bh, bb;
  int32_t n_taps, kchunks;
  int32_t tap_dw[16], tap_dh[16];
  const void* w;  int32_t N_pad, N_valid;
  const float* bias;
  void* c;  v4l_rowmap c_map;  int32_t c_f32;
  const void* mask;
  int32_t flags;
  const void* res;               /* optional fp16 residual added after the mask (addressed like c) */
} v4l_tc_gemm_args;
int v4l_tc_gemm(v4l_ctx* ctx, void* stream, const v4l_tc_gemm_args* args);
/* dw[index[n*Kp + kp]] = sum_rows X_tap[row, kp] * dY[row, n], Kp = n_taps * x_C, through
 * tcgen05 with MN-major operands (no transposed copies); x: f16 [x_B,x_H,x_W,x_C], dy: f16
 * [B,Hout,Wout,dy_C]; the row tiles are the same boxes as the forward pass.  index = the
 * weight-packing table (or NULL for dw[n*Kp + kp]).  Deterministic (fixed split order).       */
typedef struct {
  const void* x;   int32_t x_B, x_H, x_W, x_C;
  int64_t x_sW, x_sH, x_sB;      /* element strides; 0,0,0 = packed                              */
  int32_t x_estride;             /* traversal stride of X along W and H (1 or 2)                 */
  const int32_t* x_idx;          /* optional item gather; needs bb == 1                          */
  const void* dy;  int32_t dy_C;
  int64_t dy_sW, dy_sH, dy_sB;
  /* optional sub-iterations per row tile (sub-positions of a space-to-depth cell): X shifted by
   * (sub_dw, sub_dh), dY read from channel offset sub_dyc; n_sub = 0 means one plain pass      */
  int32_t n_sub;  int32_t sub_dw[4], sub_dh[4], sub_dyc[4];
  int32_t B, Hout, Wout;
  int32_t bw, bh, bb;
  int32_t n_taps;
  int32_t tap_dw[16], tap_dh[16];
  int32_t N_valid;
  const int32_t* index;
  float* dw;
  float out_scale;               /* fp32 results are multiplied by this (1/loss-scale); 0 = 1    */
  float* dbias;                  /* optional: dbias[n] = sum_rows dY[row, n] from the SAME launch
                                    (an extra K slice whose X operand is the constant 1)         */
  int32_t defer;                 /* 1: leave the split partials in the context and sum them in
                                    the next v4l_tc_wgrad_flush (one launch for a whole backward) */
  int32_t accumulate;            /* 1: dw += (gradient accumulation over micro-batches) */
} v4l_tc_wgrad_args;
int v4l_tc_wgrad(v4l_ctx* ctx, void* stream, const v4l_tc_wgrad_args* args);
int v4l_tc_wgrad_flush(v4l_ctx* ctx, void* stream);
/* Weight + bias gradient of the first NatureCNN convolution (Conv2d(4,32,8,stride 4), reference
 * torchrl/networks/base.py:317-318) on the space-to-depth layouts: x_s2d f16 [n_img,16,16,64] (v4l_ingest_img),
 * dy_cells f16 [B,8,8,128] (gradient of conv1's output stored as 2x2 cells, channel = (py*2+px)*32 + n).
 * One CTA per image range; each pixel window and each dY cell is loaded once (csrc/tc_wgrad_s2d.cu).
 * index / dw / dbias / out_scale / defer / accumulate as v4l_tc_wgrad (packed K = (dy*2+dx)*64 + c).   */
int v4l_tc_wgrad_conv1(v4l_ctx* ctx, void* stream, const void* x_s2d, int64_t n_img, const int32_t* x_idx,
                       const void* dy_cells, int B, const int32_t* index, float* dw, float* dbias,
                       float out_scale, int defer, int accumulate);

/* ---- observation pipeline on the device (csrc/obs_ops.cu) -----------------------------------
 * v4l_depth_frame: OpenGL depth-buffer values zbuf [E,64,64] -> far*near/(far-(far-near) z) -> clip [0.3,10] ->
 * sqrt(log(d+1)) into slot `head` of the per-env frame ring [E, n_slots, 64, 64] fp32 — into EVERY slot of env e when
 * reset[e] != 0 (reset may be NULL; an episode start fills the history, :635-637)
 * (reference vision4leg/envs/locomotion_gym_env_with_rich_information.py:620-633).
 * v4l_stack_frames: the observation's 4 channels = ring slots slots[e*4 + c] (the reference's deque indices
 * frame_idx, :315-336,549-554,641-648), optionally (x-1.25)/0.425 (:649-650), written as the f16 4x4
 * space-to-depth image [E,16,16,64] (v4l_ingest_img layout) and / or as fp32 CHW rows (row stride chw_stride).
 * v4l_normalizer: running-mean observation normaliser over x [n,S]: if update, merge the batch mean / population
 * variance into (mean, var, count) (torchrl/env/base_wrapper.py:44-61,84-86), then
 * out = clip((x - mean) / (sqrt(var) + 1e-4), +-clip) (:88-90,119-122).  mean/var: double[S] on the device; count (1e-4 + rows merged so far) is tracked by the caller,
 * who adds n after an updating call.  out may be NULL (update only) or alias x.  update = 2: write the BATCH mean /
 * population variance of x into mean / var and do nothing else (data-parallel callers merge them across ranks). */
int v4l_depth_frame(v4l_ctx* ctx, void* stream, const float* zbuf, float* ring, const uint8_t* reset, int E,
                    int n_slots, int head, float near_plane, float far_plane);
int v4l_stack_frames(v4l_ctx* ctx, void* stream, const float* ring, const int32_t* slots, int E, int n_slots,
                     int normalise, void* out_s2d, float* out_chw, int64_t chw_stride);
int v4l_normalizer(v4l_ctx* ctx, void* stream, const float* x, int n, int S, double* mean, double* var,
                   double count, int update, float clip, float* out);

/* ---- a chain of up to three Linear layers in one launch (csrc/tc_mlp.cu): the actor / critic MLP heads
 * (reference torchrl/networks/nets.py:973-992,1036), the proprio MLP + projector (base.py:8-44,209-230) and
 * their data-gradient chains.  Per 128-row tile: h = act(x W1^T + b1) [* (mask1 > 0)] stays in shared memory
 * as the next layer's operand; every layer can also store its result (saved activation / row gradient / fp32
 * output through a row map).  x: f16 [M, x_cols] with row pitch x_ld; layer l: packed f16 W [N_pad, K]
 * (v4l_pack_f16 layout, K a multiple of 64; K_l <= 256 for l > 0 = the previous layer's width padded to 64). */
typedef struct {
  const void* w; int32_t K, N_pad, N_valid; const float* bias; int32_t relu;
  const void* mask; int64_t mask_ld;        /* optional f16 [M, mask_ld]: result *= (mask > 0) */
  void* out; int32_t out_f32; v4l_rowmap out_map;   /* optional global output */
} v4l_tc_mlp_layer;
typedef struct {
  const void* x; int32_t M, x_cols; int64_t x_ld;
  int32_t n_layers;
  v4l_tc_mlp_layer layer[3];
} v4l_tc_mlp_chain_args;
int v4l_tc_mlp_chain(v4l_ctx* ctx, void* stream, const v4l_tc_mlp_chain_args* args);

/* ---- fused optimiser tail of one network pass (tensor-core tier; csrc/step_ops.cu):
 * phase 1  split-K reduction of all deferred weight-gradient partials into the fp32 gradient bucket
 *          (accumulating the squared norm of what it writes);
 * phase 2  clip_grad_norm_(max_norm) + Adam over the flat bucket (same arithmetic as v4l_clip_adam,
 *          reference ppo.py:71-75,116-120; a2c.py:30-40), pre-clip norm into info[norm_slot]; every
 *          updated parameter i is also written, rounded to f16, at scatter[i] = (a, b, c, d): positions
 *          a, b of packed_self and c, d of packed_other (-1 = none) — the tap-major operand copies
 *          (v4l_pack_f16 layouts) of this network and, for shared-encoder weights, of the other one;
 *          finally the Adam step counter and the optional minibatch slot advance.
 * `phases` is a bit mask; 1|2 is two launches (reduction + norm partials, then the step).  Data-parallel
 * runs launch phase 1, all-reduce the bucket, then launch phase 2 (which takes the norm from the bucket).  extra_lo/extra_n: the range of the
 * bucket that is NOT written by a reduction job (logstd) and must be added to the norm in 1|2.      */
typedef struct {
  int32_t phases;
  float* param; float* grad; float* m; float* v; int64_t n;
  float* hyper;                 /* as v4l_clip_adam */
  float* info; const int32_t* slot; int32_t norm_slot;
  int64_t extra_lo, extra_n;
  const int32_t* scatter;       /* int32 [n][4] or NULL */
  void* packed_self; void* packed_other;
  int32_t* slot_advance;        /* optional: incremented by 1 at the very end */
} v4l_opt_tail_args;
int v4l_opt_tail(v4l_ctx* ctx, void* stream, const v4l_opt_tail_args* args);
/* 1 if a device-wide barrier of v4l_opt_tail ever timed out (synchronises the device)           */
int v4l_opt_tail_error(v4l_ctx* ctx);
/* minibatch prologue in one launch: cur_idx[i] = flat_idx[(*slot)*n + i] (v4l_select_rows), the
 * advantage statistics of those rows (v4l_adv_stats) and, if state_f16 != NULL, the proprio rows
 * state[cur_idx[i], :S] converted to f16 and zero padded to Sp columns                           */
int v4l_mb_begin(v4l_ctx* ctx, void* stream, const int32_t* flat_idx, const int32_t* slot,
                 int32_t* cur_idx, int n, const float* adv, double* stats, const float* state,
                 int S, void* state_f16, int Sp);
/* out[n] = sum_m sum_f dy(m, f*N + n) for a row-mapped f16 [M, N*fold <= 256] view (bias
 * gradients; fold > 1 sums the sub-positions of a space-to-depth cell)                         */
int v4l_colsum_f16(v4l_ctx* ctx, void* stream, const void* dy, const v4l_rowmap* map, int M, int N,
                   int fold, float out_scale, float* out);
/* dst_f16[i] = index ? (index[i] >= 0 ? src[index[i]] : 0) : src[i]  — weight packing /
 * fp32 -> f16 conversion for the tensor-core tier                                            */
int v4l_pack_f16(v4l_ctx* ctx, void* stream, const float* src, const int32_t* index, void* dst,
                  int64_t n);

/* fp32 CHW [n,4,64,64] depth stack -> f16 4x4 space-to-depth NHWC [n,16,16,64]
 * (channel = (py*4+px)*4+c): the layout the tensor-core conv1 reads (one swizzle atom per tap) */
int v4l_ingest_img(v4l_ctx* ctx, void* stream, const float* img, void* out_s2d, int64_t n_img,
                   const int32_t* idx /* optional list of the n_img image rows to convert */);
/* Same re-ordering from an fp16 [n, 4, 64, 64] source: the replay buffer's half-precision staging copy
 * of the depth stack (vision4leg_b200/replay_buffers/on_policy.py), which halves the host->device
 * bytes of the streamed ingest.  idx: optional row list.                                         */
int v4l_ingest_img_f16(v4l_ctx* ctx, void* stream, const void* img_f16, void* out_s2d, int64_t n_img,
                       const int32_t* idx);
/* Zero-copy rollout ingest: row n (= idx[i] or i) of the [*, row_stride] fp32 observation matrix in
 * PINNED HOST (or device) memory -> state_out[n, 0:S] fp32, img_out[n, 0:16384] fp32 (optional),
 * s2d_out[n] fp16 [16,16,64] (optional), one CTA per row, PCIe reads issued by the SMs.
 * Replaces reference on_policy.py:83-89 + ppo.py:136-140 (float64 gather, f64->f32, H2D).       */
int v4l_ingest_rows(v4l_ctx* ctx, void* stream, const float* obs, int64_t row_stride, int S,
                    const int32_t* idx, int64_t n_rows, float* state_out, float* img_out,
                    void* s2d_out);
/* dst_f16[i, 0:dst_cols] = src[idx ? idx[i] : i, 0:src_cols] zero padded (proprio rows -> K-padded
 * f16 operand; also fp32 -> f16 conversion of loss gradients)                                */
int v4l_gather_rows_f16(v4l_ctx* ctx, void* stream, const void* src, int src_is_f32,
                        const int32_t* idx, void* dst, int rows, int src_cols, int64_t src_stride,
                        int dst_cols, float scale);
int v4l_relu_bwd_f16(v4l_ctx* ctx, void* stream, const void* dy, const v4l_rowmap* dy_map,
                      const void* act, const v4l_rowmap* act_map, void* out,
                      const v4l_rowmap* out_map, int M, int N);

/* ---- rollout ingest: strided host->device copy (pinned host rows -> aligned device planes);
 *      replaces the float64 fancy-index copy + torch.Tensor(...).to(device) of
 *      reference on_policy.py:83-89 / ppo.py:136-140.  Sizes in bytes.                       */
int v4l_h2d_2d(void* stream, void* dst, size_t dpitch, const void* h_src, size_t spitch,
               size_t width, size_t height);
/* Streamed ingest, copy-engine leg: rows[i] (host int32, any order) of a pinned host matrix with
 * row_bytes-byte rows -> the same rows of a device matrix; one batched copy per call.            */
int v4l_h2d_rows(void* stream, void* dst_base, const void* src_base, const int32_t* rows, int n_rows,
                 size_t row_bytes);

#ifdef __cplusplus
}
#endif
#endif  /* V4L_B200_H_ */
