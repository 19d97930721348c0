/*
 * probunet_b200 -- C ABI of the hand-written sm_100a kernels behind the Probabilistic U-Net hot path.
 *
 * The reference (pierrelouislemaire/prob-unet-mds) is pure Python on top of PyTorch and has no FFI of its
 * own; each entry point below replaces the ATen/cuDNN/cuBLAS call(s) that the cited reference line issues.
 * Conventions:
 *   - every function returns 0 on success, <0 on error; pu_last_error() gives the message (thread local);
 *   - activations are NHWC ("pixels x channels"), dtype PU_F32 or PU_BF16; reductions/statistics/grads of
 *     parameters are always fp32 (GroupNorm sums: fp64);
 *   - the library never allocates or frees user-visible memory; every pointer is a device pointer owned by
 *     the caller (PyTorch), every launch goes to the cudaStream_t passed as `stream` (void*);
 *   - "packed" conv weights are [Cout][kh][kw][Cin] (forward) in the activation dtype, see pu_pack_conv_weight.
 */
#ifndef PROBUNET_B200_H
#define PROBUNET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PU_F32 0
#define PU_BF16 1

#define PU_RS_NONE 0
#define PU_RS_UP 1   /* nearest x2: networks.py:82-83 (conv_transpose2d with an all-ones 2x2 filter) */
#define PU_RS_DOWN 2 /* 2x2 mean : networks.py:84-85 (depthwise stride-2 conv with a 0.25 filter)    */

#define PU_CONV_RELU 1     /* epilogue ReLU                 (prob_unet.py:34, :94, :96)              */
#define PU_CONV_FORCE_SIMPLE 2 /* use the CUDA-core kernel even where the tcgen05 kernel applies     */
#define PU_CONV_FORCE_TC 4     /* fail instead of falling back to the CUDA-core kernel               */

const char* pu_last_error(void);
int pu_version(void);
/* 1 when the tcgen05 kernels can run on the current device (compute capability 10.x) */
int pu_device_supports_tc(void);
/* number of kernels launched by this library since the last reset (per process) */
long long pu_launch_count(int reset);
/* Stream-ordered zero fill of `bytes` bytes (cudaMemsetAsync): the loss / statistics accumulators and the zero-padded
 * input channels that the host side would otherwise clear with torch.zeros (an ATen fill kernel). */
int pu_zero(void* dst, long long bytes, void* stream);
/* Stream-ordered device-to-device copy (cudaMemcpyAsync), e.g. the skip-conv bias gradient, which has the same values as
 * conv1's bias gradient (networks.py:176-177: both biases are added to the same sum) but needs its own storage. */
int pu_copy(void* dst, const void* src, long long bytes, void* stream);

/* ---------------- layout / packing ---------------- */
/* fp32 NCHW [N,C,H,W] -> NHWC dst[..., c_off : c_off+C] of a tensor with Cdst channels. Replaces the implicit
 * layout of torch.cat([x, target], 1) (prob_unet.py:58) and the fp32->bf16 cast of the model inputs.            */
int pu_nchw_to_nhwc(const float* src, void* dst, int N, int C, int H, int W, int Cdst, int c_off, int dst_dtype,
                    void* stream);
/* NHWC (dtype) -> fp32 NCHW; the model output handed back to the caller (prob_unet.py:195-196).  The source has Csrc >= C
 * channels per pixel and its first C are taken (a 3-channel output head computed in a zero-padded 64-channel tile). */
int pu_nhwc_to_nchw(const void* src, float* dst, int N, int C, int H, int W, int Csrc, int src_dtype, void* stream);
/* fp32 NCHW grad <- NHWC fp32 etc. are not needed: the reference never asks for input gradients. */

/* OIHW fp32 master weight -> packed weight in `dtype`.
 *   mode 0 (forward): dst[co'][ky][kx][ci]           = src[perm(co')][ci][ky][kx]
 *   mode 1 (dgrad)  : dst[ci][ky][kx][co']           = src[perm(co')][ci][k-1-ky][k-1-kx]
 *   mode 2 (forward, bf16 hi/lo split): dst[co'][ky][kx][0..Ci_pad) = bf16(w), [Ci_pad..2*Ci_pad) = bf16(w - bf16(w));
 *           used with pu_conv2d(src0 = x, src1 = x) so that (w_hi + w_lo) * x is accumulated: the prior / posterior
 *           encoders (prob_unet.py:44-78) whose bf16 weight rounding would otherwise cost 1e-3..2e-3 of the KL term
 * ci is padded with zeros up to Ci_pad (mode 0) ; out_perm (device int32[Co]) may be NULL (identity).
 * Replaces `self.weight.to(x.dtype)` (networks.py:69).                                                           */
int pu_pack_conv_weight(const float* src, void* dst, int Co, int Ci, int k, int Ci_pad, int mode,
                        const int* out_perm, long long src_co_stride /* 0: Ci*k*k */, int dtype, void* stream);
/* The same packing for MANY weights in one launch (the refresh of every packed copy after an optimizer step, SURVEY
 * 8f-1): `items` is a DEVICE array; item i owns the 32x32 (co, ci) tiles [tile_begin, next item's tile_begin), i.e.
 * ceil(Co/32) * ceil(Ci_pad/32) of them; total_tiles is their sum.  Field meanings as in pu_pack_conv_weight. */
typedef struct PuPackItem {
    const float* src;
    void* dst;
    const int* perm;          /* or NULL */
    long long src_co_stride;  /* elements; Ci*k*k for a plain OIHW tensor */
    int Co, Ci, k, Ci_pad;
    int mode, dtype;
    int tile_begin, pad_;
} PuPackItem;
int pu_pack_conv_weights_multi(const PuPackItem* items, int n_items, int total_tiles, void* stream);
/* packed fp32 weight gradient [Co'][k][k][Ci_pad] -> OIHW fp32 .grad, dst = (accumulate ? dst : 0) + src */
int pu_unpack_conv_wgrad(const float* src, float* dst, int Co, int Ci, int k, int Ci_pad, const int* out_perm,
                         long long dst_co_stride /* 0: Ci*k*k */, int accumulate, void* stream);
/* dst[i] = src[perm[i]] (gather) or dst[perm[i]] (+)= src[i] (scatter); fp32 vectors (qkv bias permutation) */
int pu_gather_f32(const float* src, const int* perm, float* dst, int n, void* stream);
int pu_scatter_f32(const float* src, const int* perm, float* dst, int n, int accumulate, void* stream);

/* ---------------- convolution (networks.py:87 F.conv2d; prob_unet.py:33,41-42,93-97 nn.Conv2d) ---------------- */
/* Optional epilogue of a DATA-GRADIENT conv whose output is the gradient wrt the output y of
 *     y = dropout(act(u)),  u = xhat * gamma' + beta',  xhat = GroupNorm(x)        (networks.py:164-175)
 * i.e. of the conv that consumes a GroupNorm(+SiLU)(+dropout) result.  The epilogue turns the accumulator g = dL/dy
 * into du = dL/du (dropout mask, SiLU derivative) while it is still in registers, stores du instead of g, and
 * accumulates the two per-(sample, channel) sums of the GroupNorm backward -- the first pass of pu_gn_bwd, which is
 * then called with du_ready = 1 and only runs its second pass.  Requires the tcgen05 kernel (bf16, 64-multiples). */
typedef struct PuConvGnBwd {
    const void* x0;       /* the normalised tensor x (pre-norm activation), NHWC, C0 channels ...          */
    const void* x1;       /* ... and the second source of a concatenation (C1 channels) or NULL           */
    int C0, C1;           /* C0 + C1 == Cout of the conv                                                  */
    const float* consts;  /* [N][C][4] = (rstd*gamma', beta' - mean*rstd*gamma', rstd, -mean*rstd): pu_gn_bwd_consts */
    double* sums;         /* [N][C][2]: receives sum_pixels du and sum_pixels du*xhat; zeroed by the call  */
    int silu;
    float dropout_p;
    unsigned long long seed;
    const void* keep_mask; /* optional: the keep bits pu_gn_apply stored (PuGnArgs.keep_mask); NULL: Philox(seed, index) */
} PuConvGnBwd;
typedef struct PuConvArgs {
    int N, H, W;          /* output == input spatial size (stride 1, padding k/2)                     */
    int C0, C1;           /* input channels of src0 / src1 (C1 = 0: single source).  src0||src1 is the */
                          /* torch.cat([x, skip], 1) of networks.py:330, never materialised            */
    int Cout;
    int ksize;            /* 1 or 3                                                                    */
    int dtype;            /* PU_F32 / PU_BF16 for src, weight, residual, out                           */
    int flags;            /* PU_CONV_*                                                                 */
    int bias_per_sample;  /* 0: bias[Cout]; 1: bias[N][Cout] (Fcomb: W0z.z_n + b0)                     */
    const void* src0;
    const void* src1;
    const void* weight;   /* packed [Cout][k][k][C0+C1]                                                */
    const float* bias;    /* fp32 or NULL (networks.py:88-89 x.add_(b))                                */
    const void* residual; /* NHWC [N,H,W,Cout] added in the epilogue or NULL (networks.py:176,183)     */
    void* out;            /* NHWC [N,H,W,Cout]; may alias residual                                     */
    double* qstats;       /* optional [N][Cout/4][2] fp64 (Cout % 4 == 0): receives (sum, sum of squares) of the */
                          /* STORED output over the pixels of sample n, per quad of 4 consecutive channels --    */
                          /* the GroupNorm statistics of the consumer (networks.py:104) taken in the producing   */
                          /* conv's epilogue instead of a separate pass; zeroed by the call.  Groups of any size  */
                          /* that is a multiple of 4, also over a channel concatenation of two such tensors, are  */
                          /* formed from the quads by pu_gn_stats_from_quads                                      */
    int reserved;         /* must be 0                                                                            */
    const PuConvGnBwd* gn_bwd; /* optional GroupNorm-backward epilogue (see above) or NULL                          */
} PuConvArgs;
/* forward conv and, with a mode-1 packed weight, data gradient (convolution_backward's grad_input) */
int pu_conv2d(const PuConvArgs* a, void* stream);

/* weight gradient (convolution_backward's grad_weight): dw[Cout][k][k][C0+C1] (fp32, packed) (+)=
 * sum_pixels dy[p][co] * src[p + tap][ci].  `accumulate` = 0 zeroes dw first.                        */
typedef struct PuWgradArgs {
    int N, H, W, C0, C1, Cout, ksize, dtype, flags;
    const void* src0;
    const void* src1;
    const void* dy;
    float* dw;
    int accumulate;
} PuWgradArgs;
int pu_conv2d_wgrad(const PuWgradArgs* a, void* stream);
/* db[c] (+)= sum over pixels of dy[p][c]   (bias gradient of x.add_(b)) */
int pu_bias_grad(const void* dy, float* db, long long pixels, int C, int dtype, int accumulate, void* stream);

/* ---------------- GroupNorm + SiLU (+adaptive scale/shift, dropout, resample) ---------------- */
/* stats[n][g] = (sum, sumsq) in fp64 over the group's channels of src0||src1 (F.group_norm, networks.py:104) */
int pu_gn_stats(const void* src0, const void* src1, int C0, int C1, int N, int HW, int G, int dtype,
                double* stats, void* stream);
/* stats[n][g] = sum of the quad statistics (PuConvArgs.qstats) of the group's channels; q1 / C1 describe the second
 * source of a channel concatenation (NULL / 0: single source); (C0 + C1) / G must be a multiple of 4 */
int pu_gn_stats_from_quads(const double* q0, const double* q1, int C0, int C1, int N, int G, double* stats, void* stream);
typedef struct PuGnArgs {
    int N, H, W;          /* spatial size of the normalised tensor x                                   */
    int C0, C1, G;
    int dtype;
    int silu;             /* 1: apply SiLU (networks.py:166,171,332)                                   */
    int resample;         /* PU_RS_*: y is written at (2H,2W) / (H/2,W/2)                              */
    float eps;
    float dropout_p;      /* 0 = off; keep mask from Philox(seed, element index), scaled by 1/(1-p)    */
    unsigned long long seed;
    const void* src0;
    const void* src1;
    const double* stats;  /* [N][G][2]                                                                 */
    const float* gamma;   /* [C]                                                                       */
    const float* beta;    /* [C]                                                                       */
    const float* ada;     /* NULL or [2C] = (scale, shift) == affine.bias (networks.py:168-171)        */
    void* y;              /* NHWC [N,H',W',C]                                                          */
    void* keep_mask;      /* optional uint8 [N*H*W*C/8], written by pu_gn_apply when dropout_p > 0: bit e of byte i */
                          /* = element 8i+e of y is kept.  Lets the PuConvGnBwd epilogue read the mask (4 bytes per */
                          /* 32 channels) instead of regenerating it with Philox; pu_gn_bwd reads it too when given */
} PuGnArgs;
int pu_gn_apply(const PuGnArgs* a, void* stream);

typedef struct PuGnBwdArgs {
    PuGnArgs f;           /* the forward call (y unused)                                               */
    const void* dy;       /* gradient wrt y, at the resampled resolution.  SCRATCH: when no resampling is */
                          /* involved the buffer is overwritten (with d loss / d pre-activation)          */
    const void* dres;     /* optional extra gradient added to dx (skip path), layout per dres_resample */
    int dres_resample;    /* PU_RS_NONE: dres is [N,H,W,C]; else it is at y's resolution               */
    double* sums;         /* workspace [N][C][2] fp64 (sum du, sum du*xhat)                            */
    void* dx0;            /* gradient wrt src0 [N,H,W,C0]                                              */
    void* dx1;            /* gradient wrt src1 [N,H,W,C1] or NULL                                      */
    int acc0, acc1;       /* 1: dx += ...                                                              */
    float* dgamma;        /* [C]  (+)=                                                                 */
    float* dbeta;         /* [C]                                                                       */
    float* dada;          /* [2C] or NULL                                                              */
    int acc_params;
    float* colsum0;       /* optional [C0]: sum over (n, pixels) of the final dx0 values written (incl. dres  */
    float* colsum1;       /* and the accumulated old value) = bias gradient of the conv that produced src0/1  */
    int du_ready;         /* 1: dy already holds du and sums is filled (PuConvArgs.gn_bwd epilogue of the conv that */
                          /* produced dy): only the second pass runs.  Needs resample == dres_resample == NONE ... */
} PuGnBwdArgs;
int pu_gn_bwd(const PuGnBwdArgs* a, void* stream);
/* consts[n][c] = the four per-(sample, channel) constants of PuConvGnBwd, from the forward call's statistics and affine
 * parameters (f->y unused) */
int pu_gn_bwd_consts(const PuGnArgs* f, float* consts, void* stream);

/* ---------------- attention (networks.py:112-125,179-184) ---------------- */
/* qkv is NHWC-flattened [N][T][3*C] with channel order (j in {q,k,v}, head, d) -- the product's own order,
 * obtained by permuting the rows of qkv.weight (pu_pack_conv_weight out_perm); d = 64.
 * out [N][T][C] with channel order (head, d).  lse [N][heads][T] fp32 = log-sum-exp of the scaled logits.  */
int pu_attention_fwd(const void* qkv, void* out, float* lse, int N, int T, int heads, int dtype, int flags,
                     void* stream);
/* workspaces: delta_ws fp32 [N][heads][T] (sum_d out*dout); dq_ws fp32 [N][T][C] (cross-key-tile reduction of dq,
 * only touched by the tcgen05 kernel; tile-major inside) followed, when dbias != NULL, by N*heads*ceil(T/128)*192
 * more floats (per-CTA partial column sums).  dbias (optional, fp32 [3C]) receives the column sums of dqkv over all
 * N*T pixels -- the gradient of the qkv conv's bias (networks.py:88-89, 179) -- from the kernels' epilogues instead
 * of a separate pass over dqkv.                                                                                   */
int pu_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                     float* delta_ws, float* dq_ws, float* dbias, int N, int T, int heads, int dtype, int flags,
                     void* stream);

/* ---------------- prior / posterior encoders (prob_unet.py:32-36,60-72) ---------------- */
/* y[N,2H,2W,C] = nearest-neighbour x2 of x (skip branch of the "up" blocks, networks.py:82-83,156) */
int pu_upsample2(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream);
/* transpose of the 2x resampling between a block's first GroupNorm(+SiLU) and its conv0 (networks.py:82-87, 164-166):
 * g [N,H,W,C] = gradient wrt the pre-resample activation from dy wrt the resampled one.  resample = PU_RS_UP: dy is
 * [N,2H,2W,C], g = sum of the four children; PU_RS_DOWN: dy is [N,H/2,W/2,C], g = 0.25 * dy[parent].  Lets pu_gn_bwd run
 * its streaming (resample = PU_RS_NONE) kernels for the up / down blocks. */
int pu_resample_grad(const void* dy, void* g, int N, int H, int W, int C, int resample, int dtype, void* stream);
/* y[N,H/2,W/2,C] = 2x2 mean of x */
int pu_avgpool2(const void* x, void* y, int N, int H, int W, int C, int dtype, void* stream);
/* dr = 0.25 * dp[parent] * (r > 0)   (backward of AvgPool2d(ReLU(.))) */
int pu_relu_pool_bwd(const void* dp, const void* r, void* dr, int N, int H, int W, int C, int dtype, void* stream);
/* m[n][c] = mean over pixels of x (fp32) */
int pu_global_mean(const void* x, float* m, int N, int HW, int C, int dtype, void* stream);
/* dr[n,p,c] = dm[n][c] / HW * (r > 0) */
int pu_relu_mean_bwd(const float* dm, const void* r, void* dr, int N, int HW, int C, int dtype, void* stream);
/* mu/log_sigma heads: out[n][l] = W[l] . m[n] + b[l]  for the stacked [2L][C] weight (conv_mu ; conv_log_sigma) */
int pu_heads_fwd(const float* m, const float* w, const float* b, float* out, int N, int C, int L2, void* stream);
int pu_heads_bwd(const float* m, const float* w, const float* dout, float* dm, float* dw, float* db, int N, int C,
                 int L2, int accumulate, int acc_dm, void* stream);

/* dy masked by (y > 0): ReLU backward (prob_unet.py:94,96), element count n must be a multiple of 8 */
int pu_relu_mask(const void* dy, const void* y, void* out, long long n, int dtype, void* stream);

/* ---------------- latent kernels (prob_unet.py:77,188,193,221,230) ---------------- */
/* z = mu + eps * exp(log_sigma) with separately rounded mul and add (bit-exact vs torch's loc + eps*scale);
 * sigma_out (optional) = exp(log_sigma); flag[0] |= 1 if any mu is non-finite or sigma <= 0 / non-finite
 * (the reference's Normal(validate_args) check, prob_unet.py:77)                                              */
int pu_rsample(const float* mu, const float* log_sigma, const float* eps, float* z, float* sigma_out, int* flag,
               int n, void* stream);
/* dmu += dz ; dls += dz * eps * sigma   (backward of rsample) */
int pu_rsample_bwd(const float* dz, const float* eps, const float* sigma, float* dmu, float* dls, int n, void* stream);
/* kl_acc[0] += sum_n sum_l KL(N(mu_q,sig_q) || N(mu_p,sig_p)) (fp64 accumulator); analytic gradients times *gscale
 * (device scalar, NULL = 1) written to dmu_q, dls_q, dmu_p, dls_p (any may be NULL)                            */
int pu_kl_fwd_bwd(const float* mu_q, const float* ls_q, const float* mu_p, const float* ls_p, double* kl_acc,
                  float* dmu_q, float* dls_q, float* dmu_p, float* dls_p, const float* gscale, int n, void* stream);
/* recon_acc[0] += sum (out - target)^2 over fp32 NCHW tensors (MSELoss(reduction='sum'), prob_unet.py:227);
 * dlogits (optional, NHWC [N][HW][Cdst >= C] in `dtype`) = *gscale * 2 * (out - target) in its first C channels; the
 * caller zero-fills the padding (3 logit channels in a 64-channel tile keep Fcomb's backward on the tensor-core kernels) */
int pu_mse_fwd_bwd(const float* out_nchw, const float* target, double* recon_acc, void* dlogits, const float* gscale,
                   int N, int C, int HW, int Cdst, int dtype, void* stream);
/* (total, recon, kl) = (acc[0] + beta*acc[1], acc[0], acc[1]) as fp32 scalars (prob_unet.py:232-234) */
int pu_loss_finalize(const double* acc, float beta, float* total, float* recon, float* kl, void* stream);
/* out2 = (g_total + g_recon, beta*g_total + g_kl): seeds of the two loss branches from the upstream gradients
 * of the three returned scalars (device pointers, any may be NULL = 0)                                        */
int pu_loss_bwd_scales(const float* g_total, const float* g_recon, const float* g_kl, float beta, float* out2,
                       void* stream);

/* ---------------- Fcomb (prob_unet.py:100-121) ---------------- */
typedef struct PuFcombArgs {
    int N, HW, L, dtype;
    int S;                 /* latent samples per input (ensemble members); z is [N][S][L]                 */
    int num_classes;       /* <= 3                                                                        */
    const void* feat;      /* [N][HW][64] in dtype                                                        */
    const float* z;
    const float* w0;       /* fp32 OIHW master weights: [64][64+L], [64][64], [num_classes][64]           */
    const float* b0;
    const float* w1;
    const float* b1;
    const float* w2;
    const float* b2;
    float* out_nchw;       /* [N][S][num_classes][HW] fp32 (S == 1: the reference's [N,C,H,W] output)     */
    void* h1_out;          /* optional (S == 1): hidden activations [N][HW][64] in dtype, kept for backward */
    void* h2_out;
} PuFcombArgs;
int pu_fcomb_fwd(const PuFcombArgs* a, void* stream);
/* gradients through the z half of layer 0, from rmean[n][o] = mean over pixels of d pre-activation-1:
 * dw0[:, 64:] and db0 (+)=, dz[n][l] =                                                                    */
int pu_fcomb_z_bwd(const float* rmean, float hw, const float* z, const float* w0, float* dz, float* dw0, float* db0,
                   int N, int L, int accumulate, void* stream);

/* ---------------- optimiser (fused AdamW, torch.optim.AdamW semantics, main.py:95) ---------------- */
int pu_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
             float eps, float weight_decay, int step, void* stream);
/* Multi-tensor form: one launch for the whole model.  `chunks` is a DEVICE array; each entry is a run of n <= 65536
 * fp32 elements of one parameter (device addresses of the parameter, its gradient and the two moment buffers at the same
 * offset).  Replaces the 446-tensor loop of torch.optim.AdamW.step (main.py:95, train_prob_unet_model.py:92). */
typedef struct PuAdamWChunk {
    unsigned long long p, g, m, v;
    int n;
    int pad_;
} PuAdamWChunk;
int pu_adamw_multi(const PuAdamWChunk* chunks, int nchunks, double lr, double beta1, double beta2, double eps,
                   double weight_decay, int step, void* stream);

/* ---------------- data preparation ----------------
 * pu_climex_prepare: climex2torch.__getitem__ (climex_utils.py:122-162) for a batch on the device:
 *   lr = AvgPool2d(scale)(hr); lrinterp = interpolate(lr, scale_factor=scale, mode="bilinear");
 *   inputs = stand(lrinterp); targets = stand(hr) - stand(lrinterp).            All tensors fp32 NCHW.
 * Statistics (climex_utils.py:165-195): PERPIXEL s0 = mean, s1 = std as [C][H][W]; PERTIMESTEP s0 = mean, s1 = std as
 * [N][C]; MINMAX s0 = min, s1 = max as [N][C]; NONE: null.  eps is the reference's 1e-10.
 * pu_climex_residual_to_hr: climex_utils.py:198-211, hr_pred = lrinterp + residual * (s1 [- s0] + eps). */
#define PU_STAND_NONE 0
#define PU_STAND_PERPIXEL 1
#define PU_STAND_PERTIMESTEP 2
#define PU_STAND_MINMAX 3
int pu_climex_prepare(const float* hr, const float* s0, const float* s1, int stand_mode, float eps, int N, int C, int H,
                      int W, int scale, float* lr, float* lrinterp, float* inputs, float* targets, void* stream);
int pu_climex_residual_to_hr(const float* residual, const float* lrinterp, const float* s0, const float* s1, int stand_mode,
                             float eps, int N, int C, int H, int W, float* hr_pred, void* stream);

/* ---------------- ensemble metric ----------------
 * trainmodel.crps_empirical (trainmodel.py:66-110): out = mean_s |pred_s - truth| - sum_{i<j} |pred_i - pred_j| / S^2.
 * Element (o, r), o < outer, r < inner: members at pred[o * outer_stride + s * member_stride + r], truth / out at
 * [o * inner + r].  Reference layout pred [S, ...]: outer = 1, member_stride = inner.  Ensemble [B, S, C, H, W] against
 * truth [B, C, H, W]: outer = B, inner = member_stride = C*H*W, outer_stride = S*C*H*W. */
int pu_crps_empirical(const float* pred, const float* truth, float* out, int S, long long outer, long long inner,
                      long long member_stride, long long outer_stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif
