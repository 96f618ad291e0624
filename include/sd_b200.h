/*
 * sd_b200.h — C ABI of libsd_b200.so: the sm_100a kernels behind soccerdiffusion_b200.
 *
 * The reference (bit-bots/SoccerDiffusion) has NO native/FFI layer: its hot path is Python calling
 * torch.nn and diffusers.  The drop-in boundary is therefore the Python surface of
 * soccer_diffusion/ml (mirrored by soccerdiffusion_b200/), and THIS header is what that Python
 * mirror binds with ctypes (soccerdiffusion_b200/_lib.py; INTEGRATION.md shows the stub).  Each entry
 * point names the reference code it replaces (paths relative to /root/reference/soccer_diffusion/
 * unless they start with torch/ or diffusers/).
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t, or a negative SD_ERR_* code;
 *     sd_error_string() translates either.  Nothing throws, nothing aborts.
 *   - all pointers are DEVICE pointers to fp32 unless stated; sizes in elements; `ld*` = row stride
 *     in elements; `stream` is a cudaStream_t passed as void*.
 *   - no hidden allocation or global state, except inside an explicit sd_plan (sampler state).
 *   - not thread-safe per plan; one plan per process/device (matches ml/inference/ros.py:155-163,
 *     where Inference.step is in a MutuallyExclusiveCallbackGroup).
 */
#ifndef SD_B200_H
#define SD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SD_B200_ABI_VERSION 1

/* error codes (negative; positive values are cudaError_t) */
#define SD_E_BAD_ARG (-1)
#define SD_E_UNSUPPORTED (-2)
#define SD_E_NO_PLAN (-3)

#define SD_PREC_FP32 0 /* true-fp32 FFMA: the 1e-4 mode */
#define SD_PREC_BF16 1 /* bf16 operands on tcgen05 tensor cores, fp32 accumulate in TMEM: the 2e-2 mode */

#define SD_LAYOUT_MK 0 /* A stored (M x K) row-major */
#define SD_LAYOUT_KM 1 /* A stored (K x M) row-major */
#define SD_LAYOUT_NK 0 /* B stored (N x K) row-major (torch Linear weight) */
#define SD_LAYOUT_KN 1 /* B stored (K x N) row-major */

#define SD_ACT_NONE 0
#define SD_ACT_GELU 1 /* exact erf GELU (activation="gelu" in base.py:36 / decoder.py:31) */

int sd_abi_version(void);
const char* sd_error_string(int code);

/* ---------------------------------------------------------------------------------------------
 * Fused GEMM:  C = epilogue( opA(A) * opB(B) ).
 * Replaces nn.Linear / nn.Conv1d(k=stride=patch) / the packed in_proj of nn.MultiheadAttention and
 * their autograd (encoder/base.py:28,49; decoder.py:23,36,48,54; torch/nn/functional.py:5849-5855),
 * with nn.LayerNorm folded into the operand load (norm_first layers, torch/nn/modules/
 * transformer.py:944-950, 1131-1143), and bias / erf-GELU / dropout / positional encoding
 * (misc.py:57-65) / residual add folded into the store.
 */
typedef struct sd_gemm_desc {
    const float* A; long long lda; int a_layout;
    const float* B; long long ldb; int b_layout;
    float* C; long long ldc;
    int M, N, K;
    int precision;               /* SD_PREC_* */
    /* LayerNorm-on-load of the activation operand (A when MK; B when A is KM and B is KN): */
    const float* ln_mean; const float* ln_rstd; const float* ln_gamma; const float* ln_beta;
    /* epilogue, applied in this order: */
    float alpha;                 /* 0 means 1 */
    const float* bias;           /* [N] */
    float* pre_out; long long ldp;             /* store value before activation (for backward) */
    int act;                     /* SD_ACT_* */
    const float* gelu_grad_src; long long ldg; /* multiply by gelu'(src[m][n]) (FFN backward) */
    float dropout_p; unsigned long long dropout_seed; unsigned int dropout_stream; /* element idx = m*N+n */
    const float* pe; int pe_period;            /* + pe[m % period][n]  (table is [period][N]) */
    const float* residual; long long ldr;      /* + residual[m][n] */
    int accumulate;              /* C += result instead of C = result */
} sd_gemm_desc;

int sd_gemm(const sd_gemm_desc* desc, void* stream);

/* LayerNorm statistics / backward (eps as nn.LayerNorm: 1e-5). d in {32,64,128,256,512}. */
int sd_ln_stats(const float* x, long long ld, long long M, int d, float* mean, float* rstd, float eps, void* stream);
int sd_ln_bwd(const float* g, long long ldg, const float* x, long long ldx, const float* mean, const float* rstd,
              const float* gamma, const float* dres, long long ldres, float* dx, long long lddx, float* dgamma,
              float* dbeta, long long M, int d, void* stream);

/* Attention core softmax(QK^T/sqrt(dh))V over packed head slices; rows are (b*T+t) / (b*M+m), head h
 * occupies columns [h*dh,(h+1)*dh).  Replaces F.scaled_dot_product_attention inside
 * nn.MultiheadAttention (torch/nn/functional.py:6623-6690). dh in {4,8,16,32,64,128}.
 * lse: (B,H,T) log-sum-exp saved for the backward pass (may be NULL in inference). */
int sd_attention_fwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv,
                     float* O, long long ldo, float* lse, int B, int H, int T, int M, int dh, float dropout_p,
                     unsigned long long seed, unsigned int stream_id, void* stream);
int sd_attention_bwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv,
                     const float* O, long long ldo, const float* dO, long long lddo, const float* lse, float* dQ,
                     long long lddq, float* dK, long long lddk, float* dV, long long lddv, int B, int H, int T, int M,
                     int dh, float dropout_p, unsigned long long seed, unsigned int stream_id, void* stream);

/* The same attention core on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM, fp32 softmax) for
 * T <= 128, M <= 128, dh in {16,32,64}, 16-byte aligned pointers and strides that are multiples of 4: the bf16
 * (2e-2) mode of the encoders' and the denoiser's self-attention.  sd_attention_tc_supported -> 1/0. */
int sd_attention_tc_supported(int T, int M, int dh);
int sd_attention_tc_fwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv,
                        float* O, long long ldo, float* lse, int B, int H, int T, int M, int dh, float dropout_p,
                        unsigned long long seed, unsigned int stream_id, void* stream);
int sd_attention_tc_bwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V, long long ldv,
                        const float* O, long long ldo, const float* dO, long long lddo, const float* lse, float* dQ,
                        long long lddq, float* dK, long long lddk, float* dV, long long lddv, int B, int H, int T, int M,
                        int dh, float dropout_p, unsigned long long seed, unsigned int stream_id, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Elementwise / scheduler kernels
 */
/* StepToken.forward (ml/model/misc.py:25-35): out[b] = [sin(t f) | cos(t f) | token]; t int64 or f32 */
int sd_step_token(const void* t, int t_is_float, const float* freqs, const float* token, float* out, long long ld_out,
                  int B, int d, void* stream);
int sd_step_token_bwd(const float* dout, long long ld, int B, int d, float* dtoken, void* stream);
/* Normalizer.normalize (dataset/pytorch.py:410-411) + DDIMScheduler.add_noise (train.py:204,218) */
int sd_q_sample(const float* joint_command, const float* mean, const float* std, const float* noise, const long long* t,
                const float* alphas_cumprod, int n_train, float* x0_out, float* xt_out, int B, int inner, int J,
                void* stream);
/* DDIMScheduler.step, eta=0 (ros.py:310, distill.py:189, plot.py:131) with host-side sqrt coefficients */
int sd_ddim_step(const float* x, const float* eps, float* prev, float* x0_pred, long long n, float sqrt_beta_t,
                 float sqrt_alpha_t, float sqrt_alpha_prev, float sqrt_beta_prev, void* stream);
/* F.mse_loss (train.py:229, distill.py:198) forward/backward */
int sd_mse_fwd(const float* pred, const float* target, long long n, float* loss_out, void* stream);
int sd_mse_bwd(const float* pred, const float* target, long long n, const float* grad_loss, float* grad_pred,
               void* stream);
/* Normalizer.normalize (mode 0) / denormalize (mode 1) (dataset/pytorch.py:410-414; ros.py:313) */
int sd_affine_joints(const float* x, const float* mean, const float* std, float* out, long long n, int J, int mode,
                     void* stream);
/* torch.optim.AdamW step over one flat buffer (train.py:162,239) */
int sd_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, float grad_scale, void* stream);
/* the same update with the hyper-parameters in device memory: hyper_dev = {lr, beta1, beta2, eps, weight_decay,
 * lr/(1-beta1^t), 1/sqrt(1-beta2^t), grad_scale} — capturable in a CUDA graph while OneCycleLR advances on the host.
 * step_dev != NULL: t is read from device memory (a counter the captured graph itself advances) and the two bias
 * corrections are computed in the kernel; hyper_dev[5..6] are then ignored. */
int sd_adamw_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper_dev, const int* step_dev,
                      void* stream);
/* device-resident counter added to every dropout seed (NULL = off): advancing it between CUDA-graph replays gives
 * each replay fresh masks although the host-side seeds are frozen in the graph */
int sd_set_dropout_seed_offset(const unsigned long long* device_counter);
/* GameStateEncoder (ml/model/encoder/game_state.py:27) gather + its gradient */
int sd_gather_rows(const float* table, const long long* idx, int rows, float* out, long long ld_out, int B, int d,
                   int* err_flag, void* stream);
int sd_scatter_add_rows(const float* dout, long long ld, const long long* idx, int rows, float* dtable, int B, int d,
                        void* stream);
/* bias gradients: out[n] += sum_m x[m][n] */
int sd_colsum_accum(const float* x, long long ld, long long M, int N, float* out, void* stream);
/* torch.cat(context + [step_token], dim=1) (ml/model/model.py:176) as strided copies; and its gradient */
int sd_copy_rows(const float* src, long long src_batch_stride, long long src_ld, float* dst, long long dst_batch_stride,
                 long long dst_ld, int B, int rows, int cols, int accumulate, void* stream);
int sd_add(const float* a, const float* b, float* y, long long n, void* stream);
/* the exact keep/scale factors the fused kernels apply for dropout stream `stream_id` */
int sd_dropout_mask(float* out, long long n, float p, unsigned long long seed, unsigned int stream_id, void* stream);
/* y = x * (that factor): re-applies a forward dropout mask in the backward pass */
int sd_dropout_apply(const float* x, float* y, long long n, float p, unsigned long long seed, unsigned int stream_id,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * Image-trunk non-convolution layers, bf16 NHWC (x[R][C], R = N*H*W, C a multiple of 8 dividing 2048).
 * Replace nn.BatchNorm2d (train: batch statistics + running-stat update; eval: pass running stats as
 * mean/invstd) fused with ReLU and the residual add of torchvision's BasicBlock/Bottleneck, and the stem's
 * MaxPool2d(3, 2, 1) — ml/model/encoder/image.py:46-52 (convolutions remain cuDNN calls).
 * `sums` is a caller-provided scratch of 2*C doubles. */
int sd_bn_stats_nhwc_bf16(const void* x, long long R, int C, double* sums, float eps, float momentum, float* mean,
                          float* invstd, float* running_mean, float* running_var, void* stream);
/* relu_mask (optional, R*C/8 bytes): bit i of byte v = ReLU passed channel 8*(v % (C/8)) + i of row v / (C/8) */
int sd_bn_apply_nhwc_bf16(const void* x, const void* residual, const float* mean, const float* invstd, const float* gamma,
                          const float* beta, int relu, void* y, void* relu_mask, long long R, int C, void* stream);
/* dy -> (dx, dresidual = dy*relu_mask, dgamma, dbeta); relu_mask = the bytes written by the forward, or NULL;
 * beta_recompute (used when relu_mask is NULL): non-NULL = a ReLU followed the BatchNorm (no residual) and its mask
 * is recomputed as (x - mean) * invstd * gamma + beta > 0 */
int sd_bn_bwd_nhwc_bf16(const void* dy, const void* relu_mask, const void* x, const float* mean, const float* invstd,
                        const float* gamma, const float* beta_recompute, double* sums, void* dx, void* dres, float* dgamma, float* dbeta, long long R,
                        int C, void* stream);
/* Same with TWO incoming gradients (dy2 may be NULL): when the BatchNorm output feeds two consumers (a residual block's
 * output goes to the next block's convolution AND its identity path) the two gradients are summed in fp32 inside the
 * reduction and apply kernels instead of by a separate elementwise-add pass over the tensor. */
int sd_bn_bwd2_nhwc_bf16(const void* dy, const void* dy2, const void* relu_mask, const void* x, const float* mean,
                         const float* invstd, const float* gamma, const float* beta_recompute, double* sums, void* dx,
                         void* dres, float* dgamma, float* dbeta, long long R, int C, void* stream);
/* fp32 NCHW images (N,3,H,W), H and W even -> bf16 NHWC (N,(H+6)/2,(W+6)/2,16): 3-pixel zero padding + 2x2
 * space-to-depth (channel = c*4 + dy*2 + dx, 12 used), the layout in which the 7x7/s2/p3 stem convolution
 * (torchvision resnet conv1; ml/model/encoder/image.py:55-73) is a 4x4/s1 convolution with Cin=16 */
int sd_stem_pack_s2d_bf16(const float* images, void* out, int N, int H, int W, void* stream);
/* Same packing from the RAW uint8 image (N,3,H,W): the reference's host-side torchvision preprocessing
 * v2.ToDtype(float32, scale=True) -> v2.Normalize(mean, std) (dataset/pytorch.py:198-204, ml/inference/ros.py:190-196) is
 * applied on the device in the same fp32 operations, (u * fp32(1/255) - mean[c]) / std[c], so the result is bit-identical to
 * packing the host-normalised float image; the host->device copy is 4x smaller. */
int sd_stem_pack_s2d_u8(const void* images_u8, void* out, int N, int H, int W, float mean0, float mean1, float mean2,
                        float std0, float std1, float std2, void* stream);
/* Weight gradient of the stem convolution in its space-to-depth form: dw_s2d[256][64] fp32 (row = kh*64 + kw*16 + ci,
 * column = output channel) from the packed image (sd_stem_pack_s2d_bf16) and dy bf16 NHWC (N,H/2,W/2,64); tcgen05,
 * both operands MN-major.  Replaces cuDNN's conv1 wgrad (ml/model/encoder/image.py:55-73 under autograd). */
int sd_stem_wgrad_s2d_bf16(const void* xs2d, const void* dy, float* dw_s2d, int N, int H, int W, void* stream);
/* Forward of the stem convolution on the packed image: y (N,H/2,W/2,64) bf16 NHWC; w_s2d[64][256] bf16 (row = output
 * channel, column = kh*64 + kw*16 + ci).  TMA tensor map -> tcgen05, weights resident in shared memory, double-buffered
 * TMEM accumulators.  Returns SD_E_UNSUPPORTED when W/2 is not a multiple of 8 in [8,128] or the driver has no tensor-map
 * encoder (callers then use the library convolution). */
int sd_stem_fprop_s2d_bf16(const void* xs2d, const void* w_s2d, void* y, int N, int H, int W, void* stream);
/* Same, plus the per-channel sums of y and y*y over all pixels (fp32 accumulator values) into sums[2][64] (zeroed by the
 * call): bn1's batch statistics come out of the convolution's epilogue instead of another pass over the 112x112 map;
 * sd_bn_finalize turns them into mean / invstd / running statistics (nn.BatchNorm2d training semantics). */
int sd_stem_fprop_s2d_bf16_stats(const void* xs2d, const void* w_s2d, void* y, double* sums, int N, int H, int W,
                                 void* stream);
int sd_bn_finalize(const double* sums, long long R, int C, float eps, float momentum, float* mean, float* invstd,
                   float* running_mean, float* running_var, void* stream);
/* Fused stem: maxpool3x3s2(relu(bn(x))) without materialising the activated 112x112 map; backward recomputes the
 * ReLU mask from x (torchvision ResNet stem bn1 -> relu -> maxpool). mean/invstd from sd_bn_stats_nhwc_bf16.
 * With C = 64 and W <= 112 (sd_stem_band_supported) the specialised kernels of csrc/stem_band.cu run: the forward stages
 * its input rows with one TMA bulk copy per CTA and takes the 9-tap arg-max on packed bf16 pairs; the backward routes the
 * pooled gradient of the four windows reaching a 2x2 pixel block into registers (see also _bwd2 for the pooled-domain
 * reductions); other shapes use generic gather kernels. */
int sd_stem_band_supported(int H, int W, int C);
int sd_stem_bn_relu_pool_nhwc_bf16_fwd(const void* x, const float* mean, const float* invstd, const float* gamma,
                                       const float* beta, void* y, void* idx, int N, int H, int W, int C, void* stream);
int sd_stem_bn_relu_pool_nhwc_bf16_bwd(const void* dpool, const void* idx, const void* x, const float* mean,
                                       const float* invstd, const float* gamma, const float* beta, double* sums, void* dx,
                                       float* dgamma, float* dbeta, int N, int H, int W, int C, void* stream);
/* Same, given also the pooled forward OUTPUT y_pooled (N,HO,WO,C) bf16 (may be NULL): the per-channel reductions are then
 * taken in the pooled domain — sum_windows dp*(y>0) and sum_windows dp*(y>0)*(y-beta)/gamma, an exact identity because
 * only arg-max pixels receive gradient — reading 2 GB instead of 5.6 GB at bs=256.  Falls back to the exact pixel-domain
 * kernel on the device when some channel has |beta/gamma| > 16 or gamma ~ 0 (y is bf16-rounded). */
int sd_stem_bn_relu_pool_nhwc_bf16_bwd2(const void* dpool, const void* idx, const void* x, const void* y_pooled,
                                        const float* mean, const float* invstd, const float* gamma, const float* beta,
                                        double* sums, void* dx, float* dgamma, float* dbeta, int N, int H, int W, int C,
                                        void* stream);
/* idx: one byte per output element (arg-max tap 0..8) */
int sd_maxpool3x3s2_nhwc_bf16_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, void* stream);
int sd_maxpool3x3s2_nhwc_bf16_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Persistent DDIM sampler (ros.py:301-310, distill.py:179-189): one launch = all steps.
 */
typedef struct sd_plan sd_plan;

typedef struct sd_plan_config {
    int d;          /* hidden_dim */
    int heads;      /* 4 (model.py:115) */
    int layers;     /* num_decoder_layers */
    int T;          /* trajectory_prediction_length (<= 32) */
    int J;          /* num_joints */
    int ctx_tokens; /* context tokens WITHOUT the step token (311 at default.yaml) */
} sd_plan_config;

/* reference-format (state_dict) tensors of one nn.TransformerDecoderLayer */
typedef struct sd_decoder_layer_weights {
    const float *sa_in_w, *sa_in_b, *sa_out_w, *sa_out_b;   /* self_attn.{in_proj_weight (3d,d), in_proj_bias, out_proj.*} */
    const float *ca_in_w, *ca_in_b, *ca_out_w, *ca_out_b;   /* multihead_attn.* */
    const float *lin1_w, *lin1_b, *lin2_w, *lin2_b;
    const float *norm1_w, *norm1_b, *norm2_w, *norm2_b, *norm3_w, *norm3_b;
} sd_decoder_layer_weights;

int sd_plan_create(const sd_plan_config* cfg, sd_plan** out);
int sd_plan_destroy(sd_plan* plan);
/* (re)pack weights after load_state_dict / optimizer steps */
int sd_plan_set_layer(sd_plan* plan, int layer, const sd_decoder_layer_weights* w, void* stream);
int sd_plan_set_io(sd_plan* plan, const float* emb_w, const float* emb_b, const float* fc_w, const float* fc_b,
                   const float* pe, const float* step_freqs, const float* step_token, const float* mean,
                   const float* std, void* stream);
/* DDIMScheduler.set_timesteps: HOST arrays timesteps[num_steps], coef[num_steps][4] =
 * {sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)}; tabulates the step-token K/V rows */
int sd_plan_set_schedule(sd_plan* plan, int num_steps, const long long* timesteps_host, const float* coef_host,
                         void* stream);
/* model.encode_input_data output, concatenated (B, ctx_tokens, d): projects K/V for all layers once */
int sd_plan_set_context(sd_plan* plan, const float* ctx, int B, void* stream);
/* x_T (B,T,J) -> x_0 (B,T,J); eps_trace optional (steps,B,T,J); denormalize: x*std+mean (ros.py:313).
 * B = trajectories held by x_T / x_out: SD_E_BAD_ARG unless it equals the batch of the last sd_plan_set_context
 * (the reference raises a shape error for mismatched context / sample batches). */
int sd_plan_sample(sd_plan* plan, const float* x_T, float* x_out, float* eps_trace, int denormalize, int B, void* stream);
/* Sampler kernel selection: 0 = auto (cluster kernel for batches of <= 18 trajectories when the device can co-schedule it),
 * 1 = one CTA per trajectory, 2 = one 16-CTA thread-block cluster per trajectory (weights resident in the
 * cluster's shared memory, activations exchanged through distributed shared memory).  Env SD_B200_SAMPLER=
 * auto|cta|cluster sets the process default.  sd_plan_last_sampler: which one the last sd_plan_sample ran (1/2). */
int sd_plan_set_sampler(sd_plan* plan, int mode);
int sd_plan_last_sampler(const sd_plan* plan);
/* profiling aid: device buffer of num_steps*96 int64 that the cluster sampler fills with clock64() stamps of its
 * phases (CTA 0); NULL disables */
int sd_plan_set_debug_stamps(sd_plan* plan, long long* device_buffer);
/* one forward_with_context (model.py:159-179) against the cached context, per-sample t */
int sd_plan_denoise(sd_plan* plan, const float* x, const void* t, int t_is_float, float* eps_out, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Layer-fused tensor-core path (bf16 mode, d_model = ff = 128: default.yaml and the scaled-up config).
 *
 * sd_pack_weights_bf16: fp32 weight matrices [rows][K] -> rows of ONE packed bf16 matrix [*][K] that the fused
 * kernels read through a TMA tensor map (one launch for all layers of a stack).  Row order per ENCODER layer:
 * in_proj (3d: Wq;Wk;Wv) | out_proj (d) | linear1 (d) | linear2 (d) = 768 rows; per DECODER layer:
 * self_attn.in_proj (3d) | self_attn.out_proj | multihead_attn.in_proj (3d) | multihead_attn.out_proj | linear1 |
 * linear2 = 1280 rows.
 */
#define SD_PACK_MAX_SEGMENTS 64
typedef struct sd_pack_args {
    const float* src[SD_PACK_MAX_SEGMENTS];
    int rows[SD_PACK_MAX_SEGMENTS];
    int dst_row0[SD_PACK_MAX_SEGMENTS];
    int K;
    void* dst; /* bf16 */
} sd_pack_args;
int sd_pack_weights_bf16(const sd_pack_args* args, int n_segments, void* stream);

/* One pre-LN encoder layer (nn.TransformerEncoderLayer(norm_first=True, activation="gelu", dim_feedforward=d),
 * encoder/base.py:29-40; torch/nn/modules/transformer.py:944-950) as ONE kernel: LN1 -> QKV -> per-head softmax
 * attention -> out-proj + dropout + residual -> LN2 -> FC1 + GELU + dropout -> FC2 + dropout + residual.
 * x, y: fp32 [B*S][128] (may alias).  One CTA per tile of floor(128/S) samples.  Dropout streams: dropout_stream + 0
 * (attention probabilities, element ((b*H+h)*S+t)*S+m), +1 (out-proj, element row*128+c), +2 (FC1), +3 (FC2).
 * Optional saves for the backward pass (all NULL in inference): x1_save fp32 [B*S][128] (residual stream after the
 * attention block), xn1/attn/xn2/hact_save bf16 [B*S][128] (LN1(x), attention output, LN2(x1), hidden activation). */
#define SD_LAYER_SA 1  /* the self-attention block  x1 = x + Drop(OutProj(MHA(LN1 x))) */
#define SD_LAYER_FFN 2 /* the feed-forward block    y = x1 + Drop(W2 Drop(GELU(W1 LN2 x1))) */
#define SD_LAYER_FFN_FIRST 4 /* with SA | FFN, inference only: the feed-forward block (w_row_ffn, l*_b, n2_*) runs BEFORE the attention
                              * block (w_row0, in_b, out_b, n1_*) - the feed-forward half of decoder layer l and the self-attention
                              * half of layer l+1 as one launch */
typedef struct sd_enc_layer_desc {
    const float* x; float* y;
    int B, S, H;
    const void* w_packed; int w_rows_total; int w_row0; /* packed bf16 weights [w_rows_total][128]; first row of Wq|Wk|Wv|Wout */
    const float *in_b, *out_b, *l1_b, *l2_b, *n1_w, *n1_b, *n2_w, *n2_b;
    float* x1_save; void* xn1_save; void* attn_save; void* xn2_save; void* hact_save;
    float dropout_p; unsigned long long dropout_seed; unsigned int dropout_stream;
    /* A DECODER layer (decoder.py:25-35; torch transformer.py:1131-1143) runs the two halves as separate launches around its
     * cross-attention block: blocks = SD_LAYER_SA (y = x1; LN1 / attention parameters only) or SD_LAYER_FFN (x is x1; n2_* /
     * l*_b are the layer's norm3 / linear parameters).  0 = both (an encoder layer).  w_row_ffn: first packed row of W1|W2
     * (0 = w_row0 + 512); dropout_stream_ffn: first dropout stream of the feed-forward block (0 = dropout_stream + 2). */
    int blocks; int w_row_ffn; unsigned int dropout_stream_ffn;
} sd_enc_layer_desc;
int sd_enc_layer_supported(int d, int ff, int S, int H);
int sd_enc_layer_fwd(const sd_enc_layer_desc* desc, void* stream);

/* Backward of the same layer, data path, as ONE kernel (autograd of the layer above): dy -> dx (fp32 [B*S][128], may
 * alias), recomputing hpre, Q, K, V and the softmax from the forward's saves x (layer input), x1, xn1, xn2.
 * Outputs for the weight-gradient GEMMs (bf16): g2 = dy*mask3, dhpre, g1 = dx1*mask1 ([B*S][128]) and dqkv ([B*S][384]:
 * dq | dk | dv).  LayerNorm weight / bias gradients are ACCUMULATED into g_n1_w, g_n1_b, g_n2_w, g_n2_b (fp32 [128]). */
typedef struct sd_enc_layer_bwd_desc {
    const float* dy; float* dx;
    const float* x; const float* x1; const void* xn1; const void* xn2;
    void* g2; void* dhpre; void* g1; void* dqkv;
    float *g_n1_w, *g_n1_b, *g_n2_w, *g_n2_b;
    int B, S, H;
    const void* w_packed; int w_rows_total; int w_row0;
    const float *in_b, *l1_b, *n1_w, *n2_w;
    float dropout_p; unsigned long long dropout_seed; unsigned int dropout_stream;
    int blocks; int w_row_ffn; unsigned int dropout_stream_ffn; /* as in sd_enc_layer_desc; SA only: x1 unused, FFN only: x unused */
} sd_enc_layer_bwd_desc;
int sd_enc_layer_bwd(const sd_enc_layer_bwd_desc* desc, void* stream);

/* Weight gradients of linear layers from bf16 activations, TMA-fed tcgen05 GEMMs with split K over the tokens:
 *   for each job j:  dW_j[n][k] += sum_t G_j[t][n] * X_j[t][k]   and (optionally)   db_j[n] += sum_t G_j[t][n]
 * G_j: bf16 [rows][ldg] (128 columns starting at column g_col0), X_j: bf16 [rows][ldx] (128 columns starting at column
 * x_col0), dW_j: fp32 [128][ldw], db_j: fp32 [128] or NULL.  All jobs share `rows`.  The bias gradient is one more MMA
 * against a tile of ones.  Replaces the autograd of nn.Linear / packed in_proj weights (torch/nn/functional.py:5849). */
#define SD_WGRAD_MAX_JOBS 8
typedef struct sd_wgrad_job {
    const void* G; long long ldg; int g_col0;
    const void* X; long long ldx; int x_col0;
    float* dW; long long ldw;
    float* db;
} sd_wgrad_job;
int sd_wgrad_bf16(const sd_wgrad_job* jobs, int n_jobs, long long rows, void* stream);

/* ---------------------------------------------------------------------------------------------
 * The decoder layer's cross-attention block, layer-fused on tcgen05 + TMA (bf16 mode; d_model = 128, 4 heads,
 * T <= 64 query tokens in groups of <= 16, memory length M <= 384).  Replaces, per decoder layer, LN2 -> q projection -> k/v projection
 * of the memory -> nn.MultiheadAttention core -> out-projection + dropout + residual
 * (torch/nn/modules/transformer.py:1137-1139; ml/model/decoder.py:25-54) and its autograd.
 *
 * sd_cast_bf16: fp32 -> bf16 copy (n a multiple of 8, 16-byte aligned).
 * sd_kv_proj_bf16: the memory's K | V projections of ALL layers in one GEMM:
 *   kv[row][256 l + n] = sum_k mem[row][k] * W[w_row0 + l*w_stride + n][k] + biases[l][n]      (n < 256: Wk rows, Wv rows)
 *   mem bf16 [rows][128]; W = the packed bf16 weight matrix of sd_pack_weights_bf16; kv bf16 [rows][ldkv].
 * sd_kv_dgrad_bf16: dmem[row][n] (+)= sum_{l,k} dkv[row][256 l + k] * W[w_row0 + l*w_stride + k][n]   (fp32 [rows][lddmem]).
 * (The k/v weight gradients are sd_wgrad_bf16 jobs with G = dkv, X = the bf16 memory.) */
/* Programmatic dependent launch for chains of short kernels (the tensor-core sampler's 16 launches per DDIM step):
 * on != 0 -> sd_gemm (bf16 mode), sd_enc_layer_fwd, sd_ca_block_fwd, sd_bcast_row_bf16 and sd_ddim_step are launched with
 * cudaLaunchAttributeProgrammaticStreamSerialization; each of them runs its prologue under the preceding kernel's tail and
 * executes griddepcontrol.wait before touching activations.  Returns the previous setting. */
int sd_set_pdl(int on);
/* Hardware probe used by tests / tools only: D[128][64*nblk_b] = A B^T over 16*ksteps k rows, where A and B are MN-major
 * SWIZZLE_128B views of TMA-loaded bf16 [rows][64] tiles X and Y starting at row shift_a / shift_b (any row, not only
 * multiples of 8), A's two and B's nblk_b 64-wide MN blocks lbo bytes apart (blocks may overlap).  cycles: clock64 ticks of
 * `reps` repetitions of the MMA sequence (or NULL). */
int sd_debug_shifted_mma(const void* X, const void* Y, int rows, int shift_a, int lbo_a, int shift_b, int lbo_b, int nblk_b,
                         int ksteps, int use_base_offset, int reps, float* D, long long* cycles, void* stream);
/* Hardware probe used by tools only: per-warp clock64 ticks (cycles[nwarps]) of `count` TMEM loads run by nwarps <= 16 warps of
 * one CTA at once (warp w reads lane quadrant w % 4); mode 0: tcgen05.ld.32x32b.x32 + wait each, 1: two such loads in flight,
 * 2: .x16 loads + wait each.  (use_base_offset bit 2 of sd_debug_shifted_mma runs shift_a loads per warp beside the MMA chain.) */
int sd_debug_ldtm(int nwarps, int count, int mode, long long* cycles, float* sink, void* stream);
/* One launch between two denoiser evaluations of the batched DDIM loop (ros.py:301-310; decoder.py:48,54): output
 * projection eps = h fc_w^T + fc_b (h fp32 [rows][128], fc_w [J][128]), the eta=0 update x_next = sap*(x - sb*eps)/sa + sbp*eps
 * (optionally eps_out), the next step's embedding h_next = x_next emb_w^T + emb_b + pe[row % T] (emb_w [128][J]; NULL on
 * the last step; h_next may alias h), and the broadcast of the next step token's K | V row (src_row NULL: none; arguments as
 * sd_bcast_row_bf16).  J <= 32; fp32 arithmetic. */
int sd_ddim_glue(const float* h, const float* fc_w, const float* fc_b, const float* x, float* x_next, float* eps_out, int rows,
                 int J, float sqrt_beta_t, float sqrt_alpha_t, float sqrt_alpha_prev, float sqrt_beta_prev, const float* emb_w,
                 const float* emb_b, const float* pe, int T, float* h_next, void* kv, long long ldkv, long long block_rows,
                 long long row, int B, const void* src_row, int ncols, void* stream);
#define SD_KV_MAX_LAYERS 16
int sd_cast_bf16(const float* src, void* dst_bf16, long long n, void* stream);
/* dst[(b*block_rows + row)*ld + c] = src_row[c] for b < B, c < ncols (bf16; the step token's K | V row of one DDIM step
 * written into every trajectory's cached K | V, ros.py:301-310 calls the model with the same t for the whole batch) */
int sd_bcast_row_bf16(void* dst, long long ld, long long block_rows, long long row, int B, const void* src_row, int ncols,
                      void* stream);
int sd_kv_proj_bf16(const void* mem_bf16, long long rows, const void* w_packed, int w_rows_total, int w_row0, int w_stride,
                    int n_layers, const float* const* biases, void* kv_out, long long ldkv, void* stream);
int sd_kv_dgrad_bf16(const void* dkv_bf16, long long rows, long long lddkv, const void* w_packed, int w_rows_total, int w_row0,
                     int w_stride, int n_layers, float* dmem, long long lddmem, int accumulate, void* stream);

/* Data gradient of a 1x1, stride-2, padding-0 convolution (the downsample path of a ResNet stage, torchvision resnet.py
 * `downsample = conv1x1(inplanes, planes, stride)`; ml/model/encoder/image.py:55-73): dx[n][2ho][2wo][ci] = sum_co dy[n][ho][wo][co]
 * * w[co][ci], zero at the pixels the stride skips.  dy bf16 NHWC [frames][Hin/2][Win/2][Cout], w bf16 [Cout][Cin], dx bf16 NHWC
 * [frames][Hin][Win][Cin]; Hin, Win even, Cin in {64,128,256}, Cout a multiple of 64.  TMA-fed tcgen05 GEMM. */
int sd_conv1x1s2_dgrad_supported(int Hin, int Win, int Cin, int Cout);
int sd_conv1x1s2_dgrad_bf16(const void* dy, const void* w_bf16, void* dx, int frames, int Hin, int Win, int Cin, int Cout,
                            void* stream);

/* Weight gradient of a 3x3, stride-1, padding-1 convolution with 64 input and 64 output channels (ResNet18 layer1;
 * torchvision resnet.py BasicBlock, ml/model/encoder/image.py:55-73) as an implicit GEMM over the pixels on tcgen05 + TMA:
 * dW[co][ci][kh][kw] (+)= sum_{n,h,w} dy[n][h][w][co] * x[n][h+kh-1][w+kw-1][ci].  x, dy: bf16 NHWC [frames][H][W][64];
 * dW: fp32 [64][64][3][3] contiguous; scratch: sd_conv3x3_wgrad_c64_scratch_bytes() bytes of device memory (per-CTA partial
 * sums, added in a fixed order: the result is deterministic). */
int sd_conv3x3_wgrad_c64_supported(int H, int W);
int sd_conv3x3_wgrad_c64_scratch_bytes(void);
int sd_conv3x3_wgrad_c64_bf16(const void* x, const void* dy, float* dW, int frames, int H, int W, float* scratch, int accumulate,
                              void* stream);

/* y = x + Drop(OutProj(MHA(LN(x), kv))) for B samples of T query rows; one CTA per sample.  x, y fp32 [B*T][128] (may
 * alias).  kv: bf16 [B*M][ldkv], this layer's K at columns [kv_col0, kv_col0+128), V at [kv_col0+128, kv_col0+256).
 * Dropout streams: dropout_stream + 0 (attention probabilities, element ((b*4+h)*T+t)*M+m), + 1 (out-proj, row*128+c).
 * Optional saves for the backward pass (NULL in inference): xn (LN output), q (projected queries incl. bias), attn
 * (attention output) as bf16 [B*T][128]; stats fp32 [B*T][2] (mean, rstd); lse fp32 [B][4][T] (log2 domain). */
typedef struct sd_ca_block_desc {
    const float* x; float* y;
    int B, T, M;
    const void* w_packed; int w_rows_total; int w_row_q; int w_row_o;
    const void* kv; long long ldkv; int kv_col0;
    const float *q_b, *out_b, *n_w, *n_b;
    void* xn_save; void* q_save; void* attn_save; float* stats_save; float* lse_save;
    float dropout_p; unsigned long long dropout_seed; unsigned int dropout_stream;
} sd_ca_block_desc;
int sd_ca_block_supported(int d, int H, int T, int M);     /* forward + backward: d = 128, 4 heads, T <= 64 (query rows in groups of <= 16), M <= 384 */
int sd_ca_block_fwd_supported(int d, int H, int T, int M); /* the same set (kept for callers that only run the forward) */
int sd_ca_block_fwd(const sd_ca_block_desc* desc, void* stream);

/* Backward of the block (data path): dy -> dx (fp32 [B*T][128], may alias), g1 = dy*mask and dq (bf16 [B*T][128], the G
 * operands of the out-proj / q-proj weight gradients), dkv (bf16 [B*M][lddkv], this layer's dK | dV at the kv columns);
 * LayerNorm weight / bias gradients are ACCUMULATED into g_n_w / g_n_b (fp32 [128]). */
typedef struct sd_ca_block_bwd_desc {
    const float* dy; float* dx;
    const float* x; const void* q; const void* attn; const float* stats; const float* lse;
    int B, T, M;
    const void* w_packed; int w_rows_total; int w_row_q; int w_row_o;
    const void* kv; long long ldkv; int kv_col0;
    const float* n_w;
    void* g1; void* dq; void* dkv; long long lddkv;
    float *g_n_w, *g_n_b;
    float dropout_p; unsigned long long dropout_seed; unsigned int dropout_stream;
} sd_ca_block_bwd_desc;
int sd_ca_block_bwd(const sd_ca_block_bwd_desc* desc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SD_B200_H */
