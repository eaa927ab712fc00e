/* unpp.h — C ABI of libunpp.so, the sm_100a kernel library behind the UNet_Nested (UNet++) drop-in.
 *
 * Boundary contract: plain pointers and sizes only, no torch types.  The caller
 * (PyTorch) owns ALL device memory — activations, workspaces, packed weights and the flat gradient
 * buffer are torch tensors whose data_ptr() is passed in.  Every entry point enqueues work on the
 * cudaStream_t it is given and never synchronises; it returns 0 on success and a negative code on
 * failure, with a thread-local message available from unpp_last_error().  The library keeps no
 * global mutable state apart from per-device immutable caches (SM count, driver entry points).
 *
 * What each entry point replaces in the reference (file:line into the reference repository):
 *   unpp_conv_tc        nn.Conv2d 3x3 + [BatchNorm2d eval-folded] + ReLU of unetConv2 (models/unet.py:132-134,140-141),
 *                       torch.cat of unetUp.forward (unet.py:199-201) as a K-loop over source tensors,
 *                       nn.ConvTranspose2d k2 s2 of unetUp (unet.py:187) in pointwise+scatter mode,
 *                       the 1x1 heads + sigmoid (unet.py:242-244,283-286) fused into the epilogue,
 *                       and — with flipped packed weights — the dgrad of all of the above.
 *   unpp_pack_weights   (no reference counterpart: OIHW fp32 state_dict -> bf16 UMMA operand layout)
 *   unpp_nchw_to_nhwc   layout change at the API edge (reference is NCHW fp32 throughout)
 *   unpp_maxpool2x2     nn.MaxPool2d(2) (unet.py:219,258,260,262)
 *   unpp_argmax_peaks   the arg-max core of Heatmap.extract_points_ (tools/misc/heatmap.py:173-178)
 *   unpp_wgrad, unpp_bn_*, unpp_maxpool2x2_bwd, unpp_head_bwd   the autograd backward of unet.py:255-300
 *   unpp_adamw          tools/optimizers/adamw.py:38-100 (AdamW.step) over one flat buffer
 *   unpp_create_heatmap tools/misc/helper.py:87-172 (target synthesis the trainer runs on the CPU every step)
 *   unpp_bilinear_up2x(_bwd)  nn.UpsamplingBilinear2d(scale_factor=2) of unetUp with is_deconv=False (models/unet.py:189-191)
 *   unpp_optim_step     every optimizer trainer/trainer.py:344-376 can select, over one flat buffer: tools/optimizers/adamw.py,
 *                       tools/optimizers/sgdw.py, tools/optimizers/adabound.py, torch.optim.SGD, torch.optim.Adam
 */
#ifndef UNPP_H_
#define UNPP_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* unpp_stream_t; /* cudaStream_t */

enum {
  UNPP_OK = 0,
  UNPP_ERR_BAD_ARG = -1,
  UNPP_ERR_CUDA = -2,
  UNPP_ERR_UNSUPPORTED = -3,
};

#define UNPP_MAX_SRC 6

/* conv modes */
#define UNPP_MODE_CONV 0    /* out[N,H,W,n_total]                                             */
#define UNPP_MODE_DECONV 1  /* out[N,2H,2W,n_total/4]; GEMM column = (2p+q)*Cout + co          */

typedef struct UnppConvArgs {
  int32_t N, H, W;              /* pixel grid of the sources (all sources share it)             */
  int32_t nsrc;                 /* number of concatenated source tensors (K loop walks them)    */
  const void* src[UNPP_MAX_SRC];/* NHWC bf16, C = src_C[i] in {16,32,64,128}                     */
  int32_t src_C[UNPP_MAX_SRC];
  /* src_step[i] in {0,1}: dense [N,H,W,C].  2: the tensor is [N,2H,2W,C] and pixel (y,x) of the grid
   * reads (2y+src_oy[i], 2x+src_ox[i]) — the four taps of the k2s2 transposed-conv dgrad.        */
  int32_t src_step[UNPP_MAX_SRC], src_oy[UNPP_MAX_SRC], src_ox[UNPP_MAX_SRC];
  int32_t taps;                 /* 9 = 3x3 stride 1 pad 1;  1 = pointwise                        */
  int32_t n_total;              /* GEMM N: Cout (conv) or 4*Cout (deconv)                        */
  int32_t n_tile;               /* columns per CTA; divides n_total; multiple of 16; <= 256      */
  const void* wpacked;          /* bf16 [n_total/n_tile][taps][K/8][n_tile][8]  (unpp_pack_weights) */
  const float* bias;            /* fp32 [Cout] or NULL                                           */
  int32_t mode;                 /* UNPP_MODE_*                                                   */
  int32_t relu;                 /* apply max(.,0) after bias                                     */
  void* out;                    /* NHWC bf16 or NULL (head-only)                                 */
  /* fused 1x1 head + sigmoid: requires mode conv, n_total == n_tile == 16 */
  const float* head_w;          /* fp32 [classes][16] or NULL                                    */
  const float* head_b;          /* fp32 [classes]                                                */
  float* heat;                  /* fp32 NCHW [N,classes,H,W]                                     */
  float* logit;                 /* optional fp32 NCHW pre-sigmoid output (training) or NULL       */
  int32_t head_classes;         /* <= 8                                                          */
  /* head dropout (training): keep-mask (below) or NULL; scale = 1/(1-p) */
  const uint16_t* drop_mask;    /* keep-mask of the head dropout, ONE 16-bit word per pixel [N,H,W]: bit c = keep channel c */
  float drop_scale;
  /* backward-only epilogue inputs, both NHWC bf16 shaped like `out` (conv mode) */
  const void* addend;           /* added to the accumulator before masking, or NULL              */
  const void* relu_mask_src;    /* forward activation y: result is zeroed where y <= 0, or NULL   */
  /* per-channel reductions of the fp32 result over all pixels (BN batch stats / bias grads):
   * stats[0..Cout) += sum(v), stats[Cout..2Cout) += sum(v*v) or sum(v*aux) ; fp32 atomics into a
   * [gridDim][2][Cout] partial buffer, reduced deterministically by unpp_reduce_partials. */
  float* stats_partial;         /* fp32 [unpp_conv_grid()][2][Cout] or NULL                       */
  const void* stats_aux;        /* NHWC bf16 like out, or NULL (then second stat = sum v*v)       */
  const float* aux_mean;        /* with stats_aux: second stat = sum v * (aux - mean[c]) * istd[c]*/
  const float* aux_istd;
  /* 2x2 output blocking (3x3 conv, every source C = 16, n_total = 16, even H and W): one GEMM row is a
   * 2x2 pixel block, N = 4*16, K walks the 4x4 input window -> 16 instead of 36 MMAs per 512 pixels.
   * wpacked must then be packed with kind 4 (forward) or 5 (dgrad), taps = 16, n_total = n_tile = 64.
   * block2x2 = 2: the same for the network's FIRST layer (inference epilogue): the one source is NHWC bf16 with 4 channels
   * (8-byte pixels, unpp_nchw_to_nhwc with Cpad = 4), wpacked is kind 7; a K = 16 step is half a window row of 4-channel
   * pixels, read through overlapping unswizzled K-major descriptors: 8 MMAs per 512 output pixels, 4x fewer input bytes. */
  int32_t block2x2;
  /* Fused transposed conv (inference, with block2x2): conv3x3(cat[ConvTranspose2d_k2s2(low), src...]) — the k2s2
   * upsample never overlaps, so its branch collapses into a 3x3 conv over the LOW-resolution tensor with
   * composed weights whose 64 GEMM columns are (pixel of the 2x2 block, co): one more K chunk of the same
   * accumulators.  lowres_src: NHWC bf16 [N, H/2, W/2, lowres_C] (lowres_C = 32); lowres_wpacked: kind-6 packing
   * of the composed [64][lowres_C][3][3] weight (taps 9, n_total = n_tile = 64).  The upsample bias reaches
   * fewer taps at the image border: bias is then a [9][16] table indexed by (row class, column class). */
  const void* lowres_src;
  const void* lowres_wpacked;
  int32_t lowres_C;
  int32_t bias_classes;         /* 0/1: bias[Cout]; 9: bias[3*rowclass+colclass][16], class 0 first, 1 interior, 2 last */
  /* Fused nn.MaxPool2d(2) of the output (models/unet.py:219,258,260,262): NHWC bf16 [N, H/2, W/2, n_total], written next to
   * `out` by the inference epilogue (conv mode, relu = 1, even H and W, no head, no training operand).  NULL = no pooling. */
  void* pooled;
} UnppConvArgs;

const char* unpp_last_error(void);
int unpp_version(void);
int unpp_num_sms(void);

/* Fused implicit-GEMM convolution on tcgen05 tensor cores (TMA-staged NHWC bf16 halo tiles). */
int unpp_conv_tc(const UnppConvArgs* a, unpp_stream_t stream);
/* Number of CTAs unpp_conv_tc will launch for these args (size of stats_partial's leading dim).  The tiling depends
 * on which epilogue variant runs, so query with the SAME args as the launch (in particular a non-NULL stats_partial). */
int unpp_conv_grid(const UnppConvArgs* a);

/* Weight packing into the UMMA B-operand layout [n_total/n_tile][taps][K/8][n_tile][8] (bf16).
 *  kind 0: forward conv   B[n=co][tap][k=ci]   = W[co][k_begin+ci][tap] * (scale ? scale[co] : 1)
 *          src = OIHW fp32 [Cout][Cin_total][kh][kw]; K = k_count channels starting at k_begin
 *          (channels >= Cin_total pack as zero: the 3-channel first layer reads a 16-channel padded input)
 *  kind 1: dgrad of conv  B[n=ci][tap][k=co]   = W[co][n_begin+ci][taps-1-tap]  (flipped taps)
 *          n_total = number of input channels of the slice, K = Cout
 *  kind 2: deconv forward B[n=(2p+q)*Cout+co][0][k=ci] = Wd[ci][co][p][q]; src [Cin][Cout][2][2]
 *  kind 3: deconv dgrad   (one pointwise GEMM per (p,q) tap over the up-resolution gradient)
 *          B[n=ci][tap=2p+q][k=co] = Wd[ci][co][p][q]
 *  kind 4: 2x2-blocked forward conv (taps = 16 window positions dy*4+dx, n = (2*jy+jx)*16 + co):
 *          B[n][pos][k=ci] = W[co][k_begin+ci][dy-jy][dx-jx] * scale[co] if both offsets are in 0..2, else 0
 *  kind 5: 2x2-blocked dgrad: B[n=(2*jy+jx)*16+ci][pos][k=co] = W[k_begin+co][n_begin+ci][2-(dy-jy)][2-(dx-jx)] or 0
 *  kind 6: fused transposed conv of the 2x2-blocked path (taps = 9 low-resolution taps): src = composed weights
 *          [4*16][Cin][3][3] whose output channel is (2*jy+jx)*16 + co
 *          Kinds 4-6 store only the NON-ZERO (position, pixel) blocks: per 16 input channels 36 (kind 6: 16) blocks of
 *          512 B = [2 k8][16 columns][8 channels], runs of adjacent pixels contiguous (csrc/b2_blocks.h): the kernel
 *          issues one MMA (N = 16/32/64) per run.  Buffer size: (k8_total / 2) * 36 * 512 B (kind 6: * 16 * 512 B).
 *  kind 7: first layer of the 2x2-blocked path on 4-channel (8-byte) pixels (UnppConvArgs.block2x2 = 2): src [16][Cin <= 4][3][3],
 *          taps = 4 window rows, n_total = n_tile = 64, k_count = 32: per window row dy the K axis is 4 chunks of 8 = (pixel pair,
 *          4 channels) covering window columns -1 .. 6 (zero outside the 4x4 window): [dy][chunk][64 columns][8], 16 KB.
 * k_dst8 places the K range at 8-channel chunk offset k_dst8 inside a K/8 = k8_total wide buffer,
 * so that several sources (concat) or several consumers (dgrad gather) share one packed tensor. */
typedef struct UnppPackArgs {
  const float* src;
  void* dst;            /* bf16 */
  const float* scale;   /* per output channel (BN fold), kind 0 only, or NULL */
  int32_t kind;
  int32_t src_O, src_I; /* dims 0 and 1 of the source weight tensor */
  int32_t taps;
  int32_t n_total, n_tile;
  int32_t n_begin;      /* kind 1: first input channel of the slice */
  int32_t k_begin, k_count;
  int32_t k8_total, k_dst8;
} UnppPackArgs;
int unpp_pack_weights(const UnppPackArgs* a, unpp_stream_t stream);
/* The same for a whole table of n jobs resident in DEVICE memory (entries are not validated): one launch
 * re-packs every weight of a training step after the optimizer update. */
int unpp_pack_weights_batched(const UnppPackArgs* table_dev, int n, unpp_stream_t stream);

/* fp32 NCHW [N,C,H,W] -> bf16 NHWC [N,H,W,Cpad] (channels >= C zero-filled); Cpad = 16, or 4 for the first-layer mode. */
int unpp_nchw_to_nhwc(const float* x, void* out, int N, int C, int H, int W, int Cpad, unpp_stream_t stream);
/* Eval-mode BatchNorm2d (models/unet.py:133) folded into the conv in front of it: scale[c] = gamma[c] / sqrt(running_var[c] + eps)
 * (unpp_pack_weights multiplies it into the weights), bias[c] = (conv_bias[c] - running_mean[c]) * scale[c] + beta[c]. */
int unpp_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var, const float* conv_bias, float eps, int C,
                 float* scale, float* bias, unpp_stream_t stream);
/* Weight preparation of the fused transposed conv (UnppConvArgs.lowres_src): conv3x3(cat[ConvTranspose2d_k2s2(low), ...]) restricted to
 * the upsampled slice is ONE 3x3 conv over the low-resolution tensor.  w_conv fp32 [Co][Ctot][3][3] (its first Cu input channels
 * multiply the upsampled tensor, models/unet.py:199-201), w_up fp32 [Ci][Cu][2][2], b_up [Cu] (unet.py:187), b_conv [Co] ->
 * comp fp32 [4*Co][Ci][3][3] (output channel (2*jy+jx)*Co + co, taps = low-resolution offsets -1..1; pack it with kind 6) and
 * table fp32 [9][Co] = bias per (row class, column class) (UnppConvArgs.bias_classes = 9). */
int unpp_compose_deconv_conv(const float* w_conv, int Ctot, int Cu, const float* w_up, const float* b_up, const float* b_conv, int Co, int Ci,
                             float* comp, float* table, unpp_stream_t stream);
/* 8-bit images -> bf16 NHWC [N,H,W,Cpad] with torchvision ToTensor's scaling fused in (value = byte / 255 in fp32, then the bf16
 * rounding of the fp32 path): the reference converts PIL images on the CPU (datasets/datasets_base.py:71-72) and copies fp32
 * tensors to the device (trainer/trainer.py:109); copying the bytes is 4x less host-to-device traffic.
 * hwc = 1: x is [N,H,W,C] (PIL / OpenCV layout); hwc = 0: x is [N,C,H,W]. */
int unpp_u8_to_nhwc(const uint8_t* x, void* out, int N, int C, int H, int W, int Cpad, int hwc, unpp_stream_t stream);
/* bf16 NHWC 2x2/2 max pooling (H, W even). */
int unpp_maxpool2x2(const void* x, void* out, int N, int H, int W, int C, unpp_stream_t stream);
/* Per-plane arg-max of fp32 NCHW heatmaps: first maximum in row-major order -> xy[b][c] = {x, y},
 * val[b][c] = the maximum (val may be NULL). One CTA per plane, warp-shuffle reduction. */
int unpp_argmax_peaks(const float* heat, int planes, int H, int W, int32_t* xy, float* val, unpp_stream_t stream);
/* The same result (bit-identical: (value, first index) is a total order) for few, large planes — 1024x1024 at batch 16 is 64
 * planes of 4 MB: every plane is cut into `splits` segments, one CTA each, and a second launch folds the per-segment pairs.
 * workspace: planes * splits * 8 bytes of device memory owned by the caller; unpp_argmax_splits() proposes the split count
 * (1 = one CTA per plane already fills the GPU; then the call is unpp_argmax_peaks and workspace may be NULL). */
int unpp_argmax_peaks_split(const float* heat, int planes, int H, int W, int32_t* xy, float* val, void* workspace, int splits, unpp_stream_t stream);
int unpp_argmax_splits(int planes, int H, int W);
/* The multi-point form of Heatmap.extract_points_ (tools/misc/heatmap.py:148-208: threshold, regions, every region's maximum,
 * brightest `num` first, one retry at 0.9 x threshold): the `num` best strict local maxima (8-neighbourhood, order = value
 * descending then index ascending) of heat >= threshold per plane.  xy int32 [planes][num][2] as {x, y} (-1, -1 when a plane has
 * fewer peaks), val fp32 [planes][num], count int32 [planes].  Equal to the reference's point set on separated blobs; the
 * OpenCV watershed itself is host code outside the path. */
int unpp_topk_peaks(const float* heat, int planes, int H, int W, int num, float threshold, int32_t* xy, float* val, int32_t* count, unpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Training step.  Reference counterparts: autograd of models/unet.py:255-300 (cuDNN dgrad/wgrad,
 * BatchNorm / MaxPool / Dropout / sigmoid backward), nn.MSELoss (trainer/trainer.py:427) and
 * tools/optimizers/adamw.py:38-100.
 * ------------------------------------------------------------------------------------------- */

/* Weight gradient dW[tap][ci][co] = sum_pixels X[pix + tap][ci] * dZ[pix][co] as per-CTA partials
 * fp32 [unpp_wgrad_grid()][taps][sum src_C][cout], reduced by unpp_wgrad_reduce. */
typedef struct UnppWgradArgs {
  int32_t N, H, W;               /* pixel grid */
  int32_t nsrc;
  const void* src[UNPP_MAX_SRC]; /* forward inputs of the conv (virtual concat), NHWC bf16 dense    */
  int32_t src_C[UNPP_MAX_SRC];
  const void* dz;                /* NHWC bf16 [N, H*dz_step, W*dz_step, cout]                        */
  int32_t cout;
  int32_t dz_step, dz_oy, dz_ox; /* 1,0,0 = dense; 2,p,q = tap (p,q) of a k2s2 transposed conv;
                                    2,-1,-1 = all four taps in one launch: partial is [4 (2p+q)][grid][1][cin][cout] */
  int32_t taps;                  /* 9 or 1 */
  float* partial;
} UnppWgradArgs;
int unpp_wgrad(const UnppWgradArgs* a, unpp_stream_t stream);
int unpp_wgrad_grid(const UnppWgradArgs* a);
/* dst[co*s_co + ci*s_ci + tap*s_tap] = scale * sum_p partial[p][tap][ci_begin+ci][co], ci < ci_count */
int unpp_wgrad_reduce(const float* partial, int nparts, int taps, int cin_total, int cout, float* dst, int ci_begin, int ci_count,
                      long s_co, long s_ci, long s_tap, float scale, unpp_stream_t stream);
/* Many reductions of the two kinds above in ONE launch (the training step defers every reduction whose
 * result is only needed by the optimizer): job j writes
 *   dst[co*s_co + ci*s_ci + tap*s_tap] = scale * sum_p partial[p*stride + (tap*cin_total + ci_begin + ci)*cout + co]
 * for tap < taps, ci < ci_count, co < cout, with the same fixed summation order as the single-job kernels.
 * `table` is a DEVICE array of njobs jobs; block_end = exclusive running total of ceil(outputs/128) blocks. */
typedef struct UnppReduceJob {
  const float* partial;
  float* dst;
  int64_t stride, s_co, s_ci, s_tap;
  int32_t nparts, taps, cin_total, cout, ci_begin, ci_count;
  float scale;
  int32_t block_end;
} UnppReduceJob;
int unpp_reduce_batched(const UnppReduceJob* table, int njobs, int total_blocks, unpp_stream_t stream);
int unpp_sizeof_reduce_job(void);
/* out[i] (+)= scale * sum_p partial[p*stride + i], i < n (fixed order: deterministic) */
int unpp_reduce_partials(const float* partial, int nparts, long stride, int n, float scale, float* out, int accumulate,
                         unpp_stream_t stream);

/* BatchNorm2d batch statistics from the conv epilogue partials [nparts][2][C] (sum, sum of squares):
 * mean, inverse std (biased variance, eps), fused affine scale = gamma*istd, shift = beta - mean*scale,
 * and the running-stat update with the unbiased variance (running_* may be NULL). */
int unpp_bn_finalize(const float* partial, int nparts, int C, float count, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, float* mean, float* istd, float* scale, float* shift,
                     unpp_stream_t stream);
/* y = relu(z*scale + shift) on NHWC bf16; pooled (optional) = 2x2/2 max pool of y. */
int unpp_bn_relu(const void* z, const float* scale, const float* shift, void* y, void* pooled, int N, int H, int W, int C,
                 unpp_stream_t stream);
/* MaxPool2d(2) backward: dx[N,H,W,C] from dpooled[N,H/2,W/2,C] and the forward input x (first max wins). */
int unpp_maxpool2x2_bwd(const void* x, const void* dpooled, void* dx, int N, int H, int W, int C, unpp_stream_t stream);
/* BatchNorm backward apply: dz = gamma*istd*(dyh - sums[c]/count - xhat*sums[C+c]/count), xhat=(z-mean)*istd. */
int unpp_bn_bwd_apply(const void* dyh, const void* z, const float* mean, const float* istd, const float* gamma, const float* sums,
                      float count, void* dz, int N, int H, int W, int C, unpp_stream_t stream);
/* Head backward (sigmoid' + 1x1 conv dgrad/wgrad + dropout mask).  Exactly one of dheat (upstream
 * gradient, fp32 NCHW) and target is non-NULL.  With target the loss gradient is fused:
 *   loss_kind 0, nn.MSELoss (trainer/trainer.py:427): dheat = coef*(heat-target), loss partial = sum (heat-target)^2
 *   loss_kind 1, FocalLoss_BCE_2d (tools/losses/focal_loss.py:264-301, the criterion the trainer ships, trainer.py:426):
 *                a = |heat-target|, e = 1-a+1e-20, loss partial = sum -a^gamma*log(e), dheat = coef * d/dheat of that term  dx is already multiplied by the ReLU mask [x > 0] of the conv that produced x.
 * partial: fp32 [unpp_head_bwd_grid()][classes*16 + classes + 1 + 16] = dW, db, loss, per-channel sum of dx. */
int unpp_head_bwd(const float* heat, const float* dheat, const float* target, int loss_kind, float gamma, float coef, const void* x, const uint16_t* drop_mask,
                  float drop_scale, const float* head_w, int classes, void* dx, float* partial, int N, int H, int W,
                  unpp_stream_t stream);
int unpp_head_bwd_grid(int N, int H, int W);
/* The reference's AdamW on a flat fp32 buffer (decay = weight_decay * p_old, not scaled by lr);
 * step is the 1-based step count; gradients are multiplied by grad_scale first (1/world for DP). */
int unpp_adamw(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps, float weight_decay,
               int step, float grad_scale, unpp_stream_t stream);
/* Same, with the step count kept on the device (CUDA-graph friendly): increments *step_counter,
 * derives the bias-corrected step size from it into *step_size_scratch, then updates. */
int unpp_adamw_dev(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, uint64_t* step_counter, float* step_size_scratch, float grad_scale, unpp_stream_t stream);
/* Every optimizer the reference trainer can select (trainer/trainer.py:344-376), one launch over a flat fp32 buffer. */
#define UNPP_OPT_ADAMW 0     /* tools/optimizers/adamw.py:38-100 (decay = weight_decay * p_old, not scaled by lr)            */
#define UNPP_OPT_ADAM 1      /* torch.optim.Adam(lr, weight_decay): L2 decay folded into the gradient                          */
#define UNPP_OPT_ADABOUND 2  /* tools/optimizers/adabound.py:57-122 without amsbound                                           */
#define UNPP_OPT_SGD 3       /* torch.optim.SGD(lr, momentum = beta1, weight_decay), dampening 0, no Nesterov                  */
#define UNPP_OPT_SGDW 4      /* tools/optimizers/sgdw.py:77-110 AS SHIPPED: momentum buffer (beta1 = momentum, beta2 =
                                dampening) is maintained but never applied; the parameter only decays: p -= weight_decay * p  */
typedef struct UnppOptimArgs {
  int32_t kind;
  float lr;
  float beta1, beta2;     /* Adam family: betas; SGD: beta1 = momentum; SGDW: beta1 = momentum, beta2 = dampening */
  float eps;
  float weight_decay;
  float final_lr, gamma;  /* AdaBound */
  float base_lr;          /* AdaBound: the learning rate at construction (adabound.py:49,119)                     */
  float grad_scale;       /* gradients are multiplied by this first (1/world for data-parallel sums)              */
} UnppOptimArgs;
/* state1 / state2: first / second moment (Adam family) or the momentum buffer / unused (SGD, SGDW; may be NULL without
 * momentum).  Either `step` (1-based, host) or `step_counter` (device; incremented by the call, CUDA-graph friendly; then
 * scalars_scratch = fp32[4] device scratch and lr_dev may point to a device float that overrides a->lr at every launch,
 * which is how an lr scheduler drives a captured step). */
int unpp_optim_step(float* p, const float* g, float* state1, float* state2, long n, const UnppOptimArgs* a, int step, uint64_t* step_counter,
                    const float* lr_dev, float* scalars_scratch, unpp_stream_t stream);
/* The same update for many tensors in ONE launch (the drop-in optimizers on a model whose parameters are separate allocations: the 74
 * tensors of UNet_Nested; tools/optimizers/adamw.py:49-98 loops over them with ~10 launches each).  `table_dev`: DEVICE array of
 * ntensors entries; block_end = exclusive running total of ceil(n / 1024) blocks; all tensors share the step count `step` (1-based). */
typedef struct UnppOptimTensor {
  float* p;
  const float* g;
  float* state1;
  float* state2;
  int64_t n;
  int32_t block_end;
  int32_t reserved;
} UnppOptimTensor;
int unpp_optim_step_multi(const UnppOptimTensor* table_dev, int ntensors, int total_blocks, const UnppOptimArgs* a, int step, unpp_stream_t stream);
int unpp_sizeof_optim_tensor(void);
int unpp_sizeof_optim_args(void);
/* nn.UpsamplingBilinear2d(scale_factor=2) (= bilinear, align_corners=True; models/unet.py:190) on NHWC bf16:
 * y[N,2H,2W,C] from x[N,H,W,C], C % 8 == 0, with ATen's float source-index arithmetic; and its exact adjoint
 * dx[N,H,W,C] from dy[N,2H,2W,C] (a gather with the same weights: deterministic). */
int unpp_bilinear_up2x(const void* x, void* y, int N, int H, int W, int C, unpp_stream_t stream);
int unpp_bilinear_up2x_bwd(const void* dy, void* dx, int N, int H, int W, int C, unpp_stream_t stream);
/* Keep-mask for the element-wise nn.Dropout(p) in front of the 16-channel heads (models/unet.py:254,283-286): one 16-bit
 * word per pixel, bit c = 1 with probability 1-p (counter-based hash of seed, pixel, channel and, when step_counter is
 * given, the device step counter: a captured training step draws fresh masks at every replay).  npix % 8 == 0. */
int unpp_dropout_mask(uint16_t* mask, long npix, float p_drop, uint64_t seed, const uint64_t* step_counter, unpp_stream_t stream);

/* Target heat maps from key points, the trainer's per-step CPU routine (tools/misc/helper.py:87-172, called at
 * trainer/trainer.py:122-123): keypoints fp32 [N][npts][2] as (x, y) -> out fp32 [N][4][H][W]; point groups
 * {0}, {1,2,3}, {4}, {5..npts-1}; exp(-0.5*distance/3) in float64; multi-point planes divided by their maximum. */
int unpp_create_heatmap(const float* keypoints, int N, int npts, int H, int W, float* out, unpp_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * fp32 validation mode of the inference path (UNet_Nested.precision = "fp32"): the same fused plan on NCHW fp32 tensors with fp32
 * FMAs on the CUDA cores, straight from the OIHW state_dict weights — BASELINE's "<= 1e-3 relative in fp32/TF32" tolerance; used
 * to separate wiring errors from bf16 rounding, never benchmarked.  models/unet.py:121-156,182-202,242-244,283-286.
 * ------------------------------------------------------------------------------------------- */
typedef struct UnppRefConvArgs {
  int32_t N, H, W, nsrc;
  const void* src[UNPP_MAX_SRC]; /* NCHW fp32 sources, concatenated along the channel axis in this order (unet.py:199-201) */
  int32_t src_C[UNPP_MAX_SRC];
  const float* weight;           /* fp32 [cout][sum src_C][k][k], k = 3 (taps 9, pad 1) or 1 (taps 1) */
  const float* scale;            /* per output channel (eval-mode BatchNorm) or NULL */
  const float* bias;             /* per output channel or NULL */
  int32_t cout, taps, relu, sigmoid;
  float* out;                    /* NCHW fp32 [N][cout][H][W] */
} UnppRefConvArgs;
int unpp_ref_conv(const UnppRefConvArgs* a, unpp_stream_t stream);
/* nn.ConvTranspose2d(k=2, s=2) (unet.py:187): x [N,Cin,H,W], w [Cin,Cout,2,2], b [Cout] -> out [N,Cout,2H,2W], all fp32 */
int unpp_ref_deconv2x2(const float* x, const float* w, const float* b, float* out, int N, int Cin, int Cout, int H, int W, unpp_stream_t stream);
/* nn.MaxPool2d(2) on NCHW fp32 */
int unpp_ref_maxpool2x2(const float* x, float* out, int N, int C, int H, int W, unpp_stream_t stream);
int unpp_sizeof_ref_conv_args(void);

/* sizeof() of the argument structs as the C compiler sees them (binding self-check). */
int unpp_sizeof_conv_args(void);
int unpp_sizeof_pack_args(void);
int unpp_sizeof_wgrad_args(void);

#ifdef __cplusplus
}
#endif
#endif /* UNPP_H_ */
