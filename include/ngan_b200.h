/* ngan_b200 -- C ABI of the B200-native (sm_100a) kernels behind neuron-gan's progressive-growing WGAN-GP
 * training step.  The shared library is libngan_b200.so (built in-tree by __graft_entry__.build()).
 *
 * The reference (oliviertrottier/neuron-gan) has no FFI / operator interface: its hot path is Python calling
 * ATen operators.  Each entry point below therefore names the reference Python call
 * site(s) (file:line in the reference tree) whose arithmetic it replaces.  The Python mirror of the reference
 * interface (neuron_gan_b200/models.py, loss_functions.py, utils.py) calls these through ctypes; the same
 * symbols can be bound from any language.
 *
 * Conventions
 *  - every function returns 0 on success, a negative NGAN_ERR_* code otherwise; ngan_last_error() returns a
 *    thread-local message.  Nothing is allocated or freed on behalf of the caller, there are no hidden
 *    synchronisations, every launch goes to the cudaStream_t passed as `stream` (void*).
 *  - "c8" tensors are bf16 feature maps in C8-planar layout [B][C/8][H][W][8] (16-byte granules of 8 consecutive
 *    channels); C must be a multiple of 16 for the conv kernels.  Images, scores, PixelNorm scales `r`
 *    ([B][H][W]) and all parameters / gradients are fp32 in the reference's (torch) layouts.
 *  - parameter-gradient outputs (gw, gb, dw, what) are produced WITHOUT atomics, so they are bit-reproducible:
 *    the kernel writes per-block partial sums into a caller-provided fp32 `workspace` (sized by the matching
 *    *_workspace_bytes query) and a second launch adds the partials in index order; `accumulate` != 0 adds the
 *    result to the output, 0 overwrites it (no prior zeroing needed).  Two launches that accumulate into the SAME
 *    tensor must be ordered by the caller (they read-modify-write it); concurrent contributions go to separate
 *    tensors that are summed afterwards (ngan_sum_slots) -- neuron_gan_b200/train_step.py does that for the three
 *    contributions to every critic parameter (Wasserstein backward, and the two sweeps of the penalty's double
 *    backward, loss_functions.py:175 + train.py:365).
 *  - the conv kernels are launched with the programmatic-stream-serialization attribute: their prologue may overlap
 *    the tail of the previous kernel in the stream, their first global-memory access waits for it
 *    (griddepcontrol.wait), so stream order is preserved for any caller; NGAN_NO_PDL=1 in the environment disables it.
 *  - ngan_adam_multi takes at most 48 tensors per call.
 *  - `const float* dyn` after a scalar argument (the seven entry points the fade-in coefficient alpha reaches,
 *    models.py:348-350, 519-521 and their backward passes): optional DEVICE pointer; when non-NULL the effective
 *    scalar is the host value times *dyn, read by the kernel.  A launch captured in a CUDA graph then follows alpha,
 *    which the reference advances every epoch during a resolution transition (train.py:318-321).
 *  - LeakyReLU + PixelNorm always come as a pair after a conv (models.py:261-268); "pn_bwd" below means
 *    ga = mask(y) * r * (g - y * mean_c(g*y)), the gradient wrt the conv's pre-activation, computed from the
 *    saved PixelNorm output y and scale r = (mean_c(h^2) + 1e-8)^-1/2 (models.py:118, 126).
 */
#ifndef NGAN_B200_H
#define NGAN_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define NGAN_OK 0
#define NGAN_ERR_INVALID -1
#define NGAN_ERR_CUDA -2
#define NGAN_ERR_UNSUPPORTED -3

int ngan_version(void);
const char* ngan_last_error(void);

/* ---- layout conversion at the boundary (torch NCHW fp32 <-> c8 bf16) ---- */
int ngan_nchw_to_c8(const float* src, void* dst_c8, int B, int C, int H, int W, void* stream);
int ngan_c8_to_nchw(const void* src_c8, float* dst, int B, int C, int H, int W, void* stream);

/* ---- equalised-LR 3x3 convolution: Conv2d_normalized.forward, models.py:172-204 ---- */
/* fp32 weight [cout][cin][3][3] -> bf16 UMMA operand images for the forward conv and for the data-gradient conv */
int ngan_prep_conv_weight(const float* w, void* w_fwd, void* w_dgrad, int cin, int cout, void* stream);
/* 1 if the images of a cin -> cout layer use the kx-folded layout (the layer runs on the folded kernel) */
int ngan_conv_weight_is_folded(int cin, int cout);
/* y = PixelNorm(LeakyReLU(scale*conv(x,W) + bias)), r = PixelNorm scale.  models.py:203-204 + 263-268 (+110-126) */
int ngan_conv3x3_fwd(const void* x_c8, const void* w_fwd, const float* bias, float scale, float leak, void* y_c8,
                     float* r, int B, int cin, int cout, int H, int W, void* stream);
/* The generator's last conv with ToImage fused (models.py:141-149, 344-353): additionally img[b,y,x] =
 * tanh(sum_c toim_w[c] * y[b,c,y,x]) ([B][H][W]; fp32, or bf16 when img_bf16 != 0 -- generator-only inference,
 * utils.py:346-355).  y_c8 and r may be NULL when only the image is wanted (the detached generator passes of the
 * critic step, eval.py).  Folded layers only (ngan_conv_weight_is_folded). */
int ngan_conv3x3_fwd_toim(const void* x_c8, const void* w_fwd, const float* bias, float scale, float leak, void* y_c8,
                          float* r, const float* toim_w, void* img, int img_bf16, int B, int cin, int cout, int H,
                          int W, void* stream);
/* gx = scale * convT(ga, W): autograd's convolution_backward (input gradient) of models.py:204.  cin/cout are the
 * LAYER's channel counts: ga has cout channels, gx has cin. */
int ngan_conv3x3_dgrad(const void* ga_c8, const void* w_dgrad, float scale, void* gx_c8, int B, int cin, int cout,
                       int H, int W, void* stream);
/* same, with the pn_bwd of the producing layer fused: ga_prev = pn_bwd(gx; y_prev, r_prev) (+ addin).  gy_out
 * (optional) receives gx itself, which the gradient-penalty double backward needs later. */
int ngan_conv3x3_dgrad_pn(const void* ga_c8, const void* w_dgrad, float scale, float leak, const void* y_prev_c8,
                          const float* r_prev, const void* addin_c8, void* ga_prev_c8, void* gy_out_c8, int B,
                          int cin, int cout, int H, int W, void* stream);
/* Double backward through conv + LeakyReLU + PixelNorm (the create_graph=True path of loss_functions.py:175):
 * ghat_x is the cotangent on the layer's first-order input gradient; emits ghat_y (cotangent on the first-order
 * gradient wrt this layer's output, i.e. the next layer's ghat_x) and ahat (cotangent injected at this layer's
 * pre-activation for the second sweep).  y, r, gy are this layer's saved forward output / scale / first-order
 * output gradient.  Formulas: SURVEY.md section 8a row 3. */
int ngan_conv3x3_dbl(const void* ghat_x_c8, const void* w_fwd, float scale, float leak, const void* y_c8,
                     const float* r, const void* gy_c8, void* ghat_y_c8, void* ahat_c8, int B, int cin, int cout,
                     int H, int W, void* stream);
/* dw[cout][cin][3][3] (+)= scale * corr(x, ga): convolution_backward (weight gradient) of models.py:204.
 * Two launches: the persistent mma.sync kernel writes one partial dW image per pixel CTA into `workspace`
 * (ngan_conv3x3_wgrad_workspace_bytes), ngan_reduce_partials adds them in order. */
long long ngan_conv3x3_wgrad_workspace_bytes(int B, int cin, int cout, int H, int W);
int ngan_conv3x3_wgrad(const void* x_c8, const void* ga_c8, float scale, float* dw, int accumulate, float* workspace,
                       int B, int cin, int cout, int H, int W, void* stream);
/* upper bound of the workspace any of the per-pixel parameter-gradient reductions below (bias_grad, fromim_bwd,
 * fromim_dbl, toim_bwd) needs on a [B][C][H][W] tensor */
long long ngan_pixel_reduction_workspace_bytes(int B, int C, int H, int W);
/* gb[c] (+)= sum_{b,y,x} ga: bias gradient of the biased 128->128 conv, models.py:469-471 */
int ngan_bias_grad(const void* ga_c8, float* gb, int accumulate, float* workspace, int B, int C, int H, int W,
                   void* stream);
/* out[i] (+)= scale * sum_p partials[p*ld + i], i < n, rows added in index order (the second stage of every
 * parameter-gradient reduction; exported for hosts that want to place it on another stream) */
int ngan_reduce_partials(const float* partials, int n_partials, long long n, long long ld, float scale, float* out,
                         int accumulate, void* stream);
/* out[i] = scale * (slots[i] + slots[ld + i] + ...) (n_slots terms, in slot order; ld in floats) */
int ngan_sum_slots(const float* slots, int n_slots, long long n, long long ld, float scale, float* out, void* stream);
/* cudaMemsetAsync through the C ABI (gradient-slot zeroing inside a captured iteration; D.zero_grad(), train.py:357) */
int ngan_memset(void* dst, int value, long long bytes, void* stream);

/* ---- resampling: Interpolate(x2, bilinear) models.py:78-89, 257; AvgPool2d(2) models.py:254 ---- */
int ngan_upsample2x(const void* x_c8, void* out_c8, int B, int C, int H, int W, void* stream);
int ngan_avgpool2(const void* x_c8, void* out_c8, int B, int C, int H, int W, void* stream);
/* ga = pn_bwd(gscale * g) (+ addin); unpool != 0 reads g at (H/2, W/2) (adjoint of AvgPool2d(2); fold the 1/4
 * into gscale).  gy_out (optional) receives gscale*g at full resolution. */
int ngan_pn_bwd(const void* g_c8, int unpool, float gscale, const float* dyn, const void* y_c8, const float* r, const void* addin_c8,
                void* ga_c8, void* gy_out_c8, float leak, int B, int C, int H, int W, void* stream);
/* adjoint of the bilinear x2 upsample fused with pn_bwd of the layer below it; extra_pre/extra_w add the
 * faded-out ToImage branch's contribution extra_w[c]*extra_pre[b,y,x] (models.py:348).  H, W: low resolution. */
int ngan_up2_bwd_pn_bwd(const void* g_up_c8, const void* y_c8, const float* r, const float* extra_pre,
                        const float* extra_w, void* ga_c8, float leak, int B, int C, int H, int W, void* stream);

/* ---- 1-channel fp32 image helpers (fade-in paths, models.py:335, 348, 507, 519; loss_functions.py:171) ---- */
int ngan_pool_image(const float* x, float* out, int B, int H, int W, void* stream);
int ngan_unpool_image(const float* g, float* out, float scale, int B, int H, int W, void* stream);
int ngan_up2_image(const float* x, float* out, int B, int H, int W, void* stream);
int ngan_up2_image_bwd(const float* g, float* out, float scale, const float* dyn, int B, int H, int W, void* stream);
int ngan_lerp(const float* a, const float* b, float alpha, const float* dyn, float* out, long long n, void* stream);
int ngan_axpby(const float* a, float ca, const float* b, float cb, float* out, long long n, void* stream);
int ngan_interp_images(const float* real, const float* fake, const float* eps, float* out, int B,
                       long long per_sample, void* stream);
int ngan_scale_rows(const float* x, const float* coeff, float scale, float* out, int B, long long per_sample,
                    void* stream);

/* ---- FromImage / ToImage 1x1 convolutions, models.py:133-165 ---- */
int ngan_fromim_fwd(const float* xp, const float* w, const float* b, void* out_c8, int B, int C, int H, int W,
                    void* stream);
/* discriminator fade-in, models.py:519-521: out = y_start + alpha*(y_end - y_start), y_start = FromIm_old(xp);
 * b_old may be NULL (no bias: the double backward reuses the kernel for the bias-free cotangent blend) */
int ngan_d_fade_fwd(const void* y_end_c8, const float* xp, const float* w_old, const float* b_old, float alpha,
                    const float* dyn, void* out_c8, int B, int C, int H, int W, void* stream);
int ngan_fromim_bwd(const void* g_c8, int unpool, float gscale, const float* dyn, const float* xp, const float* w, float* gw, float* gb,
                    int grad_accumulate, float* workspace, float* g_img, int g_img_accumulate, int B, int C, int H,
                    int W, void* stream);
int ngan_fromim_dbl(const float* ghat_xp, float in_scale, const void* g_c8, int unpool, float gscale,
                    const float* dyn, const float* w,
                    void* ghat_out_c8, float* what, int grad_accumulate, float* workspace, int B, int C, int H, int W,
                    void* stream);
int ngan_toim_fwd(const void* y_c8, const float* w, float* img, int B, int C, int H, int W, void* stream);
/* fp32 -> bf16 (image output of the inference path when ToImage was not fused into the last conv) */
int ngan_f32_to_bf16(const float* src, void* dst_bf16, long long n, void* stream);
int ngan_toim_bwd(const float* g_img, float gscale, const float* dyn, const float* img, const void* y_c8, const float* r,
                  const float* w, void* ga_c8, float* gpre, float* gw, int grad_accumulate, float* workspace, float leak,
                  int B, int C, int H, int W, void* stream);

/* ---- critic head: Conv2d_normalized(F -> 1, kernel S x S, pad 0) + Flatten, models.py:485-490 ---- */
int ngan_head_fwd(const void* y_c8, const float* w, const float* bias, float scale, float* score, int B, int C,
                  int S, void* stream);
int ngan_head_bwd_pn(const float* gout, const float* w, float scale, const void* y_c8, const float* r, void* ga_c8,
                     void* gy_out_c8, float leak, int B, int C, int S, void* stream);
/* gw[c,p] (+)= scale * sum_b coeff[b] * t[b,c,p]; gb (optional): gb[0] (+)= sum_b coeff[b], the bias gradient */
int ngan_head_wgrad(const void* t_c8, const float* coeff, float scale, float* gw, float* gb, int accumulate, int B,
                    int C, int S, void* stream);

/* ---- generator stem: Linear_normalized + Unflatten + LeakyReLU + PixelNorm, models.py:299-311 ---- */
/* fp32 master [C*S*S][K] -> bf16 operand image [S*S][K/8][C][8] (the C weight rows of one pixel form one
 * contiguous no-swizzle K-major UMMA operand) */
int ngan_prep_linear_weight(const float* w, void* w_img, int K, int C, int S, void* stream);
/* bytes of scratch ngan_linear_fwd needs (the latent batch as a bf16 UMMA operand, padded to 128 rows) */
long long ngan_linear_fwd_workspace_bytes(int B, int K);
int ngan_linear_fwd(const float* z, const void* w_img, float scale, float leak, void* y_c8, float* r, void* workspace,
                    int B, int K, int C, int S, void* stream);
/* dW[f][k] (+)= scale * sum_b ga[b][f] * z[b][k]; accumulate == 0 overwrites dW (no read, no prior zeroing needed) */
int ngan_linear_wgrad(const void* ga_c8, const float* z, float scale, float* dw, int accumulate, int B, int K, int C,
                      int S, void* stream);

/* ---- WGAN-GP loss reductions, loss_functions.py:14-47, 59-74, 157-180 ---- */
/* out3 = {D_loss, score_real, score_fake}; g_real/g_fake = d(gscale*D_loss)/d(score) per sample */
int ngan_wloss(const float* s_real, const float* s_fake, float drift, float* out3, float* g_real, float* g_fake,
               float gscale, int B, void* stream);
int ngan_gloss(const float* s_fake, float* out1, float* g_fake, float gscale, int B, void* stream);
/* pen = lambda*mean_b((norm_scale*||g_b|| - 1)^2); coeff_b = gscale * dpen/dnorm_b / norm_b (0 where the norm is 0,
 * torch's subgradient).  workspace: ngan_gp_loss_workspace_bytes(B) bytes for the per-sample partial sums. */
long long ngan_gp_loss_workspace_bytes(int B);
int ngan_gp_loss(const float* g, float norm_scale, float lambda, float* pen, float* coeff, float gscale,
                 float* workspace, int B, long long per_sample, void* stream);

/* similarity_loss (loss_functions.py:185-205, called at train.py:379-381): out[0] = lambda / (B*(B-1)) *
 * sum_ij (cos(z_i, z_j) - cos(x_i, x_j))^2 for images x [B][per_image] and latents z [B][latent], both fp32.  Two
 * launches (per-chunk partial Gram matrices, then an ordered reduction): deterministic.  With data parallelism the
 * caller all-gathers the rows of both tensors first (the Gram couples samples across ranks). */
long long ngan_similarity_loss_workspace_bytes(int B, long long per_image);
int ngan_similarity_loss(const float* images, const float* z, float lambda, float* workspace, float* out, int B,
                         long long per_image, int latent, void* stream);

/* stats[5] = {D_loss + pen, score_real, score_fake, G_loss, pen}: train.py:362 and the six .item() reads of
 * train.py:389-394 as one packed device tensor */
int ngan_pack_stats(const float* out3, const float* out1, const float* pen, float* stats, void* stream);

/* ---- multi-tensor Adam, train.py:220-225 (torch.optim.Adam semantics, one step count per parameter) ---- */
typedef struct {
    float* p;            /* parameter, fp32, 16-byte aligned */
    const float* g;      /* gradient */
    float* m;            /* exp_avg */
    float* v;            /* exp_avg_sq */
    void* shadow_bf16;   /* optional bf16 copy of p refreshed in the same pass, or NULL */
    long long n;
    float step_size;     /* lr / (1 - beta1^t), t = this parameter's step count after the update */
    float inv_bc2_sqrt;  /* 1 / sqrt(1 - beta2^t) */
    int shadow_k, shadow_c, shadow_ss;   /* shadow_kind 1: {K, C, S*S} of ngan_prep_linear_weight's image;
                            shadow_kind 2: {cin, cout, ngan_conv_weight_is_folded(cin, cout)} of a 3x3 conv weight */
    int shadow_kind;     /* 0: shadow has p's layout; 1: shadow is the linear operand image; 2: shadow is the forward
                            image of ngan_prep_conv_weight immediately followed by the data-gradient image */
    const float* dyn;    /* optional DEVICE pointer to {step_size, inv_bc2_sqrt}; when non-NULL it overrides the two
                            fields above at run time (lets a launch captured in a CUDA graph follow the step count) */
} ngan_adam_tensor;
int ngan_adam_multi(const ngan_adam_tensor* tensors, int n_tensors, float beta1, float beta2, float eps,
                    void* stream);

/* Adam on the generator's Linear weight [C*S*S][K] with the gradient taken from its FACTORS instead of a materialised
 * tensor (models.py:240-241 backward + train.py:385):  g[f][k] = gscale * sum_b ga[b][f] * z[b][k] over Btot samples,
 * ga = gradient at the stem's pre-activation (c8, what ngan_linear_wgrad takes), z = the latent batch.  Bit-identical
 * to ngan_linear_wgrad + ngan_adam_multi, without the 67 MB gradient round trip; with data parallelism the ranks
 * all-gather the factors (1 MB + 32 KB per rank) instead of all-reducing the product.  Sample b is row b % b_per_seg
 * of segment b / b_per_seg (segments *_seg_stride BYTES apart: the layout of an all-gathered buffer; one segment:
 * b_per_seg = Btot).  shadow_img (optional): the operand image of ngan_prep_linear_weight, refreshed in the same pass;
 * g_out (optional): also write the gradient.  p == NULL (then m, v are ignored and g_out is required): only form the
 * gradient -- the tensor-core replacement of ngan_linear_wgrad for an all-gathered global batch.  step_size /
 * inv_bc2_sqrt / dyn as in ngan_adam_tensor. */
int ngan_adam_linear_factored(float* p, float* m, float* v, void* shadow_img, float* g_out, const void* ga_c8,
                              const float* z, int Btot, int b_per_seg, long long ga_seg_stride, long long z_seg_stride,
                              int K, int C, int S, float gscale, float step_size, float inv_bc2_sqrt,
                              const float* dyn, float beta1, float beta2, float eps, void* stream);

/* ---- on-device image pipeline: DatasetIterator.__next__ (data/NeuronDataset.py:170-205) applying the transform
 * list of NeuronDataset.__init__ / set_image_size (data/NeuronDataset.py:112-126, 149-164) to a whole batch:
 * RandomAffine(nearest, fill 0) -> RandomVerticalFlip -> ColorJitter(brightness, contrast) -> CenterCrop(crop) ->
 * Renormalize((-1,1),(0,1)) -> Resize(out_size, antialias=True).  Three launches.
 *   canvases  [n_images][canvas][canvas] fp32 in [0,1]: the preloaded padded images (dataset.images), resident
 *   src_index [batch] which canvas each output row is made from
 *   params    [batch][16] fp32, drawn by the host in the reference's RNG order:
 *             0..5 inverse affine matrix (torchvision _get_inverse_affine_matrix) divided by canvas/2, row-major 2x3
 *             6 flip  7 brightness factor  8 contrast factor  9 fp32(1 - contrast factor)
 *             10 order (0: brightness then contrast, 1: contrast then brightness)  11 identity (no augmentation)
 *   tap_first/tap_count [out_size], tap_weight [out_size][max_taps]: the antialias filter of ATen's
 *             _compute_indices_weights_aa for crop -> out_size (one entry of weight 1 per index when out_size == crop)
 *   workspace ngan_augment_workspace_bytes(batch, canvas, crop) bytes, 16-byte aligned (partial sums for the
 *             per-image mean that adjust_contrast needs, and the resampled crop window of every image);
 *              out [batch][1][out_size][out_size] fp32 */
long long ngan_augment_workspace_bytes(int batch, int canvas, int crop);
int ngan_augment_batch(const float* canvases, const int* src_index, const float* params, const int* tap_first,
                       const int* tap_count, const float* tap_weight, int max_taps, float* workspace, float* out,
                       int batch, int canvas, int crop, int out_size, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NGAN_B200_H */
