// extern "C" surface of libngan_b200.so (declared in include/ngan_b200.h): argument validation, error
// reporting, and forwarding to the launchers.  No torch types cross this boundary.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "../../include/ngan_b200.h"
#include "common.cuh"
#include "conv_args.cuh"
#include "kernels.h"

namespace ngan {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return NGAN_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NGAN_ERR_CUDA;
}
int check_launch(const char* what) { return check_cuda(cudaGetLastError(), what); }
bool pdl_enabled() {
    static const bool on = getenv("NGAN_NO_PDL") == nullptr;
    return on;
}

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }
static inline bool bad_c(int C) { return C <= 0 || (C % 8) != 0; }

}  // namespace ngan

using namespace ngan;

#define NGAN_REQUIRE(cond, ...)        \
    do {                               \
        if (!(cond)) {                 \
            set_error(__VA_ARGS__);    \
            return NGAN_ERR_INVALID;   \
        }                              \
    } while (0)

namespace ngan { extern long long* g_conv_trace; extern long long* g_wgrad_trace; }

extern "C" {

int ngan_version(void) { return 100; }
// undocumented debug hook (NGAN_CONV_TRACE=1): copies the clock64() trace of the last folded-conv launch
int ngan_debug_conv_trace(long long* host_out) {
    if (!ngan::g_conv_trace) return -1;
    return check_cuda(cudaMemcpy(host_out, ngan::g_conv_trace, 32 * 8 * sizeof(long long), cudaMemcpyDeviceToHost), "trace");
}
// undocumented debug hook: device buffer (8 long long per CTA) that the next wgrad launches stamp with %globaltimer
int ngan_debug_wgrad_trace(long long* device_buf) {
    ngan::g_wgrad_trace = device_buf;
    return 0;
}
const char* ngan_last_error(void) { return g_err; }

int ngan_nchw_to_c8(const float* src, void* dst, int B, int C, int H, int W, void* stream) {
    NGAN_REQUIRE(src && dst && !bad_c(C) && B > 0 && H > 0 && W > 0, "nchw_to_c8: bad arguments (C=%d)", C);
    return nchw_to_c8(src, dst, B, C, H, W, S(stream));
}
int ngan_c8_to_nchw(const void* src, float* dst, int B, int C, int H, int W, void* stream) {
    NGAN_REQUIRE(src && dst && !bad_c(C) && B > 0 && H > 0 && W > 0, "c8_to_nchw: bad arguments (C=%d)", C);
    return c8_to_nchw(src, dst, B, C, H, W, S(stream));
}
int ngan_prep_conv_weight(const float* w, void* w_fwd, void* w_dgrad, int cin, int cout, void* stream) {
    NGAN_REQUIRE(w && (w_fwd || w_dgrad) && !bad_c(cin) && !bad_c(cout), "prep_conv_weight: bad arguments");
    return prep_conv_weight(w, w_fwd, w_dgrad, cin, cout, S(stream));
}
int ngan_conv_weight_is_folded(int cin, int cout) { return conv_uses_folded_kernel(cin, cout) ? 1 : 0; }
int ngan_conv3x3_fwd(const void* x, const void* w_fwd, const float* bias, float scale, float leak, void* y, float* r,
                     int B, int cin, int cout, int H, int W, void* stream) {
    NGAN_REQUIRE(x && w_fwd && y && B > 0, "conv3x3_fwd: null pointer or empty batch");
    return conv3x3_dispatch(EPI_FWD_PN, x, w_fwd, B, cin, cout, H, W, scale, leak, bias, y, nullptr, r, nullptr,
                            nullptr, nullptr, nullptr, S(stream));
}
int ngan_conv3x3_fwd_toim(const void* x, const void* w_fwd, const float* bias, float scale, float leak, void* y, float* r,
                          const float* toim_w, void* img, int img_bf16, int B, int cin, int cout, int H, int W,
                          void* stream) {
    NGAN_REQUIRE(x && w_fwd && toim_w && img && B > 0, "conv3x3_fwd_toim: null pointer or empty batch");
    return conv3x3_dispatch(EPI_FWD_PN, x, w_fwd, B, cin, cout, H, W, scale, leak, bias, y, nullptr, r, nullptr,
                            nullptr, nullptr, nullptr, S(stream), toim_w, static_cast<float*>(img), img_bf16);
}
int ngan_conv3x3_dgrad(const void* ga, const void* w_dgrad, float scale, void* gx, int B, int cin, int cout, int H,
                       int W, void* stream) {
    NGAN_REQUIRE(ga && w_dgrad && gx && B > 0, "conv3x3_dgrad: null pointer or empty batch");
    return conv3x3_dispatch(EPI_LINEAR, ga, w_dgrad, B, cout, cin, H, W, scale, 0.f, nullptr, gx, nullptr, nullptr,
                            nullptr, nullptr, nullptr, nullptr, S(stream));
}
int ngan_conv3x3_dgrad_pn(const void* ga, const void* w_dgrad, float scale, float leak, const void* y_prev,
                          const float* r_prev, const void* addin, void* ga_prev, void* gy_out, int B, int cin,
                          int cout, int H, int W, void* stream) {
    NGAN_REQUIRE(ga && w_dgrad && y_prev && r_prev && ga_prev && B > 0, "conv3x3_dgrad_pn: null pointer");
    return conv3x3_dispatch(EPI_BWD_PN, ga, w_dgrad, B, cout, cin, H, W, scale, leak, nullptr, ga_prev, gy_out,
                            nullptr, y_prev, r_prev, nullptr, addin, S(stream));
}
int ngan_conv3x3_dbl(const void* ghat_x, const void* w_fwd, float scale, float leak, const void* y, const float* r,
                     const void* gy, void* ghat_y, void* ahat, int B, int cin, int cout, int H, int W, void* stream) {
    NGAN_REQUIRE(ghat_x && w_fwd && y && r && gy && ghat_y && ahat && B > 0, "conv3x3_dbl: null pointer");
    return conv3x3_dispatch(EPI_DBL, ghat_x, w_fwd, B, cin, cout, H, W, scale, leak, nullptr, ghat_y, ahat, nullptr, y,
                            r, gy, nullptr, S(stream));
}
long long ngan_conv3x3_wgrad_workspace_bytes(int B, int cin, int cout, int H, int W) {
    return static_cast<long long>(conv3x3_wgrad_workspace_bytes(B, cin, cout, H, W));
}
int ngan_conv3x3_wgrad(const void* x, const void* ga, float scale, float* dw, int accumulate, float* workspace, int B,
                       int cin, int cout, int H, int W, void* stream) {
    NGAN_REQUIRE(x && ga && dw && workspace && B > 0, "conv3x3_wgrad: null pointer or empty batch");
    return conv3x3_wgrad(x, ga, scale, dw, accumulate, workspace, B, cin, cout, H, W, S(stream));
}
long long ngan_pixel_reduction_workspace_bytes(int B, int C, int H, int W) {
    return static_cast<long long>(pixel_reduction_workspace_bytes(B, C, H, W));
}
int ngan_bias_grad(const void* ga, float* gb, int accumulate, float* workspace, int B, int C, int H, int W,
                   void* stream) {
    NGAN_REQUIRE(ga && gb && workspace && !bad_c(C) && B > 0, "bias_grad: bad arguments");
    return bias_grad_c8(ga, gb, accumulate, workspace, B, C, H, W, S(stream));
}
int ngan_reduce_partials(const float* partials, int n_partials, long long n, long long ld, float scale, float* out,
                         int accumulate, void* stream) {
    NGAN_REQUIRE(partials && out && n_partials > 0 && n > 0 && ld >= n, "reduce_partials: bad arguments");
    return reduce_partials(partials, n_partials, n, ld, scale, out, accumulate, S(stream));
}
int ngan_sum_slots(const float* slots, int n_slots, long long n, long long ld, float scale, float* out, void* stream) {
    NGAN_REQUIRE(slots && out && n_slots > 0 && n > 0 && ld >= n, "sum_slots: bad arguments");
    return sum_slots(slots, n_slots, n, ld, scale, out, S(stream));
}
int ngan_memset(void* dst, int value, long long bytes, void* stream) {
    NGAN_REQUIRE(dst && bytes >= 0, "memset: bad arguments");
    if (bytes == 0) return NGAN_OK;
    return check_cuda(cudaMemsetAsync(dst, value, static_cast<size_t>(bytes), S(stream)), "cudaMemsetAsync");
}
int ngan_upsample2x(const void* x, void* out, int B, int C, int H, int W, void* stream) {
    NGAN_REQUIRE(x && out && !bad_c(C) && B > 0, "upsample2x: bad arguments");
    return upsample2x_c8(x, out, B, C, H, W, S(stream));
}
int ngan_avgpool2(const void* x, void* out, int B, int C, int H, int W, void* stream) {
    NGAN_REQUIRE(x && out && !bad_c(C) && B > 0 && H % 2 == 0 && W % 2 == 0, "avgpool2: bad arguments");
    return avgpool2_c8(x, out, B, C, H, W, S(stream));
}
int ngan_pn_bwd(const void* g, int unpool, float gscale, const float* dyn, const void* y, const float* r, const void* addin, void* ga,
                void* gy_out, float leak, int B, int C, int H, int W, void* stream) {
    NGAN_REQUIRE(g && y && r && ga && !bad_c(C) && B > 0, "pn_bwd: bad arguments");
    NGAN_REQUIRE(!unpool || (H % 2 == 0 && W % 2 == 0), "pn_bwd: unpool needs even H, W");
    return pn_bwd_c8(g, unpool, gscale, dyn, y, r, addin, ga, gy_out, leak, B, C, H, W, S(stream));
}
int ngan_up2_bwd_pn_bwd(const void* g_up, const void* y, const float* r, const float* extra_pre, const float* extra_w,
                        void* ga, float leak, int B, int C, int H, int W, void* stream) {
    NGAN_REQUIRE(g_up && y && r && ga && B > 0, "up2_bwd_pn_bwd: null pointer");
    NGAN_REQUIRE((extra_pre == nullptr) == (extra_w == nullptr), "up2_bwd_pn_bwd: extra_pre/extra_w go together");
    return up2_bwd_pn_bwd_c8(g_up, y, r, extra_pre, extra_w, ga, leak, B, C, H, W, S(stream));
}
int ngan_pool_image(const float* x, float* out, int B, int H, int W, void* stream) {
    NGAN_REQUIRE(x && out && B > 0 && H % 2 == 0 && W % 2 == 0, "pool_image: bad arguments");
    return pool_image(x, out, B, H, W, S(stream));
}
int ngan_unpool_image(const float* g, float* out, float scale, int B, int H, int W, void* stream) {
    NGAN_REQUIRE(g && out && B > 0 && H % 2 == 0 && W % 2 == 0, "unpool_image: bad arguments");
    return unpool_image(g, out, scale, B, H, W, S(stream));
}
int ngan_up2_image(const float* x, float* out, int B, int H, int W, void* stream) {
    NGAN_REQUIRE(x && out && B > 0, "up2_image: bad arguments");
    return up2_image(x, out, B, H, W, S(stream));
}
int ngan_up2_image_bwd(const float* g, float* out, float scale, const float* dyn, int B, int H, int W, void* stream) {
    NGAN_REQUIRE(g && out && B > 0, "up2_image_bwd: bad arguments");
    return up2_image_bwd(g, out, scale, dyn, B, H, W, S(stream));
}
int ngan_lerp(const float* a, const float* b, float alpha, const float* dyn, float* out, long long n, void* stream) {
    NGAN_REQUIRE(a && b && out && n >= 0, "lerp: bad arguments");
    if (n == 0) return NGAN_OK;
    return lerp_f32(a, b, alpha, dyn, out, static_cast<size_t>(n), S(stream));
}
int ngan_axpby(const float* a, float ca, const float* b, float cb, float* out, long long n, void* stream) {
    NGAN_REQUIRE(a && out && n >= 0, "axpby: bad arguments");
    if (n == 0) return NGAN_OK;
    return axpby_f32(a, ca, b, cb, out, static_cast<size_t>(n), S(stream));
}
int ngan_interp_images(const float* real, const float* fake, const float* eps, float* out, int B,
                       long long per_sample, void* stream) {
    NGAN_REQUIRE(real && fake && eps && out && B > 0 && per_sample > 0, "interp_images: bad arguments");
    return interp_images(real, fake, eps, out, B, static_cast<size_t>(per_sample), S(stream));
}
int ngan_scale_rows(const float* x, const float* coeff, float scale, float* out, int B, long long per_sample,
                    void* stream) {
    NGAN_REQUIRE(x && coeff && out && B > 0 && per_sample > 0, "scale_rows: bad arguments");
    return scale_rows_f32(x, coeff, scale, out, B, static_cast<size_t>(per_sample), S(stream));
}
int ngan_fromim_fwd(const float* xp, const float* w, const float* b, void* out, int B, int C, int H, int W,
                    void* stream) {
    NGAN_REQUIRE(xp && w && b && out && !bad_c(C) && B > 0, "fromim_fwd: bad arguments");
    return fromim_fwd(xp, w, b, out, B, C, H, W, S(stream));
}
int ngan_d_fade_fwd(const void* y_end, const float* xp, const float* w_old, const float* b_old, float alpha,
                    const float* dyn, void* out, int B, int C, int H, int W, void* stream) {
    NGAN_REQUIRE(y_end && xp && w_old && out && !bad_c(C) && B > 0, "d_fade_fwd: bad arguments");
    return d_fade_fwd(y_end, xp, w_old, b_old, alpha, dyn, out, B, C, H, W, S(stream));
}
int ngan_fromim_bwd(const void* g, int unpool, float gscale, const float* dyn, const float* xp, const float* w, float* gw, float* gb,
                    int grad_accumulate, float* workspace, float* g_img, int g_img_accumulate, int B, int C, int H,
                    int W, void* stream) {
    NGAN_REQUIRE(g && xp && w && !bad_c(C) && B > 0, "fromim_bwd: bad arguments");
    NGAN_REQUIRE(workspace || !(gw || gb), "fromim_bwd: parameter gradients need a workspace");
    return fromim_bwd(g, unpool, gscale, dyn, xp, w, gw, gb, grad_accumulate, workspace, g_img, g_img_accumulate, B, C,
                      H, W, S(stream));
}
int ngan_fromim_dbl(const float* ghat_xp, float in_scale, const void* g, int unpool, float gscale, const float* dyn,
                    const float* w,
                    void* ghat_out, float* what, int grad_accumulate, float* workspace, int B, int C, int H, int W,
                    void* stream) {
    NGAN_REQUIRE(ghat_xp && g && w && !bad_c(C) && B > 0, "fromim_dbl: bad arguments");
    NGAN_REQUIRE(workspace || !what, "fromim_dbl: the weight gradient needs a workspace");
    return fromim_dbl(ghat_xp, in_scale, g, unpool, gscale, dyn, w, ghat_out, what, grad_accumulate, workspace, B, C,
                      H, W, S(stream));
}
int ngan_toim_fwd(const void* y, const float* w, float* img, int B, int C, int H, int W, void* stream) {
    NGAN_REQUIRE(y && w && img && !bad_c(C) && B > 0, "toim_fwd: bad arguments");
    return toim_fwd(y, w, img, B, C, H, W, S(stream));
}
int ngan_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
    NGAN_REQUIRE(src && dst && n >= 0, "f32_to_bf16: bad arguments");
    if (n == 0) return NGAN_OK;
    return f32_to_bf16(src, dst, static_cast<size_t>(n), S(stream));
}
int ngan_toim_bwd(const float* g_img, float gscale, const float* dyn, const float* img, const void* y, const float* r, const float* w,
                  void* ga, float* gpre, float* gw, int grad_accumulate, float* workspace, float leak, int B, int C,
                  int H, int W, void* stream) {
    NGAN_REQUIRE(g_img && img && y && w && !bad_c(C) && B > 0, "toim_bwd: bad arguments");
    NGAN_REQUIRE(!ga || r, "toim_bwd: ga requires r");
    NGAN_REQUIRE(workspace || !gw, "toim_bwd: the weight gradient needs a workspace");
    return toim_bwd(g_img, gscale, dyn, img, y, r, w, ga, gpre, gw, grad_accumulate, workspace, leak, B, C, H, W,
                    S(stream));
}
int ngan_head_fwd(const void* y, const float* w, const float* bias, float scale, float* score, int B, int C, int Sz,
                  void* stream) {
    NGAN_REQUIRE(y && w && bias && score && !bad_c(C) && B > 0, "head_fwd: bad arguments");
    return head_fwd(y, w, bias, scale, score, B, C, Sz, S(stream));
}
int ngan_head_bwd_pn(const float* gout, const float* w, float scale, const void* y, const float* r, void* ga,
                     void* gy_out, float leak, int B, int C, int Sz, void* stream) {
    NGAN_REQUIRE(gout && w && y && r && ga && !bad_c(C) && B > 0, "head_bwd_pn: bad arguments");
    return head_bwd_pn(gout, w, scale, y, r, ga, gy_out, leak, B, C, Sz, S(stream));
}
int ngan_head_wgrad(const void* t, const float* coeff, float scale, float* gw, float* gb, int accumulate, int B,
                    int C, int Sz, void* stream) {
    NGAN_REQUIRE(t && coeff && gw && !bad_c(C) && B > 0, "head_wgrad: bad arguments");
    return head_wgrad(t, coeff, scale, gw, gb, accumulate, B, C, Sz, S(stream));
}
int ngan_prep_linear_weight(const float* w, void* wb, int K, int C, int Sz, void* stream) {
    NGAN_REQUIRE(w && wb && K > 0 && C > 0 && Sz > 0, "prep_linear_weight: bad arguments");
    return prep_linear_weight(w, wb, K, C, Sz, S(stream));
}
long long ngan_linear_fwd_workspace_bytes(int B, int K) {
    return static_cast<long long>((B + 127) / 128 * 128) * K * 2;
}
int ngan_linear_fwd(const float* z, const void* wb, float scale, float leak, void* y, float* r, void* workspace, int B,
                    int K, int C, int Sz, void* stream) {
    NGAN_REQUIRE(z && wb && y && workspace && B > 0, "linear_fwd: bad arguments");
    return linear_fwd_pn(z, wb, scale, leak, y, r, workspace, B, K, C, Sz, S(stream));
}
int ngan_linear_wgrad(const void* ga, const float* z, float scale, float* dw, int accumulate, int B, int K, int C,
                      int Sz, void* stream) {
    NGAN_REQUIRE(ga && z && dw && B > 0, "linear_wgrad: bad arguments");
    return linear_wgrad(ga, z, scale, dw, accumulate, B, K, C, Sz, S(stream));
}
int ngan_wloss(const float* s_real, const float* s_fake, float drift, float* out3, float* g_real, float* g_fake,
               float gscale, int B, void* stream) {
    NGAN_REQUIRE(s_real && s_fake && out3 && B > 0, "wloss: bad arguments");
    return wloss_fwd(s_real, s_fake, drift, out3, g_real, g_fake, gscale, B, S(stream));
}
int ngan_gloss(const float* s_fake, float* out1, float* g_fake, float gscale, int B, void* stream) {
    NGAN_REQUIRE(s_fake && out1 && B > 0, "gloss: bad arguments");
    return gloss_fwd(s_fake, out1, g_fake, gscale, B, S(stream));
}
long long ngan_similarity_loss_workspace_bytes(int B, long long per_image) {
    return static_cast<long long>(similarity_workspace_bytes(B, per_image));
}
int ngan_similarity_loss(const float* images, const float* z, float lambda, float* workspace, float* out, int B,
                         long long per_image, int latent, void* stream) {
    NGAN_REQUIRE(images && z && workspace && out && B > 1 && per_image > 0 && latent > 0, "similarity_loss: bad arguments");
    return similarity_loss(images, z, lambda, workspace, out, B, per_image, latent, S(stream));
}
int ngan_pack_stats(const float* out3, const float* out1, const float* pen, float* stats, void* stream) {
    NGAN_REQUIRE(out3 && out1 && pen && stats, "pack_stats: null pointer");
    return pack_stats(out3, out1, pen, stats, S(stream));
}
long long ngan_gp_loss_workspace_bytes(int B) { return static_cast<long long>(B) * 64 * sizeof(float); }
int ngan_gp_loss(const float* g, float norm_scale, float lambda, float* pen, float* coeff, float gscale,
                 float* workspace, int B, long long per_sample, void* stream) {
    NGAN_REQUIRE(g && pen && coeff && workspace && B > 0 && per_sample > 0, "gp_loss: bad arguments");
    return gp_loss(g, norm_scale, lambda, pen, coeff, gscale, workspace, B, static_cast<size_t>(per_sample), S(stream));
}
int ngan_adam_multi(const ngan_adam_tensor* tensors, int n_tensors, float beta1, float beta2, float eps,
                    void* stream) {
    NGAN_REQUIRE(tensors || n_tensors == 0, "adam_multi: null table");
    return adam_multi_launch(reinterpret_cast<const AdamEntry*>(tensors), n_tensors, beta1, beta2, eps, S(stream));
}

int ngan_adam_linear_factored(float* p, float* m, float* v, void* shadow_img, float* g_out, const void* ga_c8,
                              const float* z, int Btot, int b_per_seg, long long ga_seg_stride, long long z_seg_stride,
                              int K, int C, int Sz, float gscale, float step_size, float inv_bc2_sqrt,
                              const float* dyn, float beta1, float beta2, float eps, void* stream) {
    NGAN_REQUIRE(ga_c8 && z && ((p && m && v) || (!p && g_out)), "adam_linear_factored: null pointer");
    return adam_linear_factored(p, m, v, shadow_img, g_out, ga_c8, z, Btot, b_per_seg, ga_seg_stride, z_seg_stride, K,
                                C, Sz * Sz, gscale, step_size, inv_bc2_sqrt, dyn, beta1, beta2, eps, S(stream));
}

/* ---- on-device image pipeline (data/NeuronDataset.py:170-205) ---- */
long long ngan_augment_workspace_bytes(int batch, int canvas, int crop) {
    return static_cast<long long>(augment_workspace_bytes(batch, canvas, crop));
}
int ngan_augment_batch(const float* canvases, const int* src_index, const float* params, const int* tap_first,
                       const int* tap_count, const float* tap_weight, int max_taps, float* workspace, float* out,
                       int batch, int canvas, int crop, int out_size, void* stream) {
    NGAN_REQUIRE(canvases && src_index && params && tap_first && tap_count && tap_weight && workspace && out,
                 "augment_batch: null pointer");
    NGAN_REQUIRE(batch > 0 && batch <= 65535 && canvas > 0 && canvas <= 32768 && crop > 0 && crop <= canvas &&
                     out_size > 0 && out_size <= crop && max_taps > 0,
                 "augment_batch: bad sizes");
    return augment_batch(canvases, src_index, params, tap_first, tap_count, tap_weight, max_taps, workspace, out,
                         batch, canvas, crop, out_size, S(stream));
}

}  // extern "C"
