// Internal (C++) launcher declarations shared between the .cu files; the public C ABI is include/ngan_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ngan {

enum ConvEpilogue { EPI_FWD_PN = 0, EPI_LINEAR = 1, EPI_BWD_PN = 2, EPI_DBL = 3 };

// conv3x3_umma.cu
int conv3x3_dispatch(int epi, const void* x, const void* wprep, int B, int cin, int cout, int H, int W, float scale,
                     float leak, const float* bias, void* out0, void* out1, float* rout, const void* y,
                     const float* r, const void* gy, const void* addin, cudaStream_t st, const float* toim_w = nullptr,
                     float* img_out = nullptr, int img_bf16 = 0);
int prep_conv_weight(const float* w, void* fwd, void* dgrad, int cin, int cout, cudaStream_t st);

// wgrad.cu
size_t conv3x3_wgrad_workspace_bytes(int B, int cin, int cout, int H, int W);
int conv3x3_wgrad(const void* x, const void* ga, float scale, float* dw, int accumulate, float* workspace, int B,
                  int cin, int cout, int H, int W, cudaStream_t st);

// elementwise.cu
int reduce_partials(const float* partials, int n_partials, long long n, long long ld, float scale, float* out,
                    int accumulate, cudaStream_t st, float* out2 = nullptr, long long n_split = 0);
int sum_slots(const float* slots, int n_slots, long long n, long long ld, float scale, float* out, cudaStream_t st);
size_t pixel_reduction_workspace_bytes(int B, int C, int H, int W);
int nchw_to_c8(const float* src, void* dst, int B, int C, int H, int W, cudaStream_t st);
int c8_to_nchw(const void* src, float* dst, int B, int C, int H, int W, cudaStream_t st);
int upsample2x_c8(const void* x, void* out, int B, int C, int H, int W, cudaStream_t st);
int avgpool2_c8(const void* x, void* out, int B, int C, int H, int W, cudaStream_t st);
int pn_bwd_c8(const void* g, int unpool, float gscale, const float* dyn, const void* y, const float* r, const void* addin, void* ga,
              void* gy_out, float leak, int B, int C, int H, int W, cudaStream_t st);
int up2_bwd_pn_bwd_c8(const void* g_up, const void* y, const float* r, const float* extra_pre, const float* extra_w,
                      void* ga, float leak, int B, int C, int H, int W, cudaStream_t st);
int pool_image(const float* x, float* out, int B, int H, int W, cudaStream_t st);
int unpool_image(const float* g, float* out, float scale, int B, int H, int W, cudaStream_t st);
int up2_image(const float* x, float* out, int B, int H, int W, cudaStream_t st);
int up2_image_bwd(const float* g, float* out, float scale, const float* dyn, int B, int H, int W, cudaStream_t st);
int lerp_f32(const float* a, const float* b, float alpha, const float* dyn, float* out, size_t n, cudaStream_t st);
int axpby_f32(const float* a, float ca, const float* b, float cb, float* out, size_t n, cudaStream_t st);
int interp_images(const float* real, const float* fake, const float* eps, float* out, int B, size_t per_sample,
                  cudaStream_t st);
int fromim_fwd(const float* xp, const float* w, const float* b, void* out, int B, int C, int H, int W,
               cudaStream_t st);
int d_fade_fwd(const void* y_end, const float* xp, const float* w_old, const float* b_old, float alpha,
               const float* dyn, void* out, int B, int C, int H, int W, cudaStream_t st);
int fromim_bwd(const void* g, int unpool, float gscale, const float* dyn, const float* xp, const float* w, float* gw, float* gb,
               int grad_accumulate, float* workspace, float* g_img, int g_img_accumulate, int B, int C, int H, int W,
               cudaStream_t st);
int fromim_dbl(const float* ghat_xp, float in_scale, const void* g, int unpool, float gscale, const float* dyn,
               const float* w,
               void* ghat_out, float* what, int grad_accumulate, float* workspace, int B, int C, int H, int W,
               cudaStream_t st);
int toim_fwd(const void* y, const float* w, float* img, int B, int C, int H, int W, cudaStream_t st);
int f32_to_bf16(const float* src, void* dst, size_t n, cudaStream_t st);
int toim_bwd(const float* g_img, float gscale, const float* dyn, const float* img, const void* y, const float* r, const float* w,
             void* ga, float* gpre, float* gw, int grad_accumulate, float* workspace, float leak, int B, int C, int H,
             int W, cudaStream_t st);
int head_fwd(const void* y, const float* w, const float* bias, float scale, float* score, int B, int C, int S,
             cudaStream_t st);
int head_bwd_pn(const float* gout, const float* w, float scale, const void* y, const float* r, void* ga,
                void* gy_out, float leak, int B, int C, int S, cudaStream_t st);
int head_wgrad(const void* t, const float* coeff, float scale, float* gw, float* gb, int accumulate, int B, int C,
               int S, cudaStream_t st);
int bias_grad_c8(const void* ga, float* gb, int accumulate, float* workspace, int B, int C, int H, int W,
                 cudaStream_t st);
int wloss_fwd(const float* s_real, const float* s_fake, float drift, float* out3, float* g_real, float* g_fake,
              float gscale, int B, cudaStream_t st);
int gloss_fwd(const float* s_fake, float* out1, float* g_fake, float gscale, int B, cudaStream_t st);
int gp_loss(const float* g, float norm_scale, float lambda, float* pen_out, float* coeff_out, float gscale,
            float* workspace, int B, size_t per_sample, cudaStream_t st);
size_t similarity_workspace_bytes(int B, long long P);
int similarity_loss(const float* x, const float* z, float lambda, float* workspace, float* out, int B, long long P,
                    int L, cudaStream_t st);
int pack_stats(const float* out3, const float* out1, const float* pen, float* stats, cudaStream_t st);
int scale_rows_f32(const float* x, const float* coeff, float scale, float* out, int B, size_t per_sample,
                   cudaStream_t st);

// linear.cu
int prep_linear_weight(const float* w, void* wb, int K, int C, int S, cudaStream_t st);
int linear_fwd_pn(const float* z, const void* wb, float scale, float leak, void* y, float* r, void* workspace, int B,
                  int K, int C, int S, cudaStream_t st);
int linear_wgrad(const void* ga, const float* z, float scale, float* dw, int accumulate, int B, int K, int C, int S,
                 cudaStream_t st);

// adam.cu (AdamEntry has the layout of ngan_adam_tensor in include/ngan_b200.h)
struct AdamEntry;
int adam_linear_factored(float* p, float* m, float* v, void* shadow, float* g_out, const void* ga, const float* z,
                         int Btot, int b_per_seg, long long ga_seg_stride, long long z_seg_stride, int K, int C, int SS,
                         float gscale, float step_size, float inv_bc2_sqrt, const float* dyn, float beta1, float beta2,
                         float eps, cudaStream_t st);
int adam_multi_launch(const AdamEntry* entries, int n_tensors, float beta1, float beta2, float eps, cudaStream_t st);

// augment.cu
size_t augment_workspace_bytes(int batch, int canvas, int crop);
int augment_batch(const float* canvases, const int* src_index, const float* params, const int* tap_first,
                  const int* tap_count, const float* tap_weight, int max_taps, float* workspace, float* out, int batch,
                  int canvas, int crop, int out_size, cudaStream_t st);

}  // namespace ngan
