// Argument block shared by the 3x3 convolution kernels (conv3x3_umma.cu: per-tap kernel for wide layers,
// conv3x3_fold.cu: persistent kx-folded kernel for layers with CIN*COUT <= 4096).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ngan {

struct ConvArgs {
    int B, H, W;
    int TH, TW, Wh;
    int nMT;
    int tmem_cols;
    int n_stage;
    int n_acc;   // accumulator buffers in TMEM (persistent kernel)
    long long* dbg_clock;   // timing experiments only: per-tile clock64() trace of CTA 0, or null
    int debug;   // timing experiments only (NGAN_CONV_DEBUG): 1 = skip MMAs, 2 = skip epilogue, 4 = skip input loads
    int tiles_x, tiles_y, n_tiles;  // persistent kernel: tile grid over (B, rows, cols)
    float inv_tiles_x, inv_tiles_per_img;
    uint32_t plane_bytes;  // (TH+2)*Wh*16
    float scale, leak;
    const __nv_bfloat16* wprep;  // per-tap image [9][CIN/8][COUT][8] or folded image [3][CIN/8][3*COUT][8]
    const float* bias;           // FWD: [COUT] or null
    __nv_bfloat16* out0;
    __nv_bfloat16* out1;
    float* rout;
    const float* toim_w;         // FWD (folded kernel): fused ToImage, img = tanh(sum_c toim_w[c] * y[c]) ...
    float* img_out;              // ... written here ([B][H][W] fp32); out0 may then be null (y not stored)
    int img_bf16;                // img_out holds bf16 instead of fp32 (generator-only inference, BASELINE config 5)
    const __nv_bfloat16* y;
    const float* r;
    const __nv_bfloat16* gy;
    const __nv_bfloat16* addin;
};


// channel pairs small enough for all nine weight slabs to stay in shared memory use the folded kernel
inline bool conv_uses_folded_kernel(int cin, int cout) { return cin * cout <= 4096; }

// Element i of the fp32 master weight [COUT][CIN][3][3] -> its place in the two bf16 UMMA operand images
//   per-tap:  fwd [9][CIN/8][COUT][8],  dgrad [9][COUT/8][CIN][8] with the taps flipped (dx = convT(g, W))
//   folded:   fwd [3 ky][CIN/8][3*COUT (kx, co)][8],  dgrad [3][COUT/8][3*CIN][8]   (conv3x3_fold.cu)
// Used by the preparation kernel and by the Adam kernel, which refreshes the images in the pass that updates
// the master.
__device__ __forceinline__ void conv_image_store(__nv_bfloat16 v, int i, int cin, int cout, int folded,
                                                 __nv_bfloat16* __restrict__ fwd, __nv_bfloat16* __restrict__ dgrad) {
    const int tap = i % 9;
    const int ci = (i / 9) % cin;
    const int co = i / (9 * cin);
    if (!folded) {
        if (fwd) fwd[((static_cast<size_t>(tap) * (cin / 8) + ci / 8) * cout + co) * 8 + (ci % 8)] = v;
        if (dgrad) dgrad[((static_cast<size_t>(8 - tap) * (cout / 8) + co / 8) * cin + ci) * 8 + (co % 8)] = v;
    } else {
        const int ky = tap / 3, kx = tap % 3;
        if (fwd) fwd[((static_cast<size_t>(ky) * (cin / 8) + ci / 8) * (3 * cout) + kx * cout + co) * 8 + (ci % 8)] = v;
        if (dgrad)
            dgrad[((static_cast<size_t>(2 - ky) * (cout / 8) + co / 8) * (3 * cin) + (2 - kx) * cin + ci) * 8 +
                  (co % 8)] = v;
    }
}

int make_c8_tensor_map(CUtensorMap* map, const void* base, int B, int C, int H, int W, int box_w, int box_h,
                       int box_planes);
int conv3x3_fold_dispatch(int epi, const void* x, ConvArgs a, int B, int cin, int cout, int H, int W,
                          cudaStream_t st);

}  // namespace ngan
