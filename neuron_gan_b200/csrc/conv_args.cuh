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
    const __nv_bfloat16* y;
    const float* r;
    const __nv_bfloat16* gy;
    const __nv_bfloat16* addin;
};


// channel pairs small enough for all nine weight slabs to stay in shared memory use the folded kernel
inline bool conv_uses_folded_kernel(int cin, int cout) { return cin * cout <= 4096; }

int make_c8_tensor_map(CUtensorMap* map, const void* base, int B, int C, int H, int W, int box_w, int box_h,
                       int box_planes);
int conv3x3_fold_dispatch(int epi, const void* x, ConvArgs a, int B, int cin, int cout, int H, int W,
                          cudaStream_t st);

}  // namespace ngan
