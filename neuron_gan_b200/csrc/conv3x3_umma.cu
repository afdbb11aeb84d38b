// 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on the 5th-gen tensor cores (tcgen05.mma,
// accumulators in TMEM), fed by TMA, with the layer's pointwise tail fused into the epilogue.
//
// Replaces, for neuron-gan's Conv2d_normalized (reference models.py:172-204) and what autograd derives
// from it: the forward conv + weight_scale mul + LeakyReLU + PixelNorm (models.py:203-204, 263-268,
// 110-126), the data-gradient conv (with the PixelNorm/LeakyReLU backward of the producing layer fused),
// and the "conv of the cotangent" sweep of the gradient-penalty double backward (loss_functions.py:175).
//
// GEMM view: M = output pixels, N = COUT, K = 9 taps x CIN.  One CTA owns a TH x TW spatial tile of one
// sample.  A single TMA box copy brings the (TH+2) x (TW+2) haloed input tile (all CIN/8 channel-group
// planes, zero-filled outside the image) into shared memory; because the tensor is stored C8-planar
// (common.cuh) the tile *is* a no-swizzle K-major UMMA operand whose rows are the flattened haloed pixel
// index q.  Output row m = r*(TW+2)+c reads q = m + ky*(TW+2) + kx for tap (ky,kx): each tap is the same
// descriptor with a different 16-byte-granular start address, so the tile is staged once and read by
// nine shifted MMAs.  Rows with c >= TW are halo columns; they are computed and dropped.
// All M-tiles of the CTA keep their accumulators in TMEM at once (nMT x COUT fp32 columns), the tap loop
// is outermost and the per-tap weight slab [CIN/8][COUT][8] streams through a small ring of bulk copies.
// Several CTAs are resident per SM (<= 256 TMEM columns each), which overlaps one CTA's epilogue with
// another's loads and MMAs.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ngan {

struct ConvArgs {
    int B, H, W;
    int TH, TW, Wh;
    int nMT;
    int tmem_cols;
    int n_stage;
    int tiles_x, tiles_y, n_tiles;  // persistent kernel: tile grid over (B, rows, cols)
    uint32_t plane_bytes;  // (TH+2)*Wh*16
    float scale, leak;
    const __nv_bfloat16* wprep;  // [9][CIN/8][COUT][8]
    const float* bias;           // FWD: [COUT] or null
    __nv_bfloat16* out0;
    __nv_bfloat16* out1;
    float* rout;
    const __nv_bfloat16* y;
    const float* r;
    const __nv_bfloat16* gy;
    const __nv_bfloat16* addin;
};

constexpr int kMaxStages = 9;

// Epilogue of one CTA tile: thread (quad, lane) owns accumulator row quad*32+lane of every M-tile, i.e. one output
// pixel with all COUT channels in its TMEM lane.  `quad` must equal (warp index % 4): a warp can only read the
// TMEM lane quadrant it is bound to.
template <int COUT, int EPI>
__device__ __forceinline__ void conv_epilogue(const ConvArgs& a, uint32_t tmem_acc, int quad, int lane, int b,
                                              int tile_x, int tile_y) {
    const size_t HW = static_cast<size_t>(a.H) * a.W;
    constexpr int NCH = COUT / 8;
    const float inv_c = 1.0f / COUT;

    for (int mt = 0; mt < a.nMT; ++mt) {
        const int m = mt * 128 + quad * 32 + lane;
        const int rr = m / a.Wh, cc = m - rr * a.Wh;
        const int oy = tile_y * a.TH + rr, ox = tile_x * a.TW + cc;
        const bool valid = (cc < a.TW) && (rr < a.TH) && (oy < a.H) && (ox < a.W);
        const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + mt * COUT;
        // uint4 index of channel-group 0 of this pixel in a [B][COUT/8][H][W][8] tensor
        const size_t q0 = static_cast<size_t>(b) * NCH * HW + static_cast<size_t>(oy) * a.W + ox;
        const size_t p0 = static_cast<size_t>(b) * HW + static_cast<size_t>(oy) * a.W + ox;
        float v[16];

        if constexpr (EPI == EPI_FWD_PN) {
            float ss = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                tmem_ld16(taddr + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = a.scale * v[i] + (a.bias ? __ldg(a.bias + c0 + i) : 0.f);
                    x = x > 0.f ? x : a.leak * x;
                    ss += x * x;
                }
            }
            const float rinv = rsqrtf(ss * inv_c + 1e-8f);
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                tmem_ld16(taddr + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x = a.scale * v[i] + (a.bias ? __ldg(a.bias + c0 + i) : 0.f);
                    x = x > 0.f ? x : a.leak * x;
                    v[i] = x * rinv;
                }
                if (valid) {
                    uint4* o = reinterpret_cast<uint4*>(a.out0);
                    o[q0 + (c0 / 8) * HW] = pack8(v);
                    o[q0 + (c0 / 8 + 1) * HW] = pack8(v + 8);
                }
            }
            if (valid && a.rout) a.rout[p0] = rinv;
        } else if constexpr (EPI == EPI_LINEAR) {
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                tmem_ld16(taddr + c0, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] *= a.scale;
                if (valid) {
                    uint4* o = reinterpret_cast<uint4*>(a.out0);
                    o[q0 + (c0 / 8) * HW] = pack8(v);
                    o[q0 + (c0 / 8 + 1) * HW] = pack8(v + 8);
                }
            }
        } else if constexpr (EPI == EPI_BWD_PN) {
            // g = scale*acc is the gradient wrt y (PixelNorm output of the layer that produced this conv's
            // input); emit ga = mask(y) * r * (g - y*mean_c(g*y)) (+ addin)   [SURVEY.md 8a row 3]
            const uint4* yq = reinterpret_cast<const uint4*>(a.y);
            float t = 0.f, yv[16];
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                tmem_ld16(taddr + c0, v);
                uint4 y0 = valid ? __ldg(yq + q0 + (c0 / 8) * HW) : make_uint4(0, 0, 0, 0);
                uint4 y1 = valid ? __ldg(yq + q0 + (c0 / 8 + 1) * HW) : make_uint4(0, 0, 0, 0);
                unpack8(y0, yv);
                unpack8(y1, yv + 8);
#pragma unroll
                for (int i = 0; i < 16; ++i) t += a.scale * v[i] * yv[i];
            }
            t *= inv_c;
            const float rinv = valid ? __ldg(a.r + p0) : 0.f;
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                tmem_ld16(taddr + c0, v);
                uint4 y0 = valid ? __ldg(yq + q0 + (c0 / 8) * HW) : make_uint4(0, 0, 0, 0);
                uint4 y1 = valid ? __ldg(yq + q0 + (c0 / 8 + 1) * HW) : make_uint4(0, 0, 0, 0);
                unpack8(y0, yv);
                unpack8(y1, yv + 8);
                float g[16], ad[16];
                if (a.addin && valid) {
                    const uint4* aq = reinterpret_cast<const uint4*>(a.addin);
                    unpack8(__ldg(aq + q0 + (c0 / 8) * HW), ad);
                    unpack8(__ldg(aq + q0 + (c0 / 8 + 1) * HW), ad + 8);
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) ad[i] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    g[i] = a.scale * v[i];
                    v[i] = lrelu_mask(yv[i], a.leak) * rinv * (g[i] - yv[i] * t) + ad[i];
                }
                if (valid) {
                    uint4* o = reinterpret_cast<uint4*>(a.out0);
                    o[q0 + (c0 / 8) * HW] = pack8(v);
                    o[q0 + (c0 / 8 + 1) * HW] = pack8(v + 8);
                    if (a.out1) {
                        uint4* o1 = reinterpret_cast<uint4*>(a.out1);
                        o1[q0 + (c0 / 8) * HW] = pack8(g);
                        o1[q0 + (c0 / 8 + 1) * HW] = pack8(g + 8);
                    }
                }
            }
        } else {  // EPI_DBL
            // Double backward of lrelu+PixelNorm.  acc*scale = cotangent on ga (pre-activation grad);
            // gh^ = mask*that.  With t = mean(gy*y), u = mean(gh^*y), w = mean(gh^*gy):
            //   out0 = r*(gh^ - y*u)                                  (cotangent on gy, next layer's input)
            //   out1 = mask * (-r^2) * (t*gh^ + u*gy + (w - 3ut)*y)   (cotangent injected at the pre-activation)
            const uint4* yq = reinterpret_cast<const uint4*>(a.y);
            const uint4* gq = reinterpret_cast<const uint4*>(a.gy);
            float t = 0.f, u = 0.f, w = 0.f, yv[16], gv[16];
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                tmem_ld16(taddr + c0, v);
                const uint4 z4 = make_uint4(0, 0, 0, 0);
                unpack8(valid ? __ldg(yq + q0 + (c0 / 8) * HW) : z4, yv);
                unpack8(valid ? __ldg(yq + q0 + (c0 / 8 + 1) * HW) : z4, yv + 8);
                unpack8(valid ? __ldg(gq + q0 + (c0 / 8) * HW) : z4, gv);
                unpack8(valid ? __ldg(gq + q0 + (c0 / 8 + 1) * HW) : z4, gv + 8);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float gh = lrelu_mask(yv[i], a.leak) * a.scale * v[i];
                    t += gv[i] * yv[i];
                    u += gh * yv[i];
                    w += gh * gv[i];
                }
            }
            t *= inv_c;
            u *= inv_c;
            w *= inv_c;
            const float rinv = valid ? __ldg(a.r + p0) : 0.f;
            const float k3 = w - 3.f * u * t;
#pragma unroll
            for (int c0 = 0; c0 < COUT; c0 += 16) {
                tmem_ld16(taddr + c0, v);
                const uint4 z4 = make_uint4(0, 0, 0, 0);
                unpack8(valid ? __ldg(yq + q0 + (c0 / 8) * HW) : z4, yv);
                unpack8(valid ? __ldg(yq + q0 + (c0 / 8 + 1) * HW) : z4, yv + 8);
                unpack8(valid ? __ldg(gq + q0 + (c0 / 8) * HW) : z4, gv);
                unpack8(valid ? __ldg(gq + q0 + (c0 / 8 + 1) * HW) : z4, gv + 8);
                float o0[16], o1[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float mk = lrelu_mask(yv[i], a.leak);
                    const float gh = mk * a.scale * v[i];
                    o0[i] = rinv * (gh - yv[i] * u);
                    o1[i] = -mk * rinv * rinv * (t * gh + u * gv[i] + k3 * yv[i]);
                }
                if (valid) {
                    uint4* p0q = reinterpret_cast<uint4*>(a.out0);
                    uint4* p1q = reinterpret_cast<uint4*>(a.out1);
                    p0q[q0 + (c0 / 8) * HW] = pack8(o0);
                    p0q[q0 + (c0 / 8 + 1) * HW] = pack8(o0 + 8);
                    p1q[q0 + (c0 / 8) * HW] = pack8(o1);
                    p1q[q0 + (c0 / 8 + 1) * HW] = pack8(o1 + 8);
                }
            }
        }
    }

}

template <int CIN, int COUT, int EPI>
__global__ void __launch_bounds__(128) conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tmap, const ConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    constexpr uint32_t W_TAP_BYTES = CIN * COUT * 2;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, COUT);

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const uint32_t in_bytes = (CIN / 8) * a.plane_bytes;
    uint8_t* s_in = smem;
    uint8_t* s_w = smem + ((in_bytes + 256 + 127) & ~127u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + a.n_stage * W_TAP_BYTES);
    uint64_t* bar_in = bars;
    uint64_t* bar_mma = bars + 1;
    uint64_t* bar_wfull = bars + 2;
    uint64_t* bar_wempty = bars + 2 + kMaxStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * kMaxStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_x = blockIdx.x, tile_y = blockIdx.y, b = blockIdx.z;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmap);
        mbar_init(bar_in, 1);
        mbar_init(bar_mma, 1);
        for (int s = 0; s < a.n_stage; ++s) {
            mbar_init(bar_wfull + s, 1);
            mbar_init(bar_wempty + s, 1);
        }
        mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0 && lane == 0) {
        // ---- input tile: one TMA box {2*Wh x (TH+2) x CIN/8 x 1} of 8-byte elements
        mbar_arrive_expect_tx(bar_in, in_bytes);
        tma_load_4d(s_in, &tmap, bar_in, (tile_x * a.TW - 1) * 2, tile_y * a.TH - 1, 0, b);
        mbar_wait(bar_in, 0);
        tc_fence_after();
        const uint32_t in_base = smem_u32(s_in);
        const uint32_t w_base = smem_u32(s_w);
        for (int tap = 0; tap < 9; ++tap) {
            const int stage = tap % a.n_stage;
            mbar_wait(bar_wfull + stage, (tap / a.n_stage) & 1);
            tc_fence_after();
            const uint32_t tap_off = ((tap / 3) * a.Wh + (tap % 3)) * 16;
#pragma unroll
            for (int kc = 0; kc < CIN / 16; ++kc) {
                const uint32_t a_addr = in_base + (2 * kc) * a.plane_bytes + tap_off;
                const uint32_t b_addr = w_base + stage * W_TAP_BYTES + (2 * kc) * COUT * 16;
                const uint64_t bdesc = umma_desc(b_addr, COUT * 16, 128);
                for (int mt = 0; mt < a.nMT; ++mt) {
                    const uint64_t adesc = umma_desc(a_addr + mt * 128 * 16, a.plane_bytes, 128);
                    umma_bf16(tmem_base + mt * COUT, adesc, bdesc, IDESC, (tap | kc) != 0);
                }
            }
            umma_commit(bar_wempty + stage);  // slab free once these MMAs retire
        }
        umma_commit(bar_mma);
    } else if (warp == 1 && lane == 0) {
        // ---- weight slabs, one per tap, through a ring of n_stage buffers
        for (int tap = 0; tap < 9; ++tap) {
            const int stage = tap % a.n_stage;
            if (tap >= a.n_stage) mbar_wait(bar_wempty + stage, ((tap / a.n_stage) - 1) & 1);
            mbar_arrive_expect_tx(bar_wfull + stage, W_TAP_BYTES);
            bulk_load_1d(s_w + stage * W_TAP_BYTES, reinterpret_cast<const uint8_t*>(a.wprep) + tap * W_TAP_BYTES,
                         W_TAP_BYTES, bar_wfull + stage);
        }
    }

    // ---- epilogue: every thread owns one accumulator row (= one output pixel, all COUT channels)
    mbar_wait(bar_mma, 0);
    __syncwarp();
    tc_fence_after();

    conv_epilogue<COUT, EPI>(a, tmem_base, warp, lane, b, tile_x, tile_y);

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}


// ------------------------------------------------------------------------------------------------------------
// Persistent, warp-specialised variant for layers whose nine weight slabs fit in shared memory together
// (CIN*COUT <= 4096): one CTA per SM slot loops over tiles; warp 0 streams haloed input tiles through a ring
// of TMA stages, warp 1 issues the MMAs into one of two TMEM accumulator buffers, warps 2-5 drain the other
// buffer through the fused epilogue.  Load, MMA and epilogue of consecutive tiles overlap.
constexpr int kPersistThreads = 192;

template <int CIN, int COUT, int EPI>
__global__ void __launch_bounds__(kPersistThreads) conv3x3_umma_persist_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                               const ConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    constexpr uint32_t W_BYTES = 9 * CIN * COUT * 2;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, COUT);

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const uint32_t in_bytes = (CIN / 8) * a.plane_bytes;
    const uint32_t stage_stride = (in_bytes + 256 + 127) & ~127u;
    // Input stages first, weights after them: the last M-tile of a stage reads up to ~2 KB past the tile it was
    // given (rows that are dropped in the epilogue); that over-read must stay inside this CTA's allocation.
    uint8_t* s_in = smem;
    uint8_t* s_w = smem + a.n_stage * stage_stride;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + ((W_BYTES + 127) & ~127u));
    uint64_t* bar_w = bars;
    uint64_t* bar_full = bars + 1;               // [n_stage]
    uint64_t* bar_empty = bars + 1 + 4;          // [n_stage]
    uint64_t* bar_acc_full = bars + 1 + 8;       // [2]
    uint64_t* bar_acc_empty = bars + 1 + 10;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t buf_cols = a.tmem_cols / 2;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmap);
        mbar_init(bar_w, 1);
        for (int s = 0; s < a.n_stage; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(bar_acc_full + s, 1);
            mbar_init(bar_acc_empty + s, 4);     // one arrival per epilogue warp
        }
        mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int tiles_per_img = a.tiles_x * a.tiles_y;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_w, W_BYTES);
            bulk_load_1d(s_w, a.wprep, W_BYTES, bar_w);
            int it = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
                const int stage = it % a.n_stage;
                if (it >= a.n_stage) mbar_wait(bar_empty + stage, ((it / a.n_stage) - 1) & 1);
                const int b = tile / tiles_per_img, t2 = tile - b * tiles_per_img;
                const int tile_y = t2 / a.tiles_x, tile_x = t2 - tile_y * a.tiles_x;
                mbar_arrive_expect_tx(bar_full + stage, in_bytes);
                tma_load_4d(s_in + stage * stage_stride, &tmap, bar_full + stage, (tile_x * a.TW - 1) * 2,
                            tile_y * a.TH - 1, 0, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(bar_w, 0);
            const uint32_t w_base = smem_u32(s_w);
            int it = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
                const int stage = it % a.n_stage, buf = it & 1;
                if (it >= 2) mbar_wait(bar_acc_empty + buf, ((it >> 1) - 1) & 1);
                mbar_wait(bar_full + stage, (it / a.n_stage) & 1);
                tc_fence_after();
                const uint32_t in_base = smem_u32(s_in + stage * stage_stride);
                const uint32_t acc = tmem_base + buf * buf_cols;
#pragma unroll 1
                for (int tap = 0; tap < 9; ++tap) {
                    const uint32_t tap_off = ((tap / 3) * a.Wh + (tap % 3)) * 16;
#pragma unroll
                    for (int kc = 0; kc < CIN / 16; ++kc) {
                        const uint32_t a_addr = in_base + (2 * kc) * a.plane_bytes + tap_off;
                        const uint64_t bdesc =
                            umma_desc(w_base + (tap * (CIN / 8) + 2 * kc) * COUT * 16, COUT * 16, 128);
                        for (int mt = 0; mt < a.nMT; ++mt) {
                            const uint64_t adesc = umma_desc(a_addr + mt * 128 * 16, a.plane_bytes, 128);
                            umma_bf16(acc + mt * COUT, adesc, bdesc, IDESC, (tap | kc) != 0);
                        }
                    }
                }
                umma_commit(bar_empty + stage);      // input stage reusable once these MMAs retire
                umma_commit(bar_acc_full + buf);     // accumulators ready for the epilogue
            }
        }
    } else {
        const int quad = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            const int buf = it & 1;
            const int b = tile / tiles_per_img, t2 = tile - b * tiles_per_img;
            const int tile_y = t2 / a.tiles_x, tile_x = t2 - tile_y * a.tiles_x;
            mbar_wait(bar_acc_full + buf, (it >> 1) & 1);
            __syncwarp();
            tc_fence_after();
            conv_epilogue<COUT, EPI>(a, tmem_base + buf * buf_cols, quad, lane, b, tile_x, tile_y);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_acc_empty + buf);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------ host side

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// TMA map over a C8-planar tensor [B][C/8][H][W][8] bf16 seen as 8-byte elements: dims {2W, H, C/8, B};
// the box is one haloed tile {2*box_w, box_h, planes, 1}.  Out-of-image elements read as zero.
int make_c8_tensor_map(CUtensorMap* map, const void* base, int B, int C, int H, int W, int box_w, int box_h,
                       int box_planes) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return NGAN_ERR_CUDA;
    }
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(2 * W), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(C / 8),
                          static_cast<cuuint64_t>(B)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 16, static_cast<cuuint64_t>(H) * W * 16,
                             static_cast<cuuint64_t>(C / 8) * H * W * 16};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(2 * box_w), static_cast<cuuint32_t>(box_h),
                         static_cast<cuuint32_t>(box_planes), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    if (box[0] > 256 || box[1] > 256 || box[2] > 256) {
        set_error("TMA box too large (%u,%u,%u)", box[0], box[1], box[2]);
        return NGAN_ERR_INVALID;
    }
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for tensor B=%d C=%d H=%d W=%d box=%dx%dx%d", (int)r, B, C, H,
                  W, box_w, box_h, box_planes);
        return NGAN_ERR_CUDA;
    }
    return NGAN_OK;
}

struct TilePlan {
    int TH, TW, Wh, nMT, tmem_cols, n_stage;
    uint32_t plane_bytes, smem_bytes;
};

static int pow2_cols(int n) {
    int c = 32;
    while (c < n) c <<= 1;
    return c;
}

static bool plan_tiles(int B, int H, int W, int cin, int cout, TilePlan* p) {
    const int kSmemMax = 220 * 1024, kSmemSoft = 108 * 1024;
    const int w_tap = cin * cout * 2;
    int TW = W < 64 ? W : 64;
    int Wh = TW + 2;
    const int tiles_x = (W + TW - 1) / TW;
    int mt_cap = 256 / cout;
    if (mt_cap < 1) mt_cap = 1;
    if (mt_cap > 8) mt_cap = 8;
    bool have = false;
    // Largest tile first; shrink the tile (fewer M-tiles per CTA) while the grid would leave SMs idle.
    for (int mt = mt_cap; mt >= 1; --mt) {
        int th_cap = (mt * 128) / Wh;
        if (th_cap < 1) th_cap = 1;
        if (th_cap > H) th_cap = H;
        const int n_tiles = (H + th_cap - 1) / th_cap;
        const int TH = (H + n_tiles - 1) / n_tiles;
        const uint32_t plane = static_cast<uint32_t>(TH + 2) * Wh * 16;
        const uint32_t in_bytes = ((cin / 8) * plane + 256 + 127) & ~127u;
        const uint32_t fixed = in_bytes + 128 /*align slack*/ + (2 + 2 * kMaxStages) * 8 + 16;
        const int budget = (fixed + 2 * w_tap <= (uint32_t)kSmemSoft) ? kSmemSoft : kSmemMax;
        int stages = (budget - (int)fixed) / w_tap;
        if (stages > 9) stages = 9;
        const int nMT = (TH * Wh + 127) / 128;
        const int cols = pow2_cols(nMT * cout);
        if (stages < 2 || cols > 512) continue;
        p->TH = TH;
        p->TW = TW;
        p->Wh = Wh;
        p->nMT = nMT;
        p->tmem_cols = cols;
        p->n_stage = stages;
        p->plane_bytes = plane;
        p->smem_bytes = fixed + stages * w_tap;
        have = true;
        if (static_cast<long long>(B) * n_tiles * tiles_x >= 148) break;
    }
    return have;
}

static bool plan_persistent(int B, int H, int W, int cin, int cout, TilePlan* p, int* n_ctas) {
    const uint32_t w_bytes = ((9u * cin * cout * 2) + 127) & ~127u;
    if (cin * cout > 4096) return false;
    const int TW = W < 64 ? W : 64, Wh = TW + 2;
    const int tiles_x = (W + TW - 1) / TW;
    int mt_cap = 128 / cout;                 // two accumulator buffers of <= 128 columns each
    if (mt_cap < 1) mt_cap = 1;
    if (mt_cap > 8) mt_cap = 8;
    for (int mt = mt_cap; mt >= 1; --mt) {
        int th_cap = (mt * 128) / Wh;
        if (th_cap < 1) th_cap = 1;
        if (th_cap > H) th_cap = H;
        const int n_ty = (H + th_cap - 1) / th_cap;
        const int TH = (H + n_ty - 1) / n_ty;
        const uint32_t plane = static_cast<uint32_t>(TH + 2) * Wh * 16;
        const uint32_t stage = ((cin / 8) * plane + 256 + 127) & ~127u;
        const long long n_tiles = static_cast<long long>(B) * n_ty * tiles_x;
        if (n_tiles < 2 * 148 && mt > 1) continue;       // keep the tiles small enough to give every SM work
        for (int stages = 3; stages >= 2; --stages) {
            const uint32_t total = 128 + w_bytes + stages * stage + 17 * 8 + 16;
            if (total > (stages == 3 ? 112u : 220u) * 1024) continue;
            p->TH = TH; p->TW = TW; p->Wh = Wh;
            p->nMT = (TH * Wh + 127) / 128;
            p->tmem_cols = pow2_cols(2 * pow2_cols(p->nMT * cout) < 64 ? 64 : 2 * pow2_cols(p->nMT * cout));
            p->n_stage = stages;
            p->plane_bytes = plane;
            p->smem_bytes = total;
            const int per_sm = total <= 112 * 1024 && p->tmem_cols <= 256 ? 2 : 1;
            *n_ctas = static_cast<int>(n_tiles < 148 * per_sm ? n_tiles : 148 * per_sm);
            return true;
        }
    }
    return false;
}

template <int CIN, int COUT, int EPI>
static int launch_conv_persist(const CUtensorMap& tmap, const ConvArgs& a, const TilePlan& p, int n_ctas,
                               cudaStream_t st) {
    auto kern = conv3x3_umma_persist_kernel<CIN, COUT, EPI>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv3x3 persistent)");
        configured = true;
    }
    kern<<<n_ctas, kPersistThreads, p.smem_bytes, st>>>(tmap, a);
    return check_launch("conv3x3_umma_persist");
}

template <int CIN, int COUT, int EPI>
static int launch_conv(const CUtensorMap& tmap, const ConvArgs& a, const TilePlan& p, cudaStream_t st) {
    auto kern = conv3x3_umma_kernel<CIN, COUT, EPI>;
    static int configured = 0;
    if (configured < (int)p.smem_bytes) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv3x3)");
        configured = 227 * 1024;
    }
    dim3 grid((a.W + p.TW - 1) / p.TW, (a.H + p.TH - 1) / p.TH, a.B);
    kern<<<grid, 128, p.smem_bytes, st>>>(tmap, a);
    return check_launch("conv3x3_umma");
}

#define NGAN_CONV_PCASE(CI, CO)                                                                          \
    if (cin == CI && cout == CO) {                                                                       \
        switch (epi) {                                                                                   \
            case EPI_FWD_PN: return launch_conv_persist<CI, CO, EPI_FWD_PN>(tmap, a, plan, n_ctas, st);  \
            case EPI_LINEAR: return launch_conv_persist<CI, CO, EPI_LINEAR>(tmap, a, plan, n_ctas, st);  \
            case EPI_BWD_PN: return launch_conv_persist<CI, CO, EPI_BWD_PN>(tmap, a, plan, n_ctas, st);  \
            case EPI_DBL: return launch_conv_persist<CI, CO, EPI_DBL>(tmap, a, plan, n_ctas, st);        \
        }                                                                                                \
    }

#define NGAN_CONV_CASE(CI, CO)                                                                    \
    if (cin == CI && cout == CO) {                                                                \
        switch (epi) {                                                                            \
            case EPI_FWD_PN: return launch_conv<CI, CO, EPI_FWD_PN>(tmap, a, plan, st);           \
            case EPI_LINEAR: return launch_conv<CI, CO, EPI_LINEAR>(tmap, a, plan, st);           \
            case EPI_BWD_PN: return launch_conv<CI, CO, EPI_BWD_PN>(tmap, a, plan, st);           \
            case EPI_DBL: return launch_conv<CI, CO, EPI_DBL>(tmap, a, plan, st);                 \
        }                                                                                         \
    }

int conv3x3_dispatch(int epi, const void* x, const void* wprep, int B, int cin, int cout, int H, int W, float scale,
                     float leak, const float* bias, void* out0, void* out1, float* rout, const void* y,
                     const float* r, const void* gy, const void* addin, cudaStream_t st) {
    TilePlan plan;
    int n_ctas = 0;
    static const bool no_persist = getenv("NGAN_NO_PERSISTENT") != nullptr;
    const bool persistent = !no_persist && plan_persistent(B, H, W, cin, cout, &plan, &n_ctas);
    if (!persistent && !plan_tiles(B, H, W, cin, cout, &plan)) {
        set_error("conv3x3: no tile plan for H=%d W=%d cin=%d cout=%d", H, W, cin, cout);
        return NGAN_ERR_UNSUPPORTED;
    }
    CUtensorMap tmap;
    int rc = make_c8_tensor_map(&tmap, x, B, cin, H, W, plan.Wh, plan.TH + 2, cin / 8);
    if (rc) return rc;
    ConvArgs a;
    a.B = B; a.H = H; a.W = W;
    a.TH = plan.TH; a.TW = plan.TW; a.Wh = plan.Wh;
    a.nMT = plan.nMT; a.tmem_cols = plan.tmem_cols; a.n_stage = plan.n_stage;
    a.tiles_x = (W + plan.TW - 1) / plan.TW;
    a.tiles_y = (H + plan.TH - 1) / plan.TH;
    a.n_tiles = a.tiles_x * a.tiles_y * B;
    a.plane_bytes = plan.plane_bytes;
    a.scale = scale; a.leak = leak;
    a.wprep = static_cast<const __nv_bfloat16*>(wprep);
    a.bias = bias;
    a.out0 = static_cast<__nv_bfloat16*>(out0);
    a.out1 = static_cast<__nv_bfloat16*>(out1);
    a.rout = rout;
    a.y = static_cast<const __nv_bfloat16*>(y);
    a.r = r;
    a.gy = static_cast<const __nv_bfloat16*>(gy);
    a.addin = static_cast<const __nv_bfloat16*>(addin);
    if (persistent) {
        NGAN_CONV_PCASE(16, 16)
        NGAN_CONV_PCASE(16, 32)
        NGAN_CONV_PCASE(32, 16)
        NGAN_CONV_PCASE(32, 32)
        NGAN_CONV_PCASE(32, 64)
        NGAN_CONV_PCASE(64, 32)
        NGAN_CONV_PCASE(64, 64)
    }
    NGAN_CONV_CASE(16, 16)
    NGAN_CONV_CASE(16, 32)
    NGAN_CONV_CASE(32, 16)
    NGAN_CONV_CASE(32, 32)
    NGAN_CONV_CASE(32, 64)
    NGAN_CONV_CASE(64, 32)
    NGAN_CONV_CASE(64, 64)
    NGAN_CONV_CASE(64, 128)
    NGAN_CONV_CASE(128, 64)
    NGAN_CONV_CASE(128, 128)
    set_error("conv3x3: unsupported channel pair %d -> %d (supported: 16/32/64/128, ratio 1, 2 or 1/2)", cin, cout);
    return NGAN_ERR_UNSUPPORTED;
}

// --------------------------------------------------------------------------------- weight preparation
// fp32 master weight [COUT][CIN][3][3] (torch layout, reference models.py:172-181) ->
//   fwd   image bf16 [9][CIN/8][COUT][8]   : B(n=co, k=(tap, ci)) for y = conv(x, W)
//   dgrad image bf16 [9][COUT/8][CIN][8]   : B(n=ci, k=(tap', co)) with tap' = 8 - tap (flipped), dx = convT(g, W)
__global__ void prep_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                        __nv_bfloat16* __restrict__ dgrad, int cin, int cout) {
    const int n = cin * cout * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int tap = i % 9;
        const int ci = (i / 9) % cin;
        const int co = i / (9 * cin);
        const __nv_bfloat16 v = __float2bfloat16(w[i]);
        if (fwd) fwd[((static_cast<size_t>(tap) * (cin / 8) + ci / 8) * cout + co) * 8 + (ci % 8)] = v;
        if (dgrad) dgrad[((static_cast<size_t>(8 - tap) * (cout / 8) + co / 8) * cin + ci) * 8 + (co % 8)] = v;
    }
}

int prep_conv_weight(const float* w, void* fwd, void* dgrad, int cin, int cout, cudaStream_t st) {
    const int n = cin * cout * 9;
    int blocks = (n + 255) / 256;
    if (blocks > 592) blocks = 592;
    prep_conv_weight_kernel<<<blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(fwd),
                                                    static_cast<__nv_bfloat16*>(dgrad), cin, cout);
    return check_launch("prep_conv_weight");
}

}  // namespace ngan
