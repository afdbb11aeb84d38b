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
#include "conv_args.cuh"
#include "kernels.h"

namespace ngan {

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
    pdl_trigger();
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

    // warp index through a shuffle: the compiler then knows it is warp-uniform, keeps the descriptor arithmetic in
    // uniform registers and issues UTCHMMA / UTMALDG back to back instead of through a per-lane loop
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
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
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();        // everything above overlapped the previous kernel's tail; its results are needed from here on

    if (warp == 0) {
        // ---- input tile: one TMA box {2*Wh x (TH+2) x CIN/8 x 1} of 8-byte elements, then the MMA issue loop
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_in, in_bytes);
            tma_load_4d(s_in, &tmap, bar_in, (tile_x * a.TW - 1) * 2, tile_y * a.TH - 1, 0, b);
        }
        __syncwarp();
        mbar_wait_warp(bar_in, 0, lane);
        tc_fence_after();
        const uint32_t in_base = smem_u32(s_in);
        const uint32_t w_base = smem_u32(s_w);
        int stage = 0;
        uint32_t ph = 0;
        for (int tap = 0; tap < 9; ++tap) {
            mbar_wait_warp(bar_wfull + stage, ph, lane);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t tap_off = ((tap / 3) * a.Wh + (tap % 3)) * 16;
                const uint32_t hi = umma_desc_hi(128);
#pragma unroll
                for (int kc = 0; kc < CIN / 16; ++kc) {
                    const uint32_t a_addr = in_base + (2 * kc) * a.plane_bytes + tap_off;
                    const uint32_t b_lo = umma_desc_lo(w_base + stage * W_TAP_BYTES + (2 * kc) * COUT * 16, COUT * 16);
                    uint32_t a_lo = umma_desc_lo(a_addr, a.plane_bytes);
                    for (int mt = 0; mt < a.nMT; ++mt) {
                        umma_bf16_2x32(tmem_base + mt * COUT, a_lo, hi, b_lo, hi, IDESC, (tap | kc) != 0);
                        a_lo += (128 * 16) >> 4;
                    }
                }
                umma_commit(bar_wempty + stage);  // slab free once these MMAs retire
                if (tap == 8) umma_commit(bar_mma);
            }
            __syncwarp();
            if (++stage == a.n_stage) {
                stage = 0;
                ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ---- weight slabs, one per tap, through a ring of n_stage buffers
        if (elect_one()) {
            int stage = 0;
            uint32_t ph = 0;
            bool wrapped = false;
            for (int tap = 0; tap < 9; ++tap) {
                if (wrapped) mbar_wait(bar_wempty + stage, ph ^ 1);
                mbar_arrive_expect_tx(bar_wfull + stage, W_TAP_BYTES);
                bulk_load_1d(s_w + stage * W_TAP_BYTES, reinterpret_cast<const uint8_t*>(a.wprep) + tap * W_TAP_BYTES,
                             W_TAP_BYTES, bar_wfull + stage);
                if (++stage == a.n_stage) {
                    stage = 0;
                    ph ^= 1;
                    wrapped = true;
                }
            }
        }
        __syncwarp();
    }

    // ---- epilogue: every thread owns one accumulator row (= one output pixel, all COUT channels)
    mbar_wait_warp(bar_mma, 0, lane);
    tc_fence_after();

    conv_epilogue<COUT, EPI>(a, tmem_base, warp, lane, b, tile_x, tile_y);

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
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tmap, a);
    if (le != cudaSuccess) return check_cuda(le, "cudaLaunchKernelEx(conv3x3_umma)");
    return check_launch("conv3x3_umma");
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
                     const float* r, const void* gy, const void* addin, cudaStream_t st, const float* toim_w,
                     float* img_out, int img_bf16) {
    ConvArgs a;
    a.B = B; a.H = H; a.W = W;
    a.scale = scale; a.leak = leak;
    a.debug = 0;
    a.dbg_clock = nullptr;
    a.wprep = static_cast<const __nv_bfloat16*>(wprep);
    a.bias = bias;
    a.out0 = static_cast<__nv_bfloat16*>(out0);
    a.out1 = static_cast<__nv_bfloat16*>(out1);
    a.rout = rout;
    a.toim_w = toim_w;
    a.img_out = img_out;
    a.img_bf16 = img_bf16;
    a.y = static_cast<const __nv_bfloat16*>(y);
    a.r = r;
    a.gy = static_cast<const __nv_bfloat16*>(gy);
    a.addin = static_cast<const __nv_bfloat16*>(addin);
    if (conv_uses_folded_kernel(cin, cout)) return conv3x3_fold_dispatch(epi, x, a, B, cin, cout, H, W, st);
    if (img_out || !out0) {
        set_error("conv3x3: the fused ToImage epilogue exists only for folded layers (cin*cout <= 4096), got %d -> %d",
                  cin, cout);
        return NGAN_ERR_UNSUPPORTED;
    }

    TilePlan plan;
    if (!plan_tiles(B, H, W, cin, cout, &plan)) {
        set_error("conv3x3: no tile plan for H=%d W=%d cin=%d cout=%d", H, W, cin, cout);
        return NGAN_ERR_UNSUPPORTED;
    }
    CUtensorMap tmap;
    int rc = make_c8_tensor_map(&tmap, x, B, cin, H, W, plan.Wh, plan.TH + 2, cin / 8);
    if (rc) return rc;
    a.TH = plan.TH; a.TW = plan.TW; a.Wh = plan.Wh;
    a.nMT = plan.nMT; a.tmem_cols = plan.tmem_cols; a.n_stage = plan.n_stage;
    a.tiles_x = (W + plan.TW - 1) / plan.TW;
    a.tiles_y = (H + plan.TH - 1) / plan.TH;
    a.n_tiles = a.tiles_x * a.tiles_y * B;
    a.plane_bytes = plan.plane_bytes;
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
// or, when conv_uses_folded_kernel(cin, cout), the kx-folded images [3][K/8][3*N][8] of conv3x3_fold.cu.
__global__ void prep_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ fwd,
                                        __nv_bfloat16* __restrict__ dgrad, int cin, int cout, int folded) {
    const int n = cin * cout * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        conv_image_store(__float2bfloat16(w[i]), i, cin, cout, folded, fwd, dgrad);
}

int prep_conv_weight(const float* w, void* fwd, void* dgrad, int cin, int cout, cudaStream_t st) {
    const int n = cin * cout * 9;
    int blocks = (n + 255) / 256;
    if (blocks > 592) blocks = 592;
    prep_conv_weight_kernel<<<blocks, 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(fwd),
                                                    static_cast<__nv_bfloat16*>(dgrad), cin, cout,
                                                    conv_uses_folded_kernel(cin, cout) ? 1 : 0);
    return check_launch("prep_conv_weight");
}

}  // namespace ngan
