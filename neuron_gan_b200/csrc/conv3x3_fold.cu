// 3x3 convolution for the narrow layers (CIN*COUT <= 4096, i.e. the 16/32/64-channel layers that carry almost
// all pixels of the 128^2..512^2 phases): persistent, warp-specialised, with the three horizontal taps folded
// into the MMA's N dimension.
//
// Why folding: with N = COUT = 16 a tcgen05.mma (M=128, K=16) needs only 8 tensor cycles but must fetch a 4 KB
// A slice from shared memory; the per-tap kernel re-reads the input tile nine times and ends up bound by the
// shared-memory operand path (measured: 10 % tensor-pipe activity, 49 % of HBM bandwidth at 512^2).  Here one
// MMA computes, for every haloed pixel q, the products with all three horizontal taps at once:
//
//     D'[m, kx*COUT + co] = sum_{ky, ci} X[q(m) + ky*Wh, ci] * W[co][ci][ky][kx]        (N = 3*COUT, 3 MMAs per K=16)
//     out[m, co]          = D'[m, co] + D'[m+1, COUT + co] + D'[m+2, 2*COUT + co]       (epilogue)
//
// The vertical taps stay on the A side as descriptor start offsets (rows of the flattened haloed tile, as in
// conv3x3_umma.cu); the horizontal taps become a shift between accumulator ROWS, i.e. TMEM lanes, which the
// epilogue resolves with two warp shuffles per value.  Tiles are 30 pixels wide in a 32-wide haloed row, so one
// warp owns exactly one tile row and lanes 30/31 -- the ones a shuffle cannot serve -- are halo columns anyway.
// The input tile is fetched from shared memory 3x instead of 9x and each MMA does 3x the work.
//
// Roles (one CTA per SM, persistent over tiles): warp 0 = TMA producer (ring of haloed input tiles; the folded
// weight image is loaded once per CTA), warps 1-2 = MMA issuers, one per TMEM accumulator buffer (even / odd
// tiles), warps 3..18 = four epilogue groups draining the buffers (TMEM -> registers -> fused tail -> HBM).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "conv_args.cuh"
#include "kernels.h"

namespace ngan {

constexpr int kFoldTW = 30, kFoldWh = 32;
// epilogue groups of 4 warps (one warp per TMEM lane quadrant); the 64-channel epilogues need > 96 registers
// per thread (as does the 32-channel double-backward one), so those kernels run two groups (352 threads)
// instead of four (608)
template <int COUT, int EPI> struct FoldCfg {
    static constexpr int kEpiGroups = (COUT >= 64 || (COUT >= 32 && EPI == EPI_DBL)) ? 2 : 4;
    static constexpr int kThreads = 96 + 128 * kEpiGroups;   // producer + 2 MMA warps + epilogue groups
};
// timing-experiment hooks (NGAN_CONV_DEBUG / NGAN_CONV_TRACE) are compiled in only with -DNGAN_CONV_DEBUG_BUILD:
// their predicates cost ~10 % of the instructions of the issue-bound epilogue
// tuning switches of the epilogue <-> MMA hand-over (defaults = best measured, see profiles/README.md)
#ifndef NGAN_EPI_EARLY
#define NGAN_EPI_EARLY -1       // give the accumulator buffer back right after the TMEM loads (1), after the stores (0),
#endif                          // or per kernel as measured (-1): early for the light linear epilogue and for the
                                // 32/64-channel layers, late for the heavy 16-channel epilogues (early cost them 15 %)
#ifndef NGAN_EPI_ALLARRIVE
#define NGAN_EPI_ALLARRIVE 1    // every epilogue thread arrives on the buffer's barrier (1) or one elected lane per warp (0)
#endif
__device__ __forceinline__ void acc_release(uint64_t* bar) {
    tc_fence_before();
#if NGAN_EPI_ALLARRIVE
    mbar_arrive(bar);
#else
    __syncwarp();
    if (elect_one()) mbar_arrive(bar);
#endif
}
#ifdef NGAN_CONV_DEBUG_BUILD
constexpr bool kDebug = true;
#else
constexpr bool kDebug = false;
#endif
constexpr int kFoldMaxStages = 12;
constexpr int kFoldMaxAcc = 8;            // accumulator buffers in TMEM (ring between MMA issuers and epilogue)
constexpr uint32_t kFoldBarBytes = (1 + 2 * kFoldMaxStages + 2 * kFoldMaxAcc) * 8 + 16 + 16 + 64 * 4;   // + ToImage weights

// tile -> (sample, tile row, tile column) without integer division: (n + 0.5) * (1/d) truncated is exact for
// n*d < ~4e6, far above any tile count here (a 512x512 batch of 128 has 73728 tiles)
__device__ __forceinline__ void tile_coords(const ConvArgs& a, int tile, int& b, int& tile_y, int& tile_x) {
    b = __float2int_rz((static_cast<float>(tile) + 0.5f) * a.inv_tiles_per_img);
    const int t2 = tile - b * (a.tiles_x * a.tiles_y);
    tile_y = __float2int_rz((static_cast<float>(t2) + 0.5f) * a.inv_tiles_x);
    tile_x = t2 - tile_y * a.tiles_x;
}

// Fused pointwise tail on one output pixel held in registers (o[c] = raw accumulator sums).
template <int COUT, int EPI>
__device__ __forceinline__ void conv_tail(const ConvArgs& a, float* o, bool valid, size_t q0, size_t p0, size_t HW,
                                          const float* s_tw) {
    constexpr int NCH = COUT / 8;
    const float inv_c = 1.0f / COUT;
    if constexpr (EPI == EPI_FWD_PN) {
        float ss = 0.f;
        float k;      // multiplier that turns the values kept in o[] into the PixelNorm output
        float rinv;   // PixelNorm scale of the true pre-norm activation h = lrelu(scale*acc + bias)
        if (a.bias == nullptr) {
            // lrelu and PixelNorm commute with the positive scale: normalise lrelu(acc) directly
            // (mean(h^2) + eps = scale^2 * (mean(v^2) + eps/scale^2)); saves one multiply per channel.
            // Packed f32x2 arithmetic (FMUL2 / FFMA2): this epilogue is instruction-issue bound.
            float2* o2 = reinterpret_cast<float2*>(o);
            const float2 leak2 = make_float2(a.leak, a.leak);
            float2 ss2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int c = 0; c < COUT / 2; ++c) {
                const float2 l = __fmul2_rn(o2[c], leak2);
                const float2 v = make_float2(fmaxf(o2[c].x, l.x), fmaxf(o2[c].y, l.y));
                ss2 = __ffma2_rn(v, v, ss2);
                o2[c] = v;
            }
            ss = ss2.x + ss2.y;
            const float inv_s = 1.0f / a.scale;
            k = rsqrtf(ss * inv_c + 1e-8f * inv_s * inv_s);
            rinv = k * inv_s;
        } else {
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                const float x = fmaf(a.scale, o[c], __ldg(a.bias + c));
                const float v = fmaxf(x, a.leak * x);
                ss = fmaf(v, v, ss);
                o[c] = v;
            }
            k = rinv = rsqrtf(ss * inv_c + 1e-8f);
        }
        if (valid) {
            if (a.img_out) {
                // fused ToImage (models.py:141-149): plain 1x1 conv to one channel + tanh on the normalised activation
                // (weights from shared memory as four-float loads: 16 scalar __ldg per pixel cost the issue-bound
                // epilogue more than the y store this variant saves)
                float dot = 0.f;
                const float4* tw4 = reinterpret_cast<const float4*>(s_tw);
#pragma unroll
                for (int c = 0; c < COUT / 4; ++c) {
                    const float4 w = tw4[c];
                    dot = fmaf(w.x, o[4 * c], dot);
                    dot = fmaf(w.y, o[4 * c + 1], dot);
                    dot = fmaf(w.z, o[4 * c + 2], dot);
                    dot = fmaf(w.w, o[4 * c + 3], dot);
                }
                const float im = tanhf(k * dot);
                if (a.img_bf16) reinterpret_cast<__nv_bfloat16*>(a.img_out)[p0] = __float2bfloat16(im);
                else a.img_out[p0] = im;
            }
            if (a.out0) {
                uint4* out = reinterpret_cast<uint4*>(a.out0) + q0;
                float2* o2 = reinterpret_cast<float2*>(o);
                const float2 k2 = make_float2(k, k);
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) o2[j * 4 + e] = __fmul2_rn(o2[j * 4 + e], k2);
                    out[j * HW] = pack8(o + j * 8);
                }
            }
            if (a.rout) a.rout[p0] = rinv;
        }
    } else if constexpr (EPI == EPI_LINEAR) {
        if (valid) {
            uint4* out = reinterpret_cast<uint4*>(a.out0);
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float2* o2 = reinterpret_cast<float2*>(o);
                const float2 s2 = make_float2(a.scale, a.scale);
#pragma unroll
                for (int e = 0; e < 4; ++e) o2[j * 4 + e] = __fmul2_rn(o2[j * 4 + e], s2);
                out[q0 + j * HW] = pack8(o + j * 8);
            }
        }
    } else if constexpr (EPI == EPI_BWD_PN) {
        // ga = mask(y) * r * (g - y*mean_c(g*y)) (+ addin), g = scale*acc     [SURVEY.md 8a row 3]
        // = mask(y) * (A*acc - Bt*y) with A = r*scale, Bt = r*t, t = scale*mean_c(acc*y): packed f32x2 arithmetic.
        // Wide layers (COUT > 16) re-load y in the second pass (an L1 hit) instead of holding it in registers.
        if (!valid) return;
        constexpr bool kKeep = COUT <= 16;
        const uint4* yq = reinterpret_cast<const uint4*>(a.y) + q0;
        __align__(8) float ykeep[kKeep ? COUT : 8];
        const float rinv = __ldg(a.r + p0);
        float2* o2 = reinterpret_cast<float2*>(o);
        float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            __align__(8) float ytmp[8];
            float* yv = kKeep ? ykeep + j * 8 : ytmp;
            unpack8(__ldg(yq + j * HW), yv);
            const float2* y2 = reinterpret_cast<const float2*>(yv);
#pragma unroll
            for (int e = 0; e < 4; ++e) dot2 = __ffma2_rn(o2[j * 4 + e], y2[e], dot2);
        }
        const float t = (dot2.x + dot2.y) * a.scale * inv_c;
        const float A = rinv * a.scale, nBt = -rinv * t;
        const float2 A2 = make_float2(A, A), nBt2 = make_float2(nBt, nBt), s2 = make_float2(a.scale, a.scale);
        uint4* out = reinterpret_cast<uint4*>(a.out0);
        uint4* out1 = reinterpret_cast<uint4*>(a.out1);
        const uint4* aq = reinterpret_cast<const uint4*>(a.addin);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            __align__(8) float ytmp[8], ad[8], ga[8];
            float* yv = kKeep ? ykeep + j * 8 : ytmp;
            if constexpr (!kKeep) unpack8(__ldg(yq + j * HW), yv);
            if (aq) unpack8(__ldg(aq + q0 + j * HW), ad);
            const float2* y2 = reinterpret_cast<const float2*>(yv);
            float2* ga2 = reinterpret_cast<float2*>(ga);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float2 v = __ffma2_rn(y2[e], nBt2, __fmul2_rn(o2[j * 4 + e], A2));
                v = __fmul2_rn(v, make_float2(lrelu_mask(y2[e].x, a.leak), lrelu_mask(y2[e].y, a.leak)));
                if (aq) v = __fadd2_rn(v, reinterpret_cast<const float2*>(ad)[e]);
                ga2[e] = v;
            }
            out[q0 + j * HW] = pack8(ga);
            if (out1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) ga2[e] = __fmul2_rn(o2[j * 4 + e], s2);
                out1[q0 + j * HW] = pack8(ga);
            }
        }
    } else {  // EPI_DBL, formulas in conv3x3_umma.cu / SURVEY.md 8a row 3 (packed f32x2 arithmetic)
        //   gh = mask*scale*acc;  t = mean(gy*y), u = mean(gh*y), w = mean(gh*gy)
        //   out0 (ghat_y) = r*(gh - u*y);  out1 (ahat) = -mask*r^2*(t*gh + u*gy + (w - 3ut)*y)
        if (!valid) return;
        constexpr bool kKeep = COUT <= 16;
        const uint4* yq = reinterpret_cast<const uint4*>(a.y) + q0;
        const uint4* gq = reinterpret_cast<const uint4*>(a.gy) + q0;
        __align__(8) float ykeep[kKeep ? COUT : 8], gkeep[kKeep ? COUT : 8];
        const float rinv = __ldg(a.r + p0);
        float2* o2 = reinterpret_cast<float2*>(o);
        const float2 s2 = make_float2(a.scale, a.scale);
        float2 t2 = make_float2(0.f, 0.f), u2 = t2, w2 = t2;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            __align__(8) float ytmp[8], gtmp[8];
            float* yv = kKeep ? ykeep + j * 8 : ytmp;
            float* gv = kKeep ? gkeep + j * 8 : gtmp;
            unpack8(__ldg(yq + j * HW), yv);
            unpack8(__ldg(gq + j * HW), gv);
            const float2* y2 = reinterpret_cast<const float2*>(yv);
            const float2* g2 = reinterpret_cast<const float2*>(gv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 m = make_float2(lrelu_mask(y2[e].x, a.leak), lrelu_mask(y2[e].y, a.leak));
                const float2 gh = __fmul2_rn(__fmul2_rn(o2[j * 4 + e], s2), m);
                o2[j * 4 + e] = gh;
                t2 = __ffma2_rn(g2[e], y2[e], t2);
                u2 = __ffma2_rn(gh, y2[e], u2);
                w2 = __ffma2_rn(gh, g2[e], w2);
            }
        }
        const float t = (t2.x + t2.y) * inv_c, u = (u2.x + u2.y) * inv_c, w = (w2.x + w2.y) * inv_c;
        const float k3 = w - 3.f * u * t;
        const float nr2 = -rinv * rinv;
        const float2 r2 = make_float2(rinv, rinv), nru2 = make_float2(-rinv * u, -rinv * u);
        const float2 ct = make_float2(nr2 * t, nr2 * t), cu = make_float2(nr2 * u, nr2 * u), ck = make_float2(nr2 * k3, nr2 * k3);
        uint4* out0 = reinterpret_cast<uint4*>(a.out0);
        uint4* out1 = reinterpret_cast<uint4*>(a.out1);
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            __align__(8) float ytmp[8], gtmp[8], o0[8], o1[8];
            float* yv = kKeep ? ykeep + j * 8 : ytmp;
            float* gv = kKeep ? gkeep + j * 8 : gtmp;
            if constexpr (!kKeep) {
                unpack8(__ldg(yq + j * HW), yv);
                unpack8(__ldg(gq + j * HW), gv);
            }
            const float2* y2 = reinterpret_cast<const float2*>(yv);
            const float2* g2 = reinterpret_cast<const float2*>(gv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 gh = o2[j * 4 + e];
                reinterpret_cast<float2*>(o0)[e] = __ffma2_rn(y2[e], nru2, __fmul2_rn(gh, r2));
                const float2 m = make_float2(lrelu_mask(y2[e].x, a.leak), lrelu_mask(y2[e].y, a.leak));
                float2 v = __ffma2_rn(y2[e], ck, __ffma2_rn(g2[e], cu, __fmul2_rn(gh, ct)));
                reinterpret_cast<float2*>(o1)[e] = __fmul2_rn(v, m);
            }
            out0[q0 + j * HW] = pack8(o0);
            out1[q0 + j * HW] = pack8(o1);
        }
    }
}

// NKX = 3: the three horizontal taps are folded into the MMA's N dimension (N = 3*COUT, 3 MMAs per K = 16 slice,
//          the epilogue reads 3*COUT accumulator columns per pixel and combines them with warp shuffles);
// NKX = 1: one MMA per tap (N = COUT, 9 MMAs per K = 16 slice, A start address shifted by kx pixels), plain epilogue.
// Both read the same folded weight image [3 ky][CIN/8][3*COUT (kx, co)][8].
template <int CIN, int COUT, int EPI, int NKX>
__global__ void __launch_bounds__(FoldCfg<COUT, EPI>::kThreads) conv3x3_fold_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                    const ConvArgs a) {
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    constexpr int NF = 3 * COUT;                       // rows of the weight image per (ky, k-group): (kx, co)
    constexpr int NMMA = NKX == 3 ? NF : COUT;         // MMA N
    constexpr uint32_t W_BYTES = 9 * CIN * COUT * 2;   // [3 ky][CIN/8][NF][8] bf16
    constexpr uint32_t IDESC = umma_idesc_bf16(128, NMMA);
    constexpr int kFoldEpiGroups = FoldCfg<COUT, EPI>::kEpiGroups;
    constexpr bool kEarlyRelease = NGAN_EPI_EARLY < 0 ? (EPI == EPI_LINEAR || COUT >= 32) : (NGAN_EPI_EARLY != 0);

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const uint32_t in_bytes = (CIN / 8) * a.plane_bytes;   // plane = (TH+2)*32*16, a multiple of 128
    uint8_t* s_in = smem;
    uint8_t* s_w = smem + a.n_stage * in_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + ((W_BYTES + 127) & ~127u));
    uint64_t* bar_w = bars;
    uint64_t* bar_full = bars + 1;            // [n_stage <= kFoldMaxStages]
    uint64_t* bar_empty = bars + 1 + kFoldMaxStages;
    uint64_t* bar_acc_full = bars + 1 + 2 * kFoldMaxStages;        // [n_acc <= kFoldMaxAcc]
    uint64_t* bar_acc_empty = bar_acc_full + kFoldMaxAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + kFoldMaxAcc);
    float* s_tw = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

    // warp index through a shuffle: tells the compiler it is warp-uniform, so everything derived from it (tile
    // counters, descriptors) stays in uniform registers and the UTCHMMA / UTMALDG issue needs no per-lane loop
    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t buf_cols = a.nMT * NMMA;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmap);
        mbar_init(bar_w, 1);
        for (int s = 0; s < a.n_stage; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 1);
        }
        for (int s = 0; s < a.n_acc; ++s) {
            mbar_init(bar_acc_full + s, 1);
            mbar_init(bar_acc_empty + s, (NGAN_EPI_ALLARRIVE ? 128 : 4) * kFoldEpiGroups);
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
    if constexpr (EPI == EPI_FWD_PN) {
        if (a.img_out) {       // fused ToImage: its COUT weights, once per CTA
            if (threadIdx.x < COUT) s_tw[threadIdx.x] = __ldg(a.toim_w + threadIdx.x);
            __syncthreads();
        }
    }

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(bar_w, W_BYTES);
            bulk_load_1d(s_w, a.wprep, W_BYTES, bar_w);
            int stage = 0;
            uint32_t sph = 0;        // phase of the ring: flips every time `stage` wraps
            bool wrapped = false;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                if (wrapped) mbar_wait(bar_empty + stage, sph ^ 1);
                int b, tile_y, tile_x;
                tile_coords(a, tile, b, tile_y, tile_x);
                if (kDebug && (a.debug & 4) && wrapped) {
                    mbar_arrive(bar_full + stage);       // timing experiment: reuse whatever is in the stage
                } else {
                    mbar_arrive_expect_tx(bar_full + stage, in_bytes);
                    tma_load_4d(s_in + stage * in_bytes, &tmap, bar_full + stage, (tile_x * kFoldTW - 1) * 2,
                                tile_y * a.TH - 1, 0, b);
                }
                if (++stage == a.n_stage) {
                    stage = 0;
                    sph ^= 1;
                    wrapped = true;
                }
            }
        }
    } else if (warp <= 2) {
        // Two MMA-issuing warps taking alternate tiles: while one is parked in the mbarrier waits of its next
        // tile the other keeps the tensor core fed.  The loop is warp-uniform; one elected lane issues.
        mbar_wait_warp(bar_w, 0, lane);
        const uint32_t w_base = smem_u32(s_w);
        // ring positions advance by two per iteration (the other MMA warp takes the tiles in between)
        int stage = warp - 1, buf = warp - 1;            // n_stage >= 2, n_acc >= 2
        uint32_t sph = 0, bph = 0;
        bool reused = false;                             // this accumulator buffer has been used before
        auto advance = [](int& idx, uint32_t& ph, int n) {
            if (++idx == n) {
                idx = 0;
                ph ^= 1;
                return true;
            }
            return false;
        };
        for (int it = warp - 1;; it += 2) {
            const int tile = blockIdx.x + it * gridDim.x;
            if (tile >= a.n_tiles) break;
            const bool trace = kDebug && a.dbg_clock && blockIdx.x == 0 && it < 32 && lane == 0;
            if (trace) a.dbg_clock[it * 8 + 0] = clock64();
            if (reused) mbar_wait_warp(bar_acc_empty + buf, bph ^ 1, lane);
            if (trace) a.dbg_clock[it * 8 + 1] = clock64();
            mbar_wait_warp(bar_full + stage, sph, lane);
            tc_fence_after();
            if (trace) a.dbg_clock[it * 8 + 2] = clock64();
            if (elect_one()) {
                const uint32_t in_base = smem_u32(s_in + stage * in_bytes);
                const uint32_t acc = tmem_base + buf * buf_cols;
                const uint32_t a_hi = umma_desc_hi(128), b_hi = umma_desc_hi(128);
                if (!(kDebug && (a.debug & 1))) {
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                        for (int kx = 0; kx < (NKX == 3 ? 1 : 3); ++kx) {
#pragma unroll
                            for (int kc = 0; kc < CIN / 16; ++kc) {
                                const uint32_t a_addr =
                                    in_base + (2 * kc) * a.plane_bytes + (ky * kFoldWh + kx) * 16;
                                const uint32_t b_addr = w_base + ((ky * (CIN / 8) + 2 * kc) * NF + kx * COUT) * 16;
                                const uint32_t b_lo = umma_desc_lo(b_addr, NF * 16);
                                uint32_t a_lo = umma_desc_lo(a_addr, a.plane_bytes);
#pragma unroll 4
                                for (int mt = 0; mt < a.nMT; ++mt) {
                                    umma_bf16_2x32(acc + mt * NMMA, a_lo, a_hi, b_lo, b_hi, IDESC, (ky | kx | kc) != 0);
                                    a_lo += (128 * 16) >> 4;      // next M-tile: 128 rows further (start-address field)
                                }
                            }
                        }
                    }
                }
                umma_commit(bar_empty + stage);
                umma_commit(bar_acc_full + buf);
            }
            __syncwarp();
            if (trace) a.dbg_clock[it * 8 + 3] = clock64();
            advance(stage, sph, a.n_stage);
            advance(stage, sph, a.n_stage);
            reused |= advance(buf, bph, a.n_acc);
            reused |= advance(buf, bph, a.n_acc);
        }
    } else {
        const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
        const int group = (warp - 3) >> 2;         // epilogue group: M-tiles are interleaved between the groups
        const size_t HW = static_cast<size_t>(a.H) * a.W;
        int it = 0, buf = 0;
        uint32_t bph = 0;
        for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
            int b, tile_y, tile_x;
            tile_coords(a, tile, b, tile_y, tile_x);
            const bool trace = kDebug && a.dbg_clock && blockIdx.x == 0 && it < 32 && warp == 3 && lane == 0;
            if (trace) a.dbg_clock[it * 8 + 5] = clock64();
            mbar_wait_warp(bar_acc_full + buf, bph, lane);
            tc_fence_after();
            if (trace) a.dbg_clock[it * 8 + 6] = clock64();
            const int ox = tile_x * kFoldTW + lane;
            // M-tiles of consecutive tiles rotate over the groups (with one M-tile per tile, tiles alternate)
            const int mt0 = static_cast<int>((static_cast<unsigned>(group) - static_cast<unsigned>(it * a.nMT)) &
                                             (kFoldEpiGroups - 1));
            // The accumulator buffer goes back to the MMA warps as soon as this thread's TMEM loads have landed in
            // registers -- before the arithmetic and the stores -- so the next tile's MMAs overlap this epilogue.
            bool released = false;
            for (int mt = mt0; mt < ((kDebug && (a.debug & 2)) ? 0 : a.nMT); mt += kFoldEpiGroups) {
                const bool last_mt = mt + kFoldEpiGroups >= a.nMT;
                const int rr = mt * 4 + quad;      // row of the tile (Wh = 32: one warp = one row)
                const int oy = tile_y * a.TH + rr;
                const bool valid = (lane < kFoldTW) && (rr < a.TH) && (oy < a.H) && (ox < a.W);
                const uint32_t taddr = tmem_base + buf * buf_cols + (static_cast<uint32_t>(quad * 32) << 16) + mt * NMMA;
                __align__(8) float o[COUT];
                if constexpr (NKX == 3) {
                    __align__(8) float v1[16], v2[16];
#pragma unroll
                    for (int c0 = 0; c0 < COUT; c0 += 16) {
                        tmem_ld16_nowait(taddr + c0, o + c0);
                        tmem_ld16_nowait(taddr + COUT + c0, v1);
                        tmem_ld16_nowait(taddr + 2 * COUT + c0, v2);
                        tmem_ld_wait();
                        if (kEarlyRelease && last_mt && c0 + 16 >= COUT) {
                            acc_release(bar_acc_empty + buf);
                            released = true;
                        }
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const float2 s1 = make_float2(__shfl_down_sync(0xffffffffu, v1[i], 1),
                                                          __shfl_down_sync(0xffffffffu, v1[i + 1], 1));
                            const float2 s2 = make_float2(__shfl_down_sync(0xffffffffu, v2[i], 2),
                                                          __shfl_down_sync(0xffffffffu, v2[i + 1], 2));
                            float2* op = reinterpret_cast<float2*>(o + c0 + i);
                            *op = __fadd2_rn(*op, __fadd2_rn(s1, s2));
                        }
                    }
                } else {
#pragma unroll
                    for (int c0 = 0; c0 < COUT; c0 += 16) tmem_ld16_nowait(taddr + c0, o + c0);
                    tmem_ld_wait();
                    if (kEarlyRelease && last_mt) {
                        acc_release(bar_acc_empty + buf);
                        released = true;
                    }
                }
                const size_t q0 = static_cast<size_t>(b) * (COUT / 8) * HW + static_cast<size_t>(oy) * a.W + ox;
                const size_t p0 = static_cast<size_t>(b) * HW + static_cast<size_t>(oy) * a.W + ox;
                conv_tail<COUT, EPI>(a, o, valid, q0, p0, HW, s_tw);
            }
            if (!released) acc_release(bar_acc_empty + buf);     // (or no M-tile of this tile fell to this group)
            if (trace) a.dbg_clock[it * 8 + 7] = clock64();
            if (kDebug && a.dbg_clock && blockIdx.x == 0 && it < 32 && warp == 2 + 4 * kFoldEpiGroups && lane == 0)
                a.dbg_clock[it * 8 + 4] = clock64();      // the last epilogue warp of the CTA
            if (++buf == a.n_acc) {
                buf = 0;
                bph ^= 1;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------------------------------------ host side
long long* g_conv_trace = nullptr;
static int pow2_cols(int n) {
    int c = 32;
    while (c < n) c <<= 1;
    return c;
}

template <int CIN, int COUT, int EPI, int NKX>
static int launch_fold(const CUtensorMap& tmap, const ConvArgs& a, uint32_t smem_bytes, int n_ctas, cudaStream_t st) {
    auto kern = conv3x3_fold_kernel<CIN, COUT, EPI, NKX>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv3x3_fold)");
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_ctas);
    cfg.blockDim = dim3(FoldCfg<COUT, EPI>::kThreads);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tmap, a);
    if (le != cudaSuccess) return check_cuda(le, "cudaLaunchKernelEx(conv3x3_fold)");
    return check_launch("conv3x3_fold");
}

#define NGAN_FOLD_CASE_K(CI, CO, K)                                                                       \
    switch (epi) {                                                                                      \
        case EPI_FWD_PN: return launch_fold<CI, CO, EPI_FWD_PN, K>(tmap, a, smem_bytes, n_ctas, st);      \
        case EPI_LINEAR: return launch_fold<CI, CO, EPI_LINEAR, K>(tmap, a, smem_bytes, n_ctas, st);      \
        case EPI_BWD_PN: return launch_fold<CI, CO, EPI_BWD_PN, K>(tmap, a, smem_bytes, n_ctas, st);      \
        case EPI_DBL: return launch_fold<CI, CO, EPI_DBL, K>(tmap, a, smem_bytes, n_ctas, st);            \
    }
// The one-MMA-per-tap variant (NKX = 1) was measured and rejected (profiles/README.md); it is only built with
// -DNGAN_BUILD_NKX1 (then NGAN_FOLD_NKX=1 selects it at run time).
#ifdef NGAN_BUILD_NKX1
#define NGAN_FOLD_CASE(CI, CO)              \
    if (cin == CI && cout == CO) {          \
        if (nkx == 3) {                     \
            NGAN_FOLD_CASE_K(CI, CO, 3)     \
        } else {                            \
            NGAN_FOLD_CASE_K(CI, CO, 1)     \
        }                                   \
    }
#else
#define NGAN_FOLD_CASE(CI, CO)              \
    if (cin == CI && cout == CO) {          \
        NGAN_FOLD_CASE_K(CI, CO, 3)         \
    }
#endif

int conv3x3_fold_dispatch(int epi, const void* x, ConvArgs a, int B, int cin, int cout, int H, int W,
                          cudaStream_t st) {
    // NKX: 3 = horizontal taps folded into N, 1 = one MMA per tap (see the kernel comment)
#ifdef NGAN_BUILD_NKX1
    static const int nkx_env = getenv("NGAN_FOLD_NKX") ? atoi(getenv("NGAN_FOLD_NKX")) : 3;
    const int nkx = nkx_env == 1 ? 1 : 3;
#else
    const int nkx = 3;
#endif
    const int NF = nkx == 3 ? 3 * cout : cout;         // accumulator columns per M-tile
    // M-tiles (4 tile rows each) per CTA tile: at least two accumulator buffers of nMT*NF columns in 512 columns
    static const int nmt_env = getenv("NGAN_FOLD_NMT") ? atoi(getenv("NGAN_FOLD_NMT")) : 0;   // tuning experiments
    int nMT = 256 / NF;
    const int nmt_cap = nmt_env > 0 ? nmt_env : 4;
    if (nMT > nmt_cap) nMT = nmt_cap;
    if (nMT < 1) nMT = 1;
    const int tiles_x = (W + kFoldTW - 1) / kFoldTW;
    // shrink tiles while the grid cannot give every SM a couple of tiles
    while (nMT > 1 && static_cast<long long>(B) * ((H + 4 * nMT - 1) / (4 * nMT)) * tiles_x < 2 * 148) --nMT;
    int TH = 4 * nMT;
    if (TH > H) TH = (H + 3) / 4 * 4;
    nMT = TH / 4;
    static const int dbg = getenv("NGAN_CONV_DEBUG") ? atoi(getenv("NGAN_CONV_DEBUG")) : 0;
    a.debug = dbg;
    static long long* dbg_buf = nullptr;
    if (getenv("NGAN_CONV_TRACE") && !dbg_buf) cudaMalloc(&dbg_buf, 32 * 8 * sizeof(long long));
    a.dbg_clock = dbg_buf;
    g_conv_trace = dbg_buf;
    a.TH = TH; a.TW = kFoldTW; a.Wh = kFoldWh; a.nMT = nMT;
    static const int acc_env = getenv("NGAN_FOLD_ACC") ? atoi(getenv("NGAN_FOLD_ACC")) : kFoldMaxAcc;
    int n_acc = 512 / (nMT * NF);
    if (n_acc > kFoldMaxAcc) n_acc = kFoldMaxAcc;
    if (n_acc > acc_env) n_acc = acc_env;
    if (n_acc < 2) n_acc = 2;
    a.n_acc = n_acc;
    a.tmem_cols = pow2_cols(n_acc * nMT * NF);
    if (a.tmem_cols > 512) {
        set_error("conv3x3_fold: TMEM budget exceeded (cout=%d)", cout);
        return NGAN_ERR_UNSUPPORTED;
    }
    a.plane_bytes = static_cast<uint32_t>(TH + 2) * kFoldWh * 16;
    a.tiles_x = tiles_x;
    a.tiles_y = (H + TH - 1) / TH;
    a.n_tiles = a.tiles_x * a.tiles_y * B;
    a.inv_tiles_x = 1.0f / a.tiles_x;
    a.inv_tiles_per_img = 1.0f / (a.tiles_x * a.tiles_y);
    const uint32_t in_bytes = (cin / 8) * a.plane_bytes;
    const uint32_t w_bytes = ((9u * cin * cout * 2) + 127) & ~127u;
    // Deep ring: HBM needs ~100 KB of loads in flight per SM to run at full rate (latency x bandwidth), so the
    // producer keeps as many haloed tiles outstanding as shared memory allows.
    static const int max_stages = getenv("NGAN_FOLD_STAGES") ? atoi(getenv("NGAN_FOLD_STAGES")) : kFoldMaxStages;
    int stages = max_stages < kFoldMaxStages ? max_stages : kFoldMaxStages;
    while (stages > 2 && 128 + stages * in_bytes + w_bytes + kFoldBarBytes > 200u * 1024) --stages;
    a.n_stage = stages;
    const uint32_t smem_bytes = 128 + stages * in_bytes + w_bytes + kFoldBarBytes;
    if (smem_bytes > 227u * 1024) {
        set_error("conv3x3_fold: shared memory budget exceeded (cin=%d cout=%d)", cin, cout);
        return NGAN_ERR_UNSUPPORTED;
    }
    const int n_ctas = a.n_tiles < 148 ? a.n_tiles : 148;

    CUtensorMap tmap;
    int rc = make_c8_tensor_map(&tmap, x, B, cin, H, W, kFoldWh, TH + 2, cin / 8);
    if (rc) return rc;
    NGAN_FOLD_CASE(16, 16)
    NGAN_FOLD_CASE(16, 32)
    NGAN_FOLD_CASE(32, 16)
    NGAN_FOLD_CASE(32, 32)
    NGAN_FOLD_CASE(32, 64)
    NGAN_FOLD_CASE(64, 32)
    NGAN_FOLD_CASE(64, 64)
    set_error("conv3x3_fold: unsupported channel pair %d -> %d", cin, cout);
    return NGAN_ERR_UNSUPPORTED;
}

}  // namespace ngan
