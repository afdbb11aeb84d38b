// HBM-bound pointwise / per-pixel channel-reduction kernels of the PGGAN step (everything that is not a
// 3x3 convolution): layout conversion, bilinear x2 / 2x2-mean resampling and their adjoints, PixelNorm +
// LeakyReLU backward, FromImage / ToImage (1x1 convs, reference models.py:133-165) forward / backward /
// double backward, fade-in blends (models.py:348-350, 519-521), the 16x16 critic head (models.py:485-490),
// and the WGAN-GP loss reductions (loss_functions.py:14-47, 59-74, 157-180).
//
// All feature maps are C8-planar bf16 (common.cuh); one thread owns one pixel and walks the channel groups,
// so a warp reads/writes 32 consecutive 16-byte granules (512 B, fully coalesced) per channel group and the
// per-pixel channel reductions (PixelNorm statistics, 1x1 convs) are thread-local.
#include "common.cuh"
#include "kernels.h"

namespace ngan {

static inline int nblocks(size_t n, int threads) { return static_cast<int>((n + threads - 1) / threads); }

// torch upsample_bilinear2d, scale 2, align_corners=False: source index / weight of output o (models.py:78-89)
__device__ __forceinline__ void up2_src(int o, int n_in, int& i0, int& i1, float& l1) {
    float s = (o + 0.5f) * 0.5f - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = static_cast<int>(s);
    l1 = s - i0;
    i1 = i0 + 1 < n_in ? i0 + 1 : n_in - 1;
}
// weight with which input i contributes to output o (adjoint of the above)
__device__ __forceinline__ float up2_adj_w(int o, int i, int n_in) {
    if (o < 0 || o >= 2 * n_in) return 0.f;
    int i0, i1;
    float l1;
    up2_src(o, n_in, i0, i1, l1);
    return (i0 == i ? 1.f - l1 : 0.f) + (i1 == i ? l1 : 0.f);
}

// Sum 8 per-thread values over the warp into this warp's own row of shared memory (plain stores: the block adds
// its warps' rows in warp order afterwards, so the result does not depend on scheduling).
__device__ __forceinline__ void warp_store8(const float* v, float* row8, int lane) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float s = warp_sum(v[e]);
        if (lane == 0) row8[e] = s;
    }
}

// ------------------------------------------------------------------------------- deterministic reductions
// Parameter-gradient kernels never use atomics: every block writes its partial sums as one row of a workspace
// ([n_partials][ld] fp32) and this kernel adds the rows in index order:
//     out[i] (+)= scale * sum_p partials[p * ld + i],   i < n
// A block owns OPB outputs and splits the rows over 256 / OPB slices (slice s takes rows s, s + S, ...), the slices
// are combined in slice order: the order of additions depends only on (n_partials, OPB).
template <int OPB>
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int n_partials,
                                                              long long n, long long ld, float scale,
                                                              float* __restrict__ out, int accumulate,
                                                              float* __restrict__ out2, long long n_split) {
    pdl_trigger();
    pdl_wait();           // launched right behind the kernel that writes the partials
    constexpr int S = 256 / OPB;
    __shared__ float red[S][OPB];
    const int o = threadIdx.x % OPB, s = threadIdx.x / OPB;
    const long long i = static_cast<long long>(blockIdx.x) * OPB + o;
    float v = 0.f;
    if (i < n) {
        const float* src = partials + i;
        int p = s;
        // eight loads in flight, added in row order (the order of additions is part of the contract)
        for (; p + 7 * S < n_partials; p += 8 * S) {
            float t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = __ldg(src + static_cast<long long>(p + k * S) * ld);
#pragma unroll
            for (int k = 0; k < 8; ++k) v += t[k];
        }
        for (; p < n_partials; p += S) v += __ldg(src + static_cast<long long>(p) * ld);
    }
    red[s][o] = v;
    __syncthreads();
    if (s == 0 && i < n) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < S; ++k) t += red[k][o];
        t *= scale;
        float* dst = (out2 && i >= n_split) ? out2 + (i - n_split) : out + i;      // a row may feed two tensors
        *dst = accumulate ? *dst + t : t;
    }
}
int reduce_partials(const float* partials, int n_partials, long long n, long long ld, float scale, float* out,
                    int accumulate, cudaStream_t st, float* out2, long long n_split) {
    // outputs per block: wide results with few partial rows take 64 (coalesced 256-byte rows, 4 slices); many partial
    // rows are split over 16 or 32 slices so that no thread walks more than a few dozen rows
    cudaError_t e;
    if (n >= 2048 && n_partials <= 32)
        e = launch_pdl(reduce_partials_kernel<64>, dim3(static_cast<unsigned>((n + 63) / 64)), dim3(256), 0, st,
                       partials, n_partials, n, ld, scale, out, accumulate, out2, n_split);
    else if (n >= 2048)
        e = launch_pdl(reduce_partials_kernel<16>, dim3(static_cast<unsigned>((n + 15) / 16)), dim3(256), 0, st,
                       partials, n_partials, n, ld, scale, out, accumulate, out2, n_split);
    else
        e = launch_pdl(reduce_partials_kernel<8>, dim3(static_cast<unsigned>((n + 7) / 8)), dim3(256), 0, st, partials,
                       n_partials, n, ld, scale, out, accumulate, out2, n_split);
    if (e != cudaSuccess) return check_cuda(e, "cudaLaunchKernelEx(reduce_partials)");
    return check_launch("reduce_partials");
}
// out[i] = scale * sum_k slots[k * ld + i] in slot order (the critic's three gradient contributions; the ranks'
// all-gathered small gradients of the generator: train_step.py)
__global__ void sum_slots_kernel(const float* __restrict__ slots, int n_slots, long long n, long long ld, float scale,
                                 float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float v = slots[i];
    for (int k = 1; k < n_slots; ++k) v += slots[k * ld + i];
    out[i] = scale == 1.f ? v : scale * v;
}
int sum_slots(const float* slots, int n_slots, long long n, long long ld, float scale, float* out, cudaStream_t st) {
    sum_slots_kernel<<<nblocks(static_cast<size_t>(n), 256), 256, 0, st>>>(slots, n_slots, n, ld, scale, out);
    return check_launch("sum_slots");
}
size_t pixel_reduction_workspace_bytes(int B, int C, int H, int W) {
    size_t rows = (static_cast<size_t>(B) * H * W + 127) / 128;       // fromim_dbl: one row per 128 pixels
    if (rows < 148 * 16) rows = 148 * 16;                              // fromim_bwd / toim_bwd: at most 148*16 blocks
    return rows * 2 * C * sizeof(float);
}

// ------------------------------------------------------------------------------- layout conversion
__global__ void nchw_to_c8_kernel(const float* __restrict__ src, uint4* __restrict__ dst, int C, size_t HW,
                                  size_t total) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t pix = i % HW;
    const size_t bj = i / HW;  // b*(C/8) + j
    const int nch = C / 8;
    const size_t b = bj / nch, j = bj % nch;
    const float* s = src + (b * C + j * 8) * HW + pix;
    float f[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = s[e * HW];
    dst[i] = pack8(f);
}
__global__ void c8_to_nchw_kernel(const uint4* __restrict__ src, float* __restrict__ dst, int C, size_t HW,
                                  size_t total) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const size_t pix = i % HW;
    const size_t bj = i / HW;
    const int nch = C / 8;
    const size_t b = bj / nch, j = bj % nch;
    float f[8];
    unpack8(src[i], f);
    float* d = dst + (b * C + j * 8) * HW + pix;
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e * HW] = f[e];
}
int nchw_to_c8(const float* src, void* dst, int B, int C, int H, int W, cudaStream_t st) {
    const size_t HW = static_cast<size_t>(H) * W, total = static_cast<size_t>(B) * (C / 8) * HW;
    nchw_to_c8_kernel<<<nblocks(total, 256), 256, 0, st>>>(src, static_cast<uint4*>(dst), C, HW, total);
    return check_launch("nchw_to_c8");
}
int c8_to_nchw(const void* src, float* dst, int B, int C, int H, int W, cudaStream_t st) {
    const size_t HW = static_cast<size_t>(H) * W, total = static_cast<size_t>(B) * (C / 8) * HW;
    c8_to_nchw_kernel<<<nblocks(total, 256), 256, 0, st>>>(static_cast<const uint4*>(src), dst, C, HW, total);
    return check_launch("c8_to_nchw");
}

// 32-bit index decomposition: every tensor handled here has < 2^32 granules (largest: a 512x512 batch), and
// 64-bit div/mod costs ~100 instructions each on the GPU -- more than the rest of these kernels.
__device__ __forceinline__ void split_xyb(size_t i, int W, int H, int& x, int& y, size_t& b) {
    const unsigned iu = static_cast<unsigned>(i);
    const unsigned row = iu / static_cast<unsigned>(W);
    const unsigned bb = row / static_cast<unsigned>(H);
    x = static_cast<int>(iu - row * W);
    y = static_cast<int>(row - bb * H);
    b = bb;
}
__device__ __forceinline__ void split_bpix(size_t i, size_t HW, size_t& b, size_t& pix) {
    const unsigned iu = static_cast<unsigned>(i), hw = static_cast<unsigned>(HW);
    const unsigned bb = iu / hw;
    b = bb;
    pix = iu - bb * hw;
}

// ------------------------------------------------------------------------------- resampling (forward)
// One thread per INPUT column granule x over a strip of RP input rows (align_corners = False, scale 2:
// out[2i] = .25 x[i-1] + .75 x[i], out[2i+1] = .75 x[i] + .25 x[i+1], indices clamped, separable; the fixed tap
// pattern of models.py:78-89).  The horizontally interpolated (left, right) output columns of three consecutive input
// rows roll through registers, so a strip costs 3 (RP + 2) loads (the x -/+ 1 neighbours are L1 hits) instead of 9 RP,
// and each thread writes its two adjacent output granules as ONE 32-byte store (whole sectors per instruction).
// Measured: the row reuse is what pays (96 -> 76 us at 256 -> 512, B = 48; 6.6 TB/s), the wide store alone does not.
__device__ __forceinline__ void st_global_256(uint4* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z),
                 "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ void up2_hrow(const uint4* __restrict__ row, int xm, int ix, int xp, float* hl, float* hr) {
    float a[8], b[8], c[8];
    unpack8(__ldg(row + xm), a);
    unpack8(__ldg(row + ix), b);
    unpack8(__ldg(row + xp), c);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        hl[e] = __fmaf_rn(0.75f, b[e], 0.25f * a[e]);
        hr[e] = __fmaf_rn(0.75f, b[e], 0.25f * c[e]);
    }
}
template <int RP>
__global__ void __launch_bounds__(128) upsample2x_c8_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int H,
                                                            int W) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    if (ix >= W) return;
    const int y0 = blockIdx.y * RP;
    const size_t plane = blockIdx.z;
    const uint4* p = x + plane * H * W;
    const int xm = max(ix - 1, 0), xp = min(ix + 1, W - 1);
    float hl[3][8], hr[3][8];     // rows y-1, y, y+1 of the current input row, rolling (index = row mod 3)
    up2_hrow(p + static_cast<size_t>(max(y0 - 1, 0)) * W, xm, ix, xp, hl[0], hr[0]);
    up2_hrow(p + static_cast<size_t>(y0) * W, xm, ix, xp, hl[1], hr[1]);
    uint4* q = out + plane * 4 * H * W + static_cast<size_t>(2 * y0) * (2 * W) + 2 * ix;
#pragma unroll
    for (int r = 0; r < RP; ++r) {
        up2_hrow(p + static_cast<size_t>(min(y0 + r + 1, H - 1)) * W, xm, ix, xp, hl[(r + 2) % 3], hr[(r + 2) % 3]);
        const float *l0 = hl[r % 3], *l1 = hl[(r + 1) % 3], *l2 = hl[(r + 2) % 3];
        const float *r0 = hr[r % 3], *r1 = hr[(r + 1) % 3], *r2 = hr[(r + 2) % 3];
        float ol[8], orr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ol[e] = __fmaf_rn(0.75f, l1[e], 0.25f * l0[e]);
            orr[e] = __fmaf_rn(0.75f, r1[e], 0.25f * r0[e]);
        }
        st_global_256(q, pack8(ol), pack8(orr));
        q += 2 * W;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ol[e] = __fmaf_rn(0.75f, l1[e], 0.25f * l2[e]);
            orr[e] = __fmaf_rn(0.75f, r1[e], 0.25f * r2[e]);
        }
        st_global_256(q, pack8(ol), pack8(orr));
        q += 2 * W;
    }
}
int upsample2x_c8(const void* x, void* out, int B, int C, int H, int W, cudaStream_t st) {
    const int threads = W < 128 ? ((W + 31) / 32 * 32) : 128;
    // rows per thread, measured on B200 at the generator's shapes (scripts/bench_upsample.py, B = 48): 2 rows 75.7 /
    // 39.3 / 9.9 us at 512 / 256 / 128 outputs (4 rows 79.2 / 40.7 / 10.3), 4 rows 5.8 / 5.1 us at 64 / 32 (2 rows
    // 7.3 / 7.2); one row per thread (the round-1 arrangement) 96 / 49 / 14 / 13.6 / 13.5 us
    static const int rp_env = getenv("NGAN_UP2_ROWS") ? atoi(getenv("NGAN_UP2_ROWS")) : 0;
    const int want = rp_env > 0 ? rp_env : (H <= 32 ? 4 : 2);
    const int rp = (want >= 4 && H % 4 == 0) ? 4 : (want >= 2 && H % 2 == 0) ? 2 : 1;
    const dim3 grid((W + threads - 1) / threads, H / rp, B * (C / 8));
    if (grid.z > 65535) {
        set_error("upsample2x: too many planes (%u)", grid.z);
        return NGAN_ERR_UNSUPPORTED;
    }
    const uint4* xi = static_cast<const uint4*>(x);
    uint4* o = static_cast<uint4*>(out);
    if (rp == 4) upsample2x_c8_kernel<4><<<grid, threads, 0, st>>>(xi, o, H, W);
    else if (rp == 2) upsample2x_c8_kernel<2><<<grid, threads, 0, st>>>(xi, o, H, W);
    else upsample2x_c8_kernel<1><<<grid, threads, 0, st>>>(xi, o, H, W);
    return check_launch("upsample2x_c8");
}

__global__ void avgpool2_c8_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int H, int W, size_t total) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int OW = W / 2, OH = H / 2;
    int ox, oy;
    size_t plane;
    split_xyb(i, OW, OH, ox, oy, plane);
    const uint4* p = x + plane * H * W + static_cast<size_t>(2 * oy) * W + 2 * ox;
    float a[8], b[8], c[8], d[8], o[8];
    unpack8(__ldg(p), a);
    unpack8(__ldg(p + 1), b);
    unpack8(__ldg(p + W), c);
    unpack8(__ldg(p + W + 1), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.25f * (a[e] + b[e] + c[e] + d[e]);
    out[i] = pack8(o);
}
int avgpool2_c8(const void* x, void* out, int B, int C, int H, int W, cudaStream_t st) {
    const size_t total = static_cast<size_t>(B) * (C / 8) * (H / 2) * (W / 2);
    avgpool2_c8_kernel<<<nblocks(total, 256), 256, 0, st>>>(static_cast<const uint4*>(x), static_cast<uint4*>(out), H,
                                                             W, total);
    return check_launch("avgpool2_c8");
}

// ------------------------------------------------------------------------------- PixelNorm + LeakyReLU backward
// ga = mask(y) * r * (g - y * mean_c(g*y)) + addin, g = gscale * G[(y,x) or (y/2,x/2)]   (SURVEY.md 8a row 3).
// `unpool` reads G at half resolution: the adjoint of AvgPool2d(2) with the 1/4 folded into gscale by the caller.
// NCH (= C/8) is a template parameter so that both passes unroll and all loads of a pass are in flight together
// (one thread per pixel; with a run-time channel loop the kernel ran at 40 % of HBM bandwidth on load latency);
// up to 32 channels the granules of the first pass stay in registers for the second.
template <int NCH>
__global__ void __launch_bounds__(128) pn_bwd_c8_kernel(const uint4* __restrict__ g, int unpool, float gscale_h,
                                                        const float* __restrict__ dyn,
                                                        const uint4* __restrict__ y, const float* __restrict__ r,
                                                        const uint4* __restrict__ addin, uint4* __restrict__ ga,
                                                        uint4* __restrict__ gy_out, float leak, int H, int W,
                                                        size_t total) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float gscale = dyn ? gscale_h * __ldg(dyn) : gscale_h;
    constexpr int C = NCH * 8;
    constexpr bool kKeep = NCH <= 4;
    constexpr int U = NCH < 4 ? NCH : 4;          // granules per unrolled chunk
    const size_t HW = static_cast<size_t>(H) * W;
    int px, py;
    size_t b;
    split_xyb(i, W, H, px, py, b);
    const size_t q0 = b * NCH * HW + static_cast<size_t>(py) * W + px;
    const size_t gHW = unpool ? HW / 4 : HW;
    const size_t g0 = unpool ? b * NCH * gHW + static_cast<size_t>(py >> 1) * (W >> 1) + (px >> 1) : q0;
    const float rinv = __ldg(r + i);
    uint4 gk[kKeep ? NCH : 1], yk[kKeep ? NCH : 1];
    float t = 0.f;
#pragma unroll 1
    for (int j0 = 0; j0 < NCH; j0 += U) {
        uint4 gq[U], yq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            gq[u] = __ldg(g + g0 + (j0 + u) * gHW);
            yq[u] = __ldg(y + q0 + (j0 + u) * HW);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float gv[8], yv[8];
            unpack8(gq[u], gv);
            unpack8(yq[u], yv);
#pragma unroll
            for (int e = 0; e < 8; ++e) t += gv[e] * yv[e];
            if constexpr (kKeep) {
                gk[j0 + u] = gq[u];
                yk[j0 + u] = yq[u];
            }
        }
    }
    t *= gscale / C;
#pragma unroll 1
    for (int j0 = 0; j0 < NCH; j0 += U) {
        uint4 gq[U], yq[U], aq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if constexpr (kKeep) {
                gq[u] = gk[j0 + u];
                yq[u] = yk[j0 + u];
            } else {
                gq[u] = __ldg(g + g0 + (j0 + u) * gHW);
                yq[u] = __ldg(y + q0 + (j0 + u) * HW);
            }
            if (addin) aq[u] = __ldg(addin + q0 + (j0 + u) * HW);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float gv[8], yv[8], o[8], ad[8];
            unpack8(gq[u], gv);
            unpack8(yq[u], yv);
            if (addin) unpack8(aq[u], ad);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                gv[e] *= gscale;
                o[e] = lrelu_mask(yv[e], leak) * rinv * (gv[e] - yv[e] * t) + (addin ? ad[e] : 0.f);
            }
            ga[q0 + (j0 + u) * HW] = pack8(o);
            if (gy_out) gy_out[q0 + (j0 + u) * HW] = pack8(gv);
        }
    }
}
int pn_bwd_c8(const void* g, int unpool, float gscale, const float* dyn, const void* y, const float* r, const void* addin, void* ga,
              void* gy_out, float leak, int B, int C, int H, int W, cudaStream_t st) {
    const size_t total = static_cast<size_t>(B) * H * W;
#define NGAN_PNB(N)                                                                                                \
    case N:                                                                                                        \
        pn_bwd_c8_kernel<N><<<nblocks(total, 128), 128, 0, st>>>(                                                  \
            static_cast<const uint4*>(g), unpool, gscale, dyn, static_cast<const uint4*>(y), r,                    \
            static_cast<const uint4*>(addin), static_cast<uint4*>(ga), static_cast<uint4*>(gy_out), leak, H, W,    \
            total);                                                                                                \
        break;
    switch (C / 8) {
        NGAN_PNB(2)
        NGAN_PNB(4)
        NGAN_PNB(8)
        NGAN_PNB(16)
        NGAN_PNB(32)
        default:
            set_error("pn_bwd: unsupported channel count %d (16, 32, 64, 128, 256 are built)", C);
            return NGAN_ERR_UNSUPPORTED;
    }
#undef NGAN_PNB
    return check_launch("pn_bwd_c8");
}

// Adjoint of the bilinear x2 upsample fused with the PixelNorm/LeakyReLU backward of the layer that fed it.
// g_up: [B][C/8][2H][2W][8]; y, r, ga at H x W.  extra_pre/extra_w: the faded-out ToImage branch adds
// extra_w[c] * extra_pre[pixel] to the gradient wrt y (generator transition, models.py:348).
// One thread per (column, 8-channel group) over RP consecutive low-resolution rows; the NCH threads of a pixel are
// adjacent lanes and combine their partial sums of mean_c(g*y) with xor-shuffles.  The adjoint is separable: every
// high-resolution row is first combined horizontally (4 granules -> 1) and that row sum feeds the two low rows it
// belongs to, so RP = 2 rows cost 6 x 4 loads instead of 2 x 16; the weights are the fixed pattern
// (.25, .75, .75, .25) with the clamped borders folded in ((0, 1, ...) at index 0, (..., 1, 0) at the last index).
// ncu (round 2) showed the one-pixel-per-thread version issue-bound: 734 instructions per thread, 70 % issue
// utilisation at 37 % of DRAM bandwidth.
// (A shared-memory-tiled variant -- the block stages the haloed high-resolution tile once instead of every pixel
// reading its 4 x 4 window through L1/L2 -- was measured in round 2: 74.6 us against 63.6 us at 512 -> 256; the
// windows of a warp already overlap in L1.)
__device__ __forceinline__ void up2_adj_pattern(int i, int n, float* w) {
    w[0] = i == 0 ? 0.f : 0.25f;
    w[1] = i == 0 ? 1.f : 0.75f;
    w[2] = i == n - 1 ? 1.f : 0.75f;
    w[3] = i == n - 1 ? 0.f : 0.25f;
}
template <int NCH, int RP>
__global__ void __launch_bounds__(128, 4) up2_bwd_pn_bwd_kernel(const uint4* __restrict__ g_up,
                                                                const uint4* __restrict__ y,
                                                                const float* __restrict__ r,
                                                                const float* __restrict__ extra_pre,
                                                                const float* __restrict__ extra_w,
                                                                uint4* __restrict__ ga, float leak, int H, int W) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    constexpr int C = NCH * 8, NR = 2 * RP + 2;
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = static_cast<int>(t % NCH);
    const int pxr = static_cast<int>(t / NCH);
    const bool okx = pxr < W;
    const int px = okx ? pxr : W - 1;          // out-of-range lanes still take part in the shuffles
    const int py0 = blockIdx.y * RP;
    const size_t b = blockIdx.z;
    const size_t HW = static_cast<size_t>(H) * W;
    const int UW = 2 * W, UH = 2 * H;
    float wx[4];
    up2_adj_pattern(px, W, wx);
    int cx[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cx[k] = min(max(2 * px - 1 + k, 0), UW - 1);     // (weight 0 where clamped)
    const uint4* p = g_up + (b * NCH + j) * 4 * HW;
    float g[RP][8];
#pragma unroll
    for (int o = 0; o < RP; ++o)
#pragma unroll
        for (int e = 0; e < 8; ++e) g[o][e] = 0.f;
    float wy[RP][4];
#pragma unroll
    for (int o = 0; o < RP; ++o) up2_adj_pattern(py0 + o, H, wy[o]);
#pragma unroll
    for (int rr = 0; rr < NR; ++rr) {
        const int row = min(max(2 * py0 - 1 + rr, 0), UH - 1);
        const uint4* prow = p + static_cast<size_t>(row) * UW;
        float h[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float v[8];
            unpack8(__ldg(prow + cx[k]), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) h[e] = k == 0 ? wx[0] * v[e] : fmaf(wx[k], v[e], h[e]);
        }
#pragma unroll
        for (int o = 0; o < RP; ++o) {
            const int a = rr - 2 * o;            // position of this high row in low row o's window
            if (a >= 0 && a < 4) {
#pragma unroll
                for (int e = 0; e < 8; ++e) g[o][e] = fmaf(wy[o][a], h[e], g[o][e]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < RP; ++o) {
        const size_t i = b * HW + static_cast<size_t>(py0 + o) * W + px;
        const size_t q = (b * NCH + j) * HW + static_cast<size_t>(py0 + o) * W + px;
        if (extra_pre) {
            const float ep = extra_pre[i];
#pragma unroll
            for (int e = 0; e < 8; ++e) g[o][e] = fmaf(__ldg(extra_w + j * 8 + e), ep, g[o][e]);
        }
        float yv[8], tt = 0.f;
        unpack8(__ldg(y + q), yv);
#pragma unroll
        for (int e = 0; e < 8; ++e) tt = fmaf(g[o][e], yv[e], tt);
#pragma unroll
        for (int s = 1; s < NCH; s <<= 1) tt += __shfl_xor_sync(0xffffffffu, tt, s);
        tt *= 1.0f / C;
        const float rinv = r[i];
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o8[e] = lrelu_mask(yv[e], leak) * rinv * (g[o][e] - yv[e] * tt);
        if (okx) ga[q] = pack8(o8);
    }
}
int up2_bwd_pn_bwd_c8(const void* g_up, const void* y, const float* r, const float* extra_pre, const float* extra_w,
                      void* ga, float leak, int B, int C, int H, int W, cudaStream_t st) {
    const int nch = C / 8;
    const int lanes = W * nch;
    const int threads = lanes < 128 ? (lanes + 31) / 32 * 32 : 128;
    // rows per thread, measured (512 -> 256, B = 16): 1 row 61.8 us, 2 rows 50.2 us, 4 rows 50.6 us (spills)
    const int rp = H % 2 == 0 ? 2 : 1;
    const dim3 grid((lanes + threads - 1) / threads, H / rp, B);
    if (B > 65535) {
        set_error("up2_bwd_pn_bwd: batch %d too large for one launch", B);
        return NGAN_ERR_UNSUPPORTED;
    }
#define NGAN_UP2B(CC)                                                                                            \
    case CC:                                                                                                     \
        if (rp == 2)                                                                                             \
            up2_bwd_pn_bwd_kernel<CC / 8, 2><<<grid, threads, 0, st>>>(                                          \
                static_cast<const uint4*>(g_up), static_cast<const uint4*>(y), r, extra_pre, extra_w,            \
                static_cast<uint4*>(ga), leak, H, W);                                                            \
        else                                                                                                     \
            up2_bwd_pn_bwd_kernel<CC / 8, 1><<<grid, threads, 0, st>>>(                                          \
                static_cast<const uint4*>(g_up), static_cast<const uint4*>(y), r, extra_pre, extra_w,            \
                static_cast<uint4*>(ga), leak, H, W);                                                            \
        break;
    switch (C) {
        NGAN_UP2B(16)
        NGAN_UP2B(32)
        NGAN_UP2B(64)
        NGAN_UP2B(128)
        default:
            set_error("up2_bwd_pn_bwd: unsupported channel count %d (16, 32, 64, 128 are built)", C);
            return NGAN_ERR_UNSUPPORTED;
    }
#undef NGAN_UP2B
    return check_launch("up2_bwd_pn_bwd_c8");
}

// ------------------------------------------------------------------------------- 1-channel fp32 image ops
__global__ void pool_image_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int W, size_t total) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int OW = W / 2, OH = H / 2;
    int ox, oy;
    size_t b;
    split_xyb(i, OW, OH, ox, oy, b);
    const float* p = x + b * H * W + static_cast<size_t>(2 * oy) * W + 2 * ox;
    out[i] = 0.25f * (p[0] + p[1] + p[W] + p[W + 1]);
}
// Vectorised variants of the 1-channel image kernels (row length a multiple of 4 floats, 16-byte aligned pointers,
// batch <= 65535): (column, row, sample) come from the 3-D grid, four floats per thread.  The scalar kernels spend
// a 64-bit division per ELEMENT on recovering (sample, row, column) from a flat index -- at 4 bytes per thread that
// made them issue-bound several times over (ncu: interp 15.7 us for 50 MB, unpool_image 15.7 us for 21 MB).
static bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}
__global__ void __launch_bounds__(128) pool_image_v4_kernel(const float4* __restrict__ x, float2* __restrict__ out,
                                                            int H, int W4) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;      // four input columns -> two outputs
    if (c >= W4) return;
    const size_t oy = blockIdx.y, b = blockIdx.z;
    const float4* p = x + (b * H + 2 * oy) * W4 + c;
    const float4 a = __ldg(p), d = __ldg(p + W4);
    out[(b * (H / 2) + oy) * W4 + c] = make_float2(0.25f * (a.x + a.y + d.x + d.y), 0.25f * (a.z + a.w + d.z + d.w));
}
int pool_image(const float* x, float* out, int B, int H, int W, cudaStream_t st) {
    if (W % 4 == 0 && W >= 128 && H % 2 == 0 && B <= 65535 && aligned16(x, out)) {
        const int W4 = W / 4, threads = W4 < 128 ? (W4 + 31) / 32 * 32 : 128;
        pool_image_v4_kernel<<<dim3((W4 + threads - 1) / threads, H / 2, B), threads, 0, st>>>(
            reinterpret_cast<const float4*>(x), reinterpret_cast<float2*>(out), H, W4);
        return check_launch("pool_image");
    }
    const size_t total = static_cast<size_t>(B) * (H / 2) * (W / 2);
    pool_image_kernel<<<nblocks(total, 256), 256, 0, st>>>(x, out, H, W, total);
    return check_launch("pool_image");
}
// out[b, y, x] = scale * g[b, y/2, x/2]   (H, W are the OUTPUT dims)
__global__ void unpool_image_kernel(const float* __restrict__ g, float* __restrict__ out, float scale, int H, int W,
                                    size_t total) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int x, y;
    size_t b;
    split_xyb(i, W, H, x, y, b);
    out[i] = scale * g[b * (H / 2) * (W / 2) + static_cast<size_t>(y >> 1) * (W / 2) + (x >> 1)];
}
__global__ void __launch_bounds__(128) unpool_image_v4_kernel(const float2* __restrict__ g, float4* __restrict__ out,
                                                              float scale, int H, int W4) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;      // two inputs -> four output columns of two rows
    if (c >= W4) return;
    const size_t gy = blockIdx.y, b = blockIdx.z;
    const float2 v = __ldg(g + (b * (H / 2) + gy) * W4 + c);
    const float4 o = make_float4(scale * v.x, scale * v.x, scale * v.y, scale * v.y);
    float4* q = out + (b * H + 2 * gy) * W4 + c;
    q[0] = o;
    q[W4] = o;
}
int unpool_image(const float* g, float* out, float scale, int B, int H, int W, cudaStream_t st) {
    if (W % 4 == 0 && W >= 128 && H % 2 == 0 && B <= 65535 && aligned16(g, out)) {
        const int W4 = W / 4, threads = W4 < 128 ? (W4 + 31) / 32 * 32 : 128;
        unpool_image_v4_kernel<<<dim3((W4 + threads - 1) / threads, H / 2, B), threads, 0, st>>>(
            reinterpret_cast<const float2*>(g), reinterpret_cast<float4*>(out), scale, H, W4);
        return check_launch("unpool_image");
    }
    const size_t total = static_cast<size_t>(B) * H * W;
    unpool_image_kernel<<<nblocks(total, 256), 256, 0, st>>>(g, out, scale, H, W, total);
    return check_launch("unpool_image");
}
__global__ void up2_image_kernel(const float* __restrict__ x, float* __restrict__ out, int H, int W, size_t total) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int OW = 2 * W, OH = 2 * H;
    int ox, oy;
    size_t b;
    split_xyb(i, OW, OH, ox, oy, b);
    int y0, y1, x0, x1;
    float ly, lx;
    up2_src(oy, H, y0, y1, ly);
    up2_src(ox, W, x0, x1, lx);
    const float* p = x + b * H * W;
    out[i] = (1.f - ly) * ((1.f - lx) * p[static_cast<size_t>(y0) * W + x0] + lx * p[static_cast<size_t>(y0) * W + x1]) +
             ly * ((1.f - lx) * p[static_cast<size_t>(y1) * W + x0] + lx * p[static_cast<size_t>(y1) * W + x1]);
}
// two input columns of one input row -> four output columns of two output rows (same taps as upsample2x_c8)
__global__ void __launch_bounds__(128) up2_image_v4_kernel(const float* __restrict__ x, float4* __restrict__ out, int H,
                                                           int W) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * c >= W) return;
    const int iy = blockIdx.y;
    const size_t b = blockIdx.z;
    const float* p = x + b * H * W;
    const int x0 = 2 * c, xm = max(x0 - 1, 0), x1 = x0 + 1, xp = min(x0 + 2, W - 1);
    const int ys[3] = {max(iy - 1, 0), iy, min(iy + 1, H - 1)};
    float h[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float* row = p + static_cast<size_t>(ys[r]) * W;
        const float a = __ldg(row + xm), b0 = __ldg(row + x0), b1 = __ldg(row + x1), d = __ldg(row + xp);
        h[r][0] = __fmaf_rn(0.75f, b0, 0.25f * a);
        h[r][1] = __fmaf_rn(0.75f, b0, 0.25f * b1);
        h[r][2] = __fmaf_rn(0.75f, b1, 0.25f * b0);
        h[r][3] = __fmaf_rn(0.75f, b1, 0.25f * d);
    }
    const int W2 = W / 2;                                    // float4 per output row
    float4* q = out + (b * 2 * H + 2 * iy) * W2 + c;
    q[0] = make_float4(__fmaf_rn(0.75f, h[1][0], 0.25f * h[0][0]), __fmaf_rn(0.75f, h[1][1], 0.25f * h[0][1]),
                       __fmaf_rn(0.75f, h[1][2], 0.25f * h[0][2]), __fmaf_rn(0.75f, h[1][3], 0.25f * h[0][3]));
    q[W2] = make_float4(__fmaf_rn(0.75f, h[1][0], 0.25f * h[2][0]), __fmaf_rn(0.75f, h[1][1], 0.25f * h[2][1]),
                        __fmaf_rn(0.75f, h[1][2], 0.25f * h[2][2]), __fmaf_rn(0.75f, h[1][3], 0.25f * h[2][3]));
}
int up2_image(const float* x, float* out, int B, int H, int W, cudaStream_t st) {
    if (W % 2 == 0 && W >= 64 && B <= 65535 && aligned16(out)) {
        const int cols = W / 2, threads = cols < 128 ? (cols + 31) / 32 * 32 : 128;
        up2_image_v4_kernel<<<dim3((cols + threads - 1) / threads, H, B), threads, 0, st>>>(
            x, reinterpret_cast<float4*>(out), H, W);
        return check_launch("up2_image");
    }
    const size_t total = static_cast<size_t>(B) * H * W * 4;
    up2_image_kernel<<<nblocks(total, 256), 256, 0, st>>>(x, out, H, W, total);
    return check_launch("up2_image");
}
// adjoint of up2_image: g is [B, 2H, 2W], out [B, H, W] = scale * U^T g
__global__ void up2_image_bwd_kernel(const float* __restrict__ g, float* __restrict__ out, float scale,
                                     const float* __restrict__ dyn, int H, int W, size_t total) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int px, py;
    size_t b;
    split_xyb(i, W, H, px, py, b);
    const float* p = g + b * 4 * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float wy = up2_adj_w(2 * py - 1 + a, py, H);
        if (wy == 0.f) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float wx = up2_adj_w(2 * px - 1 + c, px, W);
            if (wx == 0.f) continue;
            acc += wy * wx * p[static_cast<size_t>(2 * py - 1 + a) * (2 * W) + (2 * px - 1 + c)];
        }
    }
    out[i] = (dyn ? scale * __ldg(dyn) : scale) * acc;
}
// the same sum with (column, row, sample) from the grid and the fixed weight pattern (up2_adj_pattern)
__global__ void __launch_bounds__(128) up2_image_bwd_grid_kernel(const float* __restrict__ g, float* __restrict__ out,
                                                                 float scale, const float* __restrict__ dyn, int H,
                                                                 int W) {
    const int px = blockIdx.x * blockDim.x + threadIdx.x;
    if (px >= W) return;
    const int py = blockIdx.y;
    const size_t b = blockIdx.z;
    float wx[4], wy[4];
    up2_adj_pattern(px, W, wx);
    up2_adj_pattern(py, H, wy);
    const float* p = g + b * 4 * H * W;
    float acc = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float* row = p + static_cast<size_t>(min(max(2 * py - 1 + a, 0), 2 * H - 1)) * (2 * W);
#pragma unroll
        for (int c = 0; c < 4; ++c) acc += wy[a] * wx[c] * __ldg(row + min(max(2 * px - 1 + c, 0), 2 * W - 1));
    }
    out[(b * H + py) * W + px] = (dyn ? scale * __ldg(dyn) : scale) * acc;
}
int up2_image_bwd(const float* g, float* out, float scale, const float* dyn, int B, int H, int W, cudaStream_t st) {
    if (W >= 64 && B <= 65535) {
        const int threads = W < 128 ? (W + 31) / 32 * 32 : 128;
        up2_image_bwd_grid_kernel<<<dim3((W + threads - 1) / threads, H, B), threads, 0, st>>>(g, out, scale, dyn, H, W);
        return check_launch("up2_image_bwd");
    }
    const size_t total = static_cast<size_t>(B) * H * W;
    up2_image_bwd_kernel<<<nblocks(total, 256), 256, 0, st>>>(g, out, scale, dyn, H, W, total);
    return check_launch("up2_image_bwd");
}
__global__ void axpby_kernel(const float* __restrict__ a, float ca, const float* __restrict__ b, float cb,
                             float* __restrict__ out, size_t n) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ca * a[i] + (b ? cb * b[i] : 0.f);
}
__global__ void axpby_v4_kernel(const float4* __restrict__ a, float ca, const float4* __restrict__ b, float cb,
                                float4* __restrict__ out, size_t n4) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 x = a[i], y = b ? b[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    out[i] = make_float4(ca * x.x + (b ? cb * y.x : 0.f), ca * x.y + (b ? cb * y.y : 0.f),
                         ca * x.z + (b ? cb * y.z : 0.f), ca * x.w + (b ? cb * y.w : 0.f));
}
int axpby_f32(const float* a, float ca, const float* b, float cb, float* out, size_t n, cudaStream_t st) {
    if (n % 4 == 0 && n >= 4096 && aligned16(a, b, out)) {
        axpby_v4_kernel<<<nblocks(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(a), ca,
                                                            reinterpret_cast<const float4*>(b), cb,
                                                            reinterpret_cast<float4*>(out), n / 4);
        return check_launch("axpby_f32");
    }
    axpby_kernel<<<nblocks(n, 256), 256, 0, st>>>(a, ca, b, cb, out, n);
    return check_launch("axpby_f32");
}
// a + alpha*(b - a): generator fade-in of the two ToImage branches (models.py:350)
__global__ void lerp_kernel(const float* __restrict__ a, const float* __restrict__ b, float alpha_h,
                            const float* __restrict__ dyn, float* __restrict__ out, size_t n) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const float alpha = dyn ? alpha_h * __ldg(dyn) : alpha_h;
    if (i < n) out[i] = a[i] + alpha * (b[i] - a[i]);
}
__global__ void lerp_v4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float alpha_h,
                               const float* __restrict__ dyn, float4* __restrict__ out, size_t n4) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const float alpha = dyn ? alpha_h * __ldg(dyn) : alpha_h;
    if (i >= n4) return;
    const float4 x = a[i], y = b[i];
    out[i] = make_float4(x.x + alpha * (y.x - x.x), x.y + alpha * (y.y - x.y), x.z + alpha * (y.z - x.z),
                         x.w + alpha * (y.w - x.w));
}
int lerp_f32(const float* a, const float* b, float alpha, const float* dyn, float* out, size_t n, cudaStream_t st) {
    if (n % 4 == 0 && n >= 4096 && aligned16(a, b, out)) {
        lerp_v4_kernel<<<nblocks(n / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(a),
                                                           reinterpret_cast<const float4*>(b), alpha, dyn,
                                                           reinterpret_cast<float4*>(out), n / 4);
        return check_launch("lerp_f32");
    }
    lerp_kernel<<<nblocks(n, 256), 256, 0, st>>>(a, b, alpha, dyn, out, n);
    return check_launch("lerp_f32");
}
// x_hat = eps_b * real + (1 - eps_b) * fake (loss_functions.py:171)
__global__ void interp_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                              const float* __restrict__ eps, float* __restrict__ out, size_t per_sample, size_t n) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float e = eps[i / per_sample];
    out[i] = e * real[i] + (1.f - e) * fake[i];
}
__global__ void interp_v4_kernel(const float4* __restrict__ real, const float4* __restrict__ fake,
                                 const float* __restrict__ eps, float4* __restrict__ out, size_t n4) {
    const size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= n4) return;
    const size_t i = blockIdx.y * n4 + k;
    const float e = __ldg(eps + blockIdx.y);
    const float4 r = real[i], f = fake[i];
    out[i] = make_float4(e * r.x + (1.f - e) * f.x, e * r.y + (1.f - e) * f.y, e * r.z + (1.f - e) * f.z,
                         e * r.w + (1.f - e) * f.w);
}
int interp_images(const float* real, const float* fake, const float* eps, float* out, int B, size_t per_sample,
                  cudaStream_t st) {
    if (per_sample % 4 == 0 && per_sample >= 4096 && B <= 65535 && aligned16(real, fake, out)) {
        const size_t n4 = per_sample / 4;
        interp_v4_kernel<<<dim3(nblocks(n4, 256), B), 256, 0, st>>>(reinterpret_cast<const float4*>(real),
                                                                   reinterpret_cast<const float4*>(fake), eps,
                                                                   reinterpret_cast<float4*>(out), n4);
        return check_launch("interp_images");
    }
    const size_t n = static_cast<size_t>(B) * per_sample;
    interp_kernel<<<nblocks(n, 256), 256, 0, st>>>(real, fake, eps, out, per_sample, n);
    return check_launch("interp_images");
}
__global__ void scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ coeff, float scale,
                                  float* __restrict__ out, size_t per_sample, size_t n) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) out[i] = scale * coeff[i / per_sample] * x[i];
}
__global__ void scale_rows_v4_kernel(const float4* __restrict__ x, const float* __restrict__ coeff, float scale,
                                     float4* __restrict__ out, size_t n4) {
    const size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= n4) return;
    const size_t i = blockIdx.y * n4 + k;
    const float c = scale * __ldg(coeff + blockIdx.y);
    const float4 v = x[i];
    out[i] = make_float4(c * v.x, c * v.y, c * v.z, c * v.w);
}
int scale_rows_f32(const float* x, const float* coeff, float scale, float* out, int B, size_t per_sample,
                   cudaStream_t st) {
    if (per_sample % 4 == 0 && per_sample >= 4096 && B <= 65535 && aligned16(x, out)) {
        const size_t n4 = per_sample / 4;
        scale_rows_v4_kernel<<<dim3(nblocks(n4, 256), B), 256, 0, st>>>(reinterpret_cast<const float4*>(x), coeff, scale,
                                                                       reinterpret_cast<float4*>(out), n4);
        return check_launch("scale_rows_f32");
    }
    const size_t n = static_cast<size_t>(B) * per_sample;
    scale_rows_kernel<<<nblocks(n, 256), 256, 0, st>>>(x, coeff, scale, out, per_sample, n);
    return check_launch("scale_rows_f32");
}

// ------------------------------------------------------------------------------- FromImage (models.py:156-165)
// F[b, c, p] = w_c * xp[b, p] + b_c on the (already pooled, if the block pools) 1-channel image.
__global__ void fromim_fwd_kernel(const float* __restrict__ xp, const float* __restrict__ w,
                                  const float* __restrict__ bias, uint4* __restrict__ out, int C, size_t HW,
                                  size_t total) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t b, pix;
    split_bpix(i, HW, b, pix);
    const int nch = C / 8;
    const float xv = xp[i];
    for (int j = 0; j < nch; ++j) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = __ldg(w + j * 8 + e) * xv + __ldg(bias + j * 8 + e);
        out[(b * nch + j) * HW + pix] = pack8(o);
    }
}
int fromim_fwd(const float* xp, const float* w, const float* b, void* out, int B, int C, int H, int W,
               cudaStream_t st) {
    const size_t HW = static_cast<size_t>(H) * W, total = B * HW;
    fromim_fwd_kernel<<<nblocks(total, 256), 256, 0, st>>>(xp, w, b, static_cast<uint4*>(out), C, HW, total);
    return check_launch("fromim_fwd");
}
// Discriminator fade-in (models.py:519-521): y = y_start + alpha*(y_end - y_start), y_start = FromIm_old(xp)
__global__ void d_fade_fwd_kernel(const uint4* __restrict__ y_end, const float* __restrict__ xp,
                                  const float* __restrict__ w, const float* __restrict__ bias, float alpha_h,
                                  const float* __restrict__ dyn, uint4* __restrict__ out, int C, size_t HW,
                                  size_t total) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const float alpha = dyn ? alpha_h * __ldg(dyn) : alpha_h;
    size_t b, pix;
    split_bpix(i, HW, b, pix);
    const int nch = C / 8;
    const float xv = xp[i];
    for (int j = 0; j < nch; ++j) {
        float o[8], ye[8];
        unpack8(__ldg(y_end + (b * nch + j) * HW + pix), ye);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const float ys = __ldg(w + j * 8 + e) * xv + (bias ? __ldg(bias + j * 8 + e) : 0.f);
            o[e] = ys + alpha * (ye[e] - ys);
        }
        out[(b * nch + j) * HW + pix] = pack8(o);
    }
}
int d_fade_fwd(const void* y_end, const float* xp, const float* w_old, const float* b_old, float alpha,
               const float* dyn, void* out, int B, int C, int H, int W, cudaStream_t st) {
    const size_t HW = static_cast<size_t>(H) * W, total = B * HW;
    d_fade_fwd_kernel<<<nblocks(total, 256), 256, 0, st>>>(static_cast<const uint4*>(y_end), xp, w_old, b_old, alpha,
                                                            dyn, static_cast<uint4*>(out), C, HW, total);
    return check_launch("d_fade_fwd");
}

// Backward of FromImage: with G = gscale * g[(y,x) or (y/2,x/2)]:
//   gw[c] += sum G_c * xp,  gb[c] += sum G_c,  g_img[b,p] (+)= sum_c w_c * G_c
// Thread layout of the 1x1-conv backward kernels below: a block of 128 threads = (128 / NCH) pixel lanes x NCH
// channel groups, the NCH threads of a pixel being adjacent lanes.  The work is cut into chunks of consecutive pixels
// of ONE image (pixel_chunks: a block takes chunk blockIdx.x, blockIdx.x + gridDim.x, ...), so all addressing
// inside the loop is "base + pixel" in 32 bits -- ncu (round 2) showed these kernels issue-bound, not HBM-bound:
// fromim_bwd spent 125 warp-instructions per (16 pixels x 16 channels), most of them 64-bit index arithmetic and
// the two integer divisions that turned a flat index back into (image, row, column) for every pixel.
// Each thread keeps the per-channel sums of its group (weight / bias gradient) in registers; per pixel, sums over
// all channels go through xor-shuffles over the NCH lanes; the per-channel sums meet once per block in shared
// memory (summed in lane order) and leave as one row of the partials workspace, which reduce_partials adds in
// block order (no atomics: bit-reproducible gradients).
struct PixelChunks {
    unsigned chunk, per_img, n;
    int blocks;
};
static PixelChunks pixel_chunks(int B, size_t HW, int pixel_lanes, int resident_per_sm) {
    // One resident wave.  If the work fits it as one chunk per block (16, 8 or 4 pixels per thread -- the smaller
    // ones only to get at least two blocks per SM on the small maps), that is the grid.  Otherwise chunks of 8
    // pixels per thread and as many blocks as are resident at once, each walking several chunks: the load is then
    // balanced to ~10 %.  (Measured, toim_bwd at 512x512: 1024 one-chunk blocks at 6 resident per SM = 1.15 waves,
    // 82 us; 888 blocks over 4096 chunks 64 us.  fromim_bwd at 256x256: 1024 one-chunk blocks in one wave 14.3 us;
    // 888 blocks over 1024 chunks 18.1 us.)
    const unsigned cap = 148u * static_cast<unsigned>(resident_per_sm);
    PixelChunks c;
    for (int iters = 16;; iters /= 2) {
        c.chunk = static_cast<unsigned>(pixel_lanes * iters);
        c.per_img = static_cast<unsigned>((HW + c.chunk - 1) / c.chunk);
        c.n = static_cast<unsigned>(B) * c.per_img;
        if (c.n >= 2 * 148u || iters == 4) break;
    }
    if (c.n > cap) {
        c.chunk = static_cast<unsigned>(pixel_lanes * 8);
        c.per_img = static_cast<unsigned>((HW + c.chunk - 1) / c.chunk);
        c.n = static_cast<unsigned>(B) * c.per_img;
    }
    c.blocks = static_cast<int>(c.n < cap ? c.n : cap);       // every block ends with one partial row
    return c;
}
static int log2_exact(int v) {
    int l = 0;
    while ((1 << l) < v) ++l;
    return (1 << l) == v ? l : -1;
}
// index of pixel p of an H x W map in the half-resolution map (the adjoint of AvgPool2d(2) reads its cotangent there)
__device__ __forceinline__ unsigned unpool_index(unsigned p, int W, int wshift) {
    unsigned py, px;
    if (wshift >= 0) {
        py = p >> wshift;
        px = p & static_cast<unsigned>(W - 1);
    } else {
        py = p / static_cast<unsigned>(W);
        px = p - py * W;
    }
    return (py >> 1) * static_cast<unsigned>(W >> 1) + (px >> 1);
}
template <int NCH>
__global__ void __launch_bounds__(128) fromim_bwd_kernel(const uint4* __restrict__ g, int unpool, float gscale_h,
                                                         const float* __restrict__ dyn,
                                                         const float* __restrict__ xp, const float* __restrict__ w,
                                                         float* __restrict__ partials,
                                                         float* __restrict__ g_img, int accumulate, int H, int W,
                                                         int wshift, unsigned chunk, unsigned per_img,
                                                         unsigned n_chunks) {
    constexpr int PL = 128 / NCH, C = NCH * 8;
    constexpr int U = 4;              // pixels in flight per thread: all loads of a batch are issued before the first use
    const float gscale = dyn ? gscale_h * __ldg(dyn) : gscale_h;
    __shared__ float red[2][PL][C];
    const int pl = threadIdx.x / NCH, j = threadIdx.x % NCH;
    const unsigned HW = static_cast<unsigned>(H) * W;
    const unsigned gHW = unpool ? HW / 4 : HW;
    float wj[8], sw[8], sb[8];        // sums of the raw cotangent; gscale is applied once at the end
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        wj[e] = __ldg(w + j * 8 + e);
        sw[e] = sb[e] = 0.f;
    }
    for (unsigned cid = blockIdx.x; cid < n_chunks; cid += gridDim.x) {
        const unsigned b = cid / per_img;
        const unsigned p0 = (cid - b * per_img) * chunk;
        const unsigned p_end = min(HW, p0 + chunk);
        const uint4* gp = g + static_cast<size_t>(b * NCH + j) * gHW;
        const float* xb = xp + static_cast<size_t>(b) * HW;
        float* gi = g_img ? g_img + static_cast<size_t>(b) * HW : nullptr;
        for (unsigned q = p0; q < p_end; q += U * PL) {        // the same trip count for every thread (shuffles inside)
            uint4 gq4[U];
            float xv4[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned p = q + u * PL + pl;
                const bool ok = p < p_end;
                gq4[u] = ok ? __ldg(gp + (unpool ? unpool_index(p, W, wshift) : p)) : make_uint4(0, 0, 0, 0);
                xv4[u] = ok ? __ldg(xb + p) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned p = q + u * PL + pl;
                float gv[8];
                unpack8(gq4[u], gv);
                const float xv = xv4[u];
                float im = 0.f;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    sw[e] = fmaf(gv[e], xv, sw[e]);
                    sb[e] += gv[e];
                    im = fmaf(wj[e], gv[e], im);
                }
#pragma unroll
                for (int o = 1; o < NCH; o <<= 1) im += __shfl_xor_sync(0xffffffffu, im, o);
                if (gi && j == 0 && p < p_end) gi[p] = accumulate ? fmaf(gscale, im, gi[p]) : gscale * im;
            }
        }
    }
    if (!partials) return;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        red[0][pl][j * 8 + e] = gscale * sw[e];
        red[1][pl][j * 8 + e] = gscale * sb[e];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
        const int which = c / C, cc = c - which * C;
        float v = 0.f;
#pragma unroll 4
        for (int q = 0; q < PL; ++q) v += red[which][q][cc];
        partials[static_cast<size_t>(blockIdx.x) * 2 * C + c] = v;
    }
}
int fromim_bwd(const void* g, int unpool, float gscale, const float* dyn, const float* xp, const float* w, float* gw, float* gb,
               int grad_accumulate, float* workspace, float* g_img, int g_img_accumulate, int B, int C, int H, int W,
               cudaStream_t st) {
    float* partials = (gw || gb) ? workspace : nullptr;
    const PixelChunks pc = pixel_chunks(B, static_cast<size_t>(H) * W, 128 / (C / 8), 7);
    const int blocks = pc.blocks;
#define NGAN_FIB(N)                                                                                                  \
    case N:                                                                                                          \
        fromim_bwd_kernel<N><<<blocks, 128, 0, st>>>(                                                                \
            static_cast<const uint4*>(g), unpool, gscale, dyn, xp, w, partials, g_img, g_img_accumulate, H, W,       \
            log2_exact(W), pc.chunk, pc.per_img, pc.n);                                                              \
        break;
    switch (C / 8) {
        NGAN_FIB(2)
        NGAN_FIB(4)
        NGAN_FIB(8)
        NGAN_FIB(16)
        default:
            set_error("fromim_bwd: unsupported channel count %d (16, 32, 64, 128 are built)", C);
            return NGAN_ERR_UNSUPPORTED;
    }
#undef NGAN_FIB
    int rc = check_launch("fromim_bwd");
    if (rc || !partials) return rc;
    if (gw && gb) return reduce_partials(partials, blocks, 2 * C, 2 * C, 1.f, gw, grad_accumulate, st, gb, C);
    if (gw) rc = reduce_partials(partials, blocks, C, 2 * C, 1.f, gw, grad_accumulate, st);
    if (!rc && gb) rc = reduce_partials(partials + C, blocks, C, 2 * C, 1.f, gb, grad_accumulate, st);
    return rc;
}
// Double backward of FromImage's input-gradient: first order was g_xp = sum_c w_c * G_c.  With cotangent
// X = in_scale * ghat_xp on g_xp:  ghat_out[c] = w_c * X (cotangent on G_c),  what[c] += sum X * G_c.
// (thread layout and chunking: see fromim_bwd_kernel)
template <int NCH>
__global__ void __launch_bounds__(128) fromim_dbl_kernel(const float* __restrict__ ghat_xp, float in_scale,
                                                         const uint4* __restrict__ g, int unpool, float gscale_h,
                                                         const float* __restrict__ dyn, const float* __restrict__ w,
                                                         uint4* __restrict__ ghat_out, float* __restrict__ partials,
                                                         int H, int W, int wshift, unsigned chunk, unsigned per_img,
                                                         unsigned n_chunks) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    constexpr int PL = 128 / NCH, C = NCH * 8;
    constexpr int U = 4;
    const float gscale = dyn ? gscale_h * __ldg(dyn) : gscale_h;
    __shared__ float red[PL][C];
    const int pl = threadIdx.x / NCH, j = threadIdx.x % NCH;
    const unsigned HW = static_cast<unsigned>(H) * W;
    const unsigned gHW = unpool ? HW / 4 : HW;
    float wj[8], sw[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        wj[e] = __ldg(w + j * 8 + e);
        sw[e] = 0.f;
    }
    for (unsigned cid = blockIdx.x; cid < n_chunks; cid += gridDim.x) {
        const unsigned b = cid / per_img;
        const unsigned p0 = (cid - b * per_img) * chunk;
        const unsigned p_end = min(HW, p0 + chunk);
        const uint4* gp = g + static_cast<size_t>(b * NCH + j) * gHW;
        const float* xb = ghat_xp + static_cast<size_t>(b) * HW;
        uint4* ob = ghat_out ? ghat_out + static_cast<size_t>(b * NCH + j) * HW : nullptr;
        for (unsigned q = p0; q < p_end; q += U * PL) {
            uint4 gq4[U];
            float X4[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned p = q + u * PL + pl;
                const bool ok = p < p_end;
                gq4[u] = (ok && partials) ? __ldg(gp + (unpool ? unpool_index(p, W, wshift) : p)) : make_uint4(0, 0, 0, 0);
                X4[u] = ok ? in_scale * __ldg(xb + p) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned p = q + u * PL + pl;
                float gv[8], o[8];
                unpack8(gq4[u], gv);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    o[e] = wj[e] * X4[u];
                    sw[e] = fmaf(X4[u], gv[e], sw[e]);
                }
                if (ob && p < p_end) ob[p] = pack8(o);
            }
        }
    }
    if (!partials) return;
#pragma unroll
    for (int e = 0; e < 8; ++e) red[pl][j * 8 + e] = gscale * sw[e];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float v = 0.f;
#pragma unroll 4
        for (int q = 0; q < PL; ++q) v += red[q][c];
        partials[static_cast<size_t>(blockIdx.x) * C + c] = v;
    }
}
int fromim_dbl(const float* ghat_xp, float in_scale, const void* g, int unpool, float gscale, const float* dyn,
               const float* w,
               void* ghat_out, float* what, int grad_accumulate, float* workspace, int B, int C, int H, int W,
               cudaStream_t st) {
    float* partials = what ? workspace : nullptr;
    const PixelChunks pc = pixel_chunks(B, static_cast<size_t>(H) * W, 128 / (C / 8), 8);
    const int blocks = pc.blocks;
#define NGAN_FID(N)                                                                                                  \
    case N:                                                                                                          \
        fromim_dbl_kernel<N><<<blocks, 128, 0, st>>>(ghat_xp, in_scale, static_cast<const uint4*>(g), unpool, gscale, \
                                                     dyn, w, static_cast<uint4*>(ghat_out), partials, H, W,          \
                                                     log2_exact(W), pc.chunk, pc.per_img, pc.n);                     \
        break;
    switch (C / 8) {
        NGAN_FID(2)
        NGAN_FID(4)
        NGAN_FID(8)
        NGAN_FID(16)
        default:
            set_error("fromim_dbl: unsupported channel count %d (16, 32, 64, 128 are built)", C);
            return NGAN_ERR_UNSUPPORTED;
    }
#undef NGAN_FID
    int rc = check_launch("fromim_dbl");
    if (rc || !partials) return rc;
    return reduce_partials(partials, blocks, C, C, 1.f, what, grad_accumulate, st);
}

// ------------------------------------------------------------------------------- ToImage (models.py:133-149)
__global__ void toim_fwd_kernel(const uint4* __restrict__ y, const float* __restrict__ w, float* __restrict__ img,
                                int C, size_t HW, size_t total) {
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    size_t b, pix;
    split_bpix(i, HW, b, pix);
    const int nch = C / 8;
    float acc = 0.f;
    for (int j = 0; j < nch; ++j) {
        float yv[8];
        unpack8(__ldg(y + (b * nch + j) * HW + pix), yv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += __ldg(w + j * 8 + e) * yv[e];
    }
    img[i] = tanhf(acc);
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = __float2bfloat16(src[i]);
}
int f32_to_bf16(const float* src, void* dst, size_t n, cudaStream_t st) {
    f32_to_bf16_kernel<<<nblocks(n, 256), 256, 0, st>>>(src, static_cast<__nv_bfloat16*>(dst), n);
    return check_launch("f32_to_bf16");
}
int toim_fwd(const void* y, const float* w, float* img, int B, int C, int H, int W, cudaStream_t st) {
    const size_t HW = static_cast<size_t>(H) * W, total = B * HW;
    toim_fwd_kernel<<<nblocks(total, 256), 256, 0, st>>>(static_cast<const uint4*>(y), w, img, C, HW, total);
    return check_launch("toim_fwd");
}
// Backward of tanh(conv1x1(y)): gpre = gscale*g_img*(1-img^2); gw[c] += sum gpre*y_c; gy_c = w_c*gpre, then
// (if ga != null) the PixelNorm/LeakyReLU backward of the layer that produced y.
// (thread layout: see fromim_bwd_kernel)
template <int NCH>
__global__ void __launch_bounds__(128) toim_bwd_kernel(const float* __restrict__ g_img, float gscale_h,
                                                       const float* __restrict__ dyn,
                                                       const float* __restrict__ img, const uint4* __restrict__ y,
                                                       const float* __restrict__ r, const float* __restrict__ w,
                                                       uint4* __restrict__ ga, float* __restrict__ gpre_out,
                                                       float* __restrict__ partials, float leak, unsigned HW,
                                                       unsigned chunk, unsigned per_img, unsigned n_chunks) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    constexpr int PL = 128 / NCH, C = NCH * 8;
    constexpr int U = 4;                                        // pixels in flight per thread (see fromim_bwd_kernel)
    const float gscale = dyn ? gscale_h * __ldg(dyn) : gscale_h;
    __shared__ float red[PL][C];
    const int pl = threadIdx.x / NCH, j = threadIdx.x % NCH;
    float wj[8], sw[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        wj[e] = __ldg(w + j * 8 + e);
        sw[e] = 0.f;
    }
    for (unsigned cid = blockIdx.x; cid < n_chunks; cid += gridDim.x) {
        const unsigned b = cid / per_img;
        const unsigned p0 = (cid - b * per_img) * chunk;
        const unsigned p_end = min(HW, p0 + chunk);
        const size_t img_off = static_cast<size_t>(b) * HW, feat_off = static_cast<size_t>(b * NCH + j) * HW;
        const float* imb = img + img_off;
        const float* gib = g_img + img_off;
        const float* rb = r ? r + img_off : nullptr;
        const uint4* yb = y + feat_off;
        uint4* gab = ga ? ga + feat_off : nullptr;
        float* gpo = gpre_out ? gpre_out + img_off : nullptr;
        for (unsigned q = p0; q < p_end; q += U * PL) {        // the same trip count for every thread (shuffles inside)
            uint4 yq4[U];
            float im4[U], gi4[U], r4[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned p = q + u * PL + pl;
                const bool ok = p < p_end;
                im4[u] = ok ? __ldg(imb + p) : 0.f;
                gi4[u] = ok ? __ldg(gib + p) : 0.f;
                yq4[u] = ok ? __ldg(yb + p) : make_uint4(0, 0, 0, 0);
                r4[u] = (ok && gab) ? __ldg(rb + p) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned p = q + u * PL + pl;
                const bool ok = p < p_end;
                const float im = im4[u];
                const float gpre = gscale * gi4[u] * (1.f - im * im);      // 0 for an out-of-range lane (gi = 0)
                if (gpo && ok && j == 0) gpo[p] = gpre;
                float yv[8], t = 0.f;
                unpack8(yq4[u], yv);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    sw[e] = fmaf(gpre, yv[e], sw[e]);
                    t = fmaf(wj[e], yv[e], t);
                }
                if (gab) {
#pragma unroll
                    for (int o = 1; o < NCH; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                    const float tp = t * gpre * (1.0f / C);
                    const float rinv = r4[u];
                    float o8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) o8[e] = lrelu_mask(yv[e], leak) * rinv * (wj[e] * gpre - yv[e] * tp);
                    if (ok) gab[p] = pack8(o8);
                }
            }
        }
    }
    if (partials) {
#pragma unroll
        for (int e = 0; e < 8; ++e) red[pl][j * 8 + e] = sw[e];
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float v = 0.f;
#pragma unroll 4
            for (int q = 0; q < PL; ++q) v += red[q][c];
            partials[static_cast<size_t>(blockIdx.x) * C + c] = v;
        }
    }
}
int toim_bwd(const float* g_img, float gscale, const float* dyn, const float* img, const void* y, const float* r, const float* w,
             void* ga, float* gpre, float* gw, int grad_accumulate, float* workspace, float leak, int B, int C, int H,
             int W, cudaStream_t st) {
    const size_t HW = static_cast<size_t>(H) * W;
    float* partials = gw ? workspace : nullptr;
    const PixelChunks pc = pixel_chunks(B, HW, 128 / (C / 8), 6);
    const int blocks = pc.blocks;
#define NGAN_TIB(N)                                                                                                 \
    case N:                                                                                                         \
        toim_bwd_kernel<N><<<blocks, 128, 0, st>>>(                                                                 \
            g_img, gscale, dyn, img, static_cast<const uint4*>(y), r, w, static_cast<uint4*>(ga), gpre, partials,   \
            leak, static_cast<unsigned>(HW), pc.chunk, pc.per_img, pc.n);                                           \
        break;
    switch (C / 8) {
        NGAN_TIB(2)
        NGAN_TIB(4)
        NGAN_TIB(8)
        NGAN_TIB(16)
        default:
            set_error("toim_bwd: unsupported channel count %d (16, 32, 64, 128 are built)", C);
            return NGAN_ERR_UNSUPPORTED;
    }
#undef NGAN_TIB
    int rc = check_launch("toim_bwd");
    if (rc || !partials) return rc;
    return reduce_partials(partials, blocks, C, C, 1.f, gw, grad_accumulate, st);
}

// ------------------------------------------------------------------------------- critic head (models.py:485-490)
// score[b] = scale * sum_{c,p} w[c,p] * y[b,c,p] + bias, w in torch layout [1][C][S][S].
__global__ void head_fwd_kernel(const uint4* __restrict__ y, const float* __restrict__ w,
                                const float* __restrict__ bias, float scale, float* __restrict__ score, int C,
                                int HW) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    const int nch = C / 8;
    float acc = 0.f;
    for (int k = threadIdx.x; k < nch * HW; k += blockDim.x) {
        const int j = k / HW, pix = k % HW;
        float yv[8];
        unpack8(__ldg(y + static_cast<size_t>(b) * nch * HW + k), yv);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc += __ldg(w + static_cast<size_t>(j * 8 + e) * HW + pix) * yv[e];
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) score[b] = scale * v + bias[0];
    }
}
int head_fwd(const void* y, const float* w, const float* bias, float scale, float* score, int B, int C, int S,
             cudaStream_t st) {
    head_fwd_kernel<<<B, 1024, 0, st>>>(static_cast<const uint4*>(y), w, bias, scale, score, C, S * S);
    return check_launch("head_fwd");
}
// gy[b,c,p] = scale * w[c,p] * gout[b]; then PixelNorm/LeakyReLU backward of the layer that produced y.
// One thread per (pixel, 8-channel group), xor-shuffle reduction over the NCH lanes of a pixel (the tensor is only
// [B][128][16][16]: one thread per pixel meant 32 CTAs of threads that walk 16 groups twice).
template <int NCH>
__global__ void __launch_bounds__(128) head_bwd_pn_kernel(const float* __restrict__ gout, const float* __restrict__ w,
                                                          float scale, const uint4* __restrict__ y,
                                                          const float* __restrict__ r, uint4* __restrict__ ga,
                                                          uint4* __restrict__ gy_out, float leak, int HW, size_t total) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t pix_id = tid / NCH;
    const int j = static_cast<int>(tid % NCH);
    const bool ok = pix_id < total;
    const size_t i = ok ? pix_id : 0;
    constexpr int C = NCH * 8;
    const size_t b = i / HW;
    const int pix = static_cast<int>(i - b * HW);
    const float go = scale * gout[b];
    const size_t q = (b * NCH + j) * HW + pix;
    float yv[8], gv[8], t = 0.f;
    unpack8(__ldg(y + q), yv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        gv[e] = go * __ldg(w + static_cast<size_t>(j * 8 + e) * HW + pix);
        t += gv[e] * yv[e];
    }
#pragma unroll
    for (int o = 1; o < NCH; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    t *= 1.0f / C;
    const float rinv = r[i];
    float o8[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o8[e] = lrelu_mask(yv[e], leak) * rinv * (gv[e] - yv[e] * t);
    if (ok) {
        ga[q] = pack8(o8);
        if (gy_out) gy_out[q] = pack8(gv);
    }
}
int head_bwd_pn(const float* gout, const float* w, float scale, const void* y, const float* r, void* ga,
                void* gy_out, float leak, int B, int C, int S, cudaStream_t st) {
    const size_t total = static_cast<size_t>(B) * S * S;
    const int blocks = nblocks(total * (C / 8), 128);
#define NGAN_HBP(N)                                                                                                  \
    case N:                                                                                                          \
        head_bwd_pn_kernel<N><<<blocks, 128, 0, st>>>(gout, w, scale, static_cast<const uint4*>(y), r,               \
                                                      static_cast<uint4*>(ga), static_cast<uint4*>(gy_out), leak,    \
                                                      S * S, total);                                                 \
        break;
    switch (C / 8) {
        NGAN_HBP(2)
        NGAN_HBP(4)
        NGAN_HBP(8)
        NGAN_HBP(16)
        default:
            set_error("head_bwd_pn: unsupported channel count %d (16, 32, 64, 128 are built)", C);
            return NGAN_ERR_UNSUPPORTED;
    }
#undef NGAN_HBP
    return check_launch("head_bwd_pn");
}
// gw[c,p] += scale * sum_b coeff[b] * t[b,c,p]: head weight gradient (t = y, coeff = gout) and its double-backward
// twin (t = cotangent on gy, coeff = gout).
// One thread per output element walks the batch in order (no atomics: concurrent contributions to one parameter
// go to separate gradient slots, train_step.py).  gb (optional): the head bias gradient, gb[0] (+)= sum_b coeff[b].
__global__ void head_wgrad_kernel(const uint4* __restrict__ t, const float* __restrict__ coeff, float scale,
                                  float* __restrict__ gw, float* __restrict__ gb, int accumulate, int B, int C,
                                  int HW) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int nch = C / 8;
    if (gb && k == 0) {
        float sum = 0.f;
        for (int b = 0; b < B; ++b) sum += coeff[b];
        gb[0] = accumulate ? gb[0] + sum : sum;
    }
    if (k >= nch * HW) return;
    const int j = k / HW, pix = k % HW;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b) {
        float tv[8];
        unpack8(__ldg(t + static_cast<size_t>(b) * nch * HW + k), tv);
        const float c = coeff[b];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += c * tv[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float* dst = gw + static_cast<size_t>(j * 8 + e) * HW + pix;
        *dst = accumulate ? *dst + scale * acc[e] : scale * acc[e];
    }
}
int head_wgrad(const void* t, const float* coeff, float scale, float* gw, float* gb, int accumulate, int B, int C,
               int S, cudaStream_t st) {
    const int n = (C / 8) * S * S;
    head_wgrad_kernel<<<nblocks(n, 128), 128, 0, st>>>(static_cast<const uint4*>(t), coeff, scale, gw, gb, accumulate,
                                                       B, C, S * S);
    return check_launch("head_wgrad");
}

// gb[c] += sum over batch and pixels of ga[b,c,p]   (bias of the 128->128 conv, models.py:469-471)
__global__ void bias_grad_c8_kernel(const uint4* __restrict__ ga, float* __restrict__ partials, int C, size_t HW,
                                    size_t total) {
    extern __shared__ float sacc[];   // [8 warps][C]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const bool ok = i < total;
    const size_t ii = ok ? i : 0;
    size_t b, pix;
    split_bpix(ii, HW, b, pix);
    const int nch = C / 8;
    for (int j = 0; j < nch; ++j) {
        float v[8];
        unpack8(ok ? __ldg(ga + (b * nch + j) * HW + pix) : make_uint4(0, 0, 0, 0), v);
        warp_store8(v, sacc + warp * C + j * 8, lane);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < C; k += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += sacc[q * C + k];
        partials[static_cast<size_t>(blockIdx.x) * C + k] = v;
    }
}
int bias_grad_c8(const void* ga, float* gb, int accumulate, float* workspace, int B, int C, int H, int W,
                 cudaStream_t st) {
    const size_t HW = static_cast<size_t>(H) * W, total = B * HW;
    const int blocks = nblocks(total, 256);
    bias_grad_c8_kernel<<<blocks, 256, 8 * C * sizeof(float), st>>>(static_cast<const uint4*>(ga), workspace, C, HW,
                                                                    total);
    int rc = check_launch("bias_grad_c8");
    if (rc) return rc;
    return reduce_partials(workspace, blocks, C, C, 1.f, gb, accumulate, st);
}

// ------------------------------------------------------------------------------- loss reductions
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int k = 0; k < (blockDim.x >> 5); ++k) s += red[k];
    return s;
}
// D_W_loss (loss_functions.py:14-47): out3 = {D_loss, score_real, score_fake}; g_* = d(gscale*D_loss)/d score
__global__ void wloss_kernel(const float* __restrict__ s_real, const float* __restrict__ s_fake, float drift,
                             float* __restrict__ out3, float* __restrict__ g_real, float* __restrict__ g_fake,
                             float gscale, int B) {
    __shared__ float red[32];
    float a = 0.f, f = 0.f, q = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        a += s_real[b];
        f += s_fake[b];
        q += s_real[b] * s_real[b];
        if (g_real) g_real[b] = gscale * (-1.f + 2.f * drift * s_real[b]) / B;
        if (g_fake) g_fake[b] = gscale / B;
    }
    a = block_sum(a, red);
    f = block_sum(f, red);
    q = block_sum(q, red);
    if (threadIdx.x == 0) {
        out3[1] = a / B;
        out3[2] = f / B;
        out3[0] = -a / B + f / B + drift * q / B;
    }
}
int wloss_fwd(const float* s_real, const float* s_fake, float drift, float* out3, float* g_real, float* g_fake,
              float gscale, int B, cudaStream_t st) {
    wloss_kernel<<<1, 256, 0, st>>>(s_real, s_fake, drift, out3, g_real, g_fake, gscale, B);
    return check_launch("wloss_fwd");
}
// G_W_loss (loss_functions.py:59-74)
__global__ void gloss_kernel(const float* __restrict__ s_fake, float* __restrict__ out1, float* __restrict__ g_fake,
                             float gscale, int B) {
    __shared__ float red[32];
    float f = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        f += s_fake[b];
        if (g_fake) g_fake[b] = -gscale / B;
    }
    f = block_sum(f, red);
    if (threadIdx.x == 0) out1[0] = -f / B;
}
int gloss_fwd(const float* s_fake, float* out1, float* g_fake, float gscale, int B, cudaStream_t st) {
    gloss_kernel<<<1, 256, 0, st>>>(s_fake, out1, g_fake, gscale, B);
    return check_launch("gloss_fwd");
}
// The iteration's statistics as one packed tensor (the six .item() reads of train.py:389-394 become one D2H):
// stats = {D_loss + penalty (train.py:362), score_real, score_fake, G_loss, penalty}
__global__ void pack_stats_kernel(const float* __restrict__ out3, const float* __restrict__ out1,
                                  const float* __restrict__ pen, float* __restrict__ stats) {
    if (threadIdx.x == 0) {
        const float p = pen[0];
        stats[0] = out3[0] + p;
        stats[1] = out3[1];
        stats[2] = out3[2];
        stats[3] = out1[0];
        stats[4] = p;
    }
}
int pack_stats(const float* out3, const float* out1, const float* pen, float* stats, cudaStream_t st) {
    pack_stats_kernel<<<1, 32, 0, st>>>(out3, out1, pen, stats);
    return check_launch("pack_stats");
}
// Gradient penalty (loss_functions.py:176): norm_b = norm_scale * ||g_b||_2, pen = lambda*mean((norm_b-1)^2),
// coeff_b = gscale * d pen / d norm_b / norm_b  (so that d pen/d g_x = coeff_b * g_x).
__global__ void gp_sumsq_kernel(const float* __restrict__ g, float* __restrict__ partial /* [B][gridDim.x] */,
                                size_t per_sample) {
    __shared__ float red[32];
    const int b = blockIdx.y;
    const float* p = g + static_cast<size_t>(b) * per_sample;
    float s = 0.f;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < per_sample;
         i += static_cast<size_t>(gridDim.x) * blockDim.x)
        s += p[i] * p[i];
    s = block_sum(s, red);
    if (threadIdx.x == 0) partial[b * gridDim.x + blockIdx.x] = s;
}
// The per-sample partial sums are added in block order (deterministic).  A sample whose input gradient is exactly
// zero gets coefficient 0 -- torch's norm backward yields the zero subgradient there (loss_functions.py:176) --
// instead of -inf, which the double backward would turn into NaN.
__global__ void gp_finish_kernel(const float* __restrict__ partial, int n_part, float* __restrict__ coeff,
                                 float norm_scale, float lambda, float* __restrict__ pen_out, float gscale, int B) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float ss = 0.f;
        for (int k = 0; k < n_part; ++k) ss += partial[b * n_part + k];
        const float n = norm_scale * sqrtf(ss);
        acc += (n - 1.f) * (n - 1.f);
        coeff[b] = n > 0.f ? gscale * lambda * 2.f * (n - 1.f) / (B * n) : 0.f;
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) pen_out[0] = lambda * acc / B;
}
int gp_loss(const float* g, float norm_scale, float lambda, float* pen_out, float* coeff_out, float gscale,
            float* workspace /* B * 64 floats */, int B, size_t per_sample, cudaStream_t st) {
    int bx = static_cast<int>((per_sample + 1023) / 1024);
    if (bx > 64) bx = 64;
    gp_sumsq_kernel<<<dim3(bx, B), 256, 0, st>>>(g, workspace, per_sample);
    gp_finish_kernel<<<1, 256, 0, st>>>(workspace, bx, coeff_out, norm_scale, lambda, pen_out, gscale, B);
    return check_launch("gp_loss");
}

// ------------------------------------------------------------------------------- similarity_loss (loss_functions.py:185-205)
// Lambda / (B*(B-1)) * sum_ij (cos(z_i, z_j) - cos(x_i, x_j))^2 over a batch of images x [B][P] and latents z [B][L].
// Stage 1: every block takes one chunk of pixels of ALL images into shared memory and writes its partial B x B Gram
// matrix of the images as one row of the workspace (no atomics).  Stage 2 (one block): adds the rows in order, forms
// the latent Gram directly, normalises both with their diagonals and reduces the squared difference.
constexpr int kSimChunk = 512;      // pixels per block and image
__global__ void __launch_bounds__(256) sim_gram_kernel(const float* __restrict__ x, int B, long long P,
                                                       float* __restrict__ partial /* [gridDim.x][B*B] */) {
    extern __shared__ float tile[];                   // [B][kSimChunk + 1]
    const long long p0 = static_cast<long long>(blockIdx.x) * kSimChunk;
    const int n = static_cast<int>(P - p0 < kSimChunk ? P - p0 : kSimChunk);
    for (int i = threadIdx.x; i < B * kSimChunk; i += blockDim.x) {
        const int b = i / kSimChunk, k = i - b * kSimChunk;
        tile[b * (kSimChunk + 1) + k] = k < n ? __ldg(x + static_cast<long long>(b) * P + p0 + k) : 0.f;
    }
    __syncthreads();
    for (int pr = threadIdx.x; pr < B * B; pr += blockDim.x) {
        const int i = pr / B, j = pr - i * B;
        float acc = 0.f;
        if (j >= i) {                                 // symmetric: the lower triangle is mirrored by stage 2
            const float* a = tile + i * (kSimChunk + 1);
            const float* c = tile + j * (kSimChunk + 1);
#pragma unroll 8
            for (int k = 0; k < kSimChunk; ++k) acc = fmaf(a[k], c[k], acc);
        }
        partial[static_cast<size_t>(blockIdx.x) * B * B + pr] = acc;
    }
}
__global__ void __launch_bounds__(256) sim_finish_kernel(const float* __restrict__ partial, int n_chunks,
                                                         const float* __restrict__ z, int B, int L, float lambda,
                                                         float* __restrict__ gram /* scratch [2][B*B] */,
                                                         float* __restrict__ out) {
    __shared__ float red[32];
    for (int pr = threadIdx.x; pr < B * B; pr += blockDim.x) {
        const int i = pr / B, j = pr - i * B;
        const int lo = i < j ? i : j, hi = i < j ? j : i;
        float gx = 0.f;
        for (int c = 0; c < n_chunks; ++c) gx += partial[static_cast<size_t>(c) * B * B + lo * B + hi];
        float gz = 0.f;
        for (int k = 0; k < L; ++k) gz = fmaf(__ldg(z + lo * L + k), __ldg(z + hi * L + k), gz);
        gram[pr] = gx;
        gram[B * B + pr] = gz;
    }
    __syncthreads();
    float acc = 0.f;
    for (int pr = threadIdx.x; pr < B * B; pr += blockDim.x) {
        const int i = pr / B, j = pr - i * B;
        const float cx = gram[pr] / (sqrtf(gram[i * B + i]) * sqrtf(gram[j * B + j]));
        const float cz = gram[B * B + pr] / (sqrtf(gram[B * B + i * B + i]) * sqrtf(gram[B * B + j * B + j]));
        acc += (cz - cx) * (cz - cx);
    }
    acc = block_sum(acc, red);
    if (threadIdx.x == 0) out[0] = lambda * acc / (static_cast<float>(B) * (B - 1));
}
size_t similarity_workspace_bytes(int B, long long P) {
    const long long n_chunks = (P + kSimChunk - 1) / kSimChunk;
    return static_cast<size_t>(n_chunks + 2) * B * B * sizeof(float);
}
int similarity_loss(const float* x, const float* z, float lambda, float* workspace, float* out, int B, long long P,
                    int L, cudaStream_t st) {
    const int n_chunks = static_cast<int>((P + kSimChunk - 1) / kSimChunk);
    const size_t smem = static_cast<size_t>(B) * (kSimChunk + 1) * sizeof(float);
    if (smem > 200u * 1024) {
        set_error("similarity_loss: batch %d too large for one shared-memory tile (max %d)", B,
                  static_cast<int>(200u * 1024 / ((kSimChunk + 1) * sizeof(float))));
        return NGAN_ERR_UNSUPPORTED;
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sim_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(sim_gram)");
        configured = true;
    }
    sim_gram_kernel<<<n_chunks, 256, smem, st>>>(x, B, P, workspace);
    float* gram = workspace + static_cast<size_t>(n_chunks) * B * B;
    sim_finish_kernel<<<1, 256, 0, st>>>(workspace, n_chunks, z, B, L, lambda, gram, out);
    return check_launch("similarity_loss");
}

}  // namespace ngan
