// Multi-tensor Adam: one launch updates every active parameter of a network (reference train.py:220-225,
// torch.optim.Adam(lr, betas=(beta1, 0.999)), eps 1e-8, no weight decay, no amsgrad):
//   m = m + (1-b1)(g - m);  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// Each parameter keeps its own step count t (torch skips parameters whose grad is None, so counts differ
// between blocks that became active at different resolution phases); the two bias-correction scalars are
// computed on the host in double precision, like torch does, and travel in the descriptor table, which is
// passed by value in kernel-parameter space (no device-side table to maintain).
// Optionally refreshes a bf16 shadow copy of the parameter in the same pass (the generator's linear weight).
#include <cstdlib>

#include "common.cuh"
#include "conv_args.cuh"
#include "kernels.h"

namespace ngan {

constexpr int kAdamMaxTensors = 48;   // 48 x 80 B: the table travels in (4 KB) kernel-parameter space
struct AdamEntry {
    float* p;
    const float* g;
    float* m;
    float* v;
    __nv_bfloat16* shadow;
    long long n;
    float step_size;      // lr / (1 - beta1^t)
    float inv_bc2_sqrt;   // 1 / sqrt(1 - beta2^t)
    int shadow_k, shadow_c, shadow_ss;   // kind 1: dims of the linear operand image [SS][K/8][C][8] (linear.cu)
                                         // kind 2: {cin, cout, folded} of a 3x3 conv weight (conv_args.cuh)
    int shadow_kind;                     // 0 = same layout as p, 1 = linear image, 2 = conv fwd + dgrad images
    const float* dyn;     // optional device pointer to {step_size, inv_bc2_sqrt}: overrides the two fields above, so a
                          // launch captured in a CUDA graph picks up the values of the current step at replay
};
struct AdamTable {
    AdamEntry e[kAdamMaxTensors];
};

// One element of the update, with every rounding spelled out: the multi-tensor kernel and the factored Linear kernel
// must produce the same bits from the same gradient (no compiler-chosen FMA contraction).
__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float beta1, float beta2, float eps,
                                      float step_size, float inv_bc2_sqrt) {
    m = __fmaf_rn(1.f - beta1, __fsub_rn(g, m), m);
    v = __fmaf_rn(1.f - beta2, __fmul_rn(g, g), __fmul_rn(beta2, v));
    const float denom = __fmaf_rn(sqrtf(v), inv_bc2_sqrt, eps);
    p = __fmaf_rn(-step_size, __fdividef(m, denom), p);
}

__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ AdamTable tab, float beta1,
                                                         float beta2, float eps) {
    const AdamEntry& t = tab.e[blockIdx.y];
    const float step_size = t.dyn ? __ldg(t.dyn) : t.step_size;
    const float inv_bc2_sqrt = t.dyn ? __ldg(t.dyn + 1) : t.inv_bc2_sqrt;
    const long long n4 = t.n >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
#define NGAN_ADAM1(c) adam1(p.c, g.c, m.c, v.c, beta1, beta2, eps, step_size, inv_bc2_sqrt);
    if (t.shadow && t.shadow_kind == 1 && (t.shadow_c & 1) == 0) {
        // The linear weight [C*SS][K] with its operand image [SS][K/8][C][8]: one thread updates the 8 k's of two
        // channels c, c+1 of one pixel -- rows f and f + SS of the master, 32 contiguous bytes each -- and writes
        // their two adjacent 16-byte granules of the image as one full 32-byte sector (4-element threads wrote
        // quarter sectors: +27 % traffic on the 16.8 M-parameter tensor).
        const int K = t.shadow_k, C = t.shadow_c, SS = t.shadow_ss, KG = K >> 3;
        const long long items = static_cast<long long>(C / 2) * SS * KG;
        for (long long it = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; it < items; it += stride) {
            const int kg = static_cast<int>(it % KG);
            const long long rest = it / KG;
            const int px = static_cast<int>(rest % SS), cp = static_cast<int>(rest / SS);
            uint4 img[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long i = ((static_cast<long long>(2 * cp + h) * SS + px) * K + kg * 8) >> 2;   // float4 index
                uint32_t packed[4];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    float4 p = reinterpret_cast<float4*>(t.p)[i + q];
                    const float4 g = __ldg(reinterpret_cast<const float4*>(t.g) + i + q);
                    float4 m = reinterpret_cast<float4*>(t.m)[i + q];
                    float4 v = reinterpret_cast<float4*>(t.v)[i + q];
                    NGAN_ADAM1(x) NGAN_ADAM1(y) NGAN_ADAM1(z) NGAN_ADAM1(w)
                    reinterpret_cast<float4*>(t.p)[i + q] = p;
                    reinterpret_cast<float4*>(t.m)[i + q] = m;
                    reinterpret_cast<float4*>(t.v)[i + q] = v;
                    packed[2 * q] = pack_bf16(p.x, p.y);
                    packed[2 * q + 1] = pack_bf16(p.z, p.w);
                }
                img[h] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
            uint4* dst = reinterpret_cast<uint4*>(t.shadow) + (static_cast<long long>(px) * KG + kg) * C + 2 * cp;
            dst[0] = img[0];
            dst[1] = img[1];
        }
        return;
    }
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 p = reinterpret_cast<float4*>(t.p)[i];
        const float4 g = __ldg(reinterpret_cast<const float4*>(t.g) + i);
        float4 m = reinterpret_cast<float4*>(t.m)[i];
        float4 v = reinterpret_cast<float4*>(t.v)[i];
        NGAN_ADAM1(x) NGAN_ADAM1(y) NGAN_ADAM1(z) NGAN_ADAM1(w)
        reinterpret_cast<float4*>(t.p)[i] = p;
        reinterpret_cast<float4*>(t.m)[i] = m;
        reinterpret_cast<float4*>(t.v)[i] = v;
        if (t.shadow) {
            uint2 o;
            o.x = pack_bf16(p.x, p.y);
            o.y = pack_bf16(p.z, p.w);
            size_t dst = static_cast<size_t>(i) * 4;
            if (t.shadow_kind == 2) {      // 3x3 conv weight: scatter into the forward and data-gradient images
                const int cin = t.shadow_k, cout = t.shadow_c;
                __nv_bfloat16* fwd = t.shadow;
                __nv_bfloat16* dgr = t.shadow + static_cast<size_t>(cin) * cout * 9;
                const int i0 = static_cast<int>(dst);
                conv_image_store(__float2bfloat16(p.x), i0, cin, cout, t.shadow_ss, fwd, dgr);
                conv_image_store(__float2bfloat16(p.y), i0 + 1, cin, cout, t.shadow_ss, fwd, dgr);
                conv_image_store(__float2bfloat16(p.z), i0 + 2, cin, cout, t.shadow_ss, fwd, dgr);
                conv_image_store(__float2bfloat16(p.w), i0 + 3, cin, cout, t.shadow_ss, fwd, dgr);
                continue;
            }
            if (t.shadow_kind == 1) {      // element (f, k) of [C*SS][K] -> [SS][K/8][C][8]; the 4 k's stay contiguous
                const size_t f = dst / t.shadow_k;
                const int k = static_cast<int>(dst - f * t.shadow_k);
                const int c = static_cast<int>(f / t.shadow_ss), px = static_cast<int>(f - static_cast<size_t>(c) * t.shadow_ss);
                dst = ((static_cast<size_t>(px) * (t.shadow_k / 8) + (k >> 3)) * t.shadow_c + c) * 8 + (k & 7);
            }
            *reinterpret_cast<uint2*>(t.shadow + dst) = o;
        }
    }
    // tail (n % 4 elements), handled by the first threads of block 0
    if (blockIdx.x == 0) {
        const long long i = (n4 << 2) + threadIdx.x;
        if (i < t.n) {
            const float g = t.g[i];
            float m = t.m[i], v = t.v[i], p = t.p[i];
            adam1(p, g, m, v, beta1, beta2, eps, step_size, inv_bc2_sqrt);
            t.p[i] = p; t.m[i] = m; t.v[i] = v;
            if (t.shadow && !t.shadow_kind) t.shadow[i] = __float2bfloat16(p);   // (image layouts have n % 4 == 0)
        }
    }
#undef NGAN_ADAM1
}

// ---------------------------------------------------------------------------------------------------------------
// Adam on the generator's Linear weight with the gradient taken from its FACTORS (reference models.py:240-241 backward +
// train.py:385):  g[f][k] = gscale * sum_b ga[b][f] * z[b][k]  over the Btot samples of the (global) batch, where ga is
// the gradient at the stem's pre-activation (C8 bf16 [b][C/8][SS][8], f = c*SS + px) and z the latent batch.
// The 16.8 M-element gradient (98 % of the generator) is a rank-Btot product: it is never written to HBM nor read back
// (-134 MB per step), and with data parallelism the ranks exchange the factors (1 MB + 32 KB per rank, all-gather)
// instead of all-reducing the 67 MB product.  One thread owns 8 channels (one C8 granule) x 8 k of one pixel: per
// sample one broadcast 16-byte load of ga and 32 bytes of z feed 64 FMAs, so the kernel stays bound by the 28 B/element
// of Adam traffic up to a global batch of ~128.  Samples are added in index order: deterministic, and bit-identical
// to ngan_linear_wgrad followed by ngan_adam_multi.
// Samples may live in per-rank segments of an all-gathered buffer: sample b is row (b % b_per_seg) of segment
// b / b_per_seg, segments ga_seg_stride / z_seg_stride BYTES apart.
struct AdamLinearArgs {
    float* p;
    float* m;
    float* v;
    __nv_bfloat16* shadow;     // operand image [SS][K/8][C][8] or null
    float* g_out;              // optional: materialise the gradient here ([C*SS][K] fp32)
    const uint4* ga;
    const float* z;
    int Btot, b_per_seg;
    long long ga_seg_stride, z_seg_stride;
    int K, C, SS;
    float gscale, step_size, inv_bc2_sqrt;
    const float* dyn;
    float beta1, beta2, eps;
};
__global__ void __launch_bounds__(256) adam_linear_factored_kernel(const AdamLinearArgs a) {
    const float step_size = a.dyn ? __ldg(a.dyn) : a.step_size;
    const float inv_bc2_sqrt = a.dyn ? __ldg(a.dyn + 1) : a.inv_bc2_sqrt;
    const int KG = a.K >> 3, NCH = a.C >> 3;
    const long long items = static_cast<long long>(NCH) * a.SS * KG;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long it = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; it < items; it += stride) {
        const int kg = static_cast<int>(it % KG);
        const long long rest = it / KG;
        const int px = static_cast<int>(rest % a.SS), j = static_cast<int>(rest / a.SS);
        float acc[8][8];                       // [channel of the granule][k]
#pragma unroll
        for (int e = 0; e < 8; ++e)
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[e][q] = 0.f;
        for (int b = 0; b < a.Btot; ++b) {
            const int seg = b / a.b_per_seg, row = b - seg * a.b_per_seg;
            const uint4* gp = reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(a.ga) + seg * a.ga_seg_stride) +
                              (static_cast<size_t>(row) * NCH + j) * a.SS + px;
            const float4* zp = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(a.z) + seg * a.z_seg_stride +
                                                               static_cast<size_t>(row) * a.K * sizeof(float)) + kg * 2;
            float gv[8];
            unpack8(__ldg(gp), gv);
            const float4 z0 = __ldg(zp), z1 = __ldg(zp + 1);
            const float zv[8] = {z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e)
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[e][q] += gv[e] * zv[q];
        }
        uint4 img[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const long long i = ((static_cast<long long>(j * 8 + e) * a.SS + px) * a.K + kg * 8) >> 2;   // float4 index
            uint32_t packed[4];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float4 p = reinterpret_cast<float4*>(a.p)[i + h];
                float4 m = reinterpret_cast<float4*>(a.m)[i + h];
                float4 v = reinterpret_cast<float4*>(a.v)[i + h];
                // __fmul_rn: the gradient is rounded to fp32 exactly as if it had been stored (no FMA contraction with
                // the moment updates below), so this pass is bit-identical to ngan_linear_wgrad + ngan_adam_multi
                const float4 g = make_float4(__fmul_rn(a.gscale, acc[e][4 * h]), __fmul_rn(a.gscale, acc[e][4 * h + 1]),
                                             __fmul_rn(a.gscale, acc[e][4 * h + 2]), __fmul_rn(a.gscale, acc[e][4 * h + 3]));
#define NGAN_ADAM1(c) adam1(p.c, g.c, m.c, v.c, a.beta1, a.beta2, a.eps, step_size, inv_bc2_sqrt);
                NGAN_ADAM1(x) NGAN_ADAM1(y) NGAN_ADAM1(z) NGAN_ADAM1(w)
#undef NGAN_ADAM1
                reinterpret_cast<float4*>(a.p)[i + h] = p;
                reinterpret_cast<float4*>(a.m)[i + h] = m;
                reinterpret_cast<float4*>(a.v)[i + h] = v;
                if (a.g_out) reinterpret_cast<float4*>(a.g_out)[i + h] = g;
                packed[2 * h] = pack_bf16(p.x, p.y);
                packed[2 * h + 1] = pack_bf16(p.z, p.w);
            }
            img[e] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
        if (a.shadow) {
            uint4* dst = reinterpret_cast<uint4*>(a.shadow) + (static_cast<long long>(px) * KG + kg) * a.C + j * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) dst[e] = img[e];
        }
    }
}
// ---- tensor-core version (C % 128 == 0, K % 128 == 0): the rank-Btot product runs on mma.sync m16n8k16, so the pass
// stays bound by Adam's 28 B / element for ANY global batch (the FMA version above needs 64 FMAs per sample and 8x8
// tile: 119 us at 64 samples, 240 us at 128 -- more than the 73 us of memory traffic).  ga is bf16 already; z is split
// into two bf16 terms z = hi + lo (|error| <= 2^-17 |z|), two MMAs per tile, fp32 accumulation: the gradient agrees
// with the fp32-z product to ~1e-5 relative.
// One block = one pixel px x 128 channels x 128 k.  Shared memory: the ga tile [b][128 ch] and the two z tiles
// [b][128 k], rows padded to 272 B (conflict-free ldmatrix).  A C8 granule row [b][8 ch] is the transpose of the mma
// A fragment and a z row [b][8 k] the transpose of the B fragment: ldmatrix.trans delivers both (as in wgrad.cu).
// Warp w owns channels 16w..16w+15 x all 128 k: 64 fp32 accumulators per thread, then the Adam update straight from
// the accumulator fragments (each quad of lanes covers one full 32-byte sector of p / m / v per row).
constexpr int kLinPitch = 272;     // bytes per shared-memory row: 128 bf16 + 16 B pad
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, "
        "%2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool ADAM>
__global__ void __launch_bounds__(256, 2) adam_linear_mma_kernel(const AdamLinearArgs a, int Bpad) {
    extern __shared__ __align__(16) uint8_t lin_smem[];
    uint8_t* s_ga = lin_smem;                                   // [Bpad][kLinPitch]
    uint8_t* s_zh = s_ga + static_cast<size_t>(Bpad) * kLinPitch;
    uint8_t* s_zl = s_zh + static_cast<size_t>(Bpad) * kLinPitch;
    const int px = blockIdx.x, k0 = blockIdx.y * 128, c0 = blockIdx.z * 128;
    const int NCH = a.C >> 3, KG = a.K >> 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // ---- g tile = ga^T z: M = 16 channels of this warp, N = 128 k (16 n-tiles), K = samples, in chunks of Bpad <= 64
    // samples staged at a time (bounded shared memory for any global batch)
    float acc[16][4];
#pragma unroll
    for (int nt = 0; nt < 16; ++nt)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[nt][q] = 0.f;
    const int mi = lane >> 3, ri = lane & 7;
    const uint32_t a_lane = smem_u32(s_ga) + ((mi >> 1) * 8 + ri) * kLinPitch + (warp * 2 + (mi & 1)) * 16;
    const uint32_t bh_lane = smem_u32(s_zh) + ((mi & 1) * 8 + ri) * kLinPitch + (mi >> 1) * 16;
    const uint32_t bl_lane = smem_u32(s_zl) + ((mi & 1) * 8 + ri) * kLinPitch + (mi >> 1) * 16;
    for (int base = 0; base < a.Btot; base += Bpad) {
        if (base) __syncthreads();                              // everyone is done reading the previous chunk
        // stage the factors: 16 granules of ga and 128 latents (as hi / lo bf16) per sample; rows past Btot are zero
        for (int i = threadIdx.x; i < Bpad * 16; i += blockDim.x) {
            const int r = i >> 4, jj = i & 15, b = base + r;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (b < a.Btot) {
                const int seg = b / a.b_per_seg, row = b - seg * a.b_per_seg;
                v = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const char*>(a.ga) + seg * a.ga_seg_stride) +
                          (static_cast<size_t>(row) * NCH + (c0 >> 3) + jj) * a.SS + px);
            }
            *reinterpret_cast<uint4*>(s_ga + r * kLinPitch + jj * 16) = v;
        }
        for (int i = threadIdx.x; i < Bpad * 32; i += blockDim.x) {
            const int r = i >> 5, q = i & 31, b = base + r;     // q: float4 index inside the 128-k slice
            float4 zv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b < a.Btot) {
                const int seg = b / a.b_per_seg, row = b - seg * a.b_per_seg;
                zv = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const char*>(a.z) + seg * a.z_seg_stride +
                                                           static_cast<size_t>(row) * a.K * sizeof(float)) + (k0 >> 2) + q);
            }
            const float f[4] = {zv.x, zv.y, zv.z, zv.w};
            uint32_t hi[2], lo[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const __nv_bfloat16 h0 = __float2bfloat16(f[2 * h]), h1 = __float2bfloat16(f[2 * h + 1]);
                hi[h] = pack_bf16(__bfloat162float(h0), __bfloat162float(h1));
                lo[h] = pack_bf16(f[2 * h] - __bfloat162float(h0), f[2 * h + 1] - __bfloat162float(h1));
            }
            *reinterpret_cast<uint2*>(s_zh + r * kLinPitch + q * 8) = make_uint2(hi[0], hi[1]);
            *reinterpret_cast<uint2*>(s_zl + r * kLinPitch + q * 8) = make_uint2(lo[0], lo[1]);
        }
        __syncthreads();
        for (int b0 = 0; b0 < Bpad; b0 += 16) {
            uint32_t af[4];
            ldsm_x4_trans(a_lane + b0 * kLinPitch, af);
#pragma unroll
            for (int np = 0; np < 8; ++np) {                    // pairs of n-tiles
                uint32_t bh[4], bl[4];
                ldsm_x4_trans(bh_lane + b0 * kLinPitch + np * 32, bh);
                ldsm_x4_trans(bl_lane + b0 * kLinPitch + np * 32, bl);
                mma16816(acc[2 * np], af, bh[0], bh[1]);
                mma16816(acc[2 * np + 1], af, bh[2], bh[3]);
                mma16816(acc[2 * np], af, bl[0], bl[1]);
                mma16816(acc[2 * np + 1], af, bl[2], bl[3]);
            }
        }
    }
    // ---- Adam straight from the accumulator fragments: c[0..1] = (channel g, k 2t..2t+1), c[2..3] = (channel g + 8, same k)
    const float step_size = a.dyn ? __ldg(a.dyn) : a.step_size;
    const float inv_bc2_sqrt = a.dyn ? __ldg(a.dyn + 1) : a.inv_bc2_sqrt;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = c0 + warp * 16 + g + 8 * h;
        const size_t row = (static_cast<size_t>(c) * a.SS + px) * a.K;
        // four n-tiles at a time: 12 independent 8-byte loads in flight per thread before the first use (with one
        // n-tile per iteration the pass ran at the latency of its loads, not at HBM bandwidth)
        if constexpr (!ADAM) {                      // gradient only: dW = gscale * ga^T z, written once
#pragma unroll
            for (int nt = 0; nt < 16; ++nt) {
                const size_t i2 = (row + k0 + nt * 8 + 2 * t) >> 1;
                reinterpret_cast<float2*>(a.g_out)[i2] =
                    make_float2(__fmul_rn(a.gscale, acc[nt][2 * h]), __fmul_rn(a.gscale, acc[nt][2 * h + 1]));
            }
            continue;
        }
#pragma unroll
        for (int n0 = 0; n0 < 16; n0 += 4) {
            float2 p[4], m[4], v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const size_t i2 = (row + k0 + (n0 + q) * 8 + 2 * t) >> 1;       // float2 index
                p[q] = reinterpret_cast<const float2*>(a.p)[i2];
                m[q] = reinterpret_cast<const float2*>(a.m)[i2];
                v[q] = reinterpret_cast<const float2*>(a.v)[i2];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int nt = n0 + q, k = k0 + nt * 8 + 2 * t;
                const size_t i2 = (row + k) >> 1;
                const float gx = __fmul_rn(a.gscale, acc[nt][2 * h]), gy = __fmul_rn(a.gscale, acc[nt][2 * h + 1]);
                adam1(p[q].x, gx, m[q].x, v[q].x, a.beta1, a.beta2, a.eps, step_size, inv_bc2_sqrt);
                adam1(p[q].y, gy, m[q].y, v[q].y, a.beta1, a.beta2, a.eps, step_size, inv_bc2_sqrt);
                reinterpret_cast<float2*>(a.p)[i2] = p[q];
                reinterpret_cast<float2*>(a.m)[i2] = m[q];
                reinterpret_cast<float2*>(a.v)[i2] = v[q];
                if (a.g_out) reinterpret_cast<float2*>(a.g_out)[i2] = make_float2(gx, gy);
                if (a.shadow) {
                    uint32_t* dst = reinterpret_cast<uint32_t*>(a.shadow) +
                                    (((static_cast<size_t>(px) * KG + (k >> 3)) * a.C + c) * 8 + (k & 7)) / 2;
                    *dst = pack_bf16(p[q].x, p[q].y);
                }
            }
        }
    }
}

int adam_linear_factored(float* p, float* m, float* v, void* shadow, float* g_out, const void* ga, const float* z,
                         int Btot, int b_per_seg, long long ga_seg_stride, long long z_seg_stride, int K, int C, int SS,
                         float gscale, float step_size, float inv_bc2_sqrt, const float* dyn, float beta1, float beta2,
                         float eps, cudaStream_t st) {
    if (K % 8 || C % 8 || Btot <= 0 || b_per_seg <= 0) {
        set_error("adam_linear_factored: bad shape K=%d C=%d Btot=%d", K, C, Btot);
        return NGAN_ERR_INVALID;
    }
    AdamLinearArgs a{p, m, v, static_cast<__nv_bfloat16*>(shadow), g_out, static_cast<const uint4*>(ga), z, Btot,
                     b_per_seg, ga_seg_stride, z_seg_stride, K, C, SS, gscale, step_size, inv_bc2_sqrt, dyn, beta1,
                     beta2, eps};
    static const bool force_fma = getenv("NGAN_LINEAR_ADAM_FMA") != nullptr;
    int Bpad = (Btot + 15) / 16 * 16;          // samples staged per pass: the whole batch up to 64, else chunks of 64
    if (Bpad > 64) Bpad = 64;
    const size_t smem = static_cast<size_t>(3) * Bpad * kLinPitch;
    if (!force_fma && C % 128 == 0 && K % 128 == 0) {
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(adam_linear_mma_kernel<true>,
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(adam_linear_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
            if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(adam_linear_mma)");
            configured = true;
        }
        if (p)
            adam_linear_mma_kernel<true><<<dim3(SS, K / 128, C / 128), 256, smem, st>>>(a, Bpad);
        else
            adam_linear_mma_kernel<false><<<dim3(SS, K / 128, C / 128), 256, smem, st>>>(a, Bpad);
        return check_launch("adam_linear_mma");
    }
    if (!p) {
        set_error("adam_linear_factored: the gradient-only mode needs C %% 128 == 0, K %% 128 == 0 (got C=%d K=%d)", C, K);
        return NGAN_ERR_UNSUPPORTED;
    }
    const long long items = static_cast<long long>(C / 8) * SS * (K / 8);
    long long bx = (items + 255) / 256;
    if (bx > 148 * 8) bx = 148 * 8;
    adam_linear_factored_kernel<<<static_cast<unsigned>(bx), 256, 0, st>>>(a);
    return check_launch("adam_linear_factored");
}

int adam_multi_launch(const AdamEntry* entries, int n_tensors, float beta1, float beta2, float eps,
                      cudaStream_t st) {
    if (n_tensors <= 0) return NGAN_OK;
    if (n_tensors > kAdamMaxTensors) {
        set_error("adam_multi: %d tensors exceed the table size %d", n_tensors, kAdamMaxTensors);
        return NGAN_ERR_INVALID;
    }
    AdamTable tab;
    long long max_n = 0;
    for (int i = 0; i < n_tensors; ++i) {
        tab.e[i] = entries[i];
        if (entries[i].n > max_n) max_n = entries[i].n;
        if ((reinterpret_cast<uintptr_t>(entries[i].p) | reinterpret_cast<uintptr_t>(entries[i].g) |
             reinterpret_cast<uintptr_t>(entries[i].m) | reinterpret_cast<uintptr_t>(entries[i].v)) & 15) {
            set_error("adam_multi: tensor %d is not 16-byte aligned", i);
            return NGAN_ERR_INVALID;
        }
    }
    for (int i = n_tensors; i < kAdamMaxTensors; ++i) tab.e[i] = AdamEntry{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0.f, 0.f, 0, 0, 0, 0, nullptr};
    long long bx = (max_n / 4 + 255) / 256;
    if (bx > 148 * 8) bx = 148 * 8;
    if (bx < 1) bx = 1;
    adam_multi_kernel<<<dim3(static_cast<unsigned>(bx), n_tensors), 256, 0, st>>>(tab, beta1, beta2, eps);
    return check_launch("adam_multi");
}

}  // namespace ngan
