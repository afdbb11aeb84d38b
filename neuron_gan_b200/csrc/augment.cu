// On-device image pipeline: what the reference's DatasetIterator.__next__ (data/NeuronDataset.py:170-205) does per
// image with the transform list of NeuronDataset.__init__ / set_image_size (data/NeuronDataset.py:112-126, 149-164)
//
//   RandomAffine (nearest, zero fill) -> RandomVerticalFlip -> ColorJitter (brightness, contrast, random order)
//   -> CenterCrop -> Renormalize to [-1, 1] -> Resize(antialias=True)
//
// as two kernels over a whole batch instead of ~25 small torch kernels per IMAGE.  The preloaded padded canvases
// stay resident in HBM ([N][P][P] fp32); the random parameters are drawn on the host with torch's CPU generator in
// the reference's order (neuron_gan_b200/data.py) and arrive as a small table.
//
//   augment_mean_kernel    per-image mean of the affine-resampled (and, if brightness comes first, brightened)
//                          canvas -- the operand of torchvision's adjust_contrast.  Deterministic: per-CTA partial
//                          sums, combined in a fixed order by the consumer.
//   augment_resize_kernel  one group of G threads per OUTPUT pixel gathers its antialias footprint straight from the
//                          canvas: inverse-affine nearest lookup, flip, jitter, renormalise, triangle-filter weights.
//                          Nothing intermediate is written; the output goes wherever the caller points (e.g. the
//                          training step's real-image slot).
//
// HBM-bound and tiny next to the training step: reads B*P*P*4 bytes (twice, the second time from L2), writes
// B*R*R*4.  The source coordinates use explicitly rounded fp32 multiplies/adds (no FMA contraction) so that the
// nearest-neighbour selection is the one the fp32 expression of torchvision's _gen_affine_grid + grid_sample makes.
#include "common.cuh"
#include "kernels.h"

namespace ngan {

constexpr int kAugParams = 16;      // floats per image in the parameter table
constexpr int kAugThreads = 256;

struct AugImage {
    float t00, t01, t02, t10, t11, t12;   // inverse affine matrix / (P/2), row-major 2x3
    float b, c, omc;                      // brightness factor, contrast factor, fp32(1 - contrast factor)
    int flip, order, identity;
};

__device__ __forceinline__ AugImage load_aug(const float* __restrict__ p) {
    AugImage a;
    a.t00 = p[0]; a.t01 = p[1]; a.t02 = p[2]; a.t10 = p[3]; a.t11 = p[4]; a.t12 = p[5];
    a.flip = p[6] != 0.f;
    a.b = p[7]; a.c = p[8]; a.omc = p[9];
    a.order = p[10] != 0.f;
    a.identity = p[11] != 0.f;
    return a;
}

// value of the affine-resampled canvas at (y, x): grid_sample(mode='nearest', padding_mode='zeros',
// align_corners=False) on the grid of _gen_affine_grid
__device__ __forceinline__ float affine_value(const float* __restrict__ img, int P, const AugImage& a, int y, int x) {
    const float half = 0.5f * P, Pf = static_cast<float>(P);
    const float xs = static_cast<float>(x) - half + 0.5f, ys = static_cast<float>(y) - half + 0.5f;
    const float gx = __fadd_rn(__fadd_rn(__fmul_rn(xs, a.t00), __fmul_rn(ys, a.t01)), a.t02);
    const float gy = __fadd_rn(__fadd_rn(__fmul_rn(xs, a.t10), __fmul_rn(ys, a.t11)), a.t12);
    const float ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), Pf), -1.f), 0.5f);
    const float iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), Pf), -1.f), 0.5f);
    const int jx = __float2int_rn(ix), jy = __float2int_rn(iy);      // nearbyint: half to even
    if (jx < 0 || jx >= P || jy < 0 || jy >= P) return 0.f;
    return __ldg(img + static_cast<size_t>(jy) * P + jx);
}
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
// torchvision _blend(img, mean, c) = clamp(c*img + (1-c)*mean), products rounded separately
__device__ __forceinline__ float contrast(float v, const AugImage& a, float mean) {
    return clamp01(__fadd_rn(__fmul_rn(a.c, v), __fmul_rn(a.omc, mean)));
}

__global__ void __launch_bounds__(kAugThreads)
augment_mean_kernel(const float* __restrict__ canvases, const int* __restrict__ src_index,
                    const float* __restrict__ params, float* __restrict__ partials, int P) {
    const int b = blockIdx.y;
    const AugImage a = load_aug(params + static_cast<size_t>(b) * kAugParams);
    __shared__ float warp_part[kAugThreads / 32];
    float s = 0.f;
    if (!a.identity) {
        const float* img = canvases + static_cast<size_t>(src_index[b]) * P * P;
        const int n = P * P;
        for (int p = blockIdx.x * kAugThreads + threadIdx.x; p < n; p += gridDim.x * kAugThreads) {
            const int y = p / P, x = p - y * P;
            float v = affine_value(img, P, a, y, x);
            if (a.order == 0) v = clamp01(__fmul_rn(a.b, v));
            s += v;
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kAugThreads / 32; ++w) t += warp_part[w];
        partials[static_cast<size_t>(b) * gridDim.x + blockIdx.x] = t;
    }
}

template <int G>
__global__ void __launch_bounds__(kAugThreads)
augment_resize_kernel(const float* __restrict__ canvases, const int* __restrict__ src_index,
                      const float* __restrict__ params, const float* __restrict__ partials, int n_partials,
                      const int* __restrict__ tap_first, const int* __restrict__ tap_count,
                      const float* __restrict__ tap_weight, int max_taps, float* __restrict__ out, int P, int top,
                      int R) {
    const int b = blockIdx.y;
    const AugImage a = load_aug(params + static_cast<size_t>(b) * kAugParams);
    __shared__ float mean_s;
    if (!a.identity) {
        if (threadIdx.x < 32) {   // fixed-order combine of the per-CTA partial sums
            float t = 0.f;
            for (int i = threadIdx.x; i < n_partials; i += 32) t += partials[static_cast<size_t>(b) * n_partials + i];
            t = warp_sum(t);
            if (threadIdx.x == 0) mean_s = __fdiv_rn(t, static_cast<float>(P) * static_cast<float>(P));
        }
        __syncthreads();
    }
    const float mean = a.identity ? 0.f : mean_s;
    const float* img = canvases + static_cast<size_t>(src_index[b]) * P * P;
    const int sub = threadIdx.x % G;
    const int o = blockIdx.x * (kAugThreads / G) + threadIdx.x / G;
    const bool live = o < R * R;          // whole groups are live or dead together (G divides the block)
    float acc = 0.f;
    int oy = 0, ox = 0;
    if (live) {
        oy = o / R;
        ox = o - oy * R;
        const int y0 = tap_first[oy], ny = tap_count[oy], x0 = tap_first[ox], nx = tap_count[ox];
        const float* wy = tap_weight + static_cast<size_t>(oy) * max_taps;
        const float* wx = tap_weight + static_cast<size_t>(ox) * max_taps;
        for (int t = sub; t < ny * nx; t += G) {
            const int ty = t / nx, tx = t - ty * nx;
            int y = top + y0 + ty;
            const int x = top + x0 + tx;
            float v;
            if (a.identity) {
                v = __ldg(img + static_cast<size_t>(y) * P + x);
            } else {
                if (a.flip) y = P - 1 - y;
                v = affine_value(img, P, a, y, x);
                if (a.order == 0) {
                    v = contrast(clamp01(__fmul_rn(a.b, v)), a, mean);
                } else {
                    v = clamp01(__fmul_rn(a.b, contrast(v, a, mean)));
                }
            }
            v = __fadd_rn(__fmul_rn(v, 2.f), -1.f);                       // Renormalize((-1, 1), (0, 1))
            acc = fmaf(__fmul_rn(wy[ty], wx[tx]), v, acc);
        }
    }
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if (live && sub == 0) out[(static_cast<size_t>(b) * R + oy) * R + ox] = acc;
}

static int mean_chunks(int canvas) {
    int n = (canvas * canvas + kAugThreads * 16 - 1) / (kAugThreads * 16);
    return n < 1 ? 1 : (n > 256 ? 256 : n);
}
size_t augment_workspace_bytes(int batch, int canvas) {
    return static_cast<size_t>(batch) * mean_chunks(canvas) * sizeof(float);
}

int augment_batch(const float* canvases, const int* src_index, const float* params, const int* tap_first,
                  const int* tap_count, const float* tap_weight, int max_taps, float* workspace, float* out,
                  int batch, int canvas, int crop, int out_size, cudaStream_t st) {
    const int chunks = mean_chunks(canvas);
    const int top = static_cast<int>(rint((canvas - crop) / 2.0));          // torchvision center_crop
    augment_mean_kernel<<<dim3(chunks, batch), kAugThreads, 0, st>>>(canvases, src_index, params, workspace, canvas);
    if (int e = check_launch("augment_mean")) return e;
    const int taps = max_taps * max_taps;
    const int px = out_size * out_size;
#define NGAN_AUG_LAUNCH(G)                                                                                        \
    augment_resize_kernel<G><<<dim3((px + kAugThreads / G - 1) / (kAugThreads / G), batch), kAugThreads, 0, st>>>( \
        canvases, src_index, params, workspace, chunks, tap_first, tap_count, tap_weight, max_taps, out, canvas,  \
        top, out_size)
    if (taps <= 2) NGAN_AUG_LAUNCH(1);
    else if (taps <= 32) NGAN_AUG_LAUNCH(4);
    else NGAN_AUG_LAUNCH(32);
#undef NGAN_AUG_LAUNCH
    return check_launch("augment_resize");
}

}  // namespace ngan
