// On-device image pipeline: what the reference's DatasetIterator.__next__ (data/NeuronDataset.py:170-205) does per
// image with the transform list of NeuronDataset.__init__ / set_image_size (data/NeuronDataset.py:112-126, 149-164)
//
//   RandomAffine (nearest, zero fill) -> RandomVerticalFlip -> ColorJitter (brightness, contrast, random order)
//   -> CenterCrop -> Renormalize to [-1, 1] -> Resize(antialias=True)
//
// as three launches over a whole batch instead of ~25 small torch kernels per IMAGE.  The preloaded padded canvases
// stay resident in HBM ([N][P][P] fp32); the random parameters are drawn on the host with torch's CPU generator in
// the reference's order (neuron_gan_b200/data.py) and arrive as a small table.
//
//   augment_affine_kernel  the only pass that evaluates the inverse affine map, once per canvas pixel: nearest
//                          lookup (+ brightness when it precedes contrast), per-CTA partial sums for the image mean
//                          that torchvision's adjust_contrast needs, and the pixels inside the crop window stored to
//                          a staging image (B x crop x crop fp32; 17 MB at the BASELINE shape, L2-resident).
//   augment_mean_finish    adds the partial sums in a fixed order (deterministic), one warp per image.
//   augment_resize_kernel  G threads per OUTPUT pixel walk its antialias footprint in the staging image: flip,
//                          rest of the jitter, renormalise, triangle-filter weights.  No affine math here: the
//                          footprints of neighbouring outputs overlap 4x, the expensive part must not be redone.
//
// Algorithmic bytes per batch: B*P*P*4 read + B*R*R*4 written.  The kernels are bound by instruction issue (about
// 40 instructions per canvas pixel around a gathered load), not by HBM; see profiles/README.md.
// The source coordinates use explicitly rounded fp32 multiplies/adds (no FMA contraction) so that the
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
    const float4 q0 = __ldg(reinterpret_cast<const float4*>(p));
    const float4 q1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
    const float4 q2 = __ldg(reinterpret_cast<const float4*>(p) + 2);
    AugImage a;
    a.t00 = q0.x; a.t01 = q0.y; a.t02 = q0.z; a.t10 = q0.w; a.t11 = q1.x; a.t12 = q1.y;
    a.flip = q1.z != 0.f;
    a.b = q1.w; a.c = q2.x; a.omc = q2.y;
    a.order = q2.z != 0.f;
    a.identity = q2.w != 0.f;
    return a;
}

// value of the affine-resampled canvas at (y, x): grid_sample(mode='nearest', padding_mode='zeros',
// align_corners=False) on the grid of _gen_affine_grid.  xs/ys are the centred coordinates x - P/2 + 0.5.
__device__ __forceinline__ float affine_value(const float* __restrict__ img, int P, float Pf, const AugImage& a,
                                              float xs, float ys) {
    const float gx = __fadd_rn(__fadd_rn(__fmul_rn(xs, a.t00), __fmul_rn(ys, a.t01)), a.t02);
    const float gy = __fadd_rn(__fadd_rn(__fmul_rn(xs, a.t10), __fmul_rn(ys, a.t11)), a.t12);
    const float ix = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gx, 1.f), Pf), -1.f), 0.5f);
    const float iy = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(gy, 1.f), Pf), -1.f), 0.5f);
    const int jx = __float2int_rn(ix), jy = __float2int_rn(iy);      // nearbyint: half to even
    if (static_cast<unsigned>(jx) >= static_cast<unsigned>(P) || static_cast<unsigned>(jy) >= static_cast<unsigned>(P))
        return 0.f;
    return __ldg(img + static_cast<unsigned>(jy * P + jx));
}
__device__ __forceinline__ float clamp01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }
// torchvision _blend(img, mean, c) = clamp(c*img + (1-c)*mean), products rounded separately
__device__ __forceinline__ float contrast(float v, const AugImage& a, float mean) {
    return clamp01(__fadd_rn(__fmul_rn(a.c, v), __fmul_rn(a.omc, mean)));
}

// Staging window in canvas coordinates: columns [top, top + crop), rows [row0, row0 + rows) -- the crop rows and
// their mirror image under the vertical flip (the same rows unless canvas - crop is odd).
struct AugWindow {
    int top, row0, rows, crop;
};

// Every warp works on an 8-wide x 4-tall patch of pixels, not a 32-pixel row: under a rotation the 32 source
// addresses of a row land in up to 32 different 128-byte lines (one L1 wavefront each), those of a patch in ~9.
// No integer division anywhere: coordinates come from the 3-D grid.
constexpr int kAffineRows = 128;      // canvas rows per CTA (CTA tile: 32 columns x kAffineRows)

__global__ void __launch_bounds__(kAugThreads)
augment_affine_kernel(const float* __restrict__ canvases, const int* __restrict__ src_index,
                      const float* __restrict__ params, float* __restrict__ partials, float* __restrict__ staging,
                      int P, AugWindow win) {
    const int b = blockIdx.z;
    const AugImage a = load_aug(params + static_cast<size_t>(b) * kAugParams);
    __shared__ float warp_part[kAugThreads / 32];
    float s = 0.f;
    if (!a.identity) {
        const float* img = canvases + static_cast<size_t>(__ldg(src_index + b)) * P * P;
        float* stage = staging + static_cast<size_t>(b) * win.rows * win.crop;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        // 8 warps: 4 across x 2 down -> a 32 x 8 pixel slab per step
        const int x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
        const int y0 = blockIdx.y * kAffineRows + (warp >> 2) * 4 + (lane >> 3);
        const int y1 = min(P, (blockIdx.y + 1) * kAffineRows);
        const float half = 0.5f * P, Pf = static_cast<float>(P);
        if (x < P) {
            const float xs = static_cast<float>(x) - half + 0.5f;
            const bool x_in = static_cast<unsigned>(x - win.top) < static_cast<unsigned>(win.crop);
            float ys = static_cast<float>(y0) - half + 0.5f;      // exact, and so is every ys + 8
            // where (y, x) lives in the staging window; only dereferenced inside it
            float* sp = stage + static_cast<ptrdiff_t>(y0 - win.row0) * win.crop + (x - win.top);
            const int sp_step = 8 * win.crop;
#pragma unroll 4
            for (int y = y0; y < y1; y += 8, ys += 8.f, sp += sp_step) {
                float v = affine_value(img, P, Pf, a, xs, ys);
                if (a.order == 0) v = clamp01(__fmul_rn(a.b, v));
                s += v;
                if (x_in && static_cast<unsigned>(y - win.row0) < static_cast<unsigned>(win.rows)) *sp = v;
            }
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kAugThreads / 32; ++w) t += warp_part[w];
        partials[(static_cast<size_t>(b) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = t;
    }
}

// fixed-order combine of the per-CTA partial sums: mean[b], one warp per image
__global__ void augment_mean_finish_kernel(const float* __restrict__ partials, int n_partials, float* __restrict__ means,
                                           float n_pixels) {
    const int b = blockIdx.x;
    float t = 0.f;
    for (int i = threadIdx.x; i < n_partials; i += 32) t += partials[static_cast<size_t>(b) * n_partials + i];
    t = warp_sum(t);
    if (threadIdx.x == 0) means[b] = __fdiv_rn(t, n_pixels);
}

// G threads per OUTPUT pixel, every group walks `rows` outputs of one column so that the per-thread set-up (parameter
// row, tap table of its column, source window) is paid once per several outputs:
//   G = 1   up to 4x4 taps (no resize: exactly one).  CTA = 32 columns x 8 groups down.
//   G = 8   up to 16x16 taps; the 8 lanes of a group read 8 consecutive pixels of a footprint row (32 bytes).
//           CTA = 8 columns x 4 groups down.
//   G = 32  larger footprints, swept in 8-wide x 4-tall patches.  CTA = 8 columns x 1.
template <int G>
__global__ void __launch_bounds__(kAugThreads)
augment_resize_kernel(const float* __restrict__ canvases, const int* __restrict__ src_index,
                      const float* __restrict__ params, const float* __restrict__ means,
                      const float* __restrict__ staging, const int* __restrict__ tap_first,
                      const int* __restrict__ tap_count, const float* __restrict__ tap_weight, int max_taps,
                      float* __restrict__ out, int P, AugWindow win, int R, int rows) {
    constexpr int kCols = G == 1 ? 32 : 8;                 // output columns per CTA
    constexpr int kDown = kAugThreads / G / kCols;         // groups stacked in y
    const int b = blockIdx.z;
    const AugImage a = load_aug(params + static_cast<size_t>(b) * kAugParams);
    const float mean = a.identity ? 0.f : __ldg(means + b);
    // identity: straight from the canvas; otherwise from the staging window (already resampled, maybe brightened)
    const float* src;
    int pitch, row_first, row_step;
    if (a.identity) {
        src = canvases + static_cast<size_t>(__ldg(src_index + b)) * P * P + win.top;
        pitch = P; row_first = win.top; row_step = 1;
    } else {
        src = staging + static_cast<size_t>(b) * win.rows * win.crop;
        pitch = win.crop;
        // crop row k is canvas row top + k, or its mirror P - 1 - (top + k) under the flip
        row_first = (a.flip ? P - 1 - win.top : win.top) - win.row0;
        row_step = a.flip ? -1 : 1;
    }
    const int group = threadIdx.x / G, sub = threadIdx.x % G;
    const int ox = blockIdx.x * kCols + group % kCols;
    const int gy = group / kCols;
    const int sub_x = G == 1 ? 0 : sub & 7, step_x = G == 1 ? 1 : 8;
    const int sub_y = G == 32 ? sub >> 3 : 0, step_y = G == 32 ? 4 : 1;
    const bool col_live = ox < R;
    int x0 = 0, nx = 0;
    const float* wx = tap_weight;
    if (col_live) {
        x0 = __ldg(tap_first + ox);
        nx = __ldg(tap_count + ox);
        wx = tap_weight + ox * max_taps;
    }
    src += x0;
    float* out_col = out + static_cast<size_t>(b) * R * R + ox;
    for (int r = 0; r < rows; ++r) {                        // uniform trip count: the shuffles below stay converged
        const int oy = (blockIdx.y * rows + r) * kDown + gy;
        const bool live = col_live && oy < R;               // whole groups are live or dead together
        float acc = 0.f;
        if (live) {
            const int y0 = __ldg(tap_first + oy), ny = __ldg(tap_count + oy);
            const float* wy = tap_weight + oy * max_taps;
            for (int ty = sub_y; ty < ny; ty += step_y) {
                const float wyv = __ldg(wy + ty);
                const float* row = src + (row_first + row_step * (y0 + ty)) * pitch;
                for (int tx = sub_x; tx < nx; tx += step_x) {
                    float v = __ldg(row + tx);
                    if (!a.identity) {
                        v = contrast(v, a, mean);                             // brightness already applied if first
                        if (a.order != 0) v = clamp01(__fmul_rn(a.b, v));
                    }
                    v = __fadd_rn(__fmul_rn(v, 2.f), -1.f);                   // Renormalize((-1, 1), (0, 1))
                    acc = fmaf(__fmul_rn(wyv, __ldg(wx + tx)), v, acc);
                }
            }
        }
#pragma unroll
        for (int d = G / 2; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
        if (live && sub == 0) out_col[oy * R] = acc;
    }
}

static dim3 affine_grid(int canvas, int batch) {
    return dim3((canvas + 31) / 32, (canvas + kAffineRows - 1) / kAffineRows, batch);
}
static AugWindow make_window(int canvas, int crop) {
    AugWindow w;
    w.crop = crop;
    w.top = static_cast<int>(rint((canvas - crop) / 2.0));          // torchvision center_crop (Python round())
    const int mirrored = canvas - w.top - crop;
    w.row0 = w.top < mirrored ? w.top : mirrored;
    w.rows = (w.top > mirrored ? w.top : mirrored) + crop - w.row0;
    return w;
}
// workspace layout (floats): [batch][n_partials] partial sums | [batch] means (padded to 4) | staging images
static size_t staging_offset(int batch, int n_partials) {
    return (static_cast<size_t>(batch) * (n_partials + 1) + 3) / 4 * 4;
}
size_t augment_workspace_bytes(int batch, int canvas, int crop) {
    const dim3 g = affine_grid(canvas, batch);
    const AugWindow w = make_window(canvas, crop);
    return (staging_offset(batch, g.x * g.y) + static_cast<size_t>(batch) * w.rows * w.crop) * sizeof(float);
}

int augment_batch(const float* canvases, const int* src_index, const float* params, const int* tap_first,
                  const int* tap_count, const float* tap_weight, int max_taps, float* workspace, float* out,
                  int batch, int canvas, int crop, int out_size, cudaStream_t st) {
    const dim3 ag = affine_grid(canvas, batch);
    const int n_partials = ag.x * ag.y;
    const AugWindow win = make_window(canvas, crop);
    float* means = workspace + static_cast<size_t>(batch) * n_partials;
    float* staging = workspace + staging_offset(batch, n_partials);
    augment_affine_kernel<<<ag, kAugThreads, 0, st>>>(canvases, src_index, params, workspace, staging, canvas, win);
    if (int e = check_launch("augment_affine")) return e;
    augment_mean_finish_kernel<<<batch, 32, 0, st>>>(workspace, n_partials, means,
                                                     static_cast<float>(canvas) * static_cast<float>(canvas));
    if (int e = check_launch("augment_mean_finish")) return e;
    const int taps = max_taps * max_taps;
    const int R = out_size;
    // outputs per group: enough to amortise the set-up, few enough to keep ~8 CTAs per SM in the grid
#define NGAN_AUG_LAUNCH(G, COLS)                                                                                   \
    do {                                                                                                           \
        const int down = kAugThreads / G / COLS;                                                                   \
        int rows = 1;                                                                                              \
        while (rows < 8 && static_cast<long long>((R + COLS - 1) / COLS) * ((R + down * rows * 2 - 1) /            \
                                                                           (down * rows * 2)) * batch >= 148 * 8)  \
            rows *= 2;                                                                                             \
        augment_resize_kernel<G><<<dim3((R + COLS - 1) / COLS, (R + down * rows - 1) / (down * rows), batch),      \
                                   kAugThreads, 0, st>>>(canvases, src_index, params, means, staging, tap_first,   \
                                                         tap_count, tap_weight, max_taps, out, canvas, win, R,     \
                                                         rows);                                                    \
    } while (0)
    if (taps <= 16) NGAN_AUG_LAUNCH(1, 32);
    else if (taps <= 256) NGAN_AUG_LAUNCH(8, 8);
    else NGAN_AUG_LAUNCH(32, 8);
#undef NGAN_AUG_LAUNCH
    return check_launch("augment_resize");
}

}  // namespace ngan
