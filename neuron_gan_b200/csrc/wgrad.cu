// Weight gradient of the equalised-LR 3x3 convolution (what autograd's convolution_backward computes for
// Conv2d_normalized, reference models.py:203-204), also used for the "wgrad of dgrad" term of the
// gradient-penalty double backward:
//     dW[co][ci][ky][kx] += scale * sum_{b,y,x} ga[b,co,y,x] * x[b,ci,y+ky-1,x+kx-1]
//
// GEMM view: M = co, N = ci (per tap), K = pixels.  COUT/CIN are as small as 16 here, below the M=64/128
// granularity of tcgen05.mma, so this reduction runs on the warp-level tensor-core path (mma.sync
// m16n8k16, bf16 -> fp32): every warp owns one 16(co) x 16(ci) x 9(tap) accumulator block in registers for
// the whole kernel.  Operand tiles arrive by TMA (double-buffered, mbarrier-signalled) straight from the
// C8-planar tensors; a C8 tile [pixel][8 channels] is the transpose of what mma wants, which is exactly what
// ldmatrix.trans delivers, and the 9 taps are 9 shifted ldmatrix addresses into the same haloed x tile.
// CTAs are persistent over pixel tiles (grid.x) and split the (co, ci) block space (grid.y); partial sums
// are combined with fp32 atomics into the (pre-zeroed or accumulating) gradient tensor.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace ngan {

int make_c8_tensor_map(CUtensorMap* map, const void* base, int B, int C, int H, int W, int box_w, int box_h,
                       int box_planes);

struct WgradArgs {
    int B, H, W, cin, cout;
    int TW;               // tile width (multiple of 16); tile height is 8
    int tiles_x, tiles_y, n_tiles;
    int ci_g, co_g;       // channels of x / ga handled by one CTA
    int n_ci_groups;
    int n_stage;          // depth of the TMA ring (2..kWgMaxStages)
    int debug;            // timing experiments (NGAN_WGRAD_DEBUG): 1 = skip the flush, 2 = skip the MMAs
    int cluster;          // CTAs per cluster along grid.x (1, 2, 4 or 8): cluster-level reduction before the atomics
    uint32_t x_stage_bytes, g_stage_bytes;
    float scale;
    float* dw;
};

constexpr int kWgTH = 8;
constexpr int kWgMaxStages = 6;

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, "
        "%2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ---- thread-block cluster helpers (distributed shared memory)
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_saddr, uint32_t rank) {
    uint32_t ra;
    float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                            const __grid_constant__ CUtensorMap tmap_g,
                                                            const WgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    // ring of n_stage {x tile, ga tile} slots followed by the mbarriers
    const uint32_t slot_bytes = a.x_stage_bytes + a.g_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.n_stage * slot_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.y;
    const int ci_group = group % a.n_ci_groups, co_group = group / a.n_ci_groups;
    const int n_ci_blk = a.ci_g / 16, n_co_blk = a.co_g / 16;
    const int n_blk = n_ci_blk * n_co_blk;  // <= 8
    const int rsplit = 8 / n_blk;
    const int blk = warp % n_blk, rs = warp / n_blk;
    const int cib = blk % n_ci_blk, cob = blk / n_ci_blk;

    const int Wh = a.TW + 2;
    const uint32_t x_plane = (kWgTH + 2) * Wh * 16, g_plane = kWgTH * a.TW * 16;
    const uint32_t stage_bytes = (a.ci_g / 8) * x_plane + (a.co_g / 8) * g_plane;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_g);
        for (int i = 0; i < a.n_stage; ++i) mbar_init(bars + i, 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int tile, int stage) {
        const int tx = tile % a.tiles_x;
        const int ty = (tile / a.tiles_x) % a.tiles_y;
        const int b = tile / (a.tiles_x * a.tiles_y);
        uint8_t* slot = smem + stage * slot_bytes;
        mbar_arrive_expect_tx(bars + stage, stage_bytes);
        tma_load_4d(slot, &tmap_x, bars + stage, (tx * a.TW - 1) * 2, ty * kWgTH - 1, ci_group * (a.ci_g / 8), b);
        tma_load_4d(slot + a.x_stage_bytes, &tmap_g, bars + stage, tx * a.TW * 2, ty * kWgTH,
                    co_group * (a.co_g / 8), b);
    };

    float acc[9][2][4];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int n = 0; n < 2; ++n)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[t][n][k] = 0.f;

    // Prologue: n_stage - 1 tiles in flight.  Iteration `it` first refills the slot that iteration it-1 consumed
    // (everyone left it at the __syncthreads that closed that iteration), then waits for its own slot: HBM needs
    // tens of KB in flight per SM, which two slots do not provide.
    int it = 0;
    if (threadIdx.x == 0)
        for (int i = 0; i < a.n_stage - 1; ++i) {
            const int t = blockIdx.x + i * gridDim.x;
            if (t < a.n_tiles) issue(t, i);
        }
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++it) {
        const int ahead = tile + (a.n_stage - 1) * gridDim.x;
        if (threadIdx.x == 0 && ahead < a.n_tiles) issue(ahead, stage == 0 ? a.n_stage - 1 : stage - 1);
        mbar_wait(bars + stage, phase);

        const uint32_t xb = smem_u32(smem + stage * slot_bytes), gb = xb + a.x_stage_bytes;
        // per-lane ldmatrix row addresses (matrix = lane/8, row = lane%8)
        const int mi = lane >> 3, rowi = lane & 7;
        // A (ga): matrix mi -> pixels +(mi/2)*8, co plane 2*cob + (mi%2)
        const uint32_t a_lane = gb + (2 * cob + (mi & 1)) * g_plane + ((mi >> 1) * 8 + rowi) * 16;
        // B (x):  matrix mi -> pixels +(mi%2)*8, ci plane 2*cib + (mi/2)
        const uint32_t b_lane = xb + (2 * cib + (mi >> 1)) * x_plane + ((mi & 1) * 8 + rowi) * 16;

        for (int r = rs; r < ((a.debug & 2) ? 0 : kWgTH); r += rsplit) {
            for (int w0 = 0; w0 < a.TW; w0 += 16) {
                uint32_t a0, a1, a2, a3;
                ldmatrix_x4_trans(a_lane + (r * a.TW + w0) * 16, a0, a1, a2, a3);
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int ky = tap / 3, kx = tap % 3;
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4_trans(b_lane + ((r + ky) * Wh + w0 + kx) * 16, b0, b1, b2, b3);
                    mma_bf16_16816(acc[tap][0], a0, a1, a2, a3, b0, b1);
                    mma_bf16_16816(acc[tap][1], a0, a1, a2, a3, b2, b3);
                }
            }
        }
        __syncthreads();  // everyone is done with this slot before the next iteration refills it
        if (++stage == a.n_stage) {
            stage = 0;
            phase ^= 1;
        }
    }

    // ---- flush: acc[tap][nb] = D[co = g (+8)][ci = nb*8 + 2t (+1)].  Every warp stores its accumulator block to its
    // own slice of shared memory (the operand slots are drained by now; plain stores -- shared-memory float atomics
    // are CAS loops), the CTA then sums the row-split slices of each output and issues one global atomic per
    // output.  Slices are laid out like the gradient tensor itself ([co][ci][tap], tap fastest), so the atomics
    // of a warp run over contiguous addresses: 144 consecutive floats per (block, co) row.
    if (a.debug & 1) return;
    float* s_red = reinterpret_cast<float*>(smem);   // [8 warps = rsplit x n_blk][16 co][16 ci][9 taps]
    const int g = lane >> 2, t = lane & 3;
    __syncthreads();
    {                                                 // (zeros when this CTA had no tile: its cluster still reads them)
        float* mine = s_red + warp * 2304;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int co = g + (k >> 1) * 8;
                    const int ci = nb * 8 + 2 * t + (k & 1);
                    mine[(co * 16 + ci) * 9 + tap] = acc[tap][nb][k];
                }
    }
    __syncthreads();
    const int n_out = n_blk * 2304;
    auto to_global = [&](int i, float v) {
        const int b = i / 2304, rem = i - b * 2304;
        const int co = rem / 144, j = rem - co * 144;        // j = ci * 9 + tap
        const int co_abs = co_group * a.co_g + (b / n_ci_blk) * 16 + co;
        const int ci0 = ci_group * a.ci_g + (b % n_ci_blk) * 16;
        atomicAdd(a.dw + (static_cast<size_t>(co_abs) * a.cin + ci0) * 9 + j, a.scale * v);
    };
    if (a.cluster == 1) {
        if (it > 0) {
            for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
                const int b = i / 2304, rem = i - b * 2304;
                float v = 0.f;
                for (int r = 0; r < rsplit; ++r) v += s_red[(r * n_blk + b) * 2304 + rem];     // warp = rs * n_blk + blk
                to_global(i, v);
            }
        }
        return;
    }
    // Cluster of a.cluster CTAs along the pixel dimension (same channel group, same output addresses): every CTA
    // first folds its row-split slices into the compact array s_red[0 .. n_out), then CTA q of the cluster sums slice
    // q of the outputs over all CTAs through distributed shared memory and issues those atomics: a.cluster times fewer
    // global atomics (they, not the MMAs, are what the low-resolution launches of this kernel cost).
    if (rsplit > 1) {
        for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
            const int b = i / 2304, rem = i - b * 2304;
            float v = 0.f;
            for (int r = 0; r < rsplit; ++r) v += s_red[(r * n_blk + b) * 2304 + rem];
            s_red[i] = v;          // slice (r = 0, b) starts at b * 2304: only this thread touches element (b, rem)
        }
    }
    cluster_sync();
    const int per = n_out / a.cluster;
    const uint32_t q = cluster_ctarank();
    const uint32_t base = smem_u32(s_red);
    for (int k = threadIdx.x; k < per; k += blockDim.x) {
        const int i = static_cast<int>(q) * per + k;
        float v = 0.f;
        for (int rk = 0; rk < a.cluster; ++rk) v += ld_dsmem_f32(base + i * 4, rk);
        to_global(i, v);
    }
    cluster_sync();               // nobody leaves while a neighbour may still read its shared memory
}

int conv3x3_wgrad(const void* x, const void* ga, float scale, float* dw, int B, int cin, int cout, int H, int W,
                  cudaStream_t st) {
    if (cin % 16 || cout % 16 || H % kWgTH || W % 16) {
        set_error("conv3x3_wgrad: unsupported shape cin=%d cout=%d H=%d W=%d", cin, cout, H, W);
        return NGAN_ERR_UNSUPPORTED;
    }
    WgradArgs a;
    a.B = B; a.H = H; a.W = W; a.cin = cin; a.cout = cout;
    a.ci_g = cin < 64 ? cin : 64;
    const int n_ci_blk = a.ci_g / 16;
    int co_cap = 16 * (8 / n_ci_blk);
    a.co_g = cout < co_cap ? cout : co_cap;
    int n_blk = n_ci_blk * (a.co_g / 16);
    if (8 % n_blk) {  // keep rsplit integral: shrink to a power-of-two block count
        set_error("conv3x3_wgrad: block split %d does not divide 8", n_blk);
        return NGAN_ERR_UNSUPPORTED;
    }
    a.n_ci_groups = cin / a.ci_g;
    const int n_groups = a.n_ci_groups * (cout / a.co_g);
    a.TW = W < 64 ? W : 64;
    if (a.ci_g >= 64 && a.TW > 32) a.TW = 32;
    a.tiles_x = W / a.TW;
    a.tiles_y = H / kWgTH;
    a.n_tiles = a.tiles_x * a.tiles_y * B;
    const uint32_t x_plane = (kWgTH + 2) * (a.TW + 2) * 16, g_plane = kWgTH * a.TW * 16;
    a.x_stage_bytes = ((a.ci_g / 8) * x_plane + 127) & ~127u;
    a.g_stage_bytes = ((a.co_g / 8) * g_plane + 127) & ~127u;
    a.scale = scale;
    a.dw = dw;
    int n_stage = static_cast<int>((112u * 1024) / (a.x_stage_bytes + a.g_stage_bytes));   // two CTAs per SM stay resident
    if (n_stage > kWgMaxStages) n_stage = kWgMaxStages;
    if (n_stage < 2) n_stage = 2;
    a.n_stage = n_stage;
    static const int dbg = getenv("NGAN_WGRAD_DEBUG") ? atoi(getenv("NGAN_WGRAD_DEBUG")) : 0;
    a.debug = dbg;
    const uint32_t stage_total = n_stage * (a.x_stage_bytes + a.g_stage_bytes) + kWgMaxStages * 8;
    const uint32_t red_bytes = 8u * 2304 * sizeof(float);   // flush slices (one per warp) alias the operand slots
    const uint32_t smem_bytes = (stage_total > red_bytes ? stage_total : red_bytes) + 128;

    CUtensorMap tmx, tmg;
    int rc = make_c8_tensor_map(&tmx, x, B, cin, H, W, a.TW + 2, kWgTH + 2, a.ci_g / 8);
    if (rc) return rc;
    rc = make_c8_tensor_map(&tmg, ga, B, cout, H, W, a.TW, kWgTH, a.co_g / 8);
    if (rc) return rc;

    static bool configured = false;
    if (!configured) {
        cudaError_t e =
            cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(wgrad)");
        configured = true;
    }
    // CTAs: every CTA ends with a flush of n_blk*2304 atomics, so the CTA count is what small problems pay for and
    // the tiles per CTA what large ones pay for.  Measured (B200, graph-timed): one resident wave (148 CTAs over
    // all channel groups) is best up to a few tiles per CTA, two CTAs per SM beyond that.
    const long long work = static_cast<long long>(a.n_tiles) * n_groups;
    static const int target_env = getenv("NGAN_WGRAD_CTAS") ? atoi(getenv("NGAN_WGRAD_CTAS")) : 0;
    const int target = target_env ? target_env : (work >= 4LL * 148 ? 2 * 148 : 148);
    int per_group = target / n_groups;
    if (per_group > a.n_tiles) per_group = a.n_tiles;
    if (per_group < 1) per_group = 1;
    // Clusters (measured, graph-timed): 4 CTAs help the 16x16 / 32x32 layers, whose launches are all flush
    // (21 -> 15 us); from 64x64 up the co-scheduling constraint costs more than the saved atomics (512x512: 56 -> 95 us).
    static const int cluster_env = getenv("NGAN_WGRAD_CLUSTER") ? atoi(getenv("NGAN_WGRAD_CLUSTER")) : -1;
    const int cluster_max = cluster_env >= 0 ? cluster_env : (H <= 32 ? 4 : 1);
    int cs = 1;
    while (cs * 2 <= cluster_max && cs * 2 <= 8 && cs * 2 <= per_group) cs *= 2;
    per_group = per_group / cs * cs;
    a.cluster = cs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(per_group, n_groups);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cs > 1 ? 1 : 0;
    cudaError_t le = cudaLaunchKernelEx(&cfg, conv3x3_wgrad_kernel, tmx, tmg, a);
    if (le != cudaSuccess) return check_cuda(le, "cudaLaunchKernelEx(conv3x3_wgrad)");
    return check_launch("conv3x3_wgrad");
}

}  // namespace ngan
