// Weight gradient of the equalised-LR 3x3 convolution (what autograd's convolution_backward computes for
// Conv2d_normalized, reference models.py:203-204), also used for the "wgrad of dgrad" term of the
// gradient-penalty double backward:
//     dW[co][ci][ky][kx] (+)= scale * sum_{b,y,x} ga[b,co,y,x] * x[b,ci,y+ky-1,x+kx-1]
//
// GEMM view: M = co, N = ci (per tap), K = pixels.  Why this reduction stays on the warp-level tensor-core
// path (mma.sync m16n8k16, bf16 -> fp32) and not tcgen05: with K = pixels both operands are MN-major C8 tiles, a
// tcgen05.mma covers only 16 pixels per instruction and always fetches a 64/128-row A operand from shared memory
// (2-4 KB per 16 pixels and tap, against the 0.5 KB the 16 real channels hold).  For the 16/32-channel layers that
// carry the pixels that is 70-140 shared-memory cycles per 16 pixels and SM where HBM leaves 44;
// mma.sync reads exactly the fragments it needs.
//
// Every warp owns one 16(co) x 16(ci) x 9(tap) accumulator block in registers for the whole kernel.  Operand
// tiles arrive by TMA (ring of stages, mbarrier-signalled) straight from the C8-planar tensors; a C8 tile
// [pixel][8 channels] is the transpose of what mma wants, which is exactly what ldmatrix.trans delivers.
// Inside a tile a warp walks a 16-pixel-wide column strip DOWN the rows: the x fragment of haloed row rho and
// horizontal tap kx serves the three vertical taps (ga rows rho, rho-1, rho-2), so a strip of RB rows costs
// 3*(RB+2) + RB ldmatrix instead of 10*RB -- the kernel was bound by that shared-memory traffic, not by HMMA.
//
// Reduction over pixels is DETERMINISTIC: CTAs are persistent over pixel tiles in a fixed order, each writes
// its partial block sums with plain coalesced stores into a workspace laid out like dW ([pixel CTA][cout][cin][9]),
// and reduce_partials (elementwise.cu) adds the partials in index order.  No atomics: the gradient is bit-identical
// from run to run, and the low-resolution launches no longer pay 1-2.6 M global atomics for a 10-600 KB result.
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
    uint32_t x_stage_bytes, g_stage_bytes;
    float* partial;       // [gridDim.x][cout][cin][9]
};

constexpr int kWgTH = 8;
constexpr int kWgMaxStages = 6;

struct WgradPlan {
    int ci_g, co_g, n_ci_groups, n_groups, n_blk, TW, tiles_x, tiles_y, n_tiles, per_group, rb;
};

static bool wgrad_plan(int B, int cin, int cout, int H, int W, WgradPlan* p) {
    if (cin % 16 || cout % 16 || H % kWgTH || W % 16 || B <= 0) return false;
    p->ci_g = cin < 64 ? cin : 64;
    const int n_ci_blk = p->ci_g / 16;
    const int co_cap = 16 * (8 / n_ci_blk);
    p->co_g = cout < co_cap ? cout : co_cap;
    p->n_blk = n_ci_blk * (p->co_g / 16);
    if (8 % p->n_blk) return false;
    p->n_ci_groups = cin / p->ci_g;
    p->n_groups = p->n_ci_groups * (cout / p->co_g);
    p->TW = W < 64 ? W : 64;
    if (p->ci_g >= 64 && p->TW > 32) p->TW = 32;
    p->tiles_x = W / p->TW;
    p->tiles_y = H / kWgTH;
    p->n_tiles = p->tiles_x * p->tiles_y * B;
    // rows per strip: with one accumulator block per CTA all eight warps split the tile, so strips are half height
    p->rb = p->n_blk == 1 ? 4 : 8;
    // CTAs along the pixel dimension.  Large problems: two CTAs per SM over all channel groups.  Small ones: every
    // pixel CTA costs one partial image of the group's block (written, then read by the reduction), so no more CTAs
    // than keep that traffic near the operand traffic, and never more than tiles.
    static const int target_env = getenv("NGAN_WGRAD_CTAS") ? atoi(getenv("NGAN_WGRAD_CTAS")) : 0;
    const long long work = static_cast<long long>(p->n_tiles) * p->n_groups;
    const int target = target_env ? target_env : (work >= 4LL * 148 ? 2 * 148 : 148);
    int per_group = target / p->n_groups;
    const long long in_bytes = static_cast<long long>(B) * H * W * (cin + cout) * 2;
    const long long blk_bytes = static_cast<long long>(cin) * cout * 9 * 4;
    long long cap = 2 * in_bytes / blk_bytes;
    if (cap < 4) cap = 4;
    if (!target_env && per_group > cap) per_group = static_cast<int>(cap);
    if (per_group > p->n_tiles) per_group = p->n_tiles;
    if (per_group < 1) per_group = 1;
    p->per_group = per_group;
    return true;
}

size_t conv3x3_wgrad_workspace_bytes(int B, int cin, int cout, int H, int W) {
    WgradPlan p;
    if (!wgrad_plan(B, cin, cout, H, W, &p)) return 0;
    return static_cast<size_t>(p.per_group) * cin * cout * 9 * sizeof(float);
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, "
        "%2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One 16-pixel-wide strip of RB output rows: haloed x rows 0..RB+1 (relative), ga rows 0..RB-1.  a_addr / b_addr are
// this lane's ldmatrix row addresses of ga row 0 / x row 0 (tap kx = 0); a_row / b_row the byte pitch of a tile row.
// The x fragment (rho, kx) meets ga row rho - ky for ky = 0..2; three ga fragments are live at a time.
template <int RB>
__device__ __forceinline__ void wgrad_strip(float (&acc)[9][2][4], uint32_t a_addr, uint32_t b_addr, uint32_t a_row,
                                            uint32_t b_row) {
    uint32_t af[3][4];
#pragma unroll
    for (int rho = 0; rho < RB + 2; ++rho) {
        if (rho < RB) ldmatrix_x4_trans(a_addr + rho * a_row, af[rho % 3]);
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            uint32_t bf[4];
            ldmatrix_x4_trans(b_addr + rho * b_row + kx * 16, bf);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int r = rho - ky;
                if (r >= 0 && r < RB) {
                    mma_bf16_16816(acc[ky * 3 + kx][0], af[r % 3], bf[0], bf[1]);
                    mma_bf16_16816(acc[ky * 3 + kx][1], af[r % 3], bf[2], bf[3]);
                }
            }
        }
    }
}

template <int RB>
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x,
                                                            const __grid_constant__ CUtensorMap tmap_g,
                                                            const WgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();        // the reduction that follows may be scheduled while this grid drains (it waits for all of it)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    // ring of n_stage {x tile, ga tile} slots followed by the mbarriers
    const uint32_t slot_bytes = a.x_stage_bytes + a.g_stage_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + a.n_stage * slot_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int group = blockIdx.y;
    const int ci_group = group % a.n_ci_groups, co_group = group / a.n_ci_groups;
    const int n_ci_blk = a.ci_g / 16, n_co_blk = a.co_g / 16;
    const int n_blk = n_ci_blk * n_co_blk;  // 1, 2, 4 or 8
    const int rsplit = 8 / n_blk;           // warps sharing one accumulator block: they split the strips of a tile
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
    pdl_wait();           // x / ga come from the kernels before this one in the stream

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
    if (threadIdx.x == 0)
        for (int i = 0; i < a.n_stage - 1; ++i) {
            const int t = blockIdx.x + i * gridDim.x;
            if (t < a.n_tiles) issue(t, i);
        }
    const int strips = a.TW >> 4;
    const int n_parts = strips * (kWgTH / RB);
    const uint32_t a_row = a.TW * 16, b_row = Wh * 16;
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
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

        for (int p = rs; p < n_parts; p += rsplit) {
            const int s = p % strips, r0 = (p / strips) * RB;
            wgrad_strip<RB>(acc, a_lane + (r0 * a.TW + s * 16) * 16, b_lane + (r0 * Wh + s * 16) * 16, a_row, b_row);
        }
        __syncthreads();  // everyone is done with this slot before the next iteration refills it
        if (++stage == a.n_stage) {
            stage = 0;
            phase ^= 1;
        }
    }

    // ---- flush: acc[tap][nb] = D[co = g (+8)][ci = nb*8 + 2t (+1)].  Every warp stores its accumulator block to its
    // own slice of shared memory (the operand slots are drained by now), the CTA sums the slices of the warps that
    // shared a block in warp order, and writes the block into this pixel CTA's partial image, which is laid out like
    // the gradient tensor itself ([co][ci][tap]): 144 consecutive floats per (block, co) row.
    float* s_red = reinterpret_cast<float*>(smem);   // [8 warps = rsplit x n_blk][16 co][16 ci][9 taps]
    const int g = lane >> 2, t = lane & 3;
    {
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
    float* out = a.partial + static_cast<size_t>(blockIdx.x) * a.cout * a.cin * 9;
    const int n_out = n_blk * 2304;
    for (int i = threadIdx.x; i < n_out; i += blockDim.x) {
        const int b = i / 2304, rem = i - b * 2304;
        float v = 0.f;
        for (int r = 0; r < rsplit; ++r) v += s_red[(r * n_blk + b) * 2304 + rem];     // warp = rs * n_blk + blk
        const int co = rem / 144, j = rem - co * 144;        // j = ci * 9 + tap
        const int co_abs = co_group * a.co_g + (b / n_ci_blk) * 16 + co;
        const int ci0 = ci_group * a.ci_g + (b % n_ci_blk) * 16;
        out[(static_cast<size_t>(co_abs) * a.cin + ci0) * 9 + j] = v;
    }
}

int conv3x3_wgrad(const void* x, const void* ga, float scale, float* dw, int accumulate, float* workspace, int B,
                  int cin, int cout, int H, int W, cudaStream_t st) {
    WgradPlan p;
    if (!wgrad_plan(B, cin, cout, H, W, &p)) {
        set_error("conv3x3_wgrad: unsupported shape cin=%d cout=%d H=%d W=%d", cin, cout, H, W);
        return NGAN_ERR_UNSUPPORTED;
    }
    WgradArgs a;
    a.B = B; a.H = H; a.W = W; a.cin = cin; a.cout = cout;
    a.ci_g = p.ci_g; a.co_g = p.co_g; a.n_ci_groups = p.n_ci_groups;
    a.TW = p.TW; a.tiles_x = p.tiles_x; a.tiles_y = p.tiles_y; a.n_tiles = p.n_tiles;
    const uint32_t x_plane = (kWgTH + 2) * (a.TW + 2) * 16, g_plane = kWgTH * a.TW * 16;
    a.x_stage_bytes = ((a.ci_g / 8) * x_plane + 127) & ~127u;
    a.g_stage_bytes = ((a.co_g / 8) * g_plane + 127) & ~127u;
    a.partial = workspace;
    int n_stage = static_cast<int>((112u * 1024) / (a.x_stage_bytes + a.g_stage_bytes));   // two CTAs per SM stay resident
    if (n_stage > kWgMaxStages) n_stage = kWgMaxStages;
    if (n_stage < 2) n_stage = 2;
    a.n_stage = n_stage;
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
        cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             227 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(conv3x3_wgrad_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(wgrad)");
        configured = true;
    }
    cudaError_t le = p.rb == 4 ? launch_pdl(conv3x3_wgrad_kernel<4>, dim3(p.per_group, p.n_groups), dim3(256),
                                            smem_bytes, st, tmx, tmg, a)
                               : launch_pdl(conv3x3_wgrad_kernel<8>, dim3(p.per_group, p.n_groups), dim3(256),
                                            smem_bytes, st, tmx, tmg, a);
    if (le != cudaSuccess) return check_cuda(le, "cudaLaunchKernelEx(conv3x3_wgrad)");
    rc = check_launch("conv3x3_wgrad");
    if (rc) return rc;
    const long long n = static_cast<long long>(cin) * cout * 9;
    return reduce_partials(workspace, p.per_group, n, n, scale, dw, accumulate, st);
}

}  // namespace ngan
