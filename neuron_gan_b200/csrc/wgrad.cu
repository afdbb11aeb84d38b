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
// Reduction over pixels is DETERMINISTIC: CTAs are persistent over pixel tiles in a fixed order, each writes its
// partial block sums with plain coalesced stores into a workspace (one image of cout*cin*9 floats per pixel CTA or
// cluster, in register order), and wgrad_reduce_kernel adds the images in index order.  No atomics: the gradient
// is bit-identical from run to run.
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
    int cluster;          // CTAs per cluster along grid.x (1, 2, 4 or 8): DSMEM reduction before the partial store
    uint32_t x_stage_bytes, g_stage_bytes;
    float* partial;       // [gridDim.x / cluster][gridDim.y groups][n_blk][18][32][4]: register order, see the flush
    long long* dbg;       // timing experiments (ngan_debug_wgrad_trace): 8 globaltimer stamps per CTA, or null
};
long long* g_wgrad_trace = nullptr;
__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define WG_STAMP(k)                                                                                        \
    do {                                                                                                   \
        if (a.dbg && threadIdx.x == 0) a.dbg[(blockIdx.y * gridDim.x + blockIdx.x) * 8 + (k)] = gtime(); \
    } while (0)

constexpr int kWgTH = 8;
constexpr int kWgMaxStages = 6;

struct WgradPlan {
    int ci_g, co_g, n_ci_groups, n_groups, n_blk, TW, tiles_x, tiles_y, n_tiles, per_group, rb, cluster;
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
    static const int tw_env = getenv("NGAN_WGRAD_TW") ? atoi(getenv("NGAN_WGRAD_TW")) : 0;    // tuning experiments
    p->TW = W < 64 ? W : 64;
    if (p->ci_g >= 64 && p->TW > 32) p->TW = 32;
    if (tw_env && p->TW > tw_env) p->TW = tw_env;
    p->tiles_x = W / p->TW;
    p->tiles_y = H / kWgTH;
    p->n_tiles = p->tiles_x * p->tiles_y * B;
    // rows per strip: with one accumulator block per CTA all eight warps split the tile, so strips are half height
    p->rb = p->n_blk == 1 ? 4 : 8;
    // CTAs along the pixel dimension: one resident wave (148 CTAs over all channel groups) up to a few tiles per
    // CTA, two CTAs per SM beyond that (measured in round 1, graph-timed).  Every pixel CTA (or cluster) ends with one
    // partial image of its group's blocks.  NGAN_WGRAD_CLUSTER=2|4|8 first adds the partials of neighbouring CTAs
    // through distributed shared memory (in rank order); measured slower than plain partial images, off by default.
    static const int target_env = getenv("NGAN_WGRAD_CTAS") ? atoi(getenv("NGAN_WGRAD_CTAS")) : 0;
    const long long work = static_cast<long long>(p->n_tiles) * p->n_groups;
    const int target = target_env ? target_env : (work >= 4LL * 148 ? 2 * 148 : 148);
    int per_group = target / p->n_groups;
    if (per_group > p->n_tiles) per_group = p->n_tiles;
    if (per_group < 1) per_group = 1;
    static const int cluster_env = getenv("NGAN_WGRAD_CLUSTER") ? atoi(getenv("NGAN_WGRAD_CLUSTER")) : -1;
    // (round 1 clustered the 16x16 / 32x32 layers to save global atomics; with partial images and plain stores the
    // cluster's co-scheduling constraint only costs: 128->64 @32x32 24.7 us alone vs 31.6 us in clusters of 4)
    const int cluster_max = cluster_env >= 0 ? cluster_env : 1;
    int cs = 1;
    while (cs * 2 <= cluster_max && cs * 2 <= 8 && cs * 2 <= per_group) cs *= 2;
    per_group = per_group / cs * cs;
    p->cluster = cs;
    p->per_group = per_group;
    return true;
}

size_t conv3x3_wgrad_workspace_bytes(int B, int cin, int cout, int H, int W) {
    WgradPlan p;
    if (!wgrad_plan(B, cin, cout, H, W, &p)) return 0;
    return static_cast<size_t>(p.per_group / p.cluster) * cin * cout * 9 * sizeof(float);
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
__device__ __forceinline__ float4 ld_dsmem_f32x4(uint32_t local_saddr, uint32_t rank) {
    uint32_t ra;
    float4 v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(rank));
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"(ra)
                 : "memory");
    return v;
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
    WG_STAMP(0);
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
    WG_STAMP(1);

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
        if (tile == static_cast<int>(blockIdx.x)) WG_STAMP(2);

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

    // ---- flush.  acc[tap][nb][k] = D[co = g + 8*(k>>1)][ci = nb*8 + 2t + (k&1)] with g = lane/4, t = lane%4.  The
    // partial image of this pixel CTA (or cluster) keeps that REGISTER ORDER -- [group][block][tap*2+nb][lane][k] --
    // so every store below is a conflict-free, fully coalesced float4 and nothing is transposed or divided here (the
    // transposing flush of round 1 cost 4-16 us per launch, more than the main loop of the low-resolution layers);
    // wgrad_reduce_kernel undoes the permutation once, when it writes the summed gradient.
    WG_STAMP(3);
    float4* s4 = reinterpret_cast<float4*>(smem);    // [8 warps = rsplit x n_blk][18][32 lanes] float4
    {
        float4* mine = s4 + warp * 576;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap)
#pragma unroll
            for (int nb = 0; nb < 2; ++nb)
                mine[(tap * 2 + nb) * 32 + lane] =
                    make_float4(acc[tap][nb][0], acc[tap][nb][1], acc[tap][nb][2], acc[tap][nb][3]);
    }
    __syncthreads();
    const int n4 = n_blk * 576;
    float4* out4 = reinterpret_cast<float4*>(a.partial) +
                   (static_cast<size_t>(blockIdx.x / a.cluster) * gridDim.y + group) * n4;
    auto fold = [&](int i) {             // sum over the warps that shared block b, in warp order
        const int b = i / 576, rem = i - b * 576;
        float4 v = s4[b * 576 + rem];    // warp = rs * n_blk + blk
        for (int r = 1; r < rsplit; ++r) {
            const float4 w = s4[(r * n_blk + b) * 576 + rem];
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        return v;
    };
    if (a.cluster == 1) {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) out4[i] = fold(i);
        WG_STAMP(4);
        return;
    }
    // Cluster of a.cluster CTAs along the pixel dimension (same channel group, same outputs): every CTA first folds
    // its warps' slices into the compact array s4[0 .. n4), then CTA q of the cluster adds slice q of the outputs
    // over all CTAs in rank order through distributed shared memory and stores it.
    if (rsplit > 1) {
        for (int i = threadIdx.x; i < n4; i += blockDim.x) s4[i] = fold(i);   // element i is only touched by this thread
    }
    cluster_sync();
    const int per = n4 / a.cluster;
    const uint32_t q = cluster_ctarank();
    const uint32_t base = smem_u32(s4);
    for (int k = threadIdx.x; k < per; k += blockDim.x) {
        const int i = static_cast<int>(q) * per + k;
        float4 v = ld_dsmem_f32x4(base + i * 16, 0);
        for (int rk = 1; rk < a.cluster; ++rk) {
            const float4 w = ld_dsmem_f32x4(base + i * 16, rk);
            v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
        }
        out4[i] = v;
    }
    cluster_sync();               // nobody leaves while a neighbour may still read its shared memory
    WG_STAMP(4);
}

// Second stage: dW (+)= scale * sum_p partial[p], rows added in index order; one thread per float4 of the register-
// order image (coalesced reads), which it scatters to its four places in the torch-layout gradient [cout][cin][3][3].
template <int OPB>
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float4* __restrict__ partials, int n_partials,
                                                           int n4_total, int n_blk, int n_ci_blk, int n_ci_groups,
                                                           int ci_g, int co_g, int cin, float scale,
                                                           float* __restrict__ dw, int accumulate) {
    pdl_trigger();
    pdl_wait();
    constexpr int S = 256 / OPB;
    __shared__ float4 red[S][OPB];
    const int o = threadIdx.x % OPB, s = threadIdx.x / OPB;
    const int e = blockIdx.x * OPB + o;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e < n4_total) {
        const float4* src = partials + e;
        int p = s;
        for (; p + 3 * S < n_partials; p += 4 * S) {      // four loads in flight, added in row order
            float4 t[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) t[k] = __ldg(src + static_cast<size_t>(p + k * S) * n4_total);
#pragma unroll
            for (int k = 0; k < 4; ++k) { v.x += t[k].x; v.y += t[k].y; v.z += t[k].z; v.w += t[k].w; }
        }
        for (; p < n_partials; p += S) {
            const float4 t = __ldg(src + static_cast<size_t>(p) * n4_total);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
    }
    red[s][o] = v;
    __syncthreads();
    if (s != 0 || e >= n4_total) return;
#pragma unroll
    for (int k = 1; k < S; ++k) { v.x += red[k][o].x; v.y += red[k][o].y; v.z += red[k][o].z; v.w += red[k][o].w; }
    const int per_group = n_blk * 576;
    const int group = e / per_group, r1 = e - group * per_group;
    const int blk = r1 / 576, r2 = r1 - blk * 576;
    const int q = r2 >> 5, lane = r2 & 31;
    const int tap = q >> 1, nb = q & 1, g = lane >> 2, t = lane & 3;
    const int co0 = (group / n_ci_groups) * co_g + (blk / n_ci_blk) * 16 + g;
    const int ci0 = (group % n_ci_groups) * ci_g + (blk % n_ci_blk) * 16 + nb * 8 + 2 * t;
    const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float* dst = dw + (static_cast<size_t>(co0 + (k >> 1) * 8) * cin + ci0 + (k & 1)) * 9 + tap;
        *dst = accumulate ? *dst + scale * vals[k] : scale * vals[k];
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
    a.cluster = p.cluster;
    a.dbg = g_wgrad_trace;
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
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.per_group, p.n_groups);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int n_attr = 0;
    if (pdl_enabled()) {
        attr[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n_attr].val.programmaticStreamSerializationAllowed = 1;
        ++n_attr;
    }
    if (p.cluster > 1) {
        attr[n_attr].id = cudaLaunchAttributeClusterDimension;
        attr[n_attr].val.clusterDim.x = p.cluster;
        attr[n_attr].val.clusterDim.y = 1;
        attr[n_attr].val.clusterDim.z = 1;
        ++n_attr;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n_attr;
    cudaError_t le = p.rb == 4 ? cudaLaunchKernelEx(&cfg, conv3x3_wgrad_kernel<4>, tmx, tmg, a)
                               : cudaLaunchKernelEx(&cfg, conv3x3_wgrad_kernel<8>, tmx, tmg, a);
    if (le != cudaSuccess) return check_cuda(le, "cudaLaunchKernelEx(conv3x3_wgrad)");
    rc = check_launch("conv3x3_wgrad");
    if (rc) return rc;
    const int n4_total = cin * cout * 9 / 4, n_part = p.per_group / p.cluster;
    const int n_ci_blk = p.ci_g / 16;
    cudaError_t re;
    if (n_part <= 32)
        re = launch_pdl(wgrad_reduce_kernel<64>, dim3((n4_total + 63) / 64), dim3(256), 0, st,
                        reinterpret_cast<const float4*>(workspace), n_part, n4_total, p.n_blk, n_ci_blk, p.n_ci_groups,
                        p.ci_g, p.co_g, cin, scale, dw, accumulate);
    else
        re = launch_pdl(wgrad_reduce_kernel<16>, dim3((n4_total + 15) / 16), dim3(256), 0, st,
                        reinterpret_cast<const float4*>(workspace), n_part, n4_total, p.n_blk, n_ci_blk, p.n_ci_groups,
                        p.ci_g, p.co_g, cin, scale, dw, accumulate);
    if (re != cudaSuccess) return check_cuda(re, "cudaLaunchKernelEx(wgrad_reduce)");
    return check_launch("wgrad_reduce");
}

}  // namespace ngan
