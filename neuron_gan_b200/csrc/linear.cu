// Generator stem: Linear_normalized(latent -> C*S*S) + Unflatten + LeakyReLU + PixelNorm
// (reference models.py:299-311, 240-241) and its weight gradient.
//
// FLOPs are negligible (2*B*16.8 MFLOP); both kernels are bound by streaming the 16.8 M-element weight
// (bf16 operand image, 33.5 MB, forward: tcgen05 GEMM per pixel) or its fp32 gradient (67 MB, backward) through
// HBM once.
#include "common.cuh"
#include "kernels.h"

namespace ngan {

// ---- bf16 operand image of the weight: [S*S pixels][K/8][C channels][8 k] -----------------------------------------
// Row f = c*S*S + p of the fp32 master [C*S*S][K] (Unflatten to [C][S][S], models.py:301) is channel c of pixel p.
// Grouping the C rows of one pixel makes (a) the B operand of that pixel's GEMM one contiguous block that is
// already in the no-swizzle K-major UMMA layout (core matrix = 8 channels x 16 B), streamed with plain bulk
// copies, and (b) the PixelNorm reduction over channels a per-thread reduction over accumulator columns.
__device__ __forceinline__ size_t linear_shadow_index(size_t i, int K, int C, int SS) {
    const size_t f = i / K;
    const int k = static_cast<int>(i - f * K);
    const int c = static_cast<int>(f / SS), p = static_cast<int>(f - static_cast<size_t>(c) * SS);
    return ((static_cast<size_t>(p) * (K / 8) + (k >> 3)) * C + c) * 8 + (k & 7);
}
__global__ void prep_linear_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, size_t n,
                                          int K, int C, int SS) {
    const size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i >= n) return;
    const float4 v = *reinterpret_cast<const float4*>(w + i);
    uint2 o;
    o.x = pack_bf16(v.x, v.y);
    o.y = pack_bf16(v.z, v.w);
    *reinterpret_cast<uint2*>(out + linear_shadow_index(i, K, C, SS)) = o;
}
int prep_linear_weight(const float* w, void* wb, int K, int C, int S, cudaStream_t st) {
    if (K % 8 || C % 8) {
        set_error("prep_linear_weight: K and C must be multiples of 8 (K=%d C=%d)", K, C);
        return NGAN_ERR_UNSUPPORTED;
    }
    const size_t n = static_cast<size_t>(C) * S * S * K;
    const size_t threads = n / 4;
    prep_linear_weight_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, st>>>(
        w, static_cast<__nv_bfloat16*>(wb), n, K, C, S * S);
    return check_launch("prep_linear_weight");
}

// ---- latent batch as a UMMA A operand: zp [K/8][Bpad rows][8] bf16, rows >= B zero --------------------------------
__global__ void linear_prep_z_kernel(const float* __restrict__ z, uint4* __restrict__ zp, int B, int K, int Bpad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (K / 8) * Bpad) return;
    const int kg = i / Bpad, b = i - kg * Bpad;
    uint4 o = make_uint4(0, 0, 0, 0);
    if (b < B) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(z + static_cast<size_t>(b) * K + kg * 8));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(z + static_cast<size_t>(b) * K + kg * 8) + 1);
        o.x = pack_bf16(v0.x, v0.y);
        o.y = pack_bf16(v0.z, v0.w);
        o.z = pack_bf16(v1.x, v1.y);
        o.w = pack_bf16(v1.z, v1.w);
    }
    zp[i] = o;
}

// ---- y = PixelNorm(LeakyReLU(scale * z W^T)) as one tcgen05 GEMM per pixel ------------------------------------------
// CTA (p, mt): D[b, c] = sum_k z[mt*128 + b, k] * W[c*SS + p, k]   (M = 128 samples, N = C channels, K = 512).
// A = the latent tile (128 KB, one set of bulk copies), B = this pixel's weight block streamed in 32 KB K-chunks
// through a 3-deep ring; accumulators in TMEM (C fp32 columns); thread b of the four epilogue warps owns sample
// b's row: LeakyReLU, PixelNorm over the C columns, C8 store.  The kernel is bound by streaming the weight
// (33.5 MB for the default 512 -> 128x16x16) once per M-tile.
constexpr int kLinStages = 3;
constexpr uint32_t kLinStageBytes = 32768;
template <int K, int C>
__global__ void __launch_bounds__(192) linear_fwd_umma_kernel(const uint4* __restrict__ zp,
                                                              const __nv_bfloat16* __restrict__ wb, float scale,
                                                              float leak, uint4* __restrict__ y, float* __restrict__ r,
                                                              int B, int Bpad, int SS) {
    pdl_trigger();   // the next kernel (a PDL-launched conv) may start its prologue now
    extern __shared__ uint8_t smem_raw[];
    constexpr int KG = K / 8;                          // k-groups (16-byte columns of the operands)
    constexpr int CHUNK_KG = kLinStageBytes / (C * 16);   // k-groups per weight chunk
    constexpr int N_CHUNK = KG / CHUNK_KG;
    constexpr uint32_t Z_BYTES = KG * 128 * 16;
    constexpr uint32_t IDESC = umma_idesc_bf16(128, C);
    static_assert(CHUNK_KG >= 2 && CHUNK_KG % 2 == 0 && KG % CHUNK_KG == 0, "unsupported K/C");

    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    uint8_t* s_z = smem;
    uint8_t* s_w = smem + Z_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_w + kLinStages * kLinStageBytes);
    uint64_t* bar_z = bars;
    uint64_t* bar_full = bars + 1;
    uint64_t* bar_empty = bars + 1 + kLinStages;
    uint64_t* bar_acc = bars + 1 + 2 * kLinStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * kLinStages);

    const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x, mt = blockIdx.y;

    if (threadIdx.x == 0) {
        mbar_init(bar_z, 1);
        for (int s = 0; s < kLinStages; ++s) {
            mbar_init(bar_full + s, 1);
            mbar_init(bar_empty + s, 1);
        }
        mbar_init(bar_acc, 1);
        mbar_fence_init();
    }
    __syncwarp();
    if (warp == 0) tmem_alloc(tmem_slot, C < 32 ? 32 : C);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    if (warp == 0) {
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(wb) + static_cast<size_t>(p) * KG * C * 16;
            // first weight chunks before the latent tile: they are the long pole
            for (int ch = 0; ch < N_CHUNK; ++ch) {
                const int stage = ch % kLinStages;
                if (ch >= kLinStages) mbar_wait(bar_empty + stage, ((ch / kLinStages) - 1) & 1);
                mbar_arrive_expect_tx(bar_full + stage, kLinStageBytes);
                bulk_load_1d(s_w + stage * kLinStageBytes, wsrc + static_cast<size_t>(ch) * kLinStageBytes,
                             kLinStageBytes, bar_full + stage);
                if (ch == 0) {
                    mbar_arrive_expect_tx(bar_z, Z_BYTES);
                    for (int kg = 0; kg < KG; ++kg)
                        bulk_load_1d(s_z + kg * 2048, zp + (static_cast<size_t>(kg) * Bpad + mt * 128), 2048, bar_z);
                }
            }
        }
    } else if (warp == 1) {
        mbar_wait_warp(bar_z, 0, lane);
        const uint32_t z_base = smem_u32(s_z), w_base = smem_u32(s_w);
        for (int ch = 0; ch < N_CHUNK; ++ch) {
            const int stage = ch % kLinStages;
            mbar_wait_warp(bar_full + stage, (ch / kLinStages) & 1, lane);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t hi = umma_desc_hi(128);
#pragma unroll
                for (int kk = 0; kk < CHUNK_KG / 2; ++kk) {
                    const uint32_t a_lo = umma_desc_lo(z_base + (ch * CHUNK_KG + 2 * kk) * 2048, 2048);
                    const uint32_t b_lo = umma_desc_lo(w_base + stage * kLinStageBytes + (2 * kk) * C * 16, C * 16);
                    umma_bf16_2x32(tmem_base, a_lo, hi, b_lo, hi, IDESC, (ch | kk) != 0);
                }
                umma_commit(bar_empty + stage);
                if (ch == N_CHUNK - 1) umma_commit(bar_acc);
            }
            __syncwarp();
        }
    } else {
        const int quad = warp & 3;
        const int b = mt * 128 + quad * 32 + lane;
        mbar_wait_warp(bar_acc, 0, lane);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        // lrelu and PixelNorm commute with the positive scale: normalise lrelu(acc) with eps / scale^2
        float ss = 0.f;
        float v[16];
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 16) {
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float h = fmaxf(v[i], leak * v[i]);
                ss = fmaf(h, h, ss);
            }
        }
        const float inv_s = 1.0f / scale;
        const float kn = rsqrtf(ss * (1.0f / C) + 1e-8f * inv_s * inv_s);
        // (tcgen05.ld is warp-collective: every lane runs the loop, rows beyond the batch only skip the stores)
#pragma unroll 1
        for (int c0 = 0; c0 < C; c0 += 16) {
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], leak * v[i]) * kn;
            if (b < B) {
                y[(static_cast<size_t>(b) * (C / 8) + c0 / 8) * SS + p] = pack8(v);
                y[(static_cast<size_t>(b) * (C / 8) + c0 / 8 + 1) * SS + p] = pack8(v + 8);
            }
        }
        if (b < B && r) r[static_cast<size_t>(b) * SS + p] = kn * inv_s;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, C < 32 ? 32 : C);
}

template <int K, int C>
static int launch_linear(const uint4* zp, const void* wb, float scale, float leak, void* y, float* r, int B, int Bpad,
                         int SS, cudaStream_t st) {
    auto kern = linear_fwd_umma_kernel<K, C>;
    const uint32_t smem = 128 + (K / 8) * 128 * 16 + kLinStages * kLinStageBytes + (2 + 2 * kLinStages) * 8 + 16;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(linear)");
        configured = true;
    }
    kern<<<dim3(SS, Bpad / 128), 192, smem, st>>>(zp, static_cast<const __nv_bfloat16*>(wb), scale, leak,
                                                  static_cast<uint4*>(y), r, B, Bpad, SS);
    return check_launch("linear_fwd");
}

// workspace: (K/8) * Bpad * 16 bytes, Bpad = B rounded up to 128 (see ngan_linear_fwd_workspace_bytes)
int linear_fwd_pn(const float* z, const void* wb, float scale, float leak, void* y, float* r, void* workspace, int B,
                  int K, int C, int S, cudaStream_t st) {
    if (K != 512 || (C != 64 && C != 128 && C != 256)) {
        set_error("linear_fwd_pn: built for latent_dim 512 and 64/128/256 channels (got K=%d C=%d)", K, C);
        return NGAN_ERR_UNSUPPORTED;
    }
    const int Bpad = (B + 127) / 128 * 128;
    uint4* zp = static_cast<uint4*>(workspace);
    const int n = (K / 8) * Bpad;
    linear_prep_z_kernel<<<(n + 255) / 256, 256, 0, st>>>(z, zp, B, K, Bpad);
    int rc = check_launch("linear_prep_z");
    if (rc) return rc;
    if (C == 64) return launch_linear<512, 64>(zp, wb, scale, leak, y, r, B, Bpad, S * S, st);
    if (C == 128) return launch_linear<512, 128>(zp, wb, scale, leak, y, r, B, Bpad, S * S, st);
    return launch_linear<512, 256>(zp, wb, scale, leak, y, r, B, Bpad, S * S, st);
}

// dW[f][k] (+)= scale * sum_b ga[b][f] * z[b][k], f = c*S*S + p; ga is C8 [B][C/8][S*S][8].  With accumulate == 0
// the 67 MB gradient is written without being read (and the caller need not zero it first).
// One block per C8 granule (8 rows of dW that share p), 128 threads x 4 consecutive k.
template <int K>
__global__ void __launch_bounds__(K / 4) linear_wgrad_kernel(const uint4* __restrict__ ga, const float* __restrict__ z,
                                                             float scale, float* __restrict__ dw, int accumulate,
                                                             int B, int C, int SS) {
    __shared__ float sg[64][8];
    const int p = blockIdx.x, j = blockIdx.y;
    const int nch = C / 8;
    float acc[8][4];
#pragma unroll
    for (int e = 0; e < 8; ++e)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[e][q] = 0.f;
    for (int b0 = 0; b0 < B; b0 += 64) {
        const int nb = min(64, B - b0);
        __syncthreads();
        if (threadIdx.x < nb) {
            float v[8];
            unpack8(__ldg(ga + (static_cast<size_t>(b0 + threadIdx.x) * nch + j) * SS + p), v);
#pragma unroll
            for (int e = 0; e < 8; ++e) sg[threadIdx.x][e] = v[e];
        }
        __syncthreads();
        for (int bb = 0; bb < nb; ++bb) {
            const float4 zv = __ldg(reinterpret_cast<const float4*>(z + static_cast<size_t>(b0 + bb) * K) + threadIdx.x);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float gv = sg[bb][e];
                acc[e][0] = __fmaf_rn(gv, zv.x, acc[e][0]);
                acc[e][1] = __fmaf_rn(gv, zv.y, acc[e][1]);
                acc[e][2] = __fmaf_rn(gv, zv.z, acc[e][2]);
                acc[e][3] = __fmaf_rn(gv, zv.w, acc[e][3]);
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float4* d = reinterpret_cast<float4*>(dw + (static_cast<size_t>(j * 8 + e) * SS + p) * K) + threadIdx.x;
        float4 o = accumulate ? *d : make_float4(0.f, 0.f, 0.f, 0.f);
        // (explicit roundings: ngan_adam_linear_factored forms the same gradient in registers, bit for bit)
        o.x = __fadd_rn(o.x, __fmul_rn(scale, acc[e][0]));
        o.y = __fadd_rn(o.y, __fmul_rn(scale, acc[e][1]));
        o.z = __fadd_rn(o.z, __fmul_rn(scale, acc[e][2]));
        o.w = __fadd_rn(o.w, __fmul_rn(scale, acc[e][3]));
        *d = o;
    }
}
int linear_wgrad(const void* ga, const float* z, float scale, float* dw, int accumulate, int B, int K, int C, int S,
                 cudaStream_t st) {
    if (K != 512 || C % 8) {
        set_error("linear_wgrad: only latent_dim 512 and C %% 8 == 0 are built (got K=%d C=%d)", K, C);
        return NGAN_ERR_UNSUPPORTED;
    }
    dim3 grid(S * S, C / 8);
    linear_wgrad_kernel<512><<<grid, 128, 0, st>>>(static_cast<const uint4*>(ga), z, scale, dw, accumulate, B, C,
                                                   S * S);
    return check_launch("linear_wgrad");
}

}  // namespace ngan
