// Generator stem: Linear_normalized(latent -> C*S*S) + Unflatten + LeakyReLU + PixelNorm
// (reference models.py:299-311, 240-241) and its weight gradient.
//
// FLOPs are negligible (2*B*16.8 MFLOP); both kernels are bound by streaming the 16.8 M-element weight
// (bf16 shadow, 33.5 MB, forward) or its fp32 gradient (67 MB, backward) through HBM once.
#include "common.cuh"
#include "kernels.h"

namespace ngan {

__global__ void f32_to_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, size_t n) {
    size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    if (i + 3 < n) {
        float4 v = *reinterpret_cast<const float4*>(w + i);
        uint2 o;
        o.x = pack_bf16(v.x, v.y);
        o.y = pack_bf16(v.z, v.w);
        *reinterpret_cast<uint2*>(out + i) = o;
    } else {
        for (; i < n; ++i) out[i] = __float2bfloat16(w[i]);
    }
}
int prep_linear_weight(const float* w, void* wb, size_t n, cudaStream_t st) {
    const size_t threads = (n + 3) / 4;
    f32_to_bf16_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, st>>>(w, static_cast<__nv_bfloat16*>(wb), n);
    return check_launch("prep_linear_weight");
}

// One block per (pixel p, chunk of 16 samples).  Weight rows f = c*S*S + p, c = 0..C-1, are the C channels of
// pixel p, so the PixelNorm reduction over channels stays inside the block.  Each warp streams whole weight
// rows (coalesced 16-byte loads), multiplies against the 16 latent vectors held in shared memory and
// warp-reduces; the [C][16] pre-activations are then normalised and written as C8 granules.
constexpr int kLinBT = 16;
template <int K>
__global__ void __launch_bounds__(256) linear_fwd_pn_kernel(const float* __restrict__ z,
                                                            const __nv_bfloat16* __restrict__ wb, float scale,
                                                            float leak, uint4* __restrict__ y, float* __restrict__ r,
                                                            int B, int C, int SS) {
    extern __shared__ float sm[];
    float* sz = sm;               // [kLinBT][K]
    float* sa = sm + kLinBT * K;  // [C][kLinBT]
    float* sr = sa + C * kLinBT;  // [kLinBT]
    const int p = blockIdx.x, b0 = blockIdx.y * kLinBT;
    const int nb = min(kLinBT, B - b0);
    // sz[bb][j][lane] = scale * z[b0+bb][lane*KPL + j]: lane-interleaved so the inner product below reads
    // consecutive banks (a plain [bb][k] layout makes the 32 lanes stride 16 floats: 16-way bank conflicts)
    for (int i = threadIdx.x; i < kLinBT * K; i += blockDim.x) {
        const int bb = i / K, k = i % K;
        const float v = bb < nb ? z[static_cast<size_t>(b0 + bb) * K + k] * scale : 0.f;
        sz[bb * K + (k % (K / 32)) * 32 + k / (K / 32)] = v;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KPL = K / 32;  // k's per lane (16 for K = 512)
    for (int c = warp; c < C; c += 8) {
        const __nv_bfloat16* row = wb + (static_cast<size_t>(c) * SS + p) * K + lane * KPL;
        float wv[KPL];
#pragma unroll
        for (int q = 0; q < KPL / 8; ++q) unpack8(__ldg(reinterpret_cast<const uint4*>(row) + q), wv + q * 8);
#pragma unroll 4
        for (int bb = 0; bb < kLinBT; ++bb) {
            const float* zz = sz + bb * K + lane;
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < KPL; ++k) acc += wv[k] * zz[k * 32];
            acc = warp_sum(acc);
            if (lane == 0) sa[c * kLinBT + bb] = acc > 0.f ? acc : leak * acc;
        }
    }
    __syncthreads();
    if (threadIdx.x < kLinBT) {
        float ss = 0.f;
        for (int c = 0; c < C; ++c) ss += sa[c * kLinBT + threadIdx.x] * sa[c * kLinBT + threadIdx.x];
        sr[threadIdx.x] = rsqrtf(ss / C + 1e-8f);
    }
    __syncthreads();
    const int nch = C / 8;
    for (int i = threadIdx.x; i < nb * nch; i += blockDim.x) {
        const int bb = i / nch, j = i % nch;
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = sa[(j * 8 + e) * kLinBT + bb] * sr[bb];
        y[(static_cast<size_t>(b0 + bb) * nch + j) * SS + p] = pack8(o);
        if (j == 0 && r) r[static_cast<size_t>(b0 + bb) * SS + p] = sr[bb];
    }
}
int linear_fwd_pn(const float* z, const void* wb, float scale, float leak, void* y, float* r, int B, int K, int C,
                  int S, cudaStream_t st) {
    if (K != 512 || C % 8) {
        set_error("linear_fwd_pn: only latent_dim 512 and C %% 8 == 0 are built (got K=%d C=%d)", K, C);
        return NGAN_ERR_UNSUPPORTED;
    }
    const size_t smem = (kLinBT * 512 + C * kLinBT + kLinBT) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(linear_fwd_pn_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             96 * 1024);
        if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(linear)");
        configured = true;
    }
    dim3 grid(S * S, (B + kLinBT - 1) / kLinBT);
    linear_fwd_pn_kernel<512><<<grid, 256, smem, st>>>(z, static_cast<const __nv_bfloat16*>(wb), scale, leak,
                                                        static_cast<uint4*>(y), r, B, C, S * S);
    return check_launch("linear_fwd_pn");
}

// dW[f][k] += scale * sum_b ga[b][f] * z[b][k], f = c*S*S + p; ga is C8 [B][C/8][S*S][8].
// One block per C8 granule (8 rows of dW that share p), 128 threads x 4 consecutive k.
template <int K>
__global__ void __launch_bounds__(K / 4) linear_wgrad_kernel(const uint4* __restrict__ ga, const float* __restrict__ z,
                                                             float scale, float* __restrict__ dw, int B, int C,
                                                             int SS) {
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
                acc[e][0] += gv * zv.x;
                acc[e][1] += gv * zv.y;
                acc[e][2] += gv * zv.z;
                acc[e][3] += gv * zv.w;
            }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float4* d = reinterpret_cast<float4*>(dw + (static_cast<size_t>(j * 8 + e) * SS + p) * K) + threadIdx.x;
        float4 o = *d;
        o.x += scale * acc[e][0];
        o.y += scale * acc[e][1];
        o.z += scale * acc[e][2];
        o.w += scale * acc[e][3];
        *d = o;
    }
}
int linear_wgrad(const void* ga, const float* z, float scale, float* dw, int B, int K, int C, int S,
                 cudaStream_t st) {
    if (K != 512 || C % 8) {
        set_error("linear_wgrad: only latent_dim 512 and C %% 8 == 0 are built (got K=%d C=%d)", K, C);
        return NGAN_ERR_UNSUPPORTED;
    }
    dim3 grid(S * S, C / 8);
    linear_wgrad_kernel<512><<<grid, 128, 0, st>>>(static_cast<const uint4*>(ga), z, scale, dw, B, C, S * S);
    return check_launch("linear_wgrad");
}

}  // namespace ngan
