// LayerNorm statistics and backward (warp-per-row, shuffle reductions, eps = 1e-5 like
// nn.LayerNorm inside nn.Transformer{En,De}coderLayer; torch/nn/modules/transformer.py:944-950).
// The normalisation itself is applied on the fly inside the GEMM operand loads (gemm_f32.cu /
// gemm_tc.cu); only the per-row mean / rstd are materialised (8 B per token instead of 4*d B).
#include "common.cuh"
#include "../../include/sd_b200.h"

using namespace sd;

namespace {

template <int NV>
__global__ void __launch_bounds__(256) ln_stats_kernel(const float* __restrict__ x, long long ld, long long M,
                                                       float* __restrict__ mean, float* __restrict__ rstd, float eps) {
    constexpr int D = NV * 32;
    const int lane = threadIdx.x & 31;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < M; r += nwarps) {
        float v[NV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            v[i] = x[r * ld + lane + 32 * i];
            s += v[i];
        }
        const float mu = warp_sum(s) * (1.0f / D);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float dlt = v[i] - mu;
            q = fmaf(dlt, dlt, q);
        }
        const float var = warp_sum(q) * (1.0f / D);
        if (lane == 0) {
            mean[r] = mu;
            rstd[r] = rsqrtf(var + eps);
        }
    }
}

// dx = dres + rstd * (g*gamma - mean_c(g*gamma) - xhat * mean_c(g*gamma*xhat));
// dgamma += sum_m g*xhat ; dbeta += sum_m g
template <int NV>
__global__ void __launch_bounds__(256, NV <= 4 ? 2 : 1) ln_bwd_kernel(const float* __restrict__ g, long long ldg,
                                                     const float* __restrict__ x, long long ldx,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ dres,
                                                     long long ldres, float* __restrict__ dx, long long lddx,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                     long long M) {
    constexpr int D = NV * 32;
    __shared__ float red_g[8][D];
    __shared__ float red_b[8][D];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float gam[NV], ag[NV], ab[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        gam[i] = gamma[lane + 32 * i];
        ag[i] = 0.f;
        ab[i] = 0.f;
    }
    // RB rows per warp iteration: every global load of the batch (g, x, the residual-path gradient, mean, rstd) is issued
    // before the first warp reduction, so a batch costs one DRAM round trip instead of two per row
    constexpr int RB = NV <= 4 ? 4 : 2;
    for (long long r0 = warp * RB; r0 < M; r0 += nwarps * RB) {
        float gv[RB][NV], xv[RB][NV], dr[RB][NV], mu[RB], rs[RB];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const long long r = r0 + b;
            if (r < M) {
                mu[b] = mean[r];
                rs[b] = rstd[r];
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    const int c = lane + 32 * i;
                    gv[b][i] = g[r * ldg + c];
                    xv[b][i] = x[r * ldx + c];
                    dr[b][i] = dres ? dres[r * ldres + c] : 0.f;
                }
            }
        }
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            const long long r = r0 + b;
            if (r >= M) break;
            float xh[NV];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                xh[i] = (xv[b][i] - mu[b]) * rs[b];
                const float gg = gv[b][i] * gam[i];
                s1 += gg;
                s2 = fmaf(gg, xh[i], s2);
                ag[i] = fmaf(gv[b][i], xh[i], ag[i]);
                ab[i] += gv[b][i];
            }
            s1 = warp_sum(s1) * (1.0f / D);
            s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = lane + 32 * i;
                dx[r * lddx + c] = rs[b] * (gv[b][i] * gam[i] - s1 - xh[i] * s2) + dr[b][i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        red_g[wib][lane + 32 * i] = ag[i];
        red_b[wib][lane + 32 * i] = ab[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float sg = 0.f, sb = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            sg += red_g[w][c];
            sb += red_b[w][c];
        }
        atomicAdd(&dgamma[c], sg);
        atomicAdd(&dbeta[c], sb);
    }
}

}  // namespace

#define SD_DISPATCH_NV(d, CALL)                    \
    switch (d) {                                   \
        case 32: { constexpr int NV = 1; CALL; } break;   \
        case 64: { constexpr int NV = 2; CALL; } break;   \
        case 128: { constexpr int NV = 4; CALL; } break;  \
        case 256: { constexpr int NV = 8; CALL; } break;  \
        case 512: { constexpr int NV = 16; CALL; } break; \
        default: return SD_ERR_UNSUPPORTED;        \
    }

extern "C" int sd_ln_stats(const float* x, long long ld, long long M, int d, float* mean, float* rstd, float eps,
                           void* stream) {
    if (M <= 0) return SD_OK;
    if (!x || !mean || !rstd) return SD_ERR_BAD_ARG;
    const int blocks = (int)min((long long)148 * 8, (M + 7) / 8);
    cudaStream_t st = (cudaStream_t)stream;
    SD_DISPATCH_NV(d, (ln_stats_kernel<NV><<<blocks, 256, 0, st>>>(x, ld, M, mean, rstd, eps)));
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_ln_bwd(const float* g, long long ldg, const float* x, long long ldx, const float* mean,
                         const float* rstd, const float* gamma, const float* dres, long long ldres, float* dx,
                         long long lddx, float* dgamma, float* dbeta, long long M, int d, void* stream) {
    if (M <= 0) return SD_OK;
    if (!g || !x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta) return SD_ERR_BAD_ARG;
    // 2 resident CTAs per SM (the row batches live in registers); rows are handed out in batches of 4 (2 for d > 128)
    const int blocks = (int)min((long long)148 * 2, (M + 31) / 32);
    cudaStream_t st = (cudaStream_t)stream;
    SD_DISPATCH_NV(d, (ln_bwd_kernel<NV><<<blocks, 256, 0, st>>>(g, ldg, x, ldx, mean, rstd, gamma, dres, ldres, dx,
                                                                 lddx, dgamma, dbeta, M)));
    SD_LAUNCH_CHECK();
    return SD_OK;
}
