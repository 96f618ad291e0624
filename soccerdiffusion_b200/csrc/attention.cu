// Multi-head attention core (softmax(Q K^T / sqrt(dh)) V), forward and backward, fp32.
//
// Semantics: torch.nn.MultiheadAttention / F.scaled_dot_product_attention as used by
// nn.TransformerEncoderLayer / nn.TransformerDecoderLayer in the reference (no masks, non-causal,
// attention-probability dropout in training; torch/nn/functional.py:6623-6690), call sites
// ml/model/encoder/base.py:30-40 (self, S<=100 tokens), ml/model/decoder.py:24-35 (self T=10,
// cross T=10 x M=312).
//
// One CTA per (sample, head).  Keys/values stream through shared memory in chunks of MC rows with
// an online softmax, so any memory length fits; queries are handled warp-per-row with the
// running (max, sum, O) state in shared memory.  The backward pass recomputes P from the saved
// log-sum-exp, keeps dK/dV accumulators in registers (warp-per-key) and never materialises the
// T x M matrices in HBM.
#include "common.cuh"
#include "../../include/sd_b200.h"

using namespace sd;

namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int MC = 64;   // key chunk
constexpr int TC = 32;   // query chunk (backward)

struct AttnParams {
    const float* Q; long long ldq;   // row (b*T + t), head slice [h*dh, (h+1)*dh)
    const float* K; long long ldk;   // row (b*M + m)
    const float* V; long long ldv;
    float* O; long long ldo;
    float* lse;                      // (B, H, T)
    int B, H, T, M, dh;
    float scale;
    Dropout drop;
};

template <int NV>
__global__ void __launch_bounds__(kThreads) attn_fwd_kernel(const AttnParams p) {
    constexpr int DH = NV * 32;
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;                          // [MC][DH+1]
    float* Vs = Ks + MC * (DH + 1);            // [MC][DH]
    float* Qs = Vs + MC * DH;                  // [T][DH]
    float* Os = Qs + (size_t)p.T * DH;         // [T][DH]
    float* mx = Os + (size_t)p.T * DH;         // [T]
    float* sm = mx + p.T;                      // [T]
    float* sc = sm + p.T;                      // [kWarps][MC]

    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* Qg = p.Q + (long long)b * p.T * p.ldq + h * DH;
    const float* Kg = p.K + (long long)b * p.M * p.ldk + h * DH;
    const float* Vg = p.V + (long long)b * p.M * p.ldv + h * DH;

    for (int i = threadIdx.x; i < p.T * DH; i += kThreads) {
        const int t = i / DH, c = i % DH;
        Qs[i] = Qg[(long long)t * p.ldq + c];
        Os[i] = 0.f;
    }
    for (int t = threadIdx.x; t < p.T; t += kThreads) {
        mx[t] = -INFINITY;
        sm[t] = 0.f;
    }

    for (int m0 = 0; m0 < p.M; m0 += MC) {
        const int mc = min(MC, p.M - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < mc * DH; i += kThreads) {
            const int m = i / DH, c = i % DH;
            Ks[m * (DH + 1) + c] = Kg[(long long)(m0 + m) * p.ldk + c];
            Vs[m * DH + c] = Vg[(long long)(m0 + m) * p.ldv + c];
        }
        __syncthreads();
        float* scw = sc + w * MC;
        for (int t = w; t < p.T; t += kWarps) {
            const float* q = Qs + (size_t)t * DH;
            float s[MC / 32];
            float cmax = -INFINITY;
#pragma unroll
            for (int j = 0; j < MC / 32; ++j) {
                const int m = lane + 32 * j;
                float a = 0.f;
                if (m < mc) {
                    const float* kr = Ks + m * (DH + 1);
#pragma unroll 8
                    for (int c = 0; c < DH; ++c) a = fmaf(q[c], kr[c], a);
                    a *= p.scale;
                    cmax = fmaxf(cmax, a);
                }
                s[j] = a;
            }
            cmax = warp_max(cmax);
            const float old = mx[t];
            const float nm = fmaxf(old, cmax);
            const float corr = (old == -INFINITY) ? 0.f : expf(old - nm);
            float lsum = 0.f;
            const uint64_t row_idx = ((uint64_t)(b * p.H + h) * p.T + t) * (uint64_t)p.M + m0;
#pragma unroll
            for (int j = 0; j < MC / 32; ++j) {
                const int m = lane + 32 * j;
                if (m < mc) {
                    const float e = expf(s[j] - nm);
                    lsum += e;
                    scw[m] = e * p.drop(row_idx + m);
                }
            }
            lsum = warp_sum(lsum);
            __syncwarp();
            float* o = Os + (size_t)t * DH;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c = lane + 32 * v;
                float a = o[c] * corr;
                for (int m = 0; m < mc; ++m) a = fmaf(scw[m], Vs[m * DH + c], a);
                o[c] = a;
            }
            if (lane == 0) {
                mx[t] = nm;
                sm[t] = sm[t] * corr + lsum;
            }
            __syncwarp();
        }
    }
    __syncthreads();
    float* Og = p.O + (long long)b * p.T * p.ldo + h * DH;
    for (int i = threadIdx.x; i < p.T * DH; i += kThreads) {
        const int t = i / DH, c = i % DH;
        Og[(long long)t * p.ldo + c] = Os[i] / sm[t];
    }
    if (p.lse)
        for (int t = threadIdx.x; t < p.T; t += kThreads)
            p.lse[((long long)b * p.H + h) * p.T + t] = mx[t] + logf(sm[t]);
}

struct AttnBwdParams {
    const float* Q; long long ldq;
    const float* K; long long ldk;
    const float* V; long long ldv;
    const float* O; long long ldo;
    const float* dO; long long lddo;
    const float* lse;
    float* dQ; long long lddq;
    float* dK; long long lddk;
    float* dV; long long lddv;
    int B, H, T, M, dh;
    float scale;
    Dropout drop;
};

template <int NV>
__global__ void __launch_bounds__(kThreads) attn_bwd_kernel(const AttnBwdParams p) {
    constexpr int DH = NV * 32;
    constexpr int KPW = MC / kWarps;  // keys per warp in phase B
    extern __shared__ __align__(16) float smem[];
    float* Ks = smem;                         // [MC][DH+1]
    float* Vs = Ks + MC * (DH + 1);           // [MC][DH+1]
    float* Qs = Vs + MC * (DH + 1);           // [TC][DH]
    float* dOs = Qs + TC * DH;                // [TC][DH]
    float* Pd = dOs + TC * DH;                // [TC][MC]   P * dropout
    float* dS = Pd + TC * MC;                 // [TC][MC]   dS * scale
    float* lses = dS + TC * MC;               // [T]
    float* delta = lses + p.T;                // [T]

    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* Qg = p.Q + (long long)b * p.T * p.ldq + h * DH;
    const float* Kg = p.K + (long long)b * p.M * p.ldk + h * DH;
    const float* Vg = p.V + (long long)b * p.M * p.ldv + h * DH;
    const float* Og = p.O + (long long)b * p.T * p.ldo + h * DH;
    const float* dOg = p.dO + (long long)b * p.T * p.lddo + h * DH;
    float* dQg = p.dQ + (long long)b * p.T * p.lddq + h * DH;
    float* dKg = p.dK + (long long)b * p.M * p.lddk + h * DH;
    float* dVg = p.dV + (long long)b * p.M * p.lddv + h * DH;

    // delta[t] = dO[t] . O[t]
    for (int t = w; t < p.T; t += kWarps) {
        float a = 0.f;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int c = lane + 32 * v;
            a = fmaf(dOg[(long long)t * p.lddo + c], Og[(long long)t * p.ldo + c], a);
        }
        a = warp_sum(a);
        if (lane == 0) {
            delta[t] = a;
            lses[t] = p.lse[((long long)b * p.H + h) * p.T + t];
        }
    }

    for (int m0 = 0; m0 < p.M; m0 += MC) {
        const int mc = min(MC, p.M - m0);
        __syncthreads();
        for (int i = threadIdx.x; i < mc * DH; i += kThreads) {
            const int m = i / DH, c = i % DH;
            Ks[m * (DH + 1) + c] = Kg[(long long)(m0 + m) * p.ldk + c];
            Vs[m * (DH + 1) + c] = Vg[(long long)(m0 + m) * p.ldv + c];
        }
        float dk[KPW][NV], dv[KPW][NV];
#pragma unroll
        for (int i = 0; i < KPW; ++i)
#pragma unroll
            for (int v = 0; v < NV; ++v) dk[i][v] = dv[i][v] = 0.f;

        for (int t0 = 0; t0 < p.T; t0 += TC) {
            const int tc = min(TC, p.T - t0);
            __syncthreads();
            for (int i = threadIdx.x; i < tc * DH; i += kThreads) {
                const int t = i / DH, c = i % DH;
                Qs[i] = Qg[(long long)(t0 + t) * p.ldq + c];
                dOs[i] = dOg[(long long)(t0 + t) * p.lddo + c];
            }
            __syncthreads();
            // phase A: rows of the query chunk, warp per row
            for (int t = w; t < tc; t += kWarps) {
                const float* q = Qs + t * DH;
                const float* go = dOs + t * DH;
                const float l = lses[t0 + t], dl = delta[t0 + t];
                const uint64_t row_idx = ((uint64_t)(b * p.H + h) * p.T + (t0 + t)) * (uint64_t)p.M + m0;
#pragma unroll
                for (int j = 0; j < MC / 32; ++j) {
                    const int m = lane + 32 * j;
                    if (m < mc) {
                        const float* kr = Ks + m * (DH + 1);
                        const float* vr = Vs + m * (DH + 1);
                        float s = 0.f, dpd = 0.f;
#pragma unroll 8
                        for (int c = 0; c < DH; ++c) {
                            s = fmaf(q[c], kr[c], s);
                            dpd = fmaf(go[c], vr[c], dpd);
                        }
                        const float pr = expf(s * p.scale - l);
                        const float dm = p.drop(row_idx + m);
                        Pd[t * MC + m] = pr * dm;
                        dS[t * MC + m] = pr * (dpd * dm - dl) * p.scale;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int c = lane + 32 * v;
                    float a = 0.f;
                    for (int m = 0; m < mc; ++m) a = fmaf(dS[t * MC + m], Ks[m * (DH + 1) + c], a);
                    float* dst = &dQg[(long long)(t0 + t) * p.lddq + c];
                    *dst = (m0 == 0) ? a : *dst + a;
                }
            }
            __syncthreads();
            // phase B: keys of the chunk, warp per key, accumulate over the query chunk
#pragma unroll
            for (int i = 0; i < KPW; ++i) {
                const int m = w + kWarps * i;
                if (m < mc) {
                    for (int t = 0; t < tc; ++t) {
                        const float ds = dS[t * MC + m], pd = Pd[t * MC + m];
#pragma unroll
                        for (int v = 0; v < NV; ++v) {
                            const int c = lane + 32 * v;
                            dk[i][v] = fmaf(ds, Qs[t * DH + c], dk[i][v]);
                            dv[i][v] = fmaf(pd, dOs[t * DH + c], dv[i][v]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < KPW; ++i) {
            const int m = w + kWarps * i;
            if (m < mc) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int c = lane + 32 * v;
                    dKg[(long long)(m0 + m) * p.lddk + c] = dk[i][v];
                    dVg[(long long)(m0 + m) * p.lddv + c] = dv[i][v];
                }
            }
        }
    }
}

// dh < 32 (image-sequence encoder: 8 heads of 16) is handled by the NV=1 kernels with lanes >= dh
// masked: simplest is a dedicated small-head path that pads dh to 32 in shared memory.  To keep one
// code path, small heads go through a "half-warp" specialisation below.
template <int DHS>
__global__ void __launch_bounds__(kThreads) attn_fwd_small_kernel(const AttnParams p) {
    // DHS in {8,16}: whole (T x M) problem is tiny (image sequence: T = M <= 32 frames).  One warp
    // per query row, lanes over keys, direct global reads (L1 resident), no chunking.
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* Qg = p.Q + (long long)b * p.T * p.ldq + h * DHS;
    const float* Kg = p.K + (long long)b * p.M * p.ldk + h * DHS;
    const float* Vg = p.V + (long long)b * p.M * p.ldv + h * DHS;
    float* Og = p.O + (long long)b * p.T * p.ldo + h * DHS;
    extern __shared__ __align__(16) float smem[];
    float* sc = smem + w * p.M;  // [kWarps][M]
    for (int t = w; t < p.T; t += kWarps) {
        float q[DHS];
#pragma unroll
        for (int c = 0; c < DHS; ++c) q[c] = Qg[(long long)t * p.ldq + c];
        float mxv = -INFINITY;
        for (int m = lane; m < p.M; m += 32) {
            float a = 0.f;
#pragma unroll
            for (int c = 0; c < DHS; ++c) a = fmaf(q[c], Kg[(long long)m * p.ldk + c], a);
            a *= p.scale;
            sc[m] = a;
            mxv = fmaxf(mxv, a);
        }
        mxv = warp_max(mxv);
        float lsum = 0.f;
        const uint64_t row_idx = ((uint64_t)(b * p.H + h) * p.T + t) * (uint64_t)p.M;
        for (int m = lane; m < p.M; m += 32) {
            const float e = expf(sc[m] - mxv);
            lsum += e;
            sc[m] = e * p.drop(row_idx + m);
        }
        lsum = warp_sum(lsum);
        __syncwarp();
        if (lane < DHS) {
            float a = 0.f;
            for (int m = 0; m < p.M; ++m) a = fmaf(sc[m], Vg[(long long)m * p.ldv + lane], a);
            Og[(long long)t * p.ldo + lane] = a / lsum;
        }
        if (lane == 0 && p.lse) p.lse[((long long)b * p.H + h) * p.T + t] = mxv + logf(lsum);
        __syncwarp();
    }
}

template <int DHS>
__global__ void __launch_bounds__(kThreads) attn_bwd_small_kernel(const AttnBwdParams p) {
    // tiny problem: P and dS for the whole (T x M) block live in shared memory.
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* Qg = p.Q + (long long)b * p.T * p.ldq + h * DHS;
    const float* Kg = p.K + (long long)b * p.M * p.ldk + h * DHS;
    const float* Vg = p.V + (long long)b * p.M * p.ldv + h * DHS;
    const float* Og = p.O + (long long)b * p.T * p.ldo + h * DHS;
    const float* dOg = p.dO + (long long)b * p.T * p.lddo + h * DHS;
    float* dQg = p.dQ + (long long)b * p.T * p.lddq + h * DHS;
    float* dKg = p.dK + (long long)b * p.M * p.lddk + h * DHS;
    float* dVg = p.dV + (long long)b * p.M * p.lddv + h * DHS;
    extern __shared__ __align__(16) float smem[];
    float* Pd = smem;                      // [T][M]
    float* dS = Pd + (size_t)p.T * p.M;    // [T][M]
    for (int t = w; t < p.T; t += kWarps) {
        float q[DHS], go[DHS];
        float dl = 0.f;
#pragma unroll
        for (int c = 0; c < DHS; ++c) {
            q[c] = Qg[(long long)t * p.ldq + c];
            go[c] = dOg[(long long)t * p.lddo + c];
            dl = fmaf(go[c], Og[(long long)t * p.ldo + c], dl);
        }
        const float l = p.lse[((long long)b * p.H + h) * p.T + t];
        const uint64_t row_idx = ((uint64_t)(b * p.H + h) * p.T + t) * (uint64_t)p.M;
        for (int m = lane; m < p.M; m += 32) {
            float s = 0.f, dpd = 0.f;
#pragma unroll
            for (int c = 0; c < DHS; ++c) {
                s = fmaf(q[c], Kg[(long long)m * p.ldk + c], s);
                dpd = fmaf(go[c], Vg[(long long)m * p.ldv + c], dpd);
            }
            const float pr = expf(s * p.scale - l);
            const float dm = p.drop(row_idx + m);
            Pd[t * p.M + m] = pr * dm;
            dS[t * p.M + m] = pr * (dpd * dm - dl) * p.scale;
        }
        __syncwarp();
        if (lane < DHS) {
            float a = 0.f;
            for (int m = 0; m < p.M; ++m) a = fmaf(dS[t * p.M + m], Kg[(long long)m * p.ldk + lane], a);
            dQg[(long long)t * p.lddq + lane] = a;
        }
    }
    __syncthreads();
    for (int m = w; m < p.M; m += kWarps) {
        if (lane < DHS) {
            float ak = 0.f, av = 0.f;
            for (int t = 0; t < p.T; ++t) {
                ak = fmaf(dS[t * p.M + m], Qg[(long long)t * p.ldq + lane], ak);
                av = fmaf(Pd[t * p.M + m], dOg[(long long)t * p.lddo + lane], av);
            }
            dKg[(long long)m * p.lddk + lane] = ak;
            dVg[(long long)m * p.lddv + lane] = av;
        }
    }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) return SD_ERR_UNSUPPORTED;
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return SD_OK;
}

}  // namespace

extern "C" int sd_attention_fwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V,
                                long long ldv, float* O, long long ldo, float* lse, int B, int H, int T, int M, int dh,
                                float dropout_p, unsigned long long seed, unsigned int stream_id, void* stream) {
    if (B <= 0 || T <= 0) return SD_OK;
    if (!Q || !K || !V || !O || H <= 0 || M <= 0) return SD_ERR_BAD_ARG;
    AttnParams p{Q, ldq, K, ldk, V, ldv, O, ldo, lse, B, H, T, M, dh, 1.0f / sqrtf((float)dh),
                 make_dropout(dropout_p, seed, stream_id)};
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = B * H;
    int rc;
#define SD_FWD_BIG(NVv)                                                                                       \
    {                                                                                                         \
        constexpr int DH = NVv * 32;                                                                          \
        const size_t bytes = sizeof(float) * ((size_t)MC * (2 * DH + 1) + 2 * (size_t)T * DH + 2 * (size_t)T + kWarps * MC); \
        if ((rc = set_smem(attn_fwd_kernel<NVv>, bytes)) != SD_OK) return rc;                                 \
        attn_fwd_kernel<NVv><<<grid, kThreads, bytes, st>>>(p);                                               \
    }
    switch (dh) {
        case 4: {
            const size_t bytes = sizeof(float) * kWarps * (size_t)M;
            if ((rc = set_smem(attn_fwd_small_kernel<4>, bytes)) != SD_OK) return rc;
            attn_fwd_small_kernel<4><<<grid, kThreads, bytes, st>>>(p);
        } break;
        case 8: {
            const size_t bytes = sizeof(float) * kWarps * (size_t)M;
            if ((rc = set_smem(attn_fwd_small_kernel<8>, bytes)) != SD_OK) return rc;
            attn_fwd_small_kernel<8><<<grid, kThreads, bytes, st>>>(p);
        } break;
        case 16: {
            const size_t bytes = sizeof(float) * kWarps * (size_t)M;
            if ((rc = set_smem(attn_fwd_small_kernel<16>, bytes)) != SD_OK) return rc;
            attn_fwd_small_kernel<16><<<grid, kThreads, bytes, st>>>(p);
        } break;
        case 32: SD_FWD_BIG(1) break;
        case 64: SD_FWD_BIG(2) break;
        case 128: SD_FWD_BIG(4) break;
        default: return SD_ERR_UNSUPPORTED;
    }
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_attention_bwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V,
                                long long ldv, const float* O, long long ldo, const float* dO, long long lddo,
                                const float* lse, float* dQ, long long lddq, float* dK, long long lddk, float* dV,
                                long long lddv, int B, int H, int T, int M, int dh, float dropout_p,
                                unsigned long long seed, unsigned int stream_id, void* stream) {
    if (B <= 0 || T <= 0) return SD_OK;
    if (!Q || !K || !V || !O || !dO || !lse || !dQ || !dK || !dV || H <= 0 || M <= 0) return SD_ERR_BAD_ARG;
    AttnBwdParams p{Q, ldq, K, ldk, V, ldv, O, ldo, dO, lddo, lse, dQ, lddq, dK, lddk, dV, lddv,
                    B, H, T, M, dh, 1.0f / sqrtf((float)dh), make_dropout(dropout_p, seed, stream_id)};
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = B * H;
    int rc;
#define SD_BWD_BIG(NVv)                                                                                        \
    {                                                                                                          \
        constexpr int DH = NVv * 32;                                                                           \
        const size_t bytes = sizeof(float) * (2 * (size_t)MC * (DH + 1) + 2 * (size_t)TC * DH + 2 * (size_t)TC * MC + 2 * (size_t)T); \
        if ((rc = set_smem(attn_bwd_kernel<NVv>, bytes)) != SD_OK) return rc;                                  \
        attn_bwd_kernel<NVv><<<grid, kThreads, bytes, st>>>(p);                                                \
    }
    switch (dh) {
        case 4: {
            const size_t bytes = sizeof(float) * 2 * (size_t)T * M;
            if ((rc = set_smem(attn_bwd_small_kernel<4>, bytes)) != SD_OK) return rc;
            attn_bwd_small_kernel<4><<<grid, kThreads, bytes, st>>>(p);
        } break;
        case 8: {
            const size_t bytes = sizeof(float) * 2 * (size_t)T * M;
            if ((rc = set_smem(attn_bwd_small_kernel<8>, bytes)) != SD_OK) return rc;
            attn_bwd_small_kernel<8><<<grid, kThreads, bytes, st>>>(p);
        } break;
        case 16: {
            const size_t bytes = sizeof(float) * 2 * (size_t)T * M;
            if ((rc = set_smem(attn_bwd_small_kernel<16>, bytes)) != SD_OK) return rc;
            attn_bwd_small_kernel<16><<<grid, kThreads, bytes, st>>>(p);
        } break;
        case 32: SD_BWD_BIG(1) break;
        case 64: SD_BWD_BIG(2) break;
        case 128: SD_BWD_BIG(4) break;
        default: return SD_ERR_UNSUPPORTED;
    }
    SD_LAUNCH_CHECK();
    return SD_OK;
}
