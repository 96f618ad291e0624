// Persistent fused DDIM sampler: ONE kernel launch runs all N denoising steps of a trajectory.
//
// Replaces the reference's Python loop (ml/inference/ros.py:301-310, ml/training/distill.py:179-189,
// ml/inference/plot.py:124-131):
//     for t in scheduler.timesteps:
//         eps = model.forward_with_context(ctx, x, t)        # ml/model/model.py:159-179
//         x   = scheduler.step(eps, t, x).prev_sample        # diffusers DDIMScheduler.step, eta=0
// with one CTA per trajectory that keeps x, the residual stream and all activations in shared
// memory for the whole loop (trajectories are independent: no inter-CTA communication, so the
// step loop lives inside the kernel and the launch/sync cost is paid once per trajectory instead
// of ~60 launches x 30 steps).
//
// Algorithmic restructuring that keeps results identical (fp32, same operation order per dot
// product up to summation order):
//   * the cross-attention K/V projections of the 311 step-invariant context rows are computed once
//     per trajectory (sd_plan_set_context) — the reference recomputes them in each of the 4 layers
//     at every step (83 % of its per-step FLOPs);
//   * the K/V rows of the step token (ml/model/misc.py:25-35) depend only on (t, weights): they are
//     tabulated for the whole schedule by sd_plan_set_schedule;
//   * the DDIM coefficients come from a device table (no per-step H2D scalar, cf. ros.py:306).
//
// Decoder layer semantics: nn.TransformerDecoderLayer(norm_first=True, activation="gelu",
// dim_feedforward=d) — torch/nn/modules/transformer.py:1131-1143; in/out projections
// ml/model/decoder.py:47-54.
#include "common.cuh"
#include "../../include/sd_b200.h"

#include <cooperative_groups.h>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace cg = cooperative_groups;

using namespace sd;

extern "C" int sd_gemm(const sd_gemm_desc* d, void* stream);
extern "C" int sd_step_token(const void* t, int t_is_float, const float* freqs, const float* token, float* out,
                             long long ld_out, int B, int d, void* stream);

namespace {

constexpr int kSamplerThreads = 512;
constexpr int kMaxLayers = 32;

struct LayerPtrs {   // all K-major ("transposed") fp32, device pointers into the plan blob
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *ln3_g, *ln3_b;
    const float *sa_wqkv_t, *sa_bqkv, *sa_wo_t, *sa_bo;
    const float *ca_wq_t, *ca_bq, *ca_wkv_t, *ca_bkv, *ca_wo_t, *ca_bo;
    const float *w1_t, *b1, *w2_t, *b2;
};

struct IoPtrs {
    const float *emb_wt, *emb_b, *pe, *fc_wt, *fc_b, *freqs, *token, *mean, *stdv;
};

struct SamplerArgs {
    const LayerPtrs* layers;
    IoPtrs io;
    int L, d, heads, dh, T, J, M, Mpad, B;
    float* Kt;            // [L][B][heads][dh][Mpad]
    float* Vc;            // [L][B][Mpad][d]
    const float* tok_kv;  // [S][L][2d]   (tok_mode 0)
    const float* coef;    // [S][4]  sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt(1-a_prev)
    int num_steps;
    int tok_mode;         // 0: schedule tables, 1: per-sample t given (single forward_with_context)
    const void* t_ptr;
    int t_is_float;
    const float* x_in;    // (B,T,J)
    float* x_out;         // (B,T,J) final sample (tok_mode 0)
    float* eps_out;       // tok_mode 1: (B,T,J);  tok_mode 0: optional trace (S,B,T,J)
    int denorm;
    long long* dbg;       // optional: clock64() stamps of the cluster kernel's phases (CTA 0, thread 0), 96 per step
};

// acc[t] += sum_{k0 <= k < k1} W[k] * x[k][t] for one output column: the weights come from L2 (__ldcg), 16 loads in flight
// per thread before the first use (the loop is L2-latency-bound: with 4 in flight a 128-deep column cost ~10 us), the
// activations are shared-memory broadcasts.
template <int TR>
__device__ __forceinline__ void cta_dot_column(float (&acc)[TR], const float* __restrict__ wp, int ldw,
                                               const float* __restrict__ xp, int k0, int k1) {
    int k = k0;
    for (; k + 16 <= k1; k += 16) {
        float w[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) w[j] = __ldcg(wp + (long long)(k + j) * ldw);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
#pragma unroll
            for (int q = 0; q < TR / 4; ++q) {
                const float4 x4 = *reinterpret_cast<const float4*>(xp + (k + j) * TR + 4 * q);
                acc[4 * q + 0] = fmaf(w[j], x4.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(w[j], x4.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(w[j], x4.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(w[j], x4.w, acc[4 * q + 3]);
            }
        }
    }
#pragma unroll 4
    for (; k < k1; ++k) {
        const float w = __ldcg(wp + (long long)k * ldw);
#pragma unroll
        for (int q = 0; q < TR / 4; ++q) {
            const float4 x4 = *reinterpret_cast<const float4*>(xp + k * TR + 4 * q);
            acc[4 * q + 0] = fmaf(w, x4.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(w, x4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w, x4.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(w, x4.w, acc[4 * q + 3]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// CTA-wide skinny GEMM:  out[t][i] = sum_k xT[b(i)][k][t] * Wt[b(i)][k][n(i)],  t < TR rows kept in
// registers, one output column per thread, K split across thread groups when columns are scarce.
template <int TR, class Epi>
__device__ __forceinline__ void cta_gemm(const float* Wt, int ldw, long long w_bstride, const float* xT,
                                         int x_bstride, int nb, int N, int K, int T, float* scr, Epi epi) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int items = nb * N;
    int G = 1;
    if (items * 2 <= nthreads) G = min(nthreads / items, K);
    if (G > 1) {
        const int i = tid % items, g = tid / items;
        if (g < G) {
            const int ks = (K + G - 1) / G;
            const int k0 = g * ks, k1 = min(K, k0 + ks);
            const int bidx = i / N, n = i % N;
            const float* wp = Wt + bidx * w_bstride + n;
            const float* xp = xT + bidx * x_bstride;
            float acc[TR];
#pragma unroll
            for (int t = 0; t < TR; ++t) acc[t] = 0.f;
            cta_dot_column<TR>(acc, wp, ldw, xp, k0, k1);
#pragma unroll
            for (int t = 0; t < TR; ++t) scr[(g * TR + t) * items + i] = acc[t];
        }
        __syncthreads();
        for (int idx = tid; idx < items * T; idx += nthreads) {
            const int t = idx / items, i = idx % items;
            float v = 0.f;
            for (int g2 = 0; g2 < G; ++g2) v += scr[(g2 * TR + t) * items + i];
            epi(i / N, i % N, t, v);
        }
    } else {
        for (int i = tid; i < items; i += nthreads) {
            const int bidx = i / N, n = i % N;
            const float* wp = Wt + bidx * w_bstride + n;
            const float* xp = xT + bidx * x_bstride;
            float acc[TR];
#pragma unroll
            for (int t = 0; t < TR; ++t) acc[t] = 0.f;
            cta_dot_column<TR>(acc, wp, ldw, xp, 0, K);
#pragma unroll
            for (int t = 0; t < TR; ++t)
                if (t < T) epi(bidx, n, t, acc[t]);
        }
    }
    __syncthreads();
}

// LayerNorm of the T rows of h (row-major [T][d]) written transposed into xT[c][t]
template <int TR>
__device__ __forceinline__ void cta_layernorm_T(const float* h, int d, int T, const float* gamma, const float* beta,
                                                float* xT) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = w; t < T; t += nw) {
        const float* r = h + t * d;
        float s = 0.f;
        for (int c = lane; c < d; c += 32) s += r[c];
        const float mu = warp_sum(s) / (float)d;
        float q = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float dl = r[c] - mu;
            q = fmaf(dl, dl, q);
        }
        const float rs = rsqrtf(warp_sum(q) / (float)d + 1e-5f);
        for (int c = lane; c < d; c += 32) xT[c * TR + t] = (r[c] - mu) * rs * __ldg(gamma + c) + __ldg(beta + c);
    }
    __syncthreads();
}

template <int TR>
__global__ void __launch_bounds__(kSamplerThreads, 1) sampler_kernel(const SamplerArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int d = a.d, T = a.T, J = a.J, H = a.heads, dh = a.dh, M = a.M, Mpad = a.Mpad;
    const int ldq = 3 * d + 1;
    const int Jp = (J + 3) & ~3;
    float* xs = smem;                         // [TR][Jp]
    float* eps = xs + TR * Jp;                // [TR][Jp]
    float* h = eps + TR * Jp;                 // [TR][d]
    float* xT = h + TR * d;                   // [d][TR]     GEMM input staging (transposed)
    float* big = xT + d * TR;                 // [TR][3d+1]  qkv | qT [d][TR] | ffT [d][TR]
    float* PT = big + ((TR * ldq + 3) & ~3);  // [H][Mpad][TR]
    float* scr = PT + H * Mpad * TR;          // [blockDim][TR]

    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float scale = rsqrtf((float)dh);

    for (int i = tid; i < T * J; i += blockDim.x) xs[(i / J) * Jp + (i % J)] = a.x_in[(long long)b * T * J + i];
    // zero the staging buffer once so that padded rows t >= T never hold NaNs
    for (int i = tid; i < d * TR; i += blockDim.x) xT[i] = 0.f;
    __syncthreads();

    for (int s = 0; s < a.num_steps; ++s) {
        // ---- embedding + positional encoding (decoder.py:48-50) -------------------------
        for (int i = tid; i < T * J; i += blockDim.x) xT[(i % J) * TR + (i / J)] = xs[(i / J) * Jp + (i % J)];
        __syncthreads();
        cta_gemm<TR>(a.io.emb_wt, d, 0, xT, 0, 1, d, J, T, scr, [&](int, int n, int t, float v) {
            h[t * d + n] = v + __ldg(a.io.emb_b + n) + __ldg(a.io.pe + t * d + n);
        });

        for (int l = 0; l < a.L; ++l) {
            const LayerPtrs& P = a.layers[l];
            // ---- self-attention block: h += Wo * SA(LN1 h) ------------------------------
            cta_layernorm_T<TR>(h, d, T, P.ln1_g, P.ln1_b, xT);
            cta_gemm<TR>(P.sa_wqkv_t, 3 * d, 0, xT, 0, 1, 3 * d, d, T, scr,
                         [&](int, int n, int t, float v) { big[t * ldq + n] = v + __ldg(P.sa_bqkv + n); });
            for (int pair = warp; pair < H * T; pair += nwarps) {
                const int hh = pair / T, t = pair % T;
                const float* q = big + t * ldq + hh * dh;
                float sc = -INFINITY;
                if (lane < T) {
                    const float* k = big + lane * ldq + d + hh * dh;
                    float acc = 0.f;
                    for (int c = 0; c < dh; ++c) acc = fmaf(q[c], k[c], acc);
                    sc = acc * scale;
                }
                const float mx = warp_max(sc);
                const float e = lane < T ? expf(sc - mx) : 0.f;
                const float p = e / warp_sum(e);
                for (int c0 = 0; c0 < dh; c0 += 32) {
                    const int c = c0 + lane;
                    float o = 0.f;
                    for (int m = 0; m < T; ++m) {
                        const float pm = __shfl_sync(0xffffffffu, p, m);
                        if (c < dh) o = fmaf(pm, big[m * ldq + 2 * d + hh * dh + c], o);
                    }
                    if (c < dh) xT[(hh * dh + c) * TR + t] = o;
                }
            }
            __syncthreads();
            cta_gemm<TR>(P.sa_wo_t, d, 0, xT, 0, 1, d, d, T, scr,
                         [&](int, int n, int t, float v) { h[t * d + n] += v + __ldg(P.sa_bo + n); });

            // ---- cross-attention block: h += Wo * CA(LN2 h, mem) ------------------------
            cta_layernorm_T<TR>(h, d, T, P.ln2_g, P.ln2_b, xT);
            float* qT = big;  // [d][TR], pre-scaled by 1/sqrt(dh)
            cta_gemm<TR>(P.ca_wq_t, d, 0, xT, 0, 1, d, d, T, scr,
                         [&](int, int n, int t, float v) { qT[n * TR + t] = (v + __ldg(P.ca_bq + n)) * scale; });
            float* Kt = a.Kt + ((long long)l * a.B + b) * (long long)H * dh * Mpad;
            float* Vc = a.Vc + ((long long)l * a.B + b) * (long long)Mpad * d;
            // step-token K/V row (memory row M-1)
            if (a.tok_mode == 0) {
                const float* tk = a.tok_kv + ((long long)s * a.L + l) * 2 * d;
                for (int n = tid; n < 2 * d; n += blockDim.x) {
                    const float v = __ldg(tk + n);
                    if (n < d) Kt[(long long)n * Mpad + (M - 1)] = v;   // n = hh*dh + c
                    else Vc[(long long)(M - 1) * d + (n - d)] = v;
                }
                __syncthreads();
            } else {
                // token = [sin(t f) | cos(t f) | learned] staged as a 1-row GEMM input
                float* tokT = xT;  // [d][TR], only t=0 used (xT is free between the q GEMM and P.V)
                const float tf = a.t_is_float ? ((const float*)a.t_ptr)[b] : (float)((const long long*)a.t_ptr)[b];
                const int half = d / 4;
                for (int c = tid; c < d; c += blockDim.x) {
                    float v;
                    if (c < half) v = sinf(tf * __ldg(a.io.freqs + c));
                    else if (c < 2 * half) v = cosf(tf * __ldg(a.io.freqs + c - half));
                    else v = __ldg(a.io.token + c - 2 * half);
                    for (int t = 0; t < TR; ++t) tokT[c * TR + t] = t == 0 ? v : 0.f;
                }
                __syncthreads();
                cta_gemm<TR>(P.ca_wkv_t, 2 * d, 0, tokT, 0, 1, 2 * d, d, 1, scr, [&](int, int n, int, float v) {
                    v += __ldg(P.ca_bkv + n);
                    if (n < d) Kt[(long long)n * Mpad + (M - 1)] = v;
                    else Vc[(long long)(M - 1) * d + (n - d)] = v;
                });
            }
            // scores (transposed): PT[hh][m][t] = q[t,hh] . K[m,hh]
            cta_gemm<TR>(Kt, Mpad, (long long)dh * Mpad, qT, dh * TR, H, M, dh, T, scr,
                         [&](int hh, int m, int t, float v) { PT[(hh * Mpad + m) * TR + t] = v; });
            // softmax over m for every (hh, t)
            for (int pair = warp; pair < H * T; pair += nwarps) {
                const int hh = pair / T, t = pair % T;
                float* col = PT + (long long)hh * Mpad * TR + t;
                float mx = -INFINITY;
                for (int m = lane; m < M; m += 32) mx = fmaxf(mx, col[m * TR]);
                mx = warp_max(mx);
                float sum = 0.f;
                for (int m = lane; m < M; m += 32) {
                    const float e = expf(col[m * TR] - mx);
                    col[m * TR] = e;
                    sum += e;
                }
                const float inv = 1.0f / warp_sum(sum);
                for (int m = lane; m < M; m += 32) col[m * TR] *= inv;
            }
            __syncthreads();
            // O = P V, written transposed as the next GEMM input
            cta_gemm<TR>(Vc, d, dh, PT, Mpad * TR, H, dh, M, T, scr,
                         [&](int hh, int c, int t, float v) { xT[(hh * dh + c) * TR + t] = v; });
            cta_gemm<TR>(P.ca_wo_t, d, 0, xT, 0, 1, d, d, T, scr,
                         [&](int, int n, int t, float v) { h[t * d + n] += v + __ldg(P.ca_bo + n); });

            // ---- feed-forward block: h += W2 gelu(W1 LN3 h) ----------------------------
            cta_layernorm_T<TR>(h, d, T, P.ln3_g, P.ln3_b, xT);
            float* ffT = big;  // [d][TR]
            cta_gemm<TR>(P.w1_t, d, 0, xT, 0, 1, d, d, T, scr,
                         [&](int, int n, int t, float v) { ffT[n * TR + t] = gelu_erf(v + __ldg(P.b1 + n)); });
            cta_gemm<TR>(P.w2_t, d, 0, ffT, 0, 1, d, d, T, scr,
                         [&](int, int n, int t, float v) { h[t * d + n] += v + __ldg(P.b2 + n); });
        }

        // ---- output projection (decoder.py:54) ------------------------------------------
        for (int i = tid; i < T * d; i += blockDim.x) xT[(i % d) * TR + (i / d)] = h[i];
        __syncthreads();
        cta_gemm<TR>(a.io.fc_wt, J, 0, xT, 0, 1, J, d, T, scr,
                     [&](int, int j, int t, float v) { eps[t * Jp + j] = v + __ldg(a.io.fc_b + j); });

        if (a.tok_mode == 1) {
            for (int i = tid; i < T * J; i += blockDim.x)
                a.eps_out[(long long)b * T * J + i] = eps[(i / J) * Jp + (i % J)];
            return;
        }
        // ---- DDIM eta=0 update ----------------------------------------------------------
        const float sb = __ldg(a.coef + 4 * s + 0), sa = __ldg(a.coef + 4 * s + 1);
        const float sap = __ldg(a.coef + 4 * s + 2), sbp = __ldg(a.coef + 4 * s + 3);
        for (int i = tid; i < T * J; i += blockDim.x) {
            const int o = (i / J) * Jp + (i % J);
            const float e = eps[o];
            if (a.eps_out) a.eps_out[((long long)s * a.B + b) * T * J + i] = e;
            const float x0 = (xs[o] - sb * e) / sa;
            xs[o] = sap * x0 + sbp * e;
        }
        __syncthreads();
    }
    for (int i = tid; i < T * J; i += blockDim.x) {
        float v = xs[(i / J) * Jp + (i % J)];
        if (a.denorm) v = v * __ldg(a.io.stdv + i % J) + __ldg(a.io.mean + i % J);
        a.x_out[(long long)b * T * J + i] = v;
    }
}

template <int TR>
size_t sampler_smem_bytes(int d, int J, int H, int Mpad) {
    const int Jp = (J + 3) & ~3;
    const int ldq = 3 * d + 1;
    size_t f = 2 * (size_t)TR * Jp + (size_t)TR * d + (size_t)d * TR + (((size_t)TR * ldq + 3) & ~(size_t)3) +
               (size_t)H * Mpad * TR + (size_t)kSamplerThreads * TR;
    return f * sizeof(float);
}


// ==========================================================================================
// Cluster sampler: ONE THREAD-BLOCK CLUSTER (16 CTAs = 16 SMs) per trajectory.
//
// bs=1 robot control is latency-bound: a single CTA leaves 147 SMs idle and streams the
// decoder weights (2 MB fp32) from L2 every step.  Here every CTA of the cluster owns 1/16 of the OUTPUT
// COLUMNS of every projection and keeps that weight slice resident in its shared memory for all steps;
// per projection each CTA computes its (T x N/16) slice and pushes it into the activation buffer of all
// 16 peers through distributed shared memory, one cluster barrier per projection (7 per layer).
// Cross-attention is split (head, key quarter) over the 16 CTAs flash-decoding style: each CTA produces
// (max, sum, partial PV) for its keys, every CTA combines the partials.  Self-attention (T x T),
// LayerNorm, residual adds, the embedding / output projections and the DDIM update are tiny and are
// computed redundantly by every CTA so that no further exchange is needed.
constexpr int kClusterSize = 16;
constexpr int kClThreads = 256;

struct ClusterLayout {   // offsets in floats inside dynamic shared memory
    int w, prm, act0, act1, h, xT, xs, eps, cab, sc, red, total;
};

__host__ __device__ inline int align4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }

// per layer: LayerNorm gamma/beta (6 d) + this CTA's bias slices (qkv 3*d/C, five projections d/C each)
__host__ __device__ inline int cluster_param_floats(int d) { return 6 * d + 8 * (d / kClusterSize); }

__host__ __device__ inline ClusterLayout cluster_layout(int d, int L, int L_res, int TR, int J, int H, int M, int T) {
    const int C = kClusterSize;
    const int dh = d / H;
    const int parts = C / H;
    const int Mq = (M + parts - 1) / parts;
    const int Jp = (J + 3) & ~3;
    int G = kClThreads / (Mq > 0 ? Mq : 1);
    if (G > dh) G = dh;
    if (G < 1) G = 1;
    const int groups2 = kClThreads / dh > 0 ? kClThreads / dh : 1;
    ClusterLayout o;
    int cur = 0;
    o.w = cur;    cur += align4(L_res * 8 * (d / C) * d);        // (3 + 5) slices of d/C columns each
    o.prm = cur;  cur += align4(L * cluster_param_floats(d));
    o.act0 = cur; cur += align4(imax(3 * d * TR, d * TR + dh * Mq));   // gathered q|k|v; q (cross) + staged K/V slice; ffn hidden
    o.act1 = cur; cur += align4(d * TR);                         // gathered residual deltas
    o.h = cur;    cur += align4(TR * d);
    o.xT = cur;   cur += align4(d * TR);
    o.xs = cur;   cur += align4(TR * Jp);
    o.eps = cur;  cur += align4(TR * Jp);
    o.cab = cur;  cur += align4(C * (2 * TR + TR * dh));
    o.sc = cur;   cur += align4(imax(TR * (Mq + 1), H * T * T));
    const int red1 = G * TR * Mq, red2 = groups2 * TR * dh;
    o.red = cur;  cur += align4(red1 > red2 ? red1 : red2);
    o.total = cur;
    return o;
}

// out[t][j] for the CTA's column slice: W(j,k) = Wb[j*sj + k*sk]; xT [K][TR].  Warp per column, lanes over k,
// butterfly reduction: afterwards EVERY lane holds the whole column out[0..TR)[j]; epi(j, column) runs on all lanes
// (lane p pushes the column to peer p with 16-byte distributed-shared-memory stores).
template <int TR, class Epi>
__device__ __forceinline__ void slice_gemm(const float* Wb, int sj, int sk, int ncols, int K, const float* xT, Epi epi) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = warp; j < ncols; j += nw) {
        float acc[TR];
#pragma unroll
        for (int t = 0; t < TR; ++t) acc[t] = 0.f;
        const float* wj = Wb + (long long)j * sj;
        for (int k = lane; k < K; k += 32) {
            const float w = wj[(long long)k * sk];
            const float4* xp = reinterpret_cast<const float4*>(xT + k * TR);
#pragma unroll
            for (int q = 0; q < TR / 4; ++q) {
                const float4 x4 = xp[q];
                acc[4 * q + 0] = fmaf(w, x4.x, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(w, x4.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(w, x4.z, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(w, x4.w, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int t = 0; t < TR; ++t) acc[t] = warp_sum(acc[t]);
        epi(j, acc);
    }
}

// LayerNorm of the T rows of h -> xT[c][t]; gamma/beta in SHARED memory (plain loads).  Half a warp per row
// (two rows per warp): with T <= 16 rows and 8 warps every row is normalised in one round.
template <int TR>
__device__ __forceinline__ void cta_layernorm_T_sm(const float* h, int d, int T, const float* gamma, const float* beta,
                                                   float* xT) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int hl = lane & 15, half = lane >> 4;
    for (int t0 = 2 * w; t0 < T; t0 += 2 * nw) {
        const int t = t0 + half;
        const bool ok = t < T;
        const float* r = h + (ok ? t : 0) * d;
        float s = 0.f;
        for (int c = hl; c < d; c += 16) s += r[c];
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);   // stays inside the half-warp
        const float mu = s / (float)d;
        float q = 0.f;
        for (int c = hl; c < d; c += 16) {
            const float dl = r[c] - mu;
            q = fmaf(dl, dl, q);
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
        const float rs = rsqrtf(q / (float)d + 1e-5f);
        if (ok)
            for (int c = hl; c < d; c += 16) xT[c * TR + t] = (r[c] - mu) * rs * gamma[c] + beta[c];
    }
    __syncthreads();
}

// D / DH / TT: compile-time hidden size, head size and trajectory length (0 = run-time values).  The default
// architecture (d=128, 4 heads, T=10) is instantiated with constants so that the index arithmetic of the many short
// loops folds to shifts/multiplies; every other shape runs the generic instantiation.
template <int TR, int D, int DH, int TT>
__global__ void __launch_bounds__(kClThreads, 1) sampler_cluster_kernel(const SamplerArgs a, const int L_res) {
    extern __shared__ __align__(16) float smem[];
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int C = kClusterSize;
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / C;
    const int d = D ? D : a.d, T = TT ? TT : a.T, dh = DH ? DH : a.dh;
    const int H = (D && DH) ? D / DH : a.heads;
    const int J = a.J, M = a.M, Mpad = a.Mpad, L = a.L;
    const int nd = d / C, nq = 3 * nd;
    const int Jp = (J + 3) & ~3;
    const ClusterLayout lo = cluster_layout(d, L, L_res, TR, J, H, M, T);
    float* Wres = smem + lo.w;
    float* prm = smem + lo.prm;
    float* act[2] = {smem + lo.act0, smem + lo.act1};
    float* h = smem + lo.h;
    float* xT = smem + lo.xT;
    float* xs = smem + lo.xs;
    float* eps = smem + lo.eps;
    float* cab = smem + lo.cab;
    float* sc = smem + lo.sc;
    float* red = smem + lo.red;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float scale = rsqrtf((float)dh);
    const int parts = C / H;
    const int Mq = (M + parts - 1) / parts;
    const int hc = rank % H, kp = rank / H;
    const int m_lo = min(M, kp * Mq), m_hi = min(M, (kp + 1) * Mq), nk = m_hi - m_lo;
    const int RS = 2 * TR + TR * dh;   // cross-attention partial record: max[TR] | sum[TR] | O[T][dh]
    const int per_layer = 8 * nd * d;
    const int PS = cluster_param_floats(d);

    // ---- one-time: weight slices (k contiguous per output column) and small parameters -> shared memory ----
    for (int l = 0; l < L_res; ++l) {
        const LayerPtrs& P = a.layers[l];
        float* wl = Wres + l * per_layer;
        const float* srcs[6] = {P.sa_wqkv_t, P.sa_wo_t, P.ca_wq_t, P.ca_wo_t, P.w1_t, P.w2_t};
        int off = 0;
        for (int g = 0; g < 6; ++g) {
            const int nc = g == 0 ? nq : nd, N = g == 0 ? 3 * d : d;
            for (int i = tid; i < nc * d; i += blockDim.x) {
                const int k = i / nc, j = i % nc;
                const int col = g == 0 ? (j / nd) * d + rank * nd + (j % nd) : rank * nd + j;
                wl[off + j * d + k] = __ldg(srcs[g] + (long long)k * N + col);
            }
            off += nc * d;
        }
    }
    for (int l = 0; l < L; ++l) {
        const LayerPtrs& P = a.layers[l];
        float* pl = prm + l * PS;
        const float* lnp[6] = {P.ln1_g, P.ln1_b, P.ln2_g, P.ln2_b, P.ln3_g, P.ln3_b};
        for (int i = tid; i < 6 * d; i += blockDim.x) pl[i] = __ldg(lnp[i / d] + (i % d));
        float* bl = pl + 6 * d;
        for (int j = tid; j < nq; j += blockDim.x) bl[j] = __ldg(P.sa_bqkv + (j / nd) * d + rank * nd + (j % nd));
        const float* bs[5] = {P.sa_bo, P.ca_bq, P.ca_bo, P.b1, P.b2};
        for (int i = tid; i < 5 * nd; i += blockDim.x) bl[nq + i] = __ldg(bs[i / nd] + rank * nd + (i % nd));
    }
    for (int i = tid; i < TR * Jp; i += blockDim.x) { xs[i] = 0.f; eps[i] = 0.f; }
    __syncthreads();
    for (int i = tid; i < T * J; i += blockDim.x) xs[(i / J) * Jp + (i % J)] = a.x_in[(long long)b * T * J + i];
    for (int i = tid; i < d * TR; i += blockDim.x) { xT[i] = 0.f; act[1][i] = 0.f; }
    for (int i = tid; i < 3 * d * TR; i += blockDim.x) act[0][i] = 0.f;
    __syncthreads();
    cluster.sync();

    // embedding weights of this thread's column stay in registers for all steps (column n = tid % d, rows tg, tg+ng, ..)
    constexpr int kMaxJ = 24;
    const bool emb_regs = d <= kClThreads && kClThreads % d == 0 && J <= kMaxJ;
    float embw[kMaxJ];
    float embb = 0.f;
    if (emb_regs) {
#pragma unroll
        for (int j = 0; j < kMaxJ; ++j) embw[j] = j < J ? __ldg(a.io.emb_wt + j * d + (tid % d)) : 0.f;
        embb = __ldg(a.io.emb_b + (tid % d));
    }

    int dbg_i = 0;
#define SD_STAMP()                                                                     \
    do {                                                                               \
        if (a.dbg && blockIdx.x == 0 && tid == 0 && dbg_i < 96) a.dbg[s * 96 + dbg_i] = clock64(); \
        ++dbg_i;                                                                       \
    } while (0)

    // weight slice accessor: resident layers read smem [j][k]; others read the K-major blob in global memory
    auto wslice = [&](int l, int g, const float* gsrc, int N, const float*& base, int& sj, int& sk) {
        if (l < L_res) {
            base = Wres + l * per_layer + (g == 0 ? 0 : nq * d + (g - 1) * nd * d);
            sj = d; sk = 1;
        } else {
            base = gsrc + rank * nd;
            sj = 1; sk = N;
        }
    };
    // lane p (< C) writes one whole column (TR floats, 16-byte stores) into peer p's buffer
    auto push_col = [&](float* buf, int col, const float* v) {
        if (lane < C) {
            float4* dst = reinterpret_cast<float4*>(cluster.map_shared_rank(buf, lane) + col * TR);
#pragma unroll
            for (int q = 0; q < TR / 4; ++q) dst[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    };

    for (int s = 0; s < a.num_steps; ++s) {
        dbg_i = 0;
        SD_STAMP();   // step start
        // ---- embedding + positional encoding (redundant in every CTA) ------------------------------
        if (emb_regs) {
            const int n = tid % d, ng = kClThreads / d;
            for (int t = tid / d; t < T; t += ng) {
                float acc = embb + __ldg(a.io.pe + t * d + n);
#pragma unroll
                for (int j = 0; j < kMaxJ; ++j) acc = fmaf(xs[t * Jp + min(j, Jp - 1)], embw[j], acc);   // embw = 0 for j >= J
                h[t * d + n] = acc;
            }
        } else {
            for (int i = tid; i < T * d; i += blockDim.x) {
                const int t = i / d, n = i % d;
                float acc = __ldg(a.io.emb_b + n) + __ldg(a.io.pe + t * d + n);
                for (int j = 0; j < J; ++j) acc = fmaf(xs[t * Jp + j], __ldg(a.io.emb_wt + j * d + n), acc);
                h[i] = acc;
            }
        }
        __syncthreads();

        for (int l = 0; l < L; ++l) {
            const LayerPtrs& P = a.layers[l];
            const float* pl = prm + l * PS;            // ln1_g ln1_b ln2_g ln2_b ln3_g ln3_b | bias slices
            const float* bl = pl + 6 * d;              // bqkv[nq] bo_sa bq_ca bo_ca b1 b2 (nd each)
            const float* wb; int sj, sk;
            SD_STAMP();   // layer start
            // ---- self-attention ---------------------------------------------------------------------
            cta_layernorm_T_sm<TR>(h, d, T, pl, pl + d, xT);
            SD_STAMP();   // after LN1
            if (l < L_res) {
                wslice(l, 0, nullptr, 0, wb, sj, sk);
                slice_gemm<TR>(wb, sj, sk, nq, d, xT, [&](int j, float* v) {
                    const float bias = bl[j];
#pragma unroll
                    for (int t = 0; t < TR; ++t) v[t] += bias;
                    push_col(act[0], (j / nd) * d + rank * nd + (j % nd), v);
                });
            } else {
                for (int part = 0; part < 3; ++part)
                    slice_gemm<TR>(P.sa_wqkv_t + part * d + rank * nd, 1, 3 * d, nd, d, xT, [&](int j, float* v) {
                        const float bias = bl[part * nd + j];
#pragma unroll
                        for (int t = 0; t < TR; ++t) v[t] += bias;
                        push_col(act[0], part * d + rank * nd + j, v);
                    });
            }
            SD_STAMP();   // after qkv gemm+push
            cluster.sync();
            SD_STAMP();   // after sync
            {
                const float* qkv = act[0];   // [3d][TR]
                // scores: one thread per (head, query, key)
                for (int idx = tid; idx < H * T * T; idx += blockDim.x) {
                    const int m = idx % T, t = (idx / T) % T, hh = idx / (T * T);
                    const float* qp = qkv + (hh * dh) * TR + t;
                    const float* kq = qkv + (d + hh * dh) * TR + m;
                    float acc = 0.f;
#pragma unroll 8
                    for (int c = 0; c < dh; ++c) acc = fmaf(qp[c * TR], kq[c * TR], acc);
                    sc[idx] = acc * scale;
                }
                __syncthreads();
                for (int row = tid; row < H * T; row += blockDim.x) {
                    float* pr = sc + row * T;
                    float mx = -INFINITY;
                    for (int m = 0; m < T; ++m) mx = fmaxf(mx, pr[m]);
                    float sum = 0.f;
                    for (int m = 0; m < T; ++m) {
                        const float e = expf(pr[m] - mx);
                        pr[m] = e;
                        sum += e;
                    }
                    const float inv = 1.0f / sum;
                    for (int m = 0; m < T; ++m) pr[m] *= inv;
                }
                __syncthreads();
                for (int idx = tid; idx < T * d; idx += blockDim.x) {
                    const int t = idx % T, n = idx / T;
                    const float* pr = sc + ((n / dh) * T + t) * T;
                    const float* vp = qkv + (2 * d + n) * TR;
                    float o = 0.f;
                    for (int m = 0; m < T; ++m) o = fmaf(pr[m], vp[m], o);
                    xT[n * TR + t] = o;
                }
                __syncthreads();
            }
            SD_STAMP();   // after self-attention core
            wslice(l, 1, P.sa_wo_t, d, wb, sj, sk);
            slice_gemm<TR>(wb, sj, sk, nd, d, xT, [&](int j, float* v) {
                const float bias = bl[nq + j];
#pragma unroll
                for (int t = 0; t < TR; ++t) v[t] += bias;
                push_col(act[1], rank * nd + j, v);
            });
            cluster.sync();
            for (int i = tid; i < T * d; i += blockDim.x) h[i] += act[1][(i % d) * TR + (i / d)];
            __syncthreads();
            // ---- cross-attention --------------------------------------------------------------------
            SD_STAMP();   // after sa out-proj + sync + residual
            cta_layernorm_T_sm<TR>(h, d, T, pl + 2 * d, pl + 3 * d, xT);
            wslice(l, 2, P.ca_wq_t, d, wb, sj, sk);
            slice_gemm<TR>(wb, sj, sk, nd, d, xT, [&](int j, float* v) {
                const float bias = bl[nq + nd + j];
#pragma unroll
                for (int t = 0; t < TR; ++t) v[t] = (v[t] + bias) * scale;
                push_col(act[0], rank * nd + j, v);
            });
            cluster.sync();
            SD_STAMP();   // after LN2 + q gemm + sync
            {
                const float* q = act[0];   // [d][TR], pre-scaled
                const float* Kt = a.Kt + ((long long)l * a.B + b) * (long long)H * dh * Mpad + (long long)hc * dh * Mpad;
                const float* Vc = a.Vc + ((long long)l * a.B + b) * (long long)Mpad * d + hc * dh;
                const float* tk = a.tok_kv + ((long long)s * L + l) * 2 * d;
                // this CTA's K slice [dh][nk] -> shared memory (free upper part of act[0]): one coalesced, fully parallel
                // round trip to L2 instead of dependent loads inside the dot products
                float* kvs = act[0] + d * TR;
                for (int i0 = tid; i0 < dh * nk; i0 += 8 * kClThreads) {   // 8 loads in flight per thread, then the stores
                    float v8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * kClThreads;
                        const int c = i / nk, mi = i - c * nk;
                        const int m = m_lo + mi;
                        v8[u] = i < dh * nk ? ((m == M - 1) ? __ldg(tk + hc * dh + c) : __ldcg(Kt + (long long)c * Mpad + m)) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * kClThreads;
                        const int c = i / nk, mi = i - c * nk;
                        if (i < dh * nk) kvs[c * Mq + mi] = v8[u];
                    }
                }
                __syncthreads();
                // scores for this CTA's keys: thread = (key, c-group)
                int G = min(dh, kClThreads / max(Mq, 1));   // same formula as cluster_layout()
                if (G < 1) G = 1;
                const int cpg = (dh + G - 1) / G;
                for (int idx = tid; idx < nk * G; idx += blockDim.x) {
                    const int mi = idx % nk, g = idx / nk;
                    float acc[TR];
#pragma unroll
                    for (int t = 0; t < TR; ++t) acc[t] = 0.f;
                    const int c1 = min(dh, (g + 1) * cpg);
                    for (int c = g * cpg; c < c1; ++c) {
                        const float kv = kvs[c * Mq + mi];
                        const float4* qp = reinterpret_cast<const float4*>(q + (hc * dh + c) * TR);
#pragma unroll
                        for (int qq = 0; qq < TR / 4; ++qq) {
                            const float4 x4 = qp[qq];
                            acc[4 * qq + 0] = fmaf(kv, x4.x, acc[4 * qq + 0]);
                            acc[4 * qq + 1] = fmaf(kv, x4.y, acc[4 * qq + 1]);
                            acc[4 * qq + 2] = fmaf(kv, x4.z, acc[4 * qq + 2]);
                            acc[4 * qq + 3] = fmaf(kv, x4.w, acc[4 * qq + 3]);
                        }
                    }
#pragma unroll
                    for (int t = 0; t < TR; ++t) red[(g * TR + t) * Mq + mi] = acc[t];
                }
                __syncthreads();
                SD_STAMP();   // after score partials
                for (int idx = tid; idx < T * nk; idx += blockDim.x) {
                    const int t = idx / nk, mi = idx % nk;
                    float v = 0.f;
                    for (int g = 0; g < G; ++g) v += red[(g * TR + t) * Mq + mi];
                    sc[t * (Mq + 1) + mi] = v;
                }
                __syncthreads();
                // partial softmax (unnormalised) per query row; record header pushed to every peer
                for (int t = warp; t < T; t += nwarps) {
                    float mx = -INFINITY;
                    for (int mi = lane; mi < nk; mi += 32) mx = fmaxf(mx, sc[t * (Mq + 1) + mi]);
                    mx = warp_max(mx);
                    float sum = 0.f;
                    for (int mi = lane; mi < nk; mi += 32) {
                        const float e = expf(sc[t * (Mq + 1) + mi] - mx);
                        sc[t * (Mq + 1) + mi] = e;
                        sum += e;
                    }
                    sum = warp_sum(sum);
                    if (lane < C) {
                        float* rc = cluster.map_shared_rank(cab, lane) + rank * RS;
                        rc[t] = mx;
                        rc[TR + t] = sum;
                    }
                }
                __syncthreads();
                SD_STAMP();   // after softmax partial
                // V slice [nk][dh] -> the same shared region (the K slice is no longer needed)
                for (int i0 = tid; i0 < nk * dh; i0 += 8 * kClThreads) {
                    float v8[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * kClThreads;
                        const int mi = i / dh, c = i - mi * dh;
                        const int m = m_lo + mi;
                        v8[u] = i < nk * dh ? ((m == M - 1) ? __ldg(tk + d + hc * dh + c) : __ldcg(Vc + (long long)m * d + c)) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int i = i0 + u * kClThreads;
                        if (i < nk * dh) kvs[i] = v8[u];
                    }
                }
                __syncthreads();
                // partial P.V : thread = (c, key group)
                const int groups2 = max(1, kClThreads / dh);
                {
                    const int c = tid % dh, g = tid / dh;
                    if (g < groups2) {
                        float acc[TR];
#pragma unroll
                        for (int t = 0; t < TR; ++t) acc[t] = 0.f;
                        for (int mi = g; mi < nk; mi += groups2) {
                            const float vv = kvs[mi * dh + c];
#pragma unroll
                            for (int t = 0; t < TR; ++t)
                                if (t < T) acc[t] = fmaf(sc[t * (Mq + 1) + mi], vv, acc[t]);
                        }
#pragma unroll
                        for (int t = 0; t < TR; ++t) red[(g * TR + t) * dh + c] = acc[t];
                    }
                }
                SD_STAMP();   // after PV partial (thread 0's own work)
                __syncthreads();
                SD_STAMP();   // after PV barrier
                // reduce over key groups, 4 channels per item, 16-byte pushes; item = (t, c4, peer quarter)
                const int dh4 = dh >> 2;
                for (int it = tid; it < T * dh4 * 4; it += blockDim.x) {
                    const int pq = it & 3, idx = it >> 2;
                    const int t = idx / dh4, c4 = idx % dh4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (int g = 0; g < groups2; ++g) {
                        const float4 r4 = *reinterpret_cast<const float4*>(red + (g * TR + t) * dh + 4 * c4);
                        v.x += r4.x; v.y += r4.y; v.z += r4.z; v.w += r4.w;
                    }
#pragma unroll
                    for (int p = 0; p < C / 4; ++p)
                        *reinterpret_cast<float4*>(cluster.map_shared_rank(cab, pq * (C / 4) + p) + rank * RS + 2 * TR + t * dh +
                                                   4 * c4) = v;
                }
            }
            SD_STAMP();   // after PV + push
            cluster.sync();
            SD_STAMP();   // after sync
            // combine the partials of all key parts -> O (every CTA, redundantly), transposed into xT
            for (int i = tid; i < T * d; i += blockDim.x) {
                const int t = i % T, n = i / T;
                const int hh = n / dh, c = n % dh;
                float mx = -INFINITY;
                for (int p2 = 0; p2 < parts; ++p2) mx = fmaxf(mx, cab[(p2 * H + hh) * RS + t]);
                float num = 0.f, den = 0.f;
                for (int p2 = 0; p2 < parts; ++p2) {
                    const float* rc = cab + (p2 * H + hh) * RS;
                    const float mk = rc[t];
                    const float wgt = (mk == -INFINITY) ? 0.f : expf(mk - mx);
                    num = fmaf(wgt, rc[2 * TR + t * dh + c], num);
                    den = fmaf(wgt, rc[TR + t], den);
                }
                xT[n * TR + t] = num / den;
            }
            __syncthreads();
            SD_STAMP();   // after combine
            wslice(l, 3, P.ca_wo_t, d, wb, sj, sk);
            slice_gemm<TR>(wb, sj, sk, nd, d, xT, [&](int j, float* v) {
                const float bias = bl[nq + 2 * nd + j];
#pragma unroll
                for (int t = 0; t < TR; ++t) v[t] += bias;
                push_col(act[1], rank * nd + j, v);
            });
            SD_STAMP();   // after ca out gemm+push
            cluster.sync();
            SD_STAMP();   // after sync
            for (int i = tid; i < T * d; i += blockDim.x) h[i] += act[1][(i % d) * TR + (i / d)];
            __syncthreads();
            // ---- feed-forward -----------------------------------------------------------------------
            SD_STAMP();   // after combine + ca out-proj + sync + residual
            cta_layernorm_T_sm<TR>(h, d, T, pl + 4 * d, pl + 5 * d, xT);
            wslice(l, 4, P.w1_t, d, wb, sj, sk);
            slice_gemm<TR>(wb, sj, sk, nd, d, xT, [&](int j, float* v) {
                const float bias = bl[nq + 3 * nd + j];
#pragma unroll
                for (int t = 0; t < TR; ++t) v[t] = gelu_erf(v[t] + bias);
                push_col(act[0], rank * nd + j, v);
            });
            cluster.sync();
            wslice(l, 5, P.w2_t, d, wb, sj, sk);
            slice_gemm<TR>(wb, sj, sk, nd, d, act[0], [&](int j, float* v) {   // gathered hidden = GEMM input layout
                const float bias = bl[nq + 4 * nd + j];
#pragma unroll
                for (int t = 0; t < TR; ++t) v[t] += bias;
                push_col(act[1], rank * nd + j, v);
            });
            cluster.sync();
            for (int i = tid; i < T * d; i += blockDim.x) h[i] += act[1][(i % d) * TR + (i / d)];
            __syncthreads();
        }

        SD_STAMP();   // after all layers
        // ---- output projection + DDIM update (redundant) --------------------------------------------
        for (int i = tid; i < T * d; i += blockDim.x) xT[(i % d) * TR + (i / d)] = h[i];
        __syncthreads();
        slice_gemm<TR>(a.io.fc_wt, 1, J, J, d, xT, [&](int j, float* v) {   // weights [d][J] in global memory
            const float bias = __ldg(a.io.fc_b + j);
            if (lane < T) {
                float e = 0.f;
#pragma unroll
                for (int t = 0; t < TR; ++t)
                    if (lane == t) e = v[t];
                eps[lane * Jp + j] = e + bias;
            }
        });
        __syncthreads();
        const float sb = __ldg(a.coef + 4 * s + 0), sa = __ldg(a.coef + 4 * s + 1);
        const float sap = __ldg(a.coef + 4 * s + 2), sbp = __ldg(a.coef + 4 * s + 3);
        for (int i = tid; i < T * J; i += blockDim.x) {
            const int o = (i / J) * Jp + (i % J);
            const float e = eps[o];
            if (a.eps_out && rank == 0) a.eps_out[((long long)s * a.B + b) * T * J + i] = e;
            const float x0 = (xs[o] - sb * e) / sa;
            xs[o] = sap * x0 + sbp * e;
        }
        __syncthreads();
        SD_STAMP();   // step end
    }
#undef SD_STAMP
    if (rank == 0) {
        for (int i = tid; i < T * J; i += blockDim.x) {
            float v = xs[(i / J) * Jp + (i % J)];
            if (a.denorm) v = v * __ldg(a.io.stdv + i % J) + __ldg(a.io.mean + i % J);
            a.x_out[(long long)b * T * J + i] = v;
        }
    }
    cluster.sync();   // no CTA may exit while peers can still address its shared memory
}

// dst[c][r] = src[r][c]
__global__ void transpose_kernel(const float* __restrict__ src, long long ld_src, float* __restrict__ dst, int rows,
                                 int cols) {
    __shared__ float tile[32][33];
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * ld_src + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < rows && c < cols) dst[(long long)c * rows + r] = tile[threadIdx.x][i];
    }
}

// kv_tmp[(b*Mc + m)][l*2d + n]  ->  Kt[l][b][hh][c][m] (n<d, n = hh*dh+c) , Vc[l][b][m][n-d]
__global__ void kv_relayout_kernel(const float* __restrict__ tmp, float* __restrict__ Kt, float* __restrict__ Vc, int L,
                                   int B, int Mc, int Mpad, int d) {
    const long long total = (long long)B * Mc * L * 2 * d;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int n2 = (int)(i % (2 * d));
        long long r = i / (2 * d);
        const int l = (int)(r % L);
        r /= L;
        const int m = (int)(r % Mc);
        const int b = (int)(r / Mc);
        const float v = tmp[i];
        if (n2 < d) Kt[(((long long)l * B + b) * d + n2) * Mpad + m] = v;
        else Vc[(((long long)l * B + b) * Mpad + m) * d + (n2 - d)] = v;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// plan
struct sd_plan {
    sd_plan_config cfg;
    int M, Mpad;           // memory length incl. step token
    float* blob = nullptr; // packed weights
    size_t blob_floats = 0;
    LayerPtrs* d_layers = nullptr;
    std::vector<LayerPtrs> h_layers;
    IoPtrs io{};
    float* wkv_all = nullptr;   // [L*2d][d]  rows d:3d of every multihead_attn.in_proj_weight (NK layout)
    float* bkv_all = nullptr;   // [L*2d]
    float* Kt = nullptr; float* Vc = nullptr; float* kv_tmp = nullptr;
    int cache_B = 0;
    float* tok_kv = nullptr; float* coef = nullptr; float* tok_tmp = nullptr; long long* d_timesteps = nullptr;
    int num_steps = 0, sched_cap = 0;
    int ctx_B = 0;
    bool layers_dirty = true;
    int sampler_mode = 0;   // 0 auto, 1 one CTA per trajectory, 2 one 16-CTA cluster per trajectory
    int last_sampler = 0;   // which kernel the last sd_plan_sample launched (1 / 2)
    long long* dbg = nullptr;   // optional phase-timestamp buffer (device), num_steps * 64 entries
};

namespace {
size_t layer_floats(int d) { return (size_t)6 * d + (size_t)3 * d * d + 3 * d + (size_t)d * d + d + (size_t)d * d + d + (size_t)2 * d * d + 2 * d + (size_t)d * d + d + (size_t)d * d + d + (size_t)d * d + d; }
}

extern "C" int sd_plan_create(const sd_plan_config* cfg, sd_plan** out) {
    if (!cfg || !out) return SD_ERR_BAD_ARG;
    if (cfg->d <= 0 || cfg->d % 4 != 0 || cfg->heads <= 0 || cfg->d % cfg->heads != 0 || cfg->layers <= 0 ||
        cfg->layers > kMaxLayers || cfg->T <= 0 || cfg->J <= 0 || cfg->ctx_tokens < 0)
        return SD_ERR_BAD_ARG;
    if (cfg->T > 32) return SD_ERR_UNSUPPORTED;
    sd_plan* p = new (std::nothrow) sd_plan();
    if (!p) return (int)cudaErrorMemoryAllocation;
    p->cfg = *cfg;
    p->M = cfg->ctx_tokens + 1;
    p->Mpad = (p->M + 3) & ~3;
    const int d = cfg->d, L = cfg->layers, T = cfg->T, J = cfg->J;
    const size_t io_floats = (size_t)J * d + d + (size_t)T * d + (size_t)d * J + 3 * (size_t)J + d / 4 + d / 2;
    // every carve below is rounded up to a multiple of 4 floats: 20 per layer + 9 io regions
    p->blob_floats = layer_floats(d) * L + io_floats + 4 * ((size_t)20 * L + 9) + 64;
    cudaError_t e;
    if ((e = cudaMalloc(&p->blob, p->blob_floats * sizeof(float))) != cudaSuccess ||
        (e = cudaMalloc(&p->d_layers, sizeof(LayerPtrs) * L)) != cudaSuccess ||
        (e = cudaMalloc(&p->wkv_all, sizeof(float) * (size_t)L * 2 * d * d)) != cudaSuccess ||
        (e = cudaMalloc(&p->bkv_all, sizeof(float) * (size_t)L * 2 * d)) != cudaSuccess) {
        sd_plan_destroy(p);
        return (int)e;
    }
    cudaMemset(p->blob, 0, p->blob_floats * sizeof(float));
    // carve
    float* cur = p->blob;
    auto take = [&](size_t n) { float* r = cur; cur += (n + 3) & ~(size_t)3; return r; };
    p->h_layers.resize(L);
    for (int l = 0; l < L; ++l) {
        LayerPtrs& q = p->h_layers[l];
        q.ln1_g = take(d); q.ln1_b = take(d); q.ln2_g = take(d); q.ln2_b = take(d); q.ln3_g = take(d); q.ln3_b = take(d);
        q.sa_wqkv_t = take((size_t)3 * d * d); q.sa_bqkv = take(3 * d);
        q.sa_wo_t = take((size_t)d * d); q.sa_bo = take(d);
        q.ca_wq_t = take((size_t)d * d); q.ca_bq = take(d);
        q.ca_wkv_t = take((size_t)2 * d * d); q.ca_bkv = take(2 * d);
        q.ca_wo_t = take((size_t)d * d); q.ca_bo = take(d);
        q.w1_t = take((size_t)d * d); q.b1 = take(d);
        q.w2_t = take((size_t)d * d); q.b2 = take(d);
    }
    p->io.emb_wt = take((size_t)J * d); p->io.emb_b = take(d); p->io.pe = take((size_t)T * d);
    p->io.fc_wt = take((size_t)d * J); p->io.fc_b = take(J); p->io.freqs = take(d / 4); p->io.token = take(d / 2);
    p->io.mean = take(J); p->io.stdv = take(J);
    if ((size_t)(cur - p->blob) > p->blob_floats) { sd_plan_destroy(p); return SD_ERR_BAD_ARG; }
    if ((e = cudaMemcpy(p->d_layers, p->h_layers.data(), sizeof(LayerPtrs) * L, cudaMemcpyHostToDevice)) != cudaSuccess) {
        sd_plan_destroy(p);
        return (int)e;
    }
    *out = p;
    return SD_OK;
}

extern "C" int sd_plan_destroy(sd_plan* p) {
    if (!p) return SD_OK;
    cudaFree(p->blob); cudaFree(p->d_layers); cudaFree(p->wkv_all); cudaFree(p->bkv_all);
    cudaFree(p->Kt); cudaFree(p->Vc); cudaFree(p->kv_tmp);
    cudaFree(p->tok_kv); cudaFree(p->coef); cudaFree(p->tok_tmp); cudaFree(p->d_timesteps);
    delete p;
    return SD_OK;
}

namespace {
int copy_vec(const float* src, const float* dst, size_t n, cudaStream_t st) {
    SD_CUDA(cudaMemcpyAsync(const_cast<float*>(dst), src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SD_OK;
}
int transpose_into(const float* src, long long ld_src, const float* dst, int rows, int cols, cudaStream_t st) {
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(src, ld_src, const_cast<float*>(dst), rows, cols);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
}  // namespace

#define SD_TRY(x)                 \
    do {                          \
        int rc__ = (x);           \
        if (rc__ != SD_OK) return rc__; \
    } while (0)

extern "C" int sd_plan_set_layer(sd_plan* p, int layer, const sd_decoder_layer_weights* w, void* stream) {
    if (!p) return SD_ERR_NO_PLAN;
    if (!w || layer < 0 || layer >= p->cfg.layers) return SD_ERR_BAD_ARG;
    const int d = p->cfg.d;
    cudaStream_t st = (cudaStream_t)stream;
    const LayerPtrs& q = p->h_layers[layer];
    SD_TRY(copy_vec(w->norm1_w, q.ln1_g, d, st)); SD_TRY(copy_vec(w->norm1_b, q.ln1_b, d, st));
    SD_TRY(copy_vec(w->norm2_w, q.ln2_g, d, st)); SD_TRY(copy_vec(w->norm2_b, q.ln2_b, d, st));
    SD_TRY(copy_vec(w->norm3_w, q.ln3_g, d, st)); SD_TRY(copy_vec(w->norm3_b, q.ln3_b, d, st));
    SD_TRY(transpose_into(w->sa_in_w, d, q.sa_wqkv_t, 3 * d, d, st)); SD_TRY(copy_vec(w->sa_in_b, q.sa_bqkv, 3 * d, st));
    SD_TRY(transpose_into(w->sa_out_w, d, q.sa_wo_t, d, d, st)); SD_TRY(copy_vec(w->sa_out_b, q.sa_bo, d, st));
    SD_TRY(transpose_into(w->ca_in_w, d, q.ca_wq_t, d, d, st)); SD_TRY(copy_vec(w->ca_in_b, q.ca_bq, d, st));
    SD_TRY(transpose_into(w->ca_in_w + (size_t)d * d, d, q.ca_wkv_t, 2 * d, d, st));
    SD_TRY(copy_vec(w->ca_in_b + d, q.ca_bkv, 2 * d, st));
    SD_TRY(transpose_into(w->ca_out_w, d, q.ca_wo_t, d, d, st)); SD_TRY(copy_vec(w->ca_out_b, q.ca_bo, d, st));
    SD_TRY(transpose_into(w->lin1_w, d, q.w1_t, d, d, st)); SD_TRY(copy_vec(w->lin1_b, q.b1, d, st));
    SD_TRY(transpose_into(w->lin2_w, d, q.w2_t, d, d, st)); SD_TRY(copy_vec(w->lin2_b, q.b2, d, st));
    // concatenated K/V projection (reference layout, NK) for the one-shot context GEMM
    SD_TRY(copy_vec(w->ca_in_w + (size_t)d * d, p->wkv_all + (size_t)layer * 2 * d * d, (size_t)2 * d * d, st));
    SD_TRY(copy_vec(w->ca_in_b + d, p->bkv_all + (size_t)layer * 2 * d, 2 * d, st));
    p->ctx_B = 0;       // caches are stale
    p->num_steps = 0;   // token tables are stale
    return SD_OK;
}

extern "C" int sd_plan_set_io(sd_plan* p, const float* emb_w, const float* emb_b, const float* fc_w,
                              const float* fc_b, const float* pe, const float* step_freqs, const float* step_token,
                              const float* mean, const float* stdv, void* stream) {
    if (!p) return SD_ERR_NO_PLAN;
    if (!emb_w || !emb_b || !fc_w || !fc_b || !pe || !step_freqs || !step_token) return SD_ERR_BAD_ARG;
    const int d = p->cfg.d, J = p->cfg.J, T = p->cfg.T;
    cudaStream_t st = (cudaStream_t)stream;
    SD_TRY(transpose_into(emb_w, J, p->io.emb_wt, d, J, st));   // (d,J) -> [J][d]
    SD_TRY(copy_vec(emb_b, p->io.emb_b, d, st));
    SD_TRY(copy_vec(pe, p->io.pe, (size_t)T * d, st));
    SD_TRY(transpose_into(fc_w, d, p->io.fc_wt, J, d, st));     // (J,d) -> [d][J]
    SD_TRY(copy_vec(fc_b, p->io.fc_b, J, st));
    SD_TRY(copy_vec(step_freqs, p->io.freqs, d / 4, st));
    SD_TRY(copy_vec(step_token, p->io.token, d / 2, st));
    if (mean && stdv) {
        SD_TRY(copy_vec(mean, p->io.mean, J, st));
        SD_TRY(copy_vec(stdv, p->io.stdv, J, st));
    }
    p->num_steps = 0;
    return SD_OK;
}

extern "C" int sd_plan_set_schedule(sd_plan* p, int num_steps, const long long* timesteps_host,
                                    const float* coef_host, void* stream) {
    if (!p) return SD_ERR_NO_PLAN;
    if (num_steps <= 0 || !timesteps_host || !coef_host) return SD_ERR_BAD_ARG;
    const int d = p->cfg.d, L = p->cfg.layers;
    cudaStream_t st = (cudaStream_t)stream;
    if (num_steps > p->sched_cap) {
        cudaFree(p->tok_kv); cudaFree(p->coef); cudaFree(p->tok_tmp); cudaFree(p->d_timesteps);
        p->tok_kv = p->coef = p->tok_tmp = nullptr; p->d_timesteps = nullptr;
        SD_CUDA(cudaMalloc(&p->tok_kv, sizeof(float) * (size_t)num_steps * L * 2 * d));
        SD_CUDA(cudaMalloc(&p->coef, sizeof(float) * (size_t)num_steps * 4));
        SD_CUDA(cudaMalloc(&p->tok_tmp, sizeof(float) * (size_t)num_steps * d));
        SD_CUDA(cudaMalloc(&p->d_timesteps, sizeof(long long) * (size_t)num_steps));
        p->sched_cap = num_steps;
    }
    // pageable host -> device copies are staged synchronously by the runtime: safe to return
    SD_CUDA(cudaMemcpyAsync(p->coef, coef_host, sizeof(float) * (size_t)num_steps * 4, cudaMemcpyHostToDevice, st));
    SD_CUDA(cudaMemcpyAsync(p->d_timesteps, timesteps_host, sizeof(long long) * (size_t)num_steps,
                            cudaMemcpyHostToDevice, st));
    SD_TRY(sd_step_token(p->d_timesteps, 0, p->io.freqs, p->io.token, p->tok_tmp, d, num_steps, d, stream));
    sd_gemm_desc g{};
    g.A = p->tok_tmp; g.lda = d; g.a_layout = SD_LAYOUT_MK;
    g.B = p->wkv_all; g.ldb = d; g.b_layout = SD_LAYOUT_NK;
    g.C = p->tok_kv; g.ldc = (long long)L * 2 * d;
    g.M = num_steps; g.N = L * 2 * d; g.K = d;
    g.bias = p->bkv_all; g.precision = SD_PREC_FP32;
    SD_TRY(sd_gemm(&g, stream));
    p->num_steps = num_steps;
    return SD_OK;
}

extern "C" int sd_plan_set_context(sd_plan* p, const float* ctx, int B, void* stream) {
    if (!p) return SD_ERR_NO_PLAN;
    if (B <= 0 || (!ctx && p->cfg.ctx_tokens > 0)) return SD_ERR_BAD_ARG;
    const int d = p->cfg.d, L = p->cfg.layers, Mc = p->cfg.ctx_tokens;
    cudaStream_t st = (cudaStream_t)stream;
    if (B > p->cache_B) {
        cudaFree(p->Kt); cudaFree(p->Vc); cudaFree(p->kv_tmp);
        p->Kt = p->Vc = p->kv_tmp = nullptr; p->cache_B = 0;
        const size_t n = (size_t)L * B * p->Mpad * d;
        SD_CUDA(cudaMalloc(&p->Kt, sizeof(float) * n));
        SD_CUDA(cudaMalloc(&p->Vc, sizeof(float) * n));
        SD_CUDA(cudaMalloc(&p->kv_tmp, sizeof(float) * (size_t)B * (Mc > 0 ? Mc : 1) * L * 2 * d));
        SD_CUDA(cudaMemsetAsync(p->Kt, 0, sizeof(float) * n, st));
        SD_CUDA(cudaMemsetAsync(p->Vc, 0, sizeof(float) * n, st));
        p->cache_B = B;
    }
    if (Mc > 0) {
        sd_gemm_desc g{};
        g.A = ctx; g.lda = d; g.a_layout = SD_LAYOUT_MK;
        g.B = p->wkv_all; g.ldb = d; g.b_layout = SD_LAYOUT_NK;
        g.C = p->kv_tmp; g.ldc = (long long)L * 2 * d;
        g.M = B * Mc; g.N = L * 2 * d; g.K = d;
        g.bias = p->bkv_all; g.precision = SD_PREC_FP32;
        SD_TRY(sd_gemm(&g, stream));
        const long long total = (long long)B * Mc * L * 2 * d;
        kv_relayout_kernel<<<min(ceil_div(total, 256), 148 * 16), 256, 0, st>>>(p->kv_tmp, p->Kt, p->Vc, L, B, Mc,
                                                                               p->Mpad, d);
        SD_LAUNCH_CHECK();
    }
    // NOTE: caches are laid out for batch size B (stride uses B): remember it
    p->ctx_B = B;
    return SD_OK;
}

namespace {
template <int TR>
int launch_sampler(sd_plan* p, SamplerArgs& a, cudaStream_t st) {
    const size_t bytes = sampler_smem_bytes<TR>(a.d, a.J, a.heads, a.Mpad);
    if (bytes > 227 * 1024) return SD_ERR_UNSUPPORTED;
    SD_CUDA(cudaFuncSetAttribute(sampler_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    sampler_kernel<TR><<<a.B, kSamplerThreads, bytes, st>>>(a);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
// 0: auto, 1: one CTA per trajectory, 2: one 16-CTA cluster per trajectory   (SD_B200_SAMPLER=auto|cta|cluster)
int sampler_mode() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("SD_B200_SAMPLER");
        mode = (!e || !strcmp(e, "auto")) ? 0 : (!strcmp(e, "cta") ? 1 : (!strcmp(e, "cluster") ? 2 : 0));
    }
    return mode;
}

template <int TR, int D, int DH, int TT>
int launch_cluster(sd_plan* p, SamplerArgs& a, cudaStream_t st, bool* launched) {
    *launched = false;
    const int C = kClusterSize;
    if (a.d % C != 0 || C % a.heads != 0 || a.T > TR || a.T > 32 || a.dh > kClThreads || a.dh % 4 != 0) return SD_OK;
    // resident layers: as many as fit beside the working buffers
    int L_res = a.L;
    size_t bytes = 0;
    for (; L_res >= 0; --L_res) {
        bytes = (size_t)cluster_layout(a.d, a.L, L_res, TR, a.J, a.heads, a.M, a.T).total * sizeof(float);
        if (bytes <= 227 * 1024) break;
    }
    if (L_res < 0) return SD_OK;
    auto kernel = sampler_cluster_kernel<TR, D, DH, TT>;
    static bool configured = false;
    static bool usable = true;
    if (!usable) return SD_OK;
    if (!configured) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            cudaGetLastError();
            usable = false;
            return SD_OK;
        }
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(a.B * C));
    cfg.blockDim = dim3(kClThreads);
    cfg.dynamicSmemBytes = bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static size_t checked_bytes = 0;
    if (checked_bytes != bytes) {
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1) {
            cudaGetLastError();
            usable = false;   // this device cannot co-schedule 16 CTAs of this size: use the single-CTA sampler
            return SD_OK;
        }
        checked_bytes = bytes;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a, L_res);
    if (e != cudaSuccess) {
        cudaGetLastError();
        usable = false;
        return SD_OK;
    }
    *launched = true;
    return SD_OK;
}

int run_sampler(sd_plan* p, SamplerArgs& a, cudaStream_t st) {
    const int mode = p->sampler_mode != 0 ? p->sampler_mode : sampler_mode();
    p->last_sampler = 1;
    // auto: ~9 clusters of 16 SMs run concurrently (3.3 ms per trajectory) vs 148 single-CTA trajectories (7.5 ms):
    // the cluster kernel wins up to two waves of clusters
    if (a.tok_mode == 0 && mode != 1 && (mode == 2 || a.B <= 18)) {
        bool launched = false;
        int rc = SD_OK;
        if (a.d == 128 && a.dh == 32 && a.T == 10) rc = launch_cluster<12, 128, 32, 10>(p, a, st, &launched);
        else if (a.T <= 12) rc = launch_cluster<12, 0, 0, 0>(p, a, st, &launched);
        else if (a.T <= 16) rc = launch_cluster<16, 0, 0, 0>(p, a, st, &launched);
        else if (a.T <= 20) rc = launch_cluster<20, 0, 0, 0>(p, a, st, &launched);
        if (rc != SD_OK) return rc;
        if (launched) {
            p->last_sampler = 2;
            return SD_OK;
        }
    }
    if (a.T <= 12) return launch_sampler<12>(p, a, st);
    if (a.T <= 20) return launch_sampler<20>(p, a, st);
    if (a.T <= 32) return launch_sampler<32>(p, a, st);
    return SD_ERR_UNSUPPORTED;
}
void fill_args(sd_plan* p, SamplerArgs& a) {
    a.layers = p->d_layers; a.io = p->io; a.L = p->cfg.layers; a.d = p->cfg.d; a.heads = p->cfg.heads;
    a.dh = p->cfg.d / p->cfg.heads; a.T = p->cfg.T; a.J = p->cfg.J; a.M = p->M; a.Mpad = p->Mpad; a.B = p->ctx_B;
    a.Kt = p->Kt; a.Vc = p->Vc; a.tok_kv = p->tok_kv; a.coef = p->coef; a.num_steps = p->num_steps;
}
}  // namespace

extern "C" int sd_plan_sample(sd_plan* p, const float* x_T, float* x_out, float* eps_trace, int denormalize, int B,
                              void* stream) {
    if (!p) return SD_ERR_NO_PLAN;
    if (!x_T || !x_out) return SD_ERR_BAD_ARG;
    if (p->ctx_B <= 0 || p->num_steps <= 0) return SD_ERR_BAD_ARG;  // set_context / set_schedule first
    if (B != p->ctx_B) return SD_ERR_BAD_ARG;   // x_T / x_out hold B trajectories: must be the context's batch
    SamplerArgs a{};
    fill_args(p, a);
    a.tok_mode = 0; a.t_ptr = nullptr; a.t_is_float = 0;
    a.x_in = x_T; a.x_out = x_out; a.eps_out = eps_trace; a.denorm = denormalize;
    a.dbg = p->dbg;
    return run_sampler(p, a, (cudaStream_t)stream);
}

extern "C" int sd_plan_set_sampler(sd_plan* p, int mode) {
    if (!p) return SD_ERR_NO_PLAN;
    if (mode < 0 || mode > 2) return SD_ERR_BAD_ARG;
    p->sampler_mode = mode;
    return SD_OK;
}

extern "C" int sd_plan_set_debug_stamps(sd_plan* p, long long* device_buffer) {
    if (!p) return SD_ERR_NO_PLAN;
    p->dbg = device_buffer;
    return SD_OK;
}

extern "C" int sd_plan_last_sampler(const sd_plan* p) { return p ? p->last_sampler : SD_ERR_NO_PLAN; }

extern "C" int sd_plan_denoise(sd_plan* p, const float* x, const void* t, int t_is_float, float* eps_out, int B,
                               void* stream) {
    if (!p) return SD_ERR_NO_PLAN;
    if (!x || !t || !eps_out) return SD_ERR_BAD_ARG;
    if (p->ctx_B <= 0 || B != p->ctx_B) return SD_ERR_BAD_ARG;
    SamplerArgs a{};
    fill_args(p, a);
    a.num_steps = 1; a.tok_mode = 1; a.t_ptr = t; a.t_is_float = t_is_float;
    a.x_in = x; a.x_out = nullptr; a.eps_out = eps_out; a.denorm = 0;
    return run_sampler(p, a, (cudaStream_t)stream);
}
