// HBM-bound kernels of the image trunk's non-convolution layers, bf16 NHWC ("channels_last"):
// training/eval BatchNorm fused with ReLU and the residual add, and the 3x3/stride-2 max-pool — forward and
// backward.  The convolutions themselves stay cuDNN library calls (SURVEY.md §8 a8', row (f)-1); these kernels
// replace torch's batch_norm_*_channels_last / max_pool_*_nhwc / elementwise kernels, which the round-1 launch
// list showed to be 52 % of the full training step at ~10 % of HBM bandwidth.
//
// Reference semantics: torchvision resnet BasicBlock/Bottleneck (bn -> relu, bn -> (+identity) -> relu,
// downsample bn), nn.BatchNorm2d training mode (batch statistics, biased variance for normalisation, unbiased
// for running_var, momentum 0.1, eps 1e-5) as run by ml/model/encoder/image.py:46-52 under train().
//
// Layout: x[R][C] bf16, R = N*H*W, C in {64,...,2048} multiple of 8.  Every thread owns 8 consecutive channels
// (one 16-byte vector) and walks rows with a stride that keeps its channel group fixed, so per-channel
// parameters / accumulators live in registers and all accesses are 16-byte, fully coalesced.
#include "common.cuh"
#include "../../include/sd_b200.h"

using namespace sd;

namespace {

constexpr int kT = 256;

struct bf8 { uint4 u; };
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __bfloat1622float2(p[i]);
        f[2 * i] = v.x;
        f[2 * i + 1] = v.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 u;
    __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// ---- per-channel sums over rows: out[0][c] += sum f0, out[1][c] += sum f1 (double atomics) -------------
// MODE 0: f0 = x, f1 = x^2                         (forward statistics)
// MODE 1: g = dy * (y > 0 if y); f0 = g, f1 = g * xhat  (backward reductions)
template <int MODE>
__global__ void __launch_bounds__(kT, MODE == 0 ? 6 : 3) bn_reduce_kernel(
    const uint4* __restrict__ x, const uint4* __restrict__ dy, const unsigned char* __restrict__ y,
    const float* __restrict__ mean, const float* __restrict__ invstd, long long nvec, int CV, double* __restrict__ out, int C,
    const float* __restrict__ gamma_rc, const float* __restrict__ beta_rc, const uint4* __restrict__ dy2) {
    __shared__ float red[2][kT][8 + 1];
    const int tid = threadIdx.x;
    const int cv = tid % CV;   // kT % CV == 0: a thread keeps its channel group for every vector it visits
    float sc[8], sh[8];
    if (MODE == 1 && beta_rc) {   // ReLU mask recomputed from x: relu(x*sc + sh) > 0
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sc[i] = invstd[cv * 8 + i] * gamma_rc[cv * 8 + i];
            sh[i] = beta_rc[cv * 8 + i] - mean[cv * 8 + i] * sc[i];
        }
    }
    // MODE 1 accumulates sum g and the RAW sum g*x; sum g*xhat = invstd * (sum g*x - mean * sum g) is formed once per
    // thread after the loop, which keeps the per-channel statistics out of the loop's live registers
    float a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
    const long long per_cta = ((nvec + gridDim.x - 1) / gridDim.x + kT - 1) / kT * kT;
    const long long v0 = blockIdx.x * per_cta, v1 = min(nvec, v0 + per_cta);
    constexpr int U = MODE == 0 ? 4 : 2;   // vectors per iteration: all their loads are issued before any use
    for (long long v = v0 + tid; v < v1; v += U * kT) {
        uint4 ux[U], ud[U], ue[U];
        unsigned mk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long w = v + (long long)u * kT;
            mk[u] = 0xFFu;
            if (w < v1) {
                ux[u] = ld_stream(x + w);
                if (MODE == 1) {
                    ud[u] = ld_stream(dy + w);
                    if (dy2) ue[u] = ld_stream(dy2 + w);
                    if (y) mk[u] = y[w];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v + (long long)u * kT >= v1) break;
            float fx[8];
            unpack8(ux[u], fx);
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { a0[i] += fx[i]; a1[i] = fmaf(fx[i], fx[i], a1[i]); }
            } else {
                float g[8];
                unpack8(ud[u], g);
                if (dy2) {   // the output had two consumers: their gradients are summed here, in fp32
                    float g2[8];
                    unpack8(ue[u], g2);
#pragma unroll
                    for (int i = 0; i < 8; ++i) g[i] += g2[i];
                }
                if (y) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) g[i] = (mk[u] >> i) & 1u ? g[i] : 0.f;
                } else if (beta_rc) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) g[i] = fmaf(fx[i], sc[i], sh[i]) > 0.f ? g[i] : 0.f;
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) { a0[i] += g[i]; a1[i] = fmaf(g[i], fx[i], a1[i]); }
            }
        }
    }
    if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a1[i] = invstd[cv * 8 + i] * fmaf(-mean[cv * 8 + i], a0[i], a1[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[0][tid][i] = a0[i]; red[1][tid][i] = a1[i]; }
    __syncthreads();
    // threads tid < CV*8*2 each finish one (which, channel) over the kT/CV row lanes
    for (int o = tid; o < 2 * CV * 8; o += kT) {
        const int which = o / (CV * 8), c = o % (CV * 8);
        const int ccv = c / 8, ci = c % 8;
        float s = 0.f;
        for (int r = ccv; r < kT; r += CV) s += red[which][r][ci];
        atomicAdd(&out[which * C + c], (double)s);
    }
}

// mean / invstd from the sums; running statistics update like nn.BatchNorm2d (momentum, unbiased running_var)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long R, int C, float eps, float momentum,
                                   float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sums[c] / (double)R;
    double var = sums[C + c] / (double)R - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
        running_mean[c] = (float)((1.0 - momentum) * running_mean[c] + momentum * m);
        running_var[c] = (float)((1.0 - momentum) * running_var[c] + momentum * unbiased);
    }
}

// y = act((x - mean) * invstd * gamma + beta (+ residual))
__global__ void __launch_bounds__(kT) bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ res,
                                                      const float* __restrict__ mean, const float* __restrict__ invstd,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      int relu, uint4* __restrict__ y, unsigned char* __restrict__ mask,
                                                      long long nvec, int CV) {
    const long long stride = (long long)gridDim.x * kT;
    const long long v0 = (long long)blockIdx.x * kT + threadIdx.x;
    const int cv = threadIdx.x % CV;   // stride % CV == 0
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cv * 8 + i;
        sc[i] = invstd[c] * gamma[c];
        sh[i] = beta[c] - mean[c] * sc[i];
    }
    for (long long v = v0; v < nvec; v += 2 * stride) {
        const long long w = v + stride;
        const bool two = w < nvec;
        uint4 ux[2], ur[2];
        ux[0] = ld_stream(x + v);
        if (two) ux[1] = ld_stream(x + w);
        if (res) {
            ur[0] = ld_stream(res + v);
            if (two) ur[1] = ld_stream(res + w);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            const long long vv = u ? w : v;
            float f[8];
            unpack8(ux[u], f);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], sc[i], sh[i]);
            if (res) {
                float r[8];
                unpack8(ur[u], r);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += r[i];
            }
            if (relu) {
                unsigned mk = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    mk |= (f[i] > 0.f ? 1u : 0u) << i;
                    f[i] = fmaxf(f[i], 0.f);
                }
                if (mask) mask[vv] = (unsigned char)mk;
            }
            y[vv] = pack8(f);
        }
    }
}

// g = dy * (y > 0);  dx = gamma * invstd * (g - sum_g/R - xhat * sum_gx/R);  dres = g
// evaluated as dx = kg * g + kx * x + kc with three per-channel constants (kg = gamma*invstd, kx = -kg*invstd*sum_gx/R,
// kc = -kg*sum_g/R - kx*mean): 24 live registers of constants instead of 48, three CTAs per SM instead of two.
__global__ void __launch_bounds__(kT, 3) bn_bwd_apply_kernel(const uint4* __restrict__ dy, const unsigned char* __restrict__ y,
                                                             const uint4* __restrict__ x, const float* __restrict__ mean,
                                                             const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                             const double* __restrict__ sums, long long R,
                                                             uint4* __restrict__ dx, uint4* __restrict__ dres, long long nvec,
                                                             int CV, int C, const float* __restrict__ beta_rc,
                                                             const uint4* __restrict__ dy2) {
    const long long stride = (long long)gridDim.x * kT;
    const long long v0 = (long long)blockIdx.x * kT + threadIdx.x;
    const int cv = threadIdx.x % CV;
    const double invR = 1.0 / (double)R;
    float kg[8], kx[8], kc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cv * 8 + i;
        const float mu = mean[c], is = invstd[c];
        kg[i] = gamma[c] * is;
        const float k1 = (float)(sums[c] * invR), k2 = (float)(sums[C + c] * invR);
        kx[i] = -kg[i] * is * k2;
        kc[i] = -kg[i] * k1 - kx[i] * mu;
        sh[i] = beta_rc ? beta_rc[c] - mu * kg[i] : 0.f;
    }
    for (long long v = v0; v < nvec; v += 2 * stride) {
        const long long w = v + stride;
        const bool two = w < nvec;
        uint4 ud[2], ux[2], ue[2];
        unsigned mk[2] = {0xFFu, 0xFFu};
        ud[0] = ld_stream(dy + v);
        ux[0] = ld_stream(x + v);
        if (dy2) ue[0] = ld_stream(dy2 + v);
        if (two) {
            ud[1] = ld_stream(dy + w);
            ux[1] = ld_stream(x + w);
            if (dy2) ue[1] = ld_stream(dy2 + w);
        }
        if (y) {
            mk[0] = y[v];
            if (two) mk[1] = y[w];
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            const long long vv = u ? w : v;
            float g[8], fx[8];
            unpack8(ud[u], g);
            unpack8(ux[u], fx);
            if (dy2) {
                float g2[8];
                unpack8(ue[u], g2);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] += g2[i];
            }
            if (y) {
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = (mk[u] >> i) & 1u ? g[i] : 0.f;
            } else if (beta_rc) {
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = fmaf(fx[i], kg[i], sh[i]) > 0.f ? g[i] : 0.f;
            }
            if (dres) dres[vv] = pack8(g);
#pragma unroll
            for (int i = 0; i < 8; ++i) fx[i] = fmaf(kg[i], g[i], fmaf(kx[i], fx[i], kc[i]));
            dx[vv] = pack8(fx);
        }
    }
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int C, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)sums[c];
    dgamma[c] = (float)sums[C + c];
}

// ---- max-pool 3x3, stride 2, padding 1 (torchvision resnet stem), NHWC bf16 ------------------------------
// forward stores the arg-max tap (0..8) per output element; backward gathers: every input pixel looks at the
// <= 4 windows that contain it (no atomics).
__global__ void __launch_bounds__(kT) maxpool_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                                                         uint2* __restrict__ idx, int N, int H, int W, int CV, int HO,
                                                         int WO) {
    const long long total = (long long)N * HO * WO * CV;
    for (long long o = (long long)blockIdx.x * kT + threadIdx.x; o < total; o += (long long)gridDim.x * kT) {
        const int cv = (int)(o % CV);
        long long r = o / CV;
        const int wo = (int)(r % WO); r /= WO;
        const int ho = (int)(r % HO);
        const int n = (int)(r / HO);
        float best[8];
        unsigned char bi[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; bi[i] = 0; }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int h = 2 * ho - 1 + dy;
            if (h < 0 || h >= H) continue;
#pragma unroll
            for (int dxx = 0; dxx < 3; ++dxx) {
                const int w = 2 * wo - 1 + dxx;
                if (w < 0 || w >= W) continue;
                float f[8];
                unpack8(x[(((long long)n * H + h) * W + w) * CV + cv], f);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (f[i] > best[i] || f[i] != f[i]) { best[i] = f[i]; bi[i] = (unsigned char)(dy * 3 + dxx); }
            }
        }
        y[o] = pack8(best);
        uint2 pk;
        pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
        pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
        idx[o] = pk;
    }
}

// one CTA per input row (n, h): the two candidate window rows are uniform for the CTA, the per-thread index math is
// shifts (CVT = compile-time channel-vector count; 0 = run time)
template <int CVT>
__global__ void __launch_bounds__(kT) maxpool_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx,
                                                         uint4* __restrict__ dx, int N, int H, int W, int CVr, int HO,
                                                         int WO) {
    const int CV = CVT ? CVT : CVr;
    const long long nrows = (long long)N * H;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int n = (int)(row / H), h = (int)(row - (long long)n * H);
        // window rows containing h: ho = h/2 (tap row h - (2ho-1) in {1,2}) and ho = (h+1)/2 when different (tap row 0)
        const int hoA = h >> 1, hoB = (h + 1) >> 1;
        const bool useA = hoA < HO, useB = hoB != hoA && hoB < HO;
        const int tdyA = h - (2 * hoA - 1), tdyB = h - (2 * hoB - 1);
        for (int i = threadIdx.x; i < W * CV; i += kT) {
            const int cv = CVT ? (i % CVT) : (i % CV);
            const int w = CVT ? (i / CVT) : (i / CV);
            const int woA = w >> 1, woB = (w + 1) >> 1;
            const bool wA = woA < WO, wB = woB != woA && woB < WO;
            const int tdxA = w - (2 * woA - 1), tdxB = w - (2 * woB - 1);
            float acc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                if (!(a ? useB : useA)) continue;
                const int ho = a ? hoB : hoA, tdy = a ? tdyB : tdyA;
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    if (!(b2 ? wB : wA)) continue;
                    const int wo = b2 ? woB : woA, tdx = b2 ? tdxB : tdxA;
                    const long long q = (((long long)n * HO + ho) * WO + wo) * CV + cv;
                    const uint2 pk = idx[q];
                    float g[8];
                    unpack8(dy[q], g);
                    const unsigned tap = (unsigned)(tdy * 3 + tdx);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const unsigned bsel = ((k < 4 ? pk.x : pk.y) >> (8 * (k & 3))) & 0xFFu;
                        if (bsel == tap) acc[k] += g[k];
                    }
                }
            }
            dx[(row * W + w) * CV + cv] = pack8(acc);
        }
    }
}

// ---- fused stem: y = maxpool3x3s2(relu(bn(x))) --------------------------------------------------------
// The normalised/activated 112x112 map (the largest tensor of the network) is never written: the forward reads
// the convolution output once and writes the pooled map + arg-max taps; the backward recomputes the ReLU mask from
// x and gathers the pooled gradient, once for the per-channel reductions and once to write dx.
template <int CVT>
__global__ void __launch_bounds__(kT) stem_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ mean,
                                                      const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, uint4* __restrict__ y,
                                                      uint2* __restrict__ idx, int N, int H, int W, int CVr, int HO, int WO) {
    // one CTA per output row (n, ho); requires kT % CV == 0 so that a thread keeps its channel group
    const int CV = CVT ? CVT : CVr;
    const int cv = threadIdx.x % CV;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cv * 8 + i;
        sc[i] = invstd[c] * gamma[c];
        sh[i] = beta[c] - mean[c] * sc[i];
    }
    const long long nrows = (long long)N * HO;
    for (long long row = blockIdx.x; row < nrows; row += gridDim.x) {
        const int n = (int)(row / HO), ho = (int)(row - (long long)n * HO);
        for (int i = threadIdx.x; i < WO * CV; i += kT) {
            const int wo = CVT ? (i / CVT) : (i / CV);
            float best[8];
            unsigned char bi[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) { best[k] = -INFINITY; bi[k] = 0; }
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const int h = 2 * ho - 1 + dy;
                if (h < 0 || h >= H) continue;
#pragma unroll
                for (int dxx = 0; dxx < 3; ++dxx) {
                    const int w = 2 * wo - 1 + dxx;
                    if (w < 0 || w >= W) continue;
                    float f[8];
                    unpack8(x[(((long long)n * H + h) * W + w) * CV + cv], f);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float a = fmaxf(fmaf(f[k], sc[k], sh[k]), 0.f);
                        if (a > best[k]) { best[k] = a; bi[k] = (unsigned char)(dy * 3 + dxx); }
                    }
                }
            }
            const long long o = (row * WO + wo) * CV + cv;
            y[o] = pack8(best);
            uint2 pk;
            pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
            pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
            idx[o] = pk;
        }
    }
}

// gradient reaching the ReLU output at input pixel (n,h,w), channel vector cv: sum over the <= 4 pooling windows
// whose arg-max is this pixel
__device__ __forceinline__ void stem_gather(const uint4* __restrict__ dp, const uint2* __restrict__ idx, int n, int h,
                                            int w, int cv, int CV, int HO, int WO, float* g) {
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = 0.f;
    const int ho0 = h / 2, ho1 = min(HO - 1, (h + 1) / 2);
    const int wo0 = w / 2, wo1 = min(WO - 1, (w + 1) / 2);
    for (int ho = ho0; ho <= ho1; ++ho) {
        const int tdy = h - (2 * ho - 1);
        for (int wo = wo0; wo <= wo1; ++wo) {
            const int tdx = w - (2 * wo - 1);
            const long long q = (((long long)n * HO + ho) * WO + wo) * CV + cv;
            const uint2 pk = idx[q];
            float d[8];
            unpack8(dp[q], d);
            const unsigned tap = (unsigned)(tdy * 3 + tdx);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned b = ((i < 4 ? pk.x : pk.y) >> (8 * (i & 3))) & 0xFFu;
                if (b == tap) g[i] += d[i];
            }
        }
    }
}

template <int PASS>   // 0: per-channel sums (sum g, sum g*xhat) ; 1: dx
__global__ void __launch_bounds__(kT) stem_bwd_kernel(const uint4* __restrict__ dp, const uint2* __restrict__ idx,
                                                      const uint4* __restrict__ x, const float* __restrict__ mean,
                                                      const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, double* __restrict__ sums,
                                                      uint4* __restrict__ dx, int N, int H, int W, int CV, int HO, int WO,
                                                      int C) {
    __shared__ float red[2][kT][8 + 1];
    const int tid = threadIdx.x;
    const int cv = tid % CV;
    const long long nvec = (long long)N * H * W * CV;
    const long long R = (long long)N * H * W;
    float mu[8], is[8], sc[8], sh[8], k0[8], k1[8], k2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cv * 8 + i;
        mu[i] = mean[c];
        is[i] = invstd[c];
        sc[i] = is[i] * gamma[c];
        sh[i] = beta[c] - mu[i] * sc[i];
        if (PASS == 1) {
            k0[i] = sc[i];
            k1[i] = (float)(sums[c] / (double)R);
            k2[i] = (float)(sums[C + c] / (double)R);
        }
    }
    float a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
    const long long per_cta = ((nvec + gridDim.x - 1) / gridDim.x + kT - 1) / kT * kT;
    const long long v0 = blockIdx.x * per_cta, v1 = min(nvec, v0 + per_cta);
    for (long long v = v0 + tid; v < v1; v += kT) {
        long long r = v / CV;
        const int w = (int)(r % W); r /= W;
        const int h = (int)(r % H);
        const int n = (int)(r / H);
        float fx[8], g[8];
        unpack8(ld_stream(x + v), fx);
        stem_gather(dp, idx, n, h, w, cv, CV, HO, WO, g);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (!(fmaf(fx[i], sc[i], sh[i]) > 0.f)) g[i] = 0.f;   // ReLU mask, recomputed
        }
        if (PASS == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { a0[i] += g[i]; a1[i] = fmaf(g[i], (fx[i] - mu[i]) * is[i], a1[i]); }
        } else {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = k0[i] * (g[i] - k1[i] - (fx[i] - mu[i]) * is[i] * k2[i]);
            dx[v] = pack8(o);
        }
    }
    if (PASS == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { red[0][tid][i] = a0[i]; red[1][tid][i] = a1[i]; }
        __syncthreads();
        for (int o = tid; o < 2 * CV * 8; o += kT) {
            const int which = o / (CV * 8), c = o % (CV * 8);
            const int ccv = c / 8, ci = c % 8;
            float s2 = 0.f;
            for (int r = ccv; r < kT; r += CV) s2 += red[which][r][ci];
            atomicAdd(&sums[which * C + c], (double)s2);
        }
    }
}

// ---- stem input packing: fp32 NCHW image -> bf16 NHWC, 2x2 space-to-depth of the 3-pixel zero-padded image -------
// out[n][hp][wp][c*4 + dy*2 + dx] = in[n][c][2*hp + dy - 3][2*wp + dx - 3]  (12 channels, zero-padded to 16):
// the operand layout under which conv1 (7x7/s2/p3, Cin=3) becomes a 4x4/s1 convolution with Cin=16.
// U8 = true: the input is the raw uint8 image (N,3,H,W); the reference's torchvision preprocessing (v2.ToDtype(float32,
// scale=True) -> v2.Normalize(mean, std); dataset/pytorch.py:198-204, ml/inference/ros.py:190-196) is applied here in the
// same fp32 operations (u * fp32(1/255), - mean, / std: torchvision's to_dtype_image / normalize_image), so the packed bf16 image is bit-identical to packing the host-normalised one
// while the host->device copy and this kernel's read shrink 4x.
struct PackNorm { float mean[3], std[3]; };
template <bool U8>
__global__ void __launch_bounds__(kT) stem_pack_kernel(const void* __restrict__ in_, uint4* __restrict__ out, int N, int H,
                                                       int W, int Hp, int Wp, PackNorm nm) {
    // uint8 input: a channel has only 256 possible normalised values — a shared-memory table built with exactly the fp32
    // operations above replaces 12 IEEE divisions per packed pixel (the kernel was issue-bound at 1.9 TB/s)
    __shared__ float lut[U8 ? 3 * 256 : 1];
    if (U8) {
        for (int i = threadIdx.x; i < 3 * 256; i += kT) {
            const int c = i >> 8;
            lut[i] = __fdiv_rn(__fsub_rn(__fmul_rn((float)(i & 255), (float)(1.0 / 255.0)), nm.mean[c]), nm.std[c]);
        }
        __syncthreads();
    }
    const long long total = (long long)N * Hp * Wp;
    for (long long o = (long long)blockIdx.x * kT + threadIdx.x; o < total; o += (long long)gridDim.x * kT) {
        const int wp = (int)(o % Wp);
        long long r = o / Wp;
        const int hp = (int)(r % Hp);
        const int n = (int)(r / Hp);
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long plane = ((long long)n * 3 + c) * H * W;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const int h = 2 * hp + dy - 3;
                if (h < 0 || h >= H) continue;
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const int w = 2 * wp + dx - 3;
                    if (w < 0 || w >= W) continue;
                    const long long i = plane + (long long)h * W + w;
                    if (U8) f[c * 4 + dy * 2 + dx] = lut[c * 256 + __ldg(reinterpret_cast<const unsigned char*>(in_) + i)];
                    else f[c * 4 + dy * 2 + dx] = __ldg(reinterpret_cast<const float*>(in_) + i);
                }
            }
        }
        out[2 * o] = pack8(f);
        out[2 * o + 1] = pack8(f + 8);
    }
}

inline int stream_grid(long long nvec) { return (int)min((long long)148 * 16, (nvec + kT - 1) / kT); }
// grid = (CTAs resident per SM) x (SM count) x waves: grid-stride / contiguous-chunk kernels finish in whole waves
template <typename K>
inline int wave_grid(K kernel, long long nvec, int waves) {
    int occ = 0, sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kT, 0) != cudaSuccess || occ < 1) occ = 2;
    return (int)min((long long)sms * occ * waves, (nvec + kT - 1) / kT);
}
inline bool ok_c(int C) { return C >= 8 && C % 8 == 0 && (kT % (C / 8) == 0); }

}  // namespace

extern "C" int sd_bn_stats_nhwc_bf16(const void* x, long long R, int C, double* sums /*[2][C], zeroed here*/, float eps,
                                     float momentum, float* mean, float* invstd, float* running_mean, float* running_var,
                                     void* stream) {
    if (R <= 0) return SD_OK;
    if (!x || !sums || !mean || !invstd || !ok_c(C)) return SD_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    SD_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    const long long nvec = R * (C / 8);
    static int occ_grid = 0;
    if (!occ_grid) occ_grid = wave_grid(bn_reduce_kernel<0>, 1ll << 40, 1);
    const int grid = (int)min((long long)occ_grid, (nvec + kT - 1) / kT);
    bn_reduce_kernel<0><<<grid, kT, 0, st>>>((const uint4*)x, nullptr, nullptr, nullptr, nullptr, nvec, C / 8, sums, C, nullptr,
                                             nullptr, nullptr);
    SD_LAUNCH_CHECK();
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, R, C, eps, momentum, mean, invstd, running_mean, running_var);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// mean / invstd / running statistics from per-channel sums produced elsewhere (the stem convolution's epilogue)
extern "C" int sd_bn_finalize(const double* sums, long long R, int C, float eps, float momentum, float* mean, float* invstd,
                              float* running_mean, float* running_var, void* stream) {
    if (R <= 0) return SD_OK;
    if (!sums || !mean || !invstd || C <= 0) return SD_ERR_BAD_ARG;
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sums, R, C, eps, momentum, mean, invstd, running_mean,
                                                                          running_var);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_bn_apply_nhwc_bf16(const void* x, const void* residual, const float* mean, const float* invstd,
                                     const float* gamma, const float* beta, int relu, void* y, void* relu_mask,
                                     long long R, int C, void* stream) {
    if (R <= 0) return SD_OK;
    if (!x || !y || !mean || !invstd || !gamma || !beta || !ok_c(C)) return SD_ERR_BAD_ARG;
    const long long nvec = R * (C / 8);
    bn_apply_kernel<<<stream_grid(nvec), kT, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)residual, mean, invstd,
                                                                       gamma, beta, relu, (uint4*)y, (unsigned char*)relu_mask, nvec, C / 8);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_bn_bwd2_nhwc_bf16(const void* dy, const void* dy2, const void* relu_mask, const void* x, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta_recompute, double* sums,
                                   void* dx, void* dres, float* dgamma, float* dbeta, long long R, int C, void* stream) {
    if (R <= 0) return SD_OK;
    if (!dy || !x || !mean || !invstd || !gamma || !sums || !dx || !dgamma || !dbeta || !ok_c(C)) return SD_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    SD_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    const long long nvec = R * (C / 8);
    static int occ_grid = 0, occ_grid_apply = 0;
    if (!occ_grid) {
        occ_grid = wave_grid(bn_reduce_kernel<1>, 1ll << 40, 1);
        occ_grid_apply = wave_grid(bn_bwd_apply_kernel, 1ll << 40, 4);
    }
    const int grid = (int)min((long long)occ_grid, (nvec + kT - 1) / kT);
    bn_reduce_kernel<1><<<grid, kT, 0, st>>>((const uint4*)x, (const uint4*)dy, (const unsigned char*)relu_mask, mean, invstd, nvec,
                                             C / 8, sums, C, gamma, relu_mask ? nullptr : beta_recompute, (const uint4*)dy2);
    SD_LAUNCH_CHECK();
    bn_param_grads_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, C, dgamma, dbeta);
    SD_LAUNCH_CHECK();
    bn_bwd_apply_kernel<<<(int)min((long long)occ_grid_apply, (nvec + kT - 1) / kT), kT, 0, st>>>((const uint4*)dy, (const unsigned char*)relu_mask, (const uint4*)x, mean,
                                                          invstd, gamma, sums, R, (uint4*)dx, (uint4*)dres, nvec, C / 8, C,
                                                          relu_mask ? nullptr : beta_recompute, (const uint4*)dy2);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_bn_bwd_nhwc_bf16(const void* dy, const void* relu_mask, const void* x, const float* mean,
                                   const float* invstd, const float* gamma, const float* beta_recompute, double* sums,
                                   void* dx, void* dres, float* dgamma, float* dbeta, long long R, int C, void* stream) {
    return sd_bn_bwd2_nhwc_bf16(dy, nullptr, relu_mask, x, mean, invstd, gamma, beta_recompute, sums, dx, dres, dgamma, dbeta, R,
                                C, stream);
}

extern "C" int sd_maxpool3x3s2_nhwc_bf16_fwd(const void* x, void* y, void* idx, int N, int H, int W, int C, void* stream) {
    if (N <= 0) return SD_OK;
    if (!x || !y || !idx || C % 8 != 0) return SD_ERR_BAD_ARG;
    const int HO = (H + 2 - 3) / 2 + 1, WO = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * HO * WO * (C / 8);
    maxpool_fwd_kernel<<<stream_grid(total), kT, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, (uint2*)idx, N, H, W,
                                                                           C / 8, HO, WO);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_maxpool3x3s2_nhwc_bf16_bwd(const void* dy, const void* idx, void* dx, int N, int H, int W, int C,
                                             void* stream) {
    if (N <= 0) return SD_OK;
    if (!dy || !idx || !dx || C % 8 != 0) return SD_ERR_BAD_ARG;
    const int HO = (H + 2 - 3) / 2 + 1, WO = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * H * W * (C / 8);
    (void)total;
    const int grid = (int)min((long long)148 * 16, (long long)N * H);
    if (C == 64)
        maxpool_bwd_kernel<8><<<grid, kT, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint2*)idx, (uint4*)dx, N, H, W,
                                                                     C / 8, HO, WO);
    else
        maxpool_bwd_kernel<0><<<grid, kT, 0, (cudaStream_t)stream>>>((const uint4*)dy, (const uint2*)idx, (uint4*)dx, N, H, W,
                                                                     C / 8, HO, WO);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_stem_band_supported(int H, int W, int C) { return stem_band_supported(H, W, C) ? 1 : 0; }

extern "C" int sd_stem_bn_relu_pool_nhwc_bf16_fwd(const void* x, const float* mean, const float* invstd, const float* gamma,
                                                  const float* beta, void* y, void* idx, int N, int H, int W, int C,
                                                  void* stream) {
    if (N <= 0) return SD_OK;
    if (!x || !mean || !invstd || !gamma || !beta || !y || !idx || !ok_c(C)) return SD_ERR_BAD_ARG;
    if (stem_band_supported(H, W, C))   // TMA-staged row bands (stem_band.cu)
        return stem_band_fwd(x, mean, invstd, gamma, beta, y, idx, N, H, W, (cudaStream_t)stream);
    const int HO = (H + 2 - 3) / 2 + 1, WO = (W + 2 - 3) / 2 + 1;
    const long long total = (long long)N * HO * WO * (C / 8);
    (void)total;
    const int grid = (int)min((long long)148 * 16, (long long)N * HO);
    if (C == 64)
        stem_fwd_kernel<8><<<grid, kT, 0, (cudaStream_t)stream>>>((const uint4*)x, mean, invstd, gamma, beta, (uint4*)y,
                                                                  (uint2*)idx, N, H, W, C / 8, HO, WO);
    else
        stem_fwd_kernel<0><<<grid, kT, 0, (cudaStream_t)stream>>>((const uint4*)x, mean, invstd, gamma, beta, (uint4*)y,
                                                                  (uint2*)idx, N, H, W, C / 8, HO, WO);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_stem_bn_relu_pool_nhwc_bf16_bwd2(const void* dpool, const void* idx, const void* x, const void* y_pooled,
                                                   const float* mean,
                                                  const float* invstd, const float* gamma, const float* beta, double* sums,
                                                  void* dx, float* dgamma, float* dbeta, int N, int H, int W, int C,
                                                  void* stream) {
    if (N <= 0) return SD_OK;
    if (!dpool || !idx || !x || !mean || !invstd || !gamma || !beta || !sums || !dx || !dgamma || !dbeta || !ok_c(C))
        return SD_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int HO = (H + 2 - 3) / 2 + 1, WO = (W + 2 - 3) / 2 + 1;
    SD_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, st));
    if (stem_band_supported(H, W, C)) {   // pooled-domain reductions + 2x2-block routing for dx (stem_band.cu)
        int rc = stem_band_bwd(dpool, idx, x, y_pooled, mean, invstd, gamma, beta, sums, nullptr, N, H, W, 0, st);
        if (rc != SD_OK) return rc;
        bn_param_grads_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, C, dgamma, dbeta);
        SD_LAUNCH_CHECK();
        return stem_band_bwd(dpool, idx, x, y_pooled, mean, invstd, gamma, beta, sums, dx, N, H, W, 1, st);
    }
    const long long nvec = (long long)N * H * W * (C / 8);
    const int grid = (int)min((long long)148 * 8, (nvec + kT - 1) / kT);
    stem_bwd_kernel<0><<<grid, kT, 0, st>>>((const uint4*)dpool, (const uint2*)idx, (const uint4*)x, mean, invstd, gamma, beta,
                                            sums, nullptr, N, H, W, C / 8, HO, WO, C);
    SD_LAUNCH_CHECK();
    bn_param_grads_kernel<<<ceil_div(C, 128), 128, 0, st>>>(sums, C, dgamma, dbeta);
    SD_LAUNCH_CHECK();
    stem_bwd_kernel<1><<<grid, kT, 0, st>>>((const uint4*)dpool, (const uint2*)idx, (const uint4*)x, mean, invstd, gamma, beta,
                                            sums, (uint4*)dx, N, H, W, C / 8, HO, WO, C);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_stem_bn_relu_pool_nhwc_bf16_bwd(const void* dpool, const void* idx, const void* x, const float* mean,
                                                  const float* invstd, const float* gamma, const float* beta, double* sums,
                                                  void* dx, float* dgamma, float* dbeta, int N, int H, int W, int C,
                                                  void* stream) {
    return sd_stem_bn_relu_pool_nhwc_bf16_bwd2(dpool, idx, x, nullptr, mean, invstd, gamma, beta, sums, dx, dgamma, dbeta, N, H, W,
                                               C, stream);
}

extern "C" int sd_stem_pack_s2d_bf16(const float* images, void* out, int N, int H, int W, void* stream) {
    if (N <= 0) return SD_OK;
    if (!images || !out || (H & 1) || (W & 1)) return SD_ERR_BAD_ARG;
    const int Hp = (H + 6) / 2, Wp = (W + 6) / 2;
    const long long total = (long long)N * Hp * Wp;
    stem_pack_kernel<false><<<stream_grid(total), kT, 0, (cudaStream_t)stream>>>(images, (uint4*)out, N, H, W, Hp, Wp, PackNorm{});
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_stem_pack_s2d_u8(const void* images_u8, void* out, int N, int H, int W, float mean0, float mean1, float mean2,
                                   float std0, float std1, float std2, void* stream) {
    if (N <= 0) return SD_OK;
    if (!images_u8 || !out || (H & 1) || (W & 1) || std0 == 0.f || std1 == 0.f || std2 == 0.f) return SD_ERR_BAD_ARG;
    const int Hp = (H + 6) / 2, Wp = (W + 6) / 2;
    const long long total = (long long)N * Hp * Wp;
    PackNorm nm{{mean0, mean1, mean2}, {std0, std1, std2}};
    stem_pack_kernel<true><<<stream_grid(total), kT, 0, (cudaStream_t)stream>>>(images_u8, (uint4*)out, N, H, W, Hp, Wp, nm);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
