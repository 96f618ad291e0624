// Band kernels for the ResNet stem after conv1 (bf16 NHWC, 64 channels): y = maxpool3x3s2(relu(bn(x))) forward and its
// backward, organised around ROW BANDS of the 112x112x64 convolution output (4.1 GB at bs=256 — the largest tensor of
// the network), so that every byte of that map moves exactly once per pass.
//
//   forward   one CTA per (image, pair of pooled rows): the 5 contiguous input rows (<= 70 KB) arrive in shared memory
//             with ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx); the 9-tap max / arg-max runs out of
//             shared memory.  Replaces a kernel that fetched every tap from global memory (2.2 TB/s).
//   backward  thread = (2x2 block of input pixels, 8 channels): the four pooling windows that reach the block route
//             their gradient into register accumulators with nine predicated adds per channel; the ReLU mask is
//             recomputed from x; PASS 0 accumulates sum g and sum g*x per channel (registers, one double atomic per
//             CTA and channel), PASS 1 writes dx.  The activated map and its gradient are never materialised.
//
// Reference semantics: torchvision ResNet stem bn1 -> relu -> maxpool(3, 2, 1) under train()
// (ml/model/encoder/image.py:46-52); arg-max taps are stored as dy*3+dx like sd_maxpool3x3s2_nhwc_bf16_fwd.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/sd_b200.h"

using namespace sd;
using namespace sdtc;

namespace {

constexpr int kT = 256;

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
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

// ---- forward ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT, 3) stem_fwd_band_kernel(const uint4* __restrict__ x, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, uint4* __restrict__ y,
                                                              uint2* __restrict__ idx, int N, int H, int W, int HO, int WO) {
    extern __shared__ __align__(128) uint8_t rows_sm[];   // [<=5 input rows][W][64 bf16]
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, cv = tid & 7;
    const int nb = (HO + 1) >> 1;
    const int n = blockIdx.x / nb, jb = blockIdx.x - n * nb;
    const int ho0 = 2 * jb, npr = min(2, HO - ho0);
    const int h_lo = max(0, 2 * ho0 - 1), h_hi = min(H - 1, 2 * (ho0 + npr - 1) + 1);
    const uint32_t bytes = (uint32_t)(h_hi - h_lo + 1) * (uint32_t)W * 128u;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar, bytes);
        tma_bulk_g2s(rows_sm, x + ((long long)n * H + h_lo) * W * 8, bytes, &bar);
    }
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cv * 8 + i;
        sc[i] = invstd[c] * gamma[c];
        sh[i] = beta[c] - mean[c] * sc[i];
    }
    uint32_t sgn[4];   // sign-bit flips that make a = x*sc + sh increasing in the flipped x, per packed channel pair
#pragma unroll
    for (int k = 0; k < 4; ++k) sgn[k] = (sc[2 * k] < 0.f ? 0x00008000u : 0u) | (sc[2 * k + 1] < 0.f ? 0x80000000u : 0u);
    mbar_wait(&bar, 0);
    const int per_row = WO * 8;
    for (int i = tid; i < npr * per_row; i += kT) {
        const int pr = i >= per_row ? 1 : 0;
        const int wo = (i - pr * per_row) >> 3;
        const int ho = ho0 + pr;
        // arg-max over the window of a = x*sc + sh.  a is monotone in x (increasing for sc >= 0, decreasing otherwise), so
        // the 9-tap max / arg-max runs on the RAW bf16 pairs with the sign bit flipped where sc < 0 — packed compare
        // (__hgt2_mask) + two bit-selects per tap and channel PAIR instead of unpack, fma, compare and two selects per
        // channel; the affine map and the ReLU are applied once to the winner.  (When every tap is <= 0 the winner's
        // gradient is masked by the recomputed ReLU mask in the backward pass anyway.)
        uint32_t bestp[4], bip[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { bestp[k] = 0xFF80FF80u; bip[k] = 0u; }   // (-inf, -inf)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int h = 2 * ho - 1 + dy;
            if (h < 0 || h >= H) continue;
            const uint8_t* rowp = rows_sm + (size_t)(h - h_lo) * W * 128 + cv * 16;
#pragma unroll
            for (int dxx = 0; dxx < 3; ++dxx) {
                const int w = 2 * wo - 1 + dxx;
                if (w < 0 || w >= W) continue;
                const uint4 u = *reinterpret_cast<const uint4*>(rowp + w * 128);
                const uint32_t xw[4] = {u.x, u.y, u.z, u.w};
                const uint32_t tap2 = (uint32_t)(dy * 3 + dxx) * 0x00010001u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint32_t xs = xw[k] ^ sgn[k];
                    const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&xs),
                                                   *reinterpret_cast<const __nv_bfloat162*>(&bestp[k]));
                    bestp[k] = (xs & m) | (bestp[k] & ~m);
                    bip[k] = (tap2 & m) | (bip[k] & ~m);
                }
            }
        }
        float best[8];
        unsigned bi[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t xr = bestp[k] ^ sgn[k];
            best[2 * k] = fmaxf(fmaf(__uint_as_float(xr << 16), sc[2 * k], sh[2 * k]), 0.f);
            best[2 * k + 1] = fmaxf(fmaf(__uint_as_float(xr & 0xFFFF0000u), sc[2 * k + 1], sh[2 * k + 1]), 0.f);
            bi[2 * k] = bip[k] & 0xFFu;
            bi[2 * k + 1] = (bip[k] >> 16) & 0xFFu;
        }
        const long long o = (((long long)n * HO + ho) * WO + wo) * 8 + cv;
        y[o] = pack8(best);
        uint2 pk;
        pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
        pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
        idx[o] = pk;
    }
}

// predicated accumulate (ISETP + @p FADD: one ALU-pipe instruction per routed tap instead of compare + select + add;
// written in PTX because the C form is compiled into a divergent jump tree)
__device__ __forceinline__ void add_if_eq(float& acc, unsigned tap, unsigned code, float g) {
    asm("{\n\t.reg .pred q;\n\tsetp.eq.u32 q, %1, %2;\n\t@q add.f32 %0, %0, %3;\n\t}" : "+f"(acc) : "r"(tap), "r"(code), "f"(g));
}
// ReLU-masked accumulation of the two BatchNorm-backward reductions: if (act > 0) { a0 += g; a1 += g * x; }
__device__ __forceinline__ void acc_if_pos(float& a0, float& a1, float act, float g, float x) {
    asm("{\n\t.reg .pred q;\n\tsetp.gt.f32 q, %2, 0f00000000;\n\t@q add.f32 %0, %0, %3;\n\t@q fma.rn.f32 %1, %3, %4, %1;\n\t}"
        : "+f"(a0), "+f"(a1) : "f"(act), "f"(g), "f"(x));
}
// v += (act > 0) ? s * g : 0
__device__ __forceinline__ void fma_if_pos(float& v, float act, float s, float g) {
    asm("{\n\t.reg .pred q;\n\tsetp.gt.f32 q, %1, 0f00000000;\n\t@q fma.rn.f32 %0, %2, %3, %0;\n\t}" : "+f"(v) : "f"(act), "f"(s), "f"(g));
}

// ---- backward reductions in the POOLED domain -----------------------------------------------------------------
// sum_pixels g*mask and sum_pixels g*mask*xhat only involve the arg-max pixel of every pooling window:
//   g[pixel] = sum of dp over the windows that selected it, mask[pixel] = (y_window > 0), and at the arg-max
//   gamma*xhat + beta = y_window when positive, so   sum g*mask*xhat = sum_windows dp * (y > 0) * (y - beta) / gamma.
// The per-channel reductions therefore need the pooled gradient and the pooled OUTPUT (both 1/4 of the map: 2 GB instead
// of 5.6 GB at bs=256) and no tap routing at all.  y is bf16-rounded, i.e. xhat carries an unbiased relative error of
// 2^-9 (|xhat| + |beta/gamma|) per element, which averages out over the millions of windows of a channel; when some
// channel has |beta/gamma| > 16 or gamma ~ 0 the exact kernel (stem_bwd_block_kernel<0>) runs instead: both kernels
// evaluate the same predicate on the parameters and exactly one of them accumulates.
__device__ __forceinline__ bool pooled_sums_ok(const float* __restrict__ gamma, const float* __restrict__ beta) {
    const int lane = threadIdx.x & 31;
    bool ok = true;
#pragma unroll
    for (int c = lane; c < 64; c += 32) {
        const float g = fabsf(gamma[c]), b = fabsf(beta[c]);
        ok = ok && g > 1e-6f && b <= 16.f * g;
    }
    return __all_sync(0xffffffffu, ok);
}

__global__ void __launch_bounds__(kT, 3) stem_bwd_sums_pooled_kernel(const uint4* __restrict__ dp, const uint4* __restrict__ y,
                                                                     const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                     double* __restrict__ sums, long long nvec) {
    __shared__ float red[2 * kT * 9];
    if (!pooled_sums_ok(gamma, beta)) return;
    const int tid = threadIdx.x, cv = tid & 7;
    float ig[8], be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        ig[i] = 1.0f / gamma[cv * 8 + i];
        be[i] = beta[cv * 8 + i];
    }
    float a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;
    const long long per_cta = ((nvec + gridDim.x - 1) / gridDim.x + kT - 1) / kT * kT;   // multiple of 8: cv stays fixed
    const long long v0 = blockIdx.x * per_cta, v1 = min(nvec, v0 + per_cta);
    for (long long v = v0 + tid; v < v1; v += 4 * kT) {
        uint4 ud[4], uy[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long w = v + (long long)u * kT;
            if (w < v1) {
                ud[u] = ld_stream(dp + w);
                uy[u] = ld_stream(y + w);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (v + (long long)u * kT >= v1) break;
            float g[8], fy[8];
            unpack8(ud[u], g);
            unpack8(uy[u], fy);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float gm = fy[i] > 0.f ? g[i] : 0.f;
                a0[i] += gm;
                a1[i] = fmaf(gm, (fy[i] - be[i]) * ig[i], a1[i]);   // xhat at the arg-max
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        red[tid * 9 + i] = a0[i];
        red[(kT + tid) * 9 + i] = a1[i];
    }
    __syncthreads();
    if (tid < 128) {
        const int which = tid >> 6, c = tid & 63;
        float s = 0.f;
        for (int r = c >> 3; r < kT; r += 8) s += red[(which * kT + r) * 9 + (c & 7)];
        atomicAdd(&sums[which * 64 + c], (double)s);
    }
}

// ---- backward --------------------------------------------------------------------------------------------------
// Thread = (2x2 block of input pixels, 8-channel vector).  The block is reached by exactly four pooling windows
// (ho in {j, j+1}, wo in {a, a+1} for block rows 2j,2j+1 and columns 2a,2a+1) and by nine (window, tap) pairs in
// total, so the pooled gradient is routed with nine predicated adds per channel into 32 register accumulators — a
// per-pixel gather would examine 4 windows x 9 taps for EVERY pixel.  No shared memory, no barriers; all twelve
// global loads of an item (4 x vectors, 4 pooled gradients, 4 tap words) are in flight together.
template <int PASS>   // 0: per-channel sums (sum g, sum g*xhat) ; 1: dx
__global__ void __launch_bounds__(kT, 2) stem_bwd_block_kernel(const uint4* __restrict__ dp, const uint2* __restrict__ idx,
                                                               const uint4* __restrict__ x, const float* __restrict__ mean,
                                                               const float* __restrict__ invstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, double* __restrict__ sums,
                                                               uint4* __restrict__ dx, int N, int H, int W, int HO, int WO,
                                                               int pooled_sums_launched) {
    __shared__ float red[PASS == 0 ? 2 * kT * 9 : 1];
    if (PASS == 0 && pooled_sums_launched && pooled_sums_ok(gamma, beta)) return;   // stem_bwd_sums_pooled_kernel did it
    const int tid = threadIdx.x, cv = tid & 7;
    const double invR = 1.0 / (double)((long long)N * H * W);
    float sc[8], sh[8], kx[8], kc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = cv * 8 + i;
        const float mu = mean[c], is = invstd[c];
        sc[i] = is * gamma[c];
        sh[i] = beta[c] - mu * sc[i];
        if (PASS == 1) {   // dx = sc*g + kx*x + kc  (see bn_bwd_apply_kernel)
            const float k1 = (float)(sums[c] * invR), k2 = (float)(sums[64 + c] * invR);
            kx[i] = -sc[i] * is * k2;
            kc[i] = -sc[i] * k1 - kx[i] * mu;
        }
    }
    float a0[8], a1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a0[i] = a1[i] = 0.f;

    const int AP = (W + 1) >> 1, JB = (H + 1) >> 1;   // column pairs per row, row pairs per image
    const unsigned per_band = (unsigned)AP * 8u;
    const unsigned total = (unsigned)N * (unsigned)JB * per_band;   // < 2^31 (checked by the launcher)
    for (unsigned item = blockIdx.x * kT + tid; item < total; item += gridDim.x * kT) {   // stride % 8 == 0: cv is fixed
        const unsigned bandl = item / per_band;
        const int a = (int)((item - bandl * per_band) >> 3);
        const int n = (int)(bandl / (unsigned)JB), j = (int)(bandl - (unsigned)n * (unsigned)JB);
        const int h0 = 2 * j, w0 = 2 * a;
        uint4 ux[2][2];
        bool vx[2][2];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                vx[r][c] = h0 + r < H && w0 + c < W;
                if (vx[r][c]) ux[r][c] = ld_stream(x + (((long long)n * H + h0 + r) * W + w0 + c) * 8 + cv);
            }
        uint4 ud[2][2];
        uint2 ui[2][2];
#pragma unroll
        for (int pr = 0; pr < 2; ++pr)
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
                ud[pr][pc] = make_uint4(0u, 0u, 0u, 0u);
                ui[pr][pc] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);   // tap 255 matches nothing
                if (j + pr < HO && a + pc < WO) {
                    const long long q = (((long long)n * HO + j + pr) * WO + a + pc) * 8 + cv;
                    ud[pr][pc] = __ldg(dp + q);
                    ui[pr][pc] = __ldg(idx + q);
                }
            }
        float acc[2][2][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[0][0][k] = acc[0][1][k] = acc[1][0][k] = acc[1][1][k] = 0.f;
        // window (ho, wo) tap (tdy, tdx) is pixel (2ho-1+tdy, 2wo-1+tdx); tap code = tdy*3 + tdx
#pragma unroll
        for (int pr = 0; pr < 2; ++pr)
#pragma unroll
            for (int pc = 0; pc < 2; ++pc) {
                float g[8];
                unpack8(ud[pr][pc], g);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const unsigned tap = ((k < 4 ? ui[pr][pc].x : ui[pr][pc].y) >> (8 * (k & 3))) & 0xFFu;
                    if (pr == 0 && pc == 0) {          // rows tdy-1, columns tdx-1
                        add_if_eq(acc[0][0][k], tap, 4u, g[k]);
                        add_if_eq(acc[0][1][k], tap, 5u, g[k]);
                        add_if_eq(acc[1][0][k], tap, 7u, g[k]);
                        add_if_eq(acc[1][1][k], tap, 8u, g[k]);
                    } else if (pr == 0 && pc == 1) {   // tdx = 0 -> column 1
                        add_if_eq(acc[0][1][k], tap, 3u, g[k]);
                        add_if_eq(acc[1][1][k], tap, 6u, g[k]);
                    } else if (pr == 1 && pc == 0) {   // tdy = 0 -> row 1
                        add_if_eq(acc[1][0][k], tap, 1u, g[k]);
                        add_if_eq(acc[1][1][k], tap, 2u, g[k]);
                    } else {
                        add_if_eq(acc[1][1][k], tap, 0u, g[k]);
                    }
                }
            }
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (!vx[r][c]) continue;
                float fx[8];
                unpack8(ux[r][c], fx);
                const float* g = acc[r][c];
                // ReLU mask recomputed from x: act = x*sc + sh > 0
                if (PASS == 0) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) acc_if_pos(a0[k], a1[k], fmaf(fx[k], sc[k], sh[k]), g[k], fx[k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float act = fmaf(fx[k], sc[k], sh[k]);
                        fx[k] = fmaf(kx[k], fx[k], kc[k]);
                        fma_if_pos(fx[k], act, sc[k], g[k]);
                    }
                    dx[(((long long)n * H + h0 + r) * W + w0 + c) * 8 + cv] = pack8(fx);
                }
            }
    }
    if (PASS == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = cv * 8 + i;
            red[tid * 9 + i] = a0[i];
            red[(kT + tid) * 9 + i] = invstd[c] * fmaf(-mean[c], a0[i], a1[i]);   // sum g*xhat from the raw sum g*x
        }
        __syncthreads();
        if (tid < 128) {
            const int which = tid >> 6, c = tid & 63;
            float s = 0.f;
            for (int r = c >> 3; r < kT; r += 8) s += red[(which * kT + r) * 9 + (c & 7)];
            atomicAdd(&sums[which * 64 + c], (double)s);
        }
    }
}

}  // namespace

namespace sd {

bool stem_band_supported(int H, int W, int C) { return C == 64 && W >= 1 && W <= 112 && H >= 1; }

int stem_band_fwd(const void* x, const float* mean, const float* invstd, const float* gamma, const float* beta, void* y,
                  void* idx, int N, int H, int W, cudaStream_t st) {
    const int HO = (H + 2 - 3) / 2 + 1, WO = (W + 2 - 3) / 2 + 1;
    const size_t smem = (size_t)5 * W * 128;
    static bool configured = false;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(stem_fwd_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 112 * 128));
        configured = true;
    }
    const long long grid = (long long)N * ((HO + 1) / 2);
    stem_fwd_band_kernel<<<(unsigned)grid, kT, smem, st>>>((const uint4*)x, mean, invstd, gamma, beta, (uint4*)y, (uint2*)idx,
                                                          N, H, W, HO, WO);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

int stem_band_bwd(const void* dpool, const void* idx, const void* x, const void* y_pooled, const float* mean,
                  const float* invstd, const float* gamma, const float* beta, double* sums, void* dx, int N, int H, int W,
                  int pass, cudaStream_t st) {
    const int HO = (H + 2 - 3) / 2 + 1, WO = (W + 2 - 3) / 2 + 1;
    static int grid0 = 0, grid1 = 0;
    if (!grid0) {
        int occ0 = 0, occ1 = 0, sms = 148, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ0, stem_bwd_block_kernel<0>, kT, 0) != cudaSuccess || occ0 < 1) occ0 = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, stem_bwd_block_kernel<1>, kT, 0) != cudaSuccess || occ1 < 1) occ1 = 1;
        grid0 = sms * occ0 * 4;
        grid1 = sms * occ1 * 4;
    }
    const long long items = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * 8;
    if (items >= (1ll << 31)) return SD_ERR_UNSUPPORTED;
    const long long ctas = (items + kT - 1) / kT;
    if (pass == 0) {
        if (y_pooled) {   // reductions from the pooled tensors; the exact kernel below returns at once unless the parameters
                          // fail pooled_sums_ok
            static int gridp = 0;
            if (!gridp) {
                int occ = 0, sms = 148, dev = 0;
                cudaGetDevice(&dev);
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stem_bwd_sums_pooled_kernel, kT, 0) != cudaSuccess || occ < 1) occ = 2;
                gridp = sms * occ;
            }
            const long long nvec = (long long)N * HO * WO * 8;
            stem_bwd_sums_pooled_kernel<<<(int)min((long long)gridp, (nvec + kT - 1) / kT), kT, 0, st>>>(
                (const uint4*)dpool, (const uint4*)y_pooled, mean, invstd, gamma, beta, sums, nvec);
            SD_LAUNCH_CHECK();
        }
        stem_bwd_block_kernel<0><<<(int)min((long long)grid0, ctas), kT, 0, st>>>(
            (const uint4*)dpool, (const uint2*)idx, (const uint4*)x, mean, invstd, gamma, beta, sums, nullptr, N, H, W, HO, WO,
            y_pooled ? 1 : 0);
    } else {
        stem_bwd_block_kernel<1><<<(int)min((long long)grid1, ctas), kT, 0, st>>>(
            (const uint4*)dpool, (const uint2*)idx, (const uint4*)x, mean, invstd, gamma, beta, sums, (uint4*)dx, N, H, W, HO, WO, 0);
    }
    SD_LAUNCH_CHECK();
    return SD_OK;
}

}  // namespace sd
