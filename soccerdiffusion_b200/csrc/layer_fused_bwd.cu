// Backward of one pre-LN transformer ENCODER layer as ONE kernel (data gradients) on tcgen05 tensor cores, weights
// streamed by TMA; companion of layer_fused_fwd.cu (same tile / thread mapping, same dropout streams):
//
//     y  = x1 + Drop3(W2 Drop2(GELU(W1 LN2 x1 + b1)) + b2)          torch/nn/modules/transformer.py:944-950
//     x1 = x  + Drop1(Wout MHA(LN1 x) + bout)                        reference call sites: ml/model/encoder/base.py:29-53
//
// One CTA (512 threads) per 128-row tile.  The fp32 gradient of the residual stream enters as dy, lives in registers and
// leaves as dx; in between, per tile:
//   FFN:  g2 = dy*mask3 ; hpre recomputed (LN2(x1) saved by the forward, one MMA) ; dhact = g2 W2 ; dhpre = dhact * mask2 *
//         gelu'(hpre) ; d(LN2 out) = dhpre W1 ; LayerNorm backward (statistics recomputed from the saved x1)
//   MHA:  g1 = dx1*mask1 ; dO = g1 Wout ; Q, K, V recomputed (LN1(x) saved by the forward) ; per head: S = Q K^T, dP = dO V^T
//         -> softmax recomputed, dS = P (dP*mask - delta) / sqrt(dh) -> dV = P'^T dO, dK = dS^T Q, dQ = dS K (full-width MMAs,
//         the head's columns are kept) written IN PLACE over the head's columns of V, K, Q ; d(LN1 out) = dQ Wq + dK Wk + dV Wv ;
//         LayerNorm backward.
// Every operand tile is used both K-major and MN-major (tc_common.cuh), so no transposed copies exist; dgrad GEMMs read
// the [out][in] weight tiles as MN-major B operands.
//
// The weight gradients are NOT computed here: the kernel stores the bf16 gradient activations g2, dhpre, g1, dq|dk|dv
// that, together with the forward's saved activations, feed the TMA-fed weight-gradient GEMM (gemm_tma.cu).  LayerNorm
// weight/bias gradients are reduced in-kernel (warp transpose-reduce, one atomic per column and CTA).
#include "layer_common.cuh"
#include "../../include/sd_b200.h"

using namespace sdlf;

namespace {

constexpr int P0 = 0, P1 = 2 * LTILE, P2 = 4 * LTILE, P3 = 6 * LTILE, P4 = 8 * LTILE, P5 = 10 * LTILE;
constexpr int OFF_U = 12 * LTILE;
constexpr int SMEM_DYN = 13 * LTILE + 1024;
constexpr float U_VAL = 32.0f;   // see layer_fused_fwd.cu: softmax mask by MMA

enum { BW_W1 = 0, BW_W2, BW_OUT, BW_Q, BW_K, BW_V, BW_Q2, BW_K2, BW_V2, B_F2, B_F4, B_A2, B_X, B_SP0, B_H0 = B_SP0 + 8,
       NBAR = B_H0 + 8 };

struct EncBwdParams {
    const float* dy;
    float* dx;
    const float* x;
    const float* x1;
    const uint4* xn1;
    const uint4* xn2;
    uint4 *g2, *dhpre, *g1, *dqkv;
    float *g_n1_w, *g_n1_b, *g_n2_w, *g_n2_b;
    const float *in_b, *l1_b, *n1_w, *n2_w;
    int B, S, spt, H, dh;
    int w_row0, w_row_ffn;
    uint32_t ffn_stream;
    Dropout drop;
};

// v[j] = value at (row = lane, column j of the warp's 32-column block)  ->  sum over the warp's 32 rows of column `lane`
__device__ __forceinline__ float warp_colsum32(float* v, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; ++j) {
            const float keep = up ? v[j + s] : v[j];
            const float send = up ? v[j] : v[j + s];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// 64 bytes of a saved bf16 activation row -> the thread's four chunks of an operand (two [128][64] swizzled tiles)
__device__ __forceinline__ void copy_row32(uint8_t* tiles, const Lane& L, const uint4* g, bool valid) {
    uint8_t* tile = tiles + (L.cq >> 1) * LTILE;
    const int c0 = (L.cq & 1) * 4;
    uint4 u[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) u[c] = valid ? g[c] : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(tile + sw128_chunk_off(L.row, c0 + c)) = u[c];
}

// LayerNorm backward on the register fragment: gr += rstd * (gg - mean(gg) - xh * mean(gg * xh)), gg = dxn * gamma;
// column sums of dxn and dxn * xh (the gradients of beta and gamma) go to colacc[0] / colacc[1]
__device__ __forceinline__ void ln_backward(float* gr, float* dxn, const float* xh, float rstd, const float* gamma_g,
                                            float* colacc_w, float* colacc_b, float* red0, float* red1, const Lane& L) {
    float ga[32], t[32];
    ldg32(gamma_g + L.col0, ga);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float gg = dxn[j] * ga[j];
        s1 += gg;
        s2 = fmaf(gg, xh[j], s2);
    }
    s1 = row_sum(s1, red0, L.tid) * (1.0f / 128.0f);
    s2 = row_sum(s2, red1, L.tid) * (1.0f / 128.0f);
#pragma unroll
    for (int j = 0; j < 32; ++j) gr[j] += rstd * (dxn[j] * ga[j] - s1 - xh[j] * s2);
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = dxn[j] * xh[j];
    const float cw = warp_colsum32(t, L.lane);
    const float cb = warp_colsum32(dxn, L.lane);   // destroys dxn
    atomicAdd(colacc_w + L.col0 + L.lane, cw);
    atomicAdd(colacc_b + L.col0 + L.lane, cb);
}

template <bool DROP, bool SA, bool FFN>
__global__ void __launch_bounds__(LNT, 1) enc_layer_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const EncBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar[NBAR];
    __shared__ float red[2][LNT];
    __shared__ float colacc[4][128];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const Lane L;
    const int tid = L.tid, r = L.row, c0 = L.col0;

    if ((L.warp == 0 && elect_one())) {
        for (int i = 0; i < NBAR; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmW);
    }
    colacc[tid >> 7][tid & 127] = 0.f;
    if (L.warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t ACC0 = 0, ACC1 = 128, ACC2 = 256, ACC3 = 384;

    // weight matrix mi of the layer (0 Wq, 1 Wk, 2 Wv, 3 Wout, 4 W1, 5 W2) -> the tile pair at byte offset `pair`
    auto load_w = [&](int mi, int pair, int b) {
        const uint32_t dst = sbase + pair;
        const int row = mi < 4 ? p.w_row0 + 128 * mi : p.w_row_ffn + 128 * (mi - 4);
        mbar_arrive_expect_tx(&bar[b], 2 * LTILE);
        tma_tile_2d(dst, &tmW, 0, row, &bar[b]);
        tma_tile_2d(dst + LTILE, &tmW, 64, row, &bar[b]);
    };
    auto load_attn_w = [&]() {
        load_w(3, P5, BW_OUT);
        load_w(0, P2, BW_Q);
        load_w(1, P3, BW_K);
        load_w(2, P4, BW_V);
    };
    if ((L.warp == 0 && elect_one())) {
        if (FFN) { load_w(4, P4, BW_W1); load_w(5, P5, BW_W2); } else load_attn_w();
    }

    const int S = p.S, H = p.H, dh = p.dh;
    const int samp0 = blockIdx.x * p.spt;
    const int nsamp = min(p.spt, p.B - samp0);
    const int Rv = nsamp * S;
    const long long grow = (long long)samp0 * S + r;
    const bool rv = r < Rv;
    const long long goff = grow * 128 + c0;
    const uint64_t dseed = DROP ? p.drop.resolve() : 0ull;
    const uint32_t id_kk = instr_desc_bf16(128, 128, 0, 0);   // A K-major, B K-major
    const uint32_t id_km = instr_desc_bf16(128, 128, 0, 1);   // A K-major, B MN-major
    const uint32_t id_mm = instr_desc_bf16(128, 128, 1, 1);   // both MN-major

    // ---- FFN backward -----------------------------------------------------------------------------------------------------
    float gr[32];   // fp32 gradient of the residual stream (this thread's 32 columns of its row)
    if (rv) {
        const float4* g = reinterpret_cast<const float4*>(p.dy + goff);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = g[j];
            gr[4 * j] = t.x; gr[4 * j + 1] = t.y; gr[4 * j + 2] = t.z; gr[4 * j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) gr[j] = 0.f;
    }
    if constexpr (FFN) {
    {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            v[j] = gr[j];
            if (DROP) v[j] *= dropout_scale(dseed, p.ffn_stream + 1, (uint64_t)goff + j, p.drop.thresh, p.drop.inv_keep);
        }
        st_row32(smem + P0, L, v, rv ? p.g2 + goff / 8 : nullptr);      // g2 = dy * mask3 : A operand of dhact, saved for dW2
        copy_row32(smem + P1, L, p.xn2 + goff / 8, rv);                // LN2(x1): A operand of the hpre recomputation
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        mbar_wait(&bar[BW_W1], 0);
        mma_k_tiles(tmem + ACC0, sbase + P1, LTILE, sbase + P4, LTILE, id_kk, 2, false);       // hpre = LN2(x1) W1^T
        mbar_wait(&bar[BW_W2], 0);
        mma_a_k_b_mn(tmem + ACC1, sbase + P0, LTILE, sbase + P5, LTILE, id_km, 8, false);      // dhact = g2 W2
        mma_commit(&bar[B_F2]);
    }
    __syncwarp();
    {
        float hp[32], dh_[32], b[32];
        ldg32(p.l1_b + c0, b);
        mbar_wait(&bar[B_F2], 0);
        tc_fence_after_sync();
        ld_acc32(tmem, L, ACC0, hp);
        ld_acc32(tmem, L, ACC1, dh_);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float t = dh_[j] * gelu_fast_grad(hp[j] + b[j]);
            if (DROP) t *= dropout_scale(dseed, p.ffn_stream, (uint64_t)goff + j, p.drop.thresh, p.drop.inv_keep);
            dh_[j] = rv ? t : 0.f;
        }
        st_row32(smem + P2, L, dh_, rv ? p.dhpre + goff / 8 : nullptr);   // dhpre: A operand of d(LN2 out), saved for dW1
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        mma_a_k_b_mn(tmem + ACC2, sbase + P2, LTILE, sbase + P4, LTILE, id_km, 8, false);      // d(LN2 out) = dhpre W1
        mma_commit(&bar[B_F4]);
        if (SA) {
            mbar_wait(&bar[B_F4], 0);   // P2 (dhpre), P4 (W1), P5 (W2) are free: prefetch the attention block's weights
            load_attn_w();
        }
    }
    __syncwarp();
    {
        // LayerNorm-2 backward: statistics recomputed from the saved x1 while the MMA runs
        float xh[32], mean, rstd;
        if (rv) {
            const float4* g = reinterpret_cast<const float4*>(p.x1 + goff);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = g[j];
                xh[4 * j] = t.x; xh[4 * j + 1] = t.y; xh[4 * j + 2] = t.z; xh[4 * j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) xh[j] = 0.f;
        }
        row_stats(xh, red[0], red[1], tid, mean, rstd);
#pragma unroll
        for (int j = 0; j < 32; ++j) xh[j] = (xh[j] - mean) * rstd;
        float dxn[32];
        mbar_wait(&bar[B_F4], 0);
        tc_fence_after_sync();
        ld_acc32(tmem, L, ACC2, dxn);
        ln_backward(gr, dxn, xh, rstd, p.n2_w, colacc[0], colacc[1], red[0], red[1], L);
    }
    }   // FFN
    if constexpr (!SA) {
        if (rv) {
            float4* g = reinterpret_cast<float4*>(p.dx + goff);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = make_float4(gr[4 * j], gr[4 * j + 1], gr[4 * j + 2], gr[4 * j + 3]);
        }
    }
    // ---- attention backward -------------------------------------------------------------------------------------------------
    if constexpr (SA) {
    {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            v[j] = rv ? gr[j] : 0.f;
            if (DROP) v[j] *= dropout_scale(dseed, p.drop.stream + 1, (uint64_t)goff + j, p.drop.thresh, p.drop.inv_keep);
        }
        st_row32(smem + P0, L, v, rv ? p.g1 + goff / 8 : nullptr);      // g1 = dx1 * mask1 : A operand of dO, saved for dWout
        copy_row32(smem + P1, L, p.xn1 + goff / 8, rv);                // LN1(x): A operand of the Q, K, V recomputation
        if (L.cq == 0) {
            float u[16];
            const int s_of_r = r / S;
#pragma unroll
            for (int j = 0; j < 16; ++j) u[j] = (rv && j == s_of_r) ? U_VAL : 0.f;
            *reinterpret_cast<uint4*>(smem + OFF_U + sw128_chunk_off(r, 0)) = pack8_bf16(u);
            *reinterpret_cast<uint4*>(smem + OFF_U + sw128_chunk_off(r, 1)) = pack8_bf16(u + 8);
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        mbar_wait(&bar[BW_OUT], 0);
        mma_a_k_b_mn(tmem + ACC0, sbase + P0, LTILE, sbase + P5, LTILE, id_km, 8, false);      // dO = g1 Wout
        mbar_wait(&bar[BW_Q], 0);
        mma_k_tiles(tmem + ACC1, sbase + P1, LTILE, sbase + P2, LTILE, id_kk, 2, false);       // Q
        mbar_wait(&bar[BW_K], 0);
        mma_k_tiles(tmem + ACC2, sbase + P1, LTILE, sbase + P3, LTILE, id_kk, 2, false);       // K
        mbar_wait(&bar[BW_V], 0);
        mma_k_tiles(tmem + ACC3, sbase + P1, LTILE, sbase + P4, LTILE, id_kk, 2, false);       // V (rows = tokens)
        mma_commit(&bar[B_A2]);
    }
    __syncwarp();
    {
        float v[32], b[32];
        mbar_wait(&bar[B_A2], 0);
        tc_fence_after_sync();
        ld_acc32(tmem, L, ACC0, v);
        st_row32(smem + P5, L, v, nullptr);                              // dO
        ldg32(p.in_b + c0, b);
        ld_acc32(tmem, L, ACC1, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
        st_row32(smem + P2, L, v, nullptr);                              // Q
        ldg32(p.in_b + 128 + c0, b);
        ld_acc32(tmem, L, ACC2, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
        st_row32(smem + P3, L, v, nullptr);                              // K
        ldg32(p.in_b + 256 + c0, b);
        ld_acc32(tmem, L, ACC3, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
        st_row32(smem + P4, L, v, nullptr);                              // V
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();

    const int Kp16 = (Rv + 15) >> 4;
    auto issue_SP = [&](int h) {
        const int hoff = h * dh;
        const uint32_t t_off = (hoff >> 6) * LTILE;
        const uint64_t k_off = (uint64_t)((hoff & 63) >> 3);
        const uint64_t dq = smem_desc_k_sw128(sbase + P2 + t_off) + k_off, dk = smem_desc_k_sw128(sbase + P3 + t_off) + k_off;
        const uint64_t dv = smem_desc_k_sw128(sbase + P4 + t_off) + k_off, dg = smem_desc_k_sw128(sbase + P5 + t_off) + k_off;
        for (int j = 0; j < dh / 16; ++j) mma_bf16_ss(tmem + ACC0, dq + 2 * j, dk + 2 * j, id_kk, j > 0);   // S = Q K^T
        const uint64_t du = smem_desc_k_sw128(sbase + OFF_U);
        mma_bf16_ss(tmem + ACC0, du, du, id_kk, 1u);                                                      // + mask
        for (int j = 0; j < dh / 16; ++j) mma_bf16_ss(tmem + ACC1, dg + 2 * j, dv + 2 * j, id_kk, j > 0);   // dP = dO V^T
        mma_commit(&bar[B_SP0 + h]);
    };
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        issue_SP(0);
    }
    __syncwarp();
    const float scale = rsqrtf((float)dh);
    const float sc = scale * 1.4426950408889634f;
    const int s_idx = r / S;
    const int lo = s_idx * S;
    const int t_tok = r - lo;
#pragma unroll 1
    for (int h = 0; h < H; ++h) {
        const int hoff = h * dh;
        {
            float s[32], dp[32];
            mbar_wait(&bar[B_SP0 + h], 0);
            tc_fence_after_sync();
            ld_acc32(tmem, L, ACC0, s);
            ld_acc32(tmem, L, ACC1, dp);
            float mx = s[0];
#pragma unroll
            for (int j = 1; j < 32; ++j) mx = fmaxf(mx, s[j]);
            mx = row_max(mx, red[0], tid) * sc;
            float sum = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                s[j] = ex2_approx(fmaf(s[j], sc, -mx));
                sum += s[j];
            }
            sum = row_sum(sum, red[1], tid);
            const float inv = rv ? 1.0f / sum : 0.f;   // padding rows contribute nothing to dV / dK
            float delta = 0.f;
            uint32_t keep = 0xffffffffu;   // bit j: key c0 + j survives the attention-probability dropout
            if (DROP) {
                const uint64_t didx = (((uint64_t)(samp0 + s_idx) * H + h) * S + t_tok) * (uint64_t)S + (uint64_t)(long long)(c0 - lo);
                keep = 0u;
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    keep |= (hash_u32(dseed, p.drop.stream, didx + j) >= p.drop.thresh ? 1u : 0u) << j;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                s[j] *= inv;                                                          // P
                if (DROP) dp[j] = ((keep >> j) & 1u) ? dp[j] * p.drop.inv_keep : 0.f;   // dP * mask
                delta = fmaf(s[j], dp[j], delta);
            }
            delta = row_sum(delta, red[0], tid);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                dp[j] = s[j] * (dp[j] - delta) * scale;                               // dS
                if (DROP) s[j] = ((keep >> j) & 1u) ? s[j] * p.drop.inv_keep : 0.f;     // P' = P * mask
            }
            st_row32(smem + P0, L, s, nullptr);    // P' = P * mask   [t][m]
            st_row32(smem + P1, L, dp, nullptr);   // dS              [t][m]
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if ((L.warp == 0 && elect_one())) {
            tc_fence_after_sync();
            // One accumulation chain per loop.  (Interleaving the dV and dK chains instruction by instruction — two
            // accumulators, four descriptors per iteration — produced a wrong result in whichever chain's descriptor
            // registers were recycled first; every contiguous chain in this library is exact.)
            // The head's dh columns of the MN-major B operands are addressed by a byte offset INSIDE the 128-byte swizzle
            // row (the hardware applies the swizzle to the final address, exactly as for the k offset of K-major operands):
            // N = dh instead of a full-width N = 128 product of which 1/H would be kept.
            const uint32_t b_off = (uint32_t)((hoff >> 6) * LTILE + (hoff & 63) * 2);
            const uint32_t id_mm_h = instr_desc_bf16(128, dh, 1, 1), id_km_h = instr_desc_bf16(128, dh, 0, 1);
            for (int j = 0; j < Kp16; ++j)   // dV[m][c] = sum_t P'[t][m] dO[t][c]   (contraction over the query rows t)
                mma_bf16_ss(tmem + ACC2 + hoff, smem_desc_mn_sw128(sbase + P0, LTILE, 1024) + 128 * (uint64_t)j,
                            smem_desc_mn_sw128(sbase + P5 + b_off, LTILE, 1024) + 128 * (uint64_t)j, id_mm_h, j > 0);
            for (int j = 0; j < Kp16; ++j)   // dK[m][c] = sum_t dS[t][m] Q[t][c]
                mma_bf16_ss(tmem + ACC3 + hoff, smem_desc_mn_sw128(sbase + P1, LTILE, 1024) + 128 * (uint64_t)j,
                            smem_desc_mn_sw128(sbase + P2 + b_off, LTILE, 1024) + 128 * (uint64_t)j, id_mm_h, j > 0);
            for (int j = 0; j < Kp16; ++j)   // dQ[t][c] = sum_m dS[t][m] K[m][c]
                mma_bf16_ss(tmem + ACC1 + hoff, smem_desc_k_sw128(sbase + P1 + (j >> 2) * LTILE) + 2 * (j & 3),
                            smem_desc_mn_sw128(sbase + P3 + b_off, LTILE, 1024) + 128 * (uint64_t)j, id_km_h, j > 0);
            mma_commit(&bar[B_H0 + h]);
        }
        __syncwarp();
        mbar_wait(&bar[B_H0 + h], 0);
        tc_fence_after_sync();
        // keep the head's columns: dQ, dK, dV overwrite Q, K, V of this head (dead from here on) and go to HBM for dW_in
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int col = c0 + 16 * g;
            if (col >= hoff && col < hoff + dh) {   // warp-uniform
                uint8_t* tq = smem + (col >> 6) * LTILE;
                const int ch = (col & 63) >> 3;
                const uint32_t o0 = sw128_chunk_off(r, ch), o1 = sw128_chunk_off(r, ch + 1);
                uint4* gq = p.dqkv + (grow * 384 + col) / 8;
                float v[16];
                tmem_ld_32x16(tmem + L.tlane + ACC1 + col, v);
                uint4 a = pack8_bf16(v), b = pack8_bf16(v + 8);
                *reinterpret_cast<uint4*>(tq + P2 + o0) = a;
                *reinterpret_cast<uint4*>(tq + P2 + o1) = b;
                if (rv) { gq[0] = a; gq[1] = b; }
                tmem_ld_32x16(tmem + L.tlane + ACC3 + col, v);
                a = pack8_bf16(v); b = pack8_bf16(v + 8);
                *reinterpret_cast<uint4*>(tq + P3 + o0) = a;
                *reinterpret_cast<uint4*>(tq + P3 + o1) = b;
                if (rv) { gq[16] = a; gq[17] = b; }
                tmem_ld_32x16(tmem + L.tlane + ACC2 + col, v);
                a = pack8_bf16(v); b = pack8_bf16(v + 8);
                *reinterpret_cast<uint4*>(tq + P4 + o0) = a;
                *reinterpret_cast<uint4*>(tq + P4 + o1) = b;
                if (rv) { gq[32] = a; gq[33] = b; }
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if ((L.warp == 0 && elect_one()) && h + 1 < H) {
            tc_fence_after_sync();
            issue_SP(h + 1);
        }
        __syncwarp();
    }
    // ---- d(LN1 out) = dQ Wq + dK Wk + dV Wv ; LayerNorm-1 backward ------------------------------------------------------------
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        load_w(0, P0, BW_Q2);   // P' / dS / dO are dead
        load_w(1, P1, BW_K2);
        load_w(2, P5, BW_V2);
        mbar_wait(&bar[BW_Q2], 0);
        mma_a_k_b_mn(tmem + ACC0, sbase + P2, LTILE, sbase + P0, LTILE, id_km, 8, false);
        mbar_wait(&bar[BW_K2], 0);
        mma_a_k_b_mn(tmem + ACC0, sbase + P3, LTILE, sbase + P1, LTILE, id_km, 8, true);
        mbar_wait(&bar[BW_V2], 0);
        mma_a_k_b_mn(tmem + ACC0, sbase + P4, LTILE, sbase + P5, LTILE, id_km, 8, true);
        mma_commit(&bar[B_X]);
    }
    __syncwarp();
    {
        float xh[32], mean, rstd;
        if (rv) {
            const float4* g = reinterpret_cast<const float4*>(p.x + goff);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 t = g[j];
                xh[4 * j] = t.x; xh[4 * j + 1] = t.y; xh[4 * j + 2] = t.z; xh[4 * j + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) xh[j] = 0.f;
        }
        row_stats(xh, red[0], red[1], tid, mean, rstd);
#pragma unroll
        for (int j = 0; j < 32; ++j) xh[j] = (xh[j] - mean) * rstd;
        float dxn[32];
        mbar_wait(&bar[B_X], 0);
        tc_fence_after_sync();
        ld_acc32(tmem, L, ACC0, dxn);
        ln_backward(gr, dxn, xh, rstd, p.n1_w, colacc[2], colacc[3], red[0], red[1], L);
        if (rv) {
            float4* g = reinterpret_cast<float4*>(p.dx + goff);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = make_float4(gr[4 * j], gr[4 * j + 1], gr[4 * j + 2], gr[4 * j + 3]);
        }
    }
    }   // SA
    tc_fence_before_sync();
    __syncthreads();
    if ((tid < 256) ? FFN : SA) {
        float* dst = (tid < 128) ? p.g_n2_w : (tid < 256) ? p.g_n2_b : (tid < 384) ? p.g_n1_w : p.g_n1_b;
        atomicAdd(dst + (tid & 127), colacc[tid >> 7][tid & 127]);
    }
    if (L.warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

namespace {
template <bool DROP, bool SA, bool FFN>
int launch_bwd(const CUtensorMap& tmW, const EncBwdParams& p, int tiles, cudaStream_t st) {
    auto kernel = enc_layer_bwd_kernel<DROP, SA, FFN>;
    static bool configured = false;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DYN));
        configured = true;
    }
    kernel<<<tiles, LNT, SMEM_DYN, st>>>(tmW, p);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
}  // namespace

extern "C" int sd_enc_layer_bwd(const sd_enc_layer_bwd_desc* d, void* stream) {
    if (!d || !d->dy || !d->dx || !d->w_packed) return SD_ERR_BAD_ARG;
    const int blocks = d->blocks == 0 ? (SD_LAYER_SA | SD_LAYER_FFN) : d->blocks;
    const bool sa = (blocks & SD_LAYER_SA) != 0, ffn = (blocks & SD_LAYER_FFN) != 0;
    if ((blocks & ~(SD_LAYER_SA | SD_LAYER_FFN)) != 0) return SD_ERR_BAD_ARG;
    if (sa && (!d->x || !d->xn1 || !d->g1 || !d->dqkv || !d->g_n1_w || !d->g_n1_b || !d->in_b || !d->n1_w)) return SD_ERR_BAD_ARG;
    if (ffn && (!d->x1 || !d->xn2 || !d->g2 || !d->dhpre || !d->g_n2_w || !d->g_n2_b || !d->l1_b || !d->n2_w)) return SD_ERR_BAD_ARG;
    if (d->B <= 0) return SD_OK;
    if (!sd_enc_layer_supported(128, 128, d->S, d->H)) return SD_ERR_UNSUPPORTED;
    const int w_row_ffn = d->w_row_ffn > 0 ? d->w_row_ffn : d->w_row0 + 512;
    if (sa && (d->w_row0 < 0 || d->w_row0 + 512 > d->w_rows_total)) return SD_ERR_BAD_ARG;
    if (ffn && (w_row_ffn < 0 || w_row_ffn + 256 > d->w_rows_total)) return SD_ERR_BAD_ARG;
    if ((((uintptr_t)d->dy) | ((uintptr_t)d->dx) | ((uintptr_t)d->x) | ((uintptr_t)d->x1) | ((uintptr_t)d->xn1) | ((uintptr_t)d->xn2) |
         ((uintptr_t)d->g2) | ((uintptr_t)d->dhpre) | ((uintptr_t)d->g1) | ((uintptr_t)d->dqkv)) & 15)
        return SD_ERR_BAD_ARG;
    CUtensorMap tmW;
    if (!encode_bf16_2d(&tmW, d->w_packed, d->w_rows_total, 128, 128, 128)) return SD_ERR_UNSUPPORTED;
    EncBwdParams p;
    p.dy = d->dy; p.dx = d->dx; p.x = d->x; p.x1 = d->x1; p.xn1 = (const uint4*)d->xn1; p.xn2 = (const uint4*)d->xn2;
    p.g2 = (uint4*)d->g2; p.dhpre = (uint4*)d->dhpre; p.g1 = (uint4*)d->g1; p.dqkv = (uint4*)d->dqkv;
    p.g_n1_w = d->g_n1_w; p.g_n1_b = d->g_n1_b; p.g_n2_w = d->g_n2_w; p.g_n2_b = d->g_n2_b;
    p.in_b = d->in_b; p.l1_b = d->l1_b; p.n1_w = d->n1_w; p.n2_w = d->n2_w;
    p.B = d->B; p.S = d->S; p.spt = 128 / d->S; p.H = d->H; p.dh = 128 / d->H; p.w_row0 = d->w_row0; p.w_row_ffn = w_row_ffn;
    p.drop = make_dropout(d->dropout_p, d->dropout_seed, d->dropout_stream);
    p.ffn_stream = d->dropout_stream_ffn != 0 ? d->dropout_stream_ffn : d->dropout_stream + 2;
    const int tiles = ceil_div(d->B, p.spt);
    cudaStream_t st = (cudaStream_t)stream;
    const bool drop = p.drop.thresh != 0;
    if (sa && ffn) return drop ? launch_bwd<true, true, true>(tmW, p, tiles, st) : launch_bwd<false, true, true>(tmW, p, tiles, st);
    if (sa) return drop ? launch_bwd<true, true, false>(tmW, p, tiles, st) : launch_bwd<false, true, false>(tmW, p, tiles, st);
    return drop ? launch_bwd<true, false, true>(tmW, p, tiles, st) : launch_bwd<false, false, true>(tmW, p, tiles, st);
}
