// One pre-LN transformer ENCODER layer forward as ONE kernel on tcgen05 tensor cores (bf16 operands, fp32 accumulation
// in TMEM), weights streamed by TMA, d_model = ff = 128:
//
//     x1 = x  + Drop(OutProj(MHA(LN1 x)))          torch/nn/modules/transformer.py:944-950 (norm_first)
//     y  = x1 + Drop(W2 Drop(GELU_erf(W1 LN2 x1)))  reference call sites: ml/model/encoder/base.py:29-53
//
// One CTA (512 threads) per 128-row tile holding floor(128 / S) whole samples (S tokens each; 1 sample at S = 100,
// 12 at S = 10).  The fp32 residual stream of the tile is read from HBM once, lives in registers (layer_common.cuh
// mapping) and is written once; everything in between stays on chip:
//
//   LN1 -> bf16 operand tiles -> Q, K (rows = tokens) and V^T (rows = features: the operand roles of the V projection
//   are swapped so that P.V needs no transposed copy) -> per head S = Q K^T in TMEM -> fp32 softmax from TMEM (block-
//   diagonal mask when a tile holds several samples) -> P (bf16) -> O += P V into the head's column range of one
//   accumulator -> out-proj + bias + dropout + residual -> LN2 -> FC1 + bias + erf-GELU + dropout -> FC2 + bias +
//   dropout + residual -> y.
//
// The six 128x128 weight matrices (bf16 copies packed by sd_pack_weights_bf16) arrive through a 2-matrix shared-memory
// ring filled by TMA (cp.async.bulk.tensor, 128-byte swizzle) while earlier phases compute.  In training mode the kernel
// also stores what the fused backward kernel and the weight-gradient GEMMs read: x1 (fp32) and LN1(x), attention output,
// LN2(x1), hidden activation (bf16).
#include "layer_common.cuh"
#include "../../include/sd_b200.h"

using namespace sdlf;

namespace {

// shared-memory map (bytes from the 1024-aligned base): operand tiles of [128][64] bf16
constexpr int OFF_XN = 0;             // LN output -> (attention) P -> attention output -> LN2 output
constexpr int OFF_QS = 2 * LTILE;     // Q  [token][feature]            -> hidden activation
constexpr int OFF_KS = 4 * LTILE;     // K  [token][feature]
constexpr int OFF_VT = 6 * LTILE;     // V^T [feature][token]
constexpr int OFF_W = 8 * LTILE;      // weight ring: slot A = tiles 0,1 ; slot B = tiles 2,3
constexpr int OFF_U = 12 * LTILE;     // sample one-hot tile (softmax mask by MMA, see below)
constexpr int SMEM_DYN = 13 * LTILE + 1024;
// Softmax mask on the tensor core: row r of U holds 32 at column (sample of r inside the tile) for valid rows, zeros for
// padding rows; one extra k step S += U U^T adds 1024 to every (query, key) pair of the SAME sample.  After the row
// maximum is subtracted, keys of other samples and padding keys sit 1024 below it and their exponentials flush to
// exactly zero: the softmax loops carry no per-element mask logic (they are issue-bound).
constexpr float U_VAL = 32.0f;

enum { BW0 = 0, BQ = 6, BK, BV, BOUT, BF1, BF2, BS0, BP0 = BS0 + 8, NBAR = BP0 + 8 };

struct EncFwdParams {
    const float* x;
    float* y;
    int B, S, spt, H, dh;
    int w_row0;      // first packed-weight row of the attention block (Wq | Wk | Wv | Wout)
    int w_row_ffn;   // first packed-weight row of the feed-forward block (W1 | W2)
    uint32_t ffn_stream;   // first dropout stream of the feed-forward block (sites +0 FC1, +1 FC2)
    const float *in_b, *out_b, *l1_b, *l2_b, *n1_w, *n1_b, *n2_w, *n2_b;
    float* x1_save;
    uint4 *xn1_save, *attn_save, *xn2_save, *hact_save;
    Dropout drop;   // .stream = first dropout stream of this layer (sites +0 attention, +1 out-proj, +2 FC1, +3 FC2)
};

// SA: run the self-attention block; FFN: run the feed-forward block (an encoder layer is both; a decoder layer runs the
// two halves as separate launches around its cross-attention block)
// FFN_FIRST (inference, SA && FFN): the feed-forward block runs BEFORE the attention block, each with its own weights — the
// feed-forward half of decoder layer l fused with the self-attention half of layer l+1 (one kernel boundary less per layer in
// the tensor-core sampler's latency-bound chain)
template <bool DROP, bool SA, bool FFN, bool FFN_FIRST = false>
__global__ void __launch_bounds__(LNT, 1) enc_layer_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const EncFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar[NBAR];
    __shared__ float red[2][LNT];
    __shared__ float rsum[8][128];
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
    if (L.warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t ACC0 = 0, ACC1 = 128, ACC2 = 256, ACC3 = 384;

    // weight matrix mi (0 Wq, 1 Wk, 2 Wv, 3 Wout, 4 W1, 5 W2) -> ring slot (mi & 1)
    auto load_w = [&](int mi) {
        const uint32_t dst = sbase + OFF_W + (mi & 1) * 2 * LTILE;
        const int row = mi < 4 ? p.w_row0 + 128 * mi : p.w_row_ffn + 128 * (mi - 4);
        mbar_arrive_expect_tx(&bar[BW0 + mi], 2 * LTILE);
        tma_tile_2d(dst, &tmW, 0, row, &bar[BW0 + mi]);
        tma_tile_2d(dst + LTILE, &tmW, 64, row, &bar[BW0 + mi]);
    };
    if ((L.warp == 0 && elect_one())) {
        if (SA && !FFN_FIRST) { load_w(0); load_w(1); } else { load_w(4); load_w(5); }   // packed weights: written by a non-triggering kernel
    }
    pdl_trigger();
    pdl_wait();   // the residual stream below comes from the preceding kernel of the chain

    const int S = p.S, H = p.H, dh = p.dh;
    const int samp0 = blockIdx.x * p.spt;
    const int nsamp = min(p.spt, p.B - samp0);
    const int Rv = nsamp * S;                          // valid rows of this tile
    const long long grow = (long long)samp0 * S + r;   // global row of this thread
    const bool rv = r < Rv;
    const long long goff = grow * 128 + c0;            // this thread's first element in every [rows][128] matrix
    const uint64_t dseed = DROP ? p.drop.resolve() : 0ull;

    // ---- residual fragment + LN1 --------------------------------------------------------------------------------------
    float xr[32];
    if (rv) {
        const float4* g = reinterpret_cast<const float4*>(p.x + goff);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float4 t = g[j];
            xr[4 * j] = t.x; xr[4 * j + 1] = t.y; xr[4 * j + 2] = t.z; xr[4 * j + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) xr[j] = 0.f;
    }
    const uint32_t id128 = instr_desc_bf16(128, 128);
#pragma unroll
    for (int ph = 0; ph < 2; ++ph) {   // two phases in compile-time order: attention then feed-forward, or the reverse
    if ((ph == 0) != FFN_FIRST) {
    if constexpr (SA) {
    {
        float mean, rstd, v[32], ga[32], be[32];
        ldg32(p.n1_w + c0, ga);
        ldg32(p.n1_b + c0, be);
        row_stats(xr, red[0], red[1], tid, mean, rstd);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rv ? fmaf((xr[j] - mean) * rstd, ga[j], be[j]) : 0.f;
        st_row32(smem + OFF_XN, L, v, (rv && p.xn1_save) ? p.xn1_save + goff / 8 : nullptr);
        if (L.cq == 0) {   // one-hot sample indicator of this row: columns 0..15 of the U tile (spt <= 16)
            float u[16];
            const int s_of_r = r / p.S;
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
        mbar_wait(&bar[BW0 + 0], 0);
        mma_k_tiles(tmem + ACC0, sbase + OFF_XN, LTILE, sbase + OFF_W, LTILE, id128, 2, false);   // Q = LN1(x) Wq^T
        mma_commit(&bar[BQ]);
        mbar_wait(&bar[BW0 + 1], 0);
        mma_k_tiles(tmem + ACC1, sbase + OFF_XN, LTILE, sbase + OFF_W + 2 * LTILE, LTILE, id128, 2, false);   // K
        mma_commit(&bar[BK]);
        mbar_wait(&bar[BQ], 0);   // slot A drained
        load_w(2);
        mbar_wait(&bar[BW0 + 2], 0);
        mma_k_tiles(tmem + ACC2, sbase + OFF_W, LTILE, sbase + OFF_XN, LTILE, id128, 2, false);   // V^T = Wv LN1(x)^T
        mma_commit(&bar[BV]);
        mbar_wait(&bar[BK], 0);   // slot B drained
        load_w(3);
    }
    __syncwarp();

    // ---- Q, K, V^T epilogues: + bias -> bf16 operand tiles -------------------------------------------------------------
    {
        float v[32], b[32];
        ldg32(p.in_b + c0, b);
        mbar_wait(&bar[BQ], 0);
        tc_fence_after_sync();
        ld_acc32(tmem, L, ACC0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
        st_row32(smem + OFF_QS, L, v, nullptr);
        ldg32(p.in_b + 128 + c0, b);
        mbar_wait(&bar[BK], 0);
        tc_fence_after_sync();
        ld_acc32(tmem, L, ACC1, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += b[j];
        st_row32(smem + OFF_KS, L, v, nullptr);
        const float bv = __ldg(p.in_b + 256 + r);
        mbar_wait(&bar[BV], 0);
        tc_fence_after_sync();
        if (FFN && !FFN_FIRST && (L.warp == 0 && elect_one())) load_w(4);   // slot A drained by the V^T MMAs
        __syncwarp();
        ld_acc32(tmem, L, ACC2, v);   // row = feature r, columns = tokens
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] += bv;
        st_row32(smem + OFF_VT, L, v, nullptr);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();

    // ---- attention, head by head ----------------------------------------------------------------------------------------
    const uint32_t idS = instr_desc_bf16(128, 128);
    const uint32_t idPV = instr_desc_bf16(128, dh);
    const int Kp16 = (Rv + 15) >> 4;   // k steps of P.V (P columns beyond Rv are zero)
    auto issue_S = [&](int h) {
        const int hoff = h * dh;
        const uint64_t da = smem_desc_k_sw128(sbase + OFF_QS + (hoff >> 6) * LTILE) + ((hoff & 63) >> 3);
        const uint64_t db = smem_desc_k_sw128(sbase + OFF_KS + (hoff >> 6) * LTILE) + ((hoff & 63) >> 3);
        for (int j = 0; j < dh / 16; ++j) mma_bf16_ss(tmem + ((h & 1) ? ACC1 : ACC0), da + 2 * j, db + 2 * j, idS, j > 0);
        const uint64_t du = smem_desc_k_sw128(sbase + OFF_U);
        mma_bf16_ss(tmem + ((h & 1) ? ACC1 : ACC0), du, du, idS, 1u);   // + 1024 on same-sample pairs
        mma_commit(&bar[BS0 + h]);
    };
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        issue_S(0);
        if (H > 1) issue_S(1);
    }
    __syncwarp();
    const float sc = rsqrtf((float)dh) * 1.4426950408889634f;   // softmax scale * log2(e)
    const int s_idx = r / S;                                     // sample of this row inside the tile
    const int lo = s_idx * S;                                    // its first key
    const int t_tok = r - lo;
#pragma unroll 1
    for (int h = 0; h < H; ++h) {
        float v[32];
        mbar_wait(&bar[BS0 + h], 0);
        tc_fence_after_sync();
        ld_acc32(tmem, L, (h & 1) ? ACC1 : ACC0, v);
        float mx = v[0];
#pragma unroll
        for (int j = 1; j < 32; ++j) mx = fmaxf(mx, v[j]);
        mx = row_max(mx, red[0], tid) * sc;
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            v[j] = ex2_approx(fmaf(v[j], sc, -mx));   // exactly 0 for keys of other samples / padding keys
            sum += v[j];
        }
        if (DROP) {
            // element ((b*H + h)*S + t)*S + m of the attention-probability dropout stream; m - lo may be out of range for
            // keys whose probability is zero anyway
            const uint64_t didx = (((uint64_t)(samp0 + s_idx) * H + h) * S + t_tok) * (uint64_t)S + (uint64_t)(long long)(c0 - lo);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= dropout_scale(dseed, p.drop.stream, didx + j, p.drop.thresh, p.drop.inv_keep);
        }
        sum = row_sum(sum, red[1], tid);
        if (L.cq == 0) rsum[h][r] = rv ? 1.0f / sum : 0.f;
        if (h > 0) mbar_wait(&bar[BP0 + h - 1], 0);   // P of the previous head has been consumed
        st_row32(smem + OFF_XN, L, v, nullptr);
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if ((L.warp == 0 && elect_one())) {
            tc_fence_after_sync();
            const int hoff = h * dh;
            for (int j = 0; j < Kp16; ++j) {   // O[:, hoff:hoff+dh] = P V_h : A = P (k = key), B = rows hoff.. of V^T
                const uint64_t da = smem_desc_k_sw128(sbase + OFF_XN + (j >> 2) * LTILE) + 2 * (j & 3);
                const uint64_t db = smem_desc_k_sw128(sbase + OFF_VT + (j >> 2) * LTILE + (hoff >> 3) * 1024) + 2 * (j & 3);
                mma_bf16_ss(tmem + ACC3 + hoff, da, db, idPV, j > 0);
            }
            mma_commit(&bar[BP0 + h]);
            if (h + 2 < H) issue_S(h + 2);   // every thread has read S_h: its accumulator is free
        }
        __syncwarp();
    }
    // ---- attention output: rows scaled by 1 / softmax sum -> bf16 A operand of the out-projection -----------------------
    mbar_wait(&bar[BP0 + H - 1], 0);
    tc_fence_after_sync();
    {
        float v[32];
        ld_acc32(tmem, L, ACC3, v);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const float inv = rsum[(c0 + 16 * g) / dh][r];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[16 * g + j] *= inv;
        }
        st_row32(smem + OFF_XN, L, v, (rv && p.attn_save) ? p.attn_save + goff / 8 : nullptr);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        mbar_wait(&bar[BW0 + 3], 0);
        mma_k_tiles(tmem + ACC0, sbase + OFF_XN, LTILE, sbase + OFF_W + 2 * LTILE, LTILE, id128, 2, false);
        mma_commit(&bar[BOUT]);
    }
    __syncwarp();
    // ---- out-projection epilogue: x1 = x + Drop(acc + b) ; LN2 -----------------------------------------------------------
    {
        float v[32], b[32];
        ldg32(p.out_b + c0, b);
        mbar_wait(&bar[BOUT], 0);
        tc_fence_after_sync();
        if (FFN && !FFN_FIRST && (L.warp == 0 && elect_one())) load_w(5);   // slot B drained by the out-projection
        __syncwarp();
        ld_acc32(tmem, L, ACC0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float t = v[j] + b[j];
            if (DROP) t *= dropout_scale(dseed, p.drop.stream + 1, (uint64_t)goff + j, p.drop.thresh, p.drop.inv_keep);
            xr[j] = rv ? xr[j] + t : 0.f;
        }
        float* x1_dst = (FFN && !FFN_FIRST) ? p.x1_save : p.y;   // attention block alone or last: its result is the output
        if (rv && x1_dst) {
            float4* g = reinterpret_cast<float4*>(x1_dst + goff);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = make_float4(xr[4 * j], xr[4 * j + 1], xr[4 * j + 2], xr[4 * j + 3]);
        }
    }
    }   // SA
    } else {
    if constexpr (FFN) {
    {
        float v[32], b[32];
        float mean, rstd;
        ldg32(p.n2_w + c0, b);
        row_stats(xr, red[0], red[1], tid, mean, rstd);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = (xr[j] - mean) * rstd * b[j];
        ldg32(p.n2_b + c0, b);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = rv ? v[j] + b[j] : 0.f;
        st_row32(smem + OFF_XN, L, v, (rv && p.xn2_save) ? p.xn2_save + goff / 8 : nullptr);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        mbar_wait(&bar[BW0 + 4], 0);
        mma_k_tiles(tmem + ACC1, sbase + OFF_XN, LTILE, sbase + OFF_W, LTILE, id128, 2, false);   // FC1
        mma_commit(&bar[BF1]);
    }
    __syncwarp();
    {
        float v[32], b[32];
        ldg32(p.l1_b + c0, b);
        mbar_wait(&bar[BF1], 0);
        tc_fence_after_sync();
        if (FFN_FIRST && (L.warp == 0 && elect_one())) load_w(0);   // slot A drained by FC1: the attention block's Wq
        __syncwarp();
        ld_acc32(tmem, L, ACC1, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float t = gelu_fast(v[j] + b[j]);
            if (DROP) t *= dropout_scale(dseed, p.ffn_stream, (uint64_t)goff + j, p.drop.thresh, p.drop.inv_keep);
            v[j] = rv ? t : 0.f;
        }
        st_row32(smem + OFF_QS, L, v, (rv && p.hact_save) ? p.hact_save + goff / 8 : nullptr);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    if ((L.warp == 0 && elect_one())) {
        tc_fence_after_sync();
        mbar_wait(&bar[BW0 + 5], 0);
        mma_k_tiles(tmem + ACC0, sbase + OFF_QS, LTILE, sbase + OFF_W + 2 * LTILE, LTILE, id128, 2, false);   // FC2
        mma_commit(&bar[BF2]);
    }
    __syncwarp();
    {
        float v[32], b[32];
        ldg32(p.l2_b + c0, b);
        mbar_wait(&bar[BF2], 0);
        tc_fence_after_sync();
        if (FFN_FIRST && (L.warp == 0 && elect_one())) load_w(1);   // slot B drained by FC2: the attention block's Wk
        __syncwarp();
        ld_acc32(tmem, L, ACC0, v);
        if (FFN_FIRST) {   // the attention block follows: the residual stream stays in registers
#pragma unroll
            for (int j = 0; j < 32; ++j) xr[j] = rv ? xr[j] + (v[j] + b[j]) : 0.f;
        } else if (rv) {
            float4* g = reinterpret_cast<float4*>(p.y + goff);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                float t = v[j] + b[j];
                if (DROP) t *= dropout_scale(dseed, p.ffn_stream + 1, (uint64_t)goff + j, p.drop.thresh, p.drop.inv_keep);
                v[j] = xr[j] + t;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
    }
    }   // FFN
    }
    }   // phases
    tc_fence_before_sync();
    __syncthreads();
    if (L.warp == 0) tmem_dealloc(tmem, 512);
}

__global__ void pack_bf16_kernel(sd_pack_args a) {
    // one segment per blockIdx.y: fp32 [rows][K] -> bf16 rows [dst_row0, dst_row0 + rows) of the packed matrix
    const int seg = blockIdx.y;
    const float* src = a.src[seg];
    const long long n = (long long)a.rows[seg] * a.K;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.dst) + (long long)a.dst_row0[seg] * a.K;
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n; i += (long long)gridDim.x * blockDim.x * 4) {
        if (i + 3 < n && ((((uintptr_t)src) & 15) == 0)) {
            const float4 t = *reinterpret_cast<const float4*>(src + i);
            __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&lo);
            u.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(dst + i) = u;
        } else {
            for (long long j = i; j < min(n, i + 4); ++j) dst[j] = __float2bfloat16_rn(src[j]);
        }
    }
}

}  // namespace

extern "C" int sd_pack_weights_bf16(const sd_pack_args* a, int nseg, void* stream) {
    if (!a || nseg <= 0 || nseg > SD_PACK_MAX_SEGMENTS || !a->dst || a->K <= 0 || a->K % 4 != 0) return SD_ERR_BAD_ARG;
    int max_rows = 0;
    for (int i = 0; i < nseg; ++i) {
        if (!a->src[i] || a->rows[i] <= 0) return SD_ERR_BAD_ARG;
        max_rows = max(max_rows, a->rows[i]);
    }
    const int blocks = max(1, min(64, ceil_div((long long)max_rows * a->K, 256 * 4)));
    pack_bf16_kernel<<<dim3(blocks, nseg), 256, 0, (cudaStream_t)stream>>>(*a);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_enc_layer_supported(int d, int ff, int S, int H) {
    if (d != 128 || ff != 128 || S < 8 || S > 128 || H < 1 || H > 8 || d % H != 0) return 0;   // S >= 8: <= 16 samples per tile
    const int dh = d / H;
    if (dh != 16 && dh != 32 && dh != 64) return 0;
    return tensor_map_encoder() != nullptr ? 1 : 0;
}

namespace {
template <bool DROP, bool SA, bool FFN, bool FFN_FIRST = false>
int launch_fwd(const CUtensorMap& tmW, const EncFwdParams& p, int tiles, cudaStream_t st) {
    auto kernel = enc_layer_fwd_kernel<DROP, SA, FFN, FFN_FIRST>;
    static bool configured = false;   // one flag per instantiation
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DYN));
        configured = true;
    }
    SD_CUDA(launch_chain(kernel, dim3(tiles), dim3(LNT), SMEM_DYN, st, tmW, p));
    return SD_OK;
}
}  // namespace

extern "C" int sd_enc_layer_fwd(const sd_enc_layer_desc* d, void* stream) {
    if (!d || !d->x || !d->y || !d->w_packed) return SD_ERR_BAD_ARG;
    const int blocks = d->blocks == 0 ? (SD_LAYER_SA | SD_LAYER_FFN) : d->blocks;
    const bool sa = (blocks & SD_LAYER_SA) != 0, ffn = (blocks & SD_LAYER_FFN) != 0, ffn_first = (blocks & SD_LAYER_FFN_FIRST) != 0;
    if ((blocks & ~(SD_LAYER_SA | SD_LAYER_FFN | SD_LAYER_FFN_FIRST)) != 0) return SD_ERR_BAD_ARG;
    // feed-forward first: inference only (no dropout, no saves), both blocks present
    if (ffn_first && (!sa || !ffn || d->dropout_p > 0.f || d->x1_save || d->xn1_save || d->attn_save || d->xn2_save || d->hact_save))
        return SD_ERR_BAD_ARG;
    if (sa && (!d->in_b || !d->out_b || !d->n1_w || !d->n1_b)) return SD_ERR_BAD_ARG;
    if (ffn && (!d->l1_b || !d->l2_b || !d->n2_w || !d->n2_b)) return SD_ERR_BAD_ARG;
    if (d->B <= 0) return SD_OK;
    if (!sd_enc_layer_supported(128, 128, d->S, d->H)) return SD_ERR_UNSUPPORTED;
    const int w_row_ffn = d->w_row_ffn > 0 ? d->w_row_ffn : d->w_row0 + 512;
    if (sa && (d->w_row0 < 0 || d->w_row0 + 512 > d->w_rows_total)) return SD_ERR_BAD_ARG;
    if (ffn && (w_row_ffn < 0 || w_row_ffn + 256 > d->w_rows_total)) return SD_ERR_BAD_ARG;
    if ((((uintptr_t)d->x) | ((uintptr_t)d->y) | ((uintptr_t)d->x1_save) | ((uintptr_t)d->xn1_save) | ((uintptr_t)d->attn_save) |
         ((uintptr_t)d->xn2_save) | ((uintptr_t)d->hact_save)) & 15)
        return SD_ERR_BAD_ARG;
    CUtensorMap tmW;
    if (!encode_bf16_2d(&tmW, d->w_packed, d->w_rows_total, 128, 128, 128)) return SD_ERR_UNSUPPORTED;
    EncFwdParams p;
    p.x = d->x; p.y = d->y; p.B = d->B; p.S = d->S; p.spt = 128 / d->S; p.H = d->H; p.dh = 128 / d->H;
    p.w_row0 = d->w_row0; p.w_row_ffn = w_row_ffn;
    p.in_b = d->in_b; p.out_b = d->out_b; p.l1_b = d->l1_b; p.l2_b = d->l2_b;
    p.n1_w = d->n1_w; p.n1_b = d->n1_b; p.n2_w = d->n2_w; p.n2_b = d->n2_b;
    p.x1_save = d->x1_save;
    p.xn1_save = (uint4*)d->xn1_save; p.attn_save = (uint4*)d->attn_save; p.xn2_save = (uint4*)d->xn2_save;
    p.hact_save = (uint4*)d->hact_save;
    p.drop = make_dropout(d->dropout_p, d->dropout_seed, d->dropout_stream);
    p.ffn_stream = d->dropout_stream_ffn != 0 ? d->dropout_stream_ffn : d->dropout_stream + 2;
    const int tiles = ceil_div(d->B, p.spt);
    cudaStream_t st = (cudaStream_t)stream;
    const bool drop = p.drop.thresh != 0;
    if (ffn_first) return launch_fwd<false, true, true, true>(tmW, p, tiles, st);
    if (sa && ffn) return drop ? launch_fwd<true, true, true>(tmW, p, tiles, st) : launch_fwd<false, true, true>(tmW, p, tiles, st);
    if (sa) return drop ? launch_fwd<true, true, false>(tmW, p, tiles, st) : launch_fwd<false, true, false>(tmW, p, tiles, st);
    return drop ? launch_fwd<true, false, true>(tmW, p, tiles, st) : launch_fwd<false, false, true>(tmW, p, tiles, st);
}
