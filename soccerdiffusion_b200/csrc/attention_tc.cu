// Multi-head attention core on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM, fp32 softmax), forward and
// backward, for sequences that fit one UMMA tile: T <= 128 queries, M <= 128 keys, head size 16/32/64 — the encoders'
// self-attention (S = 100 tokens, 10 frames) and the denoiser's self-attention (T = 10) of the reference
// (nn.MultiheadAttention / F.scaled_dot_product_attention, torch/nn/functional.py:6623-6690; call sites
// ml/model/encoder/base.py:30-40, ml/model/decoder.py:24-35).  Longer memories (cross-attention over 312 context
// tokens) stay on the fp32 kernel of attention.cu.
//
// One CTA (128 threads = the four TMEM lane quarters) per (sample, head):
//   forward   S = Q K^T  ->  row softmax (thread = query row, scores read from TMEM), dropout  ->  P (bf16, smem)
//             O = P V    ->  rows scaled by 1/sum, log-sum-exp saved
//   backward  S = Q K^T, dP = dO V^T  ->  P = exp(S - lse), dS = P (dP*mask - delta) * scale  (thread = query row)
//             dV = P'^T dO,  dK = dS^T Q,  dQ = dS K
// A [row][64-element] bf16 tile with the 128-byte swizzle is simultaneously a K-major operand (row index = M/N, the
// 64 elements = K) and an MN-major operand (row index = K, the 64 elements = M/N): Q, K, V, dO, P and dS are each
// staged ONCE and used in both roles by switching the descriptor (tc_common.cuh) — no transposed copies anywhere.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/sd_b200.h"

using namespace sd;
using namespace sdtc;

namespace {

constexpr int NT = 128;
constexpr int TILE = 128 * 128;   // bytes: 128 rows x 128 B

struct AttnTcParams {
    const float* Q; long long ldq;
    const float* K; long long ldk;
    const float* V; long long ldv;
    float* O; long long ldo;
    float* lse;
    // backward only
    const float* dO; long long lddo;
    float* dQ; long long lddq;
    float* dK; long long lddk;
    float* dV; long long lddv;
    int B, H, T, M, dh;
    float scale;
    Dropout drop;
};

// Operand staging: rows [0,R) of each source (row stride ld floats, `dh` floats per row) -> bf16 [128][64] swizzled
// tile, rows >= R zero.  The global loads of NOPS operands x IB passes are all issued before the first conversion /
// shared store, so a CTA pays one DRAM round trip per batch instead of one per pass.
struct StageSrc {
    const float* src; long long ld; int R;
};
template <int NOPS, int IB>
__device__ __forceinline__ void stage_multi(uint8_t* tile0, const StageSrc* ops, int dh) {
    const int cpr = dh >> 3;   // 16-byte chunks per row
    const int iters = (128 * cpr) / NT;
    for (int it0 = 0; it0 < iters; it0 += IB) {
        float4 lo[NOPS][IB], hi[NOPS][IB];
#pragma unroll
        for (int o = 0; o < NOPS; ++o)
#pragma unroll
            for (int i = 0; i < IB; ++i) {
                const int item = threadIdx.x + NT * (it0 + i);
                const int r = item / cpr, c = item % cpr;
                lo[o][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                hi[o][i] = lo[o][i];
                if (it0 + i < iters && r < ops[o].R) {
                    const float* g = ops[o].src + (long long)r * ops[o].ld + 8 * c;
                    lo[o][i] = *reinterpret_cast<const float4*>(g);
                    hi[o][i] = *reinterpret_cast<const float4*>(g + 4);
                }
            }
#pragma unroll
        for (int o = 0; o < NOPS; ++o)
#pragma unroll
            for (int i = 0; i < IB; ++i) {
                const int item = threadIdx.x + NT * (it0 + i);
                const int r = item / cpr, c = item % cpr;
                if (it0 + i < iters) {
                    const float f[8] = {lo[o][i].x, lo[o][i].y, lo[o][i].z, lo[o][i].w,
                                        hi[o][i].x, hi[o][i].y, hi[o][i].z, hi[o][i].w};
                    *reinterpret_cast<uint4*>(tile0 + o * TILE + sw128_chunk_off(r, c)) = pack8_bf16(f);
                }
            }
    }
}

// this thread's row of an fp32 [128][ncols] TMEM accumulator -> registers, 32 columns at a time
__device__ __forceinline__ void ld_row32(uint32_t tmem, int warp, int col0, float* v) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col0, v);
}

// write 32 consecutive values (columns col0..col0+31 of row r) as bf16 into the [128][64]-tiled, swizzled P/dS buffer
__device__ __forceinline__ void st_row32_bf16(uint8_t* buf, int r, int col0, const float* v) {
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
        const int col = col0 + 8 * c8;
        *reinterpret_cast<uint4*>(buf + (col >> 6) * TILE + sw128_chunk_off(r, (col & 63) >> 3)) = pack8_bf16(v + 8 * c8);
    }
}

__global__ void __launch_bounds__(NT) attn_tc_fwd_kernel(const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* Qs = smem;               // [t][dh]
    uint8_t* Ks = smem + TILE;        // [m][dh]
    uint8_t* Vs = smem + 2 * TILE;    // [m][dh]
    uint8_t* Ps = smem + 3 * TILE;    // [t][m] : two 64-column tiles
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int T = p.T, M = p.M, dh = p.dh;
    const int Mp = (M + 15) & ~15;
    if (warp == 0 && elect_one()) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    {   // Qs, Ks, Vs are consecutive tiles
        const StageSrc ops[3] = {{p.Q + (long long)b * T * p.ldq + h * dh, p.ldq, T},
                                 {p.K + (long long)b * M * p.ldk + h * dh, p.ldk, M},
                                 {p.V + (long long)b * M * p.ldv + h * dh, p.ldv, M}};
        stage_multi<3, 4>(Qs, ops, dh);
    }
    const uint64_t dseed = p.drop.resolve();
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 0 && elect_one()) {   // S[t][m] = sum_c Q[t][c] K[m][c]   (both K-major, k = dh)
        const uint32_t idesc = instr_desc_bf16(128, Mp, 0, 0);
        const uint64_t da = smem_desc_k_sw128(smem_u32(Qs)), db = smem_desc_k_sw128(smem_u32(Ks));
        for (int j = 0; j < dh / 16; ++j) mma_bf16_ss(tmem, da + 2 * j, db + 2 * j, idesc, j > 0);
        mma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], 0);
    tc_fence_after_sync();
    // ---- softmax of this thread's query row -----------------------------------------------------------------
    const int t = tid;
    float mx = -INFINITY;
    for (int c0 = 0; c0 < Mp; c0 += 32) {
        float v[32];
        ld_row32(tmem, warp, c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (c0 + j < M) mx = fmaxf(mx, v[j] * p.scale);
    }
    float sum = 0.f;
    const uint64_t row_idx = ((uint64_t)(b * p.H + h) * T + t) * (uint64_t)M;
    for (int c0 = 0; c0 < Mp; c0 += 32) {
        float v[32];
        ld_row32(tmem, warp, c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int m = c0 + j;
            float e = 0.f;
            if (m < M) {
                e = expf(v[j] * p.scale - mx);
                sum += e;
                if (t < T) e *= p.drop.at(dseed, row_idx + m);
            }
            v[j] = e;   // columns >= M are exactly zero: V rows >= M never contribute
        }
        st_row32_bf16(Ps, t, c0, v);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (warp == 0 && elect_one()) {   // O[t][c] = sum_m P[t][m] V[m][c] : A = P K-major (k = m), B = V MN-major (k rows = m)
        const uint32_t idesc = instr_desc_bf16(128, dh, 0, 1);
        for (int j = 0; j < Mp / 16; ++j) {
            const uint64_t da = smem_desc_k_sw128(smem_u32(Ps + (j >> 2) * TILE)) + 2 * (j & 3);
            const uint64_t db = smem_desc_mn_sw128(smem_u32(Vs), TILE, 1024) + 128 * j;
            mma_bf16_ss(tmem + 128, da, db, idesc, j > 0);
        }
        mma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tc_fence_after_sync();
    {
        const float inv = 1.0f / sum;
        for (int c0 = 0; c0 < dh; c0 += 32) {
            float v[32];
            ld_row32(tmem, warp, 128 + c0, v);
            if (t < T) {
                float* o = p.O + (long long)(b * T + t) * p.ldo + h * dh + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (c0 + j < dh) *reinterpret_cast<float4*>(o + j) = make_float4(v[j] * inv, v[j + 1] * inv, v[j + 2] * inv, v[j + 3] * inv);
            }
        }
        if (t < T && p.lse) p.lse[((long long)b * p.H + h) * T + t] = mx + logf(sum);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

__global__ void __launch_bounds__(NT) attn_tc_bwd_kernel(const AttnTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* Qs = smem;                 // [t][dh]
    uint8_t* Ks = smem + TILE;          // [m][dh]
    uint8_t* Vs = smem + 2 * TILE;      // [m][dh]
    uint8_t* dOs = smem + 3 * TILE;     // [t][dh]
    uint8_t* Ps = smem + 4 * TILE;      // P * dropout   [t][m], two tiles
    uint8_t* dSs = smem + 6 * TILE;     // dS * scale    [t][m], two tiles
    const int tid = threadIdx.x, warp = tid >> 5;
    const int b = blockIdx.x / p.H, h = blockIdx.x % p.H;
    const int T = p.T, M = p.M, dh = p.dh;
    const int Mp = (M + 15) & ~15, Tp = (T + 15) & ~15;
    if (warp == 0 && elect_one()) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    const float* Og = p.O + (long long)b * T * p.ldo + h * dh;
    const float* dOg = p.dO + (long long)b * T * p.lddo + h * dh;
    // delta[t] = dO[t] . O[t] and the saved log-sum-exp of this thread's row (loads issued ahead of the staging)
    const int t = tid;
    float delta = 0.f, lse = 0.f;
    if (t < T) {
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int c = 0; c < dh; c += 4) {
            const float4 a = *reinterpret_cast<const float4*>(dOg + (long long)t * p.lddo + c);
            const float4 o = *reinterpret_cast<const float4*>(Og + (long long)t * p.ldo + c);
            d4[0] = fmaf(a.x, o.x, d4[0]); d4[1] = fmaf(a.y, o.y, d4[1]);
            d4[2] = fmaf(a.z, o.z, d4[2]); d4[3] = fmaf(a.w, o.w, d4[3]);
        }
        delta = (d4[0] + d4[1]) + (d4[2] + d4[3]);
        lse = p.lse[((long long)b * p.H + h) * T + t];
    }
    {   // Qs, Ks, Vs, dOs are consecutive tiles
        const StageSrc ops[4] = {{p.Q + (long long)b * T * p.ldq + h * dh, p.ldq, T},
                                 {p.K + (long long)b * M * p.ldk + h * dh, p.ldk, M},
                                 {p.V + (long long)b * M * p.ldv + h * dh, p.ldv, M},
                                 {dOg, p.lddo, T}};
        stage_multi<4, 2>(Qs, ops, dh);
    }
    const uint64_t dseed = p.drop.resolve();
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (warp == 0 && elect_one()) {
        const uint32_t idesc = instr_desc_bf16(128, Mp, 0, 0);
        const uint64_t dq = smem_desc_k_sw128(smem_u32(Qs)), dk = smem_desc_k_sw128(smem_u32(Ks));
        const uint64_t dg = smem_desc_k_sw128(smem_u32(dOs)), dv = smem_desc_k_sw128(smem_u32(Vs));
        for (int j = 0; j < dh / 16; ++j) mma_bf16_ss(tmem, dq + 2 * j, dk + 2 * j, idesc, j > 0);         // S
        for (int j = 0; j < dh / 16; ++j) mma_bf16_ss(tmem + 128, dg + 2 * j, dv + 2 * j, idesc, j > 0);   // dP = dO V^T
        mma_commit(&bar[0]);
    }
    mbar_wait(&bar[0], 0);
    tc_fence_after_sync();
    const uint64_t row_idx = ((uint64_t)(b * p.H + h) * T + t) * (uint64_t)M;
    for (int c0 = 0; c0 < Mp; c0 += 32) {
        float s[32], dp[32];
        ld_row32(tmem, warp, c0, s);
        ld_row32(tmem, warp, 128 + c0, dp);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int m = c0 + j;
            float pd = 0.f, ds = 0.f;
            if (m < M && t < T) {
                const float pr = expf(s[j] * p.scale - lse);
                const float dm = p.drop.at(dseed, row_idx + m);
                pd = pr * dm;
                ds = pr * (dp[j] * dm - delta) * p.scale;
            }
            s[j] = pd;    // rows >= T and columns >= M are exactly zero
            dp[j] = ds;
        }
        st_row32_bf16(Ps, t, c0, s);
        st_row32_bf16(dSs, t, c0, dp);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();   // every thread has consumed S / dP: their TMEM columns are reused below
    tc_fence_after_sync();
    if (warp == 0 && elect_one()) {
        const uint32_t id_mn = instr_desc_bf16(128, dh, 1, 1);   // A MN-major (k rows = t), B MN-major
        const uint32_t id_kn = instr_desc_bf16(128, dh, 0, 1);   // A K-major (k = m),     B MN-major
        // dV[m][c] = sum_t P'[t][m] dO[t][c]   -> columns [0, dh)
        for (int j = 0; j < Tp / 16; ++j)
            mma_bf16_ss(tmem, smem_desc_mn_sw128(smem_u32(Ps), TILE, 1024) + 128 * j,
                        smem_desc_mn_sw128(smem_u32(dOs), TILE, 1024) + 128 * j, id_mn, j > 0);
        // dK[m][c] = sum_t dS[t][m] Q[t][c]    -> columns [64, 64+dh)
        for (int j = 0; j < Tp / 16; ++j)
            mma_bf16_ss(tmem + 64, smem_desc_mn_sw128(smem_u32(dSs), TILE, 1024) + 128 * j,
                        smem_desc_mn_sw128(smem_u32(Qs), TILE, 1024) + 128 * j, id_mn, j > 0);
        // dQ[t][c] = sum_m dS[t][m] K[m][c]    -> columns [128, 128+dh)
        for (int j = 0; j < Mp / 16; ++j)
            mma_bf16_ss(tmem + 128, smem_desc_k_sw128(smem_u32(dSs + (j >> 2) * TILE)) + 2 * (j & 3),
                        smem_desc_mn_sw128(smem_u32(Ks), TILE, 1024) + 128 * j, id_kn, j > 0);
        mma_commit(&bar[1]);
    }
    mbar_wait(&bar[1], 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < dh; c0 += 32) {
        float v[32];
        ld_row32(tmem, warp, c0, v);          // dV row m = tid
        if (tid < M) {
            float* o = p.dV + (long long)(b * M + tid) * p.lddv + h * dh + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (c0 + j < dh) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        ld_row32(tmem, warp, 64 + c0, v);     // dK row m = tid
        if (tid < M) {
            float* o = p.dK + (long long)(b * M + tid) * p.lddk + h * dh + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (c0 + j < dh) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        ld_row32(tmem, warp, 128 + c0, v);    // dQ row t = tid
        if (tid < T) {
            float* o = p.dQ + (long long)(b * T + tid) * p.lddq + h * dh + c0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (c0 + j < dh) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace

// supported: T, M in [1,128], dh in {16,32,64}, all pointers 16-byte aligned and row strides multiples of 4 floats
extern "C" int sd_attention_tc_supported(int T, int M, int dh) {
    return (T >= 1 && T <= 128 && M >= 1 && M <= 128 && (dh == 16 || dh == 32 || dh == 64)) ? 1 : 0;
}

extern "C" int sd_attention_tc_fwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V,
                                   long long ldv, float* O, long long ldo, float* lse, int B, int H, int T, int M, int dh,
                                   float dropout_p, unsigned long long seed, unsigned int stream_id, void* stream) {
    if (B <= 0 || T <= 0) return SD_OK;
    if (!Q || !K || !V || !O || H <= 0 || M <= 0) return SD_ERR_BAD_ARG;
    if (!sd_attention_tc_supported(T, M, dh)) return SD_ERR_UNSUPPORTED;
    if (!al16(Q) || !al16(K) || !al16(V) || !al16(O) || (ldq | ldk | ldv | ldo) % 4 != 0) return SD_ERR_UNSUPPORTED;
    AttnTcParams p{};
    p.Q = Q; p.ldq = ldq; p.K = K; p.ldk = ldk; p.V = V; p.ldv = ldv; p.O = O; p.ldo = ldo; p.lse = lse;
    p.B = B; p.H = H; p.T = T; p.M = M; p.dh = dh; p.scale = 1.0f / sqrtf((float)dh);
    p.drop = make_dropout(dropout_p, seed, stream_id);
    const int smem = 5 * TILE + 1024;
    static bool configured = false;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    attn_tc_fwd_kernel<<<B * H, NT, smem, (cudaStream_t)stream>>>(p);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_attention_tc_bwd(const float* Q, long long ldq, const float* K, long long ldk, const float* V,
                                   long long ldv, const float* O, long long ldo, const float* dO, long long lddo,
                                   const float* lse, float* dQ, long long lddq, float* dK, long long lddk, float* dV,
                                   long long lddv, int B, int H, int T, int M, int dh, float dropout_p,
                                   unsigned long long seed, unsigned int stream_id, void* stream) {
    if (B <= 0 || T <= 0) return SD_OK;
    if (!Q || !K || !V || !O || !dO || !lse || !dQ || !dK || !dV || H <= 0 || M <= 0) return SD_ERR_BAD_ARG;
    if (!sd_attention_tc_supported(T, M, dh)) return SD_ERR_UNSUPPORTED;
    if (!al16(Q) || !al16(K) || !al16(V) || !al16(O) || !al16(dO) || !al16(dQ) || !al16(dK) || !al16(dV) ||
        (ldq | ldk | ldv | ldo | lddo | lddq | lddk | lddv) % 4 != 0)
        return SD_ERR_UNSUPPORTED;
    AttnTcParams p{};
    p.Q = Q; p.ldq = ldq; p.K = K; p.ldk = ldk; p.V = V; p.ldv = ldv; p.O = const_cast<float*>(O); p.ldo = ldo;
    p.lse = const_cast<float*>(lse); p.dO = dO; p.lddo = lddo; p.dQ = dQ; p.lddq = lddq; p.dK = dK; p.lddk = lddk;
    p.dV = dV; p.lddv = lddv;
    p.B = B; p.H = H; p.T = T; p.M = M; p.dh = dh; p.scale = 1.0f / sqrtf((float)dh);
    p.drop = make_dropout(dropout_p, seed, stream_id);
    const int smem = 8 * TILE + 1024;
    static bool configured = false;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    attn_tc_bwd_kernel<<<B * H, NT, smem, (cudaStream_t)stream>>>(p);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
