// bf16 tensor-core GEMM family on tcgen05 (sm_100a): same contract and fused prologues/epilogues as the
// true-fp32 kernel in gemm_f32.cu (sd_gemm with precision = SD_PREC_BF16), i.e. the 2e-2 mode.
//
//   C[m][n] = epilogue( sum_k opA(A)[m][k] * opB(B)[k][n] )      A, B, C fp32 in HBM
//
// Operands are converted fp32 -> bf16 on the way into shared memory (LayerNorm-on-load applied in fp32 first),
// stored K-major with the 128-byte swizzle tcgen05 expects; tcgen05.mma (M=128, N<=256, K=16 per instruction,
// issued by one thread) accumulates in fp32 in TMEM; tcgen05.commit -> mbarrier tracks completion; the epilogue
// reads TMEM with tcgen05.ld, transposes through shared memory so that every global access (bias, saved
// pre-activation, GELU-derivative source, residual, positional encoding, output) is coalesced along n.
//
// With K = d_model = 128 these GEMMs are HBM-bound (64 flop/B): the tensor pipe makes the
// arithmetic free so that the kernel runs at the speed of streaming the activations once.
//
// Reference ops replaced: nn.Linear / Conv1d(k=stride=patch) / packed in_proj of nn.MultiheadAttention and their
// autograd (encoder/base.py:28,49; decoder.py:23,36,48,54; torch/nn/functional.py:5849-5855).
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/sd_b200.h"

#include <cstdlib>

using namespace sd;
using namespace sdtc;

namespace {

constexpr int TM = 128;     // UMMA M: output rows per CTA
constexpr int TK = 64;      // k elements per pipeline stage = one 128-byte swizzle row of bf16
constexpr int NT = 256;
constexpr int STAGES = 2;
constexpr int A_STAGE_BYTES = TM * TK * 2;

struct TcParams {
    const float* A; long long lda;
    const float* B; long long ldb;
    float* C; long long ldc;
    int M, N, K;
    int vecA, vecB;
    int vecC;        // every epilogue pointer 16-byte aligned, every row pitch and N a multiple of 4
    const float* ln_mean; const float* ln_rstd; const float* ln_gamma; const float* ln_beta;
    int ln_on_a, ln_on_b;
    const float* bias;
    const float* residual; long long ldr;
    const float* pe; int pe_period;
    float* pre_out; long long ldp;
    const float* gelu_grad_src; long long ldg;
    int act;
    int accumulate;
    int k_per_slice;
    float alpha;
    Dropout drop;
    int BN;          // tile columns (multiple of 16, <= 256)
    int tmem_cols;   // power of two >= max(32, BN)
};

// Operand staging is split into a LOAD phase (global -> registers: every load of a batch of 4 row passes is in
// flight before the first use, and the A and B batches are issued back to back) and a STORE phase (LayerNorm-on-load,
// fp32 -> bf16, swizzled 16-byte shared stores).  One DRAM round trip per k chunk instead of one per row pass.
struct Frag {
    float4 lo[4], hi[4];   // 8 consecutive source elements per pass
    float rs[4], nm[4];    // LayerNorm rstd and -mean*rstd of the pass's row (K-contiguous sources only)
};

__device__ __forceinline__ void load8(float4& lo, float4& hi, const float* __restrict__ g, bool full, int nvalid) {
    if (full) {
        lo = *reinterpret_cast<const float4*>(g);
        hi = *reinterpret_cast<const float4*>(g + 4);
    } else {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = (j < nvalid) ? g[j] : 0.f;
        lo = make_float4(t[0], t[1], t[2], t[3]);
        hi = make_float4(t[4], t[5], t[6], t[7]);
    }
}

// K-contiguous source (src[row][k]): thread -> 16-byte chunk c = tid & 7 of rows (tid >> 3) + 32 * (it0 + i).
template <bool LN>
__device__ __forceinline__ void load_k_contig(Frag& f, const float* __restrict__ src, long long ld, int row0,
                                              int rows_total, int R, int it0, int k0, int ke, bool vec,
                                              const TcParams& p) {
    const int c = threadIdx.x & 7;
    const int k = k0 + 8 * c;
    const bool full = vec && (k + 7 < ke);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (threadIdx.x >> 3) + 32 * (it0 + i);
        const int row = row0 + r;
        f.lo[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        f.hi[i] = f.lo[i];
        if (r < R && row < rows_total && k < ke) {
            load8(f.lo[i], f.hi[i], src + (long long)row * ld + k, full, ke - k);
            if (LN) {
                f.rs[i] = p.ln_rstd[row];
                f.nm[i] = -p.ln_mean[row] * f.rs[i];
            }
        }
    }
}
template <bool LN>
__device__ __forceinline__ void store_k_contig(uint8_t* tile, const Frag& f, int row0, int rows_total, int R, int it0,
                                               int k0, int ke, const TcParams& p) {
    const int c = threadIdx.x & 7;
    const int k = k0 + 8 * c;
    float ga[8], be[8];
    if (LN) {   // the thread's 8 k positions are the same for every row it stages
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ga[j] = (k + j < ke) ? __ldg(p.ln_gamma + k + j) : 0.f;
            be[j] = (k + j < ke) ? __ldg(p.ln_beta + k + j) : 0.f;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = (threadIdx.x >> 3) + 32 * (it0 + i);
        if (r >= R) break;
        float v[8] = {f.lo[i].x, f.lo[i].y, f.lo[i].z, f.lo[i].w, f.hi[i].x, f.hi[i].y, f.hi[i].z, f.hi[i].w};
        if (LN && row0 + r < rows_total && k < ke) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(fmaf(v[j], f.rs[i], f.nm[i]), ga[j], be[j]);   // 0 beyond ke: ga = be = 0
        }
        *reinterpret_cast<uint4*>(tile + sw128_chunk_off(r, c)) = pack8_bf16(v);
    }
}

// MN-contiguous source (src[k][mn]) staged WITHOUT transposing: MN-major operand layout (64 MN elements = one
// 128-byte row per k, 128-byte swizzle).  R = MN extent of the tile (multiple of 8), TK k rows; item = (k row, chunk).
__device__ __forceinline__ void load_mn_major(Frag& f, const float* __restrict__ src, long long ld, int mn0,
                                              int mn_total, int R, int it0, int k0, int ke, bool vec) {
    const int chunks = R >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int item = threadIdx.x + NT * (it0 + i);
        const int c = item % chunks, kr = item / chunks;
        const int mn = mn0 + 8 * c, k = k0 + kr;
        f.lo[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        f.hi[i] = f.lo[i];
        if (kr < TK && k < ke && mn < mn_total)
            load8(f.lo[i], f.hi[i], src + (long long)k * ld + mn, vec && mn + 7 < mn_total, mn_total - mn);
    }
}
template <bool LN>
__device__ __forceinline__ void store_mn_major(uint8_t* tile, const Frag& f, int mn0, int mn_total, int R, int it0, int k0,
                                               int ke, const TcParams& p) {
    const int chunks = R >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int item = threadIdx.x + NT * (it0 + i);
        const int c = item % chunks, kr = item / chunks;
        if (kr >= TK) break;
        const int mn = mn0 + 8 * c, k = k0 + kr;
        float v[8] = {f.lo[i].x, f.lo[i].y, f.lo[i].z, f.lo[i].w, f.hi[i].x, f.hi[i].y, f.hi[i].z, f.hi[i].w};
        if (LN && k < ke && mn < mn_total) {   // the k index is the normalised row, the MN index the feature
            const float rs = p.ln_rstd[k];
            const float nm = -p.ln_mean[k] * rs;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (mn + j < mn_total) v[j] = fmaf(fmaf(v[j], rs, nm), __ldg(p.ln_gamma + mn + j), __ldg(p.ln_beta + mn + j));
        }
        *reinterpret_cast<uint4*>(tile + sw128_mn_chunk_off(8 * c, kr, TK)) = pack8_bf16(v);
    }
}

// Stage a [R rows][64 k] bf16 tile from a matrix whose ROW index is contiguous in memory (src[k][row]):
// the transposition happens here (each thread gathers 8 k values of one row; a warp reads 128 contiguous bytes
// per k and writes conflict-free 16-byte chunks).
template <bool LN>
__device__ __forceinline__ void stage_row_contig(uint8_t* tile, const float* __restrict__ src, long long ld, int row0,
                                                 int rows_total, int R, int k0, int ke, const TcParams& p) {
    for (int item = threadIdx.x; item < R * 8; item += NT) {
        const int r = item % R, c = item / R;
        const int row = row0 + r;
        const int k = k0 + 8 * c;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
        if (row < rows_total) {
            const float* g = src + (long long)k * ld + row;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (k + j < ke) f[j] = g[(long long)j * ld];
            if (LN) {
                const float ga = __ldg(p.ln_gamma + row), be = __ldg(p.ln_beta + row);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (k + j < ke) f[j] = (f[j] - p.ln_mean[k + j]) * p.ln_rstd[k + j] * ga + be;
            }
        }
        *reinterpret_cast<uint4*>(tile + sw128_chunk_off(r, c)) = pack8_bf16(f);
    }
}

// epilogue feature mask (compile-time specialisations of the combinations the layer code uses; EPI_GENERIC keeps
// every feature a run-time flag)
constexpr int EPI_PRE = 1, EPI_GELU = 2, EPI_GG = 4, EPI_DROP = 8, EPI_PE = 16, EPI_RES = 32, EPI_ACC = 64;
constexpr int EPI_GENERIC = -1;

// MNMAJ: operands whose MN index is contiguous in memory (A_KM / B_KN) are staged untransposed and described to
// tcgen05 as MN-major; otherwise they are transposed by the staging threads into K-major tiles.
template <bool A_KM, bool B_KN, int EPI, bool MNMAJ>
__global__ void __launch_bounds__(NT) gemm_tc_kernel(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_stage[STAGES];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;

    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int BN = p.BN;
    const int stage_bytes = A_STAGE_BYTES + ((BN + 63) & ~63) * TK * 2;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0 && elect_one()) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar_stage[s], 1);
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    pdl_trigger();
    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)p.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    pdl_wait();   // every operand may come from the preceding kernel

    const int m0 = blockIdx.y * TM, n0 = blockIdx.x * BN;
    const int kb = blockIdx.z * p.k_per_slice;
    const int ke = min(p.K, kb + p.k_per_slice);
    const int nchunks = (ke - kb + TK - 1) / TK;
    constexpr bool A_MN = A_KM && MNMAJ, B_MN = B_KN && MNMAJ;
    const uint32_t idesc = instr_desc_bf16(TM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    const int BNr = (BN + 63) & ~63;   // MN-major B tiles are laid out in 64-wide blocks

    for (int ci = 0; ci < nchunks; ++ci) {
        const int s = ci % STAGES;
        const int k0 = kb + ci * TK;
        uint8_t* As = smem + s * stage_bytes;
        uint8_t* Bs = As + A_STAGE_BYTES;
        constexpr bool A_ROWC = A_KM && !MNMAJ, B_ROWC = B_KN && !MNMAJ;   // transposing staging (SD_B200_TC_MNMAJOR=0)
        const int itersB = B_MN ? (BNr >> 5) : ((BN + 31) >> 5);         // 32-row (or 256-item) passes of the B tile
        Frag fa, fb;
        // load phase: A and the first B batch in flight together
        if (A_MN) load_mn_major(fa, p.A, p.lda, m0, p.M, TM, 0, k0, ke, p.vecA);
        else if (!A_ROWC) {
            if (p.ln_on_a) load_k_contig<true>(fa, p.A, p.lda, m0, p.M, TM, 0, k0, ke, p.vecA, p);
            else load_k_contig<false>(fa, p.A, p.lda, m0, p.M, TM, 0, k0, ke, p.vecA, p);
        }
        if (B_MN) load_mn_major(fb, p.B, p.ldb, n0, p.N, BNr, 0, k0, ke, p.vecB);
        else if (!B_ROWC) load_k_contig<false>(fb, p.B, p.ldb, n0, p.N, BN, 0, k0, ke, p.vecB, p);
        // store phase (the stage must have been drained by the MMAs that last read it)
        if (ci >= STAGES) mbar_wait(&bar_stage[s], (uint32_t)((ci / STAGES - 1) & 1));
        if (A_MN) store_mn_major<false>(As, fa, m0, p.M, TM, 0, k0, ke, p);
        else if (A_ROWC) stage_row_contig<false>(As, p.A, p.lda, m0, p.M, TM, k0, ke, p);
        else {
            if (p.ln_on_a) store_k_contig<true>(As, fa, m0, p.M, TM, 0, k0, ke, p);
            else store_k_contig<false>(As, fa, m0, p.M, TM, 0, k0, ke, p);
        }
        if (B_ROWC) {
            if (p.ln_on_b) stage_row_contig<true>(Bs, p.B, p.ldb, n0, p.N, BN, k0, ke, p);
            else stage_row_contig<false>(Bs, p.B, p.ldb, n0, p.N, BN, k0, ke, p);
        } else {
            for (int it0 = 0; it0 < itersB; it0 += 4) {
                if (it0 > 0) {
                    if (B_MN) load_mn_major(fb, p.B, p.ldb, n0, p.N, BNr, it0, k0, ke, p.vecB);
                    else load_k_contig<false>(fb, p.B, p.ldb, n0, p.N, BN, it0, k0, ke, p.vecB, p);
                }
                if (B_MN) {
                    if (p.ln_on_b) store_mn_major<true>(Bs, fb, n0, p.N, BNr, it0, k0, ke, p);
                    else store_mn_major<false>(Bs, fb, n0, p.N, BNr, it0, k0, ke, p);
                } else {
                    store_k_contig<false>(Bs, fb, n0, p.N, BN, it0, k0, ke, p);
                }
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (warp == 0 && elect_one()) {
            tc_fence_after_sync();
            // K-major: +32 bytes of K per step inside the swizzle atom; MN-major: +2 atoms of 8 k-rows (2048 bytes)
            const uint64_t da = A_MN ? smem_desc_mn_sw128(smem_u32(As), TK * 128, 1024) : smem_desc_k_sw128(smem_u32(As));
            const uint64_t db = B_MN ? smem_desc_mn_sw128(smem_u32(Bs), TK * 128, 1024) : smem_desc_k_sw128(smem_u32(Bs));
            const uint64_t sa = A_MN ? 128 : 2, sb = B_MN ? 128 : 2;
            const int ksteps = (min(TK, ke - k0) + 15) / 16;
            for (int j = 0; j < ksteps; ++j)
                mma_bf16_ss(tmem, da + sa * (uint64_t)j, db + sb * (uint64_t)j, idesc, (ci > 0 || j > 0) ? 1u : 0u);
            mma_commit(&bar_stage[s]);
            if (ci == nchunks - 1) mma_commit(&bar_done);
        }
    }
    mbar_wait(&bar_done, 0);
    tc_fence_after_sync();

    // ---- epilogue: TMEM -> registers -> (transpose in smem) -> coalesced global ----------------------
    constexpr bool G = EPI == EPI_GENERIC;
    const bool f_pre = G ? p.pre_out != nullptr : (EPI & EPI_PRE) != 0;
    const bool f_gelu = G ? p.act == SD_ACT_GELU : (EPI & EPI_GELU) != 0;
    const bool f_gg = G ? p.gelu_grad_src != nullptr : (EPI & EPI_GG) != 0;
    const bool f_drop = G ? p.drop.thresh != 0 : (EPI & EPI_DROP) != 0;
    const bool f_pe = G ? p.pe != nullptr : (EPI & EPI_PE) != 0;
    const bool f_res = G ? p.residual != nullptr : (EPI & EPI_RES) != 0;
    const bool f_acc = G ? p.accumulate != 0 : (EPI & EPI_ACC) != 0;
    constexpr int BP = 36;                        // bounce row pitch in floats: 16-byte rows, conflict-free both ways
    float* bounce = reinterpret_cast<float*>(smem) + warp * (32 * BP);
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const bool split = gridDim.z > 1;
    const bool first_slice = blockIdx.z == 0;
    const int nblocks = (BN + 31) / 32;
    const int m_first = m0 + q * 32;
    const int rows = min(32, p.M - m_first);
    const uint64_t dseed = f_drop ? p.drop.seed + (p.drop.seed_dev ? __ldg(p.drop.seed_dev) : 0ull) : 0ull;
    for (int cb = warp >> 2; cb < nblocks; cb += 2) {
        if (n0 + cb * 32 >= p.N) break;
        float v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(bounce + lane * BP + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        __syncwarp();
        if (p.vecC) {
            // 8 lanes cover the 32 columns of one row with 16-byte accesses, 4 rows per pass, 8 passes; every global
            // load of a batch of passes is issued before the first dependent store (one DRAM round trip per batch)
            const int c4 = lane & 7, rsub = lane >> 3;
            const int n = n0 + cb * 32 + 4 * c4;
            if (n < p.N && cb * 32 + 4 * c4 < BN && rows > 0) {
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : z4;
                if (split && !first_slice) b4 = z4;
                constexpr int IB = G ? 4 : 8;     // passes per batch
#pragma unroll
                for (int h = 0; h < 8; h += IB) {
                    float4 r_res[IB], r_gg[IB], r_acc[IB], r_pe[IB];
                    if (!split) {
#pragma unroll
                        for (int i = 0; i < IB; ++i) {
                            const int rr = (h + i) * 4 + rsub;
                            if (rr < rows) {
                                const long long m = m_first + rr;
                                if (f_res) r_res[i] = *reinterpret_cast<const float4*>(p.residual + m * p.ldr + n);
                                if (f_gg) r_gg[i] = *reinterpret_cast<const float4*>(p.gelu_grad_src + m * p.ldg + n);
                                if (f_acc) r_acc[i] = *reinterpret_cast<const float4*>(p.C + m * p.ldc + n);
                                if (f_pe) r_pe[i] = __ldg(reinterpret_cast<const float4*>(p.pe + (long long)(m % p.pe_period) * p.N + n));
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < IB; ++i) {
                        const int rr = (h + i) * 4 + rsub;
                        if (rr >= rows) continue;
                        const long long m = m_first + rr;
                        const float4 a4 = *reinterpret_cast<const float4*>(bounce + rr * BP + 4 * c4);
                        float x[4] = {fmaf(a4.x, p.alpha, b4.x), fmaf(a4.y, p.alpha, b4.y), fmaf(a4.z, p.alpha, b4.z),
                                      fmaf(a4.w, p.alpha, b4.w)};
                        float* cp = p.C + m * p.ldc + n;
                        if (split) {
                            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(x[0]), "f"(x[1]),
                                         "f"(x[2]), "f"(x[3])
                                         : "memory");
                            continue;
                        }
                        if (f_pre) *reinterpret_cast<float4*>(p.pre_out + m * p.ldp + n) = make_float4(x[0], x[1], x[2], x[3]);
                        if (f_gelu) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) x[e] = gelu_erf(x[e]);
                        }
                        if (f_gg) {
                            x[0] *= gelu_erf_grad(r_gg[i].x); x[1] *= gelu_erf_grad(r_gg[i].y);
                            x[2] *= gelu_erf_grad(r_gg[i].z); x[3] *= gelu_erf_grad(r_gg[i].w);
                        }
                        if (f_drop) {
                            const uint64_t didx = (uint64_t)m * (uint64_t)p.N + (uint64_t)n;
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                x[e] *= dropout_scale(dseed, p.drop.stream, didx + e, p.drop.thresh, p.drop.inv_keep);
                        }
                        if (f_pe) { x[0] += r_pe[i].x; x[1] += r_pe[i].y; x[2] += r_pe[i].z; x[3] += r_pe[i].w; }
                        if (f_res) { x[0] += r_res[i].x; x[1] += r_res[i].y; x[2] += r_res[i].z; x[3] += r_res[i].w; }
                        if (f_acc) { x[0] += r_acc[i].x; x[1] += r_acc[i].y; x[2] += r_acc[i].z; x[3] += r_acc[i].w; }
                        *reinterpret_cast<float4*>(cp) = make_float4(x[0], x[1], x[2], x[3]);
                    }
                }
            }
            __syncwarp();
            continue;
        }
        // scalar path (unaligned pointers / pitches, N not a multiple of 4): one column per lane, row by row
        const int n = n0 + cb * 32 + lane;
        const bool n_ok = n < p.N && (cb * 32 + lane) < BN;
        if (n_ok && rows > 0) {
            const float bias = p.bias ? __ldg(p.bias + n) : 0.f;
            float* cp = p.C + (long long)m_first * p.ldc + n;
            if (split) {
                const float b2 = first_slice ? bias : 0.f;
                for (int rr = 0; rr < rows; ++rr, cp += p.ldc) atomicAdd(cp, fmaf(bounce[rr * BP + lane], p.alpha, b2));
            } else {
                float* prep = f_pre ? p.pre_out + (long long)m_first * p.ldp + n : nullptr;
                const float* ggp = f_gg ? p.gelu_grad_src + (long long)m_first * p.ldg + n : nullptr;
                const float* resp = f_res ? p.residual + (long long)m_first * p.ldr + n : nullptr;
                int pe_row = f_pe ? m_first % p.pe_period : 0;
                const float* pep = f_pe ? p.pe + (long long)pe_row * p.N + n : nullptr;
                uint64_t didx = (uint64_t)m_first * (uint64_t)p.N + (uint64_t)n;
#pragma unroll 4
                for (int rr = 0; rr < rows; ++rr) {
                    float x = fmaf(bounce[rr * BP + lane], p.alpha, bias);
                    if (f_pre) { *prep = x; prep += p.ldp; }
                    if (f_gelu) x = gelu_erf(x);
                    if (f_gg) { x *= gelu_erf_grad(*ggp); ggp += p.ldg; }
                    if (f_drop) { x *= dropout_scale(dseed, p.drop.stream, didx, p.drop.thresh, p.drop.inv_keep); didx += (uint64_t)p.N; }
                    if (f_pe) {
                        x += __ldg(pep);
                        if (++pe_row == p.pe_period) { pe_row = 0; pep = p.pe + n; } else pep += p.N;
                    }
                    if (f_res) { x += *resp; resp += p.ldr; }
                    if (f_acc) x += *cp;
                    *cp = x;
                    cp += p.ldc;
                }
            }
        }
        __syncwarp();
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

template <bool A_KM, bool B_KN, int EPI, bool MNMAJ>
int launch_impl(dim3 grid, size_t smem, cudaStream_t st, const TcParams& p) {
    auto kernel = gemm_tc_kernel<A_KM, B_KN, EPI, MNMAJ>;
    static bool configured = false;   // one flag per instantiation
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     STAGES * (A_STAGE_BYTES + 256 * TK * 2) + 1024));
        configured = true;
    }
    SD_CUDA(launch_chain(kernel, grid, dim3(NT), smem, st, p));
    return SD_OK;
}

bool use_mn_major() {   // SD_B200_TC_MNMAJOR=0 selects the transposing K-major staging for A_KM / B_KN operands
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SD_B200_TC_MNMAJOR");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

template <bool A_KM, bool B_KN, int EPI>
int launch(dim3 grid, size_t smem, cudaStream_t st, const TcParams& p) {
    if ((A_KM || B_KN) && use_mn_major()) return launch_impl<A_KM, B_KN, EPI, true>(grid, smem, st, p);
    return launch_impl<A_KM, B_KN, EPI, false>(grid, smem, st, p);
}

int epilogue_mask(const TcParams& p) {
    return (p.pre_out ? EPI_PRE : 0) | (p.act == SD_ACT_GELU ? EPI_GELU : 0) | (p.gelu_grad_src ? EPI_GG : 0) |
           (p.drop.thresh != 0 ? EPI_DROP : 0) | (p.pe ? EPI_PE : 0) | (p.residual ? EPI_RES : 0) |
           (p.accumulate ? EPI_ACC : 0);
}

}  // namespace

int sd_gemm_tc_dispatch(const sd_gemm_desc* d, void* stream) {
    TcParams p;
    p.A = d->A; p.lda = d->lda; p.B = d->B; p.ldb = d->ldb; p.C = d->C; p.ldc = d->ldc;
    p.M = d->M; p.N = d->N; p.K = d->K;
    p.vecA = aligned16(d->A) && (d->lda % 4 == 0);
    p.vecB = aligned16(d->B) && (d->ldb % 4 == 0);
    p.ln_mean = d->ln_mean; p.ln_rstd = d->ln_rstd; p.ln_gamma = d->ln_gamma; p.ln_beta = d->ln_beta;
    const bool ln = d->ln_mean != nullptr;
    if (ln && (!d->ln_rstd || !d->ln_gamma || !d->ln_beta)) return SD_ERR_BAD_ARG;
    p.ln_on_a = ln && d->a_layout == SD_LAYOUT_MK;
    p.ln_on_b = ln && d->a_layout == SD_LAYOUT_KM && d->b_layout == SD_LAYOUT_KN;
    if (ln && !p.ln_on_a && !p.ln_on_b) return SD_ERR_UNSUPPORTED;
    p.bias = d->bias; p.residual = d->residual; p.ldr = d->ldr; p.pe = d->pe;
    p.pe_period = d->pe_period > 0 ? d->pe_period : 1;
    p.pre_out = d->pre_out; p.ldp = d->ldp; p.gelu_grad_src = d->gelu_grad_src; p.ldg = d->ldg;
    p.act = d->act; p.accumulate = d->accumulate; p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
    p.drop = make_dropout(d->dropout_p, d->dropout_seed, d->dropout_stream);
    auto vec_ok = [](const void* ptr, long long ld) { return ptr == nullptr || (aligned16(ptr) && ld % 4 == 0); };
    p.vecC = d->N % 4 == 0 && vec_ok(d->C, d->ldc) && vec_ok(d->bias, 0) && vec_ok(d->residual, d->ldr) &&
             vec_ok(d->pe, 0) && vec_ok(d->pre_out, d->ldp) && vec_ok(d->gelu_grad_src, d->ldg);

    // tile columns: one tile when N <= 256, otherwise 128-wide tiles (multiple of 16 for UMMA M=128)
    const int n16 = ((d->N + 15) / 16) * 16;
    p.BN = n16 <= 256 ? n16 : 128;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.BN) p.tmem_cols *= 2;
    const int ntn = ceil_div(d->N, p.BN), ntm = ceil_div(d->M, TM);

    int slices = 1;
    if (d->a_layout == SD_LAYOUT_KM) {
        // wgrad: small output, long contraction over the token dimension -> split K (atomics) to fill the GPU
        if (d->pre_out || d->act != SD_ACT_NONE || d->gelu_grad_src || d->residual || d->pe || d->dropout_p > 0.f)
            return SD_ERR_UNSUPPORTED;
        const int tiles = ntn * ntm;
        slices = max(1, min(ceil_div(d->K, 4 * TK), (148 * 2 + tiles - 1) / tiles));
    }
    int kps = ceil_div(d->K, slices);
    kps = ceil_div(kps, TK) * TK;
    slices = ceil_div(d->K, kps);
    p.k_per_slice = kps;
    cudaStream_t st = (cudaStream_t)stream;
    if (slices > 1 && !d->accumulate)
        SD_CUDA(cudaMemset2DAsync(d->C, d->ldc * sizeof(float), 0, (size_t)d->N * sizeof(float), d->M, st));
    dim3 grid(ntn, ntm, slices);
    const size_t smem = (size_t)STAGES * (A_STAGE_BYTES + ((p.BN + 63) & ~63) * TK * 2) + 1024;
    const int em = slices > 1 ? 0 : epilogue_mask(p);
    if (d->a_layout == SD_LAYOUT_MK && d->b_layout == SD_LAYOUT_NK) {
        switch (em) {
            case 0: return launch<false, false, 0>(grid, smem, st, p);
            case EPI_PE: return launch<false, false, EPI_PE>(grid, smem, st, p);
            case EPI_RES: return launch<false, false, EPI_RES>(grid, smem, st, p);
            case EPI_RES | EPI_DROP: return launch<false, false, EPI_RES | EPI_DROP>(grid, smem, st, p);
            case EPI_PRE | EPI_GELU: return launch<false, false, EPI_PRE | EPI_GELU>(grid, smem, st, p);
            case EPI_PRE | EPI_GELU | EPI_DROP: return launch<false, false, EPI_PRE | EPI_GELU | EPI_DROP>(grid, smem, st, p);
            default: return launch<false, false, EPI_GENERIC>(grid, smem, st, p);
        }
    }
    if (d->a_layout == SD_LAYOUT_MK && d->b_layout == SD_LAYOUT_KN) {
        switch (em) {
            case 0: return launch<false, true, 0>(grid, smem, st, p);
            case EPI_GG: return launch<false, true, EPI_GG>(grid, smem, st, p);
            case EPI_GG | EPI_DROP: return launch<false, true, EPI_GG | EPI_DROP>(grid, smem, st, p);
            case EPI_ACC: return launch<false, true, EPI_ACC>(grid, smem, st, p);
            default: return launch<false, true, EPI_GENERIC>(grid, smem, st, p);
        }
    }
    if (d->a_layout == SD_LAYOUT_KM && d->b_layout == SD_LAYOUT_KN) return launch<true, true, EPI_GENERIC>(grid, smem, st, p);
    if (d->a_layout == SD_LAYOUT_KM && d->b_layout == SD_LAYOUT_NK) return launch<true, false, EPI_GENERIC>(grid, smem, st, p);
    return SD_ERR_BAD_ARG;
}
