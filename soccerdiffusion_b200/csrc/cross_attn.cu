// The decoder layer's CROSS-ATTENTION block on tcgen05 tensor cores, fed by TMA (bf16 operands, fp32 accumulation in TMEM),
// d_model = 128, 4 heads of 32, T <= 16 query tokens per sample, memory length M <= 384 (default.yaml: T = 10, M = 312):
//
//     y = x + Drop(OutProj(MHA(LN2 x, mem, mem)))        torch/nn/modules/transformer.py:1137-1139 (norm_first),
//                                                        reference call sites ml/model/decoder.py:25-54, model.py:176-179
//
// The memory is NOT layer-normed and is shared by all decoder layers, so its K / V projections of ALL layers are one GEMM
// (sd_kv_proj_bf16: [B*M][128] x [L*256][128]^T -> bf16 [B*M][L*256], TMA in, tcgen05, bias in the epilogue); its
// backward is one GEMM too (sd_kv_dgrad_bf16: dmem = dKV_all x W_all, K = L*256) plus sd_wgrad_bf16 jobs.
//
// sd_ca_block_fwd / sd_ca_block_bwd: one CTA (128 threads = the 128 TMEM lanes) per SAMPLE.  With only T = 10 queries the
// operand roles are swapped so that no MMA wastes its 128 rows on 10 tokens:
//   * projections are computed transposed, D^T[feature][t] = W[feature][:] . act[t][:]  (A = the weight matrix, N = 16):
//     thread = feature, every global access "row t, thread = column" is coalesced;
//   * scores are computed transposed for all heads at once, S^T[m][(h,t)] = K[m][:] . Qblk[(h,t)][:], where Qblk is the
//     block-diagonal expansion of Q (row (h,t) holds Q[t] on the columns of head h, zeros elsewhere): A = 128 keys of the
//     TMA-loaded K tile (the natural row-major layout), N = 4 heads x 16 = 64.  Softmax runs over the TMEM LANES (keys):
//     register-transposing warp reductions + one shared-memory exchange;
//   * O^T[c][(h,t)] = sum_m V[m][c] P^T[m][(h,t)]: A = the TMA-loaded V tile used MN-major, B = P^T written row per key.
// Backward uses the same forms: dP^T = V dOblk^T, dS^T elementwise per key, dV = Pd^T dOblk, dK = dS^T Qblk (N = 128),
// dQ^T += K^T dS^T, chunk by chunk over the keys (no cross-chunk dependency: the forward saved the log-sum-exp).
#include "layer_common.cuh"
#include "../../include/sd_b200.h"

using namespace sdlf;

namespace {

constexpr int CNT = 128;            // threads per CTA
constexpr int TP = 16;              // padded query tokens per sample
constexpr int NQ = 64;              // 4 heads x TP columns of the transposed score matrix
constexpr int CA_H = 4, CA_DH = 32;
constexpr int MAXCH = 3;            // key chunks of 128

// lane l ends with the reduction over the warp of v[l] (31 shuffles instead of 160)
template <bool MAX>
__device__ __forceinline__ float warp_reduce32_cols(float (&v)[32], int lane) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int k = 0; k < o; ++k) {
            const float keep = up ? v[k + o] : v[k];
            const float send = up ? v[k] : v[k + o];
            const float recv = __shfl_xor_sync(0xffffffffu, send, o);
            v[k] = MAX ? fmaxf(keep, recv) : keep + recv;
        }
    }
    return v[0];
}

__device__ __forceinline__ void ld_lane32(uint32_t tmem, int warp, int col, float* v) {
    tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, v);
}
__device__ __forceinline__ void ld_lane16(uint32_t tmem, int warp, int col, float* v) {
    tmem_ld_32x16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col, v);
}
// bf16 element (row r, column c) of a 128-wide K-major operand made of two [rows][64] swizzled tiles `tile_bytes` apart
__device__ __forceinline__ uint32_t elem_off(int r, int c, int tile_bytes) {
    return (uint32_t)((c >> 6) * tile_bytes) + sw128_chunk_off(r, (c & 63) >> 3) + (uint32_t)((c & 7) * 2);
}
__device__ __forceinline__ unsigned short bf16_bits(float f) {
    const __nv_bfloat16 h = __float2bfloat16_rn(f);
    return *reinterpret_cast<const unsigned short*>(&h);
}
__device__ __forceinline__ float bf16_to_f(unsigned short u) {
    return __uint_as_float(((uint32_t)u) << 16);
}

// =====================================================================================================================
// fp32 -> bf16 (the memory / any activation matrix): 8 elements per thread
__global__ void cast_bf16_kernel(const float* __restrict__ src, uint4* __restrict__ dst, long long n8) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const float4 a = reinterpret_cast<const float4*>(src)[2 * i], b = reinterpret_cast<const float4*>(src)[2 * i + 1];
        const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        dst[i] = pack8_bf16(f);
    }
}

// one bf16 row broadcast into row `row` of each of the B blocks of `block_rows` rows (the step token's K | V of a DDIM step)
__global__ void bcast_row_bf16_kernel(uint4* __restrict__ dst, long long ld8, long long block_rows, long long row, int B,
                                      const uint4* __restrict__ src, int n8) {
    const int b = blockIdx.x;
    pdl_trigger();
    pdl_wait();
    if (b >= B) return;
    uint4* d = dst + ((long long)b * block_rows + row) * ld8;
    for (int i = threadIdx.x; i < n8; i += blockDim.x) d[i] = src[i];
}

// =====================================================================================================================
// Between two denoiser evaluations of the batched DDIM loop (ros.py:301-310): output projection of step s, the eta = 0
// scheduler update, the embedding (+ positional encoding) of step s+1 and the broadcast of step s+1's step-token K | V row,
// as ONE launch instead of four (the loop is a latency-bound chain of short kernels).  fp32 CUDA-core arithmetic: the two
// projections have J = 20 columns / K = 20.
//   eps = h fc_w^T + fc_b ;  x0 = (x - sb eps) / sa ;  x' = sap x0 + sbp eps ;  h' = x' emb_w^T + emb_b + PE
constexpr int GL_ROWS = 8, GL_NT = 256, GL_JMAX = 32;
struct GlueParams {
    const float *h, *fc_w, *fc_b, *x;
    float *x_next, *eps_out;
    float sb, sa, sap, sbp;
    const float *emb_w, *emb_b, *pe;   // emb_w == nullptr: last step, no next embedding
    float* h_next;
    int rows, J, T, row_blocks;
    uint4* kv; long long ld8, block_rows, row; int B; const uint4* src; int n8;   // src == nullptr: no broadcast
};
__global__ void __launch_bounds__(GL_NT) ddim_glue_kernel(const GlueParams p) {
    __shared__ float fcw[GL_JMAX][129];
    __shared__ float embw[128][GL_JMAX + 1];
    __shared__ float hs[GL_ROWS][128];
    __shared__ float xs[GL_ROWS][GL_JMAX];
    const int tid = threadIdx.x;
    pdl_trigger();
    if ((int)blockIdx.x >= p.row_blocks) {   // broadcast CTAs: one sample each
        const int b = blockIdx.x - p.row_blocks;
        pdl_wait();
        if (p.src && b < p.B) {
            uint4* d = p.kv + ((long long)b * p.block_rows + p.row) * p.ld8;
            for (int i = tid; i < p.n8; i += GL_NT) d[i] = p.src[i];
        }
        return;
    }
    // parameters (never written inside the chain) before the wait
    const int J = p.J;
    for (int i = tid; i < J * 128; i += GL_NT) fcw[i >> 7][i & 127] = __ldg(p.fc_w + i);
    if (p.emb_w)
        for (int i = tid; i < 128 * J; i += GL_NT) embw[i / J][i % J] = __ldg(p.emb_w + i);
    pdl_wait();
    const int r0 = blockIdx.x * GL_ROWS;
    for (int i = tid; i < GL_ROWS * 128; i += GL_NT) {
        const int r = r0 + (i >> 7);
        hs[i >> 7][i & 127] = r < p.rows ? p.h[(long long)r * 128 + (i & 127)] : 0.f;
    }
    __syncthreads();
    {
        const int r = tid >> 5, j = tid & 31, gr = r0 + r;
        if (j < J && gr < p.rows) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
            for (int k = 0; k < 128; k += 4) {
                a0 = fmaf(hs[r][k], fcw[j][k], a0); a1 = fmaf(hs[r][k + 1], fcw[j][k + 1], a1);
                a2 = fmaf(hs[r][k + 2], fcw[j][k + 2], a2); a3 = fmaf(hs[r][k + 3], fcw[j][k + 3], a3);
            }
            const float e = (a0 + a1) + (a2 + a3) + __ldg(p.fc_b + j);
            const long long idx = (long long)gr * J + j;
            const float x0 = (p.x[idx] - p.sb * e) / p.sa;
            const float xn = p.sap * x0 + p.sbp * e;
            if (p.eps_out) p.eps_out[idx] = e;
            p.x_next[idx] = xn;
            xs[r][j] = xn;
        }
    }
    if (!p.emb_w) return;
    __syncthreads();
    {
        const int n = tid & 127;
        const float eb = __ldg(p.emb_b + n);
#pragma unroll
        for (int i = 0; i < GL_ROWS / 2; ++i) {
            const int r = (tid >> 7) + 2 * i, gr = r0 + r;
            if (gr < p.rows) {
                float a = eb + __ldg(p.pe + (long long)(gr % p.T) * 128 + n);
                for (int j = 0; j < J; ++j) a = fmaf(xs[r][j], embw[n][j], a);
                p.h_next[(long long)gr * 128 + n] = a;
            }
        }
    }
}

// =====================================================================================================================
// K/V projection of the memory for all layers: C[row][256 j + n] = sum_k A[row][k] W[w_row0 + j w_stride + n][k] + bias_j[n]
struct KvProjParams {
    long long rows;
    int nblk, w_row0, w_stride;
    const float* bias[SD_KV_MAX_LAYERS];
    __nv_bfloat16* C;
    long long ldc;
};
constexpr int KP_NT = 320;                                          // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue
constexpr int KP_SMEM = 4 * LTILE + 2 * 2 * LTILE + 4 * LTILE + 1024;   // resident weight block [256][128] + two A tiles + output staging

// TMA tile store (shared -> global, SASS: UTMASTG); completion tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src_smem, int col0, int row0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm), "r"(col0), "r"(row0),
                 "r"(src_smem)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Persistent and weight-stationary: CTA (block j, group g) keeps the 256 weight rows of block j (64 KB) in shared memory and
// streams the row tiles g, g + G, ... of A through a 2-slot TMA ring; two TMEM accumulators of 256 columns let the MMAs of
// tile i+1 run under the epilogue of tile i (8 warps: TMEM -> + bias -> bf16 -> swizzled staging -> TMA store).
__global__ void __launch_bounds__(KP_NT, 1) kv_proj_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                           const __grid_constant__ CUtensorMap tmC, const KvProjParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_w, a_full[2], a_empty[2], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int j = blockIdx.x % p.nblk, g = blockIdx.x / p.nblk, G = gridDim.x / p.nblk;
    const int ntiles_all = (int)((p.rows + 127) / 128);
    const int nt = g < ntiles_all ? (ntiles_all - g + G - 1) / G : 0;   // tiles g, g + G, ...
    if (warp == 0 && elect_one()) {
        mbar_init(&bar_w, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 1); }
        mbar_fence_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmC);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t OFF_A = 4 * LTILE, OFF_C = 8 * LTILE;

    if (warp == 0 && elect_one()) {          // ---- TMA producer: the weight block once, then the A tiles ------------------
        const int wr = p.w_row0 + j * p.w_stride;
        mbar_arrive_expect_tx(&bar_w, 4 * LTILE);
        tma_tile_2d(sbase, &tmW, 0, wr, &bar_w);                    // tiles [k half][row half]: 256 rows x 64 k contiguous
        tma_tile_2d(sbase + LTILE, &tmW, 0, wr + 128, &bar_w);
        tma_tile_2d(sbase + 2 * LTILE, &tmW, 64, wr, &bar_w);
        tma_tile_2d(sbase + 3 * LTILE, &tmW, 64, wr + 128, &bar_w);
        for (int i = 0; i < nt; ++i) {
            const int s = i & 1;
            mbar_wait(&a_empty[s], (uint32_t)(((i >> 1) & 1) ^ 1));
            const int row0 = (g + i * G) * 128;
            mbar_arrive_expect_tx(&a_full[s], 2 * LTILE);
            tma_tile_2d(sbase + OFF_A + s * 2 * LTILE, &tmA, 0, row0, &a_full[s]);
            tma_tile_2d(sbase + OFF_A + s * 2 * LTILE + LTILE, &tmA, 64, row0, &a_full[s]);
        }
    } else if (warp == 1 && elect_one()) {   // ---- MMA issuer ---------------------------------------------------------------
        const uint32_t id256 = instr_desc_bf16(128, 256);
        mbar_wait(&bar_w, 0);
        for (int i = 0; i < nt; ++i) {
            const int s = i & 1;
            mbar_wait(&a_full[s], (uint32_t)((i >> 1) & 1));
            mbar_wait(&acc_empty[s], (uint32_t)(((i >> 1) & 1) ^ 1));
            tc_fence_after_sync();
            mma_k_tiles(tmem + s * 256, sbase + OFF_A + s * 2 * LTILE, LTILE, sbase, 2 * LTILE, id256, 2, false);
            mma_commit(&a_empty[s]);
            mma_commit(&acc_full[s]);
        }
    } else if (warp >= 2) {                // ---- epilogue: 8 warps, lane quarter = warp % 4, column half = (warp - 2) / 4 ---
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int r = q * 32 + lane;
        const float* bias = p.bias[j];
        for (int i = 0; i < nt; ++i) {
            const int s = i & 1;
            mbar_wait(&acc_full[s], (uint32_t)((i >> 1) & 1));
            tc_fence_after_sync();
            if (tid == 64) tma_store_wait_read();   // the previous tile's stores have read the staging tiles
            asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll 1
            for (int qq = 0; qq < 4; ++qq) {
                const int c0 = 128 * half + 32 * qq;
                float v[32], b[32];
                if (bias) ldg32(bias + c0, b);
                tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * 256 + c0), v);
                if (bias) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) v[k] += b[k];
                }
                uint8_t* tile = smem + OFF_C + (2 * half + (qq >> 1)) * LTILE;
#pragma unroll
                for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(tile + sw128_chunk_off(r, 4 * (qq & 1) + k)) = pack8_bf16(v + 8 * k);
            }
            fence_proxy_async_smem();
            tc_fence_before_sync();
            asm volatile("bar.sync 1, 256;" ::: "memory");   // staging complete, accumulator s drained by all eight warps
            if (tid == 64) {
                mbar_arrive(&acc_empty[s]);
                const int row0 = (g + i * G) * 128;
#pragma unroll
                for (int t = 0; t < 4; ++t) tma_store_2d(&tmC, sbase + OFF_C + t * LTILE, 256 * j + 64 * t, row0);
                tma_store_commit();
            }
        }
        if (tid == 64) tma_store_wait_all();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// =====================================================================================================================
// C[row][n] (+)= sum_k A[row][k] W[k][n]: A bf16 [rows][K] K-major by TMA, W bf16 [K][N] used MN-major straight from its
// row-major matrix ([64 k][64 n] TMA boxes), N = 64 / 128 / 256, 4-stage TMA ring, warp-specialised producer / MMA issuer.
//   * sd_kv_dgrad_bf16: dmem = dKV_all x W_all (k tile i -> packed weight row w_row0 + (i / 4) * w_stride + (i % 4) * 64), fp32 out;
//   * sd_conv1x1s2_dgrad_bf16: data gradient of a 1x1 stride-2 convolution (the ResNet downsample path): row p = output pixel
//     (n, ho, wo), result written as the bf16 NHWC input pixel (n, 2 ho, 2 wo); the other pixels are zeroed beforehand.
struct MnGemmParams {
    long long rows;
    int ktiles, N;
    int w_row0, w_stride, tiles_per_blk;
    float* Cf; long long ldc; int accumulate;
    __nv_bfloat16* Cb; int Ho, Wo, Hin, Win;
};
constexpr int KD_STAGES = 4;
inline int kd_stage_bytes(int N) { return LTILE + N * 128; }   // A [128 rows][64 k] + W [64 k][N]

__global__ void __launch_bounds__(CNT, 1) mn_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                                                         const MnGemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[KD_STAGES], bar_empty[KD_STAGES], bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long row0 = (long long)blockIdx.x * 128;
    const int N = p.N, stage = LTILE + N * 128;
    if (warp == 0 && elect_one()) {
        for (int s = 0; s < KD_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, (uint32_t)N);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int nk = p.ktiles;
    if (warp == 0 && elect_one()) {          // TMA producer
        for (int i = 0; i < nk; ++i) {
            const int s = i % KD_STAGES;
            mbar_wait(&bar_empty[s], (uint32_t)(((i / KD_STAGES) & 1) ^ 1));
            const uint32_t dst = sbase + s * stage;
            const int wrow = p.w_row0 + (i / p.tiles_per_blk) * p.w_stride + (i % p.tiles_per_blk) * 64;
            mbar_arrive_expect_tx(&bar_full[s], (uint32_t)stage);
            tma_tile_2d(dst, &tmA, 64 * i, (int)row0, &bar_full[s]);
            for (int nb = 0; nb < N / 64; ++nb) tma_tile_2d(dst + LTILE + nb * (LTILE / 2), &tmW, 64 * nb, wrow, &bar_full[s]);
        }
    } else if (warp == 1 && elect_one()) {   // MMA issuer: A K-major, B MN-major ([64 k rows][64 n] blocks 8 KB apart)
        const uint32_t idesc = instr_desc_bf16(128, N, 0, 1);
        for (int i = 0; i < nk; ++i) {
            const int s = i % KD_STAGES;
            mbar_wait(&bar_full[s], (uint32_t)((i / KD_STAGES) & 1));
            tc_fence_after_sync();
            mma_a_k_b_mn(tmem, sbase + s * stage, 0, sbase + s * stage + LTILE, LTILE / 2, idesc, 4, i > 0);
            mma_commit(&bar_empty[s]);
        }
        mma_commit(&bar_done);
    }
    __syncwarp();
    mbar_wait(&bar_done, 0);
    tc_fence_after_sync();
    const long long grow = row0 + tid;
    const bool rv = grow < p.rows;
    if (p.Cb) {
        __nv_bfloat16* dst = nullptr;
        if (rv) {
            const long long hw = (long long)p.Ho * p.Wo;
            const long long n = grow / hw;
            const int rem = (int)(grow - n * hw), ho = rem / p.Wo, wo = rem - ho * p.Wo;
            dst = p.Cb + ((n * p.Hin + 2 * ho) * p.Win + 2 * wo) * N;
        }
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            ld_lane32(tmem, warp, c0, v);   // warp-collective: every lane takes part
            if (rv) {
                uint4* g = reinterpret_cast<uint4*>(dst + c0);
#pragma unroll
                for (int i = 0; i < 4; ++i) g[i] = pack8_bf16(v + 8 * i);
            }
        }
    } else {
        float* crow = p.Cf + grow * p.ldc;
#pragma unroll 1
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            ld_lane32(tmem, warp, c0, v);
            if (rv) {
                float4* g = reinterpret_cast<float4*>(crow + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    if (p.accumulate) { const float4 t = g[i]; o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w; }
                    g[i] = o;
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, (uint32_t)N);
}

// =====================================================================================================================
// forward block.  Shared memory is kept near 100 KB and TMEM at 256 columns so that TWO CTAs share an SM (one CTA is a chain of
// dependent phases: the second one fills its waits): every 32 KB operand — Wq, the K chunks, the V chunks, Wout — streams
// through ONE 2-slot TMA ring in the order it is consumed (a dedicated producer warp refills a slot as soon as the MMAs that
// read it have completed), P^T is produced chunk by chunk into two tiles (the first one takes Qblk's place once the scores
// exist), and the small projection accumulators alias the score columns.
constexpr int F_OFF_RING = 0;                    // 2 slots x (lo, hi) [128][64] tiles
constexpr int F_OFF_QB = 4 * LTILE;              // Qblk: two [64][64] tiles of 8 KB; after the scores: P^T buffer 0
constexpr int F_OFF_P1 = F_OFF_QB + LTILE;       // P^T buffer 1
constexpr int F_OFF_XB = F_OFF_P1 + LTILE;       // LN2(x) / attention output as B operand: two [16][64] tiles of 2 KB
constexpr int F_SMEM = F_OFF_XB + 4096 + 1024;
constexpr int XB_TILE = 2048, QB_TILE = 8192;
constexpr int F_NT = CNT + 32;                   // warps 0-3: the 128 TMEM lanes, warp 4: TMA producer

struct CaFwdParams {
    const float* x;
    float* y;
    int B, T, M;
    int groups, Tg;            // query-row groups per sample, rows per group (<= 16)
    int w_row_q, w_row_o;      // packed-weight rows of Wq and Wout
    int kv_col0;               // first column of this layer's K | V inside the all-layer K/V matrix
    const float *q_b, *out_b, *n_w, *n_b;
    __nv_bfloat16 *xn_save, *q_save, *attn_save;
    float *stats_save, *lse_save;   // [B*T][2] (mean, rstd) ; [B][4][T] log2-domain log-sum-exp
    Dropout drop;              // stream + 0: attention probabilities, + 1: out-projection
};

enum { FB_FULL0 = 0, FB_EMPTY0 = 2, FB_Q = 4, FB_S, FB_O, FB_Y, FB_PD0, FB_N = FB_PD0 + 2 };

template <bool DROP>
__global__ void __launch_bounds__(F_NT, 2) ca_fwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmKV,
                                                         const CaFwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar[FB_N];
    __shared__ float red[4][NQ];
    __shared__ __align__(16) float fin_max[NQ];
    __shared__ __align__(16) float fin_inv[NQ];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // a sample's T query rows are handled in p.groups groups of p.Tg <= 16 rows, one CTA each (the rows of a cross-attention
    // are independent): T = this group's row count, Tf = the sample's, t0 = the group's first row
    const int b = blockIdx.x / p.groups, t0 = (blockIdx.x % p.groups) * p.Tg, Tf = p.T, T = min(p.Tg, Tf - t0), M = p.M;
    const long long rowb = (long long)b * Tf + t0;
    const int nch = (M + 127) >> 7;
    const long long krow0 = (long long)b * M;

    if (warp == 0 && elect_one()) {
        for (int i = 0; i < FB_N; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmKV);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t ST = 0, OT = 192, QT = 0, YT = 0;   // the projection accumulators alias the score columns
    // ring items in consumption order: 0 Wq | 1 .. nch K chunks | nch+1 .. 2 nch V chunks | 2 nch + 1 Wout ; slot = item & 1
    const int IT_K = 1, IT_V = 1 + nch, IT_WO = 1 + 2 * nch;
    auto slot_addr = [&](int item) { return sbase + F_OFF_RING + (uint32_t)(item & 1) * 2 * LTILE; };
    auto wait_item = [&](int item) { mbar_wait(&bar[FB_FULL0 + (item & 1)], (uint32_t)((item >> 1) & 1)); };
    auto free_item = [&](int item) { mma_commit(&bar[FB_EMPTY0 + (item & 1)]); };   // when the MMAs issued so far have completed

    pdl_trigger();
    if (warp == 4) {
        // ---- TMA producer ---------------------------------------------------------------------------------------------------
        if (elect_one()) {
            for (int item = 0; item <= IT_WO; ++item) {
                const int s = item & 1;
                if (item >= 2) mbar_wait(&bar[FB_EMPTY0 + s], (uint32_t)(((item >> 1) & 1) ^ 1));
                if (item == 1) pdl_wait();   // K | V come from preceding kernels of the chain (the packed weights do not)
                const uint32_t dst = slot_addr(item);
                mbar_arrive_expect_tx(&bar[FB_FULL0 + s], 2 * LTILE);
                if (item == 0 || item == IT_WO) {
                    const int row = item == 0 ? p.w_row_q : p.w_row_o;
                    tma_tile_2d(dst, &tmW, 0, row, &bar[FB_FULL0 + s]);
                    tma_tile_2d(dst + LTILE, &tmW, 64, row, &bar[FB_FULL0 + s]);
                } else {
                    const bool isv = item >= IT_V;
                    const int j = isv ? item - IT_V : item - IT_K;
                    const int col = p.kv_col0 + (isv ? 128 : 0);
                    tma_tile_2d(dst, &tmKV, col, (int)(krow0 + 128 * j), &bar[FB_FULL0 + s]);
                    tma_tile_2d(dst + LTILE, &tmKV, col + 64, (int)(krow0 + 128 * j), &bar[FB_FULL0 + s]);
                }
            }
        }
    } else {
    // Qblk starts as zeros (the off-diagonal blocks stay zero)
    for (int i = tid; i < LTILE / 16; i += CNT) reinterpret_cast<uint4*>(smem + F_OFF_QB)[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint64_t dseed = DROP ? p.drop.resolve() : 0ull;
    pdl_wait();   // the residual stream comes from the preceding kernel of the chain

    // ---- LN2 of the T rows (warp per row, lane = 4 features) -> B operand [t][k] ------------------------------------------
    {
        const float4 ga = __ldg(reinterpret_cast<const float4*>(p.n_w) + lane);
        const float4 be = __ldg(reinterpret_cast<const float4*>(p.n_b) + lane);
        for (int t = warp; t < TP; t += 4) {
            float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
            const long long grow = rowb + t;
            if (t < T) xv = reinterpret_cast<const float4*>(p.x + grow * 128)[lane];
            const float mean = warp_sum((xv.x + xv.y) + (xv.z + xv.w)) * (1.0f / 128.0f);
            const float c0 = xv.x - mean, c1 = xv.y - mean, c2 = xv.z - mean, c3 = xv.w - mean;
            const float rstd = rsqrtf(warp_sum(fmaf(c0, c0, c1 * c1) + fmaf(c2, c2, c3 * c3)) * (1.0f / 128.0f) + LN_EPS);
            float v[4] = {0.f, 0.f, 0.f, 0.f};
            if (t < T) {
                v[0] = fmaf(c0 * rstd, ga.x, be.x); v[1] = fmaf(c1 * rstd, ga.y, be.y);
                v[2] = fmaf(c2 * rstd, ga.z, be.z); v[3] = fmaf(c3 * rstd, ga.w, be.w);
            }
            __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&lo);
            u.y = *reinterpret_cast<uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(smem + F_OFF_XB + elem_off(t, 4 * lane, XB_TILE)) = u;
            if (t < T) {
                if (p.xn_save) reinterpret_cast<uint2*>(p.xn_save + grow * 128)[lane] = u;
                if (p.stats_save && lane == 0) { p.stats_save[2 * grow] = mean; p.stats_save[2 * grow + 1] = rstd; }
            }
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // ---- Q^T[n][t] = Wq[n][:] . xn[t][:] ---------------------------------------------------------------------------------
    const uint32_t id16 = instr_desc_bf16(128, 16);
    if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        wait_item(0);
        mma_k_tiles(tmem + QT, slot_addr(0), LTILE, sbase + F_OFF_XB, XB_TILE, id16, 2, false);
        free_item(0);
        mma_commit(&bar[FB_Q]);
    }
    __syncwarp();
    {
        const float bq = __ldg(p.q_b + tid);
        float q[16];
        mbar_wait(&bar[FB_Q], 0);
        tc_fence_after_sync();
        ld_lane16(tmem, warp, QT, q);
        const int h = tid >> 5;
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const unsigned short u = t < T ? bf16_bits(q[t] + bq) : (unsigned short)0;
            *reinterpret_cast<unsigned short*>(smem + F_OFF_QB + elem_off(h * TP + t, tid, QB_TILE)) = u;
            if (t < T && p.q_save) reinterpret_cast<unsigned short*>(p.q_save)[(rowb + t) * 128 + tid] = u;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    asm volatile("bar.sync 1, 128;" ::: "memory");   // Qblk complete; Q^T columns consumed
    // ---- S^T chunks ------------------------------------------------------------------------------------------------------
    const uint32_t id64 = instr_desc_bf16(128, NQ);
    if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        for (int j = 0; j < nch; ++j) {
            wait_item(IT_K + j);
            mma_k_tiles(tmem + ST + NQ * j, slot_addr(IT_K + j), LTILE, sbase + F_OFF_QB, QB_TILE, id64, 2, false);
            free_item(IT_K + j);
        }
        mma_commit(&bar[FB_S]);
    }
    __syncwarp();
    mbar_wait(&bar[FB_S], 0);
    tc_fence_after_sync();
    // ---- softmax over the keys (TMEM lanes): thread = key 128 j + tid of every chunk ----------------------------------------
    const float sc = rsqrtf((float)CA_DH) * 1.4426950408889634f;
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
        float mx[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) mx[i] = -INFINITY;
        for (int j = 0; j < nch; ++j) {
            float s[32];
            ld_lane32(tmem, warp, ST + NQ * j + 32 * g, s);
            if (128 * j + tid < M) {
#pragma unroll
                for (int i = 0; i < 32; ++i) mx[i] = fmaxf(mx[i], s[i]);
            }
        }
        const float m = warp_reduce32_cols<true>(mx, lane);
        red[warp][32 * g + lane] = m;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (tid < NQ) fin_max[tid] = fmaxf(fmaxf(red[0][tid], red[1][tid]), fmaxf(red[2][tid], red[3][tid])) * sc;
    asm volatile("bar.sync 1, 128;" ::: "memory");
    red[warp][lane] = 0.f;        // from here on: this lane's running column sums of its warp's keys (one owner per entry, fixed
    red[warp][32 + lane] = 0.f;   // order of additions: results are bit-reproducible)
    // probabilities chunk by chunk: P^T_j -> buffer j & 1 (buffer 0 = Qblk's place: every score MMA has completed), its
    // O^T += V_j^T P^T_j issued right away; the column sums of the three chunks meet in shared memory
    const uint32_t idPV = instr_desc_bf16(128, NQ, 1, 1);
#pragma unroll 1
    for (int j = 0; j < nch; ++j) {
        uint8_t* Pb = smem + ((j & 1) ? F_OFF_P1 : F_OFF_QB);
        if (j >= 2) { mbar_wait(&bar[FB_PD0 + (j & 1)], (uint32_t)(((j >> 1) - 1) & 1)); }   // the MMAs that read this buffer are done
        const int m = 128 * j + tid;
        const bool kv = m < M;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
            float s[32], mxs[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 t4 = reinterpret_cast<const float4*>(fin_max)[8 * g + i];
                mxs[4 * i] = t4.x; mxs[4 * i + 1] = t4.y; mxs[4 * i + 2] = t4.z; mxs[4 * i + 3] = t4.w;
            }
            ld_lane32(tmem, warp, ST + NQ * j + 32 * g, s);
#pragma unroll
            for (int i = 0; i < 32; ++i) { s[i] = kv ? ex2_approx(fmaf(s[i], sc, -mxs[i])) : 0.f; mxs[i] = s[i]; }
            if (DROP && kv) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int h = (32 * g + i) >> 4, t = i & 15;
                    if (t < T)
                        s[i] *= dropout_scale(dseed, p.drop.stream, (((uint64_t)b * CA_H + h) * Tf + t0 + t) * (uint64_t)M + m, p.drop.thresh,
                                              p.drop.inv_keep);
                }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(Pb + sw128_chunk_off(tid, 4 * g + c)) = pack8_bf16(s + 8 * c);
            const float sm = warp_reduce32_cols<false>(mxs, lane);   // sums of the UNdropped probabilities
            red[warp][32 * g + lane] += sm;
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (warp == 0 && elect_one()) {
            tc_fence_after_sync();
            wait_item(IT_V + j);
            const int ks = min(8, (M - 128 * j + 15) >> 4);   // keys beyond M carry zero probabilities
            for (int k = 0; k < ks; ++k)
                mma_bf16_ss(tmem + OT, smem_desc_mn_sw128(slot_addr(IT_V + j), LTILE, 1024) + 128 * (uint64_t)k,
                            smem_desc_mn_sw128(smem_u32(Pb), 1024, 1024) + 128 * (uint64_t)k, idPV, (j > 0 || k > 0) ? 1u : 0u);
            free_item(IT_V + j);
            mma_commit(&bar[FB_PD0 + (j & 1)]);
            if (j == nch - 1) mma_commit(&bar[FB_O]);
        }
        __syncwarp();
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");   // every warp's column sums are complete
    if (tid < NQ) {
        const float sum = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
        const int h = tid >> 4, t = tid & 15;
        if (t < T && p.lse_save) p.lse_save[((long long)b * CA_H + h) * Tf + t0 + t] = fin_max[tid] + log2f(sum);
        fin_inv[tid] = 1.0f / sum;
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // residual rows of this thread's column: issued ahead of the waits
    float xr[TP];
#pragma unroll
    for (int t = 0; t < TP; ++t) xr[t] = t < T ? p.x[(rowb + t) * 128 + tid] : 0.f;
    {
        float o[16];
        const int h = tid >> 5;
        mbar_wait(&bar[FB_O], 0);
        tc_fence_after_sync();
        ld_lane16(tmem, warp, OT + TP * h, o);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const unsigned short u = t < T ? bf16_bits(o[t] * fin_inv[TP * h + t]) : (unsigned short)0;
            *reinterpret_cast<unsigned short*>(smem + F_OFF_XB + elem_off(t, tid, XB_TILE)) = u;
            if (t < T && p.attn_save) reinterpret_cast<unsigned short*>(p.attn_save)[(rowb + t) * 128 + tid] = u;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // ---- Y^T[n][t] = Wout[n][:] . attn[t][:] ; y = x + Drop(Y + b) ----------------------------------------------------------
    if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        wait_item(IT_WO);
        mma_k_tiles(tmem + YT, slot_addr(IT_WO), LTILE, sbase + F_OFF_XB, XB_TILE, id16, 2, false);
        mma_commit(&bar[FB_Y]);
    }
    __syncwarp();
    {
        const float bo = __ldg(p.out_b + tid);
        float v[16];
        mbar_wait(&bar[FB_Y], 0);
        tc_fence_after_sync();
        ld_lane16(tmem, warp, YT, v);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            if (t < T) {
                const long long e = (rowb + t) * 128 + tid;
                float a = v[t] + bo;
                if (DROP) a *= dropout_scale(dseed, p.drop.stream + 1, (uint64_t)e, p.drop.thresh, p.drop.inv_keep);
                p.y[e] = xr[t] + a;
            }
        }
    }
    }   // warps 0-3
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// =====================================================================================================================
// backward block
constexpr int B_OFF_KV = 0;                      // 2 stages x (K lo, K hi, V lo, V hi); the last free stage takes Wq
constexpr int B_OFF_QB = 8 * LTILE;              // Qblk  [64][128]
constexpr int B_OFF_DOB = B_OFF_QB + LTILE;      // dOblk [64][128]
constexpr int B_OFF_PD = B_OFF_DOB + LTILE;      // Pd^T [128 keys][64] ; first: Wout lo
constexpr int B_OFF_DS = B_OFF_PD + LTILE;       // dS^T [128 keys][64] ; first: Wout hi
constexpr int B_OFF_XB = B_OFF_DS + LTILE;       // g1 / dq as B operand: two [16][64] tiles
constexpr int B_SMEM = B_OFF_XB + 4096 + 1024;

struct CaBwdParams {
    const float* dy;
    float* dx;
    const float* x;
    const __nv_bfloat16 *q, *attn;
    const float *stats, *lse;
    int B, T, M;
    int t0, Tg, dkv_accumulate;   // this launch's query-row group
    int w_row_q, w_row_o, kv_col0;
    const float* n_w;
    __nv_bfloat16 *g1, *dq, *dkv;   // g1, dq: [B*T][128]; dkv: the all-layer dK | dV matrix (row stride lddkv), same columns as kv
    long long lddkv;
    float *g_n_w, *g_n_b;
    Dropout drop;
};

enum { BB_WO = 0, BB_WQ, BB_DA, BB_SD, BB_G, BB_X, BB_KV0, BB_N = BB_KV0 + 2 };

template <bool DROP>
__global__ void __launch_bounds__(CNT, 1) ca_bwd_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmKV,
                                                        const CaBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar[BB_N];
    __shared__ __align__(16) float s_lse[NQ];
    __shared__ __align__(16) float s_delta[NQ];
    __shared__ float red[4][32];
    __shared__ float fin[32];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // one launch per query-row group (p.t0 .. p.t0 + T): groups of one sample add into the same dK | dV rows, so they run as
    // consecutive launches, the later ones with p.dkv_accumulate
    const int b = blockIdx.x, Tf = p.T, t0 = p.t0, T = min(p.Tg, Tf - t0), M = p.M;
    const long long rowb = (long long)b * Tf + t0;
    const int nch = (M + 127) >> 7;
    const long long krow0 = (long long)b * M;

    if (warp == 0 && elect_one()) {
        for (int i = 0; i < BB_N; ++i) mbar_init(&bar[i], 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmKV);
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    // S^T / dV share columns, dP^T / dK share columns (one warpgroup works through the chunks serially)
    constexpr uint32_t SC = 0, DPC = 128, DVC = 0, DKC = 128, DQC = 256, DAC = 320, DXC = 352;

    auto load_chunk = [&](int j) {   // K and V tiles of chunk j -> stage j & 1
        const uint32_t dst = sbase + B_OFF_KV + (j & 1) * 4 * LTILE;
        const int row = (int)(krow0 + 128 * j);
        mbar_arrive_expect_tx(&bar[BB_KV0 + (j & 1)], 4 * LTILE);
        tma_tile_2d(dst, &tmKV, p.kv_col0, row, &bar[BB_KV0 + (j & 1)]);
        tma_tile_2d(dst + LTILE, &tmKV, p.kv_col0 + 64, row, &bar[BB_KV0 + (j & 1)]);
        tma_tile_2d(dst + 2 * LTILE, &tmKV, p.kv_col0 + 128, row, &bar[BB_KV0 + (j & 1)]);
        tma_tile_2d(dst + 3 * LTILE, &tmKV, p.kv_col0 + 192, row, &bar[BB_KV0 + (j & 1)]);
    };
    const uint32_t wq_addr = sbase + B_OFF_KV + (nch & 1) * 4 * LTILE;   // the stage the last chunk does NOT use
    auto load_wq = [&]() {
        mbar_arrive_expect_tx(&bar[BB_WQ], 2 * LTILE);
        tma_tile_2d(wq_addr, &tmW, 0, p.w_row_q, &bar[BB_WQ]);
        tma_tile_2d(wq_addr + LTILE, &tmW, 64, p.w_row_q, &bar[BB_WQ]);
    };
    if (warp == 0 && elect_one()) {
        mbar_arrive_expect_tx(&bar[BB_WO], 2 * LTILE);
        tma_tile_2d(sbase + B_OFF_PD, &tmW, 0, p.w_row_o, &bar[BB_WO]);
        tma_tile_2d(sbase + B_OFF_DS, &tmW, 64, p.w_row_o, &bar[BB_WO]);
        load_chunk(0);
        if (nch > 1) load_chunk(1);
        else load_wq();   // single chunk: stage 1 is never used by a chunk
    }
    for (int i = tid; i < 2 * LTILE / 16; i += CNT) reinterpret_cast<uint4*>(smem + B_OFF_QB)[i] = make_uint4(0u, 0u, 0u, 0u);
    const uint64_t dseed = DROP ? p.drop.resolve() : 0ull;

    // ---- g1 = dy * mask (thread = feature n) -> B operand [t][n] -----------------------------------------------------------
    float dyr[TP];
#pragma unroll
    for (int t = 0; t < TP; ++t) {
        const long long e = (rowb + t) * 128 + tid;
        dyr[t] = t < T ? p.dy[e] : 0.f;
        float g = dyr[t];
        if (DROP && t < T) g *= dropout_scale(dseed, p.drop.stream + 1, (uint64_t)e, p.drop.thresh, p.drop.inv_keep);
        const unsigned short u = bf16_bits(g);
        *reinterpret_cast<unsigned short*>(smem + B_OFF_XB + elem_off(t, tid, XB_TILE)) = u;
        if (t < T && p.g1) reinterpret_cast<unsigned short*>(p.g1)[e] = u;
    }
    if (tid < NQ) {
        const int h = tid >> 4, t = tid & 15;
        s_lse[tid] = t < T ? p.lse[((long long)b * CA_H + h) * Tf + t0 + t] : 0.f;
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    // ---- dattn^T[c][t] = sum_n Wout[n][c] g1[t][n] : A = Wout used MN-major ------------------------------------------------
    if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        mbar_wait(&bar[BB_WO], 0);
        const uint32_t idesc = instr_desc_bf16(128, 16, 1, 0);
        for (int k = 0; k < 8; ++k)
            mma_bf16_ss(tmem + DAC, smem_desc_mn_sw128(sbase + B_OFF_PD, LTILE, 1024) + 128 * (uint64_t)k,
                        smem_desc_k_sw128(sbase + B_OFF_XB + (k >> 2) * XB_TILE) + 2 * (k & 3), idesc, k > 0);
        mma_commit(&bar[BB_DA]);
    }
    __syncwarp();
    {
        float da[16];
        const int h = tid >> 5;
        // saved forward values of this thread's column (issued ahead of the wait)
        unsigned short at[TP], qv[TP];
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const long long e = (rowb + t) * 128 + tid;
            at[t] = t < T ? reinterpret_cast<const unsigned short*>(p.attn)[e] : (unsigned short)0;
            qv[t] = t < T ? reinterpret_cast<const unsigned short*>(p.q)[e] : (unsigned short)0;
        }
        mbar_wait(&bar[BB_DA], 0);
        tc_fence_after_sync();
        ld_lane16(tmem, warp, DAC, da);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            // delta[(h,t)] = sum over the head's 32 features (= this warp) of dattn * attn
            const float d = warp_sum(da[t] * bf16_to_f(at[t]));
            if (lane == 0) s_delta[TP * h + t] = d;
            *reinterpret_cast<unsigned short*>(smem + B_OFF_DOB + elem_off(h * TP + t, tid, QB_TILE)) = t < T ? bf16_bits(da[t]) : (unsigned short)0;
            *reinterpret_cast<unsigned short*>(smem + B_OFF_QB + elem_off(h * TP + t, tid, QB_TILE)) = qv[t];
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();

    const float sc = rsqrtf((float)CA_DH) * 1.4426950408889634f;
    const float scale = rsqrtf((float)CA_DH);
    const uint32_t id64 = instr_desc_bf16(128, NQ);
#pragma unroll 1
    for (int j = 0; j < nch; ++j) {
        const uint32_t kv = sbase + B_OFF_KV + (j & 1) * 4 * LTILE;
        if (warp == 0 && elect_one()) {
            tc_fence_after_sync();
            mbar_wait(&bar[BB_KV0 + (j & 1)], (uint32_t)((j >> 1) & 1));
            mma_k_tiles(tmem + SC, kv, LTILE, sbase + B_OFF_QB, QB_TILE, id64, 2, false);                  // S^T
            mma_k_tiles(tmem + DPC, kv + 2 * LTILE, LTILE, sbase + B_OFF_DOB, QB_TILE, id64, 2, false);    // dP^T = V dOblk^T
            mma_commit(&bar[BB_SD]);
        }
        __syncwarp();
        mbar_wait(&bar[BB_SD], (uint32_t)(j & 1));
        tc_fence_after_sync();
        const int m = 128 * j + tid;
        const bool kvld = m < M;
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
            float s[32], dp[32];
            ld_lane32(tmem, warp, SC + 32 * g, s);
            ld_lane32(tmem, warp, DPC + 32 * g, dp);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int c = 32 * g + i, t = i & 15, h = c >> 4;
                float pd = 0.f, ds = 0.f;
                if (kvld && t < T) {
                    const float pr = ex2_approx(fmaf(s[i], sc, -s_lse[c]));
                    float dm = 1.0f;
                    if (DROP)
                        dm = dropout_scale(dseed, p.drop.stream, (((uint64_t)b * CA_H + h) * Tf + t0 + t) * (uint64_t)M + m, p.drop.thresh,
                                           p.drop.inv_keep);
                    pd = pr * dm;
                    ds = pr * (dp[i] * dm - s_delta[c]) * scale;
                }
                s[i] = pd;
                dp[i] = ds;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                *reinterpret_cast<uint4*>(smem + B_OFF_PD + sw128_chunk_off(tid, 4 * g + c)) = pack8_bf16(s + 8 * c);
                *reinterpret_cast<uint4*>(smem + B_OFF_DS + sw128_chunk_off(tid, 4 * g + c)) = pack8_bf16(dp + 8 * c);
            }
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (warp == 0 && elect_one()) {
            tc_fence_after_sync();
            const uint32_t id_kn = instr_desc_bf16(128, 128, 0, 1);
            // dV[m][c] = sum_(h,t) Pd^T[m][(h,t)] dOblk[(h,t)][c] ; dK[m][c] = sum dS^T[m][(h,t)] Qblk[(h,t)][c]
            mma_a_k_b_mn(tmem + DVC, sbase + B_OFF_PD, 0, sbase + B_OFF_DOB, QB_TILE, id_kn, 4, false);
            mma_a_k_b_mn(tmem + DKC, sbase + B_OFF_DS, 0, sbase + B_OFF_QB, QB_TILE, id_kn, 4, false);
            // dQ^T[c][(h,t)] += sum_m K[m][c] dS^T[m][(h,t)]
            const uint32_t id_mm = instr_desc_bf16(128, NQ, 1, 1);
            const int ks = min(8, (M - 128 * j + 15) >> 4);
            for (int k = 0; k < ks; ++k)
                mma_bf16_ss(tmem + DQC, smem_desc_mn_sw128(kv, LTILE, 1024) + 128 * (uint64_t)k,
                            smem_desc_mn_sw128(sbase + B_OFF_DS, 1024, 1024) + 128 * (uint64_t)k, id_mm, (j > 0 || k > 0) ? 1u : 0u);
            mma_commit(&bar[BB_G]);
        }
        __syncwarp();
        mbar_wait(&bar[BB_G], (uint32_t)(j & 1));
        tc_fence_after_sync();
        if (warp == 0 && elect_one()) {   // stage j & 1 is free: it takes the next-but-one chunk, or Wq once no chunk needs it any more
            if (j + 2 < nch) load_chunk(j + 2);
            else if (j == nch - 2) load_wq();
        }
        __syncwarp();
        // dK | dV rows of this key -> bf16 all-layer gradient matrix
        {
            __nv_bfloat16* drow = p.dkv + (krow0 + m) * p.lddkv + p.kv_col0;
#pragma unroll 1
            for (int c0 = 0; c0 < 256; c0 += 32) {
                float v[32];
                ld_lane32(tmem, warp, (c0 < 128 ? DKC + c0 : DVC + c0 - 128), v);
                if (kvld) {
                    uint4* gp = reinterpret_cast<uint4*>(drow + c0);
                    if (p.dkv_accumulate) {   // a later query-row group of the same sample: add to what the earlier ones stored
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            float o[8];
                            unpack8_bf16(gp[i], o);
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[8 * i + e] += o[e];
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) gp[i] = pack8_bf16(v + 8 * i);
                }
            }
        }
        tc_fence_before_sync();
        __syncthreads();   // accumulators and Pd / dS tiles are free for the next chunk
        tc_fence_after_sync();
    }
    // ---- dq[t][c] (thread = feature c) -> global + B operand ---------------------------------------------------------------
    {
        float dq[16];
        const int h = tid >> 5;
        ld_lane16(tmem, warp, DQC + TP * h, dq);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const unsigned short u = t < T ? bf16_bits(dq[t]) : (unsigned short)0;
            *reinterpret_cast<unsigned short*>(smem + B_OFF_XB + elem_off(t, tid, XB_TILE)) = u;
            if (t < T && p.dq) reinterpret_cast<unsigned short*>(p.dq)[(rowb + t) * 128 + tid] = u;
        }
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    // ---- dxn^T[k][t] = sum_c Wq[c][k] dq[t][c] : A = Wq used MN-major ---------------------------------------------------------
    if (warp == 0 && elect_one()) {
        tc_fence_after_sync();
        mbar_wait(&bar[BB_WQ], 0);
        const uint32_t idesc = instr_desc_bf16(128, 16, 1, 0);
        for (int k = 0; k < 8; ++k)
            mma_bf16_ss(tmem + DXC, smem_desc_mn_sw128(wq_addr, LTILE, 1024) + 128 * (uint64_t)k,
                        smem_desc_k_sw128(sbase + B_OFF_XB + (k >> 2) * XB_TILE) + 2 * (k & 3), idesc, k > 0);
        mma_commit(&bar[BB_X]);
    }
    __syncwarp();
    // ---- LayerNorm backward (thread = feature k), dx = dy + LN'(dxn) -----------------------------------------------------------
    {
        float xh[TP], v[32];
        const float ga = __ldg(p.n_w + tid);
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            const long long grow = rowb + t;
            xh[t] = t < T ? (p.x[grow * 128 + tid] - p.stats[2 * grow]) * p.stats[2 * grow + 1] : 0.f;
        }
        float dxn[16];
        mbar_wait(&bar[BB_X], 0);
        tc_fence_after_sync();
        ld_lane16(tmem, warp, DXC, dxn);
        float gw = 0.f, gb = 0.f;
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            if (t >= T) dxn[t] = 0.f;
            gw = fmaf(dxn[t], xh[t], gw);
            gb += dxn[t];
            const float g = dxn[t] * ga;
            v[t] = g;             // sum_k g
            v[16 + t] = g * xh[t];   // sum_k g xhat
            dxn[t] = g;
        }
        const float r = warp_reduce32_cols<false>(v, lane);
        red[warp][lane] = r;
        __syncthreads();
        if (tid < 32) fin[tid] = ((red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid])) * (1.0f / 128.0f);
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TP; ++t) {
            if (t < T) {
                const long long grow = rowb + t;
                const float rstd = p.stats[2 * grow + 1];
                p.dx[grow * 128 + tid] = dyr[t] + rstd * (dxn[t] - fin[t] - xh[t] * fin[16 + t]);
            }
        }
        if (p.g_n_w) atomicAdd(p.g_n_w + tid, gw);
        if (p.g_n_b) atomicAdd(p.g_n_b + tid, gb);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

inline bool al16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace

extern "C" int sd_cast_bf16(const float* src, void* dst, long long n, void* stream) {
    if (n <= 0) return SD_OK;
    if (!src || !dst || n % 8 != 0 || !al16(src) || !al16(dst)) return SD_ERR_BAD_ARG;
    const long long n8 = n / 8;
    const int blocks = (int)std::min<long long>((n8 + 255) / 256, 148 * 16);
    cast_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, reinterpret_cast<uint4*>(dst), n8);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_bcast_row_bf16(void* dst, long long ld, long long block_rows, long long row, int B, const void* src_row,
                                 int ncols, void* stream) {
    if (B <= 0 || ncols <= 0) return SD_OK;
    if (!dst || !src_row || ld % 8 != 0 || ncols % 8 != 0 || ncols > ld || row < 0 || row >= block_rows || !al16(dst) || !al16(src_row))
        return SD_ERR_BAD_ARG;
    SD_CUDA(launch_chain(bcast_row_bf16_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, reinterpret_cast<uint4*>(dst), ld / 8,
                         block_rows, row, B, reinterpret_cast<const uint4*>(src_row), ncols / 8));
    return SD_OK;
}

extern "C" int sd_ddim_glue(const float* h, const float* fc_w, const float* fc_b, const float* x, float* x_next, float* eps_out,
                            int rows, int J, float sqrt_beta_t, float sqrt_alpha_t, float sqrt_alpha_prev, float sqrt_beta_prev,
                            const float* emb_w, const float* emb_b, const float* pe, int T, float* h_next, void* kv, long long ldkv,
                            long long block_rows, long long row, int B, const void* src_row, int ncols, void* stream) {
    if (rows <= 0) return SD_OK;
    if (!h || !fc_w || !fc_b || !x || !x_next || J < 1 || J > GL_JMAX) return SD_ERR_BAD_ARG;
    if (emb_w && (!emb_b || !pe || !h_next || T < 1)) return SD_ERR_BAD_ARG;
    if (src_row && (!kv || ldkv % 8 != 0 || ncols % 8 != 0 || ncols > ldkv || row < 0 || row >= block_rows || B < 1 || !al16(kv) ||
                    !al16(src_row)))
        return SD_ERR_BAD_ARG;
    GlueParams p{};
    p.h = h; p.fc_w = fc_w; p.fc_b = fc_b; p.x = x; p.x_next = x_next; p.eps_out = eps_out;
    p.sb = sqrt_beta_t; p.sa = sqrt_alpha_t; p.sap = sqrt_alpha_prev; p.sbp = sqrt_beta_prev;
    p.emb_w = emb_w; p.emb_b = emb_b; p.pe = pe; p.h_next = h_next;
    p.rows = rows; p.J = J; p.T = T; p.row_blocks = ceil_div(rows, GL_ROWS);
    p.kv = reinterpret_cast<uint4*>(kv); p.ld8 = ldkv / 8; p.block_rows = block_rows; p.row = row; p.B = src_row ? B : 0;
    p.src = reinterpret_cast<const uint4*>(src_row); p.n8 = ncols / 8;
    SD_CUDA(launch_chain(ddim_glue_kernel, dim3(p.row_blocks + p.B), dim3(GL_NT), 0, (cudaStream_t)stream, p));
    return SD_OK;
}

// forward with saves + backward: as the forward, T <= 64 in query-row groups of <= 16 (the backward runs one launch per group)
extern "C" int sd_ca_block_supported(int d, int H, int T, int M) {
    if (d != 128 || H != CA_H || T < 1 || T > 4 * TP || M < 1 || M > 128 * MAXCH) return 0;
    return tensor_map_encoder() != nullptr ? 1 : 0;
}

// inference (no saves): any T <= 64 — the query rows are split into groups of <= 16, one CTA each
extern "C" int sd_ca_block_fwd_supported(int d, int H, int T, int M) {
    if (d != 128 || H != CA_H || T < 1 || T > 4 * TP || M < 1 || M > 128 * MAXCH) return 0;
    return tensor_map_encoder() != nullptr ? 1 : 0;
}

extern "C" int sd_kv_proj_bf16(const void* mem_bf16, long long rows, const void* w_packed, int w_rows_total, int w_row0,
                               int w_stride, int n_layers, const float* const* biases, void* kv_out, long long ldkv,
                               void* stream) {
    if (rows <= 0) return SD_OK;
    if (!mem_bf16 || !w_packed || !kv_out || n_layers < 1 || n_layers > SD_KV_MAX_LAYERS) return SD_ERR_BAD_ARG;
    if (ldkv < 256LL * n_layers || ldkv % 8 != 0 || !al16(kv_out)) return SD_ERR_BAD_ARG;
    if (w_row0 < 0 || w_row0 + (long long)(n_layers - 1) * w_stride + 256 > w_rows_total) return SD_ERR_BAD_ARG;
    CUtensorMap tmA, tmW, tmC;
    if (!encode_bf16_2d(&tmA, mem_bf16, rows, 128, 128, 128)) return SD_ERR_UNSUPPORTED;
    if (!encode_bf16_2d(&tmW, w_packed, w_rows_total, 128, 128, 128)) return SD_ERR_UNSUPPORTED;
    if (!encode_bf16_2d(&tmC, kv_out, rows, 256LL * n_layers, ldkv, 128)) return SD_ERR_UNSUPPORTED;   // stores clip at `rows`
    KvProjParams p{};
    p.rows = rows; p.nblk = n_layers; p.w_row0 = w_row0; p.w_stride = w_stride;
    for (int i = 0; i < n_layers; ++i) p.bias[i] = biases ? biases[i] : nullptr;
    p.C = reinterpret_cast<__nv_bfloat16*>(kv_out); p.ldc = ldkv;
    static bool configured = false;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(kv_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, KP_SMEM));
        configured = true;
    }
    // one weight block per CTA, the SMs split evenly over the blocks
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int groups = std::max(1, std::min(ceil_div(rows, 128), sms / n_layers));
    kv_proj_kernel<<<groups * n_layers, KP_NT, KP_SMEM, (cudaStream_t)stream>>>(tmA, tmW, tmC, p);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

static int launch_mn_gemm(const CUtensorMap& tmA, const CUtensorMap& tmW, const MnGemmParams& p, cudaStream_t st) {
    const int smem = KD_STAGES * kd_stage_bytes(p.N) + 1024;
    static int configured = 0;
    if (configured < smem) {
        SD_CUDA(cudaFuncSetAttribute(mn_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    mn_gemm_kernel<<<ceil_div(p.rows, 128), CNT, smem, st>>>(tmA, tmW, p);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_kv_dgrad_bf16(const void* dkv_bf16, long long rows, long long lddkv, const void* w_packed, int w_rows_total,
                                int w_row0, int w_stride, int n_layers, float* dmem, long long lddmem, int accumulate,
                                void* stream) {
    if (rows <= 0) return SD_OK;
    if (!dkv_bf16 || !w_packed || !dmem || n_layers < 1 || n_layers > SD_KV_MAX_LAYERS) return SD_ERR_BAD_ARG;
    if (lddkv < 256LL * n_layers || lddmem < 128 || lddmem % 4 != 0 || !al16(dmem)) return SD_ERR_BAD_ARG;
    if (w_row0 < 0 || w_row0 + (long long)(n_layers - 1) * w_stride + 256 > w_rows_total) return SD_ERR_BAD_ARG;
    CUtensorMap tmA, tmW;
    if (!encode_bf16_2d(&tmA, dkv_bf16, rows, 256LL * n_layers, lddkv, 128)) return SD_ERR_UNSUPPORTED;
    if (!encode_bf16_2d(&tmW, w_packed, w_rows_total, 128, 128, 64)) return SD_ERR_UNSUPPORTED;
    MnGemmParams p{};
    p.rows = rows; p.ktiles = 4 * n_layers; p.N = 128; p.w_row0 = w_row0; p.w_stride = w_stride; p.tiles_per_blk = 4;
    p.Cf = dmem; p.ldc = lddmem; p.accumulate = accumulate;
    return launch_mn_gemm(tmA, tmW, p, (cudaStream_t)stream);
}

extern "C" int sd_conv1x1s2_dgrad_supported(int Hin, int Win, int Cin, int Cout) {
    if (Hin < 2 || Win < 2 || (Hin & 1) || (Win & 1) || (Cin != 64 && Cin != 128 && Cin != 256) || Cout < 64 || Cout % 64 != 0) return 0;
    return tensor_map_encoder() != nullptr ? 1 : 0;
}

extern "C" int sd_conv1x1s2_dgrad_bf16(const void* dy, const void* w_bf16, void* dx, int frames, int Hin, int Win, int Cin, int Cout,
                                       void* stream) {
    if (frames <= 0) return SD_OK;
    if (!dy || !w_bf16 || !dx) return SD_ERR_BAD_ARG;
    if (!sd_conv1x1s2_dgrad_supported(Hin, Win, Cin, Cout)) return SD_ERR_UNSUPPORTED;
    const int Ho = Hin / 2, Wo = Win / 2;
    const long long rows = (long long)frames * Ho * Wo;
    CUtensorMap tmA, tmW;
    if (!encode_bf16_2d(&tmA, dy, rows, Cout, Cout, 128)) return SD_ERR_UNSUPPORTED;
    if (!encode_bf16_2d(&tmW, w_bf16, Cout, Cin, Cin, 64)) return SD_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    SD_CUDA(cudaMemsetAsync(dx, 0, (size_t)frames * Hin * Win * Cin * 2, st));   // the three pixels of each 2x2 block the stride skips
    MnGemmParams p{};
    p.rows = rows; p.ktiles = Cout / 64; p.N = Cin; p.w_row0 = 0; p.w_stride = 0; p.tiles_per_blk = p.ktiles;
    p.Cb = reinterpret_cast<__nv_bfloat16*>(dx); p.Ho = Ho; p.Wo = Wo; p.Hin = Hin; p.Win = Win;
    return launch_mn_gemm(tmA, tmW, p, st);
}

namespace {
template <bool DROP>
int launch_ca_fwd(const CUtensorMap& tmW, const CUtensorMap& tmKV, const CaFwdParams& p, cudaStream_t st) {
    auto kernel = ca_fwd_kernel<DROP>;
    static bool configured = false;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
        SD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));   // two CTAs per SM
        configured = true;
    }
    SD_CUDA(launch_chain(kernel, dim3(p.B * p.groups), dim3(F_NT), F_SMEM, st, tmW, tmKV, p));
    return SD_OK;
}
template <bool DROP>
int launch_ca_bwd(const CUtensorMap& tmW, const CUtensorMap& tmKV, const CaBwdParams& p, cudaStream_t st) {
    auto kernel = ca_bwd_kernel<DROP>;
    static bool configured = false;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
        configured = true;
    }
    kernel<<<p.B, CNT, B_SMEM, st>>>(tmW, tmKV, p);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
}  // namespace

extern "C" int sd_ca_block_fwd(const sd_ca_block_desc* d, void* stream) {
    if (!d || !d->x || !d->y || !d->w_packed || !d->kv || !d->q_b || !d->out_b || !d->n_w || !d->n_b) return SD_ERR_BAD_ARG;
    if (d->B <= 0) return SD_OK;
    if (!sd_ca_block_fwd_supported(128, 4, d->T, d->M)) return SD_ERR_UNSUPPORTED;
    if (d->w_row_q < 0 || d->w_row_q + 128 > d->w_rows_total || d->w_row_o < 0 || d->w_row_o + 128 > d->w_rows_total) return SD_ERR_BAD_ARG;
    if (d->kv_col0 < 0 || d->kv_col0 + 256 > d->ldkv || d->kv_col0 % 8 != 0) return SD_ERR_BAD_ARG;
    if (!al16(d->x) || !al16(d->y) || !al16(d->xn_save) || !al16(d->q_save) || !al16(d->attn_save)) return SD_ERR_BAD_ARG;
    CUtensorMap tmW, tmKV;
    if (!encode_bf16_2d(&tmW, d->w_packed, d->w_rows_total, 128, 128, 128)) return SD_ERR_UNSUPPORTED;
    if (!encode_bf16_2d(&tmKV, d->kv, (long long)d->B * d->M, d->ldkv, d->ldkv, 128)) return SD_ERR_UNSUPPORTED;
    CaFwdParams p{};
    p.x = d->x; p.y = d->y; p.B = d->B; p.T = d->T; p.M = d->M;
    p.groups = (d->T + TP - 1) / TP;
    p.Tg = (d->T + p.groups - 1) / p.groups;
    p.w_row_q = d->w_row_q; p.w_row_o = d->w_row_o; p.kv_col0 = d->kv_col0;
    p.q_b = d->q_b; p.out_b = d->out_b; p.n_w = d->n_w; p.n_b = d->n_b;
    p.xn_save = (__nv_bfloat16*)d->xn_save; p.q_save = (__nv_bfloat16*)d->q_save; p.attn_save = (__nv_bfloat16*)d->attn_save;
    p.stats_save = d->stats_save; p.lse_save = d->lse_save;
    p.drop = make_dropout(d->dropout_p, d->dropout_seed, d->dropout_stream);
    return p.drop.thresh != 0 ? launch_ca_fwd<true>(tmW, tmKV, p, (cudaStream_t)stream)
                              : launch_ca_fwd<false>(tmW, tmKV, p, (cudaStream_t)stream);
}

extern "C" int sd_ca_block_bwd(const sd_ca_block_bwd_desc* d, void* stream) {
    if (!d || !d->dy || !d->dx || !d->x || !d->q || !d->attn || !d->stats || !d->lse || !d->w_packed || !d->kv || !d->dkv ||
        !d->n_w)
        return SD_ERR_BAD_ARG;
    if (d->B <= 0) return SD_OK;
    if (!sd_ca_block_supported(128, 4, d->T, d->M)) return SD_ERR_UNSUPPORTED;
    if (d->w_row_q < 0 || d->w_row_q + 128 > d->w_rows_total || d->w_row_o < 0 || d->w_row_o + 128 > d->w_rows_total) return SD_ERR_BAD_ARG;
    if (d->kv_col0 < 0 || d->kv_col0 + 256 > d->ldkv || d->kv_col0 + 256 > d->lddkv || d->kv_col0 % 8 != 0 || d->lddkv % 8 != 0)
        return SD_ERR_BAD_ARG;
    if (!al16(d->dkv)) return SD_ERR_BAD_ARG;
    CUtensorMap tmW, tmKV;
    if (!encode_bf16_2d(&tmW, d->w_packed, d->w_rows_total, 128, 128, 128)) return SD_ERR_UNSUPPORTED;
    if (!encode_bf16_2d(&tmKV, d->kv, (long long)d->B * d->M, d->ldkv, d->ldkv, 128)) return SD_ERR_UNSUPPORTED;
    CaBwdParams p{};
    p.dy = d->dy; p.dx = d->dx; p.x = d->x; p.q = (const __nv_bfloat16*)d->q; p.attn = (const __nv_bfloat16*)d->attn;
    p.stats = d->stats; p.lse = d->lse; p.B = d->B; p.T = d->T; p.M = d->M;
    p.w_row_q = d->w_row_q; p.w_row_o = d->w_row_o; p.kv_col0 = d->kv_col0; p.n_w = d->n_w;
    p.g1 = (__nv_bfloat16*)d->g1; p.dq = (__nv_bfloat16*)d->dq; p.dkv = (__nv_bfloat16*)d->dkv; p.lddkv = d->lddkv;
    p.g_n_w = d->g_n_w; p.g_n_b = d->g_n_b;
    p.drop = make_dropout(d->dropout_p, d->dropout_seed, d->dropout_stream);
    const int groups = (d->T + TP - 1) / TP;
    p.Tg = (d->T + groups - 1) / groups;
    for (int g = 0; g < groups; ++g) {
        p.t0 = g * p.Tg;
        p.dkv_accumulate = g > 0;
        const int rc = p.drop.thresh != 0 ? launch_ca_bwd<true>(tmW, tmKV, p, (cudaStream_t)stream)
                                          : launch_ca_bwd<false>(tmW, tmKV, p, (cudaStream_t)stream);
        if (rc != SD_OK) return rc;
    }
    return SD_OK;
}
