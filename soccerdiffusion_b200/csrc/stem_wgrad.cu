// Weight gradient of the ResNet stem convolution (conv1: 7x7, stride 2, pad 3, Cin=3 -> 64) in its space-to-depth
// form (4x4, stride 1 over the packed image [N][Hp][Wp][16], see sd_stem_pack_s2d_bf16) on tcgen05 tensor cores.
//
//   dW[j][c] = sum over output pixels p=(n,ho,wo) of  patch[p][j] * dy[p][c]
//   j = kh*64 + kw*16 + ci  (kh,kw in 0..3, ci in 0..15),  patch[p][kh*64 .. kh*64+63] = xs2d[n][ho+kh][wo .. wo+3][0..15]
//
// i.e. a (256 x 64) += (256 x P) * (P x 64) GEMM whose contraction runs over the 32 M output pixels of a batch.  Both
// operands are "MN-major" in memory (for one pixel the 64 patch values of a filter row are 128 contiguous bytes, and
// so are the 64 channels of dy), so they are copied with 16-byte cp.async straight into the 128-byte-swizzled
// MN-major operand layout — no transposition, no im2col buffer — through a 3-stage ring; two 128x64 fp32
// accumulators live in TMEM for the whole kernel and are reduced into the output with atomics at the end.
//
// cuDNN's wgrad for this layer (Cin=3) measured 7.3 ms per step at bs=256 (round-1 profile): 6 % of the step.
#include "common.cuh"
#include "tc_common.cuh"
#include "../../include/sd_b200.h"

#include <cuda.h>
#include <cstdlib>

using namespace sd;
using namespace sdtc;

namespace {

constexpr int NT = 256;
constexpr int KC = 64;                       // pixels per stage
constexpr int STAGES = 3;
constexpr int A_BYTES = 4 * KC * 128;        // 4 filter rows (kh) x 64 pixels x 128 B
constexpr int B_BYTES = KC * 128;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(NT, 1) stem_wgrad_kernel(const uint4* __restrict__ xs, const uint4* __restrict__ dy,
                                                           float* __restrict__ dw, int Nimg, int HO, int WO, int Hp, int Wp,
                                                           long long P, int chunks_per_cta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_stage[STAGES];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&bar_stage[s], 1);
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    const long long nchunks_total = (P + KC - 1) / KC;
    const long long c_begin = (long long)blockIdx.x * chunks_per_cta;
    const long long c_end = min(nchunks_total, c_begin + chunks_per_cta);
    const int nchunks = (int)max(0LL, c_end - c_begin);
    const uint32_t idesc = instr_desc_bf16(128, 64, 1, 1);

    // per-thread copy assignment: A: 8 items (c16 = tid&7, kh = (tid>>3)&3, pixel = (tid>>5) + 8*i); B: 2 items
    const int a_c16 = tid & 7, a_kh = (tid >> 3) & 3, a_px0 = tid >> 5;
    const int b_c16 = tid & 7, b_px0 = tid >> 3;
    const long long plane = (long long)HO * WO;

    auto issue = [&](int ci) {
        const int s = ci % STAGES;
        uint8_t* As = smem + s * STAGE_BYTES;
        uint8_t* Bs = As + A_BYTES;
        const long long p0 = (c_begin + ci) * KC;
        const int n0 = (int)(p0 / plane);
        const int rem = (int)(p0 - (long long)n0 * plane);
        const int ho0 = rem / WO, wo0 = rem - ho0 * WO;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int px = a_px0 + 8 * i;
            int wo = wo0 + px, ho = ho0, n = n0;
            while (wo >= WO) { wo -= WO; ++ho; }   // one iteration at most when WO >= KC (224x224 images: WO = 112)
            while (ho >= HO) { ho -= HO; ++n; }
            const uint32_t dst = smem_u32(As) + (uint32_t)((a_kh >> 1) * (2 * KC * 128)) +
                                 sw128_mn_chunk_off((a_kh & 1) * 64 + a_c16 * 8, px, KC);
            if (p0 + px < P) {
                const uint4* src = xs + ((((long long)n * Hp + ho + a_kh) * Wp + wo) * 2) + a_c16;
                cp_async16(dst, src);
            } else {
                asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int px = b_px0 + 32 * i;
            const uint32_t dst = smem_u32(Bs) + sw128_mn_chunk_off(b_c16 * 8, px, KC);
            if (p0 + px < P) cp_async16(dst, dy + (p0 + px) * 8 + b_c16);
            else asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
        }
        cp_async_commit();
    };

    if (nchunks > 0) issue(0);
    if (nchunks > 1) issue(1);
    for (int ci = 0; ci < nchunks; ++ci) {
        const int s = ci % STAGES;
        if (ci + 2 < nchunks) {
            const int s2 = (ci + 2) % STAGES;
            // stage s2 was read by the MMAs of chunk ci-1
            if (ci >= 1) mbar_wait(&bar_stage[s2], (uint32_t)(((ci - 1) / STAGES) & 1));
            issue(ci + 2);
            cp_async_wait<2>();
        } else if (ci + 1 < nchunks) {
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after_sync();
            const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES);
            const uint64_t db = smem_desc_mn_sw128(a0 + A_BYTES, KC * 128, 1024);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const uint64_t da = smem_desc_mn_sw128(a0 + mt * (2 * KC * 128), KC * 128, 1024);
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks)   // 16 k per MMA = 2 atoms of 8 k-rows = 2048 bytes
                    mma_bf16_ss(tmem + mt * 64, da + (uint64_t)(128 * ks), db + (uint64_t)(128 * ks), idesc,
                                (ci > 0 || ks > 0) ? 1u : 0u);
            }
            mma_commit(&bar_stage[s]);
            if (ci == nchunks - 1) mma_commit(&bar_done);
        }
    }
    if (nchunks > 0) {
        mbar_wait(&bar_done, 0);
        tc_fence_after_sync();
        // rows j = mt*128 + 32*(warp&3) + lane, 64 columns each
        const int q = warp & 3, mt = warp >> 2;
        const int j = mt * 128 + q * 32 + lane;
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
            float v[32];
            tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 64 + cb * 32), v);
#pragma unroll
            for (int c = 0; c < 32; ++c) atomicAdd(dw + j * 64 + cb * 32 + c, v[c]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- TMA-fed variant: one image row of output pixels (KC = WO, a multiple of 16, <= 128) per pipeline stage --------
// Operand tiles are fetched by the TMA engine from two tensor maps with the 128-byte swizzle the MN-major tcgen05
// descriptors expect:
//   tmA: the packed image viewed as [pixel-start q][64 elements] with a 32-byte row pitch (consecutive patch rows
//        overlap by 3 pixels): box {64, KC} at (0, (n*Hp + ho + kh)*Wp) = the kh-th filter row of KC patches;
//   tmB: dy viewed as [pixel][64]: box {64, KC} at (0, row*WO).
// Warp 0 = producer (one lane issues 5 TMA loads per stage), warp 1 = MMA issuer (one lane), full/empty mbarrier
// ring; everybody joins for the TMEM epilogue.
constexpr int TMA_STAGES = 3;

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(NT, 1) stem_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const __grid_constant__ CUtensorMap tmB, float* __restrict__ dw,
                                                               int HO, int WO, int Hp, int Wp, long long rows_total,
                                                               int rows_per_cta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[TMA_STAGES];
    __shared__ __align__(8) uint64_t bar_empty[TMA_STAGES];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KCr = WO;                       // pixels (k rows) per stage
    const int blk = KCr * 128;                // bytes of one 64-wide MN block (one filter row / dy)
    const int stage_bytes = 5 * blk;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TMA_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = min(rows_total, r_begin + rows_per_cta);
    const int nrows = (int)max(0LL, r_end - r_begin);

    if (warp == 0 && elect_one()) {
        for (int ci = 0; ci < nrows; ++ci) {
            const int s = ci % TMA_STAGES;
            mbar_wait(&bar_empty[s], (uint32_t)(((ci / TMA_STAGES) & 1) ^ 1));
            const long long r = r_begin + ci;
            const int n = (int)(r / HO), ho = (int)(r - (long long)n * HO);
            const uint32_t base = smem_u32(smem + s * stage_bytes);
            mbar_arrive_expect_tx(&bar_full[s], (uint32_t)stage_bytes);
#pragma unroll
            for (int kh = 0; kh < 4; ++kh)
                tma_load_2d(base + kh * blk, &tmA, 0, (n * Hp + ho + kh) * Wp, &bar_full[s]);
            tma_load_2d(base + 4 * blk, &tmB, 0, (int)(r * WO), &bar_full[s]);
        }
    } else if (warp == 1 && elect_one()) {
        // D[64 output channels][256 = (kh, kw, ci)] += dy^T (M = 64, MN-major A) x patches (N = 256: the four filter-row
        // tiles are the four 64-wide MN blocks of ONE MN-major B operand, LBO = blk).  One 64 x 256 x 16 instruction per 16
        // pixels instead of two 128 x 64 x 16: with N = 64 the tensor pipe ran at ~110 clk per instruction (25 % active).
        const uint32_t idesc = instr_desc_bf16(64, 256, 1, 1);
        const int ksteps = KCr / 16;
        for (int ci = 0; ci < nrows; ++ci) {
            const int s = ci % TMA_STAGES;
            mbar_wait(&bar_full[s], (uint32_t)((ci / TMA_STAGES) & 1));
            tc_fence_after_sync();
            const uint32_t base = smem_u32(smem + s * stage_bytes);
            const uint64_t da = smem_desc_mn_sw128(base + 4 * blk, (uint32_t)blk, 1024);   // dy tile: 64 channels wide
            const uint64_t db = smem_desc_mn_sw128(base, (uint32_t)blk, 1024);             // 4 x 64 patch elements
            for (int ks = 0; ks < ksteps; ++ks)
                mma_bf16_ss(tmem, da + (uint64_t)(128 * ks), db + (uint64_t)(128 * ks), idesc, (ci > 0 || ks > 0) ? 1u : 0u);
            mma_commit(&bar_empty[s]);
            if (ci == nrows - 1) mma_commit(&bar_done);
        }
    }
    __syncwarp();
    if (nrows > 0) {
        mbar_wait(&bar_done, 0);
        tc_fence_after_sync();
        // M = 64 accumulator layout: output channel c lives in TMEM lane (c % 16) + 32 * (c / 16); column = k index j
        const int q = warp & 3, half = warp >> 2;
        const int c = q * 16 + lane;
#pragma unroll
        for (int cb = 0; cb < 4; ++cb) {
            const int j0 = half * 128 + cb * 32;
            float v[32];
            tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)j0, v);
            if (lane < 16) {
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(dw + (j0 + j) * 64 + c, v[j]);
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---- paired variant: two patch tiles x two dy rows per instruction ---------------------------------------------------
// The patch tile T[t] of packed-image row t serves filter row kh of output row t - kh.  One 128 x 128 x 16 instruction per 16
// pixels covers all four filter rows:  A = (T[t], T[t+1]) (M = 128 = (a, j), the tiles adjacent in a ring),
// B = (dy[t-2], dy[t]) (N = 128 = (b, co), two ring slots apart):  D[(a,j)][(b,co)] += sum_pixels T[t+a][.][j] dy[r_b][.][co]
// = dW[kh = a + 2 (1 - b)][j][co], accumulated over t = 0 .. HO+1 of every image (dy rows outside the image arrive as zeros
// from the 4-D tensor map).  Against the 64 x 256 x 16 form: 94 instead of 128 clk per 16 pixels (an M = 64 instruction costs
// what an M = 128 one does) and 2.75 instead of 5 tile loads per output row (each tile is loaded once, plus the "shadow" copies
// that keep pairs contiguous across the ring's wrap-around).
constexpr int RT = 6, RD = 6;        // ring sizes (tiles in flight: the deeper the ring, the further the producer runs ahead)
constexpr int PT_SLOTS = RT + 1;     // patch-tile ring + shadow of slot 0 behind the last slot
constexpr int PD_SLOTS = RD + 2;     // dy ring + shadows of slots 0, 1 behind the last slot
constexpr int PSTG = 16;      // stage-completion barriers

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

__global__ void __launch_bounds__(NT, 1) stem_wgrad_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                const __grid_constant__ CUtensorMap tmB4, float* __restrict__ dw,
                                                                int Nimg, int HO, int WO, int Hp, int Wp) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_t[RT], full_d[RD], stage_done[PSTG], bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int blk = WO * 128;
    const uint32_t tbase = smem_u32(smem), dbase = tbase + PT_SLOTS * blk;
    if (tid == 0) {
        for (int i = 0; i < RT; ++i) mbar_init(&full_t[i], 1);
        for (int i = 0; i < RD; ++i) mbar_init(&full_d[i], 1);
        for (int i = 0; i < PSTG; ++i) mbar_init(&stage_done[i], 1);
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int n0 = (int)((long long)Nimg * blockIdx.x / gridDim.x), n1 = (int)((long long)Nimg * (blockIdx.x + 1) / gridDim.x);
    const int nst = HO + 2;            // stages per image
    const int ntl = HO + 3, ndl = HO + 4;   // patch tiles / dy tiles per image

    if (warp == 0 && elect_one()) {
        // ---- producer: tiles in the order the stages first need them; a slot is refilled once the last stage that read its
        //      previous tile has completed (stages complete in order: one running maximum of waited stages suffices)
        int last_t[RT], last_d[RD];   // global stage of the last read of each slot's tile
        for (int i = 0; i < RT; ++i) last_t[i] = -1;
        for (int i = 0; i < RD; ++i) last_d[i] = -1;
        int waited = -1;
        auto wait_stage = [&](int g) {
            if (g > waited) { mbar_wait(&stage_done[g % PSTG], (uint32_t)((g / PSTG) & 1)); waited = g; }
        };
        int yt = 0, wd = 0;            // global tile counters
        for (int n = n0; n < n1; ++n) {
            const int gbase = (n - n0) * nst;
            int it = 0, id = 0;        // next patch / dy tile of this image
            for (int t = 0; t < nst; ++t) {
                while (it <= t + RT - 3 && it < ntl) {  // stage t reads T[t], T[t+1]; run ahead as far as the ring allows
                    const int s = yt % RT;
                    wait_stage(last_t[s]);
                    const uint32_t bytes = (uint32_t)blk * (s == 0 ? 2u : 1u);
                    mbar_arrive_expect_tx(&full_t[s], bytes);
                    tma_load_2d(tbase + s * blk, &tmA, 0, (n * Hp + it) * Wp, &full_t[s]);
                    if (s == 0) tma_load_2d(tbase + RT * blk, &tmA, 0, (n * Hp + it) * Wp, &full_t[s]);
                    last_t[s] = gbase + min(it, nst - 1);
                    ++it; ++yt;
                }
                while (id <= t + RD - 2 && id < ndl) {  // stage t reads dy tiles t (row t-2) and t+2 (row t)
                    const int s = wd % RD;
                    wait_stage(last_d[s]);
                    const uint32_t bytes = (uint32_t)blk * (s < 2 ? 2u : 1u);
                    mbar_arrive_expect_tx(&full_d[s], bytes);
                    tma_load_4d(dbase + s * blk, &tmB4, 0, 0, id - 2, n, &full_d[s]);
                    if (s < 2) tma_load_4d(dbase + (s + RD) * blk, &tmB4, 0, 0, id - 2, n, &full_d[s]);
                    last_d[s] = gbase + min(id, nst - 1);
                    ++id; ++wd;
                }
            }
        }
    } else if (warp == 1 && elect_one()) {
        const uint32_t idesc = instr_desc_bf16(128, 128, 1, 1);
        const int ksteps = WO / 16;
        int wt = 0, wdd = 0;           // tiles whose full barrier has been waited for
        int g = 0;
        for (int n = n0; n < n1; ++n) {
            const int ybase = (n - n0) * ntl, zbase = (n - n0) * ndl;
            for (int t = 0; t < nst; ++t, ++g) {
                const int x = ybase + t, z = zbase + t;
                while (wt <= x + 1) { mbar_wait(&full_t[wt % RT], (uint32_t)((wt / RT) & 1)); ++wt; }
                while (wdd <= z + 2) { mbar_wait(&full_d[wdd % RD], (uint32_t)((wdd / RD) & 1)); ++wdd; }
                tc_fence_after_sync();
                const uint64_t da = smem_desc_mn_sw128(tbase + (x % RT) * blk, (uint32_t)blk, 1024);        // T[t], T[t+1]
                const uint64_t db = smem_desc_mn_sw128(dbase + (z % RD) * blk, (uint32_t)(2 * blk), 1024);  // dy[t-2], dy[t]
                for (int ks = 0; ks < ksteps; ++ks)
                    mma_bf16_ss(tmem, da + (uint64_t)(128 * ks), db + (uint64_t)(128 * ks), idesc, (g > 0 || ks > 0) ? 1u : 0u);
                mma_commit(&stage_done[g % PSTG]);
            }
        }
        mma_commit(&bar_done);
    }
    __syncwarp();
    if (n1 > n0 && warp < 4) {
        mbar_wait(&bar_done, 0);
        tc_fence_after_sync();
        const int a = warp >> 1, j = (warp & 1) * 32 + lane;   // accumulator row = (a, j)
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            float v[32];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            const int b = c0 >> 6, kh = a + 2 * (1 - b);
            float* dst = dw + (kh * 64 + j) * 64 + (c0 & 63);
#pragma unroll
            for (int c = 0; c < 32; ++c) atomicAdd(dst + c, v[c]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ---- conv1 forward (space-to-depth form) on the same TMA tensor map: y[p][c] = sum_j patch[p][j] * W[c][j] ------------
// The patch tile of packed-image row t ([WO pixels][64], the very boxes the weight-gradient kernel loads, read here with
// K-major descriptors) is the operand of output row t - kh for every filter row kh.  A tcgen05.mma of 128 x 64 x 16 costs
// almost as much as one of 128 x 128 x 16 (55-68 vs 64-67 clk, tools/probe_shifted_mma.py), so TWO filter rows share each
// instruction: N = 128 = [W_kh ; W_kh+1] (the weight tiles are adjacent in shared memory).  Per input tile t:
//     G1:  S(t)   = P_t [W0;W1]^T     columns 0..63 -> output row t (kh = 0), columns 64..127 -> output row t-1 (kh = 1)
//     G2:  S(t-2) += P_t [W2;W3]^T    columns 0..63 -> output row t-2 (kh = 2), columns 64..127 -> output row t-3 (kh = 3)
// i.e. accumulator slot S(t) (128 TMEM columns) ends up holding the kh in {0,2} half of output row t in its low columns
// and the kh in {1,3} half of output row t-1 in its high columns:  y[r] = S(r).lo + S(r+1).hi, complete after tile r+3.
// 8 instructions per output row instead of 16, every tile is read during one step only (the ring is a prefetch queue).
// Warp 0: TMA producer, warp 1: MMA issuer, warps 4-7: epilogue (thread = pixel: adds the two halves in registers,
// BatchNorm statistics, bf16, one 128-byte NHWC row per thread through a swizzled staging block).
constexpr int RING = 8;   // input-row tiles in flight
constexpr int NACC = 4;   // accumulator slots of 128 TMEM columns
constexpr bool kDirectStore = false;  // A/B: epilogue stores straight from registers (true) or through the swizzled staging block
constexpr int FNT = 384;  // warp 0 producer, warp 1 MMA issuer, warps 4-7 / 8-11: epilogue of channels 0-31 / 32-63

// the CTA's output rows [r_begin, r_end) as windows of consecutive rows of one image
struct RowWindow { int n, ho_a, ho_b; };
__device__ __forceinline__ bool next_window(long long& r, long long r_end, int HO, RowWindow& w) {
    if (r >= r_end) return false;
    w.n = (int)(r / HO);
    w.ho_a = (int)(r - (long long)w.n * HO);
    w.ho_b = (int)min((long long)HO - 1, w.ho_a + (r_end - r) - 1);
    r += w.ho_b - w.ho_a + 1;
    return true;
}

__global__ void __launch_bounds__(FNT, 1) stem_fprop_tma_kernel(const __grid_constant__ CUtensorMap tmA,
                                                               const uint4* __restrict__ w_s2d, uint4* __restrict__ y,
                                                               double* __restrict__ sums, int HO, int WO, int Hp, int Wp,
                                                               long long rows_total, int rows_per_cta) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[RING];
    __shared__ __align__(8) uint64_t bar_empty[RING];
    __shared__ __align__(8) uint64_t row_full[NACC];
    __shared__ __align__(8) uint64_t acc_empty[NACC];
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int blk = WO * 128;                 // one tile: WO patches x 64 elements (one filter row)
    uint8_t* Ws = smem + RING * blk;          // 4 tiles [64 channels][64 k] K-major, 8 KB each (+ slack for M=128 reads)
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < RING; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
#pragma unroll
        for (int a = 0; a < NACC; ++a) { mbar_init(&row_full[a], 1); mbar_init(&acc_empty[a], 1); }
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, NACC * 128);
    // weights: w_s2d[c][j] bf16 (j = kh*64 + kw*16 + ci) -> tile kh, row c, 16-byte chunk (j % 64) / 8
    for (int i = tid; i < 64 * 32; i += FNT) {
        const int c = i >> 5, ch = i & 31;    // 32 chunks of 8 bf16 per output channel
        *reinterpret_cast<uint4*>(Ws + (ch >> 3) * 8192 + sw128_chunk_off(c, ch & 7)) = w_s2d[i];
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    const long long r_begin = (long long)blockIdx.x * rows_per_cta;
    const long long r_end = min(rows_total, r_begin + rows_per_cta);
    const int nrows = (int)max(0LL, r_end - r_begin);

    if (warp == 0 && elect_one()) {
        int g = 0;   // tile sequence number
        long long r = r_begin;
        RowWindow w;
        while (next_window(r, r_end, HO, w))
            for (int t = w.ho_a; t <= w.ho_b + 3; ++t, ++g) {
                const int s = g % RING;
                mbar_wait(&bar_empty[s], (uint32_t)(((g / RING) & 1) ^ 1));
                mbar_arrive_expect_tx(&bar_full[s], (uint32_t)blk);
                tma_load_2d(smem_u32(smem + s * blk), &tmA, 0, (w.n * Hp + t) * Wp, &bar_full[s]);
            }
    } else if (warp == 1 && elect_one()) {
        const uint32_t idesc = instr_desc_bf16(128, 128, 0, 0);
        const uint64_t db01 = smem_desc_k_sw128(smem_u32(Ws)), db23 = smem_desc_k_sw128(smem_u32(Ws) + 2 * 8192);
        int gbase = 0, gready = -1, u = 0, v = 0;   // first tile of the window / last tile waited for / accumulator-slot / output-row sequence numbers
        long long r = r_begin;
        RowWindow w;
        // This thread issues in program order, so a wait for a free accumulator slot also delays everything behind it: G1(t)
        // (which overwrites slot S(t) = S(t-4)) waits for the epilogue of output row t-4, i.e. for G2(t-1) to complete plus the
        // hand-off (commit -> epilogue wake-up -> tcgen05.ld -> barrier -> arrive -> wake-up here, ~450 + 300 clk).  G2 of the
        // NEXT tile is therefore issued before that wait: the MMA a wait depends on was issued two iterations earlier.
        while (next_window(r, r_end, HO, w)) {
            const int u0 = u;      // S(ho_a)
            for (int t = w.ho_a - 1; t <= w.ho_b + 3; ++t) {
                const int t2 = t + 1;                                  // tile of this iteration's G2
                if (t2 - 2 >= w.ho_a && t2 <= w.ho_b + 3) {            // G2: S(t2-2) += P_t2 [W2;W3]^T
                    const int g = gbase + (t2 - w.ho_a);
                    while (gready < g) {
                        ++gready;
                        mbar_wait(&bar_full[gready % RING], (uint32_t)((gready / RING) & 1));
                    }
                    tc_fence_after_sync();
                    const uint64_t da = smem_desc_k_sw128(smem_u32(smem + (g % RING) * blk));
                    const int us = u0 + (t2 - 2 - w.ho_a);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma_bf16_ss(tmem + (us % NACC) * 128, da + 2 * ks, db23 + 2 * ks, idesc, 1u);
                    if (t2 - 3 >= w.ho_a) {        // output row t2-3 = S(t2-3).lo + S(t2-2).hi is complete
                        mma_commit(&row_full[v % NACC]);
                        ++v;
                    }
                }
                if (t < w.ho_a) continue;
                const int g = gbase + (t - w.ho_a);
                if (t <= w.ho_b + 1) {            // G1: S(t) = P_t [W0;W1]^T
                    while (gready < g) {
                        ++gready;
                        mbar_wait(&bar_full[gready % RING], (uint32_t)((gready / RING) & 1));
                    }
                    const uint64_t da = smem_desc_k_sw128(smem_u32(smem + (g % RING) * blk));
                    const int us = u0 + (t - w.ho_a);
                    mbar_wait(&acc_empty[us % NACC], (uint32_t)(((us / NACC) & 1) ^ 1));
                    tc_fence_after_sync();
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) mma_bf16_ss(tmem + (us % NACC) * 128, da + 2 * ks, db01 + 2 * ks, idesc, ks ? 1u : 0u);
                }
                mma_commit(&bar_empty[g % RING]);        // both uses of tile t have been issued
            }
            gbase += w.ho_b - w.ho_a + 4;
            u = u0 + (w.ho_b - w.ho_a + 2);
        }
    } else if (warp >= 4) {
        // the epilogue is issue-bound (per output row and pixel: 64 sums, 128 statistics FMAs, bf16 packing, staging): two
        // warps share a pixel group, warp (q, half) takes channels [32 half, 32 half + 32)
        const int q = warp & 3, half = (warp >> 2) - 1;
        const int row = q * 32 + lane;        // pixel within the image row
        // per-pixel-group staging block [32 pixels][128 B], 16-byte chunks XOR-swizzled by the pixel index: conflict-free
        // both for the row-per-thread writes and for the read-back in global-memory order (the group's 32 pixels are 4 KB
        // of contiguous NHWC output, stored with fully coalesced 512-byte requests)
        uint8_t* stg = Ws + 4 * 8192 + 4096 + q * 4096;
        float s1[32], s2[32];                  // BatchNorm statistics of this thread's pixel column over all its rows
#pragma unroll
        for (int c = 0; c < 32; ++c) s1[c] = s2[c] = 0.f;
        int u = 0, ci = 0;
        long long r = r_begin;
        RowWindow w;
        while (next_window(r, r_end, HO, w)) {
            const int u0 = u;
            for (int ho = w.ho_a; ho <= w.ho_b; ++ho, ++ci) {
                mbar_wait(&row_full[ci % NACC], (uint32_t)((ci / NACC) & 1));
                tc_fence_after_sync();
                const int slo = (u0 + ho - w.ho_a) % NACC, shi = (u0 + ho - w.ho_a + 1) % NACC;
                float v0[32];
                {
                    float h0[32];
                    tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(slo * 128 + 32 * half), v0);
                    tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(shi * 128 + 64 + 32 * half), h0);
#pragma unroll
                    for (int c = 0; c < 32; ++c) v0[c] += h0[c];
                }
                tc_fence_before_sync();
                // all eight epilogue warps have read the slots -> S(ho) (and, after a window's last row, S(ho+1)) go back to the MMA warp
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (tid == 128) {
                    mbar_arrive(&acc_empty[slo]);
                    if (ho == w.ho_b) mbar_arrive(&acc_empty[shi]);
                }
                if (sums != nullptr && row < WO) {
#pragma unroll
                    for (int c = 0; c < 32; ++c) { s1[c] += v0[c]; s2[c] = fmaf(v0[c], v0[c], s2[c]); }
                }
                if (kDirectStore) {
                    // this thread's 32 channels = 64 contiguous bytes of the NHWC row: four 16-byte stores, no staging, no barriers
                    if (row < WO) {
                        uint4* dst = y + ((r_begin + ci) * WO + row) * 8 + 4 * half;
#pragma unroll
                        for (int c8 = 0; c8 < 4; ++c8) dst[c8] = pack8_bf16(v0 + 8 * c8);
                    }
                } else {
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8)
                        *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 * half + c8) ^ (lane & 7)) << 4)) = pack8_bf16(v0 + 8 * c8);
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");   // the two warps of this pixel group
                    uint4* dst = y + ((r_begin + ci) * WO + q * 32) * 8;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int g = (4 * half + j) * 32 + lane, pr = g >> 3, c = g & 7;   // chunk g of the group's block: pixel pr, chunk c
                        if (q * 32 + pr < WO) dst[g] = *reinterpret_cast<const uint4*>(stg + pr * 128 + ((c ^ (pr & 7)) << 4));
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(2 + q) : "memory");   // staging block free for the next row
                }
            }
            u = u0 + (w.ho_b - w.ho_a + 2);
        }
        if (sums != nullptr) {
            // cross-pixel reduction of the epilogue threads through the (now idle) tile ring, one double atomic per
            // (statistic, channel) and CTA
            asm volatile("bar.sync 1, 256;" ::: "memory");   // every MMA of this CTA has completed (all rows consumed)
            float* red = reinterpret_cast<float*>(smem);      // [128 values][129]
            const int pix = q * 32 + lane;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                red[(32 * half + c) * 129 + pix] = s1[c];
                red[(64 + 32 * half + c) * 129 + pix] = s2[c];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (half == 0) {
                float tot = 0.f;
                for (int i = 0; i < 128; ++i) tot += red[pix * 129 + i];
                atomicAdd(&sums[pix], (double)tot);           // pix < 64: sum x of channel pix ; pix >= 64: sum x^2 of channel pix - 64
            }
        }
    }
    (void)nrows;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, NACC * 128);
}

bool use_tma() {   // SD_B200_STEM_WGRAD_TMA=0 selects the cp.async-fed kernel
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SD_B200_STEM_WGRAD_TMA");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled resolved at run time (the library must load on machines without libcuda.so.1); nullptr = no
EncodeFn tensor_map_encoder() {
    static EncodeFn encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess)
            encode = (EncodeFn)fn;
        else
            cudaGetLastError();
    }
    return encode;
}
// the packed image as [patch-row start q][64 elements], 32-byte pitch (overlapping rows), box {64, WO}
bool encode_patch_map(CUtensorMap* tm, const void* xs2d, int N, int Hp, int Wp, int WO) {
    EncodeFn encode = tensor_map_encoder();
    if (!encode) return false;
    bind_primary_context();
    const cuuint64_t gdim[2] = {64, (cuuint64_t)N * Hp * Wp - 3};
    const cuuint64_t gstr[1] = {32};
    const cuuint32_t box[2] = {64, (cuuint32_t)WO};
    const cuuint32_t estr[2] = {1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(xs2d), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns SD_OK and sets *done when the TMA kernel was launched; leaves *done false when this shape / driver cannot
int launch_tma(const void* xs2d, const void* dy, float* dw, int N, int HO, int WO, int Hp, int Wp, cudaStream_t st, bool* done) {
    *done = false;
    if (!use_tma() || WO % 16 != 0 || WO > 128 || WO < 16) return SD_OK;
    const size_t smem = (size_t)TMA_STAGES * 5 * WO * 128 + 1024;
    if (smem > 227 * 1024 - 256) return SD_OK;
    static bool usable = true;
    if (!usable) return SD_OK;
    bind_primary_context();
    // the driver entry point is resolved at run time (the library must load on machines without libcuda.so.1)
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            usable = false;
            return SD_OK;
        }
        encode = (EncodeFn)fn;
    }
    CUtensorMap tmA, tmB;
    {
        const cuuint64_t gdim[2] = {64, (cuuint64_t)N * Hp * Wp - 3};
        const cuuint64_t gstr[1] = {32};   // bytes between consecutive patch rows (they overlap by 3 pixels)
        const cuuint32_t box[2] = {64, (cuuint32_t)WO};
        const cuuint32_t estr[2] = {1, 1};
        if (encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(xs2d), gdim, gstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            usable = false;
            return SD_OK;
        }
    }
    {
        const cuuint64_t gdim[2] = {64, (cuuint64_t)N * HO * WO};
        const cuuint64_t gstr[1] = {128};
        const cuuint32_t box[2] = {64, (cuuint32_t)WO};
        const cuuint32_t estr[2] = {1, 1};
        if (encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy), gdim, gstr, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
            usable = false;
            return SD_OK;
        }
    }
    {   // paired kernel (SD_B200_STEM_WGRAD_PAIR=0 selects the 64 x 256 x 16 form below)
        static int use_pair = -1;
        if (use_pair < 0) { const char* e = getenv("SD_B200_STEM_WGRAD_PAIR"); use_pair = (e && e[0] == '0') ? 0 : 1; }
        const size_t psmem = (size_t)(PT_SLOTS + PD_SLOTS) * WO * 128 + 1024;
        if (use_pair && psmem <= 227 * 1024 - 512) {
            CUtensorMap tmB4;
            const cuuint64_t gdim[4] = {64, (cuuint64_t)WO, (cuuint64_t)HO, (cuuint64_t)N};
            const cuuint64_t gstr[3] = {128, (cuuint64_t)WO * 128, (cuuint64_t)HO * WO * 128};
            const cuuint32_t box[4] = {64, (cuuint32_t)WO, 1, 1};
            const cuuint32_t estr[4] = {1, 1, 1, 1};
            if (encode(&tmB4, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(dy), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS) {
                static size_t pconfigured = 0;
                if (pconfigured < psmem && cudaFuncSetAttribute(stem_wgrad_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                                (int)psmem) == cudaSuccess)
                    pconfigured = psmem;
                if (pconfigured >= psmem) {
                    stem_wgrad_pair_kernel<<<min(148, N), NT, psmem, st>>>(tmA, tmB4, dw, N, HO, WO, Hp, Wp);
                    SD_LAUNCH_CHECK();
                    *done = true;
                    return SD_OK;
                }
                cudaGetLastError();
            }
        }
    }
    static size_t configured = 0;   // dynamic + static shared memory must stay within the 227 KB opt-in limit
    if (configured < smem) {
        if (cudaFuncSetAttribute(stem_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            usable = false;
            return SD_OK;
        }
        configured = smem;
    }
    const long long rows = (long long)N * HO;
    const int grid = (int)min((long long)148, rows);
    const int per = (int)((rows + grid - 1) / grid);
    stem_wgrad_tma_kernel<<<grid, NT, smem, st>>>(tmA, tmB, dw, HO, WO, Hp, Wp, rows, per);
    SD_LAUNCH_CHECK();
    *done = true;
    return SD_OK;
}

}  // namespace

extern "C" int sd_stem_wgrad_s2d_bf16(const void* xs2d, const void* dy, float* dw_s2d, int N, int H, int W, void* stream) {
    if (N <= 0) return SD_OK;
    if (!xs2d || !dy || !dw_s2d || (H & 1) || (W & 1)) return SD_ERR_BAD_ARG;
    const int Hp = (H + 6) / 2, Wp = (W + 6) / 2, HO = H / 2, WO = W / 2;
    if (Hp != HO + 3 || Wp != WO + 3) return SD_ERR_BAD_ARG;
    const long long P = (long long)N * HO * WO;
    cudaStream_t st = (cudaStream_t)stream;
    SD_CUDA(cudaMemsetAsync(dw_s2d, 0, sizeof(float) * 256 * 64, st));
    {
        bool done = false;
        const int rc = launch_tma(xs2d, dy, dw_s2d, N, HO, WO, Hp, Wp, st, &done);
        if (rc != SD_OK) return rc;
        if (done) return SD_OK;
    }
    const long long nchunks = (P + KC - 1) / KC;
    const int grid = (int)min((long long)148, nchunks);
    const int per = (int)((nchunks + grid - 1) / grid);
    static bool configured = false;
    const int smem = STAGES * STAGE_BYTES + 1024;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    stem_wgrad_kernel<<<grid, NT, smem, st>>>((const uint4*)xs2d, (const uint4*)dy, dw_s2d, N, HO, WO, Hp, Wp, P, per);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// y (N, H/2, W/2, 64) bf16 NHWC = conv1 of the packed image with w_s2d[64][256] bf16 (row = output channel,
// column = kh*64 + kw*16 + ci).  Returns SD_E_UNSUPPORTED when the shape / driver cannot use the TMA kernel (the
// caller then runs cuDNN).
extern "C" int sd_stem_fprop_s2d_bf16_stats(const void* xs2d, const void* w_s2d, void* y, double* sums, int N, int H, int W,
                                            void* stream);
extern "C" int sd_stem_fprop_s2d_bf16(const void* xs2d, const void* w_s2d, void* y, int N, int H, int W, void* stream) {
    return sd_stem_fprop_s2d_bf16_stats(xs2d, w_s2d, y, nullptr, N, H, W, stream);
}

// Same, and the epilogue also accumulates the per-channel sums of y and y^2 (fp32 accumulator values, before the bf16
// rounding) into sums[2][64] (zeroed here): the BatchNorm statistics of bn1 without another pass over the 112x112 map.
extern "C" int sd_stem_fprop_s2d_bf16_stats(const void* xs2d, const void* w_s2d, void* y, double* sums, int N, int H, int W,
                                            void* stream) {
    if (N <= 0) return SD_OK;
    if (!xs2d || !w_s2d || !y || (H & 1) || (W & 1)) return SD_ERR_BAD_ARG;
    const int Hp = (H + 6) / 2, Wp = (W + 6) / 2, HO = H / 2, WO = W / 2;
    if (!use_tma() || WO % 8 != 0 || WO > 128 || WO < 8) return SD_ERR_UNSUPPORTED;
    const size_t smem = (size_t)RING * WO * 128 + 4 * 8192 + 4096 + 4 * 4096 + 1024;   // tile ring + weights + M=128 over-read slack + store staging
    if (smem > 227 * 1024 - 256) return SD_ERR_UNSUPPORTED;
    CUtensorMap tmA;
    if (!encode_patch_map(&tmA, xs2d, N, Hp, Wp, WO)) return SD_ERR_UNSUPPORTED;
    static size_t configured = 0;
    if (configured < smem) {
        if (cudaFuncSetAttribute(stem_fprop_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            cudaGetLastError();
            return SD_ERR_UNSUPPORTED;
        }
        configured = smem;
    }
    const long long rows = (long long)N * HO;
    const int grid = (int)min((long long)148, rows);
    const int per = (int)((rows + grid - 1) / grid);
    if (sums) SD_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 128, (cudaStream_t)stream));
    stem_fprop_tma_kernel<<<grid, FNT, smem, (cudaStream_t)stream>>>(tmA, (const uint4*)w_s2d, (uint4*)y, sums, HO, WO, Hp, Wp, rows, per);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
