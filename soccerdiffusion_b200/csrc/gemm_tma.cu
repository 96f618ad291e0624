// TMA-fed tcgen05 GEMMs over bf16 activations in HBM (companions of the layer-fused kernels).
//
// sd_wgrad_bf16 — weight gradients of the linear layers of a transformer layer, all of them in ONE launch:
//     dW_j[n][k] += sum_t G_j[t][n] X_j[t][k]        db_j[n] += sum_t G_j[t][n]
// G_j (gradient activations) and X_j (saved forward activations) are row-major [token][feature] bf16 matrices; a
// [64 tokens][64 features] TMA box with the 128-byte swizzle IS an MN-major tcgen05 operand tile whose K index is the
// token — no transposition anywhere.  grid = (K slices, jobs): each CTA contracts a slice of the tokens with a 3-stage
// TMA -> MMA ring (warp 0 lane 0 produces, warp 1 lane 0 issues), accumulates [128 n][128 k] fp32 in TMEM, and adds its
// partial to the fp32 gradient with vector reductions.  The bias gradient is one more N = 16 MMA per k step against a
// tile of ones (column sums on the tensor core instead of a separate pass over G).
// Replaces autograd of nn.Linear / packed in_proj weights (torch/nn/functional.py:5849-5855) for the bf16 mode.
#include "layer_common.cuh"
#include "../../include/sd_b200.h"

using namespace sdlf;

namespace {

constexpr int WNT = 256;
constexpr int WSTAGES = 3;                // 105 KB of shared memory: two CTAs per SM
constexpr int KT = 64;                    // tokens per stage
constexpr int BOX = KT * 128;             // bytes of one [64 tokens][64 features] box
constexpr int STAGE = 4 * BOX;            // G lo, G hi, X lo, X hi
constexpr int W_SMEM = WSTAGES * STAGE + BOX + 1024;

struct WgradMaps {
    CUtensorMap g[SD_WGRAD_MAX_JOBS];
    CUtensorMap x[SD_WGRAD_MAX_JOBS];
};
struct WgradParams {
    float* dW[SD_WGRAD_MAX_JOBS];
    long long ldw[SD_WGRAD_MAX_JOBS];
    float* db[SD_WGRAD_MAX_JOBS];
    int g_col0[SD_WGRAD_MAX_JOBS];
    int x_col0[SD_WGRAD_MAX_JOBS];
    int nblocks;             // ceil(rows / KT)
    int blocks_per_slice;
};

__global__ void __launch_bounds__(WNT, 2) wgrad_tma_kernel(const __grid_constant__ WgradMaps maps, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[WSTAGES];
    __shared__ __align__(8) uint64_t bar_empty[WSTAGES];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    uint8_t* ones = smem + WSTAGES * STAGE;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int job = blockIdx.y;
    const int b0 = blockIdx.x * p.blocks_per_slice;
    const int nb = max(0, min(p.nblocks, b0 + p.blocks_per_slice) - b0);
    const bool bias = p.db[job] != nullptr;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < WSTAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        mbar_fence_init();
        tma_prefetch_desc(&maps.g[job]);
        tma_prefetch_desc(&maps.x[job]);
    }
    for (int i = tid; i < BOX / 16; i += WNT) reinterpret_cast<uint4*>(ones)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;

    if (warp == 0 && elect_one()) {
        for (int i = 0; i < nb; ++i) {
            const int s = i % WSTAGES;
            mbar_wait(&bar_empty[s], (uint32_t)(((i / WSTAGES) & 1) ^ 1));
            const uint32_t dst = sbase + s * STAGE;
            const int row = (b0 + i) * KT;
            mbar_arrive_expect_tx(&bar_full[s], STAGE);
            tma_tile_2d(dst, &maps.g[job], p.g_col0[job], row, &bar_full[s]);
            tma_tile_2d(dst + BOX, &maps.g[job], p.g_col0[job] + 64, row, &bar_full[s]);
            tma_tile_2d(dst + 2 * BOX, &maps.x[job], p.x_col0[job], row, &bar_full[s]);
            tma_tile_2d(dst + 3 * BOX, &maps.x[job], p.x_col0[job] + 64, row, &bar_full[s]);
        }
    } else if (warp == 1 && elect_one()) {
        const uint32_t id_w = instr_desc_bf16(128, 128, 1, 1);
        const uint32_t id_b = instr_desc_bf16(128, 16, 1, 1);
        const uint64_t d1 = smem_desc_mn_sw128(smem_u32(ones), BOX, 1024);
        for (int i = 0; i < nb; ++i) {
            const int s = i % WSTAGES;
            mbar_wait(&bar_full[s], (uint32_t)((i / WSTAGES) & 1));
            tc_fence_after_sync();
            const uint32_t base = sbase + s * STAGE;
            const uint64_t dg = smem_desc_mn_sw128(base, BOX, 1024), dx = smem_desc_mn_sw128(base + 2 * BOX, BOX, 1024);
            // one accumulation chain per loop (see layer_fused_bwd.cu)
#pragma unroll
            for (int ks = 0; ks < KT / 16; ++ks) mma_bf16_ss(tmem, dg + 128 * ks, dx + 128 * ks, id_w, (i > 0 || ks > 0) ? 1u : 0u);
            if (bias) {
#pragma unroll
                for (int ks = 0; ks < KT / 16; ++ks)
                    mma_bf16_ss(tmem + 128, dg + 128 * ks, d1 + 128 * ks, id_b, (i > 0 || ks > 0) ? 1u : 0u);
            }
            mma_commit(&bar_empty[s]);
            if (i == nb - 1) mma_commit(&bar_done);
        }
    }
    __syncwarp();
    if (nb > 0) {
        mbar_wait(&bar_done, 0);
        tc_fence_after_sync();
        const int q = warp & 3, hf = warp >> 2;
        const int n = q * 32 + lane;
        float* dst = p.dW[job] + (long long)n * p.ldw[job] + 64 * hf;
#pragma unroll
        for (int cb = 0; cb < 2; ++cb) {
            float v[32];
            tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(64 * hf + 32 * cb), v);
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + 32 * cb + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3])
                             : "memory");
        }
        if (bias && hf == 0) {
            float v[16];
            tmem_ld_32x16(tmem + ((uint32_t)(q * 32) << 16) + 128u, v);
            atomicAdd(p.db[job] + n, v[0]);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

extern "C" int sd_wgrad_bf16(const sd_wgrad_job* jobs, int n_jobs, long long rows, void* stream) {
    if (!jobs || n_jobs <= 0 || n_jobs > SD_WGRAD_MAX_JOBS) return SD_ERR_BAD_ARG;
    if (rows <= 0) return SD_OK;
    if (!tensor_map_encoder()) return SD_ERR_UNSUPPORTED;
    WgradMaps maps;
    WgradParams p;
    for (int j = 0; j < n_jobs; ++j) {
        const sd_wgrad_job& jb = jobs[j];
        if (!jb.G || !jb.X || !jb.dW || jb.g_col0 < 0 || jb.x_col0 < 0 || jb.g_col0 + 128 > jb.ldg || jb.x_col0 + 128 > jb.ldx ||
            jb.ldw % 4 != 0 || (((uintptr_t)jb.dW) & 15))
            return SD_ERR_BAD_ARG;
        if (!encode_bf16_2d(&maps.g[j], jb.G, rows, jb.ldg, jb.ldg, KT)) return SD_ERR_UNSUPPORTED;
        if (!encode_bf16_2d(&maps.x[j], jb.X, rows, jb.ldx, jb.ldx, KT)) return SD_ERR_UNSUPPORTED;
        p.dW[j] = jb.dW; p.ldw[j] = jb.ldw; p.db[j] = jb.db; p.g_col0[j] = jb.g_col0; p.x_col0[j] = jb.x_col0;
    }
    p.nblocks = (int)((rows + KT - 1) / KT);
    // K slices: about two waves of CTAs over all jobs, at least 4 token blocks per slice
    int slices = max(1, min((p.nblocks + 3) / 4, (2 * 148 + n_jobs - 1) / n_jobs));
    p.blocks_per_slice = (p.nblocks + slices - 1) / slices;
    slices = (p.nblocks + p.blocks_per_slice - 1) / p.blocks_per_slice;
    static bool configured = false;
    if (!configured) {
        SD_CUDA(cudaFuncSetAttribute(wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W_SMEM));
        configured = true;
    }
    wgrad_tma_kernel<<<dim3(slices, n_jobs), WNT, W_SMEM, (cudaStream_t)stream>>>(maps, p);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
