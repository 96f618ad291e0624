// Hardware probe (test infrastructure, not on any product path): tcgen05.mma with MN-major SWIZZLE_128B operands that are
// SHIFTED VIEWS of one TMA-loaded [rows][64] tile — start addresses that are multiples of 128 B but not of the 1024-byte
// swizzle atom, and MN blocks that overlap (leading byte offset = a few rows).  An implicit-GEMM weight gradient needs exactly
// this: dW[tap] = sum_p x[p + off(tap)] dy[p] with every tap a row-shifted view of the same shared-memory halo tile.
#include "layer_common.cuh"
#include "../../include/sd_b200.h"

using namespace sdlf;

namespace {

__device__ __forceinline__ uint64_t desc_mn_sw128_off(uint32_t addr, uint32_t lbo, uint32_t sbo, int use_base_offset) {
    uint64_t d = smem_desc_mn_sw128(addr, lbo, sbo);
    if (use_base_offset) d |= (uint64_t)((addr >> 7) & 7) << 49;   // matrix-descriptor base offset: row phase inside the atom
    return d;
}

struct DbgParams {
    int rows;              // rows of the X / Y tiles (multiple of 8, <= 128)
    int shift_a, lbo_a;    // A view: first k row, byte offset between its two 64-wide MN blocks
    int shift_b, lbo_b, nblk_b;
    int ksteps, use_base_offset, reps, kmajor;
    int ld_count;          // > 0: warps 1-3 run this many tcgen05.ld (32 lanes x 32 columns) each WHILE the MMA chain executes
    float* D;              // [128][64 * nblk_b]
    long long* cycles;
};

__global__ void __launch_bounds__(128, 1) dbg_shifted_mma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY,
                                                                const DbgParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_l, bar_m, bar_x;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_l, 1); mbar_init(&bar_m, 1); mbar_init(&bar_x, 1); mbar_fence_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const int N = 64 * p.nblk_b;
    if (warp == 0 && elect_one()) {   // ONE elect for the whole issue loop (see elect_one in tc_common.cuh)
        mbar_arrive_expect_tx(&bar_l, 2 * p.rows * 128);
        tma_tile_2d(sbase, &tmX, 0, 0, &bar_l);
        tma_tile_2d(sbase + LTILE, &tmY, 0, 0, &bar_l);
        mbar_wait(&bar_l, 0);
        tc_fence_after_sync();
        const uint32_t idesc = instr_desc_bf16(128, N, p.kmajor ? 0 : 1, p.kmajor ? 0 : 1);
        const long long t0 = clock64();
        if (p.kmajor) {   // timing only: both operands K-major [rows][64], k step = 32 bytes inside the swizzle row
            // shift_b > 0: a tcgen05.commit (to a barrier nobody waits on) after every shift_b instructions
            int n = 0;
            for (int r = 0; r < p.reps; ++r)
                for (int k = 0; k < p.ksteps; ++k) {
                    mma_bf16_ss(tmem, smem_desc_k_sw128(sbase) + 2 * (k & 3), smem_desc_k_sw128(sbase + LTILE) + 2 * (k & 3), idesc,
                                (k > 0 || r > 0) ? 1u : 0u);
                    if (p.shift_b > 0 && (++n & (p.shift_b - 1)) == 0) mma_commit(&bar_x);   // shift_b: a power of two
                }
        } else
        for (int r = 0; r < p.reps; ++r)
            for (int k = 0; k < p.ksteps; ++k) {
                const uint32_t a = sbase + (p.shift_a + 16 * k) * 128, b = sbase + LTILE + (p.shift_b + 16 * k) * 128;
                mma_bf16_ss(tmem, desc_mn_sw128_off(a, p.lbo_a, 1024, p.use_base_offset), desc_mn_sw128_off(b, p.lbo_b, 1024, p.use_base_offset),
                            idesc, (k > 0 || r > 0) ? 1u : 0u);
            }
        mma_commit(&bar_m);
        mbar_wait(&bar_m, 0);
        if (p.cycles) *p.cycles = clock64() - t0;
    }
    if (warp >= 1 && p.ld_count > 0) {   // TMEM reads of columns the MMA chain does not touch, concurrent with it
        mbar_wait(&bar_l, 0);
        float keep = 0.f;
        const long long t0 = clock64();
        for (int i = 0; i < p.ld_count; ++i) {
            float v[32];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(256 + 32 * (i & 3)), v);
            keep += v[0] + v[31];
        }
        const long long t1 = clock64();
        if ((tid & 31) == 0 && p.cycles) p.cycles[warp] = t1 - t0;
        if (keep == 123456.789f) p.D[0] = keep;
    }
    __syncthreads();
    mbar_wait(&bar_m, 0);
    tc_fence_after_sync();
    for (int c0 = 0; c0 < N; c0 += 32) {
        float v[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 32; ++i) p.D[(long long)tid * N + c0 + i] = v[i];
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// TMEM read rate: `nwarps` warps (warp w reads lane quadrant w % 4) each run `count` tcgen05.ld.32x32b.x32 (4 KB per warp and
// load); mode 0: wait after every load, mode 1: two loads in flight, mode 2: .x16 loads (2 KB), wait after each
__global__ void __launch_bounds__(512, 1) dbg_ldtm_kernel(int count, int mode, long long* cycles, float* sink) {
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t t = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    float keep = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int i = 0; i < count; ++i) {
            float v[32];
            tmem_ld_32x32(t + 32 * (i & 3), v);
            keep += v[0] + v[31];
        }
    } else if (mode == 1) {
        for (int i = 0; i < count; i += 2) {
            uint32_t a[32], b[32];
            tmem_ld_32x32_issue(t + 32 * (i & 3), a);
            tmem_ld_32x32_issue(t + 32 * ((i + 1) & 3), b);
            tmem_ld_wait_32(a);
            tmem_ld_wait_32(b);
            keep += __uint_as_float(a[0]) + __uint_as_float(b[31]);
        }
    } else {
        for (int i = 0; i < count; ++i) {
            float v[16];
            tmem_ld_32x16(t + 16 * (i & 7), v);
            keep += v[0] + v[15];
        }
    }
    const long long t1 = clock64();
    if ((tid & 31) == 0) cycles[warp] = t1 - t0;
    if (keep == 123456.789f) sink[0] = keep;
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 512);
}

}  // namespace

extern "C" int sd_debug_shifted_mma(const void* X, const void* Y, int rows, int shift_a, int lbo_a, int shift_b, int lbo_b, int nblk_b,
                                    int ksteps, int use_base_offset, int reps, float* D, long long* cycles, void* stream) {
    const int kmajor = (use_base_offset & 2) ? 1 : 0;   // bit 1: K-major operands (issue-rate measurement)
    const int ld_count = (use_base_offset & 4) ? shift_a : 0;   // bit 2: concurrent TMEM-load probe, count in shift_a (cycles[1..3])
    use_base_offset &= 1;
    if (!X || !Y || !D || rows < 8 || rows > 128 || rows % 8 || nblk_b < 1 || nblk_b > 4 || ksteps < 1 || reps < 1) return SD_ERR_BAD_ARG;
    CUtensorMap tmX, tmY;
    if (!encode_bf16_2d(&tmX, X, rows, 64, 64, rows) || !encode_bf16_2d(&tmY, Y, rows, 64, 64, rows)) return SD_ERR_UNSUPPORTED;
    DbgParams p{rows, shift_a, lbo_a, shift_b, lbo_b, nblk_b, ksteps, use_base_offset, reps, kmajor, ld_count, D, cycles};
    const int smem = 2 * LTILE + 2 * LTILE + 1024;   // + slack: N = 256 K-major reads 256 rows of the Y tile
    SD_CUDA(cudaFuncSetAttribute(dbg_shifted_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    dbg_shifted_mma_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmX, tmY, p);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// cycles[16]; returns the per-warp clock counts of `count` TMEM loads run by nwarps warps at once (see dbg_ldtm_kernel)
extern "C" int sd_debug_ldtm(int nwarps, int count, int mode, long long* cycles, float* sink, void* stream) {
    if (nwarps < 1 || nwarps > 16 || count < 2 || (count & 1) || mode < 0 || mode > 2 || !cycles || !sink) return SD_ERR_BAD_ARG;
    dbg_ldtm_kernel<<<1, 32 * nwarps, 0, (cudaStream_t)stream>>>(count, mode, cycles, sink);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
