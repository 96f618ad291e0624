// Shared pieces of the layer-fused tcgen05 kernels (layer_fused_fwd.cu, layer_fused_bwd.cu, gemm_tma.cu):
// TMA tensor-map encoding on the host, TMA tile loads, and the thread <-> accumulator mapping used by every
// fused epilogue.
//
// Thread mapping of a 512-thread CTA working on one 128-row tile (row = TMEM lane):
//   warp w: lane quarter q = w & 3 (the only TMEM lanes the warp may read), column quarter cq = w >> 2
//   thread: row r = 32 q + lane, columns [32 cq, 32 cq + 32) of every 128-column fp32 accumulator.
// The fp32 residual stream of the tile lives in REGISTERS in exactly this mapping (32 floats per thread) from the
// first load to the last store of a layer; bf16 copies are written into 128-byte-swizzled [128][64] operand tiles
// (tile cq >> 1, 16-byte chunks 4 (cq & 1) .. 4 (cq & 1) + 3) for the tensor core.
// (16 warps: the epilogues are issue- and latency-bound, and a kernel this long must stay small enough for the
// instruction cache — a 256-thread / 64-column version ran 8x slower than its instruction count on fetch stalls.)
#pragma once
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace sdlf {

using namespace sd;
using namespace sdtc;

constexpr int LNT = 512;            // threads per CTA
constexpr int LTILE = 128 * 128;    // bytes of one [128 rows][64 bf16] operand tile
constexpr float LN_EPS = 1e-5f;     // nn.LayerNorm default (never overridden by the reference)

// ---- host: cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn encode = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess)
            encode = (EncodeTiledFn)fn;
        else
            cudaGetLastError();
    }
    return encode;
}
// bf16 row-major matrix [rows][cols] (row pitch ld_elems), box = {64 columns, box_rows}: one box lands in shared memory
// as a [box_rows][64] tile with the 128-byte swizzle (the K-major / MN-major operand tile of tc_common.cuh).
// Out-of-range rows / columns are zero-filled on load and dropped on store.
inline bool encode_bf16_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld_elems,
                           int box_rows) {
    EncodeTiledFn encode = tensor_map_encoder();
    if (!encode || (((uintptr_t)base) & 15) || (ld_elems * 2) % 16 != 0 || box_rows < 1 || box_rows > 256) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)ld_elems * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc == CUDA_ERROR_INVALID_CONTEXT) {
        // no context bound to this thread from the driver API's point of view (observed when the caller's other library
        // calls — cuDNN through PyTorch — ran in between): bind this runtime's primary context explicitly and retry
        bind_primary_context();
        rc = encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (rc != CUDA_SUCCESS && getenv("SD_B200_DEBUG"))
        fprintf(stderr, "[sd_b200] cuTensorMapEncodeTiled failed (%d): base %p rows %lld cols %lld ld %lld box_rows %d\n", (int)rc, base,
                rows, cols, ld_elems, box_rows);
    return rc == CUDA_SUCCESS;
}

// ---- device: TMA tile load (SASS: UTMALDG) ----------------------------------------------------------------------------
__device__ __forceinline__ void tma_tile_2d(uint32_t dst_smem, const CUtensorMap* tm, int col0, int row0, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            dst_smem),
        "l"(tm), "r"(col0), "r"(row0), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// ---- accumulator access in the fused-epilogue mapping -------------------------------------------------------------------
struct Lane {
    int tid, warp, lane, q, cq, row, col0;
    uint32_t tlane;   // TMEM lane field of this warp's quarter
    __device__ __forceinline__ Lane() {
        tid = threadIdx.x;
        warp = tid >> 5;
        lane = tid & 31;
        q = warp & 3;
        cq = warp >> 2;
        row = q * 32 + lane;
        col0 = 32 * cq;
        tlane = (uint32_t)(q * 32) << 16;
    }
};
// this thread's 32 columns of the accumulator at TMEM column `col`
__device__ __forceinline__ void ld_acc32(uint32_t tmem, const Lane& L, int col, float* v) {
    tmem_ld_32x32(tmem + L.tlane + (uint32_t)(col + L.col0), v);
}
// 32 fp32 values -> bf16 columns [32 cq, 32 cq + 32) of row r of a 128-wide operand (two [128][64] swizzled tiles at
// `tiles`), and, optionally, the same 64 bytes to global memory
__device__ __forceinline__ void st_row32(uint8_t* tiles, const Lane& L, const float* v, uint4* gsave) {
    uint8_t* tile = tiles + (L.cq >> 1) * LTILE;
    const int c0 = (L.cq & 1) * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint4 u = pack8_bf16(v + 8 * c);
        *reinterpret_cast<uint4*>(tile + sw128_chunk_off(L.row, c0 + c)) = u;
        if (gsave) gsave[c] = u;
    }
}
// 32 consecutive fp32 parameters (16-byte aligned) through the read-only path
__device__ __forceinline__ void ldg32(const float* __restrict__ p, float* out) {
    const float4* g = reinterpret_cast<const float4*>(p);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float4 t = __ldg(g + j);
        out[4 * j] = t.x; out[4 * j + 1] = t.y; out[4 * j + 2] = t.z; out[4 * j + 3] = t.w;
    }
}
__device__ __forceinline__ void unpack8_bf16(const uint4& u, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// D[128][N] (+)= A[128][K] B[N][K]^T with both operands K-major tiles ([rows][64] per 64 k), K = 64 * ktiles.
// a_addr / b_addr: shared addresses of the first k tile; consecutive k tiles are `a_step` / `b_step` bytes apart.
__device__ __forceinline__ void mma_k_tiles(uint32_t tmem_d, uint32_t a_addr, uint32_t a_step, uint32_t b_addr, uint32_t b_step,
                                            uint32_t idesc, int ktiles, bool accumulate) {
    for (int kt = 0; kt < ktiles; ++kt) {
        const uint64_t da = smem_desc_k_sw128(a_addr + kt * a_step), db = smem_desc_k_sw128(b_addr + kt * b_step);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_bf16_ss(tmem_d, da + 2 * j, db + 2 * j, idesc, (accumulate || kt > 0 || j > 0) ? 1u : 0u);
    }
}
// D[128][N] (+)= A[128][K] B[K][N] with A K-major tiles and B MN-major: B tiles are [k rows][64 n] (n contiguous), the two
// 64-wide n blocks `b_lbo` bytes apart, K = 16 * ksteps rows starting at row 0 of the B tiles; A k tile kt at a_addr + kt*a_step.
__device__ __forceinline__ void mma_a_k_b_mn(uint32_t tmem_d, uint32_t a_addr, uint32_t a_step, uint32_t b_addr, uint32_t b_lbo,
                                             uint32_t idesc, int ksteps, bool accumulate) {
    for (int j = 0; j < ksteps; ++j) {
        const uint64_t da = smem_desc_k_sw128(a_addr + (j >> 2) * a_step) + 2 * (j & 3);
        const uint64_t db = smem_desc_mn_sw128(b_addr, b_lbo, 1024) + 128 * (uint64_t)j;
        mma_bf16_ss(tmem_d, da, db, idesc, (accumulate || j > 0) ? 1u : 0u);
    }
}

// ---- cheap transcendental forms for the bf16 path (results are rounded to bf16 right after: |error| << 2^-9) ---------------
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// erf(|z|) by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7): two MUFU ops and seven FMAs instead of erff's ~40
// instructions and two branches; e2 = exp(-z^2) is returned for the derivative of GELU
__device__ __forceinline__ float erf_pos(float z, float& e2) {
    const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
    float poly = fmaf(1.061405429f, t, -1.453152027f);
    poly = fmaf(poly, t, 1.421413741f);
    poly = fmaf(poly, t, -0.284496736f);
    poly = fmaf(poly, t, 0.254829592f);
    e2 = ex2_approx(-1.4426950408889634f * z * z);
    return fmaf(-poly * t, e2, 1.0f);
}
// exact-form (erf) GELU of activation="gelu" (encoder/base.py:36, decoder.py:31), bf16-path evaluation
__device__ __forceinline__ float gelu_fast(float x) {
    float e2;
    const float er = erf_pos(fabsf(x) * 0.70710678118654752440f, e2);
    return 0.5f * x * (1.0f + copysignf(er, x));
}
// d/dx gelu(x) = Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_fast_grad(float x) {
    float e2;   // exp(-x^2 / 2)
    const float er = erf_pos(fabsf(x) * 0.70710678118654752440f, e2);
    return fmaf(0.39894228040143267794f * x, e2, 0.5f * (1.0f + copysignf(er, x)));
}

// row exchange through shared memory: the four threads that share a row (tid & 127 + 128 k) combine a partial value.
// `red` alternates between two [512]-float arrays so that consecutive exchanges need one barrier each.
__device__ __forceinline__ float row_sum(float v, float* red, int tid) {
    red[tid] = v;
    __syncthreads();
    const int b = tid & 127;
    return (red[b] + red[b + 128]) + (red[b + 256] + red[b + 384]);
}
__device__ __forceinline__ float row_max(float v, float* red, int tid) {
    red[tid] = v;
    __syncthreads();
    const int b = tid & 127;
    return fmaxf(fmaxf(red[b], red[b + 128]), fmaxf(red[b + 256], red[b + 384]));
}

// LayerNorm of the register-resident row fragment: returns mean / rstd of the full 128-wide row
__device__ __forceinline__ void row_stats(const float* xr, float* red0, float* red1, int tid, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) s += xr[j];
    mean = row_sum(s, red0, tid) * (1.0f / 128.0f);
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float c = xr[j] - mean;
        s2 = fmaf(c, c, s2);
    }
    rstd = rsqrtf(row_sum(s2, red1, tid) * (1.0f / 128.0f) + LN_EPS);
}

}  // namespace sdlf
