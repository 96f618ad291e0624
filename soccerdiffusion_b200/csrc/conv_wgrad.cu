// Weight gradient of a 3x3, stride-1, padding-1 convolution with 64 input and 64 output channels (the four convolutions of
// ResNet18 layer1 — torchvision resnet.py BasicBlock.conv1/conv2 at 56x56; ml/model/encoder/image.py:55-73) as an implicit
// GEMM on tcgen05 tensor cores fed by TMA:
//
//     dW[co][ci][kh][kw] = sum over images n and pixels p of  dy[n][p][co] * x[n][p + (kh-1, kw-1)][ci]
//
// The contraction runs over PIXELS.  Each image row is held in shared memory as Wp = W + 4 positions (one zero pad on the left,
// three on the right: 4-D TMA boxes start at w = -1, out-of-range elements arrive as zeros), so a row band is ONE flat
// sequence of positions q and a spatial shift (r, s) is the flat shift r*Wp + s; products that pair a pad position of dy with
// anything are zero.  With both operands MN-major (k = position, 128 B per position, 128-byte swizzle on absolute
// shared-memory addresses — tools/probe_shifted_mma.py: row-shifted and overlapping views of one tile are valid operands)
// all nine taps come from shifted VIEWS of the two tiles, no im2col copy:
//     A = x  at positions q + {-1, 0}      -> M = 128 rows (a, ci), the two 64-wide blocks 128 B apart
//     B = dy at positions q + {-Wp, 0, Wp} -> N = 192 columns (j, co), blocks Wp*128 B apart
//     D1[(a,ci)][(j,co)] = dW[co][ci][kh = 2-j][kw = a]          one 128x192x16 MMA per 16 positions (96 clk = the MMA floor)
//     D2[(0,ci)][(j,co)] = dW[co][ci][kh = 2-j][kw = 2]          A = x at q + {+1, +2}; rows 64..127 are discarded
// Persistent CTAs: a contiguous range of (image, 4-row band) work items each, 2-stage TMA ring (x: 4 rows, dy: 6 rows with the
// halo), accumulators stay in TMEM for the whole range; per-CTA partial sums go to a scratch buffer, a second kernel adds them
// in a fixed order (deterministic).
#include "layer_common.cuh"
#include "../../include/sd_b200.h"

using namespace sdlf;

namespace {

constexpr int WG_C = 64;            // channels (both sides)
constexpr int WG_BH = 4;            // q rows per band
constexpr int WG_NT = 192;          // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue
constexpr int WG_STAGES = 2;

typedef CUresult (*EncodeTiledFn4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// bf16 NHWC tensor [N][H][W][64]; box = {64 channels, box_w positions, box_h rows, 1 image}; out-of-range -> zeros
inline bool encode_nhwc64(CUtensorMap* tm, const void* base, long long N, int H, int W, int box_w, int box_h) {
    EncodeTiledFn4 encode = (EncodeTiledFn4)tensor_map_encoder();
    if (!encode || (((uintptr_t)base) & 15) || box_w > 256 || box_h > 256) return false;
    bind_primary_context();
    const cuuint64_t gdim[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t gstr[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
    const cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
__device__ __forceinline__ void tma_tile_4d(uint32_t dst_smem, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            dst_smem),
        "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}

struct WgradParams {
    int H, W, Wp;            // Wp = padded row length in positions (a multiple of 4 with BH = 4: whole k steps per band)
    int bands_per_image;     // ceil((H + 2) / BH): q rows h = -1 .. H
    long long nbands;
    float* partial;          // [gridDim.x][3 j][3 a][64 co][64 ci]
};

__global__ void __launch_bounds__(WG_NT, 1) conv3x3_wgrad_c64_kernel(const __grid_constant__ CUtensorMap tmX,
                                                                     const __grid_constant__ CUtensorMap tmDY, const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[WG_STAGES], bar_empty[WG_STAGES], bar_done;
    __shared__ uint32_t tmem_slot;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int Wp = p.Wp;
    const uint32_t row_bytes = (uint32_t)Wp * 128;
    const uint32_t x_bytes = WG_BH * row_bytes, dy_bytes = (WG_BH + 2) * row_bytes;
    const uint32_t stage_bytes = x_bytes + dy_bytes;
    const uint32_t GUARD = 1024;   // zeros in front of stage 0's x tile: the a = 0 view of the first k step starts one position early

    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
    }
    for (int i = tid; i < (int)GUARD / 16; i += WG_NT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t D1 = 0, D2 = 192;

    // this CTA's contiguous range of bands
    const long long b0 = p.nbands * blockIdx.x / gridDim.x, b1 = p.nbands * (blockIdx.x + 1) / gridDim.x;
    const int nb = (int)(b1 - b0);
    const int ksteps = WG_BH * Wp / 16;

    if (warp == 0 && elect_one()) {          // ---- TMA producer -------------------------------------------------------------
        for (int i = 0; i < nb; ++i) {
            const int s = i % WG_STAGES;
            mbar_wait(&bar_empty[s], (uint32_t)(((i / WG_STAGES) & 1) ^ 1));
            const long long band = b0 + i;
            const int n = (int)(band / p.bands_per_image);
            const int h0 = (int)(band % p.bands_per_image) * WG_BH - 1;   // first q row of the band (q rows start at h = -1)
            const uint32_t dst = sbase + GUARD + s * stage_bytes;
            mbar_arrive_expect_tx(&bar_full[s], stage_bytes);
            tma_tile_4d(dst, &tmX, 0, -1, h0, n, &bar_full[s]);                 // x rows h0 .. h0+3
            tma_tile_4d(dst + x_bytes, &tmDY, 0, -1, h0 - 1, n, &bar_full[s]);  // dy rows h0-1 .. h0+4
        }
    } else if (warp == 1 && elect_one()) {   // ---- MMA issuer ---------------------------------------------------------------
        const uint32_t idesc = instr_desc_bf16(128, 192, 1, 1);
        for (int i = 0; i < nb; ++i) {
            const int s = i % WG_STAGES;
            mbar_wait(&bar_full[s], (uint32_t)((i / WG_STAGES) & 1));
            tc_fence_after_sync();
            const uint32_t xs = sbase + GUARD + s * stage_bytes, ys = xs + x_bytes;
            for (int k = 0; k < ksteps; ++k) {
                const uint32_t q = (uint32_t)k * 16 * 128;
                // B: dy at positions q - Wp (tile row 0 = h0 - 1), q, q + Wp
                const uint64_t db = smem_desc_mn_sw128(ys + q, row_bytes, 1024);
                // A: x at positions q - 1, q   |   q + 1, q + 2
                mma_bf16_ss(tmem + D1, smem_desc_mn_sw128(xs + q - 128, 128, 1024), db, idesc, (i > 0 || k > 0) ? 1u : 0u);
                mma_bf16_ss(tmem + D2, smem_desc_mn_sw128(xs + q + 128, 128, 1024), db, idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
            mma_commit(&bar_empty[s]);
        }
        mma_commit(&bar_done);
    }
    __syncwarp();
    if (warp >= 2) {                       // ---- epilogue: accumulators -> this CTA's partial sums ------------------------
        const int w4 = warp & 3;           // TMEM lane quarter this warp may read
        const int row = w4 * 32 + lane;    // accumulator row = (a, ci)
        const int a = row >> 6, ci = row & 63;
        mbar_wait(&bar_done, 0);
        tc_fence_after_sync();
        float* out = p.partial + (long long)blockIdx.x * (9 * 64 * 64);
#pragma unroll 1
        for (int c0 = 0; c0 < 192; c0 += 32) {
            float v1[32], v2[32];
            tmem_ld_32x32(tmem + ((uint32_t)(w4 * 32) << 16) + D1 + c0, v1);
            tmem_ld_32x32(tmem + ((uint32_t)(w4 * 32) << 16) + D2 + c0, v2);
            const int j = c0 >> 6, co0 = c0 & 63;
            if (nb > 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    out[(((j * 3 + a) * 64) + co0 + i) * 64 + ci] = v1[i];
                    if (a == 0) out[(((j * 3 + 2) * 64) + co0 + i) * 64 + ci] = v2[i];
                }
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    out[(((j * 3 + a) * 64) + co0 + i) * 64 + ci] = 0.f;
                    if (a == 0) out[(((j * 3 + 2) * 64) + co0 + i) * 64 + ci] = 0.f;
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// dW[co][ci][kh][kw] (+)= sum over CTAs of partial[cta][j = 2 - kh][a = kw][co][ci]
__global__ void conv3x3_wgrad_reduce_kernel(const float* __restrict__ partial, int nparts, float* __restrict__ dW, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // index into [j][a][co][ci]
    if (i >= 9 * 64 * 64) return;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += partial[(long long)c * (9 * 64 * 64) + i];
    const int ci = i & 63, co = (i >> 6) & 63, ja = i >> 12, a = ja % 3, j = ja / 3;
    float* d = dW + ((co * 64 + ci) * 3 + (2 - j)) * 3 + a;
    *d = accumulate ? *d + s : s;
}

}  // namespace

extern "C" int sd_conv3x3_wgrad_c64_supported(int H, int W) {
    if (H < 1 || W < 1 || W + 4 > 256 || ((W + 4) * WG_BH) % 16 != 0) return 0;
    const size_t smem = 1024 + 1024 + (size_t)WG_STAGES * (2 * WG_BH + 2) * (W + 4) * 128;
    if (smem > 227 * 1024) return 0;
    return tensor_map_encoder() != nullptr ? 1 : 0;
}

extern "C" int sd_conv3x3_wgrad_c64_scratch_bytes(void) { return 148 * 9 * 64 * 64 * 4; }

extern "C" int sd_conv3x3_wgrad_c64_bf16(const void* x, const void* dy, float* dW, int frames, int H, int W, float* scratch,
                                         int accumulate, void* stream) {
    if (!x || !dy || !dW || !scratch || frames < 0) return SD_ERR_BAD_ARG;
    if (!sd_conv3x3_wgrad_c64_supported(H, W)) return SD_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int Wp = W + 4;
    WgradParams p{};
    p.H = H; p.W = W; p.Wp = Wp;
    p.bands_per_image = (H + 2 + WG_BH - 1) / WG_BH;
    p.nbands = (long long)frames * p.bands_per_image;
    p.partial = scratch;
    CUtensorMap tmX, tmDY;
    if (!encode_nhwc64(&tmX, x, frames, H, W, Wp, WG_BH) || !encode_nhwc64(&tmDY, dy, frames, H, W, Wp, WG_BH + 2)) return SD_ERR_UNSUPPORTED;
    const int grid = (int)std::max<long long>(1, std::min<long long>(148, p.nbands));
    const int smem = 1024 + 1024 + WG_STAGES * (2 * WG_BH + 2) * Wp * 128;
    static int configured = 0;
    if (configured < smem) {
        SD_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_c64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    conv3x3_wgrad_c64_kernel<<<grid, WG_NT, smem, st>>>(tmX, tmDY, p);
    SD_LAUNCH_CHECK();
    conv3x3_wgrad_reduce_kernel<<<(9 * 64 * 64 + 255) / 256, 256, 0, st>>>(scratch, grid, dW, accumulate);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
