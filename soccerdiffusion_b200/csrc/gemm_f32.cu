// True-fp32 (FFMA) tiled GEMM family with fused prologues/epilogues.
//
// fp32 mode must stay within 1e-4 of the reference (BASELINE.json north_star); TF32 tensor-core
// GEMMs measurably break that budget (3.4e-4), so this path is CUDA-core FFMA.
// The bf16 mode routes the same entry point to the tcgen05 kernels in gemm_tc.cu.
//
//   C[m][n] = epilogue( sum_k opA(A)[m][k] * opB(B)[k][n] )
//
//   A stored (M x K) row-major  ("MK")  or (K x M) row-major ("KM")
//   B stored (N x K) row-major  ("NK", a torch Linear weight) or (K x N) row-major ("KN")
//
//   forward   Y  = X  W^T            A=MK  B=NK
//   dgrad     dX = dY W              A=MK  B=KN
//   wgrad     dW = dY^T X            A=KM  B=KN   (split-K over the token dimension, atomics)
//
// Prologue: LayerNorm applied on the fly to the operand that holds activations (A if MK, B if KN)
//           from saved per-row mean / rstd — replaces nn.LayerNorm + nn.Linear
//           (torch/nn/modules/transformer.py:944-950 norm_first branch).
// Epilogue: +bias, save pre-activation, erf-GELU, * gelu'(aux), dropout, + positional-encoding
//           table (ml/model/misc.py:57-65), + residual.
#include "common.cuh"
#include "../../include/sd_b200.h"

using namespace sd;

namespace {

constexpr int BM = 128, BN = 128, BK = 8, PAD = 4, NT = 256;

struct GemmParams {
    const float* A; long long lda;
    const float* B; long long ldb;
    float* C; long long ldc;
    int M, N, K;
    int vecA, vecB;
    // LayerNorm-on-load
    const float* ln_mean; const float* ln_rstd; const float* ln_gamma; const float* ln_beta;
    int ln_on_a, ln_on_b;
    // epilogue
    const float* bias;
    const float* residual; long long ldr;
    const float* pe; int pe_period;
    float* pre_out; long long ldp;
    const float* gelu_grad_src; long long ldg;
    int act;
    int accumulate;   // 1: C += v (atomicAdd when k_slices > 1)
    int k_per_slice;
    float alpha;
    Dropout drop;
};

// load 4 consecutive elements along the contiguous dimension of a row-major matrix (R x Cc),
// zero-filled out of bounds
__device__ __forceinline__ float4 load4(const float* __restrict__ base, long long ld, int r, int c, int R, int Cc,
                                        bool vec) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r >= R) return v;
    const float* p = base + (long long)r * ld + c;
    if (vec && c + 3 < Cc) {
        v = *reinterpret_cast<const float4*>(p);
    } else {
        if (c + 0 < Cc) v.x = p[0];
        if (c + 1 < Cc) v.y = p[1];
        if (c + 2 < Cc) v.z = p[2];
        if (c + 3 < Cc) v.w = p[3];
    }
    return v;
}

__device__ __forceinline__ float4 ln_apply(float4 v, int r, int c, int R, int Cc, const GemmParams& p) {
    if (r >= R) return v;
    const float mu = p.ln_mean[r], rs = p.ln_rstd[r];
    if (c + 0 < Cc) v.x = (v.x - mu) * rs * p.ln_gamma[c + 0] + p.ln_beta[c + 0];
    if (c + 1 < Cc) v.y = (v.y - mu) * rs * p.ln_gamma[c + 1] + p.ln_beta[c + 1];
    if (c + 2 < Cc) v.z = (v.z - mu) * rs * p.ln_gamma[c + 2] + p.ln_beta[c + 2];
    if (c + 3 < Cc) v.w = (v.w - mu) * rs * p.ln_gamma[c + 3] + p.ln_beta[c + 3];
    return v;
}

template <bool A_KM, bool B_KN>
__global__ void __launch_bounds__(NT) gemm_f32_kernel(const GemmParams p) {
    __shared__ __align__(16) float As[2][BK][BM + PAD];
    __shared__ __align__(16) float Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int kb = blockIdx.z * p.k_per_slice;
    const int ke = min(p.K, kb + p.k_per_slice);
    const int tx = tid & 15, ty = tid >> 4;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    // per-thread global-load coordinates
    //   "row-of-4-along-k" layout (MK / NK):  r = tid/2 (0..127), kk = (tid&1)*4
    //   "row-of-4-along-mn" layout (KM / KN): kk = tid/32 (0..7), c4 = (tid&31)*4
    auto fetchA = [&](int k0) -> float4 {
        if (A_KM) {
            const int kk = tid >> 5, c4 = (tid & 31) * 4;
            const int k = k0 + kk;
            return (k < ke) ? load4(p.A, p.lda, k, m0 + c4, p.K, p.M, p.vecA) : make_float4(0, 0, 0, 0);
        } else {
            const int r = tid >> 1, kk = (tid & 1) * 4;
            float4 v = load4(p.A, p.lda, m0 + r, k0 + kk, p.M, ke, p.vecA);
            if (p.ln_on_a) v = ln_apply(v, m0 + r, k0 + kk, p.M, ke, p);
            return v;
        }
    };
    auto fetchB = [&](int k0) -> float4 {
        if (B_KN) {
            const int kk = tid >> 5, c4 = (tid & 31) * 4;
            const int k = k0 + kk;
            float4 v = (k < ke) ? load4(p.B, p.ldb, k, n0 + c4, p.K, p.N, p.vecB) : make_float4(0, 0, 0, 0);
            if (p.ln_on_b && k < ke) v = ln_apply(v, k, n0 + c4, p.K, p.N, p);
            return v;
        } else {
            const int r = tid >> 1, kk = (tid & 1) * 4;
            return load4(p.B, p.ldb, n0 + r, k0 + kk, p.N, ke, p.vecB);
        }
    };
    auto stashA = [&](int buf, float4 v) {
        if (A_KM) {
            const int kk = tid >> 5, c4 = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&As[buf][kk][c4]) = v;
        } else {
            const int r = tid >> 1, kk = (tid & 1) * 4;
            As[buf][kk + 0][r] = v.x; As[buf][kk + 1][r] = v.y; As[buf][kk + 2][r] = v.z; As[buf][kk + 3][r] = v.w;
        }
    };
    auto stashB = [&](int buf, float4 v) {
        if (B_KN) {
            const int kk = tid >> 5, c4 = (tid & 31) * 4;
            *reinterpret_cast<float4*>(&Bs[buf][kk][c4]) = v;
        } else {
            const int r = tid >> 1, kk = (tid & 1) * 4;
            Bs[buf][kk + 0][r] = v.x; Bs[buf][kk + 1][r] = v.y; Bs[buf][kk + 2][r] = v.z; Bs[buf][kk + 3][r] = v.w;
        }
    };

    float4 ra = fetchA(kb), rb = fetchB(kb);
    stashA(0, ra);
    stashB(0, rb);
    __syncthreads();

    int buf = 0;
    for (int k0 = kb; k0 < ke; k0 += BK) {
        const bool has_next = (k0 + BK) < ke;
        if (has_next) {
            ra = fetchA(k0 + BK);
            rb = fetchB(k0 + BK);
        }
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (has_next) {
            stashA(buf ^ 1, ra);
            stashB(buf ^ 1, rb);
        }
        __syncthreads();
        buf ^= 1;
    }

    // epilogue
    const bool split = gridDim.z > 1;
    const bool first_slice = blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (n >= p.N) continue;
            float v = acc[i][j] * p.alpha;
            if (!split) {
                if (p.bias) v += p.bias[n];
                if (p.pre_out) p.pre_out[(long long)m * p.ldp + n] = v;
                if (p.act == SD_ACT_GELU) v = gelu_erf(v);
                if (p.gelu_grad_src) v *= gelu_erf_grad(p.gelu_grad_src[(long long)m * p.ldg + n]);
                v *= p.drop((uint64_t)m * (uint64_t)p.N + (uint64_t)n);
                if (p.pe) v += p.pe[(long long)(m % p.pe_period) * p.N + n];
                if (p.residual) v += p.residual[(long long)m * p.ldr + n];
                float* c = &p.C[(long long)m * p.ldc + n];
                *c = p.accumulate ? *c + v : v;
            } else {
                if (first_slice && p.bias) v += p.bias[n];
                atomicAdd(&p.C[(long long)m * p.ldc + n], v);
            }
        }
    }
}

inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

}  // namespace

int sd_gemm_tc_dispatch(const sd_gemm_desc* d, void* stream);  // gemm_tc.cu

extern "C" int sd_gemm(const sd_gemm_desc* d, void* stream) {
    if (!d) return SD_ERR_BAD_ARG;
    if (d->M <= 0 || d->N <= 0) return SD_OK;
    if (d->K <= 0 || !d->A || !d->B || !d->C) return SD_ERR_BAD_ARG;
    if (d->precision == SD_PREC_BF16) return sd_gemm_tc_dispatch(d, stream);
    if (d->precision != SD_PREC_FP32) return SD_ERR_BAD_ARG;

    GemmParams p;
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
    p.bias = d->bias; p.residual = d->residual; p.ldr = d->ldr; p.pe = d->pe; p.pe_period = d->pe_period > 0 ? d->pe_period : 1;
    p.pre_out = d->pre_out; p.ldp = d->ldp; p.gelu_grad_src = d->gelu_grad_src; p.ldg = d->ldg;
    p.act = d->act; p.accumulate = d->accumulate; p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
    p.drop = make_dropout(d->dropout_p, d->dropout_seed, d->dropout_stream);

    int slices = 1;
    const int tiles = ceil_div(d->M, BM) * ceil_div(d->N, BN);
    if (d->a_layout == SD_LAYOUT_KM) {
        // wgrad: tiny output, huge K -> split K until ~2 waves of CTAs
        if (d->pre_out || d->act != SD_ACT_NONE || d->gelu_grad_src || d->residual || d->pe || d->dropout_p > 0.f)
            return SD_ERR_UNSUPPORTED;
        slices = max(1, min(ceil_div(d->K, 256), (148 * 2 + tiles - 1) / tiles));
    }
    int kps = ceil_div(d->K, slices);
    kps = ceil_div(kps, BK) * BK;
    slices = ceil_div(d->K, kps);
    p.k_per_slice = kps;
    if (slices > 1 && !d->accumulate) {
        // atomics need a zeroed destination
        SD_CUDA(cudaMemset2DAsync(d->C, d->ldc * sizeof(float), 0, (size_t)d->N * sizeof(float), d->M,
                                  (cudaStream_t)stream));
    }
    dim3 grid(ceil_div(d->N, BN), ceil_div(d->M, BM), slices);
    cudaStream_t st = (cudaStream_t)stream;
    if (d->a_layout == SD_LAYOUT_MK && d->b_layout == SD_LAYOUT_NK)
        gemm_f32_kernel<false, false><<<grid, NT, 0, st>>>(p);
    else if (d->a_layout == SD_LAYOUT_MK && d->b_layout == SD_LAYOUT_KN)
        gemm_f32_kernel<false, true><<<grid, NT, 0, st>>>(p);
    else if (d->a_layout == SD_LAYOUT_KM && d->b_layout == SD_LAYOUT_KN)
        gemm_f32_kernel<true, true><<<grid, NT, 0, st>>>(p);
    else if (d->a_layout == SD_LAYOUT_KM && d->b_layout == SD_LAYOUT_NK)
        gemm_f32_kernel<true, false><<<grid, NT, 0, st>>>(p);
    else
        return SD_ERR_BAD_ARG;
    SD_LAUNCH_CHECK();
    return SD_OK;
}
