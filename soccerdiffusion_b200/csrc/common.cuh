// Shared device helpers for the sm_100a kernels of soccerdiffusion_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#define SD_OK 0
#define SD_ERR_BAD_ARG (-1)
#define SD_ERR_UNSUPPORTED (-2)
#define SD_ERR_NO_PLAN (-3)

#define SD_LAUNCH_CHECK()                         \
    do {                                          \
        cudaError_t e__ = cudaGetLastError();     \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

#define SD_CUDA(x)                                \
    do {                                          \
        cudaError_t e__ = (x);                    \
        if (e__ != cudaSuccess) return (int)e__;  \
    } while (0)

namespace sd {

constexpr int kWarp = 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// exact (erf) GELU — torch.nn.functional.gelu(approximate="none")
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
// d/dx gelu(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
    return cdf + x * pdf;
}

// Stateless counter-based RNG for dropout: one 32-bit draw per (seed, stream, element).
// The (seed, stream) pair is mixed into two 32-bit keys (loop invariant: computed once per thread), the element index is
// combined with them and finalised with a 32-bit avalanche mixer (two multiplies, three xor-shifts) — the draw sits in
// the inner loop of every fused epilogue, which is issue-bound (the previous 64-bit splitmix finaliser cost ~25
// instructions per element).  The same function is exported through sd_dropout_mask so a test can hand the oracle
// exactly the masks the fused kernels used.
__host__ __device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint32_t stream, uint64_t idx) {
    uint64_t k = seed * 0x9E3779B97F4A7C15ull + (((uint64_t)stream << 32) | stream) * 0xD6E8FEB86659FD93ull;
    k ^= k >> 32;
    k *= 0xD6E8FEB86659FD93ull;
    k ^= k >> 32;
    uint32_t h = ((uint32_t)idx ^ (uint32_t)k) * 0x9E3779B1u + ((uint32_t)(idx >> 32) ^ (uint32_t)(k >> 32)) * 0xC2B2AE3Du;
    h ^= h >> 16;
    h *= 0x7FEB352Du;
    h ^= h >> 15;
    h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}
// returns 0 (dropped) or 1/(1-p) (kept)
__host__ __device__ __forceinline__ float dropout_scale(uint64_t seed, uint32_t stream, uint64_t idx, uint32_t thresh,
                                                        float inv_keep) {
    return hash_u32(seed, stream, idx) >= thresh ? inv_keep : 0.0f;
}
__host__ __device__ __forceinline__ uint32_t dropout_threshold(float p) {
    double t = (double)p * 4294967296.0;
    if (t < 0) t = 0;
    if (t > 4294967295.0) t = 4294967295.0;
    return (uint32_t)t;
}

struct Dropout {
    uint64_t seed;
    const unsigned long long* seed_dev;   // optional device-resident offset added to `seed` (CUDA-graph replays: the
                                          // host-side seed is frozen in the graph, the device counter advances)
    uint32_t stream;
    uint32_t thresh;   // 0 => disabled
    float inv_keep;
    __host__ __device__ __forceinline__ float operator()(uint64_t idx) const {
        if (thresh == 0) return 1.0f;
#ifdef __CUDA_ARCH__
        const uint64_t s = seed_dev ? seed + __ldg(seed_dev) : seed;
#else
        const uint64_t s = seed;
#endif
        return dropout_scale(s, stream, idx, thresh, inv_keep);
    }
#ifdef __CUDACC__
    // hot loops: read the device seed offset once per thread, then draw with at()
    __device__ __forceinline__ uint64_t resolve() const {
        return (thresh != 0 && seed_dev) ? seed + __ldg(seed_dev) : seed;
    }
    __device__ __forceinline__ float at(uint64_t resolved_seed, uint64_t idx) const {
        return thresh == 0 ? 1.0f : dropout_scale(resolved_seed, stream, idx, thresh, inv_keep);
    }
#endif
};
// process-wide device seed offset (sd_set_dropout_seed_offset); defined in elementwise.cu
extern const unsigned long long* g_dropout_seed_dev;
static inline Dropout make_dropout(float p, uint64_t seed, uint32_t stream) {
    Dropout d;
    d.seed = seed;
    d.seed_dev = g_dropout_seed_dev;
    d.stream = stream;
    d.thresh = (p > 0.f) ? dropout_threshold(p) : 0u;
    d.inv_keep = (p > 0.f) ? 1.0f / (1.0f - p) : 1.0f;
    return d;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// cuTensorMapEncodeTiled is a DRIVER call: it fails with CUDA_ERROR_INVALID_CONTEXT when no context is bound to the calling
// thread from the driver API's point of view — observed when another library (cuDNN through PyTorch) ran between two calls
// into this one and no kernel of this library's own runtime instance had been launched on the thread yet.  Entry points that
// encode tensor maps bind the runtime's primary context first (two sub-microsecond runtime calls; both are legal inside a
// stream capture — cudaFree is not, so the classic cudaFree(0) idiom is not used here).
static inline void bind_primary_context() {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaSetDevice(dev);
}

// ---- programmatic dependent launch (sd_set_pdl) ----------------------------------------------------------------------
// Chains of short kernels (the tensor-core sampler: 16 launches per DDIM step): with the launch attribute set, kernel N+1
// is scheduled as soon as every CTA of kernel N has executed pdl_trigger(), runs its prologue (barrier init, TMEM
// allocation, tensor-map prefetch, weight TMA loads) under kernel N's tail, and blocks in pdl_wait() until kernel N has
// completed and its writes are visible.  Rules kept by every kernel launched through launch_chain():
//   * pdl_wait() precedes the first read of anything a preceding kernel of the chain may have written and every global
//     write; only data produced by NON-triggering kernels (packed weights, biases, LayerNorm parameters) is read earlier
//     (a non-triggering kernel releases its successor at completion, so nothing later in the stream starts before it ends);
//   * without the attribute both instructions are no-ops.
extern int g_pdl;   // defined in elementwise.cu
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// stem_band.cu: row-band kernels for maxpool3x3s2(relu(bn(x))) with 64 channels (used by trunk_ops.cu's entry points)
bool stem_band_supported(int H, int W, int C);
int stem_band_fwd(const void* x, const float* mean, const float* invstd, const float* gamma, const float* beta, void* y,
                  void* idx, int N, int H, int W, cudaStream_t st);
int stem_band_bwd(const void* dpool, const void* idx, const void* x, const void* y_pooled, const float* mean,
                  const float* invstd, const float* gamma, const float* beta, double* sums, void* dx, int N, int H, int W,
                  int pass, cudaStream_t st);

}  // namespace sd
