// Fused, vectorised HBM-bound kernels: diffusion-step token, forward-noising (q-sample),
// DDIM update, MSE loss, joint normaliser, AdamW, embedding gather/scatter, bias-gradient
// column sums, strided row copies (context assembly) and the dropout-mask exporter.
//
// Reference semantics (file:line under /root/reference/soccer_diffusion):
//   StepToken          ml/model/misc.py:25-35
//   add_noise          diffusers DDIMScheduler.add_noise (call site ml/training/train.py:218)
//   DDIM step          diffusers DDIMScheduler.step, eta=0 (call sites ml/inference/ros.py:310,
//                      ml/training/distill.py:189)
//   Normalizer         dataset/pytorch.py:401-414
//   mse_loss           ml/training/train.py:229
//   AdamW              torch.optim.AdamW defaults (ml/training/train.py:162)
//   GameStateEncoder   ml/model/encoder/game_state.py:19-27
#include "common.cuh"
#include "../../include/sd_b200.h"

using namespace sd;

// ------------------------------------------------------------------------------------------
// StepToken:  out[b] = [ sin(t_b * f_k) | cos(t_b * f_k) | token ],  k < d/4
__global__ void step_token_kernel(const void* __restrict__ t, int t_is_float, const float* __restrict__ freqs,
                                  const float* __restrict__ token, float* __restrict__ out, long long ld_out, int B,
                                  int d) {
    const int half = d / 4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * d;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d), c = (int)(i % d);
        float v;
        if (c < 2 * half) {
            const float tf = t_is_float ? ((const float*)t)[b] : (float)((const long long*)t)[b];
            const int k = c < half ? c : c - half;
            const float arg = tf * freqs[k];          // fp32 product, as in the reference
            v = c < half ? sinf(arg) : cosf(arg);     // precise versions: |arg| reaches 999 rad
        } else {
            v = token[c - 2 * half];
        }
        out[b * ld_out + c] = v;
    }
}

extern "C" int sd_step_token(const void* t, int t_is_float, const float* freqs, const float* token, float* out,
                             long long ld_out, int B, int d, void* stream) {
    if (B <= 0) return SD_OK;
    if (d % 4 != 0 || !t || !freqs || !token || !out) return SD_ERR_BAD_ARG;
    const int threads = 128;
    const int blocks = min(ceil_div((long long)B * d, threads), 148 * 8);
    step_token_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(t, t_is_float, freqs, token, out, ld_out, B, d);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// backward of StepToken wrt the learnable half: dtoken[c] += sum_b dout[b][2*half + c]
__global__ void step_token_bwd_kernel(const float* __restrict__ dout, long long ld, int B, int d,
                                      float* __restrict__ dtoken) {
    const int half2 = 2 * (d / 4);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d - half2) return;
    float s = 0.f;
    for (int b = blockIdx.y; b < B; b += gridDim.y) s += dout[b * ld + half2 + c];
    atomicAdd(&dtoken[c], s);
}
extern "C" int sd_step_token_bwd(const float* dout, long long ld, int B, int d, float* dtoken, void* stream) {
    if (B <= 0) return SD_OK;
    const int n = d - 2 * (d / 4);
    dim3 grid(ceil_div(n, 128), min(B, 64));
    step_token_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dout, ld, B, d, dtoken);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// q-sample: x0 = (jc - mean)/std (optional) ; x_t = sqrt(acp[t]) x0 + sqrt(1-acp[t]) eps
__global__ void q_sample_kernel(const float* __restrict__ jc, const float* __restrict__ mean,
                                const float* __restrict__ stdv, const float* __restrict__ noise,
                                const long long* __restrict__ t, const float* __restrict__ acp, int n_train,
                                float* __restrict__ x0_out, float* __restrict__ xt_out, long long total, int inner,
                                int J) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / inner);
        const int j = (int)(i % J);
        float x0 = jc[i];
        if (mean) x0 = (x0 - mean[j]) / stdv[j];
        long long tb = t[b];
        tb = tb < 0 ? 0 : (tb >= n_train ? n_train - 1 : tb);
        const float a = acp[tb];
        const float sa = sqrtf(a), sb = sqrtf(1.0f - a);
        if (x0_out) x0_out[i] = x0;
        xt_out[i] = sa * x0 + sb * noise[i];
    }
}
extern "C" int sd_q_sample(const float* joint_command, const float* mean, const float* stdv, const float* noise,
                           const long long* t, const float* alphas_cumprod, int n_train, float* x0_out, float* xt_out,
                           int B, int inner, int J, void* stream) {
    if (B <= 0) return SD_OK;
    if (!joint_command || !noise || !t || !alphas_cumprod || !xt_out || inner % J != 0) return SD_ERR_BAD_ARG;
    const long long total = (long long)B * inner;
    const int threads = 256;
    const int blocks = min(ceil_div(total, threads), 148 * 8);
    q_sample_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(joint_command, mean, stdv, noise, t, alphas_cumprod,
                                                                 n_train, x0_out, xt_out, total, inner, J);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// DDIM eta=0 update with host-side coefficients:
//   x0_hat = (x - sb*eps)/sa ; prev = sap*x0_hat + sbp*eps
__global__ void ddim_step_kernel(const float* __restrict__ x, const float* __restrict__ eps, float* __restrict__ prev,
                                 float* __restrict__ x0, long long n, float sb, float sa, float sap, float sbp) {
    pdl_trigger();
    pdl_wait();
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float e = eps[i];
        const float p0 = (x[i] - sb * e) / sa;
        if (x0) x0[i] = p0;
        prev[i] = sap * p0 + sbp * e;
    }
}
extern "C" int sd_ddim_step(const float* x, const float* eps, float* prev, float* x0_pred, long long n, float sqrt_beta_t,
                            float sqrt_alpha_t, float sqrt_alpha_prev, float sqrt_beta_prev, void* stream) {
    if (n <= 0) return SD_OK;
    if (!x || !eps || !prev) return SD_ERR_BAD_ARG;
    const int threads = 256;
    const int blocks = min(ceil_div(n, threads), 148 * 8);
    SD_CUDA(launch_chain(ddim_step_kernel, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, x, eps, prev, x0_pred, n, sqrt_beta_t,
                         sqrt_alpha_t, sqrt_alpha_prev, sqrt_beta_prev));
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// MSE (mean over all elements), deterministic single-CTA reduction (n is B*T*J, at most ~1e6)
__global__ void mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                               float* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
        const float d = a[i] - b[i];
        s += (double)d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) out[0] = (float)(s / (double)n);
    }
}
extern "C" int sd_mse_fwd(const float* pred, const float* target, long long n, float* loss_out, void* stream) {
    if (n <= 0 || !pred || !target || !loss_out) return SD_ERR_BAD_ARG;
    mse_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(pred, target, n, loss_out);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
__global__ void mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                               const float* __restrict__ gout, float* __restrict__ grad) {
    const float g = (gout ? gout[0] : 1.0f) * (2.0f / (float)n);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        grad[i] = g * (a[i] - b[i]);
}
extern "C" int sd_mse_bwd(const float* pred, const float* target, long long n, const float* grad_loss, float* grad_pred,
                          void* stream) {
    if (n <= 0 || !pred || !target || !grad_pred) return SD_ERR_BAD_ARG;
    const int blocks = min(ceil_div(n, 256), 148 * 8);
    mse_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred, target, n, grad_loss, grad_pred);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// Normalizer: mode 0 (x-mean)/std, mode 1 x*std+mean over the trailing joint dimension
__global__ void affine_joints_kernel(const float* __restrict__ x, const float* __restrict__ mean,
                                     const float* __restrict__ stdv, float* __restrict__ out, long long n, int J,
                                     int mode) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % J);
        out[i] = mode == 0 ? (x[i] - mean[j]) / stdv[j] : x[i] * stdv[j] + mean[j];
    }
}
extern "C" int sd_affine_joints(const float* x, const float* mean, const float* stdv, float* out, long long n, int J,
                                int mode, void* stream) {
    if (n <= 0) return SD_OK;
    if (!x || !mean || !stdv || !out || J <= 0 || (mode != 0 && mode != 1)) return SD_ERR_BAD_ARG;
    const int blocks = min(ceil_div(n, 256), 148 * 8);
    affine_joints_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, mean, stdv, out, n, J, mode);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// AdamW over one flat fp32 buffer (decoupled weight decay; torch.optim.AdamW semantics):
//   p *= 1 - lr*wd ; m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// 28 B/param/step algorithmic traffic (read p,g,m,v; write p,m,v): float4-vectorised, grid-strided.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                             float step_size, float inv_sqrt_bc2, float gscale) {
    const long long n4 = n >> 2;
    const float decay = 1.0f - lr * wd;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 P = reinterpret_cast<float4*>(p)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i];
        float4 V = reinterpret_cast<float4*>(v)[i];
#define SD_ADAM1(c)                                                      \
    {                                                                    \
        const float gg = G.c * gscale;                                   \
        P.c *= decay;                                                    \
        M.c = b1 * M.c + (1.0f - b1) * gg;                               \
        V.c = b2 * V.c + (1.0f - b2) * gg * gg;                          \
        P.c -= step_size * (M.c / (sqrtf(V.c) * inv_sqrt_bc2 + eps));    \
    }
        SD_ADAM1(x) SD_ADAM1(y) SD_ADAM1(z) SD_ADAM1(w)
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
    }
    // tail
    const long long base = n4 << 2;
    for (long long i = base + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float gg = g[i] * gscale;
        float P = p[i] * decay;
        const float M = b1 * m[i] + (1.0f - b1) * gg;
        const float V = b2 * v[i] + (1.0f - b2) * gg * gg;
        P -= step_size * (M / (sqrtf(V) * inv_sqrt_bc2 + eps));
        p[i] = P;
        m[i] = M;
        v[i] = V;
    }
}
extern "C" int sd_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, float grad_scale, void* stream) {
    if (n <= 0) return SD_OK;
    if (!p || !g || !m || !v || step < 1) return SD_ERR_BAD_ARG;
    if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return SD_ERR_BAD_ARG;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    const float step_size = (float)((double)lr / bc1);
    const float inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const int threads = 256;
    const int blocks = min(ceil_div((n + 3) / 4, threads), 148 * 8);
    adamw_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                              step_size, inv_sqrt_bc2, grad_scale);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// Embedding gather / scatter-add (GameStateEncoder)
__global__ void gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ idx, int rows,
                                   float* __restrict__ out, long long ld_out, int B, int d, int* __restrict__ err) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * d;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d), c = (int)(i % d);
        long long r = idx[b];
        if (r < 0 || r >= rows) {
            if (err) atomicExch(err, 1);
            r = 0;
        }
        out[b * ld_out + c] = table[r * d + c];
    }
}
extern "C" int sd_gather_rows(const float* table, const long long* idx, int rows, float* out, long long ld_out, int B,
                              int d, int* err_flag, void* stream) {
    if (B <= 0) return SD_OK;
    if (!table || !idx || !out) return SD_ERR_BAD_ARG;
    const int blocks = min(ceil_div((long long)B * d, 256), 148 * 8);
    gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(table, idx, rows, out, ld_out, B, d, err_flag);
    SD_LAUNCH_CHECK();
    return SD_OK;
}
__global__ void scatter_add_rows_kernel(const float* __restrict__ dout, long long ld, const long long* __restrict__ idx,
                                        int rows, float* __restrict__ dtable, int B, int d) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * d;
         i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / d), c = (int)(i % d);
        const long long r = idx[b];
        if (r >= 0 && r < rows) atomicAdd(&dtable[r * d + c], dout[b * ld + c]);
    }
}
extern "C" int sd_scatter_add_rows(const float* dout, long long ld, const long long* idx, int rows, float* dtable,
                                   int B, int d, void* stream) {
    if (B <= 0) return SD_OK;
    if (!dout || !idx || !dtable) return SD_ERR_BAD_ARG;
    const int blocks = min(ceil_div((long long)B * d, 256), 148 * 8);
    scatter_add_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dout, ld, idx, rows, dtable, B, d);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// Column sums (bias gradients):  out[n] += sum_m dY[m][n].  Each CTA reduces a slab of rows for a
// 32-column strip in shared memory, then one atomicAdd per column.
__global__ void colsum_kernel(const float* __restrict__ x, long long ld, long long M, int N, float* __restrict__ out,
                              int rows_per_cta) {
    __shared__ float tile[8][33];
    const int col = blockIdx.x * 32 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * rows_per_cta;
    const long long r1 = min(M, r0 + rows_per_cta);
    float s = 0.f;
    if (col < N)
        for (long long r = r0 + threadIdx.y; r < r1; r += 8) s += x[r * ld + col];
    tile[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && col < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += tile[i][threadIdx.x];
        atomicAdd(&out[col], t);
    }
}
extern "C" int sd_colsum_accum(const float* x, long long ld, long long M, int N, float* out, void* stream) {
    if (M <= 0 || N <= 0) return SD_OK;
    if (!x || !out) return SD_ERR_BAD_ARG;
    const int col_blocks = ceil_div(N, 32);
    int row_blocks = max(1, min((int)ceil_div(M, 64), (148 * 4) / col_blocks + 1));
    const int rows_per_cta = ceil_div(M, row_blocks);
    row_blocks = ceil_div(M, rows_per_cta);
    colsum_kernel<<<dim3(col_blocks, row_blocks), dim3(32, 8), 0, (cudaStream_t)stream>>>(x, ld, M, N, out,
                                                                                         rows_per_cta);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// Batched strided 2-D copy: dst[b][r][c] = src[b][r][c] (context assembly without torch.cat)
__global__ void copy_rows_kernel(const float* __restrict__ src, long long sbs, long long sld, float* __restrict__ dst,
                                 long long dbs, long long dld, int B, int rows, int cols, int accumulate) {
    const long long total = (long long)B * rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cols);
        const long long br = i / cols;
        const int r = (int)(br % rows);
        const int b = (int)(br / rows);
        const float v = src[b * sbs + r * sld + c];
        float* o = &dst[b * dbs + r * dld + c];
        *o = accumulate ? *o + v : v;
    }
}
extern "C" int sd_copy_rows(const float* src, long long src_batch_stride, long long src_ld, float* dst,
                            long long dst_batch_stride, long long dst_ld, int B, int rows, int cols, int accumulate,
                            void* stream) {
    if (B <= 0 || rows <= 0 || cols <= 0) return SD_OK;
    if (!src || !dst) return SD_ERR_BAD_ARG;
    const int blocks = min(ceil_div((long long)B * rows * cols, 256), 148 * 16);
    copy_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, src_batch_stride, src_ld, dst, dst_batch_stride,
                                                              dst_ld, B, rows, cols, accumulate);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// y = a + b (residual-gradient joins in the layer backward passes), float4 when aligned
__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = a[i] + b[i];
}
extern "C" int sd_add(const float* a, const float* b, float* y, long long n, void* stream) {
    if (n <= 0) return SD_OK;
    if (!a || !b || !y) return SD_ERR_BAD_ARG;
    const int blocks = min(ceil_div(n, 256), 148 * 16);
    add_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, b, y, n);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
// Dropout-mask exporter: out[i] = 0 or 1/(1-p), the exact factor the fused kernels apply to
// element i of dropout stream `stream_id`.
__global__ void dropout_mask_kernel(float* __restrict__ out, long long n, Dropout dr) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = dr((uint64_t)i);
}
extern "C" int sd_dropout_mask(float* out, long long n, float p, unsigned long long seed, unsigned int stream_id,
                               void* stream) {
    if (n <= 0) return SD_OK;
    if (!out || p < 0.f || p >= 1.f) return SD_ERR_BAD_ARG;
    const int blocks = min(ceil_div(n, 256), 148 * 16);
    dropout_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, n, make_dropout(p, seed, stream_id));
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// ------------------------------------------------------------------------------------------
namespace sd { const unsigned long long* g_dropout_seed_dev = nullptr; int g_pdl = 0; }

extern "C" int sd_set_dropout_seed_offset(const unsigned long long* device_counter) {
    sd::g_dropout_seed_dev = device_counter;
    return SD_OK;
}

// AdamW with every hyper-parameter read from device memory: hp = {lr, beta1, beta2, eps, weight_decay, step_size,
// inv_sqrt_bc2, grad_scale} (step_size = lr / (1 - beta1^t), inv_sqrt_bc2 = 1/sqrt(1 - beta2^t)).  Lets the
// optimizer launch live inside a captured CUDA graph while the learning-rate schedule advances on the host.
__global__ void adamw_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                 float* __restrict__ v, long long n, const float* __restrict__ hp,
                                 const int* __restrict__ step_dev) {
    const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4], gscale = hp[7];
    float step_size = hp[5], inv_sqrt_bc2 = hp[6];
    if (step_dev != nullptr) {
        // the step count lives on the device (advanced inside the captured graph): bias corrections in double, once per block
        __shared__ float bc[2];
        if (threadIdx.x == 0) {
            const double t = (double)max(*step_dev, 1);
            bc[0] = (float)((double)lr / (1.0 - pow((double)b1, t)));
            bc[1] = (float)(1.0 / sqrt(1.0 - pow((double)b2, t)));
        }
        __syncthreads();
        step_size = bc[0];
        inv_sqrt_bc2 = bc[1];
    }
    const long long n4 = n >> 2;
    const float decay = 1.0f - lr * wd;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 P = reinterpret_cast<float4*>(p)[i];
        const float4 G = reinterpret_cast<const float4*>(g)[i];
        float4 M = reinterpret_cast<float4*>(m)[i];
        float4 V = reinterpret_cast<float4*>(v)[i];
#define SD_ADAM2(c)                                                      \
    {                                                                    \
        const float gg = G.c * gscale;                                   \
        P.c *= decay;                                                    \
        M.c = b1 * M.c + (1.0f - b1) * gg;                               \
        V.c = b2 * V.c + (1.0f - b2) * gg * gg;                          \
        P.c -= step_size * (M.c / (sqrtf(V.c) * inv_sqrt_bc2 + eps));    \
    }
        SD_ADAM2(x) SD_ADAM2(y) SD_ADAM2(z) SD_ADAM2(w)
        reinterpret_cast<float4*>(p)[i] = P;
        reinterpret_cast<float4*>(m)[i] = M;
        reinterpret_cast<float4*>(v)[i] = V;
    }
    const long long base = n4 << 2;
    for (long long i = base + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float gg = g[i] * gscale;
        float P = p[i] * decay;
        const float M = b1 * m[i] + (1.0f - b1) * gg;
        const float V = b2 * v[i] + (1.0f - b2) * gg * gg;
        P -= step_size * (M / (sqrtf(V) * inv_sqrt_bc2 + eps));
        p[i] = P;
        m[i] = M;
        v[i] = V;
    }
}
extern "C" int sd_adamw_step_dev(float* p, const float* g, float* m, float* v, long long n, const float* hyper_dev,
                                 const int* step_dev, void* stream) {
    if (n <= 0) return SD_OK;
    if (!p || !g || !m || !v || !hyper_dev) return SD_ERR_BAD_ARG;
    if ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) != 0) return SD_ERR_BAD_ARG;
    const int threads = 256;
    const int blocks = min(ceil_div((n + 3) / 4, threads), 148 * 8);
    adamw_dev_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(p, g, m, v, n, hyper_dev, step_dev);
    SD_LAUNCH_CHECK();
    return SD_OK;
}

extern "C" int sd_abi_version(void) { return SD_B200_ABI_VERSION; }

extern "C" const char* sd_error_string(int code) {
    switch (code) {
        case SD_OK: return "ok";
        case SD_ERR_BAD_ARG: return "sd_b200: bad argument";
        case SD_ERR_UNSUPPORTED: return "sd_b200: unsupported shape/configuration for this kernel";
        case SD_ERR_NO_PLAN: return "sd_b200: null or destroyed plan";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "sd_b200: unknown error";
    }
}

// ------------------------------------------------------------------------------------------
// y = x * dropout_factor(i)  — used by the layer backward passes to re-apply a forward mask
__global__ void dropout_apply_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, Dropout dr) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = x[i] * dr((uint64_t)i);
}
extern "C" int sd_dropout_apply(const float* x, float* y, long long n, float p, unsigned long long seed,
                                unsigned int stream_id, void* stream) {
    if (n <= 0) return SD_OK;
    if (!x || !y || p < 0.f || p >= 1.f) return SD_ERR_BAD_ARG;
    const int blocks = min(ceil_div(n, 256), 148 * 16);
    dropout_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, y, n, make_dropout(p, seed, stream_id));
    SD_LAUNCH_CHECK();
    return SD_OK;
}

// programmatic dependent launch for the kernels launched through launch_chain() (common.cuh); returns the previous setting
extern "C" int sd_set_pdl(int on) {
    const int prev = sd::g_pdl;
    sd::g_pdl = on ? 1 : 0;
    return prev;
}
