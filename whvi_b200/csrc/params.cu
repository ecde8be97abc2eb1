// Small parameter-side kernels:
//   reparameterisation g = mu + softplus(rho) * eps  (src/weights.py:43-50, :82-83, :92-93)
//   and its backward;
//   kernel (5): fused Gaussian KL + gradient (src/utils.py:49-71 as called from
//   src/weights.py:52-64; mode 0 reproduces the reference's variance interpretation F6,
//   mode 1 is the statistically consistent sigma^2 form).
// All are O(S*D) or O(D): noise next to the activation traffic, so they are written for
// determinism (fixed reduction order, no atomics), not for speed-of-light.
#include "common.cuh"
#include "engine.cuh"

namespace whvi {

// torch.nn.functional.softplus, beta = 1, threshold = 20
__device__ __forceinline__ float softplus_f(float r) { return r > 20.f ? r : log1pf(expf(r)); }
__device__ __forceinline__ float sigmoid_f(float r) { return 1.f / (1.f + expf(-r)); }

// grouped form (a Stacked layer's blocks, src/weights.py:179-180): samples s / spg share the parameter vectors at + (s / spg) * pstride
__global__ void reparam_diag_kernel(const float* __restrict__ mu, const float* __restrict__ rho,
                                    const float* __restrict__ eps, float* __restrict__ g, int64_t S, int64_t D, int64_t spg, int64_t pstride)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= D) return;
    if (spg >= S) {
        const float m = mu[i], sg = softplus_f(rho[i]);
        for (int64_t s = blockIdx.y; s < S; s += gridDim.y) g[s * D + i] = fmaf(sg, eps[s * D + i], m);
    } else {
        for (int64_t s = blockIdx.y; s < S; s += gridDim.y) {
            const int64_t pi = (s / spg) * pstride + i;
            g[s * D + i] = fmaf(softplus_f(rho[pi]), eps[s * D + i], mu[pi]);
        }
    }
}

// dmu[i] (+)= sum_s dg[s,i];  drho[i] (+)= (sum_s dg[s,i] eps[s,i]) * sigmoid(rho[i])
// 32 columns x 8 sample slices per CTA; slice sums are combined in a fixed order (bit-reproducible).
__global__ void __launch_bounds__(256)
reparam_diag_bwd_kernel(const float* __restrict__ rho, const float* __restrict__ eps, const float* __restrict__ dg,
                        float* __restrict__ dmu, float* __restrict__ drho, int64_t S, int64_t D, int accumulate, int64_t pstride)
{
    __shared__ float red[2][8][32];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t i = int64_t(blockIdx.x) * 32 + tx;
    // grouped form: blockIdx.y = parameter group; S samples per group, outputs (groups, D)
    rho += int64_t(blockIdx.y) * pstride;
    eps += int64_t(blockIdx.y) * S * D;
    dg += int64_t(blockIdx.y) * S * D;
    dmu += int64_t(blockIdx.y) * D;
    drho += int64_t(blockIdx.y) * D;
    float a = 0.f, b = 0.f;
    if (i < D) {
#pragma unroll 4
        for (int64_t s = ty; s < S; s += 8) {
            const float d = dg[s * D + i];
            a += d;
            b = fmaf(d, eps[s * D + i], b);
        }
    }
    red[0][ty][tx] = a;
    red[1][ty][tx] = b;
    __syncthreads();
    if (ty == 0 && i < D) {
#pragma unroll
        for (int j = 1; j < 8; ++j) {
            a += red[0][j][tx];
            b += red[1][j][tx];
        }
        b *= sigmoid_f(rho[i]);
        if (accumulate) {
            dmu[i] += a;
            drho[i] += b;
        } else {
            dmu[i] = a;
            drho[i] = b;
        }
    }
}

// One CTA, fixed-order tree: KL value + optional gradients (scaled by `grad_scale`, and
// accumulated into dmu/drho when accumulate != 0).
__global__ void __launch_bounds__(1024)
kl_kernel(const float* __restrict__ mu, const float* __restrict__ rho, float lambda_, int64_t D, int mode,
          float* __restrict__ out, float* __restrict__ dmu, float* __restrict__ drho, float grad_scale, int accumulate,
          int64_t Dg, int64_t pstride)   // grouped form: D = groups * Dg coordinates, group k's vectors at + k * pstride; outputs contiguous
{
    __shared__ double red[3][32];
    double s_log = 0.0, s_ratio = 0.0, s_mu = 0.0;
    const float inv_l = 1.f / lambda_;
    for (int64_t i = threadIdx.x; i < D; i += blockDim.x) {
        const int64_t pi = Dg >= D ? i : (i / Dg) * pstride + i % Dg;
        const float m = mu[pi], r = rho[pi];
        const float sg = softplus_f(r);
        const float v = mode ? sg * sg : sg;
        s_log += static_cast<double>(logf(v));
        s_ratio += static_cast<double>(v * inv_l);
        s_mu += static_cast<double>(m * m * inv_l);
        if (dmu) {
            const float gm = grad_scale * m * inv_l;
            const float dv = mode ? 2.f * sg : 1.f;
            const float gr = grad_scale * 0.5f * (inv_l - 1.f / v) * dv * sigmoid_f(r);
            if (accumulate) {
                dmu[i] += gm;
                drho[i] += gr;
            } else {
                dmu[i] = gm;
                drho[i] = gr;
            }
        }
    }
    const unsigned full = 0xffffffffu;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s_log += __shfl_xor_sync(full, s_log, o);
        s_ratio += __shfl_xor_sync(full, s_ratio, o);
        s_mu += __shfl_xor_sync(full, s_mu, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        red[0][warp] = s_log;
        red[1][warp] = s_ratio;
        red[2][warp] = s_mu;
    }
    __syncthreads();
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        double a = lane < nw ? red[0][lane] : 0.0, b = lane < nw ? red[1][lane] : 0.0, c = lane < nw ? red[2][lane] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(full, a, o);
            b += __shfl_xor_sync(full, b, o);
            c += __shfl_xor_sync(full, c, o);
        }
        if (lane == 0) {
            const double d = static_cast<double>(D);
            out[0] = static_cast<float>(0.5 * (d * log(static_cast<double>(lambda_)) - a - d + b + c));
        }
    }
}

int launch_reparam_diag(const float* mu, const float* rho, const float* eps, float* g, int64_t S, int64_t D,
                        cudaStream_t stream, int64_t groups, int64_t pstride)
{
    const int threads = 256;
    const unsigned gx = static_cast<unsigned>((D + threads - 1) / threads);
    unsigned gy = static_cast<unsigned>(S < 1 ? 1 : (S > 4096 ? 4096 : S));
    reparam_diag_kernel<<<dim3(gx, gy), threads, 0, stream>>>(mu, rho, eps, g, S, D, S / groups, pstride);
    return check_launch("reparam_diag_kernel");
}

// S: samples PER GROUP
int launch_reparam_diag_bwd(const float* rho, const float* eps, const float* dg, float* dmu, float* drho, int64_t S,
                            int64_t D, int accumulate, cudaStream_t stream, int64_t groups, int64_t pstride)
{
    reparam_diag_bwd_kernel<<<dim3(static_cast<unsigned>((D + 31) / 32), static_cast<unsigned>(groups)), 256, 0, stream>>>(
        rho, eps, dg, dmu, drho, S, D, accumulate, pstride);
    return check_launch("reparam_diag_bwd_kernel");
}

// D: coordinates PER GROUP
int launch_kl(const float* mu, const float* rho, float lambda_, int64_t D, int mode, float* out, float* dmu, float* drho,
              float grad_scale, int accumulate, cudaStream_t stream, int64_t groups, int64_t pstride)
{
    const int64_t total = D * groups;
    int threads = 32;
    while (threads < total && threads < 1024) threads <<= 1;
    kl_kernel<<<1, threads, 0, stream>>>(mu, rho, lambda_, total, mode, out, dmu, drho, grad_scale, accumulate, D, pstride);
    return check_launch("kl_kernel");
}

// ---- MC predictive moments: sum over the sample axis of y and y^2 ----------------
// One float4 column per thread, samples ascending (bit-reproducible), 4 independent loads in
// flight per thread; HBM-bound: reads 4*S*n bytes.  out = in + sum_s (in == NULL: 0).  `out` may
// live in a PEER GPU's memory (NVLink-mapped): the reduction and the scatter to the rank that owns
// these rows are then one kernel -- the stores ride over NVLink while the loads stream from HBM.
__global__ void __launch_bounds__(256)
mc_moments_kernel(const float4* __restrict__ y, int64_t stride4, const float4* __restrict__ in_y, const float4* __restrict__ in_y2,
                  float4* __restrict__ out_y, float4* __restrict__ out_y2, int64_t S, int64_t n4)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (in_y) a = in_y[i];
    if (in_y2) b = in_y2[i];
    auto add = [&](const float4 q) {
        a.x += q.x, a.y += q.y, a.z += q.z, a.w += q.w;
        b.x = fmaf(q.x, q.x, b.x), b.y = fmaf(q.y, q.y, b.y), b.z = fmaf(q.z, q.z, b.z), b.w = fmaf(q.w, q.w, b.w);
    };
    int64_t s = 0;
    for (; s + 4 <= S; s += 4) {
        const float4 q0 = ldg_stream(reinterpret_cast<const float*>(y + (s + 0) * stride4 + i));
        const float4 q1 = ldg_stream(reinterpret_cast<const float*>(y + (s + 1) * stride4 + i));
        const float4 q2 = ldg_stream(reinterpret_cast<const float*>(y + (s + 2) * stride4 + i));
        const float4 q3 = ldg_stream(reinterpret_cast<const float*>(y + (s + 3) * stride4 + i));
        add(q0), add(q1), add(q2), add(q3);
    }
    for (; s < S; ++s) add(ldg_stream(reinterpret_cast<const float*>(y + s * stride4 + i)));
    out_y[i] = a;
    if (out_y2) out_y2[i] = b;
}

int launch_mc_moments(const float* y, int64_t y_sample_stride, const float* in_y, const float* in_y2, float* out_y,
                      float* out_y2, int64_t S, int64_t n, cudaStream_t stream)
{
    const int64_t n4 = n / 4;
    const int64_t blocks = (n4 + 255) / 256;
    if (blocks > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "mc_moments: n too large");
    mc_moments_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(y), y_sample_stride / 4, reinterpret_cast<const float4*>(in_y),
        reinterpret_cast<const float4*>(in_y2), reinterpret_cast<float4*>(out_y), reinterpret_cast<float4*>(out_y2), S, n4);
    return check_launch("mc_moments_kernel");
}

// ---- fused Adam over one flat parameter buffer (the optimizer.step() of the reference's step loop,
// src/networks.py:80-82 / :92-94, with torch.optim.Adam's arithmetic: no weight decay, no amsgrad) ---------------
//   m = m + (1 - b1) (g - m);  v = b2 v + (1 - b2) g^2;  p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// `step` is the device-resident step count t >= 1 (already incremented by the caller) and `lr_dev`, when given, a
// device scalar, so that a captured CUDA graph sees learning-rate schedules and step counts without re-capture.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
            float lr_host, const float* __restrict__ lr_dev, const float* __restrict__ step, float b1, float b2, float eps,
            float grad_scale)
{
    const float t = __ldg(step);
    const float lr = lr_dev ? __ldg(lr_dev) : lr_host;
    const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(b1), static_cast<double>(t)));
    const float bc2_sqrt = sqrtf(static_cast<float>(1.0 - pow(static_cast<double>(b2), static_cast<double>(t))));
    const float step_size = lr / bc1;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float gi = g[i] * grad_scale;
        const float mi = m[i] + (1.f - b1) * (gi - m[i]);           // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(1.f - b2, gi * gi, b2 * v[i]);        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] -= step_size * (mi / denom);
    }
}

int launch_adam(float* p, const float* g, float* m, float* v, int64_t n, float lr, const float* lr_dev, const float* step,
                float b1, float b2, float eps, float grad_scale, cudaStream_t stream)
{
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    adam_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p, g, m, v, n, lr, lr_dev, step, b1, b2, eps, grad_scale);
    return check_launch("adam_kernel");
}

}  // namespace whvi
