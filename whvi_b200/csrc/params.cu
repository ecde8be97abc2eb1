// Small parameter-side kernels:
//   reparameterisation g = mu + softplus(rho) * eps  (src/weights.py:43-50, :82-83, :92-93)
//   and its backward;
//   kernel (5): fused Gaussian KL + gradient (src/utils.py:49-71 as called from
//   src/weights.py:52-64; mode 0 reproduces the reference's variance interpretation F6,
//   mode 1 is the statistically consistent sigma^2 form).
// All are O(S*D) or O(D): noise next to the activation traffic, so they are written for
// determinism (fixed reduction order, no atomics), not for speed-of-light.
#include "common.cuh"

namespace whvi {

// torch.nn.functional.softplus, beta = 1, threshold = 20
__device__ __forceinline__ float softplus_f(float r) { return r > 20.f ? r : log1pf(expf(r)); }
__device__ __forceinline__ float sigmoid_f(float r) { return 1.f / (1.f + expf(-r)); }

__global__ void reparam_diag_kernel(const float* __restrict__ mu, const float* __restrict__ rho,
                                    const float* __restrict__ eps, float* __restrict__ g, int64_t S, int64_t D)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= D) return;
    const float m = mu[i], sg = softplus_f(rho[i]);
    for (int64_t s = blockIdx.y; s < S; s += gridDim.y) g[s * D + i] = fmaf(sg, eps[s * D + i], m);
}

// dmu[i] (+)= sum_s dg[s,i];  drho[i] (+)= (sum_s dg[s,i] eps[s,i]) * sigmoid(rho[i])
__global__ void reparam_diag_bwd_kernel(const float* __restrict__ rho, const float* __restrict__ eps,
                                        const float* __restrict__ dg, float* __restrict__ dmu, float* __restrict__ drho,
                                        int64_t S, int64_t D, int accumulate)
{
    const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= D) return;
    float a = 0.f, b = 0.f;
    for (int64_t s = 0; s < S; ++s) {
        const float d = dg[s * D + i];
        a += d;
        b = fmaf(d, eps[s * D + i], b);
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

// One CTA, fixed-order tree: KL value + optional gradients (scaled by `grad_scale`, and
// accumulated into dmu/drho when accumulate != 0).
__global__ void __launch_bounds__(1024)
kl_kernel(const float* __restrict__ mu, const float* __restrict__ rho, float lambda_, int64_t D, int mode,
          float* __restrict__ out, float* __restrict__ dmu, float* __restrict__ drho, float grad_scale, int accumulate)
{
    __shared__ double red[3][32];
    double s_log = 0.0, s_ratio = 0.0, s_mu = 0.0;
    const float inv_l = 1.f / lambda_;
    for (int64_t i = threadIdx.x; i < D; i += blockDim.x) {
        const float m = mu[i], r = rho[i];
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
                        cudaStream_t stream)
{
    const int threads = 256;
    const unsigned gx = static_cast<unsigned>((D + threads - 1) / threads);
    unsigned gy = static_cast<unsigned>(S < 1 ? 1 : (S > 4096 ? 4096 : S));
    reparam_diag_kernel<<<dim3(gx, gy), threads, 0, stream>>>(mu, rho, eps, g, S, D);
    return check_launch("reparam_diag_kernel");
}

int launch_reparam_diag_bwd(const float* rho, const float* eps, const float* dg, float* dmu, float* drho, int64_t S,
                            int64_t D, int accumulate, cudaStream_t stream)
{
    const int threads = 128;
    reparam_diag_bwd_kernel<<<static_cast<unsigned>((D + threads - 1) / threads), threads, 0, stream>>>(rho, eps, dg, dmu,
                                                                                                       drho, S, D, accumulate);
    return check_launch("reparam_diag_bwd_kernel");
}

int launch_kl(const float* mu, const float* rho, float lambda_, int64_t D, int mode, float* out, float* dmu, float* drho,
              float grad_scale, int accumulate, cudaStream_t stream)
{
    int threads = 32;
    while (threads < D && threads < 1024) threads <<= 1;
    kl_kernel<<<1, threads, 0, stream>>>(mu, rho, lambda_, D, mode, out, dmu, drho, grad_scale, accumulate);
    return check_launch("kl_kernel");
}

}  // namespace whvi
