// WHVIStackedMatrix (src/weights.py:111-208) as ONE call per direction.
//
// The reference pads the input to D = next_pow2(n_in) columns (:197-198), runs its G = ceil(n_out / D) square blocks one
// after the other on the same padded input (:179-180), concatenates their outputs, adds the bias and drops the columns
// past n_out (:204-207).  Here the blocks are a second sample axis of the fused layer kernels ("virtual samples"
// s' = block * S + sample, each with its own g; s1 / s2 / bias are picked per block inside the kernels), so the whole
// stack is one reparameterisation launch, one layer launch and one interleave launch, whatever G is:
//
//   forward   [pad x]  ->  g = mu_k + softplus(rho_k) eps  ->  layer_fwd over G*S virtual samples  ->  y[s, b, k D + j]
//   backward  dy[s, b, k D + j] -> (G, S, B, D)  ->  layer_bwd (3 launches)  ->  reparam_bwd  ->  [dx = sum_k dx_k, unpadded]
//
// The bias (1, G*D) and the ReLU that follows the layer ride in the layer kernel (bias vector k D .. k D + D - 1 for
// block k).  Parameter vectors of block k are read at base + k * param_stride, so the blocks' separate nn.Parameters can
// be used where they lie as long as they are evenly spaced (they are once packed by FlatParams or by the module itself).
#include "common.cuh"
#include "engine.cuh"

namespace whvi {

// blocks (G, R, D) -> out (R, n_out), out[r, k D + j] = blocks[k, r, j] for k D + j < n_out      (R = S * B rows)
__global__ void __launch_bounds__(256)
stack_concat_kernel(const float4* __restrict__ blocks, float* __restrict__ out, int64_t R, int D4, int64_t G, int64_t n_out, int vec)
{
    const int64_t total = G * R * D4;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int j4 = static_cast<int>(i % D4);
        const int64_t kr = i / D4, r = kr % R, k = kr / R;
        const int64_t col = k * (int64_t(D4) * 4) + 4 * j4;
        if (col >= n_out) continue;
        const float4 v = blocks[i];
        float* o = out + r * n_out + col;
        if (vec) {
            *reinterpret_cast<float4*>(o) = v;
        } else {
            o[0] = v.x;
            if (col + 1 < n_out) o[1] = v.y;
            if (col + 2 < n_out) o[2] = v.z;
            if (col + 3 < n_out) o[3] = v.w;
        }
    }
}

// the inverse: in (R, n_out) -> blocks (G, R, D), zero where k D + j >= n_out.  With G = 1 this is the zero-padding of the
// input rows to D columns.
__global__ void __launch_bounds__(256)
stack_split_kernel(const float* __restrict__ in, float4* __restrict__ blocks, int64_t R, int D4, int64_t G, int64_t n_out, int vec)
{
    const int64_t total = G * R * D4;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int j4 = static_cast<int>(i % D4);
        const int64_t kr = i / D4, r = kr % R, k = kr / R;
        const int64_t col = k * (int64_t(D4) * 4) + 4 * j4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < n_out) {
            const float* p = in + r * n_out + col;
            if (vec) {
                v = *reinterpret_cast<const float4*>(p);
            } else {
                v.x = p[0];
                if (col + 1 < n_out) v.y = p[1];
                if (col + 2 < n_out) v.z = p[2];
                if (col + 3 < n_out) v.w = p[3];
            }
        }
        blocks[i] = v;
    }
}

// dx[r, j] = sum_k dxb[k, r, j] for j < n_in (blocks ascending: fixed order); rows of dx are n_in wide
__global__ void __launch_bounds__(256)
stack_sum_kernel(const float4* __restrict__ dxb, float* __restrict__ dx, int64_t R, int D4, int64_t G, int64_t n_in, int vec)
{
    const int64_t total = R * D4;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int j4 = static_cast<int>(i % D4);
        const int64_t r = i / D4, col = 4 * int64_t(j4);
        if (col >= n_in) continue;
        float4 a = dxb[i];
        for (int64_t k = 1; k < G; ++k) {
            const float4 q = dxb[k * total + i];
            a.x += q.x, a.y += q.y, a.z += q.z, a.w += q.w;
        }
        float* o = dx + r * n_in + col;
        if (vec) {
            *reinterpret_cast<float4*>(o) = a;
        } else {
            o[0] = a.x;
            if (col + 1 < n_in) o[1] = a.y;
            if (col + 2 < n_in) o[2] = a.z;
            if (col + 3 < n_in) o[3] = a.w;
        }
    }
}

static unsigned grid_for(int64_t total)
{
    int64_t b = (total + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    return static_cast<unsigned>(b < 1 ? 1 : b);
}

int launch_stack_split(const float* in, float* blocks, int64_t R, int64_t D, int64_t G, int64_t n_out, cudaStream_t stream)
{
    const int vec = (n_out % 4 == 0) && aligned16(in);
    stack_split_kernel<<<grid_for(G * R * (D / 4)), 256, 0, stream>>>(in, reinterpret_cast<float4*>(blocks), R, static_cast<int>(D / 4), G, n_out, vec);
    return check_launch("stack_split_kernel");
}

int launch_stack_concat(const float* blocks, float* out, int64_t R, int64_t D, int64_t G, int64_t n_out, cudaStream_t stream)
{
    const int vec = (n_out % 4 == 0) && aligned16(out);
    stack_concat_kernel<<<grid_for(G * R * (D / 4)), 256, 0, stream>>>(reinterpret_cast<const float4*>(blocks), out, R, static_cast<int>(D / 4), G, n_out, vec);
    return check_launch("stack_concat_kernel");
}

int launch_stack_sum(const float* dxb, float* dx, int64_t R, int64_t D, int64_t G, int64_t n_in, cudaStream_t stream)
{
    const int vec = (n_in % 4 == 0) && aligned16(dx);
    stack_sum_kernel<<<grid_for(R * (D / 4)), 256, 0, stream>>>(reinterpret_cast<const float4*>(dxb), dx, R, static_cast<int>(D / 4), G, n_in, vec);
    return check_launch("stack_sum_kernel");
}

}  // namespace whvi
