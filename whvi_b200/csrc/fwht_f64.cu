// fp64 batched FWHT: the reference dispatches double as well
// (src/fwht/cuda/fwht_cuda_kernel.cu:170) and its gradient check runs in double
// (src/fwht/grad_check.py:26).  This is parity tooling, not a bandwidth kernel: a plain
// shared-memory butterfly network per segment of up to 4096 doubles, then -- for longer rows --
// one global radix-2 stage per remaining bit.  Same Sylvester order, unnormalised.
#include "common.cuh"

namespace whvi {

constexpr int kSegLog2 = 12;  // 4096 doubles = 32 KB of shared memory

// One CTA per segment of 2^seg_k doubles (a whole row, several rows, or a piece of a long row).
__global__ void __launch_bounds__(256) fwht_f64_segment_kernel(const double* __restrict__ in, double* __restrict__ out,
                                                               int64_t total, int seg_k, int k)
{
    extern __shared__ double sm[];
    const int64_t base = int64_t(blockIdx.x) << seg_k;
    const int n = 1 << seg_k;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sm[i] = (base + i < total) ? in[base + i] : 0.0;
    __syncthreads();
    const int stages = k < seg_k ? k : seg_k;  // rows shorter than the segment: only their own bits
    for (int b = 0; b < stages; ++b) {
        const int h = 1 << b;
        for (int p = threadIdx.x; p < n / 2; p += blockDim.x) {
            const int i = ((p >> b) << (b + 1)) | (p & (h - 1));
            const double u = sm[i], v = sm[i + h];
            sm[i] = u + v;
            sm[i + h] = u - v;
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (base + i < total) out[base + i] = sm[i];
}

// One radix-2 stage on bit b (>= kSegLog2) of every row, in place.
__global__ void __launch_bounds__(256) fwht_f64_stage_kernel(double* __restrict__ x, int64_t pairs, int b)
{
    const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p >= pairs) return;
    const int64_t h = int64_t(1) << b;
    const int64_t i = ((p >> b) << (b + 1)) | (p & (h - 1));
    const double u = x[i], v = x[i + h];
    x[i] = u + v;
    x[i + h] = u - v;
}

int launch_fwht_f64(const double* in, double* out, int64_t rows, int64_t D, cudaStream_t stream)
{
    static unsigned char smem_ok[64] = {};
    const int k = ilog2(D);
    if (k > kMaxLog2Dmulti) return fail(WHVI_E_SHAPE, "fwht_f64: D = %lld exceeds the limit 2^%d", (long long)D, kMaxLog2Dmulti);
    const int64_t total = rows * D;
    int seg_k = k < kSegLog2 ? (k < 8 ? 8 : k) : kSegLog2;  // short rows: 256 doubles (several rows) per CTA
    const int64_t ctas = (total + (int64_t(1) << seg_k) - 1) >> seg_k;
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "fwht_f64: grid too large");
    const size_t smem = sizeof(double) << seg_k;
    if (int rc = ensure_smem(fwht_f64_segment_kernel, smem, smem_ok)) return rc;
    fwht_f64_segment_kernel<<<static_cast<unsigned>(ctas), 256, smem, stream>>>(in, out, total, seg_k, k);
    if (int rc = check_launch("fwht_f64_segment_kernel")) return rc;
    const int64_t pairs = total / 2;
    for (int b = kSegLog2; b < k; ++b) {
        fwht_f64_stage_kernel<<<static_cast<unsigned>((pairs + 255) / 256), 256, 0, stream>>>(out, pairs, b);
        if (int rc = check_launch("fwht_f64_stage_kernel")) return rc;
    }
    return WHVI_OK;
}

}  // namespace whvi
