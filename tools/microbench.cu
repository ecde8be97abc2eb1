// Pipe-throughput microbenchmarks that decide the FWHT engine's design on B200:
//   shuffle vs shared-memory (do they share the LSU data path?), FADD vs FADD2 issue rate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu && ./microbench
#include <cuda_runtime.h>
#include <cstdio>

constexpr int ITERS = 4096;

template <int MODE>  // 0: shfl only, 1: lds only (float), 2: shfl + lds, 3: lds.128 only, 4: sts.32 + lds.128, 5: shfl + lds128
__global__ void __launch_bounds__(256) lsu_kernel(float* out)
{
    __shared__ float4 sm4[256 * 2];
    float* sm = reinterpret_cast<float*>(sm4);
    const int t = threadIdx.x;
    sm[t] = t; sm[t + 256] = 2 * t; sm[t + 512] = 3 * t; sm[t + 768] = t;
    sm[t + 1024] = t; sm[t + 1280] = t; sm[t + 1536] = t; sm[t + 1792] = t;
    __syncthreads();
    float a0 = t, a1 = t + 1, a2 = t + 2, a3 = t + 3;
    float4 q = make_float4(0, 0, 0, 0);
    int idx = t;
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0 || MODE == 2 || MODE == 5) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, 1);
            a1 += __shfl_xor_sync(0xffffffffu, a1, 2);
            a2 += __shfl_xor_sync(0xffffffffu, a2, 4);
            a3 += __shfl_xor_sync(0xffffffffu, a3, 8);
        }
        if (MODE == 1 || MODE == 2) {
            a0 += sm[idx]; a1 += sm[idx + 256]; a2 += sm[idx + 512]; a3 += sm[idx + 768];
            idx = (idx + 32) & 255;
        }
        if (MODE == 3 || MODE == 4 || MODE == 5) {
            const float4 v = sm4[idx];
            q.x += v.x; q.y += v.y; q.z += v.z; q.w += v.w;
            idx = (idx + 8) & 255;
        }
        if (MODE == 4) {
            sm[1024 + t] = a0; sm[1280 + t] = a1; sm[1536 + t] = a2; sm[1792 + t] = a3;
            a0 += 1.f;
        }
    }
    out[blockIdx.x * 256 + t] = a0 + a1 + a2 + a3 + q.x + q.y + q.z + q.w;
}

template <int MODE>  // 0: FADD scalar, 1: FADD2 packed
__global__ void __launch_bounds__(256) fadd_kernel(float* out)
{
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            if (MODE == 0) {
                const float a = v[i], b = v[i + 1], c = v[i + 2], d = v[i + 3];
                v[i] = a + c; v[i + 1] = b + d; v[i + 2] = a - c; v[i + 3] = b - d;
            } else {
                const float2 a = make_float2(v[i], v[i + 1]), c = make_float2(v[i + 2], v[i + 3]);
                const float2 s = __fadd2_rn(a, c), d = __fadd2_rn(a, make_float2(-c.x, -c.y));
                v[i] = s.x; v[i + 1] = s.y; v[i + 2] = d.x; v[i + 3] = d.y;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] *= 0.5f;  // keep values bounded (FMUL, counted)
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <class K>
static float run(K k, float* out, int blocks)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<blocks, 256>>>(out);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<<<blocks, 256>>>(out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float));
    const int blocks = 148 * 8;  // 8 CTAs x 256 threads per SM = full occupancy
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char* names[] = {"shfl x4", "lds.32 x4", "shfl x4 + lds.32 x4", "lds.128 x1", "sts.32 x4 + lds.128 x1", "shfl x4 + lds.128 x1"};
    float ms[6];
    ms[0] = run(lsu_kernel<0>, out, blocks); ms[1] = run(lsu_kernel<1>, out, blocks); ms[2] = run(lsu_kernel<2>, out, blocks);
    ms[3] = run(lsu_kernel<3>, out, blocks); ms[4] = run(lsu_kernel<4>, out, blocks); ms[5] = run(lsu_kernel<5>, out, blocks);
    for (int i = 0; i < 6; ++i) {
        // warp-instructions per SM: 64 warps x ITERS per loop body
        const double cyc = ms[i] * 1e-3 * clk * 1e3;  // SM cycles at nominal clock
        std::printf("%-28s %8.3f ms  cycles/iter/SM (64 warps) = %8.1f\n", names[i], ms[i], cyc / ITERS);
    }
    const float f0 = run(fadd_kernel<0>, out, blocks), f1 = run(fadd_kernel<1>, out, blocks);
    std::printf("FADD  scalar butterflies  %8.3f ms  cycles/iter/SM = %8.1f (16 FADD + 16 FMUL per thread-iter)\n", f0, f0 * 1e-3 * clk * 1e3 / ITERS);
    std::printf("FADD2 packed butterflies  %8.3f ms  cycles/iter/SM = %8.1f ( 8 FADD2 + 16 FMUL per thread-iter)\n", f1, f1 * 1e-3 * clk * 1e3 / ITERS);
    return 0;
}
