// Tensor-memory (TMEM) as register-spill space: throughput of tcgen05.ld / tcgen05.st (32x32b) per SM on
// B200 and whether they share a data path with shared-memory loads/stores or the FP32 pipe.  Decides
// whether the fused backward can keep its accumulators / parked streams in TMEM (round 2).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_tmem.bin tools/microbench_tmem.cu
//   ./tools/microbench_tmem.bin
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int ITERS = 2048;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

#define R32(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), \
                  "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]), \
                  "=r"(v[o + 16]), "=r"(v[o + 17]), "=r"(v[o + 18]), "=r"(v[o + 19]), "=r"(v[o + 20]), "=r"(v[o + 21]), "=r"(v[o + 22]), "=r"(v[o + 23]), \
                  "=r"(v[o + 24]), "=r"(v[o + 25]), "=r"(v[o + 26]), "=r"(v[o + 27]), "=r"(v[o + 28]), "=r"(v[o + 29]), "=r"(v[o + 30]), "=r"(v[o + 31])
#define W32(v, o) "r"(v[o + 0]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7]), \
                  "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15]), \
                  "r"(v[o + 16]), "r"(v[o + 17]), "r"(v[o + 18]), "r"(v[o + 19]), "r"(v[o + 20]), "r"(v[o + 21]), "r"(v[o + 22]), "r"(v[o + 23]), \
                  "r"(v[o + 24]), "r"(v[o + 25]), "r"(v[o + 26]), "r"(v[o + 27]), "r"(v[o + 28]), "r"(v[o + 29]), "r"(v[o + 30]), "r"(v[o + 31])

__device__ __forceinline__ void tmem_ld32(uint32_t (&v)[32], uint32_t taddr)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : R32(v, 0)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(const uint32_t (&v)[32], uint32_t taddr)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), W32(v, 0)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE bits: 1 = tcgen05.ld x32 (2 per iter), 2 = tcgen05.st x32 (2 per iter), 4 = LDS.128 x8 per iter,
//            8 = STS.32 x32 per iter, 16 = 64 FFMA2-free FADD2 per iter (FP32 pipe)
template <int MODE>
__global__ void __launch_bounds__(512, 1) tmem_kernel(float* out, long long* cycles, int check)
{
    __shared__ uint32_t tmem_base_smem;
    extern __shared__ float4 dyn4[];
    float* sm = reinterpret_cast<float*>(dyn4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = float(i);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    const int per_quadrant = nwarps / 4;              // warps sharing one 32-lane quadrant
    const int cols = 512 / per_quadrant;              // columns owned by this warp
    const uint32_t taddr = tmem_base + (uint32_t(32 * (warp & 3)) << 16) + uint32_t((warp >> 2) * cols);

    uint32_t v[32], w[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = threadIdx.x * 32 + i, w[i] = i;
    // fill this warp's columns so that loads are defined
    for (int c = 0; c < cols; c += 32) tmem_st32(v, taddr + c);
    tmem_wait_st();
    if (check) {  // functional check: what a thread stored is what it loads back
        tmem_ld32(w, taddr + (cols - 32));
        tmem_wait_ld();
        int bad = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) bad += (w[i] != v[i]);
        if (bad) printf("TMEM readback mismatch thread %d (%d words)\n", threadIdx.x, bad);
    }
    float f[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = float(threadIdx.x + i);
    float4 q = make_float4(0, 0, 0, 0);
    int idx = threadIdx.x & 255;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const uint32_t c0 = uint32_t(it * 64) % uint32_t(cols);   // cols is a multiple of 64
        if (MODE & 1) {
            tmem_ld32(v, taddr + c0);
            tmem_ld32(w, taddr + c0 + 32);
        }
        if (MODE & 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 x = dyn4[(idx + 32 * j) & 2047];
                q.x += x.x, q.y += x.y, q.z += x.z, q.w += x.w;
            }
            idx = (idx + 8) & 255;
        }
        if (MODE & 8) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sm[((warp * 32 + j) * 32 + lane) & 8191] = f[j & 15];
        }
        if (MODE & 16) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float2 a = make_float2(f[i], f[i + 1]), c = make_float2(f[i + 2], f[i + 3]);
                    const float2 s = __fadd2_rn(a, c), d = __fadd2_rn(a, make_float2(-c.x, -c.y));
                    f[i] = s.x * 0.5f, f[i + 1] = s.y * 0.5f, f[i + 2] = d.x * 0.5f, f[i + 3] = d.y * 0.5f;
                }
            }
        }
        if (MODE & 1) tmem_wait_ld();
        if (MODE & 2) {
            tmem_st32(v, taddr + c0);
            tmem_st32(w, taddr + c0 + 32);
            tmem_wait_st();
        }
    }
    const long long t1 = clock64();
    float acc = q.x + q.y + q.z + q.w;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += f[i];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc += __uint_as_float(v[i] ^ w[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

template <int MODE>
static void run(const char* name, int threads, float* out, long long* cyc_d)
{
    const int blocks = 148;
    tmem_kernel<MODE><<<blocks, threads, 32768>>>(out, cyc_d, 1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { std::printf("%-44s FAILED: %s\n", name, cudaGetErrorString(e)); return; }
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    tmem_kernel<MODE><<<blocks, threads, 32768>>>(out, cyc_d, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long cyc[148];
    cudaMemcpy(cyc, cyc_d, sizeof(cyc), cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < blocks; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
    const double per_iter = double(mx) / ITERS;
    const int warps = threads / 32;
    std::printf("%-44s %2d warps/SM  %8.3f ms  %9.1f clk/iter/SM", name, warps, ms, per_iter);
    if (MODE & 1) std::printf("  LDTM %7.1f B/clk/SM", warps * 2 * 4096.0 / per_iter);
    if (MODE & 2) std::printf("  STTM %7.1f B/clk/SM", warps * 2 * 4096.0 / per_iter);
    if (MODE & 4) std::printf("  LDS %6.1f B/clk/SM", warps * 8 * 512.0 / per_iter);
    if (MODE & 8) std::printf("  STS %6.1f B/clk/SM", warps * 32 * 128.0 / per_iter);
    if (MODE & 16) std::printf("  FP32 %6.1f lane-ops/clk/SM", warps * 32 * (64.0 + 64.0) / per_iter);
    std::printf("\n");
}

int main()
{
    float* out; cudaMalloc(&out, 148 * 512 * sizeof(float));
    long long* cyc; cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int threads : {256, 512}) {
        run<1>("tcgen05.ld 32x32b.x32", threads, out, cyc);
        run<2>("tcgen05.st 32x32b.x32", threads, out, cyc);
        run<3>("tcgen05.ld + tcgen05.st", threads, out, cyc);
        run<4>("LDS.128 x8", threads, out, cyc);
        run<5>("tcgen05.ld + LDS.128 x8", threads, out, cyc);
        run<8>("STS.32 x32", threads, out, cyc);
        run<10>("tcgen05.st + STS.32 x32", threads, out, cyc);
        run<12>("LDS.128 x8 + STS.32 x32", threads, out, cyc);
        run<15>("tcgen05.ld/st + LDS.128 x8 + STS.32 x32", threads, out, cyc);
        run<16>("FADD2 x32 + FMUL x64", threads, out, cyc);
        run<17>("tcgen05.ld + FP32", threads, out, cyc);
        run<19>("tcgen05.ld/st + FP32", threads, out, cyc);
        run<31>("everything", threads, out, cyc);
    }
    return 0;
}
