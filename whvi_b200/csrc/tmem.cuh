// Tensor memory (TMEM, 128 lanes x 512 columns x 32 bit per SM on sm_100a) used as a second register file.
//
// tcgen05.ld / tcgen05.st in the 32x32b shape move 32-bit registers between a thread and "its" TMEM lane
// (lane 32 * (warp % 4) + laneid, any columns).  Measured on B200 (tools/microbench_tmem.cu,
// profiles/r02_microbench_tmem.txt): loads 565-790 B/clk/SM, stores 460-590 B/clk/SM -- 4-6x the 128 B/clk of
// shared memory -- and they overlap almost completely with LDS/STS traffic (LDS+STS+TMEM ld/st together run as
// fast as LDS+STS alone) and mostly with FP32 issue.  The fused backward is bound by the L1/shared-memory data
// pipe (profiles/r01_bwd_notes.md), so every per-thread array that only its owner -- or the thread with the same
// lane in a warp of the same quadrant -- touches is cheaper here than in shared memory and frees registers:
// running sums (ds1, ds2, dg, dbias), per-thread parameter vectors, and the X<->Y role exchange.
//
// Ordering rules used below (PTX ISA, tcgen05 memory consistency):
//  * a thread's own ld after its own st to the same columns: tcgen05.wait::st in between;
//  * another thread's ld: writer  st -> wait::st -> fence::before_thread_sync -> barrier,
//                         reader  barrier -> fence::after_thread_sync -> ld -> wait::ld.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace whvi {

// One warp allocates `cols` (power of two, 32..512) columns for the CTA and publishes the base address.
__device__ __forceinline__ void tm_alloc(uint32_t* base_in_smem, uint32_t cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     static_cast<uint32_t>(__cvta_generic_to_shared(base_in_smem))),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_dealloc(uint32_t base, uint32_t cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// This warp's lane quadrant: address bits [31:16] = lane, [15:0] = column.
__device__ __forceinline__ uint32_t tm_lane_base(uint32_t base) { return base + ((32u * ((threadIdx.x >> 5) & 3u)) << 16); }

// 32 consecutive columns <-> v[0..31] (v: 32 floats that live in registers after unrolling)
__device__ __forceinline__ void tm_ld32(float* v, uint32_t taddr)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
          "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
          "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tm_st32(const float* v, uint32_t taddr)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
          "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]),
          "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]),
          "f"(v[30]), "f"(v[31])
        : "memory");
}
// 16-column flavour (half the register footprint per access)
__device__ __forceinline__ void tm_ld16(float* v, uint32_t taddr)
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tm_st16(const float* v, uint32_t taddr)
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]),
          "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
        : "memory");
}

}  // namespace whvi
