// Shared device helpers of the fused layer kernels (layer_fwd.cu, layer_bwd.cu).
#pragma once
#include "common.cuh"
#include "engine.cuh"
#include "plan.cuh"

namespace whvi {

struct LossArgs {
    const float* x;
    int64_t x_sample_stride;
    const float* g;
    const float* s1;
    const float* s2;
    const float* bias;
    const float* target;
    float* dx;             // NULL: skip
    float* ws;
    float* sq_partials;    // one float per (slab, X warp)
    int64_t sample_elems;
    int n_samples;
    int ctas_per_sample;
    int iters_per_group;
    int k;
    int relu_in;
};

int launch_layer_loss_tm(const LayerLossCall& c, int64_t D, cudaStream_t stream);  // layer_loss_tm.cu

constexpr int SEQ_IN = seq_pack(V_FIRST, V_MID, V_LAST);
constexpr int SEQ_OUT = seq_pack(V_LAST, V_MID2, V_FIRST);
constexpr int SEQ2_IN = seq_pack(V_FIRST, V_MID);
constexpr int SEQ2_OUT = seq_pack(V_MID, V_FIRST);

// KT is the transform length log2(D) when it is a compile-time constant, or -1 when the
// kernel takes it at run time (`k`; one instantiation then serves a whole family of D).

// H over bits [0,k): FIRST -> MID -> LAST for a set of T threads that synchronise on named
// barrier `bar`.  bufA/bufB: two tile-sized shared buffers used in ping-pong so that one
// barrier per transposition suffices (SINGLE: bufA only, one more barrier each).
template <int N, int C, int KT, int T, bool SINGLE>
__device__ __forceinline__ void transform_in(float (&v)[1 << C], float* bufA, float* bufB, uint32_t tid, int bar, int k,
                                             uint32_t wb_fm, uint32_t wb_ml)
{
    bfly_round<N, C, KT, SEQ_IN, 0>(v, k);
    if constexpr (SINGLE) role_sync<T>(bar);  // earlier reads of bufA are done
    transpose_write<N, C, V_FIRST, V_MID>(v, bufA, wb_fm);
    role_sync<T>(bar);
    transpose_read<C>(v, bufA, tid);
    bfly_round<N, C, KT, SEQ_IN, 1>(v, k);
    if constexpr (SINGLE) role_sync<T>(bar);
    transpose_write<N, C, V_MID, V_LAST>(v, SINGLE ? bufA : bufB, wb_ml);
    role_sync<T>(bar);
    transpose_read<C>(v, SINGLE ? bufA : bufB, tid);
    bfly_round<N, C, KT, SEQ_IN, 2>(v, k);
}

// H over bits [0,k): LAST -> MID2 -> FIRST.
template <int N, int C, int KT, int T, bool SINGLE>
__device__ __forceinline__ void transform_out(float (&v)[1 << C], float* bufA, float* bufB, uint32_t tid, int bar, int k,
                                              uint32_t wb_lm, uint32_t wb_mf)
{
    bfly_round<N, C, KT, SEQ_OUT, 0>(v, k);
    if constexpr (SINGLE) role_sync<T>(bar);
    transpose_write<N, C, V_LAST, V_MID2>(v, bufA, wb_lm);
    role_sync<T>(bar);
    transpose_read<C>(v, bufA, tid);
    bfly_round<N, C, KT, SEQ_OUT, 1>(v, k);
    if constexpr (SINGLE) role_sync<T>(bar);
    transpose_write<N, C, V_MID2, V_FIRST>(v, SINGLE ? bufA : bufB, wb_mf);
    role_sync<T>(bar);
    transpose_read<C>(v, SINGLE ? bufA : bufB, tid);
    bfly_round<N, C, KT, SEQ_OUT, 2>(v, k);
}
// Ping-pong note: with two buffers no barrier is needed between consecutive transforms or
// tiles -- every thread reads bufA before the barrier that follows the bufB write, and reads
// bufB before it arrives at the next bufA barrier.

// f(m, element offset inside the tile, coordinate) for the float4s of view V.
template <int N, int C, int V, class F>
__device__ __forceinline__ void for_each_vec(uint32_t toff, uint32_t cmask, F&& f)
{
    static_for<0, (1 << C) / 4>([&](auto m_) {
        constexpr int m = decltype(m_)::value;
        constexpr uint32_t roff = tile_reg_offset<N, C, V>(m);
        f(m_, toff + roff, (toff + roff) & cmask);
    });
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Packed fp32x2 helpers on 4 consecutive registers (FFMA2 / FMUL2: half the issue slots).
// acc[0..3] += q * v[0..3]
__device__ __forceinline__ void fma4(float* acc, const float4& q, const float* v)
{
    const float2 r0 = __ffma2_rn(make_float2(q.x, q.y), make_float2(v[0], v[1]), make_float2(acc[0], acc[1]));
    const float2 r1 = __ffma2_rn(make_float2(q.z, q.w), make_float2(v[2], v[3]), make_float2(acc[2], acc[3]));
    acc[0] = r0.x;
    acc[1] = r0.y;
    acc[2] = r1.x;
    acc[3] = r1.y;
}
// v[0..3] = q * w   /   v[0..3] *= w
__device__ __forceinline__ void mul4(float* v, const float4& q, const float4& w)
{
    const float2 r0 = __fmul2_rn(make_float2(q.x, q.y), make_float2(w.x, w.y));
    const float2 r1 = __fmul2_rn(make_float2(q.z, q.w), make_float2(w.z, w.w));
    v[0] = r0.x;
    v[1] = r0.y;
    v[2] = r1.x;
    v[3] = r1.y;
}
__device__ __forceinline__ void scale4(float* v, const float4& w) { mul4(v, make_float4(v[0], v[1], v[2], v[3]), w); }

// ---- g in shared memory, in the MID view's physical order restricted to coordinates, so
// that the multiply in the middle of the 2-view kernels reads conflict-free float4s --------
template <int N, int C>
__device__ __forceinline__ void gtab_fill(float* gt, const float* __restrict__ gs, int nthreads, int k)
{
    constexpr View mid = view_mid(N, C);
    for (uint32_t c = threadIdx.x; c < (1u << k); c += nthreads) gt[view_phys(mid, c)] = gs[c];
}
template <int N, int C>
__device__ __forceinline__ uint32_t gtab_base(uint32_t tid, int k)
{
    uint32_t base = 0;
    static_for<0, N - C>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        constexpr View mid = view_mid(N, C);
        constexpr int b = mid.bit[C + j];
        constexpr uint32_t col = view_phys(mid, 1u << b);
#if WHVI_PADDED
        base += (((tid >> j) & 1u) && b < k) ? col : 0u;
#else
        base ^= (((tid >> j) & 1u) && b < k) ? col : 0u;
#endif
    });
    return base;
}
// Requires every MID register bit to be a coordinate bit (k > 6 for C = 5, k > 7 for C = 6).
template <int N, int C, class F>
__device__ __forceinline__ void gtab_for_each(const float* gt, uint32_t base, F&& f)
{
    static_for<0, (1 << C) / 4>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        constexpr View mid = view_mid(N, C);
        constexpr uint32_t pr = view_phys(mid, view_reg_logical(mid, 4 * j));
#if WHVI_PADDED
        f(j_, *reinterpret_cast<const float4*>(gt + (base + pr)));
#else
        constexpr uint32_t lo = pr & 0x1Cu, hi = pr & ~0x1Cu;
        f(j_, *reinterpret_cast<const float4*>(gt + ((base ^ lo) + hi)));
#endif
    });
}

}  // namespace whvi
