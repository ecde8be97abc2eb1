// Kernels (2) and (3): the fused WHVILinear forward and backward (PAPER semantics,
// docstring src/weights.py:77:  W = S1 H diag(g) H S2).
//
//   forward, per row (s,b):   y = s1 * H( g_s * H( s2 * x ) ) (+ bias)
//   backward: dt3 = H(s1*dy);  dg_s = sum_b dt3 * t2;  dt1 = H(g_s*dt3)
//                             dx = s2*dt1;  ds2 = sum dt1*x;  ds1 = sum dy*t4;  dbias = sum dy
//                             with t2 = H(s2*x), t4 = H(g_s*t2) recomputed, nothing saved.
//
// The reference materialises W (D x D) through four FWHTs of D x D matrices and a dense
// GEMM per MC sample (src/weights.py:73, :93); here every activation row goes through HBM
// once per pass: 8 B/elt forward, 12 B/elt backward (read x, dy; write dx).
//
// Structure: a tile (one row, or N/D rows when D < N) is held in registers by T threads,
// E floats each; H is the register/shared-memory engine of engine.cuh (views FIRST -> MID
// -> LAST on the way in, LAST -> MID2 -> FIRST on the way out), so x, y, dy, dx, s1, s2 and
// g are all touched as coalesced float4s.
#include "common.cuh"
#include "engine.cuh"

namespace whvi {

constexpr int SEQ_IN = seq_pack(V_FIRST, V_MID, V_LAST);
constexpr int SEQ_OUT = seq_pack(V_LAST, V_MID2, V_FIRST);

// H over bits [0,K): FIRST -> MID -> LAST.  bufA/bufB are two tile-sized shared buffers
// used in ping-pong so that one barrier per transposition suffices.
template <int N, int C, int K, int T, int GROUPS>
__device__ __forceinline__ void transform_in(float (&v)[1 << C], float* bufA, float* bufB, uint32_t tid, int group,
                                             uint32_t wb_fm, uint32_t wb_ml)
{
    bfly_round<N, C, K, SEQ_IN, 0>(v);
    transpose_write<N, C, V_FIRST, V_MID>(v, bufA, wb_fm);
    group_sync<T, GROUPS>(group);
    transpose_read<C>(v, bufA, tid);
    bfly_round<N, C, K, SEQ_IN, 1>(v);
    transpose_write<N, C, V_MID, V_LAST>(v, bufB, wb_ml);
    group_sync<T, GROUPS>(group);
    transpose_read<C>(v, bufB, tid);
    bfly_round<N, C, K, SEQ_IN, 2>(v);
}

// H over bits [0,K): LAST -> MID2 -> FIRST.
template <int N, int C, int K, int T, int GROUPS>
__device__ __forceinline__ void transform_out(float (&v)[1 << C], float* bufA, float* bufB, uint32_t tid, int group,
                                              uint32_t wb_lm, uint32_t wb_mf)
{
    bfly_round<N, C, K, SEQ_OUT, 0>(v);
    transpose_write<N, C, V_LAST, V_MID2>(v, bufA, wb_lm);
    group_sync<T, GROUPS>(group);
    transpose_read<C>(v, bufA, tid);
    bfly_round<N, C, K, SEQ_OUT, 1>(v);
    transpose_write<N, C, V_MID2, V_FIRST>(v, bufB, wb_mf);
    group_sync<T, GROUPS>(group);
    transpose_read<C>(v, bufB, tid);
    bfly_round<N, C, K, SEQ_OUT, 2>(v);
}

// v[4m..4m+3] (op)= vec[(off_m) & (D-1)] for the float4s of view V.
template <int N, int C, int V, int K, class F>
__device__ __forceinline__ void for_each_vec(uint32_t toff, F&& f)
{
    static_for<0, (1 << C) / 4>([&](auto m_) {
        constexpr int m = decltype(m_)::value;
        constexpr uint32_t roff = tile_reg_offset<N, C, V>(m);
        f(m_, toff + roff, (toff + roff) & ((1u << K) - 1u));
    });
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ------------------------------------------------------------------------------ forward
template <int N, int C, int K, int GROUPS>
__global__ void __launch_bounds__((1 << (N - C)) * GROUPS)
layer_fwd_kernel(const float* __restrict__ x, int64_t x_sample_stride, const float* __restrict__ g,
                 const float* __restrict__ s1, const float* __restrict__ s2, const float* __restrict__ bias,
                 float* __restrict__ y, int64_t sample_elems /* B*D */, int ctas_per_sample)
{
    constexpr int T = 1 << (N - C);
    constexpr int E = 1 << C;
    constexpr int64_t TILE = int64_t(1) << N;
    extern __shared__ float4 smem4[];
    const int group = threadIdx.x / T;
    const uint32_t tid = threadIdx.x % T;
    const int s = blockIdx.x / ctas_per_sample;
    const int64_t e0 = (int64_t(blockIdx.x % ctas_per_sample) * GROUPS + group) * TILE;  // within the sample
    if (e0 >= sample_elems) return;
    float* bufA = reinterpret_cast<float*>(smem4) + size_t(group) * 2 * TILE;
    float* bufB = bufA + TILE;
    const float* xs = x + int64_t(s) * x_sample_stride + e0;
    const float* gs = g + (int64_t(s) << K);
    float* ys = y + int64_t(s) * sample_elems + e0;
    const int64_t left = sample_elems - e0;  // valid elements in this tile (>= TILE except at the sample's tail)

    const uint32_t off_f = tile_thread_offset<N, C, V_FIRST>(tid);
    const uint32_t off_l = tile_thread_offset<N, C, V_LAST>(tid);

    float v[E];
    for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
        constexpr int m = decltype(m_)::value;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (off < left) q = ldg_stream(xs + off);
        const float4 p = ldg4(s2 + coord);
        v[4 * m + 0] = q.x * p.x;
        v[4 * m + 1] = q.y * p.y;
        v[4 * m + 2] = q.z * p.z;
        v[4 * m + 3] = q.w * p.w;
    });
    transform_in<N, C, K, T, GROUPS>(v, bufA, bufB, tid, group, transpose_writer_base<N, C, V_FIRST, V_MID>(tid),
                                     transpose_writer_base<N, C, V_MID, V_LAST>(tid));
    for_each_vec<N, C, V_LAST, K>(off_l, [&](auto m_, uint32_t, uint32_t coord) {
        constexpr int m = decltype(m_)::value;
        const float4 p = ldg4(gs + coord);
        v[4 * m + 0] *= p.x;
        v[4 * m + 1] *= p.y;
        v[4 * m + 2] *= p.z;
        v[4 * m + 3] *= p.w;
    });
    // no barrier needed between transforms: every thread read bufA before the barrier that
    // followed the bufB write, and reads bufB before it arrives at the next bufA barrier
    transform_out<N, C, K, T, GROUPS>(v, bufA, bufB, tid, group, transpose_writer_base<N, C, V_LAST, V_MID2>(tid),
                                      transpose_writer_base<N, C, V_MID2, V_FIRST>(tid));
    for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
        constexpr int m = decltype(m_)::value;
        if (off < left) {
            const float4 p = ldg4(s1 + coord);
            float4 o = make_float4(v[4 * m] * p.x, v[4 * m + 1] * p.y, v[4 * m + 2] * p.z, v[4 * m + 3] * p.w);
            if (bias != nullptr) {
                const float4 b = ldg4(bias + coord);
                o.x += b.x;
                o.y += b.y;
                o.z += b.z;
                o.w += b.w;
            }
            stg_stream(ys + off, o);
        }
    });
}

// ------------------------------------------------------------------------------ backward
// One CTA owns `tiles_per_cta` consecutive tiles of ONE sample and keeps the partial sums
// of dg (LAST layout) and ds1, ds2, dbias (FIRST layout) in registers; at the end every
// tile group writes its partials to the workspace and a second kernel reduces them in a
// fixed order (deterministic, no atomics).
// workspace layout: [S][ctas_per_sample][GROUPS][4][TILE] floats: 0 = dg, 1 = ds1, 2 = ds2, 3 = dbias
template <int N, int C, int K, int GROUPS, bool WANT_DX, bool WANT_DBIAS>
__global__ void __launch_bounds__((1 << (N - C)) * GROUPS)
layer_bwd_kernel(const float* __restrict__ x, int64_t x_sample_stride, const float* __restrict__ dy,
                 const float* __restrict__ g, const float* __restrict__ s1, const float* __restrict__ s2,
                 float* __restrict__ dx, float* __restrict__ ws, int64_t sample_elems, int ctas_per_sample,
                 int iters_per_group)
{
    constexpr int T = 1 << (N - C);
    constexpr int E = 1 << C;
    constexpr int64_t TILE = int64_t(1) << N;
    extern __shared__ float4 smem4[];
    const int group = threadIdx.x / T;
    const uint32_t tid = threadIdx.x % T;
    const int s = blockIdx.x / ctas_per_sample;
    const int cta_in_sample = blockIdx.x % ctas_per_sample;
    float* bufA = reinterpret_cast<float*>(smem4) + size_t(group) * 2 * TILE;
    float* bufB = bufA + TILE;
    const float* gs = g + (int64_t(s) << K);

    const uint32_t off_f = tile_thread_offset<N, C, V_FIRST>(tid);
    const uint32_t off_l = tile_thread_offset<N, C, V_LAST>(tid);
    const uint32_t wb_fm = transpose_writer_base<N, C, V_FIRST, V_MID>(tid);
    const uint32_t wb_ml = transpose_writer_base<N, C, V_MID, V_LAST>(tid);
    const uint32_t wb_lm = transpose_writer_base<N, C, V_LAST, V_MID2>(tid);
    const uint32_t wb_mf = transpose_writer_base<N, C, V_MID2, V_FIRST>(tid);

    float acc_g[E], acc_1[E], acc_2[E];
    float acc_b[WANT_DBIAS ? E : 1];
#pragma unroll
    for (int i = 0; i < E; ++i) acc_g[i] = acc_1[i] = acc_2[i] = 0.f;
    if constexpr (WANT_DBIAS) {
#pragma unroll
        for (int i = 0; i < E; ++i) acc_b[i] = 0.f;
    }

    for (int it = 0; it < iters_per_group; ++it) {
        const int64_t tile = (int64_t(cta_in_sample) * iters_per_group + it) * GROUPS + group;
        const int64_t e0 = tile * TILE;
        if (e0 >= sample_elems) break;  // uniform per group
        const int64_t left = sample_elems - e0;
        const float* xs = x + int64_t(s) * x_sample_stride + e0;
        const float* dys = dy + int64_t(s) * sample_elems + e0;

        float a[E], b[E];
        // a = t2 = H(s2 * x)
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off < left) q = ldg_stream(xs + off);
            const float4 p = ldg4(s2 + coord);
            a[4 * m + 0] = q.x * p.x;
            a[4 * m + 1] = q.y * p.y;
            a[4 * m + 2] = q.z * p.z;
            a[4 * m + 3] = q.w * p.w;
        });
        transform_in<N, C, K, T, GROUPS>(a, bufA, bufB, tid, group, wb_fm, wb_ml);
        // b = dt3 = H(s1 * dy)
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off < left) q = ldg_stream(dys + off);
            if constexpr (WANT_DBIAS) {
                acc_b[4 * m + 0] += q.x;
                acc_b[4 * m + 1] += q.y;
                acc_b[4 * m + 2] += q.z;
                acc_b[4 * m + 3] += q.w;
            }
            const float4 p = ldg4(s1 + coord);
            b[4 * m + 0] = q.x * p.x;
            b[4 * m + 1] = q.y * p.y;
            b[4 * m + 2] = q.z * p.z;
            b[4 * m + 3] = q.w * p.w;
        });
        transform_in<N, C, K, T, GROUPS>(b, bufA, bufB, tid, group, wb_fm, wb_ml);
        // LAST layout: dg += dt3 * t2 ; a = t3 = g * t2 ; b = dt2 = g * dt3
        for_each_vec<N, C, V_LAST, K>(off_l, [&](auto m_, uint32_t, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            const float4 p = ldg4(gs + coord);
            const float pv[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc_g[4 * m + q] = fmaf(b[4 * m + q], a[4 * m + q], acc_g[4 * m + q]);
                a[4 * m + q] *= pv[q];
                b[4 * m + q] *= pv[q];
            }
        });
        transform_out<N, C, K, T, GROUPS>(a, bufA, bufB, tid, group, wb_lm, wb_mf);  // a = t4
        // ds1 += dy * t4   (dy re-read: an L2 hit, not DRAM traffic)
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t) {
            constexpr int m = decltype(m_)::value;
            if (off < left) {
                const float4 q = ldg4(dys + off);
                acc_1[4 * m + 0] = fmaf(q.x, a[4 * m + 0], acc_1[4 * m + 0]);
                acc_1[4 * m + 1] = fmaf(q.y, a[4 * m + 1], acc_1[4 * m + 1]);
                acc_1[4 * m + 2] = fmaf(q.z, a[4 * m + 2], acc_1[4 * m + 2]);
                acc_1[4 * m + 3] = fmaf(q.w, a[4 * m + 3], acc_1[4 * m + 3]);
            }
        });
        transform_out<N, C, K, T, GROUPS>(b, bufA, bufB, tid, group, wb_lm, wb_mf);  // b = dt1
        // ds2 += dt1 * x ; dx = s2 * dt1
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            if (off < left) {
                const float4 q = ldg4(xs + off);
                acc_2[4 * m + 0] = fmaf(q.x, b[4 * m + 0], acc_2[4 * m + 0]);
                acc_2[4 * m + 1] = fmaf(q.y, b[4 * m + 1], acc_2[4 * m + 1]);
                acc_2[4 * m + 2] = fmaf(q.z, b[4 * m + 2], acc_2[4 * m + 2]);
                acc_2[4 * m + 3] = fmaf(q.w, b[4 * m + 3], acc_2[4 * m + 3]);
                if constexpr (WANT_DX) {
                    const float4 p = ldg4(s2 + coord);
                    stg_stream(dx + int64_t(s) * sample_elems + e0 + off,
                               make_float4(b[4 * m] * p.x, b[4 * m + 1] * p.y, b[4 * m + 2] * p.z, b[4 * m + 3] * p.w));
                }
            }
        });
    }

    // partial sums -> workspace (coalesced float4, each group its own slab)
    float* slab = ws + ((int64_t(blockIdx.x) * GROUPS + group) * 4) * TILE;
    for_each_vec<N, C, V_LAST, K>(off_l, [&](auto m_, uint32_t off, uint32_t) {
        constexpr int m = decltype(m_)::value;
        *reinterpret_cast<float4*>(slab + off) = make_float4(acc_g[4 * m], acc_g[4 * m + 1], acc_g[4 * m + 2], acc_g[4 * m + 3]);
    });
    for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t) {
        constexpr int m = decltype(m_)::value;
        *reinterpret_cast<float4*>(slab + TILE + off) = make_float4(acc_1[4 * m], acc_1[4 * m + 1], acc_1[4 * m + 2], acc_1[4 * m + 3]);
        *reinterpret_cast<float4*>(slab + 2 * TILE + off) = make_float4(acc_2[4 * m], acc_2[4 * m + 1], acc_2[4 * m + 2], acc_2[4 * m + 3]);
        if constexpr (WANT_DBIAS)
            *reinterpret_cast<float4*>(slab + 3 * TILE + off) = make_float4(acc_b[4 * m], acc_b[4 * m + 1], acc_b[4 * m + 2], acc_b[4 * m + 3]);
    });
}

// Second stage: fixed-order sums of the per-group slabs.
//   dg[s, i]  = sum over slabs of sample s, over the N/D row replicas inside a slab
//   ds1/ds2/dbias[i] = the same over ALL slabs
// grid.x covers D, grid.y = S + 1 (y == S handles the three sample-independent vectors).
__global__ void layer_bwd_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dg, float* __restrict__ ds1,
                                        float* __restrict__ ds2, float* __restrict__ dbias, int S, int slabs_per_sample,
                                        int64_t tile, int D)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= D) return;
    const int reps = static_cast<int>(tile / D);
    if (static_cast<int>(blockIdx.y) < S) {
        const int s = blockIdx.y;
        float acc = 0.f;
        for (int p = 0; p < slabs_per_sample; ++p) {
            const float* slab = ws + ((int64_t(s) * slabs_per_sample + p) * 4) * tile;
            for (int r = 0; r < reps; ++r) acc += slab[int64_t(r) * D + i];
        }
        dg[int64_t(s) * D + i] = acc;
    } else {
        float a1 = 0.f, a2 = 0.f, ab = 0.f;
        const int64_t slabs = int64_t(S) * slabs_per_sample;
        for (int64_t p = 0; p < slabs; ++p) {
            const float* slab = ws + (p * 4) * tile;
            for (int r = 0; r < reps; ++r) {
                a1 += slab[tile + int64_t(r) * D + i];
                a2 += slab[2 * tile + int64_t(r) * D + i];
                if (dbias) ab += slab[3 * tile + int64_t(r) * D + i];
            }
        }
        ds1[i] = a1;
        ds2[i] = a2;
        if (dbias) dbias[i] = ab;
    }
}

// ------------------------------------------------------------------------------ launchers
struct BwdPlan {
    int ctas_per_sample;
    int iters_per_group;
    int groups;
    int64_t tile;
};

template <int N, int GROUPS>
static BwdPlan make_bwd_plan(int64_t S, int64_t B, int64_t D)
{
    const int64_t tile = int64_t(1) << N;
    const int64_t tiles_per_sample = (B * D + tile - 1) / tile;
    // aim for ~8 CTAs per SM worth of work items overall, but at least 8 tiles per group so
    // that the partial-sum epilogue stays negligible
    const int64_t target_ctas = 148 * 4;
    int64_t ctas_per_sample = (target_ctas + S - 1) / S;
    const int64_t max_ctas = (tiles_per_sample + GROUPS - 1) / GROUPS;
    if (ctas_per_sample > max_ctas) ctas_per_sample = max_ctas;
    if (ctas_per_sample < 1) ctas_per_sample = 1;
    int64_t iters = (tiles_per_sample + ctas_per_sample * GROUPS - 1) / (ctas_per_sample * GROUPS);
    ctas_per_sample = (tiles_per_sample + iters * GROUPS - 1) / (iters * GROUPS);
    return BwdPlan{static_cast<int>(ctas_per_sample), static_cast<int>(iters), GROUPS, tile};
}

template <int N, int C, int K, int GROUPS>
static int launch_fwd_cfg(const float* x, int64_t xs, const float* g, const float* s1, const float* s2, const float* bias,
                          float* y, int64_t S, int64_t B, cudaStream_t stream)
{
    static unsigned char smem_ok[64] = {};
    constexpr int threads = (1 << (N - C)) * GROUPS;
    constexpr size_t smem = sizeof(float) * (size_t(2) << N) * GROUPS;
    auto kernel = layer_fwd_kernel<N, C, K, GROUPS>;
    if (int rc = ensure_smem(kernel, smem, smem_ok)) return rc;
    const int64_t D = int64_t(1) << K;
    const int64_t tile = int64_t(1) << N;
    const int64_t tiles_per_sample = (B * D + tile - 1) / tile;
    const int64_t ctas_per_sample = (tiles_per_sample + GROUPS - 1) / GROUPS;
    const int64_t ctas = ctas_per_sample * S;
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "layer_fwd: grid too large");
    kernel<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(x, xs, g, s1, s2, bias, y, B * D,
                                                                  static_cast<int>(ctas_per_sample));
    return check_launch("layer_fwd_kernel");
}

template <int N, int C, int K, int GROUPS>
static int launch_bwd_cfg(const float* x, int64_t xs, const float* dy, const float* g, const float* s1, const float* s2,
                          float* dx, float* dg, float* ds1, float* ds2, float* dbias, float* ws, size_t ws_bytes,
                          int64_t S, int64_t B, cudaStream_t stream, size_t* need_only)
{
    static unsigned char smem_ok[4][64] = {};
    constexpr int threads = (1 << (N - C)) * GROUPS;
    constexpr size_t smem = sizeof(float) * (size_t(2) << N) * GROUPS;
    const int64_t D = int64_t(1) << K;
    const BwdPlan plan = make_bwd_plan<N, GROUPS>(S, B, D);
    const size_t need = sizeof(float) * size_t(S) * plan.ctas_per_sample * GROUPS * 4 * plan.tile;
    if (need_only) {
        *need_only = need;
        return WHVI_OK;
    }
    if (ws == nullptr || ws_bytes < need)
        return fail(WHVI_E_WORKSPACE, "layer_bwd: workspace of %zu bytes needed, %zu given", need, ws_bytes);
    const int64_t ctas = int64_t(plan.ctas_per_sample) * S;
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "layer_bwd: grid too large");
    const bool want_dx = dx != nullptr, want_db = dbias != nullptr;
    auto go = [&](auto kernel, int slot) -> int {
        if (int rc = ensure_smem(kernel, smem, smem_ok[slot])) return rc;
        kernel<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(x, xs, dy, g, s1, s2, dx, ws, B * D,
                                                                      plan.ctas_per_sample, plan.iters_per_group);
        return check_launch("layer_bwd_kernel");
    };
    int rc;
    if (want_dx && want_db) rc = go(layer_bwd_kernel<N, C, K, GROUPS, true, true>, 0);
    else if (want_dx) rc = go(layer_bwd_kernel<N, C, K, GROUPS, true, false>, 1);
    else if (want_db) rc = go(layer_bwd_kernel<N, C, K, GROUPS, false, true>, 2);
    else rc = go(layer_bwd_kernel<N, C, K, GROUPS, false, false>, 3);
    if (rc) return rc;
    const int rthreads = 128;
    dim3 rgrid(static_cast<unsigned>((D + rthreads - 1) / rthreads), static_cast<unsigned>(S + 1));
    layer_bwd_reduce_kernel<<<rgrid, rthreads, 0, stream>>>(ws, dg, ds1, ds2, dbias, static_cast<int>(S),
                                                            plan.ctas_per_sample * GROUPS, plan.tile, static_cast<int>(D));
    return check_launch("layer_bwd_reduce_kernel");
}

#define WHVI_LAYER_DISPATCH(K_, CALL)                          \
    switch (K_) {                                              \
    case 2: return CALL(10, 5, 2, 4);                          \
    case 3: return CALL(10, 5, 3, 4);                          \
    case 4: return CALL(10, 5, 4, 4);                          \
    case 5: return CALL(10, 5, 5, 4);                          \
    case 6: return CALL(10, 5, 6, 4);                          \
    case 7: return CALL(10, 5, 7, 4);                          \
    case 8: return CALL(10, 5, 8, 4);                          \
    case 9: return CALL(10, 5, 9, 4);                          \
    case 10: return CALL(10, 5, 10, 4);                        \
    case 11: return CALL(11, 5, 11, 2);                        \
    case 12: return CALL(12, 5, 12, 1);                        \
    case 13: return CALL(13, 5, 13, 1);                        \
    default: break;                                            \
    }

int launch_layer_fwd(const float* x, int64_t xs, const float* g, const float* s1, const float* s2, const float* bias,
                     float* y, int64_t S, int64_t B, int64_t D, cudaStream_t stream)
{
    const int K = ilog2(D);
#define CALL_FWD(N_, C_, K__, G_) launch_fwd_cfg<N_, C_, K__, G_>(x, xs, g, s1, s2, bias, y, S, B, stream)
    WHVI_LAYER_DISPATCH(K, CALL_FWD)
#undef CALL_FWD
    return fail(WHVI_E_SHAPE, "layer_fwd: D = %lld unsupported (4 <= D <= 8192)", (long long)D);
}

int launch_layer_bwd(const float* x, int64_t xs, const float* dy, const float* g, const float* s1, const float* s2,
                     float* dx, float* dg, float* ds1, float* ds2, float* dbias, float* ws, size_t ws_bytes, int64_t S,
                     int64_t B, int64_t D, cudaStream_t stream, size_t* need_only)
{
    const int K = ilog2(D);
#define CALL_BWD(N_, C_, K__, G_) \
    launch_bwd_cfg<N_, C_, K__, G_>(x, xs, dy, g, s1, s2, dx, dg, ds1, ds2, dbias, ws, ws_bytes, S, B, stream, need_only)
    WHVI_LAYER_DISPATCH(K, CALL_BWD)
#undef CALL_BWD
    return fail(WHVI_E_SHAPE, "layer_bwd: D = %lld unsupported (4 <= D <= 8192)", (long long)D);
}

}  // namespace whvi
