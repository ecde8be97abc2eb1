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
#include <cstdlib>

namespace whvi {

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
        base ^= (((tid >> j) & 1u) && b < k) ? col : 0u;
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
        constexpr uint32_t lo = pr & 0x1Cu, hi = pr & ~0x1Cu;
        f(j_, *reinterpret_cast<const float4*>(gt + ((base ^ lo) + hi)));
    });
}

struct FwdArgs {
    const float* x;
    int64_t x_sample_stride;
    const float* g;
    const float* s1;
    const float* s2;
    const float* bias;
    float* y;
    int64_t sample_elems;  // B * D
    int ctas_per_sample;
    int iters_per_group;
    int k;                  // log2(D)
    int relu_out;           // y = max(y, 0)
    const float* target;    // HAS_TARGET: (B, D); sum (y - target)^2 goes to sq_partials[cta]
    float* sq_partials;
};

// ------------------------------------------------------------------------------ forward
// ROUNDS == 3: FIRST -> MID -> LAST, g applied in LAST (global float4 reads), LAST -> MID2 -> FIRST.
// ROUNDS == 2: FIRST -> MID, g applied in MID from the shared-memory table, MID -> FIRST.
// BUFS: tile buffers per group (2 = ping-pong, 1 = in place with one more barrier per
// transposition but half the shared memory).
// Flags that guard LOADS are template parameters: a run-time branch around a load inside the
// unrolled float4 loops stops the compiler from batching the loads (measured: 2x slower).
template <int N, int C, int KT, int GROUPS, int ROUNDS, int BUFS, int MINB, bool HAS_BIAS, bool HAS_TARGET>
__global__ void __launch_bounds__((1 << (N - C)) * GROUPS, MINB) layer_fwd_kernel(const FwdArgs a)
{
    constexpr int T = 1 << (N - C);
    constexpr int E = 1 << C;
    constexpr int64_t TILE = int64_t(1) << N;
    static_assert(ROUNDS == 3 || rounds_needed(N, C, N) <= 2, "2-view kernel needs FIRST+MID to cover all bits");
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    const int group = threadIdx.x / T;
    const uint32_t tid = threadIdx.x % T;
    const int bar = group + 1;
    const int k = KT >= 0 ? KT : a.k;
    const uint32_t cmask = (1u << k) - 1u;
    const int s = blockIdx.x / a.ctas_per_sample;
    const int cta_in_sample = blockIdx.x % a.ctas_per_sample;
    const float* __restrict__ gs = a.g + (int64_t(s) << k);
    float* gt = smem;  // ROUNDS == 2 only
    float* bufA = smem + (ROUNDS == 2 ? TILE : 0) + size_t(group) * BUFS * TILE;
    float* bufB = bufA + (BUFS == 2 ? TILE : 0);
    const float relu_floor = a.relu_out ? 0.f : -INFINITY;

    const uint32_t off_f = tile_thread_offset<N, C, V_FIRST>(tid);
    const uint32_t wb_fm = transpose_writer_base<N, C, V_FIRST, V_MID>(tid);
    uint32_t off_l = 0, wb_ml = 0, wb_lm = 0, wb_mf = 0, gbase = 0;
    if constexpr (ROUNDS == 3) {
        off_l = tile_thread_offset<N, C, V_LAST>(tid);
        wb_ml = transpose_writer_base<N, C, V_MID, V_LAST>(tid);
        wb_lm = transpose_writer_base<N, C, V_LAST, V_MID2>(tid);
        wb_mf = transpose_writer_base<N, C, V_MID2, V_FIRST>(tid);
    } else {
        wb_mf = transpose_writer_base<N, C, V_MID, V_FIRST>(tid);
        gbase = gtab_base<N, C>(tid, k);
        gtab_fill<N, C>(gt, gs, T * GROUPS, k);
        __syncthreads();
    }

    float sq = 0.f;
#pragma unroll 1
    for (int it = 0; it < a.iters_per_group; ++it) {
        const int64_t tile = (int64_t(cta_in_sample) * a.iters_per_group + it) * GROUPS + group;
        const int64_t e0 = tile * TILE;  // element offset inside the sample
        if (e0 >= a.sample_elems) break;  // uniform per group
        const int64_t left = a.sample_elems - e0;
        const float* __restrict__ xs = a.x + int64_t(s) * a.x_sample_stride + e0;
        float* __restrict__ ys = a.y + int64_t(s) * a.sample_elems + e0;

        float v[E];
        for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off < left) q = ldg_stream(xs + off);
            const float4 w = ldg4(a.s2 + coord);
            v[4 * m + 0] = q.x * w.x;
            v[4 * m + 1] = q.y * w.y;
            v[4 * m + 2] = q.z * w.z;
            v[4 * m + 3] = q.w * w.w;
        });
        if constexpr (ROUNDS == 3) {
            transform_in<N, C, KT, T, BUFS == 1>(v, bufA, bufB, tid, bar, k, wb_fm, wb_ml);
            for_each_vec<N, C, V_LAST>(off_l, cmask, [&](auto m_, uint32_t, uint32_t coord) {
                constexpr int m = decltype(m_)::value;
                const float4 w = ldg4(gs + coord);
                v[4 * m + 0] *= w.x;
                v[4 * m + 1] *= w.y;
                v[4 * m + 2] *= w.z;
                v[4 * m + 3] *= w.w;
            });
            transform_out<N, C, KT, T, BUFS == 1>(v, bufA, bufB, tid, bar, k, wb_lm, wb_mf);
        } else {
            bfly_round<N, C, KT, SEQ2_IN, 0>(v, k);
            if constexpr (BUFS == 1) role_sync<T>(bar);  // previous tile's reads of bufA are done
            transpose_write<N, C, V_FIRST, V_MID>(v, bufA, wb_fm);
            role_sync<T>(bar);
            transpose_read<C>(v, bufA, tid);
            bfly_round<N, C, KT, SEQ2_IN, 1>(v, k);
            gtab_for_each<N, C>(gt, gbase, [&](auto j_, const float4 w) {
                constexpr int j = decltype(j_)::value;
                v[4 * j + 0] *= w.x;
                v[4 * j + 1] *= w.y;
                v[4 * j + 2] *= w.z;
                v[4 * j + 3] *= w.w;
            });
            bfly_round<N, C, KT, SEQ2_OUT, 0>(v, k);
            if constexpr (BUFS == 1) role_sync<T>(bar);
            transpose_write<N, C, V_MID, V_FIRST>(v, bufB, wb_mf);
            role_sync<T>(bar);
            transpose_read<C>(v, bufB, tid);
            bfly_round<N, C, KT, SEQ2_OUT, 1>(v, k);
        }
        for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            const float4 w = ldg4(a.s1 + coord);
            float4 o = make_float4(v[4 * m] * w.x, v[4 * m + 1] * w.y, v[4 * m + 2] * w.z, v[4 * m + 3] * w.w);
            if constexpr (HAS_BIAS) {
                const float4 b = ldg4(a.bias + coord);
                o.x += b.x;
                o.y += b.y;
                o.z += b.z;
                o.w += b.w;
            }
            o.x = fmaxf(o.x, relu_floor);
            o.y = fmaxf(o.y, relu_floor);
            o.z = fmaxf(o.z, relu_floor);
            o.w = fmaxf(o.w, relu_floor);
            if (off < left) {
                if constexpr (HAS_TARGET) {
                    const float4 tg = ldg4(a.target + e0 + off);
                    const float d0 = o.x - tg.x, d1 = o.y - tg.y, d2 = o.z - tg.z, d3 = o.w - tg.w;
                    sq = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sq))));
                }
                stg_stream(ys + off, o);
            }
        });
    }

    if constexpr (HAS_TARGET) {  // fixed-order CTA reduction of the squared residuals
        __shared__ float red[32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
        __syncthreads();
        if (threadIdx.x == 0) {
            float tot = 0.f;
            for (int w = 0; w < (T * GROUPS) / 32; ++w) tot += red[w];
            a.sq_partials[blockIdx.x] = tot;
        }
    }
}

// ------------------------------------------------------------------------------ backward
struct BwdArgs {
    const float* x;
    int64_t x_sample_stride;
    const float* dy;       // (S,B,D) upstream gradient, or -- RESID -- the layer's saved output
    const float* g;
    const float* s1;
    const float* s2;
    float* dx;             // NULL: skip
    float* ws;
    int64_t sample_elems;
    int ctas_per_sample;
    int iters_per_group;
    int k;                 // log2(D)
    int relu_in;           // x is the output of a fused ReLU: dx *= (x > 0)
    const float* target;   // RESID: (B,D); dy := coef[0] * (dy_buffer - target)  (fused Gaussian-MNLL gradient)
    const float* coef;     // RESID: device scalar
};

// Stream-role specialised backward.  Every tile is worked on by a PAIR of thread sets
// running concurrently:
//   X role:  t2 = H(s2*x) -> publishes t2 -> t4 = H(g*t2) -> ds1 += dy*t4 (+ dbias += dy)
//   Y role:  dt3 = H(s1*dy) -> dg += dt3*t2 -> dt1 = H(g*dt3) -> ds2 += dt1*x, dx = s2*dt1
// Each role keeps one stream (32 floats) and its accumulators in registers (<= 128 regs),
// so two 2T-thread CTAs fit per SM and the two streams' load/transform phases overlap.
// One CTA owns `iters_per_group * PAIRS` consecutive tiles of ONE sample; at the end every
// role writes its partial sums to the workspace and a second kernel reduces them in a
// fixed order (deterministic, no atomics).
// workspace layout: [S][ctas_per_sample][PAIRS][4][TILE] floats: 0 = dg, 1 = ds1, 2 = ds2, 3 = dbias
template <int N, int C, int KT, int PAIRS, int MINB, bool WANT_DBIAS, bool RESID>
__global__ void __launch_bounds__((2 << (N - C)) * PAIRS, MINB) layer_bwd_kernel(const BwdArgs p)
{
    constexpr int T = 1 << (N - C);
    constexpr int E = 1 << C;
    constexpr int64_t TILE = int64_t(1) << N;
    extern __shared__ float4 smem4[];
    const int role = threadIdx.x / (T * PAIRS);        // 0 = X, 1 = Y (warp-uniform)
    const int pair = (threadIdx.x % (T * PAIRS)) / T;
    const uint32_t tid = threadIdx.x % T;
    const int k = KT >= 0 ? KT : p.k;
    const uint32_t cmask = (1u << k) - 1u;
    const int s = blockIdx.x / p.ctas_per_sample;
    const int cta_in_sample = blockIdx.x % p.ctas_per_sample;
    // per pair: X scratch (2 tiles), Y scratch (2 tiles), t2 stash (1 tile)
    float* pair_smem = reinterpret_cast<float*>(smem4) + size_t(pair) * 5 * TILE;
    float* bufA = pair_smem + size_t(role) * 2 * TILE;
    float* bufB = bufA + TILE;
    float* stash = pair_smem + 4 * TILE;
    // named barriers (ids 1..15): a warp-sized role needs none for its transpositions
    const int bar_role = 1 + 4 * pair + role;                              // transpositions inside a role
    const int bar_full = T == 32 ? 1 + 2 * pair : 3 + 4 * pair;            // X -> Y: t2 stash written
    const int bar_empty = T == 32 ? 2 + 2 * pair : 4 + 4 * pair;           // Y -> X: t2 stash consumed
    const float* __restrict__ gs = p.g + (int64_t(s) << k);
    float coef = 1.f;
    if constexpr (RESID) coef = __ldg(p.coef);
    const float relu_thr = p.relu_in ? 0.f : -INFINITY;

    const uint32_t off_f = tile_thread_offset<N, C, V_FIRST>(tid);
    const uint32_t off_l = tile_thread_offset<N, C, V_LAST>(tid);
    const uint32_t wb_fm = transpose_writer_base<N, C, V_FIRST, V_MID>(tid);
    const uint32_t wb_ml = transpose_writer_base<N, C, V_MID, V_LAST>(tid);
    const uint32_t wb_lm = transpose_writer_base<N, C, V_LAST, V_MID2>(tid);
    const uint32_t wb_mf = transpose_writer_base<N, C, V_MID2, V_FIRST>(tid);
    float* __restrict__ slab = p.ws + ((int64_t(blockIdx.x) * PAIRS + pair) * 4) * TILE;

    auto tile_of = [&](int it) -> int64_t { return ((int64_t(cta_in_sample) * p.iters_per_group + it) * PAIRS + pair) * TILE; };
    // upstream gradient of 4 consecutive elements (RESID: the fused MNLL residual)
    auto load_dy = [&](const float* dys, const float* tgt, uint32_t off, bool stream) -> float4 {
        float4 q = stream ? ldg_stream(dys + off) : ldg4(dys + off);
        if constexpr (RESID) {
            const float4 tg = ldg4(tgt + off);
            q = make_float4(coef * (q.x - tg.x), coef * (q.y - tg.y), coef * (q.z - tg.z), coef * (q.w - tg.w));
        }
        return q;
    };
    auto prefetch_next = [&](const float* base, int it) {  // next tile of this role's stream -> L2
        const int64_t e1 = tile_of(it + 1);
        if (tid == 0 && it + 1 < p.iters_per_group && e1 < p.sample_elems) {
            const int64_t left1 = p.sample_elems - e1;
            l2_prefetch_bulk(base + e1, static_cast<uint32_t>((left1 < TILE ? left1 : TILE) * sizeof(float)));
        }
    };

    if (role == 0) {
        // ------------------------------------------------------------------ X role
        float acc_1[E];
        float acc_b[WANT_DBIAS ? E : 1];
#pragma unroll
        for (int i = 0; i < E; ++i) acc_1[i] = 0.f;
        if constexpr (WANT_DBIAS) {
#pragma unroll
            for (int i = 0; i < E; ++i) acc_b[i] = 0.f;
        }
#pragma unroll 1
        for (int it = 0; it < p.iters_per_group; ++it) {
            const int64_t e0 = tile_of(it);
            if (e0 >= p.sample_elems) break;
            const int64_t left = p.sample_elems - e0;
            const float* xbase = p.x + int64_t(s) * p.x_sample_stride;
            const float* __restrict__ xs = xbase + e0;
            const float* __restrict__ dys = p.dy + int64_t(s) * p.sample_elems + e0;
            const float* __restrict__ tgt = RESID ? p.target + e0 : nullptr;
            prefetch_next(xbase, it);
            float a[E];
            for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t coord) {
                constexpr int m = decltype(m_)::value;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (off < left) q = ldg_stream(xs + off);
                const float4 w = ldg4(p.s2 + coord);
                a[4 * m + 0] = q.x * w.x;
                a[4 * m + 1] = q.y * w.y;
                a[4 * m + 2] = q.z * w.z;
                a[4 * m + 3] = q.w * w.w;
            });
            transform_in<N, C, KT, T, false>(a, bufA, bufB, tid, bar_role, k, wb_fm, wb_ml);  // a = t2 (LAST layout)
            if (it > 0) bar_wait<2 * T>(bar_empty);  // Y has consumed the previous tile's t2
#pragma unroll
            for (int j = 0; j < E / 4; ++j)
                *reinterpret_cast<float4*>(stash + (tid << C) + ((j ^ swz_of_tid(C, tid)) << 2)) =
                    make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
            __threadfence_block();
            bar_arrive<2 * T>(bar_full);
            for_each_vec<N, C, V_LAST>(off_l, cmask, [&](auto m_, uint32_t, uint32_t coord) {
                constexpr int m = decltype(m_)::value;
                const float4 w = ldg4(gs + coord);
                a[4 * m + 0] *= w.x;
                a[4 * m + 1] *= w.y;
                a[4 * m + 2] *= w.z;
                a[4 * m + 3] *= w.w;
            });
            transform_out<N, C, KT, T, false>(a, bufA, bufB, tid, bar_role, k, wb_lm, wb_mf);  // a = t4 (FIRST layout)
            for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t) {
                constexpr int m = decltype(m_)::value;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (off < left) q = load_dy(dys, tgt, off, false);  // dy re-read: L2, not DRAM
                acc_1[4 * m + 0] = fmaf(q.x, a[4 * m + 0], acc_1[4 * m + 0]);
                acc_1[4 * m + 1] = fmaf(q.y, a[4 * m + 1], acc_1[4 * m + 1]);
                acc_1[4 * m + 2] = fmaf(q.z, a[4 * m + 2], acc_1[4 * m + 2]);
                acc_1[4 * m + 3] = fmaf(q.w, a[4 * m + 3], acc_1[4 * m + 3]);
                if constexpr (WANT_DBIAS) {
                    acc_b[4 * m + 0] += q.x;
                    acc_b[4 * m + 1] += q.y;
                    acc_b[4 * m + 2] += q.z;
                    acc_b[4 * m + 3] += q.w;
                }
            });
        }
        for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t) {
            constexpr int m = decltype(m_)::value;
            *reinterpret_cast<float4*>(slab + TILE + off) = make_float4(acc_1[4 * m], acc_1[4 * m + 1], acc_1[4 * m + 2], acc_1[4 * m + 3]);
            if constexpr (WANT_DBIAS)
                *reinterpret_cast<float4*>(slab + 3 * TILE + off) = make_float4(acc_b[4 * m], acc_b[4 * m + 1], acc_b[4 * m + 2], acc_b[4 * m + 3]);
        });
    } else {
        // ------------------------------------------------------------------ Y role
        float acc_g[E], acc_2[E];
#pragma unroll
        for (int i = 0; i < E; ++i) acc_g[i] = acc_2[i] = 0.f;
#pragma unroll 1
        for (int it = 0; it < p.iters_per_group; ++it) {
            const int64_t e0 = tile_of(it);
            if (e0 >= p.sample_elems) break;
            const int64_t left = p.sample_elems - e0;
            const float* __restrict__ xs = p.x + int64_t(s) * p.x_sample_stride + e0;
            const float* dybase = p.dy + int64_t(s) * p.sample_elems;
            const float* __restrict__ dys = dybase + e0;
            const float* __restrict__ tgt = RESID ? p.target + e0 : nullptr;
            prefetch_next(dybase, it);
            float b[E];
            for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t coord) {
                constexpr int m = decltype(m_)::value;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (off < left) q = load_dy(dys, tgt, off, true);
                const float4 w = ldg4(p.s1 + coord);
                b[4 * m + 0] = q.x * w.x;
                b[4 * m + 1] = q.y * w.y;
                b[4 * m + 2] = q.z * w.z;
                b[4 * m + 3] = q.w * w.w;
            });
            transform_in<N, C, KT, T, false>(b, bufA, bufB, tid, bar_role, k, wb_fm, wb_ml);  // b = dt3 (LAST layout)
            bar_wait<2 * T>(bar_full);  // t2 of this tile is in the stash
            for_each_vec<N, C, V_LAST>(off_l, cmask, [&](auto m_, uint32_t, uint32_t coord) {
                constexpr int m = decltype(m_)::value;
                const float4 t2 = *reinterpret_cast<const float4*>(stash + (tid << C) + ((m ^ swz_of_tid(C, tid)) << 2));
                const float4 w = ldg4(gs + coord);
                acc_g[4 * m + 0] = fmaf(b[4 * m + 0], t2.x, acc_g[4 * m + 0]);
                acc_g[4 * m + 1] = fmaf(b[4 * m + 1], t2.y, acc_g[4 * m + 1]);
                acc_g[4 * m + 2] = fmaf(b[4 * m + 2], t2.z, acc_g[4 * m + 2]);
                acc_g[4 * m + 3] = fmaf(b[4 * m + 3], t2.w, acc_g[4 * m + 3]);
                b[4 * m + 0] *= w.x;
                b[4 * m + 1] *= w.y;
                b[4 * m + 2] *= w.z;
                b[4 * m + 3] *= w.w;
            });
            bar_arrive<2 * T>(bar_empty);  // stash may be overwritten (X waits from its 2nd tile on)
            transform_out<N, C, KT, T, false>(b, bufA, bufB, tid, bar_role, k, wb_lm, wb_mf);  // b = dt1 (FIRST layout)
            const bool want_dx = p.dx != nullptr;
            float* __restrict__ dxs = p.dx + int64_t(s) * p.sample_elems + e0;
            for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t coord) {
                constexpr int m = decltype(m_)::value;
                float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                if (off < left) q = ldg4(xs + off);  // x re-read: L2, not DRAM
                const float4 w = ldg4(p.s2 + coord);
                acc_2[4 * m + 0] = fmaf(q.x, b[4 * m + 0], acc_2[4 * m + 0]);
                acc_2[4 * m + 1] = fmaf(q.y, b[4 * m + 1], acc_2[4 * m + 1]);
                acc_2[4 * m + 2] = fmaf(q.z, b[4 * m + 2], acc_2[4 * m + 2]);
                acc_2[4 * m + 3] = fmaf(q.w, b[4 * m + 3], acc_2[4 * m + 3]);
                const float4 o = make_float4(q.x > relu_thr ? b[4 * m] * w.x : 0.f, q.y > relu_thr ? b[4 * m + 1] * w.y : 0.f,
                                             q.z > relu_thr ? b[4 * m + 2] * w.z : 0.f, q.w > relu_thr ? b[4 * m + 3] * w.w : 0.f);
                if (want_dx && off < left) stg_stream(dxs + off, o);
            });
        }
        for_each_vec<N, C, V_LAST>(off_l, cmask, [&](auto m_, uint32_t off, uint32_t) {
            constexpr int m = decltype(m_)::value;
            *reinterpret_cast<float4*>(slab + off) = make_float4(acc_g[4 * m], acc_g[4 * m + 1], acc_g[4 * m + 2], acc_g[4 * m + 3]);
        });
        for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t) {
            constexpr int m = decltype(m_)::value;
            *reinterpret_cast<float4*>(slab + 2 * TILE + off) = make_float4(acc_2[4 * m], acc_2[4 * m + 1], acc_2[4 * m + 2], acc_2[4 * m + 3]);
        });
    }
}

// Second stage: fixed-order sums of the per-group slabs.
//   dg[s, i]  = sum over slabs of sample s, over the N/D row replicas inside a slab
//   ds1/ds2/dbias[i] = the same over ALL slabs
// One warp per output coordinate: lanes stride over (slab, replica) pairs, then a fixed
// shuffle tree.  blockIdx.y < S: dg of that sample; blockIdx.y == S: ds1, ds2, dbias.
__global__ void __launch_bounds__(256)
layer_bwd_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dg, float* __restrict__ ds1,
                        float* __restrict__ ds2, float* __restrict__ dbias, int S, int slabs_per_sample, int64_t tile, int D)
{
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= D) return;
    const int reps = static_cast<int>(tile / D);
    const unsigned full = 0xffffffffu;
    if (static_cast<int>(blockIdx.y) < S) {
        const int s = blockIdx.y;
        const int64_t terms = int64_t(slabs_per_sample) * reps;
        float acc = 0.f;
        for (int64_t t = lane; t < terms; t += 32) {
            const int64_t slab = int64_t(s) * slabs_per_sample + t / reps;
            acc += ws[(slab * 4) * tile + (t % reps) * D + i];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(full, acc, o);
        if (lane == 0) dg[int64_t(s) * D + i] = acc;
    } else {
        const int64_t terms = int64_t(S) * slabs_per_sample * reps;
        float a1 = 0.f, a2 = 0.f, ab = 0.f;
        for (int64_t t = lane; t < terms; t += 32) {
            const float* base = ws + ((t / reps) * 4) * tile + (t % reps) * D + i;
            a1 += base[tile];
            a2 += base[2 * tile];
            if (dbias) ab += base[3 * tile];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a1 += __shfl_xor_sync(full, a1, o);
            a2 += __shfl_xor_sync(full, a2, o);
            ab += __shfl_xor_sync(full, ab, o);
        }
        if (lane == 0) {
            ds1[i] = a1;
            ds2[i] = a2;
            if (dbias) dbias[i] = ab;
        }
    }
}

// ------------------------------------------------------------------------------ launchers
struct Plan {
    int ctas_per_sample;
    int iters_per_group;
};

// Split the tiles of every sample over CTAs: enough CTAs to fill the chip several times
// over, but each group keeps at least `min_iters` tiles so per-CTA setup/epilogue amortise.
static Plan make_plan(int64_t S, int64_t tiles_per_sample, int groups, int64_t target_ctas, int min_iters)
{
    const int64_t max_ctas = (tiles_per_sample + groups - 1) / groups;
    int64_t ctas = (target_ctas + S - 1) / S;
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    int64_t iters = (tiles_per_sample + ctas * groups - 1) / (ctas * groups);
    if (iters < min_iters) iters = min_iters;
    ctas = (tiles_per_sample + iters * groups - 1) / (iters * groups);
    return Plan{static_cast<int>(ctas), static_cast<int>(iters)};
}

template <int N, int C, int KT, int GROUPS, int ROUNDS, int BUFS, int MINB>
static int launch_fwd_cfg(const LayerFwdCall& c, int k, cudaStream_t stream)
{
    static unsigned char smem_ok[4][64] = {};
    constexpr int threads = (1 << (N - C)) * GROUPS;
    constexpr size_t tile = size_t(1) << N;
    constexpr size_t smem = sizeof(float) * (tile * BUFS * GROUPS + (ROUNDS == 2 ? tile : 0));
    const int64_t D = int64_t(1) << k;
    const int64_t tiles_per_sample = (c.B * D + int64_t(tile) - 1) / int64_t(tile);
    // 2-view kernels amortise the g-table fill over >= 4 tiles per group; 3-view kernels are
    // plain one-tile-per-group grids (the hardware CTA scheduler overlaps their phases)
    const Plan plan = ROUNDS == 2 ? make_plan(c.S, tiles_per_sample, GROUPS, 148 * 16, 4)
                                  : make_plan(c.S, tiles_per_sample, GROUPS, int64_t(1) << 40, 1);
    const int64_t ctas = int64_t(plan.ctas_per_sample) * c.S;
    if (c.partials_needed) {
        *c.partials_needed = static_cast<size_t>(ctas);
        return WHVI_OK;
    }
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "layer_fwd: grid too large");
    FwdArgs a{c.x, c.xs, c.g, c.s1, c.s2, c.bias, c.y, c.B * D, plan.ctas_per_sample, plan.iters_per_group, k,
              c.relu_out, c.target, c.sq_partials};
    auto go = [&](auto kernel, int slot) -> int {
        if (int rc = ensure_smem(kernel, smem, smem_ok[slot])) return rc;
        kernel<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(a);
        return check_launch("layer_fwd_kernel");
    };
    const bool hb = c.bias != nullptr, ht = c.target != nullptr;
    if (hb && ht) return go(layer_fwd_kernel<N, C, KT, GROUPS, ROUNDS, BUFS, MINB, true, true>, 0);
    if (hb) return go(layer_fwd_kernel<N, C, KT, GROUPS, ROUNDS, BUFS, MINB, true, false>, 1);
    if (ht) return go(layer_fwd_kernel<N, C, KT, GROUPS, ROUNDS, BUFS, MINB, false, true>, 2);
    return go(layer_fwd_kernel<N, C, KT, GROUPS, ROUNDS, BUFS, MINB, false, false>, 3);
}

template <int N, int C, int KT, int PAIRS, int MINB>
static int launch_bwd_cfg(const LayerBwdCall& c, int k, cudaStream_t stream)
{
    static unsigned char smem_ok[4][64] = {};
    constexpr int threads = (2 << (N - C)) * PAIRS;
    constexpr size_t tile = size_t(1) << N;
    constexpr size_t smem = sizeof(float) * 5 * tile * PAIRS;
    const int64_t D = int64_t(1) << k;
    const int64_t tiles_per_sample = (c.B * D + int64_t(tile) - 1) / int64_t(tile);
    const Plan plan = make_plan(c.S, tiles_per_sample, PAIRS, 148 * 4, 8);
    const size_t need = sizeof(float) * size_t(c.S) * plan.ctas_per_sample * PAIRS * 4 * tile;
    if (c.need_only) {
        *c.need_only = need;
        return WHVI_OK;
    }
    if (c.ws == nullptr || c.ws_bytes < need)
        return fail(WHVI_E_WORKSPACE, "layer_bwd: workspace of %zu bytes needed, %zu given", need, c.ws_bytes);
    const int64_t ctas = int64_t(plan.ctas_per_sample) * c.S;
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "layer_bwd: grid too large");
    BwdArgs a{c.x, c.xs, c.dy, c.g, c.s1, c.s2, c.dx, c.ws, c.B * D, plan.ctas_per_sample, plan.iters_per_group, k,
              c.relu_in, c.target, c.coef};
    auto go = [&](auto kernel, int slot) -> int {
        if (int rc = ensure_smem(kernel, smem, smem_ok[slot])) return rc;
        kernel<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(a);
        return check_launch("layer_bwd_kernel");
    };
    const bool db = c.dbias != nullptr, rs = c.target != nullptr;
    int rc;
    if (db && rs) rc = go(layer_bwd_kernel<N, C, KT, PAIRS, MINB, true, true>, 0);
    else if (db) rc = go(layer_bwd_kernel<N, C, KT, PAIRS, MINB, true, false>, 1);
    else if (rs) rc = go(layer_bwd_kernel<N, C, KT, PAIRS, MINB, false, true>, 2);
    else rc = go(layer_bwd_kernel<N, C, KT, PAIRS, MINB, false, false>, 3);
    if (rc) return rc;
    const int warps = 8;
    dim3 rgrid(static_cast<unsigned>((D + warps - 1) / warps), static_cast<unsigned>(c.S + 1));
    layer_bwd_reduce_kernel<<<rgrid, warps * 32, 0, stream>>>(c.ws, c.dg, c.ds1, c.ds2, c.dbias, static_cast<int>(c.S),
                                                              plan.ctas_per_sample * PAIRS, int64_t(tile), static_cast<int>(D));
    return check_launch("layer_bwd_reduce_kernel");
}

int launch_layer_fwd(const LayerFwdCall& c, int64_t D, cudaStream_t stream)
{
    const int k = ilog2(D);
    // D <= 64: three views (the middle multiply reads g from global memory)
    if (k >= 2 && k <= 6) return launch_fwd_cfg<10, 5, k_family(2, 6), 4, 3, 2, 4>(c, k, stream);
    // 128 <= D <= 1024: two views, one warp per 1024-float tile (exact-K code for D = 1024:
    // the run-time guards of a family cost ~8% there)
    if (k >= 7 && k <= 9) return launch_fwd_cfg<10, 5, k_family(7, 9), 8, 2, 2, 3>(c, k, stream);
    if (k == 10) return launch_fwd_cfg<10, 5, 10, 8, 2, 2, 3>(c, k, stream);
    // D = 2048, 4096: two views with 64 floats per thread (tile = 4096 = one or two rows)
    if (k == 11) return launch_fwd_cfg<12, 6, 11, 2, 2, 1, 4>(c, k, stream);
    if (k == 12) return launch_fwd_cfg<12, 6, 12, 2, 2, 1, 4>(c, k, stream);
    // D = 8192: three views, 32 floats per thread
    if (k == 13) return launch_fwd_cfg<13, 5, 13, 1, 3, 2, 3>(c, k, stream);
    return fail(WHVI_E_SHAPE, "layer_fwd: D = %lld unsupported (4 <= D <= 8192)", (long long)D);
}

int launch_layer_bwd(const LayerBwdCall& c, int64_t D, cudaStream_t stream)
{
    const int k = ilog2(D);
    if (k >= 2 && k <= 6) return launch_bwd_cfg<10, 5, k_family(2, 6), 4, 2>(c, k, stream);
    if (k >= 7 && k <= 9) return launch_bwd_cfg<10, 5, k_family(7, 9), 4, 2>(c, k, stream);
    if (k == 10) return launch_bwd_cfg<10, 5, 10, 4, 2>(c, k, stream);
    if (k == 11) return launch_bwd_cfg<11, 5, 11, 2, 2>(c, k, stream);
    if (k == 12) return launch_bwd_cfg<12, 5, 12, 1, 2>(c, k, stream);
    if (k == 13) return launch_bwd_cfg<13, 5, 13, 1, 1>(c, k, stream);
    return fail(WHVI_E_SHAPE, "layer_bwd: D = %lld unsupported (4 <= D <= 8192)", (long long)D);
}

}  // namespace whvi
