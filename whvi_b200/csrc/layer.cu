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

// ---- two-view transforms (configurations where FIRST + MID cover all K bits) ------------
constexpr int SEQ2_IN = seq_pack(V_FIRST, V_MID);
constexpr int SEQ2_OUT = seq_pack(V_MID, V_FIRST);

// g lives in shared memory in the MID view's physical order (restricted to coordinates),
// so the multiply in the middle of the 2-view kernels reads conflict-free float4s.
template <int N, int C, int K>
__device__ __forceinline__ void gtab_fill(float* gt, const float* __restrict__ gs, int nthreads)
{
    constexpr View mid = view_mid(N, C);
    for (uint32_t c = threadIdx.x; c < (1u << K); c += nthreads) gt[view_phys(mid, c)] = gs[c];
}
template <int N, int C, int K>
__device__ __forceinline__ uint32_t gtab_base(uint32_t tid)
{
    uint32_t base = 0;
    static_for<0, N - C>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        constexpr View mid = view_mid(N, C);
        constexpr int b = mid.bit[C + j];
        constexpr uint32_t col = b < K ? view_phys(mid, 1u << b) : 0u;
        base ^= ((tid >> j) & 1u) ? col : 0u;
    });
    return base;
}
template <int N, int C, int K, class F>
__device__ __forceinline__ void gtab_for_each(const float* gt, uint32_t base, F&& f)
{
    static_for<0, (1 << C) / 4>([&](auto j_) {
        constexpr int j = decltype(j_)::value;
        constexpr View mid = view_mid(N, C);
        static_assert(mid.bit[0] < K && mid.bit[1] < K, "float4 of g must be contiguous coordinates");
        constexpr uint32_t pr = view_phys(mid, view_reg_logical(mid, 4 * j) & ((1u << K) - 1u));
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
    int relu_out;           // y = max(y, 0)
    const float* target;    // optional (B, D): accumulate sum (y - target)^2 into sq_partials[cta]
    float* sq_partials;
};

// ------------------------------------------------------------------------------ forward
// ROUNDS == 3: FIRST -> MID -> LAST, g applied in LAST (global float4 reads), LAST -> MID2 -> FIRST.
// ROUNDS == 2: FIRST -> MID, g applied in MID from the shared-memory table, MID -> FIRST.
// BUFS: tile buffers per group (2 = ping-pong, one barrier per transposition; 1 = in place,
// one more barrier per transposition but half the shared memory).
template <int N, int C, int K, int GROUPS, int ROUNDS, int BUFS, int MINB>
__global__ void __launch_bounds__((1 << (N - C)) * GROUPS, MINB) layer_fwd_kernel(const FwdArgs a)
{
    constexpr int T = 1 << (N - C);
    constexpr int E = 1 << C;
    constexpr int64_t TILE = int64_t(1) << N;
    static_assert(ROUNDS == 3 || rounds_needed(N, C, K) <= 2, "2-view kernel needs FIRST+MID to cover K bits");
    static_assert(ROUNDS == 2 || BUFS == 2, "3-view kernel is written for ping-pong buffers");
    extern __shared__ float4 smem4[];
    float* smem = reinterpret_cast<float*>(smem4);
    const int group = threadIdx.x / T;
    const uint32_t tid = threadIdx.x % T;
    const int s = blockIdx.x / a.ctas_per_sample;
    const int cta_in_sample = blockIdx.x % a.ctas_per_sample;
    const float* gs = a.g + (int64_t(s) << K);
    float* gt = smem;  // ROUNDS == 2 only
    float* bufA = smem + (ROUNDS == 2 ? TILE : 0) + size_t(group) * BUFS * TILE;
    float* bufB = bufA + (BUFS == 2 ? TILE : 0);

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
        gbase = gtab_base<N, C, K>(tid);
        gtab_fill<N, C, K>(gt, gs, T * GROUPS);
        __syncthreads();
    }

    float sq = 0.f;
#pragma unroll 1
    for (int it = 0; it < a.iters_per_group; ++it) {
        const int64_t tile = (int64_t(cta_in_sample) * a.iters_per_group + it) * GROUPS + group;
        const int64_t e0 = tile * TILE;  // element offset inside the sample
        if (e0 >= a.sample_elems) break;  // uniform per group
        const int64_t left = a.sample_elems - e0;
        const float* xs = a.x + int64_t(s) * a.x_sample_stride + e0;
        float* ys = a.y + int64_t(s) * a.sample_elems + e0;

        float v[E];
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off < left) q = ldg_stream(xs + off);
            const float4 p = ldg4(a.s2 + coord);
            v[4 * m + 0] = q.x * p.x;
            v[4 * m + 1] = q.y * p.y;
            v[4 * m + 2] = q.z * p.z;
            v[4 * m + 3] = q.w * p.w;
        });
        if constexpr (ROUNDS == 3) {
            transform_in<N, C, K, T, GROUPS>(v, bufA, bufB, tid, group, wb_fm, wb_ml);
            for_each_vec<N, C, V_LAST, K>(off_l, [&](auto m_, uint32_t, uint32_t coord) {
                constexpr int m = decltype(m_)::value;
                const float4 p = ldg4(gs + coord);
                v[4 * m + 0] *= p.x;
                v[4 * m + 1] *= p.y;
                v[4 * m + 2] *= p.z;
                v[4 * m + 3] *= p.w;
            });
            // no barrier needed between the transforms: every thread read bufA before the
            // barrier that followed the bufB write, and reads bufB before it arrives at the
            // next bufA barrier (the same argument covers consecutive tiles)
            transform_out<N, C, K, T, GROUPS>(v, bufA, bufB, tid, group, wb_lm, wb_mf);
        } else {
            if constexpr (BUFS == 1) group_sync<T, GROUPS>(group);  // previous tile's reads of bufA are done
            bfly_round<N, C, K, SEQ2_IN, 0>(v);
            transpose_write<N, C, V_FIRST, V_MID>(v, bufA, wb_fm);
            group_sync<T, GROUPS>(group);
            transpose_read<C>(v, bufA, tid);
            bfly_round<N, C, K, SEQ2_IN, 1>(v);
            gtab_for_each<N, C, K>(gt, gbase, [&](auto j_, const float4 p) {
                constexpr int j = decltype(j_)::value;
                v[4 * j + 0] *= p.x;
                v[4 * j + 1] *= p.y;
                v[4 * j + 2] *= p.z;
                v[4 * j + 3] *= p.w;
            });
            bfly_round<N, C, K, SEQ2_OUT, 0>(v);
            if constexpr (BUFS == 1) group_sync<T, GROUPS>(group);
            transpose_write<N, C, V_MID, V_FIRST>(v, bufB, wb_mf);
            group_sync<T, GROUPS>(group);
            transpose_read<C>(v, bufB, tid);
            bfly_round<N, C, K, SEQ2_OUT, 1>(v);
        }
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            if (off < left) {
                const float4 p = ldg4(a.s1 + coord);
                float4 o = make_float4(v[4 * m] * p.x, v[4 * m + 1] * p.y, v[4 * m + 2] * p.z, v[4 * m + 3] * p.w);
                if (a.bias != nullptr) {
                    const float4 b = ldg4(a.bias + coord);
                    o.x += b.x;
                    o.y += b.y;
                    o.z += b.z;
                    o.w += b.w;
                }
                if (a.relu_out) {
                    o.x = fmaxf(o.x, 0.f);
                    o.y = fmaxf(o.y, 0.f);
                    o.z = fmaxf(o.z, 0.f);
                    o.w = fmaxf(o.w, 0.f);
                }
                if (a.target != nullptr) {
                    const float4 tg = ldg4(a.target + e0 + off);
                    const float d0 = o.x - tg.x, d1 = o.y - tg.y, d2 = o.z - tg.z, d3 = o.w - tg.w;
                    sq = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sq))));
                }
                stg_stream(ys + off, o);
            }
        });
    }

    if (a.target != nullptr) {  // fixed-order CTA reduction of the squared residuals
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
    const float* dy;       // (S,B,D) upstream gradient, or -- when target != NULL -- the layer's saved output
    const float* g;
    const float* s1;
    const float* s2;
    float* dx;             // NULL: skip
    float* ws;
    int64_t sample_elems;
    int ctas_per_sample;
    int iters_per_group;
    int relu_in;           // x is the output of a fused ReLU: dx *= (x > 0)
    const float* target;   // optional (B,D): dy := coef[0] * (dy_buffer - target)  (fused Gaussian-MNLL gradient)
    const float* coef;     // device scalar, read when target != NULL
};

// One CTA owns `iters_per_group * GROUPS` consecutive tiles of ONE sample and keeps the
// partial sums of dg (LAST layout) and ds1, ds2, dbias (FIRST layout) in registers; at the
// end every tile group writes its partials to the workspace and a second kernel reduces
// them in a fixed order (deterministic, no atomics).
// workspace layout: [S][ctas_per_sample][GROUPS][4][TILE] floats: 0 = dg, 1 = ds1, 2 = ds2, 3 = dbias
template <int N, int C, int K, int GROUPS, bool WANT_DBIAS>
__global__ void __launch_bounds__((1 << (N - C)) * GROUPS) layer_bwd_kernel(const BwdArgs p)
{
    constexpr int T = 1 << (N - C);
    constexpr int E = 1 << C;
    constexpr int64_t TILE = int64_t(1) << N;
    extern __shared__ float4 smem4[];
    const int group = threadIdx.x / T;
    const uint32_t tid = threadIdx.x % T;
    const int s = blockIdx.x / p.ctas_per_sample;
    const int cta_in_sample = blockIdx.x % p.ctas_per_sample;
    float* bufA = reinterpret_cast<float*>(smem4) + size_t(group) * 2 * TILE;
    float* bufB = bufA + TILE;
    const float* gs = p.g + (int64_t(s) << K);
    const bool resid = p.target != nullptr;
    const float coef = resid ? __ldg(p.coef) : 1.f;

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

#pragma unroll 1
    for (int it = 0; it < p.iters_per_group; ++it) {
        const int64_t tile = (int64_t(cta_in_sample) * p.iters_per_group + it) * GROUPS + group;
        const int64_t e0 = tile * TILE;
        if (e0 >= p.sample_elems) break;  // uniform per group
        const int64_t left = p.sample_elems - e0;
        const float* xs = p.x + int64_t(s) * p.x_sample_stride + e0;
        const float* dys = p.dy + int64_t(s) * p.sample_elems + e0;
        const float* tgt = resid ? p.target + e0 : nullptr;

        // upstream gradient of 4 consecutive elements (optionally the fused MNLL residual)
        auto load_dy = [&](uint32_t off, bool stream) -> float4 {
            float4 q = stream ? ldg_stream(dys + off) : ldg4(dys + off);
            if (resid) {
                const float4 tg = ldg4(tgt + off);
                q = make_float4(coef * (q.x - tg.x), coef * (q.y - tg.y), coef * (q.z - tg.z), coef * (q.w - tg.w));
            }
            return q;
        };

        float a[E], b[E];
        // a = t2 = H(s2 * x)
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off < left) q = ldg_stream(xs + off);
            const float4 w = ldg4(p.s2 + coord);
            a[4 * m + 0] = q.x * w.x;
            a[4 * m + 1] = q.y * w.y;
            a[4 * m + 2] = q.z * w.z;
            a[4 * m + 3] = q.w * w.w;
        });
        transform_in<N, C, K, T, GROUPS>(a, bufA, bufB, tid, group, wb_fm, wb_ml);
        // b = dt3 = H(s1 * dy)
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
            if (off < left) q = load_dy(off, true);
            if constexpr (WANT_DBIAS) {
                acc_b[4 * m + 0] += q.x;
                acc_b[4 * m + 1] += q.y;
                acc_b[4 * m + 2] += q.z;
                acc_b[4 * m + 3] += q.w;
            }
            const float4 w = ldg4(p.s1 + coord);
            b[4 * m + 0] = q.x * w.x;
            b[4 * m + 1] = q.y * w.y;
            b[4 * m + 2] = q.z * w.z;
            b[4 * m + 3] = q.w * w.w;
        });
        transform_in<N, C, K, T, GROUPS>(b, bufA, bufB, tid, group, wb_fm, wb_ml);
        // LAST layout: dg += dt3 * t2 ; a = t3 = g * t2 ; b = dt2 = g * dt3
        for_each_vec<N, C, V_LAST, K>(off_l, [&](auto m_, uint32_t, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            const float4 w = ldg4(gs + coord);
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc_g[4 * m + q] = fmaf(b[4 * m + q], a[4 * m + q], acc_g[4 * m + q]);
                a[4 * m + q] *= wv[q];
                b[4 * m + q] *= wv[q];
            }
        });
        transform_out<N, C, K, T, GROUPS>(a, bufA, bufB, tid, group, wb_lm, wb_mf);  // a = t4
        // ds1 += dy * t4   (dy re-read: an L2 hit, not DRAM traffic)
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t) {
            constexpr int m = decltype(m_)::value;
            if (off < left) {
                const float4 q = load_dy(off, false);
                acc_1[4 * m + 0] = fmaf(q.x, a[4 * m + 0], acc_1[4 * m + 0]);
                acc_1[4 * m + 1] = fmaf(q.y, a[4 * m + 1], acc_1[4 * m + 1]);
                acc_1[4 * m + 2] = fmaf(q.z, a[4 * m + 2], acc_1[4 * m + 2]);
                acc_1[4 * m + 3] = fmaf(q.w, a[4 * m + 3], acc_1[4 * m + 3]);
            }
        });
        transform_out<N, C, K, T, GROUPS>(b, bufA, bufB, tid, group, wb_lm, wb_mf);  // b = dt1
        // ds2 += dt1 * x ; dx = s2 * dt1 (masked by x > 0 when x came out of a fused ReLU)
        for_each_vec<N, C, V_FIRST, K>(off_f, [&](auto m_, uint32_t off, uint32_t coord) {
            constexpr int m = decltype(m_)::value;
            if (off < left) {
                const float4 q = ldg4(xs + off);
                acc_2[4 * m + 0] = fmaf(q.x, b[4 * m + 0], acc_2[4 * m + 0]);
                acc_2[4 * m + 1] = fmaf(q.y, b[4 * m + 1], acc_2[4 * m + 1]);
                acc_2[4 * m + 2] = fmaf(q.z, b[4 * m + 2], acc_2[4 * m + 2]);
                acc_2[4 * m + 3] = fmaf(q.w, b[4 * m + 3], acc_2[4 * m + 3]);
                if (p.dx != nullptr) {
                    const float4 w = ldg4(p.s2 + coord);
                    float4 o = make_float4(b[4 * m] * w.x, b[4 * m + 1] * w.y, b[4 * m + 2] * w.z, b[4 * m + 3] * w.w);
                    if (p.relu_in) {
                        o.x = q.x > 0.f ? o.x : 0.f;
                        o.y = q.y > 0.f ? o.y : 0.f;
                        o.z = q.z > 0.f ? o.z : 0.f;
                        o.w = q.w > 0.f ? o.w : 0.f;
                    }
                    stg_stream(p.dx + int64_t(s) * p.sample_elems + e0 + off, o);
                }
            }
        });
    }

    // partial sums -> workspace (coalesced float4, each group its own slab)
    float* slab = p.ws + ((int64_t(blockIdx.x) * GROUPS + group) * 4) * TILE;
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

template <int N, int C, int K, int GROUPS, int ROUNDS, int BUFS, int MINB>
static int launch_fwd_cfg(const LayerFwdCall& c, cudaStream_t stream)
{
    static unsigned char smem_ok[64] = {};
    constexpr int threads = (1 << (N - C)) * GROUPS;
    constexpr size_t tile = size_t(1) << N;
    constexpr size_t smem = sizeof(float) * (tile * BUFS * GROUPS + (ROUNDS == 2 ? tile : 0));
    const int64_t D = int64_t(1) << K;
    const int64_t tiles_per_sample = (c.B * D + int64_t(tile) - 1) / int64_t(tile);
    const Plan plan = make_plan(c.S, tiles_per_sample, GROUPS, 148 * 16, ROUNDS == 2 ? 4 : 1);
    const int64_t ctas = int64_t(plan.ctas_per_sample) * c.S;
    if (c.partials_needed) {
        *c.partials_needed = static_cast<size_t>(ctas);
        return WHVI_OK;
    }
    auto kernel = layer_fwd_kernel<N, C, K, GROUPS, ROUNDS, BUFS, MINB>;
    if (int rc = ensure_smem(kernel, smem, smem_ok)) return rc;
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "layer_fwd: grid too large");
    FwdArgs a{c.x, c.xs, c.g, c.s1, c.s2, c.bias, c.y, c.B * D, plan.ctas_per_sample, plan.iters_per_group,
              c.relu_out, c.target, c.sq_partials};
    kernel<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(a);
    return check_launch("layer_fwd_kernel");
}

template <int N, int C, int K, int GROUPS>
static int launch_bwd_cfg(const LayerBwdCall& c, cudaStream_t stream)
{
    static unsigned char smem_ok[2][64] = {};
    constexpr int threads = (1 << (N - C)) * GROUPS;
    constexpr size_t tile = size_t(1) << N;
    constexpr size_t smem = sizeof(float) * 2 * tile * GROUPS;
    const int64_t D = int64_t(1) << K;
    const int64_t tiles_per_sample = (c.B * D + int64_t(tile) - 1) / int64_t(tile);
    const Plan plan = make_plan(c.S, tiles_per_sample, GROUPS, 148 * 4, 8);
    const size_t need = sizeof(float) * size_t(c.S) * plan.ctas_per_sample * GROUPS * 4 * tile;
    if (c.need_only) {
        *c.need_only = need;
        return WHVI_OK;
    }
    if (c.ws == nullptr || c.ws_bytes < need)
        return fail(WHVI_E_WORKSPACE, "layer_bwd: workspace of %zu bytes needed, %zu given", need, c.ws_bytes);
    const int64_t ctas = int64_t(plan.ctas_per_sample) * c.S;
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "layer_bwd: grid too large");
    BwdArgs a{c.x, c.xs, c.dy, c.g, c.s1, c.s2, c.dx, c.ws, c.B * D, plan.ctas_per_sample, plan.iters_per_group,
              c.relu_in, c.target, c.coef};
    auto go = [&](auto kernel, int slot) -> int {
        if (int rc = ensure_smem(kernel, smem, smem_ok[slot])) return rc;
        kernel<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(a);
        return check_launch("layer_bwd_kernel");
    };
    const int rc = c.dbias ? go(layer_bwd_kernel<N, C, K, GROUPS, true>, 0) : go(layer_bwd_kernel<N, C, K, GROUPS, false>, 1);
    if (rc) return rc;
    const int warps = 8;
    dim3 rgrid(static_cast<unsigned>((D + warps - 1) / warps), static_cast<unsigned>(c.S + 1));
    layer_bwd_reduce_kernel<<<rgrid, warps * 32, 0, stream>>>(c.ws, c.dg, c.ds1, c.ds2, c.dbias, static_cast<int>(c.S),
                                                              plan.ctas_per_sample * GROUPS, int64_t(tile), static_cast<int>(D));
    return check_launch("layer_bwd_reduce_kernel");
}

int launch_layer_fwd(const LayerFwdCall& c, int64_t D, cudaStream_t stream)
{
    switch (ilog2(D)) {
    // D <= 64: three views (the middle multiply reads g from global memory)
    case 2: return launch_fwd_cfg<10, 5, 2, 4, 3, 2, 4>(c, stream);
    case 3: return launch_fwd_cfg<10, 5, 3, 4, 3, 2, 4>(c, stream);
    case 4: return launch_fwd_cfg<10, 5, 4, 4, 3, 2, 4>(c, stream);
    case 5: return launch_fwd_cfg<10, 5, 5, 4, 3, 2, 4>(c, stream);
    case 6: return launch_fwd_cfg<10, 5, 6, 4, 3, 2, 4>(c, stream);
    // 128 <= D <= 1024: two views, one warp per 1024-float tile
    case 7: return launch_fwd_cfg<10, 5, 7, 8, 2, 2, 3>(c, stream);
    case 8: return launch_fwd_cfg<10, 5, 8, 8, 2, 2, 3>(c, stream);
    case 9: return launch_fwd_cfg<10, 5, 9, 8, 2, 2, 3>(c, stream);
    case 10: return launch_fwd_cfg<10, 5, 10, 8, 2, 2, 3>(c, stream);
    // D = 2048, 4096: two views with 64 floats per thread (tile = 4096)
    case 11: return launch_fwd_cfg<12, 6, 11, 4, 2, 1, 2>(c, stream);
    case 12: return launch_fwd_cfg<12, 6, 12, 4, 2, 1, 2>(c, stream);
    case 13: return launch_fwd_cfg<13, 5, 13, 1, 3, 2, 3>(c, stream);
    default: break;
    }
    return fail(WHVI_E_SHAPE, "layer_fwd: D = %lld unsupported (4 <= D <= 8192)", (long long)D);
}

int launch_layer_bwd(const LayerBwdCall& c, int64_t D, cudaStream_t stream)
{
    switch (ilog2(D)) {
    case 2: return launch_bwd_cfg<10, 5, 2, 4>(c, stream);
    case 3: return launch_bwd_cfg<10, 5, 3, 4>(c, stream);
    case 4: return launch_bwd_cfg<10, 5, 4, 4>(c, stream);
    case 5: return launch_bwd_cfg<10, 5, 5, 4>(c, stream);
    case 6: return launch_bwd_cfg<10, 5, 6, 4>(c, stream);
    case 7: return launch_bwd_cfg<10, 5, 7, 4>(c, stream);
    case 8: return launch_bwd_cfg<10, 5, 8, 4>(c, stream);
    case 9: return launch_bwd_cfg<10, 5, 9, 4>(c, stream);
    case 10: return launch_bwd_cfg<10, 5, 10, 4>(c, stream);
    case 11: return launch_bwd_cfg<11, 5, 11, 2>(c, stream);
    case 12: return launch_bwd_cfg<12, 5, 12, 1>(c, stream);
    case 13: return launch_bwd_cfg<13, 5, 13, 1>(c, stream);
    default: break;
    }
    return fail(WHVI_E_SHAPE, "layer_bwd: D = %lld unsupported (4 <= D <= 8192)", (long long)D);
}

}  // namespace whvi
