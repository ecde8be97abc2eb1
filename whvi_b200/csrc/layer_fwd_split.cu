// Kernel (2) for D = 8192: the fused forward with the row split into two 4096-float halves and tensor memory as the
// stash that makes the split free.
//
// A 64-float-per-thread engine covers 12 index bits with TWO views (one transposition per transform); D = 2^13 has one
// bit more, which cost the general kernel a third view, i.e. a second transposition per transform -- and the shared-memory
// data pipe, not HBM, was its limit (68% of the HBM roofline).  H_8192 = H_2 (x) H_4096: the butterfly on bit 12 is a plain
// elementwise add/subtract between the two halves of the row, and it commutes with everything that acts inside a half.
// So a 64-thread group handles one row as two halves, each thread holding the SAME 64 positions of both:
//     load:   a = s2_lo x_lo, b = s2_hi x_hi;  keep a + b in registers, park a - b in tensor memory
//     t_lo = H_4096(a + b), z_lo = g_lo t_lo  -> parked;   t_hi = H_4096(a - b), z_hi = g_hi t_hi
//     p = z_lo + z_hi (registers), q = z_lo - z_hi (parked);   y_lo = s1_lo H_4096(p),  y_hi = s1_hi H_4096(q)
// Four half-transforms of one transposition each: half the shared-memory wavefronts per element of the three-view kernel;
// the parked halves move over the tcgen05.ld/st path, which runs beside shared memory (profiles/r02_microbench_tmem.txt).
#include "layer_common.cuh"
#include "tmem.cuh"

namespace whvi {

struct FwdSplitArgs {
    const float* x;
    int64_t x_sample_stride;
    const float* g;
    const float* s1;
    const float* s2;
    const float* bias;
    float* y;
    int64_t sample_elems;   // B * D
    int n_samples;
    int ctas_per_sample;
    int iters_per_group;
    int relu_out;
    const float* target;    // HAS_TARGET: (B, D); sum (y - target)^2 -> sq_partials[cta]
    float* sq_partials;
    int spg, pstride, bstride;   // grouped launch (LayerFwdCall)
};

template <int GROUPS, int MINB, bool HAS_BIAS, bool HAS_TARGET, bool FROM_T2, class IO>
__global__ void __launch_bounds__(64 * GROUPS, MINB) layer_fwd_split_kernel(const FwdSplitArgs a)
{
    constexpr int N = 12, C = 6, KT = 12, T = 64, E = 64;
    constexpr int64_t HALF = 4096, ROW = 8192;
    static_assert(GROUPS <= 2, "one TMEM lane quadrant per warp: at most four warps per CTA");
    extern __shared__ float4 smem4[];
    __shared__ uint32_t tmem_base_smem;
    float* smem = reinterpret_cast<float*>(smem4);
    const int group = threadIdx.x / T;
    const uint32_t tid = threadIdx.x % T;
    const int bar = group + 1;
    const uint32_t cmask = uint32_t(HALF) - 1u;
    const int s = blockIdx.x % a.n_samples;   // sample-minor CTA order (layer_fwd.cu)
    const int cta_in_sample = blockIdx.x / a.n_samples;
    const float* __restrict__ gs = a.g + int64_t(s) * ROW;
    const int grp = s / a.spg, sx = s - grp * a.spg;
    const float* __restrict__ s1p = a.s1 + int64_t(grp) * a.pstride;
    const float* __restrict__ s2p = a.s2 + int64_t(grp) * a.pstride;
    const float* __restrict__ biasp = HAS_BIAS ? a.bias + int64_t(grp) * a.bstride : nullptr;
    constexpr size_t SW = scratch_words(N, C);
    float* gt = smem;                                   // g in MID order: [0, SW) low half, [SW, 2 SW) high half
    float* buf = smem + 2 * SW + size_t(group) * SW;    // this group's transposition buffer
    const float relu_floor = a.relu_out ? 0.f : -INFINITY;

    if (threadIdx.x < 32) tm_alloc(&tmem_base_smem, 128);
    gtab_fill<N, C>(gt, gs, T * GROUPS, KT);
    gtab_fill<N, C>(gt + SW, gs + HALF, T * GROUPS, KT);
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tmA = tm_lane_base(tmem_base_smem), tmB = tmA + 64;   // two parked half-rows of 64 floats per thread

    const uint32_t off_f = tile_thread_offset<N, C, V_FIRST>(tid);
    const uint32_t wb_fm = transpose_writer_base<N, C, V_FIRST, V_MID>(tid);
    const uint32_t wb_mf = transpose_writer_base<N, C, V_MID, V_FIRST>(tid);
    const uint32_t gbase = gtab_base<N, C>(tid, KT);

    // first half-transform (FIRST view -> MID view) and the multiply by this half's g
    auto half_in = [&](float (&v)[E], const float* gtab) {
        if constexpr (!FROM_T2) bfly_round<N, C, KT, SEQ2_IN, 0>(v, KT);
        role_sync<T>(bar);   // earlier reads of buf are done
        transpose_write<N, C, V_FIRST, V_MID>(v, buf, wb_fm);
        role_sync<T>(bar);
        transpose_read<C>(v, buf, tid);
        if constexpr (!FROM_T2) bfly_round<N, C, KT, SEQ2_IN, 1>(v, KT);
        gtab_for_each<N, C>(gtab, gbase, [&](auto j_, const float4 w) {
            constexpr int j = decltype(j_)::value;
            scale4(v + 4 * j, w);
        });
    };
    float sq = 0.f;
    // second half-transform (MID view -> FIRST view), s1 / bias / ReLU, store (and the squared error against the target)
    auto half_out = [&](float (&v)[E], const float* __restrict__ s1h, const float* __restrict__ biash, IO* __restrict__ yh,
                        const float* __restrict__ th) {
        bfly_round<N, C, KT, SEQ2_OUT, 0>(v, KT);
        role_sync<T>(bar);
        transpose_write<N, C, V_MID, V_FIRST>(v, buf, wb_mf);
        role_sync<T>(bar);
        transpose_read<C>(v, buf, tid);
        bfly_round<N, C, KT, SEQ2_OUT, 1>(v, KT);
        for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t) {
            constexpr int m = decltype(m_)::value;
            const float4 w = ldg4(s1h + off);
            float4 o = make_float4(v[4 * m] * w.x, v[4 * m + 1] * w.y, v[4 * m + 2] * w.z, v[4 * m + 3] * w.w);
            if constexpr (HAS_BIAS) {
                const float4 b = ldg4(biash + off);
                o.x += b.x, o.y += b.y, o.z += b.z, o.w += b.w;
            }
            o.x = fmaxf(o.x, relu_floor);
            o.y = fmaxf(o.y, relu_floor);
            o.z = fmaxf(o.z, relu_floor);
            o.w = fmaxf(o.w, relu_floor);
            if constexpr (HAS_TARGET) {
                const float4 tg = ldg_stream(th + off);
                const float d0 = o.x - tg.x, d1 = o.y - tg.y, d2 = o.z - tg.z, d3 = o.w - tg.w;
                sq = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, sq))));
            }
            Io<IO>::st4(yh + off, o);
        });
    };

#pragma unroll 1
    for (int it = 0; it < a.iters_per_group; ++it) {
        const int64_t row = (int64_t(cta_in_sample) * a.iters_per_group + it) * GROUPS + group;
        const int64_t e0 = row * ROW;   // element offset inside the sample; rows are whole, so a tile is in or out
        if (e0 >= a.sample_elems) break;
        const IO* __restrict__ xs = reinterpret_cast<const IO*>(a.x) + int64_t(sx) * a.x_sample_stride + e0;
        IO* __restrict__ ys = reinterpret_cast<IO*>(a.y) + int64_t(s) * a.sample_elems + e0;
        if (tid == 0 && it + 1 < a.iters_per_group) {   // pull the group's next row towards L2
            const int64_t e1 = e0 + GROUPS * ROW;
            if (e1 < a.sample_elems) l2_prefetch_bulk(xs + GROUPS * ROW, static_cast<uint32_t>(ROW * sizeof(IO)));
        }

        float v[E];
        // ---- load both halves; the bit-12 butterfly happens here (t2 already has it when FROM_T2)
        if constexpr (!FROM_T2) {
            float d[16];
            static_for<0, E / 4>([&](auto m_) {
                constexpr int m = decltype(m_)::value;
                constexpr uint32_t roff = tile_reg_offset<N, C, V_FIRST>(m);
                const uint32_t off = off_f + roff;
                float4 qa = Io<IO>::ld4(xs + off), qb = Io<IO>::ld4(xs + HALF + off);
                const float4 wa = ldg4(s2p + off), wb = ldg4(s2p + HALF + off);
                float pa[4], pb[4];
                mul4(pa, qa, wa);
                mul4(pb, qb, wb);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[4 * m + i] = pa[i] + pb[i];
                    d[(4 * m + i) % 16] = pa[i] - pb[i];
                }
                if constexpr (m % 4 == 3) tm_st16(d, tmA + 16 * (m / 4));
            });
        } else {
            float d[16];
            static_for<0, E / 4>([&](auto m_) {
                constexpr int m = decltype(m_)::value;
                constexpr uint32_t roff = tile_reg_offset<N, C, V_FIRST>(m);
                const uint32_t off = off_f + roff;
                const float4 qa = Io<IO>::ld4(xs + off), qb = Io<IO>::ld4(xs + HALF + off);
                v[4 * m] = qa.x, v[4 * m + 1] = qa.y, v[4 * m + 2] = qa.z, v[4 * m + 3] = qa.w;
                d[(4 * m) % 16] = qb.x, d[(4 * m) % 16 + 1] = qb.y, d[(4 * m) % 16 + 2] = qb.z, d[(4 * m) % 16 + 3] = qb.w;
                if constexpr (m % 4 == 3) tm_st16(d, tmA + 16 * (m / 4));
            });
        }
        // ---- low half: t_lo, z_lo = g_lo t_lo -> parked in B
        half_in(v, gt);
        tm_st32(v, tmB);
        tm_st32(v + 32, tmB + 32);
        tm_wait_st();   // this thread's stores to A and B are readable by its own loads from here on
        // ---- high half from A
        tm_ld32(v, tmA);
        tm_ld32(v + 32, tmA + 32);
        tm_wait_ld();
        half_in(v, gt + SW);
        // ---- bit-12 butterfly of the second transform: p = z_lo + z_hi stays, q = z_lo - z_hi is parked in A
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float zl[16];
            tm_ld16(zl, tmB + 16 * c);
            tm_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float lo = zl[i], hi = v[16 * c + i];
                v[16 * c + i] = lo + hi;
                zl[i] = lo - hi;
            }
            tm_st16(zl, tmA + 16 * c);
        }
        half_out(v, s1p, biasp, ys, a.target + e0);
        tm_wait_st();
        tm_ld32(v, tmA);
        tm_ld32(v + 32, tmA + 32);
        tm_wait_ld();
        half_out(v, s1p + HALF, biasp + HALF, ys + HALF, a.target + e0 + HALF);
    }

    if constexpr (HAS_TARGET) {   // fixed-order CTA reduction of the squared residuals
        __shared__ float red[4];
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
    tm_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tm_dealloc(tmem_base_smem, 128);
}

int launch_layer_fwd_split(const LayerFwdCall& c, cudaStream_t stream)
{
    static unsigned char smem_ok[10][64] = {};
    constexpr int GROUPS = 2, MINB = 3;
    constexpr int threads = 64 * GROUPS;
    constexpr size_t sw = scratch_words(12, 6);
    constexpr size_t smem = sizeof(float) * sw * (2 + GROUPS);
    const int64_t D = 8192;
    const int64_t rows = c.B;
    const Plan plan = make_plan(c.S, rows, GROUPS, 148 * 12, 4);
    const int64_t ctas = int64_t(plan.ctas_per_sample) * c.S;
    if (c.partials_needed) {
        *c.partials_needed = static_cast<size_t>(ctas);
        return WHVI_OK;
    }
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "layer_fwd: grid too large");
    FwdSplitArgs a{c.x, c.xs, c.g, c.s1, c.s2, c.bias, c.y, c.B * D, static_cast<int>(c.S), plan.ctas_per_sample, plan.iters_per_group,
                   c.relu_out, c.target, c.sq_partials, static_cast<int>(c.S / c.groups), static_cast<int>(c.pstride), static_cast<int>(c.bstride)};
    auto go = [&](auto kernel, int slot) -> int {
        if (int rc = ensure_smem(kernel, smem, smem_ok[slot])) return rc;
        kernel<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(a);
        return check_launch("layer_fwd_split_kernel");
    };
    const bool hb = c.bias != nullptr, ht = c.target != nullptr;
    if (c.bf16) {
        if (ht) return fail(WHVI_E_MODE, "layer_fwd: bf16 activations cannot be combined with a target");
        if (c.from_t2) {
            if (hb) return go(layer_fwd_split_kernel<GROUPS, MINB, true, false, true, __nv_bfloat16>, 6);
            return go(layer_fwd_split_kernel<GROUPS, MINB, false, false, true, __nv_bfloat16>, 7);
        }
        if (hb) return go(layer_fwd_split_kernel<GROUPS, MINB, true, false, false, __nv_bfloat16>, 8);
        return go(layer_fwd_split_kernel<GROUPS, MINB, false, false, false, __nv_bfloat16>, 9);
    }
    if (c.from_t2) {
        if (ht) return fail(WHVI_E_MODE, "layer_fwd: FROM_T2 cannot be combined with a target");
        if (hb) return go(layer_fwd_split_kernel<GROUPS, MINB, true, false, true, float>, 4);
        return go(layer_fwd_split_kernel<GROUPS, MINB, false, false, true, float>, 5);
    }
    if (hb && ht) return go(layer_fwd_split_kernel<GROUPS, MINB, true, true, false, float>, 0);
    if (hb) return go(layer_fwd_split_kernel<GROUPS, MINB, true, false, false, float>, 1);
    if (ht) return go(layer_fwd_split_kernel<GROUPS, MINB, false, true, false, float>, 2);
    return go(layer_fwd_split_kernel<GROUPS, MINB, false, false, false, float>, 3);
}

}  // namespace whvi
