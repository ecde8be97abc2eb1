// MC predictive moments fused into the layer forward (BASELINE config 5): for every input row b
//   sum_y[b,:]  (+)= sum_s y[s,b,:],   sum_y2[b,:] (+)= sum_s y[s,b,:]^2,   y[s,b,:] = s1 * H(g_s * H(s2 * x[b,:])) + bias
// -- what WHVIRegression.eval_model reduces WHVINetwork.forward's (B, out, S) tensor to (src/networks.py:101-115,
// :131-132; src/likelihoods.py:18-29) -- WITHOUT ever writing a prediction to HBM: a CTA takes one input tile, loops
// over the MC samples of the call and keeps the two running sums ON CHIP.  Per (sample, input) pair the only traffic
// is the re-read of the input tile and of g_s out of L2; HBM sees each input once and 8*D bytes of sums per input.
//
// Where the sums live: a 2^15-float row already fills the CTA's registers (64 floats x 512 threads, 128 registers per
// thread at that block size) and its transposition buffer fills shared memory (128 KB), which is why round 1 wrote
// y to HBM and reduced it in a second pass.  The two running sums go to TENSOR MEMORY (tmem.cuh): 128 columns per
// thread, exactly the 512 columns of an SM for the 16 warps of a 2^15 row; tcgen05.ld/st do not touch the
// shared-memory pipe the transform is bound by.
// s1 and bias are applied once per tile in closed form:  sum y = s1 T1 + S b,  sum y^2 = s1^2 T2 + 2 s1 b T1 + S b^2
// with T1 = sum_s t4, T2 = sum_s t4^2 (t4 = H(g_s t2)), so the sample loop is: multiply by g, transform, two FMAs.
// FROM_T2: x already holds t2 = H(s2 x) (sample-independent, computed once per input with whvi_fwht_f32).
// Otherwise x may also be per-sample, (S, B, D): the last layer of a deeper network.
#include "layer_common.cuh"
#include "tmem.cuh"

namespace whvi {

struct MomArgs {
    const float* x;
    int64_t x_sample_stride;   // 0: one (B, D) block for all samples
    const float* g;            // (S, D)
    const float* s1;
    const float* s2;
    const float* bias;         // may be NULL
    float* sum_y;
    float* sum_y2;             // may be NULL
    int64_t sample_elems;      // B * D
    int64_t tiles;
    int n_samples;
    int k;
    int accumulate;
    int reserve_sms;           // SMs to leave free (a concurrent NCCL kernel needs somewhere to run: these CTAs take whole SMs)
    const float* in_y;         // optional: sums to start from (e.g. a partner rank's partial sums, written into this GPU's
    const float* in_y2;        // memory over NVLink): out = in + this call's sums
};

// CTAs per SM are bounded by tensor-memory columns: 128 per thread, T / 128 warps per lane quadrant
constexpr int moments_ctas_per_sm(int n, int c) { return 512 / (128 * ((1 << (n - c)) / 128)); }

template <int N, int C, int KT, bool FROM_T2>
__global__ void __launch_bounds__(1 << (N - C), moments_ctas_per_sm(N, C)) layer_moments_kernel(const MomArgs a)
{
    constexpr int T = 1 << (N - C);
    constexpr int E = 1 << C;
    constexpr int64_t TILE = int64_t(1) << N;
    constexpr uint32_t TM_COLS = 128u * (T / 128);   // 128 columns per thread, T / 128 warps per lane quadrant
    static_assert(E == 64 && T >= 128 && rounds_needed(N, C, N) == 3, "layer_moments_kernel: three views, 64 floats per thread");
    extern __shared__ float4 smem4[];
    __shared__ uint32_t tmem_base_smem;
    float* buf = reinterpret_cast<float*>(smem4);
    const uint32_t tid = threadIdx.x;
    const int k = KT >= 0 ? KT : a.k;
    const uint32_t cmask = (1u << k) - 1u;
    if (threadIdx.x < 32) tm_alloc(&tmem_base_smem, TM_COLS);
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    const uint32_t tm = tm_lane_base(tmem_base_smem) + 128u * (threadIdx.x >> 7);   // [0,64) T1, [64,128) T2

    const uint32_t off_f = tile_thread_offset<N, C, V_FIRST>(tid);
    const uint32_t off_l = tile_thread_offset<N, C, V_LAST>(tid);
    const uint32_t wb_fm = transpose_writer_base<N, C, V_FIRST, V_MID>(tid);
    const uint32_t wb_ml = transpose_writer_base<N, C, V_MID, V_LAST>(tid);
    const uint32_t wb_lm = transpose_writer_base<N, C, V_LAST, V_MID2>(tid);
    const uint32_t wb_mf = transpose_writer_base<N, C, V_MID2, V_FIRST>(tid);
    const float nS = static_cast<float>(a.n_samples);

#pragma unroll 1
    for (int64_t tile = blockIdx.x; tile < a.tiles; tile += gridDim.x) {
        const int64_t e0 = tile * TILE;
        const int64_t left = a.sample_elems - e0;
        {
            float z[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = 0.f;
            tm_st32(z, tm), tm_st32(z, tm + 32), tm_st32(z, tm + 64), tm_st32(z, tm + 96);
        }
#pragma unroll 1
        for (int s = 0; s < a.n_samples; ++s) {
            const float* __restrict__ xs = a.x + int64_t(s) * a.x_sample_stride + e0;
            const float* __restrict__ gs = a.g + (int64_t(s) << k);
            float v[E];
            if constexpr (FROM_T2) {
                // t2 straight in the LAST view (float4-coalesced too), times g, one transform
                for_each_vec<N, C, V_LAST>(off_l, cmask, [&](auto m_, uint32_t off, uint32_t coord) {
                    constexpr int m = decltype(m_)::value;
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (off < left) q = ldg4(xs + off);   // re-read per sample: L1/L2-resident (one tile per CTA)
                    mul4(v + 4 * m, q, ldg4(gs + coord));
                });
            } else {
                for_each_vec<N, C, V_FIRST>(off_f, cmask, [&](auto m_, uint32_t off, uint32_t coord) {
                    constexpr int m = decltype(m_)::value;
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (off < left) q = a.x_sample_stride ? ldg_stream(xs + off) : ldg4(xs + off);
                    mul4(v + 4 * m, q, ldg4(a.s2 + coord));
                });
                transform_in<N, C, KT, T, true>(v, buf, buf, tid, 1, k, wb_fm, wb_ml);
                for_each_vec<N, C, V_LAST>(off_l, cmask, [&](auto m_, uint32_t, uint32_t coord) {
                    constexpr int m = decltype(m_)::value;
                    scale4(v + 4 * m, ldg4(gs + coord));
                });
            }
            transform_out<N, C, KT, T, true>(v, buf, buf, tid, 1, k, wb_lm, wb_mf);   // v = t4 (FIRST layout)
            tm_wait_st();   // the previous sample's stores to the running sums
#pragma unroll
            for (int c = 0; c < 4; ++c) {   // 16 columns at a time: the whole kernel must fit 128 registers per thread
                float t1[16], t2[16];
                tm_ld16(t1, tm + 16 * c);
                tm_ld16(t2, tm + 64 + 16 * c);
                tm_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    t1[i] += v[16 * c + i];
                    t2[i] = fmaf(v[16 * c + i], v[16 * c + i], t2[i]);
                }
                tm_st16(t1, tm + 16 * c);
                tm_st16(t2, tm + 64 + 16 * c);
            }
        }
        tm_wait_st();
        // closed-form output scale and bias, then (sum_y, sum_y2) for this tile
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            float t1[32], t2[32];
            tm_ld32(t1, tm + 32 * c);
            tm_ld32(t2, tm + 64 + 32 * c);
            tm_wait_ld();
            static_for<0, 8>([&](auto j_) {
                constexpr int j = decltype(j_)::value;
                const uint32_t off = off_f + tile_reg_offset<N, C, V_FIRST>(8 * c + j);
                if (off < left) {
                    const uint32_t coord = off & cmask;
                    const float4 w = ldg4(a.s1 + coord);
                    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (a.bias) b = ldg4(a.bias + coord);
                    const float ws[4] = {w.x, w.y, w.z, w.w}, bs[4] = {b.x, b.y, b.z, b.w};
                    float o1[4], o2[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float m1 = t1[4 * j + q], m2 = t2[4 * j + q];
                        o1[q] = fmaf(ws[q], m1, nS * bs[q]);
                        o2[q] = fmaf(ws[q] * ws[q], m2, fmaf(2.f * ws[q] * bs[q], m1, nS * bs[q] * bs[q]));
                    }
                    float4* py = reinterpret_cast<float4*>(a.sum_y + e0 + off);
                    float4 r1 = make_float4(o1[0], o1[1], o1[2], o1[3]);
                    if (a.in_y) {
                        const float4 old = ldg_stream(a.in_y + e0 + off);
                        r1 = make_float4(old.x + r1.x, old.y + r1.y, old.z + r1.z, old.w + r1.w);
                    }
                    if (a.accumulate) {
                        const float4 old = *py;
                        r1 = make_float4(old.x + r1.x, old.y + r1.y, old.z + r1.z, old.w + r1.w);
                    }
                    *py = r1;
                    if (a.sum_y2) {
                        float4* py2 = reinterpret_cast<float4*>(a.sum_y2 + e0 + off);
                        float4 r2 = make_float4(o2[0], o2[1], o2[2], o2[3]);
                        if (a.in_y2) {
                            const float4 old = ldg_stream(a.in_y2 + e0 + off);
                            r2 = make_float4(old.x + r2.x, old.y + r2.y, old.z + r2.z, old.w + r2.w);
                        }
                        if (a.accumulate) {
                            const float4 old = *py2;
                            r2 = make_float4(old.x + r2.x, old.y + r2.y, old.z + r2.z, old.w + r2.w);
                        }
                        *py2 = r2;
                    }
                }
            });
        }
    }
    tm_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tm_dealloc(tmem_base_smem, TM_COLS);
}

template <int N, int C, int KT>
static int launch_moments_cfg(const MomArgs& a, bool from_t2, cudaStream_t stream)
{
    static unsigned char smem_ok[2][64] = {};
    constexpr int T = 1 << (N - C);
    constexpr size_t smem = sizeof(float) * size_t(scratch_words(N, C));
    constexpr int ctas_per_sm = moments_ctas_per_sm(N, C);
    const int sms = 148 - (a.reserve_sms > 0 && a.reserve_sms < 148 ? a.reserve_sms : 0);
    int64_t grid = a.tiles < int64_t(sms) * ctas_per_sm ? a.tiles : int64_t(sms) * ctas_per_sm;
    auto go = [&](auto kernel, int slot) -> int {
        if (int rc = ensure_smem(kernel, smem, smem_ok[slot])) return rc;
        kernel<<<static_cast<unsigned>(grid), T, smem, stream>>>(a);
        return check_launch("layer_moments_kernel");
    };
    return from_t2 ? go(layer_moments_kernel<N, C, KT, true>, 0) : go(layer_moments_kernel<N, C, KT, false>, 1);
}

int launch_layer_moments(const float* x, int64_t xs, const float* g, const float* s1, const float* s2, const float* bias, float* sum_y,
                         float* sum_y2, int64_t S, int64_t B, int64_t D, int from_t2, int accumulate, int reserve_sms, cudaStream_t stream,
                         const float* in_y, const float* in_y2)
{
    const int k = ilog2(D);
    const int64_t tile = D;   // one row per tile
    MomArgs a{x, xs, g, s1, s2, bias, sum_y, sum_y2, B * D, (B * D + tile - 1) / tile, static_cast<int>(S), k, accumulate, reserve_sms,
              in_y, in_y2};
    if (k == 13) return launch_moments_cfg<13, 6, 13>(a, from_t2, stream);
    if (k == 14) return launch_moments_cfg<14, 6, 14>(a, from_t2, stream);
    if (k == 15) return launch_moments_cfg<15, 6, 15>(a, from_t2, stream);
    return fail(WHVI_E_SHAPE, "layer_moments: D = %lld unsupported (8192 <= D <= 32768)", (long long)D);
}

}  // namespace whvi
