// Kernel (1): batched in-place-capable FWHT over power-of-two D, fp32.
//
// Replaces fwht_batch1_kernel / fwht_batch2_kernel of the reference
// (src/fwht/cuda/fwht_cuda_kernel.cu:74-146, :35-67) and its host wrapper (:156-181,
// fwht_cuda.cpp:5-14).  Not a port: rows live in registers (32 or 64 floats per thread),
// all butterflies are register-local FADD2s, and the cross-thread exchanges are two
// swizzled, bank-conflict-free shared-memory transpositions (layout.cuh); global traffic
// is one 128-bit coalesced read and one 128-bit coalesced write per element (8 B/elt,
// the algorithmic minimum; the reference moves 16 B/elt because of its clone).
#include "common.cuh"
#include "engine.cuh"

namespace whvi {

// SCALED: out = H(scale * in) with a (D) fp32 vector broadcast over the rows -- the hoisted first transform t2 = H(s2 * x) of
// the MC predictive evaluation without a separate scaling pass.
template <int N, int C, int K, int GROUPS, class IO, bool SCALED = false>
__global__ void __launch_bounds__((1 << (N - C)) * GROUPS)
fwht_kernel(const IO* __restrict__ in, IO* __restrict__ out, int64_t total, const float* __restrict__ scale = nullptr)
{
    constexpr int T = 1 << (N - C);
    constexpr int E = 1 << C;
    constexpr int64_t TILE = int64_t(1) << N;
    constexpr int SEQ = seq_pack(V_FIRST, V_MID, V_LAST);
    static_assert(rounds_needed(N, C, K) <= 3, "FIRST+MID+LAST must cover every transform bit");
    // only bits 0,1 to transform: FIRST alone covers them and is already store-friendly
    constexpr bool kOneRound = (K <= 2);

    extern __shared__ float4 smem4[];
    const int group = threadIdx.x / T;
    const uint32_t tid = threadIdx.x % T;
    // N == 15: a 128 KB transposition buffer leaves one CTA per SM, so nothing overlaps a tile's load latency with another
    // tile's work; there the grid is persistent (one CTA per SM looping over tiles) and the next tile is pulled into L2
    // while the current one is transformed.  Smaller tiles have several CTAs per SM: one tile per group, no loop.
    constexpr bool PERSIST = (N >= 15) && GROUPS == 1;
    float* buf = reinterpret_cast<float*>(smem4) + size_t(group) * (kOneRound ? 0 : scratch_words(N, C));
    const uint32_t toff = tile_thread_offset<N, C, V_FIRST>(tid);
    const int64_t step = PERSIST ? int64_t(gridDim.x) * GROUPS * TILE : total;
#pragma unroll 1
    for (int64_t base = (int64_t(blockIdx.x) * GROUPS + group) * TILE; base < total; base += step) {
    if constexpr (PERSIST) {
        if (tid == 0 && base + step < total) {
            const int64_t left1 = total - (base + step);
            l2_prefetch_bulk(in + base + step, static_cast<uint32_t>((left1 < TILE ? left1 : TILE) * sizeof(IO)));
        }
        if (base != (int64_t(blockIdx.x) * GROUPS + group) * TILE) group_sync<T, GROUPS>(group);   // the previous tile's reads of buf are done
    }
    float v[E];
    static_for<0, E / 4>([&](auto m_) {
        constexpr int m = decltype(m_)::value;
        constexpr uint32_t roff = tile_reg_offset<N, C, V_FIRST>(m);
        const int64_t g = base + toff + roff;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < total) q = Io<IO>::ld4(in + g);
        if constexpr (SCALED) {
            const float4 w = __ldg(reinterpret_cast<const float4*>(scale + ((toff + roff) & ((1u << K) - 1u))));
            // __fmul_rn: never contracted into the first butterfly's add, so the result is bit-identical to scale-then-transform
            q.x = __fmul_rn(q.x, w.x), q.y = __fmul_rn(q.y, w.y), q.z = __fmul_rn(q.z, w.z), q.w = __fmul_rn(q.w, w.w);
        }
        v[4 * m + 0] = q.x;
        v[4 * m + 1] = q.y;
        v[4 * m + 2] = q.z;
        v[4 * m + 3] = q.w;
    });
    bfly_round<N, C, K, SEQ, 0>(v);

    if constexpr (kOneRound) {
        static_for<0, E / 4>([&](auto m_) {
            constexpr int m = decltype(m_)::value;
            constexpr uint32_t roff = tile_reg_offset<N, C, V_FIRST>(m);
            const int64_t g = base + toff + roff;
            if (g < total) Io<IO>::st4(out + g, make_float4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]));
        });
    } else {
        transpose_write<N, C, V_FIRST, V_MID>(v, buf, transpose_writer_base<N, C, V_FIRST, V_MID>(tid));
        group_sync<T, GROUPS>(group);
        transpose_read<C>(v, buf, tid);
        bfly_round<N, C, K, SEQ, 1>(v);
        group_sync<T, GROUPS>(group);  // every read of buf is done before it is overwritten
        transpose_write<N, C, V_MID, V_LAST>(v, buf, transpose_writer_base<N, C, V_MID, V_LAST>(tid));
        group_sync<T, GROUPS>(group);
        transpose_read<C>(v, buf, tid);
        bfly_round<N, C, K, SEQ, 2>(v);

        const uint32_t soff = tile_thread_offset<N, C, V_LAST>(tid);
        static_for<0, E / 4>([&](auto m_) {
            constexpr int m = decltype(m_)::value;
            constexpr uint32_t roff = tile_reg_offset<N, C, V_LAST>(m);
            const int64_t g = base + soff + roff;
            if (g < total) Io<IO>::st4(out + g, make_float4(v[4 * m], v[4 * m + 1], v[4 * m + 2], v[4 * m + 3]));
        });
    }
    }   // tile loop
}

// D = 1 (identity) and D = 2: too narrow for a float4; one thread per row.
template <class IO>
__global__ void fwht_tiny_kernel(const IO* __restrict__ in, IO* __restrict__ out, int64_t rows, int D)
{
    const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    if (D == 1) {
        out[r] = in[r];
    } else {
        const float a = static_cast<float>(in[2 * r]), b = static_cast<float>(in[2 * r + 1]);
        out[2 * r] = static_cast<IO>(a + b);
        out[2 * r + 1] = static_cast<IO>(a - b);
    }
}

// Multi-pass global variant for D > 2^15: after the low 15 bits have been transformed row
// segment by row segment with the single-pass kernel, the remaining high bits are done in
// place, R = 2^r (r <= 4) bits per pass.  A thread owns one float4 column and the R values
// that differ in the bit group [log2_stride, log2_stride + r): lanes run along the
// contiguous dimension, so every access is a coalesced 128-bit access.  Each extra pass costs
// one more read + write of the data (8 B/elt).
template <int R>
__global__ void __launch_bounds__(256)
fwht_strided_kernel(float4* __restrict__ data, int64_t n_threads, int log2_stride4 /* stride in float4 units */)
{
    constexpr int r = R == 2 ? 1 : (R == 4 ? 2 : (R == 8 ? 3 : 4));
    const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= n_threads) return;
    const int64_t low = idx & ((int64_t(1) << log2_stride4) - 1);
    const int64_t high = idx >> log2_stride4;
    float4* base = data + ((high << (log2_stride4 + r)) | low);
    const int64_t stride = int64_t(1) << log2_stride4;
    float4 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = base[j * stride];
#pragma unroll
    for (int h = 1; h < R; h <<= 1) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (j & h) continue;
            const float4 a = v[j], b = v[j | h];
            v[j] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
            v[j | h] = make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j) base[j * stride] = v[j];
}

static int launch_strided(float* data, int64_t total, int log2_stride, int r, cudaStream_t stream)
{
    const int64_t n_threads = (total / 4) >> r;
    const int threads = 256;
    const int64_t blocks = (n_threads + threads - 1) / threads;
    if (blocks > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "fwht: strided pass exceeds the grid limit");
    float4* d4 = reinterpret_cast<float4*>(data);
    const unsigned g = static_cast<unsigned>(blocks);
    switch (r) {
    case 1: fwht_strided_kernel<2><<<g, threads, 0, stream>>>(d4, n_threads, log2_stride - 2); break;
    case 2: fwht_strided_kernel<4><<<g, threads, 0, stream>>>(d4, n_threads, log2_stride - 2); break;
    case 3: fwht_strided_kernel<8><<<g, threads, 0, stream>>>(d4, n_threads, log2_stride - 2); break;
    default: fwht_strided_kernel<16><<<g, threads, 0, stream>>>(d4, n_threads, log2_stride - 2); break;
    }
    return check_launch("fwht_strided_kernel");
}

template <int N, int C, int K, int GROUPS, class IO, bool SCALED = false>
static int launch_cfg(const IO* in, IO* out, int64_t total, cudaStream_t stream, const float* scale = nullptr)
{
    static unsigned char smem_ok[64] = {};
    constexpr int threads = (1 << (N - C)) * GROUPS;
    constexpr size_t smem = (K <= 2) ? 0 : sizeof(float) * size_t(scratch_words(N, C)) * GROUPS;
    auto kernel = fwht_kernel<N, C, K, GROUPS, IO, SCALED>;
    if (int rc = ensure_smem(kernel, smem, smem_ok)) return rc;
    const int64_t tiles = (total + (int64_t(1) << N) - 1) >> N;
    int64_t ctas = (tiles + GROUPS - 1) / GROUPS;
    if (N >= 15 && GROUPS == 1 && ctas > 148) ctas = 148;   // persistent: one CTA per SM (see the kernel)
    if (ctas > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "fwht: %lld tiles exceed the grid limit", (long long)ctas);
    kernel<<<static_cast<unsigned>(ctas), threads, smem, stream>>>(in, out, total, scale);
    return check_launch("fwht_kernel");
}

// out = H(scale * in), fp32, the layer kernels' range of D (4 .. 2^15)
int launch_fwht_scaled(const float* in, const float* scale, float* out, int64_t rows, int64_t D, cudaStream_t stream)
{
    const int64_t total = rows * D;
    switch (ilog2(D)) {
    case 2: return launch_cfg<10, 5, 2, 8, float, true>(in, out, total, stream, scale);
    case 3: return launch_cfg<10, 5, 3, 8, float, true>(in, out, total, stream, scale);
    case 4: return launch_cfg<10, 5, 4, 8, float, true>(in, out, total, stream, scale);
    case 5: return launch_cfg<10, 5, 5, 8, float, true>(in, out, total, stream, scale);
    case 6: return launch_cfg<10, 5, 6, 8, float, true>(in, out, total, stream, scale);
    case 7: return launch_cfg<10, 5, 7, 8, float, true>(in, out, total, stream, scale);
    case 8: return launch_cfg<10, 5, 8, 8, float, true>(in, out, total, stream, scale);
    case 9: return launch_cfg<10, 5, 9, 8, float, true>(in, out, total, stream, scale);
    case 10: return launch_cfg<10, 5, 10, 8, float, true>(in, out, total, stream, scale);
    case 11: return launch_cfg<11, 5, 11, 4, float, true>(in, out, total, stream, scale);
    case 12: return launch_cfg<12, 5, 12, 2, float, true>(in, out, total, stream, scale);
    case 13: return launch_cfg<13, 5, 13, 1, float, true>(in, out, total, stream, scale);
    case 14: return launch_cfg<14, 6, 14, 1, float, true>(in, out, total, stream, scale);
    case 15: return launch_cfg<15, 6, 15, 1, float, true>(in, out, total, stream, scale);
    default: break;
    }
    return fail(WHVI_E_SHAPE, "fwht_scaled: D = %lld outside [4, 32768]", (long long)D);
}

// single-pass range (D <= 2^15); returns -1 when D is beyond it
template <class IO>
static int launch_fwht_single(const IO* in, IO* out, int64_t rows, int64_t D, cudaStream_t stream)
{
    const int K = ilog2(D);
    const int64_t total = rows * D;
    if (K <= 1) {
        const int threads = 256;
        const int64_t blocks = (rows + threads - 1) / threads;
        fwht_tiny_kernel<IO><<<static_cast<unsigned>(blocks), threads, 0, stream>>>(in, out, rows, static_cast<int>(D));
        return check_launch("fwht_tiny_kernel");
    }
    switch (K) {
    // D <= 1024: one warp per 1024-float tile (1024/D rows), 8 tiles per CTA
    case 2: return launch_cfg<10, 5, 2, 8>(in, out, total, stream);
    case 3: return launch_cfg<10, 5, 3, 8>(in, out, total, stream);
    case 4: return launch_cfg<10, 5, 4, 8>(in, out, total, stream);
    case 5: return launch_cfg<10, 5, 5, 8>(in, out, total, stream);
    case 6: return launch_cfg<10, 5, 6, 8>(in, out, total, stream);
    case 7: return launch_cfg<10, 5, 7, 8>(in, out, total, stream);
    case 8: return launch_cfg<10, 5, 8, 8>(in, out, total, stream);
    case 9: return launch_cfg<10, 5, 9, 8>(in, out, total, stream);
    case 10: return launch_cfg<10, 5, 10, 8>(in, out, total, stream);
    // one row per group of 64/128/256 threads
    case 11: return launch_cfg<11, 5, 11, 4>(in, out, total, stream);
    case 12: return launch_cfg<12, 5, 12, 2>(in, out, total, stream);
    case 13: return launch_cfg<13, 5, 13, 1>(in, out, total, stream);
    // 64 floats per thread
    case 14: return launch_cfg<14, 6, 14, 1>(in, out, total, stream);
    case 15: return launch_cfg<15, 6, 15, 1>(in, out, total, stream);
    default: break;
    }
    return -1;
}

// bf16 in HBM, fp32 butterflies in registers, one rounding (to nearest even) at the store
int launch_fwht_bf16(const void* in, void* out, int64_t rows, int64_t D, cudaStream_t stream)
{
    const int rc = launch_fwht_single(static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), rows, D, stream);
    if (rc == -1) return fail(WHVI_E_SHAPE, "fwht_bf16: D = %lld exceeds the single-pass limit 2^%d", (long long)D, kMaxLog2D);
    return rc;
}

int launch_fwht(const float* in, float* out, int64_t rows, int64_t D, cudaStream_t stream)
{
    const int K = ilog2(D);
    const int64_t total = rows * D;
    {
        const int rc = launch_fwht_single(in, out, rows, D, stream);
        if (rc != -1) return rc;
    }
    if (K > kMaxLog2Dmulti) return fail(WHVI_E_SHAPE, "fwht: D = %lld exceeds the limit 2^%d", (long long)D, kMaxLog2Dmulti);
    // D > 2^15: low 15 bits in one pass (every 2^15-float segment is a "row"), then the high bits
    if (int rc = launch_cfg<15, 6, 15, 1>(in, out, total, stream)) return rc;
    for (int b = kMaxLog2D; b < K;) {
        const int r = (K - b) < 4 ? (K - b) : 4;
        if (int rc = launch_strided(out, total, b, r, stream)) return rc;
        b += r;
    }
    return WHVI_OK;
}

}  // namespace whvi
