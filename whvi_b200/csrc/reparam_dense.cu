// Kernel (4): dense-covariance reparameterisation  G[S,D] = mu + E[S,D] . L^T  (L lower
// triangular, D x D) as a hand-written tcgen05 tensor-core GEMM.
//
// The reference's posterior is diagonal (g = mu + softplus(rho)*eps, src/weights.py:43-50,
// :82-83); this is the north-star's superset: one dense contraction over the noise
// per MC sample.  Not in the reference => no reference golden; the oracle is
// oracle_reparam(dense=1) = mu + eps @ L^T in fp64.
//
// Mapping onto the 5th-gen tensor core (cta_group::1, kind::tf32, M = 128, K = 8 per MMA):
//   D_tmem[i, s] = sum_j L[i0+i, j] * E[s0+s, j]       A = L rows (K-major), B = E rows (K-major)
// One CTA per (128-row block of L, <=256-sample block of E).  The accumulator lives in TMEM
// (128 lanes x N columns, fp32); the epilogue reads it back with tcgen05.ld (lane = row i of L =
// output coordinate), adds mu and stores G[s, i] coalesced over i.  Only the lower triangle is
// visited (K runs to i0+128), entries above the diagonal are treated as zero whatever the buffer
// holds.
// fp32 accuracy on tf32 hardware: every operand is split x = hi + lo with hi = x rounded to the
// 10-bit tf32 mantissa by truncation and lo = x - hi (exact), and hi*hi + hi*lo + lo*hi is
// accumulated in fp32 (3 MMAs per K-step; the dropped lo*lo term is ~2^-22 relative).
// Operands are staged by the CTA's threads (global -> registers -> split -> shared) into the
// canonical no-swizzle K-major layout  [k/4][row][4 floats]  (core matrix = 8 rows x 16 B
// contiguous; SBO = 128 B between 8-row groups, LBO = rows*16 B between the two 16-byte K halves
// of one MMA), two stages deep so that staging tile k+1 overlaps the MMAs of tile k
// (tcgen05.commit -> mbarrier frees a stage).
#include "common.cuh"
#include "engine.cuh"

namespace whvi {

constexpr int RD_M = 128;    // rows of L per CTA = TMEM lanes
constexpr int RD_KC = 32;    // K elements staged per pipeline step (4 MMAs of K = 8)
constexpr int RD_THREADS = 256;

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    // cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
    // layout_type [61,64) = 0 (SWIZZLE_NONE / interleave)
    return uint64_t((smem_addr >> 4) & 0x3FFFu) | (uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           (uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46);
}

__device__ __forceinline__ uint32_t umma_idesc_tf32(int m, int n)
{
    // cute::UMMA::InstrDescriptor: c_format F32 = 1 [4,6), a/b_format TF32 = 2 [7,10) [10,13),
    // a/b K-major (0) [15],[16], n>>3 [17,23), m>>4 [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo)
{
    hi.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
    hi.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
    hi.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
    hi.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
    lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
}

// smem per stage: A_hi, A_lo: [KC/4][128][4] floats each; B_hi, B_lo: [KC/4][NP][4] floats each
__global__ void __launch_bounds__(RD_THREADS, 1)
reparam_dense_kernel(const float* __restrict__ mu, const float* __restrict__ L, const float* __restrict__ eps,
                     float* __restrict__ g, int S, int D, int NP /* samples per CTA padded to 16 */)
{
    extern __shared__ float4 smem4[];
    __shared__ uint64_t empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    float* smem = reinterpret_cast<float*>(smem4);
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int i0 = blockIdx.x * RD_M;
    const int s0 = blockIdx.y * 256;
    const int ns = min(256, S - s0);            // valid samples in this CTA (<= NP)
    const int a_floats = RD_KC * RD_M;          // per hi or lo
    const int b_floats = RD_KC * NP;
    const int stage_floats = 2 * a_floats + 2 * b_floats;

    if (tid == 0) {
        mbar_init(&empty_bar[0], 1);
        mbar_init(&empty_bar[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) {  // one warp allocates 256 TMEM columns (fp32 accumulator 128 lanes x 256)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "r"(256)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;
    const uint32_t idesc = umma_idesc_tf32(RD_M, NP);

    const int nk = (i0 + RD_M) / RD_KC;  // lower triangle: columns j < i0 + 128 only
    for (int kb = 0; kb < nk; ++kb) {
        const int st = kb & 1;
        const int k0 = kb * RD_KC;
        float* a_hi = smem + st * stage_floats;
        float* a_lo = a_hi + a_floats;
        float* b_hi = a_lo + a_floats;
        float* b_lo = b_hi + b_floats;
        if (kb >= 2) mbar_wait(&empty_bar[st], ((kb >> 1) - 1) & 1);  // MMAs that read this stage are done
        // ---- stage A = L[i0 .. i0+128, k0 .. k0+32), zero above the diagonal
        for (int idx = tid; idx < RD_M * (RD_KC / 4); idx += RD_THREADS) {
            const int row = idx / (RD_KC / 4), c4 = idx % (RD_KC / 4);
            const int i = i0 + row, j = k0 + 4 * c4;
            float4 x = __ldg(reinterpret_cast<const float4*>(L + size_t(i) * D + j));
            if (j + 0 > i) x.x = 0.f;
            if (j + 1 > i) x.y = 0.f;
            if (j + 2 > i) x.z = 0.f;
            if (j + 3 > i) x.w = 0.f;
            float4 hi, lo;
            split_tf32(x, hi, lo);
            *reinterpret_cast<float4*>(a_hi + (c4 * RD_M + row) * 4) = hi;
            *reinterpret_cast<float4*>(a_lo + (c4 * RD_M + row) * 4) = lo;
        }
        // ---- stage B = E[s0 .. s0+NP, k0 .. k0+32), zero rows beyond the valid samples
        for (int idx = tid; idx < NP * (RD_KC / 4); idx += RD_THREADS) {
            const int row = idx / (RD_KC / 4), c4 = idx % (RD_KC / 4);
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < ns) x = __ldg(reinterpret_cast<const float4*>(eps + size_t(s0 + row) * D + k0 + 4 * c4));
            float4 hi, lo;
            split_tf32(x, hi, lo);
            *reinterpret_cast<float4*>(b_hi + (c4 * NP + row) * 4) = hi;
            *reinterpret_cast<float4*>(b_lo + (c4 * NP + row) * 4) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_lbo = RD_M * 16, b_lbo = NP * 16, sbo = 128;
#pragma unroll
            for (int ks = 0; ks < RD_KC / 8; ++ks) {
                // the two 16-byte K halves of this MMA are chunks 2*ks and 2*ks+1
                const uint64_t dah = umma_smem_desc(smem_u32(a_hi) + ks * 2 * a_lbo, a_lbo, sbo);
                const uint64_t dal = umma_smem_desc(smem_u32(a_lo) + ks * 2 * a_lbo, a_lbo, sbo);
                const uint64_t dbh = umma_smem_desc(smem_u32(b_hi) + ks * 2 * b_lbo, b_lbo, sbo);
                const uint64_t dbl = umma_smem_desc(smem_u32(b_lo) + ks * 2 * b_lbo, b_lbo, sbo);
                umma_tf32(tmem_base, dah, dbh, idesc, (kb | ks) != 0);
                umma_tf32(tmem_base, dah, dbl, idesc, 1);
                umma_tf32(tmem_base, dal, dbh, idesc, 1);
            }
            umma_commit(&empty_bar[st]);  // arrives when every MMA issued so far has completed
        }
    }
    // the last commit (stage (nk-1)&1) covers all MMAs
    {
        const int kb = nk - 1;
        mbar_wait(&empty_bar[kb & 1], (kb >> 1) & 1);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: warps 0..3 own TMEM lanes 32w .. 32w+31 = output coordinates i0 + 32w + lane
    if (warp < 4) {
        const int i = i0 + 32 * warp + lane;
        const float m = mu[i];
        for (int c = 0; c < NP; c += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_base + (uint32_t(32 * warp) << 16) + uint32_t(c);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 16; ++q)
                if (c + q < ns) g[size_t(s0 + c + q) * D + i] = __uint_as_float(v[q]) + m;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

int launch_reparam_dense(const float* mu, const float* L, const float* eps, float* g, int64_t S, int64_t D, cudaStream_t stream)
{
    static unsigned char smem_ok[64] = {};
    if (D % RD_M != 0) return fail(WHVI_E_SHAPE, "reparam(dense): D = %lld must be a multiple of 128", (long long)D);
    const int blocks_s = static_cast<int>((S + 255) / 256);
    const int ns_max = static_cast<int>(S < 256 ? S : 256);
    const int NP = (ns_max + 15) / 16 * 16;
    const size_t smem = sizeof(float) * 2 * (2 * RD_KC * RD_M + 2 * RD_KC * NP);
    const size_t smem_max = sizeof(float) * 2 * (2 * RD_KC * RD_M + 2 * RD_KC * 256);  // opt in once, for any NP
    if (int rc = ensure_smem(reparam_dense_kernel, smem_max, smem_ok)) return rc;
    dim3 grid(static_cast<unsigned>(D / RD_M), static_cast<unsigned>(blocks_s));
    reparam_dense_kernel<<<grid, RD_THREADS, smem, stream>>>(mu, L, eps, g, static_cast<int>(S), static_cast<int>(D), NP);
    return check_launch("reparam_dense_kernel");
}

}  // namespace whvi
