// extern "C" surface of libwhvi_b200.so (declared in include/whvi_b200.h): argument
// validation, error text, dispatch to the launchers.  Nothing here allocates or syncs.
#include "common.cuh"

namespace whvi {
char* error_buffer()
{
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace whvi

using namespace whvi;

extern "C" {

int whvi_abi_version(void) { return WHVI_ABI_VERSION; }

const char* whvi_last_error(void) { return error_buffer(); }

int64_t whvi_max_dim(void) { return int64_t(1) << kMaxLog2Dmulti; }

int whvi_fwht_f32(const float* in, float* out, int64_t rows, int64_t D, whvi_stream_t stream)
{
    if (rows < 0 || D < 1) return fail(WHVI_E_SHAPE, "fwht: rows=%lld D=%lld", (long long)rows, (long long)D);
    if (!is_pow2(D)) return fail(WHVI_E_SHAPE, "fwht: n must be a power of 2 (got %lld)", (long long)D);
    if (rows == 0) return WHVI_OK;
    if (!in || !out) return fail(WHVI_E_NULL, "fwht: null pointer");
    if (!aligned16(in) || !aligned16(out)) return fail(WHVI_E_ALIGN, "fwht: pointers must be 16-byte aligned");
    return launch_fwht(in, out, rows, D, static_cast<cudaStream_t>(stream));
}

int whvi_fwht_scaled_f32(const float* in, const float* scale, float* out, int64_t rows, int64_t D, whvi_stream_t stream)
{
    if (rows < 0 || D < 4) return fail(WHVI_E_SHAPE, "fwht_scaled: rows=%lld D=%lld (D >= 4)", (long long)rows, (long long)D);
    if (!is_pow2(D)) return fail(WHVI_E_SHAPE, "fwht_scaled: n must be a power of 2 (got %lld)", (long long)D);
    if (rows == 0) return WHVI_OK;
    if (!in || !scale || !out) return fail(WHVI_E_NULL, "fwht_scaled: null pointer");
    if (!aligned16(in) || !aligned16(out) || !aligned16(scale)) return fail(WHVI_E_ALIGN, "fwht_scaled: pointers must be 16-byte aligned");
    return launch_fwht_scaled(in, scale, out, rows, D, static_cast<cudaStream_t>(stream));
}

int whvi_fwht_bf16(const void* in, void* out, int64_t rows, int64_t D, whvi_stream_t stream)
{
    if (rows < 0 || D < 1) return fail(WHVI_E_SHAPE, "fwht_bf16: rows=%lld D=%lld", (long long)rows, (long long)D);
    if (!is_pow2(D)) return fail(WHVI_E_SHAPE, "fwht_bf16: n must be a power of 2 (got %lld)", (long long)D);
    if (rows == 0) return WHVI_OK;
    if (!in || !out) return fail(WHVI_E_NULL, "fwht_bf16: null pointer");
    if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 7u) return fail(WHVI_E_ALIGN, "fwht_bf16: pointers must be 8-byte aligned");
    return launch_fwht_bf16(in, out, rows, D, static_cast<cudaStream_t>(stream));
}

int whvi_fwht_f64(const double* in, double* out, int64_t rows, int64_t D, whvi_stream_t stream)
{
    if (rows < 0 || D < 1) return fail(WHVI_E_SHAPE, "fwht_f64: rows=%lld D=%lld", (long long)rows, (long long)D);
    if (!is_pow2(D)) return fail(WHVI_E_SHAPE, "fwht_f64: n must be a power of 2 (got %lld)", (long long)D);
    if (rows == 0) return WHVI_OK;
    if (!in || !out) return fail(WHVI_E_NULL, "fwht_f64: null pointer");
    if (!aligned16(in) || !aligned16(out)) return fail(WHVI_E_ALIGN, "fwht_f64: pointers must be 16-byte aligned");
    if (rows > (int64_t(1) << 40) / D) return fail(WHVI_E_SHAPE, "fwht_f64: too many elements");
    return launch_fwht_f64(in, out, rows, D, static_cast<cudaStream_t>(stream));
}

static int check_layer_shape(const char* who, int64_t S, int64_t B, int64_t D, int64_t xs, int64_t max_d = 8192)
{
    if (S < 0 || B < 0 || D < 1) return fail(WHVI_E_SHAPE, "%s: S=%lld B=%lld D=%lld", who, (long long)S, (long long)B, (long long)D);
    if (!is_pow2(D)) return fail(WHVI_E_SHAPE, "%s: D must be a power of 2 (got %lld)", who, (long long)D);
    if (D < 4 || D > max_d) return fail(WHVI_E_SHAPE, "%s: D = %lld outside [4, %lld]", who, (long long)D, (long long)max_d);
    if (xs != 0 && xs != B * D) return fail(WHVI_E_SHAPE, "%s: x_sample_stride must be 0 or B*D", who);
    return WHVI_OK;
}

int whvi_layer_fwd_partials(int64_t S, int64_t B, int64_t D, int64_t* count)
{
    if (!count) return fail(WHVI_E_NULL, "layer_fwd_partials: null pointer");
    if (int rc = check_layer_shape("layer_fwd_partials", S, B, D, 0, 32768)) return rc;
    *count = 0;
    if (S == 0 || B == 0) return WHVI_OK;
    size_t n = 0;
    LayerFwdCall c{};
    c.S = S;
    c.B = B;
    c.partials_needed = &n;
    const int rc = launch_layer_fwd(c, D, nullptr);
    *count = static_cast<int64_t>(n);
    return rc;
}

int whvi_layer_fwd_fused_f32(const float* x, int64_t x_sample_stride, const float* g, const float* s1, const float* s2,
                             const float* bias, float* y, int64_t S, int64_t B, int64_t D, int flags,
                             const float* target, float* sq_partials, whvi_stream_t stream)
{
    if (int rc = check_layer_shape("layer_fwd", S, B, D, x_sample_stride, 32768)) return rc;
    if (flags & ~(WHVI_LAYER_RELU_OUT | WHVI_LAYER_FROM_T2)) return fail(WHVI_E_MODE, "layer_fwd: unknown flags %d", flags);
    if (S == 0 || B == 0) return WHVI_OK;
    if (!x || !g || !s1 || !s2 || !y) return fail(WHVI_E_NULL, "layer_fwd: null pointer");
    if (target && !sq_partials) return fail(WHVI_E_NULL, "layer_fwd: target given without sq_partials");
    if (!aligned16(x) || !aligned16(g) || !aligned16(s1) || !aligned16(s2) || !aligned16(y) || !aligned16(bias) ||
        !aligned16(target))
        return fail(WHVI_E_ALIGN, "layer_fwd: pointers must be 16-byte aligned");
    LayerFwdCall c{x, g, s1, s2, bias, target, y, sq_partials, x_sample_stride, S, B, flags & WHVI_LAYER_RELU_OUT, nullptr};
    c.from_t2 = (flags & WHVI_LAYER_FROM_T2) ? 1 : 0;
    return launch_layer_fwd(c, D, static_cast<cudaStream_t>(stream));
}

int whvi_layer_fwd_bf16(const void* x, int64_t x_sample_stride, const float* g, const float* s1, const float* s2, const float* bias,
                        void* y, int64_t S, int64_t B, int64_t D, int flags, whvi_stream_t stream)
{
    if (int rc = check_layer_shape("layer_fwd_bf16", S, B, D, x_sample_stride, 32768)) return rc;
    if (flags & ~(WHVI_LAYER_RELU_OUT | WHVI_LAYER_FROM_T2)) return fail(WHVI_E_MODE, "layer_fwd_bf16: unknown flags %d", flags);
    if (S == 0 || B == 0) return WHVI_OK;
    if (!x || !g || !s1 || !s2 || !y) return fail(WHVI_E_NULL, "layer_fwd_bf16: null pointer");
    if (!aligned16(g) || !aligned16(s1) || !aligned16(s2) || !aligned16(bias) ||
        ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 7u))
        return fail(WHVI_E_ALIGN, "layer_fwd_bf16: fp32 pointers must be 16-byte aligned, bf16 pointers 8-byte aligned");
    LayerFwdCall c{static_cast<const float*>(x), g, s1, s2, bias, nullptr, static_cast<float*>(y), nullptr, x_sample_stride, S, B,
                   flags & WHVI_LAYER_RELU_OUT, nullptr};
    c.from_t2 = (flags & WHVI_LAYER_FROM_T2) ? 1 : 0;
    c.bf16 = 1;
    return launch_layer_fwd(c, D, static_cast<cudaStream_t>(stream));
}

int whvi_layer_fwd_f32(const float* x, int64_t x_sample_stride, const float* g, const float* s1, const float* s2,
                       const float* bias, float* y, int64_t S, int64_t B, int64_t D, whvi_stream_t stream)
{
    return whvi_layer_fwd_fused_f32(x, x_sample_stride, g, s1, s2, bias, y, S, B, D, 0, nullptr, nullptr, stream);
}

int whvi_layer_bwd_workspace_bytes(int64_t S, int64_t B, int64_t D, size_t* bytes)
{
    if (!bytes) return fail(WHVI_E_NULL, "layer_bwd_workspace_bytes: null pointer");
    if (int rc = check_layer_shape("layer_bwd_workspace_bytes", S, B, D, 0)) return rc;
    *bytes = 0;
    if (S == 0 || B == 0) return WHVI_OK;
    LayerBwdCall c{};
    c.S = S;
    c.B = B;
    c.need_only = bytes;
    return launch_layer_bwd(c, D, nullptr);
}

static int layer_bwd_common(const float* x, int64_t x_sample_stride, const float* dy, const float* g, const float* s1,
                            const float* s2, float* dx, float* dg, float* ds1, float* ds2, float* dbias,
                            void* workspace, size_t workspace_bytes, int64_t S, int64_t B, int64_t D, int flags,
                            const float* target, const float* coef, const float* dy_scale, whvi_stream_t stream)
{
    if (int rc = check_layer_shape("layer_bwd", S, B, D, x_sample_stride)) return rc;
    if (flags & ~WHVI_LAYER_RELU_IN) return fail(WHVI_E_MODE, "layer_bwd: unknown flags %d", flags);
    if (!dg || !ds1 || !ds2) return fail(WHVI_E_NULL, "layer_bwd: null output pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (S == 0 || B == 0) {  // empty sums
        if (S > 0) cudaMemsetAsync(dg, 0, sizeof(float) * size_t(S) * D, st);
        cudaMemsetAsync(ds1, 0, sizeof(float) * D, st);
        cudaMemsetAsync(ds2, 0, sizeof(float) * D, st);
        if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * D, st);
        return check_launch("layer_bwd(empty)");
    }
    if (!x || !dy || !g || !s1 || !s2) return fail(WHVI_E_NULL, "layer_bwd: null pointer");
    if (target && !coef) return fail(WHVI_E_NULL, "layer_bwd: target given without coef");
    if (!aligned16(x) || !aligned16(dy) || !aligned16(g) || !aligned16(s1) || !aligned16(s2) || !aligned16(dx) ||
        !aligned16(workspace) || !aligned16(target))
        return fail(WHVI_E_ALIGN, "layer_bwd: pointers must be 16-byte aligned");
    LayerBwdCall c{x, dy, g, s1, s2, target, coef, nullptr, dx, dg, ds1, ds2, dbias, static_cast<float*>(workspace),
                   workspace_bytes, x_sample_stride, S, B, flags & WHVI_LAYER_RELU_IN, nullptr};
    c.dy_scale = dy_scale;
    return launch_layer_bwd(c, D, st);
}

int whvi_layer_bwd_fused_f32(const float* x, int64_t x_sample_stride, const float* dy, const float* g, const float* s1,
                             const float* s2, float* dx, float* dg, float* ds1, float* ds2, float* dbias,
                             void* workspace, size_t workspace_bytes, int64_t S, int64_t B, int64_t D, int flags,
                             const float* target, const float* coef, whvi_stream_t stream)
{
    return layer_bwd_common(x, x_sample_stride, dy, g, s1, s2, dx, dg, ds1, ds2, dbias, workspace, workspace_bytes, S, B, D,
                            flags, target, coef, nullptr, stream);
}

int whvi_layer_bwd_scaled_f32(const float* x, int64_t x_sample_stride, const float* dy, const float* dy_scale,
                              const float* g, const float* s1, const float* s2, float* dx, float* dg, float* ds1,
                              float* ds2, float* dbias, void* workspace, size_t workspace_bytes, int64_t S, int64_t B,
                              int64_t D, int flags, whvi_stream_t stream)
{
    return layer_bwd_common(x, x_sample_stride, dy, g, s1, s2, dx, dg, ds1, ds2, dbias, workspace, workspace_bytes, S, B, D,
                            flags, nullptr, nullptr, dy_scale, stream);
}

int whvi_layer_loss_sizes(int64_t S, int64_t B, int64_t D, size_t* workspace_bytes, int64_t* sq_count)
{
    if (!workspace_bytes || !sq_count) return fail(WHVI_E_NULL, "layer_loss_sizes: null pointer");
    if (int rc = check_layer_shape("layer_loss_sizes", S, B, D, 0, 4096)) return rc;
    if (D < 128) return fail(WHVI_E_SHAPE, "layer_loss: D = %lld outside [128, 4096]", (long long)D);
    *workspace_bytes = 0;
    *sq_count = 0;
    if (S == 0 || B == 0) return WHVI_OK;
    LayerLossCall c{};
    c.S = S;
    c.B = B;
    c.need_ws = workspace_bytes;
    c.need_sq = sq_count;
    return launch_layer_loss(c, D, nullptr);
}

int whvi_layer_loss_f32(const float* x, int64_t x_sample_stride, const float* g, const float* s1, const float* s2,
                        const float* bias, const float* target, float* dx, float* dg, float* ds1, float* ds2,
                        float* dbias, float* sq_partials, void* workspace, size_t workspace_bytes, int64_t S, int64_t B,
                        int64_t D, int flags, whvi_stream_t stream)
{
    if (int rc = check_layer_shape("layer_loss", S, B, D, x_sample_stride, 4096)) return rc;
    if (D < 128) return fail(WHVI_E_SHAPE, "layer_loss: D = %lld outside [128, 4096]", (long long)D);
    if (flags & ~WHVI_LAYER_RELU_IN) return fail(WHVI_E_MODE, "layer_loss: unknown flags %d", flags);
    if (S == 0 || B == 0) return fail(WHVI_E_SHAPE, "layer_loss: empty batch");
    if (!x || !g || !s1 || !s2 || !target || !dg || !ds1 || !ds2 || !sq_partials || (bias && !dbias))
        return fail(WHVI_E_NULL, "layer_loss: null pointer");
    if (!aligned16(x) || !aligned16(g) || !aligned16(s1) || !aligned16(s2) || !aligned16(bias) || !aligned16(target) ||
        !aligned16(dx) || !aligned16(workspace))
        return fail(WHVI_E_ALIGN, "layer_loss: pointers must be 16-byte aligned");
    LayerLossCall c{x, g, s1, s2, bias, target, dx, dg, ds1, ds2, dbias, sq_partials, static_cast<float*>(workspace),
                    workspace_bytes, x_sample_stride, S, B, flags & WHVI_LAYER_RELU_IN, nullptr, nullptr};
    return launch_layer_loss(c, D, static_cast<cudaStream_t>(stream));
}

int whvi_layer_bwd_f32(const float* x, int64_t x_sample_stride, const float* dy, const float* g, const float* s1,
                       const float* s2, float* dx, float* dg, float* ds1, float* ds2, float* dbias, void* workspace,
                       size_t workspace_bytes, int64_t S, int64_t B, int64_t D, whvi_stream_t stream)
{
    return whvi_layer_bwd_fused_f32(x, x_sample_stride, dy, g, s1, s2, dx, dg, ds1, ds2, dbias, workspace,
                                    workspace_bytes, S, B, D, 0, nullptr, nullptr, stream);
}

int whvi_reparam_f32(const float* mu, const float* rho, const float* eps, float* g, int64_t S, int64_t D, int mode,
                     whvi_stream_t stream)
{
    if (S < 0 || D < 1) return fail(WHVI_E_SHAPE, "reparam: S=%lld D=%lld", (long long)S, (long long)D);
    if (mode != WHVI_REPARAM_DIAG && mode != WHVI_REPARAM_DENSE) return fail(WHVI_E_MODE, "reparam: unknown mode %d", mode);
    if (S == 0) return WHVI_OK;
    if (!mu || !rho || !eps || !g) return fail(WHVI_E_NULL, "reparam: null pointer");
    if (mode == WHVI_REPARAM_DENSE)
        return fail(WHVI_E_MODE, "reparam: the dense form needs a workspace -- call whvi_reparam_dense_f32");
    return launch_reparam_diag(mu, rho, eps, g, S, D, static_cast<cudaStream_t>(stream));
}

int whvi_reparam_dense_workspace_bytes(int64_t S, int64_t D, size_t* bytes)
{
    if (!bytes) return fail(WHVI_E_NULL, "reparam_dense_workspace_bytes: null pointer");
    if (S < 0 || D < 128 || D % 128 != 0) return fail(WHVI_E_SHAPE, "reparam(dense): S=%lld, D=%lld (D must be a multiple of 128)", (long long)S, (long long)D);
    *bytes = reparam_dense_workspace_bytes(S, D);
    return WHVI_OK;
}

int whvi_reparam_dense_f32(const float* mu, const float* L, const float* eps, float* g, int64_t S, int64_t D, void* workspace,
                           size_t workspace_bytes, whvi_stream_t stream)
{
    if (S < 0 || D < 128 || D % 128 != 0) return fail(WHVI_E_SHAPE, "reparam(dense): S=%lld, D=%lld (D must be a multiple of 128)", (long long)S, (long long)D);
    if (S == 0) return WHVI_OK;
    if (!mu || !L || !eps || !g) return fail(WHVI_E_NULL, "reparam(dense): null pointer");
    if (!aligned16(L) || !aligned16(eps) || !aligned16(workspace)) return fail(WHVI_E_ALIGN, "reparam(dense): pointers must be 16-byte aligned");
    return launch_reparam_dense(mu, L, eps, g, S, D, static_cast<float*>(workspace), workspace_bytes, static_cast<cudaStream_t>(stream));
}

int whvi_reparam_dense_bwd_f32(const float* dgT, const float* epsT, float* dL, int64_t S_padded, int64_t D, whvi_stream_t stream)
{
    if (S_padded <= 0 || S_padded % 32 != 0 || D < 128 || D % 128 != 0)
        return fail(WHVI_E_SHAPE, "reparam_bwd(dense): S_padded=%lld (multiple of 32), D=%lld (multiple of 128)", (long long)S_padded, (long long)D);
    if (!dgT || !epsT || !dL) return fail(WHVI_E_NULL, "reparam_bwd(dense): null pointer");
    if (!aligned16(dgT) || !aligned16(epsT) || !aligned16(dL)) return fail(WHVI_E_ALIGN, "reparam_bwd(dense): pointers must be 16-byte aligned");
    return launch_reparam_dense_bwd(dgT, epsT, dL, S_padded, D, static_cast<cudaStream_t>(stream));
}

int whvi_kl_dense_f32(const float* mu, const float* L, float lambda_, int64_t D, float* out_kl, float* dmu, float* dL, float grad_scale,
                      void* workspace, size_t workspace_bytes, whvi_stream_t stream)
{
    if (D < 1) return fail(WHVI_E_SHAPE, "kl(dense): D=%lld", (long long)D);
    if (!(lambda_ > 0.f)) return fail(WHVI_E_SHAPE, "kl(dense): lambda must be positive");
    if (!mu || !L || !out_kl) return fail(WHVI_E_NULL, "kl(dense): null pointer");
    if ((dmu == nullptr) != (dL == nullptr)) return fail(WHVI_E_NULL, "kl(dense): dmu and dL must both be given or both NULL");
    if (!workspace || workspace_bytes < sizeof(double) * size_t(D) || (reinterpret_cast<uintptr_t>(workspace) & 7u))
        return fail(WHVI_E_WORKSPACE, "kl(dense): workspace of %zu bytes (8-byte aligned) needed", sizeof(double) * size_t(D));
    return launch_kl_dense(mu, L, lambda_, D, out_kl, dmu, dL, grad_scale, static_cast<double*>(workspace), static_cast<cudaStream_t>(stream));
}

int whvi_reparam_bwd_f32(const float* rho, const float* eps, const float* dg, float* dmu, float* drho, int64_t S,
                         int64_t D, int mode, int accumulate, whvi_stream_t stream)
{
    if (S < 0 || D < 1) return fail(WHVI_E_SHAPE, "reparam_bwd: S=%lld D=%lld", (long long)S, (long long)D);
    if (mode != WHVI_REPARAM_DIAG) return fail(WHVI_E_MODE, "reparam_bwd: unknown mode %d", mode);
    if (!rho || !dmu || !drho || (S > 0 && (!eps || !dg))) return fail(WHVI_E_NULL, "reparam_bwd: null pointer");
    return launch_reparam_diag_bwd(rho, eps, dg, dmu, drho, S, D, accumulate, static_cast<cudaStream_t>(stream));
}

int whvi_kl_f32(const float* mu, const float* rho, float lambda_, int64_t D, int mode, float* out_kl, float* dmu,
                float* drho, float grad_scale, int accumulate, whvi_stream_t stream)
{
    if (D < 1) return fail(WHVI_E_SHAPE, "kl: D=%lld", (long long)D);
    if (mode != WHVI_KL_REFERENCE && mode != WHVI_KL_CONSISTENT) return fail(WHVI_E_MODE, "kl: unknown mode %d", mode);
    if (!(lambda_ > 0.f)) return fail(WHVI_E_SHAPE, "kl: lambda must be positive");
    if (!mu || !rho || !out_kl) return fail(WHVI_E_NULL, "kl: null pointer");
    if ((dmu == nullptr) != (drho == nullptr)) return fail(WHVI_E_NULL, "kl: dmu and drho must both be given or both NULL");
    return launch_kl(mu, rho, lambda_, D, mode, out_kl, dmu, drho, grad_scale, accumulate, static_cast<cudaStream_t>(stream));
}

// ---- WHVIStackedMatrix as one call per direction (csrc/stacked.cu) ------------------------------------------------------
static int check_stacked(const char* who, int64_t S, int64_t B, int64_t D, int64_t G, int64_t n_out, int64_t xs, int64_t pstride)
{
    if (int rc = check_layer_shape(who, S, B, D, xs)) return rc;
    if (G < 1 || n_out <= (G - 1) * D || n_out > G * D)
        return fail(WHVI_E_SHAPE, "%s: G=%lld blocks of D=%lld do not match n_out=%lld", who, (long long)G, (long long)D, (long long)n_out);
    if (G > 1 && (pstride < D || pstride % 4 != 0))
        return fail(WHVI_E_SHAPE, "%s: param_stride=%lld must be >= D and a multiple of 4", who, (long long)pstride);
    if (S * G > 0x7fffffffLL) return fail(WHVI_E_SHAPE, "%s: too many virtual samples", who);
    return WHVI_OK;
}

int whvi_pad_rows_f32(const float* in, float* out, int64_t rows, int64_t n_in, int64_t D, whvi_stream_t stream)
{
    if (rows < 0 || n_in < 1 || D < n_in || D % 4 != 0) return fail(WHVI_E_SHAPE, "pad_rows: rows=%lld n_in=%lld D=%lld", (long long)rows, (long long)n_in, (long long)D);
    if (rows == 0) return WHVI_OK;
    if (!in || !out) return fail(WHVI_E_NULL, "pad_rows: null pointer");
    if (!aligned16(out)) return fail(WHVI_E_ALIGN, "pad_rows: out must be 16-byte aligned");
    return launch_stack_split(in, out, rows, D, 1, n_in, static_cast<cudaStream_t>(stream));
}

int whvi_stacked_fwd_f32(const float* x, int64_t x_sample_stride, const float* mu, const float* rho, const float* s1, const float* s2,
                         int64_t param_stride, const float* eps, const float* bias, float* g, float* y_blocks, float* y, int64_t S,
                         int64_t B, int64_t D, int64_t G, int64_t n_out, int flags, whvi_stream_t stream)
{
    if (int rc = check_stacked("stacked_fwd", S, B, D, G, n_out, x_sample_stride, param_stride)) return rc;
    if (flags & ~WHVI_LAYER_RELU_OUT) return fail(WHVI_E_MODE, "stacked_fwd: unknown flags %d", flags);
    if (S == 0 || B == 0) return WHVI_OK;
    if (!x || !mu || !rho || !s1 || !s2 || !eps || !g || !y_blocks || !y) return fail(WHVI_E_NULL, "stacked_fwd: null pointer");
    if (!aligned16(x) || !aligned16(s1) || !aligned16(s2) || !aligned16(bias) || !aligned16(g) || !aligned16(y_blocks))
        return fail(WHVI_E_ALIGN, "stacked_fwd: pointers must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = launch_reparam_diag(mu, rho, eps, g, G * S, D, st, G, param_stride)) return rc;
    LayerFwdCall c{x, g, s1, s2, bias, nullptr, y_blocks, nullptr, x_sample_stride, G * S, B, flags & WHVI_LAYER_RELU_OUT, nullptr};
    c.groups = G;
    c.pstride = param_stride;
    c.bstride = D;
    if (int rc = launch_layer_fwd(c, D, st)) return rc;
    return launch_stack_concat(y_blocks, y, S * B, D, G, n_out, st);
}

static size_t stacked_bwd_layout(int64_t S, int64_t B, int64_t D, int64_t G, int want_dx, size_t layer_ws, size_t* off_dx, size_t* off_dg,
                                 size_t* off_ws)
{
    const size_t blocks = sizeof(float) * size_t(G) * S * B * D;
    size_t o = blocks;                       // [0, blocks): dy in the block layout
    *off_dx = o;
    if (want_dx) o += blocks;
    *off_dg = o;
    o += sizeof(float) * size_t(G) * S * D;
    o = (o + 255) / 256 * 256;
    *off_ws = o;
    return o + layer_ws;
}

int whvi_stacked_bwd_workspace_bytes(int64_t S, int64_t B, int64_t D, int64_t G, int want_dx, size_t* bytes)
{
    if (!bytes) return fail(WHVI_E_NULL, "stacked_bwd_workspace_bytes: null pointer");
    if (int rc = check_stacked("stacked_bwd_workspace_bytes", S, B, D, G, G * D, 0, D)) return rc;
    *bytes = 0;
    if (S == 0 || B == 0) return WHVI_OK;
    size_t layer_ws = 0;
    LayerBwdCall c{};
    c.S = S * G;
    c.B = B;
    c.need_only = &layer_ws;
    if (int rc = launch_layer_bwd(c, D, nullptr)) return rc;
    size_t a, b, d;
    *bytes = stacked_bwd_layout(S, B, D, G, want_dx, layer_ws, &a, &b, &d);
    return WHVI_OK;
}

int whvi_stacked_bwd_f32(const float* x, int64_t x_sample_stride, const float* dy, const float* g, const float* rho, const float* s1,
                         const float* s2, int64_t param_stride, const float* eps, float* dx, int64_t n_in, float* dmu, float* drho,
                         float* ds1, float* ds2, float* dbias, void* workspace, size_t workspace_bytes, int64_t S, int64_t B, int64_t D,
                         int64_t G, int64_t n_out, int flags, whvi_stream_t stream)
{
    if (int rc = check_stacked("stacked_bwd", S, B, D, G, n_out, x_sample_stride, param_stride)) return rc;
    if (flags & ~WHVI_LAYER_RELU_IN) return fail(WHVI_E_MODE, "stacked_bwd: unknown flags %d", flags);
    if (dx && (n_in < 1 || n_in > D)) return fail(WHVI_E_SHAPE, "stacked_bwd: n_in=%lld outside [1, D]", (long long)n_in);
    if (!dmu || !drho || !ds1 || !ds2) return fail(WHVI_E_NULL, "stacked_bwd: null output pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (S == 0 || B == 0) {
        for (float* p : {dmu, drho, ds1, ds2, dbias})
            if (p) cudaMemsetAsync(p, 0, sizeof(float) * size_t(G) * D, st);
        return check_launch("stacked_bwd(empty)");
    }
    if (!x || !dy || !g || !rho || !s1 || !s2 || !eps) return fail(WHVI_E_NULL, "stacked_bwd: null pointer");
    if (!aligned16(x) || !aligned16(g) || !aligned16(s1) || !aligned16(s2) || !aligned16(workspace))
        return fail(WHVI_E_ALIGN, "stacked_bwd: pointers must be 16-byte aligned");
    size_t layer_ws = 0;
    {
        LayerBwdCall q{};
        q.S = S * G;
        q.B = B;
        q.need_only = &layer_ws;
        if (int rc = launch_layer_bwd(q, D, nullptr)) return rc;
    }
    size_t off_dx, off_dg, off_ws;
    const size_t need = stacked_bwd_layout(S, B, D, G, dx != nullptr, layer_ws, &off_dx, &off_dg, &off_ws);
    if (!workspace || workspace_bytes < need) return fail(WHVI_E_WORKSPACE, "stacked_bwd: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    char* base = static_cast<char*>(workspace);
    float* dyb = reinterpret_cast<float*>(base);
    float* dxb = dx ? reinterpret_cast<float*>(base + off_dx) : nullptr;
    float* dg = reinterpret_cast<float*>(base + off_dg);
    if (int rc = launch_stack_split(dy, dyb, S * B, D, G, n_out, st)) return rc;
    LayerBwdCall c{x, dyb, g, s1, s2, nullptr, nullptr, nullptr, dxb, dg, ds1, ds2, dbias, reinterpret_cast<float*>(base + off_ws),
                   layer_ws, x_sample_stride, S * G, B, flags & WHVI_LAYER_RELU_IN, nullptr};
    c.groups = G;
    c.pstride = param_stride;
    if (int rc = launch_layer_bwd(c, D, st)) return rc;
    if (int rc = launch_reparam_diag_bwd(rho, eps, dg, dmu, drho, S, D, 0, st, G, param_stride)) return rc;
    if (dx) return launch_stack_sum(dxb, dx, S * B, D, G, n_in, st);
    return WHVI_OK;
}

int whvi_kl_grouped_f32(const float* mu, const float* rho, float lambda_, int64_t D, int64_t G, int64_t param_stride, int mode,
                        float* out_kl, float* dmu, float* drho, float grad_scale, whvi_stream_t stream)
{
    if (D < 1 || G < 1 || (G > 1 && param_stride < D)) return fail(WHVI_E_SHAPE, "kl_grouped: D=%lld G=%lld stride=%lld", (long long)D, (long long)G, (long long)param_stride);
    if (mode != WHVI_KL_REFERENCE && mode != WHVI_KL_CONSISTENT) return fail(WHVI_E_MODE, "kl_grouped: unknown mode %d", mode);
    if (!(lambda_ > 0.f)) return fail(WHVI_E_SHAPE, "kl_grouped: lambda must be positive");
    if (!mu || !rho || !out_kl) return fail(WHVI_E_NULL, "kl_grouped: null pointer");
    if ((dmu == nullptr) != (drho == nullptr)) return fail(WHVI_E_NULL, "kl_grouped: dmu and drho must both be given or both NULL");
    return launch_kl(mu, rho, lambda_, D, mode, out_kl, dmu, drho, grad_scale, 0, static_cast<cudaStream_t>(stream), G, param_stride);
}

// ---- WHVIColumnMatrix as one call per direction (csrc/column.cu) --------------------------------------------------------
static int check_column(const char* who, int64_t S, int64_t B, int64_t D, int64_t n, int64_t xs, int transposed)
{
    if (S < 0 || B < 0 || D < 1 || !is_pow2(D) || D > (int64_t(1) << kMaxLog2D) || n < 1 || n > D || (D > 1 && n <= D / 2))
        return fail(WHVI_E_SHAPE, "%s: S=%lld B=%lld D=%lld n=%lld (D = next_pow2(n) <= 32768)", who, (long long)S, (long long)B, (long long)D, (long long)n);
    const int64_t row = transposed ? n : 1;
    if (xs != 0 && xs != B * row) return fail(WHVI_E_SHAPE, "%s: x_sample_stride must be 0 or B * %lld", who, (long long)row);
    return WHVI_OK;
}

int whvi_column_fwd_f32(const float* x, int64_t x_sample_stride, const float* mu, const float* rho, const float* s1, const float* s2,
                        const float* eps, const float* bias, float* g, float* hg, float* y, int64_t S, int64_t B, int64_t D, int64_t n,
                        int transposed, int flags, whvi_stream_t stream)
{
    if (int rc = check_column("column_fwd", S, B, D, n, x_sample_stride, transposed)) return rc;
    if (flags & ~WHVI_LAYER_RELU_OUT) return fail(WHVI_E_MODE, "column_fwd: unknown flags %d", flags);
    if ((flags & WHVI_LAYER_RELU_OUT) && transposed) return fail(WHVI_E_MODE, "column_fwd: RELU_OUT applies to the (.., 1) -> (.., n) form only");
    if (S == 0 || B == 0) return WHVI_OK;
    if (!x || !mu || !rho || !s1 || !s2 || !eps || !g || !hg || !y) return fail(WHVI_E_NULL, "column_fwd: null pointer");
    if (!aligned16(g) || !aligned16(hg)) return fail(WHVI_E_ALIGN, "column_fwd: g and hg must be 16-byte aligned");
    return launch_column_fwd(x, x_sample_stride, mu, rho, s1, s2, eps, bias, g, hg, y, S, B, D, n, transposed, flags & WHVI_LAYER_RELU_OUT,
                             static_cast<cudaStream_t>(stream));
}

int whvi_column_bwd_workspace_bytes(int64_t S, int64_t D, int64_t n, size_t* bytes)
{
    if (!bytes) return fail(WHVI_E_NULL, "column_bwd_workspace_bytes: null pointer");
    if (S < 0 || D < 1 || n < 1 || n > D) return fail(WHVI_E_SHAPE, "column_bwd_workspace_bytes: S=%lld D=%lld n=%lld", (long long)S, (long long)D, (long long)n);
    *bytes = column_bwd_workspace_bytes(S, D, n);
    return WHVI_OK;
}

int whvi_column_bwd_f32(const float* x, int64_t x_sample_stride, const float* dy, const float* hg, const float* rho, const float* s1,
                        const float* s2, const float* eps, float* dx, float* dmu, float* drho, float* ds1, float* ds2, float* dbias,
                        void* workspace, size_t workspace_bytes, int64_t S, int64_t B, int64_t D, int64_t n, int transposed, int flags,
                        whvi_stream_t stream)
{
    if (int rc = check_column("column_bwd", S, B, D, n, x_sample_stride, transposed)) return rc;
    if (flags & ~WHVI_LAYER_RELU_IN) return fail(WHVI_E_MODE, "column_bwd: unknown flags %d", flags);
    if ((flags & WHVI_LAYER_RELU_IN) && !transposed) return fail(WHVI_E_MODE, "column_bwd: RELU_IN applies to the (.., n) -> (.., 1) form only");
    if (!dmu || !drho || !ds1 || !ds2) return fail(WHVI_E_NULL, "column_bwd: null output pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (S == 0 || B == 0) {
        for (float* p : {dmu, drho, ds1, ds2})
            cudaMemsetAsync(p, 0, sizeof(float) * size_t(D), st);
        if (dbias) cudaMemsetAsync(dbias, 0, sizeof(float) * size_t(transposed ? 1 : n), st);
        return check_launch("column_bwd(empty)");
    }
    if (!x || !dy || !hg || !rho || !s1 || !s2 || !eps) return fail(WHVI_E_NULL, "column_bwd: null pointer");
    if (!aligned16(workspace)) return fail(WHVI_E_ALIGN, "column_bwd: workspace must be 16-byte aligned");
    const size_t need = column_bwd_workspace_bytes(S, D, n);
    if (!workspace || workspace_bytes < need) return fail(WHVI_E_WORKSPACE, "column_bwd: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
    return launch_column_bwd(x, x_sample_stride, dy, hg, rho, s1, s2, eps, dx, dmu, drho, ds1, ds2, dbias, static_cast<float*>(workspace), S,
                             B, D, n, transposed, flags & WHVI_LAYER_RELU_IN, st);
}

int whvi_layer_moments_f32(const float* x, int64_t x_sample_stride, const float* g, const float* s1, const float* s2,
                           const float* bias, float* sum_y, float* sum_y2, int64_t S, int64_t B, int64_t D, int flags,
                           whvi_stream_t stream)
{
    return whvi_layer_moments_add_f32(x, x_sample_stride, g, s1, s2, bias, nullptr, nullptr, sum_y, sum_y2, S, B, D, flags, stream);
}

int whvi_layer_moments_add_f32(const float* x, int64_t x_sample_stride, const float* g, const float* s1, const float* s2,
                               const float* bias, const float* in_sum_y, const float* in_sum_y2, float* sum_y, float* sum_y2,
                               int64_t S, int64_t B, int64_t D, int flags, whvi_stream_t stream)
{
    if (int rc = check_layer_shape("layer_moments", S, B, D, x_sample_stride, 32768)) return rc;
    if (D < 8192) return fail(WHVI_E_SHAPE, "layer_moments: D = %lld outside [8192, 32768] (use whvi_layer_fwd_fused_f32 + whvi_mc_moments_f32)", (long long)D);
    if (flags & ~(WHVI_LAYER_FROM_T2 | WHVI_LAYER_ACCUMULATE | 0xFF00)) return fail(WHVI_E_MODE, "layer_moments: unknown flags %d", flags);
    if ((flags & WHVI_LAYER_FROM_T2) && x_sample_stride != 0) return fail(WHVI_E_MODE, "layer_moments: FROM_T2 needs a shared (B, D) input");
    if (B == 0) return WHVI_OK;
    if (!x || !s1 || !s2 || !sum_y || (S > 0 && !g)) return fail(WHVI_E_NULL, "layer_moments: null pointer");
    if (!aligned16(x) || !aligned16(g) || !aligned16(s1) || !aligned16(s2) || !aligned16(bias) || !aligned16(sum_y) || !aligned16(sum_y2))
        return fail(WHVI_E_ALIGN, "layer_moments: pointers must be 16-byte aligned");
    if (!aligned16(in_sum_y) || !aligned16(in_sum_y2)) return fail(WHVI_E_ALIGN, "layer_moments: pointers must be 16-byte aligned");
    if (in_sum_y2 && !sum_y2) return fail(WHVI_E_NULL, "layer_moments: in_sum_y2 given without sum_y2");
    return launch_layer_moments(x, x_sample_stride, g, s1, s2, bias, sum_y, sum_y2, S, B, D, (flags & WHVI_LAYER_FROM_T2) ? 1 : 0,
                                (flags & WHVI_LAYER_ACCUMULATE) ? 1 : 0, (flags >> 8) & 0xFF, static_cast<cudaStream_t>(stream),
                                in_sum_y, in_sum_y2);
}

int whvi_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                  const float* lr_dev, const float* step_dev, float beta1, float beta2, float eps, float grad_scale,
                  whvi_stream_t stream)
{
    if (n < 0) return fail(WHVI_E_SHAPE, "adam: n=%lld", (long long)n);
    if (n == 0) return WHVI_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !step_dev) return fail(WHVI_E_NULL, "adam: null pointer");
    if (!(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f))
        return fail(WHVI_E_MODE, "adam: betas must be in [0, 1) and eps >= 0");
    return launch_adam(param, grad, exp_avg, exp_avg_sq, n, lr, lr_dev, step_dev, beta1, beta2, eps, grad_scale,
                       static_cast<cudaStream_t>(stream));
}

int whvi_mc_moments_f32(const float* y, float* sum_y, float* sum_y2, int64_t S, int64_t n, int accumulate,
                        whvi_stream_t stream)
{
    if (S < 0 || n < 0 || (n & 3)) return fail(WHVI_E_SHAPE, "mc_moments: S=%lld n=%lld (n must be a multiple of 4)", (long long)S, (long long)n);
    if (n == 0) return WHVI_OK;
    if (!sum_y || (S > 0 && !y)) return fail(WHVI_E_NULL, "mc_moments: null pointer");
    if (!aligned16(y) || !aligned16(sum_y) || !aligned16(sum_y2)) return fail(WHVI_E_ALIGN, "mc_moments: pointers must be 16-byte aligned");
    return launch_mc_moments(y, n, accumulate ? sum_y : nullptr, accumulate ? sum_y2 : nullptr, sum_y, sum_y2, S, n,
                             static_cast<cudaStream_t>(stream));
}

int whvi_mc_moments_strided_f32(const float* y, int64_t y_sample_stride, const float* in_sum_y, const float* in_sum_y2,
                                float* out_sum_y, float* out_sum_y2, int64_t S, int64_t n, whvi_stream_t stream)
{
    if (S < 0 || n < 0 || (n & 3) || (y_sample_stride & 3) || y_sample_stride < n)
        return fail(WHVI_E_SHAPE, "mc_moments_strided: S=%lld n=%lld stride=%lld (n, stride multiples of 4, stride >= n)",
                    (long long)S, (long long)n, (long long)y_sample_stride);
    if (n == 0) return WHVI_OK;
    if (!out_sum_y || (S > 0 && !y)) return fail(WHVI_E_NULL, "mc_moments_strided: null pointer");
    if (!aligned16(y) || !aligned16(in_sum_y) || !aligned16(in_sum_y2) || !aligned16(out_sum_y) || !aligned16(out_sum_y2))
        return fail(WHVI_E_ALIGN, "mc_moments_strided: pointers must be 16-byte aligned");
    return launch_mc_moments(y, y_sample_stride, in_sum_y, in_sum_y2, out_sum_y, out_sum_y2, S, n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
