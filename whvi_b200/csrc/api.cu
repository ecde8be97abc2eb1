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

int64_t whvi_max_dim(void) { return int64_t(1) << kMaxLog2D; }

int whvi_fwht_f32(const float* in, float* out, int64_t rows, int64_t D, whvi_stream_t stream)
{
    if (rows < 0 || D < 1) return fail(WHVI_E_SHAPE, "fwht: rows=%lld D=%lld", (long long)rows, (long long)D);
    if (!is_pow2(D)) return fail(WHVI_E_SHAPE, "fwht: n must be a power of 2 (got %lld)", (long long)D);
    if (rows == 0) return WHVI_OK;
    if (!in || !out) return fail(WHVI_E_NULL, "fwht: null pointer");
    if (!aligned16(in) || !aligned16(out)) return fail(WHVI_E_ALIGN, "fwht: pointers must be 16-byte aligned");
    return launch_fwht(in, out, rows, D, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
