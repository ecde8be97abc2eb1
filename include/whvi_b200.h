/*
 * whvi_b200 -- C ABI of the B200-native WHVI hot path (libwhvi_b200.so, sm_100a).
 *
 * This is the drop-in boundary: the entry points below are what the reference's
 * Python layer binds instead of its two torch C++/CUDA extensions.  Every function
 *   - takes plain device pointers and sizes (no torch types),
 *   - is stream-ordered on the cudaStream_t passed as `stream` (NULL = legacy default
 *     stream), never synchronises, never allocates or frees device memory (the caller
 *     owns every buffer, including workspaces sized by the *_workspace_bytes calls),
 *   - returns 0 on success, a negative WHVI_E_* code for an invalid argument, or a
 *     positive cudaError_t for a CUDA failure (launch errors ARE checked, unlike the
 *     reference, src/fwht/cuda/fwht_cuda_kernel.cu:169-178); whvi_last_error() gives the
 *     thread-local message.  The Python side turns non-zero into RuntimeError, which is
 *     what the reference's TORCH_CHECKs raise (src/fwht/cuda/fwht_cuda.cpp:6-10).
 * All matrices are row-major contiguous fp32; device pointers must be 16-byte aligned.
 *
 * Reference interfaces replaced (file:line in ltdung/WHVI):
 *   whvi_fwht_f32            fwht_cuda.fwht            src/fwht/cuda/fwht_cuda.cpp:5-18
 *                            fwht_cuda_frontend        src/fwht/cuda/fwht_cuda_kernel.cu:156-181
 *                            fwht_cpp.forward/backward src/fwht/cpp/fwht.cpp:23-34
 *   whvi_layer_fwd_f32       WHVISquarePow2Matrix.sample_lrt / w_bar  src/weights.py:66-93
 *   whvi_layer_bwd_f32       autograd of the above (FWHTFunction.backward, src/fwht/cuda/fwht.py:14-16)
 *   whvi_reparam_*_f32       g_sigma + reparameterisation               src/weights.py:43-50, :82-83, :92-93
 *   whvi_kl_f32              WHVISquarePow2Matrix.kl -> kl_diag_normal  src/weights.py:52-64, src/utils.py:49-71
 */
#ifndef WHVI_B200_H
#define WHVI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WHVI_ABI_VERSION 1

#if defined(WHVI_BUILDING) && defined(__GNUC__)
#define WHVI_API __attribute__((visibility("default")))
#else
#define WHVI_API
#endif

#define WHVI_OK 0
#define WHVI_E_NULL (-1)      /* a required pointer is NULL                      */
#define WHVI_E_SHAPE (-2)     /* negative size, D not a power of two, D too big  */
#define WHVI_E_ALIGN (-3)     /* pointer not 16-byte aligned                     */
#define WHVI_E_MODE (-4)      /* unknown mode / flag                             */
#define WHVI_E_WORKSPACE (-5) /* workspace missing or too small                  */

typedef void* whvi_stream_t; /* a cudaStream_t */

/* Library / ABI identification. */
WHVI_API int whvi_abi_version(void);
/* Message for the last non-zero return on this thread ("" if none). */
WHVI_API const char* whvi_last_error(void);
/* Largest D (power of two) the FWHT and the fused layer kernels accept. */
WHVI_API int64_t whvi_max_dim(void);

/*
 * Batched fast Walsh-Hadamard transform, natural (Sylvester) order, unnormalised:
 *   out[r, :] = H_D . in[r, :]      for r in [0, rows)
 * D is a power of two, 1 <= D <= whvi_max_dim().  in == out (in place) is allowed;
 * partial overlap is not.  rows == 0 is a no-op.  fwd and bwd are the same call
 * (H is symmetric: src/fwht/cuda/fwht.py:14-16).
 */
WHVI_API int whvi_fwht_f32(const float* in, float* out, int64_t rows, int64_t D, whvi_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* WHVI_B200_H */
