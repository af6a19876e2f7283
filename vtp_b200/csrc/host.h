// vtp_b200 — host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vtp_b200.h"

namespace vtp {

// last error string, thread-local (returned by vtp_last_error())
char* err_buf();
#define VTP_FAIL(code, ...)                         \
    do {                                            \
        snprintf(vtp::err_buf(), 512, __VA_ARGS__); \
        return (code);                              \
    } while (0)
#define VTP_CHECK_ARG(cond, ...) \
    do {                         \
        if (!(cond)) VTP_FAIL(VTP_ERR_ARG, __VA_ARGS__); \
    } while (0)
#define VTP_CUDA(expr)                                                                               \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) VTP_FAIL(VTP_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)
#define VTP_LAUNCH_CHECK()                                                                        \
    do {                                                                                          \
        cudaError_t e__ = cudaGetLastError();                                                     \
        if (e__ != cudaSuccess) VTP_FAIL(VTP_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(e__)); \
    } while (0)

int num_sms();

// Encode a row-major bf16 tensor map: dims[0] is the contiguous dim. strides_bytes[i] is the stride of dims[i+1].
// box[] in elements. 128B swizzle (box[0]*2 bytes must be 128).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box);
// same for VTP_F32 | VTP_BF16 elements (box[0] * element size must be 128 bytes)
int make_tmap(CUtensorMap* out, const void* base, int dtype, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace vtp
