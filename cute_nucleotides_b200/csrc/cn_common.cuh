// cn_common.cuh -- error reporting and small helpers shared by the pieces of the C-ABI translation unit.
// Everything here lives in the anonymous namespace of cn_api.cu (single translation unit).
#pragma once
#include "../../include/cute_nucleotides_cuda.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

namespace {

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
thread_local char t_err[512] = "";
const char kPanicText[] = "The length is greater than the number of nucleotides!";   // n_to_bits.rs:53

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof t_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CN_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(CN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

std::atomic<uint64_t> g_launches{0};


inline uintptr_t addr(const void *p) { return reinterpret_cast<uintptr_t>(p); }

}  // namespace
