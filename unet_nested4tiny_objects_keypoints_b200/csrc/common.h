// common.h — error reporting and device queries shared by the libunpp translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace unpp {

// Thread-local last-error text (DataParallel drives one host thread per GPU).
char* last_error_buf();
int fail(int code, const char* fmt, ...);
int fail_cuda(const char* what);  // formats cudaGetLastError() and returns UNPP_ERR_CUDA
int num_sms();

}  // namespace unpp
