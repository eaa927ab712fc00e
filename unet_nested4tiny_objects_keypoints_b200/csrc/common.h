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
int current_device();  // ordinal of the calling thread's device, clamped to [0, 63]
int fail_cuda_err(const char* what, cudaError_t e);  // formats e and returns UNPP_ERR_CUDA
bool pdl_enabled();  // programmatic dependent launch for every libunpp kernel (UNPP_PDL=1 switches it on)

// With UNPP_PDL=1 every libunpp kernel is launched with the programmatic-stream-serialization attribute and runs
//   [on-chip prologue]  pdl_wait();  pdl_trigger();  [work]
// so the launch latency and the prologue (barrier init, TMEM allocation, descriptor prefetch) of kernel k+1
// overlap the tail of kernel k.  pdl_wait() returns when the preceding kernel has COMPLETED and its writes are
// visible; since every kernel passes its own wait before it triggers, everything older is complete as well.
// Nothing before pdl_wait() may touch global memory other kernels of the stream write.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE property of a kernel: in-process multi-GPU use
// (nn.DataParallel) must opt in once on every device.  `done` is the kernel's own static flag array.
template <typename K>
inline cudaError_t opt_in_smem(K kernel, int bytes, unsigned char (&done)[64]) {
  const int dev = current_device();
  if (done[dev]) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) done[dev] = 1;  // idempotent: a race between replicas' threads only repeats the call
  return e;
}

template <typename... P, typename... A>
inline cudaError_t launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
#endif

}  // namespace unpp
