// Host-side utilities shared by the C-ABI entry points: thread-local error string, device
// attribute cache, and the TMA tensor-map encoder (resolved through the runtime so that the
// library has no link-time dependency on libcuda and loads on a GPU-less build host).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <unordered_map>

#include "../../include/vist3a_sm100.h"

namespace v3a {

char* last_error_buf();   // thread-local, 512 bytes (defined in capi.cu)
int set_error(int code, const char* fmt, ...);
int num_sms();            // SM count of the current device (cached per device)
int check_arch();         // VIST3A_OK if the current device is sm_100, else VIST3A_ERR_ARCH
std::atomic<long long>& launch_counter();  // kernels launched by this library in this process

bool pdl_enabled();        // programmatic dependent launch on the hot kernels (vist3a_set_pdl / env VIST3A_PDL, default on)
int set_pdl(int enable);   // returns the previous setting

// Launch `kern` on `st`; with pdl (and PDL enabled) the launch carries the programmatic-stream-serialization attribute, so the
// kernel's prologue overlaps the tail of its predecessor.  Only kernels that execute pdl_wait() before their first global-memory
// access may be launched with pdl = true.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                                 int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl && pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Opt a kernel in to `bytes` of dynamic shared memory.  The attribute is per device (context), so the "already done" record is a
// per-device bit (devices >= 64: set every time); concurrent first calls from several threads set it twice, which is harmless.
template <typename K>
inline cudaError_t ensure_dynamic_smem(K kern, int bytes, std::atomic<unsigned long long>& done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const unsigned long long bit = dev < 64 ? (1ull << dev) : 0ull;
  if (bit && (done.load(std::memory_order_acquire) & bit)) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && bit) done.fetch_or(bit, std::memory_order_release);
  return e;
}

#define V3A_CUDA_OK(expr)                                                                             \
  do {                                                                                                \
    cudaError_t _e = (expr);                                                                          \
    if (_e != cudaSuccess) return v3a::set_error(VIST3A_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

#define V3A_REQUIRE(cond, code, ...)                        \
  do {                                                      \
    if (!(cond)) return v3a::set_error(code, __VA_ARGS__);  \
  } while (0)

// Encode a tiled tensor map with 128-byte swizzle (or no swizzle if swizzle128 == false).
//   rank <= 5; dims[0] is the contiguous dimension; strides_bytes[i] is the stride of dim i+1.
int encode_tensor_map(CUtensorMap* out, const void* base, int elem_bytes, bool is_float32, int rank,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, bool swizzle128);

}  // namespace v3a
