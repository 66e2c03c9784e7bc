// Shared host/device helpers for libsvdd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/svdd_b200.h"

namespace svdd {

// ---- error plumbing (thread-local last error, negative return codes) -------
void set_last_error(const char* fmt, ...);

#define SVDD_CHECK_ARG(cond, ...)                                   \
  do {                                                              \
    if (!(cond)) {                                                  \
      ::svdd::set_last_error(__VA_ARGS__);                          \
      return SVDD_ERR_INVALID_ARGUMENT;                             \
    }                                                               \
  } while (0)

#define SVDD_CUDA(expr)                                                      \
  do {                                                                       \
    cudaError_t _e = (expr);                                                 \
    if (_e != cudaSuccess) {                                                 \
      ::svdd::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,   \
                             cudaGetErrorString(_e));                        \
      return SVDD_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

#define SVDD_LAUNCH_CHECK()                                                  \
  do {                                                                       \
    cudaError_t _e = cudaGetLastError();                                     \
    if (_e != cudaSuccess) {                                                 \
      ::svdd::set_last_error("%s:%d: kernel launch -> %s", __FILE__,         \
                             __LINE__, cudaGetErrorString(_e));              \
      return SVDD_ERR_CUDA;                                                  \
    }                                                                        \
  } while (0)

#define SVDD_TRY(expr)            \
  do {                            \
    int _rc = (expr);             \
    if (_rc != SVDD_OK) return _rc; \
  } while (0)

// counts kernel launches issued by this library (bench.py's gpu_launches)
void count_launch(int n = 1);

// Debug aid: when SVDD_DEBUG_DUMP_DIR is set, writes `bytes` of device memory to
// <dir>/<name>.bin after synchronising the stream.  Never active on the hot path.
void debug_dump(const char* name, const void* dev_ptr, size_t bytes, cudaStream_t st);
bool debug_dump_enabled();

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ---- programmatic dependent launch (PDL) ------------------------------------------------
// With SVDD_PDL=1 every kernel of a reverse step is launched with programmatic stream
// serialisation: its CTAs may become resident (barrier init, TMEM allocation, tensor-map prefetch) while the
// previous kernel drains, and block in pdl_wait() until that kernel's memory is visible.
// Rule: a kernel launched through launch_k() calls pdl_wait() before its first access to
// global memory another kernel may have written (or may still be reading).
// Off by default (SVDD_PDL=1 enables the launch attribute; pdl_wait() is a no-op without it):
// inside the captured CUDA graph the kernel-to-kernel gap is already hidden and the A/B on
// the c2 step showed no gain (179.9 vs 177.7 seq/s).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            int cluster_x, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<P>(args)...);
}

constexpr int kMaskIndex = 4;            // diffusion_gosai.py:85,94-95
constexpr int kVocab = 5;                // A,C,G,T,MASK
constexpr float kNegInfinity = -1000000.0f;  // diffusion_gosai.py:156

template <typename T>
__host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }

// Token tensors cross the ABI as int64 (the reference's dtype) or uint8 (the
// engine's internal state).  Device-side accessors:
template <typename Tok>
__device__ __forceinline__ int load_tok(const Tok* p, size_t i) { return (int)p[i]; }
template <typename Tok>
__device__ __forceinline__ void store_tok(Tok* p, size_t i, int v) { p[i] = (Tok)v; }

}  // namespace svdd
