// Library-level entry points: version, error string, device check, launch counter.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace svdd {

static thread_local char g_last_error[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SVDD_PDL"); on = e ? (atoi(e) != 0) : 0; }
  return on == 1;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool debug_dump_enabled() {
  static int on = -1;
  if (on < 0) on = getenv("SVDD_DEBUG_DUMP_DIR") != nullptr ? 1 : 0;
  return on == 1;
}

void debug_dump(const char* name, const void* dev_ptr, size_t bytes, cudaStream_t st) {
  if (!debug_dump_enabled()) return;
  std::vector<char> host(bytes);
  cudaStreamSynchronize(st);
  if (cudaMemcpy(host.data(), dev_ptr, bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return;
  const std::string path = std::string(getenv("SVDD_DEBUG_DUMP_DIR")) + "/" + name + ".bin";
  if (FILE* f = fopen(path.c_str(), "wb")) {
    fwrite(host.data(), 1, bytes, f);
    fclose(f);
  }
}

}  // namespace svdd

extern "C" int svdd_version(void) { return SVDD_B200_VERSION; }

extern "C" const char* svdd_last_error(void) { return svdd::g_last_error; }

extern "C" int64_t svdd_launch_count(void) {
  return svdd::g_launches.load(std::memory_order_relaxed);
}

extern "C" int svdd_device_check(int device) {
  static int cached_dev = -1, cached_rc = 0;
  if (device == cached_dev) return cached_rc;
  int major = 0, minor = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
  if (e != cudaSuccess) {
    svdd::set_last_error("svdd_device_check(%d): %s", device, cudaGetErrorString(e));
    cudaGetLastError();
    return SVDD_ERR_CUDA;
  }
  int rc = SVDD_OK;
  if (major != 10) {
    svdd::set_last_error(
        "device %d is sm_%d%d; libsvdd_b200 is built for sm_100a only and has no fallback",
        device, major, minor);
    rc = SVDD_ERR_UNSUPPORTED_DEVICE;
  }
  cached_dev = device;
  cached_rc = rc;
  return rc;
}
