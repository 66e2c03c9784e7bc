// Weight ingestion shared by the network handles: lookup of reference-layout
// state_dict entries by name, device-side repacking to the layouts the kernels
// want (bf16, tap-major, K contiguous), and a tiny bump allocator so a handle
// owns exactly one device allocation.
#pragma once
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace svdd {

struct TensorTable {
  const svdd_tensor* t;
  int n;
  const svdd_tensor* find(const std::string& name) const {
    for (int i = 0; i < n; ++i)
      if (t[i].name != nullptr && name == t[i].name) return &t[i];
    return nullptr;
  }
  bool has(const std::string& name) const { return find(name) != nullptr; }
  // returns nullptr (and sets the error) unless present with exactly `numel` elements
  const float* get(const std::string& name, int64_t numel) const {
    const svdd_tensor* x = find(name);
    if (x == nullptr) {
      set_last_error("missing tensor '%s'", name.c_str());
      return nullptr;
    }
    int64_t ne = 1;
    for (int d = 0; d < x->ndim; ++d) ne *= x->shape[d];
    if (ne != numel) {
      set_last_error("tensor '%s' has %lld elements, expected %lld", name.c_str(), (long long)ne,
                     (long long)numel);
      return nullptr;
    }
    return x->data;
  }
  int64_t dim(const std::string& name, int d) const {
    const svdd_tensor* x = find(name);
    return (x && d < x->ndim) ? x->shape[d] : -1;
  }
};

#define SVDD_GET(var, table, name, numel)                 \
  const float* var = (table).get((name), (numel));        \
  if (var == nullptr) return SVDD_ERR_MISSING_TENSOR

// One cudaMalloc per handle; sub-allocations are 256-byte aligned.
class DeviceArena {
 public:
  ~DeviceArena() { release(); }
  void reserve(size_t bytes) { total_ += align(bytes); }
  int commit() {
    if (cudaMalloc(&base_, total_ > 0 ? total_ : 256) != cudaSuccess) {
      set_last_error("cudaMalloc of %zu bytes for packed weights failed", total_);
      cudaGetLastError();
      return SVDD_ERR_CUDA;
    }
    return SVDD_OK;
  }
  template <typename T>
  T* take(size_t count) {
    T* p = reinterpret_cast<T*>(reinterpret_cast<uint8_t*>(base_) + used_);
    used_ += align(count * sizeof(T));
    return p;
  }
  void release() {
    if (base_) cudaFree(base_);
    base_ = nullptr;
  }
  size_t bytes() const { return total_; }
  static size_t align(size_t b) { return (b + 255) & ~(size_t)255; }

 private:
  void* base_ = nullptr;
  size_t total_ = 0, used_ = 0;
};

// Workspace carving on the hot path (caller-provided scratch).
class Workspace {
 public:
  Workspace(void* base, size_t bytes) : base_(reinterpret_cast<uint8_t*>(base)), bytes_(bytes) {}
  template <typename T>
  T* take(size_t count) {
    const size_t off = used_;
    used_ += DeviceArena::align(count * sizeof(T));
    return (used_ <= bytes_ && base_ != nullptr) ? reinterpret_cast<T*>(base_ + off) : nullptr;
  }
  size_t used() const { return used_; }
  bool ok() const { return used_ <= bytes_; }

 private:
  uint8_t* base_;
  size_t bytes_, used_ = 0;
};

// ---- packing kernels ----------------------------------------------------------
// conv weight [Cout, Cin, T] fp32 -> bf16 [T, Cout, Cin]   (linear: T = 1)
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                        int Cout, int Cin, int T);
// out[i] = a[i] (fp32 copy into handle-owned memory)
__global__ void copy_f32_kernel(const float* __restrict__ a, float* __restrict__ out, int64_t n);
// BatchNorm (eval) folded to y = x*scale + shift, optionally with a preceding conv bias:
//   scale = g / sqrt(var + eps);  shift = (conv_bias - mean) * scale + beta
__global__ void fold_bn_kernel(const float* __restrict__ g, const float* __restrict__ b,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ conv_bias, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int n);

int pack_conv_weight(const float* w, __nv_bfloat16* out, int Cout, int Cin, int T, cudaStream_t st);
int copy_f32(const float* a, float* out, int64_t n, cudaStream_t st);
int fold_bn(const float* g, const float* b, const float* mean, const float* var,
            const float* conv_bias, float eps, float* scale, float* shift, int n, cudaStream_t st);

}  // namespace svdd
