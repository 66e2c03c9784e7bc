#include "weights.cuh"

namespace svdd {

__global__ void pack_conv_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                        int Cout, int Cin, int T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)Cout * Cin * T;
  if (i >= total) return;
  const int ci = (int)(i % Cin);
  const int co = (int)((i / Cin) % Cout);
  const int t = (int)(i / ((int64_t)Cin * Cout));
  out[i] = __float2bfloat16_rn(w[((int64_t)co * Cin + ci) * T + t]);
}

__global__ void copy_f32_kernel(const float* __restrict__ a, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i];
}

__global__ void fold_bn_kernel(const float* __restrict__ g, const float* __restrict__ b,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ conv_bias, float eps,
                               float* __restrict__ scale, float* __restrict__ shift, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = g[i] / sqrtf(var[i] + eps);
  const float cb = conv_bias ? conv_bias[i] : 0.0f;
  scale[i] = s;
  shift[i] = (cb - mean[i]) * s + b[i];
}

int pack_conv_weight(const float* w, __nv_bfloat16* out, int Cout, int Cin, int T, cudaStream_t st) {
  const int64_t total = (int64_t)Cout * Cin * T;
  pack_conv_weight_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, st>>>(w, out, Cout, Cin, T);
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}

int copy_f32(const float* a, float* out, int64_t n, cudaStream_t st) {
  copy_f32_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, st>>>(a, out, n);
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}

int fold_bn(const float* g, const float* b, const float* mean, const float* var,
            const float* conv_bias, float eps, float* scale, float* shift, int n, cudaStream_t st) {
  fold_bn_kernel<<<ceil_div(n, 256), 256, 0, st>>>(g, b, mean, var, conv_bias, eps, scale, shift, n);
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}

}  // namespace svdd
