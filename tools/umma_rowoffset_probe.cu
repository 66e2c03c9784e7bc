// Hardware probe: does a tcgen05 K-major SWIZZLE_128B shared-memory descriptor address
// rows correctly when its start address is offset by a whole number of 128-byte rows that
// is NOT a multiple of 8 (i.e. not aligned to the 1024-byte swizzle atom)?
//
// A [160, 64] bf16 is TMA-loaded (136 rows, 128B swizzle) into 1024-aligned shared memory;
// for r0 = 0..8 one 128x128x64 MMA reads rows r0 .. r0+127 through a descriptor whose start
// address is base + r0*128, and the result is compared with the CPU product.  Variant 1 also
// sets the descriptor's base-offset field to (start >> 7) & 7.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -o build/umma_probe tools/umma_rowoffset_probe.cu
//   gpurun -- ./build/umma_probe
#include <cuda_bf16.h>
#include <stdlib.h>

#include <vector>

#include "../svdd_b200/csrc/ptx_sm100.cuh"

using namespace svdd;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kRowsA = 160, kBoxA = 136, kN = 128, kK = 64;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* C, int r0,
             int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                    // 136 rows x 128 B = 17408 B -> pad to 18432
  uint8_t* sB = smem + 18432;            // 128 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 18432 + 16384);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bars[0], kBoxA * 128 + kN * 128);
    ptx::tma_load_2d(sA, &tmA, &bars[0], 0, 0);
    ptx::tma_load_2d(sB, &tmB, &bars[0], 0, 0);
    ptx::mbar_wait(&bars[0], 0);
    ptx::tc_fence_after();
    const uint32_t a_addr = ptx::smem_u32(sA) + (uint32_t)r0 * 128u;
    uint64_t da = ptx::make_kmajor_sw128_desc(a_addr);
    if (variant == 1) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(sB));
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, kN);
#pragma unroll
    for (int k = 0; k < kK / 16; ++k) ptx::umma_bf16(tmem_base, da + 2 * k, db + 2 * k, idesc, k > 0);
    ptx::umma_commit(&bars[1]);
  }
  __syncwarp();
  ptx::mbar_wait(&bars[1], 0);
  ptx::tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c = 0; c < kN / 32; ++c) {
    uint32_t raw[32];
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c * 32, raw);
    ptx::tmem_ld_wait();
    for (int i = 0; i < 32; ++i) C[(size_t)row * kN + c * 32 + i] = __uint_as_float(raw[i]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
  std::vector<__nv_bfloat16> hA((size_t)kRowsA * kK), hB((size_t)kN * kK);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 9 - 4); hA[i] = __float2bfloat16(fA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 7 - 3); hB[i] = __float2bfloat16(fB[i]); }
  __nv_bfloat16 *dA, *dB;
  float* dC;
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dC, (size_t)128 * kN * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tA, tB;
  cuuint32_t es[2] = {1, 1};
  {
    cuuint64_t dims[2] = {kK, kRowsA};
    cuuint64_t str[1] = {kK * 2};
    cuuint32_t box[2] = {64, kBoxA};
    CUresult r = enc(&tA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {kK, kN};
    cuuint64_t str[1] = {kK * 2};
    cuuint32_t box[2] = {64, kN};
    CUresult r = enc(&tB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  const int smem = 18432 + 16384 + 256 + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> hC((size_t)128 * kN);
  for (int variant = 0; variant < 2; ++variant) {
    for (int r0 = 0; r0 <= 8; ++r0) {
      CK(cudaMemset(dC, 0, hC.size() * 4));
      probe_kernel<<<1, 128, smem>>>(tA, tB, dC, r0, variant);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d r0 %d: kernel failed: %s\n", variant, r0, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
      int bad = 0, bad_rows = 0;
      for (int i = 0; i < 128; ++i) {
        int row_bad = 0;
        for (int n = 0; n < kN; ++n) {
          float ref = 0.0f;
          for (int k = 0; k < kK; ++k) ref += fA[(size_t)(r0 + i) * kK + k] * fB[(size_t)n * kK + k];
          if (ref != hC[(size_t)i * kN + n]) { ++bad; row_bad = 1; }
        }
        bad_rows += row_bad;
      }
      printf("variant %d (base_offset %s) r0 %d: %s (%d wrong elements in %d rows)\n", variant,
             variant ? "set" : "0", r0, bad ? "MISMATCH" : "exact", bad, bad_rows);
    }
  }
  return 0;
}
