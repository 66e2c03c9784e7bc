// Second-generation tcgen05 implicit GEMM (same contraction as conv_gemm.cuh):
//
//   * cta_group::2: a CTA pair (two SMs of one TPC) computes a 256 x BN tile with one
//     tcgen05.mma; each CTA stages its own 128 rows of A and HALF of the B tile, so the
//     L2 -> SM operand traffic per FLOP drops by a third against the 128 x 256 single-CTA
//     tile (the k5 convs were L2-feed bound);
//   * the epilogue no longer touches global memory from the math threads: accumulators go
//     TMEM -> registers -> 128B-swizzled shared-memory slabs (128 rows x 128 bytes) that two
//     dedicated store warps push out with TMA (cp.async.bulk.tensor store), and the
//     residual / pooled operand comes in the same way (TMA load into the slab, then read
//     by the thread that owns the row).  Thread-per-row global accesses cost 32 L1
//     wavefronts per instruction and made every K <= 1536 GEMM epilogue bound.
//
// Attention pooling over position pairs (enformer_pytorch AttentionPool, softmax over a pair)
// only needs the DIFFERENCE of the two logits, and Wp.y1 - Wp.y0 = Wp.(y1 - y0).  So:
//   EPI_PAIR   (the residual 1x1 conv feeding a pool): the epilogue exchanges each row with
//              its pair partner (adjacent lane) and stores y0 = y[2j] and yd = y[2j+1] - y[2j]
//              at half length instead of y (yd = 0 for the unpaired tail of an odd length);
//   EPI_POOL2  one GEMM over yd (half the rows, one accumulator) instead of two over y:
//              pooled = y0 + sigmoid(Wp.yd) * yd, then the next layer's BN+GELU.
// Half the pooling FLOPs of the two-accumulator EPI_POOL kernel, on the cta_group::2 path.
//
// Warp roles (384 threads): 0 TMA producer, 1 MMA issuer (leader CTA only), 2-9 epilogue
// math (thread = (row, column half)), 10-11 slab store / prefetch (one per column half).
//
// Slab protocol, per column half h, job j = 0,1,2,... in a fixed order (tile, slab, output):
//   buffer b = j & 1;  rin[h][b]  : store warp -> math warps  "buffer free / residual landed"
//                      rout[h][b] : math warps -> store warp  "slab written"
// The store warp provisions job j+2 right after job j's TMA store has drained the buffer.
//
// EPI_PAIR / EPI_POOL2 take 32-column slabs instead (ep.slab32: 128 rows x 64 B, FOUR buffers per
// column half in the same 64 KB, b = j & 3; by default with 16 epilogue warps, thread = (row, 16 columns),
// the eight warps of a column half sharing its slabs): a step consumes two jobs (an input and
// an output slab, or two inputs), and with two buffers per half the loads of step k+1 could only be
// issued once step k's math had released them -- ncu showed the math warps waiting for their input
// slabs ~35 % of the time (profiles/r02_summary.md).  With four, step k+1's inputs land during step k.
// The slabs are dense 64-byte rows under the 64B swizzle (tensor maps with SWIZZLE_64B): unit u (16 B)
// of row r sits at r * 64 + ((u ^ ((r >> 1) & 3)) << 4); eight consecutive rows of one unit cover all
// 32 banks once.
#pragma once
#include "conv_gemm.cuh"

namespace svdd {
namespace gemm2 {

using gemm_detail::kBK;
using gemm_detail::kBM;
using gemm_detail::kEpiWarps;
using gemm_detail::kEpiThreads;
using gemm_detail::kParamVecs;
using gemm_detail::P_BIAS;
using gemm_detail::P_SCALE;
using gemm_detail::P_SHIFT;
using gemm_detail::P_SCALE2;
using gemm_detail::P_SHIFT2;
using gemm_detail::P_HEADW;

constexpr int kStoreWarps = 2;
constexpr int kThreads = 64 + kEpiThreads + 32 * kStoreWarps;   // 384
constexpr int kSlabBytes = kBM * 128;                            // 128 rows x 128 B
constexpr int kStagingBytes = 2 * 2 * kSlabBytes;                // [half][buffer]
constexpr int kBarBytes = 512;
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;   // shared::cluster address of the pair's even CTA

template <int BN, int CG>
struct Cfg2 {
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = (BN / CG) * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kParamBytes = 2 * kParamVecs * BN * 4;
  static constexpr int kAvail = 227 * 1024 - 1024 - kStagingBytes - kParamBytes - kBarBytes;
  static constexpr int kStagesRaw = kAvail / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemCols = 2 * BN <= 32 ? 32 : 2 * BN <= 64 ? 64 : 2 * BN <= 128 ? 128
                                   : 2 * BN <= 256 ? 256 : 512;
  static constexpr int kHalf = BN / 2;
  static constexpr int kChunks = kHalf / 32;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kParamBytes + kBarBytes + 1024;
  // HALO variant (flat-row conv, the A tile is loaded ONCE per K block with taps-1 extra rows and
  // every tap is a row offset of the UMMA descriptor): separate A and B rings
  static constexpr int kHaloRowsMax = 8;                              // taps - 1 <= 8
  static constexpr int kAHBytes = (kBM + kHaloRowsMax) * kBK * 2;     // 136 rows x 128 B = 17 x 1024
  static constexpr int kStagesAH = 3;
  static constexpr int kStagesBHRaw = (kAvail - kStagesAH * kAHBytes) / kBBytes;
  static constexpr int kStagesBH = kStagesBHRaw > 8 ? 8 : kStagesBHRaw;
  static constexpr int kRingBytesH = kStagesAH * kAHBytes + kStagesBH * kBBytes;
  static constexpr int kSmemBytesH = kRingBytesH + kStagingBytes + kParamBytes + kBarBytes + 1024;
  static_assert(kAHBytes % 1024 == 0 && kBBytes % 1024 == 0, "ring stages must stay 1024-byte aligned");
  static_assert(CG == 1 || kStagesBH >= 4, "halo weight ring too shallow");
  static_assert(kSmemBytesH <= 227 * 1024, "shared memory budget exceeded (halo)");
  static_assert(BN == 128 || BN == 256, "gemm2 tiles are 128 or 256 columns wide");
  static_assert(CG == 1 || CG == 2, "cta_group is 1 or 2");
  static_assert(kStages >= 2, "operand ring too shallow");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

// ---- PTX helpers that only this kernel needs --------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* dst_smem, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)),
                 "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)),
                 "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// operand loads: with CG == 2 both CTAs signal the LEADER's barrier
template <int CG>
__device__ __forceinline__ void tma_load_2d_cg(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  if constexpr (CG == 1) {
    ptx::tma_load_2d(dst, m, bar, c0, c1);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar) & kPeerMask),
        "r"(c0), "r"(c1) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_load_3d_cg(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               int c2) {
  if constexpr (CG == 1) {
    ptx::tma_load_3d(dst, m, bar, c0, c1, c2);
  } else {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ptx::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar) & kPeerMask),
        "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_bf16_cg(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                             uint32_t accumulate) {
  if constexpr (CG == 1) {
    ptx::umma_bf16(tmem_d, da, db, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
  }
}
// the same with both descriptors given as their low words (high word = ptx::kKmajorSw128DescHi)
template <int CG>
__device__ __forceinline__ void umma_bf16_lo_cg(uint32_t tmem_d, uint32_t da_lo, uint32_t db_lo, uint32_t idesc,
                                                uint32_t accumulate) {
  if constexpr (CG == 1) {
    ptx::umma_bf16_lo(tmem_d, da_lo, db_lo, idesc, accumulate);
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(da_lo), "r"(db_lo), "r"(idesc), "r"(accumulate), "r"(ptx::kKmajorSw128DescHi)
        : "memory");
  }
}
// arrives (once all prior MMAs of this thread retire) on the barrier at this smem offset in
// every CTA of the pair
template <int CG>
__device__ __forceinline__ void umma_commit_cg(uint64_t* bar) {
  if constexpr (CG == 1) {
    ptx::umma_commit(bar);
  } else {
    const uint16_t mask = 3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(ptx::smem_u32(bar)), "h"(mask) : "memory");
  }
}
// arrive on the pair leader's copy of `bar` (own copy when CG == 1)
template <int CG>
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  if constexpr (CG == 1) {
    ptx::mbar_arrive(bar);
  } else {
    // relaxed: the accumulator reads it publishes are complete (tcgen05.wait::ld) and fenced
    // (tcgen05.fence::before_thread_sync) by the caller; a release at cluster scope lowers to
    // MEMBAR.ALL.GPU, 6.6 % of the EPI_PAIR kernel's stall samples
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(ptx::smem_u32(bar) & kPeerMask)
                 : "memory");
  }
}
// TMA store of a 3-D box from shared memory (bulk async group)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// the same box, accumulated into global memory (fp32 add performed by the memory system)
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// x * sigmoid(1.702 x) = 0.5 x (1 + tanh(0.851 x)): one MUFU op per element
__device__ __forceinline__ float gelu_tanh(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * v));
  const float h = 0.5f * v;
  return fmaf(h, t, h);
}
__device__ __forceinline__ void act32(float* v, int act) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
  } else if (act == ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gemm_detail::gelu_tanh_approx(v[i]);
  }
}

// One row's 32-column chunk <-> its place in a 128B-swizzled slab.  `row_base` = slab + r*128,
// x7 = r & 7.  bf16: the chunk is half a row (16B units 4*(c&1) .. +3); fp32: the whole row.
__device__ __forceinline__ void slab_write(uint8_t* row_base, int x7, bool f32, int chunk_in_slab, const float* v) {
  if (f32) {
#pragma unroll
    for (int u = 0; u < 8; ++u)
      ptx::sts128f(row_base + ((u ^ x7) << 4), make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]));
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = chunk_in_slab * 4 + u;
      ptx::sts128(row_base + ((j ^ x7) << 4),
          make_uint4(gemm_detail::pack_bf16x2(v[8 * u], v[8 * u + 1]), gemm_detail::pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                     gemm_detail::pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), gemm_detail::pack_bf16x2(v[8 * u + 6], v[8 * u + 7])));
    }
  }
}
// 32-column bf16 slabs (64-byte rows, 64B swizzle): pair_base = slab + (r >> 1) * 128, hi = (r & 1) << 2,
// x7 = (r >> 1) & 3
__device__ __forceinline__ void slab32_write(uint8_t* pair_base, int hi, int x7, const float* v) {
#pragma unroll
  for (int u = 0; u < 4; ++u)
    ptx::sts128(pair_base + (((hi | u) ^ x7) << 4),
        make_uint4(gemm_detail::pack_bf16x2(v[8 * u], v[8 * u + 1]), gemm_detail::pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                   gemm_detail::pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), gemm_detail::pack_bf16x2(v[8 * u + 6], v[8 * u + 7])));
}
__device__ __forceinline__ void slab32_read(const uint8_t* pair_base, int hi, int x7, float* v) {
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const uint4 q = ptx::lds128(pair_base + (((hi | u) ^ x7) << 4));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
      v[8 * u + 2 * k] = __low2float(hh);
      v[8 * u + 2 * k + 1] = __high2float(hh);
    }
  }
}
__device__ __forceinline__ void slab_read(const uint8_t* row_base, int x7, bool f32, int chunk_in_slab, float* v) {
  if (f32) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float4 f = ptx::lds128f(row_base + ((u ^ x7) << 4));
      v[4 * u] = f.x; v[4 * u + 1] = f.y; v[4 * u + 2] = f.z; v[4 * u + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = chunk_in_slab * 4 + u;
      const uint4 q = ptx::lds128(row_base + ((j ^ x7) << 4));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
        v[8 * u + 2 * k] = __low2float(hh);
        v[8 * u + 2 * k + 1] = __high2float(hh);
      }
    }
  }
}

struct TileCoord { int n0, s0, l0; bool in_range; };

// 16-column TMEM chunk (the EW = 16 epilogue keeps the kernel inside 102 registers)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 16 bf16 columns (chunk c of 4) of one row <-> its place in a 128B-swizzled slab row
__device__ __forceinline__ void slab_write16(uint8_t* row_base, int x7, int c, const float* v) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = c * 2 + u;
    ptx::sts128(row_base + ((j ^ x7) << 4),
        make_uint4(gemm_detail::pack_bf16x2(v[8 * u], v[8 * u + 1]), gemm_detail::pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                   gemm_detail::pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), gemm_detail::pack_bf16x2(v[8 * u + 6], v[8 * u + 7])));
  }
}
__device__ __forceinline__ void slab_read16(const uint8_t* row_base, int x7, int c, float* v) {
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int j = c * 2 + u;
    const uint4 q = ptx::lds128(row_base + ((j ^ x7) << 4));
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
      v[8 * u + 2 * k] = __low2float(hh);
      v[8 * u + 2 * k + 1] = __high2float(hh);
    }
  }
}
__device__ __forceinline__ void load_param16(const float* p, float* v) {   // p in shared memory
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = ptx::lds128f(p + 4 * i);
    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
  }
}
__device__ __forceinline__ void act16(float* v, int act) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gelu_tanh(v[i]);
  } else if (act == ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = gemm_detail::gelu_tanh_approx(v[i]);
  }
}

// EW = epilogue warps: 8 (thread = row x column half, 32-column chunks) or 16.  With 16, EPI_GENERIC
// (one staged bf16 output, no residual: the stem) runs thread = row x column quarter on ONE slab buffer
// per quarter, and EPI_PAIR / EPI_POOL2 run thread = row x 16 columns on the slab32 protocol of the
// 8-warp kernel (same slabs, barriers and store warps).  Two epilogue warps per scheduler do not hide
// their own latencies (30-45 % issue-active); EW = 16 doubles them inside the same 64 KB of staging.
// WIDE (EPI_PAIR / EPI_POOL2 on 8 warps only): the two 64-column slabs per half of rounds 1-2 instead of
// the four 32-column ones; instantiated for the cross-check tests (SVDD_SLAB32=0) only.
template <int BN, int MODE, int CG, bool HALO = false, int EW = 8, bool WIDE = false>
__global__ void __launch_bounds__(64 + 32 * EW + 32 * kStoreWarps, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
             const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmRes2,
             const GemmShape g, const EpiParams ep) {
  using C = Cfg2<BN, CG>;
  constexpr int kStages = HALO ? C::kStagesBH : C::kStages;          // (HALO: the weight ring)
  constexpr int kHalf = C::kHalf;
  constexpr int kChunks = C::kChunks;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;                                        // HALO: [kStagesAH] A tiles, then the B ring
  uint8_t* ringb_base = smem + C::kStagesAH * C::kAHBytes;           // HALO only
  uint8_t* staging = smem + (HALO ? C::kRingBytesH : kStages * C::kStageBytes);   // [half][buf] slabs, 1024-aligned
  float* s_param = reinterpret_cast<float*>(staging + kStagingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes + C::kParamBytes);
  uint64_t* full_bar = bars;                       // [kStages]   (leader's copy is the live one)
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;        // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]         (leader's copy is the live one)
  uint64_t* rin_bar = bars + 2 * kStages + 4;      // [half][buf]  (8 entries: slab32 has 4 buffers per half)
  uint64_t* rout_bar = bars + 2 * kStages + 12;    // [half][buf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 20);
  uint64_t* fulla_bar = bars + 2 * kStages + 22;   // [kStagesAH]  HALO only
  uint64_t* emptya_bar = fulla_bar + C::kStagesAH; // [kStagesAH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  auto stamp = [&](int slot) {
    if (ep.timeline != nullptr && blockIdx.x == 0) ep.timeline[slot] = (unsigned long long)clock64();
  };
  if (threadIdx.x == 0) stamp(0);

  const int tiles_l = ceil_div(g.L, g.BL);
  const int tiles_s = ceil_div(g.S, g.BS);
  // tile indices fit 32 bits (checked by the launcher): 64-bit div/mod is a ~150-cycle software
  // routine and tile_coords sits on the critical path before the first TMA of every launch
  const int m_tiles = tiles_l * tiles_s;
  const int mp_tiles = (m_tiles + CG - 1) / CG;
  const int n_tiles = (g.N + BN - 1) / BN;   // the last tile may be ragged (N % BN != 0): TMA clips / zero-fills
  const int n_end = g.n_off + g.N;
  const int total_tiles = mp_tiles * n_tiles;
  const int kblocks = g.K / kBK;
  const int kblocks2 = g.K2 / kBK;         // K-concatenated second operand (0 = none)
  const int first_tile = (int)(blockIdx.x / CG);
  const int tile_step = (int)(gridDim.x / CG);

  constexpr bool kPair = (MODE == EPI_PAIR), kPool2 = (MODE == EPI_POOL2);
  const bool out_f32 = !(kPair || kPool2) && ep.out_dtype == DT_F32;  // dtype of the staged slabs
  const bool has_out = (MODE == EPI_GENERIC) && ep.out != nullptr;
  const bool red_res = (MODE == EPI_GENERIC) && ep.res_reduce != 0;   // out += v by TMA reduce-add
  const bool has_res = (MODE == EPI_GENERIC || MODE == EPI_HEADDOT) && ep.res != nullptr && !red_res;
  const bool out2_staged = (MODE == EPI_GENERIC) && ep.out2 != nullptr && !has_res;
  // slab buffers used per 64/32-column step: EPI_PAIR = {residual in, (yd | y0) out},
  // EPI_POOL2 = {y0 in / next operand out, yd in}
  const int n_out = (kPair || kPool2) ? 2 : (has_out ? 1 : 0) + (out2_staged ? 1 : 0);
  // 32-column bf16 slabs, four buffers per half (see the header): EPI_PAIR / EPI_POOL2 on 8 or 16 warps
  // (WIDE = the 64-column slabs of rounds 1-2, cross-check only), and the 16-warp EPI_GENERIC variant
  // behind SVDD_EPI16=3 (WIDE there = one 64-column slab per column quarter, the default for the stem)
  constexpr bool slab32 = (kPair || kPool2 || (MODE == EPI_GENERIC && EW == 16)) && !WIDE;
  // 16 epilogue warps on the slab32 protocol: thread = (row, 16-column half of the current 32-column slab)
  constexpr bool kSlab16 = slab32 && EW == 16;
  const int slab_chunks = (out_f32 || slab32) ? 1 : 2;                // 32-column chunks per slab
  const int slab_cols = slab_chunks * 32;
  const int slab_bytes = slab32 ? kSlabBytes / 2 : kSlabBytes;
  const uint32_t nbuf_mask = slab32 ? 3u : 1u;                        // buffers per half - 1
  const int nbuf_shift = slab32 ? 2 : 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
    if (has_out || kPair) ptx::prefetch_tmap(&tmOut);
    if (out2_staged || kPair || (kPool2 && ep.out2 != nullptr)) ptx::prefetch_tmap(&tmOut2);
    if (has_res || kPair || kPool2) ptx::prefetch_tmap(&tmRes);
    if (kPool2) ptx::prefetch_tmap(&tmRes2);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        ptx::mbar_init(&full_bar[i], 1);
        ptx::mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull_bar[i], 1);
        ptx::mbar_init(&tempty_bar[i], CG * EW);
      }
      for (int i = 0; i < 8; ++i) {
        ptx::mbar_init(&rin_bar[i], 1);
        ptx::mbar_init(&rout_bar[i], kSlab16 ? 8 : kEpiWarps / 2);
      }
      if (HALO) {
        for (int i = 0; i < C::kStagesAH; ++i) {
          ptx::mbar_init(&fulla_bar[i], 1);
          ptx::mbar_init(&emptya_bar[i], 1);
        }
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc_cg<CG>(tmem_slot, C::kTmemCols);
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();       // everything above overlapped the previous kernel's tail
  pdl_trigger();
  if (threadIdx.x == 0) stamp(1);

  // coordinates of this CTA's half of pair-tile `t` (rr = rank within the pair)
  auto tile_coords = [&](int t, int rr) {
    TileCoord c;
    const int nt = (int)(t % n_tiles);
    const int mq = (int)(t / n_tiles);
    const int mt = (g.reverse ? mp_tiles - 1 - mq : mq) * CG + rr;
    c.n0 = g.n_off + nt * BN;    // global column
    c.in_range = mt < m_tiles;
    c.s0 = (int)(mt / tiles_l) * g.BS;       // >= S when out of range: TMA zero-fills / clips
    c.l0 = (int)(mt % tiles_l) * g.BL;
    return c;
  };
  // a tap is issued when it touches real data for either CTA of the pair
  auto tap_active = [&](int tap, int t) {
    bool any = false;
#pragma unroll
    for (int rr = 0; rr < CG; ++rr) {
      const TileCoord c = tile_coords(t, rr);
      const int l_start = c.l0 + (tap - g.taps / 2) * g.dil;
      any = any || (c.in_range && (l_start + g.BL > 0) && (l_start < g.L_in));
    }
    return any;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of a pair) =====================
    if (HALO && ptx::elect_one()) {
      // flat rows: one A box of BL + taps - 1 rows per K block (rows l0 - taps/2 ..; out-of-range
      // rows are zero-filled by TMA, pad rows between sequences are zero in memory), then the
      // weight tile of every tap
      uint32_t sa_i = 0, pa = 0, stage = 0, phase = 0;
      const uint32_t txa = (uint32_t)(CG * (g.BL + g.taps - 1) * kBK * 2);
      const uint32_t txb = (uint32_t)(CG * C::kBBytes);
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const TileCoord c = tile_coords(t, rank);
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&emptya_bar[sa_i], pa ^ 1);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&fulla_bar[sa_i], txa);
          tma_load_3d_cg<CG>(stage_base + sa_i * C::kAHBytes, &tmA, &fulla_bar[sa_i], kb * kBK, c.l0 - g.taps / 2, c.s0);
          if (++sa_i == C::kStagesAH) { sa_i = 0; pa ^= 1; }
          for (int tap = 0; tap < g.taps; ++tap) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], txb);
            tma_load_2d_cg<CG>(ringb_base + stage * C::kBBytes, &tmW, &full_bar[stage], kb * kBK,
                               tap * g.N_w + c.n0 + rank * (BN / CG));
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (!HALO && ptx::elect_one()) {   // elect.sync: TMA operands stay in uniform registers
      uint32_t stage = 0, phase = 0;
      const uint32_t tx_bytes = (uint32_t)(CG * (g.BL * g.BS * kBK * 2 + C::kBBytes));
      // Few-tile launches (the transformer GEMMs: one tile per CTA pair) stream a weight slice
      // nobody else has touched: through a 4-deep ring that is one DRAM round trip per 4 K
      // blocks.  Prefetching the slice into L2 up front turns those into L2 hits.
      if (g.prefetch_w && first_tile < total_tiles) {
        const TileCoord c = tile_coords(first_tile, rank);
        for (int tap = 0; tap < g.taps; ++tap)
          for (int kb = 0; kb < kblocks; ++kb)
            ptx::tma_prefetch_2d(&tmW, kb * kBK, tap * g.N_w + c.n0 + rank * (BN / CG));
      }
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const TileCoord c = tile_coords(t, rank);
        for (int tap = 0; tap < g.taps; ++tap) {
          if (!tap_active(tap, t)) continue;
          const int l_start = c.l0 + (tap - g.taps / 2) * g.dil;
          for (int kb = 0; kb < kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * C::kStageBytes;
            uint8_t* sb = sa + C::kABytes;
            if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            tma_load_3d_cg<CG>(sa, &tmA, &full_bar[stage], kb * kBK, l_start, c.s0);
            tma_load_2d_cg<CG>(sb, &tmW, &full_bar[stage], kb * kBK, tap * g.N_w + c.n0 + rank * (BN / CG));
            if (t == first_tile && tap == 0 && kb == 0) stamp(2);
            if (t == first_tile && kb == kblocks - 1) stamp(10);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
        // K-concatenated second operand (EPI_PAIR stage 0: the stem conv is recomputed inside the
        // residual 1x1 GEMM instead of being read back as a residual): A2 through tmRes, W2 [N, K2]
        // through tmRes2
        for (int kb = 0; kb < kblocks2; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * C::kStageBytes;
          uint8_t* sb = sa + C::kABytes;
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          tma_load_3d_cg<CG>(sa, &tmRes, &full_bar[stage], kb * kBK, c.l0, c.s0);
          tma_load_2d_cg<CG>(sb, &tmRes2, &full_bar[stage], kb * kBK, c.n0 + rank * (BN / CG));
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    // (Measured and not kept: ONE elected thread running the whole issue loop with descriptors advanced by
    // 32-bit adds, the change that took den_short_kernel's issuing warp off its critical path.  Here a stage is
    // 4 MMAs of 128 cycles, twice den_short's, and the per-stage elect region is not the limit: the k5
    // convolutions went from 482 to 502 us at stage 1, the summed GEMM table from 3.90 to 3.98 ms.)
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBM * CG, BN);
      uint32_t stage = 0, phase = 0;
      uint32_t sa_i = 0, pa = 0;           // HALO: A ring position
      uint32_t acc_stage = 0, acc_phase = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        ptx::mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc_stage * BN;
        uint32_t first = 1;
        auto consume_stage = [&]() {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (first && lane == 0) stamp(3);
          if (ptx::elect_one()) {      // elect.sync: ptxas keeps the descriptors in uniform registers (no per-UMMA ELECT/R2UR loop)
            const uint32_t sa = ptx::smem_u32(stage_base + stage * C::kStageBytes);
            const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
            const uint64_t db = ptx::make_kmajor_sw128_desc(sa + C::kABytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_bf16_cg<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, first ? (k > 0) : 1u);
            umma_commit_cg<CG>(&empty_bar[stage]);
          }
          __syncwarp();
          first = 0;
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        };
        if constexpr (HALO) {
          for (int kb = 0; kb < kblocks; ++kb) {
            ptx::mbar_wait(&fulla_bar[sa_i], pa);
            const uint32_t sa = ptx::smem_u32(stage_base + sa_i * C::kAHBytes);
            for (int tap = 0; tap < g.taps; ++tap) {
              ptx::mbar_wait(&full_bar[stage], phase);
              ptx::tc_fence_after();
              if (ptx::elect_one()) {
                // tap = row offset: the 128-byte swizzle is a function of the absolute smem address
                const uint64_t da = ptx::make_kmajor_sw128_desc(sa + (uint32_t)tap * 128u);
                const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(ringb_base + stage * C::kBBytes));
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k)
                  umma_bf16_cg<CG>(tmem_d, da + 2 * k, db + 2 * k, idesc, first ? (k > 0) : 1u);
                umma_commit_cg<CG>(&empty_bar[stage]);
                if (tap + 1 == g.taps) umma_commit_cg<CG>(&emptya_bar[sa_i]);
              }
              __syncwarp();
              first = 0;
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
            if (++sa_i == C::kStagesAH) { sa_i = 0; pa ^= 1; }
          }
        } else {
          for (int tap = 0; tap < g.taps; ++tap) {
            if (!tap_active(tap, t)) continue;
            for (int kb = 0; kb < kblocks; ++kb) consume_stage();
          }
          for (int kb = 0; kb < kblocks2; ++kb) consume_stage();
        }
        if (ptx::elect_one()) umma_commit_cg<CG>(&tfull_bar[acc_stage]);
        if (lane == 0 && t == first_tile) stamp(4);
        __syncwarp();
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else if (kSlab16 && warp < 2 + EW) {
    // ===== epilogue math, 16 warps on 32-column slabs (EPI_PAIR / EPI_POOL2 / the stem's EPI_GENERIC): thread = (row, 16 columns) =====
    // Same slabs, barriers and store warps as the 8-warp slab32 path; the eight warps of a column half
    // split every 32-column slab in two, so four warps per scheduler hide each other's latencies
    // (tcgen05.ld, lds, MUFU) where two could not (30-45 % issue-active).
    if constexpr (kSlab16) {
      const int ew = warp - 2;
      const int quad = warp & 3;           // TMEM lane quadrant this warp may read
      const int grp = ew >> 2;
      const int half = grp >> 1, sub = grp & 1;
      const int etid = threadIdx.x - 64;
      const int r = quad * 32 + lane;
      uint8_t* my_bufs32 = staging + half * 2 * kSlabBytes + (r >> 1) * 128;
      const int hi32 = ((r & 1) << 2) | (sub << 1), x7p = (r >> 1) & 3;
      uint64_t* my_rin = rin_bar + half * 4;
      uint64_t* my_rout = rout_bar + half * 4;
      const bool gelu_half = kPool2 && ep.scale2 != nullptr && ep.act2 == ACT_GELU;
      const float pool_half = gelu_half ? 0.5f : 1.0f;
      auto rd16 = [&](const uint8_t* pair_base, float* o) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint4 q = ptx::lds128(pair_base + (((hi32 | u) ^ x7p) << 4));
          const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
            o[8 * u + 2 * k] = __low2float(hh);
            o[8 * u + 2 * k + 1] = __high2float(hh);
          }
        }
      };
      auto wr16 = [&](uint8_t* pair_base, int hi, int x7, const float* o) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
          ptx::sts128(pair_base + (((hi | u) ^ x7) << 4),
              make_uint4(gemm_detail::pack_bf16x2(o[8 * u], o[8 * u + 1]), gemm_detail::pack_bf16x2(o[8 * u + 2], o[8 * u + 3]),
                         gemm_detail::pack_bf16x2(o[8 * u + 4], o[8 * u + 5]), gemm_detail::pack_bf16x2(o[8 * u + 6], o[8 * u + 7])));
      };
      uint32_t acc_stage = 0, acc_phase = 0, job = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const TileCoord tc = tile_coords(t, rank);
        const int n0 = tc.n0;
        const int s = tc.s0 + r / g.BL, l = tc.l0 + r % g.BL;
        const bool valid = (r < g.BL * g.BS) && (s < g.S) && (l < g.L);
        float* P = s_param + acc_stage * (kParamVecs * BN);
        for (int i = etid; i < BN; i += 32 * EW) {
          const bool in_n = n0 + i < n_end;
          if (ep.bias) P[P_BIAS * BN + i] = in_n ? ep.bias[n0 + i] : 0.0f;
          if (ep.scale) { P[P_SCALE * BN + i] = in_n ? ep.scale[n0 + i] : 0.0f; P[P_SHIFT * BN + i] = in_n ? ep.shift[n0 + i] : 0.0f; }
          if (ep.scale2) { P[P_SCALE2 * BN + i] = in_n ? pool_half * ep.scale2[n0 + i] : 0.0f; P[P_SHIFT2 * BN + i] = in_n ? pool_half * ep.shift2[n0 + i] : 0.0f; }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * BN + half * kHalf + sub * 16;
        uint32_t raw[2][16];
        ptx::mbar_wait(&tfull_bar[acc_stage], acc_phase);
        ptx::tc_fence_after();
        tmem_ld_32x16(taddr, raw[0]);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int c0 = half * kHalf + c * 32 + sub * 16;   // column within the tile
          constexpr int kJobs = (kPair || kPool2) ? 2 : 1;   // slabs per 32-column step
          uint8_t* buf0 = my_bufs32 + (job & 3) * (kSlabBytes / 2);
          uint8_t* buf1 = my_bufs32 + ((job + 1) & 3) * (kSlabBytes / 2);
          ptx::mbar_wait(&my_rin[job & 3], (job >> 2) & 1);
          if (kJobs == 2) ptx::mbar_wait(&my_rin[(job + 1) & 3], ((job + 1) >> 2) & 1);
          float v[16], pv[16];
          ptx::tmem_ld_wait();
          if (c + 1 < kChunks) tmem_ld_32x16(taddr + (c + 1) * 32, raw[(c + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[c & 1][i]);
          if (c + 1 == kChunks) {          // accumulator fully read: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader<CG>(&tempty_bar[acc_stage]);
          }
          if constexpr (kPool2) {
            float yd[16];
            rd16(buf0, pv);
            rd16(buf1, yd);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              float th;
              asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * v[i]));
              v[i] = fmaf(fmaf(0.5f, th, 0.5f), yd[i], pv[i]);
            }
            if (ep.scale2 != nullptr) {
              float ps[16];
              load_param16(P + P_SCALE2 * BN + c0, ps);
              load_param16(P + P_SHIFT2 * BN + c0, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] * ps[i] + pv[i];
            }
            if (gelu_half) {
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float tt;
                asm("tanh.approx.f32 %0, %1;" : "=f"(tt) : "f"((2.0f * 0.851f) * v[i]));
                v[i] = fmaf(v[i], tt, v[i]);
              }
            } else {
              act16(v, ep.act2);
            }
            wr16(buf0, hi32, x7p, v);      // in place: this thread owns these 32 bytes of the row
          } else if constexpr (MODE == EPI_GENERIC) {
            // one staged bf16 output, no residual (the stem): out = act(v * scale + shift + bias), or the
            // second stage act2(. * scale2 + shift2) of it when the launch stores `out2`
            (void)buf1;
            if (ep.scale != nullptr) {
              float ps[16];
              load_param16(P + P_SCALE * BN + c0, ps);
              load_param16(P + P_SHIFT * BN + c0, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] * ps[i] + pv[i];
            }
            if (ep.bias != nullptr) {
              load_param16(P + P_BIAS * BN + c0, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += pv[i];
            }
            act16(v, ep.act);
            if (ep.out2 != nullptr) {
              if (ep.scale2 != nullptr) {
                float ps[16];
                load_param16(P + P_SCALE2 * BN + c0, ps);
                load_param16(P + P_SHIFT2 * BN + c0, pv);
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = v[i] * ps[i] + pv[i];
              }
              act16(v, ep.act2);
            }
            wr16(buf0, hi32, x7p, v);
          } else {
            if (ep.scale != nullptr) {
              float ps[16];
              load_param16(P + P_SCALE * BN + c0, ps);
              load_param16(P + P_SHIFT * BN + c0, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] * ps[i] + pv[i];
            }
            if (ep.bias != nullptr) {
              load_param16(P + P_BIAS * BN + c0, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += pv[i];
            }
            if (!ep.act_after_res) act16(v, ep.act);
            if (ep.res != nullptr) {
              rd16(buf0, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += pv[i];
            }
            if (ep.act_after_res) act16(v, ep.act);
            const bool odd = (r & 1) != 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float other = __shfl_xor_sync(0xffffffffu, v[i], 1);
              v[i] = odd ? (valid ? v[i] - other : 0.0f) : v[i];
            }
            const int pr = r >> 1;
            wr16(buf1 + (odd ? 0 : (kBM / 2) * 64) + (pr >> 1) * 128 - pr * 128, ((pr & 1) << 2) | (sub << 1), (pr >> 1) & 3, v);
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(&my_rout[job & 3]);
            if (kJobs == 2) ptx::mbar_arrive(&my_rout[(job + 1) & 3]);
          }
          job += kJobs;
        }
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else if (EW == 16 && warp < 2 + EW) {
    // ===================== epilogue math, 16 warps: thread = (row, column quarter) =====================
    // EPI_GENERIC with one staged bf16 output and no residual (the stem): ONE slab buffer per quarter
    if constexpr (EW == 16 && !kSlab16) {
      static_assert(EW != 16 || kSlab16 || (BN == 256 && !HALO && MODE == EPI_GENERIC), "EW = 16 variants");
      const int ew = warp - 2;
      const int quad = warp & 3;           // TMEM lane quadrant this warp may read
      const int grp = ew >> 2;             // column quarter
      const int etid = threadIdx.x - 64;
      const int r = quad * 32 + lane;
      const int x7 = r & 7;
      uint8_t* buf = staging + grp * kSlabBytes + r * 128;
      uint32_t acc_stage = 0, acc_phase = 0, job = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        const TileCoord tc = tile_coords(t, rank);
        const int n0 = tc.n0;
        float* P = s_param + acc_stage * (kParamVecs * BN);
        for (int i = etid; i < BN; i += 32 * EW) {
          const bool in_n = n0 + i < n_end;
          if (ep.bias) P[P_BIAS * BN + i] = in_n ? ep.bias[n0 + i] : 0.0f;
          if (ep.scale) { P[P_SCALE * BN + i] = in_n ? ep.scale[n0 + i] : 0.0f; P[P_SHIFT * BN + i] = in_n ? ep.shift[n0 + i] : 0.0f; }
          if (ep.scale2) { P[P_SCALE2 * BN + i] = in_n ? ep.scale2[n0 + i] : 0.0f; P[P_SHIFT2 * BN + i] = in_n ? ep.shift2[n0 + i] : 0.0f; }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * BN + grp * 64;
        uint32_t raw[2][16];
        ptx::mbar_wait(&tfull_bar[acc_stage], acc_phase);
        ptx::tc_fence_after();
        tmem_ld_32x16(taddr, raw[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int c0 = grp * 64 + c * 16;  // column within the tile
          if (c == 0) ptx::mbar_wait(&rin_bar[grp], job & 1);
          float v[16], pv[16];
          ptx::tmem_ld_wait();
          if (c + 1 < 4) tmem_ld_32x16(taddr + (c + 1) * 16, raw[(c + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[c & 1][i]);
          if (c == 3) {                    // accumulator fully read: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader<CG>(&tempty_bar[acc_stage]);
          }
          if (ep.scale != nullptr) {
            float ps[16];
            load_param16(P + P_SCALE * BN + c0, ps);
            load_param16(P + P_SHIFT * BN + c0, pv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v[i] * ps[i] + pv[i];
          }
          if (ep.bias != nullptr) {
            load_param16(P + P_BIAS * BN + c0, pv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += pv[i];
          }
          act16(v, ep.act);
          if (ep.out2 != nullptr) {
            if (ep.scale2 != nullptr) {
              float ps[16];
              load_param16(P + P_SCALE2 * BN + c0, ps);
              load_param16(P + P_SHIFT2 * BN + c0, pv);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = v[i] * ps[i] + pv[i];
            }
            act16(v, ep.act2);
          }
          slab_write16(buf, x7, c, v);
          if (c == 3) {
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&rout_bar[grp]);
            ++job;
          }
        }
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else if (EW == 8 && warp < 2 + kEpiWarps) {
    // ===================== epilogue math: thread = (row, column half) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may read
    const int half = ew >> 2;
    const int etid = threadIdx.x - 64;
    const int r = quad * 32 + lane;      // tile row owned by this thread
    const int x7 = r & 7;
    uint8_t* my_bufs = staging + half * 2 * kSlabBytes + r * 128;     // + (job & 1) * kSlabBytes
    uint8_t* my_bufs32 = staging + half * 2 * kSlabBytes + (r >> 1) * 128;   // slab32: + (job & 3) * kSlabBytes / 2
    const int hi32 = (r & 1) << 2, x7p = (r >> 1) & 3;
    uint64_t* my_rin = rin_bar + half * (slab32 ? 4 : 2);
    uint64_t* my_rout = rout_bar + half * (slab32 ? 4 : 2);
    const bool gelu_half = kPool2 && ep.scale2 != nullptr && ep.act2 == ACT_GELU;
    const float pool_half = gelu_half ? 0.5f : 1.0f;
    uint32_t acc_stage = 0, acc_phase = 0;
    uint32_t job = 0;                    // staged-slab counter of this half
    for (int t = first_tile; t < total_tiles; t += tile_step) {
      const TileCoord tc = tile_coords(t, rank);
      const int n0 = tc.n0;
      const int s = tc.s0 + r / g.BL, l = tc.l0 + r % g.BL;
      const bool valid = (r < g.BL * g.BS) && (s < g.S) && (l < g.L);
      const int64_t row = (int64_t)s * g.L + l;

      float* P = s_param + acc_stage * (kParamVecs * BN);
      for (int i = etid; i < BN; i += kEpiThreads) {
        const bool in_n = n0 + i < n_end;
        if (ep.bias) P[P_BIAS * BN + i] = in_n ? ep.bias[n0 + i] : 0.0f;
        if (ep.scale) { P[P_SCALE * BN + i] = in_n ? ep.scale[n0 + i] : 0.0f; P[P_SHIFT * BN + i] = in_n ? ep.shift[n0 + i] : 0.0f; }
        // EPI_POOL2 feeding a GELU: the BN affine is staged HALVED, h = x / 2 exactly (power of two), and
        // gelu(x) = h + h * tanh(1.702 h) -- one multiply less per element, bit-identical to gelu_tanh(x)
        if (ep.scale2) { P[P_SCALE2 * BN + i] = in_n ? pool_half * ep.scale2[n0 + i] : 0.0f; P[P_SHIFT2 * BN + i] = in_n ? pool_half * ep.shift2[n0 + i] : 0.0f; }
        if (MODE == EPI_HEADDOT) P[P_HEADW * BN + i] = in_n ? ep.head_w[n0 + i] : 0.0f;
      }
      gemm_detail::epi_bar_sync();

      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * BN + half * kHalf;
      const int cbase = half * kHalf;
      float head_acc = 0.0f;
      uint32_t raw[2][32];
      ptx::mbar_wait(&tfull_bar[acc_stage], acc_phase);
      ptx::tc_fence_after();
      if (t == first_tile && etid == 0) stamp(5);
      ptx::tmem_ld_32x32(taddr, raw[0]);
      uint8_t* buf0 = nullptr;           // slab of the `out` job (holds the residual on entry)
      uint8_t* buf1 = nullptr;           // slab of the staged `out2` job
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int c0 = cbase + c * 32;   // column within the tile
        const int cis = c % slab_chunks; // chunk within its slab
        if (n_out > 0 && cis == 0) {
          if (slab32) {
            buf0 = my_bufs32 + (job & 3) * (kSlabBytes / 2);
            ptx::mbar_wait(&my_rin[job & 3], (job >> 2) & 1);
            buf1 = my_bufs32 + ((job + 1) & 3) * (kSlabBytes / 2);
            ptx::mbar_wait(&my_rin[(job + 1) & 3], ((job + 1) >> 2) & 1);
          } else {
          buf0 = my_bufs + (job & 1) * kSlabBytes;
          ptx::mbar_wait(&my_rin[job & 1], (job >> 1) & 1);
          if (n_out == 2) {
            buf1 = my_bufs + ((job + 1) & 1) * kSlabBytes;
            ptx::mbar_wait(&my_rin[(job + 1) & 1], ((job + 1) >> 1) & 1);
          }
          }
        }
        float v[32];
        ptx::tmem_ld_wait();
        if (c + 1 < kChunks) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, raw[(c + 1) & 1]);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[c & 1][i]);
        if (c + 1 == kChunks) {          // accumulator fully read: hand it back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader<CG>(&tempty_bar[acc_stage]);
        }
        float pv[32];
        if constexpr (kPool2) {
          // v = Wp.(y1 - y0): pooled = y0 + sigmoid(v) * (y1 - y0)
          float yd[32];
          if (slab32) {
            slab32_read(buf0, hi32, x7p, pv);
            slab32_read(buf1, hi32, x7p, yd);
          } else {
            slab_read(buf0, x7, false, cis, pv);
            slab_read(buf1, x7, false, cis, yd);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // sigmoid(v) = 0.5 + 0.5 tanh(v / 2): one MUFU op per element instead of ex2 + rcp (the
            // pool epilogue is MUFU-bound: sigmoid + the next layer's GELU on 128 columns per thread)
            float th;
            asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * v[i]));
            const float w1 = fmaf(0.5f, th, 0.5f);
            v[i] = fmaf(w1, yd[i], pv[i]);
          }
          if (ep.out != nullptr) {         // last stage: fp32 transformer stream, direct
            if (valid && n0 + c0 < n_end) gemm_detail::store_row32(ep.out, DT_F32, row * ep.ld_out + n0 + c0, v);
          }
          if (ep.out2 != nullptr) {
            if (ep.scale2 != nullptr) {
              float ps[32];
              gemm_detail::load_param32(P + P_SCALE2 * BN + c0, ps);
              gemm_detail::load_param32(P + P_SHIFT2 * BN + c0, pv);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = v[i] * ps[i] + pv[i];
            }
            if (gelu_half) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float t;
                asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"((2.0f * 0.851f) * v[i]));
                v[i] = fmaf(v[i], t, v[i]);
              }
            } else {
              act32(v, ep.act2);
            }
            if (slab32) slab32_write(buf0, hi32, x7p, v);
            else slab_write(buf0, x7, false, cis, v);     // in place: this thread owns the row
          }
        } else {
        if (ep.scale != nullptr) {
          float ps[32];
          gemm_detail::load_param32(P + P_SCALE * BN + c0, ps);
          gemm_detail::load_param32(P + P_SHIFT * BN + c0, pv);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] * ps[i] + pv[i];
        }
        if (ep.bias != nullptr) {
          gemm_detail::load_param32(P + P_BIAS * BN + c0, pv);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += pv[i];
        }
        if (!ep.act_after_res) act32(v, ep.act);
        if (has_res || (kPair && ep.res != nullptr)) {
          if (kPair && slab32) {
            slab32_read(buf0, hi32, x7p, pv);
          } else if (has_out || kPair) {
            slab_read(buf0, x7, out_f32, cis, pv);
          } else if (valid && n0 + c0 < n_end) {   // HEADDOT with a residual: direct (unused by the nets)
            gemm_detail::load_row32(ep.res, ep.res_dtype, row * ep.ld_res + n0 + c0, pv);
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += pv[i];
        }
        if (ep.act_after_res) act32(v, ep.act);
        if constexpr (MODE == EPI_HEADDOT) {
          gemm_detail::load_param32(P + P_HEADW * BN + c0, pv);
#pragma unroll
          for (int i = 0; i < 32; ++i) head_acc += v[i] * pv[i];
        }
        if constexpr (kPair) {
          // rows (2j, 2j+1) sit on adjacent lanes: even lane keeps y0 = y[2j], odd lane forms
          // yd = y[2j+1] - y[2j] (0 when 2j+1 is past the sequence end); slab row j of the
          // lower half (yd) / upper half (y0) of buf1
          const bool odd = (r & 1) != 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float other = __shfl_xor_sync(0xffffffffu, v[i], 1);
            v[i] = odd ? (valid ? v[i] - other : 0.0f) : v[i];
          }
          const int pr = r >> 1;
          if (slab32)       // rows 0..63 (yd) / 64..127 (y0) of a 64-byte-row slab; buf1 = slab + (r >> 1) * 128
            slab32_write(buf1 + (odd ? 0 : (kBM / 2) * 64) + (pr >> 1) * 128 - pr * 128, (pr & 1) << 2, (pr >> 1) & 3, v);
          else
          slab_write(buf1 + (odd ? 0 : (kBM / 2) * 128) + pr * 128 - r * 128, pr & 7, false, cis, v);
        } else {
        if (has_out) slab_write(buf0, x7, out_f32, cis, v);
        if (MODE == EPI_GENERIC && ep.out2 != nullptr) {
          if (ep.scale2 != nullptr) {
            float ps[32];
            gemm_detail::load_param32(P + P_SCALE2 * BN + c0, ps);
            gemm_detail::load_param32(P + P_SHIFT2 * BN + c0, pv);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v[i] * ps[i] + pv[i];
          }
          act32(v, ep.act2);
          if (out2_staged) {
            slab_write(has_out ? buf1 : buf0, x7, out_f32, cis, v);
          } else if (valid && n0 + c0 < n_end) {   // residual + second output (one launch per pass): direct
            gemm_detail::store_row32(ep.out2, ep.out2_dtype, row * ep.ld_out2 + n0 + c0, v);
          }
        }
        }
        }
        if (n_out > 0 && cis == slab_chunks - 1) {
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::mbar_arrive(&my_rout[job & nbuf_mask]);
            if (n_out == 2) ptx::mbar_arrive(&my_rout[(job + 1) & nbuf_mask]);
          }
          job += n_out;
        }
      }
      if (MODE == EPI_HEADDOT && valid)
        ep.partials[row * (2 * n_tiles) + 2 * ((n0 - g.n_off) / BN) + half] = head_acc;
      if (t == first_tile && etid == 0) stamp(6);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== slab store / prefetch warp (one per column half) =====================
    const int half = warp - (2 + EW);
    if constexpr (EW == 16 && !kSlab16) {
      // one buffer per column quarter, quarters 2*half and 2*half+1 served in turn
      if (ptx::elect_one()) {
        auto provision = [&](int t, int grp) {          // output-only slabs: nothing to load
          if (t >= total_tiles) return;
          ptx::mbar_arrive(&rin_bar[grp]);
        };
        provision(first_tile, 2 * half);
        provision(first_tile, 2 * half + 1);
        uint32_t j = 0;
        for (int t = first_tile; t < total_tiles; t += tile_step, ++j) {
          const TileCoord c = tile_coords(t, rank);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int grp = 2 * half + q;
            ptx::mbar_wait(&rout_bar[grp], j & 1);
            const uint8_t* buf = staging + grp * kSlabBytes;
            const int col = c.n0 + grp * 64;
            tma_store_3d(has_out ? &tmOut : &tmOut2, buf, col, c.l0, c.s0);
            bulk_commit();
          }
          bulk_wait_read0();
          provision(t + tile_step, 2 * half);
          provision(t + tile_step, 2 * half + 1);
        }
        bulk_wait_all();
      }
    } else
    if (n_out > 0 && ptx::elect_one()) {
      uint8_t* bufs = staging + half * 2 * kSlabBytes;
      uint64_t* my_rin = rin_bar + half * (slab32 ? 4 : 2);
      uint64_t* my_rout = rout_bar + half * (slab32 ? 4 : 2);
      const int slabs = kHalf / slab_cols;
      const uint32_t res_bytes = (uint32_t)(g.BL * g.BS * (slab32 ? 64 : 128));
      // job iterator: (tile, slab, buffer-of-the-step)
      struct It { int t; int slab, o; };
      auto advance = [&](It& it) {
        if (++it.o == n_out) { it.o = 0; if (++it.slab == slabs) { it.slab = 0; it.t += tile_step; } }
      };
      // what is loaded into the buffer before the math warps may touch it
      auto load_map = [&](int o) -> const CUtensorMap* {
        if (kPair) return (o == 0 && ep.res != nullptr) ? &tmRes : nullptr;
        if (kPool2) return o == 0 ? &tmRes : &tmRes2;
        return (has_res && o == 0) ? &tmRes : nullptr;
      };
      auto provision = [&](const It& it, uint32_t j) {     // make buffer j & nbuf_mask ready for job j
        if (it.t >= total_tiles) return;
        const CUtensorMap* m = load_map(it.o);
        const uint32_t b = j & nbuf_mask;
        if (m != nullptr) {
          const TileCoord c = tile_coords(it.t, rank);
          ptx::mbar_arrive_expect_tx(&my_rin[b], res_bytes);
          ptx::tma_load_3d(bufs + b * slab_bytes, m, &my_rin[b],
                           c.n0 + half * kHalf + it.slab * slab_cols, c.l0, c.s0);
        } else {
          ptx::mbar_arrive(&my_rin[b]);
        }
      };
      It cur{first_tile, 0, 0}, ahead{first_tile, 0, 0};
      for (uint32_t j = 0; j <= nbuf_mask; ++j) {
        provision(ahead, j);
        advance(ahead);
      }
      for (uint32_t j = 0; cur.t < total_tiles; ++j) {
        const TileCoord c = tile_coords(cur.t, rank);
        ptx::mbar_wait(&my_rout[j & nbuf_mask], (j >> nbuf_shift) & 1);
        const uint8_t* buf = bufs + (j & nbuf_mask) * slab_bytes;
        const int col = c.n0 + half * kHalf + cur.slab * slab_cols;
        bool stored = false;
        if (kPair) {
          if (cur.o == 1) {               // rows 0..63: yd, rows 64..127: y0, both at half length
            tma_store_3d(&tmOut2, buf, col, c.l0 >> 1, c.s0);
            tma_store_3d(&tmOut, buf + (kBM / 2) * (slab32 ? 64 : 128), col, c.l0 >> 1, c.s0);
            stored = true;
          }
        } else if (kPool2) {
          if (cur.o == 0 && ep.out2 != nullptr) { tma_store_3d(&tmOut2, buf, col, c.l0, c.s0); stored = true; }
        } else {
          const bool second = (cur.o == 1) || !has_out;      // this job carries `out2`
          if (red_res && !second) tma_reduce_add_3d(&tmOut, buf, col, c.l0, c.s0);
          else tma_store_3d(second ? &tmOut2 : &tmOut, buf, col, c.l0, c.s0);
          stored = true;
        }
        if (stored) {
          bulk_commit();
          bulk_wait_read0();
        }
        provision(ahead, j + nbuf_mask + 1);
        advance(ahead);
        advance(cur);
      }
      bulk_wait_all();
      if (half == 0) stamp(7);
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  if (threadIdx.x == 0) stamp(8);
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, C::kTmemCols);
  }
  if (threadIdx.x == 0) stamp(9);
}

}  // namespace gemm2

// The host launcher lives in conv_gemm.cu (launch_conv_gemm routes here when the shape and
// epilogue are ones this kernel handles).
}  // namespace svdd
