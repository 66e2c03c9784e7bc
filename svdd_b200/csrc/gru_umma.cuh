// The biGRU recurrence of the RNA value net on tcgen05 (reference: GRUBlock.forward
// Enformer.py:1607-1630 = torch.nn.GRU, gate order r, z, n):
//   r = s(W_ir x + b_ir + W_hr h + b_hr)      z = s(W_iz x + b_iz + W_hz h + b_hz)
//   n = tanh(W_in x + b_in + r * (W_hn h + b_hn))      h' = (1 - z) n + z h
//
// One CTA = 256 sequences x ONE direction, as two groups of 128 sequences that take turns: while
// the eight epilogue warps do the gate arithmetic of one group's step, the tensor core runs the
// other group's.  Per (group, step) one accumulator of 256 TMEM columns
//   [ r : 64 | z : 64 | gi_n : 64 | gh_n : 64 ]
// is produced by
//   x_t[128 x 64] . W_ih^T            -> columns 0..191          (bf16 x bf16)
//   h  [128 x 64] . W_hh[r,z]^T       -> += columns 0..127       (hi.hi + lo.hi + hi.lo)
//   h  [128 x 64] . W_hh[n]^T         ->    columns 192..255     (hi.hi + lo.hi + hi.lo)
// The INPUT projection is thus fused into the recurrence: the fp32 gi tensor of the mma.sync path
// (rows x L x 384 floats, written by a GEMM launch and read back here: 3.9 GB per 51 200 x 50 pass)
// no longer exists; x_t arrives by TMA as one 128-row box per step (rows = sequences, stride L).
// The recurrent product keeps fp32 accuracy through the bf16 hi / lo split of h and W_hh (the
// dropped lo.lo term is 2^-18 relative), as the mma.sync kernel did; h itself stays fp32 in the
// registers of the thread that owns (sequence, 32 hidden units).
// Roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (TMEM lane
// quadrant = warp & 3, unit half = (warp - 2) >> 2).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace svdd {
namespace gruu {

constexpr int kC = 64;
constexpr int kG3 = 3 * kC;
constexpr int kRowsG = 128;                     // sequences per group
constexpr int kGroups = 2;
constexpr int kXStages = 2;
constexpr int kTileBytes = kRowsG * 128;        // 128 rows x 64 bf16
constexpr int kWBytes = kG3 * 128;              // [192 gate rows][64] bf16
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kBiasFloats = kG3 + kC;           // folded gate biases (r, z, n_in) + b_hn
constexpr int kSmemBytes = 3 * kWBytes + kGroups * kXStages * kTileBytes + kGroups * 2 * kTileBytes +
                           kBiasFloats * 4 + 256 + 1024;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// Gate non-linearities on the fast paths of the SFU: ex2.approx (2 ulp) + approximate division.
// sigmoid: |error| <= ~2e-7; tanh(x) = 1 - 2 / (1 + e^{2x}): absolute error <= ~2.5e-7 (the
// cancellation near 0 costs relative, not absolute accuracy, and h' = (1 - z) n + z h only needs
// the latter).  The libm tanhf + IEEE divisions these replace were ~70 of the ~100 instructions per
// hidden unit and made the epilogue, not the tensor core, the pace of the recurrence (6-7 us per
// group and step against 0.9 us of MMAs).
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_tanh(float x) { return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x)); }

__device__ __forceinline__ void ld_bias16(const float* p, float* v) {    // p in shared memory
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 f = ptx::lds128f(p + 4 * i);
    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
  }
}

__global__ void __launch_bounds__(kThreads, 1)
cg_gru_umma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWx,
                   const __grid_constant__ CUtensorMap tmWhi, const __grid_constant__ CUtensorMap tmWlo,
                   const float* __restrict__ gate_b /*[2][192]*/, const float* __restrict__ bhn /*[2][64]*/,
                   float* __restrict__ y /*[2][rows*L][64]*/, int64_t rows, int L) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_wx = smem;
  uint8_t* s_whi = s_wx + kWBytes;
  uint8_t* s_wlo = s_whi + kWBytes;
  uint8_t* s_x = s_wlo + kWBytes;                                   // [group][stage]
  uint8_t* s_h = s_x + kGroups * kXStages * kTileBytes;             // [group][hi, lo]
  float* s_bias = reinterpret_cast<float*>(s_h + kGroups * 2 * kTileBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_bias + kBiasFloats);
  uint64_t* w_full = bars;
  uint64_t* x_full = bars + 1;            // [group][stage]
  uint64_t* x_empty = bars + 5;           // [group][stage]
  uint64_t* acc_full = bars + 9;          // [group]
  uint64_t* h_ready = bars + 11;          // [group]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int dir = blockIdx.y;
  const int64_t seq0 = (int64_t)blockIdx.x * (kGroups * kRowsG);
  const bool glive[2] = {seq0 < rows, seq0 + kRowsG < rows};

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX);
    ptx::prefetch_tmap(&tmWx);
    ptx::prefetch_tmap(&tmWhi);
    ptx::prefetch_tmap(&tmWlo);
  }
  if (warp == 1) {
    if (lane == 0) {
      ptx::mbar_init(w_full, 1);
      for (int i = 0; i < kGroups * kXStages; ++i) {
        ptx::mbar_init(&x_full[i], 1);
        ptx::mbar_init(&x_empty[i], 1);
      }
      for (int g = 0; g < kGroups; ++g) {
        ptx::mbar_init(&acc_full[g], 1);
        ptx::mbar_init(&h_ready[g], kEpiWarps);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // h_0 = 0
  for (int i = threadIdx.x; i < kGroups * 2 * kTileBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(s_h)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < kBiasFloats; i += kThreads)
    s_bias[i] = i < kG3 ? gate_b[dir * kG3 + i] : bhn[dir * kC + i - kG3];
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(w_full, 3 * kWBytes);
      ptx::tma_load_2d(s_wx, &tmWx, w_full, 0, dir * kG3);
      ptx::tma_load_2d(s_whi, &tmWhi, w_full, 0, dir * kG3);
      ptx::tma_load_2d(s_wlo, &tmWlo, w_full, 0, dir * kG3);
      for (int s = 0; s < L; ++s) {
        const int t = dir == 0 ? s : L - 1 - s;
        const int stage = s & 1;
        for (int g = 0; g < kGroups; ++g) {
          if (!glive[g]) continue;
          ptx::mbar_wait(&x_empty[g * kXStages + stage], ((s >> 1) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&x_full[g * kXStages + stage], kTileBytes);
          ptx::tma_load_3d(s_x + (g * kXStages + stage) * kTileBytes, &tmX, &x_full[g * kXStages + stage], 0, t,
                           (int)(seq0 + g * kRowsG));
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t id_x = ptx::make_idesc_bf16(128, kG3);
    constexpr uint32_t id_rz = ptx::make_idesc_bf16(128, 2 * kC);
    constexpr uint32_t id_n = ptx::make_idesc_bf16(128, kC);
    ptx::mbar_wait(w_full, 0);
    ptx::tc_fence_after();
    const uint64_t d_wx = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_wx));
    const uint64_t d_whi = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_whi));
    const uint64_t d_wlo = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_wlo));
    const uint64_t d_whi_n = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_whi + 2 * kC * 128));
    const uint64_t d_wlo_n = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_wlo + 2 * kC * 128));
    for (int s = 0; s < L; ++s) {
      const int stage = s & 1;
      for (int g = 0; g < kGroups; ++g) {
        if (!glive[g]) continue;
        ptx::mbar_wait(&h_ready[g], s & 1);                           // h_s written, accumulator drained
        ptx::mbar_wait(&x_full[g * kXStages + stage], (s >> 1) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t D = tmem_base + g * 256;
          const uint64_t d_x = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_x + (g * kXStages + stage) * kTileBytes));
          const uint64_t d_hhi = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_h + (g * 2 + 0) * kTileBytes));
          const uint64_t d_hlo = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_h + (g * 2 + 1) * kTileBytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(D, d_x + 2 * k, d_wx + 2 * k, id_x, k != 0);
          // recurrent part: lo.hi, hi.lo first (small terms), hi.hi last
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(D, d_hlo + 2 * k, d_whi + 2 * k, id_rz, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(D, d_hhi + 2 * k, d_wlo + 2 * k, id_rz, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(D, d_hhi + 2 * k, d_whi + 2 * k, id_rz, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(D + 3 * kC, d_hlo + 2 * k, d_whi_n + 2 * k, id_n, k != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(D + 3 * kC, d_hhi + 2 * k, d_wlo_n + 2 * k, id_n, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(D + 3 * kC, d_hhi + 2 * k, d_whi_n + 2 * k, id_n, 1u);
          ptx::umma_commit(&x_empty[g * kXStages + stage]);
          ptx::umma_commit(&acc_full[g]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue: thread = (sequence, 32 hidden units) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    const int row = quad * 32 + lane;
    const int x7 = row & 7;
    float hst[kGroups][32];
#pragma unroll
    for (int g = 0; g < kGroups; ++g)
#pragma unroll
      for (int i = 0; i < 32; ++i) hst[g][i] = 0.0f;
    if (lane == 0) {
      ptx::mbar_arrive(&h_ready[0]);       // h_0 (zeros) is in place
      ptx::mbar_arrive(&h_ready[1]);
    }
    float* ydir = y + (size_t)dir * rows * L * kC;
    for (int s = 0; s < L; ++s) {
      const int t = dir == 0 ? s : L - 1 - s;
#pragma unroll
      for (int g = 0; g < kGroups; ++g) {
        if (!glive[g]) continue;
        const int64_t seq = seq0 + g * kRowsG + row;
        const bool live = seq < rows;
        ptx::mbar_wait(&acc_full[g], s & 1);
        ptx::tc_fence_after();
        const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + g * 256;
        uint8_t* hrow_hi = s_h + (g * 2 + 0) * kTileBytes + row * 128;
        uint8_t* hrow_lo = s_h + (g * 2 + 1) * kTileBytes + row * 128;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int u0 = half * 32 + sub * 16;
          uint32_t ra[16], rb[16];
          float r[16], z[16], bb[16];
          tmem_ld16(tb + u0, ra);
          tmem_ld16(tb + kC + u0, rb);
          ptx::tmem_ld_wait();
          ld_bias16(s_bias + u0, bb);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = fast_sigmoid(__uint_as_float(ra[i]) + bb[i]);
          ld_bias16(s_bias + kC + u0, bb);
#pragma unroll
          for (int i = 0; i < 16; ++i) z[i] = fast_sigmoid(__uint_as_float(rb[i]) + bb[i]);
          tmem_ld16(tb + 2 * kC + u0, ra);       // gi_n
          tmem_ld16(tb + 3 * kC + u0, rb);       // gh_n
          ptx::tmem_ld_wait();
          float bn[16];
          ld_bias16(s_bias + 2 * kC + u0, bb);   // b_in
          ld_bias16(s_bias + kG3 + u0, bn);      // b_hn
          float hn[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float n = fast_tanh((__uint_as_float(ra[i]) + bb[i]) + r[i] * (__uint_as_float(rb[i]) + bn[i]));
            hn[i] = (1.0f - z[i]) * n + z[i] * hst[g][sub * 16 + i];
            hst[g][sub * 16 + i] = hn[i];
          }
          // h' -> bf16 hi / lo planes (the next step's A operand), 16-byte chunk j of the 128-byte row
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a = hn[8 * c + 2 * k], b = hn[8 * c + 2 * k + 1];
              const __nv_bfloat162 hh = __floats2bfloat162_rn(a, b);
              const __nv_bfloat162 ll = __floats2bfloat162_rn(a - __low2float(hh), b - __high2float(hh));
              hi[k] = *reinterpret_cast<const uint32_t*>(&hh);
              lo[k] = *reinterpret_cast<const uint32_t*>(&ll);
            }
            const int j = (u0 >> 3) + c;
            ptx::sts128(hrow_hi + ((j ^ x7) << 4), make_uint4(hi[0], hi[1], hi[2], hi[3]));
            ptx::sts128(hrow_lo + ((j ^ x7) << 4), make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
          if (live) {
            float4* yo = reinterpret_cast<float4*>(ydir + ((size_t)seq * L + t) * kC + u0);
#pragma unroll
            for (int i = 0; i < 4; ++i) yo[i] = make_float4(hn[4 * i], hn[4 * i + 1], hn[4 * i + 2], hn[4 * i + 3]);
          }
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&h_ready[g]);
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace gruu
}  // namespace svdd
