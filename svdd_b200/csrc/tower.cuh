// Persistent transformer-tower kernel: ALL blocks of the Enformer-style value net's transformer
// tower (Enformer.py:1931-1949: LN -> MHA -> +res ; LN -> Linear -> ReLU -> Linear -> +res) in ONE
// launch.  The tower sees R = candidates x n rows (n = 2 positions at L = 200), so each of its 44
// GEMMs is a single wave of 60-120 CTA-pair tiles: launched one by one, half of every launch is
// prologue, exposed epilogue and drain (profiles/r01_summary.md, "anatomy of one transformer GEMM").
//
// Here the tower is a list of work items in topological order
//
//   per block:  LN1 | QKV GEMM | attention | out-proj GEMM (+= x) | LN2 | FF1 GEMM (ReLU) | FF2 GEMM (+= x)
//   after the last block:  BN + GELU of x (the pointwise ConvBlock's operand)
//
// and the only dependency is per ROW TILE (256 rows = one CTA pair's M): phase p of row tile r
// needs phase p-1 of row tile r, nothing else (attention mixes the n positions of one sequence,
// which are adjacent rows).  Every phase has a counter per row tile in global memory; an item
// waits until the previous phase's counter reaches its item count and bumps its own when its
// stores are globally visible.  Items are dealt round-robin to the resident CTA pairs, each pair
// walks its items in increasing order, so the item with the lowest unfinished id can always run:
// no deadlock as long as all pairs are co-resident (grid <= SM count, 1 CTA / SM).
//
// GEMM items reuse the cta_group::2 machinery of conv_gemm2.cuh (256 x 256 tile per pair, TMA
// producer, one MMA issuer, 8 epilogue warps, 2 slab-store warps, double-buffered TMEM), with
// two differences: the weight tiles of the first ring stages are requested BEFORE the dependency
// wait (weights depend on nothing), and the accumulators of tile i+1 are computed while tile i's
// epilogue drains, across GEMM boundaries.  LN / attention / BN items run on the 8 epilogue warps
// (warp per row / per (sequence, head), the arithmetic of ef_ln_warp_kernel /
// ef_attention_warp_kernel), reading through L2 (ld.global.cg).
#pragma once
#include "conv_gemm2.cuh"

namespace svdd {
namespace tower {

using gemm_detail::kBK;
using gemm_detail::kBM;
using gemm_detail::kEpiThreads;
using gemm_detail::kEpiWarps;
using gemm2::kSlabBytes;
using gemm2::kStagingBytes;
using gemm2::kThreads;

constexpr int kTileRows = 2 * kBM;        // rows per CTA pair
constexpr int kSubRows = 16;              // rows per CTA of one LN / attention / BN item
constexpr int kSubItems = kBM / kSubRows; // such items per row tile
constexpr int kPhasesPerBlock = 7;
constexpr int kMaxSplitTiles = 16;        // column tiles of a split-K phase (C <= 4096)
constexpr int kLnVecs = 16;               // C <= 2048 (float4s per lane)

enum PhaseType { PH_LN = 0, PH_GEMM = 1, PH_ATTN = 2, PH_BNACT = 3 };
enum OutKind { OUT_F32_STORE = 0, OUT_F32_REDUCE = 1, OUT_BF16_RELU = 2 };
enum AMap { A_HN = 0, A_AO = 1, A_U = 2 };

struct Phase {
  int type = PH_LN;
  int items_per_rt = 0;   // column tiles (GEMM) or sub-items per row tile
  int signals = 0;        // counter increments per item (both CTAs of the pair together)
  int kblocks = 0;        // GEMM: K / 64
  int n_cols = 0;         // GEMM: N (the last column tile may be ragged: TMA zero-fills / clips)
  int a_map = 0;          // GEMM: AMap
  int out_kind = 0;       // GEMM: OutKind
  int w_idx = 0;          // GEMM: weight map of the block; LN: 0 = ln1, 1 = ln2
  int start = 0;          // first item of this phase within its block (already x RT)
  // GEMM with a reduce-add output: K split over `ksplit` items per column tile (item c = tile *
  // ksplit + ks).  The partial sums are added into `x` IN ORDER (slice ks waits for the stores of
  // slice ks-1 of the same tile), so the result does not depend on timing.
  int ksplit = 1;
};

// per transformer block, in global memory (tensor maps must be 64-byte aligned)
struct alignas(128) BlockParams {
  CUtensorMap w[4];       // wqkv [nqkv, C], wo [C, H*dv], wf1 [2C, C], wf2 [C, 2C]; box 64 x 128, 128B swizzle
  const float* ln_g[2];
  const float* ln_b[2];
  const float* bias[4];   // per GEMM (nullptr for QKV)
  const float* rcb;
  const float* rpb;
  const float* relk;      // [H][2n-1][dk]
};

struct TowerArgs {
  int R = 0, RT = 0, n_blocks = 0, n_pos = 0;
  int C = 0, nqkv = 0, H = 0, dk = 0, dv = 0;
  int items_per_block = 0, total_items = 0;
  Phase ph[kPhasesPerBlock + 1];
  const BlockParams* blocks = nullptr;
  float* xt = nullptr;
  __nv_bfloat16* hn = nullptr;
  const float* qkv = nullptr;
  __nv_bfloat16* ao = nullptr;
  const float* bn_s = nullptr;   // final BN (folded) + GELU -> hn
  const float* bn_t = nullptr;
  unsigned* flags = nullptr;     // [(n_blocks * 7) * RT], zeroed before the launch
  unsigned* sk_flags = nullptr;  // split-K order: [n_blocks][RT][kMaxSplitTiles] completed-store counters, zeroed too
  int attn_fast = 1;             // per-sequence attention items (attn_warp_seq); 0 = per-(sequence, head) tasks
  // debugging aid (SVDD_TOWER_TRACE=file): per item 8 x globaltimer ns, written by the pair's even CTA:
  // [0] wait for the dependency begins (producer / row item) [1] dependency satisfied [2] MMA may start
  // (accumulator free) [3] accumulator complete (epilogue) [4] epilogue / row math done [5] published
  // [6] pair index
  unsigned long long* trace = nullptr;
  // SVDD_TOWER_DISCARD (default on): row items drop intermediates whose last reader has finished from L2
  // (discard.global.L2) instead of letting the dead lines be written back: an attention item its
  // sequences' qkv rows and LN1 rows, an LN2 item its rows of the attention output (the out-projection
  // of the row tile is complete), an LN1 item its rows of the previous block's FFN hidden tensor.
  int discard_qkv = 0;
  __nv_bfloat16* u = nullptr;   // FFN hidden [R, 2C] (the GEMMs reach it through tm_u)
};

template <int kBN>                         // tile columns per CTA pair: 256 or 128
struct Cfg {
  static_assert(kBN == 256 || kBN == 128, "tower tiles are 256 or 128 columns wide");
  static constexpr int kABytes = kBM * kBK * 2;            // 16 KB: this CTA's 128 rows
  static constexpr int kBBytes = (kBN / 2) * kBK * 2;      // this CTA's half of the weight tile
  static constexpr int kStageBytes = kABytes + kBBytes;
#ifndef SVDD_TOWER_STAGES
#define SVDD_TOWER_STAGES 4
#endif
  static constexpr int kStages = kBN == 256 ? SVDD_TOWER_STAGES : 5;   // operand ring depth (3: +9 % tower time, measured)
  static constexpr int kBiasBytes = 2 * kBN * 4;
  static constexpr int kAttnWBytes = 16 * 128 * 4;           // softmax weights of one sequence per epilogue warp (H*N*N <= 128)
  static constexpr int kBarBytes = 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kBiasBytes + kAttnWBytes + kBarBytes + 1024;
  static constexpr int kTmemCols = 2 * kBN;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

// ---- cross-CTA flags --------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Spins until *p >= target.  A scheduling bug would hang the GPU, so it traps after ~seconds.
__device__ __forceinline__ void wait_flag(const unsigned* p, unsigned target) {
  if (p == nullptr) return;
  uint32_t spins = 0;
  while (ld_acquire_gpu(p) < target) {
    __nanosleep(64);
    if (++spins == (1u << 23)) {
      printf("svdd_b200: tower flag wait timed out (block %d thread %d target %u have %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, target, ld_acquire_gpu(p));
      __trap();
    }
  }
}

// mbarrier wait of the tower's roles: most of a role's life here is waiting (the kernel is bound
// by the dependency chain, not by issue slots), so the poll carries a suspend-time hint: the warp
// sleeps in hardware until the phase completes or the hint expires instead of spinning
// (171 M warp instructions per launch, mostly polls, before this).
#ifndef SVDD_TOWER_WAIT_HINT_NS
#define SVDD_TOWER_WAIT_HINT_NS 1000
#endif
__device__ __forceinline__ void twait(uint64_t* bar, uint32_t parity) {
#if SVDD_TOWER_WAIT_HINT_NS > 0
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(ptx::smem_u32(bar)), "r"(parity), "r"((uint32_t)SVDD_TOWER_WAIT_HINT_NS)
        : "memory");
    if (ok) return;
    if (++spins == (1u << 22)) {
      printf("svdd_b200: tower mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
#else
  ptx::mbar_wait(bar, parity);
#endif
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}

// ---- row-wise pieces (shared with the stand-alone kernels of enformer.cu) ----------------------
template <bool CG> __device__ __forceinline__ float4 ld_f4(const float4* p) { return CG ? __ldcg(p) : *p; }
template <bool CG> __device__ __forceinline__ float ld_f1(const float* p) { return CG ? __ldcg(p) : *p; }

// LayerNorm of one row by one warp: fp32 [C] -> bf16 [C]; C % 128 == 0, C <= 128 * kLnVecs.
// Two-pass variance, the row lives in registers as float4s.
template <bool CG>
__device__ __forceinline__ void ln_warp_row(const float* __restrict__ x_row, const float* __restrict__ g,
                                            const float* __restrict__ b, __nv_bfloat16* __restrict__ out_row,
                                            int C, int lane) {
  const int nvec = C >> 7;
  const float4* xr = reinterpret_cast<const float4*>(x_row);
  float4 v[kLnVecs];
  float sum = 0.0f;
#pragma unroll
  for (int i = 0; i < kLnVecs; ++i) {
    if (i < nvec) {
      v[i] = ld_f4<CG>(xr + i * 32 + lane);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float sq = 0.0f;
#pragma unroll
  for (int i = 0; i < kLnVecs; ++i) {
    if (i < nvec) {
      const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
      sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / (float)C + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  uint2* o2 = reinterpret_cast<uint2*>(out_row);
#pragma unroll
  for (int i = 0; i < kLnVecs; ++i) {
    if (i < nvec) {
      const float4 gg = __ldg(g4 + i * 32 + lane), bb = __ldg(b4 + i * 32 + lane);
      const float y0 = (v[i].x - mean) * rstd * gg.x + bb.x, y1 = (v[i].y - mean) * rstd * gg.y + bb.y;
      const float y2 = (v[i].z - mean) * rstd * gg.z + bb.z, y3 = (v[i].w - mean) * rstd * gg.w + bb.w;
      o2[i * 32 + lane] = make_uint2(gemm_detail::pack_bf16x2(y0, y1), gemm_detail::pack_bf16x2(y2, y3));
    }
  }
}

// The same with C = 128 * NVEC known at compile time (exactly NVEC float4s of registers): the
// 16-warp tower kernel runs at <= 102 registers.
template <int NVEC, int ROWS = 1>
__device__ __forceinline__ void ln_warp_row_n(const float* __restrict__ x_row, const float* __restrict__ g,
                                              const float* __restrict__ b, __nv_bfloat16* __restrict__ out_row,
                                              int lane) {
  // ROWS consecutive rows with all their loads in flight together (one L2 round trip)
  constexpr int C = NVEC * 128;
  float4 v[ROWS][NVEC];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const float4* xr = reinterpret_cast<const float4*>(x_row + (size_t)r * C);
#pragma unroll
    for (int i = 0; i < NVEC; ++i) v[r][i] = __ldcg(xr + i * 32 + lane);
  }
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) sum += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean[r] = sum / (float)C;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const float d0 = v[r][i].x - mean[r], d1 = v[r][i].y - mean[r], d2 = v[r][i].z - mean[r], d3 = v[r][i].w - mean[r];
      sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    rstd[r] = rsqrtf(sq / (float)C + 1e-5f);
  }
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const float4 gg = __ldg(g4 + i * 32 + lane), bb = __ldg(b4 + i * 32 + lane);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const float y0 = (v[r][i].x - mean[r]) * rstd[r] * gg.x + bb.x, y1 = (v[r][i].y - mean[r]) * rstd[r] * gg.y + bb.y;
      const float y2 = (v[r][i].z - mean[r]) * rstd[r] * gg.z + bb.z, y3 = (v[r][i].w - mean[r]) * rstd[r] * gg.w + bb.w;
      reinterpret_cast<uint2*>(out_row + (size_t)r * C)[i * 32 + lane] =
          make_uint2(gemm_detail::pack_bf16x2(y0, y1), gemm_detail::pack_bf16x2(y2, y3));
    }
  }
}

// enformer_pytorch Attention.forward for one (sequence, head) by one warp, N <= 4 positions:
//   logits[i,j] = (q_i*scale + rcb).k_j + (q_i*scale + rpb).relk[(j-i)+(N-1)]   (relative_shift)
// base = qkv row of the sequence's first position (fp32, leading dimension ld); out_seq = the
// matching row of the bf16 [rows, H*dv] output.
constexpr int kAttnDkPer = 4;      // dk <= 128
template <int N, bool CG>
__device__ __forceinline__ void attn_warp_task(const float* __restrict__ base, int ld,
                                               const float* __restrict__ rcb, const float* __restrict__ rpb,
                                               const float* __restrict__ relk, __nv_bfloat16* __restrict__ out_seq,
                                               int h, int H, int dk, int dv, int lane) {
  const float scale = rsqrtf((float)dk);
  float lg[N][N];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j) lg[i][j] = 0.0f;
#pragma unroll
  for (int t = 0; t < kAttnDkPer; ++t) {
    const int d = lane + 32 * t;
    if (d < dk) {
      const float cb = rcb[h * dk + d], pb = rpb[h * dk + d];
      float q[N], k[N], rk[2 * N - 1];
#pragma unroll
      for (int i = 0; i < N; ++i) {
        q[i] = ld_f1<CG>(base + i * ld + h * dk + d) * scale;
        k[i] = ld_f1<CG>(base + i * ld + H * dk + h * dk + d);
      }
#pragma unroll
      for (int p = 0; p < 2 * N - 1; ++p) rk[p] = relk[((size_t)h * (2 * N - 1) + p) * dk + d];
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) lg[i][j] += (q[i] + cb) * k[j] + (q[i] + pb) * rk[j - i + N - 1];
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) lg[i][j] += __shfl_xor_sync(0xffffffffu, lg[i][j], o);
#pragma unroll
  for (int i = 0; i < N; ++i) {
    float mx = lg[i][0];
#pragma unroll
    for (int j = 1; j < N; ++j) mx = fmaxf(mx, lg[i][j]);
    float den = 0.0f;
#pragma unroll
    for (int j = 0; j < N; ++j) { lg[i][j] = __expf(lg[i][j] - mx); den += lg[i][j]; }
    const float inv = 1.0f / den;
#pragma unroll
    for (int j = 0; j < N; ++j) lg[i][j] *= inv;
  }
  for (int d = lane; d < dv; d += 32) {
    float vv[N];
#pragma unroll
    for (int j = 0; j < N; ++j) vv[j] = ld_f1<CG>(base + j * ld + 2 * H * dk + h * dv + d);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float acc = 0.0f;
#pragma unroll
      for (int j = 0; j < N; ++j) acc += lg[i][j] * vv[j];
      out_seq[(size_t)i * (H * dv) + h * dv + d] = __float2bfloat16_rn(acc);
    }
  }
}

// LayerNorm of ROWS rows by one warp with every row's loads in flight together (the tower's row
// items are latency-bound: 8 warps per CTA).  NVEC = C / 128 float4s per lane; arithmetic of
// ln_warp_row.
template <int ROWS, int NVEC>
__device__ __noinline__ void ln_warp_rows(const float* __restrict__ x, const float* __restrict__ g,
                                             const float* __restrict__ b, __nv_bfloat16* __restrict__ out,
                                             int valid_rows, int lane) {
  constexpr int C = NVEC * 128;
  float4 v[ROWS][NVEC];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)(r < valid_rows ? r : 0) * C);
#pragma unroll
    for (int i = 0; i < NVEC; ++i) v[r][i] = __ldcg(xr + i * 32 + lane);
  }
  float mean[ROWS], rstd[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) sum += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    mean[r] = sum / (float)C;
    float sq = 0.0f;
#pragma unroll
    for (int i = 0; i < NVEC; ++i) {
      const float d0 = v[r][i].x - mean[r], d1 = v[r][i].y - mean[r], d2 = v[r][i].z - mean[r], d3 = v[r][i].w - mean[r];
      sq += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    rstd[r] = rsqrtf(sq / (float)C + 1e-5f);
  }
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < NVEC; ++i) {
    const float4 gg = __ldg(g4 + i * 32 + lane), bb = __ldg(b4 + i * 32 + lane);
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      if (r < valid_rows) {
        const float y0 = (v[r][i].x - mean[r]) * rstd[r] * gg.x + bb.x, y1 = (v[r][i].y - mean[r]) * rstd[r] * gg.y + bb.y;
        const float y2 = (v[r][i].z - mean[r]) * rstd[r] * gg.z + bb.z, y3 = (v[r][i].w - mean[r]) * rstd[r] * gg.w + bb.w;
        reinterpret_cast<uint2*>(out + (size_t)r * C)[i * 32 + lane] =
            make_uint2(gemm_detail::pack_bf16x2(y0, y1), gemm_detail::pack_bf16x2(y2, y3));
      }
    }
  }
}

// Attention of ONE sequence (N positions, all heads) by one warp, every load of a stage in flight
// together: q / k rows as float4s (T = H*dk/128 per lane; a float4 lies inside one head, the dk/4
// lanes of a head reduce by xor shuffles), softmax weights through `s_w` (H*N*N floats of this
// warp), then the v rows (U = H*dv/128 float4s per lane).  Same formula as attn_warp_task; the
// summation order over dk differs (4 consecutive channels per lane).
// `h0` / `Hs`: the warp handles heads [h0, h0 + Hs) of the H heads (T and U count the float4s of that
// subset: Hs*dk/128 and Hs*dv/128); the 16-warp tower gives each sequence to two warps.
template <int N, int T, int U>
__device__ __noinline__ void attn_warp_seq(const float* __restrict__ base, int ld, const float* __restrict__ rcb,
                                              const float* __restrict__ rpb, const float* __restrict__ relk,
                                              __nv_bfloat16* __restrict__ out_seq, int H, int dk, int dv, int lane,
                                              float* __restrict__ s_w, int h0 = 0) {
  const float* qb = base + h0 * dk;                  // q columns of the subset
  const float* kb = base + H * dk + h0 * dk;         // k columns
  const float* vb = base + 2 * H * dk + h0 * dv;     // v columns
  rcb += h0 * dk; rpb += h0 * dk;
  relk += (size_t)h0 * (2 * N - 1) * dk;
  __nv_bfloat16* ob = out_seq + h0 * dv;
  const int out_ld = H * dv;
  const float scale = rsqrtf((float)dk);
  const int lph = dk >> 2;                   // lanes per head
  float4 q[N][T], k[N][T];
#pragma unroll
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int t = 0; t < T; ++t) {
      q[i][t] = __ldcg(reinterpret_cast<const float4*>(qb + (size_t)i * ld) + lane + 32 * t);
      k[i][t] = __ldcg(reinterpret_cast<const float4*>(kb + (size_t)i * ld) + lane + 32 * t);
    }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int f = lane + 32 * t;             // float4 index within the H*dk row
    const int h = (4 * f) / dk, d = (4 * f) % dk;
    const float4 cb = __ldg(reinterpret_cast<const float4*>(rcb) + f);
    const float4 pb = __ldg(reinterpret_cast<const float4*>(rpb) + f);
    float4 rk[2 * N - 1];
#pragma unroll
    for (int p = 0; p < 2 * N - 1; ++p)
      rk[p] = __ldg(reinterpret_cast<const float4*>(relk + ((size_t)h * (2 * N - 1) + p) * dk + d));
    float lg[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
      const float qx = q[i][t].x * scale, qy = q[i][t].y * scale, qz = q[i][t].z * scale, qw = q[i][t].w * scale;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const float4 kk = k[j][t], rr = rk[j - i + N - 1];
        float acc = (qx + cb.x) * kk.x + (qx + pb.x) * rr.x;
        acc += (qy + cb.y) * kk.y + (qy + pb.y) * rr.y;
        acc += (qz + cb.z) * kk.z + (qz + pb.z) * rr.z;
        acc += (qw + cb.w) * kk.w + (qw + pb.w) * rr.w;
        lg[i][j] = acc;
      }
    }
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j)
        for (int o = lph >> 1; o > 0; o >>= 1) lg[i][j] += __shfl_xor_sync(0xffffffffu, lg[i][j], o);
#pragma unroll
    for (int i = 0; i < N; ++i) {
      float mx = lg[i][0];
#pragma unroll
      for (int j = 1; j < N; ++j) mx = fmaxf(mx, lg[i][j]);
      float den = 0.0f;
#pragma unroll
      for (int j = 0; j < N; ++j) { lg[i][j] = __expf(lg[i][j] - mx); den += lg[i][j]; }
      const float inv = 1.0f / den;
      if ((lane & (lph - 1)) == 0) {
#pragma unroll
        for (int j = 0; j < N; ++j) s_w[(h * N + i) * N + j] = lg[i][j] * inv;
      }
    }
  }
  __syncwarp();
  constexpr int UC = U <= 6 ? U : 6;         // v float4s in flight per lane and position
  static_assert(U % UC == 0, "U must be a multiple of the chunk");
#pragma unroll 1
  for (int u0 = 0; u0 < U; u0 += UC) {
    float4 v[N][UC];
#pragma unroll
    for (int j = 0; j < N; ++j)
#pragma unroll
      for (int u = 0; u < UC; ++u)
        v[j][u] = __ldcg(reinterpret_cast<const float4*>(vb + (size_t)j * ld) + lane + 32 * (u0 + u));
#pragma unroll
    for (int u = 0; u < UC; ++u) {
      const int f = lane + 32 * (u0 + u);
      const int h = (4 * f) / dv;
#pragma unroll
      for (int i = 0; i < N; ++i) {
        float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int j = 0; j < N; ++j) {
          const float w = s_w[(h * N + i) * N + j];
          acc.x += w * v[j][u].x; acc.y += w * v[j][u].y; acc.z += w * v[j][u].z; acc.w += w * v[j][u].w;
        }
        reinterpret_cast<uint2*>(ob + (size_t)i * out_ld)[f] =
            make_uint2(gemm_detail::pack_bf16x2(acc.x, acc.y), gemm_detail::pack_bf16x2(acc.z, acc.w));
      }
    }
  }
  __syncwarp();
}

// out-of-line fall-backs (general channel counts / head shapes): kept out of the kernel's register
// allocation
__device__ __noinline__ void ln_rows_generic(const float* x, const float* g, const float* b, __nv_bfloat16* out, int C,
                                             int rows, int lane) {
  for (int rr = 0; rr < rows; ++rr) ln_warp_row<true>(x + (size_t)rr * C, g, b, out + (size_t)rr * C, C, lane);
}
template <int N>
__device__ __noinline__ void attn_task_generic(const float* base, int ld, const float* rcb, const float* rpb,
                                               const float* relk, __nv_bfloat16* out_seq, int h, int H, int dk, int dv,
                                               int lane) {
  attn_warp_task<N, true>(base, ld, rcb, rpb, relk, out_seq, h, H, dk, dv, lane);
}

struct Item { int j, q, r, c; };

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// 16 fp32 columns (chunk `cis` of the 2 in a 32-column slab row) of one row
__device__ __forceinline__ void slab_write16_f32(uint8_t* row_base, int x7, int cis, const float* v) {
#pragma unroll
  for (int u = 0; u < 4; ++u)
    *reinterpret_cast<float4*>(row_base + (((cis * 4 + u) ^ x7) << 4)) =
        make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
}

// EW = epilogue warps: 8 (thread = row x column half, 32-column TMEM chunks, two slabs per half) or 16
// (thread = row x column quarter, 16-column chunks, one slab per quarter, <= 102 registers): the
// row phases and the GEMM epilogues are latency-bound per warp, so twice the warps per scheduler
// shortens every hop of the dependency chain.
template <int NPOS, int kBN, int EW = 8>
__global__ void __launch_bounds__(64 + 32 * EW + 64, 1)
tower_kernel(const __grid_constant__ CUtensorMap tm_hn, const __grid_constant__ CUtensorMap tm_ao,
             const __grid_constant__ CUtensorMap tm_u, const __grid_constant__ CUtensorMap tm_qkv,
             const __grid_constant__ CUtensorMap tm_xt, const TowerArgs a) {
  using Cfg = tower::Cfg<kBN>;
  constexpr int kStages = Cfg::kStages;
  constexpr int kChunks = kBN / 64;                 // 32-column accumulator chunks per epilogue thread (EW = 8)
  static_assert(EW == 8 || (EW == 16 && kBN == 256), "16 epilogue warps need 256-wide tiles");
  auto ebar = [] { asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory"); };   // all epilogue threads
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint8_t* staging = smem + kStages * Cfg::kStageBytes;            // [half][buf] slabs, 1024-aligned
  float* s_bias = reinterpret_cast<float*>(staging + kStagingBytes);   // [2][kBN]
  float* s_attn_w = s_bias + 2 * kBN;                                  // [kEpiWarps][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kStagingBytes + Cfg::kBiasBytes + Cfg::kAttnWBytes);
  uint64_t* full_bar = bars;                       // [kStages]   (leader's copy is the live one)
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;        // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;   // [2]         (leader's copy is the live one)
  uint64_t* rin_bar = bars + 2 * kStages + 4;      // [half][buf]
  uint64_t* rout_bar = bars + 2 * kStages + 8;     // [half][buf]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)gemm2::cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1);
  const int n_pairs = (int)(gridDim.x >> 1);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_hn);
    ptx::prefetch_tmap(&tm_ao);
    ptx::prefetch_tmap(&tm_u);
    ptx::prefetch_tmap(&tm_qkv);
    ptx::prefetch_tmap(&tm_xt);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        ptx::mbar_init(&full_bar[i], 1);
        ptx::mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull_bar[i], 1);
        ptx::mbar_init(&tempty_bar[i], 2 * EW);
      }
      for (int i = 0; i < 4; ++i) {
        ptx::mbar_init(&rin_bar[i], 1);
        ptx::mbar_init(&rout_bar[i], kEpiWarps / 2);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    gemm2::tmem_alloc_cg<2>(tmem_slot, Cfg::kTmemCols);
  }
  ptx::tc_fence_before();
  gemm2::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  auto decode = [&](int id) {
    Item it;
    const int blk_items = a.items_per_block;
    if (id >= a.n_blocks * blk_items) {
      const int rem = id - a.n_blocks * blk_items;
      it.j = a.n_blocks; it.q = kPhasesPerBlock;
      it.r = rem / kSubItems; it.c = rem % kSubItems;
      return it;
    }
    it.j = id / blk_items;
    const int rem = id % blk_items;
    int q = 0;
#pragma unroll
    for (int k = 1; k < kPhasesPerBlock; ++k) q += (rem >= a.ph[k].start) ? 1 : 0;
    it.q = q;
    const int off = rem - a.ph[q].start;
    it.r = off / a.ph[q].items_per_rt;
    it.c = off % a.ph[q].items_per_rt;
    return it;
  };
  // counter the item waits on (nullptr: none) and the value it must reach
  auto dep_flag = [&](const Item& it, unsigned* target) -> const unsigned* {
    if (it.j == 0 && it.q == 0) return nullptr;
    const int qp = (it.q == 0 || it.q == kPhasesPerBlock) ? kPhasesPerBlock - 1 : it.q - 1;
    const int jp = (it.q == 0 || it.q == kPhasesPerBlock) ? it.j - 1 : it.j;
    *target = (unsigned)(a.ph[qp].items_per_rt * a.ph[qp].signals);
    return a.flags + (size_t)(jp * kPhasesPerBlock + qp) * a.RT + it.r;
  };
  auto own_flag = [&](const Item& it) -> unsigned* {
    return a.flags + (size_t)(it.j * kPhasesPerBlock + it.q) * a.RT + it.r;
  };

  // split-K order flag of the item's column tile (store threads only)
  auto sk_flag = [&](const Item& it, const Phase& ph) -> unsigned* {
    return a.sk_flags + ((size_t)it.j * a.RT + it.r) * kMaxSplitTiles + it.c / ph.ksplit;
  };
  auto stamp = [&](int id, int slot) {
    if (a.trace != nullptr && rank == 0) a.trace[(size_t)id * 8 + slot] = slot == 6 ? (unsigned long long)pair : globaltimer_ns();
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs of the pair) =====================
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      constexpr uint32_t tx_bytes = 2u * (uint32_t)Cfg::kStageBytes;
      for (int id = pair; id < a.total_items; id += n_pairs) {
        const Item it = decode(id);
        const Phase& ph = a.ph[it.q];
        if (ph.type != PH_GEMM) continue;
        const CUtensorMap* mA = ph.a_map == A_HN ? &tm_hn : (ph.a_map == A_AO ? &tm_ao : &tm_u);
        const CUtensorMap* mW = &a.blocks[it.j].w[ph.w_idx];
        const int row0 = it.r * kTileRows + rank * kBM;
        const int wrow0 = (it.c / ph.ksplit) * kBN + rank * (kBN / 2);
        const int kb0 = (it.c % ph.ksplit) * ph.kblocks;        // first K block of this slice
        const int pre = ph.kblocks < kStages ? ph.kblocks : kStages;
        // weights of the first ring stages: no dependency, request them before the wait
        uint32_t s2 = stage, p2 = phase;
        for (int kb = 0; kb < pre; ++kb) {
          twait(&empty_bar[s2], p2 ^ 1);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[s2], tx_bytes);
          gemm2::tma_load_2d_cg<2>(stage_base + s2 * Cfg::kStageBytes + Cfg::kABytes, mW, &full_bar[s2], (kb0 + kb) * kBK, wrow0);
          if (++s2 == kStages) { s2 = 0; p2 ^= 1; }
        }
        unsigned target = 0;
        const unsigned* dep = dep_flag(it, &target);
        stamp(id, 0);
        wait_flag(dep, target);
        stamp(id, 1);
        stamp(id, 6);
        fence_proxy_async_all();
        for (int kb = 0; kb < pre; ++kb) {
          gemm2::tma_load_2d_cg<2>(stage_base + stage * Cfg::kStageBytes, mA, &full_bar[stage], (kb0 + kb) * kBK, row0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        for (int kb = pre; kb < ph.kblocks; ++kb) {
          twait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
          if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
          gemm2::tma_load_2d_cg<2>(sa, mA, &full_bar[stage], (kb0 + kb) * kBK, row0);
          gemm2::tma_load_2d_cg<2>(sa + Cfg::kABytes, mW, &full_bar[stage], (kb0 + kb) * kBK, wrow0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kBM * 2, kBN);
      uint32_t stage = 0, phase = 0, acc_stage = 0, acc_phase = 0;
      for (int id = pair; id < a.total_items; id += n_pairs) {
        const Item it = decode(id);
        const Phase& ph = a.ph[it.q];
        if (ph.type != PH_GEMM) continue;
        twait(&tempty_bar[acc_stage], acc_phase ^ 1);
        ptx::tc_fence_after();
        if (lane == 0) stamp(id, 2);
        const uint32_t tmem_d = tmem_base + acc_stage * kBN;
        for (int kb = 0; kb < ph.kblocks; ++kb) {
          twait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t sa = ptx::smem_u32(stage_base + stage * Cfg::kStageBytes);
            const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
            const uint64_t db = ptx::make_kmajor_sw128_desc(sa + Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              gemm2::umma_bf16_cg<2>(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            gemm2::umma_commit_cg<2>(&empty_bar[stage]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) gemm2::umma_commit_cg<2>(&tfull_bar[acc_stage]);
        __syncwarp();
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp < 2 + EW) {
    // ===================== epilogue math / row-wise items =====================
    const int ew = warp - 2;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may read
    const int half = ew >> 2;            // column half (EW = 8) / column quarter (EW = 16)
    const int etid = threadIdx.x - 64;
    const int r = quad * 32 + lane;      // tile row owned by this thread
    const int x7 = r & 7;
    uint8_t* my_bufs = staging + half * 2 * kSlabBytes + r * 128;
    uint64_t* my_rin = rin_bar + half * 2;
    uint64_t* my_rout = rout_bar + half * 2;
    uint32_t acc_stage = 0, acc_phase = 0, job = 0;
    for (int id = pair; id < a.total_items; id += n_pairs) {
      const Item it = decode(id);
      const Phase& ph = a.ph[it.q];
      if (ph.type == PH_GEMM) {
        const bool out_f32 = ph.out_kind != OUT_BF16_RELU;
        const int n0 = (it.c / ph.ksplit) * kBN;
        const float* bias = (it.c % ph.ksplit) == 0 ? a.blocks[it.j].bias[ph.w_idx] : nullptr;   // once per tile
        float* P = s_bias + acc_stage * kBN;
        for (int i = etid; i < kBN; i += 32 * EW) P[i] = (bias != nullptr && n0 + i < ph.n_cols) ? bias[n0 + i] : 0.0f;
        ebar();
        if constexpr (EW == 16) {
          // quarter `half` = 64 columns: 4 chunks of 16; fp32 = two 32-column slab jobs, bf16 = one
          uint8_t* buf = staging + half * kSlabBytes + r * 128;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * kBN + half * 64;
          uint32_t raw[2][16];
          twait(&tfull_bar[acc_stage], acc_phase);
          ptx::tc_fence_after();
          if (etid == 0) stamp(id, 3);
          gemm2::tmem_ld_32x16(taddr, raw[0]);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const bool job_start = out_f32 ? (c & 1) == 0 : c == 0;
            const bool job_end = out_f32 ? (c & 1) == 1 : c == 3;
            if (job_start) twait(&rin_bar[half], job & 1);
            float v[16], pv[16];
            ptx::tmem_ld_wait();
            if (c + 1 < 4) gemm2::tmem_ld_32x16(taddr + (c + 1) * 16, raw[(c + 1) & 1]);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[c & 1][i]);
            if (c == 3) {                  // accumulator fully read: hand it back to the MMA warp
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) gemm2::mbar_arrive_leader<2>(&tempty_bar[acc_stage]);
            }
            gemm2::load_param16(P + half * 64 + c * 16, pv);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += pv[i];
            if (out_f32) {
              slab_write16_f32(buf, x7, c & 1, v);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.0f);
              gemm2::slab_write16(buf, x7, c, v);
            }
            if (job_end) {
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive(&rout_bar[half]);
              ++job;
            }
          }
        } else {
        const int slab_chunks = out_f32 ? 1 : 2;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * kBN + half * (kBN / 2);
        uint32_t raw[2][32];
        twait(&tfull_bar[acc_stage], acc_phase);
        ptx::tc_fence_after();
        if (etid == 0) stamp(id, 3);
        ptx::tmem_ld_32x32(taddr, raw[0]);
        uint8_t* buf0 = nullptr;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int cis = c % slab_chunks;
          if (cis == 0) {
            buf0 = my_bufs + (job & 1) * kSlabBytes;
            twait(&my_rin[job & 1], (job >> 1) & 1);
          }
          float v[32], pv[32];
          ptx::tmem_ld_wait();
          if (c + 1 < kChunks) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, raw[(c + 1) & 1]);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[c & 1][i]);
          if (c + 1 == kChunks) {        // accumulator fully read: hand it back to the MMA warp
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) gemm2::mbar_arrive_leader<2>(&tempty_bar[acc_stage]);
          }
          gemm_detail::load_param32(P + half * (kBN / 2) + c * 32, pv);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += pv[i];
          if (ph.out_kind == OUT_BF16_RELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
          gemm2::slab_write(buf0, x7, out_f32, cis, v);
          if (cis == slab_chunks - 1) {
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&my_rout[job & 1]);
            ++job;
          }
        }
        }
        if (etid == 0) stamp(id, 4);
        if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
      } else {
        // LN / attention / BN+GELU over kSubRows rows of this CTA's half of the row tile
        if (etid == 0) {
          unsigned target = 0;
          const unsigned* dep = dep_flag(it, &target);
          stamp(id, 0);
          wait_flag(dep, target);
          stamp(id, 1);
          stamp(id, 6);
        }
        ebar();
        const int row_base = it.r * kTileRows + rank * kBM + it.c * kSubRows;
        if (ph.type == PH_LN) {
          const BlockParams& bp = a.blocks[it.j];
          constexpr int kRowsPerWarp = kSubRows / EW;                 // consecutive rows of one warp
          const int row = row_base + ew * kRowsPerWarp;
          const int valid = a.R - row < kRowsPerWarp ? a.R - row : kRowsPerWarp;
          if (valid > 0) {
            const float* xr = a.xt + (size_t)row * a.C;
            __nv_bfloat16* hr = a.hn + (size_t)row * a.C;
            if (a.C == 1536) {     // (two rows in flight together, ln_warp_row_n<12, 2>, spills 430 bytes here)
#pragma unroll 1
              for (int rr = 0; rr < valid; ++rr)
                ln_warp_row_n<12>(xr + (size_t)rr * a.C, bp.ln_g[ph.w_idx], bp.ln_b[ph.w_idx], hr + (size_t)rr * a.C, lane);
            } else if (a.C == 384) {
#pragma unroll 1
              for (int rr = 0; rr < valid; ++rr)
                ln_warp_row_n<3>(xr + (size_t)rr * a.C, bp.ln_g[ph.w_idx], bp.ln_b[ph.w_idx], hr + (size_t)rr * a.C, lane);
            } else {
              ln_rows_generic(xr, bp.ln_g[ph.w_idx], bp.ln_b[ph.w_idx], hr, a.C, valid, lane);
            }
            if (a.discard_qkv) {
              // dead now: LN2 -> these rows of the attention output (read by the out-projection, complete);
              // LN1 of block j > 0 -> these rows of the previous block's FFN hidden tensor (FF2 complete)
              const int row_bytes = ph.w_idx == 1 ? a.H * a.dv * 2 : 2 * a.C * 2;
              const char* dead = ph.w_idx == 1 ? reinterpret_cast<const char*>(a.ao + (size_t)row * a.H * a.dv)
                                               : (it.j > 0 && a.u != nullptr ? reinterpret_cast<const char*>(a.u + (size_t)row * 2 * a.C) : nullptr);
              if (dead != nullptr && (row_bytes & 127) == 0) {
                const int lines = valid * (row_bytes >> 7);
                for (int c = lane; c < lines; c += 32)
                  asm volatile("discard.global.L2 [%0], 128;" ::"l"(dead + (size_t)c * 128) : "memory");
              }
            }
          }
        } else if (ph.type == PH_ATTN) {
          const BlockParams& bp = a.blocks[it.j];
          const int hdk = a.H * a.dk, hdv = a.H * a.dv;
          const bool pow2 = (a.dk & (a.dk - 1)) == 0;
          const bool fast = a.attn_fast != 0 && NPOS <= 2 && pow2 && a.dk >= 4 && a.dk <= 128 && a.dv % 4 == 0 && hdk == 512 &&
                            (hdv == 1536 || hdv == 384) && a.H * NPOS * NPOS <= 128;
          if (fast && EW == 16 && NPOS == 2 && hdv == 1536) {
            // two warps per sequence, half of the heads each (32 + 48 registers of loads in flight)
            float* sw = s_attn_w + ew * 128;
            const int row = row_base + (ew >> 1) * NPOS;
            if constexpr (NPOS == 2) {
              if (row < a.R)
                attn_warp_seq<2, 2, 6>(a.qkv + (size_t)row * a.nqkv, a.nqkv, bp.rcb, bp.rpb, bp.relk,
                                       a.ao + (size_t)row * hdv, a.H, a.dk, a.dv, lane, sw, (ew & 1) * (a.H / 2));
            }
          } else if (fast) {
            float* sw = s_attn_w + ew * 128;
#pragma unroll 1
            for (int sl = ew; sl < kSubRows / NPOS; sl += EW) {
              const int row = row_base + sl * NPOS;
              if (row >= a.R) continue;
              const float* base = a.qkv + (size_t)row * a.nqkv;
              __nv_bfloat16* o = a.ao + (size_t)row * hdv;
              if constexpr (NPOS <= 2) {
                if (hdv == 1536) attn_warp_seq<NPOS, 4, 12>(base, a.nqkv, bp.rcb, bp.rpb, bp.relk, o, a.H, a.dk, a.dv, lane, sw);
                else attn_warp_seq<NPOS, 4, 3>(base, a.nqkv, bp.rcb, bp.rpb, bp.relk, o, a.H, a.dk, a.dv, lane, sw);
                // The sequence's qkv rows are dead now (this warp was their only reader; the next block's QKV
                // GEMM rewrites them whole): drop the lines from L2 instead of letting them be written back --
                // qkv is 26 MB of fp32 per block, the largest share of the kernel's dead write-backs.
                if (a.discard_qkv && (a.nqkv & 31) == 0) {
                  const int lines = NPOS * a.nqkv / 32;          // 128-byte lines of the NPOS contiguous rows
                  for (int c = lane; c < lines; c += 32)
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + (size_t)c * 32) : "memory");
                  if ((a.C & 63) == 0) {                         // LN1's rows: read by the QKV GEMM, complete
                    const char* hdead = reinterpret_cast<const char*>(a.hn + (size_t)row * a.C);
                    for (int c = lane; c < NPOS * (a.C >> 6); c += 32)
                      asm volatile("discard.global.L2 [%0], 128;" ::"l"(hdead + (size_t)c * 128) : "memory");
                  }
                }
              }
            }
          } else {
            const int tasks = (kSubRows / NPOS) * a.H;
#pragma unroll 1
            for (int t = ew; t < tasks; t += EW) {
              const int row = row_base + (t / a.H) * NPOS;
              if (row < a.R)
                attn_task_generic<NPOS>(a.qkv + (size_t)row * a.nqkv, a.nqkv, bp.rcb, bp.rpb, bp.relk,
                                        a.ao + (size_t)row * hdv, t % a.H, a.H, a.dk, a.dv, lane);
            }
          }
        } else {
          // pointwise ConvBlock operand: hn = GELU(BN(x))
#pragma unroll 1
          for (int rr = ew; rr < kSubRows; rr += EW) {
            const int row = row_base + rr;
            if (row >= a.R) continue;
            const float4* xr = reinterpret_cast<const float4*>(a.xt + (size_t)row * a.C);
            uint2* o2 = reinterpret_cast<uint2*>(a.hn + (size_t)row * a.C);
            for (int i = lane; i < (a.C >> 2); i += 32) {
              const float4 x = __ldcg(xr + i);
              const float4 s = __ldg(reinterpret_cast<const float4*>(a.bn_s) + i);
              const float4 t = __ldg(reinterpret_cast<const float4*>(a.bn_t) + i);
              const float y0 = gemm2::gelu_tanh(x.x * s.x + t.x), y1 = gemm2::gelu_tanh(x.y * s.y + t.y);
              const float y2 = gemm2::gelu_tanh(x.z * s.z + t.z), y3 = gemm2::gelu_tanh(x.w * s.w + t.w);
              o2[i] = make_uint2(gemm_detail::pack_bf16x2(y0, y1), gemm_detail::pack_bf16x2(y2, y3));
            }
          }
        }
        if (it.q < kPhasesPerBlock) {
          // bar.sync orders every thread's stores before thread 0's gpu-scope release (cumulativity)
          ebar();
          if (etid == 0) {
            stamp(id, 4);
            red_release_gpu_add(own_flag(it), 1u);
            stamp(id, 5);
          }
        }
      }
    }
  } else {
    // ===================== slab store warp (one per column half) =====================
    const int half = warp - (2 + EW);
    if constexpr (EW == 16) {
      // one slab per column quarter; this thread serves quarters 2*half and 2*half+1 in turn
      if (ptx::elect_one()) {
        ptx::mbar_arrive(&rin_bar[2 * half]);
        ptx::mbar_arrive(&rin_bar[2 * half + 1]);
        uint32_t j = 0;                    // slab jobs done per quarter
        for (int id = pair; id < a.total_items; id += n_pairs) {
          const Item it = decode(id);
          const Phase& ph = a.ph[it.q];
          if (ph.type != PH_GEMM) continue;
          const bool out_f32 = ph.out_kind != OUT_BF16_RELU;
          const int jobs = out_f32 ? 2 : 1;
          const CUtensorMap* mO = ph.out_kind == OUT_F32_STORE ? &tm_qkv : (ph.out_kind == OUT_F32_REDUCE ? &tm_xt : &tm_u);
          const int row0 = it.r * kTileRows + rank * kBM;
          const int ks = it.c % ph.ksplit, ctile = it.c / ph.ksplit;
          if (ks > 0) {                      // the earlier K slices of this tile have been added
            wait_flag(sk_flag(it, ph), 4u * (unsigned)ks);
            fence_proxy_async_all();
          }
          for (int sj = 0; sj < jobs; ++sj, ++j) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int grp = 2 * half + q;
              twait(&rout_bar[grp], j & 1);
              const int col = ctile * kBN + grp * 64 + sj * 32;
              if (ph.out_kind == OUT_F32_REDUCE) tma_reduce_add_2d(mO, staging + grp * kSlabBytes, col, row0);
              else tma_store_2d(mO, staging + grp * kSlabBytes, col, row0);
              gemm2::bulk_commit();
            }
            if (sj + 1 < jobs) {             // the slabs are free once read; the last job waits below
              gemm2::bulk_wait_read0();
              ptx::mbar_arrive(&rin_bar[2 * half]);
              ptx::mbar_arrive(&rin_bar[2 * half + 1]);
            }
          }
          // the tile's stores are complete (not just read): free the slabs, publish
          gemm2::bulk_wait_all();
          ptx::mbar_arrive(&rin_bar[2 * half]);
          ptx::mbar_arrive(&rin_bar[2 * half + 1]);
          fence_proxy_async_all();
          if (ph.ksplit > 1) red_release_gpu_add(sk_flag(it, ph), 1u);
          red_release_gpu_add(own_flag(it), 1u);
          if (half == 0) stamp(id, 5);
        }
      }
    } else
    if (ptx::elect_one()) {
      uint8_t* bufs = staging + half * 2 * kSlabBytes;
      uint64_t* my_rin = rin_bar + half * 2;
      uint64_t* my_rout = rout_bar + half * 2;
      ptx::mbar_arrive(&my_rin[0]);      // both slab buffers start free
      ptx::mbar_arrive(&my_rin[1]);
      uint32_t j = 0;
      for (int id = pair; id < a.total_items; id += n_pairs) {
        const Item it = decode(id);
        const Phase& ph = a.ph[it.q];
        if (ph.type != PH_GEMM) continue;
        const bool out_f32 = ph.out_kind != OUT_BF16_RELU;
        const int slab_cols = out_f32 ? 32 : 64;
        const int slabs = (kBN / 2) / slab_cols;
        const CUtensorMap* mO = ph.out_kind == OUT_F32_STORE ? &tm_qkv : (ph.out_kind == OUT_F32_REDUCE ? &tm_xt : &tm_u);
        const int row0 = it.r * kTileRows + rank * kBM;
        const int ks = it.c % ph.ksplit, ctile = it.c / ph.ksplit;
        if (ks > 0) {                        // the earlier K slices of this tile have been added
          wait_flag(sk_flag(it, ph), 4u * (unsigned)ks);
          fence_proxy_async_all();
        }
        for (int s = 0; s < slabs; ++s, ++j) {
          twait(&my_rout[j & 1], (j >> 1) & 1);
          const uint8_t* buf = bufs + (j & 1) * kSlabBytes;
          const int col = ctile * kBN + half * (kBN / 2) + s * slab_cols;
          if (ph.out_kind == OUT_F32_REDUCE) tma_reduce_add_2d(mO, buf, col, row0);
          else tma_store_2d(mO, buf, col, row0);
          gemm2::bulk_commit();
          // two stores in flight: the PREVIOUS job's slab is free once at most one group is still
          // reading (waiting for this job's own read here would serialise store and math)
          if (s > 0) {
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            ptx::mbar_arrive(&my_rin[(j - 1) & 1]);
          }
        }
        // the tile's stores are complete (not just read): free the last slab, publish
        gemm2::bulk_wait_all();
        ptx::mbar_arrive(&my_rin[(j - 1) & 1]);
        fence_proxy_async_all();
        if (ph.ksplit > 1) red_release_gpu_add(sk_flag(it, ph), 1u);
        red_release_gpu_add(own_flag(it), 1u);
        if (half == 0) stamp(id, 5);
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  gemm2::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    gemm2::tmem_dealloc_cg<2>(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace tower
}  // namespace svdd
