// Stage 1 as ONE persistent kernel: the whole MDLM denoiser (embed conv, 20 dilated conv
// layers with time bias / LayerNorm / ReLU / residual, final 1x1 convs) for one sequence per
// CTA, activations resident on the SM for the entire network (models/dnaconv.py:176-210).
//
// Why: as 22 separate launches the layers were bound by the L2 -> SM fabric, not the tensor
// pipe: every layer re-read its bf16 operand once per tap (9x) and its fp32 residual stream
// through L2, for 25 600 rows that fill the chip for barely two tiles per SM (35 us per layer,
// 217 TFLOP/s).  Here
//   * the LayerNorm'd bf16 operand of the current layer lives in shared memory, K-major with
//     the 128-byte swizzle, rows = positions (plus zero rows either side = the conv's zero
//     padding).  A dilated tap is just a ROW OFFSET of the UMMA shared-memory descriptor:
//     the swizzle is a function of the absolute smem address, so a descriptor may start at
//     any 128-byte row (probed on B200: tools/umma_rowoffset_probe.cu) -- the operand is
//     never re-read, never im2col'ed and never leaves the SM;
//   * the fp32 residual stream lives in tensor memory next to the accumulators (2 x 128
//     accumulator columns + 2 x 128 residual columns = all 512 TMEM columns), read with
//     tcgen05.ld and written back with tcgen05.st by the thread that owns the row;
//   * the only stream from L2 is the weights (one 16 KB TMA tile per (tap, K half), used by
//     both 128-row tiles), through a 6-deep mbarrier ring that runs ahead across layers.
//
// Roles (320 threads): warp 0 weight producer (TMA), warp 1 MMA issuer, warps 2-9 epilogue
// (thread = one row of one 128-row tile: the whole LayerNorm is thread-local, no exchange).
// The epilogue cannot overlap the MMAs of the next layer (every tap reads rows this epilogue
// writes).  Measured: 16 epilogue warps with half a row each (96 registers, spills, two
// block barriers per layer) were slower (0.37 ms vs 0.30 ms per B=128 pass).
// Sequences longer than 128 use two row tiles (L <= 256); sequences of at most 64 positions
// are processed two per CTA (one per tile; the zero gap between them exceeds every in-range
// tap offset).  In that mode tile 1's sequence sits in tile rows 64..127, so the real rows of
// the two tiles belong to four different TMEM lane quadrants = four epilogue warps on four
// different schedulers; epilogue warps that own no real row skip the arithmetic and only keep
// the barriers in step (before: all eight ran the full epilogue, the four useful ones two
// to a scheduler).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace svdd {
namespace denf {

constexpr int kH = 128;                 // hidden width
constexpr int kTaps = 9;
constexpr int kMaxLayers = 64;
constexpr int kStages = 6;
constexpr int kStageBytes = kH * 64 * 2;          // one (tap, K half) weight tile: 128 x 64 bf16
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kParamBytes = 2 * 4 * kH * 4;       // [parity][bias, tbias, gamma, beta][128]
constexpr int kW2Bytes = (kVocab * kH + 8) * 4;
constexpr int kXchBytes = 2 * kEpiThreads * 8 * 4;   // split epilogue: [parity][thread][8] floats
constexpr int kBarBytes = 256;

struct Args {
  const void* tokens;        // [n_rows, L] uint8 or int64
  const float* embed_w;      // [9][5][128]
  const float* embed_b;      // [128]
  const float* conv_b;       // [n_layers][128]
  const float* ln_g;         // [n_layers][128]
  const float* ln_b;         // [n_layers][128]
  const float* time_bias;    // [n_layers][128]
  const float* fc0_b;        // [128]
  const float* fc2_w;        // [5][128]
  const float* fc2_b;        // [5]
  float* logits;             // [n_rows, L, 5]
  int64_t n_rows;
  int L;
  int n_layers;
  int pad_before;            // zero rows in front of the first tile
  int a_rows;                // rows of one operand plane (multiple of 8)
  int two_seq;               // L <= 64: tile m holds sequence (2*iter + m)
  int split;                 // two_seq only: two epilogue threads per row (64 channels each)
  // two_seq + split: COMBINED mode.  A second pair of operand planes holds BOTH sequences of the
  // item 64 rows apart (A rows 0..L-1, B rows 64..64+L-1) so that ONE 128-row MMA per tap serves
  // both, for every tap whose offset fits the zero gap between them (|o| <= 64 - L).  Taps with
  // larger offsets keep one MMA per sequence on the isolated planes and accumulate into their own
  // TMEM columns (the rows of the other sequence hold garbage there and are never read).
  // two_seq + split: INTERLEAVED mode (supersedes the combined mode below).  The two sequences of an
  // item share ONE pair of operand planes with their positions interleaved: plane row 2p = sequence A,
  // position p; row 2p + 1 = sequence B, position p.  A tap offset o is then a row offset of 2 o, which
  // keeps A's rows reading A's rows and B's reading B's for EVERY offset (parity is preserved), and the
  // zero rows either side of the 2 L real rows are the conv padding of both.  One 128-row MMA per tap
  // serves both sequences for all taps: no isolated planes, no extra accumulators (TMEM: accumulator
  // + residual = 256 columns), L = 50: 141 MMA groups per item (combined mode 173, two tiles 281).
  int ilv;
  int cmb;                   // combined mode on
  int cmb_max;               // largest |tap offset| served by the combined planes
  int pad_c;                 // zero rows in front of a combined plane
  int c_rows;                // rows of one combined plane (multiple of 8)
  int iso_b;                 // isolated planes: row distance between the two sequences (legacy: 128)
  unsigned char iso[kMaxLayers + 1];   // round r has taps on the isolated planes
  int dil[kMaxLayers];
  // den_short_kernel only, tuning aid (SVDD_DEN_TRACE=<csv>, tools/den_trace.py): SM clock stamps of CTA 0's
  // third pair, [round][item][8 slots]
  unsigned long long* trace;
};

__host__ __device__ inline int smem_bytes(int a_rows, int c_rows = 0) {
  return 2 * (a_rows + c_rows) * 128 + kStages * kStageBytes + kParamBytes + kW2Bytes + kXchBytes + kBarBytes + 1024;
}

// tile m of the CTA's work item: first row in sequence coordinates
__host__ __device__ inline int tile_seq_row0(int m, int two_seq) { return two_seq ? 0 : 128 * m; }
// two sequences per CTA: tile 1's sequence sits in tile rows 64..127 (TMEM lane quadrants 2, 3),
// tile 0's in rows 0..63 (quadrants 0, 1), so the four epilogue warps that own real rows sit on
// four different warp schedulers; the other four only keep the barriers in step.
__host__ __device__ inline int tile_lane0(int m, int two_seq) { return two_seq ? 64 * m : 0; }
// first row of the operand planes (before pad_before) that tile m's MMA of tap offset o reads
__host__ __device__ inline int tile_plane_row(int m, int o, int two_seq) { return 128 * m + o - tile_lane0(m, two_seq); }
// does tap offset `o` touch real rows of tile m?
__host__ __device__ inline bool tap_hits(int m, int o, int L, int two_seq) {
  if (two_seq) return o > -L && o < L;
  const int lo = tile_seq_row0(m, two_seq) + o;
  return (tile_seq_row0(m, two_seq) < L) && (lo + 128 > 0) && (lo < L);
}

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
        "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// 32 floats from SHARED memory (explicit ld.shared: through a generic pointer ptxas emits LD.E
// plus two R2UR per load for the address-space descriptor)
__device__ __forceinline__ void ld_param32(const float* p, float* v) {
  const uint32_t a = ptx::smem_u32(p);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[4 * i]), "=f"(v[4 * i + 1]), "=f"(v[4 * i + 2]), "=f"(v[4 * i + 3])
                 : "r"(a + 16u * i));
}

// EW = epilogue warps.  8: thread = one row (or half a row, `split`).  16 (two sequences per CTA in
// combined mode only): thread = a QUARTER of a row (32 channels), four warps per TMEM lane quadrant
// -- the epilogue of a round is a latency chain (tensor-memory loads, LayerNorm statistics through
// shared memory, operand stores, barrier hops), and four warps per scheduler hide it better than
// two; <= 112 registers per thread.
template <typename Tok, int EW = 8>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
den_fused_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW0,
                 const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int plane_bytes = a.a_rows * 128;
  const int c_plane_bytes = a.cmb ? a.c_rows * 128 : 0;
  uint8_t* s_a = smem;                                   // [2 K halves][a_rows][64 bf16], 128B swizzle
  uint8_t* s_c = smem + 2 * plane_bytes;                 // combined planes [2 K halves][c_rows][64 bf16] (cmb mode)
  uint8_t* s_ring = s_c + 2 * c_plane_bytes;             // [kStages][128 x 64 bf16]
  float* s_param = reinterpret_cast<float*>(s_ring + kStages * kStageBytes);
  float* s_w2 = s_param + kParamBytes / 4;               // [5][128] + b2[5]
  float* s_xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_w2) + kW2Bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_xch) + kXchBytes);
  uint64_t* full_bar = bars;                  // [kStages]
  uint64_t* empty_bar = bars + kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;   // MMAs of a round retired
  uint64_t* aready_bar = bars + 2 * kStages + 1;   // operand of the next round written, accumulators free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L, nl = a.n_layers, two_seq = a.two_seq;
  const int64_t items = two_seq ? (a.n_rows + 1) / 2 : a.n_rows;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW);
    ptx::prefetch_tmap(&tmW0);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        ptx::mbar_init(&full_bar[i], 1);
        ptx::mbar_init(&empty_bar[i], 1);
      }
      ptx::mbar_init(tfull_bar, 1);
      ptx::mbar_init(aready_bar, EW);
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  // zero the operand planes once: pad rows and rows >= L are never written afterwards
  for (int i = threadIdx.x; i < 2 * (plane_bytes + c_plane_bytes) / 16; i += 64 + 32 * EW)
    reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < kVocab * kH; i += 64 + 32 * EW) s_w2[i] = a.fc2_w[i];
  if (threadIdx.x < kVocab) s_w2[kVocab * kH + threadIdx.x] = a.fc2_b[threadIdx.x];
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  // which tiles of work item `it` hold a sequence
  auto tile_live = [&](int64_t it, int m) -> bool {
    if (two_seq) return 2 * it + m < a.n_rows;
    return 128 * m < L;
  };

  if (warp == 0) {
    // ===================== weight producer =====================
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        const bool live1 = tile_live(it, 1);
        for (int r = 0; r <= nl; ++r) {
          const int taps = r < nl ? kTaps : 1;
          const int dil = r < nl ? a.dil[r] : 1;
          for (int t = 0; t < taps; ++t) {
            const int o = (t - taps / 2) * dil;
            if (!(tap_hits(0, o, L, two_seq) || (live1 && tap_hits(1, o, L, two_seq)))) continue;
            for (int kb = 0; kb < 2; ++kb) {
              ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
              ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
              if (r < nl)
                ptx::tma_load_2d(s_ring + stage * kStageBytes, &tmW, &full_bar[stage], kb * 64, (r * kTaps + t) * kH);
              else
                ptx::tma_load_2d(s_ring + stage * kStageBytes, &tmW0, &full_bar[stage], kb * 64, 0);
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // ONE elected thread runs the whole loop (as the producer does).  With an elect region per
    // weight stage the issuing warp spent ~60 instructions per stage on descriptor set-up, elect /
    // reconverge and loop control and was the critical path of den_short_kernel (see there); here
    // the descriptors are kept as their low words and advanced by adds.
    if (ptx::elect_one()) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, kH);
    constexpr uint32_t kStage16 = kStageBytes >> 4;
    uint32_t stage = 0, phase = 0;
    uint32_t round = 0;                      // aready phase counter
    const uint32_t a_base = ptx::smem_u32(s_a);
    const uint32_t a_lo = ptx::kmajor_sw128_desc_lo(a_base);
    const uint32_t ring_lo = ptx::kmajor_sw128_desc_lo(ptx::smem_u32(s_ring));
    const uint32_t plane16 = (uint32_t)plane_bytes >> 4;
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const bool live1 = tile_live(it, 1);
      for (int r = 0; r <= nl; ++r, ++round) {
        ptx::mbar_wait(aready_bar, round & 1);
        ptx::tc_fence_after();
        const int taps = r < nl ? kTaps : 1;
        const int dil = r < nl ? a.dil[r] : 1;
        if (a.ilv) {
          uint32_t st = 0u;
          const uint32_t aq = a_lo + (uint32_t)a.pad_before * 8u;   // one plane row = 8 descriptor units
          for (int t = 0; t < taps; ++t) {
            const int o = (t - taps / 2) * dil;
            if (!(o > -L && o < L)) continue;
            for (uint32_t kb = 0; kb < 2; ++kb) {
              ptx::mbar_wait(&full_bar[stage], phase);
              ptx::tc_fence_after();
              const uint32_t db = ring_lo + stage * kStage16;
              const uint32_t da = aq + (uint32_t)(16 * o) + kb * plane16;
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k)
                ptx::umma_bf16_lo(tmem_base, da + 2 * k, db + 2 * k, idesc, (st | k) != 0u);
              ptx::umma_commit(&empty_bar[stage]);
              st = 1u;
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        } else if (a.cmb) {
          // combined mode: acc0 (cols 0..127) <- taps on the combined planes, both sequences;
          // cols 256.. <- sequence A's isolated taps, cols 128.. <- sequence B's
          const uint32_t c_lo = ptx::kmajor_sw128_desc_lo(ptx::smem_u32(s_c));
          const uint32_t cplane16 = (uint32_t)c_plane_bytes >> 4;
          uint32_t st_c = 0u, st_a = 0u, st_b = 0u;
          for (int t = 0; t < taps; ++t) {
            const int o = (t - taps / 2) * dil;
            if (!(o > -L && o < L)) continue;
            const bool comb = (o < 0 ? -o : o) <= a.cmb_max;
            for (uint32_t kb = 0; kb < 2; ++kb) {
              ptx::mbar_wait(&full_bar[stage], phase);
              ptx::tc_fence_after();
              const uint32_t db = ring_lo + stage * kStage16;
              if (comb) {
                const uint32_t da = c_lo + kb * cplane16 + (uint32_t)((a.pad_c + o) * 8);
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                  ptx::umma_bf16_lo(tmem_base, da + 2 * k, db + 2 * k, idesc, (st_c | k) != 0u);
              } else {
                const uint32_t da0 = a_lo + kb * plane16 + (uint32_t)((a.pad_before + o) * 8);
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                  ptx::umma_bf16_lo(tmem_base + 2 * kH, da0 + 2 * k, db + 2 * k, idesc, (st_a | k) != 0u);
                if (live1) {
                  const uint32_t da1 = a_lo + kb * plane16 + (uint32_t)((a.pad_before + a.iso_b - 64 + o) * 8);
#pragma unroll
                  for (uint32_t k = 0; k < 4; ++k)
                    ptx::umma_bf16_lo(tmem_base + kH, da1 + 2 * k, db + 2 * k, idesc, (st_b | k) != 0u);
                }
              }
              ptx::umma_commit(&empty_bar[stage]);
              if (comb) st_c = 1u; else { st_a = 1u; st_b = 1u; }
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        } else {
          uint32_t started[2] = {0u, 0u};
          for (int t = 0; t < taps; ++t) {
            const int o = (t - taps / 2) * dil;
            const bool hit0 = tap_hits(0, o, L, two_seq);
            const bool hit1 = live1 && tap_hits(1, o, L, two_seq);
            if (!(hit0 || hit1)) continue;
            // 128 rows of a K half starting at tile_plane_row: the tap is a row offset
            const uint32_t da0 = a_lo + (uint32_t)((a.pad_before + tile_plane_row(0, o, two_seq)) * 8);
            const uint32_t da1 = a_lo + (uint32_t)((a.pad_before + tile_plane_row(1, o, two_seq)) * 8);
            for (uint32_t kb = 0; kb < 2; ++kb) {
              ptx::mbar_wait(&full_bar[stage], phase);
              ptx::tc_fence_after();
              const uint32_t db = ring_lo + stage * kStage16;
              if (hit0) {
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                  ptx::umma_bf16_lo(tmem_base, da0 + kb * plane16 + 2 * k, db + 2 * k, idesc, (started[0] | k) != 0u);
              }
              if (hit1) {
#pragma unroll
                for (uint32_t k = 0; k < 4; ++k)
                  ptx::umma_bf16_lo(tmem_base + kH, da1 + kb * plane16 + 2 * k, db + 2 * k, idesc, (started[1] | k) != 0u);
              }
              ptx::umma_commit(&empty_bar[stage]);
              started[0] |= hit0 ? 1u : 0u;
              started[1] |= hit1 ? 1u : 0u;
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
        ptx::umma_commit(tfull_bar);
      }
    }
    }
    __syncwarp();
  } else if constexpr (EW == 16) {
    // ===================== quad epilogue (combined mode, 16 warps) =====================
    // TMEM lane quadrant q = warp & 3 holds rows 32 (q & 1) .. +31 of sequence q >> 1; the four
    // warps that can reach a quadrant split the 128 channels in slices of 32.  LayerNorm statistics
    // of the four slices are combined through shared memory (Chan's update for four groups of 32)
    // behind one 128-thread named barrier per round.
    constexpr int kET = 32 * EW;
    const int ew = warp - 2;
    const int slice = ew >> 2;               // channels 32 * slice .. +31
    const int quad = warp & 3;
    const int m = quad >> 1;
    const int etid = threadIdx.x - 64;       // slice == etid >> 7
    const int row = (quad & 1) * 32 + lane;
    const int arow = a.pad_before + a.iso_b * m + row;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_acc = t_lane + slice * 32;
    const uint32_t t_iso = t_lane + (m == 0 ? 2 * kH : kH) + slice * 32;
    const uint32_t t_res = t_lane + 3 * kH + slice * 32;
    const int kh = slice >> 1, cj0 = (slice & 1) * 4;          // K half and first 16-byte chunk of this slice
    uint8_t* a_rowh = s_a + (size_t)kh * plane_bytes + (size_t)arow * 128;
    uint8_t* c_rowh = s_c + (size_t)kh * c_plane_bytes + (size_t)(a.pad_c + 64 * m + row) * 128;
    const int x7 = arow & 7;
    const bool wact = (row - lane) < L;
    const int ch0 = slice * 32;
    const Tok* tokens = reinterpret_cast<const Tok*>(a.tokens);
    float* s_lgx = s_xch + 2 * kET * 2;                        // logits exchange [3 slices][128 rows][5]
    uint32_t tphase = 0;
    int xpar = 0;
    auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kET) : "memory"); };
    auto grp_sync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(2 + quad) : "memory"); };

    auto write_operand = [&](float* v, const float* P, bool ln, bool valid, bool to_iso) {
      if (ln) {
        float pt[32];
        ld_param32(P + 1 * kH + ch0, pt);
        float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] += pt[i]; s4[i & 3] += v[i]; }
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        const float ml = sum * (1.0f / 32);
        float q4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 32; ++i) { const float d = v[i] - ml; q4[i & 3] += d * d; }
        const float m2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        float* mine = s_xch + ((size_t)xpar * kET + etid) * 2;
        *reinterpret_cast<float2*>(mine) = make_float2(sum, m2);
        grp_sync();
        float2 o[3];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          o[k - 1] = *reinterpret_cast<const float2*>(s_xch + ((size_t)xpar * kET + (etid ^ (k << 7))) * 2);
        xpar ^= 1;
        const float mean = ((sum + o[0].x) + (o[1].x + o[2].x)) * (1.0f / kH);
        float var = (m2 + o[0].y) + (o[1].y + o[2].y);
        {
          const float d0 = ml - mean, d1 = o[0].x * (1.0f / 32) - mean, d2 = o[1].x * (1.0f / 32) - mean,
                      d3 = o[2].x * (1.0f / 32) - mean;
          var += 32.0f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
        }
        const float rstd = rsqrtf(var * (1.0f / kH) + 1e-5f);
        float pg[32];
        ld_param32(P + 2 * kH + ch0, pg);
        ld_param32(P + 3 * kH + ch0, pt);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * pg[i] + pt[i];
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 u = make_uint4(pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                                     pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
          ptx::sts128(c_rowh + (((cj0 + j) ^ x7) << 4), u);
          if (to_iso) ptx::sts128(a_rowh + (((cj0 + j) ^ x7) << 4), u);
        }
      }
    };
    auto stage_params = [&](int r, float* P) {
      for (int i = etid; i < kH; i += kET) {
        P[i] = r < nl ? a.conv_b[r * kH + i] : a.fc0_b[i];
        if (r + 1 < nl) {
          P[1 * kH + i] = a.time_bias[(r + 1) * kH + i];
          P[2 * kH + i] = a.ln_g[(r + 1) * kH + i];
          P[3 * kH + i] = a.ln_b[(r + 1) * kH + i];
        }
      }
    };

    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t seq = 2 * it + m;
      const bool valid = (seq < a.n_rows) && (row < L);
      {
        float* P = s_param;
        for (int i = etid; i < kH; i += kET) {
          P[1 * kH + i] = a.time_bias[i];
          P[2 * kH + i] = a.ln_g[i];
          P[3 * kH + i] = a.ln_b[i];
        }
        epi_sync();
        if (wact) {
          float v[32];
          int tk[kTaps];
#pragma unroll
          for (int t = 0; t < kTaps; ++t) {
            const int li = row + t - kTaps / 2;
            tk[t] = (valid && li >= 0 && li < L) ? load_tok(tokens, (size_t)seq * L + li) : -1;
          }
          {
            const float4* b4 = reinterpret_cast<const float4*>(a.embed_b + ch0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 f = __ldg(b4 + i);
              v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
            }
#pragma unroll
            for (int t = 0; t < kTaps; ++t) {
              if (tk[t] >= 0) {
                const float4* w4 = reinterpret_cast<const float4*>(a.embed_w + (t * kVocab + tk[t]) * kH + ch0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 f = __ldg(w4 + i);
                  v[4 * i] += f.x; v[4 * i + 1] += f.y; v[4 * i + 2] += f.z; v[4 * i + 3] += f.w;
                }
              }
            }
            uint32_t raw[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = fmaxf(v[i], 0.0f);
              raw[i] = __float_as_uint(v[i]);
            }
            tmem_st_32x32(t_res, raw);
          }
          write_operand(v, P, true, valid, a.iso[0] != 0);
          tmem_st_wait();
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(aready_bar);
      }
      for (int r = 0; r < nl; ++r) {
        float* P = s_param + ((r + 1) & 1) * (4 * kH);
        stage_params(r, P);
        epi_sync();
        ptx::mbar_wait(tfull_bar, tphase);
        tphase ^= 1;
        ptx::tc_fence_after();
        if (wact) {
          float v[32];
          uint32_t racc[32], rres[32];
          ptx::tmem_ld_32x32(t_acc, racc);
          ptx::tmem_ld_32x32(t_res, rres);
          ptx::tmem_ld_wait();
          if (a.iso[r]) {
            uint32_t riso[32];
            ptx::tmem_ld_32x32(t_iso, riso);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) racc[i] = __float_as_uint(__uint_as_float(racc[i]) + __uint_as_float(riso[i]));
          }
          {
            float pb[32];
            ld_param32(P + ch0, pb);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float f = __uint_as_float(rres[i]) + fmaxf(__uint_as_float(racc[i]) + pb[i], 0.0f);
              v[i] = f;
              rres[i] = __float_as_uint(f);
            }
            tmem_st_32x32(t_res, rres);
          }
          write_operand(v, P, r + 1 < nl, valid, a.iso[r + 1] != 0);
          tmem_st_wait();
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(aready_bar);
      }
      {
        float* P = s_param + ((nl + 1) & 1) * (4 * kH);
        stage_params(nl, P);
        epi_sync();
        ptx::mbar_wait(tfull_bar, tphase);
        tphase ^= 1;
        ptx::tc_fence_after();
        if (wact) {
          float lg[kVocab];
#pragma unroll
          for (int j = 0; j < kVocab; ++j) lg[j] = 0.0f;
          {
            uint32_t racc[32];
            ptx::tmem_ld_32x32(t_acc, racc);
            ptx::tmem_ld_wait();
            float pb[32];
            ld_param32(P + ch0, pb);
            float y[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(__uint_as_float(racc[i]) + pb[i], 0.0f);
#pragma unroll
            for (int j = 0; j < kVocab; ++j) {
              float w[32];
              ld_param32(s_w2 + j * kH + ch0, w);
#pragma unroll
              for (int i = 0; i < 32; ++i) lg[j] += y[i] * w[i];
            }
          }
          const int qrow = quad * 32 + lane;
          if (slice > 0) {
            float* x = s_lgx + ((size_t)(slice - 1) * 128 + qrow) * kVocab;
#pragma unroll
            for (int j = 0; j < kVocab; ++j) x[j] = lg[j];
          }
          grp_sync();
          if (slice == 0 && valid) {
            float* o = a.logits + ((size_t)seq * L + row) * kVocab;
#pragma unroll
            for (int j = 0; j < kVocab; ++j)
              o[j] = ((lg[j] + s_lgx[(size_t)qrow * kVocab + j]) +
                      (s_lgx[(size_t)(128 + qrow) * kVocab + j] + s_lgx[(size_t)(256 + qrow) * kVocab + j])) + s_w2[kVocab * kH + j];
          }
          grp_sync();                        // the exchange area is reused by the next item
        }
        ptx::tc_fence_before();
      }
    }
  } else if (!a.split) {
    // ===================== epilogue: thread = one row of one tile =====================
    const int ew = warp - 2;
    const int quad = warp & 3;               // TMEM lane quadrant of this warp
    const int m = two_seq ? (quad >> 1) : (ew >> 2);             // tile
    const int etid = threadIdx.x - 64;
    const int r_tile = quad * 32 + lane;
    const int row = tile_seq_row0(m, two_seq) + r_tile - tile_lane0(m, two_seq);   // position within the sequence
    const int arow = a.pad_before + 128 * m + r_tile - tile_lane0(m, two_seq);     // row of the operand planes
    // does this warp own any real row?  (two_seq: warps 6..9 duplicate the quadrants of 2..5)
    const bool wact = (two_seq ? ew < 4 : true) && (row - lane < L);
    const uint32_t t_acc = tmem_base + ((uint32_t)(quad * 32) << 16) + m * kH;
    const uint32_t t_res = t_acc + 2 * kH;
    uint8_t* a_row0 = s_a + (size_t)arow * 128;
    uint8_t* a_row1 = a_row0 + plane_bytes;
    const int x7 = arow & 7;
    const Tok* tokens = reinterpret_cast<const Tok*>(a.tokens);
    uint32_t tphase = 0;

    // v = feat (fp32 row).  Writes LN(v + tbias)*gamma + beta (or plain v) as bf16 into the
    // operand planes; P = [bias, tbias, gamma, beta] slots of the NEXT round.  Sums run as four
    // independent chains (the thread has only one other warp to hide latency behind).
    auto write_operand = [&](float* v, const float* P, bool ln, bool valid) {
      if (ln) {
        float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float pt[32];
          ld_param32(P + 1 * kH + c * 32, pt);
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[c * 32 + i] += pt[i]; s4[i & 3] += v[c * 32 + i]; }
        }
        const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / kH);
        float q4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < kH; ++i) { const float d = v[i] - mean; q4[i & 3] += d * d; }
        const float rstd = rsqrtf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * (1.0f / kH) + 1e-5f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float pg[32], pb[32];
          ld_param32(P + 2 * kH + c * 32, pg);
          ld_param32(P + 3 * kH + c * 32, pb);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c * 32 + i] = (v[c * 32 + i] - mean) * rstd * pg[i] + pb[i];
        }
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {       // 16-byte chunk j of the 256-byte row
          uint8_t* base = (j < 8) ? a_row0 : a_row1;
          ptx::sts128(base + (((j & 7) ^ x7) << 4),
                      make_uint4(pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                                 pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7])));
        }
      }
    };
    // slots of round r: [bias_r, tbias_{r+1}, gamma_{r+1}, beta_{r+1}]
    auto stage_params = [&](int r, float* P) {
      for (int i = etid; i < kH; i += kEpiThreads) {
        P[i] = r < nl ? a.conv_b[r * kH + i] : a.fc0_b[i];
        if (r + 1 < nl) {
          P[1 * kH + i] = a.time_bias[(r + 1) * kH + i];
          P[2 * kH + i] = a.ln_g[(r + 1) * kH + i];
          P[3 * kH + i] = a.ln_b[(r + 1) * kH + i];
        }
      }
    };

    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t seq = two_seq ? 2 * it + m : it;
      const bool valid = (seq < a.n_rows) && (row < L);
      // ---- embed: Conv(5 -> 128, k9) as a weight gather, ReLU, LayerNorm_0 ----------------
      {
        float* P = s_param;                  // slots 1..3 <- layer 0's norm
        for (int i = etid; i < kH; i += kEpiThreads) {
          P[1 * kH + i] = a.time_bias[i];
          P[2 * kH + i] = a.ln_g[i];
          P[3 * kH + i] = a.ln_b[i];
        }
        epi_bar_sync();
        if (wact) {
        float v[kH];
        int tk[kTaps];
#pragma unroll
        for (int t = 0; t < kTaps; ++t) {
          const int li = row + t - kTaps / 2;
          tk[t] = (valid && li >= 0 && li < L) ? load_tok(tokens, (size_t)seq * L + li) : -1;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float acc[32];
          const float4* b4 = reinterpret_cast<const float4*>(a.embed_b + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 f = __ldg(b4 + i);
            acc[4 * i] = f.x; acc[4 * i + 1] = f.y; acc[4 * i + 2] = f.z; acc[4 * i + 3] = f.w;
          }
#pragma unroll
          for (int t = 0; t < kTaps; ++t) {
            if (tk[t] >= 0) {
              const float4* w4 = reinterpret_cast<const float4*>(a.embed_w + (t * kVocab + tk[t]) * kH + c * 32);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 f = __ldg(w4 + i);
                acc[4 * i] += f.x; acc[4 * i + 1] += f.y; acc[4 * i + 2] += f.z; acc[4 * i + 3] += f.w;
              }
            }
          }
          uint32_t raw[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            v[c * 32 + i] = fmaxf(acc[i], 0.0f);
            raw[i] = __float_as_uint(v[c * 32 + i]);
          }
          tmem_st_32x32(t_res + c * 32, raw);          // residual stream -> tensor memory
        }
        write_operand(v, P, true, valid);
        tmem_st_wait();
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(aready_bar);
      }
      // ---- conv layers ---------------------------------------------------------------------
      for (int r = 0; r < nl; ++r) {
        float* P = s_param + ((r + 1) & 1) * (4 * kH);
        stage_params(r, P);
        epi_bar_sync();
        ptx::mbar_wait(tfull_bar, tphase);
        tphase ^= 1;
        ptx::tc_fence_after();
        if (wact) {
        float v[kH];
        uint32_t racc[2][32], rres[2][32];
        ptx::tmem_ld_32x32(t_acc, racc[0]);
        ptx::tmem_ld_32x32(t_res, rres[0]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          ptx::tmem_ld_wait();
          if (c + 1 < 4) {                   // next chunk's loads fly while this one is computed
            ptx::tmem_ld_32x32(t_acc + (c + 1) * 32, racc[(c + 1) & 1]);
            ptx::tmem_ld_32x32(t_res + (c + 1) * 32, rres[(c + 1) & 1]);
          }
          float pb[32];
          ld_param32(P + c * 32, pb);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            // feat += relu(conv + bias)   (models/dnaconv.py:196-200)
            const float f = __uint_as_float(rres[c & 1][i]) + fmaxf(__uint_as_float(racc[c & 1][i]) + pb[i], 0.0f);
            v[c * 32 + i] = f;
            rres[c & 1][i] = __float_as_uint(f);
          }
          tmem_st_32x32(t_res + c * 32, rres[c & 1]);
        }
        write_operand(v, P, r + 1 < nl, valid);
        tmem_st_wait();
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(aready_bar);
      }
      // ---- final_conv: ReLU(1x1) then 1x1 to the 5 logits (models/dnaconv.py:163-165,201) ----
      {
        float* P = s_param + ((nl + 1) & 1) * (4 * kH);
        stage_params(nl, P);
        epi_bar_sync();
        ptx::mbar_wait(tfull_bar, tphase);
        tphase ^= 1;
        ptx::tc_fence_after();
        if (wact) {
        float lg[kVocab];
#pragma unroll
        for (int j = 0; j < kVocab; ++j) lg[j] = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t racc[32];
          ptx::tmem_ld_32x32(t_acc + c * 32, racc);
          ptx::tmem_ld_wait();
          float pb[32];
          ld_param32(P + c * 32, pb);
          float y[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) y[i] = fmaxf(__uint_as_float(racc[i]) + pb[i], 0.0f);
#pragma unroll
          for (int j = 0; j < kVocab; ++j) {
            float w[32];
            ld_param32(s_w2 + j * kH + c * 32, w);
#pragma unroll
            for (int i = 0; i < 32; ++i) lg[j] += y[i] * w[i];
          }
        }
        if (valid) {
          float* o = a.logits + ((size_t)seq * L + row) * kVocab;
#pragma unroll
          for (int j = 0; j < kVocab; ++j) o[j] = lg[j] + s_w2[kVocab * kH + j];
        }
        }
        // the next item's embed overwrites the operand planes and the residual columns: the
        // MMAs that read them have retired (tfull), the accumulator reads above are complete
        ptx::tc_fence_before();
      }
    }
  } else {
    // ===================== split epilogue (two sequences per CTA) =====================
    // Real rows occupy TMEM lane quadrants 0,1 (tile 0) and 2,3 (tile 1); each quadrant is
    // reachable by two of the eight epilogue warps (warp & 3).  Both work on the same 32 rows:
    // warps 2..5 on channels 0..63, warps 6..9 on channels 64..127, so a row's serial chain
    // (TMEM loads, LayerNorm, operand write) is half as long.  The LayerNorm statistics are
    // combined through shared memory (Chan's pairwise update of sum and centred sum of squares)
    // with one 64-thread named barrier per layer.
    const int ew = warp - 2;
    const int half = ew >> 2;
    const int quad = warp & 3;
    const int etid = threadIdx.x - 64;
    const bool ilv = a.ilv != 0;
    const bool cmb = a.cmb != 0;
    const int rq = quad * 32 + lane;                             // tile row = TMEM lane
    const int m = ilv ? (rq & 1) : (quad >> 1);                  // sequence of the item
    const int row = ilv ? (rq >> 1) : ((quad & 1) * 32 + lane);  // position within the sequence
    // plane row: interleaved = pad + tile row; otherwise the isolated planes (pad_before, iso_b: multiples of 8)
    const int arow = ilv ? a.pad_before + rq : a.pad_before + a.iso_b * m + row;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_acc = t_lane + ((cmb || ilv) ? 0 : m * kH) + half * 64;
    const uint32_t t_iso = t_lane + (m == 0 ? 2 * kH : kH) + half * 64;       // cmb: this sequence's isolated taps
    const uint32_t t_res = ilv ? t_lane + kH + half * 64                      // ilv: accumulator | residual
                               : (cmb ? t_lane + 3 * kH + half * 64 : t_acc + 2 * kH);   // cmb: one block, lanes 0..63 / 64..127
    uint8_t* a_rowh = s_a + (size_t)half * plane_bytes + (size_t)arow * 128;
    uint8_t* c_rowh = s_c + (size_t)half * c_plane_bytes + (size_t)(a.pad_c + 64 * m + row) * 128;   // combined planes
    const int x7 = arow & 7;                                     // == tile row & 7 on every kind of plane
    const bool wact = ilv ? (quad * 32 < 2 * L) : ((row - lane) < L);
    const int ch0 = half * 64;
    const Tok* tokens = reinterpret_cast<const Tok*>(a.tokens);
    uint32_t tphase = 0;
    int xpar = 0;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory"); };
    auto xch_mine = [&](int par) { return s_xch + ((size_t)par * kEpiThreads + etid) * 8; };
    auto xch_peer = [&](int par) { return s_xch + ((size_t)par * kEpiThreads + (etid ^ 128)) * 8; };

    auto write_operand = [&](float* v, const float* P, bool ln, bool valid, bool to_iso, bool to_cmb) {
      if (ln) {
        float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float pt[32];
          ld_param32(P + 1 * kH + ch0 + c * 32, pt);
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[c * 32 + i] += pt[i]; s4[i & 3] += v[c * 32 + i]; }
        }
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        const float ml = sum * (1.0f / 64);
        float q4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 64; ++i) { const float d = v[i] - ml; q4[i & 3] += d * d; }
        const float m2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        *reinterpret_cast<float2*>(xch_mine(xpar)) = make_float2(sum, m2);
        pair_sync();
        const float2 o = *reinterpret_cast<const float2*>(xch_peer(xpar));
        xpar ^= 1;
        const float mean = (sum + o.x) * (1.0f / kH);
        const float dm = (sum - o.x) * (1.0f / 64);
        const float rstd = rsqrtf((m2 + o.y + dm * dm * 32.0f) * (1.0f / kH) + 1e-5f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float pg[32], pb[32];
          ld_param32(P + 2 * kH + ch0 + c * 32, pg);
          ld_param32(P + 3 * kH + ch0 + c * 32, pb);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c * 32 + i] = (v[c * 32 + i] - mean) * rstd * pg[i] + pb[i];
        }
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {        // 16-byte chunk j of this K half's 128-byte row
          const uint4 u = make_uint4(pack2(v[8 * j], v[8 * j + 1]), pack2(v[8 * j + 2], v[8 * j + 3]),
                                     pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
          if (to_iso) ptx::sts128(a_rowh + ((j ^ x7) << 4), u);
          if (to_cmb) ptx::sts128(c_rowh + ((j ^ x7) << 4), u);
        }
      }
    };
    auto stage_params = [&](int r, float* P) {
      for (int i = etid; i < kH; i += kEpiThreads) {
        P[i] = r < nl ? a.conv_b[r * kH + i] : a.fc0_b[i];
        if (r + 1 < nl) {
          P[1 * kH + i] = a.time_bias[(r + 1) * kH + i];
          P[2 * kH + i] = a.ln_g[(r + 1) * kH + i];
          P[3 * kH + i] = a.ln_b[(r + 1) * kH + i];
        }
      }
    };

    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      const int64_t seq = 2 * it + m;
      const bool valid = (seq < a.n_rows) && (row < L);
      {
        float* P = s_param;
        for (int i = etid; i < kH; i += kEpiThreads) {
          P[1 * kH + i] = a.time_bias[i];
          P[2 * kH + i] = a.ln_g[i];
          P[3 * kH + i] = a.ln_b[i];
        }
        epi_bar_sync();
        if (wact) {
          float v[64];
          int tk[kTaps];
#pragma unroll
          for (int t = 0; t < kTaps; ++t) {
            const int li = row + t - kTaps / 2;
            tk[t] = (valid && li >= 0 && li < L) ? load_tok(tokens, (size_t)seq * L + li) : -1;
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float acc[32];
            const float4* b4 = reinterpret_cast<const float4*>(a.embed_b + ch0 + c * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 f = __ldg(b4 + i);
              acc[4 * i] = f.x; acc[4 * i + 1] = f.y; acc[4 * i + 2] = f.z; acc[4 * i + 3] = f.w;
            }
#pragma unroll
            for (int t = 0; t < kTaps; ++t) {
              if (tk[t] >= 0) {
                const float4* w4 = reinterpret_cast<const float4*>(a.embed_w + (t * kVocab + tk[t]) * kH + ch0 + c * 32);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 f = __ldg(w4 + i);
                  acc[4 * i] += f.x; acc[4 * i + 1] += f.y; acc[4 * i + 2] += f.z; acc[4 * i + 3] += f.w;
                }
              }
            }
            uint32_t raw[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[c * 32 + i] = fmaxf(acc[i], 0.0f);
              raw[i] = __float_as_uint(v[c * 32 + i]);
            }
            tmem_st_32x32(t_res + c * 32, raw);
          }
          write_operand(v, P, true, valid, !cmb || a.iso[0] != 0, cmb);
          tmem_st_wait();
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(aready_bar);
      }
      for (int r = 0; r < nl; ++r) {
        float* P = s_param + ((r + 1) & 1) * (4 * kH);
        stage_params(r, P);
        epi_bar_sync();
        ptx::mbar_wait(tfull_bar, tphase);
        tphase ^= 1;
        ptx::tc_fence_after();
        if (wact) {
          float v[64];
          uint32_t racc[2][32], rres[2][32];
          ptx::tmem_ld_32x32(t_acc, racc[0]);
          ptx::tmem_ld_32x32(t_res, rres[0]);
          ptx::tmem_ld_32x32(t_acc + 32, racc[1]);
          ptx::tmem_ld_32x32(t_res + 32, rres[1]);
          ptx::tmem_ld_wait();
          if (cmb && a.iso[r]) {             // this round also ran taps on the isolated planes
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t riso[32];
              ptx::tmem_ld_32x32(t_iso + 32 * c, riso);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i)
                racc[c][i] = __float_as_uint(__uint_as_float(racc[c][i]) + __uint_as_float(riso[i]));
            }
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float pb[32];
            ld_param32(P + ch0 + c * 32, pb);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float f = __uint_as_float(rres[c][i]) + fmaxf(__uint_as_float(racc[c][i]) + pb[i], 0.0f);
              v[c * 32 + i] = f;
              rres[c][i] = __float_as_uint(f);
            }
            tmem_st_32x32(t_res + c * 32, rres[c]);
          }
          write_operand(v, P, r + 1 < nl, valid, !cmb || a.iso[r + 1] != 0, cmb);
          tmem_st_wait();
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(aready_bar);
      }
      {
        float* P = s_param + ((nl + 1) & 1) * (4 * kH);
        stage_params(nl, P);
        epi_bar_sync();
        ptx::mbar_wait(tfull_bar, tphase);
        tphase ^= 1;
        ptx::tc_fence_after();
        if (wact) {
          float lg[kVocab];
#pragma unroll
          for (int j = 0; j < kVocab; ++j) lg[j] = 0.0f;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t racc[32];
            ptx::tmem_ld_32x32(t_acc + c * 32, racc);
            ptx::tmem_ld_wait();
            float pb[32];
            ld_param32(P + ch0 + c * 32, pb);
            float y[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) y[i] = fmaxf(__uint_as_float(racc[i]) + pb[i], 0.0f);
#pragma unroll
            for (int j = 0; j < kVocab; ++j) {
              float w[32];
              ld_param32(s_w2 + j * kH + ch0 + c * 32, w);
#pragma unroll
              for (int i = 0; i < 32; ++i) lg[j] += y[i] * w[i];
            }
          }
          if (half == 1) {
            float* x = xch_mine(xpar);
#pragma unroll
            for (int j = 0; j < kVocab; ++j) x[j] = lg[j];
          }
          pair_sync();
          if (half == 0 && valid) {
            const float* x = xch_peer(xpar);
            float* o = a.logits + ((size_t)seq * L + row) * kVocab;
#pragma unroll
            for (int j = 0; j < kVocab; ++j) o[j] = (lg[j] + x[j]) + s_w2[kVocab * kH + j];
          }
          xpar ^= 1;
        }
        ptx::tc_fence_before();
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

}  // namespace denf
}  // namespace svdd
