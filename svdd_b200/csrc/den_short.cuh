// Stage 1 for SHORT sequences (L <= 64: RNA, L = 50) -- the whole MDLM denoiser in one persistent
// kernel, FOUR sequences per CTA as two items of two interleaved sequences that take turns.
//
// den_fused.cuh keeps a sequence's activations on the SM for the whole network but runs each
// round as a ping-pong: the MMAs of a round (operands from shared memory: the tensor pipe is paced
// by the SM's 128 B/clk of operand bandwidth) and its epilogue (bias, ReLU, residual, LayerNorm, bf16
// operand for the next layer) cannot overlap, because every tap reads rows the epilogue writes.
// ncu on the L = 50 pass: 46 % of all warp-stall samples are the eight epilogue warps waiting for the
// round's MMAs, the MMA warp waits for the epilogue a third of its time, tensor pipe 35 % active.
//
// With the INTERLEAVED layout (plane row 2p / 2p+1 = position p of sequence A / B; a tap offset o is
// a row offset of 2 o, so one 128-row MMA per tap serves both sequences for every tap) an item needs
// ONE pair of operand planes, one accumulator and one residual block: 256 of the 512 TMEM columns
// and ~70 KB of shared memory.  Two items therefore fit, and they alternate: while the epilogue
// warps work on item X's round the tensor core runs item Y's, and vice versa.  The two items share
// the zero rows between them (the conv padding of both), the per-layer parameters and the weight
// stream (each round's taps are fetched once per item, in MMA order).
//
// Roles (320 threads): warp 0 weight producer (TMA), warp 1 MMA issuer, warps 2-9 epilogue
// (TMEM lane quadrant = warp & 3; two threads per row, 64 channels each, LayerNorm statistics
// combined pairwise through shared memory as in den_fused's split epilogue).
#pragma once
#include "conv_gemm2.cuh"
#include "den_fused.cuh"

namespace svdd {
namespace dens {

using denf::kH;
using denf::kTaps;
constexpr int kStageBytes = denf::kStageBytes;   // one (tap, K half) weight tile: 128 x 64 bf16
constexpr int kRingBytes = 4 * kStageBytes;      // per CTA: 4 whole tiles (CG = 1) or 8 half tiles (CG = 2)
constexpr int kEpiThreads = 256;
constexpr int kThreads = 64 + kEpiThreads;
constexpr int kParamBytes = denf::kParamBytes;       // [parity][bias, tbias, gamma, beta][128]
constexpr int kW2Bytes = denf::kW2Bytes;
constexpr int kStatBytes = 2 * kEpiThreads * 2 * 4;  // LayerNorm exchange [parity][thread][sum, m2]
constexpr int kLgxBytes = kEpiThreads * 8 * 4;       // logits exchange [thread][5 (+3)]
constexpr int kBarBytes = 256;

// rows of one K-half plane: pad | item 0 (128) | pad | item 1 (128) | pad
__host__ __device__ inline int plane_rows(int pad) { return 3 * pad + 256; }
__host__ __device__ inline int smem_bytes(int pad) {
  return 2 * plane_rows(pad) * 128 + kRingBytes + kParamBytes + kW2Bytes + kStatBytes + kLgxBytes + kBarBytes + 1024;
}

// ---- helpers for the CTA-pair variant -------------------------------------------------------------
// "this warp's rows of the item's next operand are written": on the own CTA's barrier, or (CG = 2)
// on the pair leader's, with release semantics at cluster scope -- the stores (made visible to the
// async proxy by the caller's fence.proxy.async) are read by THIS SM's tensor core on behalf of an
// MMA the leader issues after observing the barrier.
template <int CG>
__device__ __forceinline__ void arrive_operand_ready(uint64_t* bar) {
  if constexpr (CG == 1) {
    ptx::mbar_arrive(bar);
  } else {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];"
                 ::"r"(ptx::smem_u32(bar) & gemm2::kPeerMask) : "memory");
  }
}
// the leader's wait on that barrier: acquire at cluster scope
template <int CG>
__device__ __forceinline__ void mbar_wait_cg(uint64_t* bar, uint32_t parity) {
  if constexpr (CG == 1) {
    ptx::mbar_wait(bar, parity);
  } else {
    uint32_t spins = 0, ok = 0;
    while (true) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
      if (ok) break;
      __nanosleep(16);
      if (++spins == (1u << 24)) {
        printf("svdd_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
        __trap();
      }
    }
  }
}

// CG = 2 (SVDD_DEN_CG=2, an experiment kept with its cross-check): the kernel runs as CTA PAIRS (cluster of
// two) and every MMA is a cta_group::2 instruction (M = 256) that covers the same item slot of both
// CTAs: each CTA stages only HALF of every weight tile (64 of the 128 output channels; ring of 8 half
// tiles), the leader CTA's issuing thread drives both, the peer's epilogue warps arrive on the leader's
// `aready` barriers, `tfull` / ring `empty` commits are multicast.  Measured (tools/den_trace.py,
// profiles/r02_den_trace_cg{1,2}.txt): the MMA phase of a 9-tap item and round takes 3.5 us instead of
// 3.7 us and the pass is within +-2.5 % of CG = 1 -- each SM's tensor core still reads the WHOLE weight
// tile (its half locally, the other half from the peer's shared memory), so only the TMA write of
// the ring is halved; see the note on shared-memory bandwidth at the MMA issuer below.
template <typename Tok, int CG>
__global__ void __launch_bounds__(kThreads, 1)
den_short_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmW0,
                 const __grid_constant__ denf::Args a) {
  constexpr int kStages = 4 * CG;                 // ring stages
  constexpr int kStB = kStageBytes / CG;          // bytes of one stage in this CTA
  const int rank = CG == 2 ? (int)gemm2::cluster_ctarank() : 0;
  const int64_t group = blockIdx.x / CG, n_groups = gridDim.x / CG;     // CTA pairs (CG = 2) or CTAs
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int pad = a.pad_before;
  const int plane_bytes = plane_rows(pad) * 128;
  uint8_t* s_a = smem;                                   // [2 K halves][plane_rows][64 bf16], 128B swizzle
  uint8_t* s_ring = s_a + 2 * plane_bytes;               // [kStages][128 / CG x 64 bf16]
  float* s_param = reinterpret_cast<float*>(s_ring + kRingBytes);
  float* s_w2 = s_param + kParamBytes / 4;               // [5][128] + b2[5]
  float* s_stat = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(s_w2) + kW2Bytes);
  float* s_lgx = s_stat + kStatBytes / 4;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_lgx) + kLgxBytes);
  uint64_t* full_bar = bars;                  // [kStages]
  uint64_t* empty_bar = bars + kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;   // [2 items] MMAs of the item's round retired
  uint64_t* aready_bar = bars + 2 * kStages + 2;   // [2 items] operand of the item's next round written
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L, nl = a.n_layers;
  const int64_t items = (a.n_rows + 1) / 2;              // two sequences per item
  const int64_t pairs = (items + 1) / 2;                 // two items per pass of a CTA
  const int64_t passes = (pairs + CG - 1) / CG;          // a CTA pair takes two pairs per pass; a CTA beyond the last pair runs on invalid rows

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
      for (int q = 0; q < 2; ++q) {
        ptx::mbar_init(&tfull_bar[q], 1);
        ptx::mbar_init(&aready_bar[q], 8 * CG);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    gemm2::tmem_alloc_cg<CG>(tmem_slot, 512);
  }
  for (int i = threadIdx.x; i < 2 * plane_bytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(s_a)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < kVocab * kH; i += kThreads) s_w2[i] = a.fc2_w[i];
  if (threadIdx.x < kVocab) s_w2[kVocab * kH + threadIdx.x] = a.fc2_b[threadIdx.x];
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  if constexpr (CG == 2) gemm2::cluster_sync_all(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  auto tap_used = [&](int o) { return o > -L && o < L; };
  // tuning aid: SM clock at the hand-over points of one pair (the CTA's third) of CTA 0
  auto stamp = [&](int64_t p, int r, int q, int slot) {
    if (a.trace != nullptr && blockIdx.x == 0 && p == 2 * CG * n_groups)
      a.trace[(size_t)(r * 2 + q) * 8 + slot] = (unsigned long long)clock64();
  };

  if (warp == 0) {
    // ===================== weight producer: the taps of (item 0, round), (item 1, round), ... =====================
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      for (int64_t pass = group; pass < passes; pass += n_groups) {
        const int64_t p = pass * CG + rank;
        const bool live1 = CG == 2 || 2 * p + 1 < items;
        for (int r = 0; r <= nl; ++r) {
          const int taps = r < nl ? kTaps : 1;
          const int dil = r < nl ? a.dil[r] : 1;
          for (int q = 0; q < 2; ++q) {
            if (q == 1 && !live1) continue;
            for (int t = 0; t < taps; ++t) {
              if (!tap_used((t - taps / 2) * dil)) continue;
              for (int kb = 0; kb < 2; ++kb) {
                // CG = 2: this CTA's half of the tile (output channels rank * 64 ..), both halves
                // complete the LEADER's full barrier
                ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                if (rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                if (r < nl)
                  gemm2::tma_load_2d_cg<CG>(s_ring + stage * kStB, &tmW, &full_bar[stage], kb * 64,
                                            (r * kTaps + t) * kH + rank * (kH / CG));
                else
                  gemm2::tma_load_2d_cg<CG>(s_ring + stage * kStB, &tmW0, &full_bar[stage], kb * 64, rank * (kH / CG));
                if (++stage == kStages) { stage = 0; phase ^= 1; }
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // ONE elected thread runs the whole loop (as the producer does): with an elect region per
    // weight stage the issuing warp needed ~60 instructions for every 4 MMAs (descriptor set-up,
    // elect / reconverge, loop control: ~575 cycles per stage against 256 tensor cycles; ncu source
    // page, profiles/r02_ncu_full_den_short_source_before.csv) and was the kernel's critical path.
    // Descriptors are kept as their low words and advanced by adds; a tap always takes two ring
    // stages (kStages is even), so both are handled in one iteration.
    //
    // What paces the MMAs (tools/den_trace.py): shared memory, not the tensor pipe and not the weight
    // stream.  A 128 x 128 x 16 SS-mode MMA reads 8 KB of operands in its 64 cycles, the SM's whole
    // 128 B/clk, and TMA writes another 16 KB into the ring per stage: 4 x 8 + 16 KB = 384 cycles per
    // stage against 256 tensor cycles, i.e. 3.6 us per 9-tap item and round -- what the trace shows,
    // independent of the ring depth (4 / 5 stages), of the number of CTAs running (37 / 74 / 148) and
    // of the polling (spinning instead of sleeping polls: slower).
    static_assert(kStages % 2 == 0, "a tap uses two consecutive ring stages");
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(128 * CG, kH);
      constexpr uint32_t kStage16 = kStB >> 4;
      const uint32_t ring_lo = ptx::kmajor_sw128_desc_lo(ptx::smem_u32(s_ring));
      const uint32_t a_lo = ptx::kmajor_sw128_desc_lo(ptx::smem_u32(s_a));
      const uint32_t plane16 = (uint32_t)plane_bytes >> 4;
      uint32_t stage = 0, phase = 0;
      uint32_t rounds[2] = {0u, 0u};           // aready completions consumed per item slot
      for (int64_t pass = group; pass < passes; pass += n_groups) {
        const int64_t p = pass * CG;
        const bool live1 = CG == 2 || 2 * p + 1 < items;
        for (int r = 0; r <= nl; ++r) {
          const int taps = r < nl ? kTaps : 1;
          const int dil = r < nl ? a.dil[r] : 1;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (q == 1 && !live1) continue;
            mbar_wait_cg<CG>(&aready_bar[q], rounds[q] & 1);
            ++rounds[q];
            ptx::tc_fence_after();
            stamp(p, r, q, 0);
            const uint32_t tm = tmem_base + q * 2 * kH;
            // one plane row = 128 bytes = 8 descriptor units; tap offset o = 2 o rows
            const uint32_t aq = a_lo + (uint32_t)(pad + q * (128 + pad)) * 8u;
            uint32_t st = 0u;
            for (int t = 0; t < taps; ++t) {
              const int o = (t - taps / 2) * dil;
              if (!tap_used(o)) continue;
              const uint32_t da = aq + (uint32_t)(16 * o);
              const uint32_t db = ring_lo + stage * kStage16;
              ptx::mbar_wait(&full_bar[stage], phase);
              ptx::tc_fence_after();
              if (st == 0u) stamp(p, r, q, 1);
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k)
                gemm2::umma_bf16_lo_cg<CG>(tm, da + 2 * k, db + 2 * k, idesc, (st | k) != 0u);
              gemm2::umma_commit_cg<CG>(&empty_bar[stage]);
              ptx::mbar_wait(&full_bar[stage + 1], phase);
              ptx::tc_fence_after();
#pragma unroll
              for (uint32_t k = 0; k < 4; ++k)
                gemm2::umma_bf16_lo_cg<CG>(tm, da + plane16 + 2 * k, db + kStage16 + 2 * k, idesc, 1u);
              gemm2::umma_commit_cg<CG>(&empty_bar[stage + 1]);
              st = 1u;
              stage += 2;
              if (stage == kStages) { stage = 0; phase ^= 1; }
            }
            gemm2::umma_commit_cg<CG>(&tfull_bar[q]);
            stamp(p, r, q, 2);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: two threads per row (64 channels each), items in turn =====================
    const int ew = warp - 2;
    const int half = ew >> 2;
    const int quad = warp & 3;
    const int etid = threadIdx.x - 64;
    const int rq = quad * 32 + lane;                // tile row = TMEM lane
    const int m = rq & 1;                           // sequence of the item
    const int pos = rq >> 1;                        // position within the sequence
    const int x7 = rq & 7;                          // pad is a multiple of 8
    const bool wact = quad * 32 < 2 * L;            // does this warp own real rows?
    const int ch0 = half * 64;
    const Tok* tokens = reinterpret_cast<const Tok*>(a.tokens);
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16) + half * 64;
    uint32_t tphase[2] = {0u, 0u};
    int xpar = 0;
    auto epi_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); };
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory"); };

    // v = feat (this thread's 64 channels).  Writes LN(v + tbias) * gamma + beta (or plain v) as bf16
    // into the item's operand plane; P = [bias, tbias, gamma, beta] slots of the NEXT round.
    auto write_operand = [&](float* v, const float* P, bool ln, bool valid, uint8_t* a_rowh) {
      if (ln) {
        float s4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float pt[32];
          denf::ld_param32(P + 1 * kH + ch0 + c * 32, pt);
#pragma unroll
          for (int i = 0; i < 32; ++i) { v[c * 32 + i] += pt[i]; s4[i & 3] += v[c * 32 + i]; }
        }
        const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        const float ml = sum * (1.0f / 64);
        float q4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int i = 0; i < 64; ++i) { const float d = v[i] - ml; q4[i & 3] += d * d; }
        const float m2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
        *reinterpret_cast<float2*>(s_stat + ((size_t)xpar * kEpiThreads + etid) * 2) = make_float2(sum, m2);
        pair_sync();
        const float2 o = *reinterpret_cast<const float2*>(s_stat + ((size_t)xpar * kEpiThreads + (etid ^ 128)) * 2);
        xpar ^= 1;
        const float mean = (sum + o.x) * (1.0f / kH);
        const float dm = (sum - o.x) * (1.0f / 64);
        const float rstd = rsqrtf((m2 + o.y + dm * dm * 32.0f) * (1.0f / kH) + 1e-5f);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float pg[32], pb[32];
          denf::ld_param32(P + 2 * kH + ch0 + c * 32, pg);
          denf::ld_param32(P + 3 * kH + ch0 + c * 32, pb);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[c * 32 + i] = (v[c * 32 + i] - mean) * rstd * pg[i] + pb[i];
        }
      }
      if (valid) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          ptx::sts128(a_rowh + ((j ^ x7) << 4),
                      make_uint4(denf::pack2(v[8 * j], v[8 * j + 1]), denf::pack2(v[8 * j + 2], v[8 * j + 3]),
                                 denf::pack2(v[8 * j + 4], v[8 * j + 5]), denf::pack2(v[8 * j + 6], v[8 * j + 7])));
      }
    };
    // slots of round r: [bias_r, tbias_{r+1}, gamma_{r+1}, beta_{r+1}] = 512 floats, two per thread.
    // The values of round r + 1 are LOADED at the top of round r (global / L2 latency hidden behind the
    // round's epilogues) and STORED to the other buffer at its end; the block barrier at the top of
    // round r + 1 publishes them.  (Loading them at the top of their own round put ~0.5 us of load
    // latency + barrier in front of item 0's epilogue every round.)
    auto load_params = [&](int r, float& p0, float& p1) {
      const int i = etid & (kH - 1);
      if (etid < kH) {
        p0 = r < nl ? a.conv_b[r * kH + i] : a.fc0_b[i];
        p1 = r + 1 < nl ? a.ln_g[(r + 1) * kH + i] : 0.0f;
      } else {
        p0 = r + 1 < nl ? a.time_bias[(r + 1) * kH + i] : 0.0f;
        p1 = r + 1 < nl ? a.ln_b[(r + 1) * kH + i] : 0.0f;
      }
    };
    auto store_params = [&](float* P, float p0, float p1) {
      P[etid] = p0;                  // [bias | tbias]
      P[2 * kH + etid] = p1;         // [gamma | beta]
    };
    for (int64_t pass = group; pass < passes; pass += n_groups) {
      const int64_t p = pass * CG + rank;
      const bool live1 = CG == 2 || 2 * p + 1 < items;
      // ---- embed both items: Conv(5 -> 128, k9) as a weight gather, ReLU, LayerNorm_0 ----
      {
        float* P = s_param;                  // slots 1..3 <- layer 0's norm
        for (int i = etid; i < kH; i += kEpiThreads) {
          P[1 * kH + i] = a.time_bias[i];
          P[2 * kH + i] = a.ln_g[i];
          P[3 * kH + i] = a.ln_b[i];
        }
        float np0, np1;
        load_params(0, np0, np1);
        epi_sync();                          // every thread has left the previous pair's final round
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !live1) continue;
          const int64_t seq = 2 * (2 * p + q) + m;
          const bool valid = (seq < a.n_rows) && (pos < L);
          const uint32_t t_res = t_lane + q * 2 * kH + kH;
          uint8_t* a_rowh = s_a + (size_t)half * plane_bytes + (size_t)(pad + q * (128 + pad) + rq) * 128;
          if (wact) {
            float v[64];
            int tk[kTaps];
#pragma unroll
            for (int t = 0; t < kTaps; ++t) {
              const int li = pos + t - kTaps / 2;
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
              denf::tmem_st_32x32(t_res + c * 32, raw);      // residual stream -> tensor memory
            }
            write_operand(v, P, true, valid, a_rowh);
            denf::tmem_st_wait();
          }
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_operand_ready<CG>(&aready_bar[q]);
        }
        store_params(s_param + 4 * kH, np0, np1);      // round 0 reads buffer 1
      }
      // ---- conv layers: item 0's round r, item 1's round r, item 0's round r+1, ... ----
      for (int r = 0; r < nl; ++r) {
        float* P = s_param + ((r + 1) & 1) * (4 * kH);
        epi_sync();                          // round r's slots visible; round r - 1's buffer is free
        float np0, np1;
        load_params(r + 1, np0, np1);
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !live1) continue;
          const int64_t seq = 2 * (2 * p + q) + m;
          const bool valid = (seq < a.n_rows) && (pos < L);
          const uint32_t t_acc = t_lane + q * 2 * kH, t_res = t_acc + kH;
          uint8_t* a_rowh = s_a + (size_t)half * plane_bytes + (size_t)(pad + q * (128 + pad) + rq) * 128;
          if (etid == 0) stamp(p, r, q, 3);
          ptx::mbar_wait(&tfull_bar[q], tphase[q]);
          tphase[q] ^= 1;
          ptx::tc_fence_after();
          if (etid == 0) stamp(p, r, q, 4);
          if (wact) {
            float v[64];
            uint32_t racc[2][32], rres[2][32];
            ptx::tmem_ld_32x32(t_acc, racc[0]);
            ptx::tmem_ld_32x32(t_res, rres[0]);
            ptx::tmem_ld_32x32(t_acc + 32, racc[1]);
            ptx::tmem_ld_32x32(t_res + 32, rres[1]);
            ptx::tmem_ld_wait();
            if (etid == 0) stamp(p, r, q, 5);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              float pb[32];
              denf::ld_param32(P + ch0 + c * 32, pb);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                // feat += relu(conv + bias)   (models/dnaconv.py:196-200)
                const float f = __uint_as_float(rres[c][i]) + fmaxf(__uint_as_float(racc[c][i]) + pb[i], 0.0f);
                v[c * 32 + i] = f;
                rres[c][i] = __float_as_uint(f);
              }
              denf::tmem_st_32x32(t_res + c * 32, rres[c]);
            }
            write_operand(v, P, r + 1 < nl, valid, a_rowh);
            if (etid == 0) stamp(p, r, q, 6);
            denf::tmem_st_wait();
          }
          ptx::fence_proxy_async_smem();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_operand_ready<CG>(&aready_bar[q]);
          if (etid == 0) stamp(p, r, q, 7);
        }
        store_params(s_param + (r & 1) * (4 * kH), np0, np1);     // round r + 1 reads buffer (r + 2) & 1
      }
      // ---- final_conv: ReLU(1x1) then 1x1 to the 5 logits (models/dnaconv.py:163-165,201) ----
      {
        float* P = s_param + ((nl + 1) & 1) * (4 * kH);
        epi_sync();
        for (int q = 0; q < 2; ++q) {
          if (q == 1 && !live1) continue;
          const int64_t seq = 2 * (2 * p + q) + m;
          const bool valid = (seq < a.n_rows) && (pos < L);
          const uint32_t t_acc = t_lane + q * 2 * kH;
          ptx::mbar_wait(&tfull_bar[q], tphase[q]);
          tphase[q] ^= 1;
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
              denf::ld_param32(P + ch0 + c * 32, pb);
              float y[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) y[i] = fmaxf(__uint_as_float(racc[i]) + pb[i], 0.0f);
#pragma unroll
              for (int j = 0; j < kVocab; ++j) {
                float w[32];
                denf::ld_param32(s_w2 + j * kH + ch0 + c * 32, w);
#pragma unroll
                for (int i = 0; i < 32; ++i) lg[j] += y[i] * w[i];
              }
            }
            float* x = s_lgx + (size_t)(etid & 127) * 8;       // the pair's slot (same rows in both channel halves)
            if (half == 1) {
#pragma unroll
              for (int j = 0; j < kVocab; ++j) x[j] = lg[j];
            }
            pair_sync();
            if (half == 0 && valid) {
              float* o = a.logits + ((size_t)seq * L + pos) * kVocab;
#pragma unroll
              for (int j = 0; j < kVocab; ++j) o[j] = (lg[j] + x[j]) + s_w2[kVocab * kH + j];
            }
            pair_sync();                                        // the slot is reused by the other item's turn
          }
          // the next pair's embed overwrites this item's plane and residual columns: the MMAs that read
          // them have retired (tfull) and the accumulator reads above are complete
          ptx::tc_fence_before();
        }
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  if constexpr (CG == 2) gemm2::cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    gemm2::tmem_dealloc_cg<CG>(tmem_base, 512);
  }
}

}  // namespace dens
}  // namespace svdd
