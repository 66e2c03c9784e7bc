// tcgen05 / TMEM / TMA implicit-GEMM for 1-D convolutions and linears.
//
//   C[(s,l), n] = epilogue( sum_{tap<T} sum_{k<K} A[s, l + (tap - T/2)*dil, k] * W[tap*N + n, k] )
//
// A   : bf16 activations, channels-last [S, L_in, K]  (zero outside [0, L_in))
// W   : bf16 weights, tap-major [T*N, K]              (K contiguous)
// acc : fp32 in tensor memory
//
// One persistent CTA per SM, 320 threads, warp-specialised:
//   warp 0      TMA producer: per (tap, 64-wide K block) one 3-D box of A
//               (64 ch x BL positions x BS sequences, out-of-range positions are
//               zero-filled by TMA = the conv's zero padding) and one 2-D box of W,
//               both 128B-swizzled, into a kStages-deep smem ring;
//   warp 1      MMA issuer: one elected lane issues 4 x tcgen05.mma
//               (M=128, N=BN, K=16) per stage into a double-buffered TMEM
//               accumulator, tcgen05.commit releases the smem slot / signals the
//               epilogue;
//   warps 2-9   epilogue: thread = (output row, column half); tcgen05.ld 32 lanes x
//               32 columns at a time, software-pipelined against the math; the
//               per-column vectors of the tile (bias, BN scale/shift, head weights)
//               are staged in shared memory once per tile; fused bias / BN-affine /
//               activation / residual / LayerNorm (row partials exchanged between
//               the two halves through smem) / attention-pool / head-dot,
//               vectorised global stores.
// Taps whose whole A tile falls in the zero padding are skipped by both the
// producer and the issuer (dilation-64 taps on 200-long sequences).
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace svdd {

enum ActKind { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_GELU_TANH = 3 };   // 2: enformer x*sigmoid(1.702x); 3: nn.GELU('tanh')
enum DType { DT_NONE = 0, DT_BF16 = 1, DT_F32 = 2 };
enum EpiMode { EPI_GENERIC = 0, EPI_DEN_LN = 1, EPI_DEN_FINAL = 2, EPI_POOL = 3, EPI_HEADDOT = 4,
               EPI_PAIR = 5, EPI_POOL2 = 6 };   // 5, 6: conv_gemm2.cuh only

struct EpiParams {
  // generic chain on v = acc:  v = v*scale + shift ; v += bias ; [act] ; v += res ; [act]
  const float* bias = nullptr;
  const float* scale = nullptr;
  const float* shift = nullptr;
  int act = ACT_NONE;
  int act_after_res = 0;
  const void* res = nullptr;
  int res_dtype = DT_NONE;
  int64_t ld_res = 0;
  void* out = nullptr;
  int out_dtype = DT_NONE;
  int64_t ld_out = 0;
  // second output: out2 = act2(v * scale2 + shift2)   (next layer's BN+GELU operand)
  void* out2 = nullptr;
  int out2_dtype = DT_NONE;
  int64_t ld_out2 = 0;
  const float* scale2 = nullptr;
  const float* shift2 = nullptr;
  int act2 = ACT_NONE;
  // EPI_DEN_LN: out = residual stream (fp32, read+written), out2 = LN'd operand
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  const float* ln_tbias = nullptr;
  int ln_enable = 0;
  // EPI_DEN_FINAL: logits[r, j] = b2[j] + sum_c relu(acc+bias)[c] * w2[j, c]
  const float* w2 = nullptr;
  const float* b2 = nullptr;
  // EPI_POOL: values being pooled (bf16 [S*L_in, N])
  const void* pool_vals = nullptr;
  // EPI_PAIR : out = y[2j] and out2 = y[2j+1] - y[2j], both bf16 [S, ceil(L/2), N] (res = block input)
  // EPI_POOL2: A = yd; res = y0, res2 = yd (bf16 [rows, N]); pooled = y0 + sigmoid(acc) * yd ->
  //            out (fp32, optional) and out2 = act2(pooled * scale2 + shift2) (bf16, optional)
  const void* res2 = nullptr;
  int64_t ld_res2 = 0;
  // GemmShape::K2 > 0: second operand pair (bf16 [rows, K2] and bf16 [N, K2]); replaces `res`
  const void* a2 = nullptr;
  const void* w2k = nullptr;
  // rows between consecutive sequences of res / res2 / out / out2 (conv_gemm2; 0 = densely packed)
  int res_pitch = 0, res2_pitch = 0, out_pitch = 0, out2_pitch = 0;
  // EPI_HEADDOT: partials[r, 2*n_tile + half] = sum_{n in that half tile} v[n] * head_w[n]
  const float* head_w = nullptr;
  float* partials = nullptr;
  // tuning switch (set by the launcher from SVDD_EPI_PREFETCH, default on): fetch bf16
  // residual / pooled rows ahead of the accumulator wait
  int prefetch = 1;
  // conv_gemm2, set by the launcher: `res` is `out` itself (fp32, in place) -> the residual is
  // never loaded, the epilogue stores acc (+bias, act) with a TMA reduce-add into `out`
  int res_reduce = 0;
  // conv_gemm2, EPI_PAIR / EPI_POOL2 on 8 epilogue warps, set by the launcher (SVDD_SLAB32, default on):
  // 32-column slabs (128 rows x 64 B), FOUR per column half, so that the TMA loads of step k + 1 are in
  // flight while step k is computed (tensor maps of res / res2 / out / out2 carry 32-column boxes)
  int slab32 = 0;
  // debugging aid (SVDD_TIMELINE=1): CTA 0 stamps clock64() at its pipeline milestones
  unsigned long long* timeline = nullptr;
};

struct GemmShape {
  int S = 0;       // sequences
  int L = 0;       // output positions per sequence
  int L_in = 0;    // input positions per sequence (EPI_POOL: 2 inputs per output)
  int K = 0, N = 0;
  int taps = 1, dil = 1;
  int BL = 128, BS = 1;  // tile = BL positions x BS sequences (BL*BS <= 128)
  // column window of one launch (conv_gemm2 only): this launch computes columns
  // [n_off, n_off + N) of a weight / output that is N_w columns wide (0 = N)
  int n_off = 0, N_w = 0;
  // K-concatenated second operand (conv_gemm2, EPI_PAIR, 1x1): C += A2[rows, K2] * W2[N, K2]^T;
  // A2 has A's row structure, pointers in EpiParams::a2 / w2k
  int K2 = 0;
  // conv_gemm2 HALO mode (flat rows, S == 1, BL == 128, taps odd): the A tile is fetched once per
  // K block with taps-1 extra rows and each tap is a row offset of the smem descriptor.  The
  // caller guarantees `taps/2` zero rows between consecutive sequences of the flat row space.
  int halo = 0;
  // conv_gemm2: L2-prefetch this CTA's weight slice of its first tile at kernel start (set by the
  // launcher for launches with at most two tiles per CTA pair; SVDD_PREFETCH_W=0 disables)
  int prefetch_w = 0;
  // conv_gemm2: walk the row tiles from the last to the first.  A layer that consumes what the previous
  // launch wrote in increasing row order then starts on the rows that are still in L2 ("snake" order)
  int reverse = 0;
  // rows between consecutive sequences of A when they are not densely packed (0 = L_in)
  int a_pitch = 0;
  // rows that carry real data, for the FLOP accounting of padded layouts (0 = S * L)
  int64_t useful_rows = 0;
};

namespace gemm_detail {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kEpiWarps = 8;                       // 2 warps per TMEM lane quadrant (column halves)
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;         // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue
constexpr int kSmemBudget = 196 * 1024;            // operand ring
constexpr int kParamVecs = 6;                      // bias, scale, shift, scale2, shift2, head_w per tile
enum ParamSlot { P_BIAS = 0, P_SCALE = 1, P_SHIFT = 2, P_SCALE2 = 3, P_SHIFT2 = 4, P_HEADW = 5,
                 P_LN_TBIAS = 1, P_LN_GAMMA = 2, P_LN_BETA = 3 };

template <int BN, int MODE>
struct Cfg {
  static constexpr int kAcc = (MODE == EPI_POOL) ? 2 : 1;
  static constexpr int kABytes = kAcc * kBM * kBK * 2;
  static constexpr int kBBytes = BN * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kTmemColsRaw = 2 * kAcc * BN;
  static constexpr int kTmemCols = kTmemColsRaw <= 32 ? 32 : kTmemColsRaw <= 64 ? 64
                                   : kTmemColsRaw <= 128 ? 128 : kTmemColsRaw <= 256 ? 256 : 512;
  // columns per epilogue thread: half 0 takes ceil(BN/64)*32 columns, half 1 the rest, so that
  // both are whole 32-column tcgen05.ld chunks for BN in {64,128,192,224,256}
  static constexpr int kHalf = ((BN + 63) / 64) * 32;
  static constexpr int kHalf1 = BN - kHalf;
  static constexpr int kChunks = kHalf / 32;       // chunks of half 0 (>= chunks of half 1)
  static constexpr int kChunks1 = kHalf1 / 32;
  static constexpr int kParamBytes = 2 * kParamVecs * BN * 4;          // double-buffered per tile
  static constexpr int kXchBytes = 2 * kBM * 8 * 4;                    // half <-> half exchange
  static constexpr int kW2Bytes = (MODE == EPI_DEN_FINAL) ? (kVocab * 128 + 8) * 4 : 0;
  static constexpr int kExtraBytes = kParamBytes + kXchBytes + kW2Bytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kExtraBytes + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(kTmemColsRaw <= 512, "accumulators exceed tensor memory");
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "N tile must be a multiple of 32 in [64,256]");
  static_assert(MODE == EPI_GENERIC || MODE == EPI_HEADDOT || kHalf == kHalf1, "even halves required");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget exceeded");
};

// single-instruction MUFU approximations (2 ulp): enough for activations that are rounded
// to bf16 right after
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// enformer GELU: x * sigmoid(1.702 x) = x / (1 + 2^(-1.702 * log2(e) * x))
__device__ __forceinline__ float gelu_enformer(float v) {
  return v * fast_rcp(1.0f + fast_ex2(-2.4554669595930157f * v));
}
// nn.GELU(approximate='tanh'): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  (models/dit.py:224)
__device__ __forceinline__ float gelu_tanh_approx(float v) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.7978845608028654f * fmaf(0.044715f * v, v * v, v)));
  const float h = 0.5f * v;
  return fmaf(h, t, h);
}
// activation over a 32-column chunk; the kind is tested once per chunk, not per element
__device__ __forceinline__ void apply_act32(float* v, int act) {
  if (act == ACT_RELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
  } else if (act == ACT_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_enformer(v[i]);
  } else if (act == ACT_GELU_TANH) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_approx(v[i]);
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
}

// store / load 32 consecutive channels of one row
__device__ __forceinline__ void store_row32(void* base, int dtype, int64_t off, const float* v) {
  if (dtype == DT_BF16) {
    uint4* p = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      p[i] = make_uint4(pack_bf16x2(v[8 * i], v[8 * i + 1]), pack_bf16x2(v[8 * i + 2], v[8 * i + 3]),
                        pack_bf16x2(v[8 * i + 4], v[8 * i + 5]), pack_bf16x2(v[8 * i + 6], v[8 * i + 7]));
  } else {
    float4* p = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}
__device__ __forceinline__ void load_row32(const void* base, int dtype, int64_t off, float* v) {
  if (dtype == DT_BF16) {
    const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint4 u = p[i];
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        v[8 * i + 2 * j] = __low2float(h);
        v[8 * i + 2 * j + 1] = __high2float(h);
      }
    }
  } else {
    const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 f = p[i];
      v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
    }
  }
}
// 32 consecutive per-column parameters from shared memory (warp-uniform address -> broadcast)
__device__ __forceinline__ void load_param32(const float* p, float* v) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 f = ptx::lds128f(p + 4 * i);
    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
  }
}

template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const GemmShape g, const EpiParams ep) {
  using C = Cfg<BN, MODE>;
  constexpr int kStages = C::kStages;
  constexpr int kAcc = C::kAcc;
  constexpr int kHalf = C::kHalf;
  constexpr int kChunks = C::kChunks;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint8_t* extra = smem + kStages * C::kStageBytes;
  float* s_param = reinterpret_cast<float*>(extra);                                  // [2][kParamVecs][BN]
  float* s_xch = reinterpret_cast<float*>(extra + C::kParamBytes);                   // [2][128][8]
  float* s_w2 = reinterpret_cast<float*>(extra + C::kParamBytes + C::kXchBytes);     // DEN_FINAL only
  uint64_t* bars = reinterpret_cast<uint64_t*>(extra + C::kExtraBytes);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tiles_l = ceil_div(g.L, g.BL);
  const int tiles_s = ceil_div(g.S, g.BS);
  const int n_tiles = g.N / BN;
  const int64_t total_tiles = (int64_t)tiles_l * tiles_s * n_tiles;
  const int kblocks = g.K / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        ptx::mbar_init(&full_bar[i], 1);
        ptx::mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&tfull_bar[i], 1);
        ptx::mbar_init(&tempty_bar[i], kEpiWarps);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, C::kTmemCols);
    ptx::tmem_relinquish();
  }
  if constexpr (MODE == EPI_DEN_FINAL) {
    for (int i = threadIdx.x; i < kVocab * 128; i += kThreads) s_w2[i] = ep.w2[i];
    if (threadIdx.x < kVocab) s_w2[kVocab * 128 + threadIdx.x] = ep.b2[threadIdx.x];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();       // everything above overlapped the previous kernel's tail
  pdl_trigger();

  auto tile_coords = [&](int64_t t, int& n0, int& s0, int& l0) {
    const int nt = (int)(t % n_tiles);
    const int64_t mt = t / n_tiles;
    n0 = nt * BN;
    s0 = (int)(mt / tiles_l) * g.BS;
    l0 = (int)(mt % tiles_l) * g.BL;
  };
  auto tap_active = [&](int tap, int l0, int& l_start) {
    l_start = l0 + (tap - g.taps / 2) * g.dil;
    return (l_start + g.BL > 0) && (l_start < g.L_in);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (ptx::elect_one()) {
      uint32_t stage = 0, phase = 0;
      const uint32_t tx_bytes = (uint32_t)(kAcc * g.BL * g.BS * kBK * 2 + C::kBBytes);
      for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        int n0, s0, l0;
        tile_coords(t, n0, s0, l0);
        for (int tap = 0; tap < g.taps; ++tap) {
          int l_start;
          if (MODE != EPI_POOL && !tap_active(tap, l0, l_start)) continue;
          for (int kb = 0; kb < kblocks; ++kb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * C::kStageBytes;
            uint8_t* sb = sa + C::kABytes;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
            if (MODE == EPI_POOL) {
              ptx::tma_load_4d(sa, &tmA, &full_bar[stage], kb * kBK, 0, l0, s0);
              ptx::tma_load_4d(sa + kBM * kBK * 2, &tmA, &full_bar[stage], kb * kBK, 1, l0, s0);
            } else {
              ptx::tma_load_3d(sa, &tmA, &full_bar[stage], kb * kBK, l_start, s0);
            }
            ptx::tma_load_2d(sb, &tmW, &full_bar[stage], kb * kBK, tap * g.N + n0);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kBM, BN);
    uint32_t stage = 0, phase = 0;
    uint32_t acc_stage = 0, acc_phase = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n0, s0, l0;
      tile_coords(t, n0, s0, l0);
      ptx::mbar_wait(&tempty_bar[acc_stage], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc_stage * (kAcc * BN);
      uint32_t first = 1;
      for (int tap = 0; tap < g.taps; ++tap) {
        int l_start;
        if (MODE != EPI_POOL && !tap_active(tap, l0, l_start)) continue;
        for (int kb = 0; kb < kblocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t sa = ptx::smem_u32(stage_base + stage * C::kStageBytes);
            const uint32_t sb = sa + C::kABytes;
            const uint64_t da = ptx::make_kmajor_sw128_desc(sa);
            const uint64_t db = ptx::make_kmajor_sw128_desc(sb);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // +32 B per 16-element K step inside the 128 B swizzle atom
              ptx::umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, first ? (k > 0) : 1u);
              if (kAcc == 2) {
                const uint64_t da1 = ptx::make_kmajor_sw128_desc(sa + kBM * kBK * 2);
                ptx::umma_bf16(tmem_d + BN, da1 + 2 * k, db + 2 * k, idesc, first ? (k > 0) : 1u);
              }
            }
            ptx::umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          first = 0;
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
      if (ptx::elect_one()) ptx::umma_commit(&tfull_bar[acc_stage]);
      __syncwarp();
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue: 8 warps, thread = (row, column half) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;           // TMEM lane quadrant this warp may read
    const int half = ew >> 2;            // which half of the tile's columns
    const int etid = threadIdx.x - 64;
    const int r = quad * 32 + lane;      // row of the tile owned by this thread
    float* xch_mine = s_xch + (half * kBM + r) * 8;
    const float* xch_other = s_xch + ((half ^ 1) * kBM + r) * 8;
    uint32_t acc_stage = 0, acc_phase = 0;
    for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      int n0, s0, l0;
      tile_coords(t, n0, s0, l0);
      const int s = s0 + r / g.BL, l = l0 + r % g.BL;
      const bool valid = (r < g.BL * g.BS) && (s < g.S) && (l < g.L);
      const int64_t row = (int64_t)s * g.L + l;

      // per-column parameters of this N tile -> shared memory (double-buffered with the accumulator)
      float* P = s_param + acc_stage * (kParamVecs * BN);
      for (int i = etid; i < BN; i += kEpiThreads) {
        if constexpr (MODE == EPI_DEN_LN) {
          P[P_BIAS * BN + i] = ep.bias[i];
          if (ep.ln_enable) {
            P[P_LN_TBIAS * BN + i] = ep.ln_tbias[i];
            P[P_LN_GAMMA * BN + i] = ep.ln_gamma[i];
            P[P_LN_BETA * BN + i] = ep.ln_beta[i];
          }
        } else {
          if (ep.bias) P[P_BIAS * BN + i] = ep.bias[n0 + i];
          if (ep.scale) { P[P_SCALE * BN + i] = ep.scale[n0 + i]; P[P_SHIFT * BN + i] = ep.shift[n0 + i]; }
          if (ep.scale2) { P[P_SCALE2 * BN + i] = ep.scale2[n0 + i]; P[P_SHIFT2 * BN + i] = ep.shift2[n0 + i]; }
          if (MODE == EPI_HEADDOT) P[P_HEADW * BN + i] = ep.head_w[n0 + i];
        }
      }
      epi_bar_sync();
      if constexpr (MODE == EPI_DEN_LN || MODE == EPI_DEN_FINAL) {
        ptx::mbar_wait(&tfull_bar[acc_stage], acc_phase);
        ptx::tc_fence_after();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc_stage * (kAcc * BN) + half * kHalf;
      const int cbase = half * kHalf;    // first column (within the tile) of this thread
      const int my_chunks = half ? C::kChunks1 : kChunks;   // warp-uniform

      if constexpr (MODE == EPI_GENERIC || MODE == EPI_POOL || MODE == EPI_HEADDOT) {
        float head_acc = 0.0f;
        uint32_t raw[2][32];
        // bf16 residual / pooled values do not depend on the accumulator: fetch them two chunks
        // ahead so their latency hides behind the wait for the MMA and the previous chunk's math
        uint4 pre[2][2][4];
        const bool pre_res = (MODE == EPI_GENERIC || MODE == EPI_HEADDOT) && ep.res != nullptr &&
                             ep.res_dtype == DT_BF16 && valid && ep.prefetch;
        const bool has1 = (MODE == EPI_POOL) && (2 * l + 1) < g.L_in;
        const __nv_bfloat16* pre_base = nullptr;
        if (MODE == EPI_POOL) {
          if (valid) pre_base = reinterpret_cast<const __nv_bfloat16*>(ep.pool_vals) +
                                ((int64_t)s * g.L_in + 2 * l) * g.N + n0 + cbase;
        } else if (pre_res) {
          pre_base = reinterpret_cast<const __nv_bfloat16*>(ep.res) + row * ep.ld_res + n0 + cbase;
        }
        auto prefetch = [&](int c) {
          if (pre_base == nullptr || c >= kChunks || (C::kChunks1 != kChunks && c >= my_chunks)) return;
          const uint4* p = reinterpret_cast<const uint4*>(pre_base + c * 32);
#pragma unroll
          for (int i = 0; i < 4; ++i) pre[c & 1][0][i] = p[i];
          if (MODE == EPI_POOL && has1) {
            const uint4* q = reinterpret_cast<const uint4*>(pre_base + g.N + c * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i) pre[c & 1][1][i] = q[i];
          }
        };
        auto unpack = [&](const uint4* u, float* o) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
              o[8 * i + 2 * j] = __low2float(hh);
              o[8 * i + 2 * j + 1] = __high2float(hh);
            }
          }
        };
        if (ep.prefetch) {
          prefetch(0);
          prefetch(1);
        }
        ptx::mbar_wait(&tfull_bar[acc_stage], acc_phase);
        ptx::tc_fence_after();
        if constexpr (MODE != EPI_POOL) ptx::tmem_ld_32x32(taddr, raw[0]);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          if (C::kChunks1 != kChunks && c >= my_chunks) continue;   // uneven halves (BN 192/224)
          const int c0 = cbase + c * 32;   // column within the tile
          const int n = n0 + c0;           // global column
          float v[32];
          if constexpr (MODE == EPI_POOL) {
            ptx::tmem_ld_32x32(taddr + c * 32, raw[0]);
            ptx::tmem_ld_32x32(taddr + BN + c * 32, raw[1]);
            ptx::tmem_ld_wait();
            if (valid) {
              float y0[32], y1[32];
              if (ep.prefetch) {
                unpack(pre[c & 1][0], y0);
                if (has1) unpack(pre[c & 1][1], y1);
                prefetch(c + 2);
              } else {
                load_row32(pre_base, DT_BF16, c * 32, y0);
                if (has1) load_row32(pre_base, DT_BF16, g.N + c * 32, y1);
              }
              if (has1) {
                // softmax over the pair = sigmoid of the logit difference
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                  const float d = __uint_as_float(raw[1][i]) - __uint_as_float(raw[0][i]);
                  const float w0 = fast_rcp(1.0f + fast_ex2(1.4426950408889634f * d));
                  v[i] = y1[i] + w0 * (y0[i] - y1[i]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = y0[i];
              }
            }
          } else {
            ptx::tmem_ld_wait();
            if (c + 1 < my_chunks) ptx::tmem_ld_32x32(taddr + (c + 1) * 32, raw[(c + 1) & 1]);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[c & 1][i]);
          }
          if (valid) {
            float pv[32];
            if constexpr (MODE != EPI_POOL) {
              if (ep.scale != nullptr) {
                float ps[32];
                load_param32(P + P_SCALE * BN + c0, ps);
                load_param32(P + P_SHIFT * BN + c0, pv);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = v[i] * ps[i] + pv[i];
              }
              if (ep.bias != nullptr) {
                load_param32(P + P_BIAS * BN + c0, pv);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += pv[i];
              }
              if (!ep.act_after_res) apply_act32(v, ep.act);
              if (ep.res != nullptr) {
                if (pre_res) {
                  unpack(pre[c & 1][0], pv);
                  prefetch(c + 2);
                } else {
                  load_row32(ep.res, ep.res_dtype, row * ep.ld_res + n, pv);
                }
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += pv[i];
              }
              if (ep.act_after_res) apply_act32(v, ep.act);
            }
            if constexpr (MODE == EPI_HEADDOT) {
              load_param32(P + P_HEADW * BN + c0, pv);
#pragma unroll
              for (int i = 0; i < 32; ++i) head_acc += v[i] * pv[i];
            }
            if (ep.out != nullptr) store_row32(ep.out, ep.out_dtype, row * ep.ld_out + n, v);
            if (ep.out2 != nullptr) {
              if (ep.scale2 != nullptr) {
                float ps[32];
                load_param32(P + P_SCALE2 * BN + c0, ps);
                load_param32(P + P_SHIFT2 * BN + c0, pv);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = v[i] * ps[i] + pv[i];
              }
              apply_act32(v, ep.act2);
              store_row32(ep.out2, ep.out2_dtype, row * ep.ld_out2 + n, v);
            }
          }
        }
        if (MODE == EPI_HEADDOT && valid)
          ep.partials[row * (2 * n_tiles) + 2 * (n0 / BN) + half] = head_acc;
      } else if constexpr (MODE == EPI_DEN_LN) {
        // denoiser layer i: feat += relu(conv + bias); h_next = LN_{i+1}(feat + tbias_{i+1})
        // (models/dnaconv.py:189-200).  BN == N == 128; the two threads that share a row
        // (column halves) combine their LayerNorm partial sums through shared memory.
        float v[kHalf];
        float* feat = reinterpret_cast<float*>(ep.out);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          uint32_t raw[32];
          ptx::tmem_ld_32x32(taddr + c * 32, raw);
          ptx::tmem_ld_wait();
          float rr[32], pb[32];
          load_param32(P + P_BIAS * BN + cbase + c * 32, pb);
          if (valid) load_row32(feat, DT_F32, row * BN + cbase + c * 32, rr);
#pragma unroll
          for (int i = 0; i < 32; ++i)
            v[c * 32 + i] = valid ? rr[i] + fmaxf(__uint_as_float(raw[i]) + pb[i], 0.0f) : 0.0f;
        }
        if (valid) {
#pragma unroll
          for (int c = 0; c < kChunks; ++c) store_row32(feat, DT_F32, row * BN + cbase + c * 32, v + c * 32);
        }
        if (ep.ln_enable) {            // block-uniform
          float sum = 0.0f;
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            float pt[32];
            load_param32(P + P_LN_TBIAS * BN + cbase + c * 32, pt);
#pragma unroll
            for (int i = 0; i < 32; ++i) { v[c * 32 + i] += pt[i]; sum += v[c * 32 + i]; }
          }
          xch_mine[0] = sum;
          epi_bar_sync();
          const float mean = (sum + xch_other[0]) * (1.0f / BN);
          float sq = 0.0f;
#pragma unroll
          for (int i = 0; i < kHalf; ++i) { const float d = v[i] - mean; sq += d * d; }
          xch_mine[1] = sq;
          epi_bar_sync();
          const float rstd = rsqrtf((sq + xch_other[1]) * (1.0f / BN) + 1e-5f);
#pragma unroll
          for (int c = 0; c < kChunks; ++c) {
            float pg[32], pbt[32];
            load_param32(P + P_LN_GAMMA * BN + cbase + c * 32, pg);
            load_param32(P + P_LN_BETA * BN + cbase + c * 32, pbt);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[c * 32 + i] = (v[c * 32 + i] - mean) * rstd * pg[i] + pbt[i];
          }
        }
        if (valid) {
#pragma unroll
          for (int c = 0; c < kChunks; ++c)
            store_row32(ep.out2, DT_BF16, row * BN + cbase + c * 32, v + c * 32);
        }
      } else if constexpr (MODE == EPI_DEN_FINAL) {
        // final_conv: relu(1x1) then 1x1 to the 5 logits (models/dnaconv.py:163-165,201)
        float lg[kVocab];
#pragma unroll
        for (int j = 0; j < kVocab; ++j) lg[j] = 0.0f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          uint32_t raw[32];
          ptx::tmem_ld_32x32(taddr + c * 32, raw);
          ptx::tmem_ld_wait();
          float pb[32];
          load_param32(P + P_BIAS * BN + cbase + c * 32, pb);
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float y = fmaxf(__uint_as_float(raw[i]) + pb[i], 0.0f);
#pragma unroll
            for (int j = 0; j < kVocab; ++j) lg[j] += y * s_w2[j * 128 + cbase + c * 32 + i];
          }
        }
        if (half == 1) {
#pragma unroll
          for (int j = 0; j < kVocab; ++j) xch_mine[j] = lg[j];
        }
        epi_bar_sync();
        if (half == 0 && valid) {
          float* o = reinterpret_cast<float*>(ep.out) + row * kVocab;
#pragma unroll
          for (int j = 0; j < kVocab; ++j) o[j] = lg[j] + xch_other[j] + s_w2[kVocab * 128 + j];
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tempty_bar[acc_stage]);
      if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace gemm_detail

// Host-side launcher (conv_gemm.cu).  A is [S, L_in, K] bf16 (for EPI_POOL the
// rows 2l, 2l+1 of each sequence feed output row l), W is [taps*N, K] bf16.
int launch_conv_gemm(const void* A, const void* W, const GemmShape& shape, int mode,
                     const EpiParams& ep, cudaStream_t stream);

// Number of partials per row that EPI_HEADDOT writes for this shape (2 per N tile: one per
// column half).
int conv_gemm_n_tiles(const GemmShape& shape, int mode);

// 128B-swizzled 2-D tensor map over a row-major bf16 matrix [rows, inner] (inner contiguous).
int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows,
                        uint32_t box_inner, uint32_t box_rows);
// 128B-swizzled 3-D tensor map over bf16 [d2][d1][d0] (d0 contiguous; strides in bytes)
int encode_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                        uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2);
// the same over an fp32 matrix (box_inner * 4 bytes <= 128)
int encode_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows,
                       uint32_t box_inner, uint32_t box_rows);

// per-launch profile window (svdd_profile_begin / _end): kernels outside conv_gemm.cu that
// belong to the tensor-core family record themselves here
bool gemm_prof_on();
void gemm_prof_record(cudaEvent_t e0, cudaEvent_t e1, const GemmShape& g, int bn, int mode, double flops);

// Picks (BL, BS) for a conv over sequences of length L: whole-sequence tiles when
// L <= 128 (several sequences per tile), else 128-position tiles.
void choose_row_tiling(int L, int taps, GemmShape* shape);

}  // namespace svdd
