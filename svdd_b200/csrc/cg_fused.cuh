// The conv tower of the RNA value net / reward oracle as ONE persistent kernel (reference:
// ConvGRUTrunk = Stem + ConvTower, Enformer.py:1359-1426, ConvBlock order 'CDNRA' :2266-2292):
//   tokens -> Conv(4->64, k15) + ReLU -> n x [ Conv(64->64, k5) -> BN -> (+x) -> ReLU ] -> bf16 [rows, L, 64]
// for sequences of at most 62 positions (RNA: L = 50).  As six launches (one-hot stem + five
// first-generation implicit GEMMs with 128 x 64 tiles) the tower moved every activation through
// HBM / L2 once per layer and re-fetched the 40 KB of a layer's weights for each of its 20 000 tiles
// (3 GB of L2 traffic per launch at 51 200 x 50): 1.96 ms of the 3.65 ms scoring pass.
//
// Here one CTA works on items of 8 sequences = 4 tiles of 128 rows, two sequences per tile 64 rows
// apart (rows L..63 of a slot stay zero = the conv padding, so every k5 tap is a row offset of the
// UMMA shared-memory descriptor and ONE 128-row MMA serves both sequences).  The activations of an
// item never leave the SM: the bf16 operand planes live in shared memory and are updated in place
// (the residual is read from the plane row the thread is about to overwrite), accumulators in
// TMEM (4 x 64 columns), a layer's weights are fetched once per item (TMA, double-buffered by
// layer).  The stem rides the same machinery as "layer 0": the epilogue threads write the one-hot
// im2col row of their position (15 taps x 4 tokens = 60 of K = 64 features, exact in bf16) into the
// plane and the stem weights enter as a bf16 hi + lo pair of tap tiles at offset 0 (fp32-accurate:
// the dropped term is 2^-17 relative) -- as a gather-add of fp32 weight rows from shared memory the
// stem was 45 % of this kernel (4-way bank conflicts: every lane reads the row of ITS token).
// The four tiles form a software pipeline: the MMA warp issues tile t of layer l as soon
// as the epilogue of tile t, layer l-1 is done, and two epilogue teams of four warps each take the
// even / odd tiles, so MMAs, TMEM reads and the epilogue arithmetic of different tiles overlap.
#pragma once
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace svdd {
namespace cgf {

constexpr int kC = 64;
constexpr int kTiles = 4;
constexpr int kSeqPerItem = 2 * kTiles;
constexpr int kPad = 8;                              // zero rows either side of a plane (taps reach +-2)
constexpr int kPlaneRows = kPad + 128 + kPad;
constexpr int kPlaneBytes = kPlaneRows * 128;
constexpr int kTapBytes = kC * 128;                  // one tap: [64 out][64 in] bf16, K-major
constexpr int kMaxTaps = 5;
constexpr int kWLayerBytes = kMaxTaps * kTapBytes;
constexpr int kMaxLayers = 16;
constexpr int kStemTapsMax = 15;
constexpr int kParamBytes = (kMaxLayers + 1) * 2 * kC * 4;        // (scale, shift) per layer; layer 0 = the stem (1, bias)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBytes = kTiles * kPlaneBytes + 2 * kWLayerBytes + kParamBytes + 256 + 1024;

struct Args {
  const void* tokens;        // [rows, L] uint8 or int64
  const float* ss;           // [1 + n_layers][2][64]: scale, shift (layer 0 = stem: 1, bias; BN folded; scale = 1 without BN)
  __nv_bfloat16* out;        // [rows, L, 64]
  int64_t rows;
  int L, n_layers, taps, stem_taps, residual;
};

__device__ __forceinline__ uint32_t pack2(float x, float y) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  return *reinterpret_cast<const uint32_t*>(&h);
}

template <typename Tok>
__global__ void __launch_bounds__(kThreads, 1)
cg_convstack_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_plane = smem;                                           // [tile][kPlaneRows][64 bf16]
  uint8_t* s_w = s_plane + kTiles * kPlaneBytes;                     // [2][taps][64][64 bf16]
  float* s_par = reinterpret_cast<float*>(s_w + 2 * kWLayerBytes);   // [1 + layers][2][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_par) + kParamBytes);
  uint64_t* wl_full = bars;              // [2]
  uint64_t* wl_empty = bars + 2;         // [2]
  uint64_t* acc_full = bars + 4;         // [kTiles]
  uint64_t* plane_ready = bars + 8;      // [kTiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L, nl = a.n_layers + 1, taps = a.taps;      // nl counts the stem as layer 0 (two "taps": W hi, W lo)
  const int64_t items = (a.rows + kSeqPerItem - 1) / kSeqPerItem;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmW);
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        ptx::mbar_init(&wl_full[i], 1);
        ptx::mbar_init(&wl_empty[i], 1);
      }
      for (int t = 0; t < kTiles; ++t) {
        ptx::mbar_init(&acc_full[t], 1);
        ptx::mbar_init(&plane_ready[t], kEpiWarps / 2);
      }
      ptx::fence_barrier_init();
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, kTiles * kC);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < kTiles * kPlaneBytes / 16; i += kThreads)
    reinterpret_cast<uint4*>(s_plane)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = threadIdx.x; i < nl * 2 * kC; i += kThreads) s_par[i] = a.ss[i];
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===================== weight producer: one layer (taps x 8 KB) per transaction =====================
    if (ptx::elect_one()) {
      uint32_t cnt = 0;
      for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
        for (int l = 0; l < nl; ++l, ++cnt) {
          const uint32_t buf = cnt & 1;
          ptx::mbar_wait(&wl_empty[buf], ((cnt >> 1) & 1) ^ 1);
          const int nt = l == 0 ? 2 : taps;
          ptx::mbar_arrive_expect_tx(&wl_full[buf], (uint32_t)(nt * kTapBytes));
          for (int t = 0; t < nt; ++t)
            ptx::tma_load_2d(s_w + buf * kWLayerBytes + t * kTapBytes, &tmW, &wl_full[buf], 0, (l * taps + t) * kC);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, kC);
    uint32_t cnt = 0;
    uint32_t pr = 0;                        // plane_ready completions consumed per tile (same for all tiles)
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      for (int l = 0; l < nl; ++l, ++cnt, ++pr) {
        const uint32_t buf = cnt & 1;
        ptx::mbar_wait(&wl_full[buf], (cnt >> 1) & 1);
        for (int t = 0; t < kTiles; ++t) {
          ptx::mbar_wait(&plane_ready[t], pr & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t plane = ptx::smem_u32(s_plane + t * kPlaneBytes);
            const int nt = l == 0 ? 2 : taps;
            for (int tap = 0; tap < nt; ++tap) {
              const int o = l == 0 ? 0 : tap - taps / 2;
              const uint64_t da = ptx::make_kmajor_sw128_desc(plane + (uint32_t)(kPad + o) * 128u);
              const uint64_t db = ptx::make_kmajor_sw128_desc(ptx::smem_u32(s_w + buf * kWLayerBytes + tap * kTapBytes));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_bf16(tmem_base + t * kC, da + 2 * k, db + 2 * k, idesc, (uint32_t)((tap | k) != 0));
            }
            ptx::umma_commit(&acc_full[t]);
            if (t == kTiles - 1) ptx::umma_commit(&wl_empty[buf]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== epilogue: two teams of four warps, thread = one row of a tile =====================
    const int ew = warp - 2;
    const int team = ew >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;               // tile row
    const int slot = r >> 6, pos = r & 63;        // sequence slot of the tile, position within the sequence
    const int x7 = r & 7;                         // kPad is a multiple of 8
    const Tok* tokens = reinterpret_cast<const Tok*>(a.tokens);
    uint32_t af = 0;                              // acc_full completions consumed (per tile of this team)
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x) {
      // ---- stem operand: the one-hot im2col row of this position (feature 4 j + tok <- token at pos + j - half) ----
      for (int t = team; t < kTiles; t += 2) {
        const int64_t seq = it * kSeqPerItem + 2 * t + slot;
        const bool valid = seq < a.rows && pos < L;
        uint8_t* prow = s_plane + t * kPlaneBytes + (size_t)(kPad + r) * 128;
        if (valid) {
          int tk[16];
          const int half = a.stem_taps / 2;
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int p = pos + j - half;
            int v = -1;
            if (j < a.stem_taps && p >= 0 && p < L) v = load_tok(tokens, (size_t)seq * L + p);   // 4 = mask: no feature set
            tk[j] = v;
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {            // 16-byte chunk c = features 8c .. 8c+7 = taps 2c, 2c+1
            uint32_t w[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {          // word q: features 8c + 2q, 8c + 2q + 1 -> tap 2c + (q >> 1), tokens 2 (q & 1), +1
              const int tok = tk[2 * c + (q >> 1)];
              w[q] = (tok == 2 * (q & 1) ? 0x3F80u : 0u) | (tok == 2 * (q & 1) + 1 ? 0x3F800000u : 0u);
            }
            ptx::sts128(prow + ((c ^ x7) << 4), make_uint4(w[0], w[1], w[2], w[3]));
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&plane_ready[t]);
      }
      // ---- layers: 0 = stem (ReLU(acc + bias)), then the conv blocks ----
      for (int l = 0; l < nl; ++l, ++af) {
        const float* P = s_par + l * 2 * kC;
        const bool last = l + 1 == nl;
        for (int t = team; t < kTiles; t += 2) {
          const int64_t seq = it * kSeqPerItem + 2 * t + slot;
          const bool valid = seq < a.rows && pos < L;
          uint8_t* prow = s_plane + t * kPlaneBytes + (size_t)(kPad + r) * 128;
          ptx::mbar_wait(&acc_full[t], af & 1);
          ptx::tc_fence_after();
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + t * kC;
          uint32_t raw[2][32];
          ptx::tmem_ld_32x32(taddr, raw[0]);
          ptx::tmem_ld_32x32(taddr + 32, raw[1]);
          ptx::tmem_ld_wait();
          uint4 outv[8];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 sc = ptx::lds128f(P + c * 32 + 4 * i), sh = ptx::lds128f(P + kC + c * 32 + 4 * i);
              v[4 * i] = fmaf(__uint_as_float(raw[c][4 * i]), sc.x, sh.x);
              v[4 * i + 1] = fmaf(__uint_as_float(raw[c][4 * i + 1]), sc.y, sh.y);
              v[4 * i + 2] = fmaf(__uint_as_float(raw[c][4 * i + 2]), sc.z, sh.z);
              v[4 * i + 3] = fmaf(__uint_as_float(raw[c][4 * i + 3]), sc.w, sh.w);
            }
            if (a.residual && l > 0) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 q = ptx::lds128(prow + (((c * 4 + j) ^ x7) << 4));
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const __nv_bfloat162 hh = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
                  v[8 * j + 2 * k] += __low2float(hh);
                  v[8 * j + 2 * k + 1] += __high2float(hh);
                }
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              outv[c * 4 + j] = make_uint4(
                  pack2(fmaxf(v[8 * j], 0.f), fmaxf(v[8 * j + 1], 0.f)), pack2(fmaxf(v[8 * j + 2], 0.f), fmaxf(v[8 * j + 3], 0.f)),
                  pack2(fmaxf(v[8 * j + 4], 0.f), fmaxf(v[8 * j + 5], 0.f)), pack2(fmaxf(v[8 * j + 6], 0.f), fmaxf(v[8 * j + 7], 0.f)));
          }
          if (!last) {
            if (valid) {
#pragma unroll
              for (int j = 0; j < 8; ++j) ptx::sts128(prow + ((j ^ x7) << 4), outv[j]);
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&plane_ready[t]);
          } else if (valid) {
            uint4* dst = reinterpret_cast<uint4*>(a.out + ((size_t)seq * L + pos) * kC);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j] = outv[j];
          }
        }
      }
      // the next item's stem overwrites the planes: this team's tiles have been read by their last
      // MMAs (acc_full waited above) and their TMEM accumulators are drained
      ptx::tc_fence_before();
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTiles * kC);
  }
}

}  // namespace cgf
}  // namespace svdd
