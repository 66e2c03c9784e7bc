// Stage 3a: the RNA value net / reward oracle (reference: ConvGRUTrunk.forward
// Enformer.py:1411-1426 + ConvHead.forward :2166-2173), scoring N token rows in one call.
//
//   tokens -> Conv(4->C,k15)+ReLU                       cg_embed_kernel (one-hot conv = weight gather;
//                                                       mask token = all-zero row, diffusion_gosai.py:1462-1470)
//          -> (n_conv-1) x [Conv(C->C,k5) -> BN -> +res -> ReLU]   conv_gemm (tcgen05), BN folded to scale/shift
//          -> biGRU(C), fwd+bwd summed                  conv_gemm for the input projections of both directions,
//                                                       cg_gru_kernel for the 2 x L-step recurrence (fp32)
//          -> LN -> Linear(C,2C) -> ReLU -> Linear(2C,C) -> 1x1(C->1) -> mean over L
//                                                       cg_ln_kernel + conv_gemm(EPI_HEADDOT) + cg_mean_kernel;
//                                                       Linear(2C,C) and the head are both linear, so they are
//                                                       pre-multiplied into one 2C-vector at pack time.
// C must be 64 (the reference's configuration).  Rows are processed in chunks so the
// workspace stays bounded for B*M in the hundreds of thousands.
#include <cstdlib>
#include <new>

#include "cg_fused.cuh"
#include "conv_gemm.cuh"
#include "gru_umma.cuh"
#include "weights.cuh"

namespace svdd {
namespace {

constexpr int kC = 64;
constexpr int kG3 = 3 * kC;        // gates per direction (r, z, n)
constexpr int kStemTapsMax = 15;
constexpr int kMaxBlocks = 16;
constexpr int64_t kChunkRows = 16384;

// ---- stem: one warp per 8 consecutive positions, lane -> 2 channels -----------------------
// The one-hot conv is a gather-add of weight rows.  The warp loads the 8 + taps - 1 tokens its
// positions see once (one per lane) and walks the taps with the 8 positions' accumulators as
// independent chains (the first version did one position at a time, a dependent chain of `taps`
// shuffle + shared-load + add per position: 164 us per 12 800 x 50 rows against 14 us of HBM time).
constexpr int kEmbedPos = 8;
template <typename Tok>
__global__ void __launch_bounds__(256)
cg_embed_kernel(const Tok* __restrict__ tokens, const float* __restrict__ w /*[taps][4][64]*/,
                const float* __restrict__ b, __nv_bfloat16* __restrict__ out, int64_t NL, int L,
                int taps) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(8) float s_w[kStemTapsMax * 4 * kC];
  for (int i = threadIdx.x; i < taps * 4 * kC; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane * 2;
  const float2 bias = *reinterpret_cast<const float2*>(b + c);
  const int half = taps / 2;
  // grid-stride over groups of 8 positions: the 15 KB weight table is staged once per (persistent) block
  for (int64_t grp = (int64_t)blockIdx.x * 8 + warp; grp * kEmbedPos < NL; grp += (int64_t)gridDim.x * 8) {
  const int64_t pos_base = grp * kEmbedPos;
  // lane j holds the token at global position pos_base - half + j (j < kEmbedPos + taps - 1 <= 22);
  // -1 outside the tensor and for the mask token (all-zero one-hot row)
  int tokv = -1;
  {
    const int64_t gp = pos_base - half + lane;
    if (lane < kEmbedPos + taps - 1 && gp >= 0 && gp < NL) {
      const int t = load_tok(tokens, (size_t)gp);
      tokv = (t < 4) ? t : -1;
    }
  }
  const int l0 = (int)(pos_base % L);
  float2 acc[kEmbedPos];
#pragma unroll
  for (int i = 0; i < kEmbedPos; ++i) acc[i] = bias;
  for (int t = 0; t < taps; ++t) {
#pragma unroll
    for (int i = 0; i < kEmbedPos; ++i) {
      const int tk = __shfl_sync(0xffffffffu, tokv, i + t);
      // position i of the warp sits at l = (l0 + i) mod L of its sequence; tap t reads l + t - half,
      // which must stay inside that sequence (zero padding at its ends)
      int li = l0 + i;
      if (li >= L) li -= L;                      // kEmbedPos <= L is checked by the launcher
      const int lt = li + t - half;
      if (tk >= 0 && lt >= 0 && lt < L) {
        const float2 ww = *reinterpret_cast<const float2*>(&s_w[(t * 4 + tk) * kC + c]);
        acc[i].x += ww.x; acc[i].y += ww.y;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kEmbedPos; ++i) {
    const int64_t pos = pos_base + i;
    if (pos < NL)
      *reinterpret_cast<__nv_bfloat162*>(out + pos * kC + c) =
          __floats2bfloat162_rn(fmaxf(acc[i].x, 0.f), fmaxf(acc[i].y, 0.f));
  }
  }
}

// stem conv weight [64,4,taps] -> [taps][4][64]
__global__ void cg_pack_stem_kernel(const float* __restrict__ w, float* __restrict__ out, int taps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= taps * 4 * kC) return;
  const int c = i % kC, tok = (i / kC) % 4, t = i / (kC * 4);
  out[i] = w[(c * 4 + tok) * taps + t];
}

// ---- GRU recurrence ------------------------------------------------------------------
// gi  fp32 [rows*L, 384]: input projections of both directions (fwd gates 0..191, bwd 192..383),
//     b_ih folded in, plus b_hh for the r and z gates.
// One block = kSeq sequences x one direction; thread j owns gate row j of W_hh (64 registers).
// torch.nn.GRU gate order (r, z, n):  r = s(gi_r + W_hr h),  z = s(gi_z + W_hz h),
//   n = tanh(gi_n + r * (W_hn h + b_hn)),  h' = (1 - z) * n + z * h.
constexpr int kSeq = 8;

__global__ void __launch_bounds__(kG3)
cg_gru_kernel(const float* __restrict__ gi, const float* __restrict__ whh /*[2][192][64]*/,
              const float* __restrict__ bhn /*[2][64]*/, float* __restrict__ y /*[2][rows*L][64]*/,
              int64_t rows, int L) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float s_h[kSeq][kC];
  __shared__ float s_g[kSeq][kG3];
  const int dir = blockIdx.y;
  const int j = threadIdx.x;
  const int64_t seq0 = (int64_t)blockIdx.x * kSeq;
  float w[kC];
  {
    const float4* wr = reinterpret_cast<const float4*>(whh + ((size_t)dir * kG3 + j) * kC);
#pragma unroll
    for (int k = 0; k < kC / 4; ++k) {
      const float4 f = wr[k];
      w[4 * k] = f.x; w[4 * k + 1] = f.y; w[4 * k + 2] = f.z; w[4 * k + 3] = f.w;
    }
  }
  const float b_n = (j >= 2 * kC) ? bhn[dir * kC + (j - 2 * kC)] : 0.0f;
  for (int i = j; i < kSeq * kC; i += kG3) (&s_h[0][0])[i] = 0.0f;
  __syncthreads();
  float* ydir = y + (size_t)dir * rows * L * kC;
  for (int step = 0; step < L; ++step) {
    const int t = dir == 0 ? step : L - 1 - step;
    float acc[kSeq];
#pragma unroll
    for (int s = 0; s < kSeq; ++s) acc[s] = b_n;
#pragma unroll
    for (int k = 0; k < kC; k += 4) {
#pragma unroll
      for (int s = 0; s < kSeq; ++s) {
        const float4 hv = *reinterpret_cast<const float4*>(&s_h[s][k]);
        acc[s] += w[k] * hv.x + w[k + 1] * hv.y + w[k + 2] * hv.z + w[k + 3] * hv.w;
      }
    }
#pragma unroll
    for (int s = 0; s < kSeq; ++s) {
      const int64_t seq = seq0 + s;
      float v = acc[s];
      if (j < 2 * kC) {
        const float g = (seq < rows) ? __ldg(gi + ((size_t)seq * L + t) * (2 * kG3) + dir * kG3 + j) : 0.0f;
        v = 1.0f / (1.0f + __expf(-(g + v)));
      }
      s_g[s][j] = v;
    }
    __syncthreads();
    for (int i = j; i < kSeq * kC; i += kG3) {
      const int s = i / kC, u = i % kC;
      const int64_t seq = seq0 + s;
      if (seq < rows) {
        const float gn = __ldg(gi + ((size_t)seq * L + t) * (2 * kG3) + dir * kG3 + 2 * kC + u);
        const float r = s_g[s][u], z = s_g[s][kC + u];
        const float n = tanhf(gn + r * s_g[s][2 * kC + u]);
        const float hn = (1.0f - z) * n + z * s_h[s][u];
        s_h[s][u] = hn;
        ydir[((size_t)seq * L + t) * kC + u] = hn;
      }
    }
    __syncthreads();
  }
}

// ---- GRU recurrence on the warp-level tensor cores -------------------------------------
// One block (4 warps) = kSeqT sequences x one direction.  Per step the recurrent product
// h[16 x 64] . W_hh^T[64 x 192] is 6 n-tiles x 4 k-steps of mma.sync.m16n8k16 per warp (warp w
// owns hidden units 16w .. 16w+15 of all three gates, so the gate arithmetic is thread-local
// on the accumulator fragments).  fp32 accuracy is kept with a bf16 hi/lo split of both
// operands (hi.hi + lo.hi + hi.lo, fp32 accumulate: the dropped lo.lo term is 2^-18 relative).
// W_hh fragments live in registers for the whole sequence; h is double-buffered in shared
// memory as bf16 hi/lo planes (one barrier per step); the step's input projections gi are
// prefetched one step ahead, so no global-load latency sits on the recurrence.
// Same result per row whatever the batch composition (rows of an MMA are independent).
// The scalar kernel above did the product on the FMA pipe with h broadcast from shared
// memory: 3.99 ms per 16 384 x 50 chunk, against 0.6 ms of fp32 FMA work.
constexpr int kSeqT = 16;
constexpr int kHPitch = kC + 8;      // bf16 per row of an h plane: conflict-free fragment loads

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) -> bf16x2 of the values and bf16x2 of what the rounding dropped
__device__ __forceinline__ void split_bf16x2(float x, float y, uint32_t* hi, uint32_t* lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - __low2float(h), y - __high2float(h));
  *hi = *reinterpret_cast<const uint32_t*>(&h);
  *lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(128)
cg_gru_mma_kernel(const float* __restrict__ gi, const float* __restrict__ whh /*[2][192][64]*/,
                  const float* __restrict__ bhn /*[2][64]*/, float* __restrict__ y /*[2][rows*L][64]*/,
                  int64_t rows, int L) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) __nv_bfloat16 s_hi[2][kSeqT][kHPitch];
  __shared__ __align__(16) __nv_bfloat16 s_lo[2][kSeqT][kHPitch];
  const int dir = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  const int64_t seq0 = (int64_t)blockIdx.x * kSeqT;

  // B fragments of W_hh^T: tile = gate * 2 + nt covers gate columns gate*64 + 16*warp + 8*nt + (0..7)
  uint32_t wh[6][4][2], wl[6][4][2];
#pragma unroll
  for (int tile = 0; tile < 6; ++tile) {
    const int col = (tile >> 1) * kC + 16 * warp + 8 * (tile & 1) + g;
    const float* wr = whh + ((size_t)dir * kG3 + col) * kC;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float2 f = *reinterpret_cast<const float2*>(wr + kk * 16 + 8 * half + 2 * q);
        split_bf16x2(f.x, f.y, &wh[tile][kk][half], &wl[tile][kk][half]);
      }
    }
  }
  float b_n[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const float2 f = *reinterpret_cast<const float2*>(bhn + dir * kC + 16 * warp + 8 * nt + 2 * q);
    b_n[nt][0] = f.x; b_n[nt][1] = f.y;
  }
  for (int i = threadIdx.x; i < 2 * kSeqT * kHPitch; i += blockDim.x) {
    (&s_hi[0][0][0])[i] = __float2bfloat16(0.0f);
    (&s_lo[0][0][0])[i] = __float2bfloat16(0.0f);
  }
  __syncthreads();

  // this thread's 8 (sequence, unit pair) cells: sequence rows g and g + 8, units u0(nt) and u0(nt) + 1
  const bool live[2] = {seq0 + g < rows, seq0 + g + 8 < rows};
  const int u0[2] = {16 * warp + 2 * q, 16 * warp + 8 + 2 * q};
  float h[2][2][2] = {};                     // [row half][nt][pair element]
  float2 gq[3][2][2];                        // prefetched input projections [gate][row half][nt]
  auto load_gi = [&](int step) {
    const int t = dir == 0 ? step : L - 1 - step;
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const float* base = gi + ((size_t)(seq0 + g + 8 * rh) * L + t) * (2 * kG3) + dir * kG3;
#pragma unroll
      for (int gate = 0; gate < 3; ++gate)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          gq[gate][rh][nt] = live[rh] ? __ldg(reinterpret_cast<const float2*>(base + gate * kC + u0[nt]))
                                      : make_float2(0.0f, 0.0f);
    }
  };
  load_gi(0);
  float* ydir = y + (size_t)dir * rows * L * kC;

  for (int step = 0; step < L; ++step) {
    const int t = dir == 0 ? step : L - 1 - step;
    const int cur = step & 1;
    float acc[6][4];
#pragma unroll
    for (int tile = 0; tile < 6; ++tile)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[tile][c] = 0.0f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ah[4], al[4];
      const int k0 = kk * 16 + 2 * q;
      ah[0] = *reinterpret_cast<const uint32_t*>(&s_hi[cur][g][k0]);
      ah[1] = *reinterpret_cast<const uint32_t*>(&s_hi[cur][g + 8][k0]);
      ah[2] = *reinterpret_cast<const uint32_t*>(&s_hi[cur][g][k0 + 8]);
      ah[3] = *reinterpret_cast<const uint32_t*>(&s_hi[cur][g + 8][k0 + 8]);
      al[0] = *reinterpret_cast<const uint32_t*>(&s_lo[cur][g][k0]);
      al[1] = *reinterpret_cast<const uint32_t*>(&s_lo[cur][g + 8][k0]);
      al[2] = *reinterpret_cast<const uint32_t*>(&s_lo[cur][g][k0 + 8]);
      al[3] = *reinterpret_cast<const uint32_t*>(&s_lo[cur][g + 8][k0 + 8]);
#pragma unroll
      for (int tile = 0; tile < 6; ++tile) {
        mma_bf16_16816(acc[tile], al, wh[tile][kk][0], wh[tile][kk][1]);
        mma_bf16_16816(acc[tile], ah, wl[tile][kk][0], wl[tile][kk][1]);
        mma_bf16_16816(acc[tile], ah, wh[tile][kk][0], wh[tile][kk][1]);
      }
    }
    // gates of this step from the prefetched projections, then fetch the next step's
    float2 cr[2][2], cz[2][2], cn[2][2];
#pragma unroll
    for (int rh = 0; rh < 2; ++rh)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) { cr[rh][nt] = gq[0][rh][nt]; cz[rh][nt] = gq[1][rh][nt]; cn[rh][nt] = gq[2][rh][nt]; }
    if (step + 1 < L) load_gi(step + 1);
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float hn[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c = 2 * rh + e;
          const float gr = e ? cr[rh][nt].y : cr[rh][nt].x;
          const float gz = e ? cz[rh][nt].y : cz[rh][nt].x;
          const float gn = e ? cn[rh][nt].y : cn[rh][nt].x;
          const float r = 1.0f / (1.0f + __expf(-(gr + acc[nt][c])));
          const float z = 1.0f / (1.0f + __expf(-(gz + acc[2 + nt][c])));
          const float n = tanhf(gn + r * (acc[4 + nt][c] + b_n[nt][e]));
          hn[e] = (1.0f - z) * n + z * h[rh][nt][e];
          h[rh][nt][e] = hn[e];
        }
        uint32_t hi, lo;
        split_bf16x2(hn[0], hn[1], &hi, &lo);
        *reinterpret_cast<uint32_t*>(&s_hi[cur ^ 1][g + 8 * rh][u0[nt]]) = hi;
        *reinterpret_cast<uint32_t*>(&s_lo[cur ^ 1][g + 8 * rh][u0[nt]]) = lo;
        if (live[rh])
          *reinterpret_cast<float2*>(ydir + ((size_t)(seq0 + g + 8 * rh) * L + t) * kC + u0[nt]) =
              make_float2(hn[0], hn[1]);
      }
    }
    __syncthreads();
  }
}

// ---- (y_fwd + y_bwd) -> LayerNorm(64) -> bf16 : one warp per position -----------------
__global__ void __launch_bounds__(256)
cg_ln_kernel(const float* __restrict__ y, const float* __restrict__ g, const float* __restrict__ b,
             __nv_bfloat16* __restrict__ out, int64_t NL) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pos = (int64_t)blockIdx.x * 8 + warp;
  if (pos >= NL) return;
  const int c = lane * 2;
  const float2 a = *reinterpret_cast<const float2*>(y + pos * kC + c);
  const float2 d = *reinterpret_cast<const float2*>(y + (NL + pos) * kC + c);
  const float v0 = a.x + d.x, v1 = a.y + d.y;   // Enformer.py:1620
  float sum = v0 + v1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * (1.0f / kC);
  const float e0 = v0 - mean, e1 = v1 - mean;
  float sq = e0 * e0 + e1 * e1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * (1.0f / kC) + 1e-5f);
  *reinterpret_cast<__nv_bfloat162*>(out + pos * kC + c) =
      __floats2bfloat162_rn(e0 * rstd * g[c] + b[c], e1 * rstd * g[c + 1] + b[c + 1]);
}

// score[n] = const + mean_l sum_tiles partials[(n*L+l), tile]
__global__ void cg_mean_kernel(const float* __restrict__ partials, int n_tiles, float bias_const,
                               float* __restrict__ scores, int64_t rows, int L) {
  pdl_wait();
  pdl_trigger();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows) return;
  float s = 0.0f;
  const float* p = partials + (size_t)n * L * n_tiles;
  for (int i = 0; i < L * n_tiles; ++i) s += p[i];
  scores[n] = s / (float)L + bias_const;
}

// v[k] = sum_c head_w[c] * W2[c][k]   (k < 2C);  consts[0] = head_w . b2 + head_b
__global__ void cg_fold_head_kernel(const float* __restrict__ w2 /*[C][2C]*/, const float* __restrict__ b2,
                                    const float* __restrict__ hw /*[C]*/, const float* __restrict__ hb,
                                    float* __restrict__ v, float* __restrict__ consts) {
  const int k = threadIdx.x;
  if (k < 2 * kC) {
    float s = 0.0f;
    for (int c = 0; c < kC; ++c) s += hw[c] * w2[c * 2 * kC + k];
    v[k] = s;
  }
  if (k == 0) {
    float s = hb[0];
    for (int c = 0; c < kC; ++c) s += hw[c] * b2[c];
    consts[0] = s;
  }
}

// gate bias of the input projection: b_ih (+ b_hh for r,z) for both directions -> [384]
__global__ void cg_fold_gate_bias_kernel(const float* __restrict__ bih_f, const float* __restrict__ bhh_f,
                                         const float* __restrict__ bih_b, const float* __restrict__ bhh_b,
                                         float* __restrict__ gate_b, float* __restrict__ bhn) {
  const int i = threadIdx.x;  // 0..383
  const int dir = i / kG3, j = i % kG3;
  const float* bih = dir ? bih_b : bih_f;
  const float* bhh = dir ? bhh_b : bhh_f;
  gate_b[i] = bih[j] + (j < 2 * kC ? bhh[j] : 0.0f);
  if (j >= 2 * kC) bhn[dir * kC + j - 2 * kC] = bhh[j];
}

// packed stem weight [taps][4][64] fp32 -> bf16 hi / lo tiles [64 out][64 K], K = 4 * tap + token (K >= 4 * taps: zero)
__global__ void cg_pack_stem_hilo_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                         __nv_bfloat16* __restrict__ lo, int taps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kC * kC) return;
  const int c = i / kC, k = i % kC;
  const float v = k < 4 * taps ? w[(size_t)k * kC + c] : 0.0f;      // w[(tap * 4 + tok) * 64 + c]
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

__global__ void cg_fill_kernel(float* __restrict__ p, float v, int n) {
  if ((int)threadIdx.x < n) p[threadIdx.x] = v;
}

// W_hh fp32 [n] -> bf16 hi / lo planes (hi = rn(w), lo = rn(w - hi)) for the tcgen05 recurrence
__global__ void cg_split_hi_lo_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const __nv_bfloat16 h = __float2bfloat16_rn(w[i]);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(w[i] - __bfloat162float(h));
}

}  // namespace
}  // namespace svdd

using namespace svdd;

struct svdd_convgru {
  int n_blocks = 0;       // conv blocks after the stem
  int stem_taps = 15, taps = 5;
  bool has_bn = false, residual = false;
  DeviceArena arena;
  float* stem_w = nullptr;
  float* stem_b = nullptr;
  __nv_bfloat16* conv_w[kMaxBlocks] = {};
  float* conv_scale[kMaxBlocks] = {};
  float* conv_shift[kMaxBlocks] = {};   // (or plain bias when there is no BN)
  __nv_bfloat16* conv_w_all = nullptr;   // [1 + n_blocks][taps][64][64]: layer 0 = stem as (W hi, W lo) over the 60 one-hot features, then the conv weights, contiguous
  float* conv_ss_all = nullptr;          // [1 + n_blocks][2][64]: scale (1 for the stem / without BN), shift
  __nv_bfloat16* wih = nullptr;          // [384][64]
  float* gate_b = nullptr;               // [384]
  float* whh = nullptr;                  // [2][192][64]
  __nv_bfloat16* whh_hi = nullptr;       // [2][192][64]  bf16 hi / lo split of W_hh (tcgen05 recurrence)
  __nv_bfloat16* whh_lo = nullptr;
  float* bhn = nullptr;                  // [2][64]
  float* ln_g = nullptr;
  float* ln_b = nullptr;
  __nv_bfloat16* w1 = nullptr;           // [128][64]
  float* b1 = nullptr;                   // [128]
  float* headv = nullptr;                // [128]
  float* consts = nullptr;               // [1] (device)
  float head_const = 0.0f;               // host copy
};

extern "C" int svdd_convgru_create(const svdd_tensor* tensors, int n_tensors, void* stream,
                                   svdd_convgru** out) {
  SVDD_CHECK_ARG(tensors && out && n_tensors > 0, "svdd_convgru_create: null argument");
  int dev = 0;
  SVDD_CUDA(cudaGetDevice(&dev));
  SVDD_TRY(svdd_device_check(dev));
  cudaStream_t st = (cudaStream_t)stream;
  TensorTable tt{tensors, n_tensors};
  const std::string stem = "conv_tower.blocks.0.conv.";
  SVDD_CHECK_ARG(tt.dim(stem + "weight", 0) == kC && tt.dim(stem + "weight", 1) == 4,
                 "svdd_convgru_create: stem must be Conv1d(4 -> 64) (got %lld -> %lld)",
                 (long long)tt.dim(stem + "weight", 1), (long long)tt.dim(stem + "weight", 0));
  svdd_convgru* h = new (std::nothrow) svdd_convgru();
  SVDD_CHECK_ARG(h != nullptr, "out of host memory");
  auto fail = [&](int code) { delete h; return code; };
  h->stem_taps = (int)tt.dim(stem + "weight", 2);
  if (h->stem_taps < 1 || h->stem_taps > kStemTapsMax || h->stem_taps % 2 == 0) {
    set_last_error("svdd_convgru_create: unsupported stem kernel size %d", h->stem_taps);
    return fail(SVDD_ERR_INVALID_ARGUMENT);
  }
  int nb = 0;
  while (nb + 1 < kMaxBlocks && tt.has("conv_tower.blocks." + std::to_string(nb + 1) + ".conv.weight")) ++nb;
  h->n_blocks = nb;
  h->has_bn = nb > 0 && tt.has("conv_tower.blocks.1.norm.layer.running_mean");
  h->residual = h->has_bn;  // the reference sets conv_norm and residual together (Enformer.py:40-44)
  h->taps = nb > 0 ? (int)tt.dim("conv_tower.blocks.1.conv.weight", 2) : 5;

  DeviceArena& A = h->arena;
  A.reserve(sizeof(float) * h->stem_taps * 4 * kC);
  A.reserve(sizeof(float) * kC);
  for (int i = 0; i < nb; ++i) {
    A.reserve(sizeof(__nv_bfloat16) * h->taps * kC * kC);
    A.reserve(sizeof(float) * kC);
    A.reserve(sizeof(float) * kC);
  }
  A.reserve(sizeof(__nv_bfloat16) * (size_t)(nb + 1) * h->taps * kC * kC);
  A.reserve(sizeof(float) * (size_t)(nb + 1) * 2 * kC);
  A.reserve(sizeof(__nv_bfloat16) * 2 * kG3 * kC);
  A.reserve(sizeof(float) * 2 * kG3);
  A.reserve(sizeof(float) * 2 * kG3 * kC);
  A.reserve(sizeof(__nv_bfloat16) * 2 * kG3 * kC);
  A.reserve(sizeof(__nv_bfloat16) * 2 * kG3 * kC);
  A.reserve(sizeof(float) * 2 * kC);
  A.reserve(sizeof(float) * kC);
  A.reserve(sizeof(float) * kC);
  A.reserve(sizeof(__nv_bfloat16) * 2 * kC * kC);
  A.reserve(sizeof(float) * 2 * kC);
  A.reserve(sizeof(float) * 2 * kC);
  A.reserve(sizeof(float) * 8);
  int rc = A.commit();
  if (rc != SVDD_OK) return fail(rc);
  h->stem_w = A.take<float>(h->stem_taps * 4 * kC);
  h->stem_b = A.take<float>(kC);
  for (int i = 0; i < nb; ++i) {
    h->conv_w[i] = A.take<__nv_bfloat16>(h->taps * kC * kC);
    h->conv_scale[i] = A.take<float>(kC);
    h->conv_shift[i] = A.take<float>(kC);
  }
  h->conv_w_all = A.take<__nv_bfloat16>((size_t)(nb + 1) * h->taps * kC * kC);
  h->conv_ss_all = A.take<float>((size_t)(nb + 1) * 2 * kC);
  h->wih = A.take<__nv_bfloat16>(2 * kG3 * kC);
  h->gate_b = A.take<float>(2 * kG3);
  h->whh = A.take<float>(2 * kG3 * kC);
  h->whh_hi = A.take<__nv_bfloat16>(2 * kG3 * kC);
  h->whh_lo = A.take<__nv_bfloat16>(2 * kG3 * kC);
  h->bhn = A.take<float>(2 * kC);
  h->ln_g = A.take<float>(kC);
  h->ln_b = A.take<float>(kC);
  h->w1 = A.take<__nv_bfloat16>(2 * kC * kC);
  h->b1 = A.take<float>(2 * kC);
  h->headv = A.take<float>(2 * kC);
  h->consts = A.take<float>(8);

#define GET_OR_FAIL(var, name, numel)               \
  const float* var = tt.get((name), (numel));       \
  if (var == nullptr) return fail(SVDD_ERR_MISSING_TENSOR)
#define TRY_OR_FAIL(expr)                           \
  do { int _rc = (expr); if (_rc != SVDD_OK) return fail(_rc); } while (0)

  GET_OR_FAIL(sw, stem + "weight", (int64_t)kC * 4 * h->stem_taps);
  GET_OR_FAIL(sb, stem + "bias", kC);
  cg_pack_stem_kernel<<<ceil_div(h->stem_taps * 4 * kC, 256), 256, 0, st>>>(sw, h->stem_w, h->stem_taps);
  TRY_OR_FAIL(copy_f32(sb, h->stem_b, kC, st));
  for (int i = 0; i < nb; ++i) {
    const std::string p = "conv_tower.blocks." + std::to_string(i + 1) + ".";
    GET_OR_FAIL(cw, p + "conv.weight", (int64_t)kC * kC * h->taps);
    GET_OR_FAIL(cb, p + "conv.bias", kC);
    TRY_OR_FAIL(pack_conv_weight(cw, h->conv_w[i], kC, kC, h->taps, st));
    if (h->has_bn) {
      GET_OR_FAIL(g, p + "norm.layer.weight", kC);
      GET_OR_FAIL(b, p + "norm.layer.bias", kC);
      GET_OR_FAIL(m, p + "norm.layer.running_mean", kC);
      GET_OR_FAIL(v, p + "norm.layer.running_var", kC);
      TRY_OR_FAIL(fold_bn(g, b, m, v, cb, 1e-5f, h->conv_scale[i], h->conv_shift[i], kC, st));
    } else {
      TRY_OR_FAIL(copy_f32(cb, h->conv_shift[i], kC, st));
    }
  }
  if (h->taps >= 2 && h->stem_taps * 4 <= kC) {
    // layer 0 of the fused conv-stack kernel: the stem over one-hot im2col features, bf16 hi / lo tiles
    if (cudaMemsetAsync(h->conv_w_all, 0, sizeof(__nv_bfloat16) * h->taps * kC * kC, st) != cudaSuccess) return fail(SVDD_ERR_CUDA);
    cg_pack_stem_hilo_kernel<<<ceil_div(kC * kC, 256), 256, 0, st>>>(h->stem_w, h->conv_w_all, h->conv_w_all + kC * kC, h->stem_taps);
    cg_fill_kernel<<<1, kC, 0, st>>>(h->conv_ss_all, 1.0f, kC);
    if (cudaMemcpyAsync(h->conv_ss_all + kC, h->stem_b, sizeof(float) * kC, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return fail(SVDD_ERR_CUDA);
  }
  for (int i = 0; i < nb; ++i) {
    __nv_bfloat16* wdst = h->conv_w_all + (size_t)(i + 1) * h->taps * kC * kC;
    float* sdst = h->conv_ss_all + (size_t)(i + 1) * 2 * kC;
    if (cudaMemcpyAsync(wdst, h->conv_w[i], sizeof(__nv_bfloat16) * h->taps * kC * kC, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return fail(SVDD_ERR_CUDA);
    if (h->has_bn) {
      if (cudaMemcpyAsync(sdst, h->conv_scale[i], sizeof(float) * kC, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
        return fail(SVDD_ERR_CUDA);
    } else {
      cg_fill_kernel<<<1, kC, 0, st>>>(sdst, 1.0f, kC);
    }
    if (cudaMemcpyAsync(sdst + kC, h->conv_shift[i], sizeof(float) * kC, cudaMemcpyDeviceToDevice, st) != cudaSuccess)
      return fail(SVDD_ERR_CUDA);
  }
  const std::string gp = "gru_tower.gru.";
  GET_OR_FAIL(wih_f, gp + "weight_ih_l0", (int64_t)kG3 * kC);
  GET_OR_FAIL(wih_b, gp + "weight_ih_l0_reverse", (int64_t)kG3 * kC);
  GET_OR_FAIL(whh_f, gp + "weight_hh_l0", (int64_t)kG3 * kC);
  GET_OR_FAIL(whh_b, gp + "weight_hh_l0_reverse", (int64_t)kG3 * kC);
  GET_OR_FAIL(bih_f, gp + "bias_ih_l0", kG3);
  GET_OR_FAIL(bih_b, gp + "bias_ih_l0_reverse", kG3);
  GET_OR_FAIL(bhh_f, gp + "bias_hh_l0", kG3);
  GET_OR_FAIL(bhh_b, gp + "bias_hh_l0_reverse", kG3);
  TRY_OR_FAIL(pack_conv_weight(wih_f, h->wih, kG3, kC, 1, st));
  TRY_OR_FAIL(pack_conv_weight(wih_b, h->wih + kG3 * kC, kG3, kC, 1, st));
  TRY_OR_FAIL(copy_f32(whh_f, h->whh, kG3 * kC, st));
  TRY_OR_FAIL(copy_f32(whh_b, h->whh + kG3 * kC, kG3 * kC, st));
  cg_fold_gate_bias_kernel<<<1, 2 * kG3, 0, st>>>(bih_f, bhh_f, bih_b, bhh_b, h->gate_b, h->bhn);
  cg_split_hi_lo_kernel<<<ceil_div(2 * kG3 * kC, 256), 256, 0, st>>>(h->whh, h->whh_hi, h->whh_lo, 2 * kG3 * kC);
  const std::string fp = "gru_tower.ffn.";
  GET_OR_FAIL(lg, fp + "dense1.norm.layer.weight", kC);
  GET_OR_FAIL(lb, fp + "dense1.norm.layer.bias", kC);
  GET_OR_FAIL(w1, fp + "dense1.linear.weight", 2 * kC * kC);
  GET_OR_FAIL(b1, fp + "dense1.linear.bias", 2 * kC);
  GET_OR_FAIL(w2, fp + "dense2.linear.weight", 2 * kC * kC);
  GET_OR_FAIL(b2, fp + "dense2.linear.bias", kC);
  GET_OR_FAIL(hw, "head.channel_transform.conv.layer.weight", kC);
  GET_OR_FAIL(hb, "head.channel_transform.conv.layer.bias", 1);
  TRY_OR_FAIL(copy_f32(lg, h->ln_g, kC, st));
  TRY_OR_FAIL(copy_f32(lb, h->ln_b, kC, st));
  TRY_OR_FAIL(pack_conv_weight(w1, h->w1, 2 * kC, kC, 1, st));
  TRY_OR_FAIL(copy_f32(b1, h->b1, 2 * kC, st));
  cg_fold_head_kernel<<<1, 2 * kC, 0, st>>>(w2, b2, hw, hb, h->headv, h->consts);
#undef GET_OR_FAIL
#undef TRY_OR_FAIL
  if (cudaMemcpyAsync(&h->head_const, h->consts, sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) {
    set_last_error("svdd_convgru_create: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(SVDD_ERR_CUDA);
  }
  *out = h;
  return SVDD_OK;
}

extern "C" void svdd_convgru_destroy(svdd_convgru* h) { delete h; }

namespace {
struct CgWs {
  __nv_bfloat16* x[2];
  float* gi;
  float* y;
  __nv_bfloat16* z;
  float* partials;
};
bool gru_scalar() {
  static const bool v = [] { const char* e = getenv("SVDD_GRU_SCALAR"); return e && e[0] == '1'; }();
  return v;
}
// The tcgen05 recurrence (256 sequences per CTA, the step a serial chain of ~3 us) wins once there
// are enough sequences to fill the SMs with such CTAs: measured 4.29 vs 6.56 ms at 51 200 rows,
// 0.92 vs 1.22 ms at 8 192, equal at 4 096, 0.61 vs 0.45 ms at 2 048 (the mma.sync kernel spreads 16
// sequences per block over many more blocks).  SVDD_GRU_UMMA (read per call): 0 = never, 1 = always
// (the tests), unset = from kGruUmmaMinRows rows.
constexpr int64_t kGruUmmaMinRows = 6144;
constexpr int64_t kChunkRowsUmma = 65536;   // no gi tensor on this path: 904 instead of 2440 bytes per position
bool gru_umma(int64_t rows) {
  if (gru_scalar()) return false;
  const char* e = getenv("SVDD_GRU_UMMA");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;
  return rows >= kGruUmmaMinRows;
}
// SVDD_CG_FUSED=0 (read per call): stem + conv blocks as separate launches (one-hot stem kernel + one
// first-generation implicit GEMM per block) instead of the persistent conv-stack kernel
bool cg_fused_usable(const svdd_convgru* h, int L) {
  const char* e = getenv("SVDD_CG_FUSED");
  if (e && e[0] == '0') return false;
  return h->n_blocks >= 1 && h->n_blocks <= cgf::kMaxLayers && h->taps % 2 == 1 && h->taps >= 2 && h->taps <= cgf::kMaxTaps &&
         h->stem_taps * 4 <= kC && h->stem_taps <= 16 && L + h->taps / 2 <= 64 && h->taps / 2 <= cgf::kPad;
}
int launch_convstack(const svdd_convgru* h, const void* tok, int tok_dtype, __nv_bfloat16* out, int64_t rows, int L,
                     cudaStream_t st) {
  CUtensorMap tW;
  SVDD_TRY(encode_tmap_2d_bf16(&tW, h->conv_w_all, kC, (uint64_t)(h->n_blocks + 1) * h->taps * kC, kC, kC));
  cgf::Args a = {};
  a.tokens = tok; a.ss = h->conv_ss_all; a.out = out;
  a.rows = rows; a.L = L; a.n_layers = h->n_blocks; a.taps = h->taps; a.stem_taps = h->stem_taps;
  a.residual = h->residual ? 1 : 0;
  const int64_t items = ceil_div<int64_t>(rows, cgf::kSeqPerItem);
  const unsigned grid = (unsigned)(items < num_sms() ? items : num_sms());
  auto launch = [&](auto kern) -> int {
    SVDD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cgf::kSmemBytes));
    SVDD_CUDA(launch_k(kern, dim3(grid), dim3(cgf::kThreads), (size_t)cgf::kSmemBytes, st, 1, tW, a));
    return SVDD_OK;
  };
  if (tok_dtype == SVDD_TOK_I64) SVDD_TRY(launch(cgf::cg_convstack_kernel<int64_t>));
  else SVDD_TRY(launch(cgf::cg_convstack_kernel<uint8_t>));
  count_launch();
  return SVDD_OK;
}
int launch_gru_umma(const svdd_convgru* h, const __nv_bfloat16* x, float* y, int64_t rows, int L, cudaStream_t st) {
  CUtensorMap tX, tWx, tWhi, tWlo;
  SVDD_TRY(encode_tmap_3d_bf16(&tX, x, kC, (uint64_t)L, (uint64_t)rows, (uint64_t)kC * 2, (uint64_t)L * kC * 2, kC, 1,
                               gruu::kRowsG));
  SVDD_TRY(encode_tmap_2d_bf16(&tWx, h->wih, kC, 2 * kG3, kC, kG3));
  SVDD_TRY(encode_tmap_2d_bf16(&tWhi, h->whh_hi, kC, 2 * kG3, kC, kG3));
  SVDD_TRY(encode_tmap_2d_bf16(&tWlo, h->whh_lo, kC, 2 * kG3, kC, kG3));
  static bool configured = false;
  if (!configured) {
    SVDD_CUDA(cudaFuncSetAttribute(gruu::cg_gru_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gruu::kSmemBytes));
    configured = true;
  }
  const unsigned grid = (unsigned)ceil_div<int64_t>(rows, gruu::kGroups * gruu::kRowsG);
  SVDD_CUDA(launch_k(gruu::cg_gru_umma_kernel, dim3(grid, 2), dim3(gruu::kThreads), (size_t)gruu::kSmemBytes, st, 1, tX, tWx,
                     tWhi, tWlo, (const float*)h->gate_b, (const float*)h->bhn, y, rows, L));
  count_launch();
  return SVDD_OK;
}
size_t cg_carve(Workspace& W, int64_t rows, int L, CgWs* o, bool with_gi = true) {
  const size_t nl = (size_t)rows * L + 1;
  o->x[0] = W.take<__nv_bfloat16>(nl * kC);
  o->x[1] = W.take<__nv_bfloat16>(nl * kC);
  o->gi = with_gi ? W.take<float>(nl * 2 * kG3) : nullptr;
  o->y = W.take<float>(2 * nl * kC);
  o->z = W.take<__nv_bfloat16>(nl * kC);
  o->partials = W.take<float>(nl * 2);
  return W.used();
}
}  // namespace

extern "C" size_t svdd_convgru_workspace_bytes(const svdd_convgru* h, int64_t n_rows, int L) {
  (void)h;
  // enough for either recurrence path: chunks of kChunkRows with the gi tensor, or of kChunkRowsUmma without
  Workspace W(nullptr, 0), W2(nullptr, 0);
  CgWs o;
  const size_t a = cg_carve(W, n_rows < kChunkRows ? n_rows : kChunkRows, L, &o);
  const size_t b = cg_carve(W2, n_rows < kChunkRowsUmma ? n_rows : kChunkRowsUmma, L, &o, false);
  return a > b ? a : b;
}

extern "C" int svdd_convgru_score(svdd_convgru* h, const void* tokens, int tok_dtype, float* scores,
                                  int64_t n_rows, int L, void* ws, size_t ws_bytes, void* stream) {
  SVDD_CHECK_ARG(h && tokens && scores, "svdd_convgru_score: null pointer");
  SVDD_CHECK_ARG(n_rows >= 0 && L >= kEmbedPos, "svdd_convgru_score: bad shape (L >= %d)", kEmbedPos);
  SVDD_CHECK_ARG(tok_dtype == SVDD_TOK_I64 || tok_dtype == SVDD_TOK_U8, "bad tok_dtype %d", tok_dtype);
  if (n_rows == 0) return SVDD_OK;
  if (ws == nullptr || ws_bytes < svdd_convgru_workspace_bytes(h, n_rows, L)) {
    set_last_error("svdd_convgru_score: workspace too small");
    return SVDD_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t tok_bytes = tok_dtype == SVDD_TOK_I64 ? 8 : 1;
  const bool umma = gru_umma(n_rows);
  const int64_t chunk = umma ? kChunkRowsUmma : kChunkRows;
  for (int64_t r0 = 0; r0 < n_rows; r0 += chunk) {
    const int64_t rows = (n_rows - r0 < chunk) ? n_rows - r0 : chunk;
    const int64_t NL = rows * L;
    Workspace W(ws, ws_bytes);
    CgWs b;
    cg_carve(W, rows, L, &b, !umma);
    const void* tok = reinterpret_cast<const uint8_t*>(tokens) + (size_t)r0 * L * tok_bytes;
    int cur = 0;
    if (cg_fused_usable(h, L)) {
      SVDD_TRY(launch_convstack(h, tok, tok_dtype, b.x[0], rows, L, st));
    } else {
    const int64_t eg_all = ceil_div<int64_t>(NL, 64);
    const unsigned eg = (unsigned)(eg_all < 8 * (int64_t)num_sms() ? eg_all : 8 * (int64_t)num_sms());
    if (tok_dtype == SVDD_TOK_I64)
      launch_k(cg_embed_kernel<int64_t>, dim3(eg), dim3(256), 0, st, 1, (const int64_t*)tok, h->stem_w, h->stem_b, b.x[0], NL, L, h->stem_taps);
    else
      launch_k(cg_embed_kernel<uint8_t>, dim3(eg), dim3(256), 0, st, 1, (const uint8_t*)tok, h->stem_w, h->stem_b, b.x[0], NL, L, h->stem_taps);
    count_launch();
    SVDD_LAUNCH_CHECK();
    for (int i = 0; i < h->n_blocks; ++i) {
      GemmShape g;
      g.S = (int)rows; g.L = L; g.L_in = L; g.K = kC; g.N = kC; g.taps = h->taps; g.dil = 1;
      choose_row_tiling(L, h->taps, &g);
      EpiParams ep;
      if (h->has_bn) { ep.scale = h->conv_scale[i]; ep.shift = h->conv_shift[i]; }
      else           { ep.bias = h->conv_shift[i]; }
      if (h->residual) { ep.res = b.x[cur]; ep.res_dtype = DT_BF16; ep.ld_res = kC; }
      ep.act = ACT_RELU; ep.act_after_res = 1;   // order 'CDNRA' (Enformer.py:1399)
      ep.out = b.x[cur ^ 1]; ep.out_dtype = DT_BF16; ep.ld_out = kC;
      SVDD_TRY(launch_conv_gemm(b.x[cur], h->conv_w[i], g, EPI_GENERIC, ep, st));
      cur ^= 1;
    }
    }
    if (umma) {
      // tcgen05 recurrence with the input projection fused in (csrc/gru_umma.cuh): no gi tensor
      SVDD_TRY(launch_gru_umma(h, b.x[cur], b.y, rows, L, st));
    } else {
      {  // GRU input projections, both directions at once: [NL,64] x [64,384]
        GemmShape g;
        g.S = 1; g.L = (int)NL; g.L_in = (int)NL; g.K = kC; g.N = 2 * kG3; g.BL = 128; g.BS = 1;
        EpiParams ep;
        ep.bias = h->gate_b;
        ep.out = b.gi; ep.out_dtype = DT_F32; ep.ld_out = 2 * kG3;
        SVDD_TRY(launch_conv_gemm(b.x[cur], h->wih, g, EPI_GENERIC, ep, st));
      }
      if (gru_scalar())   // SVDD_GRU_SCALAR=1: the FMA-pipe kernel (A/B and cross-check of the tensor-core ones)
        launch_k(cg_gru_kernel, dim3((unsigned)ceil_div<int64_t>(rows, kSeq), 2), dim3(kG3), 0, st, 1, b.gi, h->whh, h->bhn, b.y, rows, L);
      else
        launch_k(cg_gru_mma_kernel, dim3((unsigned)ceil_div<int64_t>(rows, kSeqT), 2), dim3(128), 0, st, 1, b.gi, h->whh, h->bhn, b.y, rows, L);
      count_launch();
      SVDD_LAUNCH_CHECK();
    }
    launch_k(cg_ln_kernel, dim3((unsigned)ceil_div<int64_t>(NL, 8)), dim3(256), 0, st, 1, b.y, h->ln_g, h->ln_b, b.z, NL);
    count_launch();
    SVDD_LAUNCH_CHECK();
    int n_tiles = 1;
    {  // Linear(64 -> 128) + ReLU, dotted with the pre-multiplied (Linear(128->64), head) vector
      GemmShape g;
      g.S = 1; g.L = (int)NL; g.L_in = (int)NL; g.K = kC; g.N = 2 * kC; g.BL = 128; g.BS = 1;
      EpiParams ep;
      ep.bias = h->b1;
      ep.act = ACT_RELU;
      ep.head_w = h->headv;
      ep.partials = b.partials;
      n_tiles = conv_gemm_n_tiles(g, EPI_HEADDOT);
      SVDD_TRY(launch_conv_gemm(b.z, h->w1, g, EPI_HEADDOT, ep, st));
    }
    launch_k(cg_mean_kernel, dim3((unsigned)ceil_div<int64_t>(rows, 128)), dim3(128), 0, st, 1, b.partials, n_tiles, h->head_const,
                                                                           scores + r0, rows, L);
    count_launch();
    SVDD_LAUNCH_CHECK();
  }
  return SVDD_OK;
}
