// Stage 3b: the DNA value net / reward oracle (reference: EnformerTrunk.forward
// Enformer.py:1326-1334 + ConvHead.forward :2166-2173), scoring N token rows per call.
//
// Every dense contraction is a conv_gemm (tcgen05) launch; what the reference runs as
// separate BatchNorm / GELU / residual / attention-pool / LayerNorm-input kernels is folded
// into GEMM epilogues:
//   stem  Conv(4->C0,k15)            im2col of the one-hot (K = 15*4 -> 64) + GEMM; epilogue writes x0 and
//                                    a0 = GELU(BN(x0)) (the next conv's operand, order 'NACDR' :2266-2292)
//   1x1 + residual                   GEMM, epilogue adds bias and the block input
//   AttentionPool(2)                 GEMM over (even, odd) position pairs into two TMEM accumulators;
//                                    epilogue does the pair softmax, the weighted sum and the NEXT block's
//                                    BN+GELU (enformer_pytorch AttentionPool; Enformer.py:2447)
//   Conv(k5)                         implicit GEMM (5 shifted TMA boxes), epilogue writes z and GELU(BN(z))
//   transformer x n (n = L/2^7 = 2 positions): LN kernel, fused QKV GEMM, tiny attention kernel with the
//                                    Enformer relative-position logits (rel_k precomputed), out-proj GEMM +
//                                    residual, LN, FFN GEMMs (+ReLU, +residual)
//   pointwise 1x1 -> GELU -> head    GEMM with EPI_HEADDOT: GELU(acc+b) . head_w per row, then mean over n.
//
// Activations are bf16 channels-last [rows, L_i, f_i]; the transformer residual stream is fp32.
#include <math.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "conv_gemm.cuh"
#include "tower.cuh"
#include "weights.cuh"

namespace svdd {
namespace {

constexpr int kMaxStages = 12;
constexpr int kMaxBlocksT = 32;
constexpr int kStemK = 64;       // 15 taps x 4 bases = 60, padded to one 64-wide K block
constexpr int64_t kChunkRows = 2048;
constexpr int kMaxPos = 8;       // positions entering the transformer (2 for L=200)

// ---- stem im2col: one-hot windows as bf16 rows of 64 ---------------------------------
template <typename Tok>
__global__ void ef_im2col_kernel(const Tok* __restrict__ tokens, __nv_bfloat16* __restrict__ col,
                                 int64_t NL, int L, int taps) {
  pdl_wait();
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NL * 8) return;
  const int64_t pos = i >> 3;
  const int grp = (int)(i & 7);          // 8 bf16 = 2 taps x 4 bases
  const int l = (int)(pos % L);
  uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int tap = grp * 2 + h;
    const int li = l + tap - taps / 2;
    if (tap < taps && li >= 0 && li < L) {
      const int tok = load_tok(tokens, pos + tap - taps / 2);
      if (tok < 4) w[h * 2 + (tok >> 1)] = (tok & 1) ? 0x3F800000u : 0x00003F80u;  // bf16 1.0
    }
  }
  reinterpret_cast<uint4*>(col)[i] = make_uint4(w[0], w[1], w[2], w[3]);
}

// stem weight [C0,4,taps] fp32 -> bf16 [C0][64] with k = tap*4 + base
__global__ void ef_pack_stem_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                    int C0, int taps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C0 * kStemK) return;
  const int k = i % kStemK, c = i / kStemK;
  const int tap = k >> 2, base = k & 3;
  out[i] = __float2bfloat16_rn(tap < taps ? w[(c * 4 + base) * taps + tap] : 0.0f);
}

// ---- row LayerNorm fp32 [R,C] -> bf16 [R,C]; one block per row ---------------------------
constexpr int kLnThreads = 256;
constexpr int kLnMaxPer = 16;   // C <= 4096
__global__ void __launch_bounds__(kLnThreads)
ef_ln_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
             __nv_bfloat16* __restrict__ out, int C) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_red[kLnThreads / 32];
  __shared__ float s_stat;
  const int64_t row = blockIdx.x;
  const float* xr = x + row * C;
  float v[kLnMaxPer];
  float sum = 0.0f;
  int cnt = 0;
  for (int c = threadIdx.x; c < C; c += kLnThreads) { v[cnt] = xr[c]; sum += v[cnt]; ++cnt; }
  auto block_sum = [&](float val) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = val;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.0f;
      for (int i = 0; i < kLnThreads / 32; ++i) t += s_red[i];
      s_stat = t;
    }
    __syncthreads();
    const float r = s_stat;
    __syncthreads();
    return r;
  };
  const float mean = block_sum(sum) / (float)C;
  float sq = 0.0f;
  for (int i = 0; i < cnt; ++i) { const float d = v[i] - mean; sq += d * d; }
  const float rstd = rsqrtf(block_sum(sq) / (float)C + 1e-5f);
  cnt = 0;
  for (int c = threadIdx.x; c < C; c += kLnThreads) {
    out[row * C + c] = __float2bfloat16_rn((v[cnt] - mean) * rstd * g[c] + b[c]);
    ++cnt;
  }
}

// Warp-per-row variant for C % 128 == 0, C <= 128 * kLnWarpVecs: the row lives in registers as
// float4s (coalesced 512 B per warp per load), both statistics are warp shuffles, the bf16
// result goes out as 8-byte stores.  Same arithmetic as ef_ln_kernel (two-pass variance).
constexpr int kLnWarpVecs = tower::kLnVecs;    // C <= 2048
constexpr int kLnWarpsPerBlock = 8;
__global__ void __launch_bounds__(kLnWarpsPerBlock * 32)
ef_ln_warp_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                  __nv_bfloat16* __restrict__ out, int64_t rows, int C) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kLnWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows) return;
  tower::ln_warp_row<false>(x + row * C, g, b, out + row * C, C, lane);
}

int launch_ln(const float* x, const float* g, const float* b, __nv_bfloat16* out, int64_t rows, int C,
              cudaStream_t st) {
  if (C % 128 == 0 && C <= 128 * kLnWarpVecs)
    launch_k(ef_ln_warp_kernel, dim3((unsigned)ceil_div<int64_t>(rows, kLnWarpsPerBlock)), dim3(kLnWarpsPerBlock * 32), 0, st, 1, 
        x, g, b, out, rows, C);
  else
    launch_k(ef_ln_kernel, dim3((unsigned)rows), dim3(kLnThreads), 0, st, 1, x, g, b, out, C);
  count_launch();
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}

// ---- attention over n <= 8 positions (enformer_pytorch Attention.forward) ----------------
// qkv fp32 [rows*n, 2*H*dk + H*dv]; relk fp32 [H][2n-1][dk]; out bf16 [rows*n, H*dv].
//   logits[i,j] = (q_i*scale + rcb).k_j + (q_i*scale + rpb).relk[(j-i)+(n-1)]   (relative_shift)
__global__ void __launch_bounds__(256)
ef_attention_kernel(const float* __restrict__ qkv, const float* __restrict__ rcb,
                    const float* __restrict__ rpb, const float* __restrict__ relk,
                    __nv_bfloat16* __restrict__ out, int n, int H, int dk, int dv) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_attn[8][kMaxPos][kMaxPos];   // H <= 8 per pass
  const int64_t seq = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ld = 2 * H * dk + H * dv;
  const float* base = qkv + seq * n * ld;
  const float scale = rsqrtf((float)dk);
  for (int h0 = 0; h0 < H; h0 += 8) {
    const int h = h0 + warp;
    if (h < H) {
      for (int i = 0; i < n; ++i) {
        float lg[kMaxPos];
        for (int j = 0; j < n; ++j) {
          float acc = 0.0f;
          for (int d = lane; d < dk; d += 32) {
            const float q = base[i * ld + h * dk + d] * scale;
            const float k = base[j * ld + H * dk + h * dk + d];
            const float rk = relk[((size_t)h * (2 * n - 1) + (j - i + n - 1)) * dk + d];
            acc += (q + rcb[h * dk + d]) * k + (q + rpb[h * dk + d]) * rk;
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          lg[j] = acc;
        }
        float mx = lg[0];
        for (int j = 1; j < n; ++j) mx = fmaxf(mx, lg[j]);
        float den = 0.0f;
        for (int j = 0; j < n; ++j) { lg[j] = __expf(lg[j] - mx); den += lg[j]; }
        if (lane == 0)
          for (int j = 0; j < n; ++j) s_attn[warp][i][j] = lg[j] / den;
      }
    }
    __syncthreads();
    const int hcount = (H - h0 < 8) ? H - h0 : 8;
    for (int e = threadIdx.x; e < n * hcount * dv; e += blockDim.x) {
      const int d = e % dv, hh = (e / dv) % hcount, i = e / (dv * hcount);
      float acc = 0.0f;
      for (int j = 0; j < n; ++j)
        acc += s_attn[hh][i][j] * base[j * ld + 2 * H * dk + (h0 + hh) * dv + d];
      out[(seq * n + i) * (size_t)(H * dv) + (h0 + hh) * dv + d] = __float2bfloat16_rn(acc);
    }
    __syncthreads();
  }
}

// Warp-per-(sequence, head) variant for small n (the transformer sees n = 2 positions at
// L = 200): q/k/rel_k slices stay in registers, the n*n logits are warp-reduced, every lane
// holds the softmax and produces dv/32 output channels per position.  One pass over qkv
// (coalesced 128 B segments) instead of a block per sequence looping over heads.
constexpr int kAttnSmallN = 4;     // n <= 4
constexpr int kAttnDkPer = tower::kAttnDkPer;      // dk <= 128
constexpr int kAttnWarps = 8;
template <int N>
__global__ void __launch_bounds__(kAttnWarps * 32)
ef_attention_warp_kernel(const float* __restrict__ qkv, const float* __restrict__ rcb,
                         const float* __restrict__ rpb, const float* __restrict__ relk,
                         __nv_bfloat16* __restrict__ out, int64_t rows, int H, int dk, int dv) {
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * kAttnWarps + (threadIdx.x >> 5);
  if (w >= rows * H) return;
  const int64_t seq = w / H;
  const int h = (int)(w % H);
  const int ld = 2 * H * dk + H * dv;
  tower::attn_warp_task<N, false>(qkv + seq * N * ld, ld, rcb, rpb, relk, out + seq * N * (size_t)(H * dv), h, H, dk, dv,
                                  lane);
}

// Warp-per-sequence variant (all heads of one sequence, the q / k / v rows as float4s with every
// load of a stage in flight; tower::attn_warp_seq): H*dk == 512, H*dv in {1536, 384}, n <= 2.
template <int N, int T, int U>
__global__ void __launch_bounds__(kAttnWarps * 32)
ef_attention_seq_kernel(const float* __restrict__ qkv, const float* __restrict__ rcb,
                        const float* __restrict__ rpb, const float* __restrict__ relk,
                        __nv_bfloat16* __restrict__ out, int64_t rows, int H, int dk, int dv) {
  __shared__ float s_w[kAttnWarps][128];
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t seq = (int64_t)blockIdx.x * kAttnWarps + warp;
  if (seq >= rows) return;
  const int ld = 2 * H * dk + H * dv;
  tower::attn_warp_seq<N, T, U>(qkv + seq * N * ld, ld, rcb, rpb, relk, out + seq * N * (size_t)(H * dv), H, dk, dv, lane,
                                s_w[warp]);
}

int launch_attention(const float* qkv, const float* rcb, const float* rpb, const float* relk,
                     __nv_bfloat16* out, int64_t rows, int n, int H, int dk, int dv, cudaStream_t st) {
  const unsigned grid = (unsigned)ceil_div<int64_t>(rows * H, kAttnWarps);
  {
    const char* e = getenv("SVDD_ATTN_SEQ");
    const bool pow2 = (dk & (dk - 1)) == 0;
    if ((e ? atoi(e) : 1) && n <= 2 && pow2 && dk >= 4 && dk <= 128 && dv % 4 == 0 && H * dk == 512 &&
        (H * dv == 1536 || H * dv == 384) && H * n * n <= 128) {
      const unsigned gs = (unsigned)ceil_div<int64_t>(rows, kAttnWarps);
      const dim3 blk(kAttnWarps * 32);
      if (n == 1 && H * dv == 1536) launch_k(ef_attention_seq_kernel<1, 4, 12>, dim3(gs), blk, 0, st, 1, qkv, rcb, rpb, relk, out, rows, H, dk, dv);
      else if (n == 1) launch_k(ef_attention_seq_kernel<1, 4, 3>, dim3(gs), blk, 0, st, 1, qkv, rcb, rpb, relk, out, rows, H, dk, dv);
      else if (H * dv == 1536) launch_k(ef_attention_seq_kernel<2, 4, 12>, dim3(gs), blk, 0, st, 1, qkv, rcb, rpb, relk, out, rows, H, dk, dv);
      else launch_k(ef_attention_seq_kernel<2, 4, 3>, dim3(gs), blk, 0, st, 1, qkv, rcb, rpb, relk, out, rows, H, dk, dv);
      count_launch();
      SVDD_LAUNCH_CHECK();
      return SVDD_OK;
    }
  }
  if (dk <= 32 * kAttnDkPer && n == 1)
    launch_k(ef_attention_warp_kernel<1>, dim3(grid), dim3(kAttnWarps * 32), 0, st, 1, qkv, rcb, rpb, relk, out, rows, H, dk, dv);
  else if (dk <= 32 * kAttnDkPer && n == 2)
    launch_k(ef_attention_warp_kernel<2>, dim3(grid), dim3(kAttnWarps * 32), 0, st, 1, qkv, rcb, rpb, relk, out, rows, H, dk, dv);
  else if (dk <= 32 * kAttnDkPer && n == 3)
    launch_k(ef_attention_warp_kernel<3>, dim3(grid), dim3(kAttnWarps * 32), 0, st, 1, qkv, rcb, rpb, relk, out, rows, H, dk, dv);
  else if (dk <= 32 * kAttnDkPer && n == 4)
    launch_k(ef_attention_warp_kernel<4>, dim3(grid), dim3(kAttnWarps * 32), 0, st, 1, qkv, rcb, rpb, relk, out, rows, H, dk, dv);
  else
    launch_k(ef_attention_kernel, dim3((unsigned)rows), dim3(256), 0, st, 1, qkv, rcb, rpb, relk, out, n, H, dk, dv);
  count_launch();
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}

__global__ void ef_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                              int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i] + b[i];
}

// zero rows {0, 1, Lp-2, Lp-1} of every sequence of a bf16 [rows, Lp, C] buffer (HALO layout pads)
__global__ void ef_zero_pads_kernel(__nv_bfloat16* __restrict__ buf, int64_t rows, int Lp, int C) {
  pdl_wait();
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int cv = C / 8;                      // 16-byte vectors per row
  if (i >= rows * 4 * cv) return;
  const int v = (int)(i % cv);
  const int r4 = (int)((i / cv) % 4);
  const int64_t s = i / ((int64_t)cv * 4);
  const int l = r4 < 2 ? r4 : Lp - 4 + r4;
  reinterpret_cast<uint4*>(buf + ((size_t)s * Lp + l) * C)[v] = make_uint4(0u, 0u, 0u, 0u);
}

// relk[blk][h][p][d] = sum_f W[blk][(h*dk+d)][f] * pos[p][f]
__global__ void ef_relk_kernel(const float* __restrict__ w, const float* __restrict__ pos,
                               float* __restrict__ relk, int H, int dk, int F, int P) {
  const int blk = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * P * dk) return;
  const int d = i % dk, p = (i / dk) % P, h = i / (dk * P);
  const float* wr = w + ((size_t)blk * H * dk + h * dk + d) * F;
  float acc = 0.0f;
  for (int f = 0; f < F; ++f) acc += wr[f] * pos[p * F + f];
  relk[(size_t)blk * H * P * dk + i] = acc;
}

// score[s] = head_b + mean_p sum_tiles partials[(s*n+p), tile]
__global__ void ef_mean_kernel(const float* __restrict__ partials, int n_tiles, const float* __restrict__ hb,
                               float* __restrict__ scores, int64_t rows, int n) {
  pdl_wait();
  pdl_trigger();
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= rows) return;
  float acc = 0.0f;
  const float* p = partials + (size_t)s * n * n_tiles;
  for (int i = 0; i < n * n_tiles; ++i) acc += p[i];
  scores[s] = acc / (float)n + hb[0];
}

// enformer_pytorch get_positional_embed(n, F) (use_tf_gamma = False), computed once per n.
std::vector<float> positional_table(int n, int F) {
  const int P = 2 * n - 1, k = F / 6;
  std::vector<float> t((size_t)P * F, 0.0f);
  std::vector<double> base((size_t)3 * k);
  const double top = log((double)n) / log(2.0);
  for (int p = 0; p < P; ++p) {
    const double dist = (double)(p - (n - 1)), ad = fabs(dist);
    double gmax = 0.0;
    for (int i = 0; i < k; ++i) {
      const double frac = k > 1 ? (double)i / (k - 1) : 0.0;
      const double half_life = pow(2.0, 3.0 + (top - 3.0) * frac);
      base[i] = exp(-log(2.0) / half_life * ad);
      base[k + i] = ((pow(2.0, (double)(i + 1)) - 1.0) > ad) ? 1.0 : 0.0;
      const double stddev = (double)n / (2.0 * k);
      const double mean = (double)n / k + ((double)n - (double)n / k) * frac;
      const double conc = (mean / stddev) * (mean / stddev), rate = mean / (stddev * stddev);
      const double xlogy = (ad == 0.0) ? ((conc - 1.0) == 0.0 ? 0.0 : -INFINITY) : (conc - 1.0) * log(ad);
      const double logp = xlogy - rate * ad - (lgamma(conc) - conc * log(rate));
      base[2 * k + i] = exp(logp) + 1e-8;
      if (base[2 * k + i] > gmax) gmax = base[2 * k + i];
    }
    for (int i = 0; i < k; ++i) base[2 * k + i] /= gmax;
    const double sgn = dist > 0 ? 1.0 : (dist < 0 ? -1.0 : 0.0);
    for (int i = 0; i < 3 * k; ++i) {
      t[(size_t)p * F + i] = (float)base[i];
      t[(size_t)p * F + 3 * k + i] = (float)(sgn * base[i]);
    }
  }
  return t;
}

}  // namespace
}  // namespace svdd

using namespace svdd;

struct svdd_enformer {
  int H = 8, dk = 64, dv = 192, C = 1536, C0 = 768, F = 192;
  int n_stage = 7, n_blocks = 11, stem_taps = 15;
  int f[kMaxStages] = {};
  DeviceArena arena;
  __nv_bfloat16* stem_w = nullptr; float* stem_b = nullptr;
  __nv_bfloat16* w5[kMaxStages] = {}; float* b5[kMaxStages] = {};
  float* bn5_s[kMaxStages] = {}; float* bn5_t[kMaxStages] = {};   // BN feeding the k5 conv of stage i
  __nv_bfloat16* w1[kMaxStages] = {}; float* b1[kMaxStages] = {};
  float* b1s = nullptr;                                           // b1[0] + stem bias (K-concatenated stage 0)
  float* bn1_s[kMaxStages] = {}; float* bn1_t[kMaxStages] = {};   // BN feeding the 1x1 conv of stage i
  __nv_bfloat16* wp[kMaxStages] = {};
  struct Block {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *rcb, *rpb, *bo, *bf1, *bf2;
    __nv_bfloat16 *wqkv, *wo, *wf1, *wf2;
  } blk[kMaxBlocksT] = {};
  float* relk_w = nullptr;     // [n_blocks][H*dk][F]
  float *bnpw_s = nullptr, *bnpw_t = nullptr, *bpw = nullptr, *hw = nullptr, *hb = nullptr;
  __nv_bfloat16* wpw = nullptr;
  // lazily built for the transformer length n of the first call
  int relk_n = 0;
  float* relk = nullptr;       // [n_blocks][H][2n-1][dk]
  float* pos_dev = nullptr;
  // persistent tower kernel (tower.cuh): per-block weight tensor maps + parameter pointers
  tower::BlockParams* tower_blocks = nullptr;
  ~svdd_enformer() {
    if (relk) cudaFree(relk);
    if (pos_dev) cudaFree(pos_dev);
    if (tower_blocks) cudaFree(tower_blocks);
  }
};

extern "C" int svdd_enformer_create(const svdd_tensor* tensors, int n_tensors, int n_heads,
                                    void* stream, svdd_enformer** out) {
  SVDD_CHECK_ARG(tensors && out && n_tensors > 0, "svdd_enformer_create: null argument");
  SVDD_CHECK_ARG(n_heads >= 1 && n_heads <= 64, "svdd_enformer_create: bad n_heads %d", n_heads);
  int dev = 0;
  SVDD_CUDA(cudaGetDevice(&dev));
  SVDD_TRY(svdd_device_check(dev));
  cudaStream_t st = (cudaStream_t)stream;
  TensorTable tt{tensors, n_tensors};
  svdd_enformer* h = new (std::nothrow) svdd_enformer();
  SVDD_CHECK_ARG(h != nullptr, "out of host memory");
  auto fail = [&](int code) { delete h; return code; };
#define REQUIRE(cond, ...)                                        \
  do { if (!(cond)) { set_last_error(__VA_ARGS__); return fail(SVDD_ERR_INVALID_ARGUMENT); } } while (0)
#define GET_OR_FAIL(var, name, numel)               \
  const float* var = tt.get((name), (numel));       \
  if (var == nullptr) return fail(SVDD_ERR_MISSING_TENSOR)
#define TRY_OR_FAIL(expr)                           \
  do { int _rc = (expr); if (_rc != SVDD_OK) return fail(_rc); } while (0)

  // ---- discover the architecture from tensor shapes ------------------------------------
  const std::string ct = "conv_tower.blocks.";
  h->C0 = (int)tt.dim(ct + "0.0.weight", 0);
  h->stem_taps = (int)tt.dim(ct + "0.0.weight", 2);
  REQUIRE(h->C0 > 0 && tt.dim(ct + "0.0.weight", 1) == 4, "enformer: stem conv must take 4 channels");
  REQUIRE(h->stem_taps * 4 <= kStemK && h->stem_taps % 2 == 1, "enformer: unsupported stem kernel %d", h->stem_taps);
  int ns = 1;
  h->f[0] = h->C0;
  while (ns < kMaxStages && tt.has(ct + std::to_string(ns) + ".0.conv.weight")) {
    h->f[ns] = (int)tt.dim(ct + std::to_string(ns) + ".0.conv.weight", 0);
    ++ns;
  }
  h->n_stage = ns;
  h->C = h->f[ns - 1];
  int nb = 0;
  while (nb < kMaxBlocksT && tt.has("transformer_tower.blocks." + std::to_string(nb) + ".mha.to_q.weight")) ++nb;
  h->n_blocks = nb;
  h->H = n_heads;
  if (nb > 0) {
    h->dk = (int)tt.dim("transformer_tower.blocks.0.mha.to_q.weight", 0) / n_heads;
    h->dv = (int)tt.dim("transformer_tower.blocks.0.mha.to_v.weight", 0) / n_heads;
    h->F = (int)tt.dim("transformer_tower.blocks.0.mha.to_rel_k.weight", 1);
    REQUIRE(h->F % 6 == 0, "enformer: num_rel_pos_features %d not divisible by 6", h->F);
  }
  for (int i = 0; i < ns; ++i) REQUIRE(h->f[i] % 64 == 0, "enformer: channel count %d not a multiple of 64", h->f[i]);
  const int C = h->C, H = h->H, dk = h->dk, dv = h->dv, F = h->F;
  const int nqkv = 2 * H * dk + H * dv;
  REQUIRE(nqkv % 64 == 0 && (H * dv) % 64 == 0, "enformer: head dims must give multiples of 64");

  // ---- reserve / carve -----------------------------------------------------------------
  DeviceArena& A = h->arena;
  auto plan = [&](bool take) {
#define F32(ptr, n) do { if (take) ptr = A.take<float>(n); else A.reserve(sizeof(float) * (size_t)(n)); } while (0)
#define B16(ptr, n) do { if (take) ptr = A.take<__nv_bfloat16>(n); else A.reserve(sizeof(__nv_bfloat16) * (size_t)(n)); } while (0)
    B16(h->stem_w, (size_t)h->C0 * kStemK); F32(h->stem_b, h->C0); F32(h->b1s, h->C0);
    for (int i = 0; i < ns; ++i) {
      const size_t fi = h->f[i], fp = i > 0 ? h->f[i - 1] : 0;
      if (i > 0) { B16(h->w5[i], 5 * fi * fp); F32(h->b5[i], fi); F32(h->bn5_s[i], fp); F32(h->bn5_t[i], fp); }
      B16(h->w1[i], fi * fi); F32(h->b1[i], fi); F32(h->bn1_s[i], fi); F32(h->bn1_t[i], fi);
      B16(h->wp[i], fi * fi);
    }
    for (int j = 0; j < nb; ++j) {
      auto& b = h->blk[j];
      F32(b.ln1_g, C); F32(b.ln1_b, C); F32(b.ln2_g, C); F32(b.ln2_b, C);
      F32(b.rcb, H * dk); F32(b.rpb, H * dk); F32(b.bo, C); F32(b.bf1, 2 * C); F32(b.bf2, C);
      B16(b.wqkv, (size_t)nqkv * C); B16(b.wo, (size_t)C * H * dv);
      B16(b.wf1, (size_t)2 * C * C); B16(b.wf2, (size_t)2 * C * C);
    }
    F32(h->relk_w, (size_t)(nb > 0 ? nb : 1) * H * dk * F);
    F32(h->bnpw_s, C); F32(h->bnpw_t, C); F32(h->bpw, 2 * C); F32(h->hw, 2 * C); F32(h->hb, 8);
    B16(h->wpw, (size_t)2 * C * C);
#undef F32
#undef B16
  };
  plan(false);
  TRY_OR_FAIL(A.commit());
  plan(true);

  // ---- pack ------------------------------------------------------------------------------
  auto bn_fold = [&](const std::string& p, int n, float* s, float* t) -> int {
    const float* g = tt.get(p + "weight", n);
    const float* b = tt.get(p + "bias", n);
    const float* m = tt.get(p + "running_mean", n);
    const float* v = tt.get(p + "running_var", n);
    if (!g || !b || !m || !v) return SVDD_ERR_MISSING_TENSOR;
    return fold_bn(g, b, m, v, nullptr, 1e-5f, s, t, n, st);
  };
  {
    GET_OR_FAIL(sw, ct + "0.0.weight", (int64_t)h->C0 * 4 * h->stem_taps);
    GET_OR_FAIL(sb, ct + "0.0.bias", h->C0);
    ef_pack_stem_kernel<<<ceil_div(h->C0 * kStemK, 256), 256, 0, st>>>(sw, h->stem_w, h->C0, h->stem_taps);
    TRY_OR_FAIL(copy_f32(sb, h->stem_b, h->C0, st));
  }
  for (int i = 0; i < ns; ++i) {
    const int fi = h->f[i];
    const std::string p1 = ct + std::to_string(i) + ".1.";
    if (i > 0) {
      const int fp = h->f[i - 1];
      const std::string p0 = ct + std::to_string(i) + ".0.";
      GET_OR_FAIL(w, p0 + "conv.weight", (int64_t)fi * fp * 5);
      GET_OR_FAIL(b, p0 + "conv.bias", fi);
      TRY_OR_FAIL(pack_conv_weight(w, h->w5[i], fi, fp, 5, st));
      TRY_OR_FAIL(copy_f32(b, h->b5[i], fi, st));
      TRY_OR_FAIL(bn_fold(p0 + "norm.layer.", fp, h->bn5_s[i], h->bn5_t[i]));
    }
    GET_OR_FAIL(w, p1 + "conv.weight", (int64_t)fi * fi);
    GET_OR_FAIL(b, p1 + "conv.bias", fi);
    GET_OR_FAIL(wp, p1 + "pool.layer.to_attn_logits.weight", (int64_t)fi * fi);
    TRY_OR_FAIL(pack_conv_weight(w, h->w1[i], fi, fi, 1, st));
    TRY_OR_FAIL(copy_f32(b, h->b1[i], fi, st));
    if (i == 0) {
      ef_add_kernel<<<ceil_div(fi, 256), 256, 0, st>>>(h->b1[0], h->stem_b, h->b1s, fi);
    }
    TRY_OR_FAIL(pack_conv_weight(wp, h->wp[i], fi, fi, 1, st));
    TRY_OR_FAIL(bn_fold(p1 + "norm.layer.", fi, h->bn1_s[i], h->bn1_t[i]));
  }
  for (int j = 0; j < nb; ++j) {
    auto& b = h->blk[j];
    const std::string p = "transformer_tower.blocks." + std::to_string(j) + ".";
    GET_OR_FAIL(g1, p + "norm.layer.weight", C);
    GET_OR_FAIL(b1, p + "norm.layer.bias", C);
    GET_OR_FAIL(g2, p + "ffn.dense1.norm.layer.weight", C);
    GET_OR_FAIL(b2, p + "ffn.dense1.norm.layer.bias", C);
    GET_OR_FAIL(wq, p + "mha.to_q.weight", (int64_t)H * dk * C);
    GET_OR_FAIL(wk, p + "mha.to_k.weight", (int64_t)H * dk * C);
    GET_OR_FAIL(wv, p + "mha.to_v.weight", (int64_t)H * dv * C);
    GET_OR_FAIL(wo, p + "mha.to_out.weight", (int64_t)C * H * dv);
    GET_OR_FAIL(bo, p + "mha.to_out.bias", C);
    GET_OR_FAIL(wr, p + "mha.to_rel_k.weight", (int64_t)H * dk * F);
    GET_OR_FAIL(rc, p + "mha.rel_content_bias", H * dk);
    GET_OR_FAIL(rp, p + "mha.rel_pos_bias", H * dk);
    GET_OR_FAIL(w1, p + "ffn.dense1.linear.weight", (int64_t)2 * C * C);
    GET_OR_FAIL(c1, p + "ffn.dense1.linear.bias", 2 * C);
    GET_OR_FAIL(w2, p + "ffn.dense2.linear.weight", (int64_t)2 * C * C);
    GET_OR_FAIL(c2, p + "ffn.dense2.linear.bias", C);
    TRY_OR_FAIL(copy_f32(g1, b.ln1_g, C, st)); TRY_OR_FAIL(copy_f32(b1, b.ln1_b, C, st));
    TRY_OR_FAIL(copy_f32(g2, b.ln2_g, C, st)); TRY_OR_FAIL(copy_f32(b2, b.ln2_b, C, st));
    TRY_OR_FAIL(pack_conv_weight(wq, b.wqkv, H * dk, C, 1, st));
    TRY_OR_FAIL(pack_conv_weight(wk, b.wqkv + (size_t)H * dk * C, H * dk, C, 1, st));
    TRY_OR_FAIL(pack_conv_weight(wv, b.wqkv + (size_t)2 * H * dk * C, H * dv, C, 1, st));
    TRY_OR_FAIL(pack_conv_weight(wo, b.wo, C, H * dv, 1, st));
    TRY_OR_FAIL(copy_f32(bo, b.bo, C, st));
    TRY_OR_FAIL(copy_f32(wr, h->relk_w + (size_t)j * H * dk * F, (int64_t)H * dk * F, st));
    TRY_OR_FAIL(copy_f32(rc, b.rcb, H * dk, st)); TRY_OR_FAIL(copy_f32(rp, b.rpb, H * dk, st));
    TRY_OR_FAIL(pack_conv_weight(w1, b.wf1, 2 * C, C, 1, st)); TRY_OR_FAIL(copy_f32(c1, b.bf1, 2 * C, st));
    TRY_OR_FAIL(pack_conv_weight(w2, b.wf2, C, 2 * C, 1, st)); TRY_OR_FAIL(copy_f32(c2, b.bf2, C, st));
  }
  {
    GET_OR_FAIL(w, "pointwise_conv.conv.weight", (int64_t)2 * C * C);
    GET_OR_FAIL(b, "pointwise_conv.conv.bias", 2 * C);
    GET_OR_FAIL(hw, "head.channel_transform.conv.layer.weight", 2 * C);
    GET_OR_FAIL(hb, "head.channel_transform.conv.layer.bias", 1);
    TRY_OR_FAIL(pack_conv_weight(w, h->wpw, 2 * C, C, 1, st));
    TRY_OR_FAIL(copy_f32(b, h->bpw, 2 * C, st));
    TRY_OR_FAIL(copy_f32(hw, h->hw, 2 * C, st));
    TRY_OR_FAIL(copy_f32(hb, h->hb, 1, st));
    TRY_OR_FAIL(bn_fold("pointwise_conv.norm.layer.", C, h->bnpw_s, h->bnpw_t));
  }
#undef REQUIRE
#undef GET_OR_FAIL
#undef TRY_OR_FAIL
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("svdd_enformer_create: %s", cudaGetErrorString(e));
    return fail(SVDD_ERR_CUDA);
  }
  *out = h;
  return SVDD_OK;
}

extern "C" void svdd_enformer_destroy(svdd_enformer* h) { delete h; }

// ---- unit-test hooks for the two non-GEMM pieces -------------------------------------------
extern "C" int svdd_selftest_rel_positions(int n, int F, float* out_host) {
  SVDD_CHECK_ARG(n >= 1 && F >= 6 && F % 6 == 0 && out_host, "selftest_rel_positions: bad argument");
  std::vector<float> t = positional_table(n, F);
  memcpy(out_host, t.data(), t.size() * sizeof(float));
  return SVDD_OK;
}

extern "C" int svdd_selftest_attention(const float* qkv, const float* rcb, const float* rpb,
                                       const float* relk, void* out_bf16, int64_t rows, int n,
                                       int H, int dk, int dv, void* stream) {
  SVDD_CHECK_ARG(qkv && rcb && rpb && relk && out_bf16, "selftest_attention: null pointer");
  SVDD_CHECK_ARG(n >= 1 && n <= kMaxPos, "selftest_attention: n out of range");
  return launch_attention(qkv, rcb, rpb, relk, (__nv_bfloat16*)out_bf16, rows, n, H, dk, dv,
                          (cudaStream_t)stream);
}

namespace {
struct EfWs {
  __nv_bfloat16* col;
  __nv_bfloat16* big[3];
  float* xt;
  __nv_bfloat16* hn;
  float* qkv;
  __nv_bfloat16* ao;
  __nv_bfloat16* u;
  float* partials;
  unsigned* flags;       // persistent tower kernel: one counter per (phase, row tile)
};
bool pool2_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVDD_POOL2"); v = e ? atoi(e) : 1; }
  return v != 0;
}
bool halo_enabled() {
  static int v = -1;
  // off by default: measured on the c2 step the padded flat layout is SLOWER (L=100 conv 578 vs
  // 525 us, L=50 394 vs 369 us): the k5 convs already run at the sustained tensor peak (~63 % of
  // nominal under the power cap), so removing the per-tap re-fetch of A buys nothing and the pad
  // rows cost 4-8 % more MMA work.  Kept (SVDD_HALO=1) with its parity test.
  if (v < 0) { const char* e = getenv("SVDD_HALO"); v = e ? atoi(e) : 0; }
  return v != 0;
}
bool stem_concat_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVDD_STEM_CONCAT"); v = e ? atoi(e) : 1; }
  return v != 0;
}
int ef_final_len(const svdd_enformer* h, int L) {
  int n = L;
  for (int i = 0; i < h->n_stage; ++i) n = (n + 1) / 2;
  return n;
}
size_t ef_carve(const svdd_enformer* h, Workspace& W, int64_t rows, int L, EfWs* o) {
  const size_t NL = (size_t)rows * L;
  size_t big = 0;
  int len = L;
  for (int i = 0; i < h->n_stage; ++i) {
    const size_t need = (size_t)rows * (len + 4) * h->f[i];      // + the pad rows of the HALO layout
    if (need > big) big = need;
    len = (len + 1) / 2;
  }
  const int n = len;
  const size_t R = (size_t)rows * n;
  const int C = h->C, nqkv = 2 * h->H * h->dk + h->H * h->dv;
  o->col = W.take<__nv_bfloat16>(NL * kStemK + 64);
  for (int i = 0; i < 3; ++i) o->big[i] = W.take<__nv_bfloat16>(big + 4096);
  o->xt = W.take<float>(R * C + 64);
  o->hn = W.take<__nv_bfloat16>(R * C + 64);
  o->qkv = W.take<float>(R * nqkv + 64);
  o->ao = W.take<__nv_bfloat16>(R * h->H * h->dv + 64);
  o->u = W.take<__nv_bfloat16>(R * 2 * C + 64);
  o->partials = W.take<float>(R * (2 * C / 64) + 64);
  o->flags = W.take<unsigned>((size_t)(tower::kPhasesPerBlock + tower::kMaxSplitTiles) * kMaxBlocksT *
                              ceil_div<size_t>(R, tower::kTileRows) + 64);
  return W.used();
}
}  // namespace

namespace {
// ---- persistent tower kernel: host side ---------------------------------------------------------
int tower_bn() {               // tile width of the tower's GEMM items (read once)
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVDD_TOWER_BN"); v = (e && atoi(e) == 128) ? 128 : 256; }
  return v;
}
// epilogue warps of the tower kernel (read once).  16 (thread = row x column quarter, one row per warp
// in the LayerNorm items, two warps per sequence in the attention items) is measured SLOWER, 1.17 vs
// 1.07 ms: at the 102-register cap of a 640-thread CTA the row code and the GEMM epilogue spill
// (attention 9.4 -> 8.3 us, but LayerNorm 6.1 -> 7.3 and every GEMM phase +1.5 us).  Kept with its test.
int tower_ew() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVDD_TOWER_EW"); v = (e && atoi(e) == 16) ? 16 : 8; if (tower_bn() != 256) v = 8; }
  return v;
}
int tower_enabled() {          // read per call: the tests A/B the two paths in one process
  const char* e = getenv("SVDD_TOWER");
  return e ? atoi(e) : 1;
}
// The tower kernel variant for n positions per sequence, configured for its dynamic shared memory.
struct TowerKernel {
  void (*kern)(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap,
               const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, tower::TowerArgs);
  int smem_bytes, ew, threads;
};
TowerKernel tower_select(int n) {
  using namespace tower;
  const int bn = tower_bn(), ew = tower_ew();
  TowerKernel t;
  t.kern = ew == 16 ? (n == 1 ? tower_kernel<1, 256, 16> : (n == 2 ? tower_kernel<2, 256, 16> : tower_kernel<4, 256, 16>))
         : bn == 256 ? (n == 1 ? tower_kernel<1, 256> : (n == 2 ? tower_kernel<2, 256> : tower_kernel<4, 256>))
                     : (n == 1 ? tower_kernel<1, 128> : (n == 2 ? tower_kernel<2, 128> : tower_kernel<4, 128>));
  t.smem_bytes = bn == 256 ? Cfg<256>::kSmemBytes : Cfg<128>::kSmemBytes;
  t.ew = ew;
  t.threads = 64 + 32 * ew + 64;
  static bool configured[3] = {false, false, false};
  const int ki = n == 1 ? 0 : (n == 2 ? 1 : 2);
  if (!configured[ki]) {
    if (cudaFuncSetAttribute(t.kern, cudaFuncAttributeMaxDynamicSharedMemorySize, t.smem_bytes) == cudaSuccess)
      configured[ki] = true;
  }
  return t;
}
// How many CTA pairs (clusters of 2) of the tower kernel the device can hold at once.  The kernel's
// work items wait on each other through global flags, so every launched pair MUST be resident:
// the grid never exceeds this count, and 0 (no cluster fits: e.g. shared memory carved out by
// another context) sends the caller to the launch-per-GEMM path instead of a kernel that would
// spin until its watchdog traps.
int tower_resident_pairs(int n) {
  static int cached[3] = {-1, -1, -1};
  const int ki = n == 1 ? 0 : (n == 2 ? 1 : 2);
  if (cached[ki] >= 0) return cached[ki];
  TowerKernel t = tower_select(n);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * (num_sms() / 2)));
  cfg.blockDim = dim3((unsigned)t.threads);
  cfg.dynamicSmemBytes = (size_t)t.smem_bytes;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, (const void*)t.kern, &cfg) != cudaSuccess) { cudaGetLastError(); clusters = 0; }
  const int by_sm = num_sms() / 2;             // one CTA per SM (227 KB of shared memory each)
  cached[ki] = clusters < by_sm ? clusters : by_sm;
  return cached[ki];
}
bool tower_usable(const svdd_enformer* h, int n) {
  const int C = h->C, nqkv = 2 * h->H * h->dk + h->H * h->dv;
  return tower_enabled() && h->tower_blocks != nullptr && h->n_blocks > 0 && (n == 1 || n == 2 || n == 4) &&
         C % 128 == 0 && C <= 128 * tower::kLnVecs && nqkv % 64 == 0 && (h->H * h->dv) % 64 == 0 &&
         h->dk <= 32 * tower::kAttnDkPer && num_sms() >= 2 && tower_resident_pairs(n) >= 1;
}
int build_tower_params(svdd_enformer* h, int n) {
  if (h->tower_blocks) { cudaFree(h->tower_blocks); h->tower_blocks = nullptr; }
  const int C = h->C, H = h->H, dk = h->dk, dv = h->dv, P = 2 * n - 1;
  const int nqkv = 2 * H * dk + H * dv;
  if (h->n_blocks == 0 || C % 128 != 0 || (H * dv) % 64 != 0 || nqkv % 64 != 0) return SVDD_OK;
  std::vector<tower::BlockParams> hp((size_t)h->n_blocks);
  for (int j = 0; j < h->n_blocks; ++j) {
    const auto& b = h->blk[j];
    tower::BlockParams& bp = hp[(size_t)j];
    SVDD_TRY(encode_tmap_2d_bf16(&bp.w[0], b.wqkv, (uint64_t)C, (uint64_t)nqkv, 64, (uint32_t)(tower_bn() / 2)));
    SVDD_TRY(encode_tmap_2d_bf16(&bp.w[1], b.wo, (uint64_t)(H * dv), (uint64_t)C, 64, (uint32_t)(tower_bn() / 2)));
    SVDD_TRY(encode_tmap_2d_bf16(&bp.w[2], b.wf1, (uint64_t)C, (uint64_t)(2 * C), 64, (uint32_t)(tower_bn() / 2)));
    SVDD_TRY(encode_tmap_2d_bf16(&bp.w[3], b.wf2, (uint64_t)(2 * C), (uint64_t)C, 64, (uint32_t)(tower_bn() / 2)));
    bp.ln_g[0] = b.ln1_g; bp.ln_b[0] = b.ln1_b; bp.ln_g[1] = b.ln2_g; bp.ln_b[1] = b.ln2_b;
    bp.bias[0] = nullptr; bp.bias[1] = b.bo; bp.bias[2] = b.bf1; bp.bias[3] = b.bf2;
    bp.rcb = b.rcb; bp.rpb = b.rpb;
    bp.relk = h->relk + (size_t)j * H * P * dk;
  }
  SVDD_CUDA(cudaMalloc(&h->tower_blocks, hp.size() * sizeof(tower::BlockParams)));
  SVDD_CUDA(cudaMemcpy(h->tower_blocks, hp.data(), hp.size() * sizeof(tower::BlockParams), cudaMemcpyHostToDevice));
  return SVDD_OK;
}

// All transformer blocks + the pointwise ConvBlock's BN+GELU operand in one launch:
// in  b.xt (fp32 [R, C], the pooled conv-tower output), out b.xt (updated) and b.hn = GELU(BN(xt)).
int launch_tower(const svdd_enformer* h, const EfWs& b, int64_t R, int n, cudaStream_t st) {
  using namespace tower;
  const int C = h->C, H = h->H, dk = h->dk, dv = h->dv;
  const int nqkv = 2 * H * dk + H * dv;
  const int bn = tower_bn();
  TowerArgs a;
  a.R = (int)R; a.RT = (int)ceil_div<int64_t>(R, kTileRows); a.n_blocks = h->n_blocks; a.n_pos = n;
  a.C = C; a.nqkv = nqkv; a.H = H; a.dk = dk; a.dv = dv;
  auto gemm_phase = [&](int N, int K, int a_map, int out_kind, int w_idx) {
    Phase p;
    p.type = PH_GEMM; p.items_per_rt = ceil_div(N, bn); p.n_cols = N; p.signals = 4; p.kblocks = K / 64;
    p.a_map = a_map; p.out_kind = out_kind; p.w_idx = w_idx;
    return p;
  };
  auto row_phase = [&](int type, int w_idx) {
    Phase p;
    p.type = type; p.items_per_rt = kSubItems; p.signals = 2; p.w_idx = w_idx;
    return p;
  };
  a.ph[0] = row_phase(PH_LN, 0);
  a.ph[1] = gemm_phase(nqkv, C, A_HN, OUT_F32_STORE, 0);
  a.ph[2] = row_phase(PH_ATTN, 0);
  a.ph[3] = gemm_phase(C, H * dv, A_AO, OUT_F32_REDUCE, 1);
  a.ph[4] = row_phase(PH_LN, 1);
  a.ph[5] = gemm_phase(2 * C, C, A_HN, OUT_BF16_RELU, 2);
  a.ph[6] = gemm_phase(C, 2 * C, A_U, OUT_F32_REDUCE, 3);
  {
    // FF2 (K = 2C, the longest item of the chain): optionally K split in two slices whose partial sums
    // are added in order (SVDD_TOWER_SPLITK=2).  Off: measured 1.09 vs 1.07 ms -- the FF2 phase gets
    // shorter (21.4 -> 17.9 us) but its 120 items no longer fit one round of the 74 pairs.
    const char* e = getenv("SVDD_TOWER_SPLITK");
    const int ks = e ? atoi(e) : 1;
    if (ks == 2 && a.ph[6].kblocks % 2 == 0 && a.ph[6].items_per_rt <= kMaxSplitTiles) {
      a.ph[6].ksplit = 2;
      a.ph[6].items_per_rt *= 2;
      a.ph[6].kblocks /= 2;
    }
  }
  a.ph[7] = row_phase(PH_BNACT, 0);
  int start = 0;
  for (int q = 0; q < kPhasesPerBlock; ++q) {
    a.ph[q].start = start;
    start += a.ph[q].items_per_rt * a.RT;
  }
  a.items_per_block = start;
  a.total_items = start * a.n_blocks + kSubItems * a.RT;
  a.blocks = h->tower_blocks;
  a.xt = b.xt; a.hn = b.hn; a.qkv = b.qkv; a.ao = b.ao; a.u = b.u;
  a.bn_s = h->bnpw_s; a.bn_t = h->bnpw_t;
  a.flags = b.flags;
  a.sk_flags = b.flags + (size_t)kPhasesPerBlock * a.n_blocks * a.RT;
  { const char* e = getenv("SVDD_TOWER_ATTN_FAST"); a.attn_fast = e ? atoi(e) : 1; }
  { const char* e = getenv("SVDD_TOWER_DISCARD"); a.discard_qkv = e ? atoi(e) : 1; }

  CUtensorMap tm_hn, tm_ao, tm_u, tm_qkv, tm_xt;
  SVDD_TRY(encode_tmap_2d_bf16(&tm_hn, b.hn, (uint64_t)C, (uint64_t)R, 64, 128));
  SVDD_TRY(encode_tmap_2d_bf16(&tm_ao, b.ao, (uint64_t)(H * dv), (uint64_t)R, 64, 128));
  SVDD_TRY(encode_tmap_2d_bf16(&tm_u, b.u, (uint64_t)(2 * C), (uint64_t)R, 64, 128));
  SVDD_TRY(encode_tmap_2d_f32(&tm_qkv, b.qkv, (uint64_t)nqkv, (uint64_t)R, 32, 128));
  SVDD_TRY(encode_tmap_2d_f32(&tm_xt, b.xt, (uint64_t)C, (uint64_t)R, 32, 128));

  TowerKernel tk = tower_select(n);
  auto kern = tk.kern;
  const int smem_bytes = tk.smem_bytes, ew = tk.ew;
  // every CTA pair must be resident (items wait on each other): at most one pair per TPC
  static int max_pairs = -1;
  if (max_pairs < 0) { const char* e = getenv("SVDD_TOWER_PAIRS"); max_pairs = e ? atoi(e) : 0; }
  // ... and the grid is sized from what the driver says can be co-resident for THIS kernel and
  // shared-memory configuration (MIG slices, parts with fewer usable TPCs), not from the SM count
  int pairs = tower_resident_pairs(n);
  if (max_pairs > 0 && max_pairs < pairs) pairs = max_pairs;
  const int widest = a.ph[5].items_per_rt > a.ph[1].items_per_rt ? a.ph[5].items_per_rt : a.ph[1].items_per_rt;
  const int max_items_per_phase = (widest > kSubItems ? widest : kSubItems) * a.RT;
  if (pairs > max_items_per_phase) pairs = max_items_per_phase;
  SVDD_CUDA(cudaMemsetAsync(b.flags, 0, (size_t)(kPhasesPerBlock + kMaxSplitTiles) * a.n_blocks * a.RT * sizeof(unsigned), st));
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool prof = gemm_prof_on();
  if (prof) {
    SVDD_CUDA(cudaEventCreate(&e0));
    SVDD_CUDA(cudaEventCreate(&e1));
    SVDD_CUDA(cudaEventRecord(e0, st));
  }
  // SVDD_TOWER_TRACE=<file> (debugging): per-item timestamps of this launch, appended as CSV
  const char* trace_path = getenv("SVDD_TOWER_TRACE");
  unsigned long long* trace_dev = nullptr;
  if (trace_path != nullptr && trace_path[0] != 0) {
    SVDD_CUDA(cudaMalloc(&trace_dev, (size_t)a.total_items * 8 * sizeof(unsigned long long)));
    SVDD_CUDA(cudaMemsetAsync(trace_dev, 0, (size_t)a.total_items * 8 * sizeof(unsigned long long), st));
    a.trace = trace_dev;
  }
  SVDD_CUDA(launch_k(kern, dim3((unsigned)(2 * pairs)), dim3(64 + 32 * ew + 64), smem_bytes, st, 2,
                     tm_hn, tm_ao, tm_u, tm_qkv, tm_xt, a));
  count_launch();
  if (trace_dev != nullptr) {
    std::vector<unsigned long long> tr((size_t)a.total_items * 8);
    SVDD_CUDA(cudaStreamSynchronize(st));
    SVDD_CUDA(cudaMemcpy(tr.data(), trace_dev, tr.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(trace_dev);
    FILE* f = fopen(trace_path, "a");
    if (f != nullptr) {
      unsigned long long t0 = ~0ull;
      for (int i = 0; i < a.total_items; ++i)
        for (int k = 0; k < 6; ++k) if (tr[(size_t)i * 8 + k] != 0 && tr[(size_t)i * 8 + k] < t0) t0 = tr[(size_t)i * 8 + k];
      fprintf(f, "# tower R=%d RT=%d pairs=%d items=%d items_per_block=%d\n", a.R, a.RT, pairs, a.total_items, a.items_per_block);
      for (int i = 0; i < a.total_items; ++i) {
        int j = i / a.items_per_block, rem = i % a.items_per_block, q = 0;
        if (i >= a.n_blocks * a.items_per_block) { j = a.n_blocks; q = kPhasesPerBlock; rem = i - a.n_blocks * a.items_per_block; }
        else { for (int k = 1; k < kPhasesPerBlock; ++k) q += rem >= a.ph[k].start ? 1 : 0; rem -= a.ph[q].start; }
        const int ipr = a.ph[q].items_per_rt;
        fprintf(f, "%d,%d,%d,%d,%d,%llu", i, j, q, rem / ipr, rem % ipr, tr[(size_t)i * 8 + 6]);
        for (int k = 0; k < 6; ++k) fprintf(f, ",%lld", tr[(size_t)i * 8 + k] ? (long long)(tr[(size_t)i * 8 + k] - t0) : -1ll);
        fprintf(f, "\n");
      }
      fclose(f);
    }
  }
  if (prof) {
    SVDD_CUDA(cudaEventRecord(e1, st));
    GemmShape g;
    g.S = 1; g.L = (int)R; g.L_in = (int)R; g.K = C; g.N = nqkv + C + 4 * C; g.taps = 1;
    const double fl = 2.0 * (double)R * h->n_blocks *
                      ((double)nqkv * C + (double)C * (H * dv) + 2.0 * (double)(2 * C) * C);
    gemm_prof_record(e0, e1, g, 2256, 100, fl);
  }
  return SVDD_OK;
}
}  // namespace

extern "C" size_t svdd_enformer_workspace_bytes(const svdd_enformer* h, int64_t n_rows, int L) {
  Workspace W(nullptr, 0);
  EfWs o;
  return ef_carve(h, W, n_rows < kChunkRows ? n_rows : kChunkRows, L, &o);
}

extern "C" int svdd_enformer_score(svdd_enformer* h, const void* tokens, int tok_dtype, float* scores,
                                   int64_t n_rows, int L, void* ws, size_t ws_bytes, void* stream) {
  SVDD_CHECK_ARG(h && tokens && scores, "svdd_enformer_score: null pointer");
  SVDD_CHECK_ARG(n_rows >= 0 && L >= 1, "svdd_enformer_score: bad shape");
  SVDD_CHECK_ARG(tok_dtype == SVDD_TOK_I64 || tok_dtype == SVDD_TOK_U8, "bad tok_dtype %d", tok_dtype);
  if (n_rows == 0) return SVDD_OK;
  if (ws == nullptr || ws_bytes < svdd_enformer_workspace_bytes(h, n_rows, L)) {
    set_last_error("svdd_enformer_score: workspace too small");
    return SVDD_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int n = ef_final_len(h, L);
  SVDD_CHECK_ARG(n >= 1 && n <= kMaxPos, "svdd_enformer_score: %d positions reach the transformer (max %d)", n, kMaxPos);
  const int C = h->C, H = h->H, dk = h->dk, dv = h->dv, F = h->F;
  const int nqkv = 2 * H * dk + H * dv;
  const int P = 2 * n - 1;

  // relative-position keys for this n (first call only; not on the hot path)
  if (h->n_blocks > 0 && h->relk_n != n) {
    if (h->relk) { cudaFree(h->relk); h->relk = nullptr; }
    if (h->pos_dev) { cudaFree(h->pos_dev); h->pos_dev = nullptr; }
    std::vector<float> pos = positional_table(n, F);
    SVDD_CUDA(cudaMalloc(&h->pos_dev, pos.size() * sizeof(float)));
    SVDD_CUDA(cudaMalloc(&h->relk, (size_t)h->n_blocks * H * P * dk * sizeof(float)));
    SVDD_CUDA(cudaMemcpy(h->pos_dev, pos.data(), pos.size() * sizeof(float), cudaMemcpyHostToDevice));
    ef_relk_kernel<<<dim3(ceil_div(H * P * dk, 128), h->n_blocks), 128, 0, st>>>(h->relk_w, h->pos_dev, h->relk,
                                                                                H, dk, F, P);
    SVDD_LAUNCH_CHECK();
    h->relk_n = n;
    SVDD_TRY(build_tower_params(h, n));
  }

  auto gemm_flat = [&](const void* A, const void* Wt, int64_t R, int K, int N, const EpiParams& ep,
                       int mode) -> int {
    GemmShape g;
    g.S = 1; g.L = (int)R; g.L_in = (int)R; g.K = K; g.N = N; g.taps = 1; g.dil = 1; g.BL = 128; g.BS = 1;
    return launch_conv_gemm(A, Wt, g, mode, ep, st);
  };

  const size_t tok_bytes = tok_dtype == SVDD_TOK_I64 ? 8 : 1;
  for (int64_t r0 = 0; r0 < n_rows; r0 += kChunkRows) {
    const int64_t rows = (n_rows - r0 < kChunkRows) ? n_rows - r0 : kChunkRows;
    const int64_t NL = rows * L;
    Workspace W(ws, ws_bytes);
    EfWs b;
    ef_carve(h, W, rows, L, &b);
    const void* tok = reinterpret_cast<const uint8_t*>(tokens) + (size_t)r0 * L * tok_bytes;

    // ---- stem -----------------------------------------------------------------------------
    {
      const unsigned grid = (unsigned)ceil_div<int64_t>(NL * 8, 256);
      if (tok_dtype == SVDD_TOK_I64)
        launch_k(ef_im2col_kernel<int64_t>, dim3(grid), dim3(256), 0, st, 1, (const int64_t*)tok, b.col, NL, L, h->stem_taps);
      else
        launch_k(ef_im2col_kernel<uint8_t>, dim3(grid), dim3(256), 0, st, 1, (const uint8_t*)tok, b.col, NL, L, h->stem_taps);
      count_launch();
      SVDD_LAUNCH_CHECK();
    }
    __nv_bfloat16 *x = b.big[0], *a = b.big[1], *y = b.big[2];   // roles rotate below
    // Stage 0's residual 1x1 conv needs x0 = stem(col) only as an addend: y = W1.a0 + b1 + x0 with
    // x0 = Ws.col + bs is ONE contraction over the concatenated K = [a0 | col] (768 + 64), so x0
    // is never written (393 MB at c2) nor read back; the stem GEMM only emits a0.
    const bool stem_concat = stem_concat_enabled() && pool2_enabled() && h->C0 % 128 == 0 &&
                             (L % 2 == 0 || L + 1 <= 128);
    {
      EpiParams ep;                                  // x0 = conv + b ; a0 = GELU(BN_{0.1}(x0))
      ep.bias = h->stem_b;
      if (!stem_concat) { ep.out = x; ep.out_dtype = DT_BF16; ep.ld_out = h->C0; }
      ep.out2 = a; ep.out2_dtype = DT_BF16; ep.ld_out2 = h->C0;
      ep.scale2 = h->bn1_s[0]; ep.shift2 = h->bn1_t[0]; ep.act2 = ACT_GELU;
      SVDD_TRY(gemm_flat(b.col, h->stem_w, NL, kStemK, h->C0, ep, EPI_GENERIC));
      if (debug_dump_enabled()) {
        if (!stem_concat) debug_dump("ef_x0", x, (size_t)NL * h->C0 * 2, st);
        debug_dump("ef_a0", a, (size_t)NL * h->C0 * 2, st);
      }
    }
    int len = L;
    bool halo_prev = false;              // the previous stage ran in the padded flat layout
    // "snake" tile order (SVDD_SNAKE, read per call, default on): every other GEMM of the conv tower walks its
    // row tiles from the last to the first, so that a layer starts on the rows its producer wrote last
    // (still in L2) instead of the ones written first (evicted at c2 sizes); bit-identical
    const char* e_snake = getenv("SVDD_SNAKE");
    const int snake = (e_snake == nullptr || atoi(e_snake) != 0) ? 1 : 0;
    int li = 1;                          // the stem walked forwards
    for (int i = 0; i < h->n_stage; ++i) {
      const int fi = h->f[i];
      const int lo = (len + 1) / 2;
      const bool pair_path = pool2_enabled() && fi % 128 == 0 && (len % 2 == 0 || len + 1 <= 128);
      // HALO layout of stage i (its k5 conv, residual 1x1 conv and the pool's inputs): every
      // sequence carries 2 zero rows either side ([rows, len + 4, C] flat), so the conv's five taps
      // are five row offsets into ONE shared-memory copy of the activation tile instead of five
      // TMA fetches through L2 -- the k5 convs at L = 100 / 50 were bound by the L2 -> SM fabric.
      const bool halo = i > 0 && halo_prev;
      const int Lp = len + 4;
      if (i > 0) {
        // z = Conv_k5(a) + b ; a' = GELU(BN_{i.1}(z)).  `a` holds GELU(BN_{i.0}(pooled)).
        const int fp = h->f[i - 1];
        GemmShape g;
        g.K = fp; g.N = fi; g.taps = 5; g.dil = 1;
        if (halo) {
          g.S = 1; g.L = (int)(rows * Lp); g.L_in = g.L; g.BL = 128; g.BS = 1; g.halo = 1;
          g.useful_rows = rows * len;
        } else {
          g.S = (int)rows; g.L = len; g.L_in = len;
          choose_row_tiling(len, 5, &g);
        }
        EpiParams ep;
        ep.bias = h->b5[i];
        ep.out = x; ep.out_dtype = DT_BF16; ep.ld_out = fi;
        ep.out2 = y; ep.out2_dtype = DT_BF16; ep.ld_out2 = fi;
        ep.scale2 = h->bn1_s[i]; ep.shift2 = h->bn1_t[i]; ep.act2 = ACT_GELU;
        g.reverse = snake & (li++ & 1);
        SVDD_TRY(launch_conv_gemm(a, h->w5[i], g, EPI_GENERIC, ep, st));
        if (debug_dump_enabled()) {
          debug_dump(("ef_z" + std::to_string(i)).c_str(), x, (size_t)rows * len * fi * 2, st);
          debug_dump(("ef_ap" + std::to_string(i)).c_str(), y, (size_t)rows * len * fi * 2, st);
        }
        __nv_bfloat16* t = a; a = y; y = t;          // a <- a' ; old a is free (now y)
      }
      // does the NEXT stage take the padded layout?  (its input length lo must be even and long
      // enough for the 4 pad rows to be cheap; the debug dumps keep the dense layout)
      const bool halo_next = halo_enabled() && pair_path && i + 1 < h->n_stage && lo % 2 == 0 && lo >= 32 &&
                             h->f[i + 1] % 128 == 0 && !debug_dump_enabled();
      // Attention pooling needs only the difference of the two logits of a pair, and
      // Wp.y[2j+1] - Wp.y[2j] = Wp.(y[2j+1] - y[2j]): the 1x1 conv's epilogue emits y0 = y[2j] and
      // yd = y[2j+1] - y[2j] at half length (EPI_PAIR) and the pool is ONE GEMM over yd (EPI_POOL2)
      // instead of two over y.  Odd lengths: whole-sequence tiles padded to an even BL, yd = 0 for
      // the unpaired tail (= the reference's -inf logit on the padded slot).
      if (pair_path) {
        const int len_p = halo ? Lp : len;           // rows per sequence in the layout of x / a'
        const int lo_p = len_p / 2 + (len_p & 1);    // ... and of y0 / yd
        __nv_bfloat16* y0 = y;
        __nv_bfloat16* yd = y + (size_t)rows * lo_p * fi;
        {
          GemmShape g;
          g.K = fi; g.N = fi; g.taps = 1; g.dil = 1;
          if (len_p % 2 == 0) {
            g.S = 1; g.L = (int)(rows * len_p); g.L_in = g.L; g.BL = 128; g.BS = 1;
          } else {
            g.S = (int)rows; g.L = len; g.L_in = len; g.BL = len + 1; g.BS = 128 / (len + 1);
          }
          g.useful_rows = rows * len;
          EpiParams ep;
          if (i == 0 && stem_concat) {
            g.K2 = kStemK;
            ep.a2 = b.col; ep.w2k = h->stem_w;
            ep.bias = h->b1s;
          } else {
            ep.bias = h->b1[i];
            ep.res = x; ep.res_dtype = DT_BF16; ep.ld_res = fi;
          }
          ep.out = y0; ep.out_dtype = DT_BF16; ep.ld_out = fi;
          ep.out2 = yd; ep.out2_dtype = DT_BF16; ep.ld_out2 = fi;
          g.reverse = snake & (li++ & 1);
          SVDD_TRY(launch_conv_gemm(a, h->w1[i], g, EPI_PAIR, ep, st));
          if (debug_dump_enabled()) {
            debug_dump(("ef_y0_" + std::to_string(i)).c_str(), y0, (size_t)rows * lo * fi * 2, st);
            debug_dump(("ef_yd_" + std::to_string(i)).c_str(), yd, (size_t)rows * lo * fi * 2, st);
          }
        }
        {
          GemmShape g;
          g.K = fi; g.N = fi; g.taps = 1; g.dil = 1;
          EpiParams ep;
          const __nv_bfloat16* y0v = y0;
          const __nv_bfloat16* ydv = yd;
          if (halo || halo_next) {
            // per-sequence tiles: the inputs skip the pad pair of a padded stage, the output lands
            // between the zero rows of the next stage's padded operand
            g.S = (int)rows; g.L = lo; g.L_in = lo;
            choose_row_tiling(lo, 1, &g);
            if (halo) {
              y0v += (size_t)fi; ydv += (size_t)fi;              // valid pairs start at row 1 of lo_p
              g.a_pitch = lo_p; ep.res_pitch = lo_p; ep.res2_pitch = lo_p;
            }
          } else {
            g.S = 1; g.L = (int)(rows * lo); g.L_in = g.L; g.BL = 128; g.BS = 1;
          }
          ep.res = y0v; ep.res_dtype = DT_BF16; ep.ld_res = fi;
          ep.res2 = ydv; ep.ld_res2 = fi;
          if (i + 1 < h->n_stage) {
            ep.out2 = a; ep.out2_dtype = DT_BF16; ep.ld_out2 = fi;
            ep.scale2 = h->bn5_s[i + 1]; ep.shift2 = h->bn5_t[i + 1]; ep.act2 = ACT_GELU;
            if (halo_next) {
              const unsigned zg = (unsigned)ceil_div<int64_t>(rows * 4 * (fi / 8), 256);
              launch_k(ef_zero_pads_kernel, dim3(zg), dim3(256), 0, st, 1, a, rows, lo + 4, fi);
              count_launch();
              SVDD_LAUNCH_CHECK();
              ep.out2 = a + (size_t)2 * fi;
              ep.out2_pitch = lo + 4;
            }
          } else {
            ep.out = b.xt; ep.out_dtype = DT_F32; ep.ld_out = fi;
          }
          g.reverse = snake & (li++ & 1);
          SVDD_TRY(launch_conv_gemm(ydv, h->wp[i], g, EPI_POOL2, ep, st));
        }
      } else {
        {  // y = Conv_1x1(a) + b + x     (residual ConvBlock, Enformer.py:1836-1838 / 1866-1868)
          EpiParams ep;
          ep.bias = h->b1[i];
          ep.res = x; ep.res_dtype = DT_BF16; ep.ld_res = fi;
          ep.out = y; ep.out_dtype = DT_BF16; ep.ld_out = fi;
          SVDD_TRY(gemm_flat(a, h->w1[i], rows * len, fi, fi, ep, EPI_GENERIC));
          if (debug_dump_enabled())
            debug_dump(("ef_y" + std::to_string(i)).c_str(), y, (size_t)rows * len * fi * 2, st);
        }
        {  // attention pool over position pairs; epilogue also prepares the next conv's operand
          GemmShape g;
          g.S = (int)rows; g.L = lo; g.L_in = len; g.K = fi; g.N = fi; g.taps = 1; g.dil = 1;
          choose_row_tiling(lo, 1, &g);
          EpiParams ep;
          ep.pool_vals = y;
          if (i + 1 < h->n_stage) {
            ep.out2 = a; ep.out2_dtype = DT_BF16; ep.ld_out2 = fi;
            ep.scale2 = h->bn5_s[i + 1]; ep.shift2 = h->bn5_t[i + 1]; ep.act2 = ACT_GELU;
          } else {
            ep.out = b.xt; ep.out_dtype = DT_F32; ep.ld_out = fi;
          }
          SVDD_TRY(launch_conv_gemm(y, h->wp[i], g, EPI_POOL, ep, st));
        }
      }
      if (debug_dump_enabled()) {
        if (i + 1 < h->n_stage)
          debug_dump(("ef_a" + std::to_string(i + 1)).c_str(), a, (size_t)rows * lo * fi * 2, st);
        else
          debug_dump("ef_xt_in", b.xt, (size_t)rows * lo * fi * 4, st);
      }
      halo_prev = halo_next;
      len = lo;
    }
    // ---- transformer tower ------------------------------------------------------------------
    const int64_t R = rows * n;
    const bool fused_tower = tower_usable(h, n);
    if (fused_tower) SVDD_TRY(launch_tower(h, b, R, n, st));
    for (int j = 0; j < h->n_blocks && !fused_tower; ++j) {
      auto& blk = h->blk[j];
      SVDD_TRY(launch_ln(b.xt, blk.ln1_g, blk.ln1_b, b.hn, R, C, st));
      {
        EpiParams ep;
        ep.out = b.qkv; ep.out_dtype = DT_F32; ep.ld_out = nqkv;
        SVDD_TRY(gemm_flat(b.hn, blk.wqkv, R, C, nqkv, ep, EPI_GENERIC));
      }
      SVDD_TRY(launch_attention(b.qkv, blk.rcb, blk.rpb, h->relk + (size_t)j * H * P * dk, b.ao, rows, n, H,
                                dk, dv, st));
      {
        EpiParams ep;                                // x += to_out(attn)   (:1941-1945)
        ep.bias = blk.bo;
        ep.res = b.xt; ep.res_dtype = DT_F32; ep.ld_res = C;
        ep.out = b.xt; ep.out_dtype = DT_F32; ep.ld_out = C;
        SVDD_TRY(gemm_flat(b.ao, blk.wo, R, H * dv, C, ep, EPI_GENERIC));
      }
      if (debug_dump_enabled()) {
        debug_dump(("ef_qkv" + std::to_string(j)).c_str(), b.qkv, (size_t)R * nqkv * 4, st);
        debug_dump(("ef_ao" + std::to_string(j)).c_str(), b.ao, (size_t)R * H * dv * 2, st);
        debug_dump(("ef_xattn" + std::to_string(j)).c_str(), b.xt, (size_t)R * C * 4, st);
      }
      SVDD_TRY(launch_ln(b.xt, blk.ln2_g, blk.ln2_b, b.hn, R, C, st));
      {
        EpiParams ep;                                // ReLU(Linear(C, 2C))
        ep.bias = blk.bf1; ep.act = ACT_RELU;
        ep.out = b.u; ep.out_dtype = DT_BF16; ep.ld_out = 2 * C;
        SVDD_TRY(gemm_flat(b.hn, blk.wf1, R, C, 2 * C, ep, EPI_GENERIC));
      }
      {
        EpiParams ep;                                // x += Linear(2C, C)   (:1946-1948)
        ep.bias = blk.bf2;
        ep.res = b.xt; ep.res_dtype = DT_F32; ep.ld_res = C;
        ep.out = b.xt; ep.out_dtype = DT_F32; ep.ld_out = C;
        if (j + 1 == h->n_blocks) {                  // pointwise ConvBlock's BN+GELU operand
          ep.out2 = b.hn; ep.out2_dtype = DT_BF16; ep.ld_out2 = C;
          ep.scale2 = h->bnpw_s; ep.shift2 = h->bnpw_t; ep.act2 = ACT_GELU;
        }
        SVDD_TRY(gemm_flat(b.u, blk.wf2, R, 2 * C, C, ep, EPI_GENERIC));
      }
      if (debug_dump_enabled())
        debug_dump(("ef_xt" + std::to_string(j)).c_str(), b.xt, (size_t)R * C * 4, st);
    }
    if (h->n_blocks == 0) {
      set_last_error("svdd_enformer_score: trunk without transformer blocks is not supported");
      return SVDD_ERR_INVALID_ARGUMENT;
    }
    // ---- pointwise conv -> GELU -> head 1x1 -> mean over positions ----------------------------
    int n_tiles = 1;
    {
      GemmShape g;
      g.S = 1; g.L = (int)R; g.L_in = (int)R; g.K = C; g.N = 2 * C; g.BL = 128; g.BS = 1;
      EpiParams ep;
      ep.bias = h->bpw; ep.act = ACT_GELU;
      ep.head_w = h->hw; ep.partials = b.partials;
      n_tiles = conv_gemm_n_tiles(g, EPI_HEADDOT);
      SVDD_TRY(launch_conv_gemm(b.hn, h->wpw, g, EPI_HEADDOT, ep, st));
    }
    launch_k(ef_mean_kernel, dim3((unsigned)ceil_div<int64_t>(rows, 128)), dim3(128), 0, st, 1, b.partials, n_tiles, h->hb,
                                                                           scores + r0, rows, n);
    count_launch();
    SVDD_LAUNCH_CHECK();
  }
  return SVDD_OK;
}
