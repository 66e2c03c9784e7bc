// Stage 2 (SUBS + q_xs + Gumbel-max + carry-over), stage 4 (selection + gather)
// and the post-SUBS argmax.  HBM-bound byte/fp32 work: no tensor cores.
//
// Bit-exactness contract: every fp32 operation below is the same IEEE
// operation, in the same association, as the reference's torch expression
// (cited per line); transcendental calls are the accurate libdevice logf/expf
// that ATen's CUDA kernels call.  Explicit __f*_rn intrinsics stop nvcc from
// contracting mul+add into FMA.
#include "common.cuh"
#include "philox.cuh"

namespace svdd {
namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

// Post-SUBS probabilities of one position.
//   raw logits  : diffusion_gosai.py:289-294  logits[...,4] += -1e6; logits -= logsumexp
//   logsumexp   : ATen logsumexp = log(sum(exp(x - max))) + max
//   unmasked row: diffusion_gosai.py:300-303  -1e6 everywhere, 0 at x
__device__ __forceinline__ void subs_log_p(const float* lg, int tok, bool is_log_p,
                                           float* logp) {
  if (is_log_p) {
#pragma unroll
    for (int v = 0; v < kVocab; ++v) logp[v] = lg[v];
    return;
  }
  if (tok != kMaskIndex) {
#pragma unroll
    for (int v = 0; v < kVocab; ++v) logp[v] = (v == tok) ? 0.0f : kNegInfinity;
    return;
  }
  float l[kVocab];
#pragma unroll
  for (int v = 0; v < kVocab; ++v) l[v] = lg[v];
  l[kMaskIndex] = __fadd_rn(l[kMaskIndex], kNegInfinity);
  float mx = l[0];
#pragma unroll
  for (int v = 1; v < kVocab; ++v) mx = fmaxf(mx, l[v]);
  const float mx_used = (fabsf(mx) == INFINITY) ? 0.0f : mx;
  float s = 0.0f;
#pragma unroll
  for (int v = 0; v < kVocab; ++v) s = __fadd_rn(s, expf(__fsub_rn(l[v], mx_used)));
  const float lse = __fadd_rn(logf(s), mx_used);
#pragma unroll
  for (int v = 0; v < kVocab; ++v) logp[v] = __fsub_rn(l[v], lse);
}

// argmax_v q[v] / (1e-10 - log(u[v] + 1e-10)), first index on ties
// (diffusion_gosai.py:30-34).
__device__ __forceinline__ int gumbel_argmax5(const float* q, const float* u) {
  int best = 0;
  float best_key = 0.0f;
#pragma unroll
  for (int v = 0; v < kVocab; ++v) {
    const float g = __fsub_rn(1e-10f, logf(__fadd_rn(u[v], 1e-10f)));
    const float key = __fdiv_rn(q[v], g);
    if (v == 0 || key > best_key) { best_key = key; best = v; }
  }
  return best;
}

// The same argmax decided with one MUFU log per element and no division, when that is
// PROVABLY the exact answer; otherwise falls back to gumbel_argmax5.
//   g~ = 1e-10 - __logf(u + 1e-10) differs from the exact-path g by at most
//   e = 8e-7 * max(1, g~): __logf is within 2^-21.41 absolute on [0.5, 2] and 3 ulp
//   elsewhere (CUDA C Programming Guide, intrinsic error table), the exact path's logf and
//   subtraction within 1.5 ulp.  Candidate w is certainly the exact argmax when, for every
//   v != w,  q[w] * (g~[v] - e[v]) > q[v] * (g~[w] + e[w]) * (1 + 2^-19)
//   (cross-multiplied form of key[w] > key[v]; the last factor covers the roundings of the
//   two products and of the exact path's division).  Exact ties and near-ties fail the test
//   and take the exact path, so first-index tie-breaking is the reference's.
//   The bound is applied in the looser one-FMA form  e <= 1e-6 * (g~ + 1):
//   lo = g~ * (1 - 1e-6) - 1e-6,  hi = g~ * (1 + 1e-6) + 1e-6  (the extra 2e-7 * (g~ + 1) covers the
//   rounding of the FMA and of its constants).  A non-positive lo[v] fails the product test by
//   itself (q >= 0), so only the winner's lo is tested explicitly.  u + 1e-10 >= 1e-10 is a normal
//   number: MUFU.LG2 is used directly, without __logf's denormal pre-scaling (same value).
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ int gumbel_argmax5_checked(const float* q, const float* u) {
  float lo[kVocab], g[kVocab];
#pragma unroll
  for (int v = 0; v < kVocab; ++v) {
    g[v] = fmaf(lg2_approx(__fadd_rn(u[v], 1e-10f)), -0.693147182f, 1e-10f);   // = 1e-10 - __logf(u + 1e-10)
    lo[v] = fmaf(g[v], 0.999999f, -1e-6f);
  }
  int best = 0;
  float bq = q[0], bg = g[0], blo = lo[0];
#pragma unroll
  for (int v = 1; v < kVocab; ++v) {
    if (q[v] * bg > bq * g[v]) { best = v; bq = q[v]; bg = g[v]; blo = lo[v]; }
  }
  const float rhs_scale = fmaf(bg, 1.000001f, 1e-6f) * 1.0000020f;
  bool ok = (blo > 0.0f) && (bq > 1e-20f);   // the latter keeps both products far from the subnormal range
#pragma unroll
  for (int v = 0; v < kVocab; ++v) ok = ok && (v == best || bq * lo[v] > q[v] * rhs_scale);
  return ok ? best : gumbel_argmax5(q, u);
}

// One warp owns 32 positions and loops over the M candidates (q is computed once per
// position: the accurate expf/logf of the SUBS step are a third of the kernel's instructions,
// and the kernel is issue-bound, not DRAM-bound, once enough loads are in flight).  The noise
// of kNoiseGroup candidates (kNoiseGroup x 640 B per warp) is in flight while the previous
// group's draws are computed from shared memory.
constexpr int kNoiseGroup = 4;

template <typename Tok, bool kInjected, bool kFast>
__global__ void __launch_bounds__(kThreads)
subs_sample_kernel(const float* __restrict__ logits, int is_log_p,
                   const Tok* __restrict__ x, const float* __restrict__ U,
                   uint64_t seed, const uint64_t* __restrict__ seed_dev, uint32_t step,
                   int64_t row_offset,
                   float mc_t, float mc_s, Tok* __restrict__ cand,
                   float* __restrict__ q_out, int64_t BL, int L, int M) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_stage[kWarpsPerBlock][kNoiseGroup][32 * kVocab];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pos0 = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * 32;
  if (pos0 >= BL) return;
  const int64_t pos = pos0 + lane;
  const bool valid = pos < BL;
  const bool full = pos0 + 32 <= BL;           // warp-uniform: no bounds checks on the fast path
  const int64_t n_el = BL * kVocab;
  float* stage = s_stage[warp][0];

  // coalesced load of this warp's 32x5 logits (transposed through smem below)
  const int64_t e0 = pos0 * kVocab;
  float lraw[kVocab];
#pragma unroll
  for (int k = 0; k < kVocab; ++k) {
    const int64_t e = e0 + k * 32 + lane;
    lraw[k] = (full || e < n_el) ? __ldg(logits + e) : 0.0f;
  }
  const int tok = valid ? load_tok(x, pos) : 0;
  const bool masked = valid && tok == kMaskIndex;
  // carry-over: an unmasked token is kept whatever the draw (:1199-1203), so a warp without
  // masked positions never touches the noise.
  const unsigned any_masked = __ballot_sync(0xffffffffu, masked);

  float pre[kNoiseGroup][kVocab];
  const float* up = U + e0 + lane;             // candidate m: up + m * n_el (+ k * 32)
  auto prefetch = [&](int m0) {
#pragma unroll
    for (int gi = 0; gi < kNoiseGroup; ++gi) {
      if (m0 + gi < M) {
        const float* p = up + (int64_t)(m0 + gi) * n_el;
#pragma unroll
        for (int k = 0; k < kVocab; ++k)
          pre[gi][k] = (full || e0 + k * 32 + lane < n_el) ? __ldcs(p + k * 32) : 0.5f;
      }
    }
  };
  if (kInjected && any_masked != 0u) prefetch(0);

#pragma unroll
  for (int k = 0; k < kVocab; ++k) stage[k * 32 + lane] = lraw[k];
  __syncwarp();
  float lg[kVocab];
#pragma unroll
  for (int v = 0; v < kVocab; ++v) lg[v] = stage[lane * kVocab + v];
  __syncwarp();

  // q_xs (diffusion_gosai.py:1194-1196)
  float logp[kVocab], q[kVocab];
  subs_log_p(lg, tok, is_log_p != 0, logp);
  const float d = __fsub_rn(mc_t, mc_s);
#pragma unroll
  for (int v = 0; v < kVocab; ++v) q[v] = __fmul_rn(expf(logp[v]), d);
  q[kMaskIndex] = mc_s;
  if (q_out != nullptr) {
#pragma unroll
    for (int v = 0; v < kVocab; ++v) stage[lane * kVocab + v] = q[v];
    __syncwarp();
#pragma unroll
    for (int k = 0; k < kVocab; ++k) {
      const int64_t e = e0 + k * 32 + lane;
      if (e < n_el) q_out[e] = stage[k * 32 + lane];
    }
    __syncwarp();
  }

  Tok* cp = cand + pos;                        // candidate m: cp + m * BL
  if (any_masked == 0u) {
    if (valid)
      for (int m = 0; m < M; ++m) cp[(size_t)m * BL] = (Tok)tok;
    return;
  }

  if (kInjected) {
    for (int m0 = 0; m0 < M; m0 += kNoiseGroup) {
#pragma unroll
      for (int gi = 0; gi < kNoiseGroup; ++gi)
#pragma unroll
        for (int k = 0; k < kVocab; ++k) s_stage[warp][gi][k * 32 + lane] = pre[gi][k];
      __syncwarp();
      prefetch(m0 + kNoiseGroup);              // next group's loads fly during the draws below
#pragma unroll
      for (int gi = 0; gi < kNoiseGroup; ++gi) {
        const int m = m0 + gi;
        if (m < M && valid) {
          int draw = tok;
          if (masked) {
            float u[kVocab];
#pragma unroll
            for (int v = 0; v < kVocab; ++v) u[v] = s_stage[warp][gi][lane * kVocab + v];
            draw = kFast ? gumbel_argmax5_checked(q, u) : gumbel_argmax5(q, u);
          }
          cp[(size_t)m * BL] = (Tok)draw;
        }
      }
      __syncwarp();
    }
  } else {
    // seed_dev (optional, device): [0] is added to the key, [1] to the row offset -- both live in
    // device memory so that ONE captured CUDA graph serves every (run, batch position)
    const uint64_t key = seed + (seed_dev != nullptr ? seed_dev[0] : 0ull);
    const uint32_t key0 = (uint32_t)key, key1 = (uint32_t)(key >> 32);
    const uint32_t row = (uint32_t)(row_offset + (seed_dev != nullptr ? (int64_t)seed_dev[1] : 0) + pos / L);
    const uint32_t l = (uint32_t)(pos % L);
    for (int m = 0; m < M; ++m) {
      if (!valid) continue;
      int draw = tok;
      if (masked) {
        const Philox4 a = philox4x32_10(l, row, (uint32_t)m, step, key0, key1);
        const Philox4 e = philox4x32_10(l, row, (uint32_t)m | (1u << 16), step, key0, key1);
        const float u[kVocab] = {philox_uniform(a.x), philox_uniform(a.y),
                                 philox_uniform(a.z), philox_uniform(a.w),
                                 philox_uniform(e.x)};
        draw = kFast ? gumbel_argmax5_checked(q, u) : gumbel_argmax5(q, u);
      }
      cp[(size_t)m * BL] = (Tok)draw;
    }
  }
}

// ---- post-SUBS argmax over the 4 real tokens --------------------------------
template <typename Tok>
__global__ void __launch_bounds__(kThreads)
x0_argmax_kernel(const float* __restrict__ logits, const Tok* __restrict__ x,
                 Tok* __restrict__ out, int64_t NL) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_stage[kWarpsPerBlock][32 * kVocab];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pos0 = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * 32;
  if (pos0 >= NL) return;
  const int64_t pos = pos0 + lane;
  float* stage = s_stage[warp];
  const int64_t n_el = NL * kVocab, e0 = pos0 * kVocab;
#pragma unroll
  for (int k = 0; k < kVocab; ++k) {
    const int64_t e = e0 + k * 32 + lane;
    stage[k * 32 + lane] = (e < n_el) ? __ldg(logits + e) : 0.0f;
  }
  __syncwarp();
  if (pos >= NL) return;
  float lg[kVocab], logp[kVocab];
#pragma unroll
  for (int v = 0; v < kVocab; ++v) lg[v] = stage[lane * kVocab + v];
  const int tok = load_tok(x, pos);
  subs_log_p(lg, tok, false, logp);
  int best = 0;
#pragma unroll
  for (int v = 1; v < kMaskIndex; ++v)
    if (logp[v] > logp[best]) best = v;
  store_tok(out, pos, best);
}

// ---- post-SUBS log-probabilities (the tensor Diffusion.forward returns) ----------
template <typename Tok>
__global__ void __launch_bounds__(kThreads)
subs_log_p_kernel(const float* __restrict__ logits, const Tok* __restrict__ x,
                  float* __restrict__ out, int64_t NL) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_stage[kWarpsPerBlock][32 * kVocab];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t pos0 = ((int64_t)blockIdx.x * kWarpsPerBlock + warp) * 32;
  if (pos0 >= NL) return;
  const int64_t pos = pos0 + lane;
  float* stage = s_stage[warp];
  const int64_t n_el = NL * kVocab, e0 = pos0 * kVocab;
#pragma unroll
  for (int k = 0; k < kVocab; ++k) {
    const int64_t e = e0 + k * 32 + lane;
    stage[k * 32 + lane] = (e < n_el) ? __ldg(logits + e) : 0.0f;
  }
  __syncwarp();
  float lg[kVocab], logp[kVocab];
#pragma unroll
  for (int v = 0; v < kVocab; ++v) lg[v] = stage[lane * kVocab + v];
  __syncwarp();
  const int tok = pos < NL ? load_tok(x, pos) : 0;
  subs_log_p(lg, tok, false, logp);
#pragma unroll
  for (int v = 0; v < kVocab; ++v) stage[lane * kVocab + v] = logp[v];
  __syncwarp();
#pragma unroll
  for (int k = 0; k < kVocab; ++k) {
    const int64_t e = e0 + k * 32 + lane;
    if (e < n_el) out[e] = stage[k * 32 + lane];
  }
}

// ---- stage 4 -----------------------------------------------------------------
// One warp per sequence.  The max / sum reductions use the same butterfly
// (lanes = next_pow2(M) capped at 32, element i on lane i % lanes, xor
// shuffles from lanes/2 down to 1) as ATen's softmax_warp_forward so that the
// softmax values -- and hence first-index argmax ties -- are those of
// torch.softmax(scores, dim=1) on the same device (diffusion_gosai.py:1220,1225).
constexpr int kMaxIterSel = 8;  // M <= 256

template <typename Tok, bool kInjected>
__global__ void __launch_bounds__(kThreads)
select_gather_kernel(const float* __restrict__ scores, const Tok* __restrict__ cand,
                     float alpha, const float* __restrict__ U_sel, uint64_t seed,
                     const uint64_t* __restrict__ seed_dev, uint32_t step, int64_t row_offset,
                     Tok* __restrict__ x_out, int32_t* __restrict__ idx_out, int B,
                     int L, int M, int lanes, int iters) {
  pdl_wait();
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpsPerBlock + warp;
  if (b >= B) return;
  const bool active = lane < lanes;
  float e[kMaxIterSel];
  float mx = -INFINITY;
#pragma unroll
  for (int it = 0; it < kMaxIterSel; ++it) {
    if (it < iters) {
      const int m = lane + it * lanes;
      float s = (active && m < M) ? __ldg(scores + (size_t)m * B + b) : -INFINITY;
      if (alpha > 0.0f && s != -INFINITY) s = __fdiv_rn(s, alpha);
      e[it] = s;
      mx = fmaxf(mx, s);
    }
  }
  for (int off = lanes >> 1; off > 0; off >>= 1)
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float sum = 0.0f;
#pragma unroll
  for (int it = 0; it < kMaxIterSel; ++it) {
    if (it < iters) {
      const float v = (e[it] == -INFINITY) ? 0.0f : expf(__fsub_rn(e[it], mx));
      e[it] = v;
      sum = __fadd_rn(sum, v);
    }
  }
  for (int off = lanes >> 1; off > 0; off >>= 1)
    sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, off));

  // key per candidate: softmax value (alpha == 0) or Gumbel-max key
  float best_key = -INFINITY;
  int best_m = 0x7fffffff;
#pragma unroll
  for (int it = 0; it < kMaxIterSel; ++it) {
    if (it < iters) {
      const int m = lane + it * lanes;
      if (active && m < M) {
        float key = __fdiv_rn(e[it], sum);
        if (alpha > 0.0f) {
          float u;
          if (kInjected) {
            u = __ldg(U_sel + (size_t)b * M + m);
          } else {
            const uint64_t key = seed + (seed_dev != nullptr ? seed_dev[0] : 0ull);
            const uint32_t key0 = (uint32_t)key, key1 = (uint32_t)(key >> 32);
            const int64_t row0 = row_offset + (seed_dev != nullptr ? (int64_t)seed_dev[1] : 0);
            const Philox4 w = philox4x32_10((uint32_t)(m >> 2), (uint32_t)(row0 + b),
                                            0u, step | (1u << 24), key0, key1);
            const uint32_t words[4] = {w.x, w.y, w.z, w.w};
            u = philox_uniform(words[m & 3]);
          }
          const float g = __fsub_rn(1e-10f, logf(__fadd_rn(u, 1e-10f)));
          key = __fdiv_rn(key, g);
        }
        if (key > best_key || (key == best_key && m < best_m)) { best_key = key; best_m = m; }
      }
    }
  }
  for (int off = 16; off > 0; off >>= 1) {
    const float ok = __shfl_xor_sync(0xffffffffu, best_key, off);
    const int om = __shfl_xor_sync(0xffffffffu, best_m, off);
    if (ok > best_key || (ok == best_key && om < best_m)) { best_key = ok; best_m = om; }
  }
  if (best_m == 0x7fffffff) best_m = 0;  // all keys NaN / -inf: torch.argmax -> 0
  if (lane == 0 && idx_out != nullptr) idx_out[b] = best_m;
  const Tok* src = cand + ((size_t)best_m * B + b) * L;
  Tok* dst = x_out + (size_t)b * L;
  const size_t bytes = (size_t)L * sizeof(Tok);
  if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0) {
    // 16-byte copies, four in flight per lane before the first store
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    const int n = (int)(bytes >> 4);
    for (int i0 = lane; i0 < n; i0 += 128) {
      uint4 buf[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + 32 * j < n) buf[j] = __ldcs(s4 + i0 + 32 * j);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (i0 + 32 * j < n) d4[i0 + 32 * j] = buf[j];
    }
  } else {
    for (int l = lane; l < L; l += 32) dst[l] = src[l];
  }
}

int check_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_last_error("no CUDA device");
    return SVDD_ERR_CUDA;
  }
  return svdd_device_check(dev);
}

}  // namespace
}  // namespace svdd

using namespace svdd;

namespace {
int subs_sample_impl(const float* logits, int is_log_p, const void* x, int tok_dtype, const float* U,
                     uint64_t seed, const uint64_t* seed_dev, int step, int64_t row_offset, float mc_t,
                     float mc_s, void* cand, float* q_out, int B, int L, int M, bool fast, void* stream) {
  SVDD_CHECK_ARG(B >= 0 && L >= 0 && M >= 1, "svdd_subs_sample: bad shape B=%d L=%d M=%d", B, L, M);
  if ((int64_t)B * L == 0) return SVDD_OK;   // empty batch: nothing to do (pointers may be null)
  SVDD_CHECK_ARG(logits && x && cand, "svdd_subs_sample: null pointer");
  SVDD_CHECK_ARG(M < (1 << 16), "svdd_subs_sample: M must be < 65536");
  SVDD_CHECK_ARG(tok_dtype == SVDD_TOK_I64 || tok_dtype == SVDD_TOK_U8, "bad tok_dtype %d", tok_dtype);
  SVDD_TRY(check_device());
  const int64_t BL = (int64_t)B * L;
  const unsigned grid = (unsigned)ceil_div<int64_t>(BL, kThreads);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(Tok, INJ, FAST)                                                        \
  launch_k((subs_sample_kernel<Tok, INJ, FAST>), dim3(grid), dim3(kThreads), 0, st, 1,                       \
      logits, is_log_p, (const Tok*)x, U, seed, seed_dev, (uint32_t)step, row_offset, \
      mc_t, mc_s, (Tok*)cand, q_out, BL, L, M)
#define LAUNCH2(Tok, INJ) do { if (fast) LAUNCH(Tok, INJ, true); else LAUNCH(Tok, INJ, false); } while (0)
  if (tok_dtype == SVDD_TOK_I64) { if (U) LAUNCH2(int64_t, true); else LAUNCH2(int64_t, false); }
  else                           { if (U) LAUNCH2(uint8_t, true); else LAUNCH2(uint8_t, false); }
#undef LAUNCH2
#undef LAUNCH
  count_launch();
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}
}  // namespace

extern "C" int svdd_subs_sample(const float* logits, int is_log_p, const void* x,
                                int tok_dtype, const float* U, uint64_t seed,
                                const uint64_t* seed_dev, int step, int64_t row_offset,
                                float mc_t, float mc_s, void* cand,
                                float* q_out, int B, int L, int M, void* stream) {
  return subs_sample_impl(logits, is_log_p, x, tok_dtype, U, seed, seed_dev, step, row_offset, mc_t, mc_s,
                          cand, q_out, B, L, M, /*fast=*/true, stream);
}

// Test hook: the same stage with every draw taken on the exact (logf + division) path, to
// check that the verified fast path of svdd_subs_sample never changes a draw.
extern "C" int svdd_selftest_subs_sample_exact(const float* logits, int is_log_p, const void* x,
                                               int tok_dtype, const float* U, uint64_t seed,
                                               const uint64_t* seed_dev, int step, int64_t row_offset,
                                               float mc_t, float mc_s, void* cand, float* q_out, int B,
                                               int L, int M, void* stream) {
  return subs_sample_impl(logits, is_log_p, x, tok_dtype, U, seed, seed_dev, step, row_offset, mc_t, mc_s,
                          cand, q_out, B, L, M, /*fast=*/false, stream);
}

extern "C" int svdd_x0_argmax(const float* logits, const void* x, int tok_dtype, void* out,
                              int64_t n_rows, int L, void* stream) {
  SVDD_CHECK_ARG(n_rows >= 0 && L >= 0, "svdd_x0_argmax: bad shape");
  if (n_rows * L == 0) return SVDD_OK;
  SVDD_CHECK_ARG(logits && x && out, "svdd_x0_argmax: null pointer");
  SVDD_CHECK_ARG(tok_dtype == SVDD_TOK_I64 || tok_dtype == SVDD_TOK_U8, "bad tok_dtype %d", tok_dtype);
  SVDD_TRY(check_device());
  const int64_t NL = n_rows * L;
  if (NL == 0) return SVDD_OK;
  const unsigned grid = (unsigned)ceil_div<int64_t>(NL, kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  if (tok_dtype == SVDD_TOK_I64)
    launch_k(x0_argmax_kernel<int64_t>, dim3(grid), dim3(kThreads), 0, st, 1, logits, (const int64_t*)x, (int64_t*)out, NL);
  else
    launch_k(x0_argmax_kernel<uint8_t>, dim3(grid), dim3(kThreads), 0, st, 1, logits, (const uint8_t*)x, (uint8_t*)out, NL);
  count_launch();
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}

extern "C" int svdd_subs_log_p(const float* logits, const void* x, int tok_dtype, float* log_p,
                               int64_t n_rows, int L, void* stream) {
  SVDD_CHECK_ARG(n_rows >= 0 && L >= 0, "svdd_subs_log_p: bad shape");
  if (n_rows * L == 0) return SVDD_OK;
  SVDD_CHECK_ARG(logits && x && log_p, "svdd_subs_log_p: null pointer");
  SVDD_CHECK_ARG(tok_dtype == SVDD_TOK_I64 || tok_dtype == SVDD_TOK_U8, "bad tok_dtype %d", tok_dtype);
  SVDD_TRY(check_device());
  const int64_t NL = n_rows * L;
  if (NL == 0) return SVDD_OK;
  const unsigned grid = (unsigned)ceil_div<int64_t>(NL, kThreads);
  cudaStream_t st = (cudaStream_t)stream;
  if (tok_dtype == SVDD_TOK_I64)
    launch_k(subs_log_p_kernel<int64_t>, dim3(grid), dim3(kThreads), 0, st, 1, logits, (const int64_t*)x, log_p, NL);
  else
    launch_k(subs_log_p_kernel<uint8_t>, dim3(grid), dim3(kThreads), 0, st, 1, logits, (const uint8_t*)x, log_p, NL);
  count_launch();
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}

extern "C" int svdd_select_gather(const float* scores, const void* cand, int tok_dtype,
                                  float alpha, const float* U_sel, uint64_t seed,
                                  const uint64_t* seed_dev, int step, int64_t row_offset,
                                  void* x_out, int32_t* idx_out, int B,
                                  int L, int M, void* stream) {
  SVDD_CHECK_ARG(B >= 0 && L >= 0 && M >= 1, "svdd_select_gather: bad shape");
  if (B == 0) return SVDD_OK;
  SVDD_CHECK_ARG(scores && cand && x_out, "svdd_select_gather: null pointer");
  SVDD_CHECK_ARG(M <= 32 * kMaxIterSel, "svdd_select_gather: M=%d exceeds %d", M, 32 * kMaxIterSel);
  SVDD_CHECK_ARG(alpha >= 0.0f, "svdd_select_gather: alpha must be >= 0");
  SVDD_CHECK_ARG(tok_dtype == SVDD_TOK_I64 || tok_dtype == SVDD_TOK_U8, "bad tok_dtype %d", tok_dtype);
  SVDD_TRY(check_device());
  if (B == 0) return SVDD_OK;
  int pow2 = 1;
  while (pow2 < M) pow2 <<= 1;
  const int lanes = pow2 < 32 ? pow2 : 32;
  const int iters = pow2 / lanes;
  const unsigned grid = (unsigned)ceil_div(B, kWarpsPerBlock);
  cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH(Tok, INJ)                                                               \
  launch_k((select_gather_kernel<Tok, INJ>), dim3(grid), dim3(kThreads), 0, st, 1,                            \
      scores, (const Tok*)cand, alpha, U_sel, seed, seed_dev, (uint32_t)step,         \
      row_offset, (Tok*)x_out, idx_out, B, L, M, lanes, iters)
  const bool inj = U_sel != nullptr;
  if (tok_dtype == SVDD_TOK_I64) { if (inj) LAUNCH(int64_t, true); else LAUNCH(int64_t, false); }
  else                           { if (inj) LAUNCH(uint8_t, true); else LAUNCH(uint8_t, false); }
#undef LAUNCH
  count_launch();
  SVDD_LAUNCH_CHECK();
  return SVDD_OK;
}
