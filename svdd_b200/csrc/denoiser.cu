// Stage 1: the MDLM denoiser (reference: models/dnaconv.py:135-210, CNNModel.forward).
//
//   tokens -> Conv(5->128,k9)+ReLU                          den_embed_kernel (one-hot conv = weight gather)
//          -> 20 x [ +time-bias -> LayerNorm -> dilated Conv(128->128,k9) -> ReLU -> +residual ]
//                                                           conv_gemm (tcgen05) with the EPI_DEN_LN epilogue:
//                                                           the epilogue of layer i also produces layer i+1's
//                                                           LayerNorm'd bf16 operand, so each layer is ONE kernel
//          -> Conv1x1 + ReLU -> Conv1x1(128->5)             conv_gemm with EPI_DEN_FINAL
//
// HBM layout (workspace): residual stream `feat` fp32 [N*L,128]; operand buffers h0/h1 bf16
// [N*L,128] (ping-pong: layer i reads one through TMA while its epilogue writes the other).
#include <new>

#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "conv_gemm.cuh"
#include "den_fused.cuh"
#include "den_short.cuh"
#include "weights.cuh"

namespace svdd {
namespace {

constexpr int kH = 128;
constexpr int kTaps = 9;
constexpr int kMaxLayers = 64;
constexpr int kEmbedWarps = 8;
constexpr int kEmbedPosPerWarp = 8;

template <typename Tok>
__global__ void __launch_bounds__(kEmbedWarps * 32)
den_embed_kernel(const Tok* __restrict__ tokens, const float* __restrict__ w /*[9][5][128]*/,
                 const float* __restrict__ b, const float* __restrict__ tbias0,
                 const float* __restrict__ g0, const float* __restrict__ be0,
                 float* __restrict__ feat, __nv_bfloat16* __restrict__ h, int64_t NL, int L) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float s_w[kTaps * kVocab * kH];
  for (int i = threadIdx.x; i < kTaps * kVocab * kH; i += blockDim.x) s_w[i] = w[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = lane * 4;
  const float4 bias = *reinterpret_cast<const float4*>(b + c);
  const float4 tb = *reinterpret_cast<const float4*>(tbias0 + c);
  const float4 gg = *reinterpret_cast<const float4*>(g0 + c);
  const float4 bb = *reinterpret_cast<const float4*>(be0 + c);
  const int64_t pos_base = ((int64_t)blockIdx.x * kEmbedWarps + warp) * kEmbedPosPerWarp;
  for (int i = 0; i < kEmbedPosPerWarp; ++i) {
    const int64_t pos = pos_base + i;
    if (pos >= NL) break;
    const int l = (int)(pos % L);
    int tokv = -1;
    if (lane < kTaps) {
      const int li = l + lane - kTaps / 2;
      if (li >= 0 && li < L) tokv = load_tok(tokens, pos + lane - kTaps / 2);
    }
    float4 acc = bias;
#pragma unroll
    for (int t = 0; t < kTaps; ++t) {
      const int tk = __shfl_sync(0xffffffffu, tokv, t);
      if (tk >= 0) {
        const float4 ww = *reinterpret_cast<const float4*>(&s_w[(t * kVocab + tk) * kH + c]);
        acc.x += ww.x; acc.y += ww.y; acc.z += ww.z; acc.w += ww.w;
      }
    }
    float4 v = make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
    *reinterpret_cast<float4*>(feat + pos * kH + c) = v;
    // LayerNorm_0(feat + tbias_0)  (models/dnaconv.py:192-194)
    const float u0 = v.x + tb.x, u1 = v.y + tb.y, u2 = v.z + tb.z, u3 = v.w + tb.w;
    float sum = u0 + u1 + u2 + u3;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / kH);
    const float d0 = u0 - mean, d1 = u1 - mean, d2 = u2 - mean, d3 = u3 - mean;
    float sq = d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / kH) + 1e-5f);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(d0 * rstd * gg.x + bb.x, d1 * rstd * gg.y + bb.y);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(d2 * rstd * gg.z + bb.z, d3 * rstd * gg.w + bb.w);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&p0);
    packed.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(h + pos * kH + c) = packed;
  }
}

// linear.weight [128,5,9] -> [9][5][128]
__global__ void pack_embed_weight_kernel(const float* __restrict__ w, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= kTaps * kVocab * kH) return;
  const int c = i % kH, tok = (i / kH) % kVocab, t = i / (kH * kVocab);
  out[i] = w[(c * kVocab + tok) * kTaps + t];
}

}  // namespace
}  // namespace svdd

using namespace svdd;

struct svdd_denoiser {
  int n_layers = 0;
  int dil[kMaxLayers];
  DeviceArena arena;
  float* embed_w = nullptr;
  float* embed_b = nullptr;
  __nv_bfloat16* conv_w = nullptr;  // [n_layers][9][128][128]
  float* conv_b = nullptr;          // [n_layers][128]
  float* ln_g = nullptr;
  float* ln_b = nullptr;
  __nv_bfloat16* fc0_w = nullptr;   // [128][128]
  float* fc0_b = nullptr;
  float* fc2_w = nullptr;           // [5][128]
  float* fc2_b = nullptr;
};

extern "C" int svdd_denoiser_create(const svdd_tensor* tensors, int n_tensors, int num_cnn_stacks,
                                    void* stream, svdd_denoiser** out) {
  SVDD_CHECK_ARG(tensors && out && n_tensors > 0, "svdd_denoiser_create: null argument");
  SVDD_CHECK_ARG(num_cnn_stacks >= 1 && 5 * num_cnn_stacks <= kMaxLayers, "bad num_cnn_stacks %d",
                 num_cnn_stacks);
  int dev = 0;
  SVDD_CUDA(cudaGetDevice(&dev));
  SVDD_TRY(svdd_device_check(dev));
  cudaStream_t st = (cudaStream_t)stream;
  TensorTable tt{tensors, n_tensors};
  const int n = 5 * num_cnn_stacks;
  SVDD_CHECK_ARG(tt.dim("linear.weight", 0) == kH && tt.dim("linear.weight", 1) == kVocab &&
                     tt.dim("linear.weight", 2) == kTaps,
                 "svdd_denoiser_create: linear.weight must be [128,5,9] (hidden_dim 128 only)");
  svdd_denoiser* h = new (std::nothrow) svdd_denoiser();
  SVDD_CHECK_ARG(h != nullptr, "out of host memory");
  static const int groups[5] = {1, 1, 4, 16, 64};  // models/dnaconv.py:155-160
  h->n_layers = n;
  for (int i = 0; i < n; ++i) h->dil[i] = groups[i / num_cnn_stacks];

  DeviceArena& A = h->arena;
  A.reserve(sizeof(float) * kTaps * kVocab * kH);
  A.reserve(sizeof(float) * kH);
  A.reserve(sizeof(__nv_bfloat16) * (size_t)n * kTaps * kH * kH);
  A.reserve(sizeof(float) * n * kH);
  A.reserve(sizeof(float) * n * kH);
  A.reserve(sizeof(float) * n * kH);
  A.reserve(sizeof(__nv_bfloat16) * kH * kH);
  A.reserve(sizeof(float) * kH);
  A.reserve(sizeof(float) * kVocab * kH);
  A.reserve(sizeof(float) * 8);
  int rc = A.commit();
  if (rc != SVDD_OK) { delete h; return rc; }
  h->embed_w = A.take<float>(kTaps * kVocab * kH);
  h->embed_b = A.take<float>(kH);
  h->conv_w = A.take<__nv_bfloat16>((size_t)n * kTaps * kH * kH);
  h->conv_b = A.take<float>(n * kH);
  h->ln_g = A.take<float>(n * kH);
  h->ln_b = A.take<float>(n * kH);
  h->fc0_w = A.take<__nv_bfloat16>(kH * kH);
  h->fc0_b = A.take<float>(kH);
  h->fc2_w = A.take<float>(kVocab * kH);
  h->fc2_b = A.take<float>(8);

  auto fail = [&](int code) { delete h; return code; };
#define GET_OR_FAIL(var, name, numel)               \
  const float* var = tt.get((name), (numel));       \
  if (var == nullptr) return fail(SVDD_ERR_MISSING_TENSOR)
#define TRY_OR_FAIL(expr)                           \
  do { int _rc = (expr); if (_rc != SVDD_OK) return fail(_rc); } while (0)

  GET_OR_FAIL(lw, "linear.weight", kH * kVocab * kTaps);
  GET_OR_FAIL(lb, "linear.bias", kH);
  pack_embed_weight_kernel<<<ceil_div(kTaps * kVocab * kH, 256), 256, 0, st>>>(lw, h->embed_w);
  TRY_OR_FAIL(copy_f32(lb, h->embed_b, kH, st));
  for (int i = 0; i < n; ++i) {
    const std::string p = "convs." + std::to_string(i) + ".";
    const std::string q = "norms." + std::to_string(i) + ".";
    GET_OR_FAIL(cw, p + "weight", (int64_t)kH * kH * kTaps);
    GET_OR_FAIL(cb, p + "bias", kH);
    GET_OR_FAIL(ng, q + "weight", kH);
    GET_OR_FAIL(nb, q + "bias", kH);
    TRY_OR_FAIL(pack_conv_weight(cw, h->conv_w + (size_t)i * kTaps * kH * kH, kH, kH, kTaps, st));
    TRY_OR_FAIL(copy_f32(cb, h->conv_b + i * kH, kH, st));
    TRY_OR_FAIL(copy_f32(ng, h->ln_g + i * kH, kH, st));
    TRY_OR_FAIL(copy_f32(nb, h->ln_b + i * kH, kH, st));
  }
  GET_OR_FAIL(f0w, "final_conv.0.weight", kH * kH);
  GET_OR_FAIL(f0b, "final_conv.0.bias", kH);
  GET_OR_FAIL(f2w, "final_conv.2.weight", kVocab * kH);
  GET_OR_FAIL(f2b, "final_conv.2.bias", kVocab);
  TRY_OR_FAIL(pack_conv_weight(f0w, h->fc0_w, kH, kH, 1, st));
  TRY_OR_FAIL(copy_f32(f0b, h->fc0_b, kH, st));
  TRY_OR_FAIL(copy_f32(f2w, h->fc2_w, kVocab * kH, st));
  TRY_OR_FAIL(copy_f32(f2b, h->fc2_b, kVocab, st));
#undef GET_OR_FAIL
#undef TRY_OR_FAIL
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_last_error("svdd_denoiser_create: %s", cudaGetErrorString(e));
    return fail(SVDD_ERR_CUDA);
  }
  *out = h;
  return SVDD_OK;
}

extern "C" void svdd_denoiser_destroy(svdd_denoiser* h) { delete h; }

// The whole network in one persistent kernel (den_fused.cuh) when the sequence fits two row
// tiles and the padded operand fits shared memory; SVDD_ERR_INTERNAL = not handled here.
// SVDD_DEN_FUSED=0 forces the layer-by-layer path (kept as the cross-check in the tests).
static int den_fused_forward(svdd_denoiser* h, const void* tokens, int tok_dtype, const float* time_bias,
                             float* logits, int64_t n_rows, int L, cudaStream_t st) {
  const char* env = getenv("SVDD_DEN_FUSED");      // read per call: the tests flip it
  const int enabled = env ? atoi(env) : 1;
  if (!enabled || L > 256 || h->n_layers > denf::kMaxLayers) return SVDD_ERR_INTERNAL;
  denf::Args a = {};
  a.two_seq = (2 * L <= 128) ? 1 : 0;
  const char* env_split = getenv("SVDD_DEN_SPLIT");   // 0: one epilogue thread per row also for short sequences
  a.split = (a.two_seq && !(env_split && atoi(env_split) == 0)) ? 1 : 0;
  int pad_before = 0, max_end = 256;
  for (int i = 0; i < h->n_layers; ++i)
    for (int t = 0; t < kTaps; ++t)
      for (int m = 0; m < 2; ++m) {
        const int o = (t - kTaps / 2) * h->dil[i];
        if (!denf::tap_hits(m, o, L, a.two_seq)) continue;
        const int start = denf::tile_plane_row(m, o, a.two_seq);
        if (-start > pad_before) pad_before = -start;
        if (start + 128 > max_end) max_end = start + 128;
      }
  a.pad_before = pad_before;
  a.a_rows = (pad_before + max_end + 7) & ~7;
  a.iso_b = 128;
  int smem = denf::smem_bytes(a.a_rows);
  // Combined mode (two sequences per CTA, split epilogue): ONE 128-row MMA per tap for both
  // sequences on a second pair of planes where they sit 64 rows apart, for taps whose offset fits the
  // zero gap (|o| <= 64 - L); the other taps stay on the isolated planes.  L = 50: 173 instead of 281
  // MMA groups per item.  SVDD_DEN_CMB=0 (read per call) keeps the two-tile scheme for the A/B test.
  // Interleaved mode (two sequences per CTA, split epilogue): plane row 2p / 2p+1 = position p of
  // sequence A / B, a tap offset o is a row offset of 2 o, ONE 128-row MMA per tap serves both sequences
  // for every tap (L = 50: 141 MMA groups per item; combined mode 173, two tiles 281), one pair of planes,
  // accumulator + residual = 256 TMEM columns.  SVDD_DEN_ILV=0 (read per call) falls back to the
  // combined mode below (and SVDD_DEN_CMB=0 further to the two-tile scheme): the A/B and cross-checks.
  const char* env_ilv = getenv("SVDD_DEN_ILV");
  bool use_ilv = false;
  if (a.split && !(env_ilv && atoi(env_ilv) == 0)) {
    denf::Args c = a;
    int max_o = 0;
    for (int i = 0; i < h->n_layers; ++i)
      for (int t = 0; t < kTaps; ++t) {
        const int o = (t - kTaps / 2) * h->dil[i], ao = o < 0 ? -o : o;
        if (ao < L && ao > max_o) max_o = ao;
      }
    c.ilv = 1;
    c.pad_before = (2 * max_o + 7) & ~7;
    c.a_rows = c.pad_before + 128 + c.pad_before;
    const int smem_i = denf::smem_bytes(c.a_rows);
    if (smem_i <= 227 * 1024) { a = c; smem = smem_i; use_ilv = true; }
  }
  const char* env_cmb = getenv("SVDD_DEN_CMB");
  if (!use_ilv && a.split && !(env_cmb && atoi(env_cmb) == 0)) {
    denf::Args c = a;
    c.cmb = 1;
    c.cmb_max = 64 - L;
    c.pad_c = (c.cmb_max + 7) & ~7;
    c.c_rows = c.pad_c + 128 + c.pad_c;
    c.iso_b = ((2 * L + 7) & ~7) < 64 ? 64 : ((2 * L + 7) & ~7);
    int max_iso = 0;
    for (int i = 0; i <= h->n_layers; ++i) {
      c.iso[i] = 0;
      if (i == h->n_layers) break;
      for (int t = 0; t < kTaps; ++t) {
        const int o = (t - kTaps / 2) * h->dil[i], ao = o < 0 ? -o : o;
        if (ao < L && ao > c.cmb_max) { c.iso[i] = 1; if (ao > max_iso) max_iso = ao; }
      }
    }
    c.pad_before = (max_iso + 7) & ~7;
    // rows the VALID lanes of either sequence read; the garbage lanes of an isolated MMA may read up
    // to 128 rows further, into the memory that follows the plane (legal, never consumed)
    c.a_rows = max_iso > 0 ? ((c.pad_before + c.iso_b + L + max_iso + 7) & ~7) : 8;
    const int smem_c = denf::smem_bytes(c.a_rows, c.c_rows);
    if (smem_c <= 227 * 1024) { a = c; smem = smem_c; }
  }
  if (smem > 227 * 1024) return SVDD_ERR_INTERNAL;
  a.tokens = tokens;
  a.embed_w = h->embed_w; a.embed_b = h->embed_b;
  a.conv_b = h->conv_b; a.ln_g = h->ln_g; a.ln_b = h->ln_b;
  a.time_bias = time_bias;
  a.fc0_b = h->fc0_b; a.fc2_w = h->fc2_w; a.fc2_b = h->fc2_b;
  a.logits = logits;
  a.n_rows = n_rows; a.L = L; a.n_layers = h->n_layers;
  for (int i = 0; i < h->n_layers; ++i) a.dil[i] = h->dil[i];
  CUtensorMap tmW, tmW0;
  SVDD_TRY(encode_tmap_2d_bf16(&tmW, h->conv_w, kH, (uint64_t)h->n_layers * kTaps * kH, 64, kH));
  SVDD_TRY(encode_tmap_2d_bf16(&tmW0, h->fc0_w, kH, kH, 64, kH));
  const int64_t items = a.two_seq ? (n_rows + 1) / 2 : n_rows;
  const unsigned grid = (unsigned)(items < num_sms() ? items : num_sms());
  // Interleaved mode with TWO items (four sequences) in flight per CTA (csrc/den_short.cuh): while the
  // epilogue warps work on one item's round the tensor core runs the other's.  SVDD_DEN_PAIR=0 (read
  // per call) keeps one item per CTA (den_fused_kernel): the A/B and the cross-check in the tests.
  {
    // Measured (tools/ab_den_cmb.py): 51 200 x 50: 14.0-14.5 vs 19.2 ms; 999 x 33: 0.37 vs 0.47 ms; but 10 x 50:
    // 0.20 vs 0.14 ms -- with fewer items than SMs one item per CTA finishes sooner, so two items share a
    // CTA only when there are more items than SMs (SVDD_DEN_PAIR=1 / 0 forces either).
    const char* env_pair = getenv("SVDD_DEN_PAIR");
    const bool want_pair = env_pair ? atoi(env_pair) != 0 : items > num_sms();
    if (a.ilv && want_pair && dens::smem_bytes(a.pad_before) <= 227 * 1024) {
      const int smem2 = dens::smem_bytes(a.pad_before);
      const int64_t pairs = (items + 1) / 2;
      // SVDD_DEN_CG=2 (read per call): CTA pairs with cta_group::2 MMAs, each CTA staging half of every weight
      // tile (csrc/den_short.cuh).  Measured within +-2.5 % of independent CTAs (the default): kept as an
      // experiment with its cross-check in the tests.
      const char* env_cg = getenv("SVDD_DEN_CG");
      const int cg = env_cg && atoi(env_cg) == 2 && num_sms() >= 2 ? 2 : 1;
      const int64_t passes = (pairs + cg - 1) / cg;
      const int64_t max_groups = num_sms() / cg;
      unsigned grid2 = (unsigned)(cg * (passes < max_groups ? passes : max_groups));
      // SVDD_DEN_GRID (tuning aid): cap the number of CTAs, to tell L2 contention from per-SM limits
      if (const char* eg = getenv("SVDD_DEN_GRID")) { const int g = atoi(eg) / cg * cg; if (g > 0 && (unsigned)g < grid2) grid2 = (unsigned)g; }
      CUtensorMap tmWh, tmW0h;      // box = one CTA's share of a weight tile: 64 K x (128 / cg) output channels
      SVDD_TRY(encode_tmap_2d_bf16(&tmWh, h->conv_w, kH, (uint64_t)h->n_layers * kTaps * kH, 64, kH / cg));
      SVDD_TRY(encode_tmap_2d_bf16(&tmW0h, h->fc0_w, kH, kH, 64, kH / cg));
      auto launch2 = [&](auto kern) -> int {
        SVDD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        SVDD_CUDA(launch_k(kern, dim3(grid2), dim3(dens::kThreads), (size_t)smem2, st, cg, tmWh, tmW0h, a));
        return SVDD_OK;
      };
      // SVDD_DEN_TRACE=<csv path> (tuning aid, tools/den_trace.py): SM clock stamps of the hand-over points of
      // CTA 0's third pair; synchronises the stream, so never set it on a timed or captured path
      const char* trace_path = getenv("SVDD_DEN_TRACE");
      constexpr size_t kTraceWords = (size_t)(denf::kMaxLayers + 1) * 2 * 8;
      if (trace_path && *trace_path) {
        SVDD_CUDA(cudaMalloc(&a.trace, kTraceWords * 8));
        SVDD_CUDA(cudaMemsetAsync(a.trace, 0, kTraceWords * 8, st));
      }
      if (cg == 2) {
        if (tok_dtype == SVDD_TOK_I64) SVDD_TRY(launch2(dens::den_short_kernel<int64_t, 2>));
        else SVDD_TRY(launch2(dens::den_short_kernel<uint8_t, 2>));
      } else {
        if (tok_dtype == SVDD_TOK_I64) SVDD_TRY(launch2(dens::den_short_kernel<int64_t, 1>));
        else SVDD_TRY(launch2(dens::den_short_kernel<uint8_t, 1>));
      }
      count_launch();
      if (a.trace != nullptr) {
        std::vector<unsigned long long> host(kTraceWords);
        SVDD_CUDA(cudaStreamSynchronize(st));
        SVDD_CUDA(cudaMemcpy(host.data(), a.trace, kTraceWords * 8, cudaMemcpyDeviceToHost));
        SVDD_CUDA(cudaFree(a.trace));
        if (FILE* f = fopen(trace_path, "w")) {
          fprintf(f, "# round,item,mma_ready,mma_first,mma_issued,epi_wait,epi_tfull,epi_ld,epi_written,epi_arrived (SM clocks)\n");
          for (int r = 0; r <= a.n_layers; ++r)
            for (int q = 0; q < 2; ++q) {
              fprintf(f, "%d,%d", r, q);
              for (int k = 0; k < 8; ++k) fprintf(f, ",%llu", host[(size_t)(r * 2 + q) * 8 + k]);
              fprintf(f, "\n");
            }
          fclose(f);
        }
      }
      return SVDD_OK;
    }
  }
  // SVDD_DEN_EW=16 (read per call): combined mode with the 16-warp quad epilogue (thread = quarter
  // of a row).  Measured SLOWER than the 8-warp split epilogue -- 27.9 vs 24.2 ms per 51200 x 50 pass,
  // 0.61 vs 0.54 ms at 999 x 33: the round is a chain of fixed latencies (accumulator hand-over,
  // named barrier, async-proxy fence, two mbarrier hops), not a shortage of warps to hide them, and
  // the 640-thread CTA caps the kernel at 96 registers (spills).  Kept with its test.
  const char* env_ew = getenv("SVDD_DEN_EW");
  const bool ew16 = a.cmb && env_ew && atoi(env_ew) == 16;
  auto launch = [&](auto kern, int threads) -> int {
    SVDD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    SVDD_CUDA(launch_k(kern, dim3(grid), dim3((unsigned)threads), (size_t)smem, st, 1, tmW, tmW0, a));
    return SVDD_OK;
  };
  if (tok_dtype == SVDD_TOK_I64) {
    if (ew16) SVDD_TRY(launch(denf::den_fused_kernel<int64_t, 16>, 64 + 32 * 16));
    else SVDD_TRY(launch(denf::den_fused_kernel<int64_t, 8>, denf::kThreads));
  } else {
    if (ew16) SVDD_TRY(launch(denf::den_fused_kernel<uint8_t, 16>, 64 + 32 * 16));
    else SVDD_TRY(launch(denf::den_fused_kernel<uint8_t, 8>, denf::kThreads));
  }
  count_launch();
  return SVDD_OK;
}

extern "C" size_t svdd_denoiser_workspace_bytes(const svdd_denoiser* h, int64_t n_rows, int L) {
  (void)h;
  const size_t nl = (size_t)n_rows * L + 1;
  return DeviceArena::align(nl * kH * sizeof(float)) + 2 * DeviceArena::align(nl * kH * sizeof(__nv_bfloat16));
}

extern "C" int svdd_denoiser_forward(svdd_denoiser* h, const void* tokens, int tok_dtype,
                                     const float* time_bias, float* logits, int64_t n_rows, int L,
                                     void* ws, size_t ws_bytes, void* stream) {
  SVDD_CHECK_ARG(h && tokens && time_bias && logits, "svdd_denoiser_forward: null pointer");
  SVDD_CHECK_ARG(n_rows >= 0 && L >= 1, "svdd_denoiser_forward: bad shape");
  SVDD_CHECK_ARG(n_rows * L < (int64_t)1 << 31, "svdd_denoiser_forward: too many positions for one call");
  SVDD_CHECK_ARG(tok_dtype == SVDD_TOK_I64 || tok_dtype == SVDD_TOK_U8, "bad tok_dtype %d", tok_dtype);
  if (n_rows == 0) return SVDD_OK;
  if (ws_bytes < svdd_denoiser_workspace_bytes(h, n_rows, L) || ws == nullptr) {
    set_last_error("svdd_denoiser_forward: workspace too small (%zu < %zu)", ws_bytes,
                   svdd_denoiser_workspace_bytes(h, n_rows, L));
    return SVDD_ERR_WORKSPACE_TOO_SMALL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t NL = n_rows * L;
  {
    int rc = den_fused_forward(h, tokens, tok_dtype, time_bias, logits, n_rows, L, st);
    if (rc != SVDD_ERR_INTERNAL) return rc;     // SVDD_ERR_INTERNAL = shape not handled: layer-by-layer path
  }
  Workspace W(ws, ws_bytes);
  float* feat = W.take<float>((size_t)(NL + 1) * kH);
  __nv_bfloat16* hbuf[2];
  hbuf[0] = W.take<__nv_bfloat16>((size_t)(NL + 1) * kH);
  hbuf[1] = W.take<__nv_bfloat16>((size_t)(NL + 1) * kH);

  const unsigned grid = (unsigned)ceil_div<int64_t>(NL, kEmbedWarps * kEmbedPosPerWarp);
  if (tok_dtype == SVDD_TOK_I64)
    launch_k(den_embed_kernel<int64_t>, dim3(grid), dim3(kEmbedWarps * 32), 0, st, 1, 
        (const int64_t*)tokens, h->embed_w, h->embed_b, time_bias, h->ln_g, h->ln_b, feat, hbuf[0], NL, L);
  else
    launch_k(den_embed_kernel<uint8_t>, dim3(grid), dim3(kEmbedWarps * 32), 0, st, 1, 
        (const uint8_t*)tokens, h->embed_w, h->embed_b, time_bias, h->ln_g, h->ln_b, feat, hbuf[0], NL, L);
  count_launch();
  SVDD_LAUNCH_CHECK();

  int cur = 0;
  for (int i = 0; i < h->n_layers; ++i) {
    GemmShape g;
    g.S = (int)n_rows; g.L = L; g.L_in = L; g.K = kH; g.N = kH; g.taps = kTaps; g.dil = h->dil[i];
    choose_row_tiling(L, kTaps, &g);
    EpiParams ep;
    ep.bias = h->conv_b + i * kH;
    ep.out = feat;
    ep.out2 = hbuf[cur ^ 1];
    ep.ln_enable = (i + 1 < h->n_layers) ? 1 : 0;
    if (ep.ln_enable) {
      ep.ln_gamma = h->ln_g + (i + 1) * kH;
      ep.ln_beta = h->ln_b + (i + 1) * kH;
      ep.ln_tbias = time_bias + (i + 1) * kH;
    }
    SVDD_TRY(launch_conv_gemm(hbuf[cur], h->conv_w + (size_t)i * kTaps * kH * kH, g, EPI_DEN_LN, ep, st));
    cur ^= 1;
  }
  {
    GemmShape g;
    g.S = 1; g.L = (int)NL; g.L_in = (int)NL; g.K = kH; g.N = kH; g.taps = 1; g.dil = 1;
    g.BL = 128; g.BS = 1;
    EpiParams ep;
    ep.bias = h->fc0_b;
    ep.w2 = h->fc2_w;
    ep.b2 = h->fc2_b;
    ep.out = logits;
    SVDD_TRY(launch_conv_gemm(hbuf[cur], h->fc0_w, g, EPI_DEN_FINAL, ep, st));
  }
  return SVDD_OK;
}
