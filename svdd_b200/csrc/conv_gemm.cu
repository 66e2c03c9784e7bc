// Host launcher for the tcgen05 implicit-GEMM (conv_gemm.cuh): builds the TMA
// tensor maps, picks the N tile and launches one persistent CTA per SM.
#include <mutex>
#include <stdlib.h>

#include <vector>

#include "conv_gemm.cuh"
#include "conv_gemm2.cuh"

namespace svdd {
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_map(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, bool swizzle64 = false);
int encode_bf16_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box) {
  return encode_map(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box);
}
int encode_map(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const cuuint64_t* dims,
               const cuuint64_t* strides_bytes, const cuuint32_t* box, bool swizzle64) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
    return SVDD_ERR_CUDA;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base),
                  dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  // 64-byte-wide boxes (slab32) take the 64B swizzle: with SWIZZLE_128B a 64-byte box row
                  // still occupies a 128-byte shared-memory row (compute-sanitizer: footprint 2x)
                  swizzle64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)",
                   (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                   (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], box[1], rank > 2 ? box[2] : 0);
    return SVDD_ERR_CUDA;
  }
  return SVDD_OK;
}

// ---- optional per-launch timing (bench.py's roofline leg; off on the hot path) --------
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // start/stop pairs
  struct Rec { GemmShape g; int bn, mode; double flops; };
  std::vector<Rec> recs;
  double flops = 0.0;
  // longest launch of the last profile window (bench.py's dominant-launch roofline)
  double top_ms = 0.0, top_flops = 0.0;
  Rec top = {};
  std::mutex mu;
};
ProfState& prof() {
  static ProfState p;
  return p;
}

template <int BN, int MODE>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmW, const GemmShape& g,
                const EpiParams& ep, cudaStream_t stream) {
  using C = gemm_detail::Cfg<BN, MODE>;
  auto kern = gemm_detail::conv_gemm_kernel<BN, MODE>;
  static bool configured = false;
  if (!configured) {
    SVDD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes));
    configured = true;
  }
  const int64_t tiles = (int64_t)ceil_div(g.L, g.BL) * ceil_div(g.S, g.BS) * (g.N / BN);
  if (tiles == 0) return SVDD_OK;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  ProfState& P = prof();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (P.on) {
    SVDD_CUDA(cudaEventCreate(&e0));
    SVDD_CUDA(cudaEventCreate(&e1));
    SVDD_CUDA(cudaEventRecord(e0, stream));
  }
  SVDD_CUDA(launch_k(kern, dim3((unsigned)grid), dim3(gemm_detail::kThreads), C::kSmemBytes, stream, 1,
                     tmA, tmW, g, ep));
  count_launch();
  if (P.on) {
    SVDD_CUDA(cudaEventRecord(e1, stream));
    std::lock_guard<std::mutex> lk(P.mu);
    P.ev.push_back(e0);
    P.ev.push_back(e1);
    // nominal dense count (padding taps included), the figure SURVEY.md section 8(d) uses
    const double fl = 2.0 * (double)g.S * g.L * (C::kAcc) * (double)g.N * g.K * g.taps;
    P.flops += fl;
    P.recs.push_back({g, BN, MODE, fl});
  }
  return SVDD_OK;
}

int pick_bn(const GemmShape& g, int mode) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("SVDD_GEMM_BN");
    forced = e ? atoi(e) : 0;
  }
  const int64_t m_tiles = (int64_t)ceil_div(g.L, g.BL) * ceil_div(g.S, g.BS);
  if (mode == EPI_DEN_LN || mode == EPI_DEN_FINAL) return 128;
  if (mode == EPI_POOL) return (g.N % 128 == 0) ? 128 : 64;
  if (forced > 0 && g.N % forced == 0) return forced;
  const int64_t two_waves = 2 * (int64_t)num_sms();
  if (g.N % 256 == 0 && m_tiles * (g.N / 256) >= two_waves) return 256;
  if (g.N % 224 == 0 && g.N % 256 != 0 && m_tiles * (g.N / 224) >= two_waves) return 224;
  if (g.N % 192 == 0 && g.N % 256 != 0 && m_tiles * (g.N / 192) >= two_waves) return 192;
  if (g.N % 128 == 0) return 128;
  if (g.N % 192 == 0) return 192;
  return 64;
}


// ---- second-generation kernel (conv_gemm2.cuh) ----------------------------------------
int gemm2_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVDD_GEMM2"); v = e ? atoi(e) : 1; }
  return v;
}
int gemm2_cg() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SVDD_GEMM_CG"); v = e ? atoi(e) : 2; if (v != 1) v = 2; }
  return v;
}

template <int BN, int MODE, int CG, bool HALO = false, int EW = 8, bool WIDE = false>
int launch2_impl(const CUtensorMap& tmA, const CUtensorMap& tmW, const CUtensorMap& tmO, const CUtensorMap& tmO2,
                 const CUtensorMap& tmR, const CUtensorMap& tmR2, const GemmShape& g, const EpiParams& ep,
                 cudaStream_t stream) {
  using C = gemm2::Cfg2<BN, CG>;
  auto kern = gemm2::gemm2_kernel<BN, MODE, CG, HALO, EW, WIDE>;
  constexpr int kSmem = HALO ? C::kSmemBytesH : C::kSmemBytes;
  static bool configured = false;
  if (!configured) {
    SVDD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    configured = true;
  }
  const int64_t m_tiles = (int64_t)ceil_div(g.L, g.BL) * ceil_div(g.S, g.BS);
  const int64_t tiles = ceil_div<int64_t>(m_tiles, CG) * ceil_div(g.N, BN);
  if (tiles == 0) return SVDD_OK;
  SVDD_CHECK_ARG(m_tiles * ceil_div(g.N, BN) < ((int64_t)1 << 30), "conv_gemm: too many tiles for one launch");
  const int64_t max_clusters = num_sms() / CG;
  const int grid = (int)(tiles < max_clusters ? tiles : max_clusters) * CG;
  static int prefetch_w = -1;
  if (prefetch_w < 0) { const char* e = getenv("SVDD_PREFETCH_W"); prefetch_w = e ? atoi(e) : 0; }   // measured neutral-to-negative on the c2 step (195.8 vs 193.9 seq/s): off
  GemmShape gk = g;
  gk.prefetch_w = (prefetch_w && tiles <= 2 * max_clusters) ? 1 : 0;
  // SVDD_TIMELINE=1 (debugging): synchronise after every launch and print CTA 0's milestones
  static int timeline = -1;
  static unsigned long long* tl_dev = nullptr;
  if (timeline < 0) { const char* e = getenv("SVDD_TIMELINE"); timeline = e ? atoi(e) : 0; }
  EpiParams epk = ep;
  if (timeline) {
    if (tl_dev == nullptr) SVDD_CUDA(cudaMalloc(&tl_dev, 16 * sizeof(unsigned long long)));
    SVDD_CUDA(cudaMemsetAsync(tl_dev, 0, 16 * sizeof(unsigned long long), stream));
    epk.timeline = tl_dev;
  }
  ProfState& P = prof();
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (P.on) {
    SVDD_CUDA(cudaEventCreate(&e0));
    SVDD_CUDA(cudaEventCreate(&e1));
    SVDD_CUDA(cudaEventRecord(e0, stream));
  }
  SVDD_CUDA(launch_k(kern, dim3((unsigned)grid), dim3(64 + 32 * EW + 32 * gemm2::kStoreWarps), kSmem, stream, CG,
                     tmA, tmW, tmO, tmO2, tmR, tmR2, gk, epk));
  if (timeline) {
    unsigned long long h[16];
    SVDD_CUDA(cudaStreamSynchronize(stream));
    SVDD_CUDA(cudaMemcpy(h, tl_dev, sizeof(h), cudaMemcpyDeviceToHost));
    auto d = [&](int i) { return h[i] ? (long long)(h[i] - h[0]) : -1ll; };
    fprintf(stderr, "timeline mode %d rows %lld K %d N %d taps %d tiles %lld: setup %lld first_tma %lld last_tma_tile0 %lld "
            "first_full %lld mma_done %lld epi_start %lld epi_end %lld stores_done %lld pre_sync %lld exit %lld (cycles)\n",
            MODE, (long long)g.S * g.L, g.K, g.N, g.taps, (long long)tiles, d(1), d(2), d(10), d(3), d(4), d(5), d(6), d(7),
            d(8), d(9));
  }
  count_launch();
  if (P.on) {
    SVDD_CUDA(cudaEventRecord(e1, stream));
    std::lock_guard<std::mutex> lk(P.mu);
    P.ev.push_back(e0);
    P.ev.push_back(e1);
    const double rows_fl = g.useful_rows > 0 ? (double)g.useful_rows : (double)g.S * g.L;
    const double fl = 2.0 * rows_fl * (double)g.N * ((double)g.K * g.taps + g.K2);
    P.flops += fl;
    P.recs.push_back({g, BN + 1000 * CG + (EW == 16 ? 10000 : 0), MODE, fl});
  }
  return SVDD_OK;
}

bool gemm2_handles(const GemmShape& g, int mode, const EpiParams& ep) {
  if (mode == EPI_PAIR || mode == EPI_POOL2) return true;     // validated by the launcher
  if (!gemm2_enabled()) return false;
  if (mode != EPI_GENERIC && mode != EPI_HEADDOT) return false;
  if (g.N % 128 != 0) return false;
  auto ok_dt = [](int dt) { return dt == DT_BF16 || dt == DT_F32; };
  auto es = [](int dt) { return dt == DT_F32 ? 4 : 2; };
  auto aligned = [&](const void* p, int64_t ld, int dt) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld * es(dt)) % 16 == 0;
  };
  if (mode == EPI_HEADDOT) return ep.res == nullptr && ep.out == nullptr && ep.out2 == nullptr;
  if (ep.out == nullptr && ep.out2 == nullptr) return false;
  if (ep.out != nullptr && !(ok_dt(ep.out_dtype) && aligned(ep.out, ep.ld_out, ep.out_dtype))) return false;
  if (ep.res != nullptr) {
    if (ep.out == nullptr || ep.res_dtype != ep.out_dtype || !aligned(ep.res, ep.ld_res, ep.res_dtype)) return false;
  } else if (ep.out2 != nullptr) {
    if (!ok_dt(ep.out2_dtype) || !aligned(ep.out2, ep.ld_out2, ep.out2_dtype)) return false;
    if (ep.out != nullptr && ep.out2_dtype != ep.out_dtype) return false;
  }
  return true;
}

int pick_bn2(const GemmShape& g, int cg, int mode = EPI_GENERIC) {
  const int64_t m_tiles = (int64_t)ceil_div(g.L, g.BL) * ceil_div(g.S, g.BS);
  const int64_t mp = ceil_div<int64_t>(m_tiles, cg);
  static int forced = -1, ragged = -1;
  if (forced < 0) { const char* e = getenv("SVDD_GEMM2_BN"); forced = e ? atoi(e) : 0; }
  if (ragged < 0) { const char* e = getenv("SVDD_RAGGED_N"); ragged = e ? atoi(e) : 1; }
  if (forced > 0 && g.N % forced == 0) return forced;
  if (g.N % 256 == 0 && mp * (g.N / 256) >= (num_sms() / cg) / 2) return 256;
  // N = 896 / 1152 on the 1x1 GEMMs: 256-wide tiles with a ragged last tile (TMA zero-fills the
  // missing weight rows and clips the stores; 10-12 % of the MMA work is wasted) beat 128-wide
  // tiles (0.6 vs 0.9 PFLOP/s on these epilogue-heavy shapes)
  if (ragged && (mode == EPI_PAIR || mode == EPI_POOL2) && g.N % 256 == 128 && g.N > 256 && g.n_off == 0 &&
      (g.N_w == 0 || g.N_w == g.N) && mp * ceil_div(g.N, 256) >= (num_sms() / cg) / 2)
    return 256;
  return 128;
}

}  // namespace

int encode_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                        uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t b0, uint32_t b1, uint32_t b2) {
  const cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  const cuuint64_t str[2] = {(cuuint64_t)stride1_bytes, (cuuint64_t)stride2_bytes};
  const cuuint32_t box[3] = {b0, b1, b2};
  return encode_bf16_map(map, base, 3, dims, str, box);
}
int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows,
                        uint32_t box_inner, uint32_t box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t str[1] = {(cuuint64_t)inner * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  return encode_bf16_map(map, base, 2, dims, str, box);
}

int encode_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t inner, uint64_t rows,
                       uint32_t box_inner, uint32_t box_rows) {
  const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  const cuuint64_t str[1] = {(cuuint64_t)inner * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  return encode_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, 2, dims, str, box);
}

bool gemm_prof_on() { return prof().on; }
void gemm_prof_record(cudaEvent_t e0, cudaEvent_t e1, const GemmShape& g, int bn, int mode, double flops) {
  ProfState& P = prof();
  std::lock_guard<std::mutex> lk(P.mu);
  P.ev.push_back(e0);
  P.ev.push_back(e1);
  P.flops += flops;
  P.recs.push_back({g, bn, mode, flops});
}

int conv_gemm_n_tiles(const GemmShape& g, int mode) {
  if (mode == EPI_HEADDOT && gemm2_enabled() && g.N % 128 == 0) return 2 * (g.N / pick_bn2(g, gemm2_cg()));
  return 2 * (g.N / pick_bn(g, mode));
}

void choose_row_tiling(int L, int taps, GemmShape* g) {
  (void)taps;
  // A tile is BL positions x BS sequences (BL*BS <= 128 MMA rows).  Pick the pair that leaves
  // the fewest of the 128 rows empty: L = 200/100/50/25 -> 25 x 5 (125 rows, 97.7 %) where
  // whole-sequence or 128-position tiles only fill 78 %.  Smaller BL is taken only for a
  // gain above 3 % (fewer, longer TMA box rows otherwise).
  static int legacy = -1;
  if (legacy < 0) { const char* e = getenv("SVDD_TILING_LEGACY"); legacy = e ? atoi(e) : 0; }
  int best_bl = L <= 128 ? L : 128;
  double best_eff = 0.0;
  for (int bl = best_bl; bl >= (L >= 2 ? 2 : 1) && !legacy; --bl) {
    const int bs = 128 / bl;
    if (bs > 256) continue;
    const double eff = (double)L * bs / ((double)ceil_div(L, bl) * 128.0);
    if (eff > best_eff * 1.03) { best_eff = eff; best_bl = bl; }
  }
  g->BL = best_bl;
  g->BS = 128 / best_bl;
}

static int launch_gemm2_window(const void* A, const void* W, const GemmShape& g, int mode, const EpiParams& ep_in,
                               cudaStream_t stream);
static int launch_gemm1(const void* A, const void* W, const GemmShape& g, int mode, const EpiParams& ep_in,
                        cudaStream_t stream);

int launch_conv_gemm(const void* A, const void* W, const GemmShape& g, int mode,
                     const EpiParams& ep_in, cudaStream_t stream) {
  SVDD_CHECK_ARG(g.K > 0 && g.K % 64 == 0, "conv_gemm: K=%d must be a positive multiple of 64", g.K);
  SVDD_CHECK_ARG(g.N > 0 && g.N % 64 == 0, "conv_gemm: N=%d must be a positive multiple of 64", g.N);
  SVDD_CHECK_ARG(g.BL >= 1 && g.BS >= 1 && g.BL * g.BS <= 128, "conv_gemm: bad tile %dx%d", g.BL, g.BS);
  SVDD_CHECK_ARG(g.BL <= 256 && g.BS <= 256, "conv_gemm: TMA box too large");
  SVDD_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                 "conv_gemm: operands must be 16-byte aligned");
  if (g.S == 0 || g.L == 0) return SVDD_OK;

  if (gemm2_handles(g, mode, ep_in)) {
    // k5 convs with N = 896 / 1152 (not a multiple of 256): two launches over column windows,
    // 256-wide tiles for the first floor(N/256)*256 columns and 128-wide tiles for the rest,
    // instead of 128-wide tiles everywhere (0.85 vs 1.45 PFLOP/s on the two widths; measured
    // 516 -> 369 us and 244 -> 175 us on the c2 step).  The short-K 1x1 GEMMs lose on the extra
    // launch and stay on one width.
    static int nsplit = -1;
    if (nsplit < 0) { const char* e = getenv("SVDD_NSPLIT"); nsplit = e ? atoi(e) : 1; }
    GemmShape gw = g;
    gw.N_w = g.N;
    gw.n_off = 0;
    if (nsplit && mode == EPI_GENERIC && g.taps > 1 && g.N > 256 && g.N % 256 == 128) {
      GemmShape ga = gw;
      ga.N = (g.N / 256) * 256;
      if (pick_bn2(ga, gemm2_cg()) == 256) {
        SVDD_TRY(launch_gemm2_window(A, W, ga, mode, ep_in, stream));
        gw.n_off = ga.N;
        gw.N = g.N - ga.N;
      }
    }
    return launch_gemm2_window(A, W, gw, mode, ep_in, stream);
  }
  return launch_gemm1(A, W, g, mode, ep_in, stream);
}

static int launch_gemm2_window(const void* A, const void* W, const GemmShape& g, int mode, const EpiParams& ep_in,
                               cudaStream_t stream) {
  SVDD_CHECK_ARG(g.K2 == 0 || (mode == EPI_PAIR && g.taps == 1), "conv_gemm: K2 is an EPI_PAIR (1x1) feature");
  {
    const int cg = gemm2_cg();
    const int bn2 = pick_bn2(g, cg, mode);
    EpiParams ep2 = ep_in;
    if (ep2.out == nullptr && ep2.out2 != nullptr) ep2.out_dtype = ep2.out2_dtype;   // slab geometry follows the staged output
    {
      // x += GEMM (the transformer's out-projection and second FFN linear): no residual read, the
      // fp32 slab is accumulated into `out` by a TMA reduce-add
      static int res_reduce = -1;
      if (res_reduce < 0) { const char* e = getenv("SVDD_RES_REDUCE"); res_reduce = e ? atoi(e) : 1; }
      ep2.res_reduce = (res_reduce && mode == EPI_GENERIC && ep2.res != nullptr && ep2.res == ep2.out &&
                        ep2.out_dtype == DT_F32 && ep2.res_dtype == DT_F32 && ep2.out2 == nullptr &&
                        !ep2.act_after_res && ep2.ld_res == ep2.ld_out) ? 1 : 0;
    }
    CUtensorMap tA, tW, tO, tO2, tR, tR2;
    if (g.halo) {
      SVDD_CHECK_ARG(mode == EPI_GENERIC && cg == 2 && g.S == 1 && g.BS == 1 && g.BL == 128 && g.taps % 2 == 1 &&
                     g.taps - 1 <= 8 && g.dil == 1 && g.K2 == 0,
                     "conv_gemm: halo mode needs flat rows (S = 1, 128-row tiles), odd taps <= 9, cta_group 2");
    }
    {
      const int64_t a_pitch = g.a_pitch > 0 ? g.a_pitch : g.L_in;
      const cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.L_in, (cuuint64_t)g.S};
      const cuuint64_t str[2] = {(cuuint64_t)g.K * 2, (cuuint64_t)a_pitch * g.K * 2};
      const cuuint32_t box[3] = {64, (cuuint32_t)(g.halo ? g.BL + g.taps - 1 : g.BL), (cuuint32_t)g.BS};
      SVDD_TRY(encode_bf16_map(&tA, A, 3, dims, str, box));
      const cuuint64_t wd[2] = {(cuuint64_t)g.K, (cuuint64_t)g.taps * g.N_w};
      const cuuint64_t ws[1] = {(cuuint64_t)g.K * 2};
      const cuuint32_t wb[2] = {64, (cuuint32_t)(bn2 / cg)};
      SVDD_TRY(encode_bf16_map(&tW, W, 2, wd, ws, wb));
    }
    // EPI_PAIR / EPI_POOL2 on the 8-warp kernel: 32-column slabs, four per column half (conv_gemm2.cuh);
    // SVDD_SLAB32=0 (read per call) selects the two 64-column slabs per half of rounds 1-2 (bit-identical)
    bool use_pair16 = false, use_pool16 = false, use_gen16 = false, gen16_quarter = false;
    {
      // 16 epilogue warps on the same slabs (thread = row x 16 columns): pooling -6..-8 %, residual 1x1
      // -2..-4 % per launch on the c2 pass -- after slab32 both families run at the SM's shared-memory
      // bandwidth (operand reads of the MMAs + the slab traffic: 6.3 / 5.7 us per 256 x 256 tile at
      // K = 768 against 7.6 / 6.2 measured), so more warps buy little.  SVDD_PAIR16=0 / SVDD_POOL16=0
      // (read per call) select the 8-warp kernels; bit-identical.
      const char* ep16 = getenv("SVDD_PAIR16");
      const char* e16 = getenv("SVDD_POOL16");
      const bool shape16 = bn2 == 256 && cg == 2 && !g.halo;
      use_pair16 = mode == EPI_PAIR && shape16 && (ep16 == nullptr || atoi(ep16) != 0);
      use_pool16 = mode == EPI_POOL2 && shape16 && ep2.out == nullptr && ep2.out2 != nullptr && (e16 == nullptr || atoi(e16) != 0);
      const char* es32 = getenv("SVDD_SLAB32");
      const bool wide = es32 != nullptr && atoi(es32) == 0 && bn2 == 256 && cg == 2;   // the only WIDE instantiations
      ep2.slab32 = ((mode == EPI_PAIR || mode == EPI_POOL2) && !wide) ? 1 : 0;
      if (wide) use_pair16 = use_pool16 = false;
      // 16 epilogue warps for single-output bf16 EPI_GENERIC launches without residual (the stem, K = 64, pure
      // epilogue: 304 us on 8 warps, 145 on 16 with one 64-column slab per column quarter).  The slab32
      // protocol (four 32-column slabs per half, SVDD_EPI16=3) is measured slightly SLOWER here, 152 vs 145 us:
      // a store-only epilogue never waited for its slabs, and the narrower TMA boxes cost a little.
      // SVDD_EPI16 (read per call): 1 = quarter slabs (default), 3 = slab32, 0 = the 8-warp kernel.
      const char* e_epi = getenv("SVDD_EPI16");
      const int epi16 = e_epi ? atoi(e_epi) : 1;
      const bool one_out = (ep2.out != nullptr) != (ep2.out2 != nullptr);
      const int staged_dt = ep2.out != nullptr ? ep2.out_dtype : ep2.out2_dtype;
      use_gen16 = epi16 != 0 && shape16 && mode == EPI_GENERIC && ep2.res == nullptr && !ep2.res_reduce && one_out &&
                  staged_dt == DT_BF16 && !ep2.act_after_res;
      gen16_quarter = use_gen16 && epi16 != 3;
      if (use_gen16 && !gen16_quarter) ep2.slab32 = 1;
    }
    // [S, Lr, N] row-major view with leading dimension ld; box = 128 bytes (slab32: 64) x box_l x BS
    auto io_map = [&](CUtensorMap* m, const void* p, int dt, int64_t ld, int Lr, int box_l, int pitch = 0) -> int {
      if (p == nullptr) { *m = tA; return SVDD_OK; }
      const cuuint64_t es = dt == DT_F32 ? 4 : 2;
      const cuuint64_t dims[3] = {(cuuint64_t)g.N_w, (cuuint64_t)Lr, (cuuint64_t)g.S};
      const cuuint64_t str[2] = {(cuuint64_t)ld * es, (cuuint64_t)(pitch > 0 ? pitch : Lr) * ld * es};
      const cuuint32_t box[3] = {(cuuint32_t)((ep2.slab32 ? 64 : 128) / es), (cuuint32_t)box_l, (cuuint32_t)g.BS};
      return encode_map(m, dt == DT_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, p, 3,
                        dims, str, box, ep2.slab32 != 0);
    };
    auto aligned16 = [](const void* p, int64_t ld_elems, int es) {
      return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld_elems * es) % 16 == 0;
    };
    if (mode == EPI_PAIR) {
      SVDD_CHECK_ARG(g.N % 128 == 0 && g.BL % 2 == 0 && g.taps == 1, "conv_gemm: EPI_PAIR needs N %% 128 == 0, an even BL, 1x1");
      SVDD_CHECK_ARG(ep2.out && ep2.out2 && ep2.out_dtype == DT_BF16 && ep2.out2_dtype == DT_BF16,
                     "conv_gemm: EPI_PAIR needs bf16 out / out2");
      SVDD_CHECK_ARG(aligned16(ep2.out, ep2.ld_out, 2) && aligned16(ep2.out2, ep2.ld_out2, 2),
                     "conv_gemm: EPI_PAIR operands must be 16-byte aligned");
      const int Lo = (g.L + 1) / 2;
      SVDD_TRY(io_map(&tO, ep2.out, DT_BF16, ep2.ld_out, Lo, g.BL / 2));
      SVDD_TRY(io_map(&tO2, ep2.out2, DT_BF16, ep2.ld_out2, Lo, g.BL / 2));
      if (g.K2 > 0) {
        // the residual is recomputed from a second K-concatenated operand pair instead of read
        SVDD_CHECK_ARG(g.K2 % 64 == 0 && ep2.a2 && ep2.w2k && ep2.res == nullptr,
                       "conv_gemm: K2 needs a2 / w2k, a multiple of 64 and no residual");
        SVDD_CHECK_ARG((reinterpret_cast<uintptr_t>(ep2.a2) & 15) == 0 && (reinterpret_cast<uintptr_t>(ep2.w2k) & 15) == 0,
                       "conv_gemm: a2 / w2k must be 16-byte aligned");
        const cuuint64_t dims[3] = {(cuuint64_t)g.K2, (cuuint64_t)g.L_in, (cuuint64_t)g.S};
        const cuuint64_t str[2] = {(cuuint64_t)g.K2 * 2, (cuuint64_t)g.L_in * g.K2 * 2};
        const cuuint32_t box[3] = {64, (cuuint32_t)g.BL, (cuuint32_t)g.BS};
        SVDD_TRY(encode_bf16_map(&tR, ep2.a2, 3, dims, str, box));
        const cuuint64_t wd[2] = {(cuuint64_t)g.K2, (cuuint64_t)g.N_w};
        const cuuint64_t ws[1] = {(cuuint64_t)g.K2 * 2};
        const cuuint32_t wb[2] = {64, (cuuint32_t)(bn2 / cg)};
        SVDD_TRY(encode_bf16_map(&tR2, ep2.w2k, 2, wd, ws, wb));
      } else {
        SVDD_CHECK_ARG(ep2.res && ep2.res_dtype == DT_BF16 && aligned16(ep2.res, ep2.ld_res, 2),
                       "conv_gemm: EPI_PAIR needs a 16-byte aligned bf16 residual");
        SVDD_TRY(io_map(&tR, ep2.res, DT_BF16, ep2.ld_res, g.L, g.BL));
        tR2 = tA;
      }
    } else if (mode == EPI_POOL2) {
      SVDD_CHECK_ARG(g.N % 128 == 0 && g.taps == 1, "conv_gemm: EPI_POOL2 needs N %% 128 == 0, 1x1");
      SVDD_CHECK_ARG(ep2.res && ep2.res2 && (ep2.out || ep2.out2), "conv_gemm: EPI_POOL2 needs res, res2 and an output");
      SVDD_CHECK_ARG(ep2.out == nullptr || ep2.out_dtype == DT_F32, "conv_gemm: EPI_POOL2 `out` is fp32");
      SVDD_CHECK_ARG(ep2.out2 == nullptr || ep2.out2_dtype == DT_BF16, "conv_gemm: EPI_POOL2 `out2` is bf16");
      SVDD_CHECK_ARG(aligned16(ep2.res, ep2.ld_res, 2) && aligned16(ep2.res2, ep2.ld_res2, 2) &&
                     (ep2.out2 == nullptr || aligned16(ep2.out2, ep2.ld_out2, 2)),
                     "conv_gemm: EPI_POOL2 operands must be 16-byte aligned");
      SVDD_CHECK_ARG(ep2.out == nullptr || ep2.out_pitch == 0, "conv_gemm: EPI_POOL2 `out` is densely packed");
      tO = tA;
      SVDD_TRY(io_map(&tO2, ep2.out2, DT_BF16, ep2.ld_out2, g.L, g.BL, ep2.out2_pitch));
      SVDD_TRY(io_map(&tR, ep2.res, DT_BF16, ep2.ld_res, g.L, g.BL, ep2.res_pitch));
      SVDD_TRY(io_map(&tR2, ep2.res2, DT_BF16, ep2.ld_res2, g.L, g.BL, ep2.res2_pitch));
    } else {
      SVDD_TRY(io_map(&tO, ep2.out, ep2.out_dtype, ep2.ld_out, g.L, g.BL));
      SVDD_TRY(io_map(&tO2, ep2.res == nullptr ? ep2.out2 : nullptr, ep2.out2_dtype, ep2.ld_out2, g.L, g.BL));
      SVDD_TRY(io_map(&tR, ep2.res, ep2.res_dtype, ep2.ld_res, g.L, g.BL));
      tR2 = tA;
    }
    {
      if (use_gen16) {
        if (gen16_quarter) return launch2_impl<256, EPI_GENERIC, 2, false, 16, true>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream);
        return launch2_impl<256, EPI_GENERIC, 2, false, 16>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream);
      }
      if (use_pair16) return launch2_impl<256, EPI_PAIR, 2, false, 16>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream);
      if (use_pool16) return launch2_impl<256, EPI_POOL2, 2, false, 16>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream);
    }
    if (g.halo) {
      if (bn2 == 256) return launch2_impl<256, EPI_GENERIC, 2, true>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream);
      if (bn2 == 128) return launch2_impl<128, EPI_GENERIC, 2, true>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream);
    }
    if ((mode == EPI_PAIR || mode == EPI_POOL2) && !ep2.slab32) {     // SVDD_SLAB32=0 (cross-check tests)
      if (mode == EPI_PAIR) return launch2_impl<256, EPI_PAIR, 2, false, 8, true>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream);
      return launch2_impl<256, EPI_POOL2, 2, false, 8, true>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream);
    }
#define CASE2(BN_, MODE_, CG_) \
    if (bn2 == BN_ && mode == MODE_ && cg == CG_) return launch2_impl<BN_, MODE_, CG_>(tA, tW, tO, tO2, tR, tR2, g, ep2, stream)
    CASE2(128, EPI_GENERIC, 1); CASE2(256, EPI_GENERIC, 1);
    CASE2(128, EPI_GENERIC, 2); CASE2(256, EPI_GENERIC, 2);
    CASE2(128, EPI_HEADDOT, 1); CASE2(256, EPI_HEADDOT, 1);
    CASE2(128, EPI_HEADDOT, 2); CASE2(256, EPI_HEADDOT, 2);
    CASE2(128, EPI_PAIR, 1); CASE2(256, EPI_PAIR, 1);
    CASE2(128, EPI_PAIR, 2); CASE2(256, EPI_PAIR, 2);
    CASE2(128, EPI_POOL2, 1); CASE2(256, EPI_POOL2, 1);
    CASE2(128, EPI_POOL2, 2); CASE2(256, EPI_POOL2, 2);
#undef CASE2
    set_last_error("conv_gemm: no gemm2 kernel for BN=%d mode=%d", bn2, mode);
    return SVDD_ERR_INTERNAL;
  }
}

static int launch_gemm1(const void* A, const void* W, const GemmShape& g, int mode, const EpiParams& ep_in,
                        cudaStream_t stream) {
  const int bn = pick_bn(g, mode);
  SVDD_CHECK_ARG(g.N % bn == 0, "conv_gemm: N=%d not divisible by tile %d", g.N, bn);
  if (mode == EPI_DEN_LN || mode == EPI_DEN_FINAL)
    SVDD_CHECK_ARG(g.N == 128, "conv_gemm: denoiser epilogues need N == 128 (got %d)", g.N);

  static int prefetch_flag = -1;
  if (prefetch_flag < 0) {
    const char* e = getenv("SVDD_EPI_PREFETCH");
    prefetch_flag = e ? atoi(e) : 1;
  }
  EpiParams ep = ep_in;
  ep.prefetch = prefetch_flag;

  CUtensorMap tmA, tmW;
  if (mode == EPI_POOL) {
    SVDD_CHECK_ARG(g.taps == 1, "conv_gemm: pooling epilogue is 1x1 only");
    const cuuint64_t dims[4] = {(cuuint64_t)g.K, 2, (cuuint64_t)g.L, (cuuint64_t)g.S};
    const cuuint64_t str[3] = {(cuuint64_t)g.K * 2, (cuuint64_t)g.K * 4, (cuuint64_t)g.L_in * g.K * 2};
    const cuuint32_t box[4] = {64, 1, (cuuint32_t)g.BL, (cuuint32_t)g.BS};
    SVDD_TRY(encode_bf16_map(&tmA, A, 4, dims, str, box));
  } else {
    const cuuint64_t dims[3] = {(cuuint64_t)g.K, (cuuint64_t)g.L_in, (cuuint64_t)g.S};
    const cuuint64_t str[2] = {(cuuint64_t)g.K * 2, (cuuint64_t)g.L_in * g.K * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)g.BL, (cuuint32_t)g.BS};
    SVDD_TRY(encode_bf16_map(&tmA, A, 3, dims, str, box));
  }
  {
    const cuuint64_t dims[2] = {(cuuint64_t)g.K, (cuuint64_t)g.taps * g.N};
    const cuuint64_t str[1] = {(cuuint64_t)g.K * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)bn};
    SVDD_TRY(encode_bf16_map(&tmW, W, 2, dims, str, box));
  }

#define CASE(BN_, MODE_) \
  if (bn == BN_ && mode == MODE_) return launch_impl<BN_, MODE_>(tmA, tmW, g, ep, stream)
  CASE(64, EPI_GENERIC);
  CASE(128, EPI_GENERIC);
  CASE(192, EPI_GENERIC);
  CASE(224, EPI_GENERIC);
  CASE(256, EPI_GENERIC);
  CASE(128, EPI_DEN_LN);
  CASE(128, EPI_DEN_FINAL);
  CASE(64, EPI_POOL);
  CASE(128, EPI_POOL);
  CASE(64, EPI_HEADDOT);
  CASE(128, EPI_HEADDOT);
  CASE(256, EPI_HEADDOT);
#undef CASE
  set_last_error("conv_gemm: no kernel for BN=%d mode=%d", bn, mode);
  return SVDD_ERR_INTERNAL;
}

}  // namespace svdd

extern "C" int svdd_profile_begin(void) {
  svdd::ProfState& P = svdd::prof();
  std::lock_guard<std::mutex> lk(P.mu);
  for (cudaEvent_t e : P.ev) cudaEventDestroy(e);
  P.ev.clear();
  P.recs.clear();
  P.flops = 0.0;
  P.on = true;
  return SVDD_OK;
}

extern "C" int svdd_profile_end(double* gemm_ms, int64_t* gemm_launches, double* gemm_flops) {
  svdd::ProfState& P = svdd::prof();
  std::lock_guard<std::mutex> lk(P.mu);
  P.on = false;
  double ms = 0.0;
  P.top_ms = 0.0;
  for (size_t i = 0; i + 1 < P.ev.size(); i += 2) {
    if (cudaEventSynchronize(P.ev[i + 1]) != cudaSuccess) break;
    float t = 0.0f;
    if (cudaEventElapsedTime(&t, P.ev[i], P.ev[i + 1]) == cudaSuccess) ms += t;
    if (i / 2 < P.recs.size() && t > P.top_ms) { P.top_ms = t; P.top = P.recs[i / 2]; P.top_flops = P.recs[i / 2].flops; }
    if (getenv("SVDD_PROF_DUMP") && i / 2 < P.recs.size()) {   // per-launch table for tuning
      const svdd::ProfState::Rec& r = P.recs[i / 2];
      fprintf(stderr, "gemm %3zu mode %d BN %3d S %6d L %6d K %5d N %5d taps %d dil %2d  %8.1f us  %7.1f TFLOP/s\n",
              i / 2, r.mode, r.bn, r.g.S, r.g.L, r.g.K, r.g.N, r.g.taps, r.g.dil, t * 1e3,
              r.flops / (t * 1e-3) / 1e12);
    }
  }
  if (gemm_ms) *gemm_ms = ms;
  if (gemm_launches) *gemm_launches = (int64_t)(P.ev.size() / 2);
  if (gemm_flops) *gemm_flops = P.flops;
  for (cudaEvent_t e : P.ev) cudaEventDestroy(e);
  P.ev.clear();
  return SVDD_OK;
}

extern "C" int svdd_profile_top(double* ms, double* flops, int64_t* rows, int* K, int* N, int* taps, int* mode) {
  svdd::ProfState& P = svdd::prof();
  std::lock_guard<std::mutex> lk(P.mu);
  if (ms) *ms = P.top_ms;
  if (flops) *flops = P.top_flops;
  if (rows) *rows = (int64_t)P.top.g.S * P.top.g.L;
  if (K) *K = P.top.g.K;
  if (N) *N = P.top.g.N;
  if (taps) *taps = P.top.g.taps;
  if (mode) *mode = P.top.mode;
  return SVDD_OK;
}

namespace svdd {
// ---- plain CUDA-core reference of the same contraction (self test only) ----------
namespace {
__global__ void naive_conv_gemm_kernel(const __nv_bfloat16* __restrict__ A,
                                       const __nv_bfloat16* __restrict__ W,
                                       const float* __restrict__ bias, float* __restrict__ C, int S,
                                       int L, int K, int N, int taps, int dil) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)S * L * N) return;
  const int n = (int)(idx % N);
  const int64_t row = idx / N;
  const int l = (int)(row % L), s = (int)(row / L);
  float acc = 0.0f;
  for (int t = 0; t < taps; ++t) {
    const int li = l + (t - taps / 2) * dil;
    if (li < 0 || li >= L) continue;
    const __nv_bfloat16* a = A + ((int64_t)s * L + li) * K;
    const __nv_bfloat16* w = W + ((int64_t)t * N + n) * K;
    for (int k = 0; k < K; ++k) acc += __bfloat162float(a[k]) * __bfloat162float(w[k]);
  }
  C[idx] = acc + (bias ? bias[n] : 0.0f);
}
}  // namespace
}  // namespace svdd

using namespace svdd;

extern "C" int svdd_selftest_conv_gemm(const void* A_bf16, const void* W_bf16, const float* bias,
                                       float* C, int S, int L, int K, int N, int taps, int dil,
                                       int use_tensor_cores, void* stream) {
  SVDD_CHECK_ARG(A_bf16 && W_bf16 && C, "selftest: null pointer");
  int dev = 0;
  SVDD_CUDA(cudaGetDevice(&dev));
  SVDD_TRY(svdd_device_check(dev));
  cudaStream_t st = (cudaStream_t)stream;
  if (!use_tensor_cores) {
    const int64_t total = (int64_t)S * L * N;
    naive_conv_gemm_kernel<<<(unsigned)ceil_div<int64_t>(total, 256), 256, 0, st>>>(
        (const __nv_bfloat16*)A_bf16, (const __nv_bfloat16*)W_bf16, bias, C, S, L, K, N, taps, dil);
    count_launch();
    SVDD_LAUNCH_CHECK();
    return SVDD_OK;
  }
  GemmShape g;
  g.K = K; g.N = N; g.taps = taps; g.dil = dil;
  if (taps == 1) {  // plain GEMM: flatten rows
    g.S = 1; g.L = S * L; g.L_in = S * L; g.BL = 128; g.BS = 1;
  } else {
    g.S = S; g.L = L; g.L_in = L;
    choose_row_tiling(L, taps, &g);
  }
  EpiParams ep;
  ep.bias = bias;
  ep.out = C; ep.out_dtype = DT_F32; ep.ld_out = N;
  return launch_conv_gemm(A_bf16, W_bf16, g, EPI_GENERIC, ep, st);
}

// Fused epilogue chain of EPI_GENERIC in isolation (tests only):
//   v = acc*scale+shift ; v += bias ; [act] ; v += res ; [act] -> out ; out2 = act2(v*scale2+shift2)
// dtypes: 1 = bf16, 2 = fp32 (svdd::DType).  `flat` treats A as [S*L, K] rows (plain GEMM).
extern "C" int svdd_selftest_gemm_epilogue(const void* A_bf16, const void* W_bf16, const float* bias,
                                           const float* scale, const float* shift, int act, int act_after_res,
                                           const void* res, int res_dtype, void* out, int out_dtype,
                                           void* out2, int out2_dtype, const float* scale2,
                                           const float* shift2, int act2, int S, int L, int K, int N,
                                           int taps, int dil, int flat, void* stream) {
  SVDD_CHECK_ARG(A_bf16 && W_bf16, "selftest_gemm_epilogue: null pointer");
  int dev = 0;
  SVDD_CUDA(cudaGetDevice(&dev));
  SVDD_TRY(svdd_device_check(dev));
  GemmShape g;
  g.K = K; g.N = N; g.taps = taps; g.dil = dil;
  if (flat == 2) {
    // halo mode: A is [S*L, K] flat rows that already carry taps/2 zero rows between sequences
    g.S = 1; g.L = S * L; g.L_in = S * L; g.BL = 128; g.BS = 1; g.halo = 1;
  } else if (flat) {
    SVDD_CHECK_ARG(taps == 1, "selftest_gemm_epilogue: flat rows need taps == 1");
    g.S = 1; g.L = S * L; g.L_in = S * L; g.BL = 128; g.BS = 1;
  } else {
    g.S = S; g.L = L; g.L_in = L;
    choose_row_tiling(L, taps, &g);
  }
  EpiParams ep;
  ep.bias = bias; ep.scale = scale; ep.shift = shift; ep.act = act; ep.act_after_res = act_after_res;
  ep.res = res; ep.res_dtype = res_dtype; ep.ld_res = N;
  ep.out = out; ep.out_dtype = out_dtype; ep.ld_out = N;
  ep.out2 = out2; ep.out2_dtype = out2_dtype; ep.ld_out2 = N;
  ep.scale2 = scale2; ep.shift2 = shift2; ep.act2 = act2;
  return launch_conv_gemm(A_bf16, W_bf16, g, EPI_GENERIC, ep, (cudaStream_t)stream);
}

// Attention pooling (EPI_POOL) in isolation: y bf16 [S, L_in, C], Wp bf16 [C, C] ->
// out fp32 [S * ceil(L_in/2), C].
extern "C" int svdd_selftest_pool(const void* y_bf16, const void* Wp_bf16, float* out, int S,
                                  int L_in, int C, void* stream) {
  SVDD_CHECK_ARG(y_bf16 && Wp_bf16 && out, "selftest_pool: null pointer");
  int dev = 0;
  SVDD_CUDA(cudaGetDevice(&dev));
  SVDD_TRY(svdd_device_check(dev));
  GemmShape g;
  g.S = S; g.L = (L_in + 1) / 2; g.L_in = L_in; g.K = C; g.N = C; g.taps = 1; g.dil = 1;
  choose_row_tiling(g.L, 1, &g);
  EpiParams ep;
  ep.pool_vals = y_bf16;
  ep.out = out; ep.out_dtype = DT_F32; ep.ld_out = C;
  return launch_conv_gemm(y_bf16, Wp_bf16, g, EPI_POOL, ep, (cudaStream_t)stream);
}

// The pair-split 1x1 conv + difference pooling (EPI_PAIR -> EPI_POOL2) in isolation:
//   y = A.W1^T + bias + res  (never materialised);  y0 = y[2j], yd = y[2j+1] - y[2j];
//   pooled[j] = y0 + sigmoid(Wp.yd) * yd  = softmax-weighted sum over the pair.
// A, res bf16 [S, L, C]; W1, Wp bf16 [C, C]; y0, yd bf16 [S, ceil(L/2), C]; pooled fp32 (optional)
// and/or pooled_act bf16 = GELU(pooled * scale2 + shift2) (optional), both [S*ceil(L/2), C].
extern "C" int svdd_selftest_pair_pool(const void* A_bf16, const void* W1_bf16, const float* bias,
                                       const void* res_bf16, const void* Wp_bf16, void* y0_bf16,
                                       void* yd_bf16, float* pooled_f32, void* pooled_act_bf16,
                                       const float* scale2, const float* shift2, int S, int L, int C,
                                       void* stream) {
  SVDD_CHECK_ARG(A_bf16 && W1_bf16 && res_bf16 && Wp_bf16 && y0_bf16 && yd_bf16, "selftest_pair_pool: null pointer");
  SVDD_CHECK_ARG(pooled_f32 || pooled_act_bf16, "selftest_pair_pool: no output requested");
  SVDD_CHECK_ARG(L % 2 == 0 || L + 1 <= 128, "selftest_pair_pool: odd lengths above 127 are not supported");
  int dev = 0;
  SVDD_CUDA(cudaGetDevice(&dev));
  SVDD_TRY(svdd_device_check(dev));
  const int Lo = (L + 1) / 2;
  {
    GemmShape g;
    g.K = C; g.N = C; g.taps = 1; g.dil = 1;
    if (L % 2 == 0) { g.S = 1; g.L = S * L; g.L_in = g.L; g.BL = 128; g.BS = 1; }
    else { g.S = S; g.L = L; g.L_in = L; g.BL = L + 1; g.BS = 128 / (L + 1); }
    EpiParams ep;
    ep.bias = bias;
    ep.res = res_bf16; ep.res_dtype = DT_BF16; ep.ld_res = C;
    ep.out = y0_bf16; ep.out_dtype = DT_BF16; ep.ld_out = C;
    ep.out2 = yd_bf16; ep.out2_dtype = DT_BF16; ep.ld_out2 = C;
    SVDD_TRY(launch_conv_gemm(A_bf16, W1_bf16, g, EPI_PAIR, ep, (cudaStream_t)stream));
  }
  GemmShape g;
  g.S = 1; g.L = S * Lo; g.L_in = g.L; g.K = C; g.N = C; g.taps = 1; g.dil = 1; g.BL = 128; g.BS = 1;
  EpiParams ep;
  ep.res = y0_bf16; ep.res_dtype = DT_BF16; ep.ld_res = C;
  ep.res2 = yd_bf16; ep.ld_res2 = C;
  if (pooled_f32) { ep.out = pooled_f32; ep.out_dtype = DT_F32; ep.ld_out = C; }
  if (pooled_act_bf16) {
    ep.out2 = pooled_act_bf16; ep.out2_dtype = DT_BF16; ep.ld_out2 = C;
    ep.scale2 = scale2; ep.shift2 = shift2; ep.act2 = ACT_GELU;
  }
  return launch_conv_gemm(yd_bf16, Wp_bf16, g, EPI_POOL2, ep, (cudaStream_t)stream);
}
