// One whole reverse step behind a single C call: stage 1 -> 2 -> 3 -> 4 on the caller's stream.
// Replaces the body of Diffusion._ddpm_update_finetune_controlled (diffusion_gosai.py:1175-1228,
// SVDD-MC) and _ddpm_update_finetune_controlled_twedie (:1374-1460, SVDD-PM) for a host that does
// not want to sequence the five entry points itself.  Pure composition of the exported stage
// functions (same kernels, same order as svdd_b200/diffusion_gosai.py:_trajectory), so the tokens
// are bit-identical to calling them one by one; capture-safe (no allocation, no sync).
#include <stdint.h>

#include "common.cuh"

using namespace svdd;

namespace {

inline size_t align256(size_t n) { return (n + 255) & ~(size_t)255; }
inline size_t tok_bytes(int tok_dtype) { return tok_dtype == SVDD_TOK_I64 ? 8 : 1; }

struct Layout {
  size_t logits, cand, scores, logits2, x0, scratch, scratch_bytes, total;
};

int check(const svdd_step_args* a) {
  SVDD_CHECK_ARG(a != nullptr, "svdd_step: null argument block");
  SVDD_CHECK_ARG(a->B >= 0 && a->L >= 1 && a->M >= 1, "svdd_step: bad shape B=%d L=%d M=%d", a->B, a->L, a->M);
  SVDD_CHECK_ARG(a->tok_dtype == SVDD_TOK_I64 || a->tok_dtype == SVDD_TOK_U8, "svdd_step: bad tok_dtype %d", a->tok_dtype);
  SVDD_CHECK_ARG(a->denoiser != nullptr && a->scorer != nullptr, "svdd_step: denoiser and scorer handles are required");
  SVDD_CHECK_ARG(a->scorer_kind == SVDD_SCORER_CONVGRU || a->scorer_kind == SVDD_SCORER_ENFORMER,
                 "svdd_step: bad scorer_kind %d", a->scorer_kind);
  return SVDD_OK;
}

size_t scorer_ws(const svdd_step_args* a, int64_t rows) {
  return a->scorer_kind == SVDD_SCORER_CONVGRU
             ? svdd_convgru_workspace_bytes(static_cast<const svdd_convgru*>(a->scorer), rows, a->L)
             : svdd_enformer_workspace_bytes(static_cast<const svdd_enformer*>(a->scorer), rows, a->L);
}

Layout layout(const svdd_step_args* a) {
  Layout o = {};
  const size_t B = (size_t)a->B, L = (size_t)a->L, M = (size_t)a->M, tb = tok_bytes(a->tok_dtype);
  size_t off = 0;
  o.logits = off; off += align256(B * L * kVocab * sizeof(float));
  o.cand = off;   off += a->cand_out ? 0 : align256(M * B * L * tb);
  o.scores = off; off += a->scores_out ? 0 : align256(M * B * sizeof(float));
  if (a->tweedie) {
    o.logits2 = off; off += align256(M * B * L * kVocab * sizeof(float));
    o.x0 = off;      off += align256(M * B * L * tb);
  }
  // the networks run one after the other and share one scratch region
  size_t s = svdd_denoiser_workspace_bytes(a->denoiser, (int64_t)B, a->L);
  if (a->tweedie) {
    const size_t s2 = svdd_denoiser_workspace_bytes(a->denoiser, (int64_t)(M * B), a->L);
    if (s2 > s) s = s2;
  }
  const size_t s3 = scorer_ws(a, (int64_t)(M * B));
  if (s3 > s) s = s3;
  o.scratch = off;
  o.scratch_bytes = align256(s > 16 ? s : 16);
  o.total = off + o.scratch_bytes;
  return o;
}

}  // namespace

extern "C" size_t svdd_step_workspace_bytes(const svdd_step_args* a) {
  if (check(a) != SVDD_OK) return 0;
  return layout(a).total;
}

extern "C" int svdd_step(const svdd_step_args* a) {
  SVDD_TRY(check(a));
  if (a->B == 0) return SVDD_OK;                 // empty batch: the reference's loops do nothing
  SVDD_CHECK_ARG(a->x != nullptr && a->x_out != nullptr, "svdd_step: null token pointer");
  SVDD_CHECK_ARG(a->time_bias_t != nullptr, "svdd_step: time_bias_t is required (the rows for sigma_t; constant without time conditioning)");
  const Layout o = layout(a);
  if (a->ws == nullptr || a->ws_bytes < o.total) {
    set_last_error("svdd_step: workspace too small (%zu < %zu)", a->ws_bytes, o.total);
    return SVDD_ERR_WORKSPACE_TOO_SMALL;
  }
  uint8_t* w = static_cast<uint8_t*>(a->ws);
  SVDD_CHECK_ARG((reinterpret_cast<uintptr_t>(w) & 255) == 0, "svdd_step: workspace must be 256-byte aligned");
  float* logits = reinterpret_cast<float*>(w + o.logits);
  void* cand = a->cand_out ? a->cand_out : static_cast<void*>(w + o.cand);
  float* scores = a->scores_out ? a->scores_out : reinterpret_cast<float*>(w + o.scores);
  void* scratch = w + o.scratch;
  const int B = a->B, L = a->L, M = a->M;
  const int64_t MB = (int64_t)M * B;

  // stage 1: p(x0 | x_t)  (diffusion_gosai.py:1189-1190 / 1388-1389)
  SVDD_TRY(svdd_denoiser_forward(a->denoiser, a->x, a->tok_dtype, a->time_bias_t, logits, B, L, scratch,
                                 o.scratch_bytes, a->stream));
  // stage 2: SUBS, q_xs, M Gumbel-max draws, carry-over  (:1194-1203 / 1393-1402)
  SVDD_TRY(svdd_subs_sample(logits, 0, a->x, a->tok_dtype, a->U, a->seed, a->seed_dev, a->step, a->row_offset,
                            a->mc_t, a->mc_s, cand, a->q_out, B, L, M, a->stream));
  const void* scored = cand;
  if (a->tweedie) {
    // SVDD-PM: x0 estimate of every candidate (:1413-1419), scored by the reward oracle (:1430)
    float* logits2 = reinterpret_cast<float*>(w + o.logits2);
    void* x0 = w + o.x0;
    const float* tb_s = a->time_bias_s ? a->time_bias_s : a->time_bias_t;
    SVDD_TRY(svdd_denoiser_forward(a->denoiser, cand, a->tok_dtype, tb_s, logits2, MB, L, scratch, o.scratch_bytes,
                                   a->stream));
    SVDD_TRY(svdd_x0_argmax(logits2, cand, a->tok_dtype, x0, MB, L, a->stream));
    scored = x0;
  }
  // stage 3: one scoring call over all M * B candidates (:1207-1209 / 1430)
  if (a->scorer_kind == SVDD_SCORER_CONVGRU)
    SVDD_TRY(svdd_convgru_score(static_cast<svdd_convgru*>(a->scorer), scored, a->tok_dtype, scores, MB, L, scratch,
                                o.scratch_bytes, a->stream));
  else
    SVDD_TRY(svdd_enformer_score(static_cast<svdd_enformer*>(a->scorer), scored, a->tok_dtype, scores, MB, L, scratch,
                                 o.scratch_bytes, a->stream));
  // stage 4: softmax -> argmax (or soft resampling) -> gather  (:1219-1227 / 1451-1459)
  return svdd_select_gather(scores, cand, a->tok_dtype, a->alpha, a->U_sel, a->seed, a->seed_dev, a->step,
                            a->row_offset, a->x_out, a->idx_out, B, L, M, a->stream);
}
