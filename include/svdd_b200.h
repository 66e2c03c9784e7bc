/* libsvdd_b200 -- C ABI of the B200-native SVDD decoding engine.
 *
 * The reference (masa-ue/SVDD) is pure Python/PyTorch and has no FFI; the seam
 * this library slots under is the set of Python methods of
 * `diffusion_gosai.Diffusion` that one reverse SVDD step is made of.  Every
 * entry point below names the reference code (file:line under /root/reference)
 * it replaces.  INTEGRATION.md shows the ctypes binding a reference maintainer
 * would add.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types in any signature;
 *  - every function returns SVDD_OK (0) or a negative svdd_status;
 *    svdd_last_error() returns a thread-local, human-readable message;
 *  - all tensor arguments are DEVICE pointers owned by the caller (e.g. torch
 *    allocations); the library never frees or retains them past the call.  The
 *    only memory it owns are packed weights inside handles it created;
 *  - every launch goes onto the `stream` argument (a cudaStream_t passed as
 *    void*; NULL = legacy default stream).  No hidden synchronisation, no host
 *    allocation on the hot path: scratch memory comes from the caller
 *    (`*_workspace_bytes` + `ws`), so calls are re-entrant across streams given
 *    distinct workspaces and can be captured into CUDA graphs;
 *  - there is no CPU fallback: on a device that is not sm_100 every compute
 *    entry point fails with SVDD_ERR_UNSUPPORTED_DEVICE.
 *
 * Token tensors: `tok_dtype` selects SVDD_TOK_I64 (the reference's int64) or
 * SVDD_TOK_U8 (the engine's compact internal state).  Token values: 0..3 =
 * A,C,G,T ; 4 = MASK (diffusion_gosai.py:85,94-95).
 */
#ifndef SVDD_B200_H_
#define SVDD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVDD_B200_VERSION 100 /* 0.1.0 */

typedef enum svdd_status {
  SVDD_OK = 0,
  SVDD_ERR_INVALID_ARGUMENT = -1,
  SVDD_ERR_CUDA = -2,
  SVDD_ERR_UNSUPPORTED_DEVICE = -3,
  SVDD_ERR_WORKSPACE_TOO_SMALL = -4,
  SVDD_ERR_MISSING_TENSOR = -5,
  SVDD_ERR_INTERNAL = -6
} svdd_status;

typedef enum svdd_tok_dtype { SVDD_TOK_I64 = 0, SVDD_TOK_U8 = 1 } svdd_tok_dtype;

/* A named fp32 device tensor: one entry of a (reference-layout) state_dict.
 * Handles are built from arrays of these, looked up by the reference's own key
 * names (e.g. "convs.3.weight", "conv_tower.blocks.2.0.conv.bias"). */
typedef struct svdd_tensor {
  const char* name;
  const float* data; /* device pointer, contiguous, fp32 */
  int32_t ndim;
  int64_t shape[4];
} svdd_tensor;

/* ---- library ------------------------------------------------------------ */
int svdd_version(void);
const char* svdd_last_error(void);
/* SVDD_OK iff `device` is a compute-capability 10.x GPU. */
int svdd_device_check(int device);
/* Number of kernels this library has launched in this process (monotonic);
 * bench.py reports the delta over the timed region as "gpu_launches". */
int64_t svdd_launch_count(void);

/* Per-launch CUDA-event timing of the tensor-core GEMM kernel (measurement aid for
 * bench.py's roofline leg; adds two event records per launch, so it is only
 * enabled around a dedicated pass).  svdd_profile_end returns the summed device
 * time, the number of launches and their nominal dense FLOPs. */
int svdd_profile_begin(void);
int svdd_profile_end(double* gemm_ms, int64_t* gemm_launches, double* gemm_flops);
/* The longest single GEMM launch of the last svdd_profile_begin/end window: its device time,
 * nominal FLOPs and shape (rows = S*L, K, N, taps, epilogue mode). */
int svdd_profile_top(double* ms, double* flops, int64_t* rows, int* K, int* N, int* taps, int* mode);

/* ---- stage 2: SUBS + move-chance mixing + Gumbel-max draw + carry-over ----
 * Replaces, fused into one kernel:
 *   Diffusion._subs_parameterization          diffusion_gosai.py:286-304
 *   q_xs = exp(log_p)*(mc_t-mc_s); q[..,4]=mc_s   diffusion_gosai.py:1194-1196 (MC), 1393-1395 (PM), 1166-1169 (plain)
 *   _sample_categorical x M                   diffusion_gosai.py:30-34, 1203 / 1402 / 1170
 *   copy_flag*x + (1-copy_flag)*draw          diffusion_gosai.py:1199-1203
 *
 * logits [B,L,5] fp32 raw backbone output (is_log_p = 0) or post-SUBS
 *        log-probs as returned by Diffusion.forward (is_log_p = 1);
 * x      [B,L] current tokens; cand [M,B,L] output candidates (same dtype);
 * U      [M,B,L,5] fp32 uniforms in [0,1) in the reference's draw order (one
 *        rand_like per candidate, m-major), or NULL to use the in-kernel
 *        Philox4x32-10 stream keyed by (seed, step, global row = row_offset+b);
 * seed_dev optional DEVICE uint64[2] (or NULL): [0] is added to `seed` and [1] to
 *        `row_offset` at run time, so a CUDA graph that captured this call can be
 *        replayed with fresh noise and for another block of global rows;
 * mc_t, mc_s  fp32 move chances 1-exp(-sigma(t)), 1-exp(-sigma(t-dt)) computed
 *        on the host with the reference expression (diffusion_gosai.py:1176-1187);
 * q_out  optional [B,L,5] fp32: the q_xs tensor the reference returns (or NULL).
 */
int svdd_subs_sample(const float* logits, int is_log_p, const void* x,
                     int tok_dtype, const float* U, uint64_t seed,
                     const uint64_t* seed_dev, int step, int64_t row_offset,
                     float mc_t, float mc_s, void* cand, float* q_out, int B,
                     int L, int M, void* stream);

/* ---- stage 4: per-sequence selection + gather ----------------------------
 * Replaces diffusion_gosai.py:1219-1227 / 1451-1459:
 *   scores = stack(scores, dim=1); softmax(dim=1); argmax(dim=1);
 *   final = stack([samples[idx[j]][j,:] for j in range(B)])   (B host syncs)
 * scores [M,B] fp32, candidate-major (the order one scoring call over the
 *        [M,B,L] candidate tensor produces);
 * alpha == 0: idx = argmax(softmax(scores)) with first-index ties (reference);
 * alpha  > 0: idx = gumbel_argmax(softmax(scores/alpha), U_sel) -- build
 *        decision, the reference's multinomial lines are commented out;
 *        U_sel [B,M] fp32 or NULL for the Philox stream;
 * x_out  [B,L] selected rows; idx_out optional int32 [B].
 */
int svdd_select_gather(const float* scores, const void* cand, int tok_dtype,
                       float alpha, const float* U_sel, uint64_t seed,
                       const uint64_t* seed_dev, int step, int64_t row_offset,
                       void* x_out, int32_t* idx_out, int B, int L, int M,
                       void* stream);

/* ---- SUBS parameterisation alone --------------------------------------------
 * Replaces Diffusion._subs_parameterization (diffusion_gosai.py:286-304) as the
 * tail of Diffusion.forward (:339-357): raw logits [N,L,5] + tokens x [N,L] ->
 * log-probs [N,L,5] (mask column -1e6 - lse; unmasked rows -1e6 / 0). */
int svdd_subs_log_p(const float* logits, const void* x, int tok_dtype,
                    float* log_p, int64_t n_rows, int L, void* stream);

/* ---- argmax of post-SUBS log-probs ----------------------------------------
 * Replaces the noise-removal step diffusion_gosai.py:1049-1060
 *   x = forward(x, sigma)[:, :, :-1].argmax(-1)
 * and the Tweedie x0 estimate diffusion_gosai.py:1415-1419
 *   argmax(forward(cand, sigma_s), dim=2) combined with carried tokens.
 * logits [N,L,5] raw backbone output for tokens x [N,L]; out [N,L] in 0..3. */
int svdd_x0_argmax(const float* logits, const void* x, int tok_dtype, void* out,
                   int64_t n_rows, int L, void* stream);

/* ---- stage 1: MDLM denoiser (models/dnaconv.py:135-210 CNNModel.forward) ---
 * Tensors (names relative to the `backbone.` prefix of the Lightning
 * checkpoint, shapes as in the reference):
 *   linear.{weight[H,5,9],bias[H]}  convs.i.{weight[H,H,9],bias[H]}
 *   norms.i.{weight[H],bias[H]}     final_conv.0.{weight[H,H,1],bias[H]}
 *   final_conv.2.{weight[5,H,1],bias[5]}
 * H must be 128, n_layers = 5*num_cnn_stacks with dilation groups
 * [1,1,4,16,64] (models/dnaconv.py:155-160).  The per-layer time bias
 * time_layers[i](relu(time_embedder(sigma))) is a [n_layers,H] fp32 row block
 * computed by the host (constant when time_conditioning is False,
 * diffusion_gosai.py:334-335) and passed to each forward.
 * Arithmetic: bf16 operands, fp32 accumulation (tcgen05), fp32 LayerNorm. */
typedef struct svdd_denoiser svdd_denoiser;
int svdd_denoiser_create(const svdd_tensor* tensors, int n_tensors,
                         int num_cnn_stacks, void* stream, svdd_denoiser** out);
void svdd_denoiser_destroy(svdd_denoiser* h);
size_t svdd_denoiser_workspace_bytes(const svdd_denoiser* h, int64_t n_rows, int L);
/* tokens [N,L] -> logits [N,L,5] fp32 (raw, pre-SUBS). */
int svdd_denoiser_forward(svdd_denoiser* h, const void* tokens, int tok_dtype,
                          const float* time_bias, float* logits, int64_t n_rows,
                          int L, void* ws, size_t ws_bytes, void* stream);

/* ---- stage 1, alternative backbone: DiT (models/dit.py:214-369) -----------
 * Replaces DIT.forward (models/dit.py:355-366: vocab_embed -> n_blocks x DDiTBlock ->
 * DDitFinalLayer) for `backbone: dit` (diffusion_gosai.py:102-104).  Tensors, names relative
 * to the `backbone.` prefix, shapes as in the reference:
 *   vocab_embed.embedding[5,H]       rotary_emb.inv_freq[32]
 *   blocks.i.attn_qkv.weight[3H,H]   blocks.i.attn_out.weight[H,H]
 *   blocks.i.mlp.0.{weight[4H,H],bias[4H]}  blocks.i.mlp.2.weight[H,4H]
 *   output_layer.linear.{weight[5,H],bias[5]}
 * H = 64 * n_heads, a multiple of 128.  The conditioning c = silu(sigma_map(sigma)) is
 * batch-constant on the decode path, so everything that depends on it -- the six
 * adaLN_modulation vectors of each block and the two of the final layer, folded with the
 * LayerNorm weights and mlp.2.bias -- is computed by the host and passed per call as `mod`
 * (fp32, device, svdd_dit_mod_floats() elements; layout documented at svdd_dit_forward in
 * csrc/dit.cu and built by svdd_b200/dit.py:DIT.modulation).
 * Arithmetic as in the reference's autocast region (:362): bf16 GEMM operands and attention,
 * fp32 accumulation, LayerNorm and residual stream. */
typedef struct svdd_dit svdd_dit;
int svdd_dit_create(const svdd_tensor* tensors, int n_tensors, int n_heads, void* stream,
                    svdd_dit** out);
void svdd_dit_destroy(svdd_dit* h);
int64_t svdd_dit_mod_floats(const svdd_dit* h);
size_t svdd_dit_workspace_bytes(const svdd_dit* h, int64_t n_rows, int L);
/* tokens [N,L] -> logits [N,L,5] fp32 (raw, pre-SUBS); L <= 1024. */
int svdd_dit_forward(svdd_dit* h, const void* tokens, int tok_dtype, const float* mod,
                     float* logits, int64_t n_rows, int L, void* ws, size_t ws_bytes,
                     void* stream);

/* ---- stage 3a: RNA value net / reward oracle (ConvGRU) --------------------
 * Replaces head(embedding(transform_samples(tokens).float())).squeeze()
 * (diffusion_gosai.py:1208-1209, 1462-1470) for
 * ConvGRUTrunk.forward Enformer.py:1411-1426 + ConvHead.forward :2166-2173,
 * and reward_model(onehot.float().transpose(1,2))[:,0] (diffusion_gosai.py:1430)
 * for the gReLU ConvGRU oracle (no BN / no residual variant).
 * Tensors: the trunk's state_dict under its own key names plus the head's
 * under the prefix "head." (head.channel_transform.conv.layer.{weight,bias}). */
typedef struct svdd_convgru svdd_convgru;
int svdd_convgru_create(const svdd_tensor* tensors, int n_tensors, void* stream,
                        svdd_convgru** out);
void svdd_convgru_destroy(svdd_convgru* h);
size_t svdd_convgru_workspace_bytes(const svdd_convgru* h, int64_t n_rows, int L);
/* tokens [N,L] (4 -> all-zero one-hot row) -> scores [N] fp32. */
int svdd_convgru_score(svdd_convgru* h, const void* tokens, int tok_dtype,
                       float* scores, int64_t n_rows, int L, void* ws,
                       size_t ws_bytes, void* stream);

/* ---- stage 3b: DNA value net / reward oracle (Enformer-style) -------------
 * Same call shape as 3a for EnformerTrunk.forward Enformer.py:1326-1334
 * (conv tower :1807-1884, transformer tower :1887-2007, pointwise :1315-1323)
 * + ConvHead.  `n_heads` as passed to EnformerTrunk (decode.py:78). */
typedef struct svdd_enformer svdd_enformer;
int svdd_enformer_create(const svdd_tensor* tensors, int n_tensors, int n_heads,
                         void* stream, svdd_enformer** out);
void svdd_enformer_destroy(svdd_enformer* h);
size_t svdd_enformer_workspace_bytes(const svdd_enformer* h, int64_t n_rows, int L);
int svdd_enformer_score(svdd_enformer* h, const void* tokens, int tok_dtype,
                        float* scores, int64_t n_rows, int L, void* ws,
                        size_t ws_bytes, void* stream);

/* ---- one whole reverse step ------------------------------------------------
 * Replaces the body of Diffusion._ddpm_update_finetune_controlled
 * (diffusion_gosai.py:1175-1228, SVDD-MC: tweedie = 0) and of
 * _ddpm_update_finetune_controlled_twedie (:1374-1460 with options == "True",
 * SVDD-PM: tweedie = 1) by ONE call that sequences stages 1-4 on `stream`:
 *   MC:  denoiser(x) -> subs_sample (M candidates) -> scorer(candidates) -> select_gather
 *   PM:  denoiser(x) -> subs_sample -> denoiser(candidates) -> x0_argmax -> scorer(x0) -> select_gather
 * It is a pure composition of the entry points above (same kernels, same order), so the tokens are
 * bit-identical to calling them one by one; no allocation and no synchronisation, i.e. safe to
 * capture into a CUDA graph.  The return tuple of the reference's step functions maps to
 * x_out (final_samples), x (unchanged), q_out (q_xs, optional); copy_flag is (x != 4).
 *
 * denoiser       the CNN backbone; time_bias_t / time_bias_s: its time-conditioning rows for
 *                sigma(t) and sigma(t - dt) (the same constant rows without time conditioning,
 *                diffusion_gosai.py:334-335; time_bias_s may be NULL = time_bias_t);
 * scorer         svdd_convgru* or svdd_enformer* according to scorer_kind: the value net
 *                (head(embedding(.)), :1207-1209) for MC, the reward oracle (:1430) for PM;
 * x, x_out       [B,L] tokens in / selected tokens out; idx_out optional int32 [B];
 * mc_t, mc_s, alpha, U, U_sel, seed, seed_dev, step, row_offset: as in svdd_subs_sample /
 *                svdd_select_gather;
 * cand_out       optional [M,B,L] (the candidates; workspace if NULL), scores_out optional [M,B],
 *                q_out optional [B,L,5];
 * ws             256-byte aligned device workspace of svdd_step_workspace_bytes(args) bytes
 *                (logits, candidates, scores, PM buffers and the networks' shared scratch). */
typedef enum svdd_scorer_kind { SVDD_SCORER_CONVGRU = 0, SVDD_SCORER_ENFORMER = 1 } svdd_scorer_kind;
typedef struct svdd_step_args {
  svdd_denoiser* denoiser;
  const float* time_bias_t;
  const float* time_bias_s;
  int scorer_kind;
  void* scorer;
  int tweedie;
  const void* x;
  void* x_out;
  int tok_dtype;
  int32_t* idx_out;
  int B, L, M;
  float mc_t, mc_s, alpha;
  const float* U;
  const float* U_sel;
  uint64_t seed;
  const uint64_t* seed_dev;
  int step;
  int64_t row_offset;
  void* cand_out;
  float* scores_out;
  float* q_out;
  void* ws;
  size_t ws_bytes;
  void* stream;
} svdd_step_args;
size_t svdd_step_workspace_bytes(const svdd_step_args* args);
int svdd_step(const svdd_step_args* args);

/* Test hook for stage 2: same contract as svdd_subs_sample, but every Gumbel-max draw is
 * taken on the exact logf + IEEE-division path (the reference's arithmetic, operation for
 * operation).  svdd_subs_sample decides a draw with one approximate log per element only
 * when an error-bound test proves the exact argmax; tests assert both entry points agree
 * bit for bit, including on adversarial near-ties. */
int svdd_selftest_subs_sample_exact(const float* logits, int is_log_p, const void* x,
                                    int tok_dtype, const float* U, uint64_t seed,
                                    const uint64_t* seed_dev, int step, int64_t row_offset,
                                    float mc_t, float mc_s, void* cand, float* q_out, int B,
                                    int L, int M, void* stream);

/* ---- self test of the tcgen05 implicit-GEMM building block ----------------
 * C[r,n] = sum_{tap,k} A[seq(r), l(r)+(tap-taps/2)*dil, k] * W[tap,n,k]
 * (zero outside [0,L)) + bias[n], bf16 operands, fp32 accumulate, fp32 out.
 * A [S,L,K] bf16, W [taps,N,K] bf16, bias [N] fp32, C [S*L,N] fp32.  Used by
 * tests to check the tensor-core path against a plain CUDA-core kernel. */
int svdd_selftest_conv_gemm(const void* A_bf16, const void* W_bf16,
                            const float* bias, float* C, int S, int L, int K,
                            int N, int taps, int dil, int use_tensor_cores,
                            void* stream);

/* The fused epilogue chain of the GEMM in isolation (tests only):
 *   v = acc*scale+shift ; v += bias ; [act] ; v += res ; [act]  -> out
 *   out2 = act2(v*scale2 + shift2)
 * act: 0 none, 1 ReLU, 2 Enformer GELU x*sigmoid(1.702x); dtypes: 1 bf16, 2 fp32;
 * any of bias/scale/res/out/out2 may be NULL; res may alias out (in-place residual
 * stream).  flat != 0 treats A as [S*L, K] rows of a plain GEMM. */
int svdd_selftest_gemm_epilogue(const void* A_bf16, const void* W_bf16, const float* bias,
                                const float* scale, const float* shift, int act,
                                int act_after_res, const void* res, int res_dtype,
                                void* out, int out_dtype, void* out2, int out2_dtype,
                                const float* scale2, const float* shift2, int act2, int S,
                                int L, int K, int N, int taps, int dil, int flat,
                                void* stream);

/* Unit-test hooks for the fused pieces of stage 3b (used by tests only):
 *  - attention pooling over position pairs (enformer_pytorch AttentionPool,
 *    Enformer.py:2447): y bf16 [S,L_in,C], Wp bf16 [C,C] -> fp32 [S*ceil(L_in/2),C];
 *  - the Enformer relative-position basis (host, [2n-1,F] fp32);
 *  - attention over n positions with relative-position logits: qkv fp32
 *    [rows*n, 2*H*dk+H*dv], rel_content_bias/rel_pos_bias [H*dk], relk
 *    [H,2n-1,dk] -> bf16 [rows*n, H*dv]. */
int svdd_selftest_pool(const void* y_bf16, const void* Wp_bf16, float* out, int S,
                       int L_in, int C, void* stream);
/* The pair-split 1x1 conv + difference pooling used by svdd_enformer_score (tests only):
 * y = A.W1^T + bias + res is never materialised; y0 = y[2j], yd = y[2j+1] - y[2j] (0 for the
 * unpaired tail of an odd L); pooled[j] = y0 + sigmoid(Wp.yd) * yd -- AttentionPool(2) of
 * enformer_pytorch (Enformer.py:2447) since softmax over a pair is the sigmoid of the logit
 * difference and Wp.y1 - Wp.y0 = Wp.(y1 - y0).  A, res bf16 [S,L,C]; W1, Wp bf16 [C,C];
 * y0, yd bf16 [S,ceil(L/2),C]; pooled fp32 and/or pooled_act = GELU(pooled*scale2+shift2) bf16. */
int svdd_selftest_pair_pool(const void* A_bf16, const void* W1_bf16, const float* bias,
                            const void* res_bf16, const void* Wp_bf16, void* y0_bf16,
                            void* yd_bf16, float* pooled_f32, void* pooled_act_bf16,
                            const float* scale2, const float* shift2, int S, int L, int C,
                            void* stream);
int svdd_selftest_rel_positions(int n, int F, float* out_host);
int svdd_selftest_attention(const float* qkv, const float* rcb, const float* rpb,
                            const float* relk, void* out_bf16, int64_t rows, int n,
                            int H, int dk, int dv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVDD_B200_H_ */
