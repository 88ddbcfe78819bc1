/*
 * care_b200 — C ABI of the B200-native CARE caption-decode hot path.
 *
 * The reference (yangbang18/CARE) is pure Python/PyTorch: it has no FFI of its own.  The
 * boundary this library replaces is therefore the set of PyTorch calls the reference issues on
 * its inference path; each entry point below names the reference lines it stands in for
 * (paths relative to the reference tree).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: raw device pointers + explicit sizes/strides, no torch types.
 *   - the caller (PyTorch on the host side) allocates and owns every buffer; the library never
 *     frees or keeps a pointer beyond the call, except TMA descriptors cached inside the ctx.
 *   - every call only enqueues work on the caller's stream (`stream` is a cudaStream_t passed
 *     as void*); there is no hidden device synchronisation.
 *   - return value: 0 = ok, negative = argument/shape error, positive = cudaError_t / CUresult;
 *     care_last_error() returns a thread-local message.
 *   - dtype codes: CARE_F32 = 0, CARE_BF16 = 1, CARE_F16 = 2.  "T" below means the activation type of
 *     the precision mode: fp32 in the bit-exact parity mode, the library's 16-bit type in the
 *     throughput mode.  One build of the library has ONE 16-bit type (care_h16_dtype()):
 *     libcare_b200.so computes in IEEE fp16, libcare_b200_bf16.so (same sources, -DCARE_USE_BF16) in
 *     bf16; a call with the other 16-bit code fails with an argument error.
 *   - rows of decoder-side tensors are ordered (video, beam): row = video * K + beam.
 *   - sm_100a only.  There is no CPU path: every entry point fails if no device is present.
 */
#ifndef CARE_B200_H_
#define CARE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CARE_F32 0
#define CARE_BF16 1
#define CARE_F16 2

#define CARE_ACT_NONE 0
#define CARE_ACT_RELU 1

#define CARE_PAD 0 /* config/Constants.py:1-6 */
#define CARE_BOS 2
#define CARE_EOS 3
#define CARE_MASK 4
#define CARE_VIS 5

typedef struct care_ctx care_ctx;

/* library plumbing ------------------------------------------------------------------------- */
int care_version(void);
/* the 16-bit dtype code this build computes in (CARE_F16 or CARE_BF16) */
int care_h16_dtype(void);
const char* care_last_error(void);
int care_ctx_create(care_ctx** out, int device);
void care_ctx_destroy(care_ctx* ctx);
int care_ctx_sm_count(const care_ctx* ctx);
/* name of the kernel variant the most recent call of a family launched through this ctx: family = "gemm"
 * (care_gemm), "vocab" (care_vocab_beam_partials) or "self_attn" (care_self_attn_step); bench.py labels its
 * roofline records with it */
const char* care_ctx_last_kernel(const care_ctx* ctx, const char* family);
/* Arms ONE request: the next care_embed_ln call on this ctx also writes the live-slot records of the chunk-stream
 * self-attention (care_ctx_set_option "self_compact") for a prefix of n_pos positions of the beam state (anc [B, K, anc_stride],
 * tok_hist [B, anc_stride + 1, K], done [B] or NULL) - extra CTAs of the same launch - and the care_self_attn_step that follows
 * launches no record kernel.  Ignored (the record kernel runs as usual) when that self-attention would not use the records. */
int care_ctx_request_records(care_ctx* ctx, const uint8_t* anc, int anc_stride, const int32_t* tok_hist, const int32_t* done,
                             int B, int K, int H, int n_pos);
/* Inputs of the NEXT decode step that the beam kernel can produce while it still holds the chosen tokens: the
 * decoder input rows x0[v*K + b] = LN(word[tok] + pos[step] + gsg[v]) (Embeddings.py:134-188; what care_embed_ln
 * computes in a launch of its own).  care_ctx_set_next_step arms ONE request: the next care_beam_step_partials /
 * care_beam_first_step_partials call on this ctx also writes x0 (T16 [B*K, d], plus the fp32 copy x0_f32 when not
 * NULL) for every video that goes on, and the caller skips care_embed_ln at the start of that step.  NULL disarms.
 * Independently of this, the same kernel writes the live-slot records of the chunk-stream self-attention
 * (care_ctx_set_option "self_compact") for the next step, so that step launches no record kernel. */
typedef struct care_next_step {
  const float* word_emb;
  const float* pos_emb;
  const float* gsg;      /* fp32 [B, d] or NULL */
  const float* gamma;
  const float* beta;
  float eps;
  int32_t d;
  void* x0;
  float* x0_f32;         /* or NULL */
} care_next_step;
int care_ctx_set_next_step(care_ctx* ctx, const care_next_step* next);
/* Makes `ctx` use `other`'s table of per-shape GEMM variant picks (and its gemm_2sm / gemm_bn settings): two
 * contexts of one device that decode slices of the same batch on different streams then launch the same
 * kernel variant for the same shape, so a video's result does not depend on the slice it fell into. */
int care_ctx_share_tuning(care_ctx* ctx, care_ctx* other);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches claim) */
int64_t care_ctx_launch_count(const care_ctx* ctx);
/* Device-side early exit for a decode loop with no host polling: while `counter` is non-NULL every
 * kernel without a per-video predicate (GEMMs, embedding / residual LayerNorm) launched through this
 * ctx returns at once when *counter >= target.  The beam kernels keep care_beam_state.n_done, so
 * (st.n_done, B) makes the steps after the last video finished cost only their launches
 * (replaces the host-side `if not active: break`, Translator.py:77).  Pass NULL to disable. */
int care_ctx_set_early_exit(care_ctx* ctx, const int32_t* counter, int target);
/* implementation switches for A/B tests.  "attn_impl": 1 = TMA + tensor-core attention for T16
 * (default), 0 = the SIMT attention kernel for every dtype.  "gemm_2sm": 0 = single-CTA GEMM tiles
 * only, 1 = CTA-pair (tcgen05 cta_group::2) tiles whenever the shape allows, 2 (default) = choose per
 * (M, N, K, out dtype) by timing the variants ONCE, on the first care_gemm call with that shape - the
 * only place the library waits on the stream (never while the stream is being captured); 4 / 5 = clusters of
 * 4 / 2 CTA pairs over neighbouring n-tiles of one 256-row block, the A tile multicast inside the cluster
 * (also candidates of mode 2). */
int care_ctx_set_option(care_ctx* ctx, const char* name, int value);
/* "pdl": 1 (default) = the kernels of a decode step are launched with programmatic stream serialization: kernel
 * N+1 is scheduled while kernel N drains, runs its prologue and blocks in griddepcontrol.wait until N has completed.
 * "gemm_ln_multicast": 1 = care_gemm_add_ln fetches the A tile once per cluster and multicasts it (0, the default:
 * every CTA loads it; measured neutral - the mainloop is bound by shared-memory bandwidth, not by L2).
 * "gemm_ln_pair": care_gemm_add_ln on CTA pairs (tcgen05 cta_group::2; clusters of 2 * N/256 CTAs over 256-row blocks):
 * 0 = single-CTA clusters only, 1 = pairs whenever such a cluster fits the device, 2 (default) = choose per (M, N, K) by
 * timing both once, as "gemm_2sm" does.
 * "vocab_split": epilogue schedule of the fused vocabulary kernel (care_vocab_beam_partials): 0 = its two epilogue groups
 * take alternate tiles, 1 = both fold every tile (128 columns each: half the latency per tile), 2 (default) = column
 * halves when a CTA's run has fewer than "vocab_split_tiles" (default 24) tiles.  The record layout is the same; the
 * consumers (care_beam_step_partials, care_nar_best_partials) derive the schedule from the shape and these options, so
 * change them only between a producer / consumer pair.
 * "l2_hints": L2 eviction priorities on TMA loads; bit 0 = weight tiles evict_last (measured neutral), bit 1 (default on)
 * = cross-attention K/V tiles evict_first (the K/V stream no longer evicts what the following kernels re-read).
 * "fuse_info": 1 = the beam kernel also writes the next step's live-slot records (0, the default: a kernel of its
 * own before the self-attention; measured neutral to slightly slower when fused). */
/* "vocab_2sm": 1 (default) = the fused vocabulary kernel runs on CTA pairs when the shape has at least two
 * waves of 256 x 256 tiles, 0 = single-CTA tiles only.
 * "gemm_smallm": 1 (default) = GEMMs with M <= 16 rows (batch-1 / latency mode) use a weight-streaming
 * warp-MMA kernel with one CTA per 8 output columns, 0 = always the tcgen05 tile kernels.
 * "self_compact" (option above), T16 self-attention over the KV cache: 2 (default) = only the cache slots
 * some beam of the video still references are read, as a stream of 16-row chunks fetched by TMA row gathers
 * (prefixes of at least 6 positions, 1024 .. 16384 x H (video, head) pairs per call; otherwise the dense tile);
 * 3 = the chunk stream for every shape (tests); 1 = the earlier per-CTA gather of the live slots (slower,
 * kept for A/B runs); 0 = all K slots of every
 * position with one TMA tensor copy.  The environment variable CARE_B200_SELF_COMPACT sets the default.
 * Statistics kept on the device (reading one synchronises): "self_attn_rows" = K/V cache rows per
 * head read so far by the T16 self-attention kernels, summed over videos and steps. */
int care_ctx_counter(care_ctx* ctx, const char* name, int64_t* value);

/* C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) — every nn.Linear on the path:
 * Encoder.py:165-168 (Embedder Linear), Attention.py:53-67 (query/key/value),
 * SubLayers.py:33-38 (dense), SubLayers.py:126-135 (dense1/dense2), Head.py:26-32
 * (tgt_word_prj), pred_attribute.py:62-65,124 (prj), :257-260 (semantic2hidden).
 * A, W are `dtype` (fp32: SIMT FFMA kernel, bit-comparable accumulation; T16: tcgen05/TMEM
 * kernel fed by TMA, fp32 accumulate).  bias fp32 or NULL.  C is `out_dtype`.
 * Requirements: lda, ldw multiples of 8 elements (T16) / 4 (fp32), 16-byte aligned bases;
 * ldc a multiple of 8; columns [N, min(roundup8(N), ldc)) of C are written with zeros. */
int care_gemm(care_ctx* ctx, int dtype, const void* A, int64_t lda, const void* W, int64_t ldw,
              const float* bias, void* C, int64_t ldc, int out_dtype, int M, int N, int K, int act,
              void* stream);

/* fp32 -> 16-bit operand conversion of the incoming feature tensors (translate.py:36-38 hands fp32
 * features) and of other fp32 GEMM inputs.  src fp32 [rows, cols] (row stride ld_src); dst T16
 * [rows, terms * cols_pad].  terms = 1: plain cast (zero padded to cols_pad).  terms = 3: each row becomes
 * [hi | lo | hi * 2^-11] with hi = T16(x), lo = T16(x - hi); multiplied by a weight prepared as
 * [W_hi | W_hi | W_lo * 2^11] one K-concatenated care_gemm computes A_hi W_hi + A_lo W_hi + A_hi W_lo -
 * the fp32 product to ~2^-21 - on the tensor cores (the 2^11 keeps W_lo out of fp16's subnormals).  Used for everything upstream of the concept
 * ranking and the length prediction (Encoder.py:165-168, pred_attribute.py:124-125, pred_length.py:18-22),
 * where a 16-bit rounding would re-order the discrete top-k. */
int care_split_f32_h16(care_ctx* ctx, const float* src, int64_t ld_src, int64_t rows, int cols,
                       int cols_pad, int terms, void* dst, void* stream);

/* Encoder stream tail (Encoder.py:167: LayerNorm after Linear; Encoder.py:106: mean over time).
 * x: fp32 [B*T, d] (the Linear output incl. bias).  Writes LN(x) as T into
 * out[(v*out_rows + out_row0 + t)*d ...] when out != NULL (requires out_row0 + T <= out_rows), and
 * the per-video temporal mean of LN(x) as fp32 into mean_out[v*mean_ld + mean_col0 ...] when
 * mean_out != NULL. */
int care_encoder_ln_mean(care_ctx* ctx, int dtype, const float* x, const float* gamma,
                         const float* beta, float eps, int B, int T, int d, void* out,
                         int out_rows, int out_row0, float* mean_out, int64_t mean_ld, int mean_col0,
                         void* stream);

/* EncoderWithHighWayBN tail (Encoder.py:184-187,210-241): out = BN_eval(g*h + (1-g)*tanh(y)),
 * g = sigmoid(gpre); h, ypre, gpre fp32 [B*T, d] (Linear outputs incl. bias).  Same outputs as
 * care_encoder_ln_mean.  bn_scale/bn_shift are the folded eval-mode affine (fp32 [d]). */
int care_encoder_highway_bn_mean(care_ctx* ctx, int dtype, const float* h, const float* ypre,
                                 const float* gpre, const float* bn_mean, const float* bn_var,
                                 const float* bn_w, const float* bn_b, float bn_eps, int B, int T,
                                 int d, void* out, int out_rows, int out_row0, float* mean_out,
                                 int64_t mean_ld, int mean_col0, void* stream);

/* Concept head (pred_attribute.py:17-46 noisy-or, :262-289 SemanticContainer, Embeddings.py:53-87
 * NaiveEmbeddings, Framework.py:184-185 concat).  scores fp32 [B, ld_scores] = prj output incl. bias.
 * preds_f32 [B, n_attr]; preds_T fp32 [B, ld_preds_T] (A operand of semantic2hidden, zero padded; fp32 in
 * every mode: the small ranking heads run on the FFMA GEMM);
 * labels int64 [B, topk] sorted by (prob desc, index asc); LN(word[label]+pos[rank]) written as T
 * into memory rows mem_row0 .. mem_row0+topk-1 of each video ([B, mem_rows, d]). */
int care_concept_head(care_ctx* ctx, int dtype, const float* scores, int64_t ld_scores, int B,
                      int n_attr, int topk, const float* attr_word, const float* attr_pos,
                      const float* gamma, const float* beta, float eps, int d, float* preds_f32,
                      void* preds_T, int64_t ld_preds_T, int64_t* labels, void* memory, int mem_rows,
                      int mem_row0, void* stream);

/* Decoder input embedding for ONE position per row (Embeddings.py:134-188):
 * out[r] = LN(((word[tok[r]] + pos[position]) + add_feats[r / rows_per_video]) + gsg[r / rows_per_video]).
 * tokens int32 [R]; add_feats / gsg fp32 [n_videos, d] or NULL; out T [R, d].
 * With positions != NULL (int32 [R]) each row uses its own position (mask-predict passes).
 * out32 (fp32 [R, d] or NULL): an fp32 copy of the result, the head of an fp32 residual stream
 * (care_gemm_add_ln with residual_dtype = 0). */
int care_embed_ln(care_ctx* ctx, int dtype, const int32_t* tokens, const int32_t* positions,
                  int position, const float* word_emb, const float* pos_emb, const float* add_feats,
                  const float* gsg, int rows_per_video, const float* gamma, const float* beta,
                  float eps, int R, int d, void* out, float* out32, void* stream);

/* Post-LN residual block tail (SubLayers.py:74-79, :148-150): out = LN(x + residual).
 * x fp32 [R, d] (Linear output incl. bias), residual T [R, d], out T [R, d]. */
int care_add_ln(care_ctx* ctx, int dtype, const float* x, const void* residual, const float* gamma,
                const float* beta, float eps, int R, int d, void* out, void* stream);

/* Output projection / second FFN Linear fused with the residual LayerNorm that follows it
 * (SubLayers.py:68-79: dense -> dropout(identity) -> + residual -> LayerNorm; SubLayers.py:137-152 likewise for
 * the FFN), T16 operands only:  out = LayerNorm(A[M,K] W[N,K]^T + bias + residual) * gamma + beta.
 * N must be 512, 768 or 1024: a thread-block cluster of N/256 CTAs shares one 128-row block and exchanges the
 * per-row (sum, sum of squares) through distributed shared memory, so the fp32 pre-LayerNorm tensor that
 * care_gemm + care_add_ln pass through HBM never exists.  residual: T16 [M, N] (residual_dtype = the T16 code)
 * or fp32 [M, N] (residual_dtype = 0: the residual stream stays fp32; out32 [M, N] then receives the fp32
 * result next to the T16 copy in out16 that the next GEMM consumes).  out16: T16 [M, N], row stride N. */
int care_gemm_add_ln(care_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                     const void* residual, int residual_dtype, const float* gamma, const float* beta, float eps,
                     void* out16, float* out32, int M, int N, int K, void* stream);

/* Fused decode-step self-attention over the KV cache (Attention.py:81-129 for the newest query
 * position only; masks Transformer.py:15-47,169-174).  qkv_step: T [R, 3d] of THIS step (q|k|v);
 * cache: T [T_max, R, 3d] holding every step's qkv (the step-t slice is qkv_step itself), so the
 * cache is never reordered: anc uint8 [B, K, T_max] maps (beam, position) -> cache slot, and
 * tok_hist int32 [B, T_max+1, K] holds the token fed at (position, slot) for the PAD-key mask.
 * n_pos = number of positions to attend (= step t).  ctx_out T [R, d].  done int32 [B] or NULL. */
int care_self_attn_step(care_ctx* ctx, int dtype, const void* cache, int n_pos, int B, int K, int H,
                        int d, const uint8_t* anc, int anc_stride, const int32_t* tok_hist,
                        const int32_t* done, void* ctx_out, void* stream);

/* Fused decode-step cross-attention (Attention.py:81-129 with hybrid_bias :109-111).  q T [R, ldq];
 * kv T [B, Lm, 2d] (k|v projected ONCE per video, never replicated per beam — replaces
 * auto_enlarge, misc/utils.py:244-279, and the per-step re-projection Attention.py:63-67);
 * hybrid_bias fp32 [H, Lm] or NULL.  ctx_out T [R, d]. */
int care_cross_attn_step(care_ctx* ctx, int dtype, const void* q, int64_t ldq, const void* kv, int Lm,
                         int B, int K, int H, int d, const float* hybrid_bias, const int32_t* done,
                         void* ctx_out, void* stream);

/* Full-sequence attention over groups: every query row of group g attends the key rows
 * [g*nk, (g+1)*nk) of the key/value matrix `kv` (row stride ldkv elements; K at column k_col + 64*head,
 * V at v_col + 64*head).  Query / output rows of group g are [g*nq, (g+1)*nq).  Masks as in
 * Transformer.py:15-47,169-174 + Attention.py:104-111: key_tokens (int32 [n_groups*nk], may be NULL)
 * masks <pad> keys, causal masks keys k > query index (needs nq == nk), bias fp32 [H, nk] is added after
 * masking.  Used by the mask-predict passes (Translator.py:240-305, na_algorithms.py:67-82: self
 * attention with group = one candidate sequence; cross attention with group = one video and
 * nq = candidates*L rows sharing the video's memory) and by the stateless decoding_phase
 * (Framework.py:240-269).  q, kv, out are `dtype`. */
int care_group_attn(care_ctx* ctx, int dtype, const void* q, int64_t ldq, const void* kv, int64_t ldkv,
                    int k_col, int v_col, int n_groups, int nq, int nk, int H, int d,
                    const int32_t* key_tokens, int causal, const float* bias, void* out, void* stream);

/* out fp32 [B, d] = mean over the `rows` rows of x [B, rows, d] (Transformer.py:182-189,
 * enhance_input == 2: mean of the decoder memory added to every input embedding of a NAR pass). */
int care_rows_mean(care_ctx* ctx, int dtype, const void* x, int B, int rows, int d, float* out, void* stream);
/* out T [B, ld_out] (first d columns) = sum_s weights[s] * means[b, s*d + :]; weights is a HOST array of
 * n <= 8 floats (pred_length.py:14-17: mean over all predictor tokens, from the per-stream means). */
int care_combine_means(care_ctx* ctx, int dtype, const void* means, int B, int n, int d, const float* weights,
                       void* out, int64_t ld_out, void* stream);

/* Mask-predict bookkeeping, all device side (replaces the host loops of models/Translator.py:240-318
 * and misc/Decoding/na_algorithms.py:60-82,128-197).  R = B * n_cand candidate rows, L tokens each. */
/* predict_length_beam (Translator.py:307-311): top-n_cand classes by (logit desc, index asc), + bias, clamp */
int care_nar_length_beam(care_ctx* ctx, const float* logits, int64_t ld, int B, int n_classes, int n_cand,
                         int length_bias, int min_len, int max_len, int32_t* lengths, void* stream);
/* canvas (Translator.py:275-280; na_algorithms.py:60-65): tokens[r,p] = first_token if p < lengths[r] else
 * <pad>; positions[r*L+p] = p; probs = 0 */
int care_nar_init(care_ctx* ctx, const int32_t* lengths, int R, int L, int first_token, int32_t* tokens,
                  int32_t* positions, float* probs, void* stream);
/* generate_step_with_prob (na_algorithms.py:6-14): per token row arg max and its softmax probability,
 * from fp32 logits, or from the records of care_vocab_beam_partials (K = 1) */
int care_nar_best_logits(care_ctx* ctx, const float* logits, int64_t ldv, int rows, int V, int32_t* idx,
                         float* prob, void* stream);
int care_nar_best_partials(care_ctx* ctx, const float* partials, int nseg, int rows, int V, int32_t* idx,
                           float* prob, void* stream);
/* teacher rescoring (scoring_by_teacher, na_algorithms.py:92-126): logits fp32 [R*L, ldv] of the auto-regressive
 * teacher's teacher-forced pass over [<bos>, y_0 .. y_{L-2}]; targets int32 [R*L] = the student's tokens in the
 * teacher's vocabulary.  out[r,p] = softmax(logits[r,p])[targets[r,p]] for p < lengths[r], 1 for pad positions,
 * multiplied by probs_in[r,p] when probs_in != NULL (token_probs * corresponding_probs, :180,194). */
int care_nar_teacher_probs(care_ctx* ctx, const float* logits, int64_t ldv, const int32_t* targets,
                           const int32_t* lengths, int R, int L, int V, const float* probs_in, float* out,
                           void* stream);
/* write-back (na_algorithms.py:67-82,185-190): where mask_ind (NULL = everywhere) tokens/probs := new, pad
 * positions forced to (<pad>, 1); zero_mask_token: probs := 0 where the new token is <mask> (:64) */
int care_nar_apply(care_ctx* ctx, int32_t* tokens, float* probs, const int32_t* new_idx, const float* new_prob,
                   const uint8_t* mask_ind, const int32_t* lengths, int R, int L, int zero_mask_token,
                   void* stream);
/* what to re-predict (na_algorithms.py:128-137,172-182).  mode 0: positions holding <mask>; mode 1: the
 * max(1, num_mask_by_len[lengths[r]]) lowest-probability positions (ties: lower position), which are set
 * to <mask>.  mask_ind uint8 [R, L] out. */
int care_nar_remask(care_ctx* ctx, int32_t* tokens, const float* probs, const int32_t* lengths,
                    const int32_t* num_mask_by_len, int mode, int R, int L, uint8_t* mask_ind, void* stream);
/* final choice (Translator.py:292-303): per video argmax over candidates of sum_p log(prob) / len^alpha
 * (fp32); out_tokens int32 [B, L], out_lprobs fp32 [B, L], out_best int32 [B] (may be NULL) */
int care_nar_select(care_ctx* ctx, const int32_t* tokens, const float* probs, const int32_t* lengths, int B,
                    int n_cand, int L, float alpha, int32_t* out_tokens, float* out_lprobs, int32_t* out_best,
                    void* stream);

/* Beam state, all device resident (replaces misc/Decoding/Beam.py's per-video Python objects). */
typedef struct care_beam_state {
  int32_t B, K, T_max, V, need;   /* need = max(K, topk)  (Beam.py:10) */
  float* scores;                  /* [B, K]              Beam.scores */
  int32_t* cur_tok;               /* [B*K]               token fed at the next step */
  int32_t* tok_hist;              /* [B, T_max+1, K]     Beam.next_ys (position 0 = BOS) */
  int32_t* prev_ks;               /* [B, T_max, K]       Beam.prev_ks */
  uint8_t* anc;                   /* [B, K, T_max]       ancestry slot table for the KV cache */
  float* fin_score;               /* [B, need]           Beam.finished */
  int32_t* fin_t;                 /* [B, need] */
  int32_t* fin_k;                 /* [B, need] */
  int32_t* fin_count;             /* [B] */
  int32_t* done;                  /* [B] */
  int32_t* n_done;                /* [1] number of finished videos */
  float* scratch;                 /* [B*K*20] 32-bit words: per-row (max, sum-exp, top candidates) */
} care_beam_state;

/* reset state for a new batch: scores 0, cur_tok/tok_hist[0] = bos, counters 0 */
int care_beam_init(care_ctx* ctx, const care_beam_state* st, int bos, void* stream);

/* One beam-search step for every unfinished video (Translator.py:111-143 + Beam.py:45-85):
 * log_softmax over V, + running score, rows whose last token is <eos> := -1e20, top-K over the
 * flattened K*V candidates ordered by (value desc, flat index asc), back-pointers / tokens /
 * scores / ancestry update, finished list and the finish rule.  logits fp32 [R, ldv].
 * step = len_input_ids (1-based); at step 1 only beam row 0 is scored (Beam.py:56).
 * Optional audit outputs: cand_val fp32 [B, K+1], cand_idx int32 [B, K+1] (the K winners and the
 * runner-up), may be NULL. */
int care_beam_step(care_ctx* ctx, const care_beam_state* st, const float* logits, int64_t ldv,
                   int step, int max_len, float* cand_val, int32_t* cand_idx, void* stream);

/* Model ensembling (Translator.py:111-133, predict_word): out[r, :] = mean over the n models of
 * log_softmax(logits_i[r, :]).  `logits` is a HOST array of n <= 8 device pointers, each fp32 [R, ldv];
 * out fp32 [R, ldv].  care_beam_step_logprobs is care_beam_step for rows that already hold log-probabilities
 * (no second normalisation). */
int care_ensemble_logprobs(care_ctx* ctx, const float* const* logits, int n, int64_t ldv, int R, int V,
                           float* out, void* stream);
int care_beam_step_logprobs(care_ctx* ctx, const care_beam_state* st, const float* logprobs, int64_t ldv,
                            int step, int max_len, float* cand_val, int32_t* cand_idx, void* stream);

/* Fused vocabulary projection + beam partials, T16 only (Head.py:26-32 + Translator.py:127 + the
 * top-k half of Beam.py:45-60): the fp32 logits never reach HBM.  x T16 [R, ldx] (decoder output of
 * the newest position), W T16 [V, ldw] (cls_head.tgt_word_prj.weight).  Each row's vocabulary is
 * reduced on the tensor-core kernel's epilogue to per-segment records (max, sum-exp, top-(K+1) raw
 * logits + column ids); partials is fp32 words [R, nseg, 2 + 2*KB] with nseg =
 * care_vocab_beam_nseg(ctx, R, V) and KB = 2/4/6/9 for K <= 1/3/5/8.  care_beam_step_partials then
 * does what care_beam_step does after its own row pass. */
int care_vocab_beam_nseg(care_ctx* ctx, int R, int V);
int care_vocab_beam_partials(care_ctx* ctx, const void* x, int64_t ldx, const void* W, int64_t ldw, int R,
                             int V, int d, int K, float* partials, int nseg, void* stream);
int care_beam_step_partials(care_ctx* ctx, const care_beam_state* st, const float* partials, int nseg,
                            int step, int max_len, float* cand_val, int32_t* cand_idx, void* stream);
/* The same for step 1 computed on ONE row per video: before the first step all K beams of a video hold <bos> and
 * identical state, Beam.advance looks at beam 0 only (Beam.py:56), and every new beam's ancestor at position 0 is
 * slot 0 - so the decoder layer and the vocabulary kernel need B rows, not B*K.  partials: the records of
 * care_vocab_beam_partials run on those B rows (record row v = video v, nseg = care_vocab_beam_nseg(ctx, B, V)). */
int care_beam_first_step_partials(care_ctx* ctx, const care_beam_state* st, const float* partials,
                                  int nseg, int max_len, float* cand_val, int32_t* cand_idx,
                                  void* stream);

/* Hypothesis extraction (Translator.py:211-220, Beam.py:91-105,119-132): rank finished items by
 * score / t^alpha (double), stable, and back-walk the n_best first.  out_tokens int32
 * [B, n_best, T_max] PAD filled; out_len int32 [B, n_best] (0 = no such hypothesis);
 * out_score fp32 [B, n_best] raw cumulative log-prob; out_t int32 [B, n_best]. */
int care_beam_finalize(care_ctx* ctx, const care_beam_state* st, double alpha, int n_best,
                       int32_t* out_tokens, int32_t* out_len, float* out_score, int32_t* out_t,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CARE_B200_H_ */
