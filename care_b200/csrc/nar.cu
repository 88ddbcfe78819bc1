// Small device-side pieces of the non-autoregressive (mask-predict) path, replacing the host/Python
// logic of Translator_NARFormer (models/Translator.py:240-318) and MaskPredict
// (misc/Decoding/na_algorithms.py:60-82,128-197), plus two reductions used by its inputs.
// All of it is O(rows * L) integer/float bookkeeping with L <= 64: one thread per row.
#include <cfloat>
#include <climits>

#include "common.cuh"

namespace care {
namespace nar {

constexpr int MAX_L = 64;

// out[b, :] = mean over `rows` rows of x[b]   (Transformer.py:182-189, enhance_input == 2)
template <typename T>
__global__ void rows_mean_kernel(const T* __restrict__ x, int rows, int d, float* __restrict__ out) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    for (int r = 0; r < rows; ++r) s += Act<T>::to_float(x[((int64_t)b * rows + r) * d + c]);
    out[(int64_t)b * d + c] = s / (float)rows;
  }
}

struct Weights {
  float w[8];
};
// out[b, c] = sum_s w[s] * means[b, s*d + c]   (mean over all predictor tokens from per-stream means)
template <typename T>
__global__ void combine_means_kernel(const T* __restrict__ means, int n, int d, Weights w, T* __restrict__ out,
                                     int64_t ld_out) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += w.w[i] * Act<T>::to_float(means[(int64_t)b * n * d + (int64_t)i * d + c]);
    out[(int64_t)b * ld_out + c] = Act<T>::from_float(s);
  }
}

// top-n_cand classes by (logit desc, index asc) (+ bias, clamp)   (Translator.py:307-311)
__global__ void length_beam_kernel(const float* __restrict__ logits, int64_t ld, int B, int n_classes, int n_cand,
                                   int length_bias, int min_len, int max_len, int32_t* __restrict__ lengths) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* row = logits + (int64_t)b * ld;
  uint64_t used = 0ull;
  for (int r = 0; r < n_cand; ++r) {
    int best = -1;
    float bv = 0.f;
    for (int c = 0; c < n_classes; ++c) {
      if ((used >> c) & 1ull) continue;
      const float x = row[c];
      if (best < 0 || x > bv) {
        best = c;
        bv = x;
      }
    }
    used |= 1ull << best;
    int len = best + length_bias;
    len = len < min_len ? min_len : len;
    len = len > max_len ? max_len : len;
    lengths[(int64_t)b * n_cand + r] = len;
  }
}

__global__ void init_kernel(const int32_t* __restrict__ lengths, int R, int L, int first_token,
                            int32_t* __restrict__ tokens, int32_t* __restrict__ positions, float* __restrict__ probs) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * L) return;
  const int r = (int)(i / L), p = (int)(i - (int64_t)r * L);
  tokens[i] = p < lengths[r] ? first_token : CARE_PAD;
  positions[i] = p;
  probs[i] = 0.f;
}

// one warp per token row of fp32 logits: arg max (lowest index on ties) and softmax probability of it
__global__ void __launch_bounds__(256) best_logits_kernel(const float* __restrict__ logits, int64_t ldv, int rows, int V,
                                                          int32_t* __restrict__ idx, float* __restrict__ prob) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = logits + (int64_t)row * ldv;
  float m = -INFINITY;
  int mi = INT_MAX;
  for (int c = lane; c < V; c += 32) {
    const float v = x[c];
    if (v > m) {
      m = v;
      mi = c;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (om > m || (om == m && oi < mi)) {
      m = om;
      mi = oi;
    }
  }
  float s = 0.f;
  for (int c = lane; c < V; c += 32) s += expf(x[c] - m);
  s = warp_sum(s);
  if (lane == 0) {
    idx[row] = mi;
    prob[row] = 1.0f / s;   // softmax value of the maximum: exp(0) / sum
  }
}

// one warp per token row of the TEACHER's fp32 logits: softmax probability of a given token (scoring_by_teacher,
// na_algorithms.py:92-126), times the student's own probability when `probs_in` is given; pad positions count 1
__global__ void __launch_bounds__(256) teacher_probs_kernel(const float* __restrict__ logits, int64_t ldv,
                                                            const int32_t* __restrict__ targets,
                                                            const int32_t* __restrict__ lengths, int R, int L, int V,
                                                            const float* __restrict__ probs_in, float* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (int64_t)R * L) return;
  const int r = (int)(row / L), p = (int)(row - (int64_t)r * L);
  float pr = 1.0f;
  if (p < lengths[r]) {
    const float* x = logits + row * ldv;
    float m = -INFINITY;
    for (int c = lane; c < V; c += 32) m = fmaxf(m, x[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(x[c] - m);
    s = warp_sum(s);
    pr = expf(x[targets[row]] - m) / s;
  }
  if (lane == 0) out[row] = probs_in != nullptr ? probs_in[row] * pr : pr;
}

// same from the fused vocabulary kernel's records (KB = 2); segment existence as in beam.cu
__global__ void best_partials_kernel(const float* __restrict__ partials, int nseg, int n_tiles, int64_t T, int64_t G,
                                     int row_shift, int split, int rows, int32_t* __restrict__ idx,
                                     float* __restrict__ prob) {
  constexpr int KB = 2, W = 2 + 2 * KB;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int m_blk = r >> row_shift;
  const int c0 = (int)((((int64_t)m_blk * n_tiles + 1) * G - 1) / T);
  const int c1 = (int)((((int64_t)m_blk * n_tiles + n_tiles) * G - 1) / T);
  const int64_t mlo = (int64_t)m_blk * n_tiles, mhi = mlo + n_tiles;
  const float* base = partials + (int64_t)r * nseg * W;
  float M = -INFINITY;
  int mi = INT_MAX;
  for (int pass = 0; pass < 2; ++pass) {
    float S = 0.f;
    for (int c = c0; c <= c1; ++c) {
      const int64_t start = (int64_t)c * T / G, end = (int64_t)(c + 1) * T / G;
      const int64_t lo = start > mlo ? start : mlo, hi = end < mhi ? end : mhi;
      for (int g = 0; g < 2; ++g) {
        if (!(lo + (split ? 0 : ((g - (lo - start)) & 1)) < hi)) continue;
        const float* rec = base + (2 * (c - c0) + g) * W;
        if (pass == 0) {
          const float v = rec[2];
          const int i = reinterpret_cast<const int*>(rec)[2 + KB];
          if (v > M || (v == M && i < mi)) {
            M = v;
            mi = i;
          }
        } else {
          S += rec[1] * __expf(rec[0] - M);
        }
      }
    }
    if (pass == 1) {
      idx[r] = mi;
      prob[r] = 1.0f / S;
    }
  }
}

__global__ void apply_kernel(int32_t* __restrict__ tokens, float* __restrict__ probs, const int32_t* __restrict__ new_idx,
                             const float* __restrict__ new_prob, const uint8_t* __restrict__ mask_ind,
                             const int32_t* __restrict__ lengths, int R, int L, int zero_mask_token) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)R * L) return;
  const int r = (int)(i / L), p = (int)(i - (int64_t)r * L);
  if (mask_ind != nullptr && !mask_ind[i]) return;
  int t = new_idx[i];
  float pr = new_prob[i];
  if (p >= lengths[r]) {   // na_algorithms.py:78-80: pad positions are forced
    t = CARE_PAD;
    pr = 1.0f;
  }
  if (zero_mask_token && t == CARE_MASK) pr = 0.f;   // na_algorithms.py:64
  tokens[i] = t;
  probs[i] = pr;
}

// mode 0: re-predict positions whose token is <mask> (first refinement after coarse templates);
// mode 1: re-predict the max(1, num_mask[len]) lowest-probability positions (select_worst)
__global__ void remask_kernel(int32_t* __restrict__ tokens, const float* __restrict__ probs,
                              const int32_t* __restrict__ lengths, const int32_t* __restrict__ num_mask_by_len, int mode,
                              int R, int L, uint8_t* __restrict__ mask_ind) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int32_t* tk = tokens + (int64_t)r * L;
  const float* pr = probs + (int64_t)r * L;
  uint8_t* mk = mask_ind + (int64_t)r * L;
  if (mode == 0) {
    for (int p = 0; p < L; ++p) mk[p] = tk[p] == CARE_MASK ? 1 : 0;
    return;
  }
  int n = num_mask_by_len[lengths[r]];
  n = n < 1 ? 1 : n;
  n = n > L ? L : n;
  uint64_t used = 0ull;
  for (int it = 0; it < n; ++it) {
    int best = -1;
    float bv = 0.f;
    for (int p = 0; p < L; ++p) {
      if ((used >> p) & 1ull) continue;
      if (best < 0 || pr[p] < bv) {
        best = p;
        bv = pr[p];
      }
    }
    used |= 1ull << best;
  }
  for (int p = 0; p < L; ++p) {
    const bool on = (used >> p) & 1ull;
    mk[p] = on ? 1 : 0;
    if (on) tk[p] = CARE_MASK;
  }
}

__global__ void select_kernel(const int32_t* __restrict__ tokens, const float* __restrict__ probs,
                              const int32_t* __restrict__ lengths, int B, int n_cand, int L, float alpha,
                              int32_t* __restrict__ out_tokens, float* __restrict__ out_lprobs,
                              int32_t* __restrict__ out_best) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int best = 0;
  float bv = -INFINITY;
  for (int c = 0; c < n_cand; ++c) {
    const int r = b * n_cand + c;
    float s = 0.f;
    for (int p = 0; p < L; ++p) s += logf(probs[(int64_t)r * L + p]);
    const float avg = s / powf((float)lengths[r], alpha);
    if (avg > bv) {
      bv = avg;
      best = c;
    }
  }
  const int r = b * n_cand + best;
  for (int p = 0; p < L; ++p) {
    out_tokens[(int64_t)b * L + p] = tokens[(int64_t)r * L + p];
    out_lprobs[(int64_t)b * L + p] = logf(probs[(int64_t)r * L + p]);
  }
  if (out_best) out_best[b] = best;
}

}  // namespace nar

namespace vb {  // vocab_beam.cu
void seg_layout(const care_ctx* ctx, int R, int V, int* n_tiles, int64_t* T, int64_t* G, int* row_shift, int* split);
}
}  // namespace care

using namespace care;

extern "C" {

int care_rows_mean(care_ctx* ctx, int dtype, const void* x, int B, int rows, int d, float* out, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_rows_mean");
  CARE_CHECK_ARG(ctx && x && out && B > 0 && rows > 0 && d > 0, "care_rows_mean: bad args");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32) nar::rows_mean_kernel<float><<<B, 256, 0, s>>>((const float*)x, rows, d, out);
  else nar::rows_mean_kernel<h16><<<B, 256, 0, s>>>((const h16*)x, rows, d, out);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_combine_means(care_ctx* ctx, int dtype, const void* means, int B, int n, int d, const float* weights,
                       void* out, int64_t ld_out, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_combine_means");
  CARE_CHECK_ARG(ctx && means && out && weights && B > 0 && n >= 1 && n <= 8 && d > 0, "care_combine_means: bad args");
  nar::Weights w{};
  for (int i = 0; i < n; ++i) w.w[i] = weights[i];
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32)
    nar::combine_means_kernel<float><<<B, 256, 0, s>>>((const float*)means, n, d, w, (float*)out, ld_out);
  else
    nar::combine_means_kernel<h16><<<B, 256, 0, s>>>((const h16*)means, n, d, w,
                                                               (h16*)out, ld_out);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_nar_length_beam(care_ctx* ctx, const float* logits, int64_t ld, int B, int n_classes, int n_cand,
                         int length_bias, int min_len, int max_len, int32_t* lengths, void* stream) {
  CARE_CHECK_ARG(ctx && logits && lengths && B > 0 && n_classes >= 1 && n_classes <= 64 && n_cand >= 1 &&
                     n_cand <= n_classes,
                 "care_nar_length_beam: bad args (n_classes=%d n_cand=%d)", n_classes, n_cand);
  nar::length_beam_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, ld, B, n_classes, n_cand,
                                                                            length_bias, min_len, max_len, lengths);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_nar_init(care_ctx* ctx, const int32_t* lengths, int R, int L, int first_token, int32_t* tokens,
                  int32_t* positions, float* probs, void* stream) {
  CARE_CHECK_ARG(ctx && lengths && tokens && positions && probs && R > 0 && L >= 1 && L <= nar::MAX_L,
                 "care_nar_init: bad args (L=%d)", L);
  const int64_t n = (int64_t)R * L;
  nar::init_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(lengths, R, L, first_token, tokens,
                                                                             positions, probs);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_nar_best_logits(care_ctx* ctx, const float* logits, int64_t ldv, int rows, int V, int32_t* idx, float* prob,
                         void* stream) {
  CARE_CHECK_ARG(ctx && logits && idx && prob && rows > 0 && V > 0, "care_nar_best_logits: bad args");
  nar::best_logits_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(logits, ldv, rows, V, idx, prob);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_nar_teacher_probs(care_ctx* ctx, const float* logits, int64_t ldv, const int32_t* targets,
                           const int32_t* lengths, int R, int L, int V, const float* probs_in, float* out, void* stream) {
  CARE_CHECK_ARG(ctx && logits && targets && lengths && out && R > 0 && L > 0 && V > 0, "care_nar_teacher_probs: bad args");
  const int64_t rows = (int64_t)R * L;
  nar::teacher_probs_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(logits, ldv, targets, lengths, R,
                                                                                         L, V, probs_in, out);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_nar_best_partials(care_ctx* ctx, const float* partials, int nseg, int rows, int V, int32_t* idx,
                           float* prob, void* stream) {
  CARE_CHECK_ARG(ctx && partials && idx && prob && rows > 0 && V > 0, "care_nar_best_partials: bad args");
  int n_tiles, row_shift, split;
  int64_t T, G;
  vb::seg_layout(ctx, rows, V, &n_tiles, &T, &G, &row_shift, &split);
  nar::best_partials_kernel<<<(rows + 127) / 128, 128, 0, (cudaStream_t)stream>>>(partials, nseg, n_tiles, T, G,
                                                                                  row_shift, split, rows, idx, prob);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_nar_apply(care_ctx* ctx, int32_t* tokens, float* probs, const int32_t* new_idx, const float* new_prob,
                   const uint8_t* mask_ind, const int32_t* lengths, int R, int L, int zero_mask_token, void* stream) {
  CARE_CHECK_ARG(ctx && tokens && probs && new_idx && new_prob && lengths && R > 0 && L >= 1, "care_nar_apply: bad args");
  const int64_t n = (int64_t)R * L;
  nar::apply_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(tokens, probs, new_idx, new_prob, mask_ind,
                                                                              lengths, R, L, zero_mask_token);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_nar_remask(care_ctx* ctx, int32_t* tokens, const float* probs, const int32_t* lengths,
                    const int32_t* num_mask_by_len, int mode, int R, int L, uint8_t* mask_ind, void* stream) {
  CARE_CHECK_ARG(ctx && tokens && probs && lengths && mask_ind && R > 0 && L >= 1 && L <= nar::MAX_L,
                 "care_nar_remask: bad args (L=%d)", L);
  CARE_CHECK_ARG(mode == 0 || num_mask_by_len != nullptr, "care_nar_remask: mode 1 needs num_mask_by_len");
  nar::remask_kernel<<<(R + 127) / 128, 128, 0, (cudaStream_t)stream>>>(tokens, probs, lengths, num_mask_by_len, mode, R,
                                                                        L, mask_ind);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_nar_select(care_ctx* ctx, const int32_t* tokens, const float* probs, const int32_t* lengths, int B,
                    int n_cand, int L, float alpha, int32_t* out_tokens, float* out_lprobs, int32_t* out_best,
                    void* stream) {
  CARE_CHECK_ARG(ctx && tokens && probs && lengths && out_tokens && out_lprobs && B > 0 && n_cand >= 1 && L >= 1,
                 "care_nar_select: bad args");
  nar::select_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(tokens, probs, lengths, B, n_cand, L, alpha,
                                                                        out_tokens, out_lprobs, out_best);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // extern "C"
