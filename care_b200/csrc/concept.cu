// Fused concept head: noisy-or sigmoid (pred_attribute.py:17-46) -> top-k concepts sorted by
// (probability desc, index asc) (pred_attribute.py:264) -> concept embedding gather + position add +
// LayerNorm (Embeddings.py:53-87), written straight into rows [mem_row0, mem_row0+topk) of the
// decoder memory (Framework.py:184-185 "concat").  One CTA per video; the order comes from a bitonic sort of
// the (probability, index) pairs in shared memory.
#include <climits>

#include "common.cuh"

namespace care {
namespace concept_head {

constexpr int THREADS = 512;
constexpr int MAX_ATTR = 1024;
constexpr int MAX_TOPK = 64;
constexpr int MAX_CHUNKS = 8;  // d <= 1024

template <typename T>
__global__ void __launch_bounds__(THREADS)
concept_head_kernel(const float* __restrict__ scores, int64_t ld_scores, int n_attr, int topk,
                    const float* __restrict__ attr_word, const float* __restrict__ attr_pos,
                    const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int d,
                    float* __restrict__ preds_f32, float* __restrict__ preds_T, int64_t ld_preds_T,
                    int64_t* __restrict__ labels, T* __restrict__ memory, int mem_rows, int mem_row0) {
  __shared__ float prob[MAX_ATTR];
  __shared__ float sort_p[MAX_ATTR];
  __shared__ int sort_i[MAX_ATTR];
  __shared__ int sel[MAX_TOPK];
  const int v = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int a = tid; a < n_attr; a += THREADS) {
    // as written in the reference: 1 - exp(sum_seq log(clamp(1 - sigmoid(s), 1e-12, 1))), seq == 1
    const float s = scores[(int64_t)v * ld_scores + a];
    const float p = 1.0f / (1.0f + expf(-s));
    const float raw = logf(fminf(fmaxf(1.0f - p, 1e-12f), 1.0f));
    const float out = 1.0f - expf(raw);
    prob[a] = out;
    if (preds_f32) preds_f32[(int64_t)v * n_attr + a] = out;
    if (preds_T) preds_T[(int64_t)v * ld_preds_T + a] = out;
  }
  if (preds_T)
    for (int a = n_attr + tid; a < ld_preds_T; a += THREADS) preds_T[(int64_t)v * ld_preds_T + a] = 0.f;
  __syncthreads();

  // order the concepts by (probability desc, index asc): bitonic sort of the padded array in shared memory
  // (45 compare-exchange steps for 512 entries instead of 500 compares per thread)
  int n2 = 1;
  while (n2 < n_attr) n2 <<= 1;
  for (int a = tid; a < n2; a += THREADS) {
    sort_p[a] = a < n_attr ? prob[a] : -INFINITY;
    sort_i[a] = a < n_attr ? a : INT_MAX;
  }
  __syncthreads();
  for (int k = 2; k <= n2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < n2; t += THREADS) {
        const int u = t ^ j;
        if (u > t) {
          const float pa = sort_p[t], pb = sort_p[u];
          const int ia = sort_i[t], ib = sort_i[u];
          const bool a_first = pa > pb || (pa == pb && ia < ib);
          const bool keep = ((t & k) == 0) ? a_first : !a_first;
          if (!keep) {
            sort_p[t] = pb; sort_i[t] = ib;
            sort_p[u] = pa; sort_i[u] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int r = tid; r < topk; r += THREADS) {
    sel[r] = sort_i[r];
    if (labels) labels[(int64_t)v * topk + r] = sort_i[r];
  }
  __syncthreads();
  if (memory == nullptr) return;

  const int nch = d / 128;
  for (int r = warp; r < topk; r += THREADS / 32) {
    const int a = sel[r];
    float x[MAX_CHUNKS][4];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c)
      if (c < nch) {
        const int col = c * 128 + lane * 4;
        float w[4], q[4];
        Act<float>::load4(attr_word + (int64_t)a * d + col, w);
        Act<float>::load4(attr_pos + (int64_t)r * d + col, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          x[c][j] = w[j] + q[j];
          s += x[c][j];
        }
      }
    const float mean = warp_sum(s) / (float)d;
    float qq = 0.f;
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c)
      if (c < nch)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float t = x[c][j] - mean;
          qq += t * t;
        }
    const float rstd = 1.0f / sqrtf(warp_sum(qq) / (float)d + eps);
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c)
      if (c < nch) {
        const int col = c * 128 + lane * 4;
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
        const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
        float o[4];
        o[0] = (x[c][0] - mean) * rstd * g.x + b.x;
        o[1] = (x[c][1] - mean) * rstd * g.y + b.y;
        o[2] = (x[c][2] - mean) * rstd * g.z + b.z;
        o[3] = (x[c][3] - mean) * rstd * g.w + b.w;
        Act<T>::store4(memory + ((int64_t)v * mem_rows + mem_row0 + r) * d + col, o);
      }
  }
}

}  // namespace concept_head
}  // namespace care

using namespace care;

extern "C" int care_concept_head(care_ctx* ctx, int dtype, const float* scores, int64_t ld_scores, int B, int n_attr,
                                 int topk, const float* attr_word, const float* attr_pos, const float* gamma,
                                 const float* beta, float eps, int d, float* preds_f32, void* preds_T,
                                 int64_t ld_preds_T, int64_t* labels, void* memory, int mem_rows, int mem_row0,
                                 void* stream) {
  CARE_CHECK_ARG(ctx && scores && B > 0, "care_concept_head: bad args");
  CARE_CHECK_ARG(n_attr > 0 && n_attr <= concept_head::MAX_ATTR, "care_concept_head: n_attr=%d must be <= %d", n_attr,
                 concept_head::MAX_ATTR);
  CARE_CHECK_ARG(topk > 0 && topk <= concept_head::MAX_TOPK && topk <= n_attr, "care_concept_head: topk=%d must be <= %d",
                 topk, concept_head::MAX_TOPK);
  CARE_CHECK_ARG(memory == nullptr || (attr_word && attr_pos && gamma && beta && d % 128 == 0 && d <= 1024),
                 "care_concept_head: embedding args (d=%d must be a multiple of 128, <= 1024)", d);
  CARE_CHECK_ARG(preds_T == nullptr || ld_preds_T >= n_attr, "care_concept_head: ld_preds_T too small");
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32)
    concept_head::concept_head_kernel<float><<<B, concept_head::THREADS, 0, s>>>(
        scores, ld_scores, n_attr, topk, attr_word, attr_pos, gamma, beta, eps, d, preds_f32, (float*)preds_T,
        ld_preds_T, labels, (float*)memory, mem_rows, mem_row0);
  else if (dtype == CARE_H16)
    concept_head::concept_head_kernel<h16><<<B, concept_head::THREADS, 0, s>>>(
        scores, ld_scores, n_attr, topk, attr_word, attr_pos, gamma, beta, eps, d, preds_f32, (float*)preds_T,
        ld_preds_T, labels, (h16*)memory, mem_rows, mem_row0);
  else {
    care::set_error("care_concept_head: bad dtype %d", dtype);
    return -1;
  }
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}
