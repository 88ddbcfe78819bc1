// Row-wise HBM-bound kernels: feature cast, encoder LayerNorm(+temporal mean), HighWay+BN tail,
// decoder input embedding + LayerNorm, residual + LayerNorm.  One warp per row, 16-byte accesses,
// LayerNorm statistics in fp32 with a two-pass (mean, then centred variance) reduction.
#include "step_prologue.cuh"

namespace care {
namespace rw {

// ---------------------------------------------------------------------------------------------
// dst[r] = [hi | lo | hi * 2^-11] (TERMS == 3) or [hi] (TERMS == 1) of src[r], each part `cols_pad` wide (zero
// padded): hi = h16(x), lo = h16(x - hi).  With W' = [W_hi | W_hi | W_lo * 2^11] one K-concatenated tensor-core
// GEMM then computes A_hi W_hi + A_lo W_hi + A_hi W_lo, i.e. the fp32 product to ~2^-21 relative.  The power-of-
// two exchange between the third parts keeps W_lo (~2^-11 |W|, i.e. ~1e-5 for a typical weight) out of fp16's
// subnormal range; hi * 2^-11 only goes subnormal for |x| < 0.125, where its product with W_lo is negligible.
template <int TERMS>
__global__ void split_f32_h16_kernel(const float* __restrict__ src, int64_t ld_src, int64_t rows, int cols,
                                     int cols_pad, h16* __restrict__ dst) {
  const int groups = cols_pad / 8;
  const int64_t total = rows * groups;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t ld_dst = (int64_t)TERMS * cols_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t r = i / groups;
    const int c0 = (int)(i - r * groups) * 8;
    float v[8], lo[8];
    if (c0 + 8 <= cols && (ld_src & 3) == 0) {
      Act<float>::load8(src + r * ld_src + c0, v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = c0 + j < cols ? src[r * ld_src + c0 + j] : 0.f;
    }
    h16* o = dst + r * ld_dst + c0;
    if (TERMS == 1) {
      Act<h16>::store8(o, v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float hi = h16_to_float(float_to_h16(v[j]));
        lo[j] = v[j] - hi;
        v[j] = hi;
      }
      Act<h16>::store8(o, v);
      Act<h16>::store8(o + cols_pad, lo);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= (1.0f / 2048.0f);
      Act<h16>::store8(o + 2 * cols_pad, v);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Encoder tails.  grid = B videos, block = 8 warps; warp w handles rows w, w+8, ... of its video and
// keeps a running sum of the normalised rows; the 8 partial sums are combined through smem.
template <typename T, bool HIGHWAY>
__global__ void __launch_bounds__(256)
encoder_tail_kernel(const float* __restrict__ x, const float* __restrict__ ypre, const float* __restrict__ gpre,
                    const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                    const float* __restrict__ p3, float eps, int Tn, int d, T* __restrict__ out, int out_rows,
                    int out_row0, float* __restrict__ mean_out, int64_t mean_ld, int mean_col0) {
  __shared__ float red[8][MAX_D];
  const int v = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nch = d / 128;
  float acc[MAX_CHUNKS][4];
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
  for (int t = warp; t < Tn; t += 8) {
    const int64_t roff = ((int64_t)v * Tn + t) * d;
    float r[MAX_CHUNKS][4];
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c)
      if (c < nch) {
        const int col = c * 128 + lane * 4;
        Act<float>::load4(x + roff + col, r[c]);
        if (HIGHWAY) {
          // Encoder.py:219-226: gate*x + (1-gate)*tanh(w1 x); then BN1d eval (:229-241)
          float yv[4], gv[4];
          Act<float>::load4(ypre + roff + col, yv);
          Act<float>::load4(gpre + roff + col, gv);
          const float4 mu = __ldg(reinterpret_cast<const float4*>(p0 + col));
          const float4 var = __ldg(reinterpret_cast<const float4*>(p1 + col));
          const float4 w = __ldg(reinterpret_cast<const float4*>(p2 + col));
          const float4 b = __ldg(reinterpret_cast<const float4*>(p3 + col));
          const float mus[4] = {mu.x, mu.y, mu.z, mu.w}, vars[4] = {var.x, var.y, var.z, var.w};
          const float ws[4] = {w.x, w.y, w.z, w.w}, bs[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float y = tanhf(yv[j]);
            const float g = 1.f / (1.f + expf(-gv[j]));
            const float hmix = g * r[c][j] + (1.f - g) * y;
            r[c][j] = (hmix - mus[j]) * (1.0f / sqrtf(vars[j] + eps)) * ws[j] + bs[j];
          }
        }
      }
    if (!HIGHWAY) warp_layernorm(r, nch, d, lane, p0, p1, eps);
#pragma unroll
    for (int c = 0; c < MAX_CHUNKS; ++c)
      if (c < nch) {
        const int col = c * 128 + lane * 4;
        if (out != nullptr) Act<T>::store4(out + ((int64_t)v * out_rows + out_row0 + t) * d + col, r[c]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[c][j] += r[c][j];
      }
  }
  if (mean_out == nullptr) return;
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][c * 128 + lane * 4 + j] = acc[c][j];
  __syncthreads();
  for (int col = threadIdx.x; col < d; col += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][col];
    mean_out[(int64_t)v * mean_ld + mean_col0 + col] = s / (float)Tn;
  }
}

// ---------------------------------------------------------------------------------------------
// out[r] = LN(((word[tok] + pos[p]) + add[r/rpv]) + gsg[r/rpv])      (Embeddings.py:134-188)
// the live-slot records of the step's self-attention, written by extra CTAs of the embedding launch (care_ctx_request_records)
struct RecordsJob {
  const uint8_t* anc;
  const int32_t* tok_hist;
  const int32_t* done;
  uint32_t* info;      // NULL: no job
  int anc_stride, tok_stride, B, K, n_pos;
  int first_block;     // blocks >= first_block build records, one warp per video
};

template <typename T>
__global__ void __launch_bounds__(256)
embed_ln_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ positions, int position,
                const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ add,
                const float* __restrict__ gsg, int rpv, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int R, int d, T* __restrict__ out,
                float* __restrict__ out32, const RecordsJob job, const EarlyExit ee) {
  pdl_wait();   // first kernel of a step: the early-exit counter and the tokens come from the beam kernel just before
  pdl_launch_dependents();
  if (all_done(ee)) return;
  if (job.info != nullptr && (int)blockIdx.x >= job.first_block) {
    __shared__ uint32_t rec_all[8][attn_mma::INFO_WORDS];
    const int warp = threadIdx.x >> 5;
    const int v = ((int)blockIdx.x - job.first_block) * 8 + warp;
    if (v >= job.B) return;
    if (job.done != nullptr && job.done[v]) return;
    attn_mma::warp_compact_record(job.anc, job.anc_stride, job.tok_hist, job.tok_stride, v, job.K, job.n_pos,
                                  threadIdx.x & 31, rec_all[warp], job.info);
    return;
  }
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const int vid = row / rpv;
  warp_embed_ln_row<T>(tokens[row], positions ? positions[row] : position, word, pos,
                       add ? add + (int64_t)vid * d : nullptr, gsg ? gsg + (int64_t)vid * d : nullptr, gamma, beta, eps, d,
                       lane, out + (int64_t)row * d, out32 ? out32 + (int64_t)row * d : nullptr);
}

// out = LN(x + residual)   (SubLayers.py:74-79, 148-150)
template <typename T>
__global__ void __launch_bounds__(256)
add_ln_kernel(const float* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, int R, int d, T* __restrict__ out, const EarlyExit ee) {
  pdl_wait();
  pdl_launch_dependents();
  if (all_done(ee)) return;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const int nch = d / 128;
  float r[MAX_CHUNKS][4];
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch) {
      const int col = c * 128 + lane * 4;
      float q[4];
      Act<float>::load4(x + (int64_t)row * d + col, r[c]);
      Act<T>::load4(res + (int64_t)row * d + col, q);
#pragma unroll
      for (int j = 0; j < 4; ++j) r[c][j] += q[j];
    }
  warp_layernorm(r, nch, d, lane, gamma, beta, eps);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch) Act<T>::store4(out + (int64_t)row * d + c * 128 + lane * 4, r[c]);
}

static int check_d(int d, const char* who) {
  CARE_CHECK_ARG(d > 0 && d % 128 == 0 && d <= MAX_D, "%s: d=%d must be a multiple of 128 and <= %d", who, d, MAX_D);
  return 0;
}

}  // namespace rw
}  // namespace care

using namespace care;

extern "C" {

int care_split_f32_h16(care_ctx* ctx, const float* src, int64_t ld_src, int64_t rows, int cols, int cols_pad,
                       int terms, void* dst, void* stream) {
  CARE_CHECK_ARG(ctx && src && dst && rows >= 0 && cols > 0 && ld_src >= cols, "care_split_f32_h16: bad args");
  CARE_CHECK_ARG(cols_pad >= cols && cols_pad % 8 == 0, "care_split_f32_h16: cols_pad %d must be a multiple of 8 >= cols %d",
                 cols_pad, cols);
  CARE_CHECK_ARG(terms == 1 || terms == 3, "care_split_f32_h16: terms must be 1 or 3");
  CARE_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
                 "care_split_f32_h16: src and dst must be 16-byte aligned");
  if (rows == 0) return 0;
  const int64_t total = rows * (cols_pad / 8);
  int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 16);
  if (terms == 1)
    rw::split_f32_h16_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, cols_pad, (h16*)dst);
  else
    rw::split_f32_h16_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, cols, cols_pad, (h16*)dst);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_encoder_ln_mean(care_ctx* ctx, int dtype, const float* x, const float* gamma, const float* beta, float eps,
                         int B, int T, int d, void* out, int out_rows, int out_row0, float* mean_out, int64_t mean_ld,
                         int mean_col0, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_encoder_ln_mean");
  CARE_CHECK_ARG(ctx && x && gamma && beta && B > 0 && T > 0, "care_encoder_ln_mean: bad args");
  CARE_CHECK_ARG(out == nullptr || (out_row0 >= 0 && out_row0 + T <= out_rows),
                 "care_encoder_ln_mean: rows %d..%d do not fit the %d-row memory", out_row0, out_row0 + T, out_rows);
  CARE_CHECK_ARG(dtype == CARE_F32 || dtype == CARE_H16, "care_encoder_ln_mean: this build computes in fp32 / " CARE_H16_NAME);
  if (rw::check_d(d, "care_encoder_ln_mean")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32)
    rw::encoder_tail_kernel<float, false><<<B, 256, 0, s>>>(x, nullptr, nullptr, gamma, beta, nullptr, nullptr, eps, T,
                                                            d, (float*)out, out_rows, out_row0, mean_out, mean_ld,
                                                            mean_col0);
  else
    rw::encoder_tail_kernel<h16, false><<<B, 256, 0, s>>>(
        x, nullptr, nullptr, gamma, beta, nullptr, nullptr, eps, T, d, (h16*)out, out_rows, out_row0, mean_out,
        mean_ld, mean_col0);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_encoder_highway_bn_mean(care_ctx* ctx, int dtype, const float* h, const float* ypre, const float* gpre,
                                 const float* bn_mean, const float* bn_var, const float* bn_w, const float* bn_b,
                                 float bn_eps, int B, int T, int d, void* out, int out_rows, int out_row0,
                                 float* mean_out, int64_t mean_ld, int mean_col0, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_encoder_highway_bn_mean");
  CARE_CHECK_ARG(ctx && h && ypre && gpre && bn_mean && bn_var && bn_w && bn_b && B > 0 && T > 0,
                 "care_encoder_highway_bn_mean: bad args");
  CARE_CHECK_ARG(out == nullptr || (out_row0 >= 0 && out_row0 + T <= out_rows),
                 "care_encoder_highway_bn_mean: rows %d..%d do not fit the %d-row memory", out_row0, out_row0 + T, out_rows);
  CARE_CHECK_ARG(dtype == CARE_F32 || dtype == CARE_H16, "care_encoder_highway_bn_mean: this build computes in fp32 / " CARE_H16_NAME);
  if (rw::check_d(d, "care_encoder_highway_bn_mean")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32)
    rw::encoder_tail_kernel<float, true><<<B, 256, 0, s>>>(h, ypre, gpre, bn_mean, bn_var, bn_w, bn_b, bn_eps, T, d,
                                                           (float*)out, out_rows, out_row0, mean_out, mean_ld, mean_col0);
  else
    rw::encoder_tail_kernel<h16, true><<<B, 256, 0, s>>>(
        h, ypre, gpre, bn_mean, bn_var, bn_w, bn_b, bn_eps, T, d, (h16*)out, out_rows, out_row0, mean_out, mean_ld,
        mean_col0);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_embed_ln(care_ctx* ctx, int dtype, const int32_t* tokens, const int32_t* positions, int position,
                  const float* word_emb, const float* pos_emb, const float* add_feats, const float* gsg,
                  int rows_per_video, const float* gamma, const float* beta, float eps, int R, int d, void* out,
                  float* out32, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_embed_ln");
  CARE_CHECK_ARG(ctx && tokens && word_emb && pos_emb && gamma && beta && out && R > 0 && rows_per_video > 0,
                 "care_embed_ln: bad args");
  if (rw::check_d(d, "care_embed_ln")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  int grid = (R + 7) / 8;
  rw::RecordsJob job{};
  if (ctx->rec_req_armed) {
    ctx->rec_req_armed = false;
    const care_ctx::RecordsReq& q = ctx->rec_req;
    // the same conditions under which care_self_attn_step takes the chunk-stream kernel (attention_mma.cu: self_step)
    const bool stream_kernel = dtype == CARE_H16 && q.K <= 8 && q.n_pos * q.K <= 160 && q.n_pos <= 64 &&
                               (ctx->self_compact == 3 || (ctx->self_compact == 2 && q.n_pos >= 6 && q.B * q.H >= 1024)) &&
                               ctx->compact_info != nullptr && q.B <= ctx->compact_info_videos &&
                               (int64_t)q.anc_stride * q.B * q.K < (1LL << 31);
    if (stream_kernel) {
      job.anc = q.anc; job.tok_hist = q.tok_hist; job.done = q.done; job.info = ctx->compact_info;
      job.anc_stride = q.anc_stride; job.tok_stride = (q.anc_stride + 1) * q.K;
      job.B = q.B; job.K = q.K; job.n_pos = q.n_pos;
      job.first_block = grid;
      grid += (q.B + 7) / 8;
      ctx->info_ready_npos = q.n_pos;
      ctx->info_ready_B = q.B;
      ctx->info_ready_anc = q.anc;
    }
  }
  if (dtype == CARE_F32)
    CARE_CUDA(launch_pdl(ctx, rw::embed_ln_kernel<float>, dim3(grid), dim3(256), 0, s, tokens, positions, position, word_emb,
                         pos_emb, add_feats, gsg, rows_per_video, gamma, beta, eps, R, d, (float*)out, out32, job,
                         early_exit_of(ctx)));
  else
    CARE_CUDA(launch_pdl(ctx, rw::embed_ln_kernel<h16>, dim3(grid), dim3(256), 0, s, tokens, positions, position, word_emb,
                         pos_emb, add_feats, gsg, rows_per_video, gamma, beta, eps, R, d, (h16*)out, out32, job,
                         early_exit_of(ctx)));
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_add_ln(care_ctx* ctx, int dtype, const float* x, const void* residual, const float* gamma, const float* beta,
                float eps, int R, int d, void* out, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_add_ln");
  CARE_CHECK_ARG(ctx && x && residual && gamma && beta && out && R > 0, "care_add_ln: bad args");
  if (rw::check_d(d, "care_add_ln")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (R + 7) / 8;
  if (dtype == CARE_F32)
    CARE_CUDA(launch_pdl(ctx, rw::add_ln_kernel<float>, dim3(grid), dim3(256), 0, s, x, (const float*)residual, gamma, beta,
                         eps, R, d, (float*)out, early_exit_of(ctx)));
  else
    CARE_CUDA(launch_pdl(ctx, rw::add_ln_kernel<h16>, dim3(grid), dim3(256), 0, s, x, (const h16*)residual, gamma, beta, eps,
                         R, d, (h16*)out, early_exit_of(ctx)));
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // extern "C"
