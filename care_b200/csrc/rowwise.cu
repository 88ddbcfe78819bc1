// Row-wise HBM-bound kernels: feature cast, encoder LayerNorm(+temporal mean), HighWay+BN tail,
// decoder input embedding + LayerNorm, residual + LayerNorm.  One warp per row, 16-byte accesses,
// LayerNorm statistics in fp32 with a two-pass (mean, then centred variance) reduction.
#include "common.cuh"

namespace care {
namespace rw {

constexpr int MAX_D = 1024;           // per-lane register budget: MAX_D / 32 / 4 float4 chunks
constexpr int MAX_CHUNKS = MAX_D / 128;

// ---------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n8,
                                     int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    float v[8];
    Act<float>::load8(src + i * 8, v);
    Act<__nv_bfloat16>::store8(dst + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n8 * 8; i < n; ++i) dst[i] = __float2bfloat16_rn(src[i]);
}

// LayerNorm of a row held as `nch` float4 chunks per lane (chunk c covers columns c*128 + lane*4 .. +3).
__device__ __forceinline__ void warp_layernorm(float (&x)[MAX_CHUNKS][4], int nch, int d, int lane,
                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                               float eps) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch)
#pragma unroll
      for (int j = 0; j < 4; ++j) s += x[c][j];
  const float mean = warp_sum(s) / (float)d;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float t = x[c][j] - mean;
        q += t * t;
      }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)d + eps);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch) {
      const int col = c * 128 + lane * 4;
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
      x[c][0] = (x[c][0] - mean) * rstd * g.x + b.x;
      x[c][1] = (x[c][1] - mean) * rstd * g.y + b.y;
      x[c][2] = (x[c][2] - mean) * rstd * g.z + b.z;
      x[c][3] = (x[c][3] - mean) * rstd * g.w + b.w;
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
                    int out_row0, T* __restrict__ mean_out, int64_t mean_ld, int mean_col0) {
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
    mean_out[(int64_t)v * mean_ld + mean_col0 + col] = Act<T>::from_float(s / (float)Tn);
  }
}

// ---------------------------------------------------------------------------------------------
// out[r] = LN(((word[tok] + pos[p]) + add[r/rpv]) + gsg[r/rpv])      (Embeddings.py:134-188)
template <typename T>
__global__ void __launch_bounds__(256)
embed_ln_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ positions, int position,
                const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ add,
                const float* __restrict__ gsg, int rpv, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int R, int d, T* __restrict__ out, const EarlyExit ee) {
  if (all_done(ee)) return;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= R) return;
  const int nch = d / 128;
  const int tok = tokens[row];
  const int p = positions ? positions[row] : position;
  const int vid = row / rpv;
  float r[MAX_CHUNKS][4];
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch) {
      const int col = c * 128 + lane * 4;
      float w[4], q[4];
      Act<float>::load4(word + (int64_t)tok * d + col, w);
      Act<float>::load4(pos + (int64_t)p * d + col, q);
#pragma unroll
      for (int j = 0; j < 4; ++j) r[c][j] = w[j] + q[j];
      if (add) {
        Act<float>::load4(add + (int64_t)vid * d + col, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) r[c][j] += q[j];
      }
      if (gsg) {
        Act<float>::load4(gsg + (int64_t)vid * d + col, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) r[c][j] += q[j];
      }
    }
  warp_layernorm(r, nch, d, lane, gamma, beta, eps);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch) Act<T>::store4(out + (int64_t)row * d + c * 128 + lane * 4, r[c]);
}

// out = LN(x + residual)   (SubLayers.py:74-79, 148-150)
template <typename T>
__global__ void __launch_bounds__(256)
add_ln_kernel(const float* __restrict__ x, const T* __restrict__ res, const float* __restrict__ gamma,
              const float* __restrict__ beta, float eps, int R, int d, T* __restrict__ out, const EarlyExit ee) {
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

int care_cast_f32_bf16(care_ctx* ctx, const float* src, void* dst, int64_t n, void* stream) {
  CARE_CHECK_ARG(ctx && src && dst && n >= 0, "care_cast_f32_bf16: bad args");
  if (n == 0) return 0;
  const int64_t n8 = n / 8;
  int blocks = (int)std::min<int64_t>((n8 + 255) / 256 + 1, (int64_t)ctx->sm_count * 16);
  rw::cast_f32_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n8, n);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_encoder_ln_mean(care_ctx* ctx, int dtype, const float* x, const float* gamma, const float* beta, float eps,
                         int B, int T, int d, void* out, int out_rows, int out_row0, void* mean_out, int64_t mean_ld,
                         int mean_col0, void* stream) {
  CARE_CHECK_ARG(ctx && x && gamma && beta && B > 0 && T > 0, "care_encoder_ln_mean: bad args");
  if (rw::check_d(d, "care_encoder_ln_mean")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32)
    rw::encoder_tail_kernel<float, false><<<B, 256, 0, s>>>(x, nullptr, nullptr, gamma, beta, nullptr, nullptr, eps, T,
                                                            d, (float*)out, out_rows, out_row0, (float*)mean_out,
                                                            mean_ld, mean_col0);
  else
    rw::encoder_tail_kernel<__nv_bfloat16, false><<<B, 256, 0, s>>>(
        x, nullptr, nullptr, gamma, beta, nullptr, nullptr, eps, T, d, (__nv_bfloat16*)out, out_rows, out_row0,
        (__nv_bfloat16*)mean_out, mean_ld, mean_col0);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_encoder_highway_bn_mean(care_ctx* ctx, int dtype, const float* h, const float* ypre, const float* gpre,
                                 const float* bn_mean, const float* bn_var, const float* bn_w, const float* bn_b,
                                 float bn_eps, int B, int T, int d, void* out, int out_rows, int out_row0,
                                 void* mean_out, int64_t mean_ld, int mean_col0, void* stream) {
  CARE_CHECK_ARG(ctx && h && ypre && gpre && bn_mean && bn_var && bn_w && bn_b && B > 0 && T > 0,
                 "care_encoder_highway_bn_mean: bad args");
  if (rw::check_d(d, "care_encoder_highway_bn_mean")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32)
    rw::encoder_tail_kernel<float, true><<<B, 256, 0, s>>>(h, ypre, gpre, bn_mean, bn_var, bn_w, bn_b, bn_eps, T, d,
                                                           (float*)out, out_rows, out_row0, (float*)mean_out, mean_ld,
                                                           mean_col0);
  else
    rw::encoder_tail_kernel<__nv_bfloat16, true><<<B, 256, 0, s>>>(
        h, ypre, gpre, bn_mean, bn_var, bn_w, bn_b, bn_eps, T, d, (__nv_bfloat16*)out, out_rows, out_row0,
        (__nv_bfloat16*)mean_out, mean_ld, mean_col0);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_embed_ln(care_ctx* ctx, int dtype, const int32_t* tokens, const int32_t* positions, int position,
                  const float* word_emb, const float* pos_emb, const float* add_feats, const float* gsg,
                  int rows_per_video, const float* gamma, const float* beta, float eps, int R, int d, void* out,
                  void* stream) {
  CARE_CHECK_ARG(ctx && tokens && word_emb && pos_emb && gamma && beta && out && R > 0 && rows_per_video > 0,
                 "care_embed_ln: bad args");
  if (rw::check_d(d, "care_embed_ln")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (R + 7) / 8;
  if (dtype == CARE_F32)
    rw::embed_ln_kernel<float><<<grid, 256, 0, s>>>(tokens, positions, position, word_emb, pos_emb, add_feats, gsg,
                                                    rows_per_video, gamma, beta, eps, R, d, (float*)out,
                                                    early_exit_of(ctx));
  else
    rw::embed_ln_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(tokens, positions, position, word_emb, pos_emb, add_feats,
                                                            gsg, rows_per_video, gamma, beta, eps, R, d,
                                                            (__nv_bfloat16*)out, early_exit_of(ctx));
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_add_ln(care_ctx* ctx, int dtype, const float* x, const void* residual, const float* gamma, const float* beta,
                float eps, int R, int d, void* out, void* stream) {
  CARE_CHECK_ARG(ctx && x && residual && gamma && beta && out && R > 0, "care_add_ln: bad args");
  if (rw::check_d(d, "care_add_ln")) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (R + 7) / 8;
  if (dtype == CARE_F32)
    rw::add_ln_kernel<float><<<grid, 256, 0, s>>>(x, (const float*)residual, gamma, beta, eps, R, d, (float*)out,
                                                  early_exit_of(ctx));
  else
    rw::add_ln_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x, (const __nv_bfloat16*)residual, gamma, beta, eps, R, d,
                                                          (__nv_bfloat16*)out, early_exit_of(ctx));
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // extern "C"
