// Fused decode-step attention (HBM-bound).  One warp per (video, head) serves all K beams of the
// video together, so the video's keys/values are streamed from HBM exactly once per step:
//   cross-attention: K/V of the 114-token memory, projected once per video and never replicated
//                    per beam; per-head hybrid bias added after masking (Attention.py:104-111);
//   self-attention : the un-reordered KV cache [T, R, 3d]; beam b attends position p through the
//                    ancestry table anc[b][p] (which cache slot holds its prefix) with the PAD-key
//                    mask taken from the token history (Transformer.py:15-29,169-174).
// Lane layout: 4 key groups x 8 dim slices; each lane loads 8 consecutive head dims (16 B in bf16,
// 2 x 16 B in fp32) so a warp load covers 4 full 128/256-byte key rows.  Scores go through shared
// memory, softmax statistics through warp shuffles, everything accumulates in fp32.
#include "common.cuh"

namespace care {
namespace attn {

constexpr int DH = 64;

struct Params {
  const void* q;      // element pointer to q of (row 0, head 0)
  int64_t q_ld;       // elements between consecutive rows of q
  const void* kv;     // base of the key/value storage
  int64_t kv_ld;      // elements between consecutive storage rows
  int k_off, v_off;   // element offsets of k / v inside a storage row (before the head offset)
  int n_keys;         // keys per video: Lm (cross) or n_pos * K (self)
  int Lm;             // cross: memory length
  int R;              // self: rows per cache position (= B*K)
  int K, H, d, hpc;
  const float* bias;  // cross: [H, Lm] or NULL
  const uint8_t* anc; // self: [B, K, anc_stride]
  int anc_stride;
  const int32_t* tok_hist;  // self: [B, T_max+1, K]
  int tok_stride;           // (T_max+1)*K
  int n_pos;
  const int32_t* done;
  void* out;          // [R, d]
  int sc_ld;
};

template <typename T, int KB, bool SELF>
__global__ void __launch_bounds__(128) attn_step_kernel(const Params p) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int groups = p.H / p.hpc;
  const int v = blockIdx.x / groups;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = (blockIdx.x % groups) * p.hpc + warp;
  if (p.done != nullptr && p.done[v]) return;
  const int K = p.K, n_keys = p.n_keys, sc_ld = p.sc_ld;
  float* sc = reinterpret_cast<float*>(smem) + (size_t)warp * KB * sc_ld;
  uint8_t* valid = smem + (size_t)p.hpc * KB * sc_ld * sizeof(float);

  if (SELF) {
    // valid[b][j]: key j = (position pp, slot s) belongs to beam b's prefix and is not <pad>
    for (int idx = threadIdx.x; idx < K * n_keys; idx += blockDim.x) {
      const int b = idx / n_keys, j = idx - b * n_keys;
      const int pp = j / K, s = j - pp * K;
      const int slot = (pp == p.n_pos - 1) ? b : (int)p.anc[((int64_t)v * K + b) * p.anc_stride + pp];
      const int tok = p.tok_hist[(int64_t)v * p.tok_stride + pp * K + s];
      valid[b * sc_ld + j] = (slot == s && tok != CARE_PAD) ? 1 : 0;
    }
    __syncthreads();
  }

  const int g = lane >> 3, sub = lane & 7;
  const T* qbase = reinterpret_cast<const T*>(p.q) + h * DH + sub * 8;
  const T* kvbase = reinterpret_cast<const T*>(p.kv) + h * DH + sub * 8;
  auto key_row = [&](int j) -> int64_t {
    if (SELF) {
      const int pp = j / K, s = j - pp * K;
      return ((int64_t)pp * p.R + (int64_t)v * K + s) * p.kv_ld;
    }
    return ((int64_t)v * p.Lm + j) * p.kv_ld;
  };

  // ---- pass 1: scores -------------------------------------------------------------------------
  {
    float qf[KB][8];
#pragma unroll
    for (int b = 0; b < KB; ++b) {
      if (b < K) {
        Act<T>::load8(qbase + ((int64_t)v * K + b) * p.q_ld, qf[b]);
      } else {
#pragma unroll
        for (int x = 0; x < 8; ++x) qf[b][x] = 0.f;
      }
    }
    const float inv_scale_div = 8.0f;  // sqrt(DH); scores are divided, as in Attention.py:84
    const int n_iter = (n_keys + 3) >> 2;
#pragma unroll 4
    for (int i = 0; i < n_iter; ++i) {
      const int j = 4 * i + g;
      float kf[8];
      if (j < n_keys) {
        Act<T>::load8(kvbase + key_row(j) + p.k_off, kf);
      } else {
#pragma unroll
        for (int x = 0; x < 8; ++x) kf[x] = 0.f;
      }
      float part[KB];
#pragma unroll
      for (int b = 0; b < KB; ++b) {
        float s = 0.f;
#pragma unroll
        for (int x = 0; x < 8; ++x) s = fmaf(qf[b][x], kf[x], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        part[b] = s;
      }
      if (sub == 0 && j < n_keys) {
#pragma unroll
        for (int b = 0; b < KB; ++b) {
          if (b < K) {
            float s = part[b] / inv_scale_div;
            if (SELF) {
              if (!valid[b * sc_ld + j]) s = -1e9f;
            } else if (p.bias != nullptr) {
              s += __ldg(p.bias + (int64_t)h * p.Lm + j);
            }
            sc[b * sc_ld + j] = s;
          }
        }
      }
    }
  }
  __syncwarp();

  // ---- softmax over keys, one beam at a time ------------------------------------------------------
  for (int b = 0; b < K; ++b) {
    float* row = sc + b * sc_ld;
    float m = -INFINITY;
    for (int j = lane; j < n_keys; j += 32) m = fmaxf(m, row[j]);
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < n_keys; j += 32) {
      const float e = expf(row[j] - m);
      row[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    for (int j = lane; j < n_keys; j += 32) row[j] = row[j] / sum;
  }
  __syncwarp();

  // ---- pass 2: context = P * V -----------------------------------------------------------------------
  float acc[KB][8];
#pragma unroll
  for (int b = 0; b < KB; ++b)
#pragma unroll
    for (int x = 0; x < 8; ++x) acc[b][x] = 0.f;
  {
    const int n_iter = (n_keys + 3) >> 2;
#pragma unroll 4
    for (int i = 0; i < n_iter; ++i) {
      const int j = 4 * i + g;
      if (j < n_keys) {
        float vf[8];
        Act<T>::load8(kvbase + key_row(j) + p.v_off, vf);
#pragma unroll
        for (int b = 0; b < KB; ++b) {
          if (b < K) {
            const float pr = sc[b * sc_ld + j];
#pragma unroll
            for (int x = 0; x < 8; ++x) acc[b][x] = fmaf(pr, vf[x], acc[b][x]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < KB; ++b)
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      float a = acc[b][x];
      a += __shfl_xor_sync(0xffffffffu, a, 8);
      a += __shfl_xor_sync(0xffffffffu, a, 16);
      acc[b][x] = a;
    }
  T* out = reinterpret_cast<T*>(p.out) + h * DH + sub * 8;
#pragma unroll
  for (int b = 0; b < KB; ++b)
    if (b < K && g == (b & 3)) Act<T>::store8(out + ((int64_t)v * K + b) * p.d, acc[b]);
}

template <typename T, bool SELF>
static int launch(care_ctx* ctx, const Params& p, int B, cudaStream_t stream) {
  const int grid = B * (p.H / p.hpc);
  const int threads = 32 * p.hpc;
  auto go = [&](auto kern, int kb) -> int {
    size_t smem = (size_t)p.hpc * kb * p.sc_ld * sizeof(float) + (SELF ? (size_t)kb * p.sc_ld : 0);
    if (smem > 48 * 1024) CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, stream>>>(p);
    CARE_LAUNCH_CHECK(ctx);
    return 0;
  };
  if (p.K <= 1) return go(attn_step_kernel<T, 1, SELF>, 1);
  if (p.K <= 3) return go(attn_step_kernel<T, 3, SELF>, 3);
  if (p.K <= 5) return go(attn_step_kernel<T, 5, SELF>, 5);
  return go(attn_step_kernel<T, 8, SELF>, 8);
}

static int common_checks(const char* who, int B, int K, int H, int d) {
  CARE_CHECK_ARG(B > 0 && K > 0 && K <= 8, "%s: beam size K=%d must be in [1, 8]", who, K);
  CARE_CHECK_ARG(H > 0 && d == H * DH, "%s: head size must be 64 (d=%d, H=%d)", who, d, H);
  return 0;
}

}  // namespace attn

namespace attn_mma {  // attention_mma.cu: return 1 when the shape is not covered
int cross_step(care_ctx* ctx, const void* q, int64_t ldq, const void* kv, int Lm, int B, int K, int H, int d,
               const float* hybrid_bias, const int32_t* done, void* ctx_out, cudaStream_t stream);
int self_step(care_ctx* ctx, const void* cache, int n_pos, int B, int K, int H, int d, const uint8_t* anc,
              int anc_stride, const int32_t* tok_hist, const int32_t* done, void* ctx_out, cudaStream_t stream);
}  // namespace attn_mma
}  // namespace care

using namespace care;

extern "C" {

int care_self_attn_step(care_ctx* ctx, int dtype, const void* cache, int n_pos, int B, int K, int H, int d,
                        const uint8_t* anc, int anc_stride, const int32_t* tok_hist, const int32_t* done,
                        void* ctx_out, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_self_attn_step");
  CARE_CHECK_ARG(ctx && cache && anc && tok_hist && ctx_out && n_pos >= 1, "care_self_attn_step: bad args");
  if (attn::common_checks("care_self_attn_step", B, K, H, d)) return -1;
  if (dtype == CARE_H16 && ctx->attn_impl == 1) {
    const int rc = attn_mma::self_step(ctx, cache, n_pos, B, K, H, d, anc, anc_stride, tok_hist, done, ctx_out,
                                       (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  attn::Params p{};
  const int64_t ld = 3LL * d;
  const size_t esz = dtype == CARE_F32 ? 4 : 2;
  p.q = static_cast<const uint8_t*>(cache) + (size_t)(n_pos - 1) * B * K * ld * esz;
  p.q_ld = ld;
  p.kv = cache;
  p.kv_ld = ld;
  p.k_off = d;
  p.v_off = 2 * d;
  p.n_keys = n_pos * K;
  p.R = B * K;
  p.K = K; p.H = H; p.d = d;
  p.hpc = (H % 4 == 0) ? 4 : 1;
  p.anc = anc;
  p.anc_stride = anc_stride;
  p.tok_hist = tok_hist;
  p.tok_stride = (anc_stride + 1) * K;
  p.n_pos = n_pos;
  p.done = done;
  p.out = ctx_out;
  p.sc_ld = (p.n_keys + 3) & ~3;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32) return attn::launch<float, true>(ctx, p, B, s);
  if (dtype == CARE_H16) return attn::launch<h16, true>(ctx, p, B, s);
  care::set_error("care_self_attn_step: bad dtype %d", dtype);
  return -1;
}

int care_cross_attn_step(care_ctx* ctx, int dtype, const void* q, int64_t ldq, const void* kv, int Lm, int B, int K,
                         int H, int d, const float* hybrid_bias, const int32_t* done, void* ctx_out, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_cross_attn_step");
  CARE_CHECK_ARG(ctx && q && kv && ctx_out && Lm >= 1, "care_cross_attn_step: bad args");
  if (attn::common_checks("care_cross_attn_step", B, K, H, d)) return -1;
  CARE_CHECK_ARG(ldq % 8 == 0, "care_cross_attn_step: ldq must be a multiple of 8");
  if (dtype == CARE_H16 && ctx->attn_impl == 1) {
    const int rc = attn_mma::cross_step(ctx, q, ldq, kv, Lm, B, K, H, d, hybrid_bias, done, ctx_out,
                                        (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  attn::Params p{};
  p.q = q;
  p.q_ld = ldq;
  p.kv = kv;
  p.kv_ld = 2LL * d;
  p.k_off = 0;
  p.v_off = d;
  p.n_keys = Lm;
  p.Lm = Lm;
  p.R = B * K;
  p.K = K; p.H = H; p.d = d;
  p.hpc = (H % 4 == 0) ? 4 : 1;
  p.bias = hybrid_bias;
  p.done = done;
  p.out = ctx_out;
  p.sc_ld = (Lm + 3) & ~3;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_F32) return attn::launch<float, false>(ctx, p, B, s);
  if (dtype == CARE_H16) return attn::launch<h16, false>(ctx, p, B, s);
  care::set_error("care_cross_attn_step: bad dtype %d", dtype);
  return -1;
}

}  // extern "C"
