// Pieces of a decode step's prologue that the beam kernel can run for the NEXT step (beam.cu), shared with the
// stand-alone kernels that otherwise run them (rowwise.cu: embed_ln_kernel, attention_mma.cu: compact_info_kernel).
#pragma once
#include "common.cuh"

namespace care {
namespace rw {

constexpr int MAX_D = 1024;           // per-lane register budget: MAX_D / 32 / 4 float4 chunks
constexpr int MAX_CHUNKS = MAX_D / 128;

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


// One decoder input row by one warp (Embeddings.py:134-188): out = LN(((word[tok] + pos[p]) + add) + gsg).
template <typename T>
__device__ __forceinline__ void warp_embed_ln_row(int tok, int p, const float* __restrict__ word,
                                                  const float* __restrict__ pos, const float* __restrict__ add_row,
                                                  const float* __restrict__ gsg_row, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, float eps, int d, int lane,
                                                  T* __restrict__ out_row, float* __restrict__ out32_row) {
  const int nch = d / 128;
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
      if (add_row) {
        Act<float>::load4(add_row + col, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) r[c][j] += q[j];
      }
      if (gsg_row) {
        Act<float>::load4(gsg_row + col, q);
#pragma unroll
        for (int j = 0; j < 4; ++j) r[c][j] += q[j];
      }
    }
  warp_layernorm(r, nch, d, lane, gamma, beta, eps);
#pragma unroll
  for (int c = 0; c < MAX_CHUNKS; ++c)
    if (c < nch) {
      Act<T>::store4(out_row + c * 128 + lane * 4, r[c]);
      if (out32_row != nullptr) Act<float>::store4(out32_row + c * 128 + lane * 4, r[c]);
    }
}

}  // namespace rw

namespace attn_mma {

// Per-video record consumed by the live-slot self-attention kernels, built once per step instead of once per head:
//   word 0: n_live;  words 1..64: key masks [8 beams][8 words] in the compacted index space;
//   words 65..: uint16 rowsrc[row] = (position << 4) | slot of the cache row gathered into tile row `row`.
constexpr int INFO_WORDS = 160;   // 1 + 64 + 80 (160 uint16) padded: 640 B per video

// One warp builds the record of video v for a prefix of n_pos positions (the newest position uses slot = beam) in
// `rec` (shared memory, INFO_WORDS words) and copies it to info[v].
__device__ __forceinline__ void warp_compact_record(const uint8_t* __restrict__ anc, int anc_stride,
                                                    const int32_t* __restrict__ tok_hist, int tok_stride, int v, int K,
                                                    int n_pos, int lane, uint32_t* rec, uint32_t* __restrict__ info) {
  for (int i = lane; i < INFO_WORDS; i += 32) rec[i] = 0u;
  __syncwarp();
  uint16_t* rowsrc = reinterpret_cast<uint16_t*>(rec + 65);
  int carry = 0;
  for (int p0 = 0; p0 < n_pos; p0 += 32) {
    const int pp = p0 + lane;
    uint32_t bits = 0u;
    uint32_t slots = 0u;   // 4 bits per beam
    if (pp < n_pos) {
      for (int b = 0; b < K; ++b) {
        const uint32_t slot = (pp == n_pos - 1) ? (uint32_t)b : (uint32_t)anc[((int64_t)v * K + b) * anc_stride + pp];
        bits |= 1u << slot;
        slots |= slot << (4 * b);
      }
    }
    const int cnt = __popc(bits);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    const int off = carry + incl - cnt;
    if (pp < n_pos) {
      for (int sl = 0; sl < K; ++sl)
        if ((bits >> sl) & 1u) rowsrc[off + __popc(bits & ((1u << sl) - 1u))] = (uint16_t)((pp << 4) | sl);
      for (int b = 0; b < K; ++b) {
        const uint32_t slot = (slots >> (4 * b)) & 15u;
        const int tok = tok_hist[(int64_t)v * tok_stride + pp * K + slot];
        if (tok != CARE_PAD) {
          const int j = off + __popc(bits & ((1u << slot) - 1u));
          atomicOr(&rec[1 + b * 8 + (j >> 5)], 1u << (j & 31));
        }
      }
    }
    carry += __shfl_sync(0xffffffffu, incl, 31);
  }
  __syncwarp();
  if (lane == 0) rec[0] = (uint32_t)carry;
  __syncwarp();
  for (int i = lane; i < INFO_WORDS; i += 32) info[(int64_t)v * INFO_WORDS + i] = rec[i];
}

}  // namespace attn_mma
}  // namespace care
