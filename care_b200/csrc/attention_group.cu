// Full-sequence ("group") attention: every query row of a group attends the group's key rows.
// Used where the reference runs the decoder over whole sequences:
//   * mask-predict passes (Translator_NARFormer / MaskPredict, na_algorithms.py:67-82,152-197):
//     bidirectional self-attention with the PAD-key mask (Transformer.py:170-171) - group = one
//     candidate sequence, nq = nk = L; cross-attention - group = one video, nq = (length candidates x L)
//     query rows sharing the video's memory K/V, nk = Lm, per-head hybrid bias;
//   * the stateless teacher-forced decoding_phase(input_ids) of the Framework API (causal mask,
//     Transformer.py:15-29,169-174).
// Score rule as in Attention.py:81-111: (q.k)/sqrt(64); masked keys := -1e9; + bias; softmax.
//
// bf16: one warp per (group, head); the K/V tiles are staged once by TMA (SWIZZLE_128B) and reused by
//       all query blocks of 16 rows; mma.sync.m16n8k16 for QK^T and PV, fp32 softmax in registers.
// fp32 (parity mode) or unsupported shapes: a plain SIMT kernel, one warp per (group, head, query row).
#include "dev_util.cuh"

namespace care {
namespace attn_group {

using namespace care::dev;

constexpr int DH = 64;

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  mma_m16n8k16(c, a[0], a[1], a[2], a[3], b0, b1);
}

struct Params {
  const void* q;            // q of (row 0, head 0)
  int64_t q_ld;
  const void* kv;           // key/value source matrix (SIMT path)
  int64_t kv_ld;
  int k_col, v_col;         // element column of K / V (head 0)
  int n_groups, nq, nk, H, d;
  int rows_pad;             // nk rounded up to 16
  const int32_t* key_tokens;  // [n_groups * nk] or NULL: key masked where token == <pad>
  int causal;                 // key k > query i masked (nq == nk)
  const float* bias;          // [H, nk] or NULL
  void* out;                  // [n_groups * nq, d]
};


// NT = key tiles of 8 (nk <= 8 * NT)
template <int NT>
__global__ void __launch_bounds__(32) group_attn_mma_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const int lane = threadIdx.x;
  const int grp = blockIdx.x / p.H, h = blockIdx.x - grp * p.H;
  const int nk = p.nk, nq = p.nq;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t tile_bytes = (uint32_t)p.rows_pad * 128u;
  const uint32_t k_s = base, v_s = base + tile_bytes;
  const uint32_t bar = base + 2u * tile_bytes;
  {
    uint8_t* v_gen = smem_raw + (v_s - raw);
    const int n16 = (p.rows_pad - nk) * 8;
    for (int i = lane; i < n16; i += 32)
      *reinterpret_cast<uint4*>(v_gen + (size_t)nk * 128 + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, 2u * (uint32_t)nk * 128u);
    tma_load_2d(k_s, &tmap, bar, p.k_col + h * DH, grp * nk);
    tma_load_2d(v_s, &tmap, bar, p.v_col + h * DH, grp * nk);
  }
  const int g = lane >> 2, tig = lane & 3;
  // key mask bits (bit k set = key k is <pad>), shared by every query row of the group
  uint32_t padmask[(NT + 3) / 4];
#pragma unroll
  for (int w = 0; w < (NT + 3) / 4; ++w) {
    uint32_t bits = 0u;
    if (p.key_tokens != nullptr) {
      const int k = w * 32 + lane;
      const bool pad = k < nk && p.key_tokens[(int64_t)grp * nk + k] == CARE_PAD;
      bits = __ballot_sync(0xffffffffu, pad);
    }
    padmask[w] = bits;
  }
  __syncwarp();
  mbar_wait(bar, 0);

  const h16* qbase = static_cast<const h16*>(p.q) + h * DH;
  h16* obase = static_cast<h16*>(p.out) + h * DH;
  const int m = lane >> 3, rr = lane & 7;
  for (int qb = 0; qb * 16 < nq; ++qb) {
    const int i0 = qb * 16 + g, i1 = i0 + 8;   // this thread's two query rows (within the group)
    const bool ok0 = i0 < nq, ok1 = i1 < nq;
    const h16* q0 = qbase + ((int64_t)grp * nq + (ok0 ? i0 : 0)) * p.q_ld;
    const h16* q1 = qbase + ((int64_t)grp * nq + (ok1 ? i1 : 0)) * p.q_ld;
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a0 = *reinterpret_cast<const uint32_t*>(q0 + 16 * ks + 2 * tig);
      const uint32_t a1 = *reinterpret_cast<const uint32_t*>(q1 + 16 * ks + 2 * tig);
      const uint32_t a2 = *reinterpret_cast<const uint32_t*>(q0 + 16 * ks + 2 * tig + 8);
      const uint32_t a3 = *reinterpret_cast<const uint32_t*>(q1 + 16 * ks + 2 * tig + 8);
      qa[ks][0] = ok0 ? a0 : 0u;
      qa[ks][1] = ok1 ? a1 : 0u;
      qa[ks][2] = ok0 ? a2 : 0u;
      qa[ks][3] = ok1 ? a3 : 0u;
    }
    float s[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      if (nt * 8 < nk) {
        const int r = nt * 8 + rr;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t b[4];
          ldsm_x4(b, k_s + sw128(r, 4 * kp + m));
          mma_bf16(c, qa[2 * kp], b[0], b[1]);
          mma_bf16(c, qa[2 * kp + 1], b[2], b[3]);
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) s[nt][e] = c[e];
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = nt * 8 + 2 * tig + e;
        const bool pad = (padmask[nt >> 2] >> (k & 31)) & 1u;
        float bv = 0.f;
        if (p.bias != nullptr && k < nk) bv = __ldg(p.bias + (int64_t)h * nk + k);
        float x0 = s[nt][e] * 0.125f, x1 = s[nt][2 + e] * 0.125f;
        if (pad || (p.causal && k > i0)) x0 = -1e9f;
        if (pad || (p.causal && k > i1)) x1 = -1e9f;
        x0 += bv;
        x1 += bv;
        if (k >= nk) {
          x0 = -INFINITY;
          x1 = -INFINITY;
        }
        s[nt][e] = x0;
        s[nt][2 + e] = x1;
        mx0 = fmaxf(mx0, x0);
        mx1 = fmaxf(mx1, x1);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float e0 = __expf(s[nt][e] - mx0), e1 = __expf(s[nt][2 + e] - mx1);
        s[nt][e] = e0;
        s[nt][2 + e] = e1;
        sum0 += e0;
        sum1 += e1;
      }
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dn][e] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {
      if (kk * 16 < nk) {
        uint32_t ah[4], al[4];
        ah[0] = pack_h16(s[2 * kk][0], s[2 * kk][1]);
        ah[1] = pack_h16(s[2 * kk][2], s[2 * kk][3]);
        ah[2] = pack_h16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        ah[3] = pack_h16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        al[0] = pack_h16(s[2 * kk][0] - h16_lo(ah[0]), s[2 * kk][1] - h16_hi(ah[0]));
        al[1] = pack_h16(s[2 * kk][2] - h16_lo(ah[1]), s[2 * kk][3] - h16_hi(ah[1]));
        al[2] = pack_h16(s[2 * kk + 1][0] - h16_lo(ah[2]), s[2 * kk + 1][1] - h16_hi(ah[2]));
        al[3] = pack_h16(s[2 * kk + 1][2] - h16_lo(ah[3]), s[2 * kk + 1][3] - h16_hi(ah[3]));
        const int r = kk * 16 + 8 * (m & 1) + rr;
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b[4];
          ldsm_x4_trans(b, v_s + sw128(r, 2 * dp + (m >> 1)));
          mma_bf16(o[2 * dp], ah, b[0], b[1]);
          mma_bf16(o[2 * dp], al, b[0], b[1]);
          mma_bf16(o[2 * dp + 1], ah, b[2], b[3]);
          mma_bf16(o[2 * dp + 1], al, b[2], b[3]);
        }
      }
    }
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    if (ok0) {
      h16* orow = obase + ((int64_t)grp * nq + i0) * p.d + 2 * tig;
#pragma unroll
      for (int dn = 0; dn < 8; ++dn)
        *reinterpret_cast<h162*>(orow + 8 * dn) = floats_to_h162(o[dn][0] * inv0, o[dn][1] * inv0);
    }
    if (ok1) {
      h16* orow = obase + ((int64_t)grp * nq + i1) * p.d + 2 * tig;
#pragma unroll
      for (int dn = 0; dn < 8; ++dn)
        *reinterpret_cast<h162*>(orow + 8 * dn) = floats_to_h162(o[dn][2] * inv1, o[dn][3] * inv1);
    }
  }
}

// SIMT: one warp per (group, head, query row); scores in shared memory (nk <= MAX_NK)
constexpr int MAX_NK = 256;
template <typename T>
__global__ void __launch_bounds__(128) group_attn_simt_kernel(const Params p) {
  __shared__ float sc_all[4][MAX_NK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t item = (int64_t)blockIdx.x * 4 + warp;
  const int64_t n_items = (int64_t)p.n_groups * p.H * p.nq;
  if (item >= n_items) return;
  const int i = (int)(item % p.nq);
  const int h = (int)((item / p.nq) % p.H);
  const int grp = (int)(item / ((int64_t)p.nq * p.H));
  float* sc = sc_all[warp];
  const int nk = p.nk;
  const T* q = static_cast<const T*>(p.q) + ((int64_t)grp * p.nq + i) * p.q_ld + h * DH;
  const T* kvb = static_cast<const T*>(p.kv) + (int64_t)grp * nk * p.kv_ld + h * DH;
  float qf[DH];
#pragma unroll
  for (int x = 0; x < DH; x += 8) {
    float t8[8];
    Act<T>::load8(q + x, t8);
#pragma unroll
    for (int y = 0; y < 8; ++y) qf[x + y] = t8[y];
  }
  float mx = -INFINITY;
  for (int k = lane; k < nk; k += 32) {
    const T* kr = kvb + (int64_t)k * p.kv_ld + p.k_col;
    float acc = 0.f;
#pragma unroll
    for (int x = 0; x < DH; x += 8) {
      float t8[8];
      Act<T>::load8(kr + x, t8);
#pragma unroll
      for (int y = 0; y < 8; ++y) acc = fmaf(qf[x + y], t8[y], acc);
    }
    float sv = acc / 8.0f;
    const bool pad = p.key_tokens != nullptr && p.key_tokens[(int64_t)grp * nk + k] == CARE_PAD;
    if (pad || (p.causal && k > i)) sv = -1e9f;
    if (p.bias != nullptr) sv += __ldg(p.bias + (int64_t)h * nk + k);
    sc[k] = sv;
    mx = fmaxf(mx, sv);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int k = lane; k < nk; k += 32) {
    const float e = expf(sc[k] - mx);
    sc[k] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float a0 = 0.f, a1 = 0.f;
  for (int k = 0; k < nk; ++k) {
    const float pr = sc[k] / sum;
    const T* vr = kvb + (int64_t)k * p.kv_ld + p.v_col;
    a0 = fmaf(pr, Act<T>::to_float(vr[lane]), a0);
    a1 = fmaf(pr, Act<T>::to_float(vr[lane + 32]), a1);
  }
  T* o = static_cast<T*>(p.out) + ((int64_t)grp * p.nq + i) * p.d + h * DH;
  o[lane] = Act<T>::from_float(a0);
  o[lane + 32] = Act<T>::from_float(a1);
}

template <int NT>
static int launch_mma(care_ctx* ctx, const CUtensorMap& tmap, const Params& p, cudaStream_t stream) {
  auto kern = group_attn_mma_kernel<NT>;
  const size_t smem = (size_t)2 * p.rows_pad * 128 + 1024 + 16;
  static size_t configured_all[64] = {0};   // per device: function attributes are per device
  size_t& configured = configured_all[ctx->device & 63];
  if (smem > configured) {
    CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  kern<<<p.n_groups * p.H, 32, smem, stream>>>(tmap, p);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace attn_group
}  // namespace care

using namespace care;

extern "C" {

int care_group_attn(care_ctx* ctx, int dtype, const void* q, int64_t ldq, const void* kv, int64_t ldkv, int k_col,
                    int v_col, int n_groups, int nq, int nk, int H, int d, const int32_t* key_tokens, int causal,
                    const float* bias, void* out, void* stream) {
  CARE_CHECK_DTYPE(dtype, "care_group_attn");
  CARE_CHECK_ARG(ctx && q && kv && out && n_groups > 0 && nq > 0 && nk > 0, "care_group_attn: bad args");
  CARE_CHECK_ARG(H > 0 && d == H * attn_group::DH, "care_group_attn: head size must be 64 (d=%d, H=%d)", d, H);
  CARE_CHECK_ARG(!causal || nq == nk, "care_group_attn: causal needs nq == nk");
  CARE_CHECK_ARG(ldq % 8 == 0 && ldkv % 8 == 0 && k_col % 8 == 0 && v_col % 8 == 0,
                 "care_group_attn: ldq, ldkv, k_col, v_col must be multiples of 8");
  attn_group::Params p{};
  p.q = q; p.q_ld = ldq; p.kv = kv; p.kv_ld = ldkv; p.k_col = k_col; p.v_col = v_col;
  p.n_groups = n_groups; p.nq = nq; p.nk = nk; p.H = H; p.d = d;
  p.rows_pad = (nk + 15) & ~15;
  p.key_tokens = key_tokens; p.causal = causal; p.bias = bias; p.out = out;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == CARE_H16 && ctx->attn_impl == 1 && nk <= 128 && (reinterpret_cast<uintptr_t>(kv) & 15) == 0) {
    CUtensorMap tmap;
    const uint64_t gdim[2] = {(uint64_t)ldkv, (uint64_t)n_groups * nk};
    const uint64_t gstr[1] = {(uint64_t)ldkv * 2};
    const uint32_t box[2] = {(uint32_t)attn_group::DH, (uint32_t)nk};
    int rc = get_tmap_bf16(ctx, kv, 2, gdim, gstr, box, &tmap);
    if (rc) return rc;
    if (nk <= 32) return attn_group::launch_mma<4>(ctx, tmap, p, s);
    return attn_group::launch_mma<16>(ctx, tmap, p, s);
  }
  CARE_CHECK_ARG(nk <= attn_group::MAX_NK, "care_group_attn: nk=%d exceeds %d", nk, attn_group::MAX_NK);
  const int64_t items = (int64_t)n_groups * H * nq;
  const int grid = (int)((items + 3) / 4);
  if (dtype == CARE_F32) attn_group::group_attn_simt_kernel<float><<<grid, 128, 0, s>>>(p);
  else if (dtype == CARE_H16) attn_group::group_attn_simt_kernel<h16><<<grid, 128, 0, s>>>(p);
  else {
    care::set_error("care_group_attn: bad dtype %d", dtype);
    return -1;
  }
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // extern "C"
