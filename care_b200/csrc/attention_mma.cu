// Decode-step attention for the bf16 mode: HBM-bound, one 4-warp CTA per (video, head).
//
// The v0 SIMT kernel (attention.cu) spent ~10k issue slots per (video, head) on per-key FFMA/shuffle
// chains and reached only ~22 % of HBM bandwidth (profiles/r01_ncu_full_v0_step15.txt).  Here:
//   * the K and V tiles of one (video, head) - 114 x 64 (cross) or t*K x 64 (self) bf16 - are staged in
//     shared memory by ONE TMA tensor copy each (SWIZZLE_128B), so the bytes in flight per SM are set
//     by resident CTAs x 2 tiles, not by registers; the four warps of the CTA split the keys so the
//     tile is held for a short compute phase only;
//   * S = Q K^T and O = P V run as warp-level mma.sync.m16n8k16 (M = the K beams padded to 16) on
//     ldmatrix fragments; this is only to cut issue slots - 0.15 MFLOP per 29 KB tile is far below
//     what would justify a tcgen05/TMEM round trip - and leaves HBM as the bound;
//   * scale, PAD/ancestry mask (self) or per-head hybrid bias (cross), softmax: fp32 in registers,
//     statistics through quad shuffles; P is fed to the second MMA as a bf16 hi+lo pair (~16 bit).
// Semantics are those of attention.cu (Attention.py:81-129, Transformer.py:15-47,169-174).
#include "dev_util.cuh"
#include "step_prologue.cuh"

namespace care {
namespace attn_mma {

using namespace care::dev;

constexpr int DH = 64;

// D += A * B with A rows 8..15 all zero (only beams 0..7 exist): a1 = a3 = 0.
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  mma_m16n8k16(c, a0, 0u, a2, 0u, b0, b1);
}

struct Params {
  const h16* q;   // q of (row 0, head 0)
  int64_t q_ld;
  int k_col, v_col;         // element column of K / V (head 0) in the tensor map's row
  int n_keys;               // Lm (cross) or n_pos * K (self)
  int rows_pad;             // n_keys rounded up to 16
  int K, H, d;
  int Lm;                   // cross
  int n_pos;                // self
  const float* bias;        // cross: [H, Lm] or NULL
  const uint8_t* anc;       // self: [B, K, anc_stride]
  int anc_stride;
  const int32_t* tok_hist;  // self: [B, T_max+1, K]
  int tok_stride;
  const int32_t* done;
  unsigned long long* row_counter;   // self: statistics, K/V cache rows read per video (head 0 counts)
  h16* out;       // [R, d]
  int n_items;              // B * H
  int l2_hints;             // ctx->l2_hints (bit 1: K/V tiles and cache rows are loaded evict_first)
};



constexpr int COMB_LD = 72;   // floats per row of the merge buffer (64 + pad: <= 2-way bank conflicts)

// S = Q K^T, softmax and O = P V over the staged tiles (shared by the TMA kernel and the compacting
// self-attention kernel).  bar_k / bar_v: mbarriers to wait on, or 0 when the tiles are already in place.
template <int KKW, bool SELF, int WARPS>
__device__ __forceinline__ void attend(const Params& p, int v, int h, int warp, int lane, uint32_t (&qa)[4][2],
                                       uint32_t k_s, uint32_t v_s, int n_keys, int rows_pad, uint32_t bar_k,
                                       uint32_t bar_v, const uint32_t* mw, float* stat, float* comb,
                                       const float (&bq)[2 * KKW][2]) {
  const int K = p.K;
  const int g = lane >> 2, tig = lane & 3;
  if (g >= K) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) qa[ks][0] = qa[ks][1] = 0u;
  }
  // this warp's 16-key steps
  const int n_kk = rows_pad >> 4;
  const int kk0 = (warp * n_kk) / WARPS, kk1 = ((warp + 1) * n_kk) / WARPS;
  const int m = lane >> 3, rr = lane & 7;

  // ---- S = Q K^T over the warp's keys ----------------------------------------------------------------
  float s[2 * KKW][2];
  if (bar_k) mbar_wait(bar_k, 0);
#pragma unroll
  for (int i = 0; i < KKW; ++i) {
    const int kk = kk0 + i;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      if (kk < kk1) {
        const int r = kk * 16 + half * 8 + rr;
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t b[4];
          ldsm_x4(b, k_s + sw128(r, 4 * kp + m));
          mma_bf16(c, qa[2 * kp][0], qa[2 * kp][1], b[0], b[1]);
          mma_bf16(c, qa[2 * kp + 1][0], qa[2 * kp + 1][1], b[2], b[3]);
        }
      }
      s[2 * i + half][0] = c[0];
      s[2 * i + half][1] = c[1];
    }
  }
  // ---- scale, mask / bias, local softmax (fp32) ---------------------------------------------------
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < KKW; ++i) {
    const int kk = kk0 + i;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int j = kk * 16 + half * 8 + 2 * tig + e;
        float x = s[2 * i + half][e] * 0.125f;   // / sqrt(64), Attention.py:84
        if (SELF) {
          if (!((mw[g * 8 + ((j >> 5) & 7)] >> (j & 31)) & 1u)) x = -1e9f;
        } else {
          x += bq[2 * i + half][e];   // hybrid attention bias of key j, fetched while the K tile was in flight
        }
        if (kk >= kk1 || j >= n_keys) x = -INFINITY;
        s[2 * i + half][e] = x;
        mx = fmaxf(mx, x);
      }
    }
  }
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  const float mref = mx == -INFINITY ? 0.f : mx;   // a warp without keys contributes nothing
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < 2 * KKW; ++i) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float pe = __expf(s[i][e] - mref);
      s[i][e] = pe;
      sum += pe;
    }
  }
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);
  sum += __shfl_xor_sync(0xffffffffu, sum, 2);

  // ---- partial O = P V -----------------------------------------------------------------------------
  float o[8][4];
#pragma unroll
  for (int dn = 0; dn < 8; ++dn)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[dn][e] = 0.f;
  if (bar_v) mbar_wait(bar_v, 0);
#pragma unroll
  for (int i = 0; i < KKW; ++i) {
    const int kk = kk0 + i;
    if (kk < kk1) {
      const float p0 = s[2 * i][0], p1 = s[2 * i][1], p2 = s[2 * i + 1][0], p3 = s[2 * i + 1][1];
      const uint32_t a0h = pack_h16(p0, p1), a2h = pack_h16(p2, p3);
      const h162 h0 = *reinterpret_cast<const h162*>(&a0h);
      const h162 h2 = *reinterpret_cast<const h162*>(&a2h);
      const uint32_t a0l = pack_h16(p0 - __low2float(h0), p1 - __high2float(h0));
      const uint32_t a2l = pack_h16(p2 - __low2float(h2), p3 - __high2float(h2));
      const int r = kk * 16 + 8 * (m & 1) + rr;
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b[4];
        ldsm_x4_trans(b, v_s + sw128(r, 2 * dp + (m >> 1)));
        mma_bf16(o[2 * dp], a0h, a2h, b[0], b[1]);
        mma_bf16(o[2 * dp], a0l, a2l, b[0], b[1]);
        mma_bf16(o[2 * dp + 1], a0h, a2h, b[2], b[3]);
        mma_bf16(o[2 * dp + 1], a0l, a2l, b[2], b[3]);
      }
    }
  }
  if (WARPS == 1) {   // short key sets: one warp saw every key, no merge
    if (g < K) {
      const float inv = 1.0f / sum;
      h16* orow = p.out + (int64_t)(v * K + g) * p.d + h * DH + 2 * tig;
#pragma unroll
      for (int dn = 0; dn < 8; ++dn)
        *reinterpret_cast<h162*>(orow + 8 * dn) = floats_to_h162(o[dn][0] * inv, o[dn][1] * inv);
    }
    return;
  }
  // ---- merge the partials of the warps (the merge buffer aliases the K tile) --------------------------
  __syncthreads();
  {
    float* crow = comb + ((size_t)warp * 8 + g) * COMB_LD + 2 * tig;
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) *reinterpret_cast<float2*>(crow + 8 * dn) = make_float2(o[dn][0], o[dn][1]);
    if (tig == 0) {
      stat[(warp * 8 + g) * 2 + 0] = mx;
      stat[(warp * 8 + g) * 2 + 1] = sum;
    }
  }
  __syncthreads();
  {
    const int col = threadIdx.x & 63;
    for (int b = threadIdx.x >> 6; b < K; b += WARPS / 2) {
      float M = -INFINITY;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) M = fmaxf(M, stat[(w * 8 + b) * 2]);
      float num = 0.f, den = 0.f;
#pragma unroll
      for (int w = 0; w < WARPS; ++w) {
        const float mwv = stat[(w * 8 + b) * 2];
        const float sc = mwv == -INFINITY ? 0.f : __expf(mwv - M);
        den += stat[(w * 8 + b) * 2 + 1] * sc;
        num += comb[((size_t)w * 8 + b) * COMB_LD + col] * sc;
      }
      p.out[(int64_t)(v * K + b) * p.d + h * DH + col] = float_to_h16(num / den);
    }
  }
}

// One CTA of 4 warps per (video, head).  The keys are split over the warps in steps of 16 (flash-decoding
// inside the CTA): each warp computes S, a local softmax and a partial O for its key range, the partials
// are merged through shared memory.  KKW = 16-key steps per warp the registers are sized for
// (n_keys <= 64 * KKW).

template <int KKW, bool SELF, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
attn_mma_kernel(const __grid_constant__ CUtensorMap tmap, const Params p) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x / p.H, h = blockIdx.x - v * p.H;
  if (p.done != nullptr && p.done[v]) return;   // uniform for the CTA
  const int K = p.K, n_keys = p.n_keys;
  if (SELF && h == 0 && threadIdx.x == 0 && p.row_counter != nullptr) atomicAdd(p.row_counter, (unsigned long long)n_keys);

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t tile_bytes = (uint32_t)p.rows_pad * 128u;
  const uint32_t k_s = base, v_s = base + tile_bytes;
  const uint32_t aux = base + 2u * tile_bytes;
  const uint32_t bar_k = aux, bar_v = aux + 8u;
  uint8_t* aux_gen = smem_raw + (aux - raw);
  uint32_t* mw = reinterpret_cast<uint32_t*>(aux_gen + 16);                 // [8 beams][8 words] key masks
  float* stat = reinterpret_cast<float*>(aux_gen + 16 + 256);               // [WARPS][8][2] (max, sum)
  // [WARPS][8][COMB_LD] partial O: reuses the K tile (dead once every warp has its scores) when it is large
  // enough, else it sits behind the auxiliary block (short key sets, e.g. the 30 concept embeddings)
  constexpr uint32_t kCombBytes = WARPS * 8 * COMB_LD * 4;
  float* comb = tile_bytes >= kCombBytes ? reinterpret_cast<float*>(smem_raw + (k_s - raw))
                                         : reinterpret_cast<float*>(aux_gen + 16 + 256 + 256);

  // V rows [n_keys, rows_pad) are multiplied by P == 0: they must hold finite values
  {
    uint8_t* v_gen = smem_raw + (v_s - raw);
    const int n16 = (p.rows_pad - n_keys) * 8;
    for (int i = threadIdx.x; i < n16; i += WARPS * 32)
      *reinterpret_cast<uint4*>(v_gen + (size_t)n_keys * 128 + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  if (threadIdx.x == 0) {
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t bytes = (uint32_t)n_keys * 128u;
    mbar_expect_tx(bar_k, bytes);
    const uint64_t pol_kv = l2_policy((p.l2_hints & 2) ? 2 : 0);
    if (SELF) tma_load_3d(k_s, &tmap, bar_k, p.k_col + h * DH, v * K, 0);
    else tma_load_2d_hint(k_s, &tmap, bar_k, p.k_col + h * DH, v * p.Lm, pol_kv);
    mbar_expect_tx(bar_v, bytes);
    if (SELF) tma_load_3d(v_s, &tmap, bar_v, p.v_col + h * DH, v * K, 0);
    else tma_load_2d_hint(v_s, &tmap, bar_v, p.v_col + h * DH, v * p.Lm, pol_kv);
  }

  const int g = lane >> 2, tig = lane & 3;   // g = beam (MMA row), tig = column pair
  // Q fragments: A[beam g][dims 16ks + 2tig, +1] and [.. + 8, + 9]
  uint32_t qa[4][2];
  {
    const h16* qrow = p.q + (int64_t)(v * K + (g < K ? g : 0)) * p.q_ld + h * DH;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * tig);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * tig + 8);
    }
  }
  // self: bit j of beam b's mask = key j (position j / K, slot j % K) is on b's prefix and not <pad>
  if (SELF) {
    for (int i = threadIdx.x; i < 64; i += WARPS * 32) mw[i] = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < K * p.n_pos; i += WARPS * 32) {
      const int b = i / p.n_pos, pp = i - b * p.n_pos;
      const int slot = (pp == p.n_pos - 1) ? b : (int)p.anc[((int64_t)v * K + b) * p.anc_stride + pp];
      const int tok = p.tok_hist[(int64_t)v * p.tok_stride + pp * K + slot];
      if (tok != CARE_PAD) {
        const int j = pp * K + slot;
        atomicOr(&mw[b * 8 + (j >> 5)], 1u << (j & 31));
      }
    }
  }
  // cross: the hybrid attention bias of this thread's keys (Attention.py:104-111), loaded now so that its latency
  // hides behind the K tile's instead of sitting between the two MMA phases
  float bq[2 * KKW][2];
  {
    const int n_kk = p.rows_pad >> 4;
    const int kk0 = (warp * n_kk) / WARPS;
#pragma unroll
    for (int i = 0; i < KKW; ++i)
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int j = (kk0 + i) * 16 + half * 8 + 2 * tig + e;
          bq[2 * i + half][e] = (!SELF && p.bias != nullptr && j < n_keys) ? __ldg(p.bias + (int64_t)h * p.Lm + j) : 0.f;
        }
  }
  __syncthreads();   // mbarrier init + mask words visible to every warp
  attend<KKW, SELF, WARPS>(p, v, h, warp, lane, qa, k_s, v_s, n_keys, p.rows_pad, bar_k, bar_v, mw, stat, comb, bq);
}

template <int KKW, bool SELF, int WARPS>
static int launch(care_ctx* ctx, const CUtensorMap& tmap, const Params& p, cudaStream_t stream) {
  auto kern = attn_mma_kernel<KKW, SELF, WARPS>;
  const size_t tile = (size_t)p.rows_pad * 128;
  const size_t comb_bytes = WARPS > 1 ? (size_t)WARPS * 8 * COMB_LD * 4 : 0;
  // the merge buffer aliases the K tile when that is large enough, else it gets its own space
  const size_t smem = 2 * tile + 1024 + 16 + 256 + 256 + (tile >= comb_bytes ? 0 : comb_bytes);
  static size_t configured_all[64] = {0};   // per device: function attributes are per device
  size_t& configured = configured_all[ctx->device & 63];
  if (smem > configured) {
    CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  CARE_CUDA(launch_pdl(ctx, kern, dim3(p.n_items), dim3(WARPS * 32), smem, stream, tmap, p));
  if (SELF) ctx->last_self_attn = "attn_mma_kernel<self>";
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Self-attention with slot compaction.  The KV cache keeps K slots per position, but the K beams of a
// video share most of their ancestry: on average only ~2.5 of 5 slots per position are on some beam's
// prefix with the benchmark weights, ~1.2 with peaked distributions (scripts/ancestry_stats.py).  This
// kernel derives the live (position, slot) set from the ancestry table, gathers ONLY those K/V rows
// into the swizzled shared-memory tiles with 16-byte cp.async copies, and runs the same MMA pipeline on
// the compacted key list - HBM traffic and MMA work both shrink by K / (live slots per position).
// ---------------------------------------------------------------------------------------------
constexpr int MAX_POS = 64;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__global__ void __launch_bounds__(128)
compact_info_kernel(const uint8_t* __restrict__ anc, int anc_stride, const int32_t* __restrict__ tok_hist,
                    int tok_stride, const int32_t* __restrict__ done, int B, int K, int n_pos,
                    uint32_t* __restrict__ info) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ uint32_t rec_all[4][INFO_WORDS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x * 4 + warp;
  if (v >= B) return;
  if (done != nullptr && done[v]) return;
  warp_compact_record(anc, anc_stride, tok_hist, tok_stride, v, K, n_pos, lane, rec_all[warp], info);
}

template <int KKW>
__global__ void __launch_bounds__(128)
attn_self_compact_kernel(const Params p, const h16* __restrict__ cache, int64_t R,
                         unsigned long long* __restrict__ row_counter, const uint32_t* __restrict__ info) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int WARPS = 4;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x / p.H, h = blockIdx.x - v * p.H;
  if (p.done != nullptr && p.done[v]) return;
  const int K = p.K, n_pos = p.n_pos;

  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t tile_bytes = (uint32_t)p.rows_pad * 128u;     // sized for all n_pos * K rows
  const uint32_t k_s = base, v_s = base + tile_bytes;
  const uint32_t aux = base + 2u * tile_bytes;
  uint8_t* aux_gen = smem_raw + (aux - raw);
  uint32_t* mw = reinterpret_cast<uint32_t*>(aux_gen + 16);
  float* stat = reinterpret_cast<float*>(aux_gen + 16 + 256);
  float* comb = reinterpret_cast<float*>(aux_gen + 16 + 256 + 256);
  uint8_t* sa = aux_gen + 16 + 256 + 256 + WARPS * 8 * COMB_LD * 4;          // [8][MAX_POS] slot of (beam, pos)
  uint8_t* live = sa + 8 * MAX_POS;                                           // [MAX_POS] slot bitmask per position
  uint16_t* off = reinterpret_cast<uint16_t*>(live + MAX_POS);                // [MAX_POS + 1] first row of a position
  uint16_t* rowsrc = off + MAX_POS + 2;                                       // [rows] cache row (pos << 4 | slot)
  __shared__ int n_live_s;

  // Q fragments first: their latency overlaps the ancestry bookkeeping
  const int g = lane >> 2, tig = lane & 3;
  uint32_t qa[4][2];
  {
    const h16* qrow = p.q + (int64_t)(v * K + (g < K ? g : 0)) * p.q_ld + h * DH;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * tig);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(qrow + 16 * ks + 2 * tig + 8);
    }
  }
  int n_live, rows_pad;
  if (info != nullptr) {
    // the step's per-video record (compact_info_kernel): one coalesced read instead of the derivation below
    const uint32_t* rec = info + (int64_t)v * INFO_WORDS;
    uint32_t* rs32 = reinterpret_cast<uint32_t*>(rowsrc);
    for (int i = threadIdx.x; i < INFO_WORDS; i += 128) {
      const uint32_t w = __ldg(rec + i);
      if (i == 0) n_live_s = (int)w;
      else if (i <= 64) mw[i - 1] = w;
      else rs32[i - 65] = w;
    }
    __syncthreads();
    n_live = n_live_s;
    rows_pad = (n_live + 15) & ~15;
    if (threadIdx.x == 0 && h == 0 && row_counter != nullptr) atomicAdd(row_counter, (unsigned long long)n_live);
  } else {
    for (int i = threadIdx.x; i < 64; i += 128) mw[i] = 0u;
    for (int i = threadIdx.x; i < K * n_pos; i += 128) {
      const int b = i / n_pos, pp = i - b * n_pos;
      sa[b * MAX_POS + pp] = (pp == n_pos - 1) ? (uint8_t)b : p.anc[((int64_t)v * K + b) * p.anc_stride + pp];
    }
    __syncthreads();
    if (threadIdx.x < n_pos) {
      uint32_t bits = 0u;
      for (int b = 0; b < K; ++b) bits |= 1u << sa[b * MAX_POS + threadIdx.x];
      live[threadIdx.x] = (uint8_t)bits;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int pp = 0; pp < n_pos; ++pp) {
        off[pp] = (uint16_t)acc;
        acc += __popc((uint32_t)live[pp]);
      }
      off[n_pos] = (uint16_t)acc;
      n_live_s = acc;
      if (h == 0 && row_counter != nullptr) atomicAdd(row_counter, (unsigned long long)acc);
    }
    __syncthreads();
    n_live = n_live_s;
    rows_pad = (n_live + 15) & ~15;
    // row table + per-beam key masks in the compacted index space
    for (int i = threadIdx.x; i < K * n_pos; i += 128) {
      const int s = i / n_pos, pp = i - s * n_pos;        // here: candidate slot s of position pp
      const uint32_t bits = live[pp];
      if ((bits >> s) & 1u) rowsrc[off[pp] + __popc(bits & ((1u << s) - 1u))] = (uint16_t)((pp << 4) | s);
      // beam b == s of this loop index: its key at position pp
      const int b = s;
      const int slot = sa[b * MAX_POS + pp];
      const int tok = p.tok_hist[(int64_t)v * p.tok_stride + pp * K + slot];
      if (tok != CARE_PAD) {
        const int j = off[pp] + __popc(bits & ((1u << slot) - 1u));
        atomicOr(&mw[b * 8 + (j >> 5)], 1u << (j & 31));
      }
    }
    __syncthreads();
  }
  // gather: 16 chunks of 16 B per live row (8 of K, 8 of V)
  {
    const h16* kbase = cache + (int64_t)v * K * (3LL * p.d) + p.k_col + h * DH;
    const h16* vbase = cache + (int64_t)v * K * (3LL * p.d) + p.v_col + h * DH;
    const int n_chunks = n_live * 16;
    for (int c = threadIdx.x; c < n_chunks; c += 128) {
      const int j = c >> 4, isv = (c >> 3) & 1, c16 = c & 7;
      const int src = rowsrc[j];
      const int pp = src >> 4, sl = src & 15;
      const int64_t roff = ((int64_t)pp * R + sl) * (3LL * p.d) + c16 * 8;
      cp_async16((isv ? v_s : k_s) + sw128(j, c16), (isv ? vbase : kbase) + roff);
    }
    // rows [n_live, rows_pad) of V are multiplied by P == 0: they must hold finite values
    uint8_t* v_gen = smem_raw + (v_s - raw);
    for (int i = threadIdx.x; i < (rows_pad - n_live) * 8; i += 128)
      *reinterpret_cast<uint4*>(v_gen + (size_t)n_live * 128 + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const float no_bias[2 * KKW][2] = {};
  attend<KKW, true, WARPS>(p, v, h, warp, lane, qa, k_s, v_s, n_live, rows_pad, 0u, 0u, mw, stat, comb, no_bias);
}

template <int KKW>
static int launch_compact(care_ctx* ctx, const Params& p, const void* cache, int64_t R, cudaStream_t stream) {
  auto kern = attn_self_compact_kernel<KKW>;
  const size_t smem = (size_t)2 * p.rows_pad * 128 + 1024 + 16 + 256 + 256 + (size_t)4 * 8 * COMB_LD * 4 +
                      8 * MAX_POS + MAX_POS + (MAX_POS + 2) * 2 + 384 /* rowsrc: 190 uint16 of a record */ + 64;
  static size_t configured_all[64] = {0};   // per device: function attributes are per device
  size_t& configured = configured_all[ctx->device & 63];
  if (smem > configured) {
    CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const uint32_t* info = nullptr;
  if (ctx->compact_info != nullptr && p.n_items / p.H <= ctx->compact_info_videos) {
    const int B = p.n_items / p.H;
    CARE_CUDA(launch_pdl(ctx, compact_info_kernel, dim3((B + 3) / 4), dim3(128), 0, stream, p.anc, p.anc_stride, p.tok_hist,
                         p.tok_stride, p.done, B, p.K, p.n_pos, ctx->compact_info));
    CARE_LAUNCH_CHECK(ctx);
    info = ctx->compact_info;
  }
  CARE_CUDA(launch_pdl(ctx, kern, dim3(p.n_items), dim3(128), smem, stream, p, static_cast<const h16*>(cache), R,
                       ctx->self_attn_rows, info));
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Self-attention as a stream of gathered key chunks (option self_compact = 2).
//
// The compacting kernel above keeps the one-CTA-per-(video, head) shape: shared memory is sized for the
// worst case (every slot live), so halving the bytes per CTA halves the bytes in flight per SM, and the
// per-video record adds a dependent read to every CTA's latency chain - it measured slower than the dense
// TMA kernel.  Here every WARP walks its own strided list of (video, head) items and consumes their live
// K/V rows as fixed 16-row chunks out of a private ring of stages:
//   * rows are fetched by TMA row gathers (cp.async.bulk.tensor.2d.tile::gather4: four arbitrary rows of the
//     [T*R, 3d] cache per instruction, 128 B each, SWIZZLE_128B), issued up to STAGES chunks ahead and
//     across the boundary to the next item, whose record is prefetched two items ahead - bytes in flight
//     no longer depend on the worst-case tile;
//   * softmax is computed online across the chunks of an item (running max / sum / O in registers), so one
//     warp owns an item end to end: no block barriers, no merge buffer; 16 such warps per SM hide each
//     other's instruction latency.
// MMA fragments, masks and the hi+lo split of P are those of attend() above.
// ---------------------------------------------------------------------------------------------
namespace gs {

constexpr int CH = 16;                       // rows per chunk (one MMA k-step)
constexpr int SLOTS = 3;                     // resident records: current item, next, the one after
constexpr int WARPS = 4;                     // independent warps per CTA
constexpr uint32_t CHUNK_BYTES = CH * 128;   // one K (or V) chunk
constexpr uint32_t REC_BYTES = INFO_WORDS * 4;

template <int STAGES>
struct Cfg {
  static constexpr uint32_t RING = STAGES * 2 * CHUNK_BYTES;
  static constexpr uint32_t WARP_BYTES = (RING + SLOTS * REC_BYTES + 8 * (STAGES + SLOTS) + 1023u) & ~1023u;
  static constexpr uint32_t SMEM_BYTES = WARPS * WARP_BYTES + 1024;
};

__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* map, uint32_t bar, int col, int r0, int r1,
                                            int r2, int r3) {
  // no L2 eviction hint here (the cross-attention K/V tiles carry one): the 64-bit policy operand costs the two
  // registers this 128-register kernel does not have (24 bytes of spills when it was tried)
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
      : "memory");
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int STAGES>
__global__ void __launch_bounds__(WARPS * 32, 4)
attn_self_stream_kernel(const __grid_constant__ CUtensorMap tmap, const Params p, int R, const uint32_t* __restrict__ info,
                        unsigned long long* __restrict__ row_counter, const EarlyExit ee) {
  pdl_wait();
  pdl_launch_dependents();
  if (all_done(ee)) return;
  extern __shared__ uint8_t smem_raw[];
  using C = Cfg<STAGES>;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = p.K, H = p.H;
  const int stride = gridDim.x * WARPS;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = ((raw + 1023u) & ~1023u) + warp * C::WARP_BYTES;   // this warp's private region
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t kv_s = base;                                   // stage s: K chunk, then V chunk
  const uint32_t rec_s = base + C::RING;                        // SLOTS records
  const uint32_t bar_kv = rec_s + SLOTS * REC_BYTES;            // STAGES barriers
  const uint32_t bar_item = bar_kv + 8 * STAGES;                // SLOTS barriers
  const uint32_t* rec_gen = reinterpret_cast<const uint32_t*>(gen + C::RING);

  // items of this warp: id, id + stride, ... ; (video, head) advance without divisions
  struct Item {
    int id, v, h;
  };
  const int dv = stride / H, dh = stride - dv * H;
  auto skip_done = [&](Item& x) {
    while (x.id < p.n_items && p.done != nullptr && p.done[x.v]) {
      x.id += stride;
      x.v += dv;
      x.h += dh;
      if (x.h >= H) {
        x.h -= H;
        ++x.v;
      }
    }
  };
  auto following = [&](const Item& x) {
    Item y{x.id + stride, x.v + dv, x.h + dh};
    if (y.h >= H) {
      y.h -= H;
      ++y.v;
    }
    skip_done(y);
    return y;
  };
  Item it;
  it.id = blockIdx.x * WARPS + warp;
  it.v = it.id / H;
  it.h = it.id - it.v * H;
  skip_done(it);
  if (it.id >= p.n_items) return;

  // V rows behind the last gathered group are multiplied by P == 0: every stage starts out finite
  for (uint32_t i = lane; i < C::RING / 16; i += 32)
    *reinterpret_cast<uint4*>(gen + (size_t)i * 16) = make_uint4(0u, 0u, 0u, 0u);
  if (lane == 0) {
    for (int i = 0; i < STAGES + SLOTS; ++i) mbar_init(bar_kv + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the zero fill precedes the TMA writes
  __syncwarp();

  const int g = lane >> 2, tig = lane & 3, m = lane >> 3, rr = lane & 7;
  // ldmatrix addresses inside a chunk (SWIZZLE_128B): K row (half * 8 + rr), 16-byte column (4 kp + m);
  // V row (8 (m & 1) + rr), column (2 dp + (m >> 1)).  kp / half / dp only flip address bits.
  const uint32_t k_lane = (uint32_t)(rr * 128 + ((m ^ rr) << 4));
  const uint32_t v_lane = (uint32_t)((8 * (m & 1) + rr) * 128 + (((m >> 1) ^ rr) << 4));
  auto prefetch_record = [&](const Item& x, int slot) {
    if (lane == 0) {
      const uint32_t bar = bar_item + 8 * slot;
      mbar_expect_tx(bar, REC_BYTES);
      bulk_load(rec_s + slot * REC_BYTES, info + (int64_t)x.v * INFO_WORDS, REC_BYTES, bar);
    }
  };
  // q fragments of beam g for the item, straight from the cache rows of the newest position
  auto load_q = [&](const Item& x, uint32_t (&qf)[4][2]) {
    const uint32_t* qrow =
        reinterpret_cast<const uint32_t*>(p.q + (int64_t)(x.v * K + (g < K ? g : 0)) * p.q_ld + x.h * DH) + tig;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qf[ks][0] = g < K ? __ldg(qrow + 8 * ks) : 0u;
      qf[ks][1] = g < K ? __ldg(qrow + 8 * ks + 4) : 0u;
    }
  };
  // chunk c of the item (its record sits in `slot`) into ring stage `stage`: lane gq gathers rows 4gq .. 4gq+3
  auto issue_chunk = [&](const Item& x, int slot, int c, int stage) {
    const uint32_t* rec = rec_gen + slot * INFO_WORDS;
    const int n_live = (int)rec[0];
    const int groups = (min(CH, n_live - c * CH) + 3) >> 2;
    const uint16_t* rowsrc = reinterpret_cast<const uint16_t*>(rec + 65);
    const int src = rowsrc[min(c * CH + (lane & 15), n_live - 1)];
    const int row = (src >> 4) * R + x.v * K + (src & 15);
    const int r0 = __shfl_sync(0xffffffffu, row, tig * 4), r1 = __shfl_sync(0xffffffffu, row, tig * 4 + 1);
    const int r2 = __shfl_sync(0xffffffffu, row, tig * 4 + 2), r3 = __shfl_sync(0xffffffffu, row, tig * 4 + 3);
    const uint32_t bar = bar_kv + 8 * stage;
    const uint32_t k_dst = kv_s + stage * 2 * CHUNK_BYTES + lane * 512;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)groups * 1024u);
    __syncwarp();
    if (lane < groups) {
      tma_gather4(k_dst, &tmap, bar, p.k_col + x.h * DH, r0, r1, r2, r3);
      tma_gather4(k_dst + CHUNK_BYTES, &tmap, bar, p.v_col + x.h * DH, r0, r1, r2, r3);
    }
  };

  uint32_t item_seq = 0, issued = 0, consumed = 0;
  bool p_on_next = false;   // the producer has moved on to the next item
  int p_chunk = 0;          // next chunk the producer issues (of the current or of the next item)
  int nc_next = 0;
  prefetch_record(it, 0);
  Item nxt = following(it);
  if (nxt.id < p.n_items) prefetch_record(nxt, 1);
  uint32_t qa[4][2], qn[4][2];
  load_q(it, qa);

  while (true) {
    const int slot = item_seq % SLOTS, slot1 = (item_seq + 1) % SLOTS;
    const bool has_next = nxt.id < p.n_items;
    // the record of the item after next is requested a whole item ahead of the producer needing it
    Item nxt2 = nxt;
    if (has_next) {
      nxt2 = following(nxt);
      if (nxt2.id < p.n_items) prefetch_record(nxt2, (item_seq + 2) % SLOTS);
      load_q(nxt, qn);
    }
    mbar_wait(bar_item + 8 * slot, (item_seq / SLOTS) & 1);
    const uint32_t* rec = rec_gen + slot * INFO_WORDS;
    const int n_live = (int)rec[0];
    const int nc = (n_live + CH - 1) / CH;
    if (lane == 0 && it.h == 0 && row_counter != nullptr) atomicAdd(row_counter, (unsigned long long)n_live);

    auto issue_more = [&]() {
      while (issued - consumed < (uint32_t)STAGES) {
        if (!p_on_next) {
          if (p_chunk < nc) {
            issue_chunk(it, slot, p_chunk, issued % STAGES);
            ++issued;
            ++p_chunk;
            continue;
          }
          if (!has_next) break;
          mbar_wait(bar_item + 8 * slot1, ((item_seq + 1) / SLOTS) & 1);
          p_on_next = true;
          p_chunk = 0;
          nc_next = ((int)rec_gen[slot1 * INFO_WORDS] + CH - 1) / CH;
        }
        if (p_chunk >= nc_next) break;
        issue_chunk(nxt, slot1, p_chunk, issued % STAGES);
        ++issued;
        ++p_chunk;
      }
    };
    issue_more();

    float m_run = -INFINITY, l_run = 0.f;
    float o[8][4];
#pragma unroll
    for (int dn = 0; dn < 8; ++dn)
#pragma unroll
      for (int e = 0; e < 4; ++e) o[dn][e] = 0.f;
    const uint32_t* mask_words = rec + 1 + (g & 7) * 8;

    for (int c = 0; c < nc; ++c) {
      const int stage = consumed % STAGES;
      const uint32_t k_s = kv_s + stage * 2 * CHUNK_BYTES + k_lane;
      const uint32_t v_s = kv_s + stage * 2 * CHUNK_BYTES + CHUNK_BYTES + v_lane;
      mbar_wait(bar_kv + 8 * stage, (consumed / STAGES) & 1);
      // ---- S = Q K^T over the chunk's 16 keys ----
      float s[2][2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float cc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kp = 0; kp < 2; ++kp) {
          uint32_t b[4];
          ldsm_x4(b, (k_s + half * 1024) ^ (kp << 6));
          mma_bf16(cc, qa[2 * kp][0], qa[2 * kp][1], b[0], b[1]);
          mma_bf16(cc, qa[2 * kp + 1][0], qa[2 * kp + 1][1], b[2], b[3]);
        }
        s[half][0] = cc[0];
        s[half][1] = cc[1];
      }
      // ---- mask (bit j of beam g: compacted key j is on g's prefix and not <pad>), online softmax ----
      const uint32_t bits = mask_words[c >> 1] >> ((c & 1) * 16 + 2 * tig);
      const float x0 = (bits & 1u) ? s[0][0] * 0.125f : -INFINITY, x1 = (bits & 2u) ? s[0][1] * 0.125f : -INFINITY;
      const float x2 = (bits & 0x100u) ? s[1][0] * 0.125f : -INFINITY;
      const float x3 = (bits & 0x200u) ? s[1][1] * 0.125f : -INFINITY;
      float mx = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);
      const float mref = m_new == -INFINITY ? 0.f : m_new;
      const float scale = __expf(m_run - mref);   // 0 for the first chunk with a visible key
      const float p0 = __expf(x0 - mref), p1 = __expf(x1 - mref);
      const float p2 = __expf(x2 - mref), p3 = __expf(x3 - mref);
      float sum = (p0 + p1) + (p2 + p3);
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      sum += __shfl_xor_sync(0xffffffffu, sum, 2);
      l_run = l_run * scale + sum;
      m_run = m_new;
#pragma unroll
      for (int dn = 0; dn < 8; ++dn) {
        o[dn][0] *= scale;
        o[dn][1] *= scale;
      }
      // ---- O += P V ----
      {
        const uint32_t a0h = pack_h16(p0, p1), a2h = pack_h16(p2, p3);
        const h162 h0 = *reinterpret_cast<const h162*>(&a0h);
        const h162 h2 = *reinterpret_cast<const h162*>(&a2h);
        const uint32_t a0l = pack_h16(p0 - __low2float(h0), p1 - __high2float(h0));
        const uint32_t a2l = pack_h16(p2 - __low2float(h2), p3 - __high2float(h2));
#pragma unroll
        for (int dp = 0; dp < 4; ++dp) {
          uint32_t b[4];
          ldsm_x4_trans(b, v_s ^ (dp << 5));
          mma_bf16(o[2 * dp], a0h, a2h, b[0], b[1]);
          mma_bf16(o[2 * dp], a0l, a2l, b[0], b[1]);
          mma_bf16(o[2 * dp + 1], a0h, a2h, b[2], b[3]);
          mma_bf16(o[2 * dp + 1], a0l, a2l, b[2], b[3]);
        }
      }
      ++consumed;
      __syncwarp();   // every lane is done with the stage before it is refilled
      issue_more();
    }
    if (g < K) {
      const float inv = 1.0f / l_run;
      h16* orow = p.out + (int64_t)(it.v * K + g) * p.d + it.h * DH + 2 * tig;
#pragma unroll
      for (int dn = 0; dn < 8; ++dn)
        *reinterpret_cast<h162*>(orow + 8 * dn) = floats_to_h162(o[dn][0] * inv, o[dn][1] * inv);
    }
    if (!has_next) break;
    __syncwarp();   // the record slot of this item is free for the item three ahead
    ++item_seq;
    it = nxt;
    nxt = nxt2;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = qn[ks][0];
      qa[ks][1] = qn[ks][1];
    }
    if (p_on_next) {
      p_on_next = false;   // the chunks issued ahead now belong to the current item
    } else {
      p_chunk = 0;
    }
  }
}

template <int STAGES>
static int launch_stream_t(care_ctx* ctx, const CUtensorMap& tmap, const Params& p, int64_t R, cudaStream_t stream) {
  auto kern = attn_self_stream_kernel<STAGES>;
  static int ctas_per_sm_all[64] = {0};   // per device
  int& ctas_per_sm = ctas_per_sm_all[ctx->device & 63];
  if (ctas_per_sm == 0) {
    CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<STAGES>::SMEM_BYTES));
    int n = 0;
    CARE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, WARPS * 32, Cfg<STAGES>::SMEM_BYTES));
    ctas_per_sm = std::max(n, 1);
  }
  const int grid = std::min((p.n_items + WARPS - 1) / WARPS, ctx->sm_count * ctas_per_sm);
  CARE_CUDA(launch_pdl(ctx, kern, dim3(grid), dim3(WARPS * 32), Cfg<STAGES>::SMEM_BYTES, stream, tmap, p, (int)R,
                       (const uint32_t*)ctx->compact_info, ctx->self_attn_rows, early_exit_of(ctx)));
  ctx->last_self_attn = "attn_self_stream_kernel";
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

static int launch_stream(care_ctx* ctx, const Params& p, const void* cache, int64_t R, int T_rows, cudaStream_t stream) {
  const int B = p.n_items / p.H;
  CUtensorMap tmap;
  const uint64_t gdim[2] = {(uint64_t)3 * p.d, (uint64_t)T_rows * (uint64_t)R};
  const uint64_t gstr[1] = {(uint64_t)3 * p.d * 2};
  const uint32_t box[2] = {(uint32_t)DH, 1u};
  int rc = get_tmap_bf16(ctx, cache, 2, gdim, gstr, box, &tmap);
  if (rc) return rc;
  if (ctx->info_ready_npos == p.n_pos && ctx->info_ready_B == B && ctx->info_ready_anc == (const void*)p.anc) {
    ctx->info_ready_npos = -1;   // the beam kernel of the previous step already wrote this step's records
  } else {
    CARE_CUDA(launch_pdl(ctx, compact_info_kernel, dim3((B + 3) / 4), dim3(128), 0, stream, p.anc, p.anc_stride, p.tok_hist,
                         p.tok_stride, p.done, B, p.K, p.n_pos, ctx->compact_info));
    CARE_LAUNCH_CHECK(ctx);
  }
  static const int stages = [] {
    const char* e = getenv("CARE_B200_STREAM_STAGES");
    return e != nullptr ? atoi(e) : 2;
  }();
  if (stages == 3) return launch_stream_t<3>(ctx, tmap, p, R, stream);
  if (stages == 4) return launch_stream_t<4>(ctx, tmap, p, R, stream);
  return launch_stream_t<2>(ctx, tmap, p, R, stream);
}

}  // namespace gs

// returns 1 if the shape is not covered (caller falls back to the SIMT kernel), 0 on launch, else error
int cross_step(care_ctx* ctx, const void* q, int64_t ldq, const void* kv, int Lm, int B, int K, int H, int d,
               const float* hybrid_bias, const int32_t* done, void* ctx_out, cudaStream_t stream) {
  if (K > 8 || Lm > 128 || (ldq % 2) != 0 || (reinterpret_cast<uintptr_t>(q) & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(kv) & 15) != 0 || (reinterpret_cast<uintptr_t>(ctx_out) & 3) != 0)
    return 1;
  CUtensorMap tmap;
  const uint64_t gdim[2] = {(uint64_t)2 * d, (uint64_t)B * Lm};
  const uint64_t gstr[1] = {(uint64_t)2 * d * 2};
  const uint32_t box[2] = {(uint32_t)DH, (uint32_t)Lm};
  int rc = get_tmap_bf16(ctx, kv, 2, gdim, gstr, box, &tmap);
  if (rc) return rc;
  Params p{};
  p.q = static_cast<const h16*>(q);
  p.q_ld = ldq;
  p.k_col = 0;
  p.v_col = d;
  p.n_keys = Lm;
  p.rows_pad = (Lm + 15) & ~15;
  p.K = K; p.H = H; p.d = d; p.Lm = Lm;
  p.bias = hybrid_bias;
  p.done = done;
  p.out = static_cast<h16*>(ctx_out);
  p.n_items = B * H;
  p.l2_hints = ctx->l2_hints;
  if (Lm <= 32) return launch<2, false, 1>(ctx, tmap, p, stream);   // e.g. the 30 concept embeddings (attr_attention)
  if (Lm <= 64) return launch<2, false, 2>(ctx, tmap, p, stream);
  return launch<2, false, 4>(ctx, tmap, p, stream);   // Lm <= 128: 8 steps of 16 keys over 4 warps
}

int self_step(care_ctx* ctx, const void* cache, int n_pos, int B, int K, int H, int d, const uint8_t* anc,
              int anc_stride, const int32_t* tok_hist, const int32_t* done, void* ctx_out, cudaStream_t stream) {
  const int n_keys = n_pos * K;
  if (K > 8 || n_keys > 160 || (reinterpret_cast<uintptr_t>(cache) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(ctx_out) & 3) != 0)
    return 1;
  const int64_t R = (int64_t)B * K;
  CUtensorMap tmap;
  const uint64_t gdim[3] = {(uint64_t)3 * d, (uint64_t)R, (uint64_t)n_pos};
  const uint64_t gstr[2] = {(uint64_t)3 * d * 2, (uint64_t)R * 3 * d * 2};
  const uint32_t box[3] = {(uint32_t)DH, (uint32_t)K, (uint32_t)n_pos};
  int rc = get_tmap_bf16(ctx, cache, 3, gdim, gstr, box, &tmap);
  if (rc) return rc;
  Params p{};
  p.q = static_cast<const h16*>(cache) + (int64_t)(n_pos - 1) * R * 3 * d;
  p.q_ld = 3LL * d;
  p.k_col = d;
  p.v_col = 2 * d;
  p.n_keys = n_keys;
  p.rows_pad = (n_keys + 15) & ~15;
  p.K = K; p.H = H; p.d = d;
  p.n_pos = n_pos;
  p.anc = anc;
  p.anc_stride = anc_stride;
  p.tok_hist = tok_hist;
  p.tok_stride = (anc_stride + 1) * K;
  p.done = done;
  p.row_counter = ctx->self_attn_rows;
  p.out = static_cast<h16*>(ctx_out);
  p.n_items = B * H;
  p.l2_hints = ctx->l2_hints;
  // short prefixes: nearly every slot is still live and the dense TMA tile is cheaper than the per-item bookkeeping
  // (and few (video, head) items - latency mode - leave most of the stream kernel's warps without work)
  // (3 = the stream kernel for every shape: tests)
  if ((ctx->self_compact == 3 || (ctx->self_compact == 2 && n_pos >= 6 && p.n_items >= 1024)) &&
      ctx->compact_info != nullptr && B <= ctx->compact_info_videos &&
      (int64_t)anc_stride * R < (1LL << 31))
    return gs::launch_stream(ctx, p, cache, R, anc_stride, stream);   // rows of the cache as a 2D [T * R, 3d] tensor
  if (ctx->self_compact == 1 && n_pos >= 8 && n_pos <= MAX_POS) {
    // long enough prefixes: gather only the cache slots some beam still references
    if (n_keys <= 128) return launch_compact<2>(ctx, p, cache, R, stream);
    return launch_compact<3>(ctx, p, cache, R, stream);
  }
  if (n_keys <= 16) return launch<1, true, 1>(ctx, tmap, p, stream);
  if (n_keys <= 32) return launch<2, true, 1>(ctx, tmap, p, stream);   // short prefixes: one warp, no merge
  if (n_keys <= 64) return launch<2, true, 2>(ctx, tmap, p, stream);
  if (n_keys <= 96) return launch<3, true, 2>(ctx, tmap, p, stream);    // <= 6 steps of 16 keys: two warps balance better
  if (n_keys <= 128) return launch<2, true, 4>(ctx, tmap, p, stream);
  return launch<3, true, 4>(ctx, tmap, p, stream);   // <= 160 keys: 10 steps of 16 over 4 warps
}

}  // namespace attn_mma
}  // namespace care
