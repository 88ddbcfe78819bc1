// Device-resident beam search (replaces misc/Decoding/Beam.py's per-video Python objects and their
// per-element device syncs).  One CTA per video and step:
//   kernel 1 (one CTA per logits row): single pass, online max / sum-exp (log_softmax statistics,
//     Translator.py:127) and a register top-(K+1) of the raw logits -> warp shuffle merge -> block merge;
//   kernel 2 (one warp per video): candidate value = ((x - max) - log(sum)) + score (Beam.py:51), rows ending
//     in <eos> := -1e20 (:52-54), merge by (value desc, flat index asc), Beam.advance's bookkeeping (:61-85)
//     and the KV-cache ancestry table.
// HBM sees each logit exactly once.
#include <cfloat>
#include <climits>

#include "step_prologue.cuh"

namespace care {
namespace beam {

constexpr int THREADS = 512;
constexpr int WARPS = THREADS / 32;

__device__ __forceinline__ bool better(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

template <int KB>
struct TopList {
  float v[KB];
  int i[KB];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      v[q] = -INFINITY;
      i[q] = INT_MAX;
    }
  }
  // keeps all KB entries sorted best-first (static indexing only: the list lives in registers)
  __device__ __forceinline__ void insert(float x, int idx) {
    if (!better(x, idx, v[KB - 1], i[KB - 1])) return;
#pragma unroll
    for (int q = KB - 1; q >= 0; --q) {
      const bool here = better(x, idx, v[q], i[q]);
      const bool above = (q > 0) && better(x, idx, v[q > 0 ? q - 1 : 0], i[q > 0 ? q - 1 : 0]);
      if (here) {
        v[q] = above ? v[q > 0 ? q - 1 : 0] : x;
        i[q] = above ? i[q > 0 ? q - 1 : 0] : idx;
      }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int q = 0; q < KB - 1; ++q) {
      v[q] = v[q + 1];
      i[q] = i[q + 1];
    }
    v[KB - 1] = -INFINITY;
    i[KB - 1] = INT_MAX;
  }
};

// KB rounds of "best head across the warp"; results land in out_v/out_i (valid in every lane)
template <int KB>
__device__ __forceinline__ void warp_merge(TopList<KB>& l, float (&out_v)[KB], int (&out_i)[KB]) {
#pragma unroll
  for (int r = 0; r < KB; ++r) {
    float bv = l.v[0];
    int bi = l.i[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    out_v[r] = bv;
    out_i[r] = bi;
    if (l.i[0] == bi && bi != INT_MAX) l.pop();
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel 1: one CTA per logits row.  Single pass over the row (each logit is read from HBM exactly
// once): online (max, sum-exp) and a register top-KB of the RAW logits.  Within a row the map
// x -> ((x - max) - log(sum)) + score is monotone, so the row's best final candidates are its best raw
// logits.  Writes one small partial record per row.
// ---------------------------------------------------------------------------------------------
constexpr int ROW_THREADS = 256;
constexpr int ROW_WARPS = ROW_THREADS / 32;
constexpr int ROW_NV4 = 16;                             // float4 per thread held in registers
constexpr int ROW_SPAN = ROW_THREADS * ROW_NV4 * 4;     // 16384 logits per register-resident pass
constexpr int ROW_CAP = 64;                             // candidate list capacity

template <int KB>
__global__ void __launch_bounds__(ROW_THREADS)
beam_row_kernel(const care_beam_state st, const float* __restrict__ logits, int64_t ldv, int step, int normalized) {
  __shared__ float wl_v[ROW_WARPS][KB];
  __shared__ int wl_i[ROW_WARPS][KB];
  __shared__ float sm_red[ROW_WARPS];
  __shared__ float sm_M, sm_T, sm_S;
  __shared__ int sm_cnt;
  __shared__ float cl_v[ROW_CAP];
  __shared__ int cl_i[ROW_CAP];
  __shared__ float keep_v[KB];   // running best across spans (V > ROW_SPAN only)
  __shared__ int keep_i[KB];
  const int r = blockIdx.x;
  const int K = st.K, V = st.V;
  const int v = r / K, b = r - v * K;
  if (st.done[v]) return;
  if (step == 1 && b > 0) return;                            // Beam.py:56
  if (step > 1 && st.cur_tok[r] == CARE_EOS) return;         // Beam.py:52-54 (handled in kernel 2)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* row = logits + (int64_t)r * ldv;
  float run_M = -INFINITY, run_S = 0.f;
  if (tid < KB) {
    keep_v[tid] = -INFINITY;
    keep_i[tid] = INT_MAX;
  }

  for (int base = 0; base < V; base += ROW_SPAN) {
    // ---- the whole span goes into registers: every logit is read from HBM exactly once ----
    const int n_here = min(ROW_SPAN, V - base);
    const int n4 = n_here >> 2;
    const float4* row4 = reinterpret_cast<const float4*>(row + base);
    float4 x[ROW_NV4];
#pragma unroll
    for (int u = 0; u < ROW_NV4; ++u) {
      const int c = tid + u * ROW_THREADS;
      x[u] = (c < n4) ? __ldcs(row4 + c) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    const int tail_idx = base + n4 * 4 + tid;                // n_here % 4 leftover scalars
    const float xt = (tid < (n_here & 3)) ? row[tail_idx] : -INFINITY;
    float m = xt;
#pragma unroll
    for (int u = 0; u < ROW_NV4; ++u) m = fmaxf(m, fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)));

    // ---- block max M and threshold T = KB-th largest of the per-thread maxima (a lower bound of the
    //      KB-th largest logit of the span, so only a handful of elements pass `x >= T`) ----
    TopList<KB> tl;
    tl.init();
    if (m > -INFINITY) tl.insert(m, tid);
    float ov[KB];
    int oi[KB];
    warp_merge<KB>(tl, ov, oi);
    __syncthreads();   // previous span's readers of wl_*/sm_* are done
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        wl_v[warp][q] = ov[q];
        wl_i[warp][q] = oi[q];
      }
    }
    if (tid == 0) sm_cnt = 0;
    __syncthreads();
    if (warp == 0) {
      TopList<KB> l;
      l.init();
      if (lane < ROW_WARPS) {
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          l.v[q] = wl_v[lane][q];
          l.i[q] = wl_i[lane][q];
        }
      }
      warp_merge<KB>(l, ov, oi);
      if (lane == 0) {
        sm_M = ov[0];
        sm_T = fmaxf(ov[KB - 1], keep_v[KB - 1]);   // nothing below the running KB-th best can matter
      }
    }
    __syncthreads();
    const float M = fmaxf(sm_M, run_M), T = sm_T;

    // ---- sum of exp(x - M) (two-pass statistics, as torch.log_softmax computes them) ----
    float s = (xt > -INFINITY) ? __expf(xt - M) : 0.f;
#pragma unroll
    for (int u = 0; u < ROW_NV4; ++u) {
      if (tid + u * ROW_THREADS < n4)
        s += __expf(x[u].x - M) + __expf(x[u].y - M) + __expf(x[u].z - M) + __expf(x[u].w - M);
    }
    s = warp_sum(s);
    if (lane == 0) sm_red[warp] = s;

    // ---- candidates: elements >= T are appended to a small shared list ----
#pragma unroll
    for (int u = 0; u < ROW_NV4; ++u) {
      const int c = tid + u * ROW_THREADS;
      const float xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w};
      if (c < n4 && fmaxf(fmaxf(xs[0], xs[1]), fmaxf(xs[2], xs[3])) >= T) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (xs[e] >= T) {
            const int pos = atomicAdd(&sm_cnt, 1);
            if (pos < ROW_CAP) {
              cl_v[pos] = xs[e];
              cl_i[pos] = base + c * 4 + e;
            }
          }
      }
    }
    if (xt >= T && xt > -INFINITY) {
      const int pos = atomicAdd(&sm_cnt, 1);
      if (pos < ROW_CAP) {
        cl_v[pos] = xt;
        cl_i[pos] = tail_idx;
      }
    }
    __syncthreads();
    const int cnt = sm_cnt;
    if (cnt > ROW_CAP) {
      // degenerate span (masses of equal logits): exact but slow per-thread selection
      TopList<KB> mine;
      mine.init();
#pragma unroll
      for (int u = 0; u < ROW_NV4; ++u) {
        const int c = tid + u * ROW_THREADS;
        if (c < n4) {
          mine.insert(x[u].x, base + c * 4 + 0);
          mine.insert(x[u].y, base + c * 4 + 1);
          mine.insert(x[u].z, base + c * 4 + 2);
          mine.insert(x[u].w, base + c * 4 + 3);
        }
      }
      if (xt > -INFINITY) mine.insert(xt, tail_idx);
      warp_merge<KB>(mine, ov, oi);
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          wl_v[warp][q] = ov[q];
          wl_i[warp][q] = oi[q];
        }
      }
      __syncthreads();
    }
    if (warp == 0) {
      TopList<KB> l;
      l.init();
      if (cnt > ROW_CAP) {
        if (lane < ROW_WARPS) {
#pragma unroll
          for (int q = 0; q < KB; ++q) l.insert(wl_v[lane][q], wl_i[lane][q]);
        }
      } else {
        for (int q = lane; q < cnt; q += 32) l.insert(cl_v[q], cl_i[q]);
      }
      if (lane < KB) l.insert(keep_v[lane], keep_i[lane]);   // carry the previous spans' best
      warp_merge<KB>(l, ov, oi);
      float ssum = lane < ROW_WARPS ? sm_red[lane] : 0.f;
      ssum = warp_sum(ssum);
      __syncwarp();   // every lane has read keep_* before lane 0 overwrites it
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          keep_v[q] = ov[q];
          keep_i[q] = oi[q];
        }
        sm_S = (run_M > -INFINITY ? run_S * __expf(run_M - M) : 0.f) + ssum;
      }
    }
    __syncthreads();
    run_M = M;
    run_S = sm_S;
  }
  if (tid == 0) {
    float* rec = st.scratch + (int64_t)r * (2 + 2 * KB);
    // `normalized`: the row already holds log-probabilities (ensemble mean): candidate value = x itself
    rec[0] = normalized ? 0.f : run_M;
    rec[1] = normalized ? 1.f : run_S;
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      rec[2 + q] = keep_v[q];
      reinterpret_cast<int*>(rec)[2 + KB + q] = keep_i[q];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Kernel 2: one warp per video.  Turns the K row records into final candidates
// ((x - max) - log(sum)) + score, merges them by (value desc, flat index asc), then applies
// Beam.advance's bookkeeping (Beam.py:61-85) and maintains the KV-cache ancestry table.
// ---------------------------------------------------------------------------------------------
constexpr int UPD_WARPS = 4;

// Where the fused vocabulary kernel (vocab_beam.cu) left the per-row records: partials[row][seg],
// seg < (runs touching the row's m-block); T tiles cut into G contiguous runs, n_tiles per m-block.
struct SegLayout {
  const float* partials;   // NULL: st.scratch already holds one record per row (beam_row_kernel)
  int nseg, n_tiles;
  int64_t T, G;
  int row_shift;   // log2(rows per row block of the kernel that wrote the records: 7 single-CTA, 8 CTA pair)
  int split;       // segment g of a run = column half g of its tiles (both exist) instead of its tiles of parity g
  int rpv;         // record rows per video: K, or 1 for the compact first step (one <bos> row per video, record row v)
};

// Prologue of the NEXT decode step, run by the video's warp once its beams are chosen (both optional):
//   x0 != NULL:   the decoder input rows of the K new tokens, LN(word[tok] + pos[step] + gsg[v]) - what embed_ln_kernel
//                 would compute in a launch of its own at the start of step + 1;
//   info != NULL: the live-slot record of the video for a prefix of step + 1 positions - what compact_info_kernel
//                 would compute before the next self-attention.
struct NextStep {
  const float* word;
  const float* pos;
  const float* gsg;
  const float* gamma;
  const float* beta;
  float eps;
  int d;
  h16* x0;
  float* x0_32;
  uint32_t* info;
};

template <int KB>
__global__ void __launch_bounds__(UPD_WARPS * 32)
beam_update_kernel(const care_beam_state st, const SegLayout sl, int step, int max_len, float* __restrict__ cand_val,
                   int32_t* __restrict__ cand_idx, const NextStep ns) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ uint32_t info_rec_all[UPD_WARPS][attn_mma::INFO_WORDS];
  __shared__ uint8_t old_anc_all[UPD_WARPS][8 * 64];
  __shared__ float fin_v_all[UPD_WARPS][KB];
  __shared__ int fin_i_all[UPD_WARPS][KB];
  __shared__ float row_rec_all[UPD_WARPS][8][2 + 2 * KB];   // merged (max, sum-exp, top-KB) record of each beam row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int v = blockIdx.x * UPD_WARPS + warp;
  if (v >= st.B) return;
  if (st.done[v]) return;
  uint8_t* old_anc = old_anc_all[warp];
  float* fin_v = fin_v_all[warp];
  int* fin_i = fin_i_all[warp];
  const int K = st.K, V = st.V, T = st.T_max;
  const int nsel = K + 1;
  // everything the bookkeeping below needs from the beam state is requested now, under the record merge
  for (int idx = lane; idx < K * T; idx += 32) old_anc[idx] = st.anc[(int64_t)v * K * T + idx];
  __shared__ int row_tok_all[UPD_WARPS][8];
  __shared__ float row_score_all[UPD_WARPS][8];
  if (lane < K && step > 1) {
    row_tok_all[warp][lane] = st.cur_tok[v * K + lane];
    row_score_all[warp][lane] = st.scores[v * K + lane];
  }
  int fin_count0 = 0;
  if (lane == 0) fin_count0 = st.fin_count[v];

  const float* rec_base = st.scratch + (int64_t)v * K * (2 + 2 * KB);   // one record per row, rows of this video
  if (sl.partials != nullptr) {
    // merge the row's segment records into one (max, sum-exp, top-KB) record, kept in shared memory
    rec_base = &row_rec_all[warp][0][0];
    const int max_slots = 2 * (int)((sl.n_tiles * sl.G + sl.T - 1) / sl.T + 1);   // upper bound of segments per row
    // large batches have a few segments per row: one lane per row.  Small batches put every tile in its own short run
    // (up to ~2 x n_tiles segments per row): four lanes per row, all rows of the video at once
    const int lpr = max_slots <= 8 ? 1 : 4;
    const int b = lane / lpr, sub = lane - b * lpr;
    float M = -INFINITY, S = 0.f;
    TopList<KB> l;
    l.init();
    if (b < sl.rpv) {
      const int rr = v * sl.rpv + b;         // row of the vocabulary kernel's records
      const int m_blk = rr >> sl.row_shift;
      const int c0 = (int)((((int64_t)m_blk * sl.n_tiles + 1) * sl.G - 1) / sl.T);
      const int c1 = (int)((((int64_t)m_blk * sl.n_tiles + sl.n_tiles) * sl.G - 1) / sl.T);
      const float* base = sl.partials + (int64_t)rr * sl.nseg * (2 + 2 * KB);
      const int64_t mlo = (int64_t)m_blk * sl.n_tiles, mhi = mlo + sl.n_tiles;
      const int nslot = 2 * (c1 - c0 + 1);
      for (int sidx = sub; sidx < nslot; sidx += lpr) {
        const int c = c0 + (sidx >> 1), g = sidx & 1;
        // segment 2*(c - c0) + g exists iff epilogue group g of run c saw a tile of this m-block
        const int64_t start = (int64_t)c * sl.T / sl.G, end = (int64_t)(c + 1) * sl.T / sl.G;
        const int64_t lo = start > mlo ? start : mlo, hi = end < mhi ? end : mhi;
        if (!(lo + (sl.split ? 0 : ((g - (lo - start)) & 1)) < hi)) continue;
        const float* rec = base + sidx * (2 + 2 * KB);
        const float m = rec[0];
        if (m == -INFINITY) continue;   // a column half past the last vocabulary column
        if (m > M) {
          S *= __expf(M - m);
          M = m;
        }
        S += rec[1] * __expf(rec[0] - M);
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          const int idx = reinterpret_cast<const int*>(rec)[2 + KB + q];
          if (idx != INT_MAX) l.insert(rec[2 + q], idx);
        }
      }
    }
    if (lpr == 4) {   // warp-uniform
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float Mo = __shfl_xor_sync(0xffffffffu, M, o), So = __shfl_xor_sync(0xffffffffu, S, o);
        const float Mn = fmaxf(M, Mo);
        S = (M == -INFINITY ? 0.f : S * __expf(M - Mn)) + (Mo == -INFINITY ? 0.f : So * __expf(Mo - Mn));
        M = Mn;
        float ov[KB];
        int oi[KB];
#pragma unroll
        for (int q = 0; q < KB; ++q) {
          ov[q] = __shfl_xor_sync(0xffffffffu, l.v[q], o);
          oi[q] = __shfl_xor_sync(0xffffffffu, l.i[q], o);
        }
#pragma unroll
        for (int q = 0; q < KB; ++q)
          if (oi[q] != INT_MAX) l.insert(ov[q], oi[q]);
      }
    }
    if (b < sl.rpv && sub == 0) {
      float* out = row_rec_all[warp][b];
      out[0] = M;
      out[1] = S;
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        out[2 + q] = l.v[q];
        reinterpret_cast<int*>(out)[2 + KB + q] = l.i[q];
      }
    }
  }
  __syncwarp();

  TopList<KB> mine;
  mine.init();
  for (int c = lane; c < K * KB; c += 32) {
    const int b = c / KB, q = c - b * KB;
    if (step == 1 && b > 0) continue;
    const float score_b = step > 1 ? row_score_all[warp][b] : 0.f;
    if (step > 1 && row_tok_all[warp][b] == CARE_EOS) {
      if (q < V) mine.insert(-1e20f, b * V + q);   // the whole row is -1e20 (Beam.py:52-54)
      continue;
    }
    const float* rec = rec_base + b * (2 + 2 * KB);
    const int tok = reinterpret_cast<const int*>(rec)[2 + KB + q];
    if (tok == INT_MAX) continue;
    float val = (rec[2 + q] - rec[0]) - logf(rec[1]);
    if (step > 1) val += score_b;
    mine.insert(val, b * V + tok);
  }
  float ov[KB];
  int oi[KB];
  warp_merge<KB>(mine, ov, oi);
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      fin_v[q] = ov[q];
      fin_i[q] = oi[q];
    }
  }
  __syncwarp();

  // ancestry: new_anc[b'][p] = old_anc[parent(b')][p] for p < step-1, new_anc[b'][step-1] = parent(b')
  // (a row of NaN logits leaves no candidate: fin_i stays INT_MAX - keep the indices inside the tables)
  if (lane < nsel && (unsigned)fin_i[lane] >= (unsigned)(K * V)) fin_i[lane] = 0;
  __syncwarp();
  for (int idx = lane; idx < K * T; idx += 32) {
    const int b = idx / T, pp = idx - b * T;
    const int parent = fin_i[b] / V;
    uint8_t a = 0;
    if (pp < step - 1) a = old_anc[parent * T + pp];
    else if (pp == step - 1) a = (uint8_t)parent;
    st.anc[(int64_t)v * K * T + idx] = a;
  }
  if (lane < K) {
    const int fi = fin_i[lane];
    const int parent = fi / V;
    const int tok = fi - parent * V;
    st.scores[v * K + lane] = fin_v[lane];
    st.prev_ks[((int64_t)v * T + (step - 1)) * K + lane] = parent;
    st.tok_hist[((int64_t)v * (T + 1) + step) * K + lane] = tok;
    st.cur_tok[v * K + lane] = tok;
  }
  if (cand_val != nullptr && lane < nsel) {
    cand_val[(int64_t)v * nsel + lane] = fin_v[lane];
    cand_idx[(int64_t)v * nsel + lane] = fin_i[lane];
  }
  int vdone = 0;
  if (lane == 0) {
    int count = fin_count0;
    bool done = false;
    for (int i = 0; i < K && !done; ++i) {  // Beam.py:72-76
      const int fi = fin_i[i];
      const int tok = fi - (fi / V) * V;
      if (tok == CARE_EOS) {
        st.fin_score[v * st.need + count] = fin_v[i];
        st.fin_t[v * st.need + count] = step;
        st.fin_k[v * st.need + count] = i;
        ++count;
        done = count >= st.need;
      }
    }
    if (!done && step + 1 == max_len) {  // Beam.py:79-84
      done = true;
      if (count == 0) {
        for (int i = 0; i < K; ++i) {
          st.fin_score[v * st.need + count] = fin_v[i];
          st.fin_t[v * st.need + count] = step;
          st.fin_k[v * st.need + count] = i;
          ++count;
        }
      }
    }
    st.fin_count[v] = count;
    if (done) {
      st.done[v] = 1;
      atomicAdd(st.n_done, 1);
    }
    vdone = done ? 1 : 0;
  }
  vdone = __shfl_sync(0xffffffffu, vdone, 0);
  if (vdone) return;
  // ---- prologue of step + 1 for a video that goes on ----
  if (ns.x0 != nullptr) {
    for (int b = 0; b < K; ++b) {
      const int fi = fin_i[b];
      const int tok = fi - (fi / V) * V;
      const int64_t row = (int64_t)v * K + b;
      rw::warp_embed_ln_row<h16>(tok, step, ns.word, ns.pos, nullptr, ns.gsg ? ns.gsg + (int64_t)v * ns.d : nullptr,
                                 ns.gamma, ns.beta, ns.eps, ns.d, lane, ns.x0 + row * ns.d,
                                 ns.x0_32 ? ns.x0_32 + row * ns.d : nullptr);
    }
  }
  if (ns.info != nullptr) {
    __syncwarp();   // this warp's ancestry / token writes above are visible to all of its lanes
    attn_mma::warp_compact_record(st.anc, T, st.tok_hist, (T + 1) * K, v, K, step + 1, lane, info_rec_all[warp], ns.info);
  }
}

__global__ void beam_init_kernel(const care_beam_state st, int bos) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int B = st.B, K = st.K, T = st.T_max;
  for (int64_t i = i0; i < (int64_t)B * K; i += stride) {
    st.scores[i] = 0.f;
    st.cur_tok[i] = bos;
  }
  for (int64_t i = i0; i < (int64_t)B * (T + 1) * K; i += stride) {
    const int64_t pos = (i / K) % (T + 1);
    st.tok_hist[i] = pos == 0 ? bos : CARE_PAD;
  }
  for (int64_t i = i0; i < (int64_t)B * T * K; i += stride) {
    st.prev_ks[i] = 0;
    st.anc[i] = 0;
  }
  for (int64_t i = i0; i < (int64_t)B * st.need; i += stride) {
    st.fin_score[i] = 0.f;
    st.fin_t[i] = 0;
    st.fin_k[i] = 0;
  }
  for (int64_t i = i0; i < B; i += stride) {
    st.fin_count[i] = 0;
    st.done[i] = 0;
  }
  if (i0 == 0) *st.n_done = 0;
}

// Translator.collect_hypothesis_and_scores + Beam.sort_finished / get_hypothesis (one thread per video)
__global__ void beam_finalize_kernel(const care_beam_state st, double alpha, int n_best, int32_t* __restrict__ out_tok,
                                     int32_t* __restrict__ out_len, float* __restrict__ out_score,
                                     int32_t* __restrict__ out_t) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= st.B) return;
  const int K = st.K, T = st.T_max, need = st.need;
  const int n = st.fin_count[v];
  bool used[16];
  for (int i = 0; i < 16; ++i) used[i] = false;
  for (int r = 0; r < n_best; ++r) {
    int best = -1;
    double best_key = 0.0;
    for (int i = 0; i < n; ++i) {
      if (used[i]) continue;
      const double key = (double)st.fin_score[v * need + i] / pow((double)st.fin_t[v * need + i], alpha);
      if (best < 0 || key > best_key) {  // strict '>' keeps insertion order among ties (stable sort)
        best = i;
        best_key = key;
      }
    }
    int32_t* dst = out_tok + ((int64_t)v * n_best + r) * T;
    for (int j = 0; j < T; ++j) dst[j] = CARE_PAD;
    if (best < 0) {
      out_len[v * n_best + r] = 0;
      out_score[v * n_best + r] = 0.f;
      out_t[v * n_best + r] = 0;
      continue;
    }
    used[best] = true;
    const int len = st.fin_t[v * need + best];
    int k = st.fin_k[v * need + best];
    for (int j = len - 1; j >= 0; --j) {  // Beam.py:119-132
      dst[j] = st.tok_hist[((int64_t)v * (T + 1) + j + 1) * K + k];
      k = st.prev_ks[((int64_t)v * T + j) * K + k];
    }
    out_len[v * n_best + r] = len;
    out_score[v * n_best + r] = st.fin_score[v * need + best];
    out_t[v * n_best + r] = len;
  }
}

// Model ensembling (Translator.py:111-133): out[r, :] = mean_i log_softmax(logits_i[r, :]), one CTA per row.
constexpr int MAX_MODELS = 8;
struct ModelLogits {
  const float* p[MAX_MODELS];
};
__global__ void __launch_bounds__(256) ensemble_logprob_kernel(const ModelLogits ml, int n, int64_t ldv, int V,
                                                               float* __restrict__ out) {
  __shared__ float red[8];
  __shared__ float lse[MAX_MODELS];
  const int r = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = 0; i < n; ++i) {
    const float* x = ml.p[i] + (int64_t)r * ldv;
    float m = -INFINITY;
    for (int c = tid; c < V; c += 256) m = fmaxf(m, x[c]);
    m = warp_max(m);
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int c = tid; c < V; c += 256) s += expf(x[c] - m);
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      lse[i] = m + logf(t);   // log_softmax(x) = (x - m) - log(sum) = x - lse
    }
    __syncthreads();
  }
  for (int c = tid; c < V; c += 256) {
    float acc = 0.f;
    for (int i = 0; i < n; ++i) acc += ml.p[i][(int64_t)r * ldv + c] - lse[i];
    out[(int64_t)r * ldv + c] = acc / (float)n;
  }
}

static int check_state(const care_beam_state* st, const char* who) {
  CARE_CHECK_ARG(st != nullptr, "%s: state is NULL", who);
  CARE_CHECK_ARG(st->B > 0 && st->K >= 1 && st->K <= 8, "%s: K=%d must be in [1,8]", who, st->K);
  CARE_CHECK_ARG(st->T_max >= 1 && st->T_max <= 64, "%s: T_max=%d must be in [1,64]", who, st->T_max);
  CARE_CHECK_ARG(st->need >= st->K && st->need <= 16, "%s: need=%d must be in [K,16]", who, st->need);
  CARE_CHECK_ARG(st->V > st->K, "%s: V=%d too small", who, st->V);
  CARE_CHECK_ARG(st->scores && st->cur_tok && st->tok_hist && st->prev_ks && st->anc && st->fin_score && st->fin_t &&
                     st->fin_k && st->fin_count && st->done && st->n_done,
                 "%s: NULL state buffer", who);
  return 0;
}

}  // namespace beam

namespace vb {  // vocab_beam.cu
void seg_layout(const care_ctx* ctx, int R, int V, int* n_tiles, int64_t* T, int64_t* G, int* row_shift, int* split);
int nseg_for(const care_ctx* ctx, int R, int V);
}
}  // namespace care

using namespace care;

extern "C" {

int care_beam_init(care_ctx* ctx, const care_beam_state* st, int bos, void* stream) {
  CARE_CHECK_ARG(ctx != nullptr, "care_beam_init: ctx is NULL");
  if (beam::check_state(st, "care_beam_init")) return -1;
  const int64_t n = (int64_t)st->B * (st->T_max + 1) * st->K;
  const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8);
  ctx->info_ready_npos = -1;
  ctx->next_armed = false;
  beam::beam_init_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*st, bos);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

static int beam_step_impl(care_ctx* ctx, const care_beam_state* st, const float* logits, int64_t ldv, int step,
                          int max_len, float* cand_val, int32_t* cand_idx, void* stream, int normalized) {
  CARE_CHECK_ARG(ctx && logits, "care_beam_step: bad args");
  if (beam::check_state(st, "care_beam_step")) return -1;
  CARE_CHECK_ARG(step >= 1 && step <= st->T_max, "care_beam_step: step=%d outside [1,%d]", step, st->T_max);
  CARE_CHECK_ARG(ldv % 4 == 0 && ldv >= st->V, "care_beam_step: ldv=%lld must be a multiple of 4 and >= V",
                 (long long)ldv);
  CARE_CHECK_ARG((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "care_beam_step: logits must be 16-byte aligned");
  CARE_CHECK_ARG((cand_val == nullptr) == (cand_idx == nullptr), "care_beam_step: cand_val/cand_idx must go together");
  CARE_CHECK_ARG(st->scratch != nullptr, "care_beam_step: state.scratch is NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const int K = st->K, R = st->B * st->K;
  const int ugrid = (st->B + beam::UPD_WARPS - 1) / beam::UPD_WARPS, uthreads = beam::UPD_WARPS * 32;
  ctx->info_ready_npos = -1;   // the ancestry changes: records written for it are stale
  ctx->next_armed = false;
#define CARE_BEAM_GO(KB_)                                                                                  \
  do {                                                                                                     \
    beam::beam_row_kernel<KB_><<<R, beam::ROW_THREADS, 0, s>>>(*st, logits, ldv, step, normalized);        \
    CARE_LAUNCH_CHECK(ctx);                                                                                \
    beam::beam_update_kernel<KB_><<<ugrid, uthreads, 0, s>>>(*st, beam::SegLayout{}, step, max_len, cand_val, \
                                                              cand_idx, beam::NextStep{});                 \
  } while (0)
  if (K <= 1) CARE_BEAM_GO(2);
  else if (K <= 3) CARE_BEAM_GO(4);
  else if (K <= 5) CARE_BEAM_GO(6);
  else CARE_BEAM_GO(9);
#undef CARE_BEAM_GO
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_beam_step(care_ctx* ctx, const care_beam_state* st, const float* logits, int64_t ldv, int step, int max_len,
                   float* cand_val, int32_t* cand_idx, void* stream) {
  return beam_step_impl(ctx, st, logits, ldv, step, max_len, cand_val, cand_idx, stream, 0);
}

int care_beam_step_logprobs(care_ctx* ctx, const care_beam_state* st, const float* logprobs, int64_t ldv, int step,
                            int max_len, float* cand_val, int32_t* cand_idx, void* stream) {
  return beam_step_impl(ctx, st, logprobs, ldv, step, max_len, cand_val, cand_idx, stream, 1);
}

int care_ensemble_logprobs(care_ctx* ctx, const float* const* logits, int n, int64_t ldv, int R, int V, float* out,
                           void* stream) {
  CARE_CHECK_ARG(ctx && logits && out && n >= 1 && n <= beam::MAX_MODELS && R > 0 && V > 0,
                 "care_ensemble_logprobs: bad args (n=%d, at most %d models)", n, beam::MAX_MODELS);
  beam::ModelLogits ml{};
  for (int i = 0; i < n; ++i) {
    CARE_CHECK_ARG(logits[i] != nullptr, "care_ensemble_logprobs: logits[%d] is NULL", i);
    ml.p[i] = logits[i];
  }
  beam::ensemble_logprob_kernel<<<R, 256, 0, (cudaStream_t)stream>>>(ml, n, ldv, V, out);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

static int beam_step_partials_impl(care_ctx* ctx, const care_beam_state* st, const float* partials, int nseg, int step,
                                   int max_len, int rpv, float* cand_val, int32_t* cand_idx, void* stream, const char* who) {
  CARE_CHECK_ARG(ctx && partials && nseg >= 1, "%s: bad args", who);
  if (beam::check_state(st, who)) return -1;
  CARE_CHECK_ARG(step >= 1 && step <= st->T_max, "%s: step=%d outside [1,%d]", who, step, st->T_max);
  CARE_CHECK_ARG((cand_val == nullptr) == (cand_idx == nullptr), "%s: cand_val/cand_idx must go together", who);
  CARE_CHECK_ARG(st->scratch != nullptr, "%s: state.scratch is NULL", who);
  cudaStream_t s = (cudaStream_t)stream;
  const int K = st->K;
  beam::SegLayout sl{};
  sl.partials = partials;
  sl.nseg = nseg;
  sl.rpv = rpv;
  vb::seg_layout(ctx, st->B * rpv, st->V, &sl.n_tiles, &sl.T, &sl.G, &sl.row_shift, &sl.split);
  // the next step's prologue rides along: its input rows when the caller armed them (care_ctx_set_next_step), the
  // live-slot record when the next self-attention will be the chunk stream over this ctx's record table
  beam::NextStep ns{};
  const bool has_next = step + 1 < max_len;
  if (ctx->next_armed) {
    ctx->next_armed = false;
    if (has_next) {
      const care_next_step& n = ctx->next;
      CARE_CHECK_ARG(n.word_emb && n.pos_emb && n.gamma && n.beta && n.x0 && n.d > 0 && n.d % 128 == 0 && n.d <= 1024,
                     "%s: bad care_next_step", who);
      ns.word = n.word_emb; ns.pos = n.pos_emb; ns.gsg = n.gsg; ns.gamma = n.gamma; ns.beta = n.beta;
      ns.eps = n.eps; ns.d = n.d; ns.x0 = static_cast<h16*>(n.x0); ns.x0_32 = n.x0_f32;
    }
  }
  ctx->info_ready_npos = -1;
  if (has_next && ctx->fuse_info && ctx->compact_info != nullptr && st->B <= ctx->compact_info_videos && step + 1 <= 64 && st->K <= 8 &&
      (ctx->self_compact == 3 || (ctx->self_compact == 2 && step + 1 >= 6))) {
    ns.info = ctx->compact_info;
    ctx->info_ready_npos = step + 1;
    ctx->info_ready_B = st->B;
    ctx->info_ready_anc = st->anc;
  }
  CARE_CHECK_ARG(nseg == vb::nseg_for(ctx, st->B * rpv, st->V), "%s: nseg=%d does not match the %d-row record table", who,
                 nseg, st->B * rpv);
  const int ugrid = (st->B + beam::UPD_WARPS - 1) / beam::UPD_WARPS, uthreads = beam::UPD_WARPS * 32;
  if (K <= 1)
    CARE_CUDA(launch_pdl(ctx, beam::beam_update_kernel<2>, dim3(ugrid), dim3(uthreads), 0, s, *st, sl, step, max_len, cand_val, cand_idx, ns));
  else if (K <= 3)
    CARE_CUDA(launch_pdl(ctx, beam::beam_update_kernel<4>, dim3(ugrid), dim3(uthreads), 0, s, *st, sl, step, max_len, cand_val, cand_idx, ns));
  else if (K <= 5)
    CARE_CUDA(launch_pdl(ctx, beam::beam_update_kernel<6>, dim3(ugrid), dim3(uthreads), 0, s, *st, sl, step, max_len, cand_val, cand_idx, ns));
  else
    CARE_CUDA(launch_pdl(ctx, beam::beam_update_kernel<9>, dim3(ugrid), dim3(uthreads), 0, s, *st, sl, step, max_len, cand_val, cand_idx, ns));
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int care_beam_step_partials(care_ctx* ctx, const care_beam_state* st, const float* partials, int nseg, int step,
                            int max_len, float* cand_val, int32_t* cand_idx, void* stream) {
  return beam_step_partials_impl(ctx, st, partials, nseg, step, max_len, st ? st->K : 1, cand_val, cand_idx, stream,
                                 "care_beam_step_partials");
}

int care_beam_first_step_partials(care_ctx* ctx, const care_beam_state* st, const float* partials, int nseg, int max_len,
                                  float* cand_val, int32_t* cand_idx, void* stream) {
  return beam_step_partials_impl(ctx, st, partials, nseg, 1, max_len, 1, cand_val, cand_idx, stream,
                                 "care_beam_first_step_partials");
}

int care_beam_finalize(care_ctx* ctx, const care_beam_state* st, double alpha, int n_best, int32_t* out_tokens,
                       int32_t* out_len, float* out_score, int32_t* out_t, void* stream) {
  CARE_CHECK_ARG(ctx && out_tokens && out_len && out_score && out_t && n_best >= 1, "care_beam_finalize: bad args");
  if (beam::check_state(st, "care_beam_finalize")) return -1;
  beam::beam_finalize_kernel<<<(st->B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(*st, alpha, n_best, out_tokens,
                                                                                  out_len, out_score, out_t);
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // extern "C"
