// Vocabulary projection fused with the first half of the beam step (bf16 mode).
//
// logits[R, V] = x[R, d] * Wvocab[V, d]^T is the largest GEMM of a decode step, and the only
// consumer of the logits is log_softmax + top-k (Translator.py:127, Beam.py:45-60).  Writing the fp32
// logits and reading them back costs 2 * R * V * 4 bytes of HBM traffic per step (2.4 GB at cfg4) -
// more time than the tensor work itself.  This kernel keeps the mainloop of gemm_tcgen05.cu (TMA
// producer warp, single-thread tcgen05.mma issuer, double-buffered TMEM accumulator) and replaces
// the store epilogue: each epilogue thread owns one row (one TMEM lane) and folds the tile's 256
// columns into (running max, running sum-exp, top-KB raw logits with their column ids).
//
// Tile schedule: the m_tiles x n_tiles grid is linearised n-fastest and cut into one contiguous run
// per CTA; two groups of four epilogue warps take alternate tiles of the run (group g drains TMEM
// accumulator g), so a thread carries its row state across every other n-tile of the run (the
// top-KB list saturates after the first tile; later tiles rarely insert).  A group flushes one
// 2+2*KB-word record per row whenever it leaves an m-block: partials[row][segment], segment =
// 2 * (this CTA's index - first CTA that touches the m-block) + group.  beam.cu merges the records
// (and derives from the same arithmetic which segments exist).
//
// SPLIT schedule (short runs: strong-scaling shards and small batches): both groups work on EVERY tile, group g folding
// columns [128 g, 128 g + 128) of it, so a tile's fold takes half as long.  The MMAs of tile i + 2 reuse the accumulator
// of tile i and must wait for its fold: measured with clock stamps at 2560 rows (scripts/vb_trace.py), a 256-column
// fold takes 11-15 k cycles (26 k for the first tile of a run, while the top-KB lists fill) against 9.4 k cycles of MMAs
// per tile, and the tensor pipe idled a third of the kernel.  Records keep their shape: segment = 2 * run + g, where g
// is now the column half; both segments of a run exist whenever the run touches the m-block.
#include <cfloat>
#include <climits>

#include "tcgen05_util.cuh"

namespace care {
namespace vb {

using namespace care::tc;

constexpr int BN = 256;
constexpr int VB_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-7, 8-11: two epilogue groups

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ bool better(float va, int ia, float vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

// CTA that owns linear tile t when T tiles are cut into G balanced contiguous runs (start_c = c*T/G)
__host__ __device__ inline int run_of_tile(int64_t t, int64_t T, int64_t G) { return (int)(((t + 1) * G - 1) / T); }

// Folds one 128 x 256 accumulator tile (this thread's row = its TMEM lane) into the row state: running
// max, running sum of exp, top-KB raw logits with their column ids.
// EDGE: the tile holds the last vocabulary column (TMA zero-fills the columns past V; they must not take part).  Interior
// tiles are instantiated without the per-element column test - it was a fifth of the loop's instructions.
template <int KB, bool EDGE>
__device__ __forceinline__ void fold_tile_impl(uint32_t tmem_tile, int n_blk, int N, float& run_m, float& run_s,
                                               float (&tv)[KB], int (&ti)[KB], int c_begin, int c_end) {
  constexpr float LOG2E = 1.4426950408889634f;
#pragma unroll 1
for (int c = c_begin; c < c_end; ++c) {
  const int col0 = n_blk * BN + c * 32;
  if (EDGE && col0 >= N) break;
  uint32_t v[32];
  tmem_ld32(tmem_tile + c * 32, v);
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    x[j] = __uint_as_float(v[j]);
    if (EDGE && col0 + j >= N) x[j] = -INFINITY;
  }
  float cm = x[0];
#pragma unroll
  for (int j = 1; j < 32; ++j) cm = fmaxf(cm, x[j]);
  if (cm > run_m) {
    run_s *= ex2_approx((run_m - cm) * LOG2E);
    run_m = cm;
  }
  const float neg_m2 = -run_m * LOG2E;
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    p0 += ex2_approx(fmaf(x[j], LOG2E, neg_m2));
    p1 += ex2_approx(fmaf(x[j + 1], LOG2E, neg_m2));
    p2 += ex2_approx(fmaf(x[j + 2], LOG2E, neg_m2));
    p3 += ex2_approx(fmaf(x[j + 3], LOG2E, neg_m2));
  }
  run_s += (p0 + p1) + (p2 + p3);
  // candidates: repeatedly extract the chunk maximum while it beats the KB-th best so far
  float cur = cm;
  while (cur > tv[KB - 1]) {   // ascending columns: an equal value with a larger index loses
    int sel = 0;
#pragma unroll
    for (int j = 31; j >= 0; --j) sel = (x[j] == cur) ? j : sel;
    const int idx = col0 + sel;
#pragma unroll
    for (int q = KB - 1; q >= 0; --q) {
      const bool here = better(cur, idx, tv[q], ti[q]);
      const bool above = (q > 0) && better(cur, idx, tv[q > 0 ? q - 1 : 0], ti[q > 0 ? q - 1 : 0]);
      if (here) {
        tv[q] = above ? tv[q > 0 ? q - 1 : 0] : cur;
        ti[q] = above ? ti[q > 0 ? q - 1 : 0] : idx;
      }
    }
    float nm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      x[j] = (j == sel) ? -INFINITY : x[j];
      nm = fmaxf(nm, x[j]);
    }
    cur = nm;
  }
}
}

template <int KB>
__device__ __forceinline__ void fold_tile(uint32_t tmem_tile, int n_blk, int N, float& run_m, float& run_s,
                                          float (&tv)[KB], int (&ti)[KB], int c_begin = 0, int c_end = BN / 32) {
  if ((n_blk + 1) * BN > N) fold_tile_impl<KB, true>(tmem_tile, n_blk, N, run_m, run_s, tv, ti, c_begin, c_end);
  else fold_tile_impl<KB, false>(tmem_tile, n_blk, N, run_m, run_s, tv, ti, c_begin, c_end);
}

template <int KB>
__device__ __forceinline__ void flush_record(float* rec, float& run_m, float& run_s, float (&tv)[KB], int (&ti)[KB],
                                             bool write) {
  if (write) {
    rec[0] = run_m;
    rec[1] = run_s;
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      rec[2 + q] = tv[q];
      reinterpret_cast<int*>(rec)[2 + KB + q] = ti[q];
    }
  }
  run_m = -INFINITY;
  run_s = 0.f;
#pragma unroll
  for (int q = 0; q < KB; ++q) {
    tv[q] = -INFINITY;
    ti[q] = INT_MAX;
  }
}

template <int KB, bool SPLIT>
__global__ void __launch_bounds__(VB_THREADS, 1)
vocab_beam_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                          float* __restrict__ partials, int nseg, int M, int N, int K, const EarlyExit ee) {
  if (all_done(ee)) return;
  using cfg = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t bar_base = smem_base + cfg::STAGES * cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (cfg::STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * cfg::STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * cfg::STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * cfg::STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (N + BN - 1) / BN;
  const int64_t T = (int64_t)m_tiles * n_tiles, G = gridDim.x;
  const int t_begin = (int)((int64_t)blockIdx.x * T / G), t_end = (int)((int64_t)(blockIdx.x + 1) * T / G);
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < cfg::STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), SPLIT ? 2 * EPI_WARPS : EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(static_cast<uint32_t>(cfg::TMEM_COLS))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();   // prologue done under the previous kernel's tail; its outputs are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_w = l2_policy(ee.l2_hints & 1);
      uint32_t it = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int m_blk = t / n_tiles, n_blk = t - m_blk * n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % cfg::STAGES;
          const uint32_t ph = (it / cfg::STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), cfg::STAGE_BYTES);
          const uint32_t a_dst = smem_base + s * cfg::STAGE_BYTES;
          tma_load_2d(a_dst, &tma_a, full_bar(s), kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d_hint(a_dst + cfg::A_BYTES, &tma_b, full_bar(s), kb * BLOCK_K, n_blk * BN, pol_w);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(BLOCK_M, BN);
      uint32_t it = 0, tcount = 0;
      for (int t = t_begin; t < t_end; ++t, ++tcount) {
        const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar(acc), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % cfg::STAGES;
          const uint32_t ph = (it / cfg::STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * cfg::STAGE_BYTES;
          const uint64_t adesc = sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = sw128_kmajor_desc(a_addr + cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            tc_mma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit(empty_bar(s));
        }
        tc_commit(tfull_bar(acc));
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: one thread per row; two groups of 4 warps take alternate tiles (group g owns TMEM
    // accumulator g), so each group has two mainloop times per tile.  A thread carries (max, sum-exp,
    // top-KB) across its group's tiles of the run and flushes when it leaves the m-block. =====
    const int grp = (warp - 4) >> 2;
    const int ew = (warp - 4) & 3;   // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
    float run_m = -INFINITY, run_s = 0.f;
    float tv[KB];
    int ti[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      tv[q] = -INFINITY;
      ti[q] = INT_MAX;
    }
    uint32_t gcount = 0;
    constexpr int STEP = SPLIT ? 1 : 2;
    for (int t = t_begin + (SPLIT ? 0 : grp); t < t_end; t += STEP, ++gcount) {
      const int m_blk = t / n_tiles, n_blk = t - m_blk * n_tiles;
      const uint32_t acc = SPLIT ? (gcount & 1u) : (uint32_t)grp;
      const uint32_t aph = SPLIT ? ((gcount >> 1) & 1u) : (gcount & 1u);
      mbar_wait(tfull_bar(acc), aph);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + ew * 32 + lane;
      if constexpr (SPLIT)
        fold_tile<KB>(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN, n_blk, N, run_m, run_s, tv, ti,
                      grp * (BN / 64), (grp + 1) * (BN / 64));
      else
        fold_tile<KB>(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN, n_blk, N, run_m, run_s, tv, ti);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      // this group's next tile is outside the m-block (or the run): flush the row record, start over
      const bool last = (t + STEP >= t_end) || ((t + STEP) / n_tiles != m_blk);
      if (last) {
        const int seg = 2 * ((int)blockIdx.x - run_of_tile((int64_t)m_blk * n_tiles, T, G)) + grp;
        flush_record<KB>(partials + ((int64_t)row * nseg + seg) * (2 + 2 * KB), run_m, run_s, tv, ti, row < M);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(cfg::TMEM_COLS))
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (see gemm_tcgen05_2sm.cu for the protocol): the two SMs of a TPC share one
// 256-row x 256-column tile, each CTA stages its own 128 rows of x and HALF of the Wvocab tile and
// folds its own 128 rows.  Runs are per cluster; a record belongs to (row, 2 * (cluster - first
// cluster touching the 256-row block) + epilogue group).
// ---------------------------------------------------------------------------------------------
namespace p2 {
constexpr int PAIR_M = 2 * BLOCK_M;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGES = 6;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ void commit_2sm_multicast(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}
}  // namespace p2

// Timeline instrumentation of the CTA-pair kernel (scripts/vb_trace.py; built only with -DCARE_VB_TRACE into
// lib/libcare_b200_trace.so): SM clock stamps of the leader CTA's roles.  [cluster][role][event]: role 0 = kernel
// (0 entry, 1 roles start, 2 exit), 1 = TMA producer (first k-block of each tile issued), 2 = MMA issuer (per tile:
// accumulator free, first operands landed, last MMA issued), 3 / 4 = epilogue group 0 / 1, warp ew = 0 (per tile:
// accumulator full, fold done).
#ifdef CARE_VB_TRACE
constexpr int VBT_EVENTS = 64;
__device__ long long g_vb_trace[128][5][VBT_EVENTS];
#define VBT(role, ev)                                                                                         \
  do {                                                                                                        \
    if (rank == 0 && cluster_id < 128 && (ev) < VBT_EVENTS) g_vb_trace[cluster_id][role][ev] = clock64();     \
  } while (0)
#else
#define VBT(role, ev) do { } while (0)
#endif

template <int KB, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(VB_THREADS, 1)
vocab_beam_2sm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                      float* __restrict__ partials, int nseg, int M, int N, int K, const EarlyExit ee) {
  using namespace p2;
  if (all_done(ee)) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int m_pairs = (M + PAIR_M - 1) / PAIR_M;
  const int n_tiles = (N + BN - 1) / BN;
  const int64_t T = (int64_t)m_pairs * n_tiles, G = gridDim.x >> 1;
  const int t_begin = (int)((int64_t)cluster_id * T / G), t_end = (int)((int64_t)(cluster_id + 1) * T / G);
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  if (threadIdx.x == 0) VBT(0, 0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), SPLIT ? 4 * EPI_WARPS : 2 * EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(static_cast<uint32_t>(2 * BN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();
  pdl_launch_dependents();
  if (threadIdx.x == 0) VBT(0, 1);

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_w = l2_policy(ee.l2_hints & 1);
      uint32_t it = 0;
      for (int t = t_begin; t < t_end; ++t) {
        const int m_pair = t / n_tiles, n_blk = t - m_pair * n_tiles;
        const int a_row = m_pair * PAIR_M + (int)rank * BLOCK_M;
        const int b_row = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (kb == 0) VBT(1, t - t_begin);
          if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);
          const uint32_t a_dst = smem_base + s * STAGE_BYTES;
          tma_load_2d_2sm(a_dst, &tma_a, full_bar(s), kb * BLOCK_K, a_row);
          tma_load_2d_cta2_hint(a_dst + A_BYTES, &tma_b, full_bar(s) & PEER_MASK, kb * BLOCK_K, b_row, pol_w);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = instr_desc_bf16(PAIR_M, BN);
      uint32_t it = 0, tcount = 0;
      for (int t = t_begin; t < t_end; ++t, ++tcount) {
        const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar(acc), aph ^ 1u);
        tc_fence_after();
        VBT(2, 3 * (t - t_begin));
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          if (kb == 0) VBT(2, 3 * (t - t_begin) + 1);
          const uint32_t a_addr = smem_base + s * STAGE_BYTES;
          const uint64_t adesc = sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            mma_bf16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          commit_2sm_multicast(empty_bar(s));
        }
        commit_2sm_multicast(tfull_bar(acc));
        VBT(2, 3 * (t - t_begin) + 2);
      }
    }
  } else if (warp >= 4) {
    const int grp = (warp - 4) >> 2;
    const int ew = (warp - 4) & 3;
    float run_m = -INFINITY, run_s = 0.f;
    float tv[KB];
    int ti[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      tv[q] = -INFINITY;
      ti[q] = INT_MAX;
    }
    uint32_t gcount = 0;
    constexpr int STEP = SPLIT ? 1 : 2;
    for (int t = t_begin + (SPLIT ? 0 : grp); t < t_end; t += STEP, ++gcount) {
      const int m_pair = t / n_tiles, n_blk = t - m_pair * n_tiles;
      const uint32_t acc = SPLIT ? (gcount & 1u) : (uint32_t)grp;
      const uint32_t aph = SPLIT ? ((gcount >> 1) & 1u) : (gcount & 1u);
      mbar_wait(tfull_bar(acc), aph);
      tc_fence_after();
      if (ew == 0 && lane == 0) VBT(3 + grp, 2 * (int)gcount);
      const int row = m_pair * PAIR_M + (int)rank * BLOCK_M + ew * 32 + lane;
      if constexpr (SPLIT)
        fold_tile<KB>(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN, n_blk, N, run_m, run_s, tv, ti,
                      grp * (BN / 64), (grp + 1) * (BN / 64));
      else
        fold_tile<KB>(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BN, n_blk, N, run_m, run_s, tv, ti);
      tc_fence_before();
      __syncwarp();
      if (ew == 0 && lane == 0) VBT(3 + grp, 2 * (int)gcount + 1);
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      const bool last = (t + STEP >= t_end) || ((t + STEP) / n_tiles != m_pair);
      if (last) {
        const int seg = 2 * (cluster_id - run_of_tile((int64_t)m_pair * n_tiles, T, G)) + grp;
        flush_record<KB>(partials + ((int64_t)row * nseg + seg) * (2 + 2 * KB), run_m, run_s, tv, ti, row < M);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (threadIdx.x == 0) VBT(0, 2);
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(2 * BN))
                 : "memory");
  }
}

// Which kernel serves (R, V) and how its tiles are cut into runs.  Deterministic in the shape, because the
// consumers of the records (beam.cu, nar.cu) recompute it.
struct Layout {
  int split;       // both epilogue groups fold every tile (column halves) instead of alternate tiles
  int two_sm;      // CTA-pair kernel (256-row blocks) or single-CTA kernel (128-row blocks)
  int row_shift;   // log2(rows per block)
  int n_tiles;
  int64_t T, G;    // tiles, runs
};

static Layout layout_for(const care_ctx* ctx, int R, int V) {
  Layout l{};
  l.n_tiles = (V + BN - 1) / BN;
  const int64_t pair_tiles = (int64_t)((R + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * l.n_tiles;
  const int n_clusters = ctx->sm_count / 2;
  if (ctx->vocab_2sm && pair_tiles >= 2 * n_clusters) {
    l.two_sm = 1;
    l.row_shift = 8;
    l.T = pair_tiles;
    l.G = n_clusters;
  } else {
    l.two_sm = 0;
    l.row_shift = 7;
    l.T = (int64_t)((R + BLOCK_M - 1) / BLOCK_M) * l.n_tiles;
    l.G = std::min<int64_t>(l.T, ctx->sm_count);
  }
  // runs of a few tiles are bound by the latency of a tile's fold, long runs by its throughput (option "vocab_split":
  // 0 = alternate tiles always, 1 = column halves always, 2 = column halves for runs shorter than vocab_split_tiles)
  l.split = ctx->vocab_split == 1 || (ctx->vocab_split == 2 && l.T < (int64_t)ctx->vocab_split_tiles * l.G);
  return l;
}

void seg_layout(const care_ctx* ctx, int R, int V, int* n_tiles, int64_t* T, int64_t* G, int* row_shift, int* split) {
  const Layout l = layout_for(ctx, R, V);
  *n_tiles = l.n_tiles;
  *T = l.T;
  *G = l.G;
  *row_shift = l.row_shift;
  *split = l.split;
}

bool uses_pairs(const care_ctx* ctx, int R, int V) { return layout_for(ctx, R, V).two_sm != 0; }

// largest number of runs that touch one row block, times the two epilogue groups
int nseg_for(const care_ctx* ctx, int R, int V) {
  const Layout l = layout_for(ctx, R, V);
  const int blocks = (int)(l.T / l.n_tiles);
  int best = 1;
  for (int m = 0; m < blocks; ++m) {
    const int c0 = run_of_tile((int64_t)m * l.n_tiles, l.T, l.G);
    const int c1 = run_of_tile((int64_t)m * l.n_tiles + l.n_tiles - 1, l.T, l.G);
    best = std::max(best, c1 - c0 + 1);
  }
  return 2 * best;
}

template <int KB>
static int launch(care_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, float* partials, int nseg, int R, int V,
                  int d, cudaStream_t stream) {
  using cfg = Cfg<BN>;
  const Layout l = layout_for(ctx, R, V);
  static bool configured_all[64] = {false};   // per device: function attributes are per device
  bool& configured = configured_all[ctx->device & 63];
  if (!configured) {
    CARE_CUDA(cudaFuncSetAttribute(vocab_beam_2sm_kernel<KB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p2::SMEM_BYTES));
    CARE_CUDA(cudaFuncSetAttribute(vocab_beam_2sm_kernel<KB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p2::SMEM_BYTES));
    CARE_CUDA(cudaFuncSetAttribute(vocab_beam_tcgen05_kernel<KB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    CARE_CUDA(cudaFuncSetAttribute(vocab_beam_tcgen05_kernel<KB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    configured = true;
  }
  if (l.two_sm) {
    auto kern = l.split ? vocab_beam_2sm_kernel<KB, true> : vocab_beam_2sm_kernel<KB, false>;
    CARE_CUDA(launch_pdl(ctx, kern, dim3(2 * (int)l.G), dim3(VB_THREADS), p2::SMEM_BYTES, stream, ta, tb, partials, nseg, R, V,
                         d, early_exit_of(ctx)));
    ctx->last_vocab = "vocab_beam_2sm_kernel";
  } else {
    auto kern = l.split ? vocab_beam_tcgen05_kernel<KB, true> : vocab_beam_tcgen05_kernel<KB, false>;
    CARE_CUDA(launch_pdl(ctx, kern, dim3((int)l.G), dim3(VB_THREADS), cfg::SMEM_BYTES, stream, ta, tb, partials, nseg, R, V, d,
                         early_exit_of(ctx)));
    ctx->last_vocab = "vocab_beam_tcgen05_kernel";
  }
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace vb
}  // namespace care

using namespace care;

extern "C" {

int care_vocab_beam_nseg(care_ctx* ctx, int R, int V) {
  if (!ctx || R <= 0 || V <= 0) return -1;
  return vb::nseg_for(ctx, R, V);
}

int care_vocab_beam_partials(care_ctx* ctx, const void* x, int64_t ldx, const void* W, int64_t ldw, int R, int V,
                             int d, int K, float* partials, int nseg, void* stream) {
  CARE_CHECK_ARG(ctx && x && W && partials && R > 0 && V > 0 && d > 0, "care_vocab_beam_partials: bad args");
  CARE_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0, "care_vocab_beam_partials: ldx/ldw must be multiples of 8");
  CARE_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                 "care_vocab_beam_partials: x, W must be 16-byte aligned");
  CARE_CHECK_ARG(K >= 1 && K <= 8, "care_vocab_beam_partials: K=%d must be in [1,8]", K);
  CARE_CHECK_ARG(nseg == vb::nseg_for(ctx, R, V), "care_vocab_beam_partials: nseg=%d, expected %d", nseg,
                 vb::nseg_for(ctx, R, V));
  CUtensorMap ta, tb;
  {
    const uint64_t gdim[2] = {(uint64_t)d, (uint64_t)R};
    const uint64_t gstr[1] = {(uint64_t)ldx * 2};
    const uint32_t box[2] = {(uint32_t)tc::BLOCK_K, (uint32_t)tc::BLOCK_M};
    int rc = get_tmap_bf16(ctx, x, 2, gdim, gstr, box, &ta);
    if (rc) return rc;
  }
  {
    const uint64_t gdim[2] = {(uint64_t)d, (uint64_t)V};
    const uint64_t gstr[1] = {(uint64_t)ldw * 2};
    const uint32_t box[2] = {(uint32_t)tc::BLOCK_K, (uint32_t)(vb::uses_pairs(ctx, R, V) ? vb::BN / 2 : vb::BN)};
    int rc = get_tmap_bf16(ctx, W, 2, gdim, gstr, box, &tb);
    if (rc) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (K <= 1) return vb::launch<2>(ctx, ta, tb, partials, nseg, R, V, d, s);
  if (K <= 3) return vb::launch<4>(ctx, ta, tb, partials, nseg, R, V, d, s);
  if (K <= 5) return vb::launch<6>(ctx, ta, tb, partials, nseg, R, V, d, s);
  return vb::launch<9>(ctx, ta, tb, partials, nseg, R, V, d, s);
}

#ifdef CARE_VB_TRACE
/* debug build only: copies the timeline stamps of the last vocab_beam_2sm_kernel launch ([128][5][64] int64) and clears them */
int care_debug_vb_trace(long long* out_host) {
  if (cudaMemcpyFromSymbol(out_host, vb::g_vb_trace, sizeof(vb::g_vb_trace)) != cudaSuccess) return -1;
  static long long zeros[128 * 5 * vb::VBT_EVENTS] = {0};
  return cudaMemcpyToSymbol(vb::g_vb_trace, zeros, sizeof(zeros)) == cudaSuccess ? 0 : -1;
}
#endif

}  // extern "C"
