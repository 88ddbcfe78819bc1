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

template <int KB>
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
      mbar_init(tempty_bar(a), EPI_WARPS);
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

  if (warp == 0) {
    if (lane == 0) {
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
          tma_load_2d(a_dst + cfg::A_BYTES, &tma_b, full_bar(s), kb * BLOCK_K, n_blk * BN);
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
    constexpr float LOG2E = 1.4426950408889634f;
    float run_m = -INFINITY, run_s = 0.f;
    float tv[KB];
    int ti[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      tv[q] = -INFINITY;
      ti[q] = INT_MAX;
    }
    uint32_t gcount = 0;
    for (int t = t_begin + grp; t < t_end; t += 2, ++gcount) {
      const int m_blk = t / n_tiles, n_blk = t - m_blk * n_tiles;
      const uint32_t aph = gcount & 1u;
      mbar_wait(tfull_bar(grp), aph);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + ew * 32 + lane;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n_blk * BN + c * 32;
        if (col0 >= N) break;
        uint32_t v[32];
        tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + grp * BN + c * 32, v);
        float x[32];
        const bool edge = col0 + 32 > N;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x[j] = __uint_as_float(v[j]);
          if (edge && col0 + j >= N) x[j] = -INFINITY;   // TMA zero-filled columns past V
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
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(grp));
      // this group's next tile is outside the m-block (or the run): flush the row record, start over
      const bool last = (t + 2 >= t_end) || ((t + 2) / n_tiles != m_blk);
      if (last) {
        if (row < M) {
          const int seg = 2 * ((int)blockIdx.x - run_of_tile((int64_t)m_blk * n_tiles, T, G)) + grp;
          float* rec = partials + ((int64_t)row * nseg + seg) * (2 + 2 * KB);
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

static int grid_for(const care_ctx* ctx, int R, int V) {
  const int64_t T = (int64_t)((R + BLOCK_M - 1) / BLOCK_M) * ((V + BN - 1) / BN);
  return (int)std::min<int64_t>(T, ctx->sm_count);
}

void seg_layout(const care_ctx* ctx, int R, int V, int* n_tiles, int64_t* T, int64_t* G) {
  *n_tiles = (V + BN - 1) / BN;
  *T = (int64_t)((R + BLOCK_M - 1) / BLOCK_M) * *n_tiles;
  *G = grid_for(ctx, R, V);
}

// largest number of runs that touch one m-block
int nseg_for(const care_ctx* ctx, int R, int V) {
  const int m_tiles = (R + BLOCK_M - 1) / BLOCK_M, n_tiles = (V + BN - 1) / BN;
  const int64_t T = (int64_t)m_tiles * n_tiles, G = grid_for(ctx, R, V);
  int best = 1;
  for (int m = 0; m < m_tiles; ++m) {
    const int c0 = run_of_tile((int64_t)m * n_tiles, T, G), c1 = run_of_tile((int64_t)m * n_tiles + n_tiles - 1, T, G);
    best = std::max(best, c1 - c0 + 1);
  }
  return 2 * best;   // two epilogue groups per run, one record each
}

template <int KB>
static int launch(care_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, float* partials, int nseg, int R, int V,
                  int d, cudaStream_t stream) {
  using cfg = Cfg<BN>;
  static bool configured_all[64] = {false};   // per device: function attributes are per device
  bool& configured = configured_all[ctx->device & 63];
  auto kern = vocab_beam_tcgen05_kernel<KB>;
  if (!configured) {
    CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
    configured = true;
  }
  kern<<<grid_for(ctx, R, V), VB_THREADS, cfg::SMEM_BYTES, stream>>>(ta, tb, partials, nseg, R, V, d,
                                                                             early_exit_of(ctx));
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
    const uint32_t box[2] = {(uint32_t)tc::BLOCK_K, (uint32_t)vb::BN};
    int rc = get_tmap_bf16(ctx, W, 2, gdim, gstr, box, &tb);
    if (rc) return rc;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (K <= 1) return vb::launch<2>(ctx, ta, tb, partials, nseg, R, V, d, s);
  if (K <= 3) return vb::launch<4>(ctx, ta, tb, partials, nseg, R, V, d, s);
  if (K <= 5) return vb::launch<6>(ctx, ta, tb, partials, nseg, R, V, d, s);
  return vb::launch<9>(ctx, ta, tb, partials, nseg, R, V, d, s);
}

}  // extern "C"
