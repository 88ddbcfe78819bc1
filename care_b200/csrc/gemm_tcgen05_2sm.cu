// bf16 GEMM on CTA pairs: C[M,N] = act(A[M,K] * W[N,K]^T + bias) with tcgen05.mma.cta_group::2.
//
// The single-CTA kernel (gemm_tcgen05.cu) moves (128 + 256) x 64 bf16 = 48 KB from L2 into an SM per
// 4 MMAs; at the tensor rate that is ~92 B/clk per SM, about what the L2->SM path can deliver at all,
// so it saturates near 1.3 PFLOP/s (profiles/).  Here the two SMs of a TPC form a cluster and share
// the operands of one 256 x 256 tile: each CTA loads its own 128 rows of A and only HALF of the B tile
// (128 of the 256 W rows), 32 KB per k-block, and the leader's single thread issues
// tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared memory and writes each CTA's
// 128 x 256 half of the accumulator into that CTA's own TMEM.
//
// Roles per CTA (384 threads): warp 0 TMA producer (both CTAs; their transactions complete on the
// LEADER's full barrier), warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator
// (tcgen05.alloc.cta_group::2, both CTAs), warps 4-7 / 8-11 two epilogue groups taking alternate tiles.
// Barriers: full[s] (leader; count 1 + 64 KB of transactions from both CTAs), empty[s] (both CTAs;
// released by a multicast tcgen05.commit), tfull[a] (both CTAs; multicast commit), tempty[a] (leader;
// 8 arrivals: 4 epilogue warps of each CTA, the peer's arrive remotely).
#include "tcgen05_util.cuh"

namespace care {
namespace tc2 {

using namespace care::tc;

constexpr int BN = 256;                 // tile width (each CTA stages BN/2 rows of W)
constexpr int PAIR_M = 2 * BLOCK_M;     // 256 rows per CTA pair
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB
constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;       // 16 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;        // 32 KB per CTA
constexpr int STAGES = 6;
constexpr int EPI_BYTES = 8 * 4096;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 + 1024;
constexpr int THREADS = 384;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address -> leader CTA

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose transaction bytes complete on the leader CTA's barrier (same offset, rank bit cleared)
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
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void commit_2sm_multicast(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar & PEER_MASK) : "memory");
}

// multicast forms for clusters of several pairs (gemm_bf16_2sm_mc_kernel)
__device__ __forceinline__ void tma_load_2d_2sm_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                   uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, "
      "{%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & PEER_MASK), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void commit_2sm_mask(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}

// The same GEMM on clusters of n_pairs CTA pairs (cluster size 4 or 8) that work on n_pairs neighbouring n-tiles of ONE
// 256-row block: the A tile is the same for all of them, so CTA (pair p, row half h) fetches only rows
// [h * 128 + p * 128 / n_pairs, ...) of it and multicasts them to the CTAs of equal h (4 KB + 16 KB of W from L2 per k-block
// and CTA instead of 32 KB at n_pairs = 4).  Why: the pair kernel's mainloop needs ~690 cycles per k-block where its MMAs take
// 512 - 148 SMs x 64 B/clk is more than the L2 delivers chip-wide.  A stage is refilled when ALL
// pairs' MMAs have read it (empty barriers count n_pairs multicast commits); accumulator barriers stay per pair.

template <typename OutT>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_2sm_mc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const float* __restrict__ bias, OutT* __restrict__ C, int64_t ldc, int M, int N, int n_store,
                     int K, int relu, const EarlyExit ee) {
  if (all_done(ee)) return;   // uniform over the grid
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank(), csize = cluster_nctarank();
  const uint32_t rank = crank & 1u;              // 0 = leader of its pair
  const int pair = (int)(crank >> 1), n_pairs = (int)(csize >> 1);
  const int cluster_id = blockIdx.x / (int)csize;
  const int n_clusters = gridDim.x / (int)csize;
  const int m_pairs = (M + PAIR_M - 1) / PAIR_M;
  const int n_tiles = (N + BN - 1) / BN;
  const int n_groups = (n_tiles + n_pairs - 1) / n_pairs;   // a cluster works on n_pairs neighbouring n-tiles of one m-block
  const int num_tiles = m_pairs * n_groups;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  const uint16_t pair_mask = (uint16_t)(3u << (crank & ~1u));
  const uint16_t all_mask = (uint16_t)((1u << csize) - 1u);
  // the CTAs that hold the same 128 rows of A: this CTA's row half in every pair
  uint16_t half_mask = 0;
  for (int q = 0; q < n_pairs; ++q) half_mask |= (uint16_t)(1u << (2 * q + (int)rank));
  const int slice_rows = BLOCK_M / n_pairs;      // this CTA fetches rows [pair * slice_rows, +slice_rows) of them

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tma_b)) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), (uint32_t)n_pairs);   // every pair's MMAs have read the stage (its A slices come from all pairs)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 2 * EPI_WARPS);   // 4 epilogue warps of each CTA of the pair
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
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();   // prologue done under the previous kernel's tail; its outputs are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own 128 rows of A, own half of the W tile =====
      const uint64_t pol_w = l2_policy(ee.l2_hints & 1);
      uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m_pair = tile / n_groups, n_blk = (tile % n_groups) * n_pairs + pair;
        const int a_row = m_pair * PAIR_M + (int)rank * BLOCK_M + pair * slice_rows;
        const int b_row = n_blk * BN + (int)rank * (BN / 2);   // past N for a pair without a tile: zero-filled
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);   // both CTAs' bytes land on the leader
          const uint32_t a_dst = smem_base + s * STAGE_BYTES;
          tma_load_2d_2sm_mc(a_dst + (uint32_t)(pair * slice_rows * BLOCK_K * 2), &tma_a, full_bar(s), kb * BLOCK_K, a_row,
                             half_mask);
          tma_load_2d_cta2_hint(a_dst + A_BYTES, &tma_b, full_bar(s) & PEER_MASK, kb * BLOCK_K, b_row, pol_w);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader only): M = 256 across the pair, N = 256, K = 16 =====
      constexpr uint32_t idesc = instr_desc_bf16(PAIR_M, BN);
      uint32_t it = 0, tcount = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++tcount) {
        const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar(acc), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * STAGE_BYTES;
          const uint64_t adesc = sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            mma_bf16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          commit_2sm_mask(empty_bar(s), all_mask);   // one of n_pairs arrivals on the stage's barrier in every CTA
        }
        commit_2sm_mask(tfull_bar(acc), pair_mask);   // both CTAs' accumulator halves complete
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the pair tile =====
    const int grp = (warp - 4) >> 2;
    const int ew = (warp - 4) & 3;
    uint8_t* stage_gen = smem_gen + (epi_base - smem_base) + (warp - 4) * 4096;
    uint32_t gcount = 0;
    for (int tile = cluster_id + grp * n_clusters; tile < num_tiles; tile += 2 * n_clusters, ++gcount) {
      const int m_pair = tile / n_groups, n_blk = (tile % n_groups) * n_pairs + pair;
      mbar_wait(tfull_bar(grp), gcount & 1u);
      tc_fence_after();
      const int row_base = m_pair * PAIR_M + (int)rank * BLOCK_M + ew * 32;
      if (row_base < M) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n_blk * BN + c * 32;
          if (col0 >= n_store) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + grp * BN + c * 32, v);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stage_gen + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          if constexpr (sizeof(OutT) == 4) {
            const int ch = lane & 7, col = col0 + ch * 4;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias != nullptr) {
              if (col + 3 < N) {
                bv = __ldg(reinterpret_cast<const float4*>(bias + col));
              } else {
                if (col < N) bv.x = __ldg(bias + col);
                if (col + 1 < N) bv.y = __ldg(bias + col + 1);
                if (col + 2 < N) bv.z = __ldg(bias + col + 2);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3);
              float4 val = *reinterpret_cast<const float4*>(stage_gen + r * 128 + ((ch ^ (r & 7)) << 4));
              val.x += bv.x; val.y += bv.y; val.z += bv.z; val.w += bv.w;
              if (relu) {
                val.x = fmaxf(val.x, 0.f); val.y = fmaxf(val.y, 0.f);
                val.z = fmaxf(val.z, 0.f); val.w = fmaxf(val.w, 0.f);
              }
              const int row = row_base + r;
              if (row < M && col < n_store)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(C) + static_cast<int64_t>(row) * ldc + col) = val;
            }
          } else {
            const int c8 = lane & 3, col = col0 + c8 * 8;
            float bb[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) bb[q] = (bias != nullptr && col + q < N) ? __ldg(bias + col + q) : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 8 + (lane >> 2);
              const float4 lo = *reinterpret_cast<const float4*>(stage_gen + r * 128 + (((2 * c8) ^ (r & 7)) << 4));
              const float4 hi = *reinterpret_cast<const float4*>(stage_gen + r * 128 + (((2 * c8 + 1) ^ (r & 7)) << 4));
              float f[8] = {lo.x + bb[0], lo.y + bb[1], lo.z + bb[2], lo.w + bb[3],
                            hi.x + bb[4], hi.y + bb[5], hi.z + bb[6], hi.w + bb[7]};
              if (relu) {
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] = fmaxf(f[q], 0.f);
              }
              uint4 pk;
              h162* h = reinterpret_cast<h162*>(&pk);
#pragma unroll
              for (int q = 0; q < 4; ++q) h[q] = floats_to_h162(f[2 * q], f[2 * q + 1]);
              const int row = row_base + r;
              if (row < M && col < n_store)
                *reinterpret_cast<uint4*>(reinterpret_cast<h16*>(C) + static_cast<int64_t>(row) * ldc + col) = pk;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(grp));   // local for the leader, remote for the peer
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody may still be reading a peer's shared memory / barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(2 * BN))
                 : "memory");
  }
}


template <typename OutT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_bf16_2sm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                     const float* __restrict__ bias, OutT* __restrict__ C, int64_t ldc, int M, int N, int n_store,
                     int K, int relu, const EarlyExit ee) {
  if (all_done(ee)) return;   // uniform over the grid
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;
  const uint32_t bar_base = epi_base + EPI_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 = leader
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int m_pairs = (M + PAIR_M - 1) / PAIR_M;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_pairs * n_tiles;
  const int k_blocks = (K + BLOCK_K - 1) / BLOCK_K;

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
      mbar_init(tempty_bar(a), 2 * EPI_WARPS);   // 4 epilogue warps of each CTA of the pair
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
  cluster_sync_all();   // barriers of both CTAs initialised, TMEM allocated in both
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  pdl_wait();   // prologue done under the previous kernel's tail; its outputs are visible from here on
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): own 128 rows of A, own half of the W tile =====
      const uint64_t pol_w = l2_policy(ee.l2_hints & 1);
      uint32_t it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
        const int a_row = m_pair * PAIR_M + (int)rank * BLOCK_M;
        const int b_row = n_blk * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          if (rank == 0) mbar_expect_tx(full_bar(s), 2 * STAGE_BYTES);   // both CTAs' bytes land on the leader
          const uint32_t a_dst = smem_base + s * STAGE_BYTES;
          tma_load_2d_2sm(a_dst, &tma_a, full_bar(s), kb * BLOCK_K, a_row);
          tma_load_2d_cta2_hint(a_dst + A_BYTES, &tma_b, full_bar(s) & PEER_MASK, kb * BLOCK_K, b_row, pol_w);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader only): M = 256 across the pair, N = 256, K = 16 =====
      constexpr uint32_t idesc = instr_desc_bf16(PAIR_M, BN);
      uint32_t it = 0, tcount = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters, ++tcount) {
        const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar(acc), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * STAGE_BYTES;
          const uint64_t adesc = sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = sw128_kmajor_desc(a_addr + A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            mma_bf16_2sm(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          commit_2sm_multicast(empty_bar(s));   // frees the stage in both CTAs
        }
        commit_2sm_multicast(tfull_bar(acc));   // both CTAs' accumulator halves complete
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the pair tile =====
    const int grp = (warp - 4) >> 2;
    const int ew = (warp - 4) & 3;
    uint8_t* stage_gen = smem_gen + (epi_base - smem_base) + (warp - 4) * 4096;
    uint32_t gcount = 0;
    for (int tile = cluster_id + grp * n_clusters; tile < num_tiles; tile += 2 * n_clusters, ++gcount) {
      const int m_pair = tile / n_tiles, n_blk = tile % n_tiles;
      mbar_wait(tfull_bar(grp), gcount & 1u);
      tc_fence_after();
      const int row_base = m_pair * PAIR_M + (int)rank * BLOCK_M + ew * 32;
      if (row_base < M) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n_blk * BN + c * 32;
          if (col0 >= n_store) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + grp * BN + c * 32, v);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(stage_gen + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          __syncwarp();
          if constexpr (sizeof(OutT) == 4) {
            const int ch = lane & 7, col = col0 + ch * 4;
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (bias != nullptr) {
              if (col + 3 < N) {
                bv = __ldg(reinterpret_cast<const float4*>(bias + col));
              } else {
                if (col < N) bv.x = __ldg(bias + col);
                if (col + 1 < N) bv.y = __ldg(bias + col + 1);
                if (col + 2 < N) bv.z = __ldg(bias + col + 2);
              }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3);
              float4 val = *reinterpret_cast<const float4*>(stage_gen + r * 128 + ((ch ^ (r & 7)) << 4));
              val.x += bv.x; val.y += bv.y; val.z += bv.z; val.w += bv.w;
              if (relu) {
                val.x = fmaxf(val.x, 0.f); val.y = fmaxf(val.y, 0.f);
                val.z = fmaxf(val.z, 0.f); val.w = fmaxf(val.w, 0.f);
              }
              const int row = row_base + r;
              if (row < M && col < n_store)
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(C) + static_cast<int64_t>(row) * ldc + col) = val;
            }
          } else {
            const int c8 = lane & 3, col = col0 + c8 * 8;
            float bb[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) bb[q] = (bias != nullptr && col + q < N) ? __ldg(bias + col + q) : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 8 + (lane >> 2);
              const float4 lo = *reinterpret_cast<const float4*>(stage_gen + r * 128 + (((2 * c8) ^ (r & 7)) << 4));
              const float4 hi = *reinterpret_cast<const float4*>(stage_gen + r * 128 + (((2 * c8 + 1) ^ (r & 7)) << 4));
              float f[8] = {lo.x + bb[0], lo.y + bb[1], lo.z + bb[2], lo.w + bb[3],
                            hi.x + bb[4], hi.y + bb[5], hi.z + bb[6], hi.w + bb[7]};
              if (relu) {
#pragma unroll
                for (int q = 0; q < 8; ++q) f[q] = fmaxf(f[q], 0.f);
              }
              uint4 pk;
              h162* h = reinterpret_cast<h162*>(&pk);
#pragma unroll
              for (int q = 0; q < 4; ++q) h[q] = floats_to_h162(f[2 * q], f[2 * q + 1]);
              const int row = row_base + r;
              if (row < M && col < n_store)
                *reinterpret_cast<uint4*>(reinterpret_cast<h16*>(C) + static_cast<int64_t>(row) * ldc + col) = pk;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(grp));   // local for the leader, remote for the peer
    }
  }

  tc_fence_before();
  cluster_sync_all();   // nobody may still be reading a peer's shared memory / barriers
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(2 * BN))
                 : "memory");
  }
}

template <typename OutT>
static int launch(care_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const float* bias, void* C, int64_t ldc,
                  int M, int N, int n_store, int K, int act, cudaStream_t stream) {
  static bool configured_all[64] = {false};   // per device: function attributes are per device
  bool& configured = configured_all[ctx->device & 63];
  auto kern = gemm_bf16_2sm_kernel<OutT>;
  if (!configured) {
    CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((M + PAIR_M - 1) / PAIR_M) * ((N + BN - 1) / BN);
  const int n_clusters = std::min(tiles, ctx->sm_count / 2);
  CARE_CUDA(launch_pdl(ctx, kern, dim3(2 * n_clusters), dim3(THREADS), SMEM_BYTES, stream, ta, tb, bias,
                       reinterpret_cast<OutT*>(C), ldc, M, N, n_store, K, act == CARE_ACT_RELU ? 1 : 0, early_exit_of(ctx)));
  ctx->last_gemm = sizeof(OutT) == 4 ? "gemm_bf16_2sm_kernel<float>" : "gemm_bf16_2sm_kernel<h16>";
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

// clusters of n_pairs pairs (2 or 4) sharing the A tile by multicast; returns 1 when such a cluster does not fit the device
template <typename OutT>
static int launch_mc(care_ctx* ctx, int n_pairs, const CUtensorMap& ta, const CUtensorMap& tb, const float* bias, void* C,
                     int64_t ldc, int M, int N, int n_store, int K, int act, cudaStream_t stream) {
  static bool configured_all[64] = {false};
  static int max_clusters_all[64][5] = {{0}};   // [device][n_pairs]; -1: does not fit
  auto kern = gemm_bf16_2sm_mc_kernel<OutT>;
  if (!configured_all[ctx->device & 63]) {
    CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured_all[ctx->device & 63] = true;
  }
  const int csize = 2 * n_pairs;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.blockDim = dim3(THREADS, 1, 1);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 2 : 1;
  int& max_clusters = max_clusters_all[ctx->device & 63][n_pairs];
  if (max_clusters == 0) {
    cfg.gridDim = dim3((unsigned)(csize * (ctx->sm_count / csize)), 1, 1);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = -1;
    }
    max_clusters = n;
    if (ctx->debug) fprintf(stderr, "[care_b200] gemm_bf16_2sm_mc: %d clusters of %d CTAs fit\n", n, csize);
  }
  if (max_clusters < 0) return 1;
  const int n_tiles = (N + BN - 1) / BN;
  const int supers = ((M + PAIR_M - 1) / PAIR_M) * ((n_tiles + n_pairs - 1) / n_pairs);
  const int n_clusters = std::min(supers, max_clusters);
  cfg.gridDim = dim3((unsigned)(csize * n_clusters), 1, 1);
  CARE_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, bias, reinterpret_cast<OutT*>(C), ldc, M, N, n_store, K,
                               act == CARE_ACT_RELU ? 1 : 0, early_exit_of(ctx)));
  ctx->last_gemm = sizeof(OutT) == 4 ? "gemm_bf16_2sm_mc_kernel<float>" : "gemm_bf16_2sm_mc_kernel<h16>";
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

// returns 1 when the shape is better served by the single-CTA kernel.  mc_pairs: 0 = one pair per cluster, 2 / 4 = clusters
// of that many pairs with the A tile multicast
int gemm_bf16_2sm(care_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                  int64_t ldc, int out_dtype, int M, int N, int n_store, int K, int act, cudaStream_t stream, int mc_pairs) {
  const int tiles = ((M + PAIR_M - 1) / PAIR_M) * ((N + BN - 1) / BN);
  if (tiles < 2 * (ctx->sm_count / 2)) return 1;   // fewer than two waves of pair tiles: narrow 1-SM tiles balance better
  CUtensorMap ta, tb;
  {
    const uint64_t gdim[2] = {(uint64_t)K, (uint64_t)M};
    const uint64_t gstr[1] = {(uint64_t)lda * 2};
    const uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)(mc_pairs ? BLOCK_M / mc_pairs : BLOCK_M)};
    int rc = get_tmap_bf16(ctx, A, 2, gdim, gstr, box, &ta);
    if (rc) return rc;
  }
  {
    const uint64_t gdim[2] = {(uint64_t)K, (uint64_t)N};
    const uint64_t gstr[1] = {(uint64_t)ldw * 2};
    const uint32_t box[2] = {(uint32_t)BLOCK_K, (uint32_t)(BN / 2)};
    int rc = get_tmap_bf16(ctx, W, 2, gdim, gstr, box, &tb);
    if (rc) return rc;
  }
  if (mc_pairs) {
    if (out_dtype == CARE_F32) return launch_mc<float>(ctx, mc_pairs, ta, tb, bias, C, ldc, M, N, n_store, K, act, stream);
    return launch_mc<h16>(ctx, mc_pairs, ta, tb, bias, C, ldc, M, N, n_store, K, act, stream);
  }
  if (out_dtype == CARE_F32) return launch<float>(ctx, ta, tb, bias, C, ldc, M, N, n_store, K, act, stream);
  return launch<h16>(ctx, ta, tb, bias, C, ldc, M, N, n_store, K, act, stream);
}

}  // namespace tc2
}  // namespace care
