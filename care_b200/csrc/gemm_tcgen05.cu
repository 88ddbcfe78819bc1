// bf16 GEMM on the 5th-gen tensor cores: C[M,N] = act(A[M,K] * W[N,K]^T + bias).
//
// Both operands are K-major (activations row-major [M,K], nn.Linear weights [N,K]), i.e. the
// canonical "TN" GEMM.  Persistent, warp-specialised, one CTA per SM:
//   warp 0  : TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B tiles, 4-8 stage mbarrier ring)
//   warp 1  : MMA issuer     (one lane issues tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16)
//   warp 2  : TMEM allocator (2 accumulator stages x BN fp32 columns)
//   warps 4-7 / 8-11: two epilogue groups taking alternate tiles (tcgen05.ld 32x32b -> swizzled smem
//                     transpose -> +bias/ReLU -> coalesced 16 B stores)
// The accumulator is double buffered in TMEM (one buffer per epilogue group) so the epilogue of tile i
// overlaps the MMAs of tiles i+1 and i+2.
// Edges: TMA zero-fills out-of-bounds rows/columns of A and W; stores are predicated.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <initializer_list>

#include "tcgen05_util.cuh"

namespace care {
namespace tc2 {  // gemm_tcgen05_2sm.cu: returns 1 when the shape should use the single-CTA kernel
int gemm_bf16_2sm(care_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                  int64_t ldc, int out_dtype, int M, int N, int n_store, int K, int act, cudaStream_t stream, int mc_pairs);
}
namespace smallm {  // gemm_smallm.cu: returns 1 when the shape is not covered (M > 16, ...)
int gemm_bf16_smallm(care_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                     int64_t ldc, int out_dtype, int M, int N, int n_store, int K, int act, cudaStream_t stream);
}
namespace tc {

constexpr int GEMM_EPI_BYTES = 8 * 4096;   // one 32x32 fp32 staging block per epilogue warp
// Tile widths in steps of 32 columns: with a few thousand rows (a 512-video shard has 20 row blocks) the
// number of 128 x BN tiles decides how many of the 148 SMs work, e.g. N = 1024: 256-wide tiles -> 80 tiles,
// 160-wide -> 140 tiles in one wave.
template <int BN>
struct GCfg {
  static_assert(BN % 32 == 0 && BN >= 64 && BN <= 256, "tile width: multiple of 32 in [64, 256]");
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BN * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;     // a multiple of 1024 (SWIZZLE_128B atom)
  static constexpr int TMEM_COLS = BN <= 64 ? 128 : (BN <= 128 ? 256 : 512);   // power of two, two accumulators
  static constexpr int ACC_STRIDE = TMEM_COLS / 2;
  static constexpr int MAX_STAGES = (227 * 1024 - GEMM_EPI_BYTES - 256 - 1024) / STAGE_BYTES;
  static constexpr int STAGES = MAX_STAGES > 8 ? 8 : MAX_STAGES;
};
template <int BN>
constexpr int gemm_smem_bytes() { return GCfg<BN>::STAGES * GCfg<BN>::STAGE_BYTES + GEMM_EPI_BYTES + 256 + 1024; }
constexpr int GEMM_THREADS = 384;   // warps 0-3: TMA / MMA / TMEM alloc / idle; warps 4-7, 8-11: two epilogue groups

template <int BN, typename OutT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const float* __restrict__ bias, OutT* __restrict__ C, int64_t ldc, int M, int N,
                         int n_store, int K, int relu, const EarlyExit ee) {
  if (all_done(ee)) return;   // uniform over the grid: written by an earlier kernel of the stream
  using cfg = GCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - raw_addr);
  const uint32_t epi_base = smem_base + cfg::STAGES * cfg::STAGE_BYTES;
  const uint32_t bar_base = epi_base + GEMM_EPI_BYTES;
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
  const int num_tiles = m_tiles * n_tiles;
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
  // everything above overlapped the previous kernel's tail (programmatic dependent launch); its outputs - this
  // kernel's A operand - are complete and visible only from here on
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      const uint64_t pol_w = l2_policy(ee.l2_hints & 1);
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;  // n fastest: an A row block is read once
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
      // ===== MMA issuer =====
      constexpr uint32_t idesc = instr_desc_bf16(BLOCK_M, BN);
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar(acc), aph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * cfg::ACC_STRIDE;
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % cfg::STAGES;
          const uint32_t ph = (it / cfg::STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t a_addr = smem_base + s * cfg::STAGE_BYTES;
          const uint64_t adesc = sw128_kmajor_desc(a_addr);
          const uint64_t bdesc = sw128_kmajor_desc(a_addr + cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 bytes (16 bf16) inside the 128-byte swizzle atom: +2 in the >>4 address field
            tc_mma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          tc_commit(empty_bar(s));  // frees the smem slot once these MMAs have read it
        }
        tc_commit(tfull_bar(acc));  // accumulator complete -> epilogue
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> (+bias, ReLU) -> global.  Two groups of 4 warps take alternate
    // tiles (group g drains TMEM accumulator g), so each group has two mainloop times per tile; a thread
    // owns one row and writes its 32 consecutive columns of a chunk as 16-byte stores. =====
    const int grp = (warp - 4) >> 2;
    const int ew = (warp - 4) & 3;  // == warp % 4: this warp may access TMEM lanes [32*ew, 32*ew+32)
    uint8_t* stage_gen = smem_gen + (epi_base - smem_base) + (warp - 4) * 4096;   // 32 rows x 128 B per warp
    uint32_t gcount = 0;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, ++gcount) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;  // n fastest: an A row block is read once
      mbar_wait(tfull_bar(grp), gcount & 1u);
      tc_fence_after();
      const int row_base = m_blk * BLOCK_M + ew * 32;
      if (row_base < M) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = n_blk * BN + c * 32;
          if (col0 >= n_store) break;
          uint32_t v[32];
          tmem_ld32(tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + grp * cfg::ACC_STRIDE + c * 32, v);
          __syncwarp();   // the previous chunk's staged rows have been read
          if constexpr (sizeof(OutT) == 4) {
            // stage the 32x32 fp32 block (row = lane) with a 16-byte XOR swizzle, read it back row-major:
            // every global store instruction then writes 4 full 128-byte rows
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(stage_gen + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
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
            // bf16 out: bias / ReLU in registers needs the column's bias per element -> do it after the
            // transpose too: stage fp32, read back 8 columns per thread, convert, one 16-byte store each
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<uint4*>(stage_gen + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                  make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            __syncwarp();
            const int c8 = lane & 3, col = col0 + c8 * 8;   // 8 consecutive columns = staged chunks 2*c8, 2*c8+1
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
      if (lane == 0) mbar_arrive(tempty_bar(grp));
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

static int get_tmap(care_ctx* ctx, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                    CUtensorMap* out) {
  const uint64_t gdim[2] = {cols, rows};
  const uint64_t gstride[1] = {ld * 2};
  const uint32_t box[2] = {(uint32_t)BLOCK_K, box_rows};
  return get_tmap_bf16(ctx, ptr, 2, gdim, gstride, box, out);
}

template <int BN, typename OutT>
static int launch(care_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const float* bias, void* C, int64_t ldc,
                  int M, int N, int n_store, int K, int act, cudaStream_t stream) {
  static bool configured_all[64] = {false};   // per device: function attributes are per device
  bool& configured = configured_all[ctx->device & 63];
  auto kern = gemm_bf16_tcgen05_kernel<BN, OutT>;
  if (!configured) {
    CARE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gemm_smem_bytes<BN>()));
    configured = true;
  }
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M, n_tiles = (N + BN - 1) / BN;
  const int grid = std::min(m_tiles * n_tiles, ctx->sm_count);
  CARE_CUDA(launch_pdl(ctx, kern, dim3(grid), dim3(GEMM_THREADS), gemm_smem_bytes<BN>(), stream, ta, tb, bias,
                       reinterpret_cast<OutT*>(C), ldc, M, N, n_store, K, act == CARE_ACT_RELU ? 1 : 0, early_exit_of(ctx)));
  ctx->last_gemm = sizeof(OutT) == 4 ? "gemm_bf16_tcgen05_kernel<float>" : "gemm_bf16_tcgen05_kernel<h16>";
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

int gemm_bf16(care_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
              int64_t ldc, int out_dtype, int M, int N, int K, int act, cudaStream_t stream) {
  CARE_CHECK_ARG(lda % 8 == 0 && ldw % 8 == 0, "care_gemm(bf16): lda/ldw must be multiples of 8 (got %lld, %lld)",
                 (long long)lda, (long long)ldw);
  CARE_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(C) & 15) == 0,
                 "care_gemm(bf16): A, W, C must be 16-byte aligned");
  CARE_CHECK_ARG(ldc % 8 == 0, "care_gemm(bf16): ldc must be a multiple of 8 (got %lld)", (long long)ldc);
  const int n_pad = (N + 7) & ~7;
  CARE_CHECK_ARG(n_pad <= ldc || N % 8 == 0, "care_gemm(bf16): ldc %lld too small for N=%d", (long long)ldc, N);
  const int n_store = n_pad <= ldc ? n_pad : N;
  if (ctx->gemm_smallm) {   // latency mode: a handful of rows -> weight-streaming kernel
    const int rcs = smallm::gemm_bf16_smallm(ctx, A, lda, W, ldw, bias, C, ldc, out_dtype, M, N, n_store, K, act, stream);
    if (rcs != 1) return rcs;
  }
  int use_2sm = ctx->gemm_2sm;
  if (use_2sm == 2) {
    // per-shape choice between the CTA-pair kernel and single-CTA tiles, measured once
    const uint64_t key = ((uint64_t)(uint32_t)M << 40) ^ ((uint64_t)(uint32_t)N << 20) ^ ((uint64_t)(uint32_t)K << 1) ^
                         (uint64_t)(out_dtype & 1);
    int choice = -1;
    {
      std::lock_guard<std::mutex> g(ctx->tuning->mu);
      auto it = ctx->tuning->choice.find(key);
      if (it != ctx->tuning->choice.end()) choice = it->second;
    }
    if (choice < 0) {
      cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(stream, &cap);
      if (cap != cudaStreamCaptureStatusNone) {
        choice = 0;   // cannot time inside a capture; not cached
      } else {
        // variants: 0 = single-CTA tiles, 1 = CTA pairs, 2 = clusters of 4 pairs with the A tile multicast, 3 = of 2 pairs
        constexpr int NV = 4;
        static const int mc_of[NV] = {0, 0, 4, 2};
        float best_ms[NV] = {0.f, 0.f, 0.f, 0.f};
        bool ok[NV] = {true, true, true, true};
        cudaEvent_t e0, e1;
        CARE_CUDA(cudaEventCreate(&e0));
        CARE_CUDA(cudaEventCreate(&e1));
        const int saved = ctx->gemm_2sm;
        for (int v = 0; v < NV; ++v) {
          if (v >= 2 && !ok[1]) {   // the shape does not suit pair tiles at all
            ok[v] = false;
            continue;
          }
          ctx->gemm_2sm = 0;
          for (int rep = 0; rep < 4 && ok[v]; ++rep) {   // rep 0 = warm-up
            if (rep == 1) cudaEventRecord(e0, stream);
            int rc = 0;
            if (v >= 1) {
              rc = tc2::gemm_bf16_2sm(ctx, A, lda, W, ldw, bias, C, ldc, out_dtype, M, N, n_store, K, act, stream, mc_of[v]);
              if (rc == 1) ok[v] = false;
            } else {
              rc = gemm_bf16(ctx, A, lda, W, ldw, bias, C, ldc, out_dtype, M, N, K, act, stream);
            }
            if (rc != 0 && rc != 1) {
              ctx->gemm_2sm = saved;
              cudaEventDestroy(e0);
              cudaEventDestroy(e1);
              return rc;
            }
          }
          if (!ok[v]) continue;
          cudaEventRecord(e1, stream);
          cudaEventSynchronize(e1);
          cudaEventElapsedTime(&best_ms[v], e0, e1);
        }
        ctx->gemm_2sm = saved;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        choice = 0;
        for (int v = 1; v < NV; ++v)
          if (ok[v] && best_ms[v] < best_ms[choice]) choice = v;
        if (ctx->debug)
          fprintf(stderr, "[care_b200] gemm M=%d N=%d K=%d out=%s: single-CTA %.3f ms, CTA-pair %.3f, 4-pair clusters %.3f, "
                  "2-pair clusters %.3f (0 = n/a) -> variant %d\n", M, N, K, out_dtype == CARE_F32 ? "f32" : "bf16",
                  best_ms[0] / 3.0f, best_ms[1] / 3.0f, best_ms[2] / 3.0f, best_ms[3] / 3.0f, choice);
        {
          std::lock_guard<std::mutex> g(ctx->tuning->mu);
          ctx->tuning->choice[key] = choice;
        }
        // profiling runs replay the choices of an unprofiled run (timings taken under ncu are not representative):
        // CARE_B200_GEMM_CHOICE_FILE names a file that receives "key choice" lines and is read back by care_ctx_create
        if (const char* path = getenv("CARE_B200_GEMM_CHOICE_FILE")) {
          if (FILE* f = fopen(path, "a")) {
            fprintf(f, "%llu %d\n", (unsigned long long)key, choice);
            fclose(f);
          }
        }
        // C holds the result of the variant timed last; fall through so that THIS call, like every later
        // one, returns the chosen kernel's output (the two variants may differ in the last bit)
      }
    }
    use_2sm = choice;
  } else if (use_2sm == 4 || use_2sm == 5) {   // forced: clusters of 4 / 2 pairs (variant codes 2 / 3)
    use_2sm -= 2;
  }
  if (use_2sm >= 1) {   // 1: CTA pairs, 2 / 3: clusters of 4 / 2 pairs with the A tile multicast
    const int rc2 = tc2::gemm_bf16_2sm(ctx, A, lda, W, ldw, bias, C, ldc, out_dtype, M, N, n_store, K, act, stream,
                                       use_2sm == 2 ? 4 : (use_2sm == 3 ? 2 : 0));
    if (rc2 != 1) return rc2;
    if (use_2sm > 1) {   // the cluster does not fit: plain pairs
      const int rc3 = tc2::gemm_bf16_2sm(ctx, A, lda, W, ldw, bias, C, ldc, out_dtype, M, N, n_store, K, act, stream, 0);
      if (rc3 != 1) return rc3;
    }
  }
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  auto tiles = [&](int bn) { return m_tiles * ((N + bn - 1) / bn); };
  // tile width: fewest (waves x per-tile cost); per-tile cost ~ BN + a fixed part (A tile load, epilogue
  // set-up), so narrower tiles win when 256-wide tiles would leave most SMs idle in the last wave
  int bn = 256;
  int64_t best = INT64_MAX;
  for (int cand : {256, 224, 192, 160, 128, 96, 64}) {
    const int64_t waves = (tiles(cand) + ctx->sm_count - 1) / ctx->sm_count;
    const int64_t cost = waves * (cand + 48);
    if (cost < best) {
      best = cost;
      bn = cand;
    }
  }
  if (ctx->gemm_bn > 0) bn = ctx->gemm_bn;   // A/B runs
  CUtensorMap ta, tb;
  int rc = get_tmap(ctx, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BLOCK_M, &ta);
  if (rc) return rc;
  rc = get_tmap(ctx, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)bn, &tb);
  if (rc) return rc;
#define CARE_TC_DISPATCH(BN_)                                                                          \
  (out_dtype == CARE_F32 ? launch<BN_, float>(ctx, ta, tb, bias, C, ldc, M, N, n_store, K, act, stream) \
                         : launch<BN_, h16>(ctx, ta, tb, bias, C, ldc, M, N, n_store, K, act, stream))
  switch (bn) {
    case 256: return CARE_TC_DISPATCH(256);
    case 224: return CARE_TC_DISPATCH(224);
    case 192: return CARE_TC_DISPATCH(192);
    case 160: return CARE_TC_DISPATCH(160);
    case 128: return CARE_TC_DISPATCH(128);
    case 96: return CARE_TC_DISPATCH(96);
    default: return CARE_TC_DISPATCH(64);
  }
#undef CARE_TC_DISPATCH
}

}  // namespace tc
}  // namespace care
