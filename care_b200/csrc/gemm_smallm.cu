// bf16 GEMM for very few rows (M <= 16): C[M,N] = act(A[M,K] * W[N,K]^T + bias).
//
// Batch-1 / latency-mode decoding runs every Linear with M = beam size rows.  A 128-row tensor-core tile
// wastes >90 % of the MMA and, worse, leaves only N/64 CTAs to stream the weights, so each GEMM is bound by
// the latency of a handful of SMs walking K sequentially.  This kernel is weight-streaming bound instead:
// one CTA per 8 output columns (N/8 CTAs: 128 for d = 1024, 1844 for the vocabulary), 4 warps splitting K,
// every thread issuing 16-byte loads of W (and of the tiny, L2-resident A) and feeding them to
// mma.sync.m16n8k16.  The 8 consecutive K elements a thread loads are used as the k-fragments of two MMA
// steps for BOTH operands, i.e. the contraction index is permuted identically in A and B, which leaves the
// product unchanged and avoids any shared-memory transpose.  Partial sums of the 4 warps meet in smem.
#include "dev_util.cuh"

namespace care {
namespace smallm {

constexpr int WARPS = 4;
constexpr int COLS = 8;   // output columns (W rows) per CTA

__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  care::dev::mma_m16n8k16(c, a[0], a[1], a[2], a[3], b0, b1);
}


template <typename OutT>
__global__ void __launch_bounds__(WARPS * 32)
gemm_smallm_kernel(const h16* __restrict__ A, int64_t lda, const h16* __restrict__ W, int64_t ldw,
                   const float* __restrict__ bias, OutT* __restrict__ C, int64_t ldc, int M, int N, int n_store, int K,
                   int relu, const EarlyExit ee) {
  pdl_wait();
  pdl_launch_dependents();
  if (all_done(ee)) return;
  __shared__ float red[WARPS][16][COLS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int n0 = blockIdx.x * COLS;
  const int kw = K / WARPS;                 // K % 128 == 0 -> kw % 32 == 0
  const int k_begin = warp * kw;
  const bool row0 = g < M, row1 = g + 8 < M, wrow = n0 + g < N;
  const h16* a0p = A + (int64_t)(row0 ? g : 0) * lda + k_begin + 8 * tig;
  const h16* a1p = A + (int64_t)(row1 ? g + 8 : 0) * lda + k_begin + 8 * tig;
  const h16* wp = W + (int64_t)(wrow ? n0 + g : 0) * ldw + k_begin + 8 * tig;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int kc = 0; kc < kw; kc += 32) {
    const uint4 wa = wrow ? __ldg(reinterpret_cast<const uint4*>(wp + kc)) : zero;
    const uint4 x0 = row0 ? __ldg(reinterpret_cast<const uint4*>(a0p + kc)) : zero;
    const uint4 x1 = row1 ? __ldg(reinterpret_cast<const uint4*>(a1p + kc)) : zero;
    const uint32_t f0[4] = {x0.x, x1.x, x0.y, x1.y};
    const uint32_t f1[4] = {x0.z, x1.z, x0.w, x1.w};
    mma_bf16(c, f0, wa.x, wa.y);
    mma_bf16(c, f1, wa.z, wa.w);
  }
  red[warp][g][2 * tig] = c[0];
  red[warp][g][2 * tig + 1] = c[1];
  red[warp][g + 8][2 * tig] = c[2];
  red[warp][g + 8][2 * tig + 1] = c[3];
  __syncthreads();
  // 16 rows x 8 columns = 128 outputs, one per thread
  const int r = threadIdx.x >> 3, col = n0 + (threadIdx.x & 7);
  if (r < M && col < n_store) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) v += red[w][r][threadIdx.x & 7];
    if (bias != nullptr && col < N) v += __ldg(bias + col);
    if (relu) v = fmaxf(v, 0.f);
    if (col >= N) v = 0.f;
    C[(int64_t)r * ldc + col] = Act<OutT>::from_float(v);
  }
}

// returns 1 when the shape is not covered
int gemm_bf16_smallm(care_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
                     int64_t ldc, int out_dtype, int M, int N, int n_store, int K, int act, cudaStream_t stream) {
  if (M > 16 || K % (WARPS * 32) != 0 || lda % 8 != 0 || ldw % 8 != 0) return 1;
  const int grid = (n_store + COLS - 1) / COLS;
  const int relu = act == CARE_ACT_RELU ? 1 : 0;
  if (out_dtype == CARE_F32)
    CARE_CUDA(launch_pdl(ctx, gemm_smallm_kernel<float>, dim3(grid), dim3(WARPS * 32), 0, stream, static_cast<const h16*>(A),
                         lda, static_cast<const h16*>(W), ldw, bias, static_cast<float*>(C), ldc, M, N, n_store, K, relu,
                         early_exit_of(ctx)));
  else
    CARE_CUDA(launch_pdl(ctx, gemm_smallm_kernel<h16>, dim3(grid), dim3(WARPS * 32), 0, stream, static_cast<const h16*>(A),
                         lda, static_cast<const h16*>(W), ldw, bias, static_cast<h16*>(C), ldc, M, N, n_store, K, relu,
                         early_exit_of(ctx)));
  ctx->last_gemm = "gemm_smallm_kernel";
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace smallm
}  // namespace care
