// fp32 parity-mode GEMM: C[M,N] = act(A[M,K] * W[N,K]^T + bias), plain FFMA accumulation in fp32
// (tcgen05 has no true-fp32 MMA; the bit-exact token parity mode must not round operands to tf32).
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, k-major smem tiles (transposed on
// the way in so the inner loop reads conflict-free float4), register double buffering.
#include "common.cuh"

namespace care {
namespace f32 {

constexpr int BM = 128, BN = 128, BK = 16, TM = 8, TN = 8, THREADS = 256;

template <typename OutT>
__global__ void __launch_bounds__(THREADS)
gemm_f32_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                const float* __restrict__ bias, OutT* __restrict__ C, int64_t ldc, int M, int N, int n_store, int K,
                int relu, const EarlyExit ee) {
  if (all_done(ee)) return;
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // global->smem mapping: each thread moves 2 float4 of A and 2 of W per k-tile
  const int lrow = tid >> 2;        // 0..63
  const int lk = (tid & 3) * 4;     // 0,4,8,12
  const int tx = tid & 15, ty = tid >> 4;  // 16x16 thread grid, each 8x8 outputs

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[2], rw[2];
  auto gload = [&](int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      const int gm = m0 + r, gn = n0 + r, gk = k0 + lk;
      ra[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      rw[h] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gm < M && gk < K) ra[h] = *reinterpret_cast<const float4*>(A + (int64_t)gm * lda + gk);
      if (gn < N && gk < K) rw[h] = *reinterpret_cast<const float4*>(W + (int64_t)gn * ldw + gk);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = lrow + h * 64;
      As[buf][lk + 0][r] = ra[h].x; As[buf][lk + 1][r] = ra[h].y; As[buf][lk + 2][r] = ra[h].z; As[buf][lk + 3][r] = ra[h].w;
      Ws[buf][lk + 0][r] = rw[h].x; Ws[buf][lk + 1][r] = rw[h].y; Ws[buf][lk + 2][r] = rw[h].z; Ws[buf][lk + 3][r] = rw[h].w;
    }
  };

  const int k_tiles = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < k_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < k_tiles) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 w0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float4 w1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < k_tiles) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows {ty*4+i, 64+ty*4+i}, cols {tx*4+j, 64+tx*4+j}
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + jh * 64 + tx * 4;
      if (col >= n_store) continue;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float x = acc[i][jh * 4 + j];
        if (bias != nullptr && col + j < N) x += __ldg(bias + col + j);
        if (col + j >= N) x = 0.f;
        o[j] = relu ? fmaxf(x, 0.f) : x;
      }
      Act<OutT>::store4(C + (int64_t)row * ldc + col, o);
    }
  }
}

// Same product on 64x64 tiles (4x4 outputs per thread) for shapes whose 128x128 tiling would leave most SMs idle
// (the per-video heads at a few hundred videos).  Every output still accumulates k = 0..K-1 in order with fmaf,
// so the result is bit-identical to the large-tile kernel's.
constexpr int SB = 64, ST = 4;
template <typename OutT>
__global__ void __launch_bounds__(THREADS)
gemm_f32_small_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                      const float* __restrict__ bias, OutT* __restrict__ C, int64_t ldc, int M, int N, int n_store,
                      int K, int relu, const EarlyExit ee) {
  if (all_done(ee)) return;
  __shared__ __align__(16) float As[2][BK][SB + 4];
  __shared__ __align__(16) float Ws[2][BK][SB + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SB, n0 = blockIdx.x * SB;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;   // one float4 of A and one of W per thread per k-tile
  const int tx = tid & 15, ty = tid >> 4;
  float acc[ST][ST];
#pragma unroll
  for (int i = 0; i < ST; ++i)
#pragma unroll
    for (int j = 0; j < ST; ++j) acc[i][j] = 0.f;
  float4 ra, rw;
  auto gload = [&](int k0) {
    const int gm = m0 + lrow, gn = n0 + lrow, gk = k0 + lk;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gm < M && gk < K) ra = *reinterpret_cast<const float4*>(A + (int64_t)gm * lda + gk);
    if (gn < N && gk < K) rw = *reinterpret_cast<const float4*>(W + (int64_t)gn * ldw + gk);
  };
  auto sstore = [&](int buf) {
    As[buf][lk + 0][lrow] = ra.x; As[buf][lk + 1][lrow] = ra.y; As[buf][lk + 2][lrow] = ra.z; As[buf][lk + 3][lrow] = ra.w;
    Ws[buf][lk + 0][lrow] = rw.x; Ws[buf][lk + 1][lrow] = rw.y; Ws[buf][lk + 2][lrow] = rw.z; Ws[buf][lk + 3][lrow] = rw.w;
  };
  const int k_tiles = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < k_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < k_tiles) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < ST; ++i)
#pragma unroll
        for (int j = 0; j < ST; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kt + 1 < k_tiles) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const int col = n0 + tx * 4;
  if (col >= n_store) return;
#pragma unroll
  for (int i = 0; i < ST; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float x = acc[i][j];
      if (bias != nullptr && col + j < N) x += __ldg(bias + col + j);
      if (col + j >= N) x = 0.f;
      o[j] = relu ? fmaxf(x, 0.f) : x;
    }
    Act<OutT>::store4(C + (int64_t)row * ldc + col, o);
  }
}

// Same product on 32x32 tiles (2x2 outputs per thread): a few hundred rows of a ranking head (512 videos x 500
// concepts, K = 4096) are only 64 tiles of 64x64 - 64 CTAs walking 256 k-tiles one after the other on 148 SMs, bound by
// the latency of each k-tile, not by FFMA issue.  256 tiles put several CTAs on every SM.  Same k order per output:
// bit-identical to the other two kernels.
constexpr int TB = 32;
template <typename OutT>
__global__ void __launch_bounds__(THREADS)
gemm_f32_tiny_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ W, int64_t ldw,
                     const float* __restrict__ bias, OutT* __restrict__ C, int64_t ldc, int M, int N, int n_store, int K,
                     int relu, const EarlyExit ee) {
  if (all_done(ee)) return;
  __shared__ __align__(16) float As[2][BK][TB + 4];
  __shared__ __align__(16) float Ws[2][BK][TB + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * TB, n0 = blockIdx.x * TB;
  const bool loads_a = tid < 128;                 // threads 0..127 move A's k-tile, 128..255 W's: one float4 each
  const int lrow = (tid & 127) >> 2, lk = (tid & 3) * 4;
  const int tx = tid & 15, ty = tid >> 4;         // 16 x 16 threads, 2 x 2 outputs each
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float4 r4;
  auto gload = [&](int k0) {
    const int gk = k0 + lk;
    r4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (loads_a) {
      const int gm = m0 + lrow;
      if (gm < M && gk < K) r4 = *reinterpret_cast<const float4*>(A + (int64_t)gm * lda + gk);
    } else {
      const int gn = n0 + lrow;
      if (gn < N && gk < K) r4 = *reinterpret_cast<const float4*>(W + (int64_t)gn * ldw + gk);
    }
  };
  auto sstore = [&](int buf) {
    float(*dst)[TB + 4] = loads_a ? As[buf] : Ws[buf];
    dst[lk + 0][lrow] = r4.x; dst[lk + 1][lrow] = r4.y; dst[lk + 2][lrow] = r4.z; dst[lk + 3][lrow] = r4.w;
  };
  const int k_tiles = (K + BK - 1) / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < k_tiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < k_tiles) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float2 a2 = *reinterpret_cast<const float2*>(&As[buf][k][ty * 2]);
      const float2 w2 = *reinterpret_cast<const float2*>(&Ws[buf][k][tx * 2]);
      acc[0][0] = fmaf(a2.x, w2.x, acc[0][0]); acc[0][1] = fmaf(a2.x, w2.y, acc[0][1]);
      acc[1][0] = fmaf(a2.y, w2.x, acc[1][0]); acc[1][1] = fmaf(a2.y, w2.y, acc[1][1]);
    }
    if (kt + 1 < k_tiles) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }
  const int col = n0 + tx * 2;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = m0 + ty * 2 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (col + j >= n_store) continue;
      float x = acc[i][j];
      if (bias != nullptr && col + j < N) x += __ldg(bias + col + j);
      if (col + j >= N) x = 0.f;
      C[(int64_t)row * ldc + col + j] = Act<OutT>::from_float(relu ? fmaxf(x, 0.f) : x);
    }
  }
}

int gemm_f32(care_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
             int64_t ldc, int out_dtype, int M, int N, int K, int act, cudaStream_t stream) {
  CARE_CHECK_ARG(lda % 4 == 0 && ldw % 4 == 0 && K % 4 == 0,
                 "care_gemm(f32): lda, ldw, K must be multiples of 4 (got %lld, %lld, %d)", (long long)lda,
                 (long long)ldw, K);
  CARE_CHECK_ARG(ldc % 4 == 0, "care_gemm(f32): ldc must be a multiple of 4 (got %lld)", (long long)ldc);
  CARE_CHECK_ARG((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
                     (reinterpret_cast<uintptr_t>(C) & 15) == 0,
                 "care_gemm(f32): A, W, C must be 16-byte aligned");
  const int n_pad8 = (N + 7) & ~7;
  const int n_pad = n_pad8 <= ldc ? n_pad8 : ((N + 3) & ~3);
  CARE_CHECK_ARG(n_pad <= ldc, "care_gemm(f32): ldc %lld too small for N=%d", (long long)ldc, N);
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  if ((int)(grid.x * grid.y) * 4 <= ctx->sm_count) {   // too few large tiles to fill the machine
    dim3 sgrid((N + SB - 1) / SB, (M + SB - 1) / SB);
    if ((int)(sgrid.x * sgrid.y) < ctx->sm_count) {     // ... and too few 64x64 tiles as well
      dim3 tgrid((N + TB - 1) / TB, (M + TB - 1) / TB);
      if (out_dtype == CARE_F32)
        gemm_f32_tiny_kernel<float><<<tgrid, THREADS, 0, stream>>>((const float*)A, lda, (const float*)W, ldw, bias,
                                                                  (float*)C, ldc, M, N, n_pad, K, act == CARE_ACT_RELU,
                                                                  early_exit_of(ctx));
      else
        gemm_f32_tiny_kernel<h16><<<tgrid, THREADS, 0, stream>>>((const float*)A, lda, (const float*)W, ldw, bias,
                                                                (h16*)C, ldc, M, N, n_pad, K, act == CARE_ACT_RELU,
                                                                early_exit_of(ctx));
      ctx->last_gemm = "gemm_f32_tiny_kernel";
      CARE_LAUNCH_CHECK(ctx);
      return 0;
    }
    if (out_dtype == CARE_F32)
      gemm_f32_small_kernel<float><<<sgrid, THREADS, 0, stream>>>((const float*)A, lda, (const float*)W, ldw, bias,
                                                                 (float*)C, ldc, M, N, n_pad, K,
                                                                 act == CARE_ACT_RELU, early_exit_of(ctx));
    else
      gemm_f32_small_kernel<h16><<<sgrid, THREADS, 0, stream>>>((const float*)A, lda, (const float*)W, ldw, bias,
                                                               (h16*)C, ldc, M, N, n_pad, K, act == CARE_ACT_RELU,
                                                               early_exit_of(ctx));
    ctx->last_gemm = "gemm_f32_small_kernel";
    CARE_LAUNCH_CHECK(ctx);
    return 0;
  }
  ctx->last_gemm = "gemm_f32_kernel";
  if (out_dtype == CARE_F32)
    gemm_f32_kernel<float><<<grid, THREADS, 0, stream>>>((const float*)A, lda, (const float*)W, ldw, bias, (float*)C,
                                                         ldc, M, N, n_pad, K, act == CARE_ACT_RELU, early_exit_of(ctx));
  else
    gemm_f32_kernel<h16><<<grid, THREADS, 0, stream>>>((const float*)A, lda, (const float*)W, ldw, bias,
                                                                 (h16*)C, ldc, M, N, n_pad, K,
                                                                 act == CARE_ACT_RELU, early_exit_of(ctx));
  CARE_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace f32

namespace tc {
int gemm_bf16(care_ctx* ctx, const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C,
              int64_t ldc, int out_dtype, int M, int N, int K, int act, cudaStream_t stream);
}
}  // namespace care

extern "C" int care_gemm(care_ctx* ctx, int dtype, const void* A, int64_t lda, const void* W, int64_t ldw,
                         const float* bias, void* C, int64_t ldc, int out_dtype, int M, int N, int K, int act,
                         void* stream) {
  CARE_CHECK_ARG(ctx != nullptr, "care_gemm: ctx is NULL");
  CARE_CHECK_ARG(A && W && C, "care_gemm: NULL operand");
  CARE_CHECK_ARG(M > 0 && N > 0 && K > 0, "care_gemm: bad shape M=%d N=%d K=%d", M, N, K);
  CARE_CHECK_ARG(out_dtype == CARE_F32 || out_dtype == CARE_H16, "care_gemm: bad out_dtype %d", out_dtype);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (dtype == CARE_F32) return care::f32::gemm_f32(ctx, A, lda, W, ldw, bias, C, ldc, out_dtype, M, N, K, act, s);
  if (dtype == CARE_H16) return care::tc::gemm_bf16(ctx, A, lda, W, ldw, bias, C, ldc, out_dtype, M, N, K, act, s);
  care::set_error("care_gemm: bad dtype %d", dtype);
  return -1;
}
