// Shared plumbing for the care_b200 kernels: error convention, ctx, small device helpers.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <utility>

#include "../../include/care_b200.h"

// The 16-bit operand type of this build of the library.  Default: IEEE fp16 (10 mantissa bits - activations
// of this model are O(1) after every LayerNorm, so fp16's range is ample and its rounding error is 8x below
// bf16's; bf16-stored weights are exactly representable).  -DCARE_USE_BF16 builds the same kernels over bf16
// (libcare_b200_bf16.so).  Tensor-core throughput is identical for both (tcgen05 kind::f16, mma.sync m16n8k16).
#ifdef CARE_USE_BF16
typedef __nv_bfloat16 h16;
typedef __nv_bfloat162 h162;
#define CARE_H16 CARE_BF16
#define CARE_H16_NAME "bf16"
#define CARE_TMAP_H16 CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define CARE_MMA_H16 "bf16"
#define CARE_UMMA_FMT 1u   // cute::UMMA::F16F32Format: F16 = 0, BF16 = 1
__device__ __forceinline__ h162 floats_to_h162(float a, float b) { return __floats2bfloat162_rn(a, b); }
__device__ __forceinline__ h16 float_to_h16(float a) { return __float2bfloat16_rn(a); }
__device__ __forceinline__ float2 h162_to_float2(h162 a) { return __bfloat1622float2(a); }
__device__ __forceinline__ float h16_to_float(h16 a) { return __bfloat162float(a); }
#else
typedef __half h16;
typedef __half2 h162;
#define CARE_H16 CARE_F16
#define CARE_H16_NAME "fp16"
#define CARE_TMAP_H16 CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define CARE_MMA_H16 "f16"
#define CARE_UMMA_FMT 0u
// saturating: a value beyond fp16's range becomes +-65504 instead of inf (inf - inf = NaN inside a softmax)
__device__ __forceinline__ float sat16(float a) { return fminf(fmaxf(a, -65504.f), 65504.f); }
__device__ __forceinline__ h162 floats_to_h162(float a, float b) { return __floats2half2_rn(sat16(a), sat16(b)); }
__device__ __forceinline__ h16 float_to_h16(float a) { return __float2half_rn(sat16(a)); }
__device__ __forceinline__ float2 h162_to_float2(h162 a) { return __half22float2(a); }
__device__ __forceinline__ float h16_to_float(h16 a) { return __half2float(a); }
#endif

namespace care {

void set_error(const char* fmt, ...);

#define CARE_CHECK_ARG(cond, ...)          \
  do {                                     \
    if (!(cond)) {                         \
      ::care::set_error(__VA_ARGS__);      \
      return -1;                           \
    }                                      \
  } while (0)

#define CARE_CHECK_DTYPE(dtype, who)                                                                 \
  CARE_CHECK_ARG((dtype) == CARE_F32 || (dtype) == CARE_H16,                                         \
                 "%s: dtype code %d - this build of the library computes in fp32 (0) or " CARE_H16_NAME " (%d)", who, \
                 (int)(dtype), (int)CARE_H16)

#define CARE_CUDA(expr)                                                                   \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::care::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                        __LINE__);                                                        \
      return (int)_e;                                                                     \
    }                                                                                     \
  } while (0)

#define CARE_LAUNCH_CHECK(ctx)            \
  do {                                    \
    (ctx)->launches++;                    \
    CARE_CUDA(cudaGetLastError());        \
  } while (0)

// Cache key of an encoded TMA descriptor (any rank <= 3, bf16 elements, SWIZZLE_128B).
struct TmapKey {
  const void* ptr;
  uint32_t rank;
  uint64_t gdim[3];
  uint64_t gstride[2];  // bytes, dims 1..rank-1
  uint32_t box[3];
  bool operator==(const TmapKey& o) const {
    if (ptr != o.ptr || rank != o.rank) return false;
    for (int i = 0; i < 3; ++i)
      if (gdim[i] != o.gdim[i] || box[i] != o.box[i]) return false;
    return gstride[0] == o.gstride[0] && gstride[1] == o.gstride[1];
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = (size_t)k.ptr;
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); };
    mix(k.rank);
    for (int i = 0; i < 3; ++i) { mix(k.gdim[i]); mix(k.box[i]); }
    mix(k.gstride[0]); mix(k.gstride[1]);
    return h;
  }
};

}  // namespace care

typedef CUresult (*care_tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct care_ctx {
  int device = 0;
  int sm_count = 0;
  int64_t launches = 0;
  int attn_impl = 1;   // 1: TMA + mma.sync attention for bf16 (default), 0: SIMT kernel everywhere
  // bf16 self-attention: 2 (default) = gathered chunk stream over the cache slots still referenced by a beam,
  // 1 = per-CTA gather of those slots (slower than the dense tile, kept for A/B runs), 0 = dense TMA tiles
  int self_compact = 2;
  int gemm_smallm = 1;   // M <= 16: weight-streaming mma.sync kernel instead of 128-row tensor-core tiles
  uint32_t* compact_info = nullptr;   // [compact_info_videos][160] per-video records of the compacting kernel
  int compact_info_videos = 0;
  unsigned long long* self_attn_rows = nullptr;   // device counter: K/V cache rows read per video (counted by head 0)
  // 0: single-CTA tiles only, 1: CTA-pair (cta_group::2) tiles whenever the shape allows, 2 (default): pick per
  // (M, N, K, out dtype) by timing the variants once on the first call with that shape (skipped while capturing);
  // 4 / 5: clusters of 4 / 2 CTA pairs that share the A tile by TMA multicast (gemm_bf16_2sm_mc_kernel) whenever possible
  int gemm_2sm = 2;
  const char* last_gemm = "";    // variant names of the most recent launches (care_ctx_last_kernel)
  const char* last_vocab = "";
  const char* last_self_attn = "";
  int gemm_bn = 0;   // > 0: force this tile width in the single-CTA GEMM (A/B runs)
  int debug = 0;
  // programmatic dependent launch for the kernels of a decode step: kernel N+1 is scheduled while kernel N drains,
  // runs its prologue (barrier init, TMEM allocation, descriptor prefetch) and blocks in griddepcontrol.wait
  // until N has completed and flushed (option "pdl", env CARE_B200_PDL)
  int pdl = 1;
  // one-shot request armed by care_ctx_set_next_step: the next beam step also computes the following step's input rows
  care_next_step next = {};
  bool next_armed = false;
  // the live-slot record table already holds the records for (n_pos, B, anc): written by the last beam step
  // the beam kernel also writes the next step's live-slot records (option "fuse_info", env CARE_B200_FUSE_INFO).  Off
  // by default: measured neutral to slightly slower (4096 videos 51.2 -> 51.9 ms) - the one-warp-per-video beam kernel
  // is the latency-bound one, and with graphs + programmatic dependent launch a separate small launch costs little
  int fuse_info = 0;
  // one-shot request armed by care_ctx_request_records: the next care_embed_ln launch also writes the live-slot records of
  // the self-attention that follows it (extra CTAs of the same launch instead of a record kernel of its own)
  struct RecordsReq {
    const uint8_t* anc = nullptr;
    int anc_stride = 0;
    const int32_t* tok_hist = nullptr;
    const int32_t* done = nullptr;
    int B = 0, K = 0, H = 0, n_pos = 0;
  } rec_req;
  bool rec_req_armed = false;
  int info_ready_npos = -1;
  int info_ready_B = 0;
  const void* info_ready_anc = nullptr;
  // care_gemm_add_ln: A tile fetched once per cluster and multicast (option "gemm_ln_multicast", env CARE_B200_GEMM_LN_MC).
  // Off by default: measured neutral (49 vs 49 us at 20480 x 1024 x 1024) - the 128 x 256 single-CTA mainloop is bound by
  // shared-memory bandwidth (48 KB of operands read by the MMAs plus 48 KB written by TMA per 512 MMA cycles = 192 B/clk
  // against 128 B/clk), not by L2; only CTA pairs (cta_group::2, half the B operand per SM) lift that
  int gemm_ln_multicast = 0;
  // care_gemm_add_ln on CTA pairs (gemm_add_ln_pair_kernel: clusters of 2 * d/256 CTAs over 256-row blocks): 0 = single-CTA
  // clusters only, 1 = pairs whenever such a cluster fits the device, 2 (default) = pick per (M, N, K) by timing both once
  // (option "gemm_ln_pair", env CARE_B200_GEMM_LN_PAIR)
  int gemm_ln_pair = 2;
  // L2 eviction priorities of the TMA loads (option "l2_hints", env CARE_B200_L2_HINTS): bit 0 = weight tiles evict_last,
  // bit 1 = cross-attention K/V tiles evict_first (default: measured 8.62 -> 8.39 ms per 512-video shard, 54.6 -> 53.9 ms per 4096
  // videos; bit 0 measured neutral)
  int l2_hints = 2;
  // fused vocabulary kernel, epilogue schedule (option "vocab_split"): 0 = the two epilogue groups take alternate tiles,
  // 1 = both fold every tile (column halves), 2 (default) = column halves when a run has fewer than vocab_split_tiles
  // tiles (option "vocab_split_tiles")
  int vocab_split = 2;
  int vocab_split_tiles = 24;
  int vocab_2sm = 1;   // fused vocabulary kernel on CTA pairs when the shape has >= two waves of pair tiles
  // per-shape GEMM variant picks; contexts that must launch identical kernels (the lanes of one decode) share one
  // table (care_ctx_share_tuning)
  struct GemmTuning {
    std::mutex mu;
    std::unordered_map<uint64_t, int> choice;
  };
  std::shared_ptr<GemmTuning> tuning = std::make_shared<GemmTuning>();
  // device-side early exit: kernels without a per-video `done` predicate return at once when
  // *skip_counter >= skip_target (all videos of the batch have finished); NULL disables
  const int32_t* skip_counter = nullptr;
  int skip_target = 0;
  care_tmap_encode_fn encode = nullptr;
  std::mutex mu;
  std::unordered_map<care::TmapKey, CUtensorMap, care::TmapKeyHash> tmaps;
};

namespace care {

// Encodes (or fetches from the ctx cache) a bf16 SWIZZLE_128B tiled TMA descriptor.  gdim/box are
// fastest-dimension first; gstride_bytes has rank-1 entries (dimension 0 is dense).  api.cu.
struct EarlyExit {
  const int32_t* counter;
  int target;
  int l2_hints;   // ctx->l2_hints: bit 0 = weight tiles are loaded evict_last, bit 1 = cross-attention K/V tiles evict_first
};
inline EarlyExit early_exit_of(const care_ctx* ctx) { return EarlyExit{ctx->skip_counter, ctx->skip_target, ctx->l2_hints}; }

int get_tmap_bf16(care_ctx* ctx, const void* ptr, int rank, const uint64_t* gdim, const uint64_t* gstride_bytes,
                  const uint32_t* box, CUtensorMap* out);

// Launch through cudaLaunchKernelEx, with the programmatic-stream-serialization attribute when ctx->pdl is set.
// Every kernel launched through this helper executes pdl_wait() before it touches global memory that an earlier
// kernel of the stream produces or still reads, and pdl_launch_dependents() only AFTER its own pdl_wait(): a
// kernel that has started therefore implies that everything before its immediate predecessor has completed.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(const care_ctx* ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool all_done(const EarlyExit& e) {
  return e.counter != nullptr && *reinterpret_cast<const volatile int32_t*>(e.counter) >= e.target;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T>
struct Act;
template <>
struct Act<float> {
  static constexpr int kDtype = CARE_F32;
  // 8 consecutive elements
  static __device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
  static __device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
  static __device__ __forceinline__ float to_float(float x) { return x; }
  static __device__ __forceinline__ float from_float(float x) { return x; }
};
template <>
struct Act<h16> {
  static constexpr int kDtype = CARE_H16;
  static __device__ __forceinline__ void load8(const h16* p, float (&v)[8]) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const h162* h = reinterpret_cast<const h162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = h162_to_float2(h[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  }
  static __device__ __forceinline__ void store8(h16* p, const float (&v)[8]) {
    uint4 raw;
    h162* h = reinterpret_cast<h162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = floats_to_h162(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
  }
  static __device__ __forceinline__ void load4(const h16* p, float (&v)[4]) {
    uint2 raw = *reinterpret_cast<const uint2*>(p);
    const h162* h = reinterpret_cast<const h162*>(&raw);
    float2 a = h162_to_float2(h[0]), b = h162_to_float2(h[1]);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void store4(h16* p, const float (&v)[4]) {
    uint2 raw;
    h162* h = reinterpret_cast<h162*>(&raw);
    h[0] = floats_to_h162(v[0], v[1]);
    h[1] = floats_to_h162(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = raw;
  }
  static __device__ __forceinline__ float to_float(h16 x) { return h16_to_float(x); }
  static __device__ __forceinline__ h16 from_float(float x) { return float_to_h16(x); }
};

}  // namespace care
