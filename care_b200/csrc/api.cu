// ctx lifetime, error string, TMA descriptor cache.
#include <cstdarg>
#include <cstring>

#include "common.cuh"

namespace care {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace care

extern "C" {

int care_version(void) { return 100; }

const char* care_last_error(void) { return care::g_err; }

int care_ctx_create(care_ctx** out, int device) {
  CARE_CHECK_ARG(out != nullptr, "care_ctx_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    care::set_error("care_ctx_create: no CUDA device (%s); care_b200 has no CPU path",
                    e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    return e == cudaSuccess ? -2 : (int)e;
  }
  CARE_CHECK_ARG(device >= 0 && device < n, "care_ctx_create: bad device %d", device);
  cudaDeviceProp prop;
  CARE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    care::set_error("care_ctx_create: device %d is sm_%d%d; this library is sm_100a only", device,
                    prop.major, prop.minor);
    return -3;
  }
  CARE_CUDA(cudaSetDevice(device));
  care_ctx* c = new care_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    care::set_error("care_ctx_create: cuTensorMapEncodeTiled not available from the driver");
    delete c;
    return -4;
  }
  c->encode = (care_tmap_encode_fn)fn;
  *out = c;
  return 0;
}

void care_ctx_destroy(care_ctx* ctx) { delete ctx; }

int care_ctx_sm_count(const care_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int64_t care_ctx_launch_count(const care_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
