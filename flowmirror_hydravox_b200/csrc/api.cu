// C-ABI plumbing: engine lifecycle, tensor registry, error state (include/hydravox_b200.h).
#include "common.cuh"
#include <cstring>

namespace hvx {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace hvx

using namespace hvx;

extern "C" const char* hvx_last_error(void) { return g_err; }
extern "C" int hvx_version(void) { return 100; }

extern "C" hvx_status hvx_create(hvx_engine** out, const hvx_config* cfg) {
  HVX_CHECK(out && cfg, HVX_ERR_ARG, "hvx_create: null argument");
  int dev = 0, n = 0;
  cudaError_t ce = cudaGetDeviceCount(&n);
  HVX_CHECK(ce == cudaSuccess && n > 0, HVX_ERR_CUDA, "hvx_create: no CUDA device (%s) — this engine has no CPU fallback",
            cudaGetErrorString(ce));
  HVX_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  HVX_CUDA(cudaGetDeviceProperties(&prop, dev));
  HVX_CHECK(prop.major == 10, HVX_ERR_UNSUPPORTED, "hvx_create: device sm_%d%d is not Blackwell sm_100 (kernels are sm_100a only)",
            prop.major, prop.minor);
  hvx_engine* e = new hvx_engine();
  e->cfg = *cfg;
  e->sm_count = prop.multiProcessorCount;
  *out = e;
  return HVX_OK;
}

extern "C" hvx_status hvx_destroy(hvx_engine* e) {
  if (!e) return HVX_OK;
  hift_free(e);
  flow_free(e);
  llm_free(e);
  unet_free(e);
  delete e;
  return HVX_OK;
}

extern "C" hvx_status hvx_set_tensor(hvx_engine* e, int stage, const char* name, const void* dptr, int dtype,
                                     const int64_t* shape, int ndim) {
  HVX_CHECK(e && name && dptr && shape, HVX_ERR_ARG, "hvx_set_tensor: null argument");
  HVX_CHECK(stage >= 0 && stage < 4 && ndim >= 1 && ndim <= 4, HVX_ERR_ARG, "hvx_set_tensor(%s): bad stage/ndim", name);
  Tensor t;
  t.p = dptr; t.dtype = dtype; t.ndim = ndim;
  for (int i = 0; i < ndim; i++) t.shape[i] = shape[i];
  e->tensors[stage][name] = t;
  return HVX_OK;
}

extern "C" hvx_status hvx_finalize(hvx_engine* e, int stage) {
  HVX_CHECK(e, HVX_ERR_ARG, "hvx_finalize: null engine");
  switch (stage) {
    case HVX_STAGE_HIFT: return hift_finalize(e);
    case HVX_STAGE_FLOW: return flow_finalize(e);
    case HVX_STAGE_LLM: return llm_finalize(e);
    case HVX_STAGE_UNET: return unet_finalize(e);
  }
  set_error("hvx_finalize: bad stage %d", stage);
  return HVX_ERR_ARG;
}

extern "C" int64_t hvx_kernel_launches(hvx_engine* e) { return e ? e->launches.load() : 0; }

extern "C" hvx_status hvx_profile_enable(hvx_engine* e, int on) {
  HVX_CHECK(e, HVX_ERR_ARG, "profile: null engine");
  std::lock_guard<std::mutex> g(e->prof.mu);
  e->prof.on = on != 0;
  return HVX_OK;
}

extern "C" hvx_status hvx_profile_collect(hvx_engine* e, double* ms_out, double* work_out, int64_t* launches_out) {
  HVX_CHECK(e && ms_out && work_out && launches_out, HVX_ERR_ARG, "profile: null argument");
  HVX_CUDA(cudaDeviceSynchronize());
  Prof& p = e->prof;
  std::lock_guard<std::mutex> g(p.mu);
  for (const ProfRec& r : p.recs) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess && r.cls >= 0 && r.cls < PROF_NCLS) {
      p.ms[r.cls] += ms; p.work[r.cls] += r.work; p.n[r.cls]++;
    }
    p.free_events.push_back(r.a); p.free_events.push_back(r.b);
  }
  p.recs.clear();
  for (int i = 0; i < PROF_NCLS; i++) {
    ms_out[i] = p.ms[i]; work_out[i] = p.work[i]; launches_out[i] = p.n[i];
    p.ms[i] = 0; p.work[i] = 0; p.n[i] = 0;
  }
  return HVX_OK;
}
