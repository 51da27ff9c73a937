// Shared host-side plumbing of libhydravox_b200 (engine object, tensor registry, error state).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdio>
#include <cstdarg>
#include <atomic>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include "../../include/hydravox_b200.h"

namespace hvx {

void set_error(const char* fmt, ...);

struct Tensor {
  const void* p = nullptr;
  int dtype = HVX_F32;
  int ndim = 0;
  int64_t shape[4] = {0, 0, 0, 0};
  int64_t numel() const { int64_t n = 1; for (int i = 0; i < ndim; i++) n *= shape[i]; return n; }
  const float* f32() const { return (const float*)p; }
  const __nv_bfloat16* bf16() const { return (const __nv_bfloat16*)p; }
};

struct DevBuf {              // engine-owned scratch that grows on demand (never shrinks)
  void* p = nullptr;
  size_t bytes = 0;
  void* get(size_t need) {
    if (need > bytes) {
      if (p) cudaFree(p);
      size_t cap = need + need / 4;
      if (cudaMalloc(&p, cap) != cudaSuccess) { p = nullptr; bytes = 0; return nullptr; }
      bytes = cap;
    }
    return p;
  }
  ~DevBuf() { if (p) cudaFree(p); }
};

// Live per-class kernel timing for bench.py's roofline (include/hydravox_b200.h: hvx_profile_enable / hvx_profile_collect):
// while enabled, every launch of a class is bracketed by two cudaEventRecord on its own stream; collect() synchronises the
// device, sums event-to-event durations and work per class, and recycles the events.  Off (the default) costs one branch.
enum { PROF_GEMM = 0, PROF_ATTN = 1, PROF_HIFT_CONV = 2, PROF_LLM_STEP = 3, PROF_LAYERNORM = 4, PROF_LLM_PREFILL = 5, PROF_NCLS = 8 };
struct ProfRec { int cls; cudaEvent_t a, b; double work; };
struct Prof {
  bool on = false;
  std::mutex mu;
  std::vector<cudaEvent_t> free_events;
  std::vector<ProfRec> recs;
  double ms[PROF_NCLS] = {0}, work[PROF_NCLS] = {0};
  long long n[PROF_NCLS] = {0};
  cudaEvent_t get() {
    if (!free_events.empty()) { cudaEvent_t x = free_events.back(); free_events.pop_back(); return x; }
    cudaEvent_t x = nullptr;
    cudaEventCreate(&x);
    return x;
  }
};
// RAII bracket around the launches of one call site; inert when profiling is off or the stream is being captured into a graph
struct ProfScope {
  Prof* p = nullptr; cudaStream_t st; ProfRec r;
  ProfScope(Prof* prof_p, cudaStream_t stream, int cls, double work) : st(stream) {
    if (!prof_p) return;
    Prof& prof = *prof_p;
    if (!prof.on) return;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;
    p = &prof;
    std::lock_guard<std::mutex> g(prof.mu);
    r.cls = cls; r.work = work; r.a = prof.get(); r.b = prof.get();
    cudaEventRecord(r.a, st);
  }
  ~ProfScope() {
    if (!p) return;
    cudaEventRecord(r.b, st);
    std::lock_guard<std::mutex> g(p->mu);
    p->recs.push_back(r);
  }
};

struct HiftState;
struct FlowState;
struct LlmState;
struct UnetState;

}  // namespace hvx

struct hvx_engine {
  hvx_config cfg;
  std::unordered_map<std::string, hvx::Tensor> tensors[4];
  hvx::HiftState* hift = nullptr;
  hvx::FlowState* flow = nullptr;
  hvx::LlmState* llm = nullptr;
  hvx::UnetState* unet = nullptr;
  std::atomic<int64_t> launches{0};
  // One lock per stage (LLM, flow, HiFT, U-Net): calls into DIFFERENT stages may run concurrently from different host threads
  // (streaming: AR decode on one thread, chunked flow + vocoder on another — they touch disjoint engine state and streams);
  // calls into the same stage serialise here.  Recursive: hvx_synthesize_host holds all of them and calls the stage entries.
  std::recursive_mutex mu[4];
  std::atomic<int> llm_cancel{0};
  hvx::Prof prof;
  int prof_gemm_off = 0;           // > 0: gemm_bf16 calls are not counted as PROF_GEMM (they belong to an enclosing LLM scope)  // hvx_llm_cancel: a running hvx_llm_generate stops after its current batch of steps
  hvx::DevBuf samp_ws;             // sampler tables (llm.cu)
  hvx::DevBuf fe_ws;               // frontend spectrum scratch (frontend.cu)
  void* samp_arrive = nullptr;
  int64_t graph_launches = 0;      // kernels inside the captured decode-step graph
  int sm_count = 148;
  int sm_reserve = 0;              // SMs the persistent GEMM grids leave free (hvx_synthesize_host: room for the decode stream's kernels)
  const hvx::Tensor* find(int stage, const std::string& name) const {
    auto it = tensors[stage].find(name);
    return it == tensors[stage].end() ? nullptr : &it->second;
  }
};

#define HVX_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      hvx::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return HVX_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

#define HVX_CHECK(cond, code, ...)                 \
  do {                                             \
    if (!(cond)) { hvx::set_error(__VA_ARGS__); return (code); } \
  } while (0)

#define HVX_LOCK(e, stage) std::lock_guard<std::recursive_mutex> _hvx_lock_##stage((e)->mu[stage])

#define HVX_LAUNCH_CHECK(e)                                                                    \
  do {                                                                                         \
    (e)->launches++;                                                                           \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      hvx::set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return HVX_ERR_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

namespace hvx {
// stage entry points implemented in hift.cu / flow.cu / llm.cu
hvx_status hift_finalize(hvx_engine* e);
void hift_free(hvx_engine* e);
hvx_status flow_finalize(hvx_engine* e);
void flow_free(hvx_engine* e);
hvx_status llm_finalize(hvx_engine* e);
void llm_free(hvx_engine* e);
hvx_status unet_finalize(hvx_engine* e);
void unet_free(hvx_engine* e);
// hvx_llm_generate with a progress hook (llm.cu): called between batches of decode steps with every sequence's done flag and
// emitted-token count; a non-zero return aborts the generation with that status
typedef hvx_status (*llm_progress_fn)(void* ctx, int n_seq, const int* done, const int* n_out);
hvx_status llm_generate_progress(hvx_engine* e, int n_seq, int head_k, const hvx_sampler* sp, const float* u_dev, int u_stride,
                                 int32_t* out_tokens, int max_out, int32_t* out_counts, void* stream, llm_progress_fn progress,
                                 void* progress_ctx);

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Launch with programmatic dependent launch allowed: the kernel may become resident while its predecessor in the stream drains and
// run its prologue (barrier init, TMEM allocation, descriptor prefetch); it executes `griddepcontrol.wait` (hvx::pdl_wait) before
// it touches any global memory, so ordering with the predecessor's results is unchanged.  Only kernels that contain the wait are
// launched this way.  cluster > 1: thread-block cluster of that many CTAs along x.
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster >= 0) {                     // cluster < 0: plain launch (no programmatic serialization), cluster size -cluster
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    n++;
  } else cluster = -cluster;
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    n++;
  }
  cfg.attrs = at; cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}
#ifdef __CUDACC__
// let the next kernel of the stream start its prologue; then wait until the previous kernel has completed and its writes are visible
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
}  // namespace hvx
