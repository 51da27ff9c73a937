#include "common.cuh"
namespace hvx {
hvx_status llm_finalize(hvx_engine*) { set_error("llm not built"); return HVX_ERR_UNSUPPORTED; }
void llm_free(hvx_engine*) {}
}
using namespace hvx;
extern "C" hvx_status hvx_llm_begin(hvx_engine*, int, const int32_t*, int, int, const int32_t*, int, float, float) { set_error("llm not built"); return HVX_ERR_UNSUPPORTED; }
extern "C" hvx_status hvx_llm_generate(hvx_engine*, int, int, const hvx_sampler*, const float*, int, int32_t*, int, int32_t*, void*) { set_error("llm not built"); return HVX_ERR_UNSUPPORTED; }
extern "C" hvx_status hvx_llm_probe(hvx_engine*, int, float*, float*, void*) { set_error("llm not built"); return HVX_ERR_UNSUPPORTED; }
extern "C" hvx_status hvx_sample(hvx_engine*, const float*, int, const int32_t*, int, int, const hvx_sampler*, const float*, int, int32_t*, int32_t*, void*) { set_error("llm not built"); return HVX_ERR_UNSUPPORTED; }
extern "C" hvx_status hvx_synthesize_host(hvx_engine*, const hvx_request*, int, int, const hvx_sampler*, int, const float*, const float*, float*, int, int32_t*, int32_t*, int, int32_t*, float*, void*) { set_error("not built"); return HVX_ERR_UNSUPPORTED; }
