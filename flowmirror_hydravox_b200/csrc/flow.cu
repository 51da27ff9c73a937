#include "common.cuh"
namespace hvx {
hvx_status flow_finalize(hvx_engine*) { set_error("flow not built"); return HVX_ERR_UNSUPPORTED; }
void flow_free(hvx_engine*) {}
}
using namespace hvx;
extern "C" hvx_status hvx_flow_inference(hvx_engine*, const int32_t*, int, int, const float*, const float*, const float*, int,
                                         int, int, float*, void*) { set_error("flow not built"); return HVX_ERR_UNSUPPORTED; }
extern "C" hvx_status hvx_dit_estimator(hvx_engine*, const float*, const float*, const float*, const float*, const float*, int,
                                        int, float*, void*) { set_error("flow not built"); return HVX_ERR_UNSUPPORTED; }
