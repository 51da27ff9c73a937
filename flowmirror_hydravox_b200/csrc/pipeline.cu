// End-to-end entry with HOST buffers: token ids -> speech tokens -> mel -> waveform.
// Replaces the three-stage body of inference_zero_shot / inference_tts
// (server/model_utils/infer_speech_model.py:549-592, 631-670): llm.inference -> flow.inference ->
// F.interpolate for `speed` -> hift.inference -> .cpu().  Host<->device copies are inside the call.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

namespace hvx {

// F.interpolate(mel, size=T_out, mode='linear') (align_corners=False), mel (C, T) channel-major
// (infer_speech_model.py:584-587, 662-665)
__global__ void speed_interp_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int T, int T_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * T_out) return;
  const int c = i / T_out, t = i - c * T_out;
  const float scale = (float)T / (float)T_out;
  float src = ((float)t + 0.5f) * scale - 0.5f;
  if (src < 0.f) src = 0.f;
  const int i0 = min((int)src, T - 1);
  const int i1 = min(i0 + 1, T - 1);
  const float l1 = src - (float)i0, l0 = 1.0f - l1;
  y[i] = l0 * x[(size_t)c * T + i0] + l1 * x[(size_t)c * T + i1];
}

struct PipeState {
  DevBuf in, mid;
  cudaEvent_t ev[2] = {nullptr, nullptr};      // start, llm end
  std::vector<cudaEvent_t> gev;                // per flow group: flow start, flow end (= vocoder start), vocoder end
  cudaEvent_t group_event(size_t i) {
    while (gev.size() <= i) { cudaEvent_t x = nullptr; if (cudaEventCreate(&x) != cudaSuccess) return nullptr; gev.push_back(x); }
    return gev[i];
  }
};

// int(tts_mel.shape[2] / speed) evaluated like the reference does, on doubles, and never below one frame
// (infer_speech_model.py:584-587: F.interpolate(size=max(1, int(T / speed))))
static inline int speed_frames(int T, double speed) {
  if (speed == 1.0 || !(speed > 0.0)) return T;
  const int t = (int)((double)T / speed);
  return t < 1 ? 1 : t;
}
static PipeState* g_pipe = nullptr;      // one engine per process (server/worker.py:25-44)

}  // namespace hvx

using namespace hvx;

extern "C" hvx_status hvx_speed_interp(hvx_engine* e, const float* mel_dev, int C, int T, int T_out, float* out_dev, void* stream) {
  HVX_CHECK(e && mel_dev && out_dev && T >= 1 && T_out >= 1, HVX_ERR_ARG, "speed_interp: bad argument");
  speed_interp_kernel<<<cdiv(C * T_out, 256), 256, 0, (cudaStream_t)stream>>>(mel_dev, out_dev, C, T, T_out);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

extern "C" hvx_status hvx_synthesize_host(hvx_engine* e, const hvx_request* reqs, int n_req, int head_k, const hvx_sampler* sp,
                                          int n_timesteps, const float* noise_dev, const float* sine_table_dev, int64_t n_table_rows,
                                          float* wav_host, int wav_stride, int32_t* wav_len_host, int32_t* tokens_host, int tok_stride,
                                          int32_t* n_tokens_host, float* stage_ms_host, void* stream) {
  HVX_CHECK(e && e->llm && e->flow && e->hift, HVX_ERR_STATE, "synthesize: all three stages must be finalized");
  HVX_LOCK(e, HVX_STAGE_LLM); HVX_LOCK(e, HVX_STAGE_FLOW); HVX_LOCK(e, HVX_STAGE_HIFT);
  HVX_CHECK(reqs && sp && noise_dev && sine_table_dev && wav_host && wav_len_host, HVX_ERR_ARG, "synthesize: null argument");
  HVX_CHECK(n_req >= 1 && n_req <= e->cfg.llm_max_seqs, HVX_ERR_ARG, "synthesize: n_req=%d exceeds max_seqs=%d", n_req, e->cfg.llm_max_seqs);
  const hvx_config& c = e->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  if (!g_pipe) {
    g_pipe = new PipeState();
    for (auto& ev : g_pipe->ev) HVX_CUDA(cudaEventCreate(&ev));
  }
  PipeState* P = g_pipe;
  const int mel = c.flow_mel;
  int frame = c.hift_hop;
  for (int i = 0; i < c.hift_n_ups; i++) frame *= c.hift_ups[i];

  // ---- stage inputs: one H2D copy per host array
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  struct Off { size_t text, ps, pf, emb, u, out; int max_out; };
  std::vector<Off> of(n_req);
  int max_out = 0, u_stride = 0;
  for (int i = 0; i < n_req; i++) {
    const hvx_request& r = reqs[i];
    HVX_CHECK(r.text_ids_host && r.embedding_host && r.u_host && r.n_u > 0, HVX_ERR_ARG, "synthesize: request %d incomplete", i);
    HVX_CHECK(r.n_prompt_speech == 0 || (r.prompt_speech_host && r.prompt_feat_host), HVX_ERR_ARG, "synthesize: request %d prompt incomplete", i);
    of[i].max_out = (int)((float)r.n_text_new * r.max_ratio);
    max_out = std::max(max_out, of[i].max_out);
    u_stride = std::max(u_stride, r.n_u);
  }
  HVX_CHECK(tokens_host == nullptr || tok_stride >= max_out, HVX_ERR_ARG, "synthesize: tok_stride %d < max tokens %d", tok_stride, max_out);
  max_out = std::max(max_out, 1);
  for (int i = 0; i < n_req; i++) {
    const hvx_request& r = reqs[i];
    of[i].text = take(sizeof(int32_t) * std::max(1, r.n_text_total));
    of[i].ps = take(sizeof(int32_t) * std::max(1, r.n_prompt_speech));
    of[i].pf = take(sizeof(float) * std::max(1, 2 * r.n_prompt_speech * mel));
    of[i].emb = take(sizeof(float) * c.flow_spk_in);
  }
  const size_t o_u = take(sizeof(float) * (size_t)n_req * u_stride);
  const size_t o_tok = take(sizeof(int32_t) * (size_t)n_req * max_out), o_cnt = take(sizeof(int32_t) * n_req);
  uint8_t* din = (uint8_t*)P->in.get(off);
  HVX_CHECK(din, HVX_ERR_CUDA, "synthesize: input staging allocation failed");
  HVX_CUDA(cudaEventRecord(P->ev[0], st));
  for (int i = 0; i < n_req; i++) {
    const hvx_request& r = reqs[i];
    HVX_CUDA(cudaMemcpyAsync(din + of[i].text, r.text_ids_host, sizeof(int32_t) * r.n_text_total, cudaMemcpyHostToDevice, st));
    if (r.n_prompt_speech) {
      HVX_CUDA(cudaMemcpyAsync(din + of[i].ps, r.prompt_speech_host, sizeof(int32_t) * r.n_prompt_speech, cudaMemcpyHostToDevice, st));
      HVX_CUDA(cudaMemcpyAsync(din + of[i].pf, r.prompt_feat_host, sizeof(float) * 2 * r.n_prompt_speech * mel, cudaMemcpyHostToDevice, st));
    }
    HVX_CUDA(cudaMemcpyAsync(din + of[i].emb, r.embedding_host, sizeof(float) * c.flow_spk_in, cudaMemcpyHostToDevice, st));
    HVX_CUDA(cudaMemcpyAsync(din + o_u + sizeof(float) * (size_t)i * u_stride, r.u_host, sizeof(float) * r.n_u, cudaMemcpyHostToDevice, st));
  }
  // ---- stage 1: multi-head AR decode of all requests together
  hvx_status rc;
  for (int i = 0; i < n_req; i++) {
    const hvx_request& r = reqs[i];
    if ((rc = hvx_llm_begin(e, i, (const int32_t*)(din + of[i].text), r.n_text_total, r.n_text_new,
                            r.n_prompt_speech ? (const int32_t*)(din + of[i].ps) : nullptr, r.n_prompt_speech, r.min_ratio, r.max_ratio)))
      return rc;
  }
  int32_t* tok_dev = (int32_t*)(din + o_tok);
  int32_t* cnt_dev = (int32_t*)(din + o_cnt);
  if ((rc = hvx_llm_generate(e, n_req, head_k, sp, (const float*)(din + o_u), u_stride, tok_dev, max_out, cnt_dev, stream))) return rc;
  std::vector<int32_t> cnt(n_req);
  HVX_CUDA(cudaMemcpyAsync(cnt.data(), cnt_dev, sizeof(int32_t) * n_req, cudaMemcpyDeviceToHost, st));
  if (tokens_host)
    for (int i = 0; i < n_req; i++)
      HVX_CUDA(cudaMemcpyAsync(tokens_host + (size_t)i * tok_stride, tok_dev + (size_t)i * max_out, sizeof(int32_t) * of[i].max_out,
                               cudaMemcpyDeviceToHost, st));
  HVX_CUDA(cudaEventRecord(P->ev[1], st));
  HVX_CUDA(cudaStreamSynchronize(st));            // token counts size the next two stages
  float ms_llm = 0.f, ms_flow = 0.f, ms_hift = 0.f;
  cudaEventElapsedTime(&ms_llm, P->ev[0], P->ev[1]);
  // ---- stage 2 in groups of similar length (one pass per Euler step for the whole group, hvx_flow_inference_batch; the
  // reference's flow asserts batch 1, flow.py:387), stage 3 per utterance.  Everything below is stream-ordered: the group
  // workspace is reused by the next group only after this group's vocoder passes and D2H copies, which were enqueued before
  // it; the host waits once, at the end.
  std::vector<int> order;
  for (int i = 0; i < n_req; i++) {
    if (n_tokens_host) n_tokens_host[i] = cnt[i];
    if (cnt[i] <= 0) wav_len_host[i] = 0; else order.push_back(i);
  }
  auto frames = [&](int i) { return 2 * (reqs[i].n_prompt_speech + cnt[i]); };
  std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return frames(x) > frames(y); });
  // every utterance is validated before any of stage 2 is enqueued
  for (int i : order) {
    const int T_sp = speed_frames(2 * cnt[i], reqs[i].speed);
    HVX_CHECK((size_t)T_sp * frame <= (size_t)wav_stride, HVX_ERR_ARG, "synthesize: wav_stride %d too small for %d samples", wav_stride, T_sp * frame);
    HVX_CHECK((int64_t)T_sp * frame <= n_table_rows, HVX_ERR_ARG, "synthesize: request %d needs %lld sine-table rows, the table holds %lld",
              i, (long long)T_sp * frame, (long long)n_table_rows);
  }
  static const int group_frames = getenv("HVX_FLOW_GROUP_FRAMES") ? atoi(getenv("HVX_FLOW_GROUP_FRAMES")) : 8192;   // 0: one by one
  size_t n_groups = 0;
  for (size_t g0 = 0; g0 < order.size(); n_groups++) {
    // longest first; add utterances while the padded group stays under group_frames and padding under 20 %
    const int Tmax = frames(order[g0]);
    size_t g1 = g0 + 1;
    while (g1 < order.size() && g1 - g0 < 16 && (int)(g1 - g0 + 1) * Tmax <= group_frames && frames(order[g1]) * 5 >= Tmax * 4) g1++;
    const int U = (int)(g1 - g0);
    size_t m = 0;
    auto take2 = [&](size_t bytes) { size_t o = m; m += (bytes + 255) & ~(size_t)255; return o; };
    std::vector<size_t> o_all(U), o_mel(U), o_mel2(U), o_wav(U);
    for (int k = 0; k < U; k++) {
      const int i = order[g0 + k];
      const hvx_request& r = reqs[i];
      const int T = 2 * cnt[i];
      const int T_sp = speed_frames(T, r.speed);
      o_all[k] = take2(sizeof(int32_t) * (r.n_prompt_speech + cnt[i]));
      o_mel[k] = take2(sizeof(float) * mel * T);
      o_mel2[k] = T_sp != T ? take2(sizeof(float) * mel * T_sp) : 0;
      o_wav[k] = take2(sizeof(float) * (size_t)T_sp * frame);
    }
    cudaEvent_t ev_a = P->group_event(3 * n_groups), ev_b = P->group_event(3 * n_groups + 1), ev_c = P->group_event(3 * n_groups + 2);
    HVX_CHECK(ev_a && ev_b && ev_c, HVX_ERR_CUDA, "synthesize: event creation failed");
    // the workspace may move when it grows: everything enqueued so far must have drained first (rare: the first call of a shape)
    if (m > P->mid.bytes) HVX_CUDA(cudaStreamSynchronize(st));
    uint8_t* dm = (uint8_t*)P->mid.get(m);
    HVX_CHECK(dm, HVX_ERR_CUDA, "synthesize: intermediate allocation failed");
    std::vector<const int32_t*> toks(U);
    std::vector<const float*> embs(U), pfs(U);
    std::vector<float*> mels(U);
    std::vector<int> np(U), nt(U);
    for (int k = 0; k < U; k++) {
      const int i = order[g0 + k];
      const hvx_request& r = reqs[i];
      int32_t* all_tok = (int32_t*)(dm + o_all[k]);
      if (r.n_prompt_speech)
        HVX_CUDA(cudaMemcpyAsync(all_tok, din + of[i].ps, sizeof(int32_t) * r.n_prompt_speech, cudaMemcpyDeviceToDevice, st));
      HVX_CUDA(cudaMemcpyAsync(all_tok + r.n_prompt_speech, tok_dev + (size_t)i * max_out, sizeof(int32_t) * cnt[i], cudaMemcpyDeviceToDevice, st));
      toks[k] = all_tok; embs[k] = (const float*)(din + of[i].emb);
      pfs[k] = r.n_prompt_speech ? (const float*)(din + of[i].pf) : nullptr;
      mels[k] = (float*)(dm + o_mel[k]); np[k] = r.n_prompt_speech; nt[k] = cnt[i];
    }
    HVX_CUDA(cudaEventRecord(ev_a, st));
    if (U == 1) rc = hvx_flow_inference(e, toks[0], np[0], nt[0], embs[0], pfs[0], noise_dev, n_timesteps, 0, 1, mels[0], stream);
    else rc = hvx_flow_inference_batch(e, U, toks.data(), np.data(), nt.data(), embs.data(), pfs.data(), noise_dev, n_timesteps, 0, 1, mels.data(), stream);
    if (rc) return rc;
    HVX_CUDA(cudaEventRecord(ev_b, st));
    for (int k = 0; k < U; k++) {
      const int i = order[g0 + k];
      const hvx_request& r = reqs[i];
      const int T = 2 * cnt[i];
      const int T_sp = speed_frames(T, r.speed);
      float* mel_dev = mels[k];
      if (T_sp != T) {
        if ((rc = hvx_speed_interp(e, mel_dev, mel, T, T_sp, (float*)(dm + o_mel2[k]), stream))) return rc;
        mel_dev = (float*)(dm + o_mel2[k]);
      }
      float* wav_dev = (float*)(dm + o_wav[k]);
      if ((rc = hvx_hift_vocode(e, mel_dev, T_sp, 1, sine_table_dev, n_table_rows, nullptr, nullptr, wav_dev, nullptr, stream))) return rc;
      HVX_CUDA(cudaMemcpyAsync(wav_host + (size_t)i * wav_stride, wav_dev, sizeof(float) * (size_t)T_sp * frame, cudaMemcpyDeviceToHost, st));
      wav_len_host[i] = T_sp * frame;
    }
    HVX_CUDA(cudaEventRecord(ev_c, st));
    g0 = g1;
  }
  HVX_CUDA(cudaStreamSynchronize(st));
  for (size_t g = 0; g < n_groups; g++) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, P->gev[3 * g], P->gev[3 * g + 1]);
    cudaEventElapsedTime(&b, P->gev[3 * g + 1], P->gev[3 * g + 2]);
    ms_flow += a; ms_hift += b;
  }
  if (stage_ms_host) { stage_ms_host[0] = ms_llm; stage_ms_host[1] = ms_flow; stage_ms_host[2] = ms_hift; }
  return HVX_OK;
}
