// End-to-end entry with HOST buffers: token ids -> speech tokens -> mel -> waveform.
// Replaces the three-stage body of inference_zero_shot / inference_tts
// (server/model_utils/infer_speech_model.py:549-592, 631-670): llm.inference -> flow.inference ->
// F.interpolate for `speed` -> hift.inference -> .cpu().  Host<->device copies are inside the call.
#include "common.cuh"
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <memory>
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

// fade_in_out (cosyvoice/utils/common.py:169-177): the first `ov` samples of the new chunk are cross-faded with the previous
// chunk's tail under a Hamming window of 2*ov taps.  The reference multiplies float32 tensors by a float64 numpy window, i.e. the
// blend is evaluated in double and rounded once on the store — reproduced here.
__global__ void fade_in_out_kernel(float* __restrict__ wav, const float* __restrict__ prev_tail, const double* __restrict__ window, int ov) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ov) return;
  // three separately rounded double operations, as the reference's tensor expression evaluates them (no fused multiply-add)
  wav[i] = (float)__dadd_rn(__dmul_rn((double)wav[i], window[i]), __dmul_rn((double)prev_tail[i], window[ov + i]));
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

extern "C" hvx_status hvx_fade_in_out(hvx_engine* e, float* wav_dev, const float* prev_tail_dev, const double* window_dev, int overlap,
                                      void* stream) {
  HVX_CHECK(e && wav_dev && prev_tail_dev && window_dev && overlap >= 1, HVX_ERR_ARG, "fade_in_out: bad argument");
  fade_in_out_kernel<<<cdiv(overlap, 256), 256, 0, (cudaStream_t)stream>>>(wav_dev, prev_tail_dev, window_dev, overlap);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

extern "C" hvx_status hvx_synthesize_host(hvx_engine* e, const hvx_request* reqs, int n_req, int head_k, const hvx_sampler* sp,
                                          int n_timesteps, const float* noise_dev, const float* sine_table_dev, int64_t n_table_rows,
                                          float* wav_host, int wav_stride, int32_t* wav_len_host, int32_t* tokens_host, int tok_stride,
                                          int32_t* n_tokens_host, float* stage_ms_host, void* stream) {
  HVX_CHECK(e && e->llm && e->flow && e->hift, HVX_ERR_STATE, "synthesize: all three stages must be finalized");
  // the calling thread drives the decode (LLM lock); the worker thread below owns the flow and the vocoder for the call
  HVX_LOCK(e, HVX_STAGE_LLM);
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

  // ---- stages 2 + 3 of a group of finished utterances, enqueued on the caller's stream.  Stage 2 solves the group in one pass
  // per Euler step (hvx_flow_inference_batch; the reference's flow asserts batch 1, flow.py:387), stage 3 runs per utterance.
  // Everything is stream-ordered: the group workspace is reused by the next group only after this group's vocoder passes and
  // D2H copies, which were enqueued before it; the host waits once, at the end.
  std::vector<int32_t> cnt(n_req, 0);
  std::vector<char> finished(n_req, 0), enqueued(n_req, 0);
  size_t n_groups = 0;
  auto frames = [&](int i) { return 2 * (reqs[i].n_prompt_speech + cnt[i]); };
  auto enqueue_group = [&](const std::vector<int>& members) -> hvx_status {
    const int U = (int)members.size();
    for (int i : members) {
      const int T_sp = speed_frames(2 * cnt[i], reqs[i].speed);
      HVX_CHECK((size_t)T_sp * frame <= (size_t)wav_stride, HVX_ERR_ARG, "synthesize: wav_stride %d too small for %d samples", wav_stride, T_sp * frame);
      HVX_CHECK((int64_t)T_sp * frame <= n_table_rows, HVX_ERR_ARG, "synthesize: request %d needs %lld sine-table rows, the table holds %lld",
                i, (long long)T_sp * frame, (long long)n_table_rows);
    }
    size_t m = 0;
    auto take2 = [&](size_t bytes) { size_t o = m; m += (bytes + 255) & ~(size_t)255; return o; };
    std::vector<size_t> o_all(U), o_mel(U), o_mel2(U), o_wav(U);
    for (int k = 0; k < U; k++) {
      const int i = members[k];
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
      const int i = members[k];
      const hvx_request& r = reqs[i];
      int32_t* all_tok = (int32_t*)(dm + o_all[k]);
      if (r.n_prompt_speech)
        HVX_CUDA(cudaMemcpyAsync(all_tok, din + of[i].ps, sizeof(int32_t) * r.n_prompt_speech, cudaMemcpyDeviceToDevice, st));
      HVX_CUDA(cudaMemcpyAsync(all_tok + r.n_prompt_speech, tok_dev + (size_t)i * max_out, sizeof(int32_t) * cnt[i], cudaMemcpyDeviceToDevice, st));
      toks[k] = all_tok; embs[k] = (const float*)(din + of[i].emb);
      pfs[k] = r.n_prompt_speech ? (const float*)(din + of[i].pf) : nullptr;
      mels[k] = (float*)(dm + o_mel[k]); np[k] = r.n_prompt_speech; nt[k] = cnt[i];
    }
    hvx_status rc2;
    HVX_CUDA(cudaEventRecord(ev_a, st));
    if (U == 1) rc2 = hvx_flow_inference(e, toks[0], np[0], nt[0], embs[0], pfs[0], noise_dev, n_timesteps, 0, 1, mels[0], stream);
    else rc2 = hvx_flow_inference_batch(e, U, toks.data(), np.data(), nt.data(), embs.data(), pfs.data(), noise_dev, n_timesteps, 0, 1, mels.data(), stream);
    if (rc2) return rc2;
    HVX_CUDA(cudaEventRecord(ev_b, st));
    for (int k = 0; k < U; k++) {
      const int i = members[k];
      const hvx_request& r = reqs[i];
      const int T = 2 * cnt[i];
      const int T_sp = speed_frames(T, r.speed);
      float* mel_dev = mels[k];
      if (T_sp != T) {
        if ((rc2 = hvx_speed_interp(e, mel_dev, mel, T, T_sp, (float*)(dm + o_mel2[k]), stream))) return rc2;
        mel_dev = (float*)(dm + o_mel2[k]);
      }
      float* wav_dev = (float*)(dm + o_wav[k]);
      if ((rc2 = hvx_hift_vocode(e, mel_dev, T_sp, 1, sine_table_dev, n_table_rows, nullptr, nullptr, wav_dev, nullptr, stream))) return rc2;
      HVX_CUDA(cudaMemcpyAsync(wav_host + (size_t)i * wav_stride, wav_dev, sizeof(float) * (size_t)T_sp * frame, cudaMemcpyDeviceToHost, st));
      wav_len_host[i] = T_sp * frame;
    }
    HVX_CUDA(cudaEventRecord(ev_c, st));
    n_groups++;
    return HVX_OK;
  };
  static const int group_frames = getenv("HVX_FLOW_GROUP_FRAMES") ? atoi(getenv("HVX_FLOW_GROUP_FRAMES")) : 8192;   // 0: one by one
  // partition of the finished, not yet enqueued utterances: longest first; a group takes utterances while its padded size stays
  // under group_frames and the padding under 20 %
  auto partition = [&]() {
    std::vector<int> order;
    for (int i = 0; i < n_req; i++) if (finished[i] && !enqueued[i] && cnt[i] > 0) order.push_back(i);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return frames(x) > frames(y); });
    std::vector<std::vector<int>> groups;
    for (size_t g0 = 0; g0 < order.size();) {
      const int Tmax = frames(order[g0]);
      size_t g1 = g0 + 1;
      while (g1 < order.size() && g1 - g0 < 16 && (int)(g1 - g0 + 1) * Tmax <= group_frames && frames(order[g1]) * 5 >= Tmax * 4) g1++;
      groups.emplace_back(order.begin() + g0, order.begin() + g1);
      g0 = g1;
    }
    return groups;
  };
  // ---- stage 1: multi-head AR decode of all requests together on the decode stream.  HVX_PIPE_OVERLAP=1: whenever the
  // decode loop reports finished sequences that fill a group (>= HVX_PIPE_MIN_FRAMES padded frames or 16 utterances), that group
  // starts its flow + vocoder on the caller's stream — the latency-bound decode of the long utterances and the
  // throughput-bound flow of the short ones share the GPU.  (The reference is stage-serial per request,
  // infer_speech_model.py:549-592; results do not depend on the schedule.)
  // Measured on the batch-32 mixed-length workload (profiles/r2_pipe_overlap.txt): no gain — 21.75 s stage-serial vs 21.8 s
  // overlapped.  The decode step's ~250 small kernels (PDL keeps the next one resident while the current one runs) fragment the
  // SMs so that the flow's 198 KB one-CTA-per-SM persistent GEMMs barely advance while the decode is alive, with or without stream
  // priorities or reserved SMs.  Kept opt-in (HVX_PIPE_OVERLAP=1); results are bit-identical either way.
  static const bool overlap = getenv("HVX_PIPE_OVERLAP") && atoi(getenv("HVX_PIPE_OVERLAP")) != 0;
  static const int min_frames = getenv("HVX_PIPE_MIN_FRAMES") ? atoi(getenv("HVX_PIPE_MIN_FRAMES")) : 4096;
  // Groups are enqueued by a worker thread: one group is ~5000 kernel launches, far more than the driver's launch queue holds,
  // so the enqueueing thread blocks until the GPU has consumed most of them — the decode loop must not be that thread
  // (measured: with a single thread the decode advanced 16 steps per flow group and took as long as the whole flow).
  struct GroupQueue {
    std::mutex mu; std::condition_variable cv; std::deque<std::vector<int>> q; bool closed = false; hvx_status rc = HVX_OK;
  } gq;
  int dev_id = 0;
  HVX_CUDA(cudaGetDevice(&dev_id));
  std::thread worker([&]() {
    cudaSetDevice(dev_id);
    std::lock_guard<std::recursive_mutex> lf(e->mu[HVX_STAGE_FLOW]), lh(e->mu[HVX_STAGE_HIFT]);
    for (;;) {
      std::vector<int> g;
      { std::unique_lock<std::mutex> lk(gq.mu);
        gq.cv.wait(lk, [&] { return gq.closed || !gq.q.empty(); });
        if (gq.q.empty()) return;
        g = std::move(gq.q.front()); gq.q.pop_front(); }
      const hvx_status r = gq.rc ? gq.rc : enqueue_group(g);
      if (r) { std::lock_guard<std::mutex> lk(gq.mu); if (!gq.rc) gq.rc = r; }
    }
  });
  auto push_group = [&](const std::vector<int>& g) {
    for (int i : g) enqueued[i] = 1;                                  // claimed: the next partition() skips them
    { std::lock_guard<std::mutex> lk(gq.mu); gq.q.push_back(g); }
    gq.cv.notify_one();
  };
  struct Joiner { GroupQueue& q; std::thread& t; ~Joiner() { { std::lock_guard<std::mutex> lk(q.mu); q.closed = true; } q.cv.notify_all(); if (t.joinable()) t.join(); } } joiner{gq, worker};
  struct ProgCtx { decltype(push_group)* push; decltype(partition)* part; decltype(frames)* frames_of; std::vector<int32_t>* cnt;
                   std::vector<char>* fin; bool overlap; long long min_frames; GroupQueue* gq; };
  ProgCtx pctx{&push_group, &partition, &frames, &cnt, &finished, overlap, min_frames, &gq};
  auto progress = [](void* vctx, int n, const int* done, const int* n_out) -> hvx_status {
    ProgCtx* x = (ProgCtx*)vctx;
    for (int i = 0; i < n; i++)
      if (done[i] && !(*x->fin)[i]) { (*x->fin)[i] = 1; (*x->cnt)[i] = n_out[i]; }
    { std::lock_guard<std::mutex> lk(x->gq->mu); if (x->gq->rc) return x->gq->rc; }      // a group failed: stop decoding
    if (!x->overlap) return HVX_OK;
    // The schedule depends only on the decode's own progress (which sequences have stopped after how many steps), never on how far
    // the caller's stream has got, so the grouping — and with it every output bit — is reproducible from run to run.
    for (const std::vector<int>& g : (*x->part)()) {
      long long fr = 0;
      for (int i : g) fr += (*x->frames_of)(i);
      if (fr < x->min_frames && g.size() < 16) continue;            // wait for more finished utterances to fill this group
      (*x->push)(g);
    }
    return HVX_OK;
  };
  // while the decode runs beside the flow, the persistent GEMM grids leave a few SMs free: the decode step is a chain of ~250
  // small kernels, and a kernel that has to wait for a 148-CTA persistent grid to drain costs 50-100 us each (measured: 20 ms
  // per decode step instead of 3.7); on reserved SMs it starts at once.  Grids are fixed at enqueue time.
  static const int reserve_sms = getenv("HVX_PIPE_RESERVE_SMS") ? atoi(getenv("HVX_PIPE_RESERVE_SMS")) : 12;
  struct Reserve { hvx_engine* e; Reserve(hvx_engine* x, int n) : e(x) { e->sm_reserve = n; } ~Reserve() { e->sm_reserve = 0; } };
  std::unique_ptr<Reserve> reserve(overlap && n_req > 1 ? new Reserve(e, reserve_sms) : nullptr);
  const auto t_llm0 = std::chrono::steady_clock::now();
  if ((rc = llm_generate_progress(e, n_req, head_k, sp, (const float*)(din + o_u), u_stride, tok_dev, max_out, cnt_dev, stream,
                                  +progress, &pctx))) return rc;
  const float ms_llm = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_llm0).count();
  reserve.reset();                                  // the decode has drained: the remaining groups get every SM
  if (tokens_host)
    for (int i = 0; i < n_req; i++)
      HVX_CUDA(cudaMemcpyAsync(tokens_host + (size_t)i * tok_stride, tok_dev + (size_t)i * max_out, sizeof(int32_t) * of[i].max_out,
                               cudaMemcpyDeviceToHost, st));
  for (int i = 0; i < n_req; i++) {
    HVX_CHECK(finished[i], HVX_ERR_STATE, "synthesize: sequence %d was not reported finished", i);
    if (n_tokens_host) n_tokens_host[i] = cnt[i];
    if (cnt[i] <= 0) wav_len_host[i] = 0;
  }
  for (const std::vector<int>& g : partition()) push_group(g);
  { std::lock_guard<std::mutex> lk(gq.mu); gq.closed = true; }
  gq.cv.notify_all();
  worker.join();
  if (gq.rc) return gq.rc;
  HVX_CUDA(cudaStreamSynchronize(st));
  float ms_flow = 0.f, ms_hift = 0.f;
  for (size_t g = 0; g < n_groups; g++) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, P->gev[3 * g], P->gev[3 * g + 1]);
    cudaEventElapsedTime(&b, P->gev[3 * g + 1], P->gev[3 * g + 2]);
    ms_flow += a; ms_hift += b;
  }
  if (stage_ms_host) { stage_ms_host[0] = ms_llm; stage_ms_host[1] = ms_flow; stage_ms_host[2] = ms_hift; }
  return HVX_OK;
}
