/*
 * hydravox_b200 — C-ABI of the B200-native HydraVox hot path (AR decode -> CFM -> HiFT).
 *
 * Drop-in boundary.  Every entry point replaces one call the reference makes
 * into PyTorch on its hot path; the reference-side binding is the ctypes stub shown in
 * INTEGRATION.md (the reference is Python, so its "FFI" is ctypes / data_ptr()).
 *
 * Conventions (they mirror the reference's own raw-pointer seam for the TensorRT estimator,
 * cosyvoice/flow/flow_matching.py:126-153 and cosyvoice/utils/common.py:198-213):
 *   - plain C linkage, no exceptions; every call returns hvx_status (0 = OK) and
 *     hvx_last_error() returns a description of the last failure on the calling thread;
 *   - `*_dev` pointers are device pointers owned by the caller (torch tensors via data_ptr()),
 *     contiguous, valid for the duration of the call; `*_host` pointers are host memory
 *     (pinned for best throughput) — the *_host entry points include the H2D/D2H copies;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are
 *     stream-ordered and do not synchronise unless they return host data;
 *   - one engine per process/GPU (the reference runs one worker per GPU, server/worker.py:25-44).  Threading contract:
 *     every stage (LLM, flow, HiFT, U-Net) has its own lock inside the engine; entry points of DIFFERENT stages may be
 *     called concurrently from different host threads on different streams (the streaming path does: AR decode on one
 *     thread, chunked flow + vocoder on another, cosyvoice/cli/model.py:315-360); calls into the same stage serialise.
 *     hvx_synthesize_host takes all three locks for its duration.  hvx_set_tensor / hvx_finalize must not run concurrently
 *     with any compute call.
 */
#ifndef HYDRAVOX_B200_H
#define HYDRAVOX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int hvx_status;
enum { HVX_OK = 0, HVX_ERR_ARG = 1, HVX_ERR_CUDA = 2, HVX_ERR_STATE = 3, HVX_ERR_UNSUPPORTED = 4 };

enum { HVX_STAGE_LLM = 0, HVX_STAGE_FLOW = 1, HVX_STAGE_HIFT = 2, HVX_STAGE_UNET = 3 };
enum { HVX_F32 = 0, HVX_BF16 = 1, HVX_I32 = 2, HVX_F16 = 3 };

typedef struct hvx_engine hvx_engine;

/* Model dimensions (flowmirror_hydravox_b200/dims.py; SURVEY.md §8 header). */
typedef struct hvx_config {
  /* HiFT: cosyvoice/hifigan/generator.py:577-670 */
  int hift_mel, hift_base, hift_f0_ch, hift_harmonics, hift_sr;
  int hift_n_ups, hift_ups[4], hift_up_k[4];
  int hift_n_fft, hift_hop;
  int hift_n_rb, hift_rb_k[4], hift_n_dil, hift_rb_d[4], hift_src_k[4];
  /* flow: cosyvoice/flow/flow.py:296-310, cosyvoice/flow/DiT/dit.py:104-143 */
  int flow_mel, flow_spk_in, flow_vocab, flow_pla_ch, flow_dim, flow_depth, flow_heads, flow_dim_head;
  int flow_ff_mult, flow_chunk, flow_pos_k, flow_pos_groups, flow_noise_frames;
  float flow_cfg_rate;
  int flow_precise;  /* 0: fp16 x fp16 GEMMs (the reference's serving precision); 1: three-term split-fp16 GEMMs (parity mode, <= 1e-3 on mel) */
  /* llm: cosyvoice/llm/llm_multi_head_v3.py:622-689 + HF Qwen2Config */
  int llm_hidden, llm_layers, llm_q_heads, llm_kv_heads, llm_head_dim, llm_inter, llm_text_vocab;
  int llm_speech_vocab, llm_mtp_heads, llm_mtp_inter, llm_max_ctx, llm_max_seqs;
  float llm_rope_theta, llm_eps;
  int llm_kv_f32;   /* 0: KV cache in bf16 (serving default); 1: fp32 (parity mode, tests) */
  /* U-Net estimator: cosyvoice/flow/decoder.py:294-400 with channels == (unet_ch,); 0 channels = stage unused.
   * Precision follows flow_precise. */
  int unet_mel, unet_ch, unet_n_blocks, unet_n_mid, unet_heads, unet_ff_mult, unet_chunk;
  /* unet_noncausal = 1: the non-causal multi-level ConditionalDecoder (cosyvoice/flow/decoder.py:88-291) with unet_levels equal-width
   * levels (channels == (unet_ch,) * unet_levels: stride-2 Conv1d down / ConvTranspose1d(4,2,1) up between levels) and
   * GroupNorm(unet_groups) blocks; 0: the causal single-level variant above. */
  int unet_noncausal, unet_levels, unet_groups;
} hvx_config;

/* Sampler parameters bound per request by server/worker.py:57-65 (ras_sampling keywords,
 * cosyvoice/utils/common.py:138). */
typedef struct hvx_sampler {
  double top_p;   /* doubles: the reference compares fp32 tensors against Python floats (common.py:150,141) */
  double tau_r;
  int top_k;
  int win_size;
} hvx_sampler;

const char* hvx_last_error(void);
int hvx_version(void);

/* lifecycle — replaces ModelManager.load_models object construction
 * (server/model_utils/infer_speech_model.py:50-143). */
hvx_status hvx_create(hvx_engine** out, const hvx_config* cfg);
hvx_status hvx_destroy(hvx_engine* e);

/* Weight ingest — replaces module.load_state_dict(...) in load_models / load_pt
 * (infer_speech_model.py:69-94,169-184).  The engine borrows `dptr` (device memory kept alive
 * by the caller).  Names are the engine's packed-tensor names (flowmirror_hydravox_b200/
 * weights.py documents the mapping from the reference's state_dict keys). */
hvx_status hvx_set_tensor(hvx_engine* e, int stage, const char* name, const void* dptr, int dtype,
                          const int64_t* shape, int ndim);
hvx_status hvx_finalize(hvx_engine* e, int stage);

/* ---- zero-shot frontend features: framed signal x folded linear basis -> magnitude / power -> filterbank -> log.
 * Replaces mel_spectrogram (matcha/utils/audio.py:42-82; the flow's prompt_feat, called from
 * cosyvoice/cli/frontend.py:117-122) and kaldi.fbank + mean subtraction (cosyvoice/cli/frontend.py:108-112).
 * wav_dev (n_samples) fp32; basis_dev [frame_len][2*n_bins] fp32 = everything linear before the non-linearity (window, DC
 * removal, pre-emphasis, DFT: re columns then im columns; flowmirror_hydravox_b200/frontend.py builds it); fb_dev
 * [n_mels][n_bins].  Frames: n_frames = (n_samples + 2*pad_reflect - frame_len) / hop + 1 (checked).  power 0:
 * sqrt(re^2+im^2+mag_eps), 1: re^2+im^2.  out = log(max(fb . spectrum, log_floor)), frame-major [n_frames][n_mels] or
 * channel-major [n_mels][n_frames]; subtract_mean removes the per-bin mean over frames (frame-major only). */
hvx_status hvx_frontend_fbank(hvx_engine* e, const float* wav_dev, int n_samples, int frame_len, int hop, int pad_reflect,
                              const float* basis_dev, int n_bins, const float* fb_dev, int n_mels, int power,
                              float mag_eps, float log_floor, int subtract_mean, int channel_major, float* out_dev,
                              int n_frames, void* stream);

/* Dynamic-range step of whisper.log_mel_spectrogram (called at cosyvoice/cli/frontend.py:95) on the natural-log mel that
 * hvx_frontend_fbank wrote: y = x*scale (scale = 1/ln 10), y = max(y, max(y) - range), y = (y + add) / div, in place. */
hvx_status hvx_frontend_whisper_post(hvx_engine* e, float* logmel_dev, int n, float scale, float range, float add, float div,
                                     void* stream);

/* ---- HiFT: replaces CausalHiFTGenerator.inference (cosyvoice/hifigan/generator.py:713-726) ----
 * mel_dev (mel, T) fp32 -> wav_dev (frame*T') fp32 clamped to +-0.99, src_dev (frame*T) source.
 * finalize=0 follows the streaming branch (:676-679,708-709,725): T' = T-3-4 frames... see DESIGN.md.
 * sine_table_dev: SineGen2.sine_waves rows (n_table_rows, harmonics) uniform[0,1) (generator.py:226); the call fails with
 * HVX_ERR_ARG when the utterance needs more rows than the table holds (the reference raises on the shape mismatch at :306).
 * f0_in_dev (optional, T): pins the F0 track (parity tests); f0_out_dev (optional, T) receives it. */
hvx_status hvx_hift_vocode(hvx_engine* e, const float* mel_dev, int T, int finalize,
                           const float* sine_table_dev, int64_t n_table_rows, const float* f0_in_dev, float* f0_out_dev,
                           float* wav_dev, float* src_dev, void* stream);

/* ---- HiFT, transposed-conv variant: replaces HiFTGenerator.inference (cosyvoice/hifigan/generator.py:557-569; decode
 * :506-540; ConvRNNF0Predictor cosyvoice/hifigan/f0_predictor.py:9-55; SineGen2 with causal=False :233-317) — weight-normed
 * ConvTranspose1d(k, u, padding=(k-u)/2) up-sampling, "same"-padded ResBlocks, F0 predictor on the GPU.
 * noise_dev (frame*T, harmonics) fp32: the Gaussian draw of generator.py:310 made explicit (NULL allowed only when
 * cache_source covers the whole source); cache_source_dev (n_cache samples) overwrites the head of the source (:566-567).
 * The engine's HiFT stage must have been loaded with the transposed packing (weights.pack_hift_t). */
hvx_status hvx_hift_t_vocode(hvx_engine* e, const float* mel_dev, int T, const float* noise_dev,
                             const float* cache_source_dev, int n_cache, const float* f0_in_dev, float* f0_out_dev,
                             float* wav_dev, float* src_dev, void* stream);

/* ---- classic HiFi-GAN: replaces Generator.forward (matcha/hifigan/models.py:148-193; ResBlock1 :14-93) — conv_pre,
 * leaky-relu + weight-normed ConvTranspose1d(k, u, padding=(k-u)/2) up-sampling (v1: rates 8,8,2,2, kernels 16,16,4,4), the mean
 * of the ResBlock1 stacks per stage, conv_post, tanh.  mel_dev (mel, T) fp32 -> wav_dev (prod(rates)*T) fp32 in (-1, 1).
 * The engine's HiFT stage holds weights.pack_hifigan tensors; stage geometry comes from hvx_config.hift_*. */
hvx_status hvx_hifigan_vocode(hvx_engine* e, const float* mel_dev, int T, float* wav_dev, void* stream);

/* ---- flow: replaces CausalMaskedDiffWithDiT.inference (cosyvoice/flow/flow.py:367-430) ----
 * tokens = prompt||new speech tokens (n_prompt + n_tok), embedding (spk_in) fp32,
 * prompt_feat (2*n_prompt, mel) fp32 or NULL; noise_dev = CausalConditionalCFM.rand_noise
 * (mel, noise_frames) fp32 (flow_matching.py:200-201).  mel_out_dev (mel, 2*n_tok) fp32. */
hvx_status hvx_flow_inference(hvx_engine* e, const int32_t* tokens_dev, int n_prompt, int n_tok,
                              const float* embedding_dev, const float* prompt_feat_dev,
                              const float* noise_dev, int n_timesteps, int streaming, int finalize,
                              float* mel_out_dev, void* stream);

/* The same solve for n_utt utterances in ONE pass per Euler step (the reference solves one utterance per call, flow.py:387; a
 * serving batch of similar lengths fills the tensor cores better): utterance u occupies frames [u*Tmax, u*Tmax + T_u) of the
 * ODE state, padding frames are zero, attention keys are masked per utterance.  Arrays of n_utt host-side entries; every pointer
 * inside them is a device pointer as in hvx_flow_inference.  Results equal the per-utterance calls. */
hvx_status hvx_flow_inference_batch(hvx_engine* e, int n_utt, const int32_t* const* tokens_dev, const int* n_prompt,
                                    const int* n_tok, const float* const* embedding_dev,
                                    const float* const* prompt_feat_dev, const float* noise_dev, int n_timesteps,
                                    int streaming, int finalize, float* const* mel_out_dev, void* stream);

/* Estimator seam — replaces ConditionalCFM.forward_estimator's TensorRT branch
 * (cosyvoice/flow/flow_matching.py:126-153): x,mu,cond (2,mel,T), t (2), spks (2,mel) fp32;
 * writes dphi/dt (2,mel,T) into out_dev. */
hvx_status hvx_dit_estimator(hvx_engine* e, const float* x_dev, const float* mu_dev, const float* t_dev,
                             const float* spks_dev, const float* cond_dev, int T, int streaming,
                             float* out_dev, void* stream);

/* U-Net estimator at the same seam — replaces CausalConditionalDecoder.forward (cosyvoice/flow/decoder.py:405-494) behind
 * ConditionalCFM.forward_estimator (flow_matching.py:126-153): x, mu, cond (2, mel, T), t (2), spks (2, mel) fp32 -> dphi/dt
 * (2, mel, T) in out_dev.  The seam's mask is all-true at inference and is not an argument.  Weights: stage HVX_STAGE_UNET
 * (weights.pack_unet). */
hvx_status hvx_unet_estimator(hvx_engine* e, const float* x_dev, const float* mu_dev, const float* t_dev,
                              const float* spks_dev, const float* cond_dev, int T, int streaming,
                              float* out_dev, void* stream);
/* The same seam for the non-causal multi-level ConditionalDecoder with its padding mask (cosyvoice/flow/decoder.py:210-291,
 * forward(x, mask, mu, t, spks, cond)): mask_dev (2, 1, T) fp32 0/1, every row a prefix mask as make_pad_mask builds them, or NULL =
 * all true (what solve_euler passes, flow_matching.py:104).  dump_dev / n_dump as in hvx_unet_estimator_debug (NULL / 0: off). */
hvx_status hvx_unet_estimator_masked(hvx_engine* e, const float* x_dev, const float* mask_dev, const float* mu_dev, const float* t_dev,
                                     const float* spks_dev, const float* cond_dev, int T, float* out_dev, float* dump_dev, int n_dump,
                                     void* stream);
/* same, and copies the fp32 residual stream (2T, unet_ch) after every resnet / transformer block into dump_dev (parity tests) */
hvx_status hvx_unet_estimator_debug(hvx_engine* e, const float* x_dev, const float* mu_dev, const float* t_dev,
                                    const float* spks_dev, const float* cond_dev, int T, int streaming,
                                    float* out_dev, float* dump_dev, int n_dump, void* stream);

/* The reference's own plug-in seam, as ConditionalCFM.forward_estimator drives a TensorRT execution context
 * (cosyvoice/flow/flow_matching.py:126-153, pool object cosyvoice/utils/common.py:198-213): raw device addresses in the flow's
 * serving dtype (spks.dtype: HVX_F32, HVX_F16 or HVX_BF16) for x (2, mel, T), mu, t (2), spks (2, mel), cond; the result goes to
 * out_dev, which the reference binds to x itself (its 7th address is x.data_ptr()), so out_dev == x_dev is allowed.  The seam's
 * mask is all-true at inference (flow_matching.py:104-111) and is not an argument.  kind 0: DiT estimator, 1: U-Net estimator.
 * flowmirror_hydravox_b200/flow.py: NativeEstimatorPool wraps this as acquire_estimator()/release_estimator() + a context with
 * set_input_shape / set_tensor_address / execute_async_v3. */
hvx_status hvx_estimator_seam(hvx_engine* e, int kind, const void* x_dev, const void* mu_dev, const void* t_dev,
                              const void* spks_dev, const void* cond_dev, void* out_dev, int T, int dtype, int streaming,
                              void* stream);

/* CFM Euler solve over the U-Net estimator — replaces CausalConditionalCFM.forward + ConditionalCFM.solve_euler
 * (cosyvoice/flow/flow_matching.py:203-228,71-124) when the estimator is the U-Net: z = noise[:, :T] * temperature, cosine
 * t-schedule, per step CFG staging (row 1 has mu/spks/cond zeroed), estimator, v = (1+cfg)*v0 - cfg*v1, x += dt*v.
 * mu_dev, cond_dev (mel, T), spks_dev (mel) fp32 (cond/spks may be NULL = zeros); noise_dev (mel, noise_ld) = rand_noise;
 * mel_out_dev (mel, T) fp32.  cfg rate = hvx_config.flow_cfg_rate. */
hvx_status hvx_cfm_solve_unet(hvx_engine* e, const float* mu_dev, const float* spks_dev, const float* cond_dev,
                              const float* noise_dev, int noise_ld, int T, int n_timesteps, float temperature,
                              int streaming, float* mel_out_dev, void* stream);

/* ---- LLM: replaces CosyVoice3LM.inference / inference_wrapper
 * (cosyvoice/llm/llm_multi_head_v3.py:861-960).
 * hvx_llm_begin resets sequence slot `seq` and stores its prompt rows
 * [sos, embed(prompt_text||text), task_id, speech_emb(prompt_speech)] (:943-952).
 * hvx_llm_generate runs prefill + the multi-head loop for all begun sequences until each has
 * stopped (stop token / max_len) and writes tokens to out_tokens_dev[seq*max_out + i], counts to
 * out_counts_dev[seq].  u_dev: uniform stream per sequence (n_seq, u_stride) consumed in call
 * order by the sampler (oracle/llm_ref.py documents the order). */
hvx_status hvx_llm_begin(hvx_engine* e, int seq, const int32_t* text_ids_dev, int n_text_total,
                         int n_text_new, const int32_t* prompt_speech_dev, int n_prompt_speech,
                         float min_ratio, float max_ratio);
hvx_status hvx_llm_generate(hvx_engine* e, int n_seq, int head_k, const hvx_sampler* sp,
                            const float* u_dev, int u_stride, int32_t* out_tokens_dev, int max_out,
                            int32_t* out_counts_dev, void* stream);
/* Ask a hvx_llm_generate that is running on another host thread to stop after its current batch of decode steps (<= 16) and
 * return the tokens emitted so far — an abandoned streaming request (client disconnect) must not keep writing into the
 * token buffers the next request reuses.  Lock-free; the next hvx_llm_generate clears the flag. */
hvx_status hvx_llm_cancel(hvx_engine* e);
/* Teacher-forced probe for parity tests: final-normed hidden of the last prompt row and the
 * log-softmax of every MTP head on it (llm_multi_head_v3.py:886-888). */
hvx_status hvx_llm_probe(hvx_engine* e, int seq, float* last_hidden_dev, float* head_logp_dev, void* stream);
/* Sampler alone on caller-provided log-probs (n_heads, vocab) — parity of sampling_ids
 * (llm_multi_head_v3.py:151-166) independent of the transformer numerics. */
hvx_status hvx_sample(hvx_engine* e, const float* logp_dev, int n_heads, const int32_t* history_dev,
                      int n_history, int min_len, const hvx_sampler* sp, const float* u_dev, int n_u,
                      int32_t* out_ids_dev, int32_t* u_used_dev, void* stream);

/* ---- end to end with HOST buffers: replaces inference_zero_shot / inference_tts
 * (server/model_utils/infer_speech_model.py:523-689) from token ids to waveform.
 * Copies inputs H2D, runs LLM -> flow -> (speed interp) -> HiFT, copies wav D2H, synchronises. */
typedef struct hvx_request {
  const int32_t* text_ids_host;      int n_text_total; int n_text_new;   /* prompt_text||text */
  const int32_t* prompt_speech_host; int n_prompt_speech;
  const float* prompt_feat_host;                                         /* (2*n_prompt_speech, mel) or NULL */
  const float* embedding_host;                                           /* (spk_in) */
  const float* u_host;               int n_u;
  float min_ratio, max_ratio;
  double speed;   /* double: the reference evaluates int(T / speed) on Python floats (infer_speech_model.py:584-587) */
} hvx_request;
hvx_status hvx_synthesize_host(hvx_engine* e, const hvx_request* reqs, int n_req, int head_k,
                               const hvx_sampler* sp, int n_timesteps, const float* noise_dev,
                               const float* sine_table_dev, int64_t n_table_rows, float* wav_host, int wav_stride,
                               int32_t* wav_len_host, int32_t* tokens_host, int tok_stride,
                               int32_t* n_tokens_host, float* stage_ms_host, void* stream);

/* Incremental streaming flow: replaces the repeated `flow.inference(token=all tokens so far, streaming=True, finalize=False)` calls of
 * CosyVoice2Model.tts / token2wav (cosyvoice/cli/model.py:279-297,330-348), whose result is sliced to the new frames
 * (`tts_mel[:, :, token_offset * token_mel_ratio:]`).  A session caches, per Euler step and DiT layer, the keys / values of the
 * frames already evaluated (valid under the block-causal chunk mask, cosyvoice/utils/mask.py:127-236), and every append
 * evaluates only the new chunk(s).  begin: n_timesteps Euler steps, room for max_frames mel frames (prompt included).
 * append: tokens_dev = prompt + new tokens so far incl. the 3 look-ahead tokens; 2 * (n_prompt + n_tok - 3) must be a multiple of
 * the chunk size (the reference's hop schedule guarantees it); mel_out_dev (mel, n_new) receives the new frames,
 * *n_new_frames_host their count.  The final chunk (finalize=True) is a full-attention pass by the reference's own semantics
 * (cli/model.py:352-358): use hvx_flow_inference for it. */
hvx_status hvx_flow_stream_begin(hvx_engine* e, int n_timesteps, int max_frames, void* stream);
hvx_status hvx_flow_stream_append(hvx_engine* e, const int32_t* tokens_dev, int n_prompt, int n_tok, const float* embedding_dev,
                                  const float* prompt_feat_dev, const float* noise_dev, float* mel_out_dev, int* n_new_frames_host,
                                  void* stream);
hvx_status hvx_flow_stream_end(hvx_engine* e);

/* speed control: replaces F.interpolate(tts_mel, size=int(T/speed), mode='linear')
 * (infer_speech_model.py:584-587,662-665).  mel_dev (C, T) -> out_dev (C, T_out). */
hvx_status hvx_speed_interp(hvx_engine* e, const float* mel_dev, int C, int T, int T_out, float* out_dev, void* stream);

/* streaming overlap: replaces fade_in_out (cosyvoice/utils/common.py:169-177; called by CosyVoice2Model.token2wav,
 * cosyvoice/cli/model.py:295-308).  wav_dev[0:overlap) = wav_dev * window[0:overlap) + prev_tail_dev[0:overlap) * window[overlap:2*overlap),
 * evaluated in double like the reference's float32-tensor x float64-window product.  window_dev: 2*overlap doubles (np.hamming). */
hvx_status hvx_fade_in_out(hvx_engine* e, float* wav_dev, const float* prev_tail_dev, const double* window_dev, int overlap, void* stream);

/* ---- diagnostic entries for the kernel-level parity tests (tests/test_gemm_gpu.py) ----
 * C = act(A*B^T + bias): A (M,K) bf16, B (N,K) bf16 (nn.Linear weight layout), C bf16 or fp32.  out_f32 bit 0: fp32 output,
 * bit 1: fp16 operands, bit 2: split-precision operands A (M,2K) = [hi | lo], B (N,2K) = [hi | lo] -> three-term product. */
hvx_status hvx_gemm_bf16(hvx_engine* e, const void* A_dev, const void* B_dev, const float* bias_dev, void* C_dev,
                         int M, int N, int K, int out_f32, int act, void* stream);
/* softmax(q k^T / 8 + mask) v per (batch, head): qk (B*T, 2*H*64) = q|k, vt (B*H*64, vt_ld) = V^T. */
hvx_status hvx_attention_bf16(hvx_engine* e, const void* qk_dev, const void* vt_dev, int vt_ld, void* out_dev,
                              int B, int T, int H, int chunk, void* stream);

/* bench.py roofline helper: one class of decode-step kernels (0 whole step w/o sampler, 1 qkv, 2 attention, 3 o-proj,
 * 4 gate-up, 5 down, 6 MTP heads + logits, 7 whole step as the fused persistent kernel) repeated `reps` times on the engine's stream, CUDA-event timed;
 * ms_out[0] = ms per repetition (all layers' launches of that class). */
hvx_status hvx_llm_bench_kernels(hvx_engine* e, int n_seq, int head_k, int ctx, int which, int reps, float* ms_out);

/* Parity probe for the two decode-step implementations: one step of sequence slot 0 at context length ctx through the
 * kernel-per-op path (use_fused=0) or the persistent fused kernel (use_fused=1); n_layers>0 stops after that many layers.
 * h_out_dev [head_k][hidden], logits_out_dev [head_k][speech_vocab] (full step only). */
hvx_status hvx_llm_debug_step(hvx_engine* e, int head_k, int ctx, int use_fused, int n_layers, float* h_out_dev,
                              float* logits_out_dev);

/* Live per-class kernel timing for bench.py's roofline: while enabled, the launches of each class are bracketed by CUDA events on
 * their own stream (not under graph capture).  hvx_profile_collect synchronises the device and returns, per class, the summed
 * event-to-event milliseconds, the summed algorithmic work (FLOP for the tensor classes, bytes for the bandwidth classes) and
 * the number of bracketed call sites since the last collect, then resets.  Arrays of HVX_PROF_NCLS entries. */
enum { HVX_PROF_GEMM = 0,        /* tcgen05 GEMMs / implicit-GEMM convolutions: 2*M*N*K of the mathematical product */
       HVX_PROF_ATTN = 1,        /* DiT / U-Net attention: 4 * queries * visible keys * 64 per head */
       HVX_PROF_HIFT_CONV = 2,   /* HiFT convolutions: 2 * Cin * K * Cout * Lout */
       HVX_PROF_LLM_STEP = 3,    /* one decode step (graph launch): work = 1 per step, bytes are computed by the caller */
       HVX_PROF_LAYERNORM = 4,   /* adaLN LayerNorm + modulate: bytes read + written */
       HVX_PROF_LLM_PREFILL = 5, /* prompt prefill of one sequence: 2 * rows * weights */
       HVX_PROF_NCLS = 8 };
hvx_status hvx_profile_enable(hvx_engine* e, int on);
hvx_status hvx_profile_collect(hvx_engine* e, double* ms_out, double* work_out, int64_t* launches_out);

/* bookkeeping for bench.py: number of kernels this library has launched since creation. */
int64_t hvx_kernel_launches(hvx_engine* e);

#ifdef __cplusplus
}
#endif
#endif
