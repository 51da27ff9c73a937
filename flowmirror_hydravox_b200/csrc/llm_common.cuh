// Device-side state of the multi-head AR decode shared by the GEMV kernels (llm.cu) and the
// tensor-core GEMM epilogues (gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace hvx {

constexpr int LLM_MAX_HEADS = 8;       // inference_head_num <= 5 in the reference UI (inference_tab.py:555-561)

// One sequence slot (device memory).  Row slots of a decode step are [seq * head_k + r], r < n_new.
struct SeqState {
  int ctx;            // rows already in the KV cache
  int ctx_add;        // rows the running step appends (applied by the sampler kernel at the end of the step)
  int n_new;          // valid rows of this sequence in the running step
  int n_out;          // speech tokens emitted so far
  int min_len, max_len;
  int done;           // 1 once a stop token was sampled or max_len reached (llm_multi_head_v3.py:902-916)
  int u_pos;          // next unread element of this sequence's uniform stream
  int status;         // 0 ok, 1 = 100 EOS retries exhausted (llm_multi_head_v3.py:158-166), 2 = u-stream exhausted
  int new_tok[LLM_MAX_HEADS];
};

// Fused epilogue of the QKV projection: RoPE (HF half-split convention) on q and k, q to the q buffer,
// k and v straight into the KV cache as bf16.  Weight rows of q and k are permuted per head at pack
// time so that output features (2i, 2i+1) hold the RoPE pair (d=i, d=i+32); q.k is invariant under a
// permutation applied to both.
struct LlmQkvEpi {
  float* q_out = nullptr; int ldq = 0;          // [rows][q_dim] fp32
  void* kc = nullptr;                           // this layer: [seq][kv_head][max_ctx][64], bf16 (or fp32 if kv_f32)
  void* vc = nullptr;
  int kv_f32 = 0;
  size_t seq_stride = 0;                        // kv_heads * max_ctx * 64
  int max_ctx = 0, q_dim = 0, kv_dim = 0;
  const float* inv_freq = nullptr;              // [32]
  const SeqState* seqs = nullptr;               // decode: per-sequence state; nullptr: prefill of one sequence
  int rows_per_seq = 1;                         // decode: head_k
  int seq0 = 0, pos0 = 0;                       // prefill: sequence slot, position of row 0
  int n_rows = 0;                               // valid rows (row slots beyond are padding)
};

// one out-of-line copy of the accurate sincosf (its slow path is ~100 instructions; 16 inlined copies per chunk bloat the epilogue)
static __device__ __noinline__ void sincos_noinline(float a, float* sn, float* cs) { sincosf(a, sn, cs); }

// store the pair (n, n+1) (n even) of row `row`
__device__ __forceinline__ void llm_qkv_store(const LlmQkvEpi& q, int row, int n, float v0, float v1) {
  int seq, pos;
  if (row >= q.n_rows) return;
  if (q.seqs) {
    seq = row / q.rows_per_seq;
    const int r = row - seq * q.rows_per_seq;
    const SeqState& s = q.seqs[seq];
    if (s.done || r >= s.n_new) return;
    pos = s.ctx + r;
  } else {
    seq = q.seq0;
    pos = q.pos0 + row;
  }
  if (pos >= q.max_ctx) return;
  const bool is_q = n < q.q_dim;
  const bool is_k = !is_q && n < q.q_dim + q.kv_dim;
  if (is_q || is_k) {
    const int i = (n & 63) >> 1;
    const float a = (float)pos * q.inv_freq[i];
    float sn, cs;
    sincos_noinline(a, &sn, &cs);
    const float r0 = v0 * cs - v1 * sn;
    const float r1 = v1 * cs + v0 * sn;
    v0 = r0; v1 = r1;
  }
  if (is_q) {
    *reinterpret_cast<float2*>(q.q_out + (size_t)row * q.ldq + n) = make_float2(v0, v1);
  } else {
    const int m = n - q.q_dim - (is_k ? 0 : q.kv_dim);
    const int kvh = m >> 6, d = m & 63;
    const size_t idx = (size_t)seq * q.seq_stride + ((size_t)kvh * q.max_ctx + pos) * 64 + d;
    void* base = is_k ? q.kc : q.vc;
    if (q.kv_f32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + idx) = make_float2(v0, v1);
    else *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = __floats2bfloat162_rn(v0, v1);
  }
}

}  // namespace hvx
