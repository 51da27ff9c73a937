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
  const float2* rope = nullptr;                 // optional [max_ctx][32] (cos, sin) of pos * inv_freq[i], made with the same sincosf
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
    float sn, cs;
    if (q.rope) { const float2 t = __ldg(q.rope + (size_t)pos * 32 + i); cs = t.x; sn = t.y; }      // same values as the table path of llm_qkv_store32
    else sincos_noinline((float)pos * q.inv_freq[i], &sn, &cs);
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

// The same epilogue for a whole 32-column chunk of one row (tensor-core path: thread == row): columns [n0, n0 + 32) never
// straddle a q / k / v or head boundary (all are multiples of 64), so the rotated values go out as eight 16-byte stores instead
// of 16 scattered 8-byte ones — with one row per thread every store instruction touches 32 different rows, and the narrow
// stores made this epilogue longer than the whole k-loop of the 128-row decode GEMM (profiles/README.md, round 2).
__device__ __forceinline__ void llm_qkv_store32(const LlmQkvEpi& q, int row, int n0, float* f) {
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
  const bool is_q = n0 < q.q_dim;
  const bool is_k = !is_q && n0 < q.q_dim + q.kv_dim;
  if (is_q || is_k) {
    const int i0 = (n0 & 63) >> 1;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      float sn, cs;
      if (q.rope) { const float2 t = __ldg(q.rope + (size_t)pos * 32 + i0 + j); cs = t.x; sn = t.y; }
      else sincos_noinline((float)pos * q.inv_freq[i0 + j], &sn, &cs);
      const float v0 = f[2 * j], v1 = f[2 * j + 1];
      f[2 * j] = v0 * cs - v1 * sn;
      f[2 * j + 1] = v1 * cs + v0 * sn;
    }
  }
  if (is_q) {
    float4* o = reinterpret_cast<float4*>(q.q_out + (size_t)row * q.ldq + n0);
#pragma unroll
    for (int j = 0; j < 8; j++) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
  } else {
    const int m = n0 - q.q_dim - (is_k ? 0 : q.kv_dim);
    const int kvh = m >> 6, d = m & 63;
    const size_t idx = (size_t)seq * q.seq_stride + ((size_t)kvh * q.max_ctx + pos) * 64 + d;
    void* base = is_k ? q.kc : q.vc;
    if (q.kv_f32) {
      float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(base) + idx);
#pragma unroll
      for (int j = 0; j < 8; j++) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    } else {
      uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + idx);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const __nv_bfloat162 a = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]), b = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
        const __nv_bfloat162 c2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]), d2 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
        o[j] = make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                          *reinterpret_cast<const uint32_t*>(&c2), *reinterpret_cast<const uint32_t*>(&d2));
      }
    }
  }
}

}  // namespace hvx
