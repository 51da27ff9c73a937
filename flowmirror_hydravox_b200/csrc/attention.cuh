// DiT self-attention on tcgen05 (replaces F.scaled_dot_product_attention in
// cosyvoice/flow/DiT/modules.py:349-407; masks from cosyvoice/utils/mask.py:127-236).
//
// One CTA = (batch, head, 128-query tile): a TMA + tensor-core warp and four softmax warps (TMEM lane == query row == thread, so
// the online softmax needs no cross-thread reduction), 64-key tiles:
//   S_j (128x64 fp32, TMEM, 2 buffers)   = Q K_j^T      tcgen05.mma M128 N64 K16 x4, Q and K_j tiles by TMA (SW128)
//   P_j (16-bit, TMEM, 2 buffers)        = exp2(S_j*c - m) stored packed K-major by the row's own thread (tcgen05.st), lazy rescale
//   O (128x64 fp32, TMEM)               += P_j V_j      tcgen05.mma M128 N64 K16 x4, A from TMEM, V_j^T tile by TMA
// Four key slots and three value slots form separate TMA rings; two CTAs co-reside per SM (73 KB smem, 256 TMEM columns each) so one
// CTA's softmax overlaps the other's MMAs.  (Q as a TMEM operand too, with a single P buffer to make room: measured slower, 516 vs
// 568 TFLOP/s at T = 2298.)
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace hvx {

struct AttnArgs {
  int T;            // frames per batch row (queries == keys)
  int heads;
  int n_batch;
  int chunk;        // >0: streaming block-causal mask, key j visible to query i iff j < (i/chunk+1)*chunk
  int f16;          // 1: q/k/v/out are fp16 instead of bf16
  int ld_out;       // = heads*64 (row stride is 2*ld_out when lo_off > 0)
  const int* klen = nullptr;   // [n_batch] valid keys per batch row (device; ragged groups, v5 kernel only); nullptr: T
  double work = 0;  // FLOP of this call for the live profiler (0: 4 * n_batch * heads * T^2 * 64, halved under the chunk mask)
  // windowed call (incremental streaming flow): Tq query rows per batch row — frames q_pos0 .. q_pos0 + Tq of the sequence, read
  // from `qk` (Tq rows per batch) — against the first tk of T key rows per batch row held in k_ptr (ld_k wide) / vt
  int Tq = 0, q_pos0 = 0, tk = 0;
  const __nv_bfloat16* k_ptr = nullptr; int ld_k = 0;
  int lo_off = 0;   // > 0: output written as split precision, hi at [col], lo at [lo_off + col] (flow parity mode, v5 kernel)
  __nv_bfloat16* out;   // [n_batch*T][heads*64]
};

hvx_status dit_attention(hvx_engine* e, cudaStream_t st, const __nv_bfloat16* qk, int ld_qk, int k_col0,
                         const __nv_bfloat16* vt, int vt_ld, const AttnArgs& a);

}  // namespace hvx
