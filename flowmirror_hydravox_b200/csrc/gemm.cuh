// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores:  C[M,N] = A[M,K] * B[N,K]^T (+ fused epilogue)
// A = activations (row-major, K contiguous), B = nn.Linear weight (out x in, K contiguous).
// One CTA per 128 x BN output tile: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one
// elected lane) + TMEM owner, warps 2-5 = epilogue (tcgen05.ld -> registers -> fused math -> global).
// Operands arrive by TMA with 128B swizzle into a STAGES-deep smem ring guarded by mbarriers; the
// accumulator lives in TMEM (BN fp32 columns x 128 lanes).  Two CTAs fit per SM (3 x 32 KB ring),
// so one tile's epilogue overlaps the neighbour's main loop without a persistent scheduler.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"
#include "llm_common.cuh"

namespace hvx {

enum { EPI_BF16 = 0, EPI_F32 = 1, EPI_RESID_GATE = 2, EPI_QKV = 3, EPI_LLM_QKV = 4, EPI_SWIGLU = 5, EPI_HIFT = 6 };
enum { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_SILU = 2, ACT_MISH = 3, ACT_LRELU = 4 /* slope 0.01 */, ACT_GELU_ERF = 5 /* exact, F.gelu default */ };

// EPI_HIFT — epilogue of the HiFT implicit-GEMM convolutions (hift.cu; generator.py:110-117,672-711).  Per output element
//   v = acc + bias (+ resid[row][col]) (+ resid2[row][col])
//   out32 (optional, fp32 [rows][ld32]):  o = v * out_scale (+ out32's old value when accumulate); v := o
//       rows are shifted by row_shift, and row 1 is also stored to row 0 when dup_row1 (ReflectionPad1d((1, 0)), generator.py:686)
//   out16[k] (k < n16, fp16 [rows][ld16] as hi at [col], lo = a - hi at [lo_off + col]):  a = act_k(v) with
//       act 0: identity, 1: leaky-relu(slope), 2: Snake  a = v + inv_alpha[col] * sin(alpha[col] * v)^2   (activation.py:79-84)
struct HiftEpi {
  const float* resid = nullptr; const float* resid2 = nullptr; int ldr = 0;
  float* out32 = nullptr; int ld32 = 0; float out_scale = 1.f; int accumulate = 0; int row_shift = 0; int dup_row1 = 0;
  int n16 = 0; uint16_t* out16[3] = {nullptr, nullptr, nullptr}; int act[3] = {0, 0, 0};
  const float* alpha[3] = {nullptr, nullptr, nullptr}; const float* inv_alpha[3] = {nullptr, nullptr, nullptr};
  float slope = 0.f; int ld16 = 0; int lo_off = 0;
};

struct GemmEpi {
  int mode = EPI_BF16;
  int act = ACT_NONE;
  int f16 = 0;                      // 1: operands and 16-bit outputs are IEEE fp16 instead of bf16 (flow stage)
  const float* bias = nullptr;      // [N]
  void* out = nullptr;              // bf16 (EPI_BF16/EPI_QKV) or fp32
  int ldo = 0;
  // EPI_RESID_GATE: out_f32[row][col] += gate[(row / rows_per_batch) * gate_ld + col] * (acc + bias)
  const float* gate = nullptr;
  int gate_ld = 0;
  int rows_per_batch = 1;
  // EPI_QKV (DiT): cols < n_qk -> out (rotary on the first 64 channels of q and of k, interleaved
  // pairs, DiT/modules.py:367-373); cols >= n_qk -> V transposed per (batch, head): vt[(b*H+h)*64+d][t]
  __nv_bfloat16* vt = nullptr;
  int vt_ld = 0, T = 1, heads = 1, n_qk = 0;
  const float* rope_cos = nullptr;  // [T][32]
  const float* rope_sin = nullptr;
  // EPI_QKV, windowed call of the incremental streaming flow (flow.cu: flow_nfe_window): the rows are frames t_off .. t_off + T of
  // every batch row; q goes to `out` (ldo wide, window-local rows), k to the session's key cache k_out[(b * k_batch_rows + t_off + t)]
  // (k_ld wide) and V^T to column t_off + t of vt; rotary positions are absolute (t_off + t).  k_out == nullptr: q | k side by side in out.
  __nv_bfloat16* k_out = nullptr;
  int k_ld = 0, k_batch_rows = 0, t_off = 0;
  // EPI_F32 extras: out_f32 = act(acc+bias) (+ resid[row][col]); optional bf16 copy of the same value
  const float* resid = nullptr;
  __nv_bfloat16* out2 = nullptr;
  // 16-bit outputs (EPI_BF16, EPI_SWIGLU, out2) written as hi at [col] and lo = x - hi at [lo_off + col] when lo_off > 0
  int lo_off = 0;
  int ldo2 = 0;                     // row stride of out2 (0 -> ldo)
  // EPI_LLM_QKV (Qwen2 prefill / batched decode): RoPE + KV-cache write, see llm_common.cuh
  LlmQkvEpi llm;
  // EPI_SWIGLU: weight rows interleaved (gate_n, up_n); out16[row][n] = silu(acc[2n]) * acc[2n+1], ldo = N/2
  HiftEpi hift;                     // EPI_HIFT
};

// A-operand addressing.  A is viewed as [n_batch][a_rows][lda]; output tiles never straddle a batch and
// rows outside [0, a_rows) of the batch read as zero (TMA out-of-bounds fill), which is what gives
// convolutions their zero padding.  k-block kb reads
//   columns a_col0 + (n0/BN)*a_col_per_ntile + (kb % kb_per_tap)*64,  rows m0 + a_row0 + (kb / kb_per_tap)*a_row_step
// (kb_per_tap == 0: plain GEMM, columns kb*64, rows m0):
//   Conv1d over frame-major activations as implicit GEMM: K = taps * C, kb_per_tap = C/64, a_row_step = 1,
//     a_row0 = -(left pad)                       (PreLookaheadLayer, upsample_encoder.py:82-103)
//   causal grouped conv, 64-channel groups: kb_per_tap = 1, a_col_per_ntile = 64 (BN = 64 -> one group per n-tile)
//                                                (CausalConvPositionEmbedding, DiT/modules.py:115-144)
struct GemmAddr {
  int n_batch = 1;
  int rows_per_batch = 0;      // output rows per batch; 0 -> M
  int a_rows = 0;              // rows of the A tensor per batch; 0 -> rows_per_batch
  int a_cols = 0;              // width of the A matrix in elements; 0 -> K
  int a_col0 = 0, a_col_per_ntile = 0;
  int kb_per_tap = 0;
  int a_row0 = 0, a_row_step = 0;
  // split-precision activations: A = [hi | lo] (two bf16 halves of an fp32 value, K' = 2K) against the same
  // weights: k-block kb of A multiplies k-block kb % b_kb_mod of B (0 = off)
  int b_kb_mod = 0;
  // three-term split precision (flow parity mode): A = [hi | lo] (2K0 wide), B = [hi | lo] (2K0 wide), K' = 3*K0:
  //   k-blocks [0,n0) -> A_hi*B_hi, [n0,2n0) -> A_lo*B_hi, [2n0,3n0) -> A_hi*B_lo   with n0 = split3_kb = K0/64 (0 = off)
  int split3_kb = 0;
  int a_lo_off = 0;            // element offset of the lo half inside an A row (0 -> split3_kb * 64); convs: the row width
  // split-K for skinny problems (few output tiles, long K): split_k CTAs per output tile (grid.z) each accumulate a contiguous
  // share of the k-blocks and store an fp32 partial to out + z * split_stride (EPI_F32 without resid, bias from split 0);
  // the caller sums the partials in a fixed order (deterministic, no atomics)
  int split_k = 1;
  size_t split_stride = 0;
  // 1: launch without programmatic dependent launch.  The batched decode step is sized so that its GEMMs fill the SMs in exactly one
  // wave at two CTAs per SM; successors that become resident early (waiting in griddepcontrol.wait with their shared memory and
  // TMEM held) break that wave in two: 3.24 -> 3.54 ms per 128-row step with PDL on the decode GEMMs.
  int no_pdl = 0;
  // HVX_GEMM_TIMELINE=1 (diagnostic): %globaltimer stamps of CTA (0,0,0) of the tile kernel — [0] start, [1+kb] TMA of k-block kb
  // issued, [65+kb] its data landed (MMA side), [130] accumulator complete, [131] epilogue done
  unsigned long long* dbg = nullptr;
};

hvx_status gemm_bf16(hvx_engine* e, cudaStream_t st, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                     int M, int N, int K, const GemmEpi& epi, const GemmAddr* addr = nullptr);

}  // namespace hvx
