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

namespace hvx {

enum { EPI_BF16 = 0, EPI_F32 = 1, EPI_RESID_GATE = 2, EPI_QKV = 3 };
enum { ACT_NONE = 0, ACT_GELU_TANH = 1, ACT_SILU = 2, ACT_MISH = 3 };

struct GemmEpi {
  int mode = EPI_BF16;
  int act = ACT_NONE;
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
  // EPI_F32 extras: out_f32 = act(acc+bias) (+ resid[row][col]); optional bf16 copy of the same value
  const float* resid = nullptr;
  __nv_bfloat16* out2 = nullptr;
};

// A-operand addressing.  A is viewed as [n_batch][rows_per_batch][lda] and tiles never straddle a
// batch, so TMA zero-fills above/below each batch row range.  k-block kb reads columns
// a_col0 + (n0/BN)*a_col_per_ntile + kb*a_col_step and rows m0 + a_row0 + kb*a_row_step:
//   plain GEMM      : {0, 0, 64, 0, 0}
//   causal grouped conv as implicit GEMM (CausalConvPositionEmbedding, DiT/modules.py:115-144):
//     one k-block per tap, A rows shifted by the tap, columns = the group's 64 channels.
struct GemmAddr {
  int n_batch = 1;
  int rows_per_batch = 0;      // 0 -> M
  int a_col0 = 0, a_col_per_ntile = 0, a_col_step = 64, a_row0 = 0, a_row_step = 0;
  int a_cols = 0;              // 0 -> K (width of the A matrix in elements)
};

hvx_status gemm_bf16(hvx_engine* e, cudaStream_t st, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                     int M, int N, int K, const GemmEpi& epi, const GemmAddr* addr = nullptr);

}  // namespace hvx
