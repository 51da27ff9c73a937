// Multi-head autoregressive speech-token decode (Qwen2 backbone + MTP heads + RAS sampler) for sm_100a.
// Replaces CosyVoice3LM.inference / inference_wrapper (cosyvoice/llm/llm_multi_head_v3.py:861-960),
// Qwen2Encoder.forward_one_step (:248-260, HF Qwen2), the MTP heads (:657-667,886-888) and
// sampling_ids / ras_sampling / nucleus_sampling / random_sampling (:151-166, cosyvoice/utils/common.py:138-166)
// — restated in oracle/llm_ref.py.
//
// The reference recomputes the whole prefix every step with no KV cache (:871-882); this engine keeps a
// KV cache resident in HBM and feeds only the head_k rows emitted by the previous step (identical under
// causal attention).  Everything that happens between two steps — sampling, stop logic, history, the
// embedding fetch of the new rows — runs on the device, so a step is a fixed launch sequence with no
// host round trip; the host only polls a "sequences still running" counter every few steps.
//
// Data layout:
//   weights             bf16 [out][in] (nn.Linear layout); q/k rows permuted per head so RoPE pairs are
//                       adjacent; gate/up rows interleaved so SwiGLU fuses into the producing kernel
//   KV cache            bf16 [layer][seq][kv_head][max_ctx][64]   (a key is one 128 B line)
//   residual stream h   fp32 [seq*head_k + r][hidden]
// Decode linears are HBM-bound weight streams: llm_gemv_kernel reads every weight row once with 128-bit
// no-allocate loads and applies it to all <=8 live rows held in shared memory (RMSNorm fused into the
// prologue, bias / RoPE+cache write / SwiGLU / residual fused into the epilogue).  Prefill and decode with
// more than 8 live rows go through the tcgen05 GEMM (gemm.cu) with the same fused epilogues.
#include "gemm.cuh"
#include "llm_common.cuh"
#include <algorithm>
#include <type_traits>
#include <cmath>
#include <cstring>
#include <cstdlib>

namespace hvx {

constexpr int GEMV_KC = 2048;          // activation chunk (floats per row) staged in shared memory
constexpr int ATT_KEYS = 128;          // keys per shared-memory chunk (bf16 cache; 64 for the fp32 cache: static smem limit)
constexpr int ATT_LD = 72;             // padded row (bf16 elements): 144 B stride -> conflict-free 16 B reads
constexpr int ATT_LD32 = 68;           // fp32 cache: 272 B stride
constexpr int SAMP_MAXK = 64;          // top_k cap (UI slider goes to 50)

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// 1-D bulk copy global -> shared (TMA engine, no tensor map): bytes % 16 == 0, both addresses 16 B aligned
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
// fp32 value as two bf16 halves: row[k] = hi, row[off_lo + k] = x - hi
__device__ __forceinline__ void store_split(__nv_bfloat16* row, int off_lo, int k, float v) {
  const __nv_bfloat16 hi = __float2bfloat16(v);
  row[k] = hi;
  row[off_lo + k] = __float2bfloat16(v - __bfloat162float(hi));
}

// ------------------------------------------------------------------ GEMV over <= 8 rows
enum { GEMV_PLAIN = 0, GEMV_SWIGLU = 1, GEMV_QKV = 2 };

struct GemvArgs {
  const float* x = nullptr; int ldx = 0;            // [R][K] fp32
  const __nv_bfloat16* W = nullptr;                 // [N][K]
  const float* bias = nullptr;                      // [N]
  const float* norm_w = nullptr; float eps = 1e-6f; // x <- rmsnorm(x) * norm_w first (needs K <= GEMV_KC)
  float* out = nullptr; int ldo = 0;
  const float* resid = nullptr; int ldr = 0;        // out = resid + y (resid may alias out)
  int N = 0, K = 0, mode = GEMV_PLAIN;
  int rows = 0;                                     // live rows (<= R); rows beyond are not written
  int row0 = 0;                                     // global row slot of row 0 (row groups of 8, see launch_gemv)
  int kc = 0;                                       // activation chunk staged in shared memory (floats per row)
  int slot_bytes = 0, n_slots = 0;                  // TMA-fed variant: weight ring geometry
  int piece_bytes = 0;                              // bulk-copy granularity inside a slot (0 = one copy per slot)
  unsigned long long* dbg = nullptr;                // optional timeline of CTA 0 (globaltimer ns), scripts/bench_llm_kernels.py
  // blockIdx.y batches independent problems (the MTP heads): element strides
  size_t sW = 0, sBias = 0, sNorm = 0, sX = 0, sOut = 0, sResid = 0;
  LlmQkvEpi qkv;
};

// Each warp produces output features (2p, 2p+1) for all R rows: a RoPE pair (GEMV_QKV), a (gate, up)
// pair (GEMV_SWIGLU) or two plain features.  The grid is persistent (a few CTAs per SM, warps stride
// over the pairs) so the activation staging is paid once per CTA.  Launched with programmatic dependent
// launch: the first weight rows are requested *before* griddepcontrol.wait, i.e. while the producer of
// x is still draining, so the HBM latency of the weight stream overlaps the previous kernel's tail.
__device__ __forceinline__ void gemv_fma8(float& acc, const uint4 u, const float* xs) {
  const float4 xa = *reinterpret_cast<const float4*>(xs);
  const float4 xb = *reinterpret_cast<const float4*>(xs + 4);
  acc = fmaf(bf_lo(u.x), xa.x, acc); acc = fmaf(bf_hi(u.x), xa.y, acc);
  acc = fmaf(bf_lo(u.y), xa.z, acc); acc = fmaf(bf_hi(u.y), xa.w, acc);
  acc = fmaf(bf_lo(u.z), xb.x, acc); acc = fmaf(bf_hi(u.z), xb.y, acc);
  acc = fmaf(bf_lo(u.w), xb.z, acc); acc = fmaf(bf_hi(u.w), xb.w, acc);
}

// activations stored as two planes per row (elements k%8 < 4 | k%8 >= 4) so that the lanes' float4 reads are bank-conflict free
__device__ __forceinline__ void gemv_fma8p(float& acc, const uint4 u, const float* xa_p, const float* xb_p) {
  const float4 xa = *reinterpret_cast<const float4*>(xa_p);
  const float4 xb = *reinterpret_cast<const float4*>(xb_p);
  acc = fmaf(bf_lo(u.x), xa.x, acc); acc = fmaf(bf_hi(u.x), xa.y, acc);
  acc = fmaf(bf_lo(u.y), xa.z, acc); acc = fmaf(bf_hi(u.y), xa.w, acc);
  acc = fmaf(bf_lo(u.z), xb.x, acc); acc = fmaf(bf_hi(u.z), xb.y, acc);
  acc = fmaf(bf_lo(u.w), xb.z, acc); acc = fmaf(bf_hi(u.w), xb.w, acc);
}

template <int R>
__global__ void __launch_bounds__(256) llm_gemv_kernel(GemvArgs a) {
  extern __shared__ float sx[];                     // [R][kc] activations, then [kc] norm weights
  __shared__ __align__(8) uint64_t s_bar;
  asm volatile("griddepcontrol.launch_dependents;");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const int by = blockIdx.y;
  const float* x = a.x + by * a.sX;
  const __nv_bfloat16* W = a.W + by * a.sW;
  const int K = a.K, kc = a.kc;
  const int npairs = (a.N + 1) >> 1;
  const int first = blockIdx.x * nwarp + warp, stride = gridDim.x * nwarp;
  const bool multi = K > kc;                        // host guarantees <= 1 pair per warp in that case
  const float* nw = a.norm_w ? a.norm_w + by * a.sNorm : nullptr;
  float* snw = sx + R * kc;

  // ---- independent of the previous kernel: request the first weight rows of this warp, stage the norm weights
  uint4 p0[4], p1[4];
  {
    const int n0 = 2 * first;
    const __nv_bfloat16* w0 = W + (size_t)n0 * K;
    const __nv_bfloat16* w1 = w0 + ((n0 + 1 < a.N) ? K : 0);
    const int kn = K < kc ? K : kc;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int k = lane * 8 + i * 256;
      if (first < npairs && k < kn) { p0[i] = ldg_stream(w0 + k); p1[i] = ldg_stream(w1 + k); }
      else { p0[i] = make_uint4(0, 0, 0, 0); p1[i] = make_uint4(0, 0, 0, 0); }
    }
  }
  if (nw) for (int k = tid * 4; k < K; k += blockDim.x * 4) *reinterpret_cast<float4*>(&snw[k]) = __ldg(reinterpret_cast<const float4*>(nw + k));
  if (tid == 0) { tc::mbar_init(&s_bar, 1); tc::fence_barrier_init(); }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  __syncthreads();

  const float* bias = a.bias ? a.bias + by * a.sBias : nullptr;
  float* out = a.out + by * a.sOut;
  const float* resid = a.resid ? a.resid + by * a.sResid : nullptr;
  float acc0[R], acc1[R];
  uint32_t bar_phase = 0;

  for (int kb = 0; kb < K; kb += kc) {
    const int kn = (K - kb) < kc ? (K - kb) : kc;
    if (kb) __syncthreads();
    // activation rows -> shared memory through the bulk-copy engine: one L2 round trip for the whole chunk
    if (tid == 0) {
      tc::mbar_expect_tx(&s_bar, (uint32_t)(a.rows * kn * 4));
      for (int r = 0; r < a.rows; r++) bulk_g2s(&sx[r * kc], x + (size_t)r * a.ldx + kb, (uint32_t)(kn * 4), &s_bar);
    }
    for (int i = tid; i < (R - a.rows) * kn; i += blockDim.x) sx[a.rows * kc + (i / kn) * kc + (i % kn)] = 0.f;   // padding rows
    tc::mbar_wait(&s_bar, bar_phase);
    bar_phase ^= 1;
    if (nw) {
      // fused RMSNorm (HF Qwen2RMSNorm: fp32, eps inside rsqrt; K == kn here): x <- nw * (x * rsqrt(mean(x^2)+eps)),
      // each warp normalises a slice of every row after all warps have computed the row scales redundantly
      float sc[R];
#pragma unroll
      for (int r = 0; r < R; r++) {
        float ss = 0.f;
        for (int k = lane * 4; k < kn; k += 128) {
          const float4 v = *reinterpret_cast<const float4*>(&sx[r * kc + k]);
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        sc[r] = rsqrtf(ss / (float)K + a.eps);
      }
      __syncthreads();                              // every warp has read the raw rows
      for (int k = tid * 4; k < kn; k += blockDim.x * 4) {
        const float4 w4 = *reinterpret_cast<const float4*>(&snw[k]);
#pragma unroll
        for (int r = 0; r < R; r++) {
          float4 v = *reinterpret_cast<float4*>(&sx[r * kc + k]);
          v.x = w4.x * (v.x * sc[r]); v.y = w4.y * (v.y * sc[r]); v.z = w4.z * (v.z * sc[r]); v.w = w4.w * (v.w * sc[r]);
          *reinterpret_cast<float4*>(&sx[r * kc + k]) = v;
        }
      }
    }
    __syncthreads();
    for (int pair = first; pair < npairs; pair += stride) {
      const int n0 = 2 * pair;
      const __nv_bfloat16* w0 = W + (size_t)n0 * K + kb;
      const __nv_bfloat16* w1 = w0 + ((n0 + 1 < a.N) ? K : 0);
      if (!multi || kb == 0) {
#pragma unroll
        for (int r = 0; r < R; r++) { acc0[r] = 0.f; acc1[r] = 0.f; }
      }
      // software pipeline across pairs: the next pair's first rows are requested before this pair is reduced
      uint4 q0[4], q1[4];
      const int npair = pair + stride;
      const bool have_next = !multi && npair < npairs;
      if (have_next) {
        const __nv_bfloat16* v0 = W + (size_t)(2 * npair) * K;
        const __nv_bfloat16* v1 = v0 + ((2 * npair + 1 < a.N) ? K : 0);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int k = lane * 8 + i * 256;
          if (k < kn) { q0[i] = ldg_stream(v0 + k); q1[i] = ldg_stream(v1 + k); }
        }
      }
      const bool pre = !multi || kb == 0;           // p0/p1 hold the first 1024 columns of this pair
      for (int k0 = lane * 8; k0 < kn; k0 += 1024) {
        uint4 u0[4], u1[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int k = k0 + i * 256;
          if (pre && k0 == lane * 8) { u0[i] = p0[i]; u1[i] = p1[i]; }
          else if (k < kn) { u0[i] = ldg_stream(w0 + k); u1[i] = ldg_stream(w1 + k); }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int k = k0 + i * 256;
          if (k < kn) {
#pragma unroll
            for (int r = 0; r < R; r++) {
              gemv_fma8(acc0[r], u0[i], &sx[r * kc + k]);
              gemv_fma8(acc1[r], u1[i], &sx[r * kc + k]);
            }
          }
        }
      }
      if (have_next) {
#pragma unroll
        for (int i = 0; i < 4; i++) { p0[i] = q0[i]; p1[i] = q1[i]; }
      }
      if (multi && kb + kc < K) continue;
#pragma unroll
      for (int r = 0; r < R; r++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          acc0[r] += __shfl_xor_sync(0xffffffffu, acc0[r], o);
          acc1[r] += __shfl_xor_sync(0xffffffffu, acc1[r], o);
        }
      }
      // lane r finishes row r (bias, RoPE / SwiGLU / residual, store)
      float y0 = 0.f, y1 = 0.f;
#pragma unroll
      for (int r = 0; r < R; r++) if (lane == r) { y0 = acc0[r]; y1 = acc1[r]; }
      if (lane < a.rows) {
        const int r = lane;
        y0 += bias ? bias[n0] : 0.f;
        y1 += (bias && n0 + 1 < a.N) ? bias[n0 + 1] : 0.f;
        if (a.mode == GEMV_QKV) {
          llm_qkv_store(a.qkv, a.row0 + r, n0, y0, y1);
        } else if (a.mode == GEMV_SWIGLU) {
          out[(size_t)r * a.ldo + pair] = (y0 / (1.0f + expf(-y0))) * y1;
        } else {
          if (resid) { y0 += resid[(size_t)r * a.ldr + n0]; if (n0 + 1 < a.N) y1 += resid[(size_t)r * a.ldr + n0 + 1]; }
          out[(size_t)r * a.ldo + n0] = y0;
          if (n0 + 1 < a.N) out[(size_t)r * a.ldo + n0 + 1] = y1;
        }
      }
    }
  }
}

// ---- TMA-fed variant (K <= GT_KMAX): the weight stream is decoupled from the warps.  A CTA owns a contiguous
// range of output pairs, i.e. one contiguous byte range of W; a producer lane streams that range through a ring of
// shared-memory slots with 1-D bulk copies (cp.async.bulk) and starts doing so *before* griddepcontrol.wait, so
// bytes keep arriving while the previous kernel drains and while the activations are staged and normalised.  Eight
// consumer warps take the pairs of a landed slot round-robin and read their weights from shared memory.
constexpr int GT_SLOT = 40 * 1024;     // largest slot: 2 pairs at K=4864, 11 pairs at K=896
constexpr int GT_NS = 3;               // most slots; the launcher sizes the ring to the CTA's share so two CTAs co-reside
constexpr int GT_CONS = 16;            // consumer warps (8 warps + CTA co-residency measured slower: the pair loop, not the stream, is the critical path)
constexpr int GT_KMAX = 5120;          // a pair (2 rows x K bf16) must fit a slot

template <int R>
__global__ void __launch_bounds__((GT_CONS + 1) * 32, 1) llm_gemv_tma_kernel(GemvArgs a) {
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ __align__(8) uint64_t full_bar[GT_NS], empty_bar[GT_NS], x_bar;
  asm volatile("griddepcontrol.launch_dependents;");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int by = blockIdx.y;
  const int K = a.K;
  const float* x = a.x + by * a.sX;
  const __nv_bfloat16* W = a.W + by * a.sW;
  const float* nw = a.norm_w ? a.norm_w + by * a.sNorm : nullptr;
  const bool dbg = a.dbg && blockIdx.x == 0 && by == 0 && tid == 0;
  if (dbg) a.dbg[0] = gtime();
  uint8_t* ring = smraw;
  const int SLOT = a.slot_bytes, NS = a.n_slots;
  float* sx = reinterpret_cast<float*>(smraw + NS * SLOT);            // [R][2 planes][K/2]: conflict-free layout read by the pair loop
  float* snw = sx + R * K;                                            // [K]
  float* sraw = snw + K;                                              // [R][K]: rows as they land from global memory
  const int KH = K >> 1;
  const int npairs = (a.N + 1) >> 1;
  const int ppc = (npairs + gridDim.x - 1) / gridDim.x;               // pairs per CTA (contiguous)
  const int p_begin = blockIdx.x * ppc, p_end = min(npairs, p_begin + ppc);
  const int pps = SLOT / (4 * K);                                     // pairs per slot
  const int nslots = p_end > p_begin ? (p_end - p_begin + pps - 1) / pps : 0;

  if (tid == 0) {
    for (int i = 0; i < GT_NS; i++) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], GT_CONS); }
    tc::mbar_init(&x_bar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();

  if (warp == GT_CONS) {
    // ---------------- producer: weights only, never waits for the previous kernel
    if (lane == 0) {
      for (int i = 0; i < nslots; i++) {
        const int sl = i % NS;
        if (i >= NS) tc::mbar_wait(&empty_bar[sl], ((i / NS) - 1) & 1);
        const int p0 = p_begin + i * pps;
        const int row0 = 2 * p0, row1 = min(a.N, 2 * min(p_end, p0 + pps));
        const uint32_t bytes = (uint32_t)(row1 - row0) * (uint32_t)K * 2u;
        tc::mbar_expect_tx(&full_bar[sl], bytes);
        // several bulk copies per slot: one copy is executed by the TMA unit with limited request depth (a single 39 KB
        // copy streams at ~23 GB/s per SM), independent copies overlap
        const uint32_t piece = a.piece_bytes ? (uint32_t)a.piece_bytes : bytes;
        for (uint32_t off = 0; off < bytes; off += piece)
          bulk_g2s(ring + sl * SLOT + off, reinterpret_cast<const uint8_t*>(W + (size_t)row0 * K) + off, min(piece, bytes - off), &full_bar[sl]);
      }
    }
    return;
  }
  // ---------------- consumers
  if (nw) for (int k = tid * 4; k < K; k += GT_CONS * 128) *reinterpret_cast<float4*>(&snw[k]) = __ldg(reinterpret_cast<const float4*>(nw + k));
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (dbg) a.dbg[1] = gtime();
  if (tid == 0) {
    tc::mbar_expect_tx(&x_bar, (uint32_t)(a.rows * K * 4));
    for (int r = 0; r < a.rows; r++) bulk_g2s(&sraw[r * K], x + (size_t)r * a.ldx, (uint32_t)(K * 4), &x_bar);
  }
  for (int i = tid; i < (R - a.rows) * K; i += GT_CONS * 32) sx[a.rows * K + i] = 0.f;      // padding rows
  tc::mbar_wait(&x_bar, 0);
  if (dbg) a.dbg[2] = gtime();
  {
    float sc[R];
    if (nw) {                      // fused RMSNorm (HF Qwen2RMSNorm: fp32, eps inside rsqrt); every warp computes the row scales
#pragma unroll
      for (int r = 0; r < R; r++) {
        float ss = 0.f;
        if (r < a.rows)
          for (int k = lane * 4; k < K; k += 128) {
            const float4 v = *reinterpret_cast<const float4*>(&sraw[r * K + k]);
            ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
          }
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        sc[r] = rsqrtf(ss / (float)K + a.eps);
      }
    }
    // normalise (optional) and scatter into the two-plane layout
    for (int q = tid; q < (K >> 2); q += GT_CONS * 32) {
      const int dst = (q & 1) * KH + (q >> 1) * 4;
      float4 w4 = make_float4(1.f, 1.f, 1.f, 1.f);
      if (nw) w4 = *reinterpret_cast<const float4*>(&snw[q * 4]);
#pragma unroll
      for (int r = 0; r < R; r++) {
        if (r < a.rows) {
          float4 v = *reinterpret_cast<const float4*>(&sraw[r * K + q * 4]);
          if (nw) { v.x = w4.x * (v.x * sc[r]); v.y = w4.y * (v.y * sc[r]); v.z = w4.z * (v.z * sc[r]); v.w = w4.w * (v.w * sc[r]); }
          *reinterpret_cast<float4*>(&sx[r * K + dst]) = v;
        }
      }
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(GT_CONS * 32));
  if (dbg) a.dbg[3] = gtime();
  const float* bias = a.bias ? a.bias + by * a.sBias : nullptr;
  float* out = a.out + by * a.sOut;
  const float* resid = a.resid ? a.resid + by * a.sResid : nullptr;

  for (int i = 0; i < nslots; i++) {
    const int sl = i % NS;
    tc::mbar_wait(&full_bar[sl], (i / NS) & 1);
    if (dbg && i < 3) a.dbg[4 + i] = gtime();
    const int p0 = p_begin + i * pps;
    const int np = min(pps, p_end - p0);
    const __nv_bfloat16* ws = reinterpret_cast<const __nv_bfloat16*>(ring + sl * SLOT);
    // rotate the warp -> pair assignment from slot to slot so that warps without a pair in this slot move straight on
    // to the next one (all slots of a small matrix have already landed)
    for (int pi = (warp + i * 5) % GT_CONS; pi < np; pi += GT_CONS) {
      const int pair = p0 + pi, n0 = 2 * pair;
      const bool two = n0 + 1 < a.N;
      const __nv_bfloat16* w0 = ws + (size_t)(2 * pi) * K;
      const __nv_bfloat16* w1 = two ? w0 + K : w0;
      float acc0[R], acc1[R], bcc0[R], bcc1[R];          // two independent chains per output (even / odd k-blocks)
#pragma unroll
      for (int r = 0; r < R; r++) { acc0[r] = 0.f; acc1[r] = 0.f; bcc0[r] = 0.f; bcc1[r] = 0.f; }
      int k = lane * 8;
      for (; k + 256 < K; k += 512) {
        const uint4 u0 = *reinterpret_cast<const uint4*>(w0 + k);
        const uint4 u1 = *reinterpret_cast<const uint4*>(w1 + k);
        const uint4 v0 = *reinterpret_cast<const uint4*>(w0 + k + 256);
        const uint4 v1 = *reinterpret_cast<const uint4*>(w1 + k + 256);
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float* xa = &sx[r * K + (k >> 1)];
          gemv_fma8p(acc0[r], u0, xa, xa + KH);
          gemv_fma8p(acc1[r], u1, xa, xa + KH);
          gemv_fma8p(bcc0[r], v0, xa + 128, xa + KH + 128);
          gemv_fma8p(bcc1[r], v1, xa + 128, xa + KH + 128);
        }
      }
      if (k < K) {
        const uint4 u0 = *reinterpret_cast<const uint4*>(w0 + k);
        const uint4 u1 = *reinterpret_cast<const uint4*>(w1 + k);
#pragma unroll
        for (int r = 0; r < R; r++) {
          const float* xa = &sx[r * K + (k >> 1)];
          gemv_fma8p(acc0[r], u0, xa, xa + KH);
          gemv_fma8p(acc1[r], u1, xa, xa + KH);
        }
      }
#pragma unroll
      for (int r = 0; r < R; r++) {
        acc0[r] += bcc0[r]; acc1[r] += bcc1[r];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
          acc0[r] += __shfl_xor_sync(0xffffffffu, acc0[r], o);
          acc1[r] += __shfl_xor_sync(0xffffffffu, acc1[r], o);
        }
      }
      float y0 = 0.f, y1 = 0.f;
#pragma unroll
      for (int r = 0; r < R; r++) if (lane == r) { y0 = acc0[r]; y1 = acc1[r]; }
      if (lane < a.rows) {
        const int r = lane;
        y0 += bias ? bias[n0] : 0.f;
        y1 += (bias && two) ? bias[n0 + 1] : 0.f;
        if (a.mode == GEMV_QKV) {
          llm_qkv_store(a.qkv, a.row0 + r, n0, y0, y1);
        } else if (a.mode == GEMV_SWIGLU) {
          out[(size_t)r * a.ldo + pair] = (y0 / (1.0f + expf(-y0))) * y1;
        } else {
          if (resid) { y0 += resid[(size_t)r * a.ldr + n0]; if (two) y1 += resid[(size_t)r * a.ldr + n0 + 1]; }
          out[(size_t)r * a.ldo + n0] = y0;
          if (two) out[(size_t)r * a.ldo + n0 + 1] = y1;
        }
      }
    }
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&empty_bar[sl]);
  }
  if (dbg) a.dbg[7] = gtime();
}

// ------------------------------------------------------------------ attention over the KV cache
struct AttnDecArgs {
  const float* q = nullptr; int ldq = 0;            // [rows][q_dim] fp32 (RoPE applied)
  const void* kc = nullptr;                         // layer base [seq][kv_head][max_ctx][64] (bf16, or fp32: KV32)
  const void* vc = nullptr;
  size_t seq_stride = 0;
  int max_ctx = 0, kv_heads = 0, group = 0;         // group = q heads per kv head (blockDim = 32*group)
  const SeqState* seqs = nullptr; int rows_per_seq = 1;   // decode; nullptr: prefill
  int seq0 = 0, pos0 = 0, n_rows = 0;               // prefill
  int splits = 1;
  float* part = nullptr;                            // [row][q_head][split][68] partial (m, l, -, -, o[64])
  int* counters = nullptr;                          // [row][kv_head]
  float* out = nullptr; __nv_bfloat16* out16 = nullptr; int ldo = 0;   // [rows][q_dim]
  float scale = 0.125f;
};

// grid (splits, rows*kv_heads): one block = one row, one kv head, one slice of the visible keys; warp g of
// the block is q head kv*group+g.  Lanes own keys (a key's 128 B row is read by one lane from padded smem),
// so the only cross-lane traffic is the final merge.
// CL: the `splits` CTAs of one (row, kv head) form a thread-block cluster and merge their partial softmax results
// through distributed shared memory (one cluster barrier) instead of global partials + fence + atomic counter.
template <bool KV32, bool CL>
__global__ void __launch_bounds__(256) llm_attn_kernel(AttnDecArgs a) {
  using KT = typename std::conditional<KV32, float, __nv_bfloat16>::type;
  constexpr int LD = KV32 ? ATT_LD32 : ATT_LD;      // smem row stride in elements
  constexpr int SEGS = KV32 ? 16 : 8;               // 16 B segments per 64-element row
  constexpr int CHUNK = KV32 ? ATT_KEYS / 2 : ATT_KEYS;
  __shared__ __align__(16) KT sk[CHUNK * LD];
  __shared__ __align__(16) KT sv[CHUNK * LD];
  __shared__ int s_last;
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, g = tid >> 5;
  const int split = blockIdx.x;
  const int row = blockIdx.y / a.kv_heads, kvh = blockIdx.y - row * a.kv_heads;
  int seq, pos;
  if (a.seqs) {
    seq = row / a.rows_per_seq;
    const int r = row - seq * a.rows_per_seq;
    const SeqState& s = a.seqs[seq];
    if (s.done || r >= s.n_new) return;
    pos = s.ctx + r;
  } else {
    if (row >= a.n_rows) return;
    seq = a.seq0;
    pos = a.pos0 + row;
  }
  const int n_keys = min(pos + 1, a.max_ctx);
  const int per = (n_keys + a.splits - 1) / a.splits;
  const int k_begin = split * per, k_end = min(n_keys, k_begin + per);
  const int qh = kvh * a.group + g;
  // 4 lanes share a key (16 of the 64 dims each), 8 keys per warp iteration: the dot product needs two shuffles and
  // the final merge of the 8 key slots three
  const int sub = lane & 3, kslot = lane >> 2;
  float qv[16];
  {
    const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)row * a.ldq + qh * 64 + sub * 16);
#pragma unroll
    for (int i = 0; i < 4; i++) { const float4 t = qp[i]; qv[4 * i] = t.x * a.scale; qv[4 * i + 1] = t.y * a.scale; qv[4 * i + 2] = t.z * a.scale; qv[4 * i + 3] = t.w * a.scale; }
  }
  float m = -INFINITY, l = 0.f, o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) o[i] = 0.f;
  const KT* kb = reinterpret_cast<const KT*>(a.kc) + (size_t)seq * a.seq_stride + (size_t)kvh * a.max_ctx * 64;
  const KT* vb = reinterpret_cast<const KT*>(a.vc) + (size_t)seq * a.seq_stride + (size_t)kvh * a.max_ctx * 64;
  constexpr int EPS = 16 / (int)sizeof(KT);         // elements per 16 B segment
  for (int c0 = k_begin; c0 < k_end; c0 += CHUNK) {
    const int nk = min(CHUNK, k_end - c0);
    __syncthreads();
    for (int i = tid; i < nk * SEGS; i += blockDim.x) {
      const int key = i / SEGS, seg = i - key * SEGS;
      *reinterpret_cast<uint4*>(&sk[key * LD + seg * EPS]) = *reinterpret_cast<const uint4*>(kb + (size_t)(c0 + key) * 64 + seg * EPS);
      *reinterpret_cast<uint4*>(&sv[key * LD + seg * EPS]) = *reinterpret_cast<const uint4*>(vb + (size_t)(c0 + key) * 64 + seg * EPS);
    }
    __syncthreads();
    for (int k0 = 0; k0 < nk; k0 += 8) {
      const int key = k0 + kslot;
      const bool live = key < nk;
      float kf[16], vf[16];
      if (live) {
        if constexpr (KV32) {
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 u = *reinterpret_cast<const float4*>(&sk[key * LD + sub * 16 + 4 * i]);
            const float4 w = *reinterpret_cast<const float4*>(&sv[key * LD + sub * 16 + 4 * i]);
            kf[4 * i] = u.x; kf[4 * i + 1] = u.y; kf[4 * i + 2] = u.z; kf[4 * i + 3] = u.w;
            vf[4 * i] = w.x; vf[4 * i + 1] = w.y; vf[4 * i + 2] = w.z; vf[4 * i + 3] = w.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const uint4 u = *reinterpret_cast<const uint4*>(&sk[key * LD + sub * 16 + 8 * i]);
            const uint4 w = *reinterpret_cast<const uint4*>(&sv[key * LD + sub * 16 + 8 * i]);
            kf[8 * i] = bf_lo(u.x); kf[8 * i + 1] = bf_hi(u.x); kf[8 * i + 2] = bf_lo(u.y); kf[8 * i + 3] = bf_hi(u.y);
            kf[8 * i + 4] = bf_lo(u.z); kf[8 * i + 5] = bf_hi(u.z); kf[8 * i + 6] = bf_lo(u.w); kf[8 * i + 7] = bf_hi(u.w);
            vf[8 * i] = bf_lo(w.x); vf[8 * i + 1] = bf_hi(w.x); vf[8 * i + 2] = bf_lo(w.y); vf[8 * i + 3] = bf_hi(w.y);
            vf[8 * i + 4] = bf_lo(w.z); vf[8 * i + 5] = bf_hi(w.z); vf[8 * i + 6] = bf_lo(w.w); vf[8 * i + 7] = bf_hi(w.w);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; i++) { kf[i] = 0.f; vf[i] = 0.f; }
      }
      float sc = 0.f;
#pragma unroll
      for (int i = 0; i < 16; i++) sc = fmaf(qv[i], kf[i], sc);
      sc += __shfl_xor_sync(0xffffffffu, sc, 1);
      sc += __shfl_xor_sync(0xffffffffu, sc, 2);
      if (live) {
        if (sc > m) {
          const float al = expf(m - sc);
          l *= al;
#pragma unroll
          for (int i = 0; i < 16; i++) o[i] *= al;
          m = sc;
        }
        const float pw = expf(sc - m);
        l += pw;
#pragma unroll
        for (int i = 0; i < 16; i++) o[i] = fmaf(pw, vf[i], o[i]);
      }
    }
  }
  // merge the 8 key slots (lanes with equal `sub`)
  float M = m;
  for (int sft = 4; sft < 32; sft <<= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, sft));
  const float wgt = (m == -INFINITY) ? 0.f : expf(m - M);
  float L = l * wgt;
  for (int sft = 4; sft < 32; sft <<= 1) L += __shfl_xor_sync(0xffffffffu, L, sft);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    float v = o[i] * wgt;
    for (int sft = 4; sft < 32; sft <<= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
    o[i] = v;
  }
  const int q_heads = a.kv_heads * a.group;
  // lanes 0..3 (kslot 0) now hold dims [16*sub, 16*sub+16)
  if (a.splits == 1) {
    if (kslot == 0) {
      const float inv = 1.0f / L;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const float y = o[i] * inv;
        if (a.out) a.out[(size_t)row * a.ldo + qh * 64 + sub * 16 + i] = y;
        if (a.out16) store_split(a.out16 + (size_t)row * 2 * a.ldo, a.ldo, qh * 64 + sub * 16 + i, y);
      }
    }
    return;
  }
  if constexpr (CL) {
    __shared__ __align__(16) float s_part[8][68];
    if (kslot == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++) *reinterpret_cast<float4*>(&s_part[g][4 + sub * 16 + 4 * i]) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    if (lane == 0) { s_part[g][0] = M; s_part[g][1] = L; }
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    if ((uint32_t)(g % a.splits) == rank) {            // this CTA merges head g: read its partial from every CTA of the cluster
      const uint32_t local = tc::smem_u32(&s_part[g][0]);
      // issue every remote load before any dependent math: 16 x 4 independent DSMEM reads in flight instead of 16 round trips
      float mv[16], lv[16], x0[16], x1[16];
#pragma unroll
      for (int c = 0; c < 16; c++) {
        if (c < a.splits) {
          uint32_t ra;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local), "r"(c));
          asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(mv[c]) : "r"(ra));
          asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(lv[c]) : "r"(ra + 4));
          asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(x0[c]) : "r"(ra + 16 + 4 * lane));
          asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(x1[c]) : "r"(ra + 16 + 128 + 4 * lane));
        } else { mv[c] = -INFINITY; lv[c] = 0.f; x0[c] = 0.f; x1[c] = 0.f; }
      }
      float Mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 16; c++) Mx = fmaxf(Mx, mv[c]);
      float Ls = 0.f, a_lo = 0.f, a_hi = 0.f;
#pragma unroll
      for (int c = 0; c < 16; c++) {
        const float w = (mv[c] == -INFINITY) ? 0.f : expf(mv[c] - Mx);
        Ls += lv[c] * w; a_lo += x0[c] * w; a_hi += x1[c] * w;
      }
      const float inv = 1.0f / Ls;
      if (a.out) { a.out[(size_t)row * a.ldo + qh * 64 + lane] = a_lo * inv; a.out[(size_t)row * a.ldo + qh * 64 + 32 + lane] = a_hi * inv; }
      if (a.out16) {
        store_split(a.out16 + (size_t)row * 2 * a.ldo, a.ldo, qh * 64 + lane, a_lo * inv);
        store_split(a.out16 + (size_t)row * 2 * a.ldo, a.ldo, qh * 64 + 32 + lane, a_hi * inv);
      }
    }
    // nobody may exit while a peer can still read its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    return;
  }
  float* pp = a.part + (((size_t)row * q_heads + qh) * a.splits + split) * 68;
  if (kslot == 0) {
#pragma unroll
    for (int i = 0; i < 4; i++) *reinterpret_cast<float4*>(&pp[4 + sub * 16 + 4 * i]) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
  }
  if (lane == 0) { pp[0] = M; pp[1] = L; }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&a.counters[row * a.kv_heads + kvh], 1) == a.splits - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last block for this (row, kv head): merge the split partials of its heads
  const float* pb = a.part + ((size_t)row * q_heads + qh) * a.splits * 68;
  float Mx = -INFINITY;
  for (int s = 0; s < a.splits; s++) Mx = fmaxf(Mx, __ldcg(pb + s * 68));
  float Ls = 0.f, a_lo = 0.f, a_hi = 0.f;
  for (int s = 0; s < a.splits; s++) {
    const float ms = __ldcg(pb + s * 68);
    const float w = (ms == -INFINITY) ? 0.f : expf(ms - Mx);
    Ls += __ldcg(pb + s * 68 + 1) * w;
    a_lo += __ldcg(pb + s * 68 + 4 + lane) * w;
    a_hi += __ldcg(pb + s * 68 + 36 + lane) * w;
  }
  const float inv = 1.0f / Ls;
  if (a.out) { a.out[(size_t)row * a.ldo + qh * 64 + lane] = a_lo * inv; a.out[(size_t)row * a.ldo + qh * 64 + 32 + lane] = a_hi * inv; }
  if (a.out16) {
    store_split(a.out16 + (size_t)row * 2 * a.ldo, a.ldo, qh * 64 + lane, a_lo * inv);
    store_split(a.out16 + (size_t)row * 2 * a.ldo, a.ldo, qh * 64 + 32 + lane, a_hi * inv);
  }
  if (tid == 0) a.counters[row * a.kv_heads + kvh] = 0;
}


// ---- decode attention, one block per (sequence, kv head, key slice): the head_k rows a sequence decodes in one step
// (positions ctx .. ctx+n_new-1) attend over the SAME cache, so the K/V chunk staged in shared memory is shared by all of them —
// llm_attn_kernel gives every row its own block and re-reads the cache head_k times, which at 32 sequences x 4 rows x 2600 keys
// (BASELINE configs[2]) made the attention the largest part of the decode step.  Warp (g, p): q head kv*group+g, row pair p
// (rows 2p, 2p+1); 4 lanes share a key (16 dims each), 8 keys per warp iteration; the chunk ring is double-buffered with
// cp.async.  Partials go to `part` and the last-arriving block of a (sequence, kv head) merges them (fixed order: deterministic).
template <bool KV32>
__global__ void __launch_bounds__(512) llm_attn_seq_kernel(AttnDecArgs a) {
  using KT = typename std::conditional<KV32, float, __nv_bfloat16>::type;
  constexpr int LD = KV32 ? ATT_LD32 : ATT_LD;
  constexpr int SEGS = KV32 ? 16 : 8;
  constexpr int CHUNK = KV32 ? ATT_KEYS / 2 : ATT_KEYS;
  constexpr int EPS = 16 / (int)sizeof(KT);
  extern __shared__ __align__(16) uint8_t att_smem[];
  KT* sk = reinterpret_cast<KT*>(att_smem);                    // [2][CHUNK * LD]
  KT* sv = sk + 2 * CHUNK * LD;
  __shared__ int s_last;
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = warp % a.group, pr = warp / a.group;
  const int split = blockIdx.x;
  const int seq = blockIdx.y / a.kv_heads, kvh = blockIdx.y - seq * a.kv_heads;
  const SeqState& s = a.seqs[seq];
  if (s.done || s.n_new <= 0) return;
  const int n_new = min(s.n_new, a.rows_per_seq);
  const int n_keys = min(s.ctx + n_new, a.max_ctx);
  const int per = (n_keys + a.splits - 1) / a.splits;
  const int k_begin = split * per, k_end = min(n_keys, k_begin + per);
  const int qh = kvh * a.group + g;
  const int sub = lane & 3, kslot = lane >> 2;
  const int r0 = 2 * pr;
  const bool live0 = r0 < n_new, live1 = r0 + 1 < n_new;
  const int pos0 = s.ctx + r0, pos1 = pos0 + 1;               // last key row r sees
  float q0[16], q1[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { q0[i] = 0.f; q1[i] = 0.f; }
  if (live0) {
    const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)(seq * a.rows_per_seq + r0) * a.ldq + qh * 64 + sub * 16);
#pragma unroll
    for (int i = 0; i < 4; i++) { const float4 t = qp[i]; q0[4 * i] = t.x * a.scale; q0[4 * i + 1] = t.y * a.scale; q0[4 * i + 2] = t.z * a.scale; q0[4 * i + 3] = t.w * a.scale; }
  }
  if (live1) {
    const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)(seq * a.rows_per_seq + r0 + 1) * a.ldq + qh * 64 + sub * 16);
#pragma unroll
    for (int i = 0; i < 4; i++) { const float4 t = qp[i]; q1[4 * i] = t.x * a.scale; q1[4 * i + 1] = t.y * a.scale; q1[4 * i + 2] = t.z * a.scale; q1[4 * i + 3] = t.w * a.scale; }
  }
  float m0 = -INFINITY, l0 = 0.f, m1 = -INFINITY, l1 = 0.f, o0[16], o1[16];
#pragma unroll
  for (int i = 0; i < 16; i++) { o0[i] = 0.f; o1[i] = 0.f; }
  const KT* kb = reinterpret_cast<const KT*>(a.kc) + (size_t)seq * a.seq_stride + (size_t)kvh * a.max_ctx * 64;
  const KT* vb = reinterpret_cast<const KT*>(a.vc) + (size_t)seq * a.seq_stride + (size_t)kvh * a.max_ctx * 64;
  auto stage = [&](int c0, int buf) {
    const int nk = min(CHUNK, k_end - c0);
    KT* dk = sk + buf * CHUNK * LD;
    KT* dv = sv + buf * CHUNK * LD;
    for (int i = tid; i < nk * SEGS; i += blockDim.x) {
      const int key = i / SEGS, seg = i - key * SEGS;
      const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dk + key * LD + seg * EPS);
      const uint32_t d1 = (uint32_t)__cvta_generic_to_shared(dv + key * LD + seg * EPS);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(kb + (size_t)(c0 + key) * 64 + seg * EPS) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(vb + (size_t)(c0 + key) * 64 + seg * EPS) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int buf = 0;
  if (k_begin < k_end) stage(k_begin, 0);
  for (int c0 = k_begin; c0 < k_end; c0 += CHUNK, buf ^= 1) {
    const int nk = min(CHUNK, k_end - c0);
    if (c0 + CHUNK < k_end) { stage(c0 + CHUNK, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const KT* ck = sk + buf * CHUNK * LD;
    const KT* cv = sv + buf * CHUNK * LD;
    if (live0) {
      for (int k0 = 0; k0 < nk; k0 += 8) {
        const int key = k0 + kslot;
        const bool live = key < nk;
        const int kabs = c0 + key;
        float kf[16], vf[16];
        if (live) {
          if constexpr (KV32) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float4 u = *reinterpret_cast<const float4*>(&ck[key * LD + sub * 16 + 4 * i]);
              const float4 w = *reinterpret_cast<const float4*>(&cv[key * LD + sub * 16 + 4 * i]);
              kf[4 * i] = u.x; kf[4 * i + 1] = u.y; kf[4 * i + 2] = u.z; kf[4 * i + 3] = u.w;
              vf[4 * i] = w.x; vf[4 * i + 1] = w.y; vf[4 * i + 2] = w.z; vf[4 * i + 3] = w.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 2; i++) {
              const uint4 u = *reinterpret_cast<const uint4*>(&ck[key * LD + sub * 16 + 8 * i]);
              const uint4 w = *reinterpret_cast<const uint4*>(&cv[key * LD + sub * 16 + 8 * i]);
              kf[8 * i] = bf_lo(u.x); kf[8 * i + 1] = bf_hi(u.x); kf[8 * i + 2] = bf_lo(u.y); kf[8 * i + 3] = bf_hi(u.y);
              kf[8 * i + 4] = bf_lo(u.z); kf[8 * i + 5] = bf_hi(u.z); kf[8 * i + 6] = bf_lo(u.w); kf[8 * i + 7] = bf_hi(u.w);
              vf[8 * i] = bf_lo(w.x); vf[8 * i + 1] = bf_hi(w.x); vf[8 * i + 2] = bf_lo(w.y); vf[8 * i + 3] = bf_hi(w.y);
              vf[8 * i + 4] = bf_lo(w.z); vf[8 * i + 5] = bf_hi(w.z); vf[8 * i + 6] = bf_lo(w.w); vf[8 * i + 7] = bf_hi(w.w);
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; i++) { kf[i] = 0.f; vf[i] = 0.f; }
        }
        float sc0 = 0.f, sc1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++) { sc0 = fmaf(q0[i], kf[i], sc0); sc1 = fmaf(q1[i], kf[i], sc1); }
        sc0 += __shfl_xor_sync(0xffffffffu, sc0, 1); sc1 += __shfl_xor_sync(0xffffffffu, sc1, 1);
        sc0 += __shfl_xor_sync(0xffffffffu, sc0, 2); sc1 += __shfl_xor_sync(0xffffffffu, sc1, 2);
        if (live && kabs <= pos0) {
          if (sc0 > m0) {
            const float al = expf(m0 - sc0);
            l0 *= al;
#pragma unroll
            for (int i = 0; i < 16; i++) o0[i] *= al;
            m0 = sc0;
          }
          const float pw = expf(sc0 - m0);
          l0 += pw;
#pragma unroll
          for (int i = 0; i < 16; i++) o0[i] = fmaf(pw, vf[i], o0[i]);
        }
        if (live && live1 && kabs <= pos1) {
          if (sc1 > m1) {
            const float al = expf(m1 - sc1);
            l1 *= al;
#pragma unroll
            for (int i = 0; i < 16; i++) o1[i] *= al;
            m1 = sc1;
          }
          const float pw = expf(sc1 - m1);
          l1 += pw;
#pragma unroll
          for (int i = 0; i < 16; i++) o1[i] = fmaf(pw, vf[i], o1[i]);
        }
      }
    }
    __syncthreads();
  }
  const int q_heads = a.kv_heads * a.group;
  // merge the 8 key slots of each row (lanes with equal `sub`), publish the partial of this slice
#pragma unroll
  for (int rr = 0; rr < 2; rr++) {
    float* o = rr ? o1 : o0;
    const float m = rr ? m1 : m0, l = rr ? l1 : l0;
    const bool lv = rr ? live1 : live0;
    float M = m;
    for (int sft = 4; sft < 32; sft <<= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, sft));
    const float wgt = (m == -INFINITY) ? 0.f : expf(m - M);
    float Lr = l * wgt;
    for (int sft = 4; sft < 32; sft <<= 1) Lr += __shfl_xor_sync(0xffffffffu, Lr, sft);
#pragma unroll
    for (int i = 0; i < 16; i++) {
      float v = o[i] * wgt;
      for (int sft = 4; sft < 32; sft <<= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
      o[i] = v;
    }
    if (lv) {
      const int row = seq * a.rows_per_seq + r0 + rr;
      float* pp = a.part + (((size_t)row * q_heads + qh) * a.splits + split) * 68;
      if (kslot == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) *reinterpret_cast<float4*>(&pp[4 + sub * 16 + 4 * i]) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
      if (lane == 0) { pp[0] = M; pp[1] = Lr; }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&a.counters[seq * a.kv_heads + kvh], 1) == a.splits - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int rr = 0; rr < 2; rr++) {
    if (!(rr ? live1 : live0)) continue;
    const int row = seq * a.rows_per_seq + r0 + rr;
    const float* pb = a.part + ((size_t)row * q_heads + qh) * a.splits * 68;
    float Mx = -INFINITY;
    for (int sp = 0; sp < a.splits; sp++) Mx = fmaxf(Mx, __ldcg(pb + sp * 68));
    float Ls = 0.f, a_lo = 0.f, a_hi = 0.f;
    for (int sp = 0; sp < a.splits; sp++) {
      const float ms = __ldcg(pb + sp * 68);
      const float w = (ms == -INFINITY) ? 0.f : expf(ms - Mx);
      Ls += __ldcg(pb + sp * 68 + 1) * w;
      a_lo += __ldcg(pb + sp * 68 + 4 + lane) * w;
      a_hi += __ldcg(pb + sp * 68 + 36 + lane) * w;
    }
    const float inv = 1.0f / Ls;
    if (a.out) { a.out[(size_t)row * a.ldo + qh * 64 + lane] = a_lo * inv; a.out[(size_t)row * a.ldo + qh * 64 + 32 + lane] = a_hi * inv; }
    if (a.out16) {
      store_split(a.out16 + (size_t)row * 2 * a.ldo, a.ldo, qh * 64 + lane, a_lo * inv);
      store_split(a.out16 + (size_t)row * 2 * a.ldo, a.ldo, qh * 64 + 32 + lane, a_hi * inv);
    }
  }
  if (tid == 0) a.counters[seq * a.kv_heads + kvh] = 0;
}

// ---- batched decode attention on the tensor cores (mma.sync m16n8k16, bf16 operands, fp32 accumulate)
// One block per (sequence, kv head, key slice).  The `group` q heads of the kv head x the n_new <= 4 rows the sequence decodes
// this step form ONE 32-row query tile (row m = r * group + g), so every staged key is used by all of them from one fragment
// load: the CUDA-core kernels above spend ~6 instructions per (key, query row), this one < 1.
//   S (32 x 16 keys per warp and chunk) = Q K^T,  P = exp(S - m),  O (32 x 64) += P V        (flash-decoding, online softmax)
// Operands are split into bf16 (hi, lo) pairs and multiplied as hi*hi + lo*hi (+ hi*lo when the cache is fp32) — 16+ mantissa
// bits per operand — so the tokens of a batch stay those of the single-sequence fp32 path (tests/test_llm_gpu.py).
// 8 warps; warp w owns keys [16 w, 16 w + 16) of every 128-key chunk (cp.async double buffer); per-warp partials are merged
// through shared memory, slices through `part` + the last-arriver merge of llm_attn_kernel.
constexpr int AM_CH = 128;            // keys per chunk
constexpr int AM_LDQ = 72;            // bf16 elements per Q row (144 B: conflict-free ldmatrix)
constexpr int AM_LDK32 = 72, AM_LDV32 = 68;   // fp32 cache: K rows read as float2 per (key g, dims 2t), V as scalars per (key 2t, dim g)
constexpr int AM_LD16 = 72;           // bf16 cache

__device__ __forceinline__ void mma_bf16_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_bf16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - __low2float(h), y - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h); lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <bool KV32>
__global__ void __launch_bounds__(256, 1) llm_attn_mma_kernel(AttnDecArgs a) {
  using KT = typename std::conditional<KV32, float, __nv_bfloat16>::type;
  constexpr int LDK = KV32 ? AM_LDK32 : AM_LD16, LDV = KV32 ? AM_LDV32 : AM_LD16;
  constexpr int SEGS = KV32 ? 16 : 8, EPS = 16 / (int)sizeof(KT);
  extern __shared__ __align__(16) uint8_t am_smem[];
  __nv_bfloat16* qhi = reinterpret_cast<__nv_bfloat16*>(am_smem);          // [32][AM_LDQ]
  __nv_bfloat16* qlo = qhi + 32 * AM_LDQ;
  KT* sk = reinterpret_cast<KT*>(qlo + 32 * AM_LDQ);                       // [2][AM_CH * LDK]
  KT* sv = sk + 2 * AM_CH * LDK;                                           // [2][AM_CH * LDV]
  __shared__ int s_last;
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int split = blockIdx.x;
  const int seq = blockIdx.y / a.kv_heads, kvh = blockIdx.y - seq * a.kv_heads;
  const SeqState& s = a.seqs[seq];
  if (s.done || s.n_new <= 0) return;
  const int n_new = min(s.n_new, a.rows_per_seq);
  const int n_rows = n_new * a.group;                                      // live query rows of the tile (<= 32)
  const int n_keys = min(s.ctx + n_new, a.max_ctx);
  const int per = (n_keys + a.splits - 1) / a.splits;
  const int k_begin = split * per, k_end = min(n_keys, k_begin + per);
  // ---- Q tile: row m = r * group + gq  <-  q[(seq, r)][head kvh*group+gq] * scale, as bf16 hi / lo; zero the staging ring
  { // 32 rows x 16 float4: two per thread, both loads in flight before the first use (eight dependent scalar loads per thread
    // were the kernel's top stall, profiles/r2 decode attention)
    float4 qv[2];
#pragma unroll
    for (int it = 0; it < 2; it++) {
      const int i = tid + it * 256, m = i >> 4, d4 = (i & 15) * 4;
      qv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < n_rows) {
        const int r = m / a.group, gq = m - r * a.group;
        qv[it] = *reinterpret_cast<const float4*>(a.q + (size_t)(seq * a.rows_per_seq + r) * a.ldq + (kvh * a.group + gq) * 64 + d4);
      }
    }
#pragma unroll
    for (int it = 0; it < 2; it++) {
      const int i = tid + it * 256, m = i >> 4, d4 = (i & 15) * 4;
      const float v[4] = {qv[it].x * a.scale, qv[it].y * a.scale, qv[it].z * a.scale, qv[it].w * a.scale};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const __nv_bfloat16 h = __float2bfloat16(v[j]);
        qhi[m * AM_LDQ + d4 + j] = h;
        qlo[m * AM_LDQ + d4 + j] = __float2bfloat16(v[j] - __bfloat162float(h));
      }
    }
  }
  { uint4* z = reinterpret_cast<uint4*>(sk);
    const int nz = (int)((2 * AM_CH * (LDK + LDV) * sizeof(KT)) / 16);
    for (int i = tid; i < nz; i += 256) z[i] = make_uint4(0, 0, 0, 0); }
  __syncthreads();
  const KT* kb = reinterpret_cast<const KT*>(a.kc) + (size_t)seq * a.seq_stride + (size_t)kvh * a.max_ctx * 64;
  const KT* vb = reinterpret_cast<const KT*>(a.vc) + (size_t)seq * a.seq_stride + (size_t)kvh * a.max_ctx * 64;
  auto stage = [&](int c0, int buf) {
    const int nk = min(AM_CH, k_end - c0);
    KT* dk = sk + buf * AM_CH * LDK;
    KT* dv = sv + buf * AM_CH * LDV;
    for (int i = tid; i < nk * SEGS; i += 256) {
      const int key = i / SEGS, seg = i - key * SEGS;
      const uint32_t d0 = (uint32_t)__cvta_generic_to_shared(dk + key * LDK + seg * EPS);
      const uint32_t d1 = (uint32_t)__cvta_generic_to_shared(dv + key * LDV + seg * EPS);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(kb + (size_t)(c0 + key) * 64 + seg * EPS) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(vb + (size_t)(c0 + key) * 64 + seg * EPS) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // per-lane rows: index ri = 2 * mt + h  ->  tile row m = mt * 16 + g + 8 * h
  float O[2][8][4], mrow[4], lrow[4];
  int plim[4];
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nd = 0; nd < 8; nd++)
#pragma unroll
      for (int j = 0; j < 4; j++) O[mt][nd][j] = 0.f;
#pragma unroll
  for (int ri = 0; ri < 4; ri++) {
    const int m = (ri >> 1) * 16 + g + 8 * (ri & 1);
    mrow[ri] = -INFINITY; lrow[ri] = 0.f;
    plim[ri] = m < n_rows ? s.ctx + m / a.group : -1;                      // last visible key of the row (-1: padding row)
  }
  int buf = 0;
  if (k_begin < k_end) stage(k_begin, 0);
  for (int c0 = k_begin; c0 < k_end; c0 += AM_CH, buf ^= 1) {
    const int nk = min(AM_CH, k_end - c0);
    if (c0 + AM_CH < k_end) { stage(c0 + AM_CH, buf ^ 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int kw = warp * 16;                                              // this warp's keys inside the chunk
    if (kw < nk) {
      const KT* ck = sk + buf * AM_CH * LDK;
      const KT* cv = sv + buf * AM_CH * LDV;
      float S[2][2][4];
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
          for (int j = 0; j < 4; j++) S[mt][n][j] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ks++) {
        uint32_t ah[2][4], al[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          const int row = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = ks * 16 + (lane >> 4) * 8;
          const uint32_t ph = (uint32_t)__cvta_generic_to_shared(qhi + row * AM_LDQ + col);
          const uint32_t pl = (uint32_t)__cvta_generic_to_shared(qlo + row * AM_LDQ + col);
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(ah[mt][0]), "=r"(ah[mt][1]), "=r"(ah[mt][2]), "=r"(ah[mt][3]) : "r"(ph));
          asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(al[mt][0]), "=r"(al[mt][1]), "=r"(al[mt][2]), "=r"(al[mt][3]) : "r"(pl));
        }
#pragma unroll
        for (int n = 0; n < 2; n++) {
          const int key = kw + n * 8 + g;
          uint32_t bh0, bh1, bl0 = 0, bl1 = 0;
          if constexpr (KV32) {
            const float2 x0 = *reinterpret_cast<const float2*>(ck + key * LDK + ks * 16 + 2 * t);
            const float2 x1 = *reinterpret_cast<const float2*>(ck + key * LDK + ks * 16 + 8 + 2 * t);
            split_bf16x2(x0.x, x0.y, bh0, bl0);
            split_bf16x2(x1.x, x1.y, bh1, bl1);
          } else {
            bh0 = *reinterpret_cast<const uint32_t*>(ck + key * LDK + ks * 16 + 2 * t);
            bh1 = *reinterpret_cast<const uint32_t*>(ck + key * LDK + ks * 16 + 8 + 2 * t);
          }
#pragma unroll
          for (int mt = 0; mt < 2; mt++) {
            mma_bf16_16816(S[mt][n], ah[mt], bh0, bh1);
            mma_bf16_16816(S[mt][n], al[mt], bh0, bh1);
            if constexpr (KV32) mma_bf16_16816(S[mt][n], ah[mt], bl0, bl1);
          }
        }
      }
      // mask + online softmax.  S[mt][n][j]: row ri = 2*mt + (j>>1), key = c0 + kw + n*8 + 2t + (j&1)
      float tmax[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int ri = 2 * mt + (j >> 1);
            const int kabs = c0 + kw + n * 8 + 2 * t + (j & 1);
            const bool ok = kabs < k_end && kabs <= plim[ri];
            S[mt][n][j] = ok ? S[mt][n][j] : -INFINITY;
            tmax[ri] = fmaxf(tmax[ri], S[mt][n][j]);
          }
      float alpha[4];
#pragma unroll
      for (int ri = 0; ri < 4; ri++) {
        tmax[ri] = fmaxf(tmax[ri], __shfl_xor_sync(0xffffffffu, tmax[ri], 1));
        tmax[ri] = fmaxf(tmax[ri], __shfl_xor_sync(0xffffffffu, tmax[ri], 2));
        const float mn = fmaxf(mrow[ri], tmax[ri]);
        alpha[ri] = (mrow[ri] == -INFINITY) ? 0.f : expf(mrow[ri] - mn);     // mn == -inf only while the row has seen no key: alpha unused (O, l are 0)
        mrow[ri] = mn;
        lrow[ri] *= alpha[ri];
      }
      uint32_t ph[2][4], pl[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; mt++) {
        float p[2][4];
#pragma unroll
        for (int n = 0; n < 2; n++)
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int ri = 2 * mt + (j >> 1);
            const float e = (S[mt][n][j] == -INFINITY) ? 0.f : expf(S[mt][n][j] - mrow[ri]);
            p[n][j] = e;
            lrow[ri] += e;
          }
        split_bf16x2(p[0][0], p[0][1], ph[mt][0], pl[mt][0]);            // a0: row g,   keys 2t, 2t+1
        split_bf16x2(p[0][2], p[0][3], ph[mt][1], pl[mt][1]);            // a1: row g+8, keys 2t, 2t+1
        split_bf16x2(p[1][0], p[1][1], ph[mt][2], pl[mt][2]);            // a2: row g,   keys 2t+8, 2t+9
        split_bf16x2(p[1][2], p[1][3], ph[mt][3], pl[mt][3]);            // a3: row g+8, keys 2t+8, 2t+9
#pragma unroll
        for (int nd = 0; nd < 8; nd++) {
          O[mt][nd][0] *= alpha[2 * mt]; O[mt][nd][1] *= alpha[2 * mt];
          O[mt][nd][2] *= alpha[2 * mt + 1]; O[mt][nd][3] *= alpha[2 * mt + 1];
        }
      }
#pragma unroll
      for (int nd = 0; nd < 8; nd++) {
        // B fragment of V: b0 = (V[kw+2t][d], V[kw+2t+1][d]), b1 = (V[kw+2t+8][d], V[kw+2t+9][d]),  d = nd*8 + g
        const int d = nd * 8 + g;
        uint32_t vh0, vh1, vl0 = 0, vl1 = 0;
        if constexpr (KV32) {
          split_bf16x2(cv[(kw + 2 * t) * LDV + d], cv[(kw + 2 * t + 1) * LDV + d], vh0, vl0);
          split_bf16x2(cv[(kw + 2 * t + 8) * LDV + d], cv[(kw + 2 * t + 9) * LDV + d], vh1, vl1);
        } else {
          const uint16_t* v16 = reinterpret_cast<const uint16_t*>(cv);
          vh0 = (uint32_t)v16[(kw + 2 * t) * LDV + d] | ((uint32_t)v16[(kw + 2 * t + 1) * LDV + d] << 16);
          vh1 = (uint32_t)v16[(kw + 2 * t + 8) * LDV + d] | ((uint32_t)v16[(kw + 2 * t + 9) * LDV + d] << 16);
        }
#pragma unroll
        for (int mt = 0; mt < 2; mt++) {
          mma_bf16_16816(O[mt][nd], ph[mt], vh0, vh1);
          mma_bf16_16816(O[mt][nd], pl[mt], vh0, vh1);
          if constexpr (KV32) mma_bf16_16816(O[mt][nd], ph[mt], vl0, vl1);
        }
      }
    }
    __syncthreads();
  }
  // ---- merge the 8 warps' partials through shared memory (the staging ring is free now)
  float* Ow = reinterpret_cast<float*>(sk);                                // [8][32][64]
  float* mw = Ow + 8 * 32 * 64;                                            // [8][32]
  float* lw = mw + 8 * 32;
#pragma unroll
  for (int ri = 0; ri < 4; ri++) {
    lrow[ri] += __shfl_xor_sync(0xffffffffu, lrow[ri], 1);
    lrow[ri] += __shfl_xor_sync(0xffffffffu, lrow[ri], 2);
    const int m = (ri >> 1) * 16 + g + 8 * (ri & 1);
    if (t == 0) { mw[warp * 32 + m] = mrow[ri]; lw[warp * 32 + m] = lrow[ri]; }
  }
#pragma unroll
  for (int mt = 0; mt < 2; mt++)
#pragma unroll
    for (int nd = 0; nd < 8; nd++) {
      float* o0 = Ow + ((size_t)warp * 32 + mt * 16 + g) * 64 + nd * 8 + 2 * t;
      *reinterpret_cast<float2*>(o0) = make_float2(O[mt][nd][0], O[mt][nd][1]);
      *reinterpret_cast<float2*>(o0 + 8 * 64) = make_float2(O[mt][nd][2], O[mt][nd][3]);
    }
  __syncthreads();
  const int q_heads = a.kv_heads * a.group;
  // per-row merge weights of the 8 warps, once per row (thread = (warp slot w, row m)); mw is overwritten by the weights,
  // row maxima go to Mrow, row sums to Lrow
  float* Mrow = lw + 8 * 32;                                                // [32]
  float* Lrow = Mrow + 32;                                                  // [32]
  { const int w = tid >> 5, m = tid & 31;
    float M = -INFINITY;
#pragma unroll
    for (int k = 0; k < 8; k++) M = fmaxf(M, mw[k * 32 + m]);
    const float mv = mw[w * 32 + m];
    const float wgt = (mv == -INFINITY) ? 0.f : expf(mv - M);
    __syncthreads();                                                        // every thread has read the maxima of its row
    mw[w * 32 + m] = wgt;
    lw[w * 32 + m] *= wgt;
    if (w == 0) Mrow[m] = M;
    __syncthreads();
    if (w == 0) {
      float L = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) L += lw[k * 32 + m];
      Lrow[m] = L;
    }
    __syncthreads(); }
#pragma unroll
  for (int it = 0; it < 8; it++) {
    const int i = tid + it * 256;
    if (i < n_rows * 64) {
      const int m = i >> 6, d = i & 63;
      float acc = 0.f;
#pragma unroll
      for (int w = 0; w < 8; w++) acc += mw[w * 32 + m] * Ow[((size_t)w * 32 + m) * 64 + d];
      const int r = m / a.group, gq = m - r * a.group;
      const int row = seq * a.rows_per_seq + r, qh = kvh * a.group + gq;
      float* pp = a.part + (((size_t)row * q_heads + qh) * a.splits + split) * 68;
      pp[4 + d] = acc;
      if (d == 0) { pp[0] = Mrow[m]; pp[1] = Lrow[m]; }
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&a.counters[seq * a.kv_heads + kvh], 1) == a.splits - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // last slice of this (sequence, kv head): merge the slices in index order (deterministic) and write the attention rows.  A
  // thread owns up to 8 (row, dim) elements; the slice loop is the outer one so that the loads of all of them are in flight
  // together (one element after the other, each with its own dependent L2 round trips, cost ~10 us of the kernel's ~36)
  const float* pbs[8];
  float Mx[8], Ls[8], acc[8];
#pragma unroll
  for (int it = 0; it < 8; it++) {
    const int i = min(tid + it * 256, n_rows * 64 - 1), m = i >> 6;
    const int r = m / a.group, gq = m - r * a.group;
    pbs[it] = a.part + ((size_t)(seq * a.rows_per_seq + r) * q_heads + kvh * a.group + gq) * a.splits * 68;
    Mx[it] = -INFINITY; Ls[it] = 0.f; acc[it] = 0.f;
  }
  for (int sp = 0; sp < a.splits; sp++)
#pragma unroll
    for (int it = 0; it < 8; it++) Mx[it] = fmaxf(Mx[it], __ldcg(pbs[it] + sp * 68));
  for (int sp = 0; sp < a.splits; sp++) {
    float ms[8], ls[8], os[8];
#pragma unroll
    for (int it = 0; it < 8; it++) {
      const int d = (tid + it * 256) & 63;
      ms[it] = __ldcg(pbs[it] + sp * 68); ls[it] = __ldcg(pbs[it] + sp * 68 + 1); os[it] = __ldcg(pbs[it] + sp * 68 + 4 + d);
    }
#pragma unroll
    for (int it = 0; it < 8; it++) {
      const float w = (ms[it] == -INFINITY) ? 0.f : expf(ms[it] - Mx[it]);
      Ls[it] += ls[it] * w;
      acc[it] += os[it] * w;
    }
  }
#pragma unroll
  for (int it = 0; it < 8; it++) {
    const int i = tid + it * 256;
    if (i < n_rows * 64) {
      const int m = i >> 6, d = i & 63;
      const int r = m / a.group, gq = m - r * a.group;
      const int row = seq * a.rows_per_seq + r, qh = kvh * a.group + gq;
      const float y = acc[it] / Ls[it];
      if (a.out) a.out[(size_t)row * a.ldo + qh * 64 + d] = y;
      if (a.out16) store_split(a.out16 + (size_t)row * 2 * a.ldo, a.ldo, qh * 64 + d, y);
    }
  }
  if (tid == 0) a.counters[seq * a.kv_heads + kvh] = 0;
}

// ------------------------------------------------------------------ small kernels
// prompt rows [sos, embed_tokens(prompt_text || text), task_id, speech_embedding(prompt_speech)]  (:943-952)
__global__ void llm_prompt_rows_kernel(const int32_t* __restrict__ text, int n_text, const int32_t* __restrict__ pspeech,
                                       int n_ps, const __nv_bfloat16* __restrict__ embed, const __nv_bfloat16* __restrict__ semb,
                                       float* __restrict__ h, int H, int sos, int task, int text_vocab, int speech_vocab) {
  const int row = blockIdx.x;
  const __nv_bfloat16* src;
  if (row == 0) src = semb + (size_t)sos * H;
  else if (row <= n_text) { int id = text[row - 1]; id = min(max(id, 0), text_vocab - 1); src = embed + (size_t)id * H; }
  else if (row == n_text + 1) src = semb + (size_t)task * H;
  else { int id = pspeech[row - n_text - 2]; id = min(max(id, 0), speech_vocab - 1); src = semb + (size_t)id * H; }
  for (int i = threadIdx.x; i < H; i += blockDim.x) h[(size_t)row * H + i] = __bfloat162float(src[i]);
}

// (cos, sin)(pos * inv_freq[i]) with the very sincosf llm_qkv_store evaluates per element (bit-identical RoPE on both paths)
__global__ void llm_rope_table_kernel(const float* __restrict__ inv_freq, float2* __restrict__ tab, int max_ctx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 32 * max_ctx) return;
  float sn, cs;
  sincosf((float)(i >> 5) * inv_freq[i & 31], &sn, &cs);
  tab[i] = make_float2(cs, sn);
}

// rmsnorm rows for the tensor-core path, written as split bf16 [hi | lo] (row length 2H) so the GEMM sees
// fp32-accurate activations: out = norm_w * (x * rsqrt(mean(x^2) + eps))
__global__ void __launch_bounds__(256) llm_norm16_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          __nv_bfloat16* __restrict__ out, int rows, int H, float eps) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * H;
  float ss = 0.f;
  for (int k = lane; k < H; k += 32) { const float v = xr[k]; ss += v * v; }
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float sc = rsqrtf(ss / (float)H + eps);
  __nv_bfloat16* orow = out + (size_t)row * 2 * H;
  if ((H & 7) == 0) {
    // 8 features per lane and pass: one 16-byte store for the hi halves, one for the lo halves
    for (int k = lane * 8; k < H; k += 256) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float y0 = w[k + 2 * j] * (xr[k + 2 * j] * sc), y1 = w[k + 2 * j + 1] * (xr[k + 2 * j + 1] * sc);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(y0 - __low2float(h2), y1 - __high2float(h2));
        hi[j] = *reinterpret_cast<const uint32_t*>(&h2); lo[j] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      *reinterpret_cast<uint4*>(orow + k) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(orow + H + k) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  } else {
    for (int k = lane; k < H; k += 32) store_split(orow, H, k, w[k] * (xr[k] * sc));
  }
}

// hn[seq] = final rmsnorm of the last live row of every sequence (llm_multi_head_v3.py:258,883-886)
__global__ void __launch_bounds__(256) llm_last_norm_kernel(const float* __restrict__ h, const float* __restrict__ w,
                                                             const SeqState* __restrict__ seqs, int rows_per_seq,
                                                             float* __restrict__ hn, __nv_bfloat16* __restrict__ hn16, int H,
                                                             float eps) {
  __shared__ float red[8];
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int seq = blockIdx.x;
  const SeqState& s = seqs[seq];
  const int r = s.n_new > 0 ? s.n_new - 1 : 0;
  const float* xr = h + ((size_t)seq * rows_per_seq + r) * H;
  float ss = 0.f;
  for (int k = threadIdx.x; k < H; k += blockDim.x) { const float v = xr[k]; ss += v * v; }
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) tot += red[i];
  const float sc = rsqrtf(tot / (float)H + eps);
  for (int k = threadIdx.x; k < H; k += blockDim.x) {
    const float v = w[k] * (xr[k] * sc);
    hn[(size_t)seq * H + k] = v;
  }
}

__global__ void llm_f32_to_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16(x[i]);
}

// ------------------------------------------------------------------ sampler
struct SampArgs {
  const float* logits = nullptr;          // [head][seq][vocab]
  int n_seq = 0, head_k = 1, vocab = 0, stop_from = 0;
  int input_is_logp = 0;
  double top_p = 0.8; int top_k = 25; int win_size = 10; double rep_thr = 1.0;   // rep_thr = win_size * tau_r
  const float* u = nullptr; int u_stride = 0;
  SeqState* seqs = nullptr;
  int32_t* out_tokens = nullptr; int max_out = 0; int32_t* out_counts = nullptr;
  // next-step rows
  const __nv_bfloat16* semb = nullptr; float* h = nullptr; int H = 0;
  int* n_active = nullptr;
  int max_ctx = 0x7fffffff;               // KV-cache capacity: a sequence that would overflow it stops
  float* logp_out = nullptr;              // optional [head][seq][vocab]
  float* tables = nullptr;                // [seq][head][SAMP_TAB + vocab] scratch
  int* arrive = nullptr;                  // [seq] arrival counters (zero)
  // standalone mode (hvx_sample): explicit history instead of out_tokens
  const int32_t* history = nullptr; int n_history = 0; int min_len_override = -1;
  int32_t* ids_out = nullptr; int32_t* u_used = nullptr;
  unsigned long long* fused_bar = nullptr;  // grid-barrier counter of the fused decode-step kernel: re-armed here, between two steps
};

// block-wide reductions on double-buffered scratch: one __syncthreads each
__device__ __forceinline__ float block_max(float v, float* red /*[8]*/) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < 8; i++) r = fmaxf(r, red[i]);
  return r;
}
__device__ __forceinline__ float block_sum(float v, float* red /*[8]*/) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < 8; i++) r += red[i];
  return r;
}

constexpr int SAMP_THREADS = 256;
constexpr int SAMP_PER = 32;              // values per thread in registers: vocab <= 8192
constexpr int SAMP_TAB = 2 * SAMP_MAXK + 8;   // per (seq, head) table header: topv[64] | topi[64] | n | pad

// grid (n_seq, head_k).  Phase 1 (every block, one head of one sequence): log_softmax -> softmax (common.py:149),
// stable descending top-k prefix with cum < top_p (:150-156) and the fp64-accumulated cumulative sums needed by the
// inverse-CDF draws, written to a small table.  Phase 2 (the block that arrives last for its sequence): the serial
// part — draws on the explicit uniform stream (oracle/llm_ref.py documents the order: one u per multinomial call),
// repetition-aware fallback over the full vocabulary (:139-143), EOS retry (llm_multi_head_v3.py:151-166), the
// emit / stop logic of :902-916 and the embedding fetch of the next step's rows (:919-922).
__global__ void __launch_bounds__(SAMP_THREADS) llm_sampler_kernel(SampArgs a) {
  extern __shared__ float p[];                       // [vocab] probabilities
  __shared__ float red[4][8];
  __shared__ int redi[2][8];
  __shared__ double dwarp[8];
  __shared__ int s_flag;
  __shared__ int s_ids[LLM_MAX_HEADS], s_group, s_alive;
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int seq = blockIdx.x, j = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, V = a.vocab;
  if (a.fused_bar && seq == 0 && j == 0 && tid == 0) *a.fused_bar = 0ull;
  SeqState* st = a.seqs ? &a.seqs[seq] : nullptr;
  if (st && st->done) return;
  const int topk = a.top_k < SAMP_MAXK ? a.top_k : SAMP_MAXK;
  float* tab = a.tables + ((size_t)seq * a.head_k + j) * (SAMP_TAB + V);
  float* cum = tab + SAMP_TAB;

  // ---------------- phase 1
  {
    const float* x = a.logits + ((size_t)j * a.n_seq + seq) * V;
    float v[SAMP_PER];
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < SAMP_PER; q++) { const int i = tid + q * SAMP_THREADS; v[q] = i < V ? x[i] : -INFINITY; mx = fmaxf(mx, v[q]); }
    mx = block_max(mx, red[0]);
    float lse = 0.f;
    if (!a.input_is_logp) {
      float se = 0.f;
#pragma unroll
      for (int q = 0; q < SAMP_PER; q++) if (tid + q * SAMP_THREADS < V) se += expf(v[q] - mx);
      lse = logf(block_sum(se, red[1]));
    }
    // softmax(log_softmax(x)): the maximum of logp is (mx - mx) - lse = -lse exactly (or max(x) if x already is logp)
    const float m2 = a.input_is_logp ? mx : -lse;
    float z = 0.f;
#pragma unroll
    for (int q = 0; q < SAMP_PER; q++) {
      const int i = tid + q * SAMP_THREADS;
      if (i < V) {
        const float lp = a.input_is_logp ? v[q] : (v[q] - mx) - lse;
        if (a.logp_out) a.logp_out[((size_t)j * a.n_seq + seq) * V + i] = lp;
        v[q] = expf(lp - m2);
        z += v[q];
      } else v[q] = -1.f;
    }
    z = block_sum(z, red[2]);
    float bv = -1.f; int bq = 0;
#pragma unroll
    for (int q = 0; q < SAMP_PER; q++) {
      const int i = tid + q * SAMP_THREADS;
      if (i < V) { v[q] = v[q] / z; p[i] = v[q]; if (v[q] > bv) { bv = v[q]; bq = q; } }
    }
    __syncthreads();
    // inclusive cumulative sum accumulated in fp64, rounded to fp32 per element (torch CPU cumsum)
    {
      const int per = (V + SAMP_THREADS - 1) / SAMP_THREADS;
      const int i0 = tid * per, i1 = min(V, i0 + per);
      double s = 0.0;
      for (int i = i0; i < i1; i++) s += (double)p[i];
      double inc = s;                                   // warp-inclusive scan of the segment sums
      for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      if (lane == 31) dwarp[warp] = inc;
      __syncthreads();
      double run = inc - s;
      for (int w = 0; w < warp; w++) run += dwarp[w];
      for (int i = i0; i < i1; i++) { run += (double)p[i]; cum[i] = (float)run; }
    }
    // stable descending selection of the nucleus prefix: every thread caches the best of its own values
    float csum = 0.f;
    int n = 0;
    for (int t = 0; t < topk; t++) {
      float cv = bv; int ci = bv >= 0.f ? tid + bq * SAMP_THREADS : 0x7fffffff;
      for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, cv, o); const int oi = __shfl_xor_sync(0xffffffffu, ci, o);
        if (ov > cv || (ov == cv && oi < ci)) { cv = ov; ci = oi; }
      }
      float* rv = red[t & 1]; int* ri = redi[t & 1];
      if (lane == 0) { rv[warp] = cv; ri[warp] = ci; }
      __syncthreads();
      cv = rv[0]; ci = ri[0];
#pragma unroll
      for (int w = 1; w < 8; w++) if (rv[w] > cv || (rv[w] == cv && ri[w] < ci)) { cv = rv[w]; ci = ri[w]; }
      if (tid == 0) { tab[t] = cv; reinterpret_cast<int*>(tab)[SAMP_MAXK + t] = ci; }
      csum = csum + cv;                               // cum = cum + sv[n] in fp32, same in every thread
      n = t + 1;
      if ((ci & (SAMP_THREADS - 1)) == tid) {         // owner: retire the winner, rescan own values
        const int qq = ci / SAMP_THREADS;
        bv = -1.f; bq = 0;
#pragma unroll
        for (int q = 0; q < SAMP_PER; q++) { if (q == qq) v[q] = -1.f; if (v[q] > bv) { bv = v[q]; bq = q; } }
      }
      if (!((double)csum < a.top_p)) break;
    }
    if (tid == 0) reinterpret_cast<int*>(tab)[2 * SAMP_MAXK] = n;
  }
  // ---------------- hand over to the last block of this sequence
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(&a.arrive[seq], 1) == a.head_k - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();

  // ---------------- phase 2 (one thread; the tables are tiny)
  if (tid == 0) {
    a.arrive[seq] = 0;
    const int n_hist = st ? st->n_out : a.n_history;
    const int32_t* hist = st ? a.out_tokens + (size_t)seq * a.max_out : a.history;
    const int min_len = st ? st->min_len : a.min_len_override;
    int u_pos = st ? st->u_pos : 0;
    int status = 0;
    for (int h = 0; h < a.head_k; h++) {
      const float* tb = a.tables + ((size_t)seq * a.head_k + h) * (SAMP_TAB + V);
      const float* cm = tb + SAMP_TAB;
      const int n = __ldcg(reinterpret_cast<const int*>(tb) + 2 * SAMP_MAXK);
      float kc[SAMP_MAXK];
      { double run = 0.0; for (int t = 0; t < n; t++) { run += (double)__ldcg(tb + t); kc[t] = (float)run; } }
      const bool ignore_eos = (n_hist + h) < min_len;
      int trials = 0, tok = 0;
      while (true) {
        if (u_pos + 2 > a.u_stride) { status = 2; tok = a.stop_from; break; }
        const float target = a.u[(size_t)seq * a.u_stride + u_pos++] * kc[n - 1];
        int i = 0;
        while (i < n - 1 && !(kc[i] > target)) i++;
        tok = __ldcg(reinterpret_cast<const int*>(tb) + SAMP_MAXK + i);
        // repetition check over the last win_size emitted tokens (whole history when win_size == 0)
        const int lo = a.win_size != 0 ? max(0, n_hist - a.win_size) : 0;
        int rep = 0;
        for (int q = lo; q < n_hist; q++) rep += (hist[q] == tok);
        if ((double)rep >= a.rep_thr) {
          const float tg = a.u[(size_t)seq * a.u_stride + u_pos++] * __ldcg(cm + V - 1);
          int lo2 = 0, hi2 = V - 1;                               // first index with cum > tg
          while (lo2 < hi2) { const int mid = (lo2 + hi2) >> 1; if (__ldcg(cm + mid) > tg) hi2 = mid; else lo2 = mid + 1; }
          tok = lo2;
        }
        if (!ignore_eos || tok < a.stop_from) break;
        if (++trials > 100) { status = 1; break; }
      }
      s_ids[h] = tok;
    }
    if (a.u_used) a.u_used[seq] = u_pos;
    if (a.ids_out) for (int h = 0; h < a.head_k; h++) a.ids_out[seq * a.head_k + h] = s_ids[h];
    s_group = 0; s_alive = 0;
    if (st) {
      int n_out = st->n_out, group = 0, stop = (status != 0);
      for (int h = 0; h < a.head_k && !stop; h++) {
        const int t = s_ids[h];
        if (t >= a.stop_from) { stop = 1; break; }
        a.out_tokens[(size_t)seq * a.max_out + n_out++] = t;
        st->new_tok[group++] = t;
        if (n_out >= st->max_len || n_out >= a.max_out) { stop = 1; break; }
      }
      st->n_out = n_out;
      st->u_pos = u_pos;
      st->status = status;
      st->ctx += st->ctx_add;
      __threadfence();                              // tokens before the count: a streaming consumer polls out_counts from another stream
      if (a.out_counts) a.out_counts[seq] = n_out;
      if (stop || group == 0 || st->ctx + group > a.max_ctx) { st->done = 1; st->n_new = 0; st->ctx_add = 0; atomicSub(a.n_active, 1); }
      else { st->n_new = group; st->ctx_add = group; s_group = group; s_alive = 1; }
    }
  }
  __syncthreads();
  if (st && s_alive) {
    for (int r = 0; r < s_group; r++) {
      const __nv_bfloat16* src = a.semb + (size_t)st->new_tok[r] * a.H;
      float* dst = a.h + ((size_t)seq * a.head_k + r) * a.H;
      for (int i = tid; i < a.H; i += blockDim.x) dst[i] = __bfloat162float(src[i]);
    }
  }
}

// ------------------------------------------------------------------ host state
struct LlmLayer {
  const __nv_bfloat16 *qkv_w, *o_w, *gu_w, *down_w;
  const float *qkv_b, *ln1, *ln2;
};

}  // namespace hvx
#include "llm_fused.cuh"
namespace hvx {

struct SeqDesc {
  const int32_t* text = nullptr; int n_text_total = 0, n_text_new = 0;
  const int32_t* pspeech = nullptr; int n_ps = 0;
  float min_ratio = 2.f, max_ratio = 20.f;
  bool begun = false;
};

struct LlmState {
  LlmLayer layer[64];
  const __nv_bfloat16 *embed, *semb, *dec_w;
  const float* norm;
  const __nv_bfloat16 *m_v_w, *m_o_w, *m_gu_w, *m_down_w;     // stacked over MTP heads
  const float *m_v_b, *m_ln1, *m_ln2;
  float* inv_freq = nullptr;
  float2* rope = nullptr;                 // [max_ctx][32] (cos, sin) of pos * inv_freq[i]: the tensor-core path's RoPE operands
  uint8_t *kc = nullptr, *vc = nullptr;                       // [layer][seq][kv_head][max_ctx][64], bf16 or fp32
  int kv_f32 = 0; size_t kv_esz = 2;
  size_t layer_stride = 0, seq_stride = 0;
  SeqState* seqs = nullptr;
  int* n_active = nullptr;
  int max_ctx = 0x7fffffff;               // KV-cache capacity: a sequence that would overflow it stops
  int* n_active_host = nullptr;                               // pinned
  std::vector<SeqDesc> desc;
  DevBuf ws, pws;
  cudaStream_t own = nullptr;                                 // decode runs here: the caller's stream may be the legacy
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;              // default stream, which cannot be captured into a graph
  cudaGraphExec_t graph = nullptr; int graph_key[4] = {0, 0, 0, 0};
  SampArgs graph_samp;
  // fused persistent decode step (llm_fused.cuh)
  LlmLayer* layers_dev = nullptr;
  unsigned long long* fused_bar = nullptr;
  int* fused_abort = nullptr;
};

template <typename T>
static hvx_status lget(hvx_engine* e, const std::string& name, int dtype, const T** p, int64_t numel) {
  const Tensor* t = e->find(HVX_STAGE_LLM, name);
  HVX_CHECK(t && t->dtype == dtype, HVX_ERR_STATE, "llm: missing tensor %s (or wrong dtype)", name.c_str());
  HVX_CHECK(t->numel() == numel, HVX_ERR_STATE, "llm: tensor %s has %lld elements, expected %lld", name.c_str(),
            (long long)t->numel(), (long long)numel);
  *p = reinterpret_cast<const T*>(t->p);
  return HVX_OK;
}
#define LLM_GET(call) do { hvx_status _s = (call); if (_s) return _s; } while (0)

hvx_status llm_finalize(hvx_engine* e) {
  const hvx_config& c = e->cfg;
  const int64_t H = c.llm_hidden, QD = (int64_t)c.llm_q_heads * c.llm_head_dim, KD = (int64_t)c.llm_kv_heads * c.llm_head_dim;
  HVX_CHECK(c.llm_head_dim == 64, HVX_ERR_UNSUPPORTED, "llm: head_dim must be 64");
  HVX_CHECK(QD == H, HVX_ERR_UNSUPPORTED, "llm: q_heads*head_dim must equal hidden");
  HVX_CHECK(H % 64 == 0 && H <= 2048 && c.llm_inter % 64 == 0 && c.llm_mtp_inter % 64 == 0, HVX_ERR_UNSUPPORTED, "llm: unsupported dims");
  HVX_CHECK(c.llm_q_heads % c.llm_kv_heads == 0 && c.llm_q_heads / c.llm_kv_heads <= 8, HVX_ERR_UNSUPPORTED, "llm: GQA group > 8");
  HVX_CHECK(c.llm_layers <= 64 && c.llm_mtp_heads <= LLM_MAX_HEADS, HVX_ERR_UNSUPPORTED, "llm: too many layers/heads");
  HVX_CHECK(c.llm_max_seqs >= 1 && c.llm_max_ctx >= 16, HVX_ERR_ARG, "llm: bad max_seqs/max_ctx");
  if (!e->llm) e->llm = new LlmState();
  LlmState* L = e->llm;
  LLM_GET(lget(e, "embed", HVX_BF16, &L->embed, (int64_t)c.llm_text_vocab * H));
  LLM_GET(lget(e, "speech_emb", HVX_BF16, &L->semb, (int64_t)c.llm_speech_vocab * H));
  LLM_GET(lget(e, "dec.w", HVX_BF16, &L->dec_w, (int64_t)c.llm_speech_vocab * H));
  LLM_GET(lget(e, "norm", HVX_F32, &L->norm, H));
  for (int l = 0; l < c.llm_layers; l++) {
    const std::string p = "L" + std::to_string(l) + ".";
    LlmLayer& y = L->layer[l];
    LLM_GET(lget(e, p + "qkv.w", HVX_BF16, &y.qkv_w, (QD + 2 * KD) * H));
    LLM_GET(lget(e, p + "qkv.b", HVX_F32, &y.qkv_b, QD + 2 * KD));
    LLM_GET(lget(e, p + "o.w", HVX_BF16, &y.o_w, H * QD));
    LLM_GET(lget(e, p + "gu.w", HVX_BF16, &y.gu_w, (int64_t)2 * c.llm_inter * H));
    LLM_GET(lget(e, p + "down.w", HVX_BF16, &y.down_w, H * c.llm_inter));
    LLM_GET(lget(e, p + "ln1", HVX_F32, &y.ln1, H));
    LLM_GET(lget(e, p + "ln2", HVX_F32, &y.ln2, H));
  }
  const int64_t MH = c.llm_mtp_heads, MI = c.llm_mtp_inter;
  LLM_GET(lget(e, "mtp.v.w", HVX_BF16, &L->m_v_w, MH * H * H));
  LLM_GET(lget(e, "mtp.v.b", HVX_F32, &L->m_v_b, MH * H));
  LLM_GET(lget(e, "mtp.o.w", HVX_BF16, &L->m_o_w, MH * H * H));
  LLM_GET(lget(e, "mtp.gu.w", HVX_BF16, &L->m_gu_w, MH * 2 * MI * H));
  LLM_GET(lget(e, "mtp.down.w", HVX_BF16, &L->m_down_w, MH * H * MI));
  LLM_GET(lget(e, "mtp.ln1", HVX_F32, &L->m_ln1, MH * H));
  LLM_GET(lget(e, "mtp.ln2", HVX_F32, &L->m_ln2, MH * H));
  if (!L->kc) {
    L->seq_stride = (size_t)c.llm_kv_heads * c.llm_max_ctx * 64;
    L->layer_stride = L->seq_stride * c.llm_max_seqs;
    L->kv_f32 = c.llm_kv_f32 ? 1 : 0;
    L->kv_esz = L->kv_f32 ? 4 : 2;
    const size_t bytes = L->layer_stride * c.llm_layers * L->kv_esz;
    HVX_CUDA(cudaMalloc(&L->kc, bytes));
    HVX_CUDA(cudaMalloc(&L->vc, bytes));
    HVX_CUDA(cudaMemset(L->kc, 0, bytes));
    HVX_CUDA(cudaMemset(L->vc, 0, bytes));
    HVX_CUDA(cudaMalloc(&L->seqs, sizeof(SeqState) * c.llm_max_seqs));
    HVX_CUDA(cudaMemset(L->seqs, 0, sizeof(SeqState) * c.llm_max_seqs));
    HVX_CUDA(cudaMalloc(&L->n_active, sizeof(int)));
    HVX_CUDA(cudaMallocHost(&L->n_active_host, sizeof(int)));
    float inv[32];
    for (int i = 0; i < 32; i++) inv[i] = 1.0f / powf(c.llm_rope_theta, (float)(2 * i) / 64.0f);   // HF Qwen2RotaryEmbedding
    HVX_CUDA(cudaMalloc(&L->inv_freq, sizeof(inv)));
    HVX_CUDA(cudaMemcpy(L->inv_freq, inv, sizeof(inv), cudaMemcpyHostToDevice));
    HVX_CUDA(cudaMalloc(&L->rope, sizeof(float2) * 32 * (size_t)c.llm_max_ctx));
    llm_rope_table_kernel<<<cdiv(32 * c.llm_max_ctx, 256), 256>>>(L->inv_freq, L->rope, c.llm_max_ctx);
    HVX_CUDA(cudaDeviceSynchronize());
    L->desc.resize(c.llm_max_seqs);
    HVX_CUDA(cudaMalloc(&L->layers_dev, sizeof(LlmLayer) * 64));
    HVX_CUDA(cudaMalloc(&L->fused_bar, sizeof(unsigned long long)));
    HVX_CUDA(cudaMemset(L->fused_bar, 0, sizeof(unsigned long long)));
    HVX_CUDA(cudaMalloc(&L->fused_abort, sizeof(int)));
    HVX_CUDA(cudaMemset(L->fused_abort, 0, sizeof(int)));
    // the decode stream gets the highest priority: its small latency-bound kernels take the SMs that free up first when a
    // flow / vocoder group of finished utterances runs beside it (hvx_synthesize_host); HVX_LLM_STREAM_PRIO=0 turns that off
    { int lo_p = 0, hi_p = 0;
      HVX_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
      const bool prio = !(getenv("HVX_LLM_STREAM_PRIO") && atoi(getenv("HVX_LLM_STREAM_PRIO")) == 0);
      HVX_CUDA(cudaStreamCreateWithPriority(&L->own, cudaStreamNonBlocking, prio ? hi_p : lo_p)); }
    HVX_CUDA(cudaEventCreateWithFlags(&L->ev_in, cudaEventDisableTiming));
    HVX_CUDA(cudaEventCreateWithFlags(&L->ev_out, cudaEventDisableTiming));
  }
  if (L->graph) { cudaGraphExecDestroy(L->graph); L->graph = nullptr; }
  HVX_CUDA(cudaMemcpy(L->layers_dev, L->layer, sizeof(LlmLayer) * 64, cudaMemcpyHostToDevice));
  return HVX_OK;
}

void llm_free(hvx_engine* e) {
  LlmState* L = e->llm;
  if (!L) return;
  if (L->graph) cudaGraphExecDestroy(L->graph);
  cudaFree(L->kc); cudaFree(L->vc); cudaFree(L->seqs); cudaFree(L->n_active); cudaFree(L->inv_freq); cudaFree(L->rope);
  cudaFree(L->layers_dev); cudaFree(L->fused_bar); cudaFree(L->fused_abort);
  if (L->n_active_host) cudaFreeHost(L->n_active_host);
  if (L->own) cudaStreamDestroy(L->own);
  if (L->ev_in) cudaEventDestroy(L->ev_in);
  if (L->ev_out) cudaEventDestroy(L->ev_out);
  delete L;
  e->llm = nullptr;
}

// ---- launch helpers
// launch with programmatic dependent launch allowed: the kernel may be scheduled while its predecessor drains
// (every such kernel executes griddepcontrol.wait before it touches dependent data)
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <int R>
static hvx_status launch_gemv_tma(hvx_engine* e, cudaStream_t st, GemvArgs a, int n_batch) {
  static bool attr = false;
  if (!attr) { HVX_CUDA(cudaFuncSetAttribute(llm_gemv_tma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)); attr = true; }
  const int pairs = (a.N + 1) / 2;
  // one CTA per SM (all of them request bytes concurrently), each with a contiguous range of pairs
  const int ctas = std::max(1, std::min(e->sm_count / n_batch, pairs));
  // ring sized to the CTA's share of W (whole pairs per slot), capped so that two CTAs fit one SM: the next kernel's
  // CTA must be able to become resident (and prefetch its weights) while this one still computes
  const int ppc = cdiv(pairs, ctas);
  const size_t pair_bytes = (size_t)4 * a.K;
  const size_t x_bytes = ((size_t)2 * R * a.K + a.K) * sizeof(float) + 128;      // two-plane rows + norm weights + landing rows
  const size_t budget = 216 * 1024 - x_bytes;
  size_t slot = std::min((size_t)GT_SLOT / pair_bytes, (size_t)ppc) * pair_bytes;        // whole pairs, at most the share
  slot = std::max(slot, pair_bytes);
  int ns = (int)std::min((size_t)GT_NS, std::max((size_t)1, budget / slot));
  ns = std::min(ns, cdiv(ppc, (int)(slot / pair_bytes)));
  HVX_CHECK(slot * ns + x_bytes <= 220 * 1024, HVX_ERR_UNSUPPORTED, "gemv: shared memory budget exceeded (K=%d R=%d)", a.K, R);
  a.slot_bytes = (int)slot; a.n_slots = ns;
  { const char* pb = getenv("HVX_GEMV_PIECE"); a.piece_bytes = pb ? atoi(pb) : 4096; }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas, n_batch, 1);
  cfg.blockDim = dim3((GT_CONS + 1) * 32, 1, 1);
  cfg.dynamicSmemBytes = slot * ns + x_bytes;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  HVX_CUDA(cudaLaunchKernelEx(&cfg, llm_gemv_tma_kernel<R>, a));
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

static hvx_status launch_gemv8(hvx_engine* e, cudaStream_t st, GemvArgs a, int R, int n_batch) {
  a.rows = R;
  {
    const int Rt0 = R <= 1 ? 1 : R <= 2 ? 2 : R <= 4 ? 4 : 8;
    const size_t x_bytes = ((size_t)2 * Rt0 * a.K + a.K) * sizeof(float) + 128;
    if (a.K <= GT_KMAX && a.K % 8 == 0 && x_bytes + (size_t)4 * a.K <= 216 * 1024 && !getenv("HVX_NO_TMA_GEMV")) {
      switch (Rt0) {
        case 1: return launch_gemv_tma<1>(e, st, a, n_batch);
        case 2: return launch_gemv_tma<2>(e, st, a, n_batch);
        case 4: return launch_gemv_tma<4>(e, st, a, n_batch);
        default: return launch_gemv_tma<8>(e, st, a, n_batch);
      }
    }
  }
  const int pairs = (a.N + 1) / 2;
  const int Rt = R <= 1 ? 1 : R <= 2 ? 2 : R <= 4 ? 4 : 8;
  HVX_CHECK(a.K % 8 == 0, HVX_ERR_UNSUPPORTED, "gemv: K must be a multiple of 8");
  // activation staging: the whole row when it fits ~96 KB, else chunks (then every warp owns at most one pair)
  const int kc_max = (96 * 1024 / 4 / Rt) & ~7;
  a.kc = a.K <= kc_max ? a.K : kc_max;
  HVX_CHECK(!a.norm_w || a.K <= a.kc, HVX_ERR_UNSUPPORTED, "gemv: fused norm needs the whole row staged (K=%d)", a.K);
  const size_t smem = ((size_t)Rt * a.kc + (a.norm_w ? a.K : 0)) * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    HVX_CUDA(cudaFuncSetAttribute(llm_gemv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    HVX_CUDA(cudaFuncSetAttribute(llm_gemv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    HVX_CUDA(cudaFuncSetAttribute(llm_gemv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    HVX_CUDA(cudaFuncSetAttribute(llm_gemv_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_done = true;
  }
  int warps = 8;
  int ctas;
  if (a.K > a.kc) {
    ctas = cdiv(pairs, warps);
  } else {
    // persistent: at most two CTAs per SM; small matrices get 4-warp CTAs so that ~one CTA per SM covers them
    if (pairs * n_batch <= 4 * e->sm_count) warps = 4;
    int occ = 1;
    switch (Rt) {
      case 1: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, llm_gemv_kernel<1>, warps * 32, smem); break;
      case 2: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, llm_gemv_kernel<2>, warps * 32, smem); break;
      case 4: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, llm_gemv_kernel<4>, warps * 32, smem); break;
      default: cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, llm_gemv_kernel<8>, warps * 32, smem); break;
    }
    occ = std::max(1, std::min(occ, 2));
    ctas = std::min(cdiv(pairs, warps), std::max(1, occ * e->sm_count / n_batch));     // one resident wave
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ctas, n_batch, 1);
  cfg.blockDim = dim3(warps * 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
#define GEMV_CASE(RR, idx)                                                                                        \
  case RR:                                                                                                        \
    HVX_CUDA(cudaLaunchKernelEx(&cfg, llm_gemv_kernel<RR>, a));                                                   \
    break;
  switch (Rt) {
    GEMV_CASE(1, 0)
    GEMV_CASE(2, 1)
    GEMV_CASE(4, 2)
    GEMV_CASE(8, 3)
  }
#undef GEMV_CASE
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

// rows <= GEMV_MAX_ROWS: weight-streaming GEMV passes of 8 rows; above: tcgen05 GEMMs on split-bf16 activations.  With the
// split-K skinny GEMMs the tensor-core path wins from the second pass on (16 rows: 2.76 vs 3.06 ms per step), so the default is
// one pass; HVX_GEMV_MAX_ROWS=32 restores the multi-pass GEMV (kept and tested: tests/test_llm_gpu.py sets it)
static int gemv_max_rows() {                 // read per call: tests switch paths with the environment variable
  const char* s = getenv("HVX_GEMV_MAX_ROWS");
  return s ? std::max(8, std::min(32, atoi(s))) : 8;
}
#define GEMV_MAX_ROWS gemv_max_rows()

// rows > 8: one pass over the weights per group of 8 rows
static hvx_status launch_gemv(hvx_engine* e, cudaStream_t st, GemvArgs a, int R, int n_batch = 1) {
  for (int r0 = 0; r0 < R; r0 += 8) {
    GemvArgs g = a;
    g.x = a.x + (size_t)r0 * a.ldx;
    if (a.out) g.out = a.out + (size_t)r0 * a.ldo;
    if (a.resid) g.resid = a.resid + (size_t)r0 * a.ldr;
    g.row0 = r0;
    hvx_status rc = launch_gemv8(e, st, g, std::min(8, R - r0), n_batch);
    if (rc) return rc;
  }
  return HVX_OK;
}

struct StepBufs {          // per-step activations, `rows` row slots
  size_t part_floats = 0;  // capacity of `part` (attention partials; between attention calls it holds the split-K partial sums)
  float *h, *q, *att, *act, *hn, *m_v, *m_h1, *m_act, *m_o, *logits, *part;
  __nv_bfloat16 *x16, *att16, *act16, *hn16, *m16;
  int* counters;
};

static hvx_status llm_bufs(hvx_engine* e, LlmState* L, DevBuf& buf, int rows, int n_seq, int head_k, int splits, StepBufs* b,
                           cudaStream_t st) {
  const hvx_config& c = e->cfg;
  const size_t H = c.llm_hidden, I = c.llm_inter, MI = c.llm_mtp_inter, V = c.llm_speech_vocab;
  const size_t rp = (size_t)(rows + 127) / 128 * 128, sp = (size_t)(n_seq + 127) / 128 * 128;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_h = take(rp * H * 4), o_q = take(rp * H * 4), o_att = take(rp * H * 4), o_act = take(rp * I * 4);
  const size_t o_hn = take(sp * H * 4), o_mv = take(head_k * sp * H * 4), o_mh1 = take(head_k * sp * H * 4);
  const size_t o_mact = take(head_k * sp * MI * 4), o_mo = take(head_k * sp * H * 4), o_log = take(head_k * sp * V * 4);
  const int part_splits = std::max(splits, e->sm_count / std::max(1, c.llm_kv_heads));     // the fused step splits over all SMs
  const size_t o_part = take((size_t)rows * c.llm_q_heads * part_splits * 68 * 4), o_cnt = take((size_t)rows * c.llm_kv_heads * 4);
  b->part_floats = (size_t)rows * c.llm_q_heads * part_splits * 68;
  const size_t o_x16 = take(rp * H * 4), o_att16 = take(rp * H * 4), o_act16 = take(std::max(rp * I, sp * MI) * 4);
  const size_t o_hn16 = take(256), o_m16 = take(sp * H * 4);        // split bf16 [hi | lo] rows
  const bool grew = off > buf.bytes;
  uint8_t* w = (uint8_t*)buf.get(off);
  HVX_CHECK(w, HVX_ERR_CUDA, "llm: workspace allocation of %zu bytes failed", off);
  if (grew) HVX_CUDA(cudaMemsetAsync(w, 0, buf.bytes, st));
  b->h = (float*)(w + o_h); b->q = (float*)(w + o_q); b->att = (float*)(w + o_att); b->act = (float*)(w + o_act);
  b->hn = (float*)(w + o_hn); b->m_v = (float*)(w + o_mv); b->m_h1 = (float*)(w + o_mh1); b->m_act = (float*)(w + o_mact);
  b->m_o = (float*)(w + o_mo); b->logits = (float*)(w + o_log); b->part = (float*)(w + o_part); b->counters = (int*)(w + o_cnt);
  b->x16 = (__nv_bfloat16*)(w + o_x16); b->att16 = (__nv_bfloat16*)(w + o_att16); b->act16 = (__nv_bfloat16*)(w + o_act16);
  b->hn16 = (__nv_bfloat16*)(w + o_hn16); b->m16 = (__nv_bfloat16*)(w + o_m16);
  // the split-attention arrival counters must start at zero; the carve-up moves with (rows, n_seq, head_k)
  HVX_CUDA(cudaMemsetAsync(b->counters, 0, (size_t)rows * c.llm_kv_heads * 4, st));
  return HVX_OK;
}

static LlmQkvEpi make_qkv(hvx_engine* e, LlmState* L, int layer, float* q, const SeqState* seqs, int rows_per_seq, int seq0,
                          int pos0, int n_rows) {
  const hvx_config& c = e->cfg;
  LlmQkvEpi p;
  p.q_out = q; p.ldq = c.llm_hidden;
  p.kc = L->kc + (size_t)layer * L->layer_stride * L->kv_esz; p.vc = L->vc + (size_t)layer * L->layer_stride * L->kv_esz;
  p.kv_f32 = L->kv_f32;
  p.seq_stride = L->seq_stride; p.max_ctx = c.llm_max_ctx; p.q_dim = c.llm_q_heads * 64; p.kv_dim = c.llm_kv_heads * 64;
  p.inv_freq = L->inv_freq; p.rope = L->rope; p.seqs = seqs; p.rows_per_seq = rows_per_seq; p.seq0 = seq0; p.pos0 = pos0; p.n_rows = n_rows;
  return p;
}

static hvx_status launch_attn(hvx_engine* e, cudaStream_t st, LlmState* L, int layer, const StepBufs& b, int rows,
                              const SeqState* seqs, int rows_per_seq, int seq0, int pos0, int splits, bool want16) {
  const hvx_config& c = e->cfg;
  AttnDecArgs a;
  a.q = b.q; a.ldq = c.llm_hidden;
  a.kc = L->kc + (size_t)layer * L->layer_stride * L->kv_esz; a.vc = L->vc + (size_t)layer * L->layer_stride * L->kv_esz;
  a.seq_stride = L->seq_stride; a.max_ctx = c.llm_max_ctx; a.kv_heads = c.llm_kv_heads; a.group = c.llm_q_heads / c.llm_kv_heads;
  a.seqs = seqs; a.rows_per_seq = rows_per_seq; a.seq0 = seq0; a.pos0 = pos0; a.n_rows = rows; a.splits = splits;
  a.part = b.part; a.counters = b.counters; a.out = b.att; a.out16 = want16 ? b.att16 : nullptr; a.ldo = c.llm_hidden;
  a.scale = 1.0f / sqrtf((float)c.llm_head_dim);
  // batched decode (tensor-core path of the step): one block per (sequence, kv head, key slice); the rows of a sequence share
  // the staged cache chunk.  Default: the mma.sync kernel (<= 32 query rows per kv head); HVX_ATTN_DECODE=seq: the CUDA-core
  // per-sequence kernel; HVX_ATTN_DECODE=row: one block per row (llm_attn_kernel)
  static const char* mode_env = getenv("HVX_ATTN_DECODE");
  const int mode = !mode_env ? 2 : (!strcmp(mode_env, "row") ? 0 : (!strcmp(mode_env, "seq") ? 1 : 2));
  const int n_seq_b = rows_per_seq > 0 ? rows / rows_per_seq : 0;
  if (seqs && want16 && mode == 2 && rows_per_seq >= 1 && rows_per_seq * a.group <= 32 && rows_per_seq <= 4) {
    a.splits = std::max(1, std::min(16, e->sm_count / std::max(1, n_seq_b * c.llm_kv_heads)));
    const size_t kv_bytes = L->kv_f32 ? (size_t)2 * AM_CH * (AM_LDK32 + AM_LDV32) * 4 : (size_t)2 * AM_CH * 2 * AM_LD16 * 2;
    const size_t merge_bytes = (size_t)(8 * 32 * 64 + 2 * 8 * 32 + 64) * 4;
    const size_t smem = (size_t)2 * 32 * AM_LDQ * 2 + std::max(kv_bytes, merge_bytes);
    static bool attr_set = false;
    if (!attr_set) {
      HVX_CUDA(cudaFuncSetAttribute(llm_attn_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      HVX_CUDA(cudaFuncSetAttribute(llm_attn_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    HVX_CHECK((size_t)rows * c.llm_q_heads * a.splits * 68 <= b.part_floats, HVX_ERR_STATE, "llm: attention partial buffer too small");
    dim3 sgrid(a.splits, n_seq_b * c.llm_kv_heads);
    if (L->kv_f32) HVX_CUDA(launch_pdl(llm_attn_mma_kernel<true>, sgrid, dim3(256), smem, st, a));
    else HVX_CUDA(launch_pdl(llm_attn_mma_kernel<false>, sgrid, dim3(256), smem, st, a));
    HVX_LAUNCH_CHECK(e);
    return HVX_OK;
  }
  if (seqs && want16 && mode >= 1 && rows_per_seq >= 1 && 32 * a.group * ((rows_per_seq + 1) / 2) <= 512) {
    const int n_seq = n_seq_b, pairs = (rows_per_seq + 1) / 2;
    a.splits = std::max(1, std::min(16, e->sm_count / std::max(1, n_seq * c.llm_kv_heads)));
    const size_t elt = L->kv_f32 ? (size_t)(ATT_KEYS / 2) * ATT_LD32 * 4 : (size_t)ATT_KEYS * ATT_LD * 2;
    const size_t smem = 4 * elt;                                            // K and V, two buffers each
    static bool attr_set = false;
    if (!attr_set) {
      HVX_CUDA(cudaFuncSetAttribute(llm_attn_seq_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (ATT_KEYS / 2) * ATT_LD32 * 4));
      HVX_CUDA(cudaFuncSetAttribute(llm_attn_seq_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * ATT_KEYS * ATT_LD * 2));
      attr_set = true;
    }
    HVX_CHECK((size_t)rows * c.llm_q_heads * a.splits * 68 <= b.part_floats, HVX_ERR_STATE, "llm: attention partial buffer too small");
    dim3 sgrid(a.splits, n_seq * c.llm_kv_heads);
    if (L->kv_f32) HVX_CUDA(launch_pdl(llm_attn_seq_kernel<true>, sgrid, dim3(32 * a.group * pairs), smem, st, a));
    else HVX_CUDA(launch_pdl(llm_attn_seq_kernel<false>, sgrid, dim3(32 * a.group * pairs), smem, st, a));
    HVX_LAUNCH_CHECK(e);
    return HVX_OK;
  }
  dim3 grid(splits, rows * c.llm_kv_heads);
  if (splits > 1 && splits <= 16 && !getenv("HVX_NO_CLUSTER_ATTN")) {
    static bool np_set = false;
    if (!np_set) {
      cudaFuncSetAttribute(llm_attn_kernel<true, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(llm_attn_kernel<false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      np_set = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = dim3(32 * a.group); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = splits; at[1].val.clusterDim.y = 1; at[1].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    if (L->kv_f32) HVX_CUDA(cudaLaunchKernelEx(&cfg, llm_attn_kernel<true, true>, a));
    else HVX_CUDA(cudaLaunchKernelEx(&cfg, llm_attn_kernel<false, true>, a));
  } else if (L->kv_f32) HVX_CUDA(launch_pdl(llm_attn_kernel<true, false>, grid, dim3(32 * a.group), 0, st, a));
  else HVX_CUDA(launch_pdl(llm_attn_kernel<false, false>, grid, dim3(32 * a.group), 0, st, a));
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

// out[i] = resid[i] + partial sums of a split-K GEMM, added in split order (deterministic); optionally the split bf16 copy
// [hi | lo] of the result (row length H) for the next GEMM
__global__ void llm_splitk_reduce_kernel(float* __restrict__ out, const float* __restrict__ resid, const float* __restrict__ part, int S,
                                         size_t n4, size_t stride4, __nv_bfloat16* __restrict__ out16, int H) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = reinterpret_cast<const float4*>(resid)[i];
  for (int z = 0; z < S; z++) {
    const float4 p = reinterpret_cast<const float4*>(part)[(size_t)z * stride4 + i];
    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
  }
  reinterpret_cast<float4*>(out)[i] = a;
  if (out16) {
    const size_t e0 = i * 4, row = e0 / H, col = e0 - row * H;
    __nv_bfloat16* o = out16 + row * 2 * H;
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(a.x - __low2float(h0), a.y - __high2float(h0));
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(a.z - __low2float(h1), a.w - __high2float(h1));
    *reinterpret_cast<uint2*>(o + col) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
    *reinterpret_cast<uint2*>(o + H + col) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
  }
}

// split-K reduce of the decode QKV projection fused with its epilogue (bias, RoPE, q out, K / V straight into the cache): a thread
// per (row, even column) pair
__global__ void llm_qkv_reduce_kernel(const float* __restrict__ part, int S, size_t stride, const float* __restrict__ bias, int rows, int N,
                                      LlmQkvEpi q) {
  pdl_trigger();
  pdl_wait();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t pairs = (size_t)rows * (N >> 1);
  if (i >= pairs) return;
  const int row = (int)(i / (N >> 1)), n = 2 * (int)(i - (size_t)row * (N >> 1));
  float2 v = make_float2(bias ? bias[n] : 0.f, bias ? bias[n + 1] : 0.f);
  for (int z = 0; z < S; z++) {
    const float2 p = *reinterpret_cast<const float2*>(part + (size_t)z * stride + (size_t)row * N + n);
    v.x += p.x; v.y += p.y;
  }
  llm_qkv_store(q, row, n, v.x, v.y);
}

// h (rows x H) += A16 (split bf16 [hi | lo], rows x 2K) . W^T  for the skinny N = H projections of the tcgen05 path (o-proj,
// down-proj): 14 output tiles for 148 SMs, so the k-blocks are split over grid.z and the partials are added by
// llm_splitk_reduce_kernel (profiles/r1c_llm_batch32_kernels.txt: these two GEMMs were 34 % of a 128-row decode step)
static hvx_status llm_skinny_resid_gemm(hvx_engine* e, cudaStream_t st, const StepBufs& b, const __nv_bfloat16* a16, const __nv_bfloat16* w,
                                        int rows, int H, int K, float* out = nullptr, const float* resid = nullptr,
                                        __nv_bfloat16* out16 = nullptr, const float* norm_w = nullptr, __nv_bfloat16* norm_x16 = nullptr) {
  // norm_w / norm_x16: the RMSNorm that follows the projection in the layer (its [hi | lo] rows for the next GEMM).  A warp-per-row
  // kernel fusing the split-K reduce with this norm was measured at 27 us against 3.5 + 4.7 us for the two launches: kept separate.
  auto norm_after = [&](const float* h) -> hvx_status {
    if (!norm_w) return HVX_OK;
    HVX_CUDA(launch_pdl_ex(llm_norm16_kernel, dim3(cdiv(rows, 8)), dim3(256), (size_t)0, st, -1, h, norm_w, norm_x16, rows, H, e->cfg.llm_eps));
    HVX_LAUNCH_CHECK(e);
    return HVX_OK;
  };
  static const int want = getenv("HVX_LLM_SPLITK") ? atoi(getenv("HVX_LLM_SPLITK")) : 1;
  if (!out) out = b.h;
  if (!resid) resid = b.h;
  GemmAddr ga; ga.b_kb_mod = K / 64; ga.no_pdl = 1;
  const int nkb = 2 * K / 64;
  // >= 7 k-blocks per CTA, and enough CTAs for every SM: N = H gives only H/64 output tiles
  const int tiles = cdiv(H, 64) * cdiv(rows, 128);
  int S = !want ? 1 : std::min(64, std::max(1, std::min(nkb / 7, cdiv(2 * e->sm_count, tiles))));
  while (S > 1 && (size_t)S * rows * H > b.part_floats) S--;
  if (S <= 1 || (H & 3)) {
    GemmEpi p; p.mode = EPI_F32; p.out = out; p.ldo = H; p.resid = resid;
    if (out16) { p.out2 = out16; p.ldo2 = 2 * H; p.lo_off = H; }
    const hvx_status rc1 = gemm_bf16(e, st, a16, 2 * K, w, K, rows, H, 2 * K, p, &ga);
    return rc1 ? rc1 : norm_after(out);
  }
  ga.split_k = S; ga.split_stride = (size_t)rows * H;
  GemmEpi p; p.mode = EPI_F32; p.out = b.part; p.ldo = H;
  hvx_status rc = gemm_bf16(e, st, a16, 2 * K, w, K, rows, H, 2 * K, p, &ga);
  if (rc) return rc;
  const size_t n4 = (size_t)rows * H / 4;
  HVX_CUDA(launch_pdl_ex(llm_splitk_reduce_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), (size_t)0, st, -1, out, resid, b.part, S, n4, n4, out16, H));
  HVX_LAUNCH_CHECK(e);
  return norm_after(out);
}

// the 24 transformer layers over `rows` row slots of b.h (in place).  rows <= 8: weight-streaming GEMVs;
// otherwise the tcgen05 GEMM with bf16 activations.
static hvx_status llm_layers(hvx_engine* e, cudaStream_t st, LlmState* L, const StepBufs& b, int rows, const SeqState* seqs,
                             int rows_per_seq, int seq0, int pos0, int splits) {
  const hvx_config& c = e->cfg;
  const int H = c.llm_hidden, I = c.llm_inter, NQKV = (c.llm_q_heads + 2 * c.llm_kv_heads) * 64;
  hvx_status rc;
  const bool tc = rows > GEMV_MAX_ROWS;
  for (int l = 0; l < c.llm_layers; l++) {
    const LlmLayer& y = L->layer[l];
    const LlmQkvEpi qe = make_qkv(e, L, l, b.q, seqs, rows_per_seq, seq0, pos0, rows);
    if (!tc) {
      GemvArgs g; g.x = b.h; g.ldx = H; g.W = y.qkv_w; g.bias = y.qkv_b; g.norm_w = y.ln1; g.eps = c.llm_eps;
      g.N = NQKV; g.K = H; g.mode = GEMV_QKV; g.qkv = qe;
      if ((rc = launch_gemv(e, st, g, rows))) return rc;
      if ((rc = launch_attn(e, st, L, l, b, rows, seqs, rows_per_seq, seq0, pos0, splits, false))) return rc;
      GemvArgs o; o.x = b.att; o.ldx = H; o.W = y.o_w; o.N = H; o.K = H; o.out = b.h; o.ldo = H; o.resid = b.h; o.ldr = H;
      if ((rc = launch_gemv(e, st, o, rows))) return rc;
      GemvArgs u; u.x = b.h; u.ldx = H; u.W = y.gu_w; u.norm_w = y.ln2; u.eps = c.llm_eps; u.N = 2 * I; u.K = H;
      u.mode = GEMV_SWIGLU; u.out = b.act; u.ldo = I;
      if ((rc = launch_gemv(e, st, u, rows))) return rc;
      GemvArgs d; d.x = b.act; d.ldx = I; d.W = y.down_w; d.N = H; d.K = I; d.out = b.h; d.ldo = H; d.resid = b.h; d.ldr = H;
      if ((rc = launch_gemv(e, st, d, rows))) return rc;
    } else {
      // activations travel as split bf16 [hi | lo] (K' = 2K against the same weights) -> fp32-accurate products
      GemmAddr gh; gh.b_kb_mod = H / 64; gh.no_pdl = 1;
      GemmAddr gi; gi.b_kb_mod = I / 64; gi.no_pdl = 1;
      if (l == 0) {                                   // later layers: fused into the previous layer's down-proj reduce
        HVX_CUDA(launch_pdl_ex(llm_norm16_kernel, dim3(cdiv(rows, 8)), dim3(256), (size_t)0, st, -1, b.h, y.ln1, b.x16, rows, H, c.llm_eps));
        HVX_LAUNCH_CHECK(e);
      }
      // decode (seqs): the 18 output tiles of the QKV projection leave 130 SMs idle — split-K over grid.z, the epilogue (bias,
      // RoPE, cache write) runs in the reduce kernel.  The prefill keeps the one-pass epilogue GEMM (its accumulation order is
      // the one the C2 token-identity fixture was minted against).
      static const int qkv_splitk = getenv("HVX_LLM_QKV_SPLITK") ? atoi(getenv("HVX_LLM_QKV_SPLITK")) : 4;
      int Sq = seqs && qkv_splitk > 1 ? std::min(qkv_splitk, (2 * H / 64) / 7) : 1;
      while (Sq > 1 && (size_t)Sq * rows * NQKV > b.part_floats) Sq--;
      if (Sq > 1) {
        GemmAddr gq = gh; gq.split_k = Sq; gq.split_stride = (size_t)rows * NQKV;
        GemmEpi p; p.mode = EPI_F32; p.out = b.part; p.ldo = NQKV;
        if ((rc = gemm_bf16(e, st, b.x16, 2 * H, y.qkv_w, H, rows, NQKV, 2 * H, p, &gq))) return rc;
        const size_t pairs = (size_t)rows * (NQKV / 2);
        HVX_CUDA(launch_pdl_ex(llm_qkv_reduce_kernel, dim3((unsigned)((pairs + 255) / 256)), dim3(256), (size_t)0, st, -1, b.part, Sq, (size_t)rows * NQKV, y.qkv_b, rows, NQKV, qe));
        HVX_LAUNCH_CHECK(e);
      } else {
        GemmEpi p; p.mode = EPI_LLM_QKV; p.bias = y.qkv_b; p.llm = qe;
        if ((rc = gemm_bf16(e, st, b.x16, 2 * H, y.qkv_w, H, rows, NQKV, 2 * H, p, &gh))) return rc;
      }
      if ((rc = launch_attn(e, st, L, l, b, rows, seqs, rows_per_seq, seq0, pos0, splits, true))) return rc;
      if ((rc = llm_skinny_resid_gemm(e, st, b, b.att16, y.o_w, rows, H, H, nullptr, nullptr, nullptr, y.ln2, b.x16))) return rc;
      { GemmEpi p; p.mode = EPI_SWIGLU; p.out = b.act16; p.ldo = 2 * I; p.lo_off = I;
        if ((rc = gemm_bf16(e, st, b.x16, 2 * H, y.gu_w, H, rows, 2 * I, 2 * H, p, &gh))) return rc; }
      const float* next_ln1 = l + 1 < c.llm_layers ? L->layer[l + 1].ln1 : nullptr;
      if ((rc = llm_skinny_resid_gemm(e, st, b, b.act16, y.down_w, rows, H, I, nullptr, nullptr, nullptr, next_ln1, b.x16))) return rc;
    }
  }
  return HVX_OK;
}

// final norm of each sequence's last row -> head_k MTP heads -> llm_decoder logits   (:883-888)
static hvx_status llm_heads(hvx_engine* e, cudaStream_t st, LlmState* L, const StepBufs& b, int n_seq, int head_k,
                            int rows_per_seq) {
  const hvx_config& c = e->cfg;
  const int H = c.llm_hidden, MI = c.llm_mtp_inter, V = c.llm_speech_vocab;
  hvx_status rc;
  const bool tc = n_seq > GEMV_MAX_ROWS;
  HVX_CUDA(launch_pdl(llm_last_norm_kernel, dim3(n_seq), dim3(256), 0, st, (const float*)b.h, L->norm, (const SeqState*)L->seqs, rows_per_seq, b.hn,
                      (__nv_bfloat16*)nullptr, H, c.llm_eps));
  HVX_LAUNCH_CHECK(e);
  const size_t sH = (size_t)n_seq * H;
  if (!tc) {
    // v = Wv rmsnorm(hn) + bv ; h1 = hn + Wo v   (attention over one key == v, RoPE at position 0 == identity)
    GemvArgs v; v.x = b.hn; v.ldx = H; v.W = L->m_v_w; v.bias = L->m_v_b; v.norm_w = L->m_ln1; v.eps = c.llm_eps; v.N = H; v.K = H;
    v.out = b.m_v; v.ldo = H; v.sW = (size_t)H * H; v.sBias = H; v.sNorm = H; v.sOut = sH;
    if ((rc = launch_gemv(e, st, v, n_seq, head_k))) return rc;
    GemvArgs o; o.x = b.m_v; o.ldx = H; o.W = L->m_o_w; o.N = H; o.K = H; o.out = b.m_h1; o.ldo = H; o.resid = b.hn; o.ldr = H;
    o.sW = (size_t)H * H; o.sX = sH; o.sOut = sH;
    if ((rc = launch_gemv(e, st, o, n_seq, head_k))) return rc;
    GemvArgs u; u.x = b.m_h1; u.ldx = H; u.W = L->m_gu_w; u.norm_w = L->m_ln2; u.eps = c.llm_eps; u.N = 2 * MI; u.K = H;
    u.mode = GEMV_SWIGLU; u.out = b.m_act; u.ldo = MI; u.sW = (size_t)2 * MI * H; u.sNorm = H; u.sX = sH; u.sOut = (size_t)n_seq * MI;
    if ((rc = launch_gemv(e, st, u, n_seq, head_k))) return rc;
    GemvArgs d; d.x = b.m_act; d.ldx = MI; d.W = L->m_down_w; d.N = H; d.K = MI; d.out = b.m_o; d.ldo = H; d.resid = b.m_h1; d.ldr = H;
    d.sW = (size_t)H * MI; d.sX = (size_t)n_seq * MI; d.sOut = sH; d.sResid = sH;
    if ((rc = launch_gemv(e, st, d, n_seq, head_k))) return rc;
    // logits: one weight matrix for all heads; rows = head_k * n_seq when that fits the GEMV
    if (head_k * n_seq <= GEMV_MAX_ROWS) {
      GemvArgs g; g.x = b.m_o; g.ldx = H; g.W = L->dec_w; g.N = V; g.K = H; g.out = b.logits; g.ldo = V;
      if ((rc = launch_gemv(e, st, g, head_k * n_seq))) return rc;
    } else {
      GemvArgs g; g.x = b.m_o; g.ldx = H; g.W = L->dec_w; g.N = V; g.K = H; g.out = b.logits; g.ldo = V;
      g.sX = sH; g.sOut = (size_t)n_seq * V;
      if ((rc = launch_gemv(e, st, g, n_seq, head_k))) return rc;
    }
  } else {
    GemmAddr gh; gh.b_kb_mod = H / 64; gh.no_pdl = 1;
    GemmAddr gi; gi.b_kb_mod = MI / 64; gi.no_pdl = 1;
    for (int j = 0; j < head_k; j++) {
      HVX_CUDA(launch_pdl_ex(llm_norm16_kernel, dim3(cdiv(n_seq, 8)), dim3(256), (size_t)0, st, -1, b.hn, L->m_ln1 + (size_t)j * H, b.x16, n_seq, H, c.llm_eps));
      HVX_LAUNCH_CHECK(e);
      { GemmEpi p; p.mode = EPI_BF16; p.bias = L->m_v_b + (size_t)j * H; p.out = b.m16; p.ldo = 2 * H; p.lo_off = H;
        if ((rc = gemm_bf16(e, st, b.x16, 2 * H, L->m_v_w + (size_t)j * H * H, H, n_seq, H, 2 * H, p, &gh))) return rc; }
      { GemmEpi p; p.mode = EPI_F32; p.out = b.m_h1 + j * sH; p.ldo = H; p.resid = b.hn;
        if ((rc = gemm_bf16(e, st, b.m16, 2 * H, L->m_o_w + (size_t)j * H * H, H, n_seq, H, 2 * H, p, &gh))) return rc; }
      HVX_CUDA(launch_pdl_ex(llm_norm16_kernel, dim3(cdiv(n_seq, 8)), dim3(256), (size_t)0, st, -1, b.m_h1 + j * sH, L->m_ln2 + (size_t)j * H, b.x16, n_seq, H, c.llm_eps));
      HVX_LAUNCH_CHECK(e);
      { GemmEpi p; p.mode = EPI_SWIGLU; p.out = b.act16; p.ldo = 2 * MI; p.lo_off = MI;
        if ((rc = gemm_bf16(e, st, b.x16, 2 * H, L->m_gu_w + (size_t)j * 2 * MI * H, H, n_seq, 2 * MI, 2 * H, p, &gh))) return rc; }
      // down-proj: K = mtp_inter (22016 -> 688 k-blocks) against 14 output tiles: split over k like the layers' down-proj
      if ((rc = llm_skinny_resid_gemm(e, st, b, b.act16, L->m_down_w + (size_t)j * H * MI, n_seq, H, MI, b.m_o + j * sH, b.m_h1 + j * sH, b.m16)))
        return rc;
      { GemmEpi p; p.mode = EPI_F32; p.out = b.logits + (size_t)j * n_seq * V; p.ldo = V;
        if ((rc = gemm_bf16(e, st, b.m16, 2 * H, L->dec_w, H, n_seq, V, 2 * H, p, &gh))) return rc; }
    }
  }
  return HVX_OK;
}

static hvx_status launch_sampler(hvx_engine* e, cudaStream_t st, SampArgs a, int n_seq) {
  HVX_CHECK(a.vocab <= SAMP_THREADS * SAMP_PER, HVX_ERR_UNSUPPORTED, "sampler: vocab %d too large", a.vocab);
  const size_t smem = (size_t)a.vocab * sizeof(float);
  const size_t tab_bytes = (size_t)n_seq * a.head_k * (SAMP_TAB + a.vocab) * sizeof(float);
  const size_t cnt_off = (tab_bytes + 255) & ~(size_t)255;
  const bool grew = cnt_off + (size_t)n_seq * 4 > e->samp_ws.bytes;
  uint8_t* w = (uint8_t*)e->samp_ws.get(cnt_off + (size_t)n_seq * 4 + 256);
  HVX_CHECK(w, HVX_ERR_CUDA, "sampler: scratch allocation failed");
  a.tables = (float*)w;
  a.arrive = (int*)(w + cnt_off);
  // arrival counters start at zero and every launch leaves them at zero; (re)initialise when the scratch moved
  if (grew || e->samp_arrive != (void*)a.arrive) {
    HVX_CUDA(cudaMemsetAsync(a.arrive, 0, (size_t)n_seq * 4, st));
    e->samp_arrive = a.arrive;
  }
  static bool attr = false;
  if (!attr) { HVX_CUDA(cudaFuncSetAttribute(llm_sampler_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
  HVX_CUDA(launch_pdl(llm_sampler_kernel, dim3(n_seq, a.head_k), dim3(SAMP_THREADS), smem, st, a));
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

// ---- fused persistent decode step (llm_fused.cuh): one sequence, head_k <= 4
struct FusedPlan { bool ok = false; int R = 0, n_slots = 0, x_bytes = 0; size_t smem = 0; };

static FusedPlan fused_plan(hvx_engine* e, int n_seq, int head_k, bool force = false) {
  FusedPlan f;
  const hvx_config& c = e->cfg;
  // Opt-in (HVX_FUSED_DECODE=1): numerically equivalent to the kernel-per-op step (tests/test_llm_gpu.py), but at round 1 it
  // is slower on B200 (1739 vs 783 us per step at ctx 800, scripts/bench_llm_kernels.py: 120 grid barriers and the per-phase
  // activation re-staging cost more than the PDL-chained launches they replace)
  const char* env = getenv("HVX_FUSED_DECODE");
  if (!force && !(env && atoi(env) != 0)) return f;
  if (n_seq != 1 || head_k > 4) return f;
  const int H = c.llm_hidden, I = c.llm_inter, MI = c.llm_mtp_inter, G = e->sm_count;
  f.R = head_k <= 1 ? 1 : head_k <= 2 ? 2 : 4;
  const int group = c.llm_q_heads / c.llm_kv_heads;
  if (f.R * group > 2 * fused::NW || H % 8 || I % 8 || MI % 8 || H > 2048) return f;
  // every phase: a unit (two rows of one k-segment) fits a slot, k-split partial sums fit the scratch
  auto unit_ok = [&](int N, int K, int nb, int RB) {
    const int ks = fused::pick_ksplit(K), seg = K / ks;
    if (seg * 4 > fused::SLOT || seg % 8) return false;
    const int pairs = cdiv(((N + 1) / 2) * nb, G);
    return ks == 1 || pairs * ks * 2 * RB <= fused::RED_FLOATS;
  };
  const int NQKV = (c.llm_q_heads + 2 * c.llm_kv_heads) * 64;
  if (!unit_ok(NQKV, H, 1, f.R) || !unit_ok(H, H, 1, f.R) || !unit_ok(2 * I, H, 1, f.R) || !unit_ok(H, I, 1, f.R) ||
      !unit_ok(H, H, head_k, 1) || !unit_ok(2 * MI, H, head_k, 1) || !unit_ok(H, MI, head_k, 1) || !unit_ok(c.llm_speech_vocab, H, 1, f.R))
    return f;
  size_t xb = (size_t)f.R * std::max(H, I) * 4;
  xb = std::max(xb, (size_t)head_k * H * 4);
  xb = std::max(xb, (size_t)2 * fused::ATT_CHUNK * (c.llm_kv_f32 ? ATT_LD32 * 4 : ATT_LD * 2));
  xb = (xb + 127) & ~(size_t)127;
  const size_t budget = 216 * 1024;        // + 9 KB static shared memory <= 227 KB
  if (xb + (size_t)H * 4 + 3 * fused::SLOT > budget) return f;
  f.n_slots = (int)std::min((size_t)fused::MAXNS, (budget - xb - (size_t)H * 4) / fused::SLOT);
  f.x_bytes = (int)xb;
  f.smem = xb + (size_t)f.n_slots * fused::SLOT + (size_t)H * 4;
  f.ok = true;
  return f;
}

static hvx_status launch_fused_step(hvx_engine* e, cudaStream_t st, LlmState* L, const StepBufs& b, int head_k, const FusedPlan& f,
                                    int max_phases = 1 << 30, unsigned long long* dbg = nullptr) {
  const hvx_config& c = e->cfg;
  fused::Args a;
  a.layers = L->layers_dev; a.n_layers = c.llm_layers; a.H = c.llm_hidden; a.I = c.llm_inter; a.q_heads = c.llm_q_heads;
  a.kv_heads = c.llm_kv_heads; a.MI = c.llm_mtp_inter; a.V = c.llm_speech_vocab; a.head_k = head_k; a.max_ctx = c.llm_max_ctx;
  a.eps = c.llm_eps; a.scale = 1.0f / sqrtf((float)c.llm_head_dim);
  a.norm = L->norm; a.m_v_w = L->m_v_w; a.m_o_w = L->m_o_w; a.m_gu_w = L->m_gu_w; a.m_down_w = L->m_down_w; a.dec_w = L->dec_w;
  a.m_v_b = L->m_v_b; a.m_ln1 = L->m_ln1; a.m_ln2 = L->m_ln2;
  a.kc = L->kc; a.vc = L->vc; a.layer_stride = L->layer_stride; a.seq_stride = L->seq_stride; a.kv_f32 = L->kv_f32;
  a.inv_freq = L->inv_freq; a.seqs = L->seqs;
  a.h = b.h; a.q = b.q; a.att = b.att; a.act = b.act; a.part = b.part;
  a.m_v = b.m_v; a.m_h1 = b.m_h1; a.m_act = b.m_act; a.m_o = b.m_o; a.logits = b.logits;
  a.bar = L->fused_bar; a.abort_flag = L->fused_abort; a.n_slots = f.n_slots; a.x_bytes = f.x_bytes; a.max_phases = max_phases; a.dbg = dbg;
  { const char* pc = getenv("HVX_FUSED_PACE"); a.pace_clk = pc ? atoi(pc) : 0; }
  static bool attr = false;
  if (!attr) {
    HVX_CUDA(cudaFuncSetAttribute(fused::llm_fused_step_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    HVX_CUDA(cudaFuncSetAttribute(fused::llm_fused_step_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    HVX_CUDA(cudaFuncSetAttribute(fused::llm_fused_step_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024));
    attr = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(e->sm_count, 1, 1);
  cfg.blockDim = dim3(fused::THREADS, 1, 1);
  cfg.dynamicSmemBytes = f.smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;        // all CTAs co-resident (grid barriers), or the launch fails
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  switch (f.R) {
    case 1: HVX_CUDA(cudaLaunchKernelEx(&cfg, fused::llm_fused_step_kernel<1>, a)); break;
    case 2: HVX_CUDA(cudaLaunchKernelEx(&cfg, fused::llm_fused_step_kernel<2>, a)); break;
    default: HVX_CUDA(cudaLaunchKernelEx(&cfg, fused::llm_fused_step_kernel<4>, a)); break;
  }
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

static int attn_splits(int sm, int rows, int kv_heads) {
  int s = (2 * sm) / std::max(1, rows * kv_heads);
  const char* cap = getenv("HVX_ATTN_SPLITS");
  return std::max(1, std::min(s, cap ? atoi(cap) : 16));   // the splits of a row form one cluster (16 = non-portable maximum) and merge through DSMEM
}

}  // namespace hvx

using namespace hvx;

extern "C" hvx_status hvx_llm_begin(hvx_engine* e, int seq, const int32_t* text_ids, int n_text_total, int n_text_new,
                                    const int32_t* prompt_speech, int n_prompt_speech, float min_ratio, float max_ratio) {
  HVX_CHECK(e && e->llm, HVX_ERR_STATE, "llm stage not finalized");
  HVX_LOCK(e, HVX_STAGE_LLM);
  HVX_CHECK(seq >= 0 && seq < e->cfg.llm_max_seqs, HVX_ERR_ARG, "llm_begin: sequence slot %d out of range", seq);
  HVX_CHECK(text_ids && n_text_total >= n_text_new && n_text_new >= 0 && (n_prompt_speech == 0 || prompt_speech), HVX_ERR_ARG,
            "llm_begin: bad arguments");
  HVX_CHECK(2 + n_text_total + n_prompt_speech < e->cfg.llm_max_ctx, HVX_ERR_ARG, "llm_begin: prompt longer than max_ctx");
  SeqDesc& d = e->llm->desc[seq];
  d.text = text_ids; d.n_text_total = n_text_total; d.n_text_new = n_text_new; d.pspeech = prompt_speech; d.n_ps = n_prompt_speech;
  d.min_ratio = min_ratio; d.max_ratio = max_ratio; d.begun = true;
  return HVX_OK;
}

// prefill of every begun sequence: the prompt rows go through the tensor-core path, the last hidden row lands in
// row slot seq*head_k + head_k-1 so the head/sampler kernels treat prefill and decode alike
static hvx_status llm_prefill(hvx_engine* e, cudaStream_t st, LlmState* L, int n_seq, int head_k, const StepBufs& b) {
  const hvx_config& c = e->cfg;
  const int H = c.llm_hidden;
  const int sos = c.llm_speech_vocab - 200, task = sos + 2;
  hvx_status rc;
  std::vector<SeqState> hs(n_seq);
  int max_rows = 0;
  for (int s = 0; s < n_seq; s++) max_rows = std::max(max_rows, 2 + L->desc[s].n_text_total + L->desc[s].n_ps);
  StepBufs pb;
  if ((rc = llm_bufs(e, L, L->pws, max_rows, 1, 1, 1, &pb, st))) return rc;
  for (int s = 0; s < n_seq; s++) {
    const SeqDesc& d = L->desc[s];
    HVX_CHECK(d.begun, HVX_ERR_STATE, "llm_generate: sequence %d was not begun", s);
    const int rows = 2 + d.n_text_total + d.n_ps;
    SeqState& x = hs[s];
    memset(&x, 0, sizeof(x));
    x.ctx = rows; x.ctx_add = 0; x.n_new = head_k;
    x.min_len = (int)((float)d.n_text_new * d.min_ratio);        // int((text_len - prompt_text_len) * ratio)  (:955-956)
    x.max_len = (int)((float)d.n_text_new * d.max_ratio);
    x.done = x.max_len <= 0;
    llm_prompt_rows_kernel<<<rows, 128, 0, st>>>(d.text, d.n_text_total, d.pspeech, d.n_ps, L->embed, L->semb, pb.h, H, sos, task,
                                                c.llm_text_vocab, c.llm_speech_vocab);
    HVX_LAUNCH_CHECK(e);
    if ((rc = llm_layers(e, st, L, pb, rows, nullptr, 1, s, 0, 1))) return rc;
    HVX_CUDA(cudaMemcpyAsync(b.h + ((size_t)s * head_k + head_k - 1) * H, pb.h + (size_t)(rows - 1) * H, sizeof(float) * H,
                             cudaMemcpyDeviceToDevice, st));
  }
  int alive = 0;
  for (int s = 0; s < n_seq; s++) alive += !hs[s].done;
  HVX_CUDA(cudaMemcpyAsync(L->seqs, hs.data(), sizeof(SeqState) * n_seq, cudaMemcpyHostToDevice, st));
  HVX_CUDA(cudaMemcpyAsync(L->n_active, &alive, sizeof(int), cudaMemcpyHostToDevice, st));
  HVX_CUDA(cudaStreamSynchronize(st));                             // hs / alive are stack memory
  return HVX_OK;
}

extern "C" hvx_status hvx_llm_generate(hvx_engine* e, int n_seq, int head_k, const hvx_sampler* sp, const float* u_dev,
                                       int u_stride, int32_t* out_tokens, int max_out, int32_t* out_counts, void* stream) {
  return hvx::llm_generate_progress(e, n_seq, head_k, sp, u_dev, u_stride, out_tokens, max_out, out_counts, stream, nullptr, nullptr);
}

// hvx_llm_generate with a progress hook: after every batch of `poll` decode steps (the decode stream is drained at that point)
// progress(ctx, n_seq, done[], n_out[]) tells the caller which sequences have stopped and how many tokens each has emitted —
// hvx_synthesize_host starts the flow + vocoder of finished utterances on the caller's stream while the rest keep decoding.
hvx_status hvx::llm_generate_progress(hvx_engine* e, int n_seq, int head_k, const hvx_sampler* sp, const float* u_dev, int u_stride,
                                      int32_t* out_tokens, int max_out, int32_t* out_counts, void* stream, llm_progress_fn progress,
                                      void* progress_ctx) {
  HVX_CHECK(e && e->llm, HVX_ERR_STATE, "llm stage not finalized");
  HVX_LOCK(e, HVX_STAGE_LLM);
  HVX_CHECK(sp && u_dev && out_tokens && out_counts, HVX_ERR_ARG, "llm_generate: null argument");
  const hvx_config& c = e->cfg;
  HVX_CHECK(n_seq >= 1 && n_seq <= c.llm_max_seqs, HVX_ERR_ARG, "llm_generate: n_seq=%d exceeds max_seqs=%d", n_seq, c.llm_max_seqs);
  head_k = std::max(1, std::min(head_k, c.llm_mtp_heads));         // llm_multi_head_v3.py:866-868
  HVX_CHECK(sp->top_k >= 1 && sp->top_k <= SAMP_MAXK, HVX_ERR_UNSUPPORTED, "sampler: top_k=%d outside [1,%d]", sp->top_k, SAMP_MAXK);
  LlmState* L = e->llm;
  cudaStream_t user = (cudaStream_t)stream, st = L->own;
  e->llm_cancel.store(0);
  HVX_CUDA(cudaEventRecord(L->ev_in, user));
  HVX_CUDA(cudaStreamWaitEvent(st, L->ev_in, 0));
  const int rows = n_seq * head_k;
  const int splits = attn_splits(e->sm_count, rows, c.llm_kv_heads);
  StepBufs b;
  hvx_status rc;
  if ((rc = llm_bufs(e, L, L->ws, rows, n_seq, head_k, splits, &b, st))) return rc;
  HVX_CUDA(cudaMemsetAsync(out_counts, 0, sizeof(int32_t) * n_seq, st));
  struct GemmOff { hvx_engine* e; GemmOff(hvx_engine* x) : e(x) { e->prof_gemm_off++; } ~GemmOff() { e->prof_gemm_off--; } } gemm_off(e);
  { double rows = 0;
    for (int s = 0; s < n_seq; s++) rows += 2 + L->desc[s].n_text_total + L->desc[s].n_ps;
    ProfScope ps(&e->prof, st, PROF_LLM_PREFILL, rows);
    if ((rc = llm_prefill(e, st, L, n_seq, head_k, b))) return rc; }

  SampArgs sa;
  sa.logits = b.logits; sa.n_seq = n_seq; sa.head_k = head_k; sa.vocab = c.llm_speech_vocab; sa.stop_from = c.llm_speech_vocab - 200;
  sa.top_p = sp->top_p; sa.top_k = sp->top_k; sa.win_size = sp->win_size; sa.rep_thr = (double)sp->win_size * sp->tau_r;
  sa.u = u_dev; sa.u_stride = u_stride; sa.seqs = L->seqs; sa.out_tokens = out_tokens; sa.max_out = max_out; sa.out_counts = out_counts;
  sa.semb = L->semb; sa.h = b.h; sa.H = c.llm_hidden; sa.n_active = L->n_active; sa.max_ctx = c.llm_max_ctx;
  const FusedPlan fp = fused_plan(e, n_seq, head_k);
  if (fp.ok) { sa.fused_bar = L->fused_bar; HVX_CUDA(cudaMemsetAsync(L->fused_abort, 0, sizeof(int), st)); }

  // first sample straight from the prefill's last hidden state
  if ((rc = llm_heads(e, st, L, b, n_seq, head_k, head_k))) return rc;
  if ((rc = launch_sampler(e, st, sa, n_seq))) return rc;

  // every running step emits exactly head_k tokens per live sequence (or stops it), the first sample is already out
  int max_steps = 0;
  for (int s = 0; s < n_seq; s++) {
    const int ml = std::min((int)((float)L->desc[s].n_text_new * L->desc[s].max_ratio), max_out);
    max_steps = std::max(max_steps, cdiv(std::max(ml - head_k, 0), head_k));
  }
  const int poll = 16;
  std::vector<SeqState> prog_seqs(progress ? n_seq : 0);
  std::vector<int> prog_done(progress ? n_seq : 0), prog_out(progress ? n_seq : 0);
  if (fp.ok) {
    // one sequence: the whole step is ONE persistent cooperative kernel (llm_fused.cuh) + the sampler
    for (int step = 0; step < max_steps;) {
      const int n = std::min(poll, max_steps - step);
      for (int i = 0; i < n; i++) {
        if ((rc = launch_fused_step(e, st, L, b, head_k, fp))) return rc;
        if ((rc = launch_sampler(e, st, sa, n_seq))) return rc;
      }
      step += n;
      HVX_CUDA(cudaMemcpyAsync(L->n_active_host, L->n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      HVX_CUDA(cudaStreamSynchronize(st));
      if (*L->n_active_host <= 0) break;
    }
    int aborted = 0;
    HVX_CUDA(cudaMemcpyAsync(&aborted, L->fused_abort, sizeof(int), cudaMemcpyDeviceToHost, st));
    HVX_CUDA(cudaStreamSynchronize(st));
    HVX_CHECK(aborted == 0, HVX_ERR_CUDA, "llm: fused decode step timed out on a barrier (code %d)", aborted);
  } else {
  static const bool no_graph = getenv("HVX_LLM_NO_GRAPH") != nullptr;      // diagnostics: eager steps (HVX_GEMM_TIMELINE needs them)
  if (no_graph) {
    for (int step = 0; step < max_steps;) {
      const int n = std::min(poll, max_steps - step);
      for (int i = 0; i < n; i++) {
        if ((rc = llm_layers(e, st, L, b, rows, L->seqs, head_k, 0, 0, splits))) return rc;
        if ((rc = llm_heads(e, st, L, b, n_seq, head_k, head_k))) return rc;
        if ((rc = launch_sampler(e, st, sa, n_seq))) return rc;
      }
      step += n;
      HVX_CUDA(cudaMemcpyAsync(L->n_active_host, L->n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
      HVX_CUDA(cudaStreamSynchronize(st));
      if (*L->n_active_host <= 0 || e->llm_cancel.load()) break;
    }
  } else {
  // one decode step = fixed launch sequence -> CUDA graph (re-captured when shapes or buffers change)
  const int key[4] = {n_seq, head_k, (int)((uintptr_t)b.h >> 8), (int)((uintptr_t)out_tokens >> 4) ^ (int)((uintptr_t)u_dev >> 4) ^ (sp->top_k << 20) ^ sp->win_size};
  const bool same = L->graph && !memcmp(key, L->graph_key, sizeof(key)) && !memcmp(&sa, &L->graph_samp, sizeof(sa));
  if (!same) {
    if (L->graph) { cudaGraphExecDestroy(L->graph); L->graph = nullptr; }
    cudaGraph_t g;
    const int64_t l0 = e->launches;
    HVX_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
    rc = llm_layers(e, st, L, b, rows, L->seqs, head_k, 0, 0, splits);
    if (!rc) rc = llm_heads(e, st, L, b, n_seq, head_k, head_k);
    if (!rc) rc = launch_sampler(e, st, sa, n_seq);
    cudaError_t ce = cudaStreamEndCapture(st, &g);
    if (rc) return rc;
    HVX_CHECK(ce == cudaSuccess, HVX_ERR_CUDA, "llm: graph capture failed: %s", cudaGetErrorString(ce));
    ce = cudaGraphInstantiate(&L->graph, g, 0);
    cudaGraphDestroy(g);
    HVX_CHECK(ce == cudaSuccess, HVX_ERR_CUDA, "llm: graph instantiate failed: %s", cudaGetErrorString(ce));
    memcpy(L->graph_key, key, sizeof(key));
    L->graph_samp = sa;
    e->graph_launches = e->launches - l0;
    e->launches -= e->graph_launches;
  }
  for (int step = 0; step < max_steps;) {
    const int n = std::min(poll, max_steps - step);
    { ProfScope ps(&e->prof, st, PROF_LLM_STEP, (double)n);
      for (int i = 0; i < n; i++) HVX_CUDA(cudaGraphLaunch(L->graph, st)); }
    e->launches += (int64_t)n * e->graph_launches;
    step += n;
    HVX_CUDA(cudaMemcpyAsync(L->n_active_host, L->n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (progress) HVX_CUDA(cudaMemcpyAsync(prog_seqs.data(), L->seqs, sizeof(SeqState) * n_seq, cudaMemcpyDeviceToHost, st));
    HVX_CUDA(cudaStreamSynchronize(st));
    if (progress && *L->n_active_host > 0) {
      for (int s2 = 0; s2 < n_seq; s2++) { prog_done[s2] = prog_seqs[s2].done; prog_out[s2] = prog_seqs[s2].n_out; }
      if ((rc = progress(progress_ctx, n_seq, prog_done.data(), prog_out.data()))) return rc;
    }
    if (*L->n_active_host <= 0 || e->llm_cancel.load()) break;
  }
  }
  }
  // surface sampler failures the way the reference raises (llm_multi_head_v3.py:165)
  std::vector<SeqState> hs(n_seq);
  HVX_CUDA(cudaMemcpyAsync(hs.data(), L->seqs, sizeof(SeqState) * n_seq, cudaMemcpyDeviceToHost, st));
  HVX_CUDA(cudaStreamSynchronize(st));
  HVX_CUDA(cudaEventRecord(L->ev_out, st));
  HVX_CUDA(cudaStreamWaitEvent(user, L->ev_out, 0));
  if (progress) {                       // final report: every sequence has stopped (or the run was cancelled: n_out so far)
    for (int s2 = 0; s2 < n_seq; s2++) { prog_done[s2] = 1; prog_out[s2] = hs[s2].n_out; }
    if ((rc = progress(progress_ctx, n_seq, prog_done.data(), prog_out.data()))) return rc;
  }
  for (int s = 0; s < n_seq; s++) {
    L->desc[s].begun = false;
    HVX_CHECK(hs[s].status != 1, HVX_ERR_STATE, "sampling reaches max_trials 100 and still get eos when ignore_eos is True (sequence %d)", s);
    HVX_CHECK(hs[s].status != 2, HVX_ERR_ARG, "llm_generate: uniform stream of sequence %d exhausted (u_stride=%d)", s, u_stride);
  }
  return HVX_OK;
}

// Asks a hvx_llm_generate running on another host thread to return after its current batch of decode steps (<= 16) with the
// tokens emitted so far.  Deliberately takes no stage lock.
extern "C" hvx_status hvx_llm_cancel(hvx_engine* e) {
  HVX_CHECK(e, HVX_ERR_ARG, "llm_cancel: null engine");
  e->llm_cancel.store(1);
  return HVX_OK;
}

// Teacher-forced probe: prefill sequence slot `seq` and return the final-normed last hidden state and the
// log-softmax of every MTP head on it (llm_multi_head_v3.py:883-888).
extern "C" hvx_status hvx_llm_probe(hvx_engine* e, int seq, float* last_hidden, float* head_logp, void* stream) {
  HVX_CHECK(e && e->llm, HVX_ERR_STATE, "llm stage not finalized");
  HVX_LOCK(e, HVX_STAGE_LLM);
  HVX_CHECK(seq == 0, HVX_ERR_UNSUPPORTED, "llm_probe: only sequence slot 0 is supported");
  const hvx_config& c = e->cfg;
  LlmState* L = e->llm;
  cudaStream_t st = (cudaStream_t)stream;
  const int K = c.llm_mtp_heads, V = c.llm_speech_vocab;
  StepBufs b;
  hvx_status rc;
  if ((rc = llm_bufs(e, L, L->ws, K, 1, K, 1, &b, st))) return rc;
  if ((rc = llm_prefill(e, st, L, 1, K, b))) return rc;
  if ((rc = llm_heads(e, st, L, b, 1, K, K))) return rc;
  if (last_hidden) HVX_CUDA(cudaMemcpyAsync(last_hidden, b.hn, sizeof(float) * c.llm_hidden, cudaMemcpyDeviceToDevice, st));
  if (head_logp) {
    // log-softmax only: run the sampler kernel in standalone mode with logp_out and discard the ids
    SampArgs sa;
    sa.logits = b.logits; sa.n_seq = 1; sa.head_k = K; sa.vocab = V; sa.stop_from = V - 200; sa.top_k = 1; sa.top_p = 0.0;
    sa.win_size = 1; sa.rep_thr = 1e9; sa.u = b.hn; sa.u_stride = c.llm_hidden; sa.logp_out = head_logp; sa.min_len_override = 0;
    if ((rc = launch_sampler(e, st, sa, 1))) return rc;
  }
  L->desc[0].begun = false;
  if (L->graph) { cudaGraphExecDestroy(L->graph); L->graph = nullptr; }
  return HVX_OK;
}

extern "C" hvx_status hvx_sample(hvx_engine* e, const float* logp, int n_heads, const int32_t* history, int n_history, int min_len,
                                 const hvx_sampler* sp, const float* u, int n_u, int32_t* out_ids, int32_t* u_used, void* stream) {
  HVX_CHECK(e && logp && sp && u && out_ids, HVX_ERR_ARG, "hvx_sample: null argument");
  HVX_CHECK(n_heads >= 1 && n_heads <= LLM_MAX_HEADS, HVX_ERR_ARG, "hvx_sample: n_heads out of range");
  HVX_CHECK(sp->top_k >= 1 && sp->top_k <= SAMP_MAXK, HVX_ERR_UNSUPPORTED, "sampler: top_k=%d outside [1,%d]", sp->top_k, SAMP_MAXK);
  const hvx_config& c = e->cfg;
  SampArgs sa;
  sa.logits = logp; sa.n_seq = 1; sa.head_k = n_heads; sa.vocab = c.llm_speech_vocab; sa.stop_from = c.llm_speech_vocab - 200;
  sa.input_is_logp = 1;
  sa.top_p = sp->top_p; sa.top_k = sp->top_k; sa.win_size = sp->win_size; sa.rep_thr = (double)sp->win_size * sp->tau_r;
  sa.u = u; sa.u_stride = n_u; sa.history = history; sa.n_history = n_history; sa.min_len_override = min_len;
  sa.ids_out = out_ids; sa.u_used = u_used;
  return launch_sampler(e, (cudaStream_t)stream, sa, 1);
}


__global__ void llm_debug_rows_kernel(const __nv_bfloat16* __restrict__ semb, float* __restrict__ h, int H, int vocab) {
  const int r = blockIdx.x, id = (7 + 13 * r) % vocab;
  for (int i = threadIdx.x; i < H; i += blockDim.x) h[(size_t)r * H + i] = __bfloat162float(semb[(size_t)id * H + i]);
}

// Parity probe for the two decode-step implementations (tests/test_llm_gpu.py): one step of sequence slot 0 at context
// length `ctx` over whatever the KV cache holds, rows = speech embeddings of fixed token ids, through either the
// kernel-per-op path (fused = 0) or the persistent fused kernel (fused = 1).  n_layers > 0 stops after that many layers
// (heads skipped).  Outputs: residual stream rows [head_k][hidden] and, after a full step, logits [head_k][vocab].
extern "C" hvx_status hvx_llm_debug_step(hvx_engine* e, int head_k, int ctx, int use_fused, int n_layers, float* h_out_dev,
                                         float* logits_out_dev) {
  HVX_CHECK(e && e->llm && h_out_dev, HVX_ERR_ARG, "llm_debug_step: bad argument");
  HVX_LOCK(e, HVX_STAGE_LLM);
  hvx_config& c = e->cfg;
  HVX_CHECK(head_k >= 1 && head_k <= c.llm_mtp_heads && ctx >= 1 && ctx + head_k < c.llm_max_ctx, HVX_ERR_ARG, "llm_debug_step: bad shape");
  LlmState* L = e->llm;
  cudaStream_t st = L->own;
  const int rows = head_k, H = c.llm_hidden, V = c.llm_speech_vocab;
  const int splits = attn_splits(e->sm_count, rows, c.llm_kv_heads);
  StepBufs b;
  hvx_status rc;
  if ((rc = llm_bufs(e, L, L->ws, rows, 1, head_k, splits, &b, st))) return rc;
  SeqState x;
  memset(&x, 0, sizeof(x));
  x.ctx = ctx; x.n_new = head_k; x.max_len = 1 << 30;
  HVX_CUDA(cudaMemcpyAsync(L->seqs, &x, sizeof(x), cudaMemcpyHostToDevice, st));
  HVX_CUDA(cudaStreamSynchronize(st));
  llm_debug_rows_kernel<<<rows, 128, 0, st>>>(L->semb, b.h, H, V);
  HVX_LAUNCH_CHECK(e);
  const int full_layers = c.llm_layers;
  const bool partial = n_layers > 0 && n_layers < full_layers;
  if (partial) c.llm_layers = n_layers;
  if (use_fused) {
    const FusedPlan fp = fused_plan(e, 1, head_k, true);
    if (!fp.ok) { c.llm_layers = full_layers; HVX_CHECK(false, HVX_ERR_UNSUPPORTED, "llm_debug_step: the fused step does not cover this shape"); }
    cudaMemsetAsync(L->fused_bar, 0, sizeof(unsigned long long), st);
    cudaMemsetAsync(L->fused_abort, 0, sizeof(int), st);
    rc = launch_fused_step(e, st, L, b, head_k, fp, partial ? 5 * n_layers : (1 << 30));
  } else {
    rc = llm_layers(e, st, L, b, rows, L->seqs, head_k, 0, 0, splits);
    if (!rc && !partial) rc = llm_heads(e, st, L, b, 1, head_k, head_k);
  }
  c.llm_layers = full_layers;
  if (rc) return rc;
  HVX_CUDA(cudaMemcpyAsync(h_out_dev, b.h, sizeof(float) * rows * H, cudaMemcpyDeviceToDevice, st));
  if (logits_out_dev && !partial) HVX_CUDA(cudaMemcpyAsync(logits_out_dev, b.logits, sizeof(float) * rows * V, cudaMemcpyDeviceToDevice, st));
  int aborted = 0;
  HVX_CUDA(cudaMemcpyAsync(&aborted, L->fused_abort, sizeof(int), cudaMemcpyDeviceToHost, st));
  HVX_CUDA(cudaStreamSynchronize(st));
  HVX_CHECK(aborted == 0, HVX_ERR_CUDA, "llm_debug_step: fused decode step timed out on a barrier (code %d)", aborted);
  if (L->graph) { cudaGraphExecDestroy(L->graph); L->graph = nullptr; }
  return HVX_OK;
}

// Kernel-class timing for bench.py's roofline: runs one class of decode-step kernels back to back on the engine's own
// stream (same launch path as the decode graph: PDL, persistent grids) over the real weights, CUDA-event timed.
//   which: 0 whole step without sampler, 1 qkv gemv, 2 attention, 3 o-proj gemv, 4 gate-up gemv, 5 down gemv, 6 MTP heads + logits,
//          7 whole step without sampler as the persistent fused kernel (llm_fused.cuh)
// ms_out[0] = milliseconds per repetition (one repetition = all layers' kernels of that class)
extern "C" hvx_status hvx_llm_bench_kernels(hvx_engine* e, int n_seq, int head_k, int ctx, int which, int reps, float* ms_out) {
  HVX_CHECK(e && e->llm && ms_out && reps >= 1, HVX_ERR_ARG, "llm_bench_kernels: bad argument");
  HVX_LOCK(e, HVX_STAGE_LLM);
  const hvx_config& c = e->cfg;
  HVX_CHECK(n_seq >= 1 && n_seq <= c.llm_max_seqs && n_seq * head_k <= GEMV_MAX_ROWS && ctx + head_k < c.llm_max_ctx, HVX_ERR_ARG, "llm_bench_kernels: bad shape");
  LlmState* L = e->llm;
  cudaStream_t st = L->own;
  const int rows = n_seq * head_k, H = c.llm_hidden, I = c.llm_inter, NQKV = (c.llm_q_heads + 2 * c.llm_kv_heads) * 64;
  const int splits = attn_splits(e->sm_count, rows, c.llm_kv_heads);
  StepBufs b;
  hvx_status rc;
  if ((rc = llm_bufs(e, L, L->ws, rows, n_seq, head_k, splits, &b, st))) return rc;
  std::vector<SeqState> hs(n_seq);
  for (auto& x : hs) { memset(&x, 0, sizeof(x)); x.ctx = ctx; x.n_new = head_k; x.ctx_add = 0; x.max_len = 1 << 30; }
  HVX_CUDA(cudaMemcpyAsync(L->seqs, hs.data(), sizeof(SeqState) * n_seq, cudaMemcpyHostToDevice, st));
  HVX_CUDA(cudaStreamSynchronize(st));
  cudaEvent_t e0, e1;
  HVX_CUDA(cudaEventCreate(&e0)); HVX_CUDA(cudaEventCreate(&e1));
  unsigned long long* dbg_dev = nullptr;
  if (which == 4 && getenv("HVX_GEMV_TIMELINE")) { HVX_CUDA(cudaMalloc(&dbg_dev, 64 * 8 * 8)); HVX_CUDA(cudaMemset(dbg_dev, 0, 64 * 8 * 8)); }
  unsigned long long* fdbg = nullptr;
  if (which == 7 && getenv("HVX_FUSED_TIMELINE")) { HVX_CUDA(cudaMalloc(&fdbg, 1024 * 8)); HVX_CUDA(cudaMemset(fdbg, 0, 1024 * 8)); }
  auto body = [&]() -> hvx_status {
    if (which == 0) {
      if ((rc = llm_layers(e, st, L, b, rows, L->seqs, head_k, 0, 0, splits))) return rc;
      return llm_heads(e, st, L, b, n_seq, head_k, head_k);
    }
    if (which == 6) return llm_heads(e, st, L, b, n_seq, head_k, head_k);
    if (which == 7) {                                   // the whole step as the persistent fused kernel
      const FusedPlan fp = fused_plan(e, n_seq, head_k, true);
      HVX_CHECK(fp.ok, HVX_ERR_UNSUPPORTED, "llm_bench_kernels: the fused step does not cover this shape");
      HVX_CUDA(cudaMemsetAsync(L->fused_bar, 0, sizeof(unsigned long long), st));
      return launch_fused_step(e, st, L, b, head_k, fp, 1 << 30, fdbg);
    }
    if (which == 8) {                                   // 100 back-to-back grid barriers of the fused kernel's grid
      const FusedPlan fp = fused_plan(e, n_seq, head_k, true);
      HVX_CHECK(fp.ok, HVX_ERR_UNSUPPORTED, "llm_bench_kernels: the fused step does not cover this shape");
      HVX_CUDA(cudaMemsetAsync(L->fused_bar, 0, sizeof(unsigned long long), st));
      return launch_fused_step(e, st, L, b, head_k, fp, -100);
    }
    for (int l = 0; l < c.llm_layers; l++) {
      const LlmLayer& y = L->layer[l];
      if (which == 1) {
        GemvArgs g; g.x = b.h; g.ldx = H; g.W = y.qkv_w; g.bias = y.qkv_b; g.norm_w = y.ln1; g.eps = c.llm_eps; g.N = NQKV; g.K = H;
        g.mode = GEMV_QKV; g.qkv = make_qkv(e, L, l, b.q, L->seqs, head_k, 0, 0, rows);
        if ((rc = launch_gemv(e, st, g, rows))) return rc;
      } else if (which == 2) {
        if ((rc = launch_attn(e, st, L, l, b, rows, L->seqs, head_k, 0, 0, splits, false))) return rc;
      } else if (which == 3) {
        GemvArgs o; o.x = b.att; o.ldx = H; o.W = y.o_w; o.N = H; o.K = H; o.out = b.h; o.ldo = H; o.resid = b.h; o.ldr = H;
        if ((rc = launch_gemv(e, st, o, rows))) return rc;
      } else if (which == 4) {
        GemvArgs u; u.x = b.h; u.ldx = H; u.W = y.gu_w; u.norm_w = y.ln2; u.eps = c.llm_eps; u.N = 2 * I; u.K = H; u.mode = GEMV_SWIGLU;
        u.out = b.act; u.ldo = I;
        u.dbg = dbg_dev ? dbg_dev + (size_t)l * 8 : nullptr;
        if ((rc = launch_gemv(e, st, u, rows))) return rc;
      } else if (which == 5) {
        GemvArgs d; d.x = b.act; d.ldx = I; d.W = y.down_w; d.N = H; d.K = I; d.out = b.q; d.ldo = H; d.resid = b.h; d.ldr = H;
        if ((rc = launch_gemv(e, st, d, rows))) return rc;
      }
    }
    return HVX_OK;
  };
  if ((rc = body())) return rc;                       // warm-up
  HVX_CUDA(cudaEventRecord(e0, st));
  for (int i = 0; i < reps; i++) if ((rc = body())) return rc;
  HVX_CUDA(cudaEventRecord(e1, st));
  HVX_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  ms_out[0] = ms / reps;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (fdbg) {
    unsigned long long h[1024];
    cudaMemcpy(h, fdbg, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[5] = {"qkv", "attn", "o", "gu", "down"};
    const int nbar = 5 * c.llm_layers + 4;
    fprintf(stderr, "[fused timeline, CTA 0, ns since kernel start] consumer: exit of the grid barrier after each phase | producer: phase fully issued\n");
    for (int i = 1; i <= nbar; i++) {
      const int l = (i - 1) / 5, ph = (i - 1) % 5;
      if (l >= 2 && l < c.llm_layers - 1 && i <= 5 * c.llm_layers) continue;
      if (i <= 5 * c.llm_layers) fprintf(stderr, "  L%02d %-5s  +%6llu  (dt %5llu)\n", l, nm[ph], h[i] - h[0], h[i] - h[i - 1]);
      else fprintf(stderr, "  head phase %d  +%6llu  (dt %5llu)\n", i - 5 * c.llm_layers, h[i] - h[0], h[i] - h[i - 1]);
    }
    for (int i = 0; i < 4 * c.llm_layers + 5; i++) {
      if (i >= 12 && i < 4 * c.llm_layers - 4) continue;
      fprintf(stderr, "  producer phase %3d issued at +%6lld\n", i, (long long)(h[256 + i] - h[0]));
    }
    // SM-clock marks of CTA 0 inside layer 1 (16 per layer): stage | work done (barrier entry) | barrier exit
    const char* mn[14] = {"qkv: rows staged+normed", "qkv: dot+RoPE+cache done", "qkv: barrier", "attn: partials done", "attn: barrier",
                          "o: partials merged+staged", "o: dot done", "o: barrier", "gu: rows staged+normed", "gu: dot done", "gu: barrier", "down: act staged", "down: dot done", "down: barrier"};
    for (int i = 0; i < 14; i++) fprintf(stderr, "  L01 %-26s +%6llu clk\n", mn[i], h[600 + 14 + i] - h[600 + 13]);
    cudaFree(fdbg);
  }
  if (dbg_dev) {
    unsigned long long h[64 * 8];
    cudaMemcpy(h, dbg_dev, sizeof(h), cudaMemcpyDeviceToHost);
    for (int l = 0; l < 6; l++) {
      fprintf(stderr, "[gemv timeline layer %d] start %llu ns; +wait %llu  +x %llu  +norm %llu  +slot0 %llu  +slot1 %llu  +slot2 %llu  +end %llu   (next start +%llu)\n", l,
              h[l * 8], h[l * 8 + 1] - h[l * 8], h[l * 8 + 2] - h[l * 8], h[l * 8 + 3] - h[l * 8], h[l * 8 + 4] - h[l * 8],
              h[l * 8 + 5] > h[l * 8] ? h[l * 8 + 5] - h[l * 8] : 0, h[l * 8 + 6] > h[l * 8] ? h[l * 8 + 6] - h[l * 8] : 0, h[l * 8 + 7] - h[l * 8],
              h[(l + 1) * 8] - h[l * 8]);
    }
    cudaFree(dbg_dev);
  }
  if (L->graph) { cudaGraphExecDestroy(L->graph); L->graph = nullptr; }
  return HVX_OK;
}
