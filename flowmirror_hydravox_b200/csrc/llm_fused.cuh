// Persistent fused decode step for one sequence (batch-1 serving, inference_head_num <= 4): the 24 transformer layers,
// the final norm, the MTP heads and the llm_decoder logits of one AR step (llm_multi_head_v3.py:871-888) in ONE
// cooperative launch — included by llm.cu.
//
// Why: the step is a chain of ~100 dependent matrix-vector products whose weights (971 MB at K=2) are read exactly once.
// As separate kernels every link pays launch/drain + a cold weight stream (profiles/r1_decode_step_timeline.txt: 3-8 us per
// link against 4.6 us of HBM time for a whole layer).  Here one CTA per SM stays resident for the whole step:
//   * a producer warp streams this CTA's share of EVERY weight matrix, in schedule order, through a ring of shared-memory
//     slots with 1-D bulk copies (cp.async.bulk + mbarrier).  The share is static, so the stream never waits for
//     activations: it runs up to a ring (~180 KB/SM, ~27 MB/GPU = most of a layer) ahead of the math, across phase and
//     layer boundaries;
//   * 16 consumer warps do the dependent part: stage the phase's activation rows from L2 (ld.global.cg — they were written
//     by other SMs in the previous phase), fused RMSNorm, dot products against the landed slots, fused epilogues
//     (bias, RoPE + KV-cache write, SwiGLU, residual), then a grid-wide barrier (one release-atomic + acquire spin);
//   * attention over the KV cache is a phase of the same kernel: CTA = (q head, key split), partial softmax to global,
//     merged by the o-projection's staging.
// Phases per layer: qkv | attention | o-proj | gate/up | down  (5 grid barriers); heads: v | o | gate/up | down | logits.
// Every wait is bounded (globaltimer) and raises an abort flag instead of hanging the GPU.
#pragma once

namespace hvx {
namespace fused {

constexpr int NW = 16;                       // consumer warps
constexpr int NCONS = NW * 32;
constexpr int THREADS = NCONS + 32;          // + producer warp
constexpr int SLOT = 20 * 1024;              // ring slot: 5 pairs at K=896, one k-split pair (4 x 1216) at K=4864
constexpr int MAXNS = 10;
constexpr int RED_FLOATS = 1024;             // k-split partial sums
constexpr int ATT_CHUNK = 64;                // keys staged per chunk
constexpr long long TIMEOUT_CLK = 800ll * 1000 * 1000;   // ~0.4 s of SM clocks (clock64 is SM-local and cheap; %globaltimer costs ~1 us per read)

struct Args {
  const LlmLayer* layers = nullptr;          // device copy of the per-layer pointer table
  int n_layers = 0, H = 0, I = 0, q_heads = 0, kv_heads = 0, MI = 0, V = 0, head_k = 1, max_ctx = 0;
  float eps = 1e-6f, scale = 0.125f;
  const float* norm = nullptr;
  const __nv_bfloat16 *m_v_w = nullptr, *m_o_w = nullptr, *m_gu_w = nullptr, *m_down_w = nullptr, *dec_w = nullptr;
  const float *m_v_b = nullptr, *m_ln1 = nullptr, *m_ln2 = nullptr;
  uint8_t *kc = nullptr, *vc = nullptr;
  size_t layer_stride = 0, seq_stride = 0;   // elements
  int kv_f32 = 0;
  const float* inv_freq = nullptr;
  const SeqState* seqs = nullptr;
  float *h = nullptr, *q = nullptr, *att = nullptr, *act = nullptr, *part = nullptr;
  float *m_v = nullptr, *m_h1 = nullptr, *m_act = nullptr, *m_o = nullptr, *logits = nullptr;
  unsigned long long* bar = nullptr;         // grid-barrier counter, zero at launch (the sampler kernel resets it)
  int* abort_flag = nullptr;
  int n_slots = 0, x_bytes = 0;              // ring geometry / activation area (bytes, multiple of 128)
  int max_phases = 1 << 30;                  // debugging: stop after this many grid barriers
  int pace_clk = 0;                          // producer pacing: SM clocks per full slot (0 = issue as fast as slots free up)
  unsigned long long* dbg = nullptr;         // optional timeline of CTA 0: [0] start, [1+i] exit of grid barrier i, [256+i] producer done issuing phase i
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void cons_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory"); }

// bounded mbarrier wait: a scheduling bug must end in an error code, not in a hung GPU
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity, int* abort_flag) {
  if (tc::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  for (uint32_t it = 1;; it++) {
    if (tc::mbar_try_wait(bar, parity)) return;
    if ((it & 1023) == 0) {
      if (*reinterpret_cast<volatile int*>(abort_flag)) return;
      if (clock64() - t0 > TIMEOUT_CLK) { atomicExch(abort_flag, 2); return; }
    }
  }
}

// all consumer threads; the producer warp never takes part (weights are immutable).  Release/acquire at gpu scope on the
// counter (bar.sync makes the CTA's writes cumulative with thread 0's release); no separate fences, no value-returning atomic.
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target, int* abort_flag) {
  cons_sync();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(bar), "l"(1ULL) : "memory");
    if (ld_acquire_u64(bar) < target) {
      const long long t0 = clock64();
      for (uint32_t it = 1; ld_acquire_u64(bar) < target; it++) {
        if ((it & 1023) == 0) {
          if (*reinterpret_cast<volatile int*>(abort_flag)) break;
          if (clock64() - t0 > TIMEOUT_CLK) { atomicExch(abort_flag, 3); break; }
        }
      }
    }
  }
  cons_sync();
}

// One matrix-vector phase: nb stacked [N][K] matrices; this CTA's contiguous range of output pairs, cut into units
// (pair, k-segment) that fill ring slots in order.  Computed identically by the producer and the consumers.
struct Phase {
  const __nv_bfloat16* W;
  int N, K, nb, ks, ks_shift, seglen, unit_bytes, upu, pairsN, p_begin, n_pairs, n_units, n_slots;
};
// unit -> (batch j, pair in batch pp, k-segment): ks is a power of two and nb is small, so no integer division
__device__ __forceinline__ void unit_decode(const Phase& p, int u, int& pl, int& seg, int& j, int& pp) {
  pl = u >> p.ks_shift; seg = u & (p.ks - 1);
  pp = p.p_begin + pl; j = 0;
  while (pp >= p.pairsN) { pp -= p.pairsN; j++; }
}
__host__ __device__ inline int pick_ksplit(int K) {
  for (int c = 16; c > 1; c >>= 1)
    if (K % (8 * c) == 0 && K / c >= 1024) return c;
  return 1;
}
__device__ __noinline__ Phase make_phase(const __nv_bfloat16* W, int N, int K, int nb, int rot) {
  Phase p;
  p.W = W; p.N = N; p.K = K; p.nb = nb;
  p.ks = pick_ksplit(K);
  p.ks_shift = 31 - __clz(p.ks);
  p.seglen = K >> p.ks_shift;
  p.unit_bytes = p.seglen * 4;                                   // two rows of bf16
  p.upu = max(1, SLOT / p.unit_bytes);
  p.pairsN = (N + 1) >> 1;
  const int P = p.pairsN * nb, G = gridDim.x;
  const int c = (int)((blockIdx.x + (unsigned)rot) % (unsigned)G);   // rotate the CTAs that get the remainder pairs
  const int base = P / G, extra = P - base * G;
  p.p_begin = c * base + min(c, extra);
  p.n_pairs = base + (c < extra ? 1 : 0);
  p.n_units = p.n_pairs * p.ks;
  p.n_slots = (p.n_units + p.upu - 1) / p.upu;
  return p;
}

// producer warp: stream the phase's units into the ring; lane i issues unit i of the slot (address math in parallel — a
// single lane doing it serially was instruction-bound at ~1.6 TB/s)
// Pacing: a phase frees its slots in a burst, and an unpaced producer turns that into ~18 MB of requests queued at once —
// every latency-critical load of the consumers (activation rows, barrier polls) then waits behind microseconds of weight
// traffic.  The producer therefore issues at most one slot per `pace` clocks (~ this SM's share of the HBM rate).
__device__ __noinline__ void produce(const Phase& p, uint8_t* ring, int NS, uint64_t* full_bar, uint64_t* empty_bar,
                                        uint32_t& pseq, int* abort_flag, int pace, long long& next_issue) {
  const int lane = threadIdx.x & 31;
  for (int sl = 0; sl < p.n_slots; sl++, pseq++) {
    const int rs = (int)(pseq % (uint32_t)NS);
    if (pseq >= (uint32_t)NS) mbar_wait_b(&empty_bar[rs], ((pseq / NS) - 1) & 1, abort_flag);
    if (pace > 0) {
      long long now = clock64();
      while (now < next_issue) now = clock64();
      next_issue = now + pace;
    }
    const int u0 = sl * p.upu, nu = min(p.upu, p.n_units - u0);
    const uint32_t rb = (uint32_t)p.seglen * 2;
    // pass 1: bytes of the slot (the last pair of an odd N has one row)
    uint32_t bytes = 0;
    for (int i = lane; i < nu; i += 32) {
      int pl, seg, j, pp;
      unit_decode(p, u0 + i, pl, seg, j, pp);
      bytes += (2 * pp + 1 < p.N) ? 2 * rb : rb;
    }
    for (int o = 16; o; o >>= 1) bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
    if (lane == 0) tc::mbar_expect_tx(&full_bar[rs], bytes);
    __syncwarp();
    uint8_t* base = ring + (size_t)rs * SLOT;
    for (int i = lane; i < nu; i += 32) {
      int pl, seg, j, pp;
      unit_decode(p, u0 + i, pl, seg, j, pp);
      const int n0 = 2 * pp;
      const bool two = n0 + 1 < p.N;
      const __nv_bfloat16* src = p.W + ((size_t)j * p.N + n0) * p.K + (size_t)seg * p.seglen;
      uint8_t* dst = base + (size_t)i * p.unit_bytes;
      if (p.ks == 1) {
        bulk_g2s(dst, src, two ? 2 * rb : rb, &full_bar[rs]);     // rows 2p, 2p+1 are contiguous
      } else {
        bulk_g2s(dst, src, rb, &full_bar[rs]);
        if (two) bulk_g2s(dst + rb, src + p.K, rb, &full_bar[rs]);
      }
    }
  }
}

__device__ __forceinline__ void fma8v(float& acc, const uint4 u, const float4 xa, const float4 xb) {
  acc = fmaf(bf_lo(u.x), xa.x, acc); acc = fmaf(bf_hi(u.x), xa.y, acc);
  acc = fmaf(bf_lo(u.y), xa.z, acc); acc = fmaf(bf_hi(u.y), xa.w, acc);
  acc = fmaf(bf_lo(u.z), xb.x, acc); acc = fmaf(bf_hi(u.z), xb.y, acc);
  acc = fmaf(bf_lo(u.w), xb.z, acc); acc = fmaf(bf_hi(u.w), xb.w, acc);
}

// consumers: RB activation rows per batch.  XG = false: rows in shared memory, two-plane layout (floats k%8<4 | k%8>=4 of a
// row in separate halves: the lanes' float4 reads are bank-conflict free), batch j at xs + j*x_bstride.  XG = true: rows
// read straight from global/L2 (K too long to stage: the MTP down-projection, K = 22016).
// epi(j, pair_in_batch, n0, row, y0, y1, two) finishes output features (n0, n0+1) of one row.
// llm_qkv_store without its global reads (sequence state and RoPE frequencies are already on chip): pair (n, n+1) of row `row`
// at position `pos` -> RoPE on q/k, q to the q buffer, k/v into the cache
__device__ __forceinline__ void qkv_store_fast(const LlmQkvEpi& q, const float* s_invf, int pos, int row, int n, float v0, float v1) {
  if (pos >= q.max_ctx) return;
  const bool is_q = n < q.q_dim;
  const bool is_k = !is_q && n < q.q_dim + q.kv_dim;
  if (is_q || is_k) {
    float sn, cs;
    sincos_noinline((float)pos * s_invf[(n & 63) >> 1], &sn, &cs);
    const float r0 = v0 * cs - v1 * sn, r1 = v1 * cs + v0 * sn;
    v0 = r0; v1 = r1;
  }
  if (is_q) {
    *reinterpret_cast<float2*>(q.q_out + (size_t)row * q.ldq + n) = make_float2(v0, v1);
  } else {
    const int mm = n - q.q_dim - (is_k ? 0 : q.kv_dim);
    const size_t idx = ((size_t)(mm >> 6) * q.max_ctx + pos) * 64 + (mm & 63);      // sequence slot 0
    void* base = is_k ? q.kc : q.vc;
    if (q.kv_f32) *reinterpret_cast<float2*>(reinterpret_cast<float*>(base) + idx) = make_float2(v0, v1);
    else *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = __floats2bfloat162_rn(v0, v1);
  }
}

// Epilogue of a matrix-vector phase, selected at run time: ONE copy of the consumer loop per (rows, x source) instead of one
// per call site — the step's instruction footprint has to stay inside the instruction cache (a 17.6 k-instruction first
// version re-fetched its code from L2 on every phase of every layer).
enum { EPI_STORE = 0, EPI_SWIGLU = 1, EPI_QKV = 2 };
struct Epi {
  int mode = EPI_STORE;
  float* out = nullptr; size_t so = 0, ld = 0;     // out[j*so + row*ld + n]  (SWIGLU: index pair_in_batch instead of n)
  const float* s_add = nullptr;                    // optional shared-memory addend indexed by n (the heads' residual hn)
  const LlmQkvEpi* qe = nullptr; const float* s_invf = nullptr; int pos0 = 0;   // EPI_QKV
};
__device__ __forceinline__ void run_epi(const Epi& e, int j, int pp, int n0, int r, float y0, float y1, bool two) {
  if (e.mode == EPI_QKV) {
    qkv_store_fast(*e.qe, e.s_invf, e.pos0 + r, r, n0, y0, y1);
  } else if (e.mode == EPI_SWIGLU) {
    e.out[j * e.so + (size_t)r * e.ld + pp] = (y0 / (1.0f + expf(-y0))) * y1;
  } else {
    if (e.s_add) { y0 += e.s_add[n0]; if (two) y1 += e.s_add[n0 + 1]; }
    float* d = e.out + j * e.so + (size_t)r * e.ld + n0;
    d[0] = y0;
    if (two) d[1] = y1;
  }
}
struct Ring { uint8_t* ring; int NS; uint64_t* full_bar; uint64_t* empty_bar; uint32_t cseq; float* s_red; int* abort_flag; };

// `add`: optional per-output addend fetched BEFORE the dot product so its L2 latency hides behind it: add.bias[j*sb + n]
// (weights-like, cached) and/or add.resid[j*sr + row*ld + n] (produced in this launch -> ld.global.cg).
struct Addend { const float* bias = nullptr; size_t sb = 0; const float* resid = nullptr; size_t sr = 0, ld = 0; };
template <int RB, bool XG>
__device__ __noinline__ void consume(const Phase& p, Ring& rg, const float* xs, size_t x_bstride, int rows_live, const Addend& add,
                                     const Epi& epi) {
  uint8_t* ring = rg.ring; const int NS = rg.NS; uint64_t* full_bar = rg.full_bar; uint64_t* empty_bar = rg.empty_bar;
  const uint32_t cseq = rg.cseq; float* s_red = rg.s_red; int* abort_flag = rg.abort_flag;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = p.K, KH = K >> 1;
  // Every warp passes every slot in order (wait full -> own units -> arrive empty), also the slots that hold none of
  // its units: an arrival for the slot's NEXT use can then never overtake a slower warp's arrival for this use.
  for (int sl = 0; sl < p.n_slots; sl++) {
    const uint32_t gs = cseq + sl;
    const int rs = (int)(gs % (uint32_t)NS);
    mbar_wait_b(&full_bar[rs], (gs / NS) & 1, abort_flag);
    const int ub = sl * p.upu, ue = min(p.n_units, ub + p.upu);
   for (int u = ub + ((warp - ub) & (NW - 1)); u < ue; u += NW) {
    const __nv_bfloat16* ws = reinterpret_cast<const __nv_bfloat16*>(ring + (size_t)rs * SLOT + (size_t)(u - ub) * p.unit_bytes);
    int pl, seg, j, pp;
    unit_decode(p, u, pl, seg, j, pp);
    const int n0 = 2 * pp;
    const bool two = n0 + 1 < p.N;
    const __nv_bfloat16* w0 = ws;
    const __nv_bfloat16* w1 = two ? ws + p.seglen : ws;
    const int kb = seg * p.seglen;
    const float* xb = xs + (size_t)j * x_bstride;
    float2 pre = make_float2(0.f, 0.f);
    if (p.ks == 1 && lane < RB && lane < rows_live) {
      if (add.bias) { pre.x = __ldg(add.bias + j * add.sb + n0); if (two) pre.y = __ldg(add.bias + j * add.sb + n0 + 1); }
      if (add.resid) {
        const float* rp = add.resid + j * add.sr + (size_t)lane * add.ld + n0;
        pre.x += __ldcg(rp); if (two) pre.y += __ldcg(rp + 1);
      }
    }
    float acc0[RB], acc1[RB], bcc0[RB], bcc1[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) { acc0[r] = 0.f; acc1[r] = 0.f; bcc0[r] = 0.f; bcc1[r] = 0.f; }
    for (int kq = 0; kq < p.seglen; kq += 1024) {
      // one 1024-element block: the 8 weight loads are issued before any dependent math (memory-level parallelism instead
      // of 4 serial load->FMA rounds); chunks beyond the segment are predicated off
      uint4 wa[4], wb[4];
      bool on[4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int k = kq + c * 256 + lane * 8;
        on[c] = k < p.seglen;
        wa[c] = make_uint4(0, 0, 0, 0); wb[c] = wa[c];
        if (on[c]) { wa[c] = *reinterpret_cast<const uint4*>(w0 + k); wb[c] = *reinterpret_cast<const uint4*>(w1 + k); }
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {
        if (!on[c]) continue;
        const int k = kq + c * 256 + lane * 8;
#pragma unroll
        for (int r = 0; r < RB; r++) {
          float4 xa, xc;
          if constexpr (XG) {
            const float4* g = reinterpret_cast<const float4*>(xb + (size_t)r * K + kb + k);
            xa = __ldcg(g); xc = __ldcg(g + 1);
          } else {
            const float* sp = xb + r * K + ((kb + k) >> 1);
            xa = *reinterpret_cast<const float4*>(sp); xc = *reinterpret_cast<const float4*>(sp + KH);
          }
          if (c & 1) { fma8v(bcc0[r], wa[c], xa, xc); fma8v(bcc1[r], wb[c], xa, xc); }
          else { fma8v(acc0[r], wa[c], xa, xc); fma8v(acc1[r], wb[c], xa, xc); }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < RB; r++) {
      acc0[r] += bcc0[r]; acc1[r] += bcc1[r];
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        acc0[r] += __shfl_xor_sync(0xffffffffu, acc0[r], o);
        acc1[r] += __shfl_xor_sync(0xffffffffu, acc1[r], o);
      }
    }
    float y0 = 0.f, y1 = 0.f;
#pragma unroll
    for (int r = 0; r < RB; r++) if (lane == r) { y0 = acc0[r]; y1 = acc1[r]; }
    if (p.ks == 1) {
      if (lane < RB && lane < rows_live) run_epi(epi, j, pp, n0, lane, y0 + pre.x, y1 + pre.y, two);
    } else if (lane < RB) {
      s_red[(size_t)u * 2 * RB + lane * 2] = y0;
      s_red[(size_t)u * 2 * RB + lane * 2 + 1] = y1;
    }
   }
    __syncwarp();
    if (lane == 0) tc::mbar_arrive(&empty_bar[rs]);
  }
  rg.cseq = cseq + p.n_slots;
  if (p.ks > 1) {
    cons_sync();
    for (int t = threadIdx.x; t < p.n_pairs * RB; t += NCONS) {
      const int pl = t / RB, r = t - pl * RB;
      if (r >= rows_live) continue;
      float y0 = 0.f, y1 = 0.f;
      for (int s = 0; s < p.ks; s++) {
        y0 += s_red[(size_t)((pl << p.ks_shift) + s) * 2 * RB + r * 2];
        y1 += s_red[(size_t)((pl << p.ks_shift) + s) * 2 * RB + r * 2 + 1];
      }
      int pl2, seg2, j, pp;
      unit_decode(p, pl << p.ks_shift, pl2, seg2, j, pp);
      const int n0 = 2 * pp;
      const bool two = n0 + 1 < p.N;
      if (add.bias) { y0 += __ldg(add.bias + j * add.sb + n0); if (two) y1 += __ldg(add.bias + j * add.sb + n0 + 1); }
      if (add.resid) {
        const float* rp = add.resid + j * add.sr + (size_t)r * add.ld + n0;
        y0 += __ldcg(rp); if (two) y1 += __ldcg(rp + 1);
      }
      run_epi(epi, j, pp, n0, r, y0, y1, two);
    }
  }
}

// nrows rows of K floats (global, produced by other CTAs during this launch -> ld.global.cg) into the two-plane shared
// layout, zero rows up to RB, optional fused RMSNorm (HF Qwen2RMSNorm: fp32, eps inside rsqrt, weight applied last);
// nw_stride: norm-weight stride between rows (0: shared).
template <int RB>
__device__ __noinline__ void stage_rows(float* sx, const float* src, size_t ld, int nrows, int K, const float* nw,
                                           size_t nw_stride, float eps, float* s_part) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int KH = K >> 1, nq = K >> 2;
  float ss[RB];
#pragma unroll
  for (int r = 0; r < RB; r++) ss[r] = 0.f;
  if (nw && nq <= NCONS) {
    // one float4 per thread and row: rows and norm weights stay in registers across the block reduction (one L2 round trip,
    // one shared-memory write)
    const int q = tid;
    const bool on = q < nq;
    const int dst = (q & 1) * KH + (q >> 1) * 4;
    float4 v[RB], w4[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) {
      v[r] = make_float4(0.f, 0.f, 0.f, 0.f); w4[r] = v[r];
      if (on && r < nrows) {
        v[r] = __ldcg(reinterpret_cast<const float4*>(src + (size_t)r * ld) + q);
        w4[r] = __ldg(reinterpret_cast<const float4*>(nw + (size_t)r * nw_stride) + q);
      }
      ss[r] = v[r].x * v[r].x + v[r].y * v[r].y + v[r].z * v[r].z + v[r].w * v[r].w;
      for (int o = 16; o; o >>= 1) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
      if (lane == 0) s_part[r * NW + warp] = ss[r];
    }
    cons_sync();
#pragma unroll
    for (int r = 0; r < RB; r++) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; w++) t += s_part[r * NW + w];
      const float sc = rsqrtf(t / (float)K + eps);
      if (on) {
        float4 y;
        y.x = w4[r].x * (v[r].x * sc); y.y = w4[r].y * (v[r].y * sc); y.z = w4[r].z * (v[r].z * sc); y.w = w4[r].w * (v[r].w * sc);
        *reinterpret_cast<float4*>(&sx[r * K + dst]) = y;
      }
    }
    cons_sync();
    return;
  }
  for (int q = tid; q < nq; q += NCONS) {
    const int dst = (q & 1) * KH + (q >> 1) * 4;
#pragma unroll
    for (int r = 0; r < RB; r++) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) v = __ldcg(reinterpret_cast<const float4*>(src + (size_t)r * ld) + q);
      ss[r] += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      *reinterpret_cast<float4*>(&sx[r * K + dst]) = v;
    }
  }
  if (nw) {
#pragma unroll
    for (int r = 0; r < RB; r++) {
      for (int o = 16; o; o >>= 1) ss[r] += __shfl_xor_sync(0xffffffffu, ss[r], o);
      if (lane == 0) s_part[r * NW + warp] = ss[r];
    }
    cons_sync();
    float sc[RB];
#pragma unroll
    for (int r = 0; r < RB; r++) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < NW; w++) t += s_part[r * NW + w];
      sc[r] = rsqrtf(t / (float)K + eps);
    }
    for (int q = tid; q < nq; q += NCONS) {
      const int dst = (q & 1) * KH + (q >> 1) * 4;
#pragma unroll
      for (int r = 0; r < RB; r++) {
        if (r < nrows) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(nw + (size_t)r * nw_stride) + q);
          float4 v = *reinterpret_cast<float4*>(&sx[r * K + dst]);
          v.x = w4.x * (v.x * sc[r]); v.y = w4.y * (v.y * sc[r]); v.z = w4.z * (v.z * sc[r]); v.w = w4.w * (v.w * sc[r]);
          *reinterpret_cast<float4*>(&sx[r * K + dst]) = v;
        }
      }
    }
  }
  cons_sync();
}

// attention of the step's rows over the KV cache, one phase: CTA = (q head, key split); its 16 warps = rows x key sub-ranges,
// every warp reads its keys straight from the cache (L2) — 4 lanes per key, 8 keys per iteration — and the warps of a row
// merge through shared memory.  Partial (m, l, o[64]) per (row, q head, split) -> a.part; the <= 10 splits are merged by the
// o-projection's staging (stage_att), so there is no separate merge phase / grid barrier.
template <bool KV32>
__device__ __noinline__ void attention_phase(const Args& a, int layer, int rows, int ctx, int S, float* s_mrg /* [NW][68] */) {
  using KT = typename std::conditional<KV32, float, __nv_bfloat16>::type;
  const int task = blockIdx.x;
  if (task >= a.q_heads * S) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qh = task / S, split = task - qh * S;
  const int group = a.q_heads / a.kv_heads, kvh = qh / group;
  const int nk_max = min(ctx + rows, a.max_ctx);
  const int per = (nk_max + S - 1) / S;
  const int k_begin = split * per, k_end = min(nk_max, k_begin + per);
  const int rp2 = rows <= 1 ? 1 : rows <= 2 ? 2 : 4;           // warps per row = 16 / rp2
  const int wpr = NW / rp2;
  const int r = warp / wpr, ws = warp - r * wpr;
  const bool have = r < rows;
  const int d4 = lane & 3, kslot = lane >> 2;
  const size_t esz = sizeof(KT);
  const KT* kb = reinterpret_cast<const KT*>(a.kc + (size_t)layer * a.layer_stride * esz) + (size_t)kvh * a.max_ctx * 64;   // sequence slot 0
  const KT* vb = reinterpret_cast<const KT*>(a.vc + (size_t)layer * a.layer_stride * esz) + (size_t)kvh * a.max_ctx * 64;
  float m = -INFINITY, l = 0.f, o[16];
#pragma unroll
  for (int i = 0; i < 16; i++) o[i] = 0.f;
  if (have) {
    float qv[16];
    const float4* qp = reinterpret_cast<const float4*>(a.q + (size_t)r * a.H + qh * 64 + d4 * 16);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float4 t = __ldcg(qp + i);
      qv[4 * i] = t.x * a.scale; qv[4 * i + 1] = t.y * a.scale; qv[4 * i + 2] = t.z * a.scale; qv[4 * i + 3] = t.w * a.scale;
    }
    const int n = max(k_end - k_begin, 0), perw = (n + wpr - 1) / wpr;
    const int wb = k_begin + ws * perw, we = min(min(k_end, wb + perw), min(ctx + r + 1, a.max_ctx));
#pragma unroll 2
    for (int k0 = wb; k0 < we; k0 += 8) {
      const int key = k0 + kslot;
      const bool live = key < we;
      float kf[16], vf[16];
      if (live) {
        if constexpr (KV32) {
          const float4* kp = reinterpret_cast<const float4*>(kb + (size_t)key * 64 + d4 * 16);
          const float4* vp = reinterpret_cast<const float4*>(vb + (size_t)key * 64 + d4 * 16);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 u = __ldcg(kp + i), w = __ldcg(vp + i);
            kf[4 * i] = u.x; kf[4 * i + 1] = u.y; kf[4 * i + 2] = u.z; kf[4 * i + 3] = u.w;
            vf[4 * i] = w.x; vf[4 * i + 1] = w.y; vf[4 * i + 2] = w.z; vf[4 * i + 3] = w.w;
          }
        } else {
          const uint4* kp = reinterpret_cast<const uint4*>(kb + (size_t)key * 64 + d4 * 16);
          const uint4* vp = reinterpret_cast<const uint4*>(vb + (size_t)key * 64 + d4 * 16);
#pragma unroll
          for (int i = 0; i < 2; i++) {
            const uint4 u = __ldcg(kp + i), w = __ldcg(vp + i);
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
  // the 8 key slots of the warp -> one (m, l, o[64]) per warp in shared memory
  {
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
    if (kslot == 0) {
#pragma unroll
      for (int i = 0; i < 4; i++)
        *reinterpret_cast<float4*>(&s_mrg[warp * 68 + 4 + d4 * 16 + 4 * i]) = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    if (lane == 0) { s_mrg[warp * 68] = M; s_mrg[warp * 68 + 1] = Lr; }
  }
  cons_sync();
  if (warp < rows) {                                           // warp w merges the warps of row w
    const float* e0 = s_mrg + (size_t)warp * wpr * 68;
    float Mx = -INFINITY;
    for (int t = 0; t < wpr; t++) Mx = fmaxf(Mx, e0[t * 68]);
    float Lt = 0.f, lo = 0.f, hi = 0.f;
    for (int t = 0; t < wpr; t++) {
      const float mw = e0[t * 68];
      const float wg = (mw == -INFINITY) ? 0.f : expf(mw - Mx);
      Lt += e0[t * 68 + 1] * wg; lo += e0[t * 68 + 4 + lane] * wg; hi += e0[t * 68 + 36 + lane] * wg;
    }
    float* pp = a.part + (((size_t)warp * a.q_heads + qh) * S + split) * 68;
    if (lane == 0) { pp[0] = Mx; pp[1] = Lt; }
    pp[4 + lane] = lo;
    pp[36 + lane] = hi;
  }
}

// o-projection staging: merge the S split partials of every (row, q head) and write the attention rows in the two-plane layout
template <int RB>
__device__ __noinline__ void stage_att(float* sx, const Args& a, int rows, int S) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int K = a.H, KH = K >> 1;
  for (int i = tid; i < (RB - rows) * K; i += NCONS) sx[rows * K + i] = 0.f;
  const int npairs = rows * a.q_heads;
  for (int p = warp; p < npairs; p += NW) {
    const float* pb = a.part + (size_t)p * S * 68;
    float Mx = -INFINITY;
    for (int sp = 0; sp < S; sp++) Mx = fmaxf(Mx, __ldcg(pb + sp * 68));
    float Ls = 0.f, lo = 0.f, hi = 0.f;
#pragma unroll 4
    for (int sp = 0; sp < S; sp++) {
      const float ms = __ldcg(pb + sp * 68);
      const float w = (ms == -INFINITY) ? 0.f : expf(ms - Mx);
      Ls += __ldcg(pb + sp * 68 + 1) * w; lo += __ldcg(pb + sp * 68 + 4 + lane) * w; hi += __ldcg(pb + sp * 68 + 36 + lane) * w;
    }
    const float inv = 1.0f / Ls;
    const int r = p / a.q_heads, qh = p - r * a.q_heads;
    const int k0 = qh * 64 + lane, k1 = k0 + 32;
    sx[r * K + ((k0 >> 2) & 1) * KH + (k0 >> 3) * 4 + (k0 & 3)] = lo * inv;
    sx[r * K + ((k1 >> 2) & 1) * KH + (k1 >> 3) * 4 + (k1 & 3)] = hi * inv;
  }
  cons_sync();
}

__device__ __forceinline__ int attn_splits(const Args& a, int nk_max) {
  const int cap = max(1, (int)gridDim.x / a.q_heads);
  return max(1, min(cap, (nk_max + 127) / 128));
}

template <int R>
__global__ void __launch_bounds__(THREADS, 1) llm_fused_step_kernel(Args a) {   // 17 warps are allocated as 20: 96 registers per thread at most
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ __align__(8) uint64_t full_bar[MAXNS], empty_bar[MAXNS];
  __shared__ float s_red[RED_FLOATS];
  __shared__ float s_part[LLM_MAX_HEADS * NW];
  __shared__ __align__(16) float s_mrg[NW * 68];
  __shared__ float s_invf[32];
  const SeqState st = a.seqs[0];
  if (st.done || st.n_new <= 0) return;                      // uniform over the grid: no barrier is entered
  const int rows = min(st.n_new, R), ctx = st.ctx;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NS = a.n_slots;
  float* sx = reinterpret_cast<float*>(smraw);
  uint8_t* ring = smraw + a.x_bytes;
  float* s_hn = reinterpret_cast<float*>(ring + (size_t)NS * SLOT);     // [H] final-normed hidden state (heads)
  if (tid < 32) s_invf[tid] = a.inv_freq[tid];
  if (tid == 0) {
    for (int i = 0; i < NS; i++) { tc::mbar_init(&full_bar[i], 1); tc::mbar_init(&empty_bar[i], NW); }
    tc::fence_barrier_init();
  }
  __syncthreads();
  const int H = a.H, I = a.I, NQKV = (a.q_heads + 2 * a.kv_heads) * 64, MI = a.MI, HK = a.head_k;
  if (a.max_phases < 0) {                                   // microbenchmark: -max_phases back-to-back grid barriers, nothing else
    if (warp == NW) return;
    unsigned long long t = 0;
    for (int i = 0; i < -a.max_phases; i++) { t += gridDim.x; grid_barrier(a.bar, t, a.abort_flag); }
    return;
  }

  if (warp == NW) {
    // ------------------------------------------------ producer: the whole step's weight stream, in schedule order
    {
      uint32_t pseq = 0;
      long long next_issue = 0;
      int nb = 0, pi = 0;
      const bool pd = a.dbg && blockIdx.x == 0 && lane == 0;
#define HVX_PMARK() do { if (pd && pi < 250) a.dbg[256 + pi] = gtime(); pi++; } while (0)
      for (int l = 0; l < a.n_layers && nb < a.max_phases; l++) {
        const LlmLayer y = a.layers[l];
        if (nb < a.max_phases) produce(make_phase(y.qkv_w, NQKV, H, 1, 4 * l), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
        nb += 2;
        if (nb < a.max_phases) produce(make_phase(y.o_w, H, H, 1, 4 * l + 1), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
        nb += 1;
        if (nb < a.max_phases) produce(make_phase(y.gu_w, 2 * I, H, 1, 4 * l + 2), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
        nb += 1;
        if (nb < a.max_phases) produce(make_phase(y.down_w, H, I, 1, 4 * l + 3), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
        nb += 1;
      }
      if (nb < a.max_phases) {
        produce(make_phase(a.m_v_w, H, H, HK, 1), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
        produce(make_phase(a.m_o_w, H, H, HK, 2), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
        produce(make_phase(a.m_gu_w, 2 * MI, H, HK, 3), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
        produce(make_phase(a.m_down_w, H, MI, HK, 4), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
        produce(make_phase(a.dec_w, a.V, H, 1, 5), ring, NS, full_bar, empty_bar, pseq, a.abort_flag, a.pace_clk, next_issue);
        HVX_PMARK();
      }
    }
#undef HVX_PMARK
    return;
  }

  // ---------------------------------------------------- consumers
  if (a.dbg && blockIdx.x == 0 && tid == 0) a.dbg[0] = gtime();
  Ring rg; rg.ring = ring; rg.NS = NS; rg.full_bar = full_bar; rg.empty_bar = empty_bar; rg.cseq = 0; rg.s_red = s_red; rg.abort_flag = a.abort_flag;
  unsigned long long bt = 0;
  const unsigned long long G = gridDim.x;
  int nb = 0;                                                // grid barriers passed (== phases done)
  int mk = 0;
#define HVX_MARK() do { if (a.dbg && blockIdx.x == 0 && tid == 0 && mk < 400) a.dbg[600 + mk] = (unsigned long long)clock64(); mk++; } while (0)
#define HVX_GRID_BARRIER() do { HVX_MARK(); bt += G; grid_barrier(a.bar, bt, a.abort_flag); nb++; if (a.dbg && blockIdx.x == 0 && tid == 0 && nb < 255) a.dbg[nb] = gtime(); HVX_MARK(); } while (0)
  const int S = attn_splits(a, min(ctx + rows, a.max_ctx));
  LlmQkvEpi qe;
  qe.q_out = a.q; qe.ldq = H; qe.kv_f32 = a.kv_f32; qe.seq_stride = a.seq_stride; qe.max_ctx = a.max_ctx;
  qe.q_dim = a.q_heads * 64; qe.kv_dim = a.kv_heads * 64; qe.inv_freq = a.inv_freq; qe.seqs = a.seqs; qe.rows_per_seq = HK;
  qe.n_rows = rows;
  const size_t esz = a.kv_f32 ? 4 : 2;
  for (int l = 0; l < a.n_layers && nb < a.max_phases; l++) {
    const LlmLayer y = a.layers[l];
    // ---- qkv = W_qkv rmsnorm(h) + b, RoPE, K/V rows into the cache
    stage_rows<R>(sx, a.h, H, rows, H, y.ln1, 0, a.eps, s_part);
    HVX_MARK();
    {
      qe.kc = a.kc + (size_t)l * a.layer_stride * esz; qe.vc = a.vc + (size_t)l * a.layer_stride * esz;
      Addend ad; ad.bias = y.qkv_b;
      const Phase p = make_phase(y.qkv_w, NQKV, H, 1, 4 * l);
      Epi ep; ep.mode = EPI_QKV; ep.qe = &qe; ep.s_invf = s_invf; ep.pos0 = ctx;
      consume<R, false>(p, rg, sx, 0, rows, ad, ep);
    }
    HVX_GRID_BARRIER();
    if (nb >= a.max_phases) break;
    // ---- attention over the cache (split partials)
    if (a.kv_f32) attention_phase<true>(a, l, rows, ctx, S, s_mrg); else attention_phase<false>(a, l, rows, ctx, S, s_mrg);
    HVX_GRID_BARRIER();
    if (nb >= a.max_phases) break;
    // ---- h += W_o att (the split partials are merged while staging)
    stage_att<R>(sx, a, rows, S);
    HVX_MARK();
    {
      const Phase p = make_phase(y.o_w, H, H, 1, 4 * l + 1);
      float* h = a.h;
      Addend ad; ad.resid = a.h; ad.ld = H;
      Epi ep; ep.out = h; ep.ld = H;
      consume<R, false>(p, rg, sx, 0, rows, ad, ep);
    }
    HVX_GRID_BARRIER();
    if (nb >= a.max_phases) break;
    // ---- act = silu(gate) * up over rmsnorm(h)
    stage_rows<R>(sx, a.h, H, rows, H, y.ln2, 0, a.eps, s_part);
    HVX_MARK();
    {
      const Phase p = make_phase(y.gu_w, 2 * I, H, 1, 4 * l + 2);
      float* act = a.act;
      Epi ep; ep.mode = EPI_SWIGLU; ep.out = act; ep.ld = I;
      consume<R, false>(p, rg, sx, 0, rows, Addend(), ep);
    }
    HVX_GRID_BARRIER();
    if (nb >= a.max_phases) break;
    // ---- h += W_down act
    stage_rows<R>(sx, a.act, I, rows, I, nullptr, 0, a.eps, s_part);
    HVX_MARK();
    {
      const Phase p = make_phase(y.down_w, H, I, 1, 4 * l + 3);
      float* h = a.h;
      Addend ad; ad.resid = a.h; ad.ld = H;
      Epi ep; ep.out = h; ep.ld = H;
      consume<R, false>(p, rg, sx, 0, rows, ad, ep);
    }
    HVX_GRID_BARRIER();
  }
  if (nb >= a.max_phases) return;

  // ---------------------------------------------------- heads (llm_multi_head_v3.py:883-888; SURVEY App. A.1 closed form)
  // hn = final rmsnorm of the last live row; head j: v = W_v rmsnorm_j(hn) + b_v ; h1 = hn + W_o v ;
  // out = h1 + W_down(silu(gate) * up)(rmsnorm2_j(h1)) ; logits_j = W_dec out
  {
    const float* xr = a.h + (size_t)(rows - 1) * H;
    float ss = 0.f;
    for (int k = tid; k < H; k += NCONS) { const float v = __ldcg(xr + k); s_hn[k] = v; ss += v * v; }
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) s_part[warp] = ss;
    cons_sync();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NW; w++) t += s_part[w];
    const float sc1 = rsqrtf(t / (float)H + a.eps);
    cons_sync();                                              // s_part is reused below
    float s2 = 0.f;
    for (int k = tid; k < H; k += NCONS) { const float v = a.norm[k] * (s_hn[k] * sc1); s_hn[k] = v; s2 += v * v; }
    for (int o = 16; o; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    if (lane == 0) s_part[warp] = s2;
    cons_sync();
    t = 0.f;
#pragma unroll
    for (int w = 0; w < NW; w++) t += s_part[w];
    const float sc2 = rsqrtf(t / (float)H + a.eps);
    const int KH = H >> 1;
    for (int k = tid; k < H; k += NCONS) {
      const int q = k >> 2, dst = (q & 1) * KH + (q >> 1) * 4 + (k & 3);
      const float hv = s_hn[k] * sc2;
      for (int j = 0; j < HK; j++) sx[j * H + dst] = a.m_ln1[(size_t)j * H + k] * hv;
    }
    cons_sync();
  }
  {
    const Phase p = make_phase(a.m_v_w, H, H, HK, 1);
    float* out = a.m_v;
    Addend ad; ad.bias = a.m_v_b; ad.sb = H;
    Epi ep; ep.out = out; ep.so = H;
    consume<1, false>(p, rg, sx, H, 1, ad, ep);
  }
  HVX_GRID_BARRIER();
  for (int j = 0; j < HK; j++) stage_rows<1>(sx + j * H, a.m_v + (size_t)j * H, H, 1, H, nullptr, 0, a.eps, s_part);
  {
    const Phase p = make_phase(a.m_o_w, H, H, HK, 2);
    float* out = a.m_h1;
    Epi ep; ep.out = out; ep.so = H; ep.s_add = s_hn;
    consume<1, false>(p, rg, sx, H, 1, Addend(), ep);
  }
  HVX_GRID_BARRIER();
  for (int j = 0; j < HK; j++) stage_rows<1>(sx + j * H, a.m_h1 + (size_t)j * H, H, 1, H, a.m_ln2 + (size_t)j * H, 0, a.eps, s_part);
  {
    const Phase p = make_phase(a.m_gu_w, 2 * MI, H, HK, 3);
    float* out = a.m_act;
    Epi ep; ep.mode = EPI_SWIGLU; ep.out = out; ep.so = MI;
    consume<1, false>(p, rg, sx, H, 1, Addend(), ep);
  }
  HVX_GRID_BARRIER();
  {
    const Phase p = make_phase(a.m_down_w, H, MI, HK, 4);
    float* out = a.m_o;
    Addend ad; ad.resid = a.m_h1; ad.sr = H; ad.ld = 0;
    Epi ep; ep.out = out; ep.so = H;
    consume<1, true>(p, rg, a.m_act, MI, 1, ad, ep);
  }
  HVX_GRID_BARRIER();
  stage_rows<R>(sx, a.m_o, H, HK, H, nullptr, 0, a.eps, s_part);
  {
    const Phase p = make_phase(a.dec_w, a.V, H, 1, 5);
    float* out = a.logits; const int V = a.V;
    Epi ep; ep.out = out; ep.ld = V;
    consume<R, false>(p, rg, sx, 0, HK, Addend(), ep);
  }
#undef HVX_GRID_BARRIER
#undef HVX_MARK
}

}  // namespace fused
}  // namespace hvx
