// tcgen05 flash-style attention for the DiT estimator (see attention.cuh).
#include "attention.cuh"
#include <cstdlib>

namespace hvx {

// Tile geometry: 128 query rows per CTA, 64-key K / V^T tiles, three TMA stages.  (Four earlier generations of this kernel —
// 128-key tiles with a two-pass softmax, O in TMEM with eager rescale, 64-key tiles with a double-buffered score accumulator,
// two threads per query row — were measured against each other in round 1, profiles/r1*; only the fastest one is kept.)
constexpr int A3_Q_BYTES = 128 * 128;
constexpr int A3_K_BYTES = 64 * 128;         // 64 keys x 64 dims
constexpr int A3_V_BYTES = 64 * 128;         // 64 dims x 64 keys (V^T)
constexpr int A3_STAGES = 3;
constexpr int A3_OFF_Q = 0;
constexpr int A3_OFF_KV = A3_OFF_Q + A3_Q_BYTES;

// Warp-specialised: warp 4 is the only one that talks to the TMA unit and the tensor core (loads K / V^T tiles,
// issues Q.K^T two tiles ahead and P.V as soon as the 128 softmax threads have published P through an mbarrier); the four
// softmax warps never issue an MMA and never meet at a CTA-wide barrier, so each proceeds as soon as its own score tile
// and P buffer are ready.  P never leaves tensor memory: the softmax threads store the packed fp16 probabilities of their row with
// tcgen05.st (two 32-column buffers, so writing P_{j+1} does not wait for P.V_j) and P.V reads them as the TMEM A operand of
// tcgen05.mma — the shared-memory round trip of P (16 KB written + 32 KB of operand reads per 64-key tile) and its proxy fence were
// a third of the kernel's shared-memory traffic: 486 -> 568 TFLOP/s at T = 2298.
// Tensor memory (256 columns): S_0 [0,64) S_1 [64,128) O [128,192) P_0 [192,224) P_1 [224,256).
// K and V^T tiles travel through separate rings: a key tile is dead as soon as Q.K_j has retired, two tiles before P.V_j frees the
// value tile, so four key slots + three value slots give both loads two tile-times of TMA latency budget (one shared 3-stage ring
// refilled after P.V_j left the key tile of Q.K_{j+3} a single tile-time: the softmax warps spent 6 % of their samples waiting
// for S, profiles/r2 attention source view).
constexpr int A5_KN = 4, A5_VN = 3;
constexpr int A5_OFF_K = A3_OFF_KV;
constexpr int A5_OFF_V = A5_OFF_K + A5_KN * A3_K_BYTES;
constexpr int A5_OFF_BAR = A5_OFF_V + A5_VN * A3_V_BYTES;
constexpr int A5_OFF_XCH = A5_OFF_BAR + 256;     // [2 tile parities][2 halves][128 rows] floats: the row-pair exchange of HALVES == 2
constexpr int A5_SMEM = A5_OFF_XCH + 2048 + 1024;

// HALVES == 2 (A/B switch HVX_ATTN_HALVES=2): two threads per query row — warps w and w+4 share TMEM lanes 32*(w%4).. and take 32 of
// the 64 score columns, 16 of the 32 P columns and 32 of the 64 output dims each; the row maximum is exchanged through shared
// memory under a 64-thread named barrier per tile.  Measured: no faster than one thread per row (the exchange costs what the extra
// warps hide), slower once P stays in tensor memory.
template <int HALVES>
__global__ void __launch_bounds__(HALVES * 128 + 32, 2)
dit_attention_v5_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, int k_col0, AttnArgs a) {
  constexpr int MMA_WARP = HALVES * 4, COLS = 64 / HALVES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + A5_OFF_BAR);
  uint64_t* k_full = q_full + 1;             // [4]  TMA -> MMA warp
  uint64_t* v_full = k_full + A5_KN;         // [3]
  uint64_t* s_full = v_full + A5_VN;         // [2]  tensor core -> softmax warps
  uint64_t* p_ready = s_full + 2;            // [2]  softmax warps (128 arrivals) -> MMA warp
  uint64_t* o_done = p_ready + 2;            // [2]  P.V_j retired (parity buffers by j&1): frees P[j&1], K/V stage, O for rescale
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = a.T;                                   // key rows per batch row
  const int Tq = a.Tq > 0 ? a.Tq : T;                  // query rows per batch row (windowed call: Tq < T, queries are frames q_pos0 ..)

  if (tid == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < A5_KN; i++) tc::mbar_init(&k_full[i], 1);
    for (int i = 0; i < A5_VN; i++) tc::mbar_init(&v_full[i], 1);
    for (int i = 0; i < 2; i++) { tc::mbar_init(&s_full[i], 1); tc::mbar_init(&p_ready[i], HALVES * 128); tc::mbar_init(&o_done[i], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  pdl_wait();                                          // first global read below (klen); the prologue above overlapped the predecessor
  const int Tk = a.klen ? a.klen[b] : (a.tk > 0 ? a.tk : T);     // keys of this batch row that exist (ragged group: the rest of the slab is padding)
  const int klim_tile = a.chunk > 0 ? min(Tk, ((a.q_pos0 + min(q0 + 127, Tq - 1)) / a.chunk + 1) * a.chunk) : Tk;
  const int nkv = (klim_tile + 63) / 64;
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;
  const uint32_t idesc = a.f16 ? tc::umma_idesc_f16(128, 64) : tc::umma_idesc_bf16(128, 64);

  if (warp == MMA_WARP) {
    // ---------------- TMA + tensor-core warp (one elected lane)
    if (lane == 0) {
      auto load_k = [&](int j) {
        const int st = j % A5_KN;
        tc::mbar_expect_tx(&k_full[st], A3_K_BYTES);
        tc::tma_load_2d(smem + A5_OFF_K + st * A3_K_BYTES, &tm_k, &k_full[st], k_col0 + h * 64, b * T + j * 64);
      };
      auto load_v = [&](int j) {
        const int st = j % A5_VN;
        tc::mbar_expect_tx(&v_full[st], A3_V_BYTES);
        tc::tma_load_2d(smem + A5_OFF_V + st * A3_V_BYTES, &tm_v, &v_full[st], j * 64, (b * a.heads + h) * 64);
      };
      auto issue_qk = [&](int j) {
        const int st = j % A5_KN;
        tc::mbar_wait(&k_full[st], (j / A5_KN) & 1);
        tc::tc_fence_after();
        const uint64_t dq = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_Q));
        const uint64_t dk = tc::umma_desc_k128(tc::smem_u32(smem + A5_OFF_K + st * A3_K_BYTES));
#pragma unroll
        for (int k = 0; k < 4; k++) tc::umma_f16(tmem_base + (uint32_t)((j & 1) * 64), dq + 2 * k, dk + 2 * k, idesc, k ? 1u : 0u);
        tc::umma_commit(&s_full[j & 1]);
      };
      tc::mbar_expect_tx(q_full, A3_Q_BYTES);
      tc::tma_load_2d(smem + A3_OFF_Q, &tm_q, q_full, h * 64, b * Tq + q0);
      for (int j = 0; j < A5_KN; j++) {
        if (j < nkv) load_k(j);
        if (j < nkv && j < A5_VN) load_v(j);
      }
      tc::mbar_wait(q_full, 0);
      issue_qk(0);
      if (nkv > 1) issue_qk(1);
      for (int j = 0; j < nkv; j++) {
        // P_j published: S_j has been read, so S[j&1] may be overwritten by Q.K_{j+2} and — Q.K_j having retired — key slot j%4
        // may take tile j+4
        tc::mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
        tc::tc_fence_after();
        if (j + A5_KN < nkv) load_k(j + A5_KN);
        if (j >= 1 && j + 2 < nkv) {             // value slot (j+2)%3 == (j-1)%3 is free once P.V_{j-1} has retired (long ago)
          tc::mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          load_v(j + 2);
        }
        const int sv = j % A5_VN;
        tc::mbar_wait(&v_full[sv], (j / A5_VN) & 1);
        tc::tc_fence_after();
        const uint32_t tp = tmem_base + 192 + (uint32_t)((j & 1) * 32);      // P_j: 64 fp16 keys = 32 packed columns per row
        const uint64_t dv = tc::umma_desc_k128(tc::smem_u32(smem + A5_OFF_V + sv * A3_V_BYTES));
#pragma unroll
        for (int k = 0; k < 4; k++) tc::umma_f16_ts(tmem_o, tp + 8 * k, dv + 2 * k, idesc, (j | k) ? 1u : 0u);
        tc::umma_commit(&o_done[j & 1]);
        if (j + 2 < nkv) issue_qk(j + 2);
      }
    }
  } else {
    // ---------------- softmax warps: thread == query row == TMEM lane (HALVES == 2: two threads per row, COLS columns each)
    const int half = warp >> 2, wq = warp & 3, row = wq * 32 + lane, col_off = half * COLS;
    const int row_in_batch = q0 + row;
    const int klim_row = a.chunk > 0 ? min(Tk, ((a.q_pos0 + row_in_batch) / a.chunk + 1) * a.chunk) : Tk;
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const float sc = 0.125f * 1.4426950408889634f;     // dim_head^-0.5 * log2(e)
    float* xch = reinterpret_cast<float*>(smem + A5_OFF_XCH);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; j++) {
      tc::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc::tc_fence_after();
      uint32_t sreg[COLS];
#pragma unroll
      for (int c = 0; c < COLS; c += 32) tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)((j & 1) * 64 + col_off + c), sreg + c);
      tc::tmem_ld_wait();
      const int kbase = j * 64 + col_off;
      float mt[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      // raw scores stay in registers: the max is taken on them (sc > 0) and the scale rides in the ex2 argument as one FFMA
      if (kbase + COLS <= klim_row) {
#pragma unroll
        for (int i = 0; i < COLS; i++) mt[i & 3] = fmaxf(mt[i & 3], __uint_as_float(sreg[i]));
      } else {
#pragma unroll
        for (int i = 0; i < COLS; i++) {
          const float s = (kbase + i < klim_row) ? __uint_as_float(sreg[i]) : -INFINITY;
          sreg[i] = __float_as_uint(s);
          mt[i & 3] = fmaxf(mt[i & 3], s);
        }
      }
      float m_raw = fmaxf(fmaxf(mt[0], mt[1]), fmaxf(mt[2], mt[3]));
      if (HALVES == 2) {
        // the other half of the row: buffers alternate with the tile parity, so one barrier per tile orders both the read of
        // this tile's value and the overwrite two tiles later
        float* x = xch + (j & 1) * 256;
        x[half * 128 + row] = m_raw;
        asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
        m_raw = fmaxf(m_raw, x[(half ^ 1) * 128 + row]);
      }
      const float m_tile = m_raw * sc;
      const bool need = m_tile > m_run + 8.0f;
      if (__any_sync(0xffffffffu, need)) {
        const float m_new = need ? m_tile : m_run;
        const float alpha = need ? tc::ex2(m_run - m_new) : 1.0f;
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
          // O may only be rescaled once P.V_{j-1} has retired (and nothing newer can be in flight: P.V_j needs this P)
          tc::mbar_wait(&o_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
          tc::tc_fence_after();
#pragma unroll
          for (int c0 = 0; c0 < COLS; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem_o + lane_off + (uint32_t)(col_off + c0), v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tc::tmem_st_32x32(tmem_o + lane_off + (uint32_t)(col_off + c0), v);
          }
          tc::tmem_st_wait();
        }
      }
      // P buffer j&1 was last read by P.V_{j-2}
      if (j >= 2) tc::mbar_wait(&o_done[j & 1], ((j - 2) >> 1) & 1);
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t pk[COLS / 2];
#pragma unroll
      for (int i = 0; i < COLS; i += 2) {
        const float p0 = tc::ex2(fmaf(__uint_as_float(sreg[i]), sc, -m_run)), p1 = tc::ex2(fmaf(__uint_as_float(sreg[i + 1]), sc, -m_run));
        ps[(i >> 1) & 3] += p0 + p1;
        pk[i >> 1] = tc::pack16(p0, p1, a.f16);
      }
      // P stays in tensor memory: this thread's row (lane), keys (2c, 2c+1) packed in column c — the K-major A operand of P.V
      const uint32_t tp = tmem_base + lane_off + 192 + (uint32_t)((j & 1) * 32 + col_off / 2);
      if (HALVES == 1) tc::tmem_st_32x32(tp, pk); else tc::tmem_st_32x16(tp, pk);
      tc::tmem_st_wait();
      l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);
      tc::tc_fence_before();                   // orders this thread's TMEM accesses before the MMA warp's next issue
      tc::mbar_arrive(&p_ready[j & 1]);
    }
    tc::mbar_wait(&o_done[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
    tc::tc_fence_after();
    if (HALVES == 2) {                         // row sum = both halves (the parity buffer of tile nkv is free: see above)
      float* x = xch + (nkv & 1) * 256;
      x[half * 128 + row] = l_run;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
      l_run += x[(half ^ 1) * 128 + row];
    }
    uint32_t v[COLS];
#pragma unroll
    for (int c = 0; c < COLS; c += 32) tc::tmem_ld_32x32(tmem_o + lane_off + (uint32_t)(col_off + c), v + c);
    tc::tmem_ld_wait();
    if (row_in_batch < Tq && a.lo_off) {
      // split precision (flow parity mode): hi at [col], lo = y - hi at [lo_off + col]; 16-byte stores like the plain path
      // (element-wise 2-byte stores made this epilogue as long as the whole key loop: 249 vs 479 TFLOP/s)
      const float inv = 1.0f / l_run;
      uint16_t* o = reinterpret_cast<uint16_t*>(a.out) + (size_t)(b * Tq + row_in_batch) * (2 * a.ld_out) + h * 64 + col_off;
#pragma unroll
      for (int i = 0; i < COLS; i += 8) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; j++) y[j] = __uint_as_float(v[i + j]) * inv;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          hi[j] = tc::pack16(y[2 * j], y[2 * j + 1], a.f16);
          lo[j] = tc::pack_lo16(y[2 * j], y[2 * j + 1], hi[j], a.f16);
        }
        *reinterpret_cast<uint4*>(o + i) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(o + a.lo_off + i) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    } else if (row_in_batch < Tq) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* o = a.out + (size_t)(b * Tq + row_in_batch) * a.ld_out + h * 64 + col_off;
#pragma unroll
      for (int i = 0; i < COLS; i += 8) {
        uint4 pk;
        pk.x = tc::pack16(__uint_as_float(v[i]) * inv, __uint_as_float(v[i + 1]) * inv, a.f16);
        pk.y = tc::pack16(__uint_as_float(v[i + 2]) * inv, __uint_as_float(v[i + 3]) * inv, a.f16);
        pk.z = tc::pack16(__uint_as_float(v[i + 4]) * inv, __uint_as_float(v[i + 5]) * inv, a.f16);
        pk.w = tc::pack16(__uint_as_float(v[i + 6]) * inv, __uint_as_float(v[i + 7]) * inv, a.f16);
        *reinterpret_cast<uint4*>(o + i) = pk;
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tc::tmem_dealloc(tmem_base, 256); }
}

hvx_status dit_attention(hvx_engine* e, cudaStream_t st, const __nv_bfloat16* qk, int ld_qk, int k_col0,
                         const __nv_bfloat16* vt, int vt_ld, const AttnArgs& a) {
  CUtensorMap tq, tk64, tv;
  const uint64_t rows = (uint64_t)a.n_batch * a.T;
  const int Tq = a.Tq > 0 ? a.Tq : a.T;
  HVX_CHECK(a.q_pos0 >= 0 && a.q_pos0 + Tq <= a.T, HVX_ERR_ARG, "attention: query window [%d, %d) outside the %d key rows", a.q_pos0, a.q_pos0 + Tq, a.T);
  HVX_CHECK(make_tmap_bf16_2d(&tq, qk, (uint64_t)a.n_batch * Tq, ld_qk, ld_qk, 128, 64), HVX_ERR_CUDA, "attention: tensor map Q failed");
  if (a.k_ptr) {                                       // windowed call: keys live in their own (cache) matrix
    k_col0 = 0;
    HVX_CHECK(make_tmap_bf16_2d(&tk64, a.k_ptr, rows, a.ld_k, a.ld_k, 64, 64), HVX_ERR_CUDA, "attention: tensor map K failed");
  } else {
    HVX_CHECK(make_tmap_bf16_2d(&tk64, qk, rows, ld_qk, ld_qk, 64, 64), HVX_ERR_CUDA, "attention: tensor map K failed");
  }
  HVX_CHECK(make_tmap_bf16_2d(&tv, vt, (uint64_t)a.n_batch * a.heads * 64, vt_ld, vt_ld, 64, 64), HVX_ERR_CUDA,
            "attention: tensor map V^T failed");
  static bool attr_set = false;
  if (!attr_set) {
    HVX_CUDA(cudaFuncSetAttribute(dit_attention_v5_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, A5_SMEM));
    HVX_CUDA(cudaFuncSetAttribute(dit_attention_v5_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, A5_SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(Tq, 128), a.heads, a.n_batch);
  ProfScope prof_scope(&e->prof, st, PROF_ATTN, a.work > 0 ? a.work : 4.0 * a.n_batch * a.heads * (double)a.T * a.T * 64.0 * (a.chunk > 0 ? 0.5 : 1.0));
  // one thread per query row is the faster form once P stays in tensor memory (568 vs 526 TFLOP/s at T = 2298); HVX_ATTN_HALVES=2
  // selects two threads per row
  static const int halves = getenv("HVX_ATTN_HALVES") ? atoi(getenv("HVX_ATTN_HALVES")) : 1;
  if (halves == 2) HVX_CUDA(launch_pdl_ex(dit_attention_v5_kernel<2>, grid, dim3(288), (size_t)A5_SMEM, st, 1, tq, tk64, tv, k_col0, a));
  else HVX_CUDA(launch_pdl_ex(dit_attention_v5_kernel<1>, grid, dim3(160), (size_t)A5_SMEM, st, 1, tq, tk64, tv, k_col0, a));
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

}  // namespace hvx

using namespace hvx;

// Diagnostic entry (tests/test_attention_gpu.py): qk = [B*T][2*H*64] (q | k), vt = [B*H*64][vt_ld]
extern "C" hvx_status hvx_attention_bf16(hvx_engine* e, const void* qk, const void* vt, int vt_ld, void* out, int B, int T,
                                         int H, int chunk, void* stream) {
  HVX_CHECK(e, HVX_ERR_ARG, "null engine");
  AttnArgs a;
  a.T = T; a.heads = H; a.n_batch = B; a.chunk = chunk & 0xffff; a.f16 = (chunk >> 16) & 1;   // bit 16 of chunk: fp16 operands
  a.ld_out = H * 64; a.out = (__nv_bfloat16*)out;
  return dit_attention(e, (cudaStream_t)stream, (const __nv_bfloat16*)qk, 2 * H * 64, H * 64, (const __nv_bfloat16*)vt,
                       vt_ld, a);
}
