// tcgen05 flash-style attention for the DiT estimator (see attention.cuh).
#include "attention.cuh"
#include <cstdlib>

namespace hvx {

constexpr int AT_Q_BYTES = 128 * 128;        // 128 rows x 64 bf16
constexpr int AT_K_BYTES = 128 * 128;        // 128 keys x 64 bf16
constexpr int AT_V_BYTES = 2 * 64 * 128;     // two K-halves of V^T: 64 dims x 64 keys each
constexpr int AT_P_BYTES = 2 * 128 * 128;    // two 64-key atoms of P
constexpr int AT_OFF_Q = 0;
constexpr int AT_OFF_K = AT_OFF_Q + AT_Q_BYTES;
constexpr int AT_OFF_V = AT_OFF_K + 2 * AT_K_BYTES;
constexpr int AT_OFF_P = AT_OFF_V + 2 * AT_V_BYTES;
constexpr int AT_OFF_BAR = AT_OFF_P + AT_P_BYTES;
constexpr int AT_SMEM = AT_OFF_BAR + 128 + 1024;
constexpr uint32_t AT_TMEM_COLS = 256;

__global__ void __launch_bounds__(128)
dit_attention_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, int k_col0, AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t* kv_full = q_full + 1;    // [2]
  uint64_t* s_full = kv_full + 2;
  uint64_t* o_full = s_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = a.T;
  const int row_in_batch = q0 + tid;
  const int klim_row = a.chunk > 0 ? min(T, (row_in_batch / a.chunk + 1) * a.chunk) : T;
  const int klim_tile = a.chunk > 0 ? min(T, ((min(q0 + 127, T - 1)) / a.chunk + 1) * a.chunk) : T;
  const int nkv = (klim_tile + 127) / 128;

  if (tid == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(q_full, 1); tc::mbar_init(&kv_full[0], 1); tc::mbar_init(&kv_full[1], 1);
    tc::mbar_init(s_full, 1); tc::mbar_init(o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, AT_TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;
  const uint32_t tmem_o = tmem_s + 128;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  auto load_kv = [&](int j) {
    const int bsel = j & 1;
    tc::mbar_expect_tx(&kv_full[bsel], AT_K_BYTES + AT_V_BYTES);
    tc::tma_load_2d(smem + AT_OFF_K + bsel * AT_K_BYTES, &tm_k, &kv_full[bsel], k_col0 + h * 64, b * T + j * 128);
    uint8_t* sv = smem + AT_OFF_V + bsel * AT_V_BYTES;
    tc::tma_load_2d(sv, &tm_v, &kv_full[bsel], j * 128, (b * a.heads + h) * 64);
    tc::tma_load_2d(sv + 64 * 128, &tm_v, &kv_full[bsel], j * 128 + 64, (b * a.heads + h) * 64);
  };
  if (tid == 0) {
    tc::mbar_expect_tx(q_full, AT_Q_BYTES);
    tc::tma_load_2d(smem + AT_OFF_Q, &tm_q, q_full, h * 64, b * T + q0);
    load_kv(0);
  }

  const uint32_t idesc_s = a.f16 ? tc::umma_idesc_f16(128, 128) : tc::umma_idesc_bf16(128, 128);
  const uint32_t idesc_o = a.f16 ? tc::umma_idesc_f16(128, 64) : tc::umma_idesc_bf16(128, 64);
  const float sc = 0.125f * 1.4426950408889634f;     // dim_head^-0.5 * log2(e)
  float m_run = -INFINITY, l_run = 0.f;
  float o_acc[64];
#pragma unroll
  for (int i = 0; i < 64; i++) o_acc[i] = 0.f;
  uint8_t* sp = smem + AT_OFF_P;

  for (int j = 0; j < nkv; j++) {
    const int bsel = j & 1;
    if (tid == 0) {
      if (j == 0) tc::mbar_wait(q_full, 0);
      tc::mbar_wait(&kv_full[bsel], (j >> 1) & 1);
      tc::tc_fence_after();
      const uint64_t dq = tc::umma_desc_k128(tc::smem_u32(smem + AT_OFF_Q));
      const uint64_t dk = tc::umma_desc_k128(tc::smem_u32(smem + AT_OFF_K + bsel * AT_K_BYTES));
#pragma unroll
      for (int k = 0; k < 4; k++) tc::umma_f16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
      tc::umma_commit(s_full);
      if (j + 1 < nkv) load_kv(j + 1);
    }
    __syncwarp();
    tc::mbar_wait(s_full, j & 1);
    tc::tc_fence_after();

    // ---- online softmax on this thread's row; P written straight into the swizzled A-operand tile
    float m_new = m_run;
    const int kbase = j * 128;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_s + lane_off + (uint32_t)c0, v);
        tc::tmem_ld_wait();
        if (pass == 0) {
#pragma unroll
          for (int i = 0; i < 32; i++) {
            const float s = (kbase + c0 + i < klim_row) ? __uint_as_float(v[i]) * sc : -INFINITY;
            m_new = fmaxf(m_new, s);
          }
        } else {
          float psum = 0.f;
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float s0 = (kbase + c0 + i < klim_row) ? __uint_as_float(v[i]) * sc : -INFINITY;
            const float s1 = (kbase + c0 + i + 1 < klim_row) ? __uint_as_float(v[i + 1]) * sc : -INFINITY;
            const float p0 = exp2f(s0 - m_new), p1 = exp2f(s1 - m_new);
            psum += p0 + p1;
            pk[i >> 1] = tc::pack16(p0, p1, a.f16);
          }
          l_run += psum;
          // 32 keys = 4 chunks of 16 B in atom (c0/64), chunk index ((c0%64)/8 + q) ^ (row & 7)
          uint8_t* rowp = sp + (c0 >> 6) * (128 * 128) + tid * 128;
          const int cb = (c0 & 63) >> 3;
#pragma unroll
          for (int qd = 0; qd < 4; qd++) {
            uint4 val = make_uint4(pk[4 * qd], pk[4 * qd + 1], pk[4 * qd + 2], pk[4 * qd + 3]);
            *reinterpret_cast<uint4*>(rowp + (((cb + qd) ^ (tid & 7)) << 4)) = val;
          }
        }
      }
      if (pass == 0) {
        // rescale the running sum before adding this tile's probabilities
        const float alpha = exp2f(m_run - m_new);      // m_run=-inf on the first tile -> 0
        l_run *= alpha;
#pragma unroll
        for (int i = 0; i < 64; i++) o_acc[i] *= alpha;
        m_run = m_new;
      }
    }
    tc::fence_proxy_async();          // make the generic-proxy P stores visible to the tensor core
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t pv = tc::smem_u32(smem + AT_OFF_V + bsel * AT_V_BYTES);
      const uint32_t pp = tc::smem_u32(sp);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const uint64_t dp = tc::umma_desc_k128(pp + half * (128 * 128));
        const uint64_t dv = tc::umma_desc_k128(pv + half * (64 * 128));
#pragma unroll
        for (int k = 0; k < 4; k++) tc::umma_f16(tmem_o, dp + 2 * k, dv + 2 * k, idesc_o, (half | k) ? 1u : 0u);
      }
      tc::umma_commit(o_full);
    }
    __syncwarp();
    tc::mbar_wait(o_full, j & 1);
    tc::tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(tmem_o + lane_off + (uint32_t)c0, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i++) o_acc[c0 + i] += __uint_as_float(v[i]);
    }
    tc::tc_fence_before();
  }

  if (row_in_batch < T) {
    const float inv = 1.0f / l_run;
    __nv_bfloat16* o = a.out + (size_t)(b * T + row_in_batch) * a.ld_out + h * 64;
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
      uint4 pk;
      pk.x = tc::pack16(o_acc[i] * inv, o_acc[i + 1] * inv, a.f16); pk.y = tc::pack16(o_acc[i + 2] * inv, o_acc[i + 3] * inv, a.f16);
      pk.z = tc::pack16(o_acc[i + 4] * inv, o_acc[i + 5] * inv, a.f16); pk.w = tc::pack16(o_acc[i + 6] * inv, o_acc[i + 7] * inv, a.f16);
      *reinterpret_cast<uint4*>(o + i) = pk;
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tc::tmem_dealloc(tmem_s, AT_TMEM_COLS); }
}

// v2: one pass over S (the 128 scores of a row are held in registers), O stays in TMEM across KV tiles and is rescaled
// in place only when the running maximum grows by more than 2^8 (probabilities may exceed 1 by that factor; fp32 row sum
// and the final 1/l absorb it), so a tile costs one TMEM read of S instead of two plus an O read-back.
__global__ void __launch_bounds__(128)
dit_attention_v2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, int k_col0, AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + AT_OFF_BAR);
  uint64_t* kv_full = q_full + 1;    // [2]
  uint64_t* s_full = kv_full + 2;
  uint64_t* o_full = s_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = a.T;
  const int row_in_batch = q0 + tid;
  const int klim_row = a.chunk > 0 ? min(T, (row_in_batch / a.chunk + 1) * a.chunk) : T;
  const int klim_tile = a.chunk > 0 ? min(T, ((min(q0 + 127, T - 1)) / a.chunk + 1) * a.chunk) : T;
  const int nkv = (klim_tile + 127) / 128;

  if (tid == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(q_full, 1); tc::mbar_init(&kv_full[0], 1); tc::mbar_init(&kv_full[1], 1);
    tc::mbar_init(s_full, 1); tc::mbar_init(o_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, AT_TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_s = *tmem_slot;
  const uint32_t tmem_o = tmem_s + 128;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;

  auto load_kv = [&](int j) {
    const int bsel = j & 1;
    tc::mbar_expect_tx(&kv_full[bsel], AT_K_BYTES + AT_V_BYTES);
    tc::tma_load_2d(smem + AT_OFF_K + bsel * AT_K_BYTES, &tm_k, &kv_full[bsel], k_col0 + h * 64, b * T + j * 128);
    uint8_t* sv = smem + AT_OFF_V + bsel * AT_V_BYTES;
    tc::tma_load_2d(sv, &tm_v, &kv_full[bsel], j * 128, (b * a.heads + h) * 64);
    tc::tma_load_2d(sv + 64 * 128, &tm_v, &kv_full[bsel], j * 128 + 64, (b * a.heads + h) * 64);
  };
  if (tid == 0) {
    tc::mbar_expect_tx(q_full, AT_Q_BYTES);
    tc::tma_load_2d(smem + AT_OFF_Q, &tm_q, q_full, h * 64, b * T + q0);
    load_kv(0);
  }
  const uint32_t idesc_s = a.f16 ? tc::umma_idesc_f16(128, 128) : tc::umma_idesc_bf16(128, 128);
  const uint32_t idesc_o = a.f16 ? tc::umma_idesc_f16(128, 64) : tc::umma_idesc_bf16(128, 64);
  const float sc = 0.125f * 1.4426950408889634f;     // dim_head^-0.5 * log2(e)
  float m_run = -INFINITY, l_run = 0.f;
  uint8_t* sp = smem + AT_OFF_P;

  for (int j = 0; j < nkv; j++) {
    const int bsel = j & 1;
    if (tid == 0) {
      if (j == 0) tc::mbar_wait(q_full, 0);
      tc::mbar_wait(&kv_full[bsel], (j >> 1) & 1);
      tc::tc_fence_after();
      const uint64_t dq = tc::umma_desc_k128(tc::smem_u32(smem + AT_OFF_Q));
      const uint64_t dk = tc::umma_desc_k128(tc::smem_u32(smem + AT_OFF_K + bsel * AT_K_BYTES));
#pragma unroll
      for (int k = 0; k < 4; k++) tc::umma_f16(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, k ? 1u : 0u);
      tc::umma_commit(s_full);
    }
    __syncwarp();
    tc::mbar_wait(s_full, j & 1);      // S_j ready; tcgen05 ops retire in order, so P.V of tile j-1 has retired too
    tc::tc_fence_after();
    if (tid == 0 && j + 1 < nkv) load_kv(j + 1);       // its K/V^T/P buffers are free now

    uint32_t sreg[128];
    tc::tmem_ld_32x32(tmem_s + lane_off + 0, sreg);
    tc::tmem_ld_32x32(tmem_s + lane_off + 32, sreg + 32);
    tc::tmem_ld_32x32(tmem_s + lane_off + 64, sreg + 64);
    tc::tmem_ld_32x32(tmem_s + lane_off + 96, sreg + 96);
    tc::tmem_ld_wait();
    const int kbase = j * 128;
    // 8 independent max chains (a single 128-long fmax chain is ~500 cycles of pure latency per tile)
    float mt[8];
#pragma unroll
    for (int i = 0; i < 8; i++) mt[i] = -INFINITY;
    if (kbase + 128 <= klim_row) {                      // whole tile visible: no masking
#pragma unroll
      for (int i = 0; i < 128; i++) {
        const float s = __uint_as_float(sreg[i]) * sc;
        sreg[i] = __float_as_uint(s);
        mt[i & 7] = fmaxf(mt[i & 7], s);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 128; i++) {
        const float s = (kbase + i < klim_row) ? __uint_as_float(sreg[i]) * sc : -INFINITY;
        sreg[i] = __float_as_uint(s);
        mt[i & 7] = fmaxf(mt[i & 7], s);
      }
    }
    const float m_tile = fmaxf(fmaxf(fmaxf(mt[0], mt[1]), fmaxf(mt[2], mt[3])), fmaxf(fmaxf(mt[4], mt[5]), fmaxf(mt[6], mt[7])));
    // lazy rescale, decided per warp so the TMEM accesses stay warp-uniform
    const bool need = m_tile > m_run + 8.0f;
    if (__any_sync(0xffffffffu, need)) {
      const float m_new = need ? m_tile : m_run;
      const float alpha = need ? tc::ex2(m_run - m_new) : 1.0f;      // first tile: ex2(-inf) = 0
      m_run = m_new;
      l_run *= alpha;
      if (j > 0) {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32(tmem_o + lane_off + (uint32_t)c0, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tc::tmem_st_32x32(tmem_o + lane_off + (uint32_t)c0, v);
        }
        tc::tmem_st_wait();
      }
    }
    float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = tc::ex2(__uint_as_float(sreg[c0 + i]) - m_run), p1 = tc::ex2(__uint_as_float(sreg[c0 + i + 1]) - m_run);
        ps[(i >> 1) & 3] += p0 + p1;
        pk[i >> 1] = tc::pack16(p0, p1, a.f16);
      }
      uint8_t* rowp = sp + (c0 >> 6) * (128 * 128) + tid * 128;
      const int cb = (c0 & 63) >> 3;
#pragma unroll
      for (int qd = 0; qd < 4; qd++) {
        uint4 val = make_uint4(pk[4 * qd], pk[4 * qd + 1], pk[4 * qd + 2], pk[4 * qd + 3]);
        *reinterpret_cast<uint4*>(rowp + (((cb + qd) ^ (tid & 7)) << 4)) = val;
      }
    }
    l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const uint32_t pv = tc::smem_u32(smem + AT_OFF_V + bsel * AT_V_BYTES);
      const uint32_t pp = tc::smem_u32(sp);
#pragma unroll
      for (int half = 0; half < 2; half++) {
        const uint64_t dp = tc::umma_desc_k128(pp + half * (128 * 128));
        const uint64_t dv = tc::umma_desc_k128(pv + half * (64 * 128));
#pragma unroll
        for (int k = 0; k < 4; k++) tc::umma_f16(tmem_o, dp + 2 * k, dv + 2 * k, idesc_o, (j | half | k) ? 1u : 0u);
      }
      if (j == nkv - 1) tc::umma_commit(o_full);
    }
    __syncwarp();
  }
  tc::mbar_wait(o_full, 0);
  tc::tc_fence_after();
  {
    uint32_t v[64];
    tc::tmem_ld_32x32(tmem_o + lane_off, v);
    tc::tmem_ld_32x32(tmem_o + lane_off + 32, v + 32);
    tc::tmem_ld_wait();
    if (row_in_batch < T) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* o = a.out + (size_t)(b * T + row_in_batch) * a.ld_out + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
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
  if (warp == 0) { __syncwarp(); tc::tmem_dealloc(tmem_s, AT_TMEM_COLS); }
}

// v3: 64-key tiles with the score accumulator double-buffered in TMEM (2 x 64 + 64 columns -> 256 allocated, two CTAs
// per SM as before).  Q.K^T of tile j+1 is issued before the softmax of tile j starts and P.V of tile j runs under the
// softmax of tile j+1, so the threads never wait for the tensor pipe (v1/v2 spent ~1/3 of their samples spinning on
// s_full); K / V^T tiles stream through a 3-stage TMA ring.
constexpr int A3_Q_BYTES = 128 * 128;
constexpr int A3_K_BYTES = 64 * 128;         // 64 keys x 64 dims
constexpr int A3_V_BYTES = 64 * 128;         // 64 dims x 64 keys (V^T)
constexpr int A3_P_BYTES = 128 * 128;        // 128 rows x 64 keys
constexpr int A3_STAGES = 3;
constexpr int A3_OFF_Q = 0;
constexpr int A3_OFF_KV = A3_OFF_Q + A3_Q_BYTES;
constexpr int A3_OFF_P = A3_OFF_KV + A3_STAGES * (A3_K_BYTES + A3_V_BYTES);
constexpr int A3_OFF_BAR = A3_OFF_P + A3_P_BYTES;
constexpr int A3_SMEM = A3_OFF_BAR + 128 + 2048 + 1024;   // barriers + v4's pair-exchange scratch + alignment slack

__global__ void __launch_bounds__(128)
dit_attention_v3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, int k_col0, AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + A3_OFF_BAR);
  uint64_t* kv_full = q_full + 1;            // [3]
  uint64_t* s_full = kv_full + A3_STAGES;    // [2]
  uint64_t* o_done = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = a.T;
  const int row_in_batch = q0 + tid;
  const int klim_row = a.chunk > 0 ? min(T, (row_in_batch / a.chunk + 1) * a.chunk) : T;
  const int klim_tile = a.chunk > 0 ? min(T, ((min(q0 + 127, T - 1)) / a.chunk + 1) * a.chunk) : T;
  const int nkv = (klim_tile + 63) / 64;

  if (tid == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < A3_STAGES; i++) tc::mbar_init(&kv_full[i], 1);
    tc::mbar_init(&s_full[0], 1); tc::mbar_init(&s_full[1], 1); tc::mbar_init(o_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc = a.f16 ? tc::umma_idesc_f16(128, 64) : tc::umma_idesc_bf16(128, 64);

  auto load_kv = [&](int j) {
    const int st = j % A3_STAGES;
    uint8_t* sk = smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES);
    tc::mbar_expect_tx(&kv_full[st], A3_K_BYTES + A3_V_BYTES);
    tc::tma_load_2d(sk, &tm_k, &kv_full[st], k_col0 + h * 64, b * T + j * 64);
    tc::tma_load_2d(sk + A3_K_BYTES, &tm_v, &kv_full[st], j * 64, (b * a.heads + h) * 64);
  };
  auto issue_qk = [&](int j) {                 // S[j & 1] = Q K_j^T
    const int st = j % A3_STAGES;
    tc::mbar_wait(&kv_full[st], (j / A3_STAGES) & 1);
    tc::tc_fence_after();
    const uint64_t dq = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_Q));
    const uint64_t dk = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES)));
#pragma unroll
    for (int k = 0; k < 4; k++) tc::umma_f16(tmem_base + (uint32_t)((j & 1) * 64), dq + 2 * k, dk + 2 * k, idesc, k ? 1u : 0u);
    tc::umma_commit(&s_full[j & 1]);
  };
  if (tid == 0) {
    tc::mbar_expect_tx(q_full, A3_Q_BYTES);
    tc::tma_load_2d(smem + A3_OFF_Q, &tm_q, q_full, h * 64, b * T + q0);
    load_kv(0);
    if (nkv > 1) load_kv(1);
    tc::mbar_wait(q_full, 0);
    issue_qk(0);
  }
  const float sc = 0.125f * 1.4426950408889634f;     // dim_head^-0.5 * log2(e)
  float m_run = -INFINITY, l_run = 0.f;
  uint8_t* sp = smem + A3_OFF_P;

  for (int j = 0; j < nkv; j++) {
    if (tid == 0 && j + 1 < nkv) issue_qk(j + 1);      // S[(j+1)&1] was drained before the barrier of iteration j-1
    __syncwarp();
    tc::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
    tc::tc_fence_after();
    uint32_t sreg[64];
    tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)((j & 1) * 64), sreg);
    tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)((j & 1) * 64 + 32), sreg + 32);
    tc::tmem_ld_wait();
    const int kbase = j * 64;
    float mt[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (kbase + 64 <= klim_row) {
#pragma unroll
      for (int i = 0; i < 64; i++) { const float s = __uint_as_float(sreg[i]) * sc; sreg[i] = __float_as_uint(s); mt[i & 3] = fmaxf(mt[i & 3], s); }
    } else {
#pragma unroll
      for (int i = 0; i < 64; i++) {
        const float s = (kbase + i < klim_row) ? __uint_as_float(sreg[i]) * sc : -INFINITY;
        sreg[i] = __float_as_uint(s);
        mt[i & 3] = fmaxf(mt[i & 3], s);
      }
    }
    const float m_tile = fmaxf(fmaxf(mt[0], mt[1]), fmaxf(mt[2], mt[3]));
    // P.V of tile j-1 must have retired before P is overwritten / O is rescaled / its K,V stage is refilled
    if (j > 0) { tc::mbar_wait(o_done, (j - 1) & 1); tc::tc_fence_after(); }
    if (tid == 0 && j + 2 < nkv) load_kv(j + 2);       // stage (j+2)%3 == (j-1)%3 is free now
    const bool need = m_tile > m_run + 8.0f;
    if (__any_sync(0xffffffffu, need)) {
      const float m_new = need ? m_tile : m_run;
      const float alpha = need ? tc::ex2(m_run - m_new) : 1.0f;
      m_run = m_new;
      l_run *= alpha;
      if (j > 0) {
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 32) {
          uint32_t v[32];
          tc::tmem_ld_32x32(tmem_o + lane_off + (uint32_t)c0, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tc::tmem_st_32x32(tmem_o + lane_off + (uint32_t)c0, v);
        }
        tc::tmem_st_wait();
      }
    }
    float ps[4] = {0.f, 0.f, 0.f, 0.f};
    uint8_t* rowp = sp + tid * 128;
#pragma unroll
    for (int c0 = 0; c0 < 64; c0 += 32) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = tc::ex2(__uint_as_float(sreg[c0 + i]) - m_run), p1 = tc::ex2(__uint_as_float(sreg[c0 + i + 1]) - m_run);
        ps[(i >> 1) & 3] += p0 + p1;
        pk[i >> 1] = tc::pack16(p0, p1, a.f16);
      }
      const int cb = c0 >> 3;
#pragma unroll
      for (int qd = 0; qd < 4; qd++) {
        uint4 val = make_uint4(pk[4 * qd], pk[4 * qd + 1], pk[4 * qd + 2], pk[4 * qd + 3]);
        *reinterpret_cast<uint4*>(rowp + (((cb + qd) ^ (tid & 7)) << 4)) = val;
      }
    }
    l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const int st = j % A3_STAGES;
      const uint64_t dp = tc::umma_desc_k128(tc::smem_u32(sp));
      const uint64_t dv = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES) + A3_K_BYTES));
#pragma unroll
      for (int k = 0; k < 4; k++) tc::umma_f16(tmem_o, dp + 2 * k, dv + 2 * k, idesc, (j | k) ? 1u : 0u);
      tc::umma_commit(o_done);
    }
    __syncwarp();
  }
  tc::mbar_wait(o_done, (nkv - 1) & 1);
  tc::tc_fence_after();
  {
    uint32_t v[64];
    tc::tmem_ld_32x32(tmem_o + lane_off, v);
    tc::tmem_ld_32x32(tmem_o + lane_off + 32, v + 32);
    tc::tmem_ld_wait();
    if (row_in_batch < T) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* o = a.out + (size_t)(b * T + row_in_batch) * a.ld_out + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
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

// v4 = v3 with two threads per query row (256 threads: warps w and w+4 share TMEM lanes 32*(w%4).., each takes 32 of the 64
// score columns and 32 of the 64 output columns): twice the warps per scheduler to hide the TMEM / MUFU / shared-memory latencies.
__global__ void __launch_bounds__(256)
dit_attention_v4_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, int k_col0, AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + A3_OFF_BAR);
  uint64_t* kv_full = q_full + 1;            // [3]
  uint64_t* s_full = kv_full + A3_STAGES;    // [2]
  uint64_t* o_done = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 1);
  float* s_xch = reinterpret_cast<float*>(smem + A3_OFF_BAR + 128);      // [2 (parity)][2 (half)][128 rows] pair exchange

  const int tid = threadIdx.x, warp = tid >> 5, row = tid & 127, half = tid >> 7;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = a.T;
  const int row_in_batch = q0 + row;
  const int klim_row = a.chunk > 0 ? min(T, (row_in_batch / a.chunk + 1) * a.chunk) : T;
  const int klim_tile = a.chunk > 0 ? min(T, ((min(q0 + 127, T - 1)) / a.chunk + 1) * a.chunk) : T;
  const int nkv = (klim_tile + 63) / 64;

  if (tid == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < A3_STAGES; i++) tc::mbar_init(&kv_full[i], 1);
    tc::mbar_init(&s_full[0], 1); tc::mbar_init(&s_full[1], 1); tc::mbar_init(o_done, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;
  const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
  const uint32_t pair_bar = 1 + (warp & 3);          // named barrier shared by warps w and w+4 (64 threads)
  const uint32_t idesc = a.f16 ? tc::umma_idesc_f16(128, 64) : tc::umma_idesc_bf16(128, 64);

  auto load_kv = [&](int j) {
    const int st = j % A3_STAGES;
    uint8_t* sk = smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES);
    tc::mbar_expect_tx(&kv_full[st], A3_K_BYTES + A3_V_BYTES);
    tc::tma_load_2d(sk, &tm_k, &kv_full[st], k_col0 + h * 64, b * T + j * 64);
    tc::tma_load_2d(sk + A3_K_BYTES, &tm_v, &kv_full[st], j * 64, (b * a.heads + h) * 64);
  };
  auto issue_qk = [&](int j) {                 // S[j & 1] = Q K_j^T
    const int st = j % A3_STAGES;
    tc::mbar_wait(&kv_full[st], (j / A3_STAGES) & 1);
    tc::tc_fence_after();
    const uint64_t dq = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_Q));
    const uint64_t dk = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES)));
#pragma unroll
    for (int k = 0; k < 4; k++) tc::umma_f16(tmem_base + (uint32_t)((j & 1) * 64), dq + 2 * k, dk + 2 * k, idesc, k ? 1u : 0u);
    tc::umma_commit(&s_full[j & 1]);
  };
  if (tid == 0) {
    tc::mbar_expect_tx(q_full, A3_Q_BYTES);
    tc::tma_load_2d(smem + A3_OFF_Q, &tm_q, q_full, h * 64, b * T + q0);
    load_kv(0);
    if (nkv > 1) load_kv(1);
    tc::mbar_wait(q_full, 0);
    issue_qk(0);
  }
  const float sc = 0.125f * 1.4426950408889634f;     // dim_head^-0.5 * log2(e)
  float m_run = -INFINITY, l_run = 0.f;
  uint8_t* sp = smem + A3_OFF_P;

  for (int j = 0; j < nkv; j++) {
    if (tid == 0 && j + 1 < nkv) issue_qk(j + 1);      // S[(j+1)&1] was drained before the barrier of iteration j-1
    __syncwarp();
    tc::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
    tc::tc_fence_after();
    uint32_t sreg[32];
    tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)((j & 1) * 64 + half * 32), sreg);
    tc::tmem_ld_wait();
    const int kbase = j * 64 + half * 32;
    float mt[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (kbase + 32 <= klim_row) {
#pragma unroll
      for (int i = 0; i < 32; i++) { const float s = __uint_as_float(sreg[i]) * sc; sreg[i] = __float_as_uint(s); mt[i & 3] = fmaxf(mt[i & 3], s); }
    } else {
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const float s = (kbase + i < klim_row) ? __uint_as_float(sreg[i]) * sc : -INFINITY;
        sreg[i] = __float_as_uint(s);
        mt[i & 3] = fmaxf(mt[i & 3], s);
      }
    }
    float m_tile = fmaxf(fmaxf(mt[0], mt[1]), fmaxf(mt[2], mt[3]));
    // row maximum over both halves: exchange with the partner thread (same row, other 32 columns)
    s_xch[((j & 1) * 2 + half) * 128 + row] = m_tile;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    m_tile = fmaxf(m_tile, s_xch[((j & 1) * 2 + (half ^ 1)) * 128 + row]);
    // P.V of tile j-1 must have retired before P is overwritten / O is rescaled / its K,V stage is refilled
    if (j > 0) { tc::mbar_wait(o_done, (j - 1) & 1); tc::tc_fence_after(); }
    if (tid == 0 && j + 2 < nkv) load_kv(j + 2);       // stage (j+2)%3 == (j-1)%3 is free now
    const bool need = m_tile > m_run + 8.0f;
    if (__any_sync(0xffffffffu, need)) {
      const float m_new = need ? m_tile : m_run;
      const float alpha = need ? tc::ex2(m_run - m_new) : 1.0f;
      m_run = m_new;
      l_run *= alpha;
      if (j > 0) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_o + lane_off + (uint32_t)(half * 32), v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tc::tmem_st_32x32(tmem_o + lane_off + (uint32_t)(half * 32), v);
        tc::tmem_st_wait();
      }
    }
    float ps[4] = {0.f, 0.f, 0.f, 0.f};
    uint8_t* rowp = sp + row * 128;
    {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = tc::ex2(__uint_as_float(sreg[i]) - m_run), p1 = tc::ex2(__uint_as_float(sreg[i + 1]) - m_run);
        ps[(i >> 1) & 3] += p0 + p1;
        pk[i >> 1] = tc::pack16(p0, p1, a.f16);
      }
      const int cb = half * 4;
#pragma unroll
      for (int qd = 0; qd < 4; qd++) {
        uint4 val = make_uint4(pk[4 * qd], pk[4 * qd + 1], pk[4 * qd + 2], pk[4 * qd + 3]);
        *reinterpret_cast<uint4*>(rowp + (((cb + qd) ^ (row & 7)) << 4)) = val;
      }
    }
    l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      const int st = j % A3_STAGES;
      const uint64_t dp = tc::umma_desc_k128(tc::smem_u32(sp));
      const uint64_t dv = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES) + A3_K_BYTES));
#pragma unroll
      for (int k = 0; k < 4; k++) tc::umma_f16(tmem_o, dp + 2 * k, dv + 2 * k, idesc, (j | k) ? 1u : 0u);
      tc::umma_commit(o_done);
    }
    __syncwarp();
  }
  tc::mbar_wait(o_done, (nkv - 1) & 1);
  tc::tc_fence_after();
  {
    // total row sum = both halves
    s_xch[half * 128 + row] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
    const float l_tot = l_run + s_xch[(half ^ 1) * 128 + row];
    uint32_t v[32];
    tc::tmem_ld_32x32(tmem_o + lane_off + (uint32_t)(half * 32), v);
    tc::tmem_ld_wait();
    if (row_in_batch < T) {
      const float inv = 1.0f / l_tot;
      __nv_bfloat16* o = a.out + (size_t)(b * T + row_in_batch) * a.ld_out + h * 64 + half * 32;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
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

// v5: warp-specialised v3.  Warp 4 is the only one that talks to the TMA unit and the tensor core (loads K / V^T tiles,
// issues Q.K^T two tiles ahead and P.V as soon as the 128 softmax threads have published P through an mbarrier); the four
// softmax warps never issue an MMA and never meet at a CTA-wide barrier, so each proceeds as soon as its own score tile
// and P buffer are ready.  P is double-buffered so writing P_{j+1} does not wait for P.V_j.
constexpr int A5_P_BYTES = 2 * 128 * 128;
constexpr int A5_OFF_P = A3_OFF_KV + A3_STAGES * (A3_K_BYTES + A3_V_BYTES);
constexpr int A5_OFF_BAR = A5_OFF_P + A5_P_BYTES;
constexpr int A5_SMEM = A5_OFF_BAR + 256 + 1024;

__global__ void __launch_bounds__(160)
dit_attention_v5_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, int k_col0, AttnArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* q_full = reinterpret_cast<uint64_t*>(smem + A5_OFF_BAR);
  uint64_t* kv_full = q_full + 1;            // [3]  TMA -> MMA warp
  uint64_t* s_full = kv_full + A3_STAGES;    // [2]  tensor core -> softmax warps
  uint64_t* p_ready = s_full + 2;            // [2]  softmax warps (128 arrivals) -> MMA warp
  uint64_t* o_done = p_ready + 2;            // [2]  P.V_j retired (parity buffers by j&1): frees P[j&1], K/V stage, O for rescale
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_done + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y, b = blockIdx.z;
  const int T = a.T;
  const int Tk = a.klen ? a.klen[b] : T;     // keys of this batch row that exist (ragged group: the rest of the slab is padding)
  const int klim_tile = a.chunk > 0 ? min(Tk, ((min(q0 + 127, T - 1)) / a.chunk + 1) * a.chunk) : Tk;
  const int nkv = (klim_tile + 63) / 64;

  if (tid == 0) {
    tc::tma_prefetch_desc(&tm_q); tc::tma_prefetch_desc(&tm_k); tc::tma_prefetch_desc(&tm_v);
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < A3_STAGES; i++) tc::mbar_init(&kv_full[i], 1);
    for (int i = 0; i < 2; i++) { tc::mbar_init(&s_full[i], 1); tc::mbar_init(&p_ready[i], 128); tc::mbar_init(&o_done[i], 1); }
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;
  const uint32_t idesc = a.f16 ? tc::umma_idesc_f16(128, 64) : tc::umma_idesc_bf16(128, 64);

  if (warp == 4) {
    // ---------------- TMA + tensor-core warp (one elected lane)
    if (lane == 0) {
      auto load_kv = [&](int j) {
        const int st = j % A3_STAGES;
        uint8_t* sk = smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES);
        tc::mbar_expect_tx(&kv_full[st], A3_K_BYTES + A3_V_BYTES);
        tc::tma_load_2d(sk, &tm_k, &kv_full[st], k_col0 + h * 64, b * T + j * 64);
        tc::tma_load_2d(sk + A3_K_BYTES, &tm_v, &kv_full[st], j * 64, (b * a.heads + h) * 64);
      };
      auto issue_qk = [&](int j) {
        const int st = j % A3_STAGES;
        tc::mbar_wait(&kv_full[st], (j / A3_STAGES) & 1);
        tc::tc_fence_after();
        const uint64_t dq = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_Q));
        const uint64_t dk = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES)));
#pragma unroll
        for (int k = 0; k < 4; k++) tc::umma_f16(tmem_base + (uint32_t)((j & 1) * 64), dq + 2 * k, dk + 2 * k, idesc, k ? 1u : 0u);
        tc::umma_commit(&s_full[j & 1]);
      };
      tc::mbar_expect_tx(q_full, A3_Q_BYTES);
      tc::tma_load_2d(smem + A3_OFF_Q, &tm_q, q_full, h * 64, b * T + q0);
      for (int j = 0; j < min(nkv, A3_STAGES); j++) load_kv(j);
      tc::mbar_wait(q_full, 0);
      issue_qk(0);
      if (nkv > 1) issue_qk(1);
      for (int j = 0; j < nkv; j++) {
        // P_j published (which also means S_j has been read: S[j&1] may be overwritten by Q.K_{j+2})
        tc::mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
        tc::tc_fence_after();
        const int st = j % A3_STAGES;
        const uint64_t dp = tc::umma_desc_k128(tc::smem_u32(smem + A5_OFF_P + (j & 1) * (128 * 128)));
        const uint64_t dv = tc::umma_desc_k128(tc::smem_u32(smem + A3_OFF_KV + st * (A3_K_BYTES + A3_V_BYTES) + A3_K_BYTES));
#pragma unroll
        for (int k = 0; k < 4; k++) tc::umma_f16(tmem_o, dp + 2 * k, dv + 2 * k, idesc, (j | k) ? 1u : 0u);
        tc::umma_commit(&o_done[j & 1]);
        if (j + 2 < nkv) issue_qk(j + 2);
        if (j + A3_STAGES < nkv) {              // refill this tile's K/V stage once P.V_j has retired
          tc::mbar_wait(&o_done[j & 1], (j >> 1) & 1);
          load_kv(j + A3_STAGES);
        }
      }
    }
  } else {
    // ---------------- softmax warps: thread == query row == TMEM lane
    const int row_in_batch = q0 + tid;
    const int klim_row = a.chunk > 0 ? min(Tk, (row_in_batch / a.chunk + 1) * a.chunk) : Tk;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const float sc = 0.125f * 1.4426950408889634f;     // dim_head^-0.5 * log2(e)
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nkv; j++) {
      tc::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc::tc_fence_after();
      uint32_t sreg[64];
      tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)((j & 1) * 64), sreg);
      tc::tmem_ld_32x32(tmem_base + lane_off + (uint32_t)((j & 1) * 64 + 32), sreg + 32);
      tc::tmem_ld_wait();
      const int kbase = j * 64;
      float mt[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      // raw scores stay in registers: the max is taken on them (sc > 0) and the scale rides in the ex2 argument as one FFMA
      if (kbase + 64 <= klim_row) {
#pragma unroll
        for (int i = 0; i < 64; i++) mt[i & 3] = fmaxf(mt[i & 3], __uint_as_float(sreg[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 64; i++) {
          const float s = (kbase + i < klim_row) ? __uint_as_float(sreg[i]) : -INFINITY;
          sreg[i] = __float_as_uint(s);
          mt[i & 3] = fmaxf(mt[i & 3], s);
        }
      }
      const float m_tile = fmaxf(fmaxf(mt[0], mt[1]), fmaxf(mt[2], mt[3])) * sc;
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
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            tc::tmem_ld_32x32(tmem_o + lane_off + (uint32_t)c0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tc::tmem_st_32x32(tmem_o + lane_off + (uint32_t)c0, v);
          }
          tc::tmem_st_wait();
        }
      }
      // P buffer j&1 was last read by P.V_{j-2}
      if (j >= 2) tc::mbar_wait(&o_done[j & 1], ((j - 2) >> 1) & 1);
      float ps[4] = {0.f, 0.f, 0.f, 0.f};
      uint8_t* rowp = smem + A5_OFF_P + (j & 1) * (128 * 128) + tid * 128;
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float p0 = tc::ex2(fmaf(__uint_as_float(sreg[c0 + i]), sc, -m_run)), p1 = tc::ex2(fmaf(__uint_as_float(sreg[c0 + i + 1]), sc, -m_run));
          ps[(i >> 1) & 3] += p0 + p1;
          pk[i >> 1] = tc::pack16(p0, p1, a.f16);
        }
        const int cb = c0 >> 3;
#pragma unroll
        for (int qd = 0; qd < 4; qd++) {
          uint4 val = make_uint4(pk[4 * qd], pk[4 * qd + 1], pk[4 * qd + 2], pk[4 * qd + 3]);
          *reinterpret_cast<uint4*>(rowp + (((cb + qd) ^ (tid & 7)) << 4)) = val;
        }
      }
      l_run += (ps[0] + ps[1]) + (ps[2] + ps[3]);
      tc::fence_proxy_async();                 // generic-proxy P stores -> visible to the tensor core
      tc::tc_fence_before();                   // orders this thread's TMEM reads / rescale before the MMA warp's next issue
      tc::mbar_arrive(&p_ready[j & 1]);
    }
    tc::mbar_wait(&o_done[(nkv - 1) & 1], ((nkv - 1) >> 1) & 1);
    tc::tc_fence_after();
    uint32_t v[64];
    tc::tmem_ld_32x32(tmem_o + lane_off, v);
    tc::tmem_ld_32x32(tmem_o + lane_off + 32, v + 32);
    tc::tmem_ld_wait();
    if (row_in_batch < T && a.lo_off) {
      // split precision (flow parity mode): hi at [col], lo = y - hi at [lo_off + col]; 16-byte stores like the plain path
      // (element-wise 2-byte stores made this epilogue as long as the whole key loop: 249 vs 479 TFLOP/s)
      const float inv = 1.0f / l_run;
      uint16_t* o = reinterpret_cast<uint16_t*>(a.out) + (size_t)(b * T + row_in_batch) * (2 * a.ld_out) + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; j++) y[j] = __uint_as_float(v[i + j]) * inv;
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          hi[j] = tc::pack16(y[2 * j], y[2 * j + 1], a.f16);
          float r0, r1;
          if (a.f16) { const __half2 t = *reinterpret_cast<const __half2*>(&hi[j]); r0 = y[2 * j] - __low2float(t); r1 = y[2 * j + 1] - __high2float(t); }
          else { const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&hi[j]); r0 = y[2 * j] - __low2float(t); r1 = y[2 * j + 1] - __high2float(t); }
          lo[j] = tc::pack16(r0, r1, a.f16);
        }
        *reinterpret_cast<uint4*>(o + i) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(o + a.lo_off + i) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    } else if (row_in_batch < T) {
      const float inv = 1.0f / l_run;
      __nv_bfloat16* o = a.out + (size_t)(b * T + row_in_batch) * a.ld_out + h * 64;
#pragma unroll
      for (int i = 0; i < 64; i += 8) {
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
  CUtensorMap tq, tk, tv;
  const uint64_t rows = (uint64_t)a.n_batch * a.T;
  HVX_CHECK(make_tmap_bf16_2d(&tq, qk, rows, ld_qk, ld_qk, 128, 64), HVX_ERR_CUDA, "attention: tensor map Q failed");
  HVX_CHECK(make_tmap_bf16_2d(&tk, qk, rows, ld_qk, ld_qk, 128, 64), HVX_ERR_CUDA, "attention: tensor map K failed");
  HVX_CHECK(make_tmap_bf16_2d(&tv, vt, (uint64_t)a.n_batch * a.heads * 64, vt_ld, vt_ld, 64, 64), HVX_ERR_CUDA,
            "attention: tensor map V^T failed");
  static bool attr_set = false;
  if (!attr_set) {
    HVX_CUDA(cudaFuncSetAttribute(dit_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    HVX_CUDA(cudaFuncSetAttribute(dit_attention_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    attr_set = true;
  }
  dim3 grid(cdiv(a.T, 128), a.heads, a.n_batch);
  ProfScope prof_scope(&e->prof, st, PROF_ATTN, a.work > 0 ? a.work : 4.0 * a.n_batch * a.heads * (double)a.T * a.T * 64.0 * (a.chunk > 0 ? 0.5 : 1.0));
  HVX_CHECK(!a.klen || !(getenv("HVX_ATTN_V1") || getenv("HVX_ATTN_V2") || getenv("HVX_ATTN_V3") || getenv("HVX_ATTN_V4")), HVX_ERR_UNSUPPORTED,
            "attention: per-batch key counts need the v5 kernel");
  HVX_CHECK(!a.lo_off || !(getenv("HVX_ATTN_V1") || getenv("HVX_ATTN_V2")), HVX_ERR_UNSUPPORTED, "attention: split output needs the v5 kernel");
  if (getenv("HVX_ATTN_V1")) dit_attention_kernel<<<grid, 128, AT_SMEM, st>>>(tq, tk, tv, k_col0, a);
  else if (getenv("HVX_ATTN_V2")) dit_attention_v2_kernel<<<grid, 128, AT_SMEM, st>>>(tq, tk, tv, k_col0, a);
  else {
    static bool a3 = false;
    if (!a3) { HVX_CUDA(cudaFuncSetAttribute(dit_attention_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM)); a3 = true; }
    CUtensorMap tk64;
    HVX_CHECK(make_tmap_bf16_2d(&tk64, qk, rows, ld_qk, ld_qk, 64, 64), HVX_ERR_CUDA, "attention: tensor map K(64) failed");
    static bool a4 = false;
    if (!a4) { HVX_CUDA(cudaFuncSetAttribute(dit_attention_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM)); a4 = true; }
    static bool a5 = false;
    if (!a5) { HVX_CUDA(cudaFuncSetAttribute(dit_attention_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A5_SMEM)); a5 = true; }
    HVX_CHECK(!a.lo_off || !(getenv("HVX_ATTN_V3") || getenv("HVX_ATTN_V4")), HVX_ERR_UNSUPPORTED, "attention: split output needs the v5 kernel");
    if (getenv("HVX_ATTN_V3")) dit_attention_v3_kernel<<<grid, 128, A3_SMEM, st>>>(tq, tk64, tv, k_col0, a);
    else if (getenv("HVX_ATTN_V4")) dit_attention_v4_kernel<<<grid, 256, A3_SMEM, st>>>(tq, tk64, tv, k_col0, a);
    else dit_attention_v5_kernel<<<grid, 160, A5_SMEM, st>>>(tq, tk64, tv, k_col0, a);
  }
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
