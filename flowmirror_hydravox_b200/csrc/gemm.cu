// tcgen05 + TMA GEMM (see gemm.cuh).  Used by the DiT estimator's linears (DiT/modules.py:349-407,
// 500-530) and by anything else on the path that is a dense projection over >= 16 rows.
#include "gemm.cuh"
#include <cuda.h>
#include <mutex>
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace hvx {

// ------------------------------------------------------------------ host: tensor maps
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

bool make_tmap_bf16_2d(CUtensorMap* out, const void* gptr, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                       uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool make_tmap_bf16_3d(CUtensorMap* out, const void* gptr, uint64_t batches, uint64_t rows, uint64_t cols,
                       uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) return false;
  cuuint64_t dims[3] = {cols, rows, batches};
  cuuint64_t strides[2] = {row_stride_elems * 2, row_stride_elems * 2 * rows};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// ------------------------------------------------------------------ device
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;

// TRI = the three-term split-precision product (GemmAddr::split3_kb): one pipeline stage holds the four operand tiles of a k-block
// [A_hi | A_lo | B_hi | B_lo] and feeds three MMA groups A_hi B_hi + A_lo B_hi + A_hi B_lo from them — every operand tile crosses
// L2 -> shared memory once instead of 1.5 times (the GEMMs are L2 -> SM bandwidth bound, profiles/README.md).
template <int BN, int STAGES, bool TRI = false>
struct GemmSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = (TRI ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;   // barriers + tmem slot + alignment slack
};

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == ACT_GELU_TANH) {
    // 0.5 x (1 + tanh(u)) == x / (1 + exp(-2u)),  u = sqrt(2/pi) (x + 0.044715 x^3): one ex2 + one rcp instead of tanhf
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    const float u = k0 * (v + k1 * v * v * v);
    return __fdividef(v, 1.0f + __expf(-2.0f * u));
  }
  if (act == ACT_SILU) return v / (1.0f + expf(-v));
  if (act == ACT_MISH) { const float sp = v > 20.0f ? v : log1pf(expf(v)); return v * tanhf(sp); }
  if (act == ACT_LRELU) return v > 0.f ? v : 0.01f * v;
  if (act == ACT_GELU_ERF) {
    // exact-erf GELU (F.gelu default).  erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7, below fp32 noise of the GEMM):
    // one rcp + one ex2 + 7 FMAs instead of erff's ~35-instruction branchy polynomial — the epilogue of the 128x256 ff1
    // tile was longer than its K = 256 main loop (31 us per launch under ncu)
    const float z = fabsf(v) * 0.70710678118654752f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f); p = fmaf(p, t, -0.284496736f); p = fmaf(p, t, 0.254829592f);
    const float er = 1.0f - p * t * __expf(-z * z);          // erf(|v|/sqrt2)
    return 0.5f * v * (1.0f + copysignf(er, v));
  }
  return v;
}

// fused epilogue of one 32-column chunk of one output row: v = raw fp32 accumulators of columns [col0, col0+32)
// MODE / ACT are compile-time: one kernel instance carries exactly one epilogue.  (With every mode inlined the kernel was
// 21.6k SASS instructions and each tile's epilogue ran out of a cold instruction cache: ~25 us of fetch stalls per tile.)
// VEC_BIAS: only the persistent kernels (large problems) carry the 16-byte bias loads.  The extra code path costs the one-tile-per-CTA
// kernel of the small launches more in instruction fetch than the loads save: with it in every instance the first audio chunk of
// the streaming path went from 75 to 89 ms (scripts/first_audio.py, A/B against the previous build on one box).
template <int MODE, int ACT, bool VEC_BIAS = false>
__device__ __forceinline__ void epi_store(const GemmEpi& epi, int row, int bidx, int col0, int N, const uint32_t* v,
                                          const float* pre = nullptr) {
      float f[32];
      if (VEC_BIAS && epi.bias && col0 + 32 <= N && ((uintptr_t)epi.bias & 15) == 0) {
        // the chunk's 32 bias values as eight 16-byte loads (the same addresses for every thread of the warp: one transaction
        // each); 32 scalar loads with their dependent adds were the top stall of the epilogue warps (profiles/r2 ff1 source view)
        float4 b4[8];
#pragma unroll
        for (int j = 0; j < 8; j++) b4[j] = __ldg(reinterpret_cast<const float4*>(epi.bias + col0) + j);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          f[4 * j] = act_apply(__uint_as_float(v[4 * j]) + b4[j].x, ACT);
          f[4 * j + 1] = act_apply(__uint_as_float(v[4 * j + 1]) + b4[j].y, ACT);
          f[4 * j + 2] = act_apply(__uint_as_float(v[4 * j + 2]) + b4[j].z, ACT);
          f[4 * j + 3] = act_apply(__uint_as_float(v[4 * j + 3]) + b4[j].w, ACT);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) {
          float x = __uint_as_float(v[j]);
          if (epi.bias && col0 + j < N) x += __ldg(epi.bias + col0 + j);
          f[j] = act_apply(x, ACT);
        }
      }
      if (MODE == EPI_BF16) {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(epi.out) + (size_t)row * epi.ldo + col0;
        if (col0 + 32 <= N) {
          // a thread owns 64 contiguous bytes of its row in the hi half and 64 in the lo half: 16-byte stores for both (2-byte
          // stores of the lo half kept the L1 60 % busy and the ff1 GEMM at 21 % tensor pipe, profiles/r2k)
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            pk.x = tc::pack16(f[j], f[j + 1], epi.f16); pk.y = tc::pack16(f[j + 2], f[j + 3], epi.f16);
            pk.z = tc::pack16(f[j + 4], f[j + 5], epi.f16); pk.w = tc::pack16(f[j + 6], f[j + 7], epi.f16);
            *reinterpret_cast<uint4*>(o + j) = pk;
            if (epi.lo_off) {
              uint4 pl;
              pl.x = tc::pack_lo16(f[j], f[j + 1], pk.x, epi.f16); pl.y = tc::pack_lo16(f[j + 2], f[j + 3], pk.y, epi.f16);
              pl.z = tc::pack_lo16(f[j + 4], f[j + 5], pk.z, epi.f16); pl.w = tc::pack_lo16(f[j + 6], f[j + 7], pk.w, epi.f16);
              *reinterpret_cast<uint4*>(o + epi.lo_off + j) = pl;
            }
          }
        } else {
          #pragma unroll
          for (int j = 0; j < 32; j++) if (col0 + j < N) {
            reinterpret_cast<uint16_t*>(o)[j] = tc::cvt16(f[j], epi.f16);
            if (epi.lo_off) reinterpret_cast<uint16_t*>(o)[epi.lo_off + j] = tc::lo16(f[j], epi.f16);
          }
        }
      } else if (MODE == EPI_F32) {
        float* o = reinterpret_cast<float*>(epi.out) + (size_t)row * epi.ldo + col0;
        const bool vec = col0 + 32 <= N && (epi.ldo & 3) == 0 && ((uintptr_t)epi.out & 15) == 0 &&
                         (!epi.resid || ((uintptr_t)epi.resid & 15) == 0);          // a thread owns 128 contiguous bytes of its row
        if (vec) {
          if (epi.resid) {
            const float4* r = reinterpret_cast<const float4*>(epi.resid + (size_t)row * epi.ldo + col0);
#pragma unroll
            for (int j = 0; j < 8; j++) { const float4 rr = r[j]; f[4 * j] += rr.x; f[4 * j + 1] += rr.y; f[4 * j + 2] += rr.z; f[4 * j + 3] += rr.w; }
          }
#pragma unroll
          for (int j = 0; j < 8; j++) reinterpret_cast<float4*>(o)[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
          if (epi.resid) {
            const float* r = epi.resid + (size_t)row * epi.ldo + col0;
            #pragma unroll
            for (int j = 0; j < 32; j++) if (col0 + j < N) f[j] += r[j];
          }
          #pragma unroll
          for (int j = 0; j < 32; j++) if (col0 + j < N) o[j] = f[j];
        }
        if (epi.out2) {
          const int ld2 = epi.ldo2 ? epi.ldo2 : epi.ldo;
          __nv_bfloat16* o2 = epi.out2 + (size_t)row * ld2 + col0;
          if (col0 + 32 <= N && (ld2 & 7) == 0 && (epi.lo_off & 7) == 0 && ((uintptr_t)epi.out2 & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk;
              pk.x = tc::pack16(f[j], f[j + 1], epi.f16); pk.y = tc::pack16(f[j + 2], f[j + 3], epi.f16);
              pk.z = tc::pack16(f[j + 4], f[j + 5], epi.f16); pk.w = tc::pack16(f[j + 6], f[j + 7], epi.f16);
              *reinterpret_cast<uint4*>(o2 + j) = pk;
              if (epi.lo_off) {
                uint4 pl;
                pl.x = tc::pack_lo16(f[j], f[j + 1], pk.x, epi.f16); pl.y = tc::pack_lo16(f[j + 2], f[j + 3], pk.y, epi.f16);
                pl.z = tc::pack_lo16(f[j + 4], f[j + 5], pk.z, epi.f16); pl.w = tc::pack_lo16(f[j + 6], f[j + 7], pk.w, epi.f16);
                *reinterpret_cast<uint4*>(o2 + epi.lo_off + j) = pl;
              }
            }
          } else {
            #pragma unroll
            for (int j = 0; j < 32; j++) if (col0 + j < N) {
              reinterpret_cast<uint16_t*>(o2)[j] = tc::cvt16(f[j], epi.f16);
              if (epi.lo_off) reinterpret_cast<uint16_t*>(o2)[epi.lo_off + j] = tc::lo16(f[j], epi.f16);
            }
          }
        }
      } else if (MODE == EPI_RESID_GATE) {
        float* o = reinterpret_cast<float*>(epi.out) + (size_t)row * epi.ldo + col0;
        const float* g = epi.gate + (size_t)bidx * epi.gate_ld + col0;
        if (col0 + 32 <= N) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 cur = pre ? make_float4(pre[j], pre[j + 1], pre[j + 2], pre[j + 3]) : *reinterpret_cast<float4*>(o + j);
            const float4 gg = __ldg(reinterpret_cast<const float4*>(g + j));
            cur.x = fmaf(gg.x, f[j], cur.x); cur.y = fmaf(gg.y, f[j + 1], cur.y);
            cur.z = fmaf(gg.z, f[j + 2], cur.z); cur.w = fmaf(gg.w, f[j + 3], cur.w);
            *reinterpret_cast<float4*>(o + j) = cur;
          }
        } else {
          #pragma unroll
          for (int j = 0; j < 32; j++) if (col0 + j < N) o[j] = fmaf(__ldg(g + j), f[j], o[j]);
        }
      } else if (MODE == EPI_HIFT) {
        const HiftEpi& h = epi.hift;
        const bool full = col0 + 32 <= N;
        if (h.resid) {
          const float* r = h.resid + (size_t)row * h.ldr + col0;
          if (full && (h.ldr & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) { const float4 rr = reinterpret_cast<const float4*>(r)[j]; f[4 * j] += rr.x; f[4 * j + 1] += rr.y; f[4 * j + 2] += rr.z; f[4 * j + 3] += rr.w; }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) if (col0 + j < N) f[j] += r[j];
          }
        }
        if (h.resid2) {
          const float* r = h.resid2 + (size_t)row * h.ldr + col0;
          if (full && (h.ldr & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) { const float4 rr = reinterpret_cast<const float4*>(r)[j]; f[4 * j] += rr.x; f[4 * j + 1] += rr.y; f[4 * j + 2] += rr.z; f[4 * j + 3] += rr.w; }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) if (col0 + j < N) f[j] += r[j];
          }
        }
        if (h.out32) {
          float* o = h.out32 + (size_t)(row + h.row_shift) * h.ld32 + col0;
          if (full && (h.ld32 & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
              float4 o4 = make_float4(f[4 * j] * h.out_scale, f[4 * j + 1] * h.out_scale, f[4 * j + 2] * h.out_scale, f[4 * j + 3] * h.out_scale);
              if (h.accumulate) { const float4 c4 = reinterpret_cast<const float4*>(o)[j]; o4.x += c4.x; o4.y += c4.y; o4.z += c4.z; o4.w += c4.w; }
              reinterpret_cast<float4*>(o)[j] = o4;
              if (h.dup_row1 && row == 1) reinterpret_cast<float4*>(o - (size_t)(1 + h.row_shift) * h.ld32 + (size_t)0)[j] = o4;
              f[4 * j] = o4.x; f[4 * j + 1] = o4.y; f[4 * j + 2] = o4.z; f[4 * j + 3] = o4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) if (col0 + j < N) {
              float ov = f[j] * h.out_scale;
              if (h.accumulate) ov += o[j];
              o[j] = ov;
              if (h.dup_row1 && row == 1) o[j - (ptrdiff_t)(1 + h.row_shift) * h.ld32] = ov;
              f[j] = ov;
            }
          }
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          if (k >= h.n16) break;
          float a[32];
          if (h.act[k] == 2) {
            const float* al = h.alpha[k] + col0;
            const float* ia = h.inv_alpha[k] + col0;
#pragma unroll
            for (int j = 0; j < 32; j++) {
              const int jc = (col0 + j < N) ? j : 0;
              const float sn = sinf(f[j] * __ldg(al + jc));
              a[j] = f[j] + __ldg(ia + jc) * (sn * sn);
            }
          } else if (h.act[k] == 1) {
#pragma unroll
            for (int j = 0; j < 32; j++) a[j] = f[j] > 0.f ? f[j] : f[j] * h.slope;
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) a[j] = f[j];
          }
          uint16_t* o = h.out16[k] + (size_t)row * h.ld16 + col0;
          if (full) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 pk, pl;
              const __half2 h0 = __floats2half2_rn(a[j], a[j + 1]), h1 = __floats2half2_rn(a[j + 2], a[j + 3]);
              const __half2 h2 = __floats2half2_rn(a[j + 4], a[j + 5]), h3 = __floats2half2_rn(a[j + 6], a[j + 7]);
              pk.x = *reinterpret_cast<const uint32_t*>(&h0); pk.y = *reinterpret_cast<const uint32_t*>(&h1);
              pk.z = *reinterpret_cast<const uint32_t*>(&h2); pk.w = *reinterpret_cast<const uint32_t*>(&h3);
              const __half2 l0 = __floats2half2_rn(a[j] - __low2float(h0), a[j + 1] - __high2float(h0));
              const __half2 l1 = __floats2half2_rn(a[j + 2] - __low2float(h1), a[j + 3] - __high2float(h1));
              const __half2 l2 = __floats2half2_rn(a[j + 4] - __low2float(h2), a[j + 5] - __high2float(h2));
              const __half2 l3 = __floats2half2_rn(a[j + 6] - __low2float(h3), a[j + 7] - __high2float(h3));
              pl.x = *reinterpret_cast<const uint32_t*>(&l0); pl.y = *reinterpret_cast<const uint32_t*>(&l1);
              pl.z = *reinterpret_cast<const uint32_t*>(&l2); pl.w = *reinterpret_cast<const uint32_t*>(&l3);
              *reinterpret_cast<uint4*>(o + j) = pk;
              *reinterpret_cast<uint4*>(o + h.lo_off + j) = pl;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; j++) if (col0 + j < N) { o[j] = tc::cvt16(a[j], 1); o[h.lo_off + j] = tc::lo16(a[j], 1); }
          }
        }
      } else if (MODE == EPI_LLM_QKV) {
        if (col0 + 32 <= N) llm_qkv_store32(epi.llm, row, col0, f);
        else {
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            if (col0 + j < N) llm_qkv_store(epi.llm, row, col0 + j, f[j], f[j + 1]);
        }
      } else if (MODE == EPI_SWIGLU) {
        uint16_t* o = reinterpret_cast<uint16_t*>(epi.out) + (size_t)row * epi.ldo + (col0 >> 1);
        if (col0 + 32 <= N && (epi.ldo & 7) == 0 && (epi.lo_off & 7) == 0) {
          // 16 outputs of this row: two 16-byte stores for the hi halves, two for the lo halves (2-byte stores, one row per
          // thread, cost more than the k-loop of the 128-row decode GEMM)
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float y0 = (f[4 * j] / (1.0f + expf(-f[4 * j]))) * f[4 * j + 1];
            const float y1 = (f[4 * j + 2] / (1.0f + expf(-f[4 * j + 2]))) * f[4 * j + 3];
            hi[j] = tc::pack16(y0, y1, epi.f16);
            float r0, r1;
            if (epi.f16) { const __half2 t = *reinterpret_cast<const __half2*>(&hi[j]); r0 = y0 - __low2float(t); r1 = y1 - __high2float(t); }
            else { const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&hi[j]); r0 = y0 - __low2float(t); r1 = y1 - __high2float(t); }
            lo[j] = tc::pack16(r0, r1, epi.f16);
          }
          reinterpret_cast<uint4*>(o)[0] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          reinterpret_cast<uint4*>(o)[1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
          if (epi.lo_off) {
            reinterpret_cast<uint4*>(o + epi.lo_off)[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            reinterpret_cast<uint4*>(o + epi.lo_off)[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2)
            if (col0 + j < N) {
              const float y = (f[j] / (1.0f + expf(-f[j]))) * f[j + 1];
              o[j >> 1] = tc::cvt16(y, epi.f16);
              if (epi.lo_off) o[epi.lo_off + (j >> 1)] = tc::lo16(y, epi.f16);
            }
        }
      } else if (MODE == EPI_QKV) {
        const int t = row - bidx * epi.T;       // rows_per_batch == T
        const int tp = t + epi.t_off;           // absolute frame (windowed call: rotary position, cache row, V^T column)
        if (col0 < epi.n_qk) {
          const int dq = epi.n_qk >> 1;
          const int cl = col0 % dq;
          if (cl < 64 && epi.rope_cos) {            // null table: no rotary (U-Net estimator, unet.cu)
            const float* cs = epi.rope_cos + (size_t)tp * 32 + (cl >> 1);
            const float* sn = epi.rope_sin + (size_t)tp * 32 + (cl >> 1);
#pragma unroll
            for (int i = 0; i < 16; i++) {
              const float c = __ldg(cs + i), s = __ldg(sn + i);
              const float x1 = f[2 * i], x2 = f[2 * i + 1];
              f[2 * i] = x1 * c - x2 * s;
              f[2 * i + 1] = x2 * c + x1 * s;
            }
          }
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(epi.out) + (size_t)row * epi.ldo + col0;
          if (epi.k_out && col0 >= dq) o = epi.k_out + ((size_t)bidx * epi.k_batch_rows + tp) * epi.k_ld + (col0 - dq);
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            pk.x = tc::pack16(f[j], f[j + 1], epi.f16); pk.y = tc::pack16(f[j + 2], f[j + 3], epi.f16);
            pk.z = tc::pack16(f[j + 4], f[j + 5], epi.f16); pk.w = tc::pack16(f[j + 6], f[j + 7], epi.f16);
            *reinterpret_cast<uint4*>(o + j) = pk;
          }
        } else {
          const int cv = col0 - epi.n_qk;
          const int h = cv >> 6, d0 = cv & 63;
          __nv_bfloat16* o = epi.vt + ((size_t)(bidx * epi.heads + h) * 64 + d0) * epi.vt_ld + tp;
#pragma unroll
          for (int j = 0; j < 32; j++) reinterpret_cast<uint16_t*>(o)[(size_t)j * epi.vt_ld] = tc::cvt16(f[j], epi.f16);
        }
      }
}

template <int BN, int STAGES, int MODE, int ACT, bool TRI>
__global__ void __launch_bounds__(GEMM_THREADS)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M, int N, int K,
                 GemmEpi epi, GemmAddr ad, int tiles_per_batch) {
  using S = GemmSmem<BN, STAGES, TRI>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;
  const int batch = blockIdx.y / tiles_per_batch;
  const int m0 = (blockIdx.y - batch * tiles_per_batch) * BM;      // row inside the batch
  // split-K (ad.split_k > 1, grid.z): this CTA accumulates k-blocks [kb_lo, kb_lo + nkb) and stores its partial sum to
  // out + z * split_stride (plain EPI_F32, bias from split 0 only); the caller adds the partials in a fixed order
  const int nkb_all = TRI ? ad.split3_kb : (K + BK - 1) / BK;       // TRI: one pipeline step per k-block of the un-split product
  const int kb_lo = (int)(((long long)nkb_all * blockIdx.z) / gridDim.z);
  const int nkb = (int)(((long long)nkb_all * (blockIdx.z + 1)) / gridDim.z) - kb_lo;
  if (blockIdx.z) { epi.out = reinterpret_cast<float*>(epi.out) + (size_t)blockIdx.z * ad.split_stride; epi.bias = nullptr; }

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; s++) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, BN);
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  pdl_wait();                          // everything above overlaps the predecessor's tail; nothing global has been touched yet
  const uint32_t tmem_base = *tmem_slot;

  const bool dbg = ad.dbg && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  if (dbg && threadIdx.x == 0) ad.dbg[0] = tc::gtimer();
  if (warp == 0) {
    if (lane == 0) {
      for (int ki = 0; ki < nkb; ki++) {
        const int kb = kb_lo + ki;
        const int s = ki % STAGES;
        const uint32_t ph = (ki / STAGES) & 1;
        tc::mbar_wait(&empty_bar[s], ph ^ 1);
        if (dbg && ki < 64) ad.dbg[1 + ki] = tc::gtimer();
        tc::mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
        uint8_t* sa = smem + s * S::STAGE_BYTES;
        const int tap = ad.kb_per_tap ? kb / ad.kb_per_tap : 0;
        const int kin = ad.kb_per_tap ? kb - tap * ad.kb_per_tap : kb;
        const int acol = ad.a_col0 + blockIdx.x * ad.a_col_per_ntile + kin * BK, arow = m0 + ad.a_row0 + tap * ad.a_row_step;
        if (TRI) {
          tc::tma_load_3d(sa, &tma_a, &full_bar[s], acol, arow, batch);
          tc::tma_load_3d(sa + S::A_BYTES, &tma_a, &full_bar[s], acol + ad.a_lo_off, arow, batch);
          tc::tma_load_2d(sa + 2 * S::A_BYTES, &tma_b, &full_bar[s], kb * BK, n0);
          tc::tma_load_2d(sa + 2 * S::A_BYTES + S::B_BYTES, &tma_b, &full_bar[s], (ad.split3_kb + kb) * BK, n0);
        } else {
          const int kbb = ad.b_kb_mod ? kb % ad.b_kb_mod : kb;
          tc::tma_load_3d(sa, &tma_a, &full_bar[s], acol, arow, batch);
          tc::tma_load_2d(sa + S::A_BYTES, &tma_b, &full_bar[s], kbb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = epi.f16 ? tc::umma_idesc_f16(BM, BN) : tc::umma_idesc_bf16(BM, BN);
      for (int kb = 0; kb < nkb; kb++) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        tc::mbar_wait(&full_bar[s], ph);
        if (dbg && kb < 64) ad.dbg[65 + kb] = tc::gtimer();
        tc::tc_fence_after();
        const uint32_t sa = tc::smem_u32(smem + s * S::STAGE_BYTES);
        if (TRI) {
          const uint64_t a_hi = tc::umma_desc_k128(sa), a_lo = tc::umma_desc_k128(sa + S::A_BYTES);
          const uint64_t b_hi = tc::umma_desc_k128(sa + 2 * S::A_BYTES), b_lo = tc::umma_desc_k128(sa + 2 * S::A_BYTES + S::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; k++) tc::umma_f16(tmem_base, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BK / 16; k++) tc::umma_f16(tmem_base, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < BK / 16; k++) tc::umma_f16(tmem_base, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
        } else {
          const uint64_t adesc = tc::umma_desc_k128(sa);
          const uint64_t bdesc = tc::umma_desc_k128(sa + S::A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; k++)
            tc::umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
        }
        tc::umma_commit(&empty_bar[s]);          // frees the smem slot once these MMAs retire
      }
      tc::umma_commit(tmem_full);                // accumulator complete
    }
  } else {
    // ---------------- epilogue: warp q owns TMEM lanes [32q, 32q+32) = tile rows
    const int q = warp & 3;
    const int row_b = m0 + q * 32 + lane;
    const int row = batch * ad.rows_per_batch + row_b;
    tc::mbar_wait(tmem_full, 0);
    if (dbg && threadIdx.x == 64) ad.dbg[130] = tc::gtimer();
    tc::tc_fence_after();
    const bool row_ok = row_b < ad.rows_per_batch && row < M;
    const int bidx = row / epi.rows_per_batch;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      tc::tmem_ld_wait();
      const int col0 = n0 + c0;
      if (!row_ok || col0 >= N) continue;
      epi_store<MODE, ACT>(epi, row, bidx, col0, N, v);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) ad.dbg[131] = tc::gtimer();
  if (warp == 1) { __syncwarp(); tc::tmem_dealloc(tmem_base, BN); }
}

// ------------------------------------------------------------------ persistent variant for the big flow GEMMs
// One CTA per SM walks the output tiles (128 x PBN).  The accumulator is double-buffered in TMEM
// (2 x PBN fp32 columns = all 512 columns), so the epilogue of tile i (tcgen05.ld -> fused math -> global) runs
// while the tensor core already accumulates tile i+1; operands stream through a 4-stage TMA ring.
// 128 x 256 tiles read 96 B/cycle of operands from shared memory (A 4 KB + B 8 KB per 128-cycle MMA), under
// the 128 B/cycle port limit that caps 128 x 128 tiles.
constexpr int PBN = 256;
// TRI (three-term split precision): a stage is [A_hi | A_lo | B_hi | B_lo] = 96 KB feeding 12 MMAs (1536 tensor-core cycles), two stages
template <bool TRI>
struct PersistSmem {
  static constexpr int STAGES = TRI ? 2 : 4;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = PBN * BK * 2;
  static constexpr int STAGE_BYTES = (TRI ? 2 : 1) * (A_BYTES + B_BYTES);
  static constexpr int BAR_OFF = STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

constexpr int PERSIST_THREADS = 320;     // TMA warp, MMA warp, 8 epilogue warps
template <int MODE, int ACT, bool TRI>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M, int N, int K,
                    GemmEpi epi, GemmAddr ad, int tiles_m, int tiles_n) {
  using S = PersistSmem<TRI>;
  constexpr int PSTAGES = S::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty_bar = full_bar + PSTAGES;
  uint64_t* tmem_full = empty_bar + PSTAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = TRI ? ad.split3_kb : (K + BK - 1) / BK;          // TRI: one pipeline step per k-block of the un-split product
  const int n_tiles = tiles_m * tiles_n;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_b);
    for (int s = 0; s < PSTAGES; s++) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; s++) { tc::mbar_init(&tmem_full[s], 1); tc::mbar_init(&tmem_empty[s], 8); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 2 * PBN);
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;                                  // running k-block counter across tiles
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile / tiles_n, nt = tile - mt * tiles_n;
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % PSTAGES;
          const uint32_t ph = (it / PSTAGES) & 1;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          tc::mbar_expect_tx(&full_bar[s], S::STAGE_BYTES);
          uint8_t* sa = smem + s * S::STAGE_BYTES;
          if (TRI) {
            tc::tma_load_3d(sa, &tma_a, &full_bar[s], kb * BK, mt * BM, 0);
            tc::tma_load_3d(sa + S::A_BYTES, &tma_a, &full_bar[s], ad.a_lo_off + kb * BK, mt * BM, 0);
            tc::tma_load_2d(sa + 2 * S::A_BYTES, &tma_b, &full_bar[s], kb * BK, nt * PBN);
            tc::tma_load_2d(sa + 2 * S::A_BYTES + S::B_BYTES, &tma_b, &full_bar[s], (ad.split3_kb + kb) * BK, nt * PBN);
          } else {
            tc::tma_load_3d(sa, &tma_a, &full_bar[s], kb * BK, mt * BM, 0);
            tc::tma_load_2d(sa + S::A_BYTES, &tma_b, &full_bar[s], kb * BK, nt * PBN);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = epi.f16 ? tc::umma_idesc_f16(BM, PBN) : tc::umma_idesc_bf16(BM, PBN);
      int it = 0, t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, t++) {
        const int as = t & 1;
        tc::mbar_wait(&tmem_empty[as], ((t >> 1) & 1) ^ 1);      // epilogue drained this accumulator
        tc::tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * PBN);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % PSTAGES;
          const uint32_t ph = (it / PSTAGES) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + s * S::STAGE_BYTES);
          if (TRI) {
            const uint64_t a_hi = tc::umma_desc_k128(sa), a_lo = tc::umma_desc_k128(sa + S::A_BYTES);
            const uint64_t b_hi = tc::umma_desc_k128(sa + 2 * S::A_BYTES), b_lo = tc::umma_desc_k128(sa + 2 * S::A_BYTES + S::B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; k++) tc::umma_f16(tacc, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < BK / 16; k++) tc::umma_f16(tacc, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < BK / 16; k++) tc::umma_f16(tacc, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
          } else {
            const uint64_t adesc = tc::umma_desc_k128(sa);
            const uint64_t bdesc = tc::umma_desc_k128(sa + S::A_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; k++)
              tc::umma_f16(tacc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
          }
          tc::umma_commit(&empty_bar[s]);
        }
        tc::umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // 8 epilogue warps: warp%4 selects the TMEM lane quarter (hardware rule), (warp-2)/4 the column half
    const int q = warp & 3, half = (warp - 2) >> 2;
    int t = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, t++) {
      const int mt = tile / tiles_n, nt = tile - mt * tiles_n;
      const int as = t & 1;
      const int row = mt * BM + q * 32 + lane;
      const bool row_ok = row < M;
      const int bidx = row / epi.rows_per_batch;
      const int colh = nt * PBN + half * (PBN / 2);
      // the residual this tile updates does not depend on its accumulator: fetch it while the MMAs run
      float pre[PBN / 2];
      const bool use_pre = MODE == EPI_RESID_GATE && row_ok && colh + PBN / 2 <= N;
      if (use_pre) {
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(epi.out) + (size_t)row * epi.ldo + colh);
#pragma unroll
        for (int j = 0; j < PBN / 8; j++) { const float4 x4 = src[j]; pre[4 * j] = x4.x; pre[4 * j + 1] = x4.y; pre[4 * j + 2] = x4.z; pre[4 * j + 3] = x4.w; }
      }
      tc::mbar_wait(&tmem_full[as], (t >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int c = 0; c < PBN / 2; c += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PBN + half * (PBN / 2) + c), v);
        tc::tmem_ld_wait();
        const int col0 = colh + c;
        if (!row_ok || col0 >= N) continue;
        epi_store<MODE, ACT, true>(epi, row, bidx, col0, N, v, use_pre ? &pre[c] : nullptr);
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty[as]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) { __syncwarp(); tc::tmem_dealloc(tmem_base, 2 * PBN); }
}

template <int MODE, int ACT, bool TRI>
static hvx_status launch_gemm_persist_t(hvx_engine* e, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N,
                                        int K, const GemmEpi& epi, const GemmAddr& ad) {
  using S = PersistSmem<TRI>;
  static bool attr_set = false;
  if (!attr_set) {
    HVX_CUDA(cudaFuncSetAttribute(gemm_persist_kernel<MODE, ACT, TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  const int tiles_m = cdiv(M, BM), tiles_n = cdiv(N, PBN);
  const int grid = std::min(tiles_m * tiles_n, std::max(2, e->sm_count - e->sm_reserve));
  HVX_CUDA(launch_pdl_ex(gemm_persist_kernel<MODE, ACT, TRI>, dim3(grid), dim3(PERSIST_THREADS), S::TOTAL, st, ad.no_pdl ? -1 : 1, ta, tb, M, N, K, epi, ad, tiles_m, tiles_n));
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}


// ------------------------------------------------------------------ CTA-pair variant of the three-term persistent GEMM
// cta_group::2: the two CTAs of a cluster compute one 256 x 256 output tile with tcgen05.mma M = 256.  CTA r keeps rows
// [128 r, 128 r + 128) of A (hi and lo) and rows [128 r, 128 r + 128) of the weight tile (hi and lo) in its own shared memory —
// 64 KB per k-block and CTA for 3 x 128 x 256 x 64 MACs, against 96 KB for the single-CTA tile (the GEMMs are L2 -> SM
// bandwidth bound) — and its 128 accumulator lanes in its own TMEM (double-buffered: 2 x 256 columns).  Only the leader issues
// MMAs; both CTAs run a TMA producer (bytes credited to the leader's full barrier) and 8 epilogue warps over their own rows.
constexpr int PAIR_STAGES = 3;
struct PairSmem {
  static constexpr int A_BYTES = BM * BK * 2;            // 128 rows of A (hi or lo)
  static constexpr int B_BYTES = 128 * BK * 2;           // this CTA's half of the 256-row weight tile (hi or lo)
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int BAR_OFF = PAIR_STAGES * STAGE_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256 + 1024;
};

template <int MODE, int ACT>
__global__ void __launch_bounds__(PERSIST_THREADS, 1)
gemm_pair3_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, int M, int N, int K,
                  GemmEpi epi, GemmAddr ad, int tiles_m, int tiles_n) {
  using S = PairSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);      // leader's copy is the live one
  uint64_t* empty_bar = full_bar + PAIR_STAGES;                             // per CTA, signalled by multicast commits
  uint64_t* tmem_full = empty_bar + PAIR_STAGES;                            // [2] per CTA, multicast commits
  uint64_t* tmem_empty = tmem_full + 2;                                     // [2] leader's copy: 16 arrivals (8 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = tc::cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int nkb = ad.split3_kb;
  const int n_tiles = tiles_m * tiles_n;                                    // tiles of 256 x 256

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_b);
    for (int s = 0; s < PAIR_STAGES; s++) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; s++) { tc::mbar_init(&tmem_full[s], 1); tc::mbar_init(&tmem_empty[s], 16); }
    tc::fence_barrier_init();
  }
  if (warp == 1) tc::tmem_alloc_pair(tmem_slot, 2 * PBN);
  tc::tc_fence_before();
  pdl_trigger();
  tc::cluster_sync_all();            // barriers of both CTAs initialised before any remote arrive / TMA completion; also a CTA barrier
  tc::tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const int mt = tile / tiles_n, nt = tile - mt * tiles_n;
        const int row0 = mt * 256 + (int)rank * 128, brow0 = nt * PBN + (int)rank * 128;
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % PAIR_STAGES;
          const uint32_t ph = (it / PAIR_STAGES) & 1;
          tc::mbar_wait(&empty_bar[s], ph ^ 1);
          if (rank == 0) tc::mbar_expect_tx(&full_bar[s], 2 * S::STAGE_BYTES);      // both CTAs' loads land on this barrier
          uint8_t* sa = smem + s * S::STAGE_BYTES;
          tc::tma_load_3d_pair(sa, &tma_a, &full_bar[s], kb * BK, row0, 0);
          tc::tma_load_3d_pair(sa + S::A_BYTES, &tma_a, &full_bar[s], ad.a_lo_off + kb * BK, row0, 0);
          tc::tma_load_2d_pair(sa + 2 * S::A_BYTES, &tma_b, &full_bar[s], kb * BK, brow0);
          tc::tma_load_2d_pair(sa + 2 * S::A_BYTES + S::B_BYTES, &tma_b, &full_bar[s], (ad.split3_kb + kb) * BK, brow0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = epi.f16 ? tc::umma_idesc_f16(256, PBN) : tc::umma_idesc_bf16(256, PBN);
      int it = 0, t = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs, t++) {
        const int as = t & 1;
        tc::mbar_wait(&tmem_empty[as], ((t >> 1) & 1) ^ 1);       // both CTAs' epilogues drained this accumulator
        tc::tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * PBN);
        for (int kb = 0; kb < nkb; kb++, it++) {
          const int s = it % PAIR_STAGES;
          const uint32_t ph = (it / PAIR_STAGES) & 1;
          tc::mbar_wait(&full_bar[s], ph);
          tc::tc_fence_after();
          const uint32_t sa = tc::smem_u32(smem + s * S::STAGE_BYTES);
          const uint64_t a_hi = tc::umma_desc_k128(sa), a_lo = tc::umma_desc_k128(sa + S::A_BYTES);
          const uint64_t b_hi = tc::umma_desc_k128(sa + 2 * S::A_BYTES), b_lo = tc::umma_desc_k128(sa + 2 * S::A_BYTES + S::B_BYTES);
#pragma unroll
          for (int k = 0; k < BK / 16; k++) tc::umma_f16_pair(tacc, a_hi + 2 * k, b_hi + 2 * k, idesc, (kb | k) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BK / 16; k++) tc::umma_f16_pair(tacc, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);
#pragma unroll
          for (int k = 0; k < BK / 16; k++) tc::umma_f16_pair(tacc, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
          tc::umma_commit_pair(&empty_bar[s]);                    // frees this stage in both CTAs
        }
        tc::umma_commit_pair(&tmem_full[as]);                     // accumulator complete: wakes both CTAs' epilogue warps
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    int t = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, t++) {
      const int mt = tile / tiles_n, nt = tile - mt * tiles_n;
      const int as = t & 1;
      const int row = mt * 256 + (int)rank * 128 + q * 32 + lane;
      const bool row_ok = row < M;
      const int bidx = row / epi.rows_per_batch;
      const int colh = nt * PBN + half * (PBN / 2);
      float pre[PBN / 2];
      const bool use_pre = MODE == EPI_RESID_GATE && row_ok && colh + PBN / 2 <= N;
      if (use_pre) {
        const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(epi.out) + (size_t)row * epi.ldo + colh);
#pragma unroll
        for (int j = 0; j < PBN / 8; j++) { const float4 x4 = src[j]; pre[4 * j] = x4.x; pre[4 * j + 1] = x4.y; pre[4 * j + 2] = x4.z; pre[4 * j + 3] = x4.w; }
      }
      tc::mbar_wait(&tmem_full[as], (t >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int c = 0; c < PBN / 2; c += 32) {
        uint32_t v[32];
        tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PBN + half * (PBN / 2) + c), v);
        tc::tmem_ld_wait();
        const int col0 = colh + c;
        if (!row_ok || col0 >= N) continue;
        epi_store<MODE, ACT, true>(epi, row, bidx, col0, N, v, use_pre ? &pre[c] : nullptr);
      }
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_leader(&tmem_empty[as]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();            // nobody leaves while the peer's MMAs / commits can still touch this CTA's shared memory
  if (warp == 1) { __syncwarp(); tc::tmem_dealloc_pair(tmem_base, 2 * PBN); }
}


// the (mode, activation) pairs the path uses; anything else is a programming error
#define HVX_EPI_DISPATCH(CALL)                                                                   \
  do {                                                                                           \
    const int key = epi.mode * 8 + epi.act;                                                      \
    switch (key) {                                                                               \
      case EPI_BF16 * 8 + ACT_NONE: return CALL(EPI_BF16, ACT_NONE);                             \
      case EPI_BF16 * 8 + ACT_GELU_TANH: return CALL(EPI_BF16, ACT_GELU_TANH);                   \
      case EPI_BF16 * 8 + ACT_MISH: return CALL(EPI_BF16, ACT_MISH);                             \
      case EPI_BF16 * 8 + ACT_LRELU: return CALL(EPI_BF16, ACT_LRELU);                           \
      case EPI_BF16 * 8 + ACT_GELU_ERF: return CALL(EPI_BF16, ACT_GELU_ERF);                     \
      case EPI_F32 * 8 + ACT_NONE: return CALL(EPI_F32, ACT_NONE);                               \
      case EPI_F32 * 8 + ACT_GELU_TANH: return CALL(EPI_F32, ACT_GELU_TANH);                     \
      case EPI_F32 * 8 + ACT_MISH: return CALL(EPI_F32, ACT_MISH);                               \
      case EPI_RESID_GATE * 8 + ACT_NONE: return CALL(EPI_RESID_GATE, ACT_NONE);                 \
      case EPI_QKV * 8 + ACT_NONE: return CALL(EPI_QKV, ACT_NONE);                               \
      case EPI_LLM_QKV * 8 + ACT_NONE: return CALL(EPI_LLM_QKV, ACT_NONE);                       \
      case EPI_SWIGLU * 8 + ACT_NONE: return CALL(EPI_SWIGLU, ACT_NONE);                         \
      case EPI_HIFT * 8 + ACT_NONE: return CALL(EPI_HIFT, ACT_NONE);                             \
    }                                                                                            \
    set_error("gemm: epilogue mode %d with activation %d is not instantiated", epi.mode, epi.act); \
    return HVX_ERR_UNSUPPORTED;                                                                  \
  } while (0)

template <int MODE, int ACT>
static hvx_status launch_gemm_pair_t(hvx_engine* e, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N,
                                     int K, const GemmEpi& epi, const GemmAddr& ad) {
  using S = PairSmem;
  static bool attr_set = false;
  if (!attr_set) {
    HVX_CUDA(cudaFuncSetAttribute(gemm_pair3_kernel<MODE, ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  const int tiles_m = cdiv(M, 256), tiles_n = cdiv(N, PBN);
  const int n_pairs = std::max(1, std::min(tiles_m * tiles_n, std::max(2, e->sm_count - e->sm_reserve) / 2));
  HVX_CUDA(launch_pdl_ex(gemm_pair3_kernel<MODE, ACT>, dim3(2 * n_pairs), dim3(PERSIST_THREADS), S::TOTAL, st, 2, ta, tb, M, N, K, epi, ad, tiles_m, tiles_n));
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

static hvx_status launch_gemm_pair(hvx_engine* e, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N,
                                   int K, const GemmEpi& epi, const GemmAddr& ad) {
#define PPCALL(MD, AC) launch_gemm_pair_t<MD, AC>(e, st, ta, tb, M, N, K, epi, ad)
  HVX_EPI_DISPATCH(PPCALL);
#undef PPCALL
}

static hvx_status launch_gemm_persist(hvx_engine* e, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N,
                                      int K, const GemmEpi& epi, const GemmAddr& ad) {
  if (ad.split3_kb) {
#define PCALL3(MD, AC) launch_gemm_persist_t<MD, AC, true>(e, st, ta, tb, M, N, K, epi, ad)
    HVX_EPI_DISPATCH(PCALL3);
#undef PCALL3
  }
#define PCALL(MD, AC) launch_gemm_persist_t<MD, AC, false>(e, st, ta, tb, M, N, K, epi, ad)
  HVX_EPI_DISPATCH(PCALL);
#undef PCALL
}

template <int BN, int STAGES, int MODE, int ACT, bool TRI>
static hvx_status launch_gemm_t(hvx_engine* e, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N,
                                int K, const GemmEpi& epi, const GemmAddr& ad) {
  using S = GemmSmem<BN, STAGES, TRI>;
  static bool attr_set = false;
  if (!attr_set) {
    HVX_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, STAGES, MODE, ACT, TRI>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  const int tiles_per_batch = cdiv(ad.rows_per_batch, BM);
  dim3 grid(cdiv(N, BN), tiles_per_batch * ad.n_batch, ad.split_k);
  HVX_CUDA(launch_pdl_ex(gemm_bf16_kernel<BN, STAGES, MODE, ACT, TRI>, grid, dim3(GEMM_THREADS), S::TOTAL, st, ad.no_pdl ? -1 : 1, ta, tb, M, N, K, epi, ad, tiles_per_batch));
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

template <int BN, int STAGES, bool TRI>
static hvx_status launch_gemm(hvx_engine* e, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, int M, int N,
                              int K, const GemmEpi& epi, const GemmAddr& ad) {
#define GCALL(MD, AC) launch_gemm_t<BN, STAGES, MD, AC, TRI>(e, st, ta, tb, M, N, K, epi, ad)
  HVX_EPI_DISPATCH(GCALL);
#undef GCALL
}

hvx_status gemm_bf16(hvx_engine* e, cudaStream_t st, const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb,
                     int M, int N, int K, const GemmEpi& epi, const GemmAddr* addr) {
  HVX_CHECK(M > 0 && N > 0 && K > 0, HVX_ERR_ARG, "gemm: empty problem %dx%dx%d", M, N, K);
  GemmAddr ad;
  if (addr) ad = *addr;
  if (ad.rows_per_batch == 0) ad.rows_per_batch = M;
  if (ad.split3_kb && ad.a_lo_off == 0) ad.a_lo_off = ad.split3_kb * BK;
  if (ad.a_cols == 0) ad.a_cols = ad.split3_kb ? 2 * ad.split3_kb * BK : K;
  if (ad.a_rows == 0) ad.a_rows = ad.rows_per_batch;
  if (ad.split_k < 1) ad.split_k = 1;
  HVX_CHECK(ad.split_k == 1 || (epi.mode == EPI_F32 && !epi.resid && !epi.out2 && epi.act == ACT_NONE && ad.split_k <= (K + BK - 1) / BK &&
                                ad.split_k <= 64 && ad.split_stride >= (size_t)M * epi.ldo),
            HVX_ERR_ARG, "gemm: split-K needs a plain fp32 epilogue and room for %d partials", ad.split_k);
  HVX_CHECK(!(ad.split3_kb && ad.split_k > 1), HVX_ERR_ARG, "gemm: split-K is not combined with the three-term product");
  HVX_CHECK((lda % 8) == 0 && (ldb % 8) == 0 && ((uintptr_t)A % 16) == 0 && ((uintptr_t)B % 16) == 0, HVX_ERR_ARG,
            "gemm: operands must be 16-byte aligned with leading dims multiple of 8 (lda=%d ldb=%d)", lda, ldb);
  static const bool timeline = getenv("HVX_GEMM_TIMELINE") != nullptr;
  static unsigned long long* tl_dev = nullptr;
  if (timeline) {
    if (!tl_dev) HVX_CUDA(cudaMalloc(&tl_dev, 256 * sizeof(unsigned long long)));
    HVX_CUDA(cudaMemsetAsync(tl_dev, 0, 256 * sizeof(unsigned long long), st));
    ad.dbg = tl_dev;
  }
  struct TimelineDump {
    bool on; cudaStream_t st; const unsigned long long* d; int M, N, K;
    ~TimelineDump() {
      if (!on) return;
      unsigned long long h[256];
      cudaStreamSynchronize(st);
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      if (!h[0] || !h[131]) return;                 // not the tile kernel
      fprintf(stderr, "[gemm timeline M=%d N=%d K=%d] CTA(0,0,0): total %.2f us, accumulator complete at %.2f us\n  issue(us) / landed(us) per k-block:", M, N, K,
              (h[131] - h[0]) * 1e-3, (h[130] - h[0]) * 1e-3);
      for (int i = 0; i < 64 && h[1 + i]; i++) fprintf(stderr, " %d:%.2f/%.2f", i, (h[1 + i] - h[0]) * 1e-3, h[65 + i] ? (h[65 + i] - h[0]) * 1e-3 : -1.0);
      fprintf(stderr, "\n");
    }
  } timeline_dump{timeline, st, tl_dev, M, N, K};
  const double k_alg = ad.split3_kb ? (double)K / 3.0 : (ad.b_kb_mod ? (double)ad.b_kb_mod * BK : (double)K);
  ProfScope prof_scope(e->prof_gemm_off ? nullptr : &e->prof, st, PROF_GEMM, 2.0 * (double)M * (double)N * k_alg);
  CUtensorMap ta, tb;
  const uint64_t kB = ad.split3_kb ? (uint64_t)2 * ad.split3_kb * BK : ad.b_kb_mod ? (uint64_t)ad.b_kb_mod * BK : (uint64_t)K;      // width of the weight matrix
  HVX_CHECK(make_tmap_bf16_3d(&ta, A, ad.n_batch, ad.a_rows, ad.a_cols, lda, BM, BK), HVX_ERR_CUDA,
            "gemm: cuTensorMapEncodeTiled(A) failed");
  // big plain GEMMs (the DiT linears): persistent 128 x 256 tiles
  if (ad.split_k == 1 && ad.n_batch == 1 && ad.kb_per_tap == 0 && ad.b_kb_mod == 0 && ad.a_col0 == 0 && ad.a_col_per_ntile == 0 && ad.a_row0 == 0 &&
      N % PBN == 0 && cdiv(M, BM) * (N / PBN) >= e->sm_count / 2 && !getenv("HVX_NO_PERSIST")) {
    static const bool no_pair = getenv("HVX_NO_PAIR") != nullptr;
    if (ad.split3_kb && !no_pair && cdiv(M, 256) * (N / PBN) >= e->sm_count / 4) {
      // three-term product on CTA pairs (cta_group::2): 128-row halves of the 256-row weight tile per CTA
      HVX_CHECK(make_tmap_bf16_2d(&tb, B, N, kB, ldb, 128, BK), HVX_ERR_CUDA, "gemm: cuTensorMapEncodeTiled(B) failed");
      return launch_gemm_pair(e, st, ta, tb, M, N, K, epi, ad);
    }
    HVX_CHECK(make_tmap_bf16_2d(&tb, B, N, kB, ldb, PBN, BK), HVX_ERR_CUDA, "gemm: cuTensorMapEncodeTiled(B) failed");
    return launch_gemm_persist(e, st, ta, tb, M, N, K, epi, ad);
  }
  const int ctas128 = cdiv(N, 128) * cdiv(ad.rows_per_batch, BM) * ad.n_batch * ad.split_k;
  static const int min128 = getenv("HVX_GEMM_MIN_CTAS128") ? atoi(getenv("HVX_GEMM_MIN_CTAS128")) : 96;   // below: 128 x 64 tiles fill the SMs better
  if (N <= 64 || ad.a_col_per_ntile == 64 || ctas128 < min128) {
    HVX_CHECK(make_tmap_bf16_2d(&tb, B, N, kB, ldb, 64, BK), HVX_ERR_CUDA, "gemm: cuTensorMapEncodeTiled(B) failed");
    if (ad.split3_kb) return launch_gemm<64, 2, true>(e, st, ta, tb, M, N, K, epi, ad);        // 2 x 48 KB stages: two CTAs per SM
    return launch_gemm<64, 4, false>(e, st, ta, tb, M, N, K, epi, ad);
  }
  HVX_CHECK(make_tmap_bf16_2d(&tb, B, N, kB, ldb, 128, BK), HVX_ERR_CUDA, "gemm: cuTensorMapEncodeTiled(B) failed");
  if (ad.split3_kb) return launch_gemm<128, 3, true>(e, st, ta, tb, M, N, K, epi, ad);         // 3 x 64 KB stages: one CTA per SM
  return launch_gemm<128, 3, false>(e, st, ta, tb, M, N, K, epi, ad);
}

}  // namespace hvx

using namespace hvx;

// Diagnostic entry (tests/test_gemm_gpu.py): C = act(A*B^T + bias) with bf16 or fp32 output.
extern "C" hvx_status hvx_gemm_bf16(hvx_engine* e, const void* A, const void* B, const float* bias, void* C, int M, int N,
                                    int K, int out_f32, int act, void* stream) {
  HVX_CHECK(e, HVX_ERR_ARG, "null engine");
  GemmEpi epi;
  epi.mode = (out_f32 & 1) ? EPI_F32 : EPI_BF16;     // bit 1 of out_f32: operands are fp16
  epi.f16 = (out_f32 >> 1) & 1;
  epi.act = act;
  epi.bias = bias;
  epi.out = C;
  epi.ldo = N;
  if (out_f32 & 4) {                                 // bit 2: A (M, 2K) and B (N, 2K) are split [hi | lo]: three-term product
    GemmAddr ga; ga.split3_kb = K / 64;
    HVX_CHECK(K % 64 == 0, HVX_ERR_ARG, "gemm: the three-term product needs K %% 64 == 0");
    return gemm_bf16(e, (cudaStream_t)stream, (const __nv_bfloat16*)A, 2 * K, (const __nv_bfloat16*)B, 2 * K, M, N, 3 * K, epi, &ga);
  }
  return gemm_bf16(e, (cudaStream_t)stream, (const __nv_bfloat16*)A, K, (const __nv_bfloat16*)B, K, M, N, K, epi);
}
