// Flow stage (speech tokens -> mel) for sm_100a.
// Replaces CausalMaskedDiffWithDiT.inference (cosyvoice/flow/flow.py:367-430), PreLookaheadLayer
// (cosyvoice/transformer/upsample_encoder.py:82-103), CausalConditionalCFM.forward / solve_euler
// (cosyvoice/flow/flow_matching.py:71-124,203-228) and the DiT estimator (cosyvoice/flow/DiT/dit.py:145-176,
// DiT/modules.py) — restated in oracle/flow_ref.py.
//
// Data layout (one utterance per call, CFG rows stacked: row r<T conditional, r>=T unconditional):
//   ODE state x, mu, cond      fp32 [T][mel]      frame-major (a row is one 320 B line)
//   residual stream h          fp32 [2T][dim]
//   GEMM operands              fp16 [2T][*]       (the reference runs this stage in fp16: infer_speech_model.py:105-117)
//   V^T for attention          fp16 [2*heads*64][Tp]
//   adaLN modulations          fp32 [n_steps][depth*6*dim + 2*dim], computed once per solve (they depend on t only)
// Every dense contraction is the tcgen05/TMA GEMM of gemm.cu (fp16 in, fp32 accumulate) with the
// elementwise work fused into its epilogue (bias, Mish/GELU/leaky-relu, rotary + V transpose, gated residual);
// the two convolution families are run as implicit GEMMs through TMA addressing alone:
//   PreLookahead conv k4 / k3    = GEMM over an *overlapping-row* view (row stride = C, K = taps*C)
//   grouped causal conv k31      = one k-block per tap, A rows shifted by the tap, 64-channel groups
#include "attention.cuh"
#include "gemm.cuh"
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>

namespace hvx {

// ------------------------------------------------------------------ small kernels
// spks = Linear(normalize(emb))   (flow.py:389-390)
__global__ void flow_spk_kernel(const float* __restrict__ emb, const float* __restrict__ w, const float* __restrict__ b,
                                float* __restrict__ spks, int n_in, int n_out) {
  __shared__ float s_e[1024];
  __shared__ float s_red[32];
  float ss = 0.f;
  for (int i = threadIdx.x; i < n_in; i += blockDim.x) { float v = emb[i]; s_e[i] = v; ss += v * v; }
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); i++) tot += s_red[i];
  const float inv = 1.0f / fmaxf(sqrtf(tot), 1e-12f);                  // F.normalize eps
  for (int o = threadIdx.x; o < n_out; o += blockDim.x) {
    float acc = 0.f;
    for (int i = 0; i < n_in; i++) acc = fmaf(w[(size_t)o * n_in + i], s_e[i] * inv, acc);
    spks[o] = acc + b[o];
  }
}

// token embedding rows: fp32 copy (residual) and a fp16 copy padded to Cp columns and followed by zero
// rows (conv look-ahead pad)
__global__ void flow_embed_kernel(const int32_t* __restrict__ tok, const float* __restrict__ table, float* __restrict__ e32,
                                  __half* __restrict__ e16, int n_tok, int n_rows16, int C, int Cp, int vocab, int split) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows16 * Cp) return;
  const int r = i / Cp, c = i - r * Cp;
  float v = 0.f;
  if (r < n_tok && c < C) {
    int id = tok[r];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);               // torch.clamp(token, min=0) (flow.py:397)
    v = table[(size_t)id * C + c];
    e32[(size_t)r * C + c] = v;
  }
  const __half hi = __float2half_rn(v);
  if (split) { e16[(size_t)r * 2 * Cp + c] = hi; e16[(size_t)r * 2 * Cp + Cp + c] = __float2half_rn(v - __half2float(hi)); }
  else e16[i] = hi;
}

// mu = repeat_interleave(mu_tok, 2); cond = prompt mel | 0; x = noise^T   (flow.py:404-419, flow_matching.py:223)
__global__ void flow_init_kernel(const float* __restrict__ mu_tok, const float* __restrict__ prompt_feat,
                                 const float* __restrict__ noise, float* __restrict__ mu, float* __restrict__ cond,
                                 float* __restrict__ x, int T, int C, int mel_len1, int noise_ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * C) return;
  const int t = i / C, c = i - t * C;
  mu[i] = mu_tok[(size_t)(t >> 1) * C + c];
  cond[i] = (t < mel_len1) ? prompt_feat[i] : 0.f;
  x[i] = noise[(size_t)c * noise_ld + t];
}

// xin[r] = [x | cond | mu | spks] (DiT InputEmbedding order, dit.py:91-95); rows >= T are the CFG
// unconditional copy: same x, everything else zero (flow_matching.py:100-112)
// T = rows of the conditional block = U utterances x Tm frames; spks is per utterance [U][C]
__global__ void dit_pack_fm_kernel(const float* __restrict__ x, const float* __restrict__ cond, const float* __restrict__ mu,
                                   const float* __restrict__ spks, __half* __restrict__ xin, int T, int C, int Tm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = 4 * C;
  if (i >= 2 * T * W) return;
  const int r = i / W, j = i - r * W;
  const int t = r < T ? r : r - T;
  const int part = j / C, c = j - part * C;
  float v;
  if (part == 0) v = x[(size_t)t * C + c];
  else if (r >= T) v = 0.f;
  else if (part == 1) v = cond[(size_t)t * C + c];
  else if (part == 2) v = mu[(size_t)t * C + c];
  else v = spks[(t / Tm) * C + c];
  // split fp16 (hi | lo): the ODE state and the mel-valued conditioning reach |x| ~ 12 where one fp16 ulp is 7.8e-3
  const __half hi = __float2half_rn(v);
  xin[(size_t)r * 2 * W + j] = hi;
  xin[(size_t)r * 2 * W + W + j] = __float2half_rn(v - __half2float(hi));
}

// same from the estimator seam's channel-major (2, C, T) tensors (flow_matching.py:126-153)
__global__ void dit_pack_cm_kernel(const float* __restrict__ x, const float* __restrict__ cond, const float* __restrict__ mu,
                                   const float* __restrict__ spks, __half* __restrict__ xin, int T, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int W = 4 * C;
  if (i >= 2 * T * W) return;
  const int r = i / W, j = i - r * W;
  const int b = r / T, t = r - b * T;
  const int part = j / C, c = j - part * C;
  const size_t idx = ((size_t)b * C + c) * T + t;
  float v;
  if (part == 0) v = x[idx];
  else if (part == 1) v = cond[idx];
  else if (part == 2) v = mu[idx];
  else v = spks[b * C + c];
  const __half hi = __float2half_rn(v);
  xin[(size_t)r * 2 * W + j] = hi;
  xin[(size_t)r * 2 * W + W + j] = __float2half_rn(v - __half2float(hi));
}

// v (2T, C) frame-major -> out (2, C, T)
__global__ void dit_unpack_cm_kernel(const float* __restrict__ v, float* __restrict__ out, int T, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * T * C) return;
  const int b = i / (C * T), rem = i - b * C * T;
  const int c = rem / T, t = rem - c * T;
  out[i] = v[((size_t)b * T + t) * C + c];
}

// seam tensors arrive in the flow's serving dtype (spks.dtype, flow_matching.py:99-104): fp32 <-> fp16 / bf16 staging
__global__ void seam_cast_in_kernel(const void* __restrict__ src, float* __restrict__ dst, int n, int dtype) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = dtype == HVX_F16 ? __half2float(reinterpret_cast<const __half*>(src)[i])
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
}
__global__ void seam_cast_out_kernel(const float* __restrict__ src, void* __restrict__ dst, int n, int dtype) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (dtype == HVX_F16) reinterpret_cast<__half*>(dst)[i] = __float2half_rn(src[i]);
  else reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16(src[i]);
}

// SinusPositionEmbedding(256) -> Linear -> SiLU -> Linear, then the SiLU every adaLN applies first
// (modules.py:71-83,606-616,236).  One block per t value; a warp per output feature.
__global__ void __launch_bounds__(256) flow_time_embed_kernel(const float* __restrict__ t_dev, const float* __restrict__ freqs,
                                                               const float* __restrict__ w0, const float* __restrict__ b0,
                                                               const float* __restrict__ w2, const float* __restrict__ b2,
                                                               __half* __restrict__ st16, int dim, int split) {
  __shared__ float s_in[256];
  __shared__ float s_h[2048];
  const float t = t_dev[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < 128) {
    const float a = (1000.0f * t) * freqs[tid];
    s_in[tid] = sinf(a);
    s_in[tid + 128] = cosf(a);
  }
  __syncthreads();
  for (int j = warp; j < dim; j += 8) {
    float acc = 0.f;
    for (int i = lane; i < 256; i += 32) acc = fmaf(w0[(size_t)j * 256 + i], s_in[i], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) { const float v = acc + b0[j]; s_h[j] = v / (1.0f + expf(-v)); }
  }
  __syncthreads();
  for (int j = warp; j < dim; j += 8) {
    float acc = 0.f;
    for (int i = lane; i < dim; i += 32) acc = fmaf(w2[(size_t)j * dim + i], s_h[i], acc);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float v = acc + b2[j];
      const float y = v / (1.0f + expf(-v));
      __half* orow = st16 + (size_t)blockIdx.x * dim * (split ? 2 : 1);
      const __half hi = __float2half_rn(y);
      orow[j] = hi;
      if (split) orow[dim + j] = __float2half_rn(y - __half2float(hi));
    }
  }
}

// rotary table of the x_transformers convention: ang[t][i] = t * inv_freq[i]  (dit.py:158, modules.py:367-373)
__global__ void flow_rope_kernel(const float* __restrict__ inv_freq, float* __restrict__ cs, float* __restrict__ sn, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * 32) return;
  const float a = (float)(i >> 5) * inv_freq[i & 31];
  cs[i] = cosf(a);
  sn[i] = sinf(a);
}

// out16 = LayerNorm(h) * (1 + scale) + shift   (modules.py:230-244,262-265; eps 1e-6, no affine)
// One warp per row, D <= 1024 and a multiple of 128: a lane owns float4 columns (k*32 + lane) — 512 contiguous bytes per warp
// load instruction, 8-byte fp16 stores.
__global__ void __launch_bounds__(256) dit_ln_mod_kernel(const float* __restrict__ h, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, __half* __restrict__ out, int M,
                                                          int D, int split) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* hr = reinterpret_cast<const float4*>(h + (size_t)row * D);
  float4 v[8];
  const int n = D >> 7;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (k < n) { v[k] = hr[k * 32 + lane]; s += (v[k].x + v[k].y) + (v[k].z + v[k].w); }
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)D;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (k < n) {
      const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)D + 1e-6f);
  __half* orow = out + (size_t)row * D * (split ? 2 : 1);       // split: [hi | lo] per row (parity mode)
#pragma unroll
  for (int k = 0; k < 8; k++)
    if (k < n) {
      const int c4 = k * 32 + lane;
      const float4 sc = __ldg(reinterpret_cast<const float4*>(scale) + c4), sh = __ldg(reinterpret_cast<const float4*>(shift) + c4);
      const float y0 = (v[k].x - mean) * rstd * (1.0f + sc.x) + sh.x, y1 = (v[k].y - mean) * rstd * (1.0f + sc.y) + sh.y;
      const float y2 = (v[k].z - mean) * rstd * (1.0f + sc.z) + sh.z, y3 = (v[k].w - mean) * rstd * (1.0f + sc.w) + sh.w;
      const __half2 h01 = __floats2half2_rn(y0, y1), h23 = __floats2half2_rn(y2, y3);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&h01); pk.y = *reinterpret_cast<const uint32_t*>(&h23);
      reinterpret_cast<uint2*>(orow)[c4] = pk;
      if (split) {
        const __half2 l01 = __floats2half2_rn(y0 - __low2float(h01), y1 - __high2float(h01));
        const __half2 l23 = __floats2half2_rn(y2 - __low2float(h23), y3 - __high2float(h23));
        pk.x = *reinterpret_cast<const uint32_t*>(&l01); pk.y = *reinterpret_cast<const uint32_t*>(&l23);
        reinterpret_cast<uint2*>(orow + D)[c4] = pk;
      }
    }
}

// classifier-free guidance + Euler step (flow_matching.py:113-118)
__global__ void flow_euler_kernel(const float* __restrict__ v, float* __restrict__ x, int n, float dt, float cfg) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float d = (1.0f + cfg) * v[i] - cfg * v[n + i];
  x[i] = x[i] + dt * d;
}

// mel_out (C, T_out) = x[mel_len1:, :]^T   (flow.py:427-429)
__global__ void flow_out_kernel(const float* __restrict__ x, float* __restrict__ out, int T_out, int C, int mel_len1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T_out * C) return;
  const int c = i / T_out, t = i - c * T_out;
  out[i] = x[(size_t)(mel_len1 + t) * C + c];
}

// ------------------------------------------------------------------ state
struct FlowBlk { const __half *qkv_w, *out_w, *ff1_w, *ff2_w; const float *qkv_b, *out_b, *ff1_b, *ff2_b; };

struct FlowState {
  const float *spk_w, *spk_b, *emb, *pla1_b, *pla2_b, *tm0_w, *tm0_b, *tm2_w, *tm2_b, *in_b, *pos1_b, *pos2_b, *mod_b, *proj_b;
  const float* inv_freq;
  const __half *pla1_w, *pla2_w, *in_w, *pos1_w, *pos2_w, *mod_w, *proj_w;
  FlowBlk blk[64];
  float* freqs_dev = nullptr;
  DevBuf ws, ws_small;
  int rope_T = -1;
  int precise = 0;            // three-term split-fp16 GEMMs (hvx_config.flow_precise)
  // carve-up of ws for the current (T, U) (set by flow_plan): U utterances of T (padded) frames each, CFG rows stacked as
  // [conditional block U*T rows | unconditional block U*T rows]; batch index of a T-row slab = cfg * U + u
  int T = 0, Tp = 0, U = 1;
  const int* klen = nullptr;  // [2U] valid frames per slab (device), nullptr when every slab is full (U == 1)
  double attn_work = 0;       // FLOP of one attention call of the current plan: 2 CFG rows * heads * 4 * 64 * sum_u T_u^2
  __half *xin, *h0h, *c1, *n16, *qk, *vt, *ao, *f1;
  float *h0, *h, *v, *rope_c, *rope_s;
  // solve-level buffers (ws_small)
  float *spks, *e32, *mu_tok, *mu, *cond, *x, *mod, *t_dev;
  __half *e16, *y1p, *st16;
  // incremental streaming session (hvx_flow_stream_*): per Euler step s and layer l the keys (rotary applied) and V^T of every
  // frame evaluated so far, per step the two operands of the causal position-embedding convolutions (their 30-frame left context)
  struct Stream {
    bool open = false;
    int S = 0, Tcap = 0, Tpcap = 0, T_done = 0;
    DevBuf buf;
    size_t step_bytes = 0, conv_bytes = 0, k_bytes = 0, v_bytes = 0;
    float *rope_c = nullptr, *rope_s = nullptr;
    __half* h0h(int s) const { return (__half*)((uint8_t*)buf.p + (size_t)s * step_bytes); }
    __half* c1(int s) const { return (__half*)((uint8_t*)buf.p + (size_t)s * step_bytes + conv_bytes); }
    __half* K(int s, int l) const { return (__half*)((uint8_t*)buf.p + (size_t)s * step_bytes + 2 * conv_bytes + (size_t)l * (k_bytes + v_bytes)); }
    __half* Vt(int s, int l) const { return (__half*)((uint8_t*)K(s, l) + k_bytes); }
  } stream;
};

template <typename T>
static hvx_status get_t(hvx_engine* e, const std::string& name, int dtype, const T** p, int64_t numel = -1) {
  const Tensor* t = e->find(HVX_STAGE_FLOW, name);
  HVX_CHECK(t && t->dtype == dtype, HVX_ERR_STATE, "flow: missing tensor %s (or wrong dtype)", name.c_str());
  HVX_CHECK(numel < 0 || t->numel() == numel, HVX_ERR_STATE, "flow: tensor %s has %lld elements, expected %lld", name.c_str(),
            (long long)t->numel(), (long long)numel);
  *p = reinterpret_cast<const T*>(t->p);
  return HVX_OK;
}

#define FLOW_GET(call) do { hvx_status _s = (call); if (_s) return _s; } while (0)

hvx_status flow_finalize(hvx_engine* e) {
  const hvx_config& c = e->cfg;
  const int dim = c.flow_dim, inner = c.flow_heads * c.flow_dim_head, ff = dim * c.flow_ff_mult, mel = c.flow_mel;
  HVX_CHECK(c.flow_dim_head == 64, HVX_ERR_UNSUPPORTED, "flow: dim_head must be 64");
  HVX_CHECK(dim % 128 == 0 && dim <= 1024 && dim / c.flow_pos_groups == 64, HVX_ERR_UNSUPPORTED,
            "flow: dim must be a multiple of 128 (<=1024) with 64-channel conv groups (dim=%d groups=%d)", dim, c.flow_pos_groups);
  HVX_CHECK(c.flow_depth <= 64 && mel % 16 == 0 && c.flow_pla_ch % 64 == 0, HVX_ERR_UNSUPPORTED, "flow: unsupported dims");
  if (!e->flow) e->flow = new FlowState();
  FlowState* f = e->flow;
  const int64_t kpos = (int64_t)c.flow_pos_k * 64;
  FLOW_GET(get_t(e, "spk.w", HVX_F32, &f->spk_w, (int64_t)mel * c.flow_spk_in));
  FLOW_GET(get_t(e, "spk.b", HVX_F32, &f->spk_b, mel));
  FLOW_GET(get_t(e, "emb", HVX_F32, &f->emb, (int64_t)c.flow_vocab * mel));
  FLOW_GET(get_t(e, "pla1.w", HVX_F16, &f->pla1_w, (c.flow_precise ? 2 : 1) * (int64_t)c.flow_pla_ch * 4 * ((mel + 63) / 64 * 64)));
  FLOW_GET(get_t(e, "pla1.b", HVX_F32, &f->pla1_b, c.flow_pla_ch));
  FLOW_GET(get_t(e, "pla2.w", HVX_F16, &f->pla2_w, (c.flow_precise ? 2 : 1) * (int64_t)mel * 3 * c.flow_pla_ch));
  FLOW_GET(get_t(e, "pla2.b", HVX_F32, &f->pla2_b, mel));
  FLOW_GET(get_t(e, "tm0.w", HVX_F32, &f->tm0_w, (int64_t)dim * 256));
  FLOW_GET(get_t(e, "tm0.b", HVX_F32, &f->tm0_b, dim));
  FLOW_GET(get_t(e, "tm2.w", HVX_F32, &f->tm2_w, (int64_t)dim * dim));
  FLOW_GET(get_t(e, "tm2.b", HVX_F32, &f->tm2_b, dim));
  const int64_t wx = c.flow_precise ? 2 : 1;          // parity mode: GEMM weights arrive as [hi | lo] (twice as wide)
  f->precise = c.flow_precise ? 1 : 0;
  FLOW_GET(get_t(e, "in.w", HVX_F16, &f->in_w, wx * dim * 4 * mel));
  FLOW_GET(get_t(e, "in.b", HVX_F32, &f->in_b, dim));
  FLOW_GET(get_t(e, "pos1.w", HVX_F16, &f->pos1_w, (c.flow_precise ? 2 : 1) * dim * kpos));
  FLOW_GET(get_t(e, "pos1.b", HVX_F32, &f->pos1_b, dim));
  FLOW_GET(get_t(e, "pos2.w", HVX_F16, &f->pos2_w, (c.flow_precise ? 2 : 1) * dim * kpos));
  FLOW_GET(get_t(e, "pos2.b", HVX_F32, &f->pos2_b, dim));
  FLOW_GET(get_t(e, "rope.inv_freq", HVX_F32, &f->inv_freq, 32));
  const int64_t nmod = (int64_t)c.flow_depth * 6 * dim + 2 * dim;
  FLOW_GET(get_t(e, "mod.w", HVX_F16, &f->mod_w, wx * nmod * dim));
  FLOW_GET(get_t(e, "mod.b", HVX_F32, &f->mod_b, nmod));
  FLOW_GET(get_t(e, "proj.w", HVX_F16, &f->proj_w, wx * mel * dim));
  FLOW_GET(get_t(e, "proj.b", HVX_F32, &f->proj_b, mel));
  for (int i = 0; i < c.flow_depth; i++) {
    const std::string p = "blk" + std::to_string(i) + ".";
    FlowBlk& b = f->blk[i];
    FLOW_GET(get_t(e, p + "qkv.w", HVX_F16, &b.qkv_w, wx * 3 * inner * dim));
    FLOW_GET(get_t(e, p + "qkv.b", HVX_F32, &b.qkv_b, 3 * inner));
    FLOW_GET(get_t(e, p + "out.w", HVX_F16, &b.out_w, wx * dim * inner));
    FLOW_GET(get_t(e, p + "out.b", HVX_F32, &b.out_b, dim));
    FLOW_GET(get_t(e, p + "ff1.w", HVX_F16, &b.ff1_w, wx * ff * dim));
    FLOW_GET(get_t(e, p + "ff1.b", HVX_F32, &b.ff1_b, ff));
    FLOW_GET(get_t(e, p + "ff2.w", HVX_F16, &b.ff2_w, wx * dim * ff));
    FLOW_GET(get_t(e, p + "ff2.b", HVX_F32, &b.ff2_b, dim));
  }
  if (!f->freqs_dev) {
    // SinusPositionEmbedding: exp(arange(128) * -(ln(1e4)/127)) in fp32 (modules.py:76-80)
    float fr[128];
    const float e32 = (float)(-(std::log(10000.0) / 127.0));
    for (int k = 0; k < 128; k++) fr[k] = expf((float)k * e32);
    HVX_CUDA(cudaMalloc(&f->freqs_dev, sizeof(fr)));
    HVX_CUDA(cudaMemcpy(f->freqs_dev, fr, sizeof(fr), cudaMemcpyHostToDevice));
  }
  return HVX_OK;
}

void flow_free(hvx_engine* e) {
  if (e->flow && e->flow->freqs_dev) cudaFree(e->flow->freqs_dev);
  delete e->flow;
  e->flow = nullptr;
}

static inline size_t al256(size_t n) { return (n + 255) & ~(size_t)255; }

// estimator workspace for T frames (2T rows)
static hvx_status flow_plan(hvx_engine* e, cudaStream_t st, int T, int U = 1) {
  FlowState* f = e->flow;
  const hvx_config& c = e->cfg;
  const int dim = c.flow_dim, inner = c.flow_heads * 64, ff = dim * c.flow_ff_mult, mel = c.flow_mel;
  const size_t M = 2 * (size_t)T * U;
  const int Tp = (T + 7) & ~7;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
  const size_t o_xin = take(M * 8 * mel * 2), o_h0 = take(M * dim * 4), o_h0h = take(M * dim * 2 * (f->precise ? 2 : 1)), o_c1 = take(M * dim * 2 * (f->precise ? 2 : 1));
  const size_t ax = f->precise ? 2 : 1;
  const size_t o_h = take(M * dim * 4), o_n = take(M * dim * 2 * ax), o_qk = take(M * 2 * inner * 2);
  const size_t o_vt = take((size_t)2 * U * inner * Tp * 2), o_ao = take(M * inner * 2 * ax), o_f1 = take(M * ff * 2 * ax);
  const size_t o_v = take(M * mel * 4), o_rc = take((size_t)T * 32 * 4), o_rs = take((size_t)T * 32 * 4);
  const bool grew = off > f->ws.bytes;
  uint8_t* w = (uint8_t*)f->ws.get(off);
  HVX_CHECK(w, HVX_ERR_CUDA, "flow: workspace allocation of %zu bytes failed", off);
  if (grew) { HVX_CUDA(cudaMemsetAsync(w, 0, f->ws.bytes, st)); f->rope_T = -1; }     // V^T pad columns must be finite
  f->xin = (__half*)(w + o_xin); f->h0 = (float*)(w + o_h0); f->h0h = (__half*)(w + o_h0h); f->c1 = (__half*)(w + o_c1);
  f->h = (float*)(w + o_h); f->n16 = (__half*)(w + o_n); f->qk = (__half*)(w + o_qk); f->vt = (__half*)(w + o_vt);
  f->ao = (__half*)(w + o_ao); f->f1 = (__half*)(w + o_f1); f->v = (float*)(w + o_v);
  f->rope_c = (float*)(w + o_rc); f->rope_s = (float*)(w + o_rs);
  if (f->T != T || f->rope_T != T || f->U != U) {
    // the carve-up moved: V^T pad columns of the new layout may hold stale non-finite bit patterns
    HVX_CUDA(cudaMemsetAsync(f->vt, 0, (size_t)2 * U * inner * Tp * 2, st));
    flow_rope_kernel<<<cdiv(T * 32, 256), 256, 0, st>>>(f->inv_freq, f->rope_c, f->rope_s, T);
    HVX_LAUNCH_CHECK(e);
    f->rope_T = T;
  }
  f->T = T; f->Tp = Tp; f->U = U; f->klen = nullptr;
  f->attn_work = 2.0 * U * c.flow_heads * 4.0 * 64.0 * (double)T * T;
  return HVX_OK;
}

// y = x W^T for the DiT linears.  Fast mode: fp16 x fp16.  Parity mode: x = [hi | lo] (row stride 2K), W = [hi | lo]
// (row stride 2K), three products A_hi*W_hi + A_lo*W_hi + A_hi*W_lo accumulated in fp32 on the tensor core.
static hvx_status flow_linear(hvx_engine* e, cudaStream_t st, const __half* x, const __half* w, int M, int N, int K, const GemmEpi& p) {
  FlowState* f = e->flow;
  if (!f->precise)
    return gemm_bf16(e, st, (const __nv_bfloat16*)x, K, (const __nv_bfloat16*)w, K, M, N, K, p);
  GemmAddr ga; ga.split3_kb = K / 64;
  return gemm_bf16(e, st, (const __nv_bfloat16*)x, 2 * K, (const __nv_bfloat16*)w, 2 * K, M, N, 3 * K, p, &ga);
}

static GemmEpi epi16(void* out, int ldo, const float* bias, int act) {
  GemmEpi p; p.mode = EPI_BF16; p.f16 = 1; p.act = act; p.bias = bias; p.out = out; p.ldo = ldo; return p;
}

// one estimator evaluation: xin (2T, 4*mel) fp16 -> v (2T, mel) fp32   (DiT.forward, dit.py:145-176)
static hvx_status flow_nfe(hvx_engine* e, cudaStream_t st, const float* mod, int streaming) {
  FlowState* f = e->flow;
  const hvx_config& c = e->cfg;
  const int T = f->T, nb = 2 * f->U, M = nb * T, dim = c.flow_dim, inner = c.flow_heads * 64, ff = dim * c.flow_ff_mult, mel = c.flow_mel;
  hvx_status rc;
  // input embedding: Linear(320 -> dim) + causal grouped conv position embedding (dit.py:76-98, modules.py:115-144)
  { GemmEpi p; p.mode = EPI_F32; p.f16 = 1; p.bias = f->in_b; p.out = f->h0; p.ldo = dim; p.out2 = (__nv_bfloat16*)f->h0h;
    if (f->precise) { p.ldo2 = 2 * dim; p.lo_off = dim; }
    if (f->precise) {
      if ((rc = flow_linear(e, st, f->xin, f->in_w, M, dim, 4 * mel, p))) return rc;
    } else {
      GemmAddr gs; gs.b_kb_mod = (4 * mel) / 64;          // A = [hi | lo] against the same weights
      if ((rc = gemm_bf16(e, st, (const __nv_bfloat16*)f->xin, 8 * mel, (const __nv_bfloat16*)f->in_w, 4 * mel, M, dim, 8 * mel, p, &gs))) return rc;
    } }
  GemmAddr ga; ga.n_batch = nb; ga.rows_per_batch = T; ga.a_cols = dim; ga.a_col_per_ntile = 64; ga.kb_per_tap = 1;
  ga.a_row0 = -(c.flow_pos_k - 1); ga.a_row_step = 1;
  const int kpos = c.flow_pos_k * 64;
  const int px = f->precise ? 2 : 1;
  int kconv = kpos;
  if (f->precise) { ga.a_cols = 2 * dim; ga.a_lo_off = dim; ga.split3_kb = c.flow_pos_k; kconv = 3 * kpos; }
  { GemmEpi p = epi16(f->c1, px * dim, f->pos1_b, ACT_MISH);
    p.lo_off = f->precise ? dim : 0;
    if ((rc = gemm_bf16(e, st, (const __nv_bfloat16*)f->h0h, px * dim, (const __nv_bfloat16*)f->pos1_w, px * kpos, M, dim, kconv, p, &ga))) return rc; }
  { GemmEpi p; p.mode = EPI_F32; p.f16 = 1; p.act = ACT_MISH; p.bias = f->pos2_b; p.out = f->h; p.ldo = dim; p.resid = f->h0;
    if ((rc = gemm_bf16(e, st, (const __nv_bfloat16*)f->c1, px * dim, (const __nv_bfloat16*)f->pos2_w, px * kpos, M, dim, kconv, p, &ga))) return rc; }
  for (int i = 0; i < c.flow_depth; i++) {
    const FlowBlk& b = f->blk[i];
    const float* m = mod + (size_t)i * 6 * dim;       // shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp
    const double ln_bytes = (double)M * dim * (4 + 2 * (f->precise ? 2 : 1));
    { ProfScope ps(&e->prof, st, PROF_LAYERNORM, ln_bytes);
      HVX_CUDA(launch_pdl_ex(dit_ln_mod_kernel, dim3(cdiv(M, 8)), dim3(256), (size_t)0, st, 1, f->h, m + dim, m, f->n16, M, dim, f->precise)); }
    HVX_LAUNCH_CHECK(e);
    { GemmEpi p; p.mode = EPI_QKV; p.f16 = 1; p.bias = b.qkv_b; p.out = f->qk; p.ldo = 2 * inner; p.n_qk = 2 * inner;
      p.vt = (__nv_bfloat16*)f->vt; p.vt_ld = f->Tp; p.T = T; p.heads = c.flow_heads; p.rows_per_batch = T;
      p.rope_cos = f->rope_c; p.rope_sin = f->rope_s;
      if ((rc = flow_linear(e, st, f->n16, b.qkv_w, M, 3 * inner, dim, p))) return rc; }
    { AttnArgs a; a.T = T; a.heads = c.flow_heads; a.n_batch = nb; a.klen = f->klen; a.chunk = streaming ? c.flow_chunk : 0; a.f16 = 1;
      a.ld_out = inner; a.out = (__nv_bfloat16*)f->ao; a.lo_off = f->precise ? inner : 0; a.work = f->attn_work * (streaming ? 0.5 : 1.0);
      if ((rc = dit_attention(e, st, (const __nv_bfloat16*)f->qk, 2 * inner, inner, (const __nv_bfloat16*)f->vt, f->Tp, a))) return rc; }
    { GemmEpi p; p.mode = EPI_RESID_GATE; p.f16 = 1; p.bias = b.out_b; p.out = f->h; p.ldo = dim; p.gate = m + 2 * dim;
      p.gate_ld = 0; p.rows_per_batch = T;
      if ((rc = flow_linear(e, st, f->ao, b.out_w, M, dim, inner, p))) return rc; }
    { ProfScope ps(&e->prof, st, PROF_LAYERNORM, ln_bytes);
      HVX_CUDA(launch_pdl_ex(dit_ln_mod_kernel, dim3(cdiv(M, 8)), dim3(256), (size_t)0, st, 1, f->h, m + 4 * dim, m + 3 * dim, f->n16, M, dim, f->precise)); }
    HVX_LAUNCH_CHECK(e);
    { GemmEpi p = epi16(f->f1, f->precise ? 2 * ff : ff, b.ff1_b, ACT_GELU_TANH);
      p.lo_off = f->precise ? ff : 0;
      if ((rc = flow_linear(e, st, f->n16, b.ff1_w, M, ff, dim, p))) return rc; }
    { GemmEpi p; p.mode = EPI_RESID_GATE; p.f16 = 1; p.bias = b.ff2_b; p.out = f->h; p.ldo = dim; p.gate = m + 5 * dim;
      p.gate_ld = 0; p.rows_per_batch = T;
      if ((rc = flow_linear(e, st, f->f1, b.ff2_w, M, dim, ff, p))) return rc; }
  }
  const float* mf = mod + (size_t)c.flow_depth * 6 * dim;     // AdaLayerNormZero_Final: (scale, shift)  (modules.py:262)
  HVX_CUDA(launch_pdl_ex(dit_ln_mod_kernel, dim3(cdiv(M, 8)), dim3(256), (size_t)0, st, 1, f->h, mf, mf + dim, f->n16, M, dim, f->precise));
  HVX_LAUNCH_CHECK(e);
  { GemmEpi p; p.mode = EPI_F32; p.f16 = 1; p.bias = f->proj_b; p.out = f->v; p.ldo = mel;
    if ((rc = flow_linear(e, st, f->n16, f->proj_w, M, mel, dim, p))) return rc; }
  return HVX_OK;
}

// One estimator evaluation restricted to the frames [t0, t0 + Tw) of a streaming session at Euler step `step`: the same operator
// sequence as flow_nfe on 2*Tw rows, with (a) the position-embedding convolutions reading their left context from the session's
// per-step operand caches, (b) the QKV epilogue writing the new keys / V^T columns into the per-(step, layer) caches at their
// absolute frames, (c) attention of the Tw window queries over all t0 + Tw cached keys under the block-causal chunk mask.  Under
// that mask (and causal convolutions) the frames before t0 do not depend on the new ones, so their cached keys / values are
// exactly what a full re-evaluation would recompute (cosyvoice/flow/DiT/dit.py:145-176 with streaming=True).
static hvx_status flow_nfe_window(hvx_engine* e, cudaStream_t st, const float* mod, int step, int t0, int Tw) {
  FlowState* f = e->flow;
  FlowState::Stream& S = f->stream;
  const hvx_config& c = e->cfg;
  const int nb = 2, M = nb * Tw, dim = c.flow_dim, inner = c.flow_heads * 64, ff = dim * c.flow_ff_mult, mel = c.flow_mel;
  const int px = f->precise ? 2 : 1, T1 = t0 + Tw;
  hvx_status rc;
  { GemmEpi p; p.mode = EPI_F32; p.f16 = 1; p.bias = f->in_b; p.out = f->h0; p.ldo = dim; p.out2 = (__nv_bfloat16*)f->h0h;
    if (f->precise) { p.ldo2 = 2 * dim; p.lo_off = dim; }
    if (f->precise) {
      if ((rc = flow_linear(e, st, f->xin, f->in_w, M, dim, 4 * mel, p))) return rc;
    } else {
      GemmAddr gs; gs.b_kb_mod = (4 * mel) / 64;
      if ((rc = gemm_bf16(e, st, (const __nv_bfloat16*)f->xin, 8 * mel, (const __nv_bfloat16*)f->in_w, 4 * mel, M, dim, 8 * mel, p, &gs))) return rc;
    } }
  // window rows of a [2][Tw][w] operand -> rows t0.. of the session's [2][Tcap][w] cache
  auto to_cache = [&](const __half* local, __half* cache, int w) -> hvx_status {
    for (int b = 0; b < nb; b++)
      HVX_CUDA(cudaMemcpyAsync(cache + ((size_t)b * S.Tcap + t0) * w, local + (size_t)b * Tw * w, (size_t)Tw * w * sizeof(__half),
                               cudaMemcpyDeviceToDevice, st));
    return HVX_OK;
  };
  if ((rc = to_cache(f->h0h, S.h0h(step), px * dim))) return rc;
  GemmAddr ga; ga.n_batch = nb; ga.rows_per_batch = Tw; ga.a_rows = S.Tcap; ga.a_cols = dim; ga.a_col_per_ntile = 64; ga.kb_per_tap = 1;
  ga.a_row0 = t0 - (c.flow_pos_k - 1); ga.a_row_step = 1;
  const int kpos = c.flow_pos_k * 64;
  int kconv = kpos;
  if (f->precise) { ga.a_cols = 2 * dim; ga.a_lo_off = dim; ga.split3_kb = c.flow_pos_k; kconv = 3 * kpos; }
  { GemmEpi p = epi16(f->c1, px * dim, f->pos1_b, ACT_MISH);
    p.lo_off = f->precise ? dim : 0;
    if ((rc = gemm_bf16(e, st, (const __nv_bfloat16*)S.h0h(step), px * dim, (const __nv_bfloat16*)f->pos1_w, px * kpos, M, dim, kconv, p, &ga))) return rc; }
  if ((rc = to_cache(f->c1, S.c1(step), px * dim))) return rc;
  { GemmEpi p; p.mode = EPI_F32; p.f16 = 1; p.act = ACT_MISH; p.bias = f->pos2_b; p.out = f->h; p.ldo = dim; p.resid = f->h0;
    if ((rc = gemm_bf16(e, st, (const __nv_bfloat16*)S.c1(step), px * dim, (const __nv_bfloat16*)f->pos2_w, px * kpos, M, dim, kconv, p, &ga))) return rc; }
  // keys visible to the window under the chunk mask: all T1 of them for its last chunk (T1 is chunk-aligned)
  const double attn_work = 2.0 * c.flow_heads * 4.0 * 64.0 * (double)Tw * (t0 + 0.5 * Tw);
  for (int i = 0; i < c.flow_depth; i++) {
    const FlowBlk& b = f->blk[i];
    const float* m = mod + (size_t)i * 6 * dim;
    HVX_CUDA(launch_pdl_ex(dit_ln_mod_kernel, dim3(cdiv(M, 8)), dim3(256), (size_t)0, st, 1, f->h, m + dim, m, f->n16, M, dim, f->precise));
    HVX_LAUNCH_CHECK(e);
    { GemmEpi p; p.mode = EPI_QKV; p.f16 = 1; p.bias = b.qkv_b; p.out = f->qk; p.ldo = inner; p.n_qk = 2 * inner;
      p.k_out = (__nv_bfloat16*)S.K(step, i); p.k_ld = inner; p.k_batch_rows = S.Tcap; p.t_off = t0;
      p.vt = (__nv_bfloat16*)S.Vt(step, i); p.vt_ld = S.Tpcap; p.T = Tw; p.heads = c.flow_heads; p.rows_per_batch = Tw;
      p.rope_cos = S.rope_c; p.rope_sin = S.rope_s;
      if ((rc = flow_linear(e, st, f->n16, b.qkv_w, M, 3 * inner, dim, p))) return rc; }
    { AttnArgs a; a.T = S.Tcap; a.heads = c.flow_heads; a.n_batch = nb; a.chunk = c.flow_chunk; a.f16 = 1;
      a.Tq = Tw; a.q_pos0 = t0; a.tk = T1; a.k_ptr = (const __nv_bfloat16*)S.K(step, i); a.ld_k = inner;
      a.ld_out = inner; a.out = (__nv_bfloat16*)f->ao; a.lo_off = f->precise ? inner : 0; a.work = attn_work;
      if ((rc = dit_attention(e, st, (const __nv_bfloat16*)f->qk, inner, 0, (const __nv_bfloat16*)S.Vt(step, i), S.Tpcap, a))) return rc; }
    { GemmEpi p; p.mode = EPI_RESID_GATE; p.f16 = 1; p.bias = b.out_b; p.out = f->h; p.ldo = dim; p.gate = m + 2 * dim;
      p.gate_ld = 0; p.rows_per_batch = Tw;
      if ((rc = flow_linear(e, st, f->ao, b.out_w, M, dim, inner, p))) return rc; }
    HVX_CUDA(launch_pdl_ex(dit_ln_mod_kernel, dim3(cdiv(M, 8)), dim3(256), (size_t)0, st, 1, f->h, m + 4 * dim, m + 3 * dim, f->n16, M, dim, f->precise));
    HVX_LAUNCH_CHECK(e);
    { GemmEpi p = epi16(f->f1, f->precise ? 2 * ff : ff, b.ff1_b, ACT_GELU_TANH);
      p.lo_off = f->precise ? ff : 0;
      if ((rc = flow_linear(e, st, f->n16, b.ff1_w, M, ff, dim, p))) return rc; }
    { GemmEpi p; p.mode = EPI_RESID_GATE; p.f16 = 1; p.bias = b.ff2_b; p.out = f->h; p.ldo = dim; p.gate = m + 5 * dim;
      p.gate_ld = 0; p.rows_per_batch = Tw;
      if ((rc = flow_linear(e, st, f->f1, b.ff2_w, M, dim, ff, p))) return rc; }
  }
  const float* mf = mod + (size_t)c.flow_depth * 6 * dim;
  HVX_CUDA(launch_pdl_ex(dit_ln_mod_kernel, dim3(cdiv(M, 8)), dim3(256), (size_t)0, st, 1, f->h, mf, mf + dim, f->n16, M, dim, f->precise));
  HVX_LAUNCH_CHECK(e);
  { GemmEpi p; p.mode = EPI_F32; p.f16 = 1; p.bias = f->proj_b; p.out = f->v; p.ldo = mel;
    if ((rc = flow_linear(e, st, f->n16, f->proj_w, M, mel, dim, p))) return rc; }
  return HVX_OK;
}

// adaLN modulations of all blocks for n t-values (t_dev on device): mod (n, depth*6*dim + 2*dim)
static hvx_status flow_mods(hvx_engine* e, cudaStream_t st, const float* t_dev, int n, __half* st16, float* mod) {
  FlowState* f = e->flow;
  const hvx_config& c = e->cfg;
  const int dim = c.flow_dim;
  const int nmod = c.flow_depth * 6 * dim + 2 * dim;
  flow_time_embed_kernel<<<n, 256, 0, st>>>(t_dev, f->freqs_dev, f->tm0_w, f->tm0_b, f->tm2_w, f->tm2_b, st16, dim, f->precise);
  HVX_LAUNCH_CHECK(e);
  GemmEpi p; p.mode = EPI_F32; p.f16 = 1; p.bias = f->mod_b; p.out = mod; p.ldo = nmod;
  return flow_linear(e, st, st16, f->mod_w, n, nmod, dim, p);
}

}  // namespace hvx

using namespace hvx;

// Pre-net of one utterance (flow.py:387-419): speaker projection, token embedding, PreLookaheadLayer (two convolutions as implicit
// GEMMs), repeat_interleave(2) into mu, prompt mel into cond, the noise slice into x — into slot u of the solve-level buffers
// (frame offset xo elements).  nt = prompt + new tokens, l1 = tokens that become frames (look-ahead tokens are context only).
static hvx_status flow_prenet(hvx_engine* e, cudaStream_t st, const int32_t* tokens, int n_prompt, int nt, int l1, const float* embedding,
                              const float* prompt_feat, const float* noise, int u, size_t xo) {
  FlowState* f = e->flow;
  const hvx_config& c = e->cfg;
  const int mel = c.flow_mel, pc = c.flow_pla_ch, px = f->precise ? 2 : 1, Cp = (mel + 63) / 64 * 64, mel_len1 = 2 * n_prompt;
  hvx_status rc;
  flow_spk_kernel<<<1, 256, 0, st>>>(embedding, f->spk_w, f->spk_b, f->spks + (size_t)u * mel, c.flow_spk_in, mel);
  HVX_LAUNCH_CHECK(e);
  flow_embed_kernel<<<cdiv((nt + 3) * Cp, 256), 256, 0, st>>>(tokens, f->emb, f->e32, f->e16, nt, nt + 3, mel, Cp, c.flow_vocab, f->precise);
  HVX_LAUNCH_CHECK(e);
  { // conv1 k4, 3 look-ahead rows (zero rows after the last token when finalize): implicit GEMM, K = 4*Cp
    GemmEpi p = epi16(f->y1p, px * pc, f->pla1_b, ACT_LRELU);
    p.lo_off = f->precise ? pc : 0;
    GemmAddr ga; ga.rows_per_batch = l1; ga.a_rows = nt + 3; ga.a_cols = px * Cp; ga.kb_per_tap = Cp / 64; ga.a_row_step = 1;
    if (f->precise) { ga.a_lo_off = Cp; ga.split3_kb = 4 * Cp / 64; }
    if ((rc = gemm_bf16(e, st, (const __nv_bfloat16*)f->e16, px * Cp, (const __nv_bfloat16*)f->pla1_w, px * 4 * Cp, l1, pc, (f->precise ? 3 : 1) * 4 * Cp, p, &ga))) return rc; }
  { // conv2 k3 causal (left pad 2 = out-of-bounds rows) + residual
    GemmEpi p; p.mode = EPI_F32; p.f16 = 1; p.bias = f->pla2_b; p.out = f->mu_tok; p.ldo = mel; p.resid = f->e32;
    GemmAddr ga; ga.rows_per_batch = l1; ga.a_cols = px * pc; ga.kb_per_tap = pc / 64; ga.a_row0 = -2; ga.a_row_step = 1;
    if (f->precise) { ga.a_lo_off = pc; ga.split3_kb = 3 * pc / 64; }
    if ((rc = gemm_bf16(e, st, (const __nv_bfloat16*)f->y1p, px * pc, (const __nv_bfloat16*)f->pla2_w, px * 3 * pc, l1, mel, (f->precise ? 3 : 1) * 3 * pc, p, &ga))) return rc; }
  flow_init_kernel<<<cdiv(2 * l1 * mel, 256), 256, 0, st>>>(f->mu_tok, prompt_feat, noise, f->mu + xo, f->cond + xo, f->x + xo, 2 * l1, mel,
                                                           mel_len1, c.flow_noise_frames);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

// fp32 t-values and step sizes of the cosine schedule, carried exactly like solve_euler (flow_matching.py:93-122,225-227)
static void flow_schedule(int n_timesteps, float* tv, float* dtv) {
  float ts[65];
  for (int i = 0; i <= n_timesteps; i++) {
    const float step = 1.0f / (float)n_timesteps;
    const float lin = (i < (n_timesteps + 1) / 2) ? (float)i * step : 1.0f - (float)(n_timesteps - i) * step;   // torch.linspace
    ts[i] = 1.0f - cosf((lin * 0.5f) * 3.14159265358979323846f);
  }
  float t = ts[0], dt = ts[1] - ts[0];
  for (int s = 1; s <= n_timesteps; s++) {
    tv[s - 1] = t; dtv[s - 1] = dt;
    t = t + dt;
    if (s < n_timesteps) dt = ts[s + 1] - t;
  }
}

// solve-level buffers for U utterances of up to Tm frames / ntok_max tokens and n_timesteps modulation sets (ws_small)
static hvx_status flow_solve_bufs(hvx_engine* e, int U, int Tm, int ntok_max, int n_timesteps, int** klen_dev) {
  FlowState* f = e->flow;
  const hvx_config& c = e->cfg;
  const int mel = c.flow_mel, dim = c.flow_dim, pc = c.flow_pla_ch;
  const int nmod = c.flow_depth * 6 * dim + 2 * dim;
  const size_t nx = (size_t)U * Tm * mel;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
  const int Cp = (mel + 63) / 64 * 64;                              // embedding rows padded to whole 64-wide k-blocks
  const size_t o_spk = take((size_t)U * mel * 4), o_e32 = take((size_t)ntok_max * mel * 4), o_e16 = take((size_t)(ntok_max + 3) * Cp * 4);
  const size_t o_y1 = take((size_t)ntok_max * pc * 4), o_mut = take((size_t)ntok_max * mel * 4);
  const size_t o_mu = take(nx * 4), o_cond = take(nx * 4), o_x = take(nx * 4);
  const size_t o_mod = take((size_t)n_timesteps * nmod * 4), o_t = take(64 * 4), o_st = take((size_t)64 * dim * 4), o_kl = take((size_t)2 * U * 4);
  uint8_t* w = (uint8_t*)f->ws_small.get(off);
  HVX_CHECK(w, HVX_ERR_CUDA, "flow: buffer allocation of %zu bytes failed", off);
  f->spks = (float*)(w + o_spk); f->e32 = (float*)(w + o_e32); f->e16 = (__half*)(w + o_e16); f->y1p = (__half*)(w + o_y1);
  f->mu_tok = (float*)(w + o_mut); f->mu = (float*)(w + o_mu); f->cond = (float*)(w + o_cond); f->x = (float*)(w + o_x);
  f->mod = (float*)(w + o_mod); f->t_dev = (float*)(w + o_t); f->st16 = (__half*)(w + o_st);
  *klen_dev = (int*)(w + o_kl);
  return HVX_OK;
}

// One CFM solve for U utterances at once (U == 1: the reference's call shape).  Utterance u occupies frames [u*Tm, u*Tm + T_u) of
// the ODE state; rows beyond T_u are zero padding that stays finite and is never read by a valid row: every op of the estimator
// is row-wise, causal in time (the position-embedding conv) or key-masked per utterance (attention, klen).
static hvx_status flow_solve(hvx_engine* e, int U, const int32_t* const* tokens, const int* n_prompt, const int* n_tok,
                             const float* const* embedding, const float* const* prompt_feat, const float* noise, int n_timesteps,
                             int streaming, int finalize, float* const* mel_out, cudaStream_t st) {
  FlowState* f = e->flow;
  const hvx_config& c = e->cfg;
  const int mel = c.flow_mel, dim = c.flow_dim, pc = c.flow_pla_ch;
  const int nmod = c.flow_depth * 6 * dim + 2 * dim;
  std::vector<int> ntok(U), L1(U), Tu(U);
  int Tm = 0, ntok_max = 0;
  for (int u = 0; u < U; u++) {
    HVX_CHECK(tokens[u] && embedding[u] && mel_out[u], HVX_ERR_ARG, "flow: null argument (utterance %d)", u);
    HVX_CHECK(n_prompt[u] == 0 || prompt_feat[u], HVX_ERR_ARG, "flow: prompt tokens without prompt_feat (utterance %d)", u);
    ntok[u] = n_prompt[u] + n_tok[u];
    L1[u] = finalize ? ntok[u] : ntok[u] - 3;                     // look-ahead tokens are context only (flow.py:399-403)
    HVX_CHECK(L1[u] >= 1 && n_tok[u] - (finalize ? 0 : 3) >= 1, HVX_ERR_ARG, "flow: too few tokens (%d)", n_tok[u]);
    Tu[u] = 2 * L1[u];
    HVX_CHECK(Tu[u] <= c.flow_noise_frames, HVX_ERR_ARG, "flow: %d frames exceed the noise table (%d)", Tu[u], c.flow_noise_frames);
    Tm = std::max(Tm, Tu[u]); ntok_max = std::max(ntok_max, ntok[u]);
  }
  const size_t nx = (size_t)U * Tm * mel;                           // elements of the ODE state

  int* klen_dev = nullptr;
  hvx_status rc;
  if ((rc = flow_solve_bufs(e, U, Tm, ntok_max, n_timesteps, &klen_dev))) return rc;
  if ((rc = flow_plan(e, st, Tm, U))) return rc;
  if (U > 1) {
    HVX_CUDA(cudaMemsetAsync(f->mu, 0, (size_t)((uint8_t*)f->x - (uint8_t*)f->mu) + al256(nx * 4), st));      // mu | cond | x are adjacent: padding frames = 0
    std::vector<int> kl(2 * U);
    for (int u = 0; u < U; u++) kl[u] = kl[U + u] = Tu[u];
    HVX_CUDA(cudaMemcpyAsync(klen_dev, kl.data(), sizeof(int) * 2 * U, cudaMemcpyHostToDevice, st));
    f->klen = klen_dev;
    f->attn_work = 0;
    for (int u = 0; u < U; u++) f->attn_work += 2.0 * c.flow_heads * 4.0 * 64.0 * (double)Tu[u] * Tu[u];
  }

  // ---- pre-net, utterance by utterance (flow.py:387-419); a few small launches each
  for (int u = 0; u < U; u++)
    if ((rc = flow_prenet(e, st, tokens[u], n_prompt[u], ntok[u], L1[u], embedding[u], prompt_feat[u], noise, u, (size_t)u * Tm * mel))) return rc;

  float tv[64], dtv[64];
  flow_schedule(n_timesteps, tv, dtv);
  HVX_CUDA(cudaMemcpyAsync(f->t_dev, tv, sizeof(float) * n_timesteps, cudaMemcpyHostToDevice, st));
  if ((rc = flow_mods(e, st, f->t_dev, n_timesteps, f->st16, f->mod))) return rc;

  // ---- Euler steps over all utterances of the group
  const int Tall = U * Tm;
  for (int s = 0; s < n_timesteps; s++) {
    dit_pack_fm_kernel<<<cdiv(2 * Tall * 4 * mel, 256), 256, 0, st>>>(f->x, f->cond, f->mu, f->spks, f->xin, Tall, mel, Tm);
    HVX_LAUNCH_CHECK(e);
    if ((rc = flow_nfe(e, st, f->mod + (size_t)s * nmod, streaming))) return rc;
    flow_euler_kernel<<<cdiv(Tall * mel, 256), 256, 0, st>>>(f->v, f->x, Tall * mel, dtv[s], c.flow_cfg_rate);
    HVX_LAUNCH_CHECK(e);
  }
  for (int u = 0; u < U; u++) {
    const int mel_len1 = 2 * n_prompt[u], T_out = Tu[u] - mel_len1;
    flow_out_kernel<<<cdiv(T_out * mel, 256), 256, 0, st>>>(f->x + (size_t)u * Tm * mel, mel_out[u], T_out, mel, mel_len1);
    HVX_LAUNCH_CHECK(e);
  }
  return HVX_OK;
}

extern "C" hvx_status hvx_flow_inference(hvx_engine* e, const int32_t* tokens, int n_prompt, int n_tok, const float* embedding,
                                         const float* prompt_feat, const float* noise, int n_timesteps, int streaming,
                                         int finalize, float* mel_out, void* stream) {
  HVX_CHECK(e && e->flow, HVX_ERR_STATE, "flow stage not finalized");
  HVX_LOCK(e, HVX_STAGE_FLOW);
  HVX_CHECK(tokens && embedding && noise && mel_out, HVX_ERR_ARG, "flow: null argument");
  HVX_CHECK(n_timesteps >= 1 && n_timesteps <= 64, HVX_ERR_ARG, "flow: n_timesteps=%d out of range [1,64]", n_timesteps);
  return flow_solve(e, 1, &tokens, &n_prompt, &n_tok, &embedding, &prompt_feat, noise, n_timesteps, streaming, finalize, &mel_out,
                    (cudaStream_t)stream);
}

extern "C" hvx_status hvx_flow_inference_batch(hvx_engine* e, int n_utt, const int32_t* const* tokens, const int* n_prompt,
                                               const int* n_tok, const float* const* embedding, const float* const* prompt_feat,
                                               const float* noise, int n_timesteps, int streaming, int finalize,
                                               float* const* mel_out, void* stream) {
  HVX_CHECK(e && e->flow, HVX_ERR_STATE, "flow stage not finalized");
  HVX_LOCK(e, HVX_STAGE_FLOW);
  HVX_CHECK(tokens && n_prompt && n_tok && embedding && prompt_feat && noise && mel_out, HVX_ERR_ARG, "flow: null argument");
  HVX_CHECK(n_utt >= 1 && n_utt <= 64, HVX_ERR_ARG, "flow: %d utterances per solve (1..64)", n_utt);
  HVX_CHECK(n_timesteps >= 1 && n_timesteps <= 64, HVX_ERR_ARG, "flow: n_timesteps=%d out of range [1,64]", n_timesteps);
  return flow_solve(e, n_utt, tokens, n_prompt, n_tok, embedding, prompt_feat, noise, n_timesteps, streaming, finalize, mel_out,
                    (cudaStream_t)stream);
}

// ------------------------------------------------------------------ incremental streaming flow
// The reference re-runs flow.inference(streaming=True, finalize=False) over ALL tokens so far for every 25-token hop and keeps
// the new mel frames (cosyvoice/cli/model.py:279-297,330-348).  Under the block-causal chunk mask those calls recompute the same
// values for every earlier frame, so a session keeps what later frames need of them — per Euler step and layer the keys and V^T,
// per step the operands of the causal position-embedding convolutions — and evaluates only the new frames.
extern "C" hvx_status hvx_flow_stream_begin(hvx_engine* e, int n_timesteps, int max_frames, void* stream) {
  HVX_CHECK(e && e->flow, HVX_ERR_STATE, "flow stage not finalized");
  HVX_LOCK(e, HVX_STAGE_FLOW);
  const hvx_config& c = e->cfg;
  HVX_CHECK(n_timesteps >= 1 && n_timesteps <= 64, HVX_ERR_ARG, "flow_stream: n_timesteps=%d out of range [1,64]", n_timesteps);
  HVX_CHECK(c.flow_chunk > 0, HVX_ERR_UNSUPPORTED, "flow_stream: the flow has no streaming chunk mask (static_chunk_size = 0)");
  HVX_CHECK(max_frames >= c.flow_chunk && max_frames <= c.flow_noise_frames, HVX_ERR_ARG, "flow_stream: max_frames=%d outside [%d, %d]", max_frames,
            c.flow_chunk, c.flow_noise_frames);
  FlowState* f = e->flow;
  FlowState::Stream& S = f->stream;
  cudaStream_t st = (cudaStream_t)stream;
  const int dim = c.flow_dim, inner = c.flow_heads * 64, px = f->precise ? 2 : 1;
  S.S = n_timesteps;
  S.Tcap = (max_frames + 127) & ~127;                        // whole key tiles: a 64-key TMA box never leaves the batch row
  S.Tpcap = S.Tcap;
  S.conv_bytes = al256((size_t)2 * S.Tcap * px * dim * sizeof(__half));
  S.k_bytes = al256((size_t)2 * S.Tcap * inner * sizeof(__half));
  S.v_bytes = al256((size_t)2 * inner * S.Tpcap * sizeof(__half));
  S.step_bytes = 2 * S.conv_bytes + (size_t)c.flow_depth * (S.k_bytes + S.v_bytes);
  const size_t rope_bytes = al256((size_t)S.Tcap * 32 * sizeof(float));
  const size_t total = (size_t)n_timesteps * S.step_bytes + 2 * rope_bytes;
  uint8_t* w = (uint8_t*)S.buf.get(total);
  HVX_CHECK(w, HVX_ERR_CUDA, "flow_stream: cache allocation of %.1f GB failed (%d steps x %d frames)", total / 1e9, n_timesteps, S.Tcap);
  // keys / V^T columns past the frames seen so far are covered by the last key tile: masked, but they must be finite
  HVX_CUDA(cudaMemsetAsync(w, 0, (size_t)n_timesteps * S.step_bytes, st));
  S.rope_c = (float*)(w + (size_t)n_timesteps * S.step_bytes);
  S.rope_s = (float*)(w + (size_t)n_timesteps * S.step_bytes + rope_bytes);
  flow_rope_kernel<<<cdiv(S.Tcap * 32, 256), 256, 0, st>>>(f->inv_freq, S.rope_c, S.rope_s, S.Tcap);
  HVX_LAUNCH_CHECK(e);
  S.T_done = 0;
  S.open = true;
  return HVX_OK;
}

// tokens_dev: prompt + all new tokens so far INCLUDING the 3 look-ahead tokens (what the reference hands to flow.inference with
// finalize=False).  Writes the mel frames that are new since the previous call — frames [max(T_done, 2*n_prompt), 2*(n_prompt +
// n_tok - 3)) — to mel_out_dev as (mel, n_new) and returns n_new through n_new_frames (host).
extern "C" hvx_status hvx_flow_stream_append(hvx_engine* e, const int32_t* tokens, int n_prompt, int n_tok, const float* embedding,
                                             const float* prompt_feat, const float* noise, float* mel_out, int* n_new_frames, void* stream) {
  HVX_CHECK(e && e->flow, HVX_ERR_STATE, "flow stage not finalized");
  HVX_LOCK(e, HVX_STAGE_FLOW);
  FlowState* f = e->flow;
  FlowState::Stream& S = f->stream;
  const hvx_config& c = e->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  HVX_CHECK(S.open, HVX_ERR_STATE, "flow_stream_append: no open session (hvx_flow_stream_begin)");
  HVX_CHECK(tokens && embedding && noise && mel_out && n_new_frames, HVX_ERR_ARG, "flow_stream_append: null argument");
  HVX_CHECK(n_prompt == 0 || prompt_feat, HVX_ERR_ARG, "flow_stream_append: prompt tokens without prompt_feat");
  const int mel = c.flow_mel, dim = c.flow_dim, nt = n_prompt + n_tok, l1 = nt - 3, T1 = 2 * l1, t0 = S.T_done, Tw = T1 - t0;
  const int nmod = c.flow_depth * 6 * dim + 2 * dim;
  HVX_CHECK(n_tok - 3 >= 1 && Tw > 0, HVX_ERR_ARG, "flow_stream_append: %d tokens add no frame to the %d already done", n_tok, t0);
  HVX_CHECK(T1 % c.flow_chunk == 0, HVX_ERR_ARG, "flow_stream_append: %d frames are not whole %d-frame chunks (the caller pads the first hop, "
            "cli/model.py:332-335)", T1, c.flow_chunk);
  HVX_CHECK(T1 <= S.Tcap, HVX_ERR_ARG, "flow_stream_append: %d frames exceed the session's %d", T1, S.Tcap);
  int* klen_dev = nullptr;
  hvx_status rc;
  if ((rc = flow_solve_bufs(e, 1, T1, nt, S.S, &klen_dev))) return rc;
  if ((rc = flow_plan(e, st, Tw, 1))) return rc;                       // window-sized activations
  if ((rc = flow_prenet(e, st, tokens, n_prompt, nt, l1, embedding, prompt_feat, noise, 0, 0))) return rc;
  float tv[64], dtv[64];
  flow_schedule(S.S, tv, dtv);
  HVX_CUDA(cudaMemcpyAsync(f->t_dev, tv, sizeof(float) * S.S, cudaMemcpyHostToDevice, st));
  if ((rc = flow_mods(e, st, f->t_dev, S.S, f->st16, f->mod))) return rc;
  float* xw = f->x + (size_t)t0 * mel;
  for (int s = 0; s < S.S; s++) {
    dit_pack_fm_kernel<<<cdiv(2 * Tw * 4 * mel, 256), 256, 0, st>>>(xw, f->cond + (size_t)t0 * mel, f->mu + (size_t)t0 * mel, f->spks, f->xin, Tw, mel, Tw);
    HVX_LAUNCH_CHECK(e);
    if ((rc = flow_nfe_window(e, st, f->mod + (size_t)s * nmod, s, t0, Tw))) return rc;
    flow_euler_kernel<<<cdiv(Tw * mel, 256), 256, 0, st>>>(f->v, xw, Tw * mel, dtv[s], c.flow_cfg_rate);
    HVX_LAUNCH_CHECK(e);
  }
  const int first = std::max(t0, 2 * n_prompt), T_out = T1 - first;
  if (T_out > 0) {
    flow_out_kernel<<<cdiv(T_out * mel, 256), 256, 0, st>>>(f->x, mel_out, T_out, mel, first);
    HVX_LAUNCH_CHECK(e);
  }
  *n_new_frames = std::max(T_out, 0);
  S.T_done = T1;
  return HVX_OK;
}

extern "C" hvx_status hvx_flow_stream_end(hvx_engine* e) {
  HVX_CHECK(e && e->flow, HVX_ERR_STATE, "flow stage not finalized");
  HVX_LOCK(e, HVX_STAGE_FLOW);
  e->flow->stream.open = false;
  e->flow->stream.T_done = 0;
  return HVX_OK;
}

extern "C" hvx_status hvx_dit_estimator(hvx_engine* e, const float* x, const float* mu, const float* t, const float* spks,
                                        const float* cond, int T, int streaming, float* out, void* stream) {
  HVX_CHECK(e && e->flow, HVX_ERR_STATE, "flow stage not finalized");
  HVX_LOCK(e, HVX_STAGE_FLOW);
  HVX_CHECK(x && mu && t && spks && cond && out && T >= 1, HVX_ERR_ARG, "estimator: bad argument");
  FlowState* f = e->flow;
  const hvx_config& c = e->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  const int mel = c.flow_mel, dim = c.flow_dim;
  const int nmod = c.flow_depth * 6 * dim + 2 * dim;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += al256(bytes); return o; };
  const size_t o_mod = take((size_t)nmod * 4), o_st = take((size_t)64 * dim * 4);
  uint8_t* w = (uint8_t*)f->ws_small.get(off);
  HVX_CHECK(w, HVX_ERR_CUDA, "estimator: buffer allocation failed");
  float* mod = (float*)(w + o_mod);
  __half* st16 = (__half*)(w + o_st);
  hvx_status rc;
  if ((rc = flow_plan(e, st, T))) return rc;
  if ((rc = flow_mods(e, st, t, 1, st16, mod))) return rc;         // both CFG rows share t (flow_matching.py:104)
  dit_pack_cm_kernel<<<cdiv(2 * T * 4 * mel, 256), 256, 0, st>>>(x, cond, mu, spks, f->xin, T, mel);
  HVX_LAUNCH_CHECK(e);
  if ((rc = flow_nfe(e, st, mod, streaming))) return rc;
  dit_unpack_cm_kernel<<<cdiv(2 * T * mel, 256), 256, 0, st>>>(f->v, out, T, mel);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

// The reference's own estimator plug-in seam, exactly as ConditionalCFM.forward_estimator drives a TensorRT context
// (cosyvoice/flow/flow_matching.py:126-153): raw device addresses of x (2, mel, T), mu, t (2), spks (2, mel), cond in the flow's
// serving dtype (spks.dtype: fp32, fp16 or bf16), the result written to out_dev — which the reference binds to x itself (7th
// address = x.data_ptr(): in place).  kind 0 = DiT estimator (stage FLOW), 1 = U-Net estimator (stage UNET).
extern "C" hvx_status hvx_estimator_seam(hvx_engine* e, int kind, const void* x, const void* mu, const void* t, const void* spks,
                                         const void* cond, void* out, int T, int dtype, int streaming, void* stream) {
  HVX_CHECK(e && (kind == 0 ? (e->flow != nullptr) : (e->unet != nullptr)), HVX_ERR_STATE, "estimator seam: stage not finalized");
  HVX_CHECK(x && mu && t && spks && cond && out && T >= 1, HVX_ERR_ARG, "estimator seam: bad argument");
  HVX_CHECK(dtype == HVX_F32 || dtype == HVX_F16 || dtype == HVX_BF16, HVX_ERR_ARG, "estimator seam: dtype %d (fp32, fp16 or bf16)", dtype);
  auto run = [&](const float* x32, const float* mu32, const float* t32, const float* s32, const float* c32, float* o32) {
    return kind == 0 ? hvx_dit_estimator(e, x32, mu32, t32, s32, c32, T, streaming, o32, stream)
                     : hvx_unet_estimator(e, x32, mu32, t32, s32, c32, T, streaming, o32, stream);
  };
  if (dtype == HVX_F32)          // in place is safe: x is consumed by the pack kernel before the unpack kernel writes (stream order)
    return run((const float*)x, (const float*)mu, (const float*)t, (const float*)spks, (const float*)cond, (float*)out);
  const int mel = kind == 0 ? e->cfg.flow_mel : e->cfg.unet_mel;
  const size_t nx = (size_t)2 * mel * T, ns = (size_t)2 * mel;
  static DevBuf seam_ws;         // one engine per process
  float* w = (float*)seam_ws.get((4 * nx + ns + 2 + 64) * sizeof(float));
  HVX_CHECK(w, HVX_ERR_CUDA, "estimator seam: staging allocation failed");
  float *x32 = w, *mu32 = w + nx, *c32 = w + 2 * nx, *o32 = w + 3 * nx, *s32 = w + 4 * nx, *t32 = s32 + ns;
  cudaStream_t st = (cudaStream_t)stream;
  const void* src[5] = {x, mu, cond, spks, t};
  float* dst[5] = {x32, mu32, c32, s32, t32};
  const size_t cnt[5] = {nx, nx, nx, ns, 2};
  for (int i = 0; i < 5; i++) {
    seam_cast_in_kernel<<<cdiv((int)cnt[i], 256), 256, 0, st>>>(src[i], dst[i], (int)cnt[i], dtype);
    HVX_LAUNCH_CHECK(e);
  }
  hvx_status rc = run(x32, mu32, t32, s32, c32, o32);
  if (rc) return rc;
  seam_cast_out_kernel<<<cdiv((int)nx, 256), 256, 0, st>>>(o32, out, (int)nx, dtype);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}
