// HiFT vocoder stage (mel -> waveform) for sm_100a.
// Replaces CausalHiFTGenerator.inference / decode (cosyvoice/hifigan/generator.py:672-726),
// CausalConvRNNF0Predictor.forward (cosyvoice/hifigan/f0_predictor.py:95-103),
// SourceModuleHnNSF/SineGen2 (generator.py:233-375) and _stft/_istft (:491-505).
//
// Data layout: every activation is channel-major fp32 [C][L] (time contiguous), one utterance per
// call.  Conv weights are pre-folded (weight-norm) and packed [Cin][K][Cout] by
// flowmirror_hydravox_b200/weights.py so a (ci-chunk, all taps, 64 co) slab is one coalesced read.
//
// Kernels (all fp32 on the CUDA cores: the F0 track and the exp/sin ISTFT head are precision
// critical — generator.py:715 — and the stage is <15 % of the pipeline):
//   conv1d_tile_kernel   smem-tiled causal conv: 64 co x 256 t per CTA, 8x8 register tile per
//                        thread, fused nearest-upsample / dilation / Snake|lrelu pre-activation
//                        on the staged input tile, fused bias + ELU + residual(s) + scaled
//                        accumulate epilogue (the 3-resblock mean never materialises).
//   conv1d_strided_kernel  direct kernel for the three tiny source down-convs (18 -> C, k<=30).
//   f0_head / phase_scan / frame_sin / source_synth / stft16 / istft16 (frame + overlap-add).
#include "common.cuh"
#include "gemm.cuh"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace hvx {

constexpr int CO_TILE = 64;
constexpr int T_TILE = 256;
constexpr int CI_CHUNK = 8;
constexpr int K_MAX = 16;
constexpr int HALO_MAX = 64;          // (K-1)*dil <= 50 on this path
constexpr int XW = T_TILE + HALO_MAX;

struct ConvArgs {
  const float* x; const float* w; const float* bias; float* y;
  const float* res1; const float* res2; const float* alpha;
  int Cin, Cout, K, dil, up, pad_left, Lin, Lout, ldx;
  int pre_act;          // 0 none, 1 leaky-relu(slope), 2 snake(alpha)
  float slope;
  int post_elu /* 1 ELU, 2 tanh */, accumulate, reflect1;
  float out_scale;
  int zstuff;           // 1: the x`up` input is zero-stuffed (x[v/up] at v % up == 0, else 0) instead of nearest-repeated:
                        //    ConvTranspose1d(k, stride up) == conv of the zero-stuffed signal with the tap-reversed kernel
};

__device__ __forceinline__ float pre_activate(float v, int mode, float slope, float a) {
  if (mode == 1) return v > 0.f ? v : v * slope;
  if (mode == 2) { float s = sinf(v * a); return v + (1.0f / (a + 1e-9f)) * (s * s); }
  return v;
}

__global__ void __launch_bounds__(256) conv1d_tile_kernel(ConvArgs p) {
  __shared__ __align__(16) float sw[CI_CHUNK][K_MAX][CO_TILE];
  __shared__ float sx[CI_CHUNK][XW];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int t0 = blockIdx.x * T_TILE;
  const int co0 = blockIdx.y * CO_TILE;
  const int halo = (p.K - 1) * p.dil;
  const int W = T_TILE + halo;
  const int Lup = p.zstuff ? (p.Lin - 1) * p.up + 1 : p.Lin * p.up;
  float acc[8][8];
#pragma unroll
  for (int c = 0; c < 8; c++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[c][j] = 0.f;

  for (int ci0 = 0; ci0 < p.Cin; ci0 += CI_CHUNK) {
    // stage weights [ci][k][co]
    const int nw = CI_CHUNK * p.K * CO_TILE;
    for (int i = tid; i < nw; i += 256) {
      int c = i % CO_TILE, r = i / CO_TILE;
      int k = r % p.K, ci = r / p.K;
      float v = 0.f;
      if (ci0 + ci < p.Cin && co0 + c < p.Cout)
        v = __ldg(p.w + ((size_t)(ci0 + ci) * p.K + k) * p.Cout + co0 + c);
      sw[ci][k][c] = v;
    }
    // stage input tile with fused upsample + pre-activation
    for (int i = tid; i < CI_CHUNK * W; i += 256) {
      int ci = i / W, o = i - ci * W;
      int v = t0 + o - p.pad_left;
      float val = 0.f;
      if (ci0 + ci < p.Cin && v >= 0 && v < Lup && (!p.zstuff || v % p.up == 0)) {
        float a = (p.pre_act == 2) ? __ldg(p.alpha + ci0 + ci) : 0.f;
        val = pre_activate(__ldg(p.x + (size_t)(ci0 + ci) * p.ldx + v / p.up), p.pre_act, p.slope, a);
      }
      sx[ci][o] = val;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < CI_CHUNK; ci++) {
#pragma unroll 1
      for (int k = 0; k < p.K; k++) {
        const float4 wa = *reinterpret_cast<const float4*>(&sw[ci][k][wid * 8]);
        const float4 wb = *reinterpret_cast<const float4*>(&sw[ci][k][wid * 8 + 4]);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        float xv[8];
        const float* xr = &sx[ci][lane + k * p.dil];
#pragma unroll
        for (int j = 0; j < 8; j++) xv[j] = xr[32 * j];
#pragma unroll
        for (int c = 0; c < 8; c++)
#pragma unroll
          for (int j = 0; j < 8; j++) acc[c][j] = fmaf(wv[c], xv[j], acc[c][j]);
      }
    }
    __syncthreads();
  }
  // epilogue
  const int Ly = p.Lout + (p.reflect1 ? 1 : 0);
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const int co = co0 + wid * 8 + c;
    if (co >= p.Cout) continue;
    const float b = p.bias ? __ldg(p.bias + co) : 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int t = t0 + lane + 32 * j;
      if (t >= p.Lout) continue;
      float v = acc[c][j] + b;
      if (p.post_elu == 1) v = v > 0.f ? v : expm1f(v);
      else if (p.post_elu == 2) v = tanhf(v);
      const int nrep = (p.reflect1 && t == 1) ? 2 : 1;
      for (int r = 0; r < nrep; r++) {
        const size_t idx = (size_t)co * Ly + (r == 1 ? 0 : t + (p.reflect1 ? 1 : 0));
        float o = v;
        if (p.res1) o += p.res1[idx];
        if (p.res2) o += p.res2[idx];
        o *= p.out_scale;
        if (p.accumulate) o += p.y[idx];
        p.y[idx] = o;
      }
    }
  }
}

// y[co][t] = b[co] + sum_{ci,k} w[ci][k][co] * x[ci][t*stride + k - pad_left]   (zero outside)
__global__ void conv1d_strided_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ y, int Cin, int Cout,
                                      int K, int stride, int pad_left, int Lin, int Lout) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int co = blockIdx.y;
  if (t >= Lout) return;
  float acc = bias ? bias[co] : 0.f;
  for (int ci = 0; ci < Cin; ci++) {
    const float* xr = x + (size_t)ci * Lin;
    for (int k = 0; k < K; k++) {
      int v = t * stride + k - pad_left;
      if (v >= 0 && v < Lin) acc = fmaf(__ldg(w + ((size_t)ci * K + k) * Cout + co), xr[v], acc);
    }
  }
  y[(size_t)co * Lout + t] = acc;
}

// f0[t] = | b + sum_c w[c] * h[c][t] |      (f0_predictor.py:101-103)
__global__ void f0_head_kernel(const float* __restrict__ h, const float* __restrict__ w, const float* __restrict__ b,
                               float* __restrict__ f0, int C, int L) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= L) return;
  float acc = 0.f;
  for (int c = 0; c < C; c++) acc = fmaf(__ldg(w + c), h[(size_t)c * L + t], acc);
  f0[t] = fabsf(acc + b[0]);
}

// cumulative phase per harmonic at frame rate, accumulated in fp64 like torch's CPU cumsum
// (generator.py:239-255 after the exact x1/480 linear down-sampling).  One block, H*32 threads.
__global__ void phase_scan_kernel(const float* __restrict__ f0, float* __restrict__ cum, int T, int H, float sr) {
  extern __shared__ double part[];               // [H][32]
  const int h = threadIdx.x / 32, c = threadIdx.x % 32;
  const int chunk = (T + 31) / 32;
  const int i0 = c * chunk, i1 = min(T, i0 + chunk);
  const float hm = (float)(h + 1);
  double s = 0.0;
  for (int i = i0; i < i1; i++) s += (double)fmodf((f0[i] * hm) / sr, 1.0f);
  part[h * 32 + c] = s;
  __syncthreads();
  double run = 0.0;
  for (int k = 0; k < c; k++) run += part[h * 32 + k];
  for (int i = i0; i < i1; i++) {
    run += (double)fmodf((f0[i] * hm) / sr, 1.0f);
    cum[(size_t)i * H + h] = (float)run;
  }
}

// per frame/harmonic: 0.1*sin(((c*2)*pi)*frame)  (generator.py:255-258,300)
__global__ void frame_sin_kernel(const float* __restrict__ cum, float* __restrict__ sinamp, int n, float frame) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float ph = ((cum[i] * 2.0f) * 3.14159265358979323846f) * frame;
  sinamp[i] = sinf(ph) * 0.1f;
}

// s[n] = tanh(b + sum_h w[h]*(sinamp[i][h]*uv_i + namp_i*U[n][h]))   (generator.py:303-316,366-368)
__global__ void source_synth_kernel(const float* __restrict__ f0, const float* __restrict__ sinamp,
                                    const float* __restrict__ table, const float* __restrict__ lw,
                                    const float* __restrict__ lb, float* __restrict__ s, int n_samples, int frame,
                                    int H) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_samples) return;
  const int i = n / frame;
  const float uv = f0[i] > 10.0f ? 1.0f : 0.0f;
  const float namp = uv * 0.003f + ((1.0f - uv) * 0.1f) / 3.0f;
  float acc = 0.f;
  for (int h = 0; h < H; h++) {
    float sw = sinamp[(size_t)i * H + h] * uv + namp * __ldg(table + (size_t)n * H + h);
    acc = fmaf(__ldg(lw + h), sw, acc);
  }
  s[n] = tanhf(acc + lb[0]);
}

// ---- non-causal SineGen2 (generator.py:233-317 with causal=False): frame-rate phase, linearly up-sampled
// phase[i][h] = ((cum*2)*pi)*frame   (:257-258; same fp32 op order)
__global__ void frame_phase_kernel(const float* __restrict__ cum, float* __restrict__ phase, int n, float frame) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  phase[i] = ((cum[i] * 2.0f) * 3.14159265358979323846f) * frame;
}

// s[n] = tanh(b + sum_h w[h]*(0.1*sin(lerp(phase))*uv + namp*noise[n][h]))   (:258-260,300-316,366-368)
// lerp = F.interpolate(mode='linear', align_corners=False, scale_factor=frame): src = max(0, (n+0.5)/frame - 0.5)
__global__ void source_synth_nc_kernel(const float* __restrict__ f0, const float* __restrict__ phase,
                                       const float* __restrict__ noise, const float* __restrict__ lw,
                                       const float* __restrict__ lb, float* __restrict__ s, int n_samples, int frame,
                                       int H, int T, float inv_scale) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_samples) return;
  const int i = n / frame;
  const float uv = f0[i] > 10.0f ? 1.0f : 0.0f;
  const float namp = uv * 0.003f + ((1.0f - uv) * 0.1f) / 3.0f;
  float srcf = __fsub_rn(__fmul_rn(inv_scale, (float)n + 0.5f), 0.5f);
  if (srcf < 0.f) srcf = 0.f;
  int i0 = (int)srcf;
  if (i0 > T - 1) i0 = T - 1;
  const int i1 = i0 + (i0 < T - 1 ? 1 : 0);
  float lam = srcf - (float)i0;
  lam = fminf(fmaxf(lam, 0.f), 1.f);
  const float w0 = 1.0f - lam;
  float acc = 0.f;
  for (int h = 0; h < H; h++) {
    const float ph = __fadd_rn(__fmul_rn(w0, phase[(size_t)i0 * H + h]), __fmul_rn(lam, phase[(size_t)i1 * H + h]));
    const float sw = (sinf(ph) * 0.1f) * uv + namp * __ldg(noise + (size_t)n * H + h);
    acc = fmaf(__ldg(lw + h), sw, acc);
  }
  s[n] = tanhf(acc + lb[0]);
}

__constant__ float c_cos16[16];
__constant__ float c_sin16[16];
__constant__ float c_hann16[16];

// torch.stft(s, 16, hop 4, hann(periodic), center=True, reflect) -> out[18][F]  (generator.py:491-497)
__global__ void stft16_kernel(const float* __restrict__ s, float* __restrict__ out, int N, int F) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= F) return;
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; j++) {
    int idx = 4 * m - 8 + j;
    if (idx < 0) idx = -idx;
    if (idx >= N) idx = 2 * (N - 1) - idx;
    v[j] = s[idx] * c_hann16[j];
  }
#pragma unroll
  for (int f = 0; f < 9; f++) {
    float re = 0.f, im = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      re = fmaf(v[j], c_cos16[(f * j) & 15], re);
      im = fmaf(-v[j], c_sin16[(f * j) & 15], im);
    }
    out[(size_t)f * F + m] = re;
    out[(size_t)(9 + f) * F + m] = im;
  }
}

// windowed 16-point inverse real DFT of every frame (generator.py:499-505, 704-707)
__global__ void istft_frames_kernel(const float* __restrict__ x, float* __restrict__ fb, int F) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= F) return;
  float re[9], im[9];
#pragma unroll
  for (int f = 0; f < 9; f++) {
    float mag = fminf(expf(x[(size_t)f * F + m]), 100.0f);
    float ph = sinf(x[(size_t)(9 + f) * F + m]);
    re[f] = mag * cosf(ph);
    im[f] = mag * sinf(ph);
  }
#pragma unroll
  for (int j = 0; j < 16; j++) {
    float acc = re[0] + ((j & 1) ? -re[8] : re[8]);
#pragma unroll
    for (int f = 1; f < 8; f++)
      acc += 2.0f * (re[f] * c_cos16[(f * j) & 15] - im[f] * c_sin16[(f * j) & 15]);
    fb[(size_t)m * 16 + j] = acc * (1.0f / 16.0f) * c_hann16[j];
  }
}

// overlap-add / window-envelope normalisation / centre trim / clamp (torch.istft + generator.py:710)
__global__ void istft_ola_kernel(const float* __restrict__ fb, float* __restrict__ y, int F, int n_out, float limit) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_out) return;
  const int p = n + 8;
  int m_hi = p >> 2;
  if (m_hi > F - 1) m_hi = F - 1;
  int m_lo = (p - 15 + 3) >> 2;
  if (m_lo < 0) m_lo = 0;
  float acc = 0.f, env = 0.f;
  for (int m = m_lo; m <= m_hi; m++) {
    const int j = p - 4 * m;
    acc += fb[(size_t)m * 16 + j];
    env += c_hann16[j] * c_hann16[j];
  }
  float v = acc / env;
  y[n] = fminf(fmaxf(v, -limit), limit);
}

// ------------------------------------------------------------------------------------------
struct ConvW {
  const float* w = nullptr; const float* b = nullptr; int Cin = 0, K = 0, Cout = 0;
  const __half* w16 = nullptr; int Cinp = 0;     // tensor-core operand: [Cout][2*K*Cinp] fp16 = [hi | lo], column = tap*Cinp + ci
};
struct ResBlockW { ConvW c1[4], c2[4]; const float* a1[4]; const float* a2[4]; const float* ia1[4]; const float* ia2[4]; };

struct HiftState {
  ConvW f0c[5]; const float* cls_w; const float* cls_b;
  const float* lin_w; const float* lin_b;
  ConvW conv_pre, ups[4], sdown[4], conv_post;
  ResBlockW srb[4], rb[12];
  DevBuf ws;
  bool tables_ready = false;
  bool tc_ready = false;       // the split-fp16 implicit-GEMM operands (".w16", ".ia*") are registered: decode runs on tcgen05
};

static hvx_status get_conv(hvx_engine* e, const std::string& name, ConvW* c) {
  const Tensor* w = e->find(HVX_STAGE_HIFT, name + ".w");
  const Tensor* b = e->find(HVX_STAGE_HIFT, name + ".b");
  HVX_CHECK(w && b && w->ndim == 3 && w->dtype == HVX_F32, HVX_ERR_STATE, "hift: missing/invalid tensor %s", name.c_str());
  c->w = w->f32(); c->b = b->f32();
  c->Cin = (int)w->shape[0]; c->K = (int)w->shape[1]; c->Cout = (int)w->shape[2];
  c->w16 = nullptr; c->Cinp = 0;
  const Tensor* h = e->find(HVX_STAGE_HIFT, name + ".w16");
  if (h && h->dtype == HVX_F16 && h->ndim == 2 && h->shape[0] == c->Cout && h->shape[1] % (2 * c->K * 64) == 0 &&
      h->shape[1] / (2 * c->K) >= c->Cin) {
    c->w16 = reinterpret_cast<const __half*>(h->p);
    c->Cinp = (int)(h->shape[1] / (2 * c->K));
  }
  return HVX_OK;
}

static hvx_status get_vec(hvx_engine* e, const std::string& name, const float** p) {
  const Tensor* t = e->find(HVX_STAGE_HIFT, name);
  HVX_CHECK(t && t->dtype == HVX_F32, HVX_ERR_STATE, "hift: missing tensor %s", name.c_str());
  *p = t->f32();
  return HVX_OK;
}

static hvx_status get_rb(hvx_engine* e, const std::string& pfx, int ndil, ResBlockW* r) {
  for (int j = 0; j < ndil; j++) {
    hvx_status s;
    if ((s = get_conv(e, pfx + ".c1." + std::to_string(j), &r->c1[j]))) return s;
    if ((s = get_conv(e, pfx + ".c2." + std::to_string(j), &r->c2[j]))) return s;
    if ((s = get_vec(e, pfx + ".a1." + std::to_string(j), &r->a1[j]))) return s;
    if ((s = get_vec(e, pfx + ".a2." + std::to_string(j), &r->a2[j]))) return s;
    const Tensor* i1 = e->find(HVX_STAGE_HIFT, pfx + ".ia1." + std::to_string(j));
    const Tensor* i2 = e->find(HVX_STAGE_HIFT, pfx + ".ia2." + std::to_string(j));
    r->ia1[j] = i1 ? i1->f32() : nullptr;
    r->ia2[j] = i2 ? i2->f32() : nullptr;
  }
  return HVX_OK;
}

static bool rb_tc(const ResBlockW& r, int ndil) {
  for (int j = 0; j < ndil; j++)
    if (!r.c1[j].w16 || !r.c2[j].w16 || !r.ia1[j] || !r.ia2[j]) return false;
  return true;
}

hvx_status hift_finalize(hvx_engine* e) {
  const hvx_config& c = e->cfg;
  HVX_CHECK(c.hift_n_fft == 16 && c.hift_hop == 4, HVX_ERR_UNSUPPORTED, "hift: only n_fft=16/hop=4 ISTFT head is built");
  HVX_CHECK(c.hift_n_ups <= 4 && c.hift_n_rb <= 4 && c.hift_n_dil <= 4, HVX_ERR_UNSUPPORTED, "hift: too many stages");
  if (!e->hift) e->hift = new HiftState();
  HiftState* h = e->hift;
  hvx_status s;
  for (int i = 0; i < 5; i++)
    if ((s = get_conv(e, "f0.c" + std::to_string(i), &h->f0c[i]))) return s;
  if ((s = get_vec(e, "f0.cls.w", &h->cls_w))) return s;
  if ((s = get_vec(e, "f0.cls.b", &h->cls_b))) return s;
  if ((s = get_vec(e, "src.lin.w", &h->lin_w))) return s;
  if ((s = get_vec(e, "src.lin.b", &h->lin_b))) return s;
  if ((s = get_conv(e, "conv_pre", &h->conv_pre))) return s;
  if ((s = get_conv(e, "conv_post", &h->conv_post))) return s;
  for (int i = 0; i < c.hift_n_ups; i++) {
    if ((s = get_conv(e, "ups." + std::to_string(i), &h->ups[i]))) return s;
    if ((s = get_conv(e, "sdown." + std::to_string(i), &h->sdown[i]))) return s;
    if ((s = get_rb(e, "srb." + std::to_string(i), c.hift_n_dil, &h->srb[i]))) return s;
    for (int j = 0; j < c.hift_n_rb; j++)
      if ((s = get_rb(e, "rb." + std::to_string(i * c.hift_n_rb + j), c.hift_n_dil, &h->rb[i * c.hift_n_rb + j]))) return s;
  }
  h->tc_ready = h->conv_pre.w16 && h->conv_post.w16;
  for (int i = 0; i < c.hift_n_ups && h->tc_ready; i++) {
    h->tc_ready = h->ups[i].w16 && rb_tc(h->srb[i], c.hift_n_dil);
    for (int j = 0; j < c.hift_n_rb && h->tc_ready; j++) h->tc_ready = rb_tc(h->rb[i * c.hift_n_rb + j], c.hift_n_dil);
  }
  if (!h->tables_ready) {
    float cs[16], sn[16], hn[16];
    for (int j = 0; j < 16; j++) {
      cs[j] = (float)cos(2.0 * M_PI * j / 16.0);
      sn[j] = (float)sin(2.0 * M_PI * j / 16.0);
      hn[j] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * j / 16.0));
    }
    HVX_CUDA(cudaMemcpyToSymbol(c_cos16, cs, sizeof(cs)));
    HVX_CUDA(cudaMemcpyToSymbol(c_sin16, sn, sizeof(sn)));
    HVX_CUDA(cudaMemcpyToSymbol(c_hann16, hn, sizeof(hn)));
    h->tables_ready = true;
  }
  return HVX_OK;
}

void hift_free(hvx_engine* e) { delete e->hift; e->hift = nullptr; }

struct ConvOpt {
  int dil = 1, up = 1, pad_left = -1 /* -1: (K-1)*dil */, pre_act = 0, post_elu = 0, accumulate = 0, reflect1 = 0, zstuff = 0;
  float slope = 0.f, out_scale = 1.f;
  const float* alpha = nullptr; const float* res1 = nullptr; const float* res2 = nullptr;
  int Lout = -1, ldx = -1;
};

static hvx_status run_conv(hvx_engine* e, cudaStream_t st, const ConvW& c, const float* x, int Lin, float* y,
                           const ConvOpt& o) {
  HVX_CHECK(c.K <= K_MAX && (c.K - 1) * o.dil <= HALO_MAX, HVX_ERR_UNSUPPORTED, "conv tile: K=%d dil=%d unsupported", c.K, o.dil);
  ConvArgs a;
  a.x = x; a.w = c.w; a.bias = c.b; a.y = y; a.res1 = o.res1; a.res2 = o.res2; a.alpha = o.alpha;
  a.Cin = c.Cin; a.Cout = c.Cout; a.K = c.K; a.dil = o.dil; a.up = o.up;
  a.pad_left = o.pad_left < 0 ? (c.K - 1) * o.dil : o.pad_left;
  a.Lin = Lin; a.Lout = o.Lout < 0 ? Lin * o.up : o.Lout; a.ldx = o.ldx < 0 ? Lin : o.ldx;
  a.pre_act = o.pre_act; a.slope = o.slope; a.post_elu = o.post_elu; a.accumulate = o.accumulate;
  a.reflect1 = o.reflect1; a.out_scale = o.out_scale; a.zstuff = o.zstuff;
  dim3 grid(cdiv(a.Lout, T_TILE), cdiv(c.Cout, CO_TILE));
  ProfScope prof_scope(&e->prof, st, PROF_HIFT_CONV, 2.0 * c.Cin * c.K * c.Cout * (double)a.Lout);
  conv1d_tile_kernel<<<grid, 256, 0, st>>>(a);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

// x_out = resblock(x_in); the last conv's epilogue optionally adds `extra` and/or accumulates
// out_scale*(result) into `acc_out` instead of writing x_out.
// symmetric: "same" padding (k*d-d)/2 on both sides (ResBlock of the non-causal HiFTGenerator, generator.py:46-108) instead of causal
static hvx_status run_resblock(hvx_engine* e, cudaStream_t st, const ResBlockW& r, int ndil, const int* dils,
                               const float* x_in, float* work, float* tmp, int L, const float* extra, float* final_out,
                               int final_accumulate, float final_scale, bool symmetric = false, bool lrelu = false) {
  const float* cur = x_in;
  for (int j = 0; j < ndil; j++) {
    ConvOpt o1; o1.dil = dils[j]; o1.pre_act = 2; o1.alpha = r.a1[j];
    if (lrelu) { o1.pre_act = 1; o1.slope = 0.1f; o1.alpha = nullptr; }   // ResBlock1 of the classic HiFi-GAN (matcha/hifigan/models.py:86-93)
    if (symmetric) o1.pad_left = (r.c1[j].K - 1) * dils[j] / 2;
    hvx_status s = run_conv(e, st, r.c1[j], cur, L, tmp, o1);
    if (s) return s;
    ConvOpt o2; o2.pre_act = 2; o2.alpha = r.a2[j]; o2.res1 = cur;
    if (lrelu) { o2.pre_act = 1; o2.slope = 0.1f; o2.alpha = nullptr; }
    if (symmetric) o2.pad_left = (r.c2[j].K - 1) / 2;
    float* dst = work;
    if (j == ndil - 1) {
      o2.res2 = extra; o2.accumulate = final_accumulate; o2.out_scale = final_scale;
      dst = final_out;
    }
    s = run_conv(e, st, r.c2[j], tmp, L, dst, o2);
    if (s) return s;
    cur = work;
  }
  return HVX_OK;
}


// ------------------------------------------------------------------------------------------ tensor-core decode path
// The decode stack (conv_pre, up-sampling convs, source / main ResBlocks, conv_post: 99 % of the stage's FLOPs) as implicit
// GEMMs on tcgen05 (gemm.cu): activations frame-major, fp32 residual streams [L][C] plus split-fp16 operand copies
// [L][2*Cp] = [hi | lo] of the ACTIVATED values (Snake / leaky-relu applied once, in the producing epilogue), weights
// [Cout][2*K*Cp] = [hi | lo]; every conv = three tensor-core products A_hi W_hi + A_lo W_hi + A_hi W_lo accumulated in fp32
// in TMEM (~22 mantissa bits per operand), taps addressed through TMA row offsets (zero fill = the causal padding).

// mel (C, ldx) channel-major fp32 -> A16 [L][2*Cp] frame-major split fp16, channels >= C zero
__global__ void hift_mel_pack_kernel(const float* __restrict__ mel, __half* __restrict__ a16, int C, int Cp, int L, int ldx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L * Cp) return;
  const int t = i / Cp, ch = i - t * Cp;
  const float v = ch < C ? mel[(size_t)ch * ldx + t] : 0.f;
  const __half hi = __float2half_rn(v);
  a16[(size_t)t * 2 * Cp + ch] = hi;
  a16[(size_t)t * 2 * Cp + Cp + ch] = __float2half_rn(v - __half2float(hi));
}

// nearest x`up` + leaky-relu: x32 [L][C] -> A16 [L*up][2*Cp]   (CausalConv1dUpsample's Upsample, convolution.py:246-257)
__global__ void hift_act_up_kernel(const float* __restrict__ x, __half* __restrict__ a16, int C, int Cp, int L, int up, float slope) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)L * up * Cp) return;
  const int ch = (int)(i % Cp);
  const size_t t = i / Cp;
  float v = 0.f;
  if (ch < C) { v = x[(t / up) * C + ch]; v = v > 0.f ? v : v * slope; }
  const __half hi = __float2half_rn(v);
  a16[t * 2 * Cp + ch] = hi;
  a16[t * 2 * Cp + Cp + ch] = __float2half_rn(v - __half2float(hi));
}

// source down-conv (18 -> C, strided) from the channel-major source STFT: si32 [Lout][C] frame-major and its Snake copy
__global__ void hift_sdown_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                  float* __restrict__ y32, __half* __restrict__ a16, const float* __restrict__ alpha,
                                  const float* __restrict__ inv_alpha, int Cin, int Cout, int Cp, int K, int stride, int pad_left,
                                  int Lin, int Lout) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)Lout * Cout) return;
  const int co = (int)(i % Cout);
  const int t = (int)(i / Cout);
  float acc = bias[co];
  for (int ci = 0; ci < Cin; ci++) {
    const float* xr = x + (size_t)ci * Lin;
    for (int k = 0; k < K; k++) {
      const int v = t * stride + k - pad_left;
      if (v >= 0 && v < Lin) acc = fmaf(__ldg(w + ((size_t)ci * K + k) * Cout + co), xr[v], acc);
    }
  }
  y32[i] = acc;
  const float sn = sinf(acc * alpha[co]);
  const float a = acc + inv_alpha[co] * (sn * sn);
  const __half hi = __float2half_rn(a);
  a16[(size_t)t * 2 * Cp + co] = hi;
  a16[(size_t)t * 2 * Cp + Cp + co] = __float2half_rn(a - __half2float(hi));
}

// windowed 16-point inverse real DFT of every frame, conv_post output frame-major [F][ld] (generator.py:499-505, 704-707)
__global__ void istft_frames_fm_kernel(const float* __restrict__ x, float* __restrict__ fb, int F, int ld) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= F) return;
  float re[9], im[9];
#pragma unroll
  for (int f = 0; f < 9; f++) {
    const float mag = fminf(expf(x[(size_t)m * ld + f]), 100.0f);
    const float ph = sinf(x[(size_t)m * ld + 9 + f]);
    re[f] = mag * cosf(ph);
    im[f] = mag * sinf(ph);
  }
#pragma unroll
  for (int j = 0; j < 16; j++) {
    float acc = re[0] + ((j & 1) ? -re[8] : re[8]);
#pragma unroll
    for (int f = 1; f < 8; f++)
      acc += 2.0f * (re[f] * c_cos16[(f * j) & 15] - im[f] * c_sin16[(f * j) & 15]);
    fb[(size_t)m * 16 + j] = acc * (1.0f / 16.0f) * c_hann16[j];
  }
}

// one convolution as a three-term split-fp16 implicit GEMM: out rows = time, N = Cout, K = taps * Cp
static hvx_status tc_conv(hvx_engine* e, cudaStream_t st, const ConvW& w, const __half* A, int a_rows, int Lout, int dil, int pad_left,
                          const HiftEpi& he) {
  GemmEpi p; p.mode = EPI_HIFT; p.f16 = 1; p.bias = w.b; p.hift = he;
  GemmAddr ga; ga.rows_per_batch = Lout; ga.a_rows = a_rows; ga.a_cols = 2 * w.Cinp; ga.kb_per_tap = w.Cinp / 64; ga.a_row_step = dil;
  ga.a_row0 = -pad_left; ga.split3_kb = w.K * w.Cinp / 64; ga.a_lo_off = w.Cinp;
  ProfScope ps(&e->prof, st, PROF_HIFT_CONV, 2.0 * w.Cin * w.K * w.Cout * (double)Lout);
  e->prof_gemm_off++;
  const hvx_status rc = gemm_bf16(e, st, (const __nv_bfloat16*)A, 2 * w.Cinp, (const __nv_bfloat16*)w.w16, 2 * w.K * w.Cinp, Lout, w.Cout,
                                  3 * w.K * w.Cinp, p, &ga);
  e->prof_gemm_off--;
  return rc;
}

struct TcBufs {            // per-stage carve-up (elements): fp32 [L][C], fp16 [L][2*Cp]
  float *x_up, *si, *x, *work, *xs;
  __half *a_in, *a_cur, *a_tmp, *a_r[3];
};

// one ResBlock on tensor cores.  a_first = Snake_{a1[0]}(x_in) (split fp16, produced by the epilogue that wrote x_in).
// Last conv: v = c2(...) + cur (+ extra); out32 = final32 (scaled / accumulated); fin16[k] = activated copies of the stored value.
static hvx_status tc_resblock(hvx_engine* e, cudaStream_t st, const ResBlockW& r, int ndil, const int* dils, const float* x_in,
                              const __half* a_first, __half* a_cur, __half* a_tmp, float* work, int L, int C, int Cp,
                              const float* extra, float* final32, int final_acc, float final_scale, int n_fin16, __half* const* fin16,
                              const int* fin_act, const float* const* fin_alpha, const float* const* fin_ialpha, float fin_slope) {
  const float* cur = x_in;
  const __half* a = a_first;
  for (int j = 0; j < ndil; j++) {
    HiftEpi e1; e1.n16 = 1; e1.out16[0] = (uint16_t*)a_tmp; e1.act[0] = 2; e1.alpha[0] = r.a2[j]; e1.inv_alpha[0] = r.ia2[j];
    e1.ld16 = 2 * Cp; e1.lo_off = Cp;
    hvx_status s = tc_conv(e, st, r.c1[j], a, L, L, dils[j], (r.c1[j].K - 1) * dils[j], e1);
    if (s) return s;
    HiftEpi e2; e2.resid = cur; e2.ldr = C; e2.ld16 = 2 * Cp; e2.lo_off = Cp;
    if (j < ndil - 1) {
      e2.out32 = work; e2.ld32 = C;
      e2.n16 = 1; e2.out16[0] = (uint16_t*)a_cur; e2.act[0] = 2; e2.alpha[0] = r.a1[j + 1]; e2.inv_alpha[0] = r.ia1[j + 1];
    } else {
      e2.resid2 = extra; e2.out32 = final32; e2.ld32 = C; e2.accumulate = final_acc; e2.out_scale = final_scale;
      e2.n16 = n_fin16; e2.slope = fin_slope;
      for (int k = 0; k < n_fin16; k++) { e2.out16[k] = (uint16_t*)fin16[k]; e2.act[k] = fin_act[k]; e2.alpha[k] = fin_alpha[k]; e2.inv_alpha[k] = fin_ialpha[k]; }
    }
    s = tc_conv(e, st, r.c2[j], a_tmp, L, L, 1, r.c2[j].K - 1, e2);
    if (s) return s;
    cur = work; a = a_cur;
  }
  return HVX_OK;
}

// CausalHiFTGenerator.decode (generator.py:672-711) on tensor cores.  sstft: source STFT channel-major [18][Fs] (fp32 kernels).
static hvx_status hift_decode_tc(hvx_engine* e, cudaStream_t st, const float* mel, int T, int Lmel, int Tx, const float* sstft, int Fs,
                                 float* ws, float* post, int* post_ld) {
  const hvx_config& c = e->cfg;
  HiftState* h = e->hift;
  const int nu = c.hift_n_ups, nd = c.hift_n_dil, nr = c.hift_n_rb;
  hvx_status rc;
  // workspace carve-up: every region sized for the largest stage
  size_t max32 = (size_t)c.hift_base * Tx, max16 = (size_t)2 * h->conv_pre.Cinp * T;
  { int L = Tx;
    for (int i = 0; i < nu; i++) {
      const int Cin = c.hift_base >> i, C = c.hift_base >> (i + 1);
      const int Cp = (C + 63) / 64 * 64, Cinp = (Cin + 63) / 64 * 64;
      const size_t Lo = (size_t)L * c.hift_ups[i] + 1;
      max32 = std::max(max32, Lo * C);
      max16 = std::max(max16, std::max(Lo * 2 * Cp, (size_t)L * c.hift_ups[i] * 2 * Cinp));
      L = (int)Lo - (i == nu - 1 ? 0 : 1);
    } }
  max32 = (max32 + 63) & ~(size_t)63; max16 = (max16 + 127) & ~(size_t)127;
  float* f32 = ws;
  __half* f16 = reinterpret_cast<__half*>(ws + 5 * max32);
  TcBufs b;
  b.x_up = f32; b.si = f32 + max32; b.x = f32 + 2 * max32; b.work = f32 + 3 * max32; b.xs = f32 + 4 * max32;
  b.a_in = f16; b.a_cur = f16 + max16; b.a_tmp = f16 + 2 * max16; b.a_r[0] = f16 + 3 * max16; b.a_r[1] = f16 + 4 * max16; b.a_r[2] = f16 + 5 * max16;

  // conv_pre: k5 looking right (pad 0 left, zero fill past the last input frame), mel -> x0 [Tx][base]
  { const ConvW& w = h->conv_pre;
    hift_mel_pack_kernel<<<cdiv(Lmel * w.Cinp, 256), 256, 0, st>>>(mel, b.a_in, c.hift_mel, w.Cinp, Lmel, T);
    HVX_LAUNCH_CHECK(e);
    HiftEpi he; he.out32 = b.xs; he.ld32 = w.Cout;
    if ((rc = tc_conv(e, st, w, b.a_in, Lmel, Tx, 1, 0, he))) return rc; }
  const float* cur = b.xs;       // [L][Cin] fp32
  int L = Tx, down = 1;
  for (int i = 0; i < nu; i++) down *= c.hift_ups[i];
  for (int i = 0; i < nu; i++) {
    const int u = c.hift_ups[i], last = (i == nu - 1);
    const int Cin = c.hift_base >> i, C = c.hift_base >> (i + 1);
    const int Cp = (C + 63) / 64 * 64, Cinp = (Cin + 63) / 64 * 64;
    const int Lu = L * u, Lo = Lu + (last ? 1 : 0);
    HVX_CHECK(h->ups[i].Cinp == Cinp && h->ups[i].Cout == C, HVX_ERR_STATE, "hift: ups.%d operand shape mismatch", i);
    if (Cp != C) HVX_CUDA(cudaMemsetAsync(b.a_cur, 0, sizeof(__half) * 5 * max16, st));     // padded operand columns must be zero
    // up-sampling conv: leaky-relu(0.1) + nearest x u, causal k (generator.py:681-687)
    hift_act_up_kernel<<<(unsigned)cdiv((long long)Lu * Cinp, 256), 256, 0, st>>>(cur, b.a_in, Cin, Cinp, L, u, 0.1f);
    HVX_LAUNCH_CHECK(e);
    { HiftEpi he; he.out32 = b.x_up; he.ld32 = C; he.row_shift = last ? 1 : 0; he.dup_row1 = last;
      if ((rc = tc_conv(e, st, h->ups[i], b.a_in, Lu, Lu, 1, h->ups[i].K - 1, he))) return rc; }
    // a_in is re-used below as conv_post's operand [Lo][2*Cp]: epilogues write only the C real columns, so with padded channels
    // the pad columns (and the extra reflected row) must not keep stale bit patterns (0 * NaN)
    if (Cp != C && last) HVX_CUDA(cudaMemsetAsync(b.a_in, 0, sizeof(__half) * max16, st));
    // source branch: strided down-conv of the source STFT, then its ResBlock, "+ x_up" fused into its last epilogue, which also
    // writes Snake_{a1[0]}(x) for each of the parallel ResBlocks
    down /= u;
    const ResBlockW& srb = h->srb[i];
    { const ConvW& d = h->sdown[i];
      hift_sdown_kernel<<<(unsigned)cdiv((long long)Lo * C, 256), 256, 0, st>>>(sstft, d.w, d.b, b.si, b.a_cur, srb.a1[0], srb.ia1[0], d.Cin, C, Cp, d.K,
                                                                              down, down > 1 ? down - 1 : 0, Fs, Lo);
      HVX_LAUNCH_CHECK(e); }
    { __half* fin16[3]; int fin_act[3]; const float* fa[3]; const float* fi[3];
      for (int j = 0; j < nr; j++) { fin16[j] = b.a_r[j]; fin_act[j] = 2; fa[j] = h->rb[i * nr + j].a1[0]; fi[j] = h->rb[i * nr + j].ia1[0]; }
      if ((rc = tc_resblock(e, st, srb, nd, c.hift_rb_d, b.si, b.a_cur, b.a_cur, b.a_tmp, b.si, Lo, C, Cp, b.x_up, b.x, 0, 1.0f, nr, fin16, fin_act, fa, fi, 0.f)))
        return rc; }
    // mean of the parallel ResBlocks, accumulated in the last epilogues; the very last one also emits leaky-relu(0.01)(xs) for conv_post
    for (int j = 0; j < nr; j++) {
      const bool emit = last && j == nr - 1;
      __half* fin16[1] = {b.a_in}; int fin_act[1] = {1}; const float* fa[1] = {nullptr}; const float* fi[1] = {nullptr};
      if ((rc = tc_resblock(e, st, h->rb[i * nr + j], nd, c.hift_rb_d, b.x, b.a_r[j], b.a_r[j], b.a_tmp, b.work, Lo, C, Cp, nullptr, b.xs, j > 0,
                            1.0f / nr, emit ? 1 : 0, fin16, fin_act, fa, fi, 0.01f))) return rc;
    }
    cur = b.xs; L = Lo;
  }
  { const ConvW& w = h->conv_post;                       // leaky-relu(0.01) (applied above) + causal k7 -> [F][Cout]
    HiftEpi he; he.out32 = post; he.ld32 = w.Cout;
    if ((rc = tc_conv(e, st, w, b.a_in, L, L, 1, w.K - 1, he))) return rc;
    *post_ld = w.Cout; }
  return HVX_OK;
}

}  // namespace hvx

using namespace hvx;

extern "C" hvx_status hvx_hift_vocode(hvx_engine* e, const float* mel, int T, int finalize, const float* table,
                                      int64_t n_table_rows, const float* f0_in, float* f0_out, float* wav, float* src, void* stream) {
  HVX_CHECK(e && e->hift, HVX_ERR_STATE, "hift stage not finalized");
  HVX_LOCK(e, HVX_STAGE_HIFT);
  HVX_CHECK(mel && table && wav, HVX_ERR_ARG, "hift: null argument");
  const hvx_config& c = e->cfg;
  HiftState* h = e->hift;
  cudaStream_t st = (cudaStream_t)stream;
  int frame = c.hift_hop, up_prod = 1;
  for (int i = 0; i < c.hift_n_ups; i++) { frame *= c.hift_ups[i]; up_prod *= c.hift_ups[i]; }
  const int H = c.hift_harmonics;
  const int Tf0 = finalize ? T : T - 3;            // f0_predictor.py:96-99
  const int Tx = finalize ? T : T - 7;             // generator.py:676-679,725
  HVX_CHECK(Tx >= 1 && (finalize || Tx >= 2), HVX_ERR_ARG, "hift: T=%d too short", T);
  const int Ns = Tf0 * frame;                      // source samples
  HVX_CHECK((int64_t)Ns <= n_table_rows, HVX_ERR_ARG, "hift: the utterance needs %d sine-table rows, the table holds %lld (SineGen2.sine_waves, generator.py:226,306)",
            Ns, (long long)n_table_rows);
  const int Fs = Ns / 4 + 1;                       // stft frames of the source
  const int F = Tx * up_prod + 1;                  // frames entering the ISTFT
  const int n_out = finalize ? Tx * frame : (Tx - 1) * frame;
  const int C0 = c.hift_base, CF = c.hift_f0_ch;
  // decode stack on tensor cores (split-fp16 implicit GEMMs) unless the operands were not packed or HVX_HIFT_FP32=1 asks for the
  // fp32 CUDA-core convolutions (A/B runs)
  const char* env_fp32 = getenv("HVX_HIFT_FP32");
  const bool force_fp32 = env_fp32 && atoi(env_fp32);
  const bool tc = h->tc_ready && !force_fp32;

  // workspace carve-up (floats)
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 63) & ~(size_t)63; return o; };
  const size_t o_fa = take((size_t)CF * Tf0), o_fb = take((size_t)CF * Tf0), o_f0 = take(Tf0);
  const size_t o_cum = take((size_t)Tf0 * H), o_sin = take((size_t)Tf0 * H), o_s = take(Ns);
  const size_t o_stft = take((size_t)18 * Fs);
  const size_t o_post = take((size_t)std::max(18, h->conv_post.Cout) * F), o_fbuf = take((size_t)16 * F);
  size_t o_x0 = 0, o_x = 0, o_xs = 0, o_si = 0, o_work = 0, o_tmp = 0, o_tc = 0;
  if (tc) {
    size_t max32 = (size_t)C0 * Tx, max16 = (size_t)2 * h->conv_pre.Cinp * T;
    int L = Tx;
    for (int i = 0; i < c.hift_n_ups; i++) {
      const int Cin = C0 >> i, C = C0 >> (i + 1);
      const size_t Cp = (C + 63) / 64 * 64, Cinp = (Cin + 63) / 64 * 64;
      const size_t Lo = (size_t)L * c.hift_ups[i] + 1;
      max32 = std::max(max32, Lo * C);
      max16 = std::max(max16, std::max(Lo * 2 * Cp, (size_t)L * c.hift_ups[i] * 2 * Cinp));
      L = (int)Lo - (i == c.hift_n_ups - 1 ? 0 : 1);
    }
    max32 = (max32 + 63) & ~(size_t)63; max16 = (max16 + 127) & ~(size_t)127;
    o_tc = take(5 * max32 + 3 * max16);
  } else {
    size_t maxCL = 0;
    { int L = Tx; for (int i = 0; i < c.hift_n_ups; i++) { L = L * c.hift_ups[i]; size_t cl = (size_t)(C0 >> (i + 1)) * (L + 1); if (cl > maxCL) maxCL = cl; } }
    o_x0 = take((size_t)C0 * Tx);
    o_x = take(maxCL); o_xs = take(maxCL); o_si = take(maxCL); o_work = take(maxCL); o_tmp = take(maxCL);
  }
  float* ws = (float*)h->ws.get(off * sizeof(float));
  HVX_CHECK(ws, HVX_ERR_CUDA, "hift: workspace allocation of %zu bytes failed", off * sizeof(float));
  float *fa = ws + o_fa, *fb = ws + o_fb, *f0 = ws + o_f0, *cum = ws + o_cum, *sinamp = ws + o_sin, *s = ws + o_s;
  float *sstft = ws + o_stft, *post = ws + o_post, *fbuf = ws + o_fbuf;
  hvx_status rc;

  // ---- F0 predictor (fp32; reference runs it on the CPU for precision, generator.py:715-717)
  if (f0_in) {
    HVX_CUDA(cudaMemcpyAsync(f0, f0_in, sizeof(float) * Tf0, cudaMemcpyDeviceToDevice, st));
  } else {
    ConvOpt o; o.pad_left = 0; o.post_elu = 1; o.Lout = Tf0;      // right-causal k4 (look-ahead 3)
    if ((rc = run_conv(e, st, h->f0c[0], mel, T, fa, o))) return rc;
    float *a = fa, *b = fb;
    for (int i = 1; i < 5; i++) {
      ConvOpt oi; oi.post_elu = 1;
      if ((rc = run_conv(e, st, h->f0c[i], a, Tf0, b, oi))) return rc;
      float* t2 = a; a = b; b = t2;
    }
    f0_head_kernel<<<cdiv(Tf0, 128), 128, 0, st>>>(a, h->cls_w, h->cls_b, f0, CF, Tf0);
    HVX_LAUNCH_CHECK(e);
  }
  if (f0_out) HVX_CUDA(cudaMemcpyAsync(f0_out, f0, sizeof(float) * Tf0, cudaMemcpyDeviceToDevice, st));

  // ---- harmonic source
  phase_scan_kernel<<<1, H * 32, H * 32 * sizeof(double), st>>>(f0, cum, Tf0, H, (float)c.hift_sr);
  HVX_LAUNCH_CHECK(e);
  frame_sin_kernel<<<cdiv(Tf0 * H, 256), 256, 0, st>>>(cum, sinamp, Tf0 * H, (float)frame);
  HVX_LAUNCH_CHECK(e);
  source_synth_kernel<<<cdiv(Ns, 256), 256, 0, st>>>(f0, sinamp, table, h->lin_w, h->lin_b, s, Ns, frame, H);
  HVX_LAUNCH_CHECK(e);
  if (src) HVX_CUDA(cudaMemcpyAsync(src, s, sizeof(float) * Ns, cudaMemcpyDeviceToDevice, st));
  stft16_kernel<<<cdiv(Fs, 128), 128, 0, st>>>(s, sstft, Ns, Fs);
  HVX_LAUNCH_CHECK(e);

  // ---- decode
  if (tc) {
    int post_ld = 0;
    if ((rc = hift_decode_tc(e, st, mel, T, finalize ? T : T - 3, Tx, sstft, Fs, ws + o_tc, post, &post_ld))) return rc;
    istft_frames_fm_kernel<<<cdiv(F, 128), 128, 0, st>>>(post, fbuf, F, post_ld);
    HVX_LAUNCH_CHECK(e);
    istft_ola_kernel<<<cdiv(n_out, 256), 256, 0, st>>>(fbuf, wav, F, n_out, 0.99f);
    HVX_LAUNCH_CHECK(e);
    return HVX_OK;
  }
  float *x0 = ws + o_x0, *x = ws + o_x, *xs = ws + o_xs, *si = ws + o_si, *work = ws + o_work, *tmp = ws + o_tmp;
  { ConvOpt o; o.pad_left = 0; o.Lout = Tx; o.ldx = T;           // conv_pre k5, look-right 4
    if ((rc = run_conv(e, st, h->conv_pre, mel, finalize ? T : T - 3, x0, o))) return rc; }
  const float* cur = x0;
  int L = Tx;
  int down = up_prod;
  for (int i = 0; i < c.hift_n_ups; i++) {
    const int u = c.hift_ups[i];
    const int last = (i == c.hift_n_ups - 1);
    const int Lo = L * u + (last ? 1 : 0);
    { ConvOpt o; o.pre_act = 1; o.slope = 0.1f; o.up = u; o.reflect1 = last;
      if ((rc = run_conv(e, st, h->ups[i], cur, L, x, o))) return rc; }
    // source branch: strided causal down-conv of the source STFT, then its resblock, fused "+ x"
    down /= u;                                                     // 15, 3, 1 at full dims
    { const ConvW& d = h->sdown[i];
      const int stride = down, pad = down > 1 ? down - 1 : 0;
      dim3 grid(cdiv(Lo, 128), d.Cout);
      conv1d_strided_kernel<<<grid, 128, 0, st>>>(sstft, d.w, d.b, si, d.Cin, d.Cout, d.K, stride, pad, Fs, Lo);
      HVX_LAUNCH_CHECK(e); }
    if ((rc = run_resblock(e, st, h->srb[i], c.hift_n_dil, c.hift_rb_d, si, work, tmp, Lo, x, x, 0, 1.0f))) return rc;
    // mean of the parallel resblocks, accumulated in the final conv epilogues
    for (int j = 0; j < c.hift_n_rb; j++)
      if ((rc = run_resblock(e, st, h->rb[i * c.hift_n_rb + j], c.hift_n_dil, c.hift_rb_d, x, work, tmp, Lo, nullptr, xs,
                             j > 0, 1.0f / c.hift_n_rb))) return rc;
    cur = xs; L = Lo;     // next stage reads xs and writes x (never in place: shapes differ)
  }
  { ConvOpt o; o.pre_act = 1; o.slope = 0.01f;
    if ((rc = run_conv(e, st, h->conv_post, cur, L, post, o))) return rc; }
  istft_frames_kernel<<<cdiv(F, 128), 128, 0, st>>>(post, fbuf, F);
  HVX_LAUNCH_CHECK(e);
  istft_ola_kernel<<<cdiv(n_out, 256), 256, 0, st>>>(fbuf, wav, F, n_out, 0.99f);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

// Non-causal HiFTGenerator.inference (cosyvoice/hifigan/generator.py:557-569; decode :506-540; ConvRNNF0Predictor
// f0_predictor.py:9-55; SineGen2 with causal=False :233-317).  Same packed tensors as the causal vocoder
// (weights.pack_hift_t stores the ConvTranspose1d kernels tap-reversed), symmetric "same" padding everywhere, transposed
// convolutions as zero-stuffed convolutions.  noise_dev (frame*T, harmonics): the randn draw of :310, explicit;
// cache_source_dev (n_cache samples) overwrites the head of the source (:566-567); n_cache == frame*T needs no noise.
extern "C" hvx_status hvx_hift_t_vocode(hvx_engine* e, const float* mel, int T, const float* noise, const float* cache_source,
                                        int n_cache, const float* f0_in, float* f0_out, float* wav, float* src, void* stream) {
  HVX_CHECK(e && e->hift, HVX_ERR_STATE, "hift stage not finalized");
  HVX_LOCK(e, HVX_STAGE_HIFT);
  const hvx_config& c = e->cfg;
  HiftState* h = e->hift;
  cudaStream_t st = (cudaStream_t)stream;
  int frame = c.hift_hop, up_prod = 1;
  for (int i = 0; i < c.hift_n_ups; i++) { frame *= c.hift_ups[i]; up_prod *= c.hift_ups[i]; }
  const int H = c.hift_harmonics;
  HVX_CHECK(T >= 1 && mel && wav, HVX_ERR_ARG, "hift_t: bad arguments (T=%d)", T);
  const int Ns = T * frame, Fs = Ns / 4 + 1, F = T * up_prod + 1;
  HVX_CHECK(n_cache >= 0 && n_cache <= Ns && (n_cache == 0 || cache_source), HVX_ERR_ARG, "hift_t: cache_source length %d outside [0,%d]", n_cache, Ns);
  HVX_CHECK(noise || n_cache == Ns, HVX_ERR_ARG, "hift_t: the source needs its noise draw (or a cache_source covering it)");
  for (int i = 0; i < c.hift_n_ups; i++)
    HVX_CHECK(h->ups[i].K >= c.hift_ups[i] && (h->ups[i].K - c.hift_ups[i]) % 2 == 0, HVX_ERR_UNSUPPORTED,
              "hift_t: upsample kernel %d / rate %d: only k-u even is built (output length u*L)", h->ups[i].K, c.hift_ups[i]);
  const int C0 = c.hift_base, CF = c.hift_f0_ch;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 63) & ~(size_t)63; return o; };
  const size_t o_fa = take((size_t)CF * T), o_fb = take((size_t)CF * T), o_f0 = take(T);
  const size_t o_cum = take((size_t)T * H), o_ph = take((size_t)T * H), o_s = take(Ns);
  const size_t o_stft = take((size_t)18 * Fs), o_x0 = take((size_t)C0 * T);
  size_t maxCL = 0;
  { int L = T; for (int i = 0; i < c.hift_n_ups; i++) { L = L * c.hift_ups[i]; size_t cl = (size_t)(C0 >> (i + 1)) * (L + 1); if (cl > maxCL) maxCL = cl; } }
  const size_t o_x = take(maxCL), o_xs = take(maxCL), o_si = take(maxCL), o_work = take(maxCL), o_tmp = take(maxCL);
  const size_t o_post = take((size_t)18 * F), o_fbuf = take((size_t)16 * F);
  float* ws = (float*)h->ws.get(off * sizeof(float));
  HVX_CHECK(ws, HVX_ERR_CUDA, "hift_t: workspace allocation of %zu bytes failed", off * sizeof(float));
  float *fa = ws + o_fa, *fb = ws + o_fb, *f0 = ws + o_f0, *cum = ws + o_cum, *phase = ws + o_ph, *s = ws + o_s;
  float *sstft = ws + o_stft, *x0 = ws + o_x0, *x = ws + o_x, *xs = ws + o_xs, *si = ws + o_si, *work = ws + o_work;
  float *tmp = ws + o_tmp, *post = ws + o_post, *fbuf = ws + o_fbuf;
  hvx_status rc;

  // ---- F0 predictor: 5 x (conv k3 pad 1 + ELU), Linear, abs (f0_predictor.py:25-55)
  if (f0_in) {
    HVX_CUDA(cudaMemcpyAsync(f0, f0_in, sizeof(float) * T, cudaMemcpyDeviceToDevice, st));
  } else {
    const float* a = mel; float* b = fa;
    for (int i = 0; i < 5; i++) {
      ConvOpt o; o.post_elu = 1; o.pad_left = (h->f0c[i].K - 1) / 2;
      if ((rc = run_conv(e, st, h->f0c[i], a, T, b, o))) return rc;
      a = b; b = (b == fa) ? fb : fa;
    }
    f0_head_kernel<<<cdiv(T, 128), 128, 0, st>>>(a, h->cls_w, h->cls_b, f0, CF, T);
    HVX_LAUNCH_CHECK(e);
  }
  if (f0_out) HVX_CUDA(cudaMemcpyAsync(f0_out, f0, sizeof(float) * T, cudaMemcpyDeviceToDevice, st));

  // ---- harmonic source
  if (n_cache < Ns) {
    phase_scan_kernel<<<1, H * 32, H * 32 * sizeof(double), st>>>(f0, cum, T, H, (float)c.hift_sr);
    HVX_LAUNCH_CHECK(e);
    frame_phase_kernel<<<cdiv(T * H, 256), 256, 0, st>>>(cum, phase, T * H, (float)frame);
    HVX_LAUNCH_CHECK(e);
    source_synth_nc_kernel<<<cdiv(Ns, 256), 256, 0, st>>>(f0, phase, noise, h->lin_w, h->lin_b, s, Ns, frame, H, T, (float)(1.0 / (double)frame));
    HVX_LAUNCH_CHECK(e);
  }
  if (n_cache > 0) HVX_CUDA(cudaMemcpyAsync(s, cache_source, sizeof(float) * n_cache, cudaMemcpyDeviceToDevice, st));
  if (src) HVX_CUDA(cudaMemcpyAsync(src, s, sizeof(float) * Ns, cudaMemcpyDeviceToDevice, st));
  stft16_kernel<<<cdiv(Fs, 128), 128, 0, st>>>(s, sstft, Ns, Fs);
  HVX_LAUNCH_CHECK(e);

  // ---- decode (generator.py:506-540)
  { ConvOpt o; o.pad_left = (h->conv_pre.K - 1) / 2;
    if ((rc = run_conv(e, st, h->conv_pre, mel, T, x0, o))) return rc; }
  const float* cur = x0;
  int L = T, down = up_prod;
  for (int i = 0; i < c.hift_n_ups; i++) {
    const int u = c.hift_ups[i], K = h->ups[i].K;
    const int last = (i == c.hift_n_ups - 1);
    const int Lo = L * u + (last ? 1 : 0);
    // ConvTranspose1d(k, stride u, padding (k-u)/2): zero-stuff, left pad k-1-(k-u)/2, output length u*L
    { ConvOpt o; o.pre_act = 1; o.slope = 0.1f; o.up = u; o.zstuff = 1; o.pad_left = K - 1 - (K - u) / 2; o.Lout = L * u; o.reflect1 = last;
      if ((rc = run_conv(e, st, h->ups[i], cur, L, x, o))) return rc; }
    down /= u;
    { const ConvW& d = h->sdown[i];                                // Conv1d(18, C, 2*down, down, padding=down/2) or k1 (:452-460)
      dim3 grid(cdiv(Lo, 128), d.Cout);
      conv1d_strided_kernel<<<grid, 128, 0, st>>>(sstft, d.w, d.b, si, d.Cin, d.Cout, d.K, down, down > 1 ? down / 2 : 0, Fs, Lo);
      HVX_LAUNCH_CHECK(e); }
    if ((rc = run_resblock(e, st, h->srb[i], c.hift_n_dil, c.hift_rb_d, si, work, tmp, Lo, x, x, 0, 1.0f, true))) return rc;
    for (int j = 0; j < c.hift_n_rb; j++)
      if ((rc = run_resblock(e, st, h->rb[i * c.hift_n_rb + j], c.hift_n_dil, c.hift_rb_d, x, work, tmp, Lo, nullptr, xs,
                             j > 0, 1.0f / c.hift_n_rb, true))) return rc;
    cur = xs; L = Lo;
  }
  { ConvOpt o; o.pre_act = 1; o.slope = 0.01f; o.pad_left = (h->conv_post.K - 1) / 2;
    if ((rc = run_conv(e, st, h->conv_post, cur, L, post, o))) return rc; }
  istft_frames_kernel<<<cdiv(F, 128), 128, 0, st>>>(post, fbuf, F);
  HVX_LAUNCH_CHECK(e);
  istft_ola_kernel<<<cdiv(Ns, 256), 256, 0, st>>>(fbuf, wav, F, Ns, 0.99f);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

// Classic HiFi-GAN Generator.forward (matcha/hifigan/models.py:148-193; ResBlock1 :14-93): conv_pre k7 -> per stage
// leaky-relu(0.1) + weight-normed ConvTranspose1d(k, u, padding (k-u)/2) + mean of the ResBlock1s -> leaky-relu(0.01) ->
// conv_post k7 -> tanh.  Packed tensors (weights.pack_hifigan): conv_pre, ups.i (tap-reversed), rb.n.c1.j / rb.n.c2.j,
// conv_post; stage count / rates / kernel sizes from hvx_config.hift_* (no F0, source or ISTFT head on this variant).
extern "C" hvx_status hvx_hifigan_vocode(hvx_engine* e, const float* mel, int T, float* wav, void* stream) {
  HVX_CHECK(e && mel && wav && T >= 1, HVX_ERR_ARG, "hifigan: bad arguments (T=%d)", T);
  HVX_LOCK(e, HVX_STAGE_HIFT);
  const hvx_config& c = e->cfg;
  HVX_CHECK(c.hift_n_ups >= 1 && c.hift_n_ups <= 4 && c.hift_n_rb >= 1 && c.hift_n_rb <= 3 && c.hift_n_dil >= 1 && c.hift_n_dil <= 4,
            HVX_ERR_UNSUPPORTED, "hifigan: %d stages x %d resblocks x %d dilations unsupported", c.hift_n_ups, c.hift_n_rb, c.hift_n_dil);
  if (!e->hift) e->hift = new HiftState();
  HiftState* h = e->hift;
  cudaStream_t st = (cudaStream_t)stream;
  hvx_status rc;
  if ((rc = get_conv(e, "conv_pre", &h->conv_pre))) return rc;
  if ((rc = get_conv(e, "conv_post", &h->conv_post))) return rc;
  HVX_CHECK(h->conv_post.Cout == 1, HVX_ERR_STATE, "hifigan: conv_post has %d output channels (the stage holds HiFT weights?)", h->conv_post.Cout);
  for (int i = 0; i < c.hift_n_ups; i++) {
    if ((rc = get_conv(e, "ups." + std::to_string(i), &h->ups[i]))) return rc;
    HVX_CHECK(h->ups[i].K >= c.hift_ups[i] && (h->ups[i].K - c.hift_ups[i]) % 2 == 0, HVX_ERR_UNSUPPORTED,
              "hifigan: upsample kernel %d / rate %d: only k-u even is built (output length u*L)", h->ups[i].K, c.hift_ups[i]);
    for (int j = 0; j < c.hift_n_rb; j++) {
      ResBlockW& r = h->rb[i * c.hift_n_rb + j];
      const std::string pfx = "rb." + std::to_string(i * c.hift_n_rb + j);
      for (int d = 0; d < c.hift_n_dil; d++) {
        if ((rc = get_conv(e, pfx + ".c1." + std::to_string(d), &r.c1[d]))) return rc;
        if ((rc = get_conv(e, pfx + ".c2." + std::to_string(d), &r.c2[d]))) return rc;
        r.a1[d] = r.a2[d] = nullptr;
      }
    }
  }
  const int C0 = h->conv_pre.Cout;
  size_t off = 0, maxCL = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 63) & ~(size_t)63; return o; };
  { int L = T; for (int i = 0; i < c.hift_n_ups; i++) { L *= c.hift_ups[i]; size_t cl = (size_t)h->ups[i].Cout * L; if (cl > maxCL) maxCL = cl; } }
  const size_t o_x0 = take((size_t)C0 * T), o_x = take(maxCL), o_xs = take(maxCL), o_work = take(maxCL), o_tmp = take(maxCL);
  float* ws = (float*)h->ws.get(off * sizeof(float));
  HVX_CHECK(ws, HVX_ERR_CUDA, "hifigan: workspace allocation of %zu bytes failed", off * sizeof(float));
  float *x0 = ws + o_x0, *x = ws + o_x, *xs = ws + o_xs, *work = ws + o_work, *tmp = ws + o_tmp;

  { ConvOpt o; o.pad_left = (h->conv_pre.K - 1) / 2;
    if ((rc = run_conv(e, st, h->conv_pre, mel, T, x0, o))) return rc; }
  const float* cur = x0;
  int L = T;
  for (int i = 0; i < c.hift_n_ups; i++) {
    const int u = c.hift_ups[i], K = h->ups[i].K;
    { ConvOpt o; o.pre_act = 1; o.slope = 0.1f; o.up = u; o.zstuff = 1; o.pad_left = K - 1 - (K - u) / 2; o.Lout = L * u;
      if ((rc = run_conv(e, st, h->ups[i], cur, L, x, o))) return rc; }
    L *= u;
    for (int j = 0; j < c.hift_n_rb; j++)
      if ((rc = run_resblock(e, st, h->rb[i * c.hift_n_rb + j], c.hift_n_dil, c.hift_rb_d, x, work, tmp, L, nullptr, xs,
                             j > 0, 1.0f / c.hift_n_rb, true, true))) return rc;
    cur = xs;
  }
  { ConvOpt o; o.pre_act = 1; o.slope = 0.01f; o.pad_left = (h->conv_post.K - 1) / 2; o.post_elu = 2;
    if ((rc = run_conv(e, st, h->conv_post, cur, L, wav, o))) return rc; }
  return HVX_OK;
}
