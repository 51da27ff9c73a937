// Zero-shot frontend features on the GPU: the two filterbank front ends that sit directly before the hot
// path, as one kernel family "framed signal x folded linear basis -> |.| or |.|^2 -> filterbank -> log":
//   * the 24 kHz prompt mel that becomes the flow's prompt_feat: mel_spectrogram (matcha/utils/audio.py:42-82; reflect pad
//     (n_fft-hop)/2, hann STFT n_fft 1920 / hop 480, sqrt(re^2+im^2+1e-9), mel matmul, log(clamp(., 1e-5)));
//   * the 16 kHz Kaldi fbank fed to the speaker encoder: kaldi.fbank(num_mel_bins=80, dither=0) followed by the mean
//     subtraction of cosyvoice/cli/frontend.py:108-112 (DC removal, pre-emphasis 0.97, povey window, 512-point power
//     spectrum, mel banks, log(max(., eps))).
// Everything before the non-linearity is linear in the raw frame, so the host folds window / DC removal / pre-emphasis / DFT
// into one fp32 basis [frame_len][2*n_bins] (flowmirror_hydravox_b200/frontend.py, built in float64); n_fft = 1920 is not a
// power of two and a prompt is a few hundred frames, so the transform is a plain fp32 tiled GEMM over the overlapping-frame
// view of the waveform (row stride = hop), not an FFT.  Restated in oracle/frontend_ref.py.
#include "common.cuh"

namespace hvx {

constexpr int FE_TM = 64, FE_TN = 64, FE_TK = 16;

__device__ __forceinline__ int fe_reflect(int i, int n) {          // torch 'reflect' padding index (no edge repeat)
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// spec[f][j] = sum_n wav[reflect(f*hop + n - pad)] * basis[n][j]     (f < n_frames, j < ncol)
__global__ void __launch_bounds__(256) fe_frames_gemm_kernel(const float* __restrict__ wav, int n_samples, int hop, int pad,
                                                              const float* __restrict__ basis, int frame_len, int ncol,
                                                              float* __restrict__ spec, int n_frames) {
  __shared__ float sa[FE_TK][FE_TM + 1];
  __shared__ __align__(16) float sb[FE_TK][FE_TN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int f0 = blockIdx.y * FE_TM, j0 = blockIdx.x * FE_TN;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.f;
  for (int k0 = 0; k0 < frame_len; k0 += FE_TK) {
    for (int i = tid; i < FE_TK * FE_TM; i += 256) {
      const int k = i / FE_TM, m = i - k * FE_TM;
      const int f = f0 + m, n = k0 + k;
      float v = 0.f;
      if (f < n_frames && n < frame_len) {
        int idx = f * hop + n - pad;
        if (pad > 0) idx = fe_reflect(idx, n_samples);
        if (idx >= 0 && idx < n_samples) v = __ldg(wav + idx);
      }
      sa[k][m] = v;
    }
    for (int i = tid; i < FE_TK * FE_TN; i += 256) {
      const int k = i / FE_TN, j = i - k * FE_TN;
      sb[k][j] = (k0 + k < frame_len && j0 + j < ncol) ? __ldg(basis + (size_t)(k0 + k) * ncol + j0 + j) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < FE_TK; k++) {
      float av[4];
#pragma unroll
      for (int a = 0; a < 4; a++) av[a] = sa[k][ty * 4 + a];
      const float4 bv = *reinterpret_cast<const float4*>(&sb[k][tx * 4]);
#pragma unroll
      for (int a = 0; a < 4; a++) {
        acc[a][0] = fmaf(av[a], bv.x, acc[a][0]); acc[a][1] = fmaf(av[a], bv.y, acc[a][1]);
        acc[a][2] = fmaf(av[a], bv.z, acc[a][2]); acc[a][3] = fmaf(av[a], bv.w, acc[a][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; a++) {
    const int f = f0 + ty * 4 + a;
    if (f >= n_frames) continue;
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const int j = j0 + tx * 4 + b;
      if (j < ncol) spec[(size_t)f * ncol + j] = acc[a][b];
    }
  }
}

// out[f][m] = log(max(sum_k fb[m][k] * g(re_k, im_k), floor)),  g = re^2+im^2 (power) or sqrt(re^2+im^2+mag_eps)
__global__ void __launch_bounds__(128) fe_filterbank_kernel(const float* __restrict__ spec, int n_bins, const float* __restrict__ fb,
                                                             int n_mels, int power, float mag_eps, float floor_v,
                                                             float* __restrict__ out, int n_frames, int channel_major) {
  extern __shared__ float s_mag[];
  const int f = blockIdx.x;
  const float* sp = spec + (size_t)f * 2 * n_bins;
  for (int k = threadIdx.x; k < n_bins; k += blockDim.x) {
    const float re = sp[k], im = sp[n_bins + k];
    const float p = re * re + im * im;
    s_mag[k] = power ? p : sqrtf(p + mag_eps);
  }
  __syncthreads();
  for (int m = threadIdx.x; m < n_mels; m += blockDim.x) {
    const float* w = fb + (size_t)m * n_bins;
    float a0 = 0.f, a1 = 0.f;
    int k = 0;
    for (; k + 1 < n_bins; k += 2) { a0 = fmaf(__ldg(w + k), s_mag[k], a0); a1 = fmaf(__ldg(w + k + 1), s_mag[k + 1], a1); }
    if (k < n_bins) a0 = fmaf(__ldg(w + k), s_mag[k], a0);
    const float v = logf(fmaxf(a0 + a1, floor_v));
    if (channel_major) out[(size_t)m * n_frames + f] = v;
    else out[(size_t)f * n_mels + m] = v;
  }
}

// out[f][m] -= mean_f out[f][m]   (frame-major; cosyvoice/cli/frontend.py:112)
__global__ void __launch_bounds__(256) fe_sub_mean_kernel(float* __restrict__ out, int n_frames, int n_mels) {
  __shared__ double s_red[256];
  const int m = blockIdx.x;
  double s = 0.0;
  for (int f = threadIdx.x; f < n_frames; f += 256) s += (double)out[(size_t)f * n_mels + m];
  s_red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o; o >>= 1) { if (threadIdx.x < o) s_red[threadIdx.x] += s_red[threadIdx.x + o]; __syncthreads(); }
  const float mean = (float)(s_red[0] / (double)n_frames);
  for (int f = threadIdx.x; f < n_frames; f += 256) out[(size_t)f * n_mels + m] -= mean;
}

// whisper's dynamic-range step on a natural-log mel x (n elements): y = x * scale (log10); y = max(y, max(y) - range);
// y = (y + add) / div   (whisper/audio.py log_mel_spectrogram).  One block: n is a few 100 k.
__global__ void __launch_bounds__(1024) fe_whisper_post_kernel(float* __restrict__ x, int n, float scale, float range, float add, float div) {
  __shared__ float s_red[32];
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 1024) m = fmaxf(m, x[i] * scale);
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = s_red[0];
  for (int i = 1; i < 32; i++) m = fmaxf(m, s_red[i]);
  const float lo = m - range;
  for (int i = threadIdx.x; i < n; i += 1024) x[i] = (fmaxf(x[i] * scale, lo) + add) / div;
}

}  // namespace hvx

using namespace hvx;

extern "C" hvx_status hvx_frontend_whisper_post(hvx_engine* e, float* logmel, int n, float scale, float range, float add, float div,
                                                void* stream) {
  HVX_CHECK(e && logmel && n >= 1 && div != 0.f, HVX_ERR_ARG, "frontend: bad argument");
  fe_whisper_post_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logmel, n, scale, range, add, div);
  HVX_LAUNCH_CHECK(e);
  return HVX_OK;
}

extern "C" hvx_status hvx_frontend_fbank(hvx_engine* e, const float* wav, int n_samples, int frame_len, int hop, int pad_reflect,
                                         const float* basis, int n_bins, const float* fb, int n_mels, int power, float mag_eps,
                                         float log_floor, int subtract_mean, int channel_major, float* out, int n_frames,
                                         void* stream) {
  HVX_CHECK(e && wav && basis && fb && out, HVX_ERR_ARG, "frontend: null argument");
  HVX_CHECK(frame_len >= 1 && hop >= 1 && n_bins >= 1 && n_mels >= 1 && pad_reflect >= 0 && n_samples >= 1, HVX_ERR_ARG, "frontend: bad geometry");
  HVX_CHECK(pad_reflect < n_samples, HVX_ERR_ARG, "frontend: reflect padding %d needs more than %d samples", pad_reflect, n_samples);
  const int expect = (n_samples + 2 * pad_reflect - frame_len) / hop + 1;
  HVX_CHECK(n_samples + 2 * pad_reflect >= frame_len && n_frames >= 1 && n_frames <= expect, HVX_ERR_ARG,
            "frontend: %d samples give %d frames, caller asked for %d", n_samples, n_samples + 2 * pad_reflect >= frame_len ? expect : 0, n_frames);
  HVX_CHECK(!(subtract_mean && channel_major), HVX_ERR_UNSUPPORTED, "frontend: mean subtraction is built for the frame-major layout");
  HVX_CHECK((size_t)n_bins * sizeof(float) <= 48 * 1024, HVX_ERR_UNSUPPORTED, "frontend: %d bins exceed the shared-memory row", n_bins);
  cudaStream_t st = (cudaStream_t)stream;
  const int ncol = 2 * n_bins;
  float* spec = (float*)e->fe_ws.get((size_t)n_frames * ncol * sizeof(float));
  HVX_CHECK(spec, HVX_ERR_CUDA, "frontend: workspace allocation failed");
  dim3 grid(cdiv(ncol, FE_TN), cdiv(n_frames, FE_TM));
  fe_frames_gemm_kernel<<<grid, 256, 0, st>>>(wav, n_samples, hop, pad_reflect, basis, frame_len, ncol, spec, n_frames);
  HVX_LAUNCH_CHECK(e);
  fe_filterbank_kernel<<<n_frames, 128, n_bins * sizeof(float), st>>>(spec, n_bins, fb, n_mels, power, mag_eps, log_floor, out, n_frames,
                                                                        channel_major);
  HVX_LAUNCH_CHECK(e);
  if (subtract_mean) {
    fe_sub_mean_kernel<<<n_mels, 256, 0, st>>>(out, n_frames, n_mels);
    HVX_LAUNCH_CHECK(e);
  }
  return HVX_OK;
}
