// Feature front-end on the GPU: the reference's SpectrogramExtractor (data/data_loader.py:33-88) -- dither, pre-emphasis,
// centred STFT (reflect padding, window zero-padded to n_fft), power spectrum, mel filterbank, log1p, per-feature
// normalisation over time -- for a whole collated batch, written in the [B, F, T] fp32 layout the model's first kernel reads
// (data_loader.py:149-158).  The reference runs this per utterance, single-threaded, inside the DataLoader.
//
//   logmel_kernel    : one warp per frame: gather + pre-emphasis + window -> 2^m-point radix-2 FFT in shared memory -> power
//                      -> mel (filterbank staged in shared memory) -> log1p -> feats [B, T_max, n_mels] (time-major scratch)
//   feat_norm_kernel : per (utterance, feature): mean / unbiased std over the utterance's frames (two passes over the
//                      L2-resident scratch), normalise, transpose to [B, n_mels, T_max], zero the padding frames
#include "common.cuh"

namespace w2l {

constexpr int kFeatWarps = 8;

__device__ __forceinline__ int reflect_index(int i, int n) {   // torch 'reflect' padding (no edge repeat); n >= 2, |overshoot| < n
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

// sample of the dithered, pre-emphasised signal: y[0] = x[0], y[i] = x[i] - preemph * x[i-1]   (data_loader.py:66-67)
__device__ __forceinline__ float preemph_sample(const float* __restrict__ x, const float* __restrict__ noise, float dither, float preemph,
                                                int i) {
  float a = x[i];
  if (noise) a = fmaf(noise[i], dither, a);
  if (i == 0) return a;
  float b = x[i - 1];
  if (noise) b = fmaf(noise[i - 1], dither, b);
  return a - preemph * b;
}

__global__ void __launch_bounds__(kFeatWarps * 32)
logmel_kernel(const float* __restrict__ audio, int64_t audio_stride, const float* __restrict__ noise, const int32_t* __restrict__ audio_lens,
              int n_fft, int log2_fft, int win_length, int hop, const float* __restrict__ window, const float* __restrict__ fb, int n_mels,
              float dither, float preemph, float log_guard, float* __restrict__ feats, int T_max) {
  extern __shared__ __align__(16) float smem[];
  const int n_bins = n_fft / 2 + 1;
  const int fb_pitch = n_bins | 1;                                // odd pitch: the mel dot products read conflict-free
  float* s_fb = smem;                                             // [n_mels][fb_pitch]
  float2* s_tw = reinterpret_cast<float2*>(s_fb + n_mels * fb_pitch + ((n_mels * fb_pitch) & 1));   // [n_fft/2] twiddles
  float* s_win = reinterpret_cast<float*>(s_tw + n_fft / 2);      // [win_length]
  float2* s_buf = reinterpret_cast<float2*>(s_win + win_length + (win_length & 1));                  // [warps][n_fft]
  int* s_rng = reinterpret_cast<int*>(s_buf + kFeatWarps * n_fft);                                   // [n_mels][2] support of each filter
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < n_mels * n_bins; i += blockDim.x) {
    const int f = i / n_bins, k = i - f * n_bins;
    s_fb[f * fb_pitch + k] = fb[i];
  }
  for (int i = tid; i < n_fft / 2; i += blockDim.x) {
    float s, c;
    sincospif(-2.f * (float)i / (float)n_fft, &s, &c);
    s_tw[i] = make_float2(c, s);
  }
  for (int i = tid; i < win_length; i += blockDim.x) s_win[i] = window[i];
  __syncthreads();
  // the filters are narrow triangles: find each one's support once, the dot products below only walk [lo, hi)
  for (int f = tid; f < n_mels; f += blockDim.x) {
    const float* row = s_fb + f * fb_pitch;
    int lo = n_bins, hi = 0;
    for (int k = 0; k < n_bins; ++k)
      if (row[k] != 0.f) {
        lo = min(lo, k);
        hi = k + 1;
      }
    s_rng[2 * f] = lo;
    s_rng[2 * f + 1] = hi;
  }
  __syncthreads();

  const int b = blockIdx.y;
  const int L = audio_lens[b];
  const int n_frames = L > 0 ? 1 + L / hop : 0;
  const float* x = audio + (int64_t)b * audio_stride;
  const float* nz = noise ? noise + (int64_t)b * audio_stride : nullptr;
  float2* buf = s_buf + warp * n_fft;
  const int win_off = (n_fft - win_length) / 2;                   // torch.stft centres a short window inside n_fft
  const int frames_per_block = gridDim.x > 0 ? (T_max + gridDim.x - 1) / gridDim.x : T_max;
  const int t_begin = blockIdx.x * frames_per_block, t_end = min(T_max, t_begin + frames_per_block);
  for (int t = t_begin + warp; t < t_end; t += kFeatWarps) {
    float* out = feats + ((int64_t)b * T_max + t) * n_mels;
    if (t >= n_frames) {
      for (int f = lane; f < n_mels; f += 32) out[f] = 0.f;
      continue;
    }
    // windowed frame, stored bit-reversed for the in-place decimation-in-time FFT
    const int start = t * hop - n_fft / 2;
    for (int n = lane; n < n_fft; n += 32) {
      float v = 0.f;
      const int w = n - win_off;
      if (w >= 0 && w < win_length) v = preemph_sample(x, nz, dither, preemph, reflect_index(start + n, L)) * s_win[w];
      buf[__brev((unsigned)n) >> (32 - log2_fft)] = make_float2(v, 0.f);
    }
    __syncwarp();
    for (int s = 1; s <= log2_fft; ++s) {
      const int half = 1 << (s - 1);
      for (int i = lane; i < n_fft / 2; i += 32) {
        const int j = i & (half - 1);
        const int lo = ((i >> (s - 1)) << s) + j, hi = lo + half;
        const float2 w = s_tw[j << (log2_fft - s)];
        const float2 a = buf[lo], c = buf[hi];
        const float2 tw = make_float2(c.x * w.x - c.y * w.y, c.x * w.y + c.y * w.x);
        buf[lo] = make_float2(a.x + tw.x, a.y + tw.y);
        buf[hi] = make_float2(a.x - tw.x, a.y - tw.y);
      }
      __syncwarp();
    }
    // power spectrum as the reference forms it: sqrt(re^2 + im^2) then squared (data_loader.py:69-70); kept in buf[k].x
    for (int k = lane; k < n_bins; k += 32) {
      const float2 z = buf[k];
      const float mag = sqrtf(z.x * z.x + z.y * z.y);
      buf[k].x = mag * mag;
    }
    __syncwarp();
    for (int f = lane; f < n_mels; f += 32) {
      const float* row = s_fb + f * fb_pitch;
      float acc = 0.f;
      for (int k = s_rng[2 * f]; k < s_rng[2 * f + 1]; ++k) acc = fmaf(row[k], buf[k].x, acc);
      out[f] = log1pf(acc + log_guard);                           // np.log1p(spect + 2**-24), data_loader.py:79
    }
    __syncwarp();
  }
}

// grid (ceil(n_mels / 32), B), block (32, 8): threadIdx.x -> feature, threadIdx.y -> row lane
__global__ void __launch_bounds__(256)
feat_norm_kernel(const float* __restrict__ feats, const int32_t* __restrict__ audio_lens, int hop, int T_max, int n_mels, float eps,
                 float* __restrict__ out) {
  __shared__ float s_red[8][33];
  __shared__ float s_tile[32][33];
  const int b = blockIdx.y, f = blockIdx.x * 32 + threadIdx.x;
  const int L = audio_lens[b];
  const int n = L > 0 ? min(T_max, 1 + L / hop) : 0;
  const float* src = feats + (int64_t)b * T_max * n_mels;
  const bool ok = f < n_mels;
  auto block_sum = [&](float v) {
    s_red[threadIdx.y][threadIdx.x] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) t += s_red[y][threadIdx.x];
    __syncthreads();
    return t;
  };
  float acc = 0.f;
  if (ok)
    for (int t = threadIdx.y; t < n; t += 8) acc += src[(int64_t)t * n_mels + f];
  const float mean = n > 0 ? block_sum(acc) / (float)n : 0.f;
  acc = 0.f;
  if (ok)
    for (int t = threadIdx.y; t < n; t += 8) {
      const float d = src[(int64_t)t * n_mels + f] - mean;
      acc = fmaf(d, d, acc);
    }
  const float var = block_sum(acc) / (float)max(n - 1, 1);        // torch.std: unbiased (data_loader.py:82)
  const float inv = 1.f / (sqrtf(var) + eps);                      // std += epsilon; spect / std
  float* dst = out + (int64_t)b * n_mels * T_max;
  for (int t0 = 0; t0 < T_max; t0 += 32) {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int t = t0 + i;
      s_tile[i][threadIdx.x] = (ok && t < n) ? (src[(int64_t)t * n_mels + f] - mean) * inv : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {                    // i -> feature within the tile, threadIdx.x -> frame: coalesced rows
      const int ff = blockIdx.x * 32 + i, t = t0 + threadIdx.x;
      if (ff < n_mels && t < T_max) dst[(int64_t)ff * T_max + t] = s_tile[threadIdx.x][i];
    }
    __syncthreads();
  }
}

}  // namespace w2l

extern "C" {

size_t w2l_logmel_workspace_bytes(int32_t B, int32_t T_max, int32_t n_mels) {
  return (size_t)(B > 0 ? B : 0) * (size_t)(T_max > 0 ? T_max : 0) * (size_t)(n_mels > 0 ? n_mels : 0) * sizeof(float) + 256;
}

int w2l_logmel_features(const float* audio, int64_t audio_stride, const float* dither_noise, const int32_t* audio_lens, int32_t B,
                        int32_t n_fft, int32_t win_length, int32_t hop, const float* window, const float* mel_fb, int32_t n_mels,
                        float dither, float preemph, float log_guard, float norm_eps, float* out, int32_t T_max, void* workspace,
                        size_t workspace_bytes, void* stream) {
  using namespace w2l;
  W2L_REQUIRE(audio && audio_lens && window && mel_fb && out && workspace, "logmel_features: null pointer");
  W2L_REQUIRE(B >= 1 && B <= 65535 && T_max >= 1 && n_mels >= 1 && hop >= 1, "logmel_features: bad shape");
  W2L_REQUIRE(n_fft >= 64 && n_fft <= 2048 && (n_fft & (n_fft - 1)) == 0, "logmel_features: n_fft=%d must be a power of two in [64, 2048]", n_fft);
  W2L_REQUIRE(win_length >= 1 && win_length <= n_fft, "logmel_features: win_length=%d must be in [1, n_fft]", win_length);
  W2L_REQUIRE(workspace_bytes >= w2l_logmel_workspace_bytes(B, T_max, n_mels), "logmel_features: workspace too small");
  int log2_fft = 0;
  while ((1 << log2_fft) < n_fft) ++log2_fft;
  const int n_bins = n_fft / 2 + 1, fb_pitch = n_bins | 1;
  size_t smem = (size_t)(n_mels * fb_pitch + ((n_mels * fb_pitch) & 1)) * 4 + (size_t)(n_fft / 2) * 8 + (size_t)(win_length + (win_length & 1)) * 4 +
                (size_t)kFeatWarps * n_fft * 8 + (size_t)n_mels * 8;
  W2L_REQUIRE(smem <= 200 * 1024, "logmel_features: n_mels=%d x n_fft=%d needs %zu bytes of shared memory", n_mels, n_fft, smem);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    W2L_CUDA(cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* feats = reinterpret_cast<float*>(workspace);
  // enough blocks to fill the machine, few enough that the filterbank staging (66 KB per block) amortises
  int bx = (2 * num_sms() + B - 1) / B;
  const int max_bx = (T_max + 4 * kFeatWarps - 1) / (4 * kFeatWarps);
  if (bx > max_bx) bx = max_bx;
  if (bx < 1) bx = 1;
  logmel_kernel<<<dim3(bx, B), kFeatWarps * 32, smem, st>>>(audio, audio_stride, dither_noise, audio_lens, n_fft, log2_fft, win_length, hop,
                                                           window, mel_fb, n_mels, dither, preemph, log_guard, feats, T_max);
  int rc = after_launch("logmel_kernel");
  if (rc) return rc;
  feat_norm_kernel<<<dim3((n_mels + 31) / 32, B), dim3(32, 8), 0, st>>>(feats, audio_lens, hop, T_max, n_mels, norm_eps, out);
  return after_launch("feat_norm_kernel");
}

}  // extern "C"
