// Kaldi-compatible log-mel filterbank front-end: the feature extraction the reference runs
// on the host through torchaudio.compliance.kaldi.fbank (extract_feature.py:32-53,
// s3prl_upstream/expert.py:23-43: num_mel_bins=40, 16 kHz, hamming, 25 ms / 10 ms, everything else at the
// torchaudio defaults -- dither 0, preemphasis 0.97, remove_dc_offset, snip_edges, 512-point FFT, power
// spectrum, log) followed by the (y - mean) / std normalisation, as one kernel over a padded batch of
// waveforms.  One CTA per (frame, utterance): 400 samples -> DC removal -> pre-emphasis -> window ->
// 512-point radix-2 FFT in shared memory -> power -> mel matrix -> log -> normalise.  The work is tiny
// (100 frames per second of audio); the point is that the features are born in HBM next to the encoder.
#include "mh_b200.h"
#include "mh_common.cuh"

namespace mh {
extern long long g_launches;

constexpr int FB_NFFT = 512;
constexpr int FB_THREADS = 256;

__global__ void __launch_bounds__(FB_THREADS)
fbank_kernel(const float* __restrict__ wave, long long ld_wave, const int* __restrict__ n_samples,
             const float* __restrict__ mel_w /* [n_mel, 257] */, const float* __restrict__ mean,
             const float* __restrict__ inv_std, float* __restrict__ out /* [B, max_frames, n_mel] */, int max_frames,
             int n_mel, int frame_len, int frame_shift, float scale, float preemph, int window_type) {
  __shared__ float re[FB_NFFT], im[FB_NFFT];
  __shared__ float xs[FB_NFFT];
  __shared__ float red[FB_THREADS / 32];
  const int f = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int ns = n_samples[b];
  const int nf = ns >= frame_len ? 1 + (ns - frame_len) / frame_shift : 0;  // snip_edges = True
  float* orow = out + (static_cast<long long>(b) * max_frames + f) * n_mel;
  if (f >= nf) {  // padding frame of a shorter utterance
    for (int m = t; m < n_mel; m += FB_THREADS) orow[m] = 0.f;
    return;
  }
  const float* w = wave + b * ld_wave + static_cast<long long>(f) * frame_shift;
  // ---- load (x 2^15: the reference scales the [-1, 1) waveform to int16 range), mean
  float s = 0.f;
  for (int i = t; i < FB_NFFT; i += FB_THREADS) {
    const float v = i < frame_len ? __ldg(w + i) * scale : 0.f;
    xs[i] = v;
    s += v;
  }
  s = warp_sum(s);
  if ((t & 31) == 0) red[t >> 5] = s;
  __syncthreads();
  float mu = 0.f;
#pragma unroll
  for (int i = 0; i < FB_THREADS / 32; ++i) mu += red[i];
  mu /= frame_len;
  // ---- DC removal, pre-emphasis (first sample against itself: replicate padding), window, bit-reversed store
  for (int i = t; i < FB_NFFT; i += FB_THREADS) {
    float v = 0.f;
    if (i < frame_len) {
      const float cur = xs[i] - mu;
      const float prev = xs[i > 0 ? i - 1 : 0] - mu;
      v = cur - preemph * prev;
      float win;
      const float a = 2.0f * static_cast<float>(i) / static_cast<float>(frame_len - 1);  // angle / pi
      if (window_type == 0) win = 0.54f - 0.46f * cospif(a);                              // hamming
      else if (window_type == 1) win = 0.5f - 0.5f * cospif(a);                           // hanning
      else if (window_type == 2) win = powf(0.5f - 0.5f * cospif(a), 0.85f);              // povey
      else win = 1.0f;                                                                    // rectangular
      v *= win;
    }
    const int j = __brev(static_cast<unsigned>(i)) >> (32 - 9);
    re[j] = v;
    im[j] = 0.f;
  }
  __syncthreads();
  // ---- 512-point radix-2 decimation-in-time FFT, one butterfly per thread and stage
#pragma unroll 1
  for (int st = 0; st < 9; ++st) {
    const int half = 1 << st;
    const int j = t & (half - 1);
    const int i0 = ((t >> st) << (st + 1)) + j, i1 = i0 + half;
    float sn, cs;
    sincospif(-static_cast<float>(j) / static_cast<float>(half), &sn, &cs);  // exp(-2 pi i j / (2 half))
    const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
    const float tr = br * cs - bi * sn, ti = br * sn + bi * cs;
    re[i0] = ar + tr; im[i0] = ai + ti;
    re[i1] = ar - tr; im[i1] = ai - ti;
    __syncthreads();
  }
  // ---- power spectrum (bins 0 .. 256) into xs
  {
    const float pr = re[t], pi = im[t];
    xs[t] = pr * pr + pi * pi;
    if (t == 0) xs[256] = re[256] * re[256] + im[256] * im[256];
  }
  __syncthreads();
  // ---- mel energies: 8 lanes per filter, 32 filters per pass
  const int lane8 = t & 7;
  for (int m0 = 0; m0 < n_mel; m0 += FB_THREADS / 8) {
    const int m = m0 + (t >> 3);
    float e = 0.f;
    if (m < n_mel) {
      const float* wr = mel_w + static_cast<long long>(m) * 257;
      for (int k = lane8; k < 257; k += 8) e = fmaf(__ldg(wr + k), xs[k], e);
    }
    e += __shfl_xor_sync(0xffffffffu, e, 1);
    e += __shfl_xor_sync(0xffffffffu, e, 2);
    e += __shfl_xor_sync(0xffffffffu, e, 4);
    if (m < n_mel && lane8 == 0) {
      float v = logf(fmaxf(e, 1.1920928955078125e-07f));  // log(max(e, float32 eps)) like torchaudio
      if (mean != nullptr) v = (v - mean[m]) * inv_std[m];
      orow[m] = v;
    }
  }
}
}  // namespace mh

using namespace mh;

extern "C" int mh_fbank(const float* wave, long long ld_wave, const int* n_samples, int batch, const float* mel_weights,
                        const float* mean, const float* inv_std, float* out, int max_frames, int n_mel, int frame_len,
                        int frame_shift, float scale, float preemph, int window_type, void* stream) {
  MH_CHECK(batch > 0 && max_frames > 0 && n_mel > 0, "fbank: bad shape batch=%d frames=%d mel=%d", batch, max_frames, n_mel);
  MH_CHECK(frame_len > 1 && frame_len <= FB_NFFT && frame_shift > 0, "fbank: frame length %d must be in (1, %d]", frame_len, FB_NFFT);
  MH_CHECK((mean == nullptr) == (inv_std == nullptr), "fbank: mean and inv_std go together");
  MH_CHECK(batch <= 65535, "fbank: batch %d > 65535", batch);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  fbank_kernel<<<dim3(max_frames, batch), FB_THREADS, 0, st>>>(wave, ld_wave, n_samples, mel_weights, mean, inv_std, out,
                                                               max_frames, n_mel, frame_len, frame_shift, scale, preemph,
                                                               window_type);
  MH_LAUNCH_CHECK();
  ++g_launches;
  return 0;
}
