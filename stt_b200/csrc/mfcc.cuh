// K1: int16 PCM -> MFCC, one CTA per 32 ms analysis window.
//
// Restates, per window, the reference chain (all double precision internally, like the TFLite kernels):
//   stt.cc:105-128                    int16 -> f32 (x 1/32768), 512-sample window, hop 320, zero-padded tail
//   internal/spectrogram.cc:30-37     periodic Hann; :224-241 real FFT (any exact DFT agrees to ~1e-13);
//                                     :175-183 re^2+im^2 in double, stored as float
//   internal/mfcc_mel_filterbank.cc:172-197   sqrt, 40 triangular bands (tables built on the host exactly as
//                                     :40-167 does, in the same accumulation order per channel)
//   internal/mfcc.cc:54-60            log(max(x, 1e-12));  internal/mfcc_dct.cc:68-74  DCT-II, first n_dct coeffs
// Output: fp32 [frame, n_dct] (optional) and the fp16, 32-lane padded stream the windowed GEMM reads via TMA.
// HBM-bound in principle (372 KB per 10 s utterance) but tiny; tables stay in L1/L2.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace sttmfcc {

constexpr int kFft = 512;
constexpr int kBins = kFft / 2 + 1;
constexpr int kMaxChannels = 64;
constexpr int kFeatLanes = 32;  // padded lanes per frame in the fp16 stream

struct MfccTables {           // device pointers, built by the host (model.cc) in double precision
  const double* hann;         // [win_len]
  const double* tw_re;        // [256] cos(-2*pi*m/512)
  const double* tw_im;        // [256] sin(-2*pi*m/512)
  const double* weights;      // [kBins]
  const int* band_mapper;     // [kBins]
  const int* chan_first_bin;  // [n_channels] first bin contributing to channel c (left side), or -1
  const int* chan_last_bin;   // [n_channels] last bin contributing to channel c (right side)
  const double* cosines;      // [n_dct, n_channels]
  int win_len, win_step, n_channels, n_dct, start_index, end_index;
};

struct FrameJob {       // one analysis window
  const int16_t* pcm;   // first sample of the window
  int n_valid;          // samples available (<= win_len); the rest is zero
  float* out_f32;       // [n_dct] or nullptr
  __half* out_f16;      // [kFeatLanes] or nullptr (lanes >= n_dct are written as zero)
};

// Batch addressing without a job list: utterance b, frame f.
struct BatchJob {
  const int16_t* pcm;       // [B, stride] int16
  const int* n_samples;     // [B]
  long long stride;         // samples between utterances
  int frames_per_utt;       // grid.x = B * frames_per_utt (padded count; frames >= n_frames(b) are skipped)
  float* out_f32;           // [B, frames_per_utt, n_dct] or nullptr
  __half* out_f16;          // [B, f16_frames_per_utt, 32]; frame f is written at row f + f16_row_offset
  int f16_frames_per_utt, f16_row_offset;
};

__device__ __forceinline__ int n_frames_for(int n, int win_len, int win_step) {
  // full windows + the flush window (stt.cc:236-241 always emits one more, zero-padded)
  return (n >= win_len ? (n - win_len) / win_step + 1 : 0) + 1;
}

__device__ __forceinline__ void mfcc_window(const MfccTables& tb, const int16_t* pcm, int n_valid, float* out_f32,
                                            __half* out_f16) {
  __shared__ double s_re[kFft];
  __shared__ double s_im[kFft];
  __shared__ float s_pow[kBins];
  __shared__ double s_mel[kMaxChannels];
  const int tid = threadIdx.x;  // 256 threads
  for (int j = tid; j < kFft; j += blockDim.x) {
    double v = 0.0;
    if (j < tb.win_len && j < n_valid) v = (double)((float)pcm[j] * (1.0f / 32768.0f)) * tb.hann[j];
    const int r = __brev((unsigned)j) >> (32 - 9);
    s_re[r] = v;
    s_im[r] = 0.0;
  }
  __syncthreads();
#pragma unroll 1
  for (int len = 2; len <= kFft; len <<= 1) {
    const int half = len >> 1;
    for (int bf = tid; bf < kFft / 2; bf += blockDim.x) {
      const int k = bf & (half - 1);
      const int i = ((bf - k) << 1) + k;
      const int j = i + half;
      const int m = k * (kFft / len);
      const double wr = tb.tw_re[m], wi = tb.tw_im[m];
      const double xr = s_re[j] * wr - s_im[j] * wi, xi = s_re[j] * wi + s_im[j] * wr;
      const double ur = s_re[i], ui = s_im[i];
      s_re[j] = ur - xr;
      s_im[j] = ui - xi;
      s_re[i] = ur + xr;
      s_im[i] = ui + xi;
    }
    __syncthreads();
  }
  for (int i = tid; i < kBins; i += blockDim.x) s_pow[i] = (float)(s_re[i] * s_re[i] + s_im[i] * s_im[i]);
  __syncthreads();
  if (tid < tb.n_channels) {
    // same per-channel accumulation order as the reference's single pass over bins
    double acc = 0.0;
    const int first = tb.chan_first_bin[tid], last = tb.chan_last_bin[tid];
    if (first >= 0) {
      for (int i = first; i <= last; ++i) {
        const double spec_val = sqrt((double)s_pow[i]);
        const double weighted = spec_val * tb.weights[i];
        const int ch = tb.band_mapper[i];
        if (ch == tid) acc += weighted;
        else if (ch + 1 == tid) acc += spec_val - weighted;
      }
    }
    if (acc < 1e-12) acc = 1e-12;
    s_mel[tid] = log(acc);
  }
  __syncthreads();
  if (tid < kFeatLanes) {
    float f = 0.f;
    if (tid < tb.n_dct) {
      double sum = 0.0;
      for (int j = 0; j < tb.n_channels; ++j) sum += tb.cosines[tid * tb.n_channels + j] * s_mel[j];
      f = (float)sum;
      if (out_f32) out_f32[tid] = f;
    }
    if (out_f16) out_f16[tid] = __float2half_rn(f);
  }
}

__global__ void __launch_bounds__(256) mfcc_batch_kernel(const MfccTables tb, const BatchJob job) {
  const int b = blockIdx.x / job.frames_per_utt, f = blockIdx.x % job.frames_per_utt;
  const int n = job.n_samples[b];
  if (f >= n_frames_for(n, tb.win_len, tb.win_step)) return;
  const long long start = (long long)f * tb.win_step;
  int n_valid = (int)(n - start);
  if (n_valid > tb.win_len) n_valid = tb.win_len;
  if (n_valid < 0) n_valid = 0;
  mfcc_window(tb, job.pcm + b * job.stride + start, n_valid,
              job.out_f32 ? job.out_f32 + ((size_t)b * job.frames_per_utt + f) * tb.n_dct : nullptr,
              job.out_f16 ? job.out_f16 + ((size_t)b * job.f16_frames_per_utt + f + job.f16_row_offset) * kFeatLanes
                          : nullptr);
}

__global__ void __launch_bounds__(256) mfcc_jobs_kernel(const MfccTables tb, const FrameJob* jobs) {
  const FrameJob j = jobs[blockIdx.x];
  mfcc_window(tb, j.pcm, j.n_valid, j.out_f32, j.out_f16);
}

}  // namespace sttmfcc
