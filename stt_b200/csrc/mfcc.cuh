// K1: int16 PCM -> MFCC, one warp per 32 ms analysis window.
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

constexpr int kWarpsPerBlock = 4;

struct WarpScratch {
  double re[kFft];
  double im[kFft];
  float pow[kBins + 3];
  double mel[kMaxChannels];
};

// One WARP per analysis window (v1 used one 256-thread CTA per window and spent its time in __syncthreads; a warp
// needs only __syncwarp between FFT stages and 20 windows are in flight per SM instead of 8).
__device__ __forceinline__ void mfcc_window(const MfccTables& tb, WarpScratch& w, const int16_t* pcm, int n_valid,
                                            float* out_f32, __half* out_f16) {
  const int lane = threadIdx.x & 31;
  for (int j = lane; j < kFft; j += 32) {
    double v = 0.0;
    if (j < tb.win_len && j < n_valid) v = (double)((float)pcm[j] * (1.0f / 32768.0f)) * tb.hann[j];
    const int r = __brev((unsigned)j) >> (32 - 9);
    w.re[r] = v;
    w.im[r] = 0.0;
  }
  __syncwarp();
#pragma unroll 1
  for (int len = 2; len <= kFft; len <<= 1) {
    const int half = len >> 1;
    const int tstep = kFft / len;
#pragma unroll 4
    for (int bf = lane; bf < kFft / 2; bf += 32) {
      const int k = bf & (half - 1);
      const int i = ((bf - k) << 1) + k;
      const int j = i + half;
      const double wr = tb.tw_re[k * tstep], wi = tb.tw_im[k * tstep];
      const double xr = w.re[j] * wr - w.im[j] * wi, xi = w.re[j] * wi + w.im[j] * wr;
      const double ur = w.re[i], ui = w.im[i];
      w.re[j] = ur - xr;
      w.im[j] = ui - xi;
      w.re[i] = ur + xr;
      w.im[i] = ui + xi;
    }
    __syncwarp();
  }
  for (int i = lane; i < kBins; i += 32) w.pow[i] = (float)(w.re[i] * w.re[i] + w.im[i] * w.im[i]);
  __syncwarp();
  for (int ch = lane; ch < tb.n_channels; ch += 32) {
    // same per-channel accumulation order as the reference's single pass over bins
    double acc = 0.0;
    const int first = tb.chan_first_bin[ch], last = tb.chan_last_bin[ch];
    if (first >= 0) {
      for (int i = first; i <= last; ++i) {
        const double spec_val = sqrt((double)w.pow[i]);
        const double weighted = spec_val * tb.weights[i];
        const int m = tb.band_mapper[i];
        if (m == ch) acc += weighted;
        else if (m + 1 == ch) acc += spec_val - weighted;
      }
    }
    if (acc < 1e-12) acc = 1e-12;
    w.mel[ch] = log(acc);
  }
  __syncwarp();
  {
    float f = 0.f;
    if (lane < tb.n_dct) {
      double sum = 0.0;
      for (int j = 0; j < tb.n_channels; ++j) sum += tb.cosines[lane * tb.n_channels + j] * w.mel[j];
      f = (float)sum;
      if (out_f32) out_f32[lane] = f;
    }
    if (out_f16) out_f16[lane] = __float2half_rn(f);  // kFeatLanes == 32: lanes >= n_dct store zeros
  }
  __syncwarp();
}

// grid-stride over (utterance, frame) pairs, one warp each
__global__ void __launch_bounds__(kWarpsPerBlock * 32) mfcc_batch_kernel(const MfccTables tb, const BatchJob job, int n_utt) {
  __shared__ WarpScratch scratch[kWarpsPerBlock];
  WarpScratch& w = scratch[threadIdx.x >> 5];
  const long long n_items = (long long)n_utt * job.frames_per_utt;
  for (long long item = (long long)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); item < n_items;
       item += (long long)gridDim.x * kWarpsPerBlock) {
    const int b = (int)(item / job.frames_per_utt), f = (int)(item % job.frames_per_utt);
    const int n = job.n_samples[b];
    if (f >= n_frames_for(n, tb.win_len, tb.win_step)) continue;
    const long long start = (long long)f * tb.win_step;
    int n_valid = (int)(n - start);
    if (n_valid > tb.win_len) n_valid = tb.win_len;
    if (n_valid < 0) n_valid = 0;
    mfcc_window(tb, w, job.pcm + b * job.stride + start, n_valid,
                job.out_f32 ? job.out_f32 + ((size_t)b * job.frames_per_utt + f) * tb.n_dct : nullptr,
                job.out_f16 ? job.out_f16 + ((size_t)b * job.f16_frames_per_utt + f + job.f16_row_offset) * kFeatLanes
                            : nullptr);
  }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) mfcc_jobs_kernel(const MfccTables tb, const FrameJob* jobs, int n_jobs) {
  __shared__ WarpScratch scratch[kWarpsPerBlock];
  WarpScratch& w = scratch[threadIdx.x >> 5];
  for (int item = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5); item < n_jobs; item += gridDim.x * kWarpsPerBlock) {
    const FrameJob j = jobs[item];
    mfcc_window(tb, w, j.pcm, j.n_valid, j.out_f32, j.out_f16);
  }
}

}  // namespace sttmfcc
