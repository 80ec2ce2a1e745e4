// K1: int16 PCM -> MFCC, one warp per 32 ms analysis window.
//
// Restates, per window, the reference chain (all double precision internally, like the TFLite kernels):
//   stt.cc:105-128                    int16 -> f32 (x 1/32768), 512-sample window, hop 320, zero-padded tail
//   internal/spectrogram.cc:30-37     periodic Hann; :224-241 real FFT (any exact DFT agrees to ~1e-13; here three
//                                     radix-8 passes with register butterflies);
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

// 512 complex points, padded by one element per 8 so that the radix-8 passes below store without bank conflicts.
__device__ __forceinline__ int zidx(int i) { return i + (i >> 3); }

struct WarpScratch {
  double2 z[kFft + kFft / 8];   // FFT work array; afterwards its first kBins doubles hold sqrt(power)
  float pow[kBins + 3];
  double mel[kMaxChannels];
};

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// forward 8-point DFT in registers (decimation in frequency): a[] -> X[0..7] in natural order
__device__ __forceinline__ void dft8(double2 (&a)[8]) {
  const double c = 0.70710678118654752440;
  double2 t0 = make_double2(a[0].x + a[4].x, a[0].y + a[4].y), t4 = make_double2(a[0].x - a[4].x, a[0].y - a[4].y);
  double2 t1 = make_double2(a[1].x + a[5].x, a[1].y + a[5].y), t5 = make_double2(a[1].x - a[5].x, a[1].y - a[5].y);
  double2 t2 = make_double2(a[2].x + a[6].x, a[2].y + a[6].y), t6 = make_double2(a[2].x - a[6].x, a[2].y - a[6].y);
  double2 t3 = make_double2(a[3].x + a[7].x, a[3].y + a[7].y), t7 = make_double2(a[3].x - a[7].x, a[3].y - a[7].y);
  // odd branch twiddles: w^1 = (1 - i)/sqrt2, w^2 = -i, w^3 = (-1 - i)/sqrt2
  t5 = make_double2(c * (t5.x + t5.y), c * (t5.y - t5.x));
  t6 = make_double2(t6.y, -t6.x);
  t7 = make_double2(c * (t7.y - t7.x), -c * (t7.x + t7.y));
  // two 4-point DFTs
  auto dft4 = [](double2 b0, double2 b1, double2 b2, double2 b3, double2& y0, double2& y1, double2& y2, double2& y3) {
    const double2 u0 = make_double2(b0.x + b2.x, b0.y + b2.y), u1 = make_double2(b0.x - b2.x, b0.y - b2.y);
    const double2 u2 = make_double2(b1.x + b3.x, b1.y + b3.y);
    const double2 d = make_double2(b1.x - b3.x, b1.y - b3.y);
    const double2 u3 = make_double2(d.y, -d.x);  // (b1 - b3) * -i
    y0 = make_double2(u0.x + u2.x, u0.y + u2.y);
    y2 = make_double2(u0.x - u2.x, u0.y - u2.y);
    y1 = make_double2(u1.x + u3.x, u1.y + u3.y);
    y3 = make_double2(u1.x - u3.x, u1.y - u3.y);
  };
  dft4(t0, t1, t2, t3, a[0], a[2], a[4], a[6]);
  dft4(t4, t5, t6, t7, a[1], a[3], a[5], a[7]);
}
// a[t] *= w1^t, t = 1..7
__device__ __forceinline__ void twiddle8(double2 (&a)[8], double2 w1) {
  const double2 w2 = cmul(w1, w1), w3 = cmul(w2, w1), w4 = cmul(w2, w2);
  const double2 w5 = cmul(w4, w1), w6 = cmul(w3, w3), w7 = cmul(w4, w3);
  a[1] = cmul(a[1], w1); a[2] = cmul(a[2], w2); a[3] = cmul(a[3], w3); a[4] = cmul(a[4], w4);
  a[5] = cmul(a[5], w5); a[6] = cmul(a[6], w6); a[7] = cmul(a[7], w7);
}

// One WARP per analysis window.  The 512-point DFT is three radix-8 Stockham passes with the butterflies in registers
// (two per lane and pass); shared memory only carries the two transposes between passes -- 32 KB of traffic per
// window where the radix-2 version moved 147 KB.  The input is real; the imaginary lanes cost ALU only, and fp64 ALU
// is not what limits this kernel.
__device__ __forceinline__ void mfcc_window(const MfccTables& tb, WarpScratch& w, const int16_t* pcm, int n_valid,
                                            float* out_f32, __half* out_f16) {
  const int lane = threadIdx.x & 31;
  double2 a[2][8];
  // ---- pass 0 (Ns = 1): inputs x[j + 64 t], no twiddles, outputs at 8 j + t
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int n = j + 64 * t;
      double v = 0.0;
      if (n < tb.win_len && n < n_valid) v = (double)((float)pcm[n] * (1.0f / 32768.0f)) * tb.hann[n];
      a[h][t] = make_double2(v, 0.0);
    }
    dft8(a[h]);
#pragma unroll
    for (int t = 0; t < 8; ++t) w.z[zidx(8 * j + t)] = a[h][t];
  }
  __syncwarp();
  // ---- pass 1 (Ns = 8): k = j mod 8, twiddle exp(-2 pi i t k / 64), outputs at (j / 8) * 64 + k + 8 t
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
#pragma unroll
    for (int t = 0; t < 8; ++t) a[h][t] = w.z[zidx(j + 64 * t)];
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    const int k = j & 7;
    twiddle8(a[h], make_double2(tb.tw_re[8 * k], tb.tw_im[8 * k]));
    dft8(a[h]);
    const int base = (j >> 3) * 64 + k;
#pragma unroll
    for (int t = 0; t < 8; ++t) w.z[zidx(base + 8 * t)] = a[h][t];
  }
  __syncwarp();
  // ---- pass 2 (Ns = 64): k = j, twiddle exp(-2 pi i t j / 512), output bin j + 64 t; only bins 0..256 are needed
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
#pragma unroll
    for (int t = 0; t < 8; ++t) a[h][t] = w.z[zidx(j + 64 * t)];
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = lane + 32 * h;
    twiddle8(a[h], make_double2(tb.tw_re[j], tb.tw_im[j]));
    dft8(a[h]);
#pragma unroll
    for (int t = 0; t < 4; ++t) w.pow[j + 64 * t] = (float)(a[h][t].x * a[h][t].x + a[h][t].y * a[h][t].y);
    if (j == 0) w.pow[256] = (float)(a[h][4].x * a[h][4].x + a[h][4].y * a[h][4].y);
  }
  __syncwarp();
  double* spec = reinterpret_cast<double*>(w.z);  // the work array is dead now
  for (int i = lane; i < kBins; i += 32) spec[i] = sqrt((double)w.pow[i]);
  __syncwarp();
  for (int ch = lane; ch < tb.n_channels; ch += 32) {
    // same per-channel accumulation order as the reference's single pass over bins
    double acc = 0.0;
    const int first = tb.chan_first_bin[ch], last = tb.chan_last_bin[ch];
    if (first >= 0) {
      for (int i = first; i <= last; ++i) {
        const double spec_val = spec[i];
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
