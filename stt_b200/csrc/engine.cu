// See engine.h.  Host orchestration + kernel launches; every numeric step cites the reference in the kernel headers.
#include "engine.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "decoder.cuh"
#include "decoder_general.cuh"
#include "gemm_tc.cuh"
#include "gemm2_tc.cuh"
#include "lstm2_tc.cuh"
#ifdef STT_B200_DEV_HOOKS
#include "probe_tc.cuh"
#endif
#include "lstm_tc.cuh"
#include "mfcc.cuh"
#include "scorer_image.h"

namespace stteng {

#define CUDA_OK(expr)                                                                                   \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      fprintf(stderr, "[stt_b200] CUDA error %s at %s:%d: %s\n", #expr, __FILE__, __LINE__,            \
              cudaGetErrorString(_e));                                                                  \
      return -1;                                                                                        \
    }                                                                                                   \
  } while (0)

namespace {

// ------------------------------------------------------------------ small device helpers
__global__ void f32_to_f16_kernel(const float* in, __half* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __float2half_rn(in[i]);
}
__global__ void shift_frames_kernel(__half* feat, int n_keep, int shift, int lanes, int total_frames) {
  // feat[f] = feat[f + shift] for f < n_keep, zero the rest; single block, sequential in f to allow overlap
  for (int f = 0; f < total_frames; ++f) {
    for (int l = threadIdx.x; l < lanes; l += blockDim.x) {
      __half v = (f < n_keep) ? feat[(size_t)(f + shift) * lanes + l] : __float2half_rn(0.f);
      feat[(size_t)f * lanes + l] = v;
    }
    __syncthreads();
  }
}

// Fallback only: materialise the stacked-context windows if the driver refuses the overlapping-row tensor map.
__global__ void gather_windows_kernel(const __half* feat, __half* out, int B, int T, int rows_per_utt, int K1, int K1p) {
  const size_t row = blockIdx.x;  // t * B + b
  const int t = (int)(row / B), b = (int)(row % B);
  const __half* src = feat + ((size_t)b * rows_per_utt + t) * sttmfcc::kFeatLanes;
  for (int k = threadIdx.x; k < K1p; k += blockDim.x) out[row * K1p + k] = (k < K1) ? src[k] : __float2half_rn(0.f);
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

// fp16 tensor map, innermost dimension contiguous, 128B swizzle, box inner = 64 elements.
bool make_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box) {
  auto fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[3], gstride[2];
  cuuint32_t bdim[3], estr[3];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstride[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), gdim, gstride, bdim, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[stt_b200] cuTensorMapEncodeTiled failed: %d\n", (int)r);
    return false;
  }
  return true;
}
bool make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                  uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows}, strides[1] = {row_stride_elems * 2};
  uint32_t box[2] = {64, box_rows};
  return make_tmap(out, base, 2, dims, strides, box);
}

// Kernel attributes (opt-in shared memory) are per DEVICE: each Engine remembers which of its kernels it has configured
// on its device in a bit mask.  Setting an attribute twice is harmless, so concurrent host threads need no lock.
enum KernelBit {
  kBitGemmWin = 0, kBitGemmRelu, kBitGemmBias, kBitGemmSoftmax, kBitGemmSoftmaxWide, kBitDecGeneral, kBitGemm2Relu, kBitGemm2Bias, kBitLstm1, /* +0..3 by cluster size */ kBitLstm2 = kBitLstm1 + 4,
  kBitLstmPair = kBitLstm2 + 4, kBitLstmPP4, kBitLstmPP2, kBitLstmPP1, kBitLstmPP4Mc, kBitLstmPP4Mc4, kBitDec512Solo, kBitDec512, kBitDec512I, kBitDec2048, kBitDec2048I
};
template <class K>
int ensure_smem(std::atomic<uint32_t>* mask, int bit, K kern, int bytes) {
  if (mask->load(std::memory_order_acquire) & (1u << bit)) return 0;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  mask->fetch_or(1u << bit, std::memory_order_release);
  return 0;
}

template <int BN, int EPI, int AMODE>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const sttgemm::GemmParams& p, int num_sms,
                cudaStream_t st, std::atomic<uint32_t>* cfg_mask) {
  constexpr int STAGES = (BN == 256) ? 4 : 6;
  using L = sttgemm::SmemLayout<BN, STAGES>;
  auto kern = sttgemm::gemm_tc_kernel<BN, STAGES, EPI, AMODE>;
  constexpr int bit = (AMODE == sttgemm::kWindows3D) ? kBitGemmWin
                      : (EPI == sttgemm::kEpiClipReluF16) ? kBitGemmRelu
                      : (EPI == sttgemm::kEpiBiasF32)     ? kBitGemmBias
                      : (BN == 32)                        ? kBitGemmSoftmax
                                                          : kBitGemmSoftmaxWide;
  if (ensure_smem(cfg_mask, bit, kern, L::kTotal)) return -1;
  int n_tiles_m;
  if (AMODE == sttgemm::kWindows3D)
    n_tiles_m = ((p.T + p.t_box - 1) / p.t_box) * ((p.B + p.b_box - 1) / p.b_box);
  else
    n_tiles_m = (p.M + sttgemm::BLOCK_M - 1) / sttgemm::BLOCK_M;
  const int n_tiles = n_tiles_m * (p.N / BN);
  const int grid = std::max(1, std::min(n_tiles, num_sms));
  kern<<<grid, sttgemm::kNumThreads, L::kTotal, st>>>(ta, tb, p);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// CTA-pair 256 x 256 tiles (gemm2_tc.cuh) for the layers whose N is a multiple of 256; tb must have 128-row boxes.
template <int EPI>
int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const sttgemm::GemmParams& p, int num_sms, cudaStream_t st,
                 std::atomic<uint32_t>* cfg_mask) {
  constexpr int STAGES = 6;
  using L = sttgemm::Smem2Layout<STAGES>;
  auto kern = sttgemm::gemm2_tc_kernel<STAGES, EPI>;
  if (ensure_smem(cfg_mask, EPI == sttgemm::kEpiClipReluF16 ? kBitGemm2Relu : kBitGemm2Bias, kern, L::kTotal)) return -1;
  const int n_tiles = ((p.M + 255) / 256) * (p.N / 256);
  const int pairs = std::max(1, std::min(n_tiles, num_sms / 2));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(sttgemm::kNumThreads);
  cfg.dynamicSmemBytes = L::kTotal;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return 0;
}

}  // namespace

// ====================================================================================== DeviceScorer
// The enabled scorer as ONE ref-counted device object, like the reference's std::shared_ptr<Scorer>: the model holds the
// current one, and every stream / batch decode captures it when its DecoderState is initialised (stt.cc:542-547), so
// STT_EnableExternalScorer / STT_DisableExternalScorer while a stream is live neither frees tables under a running
// decode nor mixes dictionary states of two scorers.  Alpha / beta live IN the object (Scorer::reset_params), so
// STT_SetScorerAlphaBeta reaches the streams that share it, as in the reference.  Freed when the last holder lets go.
struct DeviceScorer {
  int device = 0;
  uint8_t* blob = nullptr;             // the .scorer file verbatim (+16 B pad)
  size_t blob_bytes = 0;
  sttscorer::ScorerView view{};        // blob = the device pointer; alpha / beta mutable
  uint2* fst_state2 = nullptr;         // per state {first arc, label mask}
  int4* fst_arc4 = nullptr;            // per arc {child dictionary state, its first arc, its label mask, word-ordinal skip} x2
  uint32_t* fst_space_skip = nullptr;  // word-ordinal tables (decoder.cuh DecodeParams), null when not applicable
  uint32_t* ord2wid = nullptr;
  uint2* gstate = nullptr;             // general kernel (decoder_general.cuh): per state {first arc, number of arcs}
  int2* garc = nullptr;                //                                        per arc {label, child dictionary state}
  uint32_t* byte_wid = nullptr;        // UTF-8 scorers: vocabulary id of each one-byte code point [256]
  std::vector<uint8_t> vocab_host;     // the vocabulary-hash section, kept on the host for hot-word id lookups
  ~DeviceScorer() {
    cudaSetDevice(device);
    for (void* p : {(void*)blob, (void*)fst_state2, (void*)fst_arc4, (void*)fst_space_skip, (void*)ord2wid, (void*)gstate, (void*)garc, (void*)byte_wid})
      if (p) cudaFree(p);
  }
};

// ====================================================================================== Engine
struct Engine {
  sttmodel::HostModel hm;
  int device = 0, num_sms = 0;
  int Hp = 0, Cp = 0, K1 = 0;  // padded hidden / cell dims, layer-1 K (= (2c+1)*32)
  int N6 = 32;                 // padded class count of the output layer: 32, or 256 for alphabets beyond 31 labels
  // weights: [N, K] fp16 K-major
  __half *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *wx = nullptr, *wh = nullptr, *w5 = nullptr, *w6 = nullptr;
  float *b1 = nullptr, *b2 = nullptr, *b3 = nullptr, *bx = nullptr, *b5 = nullptr, *b6 = nullptr;
  CUtensorMap tm_w1, tm_w2, tm_w3, tm_wx, tm_wh, tm_w5, tm_w6;
  CUtensorMap tm_w2h, tm_w3h, tm_wxh, tm_w5h;   // the same weights with 128-row boxes (CTA-pair GEMM: each CTA stages half a tile)
  int opt_gemm_pair = 1;
  int opt_lstm_small_pp = 1;   // batches of at least this many (and <= 128) utterances use the pair kernel with ONE group: measured
                               // 4.4-4.9 ms vs 6.0-6.4 ms of the one-CTA kernel at T = 500 (B = 1 .. 128), bit-identical; 0 = never
  // MFCC tables
  sttmfcc::MfccTables tables{};
  std::vector<void*> table_allocs;
  std::shared_ptr<DeviceScorer> scorer;   // the enabled scorer, or null
  // ---- launch state of THIS engine's device (nothing process-wide: one process may drive several GPUs)
  std::atomic<uint32_t> cfg_mask{0};   // KernelBit: kernels whose attributes are set on this device
  std::mutex launch_mu;                // guards the caches below
  std::map<std::pair<int, int>, int> lstm_cluster;   // (M tiles, grid) -> cluster size that is co-resident
  int lstm_pair_usable = -1, lstm_pp_usable[5] = {-1, -1, -1, -1, -1};
  bool use_overlap_view = true;
  bool lstm_noncoop = false;           // see launch_lstm_pp_inst
  // ---- options, read ONCE when the engine is created (development switches; none is needed in production)
  int opt_lstm_cluster_cap = 8, opt_lstm_pair = 1, opt_lstm_pp_mode = 4, opt_word_ordinals = 1;
  int opt_dec_flags = sttdec::kFlagHistSelect;
  int opt_lstm_exact_h = 1;
  bool verbose = false;
};

const sttmodel::HostModel& engine_model(const Engine* e) { return e->hm; }
bool engine_has_scorer(const Engine* e) { return e->scorer != nullptr; }
int engine_num_sms(const Engine* e) { return e->num_sms; }
void engine_set_alpha_beta(Engine* e, float alpha, float beta) {
  if (!e->scorer) return;
  e->scorer->view.alpha = (double)alpha;  // Scorer::reset_params(float, float), scorer.cpp:346-351
  e->scorer->view.beta = (double)beta;
}

namespace {

template <class T>
int upload(T** dst, const std::vector<T>& src, std::vector<void*>* track = nullptr) {
  CUDA_OK(cudaMalloc(reinterpret_cast<void**>(dst), std::max<size_t>(src.size(), 1) * sizeof(T)));
  CUDA_OK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  if (track) track->push_back(*dst);
  return 0;
}

// W[in, out] fp32 (TF layout) -> [Np, Kp] fp16 K-major, zero padded; optional row permutation of the output dim.
std::vector<__half> to_nk_f16(const float* W, int n_in, int n_out, int Kp, int Np) {
  std::vector<__half> o((size_t)Np * Kp, __float2half_rn(0.f));
  for (int k = 0; k < n_in; ++k)
    for (int n = 0; n < n_out; ++n) o[(size_t)n * Kp + k] = __float2half_rn(W[(size_t)k * n_out + n]);
  return o;
}

int build_mfcc_tables(Engine* e) {
  const auto& m = e->hm;
  if ((int)m.win_len > sttmfcc::kFft) return -1;
  // fft length = NextPowerOfTwo(window) (spectrogram.cc:92): the kernel is specialised for 512
  int fft = 1;
  while (fft < (int)m.win_len) fft <<= 1;
  if (fft != sttmfcc::kFft) return -1;
  const int n_bins = sttmfcc::kBins, n_ch = 40, n_dct = (int)m.n_input;
  const double pi = atan(1.0) * 4.0;
  std::vector<double> hann(m.win_len), tw_re(256), tw_im(256);
  for (uint32_t i = 0; i < m.win_len; ++i) hann[i] = 0.5 - 0.5 * cos((2.0 * pi * i) / m.win_len);
  for (int k = 0; k < 256; ++k) {
    const double ang = -2.0 * pi * k / 512.0;
    tw_re[k] = cos(ang);
    tw_im[k] = sin(ang);
  }
  // MfccMelFilterbank::Initialize (mfcc_mel_filterbank.cc:40-167): lower 20 Hz, upper sample_rate/2, 40 channels
  const double lower = 20.0, upper = m.sample_rate / 2, sr = m.sample_rate;
  auto mel = [](double f) { return 1127.0 * log1p(f / 700.0); };
  std::vector<double> center(n_ch + 1), weights(n_bins);
  std::vector<int> mapper(n_bins), first(n_ch, -1), last(n_ch, -1);
  const double mel_low = mel(lower), mel_hi = mel(upper), spacing = (mel_hi - mel_low) / (double)(n_ch + 1);
  for (int i = 0; i < n_ch + 1; ++i) center[i] = mel_low + spacing * (i + 1);
  const double hz_per_sbin = 0.5 * sr / (double)(n_bins - 1);
  const int start_index = (int)(1.5 + lower / hz_per_sbin), end_index = (int)(upper / hz_per_sbin);
  int channel = 0;
  for (int i = 0; i < n_bins; ++i) {
    const double melf = mel(i * hz_per_sbin);
    if (i < start_index || i > end_index) {
      mapper[i] = -2;
    } else {
      while (channel < n_ch && center[channel] < melf) ++channel;
      mapper[i] = channel - 1;
    }
  }
  for (int i = 0; i < n_bins; ++i) {
    channel = mapper[i];
    if (i < start_index || i > end_index) weights[i] = 0.0;
    else if (channel >= 0)
      weights[i] = (center[channel + 1] - mel(i * hz_per_sbin)) / (center[channel + 1] - center[channel]);
    else
      weights[i] = (center[0] - mel(i * hz_per_sbin)) / (center[0] - mel_low);
  }
  for (int i = start_index; i <= end_index && i < n_bins; ++i) {
    for (int ch : {mapper[i], mapper[i] + 1}) {
      if (ch < 0 || ch >= n_ch) continue;
      if (first[ch] < 0) first[ch] = i;
      last[ch] = i;
    }
  }
  std::vector<double> cosines((size_t)n_dct * n_ch);
  const double fnorm = sqrt(2.0 / n_ch), arg = pi / n_ch;
  for (int i = 0; i < n_dct; ++i)
    for (int j = 0; j < n_ch; ++j) cosines[(size_t)i * n_ch + j] = fnorm * cos(i * arg * (j + 0.5));

  double *d_hann, *d_re, *d_im, *d_w, *d_cos;
  int *d_map, *d_first, *d_last;
  if (upload(&d_hann, hann, &e->table_allocs) || upload(&d_re, tw_re, &e->table_allocs) ||
      upload(&d_im, tw_im, &e->table_allocs) || upload(&d_w, weights, &e->table_allocs) ||
      upload(&d_cos, cosines, &e->table_allocs) || upload(&d_map, mapper, &e->table_allocs) ||
      upload(&d_first, first, &e->table_allocs) || upload(&d_last, last, &e->table_allocs))
    return -1;
  e->tables = sttmfcc::MfccTables{d_hann, d_re,    d_im,     d_w,          d_map,       d_first,
                                  d_last, d_cos,   (int)m.win_len, (int)m.win_step, n_ch, n_dct,
                                  start_index, end_index};
  return 0;
}

int build_weights(Engine* e) {
  const auto& m = e->hm;
  const int H = m.n_hidden, C = m.n_cell, K = m.n_classes, ni = m.n_input, nf = 2 * m.n_context + 1;
  e->Hp = round_up(H, 256);
  e->Cp = round_up(C, 64);
  e->K1 = nf * sttmfcc::kFeatLanes;
  const int Hp = e->Hp, Cp = e->Cp, K1p = round_up(e->K1, 64);
  // layer 1: K index f*32 + i  <-  row (f*n_input + i) of w1
  std::vector<__half> w1((size_t)Hp * K1p, __float2half_rn(0.f));
  for (int f = 0; f < nf; ++f)
    for (int i = 0; i < ni; ++i)
      for (int n = 0; n < H; ++n)
        w1[(size_t)n * K1p + f * sttmfcc::kFeatLanes + i] = __float2half_rn(m.w1[(size_t)(f * ni + i) * H + n]);
  auto pad_bias = [](const std::vector<float>& b, int Np) {
    std::vector<float> o(Np, 0.f);
    std::copy(b.begin(), b.end(), o.begin());
    return o;
  };
  std::vector<__half> w2 = to_nk_f16(m.w2.data(), H, H, Hp, Hp), w3 = to_nk_f16(m.w3.data(), H, H, Hp, Hp);
  std::vector<__half> w5 = to_nk_f16(m.w5.data(), C, H, Cp, Hp);
  e->N6 = (K <= 32) ? 32 : 256;
  std::vector<__half> w6 = to_nk_f16(m.w6.data(), H, K, Hp, e->N6);
  // LSTM kernel [H + C, 4C], gate blocks i|j|f|o (rnn_cell_impl.py:1060-1061) -> gate-interleaved rows cell*4 + g
  std::vector<__half> wx((size_t)4 * Cp * Hp, __float2half_rn(0.f)), wh((size_t)4 * Cp * Cp, __float2half_rn(0.f));
  std::vector<float> bx((size_t)4 * Cp, 0.f);
  for (int cell = 0; cell < C; ++cell)
    for (int g = 0; g < 4; ++g) {
      const size_t row = (size_t)cell * 4 + g;
      const size_t col = (size_t)g * C + cell;
      for (int k = 0; k < H; ++k) wx[row * Hp + k] = __float2half_rn(m.lstm_kernel[(size_t)k * 4 * C + col]);
      for (int k = 0; k < C; ++k) wh[row * Cp + k] = __float2half_rn(m.lstm_kernel[(size_t)(H + k) * 4 * C + col]);
      bx[row] = m.lstm_bias[col];
    }
  if (upload(&e->w1, w1) || upload(&e->w2, w2) || upload(&e->w3, w3) || upload(&e->wx, wx) || upload(&e->wh, wh) ||
      upload(&e->w5, w5) || upload(&e->w6, w6))
    return -1;
  std::vector<float> b1 = pad_bias(m.b1, Hp), b2 = pad_bias(m.b2, Hp), b3 = pad_bias(m.b3, Hp), b5 = pad_bias(m.b5, Hp),
                     b6 = pad_bias(m.b6, e->N6);
  if (upload(&e->b1, b1) || upload(&e->b2, b2) || upload(&e->b3, b3) || upload(&e->bx, bx) || upload(&e->b5, b5) ||
      upload(&e->b6, b6))
    return -1;
  bool ok = make_tmap_2d(&e->tm_w1, e->w1, Hp, K1p, K1p, 256) && make_tmap_2d(&e->tm_w2, e->w2, Hp, Hp, Hp, 256) &&
            make_tmap_2d(&e->tm_w3, e->w3, Hp, Hp, Hp, 256) && make_tmap_2d(&e->tm_wx, e->wx, 4 * Cp, Hp, Hp, 256) &&
            make_tmap_2d(&e->tm_wh, e->wh, 4 * Cp, Cp, Cp, 64) && make_tmap_2d(&e->tm_w5, e->w5, Hp, Cp, Cp, 256) &&
            make_tmap_2d(&e->tm_w6, e->w6, e->N6, Hp, Hp, e->N6) &&
            make_tmap_2d(&e->tm_w2h, e->w2, Hp, Hp, Hp, 128) && make_tmap_2d(&e->tm_w3h, e->w3, Hp, Hp, Hp, 128) &&
            make_tmap_2d(&e->tm_wxh, e->wx, 4 * Cp, Hp, Hp, 128) && make_tmap_2d(&e->tm_w5h, e->w5, Hp, Cp, Cp, 128);
  return ok ? 0 : -1;
}

}  // namespace

Engine* engine_create(const sttmodel::HostModel& m, std::string* err) {
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    if (err) *err = "no CUDA device: this library has no CPU path";
    return nullptr;
  }
  Engine* e = new Engine();
  e->hm = m;
  cudaGetDevice(&e->device);
  {
    auto geti = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
    e->opt_lstm_cluster_cap = std::max(1, geti("STT_B200_LSTM_CLUSTER", 8));
    e->opt_lstm_pair = geti("STT_B200_LSTM_PAIR", 1);
    e->opt_lstm_pp_mode = geti("STT_B200_LSTM_PINGPONG", 4);
    e->opt_word_ordinals = getenv("STT_B200_NO_WORD_ORDINALS") ? 0 : 1;
    e->opt_dec_flags = geti("STT_B200_DEC_FLAGS", e->opt_dec_flags);
    e->opt_lstm_exact_h = geti("STT_B200_LSTM_EXACT_H", 1);
    e->opt_gemm_pair = geti("STT_B200_GEMM_PAIR", 1);
    e->opt_lstm_small_pp = geti("STT_B200_LSTM_SMALL_PP", 1);
    e->verbose = getenv("STT_B200_VERBOSE") != nullptr;
    // A CUDA injection library (Nsight Compute / Systems) serialises kernels and, with this driver, fails launches that
    // carry BOTH the cooperative and the cluster attribute; see launch_lstm_*.
    e->lstm_noncoop = getenv("STT_B200_LSTM_NONCOOP") != nullptr || getenv("CUDA_INJECTION64_PATH") != nullptr ||
                      getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") != nullptr;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, e->device);
  e->num_sms = prop.multiProcessorCount;
  if (prop.major < 10) {
    if (err) *err = "sm_100a (B200) required";
    delete e;
    return nullptr;
  }
  if (m.n_classes > 256 || m.n_classes < 2) {
    if (err) *err = "alphabets larger than 255 labels are not supported";
    delete e;
    return nullptr;
  }
  if (round_up(m.n_cell, 64) / sttlstm::kCellsPerCta > e->num_sms) {
    if (err) *err = "n_cell too large for the single-wave LSTM kernel";
    delete e;
    return nullptr;
  }
  if (build_mfcc_tables(e) != 0 || build_weights(e) != 0) {
    if (err) *err = "failed to build device tables / weights";
    delete e;
    return nullptr;
  }
  return e;
}

void engine_clear_scorer(Engine* e) { e->scorer.reset(); }   // live streams keep the object they captured

void engine_destroy(Engine* e) {
  if (!e) return;
  engine_clear_scorer(e);
  for (void* p : e->table_allocs) cudaFree(p);
  for (void* p : {(void*)e->w1, (void*)e->w2, (void*)e->w3, (void*)e->wx, (void*)e->wh, (void*)e->w5, (void*)e->w6,
                  (void*)e->b1, (void*)e->b2, (void*)e->b3, (void*)e->bx, (void*)e->b5, (void*)e->b6})
    if (p) cudaFree(p);
  delete e;
}

// Word ordinals of the dictionary FST (decoder.cuh DecodeParams::fst_arc_skip).  The reference builds the dictionary by
// adding every vocabulary word + the space label as a path and minimising (scorer.cpp:398-440,
// decoder_utils.cpp add_word_to_dictionary): an acyclic deterministic acceptor.  With paths(q) = number of accepting
// paths out of q, arc_skip(q, a) = [q final] + sum of paths(target) over q's arcs with a smaller label, and the sum of
// arc_skip along a word's arcs is its rank among the FST's words in label order.  ord2wid maps that rank to the KenLM
// vocabulary id of the word's bytes (the same vocab_index the walking path calls).  Anything unexpected -- a cycle, a
// word that does not end with the space label, more than 2^31 words -- leaves the tables null (walking path).
void build_word_ordinals(const Engine* eng, DeviceScorer* e, const sttscorer::ScorerView& v, const uint8_t* bytes,
                         std::vector<uint32_t>* skip_out, std::vector<uint32_t>* space_skip_out) {
  const int64_t nS = v.fst_nstates, nA = v.fst_narcs;
  if (nS <= 0 || nA <= 0 || v.fst_start < 0 || v.fst_start >= nS) return;
  struct Arc { int32_t il, nx; };
  auto state_arcs = [&](int64_t q, uint32_t* pos, uint32_t* narcs) {
    const uint8_t* srec = bytes + v.fst_states_off + (uint64_t)q * 20;
    memcpy(pos, srec + 4, 4);
    memcpy(narcs, srec + 8, 4);
  };
  auto arc_at = [&](uint64_t i) {
    Arc a;
    const uint8_t* arc = bytes + v.fst_arcs_off + i * 16;
    memcpy(&a.il, arc, 4);
    memcpy(&a.nx, arc + 12, 4);
    return a;
  };
  std::vector<uint64_t> paths((size_t)nS, 0);
  std::vector<uint8_t> color((size_t)nS, 0);  // 0 new, 1 on the stack, 2 done
  std::vector<uint32_t> skip((size_t)nA, 0), space_skip((size_t)nS, 0xffffffffu);
  // iterative post-order DFS from the start state
  struct Frame { int64_t q; uint32_t pos, narcs, a; uint64_t acc; };
  std::vector<Frame> stack;
  auto push = [&](int64_t q) {
    Frame f; f.q = q; f.a = 0;
    state_arcs(q, &f.pos, &f.narcs);
    f.acc = sttscorer::fst_is_final(v, q) ? 1 : 0;
    color[(size_t)q] = 1;
    stack.push_back(f);
  };
  push(v.fst_start);
  while (!stack.empty()) {
    Frame& f = stack.back();
    if (f.a < f.narcs) {
      const Arc a = arc_at((uint64_t)f.pos + f.a);
      if (a.nx < 0 || a.nx >= nS || (uint64_t)f.pos + f.a >= (uint64_t)nA) return;
      if (color[(size_t)a.nx] == 1) return;  // cycle
      if (color[(size_t)a.nx] == 0) { push(a.nx); continue; }
      skip[(size_t)f.pos + f.a] = (uint32_t)f.acc;
      f.acc += paths[(size_t)a.nx];
      if (f.acc >= (1ull << 31)) return;
      ++f.a;
    } else {
      paths[(size_t)f.q] = f.acc;
      color[(size_t)f.q] = 2;
      stack.pop_back();
    }
  }
  const uint64_t n_words = paths[(size_t)v.fst_start];
  if (n_words == 0) return;
  // enumerate the words: ordinal -> vocabulary id; every word must end with the space label on an arc into a final state
  std::vector<uint32_t> o2w((size_t)n_words, 0);
  struct EFrame { int64_t q; uint32_t pos, narcs, a; uint32_t ord; size_t len; };
  std::vector<EFrame> es;
  std::vector<uint8_t> word;
  auto epush = [&](int64_t q, uint32_t ord) {
    EFrame f; f.q = q; f.a = 0; f.ord = ord; f.len = word.size();
    state_arcs(q, &f.pos, &f.narcs);
    es.push_back(f);
  };
  epush(v.fst_start, 0);
  uint64_t seen = 0;
  while (!es.empty()) {
    EFrame& f = es.back();
    if (f.a == 0 && f.q != v.fst_start && sttscorer::fst_is_final(v, f.q) && f.narcs != 0) return;  // words nest past a final state
    if (f.a < f.narcs) {
      const uint64_t ai = (uint64_t)f.pos + f.a;
      const Arc a = arc_at(ai);
      ++f.a;
      word.resize(f.len);
      const int label = a.il - 1;
      if (label < 0 || label >= 256) return;
      const uint32_t ord = f.ord + skip[(size_t)ai];
      if (sttscorer::fst_is_final(v, a.nx)) {
        if ((uint32_t)label != v.space_label) return;
        uint32_t tp, tn;
        state_arcs(a.nx, &tp, &tn);
        if (tn != 0 || ord >= n_words) return;
        o2w[ord] = sttscorer::vocab_index(v, word.data(), (uint32_t)word.size());
        space_skip[(size_t)f.q] = skip[(size_t)ai];
        ++seen;
      } else {
        if ((uint32_t)label == v.space_label) return;
        for (int b = 0; b < v.label_len[label]; ++b) word.push_back(v.label_bytes[label][b]);
        if (word.size() > 4096) return;
        epush(a.nx, ord);
      }
    } else {
      es.pop_back();
    }
  }
  if (seen != n_words) return;
  if (cudaMalloc(reinterpret_cast<void**>(&e->fst_space_skip), space_skip.size() * 4) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&e->ord2wid), o2w.size() * 4) != cudaSuccess) {
    cudaGetLastError();
    if (e->fst_space_skip) cudaFree(e->fst_space_skip);
    if (e->ord2wid) cudaFree(e->ord2wid);
    e->fst_space_skip = e->ord2wid = nullptr;
    return;
  }
  cudaMemcpy(e->fst_space_skip, space_skip.data(), space_skip.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(e->ord2wid, o2w.data(), o2w.size() * 4, cudaMemcpyHostToDevice);
  *skip_out = std::move(skip);
  *space_skip_out = std::move(space_skip);
  if (eng->verbose) fprintf(stderr, "[stt_b200] dictionary word ordinals: %llu words\n", (unsigned long long)n_words);
}

int engine_set_scorer(Engine* e, const uint8_t* bytes, size_t n) {
  cudaSetDevice(e->device);  // CUDA's current device is per host thread
  sttscorer::AlphabetBytes ab;
  ab.labels = e->hm.labels;
  ab.space_label = e->hm.space_label;
  sttscorer::ScorerView v;
  int err = sttscorer::parse_scorer(bytes, n, ab, &v);
  if (err) return err;
  // Everything is parsed, validated and uploaded into a NEW object; the engine's current scorer is replaced only when
  // all of it has succeeded (a failed STT_EnableExternalScorer leaves the old scorer enabled, stt.cc:428-432).
  std::shared_ptr<DeviceScorer> ds = std::make_shared<DeviceScorer>();
  ds->device = e->device;
  if (v.fst_nstates <= 0 || v.fst_start < 0 || v.fst_start >= v.fst_nstates || v.fst_narcs < 0) return sttscorer::SCORER_INVALID_TRIE;
  // Pre-digest the dictionary FST for the decoder (semantics of PathTrie::get_path_trie, path_trie.cpp:60-88):
  // per state the set of labels with an outgoing arc, per arc the child's dictionary state, i.e. Start() when the
  // arc's target is final ("restart spell checker at the start state"), else the target.
  v.blob = bytes;  // host view for the preprocessing
  std::vector<uint2> st2((size_t)v.fst_nstates), gst((size_t)v.fst_nstates);
  std::vector<int2> ar2((size_t)v.fst_narcs);
  for (int64_t q = 0; q < v.fst_nstates; ++q) {
    const uint8_t* srec = bytes + v.fst_states_off + (uint64_t)q * 20;
    uint32_t pos, narcs;
    memcpy(&pos, srec + 4, 4);
    memcpy(&narcs, srec + 8, 4);
    if ((uint64_t)pos + narcs > (uint64_t)v.fst_narcs) return sttscorer::SCORER_INVALID_TRIE;  // before any arc is read
    uint32_t mask = 0;
    for (uint32_t a = 0; a < narcs; ++a) {
      const uint8_t* arc = bytes + v.fst_arcs_off + (uint64_t)(pos + a) * 16;
      int32_t il, nx;
      memcpy(&il, arc, 4);
      memcpy(&nx, arc + 12, 4);
      if (il >= 1 && il <= 32) mask |= 1u << (il - 1);
      if (a > 0) {
        int32_t prev;
        memcpy(&prev, arc - 16, 4);
        if (prev >= il) return sttscorer::SCORER_INVALID_TRIE;  // SortedMatcher needs ilabel-sorted, deterministic arcs
      }
      if (nx < 0 || nx >= v.fst_nstates) return sttscorer::SCORER_INVALID_TRIE;
      const bool fin = sttscorer::fst_is_final(v, nx);
      ar2[pos + a] = make_int2(il, fin ? (int32_t)v.fst_start : nx);
    }
    st2[(size_t)q] = make_uint2(pos, mask);
    gst[(size_t)q] = make_uint2(pos, narcs);
  }
  std::vector<int2> gar(ar2.size());
  for (size_t i = 0; i < ar2.size(); ++i) gar[i] = make_int2(ar2[i].x - 1, ar2[i].y);   // label = ilabel - 1 (path_trie.cpp:62)
  std::vector<uint32_t> skip, space_skip;
  build_word_ordinals(e, ds.get(), v, bytes, &skip, &space_skip);
  // two 16-byte words per arc: {child state, child's first arc, child's label mask, ordinal skip} and
  // {ordinal skip of the CHILD state's space arc (or none), -, -, -}
  std::vector<int4> ar4(2 * ar2.size());
  for (size_t i = 0; i < ar2.size(); ++i) {
    const uint2 cs = st2[(size_t)ar2[i].y];
    ar4[2 * i] = make_int4(ar2[i].y, (int)cs.x, (int)cs.y, skip.empty() ? 0 : (int)skip[i]);
    ar4[2 * i + 1] = make_int4(space_skip.empty() ? -1 : (int)space_skip[(size_t)ar2[i].y], 0, 0, 0);
  }
  if (cudaMalloc(reinterpret_cast<void**>(&ds->blob), n + 16) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&ds->fst_state2), std::max<size_t>(st2.size(), 1) * sizeof(uint2)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&ds->fst_arc4), std::max<size_t>(ar4.size(), 1) * sizeof(int4)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&ds->gstate), std::max<size_t>(gst.size(), 1) * sizeof(uint2)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&ds->garc), std::max<size_t>(gar.size(), 1) * sizeof(int2)) != cudaSuccess) {
    cudaGetLastError();
    return sttscorer::SCORER_UNREADABLE;
  }
  if (cudaMemset(ds->blob + n, 0, 16) != cudaSuccess || cudaMemcpy(ds->blob, bytes, n, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(ds->fst_state2, st2.data(), st2.size() * sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(ds->fst_arc4, ar4.data(), ar4.size() * sizeof(int4), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(ds->gstate, gst.data(), gst.size() * sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(ds->garc, gar.data(), gar.size() * sizeof(int2), cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    return sttscorer::SCORER_UNREADABLE;
  }
  if (v.is_utf8) {
    // one-byte code points (ASCII): vocabulary ids looked up once; v.blob still addresses the host copy here
    std::vector<uint32_t> bw(256, 0);
    for (int bt = 1; bt < 128; ++bt) {
      const uint8_t ch = (uint8_t)bt;
      bw[bt] = sttscorer::vocab_index(v, &ch, 1);
    }
    if (cudaMalloc(reinterpret_cast<void**>(&ds->byte_wid), 256 * 4) != cudaSuccess ||
        cudaMemcpy(ds->byte_wid, bw.data(), 256 * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      cudaGetLastError();
      return sttscorer::SCORER_UNREADABLE;
    }
  }
  ds->blob_bytes = n;
  const size_t vocab_end = std::min<size_t>(n, (size_t)(v.probing ? v.pvocab_off + v.pvocab_buckets * 12 + 16
                                                                  : v.vocab_off + v.vocab_count * 8 + 16));
  ds->vocab_host.assign(bytes, bytes + vocab_end);
  ds->vocab_host.resize(vocab_end + 16, 0);
  v.blob = ds->blob;
  ds->view = v;
  e->scorer = std::move(ds);   // the previous object lives on in the streams that captured it
  return 0;
}

// ====================================================================================== Batch
struct Batch {
  Engine* e = nullptr;
  int B_cap = 0, S_cap = 0, T_cap = 0, beam_cap = 0, dec_T_cap = 0, max_results = 0;
  int out_tok_cap = 0;         // tokens per result the output blocks hold (a stream's transcript may outgrow its arena)
  int B = 0, T_max = 0;
  uint32_t ht_gen = 0;  // generation of the decoder hash tables (decoder_reset)
  std::vector<int> T;  // timesteps per utterance
  cudaStream_t st = nullptr;
  cudaEvent_t ev[12];
  // host staging (pinned)
  int16_t* h_pcm = nullptr;
  int* h_nsamples = nullptr;
  // device
  int16_t* d_pcm = nullptr;
  int* d_nsamples = nullptr;
  int rows_per_utt = 0;        // T_cap + 2*n_context (+ pad)
  __half* d_feat = nullptr;    // [B_cap, rows_per_utt, 32]
  float* d_feat32 = nullptr;   // [B_cap, T_cap, n_input]
  __half *d_act_a = nullptr, *d_act_b = nullptr;  // [T_cap*B_cap, Hp]
  float* d_xw = nullptr;       // [T_cap*B_cap, 4*Cp]
  __half* d_hall = nullptr;    // [(T_cap+1)*B_cap, Cp]
  float *d_c = nullptr, *d_h = nullptr;  // [B_cap, Cp]
  unsigned int* d_barrier = nullptr;
  unsigned long long* d_lstm_prof = nullptr;  // [256 * 4]
  float* d_probs = nullptr;    // [B_cap, T_cap, n_classes]
  double* d_probs64 = nullptr; // optional f64 copy set by batch_set_probs64 (Python decoder API)
  bool use_probs64 = false;
  CUtensorMap tm_act_a, tm_act_b, tm_hall, tm_hall_out;
  __half* d_winmat = nullptr;  // fallback window matrix [T_cap*B_cap + 128, K1p]
  CUtensorMap tm_winmat;
  // decoder
  sttdec::Slot* d_slots = nullptr;
  std::vector<sttdec::Slot> h_slots;
  uint8_t* d_slot_mem = nullptr;
  size_t slot_bytes = 0;
  sttdec::StepInput* d_inputs = nullptr;
  sttdec::FinalOut* d_finals = nullptr;
  uint8_t* d_out_mem = nullptr;
  uint8_t* h_out_mem = nullptr;  // pinned mirror
  size_t out_bytes_per_utt = 0;
  int cur_beam = 0, cur_results = 1;
  std::vector<uint32_t> hot_ids;   // hot words of the current decode / stream (vocabulary ids)
  std::vector<float> hot_boosts;
  // streaming
  int stream_frames = 0;  // frames currently in d_feat (utterance 0)
  int16_t* d_win = nullptr;
  sttmfcc::FrameJob* d_jobs = nullptr;
  int last_run_T = 0;
  uint32_t stream_arena_count = 1, stream_ts_count = 1;   // host mirror of the stream slot's arena / timestep-tree fill
  long long stream_compactions = 0;
  StageTimes times;
  long long launches = 0;
  std::shared_ptr<DeviceScorer> scorer;   // captured by decoder_reset: the scorer this decode / stream runs with
  bool instrument = false;   // decoder statistics build (per-phase clocks, LM counters); bench.py asks for it once
  // vocabulary pruning of the decoder (get_pruned_emissions); the C API's fixed values (stt.cc:539-540) unless the
  // decoder-only Python surface sets others
  double cutoff_prob = 1.0;
  int cutoff_top_n = 40;
  uint32_t* d_gen_scratch = nullptr;   // general kernel's per-utterance work arrays, allocated on first use
};

const StageTimes& batch_times(const Batch* b) { return b->times; }
long long batch_kernel_launches(const Batch* b) { return b->launches; }
void batch_set_instrumented(Batch* b, bool on) { b->instrument = on; }
int batch_T(const Batch* b, int utt) { return (utt >= 0 && utt < b->B) ? b->T[utt] : -1; }

namespace {

int frames_for(int n, int win_len, int win_step) { return (n >= win_len ? (n - win_len) / win_step + 1 : 0) + 1; }

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int alloc_slots(Batch* b) {
  const int W = b->beam_cap, C = b->e->hm.n_classes;
  const uint32_t arena_cap = 1u + (uint32_t)W * (uint32_t)b->dec_T_cap;
  const uint32_t ts_cap = 1u + (uint32_t)W * (uint32_t)(b->dec_T_cap + 1);
  const uint32_t cand_cap = (uint32_t)W * (uint32_t)C;
  // carve one slab per slot
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return o;
  };
  const int SW = sttdec::kStateWords;
  const size_t o_nodes = take(sizeof(sttdec::Node) * (size_t)arena_cap);
  if (arena_cap >= (1u << 24)) return -1;  // hash entries hold 24-bit node ids
  uint32_t ht = 1;
  while (ht < 2u * arena_cap) ht <<= 1;
  const size_t o_ht = take(8ull * ht);
  const size_t o_lmc = take(8ull * arena_cap), o_lmsw = take(4ull * arena_cap * SW), o_lmsb = take(4ull * arena_cap * SW);
  const size_t o_lmm = take(4ull * arena_cap);
  const size_t o_tstree = take(8ull * ts_cap);
  const size_t o_sc = take(4ull * W), o_bp = take(4ull * W), o_nb = take(4ull * W), o_nd = take(4ull * W), o_lts = take(4ull * W);
  const size_t o_ck = take(8ull * cand_cap), o_p0 = take(4ull * cand_cap), o_p1 = take(4ull * cand_cap);
  const size_t o_scal = take(64), o_ph = take(64), o_aux = take(W > 512 ? 16ull * 2048 : 256);  // StepSmem::aux of the wide instantiation (WC = 2048)
  b->slot_bytes = off;
  CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b->d_slot_mem), b->slot_bytes * b->B_cap));
  CUDA_OK(cudaMemset(b->d_slot_mem, 0, b->slot_bytes * b->B_cap));  // hash tables start at generation 0 = empty
  b->h_slots.resize(b->B_cap);
  for (int u = 0; u < b->B_cap; ++u) {
    uint8_t* base = b->d_slot_mem + b->slot_bytes * u;
    sttdec::Slot& s = b->h_slots[u];
    s.nodes = (sttdec::Node*)(base + o_nodes);
    s.ht = (unsigned long long*)(base + o_ht);
    s.ht_mask = ht - 1;
    s.ht_gen = 0;
    s.lm_cond = (double*)(base + o_lmc); s.lm_sw = (uint32_t*)(base + o_lmsw); s.lm_sb = (float*)(base + o_lmsb);
    s.lm_meta = (uint32_t*)(base + o_lmm);
    s.ts_tree = (uint2*)(base + o_tstree);
    s.score = (float*)(base + o_sc); s.b_prev = (float*)(base + o_bp); s.nb_prev = (float*)(base + o_nb);
    s.node = (uint32_t*)(base + o_nd); s.ts = (uint32_t*)(base + o_lts);
    s.c_key = (unsigned long long*)(base + o_ck); s.c_p0 = (uint32_t*)(base + o_p0); s.c_p1 = (uint32_t*)(base + o_p1);
    s.aux = (uint32_t*)(base + o_aux);
    s.scalars = (uint32_t*)(base + o_scal); s.phase_cycles = (unsigned long long*)(base + o_ph);
    s.arena_cap = arena_cap; s.ts_cap = ts_cap; s.beam_cap = W; s.cand_cap = cand_cap;
  }
  CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b->d_slots), sizeof(sttdec::Slot) * b->B_cap));
  CUDA_OK(cudaMemcpy(b->d_slots, b->h_slots.data(), sizeof(sttdec::Slot) * b->B_cap, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b->d_inputs), sizeof(sttdec::StepInput) * b->B_cap));
  // outputs: per utterance [n_results i32][pad][confidence f64 x R][n_tokens i32 x R][tokens u32 x R*Tm][timesteps ...]
  b->out_tok_cap = b->dec_T_cap;
  const int R = b->max_results, Tm = b->out_tok_cap;
  size_t per = 0;
  per = align_up(per + 8, 8);
  const size_t o_conf = per; per += 8ull * R;
  const size_t o_nt = per; per += 4ull * R;
  per = align_up(per, 8);
  const size_t o_tok = per; per += 4ull * R * Tm;
  const size_t o_ts = per; per += 4ull * R * Tm;
  per = align_up(per, 256);
  b->out_bytes_per_utt = per;
  CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b->d_out_mem), per * b->B_cap));
  CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&b->h_out_mem), per * b->B_cap));
  std::vector<sttdec::FinalOut> fo(b->B_cap);
  for (int u = 0; u < b->B_cap; ++u) {
    uint8_t* base = b->d_out_mem + per * u;
    fo[u].max_results = R; fo[u].max_tokens = Tm;
    fo[u].n_results = (int*)base; fo[u].confidence = (double*)(base + o_conf); fo[u].n_tokens = (int*)(base + o_nt);
    fo[u].tokens = (uint32_t*)(base + o_tok); fo[u].timesteps = (uint32_t*)(base + o_ts);
  }
  CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b->d_finals), sizeof(sttdec::FinalOut) * b->B_cap));
  CUDA_OK(cudaMemcpy(b->d_finals, fo.data(), sizeof(sttdec::FinalOut) * b->B_cap, cudaMemcpyHostToDevice));
  return 0;
}

}  // namespace

Batch* batch_create(Engine* e, int B_cap, int max_samples, int beam_cap, int dec_T_cap, std::string* err) {
  cudaSetDevice(e->device);  // CUDA's current device is per host thread
  const auto& m = e->hm;
  Batch* b = new Batch();
  b->e = e;
  b->B_cap = B_cap;
  b->S_cap = std::max(max_samples, (int)m.win_len);
  b->T_cap = frames_for(b->S_cap, m.win_len, m.win_step);
  b->beam_cap = beam_cap;
  b->dec_T_cap = std::max(dec_T_cap, b->T_cap);
  b->max_results = 1;
  b->T.assign(B_cap, 0);
  auto fail = [&](const char* what) -> Batch* {
    if (err) *err = what;
    batch_destroy(b);
    return nullptr;
  };
  if (B_cap > 256) return fail("B_cap > 256: split the batch (one LSTM wave holds 256 rows)");
  if (cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking) != cudaSuccess) return fail("stream");
  for (auto& ev : b->ev) cudaEventCreate(&ev);
  const int Hp = e->Hp, Cp = e->Cp, C = m.n_classes;
  b->rows_per_utt = b->T_cap + 2 * m.n_context + 1;
  const size_t M_cap = (size_t)b->T_cap * B_cap;
  bool ok = true;
  ok &= cudaMallocHost((void**)&b->h_pcm, (size_t)B_cap * b->S_cap * 2) == cudaSuccess;
  ok &= cudaMallocHost((void**)&b->h_nsamples, 4ull * B_cap) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_pcm, (size_t)B_cap * b->S_cap * 2 + 1024) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_nsamples, 4ull * B_cap) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_feat, (size_t)B_cap * b->rows_per_utt * 64 + 4096) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_feat32, (size_t)B_cap * b->T_cap * m.n_input * 4) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_act_a, (M_cap + 128) * Hp * 2) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_act_b, (M_cap + 128) * Hp * 2) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_xw, (M_cap + 128) * 4ull * Cp * 4) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_hall, (M_cap + B_cap + 256) * Cp * 2) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_c, (size_t)B_cap * Cp * 4) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_h, (size_t)B_cap * Cp * 4) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_barrier, 16384) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_lstm_prof, 8 * 4 * 1024) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_probs, (size_t)B_cap * b->T_cap * C * 4) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_win, (size_t)(b->T_cap + 1) * m.win_len * 2) == cudaSuccess;
  ok &= cudaMalloc((void**)&b->d_jobs, sizeof(sttmfcc::FrameJob) * (b->T_cap + 1)) == cudaSuccess;
  if (!ok) return fail("device allocation failed");
  cudaMemset(b->d_feat, 0, (size_t)B_cap * b->rows_per_utt * 64 + 4096);
  cudaMemset(b->d_hall, 0, (M_cap + B_cap + 256) * Cp * 2);
  cudaMemset(b->d_c, 0, (size_t)B_cap * Cp * 4);
  cudaMemset(b->d_h, 0, (size_t)B_cap * Cp * 4);
  ok &= make_tmap_2d(&b->tm_act_a, b->d_act_a, M_cap + 128, Hp, Hp, 128);
  ok &= make_tmap_2d(&b->tm_act_b, b->d_act_b, M_cap + 128, Hp, Hp, 128);
  ok &= make_tmap_2d(&b->tm_hall, b->d_hall, M_cap + B_cap + 256, Cp, Cp, 128);
  if (!ok) return fail("tensor map creation failed");
  if (alloc_slots(b) != 0) return fail("decoder slot allocation failed");
  return b;
}

void batch_destroy(Batch* b) {
  if (!b) return;
  cudaSetDevice(b->e->device);
  if (b->st) cudaStreamSynchronize(b->st);
  for (void* p : {(void*)b->d_pcm, (void*)b->d_nsamples, (void*)b->d_feat, (void*)b->d_feat32, (void*)b->d_act_a,
                  (void*)b->d_act_b, (void*)b->d_xw, (void*)b->d_hall, (void*)b->d_c, (void*)b->d_h, (void*)b->d_barrier, (void*)b->d_lstm_prof, (void*)b->d_probs64,
                  (void*)b->d_probs, (void*)b->d_win, (void*)b->d_jobs, (void*)b->d_slot_mem, (void*)b->d_slots,
                  (void*)b->d_inputs, (void*)b->d_finals, (void*)b->d_out_mem, (void*)b->d_winmat, (void*)b->d_gen_scratch})
    if (p) cudaFree(p);
  if (b->h_pcm) cudaFreeHost(b->h_pcm);
  if (b->h_nsamples) cudaFreeHost(b->h_nsamples);
  if (b->h_out_mem) cudaFreeHost(b->h_out_mem);
  if (b->st) {
    for (auto& ev : b->ev) cudaEventDestroy(ev);
    cudaStreamDestroy(b->st);
  }
  delete b;
}

int16_t* batch_host_pcm(Batch* b, int utt) {
  return (utt >= 0 && utt < b->B_cap) ? b->h_pcm + (size_t)utt * b->S_cap : nullptr;
}

int batch_upload(Batch* b, const int16_t* const* pcm, const unsigned* n_samples, int B) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  if (B < 1 || B > b->B_cap) return -1;
  const auto& m = b->e->hm;
  b->B = B;
  b->T_max = 0;
  size_t to_copy = 0;
  for (int u = 0; u < B; ++u) {
    if ((int)n_samples[u] > b->S_cap) return -2;
    b->h_nsamples[u] = (int)n_samples[u];
    b->T[u] = frames_for((int)n_samples[u], m.win_len, m.win_step);
    b->T_max = std::max(b->T_max, b->T[u]);
    if (pcm[u] != b->h_pcm + (size_t)u * b->S_cap) to_copy += (size_t)n_samples[u] * 2;
  }
  // stage into pinned memory (skipped for utterances the caller already wrote into batch_host_pcm(b, u))
  auto stage = [&](int u0, int u1) {
    for (int u = u0; u < u1; ++u) {
      int16_t* dst = b->h_pcm + (size_t)u * b->S_cap;
      if (pcm[u] != dst) memcpy(dst, pcm[u], (size_t)n_samples[u] * 2);
    }
  };
  if (to_copy > (8u << 20) && B >= 8) {
    const int nthr = 8;
    std::vector<std::thread> th;
    for (int t = 0; t < nthr; ++t) th.emplace_back(stage, B * t / nthr, B * (t + 1) / nthr);
    for (auto& t : th) t.join();
  } else {
    stage(0, B);
  }
  cudaEventRecord(b->ev[0], b->st);
  // one strided copy when every utterance has the same length, else per utterance
  bool same = true;
  for (int u = 1; u < B; ++u) same &= n_samples[u] == n_samples[0];
  if (same) {
    CUDA_OK(cudaMemcpy2DAsync(b->d_pcm, (size_t)b->S_cap * 2, b->h_pcm, (size_t)b->S_cap * 2, (size_t)n_samples[0] * 2, B,
                              cudaMemcpyHostToDevice, b->st));
  } else {
    for (int u = 0; u < B; ++u)
      CUDA_OK(cudaMemcpyAsync(b->d_pcm + (size_t)u * b->S_cap, b->h_pcm + (size_t)u * b->S_cap, (size_t)n_samples[u] * 2,
                              cudaMemcpyHostToDevice, b->st));
  }
  CUDA_OK(cudaMemcpyAsync(b->d_nsamples, b->h_nsamples, 4ull * B, cudaMemcpyHostToDevice, b->st));
  cudaEventRecord(b->ev[1], b->st);
  CUDA_OK(cudaStreamSynchronize(b->st));
  cudaEventElapsedTime(&b->times.h2d, b->ev[0], b->ev[1]);
  return 0;
}

namespace {


// Every LSTM kernel is ONE launch for all T steps with a device-wide barrier per step, so all of its CTAs must be
// co-resident: the launch is cooperative (the runtime refuses it otherwise) and the occupancy query below is checked
// first.  Exception: with a profiler attached (Engine::lstm_noncoop) the cooperative attribute is dropped, because this
// driver fails cooperative + cluster launches under Nsight Compute ("LaunchFailed"); a profiler serialises kernels, so
// the co-residency that the occupancy query established still holds.  The same fallback is taken once if a cooperative
// launch is refused synchronously.
template <class K, class... Args>
int launch_coop_cluster(Engine* e, K kern, cudaLaunchConfig_t cfgl, int cluster, Args... args) {
  cudaLaunchAttribute attrs[2];
  int n = 0;
  if (!e->lstm_noncoop) {
    attrs[n].id = cudaLaunchAttributeCooperative;
    attrs[n].val.cooperative = 1;
    ++n;
  }
  if (cluster > 1) {
    attrs[n].id = cudaLaunchAttributeClusterDimension;
    attrs[n].val.clusterDim.x = cluster;
    attrs[n].val.clusterDim.y = 1;
    attrs[n].val.clusterDim.z = 1;
    ++n;
  }
  cfgl.attrs = attrs;
  cfgl.numAttrs = n;
  cudaError_t err = cudaLaunchKernelEx(&cfgl, kern, args...);
  if (err != cudaSuccess && !e->lstm_noncoop && cluster > 1) {
    cudaGetLastError();
    fprintf(stderr, "[stt_b200] cooperative + cluster launch refused (%s); co-residency was verified by the occupancy query, "
                    "retrying without the cooperative attribute\n", cudaGetErrorString(err));
    e->lstm_noncoop = true;
    attrs[0] = attrs[1];
    cfgl.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfgl, kern, args...);
  }
  if (err != cudaSuccess) {
    fprintf(stderr, "[stt_b200] LSTM launch failed: %s\n", cudaGetErrorString(err));
    return -1;
  }
  return 0;
}

// how many CTAs of `kern` (cluster size `cluster`) can be resident at once
template <class K>
long long resident_ctas(Engine* e, K kern, int threads, int smem, int grid, int cluster) {
  if (cluster > 1) {
    cudaLaunchConfig_t c{};
    c.gridDim = dim3(grid);
    c.blockDim = dim3(threads);
    c.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = cluster;
    a[0].val.clusterDim.y = 1;
    a[0].val.clusterDim.z = 1;
    c.attrs = a;
    c.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &c) != cudaSuccess) { cudaGetLastError(); n_clusters = 0; }
    return (long long)n_clusters * cluster;
  }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem) != cudaSuccess) { cudaGetLastError(); per_sm = 0; }
  return (long long)per_sm * e->num_sms;
}

template <int MT, int STAGES, int CS>
int launch_lstm_inst(Batch* b, const sttlstm::LstmParams& lp, int grid, cudaStream_t st, bool probe_only, bool* fits) {
  using L = sttlstm::SmemLayout<MT, STAGES>;
  Engine* e = b->e;
  auto kern = sttlstm::lstm_tc_kernel<MT, STAGES, CS>;
  constexpr int cs_idx = CS == 8 ? 3 : CS == 4 ? 2 : CS == 2 ? 1 : 0;
  if (ensure_smem(&e->cfg_mask, (MT == 1 ? kBitLstm1 : kBitLstm2) + cs_idx, kern, L::kTotal)) return -1;
  if (probe_only) {
    *fits = resident_ctas(e, kern, sttlstm::kNumThreads, L::kTotal, grid, CS) >= grid;
    return 0;
  }
  cudaLaunchConfig_t cfgl{};
  cfgl.gridDim = dim3(grid);
  cfgl.blockDim = dim3(sttlstm::kNumThreads);
  cfgl.dynamicSmemBytes = L::kTotal;
  cfgl.stream = st;
  // A-operand tensor map with a 128/CS-row box (each CTA of a cluster fetches one slice and multicasts it)
  CUtensorMap tm_h;
  const size_t rows = (size_t)b->T_cap * b->B_cap + b->B_cap + 256;
  if (!make_tmap_2d(&tm_h, b->d_hall, rows, e->Cp, e->Cp, 128 / CS)) return -1;
  if (CS == 1 && e->lstm_noncoop) {   // no cluster attribute to keep: plain launch (profiler only)
    kern<<<grid, sttlstm::kNumThreads, L::kTotal, st>>>(tm_h, e->tm_wh, lp);
    CUDA_OK(cudaGetLastError());
    return 0;
  }
  return launch_coop_cluster(e, kern, cfgl, CS, tm_h, e->tm_wh, lp);
}

template <int MT, int STAGES>
int launch_lstm_mt(Batch* b, const sttlstm::LstmParams& lp, int grid, cudaStream_t st) {
  // largest cluster size that divides the grid AND lets the whole grid be co-resident (GPC sizes vary per part)
  Engine* e = b->e;
  int cs = 0;
  {
    std::lock_guard<std::mutex> lk(e->launch_mu);
    auto it = e->lstm_cluster.find({MT, grid});
    if (it != e->lstm_cluster.end()) cs = it->second;
    if (cs == 0) {
      const int cap = e->opt_lstm_cluster_cap;
      bool fits = false;
      cs = 1;
      if (cap >= 8 && grid % 8 == 0 && launch_lstm_inst<MT, STAGES, 8>(b, lp, grid, st, true, &fits) == 0 && fits) cs = 8;
      else if (cap >= 4 && grid % 4 == 0 && launch_lstm_inst<MT, STAGES, 4>(b, lp, grid, st, true, &fits) == 0 && fits) cs = 4;
      else if (cap >= 2 && grid % 2 == 0 && launch_lstm_inst<MT, STAGES, 2>(b, lp, grid, st, true, &fits) == 0 && fits) cs = 2;
      else if (launch_lstm_inst<MT, STAGES, 1>(b, lp, grid, st, true, &fits) != 0 || !fits) {
        fprintf(stderr, "[stt_b200] the LSTM grid (%d CTAs) cannot be co-resident on this device\n", grid);
        return -1;
      }
      e->lstm_cluster[{MT, grid}] = cs;
      if (e->verbose) fprintf(stderr, "[stt_b200] LSTM grid %d: cluster size %d\n", grid, cs);
    }
  }
  bool unused;
  if (cs == 8) return launch_lstm_inst<MT, STAGES, 8>(b, lp, grid, st, false, &unused);
  if (cs == 4) return launch_lstm_inst<MT, STAGES, 4>(b, lp, grid, st, false, &unused);
  if (cs == 2) return launch_lstm_inst<MT, STAGES, 2>(b, lp, grid, st, false, &unused);
  return launch_lstm_inst<MT, STAGES, 1>(b, lp, grid, st, false, &unused);
}

// CTA-pair (cta_group::2) kernel for 129..256 utterances; returns 1 if it cannot be used.
int launch_lstm_pair(Batch* b, const sttlstm::LstmParams& lp, int grid, cudaStream_t st) {
  Engine* e = b->e;
  if (e->opt_lstm_pair == 0 || grid % 2 != 0) return 1;
  using L = sttlstm::PairSmem;
  auto kern = sttlstm::lstm_pair_kernel;
  {
    std::lock_guard<std::mutex> lk(e->launch_mu);
    if (e->lstm_pair_usable < 0) {
      e->lstm_pair_usable = 0;
      if (ensure_smem(&e->cfg_mask, kBitLstmPair, kern, L::kTotal) == 0)
        e->lstm_pair_usable = resident_ctas(e, kern, sttlstm::kNumThreads, L::kTotal, grid, 2) >= grid ? 1 : 0;
      if (e->verbose) fprintf(stderr, "[stt_b200] LSTM pair kernel usable=%d\n", e->lstm_pair_usable);
    }
    if (!e->lstm_pair_usable) return 1;
  }
  cudaLaunchConfig_t cfgl{};
  cfgl.gridDim = dim3(grid);
  cfgl.blockDim = dim3(sttlstm::kNumThreads);
  cfgl.dynamicSmemBytes = L::kTotal;
  cfgl.stream = st;
  CUtensorMap tm_h;
  const size_t rows = (size_t)b->T_cap * b->B_cap + b->B_cap + 256;
  if (!make_tmap_2d(&tm_h, b->d_hall, rows, e->Cp, e->Cp, 128)) return -1;
  return launch_coop_cluster(e, kern, cfgl, 2, tm_h, e->tm_wh, lp);
}

// Ping-pong CTA-pair kernel (two groups of <= 128 utterances, M = 128 MMAs); returns 1 if it cannot be used.
// CS = 8: four pairs per cluster share the h tiles by TMA multicast (lstm2_tc.cuh).
template <int KB, int STAGES, int CS = 2>
int launch_lstm_pp_inst(Batch* b, const sttlstm::LstmParams& lp, int grid, cudaStream_t st) {
  Engine* e = b->e;
  using L = sttlstm::PPSmem<KB, STAGES>;
  auto kern = sttlstm::lstm_pp_kernel<KB, STAGES, CS>;
  constexpr int idx = CS == 8 ? 3 : CS == 4 ? 4 : KB == 4 ? 0 : KB == 2 ? 1 : 2;
  if (grid % CS != 0) return 1;
  {
    std::lock_guard<std::mutex> lk(e->launch_mu);
    if (e->lstm_pp_usable[idx] < 0) {
      e->lstm_pp_usable[idx] = 0;
      if (ensure_smem(&e->cfg_mask, kBitLstmPP4 + idx, kern, L::kTotal) == 0)
      {
        const long long res = resident_ctas(e, kern, sttlstm::kPPThreads, L::kTotal, grid, CS);
        e->lstm_pp_usable[idx] = res >= grid ? 1 : 0;
        if (e->verbose) fprintf(stderr, "[stt_b200] LSTM ping-pong kernel<%d,%d,cluster %d>: %lld CTAs can be resident, %d needed\n", KB, STAGES, CS, res, grid);
      }
    }
    if (!e->lstm_pp_usable[idx]) return 1;
  }
  cudaLaunchConfig_t cfgl{};
  cfgl.gridDim = dim3(grid);
  cfgl.blockDim = dim3(sttlstm::kPPThreads);
  cfgl.dynamicSmemBytes = L::kTotal;
  cfgl.stream = st;
  // 3-D views {64 columns, rows, K block}: one request brings KB K blocks of 64 rows
  const int Cp = e->Cp;
  CUtensorMap tm_h, tm_w;
  {
    const uint64_t rows = (uint64_t)b->T_cap * b->B_cap + b->B_cap + 256;
    uint64_t dims[3] = {64, rows, (uint64_t)(Cp / 64)};
    uint64_t strides[2] = {(uint64_t)Cp * 2, 128};
    uint32_t box[3] = {64, 64, (uint32_t)(KB * 2 / CS)};   // multicast variant: this CTA's share of the K blocks
    if (!make_tmap(&tm_h, b->d_hall, 3, dims, strides, box)) return 1;
    dims[1] = (uint64_t)4 * Cp;
    box[2] = (uint32_t)KB;
    if (!make_tmap(&tm_w, e->wh, 3, dims, strides, box)) return 1;
  }
  return launch_coop_cluster(e, kern, cfgl, CS, tm_h, tm_w, lp);
}
int launch_lstm_pp(Batch* b, const sttlstm::LstmParams& lp, int grid, cudaStream_t st) {
  const int mode = b->e->opt_lstm_pp_mode;
  if (mode == 0 || grid % 2 != 0) return 1;
  if (mode == 2) return launch_lstm_pp_inst<2, 6>(b, lp, grid, st);
  if (mode == 1) return launch_lstm_pp_inst<1, 12>(b, lp, grid, st);
  if (mode == 8) {   // four pairs per cluster share the h tiles by TMA multicast; falls back when 8-CTA clusters do not fit
    const int rc = launch_lstm_pp_inst<4, 3, 8>(b, lp, grid, st);
    if (rc <= 0) return rc;
  }
  if (mode == 8 || mode == 5) {   // two pairs per cluster
    const int rc = launch_lstm_pp_inst<4, 3, 4>(b, lp, grid, st);
    if (rc <= 0) return rc;
  }
  return launch_lstm_pp_inst<4, 3>(b, lp, grid, st);
}

int launch_lstm(Batch* b, const sttlstm::LstmParams& lp, int grid, int B, cudaStream_t st) {
  if (B <= 128) {
    if (b->e->opt_lstm_small_pp && B >= b->e->opt_lstm_small_pp) {   // the pair kernel with ONE group (no ping-pong partner)
      const int rc = launch_lstm_pp(b, lp, grid, st);
      if (rc <= 0) return rc;
    }
    return launch_lstm_mt<1, 6>(b, lp, grid, st);
  }
  int rc = launch_lstm_pp(b, lp, grid, st);
  if (rc <= 0) return rc;
  rc = launch_lstm_pair(b, lp, grid, st);
  if (rc <= 0) return rc;
  return launch_lstm_mt<2, 5>(b, lp, grid, st);
}

// dense 1..3 -> xw -> LSTM -> dense 5,6 + softmax for rows (T timesteps x B utterances); features already in d_feat.
// probs are written at [b, out_t_offset + t].  LSTM initial state = (d_c, block 0 of d_hall); final state -> d_c, d_h.
int run_am(Batch* b, int B, int T, int out_t_offset, bool time_it) {
  Engine* e = b->e;
  const auto& m = e->hm;
  const int Hp = e->Hp, Cp = e->Cp;
  const int M = T * B;
  cudaStream_t st = b->st;
  // ---- layer 1 through the overlapping-window TMA view
  int b_box = 1;
  while (b_box < B && b_box < 128) b_box <<= 1;
  const int t_box = 128 / b_box;
  CUtensorMap tm_feat;
  bool use_overlap_view = e->use_overlap_view;
  if (use_overlap_view) {
    uint64_t dims[3] = {(uint64_t)e->K1, (uint64_t)(b->rows_per_utt - 2 * m.n_context), (uint64_t)b->B_cap};
    uint64_t strides[2] = {64, (uint64_t)b->rows_per_utt * 64};
    uint32_t box[3] = {64, (uint32_t)t_box, (uint32_t)b_box};
    if (!make_tmap(&tm_feat, b->d_feat, 3, dims, strides, box)) {
      fprintf(stderr, "[stt_b200] overlapping-row tensor map rejected; using the gathered window matrix\n");
      use_overlap_view = e->use_overlap_view = false;
    }
  }
  sttgemm::GemmParams p{};
  p.relu_clip = m.relu_clip;
  p.B = B; p.T = T; p.b_box = b_box; p.t_box = t_box;
  if (time_it) cudaEventRecord(b->ev[3], st);
  p.M = M; p.N = Hp; p.K = round_up(e->K1, 64); p.bias = e->b1; p.out = b->d_act_a;
  if (use_overlap_view) {
    if (launch_gemm<256, sttgemm::kEpiClipReluF16, sttgemm::kWindows3D>(tm_feat, e->tm_w1, p, e->num_sms, st, &e->cfg_mask)) return -1;
  } else {
    const int K1p = round_up(e->K1, 64);
    if (!b->d_winmat) {
      const size_t rows = (size_t)b->T_cap * b->B_cap + 128;
      CUDA_OK(cudaMalloc((void**)&b->d_winmat, rows * K1p * 2));
      CUDA_OK(cudaMemset(b->d_winmat, 0, rows * K1p * 2));
      if (!make_tmap_2d(&b->tm_winmat, b->d_winmat, rows, K1p, K1p, 128)) return -1;
    }
    gather_windows_kernel<<<M, 128, 0, st>>>(b->d_feat, b->d_winmat, B, T, b->rows_per_utt, e->K1, K1p);
    CUDA_OK(cudaGetLastError());
    if (launch_gemm<256, sttgemm::kEpiClipReluF16, sttgemm::kRows2D>(b->tm_winmat, e->tm_w1, p, e->num_sms, st, &e->cfg_mask)) return -1;
  }
  const bool pairs = e->opt_gemm_pair != 0;
  p.K = Hp; p.bias = e->b2; p.out = b->d_act_b;
  if (pairs ? launch_gemm2<sttgemm::kEpiClipReluF16>(b->tm_act_a, e->tm_w2h, p, e->num_sms, st, &e->cfg_mask)
            : launch_gemm<256, sttgemm::kEpiClipReluF16, sttgemm::kRows2D>(b->tm_act_a, e->tm_w2, p, e->num_sms, st, &e->cfg_mask)) return -1;
  p.bias = e->b3; p.out = b->d_act_a;
  if (pairs ? launch_gemm2<sttgemm::kEpiClipReluF16>(b->tm_act_b, e->tm_w3h, p, e->num_sms, st, &e->cfg_mask)
            : launch_gemm<256, sttgemm::kEpiClipReluF16, sttgemm::kRows2D>(b->tm_act_b, e->tm_w3, p, e->num_sms, st, &e->cfg_mask)) return -1;
  if (time_it) cudaEventRecord(b->ev[4], st);
  // ---- hoisted input half of the LSTM matmul (+ bias)
  p.N = 4 * Cp; p.K = Hp; p.bias = e->bx; p.out = b->d_xw;
  if (pairs ? launch_gemm2<sttgemm::kEpiBiasF32>(b->tm_act_a, e->tm_wxh, p, e->num_sms, st, &e->cfg_mask)
            : launch_gemm<256, sttgemm::kEpiBiasF32, sttgemm::kRows2D>(b->tm_act_a, e->tm_wx, p, e->num_sms, st, &e->cfg_mask)) return -1;
  if (time_it) cudaEventRecord(b->ev[5], st);
  // ---- recurrence
  {
    CUDA_OK(cudaMemsetAsync(b->d_barrier, 0, 16384, st));
    sttlstm::LstmParams lp{};
    lp.B = B; lp.T = T; lp.n_cell = Cp; lp.xw = b->d_xw; lp.h_all = b->d_hall; lp.c_state = b->d_c; lp.h_state = b->d_h;
    lp.barrier = b->d_barrier;
    lp.prof = b->d_lstm_prof;
    lp.exact_h = e->opt_lstm_exact_h;
    const int grid = Cp / sttlstm::kCellsPerCta;
    if (launch_lstm(b, lp, grid, B, st)) return -1;
  }
  if (time_it) cudaEventRecord(b->ev[6], st);
  // ---- layer 5 (A = h_1..h_T = rows B.. of h_all) and layer 6 + softmax
  CUtensorMap tm_h_out;
  if (!make_tmap_2d(&tm_h_out, b->d_hall + (size_t)B * Cp, (uint64_t)M + 128, Cp, Cp, 128)) return -1;
  p.N = Hp; p.K = Cp; p.bias = e->b5; p.out = b->d_act_b;
  if (pairs ? launch_gemm2<sttgemm::kEpiClipReluF16>(tm_h_out, e->tm_w5h, p, e->num_sms, st, &e->cfg_mask)
            : launch_gemm<256, sttgemm::kEpiClipReluF16, sttgemm::kRows2D>(tm_h_out, e->tm_w5, p, e->num_sms, st, &e->cfg_mask)) return -1;
  p.N = e->N6; p.K = Hp; p.bias = e->b6; p.out = b->d_probs; p.n_valid = m.n_classes;
  p.out_T_stride = b->T_cap; p.out_t_offset = out_t_offset;
  if (e->N6 == 32) {
    if (launch_gemm<32, sttgemm::kEpiSoftmaxF32, sttgemm::kRows2D>(b->tm_act_b, e->tm_w6, p, e->num_sms, st, &e->cfg_mask)) return -1;
  } else {
    if (launch_gemm<256, sttgemm::kEpiSoftmaxF32, sttgemm::kRows2D>(b->tm_act_b, e->tm_w6, p, e->num_sms, st, &e->cfg_mask)) return -1;
  }
  if (time_it) cudaEventRecord(b->ev[7], st);
  b->launches += 7;
  return 0;
}

}  // namespace

int batch_forward(Batch* b) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  Engine* e = b->e;
  b->use_probs64 = false;
  const auto& m = e->hm;
  if (b->B < 1) return -1;
  cudaStream_t st = b->st;
  const int B = b->B, T = b->T_max;
  cudaEventRecord(b->ev[2], st);
  // features: zero the padded stream (context zeros + lanes), then one CTA per analysis window
  CUDA_OK(cudaMemsetAsync(b->d_feat, 0, (size_t)b->B_cap * b->rows_per_utt * 64, st));
  sttmfcc::BatchJob job{};
  job.pcm = b->d_pcm; job.n_samples = b->d_nsamples; job.stride = b->S_cap; job.frames_per_utt = b->T_cap;
  job.out_f32 = b->d_feat32; job.out_f16 = b->d_feat; job.f16_frames_per_utt = b->rows_per_utt;
  job.f16_row_offset = m.n_context;
  // grid covers frames_per_utt = T_cap per utterance; blocks beyond an utterance's frame count exit immediately
  {
    sttmfcc::BatchJob j2 = job;
    const long long items = (long long)B * b->T_cap;
    const int grid = (int)std::min<long long>((items + sttmfcc::kWarpsPerBlock - 1) / sttmfcc::kWarpsPerBlock,
                                              (long long)e->num_sms * 5 * 8);
    sttmfcc::mfcc_batch_kernel<<<grid, sttmfcc::kWarpsPerBlock * 32, 0, st>>>(e->tables, j2, B);
    CUDA_OK(cudaGetLastError());
    b->launches += 1;
  }
  // offline: LSTM starts from zeros (STT_CreateStream, stt.cc:535-536)
  CUDA_OK(cudaMemsetAsync(b->d_c, 0, (size_t)b->B_cap * e->Cp * 4, st));
  CUDA_OK(cudaMemsetAsync(b->d_hall, 0, (size_t)b->B_cap * e->Cp * 2, st));
  if (run_am(b, B, T, 0, true)) return -1;
  CUDA_OK(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&b->times.mfcc, b->ev[2], b->ev[3]);
  cudaEventElapsedTime(&b->times.dense123, b->ev[3], b->ev[4]);
  cudaEventElapsedTime(&b->times.lstm_in, b->ev[4], b->ev[5]);
  cudaEventElapsedTime(&b->times.lstm, b->ev[5], b->ev[6]);
  cudaEventElapsedTime(&b->times.dense56, b->ev[6], b->ev[7]);
  return 0;
}

namespace {

sttdec::DecodeParams make_decode_params(const Batch* b, int beam) {
  const Engine* e = b->e;
  sttdec::DecodeParams dp{};
  dp.n_classes = e->hm.n_classes;
  dp.beam = beam;
  dp.space_id = (int)e->hm.space_label;
  const DeviceScorer* sc = b->scorer.get();   // the one captured when this decode's state was initialised
  dp.has_scorer = sc ? 1 : 0;
  if (sc) dp.scorer = sc->view;               // alpha / beta as they are NOW (shared object, scorer.cpp:346-351)
  dp.fst_state2 = sc ? sc->fst_state2 : nullptr;
  dp.fst_arc4 = sc ? sc->fst_arc4 : nullptr;
  dp.fst_space_skip = (sc && e->opt_word_ordinals) ? sc->fst_space_skip : nullptr;
  dp.flags = e->opt_dec_flags;
  dp.ord2wid = dp.fst_space_skip ? sc->ord2wid : nullptr;
  dp.n_hot = sc ? (int)std::min<size_t>(b->hot_ids.size(), sttdec::kMaxHotWords) : 0;
  for (int h = 0; h < dp.n_hot; ++h) {
    dp.hot_id[h] = b->hot_ids[h];
    dp.hot_boost[h] = b->hot_boosts[h];
  }
  return dp;
}

int decoder_reset(Batch* b, int n_slots) {
  cudaStream_t st = b->st;
  b->scorer = b->e->scorer;   // DecoderState::init takes its own handle on the scorer (stt.cc:542-547)
  const int32_t fst_start = b->scorer ? (int32_t)b->scorer->view.fst_start : 0;
  // (parent, label) hash tables are never cleared between decodes: entries carry a generation (decoder.cuh ht_insert)
  if (++b->ht_gen > 255u) {
    for (int u = 0; u < b->B_cap; ++u)
      CUDA_OK(cudaMemsetAsync(b->h_slots[u].ht, 0, 8ull * ((size_t)b->h_slots[u].ht_mask + 1), st));
    b->ht_gen = 1;
  }
  sttdec::decoder_init_kernel<<<(n_slots + 127) / 128, 128, 0, st>>>(b->d_slots, n_slots, fst_start, b->ht_gen);
  CUDA_OK(cudaGetLastError());
  b->launches += 1;
  return 0;
}

// Which kernel decodes: the shared-memory kernel covers <= 32 labels, word-mode scorers and no vocabulary pruning (all
// the C API reaches with the English alphabet); everything else -- wide alphabets, UTF-8 bytes-output scorers, pruning,
// or merely the sorted class order that cutoff_top_n < classes implies -- runs the general kernel.
bool needs_general_decoder(const Batch* b) {
  const int C = (int)b->e->hm.n_classes;
  const DeviceScorer* sc = b->scorer.get();
  return C > 33 || (sc && sc->view.is_utf8) || b->cutoff_prob < 1.0 || b->cutoff_top_n < C;
}

int decoder_steps(Batch* b, int n_slots, const std::vector<sttdec::StepInput>& in, int beam) {
  cudaStream_t st = b->st;
  CUDA_OK(cudaMemcpyAsync(b->d_inputs, in.data(), sizeof(sttdec::StepInput) * n_slots, cudaMemcpyHostToDevice, st));
  sttdec::DecodeParams dp = make_decode_params(b, beam);
  constexpr int NT = 512;
  Engine* e = b->e;
  if (needs_general_decoder(b)) {
    if (!b->d_gen_scratch)
      CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b->d_gen_scratch),
                         4ull * sttdec::kGenScratchArrays * (size_t)b->beam_cap * b->B_cap));
    const DeviceScorer* sc = b->scorer.get();
    sttdec::GenParams gp{};
    gp.gstate = sc ? sc->gstate : nullptr;
    gp.garc = sc ? sc->garc : nullptr;
    gp.cutoff_prob = b->cutoff_prob;
    gp.cutoff_top_n = b->cutoff_top_n;
    gp.scratch = b->d_gen_scratch;
    gp.byte_wid = sc ? sc->byte_wid : nullptr;
    dp.fst_space_skip = nullptr;   // the general kernel does not maintain word ordinals: words are hashed
    dp.ord2wid = nullptr;
    sttdec::decoder_general_kernel<256><<<n_slots, 256, 0, st>>>(b->d_slots, b->d_inputs, dp, gp);
    CUDA_OK(cudaGetLastError());
    b->launches += 1;
    return 0;
  }
  auto go = [&](auto kern, int bit, size_t smem) -> int {
    if (ensure_smem(&e->cfg_mask, bit, kern, (int)smem)) return -1;
    kern<<<n_slots, NT, smem, st>>>(b->d_slots, b->d_inputs, dp);
    return 0;
  };
  int rc;
  if (b->beam_cap <= 512) {
    constexpr size_t SM = sizeof(sttdec::StepSmem<512, 5632>);
    if (b->instrument) rc = go(sttdec::decoder_step_kernel<NT, 512, 5632, true>, kBitDec512I, SM);
    else if (n_slots <= e->num_sms)   // one CTA per SM anyway: the 128-register build (streams, small batches)
      rc = go(sttdec::decoder_step_kernel<NT, 512, 5632, false, 1>, kBitDec512Solo, SM);
    else rc = go(sttdec::decoder_step_kernel<NT, 512, 5632, false>, kBitDec512, SM);
  } else if (b->beam_cap <= 2048) {
    constexpr size_t SM = sizeof(sttdec::StepSmem<2048, 0>);
    rc = b->instrument ? go(sttdec::decoder_step_kernel<NT, 2048, 0, true>, kBitDec2048I, SM)
                       : go(sttdec::decoder_step_kernel<NT, 2048, 0, false>, kBitDec2048, SM);
  } else {
    fprintf(stderr, "[stt_b200] beam widths above 2048 are not supported by the shared-memory decoder\n");
    return -1;
  }
  if (rc) return -1;
  CUDA_OK(cudaGetLastError());
  b->launches += 1;
  return 0;
}

int decoder_finalize(Batch* b, int n_slots, int beam, int num_results) {
  sttdec::DecodeParams dp = make_decode_params(b, beam);
  if (needs_general_decoder(b)) {
    dp.fst_space_skip = nullptr;
    dp.ord2wid = nullptr;
  }
  sttdec::decoder_finalize_kernel<256><<<n_slots, 256, 0, b->st>>>(b->d_slots, b->d_finals, dp, num_results);
  CUDA_OK(cudaGetLastError());
  b->launches += 1;
  return 0;
}

int ensure_results_capacity(Batch* b, int num_results, int tok_cap = 0) {
  if (num_results <= b->max_results && tok_cap <= b->out_tok_cap) return 0;
  // re-create output buffers with a larger result count
  cudaStreamSynchronize(b->st);
  cudaFree(b->d_out_mem); b->d_out_mem = nullptr;
  cudaFreeHost(b->h_out_mem); b->h_out_mem = nullptr;
  cudaFree(b->d_finals); b->d_finals = nullptr;
  b->max_results = std::max(b->max_results, num_results);
  b->out_tok_cap = std::max(b->out_tok_cap, tok_cap);
  const int R = b->max_results, Tm = b->out_tok_cap;
  size_t per = 8;
  const size_t o_conf = per; per += 8ull * R;
  const size_t o_nt = per; per += 4ull * R;
  per = align_up(per, 8);
  const size_t o_tok = per; per += 4ull * R * Tm;
  const size_t o_ts = per; per += 4ull * R * Tm;
  per = align_up(per, 256);
  b->out_bytes_per_utt = per;
  CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b->d_out_mem), per * b->B_cap));
  CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&b->h_out_mem), per * b->B_cap));
  std::vector<sttdec::FinalOut> fo(b->B_cap);
  for (int u = 0; u < b->B_cap; ++u) {
    uint8_t* base = b->d_out_mem + per * u;
    fo[u].max_results = R; fo[u].max_tokens = Tm;
    fo[u].n_results = (int*)base; fo[u].confidence = (double*)(base + o_conf); fo[u].n_tokens = (int*)(base + o_nt);
    fo[u].tokens = (uint32_t*)(base + o_tok); fo[u].timesteps = (uint32_t*)(base + o_ts);
  }
  CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&b->d_finals), sizeof(sttdec::FinalOut) * b->B_cap));
  CUDA_OK(cudaMemcpy(b->d_finals, fo.data(), sizeof(sttdec::FinalOut) * b->B_cap, cudaMemcpyHostToDevice));
  return 0;
}

// parse one utterance's result block (host copy)
void parse_results(const Batch* b, const uint8_t* base, std::vector<Decoded>* out) {
  const int R = b->max_results, Tm = b->out_tok_cap;
  size_t per = 8;
  const size_t o_conf = per; per += 8ull * R;
  const size_t o_nt = per; per += 4ull * R;
  per = align_up(per, 8);
  const size_t o_tok = per; per += 4ull * R * Tm;
  const size_t o_ts = per;
  const int n = *reinterpret_cast<const int*>(base);
  out->clear();
  for (int r = 0; r < n; ++r) {
    Decoded d;
    d.confidence = reinterpret_cast<const double*>(base + o_conf)[r];
    const int nt = std::min(reinterpret_cast<const int*>(base + o_nt)[r], Tm);
    const uint32_t* tok = reinterpret_cast<const uint32_t*>(base + o_tok) + (size_t)r * Tm;
    const uint32_t* ts = reinterpret_cast<const uint32_t*>(base + o_ts) + (size_t)r * Tm;
    d.tokens.assign(tok, tok + nt);
    d.timesteps.assign(ts, ts + nt);
    out->push_back(std::move(d));
  }
}

}  // namespace

int batch_set_hot_words(Batch* b, const std::vector<std::string>& words, const std::vector<float>& boosts) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  // Words outside the LM vocabulary can never equal a word of a scored n-gram (every scored word passed the
  // dictionary, whose words come from the LM vocabulary), so they are dropped here.
  b->hot_ids.clear();
  b->hot_boosts.clear();
  const Engine* e = b->e;
  const DeviceScorer* sc = e->scorer.get();   // the scorer the coming decode will capture
  if (!sc) return 0;
  sttscorer::ScorerView hv = sc->view;
  hv.blob = sc->vocab_host.data();            // the vocabulary hashes sit at the start of the file
  for (size_t i = 0; i < words.size(); ++i) {
    const uint32_t id = sttscorer::vocab_index(hv, reinterpret_cast<const uint8_t*>(words[i].data()), (uint32_t)words[i].size());
    if (id != 0) {
      b->hot_ids.push_back(id);
      b->hot_boosts.push_back(boosts[i]);
    }
  }
  if (b->hot_ids.size() > (size_t)sttdec::kMaxHotWords) {
    fprintf(stderr, "[stt_b200] more than %d in-vocabulary hot words: the rest are ignored\n", sttdec::kMaxHotWords);
  }
  return 0;
}

int batch_decode(Batch* b, int beam, int num_results) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  if (b->B < 1 || beam < 1 || beam > b->beam_cap) return -1;
  if (ensure_results_capacity(b, num_results)) return -1;
  b->cur_beam = beam;
  b->cur_results = num_results;
  cudaStream_t st = b->st;
  cudaEventRecord(b->ev[8], st);
  if (decoder_reset(b, b->B)) return -1;
  std::vector<sttdec::StepInput> in(b->B);
  for (int u = 0; u < b->B; ++u) {
    in[u].probs = b->d_probs + (size_t)u * b->T_cap * b->e->hm.n_classes;
    in[u].probs64 = b->use_probs64 ? b->d_probs64 + (size_t)u * b->T_cap * b->e->hm.n_classes : nullptr;
    in[u].n_steps = b->T[u];
  }
  if (decoder_steps(b, b->B, in, beam)) return -1;
  if (decoder_finalize(b, b->B, beam, num_results)) return -1;
  cudaEventRecord(b->ev[9], st);
  CUDA_OK(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&b->times.decode, b->ev[8], b->ev[9]);
  return 0;
}

int batch_fetch(Batch* b, std::vector<std::vector<Decoded>>* out) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  cudaStream_t st = b->st;
  cudaEventRecord(b->ev[10], st);
  // header + first result rows are small; copy only what the results can occupy: [0, o_ts + used)
  CUDA_OK(cudaMemcpyAsync(b->h_out_mem, b->d_out_mem, b->out_bytes_per_utt * b->B, cudaMemcpyDeviceToHost, st));
  cudaEventRecord(b->ev[11], st);
  CUDA_OK(cudaStreamSynchronize(st));
  cudaEventElapsedTime(&b->times.d2h, b->ev[10], b->ev[11]);
  // overflow check: a decoder that ran out of arena space must fail loudly (flag written by the finalize kernel)
  for (int u = 0; u < b->B; ++u) {
    if (reinterpret_cast<const int*>(b->h_out_mem + b->out_bytes_per_utt * u)[1]) {
      fprintf(stderr, "[stt_b200] decoder capacity exceeded for utterance %d\n", u);
      return -3;
    }
  }
  out->resize(b->B);
  for (int u = 0; u < b->B; ++u) parse_results(b, b->h_out_mem + b->out_bytes_per_utt * u, &(*out)[u]);
  return 0;
}

int batch_phase_cycles(Batch* b, unsigned long long* out8) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  for (int q = 0; q < 8; ++q) out8[q] = 0;
  for (int u = 0; u < b->B; ++u) {
    unsigned long long ph[8];
    CUDA_OK(cudaMemcpy(ph, b->h_slots[u].phase_cycles, sizeof(ph), cudaMemcpyDeviceToHost));
    for (int q = 0; q < 8; ++q) out8[q] += ph[q];
  }
  return 0;
}

int batch_lstm_profile(Batch* b, unsigned long long* out3) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  const int grid = b->e->Cp / sttlstm::kCellsPerCta;
  std::vector<unsigned long long> h((size_t)grid * 4);
  CUDA_OK(cudaMemcpy(h.data(), b->d_lstm_prof, h.size() * 8, cudaMemcpyDeviceToHost));
  out3[0] = out3[1] = out3[2] = 0;
  for (int g = 0; g < grid; ++g)
    for (int k = 0; k < 3; ++k) out3[k] = std::max(out3[k], h[(size_t)g * 4 + k]);
  if (getenv("STT_B200_VERBOSE")) {  // ping-pong kernel's extra role counters (cycles per launch, max over CTAs)
    std::vector<unsigned long long> x((size_t)grid * 4);
    cudaMemcpy(x.data(), b->d_lstm_prof + 2048, x.size() * 8, cudaMemcpyDeviceToHost);
    unsigned long long mx[4] = {0, 0, 0, 0};
    for (int g = 0; g < grid; ++g)
      for (int k = 0; k < 4; ++k) mx[k] = std::max(mx[k], x[(size_t)g * 4 + k]);
    fprintf(stderr, "[stt_b200] lstm_pp roles: producer pre-phase %llu, producer main loop %llu, MMA wait for first tile %llu, epilogue wait for MMA %llu\n",
            mx[0], mx[1], mx[2], mx[3]);
  }
  return 0;
}

int batch_lm_stats(Batch* b, unsigned long long* words, unsigned long long* calls) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  *words = 0;
  *calls = 0;
  for (int u = 0; u < b->B; ++u) {
    uint32_t sc[16];
    CUDA_OK(cudaMemcpy(sc, b->h_slots[u].scalars, sizeof(sc), cudaMemcpyDeviceToHost));
    *words += sc[7];
    *calls += sc[8];
  }
  return 0;
}

int batch_set_cutoff(Batch* b, double cutoff_prob, int cutoff_top_n) {
  if (!(cutoff_prob > 0.0) || cutoff_top_n < 1) return -1;
  b->cutoff_prob = cutoff_prob;
  b->cutoff_top_n = cutoff_top_n;
  return 0;
}

int batch_decoder_scalars(Batch* b, unsigned long long* out16) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  for (int q = 0; q < 16; ++q) out16[q] = 0;
  for (int u = 0; u < b->B; ++u) {
    uint32_t sc[16];
    CUDA_OK(cudaMemcpy(sc, b->h_slots[u].scalars, sizeof(sc), cudaMemcpyDeviceToHost));
    for (int q = 0; q < 16; ++q) out16[q] += sc[q];
  }
  return 0;
}

int batch_copy_features(Batch* b, int utt, float* out) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  if (utt < 0 || utt >= b->B) return -1;
  const int ni = b->e->hm.n_input;
  CUDA_OK(cudaMemcpy(out, b->d_feat32 + (size_t)utt * b->T_cap * ni, (size_t)b->T[utt] * ni * 4, cudaMemcpyDeviceToHost));
  return b->T[utt];
}
int batch_copy_probs(Batch* b, int utt, float* out) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  if (utt < 0 || utt >= b->B) return -1;
  const int C = b->e->hm.n_classes;
  CUDA_OK(cudaMemcpy(out, b->d_probs + (size_t)utt * b->T_cap * C, (size_t)b->T[utt] * C * 4, cudaMemcpyDeviceToHost));
  return b->T[utt];
}
int batch_set_probs64(Batch* b, const double* probs, const int* T, int B, int T_stride) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  if (B < 1 || B > b->B_cap) return -1;
  const int C = b->e->hm.n_classes;
  if (!b->d_probs64) CUDA_OK(cudaMalloc((void**)&b->d_probs64, (size_t)b->B_cap * b->T_cap * C * 8));
  b->B = B;
  b->T_max = 0;
  for (int u = 0; u < B; ++u) {
    if (T[u] > b->T_cap) return -2;
    b->T[u] = T[u];
    b->T_max = std::max(b->T_max, T[u]);
    CUDA_OK(cudaMemcpy(b->d_probs64 + (size_t)u * b->T_cap * C, probs + (size_t)u * T_stride * C, (size_t)T[u] * C * 8,
                       cudaMemcpyHostToDevice));
  }
  b->use_probs64 = true;
  return 0;
}

int batch_set_probs(Batch* b, const float* probs, const int* T, int B, int T_stride) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  if (B < 1 || B > b->B_cap) return -1;
  b->use_probs64 = false;
  const int C = b->e->hm.n_classes;
  b->B = B;
  b->T_max = 0;
  for (int u = 0; u < B; ++u) {
    if (T[u] > b->T_cap) return -2;
    b->T[u] = T[u];
    b->T_max = std::max(b->T_max, T[u]);
    CUDA_OK(cudaMemcpy(b->d_probs + (size_t)u * b->T_cap * C, probs + (size_t)u * T_stride * C, (size_t)T[u] * C * 4,
                       cudaMemcpyHostToDevice));
  }
  return 0;
}

// ====================================================================================== streaming (B == 1)
int batch_stream_reset(Batch* b, int beam) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  if (beam < 1 || beam > b->beam_cap) return -1;
  Engine* e = b->e;
  b->B = 1;
  b->cur_beam = beam;
  b->stream_frames = 0;
  cudaStream_t st = b->st;
  CUDA_OK(cudaMemsetAsync(b->d_feat, 0, (size_t)b->rows_per_utt * 64, st));
  CUDA_OK(cudaMemsetAsync(b->d_c, 0, (size_t)e->Cp * 4, st));
  CUDA_OK(cudaMemsetAsync(b->d_h, 0, (size_t)e->Cp * 4, st));
  if (decoder_reset(b, 1)) return -1;
  b->stream_arena_count = 1;
  b->stream_ts_count = 1;
  // mfcc_buffer_ starts with n_context literal-zero frames (stt.cc:533)
  b->stream_frames = e->hm.n_context;
  CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int batch_stream_push_windows(Batch* b, const int16_t* windows, const int* n_valid, int n_windows, int n_zero_frames) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  Engine* e = b->e;
  const auto& m = e->hm;
  if (b->stream_frames + n_windows + n_zero_frames > b->rows_per_utt) return -2;
  cudaStream_t st = b->st;
  if (n_windows > 0) {
    if (n_windows > b->T_cap + 1) return -2;
    CUDA_OK(cudaMemcpyAsync(b->d_win, windows, (size_t)n_windows * m.win_len * 2, cudaMemcpyHostToDevice, st));
    std::vector<sttmfcc::FrameJob> jobs(n_windows);
    for (int i = 0; i < n_windows; ++i) {
      jobs[i].pcm = b->d_win + (size_t)i * m.win_len;
      jobs[i].n_valid = n_valid[i];
      jobs[i].out_f32 = nullptr;
      jobs[i].out_f16 = b->d_feat + (size_t)(b->stream_frames + i) * sttmfcc::kFeatLanes;
    }
    CUDA_OK(cudaMemcpyAsync(b->d_jobs, jobs.data(), sizeof(sttmfcc::FrameJob) * n_windows, cudaMemcpyHostToDevice, st));
    sttmfcc::mfcc_jobs_kernel<<<(n_windows + sttmfcc::kWarpsPerBlock - 1) / sttmfcc::kWarpsPerBlock, sttmfcc::kWarpsPerBlock * 32, 0, st>>>(
        e->tables, b->d_jobs, n_windows);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(st));  // `jobs` is a host temporary
    b->launches += 1;
    b->stream_frames += n_windows;
  }
  if (n_zero_frames > 0) {
    CUDA_OK(cudaMemsetAsync(b->d_feat + (size_t)b->stream_frames * sttmfcc::kFeatLanes, 0, (size_t)n_zero_frames * 64, st));
    b->stream_frames += n_zero_frames;
  }
  return 0;
}

namespace {

uint32_t host_ht_hash(unsigned long long k) {   // decoder.cuh ht_hash
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  return (uint32_t)k;
}

// Garbage collection of a stream's decoder arena.  The reference's trie never holds more than the live prefixes and
// their ancestors: PathTrie::remove (path_trie.cpp:192-209) deletes a pruned node that has no children and walks up
// its chain of dead parents, so memory is O(beam x transcript length) however long the stream runs.  The device arena
// is append-only while a chunk is being decoded; between chunks, when the next chunk might not fit, the slot is
// brought to the host, everything unreachable from the live list is dropped (nodes, their LM cache rows, timestep-tree
// nodes), ids are renumbered in creation order (a parent still precedes its children), every node's created-children
// mask and the (parent, label) hash table are rebuilt from the survivors, and the slot goes back.  A dropped node that
// re-enters the beam later is created afresh -- exactly what the reference does after having deleted it.
int stream_compact(Batch* b) {
  using sttdec::kNone;
  cudaStream_t st = b->st;
  CUDA_OK(cudaStreamSynchronize(st));
  sttdec::Slot& sl = b->h_slots[0];
  uint32_t sc[16];
  CUDA_OK(cudaMemcpy(sc, sl.scalars, sizeof(sc), cudaMemcpyDeviceToHost));
  const uint32_t n_live = sc[0], A = std::min(sc[2], sl.arena_cap), TS = std::min(sc[3], sl.ts_cap);
  const int SW = sttdec::kStateWords;
  std::vector<sttdec::Node> nodes(A);
  std::vector<double> lmc(A);
  std::vector<uint32_t> lmsw((size_t)A * SW), lmm(A), live_node(n_live), live_ts(n_live);
  std::vector<float> lmsb((size_t)A * SW);
  std::vector<uint2> tst(TS);
  CUDA_OK(cudaMemcpy(nodes.data(), sl.nodes, sizeof(sttdec::Node) * A, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(lmc.data(), sl.lm_cond, 8ull * A, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(lmsw.data(), sl.lm_sw, 4ull * A * SW, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(lmsb.data(), sl.lm_sb, 4ull * A * SW, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(lmm.data(), sl.lm_meta, 4ull * A, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(tst.data(), sl.ts_tree, 8ull * TS, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(live_node.data(), sl.node, 4ull * n_live, cudaMemcpyDeviceToHost));
  CUDA_OK(cudaMemcpy(live_ts.data(), sl.ts, 4ull * n_live, cudaMemcpyDeviceToHost));
  // ---- mark what the live prefixes reach
  std::vector<uint8_t> keep(A, 0), keep_ts(TS, 0);
  keep[0] = 1;
  if (TS) keep_ts[0] = 1;
  for (uint32_t i = 0; i < n_live; ++i) {
    for (uint32_t id = live_node[i]; id != kNone && id < A && !keep[id]; id = nodes[id].parent) keep[id] = 1;
    for (uint32_t t = live_ts[i]; t != kNone && t < TS && !keep_ts[t]; t = tst[t].x) keep_ts[t] = 1;
  }
  std::vector<uint32_t> map(A, kNone), map_ts(TS, kNone);
  uint32_t A2 = 0, TS2 = 0;
  for (uint32_t id = 0; id < A; ++id)
    if (keep[id]) map[id] = A2++;
  for (uint32_t t = 0; t < TS; ++t)
    if (keep_ts[t]) map_ts[t] = TS2++;
  // ---- rewrite (in place, front to back: map[id] <= id)
  for (uint32_t id = 0; id < A; ++id) {
    if (!keep[id]) continue;
    sttdec::Node n = nodes[id];
    n.parent = n.parent == kNone ? kNone : map[n.parent];
    n.last_space = n.last_space == kNone ? kNone : map[n.last_space];
    n.live_slot = kNone;
    n.child_mask = 0;
    const uint32_t d = map[id];
    nodes[d] = n;
    lmc[d] = lmc[id];
    lmm[d] = lmm[id];
    for (int q = 0; q < SW; ++q) {
      lmsw[(size_t)d * SW + q] = lmsw[(size_t)id * SW + q];
      lmsb[(size_t)d * SW + q] = lmsb[(size_t)id * SW + q];
    }
  }
  for (uint32_t t = 0; t < TS; ++t) {
    if (!keep_ts[t]) continue;
    uint2 v = tst[t];
    v.x = v.x == kNone ? kNone : map_ts[v.x];
    tst[map_ts[t]] = v;
  }
  const uint32_t gen = sl.ht_gen ? sl.ht_gen : b->ht_gen;
  std::vector<unsigned long long> ht((size_t)sl.ht_mask + 1, 0ull);
  for (uint32_t d = 1; d < A2; ++d) {
    const sttdec::Node& n = nodes[d];
    if (n.parent == kNone) continue;
    nodes[n.parent].child_mask |= 1u << (n.chr & 31u);
    const unsigned long long w = ((unsigned long long)gen << 56) | ((unsigned long long)(n.parent & 0xffffffu) << 32) |
                                 ((unsigned long long)(n.chr & 0xffu) << 24) | (unsigned long long)(d & 0xffffffu);
    uint32_t h = host_ht_hash(w >> 24) & sl.ht_mask;
    while (ht[h] != 0) h = (h + 1) & sl.ht_mask;   // any placement on the probe path before an empty slot is findable
    ht[h] = w;
  }
  for (uint32_t i = 0; i < n_live; ++i) {
    live_node[i] = map[live_node[i]];
    nodes[live_node[i]].live_slot = i;
    if (live_ts[i] != kNone) live_ts[i] = map_ts[live_ts[i]];
  }
  // ---- back to the device
  CUDA_OK(cudaMemcpy(sl.nodes, nodes.data(), sizeof(sttdec::Node) * A2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sl.lm_cond, lmc.data(), 8ull * A2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sl.lm_sw, lmsw.data(), 4ull * A2 * SW, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sl.lm_sb, lmsb.data(), 4ull * A2 * SW, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sl.lm_meta, lmm.data(), 4ull * A2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sl.ts_tree, tst.data(), 8ull * TS2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sl.ht, ht.data(), 8ull * ht.size(), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sl.node, live_node.data(), 4ull * n_live, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(sl.ts, live_ts.data(), 4ull * n_live, cudaMemcpyHostToDevice));
  sc[2] = A2;
  sc[3] = TS2;
  CUDA_OK(cudaMemcpy(sl.scalars, sc, sizeof(sc), cudaMemcpyHostToDevice));
  if (b->e->verbose)
    fprintf(stderr, "[stt_b200] stream arena compacted: %u -> %u nodes, %u -> %u timestep nodes\n", A, A2, TS, TS2);
  b->stream_arena_count = A2;
  b->stream_ts_count = TS2;
  ++b->stream_compactions;
  return 0;
}

}  // namespace

long long batch_stream_compactions(const Batch* b) { return b->stream_compactions; }

int batch_stream_run(Batch* b, int n_timesteps, int n_pad_rows, bool keep_last_probs) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  (void)keep_last_probs;
  Engine* e = b->e;
  const auto& m = e->hm;
  const int need = n_timesteps + 2 * (int)m.n_context;
  if (n_timesteps < 1 || n_timesteps > b->T_cap || b->stream_frames < need) return -2;
  cudaStream_t st = b->st;
  {
    // every step appends at most `beam` arena nodes and `beam` timestep-tree nodes: make room before the chunk runs
    const unsigned long long grow = (unsigned long long)b->cur_beam * (unsigned long long)n_timesteps;
    const sttdec::Slot& sl = b->h_slots[0];
    if (b->stream_arena_count + grow > sl.arena_cap || b->stream_ts_count + grow > sl.ts_cap) {
      if (stream_compact(b)) return -1;
      if (b->stream_arena_count + grow > sl.arena_cap || b->stream_ts_count + grow > sl.ts_cap) {
        fprintf(stderr, "[stt_b200] stream decoder arena too small even after compaction (%u live-trie nodes of %u): raise "
                        "STT_B200_STREAM_ARENA_SECONDS\n", b->stream_arena_count, sl.arena_cap);
        return -3;
      }
    }
  }
  // carried LSTM state: h (fp32) -> block 0 of h_all (fp16); c stays in d_c
  f32_to_f16_kernel<<<(e->Cp + 255) / 256, 256, 0, st>>>(b->d_h, b->d_hall, (size_t)e->Cp);
  b->launches += 1;
  if (run_am(b, 1, n_timesteps, 0, false)) return -1;
  // DecoderState::next on the new rows
  std::vector<sttdec::StepInput> in(1);
  in[0].probs = b->d_probs;
  in[0].probs64 = nullptr;
  in[0].n_steps = n_timesteps;
  if (decoder_steps(b, 1, in, b->cur_beam)) return -1;
  b->last_run_T = n_timesteps;
  // consume the timesteps: slide the frame stream left
  const int keep = b->stream_frames - n_timesteps;
  shift_frames_kernel<<<1, 32, 0, st>>>(b->d_feat, keep, n_timesteps, sttmfcc::kFeatLanes, b->stream_frames);
  b->launches += 1;
  b->stream_frames = keep;
  if (n_pad_rows > 0) {
    // The reference zero-pads a partial batch to n_steps rows and the LSTM state advances through the padding
    // (tflitemodelstate.cc:381; SURVEY 3.4 "trashing").  Reproduce by running the AM on all-zero windows with the
    // outputs discarded: a scratch all-zero feature stream sits after the live frames.
    // Zero windows = frames beyond stream_frames, which are zero by construction (shift kernel / reset).
    // We temporarily view the stream starting at the first all-zero region.
    const int zero_start = b->stream_frames;  // frames [zero_start, rows_per_utt) are zero
    if (zero_start + n_pad_rows + 2 * (int)m.n_context > b->rows_per_utt) return -2;
    f32_to_f16_kernel<<<(e->Cp + 255) / 256, 256, 0, st>>>(b->d_h, b->d_hall, (size_t)e->Cp);
    b->launches += 1;
    __half* saved = b->d_feat;
    b->d_feat = saved + (size_t)zero_start * sttmfcc::kFeatLanes;
    const int saved_rows = b->rows_per_utt;
    b->rows_per_utt = saved_rows - zero_start;
    // probs of these rows land after the real ones and are never read
    int rc = run_am(b, 1, n_pad_rows, n_timesteps, false);
    b->d_feat = saved;
    b->rows_per_utt = saved_rows;
    if (rc) return -1;
  }
  CUDA_OK(cudaStreamSynchronize(st));
  uint32_t sc[8];
  CUDA_OK(cudaMemcpy(sc, b->h_slots[0].scalars, sizeof(sc), cudaMemcpyDeviceToHost));
  if (sc[6]) {
    fprintf(stderr, "[stt_b200] stream decoder capacity exceeded\n");
    return -3;
  }
  b->stream_arena_count = sc[2];
  b->stream_ts_count = sc[3];
  return 0;
}

int batch_stream_decode(Batch* b, int num_results, std::vector<Decoded>* out) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  if (ensure_results_capacity(b, num_results)) return -1;
  for (;;) {
    if (decoder_finalize(b, 1, b->cur_beam, num_results)) return -1;
    CUDA_OK(cudaMemcpyAsync(b->h_out_mem, b->d_out_mem, b->out_bytes_per_utt, cudaMemcpyDeviceToHost, b->st));
    CUDA_OK(cudaStreamSynchronize(b->st));
    // A stream is unbounded, so its transcript can outgrow the output block sized when the stream was created: the
    // finalize kernel reports the full token counts, the block grows and the (const) finalize runs again.
    const int n = *reinterpret_cast<const int*>(b->h_out_mem);
    const int* nt = reinterpret_cast<const int*>(b->h_out_mem + 8 + 8ull * b->max_results);
    int longest = 0;
    for (int r = 0; r < n; ++r) longest = std::max(longest, nt[r]);
    if (longest <= b->out_tok_cap) break;
    int cap = b->out_tok_cap;
    while (cap < longest) cap *= 2;
    if (ensure_results_capacity(b, num_results, cap)) return -1;
  }
  parse_results(b, b->h_out_mem, out);
  return 0;
}

int batch_stream_frames(const Batch* b) { return b->stream_frames; }
void batch_release_scorer(Batch* b) {
  if (b->st) cudaStreamSynchronize(b->st);
  b->scorer.reset();
}

int batch_stream_last_probs(Batch* b, std::vector<double>* out, int* n_rows) {
  cudaSetDevice(b->e->device);  // CUDA's current device is per host thread
  const int C = b->e->hm.n_classes;
  std::vector<float> tmp((size_t)b->last_run_T * C);
  CUDA_OK(cudaMemcpy(tmp.data(), b->d_probs, tmp.size() * 4, cudaMemcpyDeviceToHost));
  out->assign(tmp.begin(), tmp.end());
  *n_rows = b->last_run_T;
  return 0;
}

#ifdef STT_B200_DEV_HOOKS   // unit-test / bring-up entry points: built into libstt_b200_dev.so only (Makefile)
// ====================================================================================== GEMM unit-test hook
int debug_gemm(int M, int N, int K, const uint16_t* a_f16, const uint16_t* w_f16, const float* bias, int epi,
               float relu_clip, void* out, float* ms) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // epi: 0 clipped-ReLU -> fp16, 1 bias -> f32, 2 softmax -> probs (N = 32: 29 classes; N = 256: 200 classes, the wide
  // epilogue); + 16 selects the one-CTA 128 x 256 kernel instead of the CTA-pair 256 x 256 kernel for epilogues 0 / 1
  const bool single = (epi & 16) != 0;
  epi &= 15;
  const int BN = (epi == sttgemm::kEpiSoftmaxF32) ? (N == 256 ? 256 : 32) : 256;
  if (N % BN != 0 || K % 8 != 0) return -2;
  __half *dA, *dW;
  float* dB;
  void* dO;
  const size_t out_elem = (epi == sttgemm::kEpiClipReluF16) ? 2 : 4;
  const int n_valid = (epi == sttgemm::kEpiSoftmaxF32) ? (N == 256 ? 200 : std::min(N, 29)) : N;
  const size_t out_count = (epi == sttgemm::kEpiSoftmaxF32) ? (size_t)M * n_valid : (size_t)M * N;
  CUDA_OK(cudaMalloc((void**)&dA, (size_t)(M + 256) * K * 2));
  CUDA_OK(cudaMalloc((void**)&dW, (size_t)N * K * 2));
  CUDA_OK(cudaMalloc((void**)&dB, (size_t)N * 4));
  CUDA_OK(cudaMalloc(&dO, out_count * out_elem));
  CUDA_OK(cudaMemset(dA, 0, (size_t)(M + 256) * K * 2));
  CUDA_OK(cudaMemcpy(dA, a_f16, (size_t)M * K * 2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(dW, w_f16, (size_t)N * K * 2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(dB, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
  CUtensorMap ta, tb;
  const bool pair = !single && epi != sttgemm::kEpiSoftmaxF32;
  if (!make_tmap_2d(&ta, dA, M + 128, K, K, 128) || !make_tmap_2d(&tb, dW, N, K, K, pair ? 128 : BN)) return -1;
  sttgemm::GemmParams p{};
  p.M = M; p.N = N; p.K = round_up(K, 64); p.bias = dB; p.out = dO; p.relu_clip = relu_clip; p.n_valid = n_valid;
  p.B = 1; p.T = M; p.out_T_stride = M; p.out_t_offset = 0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int rc = 0;
  std::atomic<uint32_t> mask{0};
  for (int it = 0; it < 2 && rc == 0; ++it) {  // second run is the timed one
    cudaEventRecord(e0, 0);
    if (epi == sttgemm::kEpiClipReluF16) rc = pair ? launch_gemm2<sttgemm::kEpiClipReluF16>(ta, tb, p, sms, 0, &mask)
                                                  : launch_gemm<256, sttgemm::kEpiClipReluF16, sttgemm::kRows2D>(ta, tb, p, sms, 0, &mask);
    else if (epi == sttgemm::kEpiBiasF32) rc = pair ? launch_gemm2<sttgemm::kEpiBiasF32>(ta, tb, p, sms, 0, &mask)
                                                    : launch_gemm<256, sttgemm::kEpiBiasF32, sttgemm::kRows2D>(ta, tb, p, sms, 0, &mask);
    else if (BN == 256) rc = launch_gemm<256, sttgemm::kEpiSoftmaxF32, sttgemm::kRows2D>(ta, tb, p, sms, 0, &mask);
    else rc = launch_gemm<32, sttgemm::kEpiSoftmaxF32, sttgemm::kRows2D>(ta, tb, p, sms, 0, &mask);
    cudaEventRecord(e1, 0);
    if (cudaDeviceSynchronize() != cudaSuccess) rc = -1;
  }
  if (rc == 0) {
    if (ms) cudaEventElapsedTime(ms, e0, e1);
    if (cudaMemcpy(out, dO, out_count * out_elem, cudaMemcpyDeviceToHost) != cudaSuccess) rc = -1;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(dA); cudaFree(dW); cudaFree(dB); cudaFree(dO);
  return rc;
}

// Bring-up helper: TMEM layout of a cta_group::2 MMA with the given M (128 or 256); out = [2 CTAs][128 lanes][128 cols].
int debug_pair_layout(int M, float* out) {
  if (M != 128 && M != 256) return -2;
  std::vector<__half> ha((size_t)256 * 64, __float2half(0.f)), hb((size_t)128 * 64, __float2half(0.f));
  for (int i = 0; i < 256; ++i) { ha[(size_t)i * 64 + 0] = __float2half((float)(i + 1)); ha[(size_t)i * 64 + 1] = __float2half(1.f); }
  for (int n = 0; n < 128; ++n) { hb[(size_t)n * 64 + 0] = __float2half(1024.f); hb[(size_t)n * 64 + 1] = __float2half((float)(n + 1)); }
  __half *dA, *dB;
  float* dO;
  CUDA_OK(cudaMalloc((void**)&dA, ha.size() * 2));
  CUDA_OK(cudaMalloc((void**)&dB, hb.size() * 2));
  CUDA_OK(cudaMalloc((void**)&dO, 2 * 128 * 128 * 4));
  CUDA_OK(cudaMemset(dO, 0xff, 2 * 128 * 128 * 4));
  CUDA_OK(cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap ta, tb;
  if (!make_tmap_2d(&ta, dA, 256, 64, 64, M / 2) || !make_tmap_2d(&tb, dB, 128, 64, 64, 64)) return -1;
  auto kern = sttprobe::probe_pair_kernel;
  const int smem = 16384 + 8192 + 64 + 1024;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, M, dO));
  CUDA_OK(cudaDeviceSynchronize());
  CUDA_OK(cudaMemcpy(out, dO, 2 * 128 * 128 * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dO);
  return 0;
}

#endif  // STT_B200_DEV_HOOKS

}  // namespace stteng
