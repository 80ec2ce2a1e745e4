// Device engine: weights, tables and the batched GPU pipeline (MFCC -> dense x3 -> LSTM -> dense x2 + softmax ->
// beam search).  Host-side replacement of TFLiteModelState (native_client/tflitemodelstate.cc) and of the
// per-stream plumbing in native_client/stt.cc:226-334.  No CPU fallback: every entry point needs a CUDA device
// and fails loudly without one.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "model_file.h"

namespace stteng {

struct Decoded {  // ctcdecode/output.h `Output`
  double confidence = 0.0;
  std::vector<uint32_t> tokens, timesteps;
};

struct StageTimes {  // milliseconds, CUDA events on the engine's stream
  float h2d = 0, mfcc = 0, dense123 = 0, lstm_in = 0, lstm = 0, dense56 = 0, decode = 0, d2h = 0, total = 0;
};

struct Engine;
struct Batch;

// ---- engine
Engine* engine_create(const sttmodel::HostModel& m, std::string* err);
void engine_destroy(Engine* e);
const sttmodel::HostModel& engine_model(const Engine* e);
int engine_set_scorer(Engine* e, const uint8_t* bytes, size_t n);  // 0 or an STT_ERR_SCORER_* code
void engine_clear_scorer(Engine* e);
bool engine_has_scorer(const Engine* e);
void engine_set_alpha_beta(Engine* e, float alpha, float beta);
int engine_num_sms(const Engine* e);

// ---- batch context: fixed-capacity device buffers for up to B_cap utterances of <= max_samples samples.
// dec_T_cap: decoder capacity in timesteps (>= frames of max_samples; larger for long-lived streams).
Batch* batch_create(Engine* e, int B_cap, int max_samples, int beam_cap, int dec_T_cap, std::string* err);
void batch_destroy(Batch* b);
// Offline use: upload -> forward -> decode -> fetch.
int16_t* batch_host_pcm(Batch* b, int utt);                // pinned staging row (capacity max_samples) for zero-copy upload
int batch_upload(Batch* b, const int16_t* const* pcm, const unsigned* n_samples, int B);
int batch_forward(Batch* b);                                 // MFCC + acoustic model -> probs in HBM
int batch_set_hot_words(Batch* b, const std::vector<std::string>& words, const std::vector<float>& boosts);
int batch_decode(Batch* b, int beam, int num_results);       // resets the decoder, runs all timesteps, finalises
int batch_fetch(Batch* b, std::vector<std::vector<Decoded>>* out);
const StageTimes& batch_times(const Batch* b);
long long batch_kernel_launches(const Batch* b);
void batch_set_instrumented(Batch* b, bool on);              // decoder statistics build (phase clocks, LM counters) for the next decodes
// Debug / test access
int batch_phase_cycles(Batch* b, unsigned long long* out8);  // instrumentation: summed over utterances
int batch_lstm_profile(Batch* b, unsigned long long* out3);  // max over CTAs: barrier-wait, load+MMA span, epilogue cycles
int batch_lm_stats(Batch* b, unsigned long long* words_scored, unsigned long long* lm_calls);  // instrumentation
int batch_decoder_scalars(Batch* b, unsigned long long* out16);
// vocabulary pruning of the following decodes (DecoderState::init's cutoff_prob / cutoff_top_n); C API: 1.0 / 40
int batch_set_cutoff(Batch* b, double cutoff_prob, int cutoff_top_n);   // sums over the batch of Slot::scalars (decoder.cuh)
int batch_T(const Batch* b, int utt);                        // timesteps of utterance `utt` after upload
int batch_copy_features(Batch* b, int utt, float* out);      // [T, n_input] fp32 MFCC
int batch_copy_probs(Batch* b, int utt, float* out);         // [T, n_classes]
int batch_set_probs(Batch* b, const float* probs, const int* T, int B, int T_stride);  // decoder-only use
int batch_set_probs64(Batch* b, const double* probs, const int* T, int B, int T_stride);
// Streaming use (B == 1): the caller owns framing (stt.cc buffer logic) and feeds analysis windows.
int batch_stream_reset(Batch* b, int beam);                                 // zero LSTM state, DecoderState::init
int batch_stream_push_windows(Batch* b, const int16_t* windows, const int* n_valid, int n_windows, int n_zero_frames);
int batch_stream_run(Batch* b, int n_timesteps, int n_pad_rows, bool keep_last_probs);  // one `infer` + DecoderState::next
int batch_stream_decode(Batch* b, int num_results, std::vector<Decoded>* out);
int batch_stream_frames(const Batch* b);
void batch_release_scorer(Batch* b);
long long batch_stream_compactions(const Batch* b);                         // arena garbage collections of this context so far                                        // drop the scorer captured by the last reset
int batch_stream_last_probs(Batch* b, std::vector<double>* out, int* n_rows);

// raw GEMM hook for the kernel unit tests: C = epi(A[M,K] * W[N,K]^T + bias)
#ifdef STT_B200_DEV_HOOKS
int debug_pair_layout(int M, float* out);
int debug_gemm(int M, int N, int K, const uint16_t* a_f16, const uint16_t* w_f16, const float* bias, int epi,
               float relu_clip, void* out, float* ms);
#endif

}  // namespace stteng
