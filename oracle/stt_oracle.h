/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's feature + acoustic-model + stream
 * buffering arithmetic.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this; the product library (stt_b200/) never does.
 *
 * PARITY STATUS
 *   MFCC / spectrogram : pinned by the vendored TFLite golden vectors
 *                        (tensorflow/tensorflow/lite/kernels/mfcc_test.cc:67-90,
 *                         audio_spectrogram_test.cc:64-108) -- see tests/test_oracle_mfcc.py.
 *   Acoustic model     : "parity unpinned": the reference ships no model file and no test that
 *                        pins AM numerics; TFLite itself cannot be built offline.
 *                        The restatement follows training/coqui_stt_training/deepspeech_model.py
 *                        and is cross-checked against an independent torch-CPU fp32 version.
 *   Stream buffering   : frame-count known answers 160000->500, 46797->146, 16000->50 timesteps
 *                         -- see tests/test_oracle_stream.py.
 */
#ifndef STT_ORACLE_H
#define STT_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- spectrogram: internal/spectrogram.cc:30-37 (periodic Hann), :224-241 (ProcessCoreFFT),
 *      :155-185 (squared magnitude, double -> float).  `samples` has n_samples <= window floats;
 *      the remainder of the window is zero (tflitemodelstate.cc:343-355 copy_vector_to_tensor). */
void orc_spectrogram_frame(const float* samples, int n_samples, int window, float* out_bins);

/* ---- MFCC: internal/mfcc_mel_filterbank.cc:40-167,172-197; internal/mfcc.cc:46-62;
 *      internal/mfcc_dct.cc:25-54,56-75; op wrapper mfcc.cc:106-161. */
typedef struct orc_mfcc orc_mfcc;
orc_mfcc* orc_mfcc_new(int n_bins, double sample_rate, double lower_hz, double upper_hz, int n_channels, int n_dct);
void orc_mfcc_free(orc_mfcc* m);
void orc_mfcc_compute(const orc_mfcc* m, const float* power_spectrum, float* out);

/* ---- acoustic model: deepspeech_model.py:66-89 (dense), :144-168 + rnn_cell_impl.py:1054-1079 (LSTM,
 *      gate order i,j,f,o, forget_bias 0), :204-263 (layer stack), :353-357 (softmax).
 *      All matrices are row-major [in, out] as TF stores them (y = x @ W + b). */
typedef struct {
  int n_input;   /* 26 */
  int n_context; /* 9 */
  int n_hidden;  /* layers 1,2,3,5 */
  int n_cell;    /* LSTM cell dim (= n_hidden in the reference geometry) */
  int n_classes; /* alphabet + 1 */
  float relu_clip;
  const float *w1, *b1, *w2, *b2, *w3, *b3;
  const float *lstm_kernel; /* [n_hidden + n_cell, 4*n_cell] */
  const float *lstm_bias;   /* [4*n_cell] */
  const float *w5, *b5, *w6, *b6;
} orc_am;

/* One `infer` call (tflitemodelstate.cc:369-405): x is [n_steps, (2*n_context+1)*n_input],
 * c/h [n_cell] are updated in place, probs [n_steps, n_classes]. */
void orc_am_infer(const orc_am* am, const float* x, int n_steps, float* c, float* h, float* probs);

/* ---- hybrid int8 mode (the reference's DEFAULT export arithmetic): TFLite EvalHybridDense,
 *      tensorflow/lite/kernels/fully_connected.cc:435-503 + internal/reference/portable_tensor_utils.cc:51-70,138-161.
 *      Weights are int8 per-tensor symmetric in the same [in, out] layout. */
typedef struct {
  const orc_am* am; /* geometry + fp32 biases */
  int8_t *w1, *w2, *w3, *lstm_kernel, *w5, *w6;
  float s1, s2, s3, sk, s5, s6; /* weight scales */
} orc_am_q;
void orc_quantize_weights(const float* w, size_t n, int8_t* q, float* scale);
orc_am_q* orc_am_quantize(const orc_am* am);
void orc_am_q_free(orc_am_q* q);
void orc_am_infer_hybrid(const orc_am_q* q, const float* x, int n_steps, float* c, float* h, float* probs);

/* ---- streaming runtime: native_client/stt.cc:105-128 (feedAudioContent), :226-334
 *      (processAudioWindow, flushBuffers, pushMfccBuffer, processMfccWindow, processBatch),
 *      :519-551 (STT_CreateStream buffer initialisation). */
typedef struct orc_stream orc_stream;
orc_stream* orc_stream_new(const orc_am* am, int sample_rate, int win_len, int win_step, int n_steps);
void orc_stream_free(orc_stream* s);
void orc_stream_feed(orc_stream* s, const int16_t* pcm, unsigned n);
void orc_stream_flush(orc_stream* s, int add_zero_mfcc_vectors); /* flushBuffers(bool) */
/* Emitted so far: number of timesteps, pointer to [T, n_classes] softmax rows and [F, n_input] MFCC rows. */
int orc_stream_timesteps(const orc_stream* s); /* context windows formed so far */
int orc_stream_emitted(const orc_stream* s);   /* softmax rows produced so far (<= timesteps until flushed) */
const float* orc_stream_probs(const orc_stream* s);
int orc_stream_frames(const orc_stream* s);
const float* orc_stream_mfcc(const orc_stream* s);

#ifdef __cplusplus
}
#endif
#endif
