/* TEST INFRASTRUCTURE ONLY -- see stt_oracle.h for scope, citations and parity status. */
#include "stt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ FFT
 * The reference calls Ooura's rdft (third-party fft2d, not vendored: tensorflow/workspace2.bzl:575-580).
 * Any exact real DFT agrees with it to ~1e-13 relative before the float cast (spectrogram.cc:175-183), far
 * inside the 1e-3 tolerance of the reference's own golden vectors.  Plain iterative radix-2 in double. */
static void fft_radix2(double* re, double* im, int n) {
  for (int i = 1, j = 0; i < n; ++i) {
    int bit = n >> 1;
    for (; j & bit; bit >>= 1) j ^= bit;
    j ^= bit;
    if (i < j) {
      double t = re[i]; re[i] = re[j]; re[j] = t;
      t = im[i]; im[i] = im[j]; im[j] = t;
    }
  }
  const double pi = atan(1.0) * 4.0;
  for (int len = 2; len <= n; len <<= 1) {
    int half = len >> 1;
    for (int k = 0; k < half; ++k) {
      double ang = -2.0 * pi * k / len;
      double wr = cos(ang), wi = sin(ang);
      for (int i = k; i < n; i += len) {
        int j = i + half;
        double xr = re[j] * wr - im[j] * wi, xi = re[j] * wi + im[j] * wr;
        re[j] = re[i] - xr; im[j] = im[i] - xi;
        re[i] += xr; im[i] += xi;
      }
    }
  }
}

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

void orc_spectrogram_frame(const float* samples, int n_samples, int window, float* out_bins) {
  const int fft_len = next_pow2(window);
  double* re = (double*)calloc(fft_len, sizeof(double));
  double* im = (double*)calloc(fft_len, sizeof(double));
  const double pi = atan(1.0) * 4.0;
  for (int j = 0; j < window; ++j) {
    double w = 0.5 - 0.5 * cos((2.0 * pi * j) / window);     /* GetPeriodicHann */
    float s = j < n_samples ? samples[j] : 0.0f;
    re[j] = (double)s * w;                                    /* input_queue_[j] * window_[j] */
  }
  fft_radix2(re, im, fft_len);
  for (int i = 0; i <= fft_len / 2; ++i) {
    double v = re[i] * re[i] + im[i] * im[i];
    out_bins[i] = (float)v;                                   /* double -> float, spectrogram.cc:183 */
  }
  free(re); free(im);
}

/* ------------------------------------------------------------------ MFCC */
struct orc_mfcc {
  int n_bins, n_channels, n_dct, start_index, end_index;
  int* band_mapper;
  double* weights;
  double* cosines; /* [n_dct][n_channels] */
};

static double freq_to_mel(double f) { return 1127.0 * log1p(f / 700.0); }

orc_mfcc* orc_mfcc_new(int n_bins, double sample_rate, double lower_hz, double upper_hz, int n_channels, int n_dct) {
  orc_mfcc* m = (orc_mfcc*)calloc(1, sizeof(orc_mfcc));
  m->n_bins = n_bins; m->n_channels = n_channels; m->n_dct = n_dct;
  double* center = (double*)malloc(sizeof(double) * (n_channels + 1));
  const double mel_low = freq_to_mel(lower_hz), mel_hi = freq_to_mel(upper_hz);
  const double mel_spacing = (mel_hi - mel_low) / (double)(n_channels + 1);
  for (int i = 0; i < n_channels + 1; ++i) center[i] = mel_low + mel_spacing * (i + 1);
  const double hz_per_sbin = 0.5 * sample_rate / (double)(n_bins - 1);
  m->start_index = (int)(1.5 + lower_hz / hz_per_sbin);
  m->end_index = (int)(upper_hz / hz_per_sbin);
  m->band_mapper = (int*)malloc(sizeof(int) * n_bins);
  m->weights = (double*)malloc(sizeof(double) * n_bins);
  int channel = 0;
  for (int i = 0; i < n_bins; ++i) {
    double melf = freq_to_mel(i * hz_per_sbin);
    if (i < m->start_index || i > m->end_index) {
      m->band_mapper[i] = -2;
    } else {
      while (channel < n_channels && center[channel] < melf) ++channel;
      m->band_mapper[i] = channel - 1;
    }
  }
  for (int i = 0; i < n_bins; ++i) {
    channel = m->band_mapper[i];
    if (i < m->start_index || i > m->end_index) {
      m->weights[i] = 0.0;
    } else if (channel >= 0) {
      m->weights[i] = (center[channel + 1] - freq_to_mel(i * hz_per_sbin)) / (center[channel + 1] - center[channel]);
    } else {
      m->weights[i] = (center[0] - freq_to_mel(i * hz_per_sbin)) / (center[0] - mel_low);
    }
  }
  free(center);
  m->cosines = (double*)malloc(sizeof(double) * n_dct * n_channels);
  const double fnorm = sqrt(2.0 / n_channels);
  const double pi = atan(1.0) * 4.0;
  const double arg = pi / n_channels;
  for (int i = 0; i < n_dct; ++i)
    for (int j = 0; j < n_channels; ++j) m->cosines[i * n_channels + j] = fnorm * cos(i * arg * (j + 0.5));
  return m;
}

void orc_mfcc_free(orc_mfcc* m) {
  if (!m) return;
  free(m->band_mapper); free(m->weights); free(m->cosines); free(m);
}

void orc_mfcc_compute(const orc_mfcc* m, const float* power_spectrum, float* out) {
  double* working = (double*)calloc(m->n_channels, sizeof(double));
  if (m->n_bins > m->end_index) {
    for (int i = m->start_index; i <= m->end_index; ++i) {
      double spec_val = sqrt((double)power_spectrum[i]);
      double weighted = spec_val * m->weights[i];
      int channel = m->band_mapper[i];
      if (channel >= 0) working[channel] += weighted;
      channel++;
      if (channel < m->n_channels) working[channel] += spec_val - weighted;
    }
  }
  for (int i = 0; i < m->n_channels; ++i) {
    double val = working[i];
    if (val < 1e-12) val = 1e-12;
    working[i] = log(val);
  }
  for (int i = 0; i < m->n_dct; ++i) {
    double sum = 0.0;
    for (int j = 0; j < m->n_channels; ++j) sum += m->cosines[i * m->n_channels + j] * working[j];
    out[i] = (float)sum;
  }
  free(working);
}

/* ------------------------------------------------------------------ acoustic model */
/* y[n_out] = x[n_in] @ W[n_in, n_out] + b ; fp32 accumulate, k-sequential per output. */
static void dense_row(const float* x, const float* W, const float* b, int n_in, int n_out, float* y) {
  for (int j = 0; j < n_out; ++j) y[j] = 0.0f;
  for (int k = 0; k < n_in; ++k) {
    const float xv = x[k];
    if (xv == 0.0f) continue; /* adding +-0 products never changes a finite fp32 sum */
    const float* w = W + (size_t)k * n_out;
    for (int j = 0; j < n_out; ++j) y[j] += xv * w[j];
  }
  for (int j = 0; j < n_out; ++j) y[j] += b[j];
}
static void clipped_relu(float* y, int n, float clip) {
  for (int j = 0; j < n; ++j) {
    float v = y[j] > 0.0f ? y[j] : 0.0f;
    y[j] = v < clip ? v : clip;
  }
}
static float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

void orc_am_infer(const orc_am* am, const float* x, int n_steps, float* c, float* h, float* probs) {
  const int n_in = (2 * am->n_context + 1) * am->n_input;
  const int H = am->n_hidden, C = am->n_cell, K = am->n_classes;
  float* a = (float*)malloc(sizeof(float) * H);
  float* b = (float*)malloc(sizeof(float) * H);
  float* xh = (float*)malloc(sizeof(float) * (H + C));
  float* gates = (float*)malloc(sizeof(float) * 4 * C);
  float* logits = (float*)malloc(sizeof(float) * K);
  for (int t = 0; t < n_steps; ++t) {
    dense_row(x + (size_t)t * n_in, am->w1, am->b1, n_in, H, a); clipped_relu(a, H, am->relu_clip);
    dense_row(a, am->w2, am->b2, H, H, b); clipped_relu(b, H, am->relu_clip);
    dense_row(b, am->w3, am->b3, H, H, a); clipped_relu(a, H, am->relu_clip);
    memcpy(xh, a, sizeof(float) * H);
    memcpy(xh + H, h, sizeof(float) * C);
    dense_row(xh, am->lstm_kernel, am->lstm_bias, H + C, 4 * C, gates);
    for (int j = 0; j < C; ++j) {
      const float gi = gates[j], gj = gates[C + j], gf = gates[2 * C + j], go = gates[3 * C + j];
      const float cn = sigmoidf_(gf) * c[j] + sigmoidf_(gi) * tanhf(gj);
      c[j] = cn;
      h[j] = sigmoidf_(go) * tanhf(cn);
    }
    dense_row(h, am->w5, am->b5, C, H, a); clipped_relu(a, H, am->relu_clip);
    dense_row(a, am->w6, am->b6, H, K, logits);
    float mx = logits[0];
    for (int j = 1; j < K; ++j) if (logits[j] > mx) mx = logits[j];
    float sum = 0.0f;
    for (int j = 0; j < K; ++j) { logits[j] = expf(logits[j] - mx); sum += logits[j]; }
    for (int j = 0; j < K; ++j) probs[(size_t)t * K + j] = logits[j] / sum;
  }
  free(a); free(b); free(xh); free(gates); free(logits);
}

/* ------------------------------------------------------------------ hybrid int8 mode
 * What a DEFAULT Coqui export computes (config.py:616-622 export_quantize -> export.py:145-146 Optimize.DEFAULT):
 * TFLite "hybrid" FULLY_CONNECTED, tensorflow/lite/kernels/fully_connected.cc:435-503 (EvalHybridDense):
 *   weights: per-tensor symmetric int8, scale = max|w|/127 (quantize_weights);
 *   input row: PortableSymmetricQuantizeFloats (internal/reference/portable_tensor_utils.cc:51-70): scale = max|x|/127,
 *              q = TfLiteRound(x * 127/max|x|) (round half away from zero), clamped to +-127;
 *   output = bias; all-zero input rows skip the matmul (:455-461);
 *   result += (float)int32_dot * (input_scale * weight_scale)  (:138-161).
 * The LSTM's matmul quantises concat([x_t, h]) as ONE row (rnn_cell_impl.py:1060). */
static float tfl_round(float x) { return roundf(x); } /* TfLiteRound == std::round: half away from zero */

void orc_quantize_weights(const float* w, size_t n, int8_t* q, float* scale) {
  float range = 0.0f;
  for (size_t i = 0; i < n; ++i) { const float a = fabsf(w[i]); if (a > range) range = a; }
  if (range == 0.0f) { memset(q, 0, n); *scale = 1.0f; return; }
  *scale = range / 127.0f;
  const float inv = 127.0f / range;
  for (size_t i = 0; i < n; ++i) {
    int v = (int)tfl_round(w[i] * inv);
    q[i] = (int8_t)(v > 127 ? 127 : (v < -127 ? -127 : v));
  }
}

/* y[n_out] = b + hybrid(x[n_in] @ Wq[n_in, n_out]) */
static void hybrid_row(const float* x, const int8_t* wq, float wscale, const float* b, int n_in, int n_out, float* y,
                       int8_t* xq, int32_t* acc) {
  for (int j = 0; j < n_out; ++j) y[j] = b[j];
  float range = 0.0f;
  for (int i = 0; i < n_in; ++i) { const float a = fabsf(x[i]); if (a > range) range = a; }
  if (range == 0.0f) return; /* IsZeroVector shortcut */
  const float sf = range / 127.0f, inv = 127.0f / range;
  for (int i = 0; i < n_in; ++i) {
    int v = (int)tfl_round(x[i] * inv);
    xq[i] = (int8_t)(v > 127 ? 127 : (v < -127 ? -127 : v));
  }
  memset(acc, 0, sizeof(int32_t) * n_out);
  for (int i = 0; i < n_in; ++i) {
    const int32_t xi = xq[i];
    if (!xi) continue;
    const int8_t* wr = wq + (size_t)i * n_out;
    for (int j = 0; j < n_out; ++j) acc[j] += xi * wr[j];
  }
  const float scale = sf * wscale;
  for (int j = 0; j < n_out; ++j) y[j] += (float)acc[j] * scale;
}

void orc_am_infer_hybrid(const orc_am_q* q, const float* x, int n_steps, float* c, float* h, float* probs) {
  const orc_am* am = q->am;
  const int n_in = (2 * am->n_context + 1) * am->n_input;
  const int H = am->n_hidden, C = am->n_cell, K = am->n_classes;
  const int widest = 4 * C > H ? 4 * C : H;
  float* a = (float*)malloc(sizeof(float) * H);
  float* b = (float*)malloc(sizeof(float) * H);
  float* xh = (float*)malloc(sizeof(float) * (H + C));
  float* gates = (float*)malloc(sizeof(float) * 4 * C);
  float* logits = (float*)malloc(sizeof(float) * K);
  int8_t* xq = (int8_t*)malloc((size_t)(H + C > n_in ? H + C : n_in));
  int32_t* acc = (int32_t*)malloc(sizeof(int32_t) * widest);
  for (int t = 0; t < n_steps; ++t) {
    hybrid_row(x + (size_t)t * n_in, q->w1, q->s1, am->b1, n_in, H, a, xq, acc); clipped_relu(a, H, am->relu_clip);
    hybrid_row(a, q->w2, q->s2, am->b2, H, H, b, xq, acc); clipped_relu(b, H, am->relu_clip);
    hybrid_row(b, q->w3, q->s3, am->b3, H, H, a, xq, acc); clipped_relu(a, H, am->relu_clip);
    memcpy(xh, a, sizeof(float) * H);
    memcpy(xh + H, h, sizeof(float) * C);
    hybrid_row(xh, q->lstm_kernel, q->sk, am->lstm_bias, H + C, 4 * C, gates, xq, acc);
    for (int j = 0; j < C; ++j) {
      const float gi = gates[j], gj = gates[C + j], gf = gates[2 * C + j], go = gates[3 * C + j];
      const float cn = sigmoidf_(gf) * c[j] + sigmoidf_(gi) * tanhf(gj);
      c[j] = cn;
      h[j] = sigmoidf_(go) * tanhf(cn);
    }
    hybrid_row(h, q->w5, q->s5, am->b5, C, H, a, xq, acc); clipped_relu(a, H, am->relu_clip);
    hybrid_row(a, q->w6, q->s6, am->b6, H, K, logits, xq, acc);
    float mx = logits[0];
    for (int j = 1; j < K; ++j) if (logits[j] > mx) mx = logits[j];
    float sum = 0.0f;
    for (int j = 0; j < K; ++j) { logits[j] = expf(logits[j] - mx); sum += logits[j]; }
    for (int j = 0; j < K; ++j) probs[(size_t)t * K + j] = logits[j] / sum;
  }
  free(a); free(b); free(xh); free(gates); free(logits); free(xq); free(acc);
}

orc_am_q* orc_am_quantize(const orc_am* am) {
  orc_am_q* q = (orc_am_q*)calloc(1, sizeof(orc_am_q));
  const size_t n_in = (size_t)(2 * am->n_context + 1) * am->n_input, H = am->n_hidden, C = am->n_cell, K = am->n_classes;
  q->am = am;
  q->w1 = (int8_t*)malloc(n_in * H); orc_quantize_weights(am->w1, n_in * H, q->w1, &q->s1);
  q->w2 = (int8_t*)malloc(H * H); orc_quantize_weights(am->w2, H * H, q->w2, &q->s2);
  q->w3 = (int8_t*)malloc(H * H); orc_quantize_weights(am->w3, H * H, q->w3, &q->s3);
  q->lstm_kernel = (int8_t*)malloc((H + C) * 4 * C); orc_quantize_weights(am->lstm_kernel, (H + C) * 4 * C, q->lstm_kernel, &q->sk);
  q->w5 = (int8_t*)malloc(C * H); orc_quantize_weights(am->w5, C * H, q->w5, &q->s5);
  q->w6 = (int8_t*)malloc(H * K); orc_quantize_weights(am->w6, H * K, q->w6, &q->s6);
  return q;
}
void orc_am_q_free(orc_am_q* q) {
  if (!q) return;
  free(q->w1); free(q->w2); free(q->w3); free(q->lstm_kernel); free(q->w5); free(q->w6); free(q);
}

/* ------------------------------------------------------------------ streaming runtime */
typedef struct { float* d; size_t n, cap; } fvec;
static void fv_push(fvec* v, const float* src, size_t n) {
  if (v->n + n > v->cap) {
    size_t cap = v->cap ? v->cap : 1024;
    while (cap < v->n + n) cap *= 2;
    v->d = (float*)realloc(v->d, cap * sizeof(float)); v->cap = cap;
  }
  if (src) memcpy(v->d + v->n, src, n * sizeof(float)); else memset(v->d + v->n, 0, n * sizeof(float));
  v->n += n;
}
static void fv_shift_left(fvec* v, size_t k) { memmove(v->d, v->d + k, (v->n - k) * sizeof(float)); v->n -= k; }

struct orc_stream {
  const orc_am* am;
  orc_mfcc* mfcc;
  int win_len, win_step, n_steps, n_features, n_context, feats_per_step, n_classes;
  fvec audio, mfcc_buf, batch, probs, mfcc_all;
  float *c, *h;
  int n_windows;
};

orc_stream* orc_stream_new(const orc_am* am, int sample_rate, int win_len, int win_step, int n_steps) {
  orc_stream* s = (orc_stream*)calloc(1, sizeof(orc_stream));
  s->am = am;
  s->win_len = win_len; s->win_step = win_step; s->n_steps = n_steps;
  s->n_features = am ? am->n_input : 26;
  s->n_context = am ? am->n_context : 9;
  s->n_classes = am ? am->n_classes : 0;
  s->feats_per_step = (2 * s->n_context + 1) * s->n_features;
  /* feeding.py:51-72: upper = sample_rate/2, lower 20, 40 channels (audio_ops.cc:150-153), dct = n_input */
  s->mfcc = orc_mfcc_new(next_pow2(win_len) / 2 + 1, sample_rate, 20.0, sample_rate / 2, 40, s->n_features);
  fv_push(&s->mfcc_buf, NULL, (size_t)s->n_features * s->n_context); /* stt.cc:533 literal zeros */
  if (am) { s->c = (float*)calloc(am->n_cell, sizeof(float)); s->h = (float*)calloc(am->n_cell, sizeof(float)); }
  return s;
}
void orc_stream_free(orc_stream* s) {
  if (!s) return;
  orc_mfcc_free(s->mfcc);
  free(s->audio.d); free(s->mfcc_buf.d); free(s->batch.d); free(s->probs.d); free(s->mfcc_all.d);
  free(s->c); free(s->h); free(s);
}

static void process_batch(orc_stream* s, unsigned n_steps) { /* stt.cc:311-334 */
  if (s->am) {
    /* infer zero-pads to n_steps_ rows but only n_frames rows are read (tflitemodelstate.cc:381,396);
       the LSTM state DOES advance through the padding rows. */
    size_t full = (size_t)s->n_steps * s->feats_per_step;
    float* x = (float*)calloc(full, sizeof(float));
    memcpy(x, s->batch.d, s->batch.n * sizeof(float));
    float* p = (float*)malloc(sizeof(float) * (size_t)s->n_steps * s->n_classes);
    orc_am_infer(s->am, x, s->n_steps, s->c, s->h, p);
    fv_push(&s->probs, p, (size_t)n_steps * s->n_classes);
    free(x); free(p);
  }
}
static void process_mfcc_window(orc_stream* s) { /* stt.cc:292-309 */
  s->n_windows++;
  fv_push(&s->batch, s->mfcc_buf.d, s->feats_per_step);
  if (s->batch.n == (size_t)s->n_steps * s->feats_per_step) { process_batch(s, s->n_steps); s->batch.n = 0; }
}
static void push_mfcc(orc_stream* s, const float* f) { /* stt.cc:272-290 */
  fv_push(&s->mfcc_buf, f, s->n_features);
  if (s->mfcc_buf.n == (size_t)s->feats_per_step) { process_mfcc_window(s); fv_shift_left(&s->mfcc_buf, s->n_features); }
}
static void process_audio_window(orc_stream* s) { /* stt.cc:226-234 + compute_mfcc */
  float* spec = (float*)malloc(sizeof(float) * s->mfcc->n_bins);
  float* feat = (float*)malloc(sizeof(float) * s->n_features);
  orc_spectrogram_frame(s->audio.d, (int)s->audio.n, s->win_len, spec);
  orc_mfcc_compute(s->mfcc, spec, feat);
  fv_push(&s->mfcc_all, feat, s->n_features);
  push_mfcc(s, feat);
  free(spec); free(feat);
}
void orc_stream_feed(orc_stream* s, const int16_t* pcm, unsigned n) { /* stt.cc:105-128 */
  while (n > 0) {
    while (n > 0 && s->audio.n < (size_t)s->win_len) {
      float v = (float)(*pcm) * (1.0f / (1 << 15));
      fv_push(&s->audio, &v, 1);
      ++pcm; --n;
    }
    if (s->audio.n == (size_t)s->win_len) { process_audio_window(s); fv_shift_left(&s->audio, s->win_step); }
  }
}
void orc_stream_flush(orc_stream* s, int add_zero) { /* stt.cc:236-254 */
  process_audio_window(s);
  if (add_zero) {
    float* z = (float*)calloc(s->n_features, sizeof(float));
    for (int i = 0; i < s->n_context; ++i) push_mfcc(s, z);
    free(z);
  }
  if (s->batch.n > 0) { process_batch(s, (unsigned)(s->batch.n / s->feats_per_step)); s->batch.n = 0; }
}
int orc_stream_timesteps(const orc_stream* s) { return s->n_windows; }
int orc_stream_emitted(const orc_stream* s) { return s->n_classes ? (int)(s->probs.n / s->n_classes) : 0; }
const float* orc_stream_probs(const orc_stream* s) { return s->probs.d; }
int orc_stream_frames(const orc_stream* s) { return (int)(s->mfcc_all.n / s->n_features); }
const float* orc_stream_mfcc(const orc_stream* s) { return s->mfcc_all.d; }
