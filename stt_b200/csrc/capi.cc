// STT_* / STTX_* entry points (include/stt_capi.h).  Host-side restatement of native_client/stt.cc and
// native_client/modelstate.cc over the CUDA engine; see the header for the per-function reference lines.
#include "../../include/stt_capi.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include "engine.h"
#include "model_file.h"

#define STT_B200_VERSION "1.4.0-b200.r1"

using stteng::Decoded;

// ------------------------------------------------------------------------------------------------ state
struct ModelState {  // native_client/modelstate.h:13-27
  stteng::Engine* engine = nullptr;
  unsigned int beam_width_ = 0;
  std::map<std::string, float> hot_words_;
  std::vector<stteng::Batch*> stream_pool;  // idle single-stream device contexts
  std::set<struct StreamingState*> live_streams;  // streams not yet finished / freed (orphaned by STT_FreeModel)
  bool warned_hot_words = false;
  // device context of the last STTX_SpeechToTextBatch call, kept while the next call fits in it (a context for 256 x 10 s
  // is ~13 GB of device buffers: allocating it per call would cost more than the transcription)
  struct STTX_Batch* oneshot = nullptr;
  unsigned int oneshot_utts = 0, oneshot_samples = 0;
};

struct StreamingState {  // native_client/stt.cc:60-95
  ModelState* model_ = nullptr;
  stteng::Batch* dev = nullptr;
  std::vector<int16_t> audio_buffer_;  // int16 here; the x(1/32768) float conversion (stt.cc:113) happens on device
  std::vector<int16_t> q_windows;      // analysis windows not yet on the device
  std::vector<int> q_valid;
  int q_zero_frames = 0;
  bool keep_emissions_ = false;
  std::vector<double> probs_;          // rows of the last processBatch (stt.cc:328-330)
  bool failed = false;
};

struct STTX_Batch {
  ModelState* model = nullptr;
  stteng::Batch* dev = nullptr;
  unsigned int beam_cap = 0;
  std::vector<std::vector<Decoded>> results;
};

namespace {

constexpr int kStreamChunk = 512;  // timesteps a stream context can run in one device pass (multiple of n_steps)

// Capacity of a stream's decoder arena in seconds of audio BETWEEN two garbage collections (engine.cu stream_compact):
// streams are unbounded in length, the arena only has to hold the live trie plus one chunk of new nodes.  16 s at beam
// 500 is ~45 MB per stream context.
int stream_arena_seconds() {
  const char* s = getenv("STT_B200_STREAM_ARENA_SECONDS");
  int v = s ? atoi(s) : 16;
  return v >= 12 ? v : 12;   // at least one kStreamChunk (512 timesteps = 10.24 s) of growth
}

std::string decode_tokens(const sttmodel::HostModel& m, const std::vector<uint32_t>& tokens) {
  std::string out;  // Alphabet::Decode, alphabet.cc:213-221
  for (uint32_t t : tokens)
    if (t < m.labels.size()) out += m.labels[t];
  return out;
}

char* dup_string(const std::string& s) { return strdup(s.c_str()); }

// ModelState::decode_metadata, modelstate.cc:39-76
Metadata* make_metadata(const ModelState* ms, const std::vector<Decoded>& out) {
  const sttmodel::HostModel& m = stteng::engine_model(ms->engine);
  const unsigned int num_returned = (unsigned int)out.size();
  CandidateTranscript* transcripts = (CandidateTranscript*)malloc(sizeof(CandidateTranscript) * (num_returned ? num_returned : 1));
  for (unsigned int i = 0; i < num_returned; ++i) {
    const size_t n = out[i].tokens.size();
    TokenMetadata* tokens = (TokenMetadata*)malloc(sizeof(TokenMetadata) * (n ? n : 1));
    for (size_t j = 0; j < n; ++j) {
      const unsigned int ts = j < out[i].timesteps.size() ? out[i].timesteps[j] : 0;
      const uint32_t tok = out[i].tokens[j];
      TokenMetadata token{
          strdup(tok < m.labels.size() ? m.labels[tok].c_str() : ""),
          ts,
          ts * ((float)m.win_step / m.sample_rate),
      };
      memcpy(&tokens[j], &token, sizeof(TokenMetadata));
    }
    CandidateTranscript transcript{tokens, (unsigned int)n, out[i].confidence};
    memcpy(&transcripts[i], &transcript, sizeof(CandidateTranscript));
  }
  Metadata* ret = (Metadata*)malloc(sizeof(Metadata));
  Metadata metadata{transcripts, num_returned, NULL};
  memcpy(ret, &metadata, sizeof(Metadata));
  return ret;
}

// stt.cc:139-171: attach the last batch's emissions
Metadata* attach_emissions(const StreamingState* s, Metadata* m) {
  const sttmodel::HostModel& hm = stteng::engine_model(s->model_->engine);
  const size_t alphabet_size = hm.labels.size();
  const int num_timesteps = (int)(s->probs_.size() / (alphabet_size + 1));
  AcousticModelEmissions* emissions = (AcousticModelEmissions*)malloc(sizeof(AcousticModelEmissions));
  emissions->num_symbols = (int)alphabet_size;
  emissions->num_timesteps = num_timesteps;
  emissions->symbols = (const char**)malloc(sizeof(char*) * (alphabet_size + 1));
  for (size_t i = 0; i < alphabet_size; i++) emissions->symbols[i] = strdup(hm.labels[i].c_str());
  emissions->symbols[alphabet_size] = strdup("\t");
  double* probs = (double*)malloc(sizeof(double) * (alphabet_size + 1) * (num_timesteps ? num_timesteps : 1));
  memcpy(probs, s->probs_.data(), sizeof(double) * (alphabet_size + 1) * num_timesteps);
  emissions->emissions = probs;
  Metadata* ret = (Metadata*)malloc(sizeof(Metadata));
  Metadata metadata{m->transcripts, m->num_transcripts, emissions};
  memcpy(ret, &metadata, sizeof(Metadata));
  free(m);
  return ret;
}

int create_model_impl(const uint8_t* data, size_t size, const char* path, ModelState** retval) {
  *retval = nullptr;
  // CI greps these two lines on stderr (ci_scripts/asserts.sh:284-321; stt.cc:344-345)
  fprintf(stderr, "TensorFlow: none (stt_b200 CUDA sm_100a backend)\n");
  fprintf(stderr, " Coqui STT: %s\n", STT_B200_VERSION);
  if ((path && !strlen(path)) || (!path && !size)) {
    fprintf(stderr, "No model specified, cannot continue.\n");
    return STT_ERR_NO_MODEL;
  }
  sttmodel::HostModel hm;
  int err = path ? sttmodel::load_from_file(path, &hm) : sttmodel::load_from_buffer(data, size, &hm);
  if (err) return err;
  std::string why;
  stteng::Engine* e = stteng::engine_create(hm, &why);
  if (!e) {
    fprintf(stderr, "Could not create the CUDA engine: %s\n", why.c_str());
    return STT_ERR_FAIL_CREATE_MODEL;
  }
  ModelState* ms = new ModelState();
  ms->engine = e;
  ms->beam_width_ = hm.beam_width;
  *retval = ms;
  return STT_ERR_OK;
}

int enable_scorer_impl(ModelState* ms, const uint8_t* bytes, size_t n) {
  int err = stteng::engine_set_scorer(ms->engine, bytes, n);
  if (err) return STT_ERR_INVALID_SCORER;  // stt.cc:428-430
  // live streams keep the scorer object they captured at STT_CreateStream (stt.cc:542-547 passes the shared_ptr to
  // DecoderState::init); streams created from now on get the new one
  return STT_ERR_OK;
}

stteng::Batch* acquire_stream_ctx(ModelState* ms) {
  if (!ms->stream_pool.empty()) {
    stteng::Batch* b = ms->stream_pool.back();
    ms->stream_pool.pop_back();
    return b;
  }
  const sttmodel::HostModel& m = stteng::engine_model(ms->engine);
  const int max_samples = (kStreamChunk + 40) * (int)m.win_step + (int)m.win_len;
  const int dec_T = stream_arena_seconds() * (int)(m.sample_rate / m.win_step);
  std::string why;
  stteng::Batch* b = stteng::batch_create(ms->engine, 1, max_samples, (int)std::max(1u, ms->beam_width_), dec_T, &why);
  if (!b) fprintf(stderr, "Could not allocate streaming state: %s\n", why.c_str());
  return b;
}

// ---- StreamingState member functions of stt.cc, over the device context
bool stream_drain(StreamingState* s, bool run_partial, bool pad_partial) {
  const sttmodel::HostModel& m = stteng::engine_model(s->model_->engine);
  const int n_steps = (int)m.n_steps, ctx2 = 2 * (int)m.n_context, win = (int)m.win_len;
  size_t w_done = 0;
  const size_t n_win = s->q_valid.size();
  int zeros_left = s->q_zero_frames;
  int dev_frames = -1;
  auto run_full = [&](int frames_on_dev) -> int {
    int avail = frames_on_dev - ctx2;
    int runnable = avail >= n_steps ? (avail / n_steps) * n_steps : 0;
    if (runnable > kStreamChunk) runnable = kStreamChunk;
    if (runnable > 0) {
      if (stteng::batch_stream_run(s->dev, runnable, 0, s->keep_emissions_)) return -1;
      if (s->keep_emissions_) {
        std::vector<double> all;
        int rows = 0;
        stteng::batch_stream_last_probs(s->dev, &all, &rows);
        const size_t C = m.n_classes;
        s->probs_.assign(all.begin() + (size_t)(rows - n_steps) * C, all.end());
      }
    }
    return runnable;
  };
  dev_frames = stteng::batch_stream_frames(s->dev);  // frames (incl. leading context) currently on the device
  for (;;) {
    // push as much as fits: keep live frames <= kStreamChunk + ctx2
    int room = kStreamChunk + ctx2 - dev_frames;
    int push_w = (int)std::min<size_t>(n_win - w_done, (size_t)std::max(room, 0));
    int push_z = std::min(zeros_left, std::max(room - push_w, 0));
    if (push_w > 0 || push_z > 0) {
      if (stteng::batch_stream_push_windows(s->dev, s->q_windows.data() + w_done * win, s->q_valid.data() + w_done, push_w,
                                            (w_done + push_w == n_win) ? push_z : 0))
        return false;
      if (w_done + push_w != n_win) push_z = 0;
      w_done += push_w;
      zeros_left -= push_z;
      dev_frames += push_w + push_z;
    }
    int ran = run_full(dev_frames);
    if (ran < 0) return false;
    dev_frames -= ran;
    if (w_done == n_win && zeros_left == 0 && ran == 0) break;
    if (push_w == 0 && push_z == 0 && ran == 0) break;  // no progress possible
  }
  s->q_windows.clear();
  s->q_valid.clear();
  s->q_zero_frames = 0;
  if (run_partial) {
    const int avail = dev_frames - ctx2;
    if (avail > 0) {  // stt.cc:250-253: batch_buffer_.size() > 0
      const int pad = pad_partial ? n_steps - avail : 0;
      if (stteng::batch_stream_run(s->dev, avail, pad, s->keep_emissions_)) return false;
      if (s->keep_emissions_) {
        std::vector<double> all;
        int rows = 0;
        stteng::batch_stream_last_probs(s->dev, &all, &rows);
        s->probs_ = all;
      }
    }
  }
  return true;
}

void stream_feed(StreamingState* s, const short* buffer, unsigned int buffer_size) {  // stt.cc:105-128
  if (!s->model_) return;  // orphaned by STT_FreeModel
  const sttmodel::HostModel& m = stteng::engine_model(s->model_->engine);
  const size_t win = m.win_len, step = m.win_step;
  while (buffer_size > 0) {
    const size_t take = std::min<size_t>(buffer_size, win - s->audio_buffer_.size());
    s->audio_buffer_.insert(s->audio_buffer_.end(), buffer, buffer + take);
    buffer += take;
    buffer_size -= (unsigned int)take;
    if (s->audio_buffer_.size() == win) {
      s->q_windows.insert(s->q_windows.end(), s->audio_buffer_.begin(), s->audio_buffer_.end());
      s->q_valid.push_back((int)win);
      s->audio_buffer_.erase(s->audio_buffer_.begin(), s->audio_buffer_.begin() + step);  // shift_buffer_left
    }
  }
  if (!stream_drain(s, false, false)) s->failed = true;
}

void stream_flush(StreamingState* s, bool add_zero_mfcc_vectors) {  // stt.cc:236-254
  if (!s->model_) return;
  const sttmodel::HostModel& m = stteng::engine_model(s->model_->engine);
  // processAudioWindow(audio_buffer_): the partial window, zero padded (tflitemodelstate.cc:343-355); not consumed
  std::vector<int16_t> w(m.win_len, 0);
  std::copy(s->audio_buffer_.begin(), s->audio_buffer_.end(), w.begin());
  s->q_windows.insert(s->q_windows.end(), w.begin(), w.end());
  s->q_valid.push_back((int)s->audio_buffer_.size());
  if (add_zero_mfcc_vectors) s->q_zero_frames += (int)m.n_context;
  // a non-final flush advances the LSTM through the zero padding of the partial batch; after the final flush the
  // stream is freed, so that state is unobservable and the padding pass is skipped
  if (!stream_drain(s, true, !add_zero_mfcc_vectors)) s->failed = true;
}

std::vector<Decoded> stream_decode(const StreamingState* s, unsigned int num_results) {
  std::vector<Decoded> out;
  if (stteng::batch_stream_decode(s->dev, (int)std::max(1u, num_results), &out)) out.clear();
  return out;
}

char* stream_decode_text(const StreamingState* s) {  // ModelState::decode, modelstate.cc:32-37
  if (!s->model_) return nullptr;
  std::vector<Decoded> out = stream_decode(s, 1);
  if (out.empty() || s->failed) return nullptr;
  return dup_string(decode_tokens(stteng::engine_model(s->model_->engine), out[0].tokens));
}

Metadata* stream_decode_metadata(const StreamingState* s, unsigned int num_results) {
  if (!s->model_) return nullptr;
  std::vector<Decoded> out = stream_decode(s, num_results);
  if (s->failed) return nullptr;
  Metadata* m = make_metadata(s->model_, out);
  if (s->keep_emissions_) m = attach_emissions(s, m);
  return m;
}

int create_stream_impl(ModelState* ms, StreamingState** retval, bool keep_emissions) {
  *retval = nullptr;
  std::unique_ptr<StreamingState> ctx(new StreamingState());
  ctx->model_ = ms;
  ctx->keep_emissions_ = keep_emissions;
  ctx->dev = acquire_stream_ctx(ms);
  if (!ctx->dev) return STT_ERR_FAIL_CREATE_STREAM;
  // beam width is snapshotted per stream (stt.cc:542-547); a context built for a smaller beam is replaced
  if (stteng::batch_stream_reset(ctx->dev, (int)std::max(1u, ms->beam_width_)) != 0) {
    stteng::batch_destroy(ctx->dev);
    const sttmodel::HostModel& m = stteng::engine_model(ms->engine);
    const int max_samples = (kStreamChunk + 40) * (int)m.win_step + (int)m.win_len;
    const int dec_T = stream_arena_seconds() * (int)(m.sample_rate / m.win_step);
    std::string why;
    ctx->dev = stteng::batch_create(ms->engine, 1, max_samples, (int)std::max(1u, ms->beam_width_), dec_T, &why);
    if (!ctx->dev || stteng::batch_stream_reset(ctx->dev, (int)std::max(1u, ms->beam_width_)) != 0) {
      if (ctx->dev) stteng::batch_destroy(ctx->dev);
      return STT_ERR_FAIL_CREATE_STREAM;
    }
  }
  {  // hot words are snapshotted per stream (stt.cc:542-547)
    std::vector<std::string> w;
    std::vector<float> bo;
    for (const auto& kv : ms->hot_words_) { w.push_back(kv.first); bo.push_back(kv.second); }
    stteng::batch_set_hot_words(ctx->dev, w, bo);
  }
  ctx->audio_buffer_.reserve(stteng::engine_model(ms->engine).win_len);
  ms->live_streams.insert(ctx.get());
  *retval = ctx.release();
  return STT_ERR_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------ PART 1
extern "C" {

int STT_CreateModel(const char* aModelPath, ModelState** retval) {
  if (!aModelPath) {
    *retval = nullptr;
    return STT_ERR_NO_MODEL;
  }
  return create_model_impl(nullptr, 0, aModelPath, retval);
}

int STT_CreateModelFromBuffer(const char* aModelBuffer, unsigned int aBufferSize, ModelState** retval) {
  return create_model_impl(reinterpret_cast<const uint8_t*>(aModelBuffer), aBufferSize, nullptr, retval);
}

unsigned int STT_GetModelBeamWidth(const ModelState* aCtx) { return aCtx->beam_width_; }

int STT_SetModelBeamWidth(ModelState* aCtx, unsigned int aBeamWidth) {
  aCtx->beam_width_ = aBeamWidth;
  return 0;
}

int STT_GetModelSampleRate(const ModelState* aCtx) { return (int)stteng::engine_model(aCtx->engine).sample_rate; }

void STT_FreeModel(ModelState* ctx) {
  if (!ctx) return;
  // The reference's STT_FreeStream never touches the model, so a client may free the model first: orphan such streams
  // (their device context dies with the model; STT_FreeStream then only deletes the host-side struct).
  for (StreamingState* s : ctx->live_streams) {
    if (s->dev) stteng::batch_destroy(s->dev);
    s->dev = nullptr;
    s->model_ = nullptr;
    s->failed = true;
  }
  for (stteng::Batch* b : ctx->stream_pool) stteng::batch_destroy(b);
  if (ctx->oneshot) STTX_BatchFree(ctx->oneshot);
  stteng::engine_destroy(ctx->engine);
  delete ctx;
}

int STT_EnableExternalScorer(ModelState* aCtx, const char* aScorerPath) {
  std::ifstream f(aScorerPath, std::ios::binary);
  if (!f) return STT_ERR_INVALID_SCORER;
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string bytes = ss.str();
  return enable_scorer_impl(aCtx, reinterpret_cast<const uint8_t*>(bytes.data()), bytes.size());
}

int STT_EnableExternalScorerFromBuffer(ModelState* aCtx, const char* aScorerBuffer, unsigned int aBufferSize) {
  return enable_scorer_impl(aCtx, reinterpret_cast<const uint8_t*>(aScorerBuffer), aBufferSize);
}

int STT_AddHotWord(ModelState* aCtx, const char* word, float boost) {
  if (stteng::engine_has_scorer(aCtx->engine)) {
    const size_t before = aCtx->hot_words_.size();
    aCtx->hot_words_.insert(std::pair<std::string, float>(word, boost));
    if (before == aCtx->hot_words_.size()) return STT_ERR_FAIL_INSERT_HOTWORD;
    return STT_ERR_OK;
  }
  return STT_ERR_SCORER_NOT_ENABLED;
}

int STT_EraseHotWord(ModelState* aCtx, const char* word) {
  if (stteng::engine_has_scorer(aCtx->engine)) {
    const size_t before = aCtx->hot_words_.size();
    aCtx->hot_words_.erase(word);
    if (before == aCtx->hot_words_.size()) return STT_ERR_FAIL_ERASE_HOTWORD;
    return STT_ERR_OK;
  }
  return STT_ERR_SCORER_NOT_ENABLED;
}

int STT_ClearHotWords(ModelState* aCtx) {
  if (stteng::engine_has_scorer(aCtx->engine)) {
    aCtx->hot_words_.clear();
    return STT_ERR_OK;
  }
  return STT_ERR_SCORER_NOT_ENABLED;
}

int STT_DisableExternalScorer(ModelState* aCtx) {
  if (stteng::engine_has_scorer(aCtx->engine)) {
    stteng::engine_clear_scorer(aCtx->engine);
    return STT_ERR_OK;
  }
  return STT_ERR_SCORER_NOT_ENABLED;
}

int STT_SetScorerAlphaBeta(ModelState* aCtx, float aAlpha, float aBeta) {
  if (stteng::engine_has_scorer(aCtx->engine)) {
    stteng::engine_set_alpha_beta(aCtx->engine, aAlpha, aBeta);
    return STT_ERR_OK;
  }
  return STT_ERR_SCORER_NOT_ENABLED;
}

int STT_CreateStream(ModelState* aCtx, StreamingState** retval) { return create_stream_impl(aCtx, retval, false); }

void STT_FeedAudioContent(StreamingState* aSctx, const short* aBuffer, unsigned int aBufferSize) {
  stream_feed(aSctx, aBuffer, aBufferSize);
}

char* STT_IntermediateDecode(const StreamingState* aSctx) { return stream_decode_text(aSctx); }

Metadata* STT_IntermediateDecodeWithMetadata(const StreamingState* aSctx, unsigned int aNumResults) {
  return stream_decode_metadata(aSctx, aNumResults);
}

char* STT_IntermediateDecodeFlushBuffers(StreamingState* aSctx) {
  stream_flush(aSctx, false);
  return stream_decode_text(aSctx);
}

Metadata* STT_IntermediateDecodeWithMetadataFlushBuffers(StreamingState* aSctx, unsigned int aNumResults) {
  stream_flush(aSctx, false);
  return stream_decode_metadata(aSctx, aNumResults);
}

char* STT_FinishStream(StreamingState* aSctx) {
  stream_flush(aSctx, true);
  char* str = stream_decode_text(aSctx);
  STT_FreeStream(aSctx);
  return str;
}

Metadata* STT_FinishStreamWithMetadata(StreamingState* aSctx, unsigned int aNumResults) {
  stream_flush(aSctx, true);
  Metadata* result = stream_decode_metadata(aSctx, aNumResults);
  STT_FreeStream(aSctx);
  return result;
}

char* STT_SpeechToText(ModelState* aCtx, const short* aBuffer, unsigned int aBufferSize) {
  StreamingState* ctx;
  if (STT_CreateStream(aCtx, &ctx) != STT_ERR_OK) return nullptr;
  STT_FeedAudioContent(ctx, aBuffer, aBufferSize);
  return STT_FinishStream(ctx);
}

Metadata* STT_SpeechToTextWithMetadata(ModelState* aCtx, const short* aBuffer, unsigned int aBufferSize,
                                       unsigned int aNumResults) {
  StreamingState* ctx;
  if (STT_CreateStream(aCtx, &ctx) != STT_ERR_OK) return nullptr;
  STT_FeedAudioContent(ctx, aBuffer, aBufferSize);
  return STT_FinishStreamWithMetadata(ctx, aNumResults);
}

Metadata* STT_SpeechToTextWithEmissions(ModelState* aCtx, const short* aBuffer, unsigned int aBufferSize,
                                        unsigned int aNumResults) {
  StreamingState* ctx;
  if (create_stream_impl(aCtx, &ctx, true) != STT_ERR_OK) return nullptr;
  STT_FeedAudioContent(ctx, aBuffer, aBufferSize);
  return STT_FinishStreamWithMetadata(ctx, aNumResults);
}

void STT_FreeStream(StreamingState* aSctx) {
  if (!aSctx) return;
  if (aSctx->model_) {
    aSctx->model_->live_streams.erase(aSctx);
    if (aSctx->dev) {
      stteng::batch_release_scorer(aSctx->dev);            // the stream's handle on its scorer ends with the stream
      // keep a few device contexts for the next streams (each pins tens of MB); the rest go back to the driver
      if (aSctx->model_->stream_pool.size() < 4) aSctx->model_->stream_pool.push_back(aSctx->dev);
      else stteng::batch_destroy(aSctx->dev);
    }
  }
  delete aSctx;
}

void STT_FreeMetadata(Metadata* m) {
  if (!m) return;
  for (unsigned int i = 0; i < m->num_transcripts; ++i) {
    for (unsigned int j = 0; j < m->transcripts[i].num_tokens; ++j) free((void*)m->transcripts[i].tokens[j].text);
    free((void*)m->transcripts[i].tokens);
  }
  free((void*)m->transcripts);
  if (m->emissions) {
    if (m->emissions->symbols) {
      for (int i = 0; i < m->emissions->num_symbols + 1; i++) free((void*)m->emissions->symbols[i]);
      free((void*)m->emissions->symbols);
    }
    if (m->emissions->emissions) free((void*)m->emissions->emissions);
    free((void*)m->emissions);
  }
  free(m);
}

void STT_FreeString(char* str) { free(str); }

char* STT_Version(void) { return strdup(STT_B200_VERSION); }

char* STT_ErrorCodeToErrorMessage(int aErrorCode) {
#define STT_RETURN_MESSAGE(NAME, VALUE, DESC) \
  case NAME:                                  \
    return strdup(DESC);
  switch (aErrorCode) {
    STT_FOR_EACH_ERROR(STT_RETURN_MESSAGE)
    default:
      return strdup("Unknown error, please make sure you are using the correct native binary.");
  }
#undef STT_RETURN_MESSAGE
}

// ------------------------------------------------------------------------------------------------ PART 2
int STTX_BatchCreate(ModelState* aCtx, unsigned int aMaxUtterances, unsigned int aMaxSamples, STTX_Batch** retval) {
  *retval = nullptr;
  if (aMaxUtterances < 1 || aMaxUtterances > 256) return STT_ERR_INVALID_SHAPE;
  std::string why;
  stteng::Batch* dev = stteng::batch_create(aCtx->engine, (int)aMaxUtterances, (int)aMaxSamples,
                                            (int)std::max(1u, aCtx->beam_width_), 0, &why);
  if (!dev) {
    fprintf(stderr, "STTX_BatchCreate: %s\n", why.c_str());
    return STT_ERR_FAIL_CREATE_STREAM;
  }
  STTX_Batch* b = new STTX_Batch();
  b->model = aCtx;
  b->dev = dev;
  b->beam_cap = std::max(1u, aCtx->beam_width_);
  *retval = b;
  return STT_ERR_OK;
}

void STTX_BatchFree(STTX_Batch* b) {
  if (!b) return;
  stteng::batch_destroy(b->dev);
  delete b;
}

short* STTX_BatchHostBuffer(STTX_Batch* b, unsigned int u) { return stteng::batch_host_pcm(b->dev, (int)u); }

int STTX_BatchUpload(STTX_Batch* b, const short* const* aBuffers, const unsigned int* aBufferSizes, unsigned int n) {
  return stteng::batch_upload(b->dev, aBuffers, aBufferSizes, (int)n) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
int STTX_BatchForward(STTX_Batch* b) { return stteng::batch_forward(b->dev) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK; }
int STTX_BatchDecode(STTX_Batch* b, unsigned int aNumResults) {
  const unsigned int beam = std::max(1u, b->model->beam_width_);
  if (beam > b->beam_cap) return STT_ERR_INVALID_SHAPE;
  {
    std::vector<std::string> w;
    std::vector<float> bo;
    for (const auto& kv : b->model->hot_words_) { w.push_back(kv.first); bo.push_back(kv.second); }
    stteng::batch_set_hot_words(b->dev, w, bo);
  }
  return stteng::batch_decode(b->dev, (int)beam, (int)std::max(1u, aNumResults)) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
int STTX_BatchFetch(STTX_Batch* b) { return stteng::batch_fetch(b->dev, &b->results) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK; }
int STTX_BatchNumResults(STTX_Batch* b, unsigned int u) { return u < b->results.size() ? (int)b->results[u].size() : -1; }
char* STTX_BatchTranscript(STTX_Batch* b, unsigned int u, unsigned int r) {
  if (u >= b->results.size() || r >= b->results[u].size()) return nullptr;
  return dup_string(decode_tokens(stteng::engine_model(b->model->engine), b->results[u][r].tokens));
}
int STTX_BatchTokens(STTX_Batch* b, unsigned int u, unsigned int r, unsigned int* tokens, unsigned int* timesteps,
                     unsigned int cap, double* confidence) {
  if (u >= b->results.size() || r >= b->results[u].size()) return -1;
  const Decoded& d = b->results[u][r];
  const size_t n = std::min<size_t>(cap, d.tokens.size());
  for (size_t i = 0; i < n; ++i) {
    tokens[i] = d.tokens[i];
    timesteps[i] = i < d.timesteps.size() ? d.timesteps[i] : 0;
  }
  if (confidence) *confidence = d.confidence;
  return (int)d.tokens.size();
}
int STTX_BatchGetTimings(STTX_Batch* b, STTX_Timings* out) {
  const stteng::StageTimes& t = stteng::batch_times(b->dev);
  out->h2d = t.h2d; out->mfcc = t.mfcc; out->dense123 = t.dense123; out->lstm_in = t.lstm_in; out->lstm = t.lstm;
  out->dense56 = t.dense56; out->decode = t.decode; out->d2h = t.d2h;
  out->total = t.h2d + t.mfcc + t.dense123 + t.lstm_in + t.lstm + t.dense56 + t.decode + t.d2h;
  return STT_ERR_OK;
}
long long STTX_BatchKernelLaunches(STTX_Batch* b) { return stteng::batch_kernel_launches(b->dev); }
int STTX_BatchSetInstrumented(STTX_Batch* b, int aOn) {
  stteng::batch_set_instrumented(b->dev, aOn != 0);
  return STT_ERR_OK;
}
int STTX_BatchPhaseCycles(STTX_Batch* b, unsigned long long* out8) {
  return stteng::batch_phase_cycles(b->dev, out8) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
int STTX_BatchLstmProfile(STTX_Batch* b, unsigned long long* out3) {
  return stteng::batch_lstm_profile(b->dev, out3) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
int STTX_BatchLmStats(STTX_Batch* b, unsigned long long* words_scored, unsigned long long* lm_calls) {
  return stteng::batch_lm_stats(b->dev, words_scored, lm_calls) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
int STTX_BatchSetCutoff(STTX_Batch* b, double cutoff_prob, unsigned int cutoff_top_n) {
  return stteng::batch_set_cutoff(b->dev, cutoff_prob, (int)cutoff_top_n) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
int STTX_BatchDecoderScalars(STTX_Batch* b, unsigned long long* out16) {
  return stteng::batch_decoder_scalars(b->dev, out16) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
int STTX_BatchTimesteps(STTX_Batch* b, unsigned int u) { return stteng::batch_T(b->dev, (int)u); }
int STTX_BatchCopyFeatures(STTX_Batch* b, unsigned int u, float* out) { return stteng::batch_copy_features(b->dev, (int)u, out); }
int STTX_BatchCopyProbs(STTX_Batch* b, unsigned int u, float* out) { return stteng::batch_copy_probs(b->dev, (int)u, out); }
int STTX_BatchSetProbs(STTX_Batch* b, const float* probs, const int* T, unsigned int n, unsigned int T_stride) {
  return stteng::batch_set_probs(b->dev, probs, T, (int)n, (int)T_stride) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
int STTX_BatchSetProbs64(STTX_Batch* b, const double* probs, const int* T, unsigned int n, unsigned int T_stride) {
  return stteng::batch_set_probs64(b->dev, probs, T, (int)n, (int)T_stride) ? STT_ERR_FAIL_RUN_SESS : STT_ERR_OK;
}
#ifdef STT_B200_DEV_HOOKS
int STTX_DebugPairLayout(int M, float* out) { return stteng::debug_pair_layout(M, out); }
int STTX_DebugGemm(int M, int N, int K, const unsigned short* a_f16, const unsigned short* w_f16, const float* bias,
                   int epilogue, float relu_clip, void* out, float* ms) {
  return stteng::debug_gemm(M, N, K, a_f16, w_f16, bias, epilogue, relu_clip, out, ms);
}
#endif
int STTX_ModelInfo(const ModelState* aCtx, unsigned int* n_classes, unsigned int* n_input, unsigned int* n_hidden,
                   unsigned int* n_steps, unsigned int* n_sms) {
  const sttmodel::HostModel& m = stteng::engine_model(aCtx->engine);
  if (n_classes) *n_classes = m.n_classes;
  if (n_input) *n_input = m.n_input;
  if (n_hidden) *n_hidden = m.n_hidden;
  if (n_steps) *n_steps = m.n_steps;
  if (n_sms) *n_sms = (unsigned int)stteng::engine_num_sms(aCtx->engine);
  return STT_ERR_OK;
}

// Model-file inspection without a device: parses either container (TFLite flatbuffer / .sttw) exactly as STT_CreateModel
// does and reports what was read.  aInfo[12] = sample_rate, win_len, win_step, n_input, n_context, n_hidden, n_cell,
// n_classes, n_steps, beam_width, space_label, number of labels.
int STTX_InspectModel(const char* aModelBuffer, unsigned int aBufferSize, unsigned int* aInfo, float* aReluClip) {
  sttmodel::HostModel hm;
  const int err = sttmodel::load_from_buffer(reinterpret_cast<const uint8_t*>(aModelBuffer), aBufferSize, &hm);
  if (err) return err;
  const unsigned int v[12] = {hm.sample_rate, hm.win_len, hm.win_step, hm.n_input, hm.n_context, hm.n_hidden, hm.n_cell,
                              hm.n_classes, hm.n_steps, hm.beam_width, hm.space_label, (unsigned int)hm.labels.size()};
  if (aInfo) memcpy(aInfo, v, sizeof(v));
  if (aReluClip) *aReluClip = hm.relu_clip;
  return STT_ERR_OK;
}
// One tensor of the parsed model in TF layout ([in, out], fp32): "w1","b1","w2","b2","w3","b3","lstm_kernel","lstm_bias",
// "w5","b5","w6","b6".  Returns the element count (also when aOut is NULL or aCapacity is too small), negative on error.
long long STTX_InspectModelTensor(const char* aModelBuffer, unsigned int aBufferSize, const char* aName, float* aOut,
                                  unsigned long long aCapacity) {
  sttmodel::HostModel hm;
  if (sttmodel::load_from_buffer(reinterpret_cast<const uint8_t*>(aModelBuffer), aBufferSize, &hm)) return -1;
  const std::pair<const char*, const std::vector<float>*> all[] = {
      {"w1", &hm.w1}, {"b1", &hm.b1}, {"w2", &hm.w2}, {"b2", &hm.b2}, {"w3", &hm.w3}, {"b3", &hm.b3},
      {"lstm_kernel", &hm.lstm_kernel}, {"lstm_bias", &hm.lstm_bias}, {"w5", &hm.w5}, {"b5", &hm.b5}, {"w6", &hm.w6}, {"b6", &hm.b6}};
  for (const auto& kv : all)
    if (!strcmp(kv.first, aName)) {
      if (aOut && aCapacity >= kv.second->size()) memcpy(aOut, kv.second->data(), kv.second->size() * sizeof(float));
      return (long long)kv.second->size();
    }
  return -2;
}

long long STTX_StreamArenaCompactions(const StreamingState* aSctx) {
  return aSctx && aSctx->dev ? stteng::batch_stream_compactions(aSctx->dev) : -1;
}

int STTX_SpeechToTextBatch(ModelState* aCtx, const short* const* aBuffers, const unsigned int* aBufferSizes,
                           unsigned int aNumBuffers, char** aTranscriptsOut) {
  for (unsigned int i = 0; i < aNumBuffers; ++i) aTranscriptsOut[i] = nullptr;
  unsigned int max_samples = 1;
  for (unsigned int i = 0; i < aNumBuffers; ++i) max_samples = std::max(max_samples, aBufferSizes[i]);
  const unsigned int group = std::min(256u, std::max(1u, aNumBuffers));
  STTX_Batch* b = aCtx->oneshot;
  int err = STT_ERR_OK;
  if (!b || aCtx->oneshot_utts < group || aCtx->oneshot_samples < max_samples || b->beam_cap < std::max(1u, aCtx->beam_width_)) {
    if (b) STTX_BatchFree(b);
    aCtx->oneshot = nullptr;
    err = STTX_BatchCreate(aCtx, group, max_samples, &b);
    if (err) return err;
    aCtx->oneshot = b;
    aCtx->oneshot_utts = group;
    aCtx->oneshot_samples = max_samples;
  }
  for (unsigned int base = 0; base < aNumBuffers && !err; base += group) {
    const unsigned int n = std::min(group, aNumBuffers - base);
    err = STTX_BatchUpload(b, aBuffers + base, aBufferSizes + base, n);
    if (!err) err = STTX_BatchForward(b);
    if (!err) err = STTX_BatchDecode(b, 1);
    if (!err) err = STTX_BatchFetch(b);
    for (unsigned int i = 0; i < n && !err; ++i) aTranscriptsOut[base + i] = STTX_BatchTranscript(b, i, 0);
  }
  return err;
}

}  // extern "C"
