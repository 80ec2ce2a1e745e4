/* C ABI of libstt_b200.so.
 *
 * PART 1 is binary-compatible with the reference's public header native_client/coqui-stt.h (29 exported
 * STT_* symbols, 4 POD structs, 23 error codes): a client compiled against coqui-stt.h links against this
 * library unchanged.  Each declaration cites the reference declaration (header line) and the definition it
 * replaces (native_client/stt.cc line).  Ownership rules are the reference's: strings are malloc'd and freed
 * with STT_FreeString, Metadata with STT_FreeMetadata, STT_FinishStream* and STT_SpeechToText* consume the stream.
 *
 * PART 2 (STTX_*) is additive: the reference C API is single-utterance (ModelState::BATCH_SIZE = 1,
 * native_client/modelstate.h:15-16); its only batch surface is the Python decoder
 * ctc_beam_search_decoder_batch (ctcdecode/ctc_beam_search_decoder.cpp:608-652).  STTX_* exposes the batched GPU
 * path, per-stage device timings, and test hooks.  No existing symbol changes meaning.
 */
#ifndef STT_B200_CAPI_H
#define STT_B200_CAPI_H

#ifdef __cplusplus
extern "C" {
#endif

#define STT_EXPORT __attribute__((visibility("default")))

typedef struct ModelState ModelState;         /* coqui-stt.h:22 */
typedef struct StreamingState StreamingState; /* coqui-stt.h:24 */

typedef struct TokenMetadata { /* coqui-stt.h:29-38 */
  const char* const text;
  const unsigned int timestep;
  const float start_time;
} TokenMetadata;

typedef struct CandidateTranscript { /* coqui-stt.h:44-55 */
  const TokenMetadata* const tokens;
  const unsigned int num_tokens;
  const double confidence;
} CandidateTranscript;

typedef struct AcousticModelEmissions { /* coqui-stt.h:65-74 */
  int num_symbols;
  const char** symbols;
  int num_timesteps;
  const double* emissions;
} AcousticModelEmissions;

typedef struct Metadata { /* coqui-stt.h:79-86 */
  const CandidateTranscript* const transcripts;
  const unsigned int num_transcripts;
  const AcousticModelEmissions* const emissions;
} Metadata;

/* coqui-stt.h:92-115, same numeric values and message texts */
#define STT_FOR_EACH_ERROR(APPLY)                                                                          \
  APPLY(STT_ERR_OK, 0x0000, "No error.")                                                                   \
  APPLY(STT_ERR_NO_MODEL, 0x1000, "Missing model information.")                                            \
  APPLY(STT_ERR_INVALID_ALPHABET, 0x2000, "Invalid alphabet embedded in model. (Data corruption?)")        \
  APPLY(STT_ERR_INVALID_SHAPE, 0x2001, "Invalid model shape.")                                             \
  APPLY(STT_ERR_INVALID_SCORER, 0x2002, "Invalid scorer file.")                                            \
  APPLY(STT_ERR_MODEL_INCOMPATIBLE, 0x2003, "Incompatible model.")                                         \
  APPLY(STT_ERR_SCORER_NOT_ENABLED, 0x2004, "External scorer is not enabled.")                             \
  APPLY(STT_ERR_SCORER_UNREADABLE, 0x2005, "Could not read scorer file.")                                  \
  APPLY(STT_ERR_SCORER_INVALID_LM, 0x2006, "Could not recognize language model header in scorer.")         \
  APPLY(STT_ERR_SCORER_NO_TRIE, 0x2007, "Reached end of scorer file before loading vocabulary trie.")      \
  APPLY(STT_ERR_SCORER_INVALID_TRIE, 0x2008, "Invalid magic in trie header.")                              \
  APPLY(STT_ERR_SCORER_VERSION_MISMATCH, 0x2009, "Scorer file version does not match expected version.")   \
  APPLY(STT_ERR_FAIL_INIT_MMAP, 0x3000, "Failed to initialize memory mapped model.")                       \
  APPLY(STT_ERR_FAIL_INIT_SESS, 0x3001, "Failed to initialize the session.")                               \
  APPLY(STT_ERR_FAIL_INTERPRETER, 0x3002, "Interpreter failed.")                                           \
  APPLY(STT_ERR_FAIL_RUN_SESS, 0x3003, "Failed to run the session.")                                       \
  APPLY(STT_ERR_FAIL_CREATE_STREAM, 0x3004, "Error creating the stream.")                                  \
  APPLY(STT_ERR_FAIL_READ_PROTOBUF, 0x3005, "Error reading the proto buffer model file.")                  \
  APPLY(STT_ERR_FAIL_CREATE_SESS, 0x3006, "Failed to create session.")                                     \
  APPLY(STT_ERR_FAIL_CREATE_MODEL, 0x3007, "Could not allocate model state.")                              \
  APPLY(STT_ERR_FAIL_INSERT_HOTWORD, 0x3008, "Could not insert hot-word.")                                 \
  APPLY(STT_ERR_FAIL_CLEAR_HOTWORD, 0x3009, "Could not clear hot-words.")                                  \
  APPLY(STT_ERR_FAIL_ERASE_HOTWORD, 0x3010, "Could not erase hot-word.")

enum STT_Error_Codes {
#define STT_DEFINE_ERR(NAME, VALUE, DESC) NAME = VALUE,
  STT_FOR_EACH_ERROR(STT_DEFINE_ERR)
#undef STT_DEFINE_ERR
};

/* ------------------------------------------------------------------------------------------------ PART 1 */
/* coqui-stt.h:137-138 | stt.cc:374-379 (+ CreateModelImpl :336-372).  Accepts the reference's `.tflite` flatbuffer
 * (native_client/tflitemodelstate.cc:161-338) or the native `.sttw` container. */
STT_EXPORT int STT_CreateModel(const char* aModelPath, ModelState** retval);
/* coqui-stt.h:150-152 | stt.cc:381-387.  The buffer (either format) is parsed in place; it need not outlive the call. */
STT_EXPORT int STT_CreateModelFromBuffer(const char* aModelBuffer, unsigned int aBufferSize, ModelState** retval);
/* coqui-stt.h:164 | stt.cc:389-393 */
STT_EXPORT unsigned int STT_GetModelBeamWidth(const ModelState* aCtx);
/* coqui-stt.h:176-177 | stt.cc:395-400 */
STT_EXPORT int STT_SetModelBeamWidth(ModelState* aCtx, unsigned int aBeamWidth);
/* coqui-stt.h:187 | stt.cc:402-406 */
STT_EXPORT int STT_GetModelSampleRate(const ModelState* aCtx);
/* coqui-stt.h:193 | stt.cc:408-412 */
STT_EXPORT void STT_FreeModel(ModelState* ctx);
/* coqui-stt.h:204-205 | stt.cc:435-440 (EnableExternalScorerImpl :414-433: any failure -> STT_ERR_INVALID_SCORER) */
STT_EXPORT int STT_EnableExternalScorer(ModelState* aCtx, const char* aScorerPath);
/* coqui-stt.h:217-219 | stt.cc:442-449 */
STT_EXPORT int STT_EnableExternalScorerFromBuffer(ModelState* aCtx, const char* aScorerBuffer, unsigned int aBufferSize);
/* coqui-stt.h:233-235 | stt.cc:451-466 */
STT_EXPORT int STT_AddHotWord(ModelState* aCtx, const char* word, float boost);
/* coqui-stt.h:246-247 | stt.cc:468-482 */
STT_EXPORT int STT_EraseHotWord(ModelState* aCtx, const char* word);
/* coqui-stt.h:257 | stt.cc:484-496 */
STT_EXPORT int STT_ClearHotWords(ModelState* aCtx);
/* coqui-stt.h:267 | stt.cc:498-506 */
STT_EXPORT int STT_DisableExternalScorer(ModelState* aCtx);
/* coqui-stt.h:279-281 | stt.cc:508-517 */
STT_EXPORT int STT_SetScorerAlphaBeta(ModelState* aCtx, float aAlpha, float aBeta);
/* coqui-stt.h:295-297 | stt.cc:655-662 */
STT_EXPORT char* STT_SpeechToText(ModelState* aCtx, const short* aBuffer, unsigned int aBufferSize);
/* coqui-stt.h:315-318 | stt.cc:664-672 */
STT_EXPORT Metadata* STT_SpeechToTextWithMetadata(ModelState* aCtx, const short* aBuffer, unsigned int aBufferSize,
                                                   unsigned int aNumResults);
/* coqui-stt.h:334-337 | stt.cc:674-688 */
STT_EXPORT Metadata* STT_SpeechToTextWithEmissions(ModelState* aCtx, const short* aBuffer, unsigned int aBufferSize,
                                                    unsigned int aNumResults);
/* coqui-stt.h:349-350 | stt.cc:519-551 */
STT_EXPORT int STT_CreateStream(ModelState* aCtx, StreamingState** retval);
/* coqui-stt.h:361-363 | stt.cc:588-594 */
STT_EXPORT void STT_FeedAudioContent(StreamingState* aSctx, const short* aBuffer, unsigned int aBufferSize);
/* coqui-stt.h:374 | stt.cc:596-600 */
STT_EXPORT char* STT_IntermediateDecode(const StreamingState* aSctx);
/* coqui-stt.h:389-390 | stt.cc:602-607 */
STT_EXPORT Metadata* STT_IntermediateDecodeWithMetadata(const StreamingState* aSctx, unsigned int aNumResults);
/* coqui-stt.h:408 | stt.cc:609-614 (does NOT free the stream) */
STT_EXPORT char* STT_IntermediateDecodeFlushBuffers(StreamingState* aSctx);
/* coqui-stt.h:428-429 | stt.cc:616-622 */
STT_EXPORT Metadata* STT_IntermediateDecodeWithMetadataFlushBuffers(StreamingState* aSctx, unsigned int aNumResults);
/* coqui-stt.h:443 | stt.cc:624-630 */
STT_EXPORT char* STT_FinishStream(StreamingState* aSctx);
/* coqui-stt.h:461-462 | stt.cc:632-639 */
STT_EXPORT Metadata* STT_FinishStreamWithMetadata(StreamingState* aSctx, unsigned int aNumResults);
/* coqui-stt.h:474 | stt.cc:690-694 */
STT_EXPORT void STT_FreeStream(StreamingState* aSctx);
/* coqui-stt.h:480 | stt.cc:696-726 */
STT_EXPORT void STT_FreeMetadata(Metadata* m);
/* coqui-stt.h:486 | stt.cc:727-731 */
STT_EXPORT void STT_FreeString(char* str);
/* coqui-stt.h:495 | stt.cc:733-737 */
STT_EXPORT char* STT_Version(void);
/* coqui-stt.h:504 | native_client/stt_errors.cc:5-19 */
STT_EXPORT char* STT_ErrorCodeToErrorMessage(int aErrorCode);

/* ------------------------------------------------------------------------------------------------ PART 2 */
typedef struct STTX_Batch STTX_Batch;

typedef struct STTX_Timings { /* milliseconds, CUDA events on the library's stream */
  float h2d, mfcc, dense123, lstm_in, lstm, dense56, decode, d2h, total;
} STTX_Timings;

/* One call, HOST buffers in, transcripts out (malloc'd, free each with STT_FreeString): the batched counterpart of
 * STT_SpeechToText for n utterances; utterances beyond 256 are processed in groups.  Returns an STT error code. */
STT_EXPORT int STTX_SpeechToTextBatch(ModelState* aCtx, const short* const* aBuffers, const unsigned int* aBufferSizes,
                                       unsigned int aNumBuffers, char** aTranscriptsOut);

/* Staged variant used by bench.py and the parity tests. */
STT_EXPORT int STTX_BatchCreate(ModelState* aCtx, unsigned int aMaxUtterances, unsigned int aMaxSamples, STTX_Batch** retval);
STT_EXPORT void STTX_BatchFree(STTX_Batch* b);
/* Pinned host staging row of utterance u (capacity aMaxSamples samples): PCM written here is uploaded by
 * STTX_BatchUpload without an intermediate copy when the same pointer is passed in aBuffers[u]. */
STT_EXPORT short* STTX_BatchHostBuffer(STTX_Batch* b, unsigned int u);
STT_EXPORT int STTX_BatchUpload(STTX_Batch* b, const short* const* aBuffers, const unsigned int* aBufferSizes, unsigned int n);
STT_EXPORT int STTX_BatchForward(STTX_Batch* b);                       /* MFCC + acoustic model, results stay in HBM */
STT_EXPORT int STTX_BatchDecode(STTX_Batch* b, unsigned int aNumResults); /* beam search (model beam width, scorer) */
/* Result r of utterance u: transcript (malloc'd), confidence, tokens/timesteps copied into caller arrays (cap entries). */
STT_EXPORT int STTX_BatchNumResults(STTX_Batch* b, unsigned int u);
STT_EXPORT char* STTX_BatchTranscript(STTX_Batch* b, unsigned int u, unsigned int r);
STT_EXPORT int STTX_BatchTokens(STTX_Batch* b, unsigned int u, unsigned int r, unsigned int* tokens, unsigned int* timesteps,
                                unsigned int cap, double* confidence);
STT_EXPORT int STTX_BatchFetch(STTX_Batch* b);                         /* device -> host copy of the decode results */
STT_EXPORT int STTX_BatchGetTimings(STTX_Batch* b, STTX_Timings* out);
STT_EXPORT long long STTX_BatchKernelLaunches(STTX_Batch* b);
/* statistics build of the decoder kernel (per-phase clocks, LM counters) for the following STTX_BatchDecode calls;
 * off by default: the production kernel carries no instrumentation */
STT_EXPORT int STTX_BatchSetInstrumented(STTX_Batch* b, int aOn);
/* instrumentation: words scored by the LM / LM calls in the last STTX_BatchDecode (decoder roofline's Q) */
/* instrumentation: SM cycles per decoder phase (gate, child discovery, LM, live update, children, select, commit, -),
 * summed over the batch's utterances, for the last STTX_BatchDecode */
STT_EXPORT int STTX_BatchPhaseCycles(STTX_Batch* b, unsigned long long* out8);
/* instrumentation: LSTM kernel cycles of the last forward (max over CTAs): grid-barrier wait, load+MMA span, epilogue */
STT_EXPORT int STTX_BatchLstmProfile(STTX_Batch* b, unsigned long long* out3);
STT_EXPORT int STTX_BatchLmStats(STTX_Batch* b, unsigned long long* words_scored, unsigned long long* lm_calls);
/* Sums over the batch of the decoder's 16 per-utterance counters (stt_b200/csrc/decoder.cuh Slot::scalars); the statistics
   build (STTX_BatchSetInstrumented) fills 7.. with LM calls / cache misses / per-phase cycle splits. */
STT_EXPORT int STTX_BatchDecoderScalars(STTX_Batch* b, unsigned long long* out16);
/* Vocabulary pruning of the batch's following decodes = DecoderState::init's cutoff_prob / cutoff_top_n
   (native_client/ctcdecode/ctc_beam_search_decoder.cpp:22-61, get_pruned_emissions :328-358).  The reference's C API fixes
   them at 1.0 / 40 (native_client/stt.cc:539-540), which is the default here; its decoder-only Python package
   (native_client/ctcdecode/__init__.py:122-178) passes the caller's values. */
STT_EXPORT int STTX_BatchSetCutoff(STTX_Batch* b, double cutoff_prob, unsigned int cutoff_top_n);
/* test hooks */
STT_EXPORT int STTX_BatchTimesteps(STTX_Batch* b, unsigned int u);
STT_EXPORT int STTX_BatchCopyFeatures(STTX_Batch* b, unsigned int u, float* out);  /* [T, n_input] */
STT_EXPORT int STTX_BatchCopyProbs(STTX_Batch* b, unsigned int u, float* out);     /* [T, n_classes] */
STT_EXPORT int STTX_BatchSetProbs(STTX_Batch* b, const float* probs, const int* T, unsigned int n, unsigned int T_stride);
/* decoder-only entry (the reference's Python ctc_beam_search_decoder_batch takes f64 probabilities) */
STT_EXPORT int STTX_BatchSetProbs64(STTX_Batch* b, const double* probs, const int* T, unsigned int n, unsigned int T_stride);
/* Unit-test / bring-up hooks.  NOT part of the product library: they exist only in stt_b200/libstt_b200_dev.so, the same
 * sources compiled with -DSTT_B200_DEV_HOOKS (tests/test_gpu_gemm.py, tests/native/probe_pair.py). */
#ifdef STT_B200_DEV_HOOKS
/* TMEM layout of a CTA-pair MMA (M = 128 or 256); out = float[2][128][128] */
STT_EXPORT int STTX_DebugPairLayout(int M, float* out);
/* one GEMM of the acoustic model's kernels on caller data; epilogue: 0 clipped-ReLU fp16, 1 bias f32, 2 softmax (N = 32 or
 * 256), + 16 = the one-CTA kernel instead of the CTA-pair kernel */
STT_EXPORT int STTX_DebugGemm(int M, int N, int K, const unsigned short* a_f16, const unsigned short* w_f16,
                              const float* bias, int epilogue, float relu_clip, void* out, float* ms);
#endif
/* how many times this stream's device context has garbage-collected its decoder arena (streams are unbounded in length;
 * PathTrie::remove, path_trie.cpp:192-209, is what bounds the reference's memory) */
STT_EXPORT long long STTX_StreamArenaCompactions(const StreamingState* aSctx);
/* Model-file inspection without a device: parses a TFLite flatbuffer (the reference's container,
 * native_client/tflitemodelstate.cc:161-338) or a .sttw file exactly as STT_CreateModel does.  aInfo[12] = sample_rate,
 * win_len, win_step, n_input, n_context, n_hidden, n_cell, n_classes, n_steps, beam_width, space_label, n_labels. */
STT_EXPORT int STTX_InspectModel(const char* aModelBuffer, unsigned int aBufferSize, unsigned int* aInfo, float* aReluClip);
/* fp32 copy of one tensor ("w1".."b6", "lstm_kernel", "lstm_bias") in TF layout; returns its element count */
STT_EXPORT long long STTX_InspectModelTensor(const char* aModelBuffer, unsigned int aBufferSize, const char* aName,
                                             float* aOut, unsigned long long aCapacity);
STT_EXPORT int STTX_ModelInfo(const ModelState* aCtx, unsigned int* n_classes, unsigned int* n_input, unsigned int* n_hidden,
                              unsigned int* n_steps, unsigned int* n_sms);

#undef STT_EXPORT
#ifdef __cplusplus
}
#endif
#endif /* STT_B200_CAPI_H */
