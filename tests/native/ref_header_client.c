/* Compiled against the REFERENCE's coqui-stt.h (never ours) and linked against libstt_b200.so: proves that a client
 * written for libstt builds and links unchanged.  Runs without a GPU: only the calls that need no device are made. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "coqui-stt.h"

int main(void) {
  /* struct layouts as the reference declares them */
  printf("TokenMetadata %zu %zu %zu\n", sizeof(TokenMetadata), offsetof(TokenMetadata, timestep), offsetof(TokenMetadata, start_time));
  printf("CandidateTranscript %zu %zu %zu\n", sizeof(CandidateTranscript), offsetof(CandidateTranscript, num_tokens),
         offsetof(CandidateTranscript, confidence));
  printf("AcousticModelEmissions %zu %zu %zu %zu\n", sizeof(AcousticModelEmissions), offsetof(AcousticModelEmissions, symbols),
         offsetof(AcousticModelEmissions, num_timesteps), offsetof(AcousticModelEmissions, emissions));
  printf("Metadata %zu %zu %zu\n", sizeof(Metadata), offsetof(Metadata, num_transcripts), offsetof(Metadata, emissions));
  char* v = STT_Version();
  printf("version %s\n", v);
  STT_FreeString(v);
  char* e = STT_ErrorCodeToErrorMessage(STT_ERR_INVALID_SCORER);
  printf("err %s\n", e);
  STT_FreeString(e);
  ModelState* m = NULL;
  int rc = STT_CreateModel("", &m);           /* empty path -> STT_ERR_NO_MODEL, stt.cc:352-355 */
  printf("create_empty 0x%X %d\n", rc, m == NULL);
  /* every entry point must resolve at link time; taking the addresses keeps the linker honest */
  void* fns[] = {(void*)STT_CreateModel, (void*)STT_CreateModelFromBuffer, (void*)STT_GetModelBeamWidth, (void*)STT_SetModelBeamWidth,
                 (void*)STT_GetModelSampleRate, (void*)STT_FreeModel, (void*)STT_EnableExternalScorer,
                 (void*)STT_EnableExternalScorerFromBuffer, (void*)STT_AddHotWord, (void*)STT_EraseHotWord, (void*)STT_ClearHotWords,
                 (void*)STT_DisableExternalScorer, (void*)STT_SetScorerAlphaBeta, (void*)STT_SpeechToText,
                 (void*)STT_SpeechToTextWithMetadata, (void*)STT_SpeechToTextWithEmissions, (void*)STT_CreateStream,
                 (void*)STT_FeedAudioContent, (void*)STT_IntermediateDecode, (void*)STT_IntermediateDecodeWithMetadata,
                 (void*)STT_IntermediateDecodeFlushBuffers, (void*)STT_IntermediateDecodeWithMetadataFlushBuffers,
                 (void*)STT_FinishStream, (void*)STT_FinishStreamWithMetadata,
                 (void*)STT_FreeStream, (void*)STT_FreeMetadata, (void*)STT_FreeString, (void*)STT_Version,
                 (void*)STT_ErrorCodeToErrorMessage};
  size_t n = 0;
  for (size_t i = 0; i < sizeof(fns) / sizeof(fns[0]); ++i) n += fns[i] != NULL;
  printf("resolved %zu\n", n);
  return 0;
}
