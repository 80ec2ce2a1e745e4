// `stt` command-line client over the public C ABI only (include/stt_capi.h) -- SURVEY 8(f) rank 3.
// Same flags and output formats as the reference client (native_client/client.cc:42-191 JSON / word timing helpers,
// :193-287 inference modes, :482-635 main; native_client/args.h:77-204 flags); audio is read with the reference's
// NO_SOX rule: canonical RIFF/WAVE, PCM16, mono, at the model's sample rate (client.cc:390-426).
#include <getopt.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/stt_capi.h"

namespace {

const char* model = nullptr;
const char* scorer = nullptr;
const char* audio = nullptr;
bool set_beamwidth = false, set_alphabeta = false, show_times = false, extended_metadata = false, json_output = false;
bool keep_emissions = false, init_from_bytes = false, has_versions = false;
int beam_width = 0, json_candidate_transcripts = 3, stream_size = 0, extended_stream_size = 0;
float lm_alpha = 0.f, lm_beta = 0.f;
const char* hot_words = nullptr;

void print_help(const char* bin) {
  std::cout << "Usage: " << bin
            << " --model MODEL [--scorer SCORER] --audio AUDIO [-t] [-e]\n\nRunning Coqui STT inference (B200 backend).\n\n"
               "\t--model MODEL\t\t\tPath to the model (.sttw)\n\t--scorer SCORER\t\t\tPath to the external scorer file\n"
               "\t--audio AUDIO\t\t\tPath to the audio file to run (WAV PCM16 mono)\n\t--beam_width BEAM_WIDTH\t\tBeam width\n"
               "\t--lm_alpha LM_ALPHA\t\tLanguage model weight\n\t--lm_beta LM_BETA\t\tWord insertion bonus\n"
               "\t-t\t\t\t\tOutput inference time\n\t--extended\t\t\tOutput string from extended metadata\n"
               "\t--keep_emissions\t\tSave the output of the acoustic model\n\t--json\t\t\t\tWord timings as JSON\n"
               "\t--candidate_transcripts NUMBER\tCandidate transcripts in JSON output\n"
               "\t--stream size\t\t\tStream mode, output intermediate results\n"
               "\t--extended_stream size\t\tStream mode using metadata output\n"
               "\t--hot_words\t\t\tWord:Boost pairs, comma-separated\n\t--init_from_bytes\t\tInit model and scorer from bytes\n"
               "\t--help\t\t\t\tShow help\n\t--version\t\t\tPrint version and exit\n";
  char* v = STT_Version();
  std::cerr << "Coqui STT " << v << "\n";
  STT_FreeString(v);
  exit(1);
}

bool process_args(int argc, char** argv) {
  const char* const short_opts = "m:l:a:b:c:d:tejs:w:vh";
  const option long_opts[] = {{"model", required_argument, nullptr, 'm'}, {"scorer", required_argument, nullptr, 'l'},
                              {"audio", required_argument, nullptr, 'a'}, {"beam_width", required_argument, nullptr, 'b'},
                              {"lm_alpha", required_argument, nullptr, 'c'}, {"lm_beta", required_argument, nullptr, 'd'},
                              {"t", no_argument, nullptr, 't'}, {"extended", no_argument, nullptr, 'e'},
                              {"keep_emissions", no_argument, nullptr, 'L'}, {"json", no_argument, nullptr, 'j'},
                              {"init_from_bytes", no_argument, nullptr, 'B'},
                              {"candidate_transcripts", required_argument, nullptr, 150},
                              {"stream", required_argument, nullptr, 's'}, {"extended_stream", required_argument, nullptr, 'S'},
                              {"hot_words", required_argument, nullptr, 'w'}, {"version", no_argument, nullptr, 'v'},
                              {"help", no_argument, nullptr, 'h'}, {nullptr, no_argument, nullptr, 0}};
  for (;;) {
    const int opt = getopt_long(argc, argv, short_opts, long_opts, nullptr);
    if (opt == -1) break;
    switch (opt) {
      case 'm': model = optarg; break;
      case 'l': scorer = optarg; break;
      case 'a': audio = optarg; break;
      case 'b': set_beamwidth = true; beam_width = atoi(optarg); break;
      case 'c': set_alphabeta = true; lm_alpha = (float)atof(optarg); break;
      case 'd': set_alphabeta = true; lm_beta = (float)atof(optarg); break;
      case 't': show_times = true; break;
      case 'e': extended_metadata = true; break;
      case 'L': keep_emissions = true; break;
      case 'j': json_output = true; break;
      case 'B': init_from_bytes = true; break;
      case 150: json_candidate_transcripts = atoi(optarg); break;
      case 's': stream_size = atoi(optarg); break;
      case 'S': extended_stream_size = atoi(optarg); break;
      case 'v': has_versions = true; break;
      case 'w': hot_words = optarg; break;
      default: print_help(argv[0]);
    }
  }
  if (has_versions) {
    char* v = STT_Version();
    std::cout << "Coqui " << v << "\n";
    STT_FreeString(v);
    return false;
  }
  if (!model || !audio) {
    print_help(argv[0]);
    return false;
  }
  if (stream_size < 0 || stream_size % 160 != 0 || extended_stream_size < 0 || extended_stream_size % 160 != 0) {
    std::cout << "Stream buffer size must be multiples of 160\n";
    return false;
  }
  return true;
}

struct MetaWord {
  std::string word;
  float start_time, duration;
};

std::string transcript_to_string(const CandidateTranscript* t) {
  std::string s;
  for (unsigned i = 0; i < t->num_tokens; ++i) s += t->tokens[i].text;
  return s;
}

std::vector<MetaWord> transcript_to_words(const CandidateTranscript* t) {  // client.cc:64-103
  std::vector<MetaWord> words;
  std::string word;
  float start = 0;
  for (unsigned i = 0; i < t->num_tokens; ++i) {
    const TokenMetadata& tok = t->tokens[i];
    const bool is_space = strcmp(tok.text, " ") == 0;
    if (!is_space) {
      if (word.empty()) start = tok.start_time;
      word.append(tok.text);
    }
    if (is_space || i == t->num_tokens - 1) {
      float dur = tok.start_time - start;
      if (dur < 0) dur = 0;
      words.push_back({word, start, dur});
      word.clear();
      start = 0;
    }
  }
  return words;
}

std::string transcript_to_json(const CandidateTranscript* t) {  // client.cc:105-123
  std::ostringstream out;
  std::vector<MetaWord> words = transcript_to_words(t);
  out << R"("metadata":{"confidence":)" << t->confidence << R"(},"words":[)";
  for (size_t i = 0; i < words.size(); ++i) {
    out << R"({"word":")" << words[i].word << R"(","time":)" << words[i].start_time << R"(,"duration":)"
        << words[i].duration << "}";
    if (i + 1 < words.size()) out << ",";
  }
  out << "]";
  return out.str();
}

std::string metadata_to_json(const Metadata* m) {  // client.cc:125-191
  std::ostringstream out;
  out << "{\n";
  for (unsigned j = 0; j < m->num_transcripts; ++j) {
    const CandidateTranscript* t = &m->transcripts[j];
    if (j == 0) {
      out << transcript_to_json(t);
      if (m->num_transcripts > 1) out << ",\n" << R"("alternatives")" << ":[\n";
    } else {
      out << "{" << transcript_to_json(t) << "}";
      out << (j + 1 < m->num_transcripts ? ",\n" : "\n]");
    }
  }
  if (keep_emissions && m->emissions) {
    const int nt = m->emissions->num_timesteps, ns = m->emissions->num_symbols, cd = ns + 1;
    out << ",\n" << R"("alphabet")" << ":[";
    for (int i = 0; i < cd; ++i) out << "\"" << m->emissions->symbols[i] << "\"" << (i + 1 < cd ? ", " : "");
    out << "],\n" << R"("emissions")" << ":[\n";
    for (int i = 0; i < nt; ++i) {
      out << "[";
      for (int j = 0; j < ns; ++j) out << m->emissions->emissions[i * ns + j] << (j + 1 < ns ? ", " : "");
      out << "]" << (i + 1 < nt ? "," : "") << "\n";
    }
    out << "\n]";
  }
  out << "\n}\n";
  return out.str();
}

bool read_wav(const char* path, int want_rate, std::vector<short>* pcm) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  f.seekg(0, std::ios::end);
  const unsigned long long file_size = (unsigned long long)f.tellg();
  f.seekg(0, std::ios::beg);
  char riff[12];
  f.read(riff, 12);
  if (!f || memcmp(riff, "RIFF", 4) || memcmp(riff + 8, "WAVE", 4)) return false;
  unsigned short fmt = 0, channels = 0, bits = 0;
  unsigned int rate = 0;
  for (;;) {
    char id[4];
    unsigned int size;
    f.read(id, 4);
    f.read(reinterpret_cast<char*>(&size), 4);
    if (!f) return false;
    const unsigned long long here = (unsigned long long)f.tellg();
    const unsigned long long left = file_size > here ? file_size - here : 0;
    if (!memcmp(id, "fmt ", 4)) {
      if (size < 16 || size > left) return false;   // WAVEFORMAT: 16 bytes at least
      std::vector<char> b(size);
      f.read(b.data(), size);
      if (!f) return false;
      memcpy(&fmt, &b[0], 2);
      memcpy(&channels, &b[2], 2);
      memcpy(&rate, &b[4], 4);
      memcpy(&bits, &b[14], 2);
      if (size & 1) f.seekg(1, std::ios::cur);
    } else if (!memcmp(id, "data", 4)) {
      if (fmt != 1 || channels != 1 || bits != 16 || (int)rate != want_rate) {
        fprintf(stderr, "Error: audio must be WAV PCM16 mono at %d Hz (got format %u, %u ch, %u bit, %u Hz)\n", want_rate, fmt,
                channels, bits, rate);
        return false;
      }
      // a streamed file may announce more than it holds (0xffffffff from a pipe): take what is there, whole samples only
      const unsigned long long bytes = (size < left ? size : left) & ~1ull;
      pcm->resize(bytes / 2);
      f.read(reinterpret_cast<char*>(pcm->data()), (std::streamsize)bytes);
      pcm->resize((size_t)f.gcount() / 2);
      return true;
    } else {
      if (size > left) return false;
      f.seekg((std::streamoff)size + (size & 1), std::ios::cur);
    }
  }
}

std::string read_file(const char* path) {
  std::ifstream f(path, std::ios::binary);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

std::string run_inference(ModelState* ctx, const short* buf, size_t n) {  // client.cc:193-287
  std::string out;
  if (extended_metadata && !keep_emissions) {
    Metadata* r = STT_SpeechToTextWithMetadata(ctx, buf, n, 1);
    out = transcript_to_string(&r->transcripts[0]);
    STT_FreeMetadata(r);
  } else if (json_output && !keep_emissions) {
    Metadata* r = STT_SpeechToTextWithMetadata(ctx, buf, n, json_candidate_transcripts);
    out = metadata_to_json(r);
    STT_FreeMetadata(r);
  } else if (keep_emissions) {
    Metadata* r = STT_SpeechToTextWithEmissions(ctx, buf, n, json_candidate_transcripts);
    out = metadata_to_json(r);
    STT_FreeMetadata(r);
  } else if (stream_size > 0 || extended_stream_size > 0) {
    const size_t chunk = stream_size > 0 ? stream_size : extended_stream_size;
    StreamingState* s;
    if (STT_CreateStream(ctx, &s) != STT_ERR_OK) return "";
    std::string last;
    bool have_last = false;
    for (size_t off = 0; off < n;) {
      const size_t cur = n - off > chunk ? chunk : n - off;
      STT_FeedAudioContent(s, buf + off, cur);
      off += cur;
      std::string partial;
      if (stream_size > 0) {
        char* p = STT_IntermediateDecode(s);
        partial = p ? p : "";
        STT_FreeString(p);
      } else {
        Metadata* r = STT_IntermediateDecodeWithMetadata(s, 1);
        partial = transcript_to_string(&r->transcripts[0]);
        STT_FreeMetadata(r);
      }
      if (!have_last || partial != last) {
        printf("%s\n", partial.c_str());
        last = partial;
        have_last = true;
      }
    }
    if (stream_size > 0) {
      char* p = STT_FinishStream(s);
      out = p ? p : "";
      STT_FreeString(p);
    } else {
      Metadata* r = STT_FinishStreamWithMetadata(s, 1);
      out = transcript_to_string(&r->transcripts[0]);
      STT_FreeMetadata(r);
    }
  } else {
    char* p = STT_SpeechToText(ctx, buf, n);
    out = p ? p : "";
    STT_FreeString(p);
  }
  return out;
}

}  // namespace

int main(int argc, char** argv) {
  if (!process_args(argc, argv)) return 1;
  ModelState* ctx;
  int status;
  std::string model_bytes;
  if (init_from_bytes) {
    model_bytes = read_file(model);
    status = STT_CreateModelFromBuffer(model_bytes.data(), model_bytes.size(), &ctx);
  } else {
    status = STT_CreateModel(model, &ctx);
  }
  if (status != 0) {
    char* e = STT_ErrorCodeToErrorMessage(status);
    fprintf(stderr, "Could not create model: %s\n", e);
    free(e);
    return 1;
  }
  if (set_beamwidth && STT_SetModelBeamWidth(ctx, beam_width) != 0) {
    fprintf(stderr, "Could not set model beam width.\n");
    return 1;
  }
  if (scorer) {
    if (init_from_bytes) {
      const std::string sb = read_file(scorer);
      status = STT_EnableExternalScorerFromBuffer(ctx, sb.data(), sb.size());
    } else {
      status = STT_EnableExternalScorer(ctx, scorer);
    }
    if (status != 0) {
      fprintf(stderr, "Could not enable external scorer.\n");
      return 1;
    }
    if (set_alphabeta && STT_SetScorerAlphaBeta(ctx, lm_alpha, lm_beta) != 0) {
      fprintf(stderr, "Error setting scorer alpha and beta.\n");
      return 1;
    }
  }
  if (hot_words) {  // "word:boost,word:boost" (client.cc:560-580)
    std::stringstream ss(hot_words);
    std::string pair;
    while (std::getline(ss, pair, ',')) {
      const size_t c = pair.find(':');
      if (c == std::string::npos) continue;
      if (STT_AddHotWord(ctx, pair.substr(0, c).c_str(), (float)atof(pair.substr(c + 1).c_str())) != 0) {
        fprintf(stderr, "Could not enable hot-word.\n");
        return 1;
      }
    }
  }
  std::vector<short> pcm;
  if (!read_wav(audio, STT_GetModelSampleRate(ctx), &pcm)) {
    fprintf(stderr, "Error: could not read %s\n", audio);
    return 1;
  }
  const clock_t t0 = clock();
  const std::string result = run_inference(ctx, pcm.data(), pcm.size());
  const double cpu = (double)(clock() - t0) / CLOCKS_PER_SEC;
  printf("%s\n", result.c_str());
  if (show_times) printf("cpu_time_overall=%.05f\n", cpu);
  STT_FreeModel(ctx);
  return 0;
}
