// Host harness for tests/test_loader_fuzz.py: the two file parsers of the product library (scorer_image.cc: .scorer;
// model_file.cc + tflite_reader.cc: model files) on truncated and byte-flipped copies of a valid file.  Built with
// -fsanitize=address,undefined: a parser that reads outside the buffer, overflows or loops on a damaged file aborts
// the process.  Prints "accepted N rejected M".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "../../stt_b200/csrc/model_file.h"
#include "../../stt_b200/csrc/scorer_image.h"

static std::mt19937_64 qrng(99);
static bool g_query = false;   // the buffer carries the 16 spare bytes the engine gives the view (engine.cu: file + 16 B)

static int parse(bool scorer, const sttscorer::AlphabetBytes& ab, const uint8_t* p, size_t n) {
  if (scorer) {
    sttscorer::ScorerView v;
    const int e = sttscorer::parse_scorer(p, n, ab, &v);
    if (e == 0 && g_query) {
      // an accepted file is then USED: language-model walks and dictionary lookups (the functions the CUDA decoder
      // compiles) over random words / states must stay inside the buffer and return
      v.blob = p;
      for (int q = 0; q < 40; ++q) {
        uint32_t ids[sttscorer::kMaxOrder];
        const int k = 1 + (int)(qrng() % v.order);
        for (int i = 0; i < k; ++i) ids[i] = (uint32_t)(qrng() % (v.vocab_count + 1));
        volatile double r = sttscorer::log_cond_prob_ids(v, ids, k, (qrng() & 1) != 0);
        (void)r;
      }
      for (int q = 0; q < 40 && v.fst_nstates > 0; ++q) {
        const int32_t st = (int32_t)(qrng() % (uint64_t)v.fst_nstates);
        volatile int32_t r = sttscorer::fst_find(v, st, 1 + (int32_t)(qrng() % 255));
        volatile bool fin = sttscorer::fst_is_final(v, st);
        (void)r; (void)fin;
      }
    }
    return e;
  }
  sttmodel::HostModel m;
  return sttmodel::load_from_buffer(p, n, &m);
}

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const bool scorer = !strcmp(argv[1], "scorer");
  const bool utf8 = atoi(argv[3]) != 0;
  const int n_trunc = atoi(argv[4]), n_flip = atoi(argv[5]);
  std::ifstream f(argv[2], std::ios::binary);
  if (!f) return 2;
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string s = ss.str();
  sttscorer::AlphabetBytes ab;
  if (!utf8) {
    for (const char* p = " abcdefghijklmnopqrstuvwxyz'"; *p; ++p) ab.labels.emplace_back(1, *p);
    ab.space_label = 0;
  } else {
    for (int i = 1; i < 256; ++i) ab.labels.emplace_back(1, (char)i);
    ab.space_label = 31;
  }
  std::mt19937_64 rng(argc > 6 ? atoll(argv[6]) : 1);
  int ok = 0, bad = 0;
  if (parse(scorer, ab, reinterpret_cast<const uint8_t*>(s.data()), s.size()) != 0) return 3;   // the fixture itself must load
  for (int it = 0; it < n_trunc; ++it) {
    // every short length first (headers), then random ones; an exact-size heap block, so that any over-read is seen
    const size_t n = it < n_trunc / 2 ? (size_t)it % (s.size() + 1) : rng() % (s.size() + 1);
    std::vector<uint8_t> b(s.begin(), s.begin() + n);
    (parse(scorer, ab, b.data(), n) ? bad : ok)++;
  }
  std::vector<uint8_t> b(s.begin(), s.end());
  b.resize(s.size() + 16, 0);
  const size_t bsize = s.size();
  g_query = true;
  for (int it = 0; it < n_flip; ++it) {
    // flatbuffer roots / KenLM and OpenFst headers sit at the ends and at section boundaries: flip near both ends and anywhere
    size_t pos[3];
    uint8_t old[3];
    const int flips = 1 + (int)(rng() % 3);
    for (int k = 0; k < flips; ++k) {
      const int r = (int)(rng() % 3);
      pos[k] = r == 0 ? rng() % std::min<size_t>(2048, bsize)
               : r == 1 ? bsize - 1 - rng() % std::min<size_t>(4096, bsize)
                        : rng() % bsize;
      old[k] = b[pos[k]];
      b[pos[k]] = (uint8_t)rng();
    }
    (parse(scorer, ab, b.data(), bsize) ? bad : ok)++;
    for (int k = flips - 1; k >= 0; --k) b[pos[k]] = old[k];
  }
  printf("accepted %d rejected %d\n", ok, bad);
  return 0;
}
