// Host harness: exposes the HD scorer-view functions (the same code the CUDA kernel compiles) through a
// C ABI so tests/test_scorer_lm.py can compare them with the compiled reference.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <vector>
#include "../../stt_b200/csrc/scorer_image.h"

using namespace sttscorer;
struct Handle { std::vector<uint8_t> bytes; ScorerView view; };

extern "C" {
void* sc_load(const char* path, const char* labels, int n_labels, int space, int* err) {
  std::ifstream f(path, std::ios::binary);
  if (!f) { *err = SCORER_UNREADABLE; return nullptr; }
  std::stringstream ss; ss << f.rdbuf();
  std::string s = ss.str();
  auto* h = new Handle();
  h->bytes.assign(s.begin(), s.end());
  h->bytes.resize(h->bytes.size() + 16, 0);
  AlphabetBytes ab; const char* p = labels;
  for (int i = 0; i < n_labels; ++i) { ab.labels.emplace_back(p); p += ab.labels.back().size() + 1; }
  ab.space_label = space;
  *err = parse_scorer(h->bytes.data(), s.size(), ab, &h->view);
  if (*err) { delete h; return nullptr; }
  h->view.blob = h->bytes.data();
  return h;
}
void sc_free(void* h) { delete static_cast<Handle*>(h); }
int sc_order(void* h) { return static_cast<Handle*>(h)->view.order; }
double sc_alpha(void* h) { return static_cast<Handle*>(h)->view.alpha; }
double sc_beta(void* h) { return static_cast<Handle*>(h)->view.beta; }
unsigned sc_vocab_index(void* h, const char* w) {
  return vocab_index(static_cast<Handle*>(h)->view, (const uint8_t*)w, (uint32_t)strlen(w));
}
// words: n NUL-terminated strings.
double sc_log_cond_prob(void* h, const char* words, int n, int bos) {
  const ScorerView& v = static_cast<Handle*>(h)->view;
  uint32_t ids[64]; const char* p = words;
  for (int i = 0; i < n; ++i) { size_t l = strlen(p); ids[i] = vocab_index(v, (const uint8_t*)p, (uint32_t)l); p += l + 1; }
  return log_cond_prob_ids(v, ids, n, bos != 0);
}
long sc_fst_start(void* h) { return static_cast<Handle*>(h)->view.fst_start; }
long sc_fst_nstates(void* h) { return static_cast<Handle*>(h)->view.fst_nstates; }
int sc_fst_find(void* h, int state, int label) { return fst_find(static_cast<Handle*>(h)->view, state, label); }
int sc_fst_final(void* h, int state) { return fst_is_final(static_cast<Handle*>(h)->view, state) ? 1 : 0; }
}
