// Host-side parser: `.scorer` bytes -> sttscorer::ScorerView offsets.
//
// Restates the load path of the reference (all under native_client/):
//   Scorer::load_lm_filepath / load_trie_impl          ctcdecode/scorer.cpp:108-146,177-222
//   lm::ngram::RecognizeBinary, Sanity header, FixedWidthParameters, TotalHeaderSize
//                                                       kenlm/lm/binary_format.cc:23-78,150-240, binary_format.hh:29-37
//   GenericModel::SetupMemory / LoadedBinary           kenlm/lm/model.cc:26-35,64-98
//   SortedVocabulary::{Size,LoadedBinary}              kenlm/lm/vocab.cc:113-116,218-232
//   TrieSearch::SetupMemory                            kenlm/lm/search_trie.cc:546-573
//   SeparatelyQuantize::{Size,SetupMemory,UpdateConfigFromBinary}   kenlm/lm/quantize.hh:142-147, quantize.cc:42-81
//   ArrayBhiksha::{Size,InlineBits,ctor,UpdateConfigFromBinary}, ChopBits, ArrayCount   kenlm/lm/bhiksha.cc:20-88
//   BitPacked::{BaseSize,BaseInit}, BitPackedMiddle ctor/Size        kenlm/lm/trie.cc:39-72
//   FstHeader::Read, ConstFstImpl::Read, AlignInput     third_party/openfst-1.6.7/src/lib/fst.cc:58-82,
//                                                       include/fst/const-fst.h:195-235, lib/util.cc:60-72
#include "scorer_image.h"

#include <algorithm>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

namespace sttscorer {
namespace {

const char kMagicBytes[] = "mmap lm http://kheafield.com/code format version 5\n\0";

uint8_t required_bits(uint64_t max_value) {
  if (!max_value) return 0;
  uint8_t ret = 1;
  while (max_value >>= 1) ++ret;
  return ret;
}

uint8_t chop_bits(uint64_t max_offset, uint64_t max_next, uint8_t pointer_bhiksha_bits) {
  uint8_t required = required_bits(max_next);
  uint8_t best_chop = 0;
  int64_t lowest_change = std::numeric_limits<int64_t>::max();
  for (uint8_t chop = 0; chop <= std::min(required, pointer_bhiksha_bits); ++chop) {
    int64_t change = (int64_t)((max_next >> (required - chop)) * 64) - (int64_t)max_offset * (int64_t)chop;
    if (change < lowest_change) {
      lowest_change = change;
      best_chop = chop;
    }
  }
  return best_chop;
}

template <class T>
bool rd(const uint8_t* file, size_t size, uint64_t& pos, T* out) {
  if (pos + sizeof(T) > size) return false;
  memcpy(out, file + pos, sizeof(T));
  pos += sizeof(T);
  return true;
}

bool rd_string(const uint8_t* file, size_t size, uint64_t& pos, std::string* out) {
  int32_t n;
  if (!rd(file, size, pos, &n) || n < 0 || pos + (uint64_t)n > size) return false;
  out->assign(reinterpret_cast<const char*>(file + pos), n);
  pos += n;
  return true;
}

// Every child range a lookup can obtain must lie inside the next level: the search functions of scorer_view.h (host
// and device) index records with whatever the parent's `next` pointers say, so a damaged record would send them
// outside the file -- on the GPU, an illegal address that takes the CUDA context down.  One sequential pass per level:
// the pointers (inline bits + ArrayBhiksha high part, bhiksha.hh:76-97) must be non-decreasing and end inside the
// next level; KenLM itself trusts the file here (trie.cc:74-99).
bool trie_pointers_ok(const ScorerView& v, const std::vector<uint64_t>& counts) {
  const int order = (int)v.order;
  {  // unigrams: `next` of words 0 .. counts[0] (the last one is the end pointer, search_trie.cc:517-518)
    const uint8_t* u = v.blob + v.unigram_off;
    uint64_t prev = 0;
    for (uint64_t w = 0; w <= counts[0]; ++w) {
      const uint64_t nx = load_u64(u + w * 16 + 8);
      if (nx < prev) return false;
      prev = nx;
    }
    if (prev > counts[1]) return false;
  }
  for (int i = 2; i < order; ++i) {
    const MiddleView& m = v.middle[i - 2];
    const uint8_t* base = v.blob + m.records_off;
    const uint8_t* offs = v.blob + m.offsets_off;
    if (v.bhiksha) {
      uint64_t prev = 0;
      for (uint64_t k = 0; k < m.offsets_count; ++k) {
        const uint64_t x = load_u64(offs + k * 8);
        if (x < prev || (k == 0 && x != 0)) return false;
        prev = x;
      }
      if (m.offsets_count == 0 || m.next_bits >= 57) return false;
    }
    uint64_t prev = 0, ub = 0;
    const uint64_t next_at = (uint64_t)m.word_bits + m.quant_bits;
    for (uint64_t r = 0; r <= m.n_records; ++r) {
      uint64_t nx = read_int57(base, r * m.total_bits + next_at, m.next_mask);
      if (v.bhiksha) {
        while (ub < m.offsets_count && load_u64(offs + ub * 8) <= r) ++ub;   // ub = #offsets <= r  (>= 1: offsets[0] == 0)
        nx |= (ub - 1) << m.next_bits;
      }
      if (nx < prev) return false;
      prev = nx;
    }
    if (prev > counts[i]) return false;
  }
  return true;
}

// ConstFst body: every state's arc span inside the arc array, every arc's target a state (const-fst.h:102-110)
bool fst_body_ok(const uint8_t* file, const ScorerView& v) {
  for (int64_t q = 0; q < v.fst_nstates; ++q) {
    const uint8_t* srec = file + v.fst_states_off + (uint64_t)q * 20;
    const uint64_t pos = load_u32(srec + 4), narcs = load_u32(srec + 8);
    if (pos + narcs > (uint64_t)v.fst_narcs) return false;
  }
  for (int64_t a = 0; a < v.fst_narcs; ++a) {
    const int32_t nx = (int32_t)load_u32(file + v.fst_arcs_off + (uint64_t)a * 16 + 12);
    if (nx < 0 || nx >= v.fst_nstates) return false;
  }
  return true;
}

}  // namespace

int parse_scorer(const uint8_t* file, size_t size, const AlphabetBytes& alphabet, ScorerView* v) {
  memset(v, 0, sizeof(*v));
  v->blob = nullptr;
  v->blob_size = size;

  // ---- KenLM Sanity header (binary_format.cc:48-62): 56-byte magic, f32 0,1,-0.5, u32 1,max,0, u64 1
  const size_t kSanity = 88;
  if (size <= kSanity) return SCORER_INVALID_LM;
  {
    uint8_t ref[kSanity];
    memset(ref, 0, sizeof(ref));
    memcpy(ref, kMagicBytes, sizeof(kMagicBytes));
    float f[3] = {0.0f, 1.0f, -0.5f};
    memcpy(ref + 56, f, 12);
    uint32_t w[3] = {1u, 0xffffffffu, 0u};
    memcpy(ref + 68, w, 12);
    uint64_t one = 1;
    memcpy(ref + 80, &one, 8);
    if (memcmp(ref, file, kSanity) != 0) return SCORER_INVALID_LM;
  }
  // FixedWidthParameters {u8 order; f32 probing_multiplier; i32 model_type; bool has_vocabulary; u32 search_version}
  // natural alignment: order@0, pm@4, model_type@8, has_vocab@12, search_version@16, sizeof = 20
  uint64_t pos = kSanity;
  if (pos + 20 > size) return SCORER_INVALID_LM;
  uint8_t order = file[pos];
  int32_t model_type;
  memcpy(&model_type, file + pos + 8, 4);
  uint32_t search_version;
  memcpy(&search_version, file + pos + 16, 4);
  pos += 20;
  if (order < 2 || order > kMaxOrder) return SCORER_INVALID_LM;
  if (model_type < 0 || model_type > 5 || model_type == 1) return SCORER_INVALID_LM;   // REST_PROBING: no fixture to verify against
  const bool probing = model_type == 0;                              // PROBING (model_type.hh:8-20)
  if (search_version != (probing ? 0u : 1u)) return SCORER_INVALID_LM;   // HashedSearch::kVersion 0, TrieSearch::kVersion 1
  float probing_multiplier;
  memcpy(&probing_multiplier, file + kSanity + 4, 4);
  std::vector<uint64_t> counts(order);
  if (pos + 8ull * order > size) return SCORER_INVALID_LM;
  memcpy(counts.data(), file + pos, 8ull * order);
  // a record takes at least one bit: counts beyond 8 * size cannot belong to this file (and would overflow the size
  // arithmetic below); counts[0] includes <unk>, so it is at least 1
  if (size > (1ull << 48)) return SCORER_INVALID_LM;
  for (uint64_t c : counts)
    if (c > 8ull * size) return SCORER_INVALID_LM;
  if (counts[0] == 0 || counts[0] > 0xffffffffull) return SCORER_INVALID_LM;   // WordIndex is 32 bits
  const uint64_t header_size = ((kSanity + 20 + 8ull * order - 1) / 8 + 1) * 8;  // ALIGN8

  v->order = order;
  uint64_t trie_offset = 0;
  if (probing) {
    // ---- probing-hash model (search_hashed.cc:206-220, vocab.cc:270-283): [vocabulary header 8 B + table of 12-byte
    //      entries][unigram weights x (count + 1)][middle tables, 8 + weights bytes per entry][longest table, 12 bytes]
    //      every table holds max(entries + 1, multiplier * entries) buckets (probing_hash_table.hh:108-111)
    if (!(probing_multiplier >= 1.0f) || !(probing_multiplier <= 1024.0f)) return SCORER_INVALID_LM;
    auto buckets_for = [&](uint64_t entries) {
      const uint64_t scaled = (uint64_t)(probing_multiplier * (float)entries);   // entries <= 2^51, multiplier <= 2^10
      return std::max<uint64_t>(entries + 1, scaled);
    };
    // a table must lie inside the file before it is probed; linear probing (probing_hash_table.hh:150-165) ends at an
    // empty bucket (key 0), so every table must hold one or a lookup would never return
    auto table_ok = [&](uint64_t off, uint64_t buckets, uint64_t entry_bytes) {
      if (buckets > size / entry_bytes || off > size || buckets * entry_bytes > size - off) return false;
      for (uint64_t i = 0; i < buckets; ++i) {
        uint64_t key;
        memcpy(&key, file + off + i * entry_bytes, 8);
        if (key == 0) return true;
      }
      return false;
    };
    v->probing = model_type == 0 ? 1 : 2;
    v->weights_size = model_type == 0 ? 8 : 12;
    uint64_t p = header_size;
    v->pvocab_buckets = buckets_for(counts[0]);
    v->pvocab_off = p + 8;
    if (header_size + 8 > size || !table_ok(v->pvocab_off, v->pvocab_buckets, 12)) return SCORER_INVALID_LM;
    for (uint64_t i = 0; i < v->pvocab_buckets; ++i) {   // {u64 hash, u32 word id}: ids index the unigram array
      uint32_t id;
      memcpy(&id, file + v->pvocab_off + i * 12 + 8, 4);
      if (id >= counts[0]) return SCORER_INVALID_LM;
    }
    p += 8 + v->pvocab_buckets * 12;
    v->vocab_off = v->pvocab_off;      // start of the vocabulary section (hot-word lookups keep a host copy up to here)
    v->vocab_count = counts[0];        // word ids are < counts[0]
    v->unigram_off = p;
    p += (counts[0] + 1) * v->weights_size;
    if (p > size) return SCORER_INVALID_LM;
    for (int i = 2; i < order; ++i) {
      v->ptab_off[i - 2] = p;
      v->ptab_buckets[i - 2] = buckets_for(counts[i - 1]);
      if (!table_ok(p, v->ptab_buckets[i - 2], 8 + v->weights_size)) return SCORER_INVALID_LM;
      p += v->ptab_buckets[i - 2] * (8 + v->weights_size);
    }
    v->ptab_off[order - 2] = p;
    v->ptab_buckets[order - 2] = buckets_for(counts[order - 1]);
    if (!table_ok(p, v->ptab_buckets[order - 2], 12)) return SCORER_INVALID_LM;
    p += v->ptab_buckets[order - 2] * 12;
    trie_offset = p;
    if (size <= trie_offset) return SCORER_NO_TRIE;
    v->blob = file;
    v->bos_word = vocab_index(*v, reinterpret_cast<const uint8_t*>("<s>"), 3);
    v->eos_word = vocab_index(*v, reinterpret_cast<const uint8_t*>("</s>"), 4);
    v->bos_backoff = load_f32(file + v->unigram_off + (uint64_t)v->bos_word * v->weights_size + 4);
  } else {
  v->quantized = (model_type - 2) & 1;
  v->bhiksha = ((model_type - 2) >> 1) & 1;

  // ---- vocabulary: u64 count, then sorted hashes; region sized for counts[0] entries (incl. <unk> slot)
  const uint64_t vocab_region = 8 + 8 * counts[0];
  if (header_size + vocab_region > size) return SCORER_INVALID_LM;
  uint64_t n_hash;
  memcpy(&n_hash, file + header_size, 8);
  if (n_hash > counts[0]) return SCORER_INVALID_LM;
  v->vocab_off = header_size + 8;
  v->vocab_count = n_hash;

  // ---- search section
  uint64_t p = header_size + vocab_region;
  uint8_t prob_bits = 0, backoff_bits = 0;
  if (v->quantized) {
    if (p + 3 > size) return SCORER_INVALID_LM;
    if (file[p] != 2) return SCORER_INVALID_LM;  // kSeparatelyQuantizeVersion
    prob_bits = file[p + 1];
    backoff_bits = file[p + 2];
    if (prob_bits == 0 || backoff_bits == 0 || prob_bits > 25 || backoff_bits > 25) return SCORER_INVALID_LM;
    uint64_t t = p + 8;
    for (int i = 0; i < order - 2; ++i) {
      v->quant_tables_off[i][0] = t;
      t += 4ull << prob_bits;
      v->quant_tables_off[i][1] = t;
      t += 4ull << backoff_bits;
    }
    v->quant_tables_off[order - 2][0] = t;
    const uint64_t longest_table = 4ull << prob_bits;
    const uint64_t middle_table = (4ull << backoff_bits) + longest_table;
    p += (order - 2) * middle_table + longest_table + 8;
  }
  v->prob_bits = prob_bits;
  v->backoff_bits = backoff_bits;
  v->unigram_off = p;
  p += (counts[0] + 2) * 16;

  uint8_t pointer_bhiksha_bits = 0;
  if (v->bhiksha && order > 2) {
    if (p + 2 > size) return SCORER_INVALID_LM;
    if (file[p] != 0) return SCORER_INVALID_LM;  // kArrayBhikshaVersion
    pointer_bhiksha_bits = file[p + 1];
  }
  const uint8_t word_bits = required_bits(counts[0]);
  const uint64_t word_mask = (1ull << word_bits) - 1;
  const uint8_t middle_quant_bits = v->quantized ? prob_bits + backoff_bits : 63;
  const uint8_t longest_quant_bits = v->quantized ? prob_bits : 31;
  for (int i = 2; i < order; ++i) {  // middle for n-grams of order i
    MiddleView& m = v->middle[i - 2];
    const uint64_t entries = counts[i - 1], max_next = counts[i];
    const uint64_t max_offset = entries + 1;
    uint64_t bhiksha_size = 0;
    uint8_t inline_bits;
    if (v->bhiksha) {
      const uint8_t required = required_bits(max_next);
      const uint8_t chop = chop_bits(max_offset, max_next, pointer_bhiksha_bits);
      const uint64_t array_count = (max_next >> (required - chop)) + 1;
      inline_bits = required - chop;
      bhiksha_size = 8 * (1 + array_count) + 7;
      m.offsets_off = ((p + 7) & ~7ull) + 8;  // AlignTo8(base) + 8-byte header (mapping base is page aligned)
      m.offsets_count = array_count;
    } else {
      inline_bits = required_bits(max_next);
    }
    m.records_off = p + bhiksha_size;
    m.n_records = entries;
    m.word_bits = word_bits;
    m.word_mask = word_mask;
    m.quant_bits = middle_quant_bits;
    m.next_bits = inline_bits;
    m.next_mask = (1ull << inline_bits) - 1;
    m.total_bits = word_bits + middle_quant_bits + inline_bits;
    p += bhiksha_size + ((1 + entries) * m.total_bits + 7) / 8 + 8;
  }
  v->longest.records_off = p;
  v->longest.word_bits = word_bits;
  v->longest.word_mask = word_mask;
  v->longest.total_bits = word_bits + longest_quant_bits;
  p += ((1 + counts[order - 1]) * v->longest.total_bits + 7) / 8 + 8;
  trie_offset = p;  // GetEndOfSearchOffset(), model.cc:265-267
  if (size <= trie_offset) return SCORER_NO_TRIE;

  // <s>, </s>, begin-sentence backoff (model.cc:78-84)
  v->blob = file;  // temporarily host-addressed so the view functions can be used for setup
  if (!trie_pointers_ok(*v, counts)) {
    v->blob = nullptr;
    return SCORER_INVALID_LM;
  }
  v->bos_word = vocab_index(*v, reinterpret_cast<const uint8_t*>("<s>"), 3);
  v->eos_word = vocab_index(*v, reinterpret_cast<const uint8_t*>("</s>"), 4);
  {
    NodeRange ignored;
    float prob, bo;
    unigram_find(*v, v->bos_word, ignored, prob, bo);
    v->bos_backoff = bo;
  }
  }  // trie model

  // ---- 'TRIE' header (scorer.cpp:182-211)
  pos = trie_offset;
  int32_t magic, version;
  if (!rd(file, size, pos, &magic) || magic != 0x54524945) return SCORER_INVALID_TRIE;
  if (!rd(file, size, pos, &version)) return SCORER_INVALID_TRIE;
  if (version != 6) return SCORER_VERSION_MISMATCH;
  uint8_t utf8;
  double alpha, beta;
  if (!rd(file, size, pos, &utf8) || !rd(file, size, pos, &alpha) || !rd(file, size, pos, &beta))
    return SCORER_INVALID_TRIE;
  v->is_utf8 = utf8 ? 1 : 0;
  v->alpha = (double)(float)alpha;  // reset_params(float, float), scorer.cpp:346-351
  v->beta = (double)(float)beta;

  // ---- OpenFst header + ConstFst body
  int32_t fst_magic;
  if (!rd(file, size, pos, &fst_magic) || fst_magic != 2125659606) return SCORER_INVALID_TRIE;
  std::string fsttype, arctype;
  int32_t fversion, flags;
  uint64_t properties;
  int64_t start, numstates, numarcs;
  if (!rd_string(file, size, pos, &fsttype) || !rd_string(file, size, pos, &arctype) ||
      !rd(file, size, pos, &fversion) || !rd(file, size, pos, &flags) || !rd(file, size, pos, &properties) ||
      !rd(file, size, pos, &start) || !rd(file, size, pos, &numstates) || !rd(file, size, pos, &numarcs))
    return SCORER_INVALID_TRIE;
  if (fsttype != "const" || arctype != "standard") return SCORER_INVALID_TRIE;
  if (flags & 3) return SCORER_INVALID_TRIE;  // embedded symbol tables: never written by save_dictionary
  const bool aligned = (flags & 4) || fversion == 1;
  if (aligned) pos = (pos + 15) & ~15ull;
  v->fst_states_off = pos;
  // bounds first: the products below must not wrap
  if (numstates < 0 || numarcs < 0 || (uint64_t)numstates > size / 20 || (uint64_t)numarcs > size / 16) return SCORER_INVALID_TRIE;
  if (start < -1 || start >= numstates) return SCORER_INVALID_TRIE;   // kNoStateId = -1: an empty automaton
  pos += (uint64_t)numstates * 20;
  if (aligned) pos = (pos + 15) & ~15ull;
  v->fst_arcs_off = pos;
  pos += (uint64_t)numarcs * 16;
  if (pos > size) return SCORER_INVALID_TRIE;
  v->fst_start = start;
  v->fst_nstates = numstates;
  v->fst_narcs = numarcs;
  if (!fst_body_ok(file, *v)) {
    v->blob = nullptr;
    return SCORER_INVALID_TRIE;
  }

  // ---- alphabet bytes
  v->n_labels = (uint32_t)alphabet.labels.size();
  v->space_label = alphabet.space_label;
  if (v->n_labels > 255) return SCORER_INVALID_TRIE;
  for (uint32_t i = 0; i < v->n_labels; ++i) {
    const std::string& s = alphabet.labels[i];
    if (s.size() > 4) return SCORER_INVALID_TRIE;
    v->label_len[i] = (uint8_t)s.size();
    memcpy(v->label_bytes[i], s.data(), s.size());
  }
  v->blob = nullptr;
  return SCORER_OK;
}

}  // namespace sttscorer
