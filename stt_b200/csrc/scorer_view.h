// Pointer-free, device-resident view of a Coqui `.scorer` package and the query functions over it.
//
// A `.scorer` is [KenLM trie binary]['TRIE' header][OpenFst ConstFst]  (SURVEY.md appendix C;
// native_client/ctcdecode/scorer.cpp:177-222).  The reference queries it through pointer-rich C++
// objects over an mmap; here the file bytes are copied verbatim into ONE device buffer and every
// structure is addressed by byte offset, so the same functions run on host (tests) and device
// (the beam-search kernel).  Everything below is integer / table-lookup work, so results are
// bit-identical to the reference by construction; tests/test_scorer_lm.py checks that against
// the compiled reference on every n-gram of the smoke-test LM.
//
// Restated reference functions (file:line under native_client/):
//   util::ReadInt57 / ReadInt25 / ReadFloat32 / ReadNonPositiveFloat31   kenlm/util/bit_packing.hh:62-150
//   SortedVocabulary::Index                                             kenlm/lm/vocab.hh:72-83
//   MurmurHash64A                                                       kenlm/util/murmur_hash.cc:26-80
//   trie::Unigram::Find                                                 kenlm/lm/trie.hh:66-71
//   BitPackedMiddle::Find / BitPackedLongest::Find / FindBitPacked      kenlm/lm/trie.cc:32-36,88-99,119-124
//   ArrayBhiksha::ReadNext / DontBhiksha::ReadNext                      kenlm/lm/bhiksha.hh:44-48,76-97
//   SeparatelyQuantize::{Middle,Longest}Pointer, DontQuantize::*        kenlm/lm/quantize.hh:31-80,152-216
//   GenericModel::FullScore / ScoreExceptBackoff / ResumeScore          kenlm/lm/model.cc:170-176,285-338
//   HasExtension                                                        kenlm/lm/blank.hh:33-39
//   Scorer::get_log_cond_prob                                           ctcdecode/scorer.cpp:307-344
//   ConstFst state/arc records, SortedMatcher::Find                     third_party/openfst-1.6.7/src/include/fst/const-fst.h:102-110, matcher.h:347-388
#pragma once
#include <stdint.h>

#include "hd_math.h"

namespace sttscorer {

constexpr int kMaxOrder = 6;  // KENLM_MAX_ORDER in the reference build (build_archive.py:19)
constexpr uint32_t kNotFound = 0xffffffffu;

struct MiddleView {
  uint64_t records_off;   // byte offset of the bit-packed record array (BitPacked::base_)
  uint64_t offsets_off;   // ArrayBhiksha offset_begin_ (8-aligned, after its 8-byte header); 0 if DontBhiksha
  uint64_t offsets_count; // ArrayBhiksha offset_end_ - offset_begin_
  uint64_t n_records;     // n-grams of this order (record indices run 0..n_records)
  uint64_t word_mask, next_mask;
  uint8_t word_bits, total_bits, quant_bits, next_bits;
};

struct LongestView {
  uint64_t records_off;
  uint64_t word_mask;
  uint8_t word_bits, total_bits;
};

struct FstStateRec {  // ConstFst<StdArc>::ConstState, const-fst.h:102-110
  float final_weight;
  uint32_t pos, narcs, niepsilons, noepsilons;
};
struct FstArcRec {  // StdArc
  int32_t ilabel, olabel;
  float weight;
  int32_t nextstate;
};

struct LmState {  // lm::ngram::State, kenlm/lm/state.hh:45-47
  uint32_t words[kMaxOrder - 1];
  float backoff[kMaxOrder - 1];
  uint8_t length;
};

struct ScorerView {
  const uint8_t* blob;  // the whole .scorer file, verbatim (+16 bytes of zero padding)
  uint64_t blob_size;
  // --- KenLM
  uint32_t order;
  uint32_t quantized, bhiksha;  // model_type = TRIE + 1*quantized + 2*bhiksha  (model_type.hh:8)
  uint32_t prob_bits, backoff_bits;
  uint64_t vocab_off;   // first hash of SortedVocabulary (begin_)
  uint64_t vocab_count; // end_ - begin_   (word id = index + 1; 0 = <unk>)
  uint64_t unigram_off; // trie::UnigramValue[counts[0] + 2]
  uint64_t quant_tables_off[kMaxOrder - 1][2];  // [order-2][0 = prob, 1 = backoff] float tables
  MiddleView middle[kMaxOrder - 2];
  LongestView longest;
  uint32_t bos_word, eos_word;
  float bos_backoff;  // begin_sentence.backoff[0], model.cc:84
  // --- probing-hash models (model_type PROBING / REST_PROBING, kenlm/lm/search_hashed.hh): 0 = trie model
  uint32_t probing;          // 1 PROBING (ProbBackoff weights, 8 B), 2 REST_PROBING (RestWeights, 12 B)
  uint32_t weights_size;
  uint64_t pvocab_off, pvocab_buckets;           // ProbingVocabulary table: {u64 hash, u32 id} entries of 12 B
  uint64_t ptab_off[kMaxOrder - 1];              // [0 .. order-3] middle tables {u64 key, weights}; [order-2] longest {u64 key, f32}
  uint64_t ptab_buckets[kMaxOrder - 1];
  // --- 'TRIE' header
  uint32_t is_utf8;
  double alpha, beta;  // f32-rounded values held in f64 (scorer.cpp:346-351)
  // --- dictionary FST
  uint64_t fst_states_off, fst_arcs_off;
  int64_t fst_start, fst_nstates, fst_narcs;
  // --- alphabet (label -> UTF-8 bytes) for word hashing; labels longer than 4 bytes unsupported
  uint8_t label_len[256];
  uint8_t label_bytes[256][4];
  uint32_t space_label, n_labels;
};

// ------------------------------------------------------------------ raw reads
STT_HD uint64_t load_u64(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
  const uint64_t a = reinterpret_cast<uint64_t>(p);
  const uint64_t* w = reinterpret_cast<const uint64_t*>(a & ~7ull);
  const unsigned sh = (unsigned)(a & 7) * 8;
  uint64_t lo = __ldg(w);
  if (sh == 0) return lo;
  uint64_t hi = __ldg(w + 1);
  return (lo >> sh) | (hi << (64 - sh));
#else
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
#endif
}
STT_HD uint32_t load_u32(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)load_u64(p);
#else
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
#endif
}
STT_HD float load_f32(const uint8_t* p) { return sttmath::as_f32(load_u32(p)); }

STT_HD uint64_t read_int57(const uint8_t* base, uint64_t bit_off, uint64_t mask) {
  return (load_u64(base + (bit_off >> 3)) >> (bit_off & 7)) & mask;
}
STT_HD uint32_t read_int25(const uint8_t* base, uint64_t bit_off, uint32_t mask) {
  return (load_u32(base + (bit_off >> 3)) >> (bit_off & 7)) & mask;
}

// ------------------------------------------------------------------ vocabulary
STT_HD uint64_t murmur64a(const uint8_t* key, uint32_t len, uint64_t seed) {
  const uint64_t m = 0xc6a4a7935bd1e995ULL;
  const int r = 47;
  uint64_t h = seed ^ (len * m);
  uint32_t nblocks = len / 8;
  for (uint32_t i = 0; i < nblocks; ++i) {
    uint64_t k = 0;
    for (int b = 0; b < 8; ++b) k |= (uint64_t)key[i * 8 + b] << (8 * b);
    k *= m;
    k ^= k >> r;
    k *= m;
    h ^= k;
    h *= m;
  }
  const uint8_t* d = key + nblocks * 8;
  uint32_t rem = len & 7;
  if (rem) {
    for (uint32_t b = rem; b-- > 0;) h ^= (uint64_t)d[b] << (8 * b);
    h *= m;
  }
  h ^= h >> r;
  h *= m;
  h ^= h >> r;
  return h;
}

// SortedVocabulary::Index: exact search of the 64-bit hash; any exact search returns the same slot.
STT_HD uint32_t vocab_index_from_hash(const ScorerView& v, uint64_t hash) {
  if (v.probing) {
    // ProbingVocabulary::Index (vocab.hh:162-165): linear probing from hash % buckets, key 0 = empty (probing_hash_table.hh)
    const uint8_t* tab = v.blob + v.pvocab_off;
    uint64_t i = hash % v.pvocab_buckets;
    for (;;) {
      const uint64_t k = load_u64(tab + i * 12);
      if (k == hash) return load_u32(tab + i * 12 + 8);
      if (k == 0) return 0;  // <unk>
      if (++i == v.pvocab_buckets) i = 0;
    }
  }
  const uint8_t* tab = v.blob + v.vocab_off;
  uint64_t lo = 0, hi = v.vocab_count;
  while (lo < hi) {
    uint64_t mid = (lo + hi) >> 1;
    uint64_t k = load_u64(tab + mid * 8);
    if (k < hash) lo = mid + 1;
    else hi = mid;
  }
  if (lo < v.vocab_count && load_u64(tab + lo * 8) == hash) return (uint32_t)lo + 1;
  return 0;  // <unk>
}
STT_HD uint32_t vocab_index(const ScorerView& v, const uint8_t* word, uint32_t len) {
  return vocab_index_from_hash(v, murmur64a(word, len, 0));
}

// ------------------------------------------------------------------ trie
struct NodeRange {
  uint64_t begin, end;
};

STT_HD bool has_extension(float backoff) { return sttmath::as_u32(backoff) != 0x80000000u; }

STT_HD void unigram_find(const ScorerView& v, uint32_t word, NodeRange& next, float& prob, float& backoff) {
  const uint8_t* u = v.blob + v.unigram_off + (uint64_t)word * 16;
  prob = load_f32(u);
  backoff = load_f32(u + 4);
  next.begin = load_u64(u + 8);
  next.end = load_u64(u + 24);
}

// FindBitPacked (trie.cc:32-40 -> bit_packing / UniformFind, sorted_uniform.hh:60-90): records [begin, end) hold
// strictly increasing word ids; which probe sequence finds the record does not change the answer.  Like KenLM this is
// an interpolation search: word ids are ranks of 64-bit hashes, so the ids of a node's children are uniform in
// [0, max_word] and ~log log n probes suffice where bisection needs log n -- every probe is a dependent memory access.
STT_HD bool find_bit_packed(const uint8_t* base, uint64_t word_mask, uint8_t total_bits, uint64_t begin,
                            uint64_t end, uint64_t key, uint64_t max_word, uint64_t& at) {
  // invariant: key can only be at an index in (before_it, after_it); record(before_it) = before_v < key < after_v
  uint64_t before_it = begin - 1, after_it = end;   // begin >= 0; "begin - 1" is only ever compared, never read
  uint64_t before_v = 0, after_v = max_word;
  if (key > max_word) return false;
  while (after_it - before_it > 1) {
    const uint64_t span = after_it - before_it - 1;
    uint64_t pivot;
    if (span < (1ull << 31) && after_v - before_v < (1ull << 32)) {
      pivot = before_it + 1 + (key - before_v) * span / (after_v - before_v + 1);
    } else {
      pivot = before_it + 1 + (span >> 1);
    }
    const uint64_t k = read_int57(base, pivot * total_bits, word_mask);
    if (k < key) {
      before_it = pivot;
      before_v = k;
    } else if (k > key) {
      after_it = pivot;
      after_v = k;
    } else {
      at = pivot;
      return true;
    }
  }
  return false;
}

STT_HD void read_next(const ScorerView& v, const MiddleView& m, uint64_t bit_offset, uint64_t index, NodeRange& out) {
  const uint8_t* base = v.blob + m.records_off;
  uint64_t b = read_int57(base, bit_offset, m.next_mask);
  uint64_t e = read_int57(base, bit_offset + m.total_bits, m.next_mask);
  if (v.bhiksha) {
    // ArrayBhiksha::ReadNext: last offset <= index, and last offset <= index + 1.
    const uint8_t* offs = v.blob + m.offsets_off;
    // upper_bound(index).  offsets[k] is the first record whose next pointer has high part k, and next pointers grow
    // about linearly with the record index, so an interpolated guess lands within a probe or two (every probe is a
    // dependent memory access; bisection over <= 256 offsets cost 8 of them per n-gram level).
    uint64_t lo = 0, hi = m.offsets_count;
    uint64_t vlo = 0, vhi = m.n_records + 1;
    for (int it = 0; lo < hi; ++it) {
      uint64_t est;
      if (it < 6 && vhi > vlo) est = lo + ((index >= vlo ? index - vlo : 0) * (hi - lo)) / (vhi - vlo + 1);
      else est = lo + ((hi - lo) >> 1);
      if (est >= hi) est = hi - 1;
      const uint64_t x = load_u64(offs + est * 8);
      if (x <= index) { lo = est + 1; vlo = x; }
      else { hi = est; vhi = x; }
    }
    uint64_t begin_it = lo - 1;
    uint64_t end_it = begin_it + 1;
    while (end_it < m.offsets_count && load_u64(offs + end_it * 8) <= index + 1) ++end_it;
    --end_it;
    b |= begin_it << m.next_bits;
    e |= end_it << m.next_bits;
  }
  out.begin = b;
  out.end = e;
}

// BitPackedMiddle::Find + the Quant::MiddlePointer reads.  On a hit fills prob/backoff and narrows
// `node` to the children range.
STT_HD bool middle_find(const ScorerView& v, int order_minus_2, uint32_t word, NodeRange& node, float& prob,
                        float& backoff) {
  const MiddleView& m = v.middle[order_minus_2];
  const uint8_t* base = v.blob + m.records_off;
  uint64_t at;
  if (!find_bit_packed(base, m.word_mask, m.total_bits, node.begin, node.end, word, v.vocab_count + 1, at)) return false;
  uint64_t bit = at * m.total_bits + m.word_bits;
  if (v.quantized) {
    const uint32_t bmask = (1u << v.backoff_bits) - 1, pmask = (1u << v.prob_bits) - 1;
    uint32_t bcode = read_int25(base, bit, bmask);
    uint32_t pcode = read_int25(base, bit + v.backoff_bits, pmask);
    prob = load_f32(v.blob + v.quant_tables_off[order_minus_2][0] + 4ull * pcode);
    backoff = load_f32(v.blob + v.quant_tables_off[order_minus_2][1] + 4ull * bcode);
  } else {
    uint32_t p31 = (uint32_t)(load_u64(base + (bit >> 3)) >> (bit & 7)) | 0x80000000u;
    prob = sttmath::as_f32(p31);
    uint64_t b2 = bit + 31;
    backoff = sttmath::as_f32((uint32_t)(load_u64(base + (b2 >> 3)) >> (b2 & 7)));
  }
  read_next(v, m, bit + m.quant_bits, at, node);
  return true;
}

STT_HD bool longest_find(const ScorerView& v, uint32_t word, const NodeRange& node, float& prob) {
  const LongestView& l = v.longest;
  const uint8_t* base = v.blob + l.records_off;
  uint64_t at;
  if (!find_bit_packed(base, l.word_mask, l.total_bits, node.begin, node.end, word, v.vocab_count + 1, at)) return false;
  uint64_t bit = at * l.total_bits + l.word_bits;
  if (v.quantized) {
    uint32_t pcode = read_int25(base, bit, (1u << v.prob_bits) - 1);
    prob = load_f32(v.blob + v.quant_tables_off[v.order - 2][0] + 4ull * pcode);
  } else {
    prob = sttmath::as_f32((uint32_t)(load_u64(base + (bit >> 3)) >> (bit & 7)) | 0x80000000u);
  }
  return true;
}

// ------------------------------------------------------------------ probing-hash search (kenlm/lm/search_hashed.hh)
STT_HD uint64_t combine_word_hash(uint64_t current, uint32_t next) {   // detail::CombineWordHash :26-29
  return (current * 8978948897894561157ULL) ^ ((uint64_t)(1 + next) * 17894857484156487943ULL);
}
// ProbingHashTable::Find with IdentityHash / DivMod: entry address or null
STT_HD const uint8_t* probing_find(const ScorerView& v, int table, uint32_t entry_size, uint64_t key) {
  const uint8_t* tab = v.blob + v.ptab_off[table];
  const uint64_t buckets = v.ptab_buckets[table];
  uint64_t i = key % buckets;
  for (;;) {
    const uint64_t k = load_u64(tab + i * entry_size);
    if (k == key) return tab + i * entry_size;
    if (k == 0) return nullptr;
    if (++i == buckets) i = 0;
  }
}
// GenericModel<HashedSearch<...>>::FullScore: the same ResumeScore walk (model.cc:285-338) with the hashed lookups
// (LookupUnigram / LookupMiddle / LookupLongest, search_hashed.hh:96-129); a stored probability's SIGN BIT says whether
// the n-gram extends to the left (GenericProbingProxy, value.hh:15-38), its value is the number with the sign bit set.
STT_HD float full_score_probing(const ScorerView& v, const LmState& in, uint32_t new_word, LmState& out) {
  const uint8_t* u = v.blob + v.unigram_off + (uint64_t)new_word * v.weights_size;
  const uint32_t pu = load_u32(u);
  float prob = sttmath::as_f32(pu | 0x80000000u);
  const float bo = load_f32(u + 4);
  bool independent_left = (pu & 0x80000000u) != 0;
  uint64_t node = (uint64_t)new_word;
  uint8_t ngram_length = 1;
  out.backoff[0] = bo;
  out.length = has_extension(bo) ? 1 : 0;
  out.words[0] = new_word;
  if (in.length != 0) {
    int order_minus_2 = 0;
    int hist = 0;
    float* backoff_out = out.backoff + 1;
    bool broke = false;
    const uint32_t mid_entry = 8 + v.weights_size;
    for (;; ++order_minus_2, ++hist, ++backoff_out) {
      if (hist == in.length) break;
      if (independent_left) break;
      if (order_minus_2 == (int)v.order - 2) {
        broke = true;
        break;
      }
      node = combine_word_hash(node, in.words[hist]);
      const uint8_t* e = probing_find(v, order_minus_2, mid_entry, node);
      if (!e) break;   // independent_left = true; not found
      const uint32_t pm = load_u32(e + 8);
      independent_left = (pm & 0x80000000u) != 0;
      const float b = load_f32(e + 12);
      *backoff_out = b;
      prob = sttmath::as_f32(pm | 0x80000000u);
      ngram_length = (uint8_t)(order_minus_2 + 2);
      if (has_extension(b)) out.length = ngram_length;
    }
    if (broke) {
      const uint8_t* e = probing_find(v, (int)v.order - 2, 12, combine_word_hash(node, in.words[hist]));
      if (e) {
        prob = load_f32(e + 8);
        ngram_length = (uint8_t)v.order;
      }
    }
    for (int i = 0; i + 1 < (int)out.length; ++i) out.words[i + 1] = in.words[i];
  }
  for (int i = ngram_length - 1; i < (int)in.length; ++i) prob += in.backoff[i];
  return prob;
}

// GenericModel::FullScore (model.cc:170-176) = ScoreExceptBackoff + charged backoffs.
STT_HD float full_score(const ScorerView& v, const LmState& in, uint32_t new_word, LmState& out) {
  if (v.probing) return full_score_probing(v, in, new_word, out);
  NodeRange node;
  float prob, bo;
  unigram_find(v, new_word, node, prob, bo);
  bool independent_left = (node.begin == node.end);
  uint8_t ngram_length = 1;
  out.backoff[0] = bo;
  out.length = has_extension(bo) ? 1 : 0;
  out.words[0] = new_word;
  if (in.length != 0) {
    // ResumeScore (model.cc:312-338)
    int order_minus_2 = 0;
    int hist = 0;
    float* backoff_out = out.backoff + 1;
    bool broke = false;
    for (;; ++order_minus_2, ++hist, ++backoff_out) {
      if (hist == in.length) break;
      if (independent_left) break;
      if (order_minus_2 == (int)v.order - 2) {
        broke = true;
        break;
      }
      float p, b;
      bool found = middle_find(v, order_minus_2, in.words[hist], node, p, b);
      independent_left = !found || (node.begin == node.end);
      if (!found) break;
      *backoff_out = b;
      prob = p;
      ngram_length = (uint8_t)(order_minus_2 + 2);
      if (has_extension(b)) out.length = ngram_length;
    }
    if (broke) {
      float p;
      if (longest_find(v, in.words[hist], node, p)) {
        prob = p;
        ngram_length = (uint8_t)v.order;
      }
    }
    // CopyRemainingHistory (model.cc:272-277)
    for (int i = 0; i + 1 < (int)out.length; ++i) out.words[i + 1] = in.words[i];
  }
  for (int i = ngram_length - 1; i < (int)in.length; ++i) prob += in.backoff[i];
  return prob;
}

STT_HD void begin_sentence_state(const ScorerView& v, LmState& s) {
  s.length = 1;
  s.words[0] = v.bos_word;
  s.backoff[0] = v.bos_backoff;
}
STT_HD void null_context_state(LmState& s) { s.length = 0; }

// Scorer::get_log_cond_prob (scorer.cpp:307-344) over word ids (0 = OOV).  Returns the natural-log
// conditional probability of the last word, or OOV_SCORE (-1000, scorer.h:16) if any word is OOV.
STT_HD double log_cond_prob_ids(const ScorerView& v, const uint32_t* ids, int n, bool bos) {
  LmState st[2];
  int cur = 0;
  if (bos) begin_sentence_state(v, st[0]);
  else null_context_state(st[0]);
  double cond_prob = 0.0;
  for (int i = 0; i < n; ++i) {
    if (ids[i] == 0) return -1000.0;
    cond_prob = (double)full_score(v, st[cur], ids[i], st[cur ^ 1]);
    cur ^= 1;
  }
  return cond_prob / (double)0.4342944819f;  // NUM_FLT_LOGE, decoder_utils.h:13
}

// ------------------------------------------------------------------ dictionary FST
STT_HD bool fst_is_final(const ScorerView& v, int32_t state) {
  // TropicalWeight::Zero() == +inf
  return sttmath::as_u32(load_f32(v.blob + v.fst_states_off + (uint64_t)state * 20)) != 0x7f800000u;
}
// SortedMatcher::Find(label) on `state` (MATCH_INPUT, label > 0): next state or -1.
STT_HD int32_t fst_find(const ScorerView& v, int32_t state, int32_t label) {
  const uint8_t* s = v.blob + v.fst_states_off + (uint64_t)state * 20;
  uint32_t pos = load_u32(s + 4), narcs = load_u32(s + 8);
  const uint8_t* arcs = v.blob + v.fst_arcs_off + (uint64_t)pos * 16;
  uint32_t lo = 0, hi = narcs;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    int32_t il = (int32_t)load_u32(arcs + (uint64_t)mid * 16);
    if (il < label) lo = mid + 1;
    else hi = mid;
  }
  if (lo < narcs && (int32_t)load_u32(arcs + (uint64_t)lo * 16) == label)
    return (int32_t)load_u32(arcs + (uint64_t)lo * 16 + 12);
  return -1;
}

}  // namespace sttscorer
