// K7/K8/K9: CTC prefix beam search with KenLM scorer + dictionary FST on the GPU, one CTA per utterance.
//
// Restates DecoderState::{init,next,decode} (native_client/ctcdecode/ctc_beam_search_decoder.cpp:22-61,112-276,
// 278-326), PathTrie (path_trie.cpp:37-100,159-209), prefix_compare (decoder_utils.cpp:66-88),
// Scorer::{is_scoring_boundary,make_ngram,get_log_cond_prob} (scorer.cpp:271-344,369-396) with the exact
// arithmetic types of SURVEY.md appendix A (f32 log-probs through glibc-exact logf/expf, f64 only where the
// reference uses it).  The data structure is NOT a port:
//
//   (v2: live lists and per-step candidates live in SHARED memory, the dictionary FST is pre-digested into per-state
//   label masks + pre-resolved next states, arena nodes are 32-byte records, LM results are cached per node.)
//   * The reference keeps a pointer trie and walks it depth-first every step.  Semantically each step maps a SET of
//     <= beam live prefixes to the top-beam of (live prefixes U their one-character extensions); a pruned node that
//     is later re-created gets exactly the values a fresh node gets (path_trie.cpp:45-52).  So we keep a flat
//     append-only ARENA of surviving nodes (parent, char, dictionary state, last-space link, word id), a
//     double-buffered LIVE list (score, b_prev, nb_prev, node, timestep-node) and per-step CANDIDATE arrays.
//   * One thread owns one live prefix.  Extensions that land on an already-live child are PULLED by the child's
//     thread from its parent (found through a node->live-slot table), so no two threads ever write the same prefix.
//   * Children of a prefix are enumerated by reading its dictionary-FST state's arc list (the allowed labels)
//     instead of probing all 28 labels with a matcher.
//   * Top-beam selection is an exact CTA-wide radix select on a 64-bit key (score, then the reference's
//     tie-break on character, then candidate index: a deterministic total order that agrees with prefix_compare
//     wherever prefix_compare is decisive), followed by an order-preserving compaction.
//   * LM calls of a step are gathered into a work list and run one-per-thread spread across warps.
//
// Tie caveat (SURVEY 7 hard part 1): prefix_compare is not a total order; when two distinct prefixes tie on score
// AND last character at the beam boundary the reference keeps whichever std::nth_element leaves first.  We keep the
// lower candidate index.  Such ties between live, finite-score prefixes need bit-equal f32 sums and are reported by
// the parity tests if they ever occur.
#pragma once
#include <stdint.h>

#include "hd_math.h"
#include "scorer_view.h"

namespace sttdec {

constexpr uint32_t kNone = 0xffffffffu;
constexpr uint32_t kRootChar = 0xffu;
constexpr float kNegMax = -3.402823466e+38f;  // -NUM_FLT_INF (decoder_utils.h:11)
constexpr float kFltMin = 1.175494351e-38f;   // NUM_FLT_MIN
constexpr int kMaxClasses = 64;
constexpr int kMaxWordBytes = 128;
constexpr int kStateWords = sttscorer::kMaxOrder - 1;
constexpr int kMaxHotWords = 32;
constexpr int kCommitSpan = 4096;  // candidates compacted per scan in phase 6 (4096 / NT rounds of NT)
constexpr int kSelBoundaryCap = 256;  // phase 5: boundary-bin elements resolved by pairwise ranking (more: radix passes)
// "not computed" marker of Slot::lm_cond (a quiet NaN no LM result can equal)
constexpr unsigned long long kLmUnset = 0x7ff8dead5117b200ull;
constexpr int kFlagHistSelect = 1;   // DecodeParams::flags (0: radix passes only -- kept for A/B measurements)

struct Node {            // one surviving prefix (PathTrie node), 32 bytes
  uint32_t parent;       // arena id, kNone for the root
  uint32_t chr;          // label, kRootChar for the root
  int32_t dict;          // dictionary-FST state AFTER this label (already reset to Start() after a final state)
  uint32_t last_space;   // nearest ancestor-or-self whose label is the space, or kNone
  union {
    uint32_t word_id;    // space nodes: vocabulary id of the word they terminate
    uint32_t ord;        // other nodes: ordinal of the partial word's path in the dictionary FST (DecodeParams); root 0
  };
  uint32_t live_slot;    // index in the current live list, or kNone
  uint32_t lm_wid;       // LM cache valid flag / word id of the word ending here (kNone = not computed)
  uint32_t child_mask;   // labels for which a child node has ever been created
};

// Per-utterance device state ("stream slot").  All pointers are device memory sized for (beam_cap, t_cap).
struct Slot {
  Node* nodes;           // [arena_cap]; node 0 is the root
  // A pruned node that still has live descendants must be REVIVED under its old id when its prefix re-enters the beam
  // (path_trie.cpp:45-52), so that those descendants keep merging into it: every node ever created is findable in
  // a (parent, label) -> node hash table.  Node::child_mask says whether a (parent, label) child was ever created, so
  // the common case -- a brand-new child -- needs no lookup at all, only an insert (one atomicMax, see ht_insert).
  unsigned long long* ht;  // [ht_mask + 1] packed {generation:8 | parent:24 | label:8 | node:24}; stale generation = empty
  uint32_t ht_mask;
  uint32_t ht_gen;         // 1..255, bumped by the host on every reset (tables are cleared when it wraps)
  // Per-node LM cache, keyed by the node that ENDS a word: the natural-log conditional probability of that word given
  // its history (Scorer::get_log_cond_prob's return value), its vocabulary id, and the KenLM state after it.  The
  // reference recomputes the whole <=order-word window on every call (scorer.cpp:307-344, 369-396); scoring the last
  // word from the carried state is value-identical (the KenLM state after k words is exactly the context the next
  // lookup may use; from-BOS carry == the window's BeginSentence chain when the window holds the whole history, and
  // the <=order-1 words a state can hold all lie inside the window otherwise), and a node's word/history never change.
  double* lm_cond;       // [arena_cap]
  uint32_t* lm_sw;       // [arena_cap * kStateWords] state words
  float* lm_sb;          // [arena_cap * kStateWords] state backoffs
  uint32_t* lm_meta;     // [arena_cap]  state length | oov distance << 8 | words in history (saturating) << 16
  // timestep tree (path_trie.h:17-37): ts node 0 is the root
  uint2* ts_tree;        // [ts_cap] {parent, absolute timestep}
  // live list as left by the last launch (the step kernel works on a shared-memory copy)
  float *score, *b_prev, *nb_prev;   // [beam_cap]
  uint32_t *node, *ts;               // [beam_cap]
  // per-step candidates when they do not fit in shared memory, and scratch for finalize; capacity cand_cap
  unsigned long long* c_key;
  uint32_t *c_p0, *c_p1;
  uint32_t* aux;                     // [4 * beam_cap] see StepSmem::aux
  unsigned long long* phase_cycles;  // [8] instrumentation: SM cycles per phase (thread 0's clock)
  // scalars: 0 n_live, 2 arena_count, 3 ts_count, 4 abs_time_step, 5 start_expanding, 6 overflow,
  //          7 LM words scored (reference-equivalent window sizes), 8 LM calls, 9 max candidates in a step,
  //          10 steps that spilled candidates to global memory
  uint32_t* scalars;     // [16]
  uint32_t arena_cap, ts_cap, beam_cap, cand_cap;
};

struct DecodeParams {
  int n_classes;         // alphabet + 1; blank = n_classes - 1
  int beam;              // <= beam_cap
  int space_id;
  int has_scorer;
  sttscorer::ScorerView scorer;
  // dictionary FST, pre-digested on the host (engine.cu build_fst_tables): per state {first arc, bit mask of the labels
  // that have an arc}, per arc {ilabel, dictionary state of the child = Start() if the arc's target is final}
  const uint2* fst_state2;
  const int4* fst_arc4;            // per arc {child state, child's first arc, child's label mask, word-ordinal skip}
  // Word ordinals (perfect hash of the acyclic dictionary FST, built by engine.cu): the words of the FST in
  // label-lexicographic order are numbered 0..n-1; ordinal(word) = sum over its arcs of fst_arc4[].w.  A node carries the
  // partial sum, so the KenLM id of the word that ends at a node is ONE table read instead of walking the prefix back
  // to the last space, hashing the bytes and binary-searching the vocabulary (Scorer::make_ngram + Vocabulary::Index,
  // scorer.cpp:307-344).  Null when the FST is not an acyclic acceptor whose words end with the space label: the
  // decoder then takes the walking path everywhere.
  const uint32_t* fst_space_skip;  // [n_states] skip of the state's space arc when that arc ends a word, else kNone
  const uint32_t* ord2wid;         // [n_words] KenLM vocabulary id (0 = <unk>)
  // hot words (ctc_beam_search_decoder.cpp:224-236): vocabulary ids and boosts, snapshotted when the stream starts
  int flags;             // kFlagHistSelect
  int n_hot;
  uint32_t hot_id[kMaxHotWords];
  float hot_boost[kMaxHotWords];
};

struct StepInput {
  const float* probs;     // [T, n_classes] f32 softmax rows for this slot (the C API path: stt.cc:327 widens f32 to f64)
  const double* probs64;  // or f64 rows as the Python decoder API passes them (swigwrapper.i:39-41); the class
                          // log-prob uses (float)p (:333), the gate and min_cutoff use the double (:125,:143)
  int n_steps;
};

// ------------------------------------------------------------------------------------------------ helpers
// (parent, label) -> node.  One 64-bit word per entry, generation in the top byte, so that
//   * a slot left over from an earlier decode (lower generation) reads as empty -- no per-decode memset of the tables;
//   * an insert is a single atomicMax: against an empty/stale slot it simply wins.  Against a live entry of the current
//     generation the larger word stays and the thread carries the smaller one to the next slot (linear probing; an entry
//     only ever moves forward along its own probe path, so "present between the home slot and the first empty slot"
//     keeps holding).
// Lookups happen only for children that are known to exist (Node::child_mask), so a lookup that runs into an empty slot
// has met an entry in transit between two slots and simply starts over.
__device__ __forceinline__ uint32_t ht_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  return (uint32_t)k;
}
__device__ __forceinline__ unsigned long long ht_pack(uint32_t gen, uint32_t parent, uint32_t c, uint32_t node) {
  return ((unsigned long long)gen << 56) | ((unsigned long long)(parent & 0xffffffu) << 32) |
         ((unsigned long long)(c & 0xffu) << 24) | (unsigned long long)(node & 0xffffffu);
}
__device__ __forceinline__ void ht_insert(const Slot& s, uint32_t parent, uint32_t c, uint32_t node) {
  unsigned long long w = ht_pack(s.ht_gen, parent, c, node);
  uint32_t h = ht_hash(w >> 24) & s.ht_mask;
  for (;;) {
    const unsigned long long old = atomicMax(&s.ht[h], w);
    if ((uint32_t)(old >> 56) != s.ht_gen) return;  // the slot was empty or stale
    if (old > w) {
      // the resident entry stays; ours moves on
    } else {
      w = old;  // we displaced the resident entry: carry it forward
    }
    h = (h + 1) & s.ht_mask;
  }
}
__device__ __forceinline__ uint32_t ht_find_existing(const Slot& s, uint32_t parent, uint32_t c) {
  const unsigned long long key = ht_pack(s.ht_gen, parent, c, 0) >> 24;
  for (;;) {
    uint32_t h = ht_hash(key) & s.ht_mask;
    for (;;) {
      const unsigned long long w = *reinterpret_cast<volatile unsigned long long*>(&s.ht[h]);
      if ((w >> 24) == key) return (uint32_t)(w & 0xffffffu);
      if ((uint32_t)(w >> 56) != s.ht_gen) break;  // in transit: retry from the home slot
      h = (h + 1) & s.ht_mask;
    }
  }
}
__device__ __forceinline__ uint32_t sortable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unsortable(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ unsigned long long make_key(float score, uint32_t chr, uint32_t idx) {
  // descending order == the reference's prefix_compare (score desc, then character asc), then candidate index asc
  return ((unsigned long long)sortable(score) << 32) | ((unsigned long long)(255u - (chr & 0xffu)) << 24) |
         (unsigned long long)(0xffffffu - (idx & 0xffffffu));
}
// Tie-break field of a NEW candidate's key: candidates are laid out live prefixes first, then new children ordered by
// (parent's live index, label), so 0x1000 + parent * 32 + label grows with the candidate's position -- the same total
// order as the position itself -- and lets the commit phase recover (parent, label) from the key alone.
__device__ __forceinline__ uint32_t new_cand_idx(uint32_t parent_live, uint32_t label) { return 0x1000u + parent_live * 32u + label; }
// order in which the reference's sorted loop visits two live prefixes: true if x comes before y
__device__ __forceinline__ bool visits_before(float sx, uint32_t cx, uint32_t ix, float sy, uint32_t cy, uint32_t iy) {
  if (sx != sy) return sx > sy;
  if (cx != cy) return cx < cy;
  return ix < iy;
}

// Scorer::make_ngram + get_log_cond_prob, literally as the reference does it (walk back <= order words, score the
// window from BeginSentence / NullContext).  Fallback when a history node has no cached state.
__device__ double lm_window_cond(const Slot& s, const sttscorer::ScorerView& v, uint32_t node, uint32_t* word_out,
                                 uint32_t* n_words_out) {
  uint32_t ids_rev[sttscorer::kMaxOrder];
  int n = 0;
  uint32_t first_word = 0;
  const int order = (int)v.order;
  uint32_t cur = node, term = kNone;
  while (n < order) {
    const uint32_t cc = s.nodes[cur].chr;
    if (cc == kRootChar) break;
    const uint32_t stop = s.nodes[cur].last_space;  // == cur when cur is itself a space (empty word)
    uint32_t id;
    if (cc == (uint32_t)v.space_label) {
      id = 0;  // get_prev_word returns an empty word: never in the vocabulary
    } else if (term == kNone) {
      uint8_t buf[kMaxWordBytes];
      int len = 0;
      bool too_long = false;
      for (uint32_t w = cur; w != stop && s.nodes[w].chr != kRootChar; w = s.nodes[w].parent) {
        const uint32_t c = s.nodes[w].chr;
        const int l = v.label_len[c];
        if (len + l > kMaxWordBytes) { too_long = true; break; }
        for (int q = l - 1; q >= 0; --q) buf[len++] = v.label_bytes[c][q];
      }
      for (int a = 0, b = len - 1; a < b; ++a, --b) { const uint8_t t = buf[a]; buf[a] = buf[b]; buf[b] = t; }
      id = too_long ? 0u : sttscorer::vocab_index(v, buf, (uint32_t)len);
      first_word = id;
    } else {
      id = s.nodes[term].word_id;
    }
    ids_rev[n++] = id;
    if (stop == kNone) break;
    term = stop;
    cur = s.nodes[stop].parent;
  }
  *word_out = first_word;
  *n_words_out = (uint32_t)n;
  uint32_t ids[sttscorer::kMaxOrder];
  for (int i = 0; i < n; ++i) ids[i] = ids_rev[n - 1 - i];
  return sttscorer::log_cond_prob_ids(v, ids, n, n < order);
}

constexpr uint32_t kStopUnknown = 0xfffffffdu;  // lm_eval_node: the caller does not have the node's last_space at hand

// Cached evaluation of get_log_cond_prob(make_ngram(prefix `node`), bos) -- see Slot::lm_cond.
//   * the id of the word ending at `node` was resolved when the node was created (Node::lm_wid, phase 6);
//   * the KenLM state BEFORE that word -- the state after the previous word -- was copied into the lm_* arrays of the
//     space node that separates the two words when that space node was created (phase 6), so it sits at index `stop`
//     (= the node's last_space, which the caller has in shared memory): every load below is independent of the others
//     and a first evaluation starts the trie descent after ONE round trip.
__device__ double lm_eval_node(const Slot& s, const DecodeParams& p, uint32_t node, uint32_t stop_hint, uint32_t* word_out,
                               uint32_t* n_window_out, uint32_t* miss_out = nullptr) {
  const sttscorer::ScorerView& v = p.scorer;
  const int order = (int)v.order;
  // lm_cond is both the cache's valid flag (kLmUnset = not computed) and its payload
  const unsigned long long c0bits = reinterpret_cast<const unsigned long long*>(s.lm_cond)[node];
  const uint32_t meta0 = s.lm_meta[node];   // space nodes: the carried context, see above
  const double cond0 = __longlong_as_double((long long)c0bits);
  const Node nd = s.nodes[node];
  uint32_t stop = stop_hint;
  uint32_t cmeta = kNone;
  uint32_t csw[kStateWords];
  float csb[kStateWords];
  if (stop_hint != kStopUnknown && stop_hint != kNone) {
    cmeta = s.lm_meta[stop_hint];
#pragma unroll
    for (int i = 0; i < kStateWords; ++i) {
      csw[i] = s.lm_sw[(size_t)stop_hint * kStateWords + i];
      csb[i] = s.lm_sb[(size_t)stop_hint * kStateWords + i];
    }
  }
  const uint32_t cc = nd.chr;
  if (cc == kRootChar) {
    // empty prefix: make_ngram returns no words, get_log_cond_prob of nothing is 0 (scorer.cpp:325,343)
    *word_out = 0;
    *n_window_out = 0;
    return 0.0;
  }
  if (stop == kStopUnknown) {
    stop = nd.last_space;  // == node when node is itself a space (empty word)
    if (stop != kNone) {
      cmeta = s.lm_meta[stop];
#pragma unroll
      for (int i = 0; i < kStateWords; ++i) {
        csw[i] = s.lm_sw[(size_t)stop * kStateWords + i];
        csb[i] = s.lm_sb[(size_t)stop * kStateWords + i];
      }
    }
  }
  if (cc == (uint32_t)v.space_label) {
    // a prefix that ends with a space names an empty word: never in the vocabulary, OOV_SCORE undivided
    // (scorer.cpp:328-331).  Nothing is cached: this node's lm_* arrays hold the context for the NEXT word.
    const uint32_t cw = (cmeta == kNone) ? 0u : (cmeta >> 16);
    const uint32_t nw = cw >= 0xfffeu ? 0xffffu : cw + 1;
    *word_out = 0;
    *n_window_out = nw < (uint32_t)order ? nw : (uint32_t)order;
    return -1000.0;
  }
  if (c0bits != kLmUnset) {
    *word_out = nd.lm_wid;
    const uint32_t nw = meta0 >> 16;
    *n_window_out = nw < (uint32_t)order ? nw : (uint32_t)order;
    return cond0;
  }
  // ---- the word ending at `node`
  if (miss_out) *miss_out = 1;
  uint32_t wid = nd.lm_wid;
  if (wid == kNone) {
    uint32_t sk = kNone;
    if (p.fst_space_skip) sk = __ldg(p.fst_space_skip + nd.dict);
    if (sk != kNone) {
      wid = __ldg(p.ord2wid + nd.ord + sk);
    } else {
      uint8_t buf[kMaxWordBytes];
      int len = 0;
      bool too_long = false;
      uint32_t w = node;
      Node wn = nd;
      while (w != stop && wn.chr != kRootChar) {
        const int l = v.label_len[wn.chr];
        if (len + l > kMaxWordBytes) { too_long = true; break; }
        for (int q = l - 1; q >= 0; --q) buf[len++] = v.label_bytes[wn.chr][q];
        w = wn.parent;
        if (w == kNone) break;
        wn = s.nodes[w];
      }
      for (int a = 0, b = len - 1; a < b; ++a, --b) { const uint8_t t = buf[a]; buf[a] = buf[b]; buf[b] = t; }
      wid = too_long ? 0u : sttscorer::vocab_index(v, buf, (uint32_t)len);
    }
  }
  // ---- history state
  sttscorer::LmState ctx;
  uint32_t ctx_oov = 255, ctx_words = 0;
  if (stop == kNone) {
    sttscorer::begin_sentence_state(v, ctx);
  } else if (cmeta != kNone) {
    ctx.length = (uint8_t)(cmeta & 0xffu);
    ctx_oov = (cmeta >> 8) & 0xffu;
    ctx_words = cmeta >> 16;
#pragma unroll
    for (int i = 0; i < kStateWords; ++i) {
      ctx.words[i] = csw[i];
      ctx.backoff[i] = csb[i];
    }
  } else {
    return lm_window_cond(s, v, node, word_out, n_window_out);  // no carried state (not expected): the literal way
  }
  double cond;
  sttscorer::LmState out;
  out.length = 0;
  uint32_t oov_dist;
  if (wid == 0) {
    sttscorer::null_context_state(out);
    oov_dist = 0;
    cond = -1000.0;  // OOV_SCORE, returned undivided (scorer.cpp:328-331)
  } else {
    const float p10 = sttscorer::full_score(v, ctx, wid, out);
    oov_dist = ctx_oov >= 254 ? 255u : ctx_oov + 1;
    // an OOV word among the previous order-1 words is still inside the reference's window
    cond = (oov_dist <= (uint32_t)(order - 1)) ? -1000.0 : (double)p10 / (double)0.4342944819f;
  }
  const uint32_t nwords = ctx_words >= 0xfffeu ? 0xffffu : ctx_words + 1;
  const uint32_t meta = (uint32_t)out.length | (oov_dist << 8) | (nwords << 16);
  for (int i = 0; i < (int)out.length; ++i) {
    s.lm_sw[(size_t)node * kStateWords + i] = out.words[i];
    s.lm_sb[(size_t)node * kStateWords + i] = out.backoff[i];
  }
  s.lm_meta[node] = meta;
  if (nd.lm_wid == kNone) s.nodes[node].lm_wid = wid;
  s.lm_cond[node] = cond;
  *word_out = wid;
  const uint32_t nw = meta >> 16;
  *n_window_out = nw < (uint32_t)order ? nw : (uint32_t)order;
  return cond;
}

// Hot-word boost of one LM call: every word of the <=order-word window that is a hot word adds its boost
// (ctc_beam_search_decoder.cpp:224-236; the window is make_ngram's, walked through the cached word ids).
__device__ float hot_word_boost(const Slot& s, const DecodeParams& p, uint32_t node, uint32_t first_wid) {
  float per_word[sttscorer::kMaxOrder];
  int n = 0;
  const int order = (int)p.scorer.order;
  uint32_t wid = first_wid;  // id of the word ending at `cur`
  uint32_t cur = node;
  while (n < order) {
    const Node nd = s.nodes[cur];
    if (nd.chr == kRootChar) break;
    float b = 0.0f;
    bool hit = false;
    if (wid != 0)
      for (int h = 0; h < p.n_hot; ++h)
        if (p.hot_id[h] == wid) { b = p.hot_boost[h]; hit = true; }
    per_word[n++] = hit ? b : 0.0f;
    const uint32_t stop = nd.last_space;  // the space before this word
    if (stop == kNone) break;
    const Node sp = s.nodes[stop];
    wid = sp.word_id;   // a space node carries the id of the word it terminates = the previous word of the window
    cur = sp.parent;    // ... which ends at the space's parent
  }
  // the reference adds the boosts in n-gram order, oldest word first (float accumulation)
  float boost = 0.0f;
  for (int i = n - 1; i >= 0; --i)
    if (per_word[i] != 0.0f) boost += per_word[i];
  return boost;
}

// ------------------------------------------------------------------------------------------------ UTF-8 mode helpers
// (bytes-output scorers; used by the general kernel of decoder_general.cuh and by the finalize kernel below)
// A node of a UTF-8 decode keeps, in the two words the word mode uses for its space bookkeeping:
//   Node::last_space  the CONTEXT node of the code point this byte belongs to = the parent of that code point's lead
//                     byte (the node where the previous code point ended, or the root)
//   Node::ord         distance_to_codepoint_boundary (bytes since the lead byte, inclusive) | lead byte << 8
__device__ __forceinline__ int utf8_needed_bytes(uint32_t lead) {   // scorer.cpp:280-292
  if ((lead >> 3) == 0x1Eu) return 4;
  if ((lead >> 4) == 0x0Eu) return 3;
  if ((lead >> 5) == 0x06u) return 2;
  if ((lead >> 7) == 0x00u) return 1;
  return 0;   // invalid lead byte: never a boundary
}
__device__ __forceinline__ bool utf8_is_lead(uint32_t byte) { return (byte & 0xC0u) != 0x80u; }   // byte_is_codepoint_boundary
__device__ __forceinline__ bool utf8_completes(uint32_t ord) { return utf8_needed_bytes((ord >> 8) & 0xffu) == (int)(ord & 0xffu); }
// cp fields of the child (parent record, label)
__device__ __forceinline__ void utf8_child_fields(const sttscorer::ScorerView& v, const Node& par, uint32_t par_id, uint32_t c,
                                                  uint32_t* ord_out, uint32_t* ctx_out) {
  const uint32_t byte = v.label_bytes[c][0];
  if (utf8_is_lead(byte) || par.chr == kRootChar) {
    *ord_out = 1u | (byte << 8);
    *ctx_out = par_id;
  } else {
    *ord_out = (((par.ord & 0xffu) + 1u) & 0xffu) | (par.ord & 0xff00u);
    *ctx_out = par.last_space;
  }
}

// Vocabulary id of the bytes of the (possibly incomplete) code point that ends at `node`, optionally followed by the
// bytes of one more label `extra` (a candidate child that does not exist yet; kNone = none).
__device__ uint32_t utf8_unit_id(const Slot& s, const sttscorer::ScorerView& v, uint32_t node, uint32_t n_back, uint32_t extra) {
  uint8_t buf[40];
  int len = 0;
  if (extra != kNone)
    for (int q = (int)v.label_len[extra] - 1; q >= 0; --q) buf[len++] = v.label_bytes[extra][q];
  uint32_t w = node;
  for (uint32_t k = 0; k < n_back && w != kNone && len <= 32; ++k) {
    const Node nd = s.nodes[w];
    if (nd.chr == kRootChar) break;
    for (int q = (int)v.label_len[nd.chr] - 1; q >= 0; --q) buf[len++] = v.label_bytes[nd.chr][q];
    w = nd.parent;
  }
  for (int a = 0, b = len - 1; a < b; ++a, --b) { const uint8_t t = buf[a]; buf[a] = buf[b]; buf[b] = t; }
  return sttscorer::vocab_index(v, buf, (uint32_t)len);
}

// Scorer::make_ngram + get_log_cond_prob in UTF-8 mode, literally: the <= order code points that end at the prefix
// (`node` followed by the optional candidate label `extra`), scored from BeginSentence when the window holds fewer
// than `order` units, from the null context otherwise; OOV_SCORE if a unit is not in the vocabulary.  Also returns the
// hot-word boost of the window (ctc_beam_search_decoder.cpp:224-236).  A unit's id is cached in Node::lm_wid of the
// node that ends it.
__device__ double utf8_window_cond(const Slot& s, const DecodeParams& p, uint32_t node, uint32_t extra, float* boost_out) {
  const sttscorer::ScorerView& v = p.scorer;
  const int order = (int)v.order;
  uint32_t ids_rev[sttscorer::kMaxOrder];
  int n = 0;
  uint32_t cur = node;
  if (extra != kNone) {
    // the unit that the candidate label completes (or leaves incomplete)
    const Node nd = s.nodes[node];
    uint32_t ord, ctx;
    utf8_child_fields(v, nd, node, extra, &ord, &ctx);
    ids_rev[n++] = utf8_unit_id(s, v, node, (ord & 0xffu) - 1u, extra);
    cur = ctx;
  }
  while (n < order && cur != kNone) {
    const Node nd = s.nodes[cur];
    if (nd.chr == kRootChar) break;
    uint32_t id = nd.lm_wid;
    if (id == kNone || !utf8_completes(nd.ord)) {
      id = utf8_unit_id(s, v, cur, nd.ord & 0xffu, kNone);
      if (utf8_completes(nd.ord)) s.nodes[cur].lm_wid = id;   // every writer writes the same value
    }
    ids_rev[n++] = id;
    cur = nd.last_space;
  }
  uint32_t ids[sttscorer::kMaxOrder];
  float boost = 0.0f;
  for (int i = 0; i < n; ++i) {
    ids[i] = ids_rev[n - 1 - i];
    if (p.n_hot > 0 && ids[i] != 0) {
      bool hit = false;
      float b = 0.0f;
      for (int h = 0; h < p.n_hot; ++h)
        if (p.hot_id[h] == ids[i]) { b = p.hot_boost[h]; hit = true; }
      if (hit) boost += b;
    }
  }
  if (boost_out) *boost_out = boost;
  if (n == 0) return 0.0;   // get_log_cond_prob of no words (scorer.cpp:325,343)
  return sttscorer::log_cond_prob_ids(v, ids, n, n < order);
}

// The same value from the CARRIED KenLM state (the argument of lm_eval_node, with code points for words): the unit `wid`
// scored from the state stored at its context node -- the node where the previous code point ended, whose lm_* rows were
// filled when it was created -- or from BeginSentence at the root.  One trie descent instead of `order`.  Returns false
// when the context carries no state (then the literal window above is evaluated).
__device__ bool utf8_cond_carried(const Slot& s, const DecodeParams& p, uint32_t ctx, uint32_t wid, double* cond_out,
                                  sttscorer::LmState* out, uint32_t* meta_out) {
  const sttscorer::ScorerView& v = p.scorer;
  const int order = (int)v.order;
  sttscorer::LmState c;
  uint32_t ctx_oov = 255, ctx_words = 0;
  if (ctx == kNone || s.nodes[ctx].chr == kRootChar) {
    sttscorer::begin_sentence_state(v, c);
  } else {
    const uint32_t meta = s.lm_meta[ctx];
    if (meta == kNone) return false;
    c.length = (uint8_t)(meta & 0xffu);
    ctx_oov = (meta >> 8) & 0xffu;
    ctx_words = meta >> 16;
    for (int i = 0; i < (int)c.length && i < kStateWords; ++i) {
      c.words[i] = s.lm_sw[(size_t)ctx * kStateWords + i];
      c.backoff[i] = s.lm_sb[(size_t)ctx * kStateWords + i];
    }
  }
  out->length = 0;
  uint32_t oov_dist;
  double cond;
  if (wid == 0) {
    sttscorer::null_context_state(*out);
    oov_dist = 0;
    cond = -1000.0;   // OOV_SCORE (scorer.cpp:328-331)
  } else {
    const float p10 = sttscorer::full_score(v, c, wid, *out);
    oov_dist = ctx_oov >= 254 ? 255u : ctx_oov + 1;
    cond = (oov_dist <= (uint32_t)(order - 1)) ? -1000.0 : (double)p10 / (double)0.4342944819f;
  }
  const uint32_t nwords = ctx_words >= 0xfffeu ? 0xffffu : ctx_words + 1;
  *meta_out = (uint32_t)out->length | (oov_dist << 8) | (nwords << 16);
  *cond_out = cond;
  return true;
}

// LM term (before alpha) of an EXISTING node whose last byte completes a code point; cached in Slot::lm_cond.
__device__ double utf8_node_cond(const Slot& s, const DecodeParams& p, uint32_t node) {
  if (p.n_hot == 0) {
    const unsigned long long bits = reinterpret_cast<const unsigned long long*>(s.lm_cond)[node];
    if (bits != kLmUnset) return __longlong_as_double((long long)bits);
  }
  float boost = 0.0f;
  const double cond = utf8_window_cond(s, p, node, kNone, &boost);
  if (p.n_hot == 0) s.lm_cond[node] = cond;
  return cond + (double)boost;
}

// ------------------------------------------------------------------------------------------------ init
// DecoderState::init (:22-61): root prefix with score = log_prob_b_prev = 0.
__global__ void decoder_init_kernel(Slot* slots, int n_slots, int32_t fst_start, uint32_t ht_gen) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_slots) return;
  Slot& s = slots[u];
  Node root;
  root.parent = kNone; root.chr = kRootChar; root.dict = fst_start; root.last_space = kNone; root.ord = 0;
  root.live_slot = 0; root.lm_wid = kNone; root.child_mask = 0;
  s.nodes[0] = root;
  s.lm_meta[0] = kNone;
  reinterpret_cast<unsigned long long*>(s.lm_cond)[0] = kLmUnset;
  s.ht_gen = ht_gen;
  s.ts_tree[0] = make_uint2(kNone, 0u);
  s.score[0] = 0.f;
  s.b_prev[0] = 0.f;
  s.nb_prev[0] = kNegMax;
  s.node[0] = 0;
  s.ts[0] = 0;
  for (int q = 0; q < 16; ++q) s.scalars[q] = 0;
  s.scalars[0] = 1;  // n_live
  s.scalars[2] = 1;  // arena_count
  s.scalars[3] = 1;  // ts_count
  for (int q = 0; q < 8; ++q) s.phase_cycles[q] = 0;
}

// ------------------------------------------------------------------------------------------------ step kernel
// Barrier among the NT threads of the step kernel: named barrier 1 (barrier 0 is left to __syncthreads users).
// block_scan: block-wide exclusive scan of one count per thread; returns the thread's offset and the block total.
template <int NT>
__device__ __forceinline__ void main_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
}
template <int NT, bool kTrailingBarrier = true>
__device__ __forceinline__ uint32_t block_scan(uint32_t cnt, uint32_t* warp_sums /*[NT/32 + 1]*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) warp_sums[warp] = incl;
  main_sync<NT>();
  if (warp == 0) {
    uint32_t v = (lane < NT / 32) ? warp_sums[lane] : 0, inc2 = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc2, d);
      if (lane >= d) inc2 += o;
    }
    if (lane < NT / 32) warp_sums[lane] = inc2 - v;
    if (lane == 31) warp_sums[NT / 32] = inc2;
  }
  main_sync<NT>();
  const uint32_t off = warp_sums[warp] + (incl - cnt);
  total = warp_sums[NT / 32];
  if (kTrailingBarrier) main_sync<NT>();  // callers that scan again before another barrier need warp_sums intact
  return off;
}

// Single-barrier variant: every warp redundantly scans the NT/32 warp totals, so nobody waits for warp 0.  `warp_sums`
// must not be rewritten before another barrier has passed (callers that scan in a loop alternate two buffers).
template <int NT>
__device__ __forceinline__ uint32_t block_scan1(uint32_t cnt, uint32_t* warp_sums /*[NT/32]*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  if (lane == 31) warp_sums[warp] = incl;
  main_sync<NT>();
  const uint32_t v = (lane < NT / 32) ? warp_sums[lane] : 0u;
  uint32_t inc2 = v;
#pragma unroll
  for (int d = 1; d < NT / 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc2, d);
    if (lane >= d) inc2 += o;
  }
  total = __shfl_sync(0xffffffffu, inc2, NT / 32 - 1);
  return __shfl_sync(0xffffffffu, inc2 - v, warp) + (incl - cnt);
}

// Shared-memory image of one live list.
template <int WC>
struct LiveList {
  float score[WC], b[WC], nb[WC];
  uint32_t node[WC], ts[WC], pnode[WC], lsp[WC], pos[WC], mask[WC];
  int32_t dict[WC];
  uint8_t chr[WC];
};

template <int WC, int NC>
struct StepSmem {
  LiveList<WC> live[2];
  uint32_t child[WC];     // phase 2: labels that already have a live child; phase 4+: labels to create
  uint32_t plive[WC];     // live index of the parent, or kNone
  uint32_t tsprev[WC];    // timestep-tree parent chosen for an updated live prefix (kNone = keep)
  uint32_t lmq[WC];       // LM work list
  uint32_t lmwid[WC];
  float lmterm[WC];
  // word ordinal + created-children mask of each live prefix, double buffered: [ord0 | ord1 | cmask0 | cmask1].  The
  // wide-beam instantiation (WC > 512, one CTA per SM) has no room left and keeps them in Slot::aux (global memory).
  uint32_t aux[WC <= 512 ? 4 * WC : 1];
  // per-step candidates: a 64-bit key each.  Only the live prefixes' entries carry payload (their updated blank / non-blank
  // log-probs, p0 / p1); a NEW candidate's parent and label are part of its key (new_cand_idx), so 8 bytes per candidate
  // buy room for 5632 candidates where 16 bytes bought 3072, and steps that spill to global memory become rare.
  unsigned long long key[NC > 0 ? NC : 1];
  uint32_t p0[NC > 0 ? WC : 1], p1[NC > 0 ? WC : 1];
};

// kInstr compiles in the per-phase clocks and LM counters (bench statistics).
// MINB = CTAs per SM the register budget is sized for: 2 (64 registers) when a batch needs two CTAs per SM, 1 (128 registers:
// no spills, no re-materialised thread ids -- 6.35 ms instead of 8.0 for 128 utterances) when it fits one CTA per SM.
// Measured alternatives (round 2, B200): 256 threads with two prefixes each (128 registers) 9.3 ms vs 8.8 for 256 utterances;
// 384 threads neutral (round 1).
template <int NT, int WC, int NC, bool kInstr, int MINB = 2>
__global__ void __launch_bounds__(NT, (WC <= 512 ? MINB : 1))
decoder_step_kernel(Slot* slots, const StepInput* inputs, const DecodeParams p) {
  static_assert(NT == 512 || NT == 256, "phase 6: 128 (round, warp) counts, four per lane of the scanning warps");
  constexpr int kCommitRounds = kCommitSpan / NT;   // candidate rounds (of NT) compacted per scan in phase 6
  constexpr int kSelBins = 4 * NT;                  // phase 5: bins of the one-pass score histogram (one uint4 per thread)
  __shared__ Slot s_slot;   // the slot's pointers and capacities are read all over the step loop
  const int tid = threadIdx.x;
  if (tid < (int)(sizeof(Slot) / 4)) reinterpret_cast<uint32_t*>(&s_slot)[tid] = reinterpret_cast<const uint32_t*>(&slots[blockIdx.x])[tid];
  __syncthreads();
  const Slot& s = s_slot;
  const StepInput in = inputs[blockIdx.x];
  const int C = p.n_classes;
  const int blank = C - 1;
  const int W = p.beam;
  const sttscorer::ScorerView& sv = p.scorer;
  const uint32_t all_labels = (blank >= 32) ? 0xffffffffu : ((1u << blank) - 1u);

  __shared__ float s_logp2[2][kMaxClasses];
  __shared__ double s_logblank[2];
  __shared__ uint32_t s_gate[2];
  __shared__ __align__(16) uint32_t s_hist[256];
  __shared__ uint32_t s_rs[2];
  __shared__ uint32_t s_bm[32];  // 1024-bit filter over the ids of the nodes revived by the current commit
  __shared__ double s_nextp[kMaxClasses];   // next row of probabilities, fetched with cp.async (f32 rows use the first half)   // "a node was revived" flags of the last two commits
  __shared__ __align__(16) uint32_t s_cnt[kCommitRounds * (NT / 32)];
  __shared__ uint32_t s_warp[NT / 32 + 1];
  __shared__ uint32_t s_scan[2][NT / 32];   // block_scan1 buffers (alternated)
  __shared__ float s_red[NT / 32], s_red2[NT / 32];
  __shared__ uint32_t s_u[8];
  __shared__ unsigned long long s_ph[8];
  __shared__ uint32_t s_ist[8];   // kInstr: 0 LM cache misses of the current step; totals: 1 misses, 2 steps with a miss,
                                  // 3 / 4 cycles >> 4 of phase 4a in steps with / without a miss, 5 expanding steps
  __shared__ unsigned long long s_thresh;   // phase 5: smallest selected key of the boundary bin
  extern __shared__ __align__(16) uint8_t s_dyn[];
  StepSmem<WC, NC>& sm = *reinterpret_cast<StepSmem<WC, NC>*>(s_dyn);

  uint32_t n_live = s.scalars[0];
  uint32_t arena_count = s.scalars[2];
  uint32_t ts_count = s.scalars[3];
  uint32_t abs_t = s.scalars[4];
  uint32_t start_expanding = s.scalars[5];
  uint32_t overflow = s.scalars[6];
  uint32_t max_cand = 0, spills = 0;
  if (tid == 0) {
    s_u[6] = 0;
    s_u[7] = 0;
    for (int q = 0; q < 8; ++q) { s_ph[q] = 0; s_ist[q] = 0; }
  }
  uint32_t* const aux_base = (WC <= 512) ? sm.aux : s.aux;
  uint32_t* const aux_ord[2] = {aux_base, aux_base + WC};
  uint32_t* const aux_cm[2] = {aux_base + 2 * WC, aux_base + 3 * WC};
  // ---- load the live list left by the previous launch
  int cur = 0;
  uint32_t rescan = 1, cpar = 0;  // see phase 0 / phase 6
  bool launch_start = true;
  if (tid == 0) { s_rs[0] = 0; s_rs[1] = 0; }
  for (uint32_t i = tid; i < n_live; i += NT) {
    LiveList<WC>& L = sm.live[0];
    const uint32_t nd = s.node[i];
    const Node n = s.nodes[nd];
    L.score[i] = s.score[i];
    L.b[i] = s.b_prev[i];
    L.nb[i] = s.nb_prev[i];
    L.node[i] = nd;
    L.ts[i] = s.ts[i];
    L.pnode[i] = n.parent;
    L.lsp[i] = n.last_space;
    L.dict[i] = n.dict;
    aux_ord[0][i] = ((int)n.chr == p.space_id && p.has_scorer) ? 0u : n.ord;  // a space node starts a new word
    aux_cm[0][i] = n.child_mask;
    L.chr[i] = (uint8_t)n.chr;
    uint2 st = make_uint2(0u, all_labels);
    if (p.has_scorer) st = p.fst_state2[n.dict];
    L.pos[i] = st.x;
    L.mask[i] = st.y;
  }
  const bool use64 = in.probs64 != nullptr;
  auto prob_f = [&](int row, int c) -> float {
    return use64 ? (float)in.probs64[(size_t)row * C + c] : in.probs[(size_t)row * C + c];
  };
  auto prob_d = [&](int row, int c) -> double {
    return use64 ? in.probs64[(size_t)row * C + c] : (double)in.probs[(size_t)row * C + c];
  };
  // thread c < C waits for ITS element of the prefetched row and turns it into the class log-prob; the blank's thread
  // also prepares the gate (:125) and log(blank) for min_cutoff (:143)
  const int rcls = NT - 1 - tid;   // class handled by this thread when a row of probabilities is prepared
  auto next_row_ready = [&](int buf) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const double pd = use64 ? s_nextp[rcls] : (double)reinterpret_cast<const float*>(s_nextp)[rcls];
    const float pf = use64 ? (float)s_nextp[rcls] : reinterpret_cast<const float*>(s_nextp)[rcls];
    s_logp2[buf][rcls] = sttmath::glibc_logf(pf + kFltMin);
    if (rcls == blank) {
      s_gate[buf] = pd < 0.999 ? 1u : 0u;
      s_logblank[buf] = log(pd);
    }
  };
  for (uint32_t i = tid; i < (uint32_t)WC; i += NT) sm.child[i] = 0;
  for (int h = tid; h < 256; h += NT) s_hist[h] = 0;
  if (in.n_steps > 0) {
    // class log-probs of the first row (get_pruned_emissions :328-358 with the C-API's cutoff_prob = 1.0,
    // cutoff_top_n = 40 >= n_classes: no pruning, index order, blank last); later rows are prepared one step ahead
    if (tid < C) s_logp2[0][tid] = sttmath::glibc_logf(prob_f(0, tid) + kFltMin);
    if (tid == 0) {
      const double pb = prob_d(0, blank);
      s_gate[0] = pb < 0.999 ? 1u : 0u;
      s_logblank[0] = log(pb);
    }
  }
  main_sync<NT>();
  long long ph_t0 = kInstr ? clock64() : 0;
#define PHASE_MARK(k) do { if (kInstr && tid == 0) { const long long _t = clock64(); s_ph[k] += (unsigned long long)(_t - ph_t0); ph_t0 = _t; } } while (0)

  for (int step = 0; step < in.n_steps; ++step, ++abs_t) {
    LiveList<WC>& L = sm.live[cur];
    LiveList<WC>& Nx = sm.live[cur ^ 1];
    uint32_t* const ordL = cur ? aux_ord[1] : aux_ord[0];
    uint32_t* const ordN = cur ? aux_ord[0] : aux_ord[1];
    uint32_t* const cmL = cur ? aux_cm[1] : aux_cm[0];
    uint32_t* const cmN = cur ? aux_cm[0] : aux_cm[1];
    // ---- phase 0: gate (:125-132), the beam's minimum score, and which (parent, label) pairs already have a live
    //      child.  This row's log-probs were prepared during the previous step; the next row is fetched now and turned
    //      into log-probs at the end of this step, off the critical path.
    const int cb = step & 1;
    const float* s_logp = s_logp2[cb];
    // phase 5's score histogram lives in the NEXT live list's storage (dead until phase 6 writes it): clear it now, the
    // barrier below publishes the zeros
    uint32_t* const sel_hist = reinterpret_cast<uint32_t*>(&Nx);
    unsigned long long* const sel_blist = reinterpret_cast<unsigned long long*>(sel_hist + kSelBins);
    static_assert(kSelBins == 4 * NT, "one uint4 of histogram bins per prefix thread");
    static_assert(sizeof(LiveList<WC>) >= kSelBins * 4 + kSelBoundaryCap * 8, "select scratch must fit in a live list");
    reinterpret_cast<uint4*>(sel_hist)[tid] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) s_u[1] = 0;   // length of phase 5's boundary list
    const bool have_next = (step + 1 < in.n_steps);
    if (have_next && rcls < C) {
      // asynchronous copy straight into shared memory: a register destination would be spilled, and the spill store
      // would wait out the DRAM latency right here
      if (use64) {
        const double* src = in.probs64 + (size_t)(step + 1) * C + rcls;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_nextp[rcls])), "l"(src) : "memory");
      } else {
        const float* src = in.probs + (size_t)(step + 1) * C + rcls;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(reinterpret_cast<float*>(s_nextp) + rcls)), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (start_expanding | s_gate[cb]) {
      if (rescan) {
        // first expanding step of a launch, or a pruned node was revived by the last commit: ask the arena
        for (uint32_t j = tid; j < n_live; j += NT) {
          const uint32_t pn = L.pnode[j];
          uint32_t pi = kNone;
          if (pn != kNone) {
            if (launch_start) {
              // Node::live_slot is only written when a launch ends, so it may be stale for a node that has left the
              // beam since: believe it only if that slot really holds the parent
              pi = s.nodes[pn].live_slot;
              if (pi >= n_live || L.node[pi] != pn) pi = kNone;
            } else {
              for (uint32_t i = 0; i < n_live; ++i)   // more revivals than the announcement list holds: search
                if (L.node[i] == pn) pi = i;
            }
            if (pi != kNone) atomicOr(&sm.child[pi], 1u << L.chr[j]);
          }
          sm.plive[j] = pi;
        }
        launch_start = false;
      } else {
        // the last commit left every prefix's parent slot in plive (old slot -> new slot map, phase 6)
        for (uint32_t j = tid; j < n_live; j += NT) {
          const uint32_t pi = sm.plive[j];
          if (pi != kNone) atomicOr(&sm.child[pi], 1u << L.chr[j]);
        }
      }
    }
    {
      float m = 3.402823466e+38f;
      for (uint32_t i = tid; i < n_live; i += NT) m = fminf(m, L.score[i]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, d));
      if ((tid & 31) == 0) s_red[tid >> 5] = m;
    }
    start_expanding |= s_gate[cb];
    main_sync<NT>();
    PHASE_MARK(1);
    if (!start_expanding || overflow) {
      if (have_next && rcls < C) next_row_ready(cb ^ 1);
      main_sync<NT>();
      continue;
    }
    // min_cutoff (:134-146), computed redundantly by every thread to save a barrier
    float min_cutoff = kNegMax;
    bool full_beam = false;
    if (p.has_scorer) {
      float mm = s_red[0];
#pragma unroll
      for (int w = 1; w < NT / 32; ++w) mm = fminf(mm, s_red[w]);
      const double beta_pos = sv.beta > 0.0 ? sv.beta : 0.0;
      min_cutoff = (float)((double)mm + s_logblank[cb] - beta_pos);
      full_beam = (n_live == (uint32_t)W);
    }
    PHASE_MARK(0);

    PHASE_MARK(2);

    // ---- phase 4a: which children each live prefix creates (labels allowed by the dictionary, without a live child,
    //      above the cutoff); count and scan so that candidates land at deterministic offsets
    uint32_t n_new = 0;
    {
      uint32_t run_base = 0;
      for (uint32_t base = 0; base < n_live; base += NT) {
        const uint32_t i = base + tid;
        uint32_t allow = 0;
        if (i < n_live) {
          const float si = L.score[i];
          if (si != kNegMax) {
            // LM term of a prefix that will be extended by the space this step (new child or pulled by a live one).
            // No barrier of its own: threads without LM work go on to their label masks while these wait for the
            // arena, and everybody meets at the scan below, which phase 3 / 4b (the readers of lmterm) follow.
            if (p.has_scorer && (((sm.child[i] | L.mask[i]) >> p.space_id) & 1u) &&
                !(full_beam && s_logp[p.space_id] + si < min_cutoff)) {
              uint32_t wid, nw;
              // (float)(cond * alpha): ctc_beam_search_decoder.cpp:239
              uint32_t miss = 0;
              double cond = lm_eval_node(s, p, L.node[i], L.lsp[i], &wid, &nw, kInstr ? &miss : nullptr);
              if (p.n_hot > 0) cond += (double)hot_word_boost(s, p, L.node[i], wid);
              sm.lmterm[i] = (float)(cond * sv.alpha);
              sm.lmwid[i] = wid;
              if (kInstr) {
                atomicAdd(&s_u[6], nw);
                atomicAdd(&s_u[7], 1u);
                if (miss) atomicAdd(&s_ist[0], 1u);
              }
            }
            allow = L.mask[i] & all_labels & ~sm.child[i];
            if (full_beam) {
              uint32_t m = allow;
              while (m) {
                const int c = __ffs(m) - 1;
                m &= m - 1;
                if (s_logp[c] + si < min_cutoff) allow &= ~(1u << c);
              }
            }
          }
        }
        uint32_t total;
        const uint32_t off = run_base + block_scan1<NT>(__popc(allow), s_scan[(base / NT) & 1], total);
        if (i < n_live) {
          sm.child[i] = allow;   // the live-child masks are no longer needed (phase 3 uses plive)
          sm.lmq[i] = off;       // reuse: offset of this prefix's first child among the new candidates
        }
        run_base += total;
      }
      n_new = run_base;
    }
    if (kInstr && tid == 0) {   // the scan's barriers have passed: every LM evaluation of this step is done
      const long long t4a = clock64();
      const uint32_t m = s_ist[0];
      s_ist[0] = 0;
      s_ist[1] += m;
      s_ist[2] += m ? 1u : 0u;
      s_ist[m ? 3 : 4] += (uint32_t)((unsigned long long)(t4a - ph_t0) >> 4);
      s_ist[5] += 1u;
    }
    const uint32_t N = n_live + n_new;
    if (N > s.cand_cap) { overflow = 1; main_sync<NT>(); continue; }
    max_cand = N > max_cand ? N : max_cand;
    const bool in_smem = (NC > 0) && (N <= (uint32_t)NC);
    if (!in_smem) ++spills;
    unsigned long long* const K = in_smem ? sm.key : s.c_key;
    uint32_t* const P0 = (NC > 0) ? sm.p0 : s.c_p0;   // live entries only (j < n_live <= WC)
    uint32_t* const P1 = (NC > 0) ? sm.p1 : s.c_p1;

    // ---- phase 3: updated values of the live prefixes (blank / repeat / pulled extension), :150-256
    float cand_min = 3.402823466e+38f, cand_max = kNegMax;   // over this thread's finite candidate scores (phase 5 bins)
    for (uint32_t j = tid; j < n_live; j += NT) {
      const float sj = L.score[j];
      const uint32_t cj = (L.chr[j] == (uint8_t)kRootChar) ? kRootChar : (uint32_t)L.chr[j];
      float nb = kNegMax, bcur = kNegMax;
      uint32_t ts_prev = kNone;  // kNone = keep current timesteps
      bool has_ext = false, parent_first = false;
      float lp_ext = kNegMax;
      uint32_t ts_par = kNone;
      const uint32_t pi = sm.plive[j];
      if (pi != kNone) {
        const float sp = L.score[pi];
        const float lc = s_logp[cj];
        if (sp != kNegMax && !(full_beam && lc + sp < min_cutoff)) {
          has_ext = true;
          const uint32_t cp = (L.chr[pi] == (uint8_t)kRootChar) ? kRootChar : (uint32_t)L.chr[pi];
          if (cj == cp) {
            const float bp = L.b[pi];
            lp_ext = (bp > kNegMax) ? lc + bp : kNegMax;
          } else {
            lp_ext = lc + sp;
          }
          if (p.has_scorer && (int)cj == p.space_id) {
            lp_ext += sm.lmterm[pi];
            lp_ext = (float)((double)lp_ext + sv.beta);
          }
          ts_par = L.ts[pi];
          parent_first = visits_before(sp, cp, pi, sj, cj, j);
        }
      }
      const bool alive = (sj != kNegMax);
      bool has_rep = false;
      float lp_rep = kNegMax;
      if (alive && cj != kRootChar && !(full_beam && s_logp[cj] + sj < min_cutoff)) {
        has_rep = true;
        lp_rep = s_logp[cj] + L.nb[j];
      }
      if (has_ext && parent_first) {
        if (nb < lp_ext) ts_prev = ts_par;
        nb = sttmath::log_sum_exp(nb, lp_ext);
      }
      if (has_rep) {
        if (nb < lp_rep) ts_prev = kNone;
        nb = sttmath::log_sum_exp(nb, lp_rep);
      }
      if (has_ext && !parent_first) {
        if (nb < lp_ext) ts_prev = ts_par;
        nb = sttmath::log_sum_exp(nb, lp_ext);
      }
      if (alive && !(full_beam && s_logp[blank] + sj < min_cutoff)) {
        const float lp_b = s_logp[blank] + sj;
        if (nb < lp_b) ts_prev = kNone;
        bcur = lp_b;  // log_sum_exp(-FLT_MAX, lp_b)
      }
      const float ns = sttmath::log_sum_exp(bcur, nb);
      if (ns > kNegMax) { cand_min = fminf(cand_min, ns); cand_max = fmaxf(cand_max, ns); }
      K[j] = make_key(ns, cj, j);
      P0[j] = __float_as_uint(bcur);
      P1[j] = __float_as_uint(nb);
      sm.tsprev[j] = ts_prev;
    }
    PHASE_MARK(3);

    // ---- phase 4b: write the new children.  A prefix with few children is expanded by its own thread; one with many
    //      (a word start allows every letter) would keep the other 31 lanes of its warp waiting, so those are expanded
    //      by the whole warp, one lane per label.
    for (uint32_t base = 0; base < n_live; base += NT) {
      const uint32_t i = base + tid;
      const int lane = tid & 31;
      uint32_t allow = (i < n_live) ? sm.child[i] : 0u;
      float si = 0.f, bp = 0.f, lmt = 0.f;
      uint32_t cp = kRootChar, e0 = 0;
      if (allow) {
        e0 = n_live + sm.lmq[i];
        si = L.score[i];
        bp = L.b[i];
        cp = (L.chr[i] == (uint8_t)kRootChar) ? kRootChar : (uint32_t)L.chr[i];
        if (p.has_scorer && ((allow >> p.space_id) & 1u)) lmt = sm.lmterm[i];
      }
      auto emit = [&](uint32_t ii, uint32_t e, int c, float s_i, float b_p, uint32_t c_p, float lm_t) {
        float lp;
        if ((uint32_t)c == c_p) lp = (b_p > kNegMax) ? s_logp[c] + b_p : kNegMax;
        else lp = s_logp[c] + s_i;
        if (p.has_scorer && c == p.space_id) {
          lp += lm_t;
          lp = (float)((double)lp + sv.beta);
        }
        if (lp > kNegMax) { cand_min = fminf(cand_min, lp); cand_max = fmaxf(cand_max, lp); }
        K[e] = make_key(lp, (uint32_t)c, new_cand_idx(ii, (uint32_t)c));   // the child's dictionary state is looked up in phase 6, for survivors only
      };
      const bool heavy = __popc(allow) > 6;
      uint32_t hv = __ballot_sync(0xffffffffu, heavy);
      while (hv) {
        const int src = __ffs(hv) - 1;
        hv &= hv - 1;
        const uint32_t a = __shfl_sync(0xffffffffu, allow, src);
        const uint32_t eb = __shfl_sync(0xffffffffu, e0, src);
        const float s_i = __shfl_sync(0xffffffffu, si, src), b_p = __shfl_sync(0xffffffffu, bp, src);
        const float lm_t = __shfl_sync(0xffffffffu, lmt, src);
        const uint32_t c_p = __shfl_sync(0xffffffffu, cp, src);
        if ((a >> lane) & 1u) emit(base + (tid & ~31) + src, eb + __popc(a & ((1u << lane) - 1u)), lane, s_i, b_p, c_p, lm_t);
      }
      if (!heavy) {
        uint32_t e = e0;
        while (allow) {
          const int c = __ffs(allow) - 1;
          allow &= allow - 1;
          emit(i, e, c, si, bp, cp, lmt);
          ++e;
        }
      }
    }
    {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        cand_min = fminf(cand_min, __shfl_xor_sync(0xffffffffu, cand_min, d));
        cand_max = fmaxf(cand_max, __shfl_xor_sync(0xffffffffu, cand_max, d));
      }
      if ((tid & 31) == 0) { s_red[tid >> 5] = cand_min; s_red2[tid >> 5] = cand_max; }
    }
    main_sync<NT>();
    for (uint32_t i = tid; i < (uint32_t)WC; i += NT) sm.child[i] = 0;  // ready for the next step's phase 0
    if (have_next && rcls < C) next_row_ready(cb ^ 1);
    PHASE_MARK(4);

    // ---- phase 5: exact top-W selection on the 64-bit key (:263-274 nth_element + prefix_compare).
    //      One pass: a kSelBins-bin histogram over [min, max] of the finite candidate scores (the bin is a monotone
    //      function of the score, so everything in a higher bin beats everything in a lower one), a suffix scan finds
    //      the bin the W-th best key falls in, and that bin's few members are ranked pairwise on the full key.  Five
    //      barriers and no contended atomics, instead of two barriers per 8-bit radix pass with most keys landing in
    //      one or two bins of the leading passes.  The radix passes remain as the fallback for a crowded boundary bin
    //      (hundreds of equal scores, e.g. -FLT_MAX zombies at the boundary).
    unsigned long long sel_prefix = 0, sel_mask = 0;
    int sel_mode = 0;              // 0 keep all / radix predicate, 1 bin >= sel_bin, 2 key >= sel_key
    uint32_t sel_bin = 0;
    unsigned long long sel_key = 0;
    float sel_lo = 0.f, sel_scale = 0.f;
    auto bin_of = [&](unsigned long long key) -> uint32_t {
      const float sc = unsortable((uint32_t)(key >> 32));
      if (!(sc > kNegMax)) return 0u;
      const float x = (sc - sel_lo) * sel_scale;
      int b = (int)x;
      b = b < 0 ? 0 : (b > kSelBins - 1 ? kSelBins - 1 : b);
      return (uint32_t)b;
    };
    bool need_radix = (N > (uint32_t)W);
    if (need_radix && (p.flags & kFlagHistSelect)) {
      float lo = s_red[0], hi = s_red2[0];
#pragma unroll
      for (int w = 1; w < NT / 32; ++w) { lo = fminf(lo, s_red[w]); hi = fmaxf(hi, s_red2[w]); }
      sel_lo = lo;
      sel_scale = (hi > lo) ? (float)(kSelBins - 1) / (hi - lo) : 0.f;
      for (uint32_t e = tid; e < N; e += NT) atomicAdd(&sel_hist[bin_of(K[e])], 1u);
      main_sync<NT>();
      // thread t owns bins [4t, 4t+4); higher bins = better scores, so count from the top
      const uint4 hv = reinterpret_cast<const uint4*>(sel_hist)[tid];
      const uint32_t mine = hv.x + hv.y + hv.z + hv.w;
      uint32_t suffix = mine;   // inclusive over lanes >= this one
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_down_sync(0xffffffffu, suffix, d);
        if (lane + d < 32) suffix += o;
      }
      if (lane == 0) s_warp[warp] = suffix;
      main_sync<NT>();
      // every warp locates the boundary bin by itself: which warp's 128 bins the W-th best key falls in (from the 16 warp
      // totals), then a second look at that warp's bins -- no broadcast through shared memory, no further barrier
      uint32_t k_rem, cnt;
      {
        const uint32_t wv = (lane < NT / 32) ? s_warp[lane] : 0u;
        uint32_t wsuf = wv;   // inclusive over warps >= lane
#pragma unroll
        for (int d = 1; d < NT / 32; d <<= 1) {
          const uint32_t o = __shfl_down_sync(0xffffffffu, wsuf, d);
          if (lane + d < NT / 32) wsuf += o;
        }
        const uint32_t hit = __ballot_sync(0xffffffffu, lane < NT / 32 && wsuf - wv < (uint32_t)W && wsuf >= (uint32_t)W);
        const int bw = __ffs(hit) - 1;                               // exactly one warp qualifies (N > W)
        const uint32_t above_bw = __shfl_sync(0xffffffffu, wsuf - wv, bw);
        const uint4 h2 = reinterpret_cast<const uint4*>(sel_hist)[bw * 32 + lane];
        const uint32_t m2 = h2.x + h2.y + h2.z + h2.w;
        uint32_t suf2 = m2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_down_sync(0xffffffffu, suf2, d);
          if (lane + d < 32) suf2 += o;
        }
        const uint32_t incl2 = above_bw + suf2, above2 = incl2 - m2;
        const uint32_t hit2 = __ballot_sync(0xffffffffu, above2 < (uint32_t)W && incl2 >= (uint32_t)W);
        const int bl = __ffs(hit2) - 1;
        uint32_t fb = 0, fk = 0, fc = 0;
        {
          const uint32_t hq[4] = {h2.x, h2.y, h2.z, h2.w};
          uint32_t cum = above2;
          bool found = false;
#pragma unroll
          for (int q = 3; q >= 0; --q) {
            if (!found && cum + hq[q] >= (uint32_t)W) {
              fb = (uint32_t)((bw * 32 + lane) * 4 + q);
              fk = (uint32_t)W - cum;   // still needed from this bin
              fc = hq[q];               // elements in this bin
              found = true;
            }
            cum += hq[q];
          }
        }
        sel_bin = __shfl_sync(0xffffffffu, fb, bl);
        k_rem = __shfl_sync(0xffffffffu, fk, bl);
        cnt = __shfl_sync(0xffffffffu, fc, bl);
      }
      if (cnt == k_rem) {
        sel_mode = 1;            // the whole boundary bin is selected
        need_radix = false;
      } else if (cnt <= (uint32_t)kSelBoundaryCap) {
        for (uint32_t e = tid; e < N; e += NT) {
          const unsigned long long key = K[e];
          if (bin_of(key) == sel_bin) sel_blist[atomicAdd(&s_u[1], 1u)] = key;
        }
        main_sync<NT>();
        if ((uint32_t)tid < cnt) {
          const unsigned long long mk = sel_blist[tid];
          uint32_t rank = 0;
          for (uint32_t j = 0; j < cnt; ++j) rank += (sel_blist[j] > mk) ? 1u : 0u;
          if (rank == k_rem - 1) s_thresh = mk;   // keys are unique: exactly one thread
        }
        main_sync<NT>();
        sel_key = s_thresh;
        sel_mode = 2;
        need_radix = false;
      }
      // else: crowded boundary bin -> radix passes below (every thread takes the same branch: cnt is shared)
    }
    if (need_radix) {
      uint32_t k_rem = (uint32_t)W;
      for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (uint32_t e = tid; e < N; e += NT) {
          const unsigned long long key = K[e];
          if ((key & sel_mask) == sel_prefix) atomicAdd(&s_hist[(uint32_t)(key >> shift) & 255u], 1u);
        }
        main_sync<NT>();
        if (tid < 32) {
          // lane l owns bins [8l, 8l+8); find the bin where the count from the top crosses k_rem
          uint32_t hc[8];
          {
            uint4* hv = reinterpret_cast<uint4*>(&s_hist[tid * 8]);
            const uint4 h0 = hv[0], h1 = hv[1];
            hv[0] = make_uint4(0u, 0u, 0u, 0u);
            hv[1] = make_uint4(0u, 0u, 0u, 0u);
            hc[0] = h0.x; hc[1] = h0.y; hc[2] = h0.z; hc[3] = h0.w;
            hc[4] = h1.x; hc[5] = h1.y; hc[6] = h1.z; hc[7] = h1.w;
          }
          uint32_t mine = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) mine += hc[q];
          uint32_t suffix = mine;  // inclusive sum over lanes >= tid
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_down_sync(0xffffffffu, suffix, d);
            if (tid + d < 32) suffix += o;
          }
          const uint32_t above = suffix - mine;  // elements in higher lanes
          if (above < k_rem && suffix >= k_rem) {
            uint32_t cum = above;
#pragma unroll
            for (int q = 7; q >= 0; --q) {
              const uint32_t hcount = hc[q];
              if (cum + hcount >= k_rem) {
                s_u[2] = (uint32_t)(tid * 8 + q);
                s_u[3] = k_rem - cum;   // still needed from this bin
                s_u[4] = hcount;        // elements in this bin
                break;
              }
              cum += hcount;
            }
          }
        }
        main_sync<NT>();
        sel_prefix |= (unsigned long long)s_u[2] << shift;
        sel_mask |= (unsigned long long)255u << shift;
        k_rem = s_u[3];
        if (s_u[4] == k_rem) break;  // every element with this prefix is selected
      }
    }
    PHASE_MARK(5);

    // ---- phase 6: order-preserving compaction + commit (iterate_to_vec :159-190, remove :192-209)
    if (tid == 0) s_u[5] = arena_count;
    if (tid < 32) s_bm[tid] = 0;
    // (i) shared-memory-only compaction, in candidate order (live entries precede new ones): survivors get their slot
    //     in the next live list; a new survivor parks (parent index, label) there.  Up to kCommitRounds rounds of NT
    //     candidates share ONE scan: per-round warp ballots, a 128-entry scan of the warp counts by warp 0.
    uint32_t out_base = 0;
    for (uint32_t g0 = 0; g0 < N; g0 += kCommitRounds * NT) {
      uint32_t bal[kCommitRounds];
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int q = 0; q < kCommitRounds; ++q) {
        const uint32_t e = g0 + (uint32_t)q * NT + tid;
        unsigned long long key = 0;
        bool keep = false;
        if (e < N) {
          key = K[e];
          keep = (N <= (uint32_t)W) || (sel_mode == 2 ? key >= sel_key
                                        : sel_mode == 1 ? bin_of(key) >= sel_bin
                                                        : (key & sel_mask) >= sel_prefix);
        }
        bal[q] = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_cnt[q * (NT / 32) + warp] = __popc(bal[q]);
      }
      main_sync<NT>();
      uint32_t round_off[kCommitRounds], round_total;
      {
        // entry (q, w) sits at index q * (NT / 32) + w; lane l holds entries 4l .. 4l+3
        constexpr int PER = kCommitRounds * (NT / 32) / 32;  // entries per lane
        static_assert(PER == 4, "lane <-> (round, warp) mapping below");
        const uint4 v4 = reinterpret_cast<const uint4*>(s_cnt)[lane];
        const uint32_t sum = v4.x + v4.y + v4.z + v4.w;
        uint32_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        const int sub = warp & 3;
        const uint32_t pre = (incl - sum) + (sub > 0 ? v4.x : 0u) + (sub > 1 ? v4.y : 0u) + (sub > 2 ? v4.z : 0u);
#pragma unroll
        for (int q = 0; q < kCommitRounds; ++q) round_off[q] = __shfl_sync(0xffffffffu, pre, (q * (NT / 32) + warp) >> 2);
        round_total = __shfl_sync(0xffffffffu, incl, 31);
      }
#pragma unroll
      for (int q = 0; q < kCommitRounds; ++q) {
        const uint32_t e = g0 + (uint32_t)q * NT + tid;
        if (e >= N) break;
        const bool keep = (bal[q] >> lane) & 1u;
        const uint32_t pos = out_base + round_off[q] + __popc(bal[q] & ((1u << lane) - 1u));
        const unsigned long long key = K[e];
        if (e < n_live) {
          const uint32_t nd = L.node[e];
          if (keep) {
            Nx.score[pos] = unsortable((uint32_t)(key >> 32));
            Nx.b[pos] = __uint_as_float(P0[e]);
            Nx.nb[pos] = __uint_as_float(P1[e]);
            Nx.node[pos] = nd;
            Nx.pnode[pos] = L.pnode[e];
            Nx.lsp[pos] = L.lsp[e];
            Nx.pos[pos] = L.pos[e];
            Nx.mask[pos] = L.mask[e];
            ordN[pos] = ordL[e];
            cmN[pos] = cmL[e];
            Nx.dict[pos] = L.dict[e];
            Nx.chr[pos] = L.chr[e];
            const uint32_t tp = sm.tsprev[e];
            if (tp != kNone) {
              const uint32_t id = ts_count + pos;
              if (id < s.ts_cap) s.ts_tree[id] = make_uint2(tp, abs_t);
              Nx.ts[pos] = id;
            } else {
              Nx.ts[pos] = L.ts[e];
            }
            sm.lmq[pos] = 0x80000000u | e;  // not a new node: remember the old slot
            sm.tsprev[e] = pos;             // tsprev now maps old live slot -> new live slot (kNone = pruned)
          } else {
            sm.tsprev[e] = kNone;
          }
        } else if (keep) {
          const float lp = unsortable((uint32_t)(key >> 32));
          Nx.score[pos] = lp;
          Nx.b[pos] = kNegMax;
          Nx.nb[pos] = lp;
          const uint32_t kc = 255u - ((uint32_t)(key >> 24) & 0xffu);
          const uint32_t pk = ((0xffffffu - ((uint32_t)key & 0xffffffu) - 0x1000u) >> 5) | (kc << 16);
          sm.lmq[pos] = pk;  // parent live index | label << 16
          // the (parent,label) hash slot is a DRAM miss: start it now, (ii) touches it after the barrier
          const unsigned long long hw = ht_pack(s.ht_gen, L.node[pk & 0xffffu], (pk >> 16) & 0xffu, 0u);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(s.ht + (ht_hash(hw >> 24) & s.ht_mask)));
        }
      }
      out_base += round_total;
      main_sync<NT>();
    }
    PHASE_MARK(7);
    // (ii) one thread per survivor.  A NEW survivor does the global-memory work (dictionary arc / arena node / child-list
    //      push / timestep node), all latencies overlapping; every survivor also works out where its parent sits in
    //      the next live list, which saves phase 0 a round trip to the arena.
    uint32_t my_plive[(WC + NT - 1) / NT];
    uint2* const rev = reinterpret_cast<uint2*>(sm.lmterm);   // lmterm is dead until the next LM phase
    constexpr int kRevCap = WC / 2;
#pragma unroll
    for (int r = 0; r < (WC + NT - 1) / NT; ++r) {
      const uint32_t pos = (uint32_t)r * NT + tid;
      my_plive[r] = kNone;
      if (pos >= out_base) continue;
      const uint32_t pk = sm.lmq[pos];
      if (pk & 0x80000000u) {
        const uint32_t po = sm.plive[pk & 0xffffu];
        if (po != kNone) my_plive[r] = sm.tsprev[po];
        continue;
      }
      const uint32_t pi = pk & 0xffffu, c = (pk >> 16) & 0xffu;
      const bool is_space = ((int)c == p.space_id);
      const uint32_t pnode = L.node[pi];
      const uint32_t np = sm.tsprev[pi];   // the parent's slot in the next live list, if it survived
      my_plive[r] = np;
      int4 arc = make_int4(0, 0, (int)all_labels, 0);
      const uint32_t ai = p.has_scorer ? (L.pos[pi] + __popc(L.mask[pi] & ((1u << c) - 1u))) : 0u;
      if (p.has_scorer) arc = __ldg(&p.fst_arc4[2 * ai]);
      uint32_t id = kNone, own_mask = 0;
      if ((cmL[pi] >> c) & 1u) {  // this child existed before (rare): find it, it is revived under its old identity
        id = ht_find_existing(s, pnode, c);
        own_mask = *reinterpret_cast<volatile uint32_t*>(&s.nodes[id].child_mask);
        const uint32_t ri = atomicAdd(&s_rs[cpar], 1u);   // tell the live children of this node where it sits now
        if (ri < (uint32_t)kRevCap) rev[ri] = make_uint2(id, pos);
        atomicOr(&s_bm[(id >> 5) & 31u], 1u << (id & 31u));
      }
      const uint32_t cord = (is_space || !p.fst_space_skip) ? 0u : ordL[pi] + (uint32_t)arc.w;
      const float lp = Nx.score[pos];
      if (id == kNone) {
        id = atomicAdd(&s_u[5], 1u);  // fresh arena node
        if (id < s.arena_cap) {
          Node n;
          n.parent = pnode; n.chr = c; n.dict = arc.x; n.last_space = is_space ? id : L.lsp[pi];
          n.word_id = is_space ? (p.has_scorer ? sm.lmwid[pi] : 0u) : cord;
          n.live_slot = pos; n.lm_wid = kNone; n.child_mask = 0;
          uint32_t meta_init = kNone;  // LM cache of this node: not computed
          if (p.has_scorer) {
            if (is_space) {
              // the KenLM state after the word that this space terminates = the context of the next word: carried here
              // so that the next word's first evaluation finds it without walking back (lm_eval_node)
              const uint32_t pc = L.chr[pi];
              if (pc == (uint32_t)(uint8_t)kRootChar || (int)pc == p.space_id) {
                meta_init = 0u | (0u << 8) | (1u << 16);  // empty previous word: null context, an OOV inside the window
              } else {
                meta_init = s.lm_meta[pnode];
                const uint32_t len = meta_init == kNone ? 0u : (meta_init & 0xffu);
                for (uint32_t q = 0; q < len && q < (uint32_t)kStateWords; ++q) {
                  s.lm_sw[(size_t)id * kStateWords + q] = s.lm_sw[(size_t)pnode * kStateWords + q];
                  s.lm_sb[(size_t)id * kStateWords + q] = s.lm_sb[(size_t)pnode * kStateWords + q];
                }
              }
            } else if (p.fst_space_skip) {
              const uint32_t css = (uint32_t)__ldg(&p.fst_arc4[2 * ai + 1]).x;  // skip of the child's own space arc
              if (css != kNone) n.lm_wid = __ldg(p.ord2wid + cord + css);       // the word that would end here
            }
          }
          s.nodes[id] = n;
          s.lm_meta[id] = meta_init;
          reinterpret_cast<unsigned long long*>(s.lm_cond)[id] = kLmUnset;
          ht_insert(s, pnode, c, id);
          atomicOr(&s.nodes[pnode].child_mask, 1u << c);
          if (np != kNone) atomicOr(&cmN[np], 1u << c);
        }
      }
      Nx.node[pos] = id;
      Nx.pnode[pos] = pnode;
      Nx.lsp[pos] = is_space ? id : L.lsp[pi];
      Nx.chr[pos] = (uint8_t)c;
      Nx.dict[pos] = arc.x;
      Nx.pos[pos] = (uint32_t)arc.y;
      Nx.mask[pos] = (uint32_t)arc.z;
      ordN[pos] = cord;    // 0 for a space node: the next word starts here
      cmN[pos] = own_mask;
      if (lp > kNegMax) {  // "prefix_new->log_prob_nb_cur < log_p" (:246-251)
        const uint32_t tid2 = ts_count + pos;
        if (tid2 < s.ts_cap) s.ts_tree[tid2] = make_uint2(L.ts[pi], abs_t);
        Nx.ts[pos] = tid2;
      } else {
        Nx.ts[pos] = kNone;
      }
    }
    main_sync<NT>();
    {
      // prefixes whose parent was not live may have just got it back: revived nodes announce their new slot
      const uint32_t n_rev = s_rs[cpar];
      rescan = n_rev > (uint32_t)kRevCap ? 1u : 0u;
#pragma unroll
      for (int r = 0; r < (WC + NT - 1) / NT; ++r) {
        const uint32_t pos = (uint32_t)r * NT + tid;
        if (pos >= out_base) continue;
        uint32_t pl = my_plive[r];
        if (pl == kNone && n_rev != 0 && !rescan && (sm.lmq[pos] & 0x80000000u)) {
          const uint32_t pn = Nx.pnode[pos];
          // most orphans' parents were not revived: a 1024-bit filter spares them the walk over the announcements
          if (pn != kNone && ((s_bm[(pn >> 5) & 31u] >> (pn & 31u)) & 1u)) {
            for (uint32_t k = 0; k < n_rev; ++k) {
              const uint2 rv = rev[k];
              if (rv.x == pn) { pl = rv.y; break; }
            }
          }
        }
        sm.plive[pos] = pl;
      }
      cpar ^= 1;
      if (tid == 0) s_rs[cpar] = 0;
    }
    const uint32_t n_surv = out_base;
    arena_count = s_u[5];
    ts_count += n_surv;
    if (arena_count > s.arena_cap || ts_count > s.ts_cap) overflow = 1;
    n_live = n_surv;
    cur ^= 1;
    PHASE_MARK(6);
  }
#undef PHASE_MARK

  // ---- store the live list for the next launch / finalize
  main_sync<NT>();
  {
    const LiveList<WC>& L = sm.live[cur];
    for (uint32_t i = tid; i < n_live; i += NT) {
      s.score[i] = L.score[i];
      s.b_prev[i] = L.b[i];
      s.nb_prev[i] = L.nb[i];
      s.node[i] = L.node[i];
      s.ts[i] = L.ts[i];
      s.nodes[L.node[i]].live_slot = i;   // read (and validated) by the next launch's first expanding step
    }
  }
  if (tid == 0) {
    s.scalars[0] = n_live;
    s.scalars[2] = arena_count;
    s.scalars[3] = ts_count;
    s.scalars[4] = abs_t;
    s.scalars[5] = start_expanding;
    s.scalars[6] = overflow;
    if (kInstr) {
      s.scalars[7] += s_u[6];   // words scored by the LM (reference-equivalent window sizes)
      s.scalars[8] += s_u[7];   // LM calls
    }
    if (max_cand > s.scalars[9]) s.scalars[9] = max_cand;
    s.scalars[10] += spills;
    if (kInstr) {
      for (int q = 0; q < 8; ++q) s.phase_cycles[q] += s_ph[q];
      for (int q = 1; q <= 5; ++q) s.scalars[10 + q] += s_ist[q];
    }
  }
}

// ------------------------------------------------------------------------------------------------ finalize
// DecoderState::decode(num_results) (:278-326) + ModelState::decode_metadata's token/timestep extraction.
struct FinalOut {
  int max_results, max_tokens;
  int* n_results;        // [2]: number of results, overflow flag
  double* confidence;    // [max_results]
  int* n_tokens;         // [max_results]
  uint32_t* tokens;      // [max_results, max_tokens]
  uint32_t* timesteps;   // [max_results, max_tokens]
};

template <int NT>
__global__ void __launch_bounds__(NT) decoder_finalize_kernel(Slot* slots, const FinalOut* outs, const DecodeParams p,
                                                               int num_results) {
  Slot& s = slots[blockIdx.x];
  const FinalOut o = outs[blockIdx.x];
  const int tid = threadIdx.x;
  const uint32_t n_live = s.scalars[0];
  __shared__ unsigned long long s_best[NT / 32];
  float* f_score = reinterpret_cast<float*>(s.c_p0);  // scratch (the live list itself is left untouched:
  unsigned long long* f_key = s.c_key;                //          IntermediateDecode is const)

  for (uint32_t i = tid; i < n_live; i += NT) {
    const uint32_t nd = s.node[i];
    float sc = s.score[i];
    const uint32_t c = s.nodes[nd].chr;
    if (p.has_scorer && p.scorer.is_utf8) {
      // UTF-8 mode (:291-300): prefix_boundary is the prefix itself, so the root is scored too (an empty n-gram: 0), and so
      // is every prefix whose last byte leaves a code point unfinished (its window ends with that partial unit)
      if (i < (uint32_t)p.beam && (c == kRootChar || !utf8_completes(s.nodes[nd].ord))) {
        const double cond = (c == kRootChar) ? 0.0 : utf8_window_cond(s, p, nd, kNone, nullptr);
        float add = (float)(cond * p.scorer.alpha);
        add = (float)((double)add + p.scorer.beta);
        sc = sc + add;
      }
    } else if (p.has_scorer && i < (uint32_t)p.beam && c != kRootChar && (int)c != p.space_id) {
      uint32_t wid_unused, nw_unused;
      float add = (float)(lm_eval_node(s, p, nd, kStopUnknown, &wid_unused, &nw_unused) * p.scorer.alpha);
      add = (float)((double)add + p.scorer.beta);
      sc = sc + add;
    }
    f_score[i] = sc;
    f_key[i] = make_key(sc, c, i);
  }
  __syncthreads();
  int n_ret = (int)n_live < num_results ? (int)n_live : num_results;
  if (n_ret > o.max_results) n_ret = o.max_results;
  for (int r = 0; r < n_ret; ++r) {
    unsigned long long best = 0;
    for (uint32_t i = tid; i < n_live; i += NT) {
      const unsigned long long k = f_key[i];
      if (k > best) best = k;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, d);
      if (ok > best) best = ok;
    }
    if ((tid & 31) == 0) s_best[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
      unsigned long long b = s_best[0];
      for (int w = 1; w < NT / 32; ++w) if (s_best[w] > b) b = s_best[w];
      const uint32_t i = 0xffffffu - (uint32_t)(b & 0xffffffu);
      f_key[i] = 0;  // remove from further rounds
      // backtrack tokens (get_path_vec) and timesteps (get_history)
      const uint32_t nd = s.node[i];
      int len = 0;
      for (uint32_t w = nd; s.nodes[w].chr != kRootChar; w = s.nodes[w].parent) ++len;
      o.n_tokens[r] = len;
      o.confidence[r] = (double)f_score[i];
      int k = len;
      for (uint32_t w = nd; s.nodes[w].chr != kRootChar; w = s.nodes[w].parent) {
        --k;
        if (k < o.max_tokens) o.tokens[(size_t)r * o.max_tokens + k] = s.nodes[w].chr;
      }
      const uint32_t tn = s.ts[i];
      int tlen = 0;
      for (uint32_t w = tn; w != kNone && w != 0; w = s.ts_tree[w].x) ++tlen;
      k = tlen;
      for (uint32_t w = tn; w != kNone && w != 0; w = s.ts_tree[w].x) {
        --k;
        if (k < o.max_tokens) o.timesteps[(size_t)r * o.max_tokens + k] = s.ts_tree[w].y;
      }
    }
    __syncthreads();
  }
  if (tid == 0) {
    o.n_results[0] = n_ret;
    o.n_results[1] = (int)s.scalars[6];  // decoder capacity overflow flag travels with the results
  }
}

}  // namespace sttdec
