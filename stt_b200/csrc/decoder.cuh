// K7/K8/K9: CTC prefix beam search with KenLM scorer + dictionary FST on the GPU, one CTA per utterance.
//
// Restates DecoderState::{init,next,decode} (native_client/ctcdecode/ctc_beam_search_decoder.cpp:22-61,112-276,
// 278-326), PathTrie (path_trie.cpp:37-100,159-209), prefix_compare (decoder_utils.cpp:66-88),
// Scorer::{is_scoring_boundary,make_ngram,get_log_cond_prob} (scorer.cpp:271-344,369-396) with the exact
// arithmetic types of SURVEY.md appendix A (f32 log-probs through glibc-exact logf/expf, f64 only where the
// reference uses it).  The data structure is NOT a port:
//
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

// Per-utterance device state ("stream slot").  All pointers are device memory sized for (beam_cap, t_cap).
struct Slot {
  // arena of surviving prefix nodes; node 0 is the root
  uint32_t* parent;      // [arena_cap]
  uint32_t* chr;         // [arena_cap]  label, kRootChar for the root
  int32_t* dict_state;   // [arena_cap]
  uint32_t* last_space;  // [arena_cap]  nearest ancestor-or-self whose label is the space, or kNone
  uint32_t* word_id;     // [arena_cap]  for space nodes: vocab id of the word they terminate
  uint32_t* live_slot;   // [arena_cap]  index in the current live list, or kNone
  // (parent node, label) -> node id for every node ever created: a pruned node that still has live descendants must
  // be REVIVED under its old id when its prefix re-enters the beam (path_trie.cpp:45-52), so that those
  // descendants keep merging into it.
  unsigned long long* ht_key;  // [ht_mask + 1], 0 = empty, zero-initialised by the host
  uint32_t* ht_val;
  uint32_t ht_mask;
  // timestep tree (path_trie.h:17-37): ts node 0 is the root
  uint32_t* ts_parent;   // [ts_cap]
  uint32_t* ts_val;      // [ts_cap]
  // live lists, double buffered (cur = live_sel)
  float* score[2];
  float* b_prev[2];
  float* nb_prev[2];
  uint32_t* node[2];
  uint32_t* ts[2];
  // per-step candidates, capacity cand_cap = beam_cap * n_classes
  float* c_score;
  uint32_t *c_a, *c_b, *c_c, *c_d;
  uint64_t* c_key;
  // per-live scratch
  float* lm_term;        // [beam_cap] (float)((cond_prob + boost) * alpha) for "prefix i + space"
  uint32_t* lm_word;     // [beam_cap] vocab id of the word completed by that space
  // scalars (persist across launches for streaming)
  uint32_t* scalars;     // [16]: 7 LM words scored, 8 LM calls; 0 n_live, 1 live_sel, 2 arena_count, 3 ts_count, 4 abs_time_step, 5 start_expanding, 6 overflow
  uint32_t arena_cap, ts_cap, beam_cap, cand_cap;
};

struct DecodeParams {
  int n_classes;         // alphabet + 1; blank = n_classes - 1
  int beam;              // <= beam_cap
  int space_id;
  int has_scorer;
  sttscorer::ScorerView scorer;
};

struct StepInput {
  const float* probs;    // [T, n_classes] f32 softmax rows for this slot
  int n_steps;
};

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t ht_hash(unsigned long long k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  return (uint32_t)k;
}
__device__ __forceinline__ unsigned long long ht_make_key(uint32_t parent_node, uint32_t c) {
  return ((unsigned long long)(parent_node + 1u) << 8) | (unsigned long long)(c & 0xffu);
}
__device__ __forceinline__ uint32_t ht_find(const Slot& s, unsigned long long key) {
  uint32_t h = ht_hash(key) & s.ht_mask;
  for (;;) {
    const unsigned long long k = s.ht_key[h];
    if (k == key) return s.ht_val[h];
    if (k == 0ull) return 0xffffffffu;
    h = (h + 1) & s.ht_mask;
  }
}
__device__ __forceinline__ void ht_insert(const Slot& s, unsigned long long key, uint32_t val) {
  uint32_t h = ht_hash(key) & s.ht_mask;
  for (;;) {
    const unsigned long long prev = atomicCAS(&s.ht_key[h], 0ull, key);
    if (prev == 0ull || prev == key) {
      s.ht_val[h] = val;
      return;
    }
    h = (h + 1) & s.ht_mask;
  }
}
__device__ __forceinline__ uint32_t sortable(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// order in which the reference's sorted loop visits two live prefixes: true if x comes before y
__device__ __forceinline__ bool visits_before(float sx, uint32_t cx, uint32_t ix, float sy, uint32_t cy, uint32_t iy) {
  if (sx != sy) return sx > sy;
  if (cx != cy) return cx < cy;
  return ix < iy;
}

// Scorer::make_ngram + get_log_cond_prob for "prefix `node` followed by a space" (word mode).
// Returns (float)(cond_prob * alpha) exactly as ctc_beam_search_decoder.cpp:239 computes it (hot-word boost = 0)
// and the vocabulary id of the completed word.
__device__ float lm_space_term(const Slot& s, const sttscorer::ScorerView& v, uint32_t node, uint32_t* word_out,
                               uint32_t* n_words_out = nullptr) {
  uint32_t ids_rev[sttscorer::kMaxOrder];
  int n = 0;
  uint32_t first_word = 0;
  const int order = (int)v.order;
  // Scorer::make_ngram (scorer.cpp:369-396): walk back word by word.  `term` is the space node that terminated
  // the word ending at `cur` (kNone for the first, possibly unfinished, word); its vocab id is cached on it.
  uint32_t cur = node, term = kNone;
  while (n < order) {
    const uint32_t cc = s.chr[cur];
    if (cc == kRootChar) break;
    const uint32_t stop = s.last_space[cur];  // == cur when cur is itself a space (empty word)
    uint32_t id;
    if (cc == (uint32_t)v.space_label) {
      id = 0;  // get_prev_word returns an empty word: never in the vocabulary
    } else if (term == kNone) {
      uint8_t buf[kMaxWordBytes];
      int len = 0;
      bool too_long = false;
      for (uint32_t w = cur; w != stop && s.chr[w] != kRootChar; w = s.parent[w]) {
        const uint32_t c = s.chr[w];
        const int l = v.label_len[c];
        if (len + l > kMaxWordBytes) { too_long = true; break; }
        for (int q = l - 1; q >= 0; --q) buf[len++] = v.label_bytes[c][q];
      }
      for (int a = 0, b = len - 1; a < b; ++a, --b) { const uint8_t t = buf[a]; buf[a] = buf[b]; buf[b] = t; }
      id = too_long ? 0u : sttscorer::vocab_index(v, buf, (uint32_t)len);
      first_word = id;
    } else {
      id = s.word_id[term];
    }
    ids_rev[n++] = id;
    if (stop == kNone) break;
    term = stop;
    cur = s.parent[stop];
  }
  *word_out = first_word;
  if (n_words_out) *n_words_out = (uint32_t)n;
  uint32_t ids[sttscorer::kMaxOrder];
  for (int i = 0; i < n; ++i) ids[i] = ids_rev[n - 1 - i];
  const bool bos = n < order;
  const double cond = sttscorer::log_cond_prob_ids(v, ids, n, bos);
  return (float)(cond * v.alpha);
}

// Same, for DecoderState::decode's rescoring of an unfinished last word (:288-300): make_ngram(prefix) where the
// last "word" is the partial word ending at `node`.
__device__ float lm_final_term(const Slot& s, const sttscorer::ScorerView& v, uint32_t node) {
  uint32_t dummy;
  return lm_space_term(s, v, node, &dummy);
}

// ------------------------------------------------------------------------------------------------ init
// DecoderState::init (:22-61): root prefix with score = log_prob_b_prev = 0.
__global__ void decoder_init_kernel(Slot* slots, int n_slots, int32_t fst_start) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= n_slots) return;
  Slot& s = slots[u];
  s.parent[0] = kNone;
  s.chr[0] = kRootChar;
  s.dict_state[0] = fst_start;
  s.last_space[0] = kNone;
  s.word_id[0] = 0;
  s.live_slot[0] = 0;
  s.ts_parent[0] = kNone;
  s.ts_val[0] = 0;
  s.score[0][0] = 0.f;
  s.b_prev[0][0] = 0.f;
  s.nb_prev[0][0] = kNegMax;
  s.node[0][0] = 0;
  s.ts[0][0] = 0;
  s.scalars[0] = 1;  // n_live
  s.scalars[1] = 0;  // live_sel
  s.scalars[2] = 1;  // arena_count
  s.scalars[3] = 1;  // ts_count
  s.scalars[4] = 0;  // abs_time_step
  s.scalars[5] = 0;  // start_expanding
  s.scalars[6] = 0;  // overflow flag
  s.scalars[7] = 0;
  s.scalars[8] = 0;
}

// ------------------------------------------------------------------------------------------------ step kernel
// Block-wide exclusive scan of one 0/1 flag per thread; returns the thread's offset and the block total.
template <int NT>
__device__ __forceinline__ uint32_t block_scan_flag(bool flag, uint32_t* warp_sums /*[NT/32 + 1]*/, uint32_t& total) {
  const unsigned ballot = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t in_warp = __popc(ballot & ((1u << lane) - 1));
  if (lane == 0) warp_sums[warp] = __popc(ballot);
  __syncthreads();
  if (warp == 0) {
    uint32_t v = (lane < NT / 32) ? warp_sums[lane] : 0;
    uint32_t incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += o;
    }
    if (lane < NT / 32) warp_sums[lane] = incl - v;
    if (lane == 31) warp_sums[NT / 32] = incl;
  }
  __syncthreads();
  const uint32_t off = warp_sums[warp] + in_warp;
  total = warp_sums[NT / 32];
  __syncthreads();
  return off;
}

template <int NT>
__global__ void __launch_bounds__(NT) decoder_step_kernel(Slot* slots, const StepInput* inputs, const DecodeParams p) {
  Slot& s = slots[blockIdx.x];
  const StepInput in = inputs[blockIdx.x];
  const int tid = threadIdx.x;
  const int C = p.n_classes;
  const int blank = C - 1;
  const int W = p.beam;
  const sttscorer::ScorerView& sv = p.scorer;

  __shared__ float s_logp[kMaxClasses];
  __shared__ uint32_t s_hist[256];
  __shared__ uint32_t s_warp[NT / 32 + 1];
  __shared__ float s_red[NT / 32];
  __shared__ uint32_t s_u[8];
  __shared__ float s_f[4];
  extern __shared__ uint32_t s_dyn[];  // [beam_cap] child masks, then [beam_cap] LM work list
  uint32_t* s_child = s_dyn;
  uint32_t* s_lmq = s_dyn + s.beam_cap;

  uint32_t n_live = s.scalars[0];
  uint32_t sel = s.scalars[1];
  uint32_t arena_count = s.scalars[2];
  uint32_t ts_count = s.scalars[3];
  uint32_t abs_t = s.scalars[4];
  uint32_t start_expanding = s.scalars[5];
  uint32_t overflow = s.scalars[6];
  if (threadIdx.x == 0) { s_u[6] = 0; s_u[7] = 0; }

  for (int step = 0; step < in.n_steps; ++step, ++abs_t) {
    const float* prob = in.probs + (size_t)step * C;
    // ---- phase 0: gate (:125-132) and class log-probs (get_pruned_emissions :328-358 with the C-API's
    //      cutoff_prob = 1.0, cutoff_top_n = 40 >= n_classes: no pruning, index order, blank last)
    if (tid < C) s_logp[tid] = sttmath::glibc_logf(prob[tid] + kFltMin);
    if (tid == 0) {
      if ((double)prob[blank] < 0.999) start_expanding = 1;
      s_u[0] = start_expanding;
    }
    __syncthreads();
    start_expanding = s_u[0];
    if (!start_expanding) { __syncthreads(); continue; }
    if (overflow) { __syncthreads(); continue; }

    const float* L_score = s.score[sel];
    const float* L_b = s.b_prev[sel];
    const float* L_nb = s.nb_prev[sel];
    const uint32_t* L_node = s.node[sel];
    const uint32_t* L_ts = s.ts[sel];

    // ---- phase 1: min_cutoff (:134-146)
    float min_cutoff = kNegMax;
    bool full_beam = false;
    if (p.has_scorer) {
      float m = 3.402823466e+38f;
      for (uint32_t i = tid; i < n_live; i += NT) m = fminf(m, L_score[i]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, d));
      if ((tid & 31) == 0) s_red[tid >> 5] = m;
      __syncthreads();
      if (tid == 0) {
        float mm = s_red[0];
        for (int w = 1; w < NT / 32; ++w) mm = fminf(mm, s_red[w]);
        const double beta_pos = sv.beta > 0.0 ? sv.beta : 0.0;
        s_f[0] = (float)((double)mm + log((double)prob[blank]) - beta_pos);
      }
      __syncthreads();
      min_cutoff = s_f[0];
      full_beam = (n_live == (uint32_t)W);
    }
    for (uint32_t i = tid; i < n_live; i += NT) s_child[i] = 0;
    if (tid == 0) s_u[1] = 0;  // LM queue length
    __syncthreads();

    // ---- phase 2: which (parent, label) pairs already have a live child
    for (uint32_t j = tid; j < n_live; j += NT) {
      const uint32_t nd = L_node[j];
      const uint32_t pn = s.parent[nd];
      if (pn != kNone) {
        const uint32_t pi = s.live_slot[pn];
        if (pi != kNone) atomicOr(&s_child[pi], 1u << s.chr[nd]);
      }
    }
    __syncthreads();

    // ---- phase 2b: LM work list: live prefixes that will be extended by the space this step
    if (p.has_scorer) {
      for (uint32_t i = tid; i < n_live; i += NT) {
        const float sc = L_score[i];
        if (sc == kNegMax) continue;
        if (full_beam && s_logp[p.space_id] + sc < min_cutoff) continue;
        bool need = (s_child[i] >> p.space_id) & 1u;
        if (!need) need = sttscorer::fst_find(sv, s.dict_state[L_node[i]], p.space_id + 1) >= 0;
        if (need) s_lmq[atomicAdd(&s_u[1], 1u)] = i;
      }
      __syncthreads();
      const uint32_t n_lm = s_u[1];
      // spread items across warps: item q -> thread (q % NW) * 32 + q / NW
      constexpr int NW = NT / 32;
      for (uint32_t base = 0; base < n_lm; base += NT) {
        const int w = tid >> 5, l = tid & 31;
        const uint32_t q = base + (uint32_t)(l * NW + w);
        if (q < n_lm) {
          const uint32_t i = s_lmq[q];
          uint32_t wid, nw;
          s.lm_term[i] = lm_space_term(s, sv, L_node[i], &wid, &nw);
          s.lm_word[i] = wid;
          atomicAdd(&s_u[6], nw);  // instrumentation: words scored (Q of SURVEY 8d's decoder roofline)
          atomicAdd(&s_u[7], 1u);  // LM calls
        }
      }
      __syncthreads();
    }

    // ---- phase 3: updated values of the live prefixes (blank / repeat / pulled extension), :150-256
    for (uint32_t j = tid; j < n_live; j += NT) {
      const float sj = L_score[j];
      const uint32_t nd = L_node[j];
      const uint32_t cj = s.chr[nd];
      float nb = kNegMax, bcur = kNegMax;
      uint32_t ts_prev = kNone;  // kNone = keep current timesteps
      // extension pulled from a live parent
      bool has_ext = false;
      float lp_ext = kNegMax;
      uint32_t ts_par = kNone;
      bool parent_first = false;
      const uint32_t pn = s.parent[nd];
      if (pn != kNone) {
        const uint32_t pi = s.live_slot[pn];
        if (pi != kNone) {
          const float sp = L_score[pi];
          const float lc = s_logp[cj];
          if (sp != kNegMax && !(full_beam && lc + sp < min_cutoff)) {
            has_ext = true;
            const uint32_t cp = s.chr[pn];
            if (cj == cp) {
              const float bp = L_b[pi];
              lp_ext = (bp > kNegMax) ? lc + bp : kNegMax;
            } else {
              lp_ext = lc + sp;
            }
            if (p.has_scorer && (int)cj == p.space_id) {
              lp_ext += s.lm_term[pi];
              lp_ext = (float)((double)lp_ext + sv.beta);
            }
            ts_par = L_ts[pi];
            parent_first = visits_before(sp, cp, pi, sj, cj, j);
          }
        }
      }
      const bool alive = (sj != kNegMax);
      bool has_rep = false;
      float lp_rep = kNegMax;
      if (alive && cj != kRootChar && !(full_beam && s_logp[cj] + sj < min_cutoff)) {
        has_rep = true;
        lp_rep = s_logp[cj] + L_nb[j];
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
      s.c_score[j] = ns;
      s.c_a[j] = __float_as_uint(bcur);
      s.c_b[j] = __float_as_uint(nb);
      s.c_c[j] = ts_prev;
      s.c_key[j] = ((uint64_t)sortable(ns) << 32) | ((uint64_t)(255u - cj) << 24) | (uint64_t)(0xffffffu - j);
    }

    // ---- phase 4: new children.  Count, scan, then write at deterministic offsets.
    uint32_t n_new_total = 0;
    {
      uint32_t run_base = n_live;
      for (uint32_t base = 0; base < n_live; base += NT) {
        const uint32_t i = base + tid;
        uint32_t cnt = 0;
        uint32_t allow = 0;  // bit c set: create child c
        float si = kNegMax;
        uint32_t nd = 0;
        if (i < n_live) {
          si = L_score[i];
          nd = L_node[i];
          if (si != kNegMax) {
            const uint32_t have = s_child[i];
            if (p.has_scorer) {
              const int32_t st = s.dict_state[nd];
              const uint8_t* srec = sv.blob + sv.fst_states_off + (uint64_t)st * 20;
              const uint32_t pos = sttscorer::load_u32(srec + 4), narcs = sttscorer::load_u32(srec + 8);
              const uint8_t* arcs = sv.blob + sv.fst_arcs_off + (uint64_t)pos * 16;
              for (uint32_t a = 0; a < narcs; ++a) {
                const int c = (int)sttscorer::load_u32(arcs + (uint64_t)a * 16) - 1;
                if (c < 0 || c >= blank) continue;
                if ((have >> c) & 1u) continue;
                if (full_beam && s_logp[c] + si < min_cutoff) continue;
                allow |= 1u << c;
              }
            } else {
              for (int c = 0; c < blank; ++c) {
                if ((have >> c) & 1u) continue;
                allow |= 1u << c;
              }
            }
            cnt = __popc(allow);
          }
        }
        // block exclusive scan of cnt
        uint32_t incl = cnt;
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += o;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
          uint32_t v = (lane < NT / 32) ? s_warp[lane] : 0, inc2 = v;
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, inc2, d);
            if (lane >= d) inc2 += o;
          }
          if (lane < NT / 32) s_warp[lane] = inc2 - v;
          if (lane == 31) s_warp[NT / 32] = inc2;
        }
        __syncthreads();
        uint32_t e = run_base + s_warp[warp] + (incl - cnt);
        const uint32_t round_total = s_warp[NT / 32];
        __syncthreads();
        if (cnt && e + cnt <= s.cand_cap) {
          const uint32_t cp = s.chr[nd];
          const float bp = L_b[i];
          const int32_t st = s.dict_state[nd];
          while (allow) {
            const int c = __ffs(allow) - 1;
            allow &= allow - 1;
            float lp;
            if ((uint32_t)c == cp) lp = (bp > kNegMax) ? s_logp[c] + bp : kNegMax;
            else lp = s_logp[c] + si;
            int32_t nds = 0;
            uint32_t wid = 0;
            if (p.has_scorer) {
              const int32_t nxt = sttscorer::fst_find(sv, st, c + 1);
              nds = sttscorer::fst_is_final(sv, nxt) ? (int32_t)sv.fst_start : nxt;
              if (c == p.space_id) {
                lp += s.lm_term[i];
                lp = (float)((double)lp + sv.beta);
                wid = s.lm_word[i];
              }
            }
            s.c_score[e] = lp;
            s.c_a[e] = i;
            s.c_b[e] = (uint32_t)c;
            s.c_c[e] = (uint32_t)nds;
            s.c_d[e] = wid;
            s.c_key[e] = ((uint64_t)sortable(lp) << 32) | ((uint64_t)(255u - (uint32_t)c) << 24) |
                         (uint64_t)(0xffffffu - (e & 0xffffffu));
            ++e;
          }
        }
        run_base += round_total;
      }
      n_new_total = run_base - n_live;
    }
    __syncthreads();
    const uint32_t N = n_live + n_new_total;
    if (N > s.cand_cap) { overflow = 1; if (tid == 0) s.scalars[6] = 1; __syncthreads(); continue; }

    // ---- phase 5: exact top-W radix select on the 64-bit key (:263-274 nth_element + prefix_compare)
    uint64_t sel_prefix = 0, sel_mask = 0;
    if (N > (uint32_t)W) {
      uint32_t k_rem = (uint32_t)W;
      for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        for (int h = tid; h < 256; h += NT) s_hist[h] = 0;
        __syncthreads();
        for (uint32_t e = tid; e < N; e += NT) {
          const uint64_t key = s.c_key[e];
          if ((key & sel_mask) == sel_prefix) atomicAdd(&s_hist[(uint32_t)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
          // lane l owns bins [8l, 8l+8); find the bin where the count from the top crosses k_rem
          uint32_t mine = 0;
#pragma unroll
          for (int q = 0; q < 8; ++q) mine += s_hist[tid * 8 + q];
          uint32_t suffix = mine;  // inclusive sum over lanes >= tid
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_down_sync(0xffffffffu, suffix, d);
            if (tid + d < 32) suffix += o;
          }
          const uint32_t above = suffix - mine;  // elements in higher lanes
          if (above < k_rem && suffix >= k_rem) {
            uint32_t cum = above;
            for (int q = 7; q >= 0; --q) {
              const uint32_t hcount = s_hist[tid * 8 + q];
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
        __syncthreads();
        sel_prefix |= (uint64_t)s_u[2] << shift;
        sel_mask |= (uint64_t)255u << shift;
        k_rem = s_u[3];
        const bool done = (s_u[4] == k_rem);  // every element with this prefix is selected
        __syncthreads();
        if (done) break;
      }
    }

    // ---- phase 6: order-preserving compaction + commit (iterate_to_vec :159-190, remove :192-209)
    const uint32_t nsel = sel ^ 1u;
    float* N_score = s.score[nsel];
    float* N_b = s.b_prev[nsel];
    float* N_nb = s.nb_prev[nsel];
    uint32_t* N_node = s.node[nsel];
    uint32_t* N_ts = s.ts[nsel];
    uint32_t out_base = 0;
    uint32_t n_live_surv = 0;
    if (tid == 0) s_u[5] = arena_count;
    // pass A: live entries (they precede new ones in candidate order)
    for (uint32_t base = 0; base < n_live; base += NT) {
      const uint32_t e = base + tid;
      bool keep = false;
      if (e < n_live) keep = (N <= (uint32_t)W) || ((s.c_key[e] & sel_mask) >= sel_prefix);
      uint32_t total;
      const uint32_t pos = out_base + block_scan_flag<NT>(keep, s_warp, total);
      if (e < n_live) {
        const uint32_t nd = L_node[e];
        if (keep) {
          N_score[pos] = s.c_score[e];
          N_b[pos] = __uint_as_float(s.c_a[e]);
          N_nb[pos] = __uint_as_float(s.c_b[e]);
          N_node[pos] = nd;
          const uint32_t tp = s.c_c[e];
          if (tp != kNone) {
            const uint32_t id = ts_count + pos;
            if (id < s.ts_cap) { s.ts_parent[id] = tp; s.ts_val[id] = abs_t; }
            N_ts[pos] = id;
          } else {
            N_ts[pos] = L_ts[e];
          }
          s.live_slot[nd] = pos;
        } else {
          s.live_slot[nd] = kNone;
        }
      }
      out_base += total;
    }
    n_live_surv = out_base;
    // pass B: new entries
    for (uint32_t base = n_live; base < N; base += NT) {
      const uint32_t e = base + tid;
      bool keep = false;
      if (e < N) keep = (N <= (uint32_t)W) || ((s.c_key[e] & sel_mask) >= sel_prefix);
      uint32_t total;
      const uint32_t pos = out_base + block_scan_flag<NT>(keep, s_warp, total);
      if (e < N && keep) {
        const uint32_t pi = s.c_a[e];
        const uint32_t c = s.c_b[e];
        const float lp = s.c_score[e];
        const uint32_t pnode = L_node[pi];
        const unsigned long long hk = ht_make_key(pnode, c);
        uint32_t id = ht_find(s, hk);
        if (id == kNone) {
          id = atomicAdd(&s_u[5], 1u);  // fresh arena node
          if (id < s.arena_cap) {
            s.parent[id] = pnode;
            s.chr[id] = c;
            s.dict_state[id] = (int32_t)s.c_c[e];
            s.last_space[id] = ((int)c == p.space_id) ? id : s.last_space[pnode];
            s.word_id[id] = s.c_d[e];
            ht_insert(s, hk, id);
          }
        }
        if (id < s.arena_cap) s.live_slot[id] = pos;
        N_score[pos] = lp;
        N_b[pos] = kNegMax;
        N_nb[pos] = lp;
        N_node[pos] = id;
        if (lp > kNegMax) {  // "prefix_new->log_prob_nb_cur < log_p" (:246-251)
          const uint32_t tid2 = ts_count + pos;
          if (tid2 < s.ts_cap) { s.ts_parent[tid2] = L_ts[pi]; s.ts_val[tid2] = abs_t; }
          N_ts[pos] = tid2;
        } else {
          N_ts[pos] = kNone;
        }
      }
      out_base += total;
    }
    const uint32_t n_surv = out_base;
    __syncthreads();
    arena_count = s_u[5];
    ts_count += n_surv;
    if (arena_count > s.arena_cap || ts_count > s.ts_cap) overflow = 1;
    n_live = n_surv;
    sel = nsel;
    __threadfence_block();
    __syncthreads();
  }

  if (tid == 0) {
    s.scalars[0] = n_live;
    s.scalars[1] = sel;
    s.scalars[2] = arena_count;
    s.scalars[3] = ts_count;
    s.scalars[4] = abs_t;
    s.scalars[5] = start_expanding;
    s.scalars[6] = overflow;
    s.scalars[7] += s_u[6];   // words scored by the LM in this launch
    s.scalars[8] += s_u[7];   // LM calls
  }
}

// ------------------------------------------------------------------------------------------------ finalize
// DecoderState::decode(num_results) (:278-326) + ModelState::decode_metadata's token/timestep extraction.
struct FinalOut {
  int max_results, max_tokens;
  int* n_results;        // [1]
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
  const uint32_t sel = s.scalars[1];
  const float* L_score = s.score[sel];
  const uint32_t* L_node = s.node[sel];
  const uint32_t* L_ts = s.ts[sel];
  __shared__ unsigned long long s_best[NT / 32];
  __shared__ unsigned long long s_pick;

  // final scores -> c_score, keys -> c_key (live list itself is left untouched: IntermediateDecode is const)
  for (uint32_t i = tid; i < n_live; i += NT) {
    const uint32_t nd = L_node[i];
    float sc = L_score[i];
    const uint32_t c = s.chr[nd];
    if (p.has_scorer && i < (uint32_t)p.beam && c != kRootChar && (int)c != p.space_id) {
      float add = lm_final_term(s, p.scorer, nd);
      add = (float)((double)add + p.scorer.beta);
      sc = sc + add;
    }
    s.c_score[i] = sc;
    s.c_key[i] = ((uint64_t)sortable(sc) << 32) | ((uint64_t)(255u - c) << 24) | (uint64_t)(0xffffffu - i);
  }
  __syncthreads();
  int n_ret = (int)n_live < num_results ? (int)n_live : num_results;
  if (n_ret > o.max_results) n_ret = o.max_results;
  for (int r = 0; r < n_ret; ++r) {
    unsigned long long best = 0;
    for (uint32_t i = tid; i < n_live; i += NT) {
      const unsigned long long k = s.c_key[i];
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
      s_pick = b;
      const uint32_t i = 0xffffffu - (uint32_t)(b & 0xffffffu);
      s.c_key[i] = 0;  // remove from further rounds
      // backtrack tokens (get_path_vec) and timesteps (get_history)
      uint32_t nd = L_node[i];
      int len = 0;
      for (uint32_t w = nd; s.chr[w] != kRootChar; w = s.parent[w]) ++len;
      o.n_tokens[r] = len;
      o.confidence[r] = (double)s.c_score[i];
      int k = len;
      for (uint32_t w = nd; s.chr[w] != kRootChar; w = s.parent[w]) {
        --k;
        if (k < o.max_tokens) o.tokens[(size_t)r * o.max_tokens + k] = s.chr[w];
      }
      uint32_t tn = L_ts[i];
      int tlen = 0;
      for (uint32_t w = tn; w != kNone && w != 0; w = s.ts_parent[w]) ++tlen;
      k = tlen;
      for (uint32_t w = tn; w != kNone && w != 0; w = s.ts_parent[w]) {
        --k;
        if (k < o.max_tokens) o.timesteps[(size_t)r * o.max_tokens + k] = s.ts_val[w];
      }
    }
    __syncthreads();
  }
  if (tid == 0) *o.n_results = n_ret;
}

}  // namespace sttdec
