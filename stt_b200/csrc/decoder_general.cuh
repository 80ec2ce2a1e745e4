// The GENERAL beam-search kernel: every alphabet (up to 255 labels + blank), UTF-8 bytes-output scorers, vocabulary
// pruning (cutoff_prob / cutoff_top_n) -- the configurations the shared-memory kernel of decoder.cuh (<= 32 labels, word
// scorers, no pruning: everything the C API reaches with a 28-letter alphabet) does not cover.
//
// Same data structure (arena of prefix nodes + (parent, label) hash, per-node LM cache, timestep tree, 64-bit candidate
// keys, exact top-beam selection, order-preserving compaction) and the same arithmetic as decoder.cuh; what differs:
//   * the children of a prefix are enumerated from its dictionary state's ARC LIST (label, child state) and an existing
//     child is found through the hash table -- no 32-bit label masks anywhere;
//   * live lists and candidates live in GLOBAL memory (a beam of 500 over 256 classes can raise 128 000 candidates);
//   * get_pruned_emissions (ctc_beam_search_decoder.cpp:328-358): the classes of a step are ranked by probability when
//     cutoff_prob < 1 or cutoff_top_n < classes, the kept set gates every extension, and -- because the reference then
//     walks the classes in that order -- the blank is no longer guaranteed to come last, which changes how a prefix's
//     timestep is chosen (:163-178 "the blank label comes last" no longer holds); that order is reproduced;
//   * UTF-8 mode (Scorer::is_scoring_boundary scorer.cpp:271-295, make_ngram :353-381 with get_prev_grapheme
//     path_trie.cpp:113-126, distance_to_codepoint_boundary :128-141): the scored unit is a code point, scored on the NEW
//     prefix when its last byte completes one; DecoderState::decode's end-of-utterance term follows :286-300.
// This kernel is written for coverage, not speed: 65 us per timestep with the English alphabet under
// pruning (the two enumerations of the arc lists), 216 us in UTF-8 mode over 256 classes (LM evaluations of the candidates that
// complete a code point), against 18 us for the shared-memory kernel.
#pragma once
#include "decoder.cuh"

namespace sttdec {

struct GenParams {
  const uint2* gstate;   // dictionary FST, per state {first arc, number of arcs}
  const int2* garc;      // per arc {label, child's dictionary state = Start() when the arc's target is final}
  double cutoff_prob;
  int cutoff_top_n;
  uint32_t* scratch;     // kGenScratchArrays arrays of beam_cap words per utterance
  const uint32_t* byte_wid;   // UTF-8 scorers: vocabulary id of every ONE-byte code point (0 = not in the vocabulary), [256]
};
constexpr int kGenScratchArrays = 12;

// (parent, label) -> node id, or kNone.  Only called in phases where nobody inserts (see ht_insert).
__device__ __forceinline__ uint32_t ht_find_maybe(const Slot& s, uint32_t parent, uint32_t c) {
  const unsigned long long key = ht_pack(s.ht_gen, parent, c, 0) >> 24;
  uint32_t h = ht_hash(key) & s.ht_mask;
  for (;;) {
    const unsigned long long w = s.ht[h];
    if ((w >> 24) == key) return (uint32_t)(w & 0xffffffu);
    if ((uint32_t)(w >> 56) != s.ht_gen) return kNone;
    h = (h + 1) & s.ht_mask;
  }
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int NT>
__device__ __forceinline__ uint32_t gen_scan(uint32_t cnt, uint32_t* warp_sums /*[NT/32 + 1]*/, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += o;
  }
  __syncthreads();   // warp_sums may still be read from the previous call
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) {
    const uint32_t v = warp_sums[w];
    base += (w < warp) ? v : 0u;
    tot += v;
  }
  total = tot;
  return base + (incl - cnt);
}

template <int NT>
__global__ void __launch_bounds__(NT) decoder_general_kernel(Slot* slots, const StepInput* inputs, const DecodeParams p,
                                                              const GenParams g) {
  static_assert(NT == 256, "one thread per class when a row of probabilities is prepared");
  __shared__ Slot s_slot;
  const int tid = threadIdx.x;
  if (tid < (int)(sizeof(Slot) / 4)) reinterpret_cast<uint32_t*>(&s_slot)[tid] = reinterpret_cast<const uint32_t*>(&slots[blockIdx.x])[tid];
  __syncthreads();
  const Slot& s = s_slot;
  const StepInput in = inputs[blockIdx.x];
  const int C = p.n_classes, blank = C - 1, W = p.beam;
  const sttscorer::ScorerView& sv = p.scorer;
  const bool utf8 = p.has_scorer && sv.is_utf8;
  const bool sorted_order = (g.cutoff_prob < 1.0) || (g.cutoff_top_n < C);   // :337
  const uint32_t cap = s.beam_cap;

  __shared__ float s_pf[256], s_logp[256];
  __shared__ uint16_t s_rank[256], s_sorted[256];
  __shared__ uint8_t s_ok[256];
  __shared__ double s_logblank;
  __shared__ uint32_t s_u[8];
  __shared__ uint32_t s_warp[NT / 32 + 1];
  __shared__ float s_red[NT / 32];
  __shared__ __align__(16) uint32_t s_hist[256];

  // live list, double buffered: buffer 0 = the slot's own arrays, buffer 1 = scratch
  uint32_t* const scr = g.scratch + (size_t)blockIdx.x * kGenScratchArrays * cap;
  float* const b_score[2] = {s.score, reinterpret_cast<float*>(scr)};
  float* const b_b[2] = {s.b_prev, reinterpret_cast<float*>(scr + cap)};
  float* const b_nb[2] = {s.nb_prev, reinterpret_cast<float*>(scr + 2 * cap)};
  uint32_t* const b_node[2] = {s.node, scr + 3 * cap};
  uint32_t* const b_ts[2] = {s.ts, scr + 4 * cap};
  uint32_t* const plive = scr + 5 * cap;     // live index of the parent, or kNone
  float* const ub = reinterpret_cast<float*>(scr + 6 * cap);    // updated blank / non-blank log-probs of a live prefix
  float* const unb = reinterpret_cast<float*>(scr + 7 * cap);
  uint32_t* const utsp = scr + 8 * cap;      // timestep-tree parent chosen for it (kNone = keep); later old -> new slot
  uint32_t* const cofs = scr + 9 * cap;      // offset of a prefix's first new candidate
  uint32_t* const parked = scr + 10 * cap;   // per survivor: 0x80000000 | old slot, or the candidate index of a new one
  unsigned long long* const K = s.c_key;
  uint32_t* const P0 = s.c_p0;   // new candidate: parent live index | label << 16
  uint32_t* const P1 = s.c_p1;   // new candidate: node id to revive, or kNone

  uint32_t n_live = s.scalars[0];
  uint32_t arena_count = s.scalars[2];
  uint32_t ts_count = s.scalars[3];
  uint32_t abs_t = s.scalars[4];
  uint32_t start_expanding = s.scalars[5];
  uint32_t overflow = s.scalars[6];
  int cur = 0;
  const bool use64 = in.probs64 != nullptr;

#ifdef STT_GEN_PROF
  unsigned long long gp_ph[7] = {0, 0, 0, 0, 0, 0, 0}, gp_n = 0;
  long long gp_t = clock64();
#define GEN_MARK(k) do { const long long _t = clock64(); gp_ph[k] += (unsigned long long)(_t - gp_t); gp_t = _t; } while (0)
#else
#define GEN_MARK(k) do { } while (0)
#endif
  for (int step = 0; step < in.n_steps; ++step, ++abs_t) {
    float* const score = b_score[cur];
    float* const bprev = b_b[cur];
    float* const nbprev = b_nb[cur];
    uint32_t* const node = b_node[cur];
    uint32_t* const ts = b_ts[cur];
    // ---- the row: class log-probs, the pruned class set and its order (get_pruned_emissions :328-358)
    double pd = 0.0;
    float pf = 0.0f;
    __syncthreads();
    if (tid < C) {
      pd = use64 ? in.probs64[(size_t)step * C + tid] : (double)in.probs[(size_t)step * C + tid];
      pf = (float)pd;
      s_pf[tid] = pf;
    }
    __syncthreads();
    uint32_t rank = (uint32_t)tid;
    if (sorted_order && tid < C) {
      // std::sort by probability, descending; equal probabilities keep index order here (unspecified in the reference)
      rank = 0;
      for (int c2 = 0; c2 < C; ++c2) {
        const float q = s_pf[c2];
        rank += (q > pf || (q == pf && c2 < tid)) ? 1u : 0u;
      }
      s_sorted[rank] = (uint16_t)tid;
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t cutoff_len = (uint32_t)C;
      if (g.cutoff_prob < 1.0) {
        double cum = 0.0;
        cutoff_len = 0;
        for (int i = 0; i < C; ++i) {
          cum += s_pf[s_sorted[i]];
          cutoff_len += 1;
          if (cum >= g.cutoff_prob || cutoff_len >= (uint32_t)g.cutoff_top_n) break;
        }
      }
      s_u[0] = cutoff_len;
    }
    __syncthreads();
    if (tid < C) {
      s_rank[tid] = (uint16_t)rank;
      s_ok[tid] = rank < s_u[0] ? 1 : 0;
      s_logp[tid] = sttmath::glibc_logf(pf + kFltMin);
      if (tid == blank) {
        s_u[1] = pd < 0.999 ? 1u : 0u;
        s_logblank = log(pd);
      }
    }
    {
      float m = 3.402823466e+38f;
      for (uint32_t i = tid; i < n_live; i += NT) m = fminf(m, score[i]);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, d));
      if ((tid & 31) == 0) s_red[tid >> 5] = m;
    }
    __syncthreads();
    start_expanding |= s_u[1];
    if (!start_expanding || overflow) continue;
    float min_cutoff = kNegMax;
    bool full_beam = false;
    if (p.has_scorer) {
      float mm = s_red[0];
#pragma unroll
      for (int w = 1; w < NT / 32; ++w) mm = fminf(mm, s_red[w]);
      const double beta_pos = sv.beta > 0.0 ? sv.beta : 0.0;
      min_cutoff = (float)((double)mm + s_logblank - beta_pos);
      full_beam = (n_live == (uint32_t)W);
    }
    const bool blank_ok = s_ok[blank] != 0;
    const float lp_blank = s_logp[blank];

    // ---- where each live prefix's parent sits in the live list (Node::live_slot is stamped when a step commits; a
    //      stale stamp of a prefix that has left the beam is recognised because the slot holds another node)
    for (uint32_t j = tid; j < n_live; j += NT) {
      const uint32_t pn = s.nodes[node[j]].parent;
      uint32_t pi = kNone;
      if (pn != kNone) {
        pi = s.nodes[pn].live_slot;
        if (pi >= n_live || node[pi] != pn) pi = kNone;
      }
      plive[j] = pi;
    }
    __syncthreads();

    // LM term (times alpha, boosts included) of extending live prefix `pi` by label c to the child `child_node`
    // (kNone when the child does not exist yet); *scored says whether this extension is an LM boundary
    auto lm_term = [&](uint32_t pi_node, uint32_t c, uint32_t child_node, bool* scored) -> float {
      *scored = false;
      if (!p.has_scorer) return 0.0f;
      if (!utf8) {
        if ((int)c != p.space_id) return 0.0f;   // is_scoring_boundary: new_label == SPACE_ID_ (scorer.cpp:293)
        *scored = true;
        uint32_t wid, nw;
        double cond = lm_eval_node(s, p, pi_node, kStopUnknown, &wid, &nw);
        if (p.n_hot > 0) cond += (double)hot_word_boost(s, p, pi_node, wid);
        return (float)(cond * sv.alpha);
      }
      // UTF-8: the NEW prefix is scored when its last byte completes a code point (ctc_beam_search_decoder.cpp:211-218)
      double cond;
      if (child_node != kNone) {
        const Node cn = s.nodes[child_node];
        if (!utf8_completes(cn.ord)) return 0.0f;
        cond = utf8_node_cond(s, p, child_node);
      } else {
        const Node pn = s.nodes[pi_node];
        uint32_t ord, ctx;
        utf8_child_fields(sv, pn, pi_node, c, &ord, &ctx);
        if (!utf8_completes(ord)) return 0.0f;
        bool have = false;
        if (p.n_hot == 0) {   // one descent from the carried state (hot words need the window's ids: literal path)
          sttscorer::LmState out_unused;
          uint32_t meta_unused;
          // a one-byte code point's vocabulary id is a table read (engine.cu builds the table when the scorer is enabled)
          const uint32_t wid = ((ord & 0xffu) == 1u && g.byte_wid && sv.label_len[c] == 1) ? g.byte_wid[sv.label_bytes[c][0]]
                                                                                          : utf8_unit_id(s, sv, pi_node, (ord & 0xffu) - 1u, c);
          have = utf8_cond_carried(s, p, ctx, wid, &cond, &out_unused, &meta_unused);
        }
        if (!have) {
          float boost = 0.0f;
          cond = utf8_window_cond(s, p, pi_node, c, &boost) + (double)boost;
        }
      }
      *scored = true;
      return (float)(cond * sv.alpha);
    };

    GEN_MARK(0);   // row preparation, cutoff, parent slots
    // ---- updated values of the live prefixes (:150-256), events in the order the reference's loops produce them
    for (uint32_t j = tid; j < n_live; j += NT) {
      const float sj = score[j];
      const uint32_t nj = node[j];
      const Node ndj = s.nodes[nj];
      const uint32_t cj = ndj.chr;
      float nb = kNegMax, bcur = kNegMax;
      uint32_t ts_prev = kNone;
      const bool alive = (sj != kNegMax);
      const bool blank_first = sorted_order && cj != kRootChar && s_rank[blank] < s_rank[cj];
      const bool do_blank = alive && blank_ok && !(full_beam && lp_blank + sj < min_cutoff);
      if (do_blank) bcur = lp_blank + sj;   // log_sum_exp(-FLT_MAX, lp_b); with the blank first, "nb_cur < log_p" clears a
                                            // timestep choice that has not been made yet: no effect
      if (cj != kRootChar && s_ok[cj]) {
        const float lc = s_logp[cj];
        bool has_ext = false, parent_first = false;
        float lp_ext = kNegMax;
        uint32_t ts_par = kNone;
        const uint32_t pi = plive[j];
        if (pi != kNone) {
          const float sp = score[pi];
          if (sp != kNegMax && !(full_beam && lc + sp < min_cutoff)) {
            has_ext = true;
            const uint32_t cp = s.nodes[node[pi]].chr;
            if (cj == cp) {
              const float bp = bprev[pi];
              lp_ext = (bp > kNegMax) ? lc + bp : kNegMax;
            } else {
              lp_ext = lc + sp;
            }
            bool scored;
            const float term = lm_term(node[pi], cj, nj, &scored);
            if (scored) {
              lp_ext += term;
              lp_ext = (float)((double)lp_ext + sv.beta);
            }
            ts_par = ts[pi];
            parent_first = visits_before(sp, cp, pi, sj, cj, j);
          }
        }
        bool has_rep = false;
        float lp_rep = kNegMax;
        if (alive && !(full_beam && lc + sj < min_cutoff)) {
          has_rep = true;
          lp_rep = lc + nbprev[j];
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
      }
      if (do_blank && !blank_first && nb < bcur) ts_prev = kNone;   // the blank came last (:166-170)
      const float ns = sttmath::log_sum_exp(bcur, nb);
      K[j] = make_key(ns, cj, j);
      ub[j] = bcur;
      unb[j] = nb;
      utsp[j] = ts_prev;
    }

    GEN_MARK(1);   // live update
    // ---- new children: count, scan, emit (two identical enumerations so that candidates land at deterministic offsets)
    uint32_t n_new = 0;
    for (int pass = 0; pass < 2; ++pass) {
      uint32_t run_base = 0;
      for (uint32_t base = 0; base < n_live; base += NT) {
        const uint32_t i = base + tid;
        uint32_t count = 0;
        if (i < n_live && score[i] != kNegMax) {
          const float si = score[i];
          const uint32_t ni = node[i];
          const Node ndi = s.nodes[ni];
          uint32_t a0 = 0, na = (uint32_t)(C - 1);
          if (p.has_scorer) {
            const uint2 st = g.gstate[ndi.dict];
            a0 = st.x;
            na = st.y;
          }
          uint32_t e = n_live + (pass ? cofs[i] : 0u);
          for (uint32_t a = 0; a < na; ++a) {
            uint32_t c = a;
            if (p.has_scorer) c = (uint32_t)g.garc[a0 + a].x;
            if (c >= (uint32_t)blank || !s_ok[c]) continue;
            if (full_beam && s_logp[c] + si < min_cutoff) continue;
            // Node::child_mask as a 32-bit filter over labels (bit = label mod 32, set when a child is created): nearly every
            // (prefix, label) pair has never had a node, and the filter spares those the hash probe -- a dependent global
            // round trip per arc, which is what this phase's time consisted of
            const uint32_t existing = ((ndi.child_mask >> (c & 31u)) & 1u) ? ht_find_maybe(s, ni, c) : kNone;
            if (existing != kNone) {
              const uint32_t ls = s.nodes[existing].live_slot;
              if (ls < n_live && node[ls] == existing) continue;   // a live child pulls this extension itself
            }
            if (pass) {
              float lp;
              if (c == ndi.chr) lp = (bprev[i] > kNegMax) ? s_logp[c] + bprev[i] : kNegMax;
              else lp = s_logp[c] + si;
              bool scored;
              const float term = lm_term(ni, c, existing, &scored);
              if (scored) {
                lp += term;
                lp = (float)((double)lp + sv.beta);
              }
              K[e] = make_key(lp, c, e);
              P0[e] = i | (c << 16);
              P1[e] = existing;
              ++e;
            }
            ++count;
          }
        }
        if (!pass) {
          uint32_t total;
          const uint32_t off = run_base + gen_scan<NT>(count, s_warp, total);
          if (i < n_live) cofs[i] = off;
          run_base += total;
        }
      }
      if (!pass) n_new = run_base;
      __syncthreads();
      if (!pass) GEN_MARK(6);   // the counting enumeration
      if (!pass && n_live + n_new > s.cand_cap) break;
    }
    const uint32_t N = n_live + n_new;
    if (N > s.cand_cap) { overflow = 1; continue; }
    __threadfence_block();
    __syncthreads();

    GEN_MARK(2);   // children
#ifdef STT_GEN_PROF
    gp_n += N;
#endif
    // ---- exact top-W selection on the 64-bit key: radix passes, most significant byte first (:263-274)
    unsigned long long sel_prefix = 0, sel_mask = 0;
    if (N > (uint32_t)W) {
      uint32_t k_rem = (uint32_t)W;
      for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
        s_hist[tid] = 0;
        __syncthreads();
        for (uint32_t e = tid; e < N; e += NT) {
          const unsigned long long key = K[e];
          if ((key & sel_mask) == sel_prefix) atomicAdd(&s_hist[(uint32_t)(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
          uint32_t cum = 0;
          for (int bin = 255; bin >= 0; --bin) {
            const uint32_t h = s_hist[bin];
            if (cum + h >= k_rem) {
              s_u[2] = (uint32_t)bin;
              s_u[3] = k_rem - cum;
              s_u[4] = h;
              break;
            }
            cum += h;
          }
        }
        __syncthreads();
        sel_prefix |= (unsigned long long)s_u[2] << shift;
        sel_mask |= (unsigned long long)255u << shift;
        k_rem = s_u[3];
        const bool done = (s_u[4] == k_rem);
        __syncthreads();
        if (done) break;   // every key with this prefix is selected
      }
    }

    GEN_MARK(3);   // select
    // ---- order-preserving compaction + commit (iterate_to_vec :159-190, remove :192-209)
    float* const nscore = b_score[cur ^ 1];
    float* const nbb = b_b[cur ^ 1];
    float* const nnb = b_nb[cur ^ 1];
    uint32_t* const nnode = b_node[cur ^ 1];
    uint32_t* const nts = b_ts[cur ^ 1];
    if (tid == 0) s_u[5] = arena_count;
    uint32_t out_base = 0;
    for (uint32_t g0 = 0; g0 < N; g0 += NT) {
      const uint32_t e = g0 + tid;
      bool keep = false;
      unsigned long long key = 0;
      if (e < N) {
        key = K[e];
        keep = (N <= (uint32_t)W) || ((key & sel_mask) >= sel_prefix);
      }
      uint32_t total;
      const uint32_t pos = out_base + gen_scan<NT>(keep ? 1u : 0u, s_warp, total);
      if (e < N) {
        if (e < n_live) {
          if (keep) {
            nscore[pos] = unsortable((uint32_t)(key >> 32));
            nbb[pos] = ub[e];
            nnb[pos] = unb[e];
            nnode[pos] = node[e];
            const uint32_t tp = utsp[e];
            if (tp != kNone) {
              const uint32_t id = ts_count + pos;
              if (id < s.ts_cap) s.ts_tree[id] = make_uint2(tp, abs_t);
              nts[pos] = id;
            } else {
              nts[pos] = ts[e];
            }
            parked[pos] = 0x80000000u | e;
          }
        } else if (keep) {
          const float lp = unsortable((uint32_t)(key >> 32));
          nscore[pos] = lp;
          nbb[pos] = kNegMax;
          nnb[pos] = lp;
          parked[pos] = e;
        }
      }
      out_base += total;
    }
    __threadfence_block();
    __syncthreads();
    GEN_MARK(4);   // compaction
    // one thread per survivor: a new one gets its arena node (or is revived under its old identity), everybody stamps
    // Node::live_slot
    for (uint32_t pos = tid; pos < out_base; pos += NT) {
      const uint32_t pk = parked[pos];
      if (pk & 0x80000000u) {
        s.nodes[nnode[pos]].live_slot = pos;
        continue;
      }
      const uint32_t pw = P0[pk];
      const uint32_t pi = pw & 0xffffu, c = pw >> 16;
      const uint32_t pnode = node[pi];
      const Node par = s.nodes[pnode];
      uint32_t id = P1[pk];
      const float lp = nscore[pos];
      if (id == kNone) {
        id = atomicAdd(&s_u[5], 1u);
        if (id < s.arena_cap) {
          const bool is_space = !utf8 && ((int)c == p.space_id);
          Node n;
          n.parent = pnode;
          n.chr = c;
          n.dict = 0;
          if (p.has_scorer) {
            // the child's dictionary state: the arc with this label (sorted by label)
            const uint2 st = g.gstate[par.dict];
            uint32_t lo = 0, hi = st.y;
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              if ((uint32_t)g.garc[st.x + mid].x < c) lo = mid + 1;
              else hi = mid;
            }
            n.dict = g.garc[st.x + lo].y;
          }
          n.live_slot = pos;
          n.lm_wid = kNone;
          n.child_mask = 0;
          uint32_t meta_init = kNone;
          double cond_init = 0.0;
          bool cond_known = false;
          if (utf8) {
            uint32_t ord, ctx;
            utf8_child_fields(sv, par, pnode, c, &ord, &ctx);
            n.ord = ord;
            n.last_space = ctx;
            if (p.has_scorer && utf8_completes(ord)) {
              // this node ends a code point: its LM term and the KenLM state after it (the context of the next code
              // point) are computed once, here
              const uint32_t wid = ((ord & 0xffu) == 1u && g.byte_wid && sv.label_len[c] == 1) ? g.byte_wid[sv.label_bytes[c][0]]
                                                                                              : utf8_unit_id(s, sv, pnode, (ord & 0xffu) - 1u, c);
              sttscorer::LmState st_out;
              uint32_t meta_out;
              if (utf8_cond_carried(s, p, ctx, wid, &cond_init, &st_out, &meta_out)) {
                for (int q = 0; q < (int)st_out.length && q < kStateWords; ++q) {
                  s.lm_sw[(size_t)id * kStateWords + q] = st_out.words[q];
                  s.lm_sb[(size_t)id * kStateWords + q] = st_out.backoff[q];
                }
                meta_init = meta_out;
                cond_known = true;
                n.lm_wid = wid;
              }
            }
          } else {
            n.last_space = is_space ? id : par.last_space;
            n.word_id = 0;
            if (p.has_scorer && is_space) {
              // as decoder.cuh phase 6 (ii): the id of the word this space terminates, and the KenLM state after it =
              // the context of the next word, carried in this node's LM rows
              uint32_t wid, nw;
              (void)lm_eval_node(s, p, pnode, kStopUnknown, &wid, &nw);
              n.word_id = wid;
              if (par.chr == kRootChar || (int)par.chr == p.space_id) {
                meta_init = 0u | (0u << 8) | (1u << 16);
              } else {
                meta_init = s.lm_meta[pnode];
                const uint32_t len = meta_init == kNone ? 0u : (meta_init & 0xffu);
                for (uint32_t q = 0; q < len && q < (uint32_t)kStateWords; ++q) {
                  s.lm_sw[(size_t)id * kStateWords + q] = s.lm_sw[(size_t)pnode * kStateWords + q];
                  s.lm_sb[(size_t)id * kStateWords + q] = s.lm_sb[(size_t)pnode * kStateWords + q];
                }
              }
            }
          }
          s.nodes[id] = n;
          s.lm_meta[id] = meta_init;
          if (cond_known) s.lm_cond[id] = cond_init;
          else reinterpret_cast<unsigned long long*>(s.lm_cond)[id] = kLmUnset;
          ht_insert(s, pnode, c, id);
          atomicOr(&s.nodes[pnode].child_mask, 1u << (c & 31u));
        }
      } else {
        s.nodes[id].live_slot = pos;
      }
      nnode[pos] = id;
      if (lp > kNegMax) {   // "prefix_new->log_prob_nb_cur < log_p" (:246-251)
        const uint32_t tid2 = ts_count + pos;
        if (tid2 < s.ts_cap) s.ts_tree[tid2] = make_uint2(ts[pi], abs_t);
        nts[pos] = tid2;
      } else {
        nts[pos] = kNone;
      }
    }
    __threadfence_block();
    __syncthreads();
    arena_count = s_u[5];
    ts_count += out_base;
    if (arena_count > s.arena_cap || ts_count > s.ts_cap) overflow = 1;
    n_live = out_base;
    cur ^= 1;
    GEN_MARK(5);   // commit
  }

  // ---- leave the live list in the slot's own arrays
  __syncthreads();
  if (cur == 1) {
    for (uint32_t i = tid; i < n_live; i += NT) {
      s.score[i] = b_score[1][i];
      s.b_prev[i] = b_b[1][i];
      s.nb_prev[i] = b_nb[1][i];
      s.node[i] = b_node[1][i];
      s.ts[i] = b_ts[1][i];
    }
  }
  if (tid == 0) {
    s.scalars[0] = n_live;
    s.scalars[2] = arena_count;
    s.scalars[3] = ts_count;
    s.scalars[4] = abs_t;
    s.scalars[5] = start_expanding;
    s.scalars[6] = overflow;
#ifdef STT_GEN_PROF
    for (int q = 0; q < 6; ++q) s.scalars[8 + q] = (uint32_t)(gp_ph[q] >> 10);   // kilo-cycles per phase
    s.scalars[15] = (uint32_t)(gp_ph[6] >> 10);
    s.scalars[14] = (uint32_t)(gp_n / (unsigned long long)(in.n_steps > 0 ? in.n_steps : 1));   // mean candidates per step
#endif
  }
}

}  // namespace sttdec
