"""Parity at the HEADLINE configuration (BASELINE.json configs[2]: n_hidden 2048, 256 utterances x 10 s -> T = 500,
beam 500, KenLM scorer) -- the exact kernel instances the benchmark number comes from:

  * acoustic model: `lstm_pp_kernel<4,3>` (ping-pong CTA pairs, B > 128) + the 128x256 tcgen05 GEMMs at M = 128 000,
    probabilities of utterances 0 / 127 / 128 / 255 (both ping-pong groups, both CTA ranks of a pair) against the
    oracle in the SAME precision mode (oracle/am_modes.py "f16": fp16 operands, fp32 accumulate) within 2e-3, on the
    CALIBRATED benchmark model (x200 output layer), and against the fp32 oracle within the envelope that mode itself
    keeps from fp32 (tools/precision_study.py; the reference's own default arithmetic, hybrid int8, is 25x further);
  * decoder: all 256 utterances at T = 500, beam 500 on the GPU's own probabilities, bit-exact (tokens, timesteps,
    confidence) against the GENUINE reference ctc_beam_search_decoder_batch; beam 100 and 2000 at T = 500 on a subset;
  * a 1000-case random / adversarial matrix of small decodes (near-ties, repeated characters, zombie prefixes,
    revival bursts, quantised probabilities) against the reference.
"""
import os

import numpy as np
import pytest

from conftest import SCORER

pytestmark = pytest.mark.gpu

# Tolerances.  SURVEY 8d parity gate 2 (|dp| <= 2e-3) is stated for softmax outputs of unit-scale logits and is checked
# as such on the plain random-init model.  The BENCHMARK model multiplies its output layer by 200 (synth.make_ctc_like,
# that is what makes a random network emit CTC-like text), so the same hidden-state agreement shows up 200x larger at
# the logits: there the gates are (measured on B200, profiles/r02_parity_headline.md)
#   vs the oracle in the SAME precision mode (fp16 operands / fp32 accumulate)  1.0e-2  -> gate 2e-2
#     (not 0: fp32 accumulation order differs, and a pre-activation that lands on the other side of an fp16 rounding
#      boundary moves that activation by one fp16 ulp, which 500 recurrent steps and the x200 layer amplify),
#   vs the fp32 oracle                                                           2.8e-2  -> gate 4e-2
#     (2.0e-2 of it is operand rounding alone, tools/precision_study.py; the reference's own default arithmetic, TFLite
#      hybrid int8, is 0.30-0.56 from fp32 on the same model),
#   frames whose arg-max class differs from fp32                                  0.1 %   -> gate 0.5 %,
# and transcripts are compared with the fp32 CPU path utterance by utterance (test_headline_transcripts_vs_fp32_cpu_path).
PROBS_ATOL_UNIT_SCALE = 2e-3
PROBS_ATOL_SAME_MODE = 2e-2
PROBS_ATOL_VS_FP32 = 4e-2
ARGMAX_FLIP_FRAC = 5e-3
N_SAMPLES = 160000
B = 256


@pytest.fixture(scope="module")
def headline(tmp_path_factory):
    """One forward + decode of the full headline batch; everything below reads from it."""
    from stt_b200 import Model, synth
    w = synth.bench_weights(n_hidden=2048)
    path = str(tmp_path_factory.mktemp("headline") / "bench.sttw")
    synth.write_model(path, w, beam_width=500)
    m = Model(path)
    m.enableExternalScorer(SCORER)
    pcms = [synth.make_pcm(N_SAMPLES, utt=u) for u in range(B)]
    b = m.createBatch(B, N_SAMPLES)
    b.upload(pcms)
    b.forward()
    b.decode(1)
    b.fetch()
    probs = [b.probs(u) for u in range(B)]
    results = [b.results(u) for u in range(B)]
    assert all(p.shape == (500, 29) for p in probs)
    # the same batch through the SAME kernels with the plain (unit-scale) output layer
    wp = synth.make_weights(n_hidden=2048)
    ppath = str(tmp_path_factory.mktemp("headline_plain") / "plain.sttw")
    synth.write_model(ppath, wp, beam_width=500)
    mp = Model(ppath)
    bp = mp.createBatch(B, N_SAMPLES)
    bp.upload(pcms)
    bp.forward()
    plain = {u: bp.probs(u) for u in (0, 127, 128, 255)}
    del bp, mp
    return {"model": m, "weights": w, "pcms": pcms, "probs": probs, "results": results, "plain_weights": wp,
            "plain_probs": plain}


def test_headline_am_unit_scale_vs_fp32_oracle(oracle, headline):
    """SURVEY 8d gate 2 on the headline kernel instances (B = 256, T = 500, n_hidden 2048): |dp| <= 2e-3 against the fp32
    oracle with the plain random-init output layer."""
    from oracle.am_modes import ModeAM
    full = ModeAM(headline["plain_weights"], "fp32")
    worst = 0.0
    for u, got in headline["plain_probs"].items():
        _, mfcc = oracle.features_only(headline["pcms"][u])
        d = float(np.abs(got - full.forward_features(mfcc)).max())
        print("plain model utt %3d: max|dp| vs fp32 oracle %.3e" % (u, d))
        worst = max(worst, d)
    assert worst <= PROBS_ATOL_UNIT_SCALE


def test_headline_am_vs_same_precision_oracle(oracle, headline):
    from oracle.am_modes import ModeAM
    w = headline["weights"]
    same = ModeAM(w, "f16", knobs={"weights": True, "features": True, "activations": True, "h_feedback": True,
                                   "h_output": True})   # gate functions are fp32-accurate on the GPU (exact_h)
    full = ModeAM(w, "fp32")
    worst_same = worst_fp32 = 0.0
    flips = 0
    for u in (0, 127, 128, 255):
        _, mfcc = oracle.features_only(headline["pcms"][u])
        got = headline["probs"][u]
        ps = same.forward_features(mfcc)
        pf = full.forward_features(mfcc)
        assert got.shape == ps.shape == pf.shape
        d_same, d_fp32 = float(np.abs(got - ps).max()), float(np.abs(got - pf).max())
        flips += int((got.argmax(1) != pf.argmax(1)).sum())
        big = pf > 1e-4   # log-domain distance ~ |d logit| (the softmax normaliser moves little)
        dl = float(np.abs(np.log(got[big]) - np.log(pf[big])).max())
        print("utt %3d: max|dp| vs f16-mode oracle %.3e, vs fp32 oracle %.3e, max|d ln p| vs fp32 %.3e" % (u, d_same, d_fp32, dl))
        worst_same, worst_fp32 = max(worst_same, d_same), max(worst_fp32, d_fp32)
        np.testing.assert_allclose(got.sum(1), 1.0, rtol=1e-4)
    print("headline AM: worst vs same-mode %.3e (tol %.0e), vs fp32 %.3e (tol %.0e), arg-max flips vs fp32 %d/2000"
          % (worst_same, PROBS_ATOL_SAME_MODE, worst_fp32, PROBS_ATOL_VS_FP32, flips))
    assert worst_same <= PROBS_ATOL_SAME_MODE
    assert worst_fp32 <= PROBS_ATOL_VS_FP32
    assert flips <= ARGMAX_FLIP_FRAC * 2000


def _edit_distance(a, b):
    """Levenshtein distance between two label sequences."""
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


# Gate 3 restated with measurements (B200, round 2): the synthetic benchmark model emits a "word salad" whose competing
# hypotheses are separated by LM scores alone, so a |d logit| of 0.1 flips the winning hypothesis somewhere in a 10 s
# utterance for ~30 % of the utterances -- always locally: a word or two out of ~35.  What is gated is therefore the
# DISTANCE between the product's transcripts and the fp32 CPU path's, in labels (character error rate), and that this
# distance is smaller than the one the reference's OWN default arithmetic (TFLite hybrid int8, oracle mode "hybrid8")
# keeps from the same fp32 path on the same utterances.
CER_VS_FP32_MAX = 0.05   # measured 2.56 % over 224 utterances; 3.6 % over the first 32 (a slower host gets through fewer)
IDENTICAL_FRAC_MIN = 0.55


def test_headline_transcripts_vs_fp32_cpu_path(oracle, ref_decoder, headline, english):
    """SURVEY 8d gate 3: end-to-end transcripts of the product against the CPU path (oracle MFCC + restated fp32 acoustic
    model + genuine reference decoder) on the benchmark's full-length utterances.  As many of the 256 as the host gets
    through in ~4 minutes (all of them on a 128-core box); every divergence is listed with its edit distance."""
    import time
    from oracle.am_modes import ModeAM
    from oracle.cpu_path import CpuPath
    cp = CpuPath(headline["weights"], SCORER, english, 500)
    fp32_tok = {}
    try:
        done, same, diverged = 0, 0, []
        err = ref_len = 0
        t0 = time.time()
        while done < B and time.time() - t0 < 240:
            n = min(cp.n_streams, B - done)
            _, _, res, _ = cp.run(headline["pcms"][done:done + n])
            for k in range(n):
                u = done + k
                r, g = list(res[k][0][1]), list(headline["results"][u][0][1])
                fp32_tok[u] = r
                ref_len += len(r)
                if r == g:
                    same += 1
                else:
                    d = _edit_distance(r, g)
                    err += d
                    diverged.append((u, d))
            done += n
    finally:
        cp.close()
    cer = err / float(max(1, ref_len))
    print("end-to-end transcripts identical to the fp32 CPU path: %d/%d; label error rate %.4f (%d edits / %d labels); "
          "diverged (utt, edits): %s" % (same, done, cer, err, ref_len, diverged))
    assert done >= 8
    # the reference's own default arithmetic on a few of the same utterances
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    h8 = ModeAM(headline["weights"], "hybrid8")
    h_err = h_len = g_err = 0
    for u in sorted(fp32_tok)[:4]:
        _, mfcc = oracle.features_only(headline["pcms"][u])
        tok = list(o.ref_decode(h8.forward_features(mfcc).astype(np.float64), alpha, 500, sc)[0][1])
        h_err += _edit_distance(fp32_tok[u], tok)
        g_err += _edit_distance(fp32_tok[u], list(headline["results"][u][0][1]))
        h_len += len(fp32_tok[u])
    print("same 4 utterances: product %d edits, TFLite-hybrid-int8 arithmetic %d edits, of %d labels" % (g_err, h_err, h_len))
    assert cer <= CER_VS_FP32_MAX, "label error rate vs the fp32 CPU path %.4f" % cer
    assert same >= IDENTICAL_FRAC_MIN * done, "too many transcripts differ from the fp32 CPU path: %s" % diverged
    assert g_err <= h_err, "the product should be closer to fp32 than the reference's default int8 arithmetic"


def test_headline_rows_equal_single_utterance_path(headline):
    """An LSTM couples nothing across utterances: row u of the 256-batch (ping-pong pair kernel) must equal the
    batch-of-one result (single-CTA kernel, cluster multicast) bit for bit -- same fp16 operands, same accumulation
    order along K."""
    m = headline["model"]
    b1 = m.createBatch(1, N_SAMPLES)
    for u in (0, 127, 128, 255):
        b1.upload([headline["pcms"][u]])
        b1.forward()
        np.testing.assert_array_equal(headline["probs"][u], b1.probs(0))


def test_headline_partial_batches_equal_full(headline):
    """129 and 200 utterances (second ping-pong group partly filled) give the same rows as the full batch."""
    m = headline["model"]
    for n in (129, 200):
        b = m.createBatch(n, N_SAMPLES)
        b.upload(headline["pcms"][:n])
        b.forward()
        for u in (0, 63, 64, 127, 128, n - 1):
            np.testing.assert_array_equal(headline["probs"][u], b.probs(u))


def _same(g, r):
    (gc, gt, gts), (rc, rt, rts) = g, r
    return list(gt) == list(rt) and list(gts) == list(rts) and gc == rc


def test_headline_decoder_all_256_utterances(ref_decoder, headline, english):
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    probs = np.stack(headline["probs"]).astype(np.float64)
    ref = o.ref_decode_batch(probs, [500] * B, alpha, 500, sc, num_processes=os.cpu_count() or 8)
    bad = [u for u in range(B) if not _same(headline["results"][u][0], ref[u][0])]
    n_tok = sum(len(headline["results"][u][0][1]) for u in range(B))
    print("headline decoder: %d/%d utterances bit-identical to the reference (%d tokens in total)" % (B - len(bad), B, n_tok))
    assert n_tok > 20 * B, "the benchmark model should emit text"
    assert not bad, "utterances that differ from the reference decoder: %s" % bad[:16]


@pytest.mark.parametrize("beam,n_utt", [(100, 32), (2000, 6)])
def test_headline_length_other_beams(ref_decoder, headline, english, beam, n_utt):
    """T = 500 at beam 100 (narrow) and 2000 (wide instantiation: candidates in global memory) on the GPU's probs."""
    from stt_b200 import Model
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    m = headline["model"]
    m.setBeamWidth(beam)
    try:
        probs = np.stack(headline["probs"][:n_utt])
        b = m.createBatch(n_utt, N_SAMPLES)
        b.set_probs(probs, [500] * n_utt)
        b.decode(num_results=2)
        b.fetch()
        ref = o.ref_decode_batch(probs.astype(np.float64), [500] * n_utt, alpha, beam, sc,
                                 num_processes=os.cpu_count() or 8, num_results=2)
        for u in range(n_utt):
            got = b.results(u)
            assert len(got) == len(ref[u])
            for r in range(len(got)):
                assert _same(got[r], ref[u][r]), "beam %d utt %d result %d differs" % (beam, u, r)
    finally:
        m.setBeamWidth(500)


# ------------------------------------------------------------------------------------------------ random / adversarial
def _gen_case(kind, rng, T, C, words, utt):
    from stt_b200 import synth
    blank = C - 1
    if kind == "ctc":
        return synth.make_ctc_probs(words, T, utt=utt, noise=float(rng.choice([0.005, 0.02, 0.08])))
    if kind == "dirichlet":      # flat-ish rows: wide beams full of near-equal scores
        return rng.dirichlet(np.ones(C) * float(rng.choice([0.3, 2.0, 50.0])), size=T).astype(np.float32)
    if kind == "two_class":      # one letter and the blank alternate: repeated characters, zombie prefixes (:200-207)
        c = int(rng.integers(1, 27))
        p = np.full((T, C), 1e-4, np.float32)
        for t in range(T):
            p[t, c if rng.random() < 0.6 else blank] = 1.0
            if rng.random() < 0.3:
                p[t, 0] += 0.5   # spaces: LM calls on one-letter "words"
        p *= rng.uniform(0.999, 1.001, size=p.shape).astype(np.float32)   # near-ties, not bit-equal ones (DESIGN tie caveat)
        return (p / p.sum(1, keepdims=True)).astype(np.float32)
    if kind == "quantised":      # probabilities on a coarse grid: many bit-equal class log-probs
        raw = rng.integers(1, 5, size=(T, C)).astype(np.float32)
        raw[np.arange(T), rng.integers(0, C, size=T)] += float(rng.choice([4.0, 16.0]))
        raw *= rng.uniform(0.9995, 1.0005, size=raw.shape).astype(np.float32)   # near-ties: scores 5e-4 apart
        return (raw / raw.sum(1, keepdims=True)).astype(np.float32)
    if kind == "flicker":        # the best path flips between two word hypotheses: prune + revive under old ids
        a, b2 = words[int(rng.integers(len(words)))], words[int(rng.integers(len(words)))]
        lab = {l: i for i, l in enumerate(synth.ENGLISH_LABELS)}
        p = rng.gamma(0.3, 1.0, size=(T, C)).astype(np.float32) * 0.02
        seq_a = [lab[ch] for ch in (a + " ") * T if ch in lab]
        seq_b = [lab[ch] for ch in (b2 + " ") * T if ch in lab]
        for t in range(T):
            if t % 2:
                p[t, blank] += 0.7
            else:
                p[t, seq_a[t // 2]] += 0.45
                p[t, seq_b[t // 2]] += 0.44
        return (p / p.sum(1, keepdims=True)).astype(np.float32)
    raise ValueError(kind)


@pytest.mark.parametrize("beam,scorer", [(1, True), (3, True), (25, True), (100, True), (500, True), (500, False),
                                         (16, False), (2000, True)])
def test_decoder_random_adversarial_matrix(ref_decoder, small_model, vocab_words, english, beam, scorer):
    """128 cases per (beam, scorer) point = 1024 decodes, mixed lengths 1..96, five generators."""
    from stt_b200 import Model
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha) if scorer else None
    path, _ = small_model
    m = Model(path)
    m.setBeamWidth(beam)
    if scorer:
        m.enableExternalScorer(SCORER)
    n_cases, T_max, C = (128, 96, 29) if beam <= 512 else (24, 64, 29)
    rng = np.random.default_rng(90000 + beam * 2 + int(scorer))
    kinds = ["ctc", "dirichlet", "two_class", "quantised", "flicker"]
    probs = np.zeros((n_cases, T_max, C), np.float32)
    lens = []
    for i in range(n_cases):
        T = int(rng.integers(1, T_max + 1))
        probs[i, :T] = _gen_case(kinds[i % len(kinds)], rng, T, C, vocab_words, 5000 + i)
        lens.append(T)
    b = m.createBatch(n_cases, T_max * 320)
    b.set_probs(probs, lens)
    b.decode(num_results=3)
    b.fetch()
    ref = o.ref_decode_batch(probs.astype(np.float64), lens, alpha, beam, sc, num_processes=os.cpu_count() or 8,
                             num_results=3)
    bad = []
    for i in range(n_cases):
        got = b.results(i)
        ok = len(got) == len(ref[i]) and all(_same(g, r) for g, r in zip(got, ref[i]))
        if not ok:
            bad.append((i, kinds[i % len(kinds)], lens[i]))
    assert not bad, "beam %d scorer %s: %d/%d cases differ from the reference, first: %s" % (beam, scorer, len(bad), n_cases, bad[:8])
