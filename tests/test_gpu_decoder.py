"""GPU beam search (stt_b200/csrc/decoder.cuh) vs the GENUINE reference decoder (oracle/_ref/libref_decoder.so) on
identical float32 probabilities: token ids, timesteps and confidence of the top results must be identical."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCORER

pytestmark = pytest.mark.gpu


def _model(small_model, beam, scorer=True):
    from stt_b200 import Model
    path, _ = small_model
    m = Model(path)
    m.setBeamWidth(beam)
    if scorer:
        m.enableExternalScorer(SCORER)
    return m


def _compare(gpu_results, ref_results, what, conf_exact=True):
    assert len(gpu_results) == len(ref_results), what
    for r, ((gc, gt, gts), (rc, rt, rts)) in enumerate(zip(gpu_results, ref_results)):
        assert list(gt) == list(rt), "%s: tokens of result %d differ" % (what, r)
        assert list(gts) == list(rts), "%s: timesteps of result %d differ" % (what, r)
        if conf_exact:
            assert gc == rc, "%s: confidence of result %d differs: %r vs %r" % (what, r, gc, rc)
        else:
            assert abs(gc - rc) <= 1e-4 * max(1.0, abs(rc))


@pytest.mark.parametrize("beam", [1, 8, 100, 500])
@pytest.mark.parametrize("T", [50, 200])
def test_decoder_with_scorer_matches_reference(ref_decoder, small_model, vocab_words, english, beam, T):
    from stt_b200 import synth
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    m = _model(small_model, beam)
    B = 6
    probs = np.stack([synth.make_ctc_probs(vocab_words, T, utt=100 * beam + u) for u in range(B)])
    b = m.createBatch(B, T * 320)
    b.set_probs(probs, [T] * B)
    b.decode(num_results=3)
    b.fetch()
    for u in range(B):
        ref = o.ref_decode(probs[u], alpha, beam, sc, num_results=3)
        _compare(b.results(u), ref, "beam=%d T=%d utt=%d" % (beam, T, u))


@pytest.mark.parametrize("beam", [700, 2000])
def test_wide_beam_matches_reference(ref_decoder, small_model, vocab_words, english, beam):
    """Beam widths above 512 (BASELINE config 5 sweeps to 2000) run the wide instantiation of the step kernel: one CTA
    per SM, candidates and the word-ordinal / child-mask arrays in global memory."""
    from stt_b200 import synth
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    m = _model(small_model, beam)
    B, T = 3, 60
    probs = np.stack([synth.make_ctc_probs(vocab_words, T, utt=7000 + beam + u) for u in range(B)])
    b = m.createBatch(B, T * 320)
    b.set_probs(probs, [T] * B)
    b.decode(num_results=2)
    b.fetch()
    for u in range(B):
        ref = o.ref_decode(probs[u], alpha, beam, sc, num_results=2)
        _compare(b.results(u), ref, "wide beam=%d utt=%d" % (beam, u))


@pytest.mark.parametrize("beam", [4, 64])
def test_decoder_without_scorer_matches_reference(ref_decoder, small_model, vocab_words, english, beam):
    from stt_b200 import synth
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    m = _model(small_model, beam, scorer=False)
    T, B = 60, 4
    probs = np.stack([synth.make_ctc_probs(vocab_words, T, utt=900 + u) for u in range(B)])
    b = m.createBatch(B, T * 320)
    b.set_probs(probs, [T] * B)
    b.decode(num_results=2)
    b.fetch()
    for u in range(B):
        ref = o.ref_decode(probs[u], alpha, beam, None, num_results=2)
        _compare(b.results(u), ref, "noscorer beam=%d utt=%d" % (beam, u))


def test_decoder_edge_cases(ref_decoder, small_model, vocab_words, english):
    """All-blank input (gate never opens), near-uniform softmax, ragged lengths, single frame."""
    from stt_b200 import synth
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    beam = 50
    m = _model(small_model, beam)
    T = 40
    C = 29
    rng = np.random.default_rng(5)
    blank = np.full((T, C), 1e-5, np.float32)
    blank[:, 28] = 1.0
    blank /= blank.sum(1, keepdims=True)
    uniform = rng.dirichlet(np.ones(C) * 50, size=T).astype(np.float32)
    normal = synth.make_ctc_probs(vocab_words, T, utt=77)
    probs = np.stack([blank, uniform, normal, normal])
    lens = [T, T, 1, 17]
    b = m.createBatch(4, T * 320)
    b.set_probs(probs, lens)
    b.decode(num_results=2)
    b.fetch()
    for u in range(4):
        ref = o.ref_decode(probs[u][:lens[u]], alpha, beam, sc, num_results=2)
        _compare(b.results(u), ref, "edge utt=%d" % u)


def test_decoder_alpha_beta_sweep(ref_decoder, small_model, vocab_words, english):
    """config 5 of BASELINE.json in miniature: lm_alpha / lm_beta grid, parity at every point."""
    from stt_b200 import synth
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    m = _model(small_model, 100)
    T = 120
    probs = np.stack([synth.make_ctc_probs(vocab_words, T, utt=4242)])
    for a in (0.5, 0.931289039105002, 1.5):
        for be in (0.5, 1.1834137581510284, 2.0):
            sc.set_alpha_beta(a, be)
            m.setScorerAlphaBeta(a, be)
            b = m.createBatch(1, T * 320)
            b.set_probs(probs, [T])
            b.decode(num_results=1)
            b.fetch()
            _compare(b.results(0), o.ref_decode(probs[0], alpha, 100, sc), "alpha=%g beta=%g" % (a, be))


def test_decoder_hot_words_match_reference(ref_decoder, small_model, vocab_words, english):
    """ctc_beam_search_decoder.cpp:224-239: every window word that is a hot word adds its boost before the alpha scale."""
    import ctypes
    from stt_b200 import synth
    o = ref_decoder
    R = o.ref()
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    beam, T = 100, 150
    hot = {"the": 7.5, "and": -4.0, "of": 3.25, "notaword": 9.0, vocab_words[10]: 6.0, vocab_words[200]: -2.5}
    m = _model(small_model, beam)
    for w, b in hot.items():
        m.addHotWord(w, b)
    for u in range(4):
        probs = synth.make_ctc_probs(vocab_words, T, utt=5150 + u)
        bt = m.createBatch(1, T * 320)
        bt.set_probs(probs[None], [T])
        bt.decode(num_results=2)
        bt.fetch()
        words = b"".join(w.encode() + b"\0" for w in hot)
        boosts = np.array(list(hot.values()), np.float32)
        d = R.ref_decoder_new(alpha.h, beam, 1.0, 40, sc.h, words, boosts.ctypes.data, len(hot))
        p64 = np.ascontiguousarray(probs, np.float64)
        R.ref_decoder_next(d, p64.ctypes.data, T, 29)
        conf = np.zeros(2, np.float64); nt = np.zeros(2, np.int32)
        tok = np.zeros((2, T), np.uint32); ts = np.zeros((2, T), np.uint32)
        n = R.ref_decoder_decode(d, 2, T, conf.ctypes.data, nt.ctypes.data, tok.ctypes.data, ts.ctypes.data)
        R.ref_decoder_free(d)
        ref = [(conf[r], tok[r, :nt[r]], ts[r, :nt[r]]) for r in range(n)]
        _compare(bt.results(0), ref, "hot words utt=%d" % u)
    # and a hot word that occurs in the transcript must change its confidence
    m2 = _model(small_model, beam)
    probs = synth.make_ctc_probs(vocab_words, T, utt=5150)
    b2 = m2.createBatch(1, T * 320)
    b2.set_probs(probs[None], [T]); b2.decode(1); b2.fetch()
    base_conf, base_tok, _ = b2.results(0)[0]
    first_word = alpha.decode(base_tok).split()[0]
    m2.addHotWord(first_word, 3.0)
    b3 = m2.createBatch(1, T * 320)
    b3.set_probs(probs[None], [T]); b3.decode(1); b3.fetch()
    assert b3.results(0)[0][0] > base_conf


def test_decoder_against_committed_golden(small_model):
    """Golden vectors produced by the reference decoder in the build container (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLDEN, "decoder_golden.npz"), allow_pickle=True)
    m = _model(small_model, int(g["beam"]))
    probs = g["probs"]
    B, T, _ = probs.shape
    b = m.createBatch(B, T * 320)
    b.set_probs(probs, [T] * B)
    b.decode(num_results=1)
    b.fetch()
    for u in range(B):
        conf, tok, ts = b.results(u)[0]
        assert list(tok) == list(g["tokens"][u])
        assert list(ts) == list(g["timesteps"][u])
        assert conf == float(g["confidence"][u])


@pytest.mark.parametrize("layout", ["trie", "quant_trie", "array_trie", "quant_array_trie", "probing"])
def test_decoder_with_every_kenlm_trie_layout(ref_decoder, small_model, english, layout):
    """Order-5 scorers in the four trie layouts (tests/golden/make_lm_variants.py, from kenlm's lm/test.arpa): the GPU
    decoder -- word ordinals, carried KenLM states, interpolation searches, Bhiksha next pointers -- against the
    reference decoder with the same package."""
    import os
    from conftest import GOLDEN
    from stt_b200 import Model, synth
    o = ref_decoder
    pkg = os.path.join(GOLDEN, "lm_variants", layout + ".scorer")
    vocab = [w for w in open(os.path.join(GOLDEN, "lm_variants", "vocab.txt")).read().split()
             if all(c in english and c != " " for c in w)]
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(pkg, alpha)
    path, _ = small_model
    m = Model(path)
    m.setBeamWidth(128)
    m.enableExternalScorer(pkg)
    B, T = 4, 120
    probs = np.stack([synth.make_ctc_probs(vocab, T, utt=4242 + u) for u in range(B)])
    b = m.createBatch(B, T * 320)
    b.set_probs(probs, [T] * B)
    b.decode(num_results=3)
    b.fetch()
    for u in range(B):
        ref = o.ref_decode(probs[u], alpha, 128, sc, num_results=3)
        _compare(b.results(u), ref, "layout=%s utt=%d" % (layout, u))


@pytest.mark.parametrize("order,kind,seed", [(2, "trie", 1), (2, "probing", 2), (3, "quant_array_trie", 3), (3, "array_trie", 4),
                                             (6, "quant_array_trie", 5), (6, "probing", 6), (6, "quant_trie", 7),
                                             (4, "probing", 8)])
def test_decoder_with_random_lms_of_every_order(ref_decoder, small_model, english, tmp_path, order, kind, seed):
    """Randomised KenLM models of order 2, 3 and 6 (KENLM_MAX_ORDER; the committed fixtures are orders 4 and 5), built on
    the spot with the reference's build_binary (oracle/_ref, test_scorer_fuzz.py's generator) and packaged by the
    reference's Scorer: the carried KenLM state holds order-1 words, so the state hand-over between a prefix's words is
    exercised at its shortest and longest -- shared-memory kernel and, with vocabulary pruning, the general kernel."""
    import subprocess
    from test_scorer_fuzz import BUILD_BINARY, _flags, _random_arpa
    from stt_b200 import Model, synth
    if not os.path.exists(BUILD_BINARY):
        pytest.skip("oracle/_ref/build_binary not built")
    o = ref_decoder
    rng = np.random.default_rng(seed)
    arpa = str(tmp_path / "lm.arpa")
    words, got_order = _random_arpa(rng, arpa, order, 60, 500, 0.25 if seed % 2 else 0.0)
    assert got_order == order
    typ, flags = _flags(rng, kind)
    lm = str(tmp_path / "lm.binary")
    subprocess.check_call([BUILD_BINARY] + flags + ["-v"] + typ + [arpa, lm], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    alpha = o.RefAlphabet(english)
    pkg = str(tmp_path / "lm.scorer")
    assert o.ref().ref_make_scorer_package(lm.encode(), b"".join(w.encode() + b"\0" for w in words), len(words),
                                           alpha.h, pkg.encode(), 0.8, 1.4) == 0
    sc = o.RefScorer(pkg, alpha)
    path, _ = small_model
    m = Model(path)
    m.setBeamWidth(96)
    m.enableExternalScorer(pkg)
    B, T = 4, 140
    probs = np.stack([synth.make_ctc_probs(words, T, utt=7300 + 10 * seed + u) for u in range(B)])
    b = m.createBatch(B, T * 320)
    b.set_probs(probs, [T] * B)
    b.decode(num_results=3)
    b.fetch()
    for u in range(B):
        ref = o.ref_decode(probs[u], alpha, 96, sc, num_results=3)
        _compare(b.results(u), ref, "order=%d %s utt=%d" % (order, kind, u))
    b2 = m.createBatch(B, T * 320)
    b2.set_probs(probs, [T] * B)
    b2.set_cutoff(0.995, 12)
    b2.decode(num_results=2)
    b2.fetch()
    for u in range(B):
        ref = o.ref_decode(probs[u], alpha, 96, sc, num_results=2, cutoff_prob=0.995, cutoff_top_n=12)
        for (gc, gt, gts), (rc, rt, rts) in zip(b2.results(u), ref):
            assert list(gt) == list(rt) and list(gts) == list(rts), "pruned order=%d %s utt=%d" % (order, kind, u)
            assert gc == rc or rc < -1e30
