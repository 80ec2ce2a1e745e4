"""The reference's Python decoder surface (native_client/ctcdecode/__init__.py) on the GPU decoder: f64 inputs,
batch + single entry points, alpha/beta sweep (BASELINE config 5 in miniature) -- identical to the reference."""
import numpy as np
import pytest

from conftest import SCORER

pytestmark = pytest.mark.gpu


def test_ctcdecoder_batch_matches_reference(ref_decoder, vocab_words, english):
    from stt_b200 import ctcdecoder, synth
    o = ref_decoder
    alpha = ctcdecoder.Alphabet()
    alpha.InitFromLabels(english)
    assert alpha.Decode(alpha.Encode("hello world")) == "hello world" and alpha.GetSize() == 28
    ra = o.RefAlphabet(english)
    rs = o.RefScorer(SCORER, ra)
    B, T = 5, 90
    rng = np.random.default_rng(3)
    probs = np.stack([synth.make_ctc_probs(vocab_words, T, utt=8000 + u) for u in range(B)]).astype(np.float64)
    probs += rng.uniform(0, 1e-9, size=probs.shape)   # genuinely f64 inputs (not float-representable)
    lens = [T, 60, T, 1, 33]
    for a, be, beam in ((0.75, 1.85, 64), (0.931289039105002, 1.1834137581510284, 200)):
        sc = ctcdecoder.Scorer(a, be, SCORER, alpha)
        rs.set_alpha_beta(a, be)
        got = ctcdecoder.ctc_beam_search_decoder_batch(probs, lens, alpha, beam, num_processes=4, scorer=sc, num_results=2)
        ref = o.ref_decode_batch(probs, lens, ra, beam, rs, num_processes=2, num_results=2)
        assert len(got) == B
        for u in range(B):
            assert len(got[u]) == len(ref[u])
            for g, (rc, rt, rts) in zip(got[u], ref[u]):
                assert g.tokens == list(rt) and g.timesteps == list(rts) and g.confidence == rc
                assert g.transcript == ra.decode(rt)
    one = ctcdecoder.ctc_beam_search_decoder(probs[0], alpha, 64, scorer=ctcdecoder.Scorer(0.75, 1.85, SCORER, alpha))
    rs.set_alpha_beta(0.75, 1.85)
    rc, rt, rts = o.ref_decode(probs[0], ra, 64, rs)[0]
    assert one[0].tokens == list(rt) and one[0].confidence == rc
    # vocabulary pruning runs the general kernel
    pruned = ctcdecoder.ctc_beam_search_decoder(probs[0], alpha, 8, cutoff_prob=0.99, cutoff_top_n=5, scorer=ctcdecoder.Scorer(0.75, 1.85, SCORER, alpha))
    rc, rt, rts = o.ref_decode(probs[0], ra, 8, rs, cutoff_prob=0.99, cutoff_top_n=5)[0]
    assert pruned[0].tokens == list(rt) and pruned[0].timesteps == list(rts) and pruned[0].confidence == rc
