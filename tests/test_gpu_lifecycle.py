"""Lifetime semantics of the reference API that round 1 got wrong or capped:
   * streams are unbounded in length: the decoder arena is garbage-collected between chunks like PathTrie::remove bounds
     the reference's trie (path_trie.cpp:192-209) -- a 60 s stream through a 12 s arena equals the offline decode bit
     for bit;
   * a live stream keeps the scorer it was created with (stt.cc:542-547 hands the shared_ptr<Scorer> to DecoderState::init):
     enabling another scorer, disabling the scorer, or a FAILED enable in mid-stream changes nothing for it, and a failed
     enable leaves the old scorer enabled (stt.cc:428-432);
   * STT_FreeModel before STT_FreeStream is legal (the reference's STT_FreeStream never touches the model)."""
import ctypes
import os

import numpy as np
import pytest

from conftest import GOLDEN, SCORER

pytestmark = pytest.mark.gpu
OTHER_SCORER = os.path.join(GOLDEN, "lm_variants", "trie.scorer")


@pytest.fixture(scope="module")
def ctc_model(tmp_path_factory):
    from stt_b200 import synth
    w = synth.bench_weights(n_hidden=256, seed=1234)
    p = tmp_path_factory.mktemp("lifecycle") / "ctc.sttw"
    synth.write_model(str(p), w, beam_width=100)
    return str(p), w


def test_long_stream_is_garbage_collected_and_exact(ref_decoder, ctc_model, english, monkeypatch):
    from stt_b200 import Model, synth
    monkeypatch.setenv("STT_B200_STREAM_ARENA_SECONDS", "12")
    path, _ = ctc_model
    m = Model(path)
    m.enableExternalScorer(SCORER)
    n = 60 * 16000
    pcm = np.concatenate([synth.make_pcm(160000, utt=900 + k) for k in range(6)])
    assert pcm.size == n
    s = m.createStream()
    rng = np.random.default_rng(3)
    pos = 0
    while pos < n:
        c = int(rng.integers(1000, 48000))
        s.feedAudioContent(pcm[pos:pos + c])
        pos += c
    gcs = s.arenaCompactions()
    md = s.finishStreamWithMetadata(2)
    assert gcs >= 3, "a 3000-timestep stream must have outgrown a 600-timestep arena several times (%d)" % gcs
    # offline decode of the same audio: arena sized for the whole utterance, never collected
    b = m.createBatch(1, n)
    b.upload([pcm])
    b.forward()
    b.decode(2)
    b.fetch()
    got = [(t.confidence, [x.text for x in t.tokens], [x.timestep for x in t.tokens]) for t in md.transcripts]
    off = b.results(0)
    assert len(got) == len(off) == 2
    for (gc, gtok, gts), (oc, otok, ots) in zip(got, off):
        assert "".join(gtok) == "".join(english[i] for i in otok)
        assert gts == list(ots) and gc == oc
    assert len(got[0][1]) > 300, "the stream should have produced a long transcript"
    # and the genuine reference decoder on the same probabilities
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    rc, rt, rts = o.ref_decode(b.probs(0), alpha, 100, sc)[0]
    assert list(rt) == list(off[0][1]) and list(rts) == list(off[0][2]) and rc == off[0][0]


def test_stream_keeps_its_scorer(ctc_model, tmp_path):
    from stt_b200 import Model, synth
    from stt_b200.api import STTError
    path, _ = ctc_model
    pcm = synth.make_pcm(96000, utt=77)

    def run(meddle):
        m = Model(path)
        m.enableExternalScorer(SCORER)
        s = m.createStream()
        s.feedAudioContent(pcm[:40000])
        mid = s.intermediateDecode()
        meddle(m)
        s.feedAudioContent(pcm[40000:])
        return mid, s.finishStream(), m

    base_mid, base, _ = run(lambda m: None)
    assert len(base) > 20

    def other(m):
        m.enableExternalScorer(OTHER_SCORER)

    def disable(m):
        m.disableExternalScorer()

    def broken(m):
        bad = tmp_path / "broken.scorer"
        bad.write_bytes(open(SCORER, "rb").read()[:5000])
        with pytest.raises(STTError):
            m.enableExternalScorer(str(bad))

    for meddle in (other, disable, broken):
        mid, text, m = run(meddle)
        assert (mid, text) == (base_mid, base), meddle.__name__
        if meddle is broken:
            assert m.stt(pcm) == base                      # the old scorer is still the enabled one
        if meddle is disable:
            # the reference binding returns the code instead of raising (native_client/python/__init__.py:161-174)
            assert m.setScorerAlphaBeta(1.0, 1.0) == 0x2004   # STT_ERR_SCORER_NOT_ENABLED (coqui-stt.h)
            assert m.stt(pcm) != base                      # new streams decode without a scorer
        if meddle is other:
            assert m.stt(pcm) != base                      # new streams use the new scorer


def test_alpha_beta_reach_live_streams(ctc_model):
    """alpha / beta live in the shared Scorer object (scorer.cpp:346-351): changing them affects streams already running."""
    from stt_b200 import Model, synth
    path, _ = ctc_model
    pcm = synth.make_pcm(96000, utt=78)
    m = Model(path)
    m.enableExternalScorer(SCORER)
    plain = m.stt(pcm)
    s = m.createStream()
    m.setScorerAlphaBeta(3.5, 0.1)
    s.feedAudioContent(pcm)
    changed = s.finishStream()
    m2 = Model(path)
    m2.enableExternalScorer(SCORER)
    m2.setScorerAlphaBeta(3.5, 0.1)
    assert changed == m2.stt(pcm)
    assert changed != plain


def test_free_model_before_stream(ctc_model):
    from stt_b200 import api
    L = api.lib()
    path, _ = ctc_model
    mh = ctypes.c_void_p()
    assert L.STT_CreateModel(path.encode(), ctypes.byref(mh)) == 0
    sh = ctypes.c_void_p()
    assert L.STT_CreateStream(mh, ctypes.byref(sh)) == 0
    pcm = np.zeros(16000, np.int16)
    L.STT_FeedAudioContent(sh, pcm.ctypes.data, pcm.size)
    L.STT_FreeModel(mh)
    L.STT_FeedAudioContent(sh, pcm.ctypes.data, pcm.size)   # ignored, must not crash
    assert not L.STT_IntermediateDecode(sh)
    L.STT_FreeStream(sh)
