"""The oracle is pinned before it is trusted:
  * MFCC / spectrogram restatement vs the golden vectors of the vendored TFLite op tests
    (tensorflow/tensorflow/lite/kernels/mfcc_test.cc:67-90, audio_spectrogram_test.cc:64-108, tol 1e-3),
  * stream-buffering frame counts (160000 -> 500, 46797 -> 146, 16000 -> 50 timesteps),
  * C restatement of the acoustic model vs an independent torch fp32 implementation,
  * committed golden fixtures still reproduce (guards against oracle drift)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_mfcc_op_golden_vector(oracle):
    data = np.arange(1, 514, dtype=np.float32)
    out = oracle.mfcc_from_spectrum(data, 22050.0, 20.0, 4000.0, 40, 13)
    gold = [29.13970072, -6.41568601, -0.61903012, -0.96778652, -0.26819878, -0.40907028, -0.15614748, -0.23203119,
            -0.10481487, -0.1543029, -0.0769791, -0.10806114, -0.06047613]
    np.testing.assert_allclose(out, gold, atol=1e-3)


def test_spectrogram_op_golden_vectors(oracle):
    x = np.array([-1, 0, 1, 0, -1, 0, 1, 0], np.float32)
    np.testing.assert_allclose(oracle.spectrogram_frame(x, 8), [0, 1, 4, 1, 0], atol=1e-3)       # SquaredTest
    np.testing.assert_allclose(np.sqrt(oracle.spectrogram_frame(x, 8)), [0, 1, 2, 1, 0], atol=1e-3)  # NonSquaredTest
    y = np.array([1, 0, -1, 0, 1, 0, 1, 0], np.float32)                                          # StrideTest, 2nd frame
    np.testing.assert_allclose(oracle.spectrogram_frame(y, 8), [1, 2, 1, 2, 1], atol=1e-3)


@pytest.mark.parametrize("n,T", [(160000, 500), (46797, 146), (16000, 50), (0, 1), (511, 1), (512, 2), (832, 3)])
def test_stream_frame_counts(oracle, n, T):
    from stt_b200 import synth
    got, mfcc = oracle.features_only(synth.make_pcm(n, utt=1))
    assert got == T == synth.n_timesteps(n)
    assert mfcc.shape == (T, 26)


def test_am_port_matches_torch(oracle):
    from oracle.am_torch import TorchAM
    from stt_b200 import synth
    w = synth.make_weights(n_hidden=128, seed=3)
    pcm = synth.make_pcm(12000, utt=2)
    probs_c, mfcc = oracle.PortAM(w).stream(pcm)
    probs_t = TorchAM(w).forward_features(mfcc)
    assert probs_c.shape == probs_t.shape == (37, 29)
    assert np.abs(probs_c - probs_t).max() < 1e-6


def test_stream_chunking_invariance(oracle):
    """The restated runtime gives the same probabilities for any chunking (asserts.sh:591-604 property)."""
    from stt_b200 import synth
    w = synth.make_weights(n_hidden=64, seed=4)
    pcm = synth.make_pcm(20000, utt=3)
    a, _ = oracle.PortAM(w).stream(pcm)
    b, _ = oracle.PortAM(w).stream(pcm, chunks=[320] * 70)
    np.testing.assert_array_equal(a, b)
    c, _ = oracle.PortAM(w).stream(pcm, chunks=[10000, 10000], flush_at=(0,))  # intermediate flush "trashes" state
    assert c.shape[0] == a.shape[0] + 1  # the flush emits one extra (partial-window) frame
    assert not np.array_equal(a[:c.shape[0] - 1], c[:-1])


def test_committed_mfcc_golden(oracle):
    g = np.load(os.path.join(GOLDEN, "mfcc_golden.npz"))
    T, mfcc = oracle.features_only(g["pcm"])
    assert T == int(g["timesteps"])
    np.testing.assert_array_equal(mfcc, g["mfcc"])
    from stt_b200 import synth
    np.testing.assert_allclose(synth.np_mfcc(g["pcm"]), g["mfcc"], atol=1e-5)


def test_committed_decoder_golden_reproduces(ref_decoder):
    from stt_b200 import synth
    from conftest import SCORER
    o = ref_decoder
    g = np.load(os.path.join(GOLDEN, "decoder_golden.npz"), allow_pickle=True)
    alpha = o.RefAlphabet(synth.ENGLISH_LABELS)
    sc = o.RefScorer(SCORER, alpha)
    for u in range(g["probs"].shape[0]):
        c, tok, ts = o.ref_decode(g["probs"][u], alpha, int(g["beam"]), sc)[0]
        assert list(tok) == list(g["tokens"][u]) and list(ts) == list(g["timesteps"][u]) and c == g["confidence"][u]


def test_kenlm_known_answer(ref_decoder):
    """get_log_cond_prob({"she","had"}, bos) on the smoke scorer."""
    from stt_b200 import synth
    from conftest import SCORER
    o = ref_decoder
    sc = o.RefScorer(SCORER, o.RefAlphabet(synth.ENGLISH_LABELS))
    assert abs(sc.log_cond_prob(["she", "had"], True) - (-0.970070)) < 1e-6


def test_real_utterance_frame_count_and_feature_range(oracle, ldc93s1_pcm):
    """BASELINE configs[0] plumbing on REAL speech (the reference's LDC93S1 clip): 46 797 samples -> 146 timesteps
    (SURVEY 8c known answer); features are finite and span the dynamic range speech has and synthetic tones do not."""
    T, feats = oracle.features_only(ldc93s1_pcm)
    assert T == 146 and feats.shape == (146, 26)
    assert np.isfinite(feats).all()
    assert feats[:, 0].max() - feats[:, 0].min() > 5.0   # C0 follows the energy contour of the utterance
