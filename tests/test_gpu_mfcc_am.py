"""MFCC kernel and the whole acoustic forward (GPU, through STTX_Batch*) vs the CPU oracle restatement."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# Stated tolerances (BASELINE.json north_star: "logits within a stated fp tolerance"):
MFCC_ATOL = 2e-4      # fp32 MFCC vs the double-precision oracle (device log/cos differ from glibc in the last ulp)
PROBS_ATOL = 2e-3     # softmax probabilities, fp16-operand tensor-core path vs fp32 oracle (SURVEY 8d parity gate 2)


@pytest.mark.parametrize("n_samples", [160000, 46797, 16000, 511, 512, 513, 832, 0, 1])
def test_mfcc_matches_oracle(oracle, small_model, n_samples):
    from stt_b200 import Model, synth
    path, _ = small_model
    m = Model(path)
    pcm = synth.make_pcm(n_samples, utt=3)
    b = m.createBatch(1, max(n_samples, 512))
    b.upload([pcm])
    b.forward()
    T_ref, mfcc_ref = oracle.features_only(pcm)
    assert b.timesteps(0) == T_ref == synth.n_timesteps(n_samples)
    got = b.features(0)
    assert got.shape == mfcc_ref.shape
    np.testing.assert_allclose(got, mfcc_ref, atol=MFCC_ATOL, rtol=1e-5)


def test_mfcc_ragged_batch(oracle, small_model):
    from stt_b200 import Model, synth
    path, _ = small_model
    m = Model(path)
    lens = [16000, 4000, 12345, 512, 7]
    pcms = [synth.make_pcm(n, utt=i) for i, n in enumerate(lens)]
    b = m.createBatch(len(lens), max(lens))
    b.upload(pcms)
    b.forward()
    for i, p in enumerate(pcms):
        _, ref = oracle.features_only(p)
        np.testing.assert_allclose(b.features(i), ref, atol=MFCC_ATOL, rtol=1e-5)


@pytest.mark.parametrize("n_samples", [16000, 46797])
def test_am_probs_small_model(oracle, small_model, n_samples):
    from stt_b200 import Model, synth
    path, w = small_model
    m = Model(path)
    pcm = synth.make_pcm(n_samples, utt=1)
    b = m.createBatch(1, n_samples)
    b.upload([pcm])
    b.forward()
    probs_ref, _ = oracle.PortAM(w).stream(pcm)
    got = b.probs(0)
    assert got.shape == probs_ref.shape
    assert np.abs(got - probs_ref).max() <= PROBS_ATOL
    np.testing.assert_allclose(got.sum(1), 1.0, rtol=1e-4)


def test_am_probs_batch_equals_single(small_model):
    """Utterances are independent: batched rows must equal the batch-of-one result bit for bit."""
    from stt_b200 import Model, synth
    path, _ = small_model
    m = Model(path)
    lens = [16000, 9000, 16000, 3200]
    pcms = [synth.make_pcm(n, utt=10 + i) for i, n in enumerate(lens)]
    b = m.createBatch(4, 16000)
    b.upload(pcms)
    b.forward()
    batched = [b.probs(i) for i in range(4)]
    for i, p in enumerate(pcms):
        b1 = m.createBatch(1, 16000)
        b1.upload([p])
        b1.forward()
        np.testing.assert_array_equal(batched[i], b1.probs(0))


def test_am_probs_full_size_model(oracle, tmp_path):
    """n_hidden = 2048 (English v1.x geometry), 1 s of audio: GPU vs the fp32 oracle."""
    from stt_b200 import Model, synth
    w = synth.make_weights(n_hidden=2048, seed=1234)
    path = str(tmp_path / "full.sttw")
    synth.write_model(path, w)
    m = Model(path)
    pcm = synth.make_pcm(16000, utt=5)
    b = m.createBatch(1, 16000)
    b.upload([pcm])
    b.forward()
    probs_ref, _ = oracle.PortAM(w).stream(pcm)
    got = b.probs(0)
    err = np.abs(got - probs_ref).max()
    print("full-size model max |dp| = %.3e" % err)
    assert err <= PROBS_ATOL


def test_am_large_batch_uses_pair_kernel_and_matches_single(small_model):
    """129..256 utterances run the cta_group::2 LSTM kernel: every row must equal the batch-of-one result (<= 128
    utterances, single-CTA kernel) bit for bit -- same fp16 operands, same fp32 accumulation order along K."""
    from stt_b200 import Model, synth
    path, _ = small_model
    m = Model(path)
    B = 131
    lens = [8000 if i % 3 else 6000 + 37 * i for i in range(B)]
    pcms = [synth.make_pcm(n, utt=300 + i) for i, n in enumerate(lens)]
    b = m.createBatch(B, max(lens))
    b.upload(pcms)
    b.forward()
    b1 = m.createBatch(1, max(lens))
    for i in (0, 1, 2, 64, 127, 128, 129, 130):
        b1.upload([pcms[i]])
        b1.forward()
        np.testing.assert_array_equal(b.probs(i), b1.probs(0))


def test_real_speech_features_and_probs(oracle, small_model, ldc93s1_pcm):
    """BASELINE configs[0] on the GPU: the reference's LDC93S1 recording (real speech: silences, plosives, 60 dB of
    dynamic range) through MFCC and the acoustic model, against the oracle."""
    from stt_b200 import Model
    path, w = small_model
    m = Model(path)
    b = m.createBatch(1, ldc93s1_pcm.size)
    b.upload([ldc93s1_pcm])
    b.forward()
    T_ref, mfcc_ref = oracle.features_only(ldc93s1_pcm)
    assert b.timesteps(0) == T_ref == 146
    np.testing.assert_allclose(b.features(0), mfcc_ref, atol=MFCC_ATOL, rtol=1e-5)
    probs_ref, _ = oracle.PortAM(w).stream(ldc93s1_pcm)
    assert np.abs(b.probs(0) - probs_ref).max() <= PROBS_ATOL
