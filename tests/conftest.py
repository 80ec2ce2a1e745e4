"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only build box; `-m gpu` runs on a B200 through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")
SCORER = os.path.join(GOLDEN, "pruned_lm.scorer")  # data/smoke_test/pruned_lm.scorer of the reference (fixture data)
VOCAB = os.path.join(GOLDEN, "vocab.pruned.txt")
# data/smoke_test/LDC93S1_pcms16le_1_16000.wav of the reference: the utterance of BASELINE.json configs[0]
LDC93S1_WAV = os.path.join(GOLDEN, "LDC93S1_pcms16le_1_16000.wav")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def vocab_words():
    return open(VOCAB).read().split()


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    if not o.have_port():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    return o


@pytest.fixture(scope="session")
def ref_decoder(oracle):
    if not oracle.have_ref():
        if os.path.isdir("/root/reference"):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-j8", "ref"])
        else:
            pytest.skip("oracle/_ref/libref_decoder.so not built and /root/reference absent")
    return oracle


@pytest.fixture(scope="session")
def english():
    from stt_b200 import synth
    return synth.ENGLISH_LABELS


@pytest.fixture(scope="session")
def small_model(tmp_path_factory):
    """n_hidden=256 model file (fast oracle), seed 1234."""
    from stt_b200 import synth
    w = synth.make_weights(n_hidden=256, seed=1234)
    p = tmp_path_factory.mktemp("model") / "small.sttw"
    synth.write_model(str(p), w)
    return str(p), w


@pytest.fixture(scope="session")
def ldc93s1_pcm():
    """int16 samples of the reference's smoke-test utterance (46 797 samples @ 16 kHz -> 146 timesteps)."""
    import wave
    with wave.open(LDC93S1_WAV, "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate()) == (1, 2, 16000)
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype="<i2").copy()
    assert pcm.size == 46797
    return pcm
