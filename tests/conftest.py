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
