"""The KenLM model types a .scorer can carry (the four trie layouts TRIE / QUANT_TRIE / ARRAY_TRIE / QUANT_ARRAY_TRIE and the
probing-hash model PROBING,
kenlm/lm/model_type.hh:8-20), order 5, built from kenlm's own lm/test.arpa by tests/golden/make_lm_variants.py:
the scorer view (the code the CUDA decoder compiles, here compiled for the host) must equal the compiled reference
(Scorer::get_log_cond_prob) bit for bit on every layout, and reproduce kenlm's model_test.cc known answers on the
unquantised ones."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

VARIANTS = ["trie", "quant_trie", "array_trie", "quant_array_trie", "probing"]
LN10_F32 = float(np.float32(0.4342944819))  # NUM_FLT_LOGE, decoder_utils.h:13


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("scv") / "libscorer_check.so")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "native", "scorer_check.cc"),
                           os.path.join(ROOT, "stt_b200", "csrc", "scorer_image.cc")])
    S = ctypes.CDLL(so)
    S.sc_load.restype = ctypes.c_void_p
    S.sc_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    S.sc_log_cond_prob.restype = ctypes.c_double
    S.sc_log_cond_prob.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    S.sc_order.argtypes = [ctypes.c_void_p]
    S.sc_fst_find.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    S.sc_fst_final.argtypes = [ctypes.c_void_p, ctypes.c_int]
    S.sc_fst_start.restype = ctypes.c_long
    S.sc_fst_start.argtypes = [ctypes.c_void_p]
    return S


def _load(S, name, english):
    lab = b"".join(l.encode() + b"\0" for l in english)
    err = ctypes.c_int()
    h = S.sc_load(os.path.join(GOLDEN, "lm_variants", name + ".scorer").encode(), lab, len(english), 0, ctypes.byref(err))
    assert err.value == 0 and h, (name, hex(err.value))
    return h


def _score(S, h, words, bos):
    return S.sc_log_cond_prob(h, b"".join(w.encode() + b"\0" for w in words), len(words), int(bos))


@pytest.mark.parametrize("name", VARIANTS)
def test_variant_matches_reference_bit_for_bit(checker, ref_decoder, english, name):
    S = checker
    h = _load(S, name, english)
    assert S.sc_order(h) == 5
    o = ref_decoder
    sc = o.RefScorer(os.path.join(GOLDEN, "lm_variants", name + ".scorer"), o.RefAlphabet(english))
    vocab = open(os.path.join(GOLDEN, "lm_variants", "vocab.txt")).read().split()
    vocab = [w for w in vocab if not w.startswith("<")]
    rng = np.random.default_rng(17)
    # every sentence prefix the known answers walk, then random windows up to order + 1 words with OOVs mixed in
    cases = []
    for line in ["looking on a little more loin also would consider higher looking",
                 "also would consider higher looking", "would consider higher looking on a little"]:
        ws = line.split()
        for i in range(len(ws)):
            for k in range(1, 7):
                if i + k <= len(ws):
                    cases.append((ws[i:i + k], True))
                    cases.append((ws[i:i + k], False))
    for _ in range(3000):
        k = int(rng.integers(1, 7))
        ws = [vocab[int(rng.integers(len(vocab)))] for _ in range(k)]
        if rng.random() < 0.1:
            ws[int(rng.integers(k))] = "notaword"
        cases.append((ws, bool(rng.integers(0, 2))))
    for ws, bos in cases:
        assert _score(S, h, ws, bos) == sc.log_cond_prob(ws, bos), (name, ws, bos)


@pytest.mark.parametrize("name", ["trie", "array_trie", "probing"])
def test_kenlm_known_answers_on_unquantised_layouts(checker, english, name):
    """kenlm/lm/model_test.cc:69-101 (Starters / Continuation): FullScore along '<s> looking on a little' and
    'also would consider higher looking' -- log10 values, here through get_log_cond_prob's natural-log window."""
    S = checker
    h = _load(S, name, english)
    known = [(["looking"], -0.4846522), (["looking", "on"], -0.348837), (["looking", "on", "a"], -0.0155266),
             (["looking", "on", "a", "little"], -0.00306122)]
    for ws, log10p in known:
        got = _score(S, h, ws, True) * LN10_F32
        assert abs(got - log10p) <= 1e-5 * max(1.0, abs(log10p)), (ws, got, log10p)
    # model_test.cc:108-113: the 5-gram "also would consider higher looking" scores -1 .. -5 step by step (no BOS)
    ws = "also would consider higher looking".split()
    for k, log10p in [(2, -2.0), (3, -3.0), (4, -4.0), (5, -5.0)]:
        got = _score(S, h, ws[:k], False) * LN10_F32
        assert abs(got - log10p) <= 1e-5 * abs(log10p), (ws[:k], got, log10p)


@pytest.mark.parametrize("name", VARIANTS)
def test_variant_dictionary_accepts_the_spellable_vocabulary(checker, english, name):
    S = checker
    h = _load(S, name, english)
    lab = {l: i for i, l in enumerate(english)}
    start = S.sc_fst_start(h)

    def accepts(word):
        st = start
        for ch in word + " ":
            st = S.sc_fst_find(h, st, lab[ch] + 1)
            if st < 0:
                return False
        return bool(S.sc_fst_final(h, st))
    vocab = open(os.path.join(GOLDEN, "lm_variants", "vocab.txt")).read().split()
    for w in vocab:
        if all(c in lab and c != " " for c in w):
            assert accepts(w), w
    assert not accepts("zzz") and not accepts("lookin")
