"""Host-compiled copy of the device scorer view (stt_b200/csrc/scorer_view.h + scorer_image.cc: KenLM quantised
array trie + ConstFst, the same code the CUDA decoder compiles) vs the compiled reference (Scorer::get_log_cond_prob,
scorer.cpp:301-344) and vs the committed golden values: bit-exact doubles."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, SCORER


@pytest.fixture(scope="module")
def view(tmp_path_factory, english):
    so = str(tmp_path_factory.mktemp("sc") / "libscorer_check.so")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "native", "scorer_check.cc"),
                           os.path.join(ROOT, "stt_b200", "csrc", "scorer_image.cc")])
    S = ctypes.CDLL(so)
    S.sc_load.restype = ctypes.c_void_p
    S.sc_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    S.sc_log_cond_prob.restype = ctypes.c_double
    S.sc_log_cond_prob.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]
    S.sc_alpha.restype = ctypes.c_double
    S.sc_alpha.argtypes = [ctypes.c_void_p]
    S.sc_beta.restype = ctypes.c_double
    S.sc_beta.argtypes = [ctypes.c_void_p]
    S.sc_order.argtypes = [ctypes.c_void_p]
    S.sc_fst_find.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    S.sc_fst_final.argtypes = [ctypes.c_void_p, ctypes.c_int]
    S.sc_fst_start.restype = ctypes.c_long
    S.sc_fst_start.argtypes = [ctypes.c_void_p]
    S.sc_vocab_index.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    S.sc_vocab_index.restype = ctypes.c_uint
    lab = b"".join(l.encode() + b"\0" for l in english)
    err = ctypes.c_int()
    h = S.sc_load(SCORER.encode(), lab, len(english), 0, ctypes.byref(err))
    assert err.value == 0 and h
    return S, h


def _score(S, h, words, bos):
    buf = b"".join(w.encode() + b"\0" for w in words)
    return S.sc_log_cond_prob(h, buf, len(words), int(bos))


def test_header_fields(view):
    S, h = view
    assert S.sc_order(h) == 4
    assert S.sc_alpha(h) == 0.75 and abs(S.sc_beta(h) - 1.85) < 1e-6


def test_lm_golden(view):
    S, h = view
    g = np.load(os.path.join(GOLDEN, "lm_golden.npz"))
    for gram, bos, val in zip(g["grams"], g["bos"], g["vals"]):
        assert _score(S, h, str(gram).split(), bos) == val, gram


def test_lm_matches_reference_live(view, ref_decoder, vocab_words, english):
    S, h = view
    o = ref_decoder
    sc = o.RefScorer(SCORER, o.RefAlphabet(english))
    rng = np.random.default_rng(5)
    for _ in range(5000):
        k = int(rng.integers(1, 5))
        ws = [vocab_words[int(rng.integers(len(vocab_words)))] for _ in range(k)]
        if rng.random() < 0.03:
            ws[int(rng.integers(k))] = "notaword"
        bos = bool(rng.integers(0, 2))
        assert _score(S, h, ws, bos) == sc.log_cond_prob(ws, bos), ws


def test_dictionary_fst_accepts_exactly_the_vocabulary(view, vocab_words, english):
    S, h = view
    lab = {l: i for i, l in enumerate(english)}
    start = S.sc_fst_start(h)

    def accepts(word):
        st = start
        for ch in word + " ":
            st = S.sc_fst_find(h, st, lab[ch] + 1)
            if st < 0:
                return False
        return bool(S.sc_fst_final(h, st))
    for w in vocab_words[:500]:
        assert accepts(w), w
    for w in ["zzzz", "qx", "thee" + "q"]:
        assert not accepts(w)
    assert S.sc_vocab_index(h, b"the") > 0 and S.sc_vocab_index(h, b"notaword") == 0


def test_invalid_scorers_rejected(view, tmp_path, english):
    S, _ = view
    lab = b"".join(l.encode() + b"\0" for l in english)
    data = open(SCORER, "rb").read()
    cases = {"garbage": b"hello world" * 100, "truncated_lm": data[:5000], "no_trie": data[:data.index(b"EIRT")],
             "bad_magic": data.replace(b"EIRT", b"XXXX", 1)}
    expected = {"garbage": 0x2006, "truncated_lm": 0x2006, "no_trie": 0x2007, "bad_magic": 0x2008}
    for name, blob in cases.items():
        p = tmp_path / name
        p.write_bytes(blob)
        err = ctypes.c_int()
        h = S.sc_load(str(p).encode(), lab, len(english), 0, ctypes.byref(err))
        assert not h and err.value == expected[name], (name, hex(err.value))
