"""Randomised language models through the scorer view (stt_b200/csrc/scorer_view.h + scorer_image.cc, the code the CUDA
decoder compiles, here compiled for the host) against the compiled reference (Scorer::get_log_cond_prob,
scorer.cpp:301-344 over kenlm's FullScore): orders 2..6, the five model types (kenlm/lm/model_type.hh:8-20) with
varying quantisation / pointer-compression widths, small and large vocabularies, ARPAs closed under sub-n-grams and
pruned ones (lower-order entries missing: the trie builder inserts computed entries, search_trie.cc:207-290).  Every
query must be bit-identical.  Needs oracle/_ref (build_binary + libref_decoder.so: built here from /root/reference by
`make -C oracle ref`; it travels with the tree), no GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

BUILD_BINARY = os.path.join(ROOT, "oracle", "_ref", "build_binary")
LETTERS = "abcdefghijklmnopqrstuvwxyz'"


def _random_arpa(rng, path, order, n_words, n_sent, prune, words=None):
    """An ARPA file over random sentences: all contiguous n-grams up to `order` (so every context and every suffix is
    present), random log10 probabilities and backoffs; `prune` drops a share of the order >= 2 entries afterwards
    (entries of the highest order and contexts of surviving entries alike -- what lmplz --prune leaves behind)."""
    if words is None:
        words = set()
        while len(words) < n_words:
            words.add("".join(LETTERS[int(i)] for i in rng.integers(0, len(LETTERS), int(rng.integers(1, 7)))))
    words = sorted(words)
    n_words = len(words)
    grams = [set() for _ in range(order)]
    for _ in range(n_sent):
        k = int(rng.integers(1, 11))
        # a Zipf-ish draw so that n-grams repeat and long contexts are shared
        idx = np.minimum((rng.pareto(1.1, k) * 3).astype(np.int64), n_words - 1)
        s = ["<s>"] + [words[int(i)] for i in idx] + ["</s>"]
        for n in range(1, order + 1):
            for i in range(len(s) - n + 1):
                grams[n - 1].add(tuple(s[i:i + n]))
    grams[0] |= {("<unk>",), ("<s>",), ("</s>",)} | {(w,) for w in words}
    if prune > 0:
        for n in range(2, order + 1):
            keep = set()
            for g in sorted(grams[n - 1]):
                if rng.random() >= prune:
                    keep.add(g)
            grams[n - 1] = keep
        # the probing builder refuses an n-gram whose context is absent (search_hashed.cc:36); the trie builder fills
        # in: keep contexts for every survivor so that both accept the same file
        for n in range(order, 2, -1):
            for g in grams[n - 1]:
                grams[n - 2].add(g[:-1])
    while order > 1 and not grams[order - 1]:
        order -= 1
    with open(path, "w", encoding="utf-8") as f:
        f.write("\\data\\\n")
        for n in range(1, order + 1):
            f.write("ngram %d=%d\n" % (n, len(grams[n - 1])))
        for n in range(1, order + 1):
            f.write("\n\\%d-grams:\n" % n)
            for g in sorted(grams[n - 1]):
                p = -99.0 if g == ("<s>",) else -float(rng.uniform(0.01, 6.0))
                line = "%.7g\t%s" % (p, " ".join(g))
                if n < order and g[-1] != "</s>":
                    b = 0.0 if rng.random() < 0.2 else -float(rng.uniform(0.0, 2.5))
                    line += "\t%.7g" % b
                f.write(line + "\n")
        f.write("\n\\end\\\n")
    return words, order


def _flags(rng, kind):
    if kind == "trie":
        return ["trie"], []
    if kind == "quant_trie":
        q = int(rng.integers(2, 17))
        return ["trie"], ["-q", str(q), "-b", str(int(rng.integers(2, 17)))]
    if kind == "array_trie":
        return ["trie"], ["-a", str(int(rng.choice([1, 3, 8, 22, 64, 255])))]
    if kind == "quant_array_trie":
        return ["trie"], ["-q", str(int(rng.integers(2, 17))), "-b", str(int(rng.integers(2, 17))),
                          "-a", str(int(rng.choice([1, 5, 16, 64, 255])))]
    return ["probing"], ["-p", "%.2f" % float(rng.uniform(1.2, 3.0))]


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("scf") / "libscorer_check.so")
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


CASES = [(kind, seed) for seed in range(12) for kind in ("trie", "quant_trie", "array_trie", "quant_array_trie", "probing")]


@pytest.mark.parametrize("kind,seed", CASES)
def test_random_lm_matches_reference_bit_for_bit(checker, ref_decoder, english, tmp_path, kind, seed):
    if not os.path.exists(BUILD_BINARY):
        pytest.skip("oracle/_ref/build_binary not built")
    o = ref_decoder
    rng = np.random.default_rng(1000 * seed + len(kind))
    order = int(rng.integers(2, 7))
    n_words = int(rng.choice([12, 60, 300, 1500]))
    prune = float(rng.choice([0.0, 0.0, 0.25, 0.5]))
    arpa = str(tmp_path / "lm.arpa")
    words, order = _random_arpa(rng, arpa, order, n_words, int(rng.integers(40, 900)), prune)
    typ, flags = _flags(rng, kind)
    lm = str(tmp_path / "lm.binary")
    r = subprocess.run([BUILD_BINARY] + flags + ["-v"] + typ + [arpa, lm], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    if r.returncode != 0 and kind == "probing" and b"ProbingSizeException" in r.stderr:
        pytest.skip("kenlm's probing builder ran out of blank space for this pruned file")
    assert r.returncode == 0, r.stderr.decode()[-400:]
    alpha = o.RefAlphabet(english)
    pkg = str(tmp_path / "lm.scorer")
    rc = o.ref().ref_make_scorer_package(lm.encode(), b"".join(w.encode() + b"\0" for w in words), len(words),
                                         alpha.h, pkg.encode(), 0.7, 1.3)
    assert rc == 0
    S = checker
    lab = b"".join(l.encode() + b"\0" for l in english)
    err = ctypes.c_int()
    h = S.sc_load(pkg.encode(), lab, len(english), 0, ctypes.byref(err))
    assert err.value == 0 and h, hex(err.value)
    assert S.sc_order(h) == order
    sc = o.RefScorer(pkg, alpha)
    n_bad = 0
    for _ in range(1500):
        k = int(rng.integers(1, order + 2))
        idx = np.minimum((rng.pareto(1.1, k) * 3).astype(np.int64), len(words) - 1)
        ws = [words[int(i)] for i in idx]
        if rng.random() < 0.08:
            ws[int(rng.integers(k))] = "zzzzzzzz"          # out of vocabulary
        bos = bool(rng.integers(0, 2))
        a = S.sc_log_cond_prob(h, b"".join(w.encode() + b"\0" for w in ws), len(ws), int(bos))
        b = sc.log_cond_prob(ws, bos)
        if a != b:
            n_bad += 1
            if n_bad <= 3:
                print(kind, flags, order, ws, bos, a, b)
    assert n_bad == 0
    # the dictionary automaton (Scorer::fill_dictionary, scorer.cpp:346-372): exactly the vocabulary, each word + space
    lab_of = {l: i for i, l in enumerate(english)}
    start = S.sc_fst_start(h)

    def accepts(word):
        st = start
        for ch in word + " ":
            st = S.sc_fst_find(h, st, lab_of[ch] + 1)
            if st < 0:
                return False
        return bool(S.sc_fst_final(h, st))
    wset = set(words)
    for w in words[:400]:
        assert accepts(w), w
        for cut in (w[:-1], w + "q", "q" + w):
            if cut and cut not in wset:
                assert not accepts(cut), cut
