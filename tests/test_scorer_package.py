"""stt_b200.scorer_package (the reference's generate_scorer_package, restated) against the compiled reference:
  * a package WE write is accepted by the REFERENCE loader (Scorer::load_lm + load_trie with OpenFst's ConstFst::Read),
  * its dictionary is the same minimal automaton (state / arc counts of the reference-made package) and accepts exactly
    the same words,
  * the REFERENCE decoder gives identical results with our package and with the reference-made one."""
import ctypes
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, SCORER, VOCAB

BUILD_BINARY = os.path.join(ROOT, "oracle", "_ref", "build_binary")


def _lm_bytes_of(package):
    d = open(package, "rb").read()
    return d[:d.rindex(struct.pack("<ii", 0x54524945, 6))]


def _fst_header(package):
    d = open(package, "rb").read()
    pos = d.rindex(struct.pack("<ii", 0x54524945, 6)) + 25
    magic, = struct.unpack_from("<i", d, pos)
    pos += 4
    names = []
    for _ in range(2):
        n, = struct.unpack_from("<i", d, pos)
        names.append(d[pos + 4:pos + 4 + n])
        pos += 4 + n
    version, flags, props, start, ns, na = struct.unpack_from("<iiQqqq", d, pos)
    return magic, names, version, flags, props, start, ns, na


@pytest.fixture(scope="module")
def packages(tmp_path_factory, ref_decoder, english):
    """(ours, reference-made) for the smoke-test LM + its 3540-word vocabulary."""
    from stt_b200 import scorer_package as sp
    d = tmp_path_factory.mktemp("pkg")
    lm = str(d / "lm.binary")
    open(lm, "wb").write(_lm_bytes_of(SCORER))
    words = open(VOCAB).read().split()
    ours = str(d / "ours.scorer")
    n, ns, na = sp.create_scorer_package(lm, words, english, ours, 0.75, 1.85)
    theirs = str(d / "theirs.scorer")
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    rc = o.ref().ref_make_scorer_package(lm.encode(), b"".join(w.encode() + b"\0" for w in words), len(words), alpha.h,
                                         theirs.encode(), 0.75, 1.85)
    assert rc == 0
    return ours, theirs, words, (n, ns, na)


def test_same_minimal_automaton_and_header(packages):
    ours, theirs, words, (n, ns, na) = packages
    ho, ht = _fst_header(ours), _fst_header(theirs)
    assert ho[:4] == ht[:4]                      # magic, "const"/"standard", aligned file version, flags
    assert ho[5:] == ht[5:] == (0, ns, na)       # start state, #states, #arcs: the minimal DFA is unique
    assert n == len(set(words))
    # every property bit we assert, the reference asserts too (we may leave bits unknown, never contradict)
    assert (ho[4] | ht[4]) == ht[4] or (ho[4] ^ ht[4]) in (0x4000000000 ^ 0x8000000000,)   # only top-sortedness may differ
    assert _lm_bytes_of(ours) == _lm_bytes_of(theirs)
    # the packages are the same size: same header, same alignment padding, same number of states and arcs
    assert os.path.getsize(ours) == os.path.getsize(theirs)


def test_reference_loader_accepts_our_package_and_scores_identically(packages, ref_decoder, english):
    ours, theirs, words, _ = packages
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    so, st = o.RefScorer(ours, alpha), o.RefScorer(theirs, alpha)     # raises if the reference rejects the file
    rng = np.random.default_rng(3)
    for _ in range(500):
        ws = [words[int(rng.integers(len(words)))] for _ in range(int(rng.integers(1, 5)))]
        bos = bool(rng.integers(0, 2))
        assert so.log_cond_prob(ws, bos) == st.log_cond_prob(ws, bos)


def test_reference_decoder_is_indifferent_to_who_packaged(packages, ref_decoder, english):
    from stt_b200 import synth
    ours, theirs, words, _ = packages
    o = ref_decoder
    alpha = o.RefAlphabet(english)
    so, st = o.RefScorer(ours, alpha), o.RefScorer(theirs, alpha)
    for u in range(4):
        probs = synth.make_ctc_probs(words, 120, utt=9100 + u)
        a = o.ref_decode(probs, alpha, 100, so, num_results=3)
        b = o.ref_decode(probs, alpha, 100, st, num_results=3)
        assert len(a) == len(b)
        for (ca, ta, tsa), (cb, tb, tsb) in zip(a, b):
            assert ca == cb and list(ta) == list(tb) and list(tsa) == list(tsb)


def test_our_scorer_view_reads_our_package(packages, english, tmp_path):
    ours, _, words, _ = packages
    so = str(tmp_path / "libscorer_check.so")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "native", "scorer_check.cc"),
                           os.path.join(ROOT, "stt_b200", "csrc", "scorer_image.cc")])
    S = ctypes.CDLL(so)
    S.sc_load.restype = ctypes.c_void_p
    S.sc_load.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    S.sc_fst_find.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    S.sc_fst_final.argtypes = [ctypes.c_void_p, ctypes.c_int]
    S.sc_fst_start.restype = ctypes.c_long
    S.sc_fst_start.argtypes = [ctypes.c_void_p]
    S.sc_alpha.restype = ctypes.c_double
    S.sc_alpha.argtypes = [ctypes.c_void_p]
    err = ctypes.c_int()
    h = S.sc_load(ours.encode(), b"".join(l.encode() + b"\0" for l in english), len(english), 0, ctypes.byref(err))
    assert err.value == 0 and h
    assert S.sc_alpha(h) == 0.75
    lab = {l: i for i, l in enumerate(english)}

    def accepts(word):
        st = S.sc_fst_start(h)
        for ch in word + " ":
            st = S.sc_fst_find(h, st, lab[ch] + 1)
            if st < 0:
                return False
        return bool(S.sc_fst_final(h, st))
    for w in words[::7]:
        assert accepts(w), w
    assert not accepts("zzzz") and not accepts("th")


@pytest.mark.skipif(not os.path.exists(BUILD_BINARY) or not os.path.exists("/root/reference/native_client/kenlm/lm/test.arpa"),
                    reason="needs the reference tree and oracle/_ref/build_binary")
def test_cli_tool_on_a_fresh_lm(tmp_path, ref_decoder, english):
    """End to end like data/lm/generate_lm.py + generate_scorer_package: ARPA -> build_binary -> our packaging CLI ->
    the REFERENCE loads it."""
    from stt_b200 import scorer_package as sp
    lm = str(tmp_path / "t.binary")
    subprocess.check_call([BUILD_BINARY, "-a", "255", "-q", "8", "-v", "trie", "/root/reference/native_client/kenlm/lm/test.arpa", lm],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    alpha_txt = str(tmp_path / "alphabet.txt")
    open(alpha_txt, "w").write("# test alphabet\n" + "\n".join(english) + "\n")
    vocab = str(tmp_path / "vocab.txt")
    open(vocab, "w").write("looking on a little more loin also would consider higher\n")
    pkg = str(tmp_path / "t.scorer")
    assert sp.main(["--alphabet", alpha_txt, "--lm", lm, "--vocab", vocab, "--package", pkg,
                    "--default_alpha", "0.9", "--default_beta", "1.2"]) == 0
    o = ref_decoder
    sc = o.RefScorer(pkg, o.RefAlphabet(english))
    assert sc.log_cond_prob(["looking", "on"], True) < 0.0
    with pytest.raises(ValueError):
        sp.create_scorer_package(pkg, ["a"], english, str(tmp_path / "again.scorer"), 0.9, 1.2)   # already packaged


def test_dictionary_is_exact_and_minimal_on_random_vocabularies(english):
    """build_dictionary accepts exactly {word + space}; no two states are equivalent (minimality) and every state is
    reachable and co-reachable -- on random vocabularies with shared prefixes and suffixes."""
    from stt_b200 import scorer_package as sp
    rng = np.random.default_rng(20260922)
    letters = [l for l in english if l != " "]
    lab = {l: i + 1 for i, l in enumerate(english)}
    for trial in range(20):
        stems = ["".join(rng.choice(letters[:6], size=int(rng.integers(1, 5)))) for _ in range(8)]
        sufs = ["", "s", "ing", "ed", "er"]
        words = sorted({s + sufs[int(rng.integers(len(sufs)))] for s in stems for _ in range(3)})
        start, fin, arcs, n = sp.build_dictionary(words + ["<s>", "</s>", "<unk>", "café"], english)   # skipped entries
        assert n == len(words) and start == 0

        def accepts(w):
            s = 0
            for ch in w + " ":
                nxt = dict(arcs[s]).get(lab[ch])
                if nxt is None:
                    return False
                s = nxt
            return fin[s]
        for w in words:
            assert accepts(w)
        for w in {w[:-1] for w in words if len(w) > 1} | {w + "x" for w in words}:
            assert accepts(w) == (w in words)
        # arcs sorted by label, deterministic; finals have no outgoing arcs (every word ends with the space)
        for s, row in enumerate(arcs):
            assert [l for l, _ in row] == sorted({l for l, _ in row})
            assert not (fin[s] and row)
        # minimal: the right languages of distinct states differ <=> signatures (finality, arcs to classes) are unique
        sigs = {(fin[s], tuple(row)) for s, row in enumerate(arcs)}
        assert len(sigs) == len(arcs)
        # accessible by construction (BFS numbering); co-accessible: every state reaches a final state
        good = {s for s in range(len(arcs)) if fin[s]}
        changed = True
        while changed:
            changed = False
            for s, row in enumerate(arcs):
                if s not in good and any(nx in good for _, nx in row):
                    good.add(s)
                    changed = True
        assert len(good) == len(arcs)


def test_bytes_output_mode_package_matches_the_reference(tmp_path, ref_decoder):
    """UTF-8 mode (generate_scorer_package.cpp:27-50, decoder_utils.cpp:108-131): the vocabulary of code points of the
    multilingual fixture, spelled in bytes without a trailing space; same automaton as the reference builds, accepted by
    the reference loader as a UTF-8 scorer, identical reference decodes with either package."""
    from stt_b200 import scorer_package as sp
    o = ref_decoder
    fixture = os.path.join(GOLDEN, "bytes", "multilingual.bytes.scorer")
    lm = str(tmp_path / "lm.binary")
    open(lm, "wb").write(_lm_bytes_of(fixture))
    text = open(os.path.join(GOLDEN, "bytes", "multilingual.txt"), encoding="utf-8").read()
    units = sorted({ch for ch in text if not ch.isspace()})
    assert sp.looks_char_based(units) and not sp.looks_char_based(["ab"])
    ours = str(tmp_path / "ours.scorer")
    n, ns, na = sp.create_scorer_package(lm, units, None, ours, 0.9, 1.1)      # mode inferred: bytes
    alpha = o.RefByteAlphabet()
    theirs = str(tmp_path / "theirs.scorer")
    rc = o.ref().ref_make_scorer_package_utf8(lm.encode(), b"".join(w.encode("utf-8") + b"\0" for w in units), len(units),
                                              alpha.h, theirs.encode(), 0.9, 1.1)
    assert rc == 0
    ho, ht = _fst_header(ours), _fst_header(theirs)
    assert ho[:4] == ht[:4] and ho[5:] == ht[5:] == (0, ns, na) and n == len(units)
    assert os.path.getsize(ours) == os.path.getsize(theirs)
    so, st = o.RefScorer(ours, alpha), o.RefScorer(theirs, alpha)
    assert o.ref().ref_scorer_is_utf8(so.h) == 1
    rng = np.random.default_rng(4)
    T = 60
    probs = rng.dirichlet(np.ones(256) * 0.05, size=T)
    for t, b in enumerate("日本語caféß😀".encode("utf-8")):
        probs[2 * t + 1, b - 1] += 3.0
    probs /= probs.sum(1, keepdims=True)
    ra, rb = o.ref_decode(probs, alpha, 64, so, num_results=3), o.ref_decode(probs, alpha, 64, st, num_results=3)
    assert len(ra) == len(rb) > 0
    for (ca, ta, tsa), (cb, tb, tsb) in zip(ra, rb):
        assert ca == cb and list(ta) == list(tb) and list(tsa) == list(tsb)
