"""Generates gpurun_in/lm_v300k.scorer + words3.txt for tools/large_vocab_check.py (build container only: needs
oracle/_ref/build_binary and the reference's Scorer through oracle/ref_shim.cc; ~30 s).  300 000 random words of 3-10
letters; an order-3 ARPA over 150 000 random sentences (half Zipf-ish, half uniform draws: 1.09 M bigrams, 1.43 M
trigrams, random log-probabilities / backoffs); `build_binary -q 8 -b 8 -a 255 -v trie`; packaged with all words."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402
from stt_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_in")
os.makedirs(OUT, exist_ok=True)
rng = np.random.default_rng(31)
LET = "abcdefghijklmnopqrstuvwxyz'"
words = set()
while len(words) < 300000:
    n = rng.integers(3, 11, 20000)
    codes = rng.integers(0, 27, (20000, 10))
    for i in range(20000):
        words.add("".join(LET[c] for c in codes[i, :n[i]]))
words = sorted(words)[:300000]
open(os.path.join(OUT, "words3.txt"), "w").write("\n".join(words))
order = 3
grams = [set() for _ in range(order)]
for s_i in range(150000):
    k = int(rng.integers(5, 16))
    if s_i % 2:
        idx = rng.integers(0, len(words), k)
    else:
        idx = np.minimum((rng.pareto(0.8, k) * 50).astype(np.int64), len(words) - 1)
    s = ["<s>"] + [words[int(i)] for i in idx] + ["</s>"]
    for n in range(2, order + 1):
        for i in range(len(s) - n + 1):
            grams[n - 1].add(tuple(s[i:i + n]))
grams[0] = {("<unk>",), ("<s>",), ("</s>",)} | {(w,) for w in words}
arpa = os.path.join(OUT, "lm3.arpa")
with open(arpa, "w") as f:
    f.write("\\data\\\n")
    for n in range(order):
        f.write("ngram %d=%d\n" % (n + 1, len(grams[n])))
    for n in range(1, order + 1):
        f.write("\n\\%d-grams:\n" % n)
        gs = sorted(grams[n - 1])
        ps = -rng.uniform(0.01, 6.0, len(gs))
        bs = -rng.uniform(0.0, 2.5, len(gs))
        for j, g in enumerate(gs):
            p = -99.0 if g == ("<s>",) else ps[j]
            if n < order and g[-1] != "</s>":
                f.write("%.7g\t%s\t%.7g\n" % (p, " ".join(g), bs[j]))
            else:
                f.write("%.7g\t%s\n" % (p, " ".join(g)))
    f.write("\n\\end\\\n")
lm = os.path.join(OUT, "lm_v300k.binary")
subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "build_binary"), "-q", "8", "-b", "8", "-a", "255", "-v", "trie", arpa, lm],
                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
alpha = o.RefAlphabet(synth.ENGLISH_LABELS)
pkg = os.path.join(OUT, "lm_v300k.scorer")
assert o.ref().ref_make_scorer_package(lm.encode(), b"".join(w.encode() + b"\0" for w in words), len(words), alpha.h,
                                       pkg.encode(), 0.7, 1.3) == 0
os.remove(arpa)
os.remove(lm)
print(pkg, os.path.getsize(pkg), "bytes")
