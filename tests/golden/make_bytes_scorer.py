"""Regenerates tests/golden/bytes/multilingual.bytes.scorer: a bytes-output (UTF-8) scorer whose units are code points of
1, 2, 3 and 4 bytes, so that Scorer::is_scoring_boundary / make_ngram (scorer.cpp:271-295,353-381) are exercised beyond
ASCII (the reference's own fixture, data/smoke_test/pruned_lm.bytes.scorer, copied next to it, is ASCII only).

An order-3 ARPA file is written from n-gram counts of the sentences below (add-one estimates; KenLM does not check
normalisation), turned into a quantised array trie by the reference's own build_binary (oracle/_ref/build_binary) and
packaged in UTF-8 mode by the reference's Scorer (fill_dictionary + save_dictionary via oracle/ref_shim.cc).
Run in the build container after `make -C oracle ref`."""
import collections
import math
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.realpath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402

SENTENCES = [
    "el niño comió piñata", "größe straße über äpfel", "café au lait déjà vu", "naïve façade coöperate",
    "日本語 の テスト です", "中文 测试 句子", "한국어 시험 문장", "привет мир как дела", "γειά σου κόσμε",
    "smile 😀 rocket 🚀 heart 💙", "mixed 日本 and café 😀 end", "the quick brown fox", "jumps over the lazy dog",
    "ça va très bien merci", "señor müller isst sushi 寿司", "東京 タワー と 富士山", "emoji 🚀🚀 again 😀",
]
OUT = os.path.join(HERE, "bytes")
os.makedirs(OUT, exist_ok=True)
open(os.path.join(OUT, "multilingual.txt"), "w", encoding="utf-8").write("\n".join(SENTENCES) + "\n")

ORDER = 3
counts = [collections.Counter() for _ in range(ORDER)]
for s in SENTENCES:
    units = ["<s>"] + [ch for ch in s if ch != " "] + ["</s>"]   # code points (ARPA is whitespace-delimited: no space unit)
    for n in range(1, ORDER + 1):
        for i in range(len(units) - n + 1):
            counts[n - 1][tuple(units[i:i + n])] += 1
vocab = sorted({u[0] for u in counts[0]} - {"<s>", "</s>"})
lines = ["\\data\\"]
uni = dict(counts[0])
uni[("<unk>",)] = 1
for n in range(ORDER):
    lines.append("ngram %d=%d" % (n + 1, len(uni) if n == 0 else len(counts[n])))
tot = float(sum(uni.values()))
lines += ["", "\\1-grams:"]
for (w,), c in sorted(uni.items()):
    lp = math.log10((c + 1.0) / (tot + len(uni)))
    if w == "<s>":
        lp = -99.0
    lines.append("%.6f\t%s\t%.6f" % (lp, w, -0.30103 - 0.01 * (c % 7)) if w != "</s>" and w != "<unk>" else "%.6f\t%s" % (lp, w))
for n in (2, 3):
    lines += ["", "\\%d-grams:" % n]
    for g, c in sorted(counts[n - 1].items()):
        ctx = counts[n - 2][g[:-1]]
        lp = math.log10((c + 0.5) / (ctx + 0.5 * len(vocab)))
        if n < ORDER and g[-1] != "</s>":
            lines.append("%.6f\t%s\t%.6f" % (lp, " ".join(g), -0.2 - 0.013 * (c % 5)))
        else:
            lines.append("%.6f\t%s" % (lp, " ".join(g)))
lines += ["", "\\end\\", ""]
arpa = os.path.join(OUT, "multilingual.arpa")
open(arpa, "w", encoding="utf-8").write("\n".join(lines))
lm = os.path.join(OUT, "multilingual.binary")
subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "build_binary"), "-q", "8", "-b", "8", "-a", "255", "-v", "trie", arpa, lm],
                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
alpha = o.RefByteAlphabet()
pkg = os.path.join(OUT, "multilingual.bytes.scorer")
rc = o.ref().ref_make_scorer_package_utf8(lm.encode(), b"".join(w.encode("utf-8") + b"\0" for w in vocab), len(vocab), alpha.h,
                                           pkg.encode(), 0.9, 1.1)
assert rc == 0, rc
os.remove(lm)
os.remove(arpa)
sc = o.RefScorer(pkg, alpha)
assert o.ref().ref_scorer_is_utf8(sc.h) == 1
print(os.path.getsize(pkg), "bytes,", len(vocab), "code points, order", o.ref().ref_scorer_max_order(sc.h))
