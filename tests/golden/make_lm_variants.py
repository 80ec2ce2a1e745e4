"""Regenerates tests/golden/lm_variants/*.scorer: the four KenLM trie layouts the scorer view supports, built from the
reference's own `native_client/kenlm/lm/test.arpa` (the 5-gram model behind kenlm's model_test.cc known answers) with
the reference's own `build_binary` (oracle/_ref/build_binary) and packaged by the reference's Scorer
(fill_dictionary + save_dictionary through oracle/ref_shim.cc).  Run in the build container after `make -C oracle ref`.

  trie.scorer              model_type TRIE              (plain probabilities, inline next pointers)
  quant_trie.scorer        QUANT_TRIE        -q 8 -b 8
  array_trie.scorer        ARRAY_TRIE        -a 64      (Bhiksha-compressed next pointers)
  quant_array_trie.scorer  QUANT_ARRAY_TRIE  -q 8 -b 8 -a 255   (what released .scorer files use)
  probing.scorer           PROBING           -p 1.5             (linear-probing hash tables, kenlm/lm/search_hashed.hh)
(all with -v: no vocabulary strings after the search section, as data/lm/generate_lm.py builds them)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.realpath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402
from stt_b200 import synth  # noqa: E402

ARPA = "/root/reference/native_client/kenlm/lm/test.arpa"
BUILD_BINARY = os.path.join(ROOT, "oracle", "_ref", "build_binary")
OUT = os.path.join(HERE, "lm_variants")
os.makedirs(OUT, exist_ok=True)

# vocabulary = the ARPA's unigrams that can be spelled with the alphabet (Scorer::fill_dictionary skips the rest)
letters = set(synth.ENGLISH_LABELS) - {" "}
words = []
in_uni = False
for line in open(ARPA):
    line = line.rstrip("\n")
    if line.startswith("\\1-grams:"):
        in_uni = True
        continue
    if in_uni and line.startswith("\\"):
        break
    parts = line.split("\t")
    if in_uni and len(parts) >= 2:
        words.append(parts[1])
open(os.path.join(OUT, "vocab.txt"), "w").write("\n".join(words) + "\n")
spellable = [w for w in words if w and all(c in letters for c in w)]

alpha = o.RefAlphabet(synth.ENGLISH_LABELS)
variants = {"trie": [], "quant_trie": ["-q", "8", "-b", "8"], "array_trie": ["-a", "64"],
            "quant_array_trie": ["-q", "8", "-b", "8", "-a", "255"],
            "probing": ["-p", "1.5"]}          # model_type PROBING: hash tables instead of the trie (kenlm/lm/search_hashed.hh)
for name, flags in variants.items():
    lm = os.path.join(OUT, name + ".binary")
    subprocess.check_call([BUILD_BINARY] + flags + ["-v", "probing" if name == "probing" else "trie", ARPA, lm],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    pkg = os.path.join(OUT, name + ".scorer")
    rc = o.ref().ref_make_scorer_package(lm.encode(), b"".join(w.encode() + b"\0" for w in spellable), len(spellable),
                                         alpha.h, pkg.encode(), 0.9, 1.2)
    assert rc == 0, (name, rc)
    os.remove(lm)
    print(name, os.path.getsize(pkg), "bytes")
