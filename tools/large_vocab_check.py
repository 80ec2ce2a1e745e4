"""Scale check of the scorer path: a 300 000-word vocabulary (order-3 quantised array trie, 2.8 M n-grams, 32 MB package;
generated in the build container: random words, `oracle/_ref/build_binary -q 8 -b 8 -a 255`, packaged by the
reference's Scorer) -- the dictionary automaton, word-ordinal tables and LM lookups at a vocabulary 60x the test
fixtures'.  GPU decodes (shared-memory kernel; general kernel under pruning) must equal the reference decoder.
Needs gpurun_in/lm_v300k.scorer + words3.txt (not committed: 32 MB).  Writes one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402  (checker)
from stt_b200 import Model, synth  # noqa: E402

pkg = os.path.join(ROOT, "gpurun_in", "lm_v300k.scorer")
words = open(os.path.join(ROOT, "gpurun_in", "words3.txt")).read().split("\n")
w = synth.make_weights(n_hidden=64, seed=3)
m = Model(synth.model_bytes(w, beam_width=500))
t0 = time.time()
m.enableExternalScorer(pkg)
t_enable = time.time() - t0
alpha = o.RefAlphabet(synth.ENGLISH_LABELS)
t0 = time.time()
sc = o.RefScorer(pkg, alpha)
t_ref_load = time.time() - t0
B, T = 16, 300
rng = np.random.default_rng(2)
sub = [words[int(i)] for i in rng.integers(0, len(words), 4000)]
probs = np.stack([synth.make_ctc_probs(sub, T, utt=9100 + u) for u in range(B)])
out = {"vocabulary": len(words), "scorer_mb": os.path.getsize(pkg) / 1e6, "enable_scorer_s": t_enable, "reference_load_s": t_ref_load,
       "cases": []}
for beam, cp, tn in ((500, 1.0, 40), (64, 1.0, 40), (500, 0.99, 15)):
    m.setBeamWidth(beam)
    b = m.createBatch(B, T * 320)
    b.set_probs(probs, [T] * B)
    b.set_cutoff(cp, tn)
    t0 = time.time()
    b.decode(num_results=3)
    b.fetch()
    dt = time.time() - t0
    same = 0
    for u in range(B):
        ref = o.ref_decode(probs[u], alpha, beam, sc, num_results=3, cutoff_prob=cp, cutoff_top_n=tn)
        got = b.results(u)
        ok = len(got) == len(ref) and all(list(g[1]) == list(r[1]) and list(g[2]) == list(r[2]) and (g[0] == r[0])
                                          for g, r in zip(got, ref))
        same += int(ok)
    txt = alpha.decode(b.results(0)[0][1])
    out["cases"].append({"beam": beam, "cutoff_prob": cp, "cutoff_top_n": tn, "identical": same, "of": B, "decode_s": dt, "first": txt[:60]})
print(json.dumps(out))
