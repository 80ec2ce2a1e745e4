#!/usr/bin/env python
"""Speed of the general decoder kernel (decoder_general.cuh) for orientation: bytes-output mode (256 classes, the multilingual
UTF-8 scorer of tests/golden/bytes) and the English alphabet under vocabulary pruning, beam 500.  One JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
from stt_b200 import Model, synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
out = {}


def ctc_rows(seq, C, rng, reps):
    rows = []
    for _ in range(reps):
        for lab in seq:
            r = rng.gamma(0.3, 1.0, size=C) * 0.03
            r[lab] += rng.uniform(0.5, 0.95)
            rows.append(r / r.sum())
            r = rng.gamma(0.3, 1.0, size=C) * 0.03
            r[C - 1] += 0.9
            rows.append(r / r.sum())
    return np.asarray(rows, np.float32)


def run(name, labels, scorer, probs, beam, cutoff):
    w = synth.make_weights(n_hidden=16, n_classes=len(labels) + 1, seed=0)
    m = Model(synth.model_bytes(w, labels=labels, beam_width=beam))
    if scorer:
        m.enableExternalScorer(scorer)
    B, T, C = probs.shape
    b = m.createBatch(B, T * 320 + 512)
    b.set_probs(probs, [T] * B)
    b.set_cutoff(*cutoff)
    for _ in range(3):
        b.decode(1)
    ms = b.timings()["decode"]
    b.fetch()
    out[name] = {"utterances": B, "timesteps": T, "classes": C, "beam": beam, "cutoff": cutoff, "decode_ms": ms,
                 "us_per_timestep": 1000.0 * ms / T, "tokens_first": len(b.results(0)[0][1])}
    if os.environ.get("STT_GEN_PROF"):   # variant build with phase clocks (-DSTT_GEN_PROF): kilo-cycles per phase, summed over utterances
        sc = b.decoder_scalars()
        out[name]["phase_kcycles"] = dict(zip(("row+cutoff+parents", "live_update", "children", "select", "compaction", "commit"), sc[8:14]))
        out[name]["mean_candidates_per_step"] = sc[14] / float(B)
        out[name]["phase_kcycles"]["children_count_pass"] = sc[15]


rng = np.random.default_rng(1)
text = open(os.path.join(GOLDEN, "bytes", "multilingual.txt"), encoding="utf-8").read().replace(" ", "").split("\n")[:-1]
seqs = [[x - 1 for x in t.encode("utf-8")] for t in text[:16]]
T = 400
probs = np.stack([ctc_rows(s, 256, rng, 1 + T // (2 * len(s)))[:T] for s in seqs])
run("bytes_mode_256_classes", [bytes([i + 1]) for i in range(255)], os.path.join(GOLDEN, "bytes", "multilingual.bytes.scorer"), probs, 500, (1.0, 40))
words = open(os.path.join(GOLDEN, "vocab.pruned.txt")).read().split()
eng = np.stack([synth.make_ctc_probs(words, T, utt=100 + u) for u in range(16)])
run("english_pruned", synth.ENGLISH_LABELS, os.path.join(GOLDEN, "pruned_lm.scorer"), eng, 500, (0.99, 40))
run("english_unpruned_shared_memory_kernel", synth.ENGLISH_LABELS, os.path.join(GOLDEN, "pruned_lm.scorer"), eng, 500, (1.0, 40))
print(json.dumps(out))
