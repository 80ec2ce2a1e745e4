"""Regenerates the committed golden fixtures.  Run in the build container (needs /root/reference for the
reference decoder: `make -C oracle ref`).

  decoder_golden.npz : seeded CTC-like probs [4, 150, 29] + the GENUINE reference decoder's top-1 tokens /
                       timesteps / confidence (beam 100, data/smoke_test/pruned_lm.scorer, cutoff 1.0 / 40)
  mfcc_golden.npz    : seeded PCM (4000 samples) + the oracle restatement's MFCC frames
  lm_golden.npz      : n-grams + Scorer::get_log_cond_prob values from the reference (scorer.cpp:301-344)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.realpath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as o  # noqa: E402
from stt_b200 import synth  # noqa: E402

SCORER = os.path.join(HERE, "pruned_lm.scorer")
words = open(os.path.join(HERE, "vocab.pruned.txt")).read().split()

alpha = o.RefAlphabet(synth.ENGLISH_LABELS)
sc = o.RefScorer(SCORER, alpha)
B, T, beam = 4, 150, 100
probs = np.stack([synth.make_ctc_probs(words, T, utt=31337 + u) for u in range(B)])
tokens, timesteps, conf = [], [], []
for u in range(B):
    c, tk, ts = o.ref_decode(probs[u], alpha, beam, sc)[0]
    tokens.append(tk.astype(np.uint32)); timesteps.append(ts.astype(np.uint32)); conf.append(c)
    print(u, repr(alpha.decode(tk)), c)
np.savez_compressed(os.path.join(HERE, "decoder_golden.npz"), probs=probs, beam=beam,
                    tokens=np.array(tokens, dtype=object), timesteps=np.array(timesteps, dtype=object),
                    confidence=np.array(conf))

pcm = synth.make_pcm(4000, utt=99)
Tn, mfcc = o.features_only(pcm)
np.savez_compressed(os.path.join(HERE, "mfcc_golden.npz"), pcm=pcm, mfcc=mfcc, timesteps=Tn)

rng = np.random.default_rng(11)
grams, bos, vals = [], [], []
for _ in range(400):
    k = int(rng.integers(1, 5))
    ws = [words[int(rng.integers(len(words)))] for _ in range(k)]
    if rng.random() < 0.05:
        ws[int(rng.integers(k))] = "zzzqqq"
    b = int(rng.integers(0, 2))
    grams.append(" ".join(ws)); bos.append(b); vals.append(sc.log_cond_prob(ws, b))
for line in ["she had your dark suit in greasy wash water all year", "we must find a new home in the stars"]:
    ws = line.split()
    for i in range(len(ws)):
        for k in range(1, 5):
            if i + k <= len(ws):
                grams.append(" ".join(ws[i:i + k])); bos.append(int(k < 4)); vals.append(sc.log_cond_prob(ws[i:i + k], k < 4))
np.savez_compressed(os.path.join(HERE, "lm_golden.npz"), grams=np.array(grams), bos=np.array(bos), vals=np.array(vals))
print("wrote golden fixtures")
