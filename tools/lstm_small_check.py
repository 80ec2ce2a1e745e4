"""A/B of the LSTM kernels for batches of <= 128 utterances: default (one-CTA kernel, cluster multicast) (STT_B200_LSTM_SMALL_PP=0) vs the CTA-pair kernel
with one group (the default since round 2).  Probabilities must be bit-identical.  usage: python tools/lstm_small_check.py"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
from stt_b200 import Model, synth  # noqa: E402

w = synth.bench_weights(n_hidden=2048)
path = os.path.join(tempfile.mkdtemp(), "m.sttw")
synth.write_model(path, w)
for B in (128, 64, 16, 1):
    pcms = [synth.make_pcm(160000, utt=u) for u in range(B)]
    res = {}
    for env in (None, "1"):
        os.environ["STT_B200_LSTM_SMALL_PP"] = "0" if env is None else env
        m = Model(path)
        b = m.createBatch(B, 160000)
        b.upload(pcms)
        for _ in range(4):
            b.forward()
        res[env] = (b.timings()["lstm"], [b.probs(u) for u in (0, B - 1)])
        del b, m
    same = all(np.array_equal(x, y) for x, y in zip(res[None][1], res["1"][1]))
    print("B=%3d  one-CTA kernel %.3f ms   pair kernel, one group %.3f ms   probabilities %s" % (B, res[None][0], res["1"][0], "bit-identical" if same else "DIFFERENT"), flush=True)
