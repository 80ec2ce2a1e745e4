"""A/B check of an LSTM launch mode (STT_B200_LSTM_PINGPONG=<mode>) against the default: probabilities of the headline batch must
be bit-identical (same operands, same accumulation order along K), then the per-stage times of both.
usage: python tools/lstm_mode_check.py <mode> [n_hidden] [batch]"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)
mode = sys.argv[1]
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
B = int(sys.argv[3]) if len(sys.argv) > 3 else 256
from stt_b200 import Model, synth  # noqa: E402

w = synth.bench_weights(n_hidden=H)
path = os.path.join(tempfile.mkdtemp(), "m.sttw")
synth.write_model(path, w)
pcms = [synth.make_pcm(160000, utt=u) for u in range(B)]


def run(env):
    if env is None:
        os.environ.pop("STT_B200_LSTM_PINGPONG", None)
    else:
        os.environ["STT_B200_LSTM_PINGPONG"] = env
    m = Model(path)
    b = m.createBatch(B, 160000)
    b.upload(pcms)
    b.forward()
    for _ in range(3):
        b.forward()
    t = b.timings()
    probs = [b.probs(u) for u in (0, 63, 64, 127, 128, B - 1)]
    return probs, t


base, tb = run(None)
print("default: lstm %.3f ms" % tb["lstm"], flush=True)
got, tg = run(mode)
print("mode %s: lstm %.3f ms" % (mode, tg["lstm"]), flush=True)
worst = max(float(np.abs(a - b).max()) for a, b in zip(base, got))
print("max |dp| vs default:", worst, "bit-identical" if worst == 0.0 else "DIFFERENT")
sys.exit(0 if worst == 0.0 else 1)
