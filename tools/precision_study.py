#!/usr/bin/env python
"""Where does the distance between the B200 path's probabilities and the fp32 restatement come from?

CPU-only study on the BENCHMARK model (synth.bench_weights: random init + calibrated CTC-like output layer that multiplies
hidden-state differences by 200).  For each arithmetic mode of oracle/am_modes.py it reports, against mode "fp32":
max |d logit|, max |d prob|, frames whose arg-max class changes, and whether the reference decoder's transcript changes.
Rounding sources of the "f16" mode are also switched on one at a time.  `hybrid8` is the reference's DEFAULT export
arithmetic (TFLite hybrid int8), i.e. the distance the reference itself keeps from fp32.

  python tools/precision_study.py [--n-hidden 2048] [--utts 2] [--seconds 2.5] [--out profiles/r02_precision_study.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.realpath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-hidden", type=int, default=2048)
    ap.add_argument("--utts", type=int, default=2)
    ap.add_argument("--seconds", type=float, default=2.5)
    ap.add_argument("--beam", type=int, default=500)
    ap.add_argument("--plain", action="store_true", help="uncalibrated random-init output layer (synth.make_weights)")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from oracle import oracle as o
    from oracle.am_modes import F16_KNOBS, ModeAM
    from stt_b200 import synth
    w = synth.make_weights(n_hidden=a.n_hidden) if a.plain else synth.bench_weights(n_hidden=a.n_hidden)
    n = int(a.seconds * 16000)
    feats = [o.features_only(synth.make_pcm(n, utt=u))[1] for u in range(a.utts)]
    alpha = sc = None
    if o.have_ref():
        alpha = o.RefAlphabet(synth.ENGLISH_LABELS)
        sc = o.RefScorer(os.path.join(ROOT, "tests", "golden", "pruned_lm.scorer"), alpha)

    def run(am):
        outs = [am.forward_features(f, return_logits=True) for f in feats]
        txt = [alpha.decode(o.ref_decode(p, alpha, a.beam, sc)[0][1]) for p, _ in outs] if alpha else None
        return outs, txt

    base, base_txt = run(ModeAM(w, "fp32"))
    rows = {}

    def report(name, am):
        outs, txt = run(am)
        dl = max(float(np.abs(l - bl).max()) for (_, l), (_, bl) in zip(outs, base))
        dp = max(float(np.abs(p - bp).max()) for (p, _), (bp, _) in zip(outs, base))
        flips = sum(int((p.argmax(1) != bp.argmax(1)).sum()) for (p, _), (bp, _) in zip(outs, base))
        frames = sum(p.shape[0] for p, _ in outs)
        same = sum(int(x == y) for x, y in zip(txt, base_txt)) if txt else None
        rows[name] = {"max_abs_dlogit": dl, "max_abs_dprob": dp, "argmax_flips": flips, "frames": frames,
                      "transcripts_equal_to_fp32": ("%d/%d" % (same, len(feats))) if txt else None}
        print("%-28s max|dlogit| %.3e  max|dp| %.3e  argmax flips %d/%d  transcripts %s" %
              (name, dl, dp, flips, frames, rows[name]["transcripts_equal_to_fp32"]), flush=True)

    report("f16 (all sources)", ModeAM(w, "f16"))
    for k in F16_KNOBS:
        report("f16: only " + k, ModeAM(w, "f16", knobs={k: True}))
    report("hybrid8 (symmetric)", ModeAM(w, "hybrid8"))
    report("hybrid8 (asymmetric)", ModeAM(w, "hybrid8", asymmetric=True))
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"model": "plain" if a.plain else "bench (calibrated x200 output layer)", "n_hidden": a.n_hidden,
                       "utterances": a.utts, "seconds": a.seconds, "beam": a.beam, "vs": "mode fp32", "rows": rows}, f,
                      indent=1)


if __name__ == "__main__":
    main()
