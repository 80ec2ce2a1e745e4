#!/usr/bin/env python
"""BASELINE.json configs[1]: English-size acoustic model (n_hidden=2048), ONE stream fed in real-time-sized chunks
through the reference's streaming C API (STT_CreateStream / STT_FeedAudioContent / STT_IntermediateDecode /
STT_FinishStream), beam_width=500, KenLM scorer.  Prints one JSON line: real-time factor of the single stream and the
latency of the calls a live client waits on.  Synthetic weights and PCM (see bench.py)."""
import argparse
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stt_b200 import Model, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=10.0)
ap.add_argument("--chunk-ms", type=int, default=20)
ap.add_argument("--beam", type=int, default=500)
ap.add_argument("--n-hidden", type=int, default=2048)
ap.add_argument("--repeats", type=int, default=5)
args = ap.parse_args()

weights = synth.bench_weights(n_hidden=args.n_hidden)
path = os.path.join(tempfile.mkdtemp(), "m.sttw")
synth.write_model(path, weights, beam_width=args.beam)
m = Model(path)
m.enableExternalScorer(os.path.join(ROOT, "tests", "golden", "pruned_lm.scorer"))
pcm = synth.make_pcm(int(args.seconds * 16000), utt=7)
chunk = 16 * args.chunk_ms
feeds, finals, totals, inter = [], [], [], []
text = ""
for rep in range(args.repeats + 1):
    st = m.createStream()
    t_start = time.perf_counter()
    for i, o in enumerate(range(0, len(pcm), chunk)):
        t0 = time.perf_counter()
        st.feedAudioContent(pcm[o:o + chunk])
        dt = time.perf_counter() - t0
        if rep:
            feeds.append(dt)
        if rep and i and i % 100 == 0:
            t0 = time.perf_counter()
            st.intermediateDecode()
            inter.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    text = st.finishStream()
    if rep:
        finals.append(time.perf_counter() - t0)
        totals.append(time.perf_counter() - t_start)
feeds = np.array(feeds) * 1e3
print(json.dumps({
    "config": "batch=1 streaming, n_hidden=%d, beam_width=%d, %d ms chunks, %.0f s utterance, 1xB200" % (
        args.n_hidden, args.beam, args.chunk_ms, args.seconds),
    "rtfx_single_stream": args.seconds / float(np.mean(totals)),
    "feed_ms": {"mean": float(feeds.mean()), "p50": float(np.percentile(feeds, 50)), "p99": float(np.percentile(feeds, 99)),
                "max": float(feeds.max())},
    "feed_ms_note": "16 of every 17 feeds only buffer samples; the 17th (one 16-timestep acoustic chunk = 320 ms of audio) "
                    "runs MFCC + acoustic model + 16 decoder steps",
    "intermediate_decode_ms": float(np.mean(inter) * 1e3) if inter else None,
    "finish_stream_ms": float(np.mean(finals) * 1e3),
    "transcript_head": text[:60]}))
