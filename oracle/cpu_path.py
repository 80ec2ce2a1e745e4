"""TEST INFRASTRUCTURE ONLY.  The reference's CPU path for bench.py's cpu_baseline / `--impl reference` legs:

  PCM -> MFCC (oracle/stt_oracle.c) -> restated acoustic model (oracle/am_torch.py, fp32) ->
  GENUINE reference ctc_beam_search_decoder_batch (oracle/_ref/libref_decoder.so)

configured like the reference runs on a many-core host: one stream per worker process with 4 intra-op threads
(native_client/tflitemodelstate.cc:200 `SetNumThreads(4)`), cores/4 streams in flight, then one batch decode with
num_processes = cores (ctc_beam_search_decoder.cpp:608-652).  TFLite itself cannot be built offline, so
the acoustic half is the restatement; the decoder half is the reference's own code.
"""
import multiprocessing as mp
import os
import tempfile
import time

import numpy as np

_AM = None


def _init_worker(weights_path, threads, counter):
    global _AM
    # one stream per worker, pinned to its own `threads` cores: the arm's speed must not depend on how the host's
    # scheduler migrates 32 x 4 unpinned threads (round 1 saw 5.4 vs 25 RTFx on two "128-core" boxes)
    try:
        with counter.get_lock():
            k = counter.value
            counter.value += 1
        cores = sorted(os.sched_getaffinity(0))
        mine = cores[(k * threads) % len(cores):(k * threads) % len(cores) + threads]
        if len(mine) == threads:
            os.sched_setaffinity(0, mine)
    except (AttributeError, OSError):
        pass
    import torch
    torch.set_num_threads(threads)
    from oracle.am_torch import TorchAM
    w = dict(np.load(weights_path))
    _AM = TorchAM(w)


def _run_stream(pcm):
    from oracle import oracle as o
    t0 = time.perf_counter()
    _, mfcc = o.features_only(pcm)
    t1 = time.perf_counter()
    probs = _AM.forward_features(mfcc)
    return probs, t1 - t0, time.perf_counter() - t1


class CpuPath(object):
    def __init__(self, weights, scorer_path, labels, beam, threads_per_stream=4, max_streams=64):
        from oracle import oracle as o
        try:
            self.cores = len(os.sched_getaffinity(0))
        except (AttributeError, OSError):
            self.cores = os.cpu_count() or 1
        self.tps = max(1, min(threads_per_stream, self.cores))
        self.n_streams = max(1, min(max_streams, self.cores // self.tps))
        self.beam = beam
        self.alpha = o.RefAlphabet(labels)
        self.scorer = o.RefScorer(scorer_path, self.alpha)
        self.o = o
        fd, self.wpath = tempfile.mkstemp(suffix=".npz")
        os.close(fd)
        np.savez(self.wpath, **weights)
        ctx = mp.get_context("spawn")
        counter = ctx.Value("i", 0)
        self.pool = ctx.Pool(self.n_streams, initializer=_init_worker, initargs=(self.wpath, self.tps, counter))
        # make sure every worker is initialised before anything is timed
        self.pool.map(_noop, range(self.n_streams * 2))

    def close(self):
        self.pool.close()
        self.pool.join()
        try:
            os.unlink(self.wpath)
        except OSError:
            pass

    def run(self, pcms):
        """One bounded sample: returns (wall seconds, probs list, decode results, breakdown)."""
        t0 = time.perf_counter()
        outs = self.pool.map(_run_stream, pcms, chunksize=1)
        t1 = time.perf_counter()
        probs = [p for p, _, _ in outs]
        T = max(p.shape[0] for p in probs)
        C = probs[0].shape[1]
        arr = np.zeros((len(probs), T, C), np.float64)
        for i, p in enumerate(probs):
            arr[i, :p.shape[0]] = p
        res = self.o.ref_decode_batch(arr, [p.shape[0] for p in probs], self.alpha, self.beam, self.scorer,
                                      num_processes=self.cores)
        t2 = time.perf_counter()
        return t2 - t0, probs, res, {"mfcc_am_wall": t1 - t0, "decode_wall": t2 - t1,
                                     "am_cpu_s_per_stream": float(np.mean([a for _, _, a in outs])),
                                     "mfcc_cpu_s_per_stream": float(np.mean([m for _, m, _ in outs]))}


def _noop(_):
    return 0
