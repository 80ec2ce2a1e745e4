"""N > 1 host logic on CPU: world_size-2 gloo run of the utterance sharding / gather / max-over-ranks helpers."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _fake_transcribe(bufs):
    return ["%d:%d" % (len(b), int(np.asarray(b, np.int64).sum())) for b in bufs]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from stt_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    bufs = [rng.integers(-100, 100, size=int(n)).astype(np.int16) for n in rng.integers(10, 5000, size=23)]
    out = parallel.transcribe_sharded(_fake_transcribe, bufs)
    slow = parallel.max_over_ranks(1.0 + rank)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out, slow))


def test_sharded_transcribe_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    bufs = [rng.integers(-100, 100, size=int(n)).astype(np.int16) for n in rng.integers(10, 5000, size=23)]
    expect = _fake_transcribe(bufs)
    for rank, out, slow in res:
        assert out == expect
        assert slow == 2.0


def test_lpt_assignment_properties():
    sys.path.insert(0, ROOT)
    from stt_b200 import parallel
    lengths = [160000] * 7 + [16000, 500, 0, 48000, 99999]
    for n in (1, 2, 3, 8):
        parts = parallel.assign_lpt(lengths, n)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(lengths)))
        loads = [sum(lengths[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(lengths)
    assert parallel.assign_lpt(lengths, 4) == parallel.assign_lpt(lengths, 4)
