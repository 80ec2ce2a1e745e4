"""Multi-GPU use: utterances are independent end to end (own LSTM state, own decoder state; weights and scorer are
read-only replicas), so the path shards with NO data-path collective (the reference does the same at
process level, transcribe.py:50-56).  One process per GPU; torch.distributed is only the control plane
(barrier, max-over-ranks timing, gathering transcripts to rank 0)."""
import heapq


def assign_lpt(lengths, n_parts):
    """Longest-processing-time-first assignment of utterances (by sample count) to `n_parts` ranks.
    Returns a list of index lists; deterministic; every index appears exactly once."""
    parts = [[] for _ in range(n_parts)]
    heap = [(0, r) for r in range(n_parts)]
    heapq.heapify(heap)
    for idx in sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i)):
        load, r = heapq.heappop(heap)
        parts[r].append(idx)
        heapq.heappush(heap, (load + int(lengths[idx]), r))
    return [sorted(p) for p in parts]


def max_over_ranks(value, device=None):
    """MAX all-reduce of a python float (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def transcribe_sharded(transcribe_fn, audio_buffers):
    """Run `transcribe_fn(list_of_buffers) -> list_of_results` on this rank's LPT shard and gather all results on
    every rank in the original order.  With no process group this is just transcribe_fn(audio_buffers)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return transcribe_fn(audio_buffers)
    world, rank = dist.get_world_size(), dist.get_rank()
    parts = assign_lpt([len(a) for a in audio_buffers], world)
    mine = transcribe_fn([audio_buffers[i] for i in parts[rank]]) if parts[rank] else []
    gathered = [None] * world
    dist.all_gather_object(gathered, list(zip(parts[rank], mine)))
    out = [None] * len(audio_buffers)
    for chunk in gathered:
        for i, r in chunk:
            out[i] = r
    return out
