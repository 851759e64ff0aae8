"""Multi-GPU plumbing of the VQ path: one process per GPU, frames sharded by utterance, codebook replicated.

Encode / decode need no communication.  Training exchanges exactly one buffer per step: the packed EMA
statistics ``[K*D sums | K counts]`` (reference: two all_reduce calls, bottleneck.py:74-75) with rank 0's
restart rows riding along as a zero-padded slab (reference: a broadcast, bottleneck.py:73) -- a SUM over
``{rows, 0, 0, ...}`` is a broadcast.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as distributed

FOLD_LIMIT = 1 << 20      # fold k_rand into the all-reduce when K*D is at most this many floats (4 MB)


def world():
    if distributed.is_available() and distributed.is_initialized():
        return distributed.get_world_size(), distributed.get_rank()
    return 1, 0


def fold_restart_rows(k_bins, emb_width):
    return world()[0] > 1 and k_bins * emb_width <= FOLD_LIMIT


def stats_numel(k_bins, emb_width):
    """Length of the statistics buffer a rank allocates (and zeroes) before vq_ema_accumulate."""
    base = k_bins * emb_width + k_bins
    return base + (k_bins * emb_width if fold_restart_rows(k_bins, emb_width) else 0)


def allreduce_statistics(stats, k_rand, k_bins, emb_width, async_op=False):
    """SUM the local statistics over all ranks and make rank 0's ``k_rand`` everyone's.  Returns the
    ``k_rand`` to use (a view into ``stats`` when folded); with ``async_op`` returns ``(k_rand, work)`` and the
    collective runs on the backend's own stream until ``work.wait()`` (NCCL: the CURRENT stream then waits for it, the
    host does not), which is how the module overlaps it with K2.  No-op for a single process."""
    n_ranks, rank = world()
    if n_ranks == 1:
        return (k_rand, None) if async_op else k_rand
    base = k_bins * emb_width + k_bins
    if fold_restart_rows(k_bins, emb_width):
        assert stats.numel() == base + k_bins * emb_width
        if rank == 0:
            stats[base:].copy_(k_rand.reshape(-1))
        else:
            stats[base:].zero_()
        work = distributed.all_reduce(stats, distributed.ReduceOp.SUM, async_op=async_op)
        k_rand = stats[base:].view(k_bins, emb_width)
        return (k_rand, work) if async_op else k_rand
    k_rand = k_rand.contiguous()
    distributed.broadcast(k_rand, 0)
    work = distributed.all_reduce(stats, distributed.ReduceOp.SUM, async_op=async_op)
    return (k_rand, work) if async_op else k_rand


def shard_range(n_items, n_ranks, rank):
    """Contiguous [start, stop) of ``n_items`` utterances owned by ``rank`` (sizes differ by at most one)."""
    q, r = divmod(n_items, n_ranks)
    start = rank * q + min(rank, r)
    return start, start + q + (1 if rank < r else 0)
