"""Multi-GPU plumbing of the VQ path: one process per GPU, frames sharded by utterance, codebook replicated.

Encode / decode need no communication.  Training exchanges exactly one buffer per step: the packed EMA
statistics ``[K*D sums | K counts]`` (reference: two all_reduce calls, bottleneck.py:74-75) with rank 0's
restart rows riding along as a zero-padded slab (reference: a broadcast, bottleneck.py:73) -- a SUM over
``{rows, 0, 0, ...}`` is a broadcast.  Works on any backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import ctypes
import os
import socket

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


class _DeviceView:
    """Zero-copy view of raw device memory for ``torch.as_tensor`` (CUDA array interface)."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


class PeerExchange:
    """The EMA statistics exchange over NVLink peer memory (csrc/k3_p2p.cuh) instead of NCCL: every rank exports one region with
    CUDA IPC, maps its peers' regions once, and from then on a step costs three tiny kernels (publish a flag in every peer,
    wait for all flags, sum the peers' statistics in rank order).  ``create`` is a collective (handle exchange through the
    default process group) and returns None on every rank when any rank cannot take part (other backend, several nodes,
    IPC / peer access unavailable, ``VQ_P2P=0``): the module then keeps the NCCL all-reduce."""

    def __init__(self):
        self.regions = None

    def __deepcopy__(self, memo):          # mappings are per process and per module: a copied module sets up its own on first use
        return None

    @classmethod
    def create(cls, k_bins, emb_width, device):
        from . import _lib
        n_ranks, rank = world()
        if n_ranks == 1 or device.type != "cuda":
            return None
        lib = _lib.load()
        self = cls()
        self.lib, self.k_bins, self.emb_width, self.device = lib, k_bins, emb_width, device
        self.n_ranks, self.rank, self.step = n_ranks, rank, 0
        ok = os.environ.get("VQ_P2P", "1") != "0" and distributed.get_backend() == "nccl" and n_ranks <= 16
        region, handle = ctypes.c_void_p(), (ctypes.c_ubyte * 64)()
        if ok:
            with torch.cuda.device(device):
                ok = lib.vq_p2p_alloc(int(lib.vq_p2p_region_bytes(k_bins, emb_width)), ctypes.byref(region), handle) == 0
        infos = [None] * n_ranks
        distributed.all_gather_object(infos, (bool(ok), socket.gethostname(), bytes(handle)))
        ok = all(i[0] for i in infos) and len({i[1] for i in infos}) == 1
        regions = (ctypes.c_void_p * n_ranks)()
        if ok:
            with torch.cuda.device(device):
                for r, (_, _, h) in enumerate(infos):
                    if r == rank:
                        regions[r] = region.value
                    else:
                        peer = ctypes.c_void_p()
                        buf = (ctypes.c_ubyte * 64).from_buffer_copy(h)
                        if lib.vq_p2p_open(buf, ctypes.byref(peer)) != 0:
                            ok = False
                            break
                        regions[r] = peer.value
        flags = [None] * n_ranks
        distributed.all_gather_object(flags, bool(ok))
        if not all(flags):
            return None
        self.regions, self.own = regions, region.value
        # the four slot views and the two pairs of output buffers are made once: a step must not pay for tensor construction
        kd = k_bins * emb_width
        self._slots, self._outs = [], []
        for parity in (0, 1):
            st = self._view(lib.vq_p2p_stats_slot(self.own, parity, k_bins, emb_width), kd + k_bins)
            kr = self._view(lib.vq_p2p_krand_slot(self.own, parity, k_bins, emb_width), kd).view(k_bins, emb_width)
            self._slots.append((st, kr))
            self._outs.append((torch.empty(kd + k_bins, dtype=torch.float32, device=device),
                               torch.empty((k_bins, emb_width), dtype=torch.float32, device=device)))
        return self

    def _view(self, ptr, numel):
        return torch.as_tensor(_DeviceView(ptr, numel), device=self.device)

    def begin_step(self):
        """Next step: (statistics slot [K*D + K], restart-row slot [K, D]) of this rank's region, for K3a / the restart rows."""
        self.step += 1
        return self._slots[self.step & 1]

    def publish(self, stream):
        """Tell every peer that this rank's slots of the current step are complete (enqueue right after K3a / the restart rows,
        BEFORE K2: the peers' flags then arrive while K2 runs)."""
        from ._lib import check
        with torch.cuda.device(self.device):
            check(self.lib.vq_p2p_publish(self.regions, self.n_ranks, self.rank, self.step, stream), "vq_p2p_publish")
        self._published = self.step

    def collect(self, stream):
        """Wait for all peers' flags of the current step, return (statistics summed in rank order, rank 0's restart rows) --
        buffers owned by this object, valid until the collect after next."""
        from ._lib import check
        if getattr(self, "_published", 0) != self.step:
            self.publish(stream)
        stats, k_rand = self._outs[self.step & 1]
        with torch.cuda.device(self.device):
            check(self.lib.vq_p2p_collect(self.regions, self.n_ranks, self.rank, self.step, self.k_bins, self.emb_width,
                                          stats.data_ptr(), k_rand.data_ptr(), stream), "vq_p2p_collect")
        return stats, k_rand


def shard_range(n_items, n_ranks, rank):
    """Contiguous [start, stop) of ``n_items`` utterances owned by ``rank`` (sizes differ by at most one)."""
    q, r = divmod(n_items, n_ranks)
    start = rank * q + min(rank, r)
    return start, start + q + (1 if rank < r else 0)
