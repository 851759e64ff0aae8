"""B200-native drop-in for ``models/vqvae/bottleneck.py`` of vliu15/speech-masters-thesis.

Same classes, constructor arguments, attributes (``k`` buffer, ``init``, ``k_sum``, ``k_elem``, ``mu``,
``threshold``, ``k_bins``, ``emb_width``, ``level_blocks``, ``levels``), methods and return values as the
reference (file:line cited per method).  Underneath, every tensor op of the reference's hot path is
replaced by calls into ``libvqb200.so`` (include/vqb200.h) through a ``torch.autograd.Function``:

    K1  vq_assign          distance + argmin              (bottleneck.py:92-100,126-141)
    K2  vq_gather_st_fwd   gather + straight-through + commit loss, NCT in / NCT out (:143-145,194-201)
        vq_gather_st_bwd   its gradient
    K3  vq_ema_accumulate  per-code sums / counts         (:64-68)
        vq_ema_finalize    EMA, revival, metrics          (:78-89)

PyTorch is used for device memory, streams, the CPU RNG calls that must replay the reference's
(``randperm`` / ``randn_like``) and ``torch.distributed``.  There is no fallback path: tensors must live on
a CUDA device of compute capability 10.x and the library must be built, otherwise a RuntimeError is raised.
"""
import math

import torch
import torch.distributed as distributed
import torch.nn as nn

from . import _lib, dist
from ._lib import check, ptr


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"vqb200: {what} must be a CUDA tensor (this implementation is sm_100a-only, "
                           "there is no CPU fallback)")


class _Workspace:
    """Scratch for K1 (FP16 codebook image, norms, fallback worklist), one buffer per (device, stream) -- two streams
    running vq_assign concurrently must not share the fallback worklist; grows on demand.  Remembers which codebook it
    was last prepared for, so a frozen codebook (the generate_vq_dataset loop) is prepared once."""
    _cache = {}
    _prepared = {}

    @staticmethod
    def _key(device):
        return (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)

    @classmethod
    def get(cls, device, n, t, k, d):
        lib = _lib.load()
        need = int(lib.vq_workspace_bytes(n, t, k, d))
        key = cls._key(device)
        buf = cls._cache.get(key)
        if buf is None or buf.numel() < need:
            buf = torch.empty(max(need, 1 << 20), dtype=torch.uint8, device=device)
            cls._cache[key] = buf
            cls._prepared.pop(key, None)
        return buf

    _counter = 0

    @classmethod
    def prepared_flag(cls, device, ws, k):
        """ALGO_PREPARED when `ws` still holds the operands of this very tensor OBJECT at its current version.

        The tag lives on the tensor object (an address can be recycled by the caching allocator, an object cannot be
        confused with its successor) and is matched against the token of the last preparation done in `ws`.
        Writes that bypass the version counter (``k.data.copy_()``, raw-pointer writes through the C ABI) are not
        seen: call ``invalidate_prepared()`` after those."""
        key = cls._key(device)
        tag = getattr(k, "_vqb200_prepared", None)
        if tag is not None and tag == (cls._prepared.get(key), ws.data_ptr(), k._version):
            return _lib.ALGO_PREPARED
        cls._counter += 1
        cls._prepared[key] = cls._counter
        try:
            k._vqb200_prepared = (cls._counter, ws.data_ptr(), k._version)
        except AttributeError:
            pass
        return 0


    @classmethod
    def invalidate(cls):
        cls._prepared.clear()


def invalidate_prepared():
    """Forget every prepared-codebook tag (needed only after writing a codebook through ``.data`` or a raw pointer,
    which does not move the tensor's version counter)."""
    _Workspace.invalidate()


# --------------------------------------------------------------------------------------- raw ops
def assign(x, k, algo="auto", want_min_d=False, scalars=None):
    """K1 on an NCT tensor (fp32, or bf16 as a bf16-autocast encoder emits it -- no up-cast pass): returns
    (idx [N,T] int64, min_d [N,T] fp32 or None)."""
    lib = _lib.load()
    _require_cuda(x, "x")
    entry = lib.vq_assign_bf16 if x.dtype == torch.bfloat16 else lib.vq_assign
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    n, d, t = x.shape
    kk = k.shape[0]
    idx = torch.empty((n, t), dtype=torch.int64, device=x.device)
    min_d = torch.empty((n, t), dtype=torch.float32, device=x.device) if want_min_d else None
    if n * t == 0:
        return idx, min_d
    ws = _Workspace.get(x.device, n, t, kk, d)
    flag = _Workspace.prepared_flag(x.device, ws, k)
    with torch.cuda.device(x.device):
        check(entry(ptr(x), n, d, t, ptr(k), kk, ptr(idx), ptr(min_d), ptr(scalars), ptr(ws), ws.numel(),
                    _lib.ALGOS[algo] | flag, _stream(x)), "vq_assign")
    return idx, min_d


def assign_grouped(x, k, tok, n_vocab, l_bins, scalars=None):
    """Grouped K1 on an NCT tensor (models/vqtts/bottleneck.py:38-58): returns (q_rel, q_abs), both [N, T] int64."""
    lib = _lib.load()
    _require_cuda(x, "x")
    n, d, t = x.shape
    q_rel = torch.empty((n, t), dtype=torch.int64, device=x.device)
    q_abs = torch.empty((n, t), dtype=torch.int64, device=x.device)
    if n * t == 0:
        return q_rel, q_abs
    ws = _Workspace.get(x.device, n, t, n_vocab * l_bins, d)
    _Workspace._prepared.pop(_Workspace._key(x.device), None)        # this call re-prepares the workspace for another codebook
    tok = tok.reshape(n, t).to(torch.int64).contiguous()
    with torch.cuda.device(x.device):
        check(lib.vq_assign_grouped(ptr(x), n, d, t, ptr(k), n_vocab, l_bins, ptr(tok), ptr(q_rel), ptr(q_abs), None, ptr(scalars),
                                    ptr(ws), ws.numel(), _stream(x)), "vq_assign_grouped")
    return q_rel, q_abs


def decode_nct(idx, k):
    """idx [N,T] int64 -> [N,D,T] fp32 (bottleneck.py:160-169)."""
    lib = _lib.load()
    _require_cuda(idx, "indices")
    n, t = idx.shape
    kk, d = k.shape
    out = torch.empty((n, d, t), dtype=torch.float32, device=idx.device)
    if n * t:
        with torch.cuda.device(idx.device):
            check(lib.vq_decode(ptr(idx), ptr(k), n, d, t, kk, ptr(out), _stream(idx)), "vq_decode")
    return out


def gather_rows(x, rows):
    """out[j] = x[n_j, :, t_j] for flat row ids n*T+t."""
    lib = _lib.load()
    n, d, t = x.shape
    out = torch.empty((rows.numel(), d), dtype=torch.float32, device=x.device)
    if rows.numel():
        with torch.cuda.device(x.device):
            check(lib.vq_gather_rows(ptr(x), ptr(rows), rows.numel(), n, d, t, ptr(out), _stream(x)), "vq_gather_rows")
    return out


def restart_rows_device(x, mask, k_bins, counter):
    """K rows of the NCT tensor x drawn uniformly (with replacement) among the frames with mask != 0, on the device.
    `counter` is a 1-element int64 CUDA tensor; it selects the random stream together with torch.initial_seed() and is
    incremented here (on the device), so consecutive calls -- and consecutive replays of a captured graph -- differ."""
    lib = _lib.load()
    _require_cuda(x, "x")
    n, d, t = x.shape
    out = torch.empty((k_bins, d), dtype=torch.float32, device=x.device)
    scratch = torch.empty(n + 1, dtype=torch.int64, device=x.device)
    m = mask.reshape(n, t) if mask is not None else None
    if m is not None and (m.dtype != torch.float32 or not m.is_contiguous()):
        m = m.to(torch.float32).contiguous()
    with torch.cuda.device(x.device):
        check(lib.vq_restart_rows_device(ptr(x), ptr(m), n, d, t, k_bins, torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, ptr(counter),
                                         ptr(out), ptr(scratch), _stream(x)), "vq_restart_rows_device")
    counter.add_(1)
    return out


class _QuantizeST(torch.autograd.Function):
    """Forward: K1 + K2.  Backward: straight-through + commitment gradient (only ``x`` gets a gradient)."""

    @staticmethod
    def forward(ctx, x, mask, k, algo, after_assign, fused_stats, group=None, x_assign=None):
        lib = _lib.load()
        n, d, t = x.shape
        kk = k.shape[0]
        scalars = torch.zeros(_lib.NUM_SCALARS, dtype=torch.float64, device=x.device)
        results = torch.zeros(_lib.NUM_RESULTS, dtype=torch.float32, device=x.device)
        rel = None
        if group is not None:                  # (tok, n_vocab, l_bins): phoneme-conditioned codebook subsets
            rel, idx = assign_grouped(x, k, *group)
        else:
            # indices only; K2 accumulates the `fit` numerator.  (x_assign: the bf16 original of x when the caller up-cast it
            # for K2 / K3 -- K1 reads the bf16 tensor directly)
            idx, _ = assign(x if x_assign is None else x_assign, k, algo)
        if after_assign is not None:
            # EMA statistics (K3a) and their all-reduce are issued here, BEFORE K2: the collective then overlaps K2,
            # which does not depend on it (K1 -> K3a -> {all-reduce || K2} -> K3b)
            after_assign(idx)
        x_q = torch.empty_like(x)
        if n * t and fused_stats is not None:
            # K2 and K3a in one pass over x: K1 -> K2+K3a -> all-reduce -> K3b
            with torch.cuda.device(x.device):
                check(lib.vq_gather_st_fwd_ema(ptr(x), ptr(idx), ptr(mask), ptr(k), n, d, t, kk, ptr(x_q), ptr(scalars),
                                               ptr(results), ptr(fused_stats), _stream(x)), "vq_gather_st_fwd_ema")
        elif n * t:
            with torch.cuda.device(x.device):
                check(lib.vq_gather_st_fwd(ptr(x), ptr(idx), ptr(mask), ptr(k), n, d, t, kk, ptr(x_q), ptr(scalars),
                                           ptr(results), _stream(x)), "vq_gather_st_fwd")
        else:
            results.fill_(float("nan"))
        ctx.save_for_backward(x, idx, mask, k, scalars)
        if rel is None:
            rel = idx
        ctx.mark_non_differentiable(idx, scalars, results, rel)
        return idx, x_q, results[_lib.R_COMMIT], scalars, results, rel

    @staticmethod
    def backward(ctx, _g_idx, g_xq, g_commit, _g_scalars, _g_results, _g_rel):
        lib = _lib.load()
        x, idx, mask, k, scalars = ctx.saved_tensors
        n, d, t = x.shape
        grad_x = torch.empty_like(x)
        if g_xq is not None:
            g_xq = g_xq.contiguous()
        g_commit = (torch.zeros((), dtype=torch.float32, device=x.device) if g_commit is None
                    else g_commit.to(torch.float32).contiguous())
        if n * t:
            with torch.cuda.device(x.device):
                check(lib.vq_gather_st_bwd(ptr(x), ptr(idx), ptr(mask), ptr(k), ptr(g_xq), ptr(g_commit), ptr(scalars),
                                           n, d, t, k.shape[0], ptr(grad_x), _stream(x)), "vq_gather_st_bwd")
        return grad_x, None, None, None, None, None, None, None


# --------------------------------------------------------------------------------------- modules
class BottleneckBlock(nn.Module):
    """Drop-in for ``BottleneckBlock`` (bottleneck.py:10-201)."""

    def __init__(self, k_bins: int, emb_width: int, mu: float, threshold: float, laplace_eps: float = 0.0,
                 algo: str = "auto", rng_parity: bool = True):
        super().__init__()
        self.k_bins = k_bins
        self.emb_width = emb_width
        self.mu = mu
        self.threshold = threshold
        self.laplace_eps = laplace_eps        # 0.0 == the reference (it has no Laplace smoothing)
        self.algo = algo
        # True: restart rows replay the reference's RNG calls (CPU randperm over all valid rows: one host sync and ~2 ms of
        # host time at 300 k frames).  False: K rows are drawn on the device (with replacement, no sync); same distribution,
        # different random stream.
        self.rng_parity = rng_parity
        # True: K2 + K3a in one kernel (vq_gather_st_fwd_ema) when the shape allows it.  Off by default: measured slower than
        # the two separate kernels (0.19-0.22 ms against 0.098 + 0.095 ms at the bench shape), and the
        # separate order lets the all-reduce overlap K2.
        self.fuse_ema = False
        self.reset_k()

    # ---- state (bottleneck.py:20-24)
    def _device_rng_counter(self, device):
        c = getattr(self, "_rng_counter", None)
        if c is None or c.device != device:
            c = torch.zeros(1, dtype=torch.int64, device=device)
            self._rng_counter = c
        return c

    def reset_k(self):
        self.init = False
        self.k_sum = None
        self.k_elem = None
        self.register_buffer("k", torch.zeros(self.k_bins, self.emb_width))

    # ---- restart rows (bottleneck.py:26-33, 39-40, 69-70)
    def _tile(self, x):
        d, ew = x.shape
        if d < self.k_bins:
            n_repeats = (self.k_bins + d - 1) // d
            std = 0.01 / math.sqrt(ew)
            x = x.repeat(n_repeats, 1)
            x = x + torch.randn_like(x) * std
        return x

    def _restart_rows_nct(self, x, mask):
        """K rows of the (tiled) valid rows in random order, drawn with the reference's own RNG calls
        (CPU ``randperm``; ``randn_like`` on x's device only when there are fewer valid rows than codes),
        but gathering K rows instead of permuting the whole batch."""
        n, d, t = x.shape
        flat = mask.reshape(-1) if mask is not None else None
        if not self.rng_parity:
            # device-only (vq_restart_rows_device): K uniform draws among the valid frames, no host sync, graph-capturable
            return restart_rows_device(x, mask, self.k_bins, self._device_rng_counter(x.device))
        valid_pos = (torch.nonzero(flat != 0)[:, 0] if flat is not None
                     else torch.arange(n * t, device=x.device))          # host sync (the reference has three)
        m = valid_pos.numel()
        if m >= self.k_bins:
            perm = torch.randperm(m)[:self.k_bins]
            return gather_rows(x, valid_pos[perm.to(x.device)])
        y = self._tile(gather_rows(x, valid_pos))
        return y[torch.randperm(y.shape[0])][:self.k_bins].contiguous()

    def _set_codebook(self, k_rand):
        if distributed.is_initialized():
            distributed.broadcast(k_rand, 0)                             # bottleneck.py:42
        self.init = True
        self.k = k_rand
        assert self.k.shape == (self.k_bins, self.emb_width)
        self.k_sum = self.k.clone()       # the reference aliases k_sum to k (:45); we update in place, so copy
        self.k_elem = torch.ones(self.k_bins, device=self.k.device)

    def init_k(self, x):
        """bottleneck.py:35-46; ``x`` is the [M, D] matrix of valid rows."""
        y = self._tile(x)
        self._set_codebook(y[torch.randperm(y.shape[0])][:self.k_bins].contiguous())

    def restore_k(self, num_tokens=None, threshold=1.0):
        """bottleneck.py:48-58."""
        self.init = True
        assert self.k.shape == (self.k_bins, self.emb_width)
        self.k_sum = self.k.clone()
        self.k_elem = torch.ones(self.k_bins, device=self.k.device)
        if num_tokens is not None:
            expected_usage = num_tokens / self.k_bins
            self.k_elem.data.mul_(expected_usage)
            self.k_sum.data.mul_(expected_usage)
        self.threshold = threshold

    # ---- EMA (bottleneck.py:60-90)
    def _fuse_ema_ok(self, x):
        """K2 + K3a in one kernel (one pass over x, per-code sums in a shared-memory slab): only when ``fuse_ema`` is set and
        the shape allows it; otherwise K3a runs before K2 (so the all-reduce overlaps K2)."""
        if not self.fuse_ema or x.numel() == 0 or x.data_ptr() % 16:
            return False
        n, d, t = x.shape
        return bool(_lib.load().vq_gather_st_fwd_ema_supported(d, t, self.k_bins))

    def _peer_exchange(self, device):
        """The NVLink peer-memory exchange (dist.PeerExchange), created on the first multi-rank training step (a collective)."""
        if getattr(self, "_peer", None) is None:
            self._peer = dist.PeerExchange.create(self.k_bins, self.emb_width, device) or False
        return self._peer or None

    def _ema_begin(self, x, x_l, mask, k_rand=None, stats=None):
        """K3a on the current stream, then the ONE exchange of the path.  Returns the pending state for ``_ema_finish``.
        Multi-rank: statistics go straight into this rank's peer-visible slot (dist.PeerExchange) and are summed over NVLink
        in ``_ema_finish``; where that is unavailable, ONE NCCL all-reduce is issued here, asynchronously, so that it overlaps K2.
        With ``rng_parity`` the restart rows need a host round trip (``nonzero`` + CPU ``randperm``); that is deferred to
        ``_ema_finish`` so K1 / K3a / K2 are all in flight before the host blocks."""
        lib = _lib.load()
        n, d, t = x.shape
        kk = self.k_bins
        with torch.no_grad():
            peer = self._peer_exchange(x.device) if dist.world()[0] > 1 else None
            k_slot = None
            if stats is None:                    # (else: the fused K2 + K3a kernel already accumulated into `stats`)
                if peer is not None:
                    stats, k_slot = peer.begin_step()
                    stats.zero_()
                else:
                    stats = torch.zeros(dist.stats_numel(kk, d), dtype=torch.float32, device=x.device)
                with torch.cuda.device(x.device):
                    scratch = torch.empty(n * ((t + 63) // 64), dtype=torch.uint8, device=x.device) if mask is not None else None
                    check(lib.vq_ema_accumulate(ptr(x), ptr(x_l), ptr(mask), n, d, t, kk, ptr(stats), ptr(scratch), _stream(x)),
                          "vq_ema_accumulate")
            elif peer is not None:
                own, k_slot = peer.begin_step()
                own.copy_(stats[:own.numel()])
            pending = dict(stats=stats, k_rand=k_rand, work=None, x=x, mask=mask, peer=peer, k_slot=k_slot)
            if k_rand is None and not self.rng_parity:
                pending["k_rand"] = self._restart_rows_nct(x, mask)        # device-side draw: no host sync
            if pending["k_rand"] is not None and peer is not None:
                k_slot.copy_(pending["k_rand"])                            # (only rank 0's rows are read by the peers)
                peer.publish(_stream(x))                                   # before K2: the peers' latency hides behind it
            if pending["k_rand"] is not None and peer is None:
                # reference: broadcast(k_rand) + all_reduce(k_sum) + all_reduce(k_elem) (bottleneck.py:73-75);
                # here ONE all-reduce of the packed buffer (see dist.py), overlapped with K2
                pending["k_rand"], pending["work"] = dist.allreduce_statistics(stats, pending["k_rand"], kk, d, async_op=True)
        return pending

    def _ema_finish(self, pending, scalars, results):
        lib = _lib.load()
        stats, k_rand, x, peer = pending["stats"], pending["k_rand"], pending["x"], pending["peer"]
        kk, d = self.k_bins, self.emb_width
        with torch.no_grad():
            if k_rand is None:                                               # rng_parity: host RNG replay, after K2 was launched
                k_rand = self._restart_rows_nct(x, pending["mask"])
                if peer is None:
                    k_rand, pending["work"] = dist.allreduce_statistics(stats, k_rand, kk, d, async_op=True)
            if peer is not None:
                if pending["k_rand"] is None:                                # (rng_parity: the rows only exist now)
                    pending["k_slot"].copy_(k_rand)
                stats, k_rand = peer.collect(_stream(x))                     # (publish if not done,) wait for the peers, sum in rank order
            elif pending["work"] is not None:
                pending["work"].wait()                                       # the current stream waits for the collective
            k_new = torch.empty_like(self.k)
            used_curr = torch.empty((), dtype=torch.int64, device=x.device)
            if self.k_sum.data_ptr() == self.k.data_ptr():
                self.k_sum = self.k_sum.clone()
            with torch.cuda.device(x.device):
                check(lib.vq_ema_finalize(ptr(stats), ptr(k_rand), ptr(self.k), ptr(k_new), ptr(self.k_sum),
                                          ptr(self.k_elem), kk, d, float(self.mu), float(self.threshold),
                                          float(self.laplace_eps), ptr(scalars), ptr(results), ptr(used_curr),
                                          _stream(x)), "vq_ema_finalize")
            self.k = k_new          # rebind like the reference does (:82); autograd may still hold the old tensor
        return dict(entropy=results[_lib.R_ENTROPY], used_curr=used_curr, usage=results[_lib.R_USAGE],
                    dk=results[_lib.R_DK])

    def _update_k_nct(self, x, x_l, mask, scalars, results, k_rand=None):
        return self._ema_finish(self._ema_begin(x, x_l, mask, k_rand), scalars, results)

    def update_k(self, x, x_l):
        """bottleneck.py:60-90 with the reference's signature: ``x`` [M, D] valid rows, ``x_l`` [M] codes."""
        _require_cuda(x, "x")
        x_nct = x.detach().float().t().contiguous().unsqueeze(0)          # [1, D, M]
        scalars = torch.zeros(_lib.NUM_SCALARS, dtype=torch.float64, device=x.device)
        results = torch.zeros(_lib.NUM_RESULTS, dtype=torch.float32, device=x.device)
        y = self._tile(x.detach().float())
        k_rand = y[torch.randperm(y.shape[0])][:self.k_bins].contiguous()
        return self._update_k_nct(x_nct, x_l.reshape(1, -1).contiguous(), None, scalars, results, k_rand)

    # ---- layout helpers kept for API parity (bottleneck.py:92-124); the kernels read and write NCT directly
    @staticmethod
    def _spread(rows):
        """The ``prenorm`` statistic of bottleneck.py:104 (every caller discards it)."""
        return torch.norm(rows - torch.mean(rows)) / math.sqrt(rows.numel())

    def preprocess(self, x, mask):
        rows = x.transpose(1, 2).reshape(-1, x.shape[1])
        mcol = mask.transpose(1, 2).reshape(-1, 1)
        keep = mcol[:, 0] != 0
        if rows.shape[-1] not in (self.emb_width, 2 * self.emb_width):
            raise AssertionError(f"Expected {rows.shape[-1]} to be (1 or 2) * {self.emb_width}")
        halves = rows.split(self.emb_width, dim=-1)
        prenorm = sum(self._spread(h[keep]) for h in halves)
        rows = halves[0] if len(halves) == 1 else halves[0] + halves[1]
        return rows, prenorm, mcol

    def postprocess(self, x_l, x_d, x_shape, mask):
        n, t = x_shape
        to_nct = lambda a: a.reshape(n, t, -1).transpose(1, 2).contiguous()
        return x_l.reshape(n, t), to_nct(x_d), to_nct(mask)

    # ---- row-major entry points (bottleneck.py:126-145)
    def quantize(self, x, mask=None):
        """``x`` [M, D] rows -> (x_l [M] int64, fit).  ``fit`` follows the reference formula, including its
        (NT,)x(NT,1) broadcast, which reduces to sum(min_d)/K when a mask is given and mean(min_d) otherwise."""
        _require_cuda(x, "x")
        x_nct = x.detach().float().t().contiguous().unsqueeze(0)
        scalars = torch.zeros(_lib.NUM_SCALARS, dtype=torch.float64, device=x.device)
        idx, _ = assign(x_nct, self.k, self.algo, scalars=scalars)
        total = scalars[_lib.S_SUM_MIN_D]
        fit = (total / x.shape[0]) if mask is None else (total / self.k_bins)
        return idx.view(-1), fit.float()

    def dequantize(self, x_l):
        """F.embedding(x_l, k) (bottleneck.py:143-145) for any index shape -> [..., D]."""
        flat = x_l.reshape(1, -1).contiguous()
        out = decode_nct(flat, self.k)                                    # [1, D, M]
        return out[0].t().contiguous().view(*x_l.shape, self.emb_width)

    # ---- NCT entry points
    def _check_input(self, x, mask, keep_bf16=False):
        _require_cuda(x, "x")
        if x.shape[1] == 2 * self.emb_width:                              # bottleneck.py:105-113 (unused by the configs)
            x = x[:, :self.emb_width] + x[:, self.emb_width:]
        assert x.shape[1] == self.emb_width, f"Expected {x.shape[1]} to be (1 or 2) * {self.emb_width}"
        x = x.contiguous()
        if x.dtype != torch.float32 and not (keep_bf16 and x.dtype == torch.bfloat16):
            x = x.float()
        if mask is not None:
            mask = mask.to(torch.float32).contiguous()
            assert mask.numel() == x.shape[0] * x.shape[2]
        return x, mask

    def encode(self, x, mask):
        """bottleneck.py:147-158: [N, D, T] -> [N, T] int64 (K1 only; padded frames get an index too)."""
        x, mask = self._check_input(x.detach(), mask, keep_bf16=True)     # bf16 latents go to K1 as they are
        idx, _ = assign(x, self.k, self.algo)
        return idx

    def decode(self, x_l):
        """bottleneck.py:160-169: [N, T] int64 -> [N, D, T]."""
        _require_cuda(x_l, "x_l")
        return decode_nct(x_l.contiguous(), self.k)

    def forward(self, x, mask, update_k=True):
        """bottleneck.py:171-201 -> (x_l [N,T] int64, x_q [N,D,T], commit_loss, metrics)."""
        x_bf16 = x.detach().contiguous() if (x.dtype == torch.bfloat16 and x.shape[1] == self.emb_width) else None
        x, mask = self._check_input(x, mask)          # (K2 / K3 read FP32: bf16 latents are up-cast for them; K1 reads the original)
        if update_k and not self.init:
            with torch.no_grad():
                self._set_codebook(self._restart_rows_nct(x.detach(), mask))      # init_k (:179-180)
        k = self.k if self.k.dtype == torch.float32 else self.k.float()
        pending, hook, fused_stats = [], None, None
        if update_k and self._fuse_ema_ok(x):
            fused_stats = torch.zeros(dist.stats_numel(self.k_bins, self.emb_width), dtype=torch.float32, device=x.device)
        elif update_k:
            hook = lambda x_l: pending.append(self._ema_begin(x.detach(), x_l, mask))
        x_l, x_q, commit_loss, scalars, results, _ = _QuantizeST.apply(x, mask, k.contiguous(), self.algo, hook, fused_stats, None, x_bf16)
        if update_k:
            if fused_stats is not None:
                pending.append(self._ema_begin(x.detach(), x_l, mask, stats=fused_stats))
            update_metrics = self._ema_finish(pending[0], scalars, results)
        else:
            update_metrics = {}
        return x_l, x_q, commit_loss, dict(fit=results[_lib.R_FIT], **update_metrics)


class GroupedBottleneck(BottleneckBlock):
    """Drop-in for the phoneme-conditioned quantiser ``models/vqtts/bottleneck.py::Bottleneck``: one codebook of
    ``n_vocab * l_bins`` codes, frame j only competes among the ``l_bins`` codes of its aligned token.

    ``forward(y_enc [b, c, ty], x_id [b, tx], attn [b, tx, ty], update_k=True)`` ->
    ``(q_rel [b, ty] int64, y_d [b, c, ty], commit_loss, metrics)`` (vqtts/bottleneck.py:19-77).  Quirks kept: init_k sees all
    frames, padded ones included (:35-36); the EMA update runs iff ``self.training`` (:62-63); ``fit`` is
    sum_all(min_d) / l_bins (the reference's (NT,)x(NT,1) broadcast, :54).  One extension: the reference's
    ``matmul(x_id, attn)`` (:28) only reshapes for b == 1 (it broadcasts to [b, b, ty]); here every utterance is aligned
    with its own ids, which is that expression for b == 1."""

    def __init__(self, n_vocab: int, l_bins: int, emb_width: int, mu: float, threshold: float, **block_kwargs):
        super().__init__(k_bins=n_vocab * l_bins, emb_width=emb_width, mu=mu, threshold=threshold, **block_kwargs)
        self.n_vocab = n_vocab
        self.l_bins = l_bins

    def forward(self, y_enc, x_id, attn, update_k=True):
        _require_cuda(y_enc, "y_enc")
        b, tx, ty = attn.shape
        mask = attn.sum(1).reshape(b, 1, ty)                                          # :25
        tok = torch.matmul(x_id.to(attn.dtype).unsqueeze(1), attn).squeeze(1).long()  # :28  [b, ty]
        x, mask = self._check_input(y_enc, mask)
        if update_k and not self.init:
            with torch.no_grad():
                self._set_codebook(self._restart_rows_nct(x.detach(), None))          # init_k on ALL frames (:35-36)
        k = (self.k if self.k.dtype == torch.float32 else self.k.float()).contiguous()
        pending, hook = [], None
        if self.training:
            hook = lambda q_abs: pending.append(self._ema_begin(x.detach(), q_abs, mask))
        q_abs, x_q, commit_loss, scalars, results, q_rel = _QuantizeST.apply(x, mask, k, self.algo, hook, None,
                                                                             (tok, self.n_vocab, self.l_bins))
        metrics = self._ema_finish(pending[0], scalars, results) if self.training else {}
        fit = (scalars[_lib.S_SUM_MIN_D] / self.l_bins).float()                       # :54
        return q_rel, x_q, commit_loss, dict(fit=fit, **metrics)


class Bottleneck(nn.Module):
    """Drop-in for ``Bottleneck`` (bottleneck.py:204-238)."""

    def __init__(self, l_bins, emb_width, mu, levels, threshold, **block_kwargs):
        super().__init__()
        self.levels = levels
        self.level_blocks = nn.ModuleList()
        for _ in range(self.levels):
            self.level_blocks.append(BottleneckBlock(l_bins, emb_width, mu, threshold, **block_kwargs))

    def encode(self, xs, x_masks=None):
        """bottleneck.py:214-216 omits the mask and raises TypeError as shipped; here the mask is optional."""
        if x_masks is None:
            x_masks = [None] * len(xs)
        return [blk.encode(x, m) for blk, x, m in zip(self.level_blocks, xs, x_masks)]

    def decode(self, zs, start_level=0, end_level=None):
        if end_level is None:
            end_level = self.levels
        return [blk.decode(z) for blk, z in zip(self.level_blocks[start_level:end_level], zs)]

    def forward(self, xs, x_masks):
        zs, xs_quantized, commit_losses, metrics = [], [], [], []
        for level in range(self.levels):
            z, x_q, commit_loss, metric = self.level_blocks[level](xs[level], x_masks[level], update_k=self.training)
            zs.append(z)
            if not self.training:
                x_q = x_q.detach()
            xs_quantized.append(x_q)
            commit_losses.append(commit_loss)
            if self.training:
                metrics.append(metric)
        return zs, xs_quantized, commit_losses, metrics


class NoBottleneckBlock(nn.Module):
    """Identity block used when ``use_bottleneck: false`` (bottleneck.py:241-247)."""

    def forward(self, x, mask, update_k=True):
        return x, x, 0, {}

    def restore_k(self):
        return None


class NoBottleneck(nn.Module):
    """Identity wrapper (bottleneck.py:250-269): passes latents through, reports zero losses and metrics."""
    METRIC_KEYS = ("entropy", "usage", "used_curr", "pn", "dk")

    def __init__(self, levels):
        super().__init__()
        self.levels = levels
        self.level_blocks = nn.ModuleList(NoBottleneckBlock() for _ in range(levels))

    def encode(self, xs):
        return xs

    def decode(self, zs, start_level=0, end_level=None):
        return zs

    def forward(self, xs, x_masks):
        zero = torch.zeros((), device=xs[0].device)
        return xs, xs, [zero] * self.levels, [dict.fromkeys(self.METRIC_KEYS, zero) for _ in range(self.levels)]
