"""CPU oracle for the VQ bottleneck hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  The shipped path (``speech-masters-thesis_b200``)
never imports anything from ``oracle/`` and fails loudly when its CUDA library
is missing.

What it restates: ``models/vqvae/bottleneck.py`` of vliu15/speech-masters-thesis
(line numbers below refer to that file).  The reference is pure Python whose
arithmetic lives in PyTorch (``requirements.txt:18`` pins only ``torch>=1.8.0``;
this image has torch 2.11.0), so the restatement uses the same torch CPU
primitives for the numerically sensitive steps (``matmul`` + ``min`` for the
assignment, one-hot ``matmul`` for the EMA sums) and keeps everything else as
plain functions over an explicit ``CodebookState``.

Parity pin: the reference ships no tests or golden vectors, so the oracle is
pinned against the reference itself, imported from ``/root/reference`` in the
authoring container by ``tests/golden/make_golden.py``; the vectors that script
wrote are committed under ``tests/golden/`` and ``tests/test_oracle_golden.py``
replays them (bit-exact for indices / straight-through latents, 1e-6 relative
for scalars).  ``/root/reference`` does not exist on the GPU box; nothing here
reads it.

Quirks of the reference that are reproduced on purpose:
  * ``fit`` uses a (NT,)x(NT,1) broadcast (bottleneck.py:140) and therefore
    equals sum_over_ALL_rows(min_d) / K;  ``faithful_fit=True`` materialises the
    NT x NT temporary exactly like the reference (what a user pays for on CPU),
    the default computes the same number without the temporary.
  * x_d and commit_loss use the codebook from BEFORE the EMA update
    (bottleneck.py:184-189).
  * there is no Laplace smoothing; dead codes are re-seeded from ``k_rand``
    by a usage threshold (bottleneck.py:81-83).
  * ``randperm`` runs on the CPU generator whatever the device (bottleneck.py:40,70).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- state
@dataclass
class CodebookState:
    """Mirror of the attributes ``BottleneckBlock`` keeps (bottleneck.py:12-24)."""
    k_bins: int
    emb_width: int
    mu: float = 0.99
    threshold: float = 1.0
    k: torch.Tensor = field(default=None)        # [K, D] fp32 (the registered buffer)
    k_sum: Optional[torch.Tensor] = None         # [K, D]
    k_elem: Optional[torch.Tensor] = None        # [K]
    init: bool = False

    def __post_init__(self):
        if self.k is None:
            self.k = torch.zeros(self.k_bins, self.emb_width)

    def clone(self) -> "CodebookState":
        c = lambda t: None if t is None else t.clone()
        return CodebookState(self.k_bins, self.emb_width, self.mu, self.threshold,
                             c(self.k), c(self.k_sum), c(self.k_elem), self.init)


def safe_log(p: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """utils/torch_utils.py:4-5."""
    return torch.log(torch.clamp(p, min=eps))


# --------------------------------------------------------------------------- layout
def flatten_nct(x: torch.Tensor, mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """bottleneck.py:92-100 (without the discarded ``prenorm``): NCT -> [N*T, C].

    Returns (rows [NT, C], mask column [NT, 1], boolean valid [NT])."""
    rows = x.permute(0, 2, 1).contiguous().view(-1, x.shape[1])
    mcol = mask.permute(0, 2, 1).contiguous().reshape(-1, 1)
    return rows, mcol, (mcol != 0)[:, 0]


def prenorm(rows: torch.Tensor, valid: torch.Tensor) -> torch.Tensor:
    """bottleneck.py:103-104: the metric every caller throws away (kept for completeness)."""
    sel = rows[valid]
    return torch.norm(sel - torch.mean(sel)) / math.sqrt(sel.numel())


def unflatten(x_l: torch.Tensor, x_d: torch.Tensor, n: int, t: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """bottleneck.py:118-124."""
    return x_l.view(n, t), x_d.view(n, t, -1).permute(0, 2, 1).contiguous()


# --------------------------------------------------------------------------- assignment
def distances(rows: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """bottleneck.py:128-133: ||x||^2 - 2 x E^T + ||E||^2 in fp32, expanded form,
    evaluated left to right exactly like the reference expression."""
    k_w = k.t()
    xx = torch.sum(rows ** 2, dim=-1, keepdim=True)
    ee = torch.sum(k_w ** 2, dim=0, keepdim=True)
    return xx - 2 * torch.matmul(rows, k_w) + ee


def assign(rows: torch.Tensor, k: torch.Tensor, mcol: Optional[torch.Tensor] = None,
           faithful_fit: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """bottleneck.py:126-141 ``quantize``.  Returns (x_l int64 [NT], fit 0-dim, min_d [NT]).

    ``torch.min`` returns the lowest index on exact ties."""
    dist = distances(rows, k)
    min_d, x_l = torch.min(dist, dim=-1)
    if mcol is None:
        fit = torch.mean(min_d)                                       # :138
    elif faithful_fit:
        fit = torch.sum(min_d * mcol) / (mcol.sum() * dist.shape[-1])  # :140, NT x NT temp
    else:
        # identical value: sum_ij min_d[j] * m[i] / (sum(m) K) = sum_j min_d[j] / K
        fit = torch.sum(min_d) / dist.shape[-1]
    return x_l, fit, min_d


def assign_chunked(rows: torch.Tensor, k: torch.Tensor, chunk: int = 8192) -> Tuple[torch.Tensor, torch.Tensor]:
    """``quantize(x, mask=None)`` (bottleneck.py:128-134) over row chunks, so that shapes whose [NT, K] distance matrix
    (or the NT x NT ``fit`` temporary) would not fit host memory can still be audited row for row.  Each row's
    distances are the same fp32 expression as in ``assign``.  Returns (x_l int64 [NT], min_d [NT])."""
    idx = torch.empty(rows.shape[0], dtype=torch.int64)
    min_d = torch.empty(rows.shape[0], dtype=rows.dtype)
    for a in range(0, rows.shape[0], chunk):
        m, i = torch.min(distances(rows[a:a + chunk], k), dim=-1)
        idx[a:a + chunk], min_d[a:a + chunk] = i, m
    return idx, min_d


def gather(x_l: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """bottleneck.py:143-145 ``dequantize``."""
    return F.embedding(x_l, k)


def encode(state: CodebookState, x: torch.Tensor, mask: torch.Tensor, faithful_fit: bool = False) -> torch.Tensor:
    """bottleneck.py:147-158."""
    n, _, t = x.shape
    rows, mcol, _ = flatten_nct(x, mask)
    x_l, _, _ = assign(rows, state.k, mcol, faithful_fit)
    return x_l.view(n, t)


def decode(state: CodebookState, x_l: torch.Tensor) -> torch.Tensor:
    """bottleneck.py:160-169."""
    n, t = x_l.shape
    return gather(x_l, state.k).view(n, t, state.emb_width).permute(0, 2, 1).contiguous()


# --------------------------------------------------------------------------- restart rows (RNG)
def tile_rows(rows: torch.Tensor, k_bins: int) -> torch.Tensor:
    """bottleneck.py:26-33 ``_tile``: repeat + jitter when there are fewer rows than codes."""
    d, ew = rows.shape
    if d < k_bins:
        reps = (k_bins + d - 1) // d
        rows = rows.repeat(reps, 1)
        rows = rows + torch.randn_like(rows) * (0.01 / math.sqrt(ew))
    return rows


def draw_restart_rows(valid_rows: torch.Tensor, k_bins: int) -> torch.Tensor:
    """bottleneck.py:39-40 / :69-70: K random rows of the (tiled) batch.  Consumes the
    global torch RNG exactly like the reference: optional ``randn_like`` then ``randperm``."""
    y = tile_rows(valid_rows, k_bins)
    return y[torch.randperm(y.shape[0])][:k_bins]


def init_codebook(state: CodebookState, valid_rows: torch.Tensor, k_rand: Optional[torch.Tensor] = None) -> None:
    """bottleneck.py:35-46.  (Rank-0 broadcast is the caller's business.)"""
    if k_rand is None:
        k_rand = draw_restart_rows(valid_rows, state.k_bins)
    state.init = True
    state.k = k_rand
    state.k_sum = state.k                      # aliased in the reference too (:45)
    state.k_elem = torch.ones(state.k_bins)


def restore_codebook(state: CodebookState, num_tokens: Optional[float] = None, threshold: float = 1.0) -> None:
    """bottleneck.py:48-58."""
    state.init = True
    state.k_sum = state.k.clone()
    state.k_elem = torch.ones(state.k_bins)
    if num_tokens is not None:
        u = num_tokens / state.k_bins
        state.k_elem = state.k_elem * u
        state.k_sum = state.k_sum * u
    state.threshold = threshold


# --------------------------------------------------------------------------- EMA
def local_statistics(valid_rows: torch.Tensor, valid_idx: torch.Tensor, k_bins: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """bottleneck.py:64-68: per-code embedding sums and counts via a dense one-hot GEMM."""
    onehot = torch.zeros(k_bins, valid_rows.shape[0])
    onehot.scatter_(0, valid_idx.view(1, -1), 1)
    return torch.matmul(onehot, valid_rows), onehot.sum(dim=-1)


def ema_update(state: CodebookState, s_sum: torch.Tensor, s_elem: torch.Tensor, k_rand: torch.Tensor) -> Dict[str, torch.Tensor]:
    """bottleneck.py:78-90, given the (already all-reduced) statistics and restart rows."""
    K, D, mu = state.k_bins, state.emb_width, state.mu
    old_k = state.k
    state.k_sum = mu * state.k_sum + (1.0 - mu) * s_sum
    state.k_elem = mu * state.k_elem + (1.0 - mu) * s_elem
    usage = (state.k_elem.view(K, 1) >= state.threshold).float()
    state.k = usage * (state.k_sum.view(K, D) / state.k_elem.view(K, 1)) + (1 - usage) * k_rand
    prob = s_elem / torch.sum(s_elem)
    entropy = -torch.sum(prob * safe_log(prob))
    used_curr = (s_elem >= state.threshold).sum()
    dk = torch.norm(state.k - old_k) / math.sqrt(K * D)
    return dict(entropy=entropy, used_curr=used_curr, usage=torch.sum(usage), dk=dk)


def update_codebook(state: CodebookState, valid_rows: torch.Tensor, valid_idx: torch.Tensor,
                    k_rand: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """bottleneck.py:60-90 ``update_k`` for one process (no collectives)."""
    with torch.no_grad():
        s_sum, s_elem = local_statistics(valid_rows, valid_idx, state.k_bins)
        if k_rand is None:
            k_rand = draw_restart_rows(valid_rows, state.k_bins)
        return ema_update(state, s_sum, s_elem, k_rand)


# --------------------------------------------------------------------------- full forward
def forward(state: CodebookState, x: torch.Tensor, mask: torch.Tensor, update_k: bool = True,
            k_rand: Optional[torch.Tensor] = None, faithful_fit: bool = False):
    """bottleneck.py:171-201.  ``x`` may require grad; returns
    (x_l [N,T] int64, x_q [N,D,T] fp32, commit_loss 0-dim, metrics dict)."""
    n, _, t = x.shape
    rows, mcol, valid = flatten_nct(x, mask)
    if update_k and not state.init:
        init_codebook(state, rows[valid].detach(), k_rand)
    with torch.no_grad():
        x_l, fit, _ = assign(rows, state.k, mcol, faithful_fit)
        x_d = gather(x_l, state.k)
    metrics = {}
    if update_k:
        metrics = update_codebook(state, rows[valid].detach(), x_l[valid], k_rand)
    commit = torch.norm(x_d[valid].detach() - rows[valid]) ** 2 / (mcol.sum() * rows.shape[1])
    x_st = rows + (x_d - rows).detach()
    x_l2, x_q = unflatten(x_l, x_st, n, t)
    mask_nct = mcol.view(n, t, 1).permute(0, 2, 1).contiguous()
    return x_l2, x_q * mask_nct, commit, dict(fit=fit, **metrics)


# --------------------------------------------------------------------------- grouped (phoneme-conditioned) quantiser
def align_tokens(x_id: torch.Tensor, attn: torch.Tensor) -> torch.Tensor:
    """models/vqtts/bottleneck.py:28: the token id of every frame, ``x_id @ attn`` -> [b, ty] int64.

    As shipped the reference multiplies a [b, tx] id matrix with the [b, tx, ty] alignment, which broadcasts to
    [b, b, ty] and only reshapes for b == 1 (the authors train with batch size 1, scripts/train_vqvae.sh:14).  For b == 1
    this is that expression; for b > 1 it is the evident per-utterance intent."""
    return torch.matmul(x_id.to(attn.dtype).unsqueeze(1), attn).squeeze(1).long()


def grouped_forward(state: CodebookState, y_enc: torch.Tensor, x_id: torch.Tensor, attn: torch.Tensor, n_vocab: int,
                    l_bins: int, training: bool = True, update_k: bool = True, k_rand: Optional[torch.Tensor] = None):
    """models/vqtts/bottleneck.py:19-77 ``Bottleneck.forward`` (K = n_vocab * l_bins codes, frame j only competes among
    the l_bins codes of its aligned token).  Returns (q_rel [b, ty] int64, y_d [b, c, ty], commit_loss, metrics).
    Quirks reproduced: init_k sees ALL rows (padded ones too, :35-36); the EMA update runs iff ``training`` (the
    ``update_k`` argument only gates init, :35,63); ``fit`` carries the (NT,)x(NT,1) broadcast, i.e. sum_all(min_d)/l_bins."""
    b, tx, ty = attn.shape
    c = y_enc.shape[1]
    mask = attn.sum(1).reshape(b * ty, 1)                                        # :25
    valid = mask.bool().flatten()
    tok = align_tokens(x_id, attn).reshape(b * ty)                               # :28,32
    rows = y_enc.permute(0, 2, 1).reshape(b * ty, c)                             # :31
    if update_k and not state.init:
        init_codebook(state, rows.detach(), k_rand)                              # :35-36
    with torch.no_grad():
        k = state.k.reshape(n_vocab, l_bins, c)[tok]                             # :39-40  [b*ty, l, c]
        dist = (torch.sum(rows.unsqueeze(1) ** 2, dim=-1) - 2 * torch.bmm(rows.unsqueeze(1), k.transpose(1, 2)).squeeze(1)
                + torch.sum(k ** 2, dim=-1))                                     # :44-49
        min_d, q_rel = torch.min(dist, dim=-1)                                   # :52
        fit = torch.sum(min_d) / l_bins                                          # :54 (== sum(min_d * mask) / (mask.sum() * l) through the broadcast)
        q_abs = (tok * l_bins + q_rel).long()                                    # :58
        y_d = gather(q_abs, state.k)                                             # :60
    metrics = {}
    if training:
        metrics = update_codebook(state, rows[valid].detach(), q_abs[valid], k_rand)    # :62-63
    commit = torch.norm(y_d[valid].detach() - rows[valid]) ** 2 / (mask.sum() * c)      # :68
    y_st = rows + (y_d - rows).detach()                                          # :71
    y_out = (y_st * mask).reshape(b, ty, c).permute(0, 2, 1)                     # :74
    return q_rel.reshape(b, ty), y_out, commit, dict(fit=fit, **metrics)


def backward_wrt_x(x: torch.Tensor, mask: torch.Tensor, x_l: torch.Tensor, k: torch.Tensor,
                   grad_xq: torch.Tensor, grad_commit: float) -> torch.Tensor:
    """Closed form of the autograd contract: only x receives gradient,
    d/dx = mask * grad_xq + grad_commit * 2 (x - e) / (M D) on valid rows."""
    n, d, t = x.shape
    e = decode(CodebookState(k.shape[0], d, k=k), x_l)
    valid = (mask != 0).to(x.dtype)
    m = valid.sum()
    return mask * grad_xq + valid * (2.0 * grad_commit / (m * d)) * (x - e)


# --------------------------------------------------------------------------- fp64 truth / near-tie audit
NEAR_TIE_REL_EPS = 2.0 ** -18


def near_tie_eps(rows: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    """Per-row epsilon for the near-tie rule: eps = 2^-18 (||x||^2 + max_c ||e_c||^2).

    The reference's fp32 expanded-form distance carries an absolute error of a few
    ulp(||x||^2 + ||e||^2) (max |d32 - d64| 1..6e-4 at D=128, i.e. up to
    ~2^-19 relative), so two fp32 implementations with different summation orders can
    only disagree on the argmin when the true gap is below about twice that."""
    xx = (rows.double() ** 2).sum(-1)
    ee = (k.double() ** 2).sum(-1).max()
    return NEAR_TIE_REL_EPS * (xx + ee)


def audit_indices(rows: torch.Tensor, k: torch.Tensor, idx_a: torch.Tensor, idx_b: torch.Tensor) -> Dict[str, float]:
    """Compare two index vectors; every disagreement is re-evaluated in fp64 and classified
    as a near-tie (|d64[a] - d64[b]| <= eps) or a real error."""
    diff = torch.nonzero(idx_a.view(-1) != idx_b.view(-1))[:, 0]
    out = dict(rows=int(idx_a.numel()), mismatches=int(diff.numel()), near_ties=0, errors=0, worst_gap=0.0)
    if diff.numel():
        r = rows[diff].double()
        ka, kb = k[idx_a.view(-1)[diff]].double(), k[idx_b.view(-1)[diff]].double()
        gap = (((r - ka) ** 2).sum(-1) - ((r - kb) ** 2).sum(-1)).abs()
        eps = near_tie_eps(rows[diff], k)
        out["near_ties"] = int((gap <= eps).sum())
        out["errors"] = int((gap > eps).sum())
        out["worst_gap"] = float(gap.max())
        out["worst_gap_over_eps"] = float((gap / eps).max())
    out["match"] = 1.0 - out["mismatches"] / max(1, out["rows"])
    return out


# --------------------------------------------------------------------------- synthetic inputs
def ljspeech_like_lengths(n: int, gen: torch.Generator) -> torch.Tensor:
    """Frame counts of ``n`` LJSpeech-like utterances: durations ~ clipped N(6.57 s, 2.19 s) on
    [1.11, 10.10] s at 22 050 Hz, truncated to a multiple of 512 samples (datasets/ljspeech.py:14,82),
    divided by the 128x compression (configs/models/vqvae.yaml:5-6)."""
    dur = (6.57 + 2.19 * torch.randn(n, generator=gen)).clamp_(1.11, 10.10)
    samples = (dur * 22050).long() // 512 * 512
    return samples // 128


def synthetic_batch(lengths: torch.Tensor, d: int, gen: torch.Generator, codebook: Optional[torch.Tensor] = None,
                    pad_value: float = 0.25) -> Tuple[torch.Tensor, torch.Tensor]:
    """Latents x[N, D, T] (NCT, fp32) padded to the batch max like ``LJSpeech.collate`` and the float
    mask [N, 1, T].  Gaussian when ``codebook`` is None, else clustered: E[j] + 0.5 N(0,1) with a
    Zipf-skewed j.  Padded frames hold one constant vector (the encoder's bias-only output)."""
    n, t = lengths.numel(), int(lengths.max())
    if codebook is None:
        x = torch.randn(n, d, t, generator=gen)
    else:
        kk = codebook.shape[0]
        w = 1.0 / torch.arange(1, kk + 1, dtype=torch.float64)
        j = torch.multinomial(w / w.sum(), n * t, replacement=True, generator=gen).view(n, t)
        x = codebook[j].permute(0, 2, 1).contiguous() + 0.5 * torch.randn(n, d, t, generator=gen)
    ar = torch.arange(t).view(1, 1, t)
    mask = (ar < lengths.view(n, 1, 1)).float()
    x = x * mask + pad_value * (1 - mask)
    return x.contiguous(), mask
