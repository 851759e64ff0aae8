"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference
(`/root/reference/models/vqvae/bottleneck.py`) on small seeded inputs.

Run in the authoring container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Each case is one ``.npz`` holding the inputs, the codebook state before the call, every
output of the reference call and the state after it.  ``tests/test_oracle_golden.py`` replays
them against ``oracle/vq_oracle.py`` (CPU) and ``tests/test_gpu_parity.py`` against the CUDA
path.  ``gap`` is the per-row difference between the two smallest fp32 distances, stored so
that a replay on another CPU can tell a legitimate near-tie from an error.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
sys.path.insert(0, REF)
from models.vqvae.bottleneck import Bottleneck, BottleneckBlock  # noqa: E402  (the reference)


def lengths_mask(lengths, t):
    ar = torch.arange(t).view(1, 1, t)
    return (ar < torch.tensor(lengths).view(-1, 1, 1)).float()


def make_inputs(seed, n, d, t, lengths, k_bins, clustered=False, pad_value=0.25):
    g = torch.Generator().manual_seed(seed)
    code = torch.randn(k_bins, d, generator=g)
    if clustered:
        j = torch.randint(0, k_bins, (n, t), generator=g)
        x = code[j].permute(0, 2, 1).contiguous() + 0.5 * torch.randn(n, d, t, generator=g)
    else:
        x = torch.randn(n, d, t, generator=g)
    mask = lengths_mask(lengths, t)
    x = (x * mask + pad_value * (1 - mask)).contiguous()
    return x, mask, code


def seeded_block(k_bins, d, mu, thr, code, elem_scale=1.0):
    blk = BottleneckBlock(k_bins, d, mu, thr)
    blk.k = code.clone()
    blk.init = True
    blk.k_sum = code.clone() * elem_scale
    blk.k_elem = torch.ones(k_bins) * elem_scale
    return blk


def row_gaps(blk, x):
    rows = x.permute(0, 2, 1).contiguous().view(-1, x.shape[1])
    k_w = blk.k.t()
    dist = torch.sum(rows ** 2, -1, keepdim=True) - 2 * torch.matmul(rows, k_w) + torch.sum(k_w ** 2, 0, keepdim=True)
    two = torch.topk(dist, 2, dim=-1, largest=False).values
    return (two[:, 1] - two[:, 0]).numpy()


def np_(t):
    return None if t is None else t.detach().cpu().numpy()


def case_forward(name, seed, n, d, t, lengths, k_bins, mu=0.99, thr=1.0, update_k=True, clustered=False,
                 elem_scale=1.0, rng_seed=1234, dup_codes=False):
    x, mask, code = make_inputs(seed, n, d, t, lengths, k_bins, clustered)
    if dup_codes:                       # exact ties: every odd code duplicates its even neighbour
        code[1::2] = code[0::2]
    blk = seeded_block(k_bins, d, mu, thr, code, elem_scale)
    blk.train()
    gap = row_gaps(blk, x)
    before = dict(k0=np_(blk.k), k_sum0=np_(blk.k_sum), k_elem0=np_(blk.k_elem))
    xg = x.clone().requires_grad_(True)
    torch.manual_seed(rng_seed)
    x_l, x_q, commit, metrics = blk(xg, mask, update_k=update_k)
    # gradient of  sum(w * x_q) + 0.7 * commit  wrt x
    gw = torch.Generator().manual_seed(seed + 99)
    w = torch.randn(x_q.shape, generator=gw)
    ((w * x_q).sum() + 0.7 * commit).backward()
    out = dict(x=np_(x), mask=np_(mask), gap=gap, **before, x_l=np_(x_l), x_q=np_(x_q), commit=np_(commit),
               grad_w=np_(w), grad_commit=np.float32(0.7), grad_x=np_(xg.grad),
               k1=np_(blk.k), k_sum1=np_(blk.k_sum), k_elem1=np_(blk.k_elem),
               mu=np.float32(mu), threshold=np.float32(thr), update_k=np.int32(update_k), rng_seed=np.int64(rng_seed))
    for key, val in metrics.items():
        out["metric_" + key] = np_(val)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "rows", n * t, "metrics", {k: float(v) for k, v in metrics.items()})


def case_init(name, seed, n, d, t, lengths, k_bins, rng_seed=77):
    """First training call: init_k (randperm / _tile) followed by the normal step."""
    x, mask, _ = make_inputs(seed, n, d, t, lengths, k_bins)
    blk = BottleneckBlock(k_bins, d, 0.99, 1.0)
    blk.train()
    torch.manual_seed(rng_seed)
    x_l, x_q, commit, metrics = blk(x, mask, update_k=True)
    out = dict(x=np_(x), mask=np_(mask), x_l=np_(x_l), x_q=np_(x_q), commit=np_(commit),
               k1=np_(blk.k), k_sum1=np_(blk.k_sum), k_elem1=np_(blk.k_elem), rng_seed=np.int64(rng_seed),
               mu=np.float32(0.99), threshold=np.float32(1.0))
    for key, val in metrics.items():
        out["metric_" + key] = np_(val)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "rows", n * t, "valid", int(mask.sum()), "metrics", {k: float(v) for k, v in metrics.items()})


def case_encode_decode(name, seed, n, d, t, lengths, k_bins):
    x, mask, code = make_inputs(seed, n, d, t, lengths, k_bins, clustered=True)
    blk = seeded_block(k_bins, d, 0.99, 1.0, code)
    blk.eval()
    gap = row_gaps(blk, x)
    with torch.no_grad():
        z = blk.encode(x, mask)
        xd = blk.decode(z)
        # wrapper in eval mode: detached x_q, empty metrics list
        wrap = Bottleneck(k_bins, d, 0.99, 1, 1.0)
        wrap.level_blocks[0] = blk
        wrap.eval()
        zs, xqs, commits, mets = wrap([x], [mask])
    out = dict(x=np_(x), mask=np_(mask), gap=gap, k0=np_(code), z=np_(z), x_dec=np_(xd),
               wrap_z=np_(zs[0]), wrap_xq=np_(xqs[0]), wrap_commit=np_(commits[0]), wrap_n_metrics=np.int32(len(mets)))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "rows", n * t)


def grouped_inputs(seed, b, c, tx, ty, n_vocab, l_bins, x_lens, y_lens):
    """Inputs of the phoneme-conditioned quantiser (models/vqtts/bottleneck.py:19): latents y_enc [b, c, ty], token ids
    x_id [b, tx] and a hard monotonic alignment attn [b, tx, ty] (frame j of utterance i belongs to exactly one token; padded
    frames / tokens have all-zero columns / rows)."""
    g = torch.Generator().manual_seed(seed)
    code = torch.randn(n_vocab * l_bins, c, generator=g)
    x_id = torch.randint(0, n_vocab, (b, tx), generator=g)
    attn = torch.zeros(b, tx, ty)
    for i in range(b):
        cuts = torch.sort(torch.randperm(y_lens[i] - 1, generator=g)[:x_lens[i] - 1] + 1).values.tolist()
        bounds = [0] + cuts + [y_lens[i]]
        for j in range(x_lens[i]):
            attn[i, j, bounds[j]:bounds[j + 1]] = 1.0
        x_id[i, x_lens[i]:] = 0
    tok = torch.matmul(x_id.float().unsqueeze(1), attn).squeeze(1).long()            # [b, ty] token of every frame
    rel = torch.randint(0, l_bins, (b, ty), generator=g)
    y_enc = code[tok * l_bins + rel].permute(0, 2, 1).contiguous() + 0.4 * torch.randn(b, c, ty, generator=g)
    return y_enc, x_id, attn, code


def case_grouped(name, seed, b, c, tx, ty, n_vocab, l_bins, x_lens, y_lens, training=True, thr=1.0, elem_scale=4.0, rng_seed=4321):
    from models.vqtts.bottleneck import Bottleneck as GroupedBottleneck          # the reference
    y_enc, x_id, attn, code = grouped_inputs(seed, b, c, tx, ty, n_vocab, l_bins, x_lens, y_lens)
    blk = GroupedBottleneck(n_vocab, l_bins, c, 0.99, thr)
    blk.k = code.clone()
    blk.init = True
    blk.k_sum = code.clone() * elem_scale
    blk.k_elem = torch.ones(n_vocab * l_bins) * elem_scale
    blk.train(training)
    before = dict(k0=np_(blk.k), k_sum0=np_(blk.k_sum), k_elem0=np_(blk.k_elem))
    yg = y_enc.clone().requires_grad_(True)
    torch.manual_seed(rng_seed)
    q_rel, y_d, commit, metrics = blk(yg, x_id, attn, update_k=training)
    gw = torch.Generator().manual_seed(seed + 99)
    w = torch.randn(y_d.shape, generator=gw)
    ((w * y_d).sum() + 0.7 * commit).backward()
    out = dict(y_enc=np_(y_enc), x_id=np_(x_id), attn=np_(attn), **before, q_rel=np_(q_rel), y_d=np_(y_d.contiguous()),
               commit=np_(commit), grad_w=np_(w), grad_commit=np.float32(0.7), grad_y=np_(yg.grad),
               k1=np_(blk.k), k_sum1=np_(blk.k_sum), k_elem1=np_(blk.k_elem), n_vocab=np.int32(n_vocab), l_bins=np.int32(l_bins),
               mu=np.float32(0.99), threshold=np.float32(thr), training=np.int32(training), rng_seed=np.int64(rng_seed))
    for key, val in metrics.items():
        out["metric_" + key] = np_(val)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "frames", b * ty, "metrics", {k: float(v) for k, v in metrics.items()})


if __name__ == "__main__":
    torch.set_num_threads(1)          # fixed summation order for the generator run
    if "--grouped" in sys.argv:       # only the phoneme-conditioned quantiser cases (added in round 2)
        # (b = 1 only: the reference's `torch.matmul(x_id, attn)` (vqtts/bottleneck.py:28) broadcasts a [b, tx] id matrix against
        #  the [b, tx, ty] alignment into [b, b, ty] and fails to reshape for b > 1; the authors train it with batch size 1,
        #  scripts/train_vqvae.sh:14)
        case_grouped("grouped_v11_l32_d16", seed=21, b=1, c=16, tx=9, ty=60, n_vocab=11, l_bins=32, x_lens=[9], y_lens=[60])
        case_grouped("grouped_padded_v4_l64_d128", seed=24, b=1, c=128, tx=12, ty=96, n_vocab=4, l_bins=64, x_lens=[10], y_lens=[77])
        case_grouped("grouped_eval_v7_l64_d32", seed=22, b=1, c=32, tx=7, ty=48, n_vocab=7, l_bins=64, x_lens=[7], y_lens=[48],
                     training=False)
        case_grouped("grouped_revival_v5_l16_d8", seed=23, b=1, c=8, tx=6, ty=40, n_vocab=5, l_bins=16, x_lens=[5], y_lens=[33],
                     thr=2.0, elem_scale=1.0)
        sys.exit(0)
    # default width, ragged lengths, training step with EMA
    case_forward("train_k512_d128", seed=1, n=3, d=128, t=100, lengths=[100, 64, 37], k_bins=512, elem_scale=1.0)
    # clustered data, well-used codes (k_elem scaled so nothing is revived)
    case_forward("train_clustered_k64_d32", seed=2, n=4, d=32, t=90, lengths=[90, 88, 51, 3], k_bins=64,
                 clustered=True, elem_scale=5.0)
    # T not a multiple of 4, D not a multiple of 32
    case_forward("train_odd_t37_d48", seed=3, n=2, d=48, t=37, lengths=[37, 19], k_bins=40, elem_scale=2.0)
    # eval-mode forward (update_k=False): metrics = {fit} only
    case_forward("eval_k128_d64", seed=4, n=2, d=64, t=72, lengths=[72, 40], k_bins=128, update_k=False)
    # dead-code revival: high threshold so most codes are re-seeded from k_rand
    case_forward("train_revival_k96_d16", seed=5, n=2, d=16, t=80, lengths=[80, 61], k_bins=96, thr=3.0)
    # fewer valid rows than codes -> _tile repeats rows and adds randn jitter
    case_forward("train_tile_k256_d16", seed=6, n=1, d=16, t=40, lengths=[33], k_bins=256, thr=1.5)
    # exact ties between duplicated codes: lowest index must win
    case_forward("train_ties_k32_d8", seed=7, n=2, d=8, t=50, lengths=[50, 22], k_bins=32, dup_codes=True, elem_scale=3.0)
    # first call: init_k from the batch
    case_init("init_k64_d32", seed=8, n=2, d=32, t=80, lengths=[80, 45], k_bins=64)
    case_init("init_tile_k128_d8", seed=9, n=1, d=8, t=30, lengths=[21], k_bins=128)
    # encode / decode / eval wrapper
    case_encode_decode("encode_k512_d128", seed=10, n=2, d=128, t=64, lengths=[64, 50], k_bins=512)
