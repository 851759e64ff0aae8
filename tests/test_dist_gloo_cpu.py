"""World-size-2 gloo test (CPU) of the one exchange step the training path has: the packed statistics
all-reduce with rank 0's restart rows folded in.  Local statistics come from the oracle; the reduced result
must equal the single-process statistics of the concatenated batch, and every rank must end up with rank 0's
restart rows (what bottleneck.py:73-75 achieves with three collectives)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

K, D = 32, 8


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data():
    g = torch.Generator().manual_seed(11)
    rows = torch.randn(400, D, generator=g)
    idx = torch.randint(0, K, (400,), generator=g)
    return rows, idx


def _worker(rank, world, port, fold_limit, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vqb200
        from oracle import vq_oracle as O
        vqb200.dist.FOLD_LIMIT = fold_limit
        rows, idx = _data()
        a, b = vqb200.dist.shard_range(rows.shape[0], world, rank)
        s_sum, s_elem = O.local_statistics(rows[a:b], idx[a:b], K)
        stats = torch.zeros(vqb200.dist.stats_numel(K, D))
        stats[:K * D] = s_sum.reshape(-1)
        stats[K * D:K * D + K] = s_elem
        k_rand = torch.full((K, D), float(rank + 1))
        k_rand = vqb200.dist.allreduce_statistics(stats, k_rand, K, D)
        torch.save(dict(stats=stats[:K * D + K].clone(), k_rand=k_rand.clone()), f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


def _run(tmp_path, fold_limit):
    from oracle import vq_oracle as O
    out = str(tmp_path / f"res{fold_limit}")
    mp.spawn(_worker, args=(2, _free_port(), fold_limit, out), nprocs=2, join=True)
    rows, idx = _data()
    s_sum, s_elem = O.local_statistics(rows, idx, K)
    for r in range(2):
        got = torch.load(f"{out}.{r}")
        assert torch.allclose(got["stats"][:K * D].view(K, D), s_sum, rtol=1e-5, atol=1e-5)
        assert torch.equal(got["stats"][K * D:], s_elem)
        assert torch.equal(got["k_rand"], torch.ones(K, D))          # rank 0's rows everywhere


def test_folded_single_allreduce(tmp_path):
    _run(tmp_path, 1 << 20)


def test_large_codebook_uses_broadcast_plus_allreduce(tmp_path):
    _run(tmp_path, 0)


def test_shard_range_partitions_everything():
    import vqb200
    for n in (0, 1, 7, 13100):
        for w in (1, 2, 4, 8):
            spans = [vqb200.dist.shard_range(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
