"""Training-mode module forward, back to back, at N ranks (torchrun): ms per forward with the NVLink peer-memory exchange
(VQ_P2P=1, default) or the NCCL all-reduce (VQ_P2P=0); prints one JSON line from rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import vqb200
    from oracle import vq_oracle as O
    dev = torch.device("cuda", local)
    K, D = 512, 128
    gen = torch.Generator().manual_seed(rank)
    code = torch.randn(K, D, generator=torch.Generator().manual_seed(0))
    lengths = O.ljspeech_like_lengths(256, gen)
    x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
    xd, md = x.to(dev), mask.to(dev)
    out = {"world": world, "p2p_env": os.environ.get("VQ_P2P", "1")}
    same_block = os.environ.get("TFB_SAME_BLOCK") == "1"       # bench.py's order: ONE block, rng_parity first, then the device RNG
    order = (("rng_parity", True), ("device_rng", False)) if same_block else (("device_rng", False), ("rng_parity", True))
    blk = None
    for name, parity in order:
        if blk is None or not same_block:
            blk = vqb200.BottleneckBlock(K, D, 0.99, 1.0, rng_parity=parity).to(dev)
            blk.k, blk.k_sum, blk.k_elem, blk.init = code.to(dev), (code * 4).to(dev), torch.full((K,), 4.0, device=dev), True
            blk.train()
        blk.rng_parity = parity
        for _ in range(3):
            blk(xd, md, update_k=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(10):
                blk(xd, md, update_k=True)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 10)
        t = torch.tensor([best], device=dev)
        if world > 1:
            dist.all_reduce(t, dist.ReduceOp.MAX)
        out[name + "_ms"] = float(t)
        out["used_p2p"] = bool(getattr(blk, "_peer", None))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
