"""Times the reference path's own torch ops (via the oracle restatement, which uses the same ops:
matmul + min, F.embedding, one-hot matmul) on the B200 with stock PyTorch CUDA kernels -- "the kernel to
beat on the same box" (BASELINE.md section 2).  Reporting only; writes gpurun_out/torch_cuda_baseline.json."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vq_oracle as O  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    dev = torch.device("cuda:0")
    out = {"torch": torch.__version__, "gpu": torch.cuda.get_device_name(0), "allow_tf32": torch.backends.cuda.matmul.allow_tf32}
    gen = torch.Generator().manual_seed(0)
    K, D = 512, 128
    code = torch.randn(K, D, generator=gen)
    for n_utt in (16, 256):
        lengths = O.ljspeech_like_lengths(n_utt, gen)
        x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
        xd, md, kd = x.to(dev), mask.to(dev), code.to(dev)
        rows = x.shape[0] * x.shape[2]

        def flatten():
            return O.flatten_nct(xd, md)

        r, mcol, valid = flatten()

        def quant():
            return O.assign(r, kd, None)

        idx, _, _ = quant()

        def encode_like():           # preprocess + quantize without the NT x NT temp (it OOMs at 256 utterances)
            rr, mm, _ = O.flatten_nct(xd, md)
            return O.assign(rr, kd, mm, faithful_fit=False)

        def deq():
            return O.unflatten(idx, O.gather(idx, kd), x.shape[0], x.shape[2])

        def onehot_stats():
            oh = torch.zeros(K, r.shape[0], device=dev)
            oh.scatter_(0, idx.view(1, -1), 1)
            return torch.matmul(oh, r), oh.sum(-1)

        res = {"rows": rows, "flatten_ms": timeit(flatten), "quantize_ms": timeit(quant), "encode_like_ms": timeit(encode_like),
               "dequantize_postprocess_ms": timeit(deq), "onehot_stats_ms": timeit(onehot_stats)}
        res["encode_like_rows_per_s"] = rows / (res["encode_like_ms"] * 1e-3)
        out[f"utt{n_utt}"] = res
        del xd, md, r, mcol, valid, idx
        torch.cuda.empty_cache()
    # the UNMODIFIED reference module itself (staged copy, baseline/_ref) with stock PyTorch CUDA ops, BASELINE.json configs[0] shape
    from oracle import stage_ref
    if stage_ref.activate():
        from models.vqvae.bottleneck import BottleneckBlock
        lengths = O.ljspeech_like_lengths(16, gen)
        x, mask = O.synthetic_batch(lengths, D, gen, codebook=code)
        xd, md = x.to(dev), mask.to(dev)
        blk = BottleneckBlock(K, D, 0.99, 1.0).to(dev)
        blk.k = code.clone().to(dev)
        blk.k_sum, blk.k_elem, blk.init = (code * 4).to(dev), torch.full((K,), 4.0, device=dev), True
        blk.train()
        with torch.no_grad():
            fwd = timeit(lambda: blk(xd, md, update_k=True), reps=5)
            enc = timeit(lambda: blk.encode(xd, md), reps=5)
        valid = int(lengths.sum())
        out["reference_module_cuda_b16"] = {"rows": x.shape[0] * x.shape[2], "valid_frames": valid, "forward_train_ms": fwd, "encode_ms": enc,
                                            "forward_train_valid_frames_per_s": valid / (fwd * 1e-3), "encode_valid_frames_per_s": valid / (enc * 1e-3),
                                            "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/torch_cuda_baseline.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
