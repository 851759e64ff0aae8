"""BASELINE.json configs[1] end to end: encode a synthetic LJSpeech-sized corpus (13 100 utterances, ~14.8 M frames, ~24 h of
audio at 172 frames/s) from latents in HOST memory to a compact code shard on disk, the way scripts/generate_vq_dataset.py
walks the corpus (batches of utterances padded to the batch maximum), and time the reference's own dump path
(`.tolist()` + pickle per utterance, generate_vq_dataset.py:83-90) on a sample next to it.

    python tools/corpus_encode.py [--batch 256] [--out gpurun_out/corpus_encode.json]
    torchrun --nproc-per-node N tools/corpus_encode.py        # utterance-sharded: one shard file per rank, no collective

The latents are synthetic (one pinned batch per distinct batch shape is generated once and re-used, so host RNG time is not
part of the measurement); every batch is H2D-copied, quantised and its codes D2H-copied and appended to the shard."""
import argparse
import json
import os
import pickle
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as ge  # noqa: E402

N_UTT, K, D = 13100, 512, 128


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--out", default="gpurun_out/corpus_encode.json")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    if rank == 0:
        ge.build()
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
        torch.cuda.set_device(local)
        dist.barrier()
    import vqb200
    from oracle import vq_oracle as O          # synthetic-input recipe only
    gen = torch.Generator().manual_seed(0)
    code = torch.randn(K, D, generator=gen)
    lengths = O.ljspeech_like_lengths(N_UTT, gen).numpy().astype(np.int32)
    a, b = vqb200.dist.shard_range(N_UTT, world, rank)
    mine = lengths[a:b]
    t_max = int(lengths.max())
    enc = vqb200.HostEncoder(local, args.batch * t_max, code.numpy())
    # one synthetic batch of clustered latents, generated once at the largest shape and sliced per batch
    lens0 = torch.from_numpy(np.sort(lengths)[-args.batch:].copy()).long()
    x0, _ = O.synthetic_batch(lens0, D, gen, codebook=code)
    x0 = x0.numpy()
    tmp = tempfile.mkdtemp(prefix="vqb2_")
    path = os.path.join(tmp, f"train.{rank:02d}.vqb2")
    writer = vqb200.CodeShardWriter(path, vocab_size=K, compression_factor=128)
    torch.cuda.synchronize()
    t_fill = t_enc = 0.0
    t0 = time.perf_counter()
    frames = 0
    for s in range(0, len(mine), args.batch):
        ql = mine[s:s + args.batch]
        n, t = len(ql), int(ql.max())
        t1 = time.perf_counter()
        stage = enc.x_staging(n, t)
        stage[...] = x0[:n, :, :t]                       # stands in for the encoder's output landing in pinned memory
        t2 = time.perf_counter()
        codes, lens = enc.encode(stage, ql)
        writer.append_batch(codes, lens)
        t_fill += t2 - t1
        t_enc += time.perf_counter() - t2
        frames += int(ql.sum())
    t3 = time.perf_counter()
    writer.close()
    t_write = time.perf_counter() - t3
    wall = time.perf_counter() - t0
    size = os.path.getsize(path)
    # the reference's dump path on a sample: python lists + one pickle per utterance (codes AND the raw audio, as shipped)
    sample = 16
    shard = vqb200.CodeShard(path)
    audio = [torch.randn(int(l) * 128) for l in mine[:sample]]
    t4 = time.perf_counter()
    for i in range(sample):
        q = torch.from_numpy(shard.tokens(i).astype(np.int64))
        with open(os.path.join(tmp, f"{i:05d}.pkl"), "wb") as f:
            pickle.dump({"x": audio[i].flatten().tolist(), "q": q.flatten().tolist()}, f)       # generate_vq_dataset.py:86-89
    t_ref = (time.perf_counter() - t4) / sample
    t5 = time.perf_counter()
    for i in range(sample):
        q = torch.from_numpy(shard.tokens(i).astype(np.int64))
        with open(os.path.join(tmp, f"q{i:05d}.pkl"), "wb") as f:
            pickle.dump({"q": q.flatten().tolist()}, f)
    t_ref_codes = (time.perf_counter() - t5) / sample
    res = {"rank": rank, "world": world, "utterances": int(len(mine)), "valid_frames": frames, "batch": args.batch,
           "wall_s": wall, "fill_pinned_s": t_fill, "encode_h2d_k1_pack_d2h_s": t_enc, "write_file_s": t_write,
           "frames_per_s_wall": frames / wall, "frames_per_s_encode": frames / t_enc, "shard_bytes": size,
           "bytes_per_frame_on_disk": size / frames,
           "reference_dump_s_per_utterance": {"codes_and_audio_as_shipped": t_ref, "codes_only": t_ref_codes,
                                              "extrapolated_corpus_s_one_process": t_ref * N_UTT}}
    if world > 1:
        import torch.distributed as dist
        allres = [None] * world
        dist.all_gather_object(allres, res)
        if rank == 0:
            res = {"ranks": allres, "world": world, "valid_frames": sum(r["valid_frames"] for r in allres),
                   "wall_s_max": max(r["wall_s"] for r in allres),
                   "frames_per_s_wall": sum(r["valid_frames"] for r in allres) / max(r["wall_s"] for r in allres)}
        dist.destroy_process_group()
    if rank == 0:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump(res, open(args.out, "w"), indent=1)
        print(json.dumps(res))


if __name__ == "__main__":
    main()
