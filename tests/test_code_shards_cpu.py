"""The compact code-dump format against the reference's pickle format: CPU only.

``dump_batch_to_pickle`` (scripts/generate_vq_dataset.py:83-90) is restated below in two lines (the script itself needs
omegaconf); ``VQLatent.__getitem__`` reads ``pkl["x"]`` / ``pkl["q"]`` (datasets/vqlatent.py:61-68)."""
import json
import os
import pickle

import numpy as np


def reference_pickle_dict(x, xl, q, ql):
    return {"x": x[:xl].flatten().tolist(), "q": q[:ql].flatten().tolist()}          # generate_vq_dataset.py:86-87


def test_shard_round_trip_matches_the_pickle_format(tmp_path):
    import torch
    import vqb200
    rng = np.random.default_rng(0)
    n, t_max, k = 7, 50, 512
    ql = rng.integers(1, t_max + 1, n)
    ql[3] = 0                                                      # an empty utterance
    q = rng.integers(0, k, (n, t_max))
    xl = ql * 128
    x = rng.standard_normal((n, int(xl.max()) + 5)).astype(np.float32)
    w = vqb200.CodeShardWriter(str(tmp_path / "train.vqb2"), vocab_size=k, compression_factor=128)
    # two batches, ragged codes as HostEncoder.encode returns them
    for a, b in ((0, 4), (4, n)):
        ragged = np.concatenate([q[i, :ql[i]] for i in range(a, b)]).astype(np.uint16)
        w.append_batch(ragged, ql[a:b], audio=[x[i] for i in range(a, b)], audio_lengths=xl[a:b])
    w.close()
    shard = vqb200.CodeShard(str(tmp_path / "train.vqb2"))
    assert len(shard) == n and shard.metadata() == {"compression_factor": 128, "vocab_size": k}
    for i in range(n):
        want = reference_pickle_dict(torch.from_numpy(x[i]), int(xl[i]), torch.from_numpy(q[i]), int(ql[i]))
        got = shard[i]
        assert got["q"] == want["q"] and got["x"] == want["x"]
        assert all(isinstance(v, int) for v in got["q"])
    # the reference's own directory layout, read back the way datasets/vqlatent.py does
    shard.export_pickles(str(tmp_path / "dump"), "train")
    assert json.load(open(tmp_path / "dump" / "metadata.json")) == {"compression_factor": 128, "vocab_size": k}
    files = sorted(f for f in os.listdir(tmp_path / "dump" / "train") if f.endswith(".pkl"))
    assert files == [f"{i:05d}.pkl" for i in range(n)]
    with open(tmp_path / "dump" / "train" / files[2], "rb") as f:
        pkl = pickle.load(f)
    assert pkl["q"] == q[2, :ql[2]].tolist() and len(pkl["x"]) == int(xl[2])
    # codes-only shard (audio by reference): x is empty, q unchanged
    w2 = vqb200.CodeShardWriter(str(tmp_path / "codes_only.vqb2"), vocab_size=k)
    w2.append_batch(np.concatenate([q[i, :ql[i]] for i in range(n)]).astype(np.uint16), ql)
    w2.close()
    s2 = vqb200.CodeShard(str(tmp_path / "codes_only.vqb2"))
    assert s2[1]["q"] == q[1, :ql[1]].tolist() and s2[1]["x"] == [] and s2.audio is None
    assert os.path.getsize(tmp_path / "codes_only.vqb2") < 64 + n * 12 + 8 + 2 * int(ql.sum()) + 16


def test_migrating_a_reference_dump(tmp_path):
    """generate_vq_dataset.py's layout -> one shard -> the same items, in the order VQLatent indexes them."""
    import json
    import pickle
    import numpy as np
    from vqb200 import CodeShard
    rng = np.random.default_rng(3)
    dump = tmp_path / "dump"
    (dump / "train").mkdir(parents=True)
    items = []
    for i in range(7):
        n = int(rng.integers(1, 40))
        item = {"x": rng.standard_normal(n * 128).astype(np.float32).tolist(), "q": rng.integers(0, 512, n).tolist()}
        items.append(item)
        with open(dump / "train" / f"{i:05d}.pkl", "wb") as f:
            pickle.dump(item, f)
    with open(dump / "metadata.json", "w") as f:
        json.dump({"compression_factor": 128, "vocab_size": 512}, f)
    shard = CodeShard.from_pickles(str(dump), "train", str(tmp_path / "train.vqb2"))
    assert len(shard) == 7 and shard.metadata() == {"compression_factor": 128, "vocab_size": 512}
    for i, item in enumerate(items):
        got = shard.item(i)
        assert got["q"] == item["q"]
        assert np.allclose(got["x"], item["x"])
    no_audio = CodeShard.from_pickles(str(dump), "train", str(tmp_path / "codes_only.vqb2"), keep_audio=False)
    assert no_audio.item(3) == {"x": [], "q": items[3]["q"]}
    # and back: the exported files are what the reference's reader unpickles
    shard.export_pickles(str(tmp_path / "again"), "train")
    with open(tmp_path / "again" / "train" / "00002.pkl", "rb") as f:
        back = pickle.load(f)
    assert back["q"] == items[2]["q"]
