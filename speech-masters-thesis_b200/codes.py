"""Compact code dump / load for the latent-encoding loop.

The reference writes one pickle per utterance, ``{"x": audio.tolist(), "q": codes.tolist()}``
(scripts/generate_vq_dataset.py:83-90, through a process pool, :111-121) plus ``metadata.json``
(``{"compression_factor", "vocab_size"}``, :216-220); ``datasets/vqlatent.py:61-68`` reads them back.  Once the
quantiser runs at GPU speed that loop is all ``.cpu()`` + ``.tolist()`` + pickle.  Here:

* ``HostEncoder`` -- the host-buffer C ABI (``vq_encode_host_u16``): latents in host memory -> ragged ``uint16`` codes in
  pinned host memory, 2 bytes per VALID frame over PCIe (the reference moves 8 per padded frame and then slices);
* ``CodeShardWriter`` / ``CodeShard`` -- one binary file per shard: header, per-utterance lengths and offsets, all codes
  as one ``uint16`` array (optionally the audio as one ``float32`` array); memory-mapped on load;
* ``CodeShard.item(i)`` yields what ``VQLatent.__getitem__`` unpickles (``{"x": [...], "q": [...]}``),
  ``export_pickles`` writes the reference's own directory layout, so ``datasets/vqlatent.py`` keeps working unchanged, and
  ``CodeShard.from_pickles`` migrates an existing dump of the reference into a shard.
"""
import ctypes
import json
import os
import pickle
import struct

import numpy as np

from . import _lib

MAGIC = b"VQB2CODE"
VERSION = 1
_HEADER = struct.Struct("<8sIIIIQQQ")          # magic, version, vocab_size, compression_factor, flags, n_utt, total_codes, total_audio
FLAG_AUDIO = 1


class HostEncoder:
    """Encode latents that live in HOST memory: H2D copy, K1, device-side packing to ragged uint16, D2H copy, chunked and
    double-buffered over two streams (``vq_host_ctx_*`` / ``vq_encode_host_u16`` in include/vqb200.h)."""

    def __init__(self, device_index, max_rows, codebook):
        self.lib = _lib.load()
        codebook = np.ascontiguousarray(codebook, dtype=np.float32)
        self.k_bins, self.emb_width = codebook.shape
        self.max_rows = int(max_rows)
        self.ctx = self.lib.vq_host_ctx_create(int(device_index), self.max_rows, self.k_bins, self.emb_width)
        if not self.ctx:
            raise RuntimeError("vqb200: " + self.lib.vq_last_error().decode())
        _lib.check(self.lib.vq_host_ctx_set_codebook(self.ctx, codebook.ctypes.data), "vq_host_ctx_set_codebook")
        self._x = self.lib.vq_host_ctx_x_staging(self.ctx)
        self._codes = self.lib.vq_host_ctx_codes_staging(self.ctx)
        if not self._codes:
            raise RuntimeError("vqb200: " + self.lib.vq_last_error().decode())

    def x_staging(self, n_utt, t_frames):
        """Pinned [n_utt, D, t_frames] float32 view of the context's staging buffer (fill it in place: no extra copy)."""
        count = n_utt * self.emb_width * t_frames
        assert n_utt * t_frames <= self.max_rows
        buf = (ctypes.c_float * count).from_address(self._x)
        return np.frombuffer(buf, dtype=np.float32).reshape(n_utt, self.emb_width, t_frames)

    def encode(self, x_host, lengths=None):
        """x_host [N, D, T] float32 (any host memory; the pinned ``x_staging`` view avoids a pageable copy) ->
        (codes uint16 [sum(lengths)] -- a VIEW of pinned memory, valid until the next call --, lengths int32 [N])."""
        x_host = np.ascontiguousarray(x_host, dtype=np.float32)
        n, d, t = x_host.shape
        assert d == self.emb_width
        if lengths is None:
            lengths = np.full(n, t, dtype=np.int32)
        lengths = np.ascontiguousarray(lengths, dtype=np.int32)
        total = ctypes.c_int64(0)
        _lib.check(self.lib.vq_encode_host_u16(self.ctx, x_host.ctypes.data, n, t, lengths.ctypes.data, self._codes,
                                               ctypes.byref(total)), "vq_encode_host_u16")
        buf = (ctypes.c_uint16 * max(total.value, 1)).from_address(self._codes)
        return np.frombuffer(buf, dtype=np.uint16)[:total.value], lengths

    def close(self):
        if self.ctx:
            self.lib.vq_host_ctx_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:          # noqa: BLE001  (interpreter shutdown)
            pass


class CodeShardWriter:
    """Accumulates utterances and writes ONE binary file on ``close``."""

    def __init__(self, path, vocab_size, compression_factor=128):
        self.path, self.vocab_size, self.compression_factor = path, int(vocab_size), int(compression_factor)
        self._codes, self._lengths, self._audio, self._audio_lengths = [], [], [], []

    def append_batch(self, codes, lengths, audio=None, audio_lengths=None):
        """codes: ragged uint16 [sum(lengths)] as ``HostEncoder.encode`` returns it (copied); audio: optional list of 1-D arrays."""
        self._codes.append(np.array(codes, dtype=np.uint16, copy=True))
        self._lengths.append(np.asarray(lengths, dtype=np.int32))
        if audio is not None:
            for a, n in zip(audio, audio_lengths if audio_lengths is not None else [len(a) for a in audio]):
                self._audio.append(np.asarray(a, dtype=np.float32).reshape(-1)[:int(n)])
                self._audio_lengths.append(int(n))

    def close(self):
        lengths = np.concatenate(self._lengths) if self._lengths else np.zeros(0, np.int32)
        codes = np.concatenate(self._codes) if self._codes else np.zeros(0, np.uint16)
        assert int(lengths.sum()) == codes.size
        has_audio = len(self._audio) > 0
        assert not has_audio or len(self._audio) == lengths.size, "audio must be given for every utterance or for none"
        offsets = np.zeros(lengths.size + 1, np.int64)
        np.cumsum(lengths, out=offsets[1:])
        a_off = np.zeros(lengths.size + 1, np.int64)
        if has_audio:
            np.cumsum(np.asarray(self._audio_lengths, np.int64), out=a_off[1:])
        with open(self.path, "wb") as f:
            f.write(_HEADER.pack(MAGIC, VERSION, self.vocab_size, self.compression_factor, FLAG_AUDIO if has_audio else 0,
                                 lengths.size, codes.size, int(a_off[-1])))
            f.write(lengths.tobytes())
            f.write(b"\0" * ((-lengths.nbytes) % 8))
            f.write(offsets.tobytes())
            if has_audio:
                f.write(a_off.tobytes())
            f.write(codes.tobytes())
            f.write(b"\0" * ((-codes.nbytes) % 8))
            for a in self._audio:
                f.write(a.tobytes())
        return self.path


class CodeShard:
    """Memory-mapped reader of a shard written by ``CodeShardWriter``."""

    def __init__(self, path):
        self.path = path
        raw = np.memmap(path, dtype=np.uint8, mode="r")
        magic, version, self.vocab_size, self.compression_factor, flags, n, total, total_audio = _HEADER.unpack(bytes(raw[:_HEADER.size]))
        if magic != MAGIC or version != VERSION:
            raise ValueError(f"{path}: not a vqb200 code shard")
        pos = _HEADER.size
        self.lengths = np.frombuffer(raw, np.int32, n, pos)
        pos += n * 4 + ((-n * 4) % 8)
        self.offsets = np.frombuffer(raw, np.int64, n + 1, pos)
        pos += (n + 1) * 8
        self.audio_offsets = None
        if flags & FLAG_AUDIO:
            self.audio_offsets = np.frombuffer(raw, np.int64, n + 1, pos)
            pos += (n + 1) * 8
        self.codes = np.frombuffer(raw, np.uint16, total, pos)
        pos += total * 2 + ((-total * 2) % 8)
        self.audio = np.frombuffer(raw, np.float32, total_audio, pos) if flags & FLAG_AUDIO else None

    def __len__(self):
        return int(self.lengths.size)

    def tokens(self, i):
        """Codes of utterance i as a uint16 array (zero copy)."""
        return self.codes[self.offsets[i]:self.offsets[i + 1]]

    def item(self, i):
        """What ``VQLatent.__getitem__`` gets from ``pickle.load`` (datasets/vqlatent.py:61-68): python lists."""
        x = self.audio[self.audio_offsets[i]:self.audio_offsets[i + 1]].tolist() if self.audio is not None else []
        return {"x": x, "q": self.tokens(i).tolist()}

    __getitem__ = item

    def metadata(self):
        return {"compression_factor": self.compression_factor, "vocab_size": self.vocab_size}

    @staticmethod
    def from_pickles(dump_dir, split, path, keep_audio=True):
        """Migrate an existing reference dump (``{split}/NNNNN.pkl`` + ``metadata.json``, generate_vq_dataset.py:86-89,216-220)
        into one shard at ``path``.  Files are taken in name order (the order ``VQLatent`` indexes them in); codes must fit
        ``uint16`` (``vocab_size`` <= 65 536).  Returns the opened shard."""
        with open(os.path.join(dump_dir, "metadata.json")) as f:
            meta = json.load(f)
        if int(meta["vocab_size"]) > 65536:
            raise ValueError("codes of this dump do not fit uint16")
        writer = CodeShardWriter(path, meta["vocab_size"], meta.get("compression_factor", 128))
        names = sorted(n for n in os.listdir(os.path.join(dump_dir, split)) if n.endswith(".pkl"))
        for name in names:
            with open(os.path.join(dump_dir, split, name), "rb") as f:
                item = pickle.load(f)
            q = np.asarray(item["q"], dtype=np.int64).reshape(-1)
            if q.size and (q.min() < 0 or q.max() >= int(meta["vocab_size"])):
                raise ValueError(f"{name}: code outside [0, vocab_size)")
            audio = [np.asarray(item.get("x", []), dtype=np.float32)] if keep_audio else None
            writer.append_batch(q.astype(np.uint16), [q.size], audio=audio)
        writer.close()
        return CodeShard(path)

    def export_pickles(self, dump_dir, split, start_index=0):
        """The reference's own layout (``{split}/NNNNN.pkl`` + ``metadata.json``, generate_vq_dataset.py:86-89,216-220),
        for consumers that must stay on ``datasets/vqlatent.py`` unchanged."""
        os.makedirs(os.path.join(dump_dir, split), exist_ok=True)
        for i in range(len(self)):
            with open(os.path.join(dump_dir, split, f"{start_index + i:05d}.pkl"), "wb") as f:
                pickle.dump(self.item(i), f)
        with open(os.path.join(dump_dir, "metadata.json"), "w") as f:
            json.dump(self.metadata(), f)
