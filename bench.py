#!/usr/bin/env python
"""bench.py -- VQ frames/s of the B200 quantiser on BASELINE.json's workload.

Workload (config.workload): ``generate_vq_dataset``-style encode (BASELINE.json configs[1]) of a synthetic
LJSpeech-like corpus at the repo's default quantiser shape K=512, D=128: one STEP = K1 (distance + argmin,
``BottleneckBlock.encode``) over one batch of 256 utterances padded to the batch maximum (about 0.44 M rows,
228 MB of FP32 latents -- larger than the 126 MB L2, so consecutive steps cannot be served from cache).
``value`` counts VALID (unpadded) frames; padded rows are computed too, exactly like the reference does.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1: launched by torchrun, one rank per GPU; every rank encodes its own batch (utterance-sharded corpus,
codebook replicated, no collective on the encode path) -> weak scaling; time = max over ranks.
--impl reference: the CPU port of the reference path (oracle/vq_oracle.py, same torch CPU ops as the reference,
including its NT x NT `fit` temporary) on the host cores, batch size 8 like the script's default.
"""
import argparse
import ctypes
import json
import os
import signal
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K_BINS, EMB = 512, 128
UTT_PER_STEP = 256
METRIC = "vq_frames_per_s"
UNIT = "frames/s"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def make_batch(n_utt, seed, clustered=True):
    """SURVEY.md 8d inputs: (C) clustered latents E[j] + 0.5 N(0,1), Zipf-skewed j (speech-like) or (G) i.i.d. N(0,1)
    (the adversarial case for near-ties)."""
    import torch
    from oracle import vq_oracle as O          # synthetic-input recipe only; not on the timed path
    gen = torch.Generator().manual_seed(seed)
    torch.randn(K_BINS, EMB, generator=gen)     # (keeps rank 0's batch what it always was)
    # ONE codebook for every rank: a rank whose latents cluster around another codebook than the one it quantises against
    # (the replicas share rank 0's, bottleneck.py:41) sees an adversarial batch and drags the whole job onto the re-scan path
    code = torch.randn(K_BINS, EMB, generator=torch.Generator().manual_seed(0))
    lengths = O.ljspeech_like_lengths(n_utt, gen)
    x, mask = O.synthetic_batch(lengths, EMB, gen, codebook=code if clustered else None)
    return x, mask, lengths, code


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [s.strip() for s in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def reference_block(code):
    """The CPU arm's quantiser: the UNMODIFIED reference ``BottleneckBlock`` when baseline/_ref is staged
    (kind "reference"), else the oracle port (kind "port").  Returns (kind, encode(x, mask), forward_train(x, mask))."""
    import torch
    from oracle import stage_ref, vq_oracle as O
    if stage_ref.activate():
        from models.vqvae.bottleneck import BottleneckBlock
        blk = BottleneckBlock(K_BINS, EMB, 0.99, 1.0)
        blk.k = code.clone()
        blk.k_sum, blk.k_elem, blk.init = code.clone() * 4, torch.full((K_BINS,), 4.0), True
        blk.train()

        def fwd(x, mask):
            with torch.no_grad():
                return blk(x, mask, update_k=True)

        return "reference", (lambda x, mask: blk.encode(x, mask)), fwd
    st = O.CodebookState(K_BINS, EMB, 0.99, 1.0, code.clone(), code.clone() * 4, torch.full((K_BINS,), 4.0), True)
    return ("port", (lambda x, mask: O.encode(st, x, mask, faithful_fit=True)),
            (lambda x, mask: O.forward(st, x, mask, update_k=True, faithful_fit=True)))


def _rate(fn, frames, budget_s, max_reps=50):
    fn()                                                        # warm-up
    t0, n = time.perf_counter(), 0
    while True:
        fn()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= max_reps:
            break
    return frames * n / (time.perf_counter() - t0)


def cpu_reference_rates(budget_s):
    """Frames/s of the reference's CPU path on this box's host cores, bounded to ~budget_s seconds:
    BASELINE.json configs[0] (forward + EMA update, batch 16), the generate-script encode (batch 8, script default), and
    the port's encode without the NT x NT `fit` temporary (so that artefact is visible separately)."""
    import torch
    from oracle import vq_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    x8, m8, l8, code = make_batch(8, 100)
    x16, m16, l16, _ = make_batch(16, 101)
    kind, enc, fwd = reference_block(code)
    st = O.CodebookState(K_BINS, EMB, k=code, init=True)
    out = {
        "forward_train_b16": _rate(lambda: fwd(x16, m16), int(l16.sum()), budget_s / 3),
        "encode_b8": _rate(lambda: enc(x8, m8), int(l8.sum()), budget_s / 3),
        "encode_b8_without_nxn_temp": _rate(lambda: O.encode(st, x8, m8, faithful_fit=False), int(l8.sum()), budget_s / 3),
    }
    sample = (f"encode: 8 utterances/batch ({int(l8.sum())} valid frames, {x8.shape[0] * x8.shape[2]} rows); "
              f"forward+EMA (configs[0]): 16 utterances ({int(l16.sum())} valid frames, {x16.shape[0] * x16.shape[2]} rows)")
    return out, kind, torch.get_num_threads(), sample


def run_reference(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    batches = []
    per_step_utts = 32                      # bounded sample of the 256-utterance step, in script-default batches of 8
    code = None
    for b in range(per_step_utts // 8):
        x, mask, lengths, c = make_batch(8, 1000 + b)
        code = c if code is None else code
        batches.append((x, mask, int(lengths.sum())))
    kind, enc, _ = reference_block(code)
    frames = sum(b[2] for b in batches)

    def step():
        for x, mask, _ in batches:
            enc(x, mask)

    for _ in range(min(args.warmup, 2)):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):          # each step is ~2 s of host work: stop after ~90 s so the arm ends within minutes
        step()
        done += 1
        if time.perf_counter() - t0 > 90.0:
            break
    dt = time.perf_counter() - t0
    value = frames * done / dt
    what = ("the unmodified reference BottleneckBlock.encode (baseline/_ref)" if kind == "reference"
            else "CPU port of the reference (oracle/vq_oracle.py)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "steps_timed": done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ljspeech-like corpus encode (BottleneckBlock.encode), K=512 D=128, " + what,
                   "k_bins": K_BINS, "emb_width": EMB, "utterances_per_step": per_step_utts, "batch_size": 8,
                   "frames_per_step": frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
                         "sample": f"{per_step_utts} utterances per step in batches of 8 (script default), as shipped "
                                   "(NT x NT fit temporary included)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    signal.alarm(env_int("VQ_BENCH_TIMEOUT", 900))      # a wedged collective must not eat the GPU lease
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    import vqb200
    lib = vqb200._lib.load()
    from oracle import vq_oracle as O
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- this rank's shard of the corpus: one 256-utterance batch per step (weak scaling); (C) clustered latents are the
    # headline, (G) i.i.d. Gaussian latents the adversarial case (2 % of the frames take the exact re-scan)
    class Job:
        def __init__(self, clustered):
            self.x, self.mask, self.lengths, self.code = make_batch(UTT_PER_STEP, seed=rank, clustered=clustered)
            self.n, self.d, self.t = self.x.shape
            self.valid, self.rows = int(self.lengths.sum()), self.n * self.t
            self.xd, self.kd = self.x.to(dev), self.code.to(dev)
            self.idx = torch.empty(self.n, self.t, dtype=torch.int64, device=dev)
            self.ws = torch.empty(int(lib.vq_workspace_bytes(self.n, self.t, K_BINS, EMB)), dtype=torch.uint8, device=dev)
            self.prepared = 0

        def step(self):
            # frozen codebook, many batches (generate_vq_dataset.py): the codebook operands are prepared by the first call only
            rc = lib.vq_assign(self.xd.data_ptr(), self.n, self.d, self.t, self.kd.data_ptr(), K_BINS, self.idx.data_ptr(), None,
                               None, self.ws.data_ptr(), self.ws.numel(), self.prepared, stream)
            self.prepared = 256
            if rc:
                raise RuntimeError(lib.vq_last_error().decode())

        def timed(self, steps, warmup):
            for _ in range(max(3, warmup)):
                self.step()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                self.step()
            e1.record()
            barrier()
            return e0.elapsed_time(e1)

        def profile(self, steps):
            # per-kernel times from a second, short pass with the library's event profiler on (its event records between the
            # kernels would otherwise sit inside the timed region and serialise the dependent launch of the re-scan)
            lib.vq_profile_enable(1)
            for _ in range(min(steps, 32)):
                self.step()
            torch.cuda.synchronize()
            prof = (ctypes.c_float * 4)()
            ok = lib.vq_profile_read(prof) == 0
            lib.vq_profile_enable(0)
            return [float(v) for v in prof] if ok else None

        def unsafe_rows(self):
            scalars = torch.zeros(16, dtype=torch.float64, device=dev)
            lib.vq_assign(self.xd.data_ptr(), self.n, self.d, self.t, self.kd.data_ptr(), K_BINS, self.idx.data_ptr(), None,
                          scalars.data_ptr(), self.ws.data_ptr(), self.ws.numel(), 0, stream)
            self.prepared = 0
            return float(scalars[vqb200._lib.S_UNSAFE_ROWS].item())

    job = Job(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = job.timed(args.steps, args.warmup)
    prof = job.profile(args.steps)
    unsafe = job.unsafe_rows()
    if args.profile_only:
        print(json.dumps({"profile_only": True, "ms_per_step": ms / args.steps, "k1_ms": prof[1] if prof else None}))
        return 0
    # sustained: the same step back to back for >= 1 s (the headline region is only steps x 0.07 ms)
    sus_steps = max(args.steps, int(1.2e3 / max(ms / args.steps, 1e-3)))
    sus_ms = job.timed(sus_steps, 0)
    clocks = sampler.stop() if rank == 0 else None
    gjob = Job(False)
    g_ms = gjob.timed(args.steps, args.warmup)
    g_prof = gjob.profile(args.steps)
    g_unsafe = gjob.unsafe_rows()
    n, d, t, rows, valid_frames = job.n, job.d, job.t, job.rows, job.valid
    x, mask, code, xd, kd, idx = job.x, job.mask, job.code, job.xd, job.kd, job.idx

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside the timed region)
    ctx = lib.vq_host_ctx_create(local, rows, K_BINS, EMB)
    if not ctx:
        raise RuntimeError(lib.vq_last_error().decode())
    lib.vq_host_ctx_set_codebook(ctx, code.numpy().ctypes.data)
    hx, hidx = lib.vq_host_ctx_x_staging(ctx), lib.vq_host_ctx_idx_staging(ctx)
    ctypes.memmove(hx, x.numpy().ctypes.data, rows * d * 4)
    for _ in range(2):
        lib.vq_encode_host(ctx, hx, n, t, hidx, None)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        if lib.vq_encode_host(ctx, hx, n, t, hidx, None):
            raise RuntimeError(lib.vq_last_error().decode())
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    idx_host = torch.frombuffer((ctypes.c_int64 * rows).from_address(hidx), dtype=torch.int64).clone()
    lib.vq_host_ctx_destroy(ctx)

    # whole training-mode forward of the module on every rank (K1 + K3a + {ONE all-reduce || K2} + K3b + restart-row glue)
    def timed_all(fn, reps=5, rounds=3):
        for _ in range(3):
            fn()
        barrier()
        best = 1e9
        for _ in range(rounds):                  # `reps` back-to-back forwards per measurement: launches overlap execution
            if world > 1:
                barrier()                        # all ranks start a measurement together
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / reps)
        return best

    md_all = mask.to(dev)
    blk = vqb200.BottleneckBlock(K_BINS, EMB, 0.99, 1.0).to(dev)
    k0 = kd.clone()
    if world > 1:
        dist.broadcast(k0, 0)                 # one codebook for all replicas (every rank's synthetic batch has its own)
    blk.k, blk.k_sum, blk.k_elem, blk.init = k0, k0.clone() * 4, torch.full((K_BINS,), 4.0, device=dev), True
    blk.train()
    fwd_ms = timed_all(lambda: blk(xd, md_all, update_k=True), reps=10)
    blk.rng_parity = False                   # restart rows drawn on the device: no host sync, no CPU randperm
    # 20 back to back, best of 5: with fewer, the start-up skew between the ranks (each leaves its own synchronize at its own
    # time, and a step cannot finish before the slowest rank has published its statistics) dominates a 0.35 ms step
    fwd_fast_ms = timed_all(lambda: blk(xd, md_all, update_k=True), reps=20, rounds=5)
    # replicas must still be bit-identical after all those all-reduced updates (bottleneck.py:72-75)
    replicas_identical = None
    if world > 1:
        digest = torch.stack([blk.k.double().sum(), blk.k_sum.double().sum(), blk.k_elem.double().sum(),
                              (blk.k.double() ** 2).sum()])
        all_digests = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(all_digests, digest)
        replicas_identical = all(torch.equal(all_digests[0], g) for g in all_digests)
        assert replicas_identical, "codebook replicas diverged across ranks"
    # the same forward replayed from a CUDA graph (no host launch cost; static input addresses, as any graphed module)
    fwd_graph_ms = None
    try:
        if world > 1:
            raise RuntimeError("single-GPU only (the forward contains an NCCL all-reduce when world_size > 1)")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(3):
                blk(xd, md_all, update_k=True)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(graph):
            g_out = blk(xd, md_all, update_k=True)
        fwd_graph_ms = timed_all(graph.replay, reps=20, rounds=5)
        del graph, g_out
    except Exception:                                          # noqa: BLE001  (reported as null, never fatal for the bench line)
        fwd_graph_ms = None
    del blk

    # ---- BASELINE.json configs[4]: quantise + index gather for 256 utterances IN TOTAL (256 / world per GPU), through the module
    # API (BottleneckBlock.encode + .decode: what TransformerLM.reconstruct / sample consume, transformer_lm.py:101-108)
    per_rank = max(1, UTT_PER_STEP // world)
    eblk = vqb200.BottleneckBlock(K_BINS, EMB, 0.99, 1.0).to(dev)
    eblk.k, eblk.init = kd.clone(), True
    eblk.eval()
    x5, m5 = xd[:per_rank].contiguous(), md_all[:per_rank].contiguous()
    codes5 = int(job.lengths[:per_rank].sum())

    def quantise_and_gather():
        with torch.no_grad():
            return eblk.decode(eblk.encode(x5, m5))

    c5_ms = timed_all(quantise_and_gather, reps=10)
    # ---- the phoneme-conditioned quantiser at the TTS config's codebook shape (149 x 512 codes, D = 128), 16 utterances x 800 frames
    gg = torch.Generator().manual_seed(1234 + rank)
    g_nv, g_lb, g_n, g_t = 149, 512, 16, 800
    g_code = torch.randn(g_nv * g_lb, EMB, generator=gg).to(dev)
    g_x = torch.randn(g_n, EMB, g_t, generator=gg).to(dev)
    g_tok = torch.randint(0, g_nv, (g_n, g_t), generator=gg).to(dev)
    grouped_ms = timed_all(lambda: vqb200.assign_grouped(g_x, g_code, g_tok, g_nv, g_lb), reps=5)
    del g_code, g_x, g_tok

    times = torch.tensor([ms, e2e_s * 1e3, fwd_ms, fwd_fast_ms, fwd_graph_ms if fwd_graph_ms else 0.0, sus_ms, g_ms, c5_ms, grouped_ms],
                         dtype=torch.float64, device=dev)
    frames = torch.tensor([float(valid_frames), float(rows), float(gjob.valid), float(gjob.rows), float(codes5)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, dist.ReduceOp.MAX)
        dist.all_reduce(frames, dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    if world > 1:
        dist.destroy_process_group()      # no collective after this point: the CPU baseline below runs the reference module
        # (it would otherwise try to broadcast CPU tensors over NCCL, bottleneck.py:72-75)
    ms, e2e_ms, fwd_ms, fwd_fast_ms, fwd_graph_ms, sus_ms, g_ms, c5_ms, grouped_ms = (float(times[i]) for i in range(9))
    tot_codes5 = float(frames[4])
    tot_valid, tot_rows, g_tot_valid = float(frames[0]), float(frames[1]), float(frames[2])
    value = tot_valid * args.steps / (ms * 1e-3)
    e2e_value = tot_valid * args.steps / (e2e_ms * 1e-3)

    # ---- parity audit against the oracle, EVERY row of both batches (outside every timed region): the oracle's
    # quantize(x, mask=None) over 8k-row chunks; each disagreement is classified in fp64 (near-tie or error)
    def full_audit(j, got):
        rows_cpu = j.x.permute(0, 2, 1).reshape(-1, j.d)
        o_l, _ = O.assign_chunked(rows_cpu, j.code)
        return O.audit_indices(rows_cpu, j.code, o_l, got.cpu().reshape(-1))

    audit = full_audit(job, idx)
    audit_host = full_audit(job, idx_host)
    audit_gauss = full_audit(gjob, gjob.idx)
    assert audit["rows"] == rows and audit["errors"] == 0 and audit_host["errors"] == 0 and audit_gauss["errors"] == 0, \
        (audit, audit_host, audit_gauss)

    # ---- roofline of the dominant kernel (K1): algorithmic flops = 2 * rows * K * D per launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    peak_sus = float(peaks.get("bf16_tflops_sustained", peak_tf))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590 TFLOP/s"
    have_prof = prof is not None and prof[1] > 0
    k1_ms = prof[1] if have_prof else ms / args.steps
    flops = 2.0 * rows * K_BINS * EMB
    achieved = flops / (k1_ms * 1e-3) / 1e12
    hbm_bytes = rows * (4 * EMB + 8) + 4 * K_BINS * EMB
    traffic = None
    for name in ("r02_k1_traffic.json", "r01_k1_traffic.json"):
        try:    # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", name)))["traffic_bytes_per_launch"]
            break
        except (OSError, ValueError, KeyError):
            pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "algorithmic_flop_per_launch": flops, "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peak_src,
                "kernel": "K1 distance+argmin (vq_assign main kernel), clustered latents",
                "kernel_ms": k1_ms, "kernel_ms_source": "library CUDA events around the kernel" if have_prof
                else "whole vq_assign step (prep + main + fallback kernels)",
                "hbm_secondary": {"achieved_GBps": hbm_bytes / (k1_ms * 1e-3) / 1e9, "peak_GBps": hbm_peak}}
    if have_prof:
        roofline["step_breakdown_ms"] = {"codebook_prepare": prof[0], "assign_main": prof[1], "exact_fallback": prof[2]}
    sus_step_ms = sus_ms / sus_steps
    roofline["sustained"] = {"steps": sus_steps, "seconds": sus_ms * 1e-3, "ms_per_step": sus_step_ms,
                             "tflops_whole_step": flops / (sus_step_ms * 1e-3) / 1e12,
                             "frac_of_burst_peak": flops / (sus_step_ms * 1e-3) / 1e12 / peak_tf,
                             "frac_of_sustained_peak": flops / (sus_step_ms * 1e-3) / 1e12 / peak_sus,
                             "sm_mhz_median": clocks.get("sm_mhz") if clocks else None}
    # the adversarial distribution as a first-class number: whole step = main kernel + exact re-scan of the unsafe frames
    g_flops = 2.0 * gjob.rows * K_BINS * EMB
    g_step_ms = g_ms / args.steps
    gaussian = {"value": g_tot_valid * args.steps / (g_ms * 1e-3), "unit": UNIT, "ms_per_step": g_step_ms,
                "unsafe_rows_per_step": g_unsafe, "unsafe_frac": g_unsafe / gjob.rows,
                "roofline_whole_step": {"achieved": g_flops / (g_step_ms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                                        "frac": g_flops / (g_step_ms * 1e-3) / 1e12 / peak_tf},
                "index_match": audit_gauss}
    if g_prof:
        gaussian["step_breakdown_ms"] = {"codebook_prepare": g_prof[0], "assign_main": g_prof[1], "exact_fallback": g_prof[2]}

    # ---- the other kernels of the training path at the same batch: each timed alone, CUDA events around 10 back-to-back
    # calls (the 228 MB input exceeds L2, so every call streams from HBM), best of 3
    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / reps)
        return best

    md = mask.to(dev)
    scalars = torch.zeros(16, dtype=torch.float64, device=dev)
    x_q = torch.empty_like(xd)
    res = torch.zeros(8, device=dev)
    stats = torch.zeros(K_BINS * EMB + K_BINS, device=dev)
    g_commit = torch.ones((), device=dev)
    k3_scratch = torch.empty(n * ((t + 63) // 64), dtype=torch.uint8, device=dev)
    k_sum, k_elem, k_new = kd.clone(), torch.ones(K_BINS, device=dev), torch.empty_like(kd)
    k2_fwd = timed(lambda: lib.vq_gather_st_fwd(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), n, d, t, K_BINS,
                                                x_q.data_ptr(), scalars.data_ptr(), res.data_ptr(), stream))
    stats_f = torch.zeros(K_BINS * EMB + K_BINS, device=dev)
    k2_fused = timed(lambda: lib.vq_gather_st_fwd_ema(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), n, d, t, K_BINS,
                                                      x_q.data_ptr(), scalars.data_ptr(), res.data_ptr(), stats_f.data_ptr(), stream))
    k2_bwd = timed(lambda: lib.vq_gather_st_bwd(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), x_q.data_ptr(),
                                                g_commit.data_ptr(), scalars.data_ptr(), n, d, t, K_BINS, x_q.data_ptr(), stream))
    k2_dec = timed(lambda: lib.vq_decode(idx.data_ptr(), kd.data_ptr(), n, d, t, K_BINS, x_q.data_ptr(), stream))
    k3_acc = timed(lambda: lib.vq_ema_accumulate(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), n, d, t, K_BINS, stats.data_ptr(), k3_scratch.data_ptr(), stream))
    k3_fin = timed(lambda: lib.vq_ema_finalize(stats.data_ptr(), kd.data_ptr(), kd.data_ptr(), k_new.data_ptr(), k_sum.data_ptr(),
                                               k_elem.data_ptr(), K_BINS, EMB, 0.99, 1.0, 0.0, scalars.data_ptr(), res.data_ptr(), None, stream))

    def hbm_line(name, ms_, nbytes):
        gbps = nbytes / (ms_ * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "ms": ms_, "algorithmic_bytes": nbytes, "achieved": gbps, "peak": hbm_peak,
                "unit": "GB/s", "frac": gbps / hbm_peak}

    rooflines = [
        hbm_line("K2 vq_gather_st_fwd", k2_fwd, rows * (8 * EMB + 12)),
        hbm_line("K2+K3a vq_gather_st_fwd_ema (fused, opt-in: slower than K2 + K3a apart)", k2_fused, rows * (8 * EMB + 12) + 4 * K_BINS * (EMB + 1)),
        hbm_line("K2 vq_gather_st_bwd", k2_bwd, rows * (12 * EMB + 12)),
        hbm_line("K2 vq_decode", k2_dec, rows * (4 * EMB + 8)),
        hbm_line("K3a vq_ema_accumulate", k3_acc, valid_frames * (4 * EMB + 8) + rows * 4 + 4 * K_BINS * (EMB + 1)),
        hbm_line("K3b vq_ema_finalize", k3_fin, 4 * K_BINS * (6 * EMB + 3)),
    ]
    training = {
        "module_forward_train": {"ms": fwd_ms, "valid_frames_per_s": tot_valid / (fwd_ms * 1e-3),
                                 "note": "whole nn.Module training forward per rank (K1+K3a+{all-reduce||K2}+K3b+restart-row glue, "
                                         "rng_parity=True: CPU randperm replay), max over ranks"},
        "module_forward_train_device_rng": {"ms": fwd_fast_ms, "valid_frames_per_s": tot_valid / (fwd_fast_ms * 1e-3),
                                            "note": "same with rng_parity=False (restart rows drawn on the device, no host sync)"},
        "replicas_bit_identical_after_allreduce": replicas_identical,
    }
    if fwd_graph_ms:
        training["module_forward_train_cuda_graph"] = {"ms": fwd_graph_ms, "valid_frames_per_s": tot_valid / (fwd_graph_ms * 1e-3),
                                                       "note": "rng_parity=False forward captured once with torch.cuda.graph and replayed"}

    cpu, cpu_kind, cores, sample_desc = cpu_reference_rates(budget_s=18.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 tensor-core shortlist + f32 exact rescoring/fallback", "data": "synthetic",
        "config": {"workload": "ljspeech-like corpus encode (BottleneckBlock.encode / generate_vq_dataset.py:69), K=512 D=128",
                   "k_bins": K_BINS, "emb_width": EMB, "utterances_per_step_per_gpu": UTT_PER_STEP,
                   "rows_per_step_per_gpu": rows, "valid_frames_per_step_per_gpu": valid_frames, "layout": "NCT fp32",
                   "latents": "clustered (speech-like); the i.i.d. Gaussian batch is reported under `gaussian`",
                   "l2": "inputs (228 MB per step) exceed the 126 MB L2; no flush needed", "parallelism": f"frames sharded x{world}, codebook replicated"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rows * d * 4, "d2h_bytes_per_step": rows * 8,
                "ms_per_step": e2e_ms / args.steps, "api": "vq_encode_host (pinned host buffers, chunked double-buffered copies)",
                "timer": "host wall clock around synchronous calls"},
        "gpu_launches": args.steps * 2,   # assign_tc + exact-fallback kernel per step (the codebook is prepared once, before the loop)
        "clocks": clocks,
        "roofline": roofline,
        "rooflines": rooflines,
        "gaussian": gaussian,
        "cpu_baseline": {"value": cpu["encode_b8"], "unit": UNIT, "cores": cores, "kind": cpu_kind, "sample": sample_desc,
                         "configs0_forward_train_b16": cpu["forward_train_b16"],
                         "encode_without_nxn_temp": cpu["encode_b8_without_nxn_temp"]},
        "index_match": {"device_path": audit, "host_path": audit_host, "gaussian": audit_gauss},
        "training_path": training,
        "config5_quantise_plus_gather": {"value": tot_codes5 / (c5_ms * 1e-3), "unit": "codes/s", "ms": c5_ms,
                                         "utterances_total": per_rank * world, "utterances_per_gpu": per_rank, "valid_codes_total": tot_codes5,
                                         "api": "BottleneckBlock.encode + BottleneckBlock.decode (K1 + decode gather), inputs resident, max over ranks"},
        "grouped_tts_quantiser": {"ms": grouped_ms, "frames_per_s_per_gpu": g_n * g_t / (grouped_ms * 1e-3), "n_vocab": g_nv, "l_bins": g_lb,
                                  "emb_width": EMB, "frames": g_n * g_t, "api": "vq_assign_grouped (exact FP32, one warp per frame)"},
        "unsafe_rows_per_step": unsafe,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--profile-only", action="store_true", help="device-timed loop only (for ncu captures)")
    args = ap.parse_args()
    return run_reference(args) if args.impl == "reference" else run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
