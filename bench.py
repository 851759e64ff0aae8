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


def make_batch(n_utt, seed):
    import torch
    from oracle import vq_oracle as O          # synthetic-input recipe only; not on the timed path
    gen = torch.Generator().manual_seed(seed)
    code = torch.randn(K_BINS, EMB, generator=gen)
    lengths = O.ljspeech_like_lengths(n_utt, gen)
    x, mask = O.synthetic_batch(lengths, EMB, gen, codebook=code)
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


def cpu_reference_rate(budget_s, batch_utt=8, seed=100):
    """Frames/s of the CPU port of the reference encode path, as shipped (NT x NT `fit` temporary included),
    and of the same path without that temporary.  Bounded to ~budget_s seconds."""
    import torch
    from oracle import vq_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    x, mask, lengths, code = make_batch(batch_utt, seed)
    st = O.CodebookState(K_BINS, EMB, k=code, init=True)
    valid = int(lengths.sum())
    out = {}
    for name, faithful in (("as_shipped", True), ("without_nxn_temp", False)):
        O.encode(st, x, mask, faithful_fit=faithful)            # warm-up
        t0, n = time.perf_counter(), 0
        while True:
            O.encode(st, x, mask, faithful_fit=faithful)
            n += 1
            if time.perf_counter() - t0 > budget_s / 2 or n >= 50:
                break
        out[name] = valid * n / (time.perf_counter() - t0)
    return out, torch.get_num_threads(), f"{batch_utt} utterances/batch ({valid} valid frames, {x.shape[0] * x.shape[2]} rows)"


def run_reference(args):
    rank, world = env_int("RANK", 0), env_int("WORLD_SIZE", 1)
    if rank != 0:
        return 0
    import torch
    from oracle import vq_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    st = None
    batches = []
    per_step_utts = 32                      # bounded sample of the 256-utterance step, in script-default batches of 8
    for b in range(per_step_utts // 8):
        x, mask, lengths, code = make_batch(8, 1000 + b)
        batches.append((x, mask, int(lengths.sum())))
        st = st or O.CodebookState(K_BINS, EMB, k=code, init=True)
    frames = sum(b[2] for b in batches)

    def step():
        for x, mask, _ in batches:
            O.encode(st, x, mask, faithful_fit=True)

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
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / done, "steps_timed": done, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ljspeech-like corpus encode (BottleneckBlock.encode), K=512 D=128, CPU port of the reference",
                   "k_bins": K_BINS, "emb_width": EMB, "utterances_per_step": per_step_utts, "batch_size": 8,
                   "frames_per_step": frames},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
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
    signal.alarm(env_int("VQ_BENCH_TIMEOUT", 540))      # a wedged collective must not eat the GPU lease
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

    # ---- this rank's shard of the corpus: one 256-utterance batch per step (weak scaling)
    x, mask, lengths, code = make_batch(UTT_PER_STEP, seed=rank)
    n, d, t = x.shape
    valid_frames, rows = int(lengths.sum()), n * t
    xd, kd = x.to(dev), code.to(dev)
    idx = torch.empty(n, t, dtype=torch.int64, device=dev)
    ws = torch.empty(int(lib.vq_workspace_bytes(n, t, K_BINS, EMB)), dtype=torch.uint8, device=dev)
    scalars = torch.zeros(16, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    prepared = [0]

    def step():
        # frozen codebook, many batches (generate_vq_dataset.py): the codebook operands are prepared by the first call only
        rc = lib.vq_assign(xd.data_ptr(), n, d, t, kd.data_ptr(), K_BINS, idx.data_ptr(), None, None,
                           ws.data_ptr(), ws.numel(), prepared[0], stream)
        prepared[0] = 256
        if rc:
            raise RuntimeError(lib.vq_last_error().decode())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # per-kernel times of the same step from a second, short pass with the library's event profiler on (its event records
    # between the kernels would otherwise sit inside the timed region and serialise the dependent launch of the re-scan)
    lib.vq_profile_enable(1)
    for _ in range(min(args.steps, 32)):
        step()
    torch.cuda.synchronize()
    prof = (ctypes.c_float * 4)()
    have_prof = lib.vq_profile_read(prof) == 0
    lib.vq_profile_enable(0)
    # one untimed call with the scalar block attached: how many frames took the exact fallback
    lib.vq_assign(xd.data_ptr(), n, d, t, kd.data_ptr(), K_BINS, idx.data_ptr(), None, scalars.data_ptr(),
                  ws.data_ptr(), ws.numel(), 0, stream)
    unsafe = float(scalars[vqb200._lib.S_UNSAFE_ROWS].item())
    if args.profile_only:
        print(json.dumps({"profile_only": True, "ms_per_step": ms / args.steps, "k1_ms": float(prof[1])}))
        return 0

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
    clocks = sampler.stop() if rank == 0 else None
    idx_host = torch.frombuffer((ctypes.c_int64 * rows).from_address(hidx), dtype=torch.int64).clone()
    lib.vq_host_ctx_destroy(ctx)

    # whole training-mode forward of the module on every rank (K1 + K2 + K3a + ONE all-reduce + K3b + restart-row glue)
    def timed_all(fn, reps=5):
        for _ in range(2):
            fn()
        barrier()
        best = 1e9
        for _ in range(3):                       # `reps` back-to-back forwards per measurement: launches overlap execution
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
    blk.k, blk.k_sum, blk.k_elem, blk.init = kd.clone(), kd.clone() * 4, torch.full((K_BINS,), 4.0, device=dev), True
    blk.train()
    fwd_ms = timed_all(lambda: blk(xd, md_all, update_k=True))
    blk.rng_parity = False                   # restart rows drawn on the device: no host sync, no CPU randperm
    fwd_fast_ms = timed_all(lambda: blk(xd, md_all, update_k=True))
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
        fwd_graph_ms = timed_all(graph.replay)
        del graph, g_out
    except Exception as exc:                                  # noqa: BLE001  (reported, never fatal for the bench line)
        fwd_graph_ms = None
        graph_error = repr(exc)[:200]
    del blk

    times = torch.tensor([ms, e2e_s * 1e3, fwd_ms, fwd_fast_ms, fwd_graph_ms if fwd_graph_ms else 0.0], dtype=torch.float64, device=dev)
    frames = torch.tensor([float(valid_frames), float(rows)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, dist.ReduceOp.MAX)
        dist.all_reduce(frames, dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    ms, e2e_ms, fwd_ms, fwd_fast_ms, fwd_graph_ms = (float(times[i]) for i in range(5))
    tot_valid, tot_rows = float(frames[0]), float(frames[1])
    value = tot_valid * args.steps / (ms * 1e-3)
    e2e_value = tot_valid * args.steps / (e2e_ms * 1e-3)

    # ---- parity spot check against the oracle (outside every timed region)
    sample = 4
    rows_cpu, _, _ = O.flatten_nct(x[:sample], mask[:sample])
    o_l, _, _ = O.assign(rows_cpu, code)
    audit = O.audit_indices(rows_cpu, code, o_l, idx[:sample].cpu().reshape(-1))
    audit_host = O.audit_indices(rows_cpu, code, o_l, idx_host.view(n, t)[:sample].reshape(-1))

    # ---- roofline of the dominant kernel (K1): algorithmic flops = 2 * rows * K * D per launch
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    peak_tf = float(peaks.get("bf16_tflops", 1590.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst)" if peaks else "fallback 1590 TFLOP/s"
    k1_ms = float(prof[1]) if have_prof and prof[1] > 0 else ms / args.steps
    flops = 2.0 * rows * K_BINS * EMB
    achieved = flops / (k1_ms * 1e-3) / 1e12
    hbm_bytes = rows * (4 * EMB + 8) + 4 * K_BINS * EMB
    traffic = None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture
        traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_k1_traffic.json")))["traffic_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": traffic, "algorithmic_flop_per_launch": flops, "algorithmic_bytes_per_launch": hbm_bytes, "peak_source": peak_src, "kernel": "K1 distance+argmin (vq_assign main kernel)",
                "kernel_ms": k1_ms, "kernel_ms_source": "library CUDA events around the kernel" if have_prof and prof[1] > 0
                else "whole vq_assign step (prep + main + fallback kernels)",
                "hbm_secondary": {"achieved_GBps": hbm_bytes / (k1_ms * 1e-3) / 1e9, "peak_GBps": float(peaks.get("hbm_gbs", 6650.0))}}
    if have_prof:
        roofline["step_breakdown_ms"] = {"codebook_prepare": float(prof[0]), "assign_main": float(prof[1]),
                                         "exact_fallback": float(prof[2])}

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
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    x_q = torch.empty_like(xd)
    res = torch.zeros(8, device=dev)
    stats = torch.zeros(K_BINS * EMB + K_BINS, device=dev)
    g_commit = torch.ones((), device=dev)
    k3_scratch = torch.empty(n * ((t + 63) // 64), dtype=torch.uint8, device=dev)
    k_sum, k_elem, k_new = kd.clone(), torch.ones(K_BINS, device=dev), torch.empty_like(kd)
    k2_fwd = timed(lambda: lib.vq_gather_st_fwd(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), n, d, t, K_BINS,
                                                x_q.data_ptr(), scalars.data_ptr(), res.data_ptr(), stream))
    k2_bwd = timed(lambda: lib.vq_gather_st_bwd(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), kd.data_ptr(), x_q.data_ptr(),
                                                g_commit.data_ptr(), scalars.data_ptr(), n, d, t, K_BINS, x_q.data_ptr(), stream))
    k2_dec = timed(lambda: lib.vq_decode(idx.data_ptr(), kd.data_ptr(), n, d, t, K_BINS, x_q.data_ptr(), stream))
    k3_acc = timed(lambda: lib.vq_ema_accumulate(xd.data_ptr(), idx.data_ptr(), md.data_ptr(), n, d, t, K_BINS, stats.data_ptr(), k3_scratch.data_ptr(), stream))
    k3_fin = timed(lambda: lib.vq_ema_finalize(stats.data_ptr(), kd.data_ptr(), kd.data_ptr(), k_new.data_ptr(), k_sum.data_ptr(),
                                               k_elem.data_ptr(), K_BINS, EMB, 0.99, 1.0, 0.0, scalars.data_ptr(), res.data_ptr(), None, stream))

    def gbps(nbytes, ms_):
        return nbytes / (ms_ * 1e-3) / 1e9

    other = {
        "K2_gather_st_fwd": {"ms": k2_fwd, "algorithmic_bytes": rows * (8 * EMB + 12), "GBps": gbps(rows * (8 * EMB + 12), k2_fwd)},
        "K2_gather_st_bwd": {"ms": k2_bwd, "algorithmic_bytes": rows * (12 * EMB + 12), "GBps": gbps(rows * (12 * EMB + 12), k2_bwd)},
        "K2_decode": {"ms": k2_dec, "algorithmic_bytes": rows * (4 * EMB + 8), "GBps": gbps(rows * (4 * EMB + 8), k2_dec)},
        "K3_ema_accumulate": {"ms": k3_acc, "algorithmic_bytes": valid_frames * (4 * EMB + 8) + rows * 4 + 4 * K_BINS * (EMB + 1),
                              "GBps": gbps(valid_frames * (4 * EMB + 8) + rows * 4 + 4 * K_BINS * (EMB + 1), k3_acc)},
        "K3_ema_finalize": {"ms": k3_fin, "algorithmic_bytes": 4 * K_BINS * (6 * EMB + 3), "GBps": gbps(4 * K_BINS * (6 * EMB + 3), k3_fin)},
    }
    for v in other.values():
        v["frac_of_hbm_peak"] = v["GBps"] / hbm_peak
    other["module_forward_train"] = {"ms": fwd_ms, "valid_frames_per_s": tot_valid / (fwd_ms * 1e-3),
                                     "note": "whole nn.Module training forward per rank (K1+K2+K3a+all-reduce+K3b+restart-row glue), max over ranks"}
    other["module_forward_train_device_rng"] = {"ms": fwd_fast_ms, "valid_frames_per_s": tot_valid / (fwd_fast_ms * 1e-3),
                                                "note": "same with rng_parity=False (restart rows drawn on the device, no host sync)"}
    if fwd_graph_ms:
        other["module_forward_train_cuda_graph"] = {"ms": fwd_graph_ms, "valid_frames_per_s": tot_valid / (fwd_graph_ms * 1e-3),
                                                    "note": "rng_parity=False forward captured once with torch.cuda.graph and replayed"}

    cpu, cores, sample_desc = cpu_reference_rate(budget_s=16.0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 tensor-core shortlist + f32 exact rescoring/fallback", "data": "synthetic",
        "config": {"workload": "ljspeech-like corpus encode (BottleneckBlock.encode / generate_vq_dataset.py:69), K=512 D=128",
                   "k_bins": K_BINS, "emb_width": EMB, "utterances_per_step_per_gpu": UTT_PER_STEP,
                   "rows_per_step_per_gpu": rows, "valid_frames_per_step_per_gpu": valid_frames, "layout": "NCT fp32",
                   "l2": "inputs (228 MB per step) exceed the 126 MB L2; no flush needed", "parallelism": f"frames sharded x{world}, codebook replicated"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": rows * d * 4, "d2h_bytes_per_step": rows * 8,
                "ms_per_step": e2e_ms / args.steps, "api": "vq_encode_host (pinned host buffers, chunked double-buffered copies)",
                "timer": "host wall clock around synchronous calls"},
        "gpu_launches": args.steps * 2,   # assign_tc + exact-fallback kernel per step (the codebook is prepared once, before the loop)
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu["as_shipped"], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc,
                         "without_nxn_temp": cpu["without_nxn_temp"]},
        "index_match": {"device_path": audit, "host_path": audit_host},
        "training_path_kernels": other,
        "unsafe_rows_per_step": unsafe,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
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
