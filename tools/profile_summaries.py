"""Turn the ncu reports / launch list in gpurun_out/ into the committed summaries under profiles/ (run in the authoring
container: `ncu -i` needs no GPU).  python tools/profile_summaries.py r02"""
import csv
import io
import json
import os
import subprocess
import sys
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def summarise(rep, command, out_name):
    hdr, units, rows = raw_rows(rep)
    res = []
    for r in rows:
        m = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                m[w] = f"{r[i]} {units[i]}".strip()
        res.append({"kernel": r[hdr.index("Kernel Name")], "metrics": m})
    json.dump({"report": os.path.relpath(rep, ROOT), "command": command, "launches": res},
              open(os.path.join(ROOT, "profiles", out_name), "w"), indent=1)
    return res


def launch_list(path, command, out_name):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= iv:
            continue
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[ik], [0, 0.0])
        a[0] += 1
        a[1] += v
    unit = rows[start + 1][hdr.index("Metric Unit")] if len(rows) > start + 1 else "ns"
    scale = {"ns": 1.0, "us": 1e3, "usecond": 1e3, "nsecond": 1.0, "ms": 1e6, "msecond": 1e6}.get(unit, 1.0)
    total = sum(a[1] for a in agg.values()) * scale
    kernels = [{"kernel": k, "launches": a[0], "total_ns": a[1] * scale, "avg_ns": a[1] * scale / a[0], "share": a[1] * scale / total}
               for k, a in agg.items()]
    kernels.sort(key=lambda e: -e["total_ns"])
    json.dump({"command": command, "note": "cold-cache serialised per-launch times: compare shares, not absolutes", "kernels": kernels},
              open(os.path.join(ROOT, "profiles", out_name), "w"), indent=1)
    return kernels


if __name__ == "__main__":
    g = os.path.join(ROOT, "gpurun_out")
    k1 = summarise(os.path.join(g, "prof_k1.ncu-rep"),
                   "ncu --set full --clock-control none --import-source on -k regex:assign_tc -s 4 -c 1 python bench.py --steps 3 --warmup 3 --profile-only",
                   f"{TAG}_k1_full_summary.json")
    m = k1[0]["metrics"]
    rd = float(m["dram__bytes_read.sum"].split()[0]) * (1e6 if "Mbyte" in m["dram__bytes_read.sum"] else 1.0)
    wr = float(m["dram__bytes_write.sum"].split()[0]) * (1e6 if "Mbyte" in m["dram__bytes_write.sum"] else 1.0)
    json.dump({"kernel": k1[0]["kernel"], "source": f"profiles/{TAG}_k1_full.ncu-rep (ncu --set full --clock-control none, one launch of the bench workload)",
               "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "traffic_bytes_per_launch": int(rd + wr),
               "algorithmic_bytes_per_launch": 444416 * (4 * 128 + 8) + 4 * 512 * 128},
              open(os.path.join(ROOT, "profiles", f"{TAG}_k1_traffic.json"), "w"), indent=1)
    summarise(os.path.join(g, "prof_k23.ncu-rep"),
              "K23_ONLY=k2_fwd,k3a K23_REPS=2 ncu --set full --clock-control none -k regex:'gather_async_kernel|ema_accumulate_runs' -s 6 -c 2 python tools/k23_bench.py",
              f"{TAG}_k2_k3_full_summary.json")
    if os.path.exists(os.path.join(g, "prof_k3a.ncu-rep")):
        summarise(os.path.join(g, "prof_k3a.ncu-rep"),
                  "K23_ONLY=k3a K23_REPS=2 ncu --set full --clock-control none -k regex:ema_accumulate_runs -s 3 -c 1 python tools/k23_bench.py",
                  f"{TAG}_k3a_full_summary.json")
    summarise(os.path.join(g, "prof_list.ncu-rep"),
              "ncu --set full --clock-control none -k regex:assign_list -c 1 python tools/list_debug.py   (i.i.d. Gaussian batch, 5576 frames re-scanned)",
              f"{TAG}_rescan_full_summary.json")
    if os.path.exists(os.path.join(g, "prof_fused.ncu-rep")):
        summarise(os.path.join(g, "prof_fused.ncu-rep"), "K23_ONLY=k2_fused ncu --set full -k regex:gather_fwd_ema python tools/k23_bench.py  (v3, owner-computes)",
                  f"{TAG}_k2_fused_ema_full_summary.json")
    ll = launch_list(os.path.join(g, "launches.csv"), "ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3",
                     f"{TAG}_launch_list.json")
    for e in ll[:8]:
        print(e["kernel"][:70], e["launches"], round(e["avg_ns"] / 1e3, 2), "us", round(e["share"], 3))
    print(json.dumps(k1[0]["metrics"], indent=1))
