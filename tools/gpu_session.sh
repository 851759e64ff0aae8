#!/bin/bash
# one GPU-box session: tests, K1 timing experiments, A/B, bench, ncu captures; everything lands in gpurun_out/
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/k23_bench.py 2>&1 | tail -1 | tee gpurun_out/k23.log
python tools/ab_k1.py ab/libvqb200_r1.so speech-masters-thesis_b200/lib/libvqb200.so 2>&1 | tail -1 | tee gpurun_out/ab.log
EXPERIMENTS="${EXPERIMENTS:-2 3 10 11 0}" timeout 900 bash tools/experiment.sh > gpurun_out/e.log 2>&1; cat gpurun_out/e.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'k1_ms',d['roofline']['kernel_ms'])
print('gaussian',{k:v for k,v in d['gaussian'].items() if k!='index_match'})
for r in d['rooflines']: print(r['kernel'], round(r['ms'],4), round(r['frac'],3))
print('training',d['training_path'])
PY
# ncu (one tool per call): the fused K2+K3a kernel, the exact re-scan kernel, K1
K23_ONLY=k2_fused K23_REPS=2 ncu --set full --clock-control none --import-source on -k regex:gather_fwd_ema -s 2 -c 2 -o gpurun_out/prof_fused python tools/k23_bench.py > gpurun_out/ncu_fused.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assign_list -c 1 -o gpurun_out/prof_list python tools/list_debug.py > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assign_tc -s 4 -c 1 -o gpurun_out/prof_k1 python bench.py --steps 3 --warmup 3 --profile-only > gpurun_out/ncu_k1.log 2>&1
ls -la gpurun_out/*.ncu-rep
