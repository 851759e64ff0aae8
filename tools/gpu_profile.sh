#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, full captures of K1 (easy variant), K2 fwd, K3a, re-scan kernel
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_ll.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:assign_tc -s 4 -c 1 -o gpurun_out/prof_k1 python bench.py --steps 3 --warmup 3 --profile-only > gpurun_out/ncu_k1.log 2>&1
K23_ONLY=k2_fwd K23_REPS=2 ncu --set full --clock-control none --import-source on -f -k regex:gather_async_kernel -s 3 -c 1 -o gpurun_out/prof_k23 python tools/k23_bench.py > gpurun_out/ncu_k23.log 2>&1
K23_ONLY=k3a K23_REPS=2 ncu --set full --clock-control none --import-source on -f -k regex:ema_accumulate_runs -s 3 -c 1 -o gpurun_out/prof_k3a python tools/k23_bench.py > gpurun_out/ncu_k3a.log 2>&1
VQ_K1_HARD=1 ncu --set full --clock-control none --import-source on -f -k regex:assign_list -c 1 -o gpurun_out/prof_list python tools/list_debug.py > gpurun_out/ncu_list.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
