#!/bin/bash
mkdir -p gpurun_out
python tools/ab_k1.py speech-masters-thesis_b200/lib/libvqb200.so ab/libvqb200_prev.so 2>&1 | tail -1 | tee gpurun_out/experiment_foldtpl.log
python tools/ab_k1.py speech-masters-thesis_b200/lib/libvqb200.so ab/libvqb200_prev.so gaussian 2>&1 | tail -1 | tee -a gpurun_out/experiment_foldtpl.log
timeout 900 python -m pytest tests/test_gpu_audit.py tests/test_gpu_parity.py -m gpu -x -q -k "every_row or sweep_corner or golden or random or tcgen05_route or bf16" 2>&1 | tail -3
timeout 600 python tools/sweep.py --quick > gpurun_out/sweepq.log 2>&1; echo "sweep rc=$?"
python - <<'PY'
import json
for p in [q for q in json.load(open('gpurun_out/sweep.json'))['results'] if 'frac_of_peak_main' in q]: print(p['K'],p['D'],round(p['frac_of_peak_main'],3),round(p['assign_main_ms'],4),p['check'].get('mismatches'))
PY
