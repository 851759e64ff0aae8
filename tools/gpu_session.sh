#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python tools/ab_k1.py ab/libvqb200_r1.so speech-masters-thesis_b200/lib/libvqb200.so 2>&1 | tail -1 | tee gpurun_out/ab.log
python tools/ab_k1.py ab/libvqb200_r1.so speech-masters-thesis_b200/lib/libvqb200.so gaussian 2>&1 | tail -1 | tee -a gpurun_out/ab.log
for v in "VQ_K1_HARD=0" "VQ_K1_HARD=1" "VQ_K1_HARD=0"; do echo "$v: $(env $v python bench.py --steps 50 --warmup 5 --profile-only 2>&1 | tail -1)"; done | tee -a gpurun_out/ab.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'k1_ms',d['roofline']['kernel_ms'])
print('gaussian',{k:v for k,v in d['gaussian'].items() if k!='index_match'})
for r in d['rooflines']: print(r['kernel'], round(r['ms'],4), round(r['frac'],3))
print('training',d['training_path'])
print('c5',d['config5_quantise_plus_gather']); print('grouped',d['grouped_tts_quantiser'])
PY
