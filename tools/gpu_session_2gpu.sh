#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_audit.py -m gpu -q -k "nccl or second_device or two_streams or revival" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_2gpu.log
tail -4 gpurun_out/pytest_2gpu.log

timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"
tail -c 400 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_2gpu.log').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'n',d['n_gpus'],'e2e',d['e2e']['value'])
print('training',d['training_path'])
print('c5',d['config5_quantise_plus_gather'])
PY
