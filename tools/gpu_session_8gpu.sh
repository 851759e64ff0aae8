#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"
tail -c 300 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.log').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'n',d['n_gpus'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'])
print('training',d['training_path'])
print('c5',d['config5_quantise_plus_gather'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/corpus_encode.py --out gpurun_out/corpus_encode_${N}gpu.json 2>&1 | tail -1 | cut -c1-600
