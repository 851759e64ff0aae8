#!/bin/bash
# final multi-GPU validation: the 2-GPU-only tests (TESTS=1), then bench.py the way the driver launches it (both arms)
mkdir -p gpurun_out
N=${NGPU:-2}
if [ "${TESTS:-0}" = "1" ]; then timeout 900 python -m pytest tests -m gpu -q -k "nccl or p2p or second_device or revival or two_streams" 2>&1 | tail -3; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"
tail -c 300 gpurun_out/bench_${N}gpu.err
if [ "${REF:-0}" = "1" ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_${N}gpu.log 2> gpurun_out/bench_ref_${N}gpu.err; echo "ref$N rc=$?"; fi
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${N}gpu.log') if l.startswith('{')][-1])
print('value',d['value'],'ms',d['ms_per_step'],'n',d['n_gpus'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'])
print('training',{k:v['ms'] for k,v in d['training_path'].items() if isinstance(v,dict)})
print('c5',d['config5_quantise_plus_gather']['ms'])
PY
