#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_audit.py -m gpu -q -k "nccl or second_device or revival" 2>&1 | tail -3; fi
for p in 1 0; do VQ_P2P=$p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$p tools/train_fwd_bench.py 2>/dev/null | tail -1; done | tee gpurun_out/train_fwd_${N}gpu.log
