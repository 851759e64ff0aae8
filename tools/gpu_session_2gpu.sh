#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
for sb in 0 1; do TFB_SAME_BLOCK=$sb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$sb tools/train_fwd_bench.py 2>/dev/null | tail -1; done | tee gpurun_out/train_fwd_${N}gpu_order.log
