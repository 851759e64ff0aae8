#!/bin/bash
for e in 256 0; do
  VQ_EXPERIMENT=$e python speech-masters-thesis_b200/build.py --force > /dev/null 2>&1
  echo "experiment $e: $(python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['training_path_kernels']['K3_ema_accumulate'])")"
done
