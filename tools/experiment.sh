#!/bin/bash
# timing experiments on the tcgen05 kernel (results are wrong by construction; only the time matters)
# bit0: no scan arithmetic, bit1: no TMEM reads, bit2: never fall back, bit3: no conversion work
for e in ${EXPERIMENTS:-4 5 6 7 15 12 0}; do
  VQ_EXPERIMENT=$e python speech-masters-thesis_b200/build.py --force > /dev/null 2>&1
  echo "experiment $e: $(python bench.py --steps 20 --warmup 3 --profile-only 2>&1 | tail -1)"
done
