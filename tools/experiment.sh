#!/bin/bash
# timing experiments on the tcgen05 kernel (results are wrong by construction; only the time matters)
# bit0: no scan arithmetic, bit1: no TMEM reads, bit2: never fall back, bit3: no conversion work, bit5 (32): no x loads,
# bit8 (256): the back stage only shakes hands, bit9 (512, with bit0): the scan warps sleep instead of scanning
# (results of round 2: profiles/r02_k1_experiments.json; A/B of two builds on one box: tools/ab_k1.py)
for e in ${EXPERIMENTS:-4 5 6 7 15 12 0}; do
  VQ_EXPERIMENT=$e python speech-masters-thesis_b200/build.py --force > /dev/null 2>&1
  echo "experiment $e: $(python bench.py --steps 20 --warmup 3 --profile-only 2>&1 | tail -1)"
done
