#!/bin/bash
# K3a occupancy experiment: 2 CTAs per SM with 32-deep slices (build-time knobs), A/B against the product
mkdir -p gpurun_out
for l in "" ab/libvqb200_k3a.so ab/libvqb200_k3b.so; do
  echo "lib=${l:-product}: $(VQB200_LIB=${l:+$PWD/$l} K23_ONLY=k3a python tools/k23_bench.py 2>&1 | tail -1 | cut -c1-300)"
done | tee gpurun_out/experiment_k3a.log
VQB200_LIB=$PWD/ab/libvqb200_k3a.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ema or golden or random" 2>&1 | tail -2
