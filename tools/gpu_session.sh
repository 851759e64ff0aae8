#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_audit.py tests/test_gpu_parity.py -m gpu -x -q -k "every_row or sweep_corner or golden or random" 2>&1 | tail -3
for o in 1 0; do echo "own $o: $(VQ_K1_OWN=$o python tools/ab_k1.py speech-masters-thesis_b200/lib/libvqb200.so 2>&1 | tail -1)"; done | tee gpurun_out/experiment_own.log
for o in 1 0; do echo "own $o gaussian: $(VQ_K1_OWN=$o python tools/ab_k1.py speech-masters-thesis_b200/lib/libvqb200.so gaussian 2>&1 | tail -1)"; done | tee -a gpurun_out/experiment_own.log
python tools/tc_timeline.py > gpurun_out/tl.log 2>&1; tail -33 gpurun_out/tl.log | head -17
