#!/bin/bash
mkdir -p gpurun_out
python tools/ab_k1.py speech-masters-thesis_b200/lib/libvqb200.so ab/libvqb200_prev.so 2>&1 | tail -1 | tee gpurun_out/experiment_tracetpl.log
python tools/ab_k1.py speech-masters-thesis_b200/lib/libvqb200.so ab/libvqb200_prev.so gaussian 2>&1 | tail -1 | tee -a gpurun_out/experiment_tracetpl.log
timeout 900 python -m pytest tests/test_gpu_audit.py tests/test_gpu_parity.py -m gpu -x -q -k "every_row or sweep_corner or golden or random or tcgen05_route or bf16" 2>&1 | tail -3
python tools/tc_timeline.py > gpurun_out/tl.log 2>&1; tail -34 gpurun_out/tl.log | head -10
