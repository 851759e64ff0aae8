#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_audit.py -m gpu -x -q -k "every_row or sweep_corner" 2>&1 | tail -3
python tools/ab_k1.py ab/libvqb200_e0.so ab/libvqb200_e1024.so ab/libvqb200_e2048.so ab/libvqb200_e3072.so 2>&1 | tail -1 | tee gpurun_out/experiment_scan.log
python tools/ab_k1.py ab/libvqb200_e0.so ab/libvqb200_e3072.so gaussian 2>&1 | tail -1 | tee -a gpurun_out/experiment_scan.log
python tools/tc_timeline.py > gpurun_out/tl.log 2>&1; tail -33 gpurun_out/tl.log | head -12
