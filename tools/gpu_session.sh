#!/bin/bash
mkdir -p gpurun_out
python tools/ab_k1.py ab/libvqb200_i24_s184.so ab/libvqb200_prev.so 2>&1 | tail -1 | tee gpurun_out/experiment_regs.log
