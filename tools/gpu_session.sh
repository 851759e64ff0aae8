#!/bin/bash
mkdir -p gpurun_out
VQ_K1_HARD=1 ncu --set full --clock-control none --import-source on -k regex:assign_list -c 1 -o gpurun_out/prof_list python tools/list_debug.py > gpurun_out/ncu_list.log 2>&1
tail -3 gpurun_out/ncu_list.log | cut -c1-300
