#!/bin/bash
# large-K offsets staged per scan warp: parity first, then the sweep
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_audit.py -m gpu -x -q -k "sweep_corner or tcgen05_route or bf16" 2>&1 | tail -4
timeout 1200 python tools/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/sweep.log | cut -c1-300
timeout 600 python tools/sweep.py --gaussian --quick > gpurun_out/sweepg.log 2>&1; echo "sweepg rc=$?"
