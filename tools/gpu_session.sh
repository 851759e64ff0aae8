#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -c 600 gpurun_out/bench_ref.log
timeout 300 python tools/torch_cuda_baseline.py > gpurun_out/torch_base.log 2>&1; tail -c 700 gpurun_out/torch_base.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
