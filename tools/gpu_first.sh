#!/bin/bash
# first GPU contact: parity tests + integer-pipe microbenchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python tools/microbench.py > gpurun_out/microbench.json 2> gpurun_out/microbench.err
cat gpurun_out/microbench.json
nproc > gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/host.txt
