#!/bin/bash
# FIFO push on the side stream, classifier kernel, vectorised LN partial sums: tests, bench with each toggle
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/t_gpu.log
echo "== bench default"; timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n_default.json | cut -c1-260
echo "== bench FIFO_OVERLAP=0"; TDNET_B200_FIFO_OVERLAP=0 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n_nofifo.json | cut -c1-260
echo "== bench SMALL_LINEAR=0"; TDNET_B200_SMALL_LINEAR=0 timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n_nolinear.json | cut -c1-260
echo "== sweep (published-table size)"; timeout 600 python tools/config_sweep.py --only 6 7 8 2>&1 | grep -v "^\[build\]" | tail -4 | cut -c1-300 | tee gpurun_out/config_sweep_769.jsonl
