#!/bin/bash
# Pair kernel on by default: model parity, bench, ncu of the pair kernel on the layer-4 conv.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== model parity"; timeout 400 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "golden or full_size" 2>&1 | tail -4
echo "== bench"; timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_pair.json | cut -c1-330
echo "== ncu full: pair conv"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_pair_kernel -s 3 -c 1 -f -o gpurun_out/prof_conv_pair python tools/tc_probe.py --one layer4_perf > gpurun_out/ncu_conv_pair.log 2>&1; tail -2 gpurun_out/ncu_conv_pair.log | cut -c1-200
