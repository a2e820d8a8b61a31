#!/bin/bash
# TD2-FANet first GPU validation: new kernels vs torch, model vs reference fixtures / oracle, timing, launch list
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== fanet tests"; timeout 500 python -m pytest tests/test_fanet_gpu.py -q -m gpu --tb=short 2>&1 | tail -70
echo "== timing"; timeout 200 python tools/fanet_time.py 2>&1 | tail -2 | cut -c1-400 | tee gpurun_out/fanet_time.json
echo "== launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/fanet_launches.csv python tools/fanet_time.py --steps 2 --warmup 1 > gpurun_out/fanet_ncu.log 2>&1; tail -1 gpurun_out/fanet_ncu.log | cut -c1-200
