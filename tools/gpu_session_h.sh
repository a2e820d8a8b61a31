#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== op tests: pair"; TDNET_TC_PAIR_VERBOSE=1 timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "pair or variant" -s 2>&1 | tail -8
echo "== probes"; for v in 0 1; do TDNET_TC_PAIR_SPLIT=$v timeout 100 python tools/tc_probe.py --one layer4_perf 2>&1 | tail -1 | cut -c1-330; done
echo "== model parity"; timeout 400 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "golden or full_size" 2>&1 | tail -3
echo "== bench"; timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_split.json | cut -c1-330
