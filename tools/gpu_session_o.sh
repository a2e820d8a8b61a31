#!/bin/bash
# Attention tail split (ragged last round as 128-channel items): op tests, probe with / without, model parity, bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== op tests: attention"; timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -4
echo "== probe"; for v in 0 1; do TDNET_ATTN_TAIL=$v timeout 100 python tools/tc_probe.py --one attention_big 2>&1 | tail -1 | cut -c1-330; done
echo "== model parity"; timeout 600 python -m pytest tests/test_model_gpu.py tests/test_fanet_gpu.py -x -q -m gpu -k "golden or two_cycles or full_size or config5" 2>&1 | tail -3
echo "== bench"; timeout 300 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_o.json | cut -c1-260
