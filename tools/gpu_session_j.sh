#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== op tests: attention"; timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -4
echo "== probe"; for e in attention_big attention_ragged attention_small; do timeout 100 python tools/tc_probe.py --one $e 2>&1 | tail -1 | cut -c1-330; done
echo "== model parity"; timeout 400 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "golden or two_cycles" 2>&1 | tail -3
