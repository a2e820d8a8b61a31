#!/bin/bash
# BASELINE configs[0], [3], [4] + comparison models: full-size stream-independence property test and throughput sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== config-5 property test"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --tb=short -k "config5" 2>&1 | tail -15
echo "== sweep"; timeout 900 python tools/config_sweep.py 2>&1 | grep -v "^\[build\]" | tail -8 | cut -c1-300 | tee gpurun_out/config_sweep.jsonl
