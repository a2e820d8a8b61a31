#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; cut -c1-420 gpurun_out/probe.log
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "stem" 2>&1 | tail -15
