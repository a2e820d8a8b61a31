#!/bin/bash
# CTA-pair conv kernel: correctness vs the single-CTA kernel, then the perf sweep.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== op tests: pair variant"; timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "pair or variant" 2>&1 | tail -15
echo "== probes: pair sweep"; TDNET_PROBE_PAIR_SWEEP=1 timeout 500 python tools/tc_probe.py > gpurun_out/probe_pair.log 2>&1; cut -c1-330 gpurun_out/probe_pair.log
