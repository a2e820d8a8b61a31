#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== op tests: attention"; timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -4
echo "== probe"; for e in attention_big attention_ragged; do timeout 100 python tools/tc_probe.py --one $e 2>&1 | tail -1 | cut -c1-330; done
echo "== ncu full: attention"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_attn_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn python tools/tc_probe.py --one attention_big > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log | cut -c1-200
