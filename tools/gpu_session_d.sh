#!/bin/bash
# Validation of the two newest kernels (tcgen05 stem, halo-region 3x3 conv) + A/B benches + the full GPU suite.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== probes: tc stem + halo sweep"; TDNET_PROBE_HALO_SWEEP=1 timeout 500 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; cut -c1-330 gpurun_out/probe.log
echo "== op tests: stems"; timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "stem" 2>&1 | tail -6
echo "== model parity with both variants on"; TDNET_B200_TC_STEM=1 TDNET_TC_HALO=1 timeout 400 python -m pytest tests/test_model_gpu.py -x -q -m gpu -k "golden" 2>&1 | tail -6 | tee gpurun_out/t_model_variants.log
echo "== bench default"; timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_default.json | cut -c1-260
echo "== bench tc stem"; TDNET_B200_TC_STEM=1 timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_tcstem.json | cut -c1-260
echo "== bench halo"; TDNET_TC_HALO=1 timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_halo.json | cut -c1-260
echo "== bench both"; TDNET_B200_TC_STEM=1 TDNET_TC_HALO=1 timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_both.json | cut -c1-260
echo "== pytest -m gpu (as the driver runs it)"; timeout 900 python -m pytest tests/ -x -q -m gpu --durations=8 2>&1 | tail -22 | tee gpurun_out/t_gpu.log
