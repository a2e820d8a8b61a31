#!/bin/bash
# One batched GPU session (gpurun calls are expensive to acquire): probe, tests, bench, launch list.
mkdir -p gpurun_out
echo "== tc_probe"; timeout 1000 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; tail -15 gpurun_out/probe.log
echo "== ops (simt + split16)"; timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "not tc_conv" 2>&1 | tail -8 | tee gpurun_out/t_ops.log
echo "== ops (tcgen05)"; timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "tc_conv" 2>&1 | tail -12 | tee gpurun_out/t_ops_tc.log
echo "== model simt"; timeout 400 python -m pytest tests/test_model_gpu.py -m gpu -q -k "simt" 2>&1 | tail -8 | tee gpurun_out/t_model_simt.log
echo "== model tc"; timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -k "not simt" 2>&1 | tail -25 | tee gpurun_out/t_model_tc.log
echo "== bench tc"; timeout 300 python bench.py --steps 30 --warmup 8 2>&1 | tail -2 | tee gpurun_out/bench_tc.json
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
