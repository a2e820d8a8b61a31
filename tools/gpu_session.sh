#!/bin/bash
# One batched GPU session (gpurun calls are expensive to acquire): probe, tests, bench, profiles.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
if [ "$1" != "noprobe" ]; then
echo "== tc_probe"; timeout 1200 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; tail -18 gpurun_out/probe.log | cut -c1-400
fi
echo "== pytest -m gpu (as the driver runs it)"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/t_gpu.log
echo "== model tc, unfused attention chain, no graphs / side stream"; TDNET_B200_FUSED_ATTN=0 TDNET_B200_CUDA_GRAPH=0 timeout 400 python -m pytest tests/test_model_gpu.py -m gpu -q -k "golden and tc" 2>&1 | tail -8 | tee gpurun_out/t_model_unfused.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench tc"; timeout 300 python bench.py --steps 40 --warmup 8 2>&1 | tail -1 | tee gpurun_out/bench_tc.json | cut -c1-1500
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(tc_|stem_|conv_simt|bilinear|copy_nhwc|maxpool|psp_|ln_|upsample|softmax|image_to)' -s 450 -c 360 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full: conv"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 3 -c 1 -f -o gpurun_out/prof_conv python tools/tc_probe.py --one layer4_perf > gpurun_out/ncu_conv.log 2>&1; tail -2 gpurun_out/ncu_conv.log | cut -c1-200
echo "== ncu full: attention"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_attn_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn python tools/tc_probe.py --one attention_big > gpurun_out/ncu_attn.log 2>&1; tail -2 gpurun_out/ncu_attn.log | cut -c1-200
ls -la gpurun_out | head -30
