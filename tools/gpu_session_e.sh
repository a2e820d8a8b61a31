#!/bin/bash
# Full GPU suite with the new defaults (tcgen05 stem, halo rule, pspnet), bench, ncu launch list, ncu of the stem.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
echo "== pytest -m gpu (as the driver runs it)"; timeout 900 python -m pytest tests/ -x -q -m gpu --durations=6 2>&1 | tail -16 | tee gpurun_out/t_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 400 python bench.py --steps 40 --warmup 8 2>&1 | tail -1 | tee gpurun_out/bench_tc.json | cut -c1-600
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(tc_|stem_|conv_simt|bilinear|copy_nhwc|maxpool|psp_|ln_|upsample|softmax|image_to)' -s 700 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu full: tc stem"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_stem_kernel -s 1 -c 1 -f -o gpurun_out/prof_tcstem python tools/tc_probe.py --one stem_tc_perf > gpurun_out/ncu_tcstem.log 2>&1; tail -2 gpurun_out/ncu_tcstem.log | cut -c1-200
ls -la gpurun_out | head -30
