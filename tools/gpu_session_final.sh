#!/bin/bash
# Round-end batch on one B200: the driver's GPU test command, smoke, bench (with the CPU baseline leg), launch lists
# of the td4-psp18 bench and of one TD2-FANet call, sanitizer passes over the new FANet kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
echo "== pytest -m gpu (as the driver runs it)"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/t_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 400 python bench.py --steps 40 --warmup 8 2>&1 | tail -1 | tee gpurun_out/bench_final.json | cut -c1-600
echo "== fanet timing"; timeout 200 python tools/fanet_time.py 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/fanet_time.json
timeout 200 python tools/fanet_time.py --backbone resnet34 2>&1 | tail -1 | cut -c1-400 | tee -a gpurun_out/fanet_time.json
echo "== ncu launch list: bench"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(tc_|stem_|conv_simt|bilinear|copy_nhwc|maxpool|psp_|ln_|upsample|softmax|image_to|fa_|add_up)' -s 450 -c 360 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -1 gpurun_out/ncu_bench.log | cut -c1-200
echo "== ncu launch list: fanet"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^(tc_|stem_|conv_simt|bilinear|copy_nhwc|maxpool|psp_|ln_|upsample|softmax|image_to|fa_|add_up)' -c 600 --csv --log-file gpurun_out/fanet_launches.csv python tools/fanet_time.py --steps 2 --warmup 2 > gpurun_out/fanet_ncu.log 2>&1; tail -1 gpurun_out/fanet_ncu.log | cut -c1-200
export TDNET_B200_CUDA_GRAPH=0
echo "== memcheck: fanet kernels + smallest fanet model case"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 1 python -m pytest tests/test_fanet_gpu.py -q -m gpu -x -k "(fa_linear and not 128-256) or add_upsampled or (golden and r50)" 2>&1 | tail -8 | tee gpurun_out/sanitize_memcheck_fanet.log
echo "== racecheck: fanet kernels"; timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_fanet_gpu.py -q -m gpu -x -k "(fa_linear and not 128-256 and not 2048) or add_upsampled" 2>&1 | tail -6 | tee gpurun_out/sanitize_racecheck_fanet.log
