#!/bin/bash
# Last batch of the round: the driver's GPU test command on the final tree, sanitizer over the kernels added last
# (resize, sampled arg-max, classifier, PDL launches), one ncu --set full capture of the attention op with the tail split
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu (as the driver runs it)"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/t_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== ncu full: attention (two launches)"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_attn_kernel -s 4 -c 2 -f -o gpurun_out/prof_attn_split python tools/tc_probe.py --one attention_big > gpurun_out/ncu_attn_split.log 2>&1; tail -2 gpurun_out/ncu_attn_split.log | cut -c1-200
export TDNET_B200_CUDA_GRAPH=0
echo "== memcheck: ingest / preview / classifier / small model (PDL launches)"; timeout 600 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 1 python -m pytest tests/test_ingest_gpu.py tests/test_ops_gpu.py -q -m gpu -x -k "(resize and not 1024) or preview or pointwise_linear and not 128-256" 2>&1 | tail -6 | tee gpurun_out/sanitize_memcheck_ingest.log
