#!/bin/bash
# In-frame ncu capture of the dominant conv + bench sanity after the N-tile rule change.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bench"; timeout 300 python bench.py --steps 40 --warmup 8 2>&1 | tail -1 | tee gpurun_out/bench_tc.json | cut -c1-400
echo "== ncu full: layer4 conv inside running frames (launch 19 of a frame = layer4.1.conv2; skip 8 warm frames)"
# per steady frame there are 28 tc_conv_kernel launches; frames 0-2 are warm-up plans with fewer. Skip ~ 9 frames.
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_conv_kernel -s 265 -c 4 -f -o gpurun_out/prof_conv_inframe python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_inframe.log 2>&1; tail -2 gpurun_out/ncu_inframe.log | cut -c1-200
echo "== ncu full: stem"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:stem_conv_pool -s 6 -c 1 -f -o gpurun_out/prof_stem python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_stem.log 2>&1; tail -1 gpurun_out/ncu_stem.log | cut -c1-200
ls -la gpurun_out | head
