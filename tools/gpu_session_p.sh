#!/bin/bash
# Programmatic dependent launch on the tcgen05 kernels: full tests with it on, bench / FANet timing on vs off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== smoke (PDL on)"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest -m gpu (PDL on)"; timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/t_gpu.log
for v in 0 1; do
echo "== bench TDNET_PDL=$v"; TDNET_PDL=$v timeout 200 python bench.py --steps 40 --warmup 8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_pdl$v.json | cut -c1-230
echo "== fanet TDNET_PDL=$v"; TDNET_PDL=$v timeout 120 python tools/fanet_time.py 2>&1 | tail -1 | cut -c1-200
done
