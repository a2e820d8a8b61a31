#!/bin/bash
# One B200 session: the driver's GPU test command, smoke, bench (default line), the reference arm.  Usage: bash tools/gpu_session.sh [tag]
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
rm -f gpurun_out/parity_metrics.jsonl
echo "== pytest -m gpu (as the driver runs it)"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/${TAG}_t_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 40 --warmup 8 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-300
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json | cut -c1-400
