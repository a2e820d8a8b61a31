#!/bin/bash
# One B200 session: the driver's GPU test command, smoke, the attention probe (kernel families), bench (default line) with
# either attention variant, the unfiltered launch list of steady frames.  Usage: bash tools/gpu_session.sh [tag] [steps...]
# steps: tests smoke probe bench bench_tq launches reference ncu:<kernel regex> (default: the first six)
cd "$(dirname "$0")/.."
TAG=${1:-r02}
shift
STEPS=${*:-tests smoke probe bench bench_tq launches}
mkdir -p gpurun_out
has() { [[ " $STEPS " == *" $1 "* ]]; }
if has tests; then
  rm -f gpurun_out/parity_metrics.jsonl
  echo "== pytest -m gpu (as the driver runs it)"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/${TAG}_t_gpu.log
  cp gpurun_out/parity_metrics.jsonl gpurun_out/${TAG}_parity_metrics.jsonl 2>/dev/null
fi
if has smoke; then echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; fi
if has probe; then
  echo "== attention probe"; ATTN_PROBE_FAMILIES=ss,ts,tq timeout 300 python tools/attn_probe.py gpurun_out/${TAG}_attn_probe.jsonl 2>&1 | tail -9 | cut -c1-420
  echo "== attention probe, sustained"; ATTN_PROBE_FAMILIES=ts,tq timeout 120 python tools/attn_probe.py --sustain gpurun_out/${TAG}_attn_probe_sustained.jsonl 2>&1 | tail -2 | cut -c1-600
fi
if has bench; then echo "== bench"; timeout 600 python bench.py --steps 40 --warmup 8 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench.json | cut -c1-1200; fi
if has bench_tq; then echo "== bench, TDNET_ATTN_TS=3"; TDNET_ATTN_TS=3 timeout 600 python bench.py --steps 40 --warmup 8 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_tq.json | cut -c1-1200; fi
if has launches; then
  echo "== launch list"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_frame_all.csv python tools/frame_launches.py > gpurun_out/${TAG}_frame_all.log 2>&1
  python tools/launch_summary.py gpurun_out/${TAG}_frame_all.csv | head -24
fi
# ncu:<regex> -- one full-set capture (source-level stall samples included) of the first matching launch inside steady frames
for st in $STEPS; do
  if [[ $st == ncu:* ]]; then
    K=${st#ncu:}
    echo "== ncu --set full of $K"
    timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name regex:$K --launch-count 1 \
      -f -o gpurun_out/${TAG}_prof_${K} python tools/frame_launches.py > gpurun_out/${TAG}_prof_${K}.log 2>&1
    tail -2 gpurun_out/${TAG}_prof_${K}.log
  fi
done
if has reference; then echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/${TAG}_bench_reference.json | cut -c1-400; fi
