for rep in 1 2; do
for t in 0 1; do
  echo "TAIL=$t"; TDNET_TC_PAIR_TAIL=$t timeout 300 python bench.py --steps 60 --warmup 8 --no-cpu-baseline --no-fast-mode --sustain-seconds 0.5 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],2), d['ms_per_step'], round(d['e2e']['value'],1), d['roofline']['ms_per_launch'], d['roofline_attention']['ms_per_launch'])"
done; done
