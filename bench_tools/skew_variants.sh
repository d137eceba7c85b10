for v in 34 52 53 54; do timeout 100 python bench_tools/skew.py --variant $v --reps 2 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['variant'], d['kind'], d['sorted'], d['stage_ms'][1], d['gkeys_s'])
"; done
