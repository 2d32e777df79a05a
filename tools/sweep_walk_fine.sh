for wc in 8 12 16 20 24 28 32; do
  timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu --no-variants --workload config4 --histories 20000000 --walk-cap $wc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('config4(2e7) walk_cap=$wc', round(d['value']/1e6,1), 'Mh/s')"
done
