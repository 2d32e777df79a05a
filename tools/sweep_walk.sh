for wc in -1 3 4 5 6 8; do
  timeout 60 python bench.py --steps 8 --warmup 3 --no-cpu --no-variants --walk-cap $wc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('config3 walk_cap=$wc', round(d['value']/1e6,1), 'Mh/s')"
done
for wc in -1 20 40 80; do
  timeout 60 python bench.py --steps 3 --warmup 3 --no-cpu --no-variants --workload config4 --histories 20000000 --walk-cap $wc 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('config4(2e7) walk_cap=$wc', round(d['value']/1e6,1), 'Mh/s')"
done
