for t in woodcock surface; do for sb in 1 2 3 4 6 8 12; do
  python bench.py --steps 8 --warmup 3 --no-cpu --no-variants --tracking $t --spawn-batch $sb 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$t spawn_batch=$sb', round(d['value']/1e6,1), 'Mh/s')"
done; done
