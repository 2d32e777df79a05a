for t in woodcock surface; do for cfg in "512 3" "384 4" "448 3" "512 2" "256 6"; do
  set -- $cfg
  timeout 60 python bench.py --steps 8 --warmup 3 --no-cpu --no-variants --tracking $t --threads $1 --blocks-per-sm $2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$t $cfg', round(d['value']/1e6,1), 'Mh/s', d['config']['launch']['grid'], d['config']['launch']['block'])"
done; done
