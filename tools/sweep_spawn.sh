for t in woodcock surface; do for sb in 1 2 4; do for ch in 32 64 128; do
  python bench.py --steps 8 --warmup 3 --no-cpu --no-variants --tracking $t --spawn-batch $sb --chunk $ch 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$t spawn_batch=$sb chunk=$ch', round(d['value']/1e6,1), 'Mh/s')"
done; done; done
