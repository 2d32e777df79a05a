mkdir -p gpurun_out
for cfg in "1024 1 64" "512 2 64" "256 4 64" "128 8 64" "512 2 32" "512 2 128" "512 2 256" "384 2 64" "640 1 64" "768 1 64"; do
  set -- $cfg
  python bench.py --steps 6 --warmup 3 --no-cpu --threads $1 --blocks-per-sm $2 --chunk $3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$cfg', round(d['value']/1e6,1), 'Mh/s', d['config']['launch'])"
done
python bench.py --steps 3 --warmup 3 --no-cpu --workload config4 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('config4', round(d['value']/1e6,1), 'Mh/s', d['ms_per_step'], d['config']['launch'], d['roofline']['collisions_per_history'])"
