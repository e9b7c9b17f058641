for L in 2 4 5; do
timeout 120 python bench.py --lanes $L --steps 100 --no-cpu-baseline --no-layer-profile 2>/dev/null | python -c "
import sys, json
j=json.loads(sys.stdin.read()); print('lanes', $L, 'value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), j['config']['timed_regions']['device_ms_per_step_min_med_max'])"
done
