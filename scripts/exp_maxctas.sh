# conv launches capped at N SMs (MSI_CONV_MAX_CTAS): do two frames' conv kernels side by side beat one after the other?
mkdir -p gpurun_out
IFS=","; for cfg in ${CFGS:-"0 4" "74 4" "74 6" "100 4" "112 4" "74 8"}; do
IFS=" "; set -- $cfg
MSI_CONV_MAX_CTAS=$1 timeout 150 python bench.py --steps 50 --lanes $2 --no-cpu-baseline > gpurun_out/exp_maxctas_$1_l$2.json 2> gpurun_out/exp_maxctas_$1_l$2.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/exp_maxctas_$1_l$2.json")); r=j['roofline']
    print('max_ctas', $1, 'lanes', $2, 'value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), '1lane', round(j['config']['one_frame_at_a_time']['value'],1), 'conv', round(r['kernel_ms_per_step'],4), j['clocks']['sm_mhz'], (j.get('energy') or {}).get('frames_per_s_per_gpu'))
except Exception as e: print('failed', e)
PY
done
