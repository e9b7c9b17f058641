# conv CTAs leave R bytes of shared memory free on their SM (co-residence of another lane's LayerNorm / sweep blocks)
#   gpurun --timeout 600 -- 'bash scripts/exp_reserve.sh "0 4096 8192"'
mkdir -p gpurun_out
for R in ${1:-16384 32768}; do
MSI_CONV_SMEM_RESERVE=$R timeout 150 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/exp_reserve_$R.json 2> gpurun_out/exp_reserve_$R.err
tail -2 gpurun_out/exp_reserve_$R.err
python - <<PY
import json
try:
    j=json.load(open("gpurun_out/exp_reserve_$R.json")); r=j['roofline']
    print('reserve', $R, 'value', round(j['value'],1), 'e2e', round(j['e2e']['value'],1), '1lane', round(j['config']['one_frame_at_a_time']['value'],1), 'conv', round(r['kernel_ms_per_step'],4), j['clocks'])
except Exception as e: print('failed', e)
PY
done
