#!/bin/bash
# Round-2 final evidence (1 GPU): full GPU suite, the bench as the driver runs it (both arms), a long-region bench,
# ncu --set full of the 18 conv launches (default precision) and of the geometry kernels, launch list of one step.
#   gpurun --timeout 1800 -- 'bash scripts/r2_final.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/r2_final_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r2_final_pytest.log | tail -5
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_bench_reference.json 2> gpurun_out/r2_final_bench_reference.err
tail -1 gpurun_out/r2_final_bench_reference.json | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_final_bench_n1.json 2> gpurun_out/r2_final_bench_n1.err
timeout 300 python bench.py --steps 200 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_bench_n1_steps200.json 2> gpurun_out/r2_final_bench_n1_steps200.err
for f in r2_final_bench_n1 r2_final_bench_n1_steps200; do
python - <<PY
import json
try:
    j = json.load(open("gpurun_out/$f.json")); r = j["roofline"]
    print("$f value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "1lane", round(j["config"]["one_frame_at_a_time"]["value"], 1),
          {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in r["hbm_kernels"].items()}, "conv_ms", round(r["kernel_ms_per_step"], 4), "frac", round(r["frac"], 3))
    print("   clocks", j["clocks"], "energy", j.get("energy"))
    print("   parity", {k: v for k, v in (j.get("parity") or {}).items() if not isinstance(v, (dict, str))})
except Exception as e:
    print("$f failed", e)
PY
done
if [ -z "$SKIP_NCU_FULL" ]; then
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 18 -c 18 -f -o gpurun_out/r2_final_conv \
    python scripts/one_forward.py > gpurun_out/r2_final_ncu_conv.log 2>&1
tail -2 gpurun_out/r2_final_ncu_conv.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"psv_gather_pair|render_composite_v2|prep_images" -s 3 -c 3 -f \
    -o gpurun_out/r2_final_geom python scripts/one_frame.py > gpurun_out/r2_final_ncu_geom.log 2>&1
tail -2 gpurun_out/r2_final_ncu_geom.log
fi
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv \
    python bench.py --steps 2 --warmup 1 --lanes 1 --no-graph --no-layer-profile --no-cpu-baseline > gpurun_out/r2_final_ncu_list.log 2>&1
wc -l gpurun_out/r2_final_launches.csv; ls -la gpurun_out/r2_final_conv.ncu-rep gpurun_out/r2_final_geom.ncu-rep
