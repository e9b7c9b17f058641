#!/bin/bash
# Round 2, GPU session 2: parity of the cached-coordinate sweep (K1) and the fast-chain render (K5), bench A/B,
# and one full ncu capture of the two new kernels.
#   gpurun --timeout 900 -- 'bash scripts/r2_session1.sh'
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2_s2_pytest.log 2>&1
tail -5 gpurun_out/r2_s2_pytest.log
grep "render fast chain" gpurun_out/r2_s2_pytest.log
timeout 200 python bench.py > gpurun_out/r2_s2_bench.json 2> gpurun_out/r2_s2_bench.err
cut -c1-400 gpurun_out/r2_s2_bench.json
MSI_RENDER_V1=1 timeout 120 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2_s2_bench_renderv1.json 2> gpurun_out/r2_s2_bench_renderv1.err
python - <<'PY'
import json
for f in ("r2_s2_bench", "r2_s2_bench_renderv1"):
    try:
        j = json.load(open(f"gpurun_out/{f}.json"))
        r = j["roofline"]
        print(f, "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "1lane", round(j["config"]["one_frame_at_a_time"]["value"], 1),
              {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in r["hbm_kernels"].items()}, "conv_ms", round(r["kernel_ms_per_step"], 4))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"psv_gather_pair|render_composite_v2|prep_images" -s 3 -c 3 -f \
    -o gpurun_out/r2_s2_geom python scripts/one_frame.py > gpurun_out/r2_s2_ncu_geom.log 2>&1
tail -2 gpurun_out/r2_s2_ncu_geom.log
ls -la gpurun_out/r2_s2_geom.ncu-rep
