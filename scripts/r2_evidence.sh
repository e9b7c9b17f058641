#!/bin/bash
# Round 2 evidence run (1 GPU): sanitizers on the shipped kernels, pair-kernel trace, full ncu of the 18 conv
# launches, ncu of the geometry kernels, launch list of one step.
#   gpurun --timeout 1500 -- 'bash scripts/r2_evidence.sh <tag>'
TAG=${1:-r2}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/${TAG}_compute_sanitizer_$tool.log 2>&1
  tail -2 gpurun_out/${TAG}_compute_sanitizer_$tool.log
done
MSI_TC_TRACE=conv2_1 timeout 120 python scripts/trace_conv.py > gpurun_out/${TAG}_trace_conv2_1_pair.log 2>&1
tail -3 gpurun_out/${TAG}_trace_conv2_1_pair.log
timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 18 -c 18 -f -o gpurun_out/${TAG}_conv \
    python scripts/one_forward.py > gpurun_out/${TAG}_ncu_conv.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_conv.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"psv_gather_pair|render_composite_v2|prep_images" -s 3 -c 3 -f \
    -o gpurun_out/${TAG}_geom python scripts/one_frame.py > gpurun_out/${TAG}_ncu_geom.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_geom.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --lanes 1 --no-graph --no-layer-profile --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
wc -l gpurun_out/${TAG}_launches.csv; ls -la gpurun_out/${TAG}_conv.ncu-rep gpurun_out/${TAG}_geom.ncu-rep
