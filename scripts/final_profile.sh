#!/bin/bash
# round-end evidence: bench line, ncu --set full of one forward's conv launches, ncu launch list of one step
mkdir -p gpurun_out
timeout 100 python bench.py > gpurun_out/r1_v10_bench_n1.json 2> gpurun_out/r1_v10_bench_n1.err
cut -c1-300 gpurun_out/r1_v10_bench_n1.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 18 -c 18 -f -o gpurun_out/r1_v10_conv \
    python scripts/one_forward.py > gpurun_out/r1_v10_ncu_full.log 2>&1
tail -2 gpurun_out/r1_v10_ncu_full.log
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_v10_launches.csv \
    python bench.py --steps 2 --warmup 1 --lanes 1 --no-graph --no-layer-profile --no-cpu-baseline > gpurun_out/r1_v10_ncu_list.log 2>&1
wc -l gpurun_out/r1_v10_launches.csv; ls -la gpurun_out/r1_v10_conv.ncu-rep
