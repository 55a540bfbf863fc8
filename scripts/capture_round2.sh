#!/bin/bash
# Round-2 evidence from ONE B200 (run through gpurun): the bench line, the ncu launch list of the same command, and one
# `ncu --set full` capture per hot kernel on view 0 of the headline workload.  Outputs land in gpurun_out/ (r02_*);
# scripts/ncu_summary.py turns the reports into the text summaries committed under profiles/.
set -u
tag=${1:-final}
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1_${tag}.json 2> gpurun_out/r02_bench_n1_${tag}.err
tail -c 400 gpurun_out/r02_bench_n1_${tag}.err
LVDGS_BENCH_1M=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file gpurun_out/r02_launches_${tag}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_bench_${tag}.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"blend_forward|blend_backward|preprocess_backward|preprocess_forward|emit_keys|tile_sort_short|binning" -s 16 -c 8 \
    -o gpurun_out/r02_prof_${tag} -f python scripts/profile_step.py 500000 kitti 1 > gpurun_out/r02_ncu_${tag}.log 2>&1
tail -2 gpurun_out/r02_ncu_${tag}.log
