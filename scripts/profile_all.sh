#!/bin/bash
# ncu evidence for profiles/ (one GPU; each ncu run follows a plain run of the same command that exited 0):
#   /usr/local/graft/bin/gpurun --timeout 1200 -- 'bash scripts/profile_all.sh'
set -u
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --pairs 8 --no-cpu --no-e2e --inversion none"
$B > gpurun_out/r2_prof_bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_bench_launches.csv $B > gpurun_out/r2_prof_bench_ncu.log 2>&1
S="python scripts/pbs_lat.py w4 8 3 pairs"
$S > gpurun_out/r2_prof_split_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pbs_split -s 1 -c 1 -o gpurun_out/r2_final_split_w4_pairs_8 $S > gpurun_out/r2_prof_split_ncu.log 2>&1
tail -2 gpurun_out/r2_prof_split_ncu.log
