#!/bin/bash
# Everything the DESIGN.md tables quote, in one GPU call (about 8 minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash scripts/measure_all.sh'
# Results land in gpurun_out/measure_*.{log,jsonl,json}; copy what is to be tracked into profiles/.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/measure_gpu_tests.log 2>&1; tail -2 gpurun_out/measure_gpu_tests.log
timeout 300 python scripts/inversion_time.py inv2_low inv2_low_prefix inv3_low inv3_low_prefix inv3_medium inv4_high inv4_high_prefix > gpurun_out/measure_inversions.jsonl 2> gpurun_out/measure_inversions.err
cat gpurun_out/measure_inversions.jsonl
timeout 60 python scripts/pair_check.py time > gpurun_out/measure_pair_check.jsonl 2>&1; tail -3 gpurun_out/measure_pair_check.jsonl
timeout 120 python scripts/inversion_batch.py inv3_low 8 > gpurun_out/measure_batch8.jsonl 2> gpurun_out/measure_batch8.err; cat gpurun_out/measure_batch8.jsonl
timeout 300 python bench.py > gpurun_out/measure_bench.json 2> gpurun_out/measure_bench.err; cat gpurun_out/measure_bench.json
