#!/bin/bash
# Every single-GPU number DESIGN.md and README.md quote, in one call (about 10 minutes on one B200):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/measure_all.sh'
# Results land in gpurun_out/r2_final_*; copy what is to be tracked into profiles/.
# (8 GPUs: scripts/measure_8gpu.sh; N GPUs, inversions only: torchrun ... scripts/inversion_multi.py NAME ...)
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_final_gpu_tests.log 2>&1; tail -3 gpurun_out/r2_final_gpu_tests.log
timeout 300 python bench.py > gpurun_out/r2_final_bench_1gpu.json 2> gpurun_out/r2_final_bench_1gpu.err; tail -c 700 gpurun_out/r2_final_bench_1gpu.json
timeout 300 python scripts/inversion_multi.py inv2_low inv2_low_prefix inv3_low inv3_low_prefix inv3_medium inv4_high inv4_high_prefix > gpurun_out/r2_final_inversions_1gpu.jsonl 2> gpurun_out/r2_final_inversions_1gpu.err
cat gpurun_out/r2_final_inversions_1gpu.jsonl | cut -c1-330
SWEEP_PAIRS=1 SWEEP_MODES=0,1,2,3 SWEEP_COUNTS=1,8,16,33,74,148,296,592,1184,2368 timeout 200 python scripts/pbs_sweep.py w4 > gpurun_out/r2_final_sweep_pairs_w4.jsonl 2> gpurun_out/r2_final_sweep.err
timeout 400 python bench.py --workload inv3_medium_batch --lanes 32 --steps 1 > gpurun_out/r2_final_inv3_medium_batch_1gpu.json 2> gpurun_out/r2_final_inv3_medium_batch_1gpu.err; cat gpurun_out/r2_final_inv3_medium_batch_1gpu.json | cut -c1-900
