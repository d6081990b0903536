#!/bin/bash
# Everything that needs the 8 GPUs of one box, in one call:
#   /usr/local/graft/bin/gpurun --gpus 8 --timeout 1500 -- 'bash scripts/measure_8gpu.sh'
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29601 bench.py --gpus 8 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err; tail -c 600 gpurun_out/r2_bench_8gpu.json
timeout 400 $TR --master-port 29602 scripts/inversion_multi.py inv3_low_prefix inv3_medium inv4_high_prefix > gpurun_out/r2_inversions_8gpu.jsonl 2> gpurun_out/r2_inversions_8gpu.err; cat gpurun_out/r2_inversions_8gpu.jsonl
timeout 120 $TR --master-port 29603 bench.py --gpus 8 --workload pbs_sweep > gpurun_out/r2_pbs_sweep_8gpu.json 2> gpurun_out/r2_pbs_sweep_8gpu.err; tail -c 400 gpurun_out/r2_pbs_sweep_8gpu.json
timeout 400 $TR --master-port 29604 bench.py --gpus 8 --workload inv3_medium_batch --lanes 32 --steps 1 > gpurun_out/r2_inv3_medium_batch_8gpu.json 2> gpurun_out/r2_inv3_medium_batch_8gpu.err; cat gpurun_out/r2_inv3_medium_batch_8gpu.json
timeout 600 $TR --master-port 29605 bench.py --gpus 8 --pairs 512 --steps 1 --warmup 0 --no-cpu --no-e2e --inversion none > gpurun_out/r2_bench_8gpu_full4096pairs.json 2> gpurun_out/r2_bench_8gpu_full4096pairs.err; tail -c 600 gpurun_out/r2_bench_8gpu_full4096pairs.json
