"""wall time of encrypted inversions with every level's lookups sharded over the ranks (torchrun), or on one GPU.
usage: [torchrun --nproc-per-node N] scripts/inversion_multi.py NAME ...     (NAME: a program in tests/golden)"""
import json
import os
import sys

sys.path.insert(0, ".")
import torch
import torch.distributed as dist

import bench
from bounty_matrix_inversion_b200 import fhe, params as PR

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for name in sys.argv[1:]:
    rec = bench.time_inversion(name, fhe, PR, local, rank, world, dist if world > 1 else None)
    if rank == 0:
        print(json.dumps(rec), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
