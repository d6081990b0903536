"""throughput of independent encrypted inversions run as batch lanes of one program.  usage: inversion_batch.py NAME LANES"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from bounty_matrix_inversion_b200 import fhe
from bounty_matrix_inversion_b200.fhe.program import Program

name, lanes = sys.argv[1], int(sys.argv[2])
path = os.path.join("tests", "golden", name + ".npz")
z, prog = np.load(path), Program.load(path)
c = fhe.Circuit.from_program(prog, configuration=fhe.Configuration(seed=78))
c.keygen()
x, want = z["golden_inputs"].astype(np.int64), z["golden_outputs"].astype(np.int64)
rows = [x[i % len(x)] for i in range(lanes)]
enc = c.encrypt_batch([(r,) for r in rows])
c.run(enc)                 # warm-up: builds the executor for this many lanes and captures the CUDA graph
torch.cuda.synchronize()
t0 = time.time()
out = c.run(enc)
torch.cuda.synchronize()
wall = time.time() - t0
got = np.stack(c.decrypt(out))
ok = bool(np.array_equal(got, np.stack([want[i % len(x)] for i in range(lanes)])))
print(json.dumps({"program": name, "lanes": lanes, "wall_s": wall, "inversions_per_s": lanes / wall, "s_per_inversion": wall / lanes,
                  "pbs_per_s": lanes * prog.n_pbs / wall, "digits_match_reference_clear_path": ok, "params": c.params.name}))
