"""where an encrypted inversion's wall time goes: PBS kernel time by level size, everything else.  usage: NAME"""
import json
import os
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from bounty_matrix_inversion_b200 import fhe
from bounty_matrix_inversion_b200.fhe.program import Program

name = sys.argv[1]
path = os.path.join("tests", "golden", name + ".npz")
z, prog = np.load(path), Program.load(path)
c = fhe.Circuit.from_program(prog, configuration=fhe.Configuration(seed=77))
c.keygen()
ex = c.executor()
enc = c.encrypt(z["golden_inputs"].astype(np.int64)[0])
c.run(enc)
torch.cuda.synchronize()
ex.profile = []
t0 = time.time()
out = c.run(enc)
torch.cuda.synchronize()
wall = time.time() - t0
prof = ex.profile
ex.profile = None
buckets = {}
for (a, b, jobs) in prof:
    k = 1 if jobs <= 1 else 2 ** int(np.ceil(np.log2(jobs)))
    d = buckets.setdefault(k, [0, 0.0, 0])
    d[0] += 1; d[1] += a.elapsed_time(b); d[2] += jobs
pbs_ms = sum(v[1] for v in buckets.values())
print(json.dumps({"program": name, "wall_s": wall, "pbs_kernel_s": pbs_ms / 1e3, "other_s": wall - pbs_ms / 1e3, "levels": len(prof),
                  "by_level_size(<=jobs: levels, ms, jobs)": {k: [v[0], round(v[1], 1), v[2]] for k, v in sorted(buckets.items())}}))
# untimed python-only pass: how long the host loop takes to ISSUE the work
ex.profile = None
t0 = time.time()
ex.run_device(1)
issue = time.time() - t0
torch.cuda.synchronize()
print(json.dumps({"host_issue_s": issue}))
