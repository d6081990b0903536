"""one PBS launch in a chosen kernel build (for ncu).  usage: pbs_lat.py <set> <count> <mode> [pairs]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
from bounty_matrix_inversion_b200 import native, params as PR
from pbs_sweep import SETS

prm = SETS[sys.argv[1]]
count, mode = int(sys.argv[2]), int(sys.argv[3])
pairs = len(sys.argv) > 4 and sys.argv[4] == "pairs"
keys = native.ClientKeys(prm, seed=5, pairs=pairs)
eng = native.Engine(prm, 0)
eng.load_keys(keys.bsk, keys.ksk, bskp=keys.bskp)
eng.set_pbs_mode(mode)
eng.load_luts(np.stack([PR.lut_polynomial([PR.encode(t, 3) for t in range(8)], 3, prm.N)]))
cts = np.tile(keys.encrypt([PR.encode(i % 8, 3) for i in range(8)]), (count // 8 + 1, 1))[:count]
big = torch.from_numpy(cts.view(np.int64)).cuda()
small = torch.zeros((count, prm.n + 1), dtype=torch.int64, device="cuda")
out = torch.zeros((count, prm.big_dim + 1), dtype=torch.int64, device="cuda")
idx = torch.arange(count, dtype=torch.int32, device="cuda")
lut = torch.zeros(count, dtype=torch.int32, device="cuda")
for _ in range(2):
    eng.keyswitch(big, small, count)
    eng.pbs(small, idx, lut, idx, out, count)
torch.cuda.synchronize()
print([PR.decode(int(p), 3) for p in keys.phase(out[:8].cpu().numpy().view(np.uint64))])
