"""wall time of encrypted inversions (compiled programs in tests/golden) on one GPU.  usage: inversion_time.py NAME ..."""
import json
import sys

sys.path.insert(0, ".")
import torch

import bench
from bounty_matrix_inversion_b200 import fhe, params as PR

torch.cuda.set_device(0)
for name in sys.argv[1:]:
    print(json.dumps(bench.time_inversion(name, fhe, PR, 0, 0, 1, None)), flush=True)
