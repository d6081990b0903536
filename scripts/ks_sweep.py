import json, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
from bounty_matrix_inversion_b200 import native, params as PR
prm = PR.for_width(4, 400.0, bsk_group=2)
keys = native.ClientKeys(prm, seed=5, evaluation_keys=True, pairs=True)
eng = native.Engine(prm, 0)
eng.load_keys(None, keys.ksk, bskp=keys.bskp)
res = {}
for count in (4, 8, 16, 26, 33, 48, 74, 148, 592):
    cts = np.tile(keys.encrypt([PR.encode(i % 16, 4) for i in range(16)]), (count // 16 + 1, 1))[:count]
    big = torch.from_numpy(cts.view(np.int64)).cuda()
    small = torch.zeros((count, prm.n + 1), dtype=torch.int64, device="cuda")
    for _ in range(3):
        eng.keyswitch(big, small, count)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        eng.keyswitch(big, small, count)
    b.record(); torch.cuda.synchronize()
    res[count] = round(a.elapsed_time(b) / 20, 4)
    chk = int(small[:4].sum().item())
print(json.dumps({"ctas_per_sm": os.environ.get("BMI_KS_CTAS_PER_SM", "3"), "ks_ms": res, "chk": chk}))
