"""A/B of TMA staging in the pair blind rotation (latency build, bmi_ctx_set_tma_stage): the three GGSWs of each step
brought into shared memory by cp.async.bulk one step ahead, against per-thread coalesced ld.global.nc.
Prints bit-exactness against the oracle on toy sets, then CUDA-event timings on the 128-bit set w4."""
import json
import sys

sys.path.insert(0, ".")
import numpy as np
import torch

from bounty_matrix_inversion_b200 import native, params as PR
from oracle import oracle as orc

for prm in (PR.TOY_1024_L1, PR.TOY_2048_L1):
    keys = native.ClientKeys(prm, seed=2024, pairs=True)
    eng = native.Engine(prm, 0)
    eng.load_keys(None, keys.ksk, bskp=keys.bskp)
    luts = np.stack([PR.lut_polynomial([PR.encode((3 * m + 1) % 16, 4) for m in range(8)], 3, prm.N)])
    eng.load_luts(luts)
    cts = keys.encrypt([PR.encode(m, 3) for m in (0, 1, 5, 7, 2)])
    small = np.stack([orc.keyswitch(prm, keys.ksk, c) for c in cts])
    small[4, 2] = 0; small[4, 3] = 0          # a whole pair that switches to 0 is skipped
    want = np.stack([orc.pbs_pairs(prm, keys.bskp, luts[0], s) for s in small])
    eng.set_pbs_mode(1)
    res = {}
    for stage in (False, True):
        eng.set_tma_stage(stage)
        out = torch.zeros((5, prm.big_dim + 1), dtype=torch.int64, device="cuda")
        idx = torch.arange(5, dtype=torch.int32, device="cuda")
        eng.pbs(torch.from_numpy(small.view(np.int64)).cuda(), idx, torch.zeros(5, dtype=torch.int32, device="cuda"), idx, out, 5)
        torch.cuda.synchronize()
        res["staged" if stage else "direct"] = bool(np.array_equal(out.cpu().numpy().view(np.uint64), want))
    print(json.dumps({"params": prm.name, "bit_exact": res}), flush=True)
    eng.close()

prm = PR.for_width(4, 400.0, bsk_group=2)
keys = native.ClientKeys(prm, seed=5, pairs=True)
eng = native.Engine(prm, 0)
eng.load_keys(None, keys.ksk, bskp=keys.bskp)
eng.load_luts(np.stack([PR.lut_polynomial([PR.encode(t, 4) for t in range(16)], 4, prm.N)]))
eng.set_pbs_mode(1)
rng = np.random.default_rng(1)
row = {"params": prm.name}
for count in (1, 8, 74, 148):
    small = torch.from_numpy((rng.integers(0, 2 ** 62, size=(count, prm.n + 1), dtype=np.uint64)).view(np.int64)).cuda()
    out = torch.zeros((count, prm.big_dim + 1), dtype=torch.int64, device="cuda")
    idx = torch.arange(count, dtype=torch.int32, device="cuda")
    lut = torch.zeros(count, dtype=torch.int32, device="cuda")
    outs = {}
    for stage in (False, True):
        eng.set_tma_stage(stage)
        for _ in range(2):
            eng.pbs(small, idx, lut, idx, out, count)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            eng.pbs(small, idx, lut, idx, out, count)
        b.record()
        torch.cuda.synchronize()
        row[f"{'staged' if stage else 'direct'}_{count}_ms"] = round(a.elapsed_time(b) / 5, 3)
        outs[stage] = out.clone()
    row[f"same_output_{count}"] = bool(torch.equal(outs[False], outs[True]))
print(json.dumps(row), flush=True)
