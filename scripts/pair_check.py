"""pair blind rotation on the GPU against the oracle, every kernel build, every l = 1 toy set; then timing on a
128-bit set.  usage: pair_check.py [time]"""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch

from bounty_matrix_inversion_b200 import native, params as PR
from oracle import oracle as orc

P = PR.P


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64) if a.dtype == np.uint64 else np.ascontiguousarray(a)).cuda()


def rand_field(rng, shape):
    return (rng.integers(0, 2 ** 63, size=shape, dtype=np.uint64) * np.uint64(2)
            + rng.integers(0, 2, size=shape, dtype=np.uint64)) % np.uint64(P)


all_ok = True
for prm in (PR.TOY_1024_L1, PR.TOY_2048_L1, PR.TOY_4096, PR.TOY_8192_L1, PR.TOY_16384_L1):
    keys = native.ClientKeys(prm, seed=2024, pairs=True)
    eng = native.Engine(prm, 0)
    if prm.N <= 8192:
        eng.set_pbs_mode(5)          # builds the key layout of the two-points-per-thread kernel as well
    eng.load_keys(None, keys.ksk, bskp=keys.bskp)
    rng = np.random.default_rng(9)
    tables = [[(3 * m + 1) % 16 for m in range(8)], [m * m % 16 for m in range(8)]]
    luts = np.stack([PR.lut_polynomial([PR.encode(t, 4) for t in tb], 3, prm.N) for tb in tables])
    eng.load_luts(luts)
    msgs = [0, 1, 5, 7, 2]
    cts = keys.encrypt([PR.encode(m, 3) for m in msgs])
    small = np.stack([orc.keyswitch(prm, keys.ksk, c) for c in cts])
    small[4] = rand_field(rng, prm.n + 1)
    small[4, 2] = 0; small[4, 3] = 0          # a whole pair that switches to 0 is skipped
    small[3, 5] = 0                           # half a pair
    lut_idx = np.array([0, 1, 0, 1, 1], np.int32)
    want = np.stack([orc.pbs_pairs(prm, keys.bskp, luts[lut_idx[i]], small[i]) for i in range(5)])
    dec_ok = [PR.decode(int(keys.phase(want[i])[0]), 4) == tables[lut_idx[i]][msgs[i]] for i in range(4)]
    for mode in (1, 2, 3, 5):
        if prm.N > 8192 and mode != 3:
            continue
        eng.set_pbs_mode(mode)
        out = torch.zeros((5, prm.big_dim + 1), dtype=torch.int64, device="cuda")
        idx = torch.arange(5, dtype=torch.int32, device="cuda")
        eng.pbs(dev(small), idx, dev(lut_idx), idx, out, 5)
        torch.cuda.synchronize()
        got = out.cpu().numpy().view(np.uint64)
        bad = [int((got[i] != want[i]).sum()) for i in range(5)]
        dec = [PR.decode(int(keys.phase(got[i])[0]), 4) for i in range(4)]
        ok = not any(bad)
        all_ok &= ok
        print(json.dumps({"params": prm.name, "mode": mode, "bit_exact": ok, "mismatching_words": bad, "decoded": dec,
                          "expected": [tables[lut_idx[i]][msgs[i]] for i in range(4)], "oracle_decodes": dec_ok}), flush=True)
    eng.close()

if len(sys.argv) > 1 and all_ok:
    for w, counts in ((4, (1, 14, 592, 2368)), (5, (1, 296))):
        prm = PR.for_width(w, 400.0)
        res = {"params": prm.name}
        for pairs in (False, True):
            keys = native.ClientKeys(prm, seed=5, pairs=pairs)
            eng = native.Engine(prm, 0)
            eng.load_keys(keys.bsk, keys.ksk, bskp=keys.bskp)
            eng.load_luts(np.stack([PR.lut_polynomial([PR.encode(t, w) for t in range(1 << w)], w, prm.N)]))
            for count in counts:
                small = dev(rand_field(np.random.default_rng(1), (count, prm.n + 1)))
                out = torch.zeros((count, prm.big_dim + 1), dtype=torch.int64, device="cuda")
                idx = torch.arange(count, dtype=torch.int32, device="cuda")
                lut = torch.zeros(count, dtype=torch.int32, device="cuda")
                for _ in range(2):
                    eng.pbs(small, idx, lut, idx, out, count)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                reps = 3
                for _ in range(reps):
                    eng.pbs(small, idx, lut, idx, out, count)
                b.record()
                torch.cuda.synchronize()
                ms = a.elapsed_time(b) / reps
                res[f"{'pairs' if pairs else 'single'}_{count}"] = {"ms": round(ms, 3), "pbs_per_s": round(count / ms * 1e3, 1)}
            eng.close()
        print(json.dumps(res), flush=True)
print("ALL_OK" if all_ok else "MISMATCH")
