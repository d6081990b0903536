"""Raw keyswitch / PBS batch-size sweep on one GPU (CUDA-event timing).  usage: pbs_sweep.py [w ...]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bounty_matrix_inversion_b200 import native, params as PR

SETS = {
    "w4": PR.TfheParams("w4_n752_N2048", 752, 1, 2048, 20, 1, 3, 5, PR.secure_std(752) * PR.TWO64, PR.secure_std(2048) * PR.TWO64),
    "w5": PR.TfheParams("w5_n840_N4096", 840, 1, 4096, 23, 1, 4, 4, PR.secure_std(840) * PR.TWO64, PR.secure_std(4096) * PR.TWO64),
    "w6": PR.TfheParams("w6_n912_N8192", 912, 1, 8192, 25, 1, 4, 4, PR.secure_std(912) * PR.TWO64, PR.secure_std(8192) * PR.TWO64),
}


def main():
    names = sys.argv[1:] or ["w4", "w6"]
    for name in names:
        prm = SETS[name]
        t0 = time.time()
        pairs = os.environ.get("SWEEP_PAIRS", "0") == "1"
        keys = native.ClientKeys(prm, seed=5, pairs=pairs)
        t_keys = time.time() - t0
        eng = native.Engine(prm, 0)
        t0 = time.time()
        eng.load_keys(keys.bsk, keys.ksk, bskp=keys.bskp)
        t_load = time.time() - t0
        w = 3
        luts = np.stack([PR.lut_polynomial([PR.encode(t, w) for t in range(8)], w, prm.N)])
        eng.load_luts(luts)
        modes = [int(v) for v in os.environ.get("SWEEP_MODES", "0,1,2,3").split(",")]
        counts = [int(v) for v in os.environ.get("SWEEP_COUNTS", "1,8,74,148,296,592,1184").split(",")]
        for mode, count in [(m, c) for m in modes for c in counts]:
            eng.set_pbs_mode(mode)
            cts = keys.encrypt([PR.encode(i % 8, w) for i in range(min(count, 16))])
            cts = np.tile(cts, (count // len(cts) + 1, 1))[:count]
            big = torch.from_numpy(cts.view(np.int64)).cuda()
            small = torch.zeros((count, prm.n + 1), dtype=torch.int64, device="cuda")
            out = torch.zeros((count, prm.big_dim + 1), dtype=torch.int64, device="cuda")
            idx = torch.arange(count, dtype=torch.int32, device="cuda")
            lut = torch.zeros(count, dtype=torch.int32, device="cuda")
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            for it in range(3):
                ev[0].record()
                eng.keyswitch(big, small, count)
                ev[1].record()
                eng.pbs(small, idx, lut, idx, out, count)
                ev[2].record()
                torch.cuda.synchronize()
            ks_ms, pbs_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
            dec = [PR.decode(int(p), w) for p in keys.phase(out[:4].cpu().numpy().view(np.uint64))]
            print(json.dumps({"set": prm.name, "pairs": pairs, "mode": mode, "count": count, "ks_ms": round(ks_ms, 3), "pbs_ms": round(pbs_ms, 3),
                              "pbs_per_s": round(count / (pbs_ms + ks_ms) * 1e3, 1), "dec": dec,
                              "keygen_s": round(t_keys, 2), "load_s": round(t_load, 2)}), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
