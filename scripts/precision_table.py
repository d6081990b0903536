"""README Table 1 of the reference (precision against scipy.linalg.inv) reproduced through the compiled programs:
mean error, big-error rate (mean abs error > 1), plus the share of inputs that leave the value ranges seen on the
100-sample compile-time inputset (those lanes would decrypt differently from the clear path, for Concrete as well).
usage: precision_table.py [count]"""
import json
import os
import sys

import numpy as np
import scipy.linalg

sys.path.insert(0, ".")
from bounty_matrix_inversion_b200.fhe.program import Program

README = {"inv2_low": (8.19e-2, 0.04), "inv2_medium": (7.6e-3, 0.02), "inv3_low": (9.93e-2, 1.34), "inv3_medium": (4.43e-2, 0.4),
          "inv4_high": (8.6e-6, 0.0)}


def quantise(M, qlen, ints):
    mag = np.floor(np.abs(M) * 2.0 ** (qlen - ints)).astype(np.int64)
    digits = (mag[..., None] >> np.arange(qlen - 1, -1, -1)) & 1
    signs = np.where(M < 0, -1, 1).astype(np.int64)
    return np.concatenate([digits.reshape(M.shape[0], -1), signs.reshape(M.shape[0], -1)], axis=1)


def main():
    count = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    rs = np.random.RandomState(12345)
    for name in ("inv2_low", "inv2_medium", "inv3_low", "inv3_medium", "inv4_high"):
        path = os.path.join("tests", "golden", name + ".npz")
        z, prog = np.load(path), Program.load(path)
        n, qlen, ints = int(z["meta_n"]), int(z["meta_qfloat_len"]), int(z["meta_qfloat_ints"])
        cnt = count if prog.n_pbs < 100000 else max(50, count // 10)
        M = rs.randn(cnt, n, n) * 100
        x = quantise(M.reshape(cnt, n * n), qlen, ints)
        out, bad = prog.evaluate_clear(x, strict=False)
        rows = out.reshape(cnt, n * n, qlen + 1)
        w = 2.0 ** (ints - 1 - np.arange(qlen))
        inv = (rows[..., :qlen] @ w * rows[..., qlen]).reshape(cnt, n, n)
        err = np.array([np.abs(inv[i] - scipy.linalg.inv(M[i])).mean() for i in range(cnt)])
        good = ~bad
        print(json.dumps({"program": name, "matrices": cnt, "mean_error": float(err[good].mean()),
                          "big_error_rate_pct": float(100 * (err[good] > 1).mean()), "out_of_inputset_range_pct": float(100 * bad.mean()),
                          "readme_mean_error": README[name][0], "readme_big_error_rate_pct": README[name][1]}), flush=True)


if __name__ == "__main__":
    main()
