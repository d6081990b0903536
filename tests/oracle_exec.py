"""Test helper: runs a levelled Program on ENCRYPTED data with the CPU oracle (lincomb ->
keyswitch -> PBS per level).  Test infrastructure only: it is the checker for the GPU executor."""
import numpy as np

from bounty_matrix_inversion_b200 import params as PR


def run_program_oracle(orc, prog, prm, keys_bsk, keys_ksk, input_cts, threads=8, keys_bskp=None, definitional=False):
    """input_cts [n_inputs][kN+1] uint64 -> output ciphertexts [n_out][kN+1].
    keys_bskp: pair bootstrapping key -> every bootstrap is the oracle's two-bits-per-step blind rotation.
    definitional: pair bootstraps by tfhe_oracle.c's orc_pbs_pairs instead of the tuned leg (slow; small programs)"""
    W1 = prm.big_dim + 1
    fast = orc.Fast(prm, keys_bsk, keys_ksk, bskp=keys_bskp) if not definitional else None
    luts = prog.lut_polynomials(prm.N)
    vals = np.zeros((prog.n_slots, W1), np.uint64)
    vals[prog.input_slots] = input_cts
    W = prog.width
    for lv in prog.levels:
        n_ks = len(lv.konst)
        big = np.zeros((n_ks, W1), np.uint64)
        for r in range(n_ks):
            a, b = lv.row_ptr[r], lv.row_ptr[r + 1]
            big[r] = orc.lincomb(vals, lv.idx[a:b], lv.coef[a:b], PR.encode(int(lv.konst[r]), W))
        if fast is not None:
            outs = fast.batch(luts, lv.job_lut, big[lv.job_ks], with_ks=True, threads=threads)
        else:
            small = [orc.keyswitch(prm, keys_ksk, b) for b in big]
            outs = np.stack([orc.pbs_pairs(prm, keys_bskp, luts[t], small[r]) for r, t in zip(lv.job_ks, lv.job_lut)])
        vals[lv.job_out] = outs
    res = np.zeros((len(prog.out_konst), W1), np.uint64)
    for r in range(len(prog.out_konst)):
        a, b = prog.out_row_ptr[r], prog.out_row_ptr[r + 1]
        res[r] = orc.lincomb(vals, prog.out_idx[a:b], prog.out_coef[a:b], PR.encode(int(prog.out_konst[r]), W))
    return res
