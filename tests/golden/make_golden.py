"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the build container only:
/root/reference does not exist on the GPU box).

For every case it
  1. traces the reference's own circuit function (qfloat_matrix_inverse or a QFloat operator written
     exactly like /root/reference/tests/test_qfloat_fhe.py's circuits) on this repo's `fhe` front-end,
     with the reference's own inputset recipe (qfloat_matrix_inversion.py:981-987), and saves the
     lowered Program;
  2. evaluates the reference's CLEAR Python path (the same functions on numpy arrays, as
     run_qfloat_inverse_python does, qfloat_matrix_inversion.py:831-845) on fresh seeded inputs and
     saves inputs + expected outputs.
Tests then check   program(clear) == expected   on CPU and   decrypt(GPU(encrypt(inputs))) == expected   on the B200.

usage: python tests/golden/make_golden.py [case ...]
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "bounty_matrix_inversion_b200", "compat"))   # `from concrete import fhe` -> this repo
sys.path.insert(0, "/root/reference/matrix_inversion")

from concrete import fhe  # noqa: E402
import qfloat_matrix_inversion as qmi  # noqa: E402  (reference, unmodified)
from qfloat import QFloat, SignedBinary, Zero  # noqa: E402  (reference, unmodified)

PRECISIONS = {"low": (23, 9, False), "medium": (31, 16, False), "mediumplus": (31, 16, True), "high": (40, 20, True)}


def inversion_case(n, precision, tensorize=False, n_golden=8, seed=0):
    qlen, qints, true_div = PRECISIONS[precision]
    rs = np.random.RandomState(seed)
    sampler = lambda: rs.randn(n, n) * 100              # normal(0,100), as main.py:119 / README
    params = [n, qlen, qints, 2, true_div, tensorize]
    inputset = [qmi.float_matrix_to_qfloat_arrays(sampler(), qlen, qints, 2) for _ in range(100)]
    fn = lambda x, y: qmi.qfloat_matrix_inverse(x, y, *params)
    mats = [sampler() for _ in range(n_golden)]
    ins = [qmi.float_matrix_to_qfloat_arrays(M, qlen, qints, 2) for M in mats]
    return fn, inputset, ins, dict(n=n, qfloat_len=qlen, qfloat_ints=qints, qfloat_base=2, true_division=true_div,
                                   matrices=np.stack(mats))


def _lists(arrays, signs, ints, base):
    return [QFloat(arrays[i, :], ints, base, True, signs[i]) for i in range(arrays.shape[0])]


def _pack(results, qlen, ints):
    """same packing as the reference's qfloat_list_to_qfloat_arrays (tests/test_qfloat_fhe.py:82-103)"""
    out = fhe.zeros((len(results), qlen + 1))
    for i, r in enumerate(results):
        if isinstance(r, QFloat):
            out[i, :-1] = r.to_array()
            out[i, -1] = r.sign
        elif isinstance(r, SignedBinary):
            out[i, ints - 1] = r.value
            out[i, -1] = r.value
        elif isinstance(r, Zero):
            pass
        else:
            out[i, ints - 1] = r
            out[i, -1] = np.sign(r)
    return out


def qfloat_op_case(op, precision, n_golden=16, seed=1):
    """one QFloat operator on a pair, the circuits of the reference's tests/test_qfloat_fhe.py"""
    qlen, qints, _ = PRECISIONS[precision]
    rs = np.random.RandomState(seed)

    def fn(arrays, signs):
        a, b = _lists(arrays, signs, qints, 2)
        if op == "add":
            res = a + b
        elif op == "sub":
            res = a - b
        elif op == "mul":
            res = a * b
        elif op == "div":
            res = a / b
        elif op == "multi":                       # test_multi_fhe: a + a + a - b, then * a
            res = (a + a + a - b) * a
        else:
            raise ValueError(op)
        return _pack([res], qlen, qints)

    def sample():
        fs = rs.uniform(0, 100, 2) * rs.choice([-1, 1], 2)
        qs = [QFloat.from_float(f, qlen, qints, 2) for f in fs]
        return (np.stack([q.to_array() for q in qs]).astype(np.int64), np.array([q.sign for q in qs], dtype=np.int64))

    inputset = [sample() for _ in range(100)]
    ins = [sample() for _ in range(n_golden)]
    return fn, inputset, ins, dict(op=op, qfloat_len=qlen, qfloat_ints=qints, qfloat_base=2)


CASES = {
    "inv2_low": lambda: inversion_case(2, "low"),
    "inv2_low_tensorized": lambda: inversion_case(2, "low", tensorize=True),
    "inv2_low_quarter_square": lambda: inversion_case(2, "low"),       # compiled with Concrete's two-lookup product lowering
    "inv2_low_prefix": lambda: inversion_case(2, "low"),               # borrow chains by parallel prefix (latency-oriented)
    "inv2_medium": lambda: inversion_case(2, "medium"),
    "inv3_low": lambda: inversion_case(3, "low", n_golden=8),
    "inv3_low_prefix": lambda: inversion_case(3, "low", n_golden=8),
    "inv3_medium": lambda: inversion_case(3, "medium", n_golden=8),
    "inv4_high": lambda: inversion_case(4, "high", n_golden=8),
    "inv4_high_prefix": lambda: inversion_case(4, "high", n_golden=8),
    "qf_add_medium": lambda: qfloat_op_case("add", "medium"),
    "qf_sub_medium": lambda: qfloat_op_case("sub", "medium"),
    "qf_mul_medium": lambda: qfloat_op_case("mul", "medium"),
    "qf_div_medium": lambda: qfloat_op_case("div", "medium"),
    "qf_multi_low": lambda: qfloat_op_case("multi", "low"),
}


def main():
    names = sys.argv[1:] or list(CASES)
    for name in names:
        t0 = time.time()
        fn, inputset, ins, meta = CASES[name]()
        comp = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"} if "inv" in name else {"arrays": "encrypted", "signs": "encrypted"})
        circuit = comp.compile(inputset, fhe.Configuration(
            tfhe_params="deferred", multiplication="quarter_square" if name.endswith("quarter_square") else "auto",
            collapse_borrows="prefix" if name.endswith("_prefix") else True))
        prog = circuit.program
        expected = np.stack([np.asarray(fn(a, s)).astype(np.int64).reshape(-1) for a, s in ins])    # reference clear path
        flat_in = np.stack([np.concatenate([np.asarray(a).reshape(-1), np.asarray(s).reshape(-1)]) for a, s in ins]).astype(np.int64)
        got = prog.evaluate_clear(flat_in)
        assert np.array_equal(got, expected), f"{name}: lowered program differs from the reference's clear path"
        path = os.path.join(HERE, name + ".npz")
        prog.save(path, golden_inputs=flat_in.astype(np.int8), golden_outputs=expected.astype(np.int8),
                  out_rows=np.array(np.asarray(fn(*ins[0])).shape, np.int64),
                  **{"meta_" + k: np.asarray(v) for k, v in meta.items()})
        print(f"{name}: {prog.stats}  -> {os.path.getsize(path) / 1e6:.2f} MB in {time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
