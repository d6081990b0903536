"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol
include/bmi_tfhe.h declares, and its client functions (keygen / encrypt / phase) are
bit-identical to the oracle's restatement of the same spec.  No GPU compute is called."""
import ctypes
import os
import re

import numpy as np

from bounty_matrix_inversion_b200 import params as PR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native):
    header = open(os.path.join(ROOT, "include", "bmi_tfhe.h")).read()
    declared = set(re.findall(r"\b(bmi_[a-z0-9_]+)\s*\(", header))
    assert declared == set(native.exported_symbols())
    lib = native.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.bmi_version()


def test_invalid_arguments_fail_loudly(native):
    bad = native.BmiParams(0, 1, 1000, 8, 3, 4, 5, 1.0, 1.0)
    s = np.zeros(8, np.uint64)
    rc = native.lib().bmi_keygen_lwe(ctypes.byref(bad), 1, s.ctypes.data_as(ctypes.c_void_p))
    assert rc == -1 and b"invalid" in native.lib().bmi_last_error()


def test_client_keys_match_oracle(native, oracle):
    for prm in (PR.TOY_1024, PR.TfheParams("t", 20, 1, 2048, 12, 2, 5, 3, 2.0 ** 30, 2.0 ** 20)):
        ck = native.ClientKeys(prm, seed=42, threads=3)
        ok = oracle.Keys(prm, seed=42)
        assert np.array_equal(ck.s, ok.s) and np.array_equal(ck.S, ok.S)
        assert np.array_equal(ck.ksk, ok.ksk)
        assert np.array_equal(ck.bsk, ok.bsk)
        assert 0.3 < ck.S.mean() < 0.7 and ck.bsk.max() < PR.P


def test_client_encrypt_phase_match_oracle(native, oracle):
    prm = PR.TOY_1024
    ck = native.ClientKeys(prm, seed=7, evaluation_keys=False)
    pts = [PR.encode(m, 3) for m in (0, 1, 7, -3)]
    cts = ck.encrypt(pts, ct_index0=100)
    for i, pt in enumerate(pts):
        assert np.array_equal(cts[i], oracle.encrypt_big(prm, ck.S, 7, 100 + i, pt))
    ph = ck.phase(cts)
    assert [PR.decode_signed(int(p), 3) for p in ph] == [0, 1, 7, -3]
    assert [PR.decode(int(p), 3) for p in ph] == [0, 1, 7, 13]
    assert [int(p) for p in ph] == [oracle.phase(ck.S, c) for c in cts]


def test_missing_library_is_an_error(native, monkeypatch, tmp_path):
    from bounty_matrix_inversion_b200 import _build
    monkeypatch.setattr(_build, "LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(native, "_lib", None)
    try:
        native.lib()
        assert False, "expected NativeError"
    except native.NativeError as e:
        assert "no CPU fallback" in str(e)
