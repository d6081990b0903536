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


def test_client_pair_key_matches_oracle_and_bootstraps(native, oracle):
    """pair key (two key bits per blind-rotation step): same bytes as the oracle's keygen, and the oracle's pair
    bootstrap with it decrypts to the table entry"""
    prm = PR.TOY_1024_L1
    ck = native.ClientKeys(prm, seed=42, threads=3, pairs=True)
    assert ck.bsk is None and ck.bskp.shape == (prm.n // 2, 3, 2, 2, prm.N)
    assert np.array_equal(ck.bskp, oracle.keygen_bsk_pairs(prm, 42, ck.s, ck.S))
    table = [(5 * m + 2) % 8 for m in range(8)]
    lut = PR.lut_polynomial([PR.encode(t, 3) for t in table], 3, prm.N)
    for m in (0, 3, 7):
        small = oracle.keyswitch(prm, ck.ksk, ck.encrypt([PR.encode(m, 3)])[0])
        out = oracle.pbs_pairs(prm, ck.bskp, lut, small)
        assert PR.decode(int(ck.phase(out)[0]), 3) == table[m]
    odd = PR.TfheParams("odd", 41, 1, 1024, 22, 1, 4, 5, 2.0 ** 24, 2.0 ** 10)
    try:
        native.ClientKeys(odd, seed=1, pairs=True)
        assert False, "expected NativeError"
    except native.NativeError as e:
        assert "even" in str(e)


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
