"""The reference's circuits, traced from its UNMODIFIED source by tests/golden/make_golden.py, against
the reference's own clear QFloat path (golden outputs): digit-for-digit.  CPU only."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from bounty_matrix_inversion_b200 import params as PR
from bounty_matrix_inversion_b200.fhe.program import Program

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
REFERENCE = "/root/reference/matrix_inversion"


def test_fixtures_present():
    names = {os.path.basename(g)[:-4] for g in GOLDEN}
    assert {"inv2_low", "inv3_low", "qf_add_medium", "qf_mul_medium", "qf_div_medium"} <= names


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_program_matches_reference_clear_path(path):
    z = np.load(path)
    prog = Program.load(path)
    got = prog.evaluate_clear(z["golden_inputs"].astype(np.int64))
    assert np.array_equal(got, z["golden_outputs"].astype(np.int64))
    assert prog.n_pbs == prog.stats["pbs"] > 0 and prog.width <= 6


def test_inversion_golden_is_close_to_scipy():
    """the README's precision table metric: error of the decoded QFloat inverse against scipy.linalg.inv"""
    import scipy.linalg
    z = np.load(os.path.join(HERE, "golden", "inv2_medium.npz"))
    qlen, ints = int(z["meta_qfloat_len"]), int(z["meta_qfloat_ints"])
    errs = []
    for M, out in zip(z["meta_matrices"], z["golden_outputs"].astype(np.int64)):
        rows = out.reshape(-1, qlen + 1)
        w = 2.0 ** (ints - 1 - np.arange(qlen))
        inv = (rows[:, :qlen] @ w * rows[:, qlen]).reshape(M.shape)
        errs.append(np.abs(inv - scipy.linalg.inv(M)).mean())
    assert np.mean(errs) < 0.05


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_reference_unit_tests_pass_on_the_shim():
    """the reference's own clear-mode test-suite (tests/test_qfloat.py) imports `concrete.fhe`; run it
    UNMODIFIED against this repo's shim"""
    env = dict(os.environ)
    compat = os.path.join(os.path.dirname(HERE), "bounty_matrix_inversion_b200", "compat")
    env["PYTHONPATH"] = os.pathsep.join([compat, REFERENCE, env.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, "tests/test_qfloat.py"], cwd="/root/reference", env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "OK" in r.stderr


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_fixture_is_reproducible_from_reference():
    """re-trace one case from the reference source and compare with the committed program"""
    sys.path.insert(0, os.path.join(HERE, "golden"))
    import importlib
    mg = importlib.import_module("make_golden")
    fn, inputset, ins, _ = mg.CASES["qf_add_medium"]()
    circ = mg.fhe.Compiler(fn, {"arrays": "encrypted", "signs": "encrypted"}).compile(inputset, mg.fhe.Configuration(tfhe_params="deferred"))
    stored = Program.load(os.path.join(HERE, "golden", "qf_add_medium.npz"))
    keys = ("pbs", "keyswitches", "levels", "tables", "width", "nu2")
    assert {k: circ.program.stats[k] for k in keys} == {k: stored.stats[k] for k in keys}
    x = np.stack([np.concatenate([a.reshape(-1), s.reshape(-1)]) for a, s in ins]).astype(np.int64)
    assert np.array_equal(circ.program.evaluate_clear(x), stored.evaluate_clear(x))


def test_encrypted_execution_on_the_oracle(oracle):
    """whole pipeline under encryption on the CPU oracle (toy parameters): encode, offsets, LUT polynomials,
    levels, slots -- decrypts to the reference's clear digits"""
    from oracle_exec import run_program_oracle
    path = os.path.join(HERE, "golden", "qf_add_medium.npz")
    z, prog = np.load(path), Program.load(path)
    prm = PR.TOY_1024
    keys = oracle.Keys(prm, seed=3)
    x = z["golden_inputs"].astype(np.int64)[0]
    cts = np.stack([oracle.encrypt_big(prm, keys.S, 3, i, PR.encode(int(m), prog.width)) for i, m in enumerate(x)])
    out = run_program_oracle(oracle, prog, prm, keys.bsk, keys.ksk, cts)
    dec = np.array([PR.decode_signed(oracle.phase(keys.S, c), prog.width) for c in out])
    assert np.array_equal(dec, z["golden_outputs"].astype(np.int64)[0])
