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


def wide_sum_circuit():
    """a circuit whose one wide lookup source (sum of 16 bits: 17 values) is lowered to sign + two half-unit tables"""
    from bounty_matrix_inversion_b200 import fhe
    rng = np.random.default_rng(11)

    def fn(x, y):
        s = np.sum(x) + np.sum(y)
        out = fhe.zeros(3)
        out[0], out[1], out[2] = s // 3, fhe.univariate(lambda v: (v * v) % 7 - 3)(s), s % 2
        return out + x[:3] % 2

    inputset = [(rng.integers(0, 2, 8), rng.integers(0, 2, 8)) for _ in range(40)]
    inputset += [(np.zeros(8, np.int64), np.zeros(8, np.int64)), (np.ones(8, np.int64), np.ones(8, np.int64))]
    circuit = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"}).compile(
        inputset, fhe.Configuration(tfhe_params=PR.TOY_1024, split_wide=True, split_guard=0))
    return fn, inputset, circuit


def test_split_wide_lookup_under_encryption_on_the_oracle(oracle):
    """padding-bit sign bootstrap and half-scale accumulators at the ciphertext level (program.lower docstring)"""
    from oracle_exec import run_program_oracle
    fn, inputset, circuit = wide_sum_circuit()
    prog, prm = circuit.program, PR.TOY_1024
    assert prog.width == 4 and prog.stats["split_lookups"] == 3 and prog.table_half.any()
    keys = oracle.Keys(prm, seed=5)
    for x, y in (inputset[0], inputset[-1], inputset[-2], inputset[3]):        # includes sums 0 and 16
        flat = np.concatenate([x, y])
        cts = np.stack([oracle.encrypt_big(prm, keys.S, 5, i, PR.encode(int(m), prog.width)) for i, m in enumerate(flat)])
        out = run_program_oracle(oracle, prog, prm, keys.bsk, keys.ksk, cts)
        dec = np.array([PR.decode_signed(oracle.phase(keys.S, c), prog.width) for c in out])
        assert np.array_equal(dec, fn(x, y)), (x, y)


def test_encrypted_execution_on_the_oracle_pair_blind_rotation(oracle):
    """same pipeline with the pair bootstrapping key (two key bits per blind-rotation step)"""
    from oracle_exec import run_program_oracle
    path = os.path.join(HERE, "golden", "qf_add_medium.npz")
    z, prog = np.load(path), Program.load(path)
    prm = PR.TOY_1024_L1
    keys = oracle.Keys(prm, seed=3)
    bskp = oracle.keygen_bsk_pairs(prm, 3, keys.s, keys.S)
    x = z["golden_inputs"].astype(np.int64)[0]
    cts = np.stack([oracle.encrypt_big(prm, keys.S, 3, i, PR.encode(int(m), prog.width)) for i, m in enumerate(x)])
    out = run_program_oracle(oracle, prog, prm, None, keys.ksk, cts, keys_bskp=bskp)
    dec = np.array([PR.decode_signed(oracle.phase(keys.S, c), prog.width) for c in out])
    assert np.array_equal(dec, z["golden_outputs"].astype(np.int64)[0])
    # the tuned pair leg and the definitional one give the same ciphertexts through the whole program
    assert np.array_equal(out, run_program_oracle(oracle, prog, prm, None, keys.ksk, cts, keys_bskp=bskp, definitional=True))


def test_prefix_compiled_inversion_under_encryption_on_the_oracle(oracle):
    """same inversion compiled with collapse_borrows="prefix" (fewer levels, more lookups): encrypted execution on the CPU
    oracle decrypts to the same reference digits"""
    from oracle_exec import run_program_oracle
    path = os.path.join(HERE, "golden", "inv2_low_prefix.npz")
    z, prog = np.load(path), Program.load(path)
    base = Program.load(os.path.join(HERE, "golden", "inv2_low.npz"))
    assert len(prog.levels) < len(base.levels) and base.n_pbs < prog.n_pbs < 1.15 * base.n_pbs
    prm = PR.TOY_2048_L1
    keys = oracle.Keys(prm, seed=13)
    bskp = oracle.keygen_bsk_pairs(prm, 13, keys.s, keys.S)
    x = z["golden_inputs"].astype(np.int64)[1]
    cts = np.stack([oracle.encrypt_big(prm, keys.S, 13, i, PR.encode(int(m), prog.width)) for i, m in enumerate(x)])
    out = run_program_oracle(oracle, prog, prm, None, keys.ksk, cts, keys_bskp=bskp, threads=os.cpu_count() or 4)
    dec = np.array([PR.decode_signed(oracle.phase(keys.S, c), prog.width) for c in out])
    assert np.array_equal(dec, z["golden_outputs"].astype(np.int64)[1])


def test_whole_inversion_under_encryption_on_the_oracle_default_pipeline(oracle):
    """the reference's 2x2 inversion as compiled by default (packed products, same-source fusion, borrow chains three
    digits per level) executed on ENCRYPTED inputs with the pair blind rotation by the CPU oracle: decrypted digits equal
    the reference's clear path -- the same check tests/test_gpu_circuits.py makes on the B200"""
    from oracle_exec import run_program_oracle
    path = os.path.join(HERE, "golden", "inv2_low.npz")
    z, prog = np.load(path), Program.load(path)
    assert prog.width == 4 and prog.stats["collapsed_borrows"] > 400
    prm = PR.TOY_2048_L1
    keys = oracle.Keys(prm, seed=12)
    bskp = oracle.keygen_bsk_pairs(prm, 12, keys.s, keys.S)
    x = z["golden_inputs"].astype(np.int64)[2]
    cts = np.stack([oracle.encrypt_big(prm, keys.S, 12, i, PR.encode(int(m), prog.width)) for i, m in enumerate(x)])
    out = run_program_oracle(oracle, prog, prm, None, keys.ksk, cts, keys_bskp=bskp, threads=os.cpu_count() or 4)
    dec = np.array([PR.decode_signed(oracle.phase(keys.S, c), prog.width) for c in out])
    assert np.array_equal(dec, z["golden_outputs"].astype(np.int64)[2])


def test_division_with_collapsed_borrow_chains_under_encryption_on_the_oracle(oracle):
    """the reference's QFloat division (long division = nested borrow chains, two digits per level after the
    collapse) executed on encrypted data by the CPU oracle decrypts to the reference's digits"""
    from oracle_exec import run_program_oracle
    path = os.path.join(HERE, "golden", "qf_div_medium.npz")
    z, prog = np.load(path), Program.load(path)
    assert prog.stats["collapsed_borrows"] > 500 and prog.stats["levels"] < 600
    prm = PR.TOY_1024_L1
    keys = oracle.Keys(prm, seed=9)
    for lane in (0, 5):
        x = z["golden_inputs"].astype(np.int64)[lane]
        cts = np.stack([oracle.encrypt_big(prm, keys.S, 9, 1000 * lane + i, PR.encode(int(m), prog.width)) for i, m in enumerate(x)])
        out = run_program_oracle(oracle, prog, prm, keys.bsk, keys.ksk, cts)
        dec = np.array([PR.decode_signed(oracle.phase(keys.S, c), prog.width) for c in out])
        assert np.array_equal(dec, z["golden_outputs"].astype(np.int64)[lane]), lane
