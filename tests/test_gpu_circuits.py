"""GPU: the reference's circuits (compiled programs in tests/golden/) executed under encryption through the
C ABI; decrypted digits must equal the reference's clear QFloat path, and the ciphertexts must equal the
CPU oracle's execution of the same program bit for bit."""
import glob
import os

import numpy as np
import pytest

from bounty_matrix_inversion_b200 import fhe, params as PR
from bounty_matrix_inversion_b200.fhe.program import Program

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
# one decomposition level where a toy set has it: those circuits run the pair blind rotation (the default)
TOY_FOR_WIDTH = {3: PR.TOY_1024, 4: PR.TOY_2048_L1, 5: PR.TOY_4096, 6: PR.TOY_8192}


def load(name):
    path = os.path.join(HERE, "golden", name + ".npz")
    z = np.load(path)
    return Program.load(path), z["golden_inputs"].astype(np.int64), z["golden_outputs"].astype(np.int64)


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_encrypted_run_matches_reference_digits(path, native):
    name = os.path.basename(path)[:-4]
    prog, x, want = load(name)
    circuit = fhe.Circuit.from_program(prog, TOY_FOR_WIDTH[prog.width])
    enc = circuit.encrypt_batch([(row,) for row in x])            # every golden input is one batch lane
    got = circuit.decrypt(circuit.run(enc))
    assert np.array_equal(np.stack(got), want)


def test_ciphertexts_equal_oracle_execution(native, oracle):
    from oracle_exec import run_program_oracle
    prog, x, want = load("qf_add_medium")
    prm = PR.TOY_1024
    circuit = fhe.Circuit.from_program(prog, prm)
    enc = circuit.encrypt(x[0])
    out = circuit.run(enc)
    ref = run_program_oracle(oracle, prog, prm, circuit.keys.bsk, circuit.keys.ksk, enc.cts)
    assert np.array_equal(out.cts, ref)
    assert np.array_equal(circuit.decrypt(out), want[0])


def test_ciphertexts_equal_oracle_execution_pair_blind_rotation(native, oracle):
    """same program, both blind rotations: each equals the oracle's execution with the same key kind bit for bit, and
    both decrypt to the reference's digits"""
    from oracle_exec import run_program_oracle
    prog, x, want = load("qf_add_medium")
    prm = PR.TOY_1024_L1
    pairs = fhe.Circuit.from_program(prog, prm)
    single = fhe.Circuit.from_program(prog, prm, configuration=fhe.Configuration(blind_rotation="single"))
    assert pairs.params.bsk_group == 2 and single.params.bsk_group == 1
    enc = pairs.encrypt(x[1])
    out = pairs.run(enc)
    ref = run_program_oracle(oracle, prog, prm, None, pairs.keys.ksk, enc.cts, keys_bskp=pairs.keys.bskp)
    assert np.array_equal(out.cts, ref)
    assert np.array_equal(pairs.decrypt(out), want[1])
    enc1 = single.encrypt(x[1])
    out1 = single.run(enc1)
    assert np.array_equal(out1.cts, run_program_oracle(oracle, prog, prm, single.keys.bsk, single.keys.ksk, enc1.cts))
    assert np.array_equal(single.decrypt(out1), want[1])
    assert not np.array_equal(out.cts, out1.cts)            # same message, different noise


def test_whole_inversion_ciphertexts_equal_oracle_execution(native, oracle):
    """the reference's 2x2 inversion, default pipeline (pair blind rotation): 3,242 bootstraps over 220 levels leave
    exactly the ciphertexts the CPU oracle computes"""
    from oracle_exec import run_program_oracle
    prog, x, want = load("inv2_low")
    prm = TOY_FOR_WIDTH[prog.width]
    circuit = fhe.Circuit.from_program(prog, prm)
    assert circuit.params.bsk_group == 2
    enc = circuit.encrypt(x[3])
    out = circuit.run(enc)
    ref = run_program_oracle(oracle, prog, prm, None, circuit.keys.ksk, enc.cts, keys_bskp=circuit.keys.bskp,
                             threads=os.cpu_count() or 4)
    assert np.array_equal(out.cts, ref)
    assert np.array_equal(circuit.decrypt(out), want[3])


def test_secure_parameters_qfloat_add(native):
    """128-bit parameter set chosen by the noise model for this circuit's width and norm"""
    prog, x, want = load("qf_add_medium")
    circuit = fhe.Circuit.from_program(prog)
    assert circuit.params.N >= 1024 and PR.failure_sigmas(circuit.params, prog.width, prog.nu2) >= 6.5
    got = circuit.decrypt(circuit.run(circuit.encrypt_batch([(row,) for row in x[:4]])))
    assert np.array_equal(np.stack(got), want[:4])


def test_compile_and_run_small_function(native):
    """fhe.Compiler -> encrypt -> run -> decrypt on a function written against the front-end"""
    rng = np.random.default_rng(1)

    def fn(x, y):
        s = x + y
        return np.concatenate(((s // 2) * (x > y), (s % 2).reshape(-1)), axis=0)

    inputset = [(rng.integers(0, 4, 3), rng.integers(0, 4, 3)) for _ in range(40)]
    circuit = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"}).compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024))
    for x, y in inputset[:5]:
        assert np.array_equal(circuit.encrypt_run_decrypt(x, y), fn(x, y))


def test_purely_leveled_circuit_needs_no_bootstrap(native):
    """edge case: a circuit without table lookups is one lincomb launch (no keyswitch, no PBS)"""
    rng = np.random.default_rng(2)
    fn = lambda x, y: 3 * x - y + np.sum(y) - 2
    inputset = [(rng.integers(-3, 4, 4), rng.integers(-3, 4, 4)) for _ in range(30)]
    circuit = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"}).compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024))
    assert circuit.statistics["pbs"] == 0
    for x, y in inputset[:4]:
        assert np.array_equal(circuit.encrypt_run_decrypt(x, y), fn(x, y))


def test_ragged_batch_and_every_kernel_build(native):
    """5 lanes (not a multiple of the keyswitch tile) through the automatic, latency, throughput and 8-CTA builds"""
    prog, x, want = load("qf_sub_medium")
    circuit = fhe.Circuit.from_program(prog, PR.TOY_1024_L1)
    enc = circuit.encrypt_batch([(row,) for row in x[:5]])
    ex = circuit.executor()
    for mode in (0, 1, 2, 3):
        ex.eng.set_pbs_mode(mode)
        got = circuit.decrypt(circuit.run(enc))
        assert np.array_equal(np.stack(got), want[:5]), mode
    ex.eng.set_pbs_mode(0)


def test_split_wide_lookup_matches_oracle_and_clear_path(native, oracle):
    """lookup one bit wider than the program: sign through the padding bit + negacyclic and cyclic half tables"""
    from oracle_exec import run_program_oracle
    from test_golden_programs import wide_sum_circuit
    fn, inputset, circuit = wide_sum_circuit()
    prm = PR.TOY_1024
    for x, y in (inputset[0], inputset[-1], inputset[-2]):
        enc = circuit.encrypt(x, y)
        out = circuit.run(enc)
        assert np.array_equal(circuit.decrypt(out), fn(x, y))
    ref = run_program_oracle(oracle, circuit.program, prm, circuit.keys.bsk, circuit.keys.ksk, enc.cts)
    assert np.array_equal(out.cts, ref)
    lanes = inputset[:9]
    got = circuit.decrypt(circuit.run(circuit.encrypt_batch(lanes)))
    assert np.array_equal(np.stack(got), np.stack([fn(x, y) for x, y in lanes]))
