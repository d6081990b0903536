"""The reference's own entry class, `EncryptedMatrixInversion` (/root/reference/matrix_inversion/main.py:17-116), built and
driven on this repo's `concrete.fhe` shim.  main.py runs its benchmark loop at import time, so the class is taken from
its source text (everything above the samplers) and executed UNMODIFIED in a subprocess whose `concrete` resolves to the
shim.  CPU only: compile (trace -> lower -> parameter selection), keygen, encrypt / decrypt round trip through the C ABI
client, and the class's `run(matrix, simulate=True)` against numpy's inverse and the reference's clear QFloat path.
The encrypted `run` of the very same compiled circuit is what tests/test_gpu_circuits.py and
tests/test_gpu_secure_pairs.py execute on the GPU from the inv2_low fixture (the GPU box has no reference checkout)."""
import os
import subprocess
import sys
import textwrap

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference/matrix_inversion"

DRIVER = textwrap.dedent('''
    import sys, numpy as np
    src = open("/root/reference/matrix_inversion/main.py").read()
    cut = src.index("normal_sampler =")                      # the class and its imports; not the benchmark loop below it
    ns = {"__name__": "reference_main"}
    exec(compile(src[:cut], "main.py", "exec"), ns)
    EncryptedMatrixInversion = ns["EncryptedMatrixInversion"]
    from concrete import fhe
    import bounty_matrix_inversion_b200.fhe as ours
    assert fhe is ours, "the reference must be running on the shim"
    n = 2
    np.random.seed(7)
    sampler = lambda: np.random.randn(n, n) * 100
    emi = EncryptedMatrixInversion(n, sampler, qfloat_base=2, qfloat_len=23, qfloat_ints=9, true_division=False, tensorize=False)
    circuit = emi.circuit
    assert isinstance(circuit, fhe.Circuit) and circuit.program.n_pbs > 1000
    M = sampler()
    inv_sim = emi.run(M, simulate=True)
    err = np.abs(inv_sim - np.linalg.inv(M)).mean()
    assert err < 0.05, err
    # the reference's clear path on the same quantised input gives the same digits as the compiled circuit
    from qfloat_matrix_inversion import qfloat_matrix_inverse
    q, s = emi.quantize(M)
    clear = qfloat_matrix_inverse(q, s, n, 23, 9, 2, False, False)
    assert np.array_equal(np.asarray(circuit.simulate(q, s)), np.asarray(clear))
    # client side through the C ABI: keygen (OS entropy, like Concrete's), encrypt, decrypt of what was encrypted
    circuit.keygen()
    assert circuit.keys.deterministic is False and circuit.params.bsk_group == 2 and circuit.params.N == 2048
    enc = emi.encrypt(q, s)
    assert isinstance(enc, fhe.PublicArguments) and enc.cts.shape == (q.size + s.size, circuit.params.big_dim + 1)
    from bounty_matrix_inversion_b200 import params as PR
    ph = circuit.keys.phase(enc.cts)
    back = np.array([PR.decode_signed(int(p), circuit.program.width) for p in ph])
    assert np.array_equal(back, np.concatenate([q.reshape(-1), s.reshape(-1)]))
    print("REFERENCE_ENTRY_OK", circuit.statistics["pbs"], circuit.params.name, round(float(err), 6))
''')


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_encrypted_matrix_inversion_class_runs_on_the_shim():
    env = dict(os.environ)
    root = os.path.dirname(HERE)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(root, "bounty_matrix_inversion_b200", "compat"), REFERENCE, root,
                                         env.get("PYTHONPATH", "")])
    r = subprocess.run([sys.executable, "-c", DRIVER], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    assert "REFERENCE_ENTRY_OK" in r.stdout
