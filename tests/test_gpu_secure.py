"""GPU at a full-size 128-bit parameter set: bit-exact against the oracle's fast leg and the
size-independent property decrypt(PBS(enc m)) == table[m] for every message."""
import numpy as np
import pytest

from bounty_matrix_inversion_b200 import params as PR

pytestmark = pytest.mark.gpu


def test_secure_w4_all_messages(native, oracle):
    prm = PR.TfheParams("w4_n752_N2048", 752, 1, 2048, 20, 1, 3, 5,
                        PR.secure_std(752) * PR.TWO64, PR.secure_std(2048) * PR.TWO64)
    keys = native.ClientKeys(prm, seed=77)
    eng = native.Engine(prm, 0)
    eng.load_keys(keys.bsk, keys.ksk)
    w = 4
    table = [(m * 7 + 3) % 16 for m in range(16)]
    luts = np.stack([PR.lut_polynomial([PR.encode(t, w) for t in table], w, prm.N)])
    eng.load_luts(luts)
    msgs = [m % 16 for m in range(64)]
    cts = keys.encrypt([PR.encode(m, w) for m in msgs])
    got = eng.ks_pbs_host(cts, np.zeros(len(msgs), np.int32))
    assert [PR.decode(int(p), w) for p in keys.phase(got)] == [table[m] for m in msgs]
    fast = oracle.Fast(prm, keys.bsk, keys.ksk)
    want = fast.batch(luts, np.zeros(4, np.int32), cts[:4], with_ks=True, threads=4)
    assert np.array_equal(got[:4], want)
    # measured output noise stays inside the model the parameter search uses
    ph = keys.phase(got).astype(np.float64)
    err = ph - np.array([PR.encode(table[m], w) for m in msgs], dtype=np.float64)
    err = (err + 2.0 ** 63) % 2.0 ** 64 - 2.0 ** 63
    assert np.std(err / 2.0 ** 64) < 3 * PR.variance_blind_rotate(prm) ** 0.5
    eng.close()
