"""GPU parity of the pair blind rotation (two key bits per step, bmi_ctx_load_bsk_pairs): every kernel build against
the oracle's definitional restatement (oracle/tfhe_oracle.c orc_pbs_pairs), bit for bit."""
import numpy as np
import pytest

from bounty_matrix_inversion_b200 import params as PR
from test_gpu_parity import dev, host_u64, rand_field

pytestmark = pytest.mark.gpu
TOYS = [PR.TOY_1024_L1, PR.TOY_2048_L1, PR.TOY_4096, PR.TOY_8192_L1, PR.TOY_16384_L1]


@pytest.fixture(scope="module", params=TOYS, ids=lambda p: p.name)
def setup(request, native, oracle):
    prm = request.param
    keys = native.ClientKeys(prm, seed=2024, pairs=True)
    eng = native.Engine(prm, 0)
    eng.load_keys(None, keys.ksk, bskp=keys.bskp)
    yield prm, keys, eng
    eng.close()


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["latency_build", "throughput_build", "split_8cta"])
def test_pair_pbs_matches_oracle_bit_exact(setup, oracle, mode):
    import torch
    prm, keys, eng = setup
    if prm.N > 8192 and mode != 3:
        pytest.skip("N = 16384 exists only on the split kernels")
    eng.set_pbs_mode(mode)
    rng = np.random.default_rng(9)
    tables = [[(3 * m + 1) % 16 for m in range(8)], [m * m % 16 for m in range(8)]]
    luts = np.stack([PR.lut_polynomial([PR.encode(t, 4) for t in tb], 3, prm.N) for tb in tables])
    eng.load_luts(luts)
    msgs = [0, 1, 5, 7, 2, 6]
    cts = keys.encrypt([PR.encode(m, 3) for m in msgs])
    small = np.stack([oracle.keyswitch(prm, keys.ksk, c) for c in cts])
    small[4] = rand_field(rng, prm.n + 1)                           # arbitrary words, not a real encryption
    small[4, 2] = 0; small[4, 3] = 0                                # a pair that switches to (0, 0) is skipped
    small[5, 5] = 0                                                 # half a pair: that monomial is X^0
    lut_idx = np.array([0, 1, 0, 1, 1, 0], np.int32)
    out = torch.zeros((len(msgs), prm.big_dim + 1), dtype=torch.int64, device="cuda")
    idx = torch.arange(len(msgs), dtype=torch.int32, device="cuda")
    eng.pbs(dev(small), idx, dev(lut_idx), idx, out, len(msgs))
    got = host_u64(out)
    for i in range(len(msgs)):
        assert np.array_equal(got[i], oracle.pbs_pairs(prm, keys.bskp, luts[lut_idx[i]], small[i])), i
    for i in range(4):
        assert PR.decode(int(keys.phase(got[i])[0]), 4) == tables[lut_idx[i]][msgs[i]]
    eng.set_pbs_mode(0)


def test_pair_ks_pbs_host_batch(setup, oracle):
    """reference-facing host entry with the pair key, enough ciphertexts to leave the 8-CTA kernel's range"""
    prm, keys, eng = setup
    table = [(5 * m + 3) % 8 for m in range(8)]
    luts = np.stack([PR.lut_polynomial([PR.encode(t, 3) for t in table], 3, prm.N)])
    eng.load_luts(luts)
    count = 21
    msgs = [i % 8 for i in range(count)]
    cts = keys.encrypt([PR.encode(m, 3) for m in msgs])
    got = eng.ks_pbs_host(cts, np.zeros(count, np.int32))
    assert [PR.decode(int(p), 3) for p in keys.phase(got)] == [table[m] for m in msgs]
    for i in (0, 20):
        small = oracle.keyswitch(prm, keys.ksk, cts[i])
        assert np.array_equal(got[i], oracle.pbs_pairs(prm, keys.bskp, luts[0], small))


def test_pair_key_needs_one_level(native):
    prm = PR.TOY_1024
    eng = native.Engine(prm, 0)
    with pytest.raises(native.NativeError, match="one decomposition level"):
        eng.load_keys(None, np.zeros((prm.big_dim, prm.ksk_l, prm.n + 1), np.uint64),
                      bskp=np.zeros((prm.n // 2, 3, 6, 2, prm.N), np.uint64))
    eng.close()
