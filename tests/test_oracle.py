"""CPU tests that pin the oracle: field arithmetic against its definition, the transform
against the schoolbook product, the fast leg against the definitional leg, and the
bootstrap against the property the reference itself tests (decrypted output == table[input])."""
import numpy as np
import pytest

from bounty_matrix_inversion_b200 import params as PR

P = PR.P
rng = np.random.default_rng(1234)


def rand_field(size):
    return (rng.integers(0, 2 ** 63, size=size, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=size, dtype=np.uint64)) % np.uint64(P)


def test_field_mul_matches_definition(oracle):
    L = oracle.lib()
    edge = [0, 1, 2, P - 1, P - 2, 2 ** 32, 2 ** 32 - 1, 2 ** 63, 2 ** 64 - 2 ** 32, 0xFFFFFFFF00000000]
    vals = edge + [int(v) for v in rand_field(200)]
    for a in vals[:40]:
        for b in vals:
            assert L.orc_fmul(a % P, b % P) == (a % P) * (b % P) % P == L.orc_fmul_def(a % P, b % P)
    assert L.orc_fpow(7, (P - 1) // 2) == P - 1          # 7 generates the multiplicative group
    assert L.orc_fpow(2, 96) == P - 1                   # 2^96 = -1
    assert L.orc_fpow(2, 64) == 2 ** 32 - 1


@pytest.mark.parametrize("N", [8, 64, 256])
def test_negacyclic_product_matches_schoolbook(oracle, N):
    a, b = rand_field(N), rand_field(N)
    assert np.array_equal(oracle.negacyclic_mul(a, b), oracle.negacyclic_mul(a, b, schoolbook=True))
    x = np.zeros(N, np.uint64); x[1] = 1            # X * X^(N-1) = -1
    y = np.zeros(N, np.uint64); y[N - 1] = 1
    want = np.zeros(N, np.uint64); want[0] = P - 1
    assert np.array_equal(oracle.negacyclic_mul(x, y), want)


def test_decomposition_recomposes(oracle):
    for bl, l in [(4, 5), (8, 3), (23, 1), (15, 2), (6, 3), (2, 6)]:
        for x in [0, 1, P - 1, 2 ** 63, 2 ** 63 - 1] + [int(v) for v in rand_field(50)]:
            d = oracle.decompose(x, bl, l)
            assert all(-(1 << (bl - 1)) <= v < (1 << (bl - 1)) for v in d)
            rec = sum(v << (64 - (j + 1) * bl) for j, v in enumerate(d))
            err = (rec - x) % (1 << 64)
            err = min(err, (1 << 64) - err)
            assert err <= 1 << (63 - bl * l), (bl, l, x, d)


def test_modswitch(oracle):
    L = oracle.lib()
    for logN in (10, 13):
        two_n = 2 << logN
        for x in [0, 1, 2 ** 63, 2 ** 64 - 1, P - 1] + [int(v) for v in rand_field(50)]:
            assert L.orc_modswitch(x, logN) == ((x * two_n + 2 ** 63) >> 64) % two_n


def test_gaussian_moments(oracle):
    L = oracle.lib()
    sigma = 2.0 ** 40
    xs = np.array([L.orc_rng_gauss(7, 4, i, sigma) for i in range(20000)], dtype=np.uint64)
    signed = np.where(xs > P // 2, xs.astype(np.float64) - float(P), xs.astype(np.float64))
    assert abs(signed.mean()) < 0.05 * sigma
    assert abs(signed.std() / sigma - 1) < 0.03


@pytest.fixture(scope="module")
def toy(oracle):
    prm = PR.TOY_1024
    return prm, oracle.Keys(prm, seed=42)


def test_encrypt_phase_roundtrip(oracle, toy):
    prm, keys = toy
    for w in (1, 3, 6):
        for m in (0, 1, (1 << w) - 1, -1, -(1 << (w - 1))):
            ct = oracle.encrypt_big(prm, keys.S, 5, 17 + w, PR.encode(m, w))
            assert PR.decode_signed(oracle.phase(keys.S, ct), w) == m


def test_keyswitch_keeps_message(oracle, toy):
    prm, keys = toy
    for m in range(8):
        ct = oracle.encrypt_big(prm, keys.S, 5, m, PR.encode(m, 3))
        small = oracle.keyswitch(prm, keys.ksk, ct)
        assert PR.decode(oracle.phase(keys.s, small), 3) == m


def test_pbs_applies_table(oracle, toy):
    """decrypt(PBS(enc m)) == table[m]: the property the reference's FHE tests rely on
    (/root/reference/tests/test_qfloat_fhe.py compares decrypted results to clear results)"""
    prm, keys = toy
    w = 3
    table = [(5 * m * m + 3) % 16 for m in range(1 << w)]       # arbitrary function, 4-bit outputs
    lut = PR.lut_polynomial([PR.encode(t, 4) for t in table], w, prm.N)
    for m in range(1 << w):
        ct = oracle.encrypt_big(prm, keys.S, 9, m, PR.encode(m, w))
        out = oracle.pbs(prm, keys.bsk, lut, oracle.keyswitch(prm, keys.ksk, ct))
        assert PR.decode(oracle.phase(keys.S, out), 4) == table[m]


def test_pbs_negacyclic_half(oracle, toy):
    """a message with the padding bit set reads the negated table (why signed inputs get an offset)"""
    prm, keys = toy
    w = 2
    lut = PR.lut_polynomial([PR.encode(t, 3) for t in (1, 2, 3, 1)], w, prm.N)
    ct = oracle.encrypt_big(prm, keys.S, 9, 0, PR.encode(5, w))          # 5 = 4 + 1: padding bit set
    out = oracle.pbs(prm, keys.bsk, lut, oracle.keyswitch(prm, keys.ksk, ct))
    assert PR.decode_signed(oracle.phase(keys.S, out), 3) == -2


def test_fast_leg_matches_definition(oracle, toy):
    prm, keys = toy
    fast = oracle.Fast(prm, keys.bsk, keys.ksk)
    lut = PR.lut_polynomial([PR.encode(t, 3) for t in range(8)], 3, prm.N)
    cts = np.stack([oracle.encrypt_big(prm, keys.S, 3, i, PR.encode(i % 8, 3)) for i in range(6)])
    for ct in cts[:3]:
        small = oracle.keyswitch(prm, keys.ksk, ct)
        assert np.array_equal(fast.keyswitch(ct), small)
        assert np.array_equal(fast.pbs(lut, small), oracle.pbs(prm, keys.bsk, lut, small))
    outs = fast.batch(lut[None, :], np.zeros(6, np.int32), cts, with_ks=True, threads=3)
    for i in range(6):
        assert np.array_equal(outs[i], oracle.pbs(prm, keys.bsk, lut, oracle.keyswitch(prm, keys.ksk, cts[i])))


def test_lincomb(oracle, toy):
    prm, keys = toy
    w = 4
    msgs = [3, 5, 1]
    cts = np.stack([oracle.encrypt_big(prm, keys.S, 11, i, PR.encode(m, w)) for i, m in enumerate(msgs)])
    out = oracle.lincomb(cts, [0, 1, 2], [2, -1, 3], PR.encode(4, w))
    assert PR.decode_signed(oracle.phase(keys.S, out), w) == 2 * 3 - 5 + 3 * 1 + 4


def test_fast_pair_leg_matches_definitional_pair_bootstrap(oracle):
    """tfhe_oracle_fast.c's pair blind rotation (monomials in the transform domain, as the kernels do it) against
    tfhe_oracle.c's orc_pbs_pairs (monomials applied to each product in the coefficient domain), incl. two levels"""
    from bounty_matrix_inversion_b200 import params as PR
    for prm in (PR.TOY_1024_L1, PR.TOY_2048):
        keys = oracle.Keys(prm, seed=21)
        bskp = oracle.keygen_bsk_pairs(prm, 21, keys.s, keys.S)
        fast = oracle.Fast(prm, None, keys.ksk, bskp=bskp)
        table = [(3 * m + 2) % 8 for m in range(8)]
        lut = PR.lut_polynomial([PR.encode(t, 3) for t in table], 3, prm.N)
        for m in (0, 5, 7):
            small = oracle.keyswitch(prm, keys.ksk, oracle.encrypt_big(prm, keys.S, 21, m, PR.encode(m, 3)))
            if m == 7:
                small[0] = small[1] = small[4] = 0                  # a skipped pair and a half pair
            want = oracle.pbs_pairs(prm, bskp, lut, small)
            assert np.array_equal(fast.pbs(lut, small), want)
            if m != 7:
                assert PR.decode(oracle.phase(keys.S, want), 3) == table[m]
