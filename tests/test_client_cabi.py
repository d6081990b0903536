"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol
include/bmi_tfhe.h declares; its generator is ChaCha20 (known answers); the keys and ciphertexts its client functions
(keygen / encrypt / phase) produce decrypt, under the oracle's arithmetic, to what the spec says with the configured
noise.  No GPU compute is called."""
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
    rc = native.lib().bmi_keygen_lwe(ctypes.byref(bad), bytes(32), s.ctypes.data_as(ctypes.c_void_p))
    assert rc == -1 and b"invalid" in native.lib().bmi_last_error()


def _chacha20_block(key: bytes, counter: int, nonce: int) -> bytes:
    """pure-Python ChaCha20 block (20 rounds, 64-bit counter and nonce), the published algorithm"""
    import struct
    M = 0xFFFFFFFF
    rot = lambda v, r: ((v << r) & M) | (v >> (32 - r))
    init = list(struct.unpack("<4I", b"expand 32-byte k")) + list(struct.unpack("<8I", key)) + [
        counter & M, counter >> 32, nonce & M, nonce >> 32]
    x = list(init)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & M; x[d] = rot(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & M; x[b] = rot(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & M; x[d] = rot(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & M; x[b] = rot(x[b] ^ x[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return struct.pack("<16I", *[(a + b) & M for a, b in zip(x, init)])


def test_client_generator_is_chacha20(native):
    """known answers: the all-zero key/nonce keystream published with the cipher (draft-agl-tls-chacha20poly1305,
    test vector 1: 76b8e0ada0f13d90...), and random-access reads against a pure-Python ChaCha20"""
    zero = native.rng_words(bytes(32), 0, 0, 8).tobytes()
    assert zero.hex().startswith("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7")
    key = bytes(range(32))
    words = native.rng_words(key, stream=0x0123456789ABCDEF, ctr0=8 * 5 + 3, count=21)
    ref = b"".join(_chacha20_block(key, blk, 0x0123456789ABCDEF) for blk in range(5, 9))
    assert words.tobytes() == ref[3 * 8: 3 * 8 + 21 * 8]
    a, b = native.random_seed(), native.random_seed()
    assert len(a) == 32 and a != b


def _centered(x):
    x = x.astype(np.uint64)
    return np.where(x > np.uint64(PR.P // 2), x.astype(np.float64) - float(PR.P), x.astype(np.float64))


def test_client_keys_are_valid_encryptions_of_the_secret_keys(native, oracle):
    """every row of the bootstrapping / keyswitch key decrypts (under the secret keys) to key bit x gadget value with
    noise of the configured standard deviation; masks are not shared between rows"""
    prm = PR.TfheParams("t", 20, 1, 2048, 12, 2, 5, 3, 2.0 ** 30, 2.0 ** 20)
    ck = native.ClientKeys(prm, seed=42, threads=3)
    assert set(np.unique(ck.s)) <= {0, 1} and set(np.unique(ck.S)) <= {0, 1}
    assert 0.3 < ck.S.mean() < 0.7 and ck.bsk.max() < PR.P
    # keyswitch key: LWE_s(S_i * 2^(64 - j*bl))
    noise = []
    for i in (0, 1, 777, prm.big_dim - 1):
        for j in range(prm.ksk_l):
            ph = int(ck.phase(ck.ksk[i, j][None], small=True)[0])
            want = int(ck.S[i]) << (64 - (j + 1) * prm.ksk_bl)
            noise.append((ph - want) % PR.P)
    noise = _centered(np.array(noise, dtype=np.uint64))
    assert np.abs(noise).max() < 6 * prm.lwe_sigma and np.abs(noise).max() > 0
    # bootstrapping key rows: GLWE_S(s_i * 2^(64 - j*bl) * X^0 on component c)
    errs = []
    for i in (0, 7, prm.n - 1):
        for r in range(2 * prm.bsk_l):
            c, j = divmod(r, prm.bsk_l)
            a, b = ck.bsk[i, r, 0], ck.bsk[i, r, 1]
            ph = (b.astype(object) - oracle.negacyclic_mul(a, ck.S).astype(object)) % PR.P
            msg = np.zeros(prm.N, dtype=object)
            g = int(ck.s[i]) << (64 - (j + 1) * prm.bsk_bl)
            if c == 0:       # gadget on the mask component: phase carries -S * g
                msg = (-(ck.S.astype(object)) * g) % PR.P
            else:
                msg[0] = g
            errs.append(_centered(np.array([(int(x) - int(m)) % PR.P for x, m in zip(ph, msg)], dtype=np.uint64)))
    errs = np.concatenate(errs)
    assert 0.9 * prm.glwe_sigma < errs.std() < 1.1 * prm.glwe_sigma and abs(errs.mean()) < 0.1 * prm.glwe_sigma
    assert len({int(ck.bsk[i, 0, 0, 0]) for i in range(prm.n)}) == prm.n
    # fixed seeds reproduce, entropy seeds do not
    again = native.ClientKeys(prm, seed=42, evaluation_keys=False)
    fresh1, fresh2 = native.ClientKeys(prm, evaluation_keys=False), native.ClientKeys(prm, evaluation_keys=False)
    assert np.array_equal(again.S, ck.S) and not np.array_equal(fresh1.S, fresh2.S)


def test_client_pair_key_bootstraps_on_the_oracle(native, oracle):
    """pair key (two key bits per blind-rotation step): the oracle's pair bootstrap with the client's key decrypts to
    the table entry, which pins the key's content (three GGSWs of the bit products per pair)"""
    prm = PR.TOY_1024_L1
    ck = native.ClientKeys(prm, seed=42, threads=3, pairs=True)
    assert ck.bsk is None and ck.bskp.shape == (prm.n // 2, 3, 2, 2, prm.N)
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


def test_client_encrypt_phase(native, oracle):
    prm = PR.TOY_1024
    ck = native.ClientKeys(prm, seed=7, evaluation_keys=False)
    pts = [PR.encode(m, 3) for m in (0, 1, 7, -3)]
    cts = ck.encrypt(pts, ct_index0=100)
    ph = ck.phase(cts)
    assert [PR.decode_signed(int(p), 3) for p in ph] == [0, 1, 7, -3]
    assert [PR.decode(int(p), 3) for p in ph] == [0, 1, 7, 13]
    assert [int(p) for p in ph] == [oracle.phase(ck.S, c) for c in cts]
    # the same position reproduces (test hook), the running counter never reuses one
    assert np.array_equal(ck.encrypt(pts[:1], ct_index0=100)[0], cts[0])
    a, b = ck.encrypt(pts[:1])[0], ck.encrypt(pts[:1])[0]
    assert not np.array_equal(a[:8], b[:8])
    fresh = native.ClientKeys(prm, evaluation_keys=False)
    try:
        fresh.encrypt(pts, ct_index0=0)
        assert False, "expected ValueError"
    except ValueError:
        pass
    noise = _centered((fresh.phase(fresh.encrypt([0] * 512)) % np.uint64(PR.P)))
    assert 0.8 * prm.glwe_sigma < noise.std() < 1.2 * prm.glwe_sigma


def test_missing_library_is_an_error(native, monkeypatch, tmp_path):
    from bounty_matrix_inversion_b200 import _build
    monkeypatch.setattr(_build, "LIB", str(tmp_path / "nope.so"))
    monkeypatch.setattr(native, "_lib", None)
    try:
        native.lib()
        assert False, "expected NativeError"
    except native.NativeError as e:
        assert "no CPU fallback" in str(e)
