"""GPU parity: every kernel behind the C ABI against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from bounty_matrix_inversion_b200 import params as PR

pytestmark = pytest.mark.gpu
P = PR.P
TOYS = [PR.TOY_1024, PR.TOY_2048, PR.TOY_4096, PR.TOY_8192, PR.TOY_1024_L1, PR.TOY_2048_L1, PR.TOY_8192_L1, PR.TOY_16384_L1]


def rand_field(rng, shape):
    return (rng.integers(0, 2 ** 63, size=shape, dtype=np.uint64) * np.uint64(2)
            + rng.integers(0, 2, size=shape, dtype=np.uint64)) % np.uint64(P)


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64) if a.dtype == np.uint64 else np.ascontiguousarray(a)).cuda()


def host_u64(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.fixture(scope="module", params=TOYS, ids=lambda p: p.name)
def setup(request, native, oracle):
    prm = request.param
    keys = native.ClientKeys(prm, seed=2024)
    eng = native.Engine(prm, 0)
    eng.load_keys(keys.bsk, keys.ksk)
    yield prm, keys, eng
    eng.close()


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["latency_build", "throughput_build", "split_4cta"])
def test_polymul_matches_oracle(setup, oracle, mode):
    prm, _, eng = setup
    if prm.N > 8192 and mode != 3:
        pytest.skip("N = 16384 exists only on the split kernels")
    eng.set_pbs_mode(mode)
    rng = np.random.default_rng(prm.N)
    a, b = rand_field(rng, (3, prm.N)), rand_field(rng, (3, prm.N))
    a[2] = 0; a[2, 1] = 1; b[2] = 0; b[2, prm.N - 1] = 1          # X * X^(N-1) = -1
    got = eng.polymul_host(a, b)
    for i in range(3):
        assert np.array_equal(got[i], oracle.negacyclic_mul(a[i], b[i]))
    assert got[2, 0] == P - 1 and not got[2, 1:].any()
    eng.set_pbs_mode(0)


def test_keyswitch_matches_oracle(setup, oracle):
    prm, keys, eng = setup
    rng = np.random.default_rng(5)
    count = 11                                                     # ragged against the 8-ciphertext tile
    big = rand_field(rng, (count, prm.big_dim + 1))
    big[0] = 0
    big[1] = P - 1
    import torch
    small = torch.zeros((count, prm.n + 1), dtype=torch.int64, device="cuda")
    eng.keyswitch(dev(big), small, count)
    got = host_u64(small)
    for i in range(count):
        assert np.array_equal(got[i], oracle.keyswitch(prm, keys.ksk, big[i])), i


@pytest.mark.parametrize("mode", [1, 2, 3], ids=["latency_build", "throughput_build", "split_8cta"])
def test_pbs_matches_oracle_bit_exact(setup, oracle, mode):
    prm, keys, eng = setup
    if mode == 3 and prm.bsk_l != 1:
        pytest.skip("the 8-CTA kernel is built for one decomposition level")
    if prm.N > 8192 and mode != 3:
        pytest.skip("N = 16384 exists only on the split kernels")
    eng.set_pbs_mode(mode)
    rng = np.random.default_rng(9)
    w = 3
    tables = [[(3 * m + 1) % 16 for m in range(8)], [m * m % 16 for m in range(8)]]
    luts = np.stack([PR.lut_polynomial([PR.encode(t, 4) for t in tb], w, prm.N) for tb in tables])
    eng.load_luts(luts)
    msgs = [0, 1, 5, 7, 2]
    cts = keys.encrypt([PR.encode(m, w) for m in msgs])
    small = np.stack([oracle.keyswitch(prm, keys.ksk, c) for c in cts])
    small[4] = rand_field(rng, prm.n + 1)                           # arbitrary words, not a real encryption
    small[4, 3] = 0                                                 # a mask word that switches to 0 is skipped
    lut_idx = np.array([0, 1, 0, 1, 1], np.int32)
    import torch
    out = torch.zeros((len(msgs), prm.big_dim + 1), dtype=torch.int64, device="cuda")
    idx = torch.arange(len(msgs), dtype=torch.int32, device="cuda")
    eng.pbs(dev(small), idx, dev(lut_idx), idx, out, len(msgs))
    got = host_u64(out)
    for i in range(len(msgs)):
        want = oracle.pbs(prm, keys.bsk, luts[lut_idx[i]], small[i])
        assert np.array_equal(got[i], want), i
    for i in range(4):
        assert PR.decode(int(keys.phase(got[i])[0]), 4) == tables[lut_idx[i]][msgs[i]]
    eng.set_pbs_mode(0)


def test_pbs_tma_staged_rows_match_oracle(setup, oracle):
    """the TMA-staged variant of the bootstrap (bmi_ctx_set_tma_stage) is bit-identical too"""
    prm, keys, eng = setup
    if prm.bsk_l != 1:
        pytest.skip("row staging is built for one decomposition level")
    luts = np.stack([PR.lut_polynomial([PR.encode(t, 3) for t in (3, 1, 4, 1, 5, 9, 2, 6)], 3, prm.N)])
    eng.load_luts(luts)
    cts = keys.encrypt([PR.encode(m, 3) for m in range(5)])
    small = np.stack([oracle.keyswitch(prm, keys.ksk, c) for c in cts])
    import torch
    idx = torch.arange(5, dtype=torch.int32, device="cuda")
    lut = torch.zeros(5, dtype=torch.int32, device="cuda")
    eng.set_tma_stage(True)
    try:
        for mode in (1, 2):
            eng.set_pbs_mode(mode)
            out = torch.zeros((5, prm.big_dim + 1), dtype=torch.int64, device="cuda")
            eng.pbs(dev(small), idx, lut, idx, out, 5)
            got = host_u64(out)
            for i in range(5):
                assert np.array_equal(got[i], oracle.pbs(prm, keys.bsk, luts[0], small[i])), (mode, i)
    finally:
        eng.set_tma_stage(False)
        eng.set_pbs_mode(0)


def test_pbs_batch_lanes(setup, oracle):
    """batch > 1: job q lane b reads row job_in[q]*batch+b and writes row job_out[q]*batch+b"""
    prm, keys, eng = setup
    w, batch = 2, 3
    luts = np.stack([PR.lut_polynomial([PR.encode(t, 2) for t in (1, 3, 0, 2)], w, prm.N)])
    eng.load_luts(luts)
    msgs = np.array([[0, 1, 2], [3, 2, 1]])                          # [row][lane]
    cts = keys.encrypt([PR.encode(int(m), w) for m in msgs.reshape(-1)])
    import torch
    small = torch.zeros((6, prm.n + 1), dtype=torch.int64, device="cuda")
    eng.keyswitch(dev(cts), small, 6)
    out = torch.zeros((3 * batch, prm.big_dim + 1), dtype=torch.int64, device="cuda")
    job_in = torch.tensor([1, 0], dtype=torch.int32, device="cuda")
    job_out = torch.tensor([0, 2], dtype=torch.int32, device="cuda")
    job_lut = torch.zeros(2, dtype=torch.int32, device="cuda")
    eng.pbs(small, job_in, job_lut, job_out, out, 2, batch)
    got = host_u64(out).reshape(3, batch, -1)
    table = (1, 3, 0, 2)
    for q, (ji, jo) in enumerate([(1, 0), (0, 2)]):
        for b in range(batch):
            assert PR.decode(int(keys.phase(got[jo, b])[0]), 2) == table[msgs[ji, b]]
    assert not got[1].any()                                          # untouched row


def test_lincomb_matches_oracle(setup, oracle):
    prm, keys, eng = setup
    rng = np.random.default_rng(3)
    vals = rand_field(rng, (5, prm.big_dim + 1))
    rows = [([0, 1, 2], [1, -1, 3], 17), ([4], [-2], 0), ([], [], PR.encode(3, 3)), ([3, 3, 1], [1, 1, -1], P - 1)]
    row_ptr = np.cumsum([0] + [len(r[0]) for r in rows]).astype(np.int32)
    idx = np.array(sum((r[0] for r in rows), []), np.int32)
    coef = np.array([c % P for r in rows for c in r[1]], np.uint64)
    konst = np.array([r[2] for r in rows], np.uint64)
    import torch
    out = torch.zeros((len(rows), prm.big_dim + 1), dtype=torch.int64, device="cuda")
    eng.lincomb(dev(vals), dev(row_ptr), dev(idx), dev(coef), dev(konst), out, len(rows))
    got = host_u64(out)
    for j, (ii, cc, kk) in enumerate(rows):
        assert np.array_equal(got[j], oracle.lincomb(vals, ii, cc, kk)), j


def test_host_buffer_path_end_to_end(setup, oracle):
    """encrypt -> bmi_ks_pbs_host -> decrypt == table[m]; and equal to the oracle's ciphertexts"""
    prm, keys, eng = setup
    w = 3
    table = [(7 - m) % 8 for m in range(8)]
    luts = np.stack([PR.lut_polynomial([PR.encode(t, 3) for t in table], w, prm.N)])
    eng.load_luts(luts)
    msgs = list(range(8)) + [3]
    cts = keys.encrypt([PR.encode(m, w) for m in msgs], ct_index0=50)
    before = eng.launch_count
    got = eng.ks_pbs_host(cts, np.zeros(len(msgs), np.int32))
    assert eng.launch_count >= before + 2
    assert [PR.decode(int(p), 3) for p in keys.phase(got)] == [table[m] for m in msgs]
    fast = oracle.Fast(prm, keys.bsk, keys.ksk)
    want = fast.batch(luts, np.zeros(len(msgs), np.int32), cts, with_ks=True, threads=4)
    assert np.array_equal(got, want)


def test_errors(setup, native):
    prm, keys, eng = setup
    with pytest.raises(native.NativeError):
        eng.ks_pbs_host(keys.encrypt([0]), np.array([99], np.int32))   # LUT index out of range
    with pytest.raises(native.NativeError):
        native.Engine(PR.TfheParams("bad", 10, 2, 1024, 8, 3, 4, 5, 1.0, 1.0), 0)   # k != 1


def test_keyswitch_batch_beyond_65535_tiles(native, oracle):
    """more ciphertext tiles than gridDim.y holds (8 x 65535 rows): the tiles run along gridDim.x.  The batch is 13
    distinct random ciphertexts repeated, so every output row must equal the oracle's keyswitch of its source row."""
    import torch
    prm = PR.TOY_1024
    keys = native.ClientKeys(prm, seed=3, evaluation_keys=True)
    eng = native.Engine(prm, 0)
    eng.load_keys(keys.bsk, keys.ksk)
    rng = np.random.default_rng(17)
    base = rand_field(rng, (13, prm.big_dim + 1))
    count = 8 * 65535 + 9
    big = dev(base).repeat((count + 12) // 13, 1)[:count].contiguous()
    small = torch.zeros((count, prm.n + 1), dtype=torch.int64, device="cuda")
    eng.keyswitch(big, small, count)
    torch.cuda.synchronize()
    want = np.stack([oracle.keyswitch(prm, keys.ksk, base[i]) for i in range(13)])
    got = small.view(-1, prm.n + 1)
    want_dev = dev(want)
    idx = torch.arange(count, device="cuda") % 13
    assert bool(torch.equal(got, want_dev[idx]))
    eng.close()


def test_scatter_rows(native):
    import torch
    prm = PR.TOY_1024
    eng = native.Engine(prm, 0)
    W = prm.big_dim + 1
    src = torch.arange(5 * 3 * W, dtype=torch.int64, device="cuda").view(5, 3, W)
    dst = torch.zeros((9, 3, W), dtype=torch.int64, device="cuda")
    rows = torch.tensor([7, 0, 3, 8, 1], dtype=torch.int32, device="cuda")
    eng.scatter_rows(src, rows, dst, 5, batch=3)
    torch.cuda.synchronize()
    assert bool(torch.equal(dst[rows.long()], src)) and int(dst[2].abs().sum()) == 0
    eng.close()
