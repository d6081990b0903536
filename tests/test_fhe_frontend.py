"""CPU tests of the tracing front-end and the lowering (no GPU, no encryption): the compiled program's
clear evaluation must equal plain numpy on the same function, and the lookup counts must show fusing."""
import numpy as np
import pytest

from bounty_matrix_inversion_b200 import fhe, params as PR

rng = np.random.default_rng(7)


def compile_(fn, inputset, names=("x", "y"), **options):
    return fhe.Compiler(fn, {n: "encrypted" for n in names}).compile(
        inputset, fhe.Configuration(tfhe_params=PR.TOY_1024, **options))


def pairs(lo, hi, shape, count=60):
    return [(rng.integers(lo, hi, shape), rng.integers(lo, hi, shape)) for _ in range(count)]


def check(fn, inputset, names=("x", "y")):
    c = compile_(fn, inputset, names)
    for args in inputset[:25]:
        want = np.asarray(fn(*args))
        assert np.array_equal(c.simulate(*args), want), (args, c.simulate(*args), want)
    return c


def test_outside_tracing_zeros_are_numpy():
    z, o = fhe.zeros((2, 3)), fhe.ones(4)
    assert isinstance(z, np.ndarray) and z.dtype == np.int64 and not z.any() and o.sum() == 4


def test_leveled_ops_cost_no_lookup():
    c = check(lambda x, y: 3 * x - y + 2 + np.sum(x) - (-y), pairs(-4, 5, (3,)))
    assert c.statistics["pbs"] == 0 and c.statistics["levels"] == 0


def test_univariate_chain_fuses_into_one_lookup():
    c = check(lambda x, y: (np.abs(x - y) // 2) * 3 + 1, pairs(0, 8, (4,)))
    assert c.statistics["pbs"] == 4                     # one per element, not three


def test_comparison_family():
    def fn(x, y):
        out = fhe.zeros((8, 3))
        out[0] = x < y; out[1] = x <= y; out[2] = x > y; out[3] = x >= y
        out[4] = x == y; out[5] = x != y; out[6] = x < 2; out[7] = 1 - (x >= 1)
        return out
    c = check(fn, pairs(-3, 4, (3,)))
    assert c.statistics["keyswitches"] <= 3 * 2          # x-y shared by six tables, x by two


def test_encrypted_product_is_two_lookups_each():
    c = check(lambda x, y: x * y, pairs(-3, 4, (5,)))
    assert c.statistics["pbs"] == 10                     # wide factors: quarter squares


def test_bit_times_small_value_is_one_lookup():
    """three quarters of the reference's lookups are digit * bit products (base_p_arrays.py:197-198, qfloat.py:892):
    packed into one lookup by default, two with Concrete's lowering; same values"""
    fn = lambda x, y: x * (y > 1) + (x - 1) * (1 - (y > 0))
    inputset = [(rng.integers(0, 2, 4), rng.integers(0, 4, 4)) for _ in range(60)]
    auto = check(fn, inputset)
    qs = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"}).compile(
        inputset, fhe.Configuration(tfhe_params=PR.TOY_1024, multiplication="quarter_square"))
    assert auto.statistics["pbs"] == 4 * 4 and qs.statistics["pbs"] == 4 * 6
    for args in inputset[:20]:
        assert np.array_equal(auto.simulate(*args), qs.simulate(*args))
    # one value beyond the range seen on the inputset is still handled (guard band of the packed lookup)
    assert np.array_equal(auto.simulate(np.array([2, 0, 1, 2]), np.array([3, 3, 0, 0])), fn(np.array([2, 0, 1, 2]), np.array([3, 3, 0, 0])))


def test_floor_div_mod_sign_abs():
    check(lambda x, y: (x + y) // 3 + (x + y) % 3 * 10 + np.sign(x - y) * 100 + abs(x), pairs(-5, 6, (2, 2)))


def test_bitwise_between_encrypted_bits():
    def fn(x, y):
        a, b = x > 0, y > 0
        return (a & b) + 2 * (a | b) + 4 * (a ^ b) + 8 * (a & 1)
    check(fn, pairs(-2, 3, (4,)))


def test_indexing_assignment_concatenate_reshape():
    def fn(x, y):
        z = np.concatenate((x[1:], y[:2], fhe.zeros(1)), axis=0)
        m = fhe.zeros((2, 3))
        m[0, :] = z[:3]
        m[1, 1:] = (z[3:5] * 2).reshape(2)
        m[:, 0] = m[:, 0] + 1
        t = m.flatten()
        t[-1] += x[0]
        return np.reshape(t, (3, 2))
    check(fn, pairs(0, 5, (3,)))


def test_carry_propagation_like_base_p_addition():
    """the reference's base_p_addition loop (base_p_arrays.py:100-103) written against this front-end"""
    def fn(x, y):
        out, carry = fhe.zeros(x.size), 0
        for i in range(x.size):
            s = x[-i - 1] + y[-i - 1] + carry
            out[-i - 1] = s % 2
            carry = s // 2
        return out
    c = check(fn, pairs(0, 2, (6,)))
    assert c.statistics["levels"] == 6 and c.statistics["pbs"] == 11      # last carry is dead code


def test_common_lookups_are_shared_and_dead_ones_dropped():
    def fn(x, y):
        a = (x + y) // 2
        b = (x + y) // 2          # same table on the same input
        _dead = (x - y) % 3
        return a + b
    c = check(fn, pairs(0, 6, (2,)))
    assert c.statistics["traced_lookups"] == 4 and c.statistics["pbs"] == 2


def test_fhe_univariate_and_where():
    f = fhe.univariate(lambda v: np.where(v & 1 == 0, 0, v >> 1))
    check(lambda x, y: f(x * 2 + (y > 2)), pairs(0, 6, (3,)))


def test_constant_folding_of_zero_arrays():
    c = check(lambda x, y: fhe.zeros(3) * x + fhe.ones(3) * y, pairs(0, 4, (3,)))
    assert c.statistics["pbs"] == 0


def test_errors():
    c = fhe.Compiler(lambda x: x, {"x": "encrypted"})
    with pytest.raises(ValueError):
        c.compile([])
    with pytest.raises(NotImplementedError):
        fhe.Compiler(lambda x: x, {"x": "clear"})
    with pytest.raises(TypeError):
        compile_(lambda x, y: x if x[0] else y, pairs(0, 2, (2,)))
    with pytest.raises(TypeError):
        compile_(lambda x, y: x / y, pairs(1, 3, (2,)))
    circ = compile_(lambda x, y: x + y, pairs(0, 2, (2,)))
    with pytest.raises(ValueError):
        circ.simulate(np.zeros(3), np.zeros(2))


def test_width_and_norm_drive_parameter_choice():
    extremes = [(np.zeros(8, np.int64), np.zeros(8, np.int64)), (np.ones(8, np.int64), np.ones(8, np.int64))]
    c = compile_(lambda x, y: (np.sum(x) + np.sum(y)) // 4, pairs(0, 2, (8,)) + extremes, split_wide=False)
    assert c.program.width == 5 and c.program.nu2 == 16            # values 0..16 need 17 table entries
    p = PR.optimize(3, 4)
    assert p.k == 1 and PR.failure_sigmas(p, 3, 4) >= 6.5 and p.N >= 1024
    assert PR.failure_sigmas(p, 4, 4) < 6.5 or PR.cost(PR.optimize(4, 4)) >= PR.cost(p)


def test_program_roundtrip(tmp_path):
    c = compile_(lambda x, y: (x * y + 3) // 2, pairs(0, 4, (3,)))
    path = str(tmp_path / "p.npz")
    c.program.save(path)
    from bounty_matrix_inversion_b200.fhe.program import Program
    q = Program.load(path)
    x = rng.integers(0, 4, (5, 6))
    assert np.array_equal(q.evaluate_clear(x), c.program.evaluate_clear(x))
    assert q.stats == c.program.stats


def test_slack_bits_widen_the_message_space():
    fn = lambda x, y: (x + y) // 2
    base = compile_(fn, pairs(0, 4, (2,)))
    wide = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"}).compile(
        pairs(0, 4, (2,)), fhe.Configuration(tfhe_params=PR.TOY_1024, slack_bits=1))
    assert wide.program.width == base.program.width + 1
    x, y = np.array([3, 1]), np.array([3, 2])
    assert np.array_equal(wide.simulate(x, y), fn(x, y))


def test_non_strict_evaluation_flags_out_of_range_lanes():
    c = compile_(lambda x, y: (x + y) // 2, pairs(0, 4, (2,)))        # sums 0..6 seen -> 3-bit window
    inside, outside = np.array([[3, 3, 3, 3]]), np.array([[3, 3, 40, 40]])
    out, bad = c.program.evaluate_clear(np.concatenate([inside, outside]), strict=False)
    assert list(bad) == [False, True] and np.array_equal(out[0], [3, 3])
    with pytest.raises(OverflowError):
        c.program.evaluate_clear(outside[0])


def test_functions_of_the_same_input_fuse_across_a_product():
    """(abs(c) // 2) * sign(c) and c * (c >= 0) are univariate in c: one lookup each (qfloat.py:619, 663)"""
    def fn(x, y):
        c = x - y
        return np.concatenate(((np.abs(c) // 2) * np.sign(c), c * (c >= 0), (c > 0) & (c < 3)), axis=0)
    c = check(fn, pairs(-4, 5, (3,)))
    assert c.statistics["pbs"] == 9 and c.statistics["levels"] == 1


# ---------------------------------------------------------------- wide-lookup splitting (program.lower docstring)
def _wide_circuit(x, y):
    s = np.sum(x) + np.sum(y)                    # 0..16: one value more than 4 bits hold
    out = fhe.zeros(3)
    out[0], out[1], out[2] = s // 3, fhe.univariate(lambda v: (v * v) % 7 - 3)(s), s % 2
    return out + x[:3] % 2


def test_wide_lookup_is_split_into_three_narrow_ones():
    extremes = [(np.zeros(8, np.int64), np.zeros(8, np.int64)), (np.ones(8, np.int64), np.ones(8, np.int64))]
    inputset = pairs(0, 2, (8,)) + extremes
    wide = compile_(_wide_circuit, inputset, split_wide=False)
    split = compile_(_wide_circuit, inputset, split_wide=True, split_guard=0)
    auto = compile_(_wide_circuit, inputset)
    assert wide.program.width == 5 and split.program.width == 4
    assert auto.program.width == 5 and auto.statistics["split_lookups"] == 0        # 3 of 6 lookups are wide: not worth it
    st = split.statistics
    # three wide lookups of one source: ONE shared sign bootstrap, one negacyclic + one cyclic half each
    assert st["split_lookups"] == 3 and st["pbs"] == wide.statistics["pbs"] + 3 + 1 and st["levels"] == 2
    assert split.program.table_half.sum() >= 2 and sum(int(l.full.sum()) for l in split.program.levels) == 1
    for args in inputset:
        want = _wide_circuit(*args)
        assert np.array_equal(split.simulate(*args), want) and np.array_equal(wide.simulate(*args), want)
    # every value of the 17-entry range, both sides of the padding bit
    for total in range(17):
        x = np.array([1] * min(total, 8) + [0] * (8 - min(total, 8)))
        y = np.array([1] * max(total - 8, 0) + [0] * (8 - max(total - 8, 0)))
        assert np.array_equal(split.simulate(x, y), _wide_circuit(x, y)), total


def test_split_program_round_trips_through_npz(tmp_path):
    extremes = [(np.zeros(8, np.int64), np.zeros(8, np.int64)), (np.ones(8, np.int64), np.ones(8, np.int64))]
    c = compile_(_wide_circuit, pairs(0, 2, (8,)) + extremes, split_wide=True)
    path = str(tmp_path / "p.npz")
    c.program.save(path)
    from bounty_matrix_inversion_b200.fhe.program import Program
    q = Program.load(path)
    assert np.array_equal(q.table_half, c.program.table_half) and np.array_equal(q.tables, c.program.tables)
    assert all(np.array_equal(a.full, b.full) for a, b in zip(q.levels, c.program.levels))
    x = np.ones((5, 16), np.int64)
    assert np.array_equal(q.evaluate_clear(x), c.program.evaluate_clear(x))
    # the accumulators of half-unit tables carry odd multiples of half the plaintext scale
    polys = q.lut_polynomials(1024)
    half_scale = PR.delta(q.width) // 2
    odd = [(int(v) if int(v) < PR.P // 2 else int(v) - PR.P) // half_scale % 2 for v in polys[q.table_half].reshape(-1)[::32]]
    assert any(odd) or not (q.tables[q.table_half] % 2).any()


def test_split_overflow_is_still_detected():
    extremes = [(np.zeros(8, np.int64), np.zeros(8, np.int64)), (np.ones(8, np.int64), np.ones(8, np.int64))]
    c = compile_(_wide_circuit, pairs(0, 2, (8,)) + extremes, split_wide=True)
    # the split lookup itself accepts the whole torus; a narrow one (x % 2, inputs 0..1) still reports leaving its range
    with pytest.raises(OverflowError):
        c.program.evaluate_clear(np.array([12] + [0] * 15, np.int64))


def test_blind_rotation_choice_follows_configuration_and_parameter_set():
    inputset = pairs(0, 4, (2,))
    fn = lambda x, y: (x + y) // 2
    comp = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"})
    assert comp.compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024_L1)).params.bsk_group == 2
    assert comp.compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024)).params.bsk_group == 1          # three levels
    assert comp.compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024_L1, blind_rotation="single")).params.bsk_group == 1
    with pytest.raises(ValueError):
        fhe.Configuration(blind_rotation="triples")


@pytest.mark.parametrize("seed", range(6))
def test_wide_split_random_tables(seed):
    """random tables on sums one bit wider than everything else: the split lowering (sign through the padding bit,
    half-integer negacyclic / cyclic halves) must reproduce every table entry, odd differences included"""
    r = np.random.default_rng(100 + seed)
    terms = int(r.integers(17, 31))                       # sum of `terms` bits: 18..31 distinct values -> 5 bits
    tables = [r.integers(-8, 8, terms + 1) for _ in range(3)]
    nx = terms // 2

    def fn(x, y):
        s = np.sum(x) + np.sum(y)
        out = fhe.zeros(4)
        for i, t in enumerate(tables):
            out[i] = fhe.univariate(lambda v, t=t: t[np.clip(v, 0, len(t) - 1)])(s)
        out[3] = (x[0] + y[0]) % 2                         # a narrow lookup next to the wide ones
        return out + x[:4]

    shape_x, shape_y = (nx,), (terms - nx,)
    inputset = [(r.integers(0, 2, shape_x), r.integers(0, 2, shape_y)) for _ in range(60)]
    inputset += [(np.zeros(shape_x, np.int64), np.zeros(shape_y, np.int64)), (np.ones(shape_x, np.int64), np.ones(shape_y, np.int64))]
    comp = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"})
    split = comp.compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024, split_wide=True, split_guard=0))
    wide = comp.compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024, split_wide=False))
    assert wide.program.width == 5 and split.program.width == 4 and split.statistics["split_lookups"] >= 1
    for total in range(terms + 1):                          # every reachable sum
        bits = np.array([1] * total + [0] * (terms - total))
        x, y = bits[:nx], bits[nx:]
        want = fn(x, y)
        assert np.array_equal(split.simulate(x, y), want), total
        assert np.array_equal(wide.simulate(x, y), want), total


def test_clear_evaluation_is_modular_like_the_ciphertexts():
    """table outputs are stored reduced mod 2^(W+1); sums of them that wrap are still right at the next lookup"""
    r = np.random.default_rng(5)

    def fn(x, y):
        a = fhe.univariate(lambda v: v + 16)(x[0])          # 16..19: outside a 3-bit message space
        b = fhe.univariate(lambda v: v % 2 - 16)(y[0])
        return (a + b + x[0]) // 2                          # the sum is small again: 0..7

    inputset = [(r.integers(0, 4, 1), r.integers(0, 4, 1)) for _ in range(60)]
    c = fhe.Compiler(fn, {"x": "encrypted", "y": "encrypted"}).compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024))
    assert c.program.width == 3 and np.abs(c.program.tables).max() <= 1 << c.program.width
    for x, y in inputset:
        assert np.array_equal(c.simulate(x, y), fn(x, y))


# ---------------------------------------------------------------- borrow-chain collapse (program.lower docstring)
def _subtract_digits(a, b):
    """schoolbook base-2 subtraction a - b, least significant digit last; returns digits and the final borrow"""
    out = fhe.zeros(a.size + 1)
    borrow = 0
    d = a - b
    for i in range(a.size):
        t = d[-i - 1] - borrow
        borrow = t < 0
        out[-i - 2] = t + 2 * borrow
    out[-1] = borrow
    return out


def test_borrow_chain_resolves_several_digits_per_level():
    n = 8
    r = np.random.default_rng(21)
    inputset = [(r.integers(0, 2, n), r.integers(0, 2, n)) for _ in range(80)]
    comp = fhe.Compiler(_subtract_digits, {"a": "encrypted", "b": "encrypted"})
    narrow = comp.compile(inputset, fhe.Configuration(tfhe_params=PR.TOY_1024))
    assert narrow.program.width == 2 and narrow.statistics["collapsed_borrows"] == 0      # no room in a 2-bit message space
    cfg = lambda **kw: fhe.Configuration(tfhe_params=PR.TOY_1024, slack_bits=2, **kw)      # 4 bits, like the reference's circuits
    plain, two, three = (comp.compile(inputset, cfg(collapse_borrows=c)) for c in (False, 2, True))
    # per digit the tracer emits [t < 0] (the next digit's borrow) and 2 [t < 0] (folded into this digit): both collapse
    assert [c.statistics["levels"] for c in (plain, two, three)] == [n, n // 2, (n + 2) // 3]
    assert [c.statistics["collapsed_borrows"] for c in (plain, two, three)] == [0, n, 10]
    assert all(c.statistics["pbs"] == 2 * n and c.program.width == 4 for c in (plain, two, three))
    # all 2^16 digit patterns, vectorised: identical to the uncollapsed program, and to the function itself on a sample
    grid = np.array([[(v >> k) & 1 for k in range(2 * n)] for v in range(1 << (2 * n))], dtype=np.int64)
    want = np.stack([np.asarray(_subtract_digits(row[:n], row[n:])) for row in grid[:: 257]])
    base = plain.program.evaluate_clear(grid)
    for c in (two, three):
        assert np.array_equal(c.program.evaluate_clear(grid[:: 257]), want)
        assert np.array_equal(c.program.evaluate_clear(grid), base)


@pytest.mark.parametrize("n", [5, 9, 24, 31])
def test_borrow_chain_by_parallel_prefix(n):
    """collapse_borrows="prefix": block signs sign(4 x2 + 2 x1 + x0) compose lexicographically, so every borrow of an
    n-digit subtraction is out after about 2 log3(n) levels (lookups + a third); same digits as the plain chain"""
    r = np.random.default_rng(30 + n)
    inputset = [(r.integers(0, 2, n), r.integers(0, 2, n)) for _ in range(80)]
    comp = fhe.Compiler(_subtract_digits, {"a": "encrypted", "b": "encrypted"})
    cfg = lambda **kw: fhe.Configuration(tfhe_params=PR.TOY_1024, slack_bits=2, **kw)
    plain, chain3, prefix = (comp.compile(inputset, cfg(collapse_borrows=c)) for c in (False, True, "prefix"))
    depth = {5: 2, 9: 3, 24: 5, 31: 5}[n]
    assert prefix.statistics["levels"] == depth <= chain3.statistics["levels"] == (n + 2) // 3
    assert plain.statistics["pbs"] <= prefix.statistics["pbs"] <= 1.5 * plain.statistics["pbs"]
    x = r.integers(0, 2, (20000, 2 * n))
    x[0], x[1] = 0, 1                                           # all borrows propagate / none
    x[2, :n], x[2, n:] = 0, 1                                   # 0 - 11..1: a borrow out of every digit
    x[3, :n], x[3, n:] = 1, 0
    x[4] = 0; x[4, 2 * n - 1] = 1                               # 0 - 1: one borrow rippling through every block
    base = plain.program.evaluate_clear(x)
    assert np.array_equal(prefix.program.evaluate_clear(x), base)
    assert np.array_equal(chain3.program.evaluate_clear(x), base)
    want = np.stack([np.asarray(_subtract_digits(row[:n], row[n:])) for row in x[:200]])
    assert np.array_equal(base[:200], want)


def test_borrow_collapse_leaves_other_threshold_chains_alone():
    """[A + b < 0], [A - 2 b < 0] and [A - f(S) < 0] with f not a sign test are not the borrow pattern"""
    r = np.random.default_rng(22)

    def fn(x, y):
        b = x[0] - y[0] < 0
        out = fhe.zeros(3)
        out[0] = (x[1] - y[1] + b) < 0
        out[1] = (x[2] - y[2] - 2 * b) < 0
        out[2] = (x[3] - y[3] - (x[0] - y[0] < 1)) < 0
        return out

    inputset = [(r.integers(0, 3, 4), r.integers(0, 3, 4)) for _ in range(80)]
    c = check(fn, inputset)
    assert c.statistics["collapsed_borrows"] == 0 and c.statistics["levels"] == 2


def _subtract_or_keep(a, b, e):
    """one step of a restoring division: a - b if it does not underflow (and no extra digit e is set), else a"""
    n = a.size
    borrow = 0
    d = a - b
    diff = fhe.zeros(n)
    for i in range(n):
        t = d[-i - 1] - borrow
        borrow = t < 0
        diff[-i - 1] = t + 2 * borrow
    lt = borrow | (np.sum(e) > 0)
    return diff * (1 - lt) + a * lt


@pytest.mark.parametrize("mode, n", [(True, 8), ("prefix", 9), ("prefix", 8)])
def test_overflow_test_joins_the_borrow_chain(mode, n):
    """(a chain of three digits per level has room for the extra digit only where its last group is not full)"""
    r = np.random.default_rng(3)
    sample = lambda: (r.integers(0, 2, n), r.integers(0, 2, n), r.integers(0, 2, 2) * r.integers(0, 2, 2))
    inputset = [sample() for _ in range(200)]
    comp = fhe.Compiler(_subtract_or_keep, {"a": "encrypted", "b": "encrypted", "e": "encrypted"})
    cfg = lambda **kw: fhe.Configuration(tfhe_params=PR.TOY_1024, slack_bits=1, **kw)
    plain = comp.compile(inputset, cfg(collapse_borrows=False))
    fast = comp.compile(inputset, cfg(collapse_borrows=mode))
    assert plain.program.width == fast.program.width == 4
    assert fast.statistics["bitwise_folded"] >= 1 and fast.statistics["levels"] <= 5 < plain.statistics["levels"]
    x = np.stack([np.concatenate(sample()) for _ in range(5000)])
    got = fast.program.evaluate_clear(x)
    assert np.array_equal(got, plain.program.evaluate_clear(x))
    for row, out in zip(x[:300], got[:300]):
        assert np.array_equal(out, _subtract_or_keep(row[:n], row[n:2 * n], row[2 * n:]))
