"""Parameter selection, message encoding and accumulator layout (host logic, no GPU).

The reference delegates these to Concrete's optimizer and runtime
(/root/reference/matrix_inversion/qfloat_matrix_inversion.py:989-1004); the properties
checked here are the ones its circuits rely on.
"""
import math

import numpy as np
import pytest

from bounty_matrix_inversion_b200 import params as PR


@pytest.mark.parametrize("width", [1, 3, 5, 8])
def test_encode_decode_round_trip_with_noise(width):
    rng = np.random.default_rng(width)
    # negative messages wrap modulo p, not 2^64: the 2^32 - 1 difference eats that much of the noise margin
    half_step = PR.delta(width) // 2 - (1 << 32)
    for m in range(-(1 << width), 1 << width):
        for noise in (0, half_step - 1, -(half_step - 1), int(rng.integers(-half_step + 1, half_step))):
            phase = (PR.encode(m, width) + noise) % PR.P
            assert PR.decode_signed(phase, width) == m or (m == -(1 << width) and PR.decode(phase, width) == 1 << width)


def test_lut_polynomial_is_half_box_rotated_and_negacyclic():
    N, w = 1024, 3
    table = [PR.encode(t * t % 8, w) for t in range(1 << w)]
    acc = PR.lut_polynomial(table, w, N)
    box = N >> w
    # coefficient j of X^-r * acc, r = round(phase * 2N / p) of message t with padding bit clear, is table[t]:
    # check the window each message's rotation can land on
    for t in range(1 << w):
        for r in (t * box - box // 2, t * box, t * box + box // 2 - 1):
            c = acc[r] if r >= 0 else (PR.P - int(acc[N + r])) % PR.P
            assert int(c) == table[t], (t, r)


def test_lut_polynomial_rejects_tables_wider_than_the_ring():
    with pytest.raises(AssertionError):
        PR.lut_polynomial([0] * 1024, 10, 1024)


def test_security_curve_is_monotone_and_clamped():
    stds = [PR.secure_std(d) for d in range(400, 20000, 200)]
    assert all(a >= b for a, b in zip(stds, stds[1:]))
    assert stds[-1] == 2.0 ** PR.MIN_STD_LOG2
    assert PR.secure_std(2048) == 2.0 ** (PR.SEC_SLOPE * 2048 + PR.SEC_BIAS)


@pytest.mark.parametrize("width, n, N", [(4, 752, 2048), (5, 840, 4096)])
def test_optimizer_reproduces_documented_sets(width, n, N):
    """DESIGN.md's parameter table: the sets every measurement in profiles/ was taken on."""
    p = PR.optimize(width, 1.0)
    assert (p.n, p.N, p.k, p.bsk_l) == (n, N, 1, 1)
    assert PR.failure_sigmas(p, width, 1.0) >= 6.5
    # both key noises sit exactly on the security curve
    assert math.isclose(p.lwe_sigma, PR.secure_std(p.n) * PR.TWO64)
    assert math.isclose(p.glwe_sigma, PR.secure_std(p.N) * PR.TWO64)
    # and the set is locally minimal: a smaller LWE dimension at the same ring fails the noise bound
    smaller = PR.TfheParams("", p.n - 64, 1, p.N, p.bsk_bl, p.bsk_l, p.ksk_bl, p.ksk_l,
                            PR.secure_std(p.n - 64) * PR.TWO64, p.glwe_sigma)
    assert PR.failure_sigmas(smaller, width, 1.0) < 6.5


def test_noise_grows_with_the_leveled_norm():
    p = PR.optimize(4, 1.0)
    assert PR.failure_sigmas(p, 4, 64.0) < PR.failure_sigmas(p, 4, 1.0)
    q = PR.optimize(4, 64.0)
    assert PR.failure_sigmas(q, 4, 64.0) >= 6.5 and PR.cost(q) >= PR.cost(p)


def test_for_width_caches_by_width_and_norm():
    a = PR.for_width(4, 1.0)
    assert PR.for_width(4, 1.0) is a


def test_key_sizes():
    p = PR.TOY_2048_L1
    assert p.bsk_bytes() == p.n * 2 * p.bsk_l * 2 * p.N * 8
    assert p.ksk_bytes() == p.N * p.ksk_l * (p.n + 1) * 8


def test_pair_blind_rotation_noise_and_sets():
    """pair key: three GGSW products per two key bits, each times a monomial minus one -> 3x the blind-rotation noise;
    the optimizer then only picks one-level sets with an even LWE dimension"""
    p1 = PR.optimize(4, 400.0)
    p2 = PR.optimize(4, 400.0, bsk_group=2)
    assert p2.bsk_group == 2 and p2.bsk_l == 1 and p2.n % 2 == 0 and p2.name.endswith("_pairs")
    assert PR.failure_sigmas(p2, 4, 400.0) >= 6.5
    import dataclasses
    assert math.isclose(PR.variance_blind_rotate(dataclasses.replace(p1, bsk_group=2)), 3 * PR.variance_blind_rotate(p1))
    assert p2.bsk_bytes() == 3 * (p2.n // 2) * 4 * p2.N * 8
    assert PR.cost(p2) < PR.cost(dataclasses.replace(p2, bsk_group=1))
    assert PR.for_width(4, 400.0, bsk_group=2) is PR.for_width(4, 400.0, bsk_group=2)
