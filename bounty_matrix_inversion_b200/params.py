"""TFHE parameter sets, message encoding and the noise model that selects them.

The reference leaves parameter selection to Concrete's optimizer
(`fhe.Compiler.compile`, /root/reference/matrix_inversion/qfloat_matrix_inversion.py:989-1004),
which picks (LWE dim n, GLWE dim k, polynomial size N, PBS base/level, keyswitch
base/level) from the circuit's largest table-lookup bit-width and the largest
2-norm of the leveled linear combinations between lookups.  `optimize()` below is
the same decision procedure (security curve -> noise of KS / mod-switch / blind
rotation -> cheapest set that keeps the failure probability under the target),
restated for this engine's ciphertext modulus p = 2^64 - 2^32 + 1 and its exact
(noise-free) NTT external product.

Ciphertext layout everywhere: little arrays of uint64 in [0, p);  LWE = mask words
followed by the body.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, replace

import numpy as np

P = 0xFFFFFFFF00000001          # 2^64 - 2^32 + 1
TWO64 = float(2 ** 64)


@dataclass(frozen=True)
class TfheParams:
    name: str
    n: int              # small LWE dimension
    k: int              # GLWE dimension (this engine's kernels: k == 1)
    N: int              # polynomial size
    bsk_bl: int         # PBS decomposition base log
    bsk_l: int          # PBS decomposition levels
    ksk_bl: int         # keyswitch base log
    ksk_l: int          # keyswitch levels
    lwe_sigma: float    # small-key noise std, units of 2^-64
    glwe_sigma: float   # GLWE / big-key noise std, units of 2^-64
    bsk_group: int = 1  # key bits per blind-rotation step: 1 = one GGSW per bit, 2 = pair key (3 GGSWs per 2 bits)

    @property
    def big_dim(self):
        return self.k * self.N

    @property
    def logN(self):
        return self.N.bit_length() - 1

    def bsk_bytes(self):
        ggsws = self.n if self.bsk_group == 1 else 3 * (self.n // 2)
        return ggsws * (self.k + 1) * self.bsk_l * (self.k + 1) * self.N * 8

    def ksk_bytes(self):
        return self.big_dim * self.ksk_l * (self.n + 1) * 8


# ---------------------------------------------------------------- encoding
def delta(width: int) -> int:
    """plaintext scale of a `width`-bit message with one padding bit"""
    return 1 << (63 - width)


def encode(m: int, width: int) -> int:
    """signed integer message -> field element m * 2^(63-width) mod p"""
    return (int(m) * delta(width)) % P


def decode(phase: int, width: int) -> int:
    """phase -> message in [0, 2^(width+1)) (padding bit included)"""
    sh = 63 - width
    return ((int(phase) + (1 << (sh - 1))) >> sh) & ((1 << (width + 1)) - 1)


def decode_signed(phase: int, width: int) -> int:
    m = decode(phase, width)
    return m - (1 << (width + 1)) if m >= (1 << width) else m


def lut_polynomial(table, in_width: int, N: int) -> np.ndarray:
    """Accumulator polynomial for a table of 2^in_width field elements (already
    encoded at the output scale): each entry fills a box of N / 2^in_width
    coefficients, the whole rotated by half a box so rounding is centred."""
    size = 1 << in_width
    assert len(table) == size and N % size == 0 and N // size >= 2, (len(table), size, N)
    box = N // size
    flat = np.repeat(np.asarray([int(t) % P for t in table], dtype=np.uint64), box)
    half = box // 2
    out = np.empty(N, np.uint64)
    out[: N - half] = flat[half:]
    tail = flat[:half]
    out[N - half:] = np.where(tail == 0, np.uint64(0), np.uint64(P) - tail)
    return out


# ------------------------------------------------------------- noise model
# 128-bit security, binary keys: log2(min noise std) = SLOPE * dim + BIAS
# (linear fit of the lattice-estimator curve used by concrete-security-curves).
SEC_SLOPE, SEC_BIAS = -0.026599462343105267, 2.981543184145991
MIN_STD_LOG2 = -62.0


def secure_std(dim: int) -> float:
    """smallest secure noise std for an LWE/GLWE instance of total dimension dim (torus units)"""
    return 2.0 ** max(SEC_SLOPE * dim + SEC_BIAS, MIN_STD_LOG2)


def variance_blind_rotate(p: TfheParams) -> float:
    """one GGSW external product per key bit; with the pair key each step of two bits adds three products, each
    multiplied by a monomial minus one (two coefficients: twice the variance): 3x the noise per bit"""
    B2 = 4.0 ** p.bsk_bl
    vb = (p.glwe_sigma / TWO64) ** 2
    per = p.bsk_l * (p.k + 1) * p.N * (B2 + 2) / 12.0 * vb + (1 + p.k * p.N / 2.0) / (24.0 * B2 ** p.bsk_l)
    return p.n * per * (3.0 if p.bsk_group == 2 else 1.0)


def variance_keyswitch(p: TfheParams) -> float:
    B2 = 4.0 ** p.ksk_bl
    vk = (p.lwe_sigma / TWO64) ** 2
    return p.big_dim * (p.ksk_l * (B2 + 2) / 12.0 * vk + 1.0 / (24.0 * B2 ** p.ksk_l))


def variance_modswitch(p: TfheParams) -> float:
    return (1 + p.n / 2.0) / (12.0 * (2.0 * p.N) ** 2)


def variance_pbs_input(p: TfheParams, nu2: float) -> float:
    """phase noise right before the blind rotation of a lookup whose input is a linear
    combination (squared 2-norm nu2) of earlier lookup outputs"""
    return nu2 * variance_blind_rotate(p) + variance_keyswitch(p) + variance_modswitch(p)


def failure_sigmas(p: TfheParams, width: int, nu2: float) -> float:
    """how many standard deviations fit in half a plaintext step"""
    return (0.5 ** (width + 2)) / math.sqrt(variance_pbs_input(p, nu2))


def cost(p: TfheParams) -> float:
    """integer multiply count of keyswitch + PBS (what both the CPU and the GPU path pay)"""
    ntt = (p.k + 1) * (p.bsk_l + 1) * (p.N / 2) * p.logN
    pointwise = (p.k + 1) ** 2 * p.bsk_l * p.N
    steps = p.n if p.bsk_group == 1 else p.n / 2 * (1 + 4 * pointwise / (ntt + pointwise))     # 3 keys combined per slot
    return steps * (ntt + pointwise) + p.big_dim * p.ksk_l * (p.n + 1)


def optimize(width: int, nu2: float = 1.0, z: float = 6.5, k: int = 1, max_logN: int = 14, bsk_group: int = 1) -> TfheParams:
    """cheapest 128-bit-secure set whose lookups of `width`-bit messages fail with
    probability < erfc(z / sqrt 2); bsk_group = 2 selects among sets the pair blind rotation runs (one
    decomposition level, even n) with its noise"""
    best = None
    for logN in range(max(width + 1, 9), max_logN + 1):
        N = 1 << logN
        gs = secure_std(k * N)
        for n in range(480, 1201, 8):
            ls = secure_std(n)
            for bl_l in ((bl, l) for l in range(1, 5 if bsk_group == 1 else 2) for bl in range(4, 33) if bl * l <= 48):
                base = TfheParams("", n, k, N, bl_l[0], bl_l[1], 1, 1, ls * TWO64, gs * TWO64, bsk_group)
                vbr = nu2 * variance_blind_rotate(base) + variance_modswitch(base)
                budget = (0.5 ** (width + 2) / z) ** 2 - vbr
                if budget <= 0:
                    continue
                for kl in range(1, 9):
                    found = False
                    for kbl in range(min(24, 48 // kl), 0, -1):
                        cand = replace(base, ksk_bl=kbl, ksk_l=kl)
                        if variance_keyswitch(cand) <= budget:
                            c = cost(cand)
                            if best is None or c < best[0]:
                                best = (c, cand)
                            found = True
                            break
                    if found:
                        break
    if best is None:
        raise ValueError(f"no parameter set for width={width} nu2={nu2}")
    p = best[1]
    return replace(p, name=f"opt_w{width}_nu{int(nu2)}_n{p.n}_N{p.N}" + ("_pairs" if bsk_group == 2 else ""))


# ----------------------------------------------------------------- presets
# Insecure toy sets: small enough for the CPU oracle to bootstrap in milliseconds; the
# kernels run exactly the same code on them.  Noise is tiny so lookups decrypt correctly.
TOY_1024 = TfheParams("toy_n48_N1024", 48, 1, 1024, 8, 3, 4, 5, 2.0 ** 24, 2.0 ** 14)
TOY_2048 = TfheParams("toy_n40_N2048", 40, 1, 2048, 12, 2, 4, 5, 2.0 ** 24, 2.0 ** 14)
TOY_4096 = TfheParams("toy_n24_N4096", 24, 1, 4096, 23, 1, 6, 3, 2.0 ** 24, 2.0 ** 10)
TOY_8192 = TfheParams("toy_n16_N8192", 16, 1, 8192, 15, 2, 4, 5, 2.0 ** 24, 2.0 ** 10)
# one decomposition level (what every 128-bit set uses): exercises the l == 1 kernel paths at every polynomial size
TOY_1024_L1 = TfheParams("toy_n40_N1024_l1", 40, 1, 1024, 22, 1, 4, 5, 2.0 ** 24, 2.0 ** 10)
TOY_2048_L1 = TfheParams("toy_n32_N2048_l1", 32, 1, 2048, 22, 1, 4, 5, 2.0 ** 24, 2.0 ** 10)
TOY_8192_L1 = TfheParams("toy_n16_N8192_l1", 16, 1, 8192, 24, 1, 4, 5, 2.0 ** 24, 2.0 ** 10)
TOY_16384_L1 = TfheParams("toy_n10_N16384_l1", 10, 1, 16384, 24, 1, 4, 5, 2.0 ** 24, 2.0 ** 10)


def _secure(name, n, N, bbl, bl, kbl, kl):
    return TfheParams(name, n, 1, N, bbl, bl, kbl, kl, secure_std(n) * TWO64, secure_std(N) * TWO64)


# 128-bit sets produced by optimize() (z = 6.5) for the widths the QFloat circuits need;
# frozen here so keys and benchmarks are reproducible.  tests/test_params.py re-derives them.
SECURE = {}


def for_width(width: int, nu2: float = 1.0, bsk_group: int = 1) -> TfheParams:
    key = (width, int(math.ceil(nu2)), bsk_group)
    if key not in SECURE:
        try:
            SECURE[key] = optimize(width, nu2, bsk_group=bsk_group)
        except ValueError:
            if bsk_group == 1:
                raise
            SECURE[key] = optimize(width, nu2)          # no one-level set for this width: one GGSW per bit
    return SECURE[key]
