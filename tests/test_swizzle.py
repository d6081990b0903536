"""The shared-memory XOR swizzles compiled into csrc/ntt.cuh are bank-conflict free for every pass of every
transform configuration the kernels instantiate (brute force over all half-warps, same code that found them)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from find_swizzle import conflict_free  # noqa: E402


def cols_from_source(L, E):
    """mirror of swz_cols<L, E>() in csrc/ntt.cuh, parsed from the source so the test follows the kernels"""
    src = open(os.path.join(ROOT, "bounty_matrix_inversion_b200", "csrc", "ntt.cuh")).read()
    e4 = int(re.search(r"if \(E == 4\) return (0x[0-9A-Fa-f]+);", src).group(1), 16)
    m2 = re.search(r"if \(E == 2\) return \(L & 1\) \? (0x[0-9A-Fa-f]+) : (0x[0-9A-Fa-f]+);", src)
    m3 = re.search(r"return L == 10 \? (0x[0-9A-Fa-f]+) : L == 11 \? (0x[0-9A-Fa-f]+) : L == 12 \? (0x[0-9A-Fa-f]+) : (0x[0-9A-Fa-f]+);", src)
    if E == 4:
        c = e4
    elif E == 2:
        c = int(m2.group(1 if L & 1 else 2), 16)
    else:
        c = int(m3.group({10: 1, 11: 2, 12: 3}.get(L, 4)), 16)
    cols = [(c >> (4 * b)) & 15 for b in range(4)]
    return cols + [0] * max(0, L - 8)


def test_every_instantiated_transform_is_conflict_free():
    configs = [(L, 4) for L in (10, 11, 12, 13)]                 # first-generation 16-per-thread transforms
    configs += [(10, 2), (11, 2), (12, 2), (10, 3), (11, 3), (12, 3), (13, 3)]     # latency / throughput builds
    configs += [(8, 2), (9, 2), (10, 2), (11, 2)]                # local transforms of the 8-CTA split kernel
    for L, E in configs:
        assert conflict_free(L, E, cols_from_source(L, E)), (L, E)


def test_identity_swizzle_would_conflict():
    assert not conflict_free(11, 2, [0] * 7)
