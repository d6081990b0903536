"""Searches, for every (L, E), the GF(2) columns that fold index bits >= 4 into the 4 bank-select bits so that
every pass of the register-radix NTT touches 16 distinct 8-byte banks per half-warp; verifies by brute force.
Prints the constexpr table pasted into csrc/ntt.cuh."""
import itertools
import sys


def slot_index(L, E, p, q, tid):
    T, FULL, R = (1 << L) >> E, L // E, L % E
    if p < FULL:
        sh = L - E * p - E
        lo, hi = tid & ((1 << sh) - 1), tid >> sh
        return (hi << (sh + E)) | (q << sh) | lo
    g, e = q >> R, q & ((1 << R) - 1)
    return ((g * T + tid) << R) | e


def phys(i, cols):
    f = 0
    for b, c in enumerate(cols):
        if (i >> (4 + b)) & 1:
            f ^= c
    return i ^ f


def conflict_free(L, E, cols):
    T, FULL, R = (1 << L) >> E, L // E, L % E
    npass = FULL + (1 if R else 0)
    for p in range(npass):
        for q in range(1 << E):
            for h0 in range(0, T, 16):
                banks = {phys(slot_index(L, E, p, q, t), cols) & 15 for t in range(h0, min(h0 + 16, T))}
                if len(banks) != min(16, T - h0):
                    return False
    # the transform's input/output side: slot q of thread tid <-> coefficient q*T + tid is register-only, no check needed
    return True


def search(L, E):
    nb = L - 4
    # prefer few non-zero columns: try all assignments of the first 4 high bits, rest zero, then widen
    for width in range(1, nb + 1):
        for combo in itertools.product(range(16), repeat=width):
            cols = list(combo) + [0] * (nb - width)
            if conflict_free(L, E, cols):
                return cols
    return None


if __name__ == "__main__":
    for E in (2, 3, 4):
        for L in (10, 11, 12, 13):
            if (1 << L) >> E > 1024:
                print(f"// L={L} E={E}: more than 1024 threads, not used")
                continue
            cols = search(L, E)
            print(f"L={L} E={E} cols={cols}", flush=True)
