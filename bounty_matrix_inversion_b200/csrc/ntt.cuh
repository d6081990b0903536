// Negacyclic NTT of size N = 2^L over Z_p held by one CTA of T = N / 2^E threads.
//
// Every thread owns 2^E coefficients in registers.  A "pass" runs up to E consecutive radix-2 stages entirely
// in registers (a radix-2^E butterfly network); between passes the values go through one shared-memory buffer
// of N words.  E = 4 (16 coefficients per thread) minimises shared-memory round trips and is used when many
// ciphertexts are in flight; E = 2 or 3 spreads one transform over 4x / 2x the warps, which is what hides the
// dependent-instruction latency when a launch has few ciphertexts.
//
// The buffer is indexed through an XOR swizzle: bank-select bits (index bits 0..3, 8-byte words) are XORed with
// a GF(2)-linear fold of the upper index bits, chosen per (L, E) by scripts/find_swizzle.py, which also verifies
// by brute force that every pass's 64-bit accesses are bank-conflict free for every half-warp.
//
// Forward = Cooley-Tukey with merged psi twiddles (natural order in, bit-reversed out);
// inverse = Gentleman-Sande (bit-reversed in, natural out, NOT scaled by 1/N: the
// bootstrapping key is pre-scaled instead).  tw[m + i] = psi^brv(m + i).
// Values inside a transform are "lazy" (any u64 congruent to the element, see field.cuh): inputs may be
// lazy, outputs are lazy; only the multiplied operand of each butterfly is made canonical.
#pragma once
#include "field.cuh"

template <int L, int E>
struct NttCfg {
    static constexpr int N = 1 << L;
    static constexpr int EPT = 1 << E;         // coefficients per thread
    static constexpr int T = N >> E;           // threads per CTA
    static constexpr int FULL = L / E;         // radix-2^E passes
    static constexpr int R = L % E;            // stages of the trailing partial pass
    static constexpr int NPASS = FULL + (R ? 1 : 0);
    static_assert(T <= 1024, "too many threads");
};

// fold columns for index bits 4, 5, 6, 7 (scripts/find_swizzle.py); higher bits need none
template <int L, int E>
__host__ __device__ constexpr int swz_cols() {
    if (E == 4) return 0x8421;
    if (E == 2) return (L & 1) ? 0x00C3 : 0x00A5;
    return L == 10 ? 0x0843 : L == 11 ? 0x0861 : L == 12 ? 0x0C21 : 0x0843;   // E == 3
}
template <int L, int E>
__host__ __device__ constexpr int swz(int i) {
    constexpr int c = swz_cols<L, E>();
    return i ^ ((((i >> 4) & 1) * (c & 15)) ^ (((i >> 5) & 1) * ((c >> 4) & 15)) ^ (((i >> 6) & 1) * ((c >> 8) & 15)) ^
                (((i >> 7) & 1) * ((c >> 12) & 15)));
}

// logical index of register slot q of thread tid in pass p
template <int L, int E>
__host__ __device__ constexpr int slot_index(int p, int q, int tid) {
    using C = NttCfg<L, E>;
    if (p < C::FULL) {
        const int sh = L - E * p - E;
        const int lo = tid & ((1 << sh) - 1), hi = tid >> sh;
        return (hi << (sh + E)) | (q << sh) | lo;
    }
    // partial pass: 2^(E-R) groups of 2^R contiguous words
    const int g = q >> C::R, e = q & ((1 << C::R) - 1);
    return ((g * C::T + tid) << C::R) | e;
}

// Bits contributed by the thread (q = 0) and by the register slot are disjoint, and swz() is linear over
// XOR, so swz(index) = swz(thread part) ^ swz(slot part): one XOR per access, the slot part a constant.
template <int L, int E>
__device__ __forceinline__ void store_pass(const u64 (&x)[1 << E], u64* buf, int p, int tid) {
    const int sb = swz<L, E>(slot_index<L, E>(p, 0, tid));
#pragma unroll
    for (int q = 0; q < (1 << E); q++) buf[sb ^ swz<L, E>(slot_index<L, E>(p, q, 0))] = x[q];
}
template <int L, int E>
__device__ __forceinline__ void load_pass(u64 (&x)[1 << E], const u64* buf, int p, int tid) {
    const int sb = swz<L, E>(slot_index<L, E>(p, 0, tid));
#pragma unroll
    for (int q = 0; q < (1 << E); q++) x[q] = buf[sb ^ swz<L, E>(slot_index<L, E>(p, q, 0))];
}

template <bool INV>
__device__ __forceinline__ void butterfly(u64& a, u64& b, u64 w) {
    const u64 u = a, v = b;
    if (!INV) {
        const u64 t = fmul_c(v, w);
        a = fadd_l(u, t);
        b = fsub_l(u, t);
    } else {
        const u64 vc = fcanon(v);
        a = fadd_l(u, vc);
        b = fmul_l(fsub_l(u, vc), w);
    }
}

// stages of pass p on registers; INV selects the Gentleman-Sande form and reversed stage order
template <int L, int E, bool INV>
__device__ __forceinline__ void pass_compute(u64 (&x)[1 << E], const u64* __restrict__ tw, int p, int tid) {
    using C = NttCfg<L, E>;
    constexpr int EPT = C::EPT;
    if (p < C::FULL) {
        const int sh = L - E * p - E;
        const int hi = tid >> sh;
#pragma unroll
        for (int dd = 0; dd < E; dd++) {
            const int d = INV ? E - 1 - dd : dd;
            const int half = EPT >> (d + 1);
            const int base = (1 << (E * p + d)) + (hi << d);
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                if (q & half) continue;
                butterfly<INV>(x[q], x[q + half], __ldg(tw + base + (q >> (E - d))));
            }
        }
    } else {
        constexpr int R = C::R;
#pragma unroll
        for (int dd = 0; dd < R; dd++) {
            const int d = INV ? R - 1 - dd : dd;
            const int half = 1 << (R - 1 - d);
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                const int g = q >> R, e = q & ((1 << R) - 1);
                if (e & half) continue;
                butterfly<INV>(x[q], x[q + half], __ldg(tw + (1 << (L - R + d)) + ((g * C::T + tid) << d) + (e >> (R - d))));
            }
        }
    }
}

// in: x in pass-0 layout (slot q <-> coefficient q*T + tid); out: x in last-pass layout
template <int L, int E>
__device__ __forceinline__ void ntt_forward(u64 (&x)[1 << E], u64* buf, const u64* __restrict__ tw, int tid) {
    using C = NttCfg<L, E>;
    pass_compute<L, E, false>(x, tw, 0, tid);
    __syncthreads();   // earlier readers of buf are done
#pragma unroll
    for (int p = 1; p < C::NPASS; p++) {
        store_pass<L, E>(x, buf, p - 1, tid);
        __syncthreads();
        load_pass<L, E>(x, buf, p, tid);
        pass_compute<L, E, false>(x, tw, p, tid);
    }
}

// in: x in last-pass layout; out: x in pass-0 layout, scaled by N
template <int L, int E>
__device__ __forceinline__ void ntt_inverse(u64 (&x)[1 << E], u64* buf, const u64* __restrict__ twi, int tid) {
    using C = NttCfg<L, E>;
    pass_compute<L, E, true>(x, twi, C::NPASS - 1, tid);
    __syncthreads();
#pragma unroll
    for (int p = C::NPASS - 2; p >= 0; p--) {
        store_pass<L, E>(x, buf, p + 1, tid);
        __syncthreads();
        load_pass<L, E>(x, buf, p, tid);
        pass_compute<L, E, true>(x, twi, p, tid);
    }
}
