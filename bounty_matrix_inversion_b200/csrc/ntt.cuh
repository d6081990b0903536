// Negacyclic NTT of size N = 2^L over Z_p held by one CTA of T = N/16 threads.
//
// Every thread owns 16 coefficients in registers.  A "pass" runs up to four
// consecutive radix-2 stages entirely in registers (a radix-16 butterfly); between
// passes the 16 values go through one shared-memory buffer of N words.  The buffer is
// indexed through an XOR swizzle (low nibble ^= fold of the upper nibbles) which makes
// every pass's 64-bit accesses bank-conflict free: in each pass the 16 threads of a
// half-warp differ in four index bits that sit at four different positions mod 4.
//
// Forward = Cooley-Tukey with merged psi twiddles (natural order in, bit-reversed out);
// inverse = Gentleman-Sande (bit-reversed in, natural out, NOT scaled by 1/N: the
// bootstrapping key is pre-scaled instead).  tw[m + i] = psi^brv(m + i).
// Values inside a transform are "lazy" (any u64 congruent to the element, see field.cuh): inputs may be
// lazy, outputs are lazy; only the multiplied operand of each butterfly is made canonical.
#pragma once
#include "field.cuh"

template <int L>
struct NttCfg {
    static constexpr int N = 1 << L;
    static constexpr int T = N / 16;           // threads per CTA
    static constexpr int FULL = L / 4;         // radix-16 passes
    static constexpr int R = L % 4;            // stages of the trailing partial pass
    static constexpr int NPASS = FULL + (R ? 1 : 0);
};

__device__ __forceinline__ int swz(int i) { return i ^ (((i >> 4) ^ (i >> 8) ^ (i >> 12)) & 15); }

// logical index of register slot q of thread tid in pass p
template <int L>
__device__ __forceinline__ int slot_index(int p, int q, int tid) {
    using C = NttCfg<L>;
    if (p < C::FULL) {
        const int sh = L - 4 * p - 4;
        const int lo = tid & ((1 << sh) - 1), hi = tid >> sh;
        return (hi << (sh + 4)) | (q << sh) | lo;
    } else {  // partial pass: 2^(4-R) groups of 2^R contiguous words
        const int g = q >> C::R, e = q & ((1 << C::R) - 1);
        return ((g * C::T + tid) << C::R) | e;
    }
}

// Bits contributed by the thread (q = 0) and by the register slot are disjoint, and swz() is linear over
// XOR, so swz(index) = swz(thread part) ^ swz(slot part): one XOR per access, the slot part a constant.
template <int L>
__device__ __forceinline__ constexpr int slot_bits(int p, int q) {
    using C = NttCfg<L>;
    if (p < C::FULL) return q << (L - 4 * p - 4);
    return (((q >> C::R) * C::T) << C::R) | (q & ((1 << C::R) - 1));
}
__device__ __forceinline__ constexpr int swz_c(int i) { return i ^ (((i >> 4) ^ (i >> 8) ^ (i >> 12)) & 15); }

template <int L>
__device__ __forceinline__ void store_pass(const u64 (&x)[16], u64* buf, int p, int tid) {
    const int sb = swz(slot_index<L>(p, 0, tid));
#pragma unroll
    for (int q = 0; q < 16; q++) buf[sb ^ swz_c(slot_bits<L>(p, q))] = x[q];
}
template <int L>
__device__ __forceinline__ void load_pass(u64 (&x)[16], const u64* buf, int p, int tid) {
    const int sb = swz(slot_index<L>(p, 0, tid));
#pragma unroll
    for (int q = 0; q < 16; q++) x[q] = buf[sb ^ swz_c(slot_bits<L>(p, q))];
}

// stages of pass p on registers; INV selects the Gentleman-Sande form and reversed stage order
template <int L, bool INV>
__device__ __forceinline__ void pass_compute(u64 (&x)[16], const u64* __restrict__ tw, int p, int tid) {
    using C = NttCfg<L>;
    if (p < C::FULL) {
        const int sh = L - 4 * p - 4;
        const int hi = tid >> sh;
#pragma unroll
        for (int dd = 0; dd < 4; dd++) {
            const int d = INV ? 3 - dd : dd;
            const int half = 8 >> d;
            const int base = (1 << (4 * p + d)) + (hi << d);
#pragma unroll
            for (int q = 0; q < 16; q++) {
                if (q & half) continue;
                const u64 w = __ldg(tw + base + (q >> (4 - d)));
                const u64 u = x[q], v = x[q + half];
                if (!INV) {
                    const u64 t = fmul_c(v, w);
                    x[q] = fadd_l(u, t);
                    x[q + half] = fsub_l(u, t);
                } else {
                    const u64 vc = fcanon(v);
                    x[q] = fadd_l(u, vc);
                    x[q + half] = fmul_l(fsub_l(u, vc), w);
                }
            }
        }
    } else {
        constexpr int R = C::R;
#pragma unroll
        for (int dd = 0; dd < R; dd++) {
            const int d = INV ? R - 1 - dd : dd;
            const int half = 1 << (R - 1 - d);
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const int g = q >> R, e = q & ((1 << R) - 1);
                if (e & half) continue;
                const u64 w = __ldg(tw + (1 << (L - R + d)) + ((g * C::T + tid) << d) + (e >> (R - d)));
                const u64 u = x[q], v = x[q + half];
                if (!INV) {
                    const u64 t = fmul_c(v, w);
                    x[q] = fadd_l(u, t);
                    x[q + half] = fsub_l(u, t);
                } else {
                    const u64 vc = fcanon(v);
                    x[q] = fadd_l(u, vc);
                    x[q + half] = fmul_l(fsub_l(u, vc), w);
                }
            }
        }
    }
}

// in: x in pass-0 layout (slot q <-> coefficient q*T + tid); out: x in last-pass layout
template <int L>
__device__ __forceinline__ void ntt_forward(u64 (&x)[16], u64* buf, const u64* __restrict__ tw, int tid) {
    using C = NttCfg<L>;
    pass_compute<L, false>(x, tw, 0, tid);
    __syncthreads();   // earlier readers of buf are done
#pragma unroll
    for (int p = 1; p < C::NPASS; p++) {
        store_pass<L>(x, buf, p - 1, tid);
        __syncthreads();
        load_pass<L>(x, buf, p, tid);
        pass_compute<L, false>(x, tw, p, tid);
    }
}

// in: x in last-pass layout; out: x in pass-0 layout, scaled by N
template <int L>
__device__ __forceinline__ void ntt_inverse(u64 (&x)[16], u64* buf, const u64* __restrict__ twi, int tid) {
    using C = NttCfg<L>;
    pass_compute<L, true>(x, twi, C::NPASS - 1, tid);
    __syncthreads();
#pragma unroll
    for (int p = C::NPASS - 2; p >= 0; p--) {
        store_pass<L>(x, buf, p + 1, tid);
        __syncthreads();
        load_pass<L>(x, buf, p, tid);
        pass_compute<L, true>(x, twi, p, tid);
    }
}
