// Arithmetic in Z_p, p = 2^64 - 2^32 + 1 (2^64 = 2^32 - 1, 2^96 = -1 mod p).
// All values handed between functions are canonical (< p) unless a comment says otherwise.
#pragma once
#include <stdint.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned int u32;

#define BMI_P 0xFFFFFFFF00000001ULL
#define BMI_EPS 0xFFFFFFFFULL

#ifdef __CUDACC__
#define BMI_HD __host__ __device__ __forceinline__
#else
#define BMI_HD inline
#endif

BMI_HD u64 fadd(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a || s >= BMI_P) s -= BMI_P;
    return s;
}
BMI_HD u64 fsub(u64 a, u64 b) {
    u64 d = a - b;
    if (a < b) d += BMI_P;
    return d;
}
BMI_HD u64 fneg(u64 a) { return a ? BMI_P - a : 0; }

// (hi:lo) mod p, any 128-bit input
BMI_HD u64 freduce128(u64 lo, u64 hi) {
    u64 hh = hi >> 32, hl = hi & BMI_EPS;
    u64 t0 = lo - hh;
    if (lo < hh) t0 -= BMI_EPS;
    u64 t1 = (hl << 32) - hl;
    u64 t2 = t0 + t1;
    if (t2 < t1) t2 += BMI_EPS;
    if (t2 >= BMI_P) t2 -= BMI_P;
    return t2;
}

BMI_HD u64 fmul(u64 a, u64 b) {
#ifdef __CUDA_ARCH__
    return freduce128(a * b, __umul64hi(a, b));
#else
    unsigned __int128 x = (unsigned __int128)a * b;
    return freduce128((u64)x, (u64)(x >> 64));
#endif
}

BMI_HD u64 fpow(u64 b, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = fmul(r, b);
        b = fmul(b, b);
        e >>= 1;
    }
    return r;
}

BMI_HD u64 from_i64(i64 v) { return v >= 0 ? (u64)v : BMI_P - (u64)(-v); }   // |v| < p

// round(x * 2N / 2^64) mod 2N
BMI_HD u32 modswitch(u64 x, int logN) { return (u32)((((x >> (62 - logN)) + 1) >> 1) & ((2ULL << logN) - 1)); }

// closest multiple of 2^(64 - bl*l), as a (bl*l)-bit integer
BMI_HD u64 round_top(u64 x, int tot) {
    u64 r = ((x >> (63 - tot)) + 1) >> 1;
    return tot < 64 ? (r & ((1ULL << tot) - 1)) : r;
}
// level j (1 = most significant) balanced digit in [-B/2, B/2) of a rounded value, as a field element
BMI_HD u64 digit_of(u64 r, int bl, int l, int j) {
    u64 B = 1ULL << bl, d = 0;
    for (int lev = l; lev >= j; lev--) {
        d = r & (B - 1);
        r >>= bl;
        if (d >= (B >> 1)) { r += 1; d = (lev == j) ? BMI_P - (B - d) : d; }
    }
    return d;
}
