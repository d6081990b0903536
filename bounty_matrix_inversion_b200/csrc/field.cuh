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

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Device fast path.  "lazy" values are any u64 congruent to the field element (possibly >= p).
// Carry chains are written in PTX so that ptxas keeps them as IADD3/IADD3.X pairs and the
// +-EPS corrections as one IMAD.WIDE; see DESIGN.md section 5 for the instruction budget.
#define BMI_LO(x) ((u32)(x))
#define BMI_HI(x) ((u32)((x) >> 32))
__device__ __forceinline__ u64 bmi_pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// canonical representative of a lazy value
__device__ __forceinline__ u64 fcanon(u64 x) {
    u32 t0, t1, c;
    asm("add.cc.u32 %0,%3,0xFFFFFFFF; addc.cc.u32 %1,%4,0; madc.lo.u32 %2,0,0,0;"
        : "=r"(t0), "=r"(t1), "=r"(c) : "r"(BMI_LO(x)), "r"(BMI_HI(x)));
    (void)t0; (void)t1;
    return x + (u64)c * 0xFFFFFFFFu;          // x >= p  <=>  x + EPS carries; then x - p == x + EPS (mod 2^64)
}
// a + b, lazy result; at least one operand must be canonical (then the +EPS cannot carry again)
__device__ __forceinline__ u64 fadd_l(u64 a, u64 b) {
    u32 s0, s1, c;
    asm("add.cc.u32 %0,%3,%5; addc.cc.u32 %1,%4,%6; madc.lo.u32 %2,0,0,0;"
        : "=r"(s0), "=r"(s1), "=r"(c) : "r"(BMI_LO(a)), "r"(BMI_HI(a)), "r"(BMI_LO(b)), "r"(BMI_HI(b)));
    return bmi_pack(s0, s1) + (u64)c * 0xFFFFFFFFu;
}
// a - b, lazy result; b must be canonical (then the -EPS cannot borrow again)
__device__ __forceinline__ u64 fsub_l(u64 a, u64 b) {
    u32 s0, s1, m;
    asm("sub.cc.u32 %0,%3,%5; subc.cc.u32 %1,%4,%6; subc.u32 %2,0,0;"
        : "=r"(s0), "=r"(s1), "=r"(m) : "r"(BMI_LO(a)), "r"(BMI_HI(a)), "r"(BMI_LO(b)), "r"(BMI_HI(b)));
    asm("sub.cc.u32 %0,%0,%2; subc.u32 %1,%1,0;" : "+r"(s0), "+r"(s1) : "r"(m));
    return bmi_pack(s0, s1);
}
// a * b, lazy operands, lazy result.  One PTX block: four 32x32->64 partial products (IMAD.WIDE), merged by two
// carry chains that ptxas folds into 3-input IADD3s, then x = x0 + x1 2^32 + x2 2^64 + x3 2^96
//   = (x1:x0) - x3 + x2 * EPS  (mod p)
__device__ __forceinline__ u64 fmul_l(u64 a, u64 b) {
    u32 r0, r1, x2, m;
    asm("{ .reg .u32 a0,a1,b0,b1,l0,h0,l1,h1,l2,h2,l3,h3,y1,y2,y3; .reg .u64 p;\n"
        "mov.b64 {a0,a1},%4; mov.b64 {b0,b1},%5;\n"
        "mul.wide.u32 p,a0,b0; mov.b64 {l0,h0},p;\n"
        "mul.wide.u32 p,a0,b1; mov.b64 {l1,h1},p;\n"
        "mul.wide.u32 p,a1,b0; mov.b64 {l2,h2},p;\n"
        "mul.wide.u32 p,a1,b1; mov.b64 {l3,h3},p;\n"
        "add.cc.u32 y1,h0,l1; addc.cc.u32 y2,l3,h1; addc.u32 y3,h3,0;\n"
        "add.cc.u32 y1,y1,l2; addc.cc.u32 y2,y2,h2; addc.u32 y3,y3,0;\n"
        "sub.cc.u32 %0,l0,y3; subc.cc.u32 %1,y1,0; subc.u32 %3,0,0;\n"      // (x1:x0) - x3, m = -borrow
        "sub.cc.u32 %0,%0,%3; subc.u32 %1,%1,0;\n"                          // - EPS on borrow (x3 < 2^32: no second borrow)
        "mov.u32 %2,y2; }"
        : "=r"(r0), "=r"(r1), "=r"(x2), "=r"(m) : "l"(a), "l"(b));
    (void)m;
    const u64 t1 = (u64)x2 * 0xFFFFFFFFu;      // < p, so the add below meets fadd_l's requirement
    return fadd_l(bmi_pack(r0, r1), t1);
}
__device__ __forceinline__ u64 fmul_c(u64 a, u64 b) { return fcanon(fmul_l(a, b)); }
#endif

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
