// Keyswitch and leveled (linear) kernels of the hot path: not templated on the polynomial size, compiled once with
// the context unit (engine.cu).
#pragma once
#include "field.cuh"

// ----------------------------------------------------------------------------
// keyswitch big (kN) -> small (n): out = (0, b) - sum_i sum_j digit_ij * ksk[i][j]
// The balanced digits d in [-B/2, B/2) are used as d' = d + B/2 in [0, B): sum d*ksk = sum d'*ksk - corr with
// corr[t] = B/2 * sum_ij ksk[i][j][t], a per-key constant (keyswitch_corr_kernel).  All products are then unsigned
// 32 x 64 bit and accumulate unreduced in 96 bits: five integer instructions per multiply-accumulate.
// CTA = 128 output columns x KS_JT ciphertexts x one slice of the kN input coefficients (blockIdx.z); digits of a
// 64-coefficient chunk are staged in shared memory, so every keyswitch-key word read from L2/HBM is used for
// KS_JT ciphertexts.  With more than one slice the per-slice sums go to `partial` and keyswitch_finish_kernel
// folds them.
constexpr int KS_COLS = 128, KS_JT = 8, KS_CHUNK = 64;

__global__ void __launch_bounds__(256) keyswitch_corr_kernel(const u64* __restrict__ ksk, u64* __restrict__ corr, int rows, int n, int bl) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n) return;
    u64 s = 0;
    for (int r = 0; r < rows; r++) s = fadd(s, ksk[(size_t)r * (n + 1) + t]);
    corr[t] = fmul(s, 1ULL << (bl - 1));
}

__global__ void __launch_bounds__(KS_COLS) keyswitch_kernel(const u64* __restrict__ in, const u64* __restrict__ ksk,
                                                            const u64* __restrict__ corr, u64* __restrict__ out,
                                                            u64* __restrict__ partial, int M, int kN, int n, int bl, int l,
                                                            int slice) {
    extern __shared__ u32 dg[];   // [KS_CHUNK][l][KS_JT] shifted digits d' = d + B/2
    // ciphertext tiles along gridDim.x (2^31 - 1 blocks), output-column groups along y, input slices along z
    const int tid = threadIdx.x, t = blockIdx.y * KS_COLS + tid;
    const int job0 = blockIdx.x * KS_JT;
    const int njob = min(KS_JT, M - job0);
    const int tot = bl * l;
    const bool live = t <= n;
    const int i_begin = blockIdx.z * slice, i_end = min(kN, i_begin + slice);
    u32 a0[KS_JT], a1[KS_JT], a2[KS_JT];
#pragma unroll
    for (int jb = 0; jb < KS_JT; jb++) { a0[jb] = 0; a1[jb] = 0; a2[jb] = 0; }

    for (int i0 = i_begin; i0 < i_end; i0 += KS_CHUNK) {
        const int ci = min(KS_CHUNK, i_end - i0);
        __syncthreads();
        for (int e = tid; e < KS_CHUNK * KS_JT; e += KS_COLS) {
            const int jb = e / KS_CHUNK, ii = e % KS_CHUNK;
            const bool valid = jb < njob && ii < ci;
            u64 r = valid ? round_top(in[(size_t)(job0 + jb) * (kN + 1) + i0 + ii], tot) : 0;
            const u32 B = 1u << bl;
            for (int lev = l; lev >= 1; lev--) {
                const u32 d = (u32)r & (B - 1);
                r >>= bl;
                u32 dp = d + (B >> 1);                   // balanced digit d (or d - B) plus B/2
                if (d >= (B >> 1)) { r += 1; dp = d - (B >> 1); }
                dg[(ii * l + (lev - 1)) * KS_JT + jb] = valid ? dp : 0;   // rows beyond the slice contribute nothing
            }
        }
        __syncthreads();
        if (live) {
            const u64* kp = ksk + (size_t)i0 * l * (n + 1) + t;
            const int rows = ci * l;
#pragma unroll 4
            for (int r = 0; r < rows; r++) {
                const u64 kv = __ldg(kp + (size_t)r * (n + 1));
                const u32 k0 = (u32)kv, k1 = (u32)(kv >> 32);
                const uint4* d4 = reinterpret_cast<const uint4*>(dg + r * KS_JT);
                const uint4 dA = d4[0], dB = d4[1];
                const u32 d[KS_JT] = {dA.x, dA.y, dA.z, dA.w, dB.x, dB.y, dB.z, dB.w};
#pragma unroll
                for (int jb = 0; jb < KS_JT; jb++) {
                    asm("mad.lo.cc.u32 %0,%3,%4,%0; madc.hi.cc.u32 %1,%3,%4,%1; addc.u32 %2,%2,0;"
                        "mad.lo.cc.u32 %1,%3,%5,%1; madc.hi.u32 %2,%3,%5,%2;"
                        : "+r"(a0[jb]), "+r"(a1[jb]), "+r"(a2[jb]) : "r"(d[jb]), "r"(k0), "r"(k1));
                }
            }
        }
    }
    if (live) {
#pragma unroll
        for (int jb = 0; jb < KS_JT; jb++) {
            if (jb >= njob) break;
            const u64 sum = freduce128(((u64)a1[jb] << 32) | a0[jb], (u64)a2[jb]);
            if (gridDim.z == 1) {
                const u64 base = t == n ? in[(size_t)(job0 + jb) * (kN + 1) + kN] : 0;
                out[(size_t)(job0 + jb) * (n + 1) + t] = fsub(fadd(base, corr[t]), sum);
            } else {
                partial[((size_t)blockIdx.z * M + job0 + jb) * (n + 1) + t] = sum;
            }
        }
    }
}

__global__ void __launch_bounds__(256) keyswitch_finish_kernel(const u64* __restrict__ in, const u64* __restrict__ partial,
                                                               const u64* __restrict__ corr, u64* __restrict__ out, int M,
                                                               int kN, int n, int slices) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (size_t)M * (n + 1)) return;
    const int job = (int)(e / (n + 1)), t = (int)(e % (n + 1));
    u64 s = 0;
    for (int z = 0; z < slices; z++) s = fadd(s, partial[(size_t)z * M * (n + 1) + e]);
    const u64 base = t == n ? in[(size_t)job * (kN + 1) + kN] : 0;
    out[e] = fsub(fadd(base, corr[t]), s);
}

// ----------------------------------------------------------------------------
// leveled linear combinations: out[j][b] = sum_t coef[t] * vals[idx[t]][b] + konst[j] (on the body)
__global__ void __launch_bounds__(256) lincomb_kernel(const u64* __restrict__ vals, const int* __restrict__ row_ptr,
                                                      const int* __restrict__ idx, const u64* __restrict__ coef,
                                                      const u64* __restrict__ konst, u64* __restrict__ out,
                                                      int W, int batch) {
    const int f = blockIdx.x;
    const int j = f / batch, b = f - j * batch;
    const int t0 = row_ptr[j], t1 = row_ptr[j + 1];
    u64* o = out + (size_t)f * W;
    for (int w = blockIdx.y * blockDim.x + threadIdx.x; w < W; w += gridDim.y * blockDim.x) {
        u64 s = w == W - 1 ? konst[j] : 0;
        for (int t = t0; t < t1; t++) {
            const u64 c = coef[t];
            const u64 v = vals[((size_t)idx[t] * batch + b) * W + w];
            if (c == 1) s = fadd(s, v);
            else if (c == BMI_P - 1) s = fsub(s, v);
            else s = fadd(s, fmul(c, v));
        }
        o[w] = s;
    }
}

// ----------------------------------------------------------------------------
// pair key, transform domain, any layout: [pair][K11, K10, K01][2 rows][2 outputs][N] -> [pair][K11, K11+K10, K11+K01][..]
__global__ void __launch_bounds__(256) pair_key_sums_kernel(u64* __restrict__ key, size_t per_pair_third, size_t pairs) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= per_pair_third * pairs) return;
    const size_t q = e / per_pair_third, w = e % per_pair_third;
    u64* base = key + q * 3 * per_pair_third + w;
    const u64 k11 = base[0];
    base[per_pair_third] = fadd(k11, base[per_pair_third]);
    base[2 * per_pair_third] = fadd(k11, base[2 * per_pair_third]);
}

// ----------------------------------------------------------------------------
// rows of W words: dst[dst_row[j] * batch + b] = src[j * batch + b]  (all-gathered bootstrap outputs -> value slots)
__global__ void __launch_bounds__(256) scatter_rows_kernel(const u64* __restrict__ src, const int* __restrict__ dst_row,
                                                           u64* __restrict__ dst, int W, int batch) {
    const int f = blockIdx.x;
    const int j = f / batch, b = f - j * batch;
    const u64* s = src + (size_t)f * W;
    u64* d = dst + ((size_t)dst_row[j] * batch + b) * W;
    for (int w = blockIdx.y * blockDim.x + threadIdx.x; w < W; w += gridDim.y * blockDim.x) d[w] = s[w];
}
