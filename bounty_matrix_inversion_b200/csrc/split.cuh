// Bootstrap on an 8-CTA cluster: each of the two accumulator polynomials is spread over S = 4 CTAs (CTA r of a
// group holds the coefficients i = r mod 4), so one ciphertext uses 8 SMs and each CTA transforms only N/4 points.
//
// Distributed negacyclic NTT of size N = 2^L over the 4 CTAs of a group (M = N/4 points, T = M/4 threads each):
//   forward : the first L-2 Cooley-Tukey stages pair positions that are congruent mod 4, i.e. they are a size-M
//             transform local to each CTA (the twiddles tw[1..M) of the size-N table ARE the size-M table);
//             one all-to-all over DSMEM then hands every CTA complete blocks of 4 consecutive positions, on which
//             the last two stages run in registers (one block per thread).
//   inverse : the mirror image (two Gentleman-Sande stages on blocks, all-to-all back, local size-M inverse).
// After the forward transform CTA r holds positions [r*M, (r+1)*M); thread t holds (r*T + t)*4 + e, e = 0..3.
#pragma once
#include "kernels.cuh"

constexpr int kSplit = 4;        // CTAs per polynomial
constexpr int kSplitE = 2;       // coefficients per thread (log2) == log2(kSplit): one block per thread

template <int L>
struct SplitCfg {
    static constexpr int LL = L - 2;             // local transform size (log2)
    static constexpr int N = 1 << L, M = N / 4, T = M / 4;
    using Local = NttCfg<L - 2, kSplitE>;
};


// x: 4 registers in pass-0 layout of the local transform (slot q <-> local coefficient q*T + tid <-> global
// coefficient (q*T + tid)*4 + r).  Out: y[e] = transform value at position (r*T + tid)*4 + e.
// gbuf: [M] words in every CTA of the group; group0 = cluster rank of the group's CTA 0.
template <int L>
__device__ __forceinline__ void split_forward(u64 (&x)[4], u64* buf, u64* gbuf, const u64* __restrict__ tw, int tid,
                                              u32 group0, u32 r) {
    using C = SplitCfg<L>;
    using LC = typename C::Local;
    ntt_forward<C::LL, kSplitE>(x, buf, tw, tid);
#pragma unroll
    for (int q = 0; q < 4; q++) {     // all-to-all: local position jl goes to the CTA that owns block jl
        const int jl = slot_index<C::LL, kSplitE>(LC::NPASS - 1, q, tid);
        const u32 dst = (u32)(jl >> (C::LL - 2));
        const int b = jl & (C::T - 1);
        cluster_map(gbuf, group0 + dst)[b * 4 + (int)r] = x[q];
    }
    cluster_sync_all();
    const ulonglong2 v01 = *reinterpret_cast<const ulonglong2*>(gbuf + tid * 4);
    const ulonglong2 v23 = *reinterpret_cast<const ulonglong2*>(gbuf + tid * 4 + 2);
    x[0] = v01.x; x[1] = v01.y; x[2] = v23.x; x[3] = v23.y;
    const int blk = (int)r * C::T + tid;
    const u64 wa = __ldg(tw + (1 << (L - 2)) + blk);
    butterfly<false>(x[0], x[2], wa);
    butterfly<false>(x[1], x[3], wa);
    butterfly<false>(x[0], x[1], __ldg(tw + (1 << (L - 1)) + 2 * blk));
    butterfly<false>(x[2], x[3], __ldg(tw + (1 << (L - 1)) + 2 * blk + 1));
}

// in: z[e] at position (r*T + tid)*4 + e.  Out: 4 registers in pass-0 layout of the local transform, times N.
// lbuf: [M] words in every CTA of the group.
template <int L>
__device__ __forceinline__ void split_inverse(u64 (&z)[4], u64* buf, u64* lbuf, const u64* __restrict__ twi, int tid,
                                              u32 group0, u32 r) {
    using C = SplitCfg<L>;
    using LC = typename C::Local;
    const int blk = (int)r * C::T + tid;
    butterfly<true>(z[0], z[1], __ldg(twi + (1 << (L - 1)) + 2 * blk));
    butterfly<true>(z[2], z[3], __ldg(twi + (1 << (L - 1)) + 2 * blk + 1));
    const u64 wa = __ldg(twi + (1 << (L - 2)) + blk);
    butterfly<true>(z[0], z[2], wa);
    butterfly<true>(z[1], z[3], wa);
#pragma unroll
    for (int e = 0; e < 4; e++) cluster_map(lbuf, group0 + e)[blk] = z[e];   // position blk*4+e lives in CTA e, local blk
    cluster_sync_all();
#pragma unroll
    for (int q = 0; q < 4; q++) z[q] = lbuf[slot_index<C::LL, kSplitE>(LC::NPASS - 1, q, tid)];
    ntt_inverse<C::LL, kSplitE>(z, buf, twi, tid);
}

// self-test: c = a * b mod (X^N + 1) for one polynomial pair per 4-CTA cluster
template <int L>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(SplitCfg<L>::T) polymul_split_kernel(
    const u64* __restrict__ a, const u64* __restrict__ b, u64* __restrict__ c, const u64* __restrict__ tw,
    const u64* __restrict__ twi, u64 ninv) {
    using C = SplitCfg<L>;
    extern __shared__ u64 smem[];
    u64 *buf = smem, *gbuf = smem + C::M, *lbuf = smem + 2 * C::M;
    const int tid = threadIdx.x;
    const u32 r = cluster_rank();
    const size_t off = (size_t)(blockIdx.x / 4) * C::N;
    u64 x[4], y[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int gi = (q * C::T + tid) * 4 + (int)r;
        x[q] = a[off + gi];
        y[q] = b[off + gi];
    }
    split_forward<L>(x, buf, gbuf, tw, tid, 0, r);
    cluster_sync_all();               // gbuf is reused by the second transform
    split_forward<L>(y, buf, gbuf, tw, tid, 0, r);
#pragma unroll
    for (int e = 0; e < 4; e++) x[e] = fmul_l(fmul_l(x[e], y[e]), ninv);
    split_inverse<L>(x, buf, lbuf, twi, tid, 0, r);
#pragma unroll
    for (int q = 0; q < 4; q++) c[off + (q * C::T + tid) * 4 + (int)r] = fcanon(x[q]);
}

// ----------------------------------------------------------------------------
// key conversion for the split layout: [poly][r][e][t] <- transform value at position (r*T + t)*4 + e
template <int L, int E>
__global__ void __launch_bounds__(NttCfg<L, E>::T) bsk_convert_split_kernel(const u64* __restrict__ src, u64* __restrict__ dst,
                                                                            const u64* __restrict__ tw, u64 ninv) {
    using C = NttCfg<L, E>;
    using S = SplitCfg<L>;
    extern __shared__ u64 smem[];
    const int tid = threadIdx.x;
    const u64* s = src + (size_t)blockIdx.x * C::N;
    u64* d = dst + (size_t)blockIdx.x * C::N;
    u64 x[C::EPT];
#pragma unroll
    for (int q = 0; q < C::EPT; q++) x[q] = s[q * C::T + tid];
    ntt_forward<L, E>(x, smem, tw, tid);
#pragma unroll
    for (int q = 0; q < C::EPT; q++) {
        const int P = slot_index<L, E>(C::NPASS - 1, q, tid);
        const int blk = P >> 2, e = P & 3;
        d[(blk / S::T) * S::M + e * S::T + (blk % S::T)] = fmul_c(x[q], ninv);
    }
}

// ----------------------------------------------------------------------------
// programmable bootstrap, one 8-CTA cluster per ciphertext (l == 1, k == 1).  Cluster rank = c*4 + r:
// c = accumulator polynomial (0 mask, 1 body), r = coefficient residue mod 4 held by the CTA.
template <int L>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(SplitCfg<L>::T, 1) pbs_split_kernel(const PbsArgs a) {
    using C = SplitCfg<L>;
    constexpr int N = C::N, M = C::M, T = C::T;
    extern __shared__ u64 smem[];
    u64* acc = smem;                  // [M] coefficients i = r mod 4 of polynomial c, local index i / 4 (canonical)
    u64* buf = smem + M;              // [M] local transform exchange buffer
    u64* gbuf = smem + 2 * M;         // [M] forward all-to-all landing zone
    u64* lbuf = smem + 3 * M;         // [M] inverse all-to-all landing zone
    u64* recv = smem + 4 * M;         // [M] partial sums from the partner polynomial's CTA
    unsigned short* rot = reinterpret_cast<unsigned short*>(smem + 5 * M);
    const int tid = threadIdx.x;
    const u32 rank = cluster_rank(), c = rank >> 2, r = rank & 3, group0 = c << 2;
    u64* peer_recv = cluster_map(recv, ((c ^ 1) << 2) + r);
    const int n = a.n, bl = a.bl;
    const int total = a.njobs * a.batch;

    for (int f = blockIdx.x >> 3; f < total; f += gridDim.x >> 3) {
        const int q0 = f / a.batch, b0 = f - q0 * a.batch;
        const u64* in = a.small + ((size_t)a.job_in[q0] * a.batch + b0) * (n + 1);
        const u64* lut = a.luts + (size_t)a.job_lut[q0] * N;
        u64* out = a.out + ((size_t)a.job_out[q0] * a.batch + b0) * (N + 1);

        cluster_sync_all();           // every CTA is done with the previous ciphertext's accumulator
        for (int i = tid; i < n; i += T) rot[i] = (unsigned short)modswitch(in[i], L);
        {   // acc = (0, X^{-b~} * lut), this CTA's residue class
            const u32 r0 = (2 * N - modswitch(in[n], L)) & (2 * N - 1);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int lc = q * T + tid, gi = lc * 4 + (int)r;
                const u32 u = (gi + 2 * N - r0) & (2 * N - 1);
                acc[lc] = c == 0 ? 0 : (u < N ? lut[u] : fneg(lut[u - N]));
            }
        }
        __syncthreads();

        for (int i = 0; i < n; i++) {
            const u32 at = rot[i];
            if (at == 0) continue;    // same decision in all 8 CTAs
            cluster_sync_all();       // accumulator updates of the previous CMUX are visible cluster-wide
            u64 x[4];
            {   // digit of (X^at - 1) * acc: the rotated coefficients of this residue class all live in ONE CTA of the group
                const u64* src = cluster_map(acc, group0 + ((r - at) & 3));
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int lc = q * T + tid, gi = lc * 4 + (int)r;
                    const u32 u = (gi + 2 * N - at) & (2 * N - 1);
                    const u64 v = src[(u & (N - 1)) >> 2];
                    const u64 rv = u < N ? v : fneg(v);
                    x[q] = digit_of(round_top(fsub(rv, acc[lc]), bl), bl, 1, 1);
                }
            }
            split_forward<L>(x, buf, gbuf, a.tw, tid, group0, r);
            const u64* g = a.bsk_hat + ((size_t)i * 2 + c) * 2 * N + (size_t)r * M;
            u64 own[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
                peer_recv[e * T + tid] = fmul_c(x[e], __ldg(g + (size_t)(c ^ 1) * N + e * T + tid));
                own[e] = fmul_l(x[e], __ldg(g + (size_t)c * N + e * T + tid));
            }
            cluster_sync_all();
#pragma unroll
            for (int e = 0; e < 4; e++) own[e] = fadd_l(own[e], recv[e * T + tid]);
            split_inverse<L>(own, buf, lbuf, a.twi, tid, group0, r);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int lc = q * T + tid;
                acc[lc] = fcanon(fadd_l(own[q], acc[lc]));
            }
        }

        cluster_sync_all();
        // sample extraction: out[0] = A[0], out[t] = -A[N - t]; coefficient N - t is held by residue (-t) mod 4
        if (c == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int lc = q * T + tid, gi = lc * 4 + (int)r;      // this CTA's coefficient gi goes to out[(N - gi) % N]
                const u64 v = acc[lc];
                if (gi == 0) out[0] = v; else out[N - gi] = fneg(v);
            }
        } else if (r == 0 && tid == 0) {
            out[N] = acc[0];
        }
    }
    cluster_sync_all();               // nobody exits while its shared memory may still be read
}


// ----------------------------------------------------------------------------
// same bootstrap, synchronised by mbarriers instead of cluster barriers.  Per CMUX and CTA:
//   acc_bar  (4 arrivals)  every CTA of the group has finished updating its accumulator  -> rotated reads may start
//   fwd_bar  (M*8 bytes)   the forward all-to-all has landed in gbuf
//   rcv_bar  (M*8 bytes)   the partner polynomial's partial sums have landed in recv[parity]
//   inv_bar  (M*8 bytes)   the inverse all-to-all has landed in lbuf
// Buffer reuse is safe without back-pressure: a CTA can only reach the next write of a buffer after waiting for data
// that its reader sends AFTER the read (gbuf, lbuf, acc), or the buffer is double-buffered (recv).
// (Folding the partial-sum exchange into the inverse all-to-all via linearity of the inverse transform was tried:
//  same time, more DSMEM traffic -- not kept.)
// PAIR: two key bits per step with the pair key (kernels.cuh).  The accumulator itself is decomposed, so no CTA reads
// another CTA's accumulator and the acc_bar round disappears along with half of the steps.
#ifndef BMI_SPLIT_MINB
#define BMI_SPLIT_MINB 1
#endif
template <int L, bool PAIR = false>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(SplitCfg<L>::T, BMI_SPLIT_MINB) pbs_split_async_kernel(const PbsArgs a) {
    using C = SplitCfg<L>;
    using LC = typename C::Local;
    constexpr int N = C::N, M = C::M, T = C::T;
    extern __shared__ u64 smem[];
    u64* acc = smem;
    u64* buf = smem + M;
    u64* gbuf = smem + 2 * M;
    u64* lbuf = smem + 3 * M;
    u64* recv = smem + 4 * M;         // [2][M]
    constexpr bool PW_SMEM = PAIR && L <= 12;  // the 2N powers of psi in shared memory (32 KB at N = 2048)
    u64* pws = smem + 6 * M;
    unsigned short* rot = reinterpret_cast<unsigned short*>(smem + 6 * M + (PW_SMEM ? 2 * N : 0));
    __shared__ __align__(8) u64 bars[4];      // acc, fwd, rcv, inv
    const int tid = threadIdx.x;
    const u32 rank = cluster_rank(), c = rank >> 2, r = rank & 3, group0 = c << 2, partner = ((c ^ 1) << 2) + r;
    const int n = a.n, bl = a.bl;
    const int total = a.njobs * a.batch;

    if (tid == 0) {
        mbar_init(&bars[0], 4);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
        mbar_init(&bars[3], 1);
    }
    if constexpr (PW_SMEM)
        for (int i = tid; i < 2 * N; i += T) pws[i] = __ldg(a.pw + i);
    __syncthreads();
    cluster_sync_all();               // every CTA's barriers exist before anyone signals them
    // per-thread constants of the whole kernel: the twiddles of the two block stages, and (pair rotation) the
    // exponents of this thread's four transform slots
    const int blk = (int)r * T + tid;
    const u64 cwa = __ldg(a.tw + (1 << (L - 2)) + blk), cwb0 = __ldg(a.tw + (1 << (L - 1)) + 2 * blk), cwb1 = __ldg(a.tw + (1 << (L - 1)) + 2 * blk + 1);
    const u64 ciwa = __ldg(a.twi + (1 << (L - 2)) + blk), ciwb0 = __ldg(a.twi + (1 << (L - 1)) + 2 * blk), ciwb1 = __ldg(a.twi + (1 << (L - 1)) + 2 * blk + 1);
    u32 ex[4] = {0, 0, 0, 0};
    if constexpr (PAIR) {
#pragma unroll
        for (int e = 0; e < 4; e++) ex[e] = __ldg(a.expo + (u32)r * M + e * T + tid);
    }
    const u64* pwt = PW_SMEM ? pws : a.pw;
    u32 gbuf_r[4], lbuf_r[4], fwd_bar_r[4], inv_bar_r[4], acc_bar_r[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        gbuf_r[g] = map_shared_u32(gbuf, group0 + g);
        lbuf_r[g] = map_shared_u32(lbuf, group0 + g);
        acc_bar_r[g] = map_shared_u32(&bars[0], group0 + g);
        fwd_bar_r[g] = map_shared_u32(&bars[1], group0 + g);
        inv_bar_r[g] = map_shared_u32(&bars[3], group0 + g);
    }
    const u32 recv_r = map_shared_u32(recv, partner), rcv_bar_r = map_shared_u32(&bars[2], partner);
    u32 ph_acc = 0, ph = 0;           // ph: common parity of fwd/rcv/inv (each used exactly once per CMUX)

    for (int f = blockIdx.x >> 3; f < total; f += gridDim.x >> 3) {
        const int q0 = f / a.batch, b0 = f - q0 * a.batch;
        const u64* in = a.small + ((size_t)a.job_in[q0] * a.batch + b0) * (n + 1);
        const u64* lut = a.luts + (size_t)a.job_lut[q0] * N;
        u64* out = a.out + ((size_t)a.job_out[q0] * a.batch + b0) * (N + 1);

        for (int i = tid; i < n; i += T) rot[i] = (unsigned short)modswitch(in[i], L);
        {
            const u32 r0 = (2 * N - modswitch(in[n], L)) & (2 * N - 1);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int lc = q * T + tid, gi = lc * 4 + (int)r;
                const u32 u = (gi + 2 * N - r0) & (2 * N - 1);
                acc[lc] = c == 0 ? 0 : (u < N ? lut[u] : fneg(lut[u - N]));
            }
        }
        __syncthreads();
        if (!PAIR && tid < 4) mbar_arrive_remote(acc_bar_r[tid]);      // my accumulator is ready

        for (int i = 0; i < n; i += PAIR ? 2 : 1) {
            const u32 at = rot[i], at2 = PAIR ? rot[i + 1] : 0;
            if ((at | at2) == 0) continue;
            if (tid == 0) {           // post what this CMUX will receive
                mbar_expect(&bars[1], M * 8);
                mbar_expect(&bars[2], M * 8);
                mbar_expect(&bars[3], M * 8);
            }
            u64 x[4];
            if (PAIR) {
#pragma unroll
                for (int q = 0; q < 4; q++) x[q] = digit_of(round_top(acc[q * T + tid], bl), bl, 1, 1);
            } else {
                mbar_wait_cluster(&bars[0], ph_acc);
                ph_acc ^= 1;
                const u64* src = cluster_map(acc, group0 + ((r - at) & 3));
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int lc = q * T + tid, gi = lc * 4 + (int)r;
                    const u32 u = (gi + 2 * N - at) & (2 * N - 1);
                    const u64 v = src[(u & (N - 1)) >> 2];
                    const u64 rv = u < N ? v : fneg(v);
                    x[q] = digit_of(round_top(fsub(rv, acc[lc]), bl), bl, 1, 1);
                }
            }
            // ---- forward: local stages, all-to-all, two block stages
            ntt_forward<C::LL, kSplitE>(x, buf, a.tw, tid);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int jl = slot_index<C::LL, kSplitE>(LC::NPASS - 1, q, tid);
                const int dst = jl >> (C::LL - 2), b = jl & (T - 1);
                st_async_u64(gbuf_r[dst] + (u32)(b * 4 + (int)r) * 8, x[q], fwd_bar_r[dst]);
            }
            mbar_wait(&bars[1], ph);
            {
                const ulonglong2 v01 = *reinterpret_cast<const ulonglong2*>(gbuf + tid * 4);
                const ulonglong2 v23 = *reinterpret_cast<const ulonglong2*>(gbuf + tid * 4 + 2);
                x[0] = v01.x; x[1] = v01.y; x[2] = v23.x; x[3] = v23.y;
            }
            butterfly<false>(x[0], x[2], cwa);
            butterfly<false>(x[1], x[3], cwa);
            butterfly<false>(x[0], x[1], cwb0);
            butterfly<false>(x[2], x[3], cwb1);
            // ---- pointwise products; the partner polynomial's share goes to the partner CTA
            u64 own[4];
            const u32 rbase = recv_r + (ph ? (u32)M * 8 : 0);
            if (PAIR) {
                // pair key: [pair][K11, K10, K01][row = decomposed polynomial][output polynomial][N], split layout within N
                const u64* gp = a.bsk_hat + ((size_t)(i >> 1) * 6 + c) * 2 * N + (size_t)r * M;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int idx = e * T + tid;
                    PairMono m;
                    m.m10 = fsub_l(pwt[(at * ex[e]) & (2 * N - 1)], 1);
                    m.m01 = fsub_l(pwt[(at2 * ex[e]) & (2 * N - 1)], 1);
                    const size_t oc = (size_t)(c ^ 1) * N + idx, mc = (size_t)c * N + idx;
                    const u64 ko = pair_combine(m, __ldg(gp + oc), __ldg(gp + 4 * N + oc), __ldg(gp + 8 * N + oc));
                    st_async_u64(rbase + (u32)idx * 8, fmul_c(x[e], ko), rcv_bar_r);
                    const u64 km = pair_combine(m, __ldg(gp + mc), __ldg(gp + 4 * N + mc), __ldg(gp + 8 * N + mc));
                    own[e] = fmul_l(x[e], km);
                }
            } else {
                const u64* g = a.bsk_hat + ((size_t)i * 2 + c) * 2 * N + (size_t)r * M;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    st_async_u64(rbase + (u32)(e * T + tid) * 8, fmul_c(x[e], __ldg(g + (size_t)(c ^ 1) * N + e * T + tid)), rcv_bar_r);
                    own[e] = fmul_l(x[e], __ldg(g + (size_t)c * N + e * T + tid));
                }
            }
            mbar_wait(&bars[2], ph);
            {
                const u64* rb = recv + (ph ? M : 0);
#pragma unroll
                for (int e = 0; e < 4; e++) own[e] = fadd_l(own[e], rb[e * T + tid]);
            }
            // ---- inverse: two block stages, all-to-all, local stages
            butterfly<true>(own[0], own[1], ciwb0);
            butterfly<true>(own[2], own[3], ciwb1);
            butterfly<true>(own[0], own[2], ciwa);
            butterfly<true>(own[1], own[3], ciwa);
#pragma unroll
            for (int e = 0; e < 4; e++) st_async_u64(lbuf_r[e] + (u32)blk * 8, own[e], inv_bar_r[e]);
            mbar_wait(&bars[3], ph);
            ph ^= 1;
#pragma unroll
            for (int q = 0; q < 4; q++) own[q] = lbuf[slot_index<C::LL, kSplitE>(LC::NPASS - 1, q, tid)];
            ntt_inverse<C::LL, kSplitE>(own, buf, a.twi, tid);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int lc = q * T + tid;
                acc[lc] = fcanon(fadd_l(own[q], acc[lc]));
            }
            __syncthreads();
            if (!PAIR && tid < 4) mbar_arrive_remote(acc_bar_r[tid]);
        }

        if (!PAIR) {
            mbar_wait_cluster(&bars[0], ph_acc);      // consume the last "accumulator ready" round
            ph_acc ^= 1;
        }
        if (c == 0) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int lc = q * T + tid, gi = lc * 4 + (int)r;
                const u64 v = acc[lc];
                if (gi == 0) out[0] = v; else out[N - gi] = fneg(v);
            }
        } else if (r == 0 && tid == 0) {
            out[N] = acc[0];
        }
        __syncthreads();              // extraction done before the next ciphertext overwrites acc / rot
    }
    cluster_sync_all();
}
