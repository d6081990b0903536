// Lowest-latency bootstrap: one 8-CTA cluster per ciphertext, TWO points per thread, pair blind rotation.
//
// Same distribution as split.cuh (each accumulator polynomial over 4 CTAs, CTA r holds the coefficients
// i = r mod 4, M = N/4 of them), but every butterfly of a stage has its own thread (T = M/2 threads per CTA,
// two warps per scheduler at N = 2048), so the dependent chain per thread is one butterfly per stage:
//   local stages   : radix-2, both points of a butterfly in one thread; between stages ONE value per thread moves
//                    to the partner thread -- through shared memory while the partner is in another warp (the first
//                    L-8 exchanges), then by warp shuffles (the last five): no barrier in the shuffle stages
//   cross-CTA part : as in split.cuh (st.async all-to-all into blocks of 4 consecutive positions, mbarrier
//                    completion), a block being handled by two neighbouring lanes with one shuffle between the
//                    last two stages
//   tables         : every twiddle a thread ever uses is fixed for the whole kernel and lives in registers; the
//                    2N powers of psi sit in shared memory; the step's 12 key words are fetched before the
//                    forward transform starts.  The loop body waits on no global load.
//   pointwise      : with KA = K11 + K10 and KB = K11 + K01 (formed once at key load) the combined GGSW of a step is
//                    m10 (m01 K11 + KA) + m01 KB  with m = X^a - 1 at the slot: eight products per slot and row.
// After the forward transform thread t of CTA (c, r) holds positions 2 (r T + t) + e, e = 0, 1.
#pragma once
#include "split.cuh"

template <int L>
struct Split2Cfg {
    static constexpr int LL = L - 2;                 // log2 of the local transform
    static constexpr int N = 1 << L, M = N / 4, T = M / 2;
    static constexpr int TB = LL - 1;                // thread-index bits
    static constexpr int NSM = TB > 5 ? TB - 5 : 0;  // exchanges that cross warps (shared memory)
    static constexpr bool PW_SMEM = L <= 12;         // 2N psi powers in shared memory (32 KB at N = 2048)
    static_assert(T <= 1024, "polynomial too large for two points per thread");
    // u64 words of dynamic shared memory before the switched mask
    static constexpr int WORDS = 5 * M + 2 * NSM * T + (PW_SMEM ? 2 * N : 0);
};

// Thread `tid` swaps ONE of its two values with the thread whose index differs in bit `bit`: the thread with the bit
// clear gives its second value and receives the partner's first, and vice versa (an involution).
template <int BIT>
__device__ __forceinline__ void swap_shfl(u64& r0, u64& r1, int tid) {
    const bool up = (tid >> BIT) & 1;
    const u64 send = up ? r0 : r1;
    const u64 got = __shfl_xor_sync(0xffffffffu, send, 1 << BIT);
    if (up) r0 = got; else r1 = got;
}
template <int BIT>
__device__ __forceinline__ void swap_smem(u64& r0, u64& r1, int tid, u64* xb) {
    const bool up = (tid >> BIT) & 1;
    xb[tid] = up ? r0 : r1;
    __syncthreads();
    const u64 got = xb[tid ^ (1 << BIT)];
    if (up) r0 = got; else r1 = got;
}
template <int BIT, int NSM, int T>
__device__ __forceinline__ void swap_bit(u64& r0, u64& r1, int tid, u64* xb) {
    if constexpr (BIT >= 5) swap_smem<BIT>(r0, r1, tid, xb + (BIT - 5) * T);
    else swap_shfl<BIT>(r0, r1, tid);
}

// Local size-M transform.  In: (r0, r1) = local coefficients (tid, tid + T); out: in-place positions (2 tid, 2 tid + 1).
// twf[s] = tw[2^s + (tid >> (LL-1-s))] (held in registers by the caller).
template <int L, int S = 0>
__device__ __forceinline__ void local_forward(u64& r0, u64& r1, const u64 (&twf)[L - 2], int tid, u64* xb) {
    using C = Split2Cfg<L>;
    butterfly<false>(r0, r1, twf[S]);
    if constexpr (S < C::LL - 1) {
        swap_bit<C::LL - 2 - S, C::NSM, C::T>(r0, r1, tid, xb);
        local_forward<L, S + 1>(r0, r1, twf, tid, xb);
    }
}
// mirror image; out: local coefficients (tid, tid + T), times M
template <int L, int S = L - 3>
__device__ __forceinline__ void local_inverse(u64& r0, u64& r1, const u64 (&twi)[L - 2], int tid, u64* xb) {
    using C = Split2Cfg<L>;
    if constexpr (S < C::LL - 1) swap_bit<C::LL - 2 - S, C::NSM, C::T>(r0, r1, tid, xb);
    butterfly<true>(r0, r1, twi[S]);
    if constexpr (S > 0) local_inverse<L, S - 1>(r0, r1, twi, tid, xb);
}

// key conversion for this kernel's layout: [poly][r][e][t] <- transform value at position 2 (r T + t) + e
template <int L, int E>
__global__ void __launch_bounds__(NttCfg<L, E>::T) bsk_convert_split2_kernel(const u64* __restrict__ src, u64* __restrict__ dst,
                                                                             const u64* __restrict__ tw, u64 ninv) {
    using C = NttCfg<L, E>;
    using S = Split2Cfg<L>;
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
        const int u = P >> 1, e = P & 1;
        d[(u / S::T) * S::M + e * S::T + (u % S::T)] = fmul_c(x[q], ninv);
    }
}

template <int L>
__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(Split2Cfg<L>::T, 1) pbs_split2_kernel(const PbsArgs a) {
    using C = Split2Cfg<L>;
    constexpr int N = C::N, M = C::M, T = C::T, LL = C::LL, Q = M / 4;     // Q blocks of 4 positions per CTA
    extern __shared__ u64 smem[];
    u64* acc = smem;                              // [M] local coefficients (canonical)
    u64* gbuf = smem + M;                         // [M] forward all-to-all landing zone
    u64* lbuf = smem + 2 * M;                     // [M] inverse all-to-all landing zone
    u64* recv = smem + 3 * M;                     // [2][M] partner polynomial's partial sums, by step parity
    u64* xbf = smem + 5 * M;                      // [NSM][T] cross-warp exchange, forward transform
    u64* xbi = xbf + C::NSM * T;                  // [NSM][T] inverse transform
    u64* pws = xbi + C::NSM * T;                  // [2N] powers of psi (PW_SMEM)
    unsigned short* rot = reinterpret_cast<unsigned short*>(smem + C::WORDS);
    __shared__ __align__(8) u64 bars[3];          // fwd, rcv, inv
    const int tid = threadIdx.x;
    const u32 rank = cluster_rank(), c = rank >> 2, r = rank & 3, group0 = c << 2, partner = ((c ^ 1) << 2) + r;
    const int n = a.n, bl = a.bl;
    const int total = a.njobs * a.batch;

    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        mbar_init(&bars[2], 1);
    }
    if constexpr (C::PW_SMEM)
        for (int i = tid; i < 2 * N; i += T) pws[i] = __ldg(a.pw + i);
    const u64* pw = C::PW_SMEM ? pws : a.pw;
    __syncthreads();
    cluster_sync_all();               // every CTA's barriers exist before anyone signals them

    // ---- per-thread constants of the whole kernel
    u64 twf[LL], twi[LL];
#pragma unroll
    for (int s = 0; s < LL; s++) {
        twf[s] = __ldg(a.tw + (1 << s) + (tid >> (LL - 1 - s)));
        twi[s] = __ldg(a.twi + (1 << s) + (tid >> (LL - 1 - s)));
    }
    const int bq = tid >> 1, h = tid & 1;         // block of 4 positions within the CTA, half of it
    const int blk = (int)r * Q + bq;              // block within the polynomial
    const u64 cwa = __ldg(a.tw + (1 << (L - 2)) + blk), cwb = __ldg(a.tw + (1 << (L - 1)) + 2 * blk + h);
    const u64 ciwa = __ldg(a.twi + (1 << (L - 2)) + blk), ciwb = __ldg(a.twi + (1 << (L - 1)) + 2 * blk + h);
    const u32 slot0 = r * M + tid, slot1 = slot0 + T;                  // this thread's two transform slots, [r][e][t]
    const u32 ex0 = __ldg(a.expo + slot0), ex1 = __ldg(a.expo + slot1);
    // forward all-to-all: local position 2 tid + e goes to the CTA owning block (2 tid + e) >> 2 ... of Q per CTA
    const int jl0 = 2 * tid, dstf = jl0 / Q;                            // both positions of a thread share the block
    const u32 gdst = map_shared_u32(gbuf, group0 + dstf) + (u32)(((jl0 % Q) * 4 + (int)r) * 8);
    const u32 fwd_bar_r = map_shared_u32(&bars[0], group0 + dstf);
    const u32 ldst0 = map_shared_u32(lbuf, group0 + h) + (u32)blk * 8, ldst1 = map_shared_u32(lbuf, group0 + h + 2) + (u32)blk * 8;
    const u32 inv_bar_r0 = map_shared_u32(&bars[2], group0 + h), inv_bar_r1 = map_shared_u32(&bars[2], group0 + h + 2);
    const u32 recv_r = map_shared_u32(recv, partner), rcv_bar_r = map_shared_u32(&bars[1], partner);
    u32 ph = 0;                       // common parity of the three barriers (each completes once per step)

    for (int f = blockIdx.x >> 3; f < total; f += gridDim.x >> 3) {
        const int q0 = f / a.batch, b0 = f - q0 * a.batch;
        const u64* in = a.small + ((size_t)a.job_in[q0] * a.batch + b0) * (n + 1);
        const u64* lut = a.luts + (size_t)a.job_lut[q0] * N;
        u64* out = a.out + ((size_t)a.job_out[q0] * a.batch + b0) * (N + 1);

        __syncthreads();              // previous ciphertext's extraction is done with acc / rot
        for (int i = tid; i < n; i += T) rot[i] = (unsigned short)modswitch(in[i], L);
        {
            const u32 r0 = (2 * N - modswitch(in[n], L)) & (2 * N - 1);
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int lc = q * T + tid, gi = lc * 4 + (int)r;
                const u32 u = (gi + 2 * N - r0) & (2 * N - 1);
                acc[lc] = c == 0 ? 0 : (u < N ? lut[u] : fneg(lut[u - N]));
            }
        }
        __syncthreads();

        for (int i = 0; i < n; i += 2) {
            const u32 a1 = rot[i], a2 = rot[i + 1];
            if ((a1 | a2) == 0) continue;
            if (tid == 0) {           // post what this step will receive
                mbar_expect(&bars[0], M * 8);
                mbar_expect(&bars[1], M * 8);
                mbar_expect(&bars[2], M * 8);
            }
            // this step's key words: [pair][K11, KA, KB][row = c][output polynomial][r][e][t]
            const u64* gp = a.bsk_hat + ((size_t)(i >> 1) * 6 + c) * 2 * N + slot0;
            u64 kk[2][3][2];          // [output: 0 = partner's polynomial, 1 = mine][key][e]
#pragma unroll
            for (int o = 0; o < 2; o++)
#pragma unroll
                for (int k = 0; k < 3; k++)
#pragma unroll
                    for (int e = 0; e < 2; e++)
                        kk[o][k][e] = __ldg(gp + (size_t)k * 4 * N + (size_t)(o ? c : (c ^ 1)) * N + e * T);

            u64 x0 = digit_of(round_top(acc[tid], bl), bl, 1, 1), x1 = digit_of(round_top(acc[tid + T], bl), bl, 1, 1);
            // ---- forward: local stages, all-to-all, two block stages
            local_forward<L>(x0, x1, twf, tid, xbf);
            st_async_u64(gdst, x0, fwd_bar_r);
            st_async_u64(gdst + 32, x1, fwd_bar_r);            // position 2 tid + 1: next block row of 4 words
            mbar_wait(&bars[0], ph);
            x0 = gbuf[bq * 4 + h];
            x1 = gbuf[bq * 4 + h + 2];
            butterfly<false>(x0, x1, cwa);
            swap_shfl<0>(x0, x1, tid);
            butterfly<false>(x0, x1, cwb);
            // ---- pointwise: combined key of the step at my two slots, for both output polynomials
            u64 own0, own1;
            {
                const u32 rbase = recv_r + (ph ? (u32)M * 8 : 0);
                const u32 mask = 2 * N - 1;
                const u64 m10a = fsub_l(pw[(a1 * ex0) & mask], 1), m01a = fsub_l(pw[(a2 * ex0) & mask], 1);
                const u64 m10b = fsub_l(pw[(a1 * ex1) & mask], 1), m01b = fsub_l(pw[(a2 * ex1) & mask], 1);
                const u64 ko0 = fadd_l(fmul_l(m10a, fadd_l(fmul_l(m01a, kk[0][0][0]), kk[0][1][0])), fmul_c(m01a, kk[0][2][0]));
                const u64 ko1 = fadd_l(fmul_l(m10b, fadd_l(fmul_l(m01b, kk[0][0][1]), kk[0][1][1])), fmul_c(m01b, kk[0][2][1]));
                st_async_u64(rbase + (u32)tid * 8, fmul_c(x0, ko0), rcv_bar_r);
                st_async_u64(rbase + (u32)(tid + T) * 8, fmul_c(x1, ko1), rcv_bar_r);
                const u64 km0 = fadd_l(fmul_l(m10a, fadd_l(fmul_l(m01a, kk[1][0][0]), kk[1][1][0])), fmul_c(m01a, kk[1][2][0]));
                const u64 km1 = fadd_l(fmul_l(m10b, fadd_l(fmul_l(m01b, kk[1][0][1]), kk[1][1][1])), fmul_c(m01b, kk[1][2][1]));
                own0 = fmul_l(x0, km0);
                own1 = fmul_l(x1, km1);
            }
            mbar_wait(&bars[1], ph);
            {
                const u64* rb = recv + (ph ? M : 0);
                own0 = fadd_l(own0, rb[tid]);
                own1 = fadd_l(own1, rb[tid + T]);
            }
            // ---- inverse: two block stages, all-to-all, local stages
            butterfly<true>(own0, own1, ciwb);
            swap_shfl<0>(own0, own1, tid);
            butterfly<true>(own0, own1, ciwa);
            st_async_u64(ldst0, own0, inv_bar_r0);
            st_async_u64(ldst1, own1, inv_bar_r1);
            mbar_wait(&bars[2], ph);
            ph ^= 1;
            own0 = lbuf[2 * tid];
            own1 = lbuf[2 * tid + 1];
            local_inverse<L>(own0, own1, twi, tid, xbi);
            acc[tid] = fcanon(fadd_l(own0, acc[tid]));
            acc[tid + T] = fcanon(fadd_l(own1, acc[tid + T]));
        }

        // sample extraction: out[0] = A[0], out[t] = -A[N - t]; coefficient gi of the mask polynomial goes to out[(N - gi) % N]
        if (c == 0) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int lc = q * T + tid, gi = lc * 4 + (int)r;
                const u64 v = acc[lc];
                if (gi == 0) out[0] = v; else out[N - gi] = fneg(v);
            }
        } else if (r == 0 && tid == 0) {
            out[N] = acc[0];
        }
    }
    cluster_sync_all();               // nobody exits while its shared memory may still be written
}
