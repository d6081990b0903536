// Device kernels of the TFHE hot path (sm_100a).
#pragma once
#include "ntt.cuh"

struct PbsArgs {
    const u64* __restrict__ bsk_hat;   // [n][2l][2][2^E][T]  transform domain, x 1/N, in the layout of the kernel's E
    const u64* __restrict__ tw;        // [N] psi^brv(i)
    const u64* __restrict__ twi;       // [N] psi^-brv(i)
    const u64* __restrict__ luts;      // [n_luts][N]
    const u64* __restrict__ small;     // [rows][n+1] keyswitched inputs
    const int* __restrict__ job_in;    // [njobs] row (before batch expansion) into small
    const int* __restrict__ job_lut;   // [njobs]
    const int* __restrict__ job_out;   // [njobs] row (before batch expansion) into out
    u64* __restrict__ out;             // [rows][N+1] big-key LWE
    int njobs, batch;
    int n, bl, l;
    // pair blind rotation (two key bits per step): bsk_hat is then the pair key [n/2][K11, K11+K10, K11+K01][2][2][N]
    const u32* __restrict__ expo;      // [N] transform slot (in bsk_hat's layout) -> r with evaluation point psi^r
    const u64* __restrict__ pw;        // [2N] psi^t
};

// Pair blind rotation, transform domain: the GGSW the step multiplies the decomposed accumulator with is
//   (X^(a1+a2) - 1) K11 + (X^a1 - 1) K10 + (X^a2 - 1) K01  =  m10 (m01 K11 + KA) + m01 KB,
// m10 = X^a1 - 1, m01 = X^a2 - 1, KA = K11 + K10, KB = K11 + K01 (the sums are formed once when the key is loaded,
// pair_key_sums_kernel), and a monomial X^e is, at the slot whose evaluation point is psi^r, the scalar
// psi^(e r mod 2N).  Three products per slot and output polynomial.
struct PairMono {
    u64 m10, m01;      // lazy
};
__device__ __forceinline__ PairMono pair_monomials(const PbsArgs& a, u32 slot, u32 a1, u32 a2, u32 mask2n) {
    const u32 e = __ldg(a.expo + slot);
    PairMono m;
    m.m10 = fsub_l(__ldg(a.pw + ((a1 * e) & mask2n)), 1);
    m.m01 = fsub_l(__ldg(a.pw + ((a2 * e) & mask2n)), 1);
    return m;
}
// m10 (m01 k11 + ka) + m01 kb, lazy; the key words are canonical
__device__ __forceinline__ u64 pair_combine(const PairMono& m, u64 k11, u64 ka, u64 kb) {
    return fadd_l(fmul_l(m.m10, fadd_l(fmul_l(m.m01, k11), ka)), fmul_c(m.m01, kb));
}

// The two builds of the bootstrap kernel per polynomial size: coefficients per thread (log2) and the CTAs per SM
// the register budget is sized for.  Latency build: as many warps per transform as a CTA allows, all in registers.
// Throughput build: the configuration that bootstraps the most ciphertexts per second with the GPU full
// (measured, scripts/pbs_sweep.py).
#ifndef BMI_LAT_E11
#define BMI_LAT_E11 2
#define BMI_LAT_E12 3
#endif
template <int L>
constexpr int latency_e() { return L <= 11 ? BMI_LAT_E11 : L == 12 ? BMI_LAT_E12 : 3; }
#ifndef BMI_TP_E11
#define BMI_TP_E11 3
#define BMI_TP_B11 4
#define BMI_TP_E12 3
#define BMI_TP_B12 2
#define BMI_TP_E13 3
#define BMI_TP_B13 1
#endif
template <int L>
constexpr int throughput_e() { return L <= 11 ? BMI_TP_E11 : L == 12 ? BMI_TP_E12 : BMI_TP_E13; }
template <int L>
constexpr int throughput_ctas_per_sm() { return L <= 11 ? BMI_TP_B11 : L == 12 ? BMI_TP_B12 : BMI_TP_B13; }

// ----------------------------------------------------------------------------
// bootstrapping-key conversion: standard-domain polynomial -> transform domain in the layout of build E
template <int L, int E>
__global__ void __launch_bounds__(NttCfg<L, E>::T) bsk_convert_kernel(const u64* __restrict__ src, u64* __restrict__ dst,
                                                                      const u64* __restrict__ tw, u64 ninv) {
    using C = NttCfg<L, E>;
    extern __shared__ u64 smem[];
    const int tid = threadIdx.x;
    const u64* s = src + (size_t)blockIdx.x * C::N;
    u64* d = dst + (size_t)blockIdx.x * C::N;
    u64 x[C::EPT];
#pragma unroll
    for (int q = 0; q < C::EPT; q++) x[q] = s[q * C::T + tid];
    ntt_forward<L, E>(x, smem, tw, tid);
#pragma unroll
    for (int q = 0; q < C::EPT; q++) d[q * C::T + tid] = fmul_c(x[q], ninv);
}

// c = a * b mod (X^N + 1): exercises forward, pointwise and inverse transforms (self-test entry)
template <int L, int E>
__global__ void __launch_bounds__(NttCfg<L, E>::T) polymul_kernel(const u64* __restrict__ a, const u64* __restrict__ b,
                                                                  u64* __restrict__ c, const u64* __restrict__ tw,
                                                                  const u64* __restrict__ twi, u64 ninv) {
    using C = NttCfg<L, E>;
    extern __shared__ u64 smem[];
    const int tid = threadIdx.x;
    const size_t off = (size_t)blockIdx.x * C::N;
    u64 x[C::EPT], y[C::EPT];
#pragma unroll
    for (int q = 0; q < C::EPT; q++) { x[q] = a[off + q * C::T + tid]; y[q] = b[off + q * C::T + tid]; }
    ntt_forward<L, E>(x, smem, tw, tid);
    ntt_forward<L, E>(y, smem, tw, tid);
#pragma unroll
    for (int q = 0; q < C::EPT; q++) x[q] = fmul_l(fmul_l(x[q], y[q]), ninv);
    ntt_inverse<L, E>(x, smem, twi, tid);
#pragma unroll
    for (int q = 0; q < C::EPT; q++) c[off + q * C::T + tid] = fcanon(x[q]);
}

// ----------------------------------------------------------------------------
// programmable bootstrap (modulus switch -> blind rotation -> sample extraction) on a 2-CTA cluster:
// CTA c owns accumulator polynomial c (0 = mask, 1 = body).  Per CMUX each CTA decomposes and forward-transforms
// only ITS polynomial, multiplies it by both columns of its GGSW rows, keeps the partial sum for its own output
// polynomial and pushes the partial sum for the partner's polynomial straight into the partner's shared memory
// (DSMEM).  After the cluster barrier each CTA adds what it received, inverse-transforms one polynomial and
// updates its accumulator.
//   E         coefficients per thread (log2): 4 = throughput build, 2/3 = latency build (more warps per transform)
//   MINB      CTAs per SM the register budget is sized for
//   ONE_LEVEL l == 1 (every 128-bit parameter set this engine selects): the partner's partial sum is a single
//             product per coefficient and goes straight to DSMEM instead of being accumulated in registers
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() { cluster_arrive(); cluster_wait(); }
__device__ __forceinline__ u32 cluster_rank() { u32 r; asm("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
// address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ u64* cluster_map(u64* p, u32 rank) {
    u64 out;
    asm("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"((u64)p), "r"(rank));
    return reinterpret_cast<u64*>(out);
}

// ---- TMA bulk copy global -> shared with mbarrier completion (SASS: UBLKCP / SYNCS)
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// (A suspend-time hint on try_wait was measured: no change in either bootstrap kernel, so the plain form stays.)
// bulk copy counted on `bar` without an arrival of its own (several copies behind one mbar_expect)
__device__ __forceinline__ void tma_copy_1d(void* dst, const void* src, u32 bytes, u64* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, u32 parity) {
    asm volatile("{ .reg .pred p;\n"
                 "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra D;\n bra W;\n D: }" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

// ----------------------------------------------------------------------------
// DSMEM signalling without cluster-wide barriers: data is pushed with st.async, which counts its bytes on an
// mbarrier in the DESTINATION CTA; the consumer waits on its own mbarrier for the bytes it expects.  No release
// fence on the producer side (the cluster-barrier version spent 29 % of its issue slots in `membar`, see profiles/).
__device__ __forceinline__ u32 map_shared_u32(const void* p, u32 rank) {
    u32 out;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(out) : "r"(smem_addr(p)), "r"(rank));
    return out;
}
__device__ __forceinline__ void st_async_u64(u32 remote_addr, u64 v, u32 remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.u64 [%0], %1, [%2];"
                 ::"r"(remote_addr), "l"(v), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_expect(u64* bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(u32 remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(u64* bar, u32 parity) {
    asm volatile("{ .reg .pred p;\n"
                 "WC: mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DC;\n bra WC;\n DC: }" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}

//   STAGE     (ONE_LEVEL only) the GGSW row polynomials of the next step are brought into shared memory by TMA bulk
//             copies issued a whole step ahead (2N words; with PAIR the three keys of the pair, 6N words), instead of
//             per-thread global loads after the transform.  An A/B option (bmi_ctx_set_tma_stage), off by default.
//   PAIR      (ONE_LEVEL) two key bits per step with the pair key: the accumulator itself is decomposed (no
//             rotated reads), one forward/inverse transform serves both bits, the monomials enter the pointwise stage
template <int L, int E, int MINB, bool ONE_LEVEL, bool STAGE, bool PAIR = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NttCfg<L, E>::T, MINB) pbs_cluster_kernel(const PbsArgs a) {
    static_assert(!PAIR || ONE_LEVEL, "pair blind rotation: one decomposition level");
    constexpr int STAGE_WORDS = STAGE ? (PAIR ? 6 : 2) * NttCfg<L, E>::N : 0;
    using C = NttCfg<L, E>;
    constexpr int N = C::N, T = C::T, EPT = C::EPT;
    extern __shared__ u64 smem[];
    u64* acc = smem;                  // [N] this CTA's accumulator polynomial (canonical values)
    u64* buf = smem + N;              // [N] transform exchange buffer (swizzled)
    u64* recv = smem + 2 * N;         // [N] partial sums pushed by the partner CTA
    u64* stage = smem + 3 * N;        // [2N | 6N] GGSW rows of the coming step (STAGE only)
    unsigned short* rot = reinterpret_cast<unsigned short*>(smem + 3 * N + STAGE_WORDS);   // [n] switched mask
    __shared__ __align__(8) u64 stage_bar;
    const int tid = threadIdx.x;
    const u32 me = cluster_rank(), other = me ^ 1;
    const int n = a.n, bl = a.bl, l = a.l, tot = bl * l;
    const int total = a.njobs * a.batch;
    u32 stage_phase = 0;
    // partner exchange by mbarriers: rcv_bar counts the bytes the partner pushes with st.async, free_bar is the
    // partner's "I have consumed your previous push" (no cluster-wide barrier, no release fence on the hot path)
    __shared__ __align__(8) u64 xbar[2];
    if (tid == 0) {
        if (STAGE) mbar_init(&stage_bar, 1);
        mbar_init(&xbar[0], 1);
        mbar_init(&xbar[1], 1);
    }
    __syncthreads();
    cluster_sync_all();
    const u32 peer_recv_a = map_shared_u32(recv, other), peer_rcv_bar = map_shared_u32(&xbar[0], other),
              peer_free_bar = map_shared_u32(&xbar[1], other);
    u32 ph_x = 0;
    if (tid == 0) mbar_arrive_remote(peer_free_bar);      // the partner's first push needs no waiting
    // thread 0: start the bulk copy of the rows of the first CMUX at or after `from` that is not skipped
    auto prefetch_rows = [&](int from) {
        if (PAIR) {
            for (int i = from; i < n; i += 2)
                if ((rot[i] | rot[i + 1]) != 0) {     // [pair][K11, KA, KB][row = me][both output polynomials][N]
                    mbar_expect(&stage_bar, 6 * N * 8);
#pragma unroll
                    for (int k = 0; k < 3; k++)
                        tma_copy_1d(stage + k * 2 * N, a.bsk_hat + ((size_t)(i >> 1) * 6 + k * 2 + me) * 2 * N, 2 * N * 8, &stage_bar);
                    return;
                }
            return;
        }
        for (int i = from; i < n; i++)
            if (rot[i] != 0) {
                tma_load_1d(stage, a.bsk_hat + ((size_t)i * 2 + me) * 2 * N, 2 * N * 8, &stage_bar);
                return;
            }
    };

    for (int f = blockIdx.x >> 1; f < total; f += gridDim.x >> 1) {
        const int q0 = f / a.batch, b0 = f - q0 * a.batch;
        const u64* in = a.small + ((size_t)a.job_in[q0] * a.batch + b0) * (n + 1);
        const u64* lut = a.luts + (size_t)a.job_lut[q0] * N;
        u64* out = a.out + ((size_t)a.job_out[q0] * a.batch + b0) * (N + 1);

        __syncthreads();              // previous ciphertext fully written out
        for (int i = tid; i < n; i += T) rot[i] = (unsigned short)modswitch(in[i], L);
        {   // acc = (0, X^{-b~} * lut)
            const u32 r0 = (2 * N - modswitch(in[n], L)) & (2 * N - 1);
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                const int idx = q * T + tid;
                const u32 u = (idx + 2 * N - r0) & (2 * N - 1);
                acc[idx] = me == 0 ? 0 : (u < N ? lut[u] : fneg(lut[u - N]));
            }
        }
        __syncthreads();
        if (STAGE && tid == 0) prefetch_rows(0);

        for (int i = 0; i < n; i += PAIR ? 2 : 1) {
            const u32 at = rot[i], at2 = PAIR ? rot[i + 1] : 0;
            if ((at | at2) == 0) continue;    // X^0 - 1 = 0: nothing to add (same decision in both CTAs)
            if (tid == 0) mbar_expect(&xbar[0], N * 8);
            u64 own[EPT];
            const u64* g = a.bsk_hat + ((size_t)i * (2 * l) + me * l) * 2 * N;
            if (PAIR) {
                u64 x[EPT];
#pragma unroll
                for (int q = 0; q < EPT; q++) x[q] = digit_of(round_top(acc[q * T + tid], tot), bl, 1, 1);
                ntt_forward<L, E>(x, buf, a.tw, tid);
                mbar_wait_cluster(&xbar[1], ph_x);     // partner has consumed what I pushed for the previous step
                // pair key: [pair][K11, KA, KB][row = decomposed polynomial][output polynomial][N]
                const u64* gp = a.bsk_hat + ((size_t)(i >> 1) * 6 + me) * 2 * N;
                if (STAGE) {
                    mbar_wait(&stage_bar, stage_phase);
                    stage_phase ^= 1;
                }
#pragma unroll
                for (int q = 0; q < EPT; q++) {
                    const int idx = q * T + tid;
                    const PairMono m = pair_monomials(a, idx, at, at2, 2 * N - 1);
                    u64 ko, km;
                    if (STAGE) {      // stage: [key][output polynomial][N]
                        ko = pair_combine(m, stage[other * N + idx], stage[2 * N + other * N + idx], stage[4 * N + other * N + idx]);
                        km = pair_combine(m, stage[me * N + idx], stage[2 * N + me * N + idx], stage[4 * N + me * N + idx]);
                    } else {
                        ko = pair_combine(m, __ldg(gp + other * N + idx), __ldg(gp + 4 * N + other * N + idx), __ldg(gp + 8 * N + other * N + idx));
                        km = pair_combine(m, __ldg(gp + me * N + idx), __ldg(gp + 4 * N + me * N + idx), __ldg(gp + 8 * N + me * N + idx));
                    }
                    st_async_u64(peer_recv_a + (u32)idx * 8, fmul_c(x[q], ko), peer_rcv_bar);
                    own[q] = fmul_l(x[q], km);
                }
            } else if (ONE_LEVEL) {
                u64 x[EPT];
#pragma unroll
                for (int q = 0; q < EPT; q++) {   // the digit of (X^at - 1) * acc
                    const int idx = q * T + tid;
                    const u32 u = (idx + 2 * N - at) & (2 * N - 1);
                    const u64 r = u < N ? acc[u] : fneg(acc[u - N]);
                    x[q] = digit_of(round_top(fsub(r, acc[idx]), tot), bl, 1, 1);
                }
                ntt_forward<L, E>(x, buf, a.tw, tid);
                mbar_wait_cluster(&xbar[1], ph_x);     // partner has consumed what I pushed for the previous CMUX
                if (STAGE) {
                    mbar_wait(&stage_bar, stage_phase);
                    stage_phase ^= 1;
#pragma unroll
                    for (int q = 0; q < EPT; q++) {
                        st_async_u64(peer_recv_a + (u32)(q * T + tid) * 8, fmul_c(x[q], stage[other * N + q * T + tid]), peer_rcv_bar);
                        own[q] = fmul_l(x[q], stage[me * N + q * T + tid]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < EPT; q++) {
                        st_async_u64(peer_recv_a + (u32)(q * T + tid) * 8, fmul_c(x[q], __ldg(g + other * N + q * T + tid)), peer_rcv_bar);
                        own[q] = fmul_l(x[q], __ldg(g + me * N + q * T + tid));
                    }
                }
            } else {
                u64 oth[EPT];
#pragma unroll
                for (int q = 0; q < EPT; q++) { own[q] = 0; oth[q] = 0; }
                for (int j = 1; j <= l; j++) {
                    u64 x[EPT];
#pragma unroll
                    for (int q = 0; q < EPT; q++) {
                        const int idx = q * T + tid;
                        const u32 u = (idx + 2 * N - at) & (2 * N - 1);
                        const u64 r = u < N ? acc[u] : fneg(acc[u - N]);
                        x[q] = digit_of(round_top(fsub(r, acc[idx]), tot), bl, l, j);
                    }
                    ntt_forward<L, E>(x, buf, a.tw, tid);
                    const u64* row = g + (size_t)(j - 1) * 2 * N;
#pragma unroll
                    for (int q = 0; q < EPT; q++) {
                        own[q] = fadd_l(own[q], fmul_c(x[q], __ldg(row + me * N + q * T + tid)));
                        oth[q] = fadd_l(oth[q], fmul_c(x[q], __ldg(row + other * N + q * T + tid)));
                    }
                }
                mbar_wait_cluster(&xbar[1], ph_x);     // partner has consumed what I pushed for the previous CMUX
#pragma unroll
                for (int q = 0; q < EPT; q++) st_async_u64(peer_recv_a + (u32)(q * T + tid) * 8, fcanon(oth[q]), peer_rcv_bar);
            }
            mbar_wait(&xbar[0], ph_x);                 // the partner's partial sums have landed
            ph_x ^= 1;
#pragma unroll
            for (int q = 0; q < EPT; q++) own[q] = fadd_l(own[q], recv[q * T + tid]);
            ntt_inverse<L, E>(own, buf, a.twi, tid);   // (its first __syncthreads orders every thread's reads of recv / stage)
            if (tid == 0) mbar_arrive_remote(peer_free_bar);
            if (STAGE && tid == 0) prefetch_rows(i + (PAIR ? 2 : 1));
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                const int idx = q * T + tid;
                acc[idx] = fcanon(fadd_l(own[q], acc[idx]));
            }
            __syncthreads();
        }

        // sample extraction of the constant coefficient
        if (me == 0) {
#pragma unroll
            for (int q = 0; q < EPT; q++) {
                const int t = q * T + tid;
                out[t] = t == 0 ? acc[0] : fneg(acc[N - t]);
            }
        } else if (tid == 0) {
            out[N] = acc[0];
        }
    }
    cluster_sync_all();               // nobody exits while the partner may still push into its shared memory
}
