/*
 * tfhe_oracle_fast.c -- CPU ORACLE, optimised leg (test infrastructure, NOT product code).
 *
 * Same functions as tfhe_oracle.c (keyswitch, PBS) restated the way a tuned CPU
 * backend runs them (what concrete-cpu / tfhe-rs do for the reference's
 * circuit.run(), qfloat_matrix_inversion.py:1034): bootstrapping key held in the
 * transform domain, one forward transform per decomposed polynomial shared by all
 * output polynomials, merged-twiddle in-place butterflies, jobs fanned out over
 * host threads.  tests/test_oracle.py pins every function here bit-for-bit
 * against the definitional versions in tfhe_oracle.c.  bench.py times THIS file
 * as `cpu_baseline` (kind "port") and as `--impl reference`.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned __int128 u128;
#define P 0xFFFFFFFF00000001ULL
#define EPS 0xFFFFFFFFULL

typedef struct {
    int n, k, N, bsk_bl, bsk_l, ksk_bl, ksk_l;
    double lwe_sigma, glwe_sigma;
} params_t;

static inline u64 fadd(u64 a, u64 b) { u64 s = a + b; if (s < a || s >= P) s -= P; return s; }
static inline u64 fsub(u64 a, u64 b) { return a >= b ? a - b : a + (P - b); }
static inline u64 fneg(u64 a) { return a ? P - a : 0; }
static inline u64 fmul(u64 a, u64 b) {
    u128 x = (u128)a * b;
    u64 lo = (u64)x, hi = (u64)(x >> 64), hh = hi >> 32, hl = hi & EPS;
    u64 t0 = lo - hh; if (lo < hh) t0 -= EPS;
    u64 t1 = hl * EPS;
    u64 t2 = t0 + t1; if (t2 < t1) t2 += EPS;
    return t2 >= P ? t2 - P : t2;
}
static u64 fpow(u64 b, u64 e) { u64 r = 1; while (e) { if (e & 1) r = fmul(r, b); b = fmul(b, b); e >>= 1; } return r; }
static int ilog2(u64 n) { int l = 0; while ((1ULL << l) < n) l++; return l; }
static int brv(int x, int L) { int r = 0; for (int b = 0; b < L; b++) if (x >> b & 1) r |= 1 << (L - 1 - b); return r; }

typedef struct {
    params_t pp;
    int logN;
    u64 *tw, *twi;      /* psi^{brv(i)}, psi^{-brv(i)} */
    u64 *bsk_hat;       /* [n][rows][k+1][N], transform domain, pre-scaled by 1/N; pair key: [n/2][3][rows][k+1][N] */
    const u64 *ksk;     /* borrowed */
    int pairs;          /* blind rotation two key bits per step with the pair key (tfhe_oracle.c orc_pbs_pairs) */
    u64 *pw;            /* [2N] psi^t */
    int *expo;          /* [N] transform slot u evaluates at psi^expo[u] */
} ctx_t;

static void fwd(const ctx_t *c, u64 *a) {
    int N = c->pp.N, t = N;
    for (int m = 1; m < N; m <<= 1) {
        t >>= 1;
        for (int i = 0; i < m; i++) {
            u64 w = c->tw[m + i]; u64 *x = a + 2 * i * t, *y = x + t;
            for (int j = 0; j < t; j++) { u64 u = x[j], v = fmul(y[j], w); x[j] = fadd(u, v); y[j] = fsub(u, v); }
        }
    }
}
static void inv(const ctx_t *c, u64 *a) {   /* unscaled: the 1/N lives in bsk_hat */
    int N = c->pp.N, t = 1;
    for (int m = N >> 1; m >= 1; m >>= 1) {
        for (int i = 0; i < m; i++) {
            u64 w = c->twi[m + i]; u64 *x = a + 2 * i * t, *y = x + t;
            for (int j = 0; j < t; j++) { u64 u = x[j], v = y[j]; x[j] = fadd(u, v); y[j] = fmul(fsub(u, v), w); }
        }
        t <<= 1;
    }
}

static void *create(const params_t *pp, const u64 *bsk, const u64 *ksk, int pairs) {
    ctx_t *c = calloc(1, sizeof(ctx_t));
    c->pp = *pp; c->logN = ilog2(pp->N); c->ksk = ksk; c->pairs = pairs;
    int N = pp->N, L = c->logN;
    u64 psi = fpow(7, (P - 1) / (2 * (u64)N)), psii = fpow(psi, P - 2);
    c->tw = malloc(sizeof(u64) * N); c->twi = malloc(sizeof(u64) * N);
    for (int i = 0; i < N; i++) { c->tw[i] = fpow(psi, brv(i, L)); c->twi[i] = fpow(psii, brv(i, L)); }
    if (pairs) {
        /* evaluation point of every transform slot = transform of the monomial X; its discrete log by table */
        c->pw = malloc(sizeof(u64) * 2 * N); c->expo = malloc(sizeof(int) * N);
        c->pw[0] = 1;
        for (int t = 1; t < 2 * N; t++) c->pw[t] = fmul(c->pw[t - 1], psi);
        u64 *x = calloc(N, sizeof(u64)); x[1] = 1;
        fwd(c, x);
        for (int u = 0; u < N; u++) {
            c->expo[u] = -1;
            for (int t = 1; t < 2 * N; t += 2) if (c->pw[t] == x[u]) { c->expo[u] = t; break; }
        }
        free(x);
    }
    u64 polys = (u64)(pairs ? 3 * (pp->n / 2) : pp->n) * (pp->k + 1) * pp->bsk_l * (pp->k + 1);
    c->bsk_hat = malloc(sizeof(u64) * polys * N);
    u64 ninv = fpow((u64)N, P - 2);
    for (u64 q = 0; q < polys; q++) {
        u64 *d = c->bsk_hat + q * N; memcpy(d, bsk + q * N, sizeof(u64) * N);
        fwd(c, d);
        for (int t = 0; t < N; t++) d[t] = fmul(d[t], ninv);
    }
    return c;
}
void *orcf_create(const params_t *pp, const u64 *bsk, const u64 *ksk) { return create(pp, bsk, ksk, 0); }
void *orcf_create_pairs(const params_t *pp, const u64 *bskp, const u64 *ksk) { return create(pp, bskp, ksk, 1); }
void orcf_destroy(void *h) { ctx_t *c = h; free(c->tw); free(c->twi); free(c->bsk_hat); free(c->pw); free(c->expo); free(c); }

static inline u64 modswitch(u64 x, int logN) { return (((x >> (62 - logN)) + 1) >> 1) & ((2ULL << logN) - 1); }

/* pair blind rotation: per step the accumulator itself is decomposed and transformed once; the three GGSWs of the
 * pair enter as  m11 K11 + m10 K10 + m01 K01  with the monomials X^e - 1 evaluated per slot (psi^(e r) - 1) */
static void pbs_pairs(const ctx_t *c, const u64 *lut, const u64 *in, u64 *out) {
    const params_t *pp = &c->pp;
    int n = pp->n, k = pp->k, N = pp->N, l = pp->bsk_l, bl = pp->bsk_bl, tot = bl * l;
    u64 sz = (u64)(k + 1) * N, B = 1ULL << bl, twoN = 2 * (u64)N, ggsw = (u64)(k + 1) * l * sz;
    u64 *acc = calloc(sz, sizeof(u64)), *rnd = malloc(sizeof(u64) * sz);
    u64 *dig = malloc(sizeof(u64) * N), *sum = malloc(sizeof(u64) * sz), *m = malloc(sizeof(u64) * 3 * N);
    u64 rb = (twoN - modswitch(in[n], c->logN)) & (twoN - 1);
    for (int t = 0; t < N; t++) { u64 u = ((u64)t + twoN - rb) & (twoN - 1); acc[(u64)k * N + t] = u < (u64)N ? lut[u] : fneg(lut[u - N]); }
    for (int q = 0; q < n / 2; q++) {
        u64 a1 = modswitch(in[2 * q], c->logN), a2 = modswitch(in[2 * q + 1], c->logN);
        if (!a1 && !a2) continue;
        for (int t = 0; t < N; t++) {
            u64 p10 = c->pw[(a1 * (u64)c->expo[t]) & (twoN - 1)], p01 = c->pw[(a2 * (u64)c->expo[t]) & (twoN - 1)];
            m[t] = fsub(fmul(p10, p01), 1); m[N + t] = fsub(p10, 1); m[2 * N + t] = fsub(p01, 1);
        }
        for (u64 t = 0; t < sz; t++) { u64 r = ((acc[t] >> (63 - tot)) + 1) >> 1; if (tot < 64) r &= (1ULL << tot) - 1; rnd[t] = r; }
        memset(sum, 0, sizeof(u64) * sz);
        const u64 *g = c->bsk_hat + (u64)q * 3 * ggsw;
        for (int j = l; j >= 1; j--) for (int cc = 0; cc <= k; cc++) {
            u64 *r = rnd + (u64)cc * N;
            for (int t = 0; t < N; t++) {
                u64 d = r[t] & (B - 1); r[t] >>= bl;
                if (d >= B / 2) { dig[t] = P - (B - d); r[t] += 1; } else dig[t] = d;
            }
            fwd(c, dig);
            u64 row = (u64)(cc * l + (j - 1)) * sz;
            for (int o = 0; o <= k; o++) {
                u64 *s = sum + (u64)o * N;
                const u64 *b11 = g + row + (u64)o * N, *b10 = b11 + ggsw, *b01 = b10 + ggsw;
                for (int t = 0; t < N; t++) {
                    u64 key = fadd(fadd(fmul(m[t], b11[t]), fmul(m[N + t], b10[t])), fmul(m[2 * N + t], b01[t]));
                    s[t] = fadd(s[t], fmul(dig[t], key));
                }
            }
        }
        for (int o = 0; o <= k; o++) {
            inv(c, sum + (u64)o * N);
            for (int t = 0; t < N; t++) acc[(u64)o * N + t] = fadd(acc[(u64)o * N + t], sum[(u64)o * N + t]);
        }
    }
    for (int cc = 0; cc < k; cc++) {
        const u64 *A = acc + (u64)cc * N; u64 *o = out + (u64)cc * N;
        o[0] = A[0];
        for (int t = 1; t < N; t++) o[t] = fneg(A[N - t]);
    }
    out[(u64)k * N] = acc[(u64)k * N];
    free(acc); free(rnd); free(dig); free(sum); free(m);
}

void orcf_pbs(void *h, const u64 *lut, const u64 *in, u64 *out) {
    const ctx_t *c = h; const params_t *pp = &c->pp;
    if (c->pairs) { pbs_pairs(c, lut, in, out); return; }
    int n = pp->n, k = pp->k, N = pp->N, l = pp->bsk_l, bl = pp->bsk_bl, tot = bl * l;
    u64 sz = (u64)(k + 1) * N, B = 1ULL << bl, twoN = 2 * (u64)N;
    u64 *acc = calloc(sz, sizeof(u64)), *rnd = malloc(sizeof(u64) * sz);
    u64 *dig = malloc(sizeof(u64) * N), *sum = malloc(sizeof(u64) * sz);
    u64 rb = (twoN - modswitch(in[n], c->logN)) & (twoN - 1);
    for (int t = 0; t < N; t++) { u64 u = ((u64)t + twoN - rb) & (twoN - 1); acc[(u64)k * N + t] = u < (u64)N ? lut[u] : fneg(lut[u - N]); }
    for (int i = 0; i < n; i++) {
        u64 at = modswitch(in[i], c->logN);
        if (!at) continue;
        /* rounded (X^at - 1) * acc, kept as tot-bit integers */
        for (int cc = 0; cc <= k; cc++) for (int t = 0; t < N; t++) {
            const u64 *A = acc + (u64)cc * N;
            u64 u = ((u64)t + twoN - at) & (twoN - 1);
            u64 d = fsub(u < (u64)N ? A[u] : fneg(A[u - N]), A[t]);
            u64 r = ((d >> (63 - tot)) + 1) >> 1; if (tot < 64) r &= (1ULL << tot) - 1;
            rnd[(u64)cc * N + t] = r;
        }
        memset(sum, 0, sizeof(u64) * sz);
        const u64 *g = c->bsk_hat + (u64)i * (k + 1) * l * sz;
        for (int j = l; j >= 1; j--) for (int cc = 0; cc <= k; cc++) {
            u64 *r = rnd + (u64)cc * N;
            for (int t = 0; t < N; t++) {
                u64 d = r[t] & (B - 1); r[t] >>= bl;
                if (d >= B / 2) { dig[t] = P - (B - d); r[t] += 1; } else dig[t] = d;
            }
            fwd(c, dig);
            const u64 *row = g + (u64)(cc * l + (j - 1)) * sz;
            for (int o = 0; o <= k; o++) {
                u64 *s = sum + (u64)o * N; const u64 *b = row + (u64)o * N;
                for (int t = 0; t < N; t++) s[t] = fadd(s[t], fmul(dig[t], b[t]));
            }
        }
        for (int o = 0; o <= k; o++) {
            inv(c, sum + (u64)o * N);
            for (int t = 0; t < N; t++) acc[(u64)o * N + t] = fadd(acc[(u64)o * N + t], sum[(u64)o * N + t]);
        }
    }
    for (int cc = 0; cc < k; cc++) {
        const u64 *A = acc + (u64)cc * N; u64 *o = out + (u64)cc * N;
        o[0] = A[0];
        for (int t = 1; t < N; t++) o[t] = fneg(A[N - t]);
    }
    out[(u64)k * N] = acc[(u64)k * N];
    free(acc); free(rnd); free(dig); free(sum);
}

void orcf_keyswitch(void *h, const u64 *in, u64 *out) {
    const ctx_t *c = h; const params_t *pp = &c->pp;
    int n = pp->n, kN = pp->k * pp->N, l = pp->ksk_l, bl = pp->ksk_bl, tot = bl * l;
    u64 B = 1ULL << bl;
    /* 128-bit lazy accumulators: positive and negative digit products kept apart */
    u128 *pos = calloc(n + 1, sizeof(u128)), *neg = calloc(n + 1, sizeof(u128));
    for (int i = 0; i < kN; i++) {
        u64 r = ((in[i] >> (63 - tot)) + 1) >> 1; if (tot < 64) r &= (1ULL << tot) - 1;
        for (int j = l; j >= 1; j--) {
            u64 d = r & (B - 1); r >>= bl;
            const u64 *ct = c->ksk + ((u64)i * l + (j - 1)) * (n + 1);
            if (d >= B / 2) { u64 a = B - d; r += 1; for (int t = 0; t <= n; t++) neg[t] += (u128)a * ct[t]; }
            else if (d) { for (int t = 0; t <= n; t++) pos[t] += (u128)d * ct[t]; }
        }
    }
    for (int t = 0; t <= n; t++) {
        u64 pv = (u64)(pos[t] % P), nv = (u64)(neg[t] % P);
        u64 base = t == n ? in[kN] : 0;
        out[t] = fsub(fadd(base, nv), pv);
    }
    free(pos); free(neg);
}

/* ------------------------------------------------------- threaded batches */
typedef struct { void *h; const u64 *luts; const int *lut_idx; const u64 *ins; u64 *outs; int lo, hi; int with_ks; } job_t;
static void *worker(void *arg) {
    job_t *j = arg; const ctx_t *c = j->h; const params_t *pp = &c->pp;
    int big = pp->k * pp->N + 1, small = pp->n + 1;
    u64 *tmp = malloc(sizeof(u64) * small);
    for (int q = j->lo; q < j->hi; q++) {
        const u64 *in = j->ins + (u64)q * (j->with_ks ? big : small);
        if (j->with_ks) { orcf_keyswitch(j->h, in, tmp); in = tmp; }
        orcf_pbs(j->h, j->luts + (u64)j->lut_idx[q] * pp->N, in, j->outs + (u64)q * big);
    }
    free(tmp);
    return NULL;
}
/* with_ks = 1: inputs are big-key LWEs (keyswitch + PBS); 0: small-key LWEs (PBS only) */
void orcf_batch(void *h, const u64 *luts, const int *lut_idx, const u64 *ins, u64 *outs,
                int count, int with_ks, int threads) {
    if (threads < 1) threads = 1;
    if (threads > count) threads = count;
    pthread_t *th = malloc(sizeof(pthread_t) * threads); job_t *jb = malloc(sizeof(job_t) * threads);
    for (int t = 0; t < threads; t++) {
        jb[t] = (job_t){h, luts, lut_idx, ins, outs, (int)((i64)count * t / threads), (int)((i64)count * (t + 1) / threads), with_ks};
        pthread_create(&th[t], NULL, worker, &jb[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th); free(jb);
}
