// Client side of the engine (host only): secret keys, evaluation keys, encryption, phases.
// Mirrors what the reference obtains from circuit.keygen()/encrypt()/decrypt()
// (/root/reference/matrix_inversion/main.py:177, qfloat_matrix_inversion.py:1032,1035).
// Randomness: ChaCha20 keystreams keyed by 32-byte seeds (OS entropy by default, see bmi_random_seed); the secret keys
// are drawn from their own seed, independent of the one behind the public masks and the noise.
#include <sys/random.h>

#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/bmi_tfhe.h"
#include "field.cuh"
#include "host_common.h"

namespace {

// ChaCha20 (D. J. Bernstein; 20 rounds, 256-bit key, 64-bit block counter, 64-bit nonce) as a seekable
// keystream: word `ctr` of stream `stream` is 64-bit word (ctr & 7) of block (ctr >> 3) under nonce = stream.
// Random access keeps key generation parallel and reproducible for a given seed; the seed is 32 bytes of OS
// entropy unless the caller passes a fixed one (tests, benchmarks).
inline u32 rotl32(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
#define BMI_QR(a, b, c, d) \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);
inline void chacha20_block(const u32 key[8], u64 counter, u64 nonce, u64 out[8]) {
    u32 in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3], key[4], key[5],
                  key[6], key[7], (u32)counter, (u32)(counter >> 32), (u32)nonce, (u32)(nonce >> 32)};
    u32 x[16];
    for (int i = 0; i < 16; i++) x[i] = in[i];
    for (int r = 0; r < 10; r++) {
        BMI_QR(x[0], x[4], x[8], x[12]) BMI_QR(x[1], x[5], x[9], x[13]) BMI_QR(x[2], x[6], x[10], x[14]) BMI_QR(x[3], x[7], x[11], x[15])
        BMI_QR(x[0], x[5], x[10], x[15]) BMI_QR(x[1], x[6], x[11], x[12]) BMI_QR(x[2], x[7], x[8], x[13]) BMI_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 8; i++) out[i] = (u64)(x[2 * i] + in[2 * i]) | ((u64)(x[2 * i + 1] + in[2 * i + 1]) << 32);
}
#undef BMI_QR

// One keystream per (seed, stream).  Copies are cheap and each copy caches its last block, so every worker thread
// takes its own copy and walks its counters mostly sequentially.
struct Rng {
    u32 key[8];
    u64 stream;
    u64 cached = ~0ULL;
    u64 block[8];
    Rng(const uint8_t* seed, u64 st) : stream(st) {
        for (int i = 0; i < 8; i++) key[i] = (u32)seed[4 * i] | ((u32)seed[4 * i + 1] << 8) | ((u32)seed[4 * i + 2] << 16) | ((u32)seed[4 * i + 3] << 24);
    }
    u64 raw(u64 ctr) {
        const u64 b = ctr >> 3;
        if (b != cached) { chacha20_block(key, b, stream, block); cached = b; }
        return block[ctr & 7];
    }
    // uniform in [0, p): rejection sampling; a rejected draw (probability 2^-32) moves to a disjoint counter range
    u64 field(u64 ctr) {
        for (u64 attempt = 0;; attempt++) {
            const u64 r = raw(ctr + (attempt << 58));
            if (r < BMI_P) return r;
        }
    }
    u64 bit(u64 ctr) { return raw(ctr) >> 63; }
    u64 gauss(u64 ctr, double sigma) {
        const u64 r1 = raw(2 * ctr), r2 = raw(2 * ctr + 1);
        const double u1 = (double)((r1 >> 11) + 1) * 0x1.0p-53, u2 = (double)(r2 >> 11) * 0x1.0p-53;
        const double z = std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925286766559 * u2);
        return from_i64((i64)std::llround(z * sigma));
    }
};
enum { ST_LWE_KEY = 1, ST_GLWE_KEY = 2, ST_BSK_MASK = 3, ST_BSK_NOISE = 4, ST_KSK_MASK = 5, ST_KSK_NOISE = 6,
       ST_ENC_MASK = 7, ST_ENC_NOISE = 8, ST_BSKP_MASK = 9, ST_BSKP_NOISE = 10 };

// host negacyclic transform (merged-twiddle butterflies), used only to multiply masks by key polynomials
struct HostNtt {
    int N, L;
    std::vector<u64> tw, twi;
    u64 ninv;
    explicit HostNtt(int n) : N(n), L(0) {
        while ((1 << L) < N) L++;
        bmi_host::twiddles(N, tw, twi);
        ninv = fpow((u64)N, BMI_P - 2);
    }
    void fwd(u64* a) const {
        int t = N;
        for (int m = 1; m < N; m <<= 1) {
            t >>= 1;
            for (int i = 0; i < m; i++) {
                const u64 w = tw[m + i];
                u64 *x = a + 2 * i * t, *y = x + t;
                for (int j = 0; j < t; j++) { const u64 u = x[j], v = fmul(y[j], w); x[j] = fadd(u, v); y[j] = fsub(u, v); }
            }
        }
    }
    void inv(u64* a) const {
        int t = 1;
        for (int m = N >> 1; m >= 1; m >>= 1) {
            for (int i = 0; i < m; i++) {
                const u64 w = twi[m + i];
                u64 *x = a + 2 * i * t, *y = x + t;
                for (int j = 0; j < t; j++) { const u64 u = x[j], v = y[j]; x[j] = fadd(u, v); y[j] = fmul(fsub(u, v), w); }
            }
            t <<= 1;
        }
        for (int j = 0; j < N; j++) a[j] = fmul(a[j], ninv);
    }
};

template <class F>
void parallel_for(int64_t count, int threads, F f) {
    threads = std::max(1, std::min<int>(threads, (int)std::min<int64_t>(count, 256)));
    if (threads == 1) { f(0, count); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(f, count * t / threads, count * (t + 1) / threads);
    for (auto& th : pool) th.join();
}

bool check(const bmi_params* p) {
    if (!p || p->n < 1 || p->k < 1 || p->N < 2 || (p->N & (p->N - 1)) || p->bsk_l < 1 || p->ksk_l < 1 ||
        p->bsk_bl < 1 || p->ksk_bl < 1 || p->bsk_bl * p->bsk_l > 63 || p->ksk_bl * p->ksk_l > 63) {
        bmi_host::set_error("invalid bmi_params");
        return false;
    }
    return true;
}

// `count` GGSW ciphertexts of the bits msg[i]: out[i][r][comp][t], r = c*l + (j-1), body last
void gen_ggsw(const bmi_params* p, const uint8_t* seed, u64 st_mask, u64 st_noise, const u64* S, const u64* msg, int count, u64* out, int threads) {
    const int k = p->k, N = p->N, l = p->bsk_l, rows = (k + 1) * l;
    const HostNtt ntt(N);
    std::vector<u64> Shat((size_t)k * N);
    for (int m = 0; m < k; m++) {
        std::memcpy(&Shat[(size_t)m * N], S + (size_t)m * N, sizeof(u64) * N);
        ntt.fwd(&Shat[(size_t)m * N]);
    }
    const Rng rm0(seed, st_mask), rn0(seed, st_noise);
    parallel_for((int64_t)count * rows, threads, [&](int64_t lo, int64_t hi) {
        Rng rm = rm0, rn = rn0;
        std::vector<u64> sum(N), tmp(N);
        for (int64_t id = lo; id < hi; id++) {
            u64* row = out + (size_t)id * (k + 1) * N;
            u64* body = row + (size_t)k * N;
            std::fill(sum.begin(), sum.end(), 0);
            for (int m = 0; m < k; m++) {
                u64* A = row + (size_t)m * N;
                for (int t = 0; t < N; t++) A[t] = rm.field(((u64)id * k + m) * N + t);
                std::memcpy(tmp.data(), A, sizeof(u64) * N);
                ntt.fwd(tmp.data());
                for (int t = 0; t < N; t++) sum[t] = fadd(sum[t], fmul(tmp[t], Shat[(size_t)m * N + t]));
            }
            ntt.inv(sum.data());
            for (int t = 0; t < N; t++) body[t] = fadd(sum[t], rn.gauss((u64)id * N + t, p->glwe_sigma));
            const int i = (int)(id / rows), r = (int)(id % rows), c = r / l, j = r % l + 1;
            if (msg[i]) row[(size_t)c * N] = fadd(row[(size_t)c * N], 1ULL << (64 - j * p->bsk_bl));
        }
    });
}

}  // namespace

extern "C" {

int bmi_random_seed(uint8_t* seed) {
    if (!seed) { bmi_host::set_error("null argument"); return BMI_EINVAL; }
    size_t got = 0;
    while (got < 32) {
        const ssize_t r = getrandom(seed + got, 32 - got, 0);
        if (r < 0) { bmi_host::set_error("getrandom failed"); return BMI_ESTATE; }
        got += (size_t)r;
    }
    return BMI_OK;
}

int bmi_rng_words(const uint8_t* seed, uint64_t stream, uint64_t ctr0, uint64_t* out, int64_t count) {
    if (!seed || !out || count < 0) { bmi_host::set_error("invalid argument"); return BMI_EINVAL; }
    Rng r(seed, stream);
    for (int64_t i = 0; i < count; i++) out[i] = r.raw(ctr0 + (u64)i);
    return BMI_OK;
}

int bmi_keygen_lwe(const bmi_params* p, const uint8_t* seed, uint64_t* s) {
    if (!check(p) || !s || !seed) return BMI_EINVAL;
    Rng r(seed, ST_LWE_KEY);
    for (int i = 0; i < p->n; i++) s[i] = r.bit(i);
    return BMI_OK;
}

int bmi_keygen_glwe(const bmi_params* p, const uint8_t* seed, uint64_t* S) {
    if (!check(p) || !S || !seed) return BMI_EINVAL;
    Rng r(seed, ST_GLWE_KEY);
    for (int i = 0; i < p->k * p->N; i++) S[i] = r.bit(i);
    return BMI_OK;
}

int bmi_keygen_bsk(const bmi_params* p, const uint8_t* seed, const uint64_t* s, const uint64_t* S, uint64_t* bsk, int threads) {
    if (!check(p) || !s || !S || !bsk || !seed) return BMI_EINVAL;
    gen_ggsw(p, seed, ST_BSK_MASK, ST_BSK_NOISE, S, s, p->n, bsk, threads);
    return BMI_OK;
}

int bmi_keygen_bsk_pairs(const bmi_params* p, const uint8_t* seed, const uint64_t* s, const uint64_t* S, uint64_t* bskp, int threads) {
    if (!check(p) || !s || !S || !bskp || !seed) return BMI_EINVAL;
    if (p->n % 2) { bmi_host::set_error("pair key needs an even LWE dimension"); return BMI_EINVAL; }
    std::vector<u64> msg((size_t)3 * (p->n / 2));
    for (int q = 0; q < p->n / 2; q++) {
        const u64 a = s[2 * q], b = s[2 * q + 1];
        msg[3 * q] = a & b; msg[3 * q + 1] = a & (b ^ 1); msg[3 * q + 2] = (a ^ 1) & b;
    }
    gen_ggsw(p, seed, ST_BSKP_MASK, ST_BSKP_NOISE, S, msg.data(), (int)msg.size(), bskp, threads);
    return BMI_OK;
}

int bmi_keygen_ksk(const bmi_params* p, const uint8_t* seed, const uint64_t* s, const uint64_t* S, uint64_t* ksk, int threads) {
    if (!check(p) || !s || !S || !ksk || !seed) return BMI_EINVAL;
    const int n = p->n, l = p->ksk_l;
    const Rng rm0(seed, ST_KSK_MASK), rn0(seed, ST_KSK_NOISE);
    parallel_for((int64_t)p->k * p->N * l, threads, [&](int64_t lo, int64_t hi) {
        Rng rm = rm0, rn = rn0;
        for (int64_t id = lo; id < hi; id++) {
            u64* ct = ksk + (size_t)id * (n + 1);
            u64 b = rn.gauss((u64)id, p->lwe_sigma);
            for (int t = 0; t < n; t++) {
                ct[t] = rm.field((u64)id * n + t);
                if (s[t]) b = fadd(b, ct[t]);
            }
            const int i = (int)(id / l), j = (int)(id % l) + 1;
            if (S[i]) b = fadd(b, 1ULL << (64 - j * p->ksk_bl));
            ct[n] = b;
        }
    });
    return BMI_OK;
}

int bmi_lwe_encrypt(const bmi_params* p, const uint8_t* seed, uint64_t ct_index0, const uint64_t* S, const uint64_t* pt,
                    int64_t count, uint64_t* out) {
    if (!check(p) || !S || !pt || !out || !seed || count < 0) return BMI_EINVAL;
    const int dim = p->k * p->N;
    const Rng rm0(seed, ST_ENC_MASK), rn0(seed, ST_ENC_NOISE);
    parallel_for(count, 8, [&](int64_t lo, int64_t hi) {
        Rng rm = rm0, rn = rn0;
        for (int64_t q = lo; q < hi; q++) {
            const u64 id = ct_index0 + (u64)q;
            u64* ct = out + (size_t)q * (dim + 1);
            u64 b = fadd(rn.gauss(id, p->glwe_sigma), pt[q] % BMI_P);
            for (int t = 0; t < dim; t++) {
                ct[t] = rm.field(id * dim + t);
                if (S[t]) b = fadd(b, ct[t]);
            }
            ct[dim] = b;
        }
    });
    return BMI_OK;
}

int bmi_lwe_phase(const uint64_t* key, int32_t dim, const uint64_t* ct, int64_t count, uint64_t* phase) {
    if (!key || !ct || !phase || dim < 1 || count < 0) { bmi_host::set_error("invalid argument"); return BMI_EINVAL; }
    for (int64_t q = 0; q < count; q++) {
        const u64* c = ct + (size_t)q * (dim + 1);
        u64 ph = c[dim];
        for (int t = 0; t < dim; t++) if (key[t]) ph = fsub(ph, c[t]);
        phase[q] = ph;
    }
    return BMI_OK;
}

}  // extern "C"
