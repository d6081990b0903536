/*
 * bmi_tfhe.h -- C ABI of the B200-native TFHE execution engine (libbmi_tfhe.so).
 *
 * This is the drop-in boundary for the reference's hot path.  The reference reaches
 * the same functionality through concrete-python's FFI into its compiled runtime:
 *
 *   reference call site (matrix_inversion/...)                    replaced here by
 *   ------------------------------------------------------------  --------------------------
 *   circuit.keygen()            main.py:177                       bmi_keygen_*
 *   circuit.encrypt(x, y)       qfloat_matrix_inversion.py:1032   bmi_lwe_encrypt
 *   circuit.run(encrypted)      qfloat_matrix_inversion.py:1034   one bmi_lincomb + bmi_keyswitch + bmi_pbs per level
 *     every table lookup (%, //, <, abs, sign, enc*enc ...)       bmi_keyswitch + bmi_pbs
 *       base_p_arrays.py:102-103,122,197-198  qfloat.py:619,663-670
 *     every leveled add / sub / scalar multiply                   bmi_lincomb
 *       base_p_arrays.py:101,118,121  qfloat.py:618,620,826,901
 *   circuit.decrypt(result)     qfloat_matrix_inversion.py:1035   bmi_lwe_phase (+ host decode)
 *
 * Plain pointers and sizes only.  All ciphertext words are uint64 in [0, p),
 * p = 2^64 - 2^32 + 1.  "d_" arguments are CUDA device pointers, "h_" host pointers;
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).
 * Every function returns 0 on success, a negative bmi_status otherwise;
 * bmi_last_error() describes the last failure on the calling thread.
 */
#ifndef BMI_TFHE_H
#define BMI_TFHE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bmi_params {
    int32_t n;          /* small LWE dimension */
    int32_t k;          /* GLWE dimension (kernels support k == 1) */
    int32_t N;          /* polynomial size, 1024 | 2048 | 4096 | 8192 | 16384 (16384: bsk_l == 1 only) */
    int32_t bsk_bl;     /* PBS decomposition base log */
    int32_t bsk_l;      /* PBS decomposition levels */
    int32_t ksk_bl;     /* keyswitch base log */
    int32_t ksk_l;      /* keyswitch levels */
    double lwe_sigma;   /* small-key noise std, units of 2^-64 */
    double glwe_sigma;  /* GLWE / big-key noise std, units of 2^-64 */
} bmi_params;

enum bmi_status { BMI_OK = 0, BMI_EINVAL = -1, BMI_ECUDA = -2, BMI_ENOMEM = -3, BMI_ESTATE = -4 };

typedef struct bmi_ctx bmi_ctx;

const char* bmi_version(void);
const char* bmi_last_error(void);

/* ---- client side (host only, no GPU needed): keys, encryption, decryption ----
 * All randomness is ChaCha20 keystream under a 32-byte seed.  Production callers fill every seed with
 * bmi_random_seed (OS entropy) and use DIFFERENT seeds for the secret keys (bmi_keygen_lwe / _glwe), for the
 * evaluation keys (bmi_keygen_bsk* / _ksk) and for encryption; fixed seeds are for tests and benchmarks only.
 * Concrete's circuit.keygen() takes no seed either (main.py:177). */
int bmi_random_seed(uint8_t* seed /* [32] */);
/* raw generator output, words ctr0 .. ctr0+count-1 of keystream `stream` (self-test entry: known-answer vectors) */
int bmi_rng_words(const uint8_t* seed, uint64_t stream, uint64_t ctr0, uint64_t* h_out, int64_t count);
int bmi_keygen_lwe(const bmi_params* p, const uint8_t* seed, uint64_t* h_s /* [n] */);
int bmi_keygen_glwe(const bmi_params* p, const uint8_t* seed, uint64_t* h_S /* [k*N] */);
/* bsk[i][r][comp][t]: i<n, r = c*l+(j-1) < (k+1)*l, comp<=k (k = body), coefficient domain */
int bmi_keygen_bsk(const bmi_params* p, const uint8_t* seed, const uint64_t* h_s, const uint64_t* h_S, uint64_t* h_bsk, int threads);
/* pair key for blind rotation two key bits per step (n even): bskp[q][x][r][comp][t], q<n/2,
 * x = 0: GGSW(s_2q s_2q+1), 1: GGSW(s_2q (1 - s_2q+1)), 2: GGSW((1 - s_2q) s_2q+1); 1.5x the size of bsk */
int bmi_keygen_bsk_pairs(const bmi_params* p, const uint8_t* seed, const uint64_t* h_s, const uint64_t* h_S, uint64_t* h_bskp, int threads);
/* ksk[i][j-1][0..n]: i<k*N, body last */
int bmi_keygen_ksk(const bmi_params* p, const uint8_t* seed, const uint64_t* h_s, const uint64_t* h_S, uint64_t* h_ksk, int threads);
/* `count` big-key encryptions of the plaintexts h_pt (field elements); ciphertext q uses counter ct_index0+q of the
 * seed's mask and noise keystreams: a (seed, counter) pair must never be used for two ciphertexts */
int bmi_lwe_encrypt(const bmi_params* p, const uint8_t* seed, uint64_t ct_index0, const uint64_t* h_S,
                    const uint64_t* h_pt, int64_t count, uint64_t* h_ct /* [count][k*N+1] */);
/* phases b - <a, key> of `count` ciphertexts of dimension dim */
int bmi_lwe_phase(const uint64_t* h_key, int32_t dim, const uint64_t* h_ct, int64_t count, uint64_t* h_phase);

/* ---- server side (GPU) ---- */
int bmi_ctx_create(const bmi_params* p, int device, bmi_ctx** out);
int bmi_ctx_destroy(bmi_ctx* ctx);
int bmi_ctx_load_bsk(bmi_ctx* ctx, const uint64_t* h_bsk);   /* upload + convert to the transform-domain layout */
/* upload + convert the pair key; from then on bmi_pbs / bmi_ks_pbs_host rotate two key bits per step (bsk_l == 1 only):
 * half the sequential transforms per bootstrap, same message, slightly different noise */
int bmi_ctx_load_bsk_pairs(bmi_ctx* ctx, const uint64_t* h_bskp);
int bmi_ctx_load_ksk(bmi_ctx* ctx, const uint64_t* h_ksk);
int bmi_ctx_load_luts(bmi_ctx* ctx, const uint64_t* h_luts, int32_t n_luts);   /* [n_luts][N] accumulator polynomials */
int64_t bmi_ctx_launch_count(const bmi_ctx* ctx);            /* kernels launched by this context so far */
/* bootstrap kernel build: 0 = automatic (picked per launch size), 1 = latency build (4 or 8 coefficients per thread:
 * most warps per transform), 2 = throughput build, 3 = 8-CTA cluster per ciphertext (each polynomial over 4 SMs; one
 * decomposition level only), 5 = the 8-CTA kernel with two points per thread and warp-shuffle butterflies (pair key
 * only; select it BEFORE bmi_ctx_load_bsk_pairs; measured slower than 3, kept for comparison);
 * bmi_polymul_host follows the same choice */
int bmi_ctx_set_pbs_mode(bmi_ctx* ctx, int32_t mode);
/* 1 = bring each step's GGSW rows into shared memory with TMA bulk copies issued one step ahead (where they fit);
 * 0 = per-thread coalesced global loads.  Defaults follow the measurements on B200: on for the
 * pair rotation's latency build (2-3 % faster), off for the one-GGSW-per-bit key (4-8 % slower).  The call sets both. */
int bmi_ctx_set_tma_stage(bmi_ctx* ctx, int32_t on);

/* out[j][b] = sum_t coef[t] * vals[idx[t]][b] + konst[j], rows of k*N+1 words.
 * CSR: d_row_ptr [njobs+1] int32, d_idx int32 (row of d_vals before batch expansion), d_coef uint64 field elements */
int bmi_lincomb(bmi_ctx* ctx, const uint64_t* d_vals, const int32_t* d_row_ptr, const int32_t* d_idx,
                const uint64_t* d_coef, const uint64_t* d_konst, uint64_t* d_out, int32_t njobs, int32_t batch, void* stream);
/* rows of k*N+1 words: d_dst row d_dst_row[j]*batch+b <- d_src row j*batch+b (e.g. all-gathered bootstrap outputs of
 * a level into their value slots) */
int bmi_scatter_rows(bmi_ctx* ctx, const uint64_t* d_src, const int32_t* d_dst_row, uint64_t* d_dst, int32_t count, int32_t batch, void* stream);
/* how many ciphertexts one bmi_pbs launch bootstraps at its minimum latency (the 8-CTA clusters this GPU keeps resident);
 * 0 when the parameter set has no such kernel.  The host-side scheduler sizes circuit levels with it. */
int32_t bmi_ctx_pbs_capacity(bmi_ctx* ctx);
/* count big-key LWEs [count][k*N+1] -> small-key LWEs [count][n+1] */
int bmi_keyswitch(bmi_ctx* ctx, const uint64_t* d_in, uint64_t* d_out, int64_t count, void* stream);
/* njobs*batch bootstraps: job q, lane b reads d_small row job_in[q]*batch+b, applies LUT job_lut[q],
 * writes the big-key LWE to d_out row job_out[q]*batch+b */
int bmi_pbs(bmi_ctx* ctx, const uint64_t* d_small, const int32_t* d_job_in, const int32_t* d_job_lut,
            const int32_t* d_job_out, uint64_t* d_out, int32_t njobs, int32_t batch, void* stream);

/* reference-facing convenience with HOST buffers: keyswitch + PBS of `count` big-key LWEs,
 * ciphertext q looked up through LUT h_lut_idx[q]; copies in and out included */
int bmi_ks_pbs_host(bmi_ctx* ctx, const uint64_t* h_in, const int32_t* h_lut_idx, uint64_t* h_out, int64_t count);

/* self-test entry: c = a*b mod (X^N+1, p) for `count` polynomial pairs, through the device transforms */
int bmi_polymul_host(bmi_ctx* ctx, const uint64_t* h_a, const uint64_t* h_b, uint64_t* h_c, int32_t count);

#ifdef __cplusplus
}
#endif
#endif
