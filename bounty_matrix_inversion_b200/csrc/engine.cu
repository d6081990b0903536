// Server side of the engine: device context, key upload/conversion, kernel launchers, C ABI.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "launchers.cuh"
#include "leveled.cuh"

namespace bmi_host {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
}  // namespace bmi_host
#define BMI_EXTERN_L(L)                                                                                  \
    extern template int setup_attrs<L>(const bmi_ctx*);                                                  \
    extern template int launch_convert<L>(bmi_ctx*, const u64*, u64* const*, int64_t, int64_t, cudaStream_t); \
    extern template int launch_pbs<L>(bmi_ctx*, PbsArgs, cudaStream_t);                                  \
    extern template int launch_polymul<L>(bmi_ctx*, const u64*, const u64*, u64*, int, cudaStream_t); \
    extern template int64_t split_capacity<L>(bmi_ctx*);
BMI_EXTERN_L(10) BMI_EXTERN_L(11) BMI_EXTERN_L(12) BMI_EXTERN_L(13) BMI_EXTERN_L(14)

namespace {

// Every ABI entry point that launches or allocates runs with the context's device current and restores the caller's.
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != device && cudaSetDevice(device) != cudaSuccess) { ok = false; set_error("cudaSetDevice failed"); }
        if (prev == device) prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define GUARD(c)                    \
    DeviceGuard guard_((c)->device); \
    if (!guard_.ok) return BMI_ECUDA

#define DISPATCH_L(c, expr)                                  \
    switch ((c)->logN) {                                     \
        case 10: { constexpr int L = 10; return expr; }      \
        case 11: { constexpr int L = 11; return expr; }      \
        case 12: { constexpr int L = 12; return expr; }      \
        case 13: { constexpr int L = 13; return expr; }      \
        case 14: { constexpr int L = 14; return expr; }      \
        default: set_error("unsupported polynomial size"); return BMI_EINVAL; \
    }

int do_setup(bmi_ctx* c) { DISPATCH_L(c, setup_attrs<L>(c)); }
int do_convert(bmi_ctx* c, const u64* s, u64* const* dst, int64_t p0, int64_t polys, cudaStream_t st) { DISPATCH_L(c, launch_convert<L>(c, s, dst, p0, polys, st)); }
int do_pbs(bmi_ctx* c, const PbsArgs& a, cudaStream_t st) { DISPATCH_L(c, launch_pbs<L>(c, a, st)); }
int64_t do_capacity(bmi_ctx* c) { DISPATCH_L(c, split_capacity<L>(c)); }
int do_polymul(bmi_ctx* c, const u64* a, const u64* b, u64* o, int n, cudaStream_t st) { DISPATCH_L(c, launch_polymul<L>(c, a, b, o, n, st)); }

// upload a standard-domain key of `polys` polynomials in slices through a bounded staging buffer, converting slice by
// slice into every transform-domain layout this parameter set can run (dst[0..2])
int upload_converted(bmi_ctx* c, const u64* h_key, int64_t polys, u64** dst, int layouts) {
    const size_t bytes = (size_t)polys * c->p.N * 8;
    for (int v = 0; v < layouts; v++) {
        // layout 3 (two-points-per-thread kernel) only when that kernel was selected before the key is loaded
        const bool needed = v == 3 ? (c->logN <= kMaxSplit2L && c->pbs_mode == 5) : v == 2 ? c->p.bsk_l == 1 : c->logN <= kMaxClusterL;
        if (needed && !dst[v]) CK(cudaMalloc(&dst[v], bytes));
    }
    const int64_t slice = std::min<int64_t>(polys, 4096);
    u64* stage = nullptr;
    CK(cudaMalloc(&stage, (size_t)slice * c->p.N * 8));
    for (int64_t p0 = 0; p0 < polys; p0 += slice) {
        const int64_t cnt = std::min(slice, polys - p0);
        CK(cudaMemcpy(stage, h_key + (size_t)p0 * c->p.N, (size_t)cnt * c->p.N * 8, cudaMemcpyHostToDevice));
        int rc = do_convert(c, stage, dst, p0, cnt, 0);
        if (rc) { cudaFree(stage); return rc; }
        CK(cudaDeviceSynchronize());
    }
    cudaFree(stage);
    return BMI_OK;
}

int ensure_scratch(bmi_ctx* c, int64_t count) {
    if (count <= c->w_cap) return BMI_OK;
    cudaFree(c->w_in); cudaFree(c->w_small); cudaFree(c->w_out); cudaFree(c->w_idx); cudaFree(c->w_lut);
    c->w_in = c->w_small = c->w_out = nullptr;      // a failed allocation below must not leave dangling pointers behind
    c->w_idx = c->w_lut = nullptr;
    c->w_cap = 0;
    const size_t big = (size_t)c->p.k * c->p.N + 1;
    CK(cudaMalloc(&c->w_in, count * big * 8));
    CK(cudaMalloc(&c->w_small, count * (size_t)(c->p.n + 1) * 8));
    CK(cudaMalloc(&c->w_out, count * big * 8));
    CK(cudaMalloc(&c->w_idx, count * sizeof(int)));
    CK(cudaMalloc(&c->w_lut, count * sizeof(int)));
    std::vector<int> iota(count);
    for (int64_t i = 0; i < count; i++) iota[i] = (int)i;
    CK(cudaMemcpy(c->w_idx, iota.data(), count * sizeof(int), cudaMemcpyHostToDevice));
    c->w_cap = count;
    return BMI_OK;
}

int ctx_init_device(bmi_ctx* c) {
    std::vector<u64> tw, twi;
    bmi_host::twiddles(c->p.N, tw, twi);
    CK(cudaMalloc(&c->d_tw, c->p.N * 8));
    CK(cudaMalloc(&c->d_twi, c->p.N * 8));
    CK(cudaMemcpy(c->d_tw, tw.data(), c->p.N * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->d_twi, twi.data(), c->p.N * 8, cudaMemcpyHostToDevice));
    return do_setup(c);
}

}  // namespace

extern "C" {

const char* bmi_version(void) { return "bmi_tfhe 0.1 (sm_100a)"; }
const char* bmi_last_error(void) { return bmi_host::g_err.c_str(); }

int bmi_ctx_create(const bmi_params* p, int device, bmi_ctx** out) {
    if (!p || !out) { set_error("null argument"); return BMI_EINVAL; }
    if (p->k != 1) { set_error("kernels support GLWE dimension k == 1 only"); return BMI_EINVAL; }
    int logN = 0;
    while ((1 << logN) < p->N) logN++;
    if ((1 << logN) != p->N || logN < 10 || logN > 14) { set_error("polynomial size must be 1024..16384"); return BMI_EINVAL; }
    if (logN == 14 && p->bsk_l != 1) { set_error("N = 16384 runs on the 8-CTA split kernel, which needs one decomposition level"); return BMI_EINVAL; }
    int ks_rows_log = 0;
    while ((1LL << ks_rows_log) < (long long)p->k * p->N * p->ksk_l) ks_rows_log++;
    if (p->bsk_bl * p->bsk_l > 63 || p->ksk_bl * p->ksk_l > 63 || ks_rows_log + p->ksk_bl > 32 || p->n < 1 || p->n > 4096) {
        set_error("unsupported decomposition / dimension");
        return BMI_EINVAL;
    }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) { set_error("no such CUDA device"); return BMI_ECUDA; }
    DeviceGuard guard(device);
    if (!guard.ok) return BMI_ECUDA;
    bmi_ctx* c = new bmi_ctx();
    c->p = *p; c->device = device; c->logN = logN;
    c->ninv = fpow((u64)p->N, BMI_P - 2);
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (const char* v = getenv("BMI_TMA_STAGE")) c->tma_stage = c->tma_stage_pairs = v[0] == '1';
    if (const char* v = getenv("BMI_SPLIT_ASYNC")) c->split_async = v[0] != '0';
    if (const char* v = getenv("BMI_KS_CTAS_PER_SM")) c->ks_ctas_per_sm = std::max(1, atoi(v));
    int rc = ctx_init_device(c);
    if (rc) { bmi_ctx_destroy(c); return rc; }      // nothing allocated so far outlives a failed create
    *out = c;
    return BMI_OK;
}

int bmi_ctx_destroy(bmi_ctx* c) {
    if (!c) return BMI_OK;
    DeviceGuard guard(c->device);
    cudaFree(c->d_tw); cudaFree(c->d_twi); cudaFree(c->d_bsk[0]); cudaFree(c->d_bsk[1]); cudaFree(c->d_bsk[2]); cudaFree(c->d_ksk); cudaFree(c->d_luts);
    cudaFree(c->w_in); cudaFree(c->w_small); cudaFree(c->w_out); cudaFree(c->w_idx); cudaFree(c->w_lut);
    cudaFree(c->ks_partial); cudaFree(c->d_ks_corr);
    for (int v = 0; v < 4; v++) { cudaFree(c->d_bskp[v]); cudaFree(c->d_expo[v]); }
    cudaFree(c->d_pw);
    delete c;
    return BMI_OK;
}

int bmi_ctx_load_bsk(bmi_ctx* c, const uint64_t* h_bsk) {
    if (!c || !h_bsk) { set_error("null argument"); return BMI_EINVAL; }
    GUARD(c);
    u64* dst[4] = {c->d_bsk[0], c->d_bsk[1], c->d_bsk[2], nullptr};
    const int rc = upload_converted(c, h_bsk, (int64_t)c->p.n * 2 * c->p.bsk_l * 2, dst, 3);
    for (int v = 0; v < 3; v++) c->d_bsk[v] = dst[v];
    return rc;
}

int bmi_ctx_load_bsk_pairs(bmi_ctx* c, const uint64_t* h_bskp) {
    if (!c || !h_bskp) { set_error("null argument"); return BMI_EINVAL; }
    if (c->p.bsk_l != 1 || c->p.n % 2) { set_error("pair blind rotation needs one decomposition level and an even LWE dimension"); return BMI_EINVAL; }
    GUARD(c);
    int rc = upload_converted(c, h_bskp, (int64_t)(c->p.n / 2) * 3 * 2 * 2, c->d_bskp, 4);
    if (rc) return rc;
    for (int v = 0; v < 4; v++) {   // every layout: K10 <- K11 + K10, K01 <- K11 + K01 (what the pointwise stage multiplies with)
        if (!c->d_bskp[v]) continue;
        const size_t third = (size_t)4 * c->p.N, pairs = (size_t)c->p.n / 2;
        pair_key_sums_kernel<<<(unsigned)((third * pairs + 255) / 256), 256>>>(c->d_bskp[v], third, pairs);
        c->launches++;
        CK(cudaGetLastError());
    }
    CK(cudaDeviceSynchronize());
    // powers of psi, and for every layout the exponent of each transform slot's evaluation point: the transform of
    // the monomial X holds the evaluation points themselves (times 1/N, which the conversion folds in)
    const int N = c->p.N;
    const u64 psi = fpow(7, (BMI_P - 1) / (2 * (u64)N));
    std::vector<u64> pw(2 * (size_t)N);
    pw[0] = 1;
    for (int t = 1; t < 2 * N; t++) pw[t] = fmul(pw[t - 1], psi);
    if (!c->d_pw) CK(cudaMalloc(&c->d_pw, pw.size() * 8));
    CK(cudaMemcpy(c->d_pw, pw.data(), pw.size() * 8, cudaMemcpyHostToDevice));
    std::vector<std::pair<u64, u32>> index(pw.size());
    for (u32 t = 0; t < pw.size(); t++) index[t] = {pw[t], t};
    std::sort(index.begin(), index.end());
    std::vector<u64> mono(N, 0);
    mono[1] = 1;
    u64 *d_src = nullptr, *d_dst[4] = {nullptr, nullptr, nullptr, nullptr};
    CK(cudaMalloc(&d_src, (size_t)N * 8));
    CK(cudaMemcpy(d_src, mono.data(), (size_t)N * 8, cudaMemcpyHostToDevice));
    for (int v = 0; v < 4; v++) CK(cudaMalloc(&d_dst[v], (size_t)N * 8));
    rc = do_convert(c, d_src, d_dst, 0, 1, 0);
    if (!rc && cudaDeviceSynchronize() != cudaSuccess) { set_error("transform of X failed"); rc = BMI_ECUDA; }
    std::vector<u64> pts(N);
    std::vector<u32> expo(N);
    for (int v = 0; v < 4 && !rc; v++) {
        if (!c->d_bskp[v]) continue;
        if (cudaMemcpy(pts.data(), d_dst[v], (size_t)N * 8, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("copy of transform of X failed"); rc = BMI_ECUDA; break; }
        for (int u = 0; u < N; u++) {
            const u64 point = fmul(pts[u], (u64)N);
            auto it = std::lower_bound(index.begin(), index.end(), std::make_pair(point, (u32)0));
            if (it == index.end() || it->first != point) { set_error("transform slot is not an evaluation at a power of psi"); rc = BMI_ESTATE; break; }
            expo[u] = it->second;
        }
        if (rc) break;
        if (!c->d_expo[v] && cudaMalloc(&c->d_expo[v], (size_t)N * 4) != cudaSuccess) { set_error("out of device memory"); rc = BMI_ENOMEM; break; }
        if (cudaMemcpy(c->d_expo[v], expo.data(), (size_t)N * 4, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("upload of exponents failed"); rc = BMI_ECUDA; }
    }
    cudaFree(d_src);
    for (int v = 0; v < 4; v++) cudaFree(d_dst[v]);
    if (!rc) c->pairs = true;
    return rc;
}

int bmi_ctx_load_ksk(bmi_ctx* c, const uint64_t* h_ksk) {
    if (!c || !h_ksk) { set_error("null argument"); return BMI_EINVAL; }
    GUARD(c);
    const size_t bytes = (size_t)c->p.k * c->p.N * c->p.ksk_l * (c->p.n + 1) * 8;
    if (!c->d_ksk) CK(cudaMalloc(&c->d_ksk, bytes));
    CK(cudaMemcpy(c->d_ksk, h_ksk, bytes, cudaMemcpyHostToDevice));
    if (!c->d_ks_corr) CK(cudaMalloc(&c->d_ks_corr, (size_t)(c->p.n + 1) * 8));
    keyswitch_corr_kernel<<<(c->p.n + 1 + 255) / 256, 256>>>(c->d_ksk, c->d_ks_corr, c->p.k * c->p.N * c->p.ksk_l, c->p.n, c->p.ksk_bl);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    return BMI_OK;
}

int bmi_ctx_load_luts(bmi_ctx* c, const uint64_t* h_luts, int32_t n_luts) {
    if (!c || !h_luts || n_luts < 1) { set_error("invalid argument"); return BMI_EINVAL; }
    GUARD(c);
    cudaFree(c->d_luts);
    c->d_luts = nullptr;
    CK(cudaMalloc(&c->d_luts, (size_t)n_luts * c->p.N * 8));
    CK(cudaMemcpy(c->d_luts, h_luts, (size_t)n_luts * c->p.N * 8, cudaMemcpyHostToDevice));
    c->n_luts = n_luts;
    return BMI_OK;
}

int64_t bmi_ctx_launch_count(const bmi_ctx* c) { return c ? c->launches : 0; }

int bmi_ctx_set_tma_stage(bmi_ctx* c, int32_t on) {
    if (!c) { set_error("invalid argument"); return BMI_EINVAL; }
    c->tma_stage = c->tma_stage_pairs = on != 0;
    return BMI_OK;
}

int bmi_ctx_set_pbs_mode(bmi_ctx* c, int32_t mode) {
    if (!c || mode < 0 || mode > 5 || mode == 4) { set_error("invalid argument"); return BMI_EINVAL; }
    if (mode == 5 && c->pairs && !c->d_bskp[3]) { set_error("select mode 5 before bmi_ctx_load_bsk_pairs: its key layout is built at load time"); return BMI_ESTATE; }
    c->pbs_mode = mode;
    return BMI_OK;
}

int bmi_lincomb(bmi_ctx* c, const uint64_t* d_vals, const int32_t* d_row_ptr, const int32_t* d_idx, const uint64_t* d_coef,
                const uint64_t* d_konst, uint64_t* d_out, int32_t njobs, int32_t batch, void* stream) {
    if (!c || njobs < 0 || batch < 1) { set_error("invalid argument"); return BMI_EINVAL; }
    if (njobs == 0) return BMI_OK;
    GUARD(c);
    const int W = c->p.k * c->p.N + 1;
    dim3 grid((unsigned)((int64_t)njobs * batch), (W + 2047) / 2048);
    lincomb_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_vals, d_row_ptr, d_idx, d_coef, d_konst, d_out, W, batch);
    c->launches++;
    CK(cudaGetLastError());
    return BMI_OK;
}

int bmi_scatter_rows(bmi_ctx* c, const uint64_t* d_src, const int32_t* d_dst_row, uint64_t* d_dst, int32_t count, int32_t batch, void* stream) {
    if (!c || count < 0 || batch < 1) { set_error("invalid argument"); return BMI_EINVAL; }
    if (count == 0) return BMI_OK;
    GUARD(c);
    const int W = c->p.k * c->p.N + 1;
    dim3 grid((unsigned)((int64_t)count * batch), (W + 2047) / 2048);
    scatter_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_src, d_dst_row, d_dst, W, batch);
    c->launches++;
    CK(cudaGetLastError());
    return BMI_OK;
}

int32_t bmi_ctx_pbs_capacity(bmi_ctx* c) {
    if (!c) return 0;
    if (c->p.bsk_l != 1) return 0;
    DeviceGuard guard(c->device);
    return (int32_t)do_capacity(c);
}

int bmi_keyswitch(bmi_ctx* c, const uint64_t* d_in, uint64_t* d_out, int64_t count, void* stream) {
    if (!c || count < 0) { set_error("invalid argument"); return BMI_EINVAL; }
    if (!c->d_ksk) { set_error("keyswitch key not loaded"); return BMI_ESTATE; }
    if (count == 0) return BMI_OK;
    GUARD(c);
    const int cols = (c->p.n + 1 + KS_COLS - 1) / KS_COLS, tiles = (int)((count + KS_JT - 1) / KS_JT);
    const int kN = c->p.k * c->p.N;
    // enough CTAs for ks_ctas_per_sm per SM: batches split the input coefficients over blockIdx.z
    int slices = std::max(1, std::min(kN / KS_CHUNK, (c->ks_ctas_per_sm * c->num_sms) / std::max(1, cols * tiles)));
    int slice = ((kN + slices - 1) / slices + KS_CHUNK - 1) / KS_CHUNK * KS_CHUNK;
    slices = (kN + slice - 1) / slice;
    if (slices > 1) {
        const size_t need = (size_t)slices * count * (c->p.n + 1);
        if (need > c->ks_partial_cap) {
            cudaFree(c->ks_partial);
            c->ks_partial = nullptr; c->ks_partial_cap = 0;
            CK(cudaMalloc(&c->ks_partial, need * 8));
            c->ks_partial_cap = need;
        }
    }
    dim3 grid(tiles, cols, slices);
    const size_t smem = (size_t)KS_CHUNK * c->p.ksk_l * KS_JT * sizeof(int);
    keyswitch_kernel<<<grid, KS_COLS, smem, (cudaStream_t)stream>>>(d_in, c->d_ksk, c->d_ks_corr, d_out, c->ks_partial, (int)count, kN, c->p.n,
                                                                   c->p.ksk_bl, c->p.ksk_l, slice);
    c->launches++;
    CK(cudaGetLastError());
    if (slices > 1) {
        const size_t elems = (size_t)count * (c->p.n + 1);
        keyswitch_finish_kernel<<<(unsigned)((elems + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_in, c->ks_partial, c->d_ks_corr, d_out, (int)count,
                                                                                                  kN, c->p.n, slices);
        c->launches++;
        CK(cudaGetLastError());
    }
    return BMI_OK;
}

int bmi_pbs(bmi_ctx* c, const uint64_t* d_small, const int32_t* d_job_in, const int32_t* d_job_lut, const int32_t* d_job_out,
            uint64_t* d_out, int32_t njobs, int32_t batch, void* stream) {
    if (!c || njobs < 0 || batch < 1) { set_error("invalid argument"); return BMI_EINVAL; }
    if (!(c->d_bsk[0] || c->d_bsk[2] || c->pairs) || !c->d_luts) { set_error("bootstrapping key / LUTs not loaded"); return BMI_ESTATE; }
    if (njobs == 0) return BMI_OK;
    GUARD(c);
    PbsArgs a;
    a.bsk_hat = nullptr; a.tw = c->d_tw; a.twi = c->d_twi; a.luts = c->d_luts; a.small = d_small;
    a.job_in = d_job_in; a.job_lut = d_job_lut; a.job_out = d_job_out; a.out = d_out;
    a.njobs = njobs; a.batch = batch; a.n = c->p.n; a.bl = c->p.bsk_bl; a.l = c->p.bsk_l;
    a.expo = nullptr; a.pw = nullptr;
    return do_pbs(c, a, (cudaStream_t)stream);
}

int bmi_ks_pbs_host(bmi_ctx* c, const uint64_t* h_in, const int32_t* h_lut_idx, uint64_t* h_out, int64_t count) {
    if (!c || !h_in || !h_lut_idx || !h_out || count < 0) { set_error("invalid argument"); return BMI_EINVAL; }
    if (count == 0) return BMI_OK;
    GUARD(c);
    for (int64_t q = 0; q < count; q++)
        if (h_lut_idx[q] < 0 || h_lut_idx[q] >= c->n_luts) { set_error("LUT index out of range"); return BMI_EINVAL; }
    int rc = ensure_scratch(c, count);
    if (rc) return rc;
    const size_t big = (size_t)c->p.k * c->p.N + 1;
    CK(cudaMemcpyAsync(c->w_in, h_in, count * big * 8, cudaMemcpyHostToDevice, 0));
    CK(cudaMemcpyAsync(c->w_lut, h_lut_idx, count * sizeof(int), cudaMemcpyHostToDevice, 0));
    if ((rc = bmi_keyswitch(c, c->w_in, c->w_small, count, nullptr))) return rc;
    if ((rc = bmi_pbs(c, c->w_small, c->w_idx, c->w_lut, c->w_idx, c->w_out, (int)count, 1, nullptr))) return rc;
    CK(cudaMemcpyAsync(h_out, c->w_out, count * big * 8, cudaMemcpyDeviceToHost, 0));
    CK(cudaStreamSynchronize(0));
    return BMI_OK;
}

int bmi_polymul_host(bmi_ctx* c, const uint64_t* h_a, const uint64_t* h_b, uint64_t* h_c, int32_t count) {
    if (!c || !h_a || !h_b || !h_c || count < 1) { set_error("invalid argument"); return BMI_EINVAL; }
    GUARD(c);
    const size_t bytes = (size_t)count * c->p.N * 8;
    u64 *a = nullptr, *b = nullptr, *o = nullptr;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&o, bytes));
    CK(cudaMemcpy(a, h_a, bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(b, h_b, bytes, cudaMemcpyHostToDevice));
    int rc = do_polymul(c, a, b, o, count, 0);
    if (!rc) {
        cudaError_t e = cudaMemcpy(h_c, o, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { set_error(cudaGetErrorString(e)); rc = BMI_ECUDA; }
    }
    cudaFree(a); cudaFree(b); cudaFree(o);
    return rc;
}

}  // extern "C"
