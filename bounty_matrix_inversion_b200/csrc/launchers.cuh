// Per-polynomial-size launchers of the device kernels.  Each size is instantiated in its own translation unit
// (launch_inst.cu compiled with -DBMI_INST_L=<log2 N>) so the sizes build in parallel; engine.cu sees only the
// declarations.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>

#include "../../include/bmi_tfhe.h"
#include "host_common.h"
#include "split2.cuh"

using bmi_host::set_error;

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                         \
            return BMI_ECUDA;                                                                      \
        }                                                                                          \
    } while (0)

struct bmi_ctx {
    bmi_params p;
    int device, logN;
    u64 ninv;
    u64 *d_tw = nullptr, *d_twi = nullptr, *d_ksk = nullptr, *d_luts = nullptr;
    u64* d_bsk[3] = {nullptr, nullptr, nullptr};   // transform-domain key: throughput build / latency build / 8-CTA split kernel
    // pair blind rotation (bmi_ctx_load_bsk_pairs): the pair key in the same three layouts, per layout the exponent of
    // every transform slot's evaluation point, and the powers of psi
    // [3]: two-points-per-thread 8-CTA kernel (split2.cuh), keys stored as K11, K11+K10, K11+K01
    u64* d_bskp[4] = {nullptr, nullptr, nullptr, nullptr};
    u32* d_expo[4] = {nullptr, nullptr, nullptr, nullptr};
    u64* d_pw = nullptr;
    bool pairs = false;
    int n_luts = 0;
    int num_sms = 148;
    int64_t launches = 0;
    int split_clusters[2] = {-1, -1};   // resident 8-CTA clusters of the split kernel, single / pair rotation (queried once)
    int latency_resident = -1;          // CTAs per SM of the latency build (queried once)
    bool split_async = true;  // split kernel synchronised by mbarriers + st.async (BMI_SPLIT_ASYNC=0: cluster barriers)
    bool tma_stage = false;  // one GGSW per bit: stage its rows with TMA bulk copies where shared memory allows (measured 4-8 % slower: off)
    bool tma_stage_pairs = true;   // pair rotation, latency build: the step's three GGSWs by TMA one step ahead (measured 2-3 % faster: on)
    // 0 auto (build chosen per launch), 1 latency build, 2 throughput build, 3 8-CTA split kernel,
    // 5 8-CTA kernel with two points per thread and warp-shuffle stages (split2.cuh; measured slower, kept for A/B)
    int pbs_mode = 0;
    // scratch for the host-buffer convenience path
    u64 *w_in = nullptr, *w_small = nullptr, *w_out = nullptr;
    int *w_idx = nullptr, *w_lut = nullptr;
    int64_t w_cap = 0;
    u64* ks_partial = nullptr;   // per-slice keyswitch sums (small batches)
    u64* d_ks_corr = nullptr;    // [n+1] B/2 * column sums of the keyswitch key (digits are used shifted by B/2)
    size_t ks_partial_cap = 0;
    int ks_ctas_per_sm = 16;     // keyswitch batches split the input dimension until the grid has this many CTAs per SM (measured: 3 -> 16 is 26 % faster at 33 rows, 13 % at 592)
};

constexpr int kMaxSmem = 227 * 1024;   // dynamic shared memory a CTA can opt into on sm_100

inline size_t pbs_smem(const bmi_ctx* c) { return (size_t)3 * c->p.N * 8 + (((size_t)c->p.n * 2 + 15) & ~(size_t)15); }

inline size_t split_smem(const bmi_ctx* c) { return (size_t)6 * (c->p.N / 4) * 8 + (((size_t)c->p.n * 2 + 15) & ~(size_t)15); }
// pair rotation: plus the 2N powers of psi up to N = 4096
inline size_t split_smem_pairs(const bmi_ctx* c) { return split_smem(c) + (c->logN <= 12 ? (size_t)2 * c->p.N * 8 : 0); }
template <int L>
inline size_t split2_smem(const bmi_ctx* c) { return (size_t)Split2Cfg<L>::WORDS * 8 + (((size_t)c->p.n * 2 + 15) & ~(size_t)15); }
constexpr int kMaxSplit2L = 13;    // two points per thread: N/8 threads per CTA
inline size_t pbs_smem_staged(const bmi_ctx* c) { return pbs_smem(c) + (size_t)2 * c->p.N * 8; }
inline size_t pbs_smem_staged_pairs(const bmi_ctx* c) { return pbs_smem(c) + (size_t)6 * c->p.N * 8; }   // the three keys of a pair step

template <int L>
constexpr int split_convert_e() { return L <= 12 ? 2 : L == 13 ? 3 : 4; }
constexpr int kMaxClusterL = 13;   // largest polynomial the 2-CTA cluster kernels hold in shared memory; above: split kernel only

template <int L>
int setup_attrs(const bmi_ctx* c) {
    CK(cudaFuncSetAttribute(pbs_split_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)split_smem(c)));
    CK(cudaFuncSetAttribute(pbs_split_async_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)split_smem(c)));
    CK(cudaFuncSetAttribute(pbs_split_async_kernel<L, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)split_smem_pairs(c)));
    CK(cudaFuncSetAttribute(polymul_split_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * SplitCfg<L>::M * 8));
    CK(cudaFuncSetAttribute(bsk_convert_split_kernel<L, split_convert_e<L>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << L) * 8));
    if constexpr (L <= kMaxSplit2L) {
        CK(cudaFuncSetAttribute(pbs_split2_kernel<L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)split2_smem<L>(c)));
        CK(cudaFuncSetAttribute(bsk_convert_split2_kernel<L, split_convert_e<L>()>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << L) * 8));
    }
    if constexpr (L <= kMaxClusterL) {
        constexpr int TP = throughput_ctas_per_sm<L>(), EL = latency_e<L>(), ET = throughput_e<L>();
        const int sm = (int)pbs_smem(c), sms = (int)pbs_smem_staged(c);
        CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, EL, 1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, EL, 1, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, ET, TP, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, ET, TP, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, EL, 1, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, ET, TP, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        if (sms <= kMaxSmem) CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, EL, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sms));
        if ((int)pbs_smem_staged_pairs(c) <= kMaxSmem)
            CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, EL, 1, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pbs_smem_staged_pairs(c)));
        if (sms * TP <= kMaxSmem) CK(cudaFuncSetAttribute(pbs_cluster_kernel<L, ET, TP, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sms));
        CK(cudaFuncSetAttribute(bsk_convert_kernel<L, EL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << L) * 8));
        CK(cudaFuncSetAttribute(bsk_convert_kernel<L, ET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << L) * 8));
        CK(cudaFuncSetAttribute(polymul_kernel<L, EL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << L) * 8));
        CK(cudaFuncSetAttribute(polymul_kernel<L, ET>, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 << L) * 8));
    }
    return BMI_OK;
}

// both layouts of the transform-domain key: [0] throughput build (E = 4), [1] latency build
template <int L>
int launch_convert(bmi_ctx* c, const u64* src, u64* const* dst, int64_t p0, int64_t polys, cudaStream_t st) {
    const size_t off = (size_t)p0 * c->p.N;
    if constexpr (L <= kMaxClusterL) {
        constexpr int EL = latency_e<L>(), ET = throughput_e<L>();
        bsk_convert_kernel<L, ET><<<(unsigned)polys, NttCfg<L, ET>::T, (1 << L) * 8, st>>>(src, dst[0] + off, c->d_tw, c->ninv);
        bsk_convert_kernel<L, EL><<<(unsigned)polys, NttCfg<L, EL>::T, (1 << L) * 8, st>>>(src, dst[1] + off, c->d_tw, c->ninv);
        c->launches += 2;
    }
    if (c->p.bsk_l == 1) {
        constexpr int EC = split_convert_e<L>();
        bsk_convert_split_kernel<L, EC><<<(unsigned)polys, NttCfg<L, EC>::T, (1 << L) * 8, st>>>(src, dst[2] + off, c->d_tw, c->ninv);
        c->launches++;
        if constexpr (L <= kMaxSplit2L) {
            if (dst[3]) {
                bsk_convert_split2_kernel<L, EC><<<(unsigned)polys, NttCfg<L, EC>::T, (1 << L) * 8, st>>>(src, dst[3] + off, c->d_tw, c->ninv);
                c->launches++;
            }
        }
    }
    CK(cudaGetLastError());
    return BMI_OK;
}

// how many 8-CTA clusters of the split kernel the GPU keeps resident at once (cluster placement is per GPC)
template <int L>
int64_t split_capacity(bmi_ctx* c) {
    int& cached = c->split_clusters[c->pairs ? 1 : 0];      // of the kernel launch_split actually launches
    if (cached < 0) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(8 * 64);
        cfg.blockDim = dim3(SplitCfg<L>::T);
        cfg.dynamicSmemBytes = c->pairs ? split_smem_pairs(c) : split_smem(c);
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 8; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = 0;
        const cudaError_t e = c->pairs ? cudaOccupancyMaxActiveClusters(&n, pbs_split_async_kernel<L, true>, &cfg)
                                       : cudaOccupancyMaxActiveClusters(&n, pbs_split_async_kernel<L, false>, &cfg);
        if (e != cudaSuccess || n < 1) { cudaGetLastError(); n = c->num_sms / 10; }
        cached = n;
    }
    return cached;
}

template <int L>
int launch_split(bmi_ctx* c, PbsArgs a, int64_t total, cudaStream_t st) {
    if constexpr (L <= kMaxSplit2L) {
        if (c->pairs && c->d_bskp[3] && c->pbs_mode == 5) {
            a.bsk_hat = c->d_bskp[3]; a.expo = c->d_expo[3]; a.pw = c->d_pw;
            pbs_split2_kernel<L><<<8 * (unsigned)std::min<int64_t>(total, 1 << 16), Split2Cfg<L>::T, split2_smem<L>(c), st>>>(a);
            c->launches++;
            CK(cudaGetLastError());
            return BMI_OK;
        }
    }
    if (c->pairs) {
        a.bsk_hat = c->d_bskp[2]; a.expo = c->d_expo[2]; a.pw = c->d_pw;
        pbs_split_async_kernel<L, true><<<8 * (unsigned)std::min<int64_t>(total, 1 << 16), SplitCfg<L>::T, split_smem_pairs(c), st>>>(a);
        c->launches++;
        CK(cudaGetLastError());
        return BMI_OK;
    }
    a.bsk_hat = c->d_bsk[2];
    if (c->split_async) pbs_split_async_kernel<L><<<8 * (unsigned)std::min<int64_t>(total, 1 << 16), SplitCfg<L>::T, split_smem(c), st>>>(a);
    else pbs_split_kernel<L><<<8 * (unsigned)std::min<int64_t>(total, 1 << 16), SplitCfg<L>::T, split_smem(c), st>>>(a);
    c->launches++;
    CK(cudaGetLastError());
    return BMI_OK;
}

template <int L>
int launch_pbs(bmi_ctx* c, PbsArgs a, cudaStream_t st) {
    const int64_t total = (int64_t)a.njobs * a.batch;
    const bool one = a.l == 1;
    if constexpr (L > kMaxClusterL) {
        return launch_split<L>(c, a, total, st);
    } else {
    constexpr int TP = throughput_ctas_per_sm<L>(), EL = latency_e<L>(), ET = throughput_e<L>();
    const unsigned grid = 2 * (unsigned)std::min<int64_t>(total, 1 << 20);      // one CTA pair per ciphertext
    // While the launch fits the CTA pairs the latency build keeps resident, latency wins; beyond one wave the
    // 16-coefficients-per-thread build (fewer shared-memory round trips, more ciphertexts per SM) does.
    if (c->latency_resident < 0) {
        int resident = 1;
        if (one) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, pbs_cluster_kernel<L, EL, 1, true, false>, NttCfg<L, EL>::T, pbs_smem(c));
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, pbs_cluster_kernel<L, EL, 1, false, false>, NttCfg<L, EL>::T, pbs_smem(c));
        c->latency_resident = std::max(resident, 1);
    }
    const int64_t one_wave = (int64_t)c->latency_resident * c->num_sms / 2;
    // Automatic choice: the build with the smallest estimated time for this launch size.  Measured on B200 at N = 2048
    // with the pair rotation (profiles/r2_pbs_sweep_pairs_w4.jsonl): one wave of the 8-CTA kernel 2.5 ms, one wave of
    // the latency build 4.3 ms, the throughput build 9.4 ms up to half a GPU of ciphertexts and 0.05 ms per ciphertext
    // beyond; the ratios carry over to the other sizes.
    int pick = c->pbs_mode;
    if (pick == 0) {
        const int64_t cap = one ? split_capacity<L>(c) : 0;
        const double t_split = cap > 0 ? (double)((total + cap - 1) / cap) : 1e30;
        const double t_lat = 1.72 * (double)((total + one_wave - 1) / one_wave);
        const double t_tp = std::max(3.76, 0.0199 * (double)total);
        pick = t_split <= t_lat && t_split <= t_tp ? 3 : (t_lat <= t_tp ? 1 : 2);
    }
    // A handful of ciphertexts: spread each over an 8-CTA cluster (4 CTAs per polynomial), lowest latency.
    if (one && pick >= 3) return launch_split<L>(c, a, total, st);
    const bool latency = pick == 1;
    a.bsk_hat = c->d_bsk[latency ? 1 : 0];
    const size_t sm = pbs_smem(c), sms = pbs_smem_staged(c);
    if (c->pairs) {
        a.bsk_hat = c->d_bskp[latency ? 1 : 0]; a.expo = c->d_expo[latency ? 1 : 0]; a.pw = c->d_pw;
        // A/B option: the three keys of each pair step staged by TMA (latency build, where the 6N words fit)
        if (latency && c->tma_stage_pairs && (int)pbs_smem_staged_pairs(c) <= kMaxSmem)
            pbs_cluster_kernel<L, EL, 1, true, true, true><<<grid, NttCfg<L, EL>::T, pbs_smem_staged_pairs(c), st>>>(a);
        else if (latency) pbs_cluster_kernel<L, EL, 1, true, false, true><<<grid, NttCfg<L, EL>::T, sm, st>>>(a);
        else pbs_cluster_kernel<L, ET, TP, true, false, true><<<grid, NttCfg<L, ET>::T, sm, st>>>(a);
        c->launches++;
        CK(cudaGetLastError());
        return BMI_OK;
    }
    // GGSW rows staged by TMA whenever the extra 2N words still leave room for the CTAs per SM the build is sized for
    const bool stage = one && c->tma_stage && (latency ? (int)sms <= kMaxSmem : (int)sms * TP <= kMaxSmem);
    if (latency && stage) pbs_cluster_kernel<L, EL, 1, true, true><<<grid, NttCfg<L, EL>::T, sms, st>>>(a);
    else if (latency && one) pbs_cluster_kernel<L, EL, 1, true, false><<<grid, NttCfg<L, EL>::T, sm, st>>>(a);
    else if (latency) pbs_cluster_kernel<L, EL, 1, false, false><<<grid, NttCfg<L, EL>::T, sm, st>>>(a);
    else if (stage) pbs_cluster_kernel<L, ET, TP, true, true><<<grid, NttCfg<L, ET>::T, sms, st>>>(a);
    else if (one) pbs_cluster_kernel<L, ET, TP, true, false><<<grid, NttCfg<L, ET>::T, sm, st>>>(a);
    else pbs_cluster_kernel<L, ET, TP, false, false><<<grid, NttCfg<L, ET>::T, sm, st>>>(a);
    c->launches++;
    CK(cudaGetLastError());
    return BMI_OK;
    }
}

template <int L>
int launch_polymul(bmi_ctx* c, const u64* a, const u64* b, u64* out, int count, cudaStream_t st) {
    if (c->pbs_mode == 3 || L > kMaxClusterL) {
        polymul_split_kernel<L><<<4 * count, SplitCfg<L>::T, 3 * SplitCfg<L>::M * 8, st>>>(a, b, out, c->d_tw, c->d_twi, c->ninv);
        c->launches++;
        CK(cudaGetLastError());
        return BMI_OK;
    }
    if constexpr (L <= kMaxClusterL) {
        constexpr int EL = latency_e<L>(), ET = throughput_e<L>();
        if (c->pbs_mode == 1) polymul_kernel<L, EL><<<count, NttCfg<L, EL>::T, (1 << L) * 8, st>>>(a, b, out, c->d_tw, c->d_twi, c->ninv);
        else polymul_kernel<L, ET><<<count, NttCfg<L, ET>::T, (1 << L) * 8, st>>>(a, b, out, c->d_tw, c->d_twi, c->ninv);
        c->launches++;
        CK(cudaGetLastError());
    }
    return BMI_OK;
}

