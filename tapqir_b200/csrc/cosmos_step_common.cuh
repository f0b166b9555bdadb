// Structures and per-site plumbing shared by the kernels of one cosmos SVI step (cosmos_step.cu: one kernel per stage;
// cosmos_fused.cu: sites -> likelihood -> post in ONE persistent kernel).
#pragma once
#include "common.cuh"
#include "cosmos_globals.cuh"
#include "cosmos_sites_fast.cuh"
#include "cosmos_hmm.cuh"

namespace tq {

using Acc = double;  // arithmetic type of the per-unit local terms ("precision")

// flat local-parameter buffer (tapqir_b200/models/layout.py LocalLayout)
struct LocalOffsets {
    int64_t Nt, F, C;
    __host__ __device__ int64_t tensor_off(int t) const {
        const int64_t aoi = Nt * C, unit = Nt * F * C;
        if (t < 2) return t * aoi;
        if (t < 4) return 2 * aoi + (t - 2) * unit;
        return 2 * aoi + 2 * unit + (int64_t)(t - 4) * kK * unit;
    }
    // flat index of local record entry i (LP_* order) for (aoi n, frame f, channel c).  Entries 2.. are
    // (Nt, F, C) slabs in LP_* order -- the K slabs of a (K, Nt, F, C) tensor are consecutive entries -- so
    // one multiply-add serves them all
    __host__ __device__ int64_t index(int i, int64_t n, int64_t f, int64_t c) const {
        if (i < 2) return (int64_t)i * (Nt * C) + n * C + c;
        return 2 * (Nt * C) + (int64_t)(i - 2) * (Nt * F * C) + (n * F + f) * C + c;
    }
    __host__ __device__ int64_t numel() const { return tensor_off(12); }
};

// per-step state kept on the device so that a captured CUDA graph can be replayed unchanged
struct StepState {
    unsigned long long step;  // SVI iteration counter: Philox stream + Adam bias correction
};

constexpr int kLocalBlock = 128;

template <typename T> struct LocalArgs {
    tq_patch_view v;
    LocalOffsets lo;
    ModelConst mc;
    const T* lparams;
    const GlobalTables<double>* tables;
    int64_t U;
    int64_t aoi_offset;          // global index of this rank's AOI 0 (RNG stream identity)
    unsigned long long seed;
    const StepState* state;
    // sites
    const T* noise_in;           // (NSAMP, U) base variates or NULL -> Philox
    T* samples;                  // (NSAMP, U)
    T* qm;                       // (kM, U)
    T* rec;                      // (NREC, U) site records
    // post
    const T* L;                  // (kM, U)
    const T* gs;                 // (NSAMP, U) d/d sample from the likelihood kernel, S_* order
    const T* g_rate;             // (U,)
    double sN, sF;
    T* lgrads;                   // flat, LocalOffsets layout
    double* aoi_partial;         // (2, U): per-unit contributions to d/d(bm, bs)
    double* block_partial;       // (gridDim.x, C, NACC)
    // sites outside the fp32 forms: appended here by site_fast_kernel, redone in double by site_worklist_kernel
    uint32_t* worklist;          // (NSAMP * U) entries s * U + u, or NULL (block-local compaction instead)
    unsigned int* work_count;    // [0]: entries appended this launch
    // hmm variant (cosmos_hmm.cuh): the extra local slabs live behind the cosmos layout in the same flat buffers --
    // m_probs[z = 1] (K slabs of (Nt, F, C)), then z_trans (Nt, F, C, 2, 2); m_probs[z = 0] are the cosmos m_probs slabs
    const double* hmm_a;         // (kZ, U) forward marginals of the guide's chain
    T* hmm_v;                    // (kZ, U) centred emission values V_f(z)
    __host__ __device__ int64_t hmm_mprobs1(int k, int64_t n, int64_t f, int64_t c) const {
        return lo.numel() + (int64_t)k * (lo.Nt * lo.F * lo.C) + (n * lo.F + f) * lo.C + c;
    }
    __host__ __device__ int64_t hmm_ztrans(int64_t n, int64_t f, int64_t c) const {   // + z' * 2 + z
        return lo.numel() + (int64_t)kK * (lo.Nt * lo.F * lo.C) + ((n * lo.F + f) * lo.C + c) * (kZ * kZ);
    }
};

// ---- sites: one thread per (site, unit), site-major so that a warp evaluates one family ------------------
// grid = (blocks over units, site): no division to find the site, 32-bit index arithmetic (the host checks
// U < 2^31).  Three kernels share the gather / scatter below:
//   site_kernel<T>           the double-precision form for every site (dtype "double")
//   site_fast_kernel         production: fp32 forms of cosmos_sites_fast.cuh; a site outside their regimes
//                            leaves a NaN in its log-q slot ...
//   site_fallback_kernel     ... and is redone here in double (same Philox stream, so the same draw whichever
//                            kernel makes it).  Keeping the two apart holds the hot kernel at 66 registers.
template <typename T> struct SiteInputs {
    UnitIndex ui;
    int64_t f;
    T p0, p1, pbm, pbs;
    unsigned long long rng_offset;
};

template <typename T>
__device__ __forceinline__ SiteInputs<T> site_gather(const LocalArgs<T>& a, int s, uint32_t u32) {
    SiteInputs<T> in;
    in.ui = locate_unit32(u32, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
    in.f = a.v.fdx ? a.v.fdx[in.ui.fi] : in.ui.fi;
    in.p0 = a.lparams[a.lo.index(site_param0(s), in.ui.aoi, in.f, in.ui.c)];
    in.p1 = a.lparams[a.lo.index(site_param1(s), in.ui.aoi, in.f, in.ui.c)];
    in.pbm = in.pbs = T(0);
    if (s == S_B) {
        in.pbm = a.lparams[a.lo.index(LP_BM, in.ui.aoi, in.f, in.ui.c)];
        in.pbs = a.lparams[a.lo.index(LP_BS, in.ui.aoi, in.f, in.ui.c)];
    }
    const unsigned long long gid = (((unsigned long long)(a.aoi_offset + in.ui.aoi)) * a.v.F + in.f) * a.v.C + in.ui.c;
    in.rng_offset = ((gid + 1ull) << 12) + ((unsigned long long)s << 8);
    return in;
}

// The reference clamps a sample into its support with the margins of ITS dtype (double: tiny for Gamma.rsample,
// eps * scale for pyro's AffineBeta.rsample).  Stored as float those margins round away -- the sample would sit exactly
// ON the bound (height 0, x = -(P+1)/2) and the model's log-densities there are infinite -- so the float sample gets the
// float-sized margin.  Only samples that were on the clamp anyway are touched.
template <typename T>
__device__ __forceinline__ T sample_into_support(int s, T v, const ModelConst& mc) {
    if (sizeof(T) != sizeof(float)) return v;
    if (site_is_gamma(s)) return v > T(1.17549435e-38f) ? v : T(1.17549435e-38f);
    const T lo = s < S_X ? (T)mc.width_min : T(-0.5) * T(mc.P + 1), hi = s < S_X ? (T)mc.width_max : T(0.5) * T(mc.P + 1);
    const T margin = (hi - lo) * T(1.1920929e-7f);
    return v < lo + margin ? lo + margin : (v > hi - margin ? hi - margin : v);
}

template <typename T>
__device__ __forceinline__ void site_scatter(const LocalArgs<T>& a, int s, int64_t u, T v, const T* rec, const T* extra) {
    a.samples[(int64_t)s * a.U + u] = sample_into_support(s, v, a.mc);
#pragma unroll
    for (int j = 0; j < NSO; ++j) a.rec[((int64_t)s * NSO + j) * a.U + u] = rec[j];
    if (s == S_B) {
#pragma unroll
        for (int j = 0; j < NEX; ++j) a.rec[((int64_t)NSAMP * NSO + j) * a.U + u] = extra[j];
    }
}

// weights of the likelihood kernel: q(m) from the unconstrained m_probs (written by the background site's thread)
template <typename T>
__device__ __forceinline__ void write_presence_weights(const LocalArgs<T>& a, const SiteInputs<T>& in, int64_t u) {
    T q1[kK], q0[kK], qm[kM];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const SpotPresence<T> sp((T)a.lparams[a.lo.index(LP_M_PROBS + k, in.ui.aoi, in.f, in.ui.c)], a.mc);
        q1[k] = sp.q1; q0[k] = sp.q0;
    }
    presence_weights<T>(q1, q0, qm);
#pragma unroll
    for (int m = 0; m < kM; ++m) a.qm[m * a.U + u] = qm[m];
}

template <typename T>
__device__ __forceinline__ void site_double(const LocalArgs<T>& a, int s, uint32_t u32) {
    const SiteInputs<T> in = site_gather(a, s, u32);
    const bool use_rng = a.noise_in == nullptr;
    Philox rng(a.seed, a.state->step, in.rng_offset);
    double variate = use_rng ? 0.0 : (double)a.noise_in[(int64_t)s * a.U + u32];
    double drec[NSO], dextra[NEX];
    const T v = (T)site_eval(s, (double)in.p0, (double)in.p1, (double)in.pbm, (double)in.pbs, a.mc, use_rng, &rng, variate, drec, dextra);
    T rec[NSO], extra[NEX];
#pragma unroll
    for (int j = 0; j < NSO; ++j) rec[j] = (T)drec[j];
#pragma unroll
    for (int j = 0; j < NEX; ++j) extra[j] = (T)dextra[j];
    site_scatter(a, s, (int64_t)u32, v, rec, extra);
}

constexpr int kPostRed = NACC + 2;

}  // namespace tq
