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
    // the same for entries 2.. from the unit's store offset  unit = (n F + f) C + c
    __host__ __device__ int64_t slab(int i) const { return 2 * (Nt * C) + (int64_t)(i - 2) * (Nt * F * C); }
    __host__ __device__ int64_t numel() const { return tensor_off(12); }
};

// per-step state kept on the device so that a captured CUDA graph can be replayed unchanged
struct StepState {
    unsigned long long step;  // SVI iteration counter: Philox stream + Adam bias correction
    // deferred AOI-local Adam (tq_cosmos_sites_adam): the gradients of the previous step are still to be applied, with
    // these bias-corrected constants.  Callers that never defer may keep passing an 8-byte state.
    unsigned int pending;
    float step_size, inv_sqrt_bc2;
};
constexpr int kStepStateBytes = 24;
static_assert(sizeof(StepState) == kStepStateBytes, "StepState layout is part of the C ABI (tq_sizeof_step_state)");

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
    // deferred Adam of the parameters a site's thread owns (site_fast_kernel; NULL = parameters are read-only here)
    T* adam_p;                   // == lparams, writable
    const T* adam_g;             // lgrads of the previous step
    T* adam_m;
    T* adam_v;
    float adam_b1, adam_b2, adam_eps;
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
    int64_t unit;                // (n F + f) C + c: offset of the unit in every (Nt, F, C) slab of the flat local buffers
    T p0, p1, pbm, pbs;
    unsigned long long rng_offset;
};

// A full-frame minibatch without index lists (every full-batch step; fb == F, AOIs 0 .. nb-1) has unit == u: no
// division, no index lists; only the background site needs the AOI (its prior's two per-AOI parameters).
template <typename T>
__device__ __forceinline__ SiteInputs<T> site_gather(const LocalArgs<T>& a, int s, uint32_t u32) {
    SiteInputs<T> in;
    int64_t aoi_c;               // n C + c (used by the background site only)
    if (a.v.ndx == nullptr && a.v.fdx == nullptr && a.v.fb == a.v.F) {
        in.unit = (int64_t)u32;
        aoi_c = 0;
        if (s == S_B) {
            const uint32_t C = (uint32_t)a.v.C, n = u32 / ((uint32_t)a.v.F * C);
            aoi_c = (int64_t)(n * C + (C == 1u ? 0u : u32 % C));
        }
    } else {
        const UnitIndex ui = locate_unit32(u32, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
        in.unit = ui.patch;
        aoi_c = (int64_t)ui.aoi * a.v.C + ui.c;
    }
    in.p0 = a.lparams[a.lo.slab(site_param0(s)) + in.unit];
    in.p1 = a.lparams[a.lo.slab(site_param1(s)) + in.unit];
    in.pbm = in.pbs = T(0);
    if (s == S_B) {
        in.pbm = a.lparams[aoi_c];                      // LP_BM, LP_BS: (Nt, C) tables at the front
        in.pbs = a.lparams[a.lo.Nt * a.lo.C + aoi_c];
    }
    const unsigned long long gid = (unsigned long long)a.aoi_offset * (unsigned long long)(a.v.F * a.v.C) + (unsigned long long)in.unit;
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
        const SpotPresence<T> sp((T)a.lparams[a.lo.slab(LP_M_PROBS + k) + in.unit], a.mc);
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

// Explicit fused multiply-adds: the update is evaluated in two kernels (adam_kernel, and site_fast_kernel for the
// deferred update of the parameters a site owns) and both must round identically, whatever the compiler would contract.
template <typename T>
__device__ __forceinline__ void adam_update(T& p, T g, T& m, T& v, T step_size, T inv_sqrt_bc2, T tb1, T tb2, T teps) {
    m = fma(g - m, T(1) - tb1, m);                       // exp_avg.lerp_(grad, 1 - beta1)
    v = fma((T(1) - tb2) * g, g, v * tb2);               // exp_avg_sq.mul_(b2).addcmul_(g, g, 1 - b2)
    const T denom = fma(Real<T>::sqrt(v), inv_sqrt_bc2, teps);
    p = fma(-step_size, m / denom, p);
}
// float: the square root and the reciprocal on the MUFU (2 ulp; the update they scale is 1e-3 of the parameter).  IEEE
// sqrt + division are ~20 instructions per element -- invisible in the bandwidth-bound adam_kernel, 6 % of the
// issue-bound site kernel when the update is deferred into it.
__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, float step_size, float inv_sqrt_bc2, float tb1,
                                            float tb2, float teps) {
    m = fmaf(g - m, 1.0f - tb1, m);
    v = fmaf((1.0f - tb2) * g, g, v * tb2);
    float sq, r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sq) : "f"(v));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(sq, inv_sqrt_bc2, teps)));
    p = fmaf(-step_size * m, r, p);
}

constexpr int kPostRed = NACC + 2;

}  // namespace tq
