// Kernels of one cosmos SVI step around the likelihood kernel (ksmogn.cu):
//
//   globals_sample -> local_pre -> [tq_ksmogn_fwd_bwd] -> local_post -> reduce -> [allreduce]
//   -> globals_grad -> adam
//
// Replaces, for this path, what Pyro's SVI.step does around models/cosmos.py:82-462 (guide/model
// execution, TraceEnum_ELBO contraction, autograd backward; models/model.py:212) -- SURVEY.md
// App. A.  The arithmetic is in cosmos_local.cuh / cosmos_globals.cuh (also compiled for the CPU by
// tests/hostcheck); this file is gather/scatter, reductions and launch plumbing.
#include "cosmos_step_common.cuh"

namespace tq {

// ---- globals: sample + tables; one block per global site --------------------------------------------------
// (the global parameters, their gradients and Adam moments are ALWAYS float64 buffers, whatever `dtype` the AOI-local
// buffers use: a dozen scalars, and the concentration-like ones (gain_beta, ...) have gradients that cancel ~1e3-fold in
// the base variate, so rounding their parameters to fp32 alone moves those gradients by 1e-4; T = type of gain_out)
template <typename T>
__global__ void globals_sample_kernel(const double* __restrict__ gparams, int Q, bool hmm, ModelConst mc,
                                      const double* __restrict__ noise_in, unsigned long long seed,
                                      const StepState* __restrict__ state, double* __restrict__ gstate,
                                      GlobalTables<double>* __restrict__ tables, T* __restrict__ gain_out) {
    // one BLOCK per site: lanes of one warp would serialise the divergent per-site code paths
    const int site = blockIdx.x;
    if (threadIdx.x != 0 || site >= global_site_count(Q, hmm)) return;
    GlobalLayout gl{Q, hmm};
    double u[kMaxGlobals];
    for (int i = 0; i < gl.count(); ++i) u[i] = (double)gparams[i];
    const bool use_rng = noise_in == nullptr;
    Philox rng(seed, state->step, (unsigned long long)site << 8);
    // every lane works on the shared device buffers directly: the entries of different sites are disjoint
    double* variate = gstate;
    double* sample = gstate + kMaxGlobalNoise;
    if (!use_rng) {
        if (site == 0) variate[gl.n_gain()] = noise_in[gl.n_gain()];
        else if (site == 1) variate[gl.n_prox()] = noise_in[gl.n_prox()];
        else if (site < 2 + Q) { for (int z = 0; z < kZ; ++z) variate[gl.n_pi(site - 2, z)] = noise_in[gl.n_pi(site - 2, z)]; }
        else if (site < 2 + 2 * Q) variate[gl.n_lamda(site - 2 - Q)] = noise_in[gl.n_lamda(site - 2 - Q)];
        else {
            const int idx = site - 2 - 2 * Q;
            for (int z = 0; z < kZ; ++z) variate[gl.n_trans(idx / kZ, idx % kZ, z)] = noise_in[gl.n_trans(idx / kZ, idx % kZ, z)];
        }
    }
    globals_pre_site(site, u, gl, mc, use_rng, &rng, variate, sample, *tables);
    if (site == 0) gain_out[0] = (T)tables->gain;
}

template <typename T>
__global__ void __launch_bounds__(kLocalBlock) site_kernel(const LocalArgs<T> a) {
    const int s = blockIdx.y;
    const uint32_t u32 = blockIdx.x * (uint32_t)kLocalBlock + threadIdx.x;
    if (u32 >= (uint32_t)a.U) return;
    site_double(a, s, u32);
    if (s == S_B) write_presence_weights(a, site_gather(a, s, u32), (int64_t)u32);
}

// Two passes per block.  Pass 1: every thread draws its site's sample and classifies it; the bulk regime (what every site is
// in at the initial point) is evaluated on the spot, the rest -- small concentrations of absent spots' guides, draws in the
// tails, ... -- are compacted by regime class into shared memory.  Pass 2: the block's first threads replay those sites,
// neighbours in the same regime.  (Without the compaction a trained model ran this kernel at 13.6 of 32 lanes per
// instruction: every warp held a few sites of every regime and walked through all of them.)
__device__ __forceinline__ void site_fast_finish(const LocalArgs<float>& a, int s, uint32_t u32, int status, float v, const float* rec,
                                                 const float* extra) {
    if (status == SITE_DONE) site_scatter(a, s, (int64_t)u32, v, rec, extra);
    else if (a.worklist) a.worklist[atomicAdd(a.work_count, 1u)] = (uint32_t)s * (uint32_t)a.U + u32;
    else a.rec[((int64_t)s * NSO + SO_LQ) * a.U + u32] = nanf("");   // marker for site_fallback_kernel
}

// Deferred Adam (tq_cosmos_sites_adam; full-batch steps).  The dense Adam over the AOI-local buffer is pure HBM traffic
// (28 B per element, 0.39 ms at 1000 AOIs x 5000 frames) while this kernel is issue-bound with the memory system at 13 %:
// the thread of site s applies the previous step's update to the parameters it is about to read -- its own two (x and y
// share `size`, which therefore stays with tq_adam_dense at the end of the step, as do the per-AOI background
// parameters) and, on the background site, the spot-presence logits it turns into q(m).  Same arithmetic, same constants
// (StepState), hence the same bits as adam_kernel.
struct DeferredAdamLoads {   // gradient and moments of the (up to) two parameters a site owns
    float g0, m0, v0, g1, m1, v1;
};
__device__ __forceinline__ float deferred_adam_at(const LocalArgs<float>& a, int64_t i, float p, float g, float m, float v, float ss, float isb) {
    adam_update(p, g, m, v, ss, isb, a.adam_b1, a.adam_b2, a.adam_eps);
    a.adam_p[i] = p; a.adam_m[i] = m; a.adam_v[i] = v;
    return p;
}
// issue the loads ...
__device__ __forceinline__ DeferredAdamLoads site_deferred_adam_load(const LocalArgs<float>& a, int s, const SiteInputs<float>& in) {
    DeferredAdamLoads l;
    const int64_t i0 = a.lo.slab(site_param0(s)) + in.unit;
    l.g0 = a.adam_g[i0]; l.m0 = a.adam_m[i0]; l.v0 = a.adam_v[i0];
    l.g1 = l.m1 = l.v1 = 0.0f;
    if (s < S_X) {
        const int64_t i1 = a.lo.slab(site_param1(s)) + in.unit;
        l.g1 = a.adam_g[i1]; l.m1 = a.adam_m[i1]; l.v1 = a.adam_v[i1];
    }
    return l;
}
// ... and use them after the (parameter-independent) Philox block of the site's first trials has been generated
__device__ __forceinline__ void site_deferred_adam(const LocalArgs<float>& a, int s, SiteInputs<float>& in, const DeferredAdamLoads& l) {
    const float ss = a.state->step_size, isb = a.state->inv_sqrt_bc2;   // (re-read per unit: two registers less across the loop)
    in.p0 = deferred_adam_at(a, a.lo.slab(site_param0(s)) + in.unit, in.p0, l.g0, l.m0, l.v0, ss, isb);
    if (s < S_X) in.p1 = deferred_adam_at(a, a.lo.slab(site_param1(s)) + in.unit, in.p1, l.g1, l.m1, l.v1, ss, isb);
    if (s == S_B) {
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const int64_t i = a.lo.slab(LP_M_PROBS + k) + in.unit;
            deferred_adam_at(a, i, a.adam_p[i], a.adam_g[i], a.adam_m[i], a.adam_v[i], ss, isb);
        }
    }
}

// (8 resident blocks per SM = 64 registers, no spills: 5 us faster on a trained model than 6 blocks at 71 registers)
// A block takes kSiteUPT x 128 consecutive units of one site: with 128 a trained model's deferred sites (a quarter of
// them, four regime classes) left each class with a handful of sites per block -- partially filled warps in pass 2
// (20 of 32 lanes per instruction over the kernel); four times the units fill them.
// Measured (B200, profiles/r2_bulk_ab.sh): C3 (5 M units) step 6.48 -> 6.40 ms, after 1000 iterations 7.09 -> 6.95 ms;
// at C2 (100 k units) the four-times-smaller grid is 1.5 waves and 5 % slower, so small launches keep one unit per thread.
template <int kSiteUPT>
__global__ void __launch_bounds__(kLocalBlock, 8) site_fast_kernel(const LocalArgs<float> a) {
    constexpr int kSiteSpan = kLocalBlock * kSiteUPT;
    __shared__ unsigned int cnt[kSiteClasses], off[kSiteClasses + 1], fill[kSiteClasses], n_def;
    __shared__ double t_var[kSiteSpan], s_var[kSiteSpan];
    __shared__ unsigned short t_idx[kSiteSpan], s_idx[kSiteSpan];
    __shared__ unsigned char t_cls[kSiteSpan];
    const int s = blockIdx.y;
    const uint32_t base = blockIdx.x * (uint32_t)kSiteSpan;
    static_assert(kSiteClasses <= kLocalBlock, "one thread per class counter");
    if (threadIdx.x < kSiteClasses) { cnt[threadIdx.x] = 0u; fill[threadIdx.x] = 0u; }
    if (threadIdx.x == 0) n_def = 0u;
    __syncthreads();
    const bool use_rng = a.noise_in == nullptr;
    // the previous step's Adam update of the parameters this thread is about to read (pass 2 and the double-precision
    // kernels behind this one re-read them from memory, updated)
    const bool adam_pending = a.adam_p != nullptr && a.state->pending != 0u;
    bool deferred = false;
#pragma unroll 1
    for (int j = 0; j < kSiteUPT; ++j) {
        const uint32_t local = (uint32_t)j * kLocalBlock + threadIdx.x, u32 = base + local;
        if (u32 >= (uint32_t)a.U) break;
        SiteInputs<float> in = site_gather(a, s, u32);
        DeferredAdamLoads al;
        if (adam_pending) al = site_deferred_adam_load(a, s, in);
        Philox rng(a.seed, a.state->step, in.rng_offset);
        GammaTrials trials;
        // ~100 instructions that need no parameter: issued under the loads above (C3 on a B200: 1643 -> 1613 us; an explicit
        // prefetch.global.L2 of the thread's next unit on top of it: 1670 us -- the kernel is issue-bound)
        if (use_rng) trials.preload(rng);
        if (adam_pending) site_deferred_adam(a, s, in, al);
        double variate = use_rng ? 0.0 : (double)a.noise_in[(int64_t)s * a.U + u32];
        float v = 0.0f, rec[NSO], extra[NEX];
        int cls = 0;
        const int status = site_eval_fast_t<1>(s, in.p0, in.p1, in.pbm, in.pbs, a.mc, use_rng, &rng, trials, variate, v, rec, extra, cls);
        if (status == SITE_DEFER) {
            deferred = true;
            const unsigned int pos = atomicAdd(&n_def, 1u);
            atomicAdd(&cnt[cls], 1u);
            t_var[pos] = variate;
            t_idx[pos] = (unsigned short)local;
            t_cls[pos] = (unsigned char)cls;
        } else {
            site_fast_finish(a, s, u32, status, v, rec, extra);
        }
        if (s == S_B) write_presence_weights(a, in, (int64_t)u32);
    }
    if (!__syncthreads_or(deferred)) return;   // every site was in the bulk regime (always so at the initial point)
    if (threadIdx.x == 0) {
        unsigned int acc = 0u;
        for (int c = 0; c < kSiteClasses; ++c) { off[c] = acc; acc += cnt[c]; }
        off[kSiteClasses] = acc;
    }
    __syncthreads();
    const unsigned int total = off[kSiteClasses];
    for (unsigned int i = threadIdx.x; i < total; i += kLocalBlock) {   // counting sort by regime class
        const int c = t_cls[i];
        const unsigned int pos = off[c] + atomicAdd(&fill[c], 1u);
        s_var[pos] = t_var[i];
        s_idx[pos] = t_idx[i];
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < total; i += kLocalBlock) {
        const uint32_t u = base + s_idx[i];
        const SiteInputs<float> in = site_gather(a, s, u);
        double var = s_var[i];
        float v = 0.0f, rec[NSO], extra[NEX];
        int cls = 0;
        const int status = site_eval_fast_t<2>(s, in.p0, in.p1, in.pbm, in.pbs, a.mc, false, nullptr, var, v, rec, extra, cls);
        site_fast_finish(a, s, u, status, v, rec, extra);
    }
}

// ---- the two passes as two launches (TQ_SITE_SPLIT=1; profiles/r2s2_sites_icache.md) ------------------------------------------
// On a trained model the kernel above is instruction-fetch bound: 106 KB of SASS, and an SM's eight blocks are spread over the
// sampler, the bulk forms (pass 1) and the regime classes of the replay (pass 2).  Split, every SM runs pass-1 code OR pass-2 code.
// A deferred site parks its base variate (8 bytes) and its class in its OWN, not yet written, rows of the record buffer
// -- no list, no extra memory -- under a NaN marker with a payload; the second launch (a resident wave of blocks walking
// the spans of the first) finds the parked sites of a span, sorts them by class in shared memory and replays them exactly as
// pass 2 above does.  A launch with nothing deferred costs its blocks one read of a counter.
constexpr unsigned int kParkMarker = 0x7FD00000u;          // quiet NaN, payload = regime class in the low byte
__device__ unsigned int g_site_deferred[8];                 // sites deferred by the current launch, per stream slot

template <int kSiteUPT>
__global__ void __launch_bounds__(kLocalBlock, 8) site_pass1_kernel(const LocalArgs<float> a, unsigned int* __restrict__ n_deferred) {
    constexpr int kSiteSpan = kLocalBlock * kSiteUPT;
    const int s = blockIdx.y;
    const uint32_t base = blockIdx.x * (uint32_t)kSiteSpan;
    const bool use_rng = a.noise_in == nullptr;
    const bool adam_pending = a.adam_p != nullptr && a.state->pending != 0u;
    unsigned int parked = 0u;
#pragma unroll 1
    for (int j = 0; j < kSiteUPT; ++j) {
        const uint32_t u32 = base + (uint32_t)j * kLocalBlock + threadIdx.x;
        if (u32 >= (uint32_t)a.U) break;
        SiteInputs<float> in = site_gather(a, s, u32);
        DeferredAdamLoads al;
        if (adam_pending) al = site_deferred_adam_load(a, s, in);
        Philox rng(a.seed, a.state->step, in.rng_offset);
        GammaTrials trials;
        if (use_rng) trials.preload(rng);
        if (adam_pending) site_deferred_adam(a, s, in, al);
        double variate = use_rng ? 0.0 : (double)a.noise_in[(int64_t)s * a.U + u32];
        float v = 0.0f, rec[NSO], extra[NEX];
        int cls = 0;
        const int status = site_eval_fast_t<1>(s, in.p0, in.p1, in.pbm, in.pbs, a.mc, use_rng, &rng, trials, variate, v, rec, extra, cls);
        if (status == SITE_DEFER) {
            const long long bits = __double_as_longlong(variate);
            float* r = a.rec + (int64_t)s * NSO * a.U + u32;
            r[(int64_t)SO_A0 * a.U] = __int_as_float((int)(bits & 0xffffffffll));
            r[(int64_t)SO_B0 * a.U] = __int_as_float((int)(bits >> 32));
            r[(int64_t)SO_LQ * a.U] = __uint_as_float(kParkMarker | (unsigned int)cls);
            ++parked;
        } else {
            site_fast_finish(a, s, u32, status, v, rec, extra);
        }
        if (s == S_B) write_presence_weights(a, in, (int64_t)u32);
    }
    const unsigned int warp_parked = __reduce_add_sync(0xffffffffu, parked);
    if ((threadIdx.x & 31) == 0 && warp_parked) atomicAdd(n_deferred, warp_parked);
}

template <int kSiteUPT>
__global__ void __launch_bounds__(kLocalBlock, 8) site_pass2_kernel(const LocalArgs<float> a, const unsigned int* __restrict__ n_deferred,
                                                                    unsigned int spans_per_site) {
    constexpr int kSiteSpan = kLocalBlock * kSiteUPT;
    if (*n_deferred == 0u) return;
    __shared__ unsigned int cnt[kSiteClasses], off[kSiteClasses + 1], fill[kSiteClasses], n_def;
    __shared__ double t_var[kSiteSpan], s_var[kSiteSpan];
    __shared__ unsigned short t_idx[kSiteSpan], s_idx[kSiteSpan];
    __shared__ unsigned char t_cls[kSiteSpan];
    const unsigned int n_spans = spans_per_site * NSAMP;
    for (unsigned int span = blockIdx.x; span < n_spans; span += gridDim.x) {
        const int s = (int)(span / spans_per_site);
        const uint32_t base = (span - (unsigned int)s * spans_per_site) * (uint32_t)kSiteSpan;
        if (threadIdx.x < kSiteClasses) { cnt[threadIdx.x] = 0u; fill[threadIdx.x] = 0u; }
        if (threadIdx.x == 0) n_def = 0u;
        __syncthreads();
        const float* r = a.rec + (int64_t)s * NSO * a.U;
        bool any = false;
#pragma unroll
        for (int j = 0; j < kSiteUPT; ++j) {
            const uint32_t local = (uint32_t)j * kLocalBlock + threadIdx.x, u32 = base + local;
            if (u32 < (uint32_t)a.U) {
                const unsigned int m = __float_as_uint(r[(int64_t)SO_LQ * a.U + u32]);
                if ((m & 0xffffff00u) == kParkMarker) {
                    const unsigned int lo = __float_as_uint(r[(int64_t)SO_A0 * a.U + u32]), hi = __float_as_uint(r[(int64_t)SO_B0 * a.U + u32]);
                    const int cls = (int)(m & 0xffu) < kSiteClasses ? (int)(m & 0xffu) : 0;
                    const unsigned int pos = atomicAdd(&n_def, 1u);
                    atomicAdd(&cnt[cls], 1u);
                    t_var[pos] = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
                    t_idx[pos] = (unsigned short)local;
                    t_cls[pos] = (unsigned char)cls;
                    any = true;
                }
            }
        }
        if (__syncthreads_or(any)) {
            if (threadIdx.x == 0) {
                unsigned int acc = 0u;
                for (int c = 0; c < kSiteClasses; ++c) { off[c] = acc; acc += cnt[c]; }
                off[kSiteClasses] = acc;
            }
            __syncthreads();
            const unsigned int total = off[kSiteClasses];
            for (unsigned int i = threadIdx.x; i < total; i += kLocalBlock) {   // counting sort by regime class
                const int c = t_cls[i];
                const unsigned int pos = off[c] + atomicAdd(&fill[c], 1u);
                s_var[pos] = t_var[i];
                s_idx[pos] = t_idx[i];
            }
            __syncthreads();
            for (unsigned int i = threadIdx.x; i < total; i += kLocalBlock) {
                const uint32_t u = base + s_idx[i];
                const SiteInputs<float> in = site_gather(a, s, u);
                double var = s_var[i];
                float v = 0.0f, rec[NSO], extra[NEX];
                int cls = 0;
                const int status = site_eval_fast_t<2>(s, in.p0, in.p1, in.pbm, in.pbs, a.mc, false, nullptr, var, v, rec, extra, cls);
                if (status == SITE_DONE) site_scatter(a, s, (int64_t)u, v, rec, extra);
                else {
                    // leaves the fp32 forms after all: to the double worklist, and the marker must not be found again
                    a.rec[((int64_t)s * NSO + SO_LQ) * a.U + u] = 0.0f;
                    a.worklist[atomicAdd(a.work_count, 1u)] = (uint32_t)s * (uint32_t)a.U + u;
                }
            }
        }
        __syncthreads();   // the shared lists are reused by the next span
    }
}

// A block scans kFallbackUPT * 128 markers of one site, compacts the hits into shared memory and then works through
// them with dense warps: a trained model leaves the fp32 forms at a percent or so of its sites (draws in the far tail,
// Rice expansion at small concentrations), and scattered over the warps every one of them would drag 31 idle lanes
// through the ~3000-instruction double form.
constexpr int kFallbackUPT = 16;
__global__ void __launch_bounds__(kLocalBlock) site_fallback_kernel(const LocalArgs<float> a) {
    __shared__ uint32_t hits[kLocalBlock * kFallbackUPT];
    __shared__ unsigned int n_hits;
    const int s = blockIdx.y;
    if (threadIdx.x == 0) n_hits = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * (uint32_t)(kLocalBlock * kFallbackUPT);
    const float* marker = a.rec + ((int64_t)s * NSO + SO_LQ) * a.U;
#pragma unroll 4
    for (int j = 0; j < kFallbackUPT; ++j) {
        const uint32_t u32 = base + j * kLocalBlock + threadIdx.x;
        if (u32 < (uint32_t)a.U) {
            const float m = marker[u32];
            if (m != m) hits[atomicAdd(&n_hits, 1u)] = u32;
        }
    }
    __syncthreads();
    for (unsigned int i = threadIdx.x; i < n_hits; i += kLocalBlock) site_double(a, s, hits[i]);
}

// The same with a device-wide worklist (tq_cosmos_sites_ws): every thread of the fixed-size grid takes entries, so the
// double form runs in full warps at full occupancy whatever the hit rate (block-local compaction above: 75 us for the
// 4.5 % of sites a trained C2 model sends here; this: the cost of the work itself).
__global__ void __launch_bounds__(kLocalBlock) site_worklist_kernel(const LocalArgs<float> a) {
    const unsigned int n = *a.work_count;
    const uint32_t U = (uint32_t)a.U;
    for (unsigned int i = blockIdx.x * kLocalBlock + threadIdx.x; i < n; i += gridDim.x * kLocalBlock) {
        const uint32_t e = a.worklist[i];
        const uint32_t s = e / U;
        site_double(a, (int)s, e - s * U);
    }
}

// ---- post: blocks per (AOI, channel) chunk of frames, kPostUPT units per thread; the cross-unit sums are fused in ----
// Every block reduces its NACC channel accumulators and the two AOI-level gradient sums in a fixed order into
// block_partial; the last block of an (AOI, channel) to finish (atomic ticket) adds that AOI's partials in index
// order and writes d loss / d (background_mean_loc, background_std_loc); the last block of the launch does the
// same for the channel accumulators.  Fixed summation order => run-to-run deterministic, whichever block is last.
// `tickets`: (nb * C + 1) zero-initialised counters, left zeroed for the next launch.

#ifndef TQ_POST_UPT_MIN
#define TQ_POST_UPT_MIN 888   // two waves of 3 blocks per SM (B200: 125 AOIs x 5000 frames, one rank of 8: 105 -> 74 us)
#endif
// units per thread: 4 amortises the block reduction when there are enough blocks to fill the GPU, else 1
__host__ __device__ inline int post_upt(int nb, int fb, int C) {
    return (int64_t)nb * C * ((fb + kLocalBlock * 4 - 1) / (kLocalBlock * 4)) >= TQ_POST_UPT_MIN ? 4 : 1;
}
__host__ __device__ inline int post_chunks(int fb, int upt) { return (fb + kLocalBlock * upt - 1) / (kLocalBlock * upt); }

template <typename T, int kPostUPT, bool HMM = false>
#ifndef TQ_POST_MINB
#define TQ_POST_MINB 3
#endif
__global__ void __launch_bounds__(kLocalBlock, TQ_POST_MINB) local_post_kernel(const LocalArgs<T> a, int chunks, unsigned int* __restrict__ tickets,
                                                                double* __restrict__ acc_out) {
    __shared__ double red[kLocalBlock / 32][kPostRed];
    __shared__ GlobalTables<T> gt;
    __shared__ int last_flags[2];
    if (threadIdx.x == 0) gt.convert_from(*a.tables);
    __syncthreads();
    const int fb = a.v.fb, C = a.v.C;
    const int chunk = blockIdx.x % chunks;
    const int nc = blockIdx.x / chunks;          // (ni, c) pair
    const int c = nc % C, ni = nc / C;
    const int64_t n = a.v.ndx ? a.v.ndx[ni] : ni;
    const double mu = a.v.mask[n] ? 1.0 : 0.0;
    const bool ontarget = a.v.is_ontarget[n] != 0;
    const T u_bm = a.lparams[a.lo.index(LP_BM, n, 0, c)], u_bs = a.lparams[a.lo.index(LP_BS, n, 0, c)];
    const T scale = (T)(-a.sN * a.sF * mu);  // loss = -ELBO
    T sum[kPostRed];   // this thread's kPostUPT units in T (a handful of terms); everything across threads in double
#pragma unroll
    for (int i = 0; i < kPostRed; ++i) sum[i] = T(0);
#pragma unroll 1
    for (int t = 0; t < kPostUPT; ++t) {
        const int fi = (chunk * kPostUPT + t) * kLocalBlock + threadIdx.x;
        if (fi >= fb) break;
        const int64_t u = ((int64_t)ni * fb + fi) * C + c;
        const int64_t f = a.v.fdx ? a.v.fdx[fi] : fi;
        T rec[NREC], sample[NSAMP], gs[NSAMP], L[kM], u_mp[kK];
        {   // SoA rows are U apart: running pointers instead of a 64-bit multiply per load
            const T* pr = a.rec + u;
#pragma unroll
            for (int i = 0; i < NREC; ++i, pr += a.U) rec[i] = *pr;
            const T* ps = a.samples + u;
            const T* pg = a.gs + u;
#pragma unroll
            for (int i = 0; i < NSAMP; ++i, ps += a.U, pg += a.U) { sample[i] = *ps; gs[i] = *pg; }
            const T* pl = a.L + u;
#pragma unroll
            for (int m = 0; m < kM; ++m, pl += a.U) L[m] = *pl;
        }
#pragma unroll
        for (int k = 0; k < kK; ++k) u_mp[k] = a.lparams[a.lo.index(LP_M_PROBS + k, n, f, c)];
        UnitGrads<T> ug;
        if (HMM) {
            T ump2[kZ][kK], az[kZ];
            HmmUnitOut<T> ho;
#pragma unroll
            for (int k = 0; k < kK; ++k) { ump2[0][k] = u_mp[k]; ump2[1][k] = a.lparams[a.hmm_mprobs1(k, n, f, c)]; }
#pragma unroll
            for (int z = 0; z < kZ; ++z) az[z] = (T)a.hmm_a[(int64_t)z * a.U + u];
            unit_post_hmm<T>(rec, sample, L, gs, a.g_rate[u], ump2, az, u_bm, u_bs, a.mc, gt, c, fi == 0, ug, ho);
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                ug.g[LP_M_PROBS + k] = ho.gmp[0][k];
                a.lgrads[a.hmm_mprobs1(k, n, f, c)] = scale * ho.gmp[1][k];
            }
#pragma unroll
            for (int z = 0; z < kZ; ++z) a.hmm_v[(int64_t)z * a.U + u] = ho.v[z];
        } else {
            unit_post<T>(rec, sample, L, gs, a.g_rate[u], u_mp, u_bm, u_bs, a.mc, gt, c, ontarget, fi == 0, ug);
        }
#pragma unroll
        for (int i = LP_B_LOC; i < NLOCAL; ++i) a.lgrads[a.lo.index(i, n, f, c)] = scale * ug.g[i];
#pragma unroll
        for (int i = 0; i < NACC; ++i) sum[i] += ug.acc[i];
        sum[NACC] += ug.g[LP_BM];
        sum[NACC + 1] += ug.g[LP_BS];
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < kPostRed; ++i) {
        const double v = warp_sum((double)sum[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < kPostRed) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kLocalBlock / 32; ++w) v += red[w][threadIdx.x];
        a.block_partial[(int64_t)blockIdx.x * kPostRed + threadIdx.x] = mu * v;
    }
    // publish: the barrier orders the writers before thread 0, whose (cumulative) fence + ticket make them visible
    // device-wide -- the pattern of a cooperative-groups grid barrier.  (Fencing in every thread instead makes each
    // warp wait for its own gradient stores to drain.)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t1 = atomicAdd(&tickets[nc], 1u);
        last_flags[0] = (t1 == (unsigned int)chunks - 1u);
        if (last_flags[0]) { tickets[nc] = 0u; __threadfence(); }
    }
    __syncthreads();
    if (!last_flags[0]) return;

    // ---- last block of this (AOI, channel): add its chunks in index order --------------------------------------------
    const int n_nc = a.v.nb * C;
    double* aoi_sums = a.block_partial + (int64_t)n_nc * chunks * kPostRed;   // (nb * C, kPostRed), after the chunk partials
    if (threadIdx.x < kPostRed) {
        const double* bp = a.block_partial + (int64_t)nc * chunks * kPostRed + threadIdx.x;
        double v = 0.0;
        for (int k0 = 0; k0 < chunks; k0 += 8) {
            double t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = __ldcg(bp + (int64_t)min(k0 + j, chunks - 1) * kPostRed);   // loads first ...
#pragma unroll
            for (int j = 0; j < 8; ++j) v += (k0 + j < chunks) ? t[j] : 0.0;                                 // ... then a fixed-order sum
        }
        aoi_sums[(int64_t)nc * kPostRed + threadIdx.x] = v;
        if (threadIdx.x >= NACC) {
            // d loss / d (background_mean_loc, background_std_loc)[n, 0, c]: its frames + the AOI-level prior
            double pbm, pbs;
            aoi_prior_grad((double)u_bm, (double)u_bs, a.mc, pbm, pbs);
            const bool is_bm = threadIdx.x == NACC;
            a.lgrads[a.lo.index(is_bm ? LP_BM : LP_BS, n, 0, c)] = (T)(-(a.sN * a.sF * v + a.sN * mu * (is_bm ? pbm : pbs)));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t2 = atomicAdd(&tickets[n_nc], 1u);
        last_flags[1] = (t2 == (unsigned int)n_nc - 1u);
        if (last_flags[1]) { tickets[n_nc] = 0u; __threadfence(); }
    }
    __syncthreads();
    if (!last_flags[1]) return;

    // ---- last block of the launch: channel accumulators = sum over AOIs, in index order ---------------------------------
    // thread (cc, i, part) adds every kParts-th AOI of channel cc, loads batched eight at a time; the kParts partial sums
    // are then combined in a fixed order
    constexpr int kParts = 4;
    __shared__ double fin[kMaxC * NACC][kParts];
    for (int w = threadIdx.x; w < C * NACC * kParts; w += kLocalBlock) {
        const int part = w % kParts, vi = w / kParts;
        const int cc = vi / NACC, i = vi - cc * NACC;
        double v = 0.0;
        for (int q0 = part; q0 < a.v.nb; q0 += kParts * 8) {
            double t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int q = min(q0 + kParts * j, a.v.nb - 1);
                t[j] = __ldcg(aoi_sums + ((int64_t)q * C + cc) * kPostRed + i);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) v += (q0 + kParts * j < a.v.nb) ? t[j] : 0.0;
        }
        fin[vi][part] = v;
    }
    __syncthreads();
    if (threadIdx.x < C * NACC) acc_out[threadIdx.x] = (fin[threadIdx.x][0] + fin[threadIdx.x][1]) + (fin[threadIdx.x][2] + fin[threadIdx.x][3]);
}

// ---- hmm: the guide's chain (cosmos_hmm.cuh) -------------------------------------------------------------------------
// Everything the recursions need from one frame is local to it, so a per-unit kernel prepares it in parallel (the
// exps / logs of the softmax rows), and the recursions themselves are warp-per-chain chunked scans of cheap maps:
// lane l takes frames [l L, (l+1) L), composes its chunk, the 32 chunk maps are scanned with shuffles, and a second
// sweep over the chunk produces the per-frame values.  All in double (a few dozen flops per unit; only differences of
// the big emission values propagate, see HmmUnitOut::v).  One thread per chain walking 2000 frames cost 5.4 ms at C5.
//
// rows (6, U): q(1|z'=0), q(1|z'=1), rho(0), rho(1), rd(0), rd(1) with, for row z' of frame f,
//   rho(z') = sum_z q(z|z') R(z', z),  rd(z') = R(z', 1) - R(z', 0),  R = log p_f(z|z') - log q_f(z|z')
// (log p_f: the initial distribution at f = 0 -- only row 0 carries weight there -- else the transition matrix).
enum { ROW_Q0 = 0, ROW_Q1, ROW_RHO0, ROW_RHO1, ROW_RD0, ROW_RD1, NROW };

template <typename T>
__global__ void __launch_bounds__(kLocalBlock) hmm_rows_kernel(const LocalArgs<T> a, double* __restrict__ rows) {
    const uint32_t u32 = blockIdx.x * (uint32_t)kLocalBlock + threadIdx.x;
    if (u32 >= (uint32_t)a.U) return;
    const UnitIndex ui = locate_unit32(u32, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
    const int f = ui.fi;
    const int ot = a.v.is_ontarget[ui.aoi] ? 1 : 0;
    const ChannelTables<double>& ct = a.tables->ch[ui.c];
    const T* zt = a.lparams + a.hmm_ztrans(ui.aoi, f, ui.c);
    const ChainRow r[kZ] = {chain_row((double)zt[0], (double)zt[1], a.mc), chain_row((double)zt[2], (double)zt[3], a.mc)};
#pragma unroll
    for (int zr = 0; zr < kZ; ++zr) {
        double R[kZ];
#pragma unroll
        for (int z = 0; z < kZ; ++z) R[z] = (f == 0 ? ct.logpz[ot][z] : ct.logptrans[ot][zr][z]) - r[zr].lq[z];
        rows[(int64_t)(ROW_Q0 + zr) * a.U + u32] = r[zr].q[1];
        rows[(int64_t)(ROW_RHO0 + zr) * a.U + u32] = r[zr].q[0] * R[0] + r[zr].q[1] * R[1];
        rows[(int64_t)(ROW_RD0 + zr) * a.U + u32] = R[1] - R[0];
    }
}

// One BLOCK of kChainThreads threads per chain: thread t takes frames [t L, (t+1) L), L = ceil(F / kChainThreads);
// chunk maps are scanned inside each warp with shuffles and across the warps through shared memory.
constexpr int kChainThreads = 128;
constexpr int kChainWarps = kChainThreads / 32;
__device__ __forceinline__ void chain_chunk(int F, int t, int& f_lo, int& f_hi) {
    const int L = (F + kChainThreads - 1) / kChainThreads;
    f_lo = t * L < F ? t * L : F;
    f_hi = (t + 1) * L < F ? (t + 1) * L : F;
}

// forward marginals a_f(z): a_f = a_{f-1} T_f, T_f = [[1-p, p], [1-r, r]], a_{-1} = e_0 (row 0 of z_trans is the
// initial distribution, hmm.py:355-359).  Products of row-stochastic 2x2 matrices stay row-stochastic: (p, r) suffice.
template <typename T>
__global__ void __launch_bounds__(kChainThreads) hmm_forward_kernel(const LocalArgs<T> a, const double* __restrict__ rows,
                                                                    double* __restrict__ a_out) {
    __shared__ double wp[kChainWarps], wr[kChainWarps];
    const int chain = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = a.v.C, F = a.v.F;
    const int ni = chain / C, c = chain - ni * C;
    int f_lo, f_hi;
    chain_chunk(F, threadIdx.x, f_lo, f_hi);
    const int64_t u0 = ((int64_t)ni * F) * C + c;
    const double* q0 = rows + (int64_t)ROW_Q0 * a.U + u0;
    const double* q1 = rows + (int64_t)ROW_Q1 * a.U + u0;
    double p = 0.0, r = 1.0;   // identity
    for (int f = f_lo; f < f_hi; ++f) {
        const double p2 = q0[(int64_t)f * C], r2 = q1[(int64_t)f * C];
        const double np = (1.0 - p) * p2 + p * r2, nr = (1.0 - r) * p2 + r * r2;
        p = np; r = nr;
    }
    // inclusive scan of the chunk products inside the warp (earlier chunks multiply from the left)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double po = __shfl_up_sync(0xffffffffu, p, o), ro = __shfl_up_sync(0xffffffffu, r, o);
        if (lane >= o) {
            const double np = (1.0 - po) * p + po * r, nr = (1.0 - ro) * p + ro * r;
            p = np; r = nr;
        }
    }
    if (lane == 31) { wp[warp] = p; wr[warp] = r; }
    // exclusive within the warp: product of the earlier lanes
    double ep = __shfl_up_sync(0xffffffffu, p, 1), er = __shfl_up_sync(0xffffffffu, r, 1);
    if (lane == 0) { ep = 0.0; er = 1.0; }
    __syncthreads();
    // e_0 (earlier warps' product) (earlier lanes' product): only the probability of state 1 is needed
    double a1 = 0.0;   // e_0 -> state 1 with probability 0
    for (int w = 0; w < warp; ++w) a1 = (1.0 - a1) * wp[w] + a1 * wr[w];
    a1 = (1.0 - a1) * ep + a1 * er;
    for (int f = f_lo; f < f_hi; ++f) {
        const double p2 = q0[(int64_t)f * C], r2 = q1[(int64_t)f * C];
        a1 = (1.0 - a1) * p2 + a1 * r2;
        a_out[u0 + (int64_t)f * C] = 1.0 - a1;
        a_out[a.U + u0 + (int64_t)f * C] = a1;
    }
}

// likelihood-kernel weights W(m) = sum_z a(z) q(m | z) per unit
template <typename T>
__global__ void __launch_bounds__(kLocalBlock) hmm_weights_kernel(const LocalArgs<T> a) {
    const uint32_t u32 = blockIdx.x * (uint32_t)kLocalBlock + threadIdx.x;
    if (u32 >= (uint32_t)a.U) return;
    const UnitIndex ui = locate_unit32(u32, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
    T ump[kZ][kK], az[kZ], qm[kM];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        ump[0][k] = a.lparams[a.lo.index(LP_M_PROBS + k, ui.aoi, ui.fi, ui.c)];
        ump[1][k] = a.lparams[a.hmm_mprobs1(k, ui.aoi, ui.fi, ui.c)];
    }
#pragma unroll
    for (int z = 0; z < kZ; ++z) az[z] = (T)a.hmm_a[(int64_t)z * a.U + u32];
    hmm_presence_weights<T>(ump, az, a.mc, qm);
#pragma unroll
    for (int m = 0; m < kM; ++m) a.qm[m * a.U + u32] = qm[m];
}

// backward recursion: gradients of the unconstrained z_trans, the chain's ELBO terms and the expected
// initial-state / transition counts of this (AOI, channel) -> hpartial[(ni * C + c)][NHACC]
//   Delta_f = V_f(1) - V_f(0) + [rho_{f+1}(1) - rho_{f+1}(0)] + [q_{f+1}(1|1) - q_{f+1}(1|0)] Delta_{f+1}
//           = c_f + kappa_f Delta_{f+1}                                  (an affine map per frame: scanned right to left)
//   d ELBO / d u_f(z', 1) = a_{f-1}(z') q_f(1|z') q_f(0|z') [ rd_f(z') + Delta_f ] = - d / d u_f(z', 0)
template <typename T>
__global__ void __launch_bounds__(kChainThreads) hmm_backward_kernel(const LocalArgs<T> a, const double* __restrict__ rows,
                                                                     double* __restrict__ hpartial) {
    __shared__ double wA[kChainWarps], wB[kChainWarps], wacc[kChainWarps][NHACC];
    const int chain = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C = a.v.C, F = a.v.F;
    const int ni = chain / C, c = chain - ni * C;
    const int64_t n = a.v.ndx ? a.v.ndx[ni] : ni;
    const double mu = a.v.mask[n] ? 1.0 : 0.0;
    const bool ot = a.v.is_ontarget[n] != 0;
    int f_lo, f_hi;
    chain_chunk(F, threadIdx.x, f_lo, f_hi);
    const int64_t u0 = ((int64_t)ni * F) * C + c;
    auto row = [&](int j, int f) { return rows[(int64_t)j * a.U + u0 + (int64_t)f * C]; };
    auto c_of = [&](int f) {   // c_f and kappa_f
        const int64_t u = u0 + (int64_t)f * C;
        double cf = (double)a.hmm_v[a.U + u] - (double)a.hmm_v[u], kf = 0.0;
        if (f + 1 < F) { cf += row(ROW_RHO1, f + 1) - row(ROW_RHO0, f + 1); kf = row(ROW_Q1, f + 1) - row(ROW_Q0, f + 1); }
        return make_double2(cf, kf);
    };
    // this chunk as one map  Delta_first = A + B Delta_(first frame of the next chunk)
    double A = 0.0, B = 1.0;
    for (int f = f_hi - 1; f >= f_lo; --f) {
        const double2 ck = c_of(f);
        A = ck.x + ck.y * A;
        B = ck.y * B;
    }
    // inclusive suffix scan inside the warp: G_l = map_l o map_{l+1} o ... o map_31
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double Ao = __shfl_down_sync(0xffffffffu, A, o), Bo = __shfl_down_sync(0xffffffffu, B, o);
        if (lane + o < 32) { A = A + B * Ao; B = B * Bo; }
    }
    if (lane == 0) { wA[warp] = A; wB[warp] = B; }
    double eA = __shfl_down_sync(0xffffffffu, A, 1), eB = __shfl_down_sync(0xffffffffu, B, 1);   // later lanes of this warp
    if (lane == 31) { eA = 0.0; eB = 1.0; }
    __syncthreads();
    // Delta at the first frame after this thread's chunk: later warps applied to Delta = 0 beyond the last frame, then the
    // later lanes of this warp
    double later = 0.0;
    for (int w = kChainWarps - 1; w > warp; --w) later = wA[w] + wB[w] * later;
    double delta_next = eA + eB * later;
    double acc[NHACC];
#pragma unroll
    for (int i = 0; i < NHACC; ++i) acc[i] = 0.0;
    const double scale = -a.sN * mu;   // loss = -ELBO; every frame is used: no frame scale
    for (int f = f_hi - 1; f >= f_lo; --f) {
        const double2 ck = c_of(f);
        const double delta = ck.x + ck.y * delta_next;
        delta_next = delta;
        const int64_t u = u0 + (int64_t)f * C;
        double ap0 = 1.0, ap1 = 0.0;   // a_{f-1}
        if (f > 0) { ap0 = a.hmm_a[u - C]; ap1 = a.hmm_a[a.U + u - C]; }
        const double q01 = row(ROW_Q0, f), q11 = row(ROW_Q1, f);
        const double g0 = ap0 * q01 * (1.0 - q01) * (row(ROW_RD0, f) + delta);
        const double g1 = ap1 * q11 * (1.0 - q11) * (row(ROW_RD1, f) + delta);
        const int64_t iz = a.hmm_ztrans(n, f, c);
        a.lgrads[iz + 0] = (T)(-scale * g0);
        a.lgrads[iz + 1] = (T)(scale * g0);
        a.lgrads[iz + 2] = (T)(-scale * g1);
        a.lgrads[iz + 3] = (T)(scale * g1);
        acc[HACC_ELBO] += ap0 * row(ROW_RHO0, f) + ap1 * row(ROW_RHO1, f);
        if (ot) {
            if (f == 0) {
                acc[HACC_INIT + 0] += ap0 * (1.0 - q01);
                acc[HACC_INIT + 1] += ap0 * q01;
            } else {
                acc[HACC_TRANS + 0] += ap0 * (1.0 - q01);
                acc[HACC_TRANS + 1] += ap0 * q01;
                acc[HACC_TRANS + kZ + 0] += ap1 * (1.0 - q11);
                acc[HACC_TRANS + kZ + 1] += ap1 * q11;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NHACC; ++i) {
        const double v = warp_sum(acc[i]);   // fixed shuffle order: deterministic
        if (lane == 0) wacc[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NHACC) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kChainWarps; ++w) v += wacc[w][threadIdx.x];
        hpartial[(int64_t)chain * NHACC + threadIdx.x] = mu * v;
    }
}

// theta posterior of one particle given the MAP chain state, accumulated into the running mean (hmm.py:541-625)
template <typename T>
__global__ void __launch_bounds__(kLocalBlock) hmm_theta_kernel(const LocalArgs<T> a, const uint8_t* __restrict__ z_map, T weight,
                                                               T* __restrict__ theta_probs) {
    __shared__ GlobalTables<T> gt;
    if (threadIdx.x == 0) gt.convert_from(*a.tables);
    __syncthreads();
    const uint32_t u32 = blockIdx.x * (uint32_t)kLocalBlock + threadIdx.x;
    if (u32 >= (uint32_t)a.U) return;
    const UnitIndex ui = locate_unit32(u32, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
    const int z = z_map[u32] ? 1 : 0;
    T x[kK], y[kK], u_mp[kK], pth[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        x[k] = a.samples[(int64_t)(S_X + k) * a.U + u32];
        y[k] = a.samples[(int64_t)(S_Y + k) * a.U + u32];
        u_mp[k] = z ? a.lparams[a.hmm_mprobs1(k, ui.aoi, ui.fi, ui.c)] : a.lparams[a.lo.index(LP_M_PROBS + k, ui.aoi, ui.fi, ui.c)];
    }
    unit_theta_given_z<T>(x, y, u_mp, z, a.mc, gt, ui.c, pth);
#pragma unroll
    for (int k = 0; k < kK; ++k) theta_probs[(int64_t)k * a.U + u32] += weight * pth[k];
}

// hacc[c][i] = sum over the minibatch AOIs: one warp per value, lanes stride over the AOIs, fixed combination order
__global__ void __launch_bounds__(256) hmm_reduce_kernel(const double* __restrict__ hpartial, int nb, int C, double* __restrict__ hacc) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (blockIdx.x != 0) return;
    for (int t = warp; t < C * NHACC; t += 8) {
        const int c = t / NHACC, i = t - c * NHACC;
        double s = 0.0;
        for (int ni = lane; ni < nb; ni += 32) s += hpartial[((int64_t)ni * C + c) * NHACC + i];
        s = warp_sum(s);
        if (lane == 0) hacc[t] = s;
    }
}

// ---- z / theta posterior of one particle, accumulated into the running means (row N1) ---------------------
template <typename T>
__global__ void __launch_bounds__(kLocalBlock) zprobs_kernel(const LocalArgs<T> a, T weight, T* __restrict__ z_probs,
                                                            T* __restrict__ theta_probs) {
    __shared__ GlobalTables<T> gt;
    if (threadIdx.x == 0) gt.convert_from(*a.tables);
    __syncthreads();
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= a.U) return;
    const UnitIndex ui = locate_unit(u, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
    const int64_t f = a.v.fdx ? a.v.fdx[ui.fi] : ui.fi;
    T x[kK], y[kK], u_mp[kK], pz[kZ], pth[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        x[k] = a.samples[(int64_t)(S_X + k) * a.U + u];
        y[k] = a.samples[(int64_t)(S_Y + k) * a.U + u];
        u_mp[k] = a.lparams[a.lo.index(LP_M_PROBS + k, ui.aoi, f, ui.c)];
    }
    unit_ztheta_posterior<T>(x, y, u_mp, a.mc, gt, ui.c, a.v.is_ontarget[ui.aoi] != 0, pz, pth);
#pragma unroll
    for (int z = 0; z < kZ; ++z) z_probs[u * kZ + z] += weight * pz[z];
#pragma unroll
    for (int k = 0; k < kK; ++k) theta_probs[(int64_t)k * a.U + u] += weight * pth[k];
}

// ---- globals: reverse mode; one block per global site, then a fixed-order sum of the ELBO parts -----------------
__global__ void globals_grad_kernel(const double* __restrict__ gparams, int Q, ModelConst mc,
                                    const double* __restrict__ gstate, const double* __restrict__ acc,
                                    double sN, double sF, double* __restrict__ ggrads, double* __restrict__ elbo_parts) {
    const int site = blockIdx.x;
    if (threadIdx.x != 0 || site >= global_site_count(Q)) return;
    GlobalLayout gl{Q};
    double u[kMaxGlobals], grad[kMaxGlobals];
    for (int i = 0; i < gl.count(); ++i) { u[i] = (double)gparams[i]; grad[i] = 0.0; }
    elbo_parts[site] = globals_post_site(site, u, gl, mc, gstate + kMaxGlobalNoise, acc, sN, sF, grad);
    // each site owns its parameters' gradient entries
    if (site == 0) { ggrads[gl.gain_loc()] = grad[gl.gain_loc()]; ggrads[gl.gain_beta()] = grad[gl.gain_beta()]; }
    else if (site == 1) { ggrads[gl.prox_loc()] = grad[gl.prox_loc()]; ggrads[gl.prox_size()] = grad[gl.prox_size()]; }
    else if (site < 2 + Q) {
        const int q = site - 2;
        ggrads[gl.pi_mean(q, 0)] = grad[gl.pi_mean(q, 0)];
        ggrads[gl.pi_mean(q, 1)] = grad[gl.pi_mean(q, 1)];
        ggrads[gl.pi_size(q)] = grad[gl.pi_size(q)];
    } else {
        const int q = site - 2 - Q;
        ggrads[gl.lamda_loc(q)] = grad[gl.lamda_loc(q)];
        ggrads[gl.lamda_beta(q)] = grad[gl.lamda_beta(q)];
    }
}

__global__ void finalize_loss_kernel(const double* __restrict__ elbo_parts, int n, double* __restrict__ loss) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double e = 0.0;
    for (int i = 0; i < n; ++i) e += elbo_parts[i];
    loss[0] = -e;
}

// ---- globals, split reverse mode (see cosmos_globals.cuh): prepare runs right after sampling on the side stream,
// finish after the accumulators exist.  prepare: one block per site, lanes 0..2 evaluate drive = 0, e0, e1.
__global__ void globals_prepare_kernel(const double* __restrict__ gparams, int Q, bool hmm, ModelConst mc,
                                       const double* __restrict__ gstate, GlobalPrep* __restrict__ prep) {
    const int site = blockIdx.x, lane = threadIdx.x;
    if (lane >= 3 || site >= global_site_count(Q, hmm)) return;
    GlobalLayout gl{Q, hmm};
    double u[kMaxGlobals];
    for (int i = 0; i < gl.count(); ++i) { u[i] = (double)gparams[i]; prep->grad[site][lane][i] = 0.0; }
    const double drive[2] = {lane == 1 ? 1.0 : 0.0, lane == 2 ? 1.0 : 0.0};
    const double e = globals_post_site_driven(site, u, gl, mc, gstate + kMaxGlobalNoise, drive, prep->grad[site][lane]);
    if (lane == 0) prep->elbo[site] = e;
}

// finish: thread i owns global parameter i; thread 0 also sums the ELBO parts in a fixed order
__global__ void globals_finish_kernel(int Q, bool hmm, ModelConst mc, const double* __restrict__ gstate,
                                      const GlobalPrep* __restrict__ prep, const double* __restrict__ acc,
                                      const double* __restrict__ hacc, double sN,
                                      double sF, double* __restrict__ ggrads, double* __restrict__ loss) {
    GlobalLayout gl{Q, hmm};
    const int i = threadIdx.x;
    if (blockIdx.x != 0) return;
    if (i < gl.count()) {
        const int site = global_param_site(i, Q);
        double drive[2], elbo_data;
        globals_drive(site, gl, mc, gstate + kMaxGlobalNoise, acc, hacc, sN, sF, drive, elbo_data);
        const double g0 = prep->grad[site][0][i];
        ggrads[i] = g0 + drive[0] * (prep->grad[site][1][i] - g0) + drive[1] * (prep->grad[site][2][i] - g0);
    }
    if (i == 0) {
        double drive[2], e;
        globals_drive(0, gl, mc, gstate + kMaxGlobalNoise, acc, hacc, sN, sF, drive, e);
        for (int site = 0; site < global_site_count(Q, hmm); ++site) e += prep->elbo[site];
        loss[0] = -e;
    }
}

// ---- dense Adam (torch.optim.Adam defaults: no weight decay, no amsgrad) --------------------------------------
// models/model.py:168-171: pyro.optim.Adam({"lr", "betas": [0.9, 0.999]}) on every unconstrained tensor,
// dense over the whole tensor.  `state->step` is the number of completed steps.
template <typename T> struct alignas(4 * sizeof(T)) Vec4 { T x, y, z, w; };

// bias-corrected constants of update number t (1-based), rounded to T exactly once (shared by both evaluations)
template <typename T>
__device__ __forceinline__ void adam_constants(double lr, double b1, double b2, double t, T& step_size, T& inv_sqrt_bc2) {
    const double bc1 = 1.0 - pow(b1, t), bc2 = 1.0 - pow(b2, t);
    step_size = (T)(lr / bc1);
    inv_sqrt_bc2 = (T)(1.0 / sqrt(bc2));
}

// 28 B of HBM traffic per element (p, g, m, v read; p, m, v written): four elements per thread through 16-byte
// (float) / 32-byte (double) accesses, scalar tail.
// DEFERRED: the constants come from StepState (written by step_advance_deferred_kernel with the update they belong to)
// and nothing happens unless an update is pending -- tq_adam_deferred_flush.
template <typename T, bool VEC, bool DEFERRED = false>
__global__ void __launch_bounds__(256) adam_kernel(int64_t n, T* __restrict__ p, const T* __restrict__ g, T* __restrict__ m,
                                                   T* __restrict__ v, double lr, double b1, double b2, double eps,
                                                   const StepState* __restrict__ state) {
    T step_size, inv_sqrt_bc2;
    if (DEFERRED) {
        if (state->pending == 0u) return;
        step_size = (T)state->step_size;
        inv_sqrt_bc2 = (T)state->inv_sqrt_bc2;
    } else {
        adam_constants<T>(lr, b1, b2, (double)(state->step + 1ull), step_size, inv_sqrt_bc2);
    }
    const T tb1 = (T)b1, tb2 = (T)b2, teps = (T)eps;
    const int64_t n4 = VEC ? n / 4 : 0;
    Vec4<T>* p4 = reinterpret_cast<Vec4<T>*>(p);
    const Vec4<T>* g4 = reinterpret_cast<const Vec4<T>*>(g);
    Vec4<T>* m4 = reinterpret_cast<Vec4<T>*>(m);
    Vec4<T>* v4 = reinterpret_cast<Vec4<T>*>(v);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        Vec4<T> pi = p4[i], mi = m4[i], vi = v4[i];
        const Vec4<T> gi = g4[i];
        adam_update(pi.x, gi.x, mi.x, vi.x, step_size, inv_sqrt_bc2, tb1, tb2, teps);
        adam_update(pi.y, gi.y, mi.y, vi.y, step_size, inv_sqrt_bc2, tb1, tb2, teps);
        adam_update(pi.z, gi.z, mi.z, vi.z, step_size, inv_sqrt_bc2, tb1, tb2, teps);
        adam_update(pi.w, gi.w, mi.w, vi.w, step_size, inv_sqrt_bc2, tb1, tb2, teps);
        p4[i] = pi; m4[i] = mi; v4[i] = vi;
    }
    for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        adam_update(p[i], g[i], m[i], v[i], step_size, inv_sqrt_bc2, tb1, tb2, teps);
}

__global__ void step_advance_kernel(StepState* state) {
    if (blockIdx.x == 0 && threadIdx.x == 0) state->step += 1ull;
}

// End of a step whose AOI-local Adam update is deferred into the next step's site kernel: record the constants of THIS
// update (number step + 1) with the pending flag, then count the step.
__global__ void step_advance_deferred_kernel(StepState* state, double lr, double b1, double b2) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    float ss, isb;
    adam_constants<float>(lr, b1, b2, (double)(state->step + 1ull), ss, isb);
    state->step_size = ss;
    state->inv_sqrt_bc2 = isb;
    state->pending = 1u;
    state->step += 1ull;
}

__global__ void step_clear_pending_kernel(StepState* state) {
    if (blockIdx.x == 0 && threadIdx.x == 0) state->pending = 0u;
}

static int local_blocks(int64_t U) { return (int)((U + kLocalBlock - 1) / kLocalBlock); }

template <typename T>
static void fill_common(LocalArgs<T>& a, const tq_patch_view* view, int64_t Nt, const ModelConst* mc, const void* lparams,
                        const void* tables, int64_t aoi_offset, unsigned long long seed, const void* state) {
    a.v = *view;
    a.lo = LocalOffsets{Nt, (int64_t)view->F, (int64_t)view->C};
    a.mc = *mc;
    a.lparams = (const T*)lparams;
    a.tables = (const GlobalTables<double>*)tables;
    a.U = (int64_t)view->nb * view->fb * view->C;
    a.aoi_offset = aoi_offset;
    a.seed = seed;
    a.state = (const StepState*)state;
}

}  // namespace tq

using namespace tq;

extern "C" int tq_sizeof_tables(void) { return (int)sizeof(GlobalTables<double>); }
extern "C" int tq_sizeof_gstate(void) { return (int)(2 * kMaxGlobalNoise * sizeof(double)); }
extern "C" int tq_sizeof_model_const(void) { return (int)sizeof(ModelConst); }
// doubles of `block_partial` scratch and 8-byte slots of zero-initialised `aoi_partial` (ticket) scratch that
// tq_cosmos_local_post needs for a minibatch of nb AOIs x fb frames x C channels
extern "C" int64_t tq_local_post_scratch(int nb, int fb, int C) {
    return (int64_t)nb * C * (post_chunks(fb, post_upt(nb, fb, C)) + 1) * kPostRed;   // per-chunk partials + per-(AOI, channel) sums
}
extern "C" int64_t tq_local_post_tickets(int nb, int fb, int C) { (void)fb; return ((int64_t)nb * C + 1 + 1) / 2; }
extern "C" int tq_site_record_rows(void) { return NREC; }

static int globals_sample_impl(int dtype, int Q, bool hmm, const void* gparams, const void* mc, const double* noise_in,
                               uint64_t seed, const void* state, double* gstate, void* tables, void* gain_out, void* stream) {
    TQ_CHECK_ARG(Q >= 1 && Q <= kMaxC, "Q (channels) must be in [1, 4]");
    TQ_CHECK_ARG(gparams && mc && state && gstate && tables && gain_out, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const ModelConst m = *(const ModelConst*)mc;
    const int sites = global_site_count(Q, hmm);
    if (dtype == TQ_F32)
        globals_sample_kernel<float><<<sites, 32, 0, st>>>((const double*)gparams, Q, hmm, m, noise_in, seed, (const StepState*)state,
                                                      gstate, (GlobalTables<double>*)tables, (float*)gain_out);
    else if (dtype == TQ_F64)
        globals_sample_kernel<double><<<sites, 32, 0, st>>>((const double*)gparams, Q, hmm, m, noise_in, seed, (const StepState*)state,
                                                       gstate, (GlobalTables<double>*)tables, (double*)gain_out);
    else { set_error("bad dtype %d", dtype); return TQ_ERR_ARG; }
    TQ_LAUNCH_CHECK("globals_sample_kernel launch");
    return TQ_OK;
}

extern "C" int tq_cosmos_globals_sample(int dtype, int Q, const void* gparams, const void* mc, const double* noise_in,
                                        uint64_t seed, const void* state, double* gstate, void* tables,
                                        void* gain_out, void* stream) {
    return globals_sample_impl(dtype, Q, false, gparams, mc, noise_in, seed, state, gstate, tables, gain_out, stream);
}
extern "C" int tq_hmm_globals_sample(int dtype, int Q, const void* gparams, const void* mc, const double* noise_in,
                                     uint64_t seed, const void* state, double* gstate, void* tables,
                                     void* gain_out, void* stream) {
    return globals_sample_impl(dtype, Q, true, gparams, mc, noise_in, seed, state, gstate, tables, gain_out, stream);
}

struct DeferredAdam {   // tq_cosmos_sites_adam
    const void* grads;
    void* exp_avg;
    void* exp_avg_sq;
    double beta1, beta2, eps;
};

template <typename T>
static int run_sites(const tq_patch_view* view, int64_t Nt, const ModelConst* mc, const void* lparams,
                     int64_t aoi_offset, uint64_t seed, const void* state, const void* noise_in, void* samples,
                     void* qm, void* rec, void* worklist, void* work_count, cudaStream_t st, const DeferredAdam* adam = nullptr) {
    LocalArgs<T> a{};
    fill_common(a, view, Nt, mc, lparams, nullptr, aoi_offset, seed, state);
    if (adam) {
        a.adam_p = (T*)const_cast<void*>(lparams);
        a.adam_g = (const T*)adam->grads;
        a.adam_m = (T*)adam->exp_avg;
        a.adam_v = (T*)adam->exp_avg_sq;
        a.adam_b1 = (float)adam->beta1; a.adam_b2 = (float)adam->beta2; a.adam_eps = (float)adam->eps;
    }
    a.noise_in = (const T*)noise_in;
    a.samples = (T*)samples;
    a.qm = (T*)qm;
    a.rec = (T*)rec;
    if (a.U == 0) return TQ_OK;
    if (a.U >= (int64_t)1 << 31) { set_error("minibatch of %lld units exceeds the 2^31 limit of one launch", (long long)a.U); return TQ_ERR_ARG; }
    const dim3 grid(local_blocks(a.U), NSAMP);
    if constexpr (sizeof(T) == sizeof(float)) {
        const bool ws = worklist != nullptr && work_count != nullptr && a.U * NSAMP < ((int64_t)1 << 32);
        a.worklist = ws ? (uint32_t*)worklist : nullptr;
        a.work_count = ws ? (unsigned int*)work_count : nullptr;
        if (ws) {
            int stm = cuda_status(cudaMemsetAsync(work_count, 0, sizeof(unsigned int), st), "cudaMemsetAsync(work_count)");
            if (stm != TQ_OK) return stm;
        }
        // eight units per thread once the grid still covers the GPU several times over (>= 4 waves of 8 blocks per SM)
#ifndef TQ_SITE_UPT_BIG
#define TQ_SITE_UPT_BIG 8   // B200, one rank's C3 shard after 3000 iterations: 608 (1 x 128 sites, 4 classes) -> 588 (4 x, 64 classes) -> 550 us (8 x)
#endif
        const int64_t blocks4 = (a.U + TQ_SITE_UPT_BIG * kLocalBlock - 1) / (TQ_SITE_UPT_BIG * kLocalBlock);
        // TQ_SITE_SPLIT=1: the two passes as two launches (same bits either way).  Opt-in: measured on a B200 (one 8-GPU rank's
        // C3 shard, site kernels at the initial point / after 1000 / 3000 iterations) 216 / 241 / 396 us against 190 / 236 / 438
        // in one launch -- it wins once most height / width sites are deferred (a fit beyond ~2000 iterations) and loses before,
        // where the few replays of a block overlap with other blocks' first pass only in the one-launch form
        static const bool split = [] { const char* e = getenv("TQ_SITE_SPLIT"); return e && e[0] == '1'; }();
        if (blocks4 * NSAMP >= (int64_t)sm_count() * 8 * 4 && split && ws) {
            void* sym = nullptr;
            int stc = cuda_status(cudaGetSymbolAddress(&sym, g_site_deferred), "cudaGetSymbolAddress(g_site_deferred)");
            if (stc != TQ_OK) return stc;
            unsigned int* n_deferred = static_cast<unsigned int*>(sym) + (((uintptr_t)st >> 8) & 7u);
            stc = cuda_status(cudaMemsetAsync(n_deferred, 0, sizeof(unsigned int), st), "cudaMemsetAsync(n_deferred)");
            if (stc != TQ_OK) return stc;
            site_pass1_kernel<TQ_SITE_UPT_BIG><<<dim3((unsigned)blocks4, NSAMP), kLocalBlock, 0, st>>>(a, n_deferred);
            TQ_LAUNCH_CHECK("site_pass1_kernel launch");
            // one block per span, as in pass 1 (a single resident wave walking the spans was latency-bound: 69 us for the
            // scan alone at 625 k units)
            site_pass2_kernel<TQ_SITE_UPT_BIG><<<(unsigned)(blocks4 * NSAMP), kLocalBlock, 0, st>>>(a, n_deferred, (unsigned)blocks4);
        } else if (blocks4 * NSAMP >= (int64_t)sm_count() * 8 * 4)
            site_fast_kernel<TQ_SITE_UPT_BIG><<<dim3((unsigned)blocks4, NSAMP), kLocalBlock, 0, st>>>(a);
        else
            site_fast_kernel<1><<<grid, kLocalBlock, 0, st>>>(a);
        TQ_LAUNCH_CHECK("site_fast_kernel launch");
        if (ws) {
            site_worklist_kernel<<<sm_count() * 8, kLocalBlock, 0, st>>>(a);
            TQ_LAUNCH_CHECK("site_worklist_kernel launch");
        } else {
            site_fallback_kernel<<<dim3((grid.x + kFallbackUPT - 1) / kFallbackUPT, NSAMP), kLocalBlock, 0, st>>>(a);
            TQ_LAUNCH_CHECK("site_fallback_kernel launch");
        }
    } else {
        site_kernel<T><<<grid, kLocalBlock, 0, st>>>(a);
        TQ_LAUNCH_CHECK("site_kernel launch");
    }
    return TQ_OK;
}

static int sites_impl(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                      int64_t aoi_offset, uint64_t seed, const void* state, const void* noise_in,
                      void* samples, void* qm, void* rec, void* worklist, void* work_count, void* stream,
                      const DeferredAdam* adam = nullptr) {
    TQ_CHECK_ARG(view && mc && lparams && state && samples && qm && rec, "NULL pointer");
    TQ_CHECK_ARG(view->C >= 1 && view->C <= kMaxC, "C (channels) must be in [1, 4]");
    cudaStream_t st = (cudaStream_t)stream;
    if (adam) {
        TQ_CHECK_ARG(dtype == TQ_F32, "the deferred Adam update lives in the fp32 site kernel");
        TQ_CHECK_ARG(adam->grads && adam->exp_avg && adam->exp_avg_sq, "NULL pointer");
        TQ_CHECK_ARG(view->ndx == nullptr && view->fdx == nullptr && view->nb == Nt && view->fb == view->F,
                     "the deferred Adam update needs a full-batch step (a subsampled step visits only its minibatch, the update is dense)");
    }
    if (dtype == TQ_F32) return run_sites<float>(view, Nt, (const ModelConst*)mc, lparams, aoi_offset, seed, state, noise_in, samples, qm, rec, worklist, work_count, st, adam);
    if (dtype == TQ_F64) return run_sites<double>(view, Nt, (const ModelConst*)mc, lparams, aoi_offset, seed, state, noise_in, samples, qm, rec, worklist, work_count, st);
    set_error("bad dtype %d", dtype);
    return TQ_ERR_ARG;
}

extern "C" int tq_cosmos_sites(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                               int64_t aoi_offset, uint64_t seed, const void* state, const void* noise_in,
                               void* samples, void* qm, void* rec, void* stream) {
    return sites_impl(dtype, view, Nt, mc, lparams, aoi_offset, seed, state, noise_in, samples, qm, rec, nullptr, nullptr, stream);
}

extern "C" int tq_cosmos_sites_ws(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                                  int64_t aoi_offset, uint64_t seed, const void* state, const void* noise_in,
                                  void* samples, void* qm, void* rec, void* worklist, void* work_count, void* stream) {
    TQ_CHECK_ARG(worklist && work_count, "NULL workspace");
    return sites_impl(dtype, view, Nt, mc, lparams, aoi_offset, seed, state, noise_in, samples, qm, rec, worklist, work_count, stream);
}

// tq_cosmos_sites_ws for a FULL-BATCH fp32 step, with the previous step's dense Adam update of the AOI-local parameters
// folded in: if StepState says an update is pending (tq_step_advance_deferred), every site's thread applies it to the
// parameters it owns before reading them.  Covers the flat range tq_local_deferred_range reports; the rest of the
// buffer (per-AOI background parameters, `size`) is updated by tq_adam_dense at the end of the step as before.
extern "C" int tq_cosmos_sites_adam(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, void* lparams,
                                    int64_t aoi_offset, uint64_t seed, const void* state, const void* noise_in,
                                    void* samples, void* qm, void* rec, void* worklist, void* work_count,
                                    const void* lgrads, void* exp_avg, void* exp_avg_sq, double beta1, double beta2,
                                    double eps, void* stream) {
    TQ_CHECK_ARG(worklist && work_count, "NULL workspace");
    const DeferredAdam adam{lgrads, exp_avg, exp_avg_sq, beta1, beta2, eps};
    return sites_impl(dtype, view, Nt, mc, lparams, aoi_offset, seed, state, noise_in, samples, qm, rec, worklist, work_count, stream, &adam);
}

// [*begin, *end): the entries of the flat AOI-local buffer whose update tq_cosmos_sites_adam performs
extern "C" int tq_local_deferred_range(int64_t Nt, int64_t F, int64_t C, int64_t* begin, int64_t* end) {
    TQ_CHECK_ARG(begin && end && Nt >= 0 && F >= 0 && C >= 1, "bad arguments");
    const LocalOffsets lo{Nt, F, C};
    *begin = lo.index(LP_B_LOC, 0, 0, 0);
    *end = lo.index(LP_SIZE, 0, 0, 0);
    return TQ_OK;
}
extern "C" int tq_sizeof_step_state(void) { return (int)sizeof(StepState); }

template <typename T>
static int run_local_post(const tq_patch_view* view, int64_t Nt, const ModelConst* mc, const void* lparams, const void* tables,
                          const void* samples, const void* rec, const void* L, const void* gs, const void* g_rate, double sN, double sF,
                          void* lgrads, double* aoi_partial, double* block_partial, double* acc, cudaStream_t st) {
    LocalArgs<T> a{};
    fill_common(a, view, Nt, mc, lparams, tables, 0, 0, nullptr);
    a.samples = (T*)samples;
    a.rec = (T*)rec;
    a.L = (const T*)L;
    a.gs = (const T*)gs;  // (NSAMP, U) in S_* order: background, height_k, width_k, x_k, y_k
    a.g_rate = (const T*)g_rate;
    a.sN = sN; a.sF = sF;
    a.lgrads = (T*)lgrads;
    a.aoi_partial = aoi_partial;
    a.block_partial = block_partial;
    const int upt = post_upt(view->nb, view->fb, view->C);
    const int chunks = post_chunks(view->fb, upt);
    const int64_t nblocks = (int64_t)view->nb * view->C * chunks;
    if (a.U > 0) {
        if (nblocks >= (int64_t)1 << 31) { set_error("minibatch too large for one launch"); return TQ_ERR_ARG; }
        if (upt == 4) local_post_kernel<T, 4><<<(int)nblocks, kLocalBlock, 0, st>>>(a, chunks, (unsigned int*)aoi_partial, acc);
        else local_post_kernel<T, 1><<<(int)nblocks, kLocalBlock, 0, st>>>(a, chunks, (unsigned int*)aoi_partial, acc);
        TQ_LAUNCH_CHECK("local_post_kernel launch");
    } else {
        cudaMemsetAsync(acc, 0, sizeof(double) * view->C * NACC, st);
    }
    return TQ_OK;
}

extern "C" int tq_cosmos_local_post(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                                    const void* tables, const void* samples, const void* rec, const void* L, const void* gs,
                                    const void* g_rate, double sN, double sF, void* lgrads, double* aoi_partial,
                                    double* block_partial, double* acc, void* stream) {
    TQ_CHECK_ARG(view && mc && lparams && tables && samples && rec && L && gs && g_rate, "NULL input pointer");
    TQ_CHECK_ARG(lgrads && aoi_partial && block_partial && acc, "NULL output pointer");
    TQ_CHECK_ARG(view->mask && view->is_ontarget, "view needs mask and is_ontarget");
    TQ_CHECK_ARG(view->C >= 1 && view->C <= kMaxC, "C (channels) must be in [1, 4]");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TQ_F32) return run_local_post<float>(view, Nt, (const ModelConst*)mc, lparams, tables, samples, rec, L, gs, g_rate, sN, sF, lgrads, aoi_partial, block_partial, acc, st);
    if (dtype == TQ_F64) return run_local_post<double>(view, Nt, (const ModelConst*)mc, lparams, tables, samples, rec, L, gs, g_rate, sN, sF, lgrads, aoi_partial, block_partial, acc, st);
    set_error("bad dtype %d", dtype);
    return TQ_ERR_ARG;
}

// ---- hmm variant: host side ---------------------------------------------------------------------------------------------------
template <typename T>
static int run_hmm_forward(const tq_patch_view* view, int64_t Nt, const ModelConst* mc, const void* lparams, const void* tables,
                           double* rows, double* a_out, void* qm, cudaStream_t st) {
    LocalArgs<T> a{};
    fill_common(a, view, Nt, mc, lparams, tables, 0, 0, nullptr);
    a.hmm_a = a_out;
    a.qm = (T*)qm;
    if (a.U == 0) return TQ_OK;
    const int chains = view->nb * view->C;
    hmm_rows_kernel<T><<<local_blocks(a.U), kLocalBlock, 0, st>>>(a, rows);
    TQ_LAUNCH_CHECK("hmm_rows_kernel launch");
    hmm_forward_kernel<T><<<chains, kChainThreads, 0, st>>>(a, rows, a_out);
    TQ_LAUNCH_CHECK("hmm_forward_kernel launch");
    hmm_weights_kernel<T><<<local_blocks(a.U), kLocalBlock, 0, st>>>(a);
    TQ_LAUNCH_CHECK("hmm_weights_kernel launch");
    return TQ_OK;
}

static int hmm_check_view(const tq_patch_view* view) {
    TQ_CHECK_ARG(view != nullptr, "NULL view");
    TQ_CHECK_ARG(view->C >= 1 && view->C <= kMaxC, "C (channels) must be in [1, 4]");
    TQ_CHECK_ARG(view->fb == view->F && view->fdx == nullptr, "the hmm variant uses every frame (models/hmm.py:127-131): fb == F, fdx == NULL");
    TQ_CHECK_ARG((int64_t)view->nb * view->fb * view->C < ((int64_t)1 << 31), "minibatch too large for one launch");
    return TQ_OK;
}

extern "C" int64_t tq_hmm_local_numel(int64_t Nt, int64_t F, int64_t C) {
    LocalOffsets lo{Nt, F, C};
    return lo.numel() + (int64_t)kK * Nt * F * C + Nt * F * C * kZ * kZ;
}
extern "C" int tq_hmm_chain_sums(void) { return NHACC; }

extern "C" int tq_hmm_chain_rows(void) { return NROW; }

extern "C" int tq_hmm_forward(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                              const void* tables, double* rows, double* a_out, void* qm, void* stream) {
    int stv = hmm_check_view(view);
    if (stv != TQ_OK) return stv;
    TQ_CHECK_ARG(mc && lparams && tables && rows && a_out && qm, "NULL pointer");
    TQ_CHECK_ARG(view->is_ontarget, "view needs is_ontarget");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TQ_F32) return run_hmm_forward<float>(view, Nt, (const ModelConst*)mc, lparams, tables, rows, a_out, qm, st);
    if (dtype == TQ_F64) return run_hmm_forward<double>(view, Nt, (const ModelConst*)mc, lparams, tables, rows, a_out, qm, st);
    set_error("bad dtype %d", dtype);
    return TQ_ERR_ARG;
}

template <typename T>
static int run_hmm_post(const tq_patch_view* view, int64_t Nt, const ModelConst* mc, const void* lparams, const void* tables,
                        const void* samples, const void* rec, const void* L, const void* gs, const void* g_rate, const double* a_in,
                        double sN, void* lgrads, void* v_out, double* tickets, double* block_partial, double* acc, cudaStream_t st) {
    LocalArgs<T> a{};
    fill_common(a, view, Nt, mc, lparams, tables, 0, 0, nullptr);
    a.samples = (T*)samples;
    a.rec = (T*)rec;
    a.L = (const T*)L;
    a.gs = (const T*)gs;
    a.g_rate = (const T*)g_rate;
    a.sN = sN; a.sF = 1.0;
    a.lgrads = (T*)lgrads;
    a.block_partial = block_partial;
    a.hmm_a = a_in;
    a.hmm_v = (T*)v_out;
    const int upt = post_upt(view->nb, view->fb, view->C);
    const int chunks = post_chunks(view->fb, upt);
    const int64_t nblocks = (int64_t)view->nb * view->C * chunks;
    if (a.U == 0) { cudaMemsetAsync(acc, 0, sizeof(double) * view->C * NACC, st); return TQ_OK; }
    if (nblocks >= (int64_t)1 << 31) { set_error("minibatch too large for one launch"); return TQ_ERR_ARG; }
    if (upt == 4) local_post_kernel<T, 4, true><<<(int)nblocks, kLocalBlock, 0, st>>>(a, chunks, (unsigned int*)tickets, acc);
    else local_post_kernel<T, 1, true><<<(int)nblocks, kLocalBlock, 0, st>>>(a, chunks, (unsigned int*)tickets, acc);
    TQ_LAUNCH_CHECK("local_post_kernel<hmm> launch");
    return TQ_OK;
}

extern "C" int tq_hmm_local_post(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                                 const void* tables, const void* samples, const void* rec, const void* L, const void* gs,
                                 const void* g_rate, const double* a_in, double sN, void* lgrads, void* v_out, double* tickets,
                                 double* block_partial, double* acc, void* stream) {
    int stv = hmm_check_view(view);
    if (stv != TQ_OK) return stv;
    TQ_CHECK_ARG(mc && lparams && tables && samples && rec && L && gs && g_rate && a_in, "NULL input pointer");
    TQ_CHECK_ARG(lgrads && v_out && tickets && block_partial && acc, "NULL output pointer");
    TQ_CHECK_ARG(view->mask && view->is_ontarget, "view needs mask and is_ontarget");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TQ_F32) return run_hmm_post<float>(view, Nt, (const ModelConst*)mc, lparams, tables, samples, rec, L, gs, g_rate, a_in, sN, lgrads, v_out, tickets, block_partial, acc, st);
    if (dtype == TQ_F64) return run_hmm_post<double>(view, Nt, (const ModelConst*)mc, lparams, tables, samples, rec, L, gs, g_rate, a_in, sN, lgrads, v_out, tickets, block_partial, acc, st);
    set_error("bad dtype %d", dtype);
    return TQ_ERR_ARG;
}

template <typename T>
static int run_hmm_backward(const tq_patch_view* view, int64_t Nt, const ModelConst* mc, const void* lparams, const void* tables,
                            const double* rows, const double* a_in, const void* v_in, double sN, void* lgrads, double* hpartial,
                            double* hacc, cudaStream_t st) {
    LocalArgs<T> a{};
    fill_common(a, view, Nt, mc, lparams, tables, 0, 0, nullptr);
    a.sN = sN; a.sF = 1.0;
    a.lgrads = (T*)lgrads;
    a.hmm_a = a_in;
    a.hmm_v = (T*)v_in;
    const int chains = view->nb * view->C;
    if (chains > 0 && a.U > 0) {
        hmm_backward_kernel<T><<<chains, kChainThreads, 0, st>>>(a, rows, hpartial);
        TQ_LAUNCH_CHECK("hmm_backward_kernel launch");
    }
    hmm_reduce_kernel<<<1, 256, 0, st>>>(hpartial, a.U > 0 ? view->nb : 0, view->C, hacc);
    TQ_LAUNCH_CHECK("hmm_reduce_kernel launch");
    return TQ_OK;
}

extern "C" int tq_hmm_backward(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                               const void* tables, const double* rows, const double* a_in, const void* v_in, double sN,
                               void* lgrads, double* hpartial, double* hacc, void* stream) {
    int stv = hmm_check_view(view);
    if (stv != TQ_OK) return stv;
    TQ_CHECK_ARG(mc && lparams && tables && rows && a_in && v_in && lgrads && hpartial && hacc, "NULL pointer");
    TQ_CHECK_ARG(view->mask && view->is_ontarget, "view needs mask and is_ontarget");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TQ_F32) return run_hmm_backward<float>(view, Nt, (const ModelConst*)mc, lparams, tables, rows, a_in, v_in, sN, lgrads, hpartial, hacc, st);
    if (dtype == TQ_F64) return run_hmm_backward<double>(view, Nt, (const ModelConst*)mc, lparams, tables, rows, a_in, v_in, sN, lgrads, hpartial, hacc, st);
    set_error("bad dtype %d", dtype);
    return TQ_ERR_ARG;
}

template <typename T>
static int run_hmm_theta(const tq_patch_view* view, int64_t Nt, const ModelConst* mc, const void* lparams, const void* tables,
                         const void* samples, const void* z_map, double weight, void* theta_probs, cudaStream_t st) {
    LocalArgs<T> a{};
    fill_common(a, view, Nt, mc, lparams, tables, 0, 0, nullptr);
    a.samples = (T*)samples;
    if (a.U == 0) return TQ_OK;
    hmm_theta_kernel<T><<<local_blocks(a.U), kLocalBlock, 0, st>>>(a, (const uint8_t*)z_map, (T)weight, (T*)theta_probs);
    TQ_LAUNCH_CHECK("hmm_theta_kernel launch");
    return TQ_OK;
}

extern "C" int tq_hmm_theta_probs(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                                  const void* tables, const void* samples, const void* z_map, double weight,
                                  void* theta_probs, void* stream) {
    int stv = hmm_check_view(view);
    if (stv != TQ_OK) return stv;
    TQ_CHECK_ARG(mc && lparams && tables && samples && z_map && theta_probs, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TQ_F32) return run_hmm_theta<float>(view, Nt, (const ModelConst*)mc, lparams, tables, samples, z_map, weight, theta_probs, st);
    if (dtype == TQ_F64) return run_hmm_theta<double>(view, Nt, (const ModelConst*)mc, lparams, tables, samples, z_map, weight, theta_probs, st);
    set_error("bad dtype %d", dtype);
    return TQ_ERR_ARG;
}

template <typename T>
static int run_zprobs(const tq_patch_view* view, int64_t Nt, const ModelConst* mc, const void* lparams, const void* tables,
                      const void* samples, double weight, void* z_probs, void* theta_probs, cudaStream_t st) {
    LocalArgs<T> a{};
    fill_common(a, view, Nt, mc, lparams, tables, 0, 0, nullptr);
    a.samples = (T*)samples;
    if (a.U == 0) return TQ_OK;
    zprobs_kernel<T><<<local_blocks(a.U), kLocalBlock, 0, st>>>(a, (T)weight, (T*)z_probs, (T*)theta_probs);
    TQ_LAUNCH_CHECK("zprobs_kernel launch");
    return TQ_OK;
}

extern "C" int tq_cosmos_zprobs(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                                const void* tables, const void* samples, double weight, void* z_probs,
                                void* theta_probs, void* stream) {
    TQ_CHECK_ARG(view && mc && lparams && tables && samples && z_probs && theta_probs, "NULL pointer");
    TQ_CHECK_ARG(view->is_ontarget, "view needs is_ontarget");
    TQ_CHECK_ARG(view->C >= 1 && view->C <= kMaxC, "C (channels) must be in [1, 4]");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TQ_F32) return run_zprobs<float>(view, Nt, (const ModelConst*)mc, lparams, tables, samples, weight, z_probs, theta_probs, st);
    if (dtype == TQ_F64) return run_zprobs<double>(view, Nt, (const ModelConst*)mc, lparams, tables, samples, weight, z_probs, theta_probs, st);
    set_error("bad dtype %d", dtype);
    return TQ_ERR_ARG;
}

extern "C" int tq_cosmos_globals_grad(int dtype, int Q, const void* gparams, const void* mc, const double* gstate,
                                      const double* acc, double sN, double sF, void* ggrads, double* elbo_parts,
                                      double* loss, void* stream) {
    TQ_CHECK_ARG(Q >= 1 && Q <= kMaxC, "Q (channels) must be in [1, 4]");
    TQ_CHECK_ARG(gparams && mc && gstate && acc && ggrads && elbo_parts && loss, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const ModelConst m = *(const ModelConst*)mc;
    if (dtype != TQ_F32 && dtype != TQ_F64) { set_error("bad dtype %d", dtype); return TQ_ERR_ARG; }
    globals_grad_kernel<<<global_site_count(Q), 32, 0, st>>>((const double*)gparams, Q, m, gstate, acc, sN, sF, (double*)ggrads, elbo_parts);
    TQ_LAUNCH_CHECK("globals_grad_kernel launch");
    finalize_loss_kernel<<<1, 32, 0, st>>>(elbo_parts, global_site_count(Q), loss);
    TQ_LAUNCH_CHECK("finalize_loss_kernel launch");
    return TQ_OK;
}

extern "C" int tq_sizeof_gprep(void) { return (int)sizeof(GlobalPrep); }

static int globals_prepare_impl(int dtype, int Q, bool hmm, const void* gparams, const void* mc, const double* gstate,
                                void* gprep, void* stream) {
    TQ_CHECK_ARG(Q >= 1 && Q <= kMaxC, "Q (channels) must be in [1, 4]");
    TQ_CHECK_ARG(gparams && mc && gstate && gprep, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const ModelConst m = *(const ModelConst*)mc;
    const int sites = global_site_count(Q, hmm);
    if (dtype != TQ_F32 && dtype != TQ_F64) { set_error("bad dtype %d", dtype); return TQ_ERR_ARG; }
    globals_prepare_kernel<<<sites, 32, 0, st>>>((const double*)gparams, Q, hmm, m, gstate, (GlobalPrep*)gprep);
    TQ_LAUNCH_CHECK("globals_prepare_kernel launch");
    return TQ_OK;
}

static int globals_finish_impl(int dtype, int Q, bool hmm, const void* mc, const double* gstate, const void* gprep,
                               const double* acc, const double* hacc, double sN, double sF, void* ggrads, double* loss, void* stream) {
    TQ_CHECK_ARG(Q >= 1 && Q <= kMaxC, "Q (channels) must be in [1, 4]");
    TQ_CHECK_ARG(mc && gstate && gprep && acc && ggrads && loss && (hacc || !hmm), "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const ModelConst m = *(const ModelConst*)mc;
    if (dtype != TQ_F32 && dtype != TQ_F64) { set_error("bad dtype %d", dtype); return TQ_ERR_ARG; }
    globals_finish_kernel<<<1, 64, 0, st>>>(Q, hmm, m, gstate, (const GlobalPrep*)gprep, acc, hacc, sN, sF, (double*)ggrads, loss);
    TQ_LAUNCH_CHECK("globals_finish_kernel launch");
    return TQ_OK;
}

extern "C" int tq_cosmos_globals_prepare(int dtype, int Q, const void* gparams, const void* mc, const double* gstate,
                                         void* gprep, void* stream) {
    return globals_prepare_impl(dtype, Q, false, gparams, mc, gstate, gprep, stream);
}
extern "C" int tq_cosmos_globals_finish(int dtype, int Q, const void* mc, const double* gstate, const void* gprep,
                                        const double* acc, double sN, double sF, void* ggrads, double* loss, void* stream) {
    return globals_finish_impl(dtype, Q, false, mc, gstate, gprep, acc, nullptr, sN, sF, ggrads, loss, stream);
}
extern "C" int tq_hmm_globals_prepare(int dtype, int Q, const void* gparams, const void* mc, const double* gstate,
                                      void* gprep, void* stream) {
    return globals_prepare_impl(dtype, Q, true, gparams, mc, gstate, gprep, stream);
}
extern "C" int tq_hmm_globals_finish(int dtype, int Q, const void* mc, const double* gstate, const void* gprep,
                                     const double* acc, const double* hacc, double sN, void* ggrads, double* loss, void* stream) {
    return globals_finish_impl(dtype, Q, true, mc, gstate, gprep, acc, hacc, sN, 1.0, ggrads, loss, stream);
}

extern "C" int tq_adam_dense(int dtype, int64_t n, void* params, const void* grads, void* exp_avg, void* exp_avg_sq,
                             double lr, double beta1, double beta2, double eps, const void* state, void* stream) {
    TQ_CHECK_ARG(n >= 0, "negative size");
    if (n == 0) return TQ_OK;
    TQ_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && state, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 256;
    int64_t grid = (n / 4 + block - 1) / block;
    if (grid < 1) grid = 1;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (grid > cap) grid = cap;
    // the four-wide path needs 16-byte (float) / 32-byte (double) aligned buffers; anything else takes the scalar path
    const size_t esz = dtype == TQ_F64 ? 8 : 4;
    const bool vec = (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % (4 * esz)) == 0;
    if (!vec) { grid = (n + block - 1) / block; if (grid > cap) grid = cap; }
    if (dtype == TQ_F32) {
        if (vec) adam_kernel<float, true><<<(int)grid, block, 0, st>>>(n, (float*)params, (const float*)grads, (float*)exp_avg, (float*)exp_avg_sq, lr, beta1, beta2, eps, (const StepState*)state);
        else adam_kernel<float, false><<<(int)grid, block, 0, st>>>(n, (float*)params, (const float*)grads, (float*)exp_avg, (float*)exp_avg_sq, lr, beta1, beta2, eps, (const StepState*)state);
    } else if (dtype == TQ_F64) {
        if (vec) adam_kernel<double, true><<<(int)grid, block, 0, st>>>(n, (double*)params, (const double*)grads, (double*)exp_avg, (double*)exp_avg_sq, lr, beta1, beta2, eps, (const StepState*)state);
        else adam_kernel<double, false><<<(int)grid, block, 0, st>>>(n, (double*)params, (const double*)grads, (double*)exp_avg, (double*)exp_avg_sq, lr, beta1, beta2, eps, (const StepState*)state);
    } else { set_error("bad dtype %d", dtype); return TQ_ERR_ARG; }
    TQ_LAUNCH_CHECK("adam_kernel launch");
    return TQ_OK;
}

// End of a step whose local update is deferred: StepState (tq_sizeof_step_state bytes) receives the constants of this
// update and the pending flag, then the step count advances.
extern "C" int tq_step_advance_deferred(void* state, double lr, double beta1, double beta2, void* stream) {
    TQ_CHECK_ARG(state != nullptr, "NULL pointer");
    step_advance_deferred_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((StepState*)state, lr, beta1, beta2);
    TQ_LAUNCH_CHECK("step_advance_deferred_kernel launch");
    return TQ_OK;
}

// Apply a pending deferred update NOW (before anything but tq_cosmos_sites_adam reads the parameters: checkpoints,
// statistics, a subsampled step ...) over n entries -- the range of tq_local_deferred_range -- and clear the flag.
// No-op when nothing is pending.
extern "C" int tq_adam_deferred_flush(int dtype, int64_t n, void* params, const void* grads, void* exp_avg, void* exp_avg_sq,
                                      double beta1, double beta2, double eps, void* state, void* stream) {
    TQ_CHECK_ARG(dtype == TQ_F32, "the deferred Adam update is fp32");
    TQ_CHECK_ARG(n >= 0 && state, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n > 0) {
        TQ_CHECK_ARG(params && grads && exp_avg && exp_avg_sq, "NULL pointer");
        const int block = 256;
        const int64_t cap = (int64_t)sm_count() * 8;
        const bool vec = (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16) == 0;
        int64_t grid = ((vec ? n / 4 : n) + block - 1) / block;
        if (grid < 1) grid = 1;
        if (grid > cap) grid = cap;
        if (vec) adam_kernel<float, true, true><<<(int)grid, block, 0, st>>>(n, (float*)params, (const float*)grads, (float*)exp_avg, (float*)exp_avg_sq, 0.0, beta1, beta2, eps, (const StepState*)state);
        else adam_kernel<float, false, true><<<(int)grid, block, 0, st>>>(n, (float*)params, (const float*)grads, (float*)exp_avg, (float*)exp_avg_sq, 0.0, beta1, beta2, eps, (const StepState*)state);
        TQ_LAUNCH_CHECK("adam_kernel (deferred flush) launch");
    }
    step_clear_pending_kernel<<<1, 1, 0, st>>>((StepState*)state);
    TQ_LAUNCH_CHECK("step_clear_pending_kernel launch");
    return TQ_OK;
}

extern "C" int tq_step_advance(void* state, void* stream) {
    TQ_CHECK_ARG(state != nullptr, "NULL pointer");
    step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((StepState*)state);
    TQ_LAUNCH_CHECK("step_advance_kernel launch");
    return TQ_OK;
}

// ---- minibatch subsampling (pyro.plate(subsample_size=...) [third party]: randperm(size)[:n]) -----------
// A uniform ordered sample without replacement = the n_pick smallest of n_total i.i.d. random keys, in key order.
// Every block derives all keys (Philox keyed by (seed, stream, step), counter = element) into shared memory and each
// thread ranks its own element against them (shared-memory broadcast reads): a few microseconds for the sizes of a
// minibatch axis, where a serial Fisher-Yates took 85 us for 512 of 1000 frames -- most of the reference-default
// 10 x 512 step.  Axes longer than kSubsampleMaxRank keep the serial partial Fisher-Yates on a persistent permutation
// (after n_pick swaps the first n_pick entries are a uniform sample whatever permutation the array held before).
namespace tq {
constexpr int kSubsampleMaxRank = 16384;
constexpr int kSubsampleBlock = 256;

__global__ void __launch_bounds__(kSubsampleBlock) subsample_rank_kernel(int n_total, int n_pick, unsigned long long seed,
                                                                         const StepState* state, unsigned long long stream_id,
                                                                         int32_t* __restrict__ out) {
    extern __shared__ unsigned int keys[];
    const unsigned long long sd = seed ^ (0x9E3779B97F4A7C15ull * (stream_id + 1ull)), step = state->step;
    for (int j = threadIdx.x; j < n_total; j += kSubsampleBlock) {
        Philox rng(sd, step, (unsigned long long)j);
        keys[j] = rng.next();
    }
    __syncthreads();
    const int i = blockIdx.x * kSubsampleBlock + threadIdx.x;
    if (i >= n_total) return;
    const unsigned int ki = keys[i];
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < n_total; ++j) {
        const unsigned int kj = keys[j];
        rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;   // ties (2^-32 per pair) broken by position
    }
    if (rank < n_pick) out[rank] = i;
}

// the AOI and the frame draw of one step in ONE launch (blockIdx.y selects the axis): the reference-default minibatch
// step is launch-latency bound, and the two draws head its critical path
struct SubsampleAxis { int n_total, n_pick; unsigned long long stream_id; int32_t* out; };
__global__ void __launch_bounds__(kSubsampleBlock) subsample_rank_pair_kernel(SubsampleAxis ax0, SubsampleAxis ax1, unsigned long long seed,
                                                                              const StepState* state) {
    extern __shared__ unsigned int keys[];
    const SubsampleAxis ax = blockIdx.y == 0 ? ax0 : ax1;
    if ((int)(blockIdx.x * kSubsampleBlock) >= ax.n_total) return;
    const unsigned long long sd = seed ^ (0x9E3779B97F4A7C15ull * (ax.stream_id + 1ull)), step = state->step;
    for (int j = threadIdx.x; j < ax.n_total; j += kSubsampleBlock) {
        Philox rng(sd, step, (unsigned long long)j);
        keys[j] = rng.next();
    }
    __syncthreads();
    const int i = blockIdx.x * kSubsampleBlock + threadIdx.x;
    if (i >= ax.n_total) return;
    const unsigned int ki = keys[i];
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j < ax.n_total; ++j) {
        const unsigned int kj = keys[j];
        rank += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
    }
    if (rank < ax.n_pick) ax.out[rank] = i;
}

__global__ void subsample_kernel(int n_total, int n_pick, unsigned long long seed, const StepState* state,
                                 unsigned long long stream_id, int32_t* __restrict__ perm, int32_t* __restrict__ out) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    Philox rng(seed ^ (0x9E3779B97F4A7C15ull * (stream_id + 1ull)), state->step, 0ull);
    for (int i = 0; i < n_pick; ++i) {
        const unsigned int span = (unsigned int)(n_total - i);
        const unsigned int j = i + (unsigned int)(((unsigned long long)rng.next() * span) >> 32);
        const int32_t t = perm[i];
        perm[i] = perm[j];
        perm[j] = t;
        out[i] = perm[i];
    }
}
}  // namespace tq

// perm: (n_total,) int32 scratch holding a permutation of 0..n_total-1 (initialise to arange once; only touched when
// n_total > 16384); out: (n_pick,) int32.  stream_id separates independent draws (AOIs vs frames, ranks).
extern "C" int tq_subsample(int n_total, int n_pick, uint64_t seed, const void* state, uint64_t stream_id, void* perm,
                            void* out, void* stream) {
    TQ_CHECK_ARG(n_total >= 1 && n_pick >= 0 && n_pick <= n_total, "bad sizes");
    TQ_CHECK_ARG(state && perm && out, "NULL pointer");
    if (n_pick == 0) return TQ_OK;
    if (n_total <= tq::kSubsampleMaxRank) {
        const size_t smem = sizeof(unsigned int) * (size_t)n_total;
        if (smem > 48 * 1024) {
            int st2 = tq::cuda_status(cudaFuncSetAttribute(tq::subsample_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                      "cudaFuncSetAttribute(subsample_rank)");
            if (st2 != TQ_OK) return st2;
        }
        const int grid = (n_total + tq::kSubsampleBlock - 1) / tq::kSubsampleBlock;
        tq::subsample_rank_kernel<<<grid, tq::kSubsampleBlock, smem, (cudaStream_t)stream>>>(n_total, n_pick, seed, (const tq::StepState*)state,
                                                                                           stream_id, (int32_t*)out);
        TQ_LAUNCH_CHECK("subsample_rank_kernel launch");
        return TQ_OK;
    }
    tq::subsample_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(n_total, n_pick, seed, (const tq::StepState*)state, stream_id,
                                                          (int32_t*)perm, (int32_t*)out);
    TQ_LAUNCH_CHECK("subsample_kernel launch");
    return TQ_OK;
}

// Two independent draws (tq_subsample semantics, same keys: identical results) in one launch; both axes must be
// <= 16384 long (tq_subsample_pair_supported), else call tq_subsample twice.
extern "C" int tq_subsample_pair_supported(int n_total0, int n_total1) {
    return n_total0 >= 1 && n_total1 >= 1 && n_total0 <= tq::kSubsampleMaxRank && n_total1 <= tq::kSubsampleMaxRank ? 1 : 0;
}
extern "C" int tq_subsample_pair(int n_total0, int n_pick0, uint64_t stream_id0, void* out0, int n_total1, int n_pick1,
                                 uint64_t stream_id1, void* out1, uint64_t seed, const void* state, void* stream) {
    TQ_CHECK_ARG(tq_subsample_pair_supported(n_total0, n_total1), "axes must hold 1..16384 entries");
    TQ_CHECK_ARG(n_pick0 >= 1 && n_pick0 <= n_total0 && n_pick1 >= 1 && n_pick1 <= n_total1, "bad sizes");
    TQ_CHECK_ARG(state && out0 && out1, "NULL pointer");
    const int n_max = n_total0 > n_total1 ? n_total0 : n_total1;
    const size_t smem = sizeof(unsigned int) * (size_t)n_max;
    if (smem > 48 * 1024) {
        int st2 = tq::cuda_status(cudaFuncSetAttribute(tq::subsample_rank_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                  "cudaFuncSetAttribute(subsample_rank_pair)");
        if (st2 != TQ_OK) return st2;
    }
    const dim3 grid((n_max + tq::kSubsampleBlock - 1) / tq::kSubsampleBlock, 2);
    tq::subsample_rank_pair_kernel<<<grid, tq::kSubsampleBlock, smem, (cudaStream_t)stream>>>(
        tq::SubsampleAxis{n_total0, n_pick0, stream_id0, (int32_t*)out0}, tq::SubsampleAxis{n_total1, n_pick1, stream_id1, (int32_t*)out1},
        seed, (const tq::StepState*)state);
    TQ_LAUNCH_CHECK("subsample_rank_pair_kernel launch");
    return TQ_OK;
}
