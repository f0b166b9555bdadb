// One persistent kernel for the AOI-local part of a cosmos SVI step: guide sites -> rendered likelihood (forward +
// reverse) -> priors, (z, theta) log-sum-exp, chain rule and the cross-unit sums.  Replaces, for dtype float, the
// sequence  site_fast_kernel (+ site_worklist_kernel) -> ksmogn_stream_kernel -> local_post_kernel  of cosmos_step.cu /
// ksmogn.cu and the 340 B/unit of SoA scratch they hand to each other through HBM (samples, q(m), 58-float site
// records, L, sample gradients): here those live in shared memory for the 64 units a block works on.
//
// Reference path: what Pyro's SVI.step evaluates for models/cosmos.py:329-462 (guide) and :82-327 (model) under
// TraceEnum_ELBO (cosmos.py:600-607); the arithmetic is unchanged -- site_eval_fast_t
// (cosmos_sites_fast.cuh), the packed pixel sweeps (ksmogn_fast.cuh / ksmogn_sweep.cuh), unit_post (cosmos_local.cuh).
//
// Work unit = a BATCH: 64 consecutive units of ONE minibatch AOI (all its channels interleaved), handed out through a
// device counter (the first batch of a block by position).  Per batch a block of 4 warps runs three phases:
//   S  thread per (site, unit), site-major so a warp evaluates one family: 9 x 64 site evaluations in 4.5 passes; sites
//      outside the bulk regime are collected and replayed together, sites outside every fp32 form go through the
//      double-precision site_eval (rare: none on a trained C2 model);
//   L  the likelihood sweep exactly as in ksmogn_stream_kernel: 8 lanes per patch, 4 patches per warp, 4 rounds; the
//      pixels of the next round (or of the next batch's first round) are in flight by cp.async meanwhile;
//   P  thread per unit: unit_post, gradients of the 18 AOI-local parameters of the unit straight to HBM, block sums of
//      the accumulators; the last batch of an AOI adds that AOI's batches in index order, the last AOI of the launch adds
//      the AOIs in index order (tickets, as in local_post_kernel): fixed summation order, run-to-run deterministic.
// The four resident blocks of an SM are in different phases at any time, so the MUFU / FP64 / latency-bound site and post
// phases of one block fill the issue slots that the FMA-bound sweep of another leaves idle.
#include "cosmos_step_common.cuh"
#include "ksmogn_sweep.cuh"
#include <mutex>
#include <unordered_map>

namespace tq {

constexpr int kFB = 64;                        // units per batch
constexpr int kFThreads = kWarpsPerBlock * 32; // 128
constexpr int kFSlots = kFThreads / kSub;      // 16 patches swept concurrently
constexpr int kFRounds = kFB / kFSlots;        // 4
constexpr int kFStage = 416;                   // bytes per staging slot: 392 B of pixels, 8 B target position, padding
constexpr int kFUnitRows = NSAMP + kM + kM + NSAMP + 1;   // samples, q(m), L, sample gradients, rate gradient

struct FusedArgs {
    tq_patch_view v;
    LocalOffsets lo;
    ModelConst mc;
    const float* lparams;
    const GlobalTables<double>* tables;
    const float* gain;
    int64_t aoi_offset;
    unsigned long long seed;
    const StepState* state;
    const float* noise_in;        // (NSAMP, U) base variates or NULL -> Philox
    double sN, sF;
    float* lgrads;
    double* block_partial;        // (n_batches + nb, C, kPostRed)
    unsigned int* tickets;        // (nb + 1), zero, left zeroed
    double* acc_out;              // (C, NACC)
    unsigned int* counters;       // {batches handed out beyond the first per block, blocks finished}
    int chunks;                   // batches per minibatch AOI
    uint32_t n_batches;
    uint32_t row;                 // fb * C: units per minibatch AOI
    int64_t U;
    // optional copies of the per-unit intermediates for tests / compute_probs-style consumers (NULL in production)
    float* samples_out;           // (NSAMP, U)
    float* L_out;                 // (kM, U)
};

// strided view of one unit's column in a (rows, kFB) shared-memory array
struct SmemCol {
    const float* base;
    __device__ __forceinline__ float operator[](int i) const { return base[i * kFB]; }
};

struct BatchLoc {
    uint32_t ni, j0, count;       // minibatch AOI, first unit within the AOI's row, live units
    int64_t n;                    // store AOI
    int64_t u0;                   // first unit's index in (NSAMP, U)-shaped arrays
};

__device__ __forceinline__ BatchLoc locate_batch(const FusedArgs& a, uint32_t b) {
    BatchLoc L;
    L.ni = b / (uint32_t)a.chunks;
    L.j0 = (b - L.ni * (uint32_t)a.chunks) * (uint32_t)kFB;
    L.count = min((uint32_t)kFB, a.row - L.j0);
    L.n = a.v.ndx ? a.v.ndx[L.ni] : (int64_t)L.ni;
    L.u0 = (int64_t)L.ni * a.row + L.j0;
    return L;
}

// unit t of a batch -> frame, channel, store patch
struct UnitLoc { int fi, c; int64_t f, patch; };
__device__ __forceinline__ UnitLoc locate_in_batch(const FusedArgs& a, const BatchLoc& B, uint32_t t) {
    UnitLoc u;
    const uint32_t idx = B.j0 + t;
    if (a.v.C == 1) { u.fi = (int)idx; u.c = 0; }
    else { u.fi = (int)(idx / (uint32_t)a.v.C); u.c = (int)(idx - (uint32_t)u.fi * (uint32_t)a.v.C); }
    u.f = a.v.fdx ? a.v.fdx[u.fi] : u.fi;
    u.patch = (B.n * a.v.F + u.f) * a.v.C + u.c;
    return u;
}

struct FusedSite { float p0, p1, pbm, pbs; unsigned long long rng_offset; };
__device__ __forceinline__ FusedSite fused_site_gather(const FusedArgs& a, const BatchLoc& B, const UnitLoc& u, int s) {
    FusedSite in;
    in.p0 = a.lparams[a.lo.index(site_param0(s), B.n, u.f, u.c)];
    in.p1 = a.lparams[a.lo.index(site_param1(s), B.n, u.f, u.c)];
    in.pbm = in.pbs = 0.0f;
    if (s == S_B) {
        in.pbm = a.lparams[a.lo.index(LP_BM, B.n, u.f, u.c)];
        in.pbs = a.lparams[a.lo.index(LP_BS, B.n, u.f, u.c)];
    }
    const unsigned long long gid = (((unsigned long long)(a.aoi_offset + B.n)) * a.v.F + u.f) * a.v.C + u.c;
    in.rng_offset = ((gid + 1ull) << 12) + ((unsigned long long)s << 8);   // same stream identity as site_gather
    return in;
}

__device__ __forceinline__ float fused_into_support(int s, float v, const ModelConst& mc) {
    if (site_is_gamma(s)) return v > 1.17549435e-38f ? v : 1.17549435e-38f;
    const float lo = s < S_X ? (float)mc.width_min : -0.5f * float(mc.P + 1), hi = s < S_X ? (float)mc.width_max : 0.5f * float(mc.P + 1);
    const float margin = (hi - lo) * 1.1920929e-7f;
    return v < lo + margin ? lo + margin : (v > hi - margin ? hi - margin : v);
}

// shared-memory columns of unit t
struct UnitSmem {
    float* samples;   // (NSAMP, kFB)
    float* qm;        // (kM, kFB)
    float* Lm;        // (kM, kFB)
    float* gs;        // (NSAMP, kFB)
    float* g_rate;    // (kFB)
    float* rec;       // (NREC, kFB)
};

__device__ __forceinline__ void fused_site_store(const UnitSmem& sm, int s, uint32_t t, float v, const float* rec, const float* extra,
                                                 const ModelConst& mc) {
    sm.samples[s * kFB + t] = fused_into_support(s, v, mc);
#pragma unroll
    for (int j = 0; j < NSO; ++j) sm.rec[(s * NSO + j) * kFB + t] = rec[j];
    if (s == S_B) {
#pragma unroll
        for (int j = 0; j < NEX; ++j) sm.rec[(NSAMP * NSO + j) * kFB + t] = extra[j];
    }
}

// double-precision form of one site (cosmos_local.cuh::site_eval), for what no fp32 form covers; same Philox stream,
// so the same draw.  Not inlined: its ~3000 instructions and 100+ registers stay out of the phases' main line.
__device__ __noinline__ void fused_site_double(const FusedArgs& a, UnitSmem sm, int s, uint32_t t, FusedSite in, int64_t u) {
    const bool use_rng = a.noise_in == nullptr;
    Philox rng(a.seed, a.state->step, in.rng_offset);
    double variate = use_rng ? 0.0 : (double)a.noise_in[(int64_t)s * a.U + u];
    double drec[NSO], dextra[NEX];
    const float v = (float)site_eval(s, (double)in.p0, (double)in.p1, (double)in.pbm, (double)in.pbs, a.mc, use_rng, &rng, variate, drec, dextra);
    float rec[NSO], extra[NEX];
#pragma unroll
    for (int j = 0; j < NSO; ++j) rec[j] = (float)drec[j];
#pragma unroll
    for (int j = 0; j < NEX; ++j) extra[j] = (float)dextra[j];
    fused_site_store(sm, s, t, v, rec, extra, a.mc);
}

template <typename PIX, int OC>
__global__ void __launch_bounds__(kFThreads, 4) cosmos_fused_kernel(const FusedArgs a) {
    static_assert(sizeof(PIX) == 2, "the fused kernel stages uint16 pixels");
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int O = a.v.O, offpad = (2 * O + 3) & ~3;
    float* off_s = reinterpret_cast<float*>(smem_raw);
    float* off_w2 = off_s + O;
    float* tabs = off_s + offpad;                                        // kFSlots x kTabFloats
    float* spx_all = tabs + kFSlots * kTabFloats;                  // kFSlots x 196 (phase L); scratch in S and P
    float* unit_rows = spx_all + kFSlots * 196;                          // kFUnitRows x kFB
    float* rec_rows = unit_rows + kFUnitRows * kFB;                      // NREC x kFB
    unsigned char* stage_all = reinterpret_cast<unsigned char*>(rec_rows + NREC * kFB);   // kFSlots x kFStage
    __shared__ GlobalTables<float> gt;
    __shared__ unsigned int s_next, s_ndef, s_nfb;
    __shared__ int last_flags[2];

    UnitSmem sm;
    sm.samples = unit_rows;
    sm.qm = sm.samples + NSAMP * kFB;
    sm.Lm = sm.qm + kM * kFB;
    sm.gs = sm.Lm + kM * kFB;
    sm.g_rate = sm.gs + NSAMP * kFB;
    sm.rec = rec_rows;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot = tid / kSub, sub = tid % kSub, wslot = lane / kSub;
    float* tab = tabs + slot * kTabFloats;
    float* gx = tab;
    float* gy = gx + kK * kMaxP;
    float* spx = spx_all + slot * 196;
    unsigned char* stage = stage_all + slot * kFStage;

    for (int j = tid; j < O; j += kFThreads) {
        off_s[j] = static_cast<const float*>(a.v.offset_samples)[j];
        off_w2[j] = static_cast<const float*>(a.v.offset_logits)[j] * kLog2e;
    }
    if (tid == 0) gt.convert_from(*a.tables);
    FastConst fc;
    fc.gain = a.gain[0];
    fc.rate = 1.0f / fc.gain;
    fc.rate2 = fc.rate * kLog2e;
    fc.log_rate = logf(fc.rate);

    const PIX* pixels = static_cast<const PIX*>(a.v.pixels);
    const float* xy = static_cast<const float*>(a.v.xy);
    const bool use_rng = a.noise_in == nullptr;
    const int C = a.v.C;

    // pixels (49 x 8 B) and target position (8 B) of this slot's patch in round r of batch B: issued one round ahead
    auto prefetch = [&](const BatchLoc& B, int r) {
        const uint32_t t_raw = (uint32_t)(r * kFSlots + slot), t = t_raw < B.count ? t_raw : B.count - 1u;
        const UnitLoc u = locate_in_batch(a, B, t);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(pixels + u.patch * 196);
#pragma unroll
        for (int i7 = 0; i7 < 7; ++i7) {
            const int i = sub + i7 * kSub;
            if (i < 49) cp_async_8(stage + 8 * i, src + 8 * i);
        }
        if (sub == 0) cp_async_8(stage + 392, xy + u.patch * 2);
    };

    uint32_t cur = blockIdx.x;
    if (tid == 0) s_next = gridDim.x + atomicAdd(a.counters, 1u);
    if (cur < a.n_batches) prefetch(locate_batch(a, cur), 0);
    __syncthreads();

    while (cur < a.n_batches) {
        const BatchLoc B = locate_batch(a, cur);
        const uint32_t nxt = s_next;
        if (tid == 0) { s_ndef = 0u; s_nfb = 0u; }
        __syncthreads();                                               // s_next read by everyone; lists reset
        if (tid == 0) s_next = gridDim.x + atomicAdd(a.counters, 1u);  // the batch after next: latency hidden by this batch

        // =============================== phase S: guide sites ============================================================
        // deferred / fallback lists live in the pixel area (unused until phase L)
        double* s_var = reinterpret_cast<double*>(spx_all);                              // kFB * NSAMP doubles max
        unsigned short* s_task = reinterpret_cast<unsigned short*>(s_var + kFB * NSAMP); // task ids of deferred sites
        unsigned short* s_fb = s_task + kFB * NSAMP;                                     // task ids of fallback sites
        for (int task = tid; task < NSAMP * kFB; task += kFThreads) {
            const int s = task / kFB;
            const uint32_t t = (uint32_t)(task - s * kFB);
            if (t >= B.count) continue;
            const UnitLoc u = locate_in_batch(a, B, t);
            const FusedSite in = fused_site_gather(a, B, u, s);
            Philox rng(a.seed, a.state->step, in.rng_offset);
            double variate = use_rng ? 0.0 : (double)a.noise_in[(int64_t)s * a.U + B.u0 + t];
            float v = 0.0f, rec[NSO], extra[NEX];
            int cls = 0;
            const int status = site_eval_fast_t<1>(s, in.p0, in.p1, in.pbm, in.pbs, a.mc, use_rng, &rng, variate, v, rec, extra, cls);
            if (status == SITE_DONE) {
                fused_site_store(sm, s, t, v, rec, extra, a.mc);
            } else if (status == SITE_DEFER) {
                const unsigned int pos = atomicAdd(&s_ndef, 1u);
                s_var[pos] = variate;
                s_task[pos] = (unsigned short)task;
            } else {
                s_fb[atomicAdd(&s_nfb, 1u)] = (unsigned short)task;
            }
            if (s == S_B) {   // q(m) from the unconstrained m_probs: the weights of the likelihood's reverse mode
                float q1[kK], q0[kK], qm[kM];
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    const SpotPresence<float> sp(a.lparams[a.lo.index(LP_M_PROBS + k, B.n, u.f, u.c)], a.mc);
                    q1[k] = sp.q1; q0[k] = sp.q0;
                }
                presence_weights<float>(q1, q0, qm);
#pragma unroll
                for (int m = 0; m < kM; ++m) sm.qm[m * kFB + t] = qm[m];
            }
        }
        __syncthreads();
        if (s_ndef) {   // sites outside the bulk regime (none at the initial point), replayed side by side
            const unsigned int nd = s_ndef;
            for (unsigned int i = tid; i < nd; i += kFThreads) {
                const int task = s_task[i], s = task / kFB;
                const uint32_t t = (uint32_t)(task - s * kFB);
                const UnitLoc u = locate_in_batch(a, B, t);
                const FusedSite in = fused_site_gather(a, B, u, s);
                double var = s_var[i];
                float v = 0.0f, rec[NSO], extra[NEX];
                int cls = 0;
                const int status = site_eval_fast_t<2>(s, in.p0, in.p1, in.pbm, in.pbs, a.mc, false, nullptr, var, v, rec, extra, cls);
                if (status == SITE_DONE) fused_site_store(sm, s, t, v, rec, extra, a.mc);
                else s_fb[atomicAdd(&s_nfb, 1u)] = (unsigned short)task;
            }
            __syncthreads();
        }
        if (s_nfb) {    // outside every fp32 form: the double-precision site_eval
            const unsigned int nf = s_nfb;
            for (unsigned int i = tid; i < nf; i += kFThreads) {
                const int task = s_fb[i], s = task / kFB;
                const uint32_t t = (uint32_t)(task - s * kFB);
                const UnitLoc u = locate_in_batch(a, B, t);
                fused_site_double(a, sm, s, t, fused_site_gather(a, B, u, s), B.u0 + t);
            }
            __syncthreads();
        }
        if (a.samples_out) {
            for (int task = tid; task < NSAMP * kFB; task += kFThreads) {
                const int s = task / kFB, t = task - s * kFB;
                if ((uint32_t)t < B.count) a.samples_out[(int64_t)s * a.U + B.u0 + t] = sm.samples[task];
            }
        }

        // =============================== phase L: rendered likelihood, forward + reverse ===================================
#pragma unroll 1
        for (int r = 0; r < kFRounds; ++r) {
            const uint32_t t_raw = (uint32_t)(r * kFSlots + slot);
            const bool live = t_raw < B.count;
            const uint32_t t = live ? t_raw : B.count - 1u;   // idle slots shadow the last unit and write nothing
            const bool group_live = (uint32_t)(r * kFSlots + warp * 4) < B.count;
            cp_async_wait_all();
            __syncwarp();
            float pix_min = 3.0e38f;
            float tx, ty;
            {
                const uint2* raw = reinterpret_cast<const uint2*>(stage);
#pragma unroll
                for (int i7 = 0; i7 < 7; ++i7) {
                    const int i = sub + i7 * kSub;
                    if (i < 49) {
                        const uint2 q = raw[i];
                        const float4 v = make_float4(float(q.x & 0xffffu), float(q.x >> 16), float(q.y & 0xffffu), float(q.y >> 16));
                        reinterpret_cast<float4*>(spx)[i] = v;
                        pix_min = fminf(pix_min, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
                    }
                }
                const float2 txy = *reinterpret_cast<const float2*>(stage + 392);
                tx = txy.x; ty = txy.y;
            }
            __syncwarp();   // the staging area has been read by every lane: the next round may land in it
            if (r + 1 < kFRounds) prefetch(B, r + 1);
            else if (nxt < a.n_batches) prefetch(locate_batch(a, nxt), 0);
            if (!group_live) continue;

            PatchSpots<float> s;
            float W[kM], Wr[kM];
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                s.h[k] = sm.samples[(S_H + k) * kFB + t];
                s.w[k] = sm.samples[(S_W + k) * kFB + t];
                s.cx[k] = sm.samples[(S_X + k) * kFB + t] + tx;
                s.cy[k] = sm.samples[(S_Y + k) * kFB + t] + ty;
            }
            s.b = sm.samples[S_B * kFB + t];
#pragma unroll
            for (int m = 0; m < kM; ++m) { W[m] = sm.qm[m * kFB + t]; Wr[m] = W[m] * fc.rate; }

            const bool small = s.b * fc.rate < 4.0f;
            bool pairs = false;
            if (OC > 0) {
                float max_off = off_s[0];
                for (int j = 1; j < OC; ++j) max_off = fmaxf(max_off, off_s[j]);
                pairs = __all_sync(kFull, !small && pix_min > max_off);
            }
            // separable spot factors: 2 K P exponentials per patch instead of K P P
            float norm[kK];
            if (pairs) build_tables_pairs(tab, sub, s);
            else build_tables_generic(tab, 14, sub, s, norm);
            __syncwarp();

            PatchOut<float, kM> out;
            out.zero();
            if (pairs && OC == 1)
                sweep_patch_rows_single_bin(spx, sub, tab, s, fc, off_s[0], off_w2[0] * kLn2, W, out);
            else if (pairs)
                sweep_patch_pairs<OC>(spx, sub, tab, s, fc, off_s, off_w2, W, out);
            else if (__any_sync(kFull, small))
                sweep_patch<OC, true, true, true>(spx, 14, 196, sub, gx, gy, s, norm, fc, O, off_s, off_w2, W, Wr, out);
            else
                sweep_patch<OC, true, true, false>(spx, 14, 196, sub, gx, gy, s, norm, fc, O, off_s, off_w2, W, Wr, out);

#pragma unroll
            for (int m = 0; m < kM; ++m) out.logp[m] = sub_sum(out.logp[m]);
            out.g_b = sub_sum(out.g_b);
            out.g_rate = sub_sum(out.g_rate);
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                out.g_h[k] = sub_sum(out.g_h[k]);
                out.g_w[k] = sub_sum(out.g_w[k]);
                out.g_x[k] = sub_sum(out.g_x[k]);
                out.g_y[k] = sub_sum(out.g_y[k]);
            }
            finish_spot_moments(s, out);
            if (live && sub == 0) {
#pragma unroll
                for (int m = 0; m < kM; ++m) sm.Lm[m * kFB + t] = out.logp[m];
                sm.gs[S_B * kFB + t] = out.g_b;
                sm.g_rate[t] = out.g_rate;
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    sm.gs[(S_H + k) * kFB + t] = out.g_h[k];
                    sm.gs[(S_W + k) * kFB + t] = out.g_w[k];
                    sm.gs[(S_X + k) * kFB + t] = out.g_x[k];
                    sm.gs[(S_Y + k) * kFB + t] = out.g_y[k];
                }
            }
        }
        __syncthreads();

        // =============================== phase P: priors, (z, theta) sums, chain rule, reductions ==========================
        const double mu = a.v.mask[B.n] ? 1.0 : 0.0;
        const bool ontarget = a.v.is_ontarget[B.n] != 0;
        float sum[kPostRed];
#pragma unroll
        for (int i = 0; i < kPostRed; ++i) sum[i] = 0.0f;
        int my_c = 0;
        if ((uint32_t)tid < B.count) {
            const uint32_t t = (uint32_t)tid;
            const UnitLoc u = locate_in_batch(a, B, t);
            my_c = u.c;
            float L[kM], u_mp[kK];
#pragma unroll
            for (int m = 0; m < kM; ++m) L[m] = sm.Lm[m * kFB + t];
#pragma unroll
            for (int k = 0; k < kK; ++k) u_mp[k] = a.lparams[a.lo.index(LP_M_PROBS + k, B.n, u.f, u.c)];
            const float u_bm = a.lparams[a.lo.index(LP_BM, B.n, 0, u.c)], u_bs = a.lparams[a.lo.index(LP_BS, B.n, 0, u.c)];
            if (a.L_out) {
#pragma unroll
                for (int m = 0; m < kM; ++m) a.L_out[(int64_t)m * a.U + B.u0 + t] = L[m];
            }
            UnitGrads<float> ug;
            unit_post<float>(SmemCol{sm.rec + t}, SmemCol{sm.samples + t}, L, SmemCol{sm.gs + t}, sm.g_rate[t], u_mp, u_bm, u_bs,
                             a.mc, gt, u.c, ontarget, u.fi == 0, ug);
            const float scale = (float)(-a.sN * a.sF * mu);   // loss = -ELBO
#pragma unroll
            for (int i = LP_B_LOC; i < NLOCAL; ++i) a.lgrads[a.lo.index(i, B.n, u.f, u.c)] = scale * ug.g[i];
#pragma unroll
            for (int i = 0; i < NACC; ++i) sum[i] = ug.acc[i];
            sum[NACC] = ug.g[LP_BM];
            sum[NACC + 1] = ug.g[LP_BS];
        }
        // block sums per channel, fixed shuffle order, in double
        double* red = reinterpret_cast<double*>(spx_all);   // [warp][C][kPostRed]
        for (int cc = 0; cc < C; ++cc) {
#pragma unroll
            for (int i = 0; i < kPostRed; ++i) {
                const double v = warp_sum((double)((my_c == cc) ? sum[i] : 0.0f));
                if (lane == 0) red[(warp * C + cc) * kPostRed + i] = v;
            }
        }
        __syncthreads();
        if (tid < C * kPostRed) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kWarpsPerBlock; ++w) v += red[w * C * kPostRed + tid];
            a.block_partial[(int64_t)cur * C * kPostRed + tid] = mu * v;
        }
        // publish (barrier, then thread 0's fence + ticket: the pattern of a cooperative-groups grid barrier)
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            const unsigned int t1 = atomicAdd(&a.tickets[B.ni], 1u);
            last_flags[0] = (t1 == (unsigned int)a.chunks - 1u);
            if (last_flags[0]) { a.tickets[B.ni] = 0u; __threadfence(); }
        }
        __syncthreads();
        if (last_flags[0]) {
            // ---- last batch of this minibatch AOI: add its batches in index order ------------------------------------------
            double* aoi_sums = a.block_partial + (int64_t)a.n_batches * C * kPostRed;   // (nb, C, kPostRed)
            if (tid < C * kPostRed) {
                const int cc = tid / kPostRed, i = tid - cc * kPostRed;
                const double* bp = a.block_partial + (int64_t)B.ni * a.chunks * C * kPostRed + tid;
                double v = 0.0;
                for (int k0 = 0; k0 < a.chunks; k0 += 8) {
                    double tv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) tv[j] = __ldcg(bp + (int64_t)min(k0 + j, a.chunks - 1) * C * kPostRed);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v += (k0 + j < a.chunks) ? tv[j] : 0.0;
                }
                aoi_sums[(int64_t)B.ni * C * kPostRed + tid] = v;
                if (i >= NACC) {
                    // d loss / d (background_mean_loc, background_std_loc)[n, 0, c]: its frames + the AOI-level prior
                    const float u_bm = a.lparams[a.lo.index(LP_BM, B.n, 0, cc)], u_bs = a.lparams[a.lo.index(LP_BS, B.n, 0, cc)];
                    double pbm, pbs;
                    aoi_prior_grad((double)u_bm, (double)u_bs, a.mc, pbm, pbs);
                    const bool is_bm = i == NACC;
                    a.lgrads[a.lo.index(is_bm ? LP_BM : LP_BS, B.n, 0, cc)] = (float)(-(a.sN * a.sF * v + a.sN * mu * (is_bm ? pbm : pbs)));
                }
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                const unsigned int t2 = atomicAdd(&a.tickets[a.v.nb], 1u);
                last_flags[1] = (t2 == (unsigned int)a.v.nb - 1u);
                if (last_flags[1]) { a.tickets[a.v.nb] = 0u; __threadfence(); }
            }
            __syncthreads();
            if (last_flags[1]) {
                // ---- last AOI of the launch: channel accumulators = sum over AOIs in index order -------------------------------
                constexpr int kParts = 4;
                double* fin = red;   // [C * NACC][kParts]
                for (int w = tid; w < C * NACC * kParts; w += kFThreads) {
                    const int part = w % kParts, vi = w / kParts;
                    const int cc = vi / NACC, i = vi - cc * NACC;
                    double v = 0.0;
                    for (int q0 = part; q0 < a.v.nb; q0 += kParts * 8) {
                        double tv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int q = min(q0 + kParts * j, a.v.nb - 1);
                            tv[j] = __ldcg(aoi_sums + ((int64_t)q * C + cc) * kPostRed + i);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) v += (q0 + kParts * j < a.v.nb) ? tv[j] : 0.0;
                    }
                    fin[vi * kParts + part] = v;
                }
                __syncthreads();
                if (tid < C * NACC) a.acc_out[tid] = (fin[tid * kParts + 0] + fin[tid * kParts + 1]) + (fin[tid * kParts + 2] + fin[tid * kParts + 3]);
            }
        }
        __syncthreads();   // shared memory is reused by the next batch
        cur = nxt;
    }
    if (tid == 0) {
        // the last block out re-arms the counters for the next launch on this stream
        __threadfence();
        const unsigned int done = atomicAdd(a.counters + 1, 1u);
        if (done == gridDim.x - 1u) {
            a.counters[0] = 0u;
            a.counters[1] = 0u;
        }
    }
}

// work counters: one set per stream (launches on one stream are ordered and can share; see ksmogn.cu)
constexpr int kFusedCounterSets = 64;
__device__ unsigned int g_fused_counters[kFusedCounterSets][2];

static int fused_counter_set_of(cudaStream_t st) {
    static std::mutex mu;
    static std::unordered_map<cudaStream_t, int> sets;
    std::lock_guard<std::mutex> lock(mu);
    auto it = sets.find(st);
    if (it != sets.end()) return it->second;
    const int k = (int)(sets.size() % kFusedCounterSets);
    sets.emplace(st, k);
    return k;
}

static size_t fused_smem_bytes(int O) {
    const int offpad = (2 * O + 3) & ~3;
    return sizeof(float) * ((size_t)offpad + kFSlots * kTabFloats + kFSlots * 196 + (size_t)(kFUnitRows + NREC) * kFB) +
           (size_t)kFSlots * kFStage;
}

template <int OC>
static int launch_fused(FusedArgs& a, cudaStream_t st) {
    auto kern = cosmos_fused_kernel<uint16_t, OC>;
    const size_t smem = fused_smem_bytes(a.v.O);
    static_assert(sizeof(double) * kFB * NSAMP + 2 * sizeof(unsigned short) * kFB * NSAMP <= sizeof(float) * kFSlots * 196,
                  "deferred-site lists must fit the pixel area");
    int st2 = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                          "cudaFuncSetAttribute(cosmos_fused)");
    if (st2 != TQ_OK) return st2;
    void* base = nullptr;
    st2 = cuda_status(cudaGetSymbolAddress(&base, g_fused_counters), "cudaGetSymbolAddress(g_fused_counters)");
    if (st2 != TQ_OK) return st2;
    a.counters = static_cast<unsigned int*>(base) + 2 * fused_counter_set_of(st);
    const int64_t cap = (int64_t)sm_count() * 4;
    const int grid = (int)((int64_t)a.n_batches < cap ? (int64_t)a.n_batches : cap);
    kern<<<grid, kFThreads, smem, st>>>(a);
    TQ_LAUNCH_CHECK("cosmos_fused_kernel launch");
    return TQ_OK;
}

}  // namespace tq

using namespace tq;

// scratch sizes for a minibatch of nb AOIs x fb frames x C channels
extern "C" int64_t tq_cosmos_fused_scratch(int nb, int fb, int C) {
    const int64_t row = (int64_t)fb * C, chunks = (row + kFB - 1) / kFB;
    return ((int64_t)nb * chunks + nb) * C * kPostRed;   // doubles: per-batch partials + per-AOI sums
}
extern "C" int64_t tq_cosmos_fused_tickets(int nb, int fb, int C) { (void)fb; (void)C; return ((int64_t)nb + 1 + 1) / 2; }   // 8-byte slots

// 1 when tq_cosmos_fused_step supports this view (else the caller runs the per-stage kernels)
extern "C" int tq_cosmos_fused_supported(int dtype, const tq_patch_view* view) {
    if (!view || dtype != TQ_F32) return 0;
    if (view->P != 14 || view->pixtype != TQ_PIX_U16 || view->O < 1 || view->O > 512) return 0;
    if ((int64_t)view->nb * view->fb * view->C >= ((int64_t)1 << 31)) return 0;
    return fused_smem_bytes(view->O) <= 56 * 1024 ? 1 : 0;   // four resident blocks per SM
}

// AOI-local part of one SVI step (guide sites, likelihood forward + reverse, post, cross-unit sums) in one launch.
// Inputs as tq_cosmos_sites_ws / tq_ksmogn_fwd_bwd / tq_cosmos_local_post; `gain`: 1 float written by
// tq_cosmos_globals_sample; outputs: lgrads (entries of the minibatch), acc (C, NACC).  samples_out (NSAMP, U) and
// L_out (4, U) may be NULL.
extern "C" int tq_cosmos_fused_step(int dtype, const tq_patch_view* view, int64_t Nt, const void* mc, const void* lparams,
                                    const void* tables, const void* gain, int64_t aoi_offset, uint64_t seed, const void* state,
                                    const void* noise_in, double sN, double sF, void* lgrads, double* tickets,
                                    double* block_partial, double* acc, void* samples_out, void* L_out, void* stream) {
    TQ_CHECK_ARG(view && mc && lparams && tables && gain && state, "NULL input pointer");
    TQ_CHECK_ARG(lgrads && tickets && block_partial && acc, "NULL output pointer");
    TQ_CHECK_ARG(view->mask && view->is_ontarget, "view needs mask and is_ontarget");
    TQ_CHECK_ARG(view->C >= 1 && view->C <= kMaxC, "C (channels) must be in [1, 4]");
    if (!tq_cosmos_fused_supported(dtype, view)) {
        set_error("tq_cosmos_fused_step: needs dtype float, P = 14, uint16 pixels, O <= 512 (got dtype %d, P %d, pixtype %d, O %d)",
                  dtype, view->P, view->pixtype, view->O);
        return TQ_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    FusedArgs a{};
    a.v = *view;
    a.lo = LocalOffsets{Nt, (int64_t)view->F, (int64_t)view->C};
    a.mc = *(const ModelConst*)mc;
    a.lparams = (const float*)lparams;
    a.tables = (const GlobalTables<double>*)tables;
    a.gain = (const float*)gain;
    a.aoi_offset = aoi_offset;
    a.seed = seed;
    a.state = (const StepState*)state;
    a.noise_in = (const float*)noise_in;
    a.sN = sN; a.sF = sF;
    a.lgrads = (float*)lgrads;
    a.block_partial = block_partial;
    a.tickets = (unsigned int*)tickets;
    a.acc_out = acc;
    a.row = (uint32_t)view->fb * (uint32_t)view->C;
    a.chunks = (int)((a.row + kFB - 1) / kFB);
    a.U = (int64_t)view->nb * view->fb * view->C;
    a.n_batches = (uint32_t)((int64_t)view->nb * a.chunks);
    a.samples_out = (float*)samples_out;
    a.L_out = (float*)L_out;
    if (a.U == 0) {
        cudaMemsetAsync(acc, 0, sizeof(double) * view->C * NACC, st);
        return TQ_OK;
    }
    switch (view->O <= 4 ? view->O : 0) {
        case 1: return launch_fused<1>(a, st);
        case 2: return launch_fused<2>(a, st);
        case 3: return launch_fused<3>(a, st);
        case 4: return launch_fused<4>(a, st);
        default: return launch_fused<0>(a, st);
    }
}
