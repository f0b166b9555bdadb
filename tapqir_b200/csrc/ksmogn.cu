// K1: fused spot render + offset-marginalised Gamma image likelihood, forward and reverse mode,
// one warp per (AOI, frame, channel) patch.
//
// Replaces distributions/util.py:15-64 (gaussian_spots), distributions/ksmogn.py:146-169 and
// ksmogn.py:187-238 (KSMOGN.log_prob; KeOps Genred LogSumExp or the torch branch) and their
// autograd backward.  The reference materialises (2^K, nb, fb, C, K, P, P, 2) temporaries; here a
// patch's pixels are read once from HBM (392 B as uint16 at P=14), everything else lives in
// registers / a 448 B per-warp row-column table, and the reverse pass is computed in the same
// sweep because its upstream weights are known before the sweep starts.
#include "common.cuh"
#include "ksmogn_core.cuh"
#include "ksmogn_fast.cuh"
#include "ksmogn_sweep.cuh"
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

namespace tq {

template <typename T> struct KsmognArgs {
    tq_patch_view v;
    const T* height; const T* width; const T* x; const T* y; const T* background;
    const T* gain; const T* mcfg; const T* W;
    T* logp; T* g_height; T* g_width; T* g_x; T* g_y; T* g_background; T* g_rate;
    int64_t U;
    int bulk;   // stream kernel, uint16 14 x 14 patches: stage the pixels with 1-D bulk copies (pixel base 16-byte aligned)
};

template <typename PIX> __device__ __forceinline__ float load_pixel_f(const PIX* p) { return (float)*p; }

// smem layout: [off_s (O)] [off_w (O)] [per warp: gx (K*kMaxP), gy (K*kMaxP)]
template <typename T, typename PIX, int NM, bool BWD>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
ksmogn_kernel(const KsmognArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* off_s = reinterpret_cast<T*>(smem_raw);
    T* off_w = off_s + a.v.O;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    T* gx = off_w + a.v.O + warp * (2 * kK * kMaxP);
    T* gy = gx + kK * kMaxP;

    for (int j = threadIdx.x; j < a.v.O; j += blockDim.x) {
        off_s[j] = static_cast<const T*>(a.v.offset_samples)[j];
        off_w[j] = static_cast<const T*>(a.v.offset_logits)[j];
    }
    T mcfg[NM][kK];
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int k = 0; k < kK; ++k) mcfg[m][k] = a.mcfg ? a.mcfg[m * kK + k] : T((m >> k) & 1);
    const T gain = a.gain[0];
    const T rate = T(1) / gain;
    const T log_rate = Real<T>::log(rate);
    __syncthreads();

    const int P = a.v.P, PP = P * P;
    const PIX* pixels = static_cast<const PIX*>(a.v.pixels);
    const T* xy = static_cast<const T*>(a.v.xy);

    for (int64_t u = (int64_t)blockIdx.x * kWarpsPerBlock + warp; u < a.U;
         u += (int64_t)gridDim.x * kWarpsPerBlock) {
        const UnitIndex ui = locate_unit(u, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
        PatchSpots<T> s;
        const T tx = xy[ui.patch * 2 + 0], ty = xy[ui.patch * 2 + 1];
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            s.h[k] = a.height[k * a.U + u];
            s.w[k] = a.width[k * a.U + u];
            s.cx[k] = a.x[k * a.U + u] + tx;
            s.cy[k] = a.y[k * a.U + u] + ty;
        }
        s.b = a.background[u];
        T W[NM];
#pragma unroll
        for (int m = 0; m < NM; ++m) W[m] = BWD ? a.W[m * a.U + u] : T(0);

        // separable spot factors: 2*K*P exponentials per patch instead of K*P*P
        __syncwarp();
        for (int idx = lane; idx < 2 * kK * P; idx += 32) {
            const int axis = idx / (kK * P), rem = idx - axis * (kK * P);
            const int k = rem / P, i = rem - k * P;
            if (axis == 0) gx[k * kMaxP + i] = axis_factor<T>(i, s.cx[k], s.w[k]);
            else           gy[k * kMaxP + i] = axis_factor<T>(i, s.cy[k], s.w[k]);
        }
        __syncwarp();

        PatchOut<T, NM> out;
        out.zero();
        const PIX* pix = pixels + ui.patch * PP;
        for (int p = lane; p < PP; p += 32) {
            const int row = p / P, col = p - row * P;
            T gxk[kK], gyk[kK];
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                gxk[k] = gx[k * kMaxP + col];
                gyk[k] = gy[k * kMaxP + row];
            }
            pixel_accumulate<T, NM, BWD>(T(pix[p]), gxk, gyk, col, row, s, mcfg, rate, log_rate,
                                         a.v.O, off_s, off_w, W, out);
        }
#pragma unroll
        for (int m = 0; m < NM; ++m) out.logp[m] = warp_sum(out.logp[m]);
        if (BWD) {
            out.g_b = warp_sum(out.g_b);
            out.g_rate = warp_sum(out.g_rate);
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                out.g_h[k] = warp_sum(out.g_h[k]);
                out.g_w[k] = warp_sum(out.g_w[k]);
                out.g_x[k] = warp_sum(out.g_x[k]);
                out.g_y[k] = warp_sum(out.g_y[k]);
            }
        }
        if (lane == 0) {
            if (a.logp) {
#pragma unroll
                for (int m = 0; m < NM; ++m) a.logp[m * a.U + u] = out.logp[m];
            }
            if (BWD) {
                a.g_background[u] = out.g_b;
                a.g_rate[u] = out.g_rate;
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    a.g_height[k * a.U + u] = out.g_h[k];
                    a.g_width[k * a.U + u] = out.g_w[k];
                    a.g_x[k * a.U + u] = out.g_x[k];
                    a.g_y[k * a.U + u] = out.g_y[k];
                }
            }
        }
    }
}


template <typename PIX, int OC, bool P14, bool BWD, int MINB>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB)
ksmogn_stream_kernel(const KsmognArgs<float> a, unsigned int* __restrict__ counters) {
    constexpr bool PF = P14 && sizeof(PIX) == 2;
    constexpr unsigned kFull = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // O > 4 (OC == 0): per-bin constants of the one-pass form (4 floats per bin) in front of the offset tables
    constexpr bool MANY = OC == 0 && P14 && BWD;
    const int O = a.v.O, offpad = ((2 * O + 3) & ~3) + (MANY ? 4 * O : 0);
    BinConst* bins = reinterpret_cast<BinConst*>(smem_raw);
    float* off_s = reinterpret_cast<float*>(smem_raw) + (MANY ? 4 * O : 0);
    float* off_w2 = off_s + O;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = threadIdx.x / kSub, sub = threadIdx.x % kSub, wslot = lane / kSub;
    const int P = P14 ? 14 : a.v.P, PP = P * P;
    float* tab = reinterpret_cast<float*>(smem_raw) + offpad + slot * kTabFloats;
    float* gx = tab;
    float* gy = gx + kK * kMaxP;
    float* spx = reinterpret_cast<float*>(smem_raw) + offpad + kUnitsPerBlock * kTabFloats + slot * PP;
    unsigned char* stage = reinterpret_cast<unsigned char*>(reinterpret_cast<float*>(smem_raw) + offpad + kUnitsPerBlock * (kTabFloats + PP))
                           + slot * stage_bytes<PF>();
    // one mbarrier per patch slot behind the staging areas (bulk-copy staging)
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<unsigned char*>(reinterpret_cast<float*>(smem_raw) + offpad + kUnitsPerBlock * (kTabFloats + PP))
        + kUnitsPerBlock * stage_bytes<PF>()) + slot;
    const bool bulk = PF && a.bulk != 0;
    if (bulk && sub == 0) mbar_init(bar, 1u);
    if (bulk) mbar_fence_init();
    for (int j = threadIdx.x; j < O; j += blockDim.x) {
        off_s[j] = static_cast<const float*>(a.v.offset_samples)[j];
        off_w2[j] = static_cast<const float*>(a.v.offset_logits)[j] * kLog2e;
    }
    // where each of a patch's eight lanes stores its two results (epilogue)
    __shared__ float* out_ptr[kSub][2];
    if (BWD && threadIdx.x < kSub) {
        const int l = threadIdx.x;
        const size_t Uz = (size_t)a.U;
        float* p0 = nullptr;
        float* p1 = nullptr;
        if (l < 2) { if (a.logp) { p0 = a.logp + (size_t)(2 * l) * Uz; p1 = p0 + Uz; } }
        else if (l == 2) { p0 = a.g_background; p1 = a.g_rate; }
        else if (l < 7) { p0 = l == 3 ? a.g_height : l == 4 ? a.g_width : l == 5 ? a.g_x : a.g_y; p1 = p0 + Uz; }
        out_ptr[l][0] = p0;
        out_ptr[l][1] = p1;
    }
    FastConst fc;
    fc.gain = a.gain[0];
    fc.rate = 1.0f / fc.gain;
    fc.rate2 = fc.rate * kLog2e;
    fc.log_rate = logf(fc.rate);
    __syncthreads();
    int ref_bin = 0;
    float delta_ref = 0.0f, w2_ref = 0.0f;
    bool many_ok = false;
    if (MANY) {
        ref_bin = many_bins_reference(O, off_s);
        delta_ref = off_s[ref_bin];
        w2_ref = off_w2[ref_bin];
        // the one-pass form exponentiates v_j - v_ref without a running maximum: it is bounded by c_j = (w2_j - w2_ref) +
        // rate2 (delta_j - delta_ref) (the other term is <= 0).  A histogram so wide, or a gain so small, that some c_j
        // nears the fp32 exponent range (64 bins at gain 7: 13 bits) goes through the two-pass form instead.
        float c_max = 0.0f;
        for (int j = 0; j < O; ++j) c_max = fmaxf(c_max, many_bins_const(j, ref_bin, off_s, off_w2, fc.rate2).c);
        many_ok = c_max < 100.0f;
        for (int j = threadIdx.x; j < O; j += blockDim.x) bins[j] = many_bins_const(j, ref_bin, off_s, off_w2, fc.rate2);
        __syncthreads();
    }

    const PIX* pixels = static_cast<const PIX*>(a.v.pixels);
    const float* xy = static_cast<const float*>(a.v.xy);
    const unsigned U = (unsigned)a.U, n_wg = (U + 3u) >> 2;
    const unsigned stride = gridDim.x * kWarpsPerBlock;

    const bool dense_units = a.v.ndx == nullptr && a.v.fdx == nullptr && a.v.fb == a.v.F;
    // issue the copies of group g: this lane's share of its slot's patch and two of the slot's 15 scalars
    unsigned pix_off = 0u, pix_off_next = 0u, phase = 0u;   // bulk staging: where the patch starts in its slot; mbarrier phase
    auto prefetch = [&](unsigned g) {
        const unsigned u_raw = g * 4u + wslot, u = u_raw < U ? u_raw : U - 1u;
        // whole frames of consecutive AOIs (every full-batch step): the unit index IS the patch index
        UnitIndex ui;
        if (dense_units) ui.patch = (int64_t)u;
        else ui = locate_unit32(u, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
        if (PF && bulk) {
            // 392 B at an 8-byte stride: odd patches start 8 bytes off a 16-byte boundary.  The 384 aligned bytes go by
            // one bulk copy, the other 8 (head of an odd patch, tail of an even one) by one cp.async; the patch lands
            // at offset 0 (even) or 8 (odd) of the slot's pixel area so that the bulk destination is aligned too.
            pix_off_next = (unsigned)(ui.patch & 1) * 8u;
            if (sub == 0) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(pixels + ui.patch * 196);
                unsigned char* dst = stage + kParFloats * 4;
                mbar_expect_tx(bar, 384u);
                if (pix_off_next) {
                    cp_async_8(dst + 8, src);
                    bulk_copy_g2s(dst + 16, src + 8, 384u, bar);
                } else {
                    bulk_copy_g2s(dst, src, 384u, bar);
                    cp_async_8(dst + 384, src + 384);
                }
            }
        } else if (PF) {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(pixels + ui.patch * 196);
#pragma unroll
            for (int t = 0; t < 7; ++t) {
                const int i = sub + t * kSub;
                if (i < 49) cp_async_8(stage + kParFloats * 4 + 8 * i, src + 8 * i);
            }
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const int j = sub + r * kSub;
            const float* src;
            if (j < 8) {
                const float* base = j < 2 ? a.height : j < 4 ? a.width : j < 6 ? a.x : a.y;
                src = base + (size_t)(j & 1) * U + u;
            } else if (j == 8) {
                src = a.background + u;
            } else if (j < 13) {
                src = BWD ? a.W + (size_t)(j - 9) * U + u : nullptr;
            } else {
                src = xy + ui.patch * 2 + (j - 13);
            }
            if (j < 15 && src) cp_async_4(stage + 4 * j, src);
        }
    };

    // first group by position (no dependent atomic at the head of the kernel); later ones from the counter, which
    // counts the groups handed out beyond the first `stride`
    unsigned cur = blockIdx.x * kWarpsPerBlock + warp, pending = 0;
    if (lane == 0) pending = stride + atomicAdd(counters, 1u);
    __syncthreads();   // mbarriers initialised before the first copy can complete on them
    if (cur < n_wg) prefetch(cur);

    while (cur < n_wg) {
        const unsigned u_raw = cur * 4u + wslot;
        const bool live = u_raw < U;
        const unsigned u = live ? u_raw : U - 1u;   // idle slots shadow the last patch and write nothing
        pix_off = pix_off_next;
        cp_async_wait_all();
        if (bulk) {
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        __syncwarp();
        PatchSpots<float> s;
        float W[kM], Wr[kM], tx, ty;
        {
            const float4* par = reinterpret_cast<const float4*>(stage);
            const float4 p0 = par[0], p1 = par[1], p2 = par[2], p3 = par[3];
            s.h[0] = p0.x; s.h[1] = p0.y; s.w[0] = p0.z; s.w[1] = p0.w;
            tx = p3.y; ty = p3.z;
            s.cx[0] = p1.x + tx; s.cx[1] = p1.y + tx; s.cy[0] = p1.z + ty; s.cy[1] = p1.w + ty;
            s.b = p2.x;
            W[0] = BWD ? p2.y : 0.0f; W[1] = BWD ? p2.z : 0.0f; W[2] = BWD ? p2.w : 0.0f; W[3] = BWD ? p3.x : 0.0f;
        }
#pragma unroll
        for (int m = 0; m < kM; ++m) Wr[m] = W[m] * fc.rate;
        float pix_min = 3.0e38f;   // smallest pixel of the patch (this lane's share): selects the no-clamp pair form
        if (PF) {
            const uint2* raw = reinterpret_cast<const uint2*>(stage + kParFloats * 4 + pix_off);
#pragma unroll
            for (int t = 0; t < 7; ++t) {
                const int i = sub + t * kSub;
                if (i < 49) {
                    const uint2 q = raw[i];
                    const float4 v = make_float4(float(q.x & 0xffffu), float(q.x >> 16), float(q.y & 0xffffu), float(q.y >> 16));
                    reinterpret_cast<float4*>(spx)[i] = v;
                    pix_min = fminf(pix_min, fminf(fminf(v.x, v.y), fminf(v.z, v.w)));
                }
            }
        } else {
            const UnitIndex ui = locate_unit32(u, a.v.fb, a.v.C, a.v.F, a.v.ndx, a.v.fdx);
            const PIX* src = pixels + ui.patch * PP;
            for (int p = sub; p < PP; p += kSub) {
                const float pv = float(src[p]);
                spx[p] = pv;
                pix_min = fminf(pix_min, pv);
            }
        }
        __syncwarp();   // the staging area has been read by every lane: the next group may land in it
        const unsigned nxt = __shfl_sync(kFull, pending, 0);   // asked for one group ago: the atomic's latency is hidden
        if (lane == 0) pending = stride + atomicAdd(counters, 1u);
        if (nxt < n_wg) prefetch(nxt);

        // a = image/gain is smallest without spots: one test per patch selects the Stirling variant
        const bool small = s.b * fc.rate < 4.0f;
        bool pairs = false;
        if (P14 && BWD && OC > 0) {
            // every pixel above every offset (no -inf handling needed) and no small concentration anywhere in the warp
            float max_off = off_s[0];
            for (int j = 1; j < OC; ++j) max_off = fmaxf(max_off, off_s[j]);
            pairs = __all_sync(kFull, !small && pix_min > max_off);
        }
        // many bins: every pixel above the SMALLEST offset (bins at or above a pixel drop out by underflow)
        if (MANY) pairs = many_ok && __all_sync(kFull, !small && pix_min > delta_ref);
        // separable spot factors: 2*K*P exponentials per patch instead of K*P*P (ksmogn_sweep.cuh: table layouts)
        float norm[kK];
        if (pairs) build_tables_pairs(tab, sub, s);
        else build_tables_generic(tab, P, sub, s, norm);
        __syncwarp();

        PatchOut<float, kM> out;
        out.zero();
        if (MANY && pairs)
            sweep_patch_pairs_many(spx, sub, tab, s, fc, O, bins, delta_ref, w2_ref, W, out);
        else if (pairs && OC == 1)
            sweep_patch_rows_single_bin(spx, sub, tab, s, fc, off_s[0], off_w2[0] * kLn2, W, out);
        else if (pairs)
            sweep_patch_pairs<OC>(spx, sub, tab, s, fc, off_s, off_w2, W, out);
        else if (__any_sync(kFull, small))
            sweep_patch<OC, P14, BWD, true>(spx, P, PP, sub, gx, gy, s, norm, fc, O, off_s, off_w2, W, Wr, out);
        else
            sweep_patch<OC, P14, BWD, false>(spx, P, PP, sub, gx, gy, s, norm, fc, O, off_s, off_w2, W, Wr, out);

        if (BWD) {
            // the 14 sums of a patch, reduce-scattered: lane 0 of its eight ends with logp[0], logp[1]; 1: logp[2], logp[3];
            // 2: d/d background, d/d (1/gain); 3: the height moments A0; 4: A3; 5: A1; 6: A2 (k = 0, 1 each); 7: nothing
            const float v[16] = {out.logp[0], out.logp[1], out.logp[2], out.logp[3], out.g_b, out.g_rate, out.g_h[0], out.g_h[1],
                                 out.g_w[0], out.g_w[1], out.g_x[0], out.g_x[1], out.g_y[0], out.g_y[1], 0.0f, 0.0f};
            float y[kK];
            sub_reduce_scatter16(v, sub, y[0], y[1]);
            // moments -> gradients (finish_spot_moments, the same arithmetic on every lane's own pair; d/dw also needs A0)
            PatchOut<float, kM> fin;
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                fin.g_h[k] = __shfl_sync(kFull, y[k], (lane & ~7) | 3);
                fin.g_w[k] = fin.g_x[k] = fin.g_y[k] = y[k];
            }
            finish_spot_moments(s, fin);
#pragma unroll
            for (int k = 0; k < kK; ++k) y[k] = sub == 3 ? fin.g_h[k] : sub == 4 ? fin.g_w[k] : sub >= 5 ? fin.g_x[k] : y[k];
            if (live) {
                float* const p0 = out_ptr[sub][0];
                float* const p1 = out_ptr[sub][1];
                if (p0) p0[u] = y[0];
                if (p1) p1[u] = y[1];
            }
        } else {
#pragma unroll
            for (int m = 0; m < kM; ++m) out.logp[m] = sub_sum(out.logp[m]);
            if (live && sub == 0 && a.logp) {
#pragma unroll
                for (int m = 0; m < kM; ++m) a.logp[(size_t)m * U + u] = out.logp[m];
            }
        }
        cur = nxt;
    }
    if (lane == 0) {
        // the last warp out re-arms the counters for the next launch on this stream
        __threadfence();
        const unsigned done = atomicAdd(counters + 1, 1u);
        if (done == stride - 1u) {
            counters[0] = 0u;
            counters[1] = 0u;
        }
    }
}

// Work counters of the kernel above: {groups handed out beyond the first per warp, warps finished}, re-armed by the
// kernel itself.  One set per stream (launches on a stream are ordered, so they can share; launches on different streams
// may overlap and must not).
constexpr int kCounterSets = 64;
__device__ unsigned int g_stream_counters[kCounterSets][2];

static int counter_set_of(cudaStream_t st) {
    static std::mutex mu;
    static std::unordered_map<cudaStream_t, int> sets;
    std::lock_guard<std::mutex> lock(mu);
    auto it = sets.find(st);
    if (it != sets.end()) return it->second;
    const int k = (int)(sets.size() % kCounterSets);   // more than 64 live streams: sets are shared again
    sets.emplace(st, k);
    return k;
}

template <typename PIX, int OC, bool P14, bool BWD>
static int launch_fast(const KsmognArgs<float>& a, cudaStream_t st) {
    constexpr bool PF = P14 && sizeof(PIX) == 2;
#ifndef TQ_KS_MINB
#define TQ_KS_MINB 4
#endif
    constexpr int MINB = TQ_KS_MINB;   // 4 resident blocks per SM (<= 128 registers): 5 / 6 blocks measured slower (spills)
    if (a.U > 0x7fffffff) {
        set_error("ksmogn: %lld patches in one launch (limit 2^31 - 1)", (long long)a.U);
        return TQ_ERR_ARG;
    }
    const int PP = a.v.P * a.v.P, offpad = ((2 * a.v.O + 3) & ~3) + (OC == 0 && P14 && BWD ? 4 * a.v.O : 0);
    const size_t smem = sizeof(float) * ((size_t)offpad + kUnitsPerBlock * (kTabFloats + (size_t)PP)) +
                        (size_t)kUnitsPerBlock * stage_bytes<PF>() + (size_t)kUnitsPerBlock * sizeof(unsigned long long);
    auto kern = ksmogn_stream_kernel<PIX, OC, P14, BWD, MINB>;
    if (smem > 48 * 1024) {
        int st2 = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                              "cudaFuncSetAttribute(ksmogn_stream)");
        if (st2 != TQ_OK) return st2;
    }
    void* base = nullptr;
    int st2 = cuda_status(cudaGetSymbolAddress(&base, g_stream_counters), "cudaGetSymbolAddress(g_stream_counters)");
    if (st2 != TQ_OK) return st2;
    unsigned int* counters = static_cast<unsigned int*>(base) + 2 * counter_set_of(st);
    // one resident wave of persistent warps; groups beyond each warp's first are handed out dynamically, so SMs that
    // run slower (a neighbour kernel on the side stream, clock differences) simply take fewer
    const int64_t n_wg = (a.U + 3) / 4, blocks_needed = (n_wg + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t cap = (int64_t)sm_count() * MINB;
    const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
    kern<<<grid, kWarpsPerBlock * 32, smem, st>>>(a, counters);
    TQ_LAUNCH_CHECK("ksmogn_stream_kernel launch");
    return TQ_OK;
}

template <typename PIX, bool BWD>
static int dispatch_fast(const KsmognArgs<float>& a, cudaStream_t st) {
    const bool p14 = a.v.P == 14;
    switch (a.v.O <= 4 ? a.v.O : 0) {
        case 1: return p14 ? launch_fast<PIX, 1, true, BWD>(a, st) : launch_fast<PIX, 1, false, BWD>(a, st);
        case 2: return p14 ? launch_fast<PIX, 2, true, BWD>(a, st) : launch_fast<PIX, 2, false, BWD>(a, st);
        case 3: return p14 ? launch_fast<PIX, 3, true, BWD>(a, st) : launch_fast<PIX, 3, false, BWD>(a, st);
        case 4: return p14 ? launch_fast<PIX, 4, true, BWD>(a, st) : launch_fast<PIX, 4, false, BWD>(a, st);
        default: return p14 ? launch_fast<PIX, 0, true, BWD>(a, st) : launch_fast<PIX, 0, false, BWD>(a, st);
    }
}

// float + enumerated table (mcfg == NULL): production kernel; everything else: exact generic kernel
template <bool BWD>
static int try_fast(const KsmognArgs<float>& a, int NM, cudaStream_t st, bool& handled) {
    handled = false;
    if (NM != kM || a.mcfg != nullptr || a.U == 0) return TQ_OK;
    handled = true;
    switch (a.v.pixtype) {
        case TQ_PIX_U16: return dispatch_fast<uint16_t, BWD>(a, st);
        case TQ_PIX_F32: return dispatch_fast<float, BWD>(a, st);
        default: handled = false; return TQ_OK;
    }
}
template <bool BWD> static int try_fast(const KsmognArgs<double>&, int, cudaStream_t, bool& handled) { handled = false; return TQ_OK; }

template <typename T, typename PIX, int NM, bool BWD>
static int launch_ksmogn(const KsmognArgs<T>& a, cudaStream_t st) {
    if (a.U == 0) return TQ_OK;
    const size_t smem = sizeof(T) * (2 * (size_t)a.v.O + kWarpsPerBlock * 2 * kK * kMaxP);
    auto kern = ksmogn_kernel<T, PIX, NM, BWD>;
    if (smem > 48 * 1024) {
        int st2 = cuda_status(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                              "cudaFuncSetAttribute(ksmogn)");
        if (st2 != TQ_OK) return st2;
    }
    const int64_t blocks_needed = (a.U + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const int64_t cap = (int64_t)sm_count() * 16;  // 16 x 4 warps = full residency per SM
    const int grid = (int)(blocks_needed < cap ? blocks_needed : cap);
    kern<<<grid, kWarpsPerBlock * 32, smem, st>>>(a);
    TQ_LAUNCH_CHECK("ksmogn_kernel launch");
    return TQ_OK;
}

template <typename T, int NM, bool BWD>
static int dispatch_pix(const KsmognArgs<T>& a, cudaStream_t st) {
    switch (a.v.pixtype) {
        case TQ_PIX_U16: return launch_ksmogn<T, uint16_t, NM, BWD>(a, st);
        case TQ_PIX_F32: return launch_ksmogn<T, float, NM, BWD>(a, st);
        case TQ_PIX_F64: return launch_ksmogn<T, double, NM, BWD>(a, st);
    }
    set_error("ksmogn: unknown pixtype %d", a.v.pixtype);
    return TQ_ERR_ARG;
}

template <typename T, bool BWD>
static int run_ksmogn(const tq_patch_view* view, const void* height, const void* width, const void* x,
                      const void* y, const void* background, const void* gain, const void* mcfg, int NM,
                      const void* W, void* logp, void* g_height, void* g_width, void* g_x, void* g_y,
                      void* g_background, void* g_rate, void* stream) {
    KsmognArgs<T> a;
    a.v = *view;
    a.height = (const T*)height; a.width = (const T*)width; a.x = (const T*)x; a.y = (const T*)y;
    a.background = (const T*)background; a.gain = (const T*)gain; a.mcfg = (const T*)mcfg;
    a.W = (const T*)W; a.logp = (T*)logp;
    a.g_height = (T*)g_height; a.g_width = (T*)g_width; a.g_x = (T*)g_x; a.g_y = (T*)g_y;
    a.g_background = (T*)g_background; a.g_rate = (T*)g_rate;
    a.U = (int64_t)view->nb * view->fb * view->C;
    // TQ_BULK=1: stage the pixels with 1-D bulk copies (TMA) instead of seven 8-byte cp.async per lane.  Correct
    // (tests/test_ksmogn_gpu.py) and measured 1 % SLOWER on a B200 (C3 kernel 3.85 vs 3.81 ms, profiles/r2_bulk_ab.sh): a
    // 392-byte patch at an 8-byte stride needs a bulk copy plus an 8-byte cp.async anyway, and the mbarrier wait costs more
    // than cp.async.wait_all on copies that have long landed -- so cp.async stays the default.
    const char* bulk_env = getenv("TQ_BULK");
    a.bulk = ((uintptr_t)view->pixels % 16 == 0 && bulk_env && bulk_env[0] == '1') ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    bool handled = false;
    const int fst = try_fast<BWD>(a, NM, st, handled);
    if (handled) return fst;
    if (NM == 1) return dispatch_pix<T, 1, BWD>(a, st);
    if (NM == kM) return dispatch_pix<T, kM, BWD>(a, st);
    set_error("ksmogn: NM must be 1 or %d, got %d", kM, NM);
    return TQ_ERR_UNSUPPORTED;
}

static int check_view(const tq_patch_view* v) {
    TQ_CHECK_ARG(v != nullptr, "view is NULL");
    TQ_CHECK_ARG(v->nb >= 0 && v->fb >= 0 && v->C >= 1, "bad minibatch shape");
    TQ_CHECK_ARG(v->P >= 2 && v->P <= kMaxP, "P must be in [2, 32]");
    TQ_CHECK_ARG(v->O >= 1, "need at least one offset bin");
    TQ_CHECK_ARG(v->pixels && v->xy && v->offset_samples && v->offset_logits, "NULL dataset pointer");
    return TQ_OK;
}

// ---- gaussian_spots -----------------------------------------------------------------------------
template <typename T>
__global__ void gaussian_spots_kernel(int64_t U, int K, int P, const T* __restrict__ height,
                                      const T* __restrict__ width, const T* __restrict__ x,
                                      const T* __restrict__ y, const T* __restrict__ target,
                                      const T* __restrict__ m, T* __restrict__ out) {
    const int64_t total = U * K * P * P;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % P);
        const int row = (int)((i / P) % P);
        const int k = (int)((i / ((int64_t)P * P)) % K);
        const int64_t u = i / ((int64_t)P * P * K);
        const T w = width[k * U + u];
        T h = height[k * U + u];
        if (m) h *= m[k * U + u];
        const T cx = x[k * U + u] + target[u * 2 + 0];
        const T cy = y[k * U + u] + target[u * 2 + 1];
        out[i] = h * axis_factor<T>(col, cx, w) * axis_factor<T>(row, cy, w)
                 / (T(6.283185307179586476925) * w * w);
    }
}

}  // namespace tq

using namespace tq;

extern "C" int tq_gaussian_spots(int dtype, int64_t U, int K, int P, const void* height, const void* width,
                                 const void* x, const void* y, const void* target_xy, const void* m,
                                 void* out, void* stream) {
    TQ_CHECK_ARG(U >= 0 && P >= 1 && K >= 1, "bad shape");
    if (U == 0) return TQ_OK;
    TQ_CHECK_ARG(height && width && x && y && target_xy && out, "NULL pointer");
    const int64_t total = U * K * P * P;
    const int block = 256;
    int64_t grid = (total + block - 1) / block;
    const int64_t cap = (int64_t)sm_count() * 32;
    if (grid > cap) grid = cap;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == TQ_F32)
        gaussian_spots_kernel<float><<<(int)grid, block, 0, st>>>(U, K, P, (const float*)height, (const float*)width,
            (const float*)x, (const float*)y, (const float*)target_xy, (const float*)m, (float*)out);
    else if (dtype == TQ_F64)
        gaussian_spots_kernel<double><<<(int)grid, block, 0, st>>>(U, K, P, (const double*)height, (const double*)width,
            (const double*)x, (const double*)y, (const double*)target_xy, (const double*)m, (double*)out);
    else { set_error("tq_gaussian_spots: bad dtype %d", dtype); return TQ_ERR_ARG; }
    TQ_LAUNCH_CHECK("gaussian_spots_kernel launch");
    return TQ_OK;
}

extern "C" int tq_ksmogn_fwd(int dtype, const tq_patch_view* view, const void* height, const void* width,
                             const void* x, const void* y, const void* background, const void* gain,
                             const void* mcfg, int NM, void* logp, void* stream) {
    int st = check_view(view);
    if (st != TQ_OK) return st;
    TQ_CHECK_ARG(height && width && x && y && background && gain && logp, "NULL pointer");
    TQ_CHECK_ARG(mcfg || NM == kM, "mcfg == NULL selects the enumerated 2^K table and needs NM == 4");
    if (dtype == TQ_F32)
        return run_ksmogn<float, false>(view, height, width, x, y, background, gain, mcfg, NM, nullptr, logp,
                                        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
    if (dtype == TQ_F64)
        return run_ksmogn<double, false>(view, height, width, x, y, background, gain, mcfg, NM, nullptr, logp,
                                         nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
    set_error("tq_ksmogn_fwd: bad dtype %d", dtype);
    return TQ_ERR_ARG;
}

extern "C" int tq_ksmogn_fwd_bwd(int dtype, const tq_patch_view* view, const void* height, const void* width,
                                 const void* x, const void* y, const void* background, const void* gain,
                                 const void* mcfg, int NM, const void* W, void* logp, void* g_height,
                                 void* g_width, void* g_x, void* g_y, void* g_background, void* g_rate,
                                 void* stream) {
    int st = check_view(view);
    if (st != TQ_OK) return st;
    TQ_CHECK_ARG(height && width && x && y && background && gain && W, "NULL input pointer");
    TQ_CHECK_ARG(mcfg || NM == kM, "mcfg == NULL selects the enumerated 2^K table and needs NM == 4");
    TQ_CHECK_ARG(g_height && g_width && g_x && g_y && g_background && g_rate, "NULL gradient pointer");
    if (dtype == TQ_F32)
        return run_ksmogn<float, true>(view, height, width, x, y, background, gain, mcfg, NM, W, logp,
                                       g_height, g_width, g_x, g_y, g_background, g_rate, stream);
    if (dtype == TQ_F64)
        return run_ksmogn<double, true>(view, height, width, x, y, background, gain, mcfg, NM, W, logp,
                                        g_height, g_width, g_x, g_y, g_background, g_rate, stream);
    set_error("tq_ksmogn_fwd_bwd: bad dtype %d", dtype);
    return TQ_ERR_ARG;
}
