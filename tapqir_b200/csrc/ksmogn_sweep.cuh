// Device pieces of the fp32 likelihood sweep shared by the stand-alone kernel (ksmogn.cu::ksmogn_stream_kernel) and the
// fused step kernel (cosmos_fused.cu): the 8-lanes-per-patch sweeps, their reductions and the cp.async staging helpers.
#pragma once
#include "ksmogn_core.cuh"
#include "ksmogn_fast.cuh"

namespace tq {

constexpr int kWarpsPerBlock = 4;

// ---- fp32 production kernel (ksmogn_fast.cuh) -------------------------------------------------------------
// Eight lanes per patch, four patches per warp: 196 pixels / 8 lanes = 24.5 -> 25 sweeps (98 % lane
// use, against 87.5 % for one warp per patch), the 14 per-patch sums reduce over 3 shuffle levels
// instead of 5, and the per-patch scalars live in the registers of the 8 lanes that use them.
constexpr int kSub = 8;                                   // lanes per patch
constexpr int kUnitsPerBlock = kWarpsPerBlock * 32 / kSub;

template <typename T> __device__ __forceinline__ T sub_sum(T v) {
#pragma unroll
    for (int o = kSub / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce-scatter of 16 values over the eight lanes of a patch: lane l ends with the sums of v[2 l] and v[2 l + 1].
// Same partners (4, 2, 1) and the same pairing of partial sums as the butterfly of sub_sum, so every sum has exactly the bits
// sub_sum gives it -- at 56 instructions instead of 126 for the 14 sums of a patch, and with the results already spread
// over the lanes that store them.
__device__ __forceinline__ void sub_reduce_scatter16(const float (&v)[16], int sub, float& y0, float& y1) {
    const bool b2 = (sub & 4) != 0, b1 = (sub & 2) != 0, b0 = (sub & 1) != 0;
    float w[8], x[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float keep = b2 ? v[i + 8] : v[i], send = b2 ? v[i] : v[i + 8];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float keep = b1 ? w[i + 4] : w[i], send = b1 ? w[i] : w[i + 4];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    {
        const float keep0 = b0 ? x[2] : x[0], send0 = b0 ? x[0] : x[2];
        const float keep1 = b0 ? x[3] : x[1], send1 = b0 ? x[1] : x[3];
        y0 = keep0 + __shfl_xor_sync(0xffffffffu, send0, 1);
        y1 = keep1 + __shfl_xor_sync(0xffffffffu, send1, 1);
    }
}

template <int OC, bool P14, bool BWD, bool SMALL>
__device__ __forceinline__ void sweep_patch(const float* __restrict__ pix, int P, int PP, int sub, const float* gx,
                                            const float* gy, const PatchSpots<float>& s, const float (&norm)[kK],
                                            const FastConst& fc, int O, const float* off_s,
                                            const float* off_w2, const float (&W)[kM], const float (&Wr)[kM],
                                            PatchOut<float, kM>& out) {
#pragma unroll 1
    for (int p = sub; p < PP; p += kSub) {
        const int row = P14 ? p / 14 : p / P, col = p - row * (P14 ? 14 : P);
        float gxk[kK], gyk[kK];
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            gxk[k] = gx[k * kMaxP + col];
            gyk[k] = gy[k * kMaxP + row];
        }
        pixel_accumulate_fast<kM, OC, BWD, SMALL>(pix[p], gxk, gyk, col, row, s, norm, fc, O, off_s, off_w2,
                                                  W, Wr, out);
    }
}

// ---- per-patch tables of the separable render (2 K P exponentials per patch instead of K P P) ---------------------------
// One table area of kTabFloats floats per patch slot, in one of two layouts:
//   generic   gx[k * kMaxP + i], gy[k * kMaxP + i] = exp(-(i - c)^2 / (2 w^2))             (any P; sweep_patch)
//   pairs     P = 14, packed sweep: everything a pixel pair needs that depends on its column or on its row pair alone,
//             so that the sweep's inner loop is left with the arithmetic that depends on BOTH (and with wide loads):
//             col[c] (8 floats): gxh_0, gxh_1 (column factor x norm x height), dx_0, dx_1, dx_0^2, dx_1^2, -, -
//             row[r] (8 floats), r = 0..6: gy_0(r), gy_0(r+7), gy_1(r), gy_1(r+7), dy_0(r), dy_0(r+7), dy_1(r), dy_1(r+7)
constexpr int kTabFloats = 176;   // >= 2 kK kMaxP (generic) and >= 14 * 8 + 7 * 8 (pairs), 16-byte multiple
static_assert(kTabFloats >= 2 * kK * kMaxP && kTabFloats >= 21 * 8 && kTabFloats % 4 == 0, "table area too small");

__device__ __forceinline__ void build_tables_generic(float* tab, int P, int sub, const PatchSpots<float>& s, float (&norm)[kK]) {
    float* gx = tab;
    float* gy = tab + kK * kMaxP;
    float c2[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const float iw = rcp_newton(s.w[k]);
        norm[k] = 0.15915494309189535f * iw * iw;
        c2[k] = (-0.5f * kLog2e) * iw * iw;
    }
    for (int i = sub; i < P; i += kSub) {
        const float fi = float(i);
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const float dx = fi - s.cx[k], dy = fi - s.cy[k];
            gx[k * kMaxP + i] = f_ex2(c2[k] * dx * dx);
            gy[k * kMaxP + i] = f_ex2(c2[k] * dy * dy);
        }
    }
}

__device__ __forceinline__ void build_tables_pairs(float* tab, int sub, const PatchSpots<float>& s) {
    float c2[kK], nh[kK];
#pragma unroll
    for (int k = 0; k < kK; ++k) {
        const float iw = rcp_newton(s.w[k]);
        nh[k] = 0.15915494309189535f * iw * iw * s.h[k];
        c2[k] = (-0.5f * kLog2e) * iw * iw;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int c = sub + j * kSub;
        if (c < 14) {
            const float fc_ = float(c);
            const float dx0 = fc_ - s.cx[0], dx1 = fc_ - s.cx[1];
            const float q0 = dx0 * dx0, q1 = dx1 * dx1;
            float4* dst = reinterpret_cast<float4*>(tab + c * 8);
            dst[0] = make_float4(f_ex2(c2[0] * q0) * nh[0], f_ex2(c2[1] * q1) * nh[1], dx0, dx1);
            dst[1] = make_float4(q0, q1, 0.0f, 0.0f);
        }
    }
    if (sub < 7) {
        const float fr = float(sub);
        const float a0 = fr - s.cy[0], b0 = (fr + 7.0f) - s.cy[0], a1 = fr - s.cy[1], b1 = (fr + 7.0f) - s.cy[1];
        float4* dst = reinterpret_cast<float4*>(tab + 14 * 8 + sub * 8);
        dst[0] = make_float4(f_ex2(c2[0] * a0 * a0), f_ex2(c2[0] * b0 * b0), f_ex2(c2[1] * a1 * a1), f_ex2(c2[1] * b1 * b1));
        dst[1] = make_float4(a0, b0, a1, b1);
    }
}

// P = 14 common case: lane `sub` takes the pixel pairs (row, col), (row + 7, col) with row * 14 + col = sub + 8 t,
// t = 0..12 (98 pairs over 8 lanes), in packed two-wide FP32 (ksmogn_fast.cuh); `tab`: build_tables_pairs
template <int OC>
__device__ __forceinline__ void sweep_patch_pairs(const float* __restrict__ pix, int sub, const float* __restrict__ tab,
                                                  const PatchSpots<float>& s, const FastConst& fc,
                                                  const float* off_s_sm, const float* off_w2_sm, const float (&W)[kM],
                                                  PatchOut<float, kM>& out) {
    constexpr int NC = OC > 0 ? OC : 1;
    float off_s[NC], off_w2[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) { off_s[j] = off_s_sm[j]; off_w2[j] = off_w2_sm[j]; }
    const float4* col4 = reinterpret_cast<const float4*>(tab);
    const float4* row4 = reinterpret_cast<const float4*>(tab + 14 * 8);
    int col = sub, row = 0;
    auto load = [&](float (&gxh)[kK], float (&dx)[kK], float (&dx2)[kK], F2 (&gyk)[kK], F2 (&dy)[kK]) {
        const float4 c0 = col4[col * 2], c1 = col4[col * 2 + 1], r0 = row4[row * 2], r1 = row4[row * 2 + 1];
        gxh[0] = c0.x; gxh[1] = c0.y; dx[0] = c0.z; dx[1] = c0.w; dx2[0] = c1.x; dx2[1] = c1.y;
        gyk[0] = F2{r0.x, r0.y}; gyk[1] = F2{r0.z, r0.w}; dy[0] = F2{r1.x, r1.y}; dy[1] = F2{r1.z, r1.w};
    };
    if (OC == 1) {
        const SingleBinConst sc = single_bin_const(s.b, fc);
        PairOut1 po;
        po.zero();
        int npix = 0;
#pragma unroll 1
        for (int t = 0; t < 13; ++t) {
            if (row < 7) {
                float gxh[kK], dx[kK], dx2[kK];
                F2 gyk[kK], dy[kK];
                load(gxh, dx, dx2, gyk, dy);
                const F2 D{pix[row * 14 + col], pix[(row + 7) * 14 + col]};
                pixel_pair_single_bin(D, gxh, gyk, dx, dx2, dy, s, fc, sc, off_s[0], W, po);
                npix += 2;
            }
            col += 8;
            if (col >= 14) { col -= 14; ++row; }
        }
        finish_single_bin(po, sc, fc, s.b, off_w2[0] * kLn2, W[0], npix, out);
        return;
    }
    PairOut po;
    po.zero();
#pragma unroll 1
    for (int t = 0; t < 13; ++t) {
        if (row < 7) {
            float gxh[kK], dx[kK], dx2[kK];
            F2 gyk[kK], dy[kK];
            load(gxh, dx, dx2, gyk, dy);
            const F2 D{pix[row * 14 + col], pix[(row + 7) * 14 + col]};
            pixel_pair_accumulate_fast<NC>(D, gxh, gyk, dx, dx2, dy, s, fc, off_s, off_w2, W, po);
        }
        col += 8;
        if (col >= 14) { col -= 14; ++row; }
    }
    finish_pair(po, fc.rate, out);
}

// Single offset bin, ROW mapping: lane r < 7 of the patch's eight takes the row pair (r, r + 7) and walks the 14 columns.
// Row factors and distances stay in registers for the whole patch, the column index is a compile-time constant (tables
// and pixels at immediate offsets: no index arithmetic, no branch in the loop), and the y / dy^2 moments follow from the
// height moment at the end of the row (finish_row_single_bin).  14 trips of 60 packed + ~20 other instructions (round 2:
// 75 + 21 before the sweep moved to units of a = image / gain, ksmogn_fast.cuh "ROW form") against 13 trips of 79 + 41 in the
// (row, col) = (p / 14, p % 14) walk; the eighth lane idles through the sweep.
#ifndef TQ_ROW_UNROLL
#define TQ_ROW_UNROLL 7   // two trips of seven: the fully unrolled loop (21 KB) stalled on instruction fetch (C3 kernel 3.64 vs 3.58 ms)
#endif
constexpr int kRowUnroll = TQ_ROW_UNROLL;
__device__ __forceinline__ void sweep_patch_rows_single_bin(const float* __restrict__ pix, int sub, const float* __restrict__ tab,
                                                            const PatchSpots<float>& s, const FastConst& fc, float off,
                                                            float log_w, const float (&W)[kM], PatchOut<float, kM>& out) {
    const SingleBinConst sc = single_bin_const(s.b, fc);
    const RowConst rc = row_const(off, s.b, fc);
    RowOut po;
    po.zero();
    int npix = 0;
    F2 gyr[kK] = {f2(0.0f), f2(0.0f)}, dy[kK] = {f2(0.0f), f2(0.0f)};
    if (sub < 7) {
        const float4* col4 = reinterpret_cast<const float4*>(tab);
        const float4* row4 = reinterpret_cast<const float4*>(tab + 14 * 8);
        const float4 r0 = row4[sub * 2], r1 = row4[sub * 2 + 1];
        gyr[0] = mul2(F2{r0.x, r0.y}, f2(fc.rate)); gyr[1] = mul2(F2{r0.z, r0.w}, f2(fc.rate));
        dy[0] = F2{r1.x, r1.y}; dy[1] = F2{r1.z, r1.w};
        const float* p0 = pix + sub * 14;
#pragma unroll kRowUnroll
        for (int c = 0; c < 14; ++c) {
            const float4 c0 = col4[c * 2];
            const float2 c1 = *reinterpret_cast<const float2*>(col4 + c * 2 + 1);
            const float gxh[kK] = {c0.x, c0.y}, dx[kK] = {c0.z, c0.w}, dx2[kK] = {c1.x, c1.y};
            const F2 D{p0[c], p0[c + 98]};
            row_pair_single_bin(D, gxh, gyr, dx, dx2, sc, rc, W, po);
        }
        npix = 28;
    }
    finish_row_single_bin(po, gyr, dy, sc, fc, log_w, W, npix, out);
}

// the same pair mapping for O > 4 offset bins (ksmogn_fast.cuh: "many offset bins"); bins: per-bin constants in shared memory
__device__ __forceinline__ void sweep_patch_pairs_many(const float* __restrict__ pix, int sub, const float* __restrict__ tab,
                                                       const PatchSpots<float>& s, const FastConst& fc, int O,
                                                       const BinConst* __restrict__ bins, float delta_ref, float w2_ref,
                                                       const float (&W)[kM], PatchOut<float, kM>& out) {
    const float4* col4 = reinterpret_cast<const float4*>(tab);
    const float4* row4 = reinterpret_cast<const float4*>(tab + 14 * 8);
    int col = sub, row = 0;
    PairOut po;
    po.zero();
#pragma unroll 1
    for (int t = 0; t < 13; ++t) {
        if (row < 7) {
            const float4 c0 = col4[col * 2], c1 = col4[col * 2 + 1], r0 = row4[row * 2], r1 = row4[row * 2 + 1];
            const float gxh[kK] = {c0.x, c0.y}, dx[kK] = {c0.z, c0.w}, dx2[kK] = {c1.x, c1.y};
            const F2 gyk[kK] = {F2{r0.x, r0.y}, F2{r0.z, r0.w}}, dy[kK] = {F2{r1.x, r1.y}, F2{r1.z, r1.w}};
            const F2 D{pix[row * 14 + col], pix[(row + 7) * 14 + col]};
            pixel_pair_many_bins(D, gxh, gyk, dx, dx2, dy, s, fc, O, bins, delta_ref, w2_ref, W, po);
        }
        col += 8;
        if (col >= 14) { col -= 14; ++row; }
    }
    finish_pair(po, fc.rate, out);
}

// ---- streaming form of the production kernel -----------------------------------------------------------------
// Persistent warps: every warp walks its own sequence of 4-patch groups and, while it sweeps one group, the next
// group's pixels (49 x 8 B per patch) and its 15 per-patch scalars are already in flight into a per-slot staging
// area (cp.async), so the HBM/L2 latency of a patch is hidden behind the previous patch's arithmetic instead of
// stalling the warp at the head of every block (19 % of warp time in the one-block-per-16-patches form).
__device__ __forceinline__ void cp_async_4(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- 1-D bulk copies (TMA, cp.async.bulk -> SASS UBLKCP) completing on an mbarrier ------------------------------------------
// One elected lane per patch slot moves the patch's pixels with ONE instruction instead of seven 8-byte cp.async per lane
// of the slot.  A bulk copy needs 16-byte aligned addresses and sizes; a 14 x 14 uint16 patch is 392 B at an 8-byte
// stride, so the copy takes the 384 bytes of the patch that ARE 16-byte aligned and one 8-byte cp.async the remaining
// head (odd patches) or tail (even patches) -- exact bytes, nothing read beyond the patch.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(void* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0u;
}
// wait for phase `parity`; a copy that never lands (a bug, not a data condition) traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    for (unsigned spins = 0; !mbar_try_wait(bar, parity); ++spins)
        if (spins > (1u << 24)) __trap();
}

constexpr int kParFloats = 16;   // h0 h1 w0 w1 x0 x1 y0 y1 | b W0 W1 W2 | W3 tx ty -
template <bool PF> constexpr int stage_bytes() { return kParFloats * 4 + (PF ? 400 : 0); }   // 392 B of pixels, 16 B aligned slots

}  // namespace tq
