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

// P = 14 common case: lane `sub` takes the pixel pairs (row, col), (row + 7, col) with row * 14 + col = sub + 8 t,
// t = 0..12 (98 pairs over 8 lanes), in packed two-wide FP32 (ksmogn_fast.cuh)
template <int OC>
__device__ __forceinline__ void sweep_patch_pairs(const float* __restrict__ pix, int sub, const float* gx, const float* gy,
                                                  const PatchSpots<float>& s, const float (&norm)[kK], const FastConst& fc,
                                                  const float* off_s_sm, const float* off_w2_sm, const float (&W)[kM],
                                                  PatchOut<float, kM>& out) {
    constexpr int NC = OC > 0 ? OC : 1;
    float off_s[NC], off_w2[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) { off_s[j] = off_s_sm[j]; off_w2[j] = off_w2_sm[j]; }
    int col = sub, row = 0;
    if (OC == 1) {
        const SingleBinConst sc = single_bin_const(s.b, fc);
        PairOut1 po;
        po.zero();
        int npix = 0;
#pragma unroll 1
        for (int t = 0; t < 13; ++t) {
            if (row < 7) {
                float gxn[kK], dx[kK];
                F2 gyk[kK], dy[kK];
                const float fr = float(row);
#pragma unroll
                for (int k = 0; k < kK; ++k) {
                    gxn[k] = gx[k * kMaxP + col] * norm[k];
                    gyk[k] = F2{gy[k * kMaxP + row], gy[k * kMaxP + row + 7]};
                    dx[k] = float(col) - s.cx[k];
                    dy[k] = F2{fr - s.cy[k], (fr + 7.0f) - s.cy[k]};
                }
                const F2 D{pix[row * 14 + col], pix[(row + 7) * 14 + col]};
                pixel_pair_single_bin(D, gxn, gyk, dx, dy, s, fc, sc, off_s[0], W, po);
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
            float gxn[kK], dx[kK];
            F2 gyk[kK], dy[kK];
            const float fr = float(row);
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                gxn[k] = gx[k * kMaxP + col] * norm[k];
                gyk[k] = F2{gy[k * kMaxP + row], gy[k * kMaxP + row + 7]};
                dx[k] = float(col) - s.cx[k];
                dy[k] = F2{fr - s.cy[k], (fr + 7.0f) - s.cy[k]};
            }
            const F2 D{pix[row * 14 + col], pix[(row + 7) * 14 + col]};
            pixel_pair_accumulate_fast<NC>(D, gxn, gyk, dx, dy, s, fc, off_s, off_w2, W, po);
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

constexpr int kParFloats = 16;   // h0 h1 w0 w1 x0 x1 y0 y1 | b W0 W1 W2 | W3 tx ty -
template <bool PF> constexpr int stage_bytes() { return kParFloats * 4 + (PF ? 400 : 0); }   // 392 B of pixels, 16 B aligned slots

}  // namespace tq
