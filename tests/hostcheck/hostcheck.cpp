// CPU harness around the host+device arithmetic headers of tapqir_b200/csrc (TEST INFRASTRUCTURE).
// Compiled with g++ by tests/hostcheck/__init__.py; lets `-m "not gpu"` tests compare the exact
// per-pixel / per-unit formulas the kernels execute against the oracle without a GPU.  Nothing
// in the product imports or links this file.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "ksmogn_core.cuh"
#include "ksmogn_fast.cuh"

using namespace tq;

template <typename T>
static void ksmogn_host(int64_t U, int P, int O, int NM, const T* height, const T* width, const T* x, const T* y,
                        const T* background, T gain, const T* mcfg, const T* W, const T* value, const T* target,
                        const T* off_s, const T* off_w, T* logp, T* g_h, T* g_w, T* g_x, T* g_y, T* g_b, T* g_rate) {
    const T rate = T(1) / gain, log_rate = Real<T>::log(rate);
    for (int64_t u = 0; u < U; ++u) {
        PatchSpots<T> s;
        for (int k = 0; k < kK; ++k) {
            s.h[k] = height[k * U + u]; s.w[k] = width[k * U + u];
            s.cx[k] = x[k * U + u] + target[u * 2]; s.cy[k] = y[k * U + u] + target[u * 2 + 1];
        }
        s.b = background[u];
        auto run = [&](auto nm_tag) {
            constexpr int NMc = decltype(nm_tag)::value;
            T cfg[NMc][kK], Wm[NMc];
            for (int m = 0; m < NMc; ++m) { Wm[m] = W[m * U + u]; for (int k = 0; k < kK; ++k) cfg[m][k] = mcfg[m * kK + k]; }
            PatchOut<T, NMc> out; out.zero();
            for (int row = 0; row < P; ++row)
                for (int col = 0; col < P; ++col) {
                    T gxk[kK], gyk[kK];
                    for (int k = 0; k < kK; ++k) { gxk[k] = axis_factor<T>(col, s.cx[k], s.w[k]); gyk[k] = axis_factor<T>(row, s.cy[k], s.w[k]); }
                    pixel_accumulate<T, NMc, true>(value[(u * P + row) * P + col], gxk, gyk, col, row, s, cfg, rate, log_rate, O, off_s, off_w, Wm, out);
                }
            for (int m = 0; m < NMc; ++m) logp[m * U + u] = out.logp[m];
            g_b[u] = out.g_b; g_rate[u] = out.g_rate;
            for (int k = 0; k < kK; ++k) { g_h[k * U + u] = out.g_h[k]; g_w[k * U + u] = out.g_w[k]; g_x[k * U + u] = out.g_x[k]; g_y[k * U + u] = out.g_y[k]; }
        };
        if (NM == 1) run(std::integral_constant<int, 1>{}); else run(std::integral_constant<int, kM>{});
    }
}

extern "C" {
void hc_ksmogn_f64(int64_t U, int P, int O, int NM, const double* height, const double* width, const double* x, const double* y,
                   const double* background, double gain, const double* mcfg, const double* W, const double* value, const double* target,
                   const double* off_s, const double* off_w, double* logp, double* g_h, double* g_w, double* g_x, double* g_y, double* g_b, double* g_rate) {
    ksmogn_host<double>(U, P, O, NM, height, width, x, y, background, gain, mcfg, W, value, target, off_s, off_w, logp, g_h, g_w, g_x, g_y, g_b, g_rate);
}
void hc_ksmogn_f32(int64_t U, int P, int O, int NM, const float* height, const float* width, const float* x, const float* y,
                   const float* background, float gain, const float* mcfg, const float* W, const float* value, const float* target,
                   const float* off_s, const float* off_w, float* logp, float* g_h, float* g_w, float* g_x, float* g_y, float* g_b, float* g_rate) {
    ksmogn_host<float>(U, P, O, NM, height, width, x, y, background, gain, mcfg, W, value, target, off_s, off_w, logp, g_h, g_w, g_x, g_y, g_b, g_rate);
}
// fp32 production form (ksmogn_fast.cuh) with libm standing in for the MUFU approximations
void hc_ksmogn_fast_f32(int64_t U, int P, int O, int OC, const float* height, const float* width, const float* x, const float* y,
                        const float* background, float gain, const float* W, const float* value, const float* target,
                        const float* off_s, const float* off_w, float* logp, float* g_h, float* g_w, float* g_x, float* g_y, float* g_b, float* g_rate) {
    FastConst fc; fc.gain = gain; fc.rate = 1.0f / gain; fc.rate2 = fc.rate * kLog2e; fc.log_rate = logf(fc.rate);
    std::vector<float> w2(O);
    for (int j = 0; j < O; ++j) w2[j] = off_w[j] * kLog2e;
    for (int64_t u = 0; u < U; ++u) {
        PatchSpots<float> s; float norm[kK], iw[kK], Wr[kM], Wm[kM];
        for (int k = 0; k < kK; ++k) {
            s.h[k] = height[k * U + u]; s.w[k] = width[k * U + u];
            s.cx[k] = x[k * U + u] + target[u * 2]; s.cy[k] = y[k * U + u] + target[u * 2 + 1];
            norm[k] = 1.0f / (6.283185307179586f * s.w[k] * s.w[k]); iw[k] = 1.0f / s.w[k];
        }
        s.b = background[u];
        for (int m = 0; m < kM; ++m) { Wm[m] = W[m * U + u]; Wr[m] = Wm[m] * fc.rate; }
        const bool small = s.b * fc.rate < 4.0f;
        PatchOut<float, kM> out; out.zero();
        // the kernel's common case (P = 14, cached offsets, no small concentration, every pixel above every offset):
        // packed pairs of pixels (rows r, r + 7 of one column)
        bool pairs = P == 14 && OC >= 1 && OC <= 4 && OC == O && !small;
        if (pairs) {
            float max_off = off_s[0];
            for (int j = 1; j < O; ++j) max_off = std::max(max_off, off_s[j]);
            for (int p = 0; p < P * P; ++p) pairs = pairs && value[u * P * P + p] > max_off;
        }
        // O > 4 on a 14 x 14 patch: the one-pass many-bins form (every pixel above the smallest offset)
        bool many = P == 14 && OC == 0 && O > 4 && !small;
        if (many) {
            float min_off = off_s[0];
            for (int j = 1; j < O; ++j) min_off = std::min(min_off, off_s[j]);
            for (int p = 0; p < P * P; ++p) many = many && value[u * P * P + p] > min_off;
        }
        if (many) {
            const int ref = many_bins_reference(O, off_s);
            std::vector<BinConst> bins(O);
            for (int j = 0; j < O; ++j) bins[j] = many_bins_const(j, ref, off_s, w2.data(), fc.rate2);
            PairOut po; po.zero();
            for (int row = 0; row < 7; ++row)
                for (int col = 0; col < 14; ++col) {
                    float gxh[kK], dx[kK], dx2[kK]; F2 gyk[kK], dy[kK];
                    for (int k = 0; k < kK; ++k) {
                        gxh[k] = axis_factor<float>(col, s.cx[k], s.w[k]) * norm[k] * s.h[k];
                        gyk[k] = F2{axis_factor<float>(row, s.cy[k], s.w[k]), axis_factor<float>(row + 7, s.cy[k], s.w[k])};
                        dx[k] = float(col) - s.cx[k];
                        dx2[k] = dx[k] * dx[k];
                        dy[k] = F2{float(row) - s.cy[k], float(row + 7) - s.cy[k]};
                    }
                    const F2 D{value[(u * P + row) * P + col], value[(u * P + row + 7) * P + col]};
                    pixel_pair_many_bins(D, gxh, gyk, dx, dx2, dy, s, fc, O, bins.data(), off_s[ref], w2[ref], Wm, po);
                }
            finish_pair(po, fc.rate, out);
        } else if (pairs && OC == 1) {
            const SingleBinConst sc = single_bin_const(s.b, fc);
            const RowConst rc = row_const(off_s[0], s.b, fc);
            // the kernel's row mapping (sweep_patch_rows_single_bin): one row pair at a time in units of a = image / gain,
            // folded per row (the kernel: per lane) into the scalar record; the seven rows' records are then added (the
            // kernel: across lanes)
            for (int row = 0; row < 7; ++row) {
                RowOut pr; pr.zero();
                F2 gyr[kK], dy[kK];
                for (int k = 0; k < kK; ++k) {
                    gyr[k] = mul2(F2{axis_factor<float>(row, s.cy[k], s.w[k]), axis_factor<float>(row + 7, s.cy[k], s.w[k])}, f2(fc.rate));
                    dy[k] = F2{float(row) - s.cy[k], float(row + 7) - s.cy[k]};
                }
                for (int col = 0; col < 14; ++col) {
                    float gxh[kK], dx[kK], dx2[kK];
                    for (int k = 0; k < kK; ++k) {
                        gxh[k] = axis_factor<float>(col, s.cx[k], s.w[k]) * norm[k] * s.h[k];
                        dx[k] = float(col) - s.cx[k];
                        dx2[k] = dx[k] * dx[k];
                    }
                    const F2 D{value[(u * P + row) * P + col], value[(u * P + row + 7) * P + col]};
                    row_pair_single_bin(D, gxh, gyr, dx, dx2, sc, rc, Wm, pr);
                }
                PatchOut<float, kM> part; part.zero();
                finish_row_single_bin(pr, gyr, dy, sc, fc, w2[0] * kLn2, Wm, 28, part);
                for (int m = 0; m < kM; ++m) out.logp[m] += part.logp[m];
                out.g_b += part.g_b; out.g_rate += part.g_rate;
                for (int k = 0; k < kK; ++k) {
                    out.g_h[k] += part.g_h[k]; out.g_w[k] += part.g_w[k]; out.g_x[k] += part.g_x[k]; out.g_y[k] += part.g_y[k];
                }
            }
        } else if (pairs) {
            PairOut po; po.zero();
            auto run = [&](auto oc_tag) {
                constexpr int OCc = decltype(oc_tag)::value;
                float os[OCc], ow[OCc];
                for (int j = 0; j < OCc; ++j) { os[j] = off_s[j]; ow[j] = w2[j]; }
                for (int row = 0; row < 7; ++row)
                    for (int col = 0; col < 14; ++col) {
                        float gxh[kK], dx[kK], dx2[kK]; F2 gyk[kK], dy[kK];
                        for (int k = 0; k < kK; ++k) {
                            gxh[k] = axis_factor<float>(col, s.cx[k], s.w[k]) * norm[k] * s.h[k];
                            gyk[k] = F2{axis_factor<float>(row, s.cy[k], s.w[k]), axis_factor<float>(row + 7, s.cy[k], s.w[k])};
                            dx[k] = float(col) - s.cx[k];
                            dx2[k] = dx[k] * dx[k];
                            dy[k] = F2{float(row) - s.cy[k], float(row + 7) - s.cy[k]};
                        }
                        const F2 D{value[(u * P + row) * P + col], value[(u * P + row + 7) * P + col]};
                        pixel_pair_accumulate_fast<OCc>(D, gxh, gyk, dx, dx2, dy, s, fc, os, ow, Wm, po);
                    }
            };
            switch (OC) { case 1: run(std::integral_constant<int, 1>{}); break; case 2: run(std::integral_constant<int, 2>{}); break;
                          case 3: run(std::integral_constant<int, 3>{}); break; default: run(std::integral_constant<int, 4>{}); break; }
            finish_pair(po, fc.rate, out);
        } else
        for (int row = 0; row < P; ++row)
            for (int col = 0; col < P; ++col) {
                float gxk[kK], gyk[kK];
                for (int k = 0; k < kK; ++k) { gxk[k] = axis_factor<float>(col, s.cx[k], s.w[k]); gyk[k] = axis_factor<float>(row, s.cy[k], s.w[k]); }
                const float D = value[(u * P + row) * P + col];
                switch (OC) {
                    case 3: if (small) pixel_accumulate_fast<kM, 3, true, true>(D, gxk, gyk, col, row, s, norm, fc, O, off_s, w2.data(), Wm, Wr, out); else pixel_accumulate_fast<kM, 3, true, false>(D, gxk, gyk, col, row, s, norm, fc, O, off_s, w2.data(), Wm, Wr, out); break;
                    case 4: if (small) pixel_accumulate_fast<kM, 4, true, true>(D, gxk, gyk, col, row, s, norm, fc, O, off_s, w2.data(), Wm, Wr, out); else pixel_accumulate_fast<kM, 4, true, false>(D, gxk, gyk, col, row, s, norm, fc, O, off_s, w2.data(), Wm, Wr, out); break;
                    default: if (small) pixel_accumulate_fast<kM, 0, true, true>(D, gxk, gyk, col, row, s, norm, fc, O, off_s, w2.data(), Wm, Wr, out); else pixel_accumulate_fast<kM, 0, true, false>(D, gxk, gyk, col, row, s, norm, fc, O, off_s, w2.data(), Wm, Wr, out); break;
                }
            }
        finish_spot_moments(s, out);
        for (int m = 0; m < kM; ++m) logp[m * U + u] = out.logp[m];
        g_b[u] = out.g_b; g_rate[u] = out.g_rate;
        for (int k = 0; k < kK; ++k) { g_h[k * U + u] = out.g_h[k]; g_w[k * U + u] = out.g_w[k]; g_x[k * U + u] = out.g_x[k]; g_y[k * U + u] = out.g_y[k]; }
    }
}
double hc_digamma_f64(double x) { return digamma<double>(x); }
float hc_digamma_f32(float x) { return digamma<float>(x); }
double hc_std_gamma_grad_f64(double alpha, double x) { return std_gamma_grad<double>(alpha, x); }
float hc_std_gamma_grad_f32(float alpha, float x) { return std_gamma_grad<float>(alpha, x); }
double hc_beta_grad_f64(double x, double alpha, double total) { return beta_grad<double>(x, alpha, total); }
float hc_beta_grad_f32(float x, float alpha, float total) { return beta_grad<float>(x, alpha, total); }
void hc_beta_grad_pair_f64(double x, double c1, double c0, double* g1, double* g0) { beta_grad_pair<double>(x, c1, c0, *g1, *g0); }
double hc_lgamma_pos(double x) { return lgamma_pos(x); }
void hc_philox(uint64_t seed, uint64_t stream, uint64_t offset, int n, uint32_t* out) {
    Philox rng(seed, stream, offset);
    for (int i = 0; i < n; ++i) out[i] = rng.next();
}
// n Philox blocks as (normal, uniform, normal, uniform) rows -- what the Marsaglia-Tsang trials of the guide sites consume
void hc_normal_uniform_pairs(uint64_t seed, uint64_t stream, int n, float* out) {
    Philox rng(seed, stream, 0);
    for (int i = 0; i < n; ++i) rng.normal_uniform_pairs(out[4 * i], out[4 * i + 1], out[4 * i + 2], out[4 * i + 3]);
}
// n Beta(c1, c0)-distributed pairs of gamma draws sharing their trials (the form the AffineBeta sites use): g1, g2 per row
void hc_sample_gamma_pair_f32(uint64_t seed, uint64_t stream, float c1, float c0, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        Philox rng(seed, stream, (uint64_t)i * 64);
        GammaTrials trials;
        out[2 * i] = sample_std_gamma_f32(rng, trials, c1);
        out[2 * i + 1] = sample_std_gamma_f32(rng, trials, c0);
    }
}
void hc_sample_gamma_f64(uint64_t seed, uint64_t stream, double alpha, int n, double* out) {
    for (int i = 0; i < n; ++i) { Philox rng(seed, stream, (uint64_t)i * 64); out[i] = sample_std_gamma<double>(rng, alpha); }
}
}

// ---- whole-step emulation: the same sequence of device functions the kernels run -------------------
#include "cosmos_globals.cuh"
#include "cosmos_sites_fast.cuh"

namespace {

struct LocalOffsets {
    int64_t Nt, F, C;
    int64_t tensor_off(int t) const {
        const int64_t aoi = Nt * C, unit = Nt * F * C;
        if (t < 2) return t * aoi;
        if (t < 4) return 2 * aoi + (t - 2) * unit;
        return 2 * aoi + 2 * unit + (int64_t)(t - 4) * kK * unit;
    }
    // flat index of local record entry i for (aoi n, frame f, channel c)
    int64_t index(int i, int64_t n, int64_t f, int64_t c) const {
        if (i < 2) return tensor_off(i) + n * C + c;
        if (i < 4) return tensor_off(i) + (n * F + f) * C + c;
        const int t = 4 + (i - 4) / kK, k = (i - 4) % kK;
        return tensor_off(t) + ((k * Nt + n) * F + f) * C + c;
    }
};

// T: storage + pixel arithmetic type; F: arithmetic type of unit_post (site_eval is always double)
template <typename T, typename F>
double cosmos_step_host(int nb, int fb, int Nt, int F_, int C, int P, int O, const int32_t* ndx, const int32_t* fdx,
                        const T* pixels, const T* xy, const uint8_t* ontarget, const uint8_t* mask, const T* off_s,
                        const T* off_w, const ModelConst* mcp, double sN, double sF, const T* lparams, const double* gparams,
                        const T* lnoise, const double* gnoise, T* lgrads, double* ggrads, double* acc_out, T* samples_out) {
    const ModelConst& mc = *mcp;
    GlobalLayout gl{C};
    GlobalTables<double> gtd;
    double gvar[kMaxGlobalNoise], gsamp[kMaxGlobalNoise];
    for (int i = 0; i < gl.n_count(); ++i) gvar[i] = gnoise[i];
    globals_pre(gparams, gl, mc, false, nullptr, gvar, gsamp, gtd);
    GlobalTables<F> gt;
    gt.convert_from(gtd);
    LocalOffsets lo{Nt, F_, C};
    const int64_t U = (int64_t)nb * fb * C;
    std::vector<double> acc((size_t)C * NACC, 0.0);
    const int64_t total = lo.tensor_off(12);
    for (int64_t i = 0; i < total; ++i) lgrads[i] = T(0);
    const double s = sN * sF;
    T mcfg[kM][kK];
    for (int m = 0; m < kM; ++m) for (int k = 0; k < kK; ++k) mcfg[m][k] = T((m >> k) & 1);
    for (int64_t u = 0; u < U; ++u) {
        const int c = (int)(u % C), fi = (int)((u / C) % fb), ni = (int)(u / ((int64_t)C * fb));
        const int64_t n = ndx ? ndx[ni] : ni, f = fdx ? fdx[fi] : fi;
        // sites (double), results stored in T like the (NREC, U) device buffer
        T rec[NREC], sample[NSAMP], qm[kM], u_mp[kK];
        const double ubm = (double)lparams[lo.index(LP_BM, n, f, c)], ubs = (double)lparams[lo.index(LP_BS, n, f, c)];
        for (int st = 0; st < NSAMP; ++st) {
            double r[NSO], ex[NEX];
            double variate = (double)lnoise[st * U + u];
            bool done = false;
            if (sizeof(T) == sizeof(float)) {   // production path of site_kernel<float>
                float rf[NSO], exf[NEX], vf;
                if (site_eval_fast(st, (float)lparams[lo.index(site_param0(st), n, f, c)], (float)lparams[lo.index(site_param1(st), n, f, c)],
                                   (float)ubm, (float)ubs, mc, false, nullptr, variate, vf, rf, exf) == SITE_DONE) {
                    done = true;
                    sample[st] = (T)vf;
                    for (int j2 = 0; j2 < NSO; ++j2) rec[st * NSO + j2] = (T)rf[j2];
                    if (st == S_B) for (int j2 = 0; j2 < NEX; ++j2) rec[NSAMP * NSO + j2] = (T)exf[j2];
                }
            }
            if (done) continue;
            const double v = site_eval(st, (double)lparams[lo.index(site_param0(st), n, f, c)],
                                       (double)lparams[lo.index(site_param1(st), n, f, c)], ubm, ubs, mc, false, nullptr,
                                       variate, r, ex);
            sample[st] = (T)v;
            for (int j2 = 0; j2 < NSO; ++j2) rec[st * NSO + j2] = (T)r[j2];
            if (st == S_B) for (int j2 = 0; j2 < NEX; ++j2) rec[NSAMP * NSO + j2] = (T)ex[j2];
        }
        {
            T q1[kK], q0[kK];
            for (int k = 0; k < kK; ++k) {
                u_mp[k] = lparams[lo.index(LP_M_PROBS + k, n, f, c)];
                const SpotPresence<T> sp(u_mp[k], mc);
                q1[k] = sp.q1; q0[k] = sp.q0;
            }
            presence_weights<T>(q1, q0, qm);
        }
        if (samples_out) for (int i = 0; i < NSAMP; ++i) samples_out[i * U + u] = sample[i];
        // likelihood with W = q(m)
        const int64_t patch = (n * F_ + f) * C + c;
        PatchSpots<T> sp;
        for (int k = 0; k < kK; ++k) {
            sp.h[k] = sample[S_H + k]; sp.w[k] = sample[S_W + k];
            sp.cx[k] = sample[S_X + k] + xy[patch * 2]; sp.cy[k] = sample[S_Y + k] + xy[patch * 2 + 1];
        }
        sp.b = sample[S_B];
        PatchOut<T, kM> po; po.zero();
        for (int row = 0; row < P; ++row)
            for (int col = 0; col < P; ++col) {
                T gxk[kK], gyk[kK];
                for (int k = 0; k < kK; ++k) { gxk[k] = axis_factor<T>(col, sp.cx[k], sp.w[k]); gyk[k] = axis_factor<T>(row, sp.cy[k], sp.w[k]); }
                pixel_accumulate<T, kM, true>(pixels[(patch * P + row) * P + col], gxk, gyk, col, row, sp, mcfg, (T)gtd.rate, (T)gtd.log_rate, O, off_s, off_w, qm, po);
            }
        // post in F
        F recF[NREC], sampleF[NSAMP], gsF[NSAMP], LF[kM], umpF[kK];
        for (int i = 0; i < NREC; ++i) recF[i] = (F)rec[i];
        for (int i = 0; i < NSAMP; ++i) sampleF[i] = (F)sample[i];
        gsF[S_B] = (F)po.g_b;
        for (int k = 0; k < kK; ++k) { gsF[S_H + k] = (F)po.g_h[k]; gsF[S_W + k] = (F)po.g_w[k]; gsF[S_X + k] = (F)po.g_x[k]; gsF[S_Y + k] = (F)po.g_y[k]; }
        for (int m = 0; m < kM; ++m) LF[m] = (F)po.logp[m];
        for (int k = 0; k < kK; ++k) umpF[k] = (F)u_mp[k];
        UnitGrads<F> ug;
        unit_post<F>(recF, sampleF, LF, gsF, (F)po.g_rate, umpF, (F)ubm, (F)ubs, mc, gt, c, ontarget[n] != 0, fi == 0, ug);
        const double mu = mask[n] ? 1.0 : 0.0;
        for (int i = 0; i < NACC; ++i) acc[(size_t)c * NACC + i] += mu * (double)ug.acc[i];
        for (int i = 0; i < NLOCAL; ++i) lgrads[lo.index(i, n, f, c)] += (T)(-s * mu * (double)ug.g[i]);
        if (fi == 0) {
            double gbm, gbs;
            aoi_prior_grad(ubm, ubs, mc, gbm, gbs);
            lgrads[lo.index(LP_BM, n, f, c)] += (T)(-sN * mu * gbm);
            lgrads[lo.index(LP_BS, n, f, c)] += (T)(-sN * mu * gbs);
        }
    }
    for (size_t i = 0; i < acc.size(); ++i) acc_out[i] = acc[i];
    const double elbo = globals_post(gparams, gl, mc, gsamp, acc.data(), sN, sF, ggrads);
    return -elbo;
}

}  // namespace

// ---- hmm variant: the same device functions (cosmos_hmm.cuh, cosmos_globals.cuh with the hmm layout) in the order the
// kernels of csrc/cosmos_step.cu run them; the chain recursions are walked frame by frame here (the kernels scan them).
#include "cosmos_hmm.cuh"
#include "stats_math.cuh"

namespace {

// local flat buffer of the hmm variant: cosmos layout, then m_probs[z = 1] (K slabs), then z_trans (Nt, F, C, 2, 2)
struct HmmOffsets {
    LocalOffsets lo;
    int64_t std_numel() const { return lo.tensor_off(12); }
    int64_t mprobs1(int k, int64_t n, int64_t f, int64_t c) const { return std_numel() + (int64_t)k * (lo.Nt * lo.F * lo.C) + (n * lo.F + f) * lo.C + c; }
    int64_t ztrans(int64_t n, int64_t f, int64_t c) const { return std_numel() + (int64_t)kK * (lo.Nt * lo.F * lo.C) + ((n * lo.F + f) * lo.C + c) * (kZ * kZ); }
    int64_t numel() const { return std_numel() + (int64_t)kK * lo.Nt * lo.F * lo.C + lo.Nt * lo.F * lo.C * kZ * kZ; }
};

template <typename T>
double hmm_step_host(int nb, int Nt, int F_, int C, int P, int O, const int32_t* ndx, const T* pixels, const T* xy,
                     const uint8_t* ontarget, const uint8_t* mask, const T* off_s, const T* off_w, const ModelConst* mcp, double sN,
                     const T* lparams, const double* gparams, const T* lnoise, const double* gnoise, T* lgrads, double* ggrads,
                     double* acc_out = nullptr, double* hacc_out = nullptr) {
    const ModelConst& mc = *mcp;
    GlobalLayout gl{C, true};
    GlobalTables<double> gtd;
    double gvar[kMaxGlobalNoise], gsamp[kMaxGlobalNoise];
    for (int i = 0; i < gl.n_count(); ++i) gvar[i] = gnoise[i];
    globals_pre(gparams, gl, mc, false, nullptr, gvar, gsamp, gtd);
    GlobalTables<T> gt;
    gt.convert_from(gtd);
    HmmOffsets ho{{Nt, F_, C}};
    const LocalOffsets& lo = ho.lo;
    const int fb = F_;
    const int64_t U = (int64_t)nb * fb * C;
    std::vector<double> acc((size_t)C * NACC, 0.0), hacc((size_t)C * NHACC, 0.0);
    for (int64_t i = 0; i < ho.numel(); ++i) lgrads[i] = T(0);
    T mcfg[kM][kK];
    for (int m = 0; m < kM; ++m) for (int k = 0; k < kK; ++k) mcfg[m][k] = T((m >> k) & 1);
    std::vector<double> a_fwd((size_t)2 * fb), vdiff((size_t)fb);
    for (int ni = 0; ni < nb; ++ni)
        for (int c = 0; c < C; ++c) {
            const int64_t n = ndx ? ndx[ni] : ni;
            const double mu = mask[n] ? 1.0 : 0.0;
            const int ot = ontarget[n] ? 1 : 0;
            // forward marginals of the guide's chain
            double a0 = 1.0, a1 = 0.0;
            for (int f = 0; f < fb; ++f) {
                const T* zt = lparams + ho.ztrans(n, f, c);
                const ChainRow r0 = chain_row((double)zt[0], (double)zt[1], mc), r1 = chain_row((double)zt[2], (double)zt[3], mc);
                const double b0 = a0 * r0.q[0] + a1 * r1.q[0], b1 = a0 * r0.q[1] + a1 * r1.q[1];
                a0 = b0; a1 = b1;
                a_fwd[2 * f] = a0; a_fwd[2 * f + 1] = a1;
            }
            // per unit: sites, weights, likelihood, emission terms
            for (int f = 0; f < fb; ++f) {
                const int64_t u = ((int64_t)ni * fb + f) * C + c;
                T rec[NREC], sample[NSAMP], qm[kM], ump[kZ][kK], az[kZ] = {(T)a_fwd[2 * f], (T)a_fwd[2 * f + 1]};
                const double ubm = (double)lparams[lo.index(LP_BM, n, f, c)], ubs = (double)lparams[lo.index(LP_BS, n, f, c)];
                for (int st = 0; st < NSAMP; ++st) {
                    double r[NSO], ex[NEX];
                    double variate = (double)lnoise[st * U + u];
                    bool done = false;
                    if (sizeof(T) == sizeof(float)) {
                        float rf[NSO], exf[NEX], vf;
                        if (site_eval_fast(st, (float)lparams[lo.index(site_param0(st), n, f, c)], (float)lparams[lo.index(site_param1(st), n, f, c)],
                                           (float)ubm, (float)ubs, mc, false, nullptr, variate, vf, rf, exf) == SITE_DONE) {
                            done = true;
                            sample[st] = (T)vf;
                            for (int j2 = 0; j2 < NSO; ++j2) rec[st * NSO + j2] = (T)rf[j2];
                            if (st == S_B) for (int j2 = 0; j2 < NEX; ++j2) rec[NSAMP * NSO + j2] = (T)exf[j2];
                        }
                    }
                    if (done) continue;
                    const double v = site_eval(st, (double)lparams[lo.index(site_param0(st), n, f, c)],
                                               (double)lparams[lo.index(site_param1(st), n, f, c)], ubm, ubs, mc, false, nullptr, variate, r, ex);
                    sample[st] = (T)v;
                    for (int j2 = 0; j2 < NSO; ++j2) rec[st * NSO + j2] = (T)r[j2];
                    if (st == S_B) for (int j2 = 0; j2 < NEX; ++j2) rec[NSAMP * NSO + j2] = (T)ex[j2];
                }
                for (int k = 0; k < kK; ++k) { ump[0][k] = lparams[lo.index(LP_M_PROBS + k, n, f, c)]; ump[1][k] = lparams[ho.mprobs1(k, n, f, c)]; }
                hmm_presence_weights<T>(ump, az, mc, qm);
                const int64_t patch = (n * F_ + f) * C + c;
                PatchSpots<T> sp;
                for (int k = 0; k < kK; ++k) {
                    sp.h[k] = sample[S_H + k]; sp.w[k] = sample[S_W + k];
                    sp.cx[k] = sample[S_X + k] + xy[patch * 2]; sp.cy[k] = sample[S_Y + k] + xy[patch * 2 + 1];
                }
                sp.b = sample[S_B];
                PatchOut<T, kM> po; po.zero();
                for (int row = 0; row < P; ++row)
                    for (int col = 0; col < P; ++col) {
                        T gxk[kK], gyk[kK];
                        for (int k = 0; k < kK; ++k) { gxk[k] = axis_factor<T>(col, sp.cx[k], sp.w[k]); gyk[k] = axis_factor<T>(row, sp.cy[k], sp.w[k]); }
                        pixel_accumulate<T, kM, true>(pixels[(patch * P + row) * P + col], gxk, gyk, col, row, sp, mcfg, (T)gtd.rate, (T)gtd.log_rate, O, off_s, off_w, qm, po);
                    }
                T gs[NSAMP], Lm[kM];
                gs[S_B] = po.g_b;
                for (int k = 0; k < kK; ++k) { gs[S_H + k] = po.g_h[k]; gs[S_W + k] = po.g_w[k]; gs[S_X + k] = po.g_x[k]; gs[S_Y + k] = po.g_y[k]; }
                for (int m = 0; m < kM; ++m) Lm[m] = po.logp[m];
                UnitGrads<T> ug;
                HmmUnitOut<T> hu;
                unit_post_hmm<T>(rec, sample, Lm, gs, po.g_rate, ump, az, (T)ubm, (T)ubs, mc, gt, c, f == 0, ug, hu);
                for (int k = 0; k < kK; ++k) { ug.g[LP_M_PROBS + k] = hu.gmp[0][k]; lgrads[ho.mprobs1(k, n, f, c)] += (T)(-sN * mu * (double)hu.gmp[1][k]); }
                vdiff[f] = (double)hu.v[1] - (double)hu.v[0];
                for (int i = 0; i < NACC; ++i) acc[(size_t)c * NACC + i] += mu * (double)ug.acc[i];
                for (int i = 0; i < NLOCAL; ++i) lgrads[lo.index(i, n, f, c)] += (T)(-sN * mu * (double)ug.g[i]);
                if (f == 0) {
                    double gbm, gbs;
                    aoi_prior_grad(ubm, ubs, mc, gbm, gbs);
                    lgrads[lo.index(LP_BM, n, f, c)] += (T)(-sN * mu * gbm);
                    lgrads[lo.index(LP_BS, n, f, c)] += (T)(-sN * mu * gbs);
                }
            }
            // backward recursion (hmm_backward_kernel, frame by frame)
            const ChannelTables<double>& ct = gtd.ch[c];
            double carry = 0.0;
            for (int f = fb - 1; f >= 0; --f) {
                const int64_t iz = ho.ztrans(n, f, c);
                const ChainRow r0 = chain_row((double)lparams[iz + 0], (double)lparams[iz + 1], mc);
                const ChainRow r1 = chain_row((double)lparams[iz + 2], (double)lparams[iz + 3], mc);
                const double delta = vdiff[f] + carry;
                const double ap0 = f > 0 ? a_fwd[2 * (f - 1)] : 1.0, ap1 = f > 0 ? a_fwd[2 * (f - 1) + 1] : 0.0;
                double R0[kZ], R1[kZ];
                for (int z = 0; z < kZ; ++z) {
                    R0[z] = (f == 0 ? ct.logpz[ot][z] : ct.logptrans[ot][0][z]) - r0.lq[z];
                    R1[z] = (f == 0 ? ct.logpz[ot][z] : ct.logptrans[ot][1][z]) - r1.lq[z];
                }
                const double g0 = ap0 * r0.q[1] * r0.q[0] * (R0[1] - R0[0] + delta), g1 = ap1 * r1.q[1] * r1.q[0] * (R1[1] - R1[0] + delta);
                lgrads[iz + 0] = (T)(sN * mu * g0); lgrads[iz + 1] = (T)(-sN * mu * g0);
                lgrads[iz + 2] = (T)(sN * mu * g1); lgrads[iz + 3] = (T)(-sN * mu * g1);
                const double rho0 = r0.q[0] * R0[0] + r0.q[1] * R0[1], rho1 = r1.q[0] * R1[0] + r1.q[1] * R1[1];
                hacc[(size_t)c * NHACC + HACC_ELBO] += mu * (ap0 * rho0 + ap1 * rho1);
                if (ot) for (int z = 0; z < kZ; ++z) {
                    if (f == 0) hacc[(size_t)c * NHACC + HACC_INIT + z] += mu * ap0 * r0.q[z];
                    else { hacc[(size_t)c * NHACC + HACC_TRANS + z] += mu * ap0 * r0.q[z]; hacc[(size_t)c * NHACC + HACC_TRANS + kZ + z] += mu * ap1 * r1.q[z]; }
                }
                carry = (rho1 - rho0) + (r1.q[1] - r0.q[1]) * delta;
            }
        }
    // what a rank contributes to the cross-rank sum (the sharded step: all-reduce these, then hc_hmm_globals_post)
    if (acc_out) for (size_t i = 0; i < acc.size(); ++i) acc_out[i] = acc[i];
    if (hacc_out) for (size_t i = 0; i < hacc.size(); ++i) hacc_out[i] = hacc[i];
    double elbo = 0.0;
    for (int site = 0; site < global_site_count(gl.Q, true); ++site)
        elbo += globals_post_site(site, gparams, gl, mc, gsamp, acc.data(), sN, 1.0, ggrads, hacc.data());
    return -elbo;
}

}  // namespace

extern "C" {
double hc_hmm_step_f64(int nb, int Nt, int F, int C, int P, int O, const int32_t* ndx, const double* pixels, const double* xy,
                       const uint8_t* ontarget, const uint8_t* mask, const double* off_s, const double* off_w, const ModelConst* mc,
                       double sN, const double* lparams, const double* gparams, const double* lnoise, const double* gnoise,
                       double* lgrads, double* ggrads) {
    return hmm_step_host<double>(nb, Nt, F, C, P, O, ndx, pixels, xy, ontarget, mask, off_s, off_w, mc, sN, lparams, gparams, lnoise, gnoise, lgrads, ggrads);
}
// the sharded hmm step in two halves: local part of one rank (accumulators out), then -- after the cross-rank sum of
// (C, NACC) + (C, NHACC) doubles -- the replicated reverse mode of the global sites
double hc_hmm_step_acc_f64(int nb, int Nt, int F, int C, int P, int O, const int32_t* ndx, const double* pixels, const double* xy,
                           const uint8_t* ontarget, const uint8_t* mask, const double* off_s, const double* off_w, const ModelConst* mc,
                           double sN, const double* lparams, const double* gparams, const double* lnoise, const double* gnoise,
                           double* lgrads, double* ggrads, double* acc, double* hacc) {
    return hmm_step_host<double>(nb, Nt, F, C, P, O, ndx, pixels, xy, ontarget, mask, off_s, off_w, mc, sN, lparams, gparams, lnoise, gnoise, lgrads, ggrads, acc, hacc);
}
int hc_hmm_acc_sizes(int* nacc, int* nhacc) { *nacc = NACC; *nhacc = NHACC; return 0; }
double hc_hmm_globals_post(int C, const ModelConst* mc, const double* gparams, const double* gnoise, const double* acc,
                           const double* hacc, double sN, double* ggrads) {
    GlobalLayout gl{C, true};
    GlobalTables<double> gt;
    double gvar[kMaxGlobalNoise], gsamp[kMaxGlobalNoise];
    for (int i = 0; i < gl.n_count(); ++i) gvar[i] = gnoise[i];
    globals_pre(gparams, gl, *mc, false, nullptr, gvar, gsamp, gt);
    double elbo = 0.0;
    for (int site = 0; site < global_site_count(gl.Q, true); ++site)
        elbo += globals_post_site(site, gparams, gl, *mc, gsamp, acc, sN, 1.0, ggrads, hacc);
    return -elbo;
}
double hc_hmm_step_f32(int nb, int Nt, int F, int C, int P, int O, const int32_t* ndx, const float* pixels, const float* xy,
                       const uint8_t* ontarget, const uint8_t* mask, const float* off_s, const float* off_w, const ModelConst* mc,
                       double sN, const float* lparams, const double* gparams, const float* lnoise, const double* gnoise,
                       float* lgrads, double* ggrads) {
    return hmm_step_host<float>(nb, Nt, F, C, P, O, ndx, pixels, xy, ontarget, mask, off_s, off_w, mc, sN, lparams, gparams, lnoise, gnoise, lgrads, ggrads);
}
}

extern "C" {
double hc_cosmos_step_f64(int nb, int fb, int Nt, int F, int C, int P, int O, const int32_t* ndx, const int32_t* fdx,
                          const double* pixels, const double* xy, const uint8_t* ontarget, const uint8_t* mask,
                          const double* off_s, const double* off_w, const ModelConst* mc, double sN, double sF,
                          const double* lparams, const double* gparams, const double* lnoise, const double* gnoise,
                          double* lgrads, double* ggrads, double* acc_out, double* samples_out) {
    return cosmos_step_host<double, double>(nb, fb, Nt, F, C, P, O, ndx, fdx, pixels, xy, ontarget, mask, off_s, off_w, mc, sN, sF,
                                    lparams, gparams, lnoise, gnoise, lgrads, ggrads, acc_out, samples_out);
}
double hc_cosmos_step_f32(int nb, int fb, int Nt, int F, int C, int P, int O, const int32_t* ndx, const int32_t* fdx,
                          const float* pixels, const float* xy, const uint8_t* ontarget, const uint8_t* mask,
                          const float* off_s, const float* off_w, const ModelConst* mc, double sN, double sF,
                          const float* lparams, const double* gparams, const float* lnoise, const double* gnoise,
                          float* lgrads, double* ggrads, double* acc_out, float* samples_out) {
    return cosmos_step_host<float, float>(nb, fb, Nt, F, C, P, O, ndx, fdx, pixels, xy, ontarget, mask, off_s, off_w, mc, sN, sF,
                                   lparams, gparams, lnoise, gnoise, lgrads, ggrads, acc_out, samples_out);
}
// globals only: (replayed) sample -> reverse mode from the given accumulators; returns the loss
double hc_globals_post(int C, const ModelConst* mc, const double* gparams, const double* gnoise, const double* acc,
                       double sN, double sF, double* ggrads) {
    GlobalLayout gl{C};
    GlobalTables<double> gt;
    double gvar[kMaxGlobalNoise], gsamp[kMaxGlobalNoise];
    for (int i = 0; i < gl.n_count(); ++i) gvar[i] = gnoise[i];
    globals_pre(gparams, gl, *mc, false, nullptr, gvar, gsamp, gt);
    return -globals_post(gparams, gl, *mc, gsamp, acc, sN, sF, ggrads);
}
// RNG-mode guide draws of one site type: n samples of site `s` with unconstrained params (u0, u1)
void hc_site_draws(int s, double u0, double u1, const ModelConst* mc, uint64_t seed, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        Philox rng(seed, 7ull, ((uint64_t)(i + 1) << 12) + ((uint64_t)s << 8));
        double variate = 0.0, rec[NSO], ex[NEX];
        out[i] = site_eval(s, u0, u1, 0.0, 0.0, *mc, true, &rng, variate, rec, ex);
    }
}
void hc_gamma_draws_f32(float alpha, uint64_t seed, int n, float* out) {
    for (int i = 0; i < n; ++i) { Philox rng(seed, 3ull, (uint64_t)i << 8); out[i] = sample_std_gamma<float>(rng, alpha); }
}
int hc_sizeof_model_const() { return (int)sizeof(ModelConst); }
// one site in replay mode: double reference form and fp32 production form; out = [sample, rec[NSO], extra[NEX]]
void hc_site_eval_f64(int s, double u0, double u1, double ubm, double ubs, const ModelConst* mc, double variate, double* out) {
    double rec[NSO], ex[NEX] = {0, 0, 0, 0};
    out[0] = site_eval(s, u0, u1, ubm, ubs, *mc, false, nullptr, variate, rec, ex);
    for (int j = 0; j < NSO; ++j) out[1 + j] = rec[j];
    for (int j = 0; j < NEX; ++j) out[1 + NSO + j] = ex[j];
}
int hc_site_eval_fast(int s, float u0, float u1, float ubm, float ubs, const ModelConst* mc, double variate, double* out) {
    float rec[NSO], ex[NEX] = {0, 0, 0, 0}, v = 0.0f;
    const int status = site_eval_fast(s, u0, u1, ubm, ubs, *mc, false, nullptr, variate, v, rec, ex);
    out[0] = v;
    for (int j = 0; j < NSO; ++j) out[1 + j] = rec[j];
    for (int j = 0; j < NEX; ++j) out[1 + NSO + j] = ex[j];
    return status;
}
// the two-pass form site_fast_kernel runs (MODE 1: draw + classify + bulk regime; MODE 2: replay of a deferred site), RNG mode;
// out as above, returns the final status, *deferred_cls = regime class or -1 when pass 1 finished the site
int hc_site_eval_fast_split(int s, float u0, float u1, float ubm, float ubs, const ModelConst* mc, uint64_t seed, double* out, int* deferred_cls,
                            double* variate_out) {
    Philox rng(seed, 11ull, ((uint64_t)(s + 1) << 8));
    float rec[NSO], ex[NEX] = {0, 0, 0, 0}, v = 0.0f;
    double variate = 0.0;
    int cls = -1;
    int status = site_eval_fast_t<1>(s, u0, u1, ubm, ubs, *mc, true, &rng, variate, v, rec, ex, cls);
    *deferred_cls = -1;
    if (status == SITE_DEFER) {
        *deferred_cls = cls;
        status = site_eval_fast_t<2>(s, u0, u1, ubm, ubs, *mc, false, nullptr, variate, v, rec, ex, cls);
    }
    *variate_out = variate;
    out[0] = v;
    for (int j = 0; j < NSO; ++j) out[1 + j] = rec[j];
    for (int j = 0; j < NEX; ++j) out[1 + NSO + j] = ex[j];
    return status;
}
// ... and the one-call form on the same RNG stream
int hc_site_eval_fast_rng(int s, float u0, float u1, float ubm, float ubs, const ModelConst* mc, uint64_t seed, double* out, double* variate_out) {
    Philox rng(seed, 11ull, ((uint64_t)(s + 1) << 8));
    float rec[NSO], ex[NEX] = {0, 0, 0, 0}, v = 0.0f;
    double variate = 0.0;
    const int status = site_eval_fast(s, u0, u1, ubm, ubs, *mc, true, &rng, variate, v, rec, ex);
    *variate_out = variate;
    out[0] = v;
    for (int j = 0; j < NSO; ++j) out[1 + j] = rec[j];
    for (int j = 0; j < NEX; ++j) out[1 + NSO + j] = ex[j];
    return status;
}
// status of the production form for one RNG-mode evaluation (profiles/fallback_probe.py)
int hc_site_status_rng(int s, float u0, float u1, float ubm, float ubs, const ModelConst* mc, uint64_t seed) {
    Philox rng(seed, 11ull, ((uint64_t)(s + 1) << 8));
    double variate = 0.0;
    float rec[NSO], ex[NEX], v = 0.0f;
    return site_eval_fast(s, u0, u1, ubm, ubs, *mc, true, &rng, variate, v, rec, ex);
}
// the same for n sites at once, returning the base draws too (profiles/site_regimes.py)
void hc_site_status_batch(int s, int n, const float* u0, const float* u1, const float* ubm, const float* ubs, const ModelConst* mc,
                          uint64_t seed, int* status, double* variate) {
    for (int i = 0; i < n; ++i) {
        Philox rng(seed, 11ull, ((uint64_t)(i + 1) << 12) + ((uint64_t)(s + 1) << 8));
        float rec[NSO], ex[NEX], v = 0.0f;
        variate[i] = 0.0;
        status[i] = site_eval_fast(s, u0[i], u1[i], ubm ? ubm[i] : 0.0f, ubs ? ubs[i] : 0.0f, *mc, true, &rng, variate[i], v, rec, ex);
    }
}
// RNG-mode draws through the production form (falls back like site_kernel<float>)
void hc_site_draws_fast(int s, float u0, float u1, const ModelConst* mc, uint64_t seed, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        Philox rng(seed, 7ull, ((uint64_t)(i + 1) << 12) + ((uint64_t)s << 8));
        double variate = 0.0, rec[NSO], ex[NEX];
        float recf[NSO], exf[NEX], v = 0.0f;
        const int status = site_eval_fast(s, u0, u1, 0.0f, 0.0f, *mc, true, &rng, variate, v, recf, exf);
        out[i] = status == SITE_DONE ? (double)v : site_eval(s, u0, u1, 0.0, 0.0, *mc, status == SITE_FALLBACK_DRAW, &rng, variate, rec, ex);
    }
}
}

// ---- inverse CDFs of the credible intervals (stats_math.cuh) ------------------------------------------------------------
extern "C" {
void hc_gamma_interval(int64_t n, const double* conc, const double* rate, double ci, double* lo, double* hi) {
    for (int64_t i = 0; i < n; ++i) {
        lo[i] = tq::gamma_p_inv(0.5 * (1.0 - ci), conc[i]) / rate[i];
        hi[i] = tq::gamma_p_inv(0.5 * (1.0 + ci), conc[i]) / rate[i];
    }
}
void hc_beta_interval(int64_t n, const double* c1, const double* c0, double ci, double* lo, double* hi) {
    for (int64_t i = 0; i < n; ++i) {
        lo[i] = tq::beta_i_inv(0.5 * (1.0 - ci), c1[i], c0[i]);
        hi[i] = tq::beta_i_inv(0.5 * (1.0 + ci), c1[i], c0[i]);
    }
}
}
