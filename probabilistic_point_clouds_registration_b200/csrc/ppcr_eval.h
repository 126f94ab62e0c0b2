// ppcr_eval.h -- per-source-point arithmetic of the fused "weights + normal equations" pass.
//
// One source point i with its k_i associated target points y_j produces
//   w_ij  : probabilistic_weights.hpp:48-105, evaluated from the residuals at pose_w
//           (the pose of the last WeightUpdaterCallback, weight_updater_callback.hpp:36-63), and
//   r_ij  : error_term.hpp:21-37 at pose_e (the pose Ceres is evaluating),
// folded into the 24 moments of ppcr_lm.h.  The row softmax of the reference (max-subtracted log-sum-exp over the
// row, hpp:77-99) is computed in ONE pass with a running maximum, so neither the weights nor the residuals are
// ever stored: the pass reads 12 bytes per correspondence and 20 bytes per source point.
//
// Plain C++ qualified PPCR_HD: the eval kernel and the CPU tests of the host logic share this source.
#ifndef PPCR_EVAL_H
#define PPCR_EVAL_H

#include "ppcr_lm.h"

namespace ppcr {

struct WeightCfg {
    double dof;        // v
    double t_exponent; // -(v + 3) / 2        (probabilistic_weights.hpp:37, DIMENSIONS = 3)
    double dof_plus_d; // v + 3               (numerator of the expected weight, hpp:73)
    double inv_dof;    // 1 / v is NOT used for the log argument (the reference divides); kept for the fp32 path
    int32_t is_normal; // v == +inf: Gaussian model, w = softmax(-r^2/2)
    int32_t pad;
    // float32 copies for the fast path: u = s^h with s = 1/(1 + r^2/v), h = (v+d)/2; expected weight = f_c * s
    float f_dof, f_t_exponent, f_dof_plus_d, f_inv_dof;
    float f_c;         // (v + d) / v
    float f_h;         // (v + d) / 2
    int32_t pow_int;   // floor(h) when 2h is an integer <= 31 (h by repeated multiplication, +sqrt for the half), else -1
    int32_t pow_half;  // 2h odd
};

PPCR_HD WeightCfg make_weight_cfg(double dof, int dimension = 3)
{
    const double dim = static_cast<double>(dimension);  // DIMENSIONS = 3 in production (iteration.hpp:17,29)
    WeightCfg w;
    w.is_normal = !(dof < 1.7976931348623157e308);
    w.dof = dof;
    w.t_exponent = w.is_normal ? 0.0 : -(dof + dim) / 2.0;
    w.dof_plus_d = dof + dim;
    w.inv_dof = w.is_normal ? 0.0 : 1.0 / dof;
    w.pad = 0;
    w.f_dof = w.is_normal ? 0.f : static_cast<float>(dof);
    w.f_t_exponent = static_cast<float>(w.t_exponent);
    w.f_dof_plus_d = w.is_normal ? 0.f : static_cast<float>(w.dof_plus_d);
    w.f_inv_dof = static_cast<float>(w.inv_dof);
    w.f_c = w.is_normal ? 0.f : static_cast<float>(w.dof_plus_d / dof);
    w.f_h = w.is_normal ? 0.f : static_cast<float>(w.dof_plus_d / 2.0);
    w.pow_int = -1;
    w.pow_half = 0;
    if (!w.is_normal) {
        const double twice = w.dof_plus_d;  // 2h
        const long long t = static_cast<long long>(twice);
        if (static_cast<double>(t) == twice && t >= 0 && t <= 31) {
            w.pow_int = static_cast<int32_t>(t / 2);
            w.pow_half = static_cast<int32_t>(t & 1);
        }
    }
    return w;
}

// log-probability up to the row-constant normaliser (which cancels in the softmax) and expected weight
template <bool kFast>
PPCR_HD void log_prob(const WeightCfg& wc, double r2, double* lp, double* expected)
{
    if (wc.is_normal) {
        *lp = -r2 / 2.0;
        *expected = 1.0;
    } else if (kFast) {
        const float z = static_cast<float>(r2 / wc.dof);
        *lp = static_cast<double>(static_cast<float>(wc.t_exponent) * log1pf(z));
        *expected = static_cast<double>(static_cast<float>(wc.dof_plus_d) / (static_cast<float>(wc.dof) + static_cast<float>(r2)));
    } else {
        *lp = wc.t_exponent * log1p(r2 / wc.dof);
        *expected = wc.dof_plus_d / (wc.dof + r2);
    }
}

template <bool kFast>
PPCR_HD double exp_diff(double d)
{
    if (kFast) return static_cast<double>(expf(static_cast<float>(d)));
    return exp(d);
}

struct RowAcc {  // running softmax state of one source row
    double m;    // running max of the log-probabilities
    double a0;   // sum exp(l - m)
    double a1;   // sum exp(l - m) e
    double ar[3];// sum exp(l - m) e r
    double ac;   // sum exp(l - m) e |r|^2
};

PPCR_HD void row_begin(RowAcc* a)
{
    a->m = -1.7976931348623157e308;
    a->a0 = a->a1 = a->ac = 0.0;
    a->ar[0] = a->ar[1] = a->ar[2] = 0.0;
}

// one correspondence: target point (yx,yy,yz); pe = R_e x + t_e, pw = R_w x + t_w already computed for the row
template <bool kFast>
PPCR_HD void row_add(RowAcc* a, const WeightCfg& wc, double yx, double yy, double yz, const double* pe, const double* pw)
{
    const double wx = yx - pw[0], wy = yy - pw[1], wz = yz - pw[2];
    const double r2w = wx * wx + wy * wy + wz * wz;  // squared error the callback hands to updateWeights
    double lp, ex;
    log_prob<kFast>(wc, r2w, &lp, &ex);
    const double ex_ = ex;
    const double rx = yx - pe[0], ry = yy - pe[1], rz = yz - pe[2];
    const double r2e = rx * rx + ry * ry + rz * rz;
    if (lp > a->m) {  // new row maximum: rescale what has been accumulated so far
        const double sc = exp_diff<kFast>(a->m - lp);
        a->a0 *= sc;
        a->a1 *= sc;
        a->ar[0] *= sc;
        a->ar[1] *= sc;
        a->ar[2] *= sc;
        a->ac *= sc;
        a->m = lp;
    }
    const double p = exp_diff<kFast>(lp - a->m);
    const double pe_w = p * ex_;
    a->a0 += p;
    a->a1 += pe_w;
    a->ar[0] += pe_w * rx;
    a->ar[1] += pe_w * ry;
    a->ar[2] += pe_w * rz;
    a->ac += pe_w * r2e;
}

// fold a finished row into the 24 moments; (sx,sy,sz) is the source point in double
// STRIDE: distance between consecutive moments in `acc` (1 = a plain array; the eval kernel keeps one column per
// thread in shared memory, STRIDE = block size, so that the 24 accumulators do not occupy 48 registers)
template <int STRIDE>
PPCR_HD void row_end_s(const RowAcc* a, double sx, double sy, double sz, double* acc)
{
    const double inv = a->a0 > 0.0 ? 1.0 / a->a0 : 0.0;  // every posterior underflowed: the row carries no weight
    const double W = a->a1 * inv;
    const double rho[3] = {a->ar[0] * inv, a->ar[1] * inv, a->ar[2] * inv};
    acc[(M_S0) * STRIDE] += W;
    acc[(M_S1 + 0) * STRIDE] += W * sx;
    acc[(M_S1 + 1) * STRIDE] += W * sy;
    acc[(M_S1 + 2) * STRIDE] += W * sz;
    acc[(M_S2 + 0) * STRIDE] += W * sx * sx;
    acc[(M_S2 + 1) * STRIDE] += W * sx * sy;
    acc[(M_S2 + 2) * STRIDE] += W * sx * sz;
    acc[(M_S2 + 3) * STRIDE] += W * sy * sy;
    acc[(M_S2 + 4) * STRIDE] += W * sy * sz;
    acc[(M_S2 + 5) * STRIDE] += W * sz * sz;
    acc[(M_SR + 0) * STRIDE] += rho[0];
    acc[(M_SR + 1) * STRIDE] += rho[1];
    acc[(M_SR + 2) * STRIDE] += rho[2];
    acc[(M_C + 0) * STRIDE] += sx * rho[0];
    acc[(M_C + 1) * STRIDE] += sx * rho[1];
    acc[(M_C + 2) * STRIDE] += sx * rho[2];
    acc[(M_C + 3) * STRIDE] += sy * rho[0];
    acc[(M_C + 4) * STRIDE] += sy * rho[1];
    acc[(M_C + 5) * STRIDE] += sy * rho[2];
    acc[(M_C + 6) * STRIDE] += sz * rho[0];
    acc[(M_C + 7) * STRIDE] += sz * rho[1];
    acc[(M_C + 8) * STRIDE] += sz * rho[2];
    acc[(M_COST) * STRIDE] += 0.5 * a->ac * inv;
    acc[(M_ROWS) * STRIDE] += 1.0;
}

PPCR_HD void row_end(const RowAcc* a, double sx, double sy, double sz, double* acc) { row_end_s<1>(a, sx, sy, sz, acc); }

// ---- fast path ------------------------------------------------------------------------------------------------
//
// The same row arithmetic with float32 transcendentals and float32 in-row sums.  The residuals keep (nearly) full
// float32 relative precision although they are differences of ~10..100 m coordinates: the transformed source point
// is held as an unevaluated float pair (hi + lo), and y - hi is exact or within one rounding, so
// r = (y - hi) - lo carries ~1e-7 relative error instead of 1e-7 * |y| absolute.  The weight residual (taken at
// pose_w) is r_w = r_e + (p_e - p_w): the pose difference is a per-row constant, computed in float64.  Weights then
// agree with the float64 formula to a few 1e-7 relative (the bar is 1e-5); everything that is summed ACROSS rows
// stays float64.
//
// The per-correspondence arithmetic is instantiated per weight model (WM_*) and per "pose_w == pose_e", both uniform
// over a launch, so the inner loop carries no branch on either.

enum WeightMode {
    WM_T_H4 = 0,    // t, (v + d) / 2 == 4  (the reference's default dof = 5): u = s^4, two multiplications
    WM_T_INT = 1,   // t, 2h an integer <= 31: repeated multiplication (+ a square root for the half)
    WM_T_REAL = 2,  // t, any other dof: u = exp2(h * log2(s))
    WM_GAUSS = 3,   // dof = +inf: softmax(-r^2 / 2) with a running maximum
};

PPCR_HD int weight_mode(const WeightCfg& wc)
{
    if (wc.is_normal) return WM_GAUSS;
    if (wc.pow_int == 4 && !wc.pow_half) return WM_T_H4;
    if (wc.pow_int >= 0) return WM_T_INT;
    return WM_T_REAL;
}

struct PointHL {  // p = hi + lo, component-wise
    float hi[3], lo[3];
};

PPCR_HD void split_point(const double* p, PointHL* out)
{
    for (int k = 0; k < 3; ++k) {
        out->hi[k] = static_cast<float>(p[k]);
        out->lo[k] = static_cast<float>(p[k] - static_cast<double>(out->hi[k]));
    }
}

// d = p_e - p_w, so that  y - p_w = (y - p_e) + d
PPCR_HD void pose_delta(const double* pe, const double* pw, float* d)
{
    for (int k = 0; k < 3; ++k) d[k] = static_cast<float>(pe[k] - pw[k]);
}

struct RowAccF {
    float m;      // running max of the log-probabilities (Gaussian model only)
    float a0;     // sum exp(l - m)
    float a1;     // sum exp(l - m) e
    float ar[3];  // sum exp(l - m) e r
    float ac;     // sum exp(l - m) e |r|^2
};

PPCR_HD void rowf_begin(RowAccF* a)
{
    a->m = -3.0e38f;
    a->a0 = a->a1 = a->ac = 0.f;
    a->ar[0] = a->ar[1] = a->ar[2] = 0.f;
}

PPCR_HD float residual_hl(float y, float hi, float lo)
{
#if defined(__CUDA_ARCH__)
    return __fsub_rn(__fsub_rn(y, hi), lo);
#else
    return (y - hi) - lo;
#endif
}

// 1/x for x in [1, 2^60): one MUFU.RCP on the device (<= 1 ulp), a division on the host
PPCR_HD float f_rcp(float x)
{
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}

PPCR_HD float f_exp(float x)
{
#if defined(__CUDA_ARCH__)
    return __expf(x);
#else
    return expf(x);
#endif
}

// t-distribution: unnormalised posterior u = (1 + r^2/v)^(-(v+d)/2) and u * expected weight, without log / exp.
// u <= 1 and the nearest neighbour keeps the row sum away from zero, so the max-subtraction of the reference's
// log-sum-exp (which only guards against overflow / total underflow) has nothing to do here.
template <int WM>
PPCR_HD void t_terms(const WeightCfg& wc, float r2w, float* u_out, float* ue_out)
{
    const float s = f_rcp(r2w * wc.f_inv_dof + 1.0f);
    float u;
    if (WM == WM_T_H4) {
        const float s2 = s * s;
        u = s2 * s2;
    } else if (WM == WM_T_INT) {
        const float s2 = s * s, s4 = s2 * s2, s8 = s4 * s4;
        u = (wc.pow_int & 1) ? s : 1.0f;
        if (wc.pow_int & 2) u *= s2;
        if (wc.pow_int & 4) u *= s4;
        if (wc.pow_int & 8) u *= s8;
        if (wc.pow_half) u *= sqrtf(s);
    } else {
#if defined(__CUDA_ARCH__)
        u = exp2f(wc.f_h * __log2f(s));
#else
        u = exp2f(wc.f_h * log2f(s));
#endif
    }
    *u_out = u;
    *ue_out = u * (wc.f_c * s);
}

// One correspondence.  SAME: pose_w == pose_e (first evaluation of an outer iteration), the weight residual is the
// cost residual; otherwise dw = p_e - p_w of this row (pose_delta).
template <int WM, bool SAME>
PPCR_HD void rowf_add_t(RowAccF* a, const WeightCfg& wc, float yx, float yy, float yz, const PointHL& pe, const float* dw)
{
    const float rx = residual_hl(yx, pe.hi[0], pe.lo[0]);
    const float ry = residual_hl(yy, pe.hi[1], pe.lo[1]);
    const float rz = residual_hl(yz, pe.hi[2], pe.lo[2]);
    const float r2e = rx * rx + ry * ry + rz * rz;
    float r2w = r2e;
    if (!SAME) {
        const float wx = rx + dw[0], wy = ry + dw[1], wz = rz + dw[2];
        r2w = wx * wx + wy * wy + wz * wz;
    }
    float p, pw_;
    if (WM == WM_GAUSS) {
        const float lp = -0.5f * r2w;
        if (lp > a->m) {  // new row maximum: rescale what has been accumulated so far
            const float sc = f_exp(a->m - lp);
            a->a0 *= sc;
            a->a1 *= sc;
            a->ar[0] *= sc;
            a->ar[1] *= sc;
            a->ar[2] *= sc;
            a->ac *= sc;
            a->m = lp;
        }
        p = f_exp(lp - a->m);
        pw_ = p;
    } else {
        t_terms<WM>(wc, r2w, &p, &pw_);
    }
    a->a0 += p;
    a->a1 += pw_;
    a->ar[0] += pw_ * rx;
    a->ar[1] += pw_ * ry;
    a->ar[2] += pw_ * rz;
    a->ac += pw_ * r2e;
}

// run-time dispatch of the above (host emulation, parity dumps)
PPCR_HD void rowf_add(RowAccF* a, const WeightCfg& wc, float yx, float yy, float yz, const PointHL& pe, const float* dw,
                      bool same_pose)
{
    switch (weight_mode(wc) * 2 + (same_pose ? 1 : 0)) {
        case WM_T_H4 * 2: rowf_add_t<WM_T_H4, false>(a, wc, yx, yy, yz, pe, dw); break;
        case WM_T_H4 * 2 + 1: rowf_add_t<WM_T_H4, true>(a, wc, yx, yy, yz, pe, dw); break;
        case WM_T_INT * 2: rowf_add_t<WM_T_INT, false>(a, wc, yx, yy, yz, pe, dw); break;
        case WM_T_INT * 2 + 1: rowf_add_t<WM_T_INT, true>(a, wc, yx, yy, yz, pe, dw); break;
        case WM_T_REAL * 2: rowf_add_t<WM_T_REAL, false>(a, wc, yx, yy, yz, pe, dw); break;
        case WM_T_REAL * 2 + 1: rowf_add_t<WM_T_REAL, true>(a, wc, yx, yy, yz, pe, dw); break;
        case WM_GAUSS * 2: rowf_add_t<WM_GAUSS, false>(a, wc, yx, yy, yz, pe, dw); break;
        default: rowf_add_t<WM_GAUSS, true>(a, wc, yx, yy, yz, pe, dw); break;
    }
}

// fold a finished float32 row into the 24 float64 moments: the row quotients are float32 (like the row sums), the
// products with the source point and everything summed across rows are float64
template <int STRIDE>
PPCR_HD void rowf_end_s(const RowAccF* a, double sx, double sy, double sz, double* acc)
{
    const float inv = a->a0 > 0.f ? 1.0f / a->a0 : 0.f;  // every posterior underflowed: the row carries no weight
    const double W = static_cast<double>(a->a1 * inv);
    const double rho[3] = {static_cast<double>(a->ar[0] * inv), static_cast<double>(a->ar[1] * inv),
                           static_cast<double>(a->ar[2] * inv)};
    const double Wx = W * sx, Wy = W * sy, Wz = W * sz;
    acc[(M_S0) * STRIDE] += W;
    acc[(M_S1 + 0) * STRIDE] += Wx;
    acc[(M_S1 + 1) * STRIDE] += Wy;
    acc[(M_S1 + 2) * STRIDE] += Wz;
    acc[(M_S2 + 0) * STRIDE] += Wx * sx;
    acc[(M_S2 + 1) * STRIDE] += Wx * sy;
    acc[(M_S2 + 2) * STRIDE] += Wx * sz;
    acc[(M_S2 + 3) * STRIDE] += Wy * sy;
    acc[(M_S2 + 4) * STRIDE] += Wy * sz;
    acc[(M_S2 + 5) * STRIDE] += Wz * sz;
    acc[(M_SR + 0) * STRIDE] += rho[0];
    acc[(M_SR + 1) * STRIDE] += rho[1];
    acc[(M_SR + 2) * STRIDE] += rho[2];
    acc[(M_C + 0) * STRIDE] += sx * rho[0];
    acc[(M_C + 1) * STRIDE] += sx * rho[1];
    acc[(M_C + 2) * STRIDE] += sx * rho[2];
    acc[(M_C + 3) * STRIDE] += sy * rho[0];
    acc[(M_C + 4) * STRIDE] += sy * rho[1];
    acc[(M_C + 5) * STRIDE] += sy * rho[2];
    acc[(M_C + 6) * STRIDE] += sz * rho[0];
    acc[(M_C + 7) * STRIDE] += sz * rho[1];
    acc[(M_C + 8) * STRIDE] += sz * rho[2];
    acc[(M_COST) * STRIDE] += static_cast<double>(0.5f * a->ac * inv);
    acc[(M_ROWS) * STRIDE] += 1.0;
}
PPCR_HD void rowf_end(const RowAccF* a, double sx, double sy, double sz, double* acc) { rowf_end_s<1>(a, sx, sy, sz, acc); }

// The same 24 contributions of a finished row as float32 values (v[m], moment index m as above): the source coordinates are
// float32 to begin with and the row quotients are, so every product carries one more float32 rounding (6e-8 relative) on top of
// the row sums' own.  Used by the evaluation kernel, which adds the rows of a warp in float32 (a fixed tree over 32 rows) before
// anything goes into a float64 accumulator.
PPCR_HD void rowf_end_f(const RowAccF* a, float sx, float sy, float sz, float* v)
{
    const float inv = a->a0 > 0.f ? 1.0f / a->a0 : 0.f;
    const float W = a->a1 * inv;
    const float r0 = a->ar[0] * inv, r1 = a->ar[1] * inv, r2 = a->ar[2] * inv;
    const float Wx = W * sx, Wy = W * sy, Wz = W * sz;
    v[M_S0] = W;
    v[M_S1 + 0] = Wx;
    v[M_S1 + 1] = Wy;
    v[M_S1 + 2] = Wz;
    v[M_S2 + 0] = Wx * sx;
    v[M_S2 + 1] = Wx * sy;
    v[M_S2 + 2] = Wx * sz;
    v[M_S2 + 3] = Wy * sy;
    v[M_S2 + 4] = Wy * sz;
    v[M_S2 + 5] = Wz * sz;
    v[M_SR + 0] = r0;
    v[M_SR + 1] = r1;
    v[M_SR + 2] = r2;
    v[M_C + 0] = sx * r0;
    v[M_C + 1] = sx * r1;
    v[M_C + 2] = sx * r2;
    v[M_C + 3] = sy * r0;
    v[M_C + 4] = sy * r1;
    v[M_C + 5] = sy * r2;
    v[M_C + 6] = sz * r0;
    v[M_C + 7] = sz * r1;
    v[M_C + 8] = sz * r2;
    v[M_COST] = 0.5f * a->ac * inv;
    v[M_ROWS] = 1.0f;
}

// weight of one correspondence once the row statistics are known (parity dumps only); pw = p_w as a float pair
PPCR_HD float rowf_finished_weight(const RowAccF* a, const WeightCfg& wc, float yx, float yy, float yz, const PointHL& pw)
{
    const float wx = residual_hl(yx, pw.hi[0], pw.lo[0]);
    const float wy = residual_hl(yy, pw.hi[1], pw.lo[1]);
    const float wz = residual_hl(yz, pw.hi[2], pw.lo[2]);
    const float r2w = wx * wx + wy * wy + wz * wz;
    if (wc.is_normal) return f_exp(-0.5f * r2w - a->m) / a->a0;
    float u, ue;
    switch (weight_mode(wc)) {
        case WM_T_H4: t_terms<WM_T_H4>(wc, r2w, &u, &ue); break;
        case WM_T_INT: t_terms<WM_T_INT>(wc, r2w, &u, &ue); break;
        default: t_terms<WM_T_REAL>(wc, r2w, &u, &ue); break;
    }
    return ue / a->a0;
}

PPCR_HD void apply_pose(const Pose& p, double sx, double sy, double sz, double* out)
{
    out[0] = p.R[0] * sx + p.R[1] * sy + p.R[2] * sz + p.t[0];
    out[1] = p.R[3] * sx + p.R[4] * sy + p.R[5] * sz + p.t[1];
    out[2] = p.R[6] * sx + p.R[7] * sy + p.R[8] * sz + p.t[2];
}

// weight of one correspondence once the row statistics (m, a0) are known -- used only when the weights
// themselves are requested (parity dumps); the solver never materialises them.
template <bool kFast>
PPCR_HD double finished_weight(const RowAcc* a, const WeightCfg& wc, double yx, double yy, double yz, const double* pw)
{
    const double wx = yx - pw[0], wy = yy - pw[1], wz = yz - pw[2];
    const double r2w = wx * wx + wy * wy + wz * wz;
    double lp, ex;
    log_prob<kFast>(wc, r2w, &lp, &ex);
    return exp_diff<kFast>(lp - a->m) / a->a0 * ex;
}

}  // namespace ppcr
#endif
