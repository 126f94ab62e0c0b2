// emu_host_logic.cpp -- CPU harness for the PRODUCT's scalar host/device-shared logic (test infrastructure).
//
// It compiles csrc/ppcr_lm.h (LM controller, moment expansion, outer loop) and csrc/ppcr_eval.h (per-row
// weights + moments) with g++ and drives them with a brute-force neighbour search, so the state machine that
// k_controller / k_eval run on the GPU can be compared with the oracle in the CPU-only test tier.
// The neighbour search here is deliberately naive (exact (d2, index) top-m by sorting); the GPU search kernel
// is covered by the -m gpu tests.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../probabilistic_point_clouds_registration_b200/csrc/ppcr_eval.h"
#include "../../probabilistic_point_clouds_registration_b200/csrc/ppcr_lm.h"

using namespace ppcr;

namespace {

struct Cloud {
    std::vector<float> p;  // xyzw
    int64_t n;
};

float d2f(const float* a, const float* b)
{
    float dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    float acc = dx * dx;
    acc = acc + dy * dy;
    acc = acc + dz * dz;
    return acc;
}

void search(const Cloud& src, const Cloud& tgt, double radius, int m, std::vector<int>& idx, std::vector<int>& cnt)
{
    const float r2 = static_cast<float>(radius * radius);
    idx.assign(static_cast<size_t>(src.n) * m, -1);
    cnt.assign(static_cast<size_t>(src.n), 0);
    std::vector<std::pair<float, int>> cand;
    for (int64_t i = 0; i < src.n; ++i) {
        cand.clear();
        for (int64_t j = 0; j < tgt.n; ++j) {
            float d = d2f(&src.p[4 * i], &tgt.p[4 * j]);
            if (d < r2) cand.push_back({d, static_cast<int>(j)});
        }
        std::sort(cand.begin(), cand.end());
        int c = static_cast<int>(std::min<size_t>(cand.size(), static_cast<size_t>(m)));
        cnt[i] = c;
        for (int k = 0; k < c; ++k) idx[i * m + k] = cand[k].second;
    }
}

template <bool kFast>
void eval(const Cloud& src, const Cloud& tgt, const std::vector<int>& idx, const std::vector<int>& cnt, int m,
          const PairState& st, const WeightCfg& wc, double* S)
{
    for (int k = 0; k < kNSum; ++k) S[k] = 0.0;
    for (int64_t i = 0; i < src.n; ++i) {
        if (cnt[i] == 0) continue;
        const double sx = src.p[4 * i], sy = src.p[4 * i + 1], sz = src.p[4 * i + 2];
        double pe[3], pw[3];
        apply_pose(st.pose_e, sx, sy, sz, pe);
        apply_pose(st.pose_w, sx, sy, sz, pw);
        RowAcc row;
        row_begin(&row);
        for (int k = 0; k < cnt[i]; ++k) {
            const float* y = &tgt.p[4 * static_cast<size_t>(idx[i * m + k])];
            row_add<kFast>(&row, wc, y[0], y[1], y[2], pe, pw);
        }
        row_end(&row, sx, sy, sz, S);
    }
}

}  // namespace

extern "C" {

// Full align() with the product's controller.  Returns the number of outer iterations.
int emu_align(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt, int max_neighbours, double dof,
              double radius, int n_iter, double cost_drop_thresh, double n_cost_drop_it, const double* x0,
              double function_tolerance, int fast_weights, double* history, IterStats* stats, int max_hist,
              float* out_src)
{
    Cloud src{std::vector<float>(src_xyzw, src_xyzw + 4 * n_src), n_src};
    Cloud tgt{std::vector<float>(tgt_xyzw, tgt_xyzw + 4 * n_tgt), n_tgt};
    Config cfg{};
    for (int k = 0; k < kNP; ++k) cfg.x0[k] = x0[k];
    cfg.function_tolerance = function_tolerance;
    cfg.cost_drop_thresh = cost_drop_thresh;
    cfg.n_cost_drop_it = n_cost_drop_it;
    cfg.dof = dof;
    cfg.n_iter = n_iter;
    cfg.max_lm_iterations = 2147483647;
    cfg.is_normal = !(dof < 1.7976931348623157e308);
    cfg.fast_weights = fast_weights;
    const WeightCfg wc = make_weight_cfg(dof);
    PairState st;
    std::memset(&st, 0, sizeof(st));
    state_init(&st, &cfg);
    align_begin(&st, &cfg);
    const int m = static_cast<int>(std::min<int64_t>(max_neighbours, std::max<int64_t>(n_tgt, 1)));
    std::vector<int> idx, cnt;
    std::vector<double> hist(static_cast<size_t>(std::max(1, max_hist)) * 16);
    std::vector<IterStats> stv(static_cast<size_t>(std::max(1, max_hist)));
    double S[kNSum];
    long long guard = 0;
    while (st.phase != PH_DONE && guard++ < 10000000) {
        // one tick: search (if needed), eval, controller, transform -- the order k_* kernels run in
        st.apply_dT = 0;
        if (st.phase == PH_SEARCH) {
            search(src, tgt, radius, m, idx, cnt);
            int64_t K = 0;
            for (int c : cnt) K += c;
            st.K += K;
        }
        if (fast_weights) eval<true>(src, tgt, idx, cnt, m, st, wc, S); else eval<false>(src, tgt, idx, cnt, m, st, wc, S);
        controller_tick(&st, &cfg, S, hist.data(), stv.data(), max_hist);
        if (st.apply_dT) {
            for (int64_t i = 0; i < src.n; ++i) {
                const double x = src.p[4 * i], y = src.p[4 * i + 1], z = src.p[4 * i + 2];
                for (int r = 0; r < 3; ++r) {
                    const double* T = st.dT + 4 * r;
                    double acc = T[0] * x;
                    acc = acc + T[1] * y;
                    acc = acc + T[2] * z;
                    acc = acc + T[3];
                    src.p[4 * i + r] = static_cast<float>(acc);
                }
            }
        }
    }
    const int n = std::min(st.current_iteration, max_hist);
    if (history) std::memcpy(history, hist.data(), sizeof(double) * 16 * n);
    if (stats) std::memcpy(stats, stv.data(), sizeof(IterStats) * n);
    if (out_src) std::memcpy(out_src, src.p.data(), sizeof(float) * 4 * n_src);
    return st.current_iteration;
}

// One evaluation: the 24 moments and the expanded 7x7 system, for an explicit association.
void emu_normal_eq(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt, const int* idx,
                   const int* cnt, int m, double dof, const double* pose_w, const double* pose_e, int fast_weights,
                   double* normal_eq36, double* moments24)
{
    Cloud src{std::vector<float>(src_xyzw, src_xyzw + 4 * n_src), n_src};
    Cloud tgt{std::vector<float>(tgt_xyzw, tgt_xyzw + 4 * n_tgt), n_tgt};
    std::vector<int> vi(idx, idx + n_src * m), vc(cnt, cnt + n_src);
    PairState st;
    std::memset(&st, 0, sizeof(st));
    pose_from_x(pose_e, &st.pose_e);
    pose_from_x(pose_w, &st.pose_w);
    const WeightCfg wc = make_weight_cfg(dof);
    double S[kNSum];
    if (fast_weights) eval<true>(src, tgt, vi, vc, m, st, wc, S); else eval<false>(src, tgt, vi, vc, m, st, wc, S);
    double H[kNP * kNP], g[kNP], cost;
    expand_moments(S, pose_e, H, g, &cost);
    int o = 0;
    for (int r = 0; r < kNP; ++r)
        for (int c = r; c < kNP; ++c) normal_eq36[o++] = H[r * kNP + c];
    for (int r = 0; r < kNP; ++r) normal_eq36[o++] = g[r];
    normal_eq36[o] = cost;
    if (moments24) std::memcpy(moments24, S, sizeof(S));
}

}  // extern "C"
