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
#include "../../probabilistic_point_clouds_registration_b200/csrc/ppcr_tree.h"

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

void move_cloud(Cloud& src, const double* dT)
{
    for (int64_t i = 0; i < src.n; ++i) {
        const double x = src.p[4 * i], y = src.p[4 * i + 1], z = src.p[4 * i + 2];
        for (int r = 0; r < 3; ++r) {
            const double* T = dT + 4 * r;
            double acc = T[0] * x;
            acc = acc + T[1] * y;
            acc = acc + T[2] * z;
            acc = acc + T[3];
            src.p[4 * i + r] = static_cast<float>(acc);
        }
    }
}

template <bool kFast>
void eval(const Cloud& src, const Cloud& tgt, const std::vector<int>& idx, const std::vector<int>& cnt, int m,
          const PairState& st, const WeightCfg& wc, double* S)
{
    for (int k = 0; k < kNSum; ++k) S[k] = 0.0;
    bool same = true;
    for (int k = 0; k < 12; ++k)
        same = same && (reinterpret_cast<const double*>(&st.pose_e)[k] == reinterpret_cast<const double*>(&st.pose_w)[k]);
    for (int64_t i = 0; i < src.n; ++i) {
        if (cnt[i] == 0) continue;
        const double sx = src.p[4 * i], sy = src.p[4 * i + 1], sz = src.p[4 * i + 2];
        double pe[3], pw[3];
        apply_pose(st.pose_e, sx, sy, sz, pe);
        apply_pose(st.pose_w, sx, sy, sz, pw);
        if (kFast) {  // the float32 row path of k_eval<true>
            PointHL he;
            split_point(pe, &he);
            float dw[3];
            pose_delta(pe, pw, dw);
            RowAccF row;
            rowf_begin(&row);
            for (int k = 0; k < cnt[i]; ++k) {
                const float* y = &tgt.p[4 * static_cast<size_t>(idx[i * m + k])];
                rowf_add(&row, wc, y[0], y[1], y[2], he, dw, same);
            }
            rowf_end(&row, sx, sy, sz, S);
        } else {
            RowAcc row;
            row_begin(&row);
            for (int k = 0; k < cnt[i]; ++k) {
                const float* y = &tgt.p[4 * static_cast<size_t>(idx[i * m + k])];
                row_add<false>(&row, wc, y[0], y[1], y[2], pe, pw);
            }
            row_end(&row, sx, sy, sz, S);
        }
    }
}

}  // namespace

#if defined(PPCR_TREE_STATS)
static std::vector<float> g_query_cost;
static std::vector<int> g_query_opens, g_query_leaves;
#endif

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
    std::vector<double> hist(static_cast<size_t>(std::max(1, max_hist)) * 32);  // poses, then increments
    std::vector<IterStats> stv(static_cast<size_t>(std::max(1, max_hist)));
    double S[kNSum];
    long long guard = 0;
    while (st.phase != PH_DONE && guard++ < 10000000) {
        // one tick: [move the cloud + search] (if needed), eval, controller -- the order the kernels run in
        if (st.phase == PH_SEARCH) {
            if (st.apply_dT) move_cloud(src, st.dT);
            search(src, tgt, radius, m, idx, cnt);
            int64_t K = 0;
            for (int c : cnt) K += c;
            st.K += K;
        }
        if (fast_weights) eval<true>(src, tgt, idx, cnt, m, st, wc, S); else eval<false>(src, tgt, idx, cnt, m, st, wc, S);
        controller_tick(&st, &cfg, S, hist.data(), stv.data(), max_hist);
    }
    if (st.apply_dT) {  // epilogue of align(): the increment of the last outer iteration
        move_cloud(src, st.dT);
        st.apply_dT = 0;
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
    Expanded ev;
    expand_moments(S, pose_e, &ev);
    int o = 0;
    for (int r = 0; r < kNP; ++r)
        for (int c = r; c < kNP; ++c) normal_eq36[o++] = ev.H[r * kNP + c];
    for (int r = 0; r < kNP; ++r) normal_eq36[o++] = ev.g[r];
    normal_eq36[o] = ev.cost;
    if (moments24) std::memcpy(moments24, S, sizeof(S));
}


// The product's octree (csrc/ppcr_tree.h): serial build with the same split routine the build kernel calls, then the
// same traversal the search kernel runs, one query at a time.  Rows come back sorted ascending by (d2, index).
// list_kind: 0 = sorted register list, 1 = sorted addressable list, 2 = max-heap that starts full of infinity (the search
// kernel's list for large m), 3 = unordered column + worst scan, 4 / 5 = the heap's append / bottom-up fill modes,
// 100 + e = collect + select with a column of m + e slots, 200 + c = the three phases of the queued search kernel
// (k_search_q) with room for c candidates per query.
int64_t emu_tree_search(const float* src_xyzw, int64_t n_src, const float* tgt_xyzw, int64_t n_tgt, double radius,
                        int max_nn, int leaf_cap, int list_kind, const float* bounds, int* out_idx, float* out_d2,
                        int* out_cnt, int* out_n_nodes)
{
    const int m = static_cast<int>(std::min<int64_t>(max_nn, std::max<int64_t>(n_tgt, 1)));
    const float r2f = static_cast<float>(radius * radius);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < n_tgt; ++i)
        for (int a = 0; a < 3; ++a) {
            lo[a] = std::min(lo[a], tgt_xyzw[4 * i + a]);
            hi[a] = std::max(hi[a], tgt_xyzw[4 * i + a]);
        }
    if (n_tgt == 0) { lo[0] = lo[1] = lo[2] = 0.f; hi[0] = hi[1] = hi[2] = 1.f; }
    // make_tree_geom of ppcr_capi.cu
    TreeGeom g{};
    double span = 0.0, mag = 0.0;
    for (int k = 0; k < 3; ++k) {
        span = std::max(span, static_cast<double>(hi[k]) - lo[k]);
        mag = std::max({mag, std::fabs(static_cast<double>(lo[k])), std::fabs(static_cast<double>(hi[k]))});
    }
    span = std::max(span, 1e-6 * std::max(mag, 1e-30)) * (1.0 + 1e-5);
    if (!(span > 0.0)) span = 1.0;
    g.ox = lo[0]; g.oy = lo[1]; g.oz = lo[2];
    g.inv_hf = static_cast<float>(static_cast<double>(1 << kTreeBits) / span);
    g.hf = static_cast<float>(1.0 / static_cast<double>(g.inv_hf));
    g.slack = static_cast<float>(1e-6 * span + 1e-6 * (mag + span) + 1e-30);
    g.leaf_cap = leaf_cap;
    g.n_nodes_cap = static_cast<int>(64ll + 8ll * (2ll * n_tgt / std::max(leaf_cap, 1) + 8));
    std::vector<std::pair<unsigned long long, int>> kv(static_cast<size_t>(n_tgt));
    for (int64_t i = 0; i < n_tgt; ++i)
        kv[i] = {tree_key(g, tgt_xyzw[4 * i], tgt_xyzw[4 * i + 1], tgt_xyzw[4 * i + 2]), static_cast<int>(i)};
    std::stable_sort(kv.begin(), kv.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<unsigned long long> keys(static_cast<size_t>(n_tgt));
    std::vector<float4> pts(static_cast<size_t>(std::max<int64_t>(n_tgt, 1)));
    for (int64_t j = 0; j < n_tgt; ++j) {
        keys[j] = kv[j].first;
        const float* p = tgt_xyzw + 4 * static_cast<int64_t>(kv[j].second);
        pts[j].x = p[0]; pts[j].y = p[1]; pts[j].z = p[2];
        pts[j].w = bits_float(static_cast<uint32_t>(kv[j].second));
    }
    std::vector<TreeNode> nodes(static_cast<size_t>(g.n_nodes_cap) + 8);
    std::memset(nodes.data(), 0xff, nodes.size() * sizeof(TreeNode));
    TreeNode root;
    root.begin = 0; root.end = static_cast<int>(n_tgt); root.child = -1; root.mask = 0;
    tree_node_box(g, 0, 0ull, &root);
    nodes[0] = root;
    int n_nodes = 1, lb = 0, le = 1;
    for (int level = 0; level < kTreeBits; ++level) {
        for (int ni = lb; ni < le; ++ni) {
            if (nodes[ni].end - nodes[ni].begin > g.leaf_cap) {
                const int base = n_nodes;
                n_nodes += 8;
                if (base + 8 <= g.n_nodes_cap) tree_split_node(g, keys.data(), nodes.data(), ni, base);
            }
        }
        lb = le;
        le = std::min(n_nodes, g.n_nodes_cap);
    }
    for (int ni = 0; ni < std::min(n_nodes, g.n_nodes_cap); ++ni) tree_mark_leaf_children(nodes.data(), ni);
    for (int ni = 0; ni < std::min(n_nodes, g.n_nodes_cap); ++ni) tree_box_leaf(nodes.data(), pts.data(), ni);
    if (getenv("EMU_CHECK_LEAF_BOXES")) {
        // property behind the leaf boxes: for every boxed leaf and every query, the box bound never exceeds the float32
        // distance (dist2_exact) of any point stored in the leaf -- so pruning on it cannot lose a neighbour
        // walk the tree from the root so that leaves are identified by their parents' masks
        std::vector<int> todo{0};
        if (nodes[0].child >= 0)
            while (!todo.empty()) {
                const int ni = todo.back();
                todo.pop_back();
                const TreeNode n = nodes[static_cast<size_t>(ni)];
                for (int c = 0; c < 8; ++c) {
                    if (!((n.mask >> c) & 1)) continue;
                    const int ch = n.child + c;
                    if (!((n.mask >> (16 + c)) & 1)) {
                        todo.push_back(ch);
                        continue;
                    }
                    const TreeNode leaf = nodes[static_cast<size_t>(ch)];
                    for (int64_t i = 0; i < n_src; ++i) {
                        const float* q = src_xyzw + 4 * i;
                        const float lb = leaf_box_lower_bound(leaf, q[0], q[1], q[2]);
                        for (int j = leaf.begin; j < leaf.end; ++j)
                            if (lb > dist2_exact(q[0], q[1], q[2], pts[static_cast<size_t>(j)].x, pts[static_cast<size_t>(j)].y, pts[static_cast<size_t>(j)].z)) return -2;
                    }
                }
            }
    }
    if (out_n_nodes) *out_n_nodes = n_nodes;
    int64_t total = 0;
    int stack[2 * kTreeStack];
    std::vector<unsigned long long> buf(static_cast<size_t>(heap_slots(std::max(m, 1))));
#if defined(PPCR_TREE_STATS)
    g_query_cost.assign(static_cast<size_t>(n_src), 0.f);
    g_query_opens.assign(static_cast<size_t>(n_src), 0);
    g_query_leaves.assign(static_cast<size_t>(n_src), 0);
#endif
    for (int64_t i = 0; i < n_src; ++i) {
        const float* q = src_xyzw + 4 * i;
        const float bound0 = bounds ? bounds[i] : r2f;
        std::vector<unsigned long long> found;
#if defined(PPCR_TREE_STATS)
        const TreeStats before = g_tree_stats;
        struct CostNote {
            const TreeStats& b;
            float& out;
            int& opens;
            int& leaves;
            ~CostNote()
            {
                opens = static_cast<int>(g_tree_stats.opens - b.opens);
                leaves = static_cast<int>(g_tree_stats.leaves - b.leaves);
                // rough thread-instruction weights of the search kernel (profiles/): open, scanned point, survivor, insertion
                out = 120.f * (g_tree_stats.opens - b.opens) + 14.f * (g_tree_stats.points - b.points) +
                      25.f * (g_tree_stats.survivors - b.survivors) + 65.f * (g_tree_stats.inserts - b.inserts) + 300.f;
            }
        } note{before, g_query_cost[static_cast<size_t>(i)], g_query_opens[static_cast<size_t>(i)], g_query_leaves[static_cast<size_t>(i)]};
#endif
        auto take = [&](const unsigned long long* k, int cap) {
            for (int s2 = 0; s2 < cap; ++s2) {
                const int e = m - 1 - s2;
                if (e >= 0 && k[s2] != kKeyInf) found.push_back(k[s2]);
            }
        };
        const int cap = m <= 4 ? 4 : m <= 8 ? 8 : m <= 12 ? 12 : m <= 16 ? 16 : m <= 20 ? 20 : m <= 24 ? 24 : m <= 32 ? 32 : 0;
        if (list_kind == 0 && cap > 0) {
#define EMU_RUN(C)                                                                                  \
    {                                                                                               \
        TopList<C> L;                                                                               \
        L.init(m);                                                                                  \
        tree_search(g, nodes.data(), pts.data(), q[0], q[1], q[2], r2f, bound0, L, stack);                  \
        take(L.k, C);                                                                               \
    }
            switch (cap) {
                case 4: EMU_RUN(4) break;
                case 8: EMU_RUN(8) break;
                case 12: EMU_RUN(12) break;
                case 16: EMU_RUN(16) break;
                case 20: EMU_RUN(20) break;
                case 24: EMU_RUN(24) break;
                default: EMU_RUN(32) break;
            }
#undef EMU_RUN
        } else if (list_kind == 2) {  // the search kernel's list: a max-heap (one column per thread on the device)
            HeapList<1> L;
            L.k = buf.data();
            L.init(m);
            tree_search(g, nodes.data(), pts.data(), q[0], q[1], q[2], r2f, bound0, L, stack);
            for (int s2 = L.begin(); s2 < L.end(); ++s2)
                if (buf[s2] != kKeyInf) found.push_back(buf[s2]);
        } else if (list_kind >= 200) {
            // the queued search kernel's three phases, one query at a time: leaves within the (fixed) bound -> candidate
            // positions -> the m best.  list_kind - 200 = candidate capacity; a query that overflows it (or has no finite
            // bound... the kernel falls back to tree_search for those) is searched with the heap, like the kernel does.
            const int qcap = list_kind - 200;
            std::vector<int> leaves;
            std::vector<uint32_t> cand;
            const float b0 = bound0 < r2f ? bound0 : r2f;
            auto emit = [&](int first, int mask) {
                for (int c = 0; c < 8; ++c)
                    if ((mask >> c) & 1) leaves.push_back(first + c);
                return true;
            };
            tree_collect_leaves(g, nodes.data(), q[0], q[1], q[2], b0, emit, stack);
            auto push = [&](int j0, uint32_t pass) {
                for (; pass; pass &= pass - 1) cand.push_back(static_cast<uint32_t>(j0 + lowest_bit(pass)));
            };
            for (int node : leaves) leaf_candidates(nodes.data(), pts.data(), node, q[0], q[1], q[2], candidate_limit(b0, r2f), push);
            if (static_cast<int>(cand.size()) <= qcap) {
                unsigned long long kth = 0;
                auto at = [&](int c) { return static_cast<int>(cand[static_cast<size_t>(c)]); };
                const int cnt = select_candidates<1>(pts.data(), at, static_cast<int>(cand.size()), m, q[0], q[1], q[2], buf.data(), &kth);
                for (int s2 = 0; s2 < cnt; ++s2) found.push_back(buf[s2]);
                if (!std::is_sorted(found.begin(), found.end())) return -2;  // rows come out in FLANN's order
                if ((kth != kKeyInf) != (cnt == m) || (cnt == m && kth != *std::max_element(found.begin(), found.end())))
                    return -1;
            } else {
                // the kernel's overflow path: the qcap candidates the list holds are distinct targets within the radius, so the
                // m-th smallest of their distances bounds the true one; the heap walk then starts from that (tight) bound
                float bound1 = bound0;
                if (qcap >= m) {
                    unsigned long long kth = 0;
                    auto at = [&](int c) { return static_cast<int>(cand[static_cast<size_t>(c)]); };
                    select_candidates<1>(pts.data(), at, qcap, m, q[0], q[1], q[2], buf.data(), &kth);
                    if (kth != kKeyInf && key_d2(kth) < bound1) bound1 = key_d2(kth);
                }
                HeapList<1> L;
                L.k = buf.data();
                L.init(m);
                tree_search(g, nodes.data(), pts.data(), q[0], q[1], q[2], r2f, bound1, L, stack);
                for (int s2 = L.begin(); s2 < L.end(); ++s2)
                    if (buf[s2] != kKeyInf) found.push_back(buf[s2]);
            }
        } else if (list_kind >= 100) {  // collect + select, the search kernel's list; column of m + (list_kind - 100) slots
            const int cap = m + (list_kind - 100);
            std::vector<unsigned long long> col(static_cast<size_t>(cap));
            CollectList<1> L;
            L.k = col.data();
            L.init(m, cap);
            tree_search(g, nodes.data(), pts.data(), q[0], q[1], q[2], r2f, bound0, L, stack);
            L.finish();
            for (int s2 = L.begin(); s2 < L.end(); ++s2) found.push_back(col[s2]);
            if (L.kth_key() != kKeyInf && (static_cast<int>(found.size()) != m ||
                                           L.kth_key() != *std::max_element(found.begin(), found.end())))
                return -1;  // the warm-start distance the kernel stores must be the m-th best
        } else if (list_kind == 4 || list_kind == 5) {  // the heap's other fill modes (tuning variants)
            auto run = [&](auto& L) {
                L.k = buf.data();
                L.init(m);
                tree_search(g, nodes.data(), pts.data(), q[0], q[1], q[2], r2f, bound0, L, stack);
                for (int s2 = L.begin(); s2 < L.end(); ++s2)
                    if (buf[s2] != kKeyInf) found.push_back(buf[s2]);
            };
            if (list_kind == 4) {
                HeapList<1, 1> L;
                run(L);
            } else {
                HeapList<1, 2> L;
                run(L);
            }
        } else if (list_kind == 3) {  // the search kernel's list for small m: unordered column + worst scan
            ScanList<1> L;
            L.k = buf.data();
            L.init(m);
            tree_search(g, nodes.data(), pts.data(), q[0], q[1], q[2], r2f, bound0, L, stack);
            for (int s2 = 0; s2 < L.n; ++s2) found.push_back(buf[s2]);
        } else {
            TopListDyn L;
            L.k = buf.data();
            L.init(m);
            tree_search(g, nodes.data(), pts.data(), q[0], q[1], q[2], r2f, bound0, L, stack);
            take(buf.data(), m);
        }
        std::sort(found.begin(), found.end());
        out_cnt[i] = static_cast<int>(found.size());
        for (size_t k = 0; k < found.size(); ++k) {
            out_idx[i * max_nn + static_cast<int64_t>(k)] = key_index(found[k]);
            out_d2[i * max_nn + static_cast<int64_t>(k)] = key_d2(found[k]);
        }
        total += static_cast<int64_t>(found.size());
    }
    return total;
}

#if defined(PPCR_TREE_STATS)
void emu_tree_costs(float* out, long long n)
{
    for (long long i = 0; i < n && i < static_cast<long long>(g_query_cost.size()); ++i) out[i] = g_query_cost[static_cast<size_t>(i)];
}
void emu_tree_per_query(int* opens, int* leaves, long long n)
{
    for (long long i = 0; i < n && i < static_cast<long long>(g_query_opens.size()); ++i) {
        opens[i] = g_query_opens[static_cast<size_t>(i)];
        leaves[i] = g_query_leaves[static_cast<size_t>(i)];
    }
}
void emu_tree_stats(long long* out7, int reset)
{
    const long long v[7] = {g_tree_stats.opens, g_tree_stats.leaves, g_tree_stats.leaves_skipped, g_tree_stats.points,
                            g_tree_stats.survivors, g_tree_stats.inserts, g_tree_stats.stack_skipped};
    for (int k = 0; k < 7; ++k) out7[k] = v[k];
    if (reset) g_tree_stats = TreeStats{};
}
#endif

}  // extern "C"
