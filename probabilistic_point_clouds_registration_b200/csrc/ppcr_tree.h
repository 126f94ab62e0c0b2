// ppcr_tree.h -- the neighbour-search structure: a linear octree over the Morton-sorted target cloud, and the
// per-query traversal with a register-resident top-m list.
//
// Replaces pcl::KdTreeFLANN::setInputCloud + radiusSearch (src/prob_point_cloud_registration.cc:66-67,72-81).
// Semantics reproduced exactly: d2 = ((dx*dx) + dy*dy) + dz*dz in float32 without contraction,
// membership d2 < float(radius*radius) strictly, at most m results = the m smallest under (d2, index), returned
// ascending.  The result is a pure function of the point sets, so the tree shape / traversal order cannot change it
// as long as pruning is conservative.
//
// Why a tree and not a uniform grid: LiDAR-like clouds span three orders of magnitude in density.  A cell edge that
// keeps dense regions cheap makes sparse queries walk tens of thousands of empty cells, and one that suits sparse
// regions makes dense queries test thousands of candidates.  The octree adapts: every query opens O(depth) nodes
// and scans a handful of leaves of at most `leaf_cap` points.
//
// Layout: the target is sorted by the 3*kTreeBits-bit Morton key of its quantised coordinates, so every octree
// node is one contiguous range [begin,end) of tgt_sorted; nodes are 32-byte records, the 8 children of a node are
// consecutive.  Plain C++ qualified PPCR_HD: the search kernel and the CPU tests (tests/emu) share this source.
#ifndef PPCR_TREE_H
#define PPCR_TREE_H

#include <stdint.h>

#include "ppcr_lm.h"  // PPCR_HD

#if !defined(__CUDACC__)
struct float4 {  // host-only builds (tests/emu): the 16-byte pcl::PointXYZ record
    float x, y, z, w;
};
#endif

namespace ppcr {

constexpr int kTreeBits = 16;            // quantisation bits per axis -> 48-bit Morton keys
constexpr int kTreeStack = 7 * kTreeBits + 9;
constexpr unsigned long long kKeyInf = 0xffffffffffffffffull;

struct TreeGeom {
    float ox, oy, oz;   // min corner of the root cube
    float inv_hf;       // 1 / finest cell edge
    float hf;           // finest cell edge
    float slack;        // absolute bound on the float fuzz of binning + node centres; inflates every box
    int n_nodes_cap;
    int leaf_cap;       // a node holding more points than this is split (unless it is at the finest level)
};

struct TreeNode {       // 32 bytes
    float cx, cy, cz;   // cube centre
    float half;         // half edge (not inflated)
    int begin, end;     // range in the Morton-sorted target
    int child;          // first of 8 consecutive children, -1 = leaf
    int mask;           // bits 0..7: non-empty children; bits 8..15: level; bits 16..23: children that are leaves
};

// ---- bit-exact float helpers ---------------------------------------------------------------------------------

PPCR_HD float f_sub(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
}
PPCR_HD float f_mul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}
PPCR_HD float f_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}

// FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz in float32 with NO fused multiply-add
PPCR_HD float dist2_exact(float qx, float qy, float qz, float px, float py, float pz)
{
    const float dx = f_sub(qx, px), dy = f_sub(qy, py), dz = f_sub(qz, pz);
    float acc = f_mul(dx, dx);
    acc = f_add(acc, f_mul(dy, dy));
    acc = f_add(acc, f_mul(dz, dz));
    return acc;
}

PPCR_HD uint32_t float_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c;
    c.f = f;
    return c.u;
#endif
}
PPCR_HD float bits_float(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c;
    c.u = u;
    return c.f;
#endif
}

// (d2, index) packed so that integer order = lexicographic order; +1 keeps every real key above the 0 that pins
// the unused slots of a list
PPCR_HD unsigned long long make_key(float d2, int idx)
{
    return ((static_cast<unsigned long long>(float_bits(d2)) << 32) | static_cast<uint32_t>(idx)) + 1ull;
}
PPCR_HD float key_d2(unsigned long long k) { return bits_float(static_cast<uint32_t>((k - 1ull) >> 32)); }
PPCR_HD int key_index(unsigned long long k) { return static_cast<int>(static_cast<uint32_t>(k - 1ull)); }

PPCR_HD int lowest_bit(uint32_t v)
{
#if defined(__CUDA_ARCH__)
    return __ffs(static_cast<int>(v)) - 1;
#else
    return __builtin_ctz(v);
#endif
}

// ---- Morton keys ---------------------------------------------------------------------------------------------

PPCR_HD unsigned long long spread3(uint32_t v)  // 16 bits -> every third bit
{
    unsigned long long x = v & 0xffffull;
    x = (x | (x << 16)) & 0x0000ff0000ffull;
    x = (x | (x << 8)) & 0x00f00f00f00full;
    x = (x | (x << 4)) & 0x0c30c30c30c3ull;
    x = (x | (x << 2)) & 0x249249249249ull;
    return x;
}

PPCR_HD uint32_t compact3(unsigned long long x)
{
    x &= 0x249249249249ull;
    x = (x | (x >> 2)) & 0x0c30c30c30c3ull;
    x = (x | (x >> 4)) & 0x00f00f00f00full;
    x = (x | (x >> 8)) & 0x0000ff0000ffull;
    x = (x | (x >> 16)) & 0xffffull;
    return static_cast<uint32_t>(x);
}

PPCR_HD int tree_coord(float v, float origin, float inv_hf)
{
    // monotone in v; clamped so that queries and fuzz at the faces stay inside the root cube
    float t = f_mul(f_sub(v, origin), inv_hf);
    if (!(t > 0.f)) t = 0.f;
    const float top = static_cast<float>((1 << kTreeBits) - 1);
    if (t > top) t = top;
    return static_cast<int>(t);
}

PPCR_HD unsigned long long tree_key(const TreeGeom& g, float x, float y, float z)
{
    const uint32_t ix = static_cast<uint32_t>(tree_coord(x, g.ox, g.inv_hf));
    const uint32_t iy = static_cast<uint32_t>(tree_coord(y, g.oy, g.inv_hf));
    const uint32_t iz = static_cast<uint32_t>(tree_coord(z, g.oz, g.inv_hf));
    return spread3(ix) | (spread3(iy) << 1) | (spread3(iz) << 2);  // bit 0 = x, bit 1 = y, bit 2 = z of each digit
}

// node geometry from its level and the Morton prefix of any key inside it
PPCR_HD void tree_node_box(const TreeGeom& g, int level, unsigned long long key_in_node, TreeNode* n)
{
    const int shift = 3 * (kTreeBits - level);
    const unsigned long long prefix = shift >= 48 ? 0ull : (key_in_node >> shift);
    const uint32_t ix = compact3(prefix), iy = compact3(prefix >> 1), iz = compact3(prefix >> 2);
    const float half = g.hf * static_cast<float>(1u << (kTreeBits - level)) * 0.5f;  // exact power-of-two scaling
    n->cx = g.ox + static_cast<float>(2u * ix + 1u) * half;
    n->cy = g.oy + static_cast<float>(2u * iy + 1u) * half;
    n->cz = g.oz + static_cast<float>(2u * iz + 1u) * half;
    n->half = half;
}

// first position in keys[lo,hi) whose octal digit at `shift` is >= digit (keys of one node: the digit is monotone)
PPCR_HD int tree_digit_lower_bound(const unsigned long long* keys, int lo, int hi, int shift, unsigned digit)
{
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (static_cast<unsigned>((keys[mid] >> shift) & 7ull) < digit) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

// Splits node `ni` (level < kTreeBits, more than leaf_cap points) into 8 children written at nodes[base..base+8).
PPCR_HD void tree_split_node(const TreeGeom& g, const unsigned long long* keys, TreeNode* nodes, int ni, int base)
{
    TreeNode parent = nodes[ni];
    const int level = (parent.mask >> 8) & 0xff;
    const int shift = 3 * (kTreeBits - level - 1);
    int bound[9];
    bound[0] = parent.begin;
    bound[8] = parent.end;
    for (unsigned c = 1; c < 8; ++c) bound[c] = tree_digit_lower_bound(keys, bound[c - 1], parent.end, shift, c);
    int mask = 0;
    const unsigned long long prefix = keys[parent.begin] >> (shift + 3);
    for (int c = 0; c < 8; ++c) {
        TreeNode ch;
        ch.begin = bound[c];
        ch.end = bound[c + 1];
        ch.child = -1;
        ch.mask = (level + 1) << 8;
        tree_node_box(g, level + 1, ((prefix << 3) | static_cast<unsigned long long>(c)) << shift, &ch);
        if (ch.end > ch.begin) mask |= 1 << c;
        nodes[base + c] = ch;
    }
    parent.child = base;
    parent.mask = (level << 8) | mask;
    nodes[ni] = parent;
}

// ---- loads ---------------------------------------------------------------------------------------------------
//
// On the device every node / point read is an explicit 128-bit read-only load: the pointers come out of a structure in
// global memory, so without this the compiler emits generic 32-bit loads (six per node, two per point).

PPCR_HD float4 load_point(const float4* p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

PPCR_HD TreeNode load_node(const TreeNode* n)
{
#if defined(__CUDA_ARCH__)
    const int4 a = __ldg(reinterpret_cast<const int4*>(n));
    const int4 b = __ldg(reinterpret_cast<const int4*>(n) + 1);
    TreeNode t;
    t.cx = __int_as_float(a.x);
    t.cy = __int_as_float(a.y);
    t.cz = __int_as_float(a.z);
    t.half = __int_as_float(a.w);
    t.begin = b.x;
    t.end = b.y;
    t.child = b.z;
    t.mask = b.w;
    return t;
#else
    return *n;
#endif
}

// ---- top-m list ----------------------------------------------------------------------------------------------
//
// Kept DESCENDING: k[0] is the current worst of the m best, k[m-1] the best; slots >= m are pinned to 0 (below
// every real key) so the fully unrolled insertion needs no dynamic register indexing.  Empty slots hold kKeyInf.

template <int CAP>
struct TopList {
    unsigned long long k[CAP];

    PPCR_HD void init(int m)
    {
#pragma unroll
        for (int i = 0; i < CAP; ++i) k[i] = i < m ? kKeyInf : 0ull;
    }
    PPCR_HD unsigned long long worst() const { return k[0]; }
    // pre: x < k[0]
    PPCR_HD void insert(unsigned long long x)
    {
#pragma unroll
        for (int i = 0; i < CAP - 1; ++i) k[i] = (k[i + 1] > x) ? k[i + 1] : (k[i] > x ? x : k[i]);
        k[CAP - 1] = k[CAP - 1] > x ? x : k[CAP - 1];
    }
};

// any m: the same list in addressable (local) memory, loops not unrolled
struct TopListDyn {
    unsigned long long* k;
    int m;
    PPCR_HD void init(int m_)
    {
        m = m_;
        for (int i = 0; i < m; ++i) k[i] = kKeyInf;
    }
    PPCR_HD unsigned long long worst() const { return k[0]; }
    PPCR_HD void insert(unsigned long long x)
    {
        int i = 0;
        for (; i < m - 1 && k[i + 1] > x; ++i) k[i] = k[i + 1];
        k[i] = x;
    }
};

// After the build: every inner node learns which of its children are leaves (bits 16..23 of mask), so that the
// traversal can sort children into "open later" and "scan later" without loading them.
PPCR_HD void tree_mark_leaf_children(TreeNode* nodes, int ni)
{
    const int child = nodes[ni].child;
    if (child < 0) return;
    int bits = 0;
    for (int c = 0; c < 8; ++c)
        if (nodes[child + c].child < 0) bits |= 1 << (16 + c);
    nodes[ni].mask = (nodes[ni].mask & 0xffff) | bits;
}

// ---- tight leaf boxes ------------------------------------------------------------------------------------------
//
// The octree's cells are cubes; the points of a scan lie on surfaces.  A search ball that cuts a leaf's cube usually misses
// the few points inside it: of the leaves a query of the 1M-point pair reaches, a third hold no point within its bound, and of
// the leaves reached by a query that still has fewer than m neighbours (bound = the full radius: 8 % of the rows while the
// pose is off) seven in eight (tools/tree_stats_real.py).  Once the tree is built nobody reads the cube of a LEAF any more --
// its parent computes the children's lower bounds from its own centre -- so every leaf but a root leaf stores the bounding
// box of its points in those fields: (cx, cy, cz) = min corner, (half, child, mask) = max corner (the last two as float bits).
// One 32-byte node load then answers "can this leaf hold a point within the bound" exactly: the box distance is computed with
// the operations of dist2_exact in the same order, and rounding is monotone, so it never exceeds the distance of a point inside.
// Must run after tree_mark_leaf_children has looked at every node (it recognises a leaf by child < 0, which the box overwrites).
// A thread only touches its own node.
PPCR_HD void tree_box_leaf(TreeNode* nodes, const float4* __restrict__ pts, int ni)
{
    TreeNode leaf = nodes[ni];
    if (ni == 0 || leaf.child >= 0 || leaf.end <= leaf.begin) return;  // the root, inner nodes, empty and unused slots stay as they are
    float lo[3] = {pts[leaf.begin].x, pts[leaf.begin].y, pts[leaf.begin].z};
    float hi[3] = {lo[0], lo[1], lo[2]};
    for (int j = leaf.begin + 1; j < leaf.end; ++j) {
        const float4 p = pts[j];
        lo[0] = p.x < lo[0] ? p.x : lo[0];
        lo[1] = p.y < lo[1] ? p.y : lo[1];
        lo[2] = p.z < lo[2] ? p.z : lo[2];
        hi[0] = p.x > hi[0] ? p.x : hi[0];
        hi[1] = p.y > hi[1] ? p.y : hi[1];
        hi[2] = p.z > hi[2] ? p.z : hi[2];
    }
    leaf.cx = lo[0];
    leaf.cy = lo[1];
    leaf.cz = lo[2];
    leaf.half = hi[0];
    leaf.child = static_cast<int>(float_bits(hi[1]));
    leaf.mask = static_cast<int>(float_bits(hi[2]));
    nodes[ni] = leaf;
}

// lower bound of dist2_exact(q, p) over the points p of a boxed leaf (never above any of them)
PPCR_HD float leaf_box_lower_bound(const TreeNode& leaf, float qx, float qy, float qz)
{
    const float hy = bits_float(static_cast<uint32_t>(leaf.child)), hz = bits_float(static_cast<uint32_t>(leaf.mask));
    float gx = f_sub(leaf.cx, qx), gy = f_sub(leaf.cy, qy), gz = f_sub(leaf.cz, qz);
    const float ux = f_sub(qx, leaf.half), uy = f_sub(qy, hy), uz = f_sub(qz, hz);
    gx = gx > ux ? gx : ux;
    gy = gy > uy ? gy : uy;
    gz = gz > uz ? gz : uz;
    gx = gx > 0.f ? gx : 0.f;
    gy = gy > 0.f ? gy : 0.f;
    gz = gz > 0.f ? gz : 0.f;
    float acc = f_mul(gx, gx);
    acc = f_add(acc, f_mul(gy, gy));
    acc = f_add(acc, f_mul(gz, gz));
    return acc;
}

// binary max-heap of the m best keys in addressable memory, element i at k[i * STRIDE] (the search kernel keeps one
// column per thread in shared memory, STRIDE = block size).  Once m candidates are known a better one replaces the
// root and sifts down.  How the first m get in is what FILL selects; measured on the 1M-point pair (profiles/):
//   0  the column starts full of kKeyInf and every candidate replaces the root.  Most instructions, but ONE code
//      path walked from the same slot by every thread: the threads of a warp that insert at the same time stay in
//      step (13.2 of 32 lanes active over the kernel).  The fastest of the three.
//   1  the first m are appended, the heap is built when the m-th arrives: three code paths the warp serialises on.
//   2  bottom-up: the n-th candidate goes to slot m-1-n and sifts down inside the part already filled (Floyd's
//      construction one element at a time).  13 % fewer thread instructions than 0 but every thread starts from
//      another slot: 8.1 lanes active, 40 % MORE warp instructions, 0.62 ms against 0.43.
// Valid entries are the slots [begin(), end()) that do not hold kKeyInf.
PPCR_HD constexpr int heap_slots(int m) { return m < 1 ? 1 : m; }

template <int STRIDE, int FILL = 0>
struct HeapList {
    unsigned long long* k;
    int m;
    int n;
    PPCR_HD void init(int m_, int /*cap*/ = 0)
    {
        m = m_;
        n = 0;
        if (FILL == 0) {
            n = m;
            for (int i = 0; i < m; ++i) k[i * STRIDE] = kKeyInf;
        }
    }
    PPCR_HD unsigned long long worst() const { return n == m ? k[0] : kKeyInf; }
    // puts x into the hole at i and moves it down until both children are smaller
    PPCR_HD void sift_down(int i, unsigned long long x)
    {
        for (;;) {
            const int l = 2 * i + 1;
            if (l >= m) break;
            unsigned long long vc = k[l * STRIDE];
            int c = l;
            if (l + 1 < m) {
                const unsigned long long vr = k[(l + 1) * STRIDE];
                if (vr > vc) {
                    vc = vr;
                    c = l + 1;
                }
            }
            if (vc <= x) break;
            k[i * STRIDE] = vc;
            i = c;
        }
        k[i * STRIDE] = x;
    }
    // pre: x < worst()
    PPCR_HD void insert(unsigned long long x)
    {
        if (FILL == 2) {
            int at = 0;
            if (n < m) at = m - 1 - n++;
            sift_down(at, x);
        } else if (n < m) {
            k[n * STRIDE] = x;
            if (++n == m)
                for (int i = m / 2 - 1; i >= 0; --i) sift_down(i, k[i * STRIDE]);
        } else {
            sift_down(0, x);
        }
    }
    PPCR_HD void finish() {}
    // Leaves the column (FILL == 0) in ascending (distance, index) order, unused slots (kKeyInf) last.  The layout of a heap
    // depends on the order its keys arrived in -- fixed for a given query, tree and starting bound, which is all k_search and
    // the heavy chunks of k_search_q need for a reproducible row (the order of a row's neighbours decides the last bits of its
    // float32 sums) -- but the one-by-one fallback of k_search_q starts from a bound that depends on which candidates other
    // threads pushed first, so it sorts.  (Sorting everywhere was measured: the 10M-point pair 375 -> 467 ms, the sift paths of
    // the 32 lanes of a warp all differ.)  Heap sort in place; call kth_key() BEFORE sort().
    PPCR_HD void sort()
    {
        if (FILL != 0) return;
        const int full = m;
        for (int e = full - 1; e > 0; --e) {
            const unsigned long long top = k[0], x = k[e * STRIDE];
            m = e;
            sift_down(0, x);
            k[e * STRIDE] = top;
        }
        m = full;
    }
    PPCR_HD int begin() const { return FILL == 2 ? m - n : 0; }
    PPCR_HD int end() const { return FILL == 2 ? m : n; }
    // key of the m-th best, kKeyInf when fewer than m real keys are held (before sort())
    PPCR_HD unsigned long long kth_key() const { return n == m ? k[0] : kKeyInf; }
};

// The m best keys as an UNORDERED column plus the position and key of the current worst in registers.  The first m
// candidates are appended (one store each); after that a better candidate overwrites the worst and the column is
// scanned once for the new worst.  That scan has the same trip count (m) in every thread of a warp, so -- unlike
// the sift-down of the heap, whose length differs from thread to thread -- the threads that insert at the same time
// stay together; and worst() costs nothing.  m loads + compares per replacement: the cheaper list up to m of a few
// dozen, which is where the reference's defaults live (max_neighbours = 20, the benchmark's 10); the heap takes over
// for large m.  Entries [0, n) are valid and come out unordered.
template <int STRIDE>
struct ScanList {
    unsigned long long* k;
    unsigned long long w;  // key of the worst of the m best; kKeyInf while fewer than m are known
    int m;
    int n;
    int wi;                // its position, valid when n == m
    PPCR_HD void init(int m_, int /*cap*/ = 0)
    {
        m = m_;
        n = 0;
        wi = 0;
        w = kKeyInf;
    }
    PPCR_HD unsigned long long worst() const { return w; }
    PPCR_HD void rescan()
    {
        unsigned long long best = 0ull;
        int at = 0;
#if defined(PPCR_SCAN_UNROLL)
        constexpr int kUnroll = PPCR_SCAN_UNROLL;
#pragma unroll kUnroll
#endif
        for (int i = 0; i < m; ++i) {
            const unsigned long long v = k[i * STRIDE];
            if (v > best) {
                best = v;
                at = i;
            }
        }
        w = best;
        wi = at;
    }
    // pre: x < worst()
    PPCR_HD void insert(unsigned long long x)
    {
        if (n < m) {
            k[n * STRIDE] = x;
            if (++n == m) rescan();
        } else {
            k[wi * STRIDE] = x;
            rescan();
        }
    }
    PPCR_HD void finish() {}
    PPCR_HD int begin() const { return 0; }
    PPCR_HD int end() const { return n; }
    PPCR_HD unsigned long long kth_key() const { return w; }
};

// Collect first, select later (tuning variant 16): only APPENDS while the walk runs and picks the m best when the walk
// is over -- Floyd's heap construction over the first m entries, then the rest streamed through the root -- relying on
// the warm pruning bound being good from the start (15-20 targets lie within it at m = 10).  When the column fills up
// (cap entries) the selection runs early, the m best stay in slots [0, m) as a heap, the bound drops to their worst and
// collection goes on behind them.  12 % fewer thread instructions than the heap, and slower (0.63 ms against 0.42 on
// the converged 1M-point pair, 4.5 ms against 0.97 for a search without a bound): the wider column costs resident
// blocks, and without the heap updates the threads of a warp drift apart in the leaf scans (7-8 lanes active there
// instead of 15).  cap > m.  After finish() the valid entries are [0, n), unordered.
template <int STRIDE>
struct CollectList {
    unsigned long long* k;
    int m;
    int cap;
    int n;
    bool heaped;  // slots [0, m) hold the m best seen so far as a max-heap; [m, n) are pending
    PPCR_HD void init(int m_, int cap_)
    {
        m = m_;
        cap = cap_;
        n = 0;
        heaped = false;
    }
    PPCR_HD unsigned long long worst() const { return heaped ? k[0] : kKeyInf; }
    PPCR_HD void sift_down(int i, unsigned long long x)
    {
        for (;;) {
            const int l = 2 * i + 1;
            if (l >= m) break;
            unsigned long long vc = k[l * STRIDE];
            int c = l;
            if (l + 1 < m) {
                const unsigned long long vr = k[(l + 1) * STRIDE];
                if (vr > vc) {
                    vc = vr;
                    c = l + 1;
                }
            }
            if (vc <= x) break;
            k[i * STRIDE] = vc;
            i = c;
        }
        k[i * STRIDE] = x;
    }
    // pre: n >= m.  Leaves the m best of the n entries in [0, m) as a heap.
    PPCR_HD void select()
    {
        if (!heaped) {
            for (int i = m / 2 - 1; i >= 0; --i) sift_down(i, k[i * STRIDE]);
            heaped = true;
        }
        for (int j = m; j < n; ++j) {
            const unsigned long long x = k[j * STRIDE];
            if (x < k[0]) sift_down(0, x);
        }
        n = m;
    }
    // pre: x < worst()
    PPCR_HD void insert(unsigned long long x)
    {
        k[n * STRIDE] = x;
        if (++n == cap) select();
    }
    PPCR_HD void finish()
    {
        if (n > m || (n == m && !heaped)) select();
    }
    PPCR_HD int begin() const { return 0; }
    PPCR_HD int end() const { return n; }
    // after finish(): key of the m-th best, kKeyInf when fewer than m were found
    PPCR_HD unsigned long long kth_key() const { return (n == m && heaped) ? k[0] : kKeyInf; }
};

// ---- traversal -----------------------------------------------------------------------------------------------

// work counters of the traversal, host builds with -DPPCR_TREE_STATS only (tools/tree_stats.py)
#if defined(PPCR_TREE_STATS) && !defined(__CUDA_ARCH__)
struct TreeStats {
    long long opens, leaves, leaves_skipped, points, survivors, inserts, stack_skipped;
};
inline TreeStats g_tree_stats = {};
#define PPCR_STAT(field, n) (g_tree_stats.field += (n))
#else
#define PPCR_STAT(field, n) ((void)0)
#endif

// squared distance from q to the interval [c - hi, c + hi] along one axis
PPCR_HD float axis_gap2(float q, float c, float hi)
{
    float d = fabsf(q - c) - hi;
    d = d > 0.f ? d : 0.f;
    return d * d;
}

// The pruning threshold a (not shaved) box lower bound is compared with: slightly ABOVE the bound, so that float
// rounding in the lower bound can only keep a subtree that exact arithmetic would prune, never the reverse.
PPCR_HD float prune_threshold(float bound_d2) { return bound_d2 * 1.00002f; }

// the leaf children `first + c` (c a set bit of mask) that can hold a point within bound_d2 of q
PPCR_HD int tree_prune_leaf_mask(const TreeNode* __restrict__ nodes, int first, int mask, float qx, float qy, float qz, float bound_d2)
{
    int keep = 0;
    for (int mk = mask; mk; mk &= mk - 1) {
        const int c = lowest_bit(static_cast<uint32_t>(mk));
        const TreeNode leaf = load_node(nodes + first + c);
        if (!(leaf_box_lower_bound(leaf, qx, qy, qz) > bound_d2)) keep |= 1 << c;
        else PPCR_STAT(leaves_skipped, 1);
    }
    return keep;
}

// Leaves the (at most m) nearest targets with d2 < r2f in L.  pts = Morton-sorted target, .w = original index.
// bound0 <= r2f is a caller-supplied squared distance within which at least m targets are KNOWN to lie (r2f when
// nothing is known): points farther than it cannot be among the m nearest, so subtrees beyond it are never opened.
// `stack` must hold 2 * kTreeStack ints: (node, lower bound of its box) pairs.
template <class List>
PPCR_HD void tree_search(const TreeGeom& g, const TreeNode* __restrict__ nodes, const float4* __restrict__ pts,
                         float qx, float qy, float qz, float r2f, float bound0, List& L, int* stack)
{
    const unsigned long long r2key = static_cast<unsigned long long>(float_bits(r2f)) << 32;  // keys of d2 >= r2f are > this
    float bound_d2 = bound0 < r2f ? bound0 : r2f;  // no unseen point farther than this can enter the list
    float thr = prune_threshold(bound_d2);
    // Two kinds of work alternate, each done by all threads of a warp at the same time: opening inner nodes (their
    // leaf children go to a short pending list, the inner ones back on the stack) and scanning pending leaves.
    constexpr int kPend = 16;
    int pend[2 * kPend];
    int sp = 0, np = 0;
    {
        // Start at the deepest node whose subtree holds every target within the bound, not at the root: while the
        // search ball (inflated by twice the binning slack, so that a point's side of a centre plane is certain) lies
        // on one side of all three centre planes of a node, only that child can hold candidates.  One comparison per
        // axis and level replaces opening the node (eight child tests) on the shared upper part of every query's path.
        TreeNode n = load_node(nodes);
        if (n.end <= n.begin) return;
        int at = 0;
        const float rho = sqrtf(bound_d2) * 1.00001f + 2.0f * g.slack;
        const float lox = qx - rho, hix = qx + rho, loy = qy - rho, hiy = qy + rho, loz = qz - rho, hiz = qz + rho;
        bool leaf = n.child < 0;
        while (!leaf) {
            int oct;
            if (lox >= n.cx) oct = 1;
            else if (hix < n.cx) oct = 0;
            else break;
            if (loy >= n.cy) oct |= 2;
            else if (!(hiy < n.cy)) break;
            if (loz >= n.cz) oct |= 4;
            else if (!(hiz < n.cz)) break;
            if (!((n.mask >> oct) & 1)) return;  // the only octant the ball touches is empty
            at = n.child + oct;
            leaf = ((n.mask >> (16 + oct)) & 1) != 0;
            if (!leaf) n = load_node(nodes + at);
        }
        if (leaf) {
            pend[0] = at;
            pend[1] = 0;  // float bits of 0.0f
            np = 1;
        } else {
            stack[0] = at;
            stack[1] = 0;
            sp = 1;
        }
    }
    for (;;) {
        // ---- open inner nodes until the stack is empty or the pending list could overflow ----
        while (sp > 0 && np <= kPend - 8) {
            --sp;
            // the bound may have shrunk since this node was pushed: re-test with the lower bound stored beside it
            if (bits_float(static_cast<uint32_t>(stack[2 * sp + 1])) > thr) {
                PPCR_STAT(stack_skipped, 1);
                continue;
            }
            PPCR_STAT(opens, 1);
            const TreeNode n = load_node(nodes + stack[2 * sp]);
            // Children far-to-near, so that the octant holding q is opened / scanned first.  The lower bound of a
            // child box is a sum of three per-axis gaps, each of which takes one of two values (the child's half on
            // q's side of the centre plane, or the other one): six gaps serve all eight children.
            const int oct = (qx >= n.cx ? 1 : 0) | (qy >= n.cy ? 2 : 0) | (qz >= n.cz ? 4 : 0);
            const float ch = n.half * 0.5f;
            const float hi = ch + g.slack;
            const float gx_lo = axis_gap2(qx, n.cx - ch, hi), gx_hi = axis_gap2(qx, n.cx + ch, hi);
            const float gy_lo = axis_gap2(qy, n.cy - ch, hi), gy_hi = axis_gap2(qy, n.cy + ch, hi);
            const float gz_lo = axis_gap2(qz, n.cz - ch, hi), gz_hi = axis_gap2(qz, n.cz + ch, hi);
            const float nx = (oct & 1) ? gx_hi : gx_lo, fx = (oct & 1) ? gx_lo : gx_hi;  // near / far side per axis
            const float ny = (oct & 2) ? gy_hi : gy_lo, fy = (oct & 2) ? gy_lo : gy_hi;
            const float nz = (oct & 4) ? gz_hi : gz_lo, fz = (oct & 4) ? gz_lo : gz_hi;
            const int mask = n.mask;
#pragma unroll
            for (int k = 7; k >= 0; --k) {
                const int c = oct ^ k;
                const float lb = ((k & 1) ? fx : nx) + ((k & 2) ? fy : ny) + ((k & 4) ? fz : nz);
                if (((mask >> c) & 1) && !(lb > thr)) {
                    if ((mask >> (16 + c)) & 1) {  // a leaf
                        pend[2 * np] = n.child + c;
                        pend[2 * np + 1] = static_cast<int>(float_bits(lb));
                        ++np;
                    } else {
                        stack[2 * sp] = n.child + c;
                        stack[2 * sp + 1] = static_cast<int>(float_bits(lb));
                        ++sp;
                    }
                }
            }
        }
        if (np == 0) {
            if (sp == 0) break;
            continue;
        }
        // ---- scan the pending leaves, nearest (appended last) first ----
        while (np > 0) {
            --np;
            if (bits_float(static_cast<uint32_t>(pend[2 * np + 1])) > thr) {
                PPCR_STAT(leaves_skipped, 1);
                continue;
            }
            const TreeNode n = load_node(nodes + pend[2 * np]);
            // (a leaf other than a root leaf carries the box of its points, tree_box_leaf)
            if (pend[2 * np] != 0 && leaf_box_lower_bound(n, qx, qy, qz) > bound_d2) {
                PPCR_STAT(leaves_skipped, 1);
                continue;
            }
            PPCR_STAT(leaves, 1);
            PPCR_STAT(points, n.end - n.begin);
            // 32 points at a time, in two passes so that the threads of a warp stay together: first a plain distance
            // test of every point against the current bound (a bit per survivor, nothing else), then the survivors --
            // re-read from L1 -- go through the list one after the other.  The expensive, divergent part (the heap
            // update) is thereby reached by all threads at the same time instead of point by point.
            for (int j0 = n.begin; j0 < n.end; j0 += 32) {
                const int cnt = n.end - j0 < 32 ? n.end - j0 : 32;
                uint32_t pass = 0;
                for (int t = 0; t < cnt; ++t) {
                    const float4 p = load_point(pts + j0 + t);
                    const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
                    if (d2 <= bound_d2) pass |= 1u << t;
                }
                while (pass) {
                    const int t = lowest_bit(pass);
                    pass &= pass - 1;
                    PPCR_STAT(survivors, 1);
                    const float4 p = load_point(pts + j0 + t);
                    const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
                    if (d2 <= bound_d2) {  // the bound may have shrunk since the first pass
                        const unsigned long long k2 = make_key(d2, static_cast<int>(float_bits(p.w)));
                        if (k2 <= r2key && k2 < L.worst()) {  // k2 <= r2key  <=>  d2 < r2f (keys carry +1)
                            L.insert(k2);
                            PPCR_STAT(inserts, 1);
                            const unsigned long long w = L.worst();
                            if (w != kKeyInf) {
                                const float wd = key_d2(w);
                                bound_d2 = wd < bound_d2 ? wd : bound_d2;
                                thr = prune_threshold(bound_d2);
                            }
                        }
                    }
                }
            }
        }
    }
}

// ---- traversal with a fixed bound: which leaves can hold a target within bound_d2 of q ------------------------------
//
// The deepest node whose subtree holds every target within bound_d2 of q (the start of the walk; see tree_search).
// Returns its index and whether it is a leaf; -1 when no target can lie within the bound.
PPCR_HD int tree_start_node(const TreeGeom& g, const TreeNode* __restrict__ nodes, float qx, float qy, float qz, float bound_d2,
                            bool* is_leaf)
{
    TreeNode n = load_node(nodes);
    if (n.end <= n.begin) return -1;
    int at = 0;
    const float rho = sqrtf(bound_d2) * 1.00001f + 2.0f * g.slack;
    const float lox = qx - rho, hix = qx + rho, loy = qy - rho, hiy = qy + rho, loz = qz - rho, hiz = qz + rho;
    bool leaf = n.child < 0;
    while (!leaf) {
        int oct;
        if (lox >= n.cx) oct = 1;
        else if (hix < n.cx) oct = 0;
        else break;
        if (loy >= n.cy) oct |= 2;
        else if (!(hiy < n.cy)) break;
        if (loz >= n.cz) oct |= 4;
        else if (!(hiz < n.cz)) break;
        if (!((n.mask >> oct) & 1)) return -1;  // the only octant the ball touches is empty
        at = n.child + oct;
        leaf = ((n.mask >> (16 + oct)) & 1) != 0;
        if (!leaf) n = load_node(nodes + at);
    }
    *is_leaf = leaf;
    return at;
}

// Opens inner node `node` for query q: the non-empty children whose boxes come within the pruning threshold are handed on --
// the inner ones one by one, inner(child), the leaves together, leaves(first child, bit mask of the leaf children), so that a
// consumer that queues them needs one reservation per node.  Same lower bounds as tree_search.
template <class Leaves, class Inner>
PPCR_HD void tree_open_node(const TreeGeom& g, const TreeNode* __restrict__ nodes, int node, float qx, float qy, float qz, float thr,
                            Leaves& leaves, Inner& inner)
{
    const TreeNode n = load_node(nodes + node);
    PPCR_STAT(opens, 1);
    const float ch = n.half * 0.5f;
    const float hi = ch + g.slack;
    const float gx[2] = {axis_gap2(qx, n.cx - ch, hi), axis_gap2(qx, n.cx + ch, hi)};
    const float gy[2] = {axis_gap2(qy, n.cy - ch, hi), axis_gap2(qy, n.cy + ch, hi)};
    const float gz[2] = {axis_gap2(qz, n.cz - ch, hi), axis_gap2(qz, n.cz + ch, hi)};
    const int mask = n.mask;
    int leaf_mask = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const float lb = gx[c & 1] + gy[(c >> 1) & 1] + gz[(c >> 2) & 1];
        if (((mask >> c) & 1) && !(lb > thr)) {
            if ((mask >> (16 + c)) & 1) leaf_mask |= 1 << c;
            else inner(n.child + c);
        }
    }
    if (leaf_mask) leaves(n.child, leaf_mask);
}

// First phase of the queued search kernel (k_search_q): the pruning bound is known up front (the search kernel's warm
// bound) and does not change during the walk, so the walk only has to NAME the leaves whose boxes -- emit(first, mask):
// the nodes first + c for every set bit c of mask, the leaf children of one opened node together --
// come within the bound; testing their points is somebody else's work.  Same start-node descent, same box lower bounds
// and the same pruning threshold as tree_search: a leaf tree_search would scan under this bound is always emitted.
// `stack` must hold kTreeStack ints.  Returns false when emit refused a leaf (queue full): the caller falls back.
template <class Emit>
PPCR_HD bool tree_collect_leaves(const TreeGeom& g, const TreeNode* __restrict__ nodes, float qx, float qy, float qz,
                                 float bound_d2, Emit& emit, int* stack)
{
    const float thr = prune_threshold(bound_d2);
    bool is_leaf = false;
    const int at = tree_start_node(g, nodes, qx, qy, qz, bound_d2, &is_leaf);
    if (at < 0) return true;
    if (is_leaf) return emit(at, 1);
    int sp = 0;
    stack[sp++] = at;
    bool ok = true;
    auto leaves = [&](int first, int mask) {
        mask = tree_prune_leaf_mask(nodes, first, mask, qx, qy, qz, bound_d2);
        if (mask) ok = emit(first, mask) && ok;
    };
    auto inner = [&](int child) { stack[sp++] = child; };
    while (sp > 0) tree_open_node(g, nodes, stack[--sp], qx, qy, qz, thr, leaves, inner);
    return ok;
}

// Second phase: the points of one emitted leaf against one query.  Every point of the leaf with d2 <= limit_d2 is reported,
// 32 points at a time: push(first position, survivor bit mask).  limit_d2 folds the two tests a candidate passes in
// tree_search -- d2 <= bound and d2 < r2f (strict) -- into one comparison: candidate_limit(bound, r2f).  Two passes like
// tree_search's leaf scan: the point loop only sets bits, so the caller reserves room for all survivors of the group
// with ONE atomic instead of one per survivor.
PPCR_HD float candidate_limit(float bound_d2, float r2f)
{
    if (bound_d2 < r2f) return bound_d2;
    // the largest float below r2f (r2f > 0: a squared radius): d2 <= it  <=>  d2 < r2f
    return r2f > 0.f ? bits_float(float_bits(r2f) - 1u) : -1.f;
}

template <class Push>
PPCR_HD void leaf_candidates(const TreeNode* __restrict__ nodes, const float4* __restrict__ pts, int node, float qx, float qy,
                             float qz, float limit_d2, Push& push)
{
    struct Range {
        int begin, end;
    };
#if defined(__CUDA_ARCH__)
    const int2 be = __ldg(reinterpret_cast<const int2*>(&nodes[node].begin));  // only the range of the leaf is needed
    const Range n{be.x, be.y};
#else
    const Range n{nodes[node].begin, nodes[node].end};
#endif
    PPCR_STAT(leaves, 1);
    PPCR_STAT(points, n.end - n.begin);
    const int last = n.end - 1;
    for (int j0 = n.begin; j0 < n.end; j0 += 32) {
        const int stop = n.end - j0 < 32 ? n.end - j0 : 32;
        uint32_t pass = 0;
        // four independent loads in flight per thread; past the end the index is clamped and the bits masked off below
        for (int t = 0; t < stop; t += 4) {
            float4 p[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) p[u] = load_point(pts + (j0 + t + u < last ? j0 + t + u : last));
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z) <= limit_d2) pass |= 1u << ((t + u) & 31);
        }
        if (stop < 32) pass &= (1u << stop) - 1u;
        if (pass) push(j0, pass);
    }
}

// Third phase: the m best of a query's n candidates (every one within the radius and distinct; cand(c) = position of
// candidate c in the sorted target), as keys in the heap column k[i * STRIDE].  Fewer than m: all of them.  Otherwise
// the first m are stored, made a max-heap (Floyd) and the rest streamed through its root.  The candidates arrive in no
// particular order (they were pushed by many threads), and the order of a row decides the last bits of its float32
// weight sums: the results are therefore SORTED ascending by (distance, index) -- FLANN's own order -- by an in-place
// heap sort, which makes the association a pure function of the two clouds.  Returns the number of results, left in
// slots [0, count); *kth = the key of the m-th best when m were found, kKeyInf otherwise.
template <int STRIDE, class Cand>
PPCR_HD int select_candidates(const float4* __restrict__ pts, const Cand& cand, int n, int m, float qx, float qy, float qz,
                              unsigned long long* k, unsigned long long* kth)
{
    HeapList<STRIDE, 1> H;
    H.k = k;
    H.m = n < m ? n : m;  // heap size: all n candidates when fewer than m
    H.n = 0;
    const int n0 = H.m;
    // candidates are fetched four at a time (clamped index, surplus discarded) so that their latencies overlap
    for (int c = 0; c < n0; c += 4) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u] = load_point(pts + cand(c + u < n0 ? c + u : n0 - 1));
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c + u < n0)
                k[(c + u) * STRIDE] = make_key(dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z), static_cast<int>(float_bits(p[u].w)));
    }
    for (int i = n0 / 2 - 1; i >= 0; --i) H.sift_down(i, k[i * STRIDE]);
    for (int c = m; c < n; c += 4) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u] = load_point(pts + cand(c + u < n ? c + u : n - 1));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const unsigned long long x =
                make_key(dist2_exact(qx, qy, qz, p[u].x, p[u].y, p[u].z), static_cast<int>(float_bits(p[u].w)));
            if (c + u < n && x < k[0]) {
                H.sift_down(0, x);
                PPCR_STAT(inserts, 1);
            }
        }
    }
    *kth = (n >= m && n0 > 0) ? k[0] : kKeyInf;
    // heap sort: the root (largest) goes to the end of the shrinking heap
    for (int e = n0 - 1; e > 0; --e) {
        const unsigned long long top = k[0], x = k[e * STRIDE];
        H.m = e;
        H.sift_down(0, x);
        k[e * STRIDE] = top;
    }
    return n0;
}

}  // namespace ppcr
#endif
