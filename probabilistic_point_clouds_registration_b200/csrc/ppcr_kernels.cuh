// ppcr_kernels.cuh -- sm_100a kernels of the registration hot path.
//
//   tree build      k_bbox, k_tree_keys, (radix sort), k_tree_gather, k_tree_root, k_tree_split_level
//                                                                     (replaces the kd-tree build, registration.cc:66-67)
//   radius search   k_search (first search of an align()),            (replaces the radiusSearch loop, :72-81, and the
//                   k_search_q (every search after a cloud move)       CSR assembly, :69-83)
//   weights + J^TWJ k_eval<FAST>                                      (WeightUpdaterCallback, ProbabilisticWeights,
//                                                                      ErrorTerm + Ceres' Jacobian evaluation)
//   LM controller   k_controller                                      (ceres::Solve's trust-region loop, pose
//                                                                      composition, cost drop, hasConverged)
//   cloud move      k_transform                                       (pcl::transformPointCloud, :110-112)
//   voxel filter    k_voxel_* (+ a radix sort of the voxel keys)      (pcl::VoxelGrid, :24-41)
//
// Data layout in HBM (per pair):
//   tgt_sorted  float4[n_tgt]      target points in Morton order, .w = original index (int bits)
//   tgt_raw     float4[n_tgt]      target points in caller order (coordinate gather of the found neighbours)
//   nodes       TreeNode[]         linear octree over tgt_sorted (ppcr_tree.h), 32-byte records, 8 children adjacent
//   src         float4[n_src]      the moving source cloud (filtered), Morton-sorted once so that the 32 queries of a
//                                  warp walk the same part of the tree; .w = original index
//   nbr_pos     int[m][n_pad]      slot-major association: entry (k, i) is the position IN tgt_sorted of a neighbour of
//                                  source i (one warp reads 128 contiguous bytes per slot); the evaluation gathers the
//                                  16-byte target points from tgt_sorted, which stays in L2 and -- neighbours of
//                                  neighbouring queries being neighbours in Morton order -- mostly in L1
//   inv_perm    int[n_tgt]         original target index -> position in tgt_sorted
//   nbr_cnt     int[n_pad]
//   partials    double[blocks][24] per-block moment sums, reduced in a fixed order by the controller
// No tensor cores: nothing on this path is a dense contraction; the kernels are HBM/L2-bound streaming passes.
#ifndef PPCR_KERNELS_CUH
#define PPCR_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "ppcr_eval.h"
#include "ppcr_lm.h"
#include "ppcr_tree.h"

namespace ppcr {

constexpr int kSearchThreads = 128;  // one query per thread
constexpr int kEvalThreads = 256;      // float64 ("exact") evaluation: one row per thread, grid-stride, moments in shared memory
constexpr int kEvalFastThreads = 128;  // float32-row evaluation: tiles of 128 rows staged by bulk copies, moments in registers
#ifndef PPCR_EVAL_LANEACC
#define PPCR_EVAL_LANEACC 1  // 1: the rows of a warp are added by a float32 transpose-reduction, one float64 accumulator per lane
#endif
#ifndef PPCR_EVAL_FAST_BLOCKS
#define PPCR_EVAL_FAST_BLOCKS (PPCR_EVAL_LANEACC ? 5 : 4)
#endif
constexpr int kEvalFastBlocks = PPCR_EVAL_FAST_BLOCKS;  // resident blocks per SM the fast evaluation is held to (5: 96 registers per thread, no spills; 6: 80 with spills, no faster)
#ifndef PPCR_EVAL_STAGES
#define PPCR_EVAL_STAGES 3
#endif
constexpr int kEvalStages = PPCR_EVAL_STAGES;  // staged tiles per block
#ifndef PPCR_EVAL_BATCH
#define PPCR_EVAL_BATCH 5
#endif
#ifndef PPCR_EVAL_MIN_BLOCKS
#define PPCR_EVAL_MIN_BLOCKS 4  // resident blocks per SM the register allocation of k_evalctl is held to
#endif
constexpr int kFoldGroup = 16;    // blocks per first-level group of the moment reduction
constexpr int kFoldChains = 4;    // interleaved chains of the second level
constexpr int kErrRowOverflow = 200;  // PairState::error: a row of a wide association filled up
constexpr int kMailDoubles = 32;  // 24 moments + K + sequence stamp, padded
constexpr double kMailEpochStride = 67108864.0;  // 2^26 ticks per stamp epoch (mailboxes outlive handles; a handle's ticks stay far below)
constexpr unsigned kFull = 0xffffffffu;

struct PairDev {
    const float4* tgt_sorted;
    const float4* tgt_raw;
    const TreeNode* nodes;
    TreeGeom tree;
    int n_tgt;
    float4* src;
    int n_src;
    int n_pad;
    int m;          // result capacity = min(max_neighbours, n_tgt)
    int search_cap; // slots per query in the search kernel's shared-memory column (CollectList: > m)
    int search_queued;  // 1: searches that follow a cloud move are k_search_q's, k_search only does the first of an align()
    int q_cand;                // k_search_q: candidate positions per query (search_q_cand(max_neighbours))
    float q_heavy;             // k_search_q: a query expecting more than q_heavy * q_cand candidates counts as heavy
    int overflow_at;           // a row that reaches this count may have lost neighbours (wide rows; INT_MAX otherwise)
    int q_flags;               // k_search_q tuning switches (PPCR_Q_FLAGS): 1 = fallback rows are not sorted, 2 = the fallback starts from the incoming bound
    int q_leaves;              // k_search_q: leaves one query may queue before it falls back to tree_search (kQTaskPerQuery)
    unsigned char* q_scratch;  // k_search_q's task / candidate queues: one slab per block of its grid (all pairs share the pointer)
    float r2f;      // float(radius * radius): strict membership bound (FLANN)
    int* nbr_pos;   // [m][n_pad] slot-major association: positions in tgt_sorted
    const int* inv_perm;  // [n_tgt] original index -> position in tgt_sorted
    float* nbr_d2;  // optional (stage API only), may be null
    float* nbr_kth; // d2 of the m-th neighbour found by the last search (+inf when fewer were found): warm start
    int* nbr_cnt;
    double* partials;        // [n_eval_blocks][24] per-block moment sums
    double* group_partials;  // [n_eval_groups][24] sums of kFoldGroup consecutive blocks
    int* group_ticket;       // [n_eval_groups] blocks of the group that have published
    int n_eval_blocks;
    int n_eval_groups;
    int max_hist;
    PairState* state;
    const Config* cfg;
    double* history;
    IterStats* stats;
    WeightCfg wcfg;
    double* dump_w;  // parity dumps only (ppcr_weights_normal_eq), normally null: [m][n_pad] weights written by the evaluation itself
    // sharded mode: mailbox exchange of the moment vector between ranks
    double* mailbox;          // [world][kMailDoubles] on THIS device, written by the peers
    double* peer_mailbox[8];  // the same buffer on every rank (peer-mapped), indexed by rank
    double mail_base;         // this handle's stamp epoch * kMailEpochStride: stamps are mail_base + tick number
    int rank, world;
    long long spin_limit;
};

// ------------------------------------------------------------------------------------------------------------
// bounding box
// ------------------------------------------------------------------------------------------------------------

// min / max corner of a cloud: out[0..2] = min, out[3..5] = max, encoded as order-preserving uints
__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float ord2f(unsigned u)
{
    unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#if defined(__CUDA_ARCH__)
    return __uint_as_float(v);
#else
    float f;
    memcpy(&f, &v, 4);
    return f;
#endif
}

__global__ void k_bbox(const float4* __restrict__ pts, int n, unsigned* __restrict__ out6)
{
    unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        const unsigned a[3] = {f2ord(p.x), f2ord(p.y), f2ord(p.z)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            lo[k] = min(lo[k], a[k]);
            hi[k] = max(hi[k], a[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = min(lo[k], __shfl_xor_sync(kFull, lo[k], o));
            hi[k] = max(hi[k], __shfl_xor_sync(kFull, hi[k], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(out6 + k, lo[k]);
            atomicMax(out6 + 3 + k, hi[k]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// exclusive scan, three passes; each block owns kScanTile consecutive items (used by the voxel filter)
// ------------------------------------------------------------------------------------------------------------

constexpr int kScanThreads = 256;
constexpr int kScanPerThread = 8;
constexpr int kScanTile = kScanThreads * kScanPerThread;

__global__ void k_scan_local(int* __restrict__ data, int n, int* __restrict__ block_sums)
{
    __shared__ int warp_sums[kScanThreads / 32];
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPerThread;
    int v[kScanPerThread];
    int sum = 0;
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0;
        sum += v[k];
    }
    int incl = sum;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < kScanThreads / 32) ? warp_sums[lane] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    int excl = incl - sum + (warp > 0 ? warp_sums[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k) {
        if (base + k < n) data[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == kScanThreads - 1) block_sums[blockIdx.x] = excl;
}

__global__ void k_scan_sums(int* __restrict__ block_sums, int n_blocks, int* __restrict__ total_out)
{
    // single block: serial over chunks of blockDim.x, parallel inside a chunk
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int base = 0; base < n_blocks; base += blockDim.x) {
        const int i = base + threadIdx.x;
        const int v = (i < n_blocks) ? block_sums[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = (lane < nw) ? warp_sums[lane] : 0;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, w, o);
                if (lane >= o) w += t;
            }
            if (lane < nw) warp_sums[lane] = w;
        }
        __syncthreads();
        const int excl = incl - v + (warp > 0 ? warp_sums[warp - 1] : 0) + carry;
        if (i < n_blocks) block_sums[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void k_scan_add(int* __restrict__ data, int n, const int* __restrict__ block_sums, int total_slot)
{
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanPerThread;
    const int add = block_sums[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanPerThread; ++k)
        if (base + k < n) data[base + k] += add;
    (void)total_slot;
}

// ------------------------------------------------------------------------------------------------------------
// octree build over the Morton-sorted target (ppcr_tree.h)
// ------------------------------------------------------------------------------------------------------------

__global__ void k_tree_keys(const float4* __restrict__ pts, int n, TreeGeom g, unsigned long long* __restrict__ keys,
                            unsigned* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    keys[i] = tree_key(g, p.x, p.y, p.z);
    vals[i] = static_cast<unsigned>(i);
}

// sorted[j] = pts[vals[j]] with .w = the original index.  keep_w: the input's own .w is carried instead (a cloud that
// is already tagged, i.e. re-sorting the source)
__global__ void k_tree_gather(const float4* __restrict__ pts, const unsigned* __restrict__ vals, int n, int keep_w,
                              float4* __restrict__ sorted, int* __restrict__ inv_perm)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const unsigned i = vals[j];
    float4 p = pts[i];
    if (!keep_w) p.w = __int_as_float(static_cast<int>(i));
    sorted[j] = p;
    if (inv_perm) inv_perm[i] = j;
}

// tag every point with its own index in .w (the source keeps its caller-order identity through the Morton sort)
__global__ void k_tag_index(float4* __restrict__ pts, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pts[i].w = __int_as_float(i);
}

struct TreeCounters {
    int n_nodes;
    int level_begin[kTreeBits + 3];  // nodes of level l are [level_begin[l], level_begin[l+1])
    int ticket[kTreeBits + 1];
};

__global__ void k_tree_root(TreeNode* __restrict__ nodes, int n, TreeGeom g, TreeCounters* __restrict__ tc)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    TreeNode r;
    r.begin = 0;
    r.end = n;
    r.child = -1;
    r.mask = 0;
    tree_node_box(g, 0, 0ull, &r);
    nodes[0] = r;
    tc->n_nodes = 1;
    for (int l = 0; l < kTreeBits + 3; ++l) tc->level_begin[l] = l == 0 ? 0 : 1;
    for (int l = 0; l < kTreeBits + 1; ++l) tc->ticket[l] = 0;
}

// one thread per node of `level`: split it when it holds more than leaf_cap points
__global__ void k_tree_split_level(TreeGeom g, const unsigned long long* __restrict__ keys, TreeNode* __restrict__ nodes,
                                   TreeCounters* __restrict__ tc, int level)
{
    const int lb = tc->level_begin[level], le = tc->level_begin[level + 1];
    for (int ni = lb + blockIdx.x * blockDim.x + threadIdx.x; ni < le; ni += gridDim.x * blockDim.x) {
        const int count = nodes[ni].end - nodes[ni].begin;
        if (count > g.leaf_cap) {
            const int base = atomicAdd(&tc->n_nodes, 8);
            if (base + 8 <= g.n_nodes_cap) tree_split_node(g, keys, nodes, ni, base);  // else: stays a (large) leaf
        }
    }
    __threadfence();
    __shared__ int s_last;
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&tc->ticket[level], 1) == static_cast<int>(gridDim.x) - 1);
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        // slots whose allocation did not fit were never written: the buffer is pre-filled with empty leaves
        const int nn = min(*reinterpret_cast<volatile int*>(&tc->n_nodes), g.n_nodes_cap);
        for (int l = level + 2; l < kTreeBits + 3; ++l) tc->level_begin[l] = nn;
    }
}

__global__ void k_tree_finalize(TreeNode* __restrict__ nodes, const TreeCounters* __restrict__ tc, int cap)
{
    const int n = min(tc->n_nodes, cap);
    for (int ni = blockIdx.x * blockDim.x + threadIdx.x; ni < n; ni += gridDim.x * blockDim.x) tree_mark_leaf_children(nodes, ni);
}

// Small targets: the whole build -- root, every level, leaf marks, leaf boxes -- by ONE block in one launch.  Level by level it is
// 19 launches of a few microseconds of work each: 0.30 ms for a 10k-point cloud, a tenth of its whole registration.  Same
// functions, same tree (node numbers may differ; no result depends on them).
constexpr int kTreeSmallThreads = 1024;
__global__ void __launch_bounds__(kTreeSmallThreads) k_tree_build_small(TreeNode* __restrict__ nodes, int n, TreeGeom g,
                                                                        const unsigned long long* __restrict__ keys,
                                                                        const float4* __restrict__ pts, TreeCounters* __restrict__ tc)
{
    __shared__ int s_n_nodes, s_lb, s_le;
    if (threadIdx.x == 0) {
        TreeNode r;
        r.begin = 0;
        r.end = n;
        r.child = -1;
        r.mask = 0;
        tree_node_box(g, 0, 0ull, &r);
        nodes[0] = r;
        s_n_nodes = 1;
        s_lb = 0;
        s_le = 1;
    }
    __syncthreads();
    for (int level = 0; level < kTreeBits; ++level) {
        const int lb = s_lb, le = s_le;
        if (lb >= le) break;  // (block-uniform) nothing was split on the level above
        for (int ni = lb + threadIdx.x; ni < le; ni += kTreeSmallThreads) {
            if (nodes[ni].end - nodes[ni].begin > g.leaf_cap) {
                const int base = atomicAdd(&s_n_nodes, 8);
                if (base + 8 <= g.n_nodes_cap) tree_split_node(g, keys, nodes, ni, base);  // else: stays a (large) leaf
            }
        }
        __syncthreads();  // the children written above are read by other threads of this block on the next level
        if (threadIdx.x == 0) {
            s_lb = le;
            s_le = min(s_n_nodes, g.n_nodes_cap);
        }
        __syncthreads();
    }
    const int n_nodes = min(s_n_nodes, g.n_nodes_cap);
    for (int ni = threadIdx.x; ni < n_nodes; ni += kTreeSmallThreads) tree_mark_leaf_children(nodes, ni);
    __syncthreads();  // a leaf is recognised by child < 0, which its box overwrites
    for (int ni = threadIdx.x; ni < n_nodes; ni += kTreeSmallThreads) tree_box_leaf(nodes, pts, ni);
    if (threadIdx.x == 0) tc->n_nodes = s_n_nodes;
}

// every non-empty leaf but a root leaf gets the bounding box of its points (ppcr_tree.h, "tight leaf boxes"); after k_tree_finalize
__global__ void k_tree_leaf_boxes(TreeNode* __restrict__ nodes, const float4* __restrict__ pts, const TreeCounters* __restrict__ tc, int cap)
{
    const int n = min(tc->n_nodes, cap);
    for (int ni = blockIdx.x * blockDim.x + threadIdx.x; ni < n; ni += gridDim.x * blockDim.x) tree_box_leaf(nodes, pts, ni);
}

// ------------------------------------------------------------------------------------------------------------
// radius search: one thread per query walks the octree with a register-resident sorted top-m list
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ float transform_row(const double* T, double x, double y, double z)
{
    // pcl::transformPointCloud: double arithmetic without contraction, then one rounding to float
    double acc = __dmul_rn(T[0], x);
    acc = __dadd_rn(acc, __dmul_rn(T[1], y));
    acc = __dadd_rn(acc, __dmul_rn(T[2], z));
    acc = __dadd_rn(acc, T[3]);
    return __double2float_rn(acc);
}

// A query with a non-finite coordinate has no neighbours (every distance comparison fails, as in FLANN); it must not walk
// the tree either: its box tests cannot prune anything.
__device__ __forceinline__ bool finite_query(const float4& q) { return isfinite(q.x) && isfinite(q.y) && isfinite(q.z); }

struct SearchOut {  // where a query's results go (copied out of the PairDev once per block)
    int* __restrict__ nbr_pos;
    float* __restrict__ nbr_d2;
    const int* __restrict__ inv_perm;
    size_t n_pad;
};

__device__ __forceinline__ void search_store(const SearchOut& O, int i, int e, unsigned long long key)
{
    const size_t o = static_cast<size_t>(e) * O.n_pad + i;
    O.nbr_pos[o] = __ldg(O.inv_perm + key_index(key));
    if (O.nbr_d2) O.nbr_d2[o] = key_d2(key);
}

constexpr int kSearchChunk = kSearchThreads;  // queries handed out per grab of the work cursor

// Persistent blocks pull chunks of 128 consecutive (Morton-sorted) queries from a cursor in the pair state, so dense
// and sparse parts of the cloud balance across SMs and an idle launch (LM phase) costs one block wave of early exits.
//
// Fused into the load of the query: the cloud move of the PREVIOUS outer iteration (registration.cc:110-112,
// x <- float(dT * x) in double arithmetic, written back in place) -- every search of an align() except the first
// follows a pose update -- and the warm start of the pruning bound: the m targets found last time lie within
// sqrt(d_m) + |dx| of the moved query, so nothing farther than that can be among the m nearest now.
//
// The m best candidates of a query live in a binary max-heap in shared memory (one column per thread, m * 8 bytes of
// dynamic shared memory per thread): replacing the root and sifting down costs ~log2(m) steps, about half the
// instructions of a sorted register list at m = 10..20, and works for any m.  Rows are stored in heap order; nothing
// downstream depends on the order inside a row (the host sorts rows it hands out).
// The list of the m best candidates.  VAR 0 (the product): max-heap that starts full of kKeyInf.  Tuning variants, all
// measured slower on the 1M-point pair: 1 heap filled by appends, 3 heap filled bottom-up, 4 unordered
// column + worst scan, 16 collect + select; 32 = the product without the exact warm bound.
template <int VAR>
struct SearchList {
    using type = HeapList<kSearchThreads, 0>;
};
template <>
struct SearchList<1> {
    using type = HeapList<kSearchThreads, 1>;
};
template <>
struct SearchList<3> {
    using type = HeapList<kSearchThreads, 2>;
};
template <>
struct SearchList<4> {
    using type = ScanList<kSearchThreads>;
};
template <>
struct SearchList<16> {
    using type = CollectList<kSearchThreads>;
};

#if defined(PPCR_SEARCH_MIN_BLOCKS)  // tuning builds only: ptxas' own choice (48 registers) measured fastest
#define PPCR_SEARCH_BOUNDS __launch_bounds__(kSearchThreads, PPCR_SEARCH_MIN_BLOCKS)
#else
#define PPCR_SEARCH_BOUNDS __launch_bounds__(kSearchThreads)
#endif
template <int VAR>
__global__ void PPCR_SEARCH_BOUNDS k_search(const PairDev* __restrict__ pairs)
{
    extern __shared__ unsigned long long s_heap[];
    const PairDev& P = pairs[blockIdx.y];
    PairState* st = P.state;
    if (st->phase != PH_SEARCH) return;
    __shared__ double s_T[12];
    __shared__ int s_chunk;
    const bool moving = st->apply_dT != 0;
    if (moving && P.search_queued) return;  // k_search_q, launched right behind this kernel, does this one
    if (threadIdx.x < 12) s_T[threadIdx.x] = st->dT[threadIdx.x];
    // the geometry and the pointers the walk uses, held in registers (P lives in global memory)
    const int m = P.m;
    const int cap = P.search_cap;
    const int n_src = P.n_src;
    const float r2f = P.r2f;
    const TreeGeom geom = P.tree;
    const TreeNode* __restrict__ nodes = P.nodes;
    const float4* __restrict__ tgt_sorted = P.tgt_sorted;
    const SearchOut out{P.nbr_pos, P.nbr_d2, P.inv_perm, static_cast<size_t>(P.n_pad)};
    float4* __restrict__ src = P.src;
    float* __restrict__ nbr_kth = P.nbr_kth;
    int* __restrict__ nbr_cnt = P.nbr_cnt;
    const int n_chunks = (n_src + kSearchChunk - 1) / kSearchChunk;
    int stack[2 * kTreeStack];
    int cnt_total = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(&st->search_cursor, 1);
        __syncthreads();
        const int chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const int i = chunk * kSearchChunk + threadIdx.x;
        if (i >= n_src) continue;
        float4 q = src[i];
        float bound0 = r2f;
        if (moving) {
            const double x = q.x, y = q.y, z = q.z;
            const float nx = transform_row(s_T, x, y, z), ny = transform_row(s_T + 4, x, y, z),
                        nz = transform_row(s_T + 8, x, y, z);
            const float prev = nbr_kth[i];  // d2 of the m-th neighbour of the last search, +inf if it found fewer
            const float ddx = nx - q.x, ddy = ny - q.y, ddz = nz - q.z;
            const float move = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz) * 1.000001f;
            const float reach = sqrtf(prev) * 1.000001f + move;
            bound0 = fminf(r2f, reach * reach * 1.00001f);  // inf stays inf -> r2f
            q.x = nx;
            q.y = ny;
            q.z = nz;
            src[i] = q;
            if ((VAR & 32) == 0 && prev < __int_as_float(0x7f800000)) {
                // A saturated row: its m neighbours of the last search are m distinct targets, so the largest of their
                // distances to the MOVED query (same arithmetic as the walk, hence the same bits) bounds the new m-th
                // distance -- usually far tighter than the triangle inequality above, and independent of how far the
                // cloud moved.  The positions stream from the slot-major plane, the points come from L2 / L1.
                const int* __restrict__ pp = out.nbr_pos + i;
                float far2 = 0.f;
                int k = 0;
                for (; k + 5 <= m; k += 5) {
                    int pos[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) pos[u] = pp[static_cast<size_t>(k + u) * out.n_pad];
                    float4 t[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) t[u] = load_point(tgt_sorted + pos[u]);
#pragma unroll
                    for (int u = 0; u < 5; ++u) far2 = fmaxf(far2, dist2_exact(nx, ny, nz, t[u].x, t[u].y, t[u].z));
                }
                for (; k < m; ++k) {
                    const float4 t = load_point(tgt_sorted + pp[static_cast<size_t>(k) * out.n_pad]);
                    far2 = fmaxf(far2, dist2_exact(nx, ny, nz, t.x, t.y, t.z));
                }
                bound0 = fminf(bound0, far2);
            }
        }
        int cnt = 0;
        float kth = __int_as_float(0x7f800000);
        if (!finite_query(q)) {
            nbr_cnt[i] = 0;
            nbr_kth[i] = kth;
            continue;
        }
        typename SearchList<VAR>::type L;
        L.k = s_heap + threadIdx.x;
        L.init(m, cap);
        tree_search(geom, nodes, tgt_sorted, q.x, q.y, q.z, r2f, bound0, L, stack);
        L.finish();
        if (L.kth_key() != kKeyInf) kth = key_d2(L.kth_key());
        for (int s = L.begin(); s < L.end(); ++s) {
            const unsigned long long key = L.k[s * kSearchThreads];
            if (key != kKeyInf) search_store(out, i, cnt++, key);
        }
        nbr_cnt[i] = cnt;
        nbr_kth[i] = kth;
        cnt_total += cnt;
        if (cnt >= P.overflow_at) st->row_overflow = 1;
    }
    // association size: warp sum, one atomic per warp
    for (int o = 16; o > 0; o >>= 1) cnt_total += __shfl_xor_sync(kFull, cnt_total, o);
    if ((threadIdx.x & 31) == 0 && cnt_total)
        atomicAdd(reinterpret_cast<unsigned long long*>(&st->K), static_cast<unsigned long long>(cnt_total));
}

// ------------------------------------------------------------------------------------------------------------
// radius search, queued form: every search of an align() but the first
// ------------------------------------------------------------------------------------------------------------
//
// k_search is bound by instruction issue on a divergent walk: a thread opens nodes, tests points and updates its heap in
// turns, and at any moment the 32 threads of a warp want different turns (13 of 32 lanes active on average, 7 in the heap
// updates).  Every search but the first of an align() knows a tight pruning bound per query BEFORE it starts -- the
// distance of the farthest of last iteration's neighbours to the moved query -- so the walk does not need the heap to prune.
// That allows the three kinds of work to be separated and each to be done by all threads of the block at the same time,
// for whichever query it belongs to, through two queues per block:
//   A  (thread per query)   move the query, bound from the previous neighbours, walk the octree with that fixed bound and
//                           push every leaf within it as a task (query slot, node) on the block's task queue
//   B  (thread per task)    test the leaf's points against its query; survivors' positions go to the query's candidate list
//   C  (thread per query)   the m best of the candidates: build a heap of the first m, stream the rest through its root,
//                           heap-sort them into (distance, index) order (a row must not depend on who pushed first)
// A query whose tasks or candidates overflow the queues is searched by tree_search with the heap, as in k_search; a chunk in
// which most queries expect to (the cloud moved by more than the neighbour spacing) is walked with heaps as a whole.  The
// neighbour sets are identical to k_search's: same distance arithmetic, same strict radius test, same (distance, index) order.
// Leaves one query may queue before it falls back to tree_search.  It used to be 64, which looked harmless (5.6 leaves per query
// on average) and was not: while the pose is still off, about one query per few blocks sits 0.4 m from a dense surface and its
// bound touches 70-300 leaf cubes (a thick shell of small cells, nearly all of them without a point inside the bound).  Such a
// query then walked that shell ALONE with a heap -- 130 us on one thread with 127 waiting at the barrier -- and the blocks that
// drew one finished at 540 us when every other block was done at 350: a third of the in-loop search time was that tail
// (per-block globaltimer stamps, -DPPCR_Q_PROFILE).  Through the task queue the same leaves are 2-3 rounds of phase B for the
// whole block.  The cap that remains is the block's queue (kQTaskCap).
constexpr int kQTaskPerQuery = 4096;
constexpr int kQTaskCap = 8192;                  // leaf tasks per block of 128 queries (64 per query on average)
// candidate positions per query.  128 for max_neighbours = 20 was tried: one 120k-point pair 6.2 -> 6.0 ms, but a batch of them
// 403 -> 344 pairs/s (the slabs of six lanes crowd L2).  Re-swept with the tight leaf boxes (round 2): 64 / 80 / 96 / 128 -> 15.5 /
// 15.0 / 14.8 / 14.8 ms of search per registration on the 1M-point pair, a 120k-point pair alone 3.26 -> 3.01 ms at 96, batches and
// the 10M-point pair unchanged: 96 for every m; PPCR_Q_CAND overrides it (tuning)
PPCR_HD constexpr int search_q_cand(int /*max_nn*/) { return 96; }
constexpr int kQNodeBits = 25;                   // task = query slot << 25 | node index
constexpr uint32_t kQLowMask = (1u << kQNodeBits) - 1u;
// The two queues of a block live in GLOBAL memory (a scratch slab per resident block, L2 resident, accessed with .cg
// loads / stores so that they do not displace the tree from L1): in shared memory they limited the kernel to 6 blocks per
// SM and left the walk 50 KB of L1.  That also makes them cheap to size for the expensive queries -- the ones half a
// metre off a dense surface, whose bound touches 30-40 leaves -- so that those go through the queues like all others
// instead of falling back.  Shared memory keeps the queries, the counters and the heap columns.
PPCR_HD constexpr size_t search_q_smem(int m)
{
    return static_cast<size_t>(kSearchThreads) * (16u + 4u) + 8u * static_cast<size_t>(kSearchThreads) * static_cast<size_t>(m);
}
PPCR_HD constexpr size_t search_q_scratch_per_block(int q_cand) { return 4u * kQTaskCap + 4u * kSearchThreads * static_cast<size_t>(q_cand); }

constexpr uint32_t kQNoTask = 0xffffffffu;       // a queue entry nobody filled (phase B skips it)
#if defined(PPCR_Q_FALLBACK_NOINLINE)  // measured: 137 against 130 ms over the first 12 searches of the 10M-point pair
#define PPCR_Q_FALLBACK_INLINE __device__ __noinline__
#else
#define PPCR_Q_FALLBACK_INLINE __device__ __forceinline__
#endif
struct QEmit {  // phase A -> task queue
    uint32_t* tasks;
    int* n_tasks;
    uint32_t slot;
    int mine;
    int per_query;
    // the leaf children `first + c`, c a set bit of mask, of one opened node: one reservation for all of them
    __device__ __forceinline__ bool operator()(int first, int mask)
    {
        const int n = __popc(mask);
        mine += n;
        if (mine > per_query) return false;
        int t = atomicAdd(n_tasks, n);
        if (t + n > kQTaskCap) {
            // the reservation that crosses the end of the queue owns [t, cap): phase B reads every entry below
            // min(n_tasks, cap), so the part nobody will write must not keep a task of an earlier chunk
            for (; t < kQTaskCap; ++t) __stcg(tasks + t, kQNoTask);
            return false;
        }
        for (; mask; mask &= mask - 1) __stcg(tasks + t++, (slot << kQNodeBits) | static_cast<uint32_t>(first + lowest_bit(mask)));
        return true;
    }
};
struct QPush {  // phase B -> candidate list of one query (slot-major: entry c of query ql at [c * 128 + ql])
    uint32_t* cand;
    int* cnt;
    int cap;
    __device__ __forceinline__ void operator()(int j0, uint32_t pass)
    {
        int c = atomicAdd(cnt, __popc(pass));  // room for every survivor of the group at once
        while (pass && c < cap) {
            __stcg(cand + c++ * kSearchThreads, static_cast<uint32_t>(j0 + lowest_bit(pass)));
            pass &= pass - 1;
        }
    }
};
struct QCand {  // phase C: candidate c of one query -> position in the sorted target
    const uint32_t* cand;
    __device__ __forceinline__ int operator()(int c) const { return static_cast<int>(__ldcg(cand + c * kSearchThreads)); }
};

// The one-by-one fallback of k_search_q: a query whose candidate list (or the block's task queue) overflowed is searched by its
// thread alone, 127 others waiting.  What the list does hold are DISTINCT targets within the radius, so when there are m of them
// their m-th smallest distance bounds the true one -- and it is tight (a sample of a far larger candidate set), where the bound
// the query came in with was loose enough to overflow the list: the walk then opens a handful of leaves instead of the hundreds
// the loose bound touches (it was 130 us of a 350 us block on the 1M-point pair; -DPPCR_Q_PROFILE).  The register allocation of
// the whole kernel is sensitive to this code (48 registers, 16 bytes spilled): holding the m-th key in a register across the
// store loop -- here or in the heavy path -- cost the 10M-point pair, whose chunks mostly take the heap walk, 10-35 % of its
// search time; the keys are therefore read back from the column after the stores.
PPCR_Q_FALLBACK_INLINE int search_q_fallback(const TreeGeom& geom, const TreeNode* __restrict__ nodes, const float4* __restrict__ tgt_sorted,
                                             const SearchOut& out, float4 q, float r2f, float bound0, int have, int m, int flags,
                                             const uint32_t* cand, unsigned long long* heap, int* stack, int i, float* kth)
{
    float bound1 = bound0;
    if (have >= m && !(flags & 2)) {
        unsigned long long kk;
        const QCand at{cand};
        select_candidates<kSearchThreads>(tgt_sorted, at, have, m, q.x, q.y, q.z, heap, &kk);
        if (kk != kKeyInf) bound1 = fminf(bound1, key_d2(kk));
    }
    HeapList<kSearchThreads, 0> L;
    L.k = heap;
    L.init(m);
    tree_search(geom, nodes, tgt_sorted, q.x, q.y, q.z, r2f, bound1, L, stack);
    // bound1 came from whichever candidates were pushed first, so the heap's layout differs from run to run: sorted, the row is
    // the same bits every time (the m-th key then sits in the last slot)
    const bool sorted = !(flags & 1);
    if (sorted) L.sort();
    int cnt = 0;
    for (int s = 0; s < m; ++s) {
        const unsigned long long key = L.k[s * kSearchThreads];
        if (key != kKeyInf) search_store(out, i, cnt++, key);
    }
    const unsigned long long kk = sorted ? L.k[(m - 1) * kSearchThreads] : L.kth_key();
    if (kk != kKeyInf) *kth = key_d2(kk);
    return cnt;
}

__global__ void __launch_bounds__(kSearchThreads) k_search_q(const PairDev* __restrict__ pairs)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    const PairDev& P = pairs[blockIdx.y];
    PairState* st = P.state;
    if (st->phase != PH_SEARCH || st->apply_dT == 0 || !P.search_queued) return;
    float4* s_q = reinterpret_cast<float4*>(s_raw);                       // query, .w = its pruning bound
    int* s_cnt = reinterpret_cast<int*>(s_q + kSearchThreads);            // candidates pushed per query
    unsigned long long* s_heap = reinterpret_cast<unsigned long long*>(s_cnt + kSearchThreads);
    uint32_t* s_tasks = reinterpret_cast<uint32_t*>(P.q_scratch + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) *
                                                                       search_q_scratch_per_block(P.q_cand));  // this block's slab
    uint32_t* s_cand = s_tasks + kQTaskCap;
    __shared__ double s_T[12];
    __shared__ int s_chunk, s_ntasks;
    if (threadIdx.x < 12) s_T[threadIdx.x] = st->dT[threadIdx.x];
    const int m = P.m;
    const int q_cand = P.q_cand;
    const float heavy_factor = P.q_heavy;
    const int n_src = P.n_src;
    const float r2f = P.r2f;
    const TreeGeom geom = P.tree;
    const TreeNode* __restrict__ nodes = P.nodes;
    const float4* __restrict__ tgt_sorted = P.tgt_sorted;
    const SearchOut out{P.nbr_pos, P.nbr_d2, P.inv_perm, static_cast<size_t>(P.n_pad)};
    float4* __restrict__ src = P.src;
    float* __restrict__ nbr_kth = P.nbr_kth;
    int* __restrict__ nbr_cnt = P.nbr_cnt;
    const int n_chunks = (n_src + kSearchChunk - 1) / kSearchChunk;
    const float kInf = __int_as_float(0x7f800000);
    int stack[2 * kTreeStack];
    int cnt_total = 0;
#if defined(PPCR_Q_PROFILE)
    long long t_phase[4] = {0, 0, 0, 0}, t_mark = clock64();
    unsigned long long g_begin;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_begin));
    int n_chunks_done = 0, n_heavy_chunks = 0;
    long long n_task_sum = 0, n_fall = 0, n_fall_task = 0;
#define PPCR_Q_MARK(ph) { const long long now = clock64(); t_phase[ph] += now - t_mark; t_mark = now; }
#else
#define PPCR_Q_MARK(ph)
#endif
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            s_chunk = atomicAdd(&st->search_cursor, 1);
            s_ntasks = 0;
        }
        s_cnt[threadIdx.x] = 0;
        __syncthreads();
        PPCR_Q_MARK(3)
        const int chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const int i = chunk * kSearchChunk + threadIdx.x;
        const bool valid = i < n_src;
        // ---- A: the query, its bound, its leaves ----
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        float bound0 = r2f;
        bool fallback = false, heavy = false, dead = false;  // dead: a non-finite query, no neighbours
        if (valid) {
            q = src[i];
            const double x = q.x, y = q.y, z = q.z;
            const float nx = transform_row(s_T, x, y, z), ny = transform_row(s_T + 4, x, y, z),
                        nz = transform_row(s_T + 8, x, y, z);
            const float prev = nbr_kth[i];
            q.x = nx;
            q.y = ny;
            q.z = nz;
            src[i] = q;
            dead = !finite_query(q);
            if (prev < kInf && !dead) {  // a saturated row: the farthest of its previous neighbours bounds the new m-th distance
                const int* __restrict__ pp = out.nbr_pos + i;
                float far2 = 0.f;
                int k = 0;
                for (; k + 5 <= m; k += 5) {
                    int pos[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) pos[u] = pp[static_cast<size_t>(k + u) * out.n_pad];
                    float4 t[5];
#pragma unroll
                    for (int u = 0; u < 5; ++u) t[u] = load_point(tgt_sorted + pos[u]);
#pragma unroll
                    for (int u = 0; u < 5; ++u) far2 = fmaxf(far2, dist2_exact(nx, ny, nz, t[u].x, t[u].y, t[u].z));
                }
                for (; k < m; ++k) {
                    const float4 t = load_point(tgt_sorted + pp[static_cast<size_t>(k) * out.n_pad]);
                    far2 = fmaxf(far2, dist2_exact(nx, ny, nz, t.x, t.y, t.z));
                }
                bound0 = fminf(r2f, far2);
                // On a surface the targets within the bound number about m (bound / previous m-th distance)^2: a query is
                // "heavy" when that would fill most of its candidate list.  It happens when the cloud moves by more than
                // the neighbour spacing (the 10M-point pair: spacing 1 cm, increments of 2 cm).
                heavy = static_cast<float>(m) * far2 > heavy_factor * static_cast<float>(q_cand) * prev;
            }
        }
        // A chunk made mostly of heavy queries is searched the way k_search does it -- every thread walks with its heap and
        // lets the bound shrink as candidates arrive -- because the fixed bound would push them all through the fallback.
        if (__syncthreads_count(heavy) * 2 > kSearchThreads) {
#if defined(PPCR_Q_PROFILE)
            ++n_heavy_chunks;
#endif
            if (dead) {
                nbr_cnt[i] = 0;
                nbr_kth[i] = kInf;
            } else if (valid) {
                HeapList<kSearchThreads, 0> L;
                L.k = s_heap + threadIdx.x;
                L.init(m);
                tree_search(geom, nodes, tgt_sorted, q.x, q.y, q.z, r2f, bound0, L, stack);
                int cnt = 0;
                for (int s = 0; s < m; ++s) {
                    const unsigned long long key = L.k[s * kSearchThreads];
                    if (key != kKeyInf) search_store(out, i, cnt++, key);
                }
                nbr_cnt[i] = cnt;
                nbr_kth[i] = L.kth_key() != kKeyInf ? key_d2(L.kth_key()) : kInf;
                cnt_total += cnt;
                if (cnt >= P.overflow_at) st->row_overflow = 1;
            }
            continue;
        }
        if (valid && !dead) {
            s_q[threadIdx.x] = make_float4(q.x, q.y, q.z, candidate_limit(bound0, r2f));
            QEmit emit{s_tasks, &s_ntasks, threadIdx.x, 0, P.q_leaves};
            fallback = !tree_collect_leaves(geom, nodes, q.x, q.y, q.z, bound0, emit, stack);
        }
        __syncthreads();
        PPCR_Q_MARK(0)
        // ---- B: every queued leaf against its query ----
        const int n_tasks = min(s_ntasks, kQTaskCap);
        for (int t = threadIdx.x; t < n_tasks; t += kSearchThreads) {
            const uint32_t task = __ldcg(s_tasks + t);
            if (task == kQNoTask) continue;
            const uint32_t ql = task >> kQNodeBits;
            const float4 qq = s_q[ql];
            QPush push{s_cand + ql, s_cnt + ql, q_cand};
            leaf_candidates(nodes, tgt_sorted, static_cast<int>(task & kQLowMask), qq.x, qq.y, qq.z, qq.w, push);
        }
        __syncthreads();
        PPCR_Q_MARK(1)
#if defined(PPCR_Q_PROFILE)
        ++n_chunks_done;
        n_task_sum += n_tasks;
        {
            const int n_cold = __syncthreads_count(valid && bound0 >= r2f);
            const int n_big = __syncthreads_count(valid && bound0 > 0.04f);
            if (false && threadIdx.x == 0 && (chunk % 7) == 0)
                printf("[k_search_q overflow] chunk %d tasks %d, queries with bound = r2: %d, with bound > (0.2 m)^2: %d, q0 = (%.2f %.2f %.2f) bound0 %.4f\n",
                       chunk, s_ntasks, n_cold, n_big, q.x, q.y, q.z, bound0);
        }
#endif
        // ---- C: the m best of each query's candidates ----
        if (valid) {
            const int n_c = s_cnt[threadIdx.x];
#if defined(PPCR_Q_PROFILE)
            n_fall_task += fallback ? 1 : 0;
#endif
            if (n_c > q_cand) fallback = true;
            int cnt = 0;
            float kth = kInf;
            if (!fallback) {
                unsigned long long kk;
                const QCand at{s_cand + threadIdx.x};
                const int n = select_candidates<kSearchThreads>(tgt_sorted, at, n_c, m, q.x, q.y, q.z, s_heap + threadIdx.x, &kk);
                for (int s = 0; s < n; ++s) search_store(out, i, cnt++, s_heap[threadIdx.x + s * kSearchThreads]);
                if (kk != kKeyInf) kth = key_d2(kk);
            } else {
                cnt = search_q_fallback(geom, nodes, tgt_sorted, out, q, r2f, bound0, min(n_c, q_cand), m, P.q_flags, s_cand + threadIdx.x,
                                        s_heap + threadIdx.x, stack, i, &kth);
            }
            nbr_cnt[i] = cnt;
            nbr_kth[i] = kth;
            cnt_total += cnt;
            if (cnt >= P.overflow_at) st->row_overflow = 1;
#if defined(PPCR_Q_PROFILE)
            n_fall += fallback ? 1 : 0;
#endif
        }
#if defined(PPCR_Q_PROFILE)
        __syncthreads();
        PPCR_Q_MARK(2)
#endif
    }
#if defined(PPCR_Q_PROFILE)
    __shared__ int s_prof[2];
    if (threadIdx.x == 0) s_prof[0] = s_prof[1] = 0;
    __syncthreads();
    atomicAdd(&s_prof[0], static_cast<int>(n_fall));
    atomicAdd(&s_prof[1], static_cast<int>(n_fall_task));
    __syncthreads();
    unsigned long long g_end;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
    if (threadIdx.x == 0 && ((blockIdx.x % 197) == 0 || s_prof[0] > 0 || n_heavy_chunks > 0 || (g_end - g_begin) > 450000ull))
        printf("[k_search_q block %d] cycles A %lld B %lld C %lld idle %lld; tasks %lld; fallbacks %d (task-queue overflow %d); chunks %d heavy %d; "
               "globaltimer begin %llu end %llu (%.1f us)\n", blockIdx.x, t_phase[0], t_phase[1], t_phase[2], t_phase[3], n_task_sum, s_prof[0],
               s_prof[1], n_chunks_done, n_heavy_chunks, g_begin % 100000000ull, g_end % 100000000ull, (g_end - g_begin) * 1e-3);
#endif
    for (int o = 16; o > 0; o >>= 1) cnt_total += __shfl_xor_sync(kFull, cnt_total, o);
    if ((threadIdx.x & 31) == 0 && cnt_total)
        atomicAdd(reinterpret_cast<unsigned long long*>(&st->K), static_cast<unsigned long long>(cnt_total));
}

// ------------------------------------------------------------------------------------------------------------
// weights + normal-equation moments
// ------------------------------------------------------------------------------------------------------------

PPCR_HD constexpr int eval_threads(bool fast) { return fast ? kEvalFastThreads : kEvalThreads; }

// one staged tile of the fast evaluation: [m][128] positions, [128] counts, [128] source points
PPCR_HD constexpr size_t eval_stage_bytes(int m) { return static_cast<size_t>(kEvalFastThreads) * (4u * static_cast<size_t>(m) + 20u); }
// Build-time experiment (-DPPCR_EVAL_ASYNC=1): up to kEvalAsyncMaxM neighbours per row the gathered target points are staged too,
// two buffers of [m][128] float4 filled by per-thread asynchronous copies (cp.async) one tile ahead of the arithmetic.
#ifndef PPCR_EVAL_ASYNC
#define PPCR_EVAL_ASYNC 0  // measured on the 1M-point pair: 81 us per evaluation against 63 (three blocks per SM, and the 16-byte copies take the same L1 path as the loads they replace)
#endif
constexpr int kEvalAsyncMaxM = PPCR_EVAL_ASYNC ? 12 : 0;
#ifndef PPCR_EVAL_ASYNC_BLOCKS
#define PPCR_EVAL_ASYNC_BLOCKS 3
#endif
constexpr int kEvalAsyncBlocks = PPCR_EVAL_ASYNC_BLOCKS;  // resident blocks per SM that fit with the point buffers at m = 10..12
PPCR_HD constexpr bool eval_async(int m) { return m <= kEvalAsyncMaxM; }
PPCR_HD constexpr size_t eval_points_bytes(int m) { return static_cast<size_t>(kEvalFastThreads) * 16u * static_cast<size_t>(m); }
PPCR_HD constexpr size_t eval_fast_smem(int m)
{
    // (the stages are a multiple of 16 bytes; the barriers sit behind the point buffers)
    return kEvalStages * eval_stage_bytes(m) + (eval_async(m) ? 2u * eval_points_bytes(m) : 0u) + 16u * kEvalStages;
}

// ---- asynchronous bulk copies (TMA, 1-D) completing on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ double ld_volatile_f64(const double* p)
{
    return *reinterpret_cast<const volatile double*>(p);
}

// The scalar LM / outer-loop state machine (ppcr_lm.h), one thread.  Kept out of line so that its register appetite
// (a 7x7 Cholesky and the moment expansion, all float64) does not set the register count of the streaming part.
// It works on shared-memory copies of the pair state and configuration: the state machine is a long chain of
// dependent reads and writes of those fields, and every one of them would otherwise be a global-memory round trip.
//
// Cooperative form of ppcr_lm.h::controller_tick for the first warp of the controller block: lane 0 runs the scalar
// decisions, the 7x7 Cholesky factorisation runs one row per lane (same operation order per entry as the serial
// solve_damped, hence the same bits).  The moment expansion has already been done by the whole block (ev).
struct CtrlShared {
    Expanded ev;
    double N[kExpandN];
    QuatFrame frame;
    double L[kNP][kNP];
    double D[kNP], y[kNP];
    int flag;
};

__device__ __noinline__ int ctrl_begin(PairState* st, const Config* cfg, const Expanded* ev)
{
    st->evals += 1;
    return controller_begin(st, cfg, *ev) ? 1 : 0;
}
__device__ __noinline__ int ctrl_prepare(PairState* st, const Config* cfg, double* D) { return step_prepare(st, cfg, D) ? 1 : 0; }
__device__ __noinline__ int ctrl_complete(PairState* st, const Config* cfg, int valid, const double* y)
{
    return step_complete(st, cfg, valid != 0, y);
}
__device__ __noinline__ void ctrl_finish(PairState* st, const Config* cfg, double* history, IterStats* stats, int max_hist)
{
    outer_finish(st, cfg, history, stats, max_hist);
}

// (Hs + diag(D2)) y = gs, rows of the factor spread over lanes 0..6 (same operation order per entry as the serial
// solve_damped: inverse diagonal, no division); returns validity in every lane
__device__ __forceinline__ int warp_solve_damped(const PairState* st, CtrlShared* sh, int lane)
{
    if (lane < kNP)
        for (int c = 0; c < kNP; ++c) sh->L[lane][c] = st->Hs[lane * kNP + c] + (lane == c ? sh->D[lane] : 0.0);
    if (lane == 0) sh->flag = 1;
    __syncwarp();
    for (int c = 0; c < kNP; ++c) {
        if (lane == c) {
            double d = sh->L[c][c];
            for (int k = 0; k < c; ++k) d -= sh->L[c][k] * sh->L[c][k];
            if (!(d > 0.0)) sh->flag = 0;
            else sh->L[c][c] = inv_sqrt(d);
        }
        __syncwarp();
        if (!sh->flag) break;
        if (lane > c && lane < kNP) {
            double s = sh->L[lane][c];
            for (int k = 0; k < c; ++k) s -= sh->L[lane][k] * sh->L[c][k];
            sh->L[lane][c] = s * sh->L[c][c];
        }
        __syncwarp();
    }
    if (lane == 0 && sh->flag) {
        double z[kNP];
        for (int r = 0; r < kNP; ++r) {
            double s = st->gs[r];
            for (int k = 0; k < r; ++k) s -= sh->L[r][k] * z[k];
            z[r] = s * sh->L[r][r];
        }
        for (int r = kNP - 1; r >= 0; --r) {
            double s = z[r];
            for (int k = r + 1; k < kNP; ++k) s -= sh->L[k][r] * sh->y[k];
            sh->y[r] = s * sh->L[r][r];
        }
    }
    __syncwarp();
    return sh->flag;
}

__device__ __forceinline__ void warp_controller(PairState* st, const Config* cfg, CtrlShared* sh, double* history,
                                                IterStats* stats, int max_hist, int max_ticks, int lane)
{
    int go = 0;
    if (lane == 0) go = ctrl_begin(st, cfg, &sh->ev);
    go = __shfl_sync(kFull, go, 0);
    while (go) {
        int ok = 0;
        if (lane == 0) ok = ctrl_prepare(st, cfg, sh->D);
        ok = __shfl_sync(kFull, ok, 0);  // also orders lane 0's writes of D before the other lanes' reads
        if (!ok) {
            go = 0;
            break;
        }
        __syncwarp();
        const int valid = warp_solve_damped(st, sh, lane);
        int r = 0;
        if (lane == 0) r = ctrl_complete(st, cfg, valid, sh->y);
        r = __shfl_sync(kFull, r, 0);
        if (r == 0) break;
        if (r == 2) go = 0;
    }
    if (lane == 0) {
        if (!go) ctrl_finish(st, cfg, history, stats, max_hist);
        if (st->ticks >= max_ticks && st->phase != PH_DONE) {  // never spin forever on the device
            st->error = 1;
            st->phase = PH_DONE;
        }
    }
}

static_assert(sizeof(PairState) % 8 == 0 && sizeof(Config) % 8 == 0, "copied as 64-bit words");

// The float32 row loop of k_evalctl.  A block walks tiles of 128 consecutive rows (tile = blockIdx.x, += gridDim.x).
// What a tile needs from the planes -- m slot segments of positions (512 contiguous bytes each), the counts, the source
// points -- is brought into shared memory by 1-D bulk copies (TMA) that complete on an mbarrier, kEvalStages tiles
// ahead of the arithmetic: the latency of the planes never sits on a thread's critical path, and no register is spent
// on prefetching.  The 16-byte target points are then gathered from the Morton-sorted target (L2 / L1 resident: the
// neighbours of neighbouring rows are neighbours in that order), kU at a time.  The 24 float64 moments of a thread
// stay in registers.
struct EvalStage {
    unsigned char* base;   // this block's staging area
    uint32_t bar0;         // shared-space address of the first mbarrier
    size_t stage_bytes;
    int m;
    float4* points;        // two buffers of [m][128] gathered target points (asynchronous variant), else null
};

__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src_global)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_global) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// called by the whole first warp: lane 0 arms the barrier, then the lanes issue the m + 2 copies between them
__device__ __forceinline__ void eval_issue_tile(const PairDev& P, const EvalStage& S, int stage, int tile)
{
    const int lane = threadIdx.x & 31;
    const uint32_t bar = S.bar0 + 8u * stage;
    const uint32_t dst = smem_u32(S.base + stage * S.stage_bytes);
    const size_t row0 = static_cast<size_t>(tile) * kEvalFastThreads;
    constexpr uint32_t seg = kEvalFastThreads * 4u;
    if (lane == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(bar, seg * static_cast<uint32_t>(S.m) + seg + kEvalFastThreads * 16u);
    }
    __syncwarp();
    for (int k = lane; k < S.m; k += 32) bulk_g2s(dst + seg * k, P.nbr_pos + static_cast<size_t>(k) * P.n_pad + row0, seg, bar);
    if (lane == (S.m & 31)) bulk_g2s(dst + seg * S.m, P.nbr_cnt + row0, seg, bar);
    if (lane == ((S.m + 1) & 31)) bulk_g2s(dst + seg * S.m + seg, P.src + row0, kEvalFastThreads * 16u, bar);
}

// Sum over the 32 lanes of a warp of 32 values per lane, transposed: lane L returns sum_over_lanes v[L].  Recursive halving --
// a lane keeps one half of its values and trades the other half with its partner, 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of
// 32 x 5 -- in a fixed tree, so the float32 result is the same on every run.  All 32 lanes must call it.
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int half = 16; half >= 1; half >>= 1) {
        const bool upper = (lane & half) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float keep = upper ? v[i + half] : v[i];
            const float send = upper ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(kFull, send, half);
        }
    }
    return v[0];
}

template <int WM, bool SAME, bool ASYNC>
__device__ __forceinline__ void eval_rows_fast(const PairDev& P, const EvalStage& S, const Pose& pe, const Pose& pw,
                                               const WeightCfg& wc, double* __restrict__ racc)
{
    // PPCR_EVAL_LANEACC: racc is ONE double, the lane's accumulator of moment `lane` (lanes 24..31 add zeros); else the thread's 24
    constexpr int kU = PPCR_EVAL_BATCH;
    const int n_tiles = (P.n_src + kEvalFastThreads - 1) / kEvalFastThreads;
    const int tile_step = P.n_eval_blocks;
    const float4* __restrict__ table = P.tgt_sorted;
    const int n_src = P.n_src;
    // ASYNC: this thread's row of tile `t` (staged in stage slot `jt % kEvalStages`, once its barrier has completed) names its
    // target points; fetch them into point buffer `jt & 1` without waiting.  Every thread only ever reads the points it
    // fetched itself, so completion is the thread's own cp.async group, no block-wide barrier.
    auto fetch_points = [&](int t, int jt) {
        if (t < n_tiles) {
            const int st = jt % kEvalStages;
            while (!mbar_try_wait(S.bar0 + 8u * st, static_cast<uint32_t>(jt / kEvalStages) & 1u)) {
            }
            const unsigned char* nb = S.base + st * S.stage_bytes;
            const int* n_pos = reinterpret_cast<const int*>(nb) + threadIdx.x;
            int n_cnt = reinterpret_cast<const int*>(nb + static_cast<size_t>(kEvalFastThreads) * 4u * S.m)[threadIdx.x];
            if (t * kEvalFastThreads + static_cast<int>(threadIdx.x) >= n_src) n_cnt = 0;
            float4* dst = S.points + static_cast<size_t>(jt & 1) * kEvalFastThreads * S.m + threadIdx.x;
            for (int k = 0; k < n_cnt; ++k) cp_async16(dst + k * kEvalFastThreads, table + n_pos[k * kEvalFastThreads]);
        }
        cp_async_commit();  // (an empty group when there is no such tile: the group count stays in step)
    };
    int j = 0;
    if (ASYNC) fetch_points(blockIdx.x, 0);
    for (int tile = blockIdx.x; tile < n_tiles; tile += tile_step, ++j) {
        const int stage = j % kEvalStages;
        const uint32_t parity = static_cast<uint32_t>(j / kEvalStages) & 1u;
        if (ASYNC) {
            fetch_points(tile + tile_step, j + 1);
            cp_async_wait_1();  // everything but the group just committed has landed: this tile's points are here
        } else {
            while (!mbar_try_wait(S.bar0 + 8u * stage, parity)) {
            }
        }
        const float4* s_pts = ASYNC ? S.points + static_cast<size_t>(j & 1) * kEvalFastThreads * S.m + threadIdx.x : nullptr;
        const unsigned char* sb = S.base + stage * S.stage_bytes;
        const int* s_pos = reinterpret_cast<const int*>(sb) + threadIdx.x;
        int cnt = reinterpret_cast<const int*>(sb + static_cast<size_t>(kEvalFastThreads) * 4u * S.m)[threadIdx.x];
        if (tile * kEvalFastThreads + static_cast<int>(threadIdx.x) >= n_src) cnt = 0;
#if PPCR_EVAL_LANEACC
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = 0.f;
#endif
        if (cnt > 0) {
            const float4 sp = reinterpret_cast<const float4*>(sb + static_cast<size_t>(kEvalFastThreads) * 4u * (S.m + 1))[threadIdx.x];
            const double sx = sp.x, sy = sp.y, sz = sp.z;
            double pte[3];
            apply_pose(pe, sx, sy, sz, pte);
            PointHL he;
            split_point(pte, &he);
            float dw[3] = {0.f, 0.f, 0.f};
            if (!SAME) {
                double ptw[3];
                apply_pose(pw, sx, sy, sz, ptw);
                pose_delta(pte, ptw, dw);
            }
            RowAccF row;
            rowf_begin(&row);
            int k0 = 0;
            for (; k0 + kU <= cnt; k0 += kU) {
                float4 y[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    y[u] = ASYNC ? s_pts[(k0 + u) * kEvalFastThreads] : __ldg(table + s_pos[(k0 + u) * kEvalFastThreads]);
#pragma unroll
                for (int u = 0; u < kU; ++u) rowf_add_t<WM, SAME>(&row, wc, y[u].x, y[u].y, y[u].z, he, dw);
            }
            if (k0 < cnt) {
                float4 y[kU - 1];
#pragma unroll
                for (int u = 0; u < kU - 1; ++u)
                    if (k0 + u < cnt)
                        y[u] = ASYNC ? s_pts[(k0 + u) * kEvalFastThreads] : __ldg(table + s_pos[(k0 + u) * kEvalFastThreads]);
#pragma unroll
                for (int u = 0; u < kU - 1; ++u)
                    if (k0 + u < cnt) rowf_add_t<WM, SAME>(&row, wc, y[u].x, y[u].y, y[u].z, he, dw);
            }
#if PPCR_EVAL_LANEACC
            rowf_end_f(&row, sp.x, sp.y, sp.z, v);
#else
            rowf_end_s<1>(&row, sx, sy, sz, racc);
#endif
            if (!ASYNC && P.dump_w) {
                // Parity dump (ppcr_weights_normal_eq): the weight of every correspondence of this row, from the row statistics
                // just folded into the moments and the same staged positions, residual arithmetic and weight terms.
                double* __restrict__ o = P.dump_w + static_cast<size_t>(tile) * kEvalFastThreads + threadIdx.x;
                for (int k = 0; k < cnt; ++k) {
                    const float4 y = __ldg(table + s_pos[k * kEvalFastThreads]);
                    float wx = residual_hl(y.x, he.hi[0], he.lo[0]), wy = residual_hl(y.y, he.hi[1], he.lo[1]),
                          wz = residual_hl(y.z, he.hi[2], he.lo[2]);
                    if (!SAME) {
                        wx += dw[0];
                        wy += dw[1];
                        wz += dw[2];
                    }
                    const float r2w = wx * wx + wy * wy + wz * wz;
                    float w;
                    if (WM == WM_GAUSS) {
                        w = f_exp(-0.5f * r2w - row.m) / row.a0;
                    } else {
                        float u, ue;
                        t_terms<WM>(wc, r2w, &u, &ue);
                        w = ue / row.a0;
                    }
                    o[static_cast<size_t>(k) * P.n_pad] = static_cast<double>(w);
                }
            }
        }
#if PPCR_EVAL_LANEACC
        // the 32 rows of the warp, added in float32 by a fixed tree; only the sums of whole warps meet the float64 accumulator
        racc[0] += static_cast<double>(warp_transpose_sum(v));
#endif
        __syncthreads();  // every thread is done with this stage: refill it with the tile kEvalStages ahead
        if (threadIdx.x < 32) {
            const int next = tile + kEvalStages * tile_step;
            if (next < n_tiles) eval_issue_tile(P, S, stage, next);
        }
    }
}

struct LoopCtl {   // one per engine
    int active;      // any pair still running (read back by the host-stepped driver)
    int pairs_done;  // pairs whose controller has finished this tick
};

// The per-iteration kernel: weights + moments over the association (every block), then -- in the block that
// publishes its partial sums last -- the fixed-order reduction, the cross-rank exchange (sharded mode), the LM /
// outer-loop controller and the loop condition of the tick graph.  One launch per LM iteration.
template <bool kFast>
__global__ void __launch_bounds__(eval_threads(kFast), kFast ? kEvalFastBlocks : PPCR_EVAL_MIN_BLOCKS)
    k_evalctl(const PairDev* __restrict__ pairs, int n_pairs, LoopCtl* __restrict__ loop, cudaGraphConditionalHandle cond,
              cudaGraphConditionalHandle cond_search, int use_cond, int max_ticks)
{
    constexpr int NT = eval_threads(kFast);
    const PairDev& P = pairs[blockIdx.y];
    PairState* st = P.state;
    __shared__ Pose s_pe, s_pw;
    __shared__ double s_red[NT / 32][kNSum];
    static_assert(NT / 32 >= kFoldChains && NT >= kNSum * kFoldChains, "the reduction needs kFoldChains rows of s_red");
    __shared__ double s_sum[kMailDoubles];
    __shared__ int s_flag;
    const bool live = st->phase != PH_DONE;
    if (live) {
    if (static_cast<int>(blockIdx.x) >= P.n_eval_blocks) return;
    extern __shared__ __align__(128) unsigned char s_dyn[];
    EvalStage stg;
    if constexpr (kFast) {
        // staging area + one mbarrier per stage; the first kEvalStages tiles are requested before anything else
        stg.m = P.m;
        stg.stage_bytes = eval_stage_bytes(stg.m);
        stg.base = s_dyn;
        // the asynchronous variant needs the smem the host sized for the engine's max_neighbours: use it when this pair's m fits
        stg.points = (eval_async(stg.m) && (use_cond & 8)) ? reinterpret_cast<float4*>(s_dyn + kEvalStages * stg.stage_bytes) : nullptr;
        stg.bar0 = smem_u32(s_dyn + kEvalStages * stg.stage_bytes + (stg.points ? 2u * eval_points_bytes(stg.m) : 0u));
        if (threadIdx.x < 32) {
            if (threadIdx.x == 0) {
                for (int k = 0; k < kEvalStages; ++k) mbar_init(stg.bar0 + 8u * k, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncwarp();
            const int n_tiles = (P.n_src + kEvalFastThreads - 1) / kEvalFastThreads;
            for (int k = 0; k < kEvalStages; ++k) {
                const int tile = blockIdx.x + k * P.n_eval_blocks;
                if (tile < n_tiles) eval_issue_tile(P, stg, k, tile);
            }
        }
    }
    if (threadIdx.x < 12) {
        const double* pe = reinterpret_cast<const double*>(&st->pose_e);
        const double* pw = reinterpret_cast<const double*>(&st->pose_w);
        reinterpret_cast<double*>(&s_pe)[threadIdx.x] = pe[threadIdx.x];
        reinterpret_cast<double*>(&s_pw)[threadIdx.x] = pw[threadIdx.x];
    }
    __syncthreads();
    const Pose& pe = s_pe;
    const Pose& pw = s_pw;
    const WeightCfg wc = P.wcfg;
    // fast path: the lane's accumulator (PPCR_EVAL_LANEACC) or the thread's 24 moments, in registers
    double racc[(kFast && PPCR_EVAL_LANEACC) ? 1 : kNSum];
#pragma unroll
    for (int k = 0; k < ((kFast && PPCR_EVAL_LANEACC) ? 1 : kNSum); ++k) racc[k] = 0.0;
    if constexpr (kFast) {
        // pose_w == pose_e on the first evaluation of every outer iteration: one residual serves both uses
        bool same = true;
#pragma unroll
        for (int k = 0; k < 12; ++k)
            same = same && (reinterpret_cast<const double*>(&s_pe)[k] == reinterpret_cast<const double*>(&s_pw)[k]);
        // one instantiation of the row loop per (weight model, same pose): both are uniform over the launch
        // ... and one per gather variant (uniform too: it follows from m)
#if PPCR_EVAL_ASYNC
#define PPCR_ROWS(WM, SAME)                                                     \
    if (stg.points) eval_rows_fast<WM, SAME, true>(P, stg, pe, pw, wc, racc);   \
    else eval_rows_fast<WM, SAME, false>(P, stg, pe, pw, wc, racc);             \
    break;
#else
#define PPCR_ROWS(WM, SAME)                                      \
    eval_rows_fast<WM, SAME, false>(P, stg, pe, pw, wc, racc);   \
    break;
#endif
        switch (weight_mode(wc) * 2 + (same ? 1 : 0)) {
            case WM_T_H4 * 2: PPCR_ROWS(WM_T_H4, false)
            case WM_T_H4 * 2 + 1: PPCR_ROWS(WM_T_H4, true)
            case WM_T_INT * 2: PPCR_ROWS(WM_T_INT, false)
            case WM_T_INT * 2 + 1: PPCR_ROWS(WM_T_INT, true)
            case WM_T_REAL * 2: PPCR_ROWS(WM_T_REAL, false)
            case WM_T_REAL * 2 + 1: PPCR_ROWS(WM_T_REAL, true)
            case WM_GAUSS * 2: PPCR_ROWS(WM_GAUSS, false)
            default: PPCR_ROWS(WM_GAUSS, true)
        }
#undef PPCR_ROWS
    } else {
        // float64 path: the 24 accumulators of a thread live in shared memory (one column per thread, conflict-free)
        double* acc = reinterpret_cast<double*>(s_dyn) + threadIdx.x;
#pragma unroll
        for (int k = 0; k < kNSum; ++k) acc[k * kEvalThreads] = 0.0;
        const int stride = P.n_eval_blocks * kEvalThreads;
        const size_t n_pad = P.n_pad;
        for (int i = blockIdx.x * kEvalThreads + threadIdx.x; i < P.n_src; i += stride) {
            const int cnt = P.nbr_cnt[i];
            if (cnt == 0) continue;
            const float4 sp = P.src[i];
            const double sx = sp.x, sy = sp.y, sz = sp.z;
            double pte[3], ptw[3];
            apply_pose(pe, sx, sy, sz, pte);
            apply_pose(pw, sx, sy, sz, ptw);
            RowAcc row;
            row_begin(&row);
            for (int k = 0; k < cnt; ++k) {
                const size_t o = static_cast<size_t>(k) * n_pad + i;
                const float4 y = __ldg(P.tgt_sorted + __ldg(P.nbr_pos + o));
                row_add<false>(&row, wc, y.x, y.y, y.z, pte, ptw);
            }
            row_end_s<kEvalThreads>(&row, sx, sy, sz, acc);
            if (P.dump_w) {  // parity dump: the weights this row entered the moments with
                for (int k = 0; k < cnt; ++k) {
                    const size_t o = static_cast<size_t>(k) * n_pad + i;
                    const float4 y = __ldg(P.tgt_sorted + __ldg(P.nbr_pos + o));
                    P.dump_w[o] = finished_weight<false>(&row, wc, y.x, y.y, y.z, ptw);
                }
            }
        }
#pragma unroll
        for (int k = 0; k < kNSum; ++k) racc[k] = acc[k * kEvalThreads];
    }
    // fixed-shape reduction: xor-shuffle tree inside the warp, then warps in index order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if constexpr (kFast && PPCR_EVAL_LANEACC) {
        if (lane < kNSum) s_red[warp][lane] = racc[0];  // lane L already holds the warp's sum of moment L
    } else {
#pragma unroll
        for (int k = 0; k < kNSum; ++k) {
            double v = racc[k];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
            if (lane == 0) s_red[warp][k] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < kNSum) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) v += s_red[w][threadIdx.x];
        P.partials[static_cast<size_t>(blockIdx.x) * kNSum + threadIdx.x] = v;
    }
    // ---- two-level, fixed-order reduction; the block that finishes it runs the controller ----------------------
    // Level 1: the last block of every group of kFoldGroup consecutive blocks to publish adds the group's partial sums
    // in block order.  Level 2: the last group to finish adds the group sums in kFoldChains interleaved chains.  Both
    // orders are fixed, so the result does not depend on which blocks happen to be last; two short rounds of L2 reads
    // replace one long serial pass over every block's partial sums.
    const int group = blockIdx.x / kFoldGroup;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int members = min(kFoldGroup, P.n_eval_blocks - group * kFoldGroup);
        s_flag = (atomicAdd(&P.group_ticket[group], 1) == members - 1);
    }
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    if (threadIdx.x < kNSum) {
        const int b0 = group * kFoldGroup;
        const int members = min(kFoldGroup, P.n_eval_blocks - b0);
        const double* base = P.partials + static_cast<size_t>(b0) * kNSum + threadIdx.x;
        double t[kFoldGroup];
#pragma unroll
        for (int u = 0; u < kFoldGroup; ++u) t[u] = u < members ? __ldcg(base + static_cast<size_t>(u) * kNSum) : 0.0;
        double v = 0.0;
#pragma unroll
        for (int u = 0; u < kFoldGroup; ++u)
            if (u < members) v += t[u];
        P.group_partials[static_cast<size_t>(group) * kNSum + threadIdx.x] = v;
    }
    if (threadIdx.x == 0) P.group_ticket[group] = 0;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_flag = (atomicAdd(&st->eval_ticket, 1) == P.n_eval_groups - 1);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    if (threadIdx.x == 0) st->eval_ticket = 0;
    if (use_cond & 2) return;  // timing probe (ppcr_time_kernel): the streaming part alone
    {
        // chain q adds groups q, q + kFoldChains, ... in order (all loads of a chain issued together), then the chains in order
        if (threadIdx.x < kNSum * kFoldChains) {
            const int k = threadIdx.x % kNSum, q = threadIdx.x / kNSum;
            const double* base = P.group_partials + k;
            double v = 0.0;
            int g = q;
            for (; g + 7 * kFoldChains < P.n_eval_groups; g += 8 * kFoldChains) {
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = __ldcg(base + static_cast<size_t>(g + u * kFoldChains) * kNSum);
#pragma unroll
                for (int u = 0; u < 8; ++u) v += t[u];
            }
            for (; g < P.n_eval_groups; g += kFoldChains) v += __ldcg(base + static_cast<size_t>(g) * kNSum);
            s_red[q][k] = v;
        }
        __syncthreads();
        if (threadIdx.x < kNSum) {
            double t = 0.0;
#pragma unroll
            for (int q = 0; q < kFoldChains; ++q) t += s_red[q][threadIdx.x];
            s_sum[threadIdx.x] = t;
        }
        __syncthreads();
    }

    if (P.world > 1) {
        // Sharded pair: every rank adds the other ranks' moments (and association sizes) in rank order, so all
        // ranks hold bit-identical sums and take identical decisions.  The exchange is a one-shot all-gather
        // written straight into the peers' mailboxes over NVLink; a sequence stamp doubles as the ready flag.
        const int seq = st->ticks + 1;
        const long long t_exchange = clock64();
        if (threadIdx.x < kMailDoubles) {
            double payload = 0.0;
            if (threadIdx.x < kNSum) payload = s_sum[threadIdx.x];
            else if (threadIdx.x == kNSum) payload = static_cast<double>(st->K);
            // two alternating mail slots per rank: a fast rank may be one tick ahead, never two
            const int parity = seq & 1;
            for (int r = 0; r < P.world; ++r) {
                double* dst = P.peer_mailbox[r] + (static_cast<size_t>(parity) * P.world + P.rank) * kMailDoubles;
                if (threadIdx.x != kMailDoubles - 1) dst[threadIdx.x] = payload;
            }
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            const int parity = seq & 1;
            for (int r = 0; r < P.world; ++r) {
                double* dst = P.peer_mailbox[r] + (static_cast<size_t>(parity) * P.world + P.rank) * kMailDoubles;
                *reinterpret_cast<volatile double*>(dst + kMailDoubles - 1) = P.mail_base + static_cast<double>(seq);
            }
            __threadfence_system();
        }
        __shared__ int s_timeout;
        if (threadIdx.x == 0) s_timeout = 0;
        __syncthreads();
        if (threadIdx.x < P.world) {
            const int parity = seq & 1;
            const double* src = P.mailbox + (static_cast<size_t>(parity) * P.world + threadIdx.x) * kMailDoubles;
            const long long t0 = clock64();
            while (ld_volatile_f64(src + kMailDoubles - 1) != P.mail_base + static_cast<double>(seq)) {
                if (clock64() - t0 > P.spin_limit) {
                    s_timeout = 1;
                    break;
                }
            }
        }
        __threadfence_system();
        __syncthreads();
        if (s_timeout) {
            if (threadIdx.x == 0) {
                st->error = PH_DONE + 100;
                st->phase = PH_DONE;
            }
        } else {
            if (threadIdx.x <= kNSum) {
                const int parity = seq & 1;
                double t = 0.0;
                for (int r = 0; r < P.world; ++r)
                    t += ld_volatile_f64(P.mailbox + (static_cast<size_t>(parity) * P.world + r) * kMailDoubles + threadIdx.x);
                s_sum[threadIdx.x] = t;
            }
            __syncthreads();
            if (threadIdx.x == 0 && st->phase == PH_SEARCH) st->K = static_cast<int64_t>(s_sum[kNSum]);
        }
        if (threadIdx.x == 0) st->exchange_cycles += clock64() - t_exchange;
        __syncthreads();
    }

    {
        __shared__ PairState s_state;
        __shared__ Config s_cfg;
        constexpr int kStateWords = sizeof(PairState) / 8, kCfgWords = sizeof(Config) / 8;
        for (int k = threadIdx.x; k < kStateWords; k += NT)
            reinterpret_cast<unsigned long long*>(&s_state)[k] = __ldcg(reinterpret_cast<const unsigned long long*>(st) + k);
        for (int k = threadIdx.x; k < kCfgWords; k += NT)
            reinterpret_cast<unsigned long long*>(&s_cfg)[k] = reinterpret_cast<const unsigned long long*>(P.cfg)[k];
        __shared__ CtrlShared s_ctrl;
        __syncthreads();
        if (s_state.row_overflow && s_state.phase != PH_DONE) {  // block-uniform: a row of a wide association filled up
            __syncthreads();
            if (threadIdx.x == 0) {
                s_state.error = kErrRowOverflow;
                s_state.phase = PH_DONE;
            }
            __syncthreads();
        }
        if (s_state.phase != PH_DONE) {  // block-uniform
            // moment expansion around the pose the residuals were taken at: 36 entries of N, then 27 output tasks
            if (threadIdx.x == 0) quat_frame(evaluated_at(&s_state), &s_ctrl.frame);
            __syncthreads();
            if (threadIdx.x < kExpandN) s_ctrl.N[threadIdx.x] = expand_N_entry(s_ctrl.frame, threadIdx.x);
            __syncthreads();
            if (threadIdx.x < kExpandTasks) expand_task(s_sum, s_ctrl.N, threadIdx.x, &s_ctrl.ev);
            __syncthreads();
            if (threadIdx.x < 32 && !(use_cond & 4))  // & 4: timing probe without the LM state machine
                warp_controller(&s_state, &s_cfg, &s_ctrl, P.history, P.stats, P.max_hist, max_ticks, threadIdx.x);
        }
        __syncthreads();
        for (int k = threadIdx.x; k < kStateWords; k += NT)
            reinterpret_cast<unsigned long long*>(st)[k] = reinterpret_cast<const unsigned long long*>(&s_state)[k];
        __syncthreads();
    }
    } else if (blockIdx.x != 0) {
        return;  // a finished pair: one block keeps the tick protocol going
    }
    // ---- loop condition: the last pair to finish its tick publishes "is anything still running" ----------------
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&loop->pairs_done, 1) == n_pairs - 1) {
            __threadfence();
            loop->pairs_done = 0;
            int active = 0, searching = 0;
            for (int p = 0; p < n_pairs; ++p) {
                const int phase = *reinterpret_cast<volatile int*>(&pairs[p].state->phase);
                active |= (phase != PH_DONE);
                searching |= (phase == PH_SEARCH);
            }
            loop->active = active;
            if (use_cond & 1) {
                cudaGraphSetConditional(cond, active ? 1u : 0u);                 // WHILE: another tick
                cudaGraphSetConditional(cond_search, searching ? 1u : 0u);       // IF: the next tick starts with a search
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// align() entry and exit
// ------------------------------------------------------------------------------------------------------------

// the first hasConverged() test of align() (:65); also arms the loop condition for the first tick
__global__ void k_align_begin(const PairDev* __restrict__ pairs, int n_pairs, LoopCtl* __restrict__ loop)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_pairs) align_begin(pairs[p].state, pairs[p].cfg);
    if (p == 0) {
        loop->active = 1;
        loop->pairs_done = 0;
    }
}

// Epilogue of align(): the increment of the LAST outer iteration (every earlier one is applied by the search that
// follows it).  registration.cc:110-112.
__global__ void k_transform_final(const PairDev* __restrict__ pairs)
{
    const PairDev& P = pairs[blockIdx.y];
    PairState* st = P.state;
    if (!st->apply_dT) return;
    __shared__ double T[12];
    if (threadIdx.x < 12) T[threadIdx.x] = st->dT[threadIdx.x];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n_src; i += gridDim.x * blockDim.x) {
        float4 p = P.src[i];
        const double x = p.x, y = p.y, z = p.z;
        p.x = transform_row(T, x, y, z);
        p.y = transform_row(T + 4, x, y, z);
        p.z = transform_row(T + 8, x, y, z);
        P.src[i] = p;
    }
}

__global__ void k_transform_done(const PairDev* __restrict__ pairs, int n_pairs)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_pairs) pairs[p].state->apply_dT = 0;
}

__global__ void k_transform_plain(float4* __restrict__ pts, int n, const double* __restrict__ Tm)
{
    __shared__ double T[12];
    if (threadIdx.x < 12) T[threadIdx.x] = Tm[threadIdx.x];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 p = pts[i];
        const double x = p.x, y = p.y, z = p.z;
        p.x = transform_row(T, x, y, z);
        p.y = transform_row(T + 4, x, y, z);
        p.z = transform_row(T + 8, x, y, z);
        pts[i] = p;
    }
}

// ------------------------------------------------------------------------------------------------------------
// replay of the per-iteration diagnostics (registration.cc:110-122, utilities.hpp:16-26)
// ------------------------------------------------------------------------------------------------------------
//
// One outer iteration of the reference's bookkeeping on a full-resolution cloud: x <- float(dT x) in place, and in the same
// pass the two "MSE" figures -- really mean Euclidean distances, evaluated in float32 like calculateMSE -- of the moved
// cloud to the ground truth and to where it was before the move.  Per-block sums in double; k_replay_fold adds them in
// block order, so the figures do not depend on scheduling.
constexpr int kReplayThreads = 256;

__device__ __forceinline__ float replay_distance(const float4& a, const float4& b)
{
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z);
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
}

__global__ void __launch_bounds__(kReplayThreads) k_replay_step(float4* __restrict__ pts, const float4* __restrict__ gt, int n,
                                                                const double* __restrict__ Tm, double* __restrict__ partial)
{
    __shared__ double T[12];
    __shared__ double s_red[2][kReplayThreads / 32];
    if (threadIdx.x < 12) T[threadIdx.x] = Tm[threadIdx.x];
    __syncthreads();
    double sum_gt = 0.0, sum_prev = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 old = pts[i];
        float4 p = old;
        const double x = p.x, y = p.y, z = p.z;
        p.x = transform_row(T, x, y, z);
        p.y = transform_row(T + 4, x, y, z);
        p.z = transform_row(T + 8, x, y, z);
        pts[i] = p;
        sum_prev += static_cast<double>(replay_distance(p, old));
        if (gt) sum_gt += static_cast<double>(replay_distance(p, gt[i]));
    }
    for (int o = 16; o > 0; o >>= 1) {
        sum_gt += __shfl_xor_sync(kFull, sum_gt, o);
        sum_prev += __shfl_xor_sync(kFull, sum_prev, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_red[0][threadIdx.x >> 5] = sum_gt;
        s_red[1][threadIdx.x >> 5] = sum_prev;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        double v = 0.0;
        for (int w = 0; w < kReplayThreads / 32; ++w) v += s_red[threadIdx.x][w];
        partial[2 * blockIdx.x + threadIdx.x] = v;
    }
}

// out[0] = mean distance to the ground truth, out[1] = to the previous position
__global__ void k_replay_fold(const double* __restrict__ partial, int n_blocks, int n, double* __restrict__ out)
{
    if (threadIdx.x < 2) {
        double v = 0.0;
        for (int b = 0; b < n_blocks; ++b) v += partial[2 * b + threadIdx.x];
        out[threadIdx.x] = v / static_cast<double>(n);
    }
}

// ------------------------------------------------------------------------------------------------------------
// voxel filter (pcl::VoxelGrid default settings): key, sort by key (radix sort, host side), segmented mean
// ------------------------------------------------------------------------------------------------------------

struct VoxelGeom {
    float inv_leaf;
    int minb[3];
    int mul[3];
};

__global__ void k_voxel_keys(const float4* __restrict__ pts, int n, VoxelGeom vg, unsigned* __restrict__ keys,
                             unsigned* __restrict__ vals)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    // ijk = int(floor(p * inv_leaf) - float(min_b)), PCL's float arithmetic
    const int ix = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.x, vg.inv_leaf)), static_cast<float>(vg.minb[0])));
    const int iy = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.y, vg.inv_leaf)), static_cast<float>(vg.minb[1])));
    const int iz = __float2int_rz(__fsub_rn(floorf(__fmul_rn(p.z, vg.inv_leaf)), static_cast<float>(vg.minb[2])));
    keys[i] = static_cast<unsigned>(ix * vg.mul[0] + iy * vg.mul[1] + iz * vg.mul[2]);
    vals[i] = static_cast<unsigned>(i);
}

// head flag per sorted entry -> scanned into output slots; one thread per voxel walks its run in index order
__global__ void k_voxel_heads(const unsigned* __restrict__ keys, int n, int* __restrict__ head)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void k_voxel_mean(const float4* __restrict__ pts, const unsigned* __restrict__ keys,
                             const unsigned* __restrict__ vals, const int* __restrict__ slot, int n,
                             float4* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (!(i == 0 || keys[i] != keys[i - 1])) return;
    const unsigned key = keys[i];
    float cx = 0.f, cy = 0.f, cz = 0.f;
    int k = i;
    while (k < n && keys[k] == key) {  // float32 accumulation in point-index order (the sort is stable)
        const float4 p = pts[vals[k]];
        cx = __fadd_rn(cx, p.x);
        cy = __fadd_rn(cy, p.y);
        cz = __fadd_rn(cz, p.z);
        ++k;
    }
    const float cnt = static_cast<float>(k - i);
    out[slot[i]] = make_float4(__fdiv_rn(cx, cnt), __fdiv_rn(cy, cnt), __fdiv_rn(cz, cnt), 1.0f);
}

// ------------------------------------------------------------------------------------------------------------
// closest-point metrics (utilities.hpp:28-234): statistics of the squared 1-NN distances of one cloud in another
// ------------------------------------------------------------------------------------------------------------
//
// The seven helpers of the reference differ only in what they do with the vector of nearestKSearch(k = 1) distances: its
// sum, its "median" (the reference's own index rule, one position above the textbook one), and the sum / count / median
// of the entries inside a window [median / f, median * f].  The search is k_search with m = 1 and no radius; the vector is
// sorted once (radix sort), k_closest_plan finds the medians and the windows -- contiguous index ranges of the sorted
// vector -- by binary search, k_closest_reduce adds up the three ranges per block in double, k_closest_fold adds the block
// sums in block order (fixed order: run-to-run identical) and writes the nine results.

struct ClosestPlan {
    double med_d;      // "median" of the distances as a vector<double> (robustSumSquaredError family)
    double med_f;      // ... as a vector<float>: the two middle floats are added in float32 (medianClosestDistance family)
    double robust_med; // robustMedianClosestDistance
    int lo3, hi3;      // [lo3, hi3): entries with med_d / 3 <= v <= med_d * 3
    int lof, hif;      // the same with the caller's factor
    int pad[2];
};

__device__ __forceinline__ double closest_median(const float* __restrict__ v, int n, bool add_in_float)
{
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (n % 2 != 0) {
        const int k = (n + 1) / 2;
        return k < n ? static_cast<double>(v[k]) : nan;  // (out of bounds in the reference for n = 1)
    }
    const int a = n / 2, b = n / 2 + 1;
    if (b >= n) return nan;
    if (add_in_float) return static_cast<double>(__fadd_rn(v[a], v[b])) / 2.0;
    return (static_cast<double>(v[a]) + static_cast<double>(v[b])) / 2.0;
}

// first index in [0, n) whose entry is >= lo (as doubles) / first index whose entry is > hi; NaN bounds give empty ranges
__device__ __forceinline__ void closest_window(const float* __restrict__ v, int n, double lo, double hi, int* first, int* last)
{
    int a = 0, b = n;
    while (a < b) {
        const int mid = a + ((b - a) >> 1);
        if (static_cast<double>(v[mid]) >= lo) b = mid; else a = mid + 1;
    }
    *first = a;
    int c = 0, d = n;
    while (c < d) {
        const int mid = c + ((d - c) >> 1);
        if (static_cast<double>(v[mid]) <= hi) c = mid + 1; else d = mid;
    }
    *last = c;
    if (!(lo == lo) || !(hi == hi) || *last < *first) *first = *last = 0;
}

__global__ void k_closest_plan(const float* __restrict__ sorted, int n, double factor, ClosestPlan* __restrict__ plan)
{
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    ClosestPlan p{};
    p.med_d = closest_median(sorted, n, false);
    p.med_f = closest_median(sorted, n, true);
    closest_window(sorted, n, p.med_d / 3, p.med_d * 3, &p.lo3, &p.hi3);
    closest_window(sorted, n, p.med_d / factor, p.med_d * factor, &p.lof, &p.hif);
    int a = 0, b = 0;
    closest_window(sorted, n, p.med_f / 3.0, p.med_f * 3, &a, &b);
    const int nf = b - a;
    p.robust_med = nf > 0 ? closest_median(sorted + a, nf, true) / static_cast<double>(nf) : __longlong_as_double(0x7ff8000000000000ll);
    *plan = p;
}

constexpr int kClosestThreads = 256;

__global__ void __launch_bounds__(kClosestThreads) k_closest_reduce(const float* __restrict__ sorted, int n,
                                                                    const ClosestPlan* __restrict__ plan, double* __restrict__ partial)
{
    __shared__ double s_red[3][kClosestThreads / 32];
    const ClosestPlan p = *plan;
    double s[3] = {0.0, 0.0, 0.0};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double v = static_cast<double>(sorted[i]);
        s[0] += v;
        if (i >= p.lo3 && i < p.hi3) s[1] += v;
        if (i >= p.lof && i < p.hif) s[2] += v;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(kFull, s[k], o);
        if ((threadIdx.x & 31) == 0) s_red[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int w = 0; w < kClosestThreads / 32; ++w) v += s_red[threadIdx.x][w];
        partial[3 * blockIdx.x + threadIdx.x] = v;
    }
}

// out[9]: average, sum, robust sum (3), robust sum (factor), robust averaged sum, median, robust median, window counts
__global__ void k_closest_fold(const double* __restrict__ partial, int n_blocks, int n, const ClosestPlan* __restrict__ plan,
                               double* __restrict__ out)
{
    __shared__ double s_sum[3];
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int b = 0; b < n_blocks; ++b) v += partial[3 * b + threadIdx.x];
        s_sum[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const ClosestPlan p = *plan;
        const double big = 1.7976931348623157e308;  // std::numeric_limits<double>::max(), the helpers' "too few points"
        const int n3 = p.hi3 - p.lo3, nf = p.hif - p.lof;
        out[0] = s_sum[0] / static_cast<double>(n);
        out[1] = s_sum[0];
        out[2] = n3 < 10 ? big : s_sum[1];
        out[3] = nf < 10 ? big : s_sum[2];
        out[4] = n3 < 10 ? big : s_sum[1] / static_cast<double>(n3);
        out[5] = p.med_f;
        out[6] = p.robust_med;
        out[7] = static_cast<double>(n3);
        out[8] = static_cast<double>(nf);
    }
}

// L2 flush helper for benchmarks
__global__ void k_fill(float4* __restrict__ p, size_t n, float v)
{
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        p[i] = make_float4(v, v, v, v);
}

}  // namespace ppcr
#endif
