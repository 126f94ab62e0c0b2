// ppcr_kernels.cuh -- sm_100a kernels of the registration hot path.
//
//   tree build      k_bbox, k_tree_keys, (radix sort), k_tree_gather, k_tree_root, k_tree_split_level
//                                                                     (replaces the kd-tree build, registration.cc:66-67)
//   radius search   k_search<CAP>                                     (replaces the radiusSearch loop, :72-81, and the
//                                                                      CSR assembly, :69-83)
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
constexpr int kEvalThreads = 256;
#ifndef PPCR_EVAL_BATCH
#define PPCR_EVAL_BATCH 5
#endif
#ifndef PPCR_EVAL_MIN_BLOCKS
#define PPCR_EVAL_MIN_BLOCKS 4  // resident blocks per SM the register allocation of k_evalctl is held to
#endif
constexpr int kMailDoubles = 32;  // 24 moments + K + sequence stamp, padded
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
    float r2f;      // float(radius * radius): strict membership bound (FLANN)
    int* nbr_pos;   // [m][n_pad] slot-major association: positions in tgt_sorted
    const int* inv_perm;  // [n_tgt] original index -> position in tgt_sorted
    float* nbr_d2;  // optional (stage API only), may be null
    float* nbr_kth; // d2 of the m-th neighbour found by the last search (+inf when fewer were found): warm start
    int* nbr_cnt;
    double* partials;
    int n_eval_blocks;
    int max_hist;
    PairState* state;
    const Config* cfg;
    double* history;
    IterStats* stats;
    WeightCfg wcfg;
    // sharded mode: mailbox exchange of the moment vector between ranks
    double* mailbox;          // [world][kMailDoubles] on THIS device, written by the peers
    double* peer_mailbox[8];  // the same buffer on every rank (peer-mapped), indexed by rank
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

__device__ __forceinline__ void search_store(const PairDev& P, int i, int e, unsigned long long key)
{
    const size_t o = static_cast<size_t>(e) * P.n_pad + i;
    P.nbr_pos[o] = __ldg(P.inv_perm + key_index(key));
    if (P.nbr_d2) P.nbr_d2[o] = key_d2(key);
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
template <int VAR>
__global__ void __launch_bounds__(kSearchThreads) k_search(const PairDev* __restrict__ pairs)
{
    extern __shared__ unsigned long long s_heap[];
    const PairDev& P = pairs[blockIdx.y];
    PairState* st = P.state;
    if (st->phase != PH_SEARCH) return;
    __shared__ double s_T[12];
    __shared__ int s_chunk;
    const bool moving = st->apply_dT != 0;
    if (threadIdx.x < 12) s_T[threadIdx.x] = st->dT[threadIdx.x];
    const int m = P.m;
    const int n_chunks = (P.n_src + kSearchChunk - 1) / kSearchChunk;
    int stack[2 * kTreeStack];
    int cnt_total = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_chunk = atomicAdd(&st->search_cursor, 1);
        __syncthreads();
        const int chunk = s_chunk;
        if (chunk >= n_chunks) break;
        const int i = chunk * kSearchChunk + threadIdx.x;
        if (i >= P.n_src) continue;
        float4 q = P.src[i];
        float bound0 = P.r2f;
        if (moving) {
            const double x = q.x, y = q.y, z = q.z;
            const float nx = transform_row(s_T, x, y, z), ny = transform_row(s_T + 4, x, y, z),
                        nz = transform_row(s_T + 8, x, y, z);
            const float prev = P.nbr_kth[i];  // d2 of the m-th neighbour of the last search, +inf if it found fewer
            const float ddx = nx - q.x, ddy = ny - q.y, ddz = nz - q.z;
            const float move = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz) * 1.000001f;
            const float reach = sqrtf(prev) * 1.000001f + move;
            bound0 = fminf(P.r2f, reach * reach * 1.00001f);  // inf stays inf -> r2f
            q.x = nx;
            q.y = ny;
            q.z = nz;
            P.src[i] = q;
        }
        int cnt = 0;
        float kth = __int_as_float(0x7f800000);
        HeapList<kSearchThreads, (VAR & 1) != 0> L;
        L.k = s_heap + threadIdx.x;
        L.init(m);
        tree_search(P.tree, P.nodes, P.tgt_sorted, q.x, q.y, q.z, P.r2f, bound0, L, stack);
        for (int s = 0; s < L.n; ++s) {
            const unsigned long long key = L.k[s * kSearchThreads];
            if ((VAR & 1) || key != kKeyInf) search_store(P, i, cnt++, key);
        }
        if (L.n == m && L.k[0] != kKeyInf) kth = key_d2(L.k[0]);
        P.nbr_cnt[i] = cnt;
        P.nbr_kth[i] = kth;
        cnt_total += cnt;
    }
    // association size: warp sum, one atomic per warp
    for (int o = 16; o > 0; o >>= 1) cnt_total += __shfl_xor_sync(kFull, cnt_total, o);
    if ((threadIdx.x & 31) == 0 && cnt_total)
        atomicAdd(reinterpret_cast<unsigned long long*>(&st->K), static_cast<unsigned long long>(cnt_total));
}

// ------------------------------------------------------------------------------------------------------------
// weights + normal-equation moments
// ------------------------------------------------------------------------------------------------------------

constexpr int kCtrlGroups = kEvalThreads / 32;  // partial sums are folded in kCtrlGroups interleaved chains

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

// (Hs + diag(D)^2) y = gs, rows of the factor spread over lanes 0..6; returns validity in every lane
__device__ __forceinline__ int warp_solve_damped(const PairState* st, CtrlShared* sh, int lane)
{
    if (lane < kNP)
        for (int c = 0; c < kNP; ++c) sh->L[lane][c] = st->Hs[lane * kNP + c] + (lane == c ? sh->D[lane] * sh->D[lane] : 0.0);
    if (lane == 0) sh->flag = 1;
    __syncwarp();
    for (int c = 0; c < kNP; ++c) {
        if (lane == c) {
            double d = sh->L[c][c];
            for (int k = 0; k < c; ++k) d -= sh->L[c][k] * sh->L[c][k];
            if (!(d > 0.0)) sh->flag = 0;
            else sh->L[c][c] = sqrt(d);
        }
        __syncwarp();
        if (!sh->flag) break;
        if (lane > c && lane < kNP) {
            double s = sh->L[lane][c];
            for (int k = 0; k < c; ++k) s -= sh->L[lane][k] * sh->L[c][k];
            sh->L[lane][c] = s / sh->L[c][c];
        }
        __syncwarp();
    }
    if (lane == 0 && sh->flag) {
        double z[kNP];
        for (int r = 0; r < kNP; ++r) {
            double s = st->gs[r];
            for (int k = 0; k < r; ++k) s -= sh->L[r][k] * z[k];
            z[r] = s / sh->L[r][r];
        }
        for (int r = kNP - 1; r >= 0; --r) {
            double s = z[r];
            for (int k = r + 1; k < kNP; ++k) s -= sh->L[k][r] * sh->y[k];
            sh->y[r] = s / sh->L[r][r];
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

// The float32 row loop of k_evalctl: one source row per thread, grid-stride.  The neighbour records of a row are
// kU slots apart in the slot-major planes (each a coalesced 512-byte read per warp); full batches of kU records are
// loaded without predicates before their arithmetic, the ragged tail of a row with.
template <int WM, bool SAME>
__device__ __forceinline__ void eval_rows_fast(const PairDev& P, const Pose& pe, const Pose& pw, const WeightCfg& wc, double* acc)
{
    constexpr int kU = PPCR_EVAL_BATCH;
    const int stride = P.n_eval_blocks * kEvalThreads;
    const size_t n_pad = P.n_pad;
    for (int i = blockIdx.x * kEvalThreads + threadIdx.x; i < P.n_src; i += stride) {
        const int cnt = P.nbr_cnt[i];
        if (cnt == 0) continue;
        const float4 sp = P.src[i];
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
        const int* rec = P.nbr_pos + i;
        const float4* __restrict__ table = P.tgt_sorted;
        int k0 = 0;
        for (; k0 + kU <= cnt; k0 += kU) {
            int pos[kU];
            float4 y[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) pos[u] = __ldg(rec + static_cast<size_t>(u) * n_pad);
            rec += static_cast<size_t>(kU) * n_pad;
#pragma unroll
            for (int u = 0; u < kU; ++u) y[u] = __ldg(table + pos[u]);
#pragma unroll
            for (int u = 0; u < kU; ++u) rowf_add_t<WM, SAME>(&row, wc, y[u].x, y[u].y, y[u].z, he, dw);
        }
        if (k0 < cnt) {
            int pos[kU - 1];
            float4 y[kU - 1];
#pragma unroll
            for (int u = 0; u < kU - 1; ++u)
                if (k0 + u < cnt) pos[u] = __ldg(rec + static_cast<size_t>(u) * n_pad);
#pragma unroll
            for (int u = 0; u < kU - 1; ++u)
                if (k0 + u < cnt) y[u] = __ldg(table + pos[u]);
#pragma unroll
            for (int u = 0; u < kU - 1; ++u)
                if (k0 + u < cnt) rowf_add_t<WM, SAME>(&row, wc, y[u].x, y[u].y, y[u].z, he, dw);
        }
        rowf_end_s<kEvalThreads>(&row, sx, sy, sz, acc);
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
__global__ void __launch_bounds__(kEvalThreads, PPCR_EVAL_MIN_BLOCKS) k_evalctl(const PairDev* __restrict__ pairs, int n_pairs,
                                                          LoopCtl* __restrict__ loop, cudaGraphConditionalHandle cond,
                                                          int use_cond, int max_ticks)
{
    const PairDev& P = pairs[blockIdx.y];
    PairState* st = P.state;
    __shared__ Pose s_pe, s_pw;
    __shared__ double s_red[kEvalThreads / 32][kNSum];
    __shared__ double s_sum[kMailDoubles];
    __shared__ int s_flag;
    const bool live = st->phase != PH_DONE;
    if (live) {
    if (static_cast<int>(blockIdx.x) >= P.n_eval_blocks) return;
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
    // the 24 float64 accumulators of a thread live in shared memory (one column per thread, conflict-free): they are
    // touched once per row, and keeping them out of the register file leaves room to have a whole row in flight
    extern __shared__ double s_acc[];
    double* acc = s_acc + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kNSum; ++k) acc[k * kEvalThreads] = 0.0;
    const int stride = P.n_eval_blocks * kEvalThreads;
    const size_t n_pad = P.n_pad;
    if constexpr (kFast) {
        // pose_w == pose_e on the first evaluation of every outer iteration: one residual serves both uses
        bool same = true;
#pragma unroll
        for (int k = 0; k < 12; ++k)
            same = same && (reinterpret_cast<const double*>(&s_pe)[k] == reinterpret_cast<const double*>(&s_pw)[k]);
        // one instantiation of the row loop per (weight model, same pose): both are uniform over the launch
        switch (weight_mode(wc) * 2 + (same ? 1 : 0)) {
            case WM_T_H4 * 2: eval_rows_fast<WM_T_H4, false>(P, pe, pw, wc, acc); break;
            case WM_T_H4 * 2 + 1: eval_rows_fast<WM_T_H4, true>(P, pe, pw, wc, acc); break;
            case WM_T_INT * 2: eval_rows_fast<WM_T_INT, false>(P, pe, pw, wc, acc); break;
            case WM_T_INT * 2 + 1: eval_rows_fast<WM_T_INT, true>(P, pe, pw, wc, acc); break;
            case WM_T_REAL * 2: eval_rows_fast<WM_T_REAL, false>(P, pe, pw, wc, acc); break;
            case WM_T_REAL * 2 + 1: eval_rows_fast<WM_T_REAL, true>(P, pe, pw, wc, acc); break;
            case WM_GAUSS * 2: eval_rows_fast<WM_GAUSS, false>(P, pe, pw, wc, acc); break;
            default: eval_rows_fast<WM_GAUSS, true>(P, pe, pw, wc, acc); break;
        }
    } else {
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
        }
    }
    // fixed-shape reduction: xor-shuffle tree inside the warp, then warps in index order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kNSum; ++k) {
        double v = acc[k * kEvalThreads];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        if (lane == 0) s_red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kNSum) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < kEvalThreads / 32; ++w) v += s_red[w][threadIdx.x];
        P.partials[static_cast<size_t>(blockIdx.x) * kNSum + threadIdx.x] = v;
    }
    // ---- last block standing runs the controller --------------------------------------------------------------
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_flag = (atomicAdd(&st->eval_ticket, 1) == P.n_eval_blocks - 1);
    __syncthreads();
    if (!s_flag) return;
    __threadfence();
    if (threadIdx.x == 0) st->eval_ticket = 0;
    if (use_cond & 2) return;  // timing probe (ppcr_time_kernel): the streaming part alone
    {
        // group g folds blocks g, g + G, g + 2G, ... in order (loads issued eight at a time), then the groups in order
        double v = 0.0;
        if (lane < kNSum) {
            const double* base = P.partials + lane;
            int b = warp;
            for (; b + 7 * kCtrlGroups < P.n_eval_blocks; b += 8 * kCtrlGroups) {
                double t[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) t[u] = __ldcg(base + static_cast<size_t>(b + u * kCtrlGroups) * kNSum);
#pragma unroll
                for (int u = 0; u < 8; ++u) v += t[u];
            }
            for (; b < P.n_eval_blocks; b += kCtrlGroups) v += __ldcg(base + static_cast<size_t>(b) * kNSum);
            s_red[warp][lane] = v;
        }
        __syncthreads();
        if (threadIdx.x < kNSum) {
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < kCtrlGroups; ++w) t += s_red[w][threadIdx.x];
            s_sum[threadIdx.x] = t;
        }
        __syncthreads();
    }

    if (P.world > 1) {
        // Sharded pair: every rank adds the other ranks' moments (and association sizes) in rank order, so all
        // ranks hold bit-identical sums and take identical decisions.  The exchange is a one-shot all-gather
        // written straight into the peers' mailboxes over NVLink; a sequence stamp doubles as the ready flag.
        const int seq = st->ticks + 1;
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
                *reinterpret_cast<volatile double*>(dst + kMailDoubles - 1) = static_cast<double>(seq);
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
            while (ld_volatile_f64(src + kMailDoubles - 1) != static_cast<double>(seq)) {
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
        __syncthreads();
    }

    {
        __shared__ PairState s_state;
        __shared__ Config s_cfg;
        constexpr int kStateWords = sizeof(PairState) / 8, kCfgWords = sizeof(Config) / 8;
        for (int k = threadIdx.x; k < kStateWords; k += kEvalThreads)
            reinterpret_cast<unsigned long long*>(&s_state)[k] = __ldcg(reinterpret_cast<const unsigned long long*>(st) + k);
        for (int k = threadIdx.x; k < kCfgWords; k += kEvalThreads)
            reinterpret_cast<unsigned long long*>(&s_cfg)[k] = reinterpret_cast<const unsigned long long*>(P.cfg)[k];
        __shared__ CtrlShared s_ctrl;
        __syncthreads();
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
        for (int k = threadIdx.x; k < kStateWords; k += kEvalThreads)
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
            int active = 0;
            for (int p = 0; p < n_pairs; ++p) active |= (*reinterpret_cast<volatile int*>(&pairs[p].state->phase) != PH_DONE);
            loop->active = active;
            if (use_cond & 1) cudaGraphSetConditional(cond, active ? 1u : 0u);
        }
    }
}

// weights of the current association at pose_w, written slot-major (parity dumps only)
template <bool kFast>
__global__ void k_dump_weights(const PairDev* __restrict__ pairs, double* __restrict__ out)
{
    const PairDev& P = pairs[0];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_src) return;
    const int cnt = P.nbr_cnt[i];
    if (cnt == 0) return;
    const Pose pw = P.state->pose_w;
    const float4 sp = P.src[i];
    double ptw[3];
    apply_pose(pw, sp.x, sp.y, sp.z, ptw);
    const size_t n_pad = P.n_pad;
    if constexpr (kFast) {
        PointHL hw;
        split_point(ptw, &hw);
        RowAccF row;
        rowf_begin(&row);
        for (int k = 0; k < cnt; ++k) {
            const size_t o = static_cast<size_t>(k) * n_pad + i;
            const float zero[3] = {0.f, 0.f, 0.f};
            const float4 y = P.tgt_sorted[P.nbr_pos[o]];
            rowf_add(&row, P.wcfg, y.x, y.y, y.z, hw, zero, true);
        }
        for (int k = 0; k < cnt; ++k) {
            const size_t o = static_cast<size_t>(k) * n_pad + i;
            const float4 y = P.tgt_sorted[P.nbr_pos[o]];
            out[o] = rowf_finished_weight(&row, P.wcfg, y.x, y.y, y.z, hw);
        }
    } else {
        RowAcc row;
        row_begin(&row);
        for (int k = 0; k < cnt; ++k) {
            const size_t o = static_cast<size_t>(k) * n_pad + i;
            const float4 y = P.tgt_sorted[P.nbr_pos[o]];
            row_add<false>(&row, P.wcfg, y.x, y.y, y.z, ptw, ptw);
        }
        for (int k = 0; k < cnt; ++k) {
            const size_t o = static_cast<size_t>(k) * n_pad + i;
            const float4 y = P.tgt_sorted[P.nbr_pos[o]];
            out[o] = finished_weight<false>(&row, P.wcfg, y.x, y.y, y.z, ptw);
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

// L2 flush helper for benchmarks
__global__ void k_fill(float4* __restrict__ p, size_t n, float v)
{
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        p[i] = make_float4(v, v, v, v);
}

}  // namespace ppcr
#endif
