// ppcr_kernels.cuh -- sm_100a kernels of the registration hot path.
//
//   grid build      k_bbox, k_cell_count, k_scan_*, k_cell_scatter   (replaces the kd-tree build, registration.cc:66-67)
//   radius search   k_search<R>                                       (replaces the radiusSearch loop, :72-81, and the
//                                                                      CSR assembly, :69-83)
//   weights + J^TWJ k_eval<FAST>                                      (WeightUpdaterCallback, ProbabilisticWeights,
//                                                                      ErrorTerm + Ceres' Jacobian evaluation)
//   LM controller   k_controller                                      (ceres::Solve's trust-region loop, pose
//                                                                      composition, cost drop, hasConverged)
//   cloud move      k_transform                                       (pcl::transformPointCloud, :110-112)
//   voxel filter    k_voxel_* (+ a radix sort of the voxel keys)      (pcl::VoxelGrid, :24-41)
//
// Data layout in HBM (per pair):
//   tgt_sorted  float4[n_tgt]      target points counting-sorted by grid cell, .w = original index (int bits)
//   cell_start  int[n_cells + 1]   CSR over cells, x fastest, so a run of x-adjacent cells is one contiguous range
//   src         float4[n_src]      the moving source cloud (filtered), .w = original index
//   nbr_x/y/z   float[m][n_pad]    slot-major neighbour coordinates: entry (k, i) is the k-th nearest target of
//   nbr_idx     int[m][n_pad]      source i, so one warp reads 32 consecutive floats per slot (fully coalesced)
//   nbr_cnt     int[n_pad]
//   partials    double[blocks][24] per-block moment sums, reduced in a fixed order by the controller
// No tensor cores: nothing on this path is a dense contraction; the kernels are HBM/L2-bound streaming passes.
#ifndef PPCR_KERNELS_CUH
#define PPCR_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "ppcr_eval.h"
#include "ppcr_lm.h"

namespace ppcr {

constexpr int kTileQ = 32;        // queries per search block (one output tile)
constexpr int kSearchWarps = 8;   // warps per search block
constexpr int kEvalThreads = 256;
constexpr unsigned kFull = 0xffffffffu;

struct GridDev {
    float ox, oy, oz;   // origin = min corner of the target bounding box
    float inv_h;        // 1 / cell edge
    float h_cover;      // cell edge, rounded down: used for the "covered radius" bound of the shell scan
    float cover_slack;  // absolute slack for float cell-boundary fuzz
    int nx, ny, nz;
    int n_cells;
};

struct PairDev {
    const float4* tgt_sorted;
    const int* cell_start;
    GridDev grid;
    int n_tgt;
    float4* src;
    int n_src;
    int n_pad;
    int m;          // result capacity = min(max_neighbours, n_tgt)
    float r2f;      // float(radius * radius): strict membership bound (FLANN)
    float rpad;     // radius padded upwards, for the conservative cell window
    float* nbr_x;
    float* nbr_y;
    float* nbr_z;
    int* nbr_idx;   // original target index
    float* nbr_d2;  // optional (stage API only), may be null
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

constexpr int kMailDoubles = 32;  // 24 moments + K + sequence stamp, padded

// ------------------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------------------

__device__ __forceinline__ int cell_coord(float v, float origin, float inv_h)
{
    // the same expression bins targets and locates queries; monotone in v, saturating conversion
    return __float2int_rd(__fmul_rn(__fsub_rn(v, origin), inv_h));
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ float dist2_exact(float qx, float qy, float qz, float px, float py, float pz)
{
    // FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz in float32 with NO fused multiply-add
    const float dx = __fsub_rn(qx, px), dy = __fsub_rn(qy, py), dz = __fsub_rn(qz, pz);
    float acc = __fmul_rn(dx, dx);
    acc = __fadd_rn(acc, __fmul_rn(dy, dy));
    acc = __fadd_rn(acc, __fmul_rn(dz, dz));
    return acc;
}

// ------------------------------------------------------------------------------------------------------------
// grid build
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

__device__ __forceinline__ int cell_of_point(const GridDev& g, float x, float y, float z)
{
    const int cx = clampi(cell_coord(x, g.ox, g.inv_h), 0, g.nx - 1);
    const int cy = clampi(cell_coord(y, g.oy, g.inv_h), 0, g.ny - 1);
    const int cz = clampi(cell_coord(z, g.oz, g.inv_h), 0, g.nz - 1);
    return (cz * g.ny + cy) * g.nx + cx;
}

// pass 1 of the counting sort: cell of every point and its arrival rank inside the cell
__global__ void k_cell_count(const float4* __restrict__ pts, int n, GridDev g, int* __restrict__ counts,
                             int* __restrict__ cell_of, int* __restrict__ rank)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 p = pts[i];
    const int c = cell_of_point(g, p.x, p.y, p.z);
    cell_of[i] = c;
    rank[i] = atomicAdd(counts + c, 1);
}

__global__ void k_count_occupied(const int* __restrict__ counts, int n_cells, unsigned long long* __restrict__ out)
{
    int local = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_cells; i += gridDim.x * blockDim.x) local += counts[i] > 0;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(kFull, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, static_cast<unsigned long long>(local));
}

// exclusive scan, three passes; each block owns kScanTile consecutive items
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

// pass 3 of the counting sort: scatter into cell order; .w carries the original index
__global__ void k_cell_scatter(const float4* __restrict__ pts, int n, const int* __restrict__ cell_start,
                               const int* __restrict__ cell_of, const int* __restrict__ rank,
                               float4* __restrict__ sorted)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pts[i];
    p.w = __int_as_float(i);
    sorted[cell_start[cell_of[i]] + rank[i]] = p;
}

// tag every point with its own index in .w (source cloud kept in caller order)
__global__ void k_tag_index(float4* __restrict__ pts, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pts[i].w = __int_as_float(i);
}

// ------------------------------------------------------------------------------------------------------------
// radius search: one warp per query, Chebyshev shells of grid cells, warp-resident sorted result list
// ------------------------------------------------------------------------------------------------------------

constexpr unsigned long long kKeyInf = 0xffffffffffffffffull;

template <int R>
struct WarpList {  // entry e = 32*r + lane holds the e-th smallest (d2, index) key found so far
    unsigned long long key[R];
    int spos[R];  // position of that target in tgt_sorted
};

template <int R>
__device__ __forceinline__ void list_insert(WarpList<R>& L, unsigned long long k, int sp, int lane, int m)
{
    int pos = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) pos += __popc(__ballot_sync(kFull, L.key[r] < k));
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {
        const int base = 32 * r;
        if (pos >= base + 32) continue;  // warp-uniform
        unsigned long long up_k = __shfl_up_sync(kFull, L.key[r], 1);
        int up_s = __shfl_up_sync(kFull, L.spos[r], 1);
        if (r > 0) {
            const unsigned long long ck = __shfl_sync(kFull, L.key[r > 0 ? r - 1 : 0], 31);
            const int cs = __shfl_sync(kFull, L.spos[r > 0 ? r - 1 : 0], 31);
            if (lane == 0) {
                up_k = ck;
                up_s = cs;
            }
        }
        const int e = base + lane;
        if (e > pos) {
            L.key[r] = up_k;
            L.spos[r] = up_s;
        } else if (e == pos) {
            L.key[r] = k;
            L.spos[r] = sp;
        }
        if (e >= m) L.key[r] = kKeyInf;  // the element pushed past the capacity falls off
    }
}

template <int R>
__device__ __forceinline__ unsigned long long list_worst(const WarpList<R>& L, int m)
{
    // key of entry m-1: kKeyInf until the list is full, the current worst kept candidate afterwards
    const int e = m - 1;
    unsigned long long worst = kKeyInf;
#pragma unroll
    for (int r = 0; r < R; ++r)
        if ((e >> 5) == r) worst = __shfl_sync(kFull, L.key[r], e & 31);
    return worst;
}

// Offers one candidate per lane.  tau is the strict bound a candidate must beat: the squared radius until the
// list is full, the current worst afterwards (FLANN's KNNRadiusResultSet rule).
template <int R>
__device__ __forceinline__ void offer_chunk(WarpList<R>& L, unsigned long long ckey, int csp, int lane, int m,
                                            unsigned long long r2key, unsigned long long& worst,
                                            unsigned long long& tau)
{
    unsigned mask = __ballot_sync(kFull, ckey < tau);
    while (mask) {
        const int b = __ffs(mask) - 1;
        const unsigned long long k = __shfl_sync(kFull, ckey, b);
        const int sp = __shfl_sync(kFull, csp, b);
        list_insert<R>(L, k, sp, lane, m);
        worst = list_worst<R>(L, m);
        tau = worst < r2key ? worst : r2key;
        mask &= mask - 1;
        mask &= __ballot_sync(kFull, ckey < tau);
    }
}

// Scans the target grid around query q and leaves the (at most m) nearest in-radius targets in L, sorted.
template <int R>
__device__ void warp_search(const PairDev& P, float qx, float qy, float qz, int lane, WarpList<R>& L)
{
    const GridDev& g = P.grid;
    const int m = P.m;
    const unsigned long long r2key = static_cast<unsigned long long>(__float_as_uint(P.r2f)) << 32;
    unsigned long long tau = r2key, worst = kKeyInf;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        L.key[r] = kKeyInf;
        L.spos[r] = 0;
    }
    // conservative cell window of the radius ball (directed rounding keeps it a superset)
    int lo[3], hi[3], qc[3];
    lo[0] = cell_coord(__fsub_rd(qx, P.rpad), g.ox, g.inv_h);
    hi[0] = cell_coord(__fadd_ru(qx, P.rpad), g.ox, g.inv_h);
    lo[1] = cell_coord(__fsub_rd(qy, P.rpad), g.oy, g.inv_h);
    hi[1] = cell_coord(__fadd_ru(qy, P.rpad), g.oy, g.inv_h);
    lo[2] = cell_coord(__fsub_rd(qz, P.rpad), g.oz, g.inv_h);
    hi[2] = cell_coord(__fadd_ru(qz, P.rpad), g.oz, g.inv_h);
    qc[0] = cell_coord(qx, g.ox, g.inv_h);
    qc[1] = cell_coord(qy, g.oy, g.inv_h);
    qc[2] = cell_coord(qz, g.oz, g.inv_h);
    const int dims[3] = {g.nx, g.ny, g.nz};
    int s_min = 0, s_max = 0;
    bool empty = false;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = max(lo[a], 0);
        hi[a] = min(hi[a], dims[a] - 1);
        empty |= lo[a] > hi[a];
        // keep the shell arithmetic in range for queries far outside the grid
        qc[a] = clampi(qc[a], lo[a] - (1 << 20), hi[a] + (1 << 20));
        s_min = max(s_min, max(lo[a] - qc[a], qc[a] - hi[a]));
        s_max = max(s_max, max(qc[a] - lo[a], hi[a] - qc[a]));
    }
    if (empty) return;
    s_min = max(s_min, 0);

    for (int s = s_min; s <= s_max; ++s) {
        const int side = 2 * s + 1;
        const int n_rows = side * side;
        for (int row0 = 0; row0 < n_rows; row0 += 32) {
            // lane -> one (dy, dz) row of the shell; up to two x-runs per row
            const int r = row0 + lane;
            int begA = 0, lenA = 0, begB = 0, lenB = 0;
            if (r < n_rows) {
                const int dz = r / side - s, dy = r % side - s;
                const int cy = qc[1] + dy, cz = qc[2] + dz;
                if (cy >= lo[1] && cy <= hi[1] && cz >= lo[2] && cz <= hi[2]) {
                    const int row_base = (cz * g.ny + cy) * g.nx;
                    const bool face = (abs(dy) == s) || (abs(dz) == s);
                    if (face) {
                        const int x0 = max(lo[0], qc[0] - s), x1 = min(hi[0], qc[0] + s);
                        if (x0 <= x1) {
                            begA = __ldg(P.cell_start + row_base + x0);
                            lenA = __ldg(P.cell_start + row_base + x1 + 1) - begA;
                        }
                    } else {
                        const int xa = qc[0] - s, xb = qc[0] + s;
                        if (xa >= lo[0] && xa <= hi[0]) {
                            begA = __ldg(P.cell_start + row_base + xa);
                            lenA = __ldg(P.cell_start + row_base + xa + 1) - begA;
                        }
                        if (xb >= lo[0] && xb <= hi[0]) {
                            begB = __ldg(P.cell_start + row_base + xb);
                            lenB = __ldg(P.cell_start + row_base + xb + 1) - begB;
                        }
                    }
                }
            }
            // flatten the runs of the 32 rows into one candidate stream
            const int len = lenA + lenB;
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += t;
            }
            const int total = __shfl_sync(kFull, incl, 31);
            for (int it = 0; it < total; it += 32) {
                const int gidx = it + lane;
                // owner = first lane whose inclusive prefix exceeds gidx
                int owner = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int probe = __shfl_sync(kFull, incl, owner + step - 1);
                    if (probe <= gidx) owner += step;
                }
                owner = min(owner, 31);
                const int o_incl = __shfl_sync(kFull, incl, owner);
                const int o_len = __shfl_sync(kFull, len, owner);
                const int o_begA = __shfl_sync(kFull, begA, owner);
                const int o_lenA = __shfl_sync(kFull, lenA, owner);
                const int o_begB = __shfl_sync(kFull, begB, owner);
                unsigned long long ckey = kKeyInf;
                int csp = 0;
                if (gidx < total) {
                    const int within = gidx - (o_incl - o_len);
                    csp = within < o_lenA ? o_begA + within : o_begB + (within - o_lenA);
                    const float4 p = __ldg(P.tgt_sorted + csp);
                    const float d2 = dist2_exact(qx, qy, qz, p.x, p.y, p.z);
                    ckey = (static_cast<unsigned long long>(__float_as_uint(d2)) << 32) |
                           static_cast<unsigned>(__float_as_int(p.w));
                }
                offer_chunk<R>(L, ckey, csp, lane, m, r2key, worst, tau);
            }
        }
        // every target not scanned yet lies farther than s*h (minus float slack) along some axis: once the list is
        // full and its worst entry is inside that covered radius, no later shell can change the result
        if (worst != kKeyInf) {
            const double cover = static_cast<double>(s) * g.h_cover - g.cover_slack;
            const double worst_d2 = __uint_as_float(static_cast<unsigned>(worst >> 32));
            if (cover > 0.0 && worst_d2 < cover * cover) break;
        }
    }
}

// block = 8 warps, one tile of 32 consecutive queries; results leave through shared memory so that the
// slot-major planes are written 128 bytes at a time
template <int R>
__global__ void __launch_bounds__(kSearchWarps * 32) k_search(const PairDev* __restrict__ pairs)
{
    extern __shared__ unsigned char smem_raw[];
    const PairDev& P = pairs[blockIdx.y];
    PairState* st = P.state;
    // the increment published by the previous tick has been consumed by k_transform: retire the flag
    if (blockIdx.x == 0 && threadIdx.x == 0) st->apply_dT = 0;
    if (st->phase != PH_SEARCH) return;
    const int tile0 = blockIdx.x * kTileQ;
    if (tile0 >= P.n_src) return;
    const int m = P.m;
    float* t_x = reinterpret_cast<float*>(smem_raw);
    float* t_y = t_x + m * kTileQ;
    float* t_z = t_y + m * kTileQ;
    int* t_i = reinterpret_cast<int*>(t_z + m * kTileQ);
    float* t_d = reinterpret_cast<float*>(t_i + m * kTileQ);
    __shared__ int t_cnt[kTileQ];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int qi = warp; qi < kTileQ; qi += kSearchWarps) {
        const int i = tile0 + qi;
        int cnt = 0;
        if (i < P.n_src) {
            const float4 q = P.src[i];
            WarpList<R> L;
            warp_search<R>(P, q.x, q.y, q.z, lane, L);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const bool have = L.key[r] != kKeyInf;
                cnt += __popc(__ballot_sync(kFull, have));
                const int e = 32 * r + lane;
                if (have) {
                    const float4 p = __ldg(P.tgt_sorted + L.spos[r]);
                    t_x[e * kTileQ + qi] = p.x;
                    t_y[e * kTileQ + qi] = p.y;
                    t_z[e * kTileQ + qi] = p.z;
                    t_i[e * kTileQ + qi] = __float_as_int(p.w);
                    t_d[e * kTileQ + qi] = __uint_as_float(static_cast<unsigned>(L.key[r] >> 32));
                }
            }
        }
        if (lane == 0) t_cnt[qi] = cnt;
    }
    __syncthreads();
    // coalesced write-out: one warp per slot row
    const int my_cnt = t_cnt[lane];
    const int i_out = tile0 + lane;
    for (int e = warp; e < m; e += kSearchWarps) {
        if (i_out < P.n_src && e < my_cnt) {
            const size_t o = static_cast<size_t>(e) * P.n_pad + i_out;
            P.nbr_x[o] = t_x[e * kTileQ + lane];
            P.nbr_y[o] = t_y[e * kTileQ + lane];
            P.nbr_z[o] = t_z[e * kTileQ + lane];
            P.nbr_idx[o] = t_i[e * kTileQ + lane];
            if (P.nbr_d2) P.nbr_d2[o] = t_d[e * kTileQ + lane];
        }
    }
    if (warp == 0) {
        if (i_out < P.n_src) P.nbr_cnt[i_out] = my_cnt;
        int total = (i_out < P.n_src) ? my_cnt : 0;
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(kFull, total, o);
        if (lane == 0 && total) atomicAdd(reinterpret_cast<unsigned long long*>(&st->K), static_cast<unsigned long long>(total));
    }
}

// ------------------------------------------------------------------------------------------------------------
// weights + normal-equation moments
// ------------------------------------------------------------------------------------------------------------

template <bool kFast>
__global__ void __launch_bounds__(kEvalThreads) k_eval(const PairDev* __restrict__ pairs)
{
    const PairDev& P = pairs[blockIdx.y];
    const PairState* st = P.state;
    if (st->phase == PH_DONE) return;
    if (static_cast<int>(blockIdx.x) >= P.n_eval_blocks) return;
    __shared__ Pose s_pe, s_pw;
    __shared__ double s_red[kEvalThreads / 32][kNSum];
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
    double acc[kNSum];
#pragma unroll
    for (int k = 0; k < kNSum; ++k) acc[k] = 0.0;
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
            const float yx = __ldg(P.nbr_x + o), yy = __ldg(P.nbr_y + o), yz = __ldg(P.nbr_z + o);
            row_add<kFast>(&row, wc, yx, yy, yz, pte, ptw);
        }
        row_end(&row, sx, sy, sz, acc);
    }
    // fixed-shape reduction: xor-shuffle tree inside the warp, then warps in index order
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kNSum; ++k) {
        double v = acc[k];
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
    RowAcc row;
    row_begin(&row);
    const size_t n_pad = P.n_pad;
    for (int k = 0; k < cnt; ++k) {
        const size_t o = static_cast<size_t>(k) * n_pad + i;
        row_add<kFast>(&row, P.wcfg, P.nbr_x[o], P.nbr_y[o], P.nbr_z[o], ptw, ptw);
    }
    for (int k = 0; k < cnt; ++k) {
        const size_t o = static_cast<size_t>(k) * n_pad + i;
        out[o] = finished_weight<kFast>(&row, P.wcfg, P.nbr_x[o], P.nbr_y[o], P.nbr_z[o], ptw);
    }
}

// ------------------------------------------------------------------------------------------------------------
// controller: reduce the per-block moments in a fixed order, then run the LM / outer-loop state machine
// ------------------------------------------------------------------------------------------------------------

constexpr int kCtrlThreads = 256;

__device__ __forceinline__ double ld_volatile_f64(const double* p)
{
    return *reinterpret_cast<const volatile double*>(p);
}

__global__ void __launch_bounds__(kCtrlThreads) k_controller(const PairDev* __restrict__ pairs, int max_ticks)
{
    const PairDev& P = pairs[blockIdx.x];
    PairState* st = P.state;
    if (st->phase == PH_DONE) return;
    __shared__ double s_part[kCtrlThreads / 32][kNSum];
    __shared__ double s_sum[kMailDoubles];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = 0.0;
    if (lane < kNSum)
        for (int b = warp; b < P.n_eval_blocks; b += kCtrlThreads / 32) v += P.partials[static_cast<size_t>(b) * kNSum + lane];
    if (lane < kNSum) s_part[warp][lane] = v;
    __syncthreads();
    if (threadIdx.x < kNSum) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < kCtrlThreads / 32; ++w) t += s_part[w][threadIdx.x];
        s_sum[threadIdx.x] = t;
    }
    __syncthreads();

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
            return;
        }
        if (threadIdx.x <= kNSum) {
            const int parity = seq & 1;
            double t = 0.0;
            for (int r = 0; r < P.world; ++r)
                t += ld_volatile_f64(P.mailbox + (static_cast<size_t>(parity) * P.world + r) * kMailDoubles + threadIdx.x);
            s_sum[threadIdx.x] = t;
        }
        __syncthreads();
        if (threadIdx.x == 0 && st->phase == PH_SEARCH) st->K = static_cast<int64_t>(s_sum[kNSum]);
        __syncthreads();
    }

    if (threadIdx.x == 0) {
        double S[kNSum];
        for (int k = 0; k < kNSum; ++k) S[k] = s_sum[k];
        st->evals += 1;
        controller_tick(st, P.cfg, S, P.history, P.stats, P.max_hist);
        if (st->ticks >= max_ticks && st->phase != PH_DONE) {  // never spin forever on the device
            st->error = 1;
            st->phase = PH_DONE;
        }
    }
}

// align() entry: the first hasConverged() test
__global__ void k_align_begin(const PairDev* __restrict__ pairs, int n_pairs)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n_pairs) align_begin(pairs[p].state, pairs[p].cfg);
}

// ------------------------------------------------------------------------------------------------------------
// cloud move + loop condition
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

// Applies the increment of a finished outer iteration to the source cloud in place (registration.cc:110-112).
// Block (0,0) also publishes "is any pair still running" for the loop around the tick.
__global__ void k_transform(const PairDev* __restrict__ pairs, int n_pairs, int* __restrict__ active_flag,
                            cudaGraphConditionalHandle cond, int use_cond)
{
    const PairDev& P = pairs[blockIdx.y];
    PairState* st = P.state;
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {
        int active = 0;
        for (int p = 0; p < n_pairs; ++p) active |= (pairs[p].state->phase != PH_DONE);
        *active_flag = active;
        if (use_cond) cudaGraphSetConditional(cond, active ? 1u : 0u);
    }
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
