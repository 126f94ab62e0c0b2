// ppcr_capi.cu -- host side of libppcr_cuda.so: device memory, the tick driver (host-stepped or a CUDA-graph
// WHILE loop), and the extern "C" entry points declared in include/ppcr.h.
//
// Nothing in this file computes on the CPU: it sizes buffers, launches the kernels of ppcr_kernels.cuh and copies
// results back.  If no CUDA device is usable every entry point fails (PPCR_ERR_NO_DEVICE); there is no fallback.
#include "../../include/ppcr.h"

#include <cub/device/device_radix_sort.cuh>
#include <cuda_profiler_api.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <climits>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unistd.h>
#include <vector>

#include "ppcr_kernels.cuh"

using namespace ppcr;

static_assert(sizeof(ppcr_iter_stats) == sizeof(IterStats), "stats layout");

// ------------------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------------------

static thread_local std::string g_last_error;

static ppcr_status fail(ppcr_status code, const std::string& msg)
{
    g_last_error = msg;
    return code;
}

struct CudaError {
    cudaError_t err;
    const char* what;
    int line;
};

#define CK(call)                                                 \
    do {                                                         \
        cudaError_t e__ = (call);                                \
        if (e__ != cudaSuccess) throw CudaError{e__, #call, __LINE__}; \
    } while (0)

static ppcr_status translate(const CudaError& e)
{
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at ppcr_capi.cu:%d: %s", static_cast<int>(e.err),
             cudaGetErrorString(e.err), e.line, e.what);
    cudaGetLastError();
    if (e.err == cudaErrorNoDevice || e.err == cudaErrorInsufficientDriver || e.err == cudaErrorInvalidDevice)
        return fail(PPCR_ERR_NO_DEVICE, buf);
    return fail(PPCR_ERR_CUDA, buf);
}

struct StatusError {
    ppcr_status code;
    std::string msg;
};

// ------------------------------------------------------------------------------------------------------------
// device buffers (grow-only, so batch slots can be reused without re-allocating)
// ------------------------------------------------------------------------------------------------------------

// Stream-ordered allocations from the device's default memory pool (release threshold raised to "never"), so that
// creating and destroying handles in a loop -- one per scan pair -- re-uses the same blocks without a device-wide
// synchronisation or a trip to the driver's allocator.
static thread_local cudaStream_t g_alloc_stream = nullptr;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaStream_t owner = nullptr;
    bool plain = false;  // cudaMalloc instead of the pool (memory exported over CUDA IPC)
    void reserve(size_t n)
    {
        if (n <= cap) return;
        release();
        size_t want = n + n / 8 + 64;
        if (plain) {
            CK(cudaMalloc(&p, want * sizeof(T)));
        } else {
            owner = g_alloc_stream;
            CK(cudaMallocAsync(&p, want * sizeof(T), owner));
        }
        cap = want;
    }
    void release()
    {
        if (p) {
            if (plain) cudaFree(p);
            else cudaFreeAsync(p, owner);
        }
        p = nullptr;
        cap = 0;
    }
};

// PPCR_TRACE=1: wall-clock trace of the set-up phases on stderr (each mark synchronises the stream first)
struct Trace {
    bool on;
    cudaStream_t st;
    std::chrono::steady_clock::time_point t0;
    explicit Trace(cudaStream_t s) : on(getenv("PPCR_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* what)
    {
        if (!on) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[ppcr trace] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static int ceil_div(long long a, long long b) { return static_cast<int>((a + b - 1) / b); }

static int g_sm_count = 0;

// The search kernel.  PPCR_SEARCH_VARIANT selects a tuning variant (ppcr_kernels.cuh, SearchList) and PPCR_SEARCH_CAP the
// slots per query of the collect + select variant's column; unset = the product (max-heap column of m slots).
using SearchKernel = void (*)(const PairDev*);
static int search_variant()
{
    static const int variant = getenv("PPCR_SEARCH_VARIANT") ? atoi(getenv("PPCR_SEARCH_VARIANT")) : 0;
    return variant;
}
static bool search_collects() { return search_variant() == 16; }
static SearchKernel search_kernel()
{
    switch (search_variant()) {
        case 1: return k_search<1>;
        case 3: return k_search<3>;
        case 4: return k_search<4>;
        case 16: return k_search<16>;
        case 32: return k_search<32>;  // without the exact warm bound from the previous neighbours
        default: return k_search<0>;
    }
}
// slots per query in the search kernel's shared-memory column
static int search_cap(int max_nn)
{
    if (!search_collects()) return heap_slots(max_nn);
    static const int forced = getenv("PPCR_SEARCH_CAP") ? atoi(getenv("PPCR_SEARCH_CAP")) : 0;
    return std::max(forced > 0 ? forced : 24, max_nn + 4);
}
constexpr int kSearchQueuedMaxM = 32;  // k_search_q's shared-memory heap columns: m KiB per block on top of 35 KiB of queues
constexpr size_t kEvalSmem = static_cast<size_t>(kNSum) * kEvalThreads * sizeof(double);  // per-thread moment columns
constexpr int kDefaultLeafCap = 32;
constexpr int kTreeSmallMax = 32768;  // targets up to this many points are built by one block in one launch
constexpr int kEvalBatchBlocks = 4;  // evaluation blocks per SM of a batch lane (see pair_setup)

// kernel-launch bookkeeping for ppcr_get_stage_times: every launch site outside the tick adds to the engine the
// calling thread is currently working for
static thread_local int32_t g_launch_sink_dummy = 0;
static thread_local int32_t* g_launch_sink = &g_launch_sink_dummy;
static inline void note_launches(int n) { *g_launch_sink += n; }

// ------------------------------------------------------------------------------------------------------------
// one pair: its buffers and the host mirror of its PairDev
// ------------------------------------------------------------------------------------------------------------

struct Pair {
    DevBuf<float4> src, tgt_raw, tgt_sorted, tmp_cloud;
    DevBuf<int> nbr_cnt, scan_sums;
    DevBuf<int> nbr_pos, inv_perm, group_ticket;
    DevBuf<TreeNode> nodes;
    DevBuf<TreeCounters> tree_counters;
    DevBuf<unsigned long long> sort_keys[2];
    DevBuf<unsigned> sort_vals[2];
    DevBuf<unsigned char> sort_tmp;
    DevBuf<float> nbr_d2, nbr_kth;
    DevBuf<double> partials, history;
    DevBuf<PairState> state;
    DevBuf<Config> cfg;
    DevBuf<IterStats> stats;
    DevBuf<unsigned> scratch_u;
    DevBuf<unsigned long long> scratch_ull;
    PairDev dev{};
    Config hcfg{};
    int64_t n_src = 0, n_tgt = 0;
    bool want_d2 = false;
    void release()
    {
        src.release(); tgt_raw.release(); tgt_sorted.release(); tmp_cloud.release();
        nodes.release(); tree_counters.release(); sort_keys[0].release(); sort_keys[1].release();
        sort_vals[0].release(); sort_vals[1].release(); sort_tmp.release(); nbr_pos.release(); inv_perm.release(); group_ticket.release(); nbr_cnt.release();
        scan_sums.release(); nbr_d2.release(); nbr_kth.release();
        partials.release(); history.release(); state.release(); cfg.release(); stats.release();
        scratch_u.release(); scratch_ull.release();
    }
};

struct EventPair {
    cudaEvent_t a, b;
    int stage;
};

struct Engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    ppcr_params params{};
    ppcr_options opts{};
    std::vector<Pair> pairs;
    DevBuf<PairDev> d_pairs;
    DevBuf<LoopCtl> d_loop;
    int* h_active = nullptr;  // pinned, shared by the handles of a host thread (pinned_flag)
    int eval_blocks_per_sm = PPCR_EVAL_MIN_BLOCKS, search_blocks_per_sm = 8;
    size_t search_smem = 0;  // the search kernel's per-block heap columns
    size_t eval_smem = kEvalSmem;  // float64 path: moment columns; float32 path: the staged tiles (set in engine_commit)
    // launch geometry (capacity based, so a captured graph stays valid while the slots are refilled)
    int max_tiles = 1, max_eval_blocks = 1, max_tr_blocks = 1;
    bool skip_search = false;
    bool search_queued = false;  // searches after a cloud move go through k_search_q
    size_t search_q_smem_bytes = 0;
    int q_tiles = 1;
    DevBuf<unsigned char> q_scratch;  // k_search_q's task / candidate queues
    int max_ticks = 0;
    // graph driver
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphConditionalHandle cond = 0, cond_search = 0;
    bool graph_ready = false;
    bool graph_failed = false;
    // stage timing
    std::vector<EventPair> events;
    ppcr_stage_times times{};
    // sharded mode
    int rank = 0, world = 1;
    bool batch_lane = false;  // one of several lanes of ppcr_align_batch working on the same device
    int requested_max_neighbours = 0;  // what the caller asked for (params.max_neighbours holds the row capacity)
    // L2 flush buffer for ppcr_time_kernel
    DevBuf<float4> flush;

    ~Engine()
    {
        if (g_launch_sink == &times.total_launches) g_launch_sink = &g_launch_sink_dummy;
        cudaSetDevice(device);
        const bool trace = getenv("PPCR_TRACE") != nullptr;
        auto t0 = std::chrono::steady_clock::now();
        auto mark = [&](const char* what) {
            if (!trace) return;
            const auto t1 = std::chrono::steady_clock::now();
            fprintf(stderr, "[ppcr trace] destroy: %-19s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
            t0 = t1;
        };
        for (auto& e : events) {
            cudaEventDestroy(e.a);
            cudaEventDestroy(e.b);
        }
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        if (graph) cudaGraphDestroy(graph);
        mark("events + graph");
        for (auto& p : pairs) p.release();
        d_pairs.release();
        d_loop.release();
        flush.release();
        q_scratch.release();
        mark("stream-ordered frees");
        if (stream) cudaStreamSynchronize(stream);  // the frees above are stream-ordered
        mark("stream synchronize");
        if (own_stream && stream) cudaStreamDestroy(stream);
        mark("stream destroy");
        if (g_alloc_stream == stream) g_alloc_stream = nullptr;
    }
};

struct ppcr_handle {
    Engine eng;
};

// ---- sharded pairs: the mailboxes of the moment exchange live as long as the process ----
// One small cudaMalloc block per device, exported over CUDA IPC once; the mappings of the peers' blocks are kept too.  A
// sharded handle therefore costs no allocation, no IPC open and no implicit device synchronisation (allocating, exporting and
// mapping per handle was where one repetition in eight of the 10M-point pair lost 100-300 ms).
struct ShardToken {  // what ppcr_shard_export hands out (PPCR_SHARD_TOKEN_BYTES, zero padded)
    cudaIpcMemHandle_t ipc;
    uint64_t pid;     // a peer in the same process uses `ptr` directly (cudaIpcOpenMemHandle refuses the exporting process)
    uint64_t ptr;
    int32_t device;
    uint32_t epoch;   // proposed stamp epoch
};
static_assert(sizeof(ShardToken) <= PPCR_SHARD_TOKEN_BYTES, "token size");
constexpr size_t kMailboxDoubles = 2ull * 8 * kMailDoubles;  // two alternating slots x up to 8 ranks
struct DeviceMailbox {
    double* p = nullptr;
    cudaIpcMemHandle_t ipc{};
};
struct ShardGlobals {
    std::mutex mu;
    DeviceMailbox box[64];
    std::map<std::string, void*> opened;  // peer mailboxes of other processes, by IPC handle bytes
    uint32_t epoch = 0;
};
static ShardGlobals g_shard;

// ------------------------------------------------------------------------------------------------------------
// device selection
// ------------------------------------------------------------------------------------------------------------

// Per-device one-time set-up (architecture check, kernel attributes, memory pool); a handle per scan pair must not pay
// for device queries again.
struct DeviceInfo {
    bool ready = false;
    int sm_count = 0;
    int clock_khz = 0;  // (queried once: cudaDevAttrClockRate takes tens of milliseconds on a busy host)
};
static DeviceInfo g_devices[64];
static std::mutex g_devices_mutex;

static void select_device(int device)
{
    if (device < 0 || device >= 64) throw StatusError{PPCR_ERR_NO_DEVICE, "device ordinal out of range"};
    {
        std::lock_guard<std::mutex> lock(g_devices_mutex);
        DeviceInfo& info = g_devices[device];
        if (!info.ready) {
            int count = 0;
            cudaError_t e = cudaGetDeviceCount(&count);
            if (e != cudaSuccess || count == 0) {
                cudaGetLastError();
                throw StatusError{PPCR_ERR_NO_DEVICE, "no CUDA device is visible; libppcr_cuda has no CPU fallback"};
            }
            if (device >= count) throw StatusError{PPCR_ERR_NO_DEVICE, "device ordinal out of range"};
            CK(cudaSetDevice(device));
            int major = 0, sms = 0;
            CK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
            CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
            if (major != 10) {
                cudaDeviceProp prop;
                CK(cudaGetDeviceProperties(&prop, device));
                throw StatusError{PPCR_ERR_NO_DEVICE, std::string("libppcr_cuda is built for sm_100a only; device is ") + prop.name};
            }
            CK(cudaFuncSetAttribute(k_evalctl<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kEvalSmem)));
            CK(cudaFuncSetAttribute(k_evalctl<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kEvalSmem)));
            cudaMemPool_t pool;
            CK(cudaDeviceGetDefaultMemPool(&pool, device));
            unsigned long long keep = ~0ull;
            CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
            info.sm_count = sms;
            CK(cudaDeviceGetAttribute(&info.clock_khz, cudaDevAttrClockRate, device));
            info.ready = true;
        }
        g_sm_count = info.sm_count;
    }
    CK(cudaSetDevice(device));
}

// one small pinned word per host thread for the "still running?" read-back of the host-stepped driver
struct PinnedWord {
    int* p = nullptr;
    ~PinnedWord()
    {
        if (p) cudaFreeHost(p);  // a lane thread of ppcr_align_batch ends with its call
    }
};
static int* pinned_flag()
{
    static thread_local PinnedWord w;
    if (!w.p) CK(cudaMallocHost(&w.p, 4 * sizeof(int)));
    return w.p;
}

// ------------------------------------------------------------------------------------------------------------
// scans, grid build
// ------------------------------------------------------------------------------------------------------------

// in-place exclusive scan of data[0..n); the grand total is also written to data[n] when write_total
static void exclusive_scan(int* data, int n, DevBuf<int>& sums, bool write_total, cudaStream_t st)
{
    const int n_blocks = ceil_div(n, kScanTile);
    sums.reserve(static_cast<size_t>(n_blocks) + 1);
    k_scan_local<<<n_blocks, kScanThreads, 0, st>>>(data, n, sums.p);
    k_scan_sums<<<1, 1024, 0, st>>>(sums.p, n_blocks, write_total ? data + n : nullptr);
    k_scan_add<<<n_blocks, kScanThreads, 0, st>>>(data, n, sums.p, 0);
    CK(cudaGetLastError());
    note_launches(3);
}

struct Bbox {
    float lo[3], hi[3];
};

static Bbox cloud_bbox(const float4* pts, int n, Pair& P, cudaStream_t st)
{
    P.scratch_u.reserve(8);
    unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    CK(cudaMemcpyAsync(P.scratch_u.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
    k_bbox<<<std::min(ceil_div(n, 256), 4 * std::max(g_sm_count, 1)), 256, 0, st>>>(pts, n, P.scratch_u.p);
    CK(cudaGetLastError());
    note_launches(1);
    unsigned out[6];
    CK(cudaMemcpyAsync(out, P.scratch_u.p, sizeof(out), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    Bbox b;
    for (int k = 0; k < 3; ++k) {
        b.lo[k] = ord2f(out[k]);
        b.hi[k] = ord2f(out[3 + k]);
    }
    return b;
}

// Morton keys -> radix sort -> permutation; returns the sorted permutation in P.sort_vals[1]
static void morton_sort(Pair& P, const float4* pts, int n, const TreeGeom& g, cudaStream_t st)
{
    P.sort_keys[0].reserve(n); P.sort_keys[1].reserve(n);
    P.sort_vals[0].reserve(n); P.sort_vals[1].reserve(n);
    k_tree_keys<<<ceil_div(n, 256), 256, 0, st>>>(pts, n, g, P.sort_keys[0].p, P.sort_vals[0].p);
    CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, P.sort_keys[0].p, P.sort_keys[1].p, P.sort_vals[0].p,
                                       P.sort_vals[1].p, n, 0, 3 * kTreeBits, st));
    P.sort_tmp.reserve(tmp_bytes);
    CK(cub::DeviceRadixSort::SortPairs(P.sort_tmp.p, tmp_bytes, P.sort_keys[0].p, P.sort_keys[1].p, P.sort_vals[0].p,
                                       P.sort_vals[1].p, n, 0, 3 * kTreeBits, st));
    note_launches(1 + 8);  // keys + the radix sort's passes (histogram, scan, 6 x onesweep)
}

static TreeGeom make_tree_geom(const Bbox& b, int leaf_cap, int n)
{
    TreeGeom g{};
    double span = 0.0, mag = 0.0;
    for (int k = 0; k < 3; ++k) {
        span = std::max(span, static_cast<double>(b.hi[k]) - b.lo[k]);
        mag = std::max({mag, std::fabs(static_cast<double>(b.lo[k])), std::fabs(static_cast<double>(b.hi[k]))});
    }
    span = std::max(span, 1e-6 * std::max(mag, 1e-30)) * (1.0 + 1e-5);
    if (!(span > 0.0)) span = 1.0;
    g.ox = b.lo[0];
    g.oy = b.lo[1];
    g.oz = b.lo[2];
    g.inv_hf = static_cast<float>(static_cast<double>(1 << kTreeBits) / span);
    g.hf = static_cast<float>(1.0 / static_cast<double>(g.inv_hf));
    // bound on the float fuzz of binning a point ((v - o) * inv_hf: three roundings on a value <= span) plus the
    // rounding of node centres (o + k * half: two roundings on a value <= mag + span), with a wide margin
    g.slack = static_cast<float>(1e-6 * span + 1e-6 * (mag + span) + 1e-30);
    g.leaf_cap = leaf_cap;
    g.n_nodes_cap = static_cast<int>(std::min<long long>(64ll + 8ll * (2ll * n / std::max(leaf_cap, 1) + 8), 1ll << 28));
    return g;
}

// Morton sort of the target + level-by-level octree construction (replaces the kd-tree build)
static void build_target_tree(Engine& E, Pair& P, int leaf_cap)
{
    cudaStream_t st = E.stream;
    const int n = static_cast<int>(P.n_tgt);
    const Bbox bb = cloud_bbox(P.tgt_raw.p, n, P, st);
    for (int k = 0; k < 3; ++k)
        if (!std::isfinite(bb.lo[k]) || !std::isfinite(bb.hi[k]))
            throw StatusError{PPCR_ERR_INVALID, "target cloud contains non-finite coordinates (pcl::KdTreeFLANN leaves such points out "
                                                "of its tree when the cloud is flagged !is_dense: the C++ class does, the C ABI takes "
                                                "finite targets)"};
    const TreeGeom g = make_tree_geom(bb, leaf_cap, n);
    morton_sort(P, P.tgt_raw.p, n, g, st);
    P.tgt_sorted.reserve(n);
    P.inv_perm.reserve(n);
    k_tree_gather<<<ceil_div(n, 256), 256, 0, st>>>(P.tgt_raw.p, P.sort_vals[1].p, n, 0, P.tgt_sorted.p, P.inv_perm.p);
    CK(cudaGetLastError());
    P.nodes.reserve(static_cast<size_t>(g.n_nodes_cap) + 8);
    P.tree_counters.reserve(1);
    CK(cudaMemsetAsync(P.nodes.p, 0xff, (static_cast<size_t>(g.n_nodes_cap) + 8) * sizeof(TreeNode), st));
    static const int small_max = getenv("PPCR_TREE_SMALL") ? atoi(getenv("PPCR_TREE_SMALL")) : kTreeSmallMax;  // (0: always level by level)
    if (n <= small_max) {
        k_tree_build_small<<<1, kTreeSmallThreads, 0, st>>>(P.nodes.p, n, g, P.sort_keys[1].p, P.tgt_sorted.p, P.tree_counters.p);
        CK(cudaGetLastError());
        note_launches(1);
        P.dev.tree = g;
        P.dev.nodes = P.nodes.p;
        P.dev.tgt_sorted = P.tgt_sorted.p;
        P.dev.tgt_raw = P.tgt_raw.p;
        P.dev.inv_perm = P.inv_perm.p;
        return;
    }
    k_tree_root<<<1, 32, 0, st>>>(P.nodes.p, n, g, P.tree_counters.p);
    long long level_nodes = 1;
    for (int level = 0; level < kTreeBits; ++level) {
        const int blocks = static_cast<int>(std::min<long long>((level_nodes + 127) / 128, 8ll * std::max(g_sm_count, 1)));
        k_tree_split_level<<<blocks, 128, 0, st>>>(g, P.sort_keys[1].p, P.nodes.p, P.tree_counters.p, level);
        level_nodes = std::min<long long>(level_nodes * 8, g.n_nodes_cap);
    }
    k_tree_finalize<<<4 * std::max(g_sm_count, 1), 256, 0, st>>>(P.nodes.p, P.tree_counters.p, g.n_nodes_cap);
    k_tree_leaf_boxes<<<8 * std::max(g_sm_count, 1), 256, 0, st>>>(P.nodes.p, P.tgt_sorted.p, P.tree_counters.p, g.n_nodes_cap);
    CK(cudaGetLastError());
    note_launches(4 + kTreeBits);
    P.dev.tree = g;
    P.dev.nodes = P.nodes.p;
    P.dev.tgt_sorted = P.tgt_sorted.p;
    P.dev.tgt_raw = P.tgt_raw.p;
    P.dev.inv_perm = P.inv_perm.p;
}

// Sorts the (tagged) source cloud along the target tree's Morton curve: the queries of one warp then open the same
// nodes.  Done once; the cloud moves rigidly and by little, so the locality survives the outer iterations.
static void sort_source(Engine& E, Pair& P)
{
    cudaStream_t st = E.stream;
    const int n = static_cast<int>(P.n_src);
    if (n <= 1) return;
    morton_sort(P, P.src.p, n, P.dev.tree, st);
    P.tmp_cloud.reserve(n);
    k_tree_gather<<<ceil_div(n, 256), 256, 0, st>>>(P.src.p, P.sort_vals[1].p, n, 1, P.tmp_cloud.p, nullptr);
    CK(cudaGetLastError());
    note_launches(1);
    std::swap(P.src, P.tmp_cloud);
}

// ------------------------------------------------------------------------------------------------------------
// voxel filter
// ------------------------------------------------------------------------------------------------------------

// Filters `in` (n points) into `out`; returns the filtered size, or -1 when PCL would refuse the leaf size.
static int64_t voxel_filter_device(const float4* in, int64_t n, double leaf_d, DevBuf<float4>& out, Pair& scratch,
                                   cudaStream_t st)
{
    if (n == 0) return 0;
    const float leaf = static_cast<float>(leaf_d);
    const float inv = 1.0f / leaf;
    const Bbox bb = cloud_bbox(in, static_cast<int>(n), scratch, st);
    for (int k = 0; k < 3; ++k)
        if (!std::isfinite(bb.lo[k]) || !std::isfinite(bb.hi[k]))
            throw StatusError{PPCR_ERR_INVALID, "a cloud to be voxel-filtered contains non-finite coordinates (PCL drops them when "
                                                "the cloud is flagged !is_dense: the C++ class does, the C ABI takes finite clouds)"};
    int64_t d[3];
    VoxelGeom vg{};
    vg.inv_leaf = inv;
    int div[3];
    for (int k = 0; k < 3; ++k) {
        d[k] = static_cast<int64_t>((bb.hi[k] - bb.lo[k]) * inv) + 1;
        vg.minb[k] = static_cast<int>(std::floor(bb.lo[k] * inv));
        div[k] = static_cast<int>(std::floor(bb.hi[k] * inv)) - vg.minb[k] + 1;
    }
    // formed in double: exact near the threshold, and immune to int64 wrap-around for absurd extents
    if (static_cast<double>(d[0]) * static_cast<double>(d[1]) * static_cast<double>(d[2]) > static_cast<double>(INT_MAX)) return -1;
    vg.mul[0] = 1;
    vg.mul[1] = div[0];
    vg.mul[2] = div[0] * div[1];
    DevBuf<unsigned> keys, vals, keys2, vals2;
    DevBuf<int> head;
    DevBuf<unsigned char> tmp;
    const int ni = static_cast<int>(n);
    keys.reserve(n); vals.reserve(n); keys2.reserve(n); vals2.reserve(n); head.reserve(n + 2);
    k_voxel_keys<<<ceil_div(n, 256), 256, 0, st>>>(in, ni, vg, keys.p, vals.p);
    CK(cudaGetLastError());
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys.p, keys2.p, vals.p, vals2.p, ni, 0, 32, st));
    tmp.reserve(tmp_bytes);
    CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, keys.p, keys2.p, vals.p, vals2.p, ni, 0, 32, st));  // stable
    k_voxel_heads<<<ceil_div(n, 256), 256, 0, st>>>(keys2.p, ni, head.p);
    CK(cudaGetLastError());
    exclusive_scan(head.p, ni, scratch.scan_sums, true, st);
    int n_out = 0;
    CK(cudaMemcpyAsync(&n_out, head.p + ni, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    out.reserve(static_cast<size_t>(n_out) + 1);
    k_voxel_mean<<<ceil_div(n, 256), 256, 0, st>>>(in, keys2.p, vals2.p, head.p, ni, out.p);
    CK(cudaGetLastError());
    note_launches(3 + 6);  // keys, heads, mean + the radix sort's passes (histogram, scan, 4 x onesweep)
    CK(cudaStreamSynchronize(st));
    keys.release(); vals.release(); keys2.release(); vals2.release(); head.release(); tmp.release();
    return n_out;
}

// ------------------------------------------------------------------------------------------------------------
// pair / engine setup
// ------------------------------------------------------------------------------------------------------------

// Rows of the association hold at most kMaxRow neighbours.  A request for more -- max_neighbours <= 0 is pcl's "every target
// within the radius" (registration.cc:74-75 passes it straight to radiusSearch), and so is anything above the target size -- is
// served with rows of kMaxRow as long as no query has that many targets within the radius: every row then holds ALL of its
// in-radius targets, which is what the reference would have returned.  A row that fills up stops the registration with
// PPCR_ERR_UNSUPPORTED instead of silently truncating it.
constexpr int kMaxRow = 128;
static bool wide_rows(int max_neighbours) { return max_neighbours <= 0 || max_neighbours > kMaxRow; }

static void validate_params(const ppcr_params& p, bool allow_wide)
{
    if (wide_rows(p.max_neighbours) && !allow_wide)
        throw StatusError{PPCR_ERR_UNSUPPORTED, "max_neighbours must be in [1,128] at this entry point"};
    if (!(p.radius > 0.0) || !std::isfinite(p.radius)) throw StatusError{PPCR_ERR_INVALID, "radius must be finite and > 0"};
    if (!(p.dof > 0.0)) throw StatusError{PPCR_ERR_INVALID, "dof must be > 0 (+inf selects the Gaussian model)"};
    if (p.n_iter < 0) throw StatusError{PPCR_ERR_INVALID, "n_iter must be >= 0"};
}

static void upload_cloud(DevBuf<float4>& dst, const float* src, int64_t n, bool on_device, cudaStream_t st)
{
    dst.reserve(static_cast<size_t>(std::max<int64_t>(n, 1)));
    if (n > 0)
        CK(cudaMemcpyAsync(dst.p, src, static_cast<size_t>(n) * sizeof(float4),
                           on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
}

// Everything the reference constructor does (registration.cc:15-49) plus the grid build that replaces the
// per-iteration kd-tree construction (:66-67; the target never changes after the constructor).
static void pair_setup(Engine& E, Pair& P, const float* src, int64_t n_src, const float* tgt, int64_t n_tgt,
                       bool on_device)
{
    if (n_src < 0 || n_tgt < 0 || (n_src > 0 && !src) || (n_tgt > 0 && !tgt)) throw StatusError{PPCR_ERR_INVALID, "null cloud"};
    if (n_src > INT_MAX / 2 || n_tgt > INT_MAX / 2) throw StatusError{PPCR_ERR_UNSUPPORTED, "clouds above 2^30 points are not supported"};
    cudaStream_t st = E.stream;
    Trace tr(st);
    const ppcr_params& prm = E.params;
    upload_cloud(P.src, src, n_src, on_device, st);
    upload_cloud(P.tgt_raw, tgt, n_tgt, on_device, st);
    P.n_src = n_src;
    P.n_tgt = n_tgt;
    tr.mark("upload clouds");
    if (prm.source_filter_size > 0 && n_src > 0) {
        int64_t k = voxel_filter_device(P.src.p, n_src, prm.source_filter_size, P.tmp_cloud, P, st);
        if (k >= 0) {
            std::swap(P.src, P.tmp_cloud);
            P.n_src = k;
        }
    }
    if (prm.target_filter_size > 0 && n_tgt > 0) {
        int64_t k = voxel_filter_device(P.tgt_raw.p, n_tgt, prm.target_filter_size, P.tmp_cloud, P, st);
        if (k >= 0) {
            std::swap(P.tgt_raw, P.tmp_cloud);
            P.n_tgt = k;
        }
    }
    PairDev& D = P.dev;
    D.n_src = static_cast<int>(P.n_src);
    D.n_tgt = static_cast<int>(P.n_tgt);
    // rows are padded to whole tiles of the evaluation (bulk copies read whole tiles; the padding rows carry count 0)
    D.n_pad = (D.n_src + kEvalFastThreads - 1) / kEvalFastThreads * kEvalFastThreads;
    if (D.n_pad == 0) D.n_pad = kEvalFastThreads;
    D.m = static_cast<int>(std::min<int64_t>(prm.max_neighbours, std::max<int64_t>(P.n_tgt, 1)));
    // wide rows: a row that reaches the capacity may have lost neighbours -- unless the capacity is the whole target
    const int64_t wanted = E.requested_max_neighbours <= 0 ? P.n_tgt : std::min<int64_t>(E.requested_max_neighbours, P.n_tgt);
    D.overflow_at = wanted > D.m ? D.m : INT_MAX;
    D.search_cap = search_cap(prm.max_neighbours);
    D.r2f = static_cast<float>(prm.radius * prm.radius);
    D.src = P.src.p;
    if (D.n_src > 0) {
        k_tag_index<<<ceil_div(D.n_src, 256), 256, 0, st>>>(P.src.p, D.n_src);
        CK(cudaGetLastError());
        note_launches(1);
    }
    tr.mark("voxel filters + tag");
    static const int env_leaf = getenv("PPCR_LEAF_CAP") ? atoi(getenv("PPCR_LEAF_CAP")) : 0;  // tuning
    const int leaf_cap = E.opts.leaf_capacity > 0 ? E.opts.leaf_capacity : env_leaf > 0 ? env_leaf : kDefaultLeafCap;
    if (P.n_tgt > 0) {
        build_target_tree(E, P, leaf_cap);
        tr.mark("target tree");
    } else {
        // an empty target: a root with no points, every search returns nothing
        Bbox bb{};
        bb.hi[0] = bb.hi[1] = bb.hi[2] = 1.f;
        P.nodes.reserve(16);
        P.tgt_sorted.reserve(1);
        const TreeGeom g = make_tree_geom(bb, leaf_cap, 0);
        CK(cudaMemsetAsync(P.nodes.p, 0xff, 16 * sizeof(TreeNode), st));
        D.tree = g;
        D.nodes = P.nodes.p;
        D.tgt_sorted = P.tgt_sorted.p;
        D.tgt_raw = P.tgt_raw.p;
        P.inv_perm.reserve(1);
        D.inv_perm = P.inv_perm.p;
    }
    sort_source(E, P);
    D.src = P.src.p;
    tr.mark("source sort");
    const size_t plane = static_cast<size_t>(D.m) * D.n_pad;
    P.nbr_pos.reserve(plane);
    P.nbr_cnt.reserve(D.n_pad);
    P.nbr_kth.reserve(D.n_pad);
    D.nbr_kth = P.nbr_kth.p;
    CK(cudaMemsetAsync(P.nbr_kth.p, 0x7f, static_cast<size_t>(D.n_pad) * sizeof(float), st));  // "nothing known yet"
    if (P.want_d2) P.nbr_d2.reserve(plane);
    D.nbr_pos = P.nbr_pos.p;
    D.nbr_d2 = P.want_d2 ? P.nbr_d2.p : nullptr;
    D.nbr_cnt = P.nbr_cnt.p;
    CK(cudaMemsetAsync(P.nbr_cnt.p, 0, static_cast<size_t>(D.n_pad) * sizeof(int), st));
    {
        const bool fast = !E.opts.exact_weights;
        // (the asynchronous-gather variant stages the target points too: 64 KB of shared memory per block at m = 10, three blocks per SM)
        int per_sm = fast ? (eval_async(E.params.max_neighbours) ? std::min(kEvalFastBlocks, kEvalAsyncBlocks) : kEvalFastBlocks) : E.eval_blocks_per_sm;
        // A lane of a batch shares the device with the other lanes' kernels: a smaller evaluation grid leaves them registers and
        // shared memory to run beside it (PPCR_EVAL_PER_SM / PPCR_EVAL_PER_SM_BATCH: tuning)
        static const int env_single = getenv("PPCR_EVAL_PER_SM") ? atoi(getenv("PPCR_EVAL_PER_SM")) : 0;
        static const int env_batch = getenv("PPCR_EVAL_PER_SM_BATCH") ? atoi(getenv("PPCR_EVAL_PER_SM_BATCH")) : 0;
        if (fast) {
            const int want = E.batch_lane ? (env_batch > 0 ? env_batch : kEvalBatchBlocks) : (env_single > 0 ? env_single : per_sm);
            per_sm = std::max(1, std::min(per_sm, want));
        }
        D.n_eval_blocks = std::max(1, std::min(ceil_div(std::max(D.n_src, 1), eval_threads(fast)), per_sm * std::max(g_sm_count, 1)));
    }
    D.n_eval_groups = ceil_div(D.n_eval_blocks, kFoldGroup);
    P.partials.reserve(static_cast<size_t>(D.n_eval_blocks + D.n_eval_groups) * kNSum);
    D.partials = P.partials.p;
    D.group_partials = P.partials.p + static_cast<size_t>(D.n_eval_blocks) * kNSum;
    CK(cudaMemsetAsync(P.partials.p, 0, static_cast<size_t>(D.n_eval_blocks + D.n_eval_groups) * kNSum * sizeof(double), st));
    P.group_ticket.reserve(D.n_eval_groups);
    D.group_ticket = P.group_ticket.p;
    CK(cudaMemsetAsync(P.group_ticket.p, 0, static_cast<size_t>(D.n_eval_groups) * sizeof(int), st));
    D.max_hist = std::max(1, prm.n_iter);
    P.history.reserve(static_cast<size_t>(D.max_hist) * 32);  // accumulated poses, then the increments
    P.stats.reserve(D.max_hist);
    P.state.reserve(1);
    P.cfg.reserve(1);
    D.history = P.history.p;
    D.stats = P.stats.p;
    D.state = P.state.p;
    D.cfg = P.cfg.p;
    Config& c = P.hcfg;
    for (int k = 0; k < 4; ++k) c.x0[k] = prm.initial_rotation[k];
    for (int k = 0; k < 3; ++k) c.x0[4 + k] = prm.initial_translation[k];
    c.function_tolerance = E.opts.function_tolerance > 0 ? E.opts.function_tolerance : 10e-6;
    c.cost_drop_thresh = prm.cost_drop_thresh;
    c.n_cost_drop_it = prm.n_cost_drop_it;
    c.dof = prm.dof;
    c.n_iter = prm.n_iter;
    c.max_lm_iterations = INT_MAX;
    c.is_normal = !(prm.dof < DBL_MAX);
    c.fast_weights = E.opts.exact_weights ? 0 : 1;
    D.wcfg = make_weight_cfg(prm.dof);
    PairState hs;
    memset(&hs, 0, sizeof(hs));
    state_init(&hs, &c);
    CK(cudaMemcpyAsync(P.cfg.p, &c, sizeof(c), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(P.state.p, &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
    D.dump_w = nullptr;
    D.mailbox = nullptr;
    D.mail_base = 0.0;
    D.rank = 0;
    D.world = 1;
    D.spin_limit = 4000000000ll;  // ~2 s of SM clocks
    CK(cudaStreamSynchronize(st));  // host temporaries (hs, c) must outlive the copies
    tr.mark("buffers + state");
}

static void engine_init(Engine& E, const ppcr_params& params, const ppcr_options* options, bool allow_wide = false)
{
    E.params = params;
    g_launch_sink = &E.times.total_launches;
    if (options) E.opts = *options; else memset(&E.opts, 0, sizeof(E.opts));
    validate_params(params, allow_wide);
    E.requested_max_neighbours = params.max_neighbours;
    if (wide_rows(params.max_neighbours)) E.params.max_neighbours = kMaxRow;  // every size below derives from this
    E.device = E.opts.device;
    select_device(E.device);
    if (E.opts.stream) {
        E.stream = static_cast<cudaStream_t>(E.opts.stream);
        E.own_stream = false;
    } else {
        CK(cudaStreamCreateWithFlags(&E.stream, cudaStreamNonBlocking));
        E.own_stream = true;
    }
    g_alloc_stream = E.stream;
    E.d_loop.reserve(1);
    CK(cudaMemsetAsync(E.d_loop.p, 0, sizeof(LoopCtl), E.stream));
    E.h_active = pinned_flag();
    if (E.opts.ticks_per_sync <= 0) E.opts.ticks_per_sync = 4;
    const char* env = getenv("PPCR_DRIVER");
    if (env && E.opts.driver == 0) E.opts.driver = atoi(env);
}

// publishes the PairDev array and derives the launch geometry
static void engine_commit(Engine& E)
{
    const int np = static_cast<int>(E.pairs.size());
    std::vector<PairDev> host(np);
    int max_m = 1;
    long long max_src = 1;
    int max_eval = 1;
    // The queued search (k_search_q) does every search that follows a cloud move whenever its queues fit (the heap columns in
    // shared memory grow with m; a task names a node in 25 bits): on the 1M-point pair 0.29 against 0.42 ms on a converged
    // pair, 0.50 against 0.61 ms inside the loop.  PPCR_SEARCH_QUEUED=0 switches it off (k_search then does every search).
    // Read at every commit (not cached) so that a test can switch it inside one process.
    const bool queued_wanted = !(getenv("PPCR_SEARCH_QUEUED") && atoi(getenv("PPCR_SEARCH_QUEUED")) == 0);
    bool queued = queued_wanted && search_variant() == 0 && E.params.max_neighbours <= kSearchQueuedMaxM;
    for (int p = 0; p < np; ++p)
        queued = queued && E.pairs[p].dev.tree.n_nodes_cap < (1 << kQNodeBits) && E.pairs[p].dev.n_tgt < (1 << kQNodeBits);
    E.search_queued = queued;
    for (int p = 0; p < np; ++p) {
        E.pairs[p].dev.search_queued = queued ? 1 : 0;
        E.pairs[p].dev.q_cand = getenv("PPCR_Q_CAND") ? atoi(getenv("PPCR_Q_CAND")) : search_q_cand(E.params.max_neighbours);
        E.pairs[p].dev.q_heavy = getenv("PPCR_Q_HEAVY") ? static_cast<float>(atof(getenv("PPCR_Q_HEAVY"))) : 0.5f;  // (0.75 before the leaf boxes; with them 0.5: 10M-point pair 303 -> 278 ms of search, the 1M-point pair and the 120k-point pairs unchanged; 0.35 costs the 1M-point pair 3 %)
        E.pairs[p].dev.q_flags = getenv("PPCR_Q_FLAGS") ? atoi(getenv("PPCR_Q_FLAGS")) : 0;
        E.pairs[p].dev.q_leaves = getenv("PPCR_Q_LEAVES") ? atoi(getenv("PPCR_Q_LEAVES")) : kQTaskPerQuery;
        host[p] = E.pairs[p].dev;
        max_m = std::max(max_m, host[p].m);
        max_src = std::max<long long>(max_src, host[p].n_src);
        max_eval = std::max(max_eval, host[p].n_eval_blocks);
    }
    E.search_smem = static_cast<size_t>(search_cap(E.params.max_neighbours)) * kSearchThreads * sizeof(unsigned long long);
    if (const char* pad = getenv("PPCR_SEARCH_PAD_SMEM")) E.search_smem += static_cast<size_t>(atoi(pad));  // occupancy experiments
    E.eval_smem = E.opts.exact_weights ? kEvalSmem : eval_fast_smem(E.params.max_neighbours);
    {
        // the opt-in shared-memory sizes are per function and process wide: only ever raise them (handles of several host
        // threads, e.g. the lanes of ppcr_align_batch, launch the same kernels with different sizes)
        static std::mutex attr_mutex;
        static size_t eval_fast_max[64] = {}, search_max_of[2][64] = {};
        size_t* search_max = search_max_of[search_collects() ? 1 : 0];
        std::lock_guard<std::mutex> lock(attr_mutex);
        if (!E.opts.exact_weights && E.eval_smem > eval_fast_max[E.device]) {
            CK(cudaFuncSetAttribute(k_evalctl<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(E.eval_smem)));
            eval_fast_max[E.device] = E.eval_smem;
        }
        // (the kernel also has a little static shared memory: opt in from just below the 48 KiB default limit)
        if (E.search_smem > 47 * 1024 && E.search_smem > search_max[E.device]) {
            CK(cudaFuncSetAttribute(search_kernel(), cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(E.search_smem)));
            search_max[E.device] = E.search_smem;
        }
    }
    {   // persistent grid: exactly as many blocks as the device keeps resident
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, search_kernel(), kSearchThreads, E.search_smem));
        E.search_blocks_per_sm = std::max(1, per_sm);
        // tuning: PPCR_SEARCH_BLOCKS caps the resident blocks per SM, PPCR_SEARCH_CARVEOUT (percent) sets the shared-memory
        // carve-out preference (what is not carved out is L1, which the tree walk lives on)
        if (const char* b = getenv("PPCR_SEARCH_BLOCKS")) E.search_blocks_per_sm = std::max(1, std::min(per_sm, atoi(b)));
        if (const char* c = getenv("PPCR_SEARCH_CARVEOUT"))
            CK(cudaFuncSetAttribute(search_kernel(), cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c)));
    }
    int q_tiles = 1;
    if (E.search_queued) {
        E.search_q_smem_bytes = search_q_smem(E.params.max_neighbours);
        static std::mutex q_mutex;
        static size_t q_max[64] = {};
        {
            std::lock_guard<std::mutex> lock(q_mutex);
            if (E.search_q_smem_bytes > q_max[E.device]) {
                CK(cudaFuncSetAttribute(k_search_q, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(E.search_q_smem_bytes)));
                q_max[E.device] = E.search_q_smem_bytes;
            }
        }
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_search_q, kSearchThreads, E.search_q_smem_bytes));
        if (const char* b = getenv("PPCR_SEARCH_Q_BLOCKS")) per_sm = std::min(per_sm, atoi(b));
        q_tiles = std::max(1, std::min(ceil_div(max_src, kSearchChunk), std::max(1, per_sm) * std::max(g_sm_count, 1)));
    }
    const int tiles = std::max(1, std::min(ceil_div(max_src, kSearchChunk), E.search_blocks_per_sm * std::max(g_sm_count, 1)));
    const int trb = std::max(1, std::min(ceil_div(max_src, 256), 8 * std::max(g_sm_count, 1)));
    if (tiles > E.max_tiles || max_eval > E.max_eval_blocks || trb > E.max_tr_blocks || q_tiles > E.q_tiles) {
        E.max_tiles = std::max(E.max_tiles, tiles);
        E.q_tiles = std::max(E.q_tiles, q_tiles);
        E.max_eval_blocks = std::max(E.max_eval_blocks, max_eval);
        E.max_tr_blocks = std::max(E.max_tr_blocks, trb);
        if (E.graph_exec) {  // geometry changed: the captured graph is stale
            cudaGraphExecDestroy(E.graph_exec);
            cudaGraphDestroy(E.graph);
            E.graph_exec = nullptr;
            E.graph = nullptr;
            E.graph_ready = false;
        }
    }
    if (E.search_queued) {  // one slab of queues per block of k_search_q's grid
        E.q_scratch.reserve(static_cast<size_t>(E.q_tiles) * np * search_q_scratch_per_block(E.pairs[0].dev.q_cand));
        for (int p = 0; p < np; ++p) E.pairs[p].dev.q_scratch = host[p].q_scratch = E.q_scratch.p;
    }
    E.d_pairs.reserve(np);
    CK(cudaMemcpyAsync(E.d_pairs.p, host.data(), sizeof(PairDev) * np, cudaMemcpyHostToDevice, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    // generous device-side cap: n_iter outer iterations, each at most ~64 LM evaluations in practice
    const long long cap = (static_cast<long long>(E.params.n_iter) + 2) * 256 + 64;
    E.max_ticks = static_cast<int>(std::min<long long>(cap, INT_MAX / 2));
}

// ------------------------------------------------------------------------------------------------------------
// tick driver
// ------------------------------------------------------------------------------------------------------------

enum Stage { ST_SEARCH = 0, ST_EVAL = 1, ST_CTRL = 2, ST_TRANSFORM = 3 };

static void stage_begin(Engine& E, bool rec, int stage)
{
    if (!rec) return;
    EventPair ep;
    CK(cudaEventCreate(&ep.a));
    CK(cudaEventCreate(&ep.b));
    ep.stage = stage;
    CK(cudaEventRecord(ep.a, E.stream));
    E.events.push_back(ep);
}
static void stage_end(Engine& E, bool rec)
{
    if (!rec) return;
    CK(cudaEventRecord(E.events.back().b, E.stream));
}

static void launch_search(Engine& E)
{
    const int np = static_cast<int>(E.pairs.size());
    dim3 grid(E.max_tiles, np);
    search_kernel()<<<grid, kSearchThreads, E.search_smem, E.stream>>>(E.d_pairs.p);
    // the first search of an align() is k_search's, every later one (fused with the cloud move, tight bounds known)
    // k_search_q's; each returns at once when the search is the other's
    if (E.search_queued) k_search_q<<<dim3(E.q_tiles, np), kSearchThreads, E.search_q_smem_bytes, E.stream>>>(E.d_pairs.p);
}
static int search_launches(const Engine& E) { return E.search_queued ? 2 : 1; }

// weights + moments + (in its last block) reduction, controller and loop condition
static void launch_evalctl(Engine& E, bool use_cond, int probe = 0)
{
    const int np = static_cast<int>(E.pairs.size());
    dim3 grid(E.max_eval_blocks, np);
    const int flags = (use_cond ? 1 : 0) | probe | ((!E.opts.exact_weights && eval_async(E.params.max_neighbours)) ? 8 : 0);
    if (!E.opts.exact_weights)
        k_evalctl<true><<<grid, kEvalFastThreads, E.eval_smem, E.stream>>>(E.d_pairs.p, np, E.d_loop.p, E.cond, E.cond_search, flags, E.max_ticks);
    else
        k_evalctl<false><<<grid, kEvalThreads, E.eval_smem, E.stream>>>(E.d_pairs.p, np, E.d_loop.p, E.cond, E.cond_search, flags, E.max_ticks);
}

static int launches_per_tick(const Engine& E) { return E.skip_search ? 1 : 1 + search_launches(E); }

static void launch_tick(Engine& E, bool use_cond, bool rec)
{
    if (!E.skip_search) {
        stage_begin(E, rec, ST_SEARCH);
        launch_search(E);
        stage_end(E, rec);
    }
    stage_begin(E, rec, ST_EVAL);
    launch_evalctl(E, use_cond);
    stage_end(E, rec);
    CK(cudaGetLastError());
    E.times.ticks += 1;
    E.times.total_launches += launches_per_tick(E);
}

// epilogue of align(): the cloud move of the last outer iteration
static void launch_final_transform(Engine& E)
{
    const int np = static_cast<int>(E.pairs.size());
    k_transform_final<<<dim3(E.max_tr_blocks, np), 256, 0, E.stream>>>(E.d_pairs.p);
    k_transform_done<<<ceil_div(np, 64), 64, 0, E.stream>>>(E.d_pairs.p, np);
    CK(cudaGetLastError());
    E.times.total_launches += 2;
}

static void collect_stage_times(Engine& E)
{
    static const bool trace = getenv("PPCR_TRACE_SEARCH") != nullptr;  // per-launch search times on stderr (tools/)
    for (auto& ep : E.events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ep.a, ep.b) == cudaSuccess) {
            switch (ep.stage) {
                case ST_SEARCH:
                    E.times.search_ms += ms;
                    E.times.search_launches++;
                    if (trace && ms > 0.02f) fprintf(stderr, "[ppcr trace] search launch %.3f ms\n", ms);
                    break;
                case ST_EVAL: E.times.eval_ms += ms; E.times.eval_launches++; break;
                case ST_CTRL: E.times.controller_ms += ms; E.times.controller_launches++; break;
                default: E.times.transform_ms += ms; E.times.transform_launches++; break;
            }
        }
        cudaEventDestroy(ep.a);
        cudaEventDestroy(ep.b);
    }
    E.events.clear();
}

// body graph = one tick, wrapped in a WHILE node whose condition k_evalctl sets on the device
static bool build_graph(Engine& E)
{
    if (E.graph_ready) return true;
    if (E.graph_failed) return false;
    cudaError_t e;
    e = cudaGraphCreate(&E.graph, 0);
    if (e != cudaSuccess) goto bad;
    e = cudaGraphConditionalHandleCreate(&E.cond, E.graph, 1, cudaGraphCondAssignDefault);
    if (e != cudaSuccess) goto bad;
    {
        cudaGraphNodeParams np{};
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = E.cond;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        cudaGraphNode_t node;
        e = cudaGraphAddNode(&node, E.graph, nullptr, 0, &np);
        if (e != cudaSuccess) goto bad;
        cudaGraph_t body = np.conditional.phGraph_out[0];
        // body = one tick: [IF a pair is about to search: k_search] -> k_evalctl.  Three ticks out of four are LM
        // iterations on an unchanged association; the IF node (set by the controller) saves their idle search launch.
        cudaGraphNode_t if_node = nullptr;
        if (!E.skip_search) {
            e = cudaGraphConditionalHandleCreate(&E.cond_search, E.graph, 1, cudaGraphCondAssignDefault);
            if (e != cudaSuccess) goto bad;
            cudaGraphNodeParams ip{};
            ip.type = cudaGraphNodeTypeConditional;
            ip.conditional.handle = E.cond_search;
            ip.conditional.type = cudaGraphCondTypeIf;
            ip.conditional.size = 1;
            e = cudaGraphAddNode(&if_node, body, nullptr, 0, &ip);
            if (e != cudaSuccess) goto bad;
            e = cudaStreamBeginCaptureToGraph(E.stream, ip.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
            if (e != cudaSuccess) goto bad;
            launch_search(E);
            e = cudaStreamEndCapture(E.stream, nullptr);
            if (e != cudaSuccess) goto bad;
        }
        e = cudaStreamBeginCaptureToGraph(E.stream, body, if_node ? &if_node : nullptr, nullptr, if_node ? 1 : 0,
                                          cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) goto bad;
        launch_evalctl(E, true);
        e = cudaStreamEndCapture(E.stream, nullptr);
        if (e != cudaSuccess) goto bad;
        e = cudaGetLastError();
        if (e != cudaSuccess) goto bad;
    }
    e = cudaGraphInstantiate(&E.graph_exec, E.graph, 0);
    if (e != cudaSuccess) goto bad;
    E.graph_ready = true;
    return true;
bad:
    cudaGetLastError();
    if (E.graph_exec) cudaGraphExecDestroy(E.graph_exec);
    if (E.graph) cudaGraphDestroy(E.graph);
    E.graph_exec = nullptr;
    E.graph = nullptr;
    E.graph_failed = true;
    if (E.opts.driver == 2) throw StatusError{PPCR_ERR_CUDA, std::string("CUDA graph WHILE driver unavailable: ") + cudaGetErrorString(e)};
    return false;
}

// Runs ticks until every pair of the engine is done.
static void run_to_completion(Engine& E)
{
    const int np = static_cast<int>(E.pairs.size());
    k_align_begin<<<ceil_div(np, 64), 64, 0, E.stream>>>(E.d_pairs.p, np, E.d_loop.p);
    CK(cudaGetLastError());
    E.times.total_launches += 1;
    const bool rec = E.opts.record_stage_times != 0;
    bool use_graph = (E.opts.driver == 2) || (E.opts.driver == 0 && !rec);
    Trace tr(E.stream);
    if (use_graph) use_graph = build_graph(E);
    tr.mark("graph build");
    if (use_graph) {
        CK(cudaGraphLaunch(E.graph_exec, E.stream));
        launch_final_transform(E);
        CK(cudaStreamSynchronize(E.stream));
        tr.mark("graph run");
    } else {
        long long guard = 0;
        for (;;) {
            for (int t = 0; t < E.opts.ticks_per_sync; ++t) launch_tick(E, false, rec);
            CK(cudaMemcpyAsync(E.h_active, E.d_loop.p, sizeof(int), cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            if (!E.h_active[0]) break;
            guard += E.opts.ticks_per_sync;
            if (guard > static_cast<long long>(E.max_ticks) + 64) throw StatusError{PPCR_ERR_CUDA, "tick guard exceeded"};
        }
        launch_final_transform(E);
        CK(cudaStreamSynchronize(E.stream));
        if (rec) collect_stage_times(E);
    }
}

// every entry point that works on an existing handle starts here
static void use_engine(Engine& E)
{
    CK(cudaSetDevice(E.device));
    g_alloc_stream = E.stream;
    g_launch_sink = &E.times.total_launches;
    E.h_active = pinned_flag();  // the CALLING thread's pinned word (the handle may have been created on another thread)
}

static PairState download_state(Engine& E, int p)
{
    PairState s;
    CK(cudaMemcpyAsync(&s, E.pairs[p].state.p, sizeof(s), cudaMemcpyDeviceToHost, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    return s;
}

static void check_state_error(const PairState& s)
{
    if (s.error == kErrRowOverflow)
        throw StatusError{PPCR_ERR_UNSUPPORTED, "a source point has 128 or more targets within the radius: neighbour sets beyond 128 "
                                                "(max_neighbours <= 0 or > 128) are not supported -- reduce the radius or filter the target"};
    if (s.error >= 100) throw StatusError{PPCR_ERR_TIMEOUT, "a peer rank did not arrive at the moment exchange"};
    if (s.error != 0) throw StatusError{PPCR_ERR_CUDA, "device tick cap reached before convergence"};
}

// ------------------------------------------------------------------------------------------------------------
// helpers shared by the stage-level entry points
// ------------------------------------------------------------------------------------------------------------

static void set_phase(Engine& E, int p, int phase)
{
    PairState s = download_state(E, p);
    s.phase = phase;
    s.search_cursor = 0;
    s.eval_ticket = 0;
    s.apply_dT = 0;
    CK(cudaMemcpyAsync(E.pairs[p].state.p, &s, sizeof(s), cudaMemcpyHostToDevice, E.stream));
    CK(cudaStreamSynchronize(E.stream));
}

// original (caller-order) index of every device row of the Morton-sorted source
static std::vector<int> source_order(Engine& E, int p)
{
    Pair& P = E.pairs[p];
    const int n = P.dev.n_src;
    std::vector<float4> h(static_cast<size_t>(std::max(n, 1)));
    if (n > 0) CK(cudaMemcpyAsync(h.data(), P.src.p, static_cast<size_t>(n) * sizeof(float4), cudaMemcpyDeviceToHost, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    std::vector<int> order(static_cast<size_t>(n));
    for (int j = 0; j < n; ++j) {
        int w;
        memcpy(&w, &h[j].w, 4);
        if (w < 0 || w >= n) throw StatusError{PPCR_ERR_CUDA, "corrupt source index tag"};
        order[j] = w;
    }
    return order;
}

// slot-major device planes -> row-major [n_src][max_nn] host arrays
static void download_association(Engine& E, int p, int32_t* idx, float* d2, int32_t* count, int64_t n_src, int max_nn)
{
    Pair& P = E.pairs[p];
    const PairDev& D = P.dev;
    if (n_src != D.n_src) throw StatusError{PPCR_ERR_INVALID, "n_src does not match the handle's (filtered) source size"};
    std::vector<int> h_idx(static_cast<size_t>(D.m) * D.n_pad), h_cnt(D.n_pad);
    std::vector<float4> h_tbl(static_cast<size_t>(std::max(D.n_tgt, 1)));
    std::vector<float> h_d2;
    CK(cudaMemcpyAsync(h_idx.data(), D.nbr_pos, h_idx.size() * sizeof(int), cudaMemcpyDeviceToHost, E.stream));
    if (D.n_tgt > 0)
        CK(cudaMemcpyAsync(h_tbl.data(), D.tgt_sorted, static_cast<size_t>(D.n_tgt) * sizeof(float4), cudaMemcpyDeviceToHost, E.stream));
    CK(cudaMemcpyAsync(h_cnt.data(), D.nbr_cnt, h_cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, E.stream));
    if (d2) {
        if (!D.nbr_d2) throw StatusError{PPCR_ERR_INVALID, "distances were not recorded"};
        h_d2.resize(h_idx.size());
        CK(cudaMemcpyAsync(h_d2.data(), D.nbr_d2, h_d2.size() * sizeof(float), cudaMemcpyDeviceToHost, E.stream));
    }
    CK(cudaStreamSynchronize(E.stream));
    const std::vector<int> order = source_order(E, p);
    auto original_index = [&](int pos) {  // the device stores positions in the Morton-sorted target
        int idx_out = -1;
        if (pos >= 0 && pos < D.n_tgt) memcpy(&idx_out, &h_tbl[static_cast<size_t>(pos)].w, 4);
        return idx_out;
    };
    // rows sit on the device in heap order: hand them out sorted by (d2, index) when distances were recorded
    // (FLANN's result order), by target index otherwise (the column order of the reference's CSR, :82-83)
    std::vector<std::pair<float, int>> row;
    for (int64_t j = 0; j < n_src; ++j) {  // device row j is the caller's point order[j]
        const int64_t i = order[j];
        const int c = std::min(h_cnt[j], D.m);
        count[i] = c;
        row.clear();
        for (int k = 0; k < c; ++k)
            row.push_back({d2 ? h_d2[static_cast<size_t>(k) * D.n_pad + j] : 0.f, original_index(h_idx[static_cast<size_t>(k) * D.n_pad + j])});
        std::sort(row.begin(), row.end());
        for (int k = 0; k < max_nn; ++k) {
            const bool have = k < c;
            idx[i * max_nn + k] = have ? row[k].second : -1;
            if (d2) d2[i * max_nn + k] = have ? row[k].first : 0.f;
        }
    }
}

// explicit association (row-major idx/count on the host) -> slot-major planes on the device
static void upload_association(Engine& E, int p, const float* tgt_xyzw, int64_t n_tgt, const int32_t* idx,
                               const int32_t* count, int max_nn)
{
    Pair& P = E.pairs[p];
    const PairDev& D = P.dev;
    const size_t plane = static_cast<size_t>(D.m) * D.n_pad;
    std::vector<int> hr(plane, 0);
    std::vector<int> hc(D.n_pad, 0);
    if (n_tgt != D.n_tgt) throw StatusError{PPCR_ERR_INVALID, "n_tgt does not match the handle's (filtered) target size"};
    std::vector<int> h_inv(static_cast<size_t>(std::max(D.n_tgt, 1)), 0);
    if (D.n_tgt > 0) {
        CK(cudaMemcpyAsync(h_inv.data(), D.inv_perm, static_cast<size_t>(D.n_tgt) * sizeof(int), cudaMemcpyDeviceToHost, E.stream));
        CK(cudaStreamSynchronize(E.stream));
    }
    (void)tgt_xyzw;
    int64_t K = 0;
    const std::vector<int> order = source_order(E, p);
    for (int64_t row = 0; row < D.n_src; ++row) {  // device row `row` is the caller's point order[row]
        const int64_t i = order[row];
        const int c = count[i];
        if (c < 0 || c > max_nn || c > D.m) throw StatusError{PPCR_ERR_INVALID, "association count out of range"};
        hc[row] = c;
        K += c;
        for (int k = 0; k < c; ++k) {
            const int j = idx[i * max_nn + k];
            if (j < 0 || j >= n_tgt) throw StatusError{PPCR_ERR_INVALID, "association index out of range"};
            hr[static_cast<size_t>(k) * D.n_pad + row] = h_inv[static_cast<size_t>(j)];
        }
    }
    CK(cudaMemcpyAsync(D.nbr_pos, hr.data(), plane * sizeof(int), cudaMemcpyHostToDevice, E.stream));
    CK(cudaMemcpyAsync(D.nbr_cnt, hc.data(), hc.size() * sizeof(int), cudaMemcpyHostToDevice, E.stream));
    PairState s = download_state(E, p);
    s.K = K;
    CK(cudaMemcpyAsync(P.state.p, &s, sizeof(s), cudaMemcpyHostToDevice, E.stream));
    CK(cudaStreamSynchronize(E.stream));
}

// the 24 moments, reduced exactly like the last block of k_evalctl does (interleaved chains, then chains in order)
static void reduce_partials_like_controller(Engine& E, int p, double* S)
{
    const PairDev& D = E.pairs[p].dev;
    std::vector<double> part(static_cast<size_t>(D.n_eval_blocks) * kNSum);
    CK(cudaMemcpyAsync(part.data(), D.partials, part.size() * sizeof(double), cudaMemcpyDeviceToHost, E.stream));
    CK(cudaStreamSynchronize(E.stream));
    for (int k = 0; k < kNSum; ++k) {
        double total = 0.0;
        for (int q = 0; q < kFoldChains; ++q) {
            double v = 0.0;
            for (int g = q; g < D.n_eval_groups; g += kFoldChains) {
                double gp = 0.0;  // level 1: the group's blocks in order
                for (int b = g * kFoldGroup; b < std::min(D.n_eval_blocks, (g + 1) * kFoldGroup); ++b)
                    gp += part[static_cast<size_t>(b) * kNSum + k];
                v += gp;
            }
            total += v;
        }
        S[k] = total;
    }
}

template <typename F>
static ppcr_status guarded(F&& f)
{
    try {
        f();
        return PPCR_OK;
    } catch (const CudaError& e) {
        return translate(e);
    } catch (const StatusError& e) {
        return fail(e.code, e.msg);
    } catch (const std::bad_alloc&) {
        return fail(PPCR_ERR_CUDA, "host allocation failed");
    }
}

// ------------------------------------------------------------------------------------------------------------
// extern "C"
// ------------------------------------------------------------------------------------------------------------

extern "C" {

const char* ppcr_last_error(void) { return g_last_error.c_str(); }
const char* ppcr_version(void) { return "ppcr-b200 0.1 (sm_100a)"; }

void ppcr_default_params(ppcr_params* p)
{
    memset(p, 0, sizeof(*p));
    p->max_neighbours = 20;
    p->dof = 5;
    p->radius = 1;
    p->n_iter = 1000;
    p->cost_drop_thresh = 0.01;
    p->n_cost_drop_it = 5;
    p->initial_rotation[0] = 1;
}

void ppcr_default_options(ppcr_options* o) { memset(o, 0, sizeof(*o)); }

ppcr_status ppcr_create_ex(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, const ppcr_params* params,
                           const ppcr_options* options, ppcr_handle** out)
{
    if (!params || !out) return fail(PPCR_ERR_INVALID, "null argument");
    *out = nullptr;
    ppcr_handle* h = nullptr;
    ppcr_status s = guarded([&] {
        h = new ppcr_handle();
        Engine& E = h->eng;
        engine_init(E, *params, options, true);
        E.pairs.resize(1);
        pair_setup(E, E.pairs[0], src, n_src, tgt, n_tgt, E.opts.input_on_device != 0);
        engine_commit(E);
    });
    if (s != PPCR_OK) {
        delete h;
        return s;
    }
    *out = h;
    return PPCR_OK;
}

ppcr_status ppcr_create(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, const ppcr_params* params,
                        ppcr_handle** out)
{
    return ppcr_create_ex(src, n_src, tgt, n_tgt, params, nullptr, out);
}

void ppcr_destroy(ppcr_handle* h) { delete h; }

ppcr_status ppcr_align(ppcr_handle* h)
{
    if (!h) return fail(PPCR_ERR_INVALID, "null handle");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        run_to_completion(E);
        check_state_error(download_state(E, 0));
    });
}

ppcr_status ppcr_has_converged(ppcr_handle* h, int32_t* out)
{
    if (!h || !out) return fail(PPCR_ERR_INVALID, "null argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        PairState s = download_state(E, 0);
        *out = has_converged(&s, &E.pairs[0].hcfg) ? 1 : 0;  // mutates the counter, like the reference
        CK(cudaMemcpyAsync(E.pairs[0].state.p, &s, sizeof(s), cudaMemcpyHostToDevice, E.stream));
        CK(cudaStreamSynchronize(E.stream));
    });
}

ppcr_status ppcr_history(ppcr_handle* h, double* T, int32_t* n_inout)
{
    if (!h || !n_inout) return fail(PPCR_ERR_INVALID, "null argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        PairState s = download_state(E, 0);
        const int n = std::min(s.current_iteration, E.pairs[0].dev.max_hist);
        const int take = std::min(n, *n_inout);
        if (T && take > 0) {
            CK(cudaMemcpyAsync(T, E.pairs[0].history.p, static_cast<size_t>(take) * 16 * sizeof(double), cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
        }
        *n_inout = n;
    });
}

ppcr_status ppcr_increment_history(ppcr_handle* h, double* T, int32_t* n_inout)
{
    if (!h || !n_inout) return fail(PPCR_ERR_INVALID, "null argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        PairState s = download_state(E, 0);
        const int max_hist = E.pairs[0].dev.max_hist;
        const int n = std::min(s.current_iteration, max_hist);
        const int take = std::min(n, *n_inout);
        if (T && take > 0) {
            CK(cudaMemcpyAsync(T, E.pairs[0].history.p + static_cast<size_t>(max_hist) * 16,
                               static_cast<size_t>(take) * 16 * sizeof(double), cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
        }
        *n_inout = n;
    });
}

ppcr_status ppcr_iteration_stats(ppcr_handle* h, ppcr_iter_stats* out, int32_t* n_inout)
{
    if (!h || !n_inout) return fail(PPCR_ERR_INVALID, "null argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        PairState s = download_state(E, 0);
        const int n = std::min(s.current_iteration, E.pairs[0].dev.max_hist);
        const int take = std::min(n, *n_inout);
        if (out && take > 0) {
            CK(cudaMemcpyAsync(out, E.pairs[0].stats.p, static_cast<size_t>(take) * sizeof(IterStats), cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
        }
        *n_inout = n;
    });
}

// clear_w: the cloud is the index-tagged, Morton-sorted source -- hand it back in the caller's order with w = 1
static void download_cloud(Engine& E, const float4* dev, int64_t n, float* out, int64_t* n_inout, bool clear_w)
{
    if (!n_inout) throw StatusError{PPCR_ERR_INVALID, "null argument"};
    if (*n_inout < n || !out) {
        *n_inout = n;
        if (out) throw StatusError{PPCR_ERR_SMALL_BUFFER, "output buffer too small"};
        return;
    }
    if (n > 0 && !clear_w) {
        CK(cudaMemcpyAsync(out, dev, static_cast<size_t>(n) * sizeof(float4), cudaMemcpyDeviceToHost, E.stream));
        CK(cudaStreamSynchronize(E.stream));
    } else if (n > 0) {
        std::vector<float4> h(static_cast<size_t>(n));
        CK(cudaMemcpyAsync(h.data(), dev, static_cast<size_t>(n) * sizeof(float4), cudaMemcpyDeviceToHost, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        for (int64_t j = 0; j < n; ++j) {
            int w;
            memcpy(&w, &h[j].w, 4);
            if (w < 0 || w >= n) throw StatusError{PPCR_ERR_CUDA, "corrupt source index tag"};
            out[4 * static_cast<int64_t>(w)] = h[j].x;
            out[4 * static_cast<int64_t>(w) + 1] = h[j].y;
            out[4 * static_cast<int64_t>(w) + 2] = h[j].z;
            out[4 * static_cast<int64_t>(w) + 3] = 1.0f;
        }
    }
    *n_inout = n;
}

ppcr_status ppcr_filtered_source(ppcr_handle* h, float* out, int64_t* n_inout)
{
    if (!h) return fail(PPCR_ERR_INVALID, "null handle");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        download_cloud(E, E.pairs[0].src.p, E.pairs[0].n_src, out, n_inout, true);
    });
}

ppcr_status ppcr_filtered_target(ppcr_handle* h, float* out, int64_t* n_inout)
{
    if (!h) return fail(PPCR_ERR_INVALID, "null handle");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        download_cloud(E, E.pairs[0].tgt_raw.p, E.pairs[0].n_tgt, out, n_inout, false);
    });
}

ppcr_status ppcr_association(ppcr_handle* h, int32_t* idx, int32_t* count, int64_t n_src, int32_t max_neighbours)
{
    if (!h || !idx || !count) return fail(PPCR_ERR_INVALID, "null argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        download_association(E, 0, idx, nullptr, count, n_src, max_neighbours);
    });
}

ppcr_status ppcr_get_stage_times(ppcr_handle* h, ppcr_stage_times* out)
{
    if (!h || !out) return fail(PPCR_ERR_INVALID, "null argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        *out = E.times;
        if (E.world > 1) {
            const PairState s = download_state(E, 0);
            const int khz = g_devices[E.device].clock_khz;
            out->exchanges = s.evals;
            out->exchange_wait_ms = khz > 0 ? static_cast<float>(static_cast<double>(s.exchange_cycles) / khz) : 0.f;
        }
        if (E.graph_ready) {  // the WHILE graph loops on the device: k_evalctl every tick, k_search once per outer iteration
            const PairState s = download_state(E, 0);
            out->ticks = s.ticks;
            out->total_launches = E.times.total_launches + s.ticks + (E.skip_search ? 0 : search_launches(E) * s.current_iteration);
        }
    });
}

ppcr_status ppcr_time_kernel(ppcr_handle* h, int32_t which, int32_t reps, int32_t flush_l2, float* avg_ms,
                             double* algorithmic_bytes)
{
    if (!h || !avg_ms || reps <= 0) return fail(PPCR_ERR_INVALID, "bad argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        Pair& P = E.pairs[0];
        const PairDev& D = P.dev;
        PairState saved = download_state(E, 0);
        PairState tmp = saved;
        const bool is_search = (which == 0 || which == 4);
        tmp.phase = is_search ? PH_SEARCH : PH_LM;
        tmp.search_cursor = 0;
        tmp.eval_ticket = 0;
        tmp.K = 0;
        tmp.apply_dT = (which == 2 || which == 4) ? 1 : 0;
        if (tmp.apply_dT)
            for (int k = 0; k < 16; ++k) tmp.dT[k] = (k % 5 == 0) ? 1.0 : 0.0;  // identity: the cloud is unchanged
        const size_t flush_n = (256ull << 20) / sizeof(float4);
        if (flush_l2) E.flush.reserve(flush_n);
        std::vector<cudaEvent_t> ev(2 * static_cast<size_t>(reps));
        for (auto& e : ev) CK(cudaEventCreate(&e));
        double K_sum = 0;
        // PPCR_PROFILE_KERNEL=<which>: bracket exactly these launches for `ncu --profile-from-start off`
        const char* prof_env = getenv("PPCR_PROFILE_KERNEL");
        const bool prof = prof_env && atoi(prof_env) == which;
        for (int r = 0; r < reps; ++r) {
            // every launch starts from the same state (the controller / the work cursor mutate it)
            CK(cudaMemcpyAsync(P.state.p, &tmp, sizeof(tmp), cudaMemcpyHostToDevice, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            if (flush_l2) k_fill<<<4 * std::max(g_sm_count, 1), 256, 0, E.stream>>>(E.flush.p, flush_n, static_cast<float>(r));
            if (prof) cudaProfilerStart();
            CK(cudaEventRecord(ev[2 * r], E.stream));
            switch (which) {
                case 0: case 4: launch_search(E); break;
                case 1: launch_evalctl(E, false); break;
                case 5: launch_evalctl(E, false, 2); break;  // probes: streaming part only / everything but the LM state machine
                case 6: launch_evalctl(E, false, 4); break;
                case 2: k_transform_final<<<dim3(E.max_tr_blocks, 1), 256, 0, E.stream>>>(E.d_pairs.p); break;
                default: build_target_tree(E, P, D.tree.leaf_cap); break;
            }
            CK(cudaEventRecord(ev[2 * r + 1], E.stream));
            if (prof) {
                CK(cudaStreamSynchronize(E.stream));
                cudaProfilerStop();
            }
            if (is_search) K_sum += static_cast<double>(download_state(E, 0).K);
        }
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(E.stream));
        double total = 0;
        for (int r = 0; r < reps; ++r) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, ev[2 * r], ev[2 * r + 1]));
            total += ms;
        }
        for (auto& e : ev) cudaEventDestroy(e);
        *avg_ms = static_cast<float>(total / reps);
        const double K = is_search ? K_sum / reps : static_cast<double>(saved.K);
        if (algorithmic_bytes) {
            const double ns = D.n_src, nt = D.n_tgt;
            switch (which) {
                // query float4 (+ its write-back when the cloud move is fused in), sorted target float4 +
                // index map, count and k-th distance, one 4-byte index per correspondence
                case 0: *algorithmic_bytes = 16.0 * ns + 20.0 * nt + 8.0 * ns + 4.0 * K; break;
                case 4: *algorithmic_bytes = 32.0 * ns + 20.0 * nt + 12.0 * ns + 4.0 * K; break;
                // source float4 + count; per correspondence a 4-byte index and the 16-byte target point it
                // names (the points are gathered from the L2-resident sorted target, so DRAM traffic is far below this)
                case 1: case 5: case 6: *algorithmic_bytes = 20.0 * ns + 20.0 * K; break;
                case 2: *algorithmic_bytes = 32.0 * ns; break;
                default: *algorithmic_bytes = 44.0 * nt; break;  // read 16, key+value 12, sorted write 16
            }
        }
        CK(cudaMemcpyAsync(P.state.p, &saved, sizeof(saved), cudaMemcpyHostToDevice, E.stream));
        CK(cudaStreamSynchronize(E.stream));
    });
}

// ---- stage-level entry points ------------------------------------------------------------------------------

ppcr_status ppcr_voxel_filter(const float* xyzw, int64_t n, double leaf, float* out_xyzw, int64_t* n_out)
{
    if ((n > 0 && (!xyzw || !out_xyzw)) || !n_out || n < 0 || !(leaf > 0)) return fail(PPCR_ERR_INVALID, "bad argument");
    return guarded([&] {
        ppcr_params prm;
        ppcr_default_params(&prm);
        Engine E;
        engine_init(E, prm, nullptr);
        Pair P;
        DevBuf<float4> in, out;
        upload_cloud(in, xyzw, n, false, E.stream);
        int64_t k = voxel_filter_device(in.p, n, leaf, out, P, E.stream);
        if (k < 0) {
            memcpy(out_xyzw, xyzw, static_cast<size_t>(n) * 16);
            *n_out = n;
        } else {
            if (k > 0) CK(cudaMemcpy(out_xyzw, out.p, static_cast<size_t>(k) * sizeof(float4), cudaMemcpyDeviceToHost));
            *n_out = k;
        }
        in.release();
        out.release();
        P.release();
    });
}

ppcr_status ppcr_time_voxel_filter(const float* xyzw, int64_t n, double leaf, const ppcr_options* options, int32_t reps,
                                   float* avg_ms, double* algorithmic_bytes, int64_t* n_out)
{
    if (!xyzw || n < 1 || !(leaf > 0) || reps < 1 || !avg_ms) return fail(PPCR_ERR_INVALID, "bad argument");
    return guarded([&] {
        ppcr_params prm;
        ppcr_default_params(&prm);
        ppcr_options opt;
        ppcr_default_options(&opt);
        if (options) opt = *options;
        Engine E;
        engine_init(E, prm, &opt);
        Pair P;
        DevBuf<float4> in, out;
        upload_cloud(in, xyzw, n, opt.input_on_device != 0, E.stream);
        int64_t k = voxel_filter_device(in.p, n, leaf, out, P, E.stream);  // warm-up (allocations)
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a));
        CK(cudaEventCreate(&b));
        CK(cudaEventRecord(a, E.stream));
        for (int r = 0; r < reps; ++r) k = voxel_filter_device(in.p, n, leaf, out, P, E.stream);
        CK(cudaEventRecord(b, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, a, b));
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        *avg_ms = ms / static_cast<float>(reps);
        if (n_out) *n_out = k;
        // every input point read once, every centroid written once (the sort's own traffic is not algorithmic)
        if (algorithmic_bytes) *algorithmic_bytes = 16.0 * static_cast<double>(n) + 16.0 * static_cast<double>(std::max<int64_t>(k, 0));
        in.release();
        out.release();
        P.release();
    });
}

ppcr_status ppcr_radius_search(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, double radius,
                               int32_t max_nn, int32_t leaf_capacity, int32_t* out_idx, float* out_d2, int32_t* out_count)
{
    if (!out_idx || !out_count) return fail(PPCR_ERR_INVALID, "null output");
    return guarded([&] {
        ppcr_params prm;
        ppcr_default_params(&prm);
        prm.radius = radius;
        prm.max_neighbours = max_nn;
        ppcr_options opt;
        ppcr_default_options(&opt);
        opt.leaf_capacity = leaf_capacity;
        Engine E;
        engine_init(E, prm, &opt);
        E.pairs.resize(1);
        E.pairs[0].want_d2 = true;
        pair_setup(E, E.pairs[0], src, n_src, tgt, n_tgt, false);
        engine_commit(E);
        set_phase(E, 0, PH_SEARCH);
        if (n_src > 0) launch_search(E);
        note_launches(1);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(E.stream));
        download_association(E, 0, out_idx, out_d2, out_count, n_src, max_nn);
    });
}

static void pose7_to_state(const double* pose, Pose* out) { pose_from_x(pose, out); }

ppcr_status ppcr_weights_normal_eq(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, const int32_t* idx,
                                   const int32_t* count, int32_t max_nn, double dof, int32_t dimension,
                                   const double* pose_w, const double* pose_e, int32_t fast_weights,
                                   double* weights, double* normal_eq)
{
    if (!idx || !count || !pose_w || !pose_e || !normal_eq) return fail(PPCR_ERR_INVALID, "null argument");
    return guarded([&] {
        ppcr_params prm;
        ppcr_default_params(&prm);
        prm.max_neighbours = max_nn;
        prm.dof = dof;
        prm.radius = 1.0;
        ppcr_options opt;
        ppcr_default_options(&opt);
        opt.exact_weights = fast_weights ? 0 : 1;
        Engine E;
        engine_init(E, prm, &opt);
        E.pairs.resize(1);
        pair_setup(E, E.pairs[0], src, n_src, tgt, n_tgt, false);
        E.pairs[0].dev.wcfg = make_weight_cfg(dof, dimension > 0 ? dimension : 3);
        engine_commit(E);
        upload_association(E, 0, tgt, n_tgt, idx, count, max_nn);
        PairState s = download_state(E, 0);
        s.phase = PH_LM;
        pose7_to_state(pose_e, &s.pose_e);
        pose7_to_state(pose_w, &s.pose_w);
        CK(cudaMemcpyAsync(E.pairs[0].state.p, &s, sizeof(s), cudaMemcpyHostToDevice, E.stream));
        // The weights come out of the evaluation kernel itself (PairDev::dump_w): the same staged loop, row statistics and
        // weight terms that produce the moments, not a separate test kernel.
        DevBuf<double> dw;
        const size_t plane = static_cast<size_t>(E.pairs[0].dev.m) * E.pairs[0].dev.n_pad;
        if (weights) {
            dw.reserve(plane);
            CK(cudaMemsetAsync(dw.p, 0, plane * sizeof(double), E.stream));
            E.pairs[0].dev.dump_w = dw.p;
            engine_commit(E);  // republish the PairDev
        }
        launch_evalctl(E, false);  // its controller tail runs on a scratch state; only the partial sums are used
        CK(cudaGetLastError());
        if (weights) {
            const PairDev& D = E.pairs[0].dev;
            std::vector<double> hw(plane);
            CK(cudaMemcpyAsync(hw.data(), dw.p, plane * sizeof(double), cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            const std::vector<int> order = source_order(E, 0);
            for (int64_t row = 0; row < n_src; ++row) {
                const int64_t i = order[row];
                for (int k = 0; k < max_nn; ++k)
                    weights[i * max_nn + k] = (k < count[i]) ? hw[static_cast<size_t>(k) * D.n_pad + row] : 0.0;
            }
            dw.release();
        }
        double S[kNSum];
        reduce_partials_like_controller(E, 0, S);
        Expanded ev;
        expand_moments(S, pose_e, &ev);  // the controller's expansion, same source
        int o = 0;
        for (int r = 0; r < kNP; ++r)
            for (int c = r; c < kNP; ++c) normal_eq[o++] = ev.H[r * kNP + c];
        for (int r = 0; r < kNP; ++r) normal_eq[o++] = ev.g[r];
        normal_eq[o] = ev.cost;
    });
}

ppcr_status ppcr_iteration_solve(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, const int32_t* idx,
                                 const int32_t* count, int32_t max_nn, const ppcr_params* params,
                                 const ppcr_options* options, double function_tolerance, double* out_pose, double* out_T,
                                 ppcr_iter_stats* stats)
{
    if (!idx || !count || !params) return fail(PPCR_ERR_INVALID, "null argument");
    return guarded([&] {
        ppcr_params prm = *params;
        prm.max_neighbours = max_nn;
        prm.n_iter = 1;  // exactly one ceres::Solve
        prm.source_filter_size = 0;
        prm.target_filter_size = 0;
        ppcr_options opt;
        ppcr_default_options(&opt);
        if (options) opt = *options;
        opt.function_tolerance = function_tolerance;
        opt.driver = 1;
        Engine E;
        engine_init(E, prm, &opt);
        E.pairs.resize(1);
        pair_setup(E, E.pairs[0], src, n_src, tgt, n_tgt, false);
        engine_commit(E);
        E.skip_search = true;
        // align_begin would reset K, so arm the state by hand: first loop test passed, association given
        PairState s = download_state(E, 0);
        lm_reset(&s, &E.pairs[0].hcfg);
        s.phase = PH_SEARCH;
        s.num_unuseful = 1;
        CK(cudaMemcpyAsync(E.pairs[0].state.p, &s, sizeof(s), cudaMemcpyHostToDevice, E.stream));
        CK(cudaStreamSynchronize(E.stream));
        upload_association(E, 0, tgt, n_tgt, idx, count, max_nn);
        long long guard = 0;
        for (;;) {
            for (int t = 0; t < E.opts.ticks_per_sync; ++t) launch_tick(E, false, false);
            CK(cudaMemcpyAsync(E.h_active, E.d_loop.p, sizeof(int), cudaMemcpyDeviceToHost, E.stream));
            CK(cudaStreamSynchronize(E.stream));
            if (!E.h_active[0]) break;
            if (++guard > 100000) throw StatusError{PPCR_ERR_CUDA, "inner solve did not terminate"};
        }
        PairState f = download_state(E, 0);
        check_state_error(f);
        if (out_pose) memcpy(out_pose, f.x, sizeof(double) * kNP);
        if (out_T) memcpy(out_T, f.dT, sizeof(double) * 16);
        if (stats) CK(cudaMemcpy(stats, E.pairs[0].stats.p, sizeof(IterStats), cudaMemcpyDeviceToHost));
    });
}

ppcr_status ppcr_transform_ex(float* xyzw, int64_t n, const double* T, const ppcr_options* options)
{
    if ((n > 0 && !xyzw) || !T || n < 0) return fail(PPCR_ERR_INVALID, "bad argument");
    return guarded([&] {
        const int device = options ? options->device : 0;
        const bool on_device = options && options->input_on_device;
        select_device(device);
        CK(cudaSetDevice(device));
        if (n == 0) return;
        cudaStream_t st = options && options->stream ? static_cast<cudaStream_t>(options->stream) : nullptr;
        bool own = false;
        if (!st) {
            CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            own = true;
        }
        g_alloc_stream = st;
        DevBuf<float4> d;
        DevBuf<double> dT;
        dT.reserve(16);
        CK(cudaMemcpyAsync(dT.p, T, 16 * sizeof(double), cudaMemcpyHostToDevice, st));
        float4* pts = reinterpret_cast<float4*>(xyzw);
        if (!on_device) {
            d.reserve(n);
            CK(cudaMemcpyAsync(d.p, xyzw, static_cast<size_t>(n) * 16, cudaMemcpyHostToDevice, st));
            pts = d.p;
        }
        k_transform_plain<<<std::min(ceil_div(n, 256), 8 * std::max(g_sm_count, 1)), 256, 0, st>>>(pts, static_cast<int>(n), dT.p);
        CK(cudaGetLastError());
        if (!on_device) CK(cudaMemcpyAsync(xyzw, d.p, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, st));
        d.release();
        dT.release();
        CK(cudaStreamSynchronize(st));  // T (and a host cloud) belong to the caller again when this returns
        if (own) cudaStreamDestroy(st);
        g_alloc_stream = nullptr;
    });
}

ppcr_status ppcr_transform(float* xyzw, int64_t n, const double* T) { return ppcr_transform_ex(xyzw, n, T, nullptr); }

ppcr_status ppcr_replay_metrics(ppcr_handle* h, float* cloud_xyzw, const float* gt_xyzw, int64_t n, int32_t first, int32_t count,
                                double* mse_gt, double* mse_prev)
{
    if (!h || (n > 0 && !cloud_xyzw) || n < 0 || first < 0 || count < 0) return fail(PPCR_ERR_INVALID, "bad argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        int32_t n_hist = 0;
        {
            const PairState s = download_state(E, 0);
            n_hist = s.current_iteration;
        }
        if (first + count > n_hist) throw StatusError{PPCR_ERR_INVALID, "replay range exceeds the outer iterations run"};
        if (count == 0 || n == 0) return;
        if (n > INT_MAX) throw StatusError{PPCR_ERR_INVALID, "cloud too large"};
        std::vector<double> inc(static_cast<size_t>(n_hist) * 16);
        int32_t cap = n_hist;
        if (ppcr_increment_history(h, inc.data(), &cap) != PPCR_OK) throw StatusError{PPCR_ERR_CUDA, ppcr_last_error()};
        use_engine(E);
        cudaStream_t st = E.stream;
        const int blocks = std::min(ceil_div(n, kReplayThreads), 8 * std::max(g_sm_count, 1));
        DevBuf<float4> d_cloud, d_gt;
        DevBuf<double> d_T, d_partial, d_out;
        d_cloud.reserve(n);
        if (gt_xyzw) d_gt.reserve(n);
        d_T.reserve(static_cast<size_t>(count) * 16);
        d_partial.reserve(static_cast<size_t>(blocks) * 2);
        d_out.reserve(static_cast<size_t>(count) * 2);
        CK(cudaMemcpyAsync(d_cloud.p, cloud_xyzw, static_cast<size_t>(n) * 16, cudaMemcpyHostToDevice, st));
        if (gt_xyzw) CK(cudaMemcpyAsync(d_gt.p, gt_xyzw, static_cast<size_t>(n) * 16, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d_T.p, inc.data() + static_cast<size_t>(first) * 16, static_cast<size_t>(count) * 16 * sizeof(double),
                           cudaMemcpyHostToDevice, st));
        for (int k = 0; k < count; ++k) {
            k_replay_step<<<blocks, kReplayThreads, 0, st>>>(d_cloud.p, gt_xyzw ? d_gt.p : nullptr, static_cast<int>(n),
                                                             d_T.p + static_cast<size_t>(k) * 16, d_partial.p);
            k_replay_fold<<<1, 32, 0, st>>>(d_partial.p, blocks, static_cast<int>(n), d_out.p + 2 * k);
        }
        CK(cudaGetLastError());
        note_launches(2 * count);
        std::vector<double> out(static_cast<size_t>(count) * 2);
        CK(cudaMemcpyAsync(out.data(), d_out.p, out.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(cloud_xyzw, d_cloud.p, static_cast<size_t>(n) * 16, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int k = 0; k < count; ++k) {
            if (mse_gt) mse_gt[k] = gt_xyzw ? out[2 * k] : 0.0;
            if (mse_prev) mse_prev[k] = out[2 * k + 1];
        }
        d_cloud.release(); d_gt.release(); d_T.release(); d_partial.release(); d_out.release();
    });
}

// ---- closest-point metrics ------------------------------------------------------------------------------------

ppcr_status ppcr_closest_point_metrics(const float* cloud1, int64_t n1, const float* cloud2, int64_t n2, double factor,
                                       const ppcr_options* options, ppcr_closest_metrics* out, float* out_d2)
{
    if (!cloud1 || !cloud2 || !out || n1 < 1 || n2 < 1 || !(factor > 0.0))
        return fail(PPCR_ERR_INVALID, "bad argument (both clouds need at least one point: the reference reads out of bounds otherwise)");
    return guarded([&] {
        ppcr_params prm;
        ppcr_default_params(&prm);
        prm.max_neighbours = 1;       // nearestKSearch(k = 1)
        prm.radius = 1.8e19;          // no radius: float(r * r) is the largest squared distance a float cloud can produce
        ppcr_options opt;
        ppcr_default_options(&opt);
        if (options) opt = *options;
        Engine E;
        engine_init(E, prm, &opt);
        E.pairs.resize(1);
        Pair& P = E.pairs[0];
        P.want_d2 = true;
        pair_setup(E, P, cloud1, n1, cloud2, n2, opt.input_on_device != 0);
        engine_commit(E);
        cudaStream_t st = E.stream;
        const int n = static_cast<int>(n1);
        {   // the queries must be finite too (the target was checked by the tree build)
            const Bbox bb = cloud_bbox(P.src.p, n, P, st);
            for (int k = 0; k < 3; ++k)
                if (!std::isfinite(bb.lo[k]) || !std::isfinite(bb.hi[k]))
                    throw StatusError{PPCR_ERR_INVALID, "cloud1 contains non-finite coordinates"};
        }
        set_phase(E, 0, PH_SEARCH);
        launch_search(E);
        note_launches(1);
        CK(cudaGetLastError());
        // m = 1: the distance plane is one float per query (device order = Morton order of cloud1)
        DevBuf<float> sorted;
        DevBuf<unsigned char> tmp;
        DevBuf<double> partial, d_out;
        DevBuf<ClosestPlan> plan;
        sorted.reserve(n);
        size_t tmp_bytes = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, P.nbr_d2.p, sorted.p, n, 0, 32, st));
        tmp.reserve(tmp_bytes);
        CK(cub::DeviceRadixSort::SortKeys(tmp.p, tmp_bytes, P.nbr_d2.p, sorted.p, n, 0, 32, st));
        const int blocks = std::min(ceil_div(n, kClosestThreads), 8 * std::max(g_sm_count, 1));
        partial.reserve(static_cast<size_t>(blocks) * 3);
        d_out.reserve(9);
        plan.reserve(1);
        k_closest_plan<<<1, 32, 0, st>>>(sorted.p, n, factor, plan.p);
        k_closest_reduce<<<blocks, kClosestThreads, 0, st>>>(sorted.p, n, plan.p, partial.p);
        k_closest_fold<<<1, 32, 0, st>>>(partial.p, blocks, n, plan.p, d_out.p);
        CK(cudaGetLastError());
        note_launches(3 + 6);
        double h[9];
        CK(cudaMemcpyAsync(h, d_out.p, sizeof(h), cudaMemcpyDeviceToHost, st));
        std::vector<float> h_d2;
        std::vector<int> order;
        if (out_d2) {
            h_d2.resize(static_cast<size_t>(n));
            CK(cudaMemcpyAsync(h_d2.data(), P.nbr_d2.p, static_cast<size_t>(n) * sizeof(float), cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        if (out_d2) {
            order = source_order(E, 0);
            for (int j = 0; j < n; ++j) out_d2[order[static_cast<size_t>(j)]] = h_d2[static_cast<size_t>(j)];
        }
        out->average_closest_distance = h[0];
        out->sum_squared_error = h[1];
        out->robust_sum_squared_error = h[2];
        out->robust_sum_squared_error_factor = h[3];
        out->robust_averaged_sum_squared_error = h[4];
        out->median_closest_distance = h[5];
        out->robust_median_closest_distance = h[6];
        out->n_filtered = static_cast<int64_t>(h[7]);
        out->n_filtered_factor = static_cast<int64_t>(h[8]);
        sorted.release(); tmp.release(); partial.release(); d_out.release(); plan.release();
    });
}

// ---- batch ---------------------------------------------------------------------------------------------------

ppcr_status ppcr_align_batch_devices(const ppcr_pair* pairs, int32_t n_pairs, const ppcr_params* params,
                                     const ppcr_options* options, const int32_t* device_ids, int32_t n_dev, int32_t slots,
                                     double* out_T, int32_t* out_n_outer, int64_t* out_corr)
{
    if (!pairs || n_pairs < 0 || !params || !out_T || !device_ids || n_dev < 1) return fail(PPCR_ERR_INVALID, "bad argument");
    if (n_dev > 1 && options && options->input_on_device)
        return fail(PPCR_ERR_INVALID, "device-resident pair buffers belong to one device: pass host buffers to a multi-device batch");
    return guarded([&] {
        if (n_pairs == 0) return;
        // `slots` lanes, each a host thread with its own engine and stream, pull pairs from a shared counter: the set-up
        // of one pair (upload, sorts, tree build) overlaps the iterations of the others, a 120k-point pair does not fill
        // 148 SMs on its own, and a lane that draws a pair needing few outer iterations moves on at once.
        // With several devices every device gets `slots` lanes and all lanes draw from the same counter: the pairs are dealt
        // dynamically, no device waits for another's block of the batch, and there is no exchange between them.
        if (slots <= 0) slots = 6;
        const int lanes = static_cast<int>(std::min<long long>(static_cast<long long>(slots) * n_dev, n_pairs));
        ppcr_options lane_opt;
        if (options) lane_opt = *options; else ppcr_default_options(&lane_opt);
        if (n_dev > 1) lane_opt.stream = nullptr;  // a stream belongs to one device
        if (lane_opt.stream && lanes > 1) {
            // device-resident inputs were produced on the caller's stream: wait for them, then use one stream per lane
            CK(cudaSetDevice(lane_opt.device));
            CK(cudaStreamSynchronize(static_cast<cudaStream_t>(lane_opt.stream)));
            lane_opt.stream = nullptr;
        }
        std::atomic<int> next{0};
        std::mutex err_mutex;
        StatusError first_error{PPCR_OK, ""};
        auto lane = [&](int lane_index) {
            try {
                Engine E;
                ppcr_options my_opt = lane_opt;
                my_opt.device = device_ids[lane_index % n_dev];  // lanes interleave over the devices
                engine_init(E, *params, &my_opt, true);
                E.batch_lane = lanes > 1;
                E.pairs.resize(1);
                const bool on_dev = E.opts.input_on_device != 0;
                for (;;) {
                    const int i = next.fetch_add(1);
                    if (i >= n_pairs) break;
                    {
                        std::lock_guard<std::mutex> lock(err_mutex);
                        if (first_error.code != PPCR_OK) break;
                    }
                    const ppcr_pair& pr = pairs[i];
                    pair_setup(E, E.pairs[0], pr.src_xyzw, pr.n_src, pr.tgt_xyzw, pr.n_tgt, on_dev);
                    engine_commit(E);
                    run_to_completion(E);
                    PairState f = download_state(E, 0);
                    check_state_error(f);
                    memcpy(out_T + static_cast<size_t>(i) * 16, f.T_total, sizeof(double) * 16);
                    if (out_n_outer) out_n_outer[i] = f.current_iteration;
                    if (out_corr) out_corr[i] = f.K_total;
                }
            } catch (const StatusError& e) {
                std::lock_guard<std::mutex> lock(err_mutex);
                if (first_error.code == PPCR_OK) first_error = e;
            } catch (const CudaError& e) {  // CK() throws this plain struct: it must not leave the thread
                const ppcr_status code = translate(e);  // (message lands in this lane thread's g_last_error)
                std::lock_guard<std::mutex> lock(err_mutex);
                if (first_error.code == PPCR_OK) first_error = StatusError{code, g_last_error};
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lock(err_mutex);
                if (first_error.code == PPCR_OK) first_error = StatusError{PPCR_ERR_CUDA, e.what()};
            } catch (...) {
                std::lock_guard<std::mutex> lock(err_mutex);
                if (first_error.code == PPCR_OK) first_error = StatusError{PPCR_ERR_CUDA, "unknown failure in a batch lane"};
            }
        };
        if (lanes == 1) {
            lane(0);
        } else {
            std::vector<std::thread> threads;
            for (int t = 0; t < lanes; ++t) threads.emplace_back(lane, t);
            for (auto& t : threads) t.join();
        }
        if (first_error.code != PPCR_OK) throw first_error;
    });
}

ppcr_status ppcr_align_batch(const ppcr_pair* pairs, int32_t n_pairs, const ppcr_params* params,
                             const ppcr_options* options, int32_t slots, double* out_T, int32_t* out_n_outer,
                             int64_t* out_corr)
{
    const int32_t device = options ? options->device : 0;
    return ppcr_align_batch_devices(pairs, n_pairs, params, options, &device, 1, slots, out_T, out_n_outer, out_corr);
}

// ---- one pair over several GPUs of one process ------------------------------------------------------------------

namespace {
struct HostBarrier {  // reusable barrier for the rank threads of ppcr_align_sharded
    std::mutex mu;
    std::condition_variable cv;
    int count = 0, generation = 0, parties;
    explicit HostBarrier(int n) : parties(n) {}
    void wait()
    {
        std::unique_lock<std::mutex> lock(mu);
        const int gen = generation;
        if (++count == parties) {
            count = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lock, [&] { return gen != generation; });
        }
    }
};
}  // namespace

ppcr_status ppcr_align_sharded(const float* src, int64_t n_src, const float* tgt, int64_t n_tgt, const ppcr_params* params,
                               const ppcr_options* options, const int32_t* device_ids, int32_t n_dev, double* out_T,
                               int32_t* n_inout, int64_t* out_corr)
{
    if (!params || !device_ids || n_dev < 1 || n_dev > 8 || !n_inout || n_src < 0 || n_tgt < 0 || (n_src > 0 && !src) || (n_tgt > 0 && !tgt))
        return fail(PPCR_ERR_INVALID, "bad argument");
    if (options && options->input_on_device) return fail(PPCR_ERR_INVALID, "ppcr_align_sharded takes host buffers");
    if (params->source_filter_size > 0)
        return fail(PPCR_ERR_UNSUPPORTED, "a source voxel filter would act on every share separately: filter the source first (ppcr_voxel_filter)");
    return guarded([&] {
        constexpr int64_t kRun = 8192;  // points per run of the block-cyclic deal
        std::vector<uint8_t> tokens(static_cast<size_t>(n_dev) * PPCR_SHARD_TOKEN_BYTES);
        HostBarrier barrier(n_dev);
        std::mutex err_mutex;
        StatusError first_error{PPCR_OK, ""};
        std::atomic<int> failed{0};
        const int cap = *n_inout;
        auto rank_main = [&](int r) {
            ppcr_handle* h = nullptr;
            auto note = [&](ppcr_status st) {
                if (st == PPCR_OK) return true;
                std::lock_guard<std::mutex> lock(err_mutex);
                if (first_error.code == PPCR_OK) first_error = StatusError{st, g_last_error};
                failed.store(1);
                return false;
            };
            // Nothing may leave this thread as an exception (std::terminate), and both barriers must be reached whatever happens:
            // every stage runs only while nobody has failed, inside its own try block.
            auto stage = [&](auto&& body) {
                if (failed.load()) return;
                try {
                    body();
                } catch (const std::exception& e) {
                    std::lock_guard<std::mutex> lock(err_mutex);
                    if (first_error.code == PPCR_OK) first_error = StatusError{PPCR_ERR_INVALID, e.what()};
                    failed.store(1);
                } catch (...) {
                    std::lock_guard<std::mutex> lock(err_mutex);
                    if (first_error.code == PPCR_OK) first_error = StatusError{PPCR_ERR_CUDA, "unknown failure in a rank of ppcr_align_sharded"};
                    failed.store(1);
                }
            };
            stage([&] {
                // this rank's share of the source: runs r, r + n_dev, r + 2 n_dev, ... of kRun consecutive points
                std::vector<float> share;
                const int64_t n_runs = (n_src + kRun - 1) / kRun;
                for (int64_t b = r; b < n_runs; b += n_dev) {
                    const int64_t lo = b * kRun, hi = std::min(n_src, lo + kRun);
                    share.insert(share.end(), src + 4 * lo, src + 4 * hi);
                }
                ppcr_options opt;
                if (options) opt = *options; else ppcr_default_options(&opt);
                opt.device = device_ids[r];
                opt.stream = nullptr;
                if (note(ppcr_create_ex(share.empty() ? nullptr : share.data(), static_cast<int64_t>(share.size() / 4), tgt, n_tgt, params, &opt, &h)) &&
                    n_dev > 1)
                    note(ppcr_shard_export(h, r, n_dev, tokens.data() + static_cast<size_t>(r) * PPCR_SHARD_TOKEN_BYTES));
            });
            barrier.wait();  // every token is written (or a rank has failed)
            stage([&] {
                if (n_dev > 1) note(ppcr_shard_connect(h, tokens.data()));
            });
            barrier.wait();  // every rank is connected: nobody starts writing into a mailbox whose owner is not ready
            stage([&] { note(ppcr_align(h)); });
            stage([&] {
                if (r != 0) return;
                int32_t n = cap;
                if (note(ppcr_history(h, out_T, &n))) *n_inout = n;
                if (out_corr) {
                    use_engine(h->eng);
                    *out_corr = download_state(h->eng, 0).K_total;
                }
            });
            ppcr_destroy(h);
        };
        if (n_dev == 1) {
            rank_main(0);
        } else {
            std::vector<std::thread> threads;
            for (int r = 0; r < n_dev; ++r) threads.emplace_back(rank_main, r);
            for (auto& t : threads) t.join();
        }
        if (first_error.code != PPCR_OK) throw first_error;
    });
}

// ---- page-locked host memory ----------------------------------------------------------------------------------

static std::mutex g_host_mutex;
static std::set<void*> g_host_blocks;

void* ppcr_host_alloc(size_t bytes)
{
    if (bytes == 0) return nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return nullptr;
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    std::lock_guard<std::mutex> lock(g_host_mutex);
    g_host_blocks.insert(p);
    return p;
}

int32_t ppcr_host_free(void* p)
{
    if (!p) return 0;
    {
        std::lock_guard<std::mutex> lock(g_host_mutex);
        auto it = g_host_blocks.find(p);
        if (it == g_host_blocks.end()) return 0;
        g_host_blocks.erase(it);
    }
    cudaFreeHost(p);
    cudaGetLastError();
    return 1;
}

// ---- sharded pair -------------------------------------------------------------------------------------------

ppcr_status ppcr_shard_export(ppcr_handle* h, int32_t rank, int32_t world, uint8_t* token_out)
{
    if (!h || !token_out || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(PPCR_ERR_INVALID, "bad argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        E.rank = rank;
        E.world = world;
        ShardToken tok{};
        {
            std::lock_guard<std::mutex> lock(g_shard.mu);
            DeviceMailbox& M = g_shard.box[E.device];
            if (!M.p) {
                // a dedicated allocation (IPC handles cover whole cudaMalloc blocks), made once per device and process:
                // every later sharded handle on this device re-uses it, and the peers keep their mapping of it
                CK(cudaMalloc(&M.p, kMailboxDoubles * sizeof(double)));
                CK(cudaMemset(M.p, 0, kMailboxDoubles * sizeof(double)));
                CK(cudaDeviceSynchronize());
                CK(cudaIpcGetMemHandle(&M.ipc, M.p));
            }
            tok.ipc = M.ipc;
            tok.pid = static_cast<uint64_t>(getpid());
            tok.ptr = reinterpret_cast<uint64_t>(M.p);
            tok.device = E.device;
            tok.epoch = g_shard.epoch + 1;  // proposal; the group takes the largest
        }
        memset(token_out, 0, PPCR_SHARD_TOKEN_BYTES);
        memcpy(token_out, &tok, sizeof(tok));
    });
}

ppcr_status ppcr_shard_connect(ppcr_handle* h, const uint8_t* tokens)
{
    if (!h || !tokens) return fail(PPCR_ERR_INVALID, "bad argument");
    return guarded([&] {
        Engine& E = h->eng;
        use_engine(E);
        Pair& P = E.pairs[0];
        std::lock_guard<std::mutex> lock(g_shard.mu);
        DeviceMailbox& M = g_shard.box[E.device];
        if (!M.p) throw StatusError{PPCR_ERR_INVALID, "call ppcr_shard_export first"};
        uint32_t epoch = 0;
        for (int r = 0; r < 8; ++r) P.dev.peer_mailbox[r] = nullptr;
        for (int r = 0; r < E.world; ++r) {
            ShardToken tok;
            memcpy(&tok, tokens + static_cast<size_t>(r) * PPCR_SHARD_TOKEN_BYTES, sizeof(tok));
            epoch = std::max(epoch, tok.epoch);
            if (r == E.rank) {
                P.dev.peer_mailbox[r] = M.p;
            } else if (tok.pid == static_cast<uint64_t>(getpid())) {
                // a rank of this very process (one host thread per device): plain peer access, no IPC
                if (tok.device != E.device) {
                    int can = 0;
                    CK(cudaDeviceCanAccessPeer(&can, E.device, tok.device));
                    if (!can) throw StatusError{PPCR_ERR_UNSUPPORTED, "the devices of a sharded pair need peer access to each other"};
                    const cudaError_t e = cudaDeviceEnablePeerAccess(tok.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CK(e);
                    cudaGetLastError();
                }
                P.dev.peer_mailbox[r] = reinterpret_cast<double*>(tok.ptr);
            } else {
                // another process: map its mailbox once, keep the mapping for the life of this process
                const std::string key(reinterpret_cast<const char*>(&tok.ipc), sizeof(tok.ipc));
                auto it = g_shard.opened.find(key);
                if (it == g_shard.opened.end()) {
                    void* ptr = nullptr;
                    CK(cudaIpcOpenMemHandle(&ptr, tok.ipc, cudaIpcMemLazyEnablePeerAccess));
                    it = g_shard.opened.emplace(key, ptr).first;
                }
                P.dev.peer_mailbox[r] = static_cast<double*>(it->second);
            }
        }
        // Stamps must be new to every mailbox of the group although the mailboxes outlive the handles: the group adopts the
        // largest proposal (every rank sees the same tokens, so the same value) and every rank's counter moves up to it.
        g_shard.epoch = std::max(g_shard.epoch, epoch);
        P.dev.mailbox = M.p;
        P.dev.mail_base = static_cast<double>(epoch) * kMailEpochStride;
        P.dev.rank = E.rank;
        P.dev.world = E.world;
        engine_commit(E);
    });
}

}  // extern "C"
