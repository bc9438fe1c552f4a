// b2sv.cu -- C ABI (include/b2sv.h) of the B200 statevector engine: handle management, plan
// execution and kernel launches.  Kernels: kernels.cuh.  Fusion planner: planner.hpp.
//
// Replaces, for one statevector, what tq.simulate(..., backend="qulacs") does behind
// Circuit.simulate (/root/reference/src/unitaria/circuit.py:62-70) plus the Subspace projection and
// norm of Node.simulate / simulate_norm (/root/reference/src/unitaria/nodes/node.py:363-403).
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <unistd.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/b2sv.h"
#include "kernels.cuh"
#include "planner.hpp"
#include "tileops.hpp"
#include "perm_jit.hpp"

// NVTX ranges (tracing): every C-ABI entry point that launches work and every sweep of a plan is a
// named range, so an Nsight timeline shows "b2sv_apply > plan > tile / perm / low6 ..." next to the kernels.  nvtx3 is
// header-only and resolves to a no-op call when no tool is attached.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? B2SV_ENOMEM : B2SV_ECUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                               \
    } while (0)

enum { CLS_GATE1 = 0, CLS_TILE = 1, CLS_PERM = 2, CLS_SCALE = 3, CLS_PROJECT = 4, CLS_INIT = 5, CLS_SWAP = 6, CLS_PERM_JIT = 7, CLS_LOW6 = 8, CLS_DIST = 9, CLS_N = 12 };

struct TimedLaunch {
    cudaEvent_t a, b;
    int cls;
};

}  // namespace

struct b2sv {
    int n = 0;
    int dtype = B2SV_C128;
    int device = 0;
    int amp_bytes = 16;
    uint64_t count = 0;
    cudaStream_t stream = nullptr;
    void* amps = nullptr;
    void* scratch = nullptr;       // second state-sized buffer for out-of-place permutation sweeps
    bool scratch_failed = false;
    bool pending_basis = false;    // |basis_index> has been requested but not written yet (lazy init)
    uint64_t basis_index = 0;
    // live window: the state is zero outside {i : (i & win_mask) == win_value} and only that part of the
    // buffer has been written (a basis state after one sweep that ran on the single tile owning the index,
    // when a permutation sweep follows: that sweep then reads one tile instead of a whole register of zeros)
    bool window = false;
    uint64_t win_mask = 0, win_value = 0;
    // known-zero state: every amplitude is 0 and the buffer has not been written (a shard that does not own the
    // basis state, b2sv_set_zero): every gate maps it to itself, so nothing runs until a distributed sweep brings
    // amplitudes in or a reader needs the zeros
    bool zero = false;
    // distributed execution (b2sv_ipc_* / b2sv_dist_*): the two state buffers and the hand-shake events of every
    // rank, mapped into this process
    int dist_rank = -1, dist_world = 0;
    void* own_buf[2] = {nullptr, nullptr};
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    struct Peer {
        void* buf[2] = {nullptr, nullptr};
        cudaEvent_t ready = nullptr, done = nullptr;
        bool ipc = false;  // opened through CUDA IPC (another process): close on detach
    };
    std::vector<Peer> peers;
    void* d_dist_gates = nullptr;
    size_t d_dist_gates_cap = 0;
    // pipelined results (b2sv_project_gather_async): two device result buffers drained by a copy stream while
    // the next input's sweeps run on the main stream
    void* d_col[2] = {nullptr, nullptr};
    size_t d_col_cap[2] = {0, 0};
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_col_kernel[2] = {nullptr, nullptr}, ev_col_copy[2] = {nullptr, nullptr};
    bool col_pending[2] = {false, false};
    std::vector<unsigned char> plan_gate_copy;  // the gate list the cached plan was made from (compared on a hash hit)
    bool timing_on = false;        // last b2sv_apply asked for per-launch timing: time init/project too
    double sre = 1.0, sim = 0.0;   // pending global scalar (folded uncontrolled global phases)
    void* d_ops = nullptr;         // device copy of the tile ops of the current apply
    size_t d_ops_cap = 0;
    void* d_tables = nullptr;      // device copy of the multiplexor tables of the current apply
    size_t d_tables_cap = 0;
    void* d_gather = nullptr;      // device-side result buffer of project_gather / enumerate_basis
    size_t d_gather_cap = 0;
    b2sv_subinstr* d_sub = nullptr; // device copy of the current subspace program
    int d_sub_cap = 0;
    double* d_partial = nullptr;   // projection partial sums (+1 slot for the result)
    int n_partial = 0;
    int sm_count = 148;
    // TMA tensor maps of the (up to two) state buffers: rows of 128 bytes, SWIZZLE_128B, per chunk size
    struct MapEntry {
        void* ptr;
        int rows_per_box;
        CUtensorMap map;
    };
    std::vector<MapEntry> maps;
    // the plan of the last b2sv_apply (verify-style callers apply one circuit to many inputs)
    bool plan_valid = false;
    uint64_t plan_key = 0, plan_key2 = 0;
    size_t plan_ngates = 0;
    b2svk::Plan plan;
    std::vector<size_t> plan_seg_first;
    std::vector<int> plan_seg_nops;
    std::vector<std::vector<b2svk::SlotFill>> plan_seg_slots;
    b2sv_stats_t stats{};
    std::vector<TimedLaunch> timed;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> apply_events;
    std::vector<cudaEvent_t> event_pool;
    cudaEvent_t marks[8] = {};
};

namespace {

using namespace b2svk;

// The permutation program lives in __constant__ memory, which is per device, not per handle.  Users on different streams are ordered through one event per device.
std::mutex g_const_mutex;
cudaEvent_t g_const_event[64] = {};

struct ConstGuard {
    b2sv* sv;
    explicit ConstGuard(b2sv* s) : sv(s) {
        g_const_mutex.lock();
        cudaEvent_t& ev = g_const_event[sv->device & 63];
        if (!ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        else cudaStreamWaitEvent(sv->stream, ev, 0);
    }
    ~ConstGuard() {
        cudaEventRecord(g_const_event[sv->device & 63], sv->stream);
        g_const_mutex.unlock();
    }
};

cudaEvent_t get_event(b2sv* sv) {
    if (!sv->event_pool.empty()) {
        cudaEvent_t e = sv->event_pool.back();
        sv->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// Resolve recorded event pairs into the stats (synchronises the stream).
int drain_events(b2sv* sv) {
    CU(cudaStreamSynchronize(sv->stream));
    // B2SV_DUMP_TIMES=<file>: append "class ms" of every timed launch, in launch order (profiling aid)
    static const char* dump_path = std::getenv("B2SV_DUMP_TIMES");
    FILE* dump = (dump_path && !sv->timed.empty()) ? std::fopen(dump_path, "a") : nullptr;
    for (auto& t : sv->timed) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, t.a, t.b));
        if (dump) std::fprintf(dump, "%d %.6f\n", t.cls, ms);
        sv->stats.ms_by_class[t.cls] += ms;
        sv->event_pool.push_back(t.a);
        sv->event_pool.push_back(t.b);
    }
    if (dump) {
        std::fprintf(dump, "-1 0\n");
        std::fclose(dump);
    }
    sv->timed.clear();
    for (auto& p : sv->apply_events) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, p.first, p.second));
        sv->stats.apply_ms += ms;
        sv->event_pool.push_back(p.first);
        sv->event_pool.push_back(p.second);
    }
    sv->apply_events.clear();
    return B2SV_OK;
}

struct LaunchTimer {
    b2sv* sv;
    bool on;
    TimedLaunch t{};
    LaunchTimer(b2sv* s, bool timing, int cls, double alg_bytes) : sv(s), on(timing) {
        sv->stats.launches++;
        sv->stats.launches_by_class[cls]++;
        sv->stats.alg_bytes_by_class[cls] += alg_bytes;
        if (on) {
            t.cls = cls;
            t.a = get_event(sv);
            t.b = get_event(sv);
            cudaEventRecord(t.a, sv->stream);
        }
    }
    ~LaunchTimer() {
        if (on) {
            cudaEventRecord(t.b, sv->stream);
            sv->timed.push_back(t);
        }
    }
};

int grid_for(const b2sv* sv, uint64_t work_items, int per_sm = 8) {
    uint64_t blocks = (work_items + kThreads - 1) / kThreads;
    const uint64_t cap = (uint64_t)sv->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename F>
int dispatch_dtype(const b2sv* sv, F&& f) {
    if (sv->dtype == B2SV_C128) return f((double2*)nullptr);
    return f((float2*)nullptr);
}

int ensure_gather(b2sv* sv, size_t bytes) {
    if (bytes <= sv->d_gather_cap) return B2SV_OK;
    CU(cudaStreamSynchronize(sv->stream));
    cudaFree(sv->d_gather);
    sv->d_gather = nullptr;
    sv->d_gather_cap = 0;
    CU(cudaMalloc(&sv->d_gather, bytes));
    sv->d_gather_cap = bytes;
    return B2SV_OK;
}

int ensure_scratch(b2sv* sv) {
    if (sv->scratch) return B2SV_OK;
    if (sv->scratch_failed) return B2SV_ENOMEM;
    cudaError_t e = cudaMalloc(&sv->scratch, sv->count * sv->amp_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        sv->scratch = nullptr;
        sv->scratch_failed = true;
        return B2SV_ENOMEM;
    }
    return B2SV_OK;
}

int flush_scalar(b2sv* sv, bool timing) {
    if (sv->sre == 1.0 && sv->sim == 0.0) return B2SV_OK;
    const double2 s = make_double2(sv->sre, sv->sim);
    int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, timing, CLS_SCALE, 2.0 * sv->amp_bytes * (double)sv->count);
        k_scale<A><<<grid_for(sv, sv->count), kThreads, 0, sv->stream>>>((A*)sv->amps, sv->count, s);
        return B2SV_OK;
    });
    CU(cudaGetLastError());
    sv->sre = 1.0;
    sv->sim = 0.0;
    return rc;
}

// Zero `bytes` of device memory on the handle's stream with the library's own fill kernel (buffers that are not a
// multiple of 32 bytes -- registers of one qubit -- and B2SV_MEMSET=1, the A/B baseline, take cudaMemsetAsync).
int fill_zero(b2sv* sv, void* ptr, size_t bytes) {
    static const bool use_memset = std::getenv("B2SV_MEMSET") != nullptr;
    if (use_memset || (bytes & 31) != 0 || (reinterpret_cast<uintptr_t>(ptr) & 31) != 0) {
        CU(cudaMemsetAsync(ptr, 0, bytes, sv->stream));
        return B2SV_OK;
    }
    const uint64_t n32 = bytes / 32;
    const uint64_t blocks = (n32 + 1023) / 1024;  // <= 2^21 for a 64 GiB buffer
    k_fill_zero<<<(unsigned)blocks, 256, 0, sv->stream>>>(reinterpret_cast<unsigned char*>(ptr), n32);
    sv->stats.launches++;  // a kernel of this library like any other (its time is booked by the caller's LaunchTimer)
    CU(cudaGetLastError());
    return B2SV_OK;
}

// Write the lazily requested basis state out (every reader of the amplitudes calls this first).
int materialise(b2sv* sv) {
    if (sv->zero) {
        sv->zero = false;
        sv->stats.launches_by_class[CLS_INIT]++;
        return fill_zero(sv, sv->amps, sv->count * sv->amp_bytes);
    }
    if (sv->window) {  // write the implied zeros
        sv->window = false;
        const uint64_t mask = sv->win_mask, value = sv->win_value;
        int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
            using A = std::remove_pointer_t<decltype(tag)>;
            LaunchTimer lt(sv, sv->timing_on, CLS_INIT, (double)sv->amp_bytes * (double)sv->count);
            k_window_fill<A><<<grid_for(sv, sv->count), kThreads, 0, sv->stream>>>((A*)sv->amps, sv->count, mask, value);
            return B2SV_OK;
        });
        CU(cudaGetLastError());
        if (rc) return rc;
    }
    if (!sv->pending_basis) return B2SV_OK;
    sv->pending_basis = false;
    const uint64_t index = sv->basis_index;
    int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, sv->timing_on, CLS_INIT, (double)sv->amp_bytes * (double)sv->count);
        k_set_basis<A><<<grid_for(sv, sv->count), kThreads, 0, sv->stream>>>((A*)sv->amps, sv->count, index);
        return B2SV_OK;
    });
    CU(cudaGetLastError());
    return rc;
}

// ---- launch helpers ---------------------------------------------------------------------------

int launch_gate1(b2sv* sv, const PGate& g, bool timing) {
    Gate1Args a{};
    uint64_t fixed = g.ctrl;
    a.op = g.op;
    a.ctrl = g.ctrl;
    a.bit0 = a.bit1 = 0;
    double touched = 2.0;
    switch (g.op) {
        case B2SV_OP_PHASE:
            touched = 1.0;
            break;
        case B2SV_OP_SWAP:
            a.bit0 = 1ull << g.t0;
            a.bit1 = 1ull << g.t1;
            fixed |= a.bit0 | a.bit1;
            break;
        default:
            a.bit0 = 1ull << g.t0;
            fixed |= a.bit0;
            break;
    }
    if (popcount64(fixed) > kMaxFixed) return fail(B2SV_EINVAL, "gate touches more than %d qubits", kMaxFixed);
    fill_expand(a.ex, fixed);
    a.count = 1ull << (sv->n - a.ex.m);
    a.total = sv->count;
    std::memcpy(a.m, g.m, sizeof(a.m));
    return dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, timing, CLS_GATE1, 2.0 * sv->amp_bytes * touched * (double)a.count);
        k_gate1<A><<<grid_for(sv, a.count), kThreads, 0, sv->stream>>>((A*)sv->amps, a);
        return B2SV_OK;
    });
}

template <typename A>
int set_tile_smem(int device, size_t bytes) {
    static size_t configured_dev[64] = {};  // the attribute is per device
    size_t& configured = configured_dev[device & 63];
    if (bytes > 48 * 1024 && configured < bytes) {
        CU(cudaFuncSetAttribute(k_tile<A, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        CU(cudaFuncSetAttribute(k_tile<A, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        CU(cudaFuncSetAttribute(k_tile<A, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        CU(cudaFuncSetAttribute(k_tile<A, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
    return B2SV_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// Tensor map viewing the current state buffer as rows of 128 bytes with the hardware 128-byte swizzle.
const CUtensorMap* state_tensor_map(b2sv* sv, int rows_per_box) {
    for (auto& e : sv->maps)
        if (e.ptr == sv->amps && e.rows_per_box == rows_per_box) return &e.map;
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return nullptr;
    b2sv::MapEntry e{};
    e.ptr = sv->amps;
    e.rows_per_box = rows_per_box;
    const cuuint64_t dims[2] = {16, (cuuint64_t)(sv->count * (uint64_t)sv->amp_bytes / 128)};
    const cuuint64_t strides[1] = {128};
    const cuuint32_t box[2] = {16, (cuuint32_t)rows_per_box};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&e.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, sv->amps, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return nullptr;
    if (sv->maps.size() >= 8) sv->maps.erase(sv->maps.begin());
    sv->maps.push_back(e);
    return &sv->maps.back().map;
}

template <typename A>
int set_swz_smem(int device, size_t bytes) {
    static size_t configured_dev[64] = {};  // the attribute is per device
    size_t& configured = configured_dev[device & 63];
    if (bytes > 48 * 1024 && configured < bytes) {
        CU(cudaFuncSetAttribute(k_tile_swz<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        CU(cudaFuncSetAttribute(k_tile_swz<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        configured = bytes;
    }
    return B2SV_OK;
}

// owner_only (with lazy_init): launch just the one tile that holds the pending basis index; the caller then
// records the rest of the register as "implied zeros" (b2sv::window).
int launch_tile(b2sv* sv, const Segment& seg, const TileOp* d_ops, int n_ops, const PlanConfig& cfg, int tile_io,
                bool fused_ops, bool timing, bool lazy_init, bool owner_only, const std::vector<SlotFill>& slots,
                const std::vector<double>& tables) {
    const bool bulk = tile_io != 1;
    TileArgs a{};
    fill_expand(a.tile_ex, seg.tile_mask);
    if (popcount64(seg.tile_mask | seg.common_ctrl) > kMaxFixed) return fail(B2SV_EINVAL, "tile + common controls exceed %d qubits", kMaxFixed);
    if (owner_only) {
        if (sv->n > kMaxFixed) return fail(B2SV_EINVAL, "register wider than %d qubits", kMaxFixed);
        fill_expand(a.cta_ex, low_mask(sv->n));           // every bit fixed: one tile
        a.cta_or = sv->basis_index & ~seg.tile_mask;      // ... the one that owns the basis index
    } else {
        fill_expand(a.cta_ex, seg.tile_mask | seg.common_ctrl);
        a.cta_or = seg.common_ctrl;
    }
    a.k = a.tile_ex.m;
    int lc = 0;
    while (lc < a.k && a.tile_ex.pos[lc] == lc) ++lc;  // contiguous low run of the tile
    const int lc_cap = sv->amp_bytes == 16 ? 10 : 11;  // at most 16 KiB per copy: keep enough chunks in flight
    if (lc > lc_cap) lc = lc_cap;
    if (((size_t)sv->amp_bytes << lc) < 16) return fail(B2SV_EINVAL, "tile chunk smaller than 16 bytes");
    a.lc = lc;
    a.n_ops = n_ops;
    a.ops = d_ops;
    a.tables = (const double2*)sv->d_tables;
    a.init_mode = lazy_init ? 1 : 0;
    a.init_index = sv->basis_index;
    a.tile_mask = seg.tile_mask;
    const uint64_t n_tiles = 1ull << (sv->n - a.cta_ex.m);
    const size_t smem = ((size_t)sv->amp_bytes << a.k) + (size_t)kTileMaxOps * sizeof(TileOp);
    // default path: TMA tensor copies with the hardware 128-byte swizzle (needs chunks of >= one 128-byte
    // row and a tile of >= one 1 KiB swizzle atom; tiny registers use the linear bulk-copy kernel below)
    const int row_shift = sv->amp_bytes == 16 ? 3 : 4;
    if (tile_io == 0 && lc >= row_shift && a.k >= row_shift + 3 && (1 << (lc - row_shift)) <= 256) {
        if (const CUtensorMap* map = state_tensor_map(sv, 1 << (lc - row_shift))) {
            TileArgsSwz as;
            as.map = *map;
            as.t = a;
            for (const SlotFill& f : slots) {  // the sweep's dense-block matrices ride in the kernel parameters
                if (sv->amp_bytes == 16) {
                    std::memcpy(&as.blk.m[(size_t)f.granule * kBlkGranuleElems], &tables[2 * f.mat_off], sizeof(double2) << (2 * f.kb));
                } else {  // complex64 sweeps compute in fp32: the same entries, narrowed once here
                    float2* dst = reinterpret_cast<float2*>(as.blk.m) + (size_t)f.granule * kBlkGranuleElems;
                    for (int e = 0; e < (1 << (2 * f.kb)); ++e)
                        dst[e] = make_float2((float)tables[2 * (f.mat_off + e)], (float)tables[2 * (f.mat_off + e) + 1]);
                }
            }
            const size_t swz_smem = smem + 16;  // tile + ops + the mbarrier
            return dispatch_dtype(sv, [&](auto* tag) -> int {
                using A = std::remove_pointer_t<decltype(tag)>;
                int rc = set_swz_smem<A>(sv->device, swz_smem);
                if (rc) return rc;
                // B2SV_TRACE=<file> B2SV_TRACE_LAUNCH=<i>: dump the phase timestamps of every CTA of the i-th swizzled tile
                // launch of this process (profiling aid: how load / compute / store of co-resident CTAs overlap)
                static const char* trace_path = std::getenv("B2SV_TRACE");
                static const long trace_launch = [] { const char* e = std::getenv("B2SV_TRACE_LAUNCH"); return e ? std::atol(e) : -1L; }();
                static long launch_no = 0;
                unsigned long long* d_trace = nullptr;
                if (trace_path && launch_no++ == trace_launch) {
                    CU(cudaMalloc(&d_trace, n_tiles * 5 * sizeof(unsigned long long)));
                    CU(cudaMemsetAsync(d_trace, 0, n_tiles * 5 * sizeof(unsigned long long), sv->stream));
                    as.t.trace = d_trace;
                }
                {
                    LaunchTimer lt(sv, timing, CLS_TILE, (lazy_init ? 1.0 : 2.0) * sv->amp_bytes * (double)(n_tiles << a.k));
                    if (fused_ops) k_tile_swz<A, true><<<(unsigned)n_tiles, kThreads, swz_smem, sv->stream>>>(as);
                    else k_tile_swz<A, false><<<(unsigned)n_tiles, kThreads, swz_smem, sv->stream>>>(as);
                }
                if (d_trace) {
                    std::vector<unsigned long long> h(n_tiles * 5);
                    CU(cudaStreamSynchronize(sv->stream));
                    CU(cudaMemcpy(h.data(), d_trace, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
                    cudaFree(d_trace);
                    if (FILE* f = std::fopen(trace_path, "wb")) {
                        std::fwrite(h.data(), sizeof(unsigned long long), h.size(), f);
                        std::fclose(f);
                    }
                }
                return B2SV_OK;
            });
        }
    }
    return dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        int rc = set_tile_smem<A>(sv->device, smem);
        if (rc) return rc;
        LaunchTimer lt(sv, timing, CLS_TILE, (lazy_init ? 1.0 : 2.0) * sv->amp_bytes * (double)(n_tiles << a.k));
        if (bulk && fused_ops)
            k_tile<A, true, true><<<(unsigned)n_tiles, kThreads, smem, sv->stream>>>((A*)sv->amps, a);
        else if (bulk)
            k_tile<A, true, false><<<(unsigned)n_tiles, kThreads, smem, sv->stream>>>((A*)sv->amps, a);
        else if (fused_ops)
            k_tile<A, false, true><<<(unsigned)n_tiles, kThreads, smem, sv->stream>>>((A*)sv->amps, a);
        else
            k_tile<A, false, false><<<(unsigned)n_tiles, kThreads, smem, sv->stream>>>((A*)sv->amps, a);
        return B2SV_OK;
    });
}

int launch_low6(b2sv* sv, const Segment& seg, const TileOp* d_ops, int n_ops, bool timing, bool lazy_init, bool owner_only = false) {
    if (n_ops > kTileMaxOps) return fail(B2SV_EINVAL, "low-qubit sweep with %d ops (max %d)", n_ops, kTileMaxOps);
    Low6Args a{};
    const uint64_t fixed = seg.tile_mask | seg.common_ctrl;
    if (popcount64(fixed) > kMaxFixed) return fail(B2SV_EINVAL, "block + common controls exceed %d qubits", kMaxFixed);
    if (owner_only) {  // just the 64-amplitude block that holds the pending basis index
        if (sv->n > kMaxFixed) return fail(B2SV_EINVAL, "register wider than %d qubits", kMaxFixed);
        fill_expand(a.blk_ex, low_mask(sv->n));
        a.blk_or = sv->basis_index & ~seg.tile_mask;
    } else {
        fill_expand(a.blk_ex, fixed);
        a.blk_or = seg.common_ctrl;
    }
    a.n_blocks = 1ull << (sv->n - a.blk_ex.m);
    a.n_ops = n_ops;
    a.ops = d_ops;
    a.tables = (const double2*)sv->d_tables;
    a.init_mode = lazy_init ? 1 : 0;
    a.init_index = sv->basis_index;
    const size_t smem = (size_t)kTileMaxOps * sizeof(TileOp);
    const uint64_t n_groups = (a.n_blocks + kLow6Blocks - 1) / kLow6Blocks;
    const uint64_t want = (n_groups + (kThreads / 32) - 1) / (kThreads / 32);
    const uint64_t cap = (uint64_t)sv->sm_count * 8;
    const unsigned grid = (unsigned)(want < cap ? (want ? want : 1) : cap);
    return dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, timing, CLS_LOW6, (lazy_init ? 1.0 : 2.0) * sv->amp_bytes * (double)(a.n_blocks << kLowBits));
        k_low6<A><<<grid, kThreads, smem, sv->stream>>>((A*)sv->amps, a);
        return B2SV_OK;
    });
}

// Encode one permutation gate as word ops (see kernels.cuh), in EXECUTION order.
void encode_perm_gate(const PGate& g, int n, std::vector<uint4>& out) {
    auto row = [](int r) { return (uint32_t)r * (uint32_t)kPermRowBytes; };
    const uint32_t ones = row(n);
    std::vector<int> ctrls;
    for (int q = 0; q < 64; ++q)
        if ((g.ctrl >> q) & 1) ctrls.push_back(q);
    const size_t direct = g.op == B2SV_OP_X ? 3 : 2;  // control slots left in the final op(s)
    // Long control lists: fold controls into temporaries T1 (and T2 for very long lists):
    //   T1 ^= c0&c1&c2 ; [T2 ^= T1&c3&c4 ; T1' ...] -- kept simple: chain through two temporaries.
    std::vector<uint4> prologue;
    std::vector<int> eff(ctrls.begin(), ctrls.end());  // effective control rows of the final op
    int next_temp = n + 1;
    while (eff.size() > direct) {
        // take three controls, AND them into a fresh temporary, replace them by the temporary
        const int t = next_temp++;
        uint4 op;
        op.x = row(t);
        op.y = row(eff[0]);
        op.z = row(eff[1]);
        op.w = row(eff[2]);
        prologue.push_back(op);
        eff.erase(eff.begin(), eff.begin() + 3);
        eff.insert(eff.begin(), t);
    }
    for (const uint4& op : prologue) out.push_back(op);
    auto ctl = [&](size_t i) { return i < eff.size() ? row(eff[i]) : ones; };
    if (g.op == B2SV_OP_X) {
        out.push_back(make_uint4(row(g.t0), ctl(0), ctl(1), ctl(2)));
    } else {
        out.push_back(make_uint4(row(g.t1), row(g.t0), ctl(0), ctl(1)));
        out.push_back(make_uint4(row(g.t0), row(g.t1), ctl(0), ctl(1)));
        out.push_back(make_uint4(row(g.t1), row(g.t0), ctl(0), ctl(1)));
    }
    for (size_t i = prologue.size(); i-- > 0;) out.push_back(prologue[i]);  // un-compute the temporaries
}

static_assert(kPlanMaxTileOps + (kPlanMaxTileOps + 1) / 2 <= kTileMaxOps, "planner and kernel disagree on the ops per single-batch sweep");
static_assert(kPermMaxCtrlX == 3 + 2 * kPermTemps && kPermMaxCtrlSwap == 2 + 2 * kPermTemps,
              "planner's control limits must match the scratch rows of k_perm");

// ---- run-time specialised permutation kernels (perm_jit.hpp) -------------------------------------

struct JitEntry {
    std::mutex mu;
    int state = 0;  // 0 compiling, 1 cubin ready, -1 failed
    std::vector<char> cubin;
    std::string log;
    std::map<int, std::pair<cudaLibrary_t, cudaKernel_t>> loaded;  // per device
    uint64_t uses = 0;
};
std::mutex g_jit_mutex;
std::map<std::string, std::shared_ptr<JitEntry>> g_jit;

std::string perm_key(const std::vector<PGate>& gates, int n, int dtype, int n_global = -1) {
    std::string k;
    k.reserve(gates.size() * 11 + 8);
    k.push_back((char)n);
    k.push_back((char)dtype);
    k.push_back((char)n_global);  // -1: single-register kernel, >= 0: distributed gather over 2^n_global shards
    for (const PGate& g : gates) {
        k.push_back((char)g.op);
        k.push_back((char)g.t0);
        k.push_back((char)(g.op == B2SV_OP_SWAP ? g.t1 : 0));
        k.append(reinterpret_cast<const char*>(&g.ctrl), 8);
    }
    return k;
}

void jit_compile_into(const std::shared_ptr<JitEntry>& e, std::vector<PGate> gates, int n, int dtype, int n_global) {
    std::vector<char> cubin;
    std::string log;
    const bool ok = nvrtc_compile(n_global < 0 ? perm_jit_source(gates, n, dtype) : perm_dist_source(gates, n, n_global, dtype), cubin, log);
    std::lock_guard<std::mutex> lk(e->mu);
    e->cubin = std::move(cubin);
    e->log = std::move(log);
    e->state = ok ? 1 : -1;
}

// Background compiles run on joinable threads that are joined when the library is unloaded / the process
// exits, so that no thread is still inside NVRTC (or touching g_jit) while the statics are torn down.
struct JitWorkers {
    std::mutex mu;
    std::vector<std::thread> threads;
    void spawn(const std::shared_ptr<JitEntry>& e, const std::vector<PGate>& gates, int n, int dtype) {
        std::lock_guard<std::mutex> lk(mu);
        threads.emplace_back(jit_compile_into, e, gates, n, dtype, -1);
    }
    ~JitWorkers() {
        std::lock_guard<std::mutex> lk(mu);
        for (auto& t : threads)
            if (t.joinable()) t.join();
    }
};
JitWorkers g_jit_workers;

// Returns the specialised kernel for this run if it is ready (loading it on first use), else null.
// n_global < 0: the single-register sweep (b2sv_perm_jit); >= 0: the distributed gather (b2sv_perm_dist).
cudaKernel_t jit_lookup(b2sv* sv, const std::vector<PGate>& gates, int jit_mode, int n_global = -1) {
    if (jit_mode == 1 || !nvrtc_api().ok) return nullptr;
    std::shared_ptr<JitEntry> e;
    {
        std::lock_guard<std::mutex> lk(g_jit_mutex);
        std::string key = perm_key(gates, sv->n, sv->dtype, n_global);
        auto it = g_jit.find(key);
        if (it == g_jit.end()) {
            if (g_jit.size() >= 512) {  // bound the cache: drop entries nobody holds a loaded kernel of any more
                for (auto jt = g_jit.begin(); jt != g_jit.end();) {
                    std::unique_lock<std::mutex> el(jt->second->mu, std::try_to_lock);
                    if (el.owns_lock() && jt->second->state != 0) {
                        for (auto& ld : jt->second->loaded) cudaLibraryUnload(ld.second.first);
                        el.unlock();
                        jt = g_jit.erase(jt);
                    } else {
                        ++jt;
                    }
                }
            }
            e = std::make_shared<JitEntry>();
            g_jit.emplace(std::move(key), e);
            if (jit_mode == 2) {
                jit_compile_into(e, gates, sv->n, sv->dtype, n_global);
            } else {
                g_jit_workers.spawn(e, gates, sv->n, sv->dtype);
                return nullptr;  // first sighting: the generic kernel runs while NVRTC works
            }
        } else {
            e = it->second;
        }
    }
    std::lock_guard<std::mutex> lk(e->mu);
    e->uses++;
    if (e->state != 1) return nullptr;
    auto it = e->loaded.find(sv->device);
    if (it == e->loaded.end()) {
        cudaLibrary_t lib = nullptr;
        cudaKernel_t k = nullptr;
        if (cudaLibraryLoadData(&lib, e->cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
            cudaLibraryGetKernel(&k, lib, n_global < 0 ? "b2sv_perm_jit" : "b2sv_perm_dist") != cudaSuccess) {
            cudaGetLastError();
            e->state = -1;
            e->log = "loading the specialised kernel failed";
            return nullptr;
        }
        it = e->loaded.emplace(sv->device, std::make_pair(lib, k)).first;
    }
    return it->second.second;
}

// Device copy of a permutation run as DistGate records (interpreter / scatter kernels).
int upload_dist_gates(b2sv* sv, const std::vector<PGate>& pg) {
    std::vector<DistGate> dg(pg.size());
    for (size_t j = 0; j < pg.size(); ++j) {
        std::memset(&dg[j], 0, sizeof(DistGate));
        dg[j].ctrl = pg[j].ctrl;
        dg[j].op = pg[j].op;
        dg[j].t0 = pg[j].t0;
        dg[j].t1 = pg[j].t1;
    }
    const size_t bytes = (dg.size() + 1) * sizeof(DistGate);
    if (bytes > sv->d_dist_gates_cap) {
        CU(cudaStreamSynchronize(sv->stream));
        cudaFree(sv->d_dist_gates);
        sv->d_dist_gates = nullptr;
        sv->d_dist_gates_cap = 0;
        CU(cudaMalloc(&sv->d_dist_gates, bytes * 2));
        sv->d_dist_gates_cap = bytes * 2;
    }
    if (!dg.empty()) {
        CU(cudaMemcpyAsync(sv->d_dist_gates, dg.data(), dg.size() * sizeof(DistGate), cudaMemcpyHostToDevice, sv->stream));
        CU(cudaStreamSynchronize(sv->stream));  // dg is pageable host memory that dies with this frame
    }
    return B2SV_OK;
}

constexpr int kScatterMaxBits = 16;  // window sources up to 2^16 amplitudes are scattered, larger ones gathered

int launch_perm(b2sv* sv, const Plan& plan, const Segment& seg, bool timing, int jit_mode) {
    // a small live window as the source: zero-fill the output at write bandwidth and scatter the window forward
    if (sv->window && popcount64(~sv->win_mask & low_mask(sv->n)) <= kScatterMaxBits && !std::getenv("B2SV_NO_SCATTER")) {
        std::vector<PGate> gates;
        gates.reserve(seg.gates.size());
        for (int gi : seg.gates) gates.push_back(plan.gates[gi]);
        if (int rc = upload_dist_gates(sv, gates)) return rc;
        DistPermArgs a;
        std::memset(&a, 0, sizeof(a));
        a.src[0] = sv->amps;
        a.wmask[0] = sv->win_mask;
        a.wval[0] = sv->win_value;
        a.out = sv->scratch;
        const uint64_t cnt = 1ull << popcount64(~sv->win_mask & low_mask(sv->n));
        int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
            using A = std::remove_pointer_t<decltype(tag)>;
            LaunchTimer lt(sv, timing, jit_mode == 1 ? CLS_PERM : CLS_PERM_JIT, 1.0 * sv->amp_bytes * (double)sv->count);
            if (int frc = fill_zero(sv, sv->scratch, sv->count * sv->amp_bytes)) return frc;
            k_perm_scatter<A><<<dim3((unsigned)((cnt + kThreads - 1) / kThreads), 1), kThreads, 0, sv->stream>>>(a, (const DistGate*)sv->d_dist_gates,
                                                                                                          (int)gates.size(), sv->n);
            return B2SV_OK;
        });
        if (rc) return rc;
        CU(cudaGetLastError());
        sv->window = false;
        std::swap(sv->amps, sv->scratch);
        return B2SV_OK;
    }
    // long runs on big registers: straight-line register code from NVRTC once it is ready
    if (jit_mode == 2 || (seg.gates.size() >= 64 && sv->n >= 20)) {
        std::vector<PGate> gates;
        gates.reserve(seg.gates.size());
        for (int gi : seg.gates) gates.push_back(plan.gates[gi]);
        if (cudaKernel_t k = jit_lookup(sv, gates, jit_mode)) {
            const void* in = sv->amps;
            void* out = sv->scratch;
            unsigned long long wmask = sv->window ? sv->win_mask : 0ull, wval = sv->window ? sv->win_value : 0ull;
            void* args[] = {(void*)&in, (void*)&out, (void*)&wmask, (void*)&wval};
            const uint64_t blocks = sv->count / (uint64_t)(kPermThreads * 32);
            {
                LaunchTimer lt(sv, timing, CLS_PERM_JIT, (sv->window ? 1.0 : 2.0) * sv->amp_bytes * (double)sv->count);
                CU(cudaLaunchKernel((const void*)k, dim3((unsigned)blocks), dim3(kPermThreads), args, 0, sv->stream));
            }
            sv->window = false;  // the sweep wrote every amplitude
            std::swap(sv->amps, sv->scratch);
            return B2SV_OK;
        }
    }
    // split the run so that each chunk's program fits constant memory
    size_t pos = 0;
    while (pos < seg.gates.size()) {
        std::vector<uint4> fwd;  // ops in execution order, gate by gate
        std::vector<size_t> gate_end;
        size_t e = pos;
        while (e < seg.gates.size()) {
            const size_t before = fwd.size();
            encode_perm_gate(plan.gates[seg.gates[e]], sv->n, fwd);
            if (fwd.size() > (size_t)kPermMaxOps) {
                fwd.resize(before);
                break;
            }
            gate_end.push_back(fwd.size());
            ++e;
        }
        if (e == pos) return fail(B2SV_EINVAL, "permutation gate too large");
        // The sweep gathers: out[j] = in[f^-1(j)], and f^-1 applies the (self-inverse) gates in reverse
        // order.  Each gate's op group is a palindrome (compute temporaries, act, un-compute), so
        // reversing the whole op list reverses the gates and keeps every group valid.
        std::vector<uint4> prog(fwd.rbegin(), fwd.rend());
        // SWAP groups (b^=a, a^=b, b^=a) are palindromes too.
        ConstGuard guard(sv);
        CU(cudaMemcpyToSymbolAsync(c_perm_ops, prog.data(), prog.size() * sizeof(uint4), 0, cudaMemcpyHostToDevice,
                                   sv->stream));
        const size_t smem = (size_t)(sv->n + 1 + kPermTemps) * kPermRowBytes;
        const uint64_t blocks = sv->count / (uint64_t)(kPermThreads * 32);
        int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
            using A = std::remove_pointer_t<decltype(tag)>;
            static bool configured_dev[64] = {};  // the attribute is per device
            bool& configured = configured_dev[sv->device & 63];
            if (!configured && smem > 48 * 1024) {
                CU(cudaFuncSetAttribute(k_perm<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                configured = true;
            }
            LaunchTimer lt(sv, timing, CLS_PERM, (sv->window ? 1.0 : 2.0) * sv->amp_bytes * (double)sv->count);
            k_perm<A><<<(unsigned)blocks, kPermThreads, smem, sv->stream>>>((const A*)sv->amps, (A*)sv->scratch, sv->n,
                                                                           (int)prog.size(), sv->window ? sv->win_mask : 0ull,
                                                                           sv->window ? sv->win_value : 0ull);
            return B2SV_OK;
        });
        if (rc) return rc;
        CU(cudaGetLastError());
        sv->window = false;  // the sweep wrote every amplitude
        std::swap(sv->amps, sv->scratch);
        pos = e;
    }
    return B2SV_OK;
}

int check(const b2sv* sv) {
    if (!sv) return fail(B2SV_EINVAL, "null handle");
    return B2SV_OK;
}

// Validates and uploads a subspace program; *count_bits = number of low index bits that can be
// non-zero in a member (top ZEROS runs reachable from `entry` shrink the scan range).
// A program that starts with FREE on the bits [0, len): aligned groups of up to 2^len amplitudes are members
// together, so the projection kernels walk the program once per group.
constexpr int kProjectGroupBits = 3;
static_assert((1 << kProjectGroupBits) == kProjectGroup, "group size");
int low_free_bits(const b2sv_subinstr* prog, int entry) {
    const b2sv_subinstr& in = prog[entry];
    return (in.op == B2SV_SUB_FREE && in.pos == 0) ? in.len : 0;
}

// Number of members of the subspace a (validated) program accepts: what a projection has to read.
double subprog_members(const b2sv_subinstr* prog, int n_instr, int entry) {
    std::vector<double> memo(n_instr, -1.0);
    std::vector<int> stack{entry};
    while (!stack.empty()) {
        const int pc = stack.back();
        const b2sv_subinstr& in = prog[pc];
        if (memo[pc] >= 0) {
            stack.pop_back();
            continue;
        }
        if (in.op == B2SV_SUB_END || in.op == B2SV_SUB_REJECT) {
            memo[pc] = in.op == B2SV_SUB_END ? 1.0 : 0.0;
            stack.pop_back();
            continue;
        }
        const bool branch = in.op == B2SV_SUB_BRANCH;
        if (memo[in.next0] < 0 || (branch && memo[in.next1] < 0)) {
            if ((int)stack.size() > 4 * n_instr + 8) return std::ldexp(1.0, 62);  // malformed (cyclic) program
            if (memo[in.next0] < 0) stack.push_back(in.next0);
            if (branch && memo[in.next1] < 0) stack.push_back(in.next1);
            continue;
        }
        if (branch) memo[pc] = memo[in.next0] + memo[in.next1];
        else if (in.op == B2SV_SUB_FREE) memo[pc] = std::ldexp(memo[in.next0], in.len);
        else memo[pc] = memo[in.next0];
        stack.pop_back();
    }
    return memo[entry];
}

int upload_subprog(b2sv* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits, int* count_bits,
                   int max_target_qubits = -1, bool* full_cover = nullptr, uint64_t* head_zeros = nullptr, bool* pure = nullptr) {
    if (max_target_qubits < 0) max_target_qubits = sv->n;
    if (!prog || n_instr <= 0) return fail(B2SV_EINVAL, "empty subspace program");
    if (n_instr > kMaxSubInstr) return fail(B2SV_EINVAL, "subspace program has %d instructions (max %d)", n_instr, kMaxSubInstr);
    if (entry < 0 || entry >= n_instr) return fail(B2SV_EINVAL, "bad entry %d", entry);
    if (target_qubits < 0 || target_qubits > max_target_qubits)
        return fail(B2SV_EINVAL, "target_qubits %d outside [0,%d]", target_qubits, max_target_qubits);
    uint64_t covered = 0;
    for (int i = 0; i < n_instr; ++i) {
        const b2sv_subinstr& in = prog[i];
        if (in.op > B2SV_SUB_REJECT) return fail(B2SV_EINVAL, "subspace instruction %d: bad op %d", i, (int)in.op);
        if (in.op == B2SV_SUB_END || in.op == B2SV_SUB_REJECT) continue;
        const int len = in.op == B2SV_SUB_BRANCH ? 1 : in.len;
        if (len < 1 || in.pos + len > target_qubits) return fail(B2SV_EINVAL, "subspace instruction %d: bits [%d,%d) outside the %d target qubits", i, (int)in.pos, in.pos + len, target_qubits);
        if (in.next0 >= n_instr || (in.op == B2SV_SUB_BRANCH && in.next1 >= n_instr))
            return fail(B2SV_EINVAL, "subspace instruction %d: jump out of range", i);
        covered |= (len >= 64 ? ~0ull : ((1ull << len) - 1)) << in.pos;
    }
    const uint64_t all = target_qubits >= 64 ? ~0ull : ((1ull << target_qubits) - 1);
    // single-register programs mention every target qubit; per-rank specialised programs (sharded
    // form, max_target_qubits == 64) have had their rank bits resolved on the host
    if (max_target_qubits != 64 && (covered & all) != all)
        return fail(B2SV_EINVAL, "subspace program does not constrain all %d target qubits", target_qubits);
    if (full_cover) *full_cover = (covered & all) == all;
    // shrink the scan range by the ZEROS runs at the top of the index that every path goes through:
    // conservative version -- follow the entry chain while it is made of ZEROS instructions.
    int bits = target_qubits;
    {
        // every index goes through the straight-line head of the program (ZEROS / FREE have one successor)
        uint64_t zero_mask = 0;
        int pc = entry;
        for (int guard = 0; guard < n_instr && (prog[pc].op == B2SV_SUB_ZEROS || prog[pc].op == B2SV_SUB_FREE); ++guard) {
            if (prog[pc].op == B2SV_SUB_ZEROS) zero_mask |= (prog[pc].len >= 64 ? ~0ull : ((1ull << prog[pc].len) - 1)) << prog[pc].pos;
            pc = prog[pc].next0;
        }
        while (bits > 0 && ((zero_mask >> (bits - 1)) & 1)) --bits;
        if (head_zeros) *head_zeros = zero_mask & low_mask(bits);
        // branch-free program: the straight-line head runs into END, so membership IS "the always-zero bits are clear"
        if (pure) *pure = prog[pc].op == B2SV_SUB_END;
    }
    if (count_bits) *count_bits = bits;
    if (n_instr > sv->d_sub_cap) {
        CU(cudaStreamSynchronize(sv->stream));
        cudaFree(sv->d_sub);
        sv->d_sub = nullptr;
        sv->d_sub_cap = 0;
        CU(cudaMalloc(&sv->d_sub, sizeof(b2sv_subinstr) * (size_t)(n_instr + 64)));
        sv->d_sub_cap = n_instr + 64;
    }
    CU(cudaMemcpyAsync(sv->d_sub, prog, (size_t)n_instr * sizeof(b2sv_subinstr), cudaMemcpyHostToDevice, sv->stream));
    return B2SV_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================

extern "C" {

int b2sv_abi_version(void) { return B2SV_ABI_VERSION; }

const char* b2sv_last_error(void) { return g_err.c_str(); }

int b2sv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int b2sv_create(int n_qubits, int dtype, int device, b2sv_t** out) {
    if (!out) return fail(B2SV_EINVAL, "out is null");
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > 40) return fail(B2SV_EINVAL, "n_qubits %d outside [1,40]", n_qubits);
    if (dtype != B2SV_C128 && dtype != B2SV_C64) return fail(B2SV_EINVAL, "unknown dtype %d", dtype);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(B2SV_ENODEV, "no CUDA device available (%s); this engine has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(B2SV_EINVAL, "device %d outside [0,%d)", device, ndev);
    CU(cudaSetDevice(device));
    b2sv* sv = new b2sv();
    sv->n = n_qubits;
    sv->dtype = dtype;
    sv->device = device;
    sv->amp_bytes = dtype == B2SV_C64 ? 8 : 16;
    sv->count = 1ull << n_qubits;
    cudaDeviceGetAttribute(&sv->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (sv->sm_count <= 0) sv->sm_count = 148;
    e = cudaMalloc(&sv->amps, sv->count * sv->amp_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        const unsigned long long want = (unsigned long long)(sv->count * sv->amp_bytes);
        delete sv;
        return fail(B2SV_ENOMEM, "cudaMalloc of %llu bytes failed: %s", want, cudaGetErrorString(e));
    }
    e = cudaStreamCreateWithFlags(&sv->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        cudaFree(sv->amps);
        delete sv;
        return fail(B2SV_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e));
    }
    sv->n_partial = sv->sm_count * 8;
    e = cudaMalloc(&sv->d_partial, sizeof(double) * (sv->n_partial + 1));
    if (e != cudaSuccess) {
        cudaStreamDestroy(sv->stream);
        cudaFree(sv->amps);
        delete sv;
        return fail(B2SV_ENOMEM, "cudaMalloc failed: %s", cudaGetErrorString(e));
    }
    *out = sv;
    int rc = b2sv_set_basis(sv, 0);
    if (rc) {
        b2sv_destroy(sv);
        *out = nullptr;
    }
    return rc;
}

int b2sv_destroy(b2sv_t* sv) {
    if (!sv) return B2SV_OK;
    cudaSetDevice(sv->device);
    cudaStreamSynchronize(sv->stream);
    for (auto& t : sv->timed) {
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    for (auto& p : sv->apply_events) {
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    for (auto e : sv->event_pool) cudaEventDestroy(e);
    for (auto e : sv->marks)
        if (e) cudaEventDestroy(e);
    b2sv_ipc_detach(sv);
    if (sv->copy_stream) {
        cudaStreamSynchronize(sv->copy_stream);
        cudaStreamDestroy(sv->copy_stream);
    }
    for (int i = 0; i < 2; ++i) {
        cudaFree(sv->d_col[i]);
        if (sv->ev_col_kernel[i]) cudaEventDestroy(sv->ev_col_kernel[i]);
        if (sv->ev_col_copy[i]) cudaEventDestroy(sv->ev_col_copy[i]);
    }
    if (sv->ev_ready) cudaEventDestroy(sv->ev_ready);
    if (sv->ev_done) cudaEventDestroy(sv->ev_done);
    cudaFree(sv->d_dist_gates);
    cudaFree(sv->amps);
    cudaFree(sv->scratch);
    cudaFree(sv->d_ops);
    cudaFree(sv->d_tables);
    cudaFree(sv->d_partial);
    cudaFree(sv->d_sub);
    cudaFree(sv->d_gather);
    cudaStreamDestroy(sv->stream);
    cudaGetLastError();
    delete sv;
    return B2SV_OK;
}

int b2sv_set_basis(b2sv_t* sv, uint64_t index) {
    if (int rc = check(sv)) return rc;
    if (index >= sv->count) return fail(B2SV_EINVAL, "basis index %llu >= 2^%d", (unsigned long long)index, sv->n);
    sv->sre = 1.0;
    sv->sim = 0.0;
    // lazy: the first full tile sweep of the next b2sv_apply generates the state in shared memory;
    // any other reader materialises it first
    sv->pending_basis = true;
    sv->window = false;
    sv->zero = false;
    sv->basis_index = index;
    return B2SV_OK;
}

int b2sv_set_zero(b2sv_t* sv) {
    if (int rc = check(sv)) return rc;
    CU(cudaSetDevice(sv->device));
    sv->sre = 1.0;
    sv->sim = 0.0;
    sv->pending_basis = false;
    sv->window = false;
    sv->zero = true;  // lazy: nothing is written until somebody needs the zeros (materialise)
    return B2SV_OK;
}

int b2sv_fill_random(b2sv_t* sv, uint64_t seed, int support_bits, uint64_t index_base) {
    NvtxRange nvtx_range("b2sv_fill_random");
    if (int rc = check(sv)) return rc;
    if (support_bits < 0 || support_bits > 63) return fail(B2SV_EINVAL, "support_bits %d outside [0,63]", support_bits);
    CU(cudaSetDevice(sv->device));
    sv->sre = 1.0;
    sv->sim = 0.0;
    sv->pending_basis = false;
    sv->window = false;
    sv->zero = false;
    const uint64_t support = 1ull << support_bits;
    const double scale = std::sqrt(1.5 / (double)support);  // E|psi_i|^2 = 2/3 scale^2: unit norm in expectation
    int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, sv->timing_on, CLS_INIT, (double)sv->amp_bytes * (double)sv->count);
        k_fill_random<A><<<grid_for(sv, sv->count), kThreads, 0, sv->stream>>>((A*)sv->amps, sv->count, support, seed, index_base, scale);
        return B2SV_OK;
    });
    if (rc) return rc;
    CU(cudaGetLastError());
    return B2SV_OK;
}

// Host buffers are complex128.  complex64 states convert through a bounded device staging buffer.
static const uint64_t kStageAmps = 1ull << 22;  // 64 MiB of complex128

int b2sv_upload(b2sv_t* sv, const void* host) {
    NvtxRange nvtx_range("b2sv_upload");
    if (int rc = check(sv)) return rc;
    if (!host) return fail(B2SV_EINVAL, "host buffer is null");
    CU(cudaSetDevice(sv->device));
    sv->sre = 1.0;
    sv->sim = 0.0;
    sv->pending_basis = false;
    sv->window = false;
    sv->zero = false;
    if (sv->dtype == B2SV_C128) {
        CU(cudaMemcpyAsync(sv->amps, host, sv->count * 16, cudaMemcpyHostToDevice, sv->stream));
        CU(cudaStreamSynchronize(sv->stream));
        return B2SV_OK;
    }
    const uint64_t chunk = sv->count < kStageAmps ? sv->count : kStageAmps;
    double2* stage = nullptr;
    CU(cudaMalloc(&stage, chunk * 16));
    for (uint64_t off = 0; off < sv->count; off += chunk) {
        cudaError_t e = cudaMemcpyAsync(stage, (const char*)host + off * 16, chunk * 16, cudaMemcpyHostToDevice, sv->stream);
        if (e == cudaSuccess) {
            sv->stats.launches_by_class[CLS_INIT]++;
            k_convert<float2, double2><<<grid_for(sv, chunk), kThreads, 0, sv->stream>>>((float2*)sv->amps + off, stage, chunk);
            e = cudaStreamSynchronize(sv->stream);
        }
        if (e != cudaSuccess) {
            cudaFree(stage);
            return fail(B2SV_ECUDA, "upload failed: %s", cudaGetErrorString(e));
        }
    }
    cudaFree(stage);
    return B2SV_OK;
}

int b2sv_download(b2sv_t* sv, void* host) {
    NvtxRange nvtx_range("b2sv_download");
    if (int rc = check(sv)) return rc;
    if (!host) return fail(B2SV_EINVAL, "host buffer is null");
    CU(cudaSetDevice(sv->device));
    if (int rc = materialise(sv)) return rc;
    if (int rc = flush_scalar(sv, false)) return rc;
    if (sv->dtype == B2SV_C128) {
        CU(cudaMemcpyAsync(host, sv->amps, sv->count * 16, cudaMemcpyDeviceToHost, sv->stream));
        CU(cudaStreamSynchronize(sv->stream));
        return B2SV_OK;
    }
    const uint64_t chunk = sv->count < kStageAmps ? sv->count : kStageAmps;
    double2* stage = nullptr;
    CU(cudaMalloc(&stage, chunk * 16));
    for (uint64_t off = 0; off < sv->count; off += chunk) {
        sv->stats.launches_by_class[CLS_INIT]++;
        k_convert<double2, float2><<<grid_for(sv, chunk), kThreads, 0, sv->stream>>>(stage, (const float2*)sv->amps + off, chunk);
        cudaError_t e = cudaMemcpyAsync((char*)host + off * 16, stage, chunk * 16, cudaMemcpyDeviceToHost, sv->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(sv->stream);
        if (e != cudaSuccess) {
            cudaFree(stage);
            return fail(B2SV_ECUDA, "download failed: %s", cudaGetErrorString(e));
        }
    }
    cudaFree(stage);
    return B2SV_OK;
}

int b2sv_apply(b2sv_t* sv, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts) {
    NvtxRange nvtx_range("b2sv_apply");
    if (int rc = check(sv)) return rc;
    if (n_gates > 0 && !gates) return fail(B2SV_EINVAL, "gates is null");
    CU(cudaSetDevice(sv->device));
    if (sv->zero) {  // unitary gates map the zero vector to itself
        sv->stats.gates_in += n_gates;
        return B2SV_OK;
    }
    PlanConfig cfg = default_config(sv->n, sv->dtype, opts);
    const bool timing = opts && opts->timing;
    sv->timing_on = timing;
    const int tile_io = opts ? opts->tile_io : 0;
    const int jit_mode = opts ? opts->jit : 0;
    const bool tail_window = opts && (opts->reserved & 32);
    if (sv->n < 13) cfg.perm_runs = false;
    // same gate list and planner options as last time: reuse the plan and what is already on the device
    uint64_t key = 1469598103934665603ull, key2 = 0;
    {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(gates);
        const size_t nw = n_gates * sizeof(b2sv_gate) / 8;
        for (size_t i = 0; i < nw; ++i) {
            key = (key ^ w[i]) * 1099511628211ull;
            key2 += w[i] * (2 * i + 1);
        }
        const uint64_t o = (uint64_t)cfg.fuse | ((uint64_t)cfg.perm_runs << 1) | ((uint64_t)cfg.mux << 2) | ((uint64_t)cfg.phases << 3) | ((uint64_t)cfg.low6 << 4) | ((uint64_t)cfg.mono << 5) |
                           ((uint64_t)cfg.tile_qubits << 8) | ((uint64_t)cfg.lookahead << 16) | ((uint64_t)sv->scratch_failed << 40);
        key = (key ^ o) * 1099511628211ull;
    }
    // a hash hit alone could replay another circuit's plan: compare the gate bytes too (cheap next to one sweep)
    const bool reuse = sv->plan_valid && sv->plan_key == key && sv->plan_key2 == key2 && sv->plan_ngates == n_gates &&
                       sv->plan_gate_copy.size() == n_gates * sizeof(b2sv_gate) &&
                       (n_gates == 0 || std::memcmp(sv->plan_gate_copy.data(), gates, n_gates * sizeof(b2sv_gate)) == 0);
    if (!reuse) {
        for (size_t j = 0; j < n_gates; ++j) {
            const b2sv_gate& g = gates[j];
            if (g.op > B2SV_OP_PHASE) return fail(B2SV_EINVAL, "gate %zu: unknown op %d", j, (int)g.op);
            if (g.op == B2SV_OP_SWAP && g.t0 == g.t1) return fail(B2SV_EINVAL, "gate %zu: SWAP of a qubit with itself", j);
            const uint64_t all = sv->n >= 64 ? ~0ull : ((1ull << sv->n) - 1);
            if (g.ctrl_mask & ~all) return fail(B2SV_EINVAL, "gate %zu: control outside the %d-qubit register", j, sv->n);
            if (g.op != B2SV_OP_PHASE) {
                if (g.t0 >= sv->n) return fail(B2SV_EINVAL, "gate %zu: target %d outside the register", j, (int)g.t0);
                if ((g.ctrl_mask >> g.t0) & 1) return fail(B2SV_EINVAL, "gate %zu: target %d is also a control", j, (int)g.t0);
            }
            if (g.op == B2SV_OP_SWAP) {
                if (g.t1 >= sv->n) return fail(B2SV_EINVAL, "gate %zu: target %d outside the register", j, (int)g.t1);
                if ((g.ctrl_mask >> g.t1) & 1) return fail(B2SV_EINVAL, "gate %zu: target %d is also a control", j, (int)g.t1);
            }
        }
        sv->plan_valid = false;
        bool scratch_ok = false;
        if (cfg.fuse && cfg.perm_runs) {
            // only pay for the second buffer if the circuit has a run that wants it
            size_t run = 0, best = 0;
            for (size_t j = 0; j < n_gates; ++j) {
                if (gates[j].op == B2SV_OP_X || gates[j].op == B2SV_OP_SWAP) best = std::max(best, ++run);
                else run = 0;
            }
            if ((int)best >= cfg.perm_min) scratch_ok = ensure_scratch(sv) == B2SV_OK;
        }
        NvtxRange nvtx_plan("b2sv plan (host)");
        sv->plan = make_plan(gates, n_gates, cfg, scratch_ok);
        Plan& np = sv->plan;
        // tile ops of all tile segments, one upload
        std::vector<TileOp> host_ops;
        sv->plan_seg_first.assign(np.segs.size(), 0);
        sv->plan_seg_nops.assign(np.segs.size(), 0);
        sv->plan_seg_slots.assign(np.segs.size(), {});
        for (size_t s = 0; s < np.segs.size(); ++s) {
            const Segment& seg = np.segs[s];
            if (seg.cls != SEG_TILE && seg.cls != SEG_LOW6) continue;
            sv->plan_seg_first[s] = host_ops.size();
            build_segment_ops(np, seg, cfg, host_ops, &sv->plan_seg_slots[s]);
            sv->plan_seg_nops[s] = (int)(host_ops.size() - sv->plan_seg_first[s]);
        }
        if (!host_ops.empty()) {
            const size_t bytes = host_ops.size() * sizeof(TileOp);
            if (bytes > sv->d_ops_cap) {
                CU(cudaStreamSynchronize(sv->stream));
                cudaFree(sv->d_ops);
                sv->d_ops = nullptr;
                sv->d_ops_cap = 0;
                CU(cudaMalloc(&sv->d_ops, bytes * 2));
                sv->d_ops_cap = bytes * 2;
            }
            CU(cudaMemcpyAsync(sv->d_ops, host_ops.data(), bytes, cudaMemcpyHostToDevice, sv->stream));
        }
        if (!np.mux_tables.empty()) {
            const size_t bytes = np.mux_tables.size() * sizeof(double);
            if (bytes > sv->d_tables_cap) {
                CU(cudaStreamSynchronize(sv->stream));
                cudaFree(sv->d_tables);
                sv->d_tables = nullptr;
                sv->d_tables_cap = 0;
                CU(cudaMalloc(&sv->d_tables, bytes * 2));
                sv->d_tables_cap = bytes * 2;
            }
            CU(cudaMemcpyAsync(sv->d_tables, np.mux_tables.data(), bytes, cudaMemcpyHostToDevice, sv->stream));
        }
        // pageable uploads are staged before cudaMemcpyAsync returns, so host_ops may die here
        sv->plan_key = key;
        sv->plan_key2 = key2;
        sv->plan_ngates = n_gates;
        sv->plan_gate_copy.assign(reinterpret_cast<const unsigned char*>(gates), reinterpret_cast<const unsigned char*>(gates) + n_gates * sizeof(b2sv_gate));
        sv->plan_valid = true;
    }
    const Plan& plan = sv->plan;
    const std::vector<size_t>& seg_first = sv->plan_seg_first;
    sv->stats.gates_in += plan.gates_in;
    sv->stats.gates_live += plan.gates_live;
    {  // fold the plan's scalar into the pending one
        const double r = sv->sre * plan.scalar_re - sv->sim * plan.scalar_im;
        const double i = sv->sre * plan.scalar_im + sv->sim * plan.scalar_re;
        sv->sre = r;
        sv->sim = i;
    }
    cudaEvent_t ev_a = get_event(sv), ev_b = get_event(sv);
    CU(cudaEventRecord(ev_a, sv->stream));
    for (size_t s = 0; s < plan.segs.size(); ++s) {
        const Segment& seg = plan.segs[s];
        int rc = B2SV_OK;
        NvtxRange nvtx_seg(seg.cls == SEG_TILE ? "sweep: tile" : seg.cls == SEG_LOW6 ? "sweep: low6" : seg.cls == SEG_PERM ? "sweep: perm" : "sweep: gate1q");
        // a pending basis state is generated by the first sweep if that sweep visits every tile
        const bool lazy_init = sv->pending_basis && (seg.cls == SEG_TILE || seg.cls == SEG_LOW6) && seg.common_ctrl == 0;
        // ... and if a permutation sweep comes next, only the tile that owns the index is written at all: the
        // permutation sweep reads that tile and takes zeros for everything else
        // (tail_window: the caller promises that a permutation sweep comes next -- the distributed layer's gather)
        const bool perm_next = s + 1 < plan.segs.size() ? plan.segs[s + 1].cls == SEG_PERM : tail_window;
        const bool owner_only = lazy_init && perm_next && sv->n <= kMaxFixed && !std::getenv("B2SV_NO_WINDOW");
        if ((sv->pending_basis && !lazy_init) || (sv->window && seg.cls != SEG_PERM)) {
            rc = materialise(sv);
            if (rc) return rc;
        }
        if (seg.cls == SEG_GATE1Q) {
            for (int gi : seg.gates) {
                rc = launch_gate1(sv, plan.gates[gi], timing);
                if (rc) break;
            }
        } else if (seg.cls == SEG_TILE) {
            bool fused_ops = false;
            for (int gi : seg.gates) fused_ops |= plan.gates[gi].op == B2SV_OP_MUX || plan.gates[gi].op == B2SV_OP_BLOCK;
            rc = launch_tile(sv, seg, (const TileOp*)sv->d_ops + seg_first[s], sv->plan_seg_nops[s], cfg, tile_io, fused_ops, timing, lazy_init,
                             owner_only, sv->plan_seg_slots[s], plan.mux_tables);
            if (lazy_init) sv->pending_basis = false;
        } else if (seg.cls == SEG_LOW6) {
            rc = launch_low6(sv, seg, (const TileOp*)sv->d_ops + seg_first[s], sv->plan_seg_nops[s], timing, lazy_init, owner_only);
            if (lazy_init) sv->pending_basis = false;
        } else {
            rc = launch_perm(sv, plan, seg, timing, jit_mode);
        }
        if (rc) return rc;
        CU(cudaGetLastError());
        if (owner_only) {
            sv->window = true;
            sv->win_mask = low_mask(sv->n) & ~seg.tile_mask;
            sv->win_value = sv->basis_index & ~seg.tile_mask;
        }
    }
    CU(cudaEventRecord(ev_b, sv->stream));
    sv->apply_events.push_back({ev_a, ev_b});
    // host_ops is pageable memory: the async copy above has been staged, but kernels may still run.
    // Synchronise so that errors surface here and the caller's gate buffer can be reused.
    cudaError_t e = cudaStreamSynchronize(sv->stream);
    if (e != cudaSuccess) return fail(B2SV_ECUDA, "kernel execution failed: %s", cudaGetErrorString(e));
    return B2SV_OK;
}

int b2sv_project_norm2_at(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                          uint64_t index_base, double* out) {
    NvtxRange nvtx_range("b2sv_project_norm2_at");
    if (int rc = check(sv)) return rc;
    if (!out) return fail(B2SV_EINVAL, "out is null");
    CU(cudaSetDevice(sv->device));
    if (index_base & (sv->count - 1)) return fail(B2SV_EINVAL, "index_base must have its low %d bits clear", sv->n);
    if (sv->zero) {
        *out = 0.0;
        return B2SV_OK;
    }
    if (int rc = materialise(sv)) return rc;
    int count_bits = 0;
    bool full_cover = false;
    uint64_t head_zeros = 0;
    bool pure = false;
    if (int rc = upload_subprog(sv, prog, n_instr, entry, target_qubits, &count_bits, 64, &full_cover, &head_zeros, &pure)) return rc;
    if (target_qubits < 64 && (index_base >> target_qubits) != 0) {  // this shard holds ancilla != 0 only
        *out = 0.0;
        return B2SV_OK;
    }
    if (count_bits > sv->n) count_bits = sv->n;
    head_zeros &= low_mask(count_bits);
    static const bool no_scan_kernel = std::getenv("B2SV_NO_NORM_SCAN") != nullptr;
    if (no_scan_kernel) pure = false;
    const bool grouped = !pure && head_zeros == 0 && low_free_bits(prog, entry) >= kProjectGroupBits && (1ull << count_bits) >= (uint64_t)kProjectWarpSpan;
    ExpandArgs scan{};
    if (!grouped) fill_expand(scan, head_zeros);
    // grouped: the whole range; else only the indices whose always-zero bits are clear
    const uint64_t cnt = 1ull << (count_bits - scan.m);
    const int grid = grid_for(sv, grouped ? cnt >> kProjectGroupBits : (pure ? (cnt + 7) / 8 : cnt));
    int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        // algorithmic bytes: the member amplitudes (what has to be read), never more than the scanned range
        // (per-rank specialised programs do not mention their free bits: book the scanned range for those)
        const double members = (sv->timing_on && full_cover) ? std::min((double)cnt, subprog_members(prog, n_instr, entry)) : (double)cnt;
        LaunchTimer lt(sv, sv->timing_on, CLS_PROJECT, (double)sv->amp_bytes * members);
        if (pure)  // branch-free subspace: a strided streaming reduction, no membership walk
            k_norm2_scan<A><<<grid, kThreads, 0, sv->stream>>>((const A*)sv->amps, cnt, sv->d_partial, scan);
        else if (grouped)
            k_project_norm2<A, true><<<grid, kThreads, 0, sv->stream>>>((const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1, index_base,
                                                                       sv->d_partial, scan);
        else
            k_project_norm2<A, false><<<grid, kThreads, 0, sv->stream>>>((const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1, index_base,
                                                                        sv->d_partial, scan);
        k_reduce_partials<<<1, kThreads, 0, sv->stream>>>(sv->d_partial, grid, sv->d_partial + sv->n_partial);
        sv->stats.launches++;  // the reduction is a launch too
        sv->stats.launches_by_class[CLS_PROJECT]++;
        return B2SV_OK;
    });
    if (rc) return rc;
    CU(cudaGetLastError());
    double v = 0;
    CU(cudaMemcpyAsync(&v, sv->d_partial + sv->n_partial, sizeof(double), cudaMemcpyDeviceToHost, sv->stream));
    CU(cudaStreamSynchronize(sv->stream));
    *out = v * (sv->sre * sv->sre + sv->sim * sv->sim);
    return B2SV_OK;
}

int b2sv_project_norm2(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits, double* out) {
    if (int rc = check(sv)) return rc;
    if (target_qubits > sv->n) return fail(B2SV_EINVAL, "target_qubits %d outside [0,%d]", target_qubits, sv->n);
    return b2sv_project_norm2_at(sv, prog, n_instr, entry, target_qubits, 0, out);
}

int b2sv_project_gather(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                        double scale_re, double scale_im, void* host_out, uint64_t dim) {
    NvtxRange nvtx_range("b2sv_project_gather");
    if (int rc = check(sv)) return rc;
    if (!host_out && dim > 0) return fail(B2SV_EINVAL, "host_out is null");
    CU(cudaSetDevice(sv->device));
    if (int rc = materialise(sv)) return rc;
    int count_bits = 0;
    uint64_t head_zeros = 0;
    if (int rc = upload_subprog(sv, prog, n_instr, entry, target_qubits, &count_bits, -1, nullptr, &head_zeros)) return rc;
    if (dim == 0) return B2SV_OK;
    const bool grouped = head_zeros == 0 && low_free_bits(prog, entry) >= kProjectGroupBits && (1ull << count_bits) >= (uint64_t)kProjectWarpSpan;
    ExpandArgs scan{};
    if (!grouped) fill_expand(scan, head_zeros);
    const uint64_t cnt = 1ull << (count_bits - scan.m);
    if (int rc = ensure_gather(sv, dim * 16)) return rc;
    double2* d_out = (double2*)sv->d_gather;
    cudaMemsetAsync(d_out, 0, dim * 16, sv->stream);
    // the pending global scalar rides along in the scale factor
    const double2 s = make_double2(scale_re * sv->sre - scale_im * sv->sim, scale_re * sv->sim + scale_im * sv->sre);
    dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, sv->timing_on, CLS_PROJECT, (double)sv->amp_bytes * std::min((double)cnt, (double)dim) + 16.0 * (double)dim);
        if (grouped)
            k_project_gather<A, true><<<grid_for(sv, cnt >> kProjectGroupBits), kThreads, 0, sv->stream>>>(
                (const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1, s, d_out, dim, prog[entry].weight, scan);
        else
            k_project_gather<A, false><<<grid_for(sv, cnt), kThreads, 0, sv->stream>>>((const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1,
                                                                                     s, d_out, dim, 0, scan);
        return B2SV_OK;
    });
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d_out, dim * 16, cudaMemcpyDeviceToHost, sv->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sv->stream);
    if (e != cudaSuccess) return fail(B2SV_ECUDA, "project_gather failed: %s", cudaGetErrorString(e));
    return B2SV_OK;
}

int b2sv_project_gather_sparse(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                               double scale_re, double scale_im, uint64_t dim, uint64_t cap, uint64_t* idx_out, void* val_out,
                               uint64_t* nnz_out) {
    NvtxRange nvtx_range("b2sv_project_gather_sparse");
    if (int rc = check(sv)) return rc;
    if (!nnz_out || (cap > 0 && (!idx_out || !val_out))) return fail(B2SV_EINVAL, "null output");
    *nnz_out = 0;
    CU(cudaSetDevice(sv->device));
    if (int rc = materialise(sv)) return rc;
    int count_bits = 0;
    uint64_t head_zeros = 0;
    if (int rc = upload_subprog(sv, prog, n_instr, entry, target_qubits, &count_bits, -1, nullptr, &head_zeros)) return rc;
    if (dim == 0) return B2SV_OK;
    if (cap == 0) {
        *nnz_out = 1;  // ">= cap": nothing can be listed
        return B2SV_OK;
    }
    const bool grouped = head_zeros == 0 && low_free_bits(prog, entry) >= kProjectGroupBits && (1ull << count_bits) >= (uint64_t)kProjectWarpSpan;
    ExpandArgs scan{};
    if (!grouped) fill_expand(scan, head_zeros);
    const uint64_t cnt = 1ull << (count_bits - scan.m);
    // device layout: values (cap * 16 B) | indices (cap * 8 B) | count (8 B)
    if (int rc = ensure_gather(sv, cap * 24 + 8)) return rc;
    double2* d_val = (double2*)sv->d_gather;
    SparseOut sp;
    sp.idx = reinterpret_cast<unsigned long long*>(d_val + cap);
    sp.count = sp.idx + cap;
    sp.cap = cap;
    CU(cudaMemsetAsync(sp.count, 0, 8, sv->stream));
    const double2 s = make_double2(scale_re * sv->sre - scale_im * sv->sim, scale_re * sv->sim + scale_im * sv->sre);
    dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, sv->timing_on, CLS_PROJECT, (double)sv->amp_bytes * std::min((double)cnt, (double)dim));
        if (grouped)
            k_project_gather<A, true, true><<<grid_for(sv, cnt >> kProjectGroupBits), kThreads, 0, sv->stream>>>(
                (const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1, s, d_val, dim, prog[entry].weight, scan, sp);
        else
            k_project_gather<A, false, true><<<grid_for(sv, cnt), kThreads, 0, sv->stream>>>((const A*)sv->amps, sv->d_sub, cnt, entry,
                                                                                           n_instr + 1, s, d_val, dim, 0, scan, sp);
        return B2SV_OK;
    });
    CU(cudaGetLastError());
    unsigned long long n = 0;
    CU(cudaMemcpyAsync(&n, sp.count, 8, cudaMemcpyDeviceToHost, sv->stream));
    CU(cudaStreamSynchronize(sv->stream));
    *nnz_out = n;
    if (n == 0 || n >= cap) return B2SV_OK;
    CU(cudaMemcpyAsync(idx_out, sp.idx, n * 8, cudaMemcpyDeviceToHost, sv->stream));
    CU(cudaMemcpyAsync(val_out, d_val, n * 16, cudaMemcpyDeviceToHost, sv->stream));
    CU(cudaStreamSynchronize(sv->stream));
    return B2SV_OK;
}

int b2sv_project_gather_async(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                              double scale_re, double scale_im, void* host_out, uint64_t dim, int slot) {
    NvtxRange nvtx_range("b2sv_project_gather_async");
    if (int rc = check(sv)) return rc;
    if (slot < 0 || slot > 1) return fail(B2SV_EINVAL, "slot %d outside [0,1]", slot);
    if (!host_out && dim > 0) return fail(B2SV_EINVAL, "host_out is null");
    CU(cudaSetDevice(sv->device));
    if (!sv->copy_stream) {
        CU(cudaStreamCreateWithFlags(&sv->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CU(cudaEventCreateWithFlags(&sv->ev_col_kernel[i], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&sv->ev_col_copy[i], cudaEventDisableTiming));
        }
    }
    if (int rc = materialise(sv)) return rc;
    int count_bits = 0;
    uint64_t head_zeros = 0;
    if (int rc = upload_subprog(sv, prog, n_instr, entry, target_qubits, &count_bits, -1, nullptr, &head_zeros)) return rc;
    if (dim == 0) return B2SV_OK;
    if (sv->col_pending[slot]) {
        if (dim * 16 > sv->d_col_cap[slot]) CU(cudaEventSynchronize(sv->ev_col_copy[slot]));  // about to free the buffer
        else CU(cudaStreamWaitEvent(sv->stream, sv->ev_col_copy[slot], 0));                      // about to overwrite it
        sv->col_pending[slot] = false;
    }
    if (dim * 16 > sv->d_col_cap[slot]) {
        CU(cudaStreamSynchronize(sv->stream));
        cudaFree(sv->d_col[slot]);
        sv->d_col[slot] = nullptr;
        sv->d_col_cap[slot] = 0;
        CU(cudaMalloc(&sv->d_col[slot], dim * 16));
        sv->d_col_cap[slot] = dim * 16;
    }
    const bool grouped = head_zeros == 0 && low_free_bits(prog, entry) >= kProjectGroupBits && (1ull << count_bits) >= (uint64_t)kProjectWarpSpan;
    ExpandArgs scan{};
    if (!grouped) fill_expand(scan, head_zeros);
    const uint64_t cnt = 1ull << (count_bits - scan.m);
    double2* d_out = (double2*)sv->d_col[slot];
    CU(cudaMemsetAsync(d_out, 0, dim * 16, sv->stream));
    const double2 s = make_double2(scale_re * sv->sre - scale_im * sv->sim, scale_re * sv->sim + scale_im * sv->sre);
    dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, sv->timing_on, CLS_PROJECT, (double)sv->amp_bytes * std::min((double)cnt, (double)dim) + 16.0 * (double)dim);
        if (grouped)
            k_project_gather<A, true><<<grid_for(sv, cnt >> kProjectGroupBits), kThreads, 0, sv->stream>>>(
                (const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1, s, d_out, dim, prog[entry].weight, scan);
        else
            k_project_gather<A, false><<<grid_for(sv, cnt), kThreads, 0, sv->stream>>>((const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1,
                                                                                     s, d_out, dim, 0, scan);
        return B2SV_OK;
    });
    CU(cudaGetLastError());
    CU(cudaEventRecord(sv->ev_col_kernel[slot], sv->stream));
    CU(cudaStreamWaitEvent(sv->copy_stream, sv->ev_col_kernel[slot], 0));
    CU(cudaMemcpyAsync(host_out, d_out, dim * 16, cudaMemcpyDeviceToHost, sv->copy_stream));
    CU(cudaEventRecord(sv->ev_col_copy[slot], sv->copy_stream));
    sv->col_pending[slot] = true;
    return B2SV_OK;
}

int b2sv_gather_wait(b2sv_t* sv, int slot) {
    if (int rc = check(sv)) return rc;
    if (slot < 0 || slot > 1) return fail(B2SV_EINVAL, "slot %d outside [0,1]", slot);
    if (!sv->col_pending[slot]) return B2SV_OK;
    CU(cudaSetDevice(sv->device));
    CU(cudaEventSynchronize(sv->ev_col_copy[slot]));
    sv->col_pending[slot] = false;
    return B2SV_OK;
}

int b2sv_project_gather_device(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                               double scale_re, double scale_im, void** dev_out, uint64_t dim) {
    NvtxRange nvtx_range("b2sv_project_gather_device");
    if (int rc = check(sv)) return rc;
    if (!dev_out) return fail(B2SV_EINVAL, "dev_out is null");
    *dev_out = nullptr;
    CU(cudaSetDevice(sv->device));
    if (int rc = materialise(sv)) return rc;
    int count_bits = 0;
    uint64_t head_zeros = 0;
    if (int rc = upload_subprog(sv, prog, n_instr, entry, target_qubits, &count_bits, -1, nullptr, &head_zeros)) return rc;
    if (int rc = ensure_gather(sv, (dim ? dim : 1) * 16)) return rc;
    double2* d_out = (double2*)sv->d_gather;
    *dev_out = d_out;
    if (dim == 0) return B2SV_OK;
    const bool grouped = head_zeros == 0 && low_free_bits(prog, entry) >= kProjectGroupBits && (1ull << count_bits) >= (uint64_t)kProjectWarpSpan;
    ExpandArgs scan{};
    if (!grouped) fill_expand(scan, head_zeros);
    const uint64_t cnt = 1ull << (count_bits - scan.m);
    CU(cudaMemsetAsync(d_out, 0, dim * 16, sv->stream));
    const double2 s = make_double2(scale_re * sv->sre - scale_im * sv->sim, scale_re * sv->sim + scale_im * sv->sre);
    dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, sv->timing_on, CLS_PROJECT, (double)sv->amp_bytes * std::min((double)cnt, (double)dim) + 16.0 * (double)dim);
        if (grouped)
            k_project_gather<A, true><<<grid_for(sv, cnt >> kProjectGroupBits), kThreads, 0, sv->stream>>>(
                (const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1, s, d_out, dim, prog[entry].weight, scan);
        else
            k_project_gather<A, false><<<grid_for(sv, cnt), kThreads, 0, sv->stream>>>((const A*)sv->amps, sv->d_sub, cnt, entry, n_instr + 1,
                                                                                     s, d_out, dim, 0, scan);
        return B2SV_OK;
    });
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(sv->stream));  // the caller reads the buffer from its own stream
    return B2SV_OK;
}

int b2sv_upload_device(b2sv_t* sv, const void* dev_amps_c128) {
    NvtxRange nvtx_range("b2sv_upload_device");
    if (int rc = check(sv)) return rc;
    if (!dev_amps_c128) return fail(B2SV_EINVAL, "device buffer is null");
    CU(cudaSetDevice(sv->device));
    sv->sre = 1.0;
    sv->sim = 0.0;
    sv->pending_basis = false;
    sv->window = false;
    sv->zero = false;
    if (sv->dtype == B2SV_C128) {
        CU(cudaMemcpyAsync(sv->amps, dev_amps_c128, sv->count * 16, cudaMemcpyDeviceToDevice, sv->stream));
    } else {
        sv->stats.launches_by_class[CLS_INIT]++;
        k_convert<float2, double2><<<grid_for(sv, sv->count), kThreads, 0, sv->stream>>>((float2*)sv->amps, (const double2*)dev_amps_c128, sv->count);
        CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(sv->stream));  // the source belongs to the caller's stream
    return B2SV_OK;
}

int b2sv_embed(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits, const void* host_in,
               uint64_t dim) {
    NvtxRange nvtx_range("b2sv_embed");
    if (int rc = check(sv)) return rc;
    if (!host_in && dim > 0) return fail(B2SV_EINVAL, "host_in is null");
    CU(cudaSetDevice(sv->device));
    int count_bits = 0;
    if (int rc = upload_subprog(sv, prog, n_instr, entry, target_qubits, &count_bits)) return rc;
    const uint64_t cnt = 1ull << count_bits;
    if (int rc = ensure_gather(sv, (dim ? dim : 1) * 16)) return rc;
    double2* d_in = (double2*)sv->d_gather;
    if (dim) CU(cudaMemcpyAsync(d_in, host_in, dim * 16, cudaMemcpyHostToDevice, sv->stream));
    // the register is overwritten: nothing pending survives
    sv->pending_basis = false;
    sv->window = false;
    sv->zero = false;
    sv->sre = 1.0;
    sv->sim = 0.0;
    dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, sv->timing_on, CLS_INIT, (double)sv->amp_bytes * (double)sv->count + 16.0 * (double)dim);
        k_embed<A><<<grid_for(sv, sv->count), kThreads, 0, sv->stream>>>((A*)sv->amps, sv->d_sub, sv->count, cnt, entry, n_instr + 1, d_in, dim);
        return B2SV_OK;
    });
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(sv->stream);  // host_in is borrowed for the call only
    if (e != cudaSuccess) return fail(B2SV_ECUDA, "embed failed: %s", cudaGetErrorString(e));
    return B2SV_OK;
}

int b2sv_enumerate_basis(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                         int64_t* host_out, uint64_t dim) {
    NvtxRange nvtx_range("b2sv_enumerate_basis");
    if (int rc = check(sv)) return rc;
    if (!host_out && dim > 0) return fail(B2SV_EINVAL, "host_out is null");
    CU(cudaSetDevice(sv->device));
    int count_bits = 0;
    // listing member indices touches no amplitude: any handle will do, whatever its register size
    if (int rc = upload_subprog(sv, prog, n_instr, entry, target_qubits, &count_bits, 64)) return rc;
    if (dim == 0) return B2SV_OK;
    const uint64_t cnt = 1ull << count_bits;
    if (int rc = ensure_gather(sv, dim * 8)) return rc;
    long long* d_out = (long long*)sv->d_gather;
    cudaMemsetAsync(d_out, 0xFF, dim * 8, sv->stream);
    sv->stats.launches_by_class[CLS_PROJECT]++;
    k_enumerate_basis<<<grid_for(sv, cnt), kThreads, 0, sv->stream>>>(sv->d_sub, cnt, entry, n_instr + 1, d_out, dim);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d_out, dim * 8, cudaMemcpyDeviceToHost, sv->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sv->stream);
    if (e != cudaSuccess) return fail(B2SV_ECUDA, "enumerate_basis failed: %s", cudaGetErrorString(e));
    return B2SV_OK;
}

int b2sv_device_ptr(b2sv_t* sv, void** amps, uint64_t* bytes) {
    if (int rc = check(sv)) return rc;
    CU(cudaSetDevice(sv->device));
    if (int rc = materialise(sv)) return rc;
    if (int rc = flush_scalar(sv, false)) return rc;
    CU(cudaStreamSynchronize(sv->stream));
    if (amps) *amps = sv->amps;
    if (bytes) *bytes = sv->count * sv->amp_bytes;
    return B2SV_OK;
}

int b2sv_sync(b2sv_t* sv) {
    if (int rc = check(sv)) return rc;
    CU(cudaSetDevice(sv->device));
    CU(cudaStreamSynchronize(sv->stream));
    return B2SV_OK;
}

int b2sv_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(B2SV_EINVAL, "out is null");
    *out = nullptr;
    if (bytes == 0) return fail(B2SV_EINVAL, "zero-size allocation");
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return fail(e == cudaErrorMemoryAllocation ? B2SV_ENOMEM : B2SV_ENODEV, "cudaHostAlloc(%zu) failed: %s", bytes,
                    cudaGetErrorString(e));
    }
    return B2SV_OK;
}

int b2sv_host_free(void* ptr) {
    if (!ptr) return B2SV_OK;
    cudaError_t e = cudaFreeHost(ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(B2SV_ECUDA, "cudaFreeHost failed: %s", cudaGetErrorString(e));
    }
    return B2SV_OK;
}

int b2sv_mark(b2sv_t* sv, int slot) {
    if (int rc = check(sv)) return rc;
    if (slot < 0 || slot >= 8) return fail(B2SV_EINVAL, "mark slot %d outside [0,8)", slot);
    CU(cudaSetDevice(sv->device));
    if (!sv->marks[slot]) CU(cudaEventCreate(&sv->marks[slot]));
    CU(cudaEventRecord(sv->marks[slot], sv->stream));
    return B2SV_OK;
}

int b2sv_elapsed_ms(b2sv_t* sv, int slot_a, int slot_b, double* ms) {
    if (int rc = check(sv)) return rc;
    if (!ms) return fail(B2SV_EINVAL, "ms is null");
    if (slot_a < 0 || slot_a >= 8 || slot_b < 0 || slot_b >= 8 || !sv->marks[slot_a] || !sv->marks[slot_b])
        return fail(B2SV_ESTATE, "marks %d/%d were not recorded", slot_a, slot_b);
    CU(cudaSetDevice(sv->device));
    CU(cudaEventSynchronize(sv->marks[slot_b]));
    float f = 0;
    CU(cudaEventElapsedTime(&f, sv->marks[slot_a], sv->marks[slot_b]));
    *ms = f;
    return B2SV_OK;
}

static int swap_pack_impl(b2sv_t* sv, int q, int keep_bit, void* staging, bool pack, uint64_t first, uint64_t count) {
    if (int rc = check(sv)) return rc;
    if (q < 0 || q >= sv->n) return fail(B2SV_EINVAL, "qubit %d outside the shard", q);
    if (!staging) return fail(B2SV_EINVAL, "staging is null");
    const uint64_t half = sv->count >> 1;
    if (first > half || count > half - first) return fail(B2SV_EINVAL, "range [%llu, +%llu) outside the %llu packed amplitudes",
                                                          (unsigned long long)first, (unsigned long long)count, (unsigned long long)half);
    CU(cudaSetDevice(sv->device));
    if (int rc = materialise(sv)) return rc;
    if (count == 0) return B2SV_OK;
    const uint64_t sel = keep_bit ? 0ull : (1ull << q);  // move the half whose bit differs from keep_bit
    // The shard's lazy global scalar stays pending for the half that does not move: outgoing amplitudes
    // are multiplied by it on the way out, incoming ones by its conjugate (|s| = 1) on the way in.
    const bool scale = !(sv->sre == 1.0 && sv->sim == 0.0);
    const double2 sc = pack ? make_double2(sv->sre, sv->sim) : make_double2(sv->sre, -sv->sim);
    int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, false, CLS_SWAP, 2.0 * sv->amp_bytes * (double)count);
        const int grid = grid_for(sv, (count + 3) / 4, 16);
        A* amps = (A*)sv->amps;
        A* st = (A*)staging;
        const uint64_t end = first + count;
        if (pack && scale) k_swap_pack<A, true, true><<<grid, kThreads, 0, sv->stream>>>(amps, st, first, end, q, sel, sc);
        else if (pack) k_swap_pack<A, true, false><<<grid, kThreads, 0, sv->stream>>>(amps, st, first, end, q, sel, sc);
        else if (scale) k_swap_pack<A, false, true><<<grid, kThreads, 0, sv->stream>>>(amps, st, first, end, q, sel, sc);
        else k_swap_pack<A, false, false><<<grid, kThreads, 0, sv->stream>>>(amps, st, first, end, q, sel, sc);
        return B2SV_OK;
    });
    if (rc) return rc;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(sv->stream));
    return B2SV_OK;
}

int b2sv_swap_pack(b2sv_t* sv, int q, int keep_bit, void* staging_dev) {
    if (int rc = check(sv)) return rc;
    return swap_pack_impl(sv, q, keep_bit, staging_dev, true, 0, sv->count >> 1);
}
int b2sv_swap_unpack(b2sv_t* sv, int q, int keep_bit, const void* staging_dev) {
    if (int rc = check(sv)) return rc;
    return swap_pack_impl(sv, q, keep_bit, const_cast<void*>(staging_dev), false, 0, sv->count >> 1);
}
int b2sv_swap_pack_range(b2sv_t* sv, int q, int keep_bit, void* staging_dev, uint64_t first, uint64_t count) {
    return swap_pack_impl(sv, q, keep_bit, staging_dev, true, first, count);
}
int b2sv_swap_unpack_range(b2sv_t* sv, int q, int keep_bit, const void* staging_dev, uint64_t first, uint64_t count) {
    return swap_pack_impl(sv, q, keep_bit, const_cast<void*>(staging_dev), false, first, count);
}


// ---- distributed execution: peer-mapped shards + one fused gather kernel per exchange ----------------

namespace {
struct IpcBlob {  // B2SV_IPC_BYTES bytes, plain data
    cudaIpcMemHandle_t mem[2];
    cudaIpcEventHandle_t ev[2];
    int64_t pid;
    void* raw_buf[2];
    cudaEvent_t raw_ev[2];
    int32_t device, n, dtype, pad;
};
static_assert(sizeof(IpcBlob) <= B2SV_IPC_BYTES, "IPC blob layout");
}  // namespace

int b2sv_ipc_export(b2sv_t* sv, void* blob_out) {
    if (int rc = check(sv)) return rc;
    if (!blob_out) return fail(B2SV_EINVAL, "blob_out is null");
    CU(cudaSetDevice(sv->device));
    if (ensure_scratch(sv) != B2SV_OK) return fail(B2SV_ENOMEM, "no memory for the second state buffer (distributed sweeps are out of place)");
    if (!sv->own_buf[0]) {
        sv->own_buf[0] = sv->amps;
        sv->own_buf[1] = sv->scratch;
    }
    if (!sv->ev_ready) {
        CU(cudaEventCreateWithFlags(&sv->ev_ready, cudaEventDisableTiming | cudaEventInterprocess));
        CU(cudaEventCreateWithFlags(&sv->ev_done, cudaEventDisableTiming | cudaEventInterprocess));
    }
    IpcBlob b;
    std::memset(&b, 0, sizeof(b));
    CU(cudaIpcGetMemHandle(&b.mem[0], sv->own_buf[0]));
    CU(cudaIpcGetMemHandle(&b.mem[1], sv->own_buf[1]));
    CU(cudaIpcGetEventHandle(&b.ev[0], sv->ev_ready));
    CU(cudaIpcGetEventHandle(&b.ev[1], sv->ev_done));
    b.pid = (int64_t)getpid();
    b.raw_buf[0] = sv->own_buf[0];
    b.raw_buf[1] = sv->own_buf[1];
    b.raw_ev[0] = sv->ev_ready;
    b.raw_ev[1] = sv->ev_done;
    b.device = sv->device;
    b.n = sv->n;
    b.dtype = sv->dtype;
    std::memset(blob_out, 0, B2SV_IPC_BYTES);
    std::memcpy(blob_out, &b, sizeof(b));
    return B2SV_OK;
}

int b2sv_ipc_detach(b2sv_t* sv) {
    if (int rc = check(sv)) return rc;
    if (sv->peers.empty()) return B2SV_OK;
    cudaSetDevice(sv->device);
    cudaStreamSynchronize(sv->stream);
    for (auto& p : sv->peers) {
        if (!p.ipc) continue;
        for (int i = 0; i < 2; ++i)
            if (p.buf[i]) cudaIpcCloseMemHandle(p.buf[i]);
        if (p.ready) cudaEventDestroy(p.ready);
        if (p.done) cudaEventDestroy(p.done);
    }
    cudaGetLastError();
    sv->peers.clear();
    sv->dist_rank = -1;
    sv->dist_world = 0;
    return B2SV_OK;
}

int b2sv_ipc_attach(b2sv_t* sv, int rank, int world, const void* blobs) {
    if (int rc = check(sv)) return rc;
    if (!blobs) return fail(B2SV_EINVAL, "blobs is null");
    if (world < 2 || world > kDistMaxRanks || (world & (world - 1))) return fail(B2SV_EINVAL, "world size %d must be a power of two in [2,%d]", world, kDistMaxRanks);
    if (rank < 0 || rank >= world) return fail(B2SV_EINVAL, "rank %d outside [0,%d)", rank, world);
    if (!sv->own_buf[0] || !sv->ev_ready) return fail(B2SV_ESTATE, "b2sv_ipc_export must be called first");
    CU(cudaSetDevice(sv->device));
    b2sv_ipc_detach(sv);
    sv->peers.assign(world, b2sv::Peer());
    const int64_t me = (int64_t)getpid();
    for (int r = 0; r < world; ++r) {
        IpcBlob b;
        std::memcpy(&b, (const char*)blobs + (size_t)r * B2SV_IPC_BYTES, sizeof(b));
        if (b.n != sv->n || b.dtype != sv->dtype) {
            b2sv_ipc_detach(sv);
            return fail(B2SV_EINVAL, "rank %d holds a %d-qubit dtype-%d shard, this rank a %d-qubit dtype-%d one", r, b.n, b.dtype, sv->n, sv->dtype);
        }
        b2sv::Peer& p = sv->peers[r];
        if (r == rank) {
            p.buf[0] = sv->own_buf[0];
            p.buf[1] = sv->own_buf[1];
            continue;
        }
        if (b.pid == me) {  // another handle of this process (threads driving several GPUs): plain pointers
            if (b.device != sv->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    b2sv_ipc_detach(sv);
                    return fail(B2SV_ECUDA, "no peer access from device %d to device %d: %s", sv->device, b.device, cudaGetErrorString(e));
                }
                cudaGetLastError();
            }
            p.buf[0] = b.raw_buf[0];
            p.buf[1] = b.raw_buf[1];
            p.ready = b.raw_ev[0];
            p.done = b.raw_ev[1];
            continue;
        }
        p.ipc = true;
        cudaError_t e = cudaIpcOpenMemHandle(&p.buf[0], b.mem[0], cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenMemHandle(&p.buf[1], b.mem[1], cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) e = cudaIpcOpenEventHandle(&p.ready, b.ev[0]);
        if (e == cudaSuccess) e = cudaIpcOpenEventHandle(&p.done, b.ev[1]);
        if (e != cudaSuccess) {
            cudaGetLastError();
            b2sv_ipc_detach(sv);
            return fail(B2SV_ECUDA, "mapping the shard of rank %d failed: %s", r, cudaGetErrorString(e));
        }
    }
    sv->dist_rank = rank;
    sv->dist_world = world;
    return B2SV_OK;
}

int b2sv_dist_ready(b2sv_t* sv, uint64_t* state_out) {
    NvtxRange nvtx_range("b2sv_dist_ready");
    if (int rc = check(sv)) return rc;
    if (!state_out) return fail(B2SV_EINVAL, "state_out is null");
    if (sv->dist_world == 0) return fail(B2SV_ESTATE, "not attached (b2sv_ipc_attach)");
    CU(cudaSetDevice(sv->device));
    if (sv->pending_basis) {
        // a basis state nobody has touched yet: one amplitude, everything else implied zeros
        sv->pending_basis = false;
        const uint64_t index = sv->basis_index;
        dispatch_dtype(sv, [&](auto* tag) -> int {
            using A = std::remove_pointer_t<decltype(tag)>;
            sv->stats.launches_by_class[CLS_INIT]++;
            k_set_one<A><<<1, 1, 0, sv->stream>>>((A*)sv->amps, index);
            return B2SV_OK;
        });
        CU(cudaGetLastError());
        sv->window = true;
        sv->win_mask = low_mask(sv->n);
        sv->win_value = index;
    }
    std::memset(state_out, 0, 8 * B2SV_DIST_STATE_WORDS);
    state_out[0] = (sv->amps == sv->own_buf[1] ? 1ull : 0ull) | (sv->window ? 0x100ull : 0ull) | (sv->zero ? 0x200ull : 0ull);
    if (sv->zero) {
        state_out[1] = 1;  // (i & 1) == 2 never holds: an empty window
        state_out[2] = 2;
    } else if (sv->window) {
        state_out[1] = sv->win_mask;
        state_out[2] = sv->win_value;
    }
    std::memcpy(&state_out[3], &sv->sre, 8);
    std::memcpy(&state_out[4], &sv->sim, 8);
    CU(cudaEventRecord(sv->ev_ready, sv->stream));
    return B2SV_OK;
}

int b2sv_dist_perm(b2sv_t* sv, const b2sv_gate* gates, size_t n_gates, int n_total, uint32_t flip_in, uint32_t flip_out,
                   const uint64_t* states, const b2sv_plan_opts* opts) {
    NvtxRange nvtx_range("b2sv_dist_perm");
    if (int rc = check(sv)) return rc;
    if (sv->dist_world == 0) return fail(B2SV_ESTATE, "not attached (b2sv_ipc_attach)");
    if (n_gates > 0 && !gates) return fail(B2SV_EINVAL, "gates is null");
    if (!states) return fail(B2SV_EINVAL, "states is null");
    const int g = n_total - sv->n;
    if (g < 1 || (1 << g) != sv->dist_world) return fail(B2SV_EINVAL, "n_total %d does not match %d local qubits on %d ranks", n_total, sv->n, sv->dist_world);
    if (((flip_in | flip_out) >> g) != 0) return fail(B2SV_EINVAL, "flip masks wider than the %d global bits", g);
    CU(cudaSetDevice(sv->device));
    const bool timing = opts && opts->timing;
    const int jit_mode = opts ? opts->jit : 0;
    std::vector<PGate> pg;
    pg.reserve(n_gates);
    {
        double sre = 1, sim = 0;
        const uint64_t all = low_mask(n_total);
        for (size_t j = 0; j < n_gates; ++j) {
            const b2sv_gate& gt = gates[j];
            if (gt.op > B2SV_OP_PHASE) return fail(B2SV_EINVAL, "gate %zu: unknown op %d", j, (int)gt.op);
            PGate p;
            if (!normalise_gate(gt, (int)j, p, sre, sim)) continue;
            if (!p.is_perm()) return fail(B2SV_EINVAL, "gate %zu is not an X / SWAP gate: a distributed sweep only moves amplitudes", j);
            if ((p.qmask & ~all) != 0) return fail(B2SV_EINVAL, "gate %zu: qubit outside the %d-qubit register", j, n_total);
            if ((p.ctrl & p.tmask) != 0 || (p.op == B2SV_OP_SWAP && p.t0 == p.t1)) return fail(B2SV_EINVAL, "gate %zu: bad targets", j);
            if (!pg.empty() && same_perm_gate(pg.back(), p)) {
                pg.pop_back();
                continue;
            }
            pg.push_back(p);
        }
        if (!(sre == 1.0 && sim == 0.0)) return fail(B2SV_EINVAL, "a distributed sweep cannot carry a global phase");
    }
    DistPermArgs a;
    std::memset(&a, 0, sizeof(a));
    bool any_scale = false;
    for (int r = 0; r < sv->dist_world; ++r) {
        const uint64_t* st = states + (size_t)r * B2SV_DIST_STATE_WORDS;
        a.src[r] = sv->peers[r].buf[st[0] & 1ull];
        a.wmask[r] = st[1];
        a.wval[r] = st[2];
        std::memcpy(&a.sre[r], &st[3], 8);
        std::memcpy(&a.sim[r], &st[4], 8);
        any_scale |= !(a.sre[r] == 1.0 && a.sim[r] == 0.0);
    }
    if (a.src[sv->dist_rank] != sv->amps) return fail(B2SV_ESTATE, "stale state words: call b2sv_dist_ready right before b2sv_dist_perm");
    a.out = sv->amps == sv->own_buf[0] ? sv->own_buf[1] : sv->own_buf[0];
    a.rank_logical = (unsigned)sv->dist_rank ^ flip_out;
    a.flip_in = flip_in;
    a.has_scale = any_scale ? 1 : 0;
    // every peer's shard must be final before it is read: wait (on the device) for the events the peers recorded
    // in b2sv_dist_ready -- the caller's host-side exchange of the state words guarantees they have been recorded
    for (int r = 0; r < sv->dist_world; ++r)
        if (r != sv->dist_rank) CU(cudaStreamWaitEvent(sv->stream, sv->peers[r].ready, 0));
    // every source shard is a small live window (or holds nothing at all): zero-fill + forward scatter (k_perm_scatter)
    bool scatter = !std::getenv("B2SV_NO_SCATTER");
    uint64_t scatter_max = 0;
    for (int r = 0; r < sv->dist_world && scatter; ++r) {
        if ((a.wval[r] & ~a.wmask[r]) != 0) continue;  // empty window
        const int bits = popcount64(~a.wmask[r] & low_mask(sv->n));
        scatter = bits <= kScatterMaxBits;
        scatter_max = std::max<uint64_t>(scatter_max, 1ull << bits);
    }
    cudaKernel_t k = nullptr;
    if (scatter) {
        if (int rc = upload_dist_gates(sv, pg)) return rc;
        int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
            using A = std::remove_pointer_t<decltype(tag)>;
            LaunchTimer lt(sv, timing, CLS_DIST, 1.0 * sv->amp_bytes * (double)sv->count);
            if (int frc = fill_zero(sv, a.out, sv->count * sv->amp_bytes)) return frc;
            if (scatter_max)
                k_perm_scatter<A><<<dim3((unsigned)((scatter_max + kThreads - 1) / kThreads), (unsigned)sv->dist_world), kThreads, 0, sv->stream>>>(
                    a, (const DistGate*)sv->d_dist_gates, (int)pg.size(), sv->n);
            return B2SV_OK;
        });
        if (rc) return rc;
        CU(cudaGetLastError());
    } else if (sv->n >= 13 && jit_mode != 1 && (k = jit_lookup(sv, pg, 2, g)) != nullptr) {
        const uint64_t blocks = sv->count / (uint64_t)(kPermThreads * 32);
        a.block_rot = (unsigned)((blocks / (uint64_t)sv->dist_world) * (uint64_t)sv->dist_rank);
        void* args[] = {(void*)&a};
        LaunchTimer lt(sv, timing, CLS_DIST, 2.0 * sv->amp_bytes * (double)sv->count);
        CU(cudaLaunchKernel((const void*)k, dim3((unsigned)blocks), dim3(kPermThreads), args, 0, sv->stream));
    } else {
        // interpreter: shards too small for the bit-sliced layout, or no NVRTC
        if (int rc = upload_dist_gates(sv, pg)) return rc;
        int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
            using A = std::remove_pointer_t<decltype(tag)>;
            LaunchTimer lt(sv, timing, CLS_DIST, 2.0 * sv->amp_bytes * (double)sv->count);
            k_dist_perm_generic<A><<<grid_for(sv, sv->count), kThreads, 0, sv->stream>>>(a, (const DistGate*)sv->d_dist_gates, (int)pg.size(), sv->n,
                                                                                      sv->count);
            return B2SV_OK;
        });
        if (rc) return rc;
        CU(cudaGetLastError());
    }
    sv->amps = a.out;
    sv->scratch = sv->amps == sv->own_buf[0] ? sv->own_buf[1] : sv->own_buf[0];
    sv->window = false;
    sv->zero = false;
    sv->pending_basis = false;
    sv->sre = 1.0;  // every incoming amplitude carried its source shard's scalar
    sv->sim = 0.0;
    CU(cudaEventRecord(sv->ev_done, sv->stream));
    return B2SV_OK;
}

int b2sv_dist_done(b2sv_t* sv) {
    NvtxRange nvtx_range("b2sv_dist_done");
    if (int rc = check(sv)) return rc;
    if (sv->dist_world == 0) return fail(B2SV_ESTATE, "not attached (b2sv_ipc_attach)");
    CU(cudaSetDevice(sv->device));
    // nobody may overwrite a buffer a peer is still gathering from: this rank's stream continues only when every
    // peer's sweep is complete (the caller's host barrier guarantees the events have been recorded)
    for (int r = 0; r < sv->dist_world; ++r)
        if (r != sv->dist_rank) CU(cudaStreamWaitEvent(sv->stream, sv->peers[r].done, 0));
    return B2SV_OK;
}

int b2sv_dist_source(int n_local, int n_global, int dtype, const b2sv_gate* gates, size_t n_gates, char* out, size_t cap, size_t* need) {
    if (n_local < 10 || n_global < 1 || n_local + n_global > 40) return fail(B2SV_EINVAL, "bad register shape %d + %d", n_local, n_global);
    if (n_gates > 0 && !gates) return fail(B2SV_EINVAL, "gates is null");
    std::vector<PGate> pg;
    double sre = 1, sim = 0;
    for (size_t j = 0; j < n_gates; ++j) {
        PGate g;
        if (!normalise_gate(gates[j], (int)j, g, sre, sim)) continue;
        if (!g.is_perm()) return fail(B2SV_EINVAL, "gate %zu is not a permutation gate", j);
        pg.push_back(g);
    }
    const std::string src = perm_dist_source(pg, n_local, n_global, dtype);
    if (need) *need = src.size() + 1;
    if (out && cap > 0) {
        const size_t m = src.size() < cap - 1 ? src.size() : cap - 1;
        std::memcpy(out, src.data(), m);
        out[m] = 0;
    }
    return B2SV_OK;
}

int b2sv_plan_describe(int n_qubits, int dtype, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts,
                       int64_t* out_words, size_t cap, size_t* n_words) {
    if (n_qubits < 1 || n_qubits > 40) return fail(B2SV_EINVAL, "n_qubits %d outside [1,40]", n_qubits);
    if (n_gates > 0 && !gates) return fail(B2SV_EINVAL, "gates is null");
    PlanConfig cfg = default_config(n_qubits, dtype, opts);
    if (n_qubits < 13) cfg.perm_runs = false;
    Plan plan = make_plan(gates, n_gates, cfg, true);
    size_t w = 0;
    auto put = [&](int64_t v) {
        if (out_words && w < cap) out_words[w] = v;
        ++w;
    };
    for (const Segment& seg : plan.segs) {
        std::vector<int> srcs;
        for (int gi : seg.gates) gate_sources(plan, plan.gates[gi], srcs);
        put(seg.cls);
        put((int64_t)srcs.size());
        put((int64_t)seg.tile_mask);
        put((int64_t)seg.common_ctrl);
        put((int64_t)seg.gates.size());
        for (int sidx : srcs) put(sidx);
    }
    {  // trailer: the folded global scalar, as raw IEEE bits
        int64_t re_bits, im_bits;
        std::memcpy(&re_bits, &plan.scalar_re, 8);
        std::memcpy(&im_bits, &plan.scalar_im, 8);
        put(-1);
        put(2);
        put(0);
        put(0);
        put(0);
        put(re_bits);
        put(im_bits);
    }
    {  // second trailer: gates the SWAP frame executed on renamed qubits, and the synthetic SWAPs that restore the layout
        put(-2);
        put((int64_t)plan.relabelled.size());
        put(0);
        put(0);
        put(0);
        for (const Plan::Relabel& r : plan.relabelled) {
            put(r.id);
            put(r.op);
            put(r.t0);
            put(r.t1);
            put((int64_t)r.ctrl);
        }
    }
    if (n_words) *n_words = w;
    return B2SV_OK;
}

int b2sv_perm_jit_compile_check(int n_qubits, int dtype, const b2sv_gate* gates, size_t n_gates, size_t* cubin_bytes) {
    if (n_qubits < 13 || n_qubits > 40) return fail(B2SV_EINVAL, "n_qubits %d outside [13,40]", n_qubits);
    if (n_gates > 0 && !gates) return fail(B2SV_EINVAL, "gates is null");
    std::vector<PGate> pg;
    double sre = 1, sim = 0;
    for (size_t j = 0; j < n_gates; ++j) {
        PGate g;
        if (!normalise_gate(gates[j], (int)j, g, sre, sim)) continue;
        if (!g.is_perm()) return fail(B2SV_EINVAL, "gate %zu is not a permutation gate", j);
        pg.push_back(g);
    }
    std::vector<char> cubin;
    std::string log;
    if (!nvrtc_compile(perm_jit_source(pg, n_qubits, dtype), cubin, log)) return fail(B2SV_ESTATE, "NVRTC: %s", log.c_str());
    if (cubin_bytes) *cubin_bytes = cubin.size();
    return B2SV_OK;
}

int b2sv_stats(b2sv_t* sv, b2sv_stats_t* out) {
    if (int rc = check(sv)) return rc;
    if (!out) return fail(B2SV_EINVAL, "out is null");
    CU(cudaSetDevice(sv->device));
    if (int rc = drain_events(sv)) return rc;
    *out = sv->stats;
    return B2SV_OK;
}

int b2sv_stats_reset(b2sv_t* sv) {
    if (int rc = check(sv)) return rc;
    CU(cudaSetDevice(sv->device));
    if (int rc = drain_events(sv)) return rc;
    sv->stats = b2sv_stats_t{};
    return B2SV_OK;
}

}  // extern "C"
