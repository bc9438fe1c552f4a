// b2sv.cu -- C ABI (include/b2sv.h) of the B200 statevector engine: handle management, plan
// execution and kernel launches.  Kernels: kernels.cuh.  Fusion planner: planner.hpp.
//
// Replaces, for one statevector, what tq.simulate(..., backend="qulacs") does behind
// Circuit.simulate (/root/reference/src/unitaria/circuit.py:62-70) plus the Subspace projection and
// norm of Node.simulate / simulate_norm (/root/reference/src/unitaria/nodes/node.py:363-403).
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/b2sv.h"
#include "kernels.cuh"
#include "planner.hpp"

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

enum { CLS_GATE1 = 0, CLS_TILE = 1, CLS_PERM = 2, CLS_SCALE = 3, CLS_PROJECT = 4, CLS_INIT = 5, CLS_SWAP = 6, CLS_N = 8 };

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
    double sre = 1.0, sim = 0.0;   // pending global scalar (folded uncontrolled global phases)
    void* d_ops = nullptr;         // device copy of the tile ops of the current apply
    size_t d_ops_cap = 0;
    b2sv_subnode* d_sub = nullptr; // device copy of the current subspace program
    int d_sub_cap = 0;
    double* d_partial = nullptr;   // projection partial sums (+1 slot for the result)
    int n_partial = 0;
    int sm_count = 148;
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
    for (auto& t : sv->timed) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, t.a, t.b));
        sv->stats.ms_by_class[t.cls] += ms;
        sv->event_pool.push_back(t.a);
        sv->event_pool.push_back(t.b);
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

// ---- launch helpers ---------------------------------------------------------------------------

void fill_expand(ExpandArgs& ex, uint64_t mask) {
    ex.m = 0;
    for (int q = 0; q < 64; ++q)
        if ((mask >> q) & 1) ex.pos[ex.m++] = (uint8_t)q;
}

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
    std::memcpy(a.m, g.m, sizeof(a.m));
    return dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, timing, CLS_GATE1, 2.0 * sv->amp_bytes * touched * (double)a.count);
        k_gate1<A><<<grid_for(sv, a.count), kThreads, 0, sv->stream>>>((A*)sv->amps, a);
        return B2SV_OK;
    });
}

// Translate a planned gate into a tile-local op for the tile with (ascending) qubit positions `ex`.
TileOp make_tile_op(const PGate& g, uint64_t tile_mask, const ExpandArgs& ex) {
    TileOp op{};
    auto local_pos = [&](int q) -> int {
        for (int i = 0; i < ex.m; ++i)
            if (ex.pos[i] == q) return i;
        return -1;
    };
    op.op = g.op;
    op.tpos = op.tpos2 = 0;
    op.ext_ctrl = g.ctrl & ~tile_mask;
    op.ext_tbit = 0;
    uint32_t lctrl = 0;
    for (int i = 0; i < ex.m; ++i)
        if ((g.ctrl >> ex.pos[i]) & 1) lctrl |= 1u << i;
    op.lctrl = lctrl;
    uint32_t fixed = lctrl;
    switch (g.op) {
        case B2SV_OP_PHASE:
            break;
        case B2SV_OP_DIAG1: {
            const int p = local_pos(g.t0);
            if (p < 0) {
                op.tpos = 0xFF;
                op.ext_tbit = 1ull << g.t0;
            } else {
                op.tpos = (uint8_t)p;
                fixed |= 1u << p;
            }
        } break;
        case B2SV_OP_SWAP:
            op.tpos = (uint8_t)local_pos(g.t0);
            op.tpos2 = (uint8_t)local_pos(g.t1);
            fixed |= (1u << op.tpos) | (1u << op.tpos2);
            break;
        default:
            op.tpos = (uint8_t)local_pos(g.t0);
            fixed |= 1u << op.tpos;
            break;
    }
    op.nfix = 0;
    op.fixpos = 0;
    for (int i = 0; i < ex.m; ++i)
        if ((fixed >> i) & 1) {
            op.fixpos |= (uint64_t)i << (4 * op.nfix);
            op.nfix++;
        }
    std::memcpy(op.m, g.m, sizeof(op.m));
    return op;
}

template <typename A>
int set_tile_smem(size_t bytes) {
    static size_t configured[2] = {0, 0};
    if (bytes > 48 * 1024) {
        if (configured[0] < bytes) {
            CU(cudaFuncSetAttribute(k_tile<A, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            configured[0] = bytes;
        }
        if (configured[1] < bytes) {
            CU(cudaFuncSetAttribute(k_tile<A, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
            configured[1] = bytes;
        }
    }
    return B2SV_OK;
}

int launch_tile(b2sv* sv, const Segment& seg, const TileOp* d_ops, int n_ops, const PlanConfig& cfg, bool bulk,
                bool timing) {
    TileArgs a{};
    fill_expand(a.tile_ex, seg.tile_mask);
    a.k = a.tile_ex.m;
    int lc = 0;
    while (lc < a.k && a.tile_ex.pos[lc] == lc) ++lc;  // contiguous low run of the tile
    if (lc > cfg.chunk_bits + 3) lc = cfg.chunk_bits + 3;  // keep enough chunks in flight
    if (((size_t)sv->amp_bytes << lc) < 16) return fail(B2SV_EINVAL, "tile chunk smaller than 16 bytes");
    a.lc = lc;
    a.n_ops = n_ops;
    a.ops = d_ops;
    const uint64_t n_tiles = 1ull << (sv->n - a.k);
    const size_t smem = (size_t)sv->amp_bytes << a.k;
    return dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        int rc = set_tile_smem<A>(smem);
        if (rc) return rc;
        LaunchTimer lt(sv, timing, CLS_TILE, 2.0 * sv->amp_bytes * (double)sv->count);
        if (bulk)
            k_tile<A, true><<<(unsigned)n_tiles, kThreads, smem, sv->stream>>>((A*)sv->amps, a);
        else
            k_tile<A, false><<<(unsigned)n_tiles, kThreads, smem, sv->stream>>>((A*)sv->amps, a);
        return B2SV_OK;
    });
}

// Encode a permutation run (execution order gates[first..last)) as the REVERSED word program of its
// inverse map; returns false if it does not fit `cap` ops.
void encode_perm_gate(const PGate& g, int n, std::vector<PermOp>& out) {
    std::vector<uint8_t> ctrls;
    for (int q = 0; q < 64; ++q)
        if ((g.ctrl >> q) & 1) ctrls.push_back((uint8_t)q);
    const uint8_t ones = (uint8_t)n;
    const size_t direct = g.op == B2SV_OP_X ? 3 : 2;  // control slots of the final op
    size_t used = 0;
    // long control lists: leading ACC ops take 4 controls each
    bool first_acc = true;
    while (ctrls.size() - used > direct) {
        PermOp op{};
        op.kind = first_acc ? 2 : 3;
        first_acc = false;
        uint8_t slot[4] = {ones, ones, ones, ones};
        for (int s = 0; s < 4 && ctrls.size() - used > direct; ++s) slot[s] = ctrls[used++];
        op.a = slot[0];
        op.b = slot[1];
        op.c = slot[2];
        op.d = slot[3];
        out.push_back(op);
    }
    uint8_t rest[3] = {ones, ones, ones};
    for (size_t s = 0; used < ctrls.size(); ++s) rest[s] = ctrls[used++];
    PermOp op{};
    if (g.op == B2SV_OP_X) {
        op.kind = 0;
        op.a = g.t0;
        op.b = rest[0];
        op.c = rest[1];
        op.d = rest[2];
    } else {
        op.kind = 1;
        op.a = g.t0;
        op.b = g.t1;
        op.c = rest[0];
        op.d = rest[1];
    }
    out.push_back(op);
}

int launch_perm(b2sv* sv, const Plan& plan, const Segment& seg, bool timing) {
    // split the run so that each chunk's program fits constant memory
    size_t pos = 0;
    while (pos < seg.gates.size()) {
        std::vector<PermOp> fwd;  // ops in execution order, gate by gate
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
        // reverse gate order, keeping each gate's ACC ops in front of its final op
        std::vector<PermOp> prog;
        prog.reserve(fwd.size());
        for (size_t gi = gate_end.size(); gi-- > 0;) {
            const size_t b = gi == 0 ? 0 : gate_end[gi - 1];
            for (size_t q = b; q < gate_end[gi]; ++q) prog.push_back(fwd[q]);
        }
        ConstGuard guard(sv);
        CU(cudaMemcpyToSymbolAsync(c_perm_ops, prog.data(), prog.size() * sizeof(PermOp), 0, cudaMemcpyHostToDevice,
                                   sv->stream));
        // the host vector dies at the end of this iteration: pageable copies are staged synchronously
        const size_t smem = (size_t)(sv->n + 1) * kPermRowBytes;
        const uint64_t blocks = sv->count / (uint64_t)(kPermThreads * 32);
        int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
            using A = std::remove_pointer_t<decltype(tag)>;
            static bool configured = false;
            if (!configured && smem > 48 * 1024) {
                CU(cudaFuncSetAttribute(k_perm<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
                configured = true;
            }
            LaunchTimer lt(sv, timing, CLS_PERM, 2.0 * sv->amp_bytes * (double)sv->count);
            k_perm<A><<<(unsigned)blocks, kPermThreads, smem, sv->stream>>>((const A*)sv->amps, (A*)sv->scratch, sv->n,
                                                                           (int)prog.size());
            return B2SV_OK;
        });
        if (rc) return rc;
        CU(cudaGetLastError());
        std::swap(sv->amps, sv->scratch);
        pos = e;
    }
    return B2SV_OK;
}

int check(const b2sv* sv) {
    if (!sv) return fail(B2SV_EINVAL, "null handle");
    return B2SV_OK;
}

int upload_subprog(b2sv* sv, const b2sv_subnode* prog, int n_nodes, int root_count, int target_qubits) {
    if (!prog && n_nodes > 0) return fail(B2SV_EINVAL, "null subspace program");
    if (n_nodes < 0 || n_nodes > kMaxSubNodes) return fail(B2SV_EINVAL, "subspace program has %d nodes (max %d)", n_nodes, kMaxSubNodes);
    if (root_count < 0 || root_count > n_nodes) return fail(B2SV_EINVAL, "bad root_count");
    if (target_qubits < 0 || target_qubits > sv->n) return fail(B2SV_EINVAL, "target_qubits %d outside [0,%d]", target_qubits, sv->n);
    int bits = 0;
    for (int i = 0; i < root_count; ++i) bits += prog[i].nbits;
    if (bits != target_qubits) return fail(B2SV_EINVAL, "subspace spans %d qubits, target_qubits is %d", bits, target_qubits);
    for (int i = 0; i < n_nodes; ++i) {
        const b2sv_subnode& nd = prog[i];
        if (nd.kind != 0 && nd.kind != 1) return fail(B2SV_EINVAL, "bad subspace node kind");
        if (nd.kind == 1)
            for (int c = 0; c < 2; ++c)
                if (nd.child_count[c] < 0 || nd.child_first[c] < 0 || nd.child_first[c] + nd.child_count[c] > n_nodes)
                    return fail(B2SV_EINVAL, "subspace node %d has children out of range", i);
    }
    if (n_nodes > sv->d_sub_cap) {
        CU(cudaStreamSynchronize(sv->stream));
        cudaFree(sv->d_sub);
        sv->d_sub = nullptr;
        sv->d_sub_cap = 0;
        CU(cudaMalloc(&sv->d_sub, sizeof(b2sv_subnode) * (size_t)(n_nodes + 64)));
        sv->d_sub_cap = n_nodes + 64;
    }
    if (n_nodes > 0)
        CU(cudaMemcpyAsync(sv->d_sub, prog, (size_t)n_nodes * sizeof(b2sv_subnode), cudaMemcpyHostToDevice, sv->stream));
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
    cudaFree(sv->amps);
    cudaFree(sv->scratch);
    cudaFree(sv->d_ops);
    cudaFree(sv->d_partial);
    cudaFree(sv->d_sub);
    cudaStreamDestroy(sv->stream);
    cudaGetLastError();
    delete sv;
    return B2SV_OK;
}

int b2sv_set_basis(b2sv_t* sv, uint64_t index) {
    if (int rc = check(sv)) return rc;
    if (index >= sv->count) return fail(B2SV_EINVAL, "basis index %llu >= 2^%d", (unsigned long long)index, sv->n);
    CU(cudaSetDevice(sv->device));
    sv->sre = 1.0;
    sv->sim = 0.0;
    int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, false, CLS_INIT, (double)sv->amp_bytes * (double)sv->count);
        k_set_basis<A><<<grid_for(sv, sv->count), kThreads, 0, sv->stream>>>((A*)sv->amps, sv->count, index);
        return B2SV_OK;
    });
    CU(cudaGetLastError());
    return rc;
}

// Host buffers are complex128.  complex64 states convert through a bounded device staging buffer.
static const uint64_t kStageAmps = 1ull << 22;  // 64 MiB of complex128

int b2sv_upload(b2sv_t* sv, const void* host) {
    if (int rc = check(sv)) return rc;
    if (!host) return fail(B2SV_EINVAL, "host buffer is null");
    CU(cudaSetDevice(sv->device));
    sv->sre = 1.0;
    sv->sim = 0.0;
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
    if (int rc = check(sv)) return rc;
    if (!host) return fail(B2SV_EINVAL, "host buffer is null");
    CU(cudaSetDevice(sv->device));
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
    if (int rc = check(sv)) return rc;
    if (n_gates > 0 && !gates) return fail(B2SV_EINVAL, "gates is null");
    CU(cudaSetDevice(sv->device));
    for (size_t j = 0; j < n_gates; ++j) {
        const b2sv_gate& g = gates[j];
        if (g.op > B2SV_OP_PHASE) return fail(B2SV_EINVAL, "gate %zu: unknown op %d", j, (int)g.op);
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
    PlanConfig cfg = default_config(sv->n, sv->dtype, opts);
    const bool timing = opts && opts->timing;
    const bool bulk = !(opts && opts->reserved[0] == 1);  // reserved[0] == 1: plain ld/st tile I/O (debug aid)
    bool scratch_ok = false;
    if (cfg.fuse && cfg.perm_runs && sv->n >= 13) {
        // only pay for the second buffer if the circuit has a run that wants it
        size_t run = 0, best = 0;
        for (size_t j = 0; j < n_gates; ++j) {
            if (gates[j].op == B2SV_OP_X || gates[j].op == B2SV_OP_SWAP) best = std::max(best, ++run);
            else run = 0;
        }
        if ((int)best >= cfg.perm_min) scratch_ok = ensure_scratch(sv) == B2SV_OK;
    }
    if (sv->n < 13) cfg.perm_runs = false;
    Plan plan = make_plan(gates, n_gates, cfg, scratch_ok);
    sv->stats.gates_in += plan.gates_in;
    sv->stats.gates_live += plan.gates_live;
    {  // fold the plan's scalar into the pending one
        const double r = sv->sre * plan.scalar_re - sv->sim * plan.scalar_im;
        const double i = sv->sre * plan.scalar_im + sv->sim * plan.scalar_re;
        sv->sre = r;
        sv->sim = i;
    }
    // tile ops of all tile segments, one upload
    std::vector<TileOp> host_ops;
    std::vector<size_t> seg_first(plan.segs.size(), 0);
    for (size_t s = 0; s < plan.segs.size(); ++s) {
        const Segment& seg = plan.segs[s];
        if (seg.cls != SEG_TILE) continue;
        ExpandArgs ex;
        fill_expand(ex, seg.tile_mask);
        seg_first[s] = host_ops.size();
        for (int gi : seg.gates) host_ops.push_back(make_tile_op(plan.gates[gi], seg.tile_mask, ex));
    }
    if (!host_ops.empty()) {
        const size_t bytes = host_ops.size() * sizeof(TileOp);
        if (bytes > sv->d_ops_cap) {
            // a previous apply may still be reading the old buffer
            CU(cudaStreamSynchronize(sv->stream));
            cudaFree(sv->d_ops);
            sv->d_ops = nullptr;
            sv->d_ops_cap = 0;
            CU(cudaMalloc(&sv->d_ops, bytes * 2));
            sv->d_ops_cap = bytes * 2;
        }
        CU(cudaMemcpyAsync(sv->d_ops, host_ops.data(), bytes, cudaMemcpyHostToDevice, sv->stream));
    }
    cudaEvent_t ev_a = get_event(sv), ev_b = get_event(sv);
    CU(cudaEventRecord(ev_a, sv->stream));
    for (size_t s = 0; s < plan.segs.size(); ++s) {
        const Segment& seg = plan.segs[s];
        int rc = B2SV_OK;
        if (seg.cls == SEG_GATE1Q) {
            for (int gi : seg.gates) {
                rc = launch_gate1(sv, plan.gates[gi], timing);
                if (rc) break;
            }
        } else if (seg.cls == SEG_TILE) {
            rc = launch_tile(sv, seg, (const TileOp*)sv->d_ops + seg_first[s], (int)seg.gates.size(), cfg, bulk, timing);
        } else {
            rc = launch_perm(sv, plan, seg, timing);
        }
        if (rc) return rc;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(ev_b, sv->stream));
    sv->apply_events.push_back({ev_a, ev_b});
    // host_ops is pageable memory: the async copy above has been staged, but kernels may still run.
    // Synchronise so that errors surface here and the caller's gate buffer can be reused.
    cudaError_t e = cudaStreamSynchronize(sv->stream);
    if (e != cudaSuccess) return fail(B2SV_ECUDA, "kernel execution failed: %s", cudaGetErrorString(e));
    return B2SV_OK;
}

int b2sv_project_norm2(b2sv_t* sv, const b2sv_subnode* prog, int n_nodes, int root_count, int target_qubits, double* out) {
    if (int rc = check(sv)) return rc;
    if (!out) return fail(B2SV_EINVAL, "out is null");
    CU(cudaSetDevice(sv->device));
    if (int rc = upload_subprog(sv, prog, n_nodes, root_count, target_qubits)) return rc;
    const uint64_t cnt = 1ull << target_qubits;
    const int grid = grid_for(sv, cnt);
    int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, false, CLS_PROJECT, (double)sv->amp_bytes * (double)cnt);
        k_project_norm2<A><<<grid, kThreads, 0, sv->stream>>>((const A*)sv->amps, sv->d_sub, cnt, root_count, sv->d_partial);
        k_reduce_partials<<<1, kThreads, 0, sv->stream>>>(sv->d_partial, grid, sv->d_partial + sv->n_partial);
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

int b2sv_project_gather(b2sv_t* sv, const b2sv_subnode* prog, int n_nodes, int root_count, int target_qubits,
                        double scale_re, double scale_im, void* host_out, uint64_t dim) {
    if (int rc = check(sv)) return rc;
    if (!host_out && dim > 0) return fail(B2SV_EINVAL, "host_out is null");
    CU(cudaSetDevice(sv->device));
    if (int rc = upload_subprog(sv, prog, n_nodes, root_count, target_qubits)) return rc;
    if (dim == 0) return B2SV_OK;
    const uint64_t cnt = 1ull << target_qubits;
    double2* d_out = nullptr;
    CU(cudaMalloc(&d_out, dim * 16));
    cudaMemsetAsync(d_out, 0, dim * 16, sv->stream);
    // the pending global scalar rides along in the scale factor
    const double2 s = make_double2(scale_re * sv->sre - scale_im * sv->sim, scale_re * sv->sim + scale_im * sv->sre);
    dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, false, CLS_PROJECT, (double)sv->amp_bytes * (double)cnt + 16.0 * (double)dim);
        k_project_gather<A><<<grid_for(sv, cnt), kThreads, 0, sv->stream>>>((const A*)sv->amps, sv->d_sub, cnt, root_count, s, d_out, dim);
        return B2SV_OK;
    });
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d_out, dim * 16, cudaMemcpyDeviceToHost, sv->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sv->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(B2SV_ECUDA, "project_gather failed: %s", cudaGetErrorString(e));
    return B2SV_OK;
}

int b2sv_enumerate_basis(b2sv_t* sv, const b2sv_subnode* prog, int n_nodes, int root_count, int target_qubits,
                         int64_t* host_out, uint64_t dim) {
    if (int rc = check(sv)) return rc;
    if (!host_out && dim > 0) return fail(B2SV_EINVAL, "host_out is null");
    CU(cudaSetDevice(sv->device));
    if (int rc = upload_subprog(sv, prog, n_nodes, root_count, target_qubits)) return rc;
    if (dim == 0) return B2SV_OK;
    const uint64_t cnt = 1ull << target_qubits;
    long long* d_out = nullptr;
    CU(cudaMalloc(&d_out, dim * 8));
    cudaMemsetAsync(d_out, 0xFF, dim * 8, sv->stream);
    sv->stats.launches_by_class[CLS_PROJECT]++;
    k_enumerate_basis<<<grid_for(sv, cnt), kThreads, 0, sv->stream>>>(sv->d_sub, cnt, root_count, d_out, dim);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, d_out, dim * 8, cudaMemcpyDeviceToHost, sv->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(sv->stream);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(B2SV_ECUDA, "enumerate_basis failed: %s", cudaGetErrorString(e));
    return B2SV_OK;
}

int b2sv_device_ptr(b2sv_t* sv, void** amps, uint64_t* bytes) {
    if (int rc = check(sv)) return rc;
    CU(cudaSetDevice(sv->device));
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

static int swap_pack_impl(b2sv_t* sv, int q, int keep_bit, void* staging, bool pack) {
    if (int rc = check(sv)) return rc;
    if (q < 0 || q >= sv->n) return fail(B2SV_EINVAL, "qubit %d outside the shard", q);
    if (!staging) return fail(B2SV_EINVAL, "staging is null");
    CU(cudaSetDevice(sv->device));
    if (int rc = flush_scalar(sv, false)) return rc;
    const uint64_t half = sv->count >> 1;
    const uint64_t sel = keep_bit ? 0ull : (1ull << q);  // move the half whose bit differs from keep_bit
    int rc = dispatch_dtype(sv, [&](auto* tag) -> int {
        using A = std::remove_pointer_t<decltype(tag)>;
        LaunchTimer lt(sv, false, CLS_SWAP, 2.0 * sv->amp_bytes * (double)half);
        if (pack)
            k_swap_pack<A, true><<<grid_for(sv, half), kThreads, 0, sv->stream>>>((A*)sv->amps, (A*)staging, half, q, sel);
        else
            k_swap_pack<A, false><<<grid_for(sv, half), kThreads, 0, sv->stream>>>((A*)sv->amps, (A*)staging, half, q, sel);
        return B2SV_OK;
    });
    if (rc) return rc;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(sv->stream));
    return B2SV_OK;
}

int b2sv_swap_pack(b2sv_t* sv, int q, int keep_bit, void* staging_dev) { return swap_pack_impl(sv, q, keep_bit, staging_dev, true); }
int b2sv_swap_unpack(b2sv_t* sv, int q, int keep_bit, const void* staging_dev) {
    return swap_pack_impl(sv, q, keep_bit, const_cast<void*>(staging_dev), false);
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
        put(seg.cls);
        put((int64_t)seg.gates.size());
        put((int64_t)seg.tile_mask);
        for (int gi : seg.gates) put(plan.gates[gi].src);
    }
    {  // trailer: the folded global scalar, as raw IEEE bits
        int64_t re_bits, im_bits;
        std::memcpy(&re_bits, &plan.scalar_re, 8);
        std::memcpy(&im_bits, &plan.scalar_im, 8);
        put(-1);
        put(2);
        put(re_bits);
        put(im_bits);
    }
    if (n_words) *n_words = w;
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
