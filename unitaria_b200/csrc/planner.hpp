// planner.hpp -- host-side fusion planner of the B200 statevector engine (no CUDA in this file).
//
// Input: the lowered gate list handed to b2sv_apply -- one entry per tequila gate of the circuit
// that unitaria's Circuit.simulate would have passed to tq.simulate
// (/root/reference/src/unitaria/circuit.py:62-70).  The reference's simulator (qulacs) makes one
// memory sweep per gate; the planner's job is to make as few sweeps as possible:
//
//   * exact no-ops are dropped (I pads, zero angles, X.X pairs from
//     nodes/basic/modify_control.py:70-78), uncontrolled global phases fold into one scalar;
//   * long runs of X / multi-controlled X / (controlled) SWAP gates -- the ripple-carry arithmetic
//     of circuits/arithmetic.py -- become ONE out-of-place index-map sweep (SEG_PERM);
//   * everything else is packed greedily, with a commutation-aware look-ahead, into shared-memory
//     tile sweeps (SEG_TILE): a sweep owns a set T of `tile_qubits` qubits (always containing the
//     low `chunk_bits` qubits so every global access is a long contiguous run); a gate can ride in
//     the sweep when its non-diagonal targets are in T.  Controls are free (tile predicates) and
//     diagonal gates ride anywhere;
//   * a sweep that would touch only a small fraction of the state (few, heavily controlled gates)
//     is issued as stand-alone gate kernels instead (SEG_GATE1Q) which only visit the 2^(n-c)
//     controlled amplitudes.
//
// The plan never reorders two gates unless they commute by the rule "on every shared qubit both
// gates act diagonally (as a control or as a diagonal target)".
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/b2sv.h"

namespace b2svk {

enum SegClass { SEG_GATE1Q = 0, SEG_TILE = 1, SEG_PERM = 2, SEG_LOW6 = 3 };
constexpr int kLowBits = 6;           // SEG_LOW6: every target within the low 6 qubits -> warp-shuffle sweep
constexpr int kPermMaxCtrlX = 7, kPermMaxCtrlSwap = 6;
constexpr int B2SV_OP_MUX = 5;        // internal: fused same-target run, a table of 2x2 matrices indexed by control bits
constexpr int kMuxMaxCtrl = 6;        // table of at most 64 matrices (4 KiB)
constexpr int B2SV_OP_BLOCK = 6;      // internal: fused run on <= 3 qubits, one dense 2^kb x 2^kb matrix
constexpr int kBlockMaxQubits = 3;
constexpr int kPlanMaxTileOps = 74;   // gates per LOW6 sweep / per tile sweep without monomial phases: with one
                                      // register-phase header per two gates they still fit the 112 op slots
                                      // staged in shared memory
constexpr int kPlanMaxTileGates = 1536;  // gates per tile sweep (the kernels stage long programs in batches)

struct PGate {
    uint8_t op;
    uint8_t t0, t1;
    uint64_t ctrl;
    uint64_t tmask;   // targets that must be tile-resident (non-diagonal action)
    uint64_t qmask;   // every qubit the gate touches
    double m[8];
    int src;          // index in the caller's list
    int mux = -1;     // B2SV_OP_MUX: index into Plan::mux
    int blk = -1;     // B2SV_OP_BLOCK: index into Plan::blocks
    bool is_perm() const { return op == B2SV_OP_X || op == B2SV_OP_SWAP; }
    // fits the word-op encoding of the permutation-run kernel (3 direct control slots, 2 for SWAP,
    // plus two per scratch row)
    bool perm_ok() const {
        const int c = __builtin_popcountll(ctrl);
        return (op == B2SV_OP_X && c <= kPermMaxCtrlX) || (op == B2SV_OP_SWAP && c <= kPermMaxCtrlSwap);
    }
};

struct Segment {
    int cls;
    uint64_t tile_mask;  // SEG_TILE only
    uint64_t common_ctrl = 0;  // SEG_TILE: controls shared by EVERY gate of the sweep that lie outside the
                               // tile -- only the 2^-|common| tiles with those bits set are launched
    std::vector<int> gates;  // indices into Plan::gates, execution order
};

struct PlanConfig {
    int n = 1;
    int amp_bytes = 16;
    bool fuse = true;
    bool perm_runs = true;
    int tile_qubits = 12;
    int chunk_bits = 7;
    int lookahead = 96;
    int perm_min = 12;          // shortest run worth an out-of-place sweep
    double perm_gate_cost = 1.0 / 300.0;  // cost of one perm gate in units of one full sweep
    bool always_tile_below = true;        // small states: prefer the fewest launches
    bool mux = true;                      // fuse same-target runs into multiplexed 2x2 ops
    bool phases = true;                   // group plain gates of a sweep into register phases
    bool low6 = true;                     // sweeps confined to the low 6 qubits use the warp-shuffle kernel
    bool mono = true;                     // monomial phases: runs of X / SWAP / phase gates inside a tile sweep cost
                                          // index arithmetic only, so such runs ride in tile sweeps
    bool frame = true;                    // SWAP frame: (controlled) SWAPs relabel qubits at plan time instead of moving data
    double mono_gate_cost = 0.001;        // one gate of a monomial phase in units of one full sweep: the index walks
                                          // run on the host, the device pays per phase, not per gate
};

// A run of consecutive gates on ONE target whose other qubits are only controls multiplies out to
// sum_c |c><c| (x) M_c: one 2x2 matrix per value c of the union of the control qubits.  This is what
// Moettoenen state preparation emits (circuits/multiplexed_rot.py:8-77: Ry/Rz interleaved with CNOTs
// onto the same target) and what LCU / cosine-sine rotations look like after add_controls.
struct MuxInfo {
    uint64_t ctrl_mask = 0;      // union of the controls, table index = those bits in ascending order
    int nbits = 0;
    size_t table_off = 0;        // first entry in Plan::mux_tables (8 doubles per entry)
    std::vector<int> sources;    // caller's gate indices, execution order
};

// A run of consecutive gates confined to <= 3 qubits (apart from controls they ALL share) is one
// dense 2^kb x 2^kb matrix on those qubits under the shared controls: the cosine-sine two-qubit
// unitaries of circuits/generic_unitary.py, LCU rotations with their X sandwiches, ...
struct BlockInfo {
    int kb = 0;
    uint8_t q[kBlockMaxQubits] = {0, 0, 0};  // block qubits, ascending; matrix index bit i <-> q[i]
    size_t mat_off = 0;                      // first double2 of the row-major matrix in Plan::mux_tables
    std::vector<int> sources;
};

// walk results of a monomial phase already in Plan::mux_tables (block-encoding circuits repeat the same
// arithmetic dozens of times: QSVT steps, LCU terms), keyed by everything the results depend on
struct MonoTableRef {
    uint64_t mask_off = 0, fac_off = 0;
    uint32_t identity = 0;
};

struct Plan {
    std::map<std::string, MonoTableRef> mono_tables;
    std::vector<BlockInfo> blocks;
    std::vector<MuxInfo> mux;
    std::vector<double> mux_tables;
    std::vector<PGate> gates;
    std::vector<Segment> segs;
    double scalar_re = 1.0, scalar_im = 0.0;  // folded uncontrolled global phases
    uint64_t gates_in = 0, gates_live = 0;
    // SWAP frame (make_plan): caller gates that were executed on relabelled qubits, and the synthetic SWAPs that
    // restore the layout (source ids >= gates_in).  {id, op, t0, t1, ctrl} -- exported by b2sv_plan_describe so that a
    // plan can be replayed gate by gate.
    struct Relabel {
        int64_t id;
        uint8_t op, t0, t1;
        uint64_t ctrl;
    };
    std::vector<Relabel> relabelled;
};

inline int popcount64(uint64_t x) { return __builtin_popcountll(x); }
inline uint64_t low_mask(int b) { return b >= 64 ? ~0ull : ((1ull << b) - 1); }

inline PlanConfig default_config(int n, int dtype, const b2sv_plan_opts* o) {
    PlanConfig c;
    c.n = n;
    c.amp_bytes = dtype == B2SV_C64 ? 8 : 16;
    c.tile_qubits = dtype == B2SV_C64 ? 13 : 12;  // 64 KiB of shared memory either way
    // low qubits every tile contains, i.e. the contiguous run one TMA box moves: 256 bytes.  Measured on
    // B200: sweeps stream at the same rate with 256 B .. 2 KiB runs (cfg 2: 7.08 vs 7.14 ms), while every
    // forced low qubit takes a tile position away from the gates' targets (BPX L=8: 289 sweeps / 111 ms
    // with 2 KiB runs, 141 sweeps / 92 ms with 256 B runs)
    c.chunk_bits = dtype == B2SV_C64 ? 5 : 4;
    {   // B2SV_CHUNK_BITS: experiment knob (fewer forced low qubits leave more tile qubits for targets)
        static const int env_chunk = [] {
            const char* e = std::getenv("B2SV_CHUNK_BITS");
            return e ? std::atoi(e) : 0;
        }();
        if (env_chunk > 0) c.chunk_bits = env_chunk;
    }
    // B2SV_AB: A/B switch bits (same meaning as b2sv_plan_opts::reserved) for runs that use default options
    static const int env_ab = [] {
        const char* e = std::getenv("B2SV_AB");
        return e ? std::atoi(e) : 0;
    }();
    c.mux = !(env_ab & 1);
    c.phases = !(env_ab & 2);
    c.low6 = !(env_ab & 4);
    c.mono = !(env_ab & 8);
    c.frame = !(env_ab & 16);
    if (o) {
        c.fuse = o->fuse != 0;
        c.perm_runs = o->perm_runs != 0;
        if (o->tile_qubits > 0) c.tile_qubits = o->tile_qubits;
        if (o->lookahead > 0) c.lookahead = o->lookahead;
        const int ab = o->reserved | env_ab;
        c.mux = !(ab & 1);     // reserved bit 0: no multiplexor / block fusion (A/B switch)
        c.phases = !(ab & 2);  // reserved bit 1: no register phases (A/B switch)
        c.low6 = !(ab & 4);    // reserved bit 2: no warp-shuffle low-qubit sweeps (A/B switch)
        c.mono = !(ab & 8);    // reserved bit 3: no monomial phases (A/B switch)
        c.frame = !(ab & 16);  // reserved bit 4: no SWAP frame (A/B switch)
    }
    if (!c.phases) c.mono = false;
    if (!c.fuse) c.frame = false;
    if (c.tile_qubits > 13 && c.amp_bytes == 16) c.tile_qubits = 13;
    if (c.tile_qubits > 14) c.tile_qubits = 14;
    if (c.tile_qubits < 3) c.tile_qubits = 3;
    if (c.tile_qubits > n) c.tile_qubits = n;
    if (c.chunk_bits > c.tile_qubits - 2) c.chunk_bits = std::max(0, c.tile_qubits - 2);
    if (c.amp_bytes << c.chunk_bits < 16) c.chunk_bits = c.amp_bytes == 8 ? 1 : 0;  // bulk copies need 16 B
    if (c.chunk_bits > c.tile_qubits) c.chunk_bits = c.tile_qubits;
    return c;
}

// ---- normalisation ----------------------------------------------------------------------------

inline bool is_one(double re, double im) { return re == 1.0 && im == 0.0; }
inline bool is_zero(double re, double im) { return re == 0.0 && im == 0.0; }

// Returns false when the gate is an exact no-op.
inline bool normalise_gate(const b2sv_gate& g, int src, PGate& out, double& sre, double& sim) {
    out.op = g.op;
    out.t0 = g.t0;
    out.t1 = g.t1;
    out.ctrl = g.ctrl_mask;
    out.src = src;
    std::memcpy(out.m, g.m, sizeof(out.m));
    if (out.op == B2SV_OP_MAT1) {
        const double* m = out.m;
        if (is_zero(m[2], m[3]) && is_zero(m[4], m[5])) {
            out.op = B2SV_OP_DIAG1;  // {d0, d1} = {m00, m11}
            out.m[2] = m[6];
            out.m[3] = m[7];
        } else if (is_zero(m[0], m[1]) && is_zero(m[6], m[7]) && is_one(m[2], m[3]) && is_one(m[4], m[5])) {
            out.op = B2SV_OP_X;
        }
    }
    if (out.op == B2SV_OP_DIAG1) {
        if (is_one(out.m[0], out.m[1]) && is_one(out.m[2], out.m[3])) return false;
        if (out.m[0] == out.m[2] && out.m[1] == out.m[3]) {  // scalar on the controlled block
            out.op = B2SV_OP_PHASE;
        } else if (is_one(out.m[0], out.m[1])) {  // diag(1, w): a phase on the block where the target is 1 too
            out.op = B2SV_OP_PHASE;
            out.ctrl |= 1ull << out.t0;
            out.m[0] = out.m[2];
            out.m[1] = out.m[3];
        }
    }
    if (out.op == B2SV_OP_PHASE) {
        if (is_one(out.m[0], out.m[1])) return false;
        if (out.ctrl == 0) {  // uncontrolled global phase: fold
            double r = sre * out.m[0] - sim * out.m[1];
            double i = sre * out.m[1] + sim * out.m[0];
            sre = r;
            sim = i;
            return false;
        }
        out.t0 = out.t1 = 0;
        out.tmask = 0;
        out.qmask = out.ctrl;
        return true;
    }
    if (out.op == B2SV_OP_SWAP) {
        if (out.t0 == out.t1) return false;
        out.tmask = (1ull << out.t0) | (1ull << out.t1);
    } else if (out.op == B2SV_OP_DIAG1) {
        out.tmask = 0;
    } else {
        out.tmask = 1ull << out.t0;
    }
    out.qmask = out.ctrl | (1ull << out.t0);
    if (out.op == B2SV_OP_SWAP) out.qmask |= 1ull << out.t1;
    return true;
}

inline bool same_perm_gate(const PGate& a, const PGate& b) {
    if (a.op != b.op || a.ctrl != b.ctrl) return false;
    if (a.op == B2SV_OP_X) return a.t0 == b.t0;
    return (a.t0 == b.t0 && a.t1 == b.t1) || (a.t0 == b.t1 && a.t1 == b.t0);
}

inline double gate_sweep_fraction(const PGate& g) {
    int c = popcount64(g.ctrl);
    if (g.op == B2SV_OP_SWAP) c += 1;
    return std::ldexp(1.0, -c);
}

// ---- same-target multiplexor fusion -------------------------------------------------------------

inline bool single_target_gate(const PGate& g) {
    return g.op == B2SV_OP_MAT1 || g.op == B2SV_OP_DIAG1 || g.op == B2SV_OP_X;
}

inline void gate_matrix2(const PGate& g, double m[8]) {
    for (int i = 0; i < 8; ++i) m[i] = 0.0;
    if (g.op == B2SV_OP_MAT1) {
        for (int i = 0; i < 8; ++i) m[i] = g.m[i];
    } else if (g.op == B2SV_OP_DIAG1) {
        m[0] = g.m[0];
        m[1] = g.m[1];
        m[6] = g.m[2];
        m[7] = g.m[3];
    } else {  // X
        m[2] = 1.0;
        m[4] = 1.0;
    }
}

// c = a * b (2x2 complex, row-major, (re,im) pairs)
inline void matmul2(const double a[8], const double b[8], double c[8]) {
    for (int r = 0; r < 2; ++r)
        for (int col = 0; col < 2; ++col) {
            double re = 0, im = 0;
            for (int k = 0; k < 2; ++k) {
                const double ar = a[2 * (2 * r + k)], ai = a[2 * (2 * r + k) + 1];
                const double br = b[2 * (2 * k + col)], bi = b[2 * (2 * k + col) + 1];
                re += ar * br - ai * bi;
                im += ar * bi + ai * br;
            }
            c[2 * (2 * r + col)] = re;
            c[2 * (2 * r + col) + 1] = im;
        }
}

// Returns false when the gate is an exact no-op (same contract as normalise_gate) -- used to
// re-normalise the product of a fused run.
inline bool renormalise(PGate& g, double& sre, double& sim) {
    b2sv_gate raw{};
    raw.op = g.op;
    raw.t0 = g.t0;
    raw.t1 = g.t1;
    raw.ctrl_mask = g.ctrl;
    std::memcpy(raw.m, g.m, sizeof(raw.m));
    const int src = g.src;
    return normalise_gate(raw, src, g, sre, sim);
}

inline void fuse_multiplexors(Plan& plan) {
    std::vector<PGate> out;
    out.reserve(plan.gates.size());
    const std::vector<PGate>& G = plan.gates;
    size_t i = 0;
    while (i < G.size()) {
        const PGate& g = G[i];
        if (!single_target_gate(g)) {
            out.push_back(g);
            ++i;
            continue;
        }
        uint64_t U = g.ctrl, K = g.ctrl;  // union / intersection of the controls of the run
        bool nonperm = !g.is_perm();
        double separate = gate_sweep_fraction(g);
        size_t j = i + 1;
        while (j < G.size() && single_target_gate(G[j]) && G[j].t0 == g.t0 &&
               popcount64((U | G[j].ctrl) & ~(K & G[j].ctrl)) <= kMuxMaxCtrl) {
            U |= G[j].ctrl;
            K &= G[j].ctrl;
            nonperm |= !G[j].is_perm();
            separate += gate_sweep_fraction(G[j]);
            ++j;
        }
        const size_t len = j - i;
        const uint64_t C = U & ~K;  // the table is indexed by the non-common controls only
        const double scale = std::ldexp(1.0, -popcount64(K));
        // one table-driven pass over the K-controlled block costs about 2.5 plain passes; a product
        // with a single entry is just another 2x2 gate
        // measured (profiles/fusion_r1.txt): the table-driven pass only pays for long multiplexor runs
        // (state preparation); short controlled runs are cheaper as plain gates
        // (tried in round 2: leaving 1-2 bit multiplexors as plain gates for a register phase -- config 3's sweep went
        // from 2.24 ms to 2.7 ms, the per-gate loop of the phase costs more than the table look-ups)
        const bool worth = nonperm && len >= 2 && (C == 0 ? true : (len >= 6 && separate > 1.8 * scale));
        if (!worth) {
            out.push_back(g);
            ++i;
            continue;
        }
        int cpos[kMuxMaxCtrl];
        int nb = 0;
        for (int q = 0; q < 64; ++q)
            if ((C >> q) & 1) cpos[nb++] = q;
        std::vector<double> table;
        table.reserve((size_t)8 << nb);
        for (uint32_t c = 0; c < (1u << nb); ++c) {
            uint64_t pattern = K;
            for (int b = 0; b < nb; ++b)
                if ((c >> b) & 1) pattern |= 1ull << cpos[b];
            double M[8] = {1, 0, 0, 0, 0, 0, 1, 0};
            for (size_t q = i; q < j; ++q) {
                if ((G[q].ctrl & pattern) != G[q].ctrl) continue;
                double Ug[8], R[8];
                gate_matrix2(G[q], Ug);
                matmul2(Ug, M, R);
                for (int e = 0; e < 8; ++e) M[e] = R[e];
            }
            table.insert(table.end(), M, M + 8);
        }
        PGate f;
        f.t0 = g.t0;
        f.t1 = 0;
        f.ctrl = K;  // a real predicate: only the K-controlled block is touched
        f.src = g.src;
        MuxInfo info;
        for (size_t q = i; q < j; ++q) info.sources.push_back(G[q].src);
        if (nb == 0) {
            // plain product: still one gate, but remember what it stands for
            f.op = B2SV_OP_MAT1;
            for (int e = 0; e < 8; ++e) f.m[e] = table[e];
            if (renormalise(f, plan.scalar_re, plan.scalar_im)) {
                info.ctrl_mask = 0;
                info.nbits = -1;  // marks "sources only"
                f.mux = (int)plan.mux.size();
                plan.mux.push_back(std::move(info));
                out.push_back(f);
            }
            i = j;
            continue;
        }
        info.ctrl_mask = C;
        info.nbits = nb;
        info.table_off = plan.mux_tables.size() / 8;
        plan.mux_tables.insert(plan.mux_tables.end(), table.begin(), table.end());
        f.op = B2SV_OP_MUX;
        f.tmask = 1ull << g.t0;
        f.qmask = U | f.tmask;
        for (int e = 0; e < 8; ++e) f.m[e] = 0;
        f.mux = (int)plan.mux.size();
        plan.mux.push_back(std::move(info));
        out.push_back(f);
        i = j;
    }
    plan.gates.swap(out);
}

// How much more the separate gates must cost than the fused op before fusing (tunable for experiments
// through B2SV_FUSE_MARGIN; the default comes from the sweep in profiles/fusion_r1.txt).
inline double fuse_margin() {
    static const double v = [] {
        const char* e = std::getenv("B2SV_FUSE_MARGIN");
        const double x = e ? std::atof(e) : 0.0;
        return x > 0.0 ? x : 1.0;
    }();
    return v;
}

// ---- dense few-qubit block fusion ---------------------------------------------------------------

inline bool blockable_gate(const PGate& g) {
    return g.op == B2SV_OP_MAT1 || g.op == B2SV_OP_DIAG1 || g.op == B2SV_OP_X || g.op == B2SV_OP_SWAP || g.op == B2SV_OP_PHASE;
}

inline uint64_t gate_targets(const PGate& g) {
    if (g.op == B2SV_OP_PHASE) return 0;
    uint64_t t = 1ull << g.t0;
    if (g.op == B2SV_OP_SWAP) t |= 1ull << g.t1;
    return t;
}

// A small multiplexor (target + index bits on <= 3 qubits) is a block-diagonal 4x4 / 8x8 matrix on those qubits: a dense
// block that already covers them absorbs it for free.  Config 3's closing sweep was MUX (index bits 1, 2 -> target 0) +
// an 8x8 block on {0,1,2} + a monomial phase: three shared-memory passes where two do (the whole Moettoenen rotation
// tree of circuits/multiplexed_rot.py on three qubits is ONE real 8x8 matrix).  For the block DP the index bits count
// as block qubits (they select the matrix, so they can never be "common controls").
inline bool blockable_mux(const Plan& plan, const PGate& g) {
    static const bool disabled = std::getenv("B2SV_NO_MUX_IN_BLOCK") != nullptr;
    return !disabled && g.op == B2SV_OP_MUX && g.mux >= 0 && plan.mux[g.mux].nbits >= 1 && plan.mux[g.mux].nbits + 1 <= kBlockMaxQubits;
}
inline bool blockable_gate(const Plan& plan, const PGate& g) { return blockable_gate(g) || blockable_mux(plan, g); }
inline uint64_t block_targets(const Plan& plan, const PGate& g) {
    return g.op == B2SV_OP_MUX ? (1ull << g.t0) | plan.mux[g.mux].ctrl_mask : gate_targets(g);
}
inline bool block_gate_is_real(const Plan& plan, const PGate& g);

// Apply gate g to a 2^kb vector (interleaved re,im) whose bit i stands for qubit q[i]; controls in
// `common` are assumed satisfied.
inline void block_apply_gate(double* v, int kb, const uint8_t* q, const PGate& g, uint64_t common, const Plan* plan = nullptr) {
    auto local = [&](int qubit) {
        for (int i = 0; i < kb; ++i)
            if (q[i] == qubit) return i;
        return -1;
    };
    uint32_t lc = 0;
    for (int b = 0; b < 64; ++b)
        if (((g.ctrl & ~common) >> b) & 1) lc |= 1u << local(b);
    const uint32_t dim = 1u << kb;
    if (g.op == B2SV_OP_MUX) {  // table entry chosen by the index bits (ascending qubit order), then a plain 2x2
        const MuxInfo& mi = plan->mux[g.mux];
        int ipos[kMuxMaxCtrl], nb = 0;
        for (int b = 0; b < 64; ++b)
            if ((mi.ctrl_mask >> b) & 1) ipos[nb++] = local(b);
        const uint32_t bt = 1u << local(g.t0);
        for (uint32_t i = 0; i < dim; ++i)
            if ((i & lc) == lc && !(i & bt)) {
                uint32_t idx = 0;
                for (int b = 0; b < nb; ++b) idx |= ((i >> ipos[b]) & 1u) << b;
                const double* U = plan->mux_tables.data() + 8 * (mi.table_off + idx);
                const double a0r = v[2 * i], a0i = v[2 * i + 1], a1r = v[2 * (i | bt)], a1i = v[2 * (i | bt) + 1];
                v[2 * i] = U[0] * a0r - U[1] * a0i + U[2] * a1r - U[3] * a1i;
                v[2 * i + 1] = U[0] * a0i + U[1] * a0r + U[2] * a1i + U[3] * a1r;
                v[2 * (i | bt)] = U[4] * a0r - U[5] * a0i + U[6] * a1r - U[7] * a1i;
                v[2 * (i | bt) + 1] = U[4] * a0i + U[5] * a0r + U[6] * a1i + U[7] * a1r;
            }
        return;
    }
    if (g.op == B2SV_OP_PHASE) {
        for (uint32_t i = 0; i < dim; ++i)
            if ((i & lc) == lc) {
                const double re = v[2 * i] * g.m[0] - v[2 * i + 1] * g.m[1];
                const double im = v[2 * i] * g.m[1] + v[2 * i + 1] * g.m[0];
                v[2 * i] = re;
                v[2 * i + 1] = im;
            }
        return;
    }
    if (g.op == B2SV_OP_SWAP) {
        const uint32_t ba = 1u << local(g.t0), bb = 1u << local(g.t1);
        for (uint32_t i = 0; i < dim; ++i)
            if ((i & lc) == lc && (i & ba) && !(i & bb)) {
                const uint32_t j = i ^ ba ^ bb;
                std::swap(v[2 * i], v[2 * j]);
                std::swap(v[2 * i + 1], v[2 * j + 1]);
            }
        return;
    }
    double U[8];
    gate_matrix2(g, U);
    const uint32_t bt = 1u << local(g.t0);
    for (uint32_t i = 0; i < dim; ++i)
        if ((i & lc) == lc && !(i & bt)) {
            const double a0r = v[2 * i], a0i = v[2 * i + 1], a1r = v[2 * (i | bt)], a1i = v[2 * (i | bt) + 1];
            v[2 * i] = U[0] * a0r - U[1] * a0i + U[2] * a1r - U[3] * a1i;
            v[2 * i + 1] = U[0] * a0i + U[1] * a0r + U[2] * a1i + U[3] * a1r;
            v[2 * (i | bt)] = U[4] * a0r - U[5] * a0i + U[6] * a1r - U[7] * a1i;
            v[2 * (i | bt) + 1] = U[4] * a0i + U[5] * a0r + U[6] * a1i + U[7] * a1r;
        }
}

inline void gate_sources(const Plan& plan, const PGate& g, std::vector<int>& out) {
    if (g.mux >= 0) out.insert(out.end(), plan.mux[g.mux].sources.begin(), plan.mux[g.mux].sources.end());
    else if (g.blk >= 0) out.insert(out.end(), plan.blocks[g.blk].sources.begin(), plan.blocks[g.blk].sources.end());
    else if (g.src >= 0) out.push_back(g.src);  // (src < 0: a phase factored out of dense blocks, see fuse_blocks)
}

// FP64 work of a plain gate inside a register phase, in DFMA per 8 amplitudes (`common`: controls every gate of the
// run shares -- they scale both alternatives alike).  Register phases apply a gate to the four pairs a thread
// holds: a complex 2x2 is 4 x 16 DFMA, a real one half of that, a diagonal 4 x 8; X and SWAP are register moves.
inline bool gate_is_real(const PGate& g) {
    if (g.op == B2SV_OP_MAT1) return g.m[1] == 0.0 && g.m[3] == 0.0 && g.m[5] == 0.0 && g.m[7] == 0.0;
    if (g.op == B2SV_OP_DIAG1) return g.m[1] == 0.0 && g.m[3] == 0.0;
    if (g.op == B2SV_OP_PHASE) return g.m[1] == 0.0;
    return true;
}
inline bool block_gate_is_real(const Plan& plan, const PGate& g) {
    if (g.op != B2SV_OP_MUX) return gate_is_real(g);
    const MuxInfo& mi = plan.mux[g.mux];
    const double* t = plan.mux_tables.data() + 8 * mi.table_off;
    for (size_t e = 0; e < ((size_t)4 << mi.nbits); ++e)
        if (t[2 * e + 1] != 0.0) return false;
    return true;
}
// What a gate costs as a register-phase member beyond its arithmetic (decode, predicates, the idle threads of a phase
// whose members share controls): it does NOT shrink with the number of controls.  Measured on the 30-qubit BPX sweeps:
// 29 mostly doubly-controlled gates in six phases took 6.7 ms of compute where their DFMA count predicted 1.1 ms.
inline double member_overhead() {
    static const double v = [] {
        const char* e = std::getenv("B2SV_MEMBER_COST");
        return e ? std::atof(e) : 64.0;
    }();
    return v;
}
inline double gate_fp64_cost(const PGate& g, uint64_t common) {
    const double f = std::ldexp(1.0, -popcount64(g.ctrl & ~common));
    switch (g.op) {
        case B2SV_OP_MAT1: return f * (gate_is_real(g) ? 32.0 : 64.0) + member_overhead();
        case B2SV_OP_DIAG1:
        case B2SV_OP_PHASE: return f * 32.0 + member_overhead();
        default: return member_overhead();
    }
}
// one dense block pass: 2^kb complex FMAs per amplitude (half for a real matrix) plus a shared-memory round trip of
// its own (a register phase would have carried the gates along with its neighbours)
inline double block_fp64_cost(int kb, bool real) { return (kb == 3 ? 256.0 : 128.0) * (real ? 0.5 : 1.0) * fuse_margin() + 96.0; }

// Build the dense matrix of G[a, b) on its <= 3 qubits and append the BLOCK gate to `out`.
// Largest imaginary residue (after dividing by the common phase) that is treated as rounding noise of the host-side
// matrix product and dropped.  The residues of the BPX / Pseudoinverse blocks are <= 5e-16 (products of ~25 gates).
inline double block_phase_tol() {
    static const double v = [] {
        const char* e = std::getenv("B2SV_BLOCK_PHASE_TOL");
        return e ? std::atof(e) : 2e-15;
    }();
    return v;
}

// Returns the unit scalar factored out of the matrix in (*ph_re, *ph_im): a block whose entries are all real multiples
// of ONE phase e^{i phi} (the cosine-sine two-qubit unitaries of circuits/generic_unitary.py are: 46 of the 50 complex
// 4x4 blocks of BPX L=8, all 684 of Pseudoinverse(BPX L=4)) is stored as the real matrix -- half the FP64 work of the
// pass -- and the phase is left to the caller, who merges the phases of a whole chain of blocks into one gate.
inline void emit_block(Plan& plan, const std::vector<PGate>& G, size_t a, size_t b, std::vector<PGate>& out, double* ph_re, double* ph_im) {
    uint64_t K = G[a].ctrl, T = block_targets(plan, G[a]), U = G[a].ctrl;
    for (size_t q = a + 1; q < b; ++q) {
        K &= G[q].ctrl;
        T |= block_targets(plan, G[q]);
        U |= G[q].ctrl;
    }
    const uint64_t B = T | (U & ~K);
    BlockInfo info;
    info.kb = popcount64(B);
    int nb = 0;
    for (int bit = 0; bit < 64; ++bit)
        if ((B >> bit) & 1) info.q[nb++] = (uint8_t)bit;
    const uint32_t dim = 1u << info.kb;
    std::vector<double> mat((size_t)2 * dim * dim, 0.0);  // row-major, (re,im)
    for (uint32_t col = 0; col < dim; ++col) {
        double v[2 * 8] = {0};
        v[2 * col] = 1.0;
        for (size_t qg = a; qg < b; ++qg) block_apply_gate(v, info.kb, info.q, G[qg], K, &plan);
        for (uint32_t row = 0; row < dim; ++row) {
            mat[2 * ((size_t)row * dim + col)] = v[2 * row];
            mat[2 * ((size_t)row * dim + col) + 1] = v[2 * row + 1];
        }
    }
    *ph_re = 1.0;
    *ph_im = 0.0;
    {
        bool real = true;
        double best = 0.0, pr = 1.0, pi = 0.0;
        for (size_t e = 0; e < (size_t)dim * dim; ++e) {
            real = real && mat[2 * e + 1] == 0.0;
            const double mag = std::hypot(mat[2 * e], mat[2 * e + 1]);
            if (mag > best) {
                best = mag;
                pr = mat[2 * e] / mag;
                pi = mat[2 * e + 1] / mag;
            }
        }
        if (!real && best > 0.0) {
            bool up_to_phase = true;  // every entry * conj(phase) is real (to rounding: the entries are <= 1 in magnitude)
            for (size_t e = 0; e < (size_t)dim * dim && up_to_phase; ++e)
                up_to_phase = std::fabs(mat[2 * e + 1] * pr - mat[2 * e] * pi) <= block_phase_tol();
            if (up_to_phase) {
                for (size_t e = 0; e < (size_t)dim * dim; ++e) {
                    mat[2 * e] = mat[2 * e] * pr + mat[2 * e + 1] * pi;
                    mat[2 * e + 1] = 0.0;
                }
                *ph_re = pr;
                *ph_im = pi;
            }
        }
    }
    while (plan.mux_tables.size() % 8) plan.mux_tables.push_back(0.0);  // keep multiplexor entries aligned
    info.mat_off = plan.mux_tables.size() / 2;
    plan.mux_tables.insert(plan.mux_tables.end(), mat.begin(), mat.end());
    for (size_t qg = a; qg < b; ++qg) gate_sources(plan, G[qg], info.sources);
    PGate f;
    f.op = B2SV_OP_BLOCK;
    f.t0 = info.q[0];
    f.t1 = 0;
    f.ctrl = K;
    f.tmask = B;
    f.qmask = B | K;
    for (int e = 0; e < 8; ++e) f.m[e] = 0;
    f.src = G[a].src;
    f.blk = (int)plan.blocks.size();
    plan.blocks.push_back(std::move(info));
    out.push_back(f);
}

// A dense-block pass is bound by its shared-memory round trip, not by its arithmetic (ncu on the BPX chains of 4x4
// blocks: L1 data pipe 75 % busy, FP64 pipe 48 %; storing the blocks as real matrices -- half the DFMAs -- did not move
// the time).  Two blocks under the same controls whose qubit sets overlap in one qubit ({1,5} then {1,7}: the chains
// of circuits/generic_unitary.py alternate between two such families) cost the same arithmetic as ONE 8x8 block on the
// union but half the round trips.  They are rarely adjacent in the gate list, so the DP cannot see them: this pass
// moves a block up to an earlier one when every gate in between commutes with it for the plain reason that they share
// no qubit (control-control overlaps are fine), and multiplies the two out.
inline void merge_commuting_blocks(Plan& plan, std::vector<PGate>& L) {
    static const bool disabled = std::getenv("B2SV_NO_BLOCK_MERGE") != nullptr;
    if (disabled) return;
    auto commutes = [](const PGate& g, const PGate& b) { return (g.qmask & b.tmask) == 0 && (g.tmask & b.qmask) == 0; };
    for (size_t j = 1; j < L.size(); ++j) {
        if (L[j].op != B2SV_OP_BLOCK) continue;
        // walk back over commuting gates to the nearest earlier block that can absorb L[j]
        for (size_t i = j; i-- > 0 && j - i <= 4096;) {
            const PGate& cand = L[i];
            if (cand.op == B2SV_OP_BLOCK && cand.ctrl == L[j].ctrl && popcount64(cand.tmask | L[j].tmask) <= kBlockMaxQubits &&
                (cand.tmask & L[j].tmask) != 0) {
                const BlockInfo a = plan.blocks[cand.blk], b = plan.blocks[L[j].blk];
                const uint64_t B = cand.tmask | L[j].tmask;
                BlockInfo info;
                info.kb = popcount64(B);
                int nb = 0;
                for (int bit = 0; bit < 64; ++bit)
                    if ((B >> bit) & 1) info.q[nb++] = (uint8_t)bit;
                const int dim = 1 << info.kb;
                auto apply = [&](const BlockInfo& blk, double* v) {  // v: 2^kb complex entries indexed by info.q
                    int pos[3];
                    for (int t = 0; t < blk.kb; ++t)
                        for (int u = 0; u < info.kb; ++u)
                            if (info.q[u] == blk.q[t]) pos[t] = u;
                    const int bd = 1 << blk.kb;
                    const double* M = plan.mux_tables.data() + 2 * blk.mat_off;
                    double w[16];
                    for (int rest = 0; rest < dim; ++rest) {
                        bool base = true;
                        for (int t = 0; t < blk.kb; ++t) base = base && !((rest >> pos[t]) & 1);
                        if (!base) continue;
                        int idx[8];
                        for (int c = 0; c < bd; ++c) {
                            idx[c] = rest;
                            for (int t = 0; t < blk.kb; ++t)
                                if ((c >> t) & 1) idx[c] |= 1 << pos[t];
                        }
                        for (int r = 0; r < bd; ++r) {
                            double re = 0, im = 0;
                            for (int c = 0; c < bd; ++c) {
                                const double mr = M[2 * (r * bd + c)], mi = M[2 * (r * bd + c) + 1];
                                re += mr * v[2 * idx[c]] - mi * v[2 * idx[c] + 1];
                                im += mr * v[2 * idx[c] + 1] + mi * v[2 * idx[c]];
                            }
                            w[2 * r] = re;
                            w[2 * r + 1] = im;
                        }
                        for (int r = 0; r < bd; ++r) {
                            v[2 * idx[r]] = w[2 * r];
                            v[2 * idx[r] + 1] = w[2 * r + 1];
                        }
                    }
                };
                std::vector<double> mat((size_t)2 * dim * dim, 0.0);
                for (int col = 0; col < dim; ++col) {
                    double v[16] = {0};
                    v[2 * col] = 1.0;
                    apply(a, v);
                    apply(b, v);
                    for (int row = 0; row < dim; ++row) {
                        mat[2 * ((size_t)row * dim + col)] = v[2 * row];
                        mat[2 * ((size_t)row * dim + col) + 1] = v[2 * row + 1];
                    }
                }
                while (plan.mux_tables.size() % 8) plan.mux_tables.push_back(0.0);
                info.mat_off = plan.mux_tables.size() / 2;
                plan.mux_tables.insert(plan.mux_tables.end(), mat.begin(), mat.end());
                info.sources = a.sources;
                info.sources.insert(info.sources.end(), b.sources.begin(), b.sources.end());
                PGate f = cand;
                f.t0 = info.q[0];
                f.tmask = B;
                f.qmask = B | f.ctrl;
                f.blk = (int)plan.blocks.size();
                plan.blocks.push_back(std::move(info));
                L[i] = f;
                L.erase(L.begin() + (long)j);
                --j;
                break;
            }
            if (!commutes(cand, L[j])) break;
        }
    }
}

// The gate list is cut into pieces, each either left as plain gates (they ride in register phases) or multiplied out
// into one dense 4x4 / 8x8 block, so that the FP64 work is smallest: a dynamic programme over the whole list
// (dp[p] = least cost of the first p gates; a block may end at p and start at any q for which G[q, p) stays on <= 3
// qubits).  Costs are DFMA per 8 amplitudes of the FULL state: a gate with c controls costs 2^-c of its uncontrolled
// self, a block 2^-|K| with K the controls all its gates share.  ncu on the BPX sweeps (profiles/ncu_r2_bpx.txt) showed
// what the greedy "one block per maximal run" of round 1 cost: the 2-qubit cosine-sine unitaries of
// circuits/generic_unitary.py (6 Ry, 9 Rz, 6 phases, 6 CNOT: 672 DFMA per 8 amplitudes as plain gates, 128 as a 4x4
// block) were followed by ONE SWAP with a third qubit, which turned each of them into an 8x8 block at twice the work.
inline void fuse_blocks(Plan& plan) {
    const std::vector<PGate>& G = plan.gates;
    const size_t N = G.size();
    if (N == 0) return;
    // Two states per prefix: the last piece is a plain stretch (P) or a block (B).  A plain stretch that follows a block
    // (or starts the list) opens a shared-memory pass of its own, like every block does: without that term the DP cut
    // "block + one left-over gate" where one register phase does the same work in one round trip (config 2's sweeps).
    const double inf = 1e300;
    std::vector<double> dpP(N + 1, inf), dpB(N + 1, inf);
    std::vector<uint32_t> cutB(N + 1, 0);
    std::vector<char> fromB_P(N + 1, 0), fromB_B(N + 1, 0);  // predecessor state of the P / B state at p (1: block)
    dpB[0] = 0.0;
    constexpr size_t kMaxBlockGates = 512;
    static const double kPassCost = [] { const char* e = std::getenv("B2SV_PASS_COST"); return e ? std::atof(e) : 48.0; }();  // A/B over 0/24/48/96: 96 cost the Pseudoinverse 3 %
    for (size_t p = 1; p <= N; ++p) {
        const PGate& last = G[p - 1];
        const double open = dpB[p - 1] + kPassCost;  // (not scaled by the gate's controls: barrier + latency of a pass are fixed)
        fromB_P[p] = open < dpP[p - 1] ? 1 : 0;
        dpP[p] = gate_fp64_cost(last, 0) + (fromB_P[p] ? open : dpP[p - 1]);
        if (!blockable_gate(plan, last)) continue;
        uint64_t K = last.ctrl, T = block_targets(plan, last), U = last.ctrl;
        bool nonperm = !last.is_perm(), real = block_gate_is_real(plan, last);
        for (size_t q = p - 1; q-- > 0 && p - q <= kMaxBlockGates;) {
            const PGate& g = G[q];
            if (!blockable_gate(plan, g)) break;
            K &= g.ctrl;
            T |= block_targets(plan, g);
            U |= g.ctrl;
            const int kb = popcount64(T | (U & ~K));
            if (kb > kBlockMaxQubits || (T & K) != 0) break;
            nonperm |= !g.is_perm();
            real = real && block_gate_is_real(plan, g);
            if (p - q < 3 || kb < 2 || !nonperm) continue;
            const bool prevB = dpB[q] < dpP[q];
            const double c = (prevB ? dpB[q] : dpP[q]) + std::ldexp(block_fp64_cost(kb, real), -popcount64(K));
            if (c < dpB[p]) {
                dpB[p] = c;
                cutB[p] = (uint32_t)q;
                fromB_B[p] = prevB ? 1 : 0;
            }
        }
    }
    std::vector<std::pair<size_t, size_t>> pieces;
    std::vector<char> blk(N + 1, 0);
    {
        size_t p = N;
        bool inB = dpB[N] < dpP[N];
        while (p > 0) {
            if (inB) {
                pieces.push_back({cutB[p], p});
                blk[p] = 1;
                inB = fromB_B[p] != 0;
                p = cutB[p];
            } else {
                pieces.push_back({p - 1, p});
                inB = fromB_P[p] != 0;
                p = p - 1;
            }
        }
    }
    std::vector<PGate> out;
    out.reserve(N);
    // Phases factored out of blocks wait here, keyed by the blocks' shared controls K: a phase on the K-controlled
    // subspace is diagonal, so it commutes with everything that does not act non-diagonally on a qubit of K and the
    // phases of a chain of blocks under the same selector controls merge into ONE gate (emitted with src = -1: the
    // blocks' source lists already stand for the whole unitaries).
    struct Pending {
        uint64_t K;
        double re, im;
    };
    std::vector<Pending> pending;
    auto emit_phase = [&](const Pending& ph) {
        if (ph.re == 1.0 && ph.im == 0.0) return;
        PGate f{};
        f.op = B2SV_OP_PHASE;
        f.t0 = f.t1 = 0;
        f.ctrl = ph.K;
        f.tmask = 0;
        f.qmask = ph.K;
        f.m[0] = ph.re;
        f.m[1] = ph.im;
        f.src = -1;
        out.push_back(f);
    };
    auto flush_touching = [&](uint64_t nondiag) {
        for (size_t i = 0; i < pending.size();) {
            if (pending[i].K & nondiag) {
                emit_phase(pending[i]);
                pending.erase(pending.begin() + (long)i);
            } else {
                ++i;
            }
        }
    };
    for (size_t k = pieces.size(); k-- > 0;) {
        const size_t a = pieces[k].first, b = pieces[k].second;
        if (blk[b]) {
            double pr = 1.0, pi = 0.0;
            const size_t at = out.size();
            emit_block(plan, G, a, b, out, &pr, &pi);
            const PGate blockgate = out[at];
            if (pr != 1.0 || pi != 0.0) {
                if (blockgate.ctrl == 0) {  // unconditioned: a global phase, folded into the lazy scalar
                    const double r = plan.scalar_re * pr - plan.scalar_im * pi, i = plan.scalar_re * pi + plan.scalar_im * pr;
                    plan.scalar_re = r;
                    plan.scalar_im = i;
                } else {
                    // the block itself may act on a qubit of an older pending phase's K: that phase goes in FRONT of it
                    out.pop_back();
                    flush_touching(blockgate.tmask);
                    out.push_back(blockgate);
                    bool merged = false;
                    for (Pending& ph : pending)
                        if (ph.K == blockgate.ctrl) {
                            const double r = ph.re * pr - ph.im * pi, i = ph.re * pi + ph.im * pr;
                            ph.re = r;
                            ph.im = i;
                            merged = true;
                        }
                    if (!merged) pending.push_back({blockgate.ctrl, pr, pi});
                }
            } else if (!pending.empty()) {
                out.pop_back();
                flush_touching(blockgate.tmask);
                out.push_back(blockgate);
            }
        } else {
            for (size_t q = a; q < b; ++q) {
                if (!pending.empty()) flush_touching(G[q].tmask);
                out.push_back(G[q]);
            }
        }
    }
    for (const Pending& ph : pending) emit_phase(ph);
    merge_commuting_blocks(plan, out);
    plan.gates.swap(out);
}

// ---- tile packing -----------------------------------------------------------------------------

// Cost of a sweep in units of one full pass over the state.
inline double sweep_cost(uint64_t common_outside) { return std::ldexp(1.0, -popcount64(common_outside)); }

inline void plan_tiles(const std::vector<PGate>& G, const std::vector<int>& idx, const PlanConfig& cfg,
                       std::vector<Segment>& out) {
    const int n = cfg.n, k = cfg.tile_qubits;
    std::vector<char> done(idx.size(), 0);
    size_t head = 0;
    const uint64_t base = low_mask(std::min(cfg.chunk_bits, n));
    while (head < idx.size()) {
        if (done[head]) {
            ++head;
            continue;
        }
        Segment seg;
        seg.cls = SEG_TILE;
        uint64_t T = base;
        uint64_t K = ~0ull;  // controls common to every gate taken so far
        uint64_t blocked_any = 0, blocked_nd = 0;
        // the look-ahead window counts the gates passed over, not the ones taken: with monomial phases a
        // sweep can carry hundreds of gates
        int scanned = 0;
        const int max_gates = cfg.mono ? kPlanMaxTileGates : kPlanMaxTileOps;
        for (size_t p = head; p < idx.size() && scanned < cfg.lookahead && (int)seg.gates.size() < max_gates; ++p) {
            if (done[p]) continue;
            if (!cfg.mono) ++scanned;
            const PGate& g = G[idx[p]];
            const bool commutes = ((g.tmask & blocked_any) == 0) && ((g.qmask & blocked_nd) == 0);
            const uint64_t T2 = T | g.tmask;
            const bool fits = popcount64(T2) <= k;
            // Block-encoding circuits put whole stretches under the same selector controls
            // (BlockDiagonal / LCU / QSVT flag qubits).  A sweep whose gates ALL share controls outside
            // the tile only has to visit the tiles where those bits are 1, so do not let one gate
            // with fewer controls make everybody else's sweep bigger.
            bool cheap = true;
            if (!seg.gates.empty()) {
                const double now = sweep_cost(K & ~T2);
                const double joined = sweep_cost(K & g.ctrl & ~T2);
                const double alone = sweep_cost(g.ctrl & ~(base | g.tmask));
                cheap = joined <= now + 0.5 * alone;
            }
            if (commutes && fits && cheap) {
                T = T2;
                K &= g.ctrl;
                seg.gates.push_back(idx[p]);
                done[p] = 1;
            } else {
                if (cfg.mono) ++scanned;
                blocked_any |= g.qmask;
                blocked_nd |= g.tmask;
                if (p == head) break;  // cannot happen (k >= chunk_bits + 2), but never spin
            }
        }
        if (seg.gates.empty()) {  // defensive: issue the head gate alone
            seg.gates.push_back(idx[head]);
            done[head] = 1;
            T |= G[idx[head]].tmask;
            K = G[idx[head]].ctrl;
        }
        // pad with the lowest qubits that are not common controls (a bit fixed to 1 is wasted in a tile)
        for (int q = 0; q < n && popcount64(T) < k; ++q)
            if (!((K >> q) & 1)) T |= 1ull << q;
        for (int q = 0; q < n && popcount64(T) < k; ++q) T |= 1ull << q;
        seg.tile_mask = T;
        seg.common_ctrl = K & ~T;
        // A sweep that touches little of the state is cheaper as stand-alone controlled kernels.
        double frac = 0;
        for (int gi : seg.gates) frac += gate_sweep_fraction(G[gi]);
        const bool small_state = cfg.always_tile_below && n <= 16;
        bool has_mux = false;
        for (int gi : seg.gates) has_mux |= G[gi].op == B2SV_OP_MUX || G[gi].op == B2SV_OP_BLOCK;
        if (!small_state && !has_mux && frac <= 0.55 * sweep_cost(seg.common_ctrl) && seg.gates.size() <= 8) {
            for (int gi : seg.gates) {
                Segment one;
                one.cls = SEG_GATE1Q;
                one.tile_mask = 0;
                one.gates.push_back(gi);
                out.push_back(std::move(one));
            }
        } else {
            // state-preparation style sweeps: every (non-diagonal) target sits in the low kLowBits qubits,
            // i.e. inside one 64-amplitude block a warp can hold in registers and exchange by shuffles
            bool low = cfg.low6 && n >= kLowBits + 1 && (int)seg.gates.size() <= kPlanMaxTileOps;
            for (int gi : seg.gates) {
                const PGate& g = G[gi];
                low = low && (g.tmask >> kLowBits) == 0 && g.op != B2SV_OP_SWAP && g.op != B2SV_OP_BLOCK;
            }
            if (low) {
                seg.cls = SEG_LOW6;
                seg.tile_mask = low_mask(kLowBits);
                uint64_t Kc = ~0ull;
                for (int gi : seg.gates) Kc &= G[gi].ctrl;
                seg.common_ctrl = Kc & ~seg.tile_mask;
            }
            out.push_back(std::move(seg));
        }
    }
}

// ---- whole plan -------------------------------------------------------------------------------

// Inside a sweep the op program pays per PASS (a shared-memory round trip), and a pass is a run of like gates: a
// monomial phase takes any number of consecutive X / SWAP / phase gates, a register phase consecutive plain gates on <=
// 3 targets.  The arithmetic-plus-rotation stretches of the QSVT circuits alternate the two kinds gate by gate (a
// 41-gate sweep of Pseudoinverse(BPX) became 15 passes, 3.7 x its HBM time), although most neighbours commute for the
// plain reason that neither acts non-diagonally on a qubit the other touches.  This pass moves every gate back over such
// commuting neighbours to the nearest earlier gate of its own kind, so the kinds cluster and the runs get long.
inline void cluster_commuting(const std::vector<PGate>& G, Segment& seg) {
    static const bool disabled = std::getenv("B2SV_NO_CLUSTER") != nullptr;
    if (disabled || seg.gates.size() < 3) return;
    auto kind = [&](const PGate& g) { return (g.op == B2SV_OP_X || g.op == B2SV_OP_SWAP || g.op == B2SV_OP_PHASE || g.op == B2SV_OP_DIAG1) ? 0 : 1; };
    auto commutes = [](const PGate& g, const PGate& h) { return (g.tmask & h.qmask) == 0 && (h.tmask & g.qmask) == 0; };
    std::vector<int> out;
    out.reserve(seg.gates.size());
    for (int gi : seg.gates) {
        const PGate& g = G[gi];
        const int kg = kind(g);
        size_t at = out.size();
        for (size_t k = out.size(); k-- > 0 && out.size() - k <= 256;) {
            const PGate& h = G[out[k]];
            if (kind(h) == kg) {
                at = k + 1;
                break;
            }
            if (!commutes(g, h)) break;
        }
        out.insert(out.begin() + (long)at, gi);
    }
    seg.gates.swap(out);
}

inline Plan make_plan(const b2sv_gate* gates, size_t n_gates, const PlanConfig& cfg, bool scratch_available) {
    Plan plan;
    plan.gates_in = n_gates;
    plan.gates.reserve(n_gates);
    // SWAP frame.  A (controlled) SWAP does not have to move amplitudes: it renames two qubits.  While every following
    // gate carries the SWAP's controls C (then it only acts where the SWAP did), the rename is applied to the gates
    // instead -- phys[q] is where logical qubit q lives -- and the data moves once, when a gate without those controls
    // arrives or the circuit ends (a few SWAPs that ride in one monomial phase).  Block-encoding circuits are full of
    // this: nodes/permutation/permutation.py emits SWAP networks between the two-qubit unitaries of an LCU term, all
    // under the term's selector controls; ncu / per-sweep timings of BPX L=8 showed one shared-memory pass per SWAP
    // chain (13 ms of a 69 ms step) and, worse, the chains kept neighbouring unitaries from fusing.
    struct Frame {
        bool active = false;
        uint64_t C = 0, moved = 0;
        uint8_t phys[64];
    } fr;
    int64_t next_virtual = (int64_t)n_gates;
    uint64_t absorbed = 0;  // SWAPs executed by renaming (they are live gates of the circuit all the same)
    auto push_gate = [&](const b2sv_gate& raw, int64_t id) {
        PGate g;
        if (!normalise_gate(raw, (int)id, g, plan.scalar_re, plan.scalar_im)) return;
        // adjacent self-inverse pairs cancel (X.X, SWAP.SWAP with identical controls)
        if (g.is_perm() && !plan.gates.empty() && same_perm_gate(plan.gates.back(), g)) {
            plan.gates.pop_back();
            return;
        }
        plan.gates.push_back(g);
    };
    auto flush_frame = [&]() {
        if (!fr.active) return;
        fr.active = false;
        if (!fr.moved) return;
        uint8_t inv[64];
        for (int q = 0; q < 64; ++q) inv[fr.phys[q]] = (uint8_t)q;
        for (int q = 0; q < 64; ++q) {
            if (fr.phys[q] == q) continue;
            const int x = fr.phys[q], other = inv[q];  // logical q lives at x; position q holds logical `other`
            b2sv_gate sw{};
            sw.op = B2SV_OP_SWAP;
            sw.n_targets = 2;
            sw.t0 = (uint8_t)q;
            sw.t1 = (uint8_t)x;
            sw.ctrl_mask = fr.C;
            plan.relabelled.push_back({next_virtual, B2SV_OP_SWAP, sw.t0, sw.t1, sw.ctrl_mask});
            push_gate(sw, next_virtual++);
            fr.phys[q] = (uint8_t)q;
            fr.phys[other] = (uint8_t)x;
            inv[q] = (uint8_t)q;
            inv[x] = (uint8_t)other;
        }
        fr.moved = 0;
    };
    for (size_t j = 0; j < n_gates; ++j) {
        b2sv_gate raw = gates[j];
        if (cfg.frame) {
            const bool is_swap = raw.op == B2SV_OP_SWAP && raw.t0 != raw.t1;
            uint64_t qm = raw.ctrl_mask;
            if (raw.op != B2SV_OP_PHASE) qm |= 1ull << raw.t0;
            if (raw.op == B2SV_OP_SWAP) qm |= 1ull << raw.t1;
            for (int pass = 0; pass < 2; ++pass) {  // second pass: the same gate after a flush
                if (!fr.active) {
                    if (is_swap) {  // start a frame with this SWAP
                        fr.active = true;
                        fr.C = raw.ctrl_mask;
                        fr.moved = 0;
                        for (int q = 0; q < 64; ++q) fr.phys[q] = (uint8_t)q;
                    } else {
                        break;  // no frame, ordinary gate
                    }
                }
                if (is_swap && raw.ctrl_mask == fr.C) {  // absorbed: rename
                    std::swap(fr.phys[raw.t0], fr.phys[raw.t1]);
                    fr.moved |= (1ull << raw.t0) | (1ull << raw.t1);
                    bool ident = true;
                    for (int q = 0; q < 64 && ident; ++q) ident = fr.phys[q] == q;
                    if (ident) fr.moved = 0;
                    ++absorbed;
                    raw.op = 0xFF;  // nothing to emit
                    break;
                }
                if ((raw.ctrl_mask & fr.C) == fr.C) {  // acts only where the deferred SWAPs did: rename its qubits
                    if (fr.moved & qm) {
                        uint64_t c2 = 0;
                        for (int q = 0; q < 64; ++q)
                            if ((raw.ctrl_mask >> q) & 1) c2 |= 1ull << fr.phys[q];
                        raw.ctrl_mask = c2;
                        if (raw.op != B2SV_OP_PHASE) raw.t0 = fr.phys[raw.t0];
                        if (raw.op == B2SV_OP_SWAP) raw.t1 = fr.phys[raw.t1];
                        plan.relabelled.push_back({(int64_t)j, raw.op, raw.t0, raw.t1, raw.ctrl_mask});
                    }
                    break;
                }
                if (((fr.moved | fr.C) & qm) == 0) break;  // touches neither the renamed qubits nor the controls: commutes
                flush_frame();                             // the data has to move now; then look at the gate again
            }
            if (raw.op == 0xFF) continue;
        }
        push_gate(raw, (int64_t)j);
    }
    flush_frame();
    {   // live gates of the CALLER's circuit: what survived normalisation, plus renamed SWAPs, minus synthetic ones
        uint64_t synthetic = 0;
        for (const PGate& g : plan.gates) synthetic += (uint64_t)g.src >= n_gates ? 1 : 0;
        plan.gates_live = plan.gates.size() + absorbed - synthetic;
    }
    if (cfg.fuse && cfg.mux) {
        fuse_multiplexors(plan);
        fuse_blocks(plan);
    }
    const std::vector<PGate>& G = plan.gates;
    if (!cfg.fuse) {
        for (size_t j = 0; j < G.size(); ++j) {
            Segment s;
            s.cls = SEG_GATE1Q;
            s.tile_mask = 0;
            s.gates.push_back((int)j);
            plan.segs.push_back(std::move(s));
        }
        return plan;
    }
    // split into stretches at long permutation runs
    std::vector<int> stretch;
    auto flush_stretch = [&]() {
        if (!stretch.empty()) {
            plan_tiles(G, stretch, cfg, plan.segs);
            stretch.clear();
        }
    };
    size_t j = 0;
    while (j < G.size()) {
        if (G[j].perm_ok() && cfg.perm_runs && scratch_available) {
            size_t e = j;
            while (e < G.size() && G[e].perm_ok()) ++e;
            const size_t len = e - j;
            bool use_perm = false;
            if ((int)len >= cfg.perm_min) {
                std::vector<int> run(len);
                for (size_t q = 0; q < len; ++q) run[q] = (int)(j + q);
                std::vector<Segment> alt;
                plan_tiles(G, run, cfg, alt);
                double tile_cost = 0;
                for (const Segment& s : alt) {
                    if (s.cls == SEG_TILE || s.cls == SEG_LOW6)
                        tile_cost += sweep_cost(s.common_ctrl) * (1.0 + (cfg.mono ? cfg.mono_gate_cost * (double)s.gates.size() : 0.0));
                    else tile_cost += gate_sweep_fraction(G[s.gates[0]]);
                }
                double perm_cost = 1.0 + cfg.perm_gate_cost * (double)len;
                uint64_t run_targets = 0;
                for (size_t q = j; q < e; ++q) run_targets |= G[q].tmask;
                if (cfg.mono && popcount64(low_mask(std::min(cfg.chunk_bits, cfg.n)) | run_targets) <= cfg.tile_qubits) {
                    // a compact run (its targets fit one tile) rides with its neighbours as a monomial phase:
                    // what it adds to their sweep is one shared-memory round trip (plus index arithmetic when
                    // the walks run on the device), while a permutation sweep of its own also cuts the
                    // neighbours' sweep in two
                    tile_cost = 0.3 + cfg.mono_gate_cost * (double)len;
                    perm_cost += 0.3;
                }
                use_perm = perm_cost < tile_cost;
                if (cfg.n <= 16) use_perm = alt.size() > 1;  // launch-bound regime: fewest launches
            }
            if (use_perm) {
                flush_stretch();
                Segment s;
                s.cls = SEG_PERM;
                s.tile_mask = 0;
                for (size_t q = j; q < e; ++q) s.gates.push_back((int)q);
                plan.segs.push_back(std::move(s));
            } else {
                for (size_t q = j; q < e; ++q) stretch.push_back((int)q);
            }
            j = e;
        } else {
            stretch.push_back((int)j);
            ++j;
        }
    }
    flush_stretch();
    for (Segment& seg : plan.segs)
        if (seg.cls == SEG_TILE) cluster_commuting(plan.gates, seg);
    return plan;
}

}  // namespace b2svk
