// tileops.hpp -- host side of the tile sweeps: turns the planned gates of one sweep into the tile-local
// op program the kernels of kernels.cuh interpret (no CUDA calls in this file; the tests also build
// it into a host-only emulator of that program, tests/emul/).
//
//   make_tile_op         one planned gate -> one TileOp (targets / controls as tile-local positions)
//   append_with_phases   groups consecutive ops into register phases (<= 3 target qubits, any plain
//                        gate) and monomial phases (X / SWAP / PHASE / DIAG1 touching up to 11 tile qubits;
//                        their index walks are done here, on the host: mono_build_table)
//   layout_batches       pads the program so that no phase straddles a kTileMaxOps batch boundary
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "planner.hpp"

namespace b2svk {

inline void fill_expand(ExpandArgs& ex, uint64_t mask) {
    ex.m = 0;
    for (int q = 0; q < 64; ++q)
        if ((mask >> q) & 1) ex.pos[ex.m++] = (uint8_t)q;
}

// Translate a planned gate into a tile-local op for the tile with (ascending) qubit positions `ex`.
inline TileOp make_tile_op(const Plan& plan, const PGate& g, uint64_t tile_mask, const ExpandArgs& ex) {
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
        case B2SV_OP_BLOCK: {
            const BlockInfo& bi = plan.blocks[g.blk];
            for (int i = 0; i < bi.kb; ++i) fixed |= 1u << local_pos(bi.q[i]);
        } break;
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
    if (g.op == B2SV_OP_BLOCK) {
        const BlockInfo& bi = plan.blocks[g.blk];
        op.blk.mat_off = bi.mat_off;
        op.blk.kb = (uint8_t)bi.kb;
        for (int i = 0; i < bi.kb; ++i) op.blk.bpos[i] = (uint8_t)local_pos(bi.q[i]);
        op.blk.slot = 0xFF;  // build_segment_ops hands out the launch's constant-bank slots
        bool real = true;
        for (int e = 0; e < (1 << (2 * bi.kb)); ++e) real = real && plan.mux_tables[2 * (bi.mat_off + e) + 1] == 0.0;
        op.blk.real = real ? 1 : 0;
    }
    if (g.op == B2SV_OP_MUX) {
        const MuxInfo& mi = plan.mux[g.mux];
        // g.ctrl (the run's common controls) is already a predicate in lctrl / ext_ctrl
        op.mux.table_off = mi.table_off;
        op.mux.nmux = (uint8_t)mi.nbits;
        int j = 0;
        for (int q = 0; q < 64 && j < mi.nbits; ++q)
            if ((mi.ctrl_mask >> q) & 1) {
                const int lp = local_pos(q);
                op.mux.mpos[j] = lp < 0 ? 0xFF : (uint8_t)lp;
                op.mux.mq[j] = (uint8_t)q;
                ++j;
            }
        // the index bits inside the tile join the fixed positions and are enumerated last (kernels.cuh, kOpMux)
        uint32_t mfixed = fixed;
        op.mux.nl = 0;
        op.mux.lppack = op.mux.ljpack = 0;
        for (int jj = 0; jj < mi.nbits; ++jj)
            if (op.mux.mpos[jj] != 0xFF) {
                mfixed |= 1u << op.mux.mpos[jj];
                op.mux.lppack |= (uint32_t)op.mux.mpos[jj] << (4 * op.mux.nl);
                op.mux.ljpack |= (uint32_t)jj << (4 * op.mux.nl);
                op.mux.nl++;
            }
        op.mux.mfixpos = 0;
        op.mux.mnfix = 0;
        for (int i = 0; i < ex.m; ++i)
            if ((mfixed >> i) & 1) {
                op.mux.mfixpos |= (uint64_t)i << (4 * op.mux.mnfix);
                op.mux.mnfix++;
            }
    }
    return op;
}

constexpr int kPhaseMaxMembers = kTileMaxOps - 1; // header + members fit one staged batch
constexpr size_t kMonoMaxTableBytes = 256 * 1024; // walk results of one monomial phase

// Host side of a monomial phase: walk every (thread, walk) index of the phase through the
// members, once per value of the `ebits` outside qubits the members test (beyond the sweep's common
// controls, which hold in every launched tile).  The walks run bit-sliced, 64 at a time: word b holds
// bit b of 64 indices, an X is `word[t] ^= AND of the control words`.  Results: per walk the XOR mask
// from source to destination index, and (fac) the product of the phase factors it collected.
inline void mono_build_table(const TileOp* members, size_t count, int k, uint32_t tmask, uint32_t qlmask, int V, bool fac,
                             uint64_t common, const uint8_t* epos, int ebits, std::vector<uint16_t>& masks,
                             std::vector<double>& factors, uint32_t& identity) {
    const uint32_t nthreads = 1u << __builtin_popcount(tmask);
    const size_t W = (size_t)nthreads * V;
    std::vector<uint32_t> idx0(W);
    for (uint32_t t = 0; t < nthreads; ++t) {
        const uint32_t tpart = mono_deposit(t, tmask);
        uint32_t q = 0;
        for (int v = 0; v < V; ++v) {
            idx0[(size_t)t * V + v] = tpart | q;
            q = mono_next_subset(q, qlmask);
        }
    }
    masks.assign(W << ebits, 0);
    if (fac) factors.assign((W << ebits) * 2, 0.0);
    identity = 0;
    for (uint32_t slice = 0; slice < (1u << ebits); ++slice) {
        uint64_t have = common;  // outside bits known to be 1 in this slice
        for (int i = 0; i < ebits; ++i)
            if ((slice >> i) & 1) have |= 1ull << epos[i];
        bool ident = true;
        for (size_t w0 = 0; w0 < W; w0 += 64) {
            const int nw = (int)std::min<size_t>(64, W - w0);
            uint64_t word[16] = {0};
            for (int i = 0; i < nw; ++i)
                for (int b = 0; b < k; ++b) word[b] |= (uint64_t)((idx0[w0 + i] >> b) & 1u) << i;
            double2 f[64];
            for (int i = 0; i < 64; ++i) f[i] = make_double2(1.0, 0.0);
            const uint64_t live = nw == 64 ? ~0ull : ((1ull << nw) - 1);
            for (size_t mi = 0; mi < count; ++mi) {
                const TileOp& op = members[mi];
                if ((op.ext_ctrl & ~have) != 0) continue;  // an outside control is 0 in this slice
                uint64_t cond = live;
                for (uint32_t lc = op.lctrl; lc; lc &= lc - 1) cond &= word[__builtin_ctz(lc)];
                if (!cond) continue;
                if (op.op == B2SV_OP_X) {
                    word[op.tpos] ^= cond;
                } else if (op.op == B2SV_OP_SWAP) {
                    const uint64_t d = (word[op.tpos] ^ word[op.tpos2]) & cond;
                    word[op.tpos] ^= d;
                    word[op.tpos2] ^= d;
                } else if (op.op == B2SV_OP_PHASE) {
                    const double2 wv = make_double2(op.m[0], op.m[1]);
                    for (uint64_t c = cond; c; c &= c - 1) {
                        const int i = __builtin_ctzll(c);
                        f[i] = cmul(wv, f[i]);
                    }
                } else {  // DIAG1
                    const double2 d0 = make_double2(op.m[0], op.m[1]), d1 = make_double2(op.m[2], op.m[3]);
                    if (op.tpos == 0xFF) {
                        const double2 d = (have & op.ext_tbit) ? d1 : d0;
                        for (uint64_t c = cond; c; c &= c - 1) {
                            const int i = __builtin_ctzll(c);
                            f[i] = cmul(d, f[i]);
                        }
                    } else {
                        for (uint64_t c = cond; c; c &= c - 1) {
                            const int i = __builtin_ctzll(c);
                            f[i] = cmul(((word[op.tpos] >> i) & 1ull) ? d1 : d0, f[i]);
                        }
                    }
                }
            }
            for (int i = 0; i < nw; ++i) {
                uint32_t fin = 0;
                for (int b = 0; b < k; ++b) fin |= (uint32_t)((word[b] >> i) & 1ull) << b;
                const uint16_t xm = (uint16_t)(fin ^ idx0[w0 + i]);
                const size_t at = ((size_t)slice * W) + w0 + i;
                masks[at] = xm;
                if (xm) ident = false;
                if (fac) {
                    factors[2 * at] = f[i].x;
                    factors[2 * at + 1] = f[i].y;
                    if (!(f[i].x == 1.0 && f[i].y == 0.0)) ident = false;
                }
            }
        }
        if (ident) identity |= 1u << slice;
    }
}

// Group the ops of one tile sweep into phases (kernels.cuh: tile_apply_phase / tile_apply_mono).
//   register phase: maximal run of consecutive plain gates whose targets fall on at most three tile-local
//                   positions -- applied in registers, one shared-memory round trip for the run;
//   monomial phase: maximal run of consecutive X / SWAP / PHASE / DIAG1 gates touching (as target or tile-local
//                   control) at most 11 tile positions (10 with phase factors) -- the index walks are done here,
//                   on the host (mono_build_table); the device fetches their results and moves the amplitudes.
// At every position the kind that swallows more gates wins; a monomial phase of 16 or more gates always does
// (its cost does not grow with the run, a register phase pays per gate).
// `common`: the sweep's common outside controls (set in every launched tile); `tables`: where monomial
// phases put their walk results (Plan::mux_tables).
inline void append_with_phases(const std::vector<TileOp>& ops, int k, bool enable, bool mono, uint64_t common,
                               std::vector<double>* tables, std::vector<TileOp>& out,
                               std::map<std::string, MonoTableRef>* table_cache = nullptr) {
    auto plain = [](const TileOp& o) {
        return o.op == B2SV_OP_MAT1 || o.op == B2SV_OP_DIAG1 || o.op == B2SV_OP_X || o.op == B2SV_OP_SWAP || o.op == B2SV_OP_PHASE;
    };
    auto monomial = [](const TileOp& o) {
        return o.op == B2SV_OP_DIAG1 || o.op == B2SV_OP_X || o.op == B2SV_OP_SWAP || o.op == B2SV_OP_PHASE;
    };
    auto targets = [](const TileOp& o) -> uint32_t {
        uint32_t t = 0;
        if (o.op == B2SV_OP_PHASE) return 0;
        if (o.tpos != 0xFF) t |= 1u << o.tpos;
        if (o.op == B2SV_OP_SWAP) t |= 1u << o.tpos2;
        return t;
    };
    // what the op costs as a stand-alone pass, in units of one read+write pass over the whole tile
    // Measured (scripts/microbench_passes.py, profiles/microbench_passes_r2.txt): inside a sweep that is bound by its passes
    // a stand-alone X / SWAP pass on a QUARTER of the tile still costs 0.13 HBM sweeps and a full-tile 2x2 pass 0.23, against
    // 0.36 for a monomial phase (= 2.5 units here): 0.8 + 0.5 * fraction.  The round-1 model (fraction + 0.08) let runs of
    // three or four lightly controlled SWAPs / Xs go out as passes of their own; with the measured cost they ride in one
    // monomial phase (Pseudoinverse(BPX L=4) at 30 qubits: 4857 -> 4601 passes, tile time 7553 -> 7302 ms; A/B through
    // B2SV_LIGHT_PASS / B2SV_LIGHT_SLOPE, gpurun_out/r2am_*).
    static const double kLightPass = [] { const char* e = std::getenv("B2SV_LIGHT_PASS"); return e ? std::atof(e) : 0.8; }();
    static const double kLightSlope = [] { const char* e = std::getenv("B2SV_LIGHT_SLOPE"); return e ? std::atof(e) : 0.5; }();
    auto alone_cost = [](const TileOp& o) {
        const int c = __builtin_popcount(o.lctrl) + (o.op == B2SV_OP_SWAP ? 1 : 0);
        return kLightSlope * std::ldexp(1.0, -c) + kLightPass;
    };
    auto pass_cost = [](uint32_t common) { return std::ldexp(1.0, -__builtin_popcount(common)); };
    size_t i = 0;
    while (i < ops.size()) {
        if (!enable || k < 3 || !plain(ops[i])) {
            out.push_back(ops[i++]);
            continue;
        }
        // ---- register-phase candidate [i, jr)
        uint32_t R = targets(ops[i]);
        uint32_t K = ops[i].lctrl;  // tile-local controls shared by every gate taken so far
        size_t jr = i + 1;
        while (jr < ops.size() && plain(ops[jr]) && (int)(jr - i) < kPhaseMaxMembers) {
            const uint32_t R2 = R | targets(ops[jr]);
            if (__builtin_popcount(R2) > 3) break;
            // like the sweeps themselves: do not let one lightly controlled gate make a heavily
            // controlled stretch walk over the whole tile
            const double now = pass_cost(K & ~R2), joined = pass_cost(K & ops[jr].lctrl & ~R2), alone = pass_cost(ops[jr].lctrl & ~R2);
            if (joined > now + 0.5 * alone) break;
            R = R2;
            K &= ops[jr].lctrl;
            ++jr;
        }
        // ---- monomial-phase candidate [i, jm)
        size_t jm = i;
        uint32_t Q = 0;      // tile positions the run reads or moves
        uint64_t EX = 0;     // outside qubits the run tests beyond the sweep's common controls
        bool fac = false;
        bool mono_worth = false;
        const int tb = kMonoThreadBits, lb = k - tb;
        auto touched = [&](const TileOp& o) { return o.lctrl | targets(o); };
        auto outside = [&](const TileOp& o) { return (o.ext_ctrl & ~common) | o.ext_tbit; };
        // the walks of a thread: 8 for a pure permutation, 4 when amplitudes pick up factors; at most 16 table
        // slices, and a bounded table
        auto fits = [&](uint32_t q, bool f, uint64_t ex) {
            const int vbits = f ? 2 : 3;
            if (__builtin_popcount(q) > tb + vbits || popcount64(ex) > kMonoMaxEbits) return false;
            const size_t walks = ((size_t)1 << (tb + vbits)) << popcount64(ex);
            return walks * (f ? 18 : 2) <= kMonoMaxTableBytes;
        };
        if (mono && tables && lb >= 3 && k <= 16 && monomial(ops[i])) {
            Q = touched(ops[i]);
            EX = outside(ops[i]);
            fac = !(ops[i].op == B2SV_OP_X || ops[i].op == B2SV_OP_SWAP);
            double separate = alone_cost(ops[i]);
            jm = i + 1;
            // the phase has no members on the device, so it may be as long as the run
            while (fits(Q, fac, EX) && jm < ops.size() && monomial(ops[jm])) {
                const uint32_t Q2 = Q | touched(ops[jm]);
                const uint64_t EX2 = EX | outside(ops[jm]);
                const bool fac2 = fac || !(ops[jm].op == B2SV_OP_X || ops[jm].op == B2SV_OP_SWAP);
                if (!fits(Q2, fac2, EX2)) break;  // close the phase here, the next one starts with this gate
                Q = Q2;
                EX = EX2;
                fac = fac2;
                separate += alone_cost(ops[jm]);
                ++jm;
            }
            // measured (scripts/microbench_mono.py): a phase costs ~0.45 sweep whatever its length = ~2.5 passes
            mono_worth = fits(Q, fac, EX) && jm - i >= 2 && separate > 2.5;
        }
        // a register phase is the cheaper round trip (no bank conflicts, ~1 pass) but pays per gate; a monomial
        // phase costs the same whatever its length
        if (mono_worth && (jm >= jr || jm - i >= 16)) {
            TileOp h{};
            h.op = kOpMono;
            h.mono.count = 0;
            h.mono.fac = fac ? 1 : 0;
            h.mono.lb = (uint8_t)lb;
            h.mono.vbits = (uint8_t)(fac ? 2 : 3);
            // thread bits: Q positions first (ascending), filled up with the lowest other positions; the lowest
            // positions end up in the lane number, so a warp reads whole 128-byte rows whenever it can
            uint32_t tmask = 0;
            for (int p = 0; p < k && __builtin_popcount(tmask) < tb; ++p)
                if ((Q >> p) & 1) tmask |= 1u << p;
            for (int p = 0; p < k && __builtin_popcount(tmask) < tb; ++p)
                if (!((tmask >> p) & 1)) tmask |= 1u << p;
            const uint32_t all = (1u << k) - 1u;
            // walk bits: the Q positions that did not fit the thread number, filled up with the highest others
            uint32_t qlmask = Q & ~tmask & all;
            for (int p = k - 1; p >= 0 && __builtin_popcount(qlmask) < h.mono.vbits; --p)
                if (!((tmask >> p) & 1) && !((qlmask >> p) & 1)) qlmask |= 1u << p;
            h.mono.tmask = tmask;
            h.mono.qlmask = qlmask;
            h.mono.nmask = ~qlmask & ~tmask & all;
            int ne = 0;
            for (int q = 0; q < 64; ++q)
                if ((EX >> q) & 1) h.mono.epos[ne++] = (uint8_t)q;
            h.mono.ebits = (uint8_t)ne;
            // the same run under the same tile layout has been walked before (QSVT steps, LCU terms repeat their
            // arithmetic verbatim): share its tables
            std::string key;
            if (table_cache) {
                key.reserve((jm - i) * 56 + 32);
                auto put = [&](const void* p, size_t nbytes) { key.append(reinterpret_cast<const char*>(p), nbytes); };
                put(&k, sizeof(k));
                put(&tmask, sizeof(tmask));
                put(&qlmask, sizeof(qlmask));
                put(&common, sizeof(common));
                for (size_t q = i; q < jm; ++q) {
                    const TileOp& o = ops[q];
                    put(&o.op, 1);
                    put(&o.tpos, 1);
                    put(&o.tpos2, 1);
                    put(&o.lctrl, sizeof(o.lctrl));
                    put(&o.ext_ctrl, sizeof(o.ext_ctrl));
                    put(&o.ext_tbit, sizeof(o.ext_tbit));
                    if (o.op == B2SV_OP_PHASE || o.op == B2SV_OP_DIAG1) put(o.m, 4 * sizeof(double));
                }
            }
            const auto hit = table_cache ? table_cache->find(key) : decltype(table_cache->end())();
            if (table_cache && hit != table_cache->end()) {
                h.mono.mask_off = hit->second.mask_off;
                h.mono.fac_off = hit->second.fac_off;
                h.mono.identity = hit->second.identity;
            } else {
                std::vector<uint16_t> masks;
                std::vector<double> factors;
                mono_build_table(&ops[i], jm - i, k, tmask, qlmask, 1 << h.mono.vbits, fac, common, h.mono.epos, ne, masks, factors,
                                 h.mono.identity);
                while (tables->size() % 2) tables->push_back(0.0);  // offsets count 16-byte units
                h.mono.mask_off = tables->size() / 2;
                const size_t mask_doubles = (masks.size() * sizeof(uint16_t) + 15) / 16 * 2;
                tables->resize(tables->size() + mask_doubles, 0.0);
                std::memcpy(reinterpret_cast<char*>(tables->data()) + h.mono.mask_off * 16, masks.data(),
                            masks.size() * sizeof(uint16_t));
                if (fac) {
                    h.mono.fac_off = tables->size() / 2;
                    tables->insert(tables->end(), factors.begin(), factors.end());
                }
                if (table_cache) (*table_cache)[key] = MonoTableRef{h.mono.mask_off, h.mono.fac_off, h.mono.identity};
            }
            out.push_back(h);
            i = jm;
            continue;
        }
        if (jr - i < 2) {
            out.push_back(ops[i++]);
            continue;
        }
        // pad R to three positions with the HIGHEST free tile positions that are not shared controls:
        // a thread's eight amplitudes then sit far apart and neighbouring lanes read neighbouring
        // addresses (no bank conflicts)
        for (int p = k - 1; p >= 0 && __builtin_popcount(R) < 3; --p)
            if (!((R >> p) & 1) && !((K >> p) & 1)) R |= 1u << p;
        for (int p = k - 1; p >= 0 && __builtin_popcount(R) < 3; --p)
            if (!((R >> p) & 1)) R |= 1u << p;
        K &= ~R;
        TileOp h{};
        h.op = kOpPhase;
        h.lctrl = K;
        h.ph.count = (uint16_t)(jr - i);
        int nr = 0;
        for (int p = 0; p < k; ++p)
            if ((R >> p) & 1) h.ph.rpos[nr++] = (uint8_t)p;
        h.nfix = 0;
        for (int p = 0; p < k; ++p)
            if (((R | K) >> p) & 1) {
                h.fixpos |= (uint64_t)p << (4 * h.nfix);
                h.nfix++;
            }
        out.push_back(h);
        // Members are pre-decoded for the kernel's inner loop (which runs once per 8-amplitude group: ncu on the BPX
        // phase-only sweeps showed them instruction-issue bound, ~70 instructions per member per group): the fields
        // a member does not need inside a phase carry
        //   nfix          bit j: element j of a group (bit 0/1/2 of j <-> rpos[0/1/2]) satisfies the controls that lie in R
        //   fixpos 0-31   the tile-local controls outside R (one mask test per group)
        //   fixpos 32-33  which register position the target is (3: none), 34-35 the same for a SWAP's second target
        //   fixpos 36     MAT1 with a real matrix (Ry, H, X powers of pi: half the FP64 work)
        for (size_t q = i; q < jr; ++q) {
            TileOp m = ops[q];
            uint32_t ok = 0xFFu;
            for (int b = 0; b < 3; ++b)
                if ((m.lctrl >> h.ph.rpos[b]) & 1u) ok &= b == 0 ? 0xAAu : (b == 1 ? 0xCCu : 0xF0u);
            auto reg_of = [&](uint8_t pos) -> uint64_t {
                for (int b = 0; b < 3; ++b)
                    if (h.ph.rpos[b] == pos) return (uint64_t)b;
                return 3;
            };
            const bool has_t = m.op != B2SV_OP_PHASE && m.tpos != 0xFF;
            const bool real = m.op == B2SV_OP_MAT1 && m.m[1] == 0.0 && m.m[3] == 0.0 && m.m[5] == 0.0 && m.m[7] == 0.0;
            m.nfix = (uint8_t)ok;
            m.fixpos = (uint64_t)(m.lctrl & ~R) | ((has_t ? reg_of(m.tpos) : 3ull) << 32) |
                       ((m.op == B2SV_OP_SWAP ? reg_of(m.tpos2) : 3ull) << 34) | ((real ? 1ull : 0ull) << 36);
            out.push_back(m);
        }
        i = jr;
    }
}

// The kernels stage the program kTileMaxOps slots at a time; a phase (header + members) must sit inside
// one batch, so pad with no-ops in front of a phase that would straddle a boundary.
inline void layout_batches(const std::vector<TileOp>& in, std::vector<TileOp>& out) {
    const size_t first = out.size();
    size_t i = 0;
    while (i < in.size()) {
        const bool hdr = in[i].op == kOpPhase || in[i].op == kOpMono;
        const size_t unit = hdr ? (size_t)in[i].ph.count + 1 : 1;
        const size_t used = (out.size() - first) % kTileMaxOps;
        if (used + unit > (size_t)kTileMaxOps) {
            TileOp nop{};
            nop.op = kOpNop;
            for (size_t p = used; p < (size_t)kTileMaxOps; ++p) out.push_back(nop);
        }
        for (size_t q = 0; q < unit; ++q) out.push_back(in[i + q]);
        i += unit;
    }
}

// Which dense-block matrix goes into which granule of the launch's BlkSlots (kernels.cuh).
struct SlotFill {
    uint8_t granule;
    uint8_t kb;
    size_t mat_off;  // first double2 of the matrix in Plan::mux_tables
};

// Hand the launch's 16 constant-bank granules to the segment's block ops: 8x8 matrices first (four aligned granules
// each), 4x4 matrices into what is left; identical matrices share their slot, ops that find no room keep slot 0xFF
// (the kernel then reads their matrix through the read-only path).
inline void assign_block_slots(std::vector<TileOp>& ops, std::vector<SlotFill>& fills) {
    static const bool disabled = std::getenv("B2SV_NO_BLK_SLOTS") != nullptr;  // A/B switch: matrices through __ldg as in round 1
    if (disabled) return;
    bool used[kBlkGranules] = {};
    auto find = [&](const TileOp& op) -> int {
        for (const SlotFill& f : fills)
            if (f.mat_off == op.blk.mat_off && f.kb == op.blk.kb) return f.granule;
        return -1;
    };
    for (int kb = 3; kb >= 2; --kb)
        for (TileOp& op : ops) {
            if (op.op != kOpBlock || op.blk.kb != kb) continue;
            int g = find(op);
            if (g < 0) {
                const int need = kb == 3 ? 4 : 1;
                for (int c = 0; c + need <= kBlkGranules && g < 0; c += need) {
                    bool free_run = true;
                    for (int i = 0; i < need; ++i) free_run = free_run && !used[c + i];
                    if (free_run) g = c;
                }
                if (g < 0) continue;
                for (int i = 0; i < need; ++i) used[g + i] = true;
                fills.push_back(SlotFill{(uint8_t)g, (uint8_t)kb, (size_t)op.blk.mat_off});
            }
            op.blk.slot = (uint8_t)g;
        }
}

// Lane assignment per pass (kernels.cuh: lane_swaps).  `fixed`: the tile positions the pass does NOT enumerate (targets,
// controls).  The lowest lane bits -- 3 for 16-byte amplitudes (a shared-memory phase is 8 lanes), 4 for 8-byte ones (16
// lanes) -- should land on free positions whose bank-group bits differ: with the 128-byte swizzle the bank group of a
// complex128 element is (l & 7) ^ ((l >> 3) & 7), i.e. position p < 6 drives bit p mod 3; of a complex64 element
// (l & 15) ^ (((l >> 4) & 7) << 1), i.e. positions {0}, {1,4}, {2,5}, {3,6} drive bits 0..3.
inline uint32_t lane_swap_code(uint32_t fixed, int k, int amp_bytes) {
    int F[16], nf = 0;
    for (int p = 0; p < k && nf < 16; ++p)
        if (!((fixed >> p) & 1u)) F[nf++] = p;
    const int lane_bits = amp_bytes == 16 ? 3 : 4, lim = amp_bytes == 16 ? 6 : 7;
    auto cls = [&](int p) { return amp_bytes == 16 ? p % 3 : (p == 0 ? 0 : 1 + (p - 1) % 3); };
    const int m = nf < 8 ? nf : 8;  // only the low 8 bits of the thread number are permuted, and only below log2(groups)
    int chosen[4], nc = 0;
    bool used[8] = {}, have[4] = {};
    for (int i = 0; i < m && nc < lane_bits; ++i)
        if (F[i] < lim && !have[cls(F[i])]) {
            have[cls(F[i])] = true;
            used[i] = true;
            chosen[nc++] = i;
        }
    for (int i = 0; i < m && nc < lane_bits; ++i)
        if (!used[i]) {
            used[i] = true;
            chosen[nc++] = i;
        }
    // thread bit j normally feeds F[j]; make thread bit kk feed F[chosen[kk]] by successive swaps
    int cur[8];
    for (int j = 0; j < 8; ++j) cur[j] = j;
    uint32_t code = 0;
    int ns = 0;
    for (int kk = 0; kk < nc; ++kk) {
        int at = 0;
        while (cur[at] != kk) ++at;
        if (at == chosen[kk]) continue;
        std::swap(cur[at], cur[chosen[kk]]);
        code |= ((uint32_t)at | ((uint32_t)chosen[kk] << 3)) << (6 * ns);
        ++ns;
    }
    return code;
}

inline void assign_lane_swaps(std::vector<TileOp>& ops, int k, int amp_bytes) {
    static const bool disabled = std::getenv("B2SV_NO_LANE_SWAPS") != nullptr;
    if (disabled) return;
    auto fixed_of = [](uint64_t fixpos, int nfix) {
        uint32_t f = 0;
        for (int i = 0; i < nfix; ++i) f |= 1u << ((fixpos >> (4 * i)) & 15);
        return f;
    };
    for (size_t o = 0; o < ops.size(); ++o) {
        TileOp& op = ops[o];
        if (op.op == kOpPhase) {
            op.ext_tbit = lane_swap_code(fixed_of(op.fixpos, op.nfix), k, amp_bytes);
            o += op.ph.count;  // members keep their fields
            continue;
        }
        if (op.op == kOpMono || op.op == kOpNop) continue;
        if (op.op == B2SV_OP_DIAG1 && op.tpos == 0xFF) continue;  // ext_tbit is that gate's target bit
        if (op.op == kOpMux) op.ext_tbit = lane_swap_code(fixed_of(op.mux.mfixpos, op.mux.mnfix), k, amp_bytes);
        else op.ext_tbit = lane_swap_code(fixed_of(op.fixpos, op.nfix), k, amp_bytes);
    }
}

// The complete op program of one SEG_TILE / SEG_LOW6 segment, appended to `out`.
// Table-driven monomial phases append their walk results to plan.mux_tables.
inline void build_segment_ops(Plan& plan, const Segment& seg, const PlanConfig& cfg, std::vector<TileOp>& out,
                              std::vector<SlotFill>* slots = nullptr) {
    ExpandArgs ex;
    fill_expand(ex, seg.tile_mask);
    std::vector<TileOp> seg_ops, grouped;
    seg_ops.reserve(seg.gates.size());
    for (int gi : seg.gates) seg_ops.push_back(make_tile_op(plan, plan.gates[gi], seg.tile_mask, ex));
    if (slots && seg.cls == SEG_TILE) assign_block_slots(seg_ops, *slots);
    const bool tile = seg.cls == SEG_TILE;
    append_with_phases(seg_ops, ex.m, cfg.phases && tile, cfg.mono && tile, seg.common_ctrl, &plan.mux_tables, grouped, &plan.mono_tables);
    if (tile) assign_lane_swaps(grouped, ex.m, cfg.amp_bytes);
    if (tile) layout_batches(grouped, out);
    else out.insert(out.end(), grouped.begin(), grouped.end());
}

}  // namespace b2svk
