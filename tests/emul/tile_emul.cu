// tile_emul.cu -- TEST INFRASTRUCTURE, never shipped and never loaded by unitaria_b200.
//
// Host-only interpreter of the tile-op programs that unitaria_b200/csrc/tileops.hpp builds for the
// CUDA kernels of kernels.cuh.  The GPU is scarce while developing; this lets the CPU test-suite check
// everything the host prepares for a tile sweep -- tile-local positions, fixed-position enumeration,
// register-phase and monomial-phase headers with their host-built walk tables, batch padding -- by
// running the very same TileOp arrays and tables over a small state and comparing with the CPU oracle.
// It includes the product's planner and op builders unchanged; only the interpretation below is written
// a second time.
//
// Built by tests/test_tile_emul.py with nvcc (host code only; no kernel is launched).
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../unitaria_b200/csrc/kernels.cuh"
#include "../../unitaria_b200/csrc/planner.hpp"
#include "../../unitaria_b200/csrc/tileops.hpp"

using namespace b2svk;
typedef std::complex<double> cplx;

namespace {

inline cplx c_of(const double* m, int i) { return cplx(m[2 * i], m[2 * i + 1]); }

// reference application of one planned plain gate on the full state (SEG_GATE1Q / SEG_PERM segments)
void apply_plain(std::vector<cplx>& psi, int n, const PGate& g) {
    const uint64_t N = 1ull << n;
    const uint64_t bt = g.op == B2SV_OP_PHASE ? 0 : 1ull << g.t0;
    for (uint64_t i = 0; i < N; ++i) {
        if ((i & g.ctrl) != g.ctrl) continue;
        switch (g.op) {
            case B2SV_OP_PHASE:
                psi[i] *= c_of(g.m, 0);
                break;
            case B2SV_OP_X:
                if (!(i & bt)) std::swap(psi[i], psi[i | bt]);
                break;
            case B2SV_OP_SWAP: {
                const uint64_t ba = 1ull << g.t0, bb = 1ull << g.t1;
                if ((i & ba) && !(i & bb)) std::swap(psi[i], psi[i ^ ba ^ bb]);
            } break;
            case B2SV_OP_DIAG1:
                psi[i] *= (i & bt) ? c_of(g.m, 1) : c_of(g.m, 0);
                break;
            case B2SV_OP_MAT1:
                if (!(i & bt)) {
                    const cplx a = psi[i], b = psi[i | bt];
                    psi[i] = c_of(g.m, 0) * a + c_of(g.m, 1) * b;
                    psi[i | bt] = c_of(g.m, 2) * a + c_of(g.m, 3) * b;
                }
                break;
            default:
                break;
        }
    }
}

struct Counters {
    int64_t tile = 0, perm = 0, gate1q = 0, low6 = 0, mono_phases = 0, mono_members = 0, reg_phases = 0, reg_members = 0, max_ops = 0,
            nops = 0, mono_rounds_max = 0, mono_gates = 0;
};

// one op (not a header) on a tile held as tile[l], l = tile-local index
void run_op(std::vector<cplx>& tile, const TileOp& op, int k, uint64_t base, const Plan& plan) {
    if ((base & op.ext_ctrl) != op.ext_ctrl) return;
    const uint32_t cnt = 1u << (k - op.nfix);
    const double2* tables = reinterpret_cast<const double2*>(plan.mux_tables.data());
    for (uint32_t p = 0; p < cnt; ++p) {
        const uint32_t l = expand_local(p, op.fixpos, op.nfix) | op.lctrl;
        switch (op.op) {
            case B2SV_OP_X: {
                const uint32_t bt = 1u << op.tpos;
                std::swap(tile[l], tile[l | bt]);
            } break;
            case B2SV_OP_SWAP: {
                const uint32_t ba = 1u << op.tpos, bb = 1u << op.tpos2;
                std::swap(tile[l | ba], tile[l | bb]);
            } break;
            case B2SV_OP_MAT1: {
                const uint32_t bt = 1u << op.tpos;
                const cplx a = tile[l], b = tile[l | bt];
                tile[l] = c_of(op.m, 0) * a + c_of(op.m, 1) * b;
                tile[l | bt] = c_of(op.m, 2) * a + c_of(op.m, 3) * b;
            } break;
            case B2SV_OP_DIAG1: {
                if (op.tpos == 0xFF) {
                    tile[l] *= (base & op.ext_tbit) ? c_of(op.m, 1) : c_of(op.m, 0);
                } else {
                    const uint32_t bt = 1u << op.tpos;
                    tile[l] *= c_of(op.m, 0);
                    tile[l | bt] *= c_of(op.m, 1);
                }
            } break;
            case kOpMux: {
                const uint32_t bt = 1u << op.tpos;
                uint32_t idx = 0;
                for (int j = 0; j < op.mux.nmux; ++j) {
                    if (op.mux.mpos[j] == 0xFF) idx |= (uint32_t)((base >> op.mux.mq[j]) & 1ull) << j;
                    else idx |= ((l >> op.mux.mpos[j]) & 1u) << j;
                }
                const double2* M = tables + (op.mux.table_off + idx) * 4;
                const cplx m00(M[0].x, M[0].y), m01(M[1].x, M[1].y), m10(M[2].x, M[2].y), m11(M[3].x, M[3].y);
                const cplx a = tile[l], b = tile[l | bt];
                tile[l] = m00 * a + m01 * b;
                tile[l | bt] = m10 * a + m11 * b;
            } break;
            case kOpBlock: {
                const int kb = op.blk.kb;
                const uint32_t dim = 1u << kb;
                const double2* U = tables + op.blk.mat_off;
                uint32_t off[8];
                cplx v[8], w[8];
                for (uint32_t j = 0; j < dim; ++j) {
                    off[j] = l;
                    for (int b = 0; b < kb; ++b)
                        if ((j >> b) & 1) off[j] |= 1u << op.blk.bpos[b];
                    v[j] = tile[off[j]];
                }
                for (uint32_t r = 0; r < dim; ++r) {
                    w[r] = 0;
                    for (uint32_t c = 0; c < dim; ++c) w[r] += cplx(U[r * dim + c].x, U[r * dim + c].y) * v[c];
                }
                for (uint32_t j = 0; j < dim; ++j) tile[off[j]] = w[j];
            } break;
            default:  // PHASE
                tile[l] *= c_of(op.m, 0);
                break;
        }
    }
}

void run_reg_phase(std::vector<cplx>& tile, const TileOp& hdr, const TileOp* ops, int k, uint64_t base) {
    const uint32_t r0 = hdr.ph.rpos[0], r1 = hdr.ph.rpos[1], r2 = hdr.ph.rpos[2];
    const uint32_t b0 = 1u << r0, b1 = 1u << r1, b2 = 1u << r2;
    const uint32_t cnt = 1u << (k - hdr.nfix);
    for (uint32_t g = 0; g < cnt; ++g) {
        const uint32_t l = expand_local(g, hdr.fixpos, hdr.nfix) | hdr.lctrl;
        uint32_t off[8];
        cplx v[8];
        for (int j = 0; j < 8; ++j) {
            off[j] = l | ((j & 1) ? b0 : 0u) | ((j & 2) ? b1 : 0u) | ((j & 4) ? b2 : 0u);
            v[j] = tile[off[j]];
        }
        for (int mi = 0; mi < hdr.ph.count; ++mi) {
            const TileOp& op = ops[mi];
            if ((base & op.ext_ctrl) != op.ext_ctrl) continue;
            // the kernel reads the host's pre-decoded member fields (tileops.hpp): so does the emulator
            const uint32_t lco = (uint32_t)op.fixpos, ok = op.nfix;
            if ((l & lco) != lco) continue;
            const int rp = (int)((op.fixpos >> 32) & 3u), rq = (int)((op.fixpos >> 34) & 3u);
            auto jbit = [&](uint32_t pos) { return pos == r0 ? 1 : (pos == r1 ? 2 : 4); };
            if (op.op != B2SV_OP_PHASE && op.tpos != 0xFF && (1 << rp) != jbit(op.tpos)) std::abort();
            if (op.op == B2SV_OP_SWAP && (1 << rq) != jbit(op.tpos2)) std::abort();
            if (((op.fixpos >> 36) & 1u) && (op.op != B2SV_OP_MAT1 || op.m[1] != 0.0 || op.m[3] != 0.0 || op.m[5] != 0.0 || op.m[7] != 0.0)) std::abort();
            for (int j = 0; j < 8; ++j) {
                if (!((ok >> j) & 1u)) continue;
                switch (op.op) {
                    case B2SV_OP_PHASE:
                        v[j] *= c_of(op.m, 0);
                        break;
                    case B2SV_OP_DIAG1:
                        if (op.tpos == 0xFF) v[j] *= (base & op.ext_tbit) ? c_of(op.m, 1) : c_of(op.m, 0);
                        else v[j] *= (j & jbit(op.tpos)) ? c_of(op.m, 1) : c_of(op.m, 0);
                        break;
                    case B2SV_OP_X: {
                        const int tb = jbit(op.tpos);
                        if (!(j & tb)) std::swap(v[j], v[j | tb]);
                    } break;
                    case B2SV_OP_SWAP: {
                        const int ta = jbit(op.tpos), tb = jbit(op.tpos2);
                        if ((j & ta) && !(j & tb)) std::swap(v[j], v[j ^ ta ^ tb]);
                    } break;
                    case B2SV_OP_MAT1: {
                        const int tb = jbit(op.tpos);
                        if (!(j & tb)) {
                            const cplx a = v[j], b = v[j | tb];
                            v[j] = c_of(op.m, 0) * a + c_of(op.m, 1) * b;
                            v[j | tb] = c_of(op.m, 2) * a + c_of(op.m, 3) * b;
                        }
                    } break;
                    default:
                        break;
                }
            }
        }
        for (int j = 0; j < 8; ++j) tile[off[j]] = v[j];
    }
}

// mirrors tile_apply_mono: table look-up of the per-thread walk results, rounds of kMonoElems amplitudes per
// thread, all reads of a round before its writes
int run_mono_phase(std::vector<cplx>& tile, const TileOp& hdr, int k, uint64_t base, uint32_t slice, const Plan& plan, Counters& c) {
    const int E = kMonoElems;
    const int V = hdr.mono.fac ? kMonoVFac : kMonoV;
    const uint32_t tmask = hdr.mono.tmask, qlmask = hdr.mono.qlmask, nmask = hdr.mono.nmask;
    const int tb = __builtin_popcount(tmask), lb = hdr.mono.lb, vbits = hdr.mono.vbits;
    if (tb + lb != k || (1 << vbits) != V || lb < 3 || tb != kMonoThreadBits || hdr.mono.count != 0) return -1;
    if (hdr.mono.ebits > kMonoMaxEbits || slice >= (1u << hdr.mono.ebits)) return -9;
    // thread, walk and untouched loop positions partition the tile positions
    if ((tmask | qlmask | nmask) != (1u << k) - 1u || (tmask & qlmask) || (tmask & nmask) || (qlmask & nmask)) return -5;
    if (__builtin_popcount(qlmask) != vbits) return -5;
    const uint32_t n_threads = 1u << tb, per_thread = 1u << lb;
    const uint32_t rounds = per_thread / E;
    if ((int64_t)rounds > c.mono_rounds_max) c.mono_rounds_max = rounds;
    const double2* tables = reinterpret_cast<const double2*>(plan.mux_tables.data());
    std::vector<uint32_t> qsrc((size_t)n_threads * V), qx((size_t)n_threads * V);
    std::vector<cplx> fac((size_t)n_threads * V, cplx(1.0, 0.0));
    for (uint32_t t = 0; t < n_threads; ++t) {
        const uint32_t tpart = mono_deposit(t, tmask);
        const size_t first = (((size_t)slice << tb) + t) * V;
        if ((hdr.mono.mask_off * 16 + (first + V) * 2) > plan.mux_tables.size() * 8) return -11;
        const uint16_t* xm = reinterpret_cast<const uint16_t*>(tables + hdr.mono.mask_off) + first;
        uint32_t q = 0;
        for (int v = 0; v < V; ++v) {
            qsrc[(size_t)t * V + v] = tpart | q;
            q = mono_next_subset(q, qlmask);
            qx[(size_t)t * V + v] = xm[v];
            if (hdr.mono.fac) {
                if ((hdr.mono.fac_off + first + V) * 2 > plan.mux_tables.size()) return -12;
                const double2 fv = (tables + hdr.mono.fac_off + first)[v];
                fac[(size_t)t * V + v] = cplx(fv.x, fv.y);
            }
        }
    }
    std::vector<char> written(tile.size(), 0);
    std::vector<uint32_t> dst;
    std::vector<cplx> val;
    std::vector<uint32_t> ncur(n_threads, 0);
    for (uint32_t j0 = 0; j0 < per_thread; j0 += E) {
        dst.clear();
        val.clear();
        std::vector<char> in_round(tile.size(), 0);
        for (uint32_t t = 0; t < n_threads; ++t)
            for (int e = 0; e < E; ++e) {
                const int v = e & (V - 1);
                const uint32_t src = qsrc[(size_t)t * V + v] | ncur[t];
                if (src >= tile.size() || in_round[src]) return -6;  // every amplitude is read exactly once
                in_round[src] = 1;
                dst.push_back(src ^ qx[(size_t)t * V + v]);
                val.push_back(hdr.mono.fac ? fac[(size_t)t * V + v] * tile[src] : tile[src]);
                if (v == V - 1) ncur[t] = mono_next_subset(ncur[t], nmask);
            }
        for (size_t q = 0; q < dst.size(); ++q) {
            if (dst[q] >= tile.size() || written[dst[q]]) return -2;  // a bijection
            if (!in_round[dst[q]]) return -3;                          // the round is closed under the run
            written[dst[q]] = 1;
            tile[dst[q]] = val[q];
        }
    }
    for (char w : written)
        if (!w) return -7;
    return 0;
}

int run_program(std::vector<cplx>& tile, const TileOp* ops, int n_ops, int k, uint64_t base, const Plan& plan, Counters& c) {
    for (int b0 = 0; b0 < n_ops; b0 += kTileMaxOps) {
        const int nb = n_ops - b0 < kTileMaxOps ? n_ops - b0 : kTileMaxOps;
        const TileOp* s_ops = ops + b0;
        for (int o = 0; o < nb; ++o) {
            const TileOp& op = s_ops[o];
            if (op.op == kOpMono) {
                uint32_t slice = 0;
                for (int i = 0; i < op.mono.ebits; ++i) slice |= (uint32_t)((base >> op.mono.epos[i]) & 1ull) << i;
                if (!((op.mono.identity >> slice) & 1u))
                    if (int rc = run_mono_phase(tile, op, k, base, slice, plan, c)) return rc;
                continue;
            }
            if (op.op == kOpPhase) {
                if (o + op.ph.count >= nb) return -10;  // a phase straddles a batch boundary
                bool any = false;
                for (int mi = 1; mi <= op.ph.count; ++mi) any |= (base & s_ops[o + mi].ext_ctrl) == s_ops[o + mi].ext_ctrl;
                if (any) run_reg_phase(tile, op, s_ops + o + 1, k, base);
                o += op.ph.count;
                continue;
            }
            if (op.op == kOpNop) continue;
            run_op(tile, op, k, base, plan);
        }
    }
    return 0;
}

}  // namespace

extern "C" int emul_apply(int n, int dtype, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts, double* amps,
                          int64_t* info /* 12 counters */) {
    PlanConfig cfg = default_config(n, dtype, opts);
    if (n < 13) cfg.perm_runs = false;  // as b2sv_apply
    Plan plan = make_plan(gates, n_gates, cfg, true);
    const uint64_t N = 1ull << n;
    std::vector<cplx> psi(N);
    for (uint64_t i = 0; i < N; ++i) psi[i] = cplx(amps[2 * i], amps[2 * i + 1]);
    Counters c;
    for (const Segment& seg : plan.segs) {
        if (seg.cls == SEG_GATE1Q || seg.cls == SEG_PERM) {
            (seg.cls == SEG_PERM ? c.perm : c.gate1q) += 1;
            for (int gi : seg.gates) {
                const PGate& g = plan.gates[gi];
                if (g.op > B2SV_OP_PHASE) return -20;  // fused ops only ride in tile sweeps
                apply_plain(psi, n, g);
            }
            continue;
        }
        (seg.cls == SEG_TILE ? c.tile : c.low6) += 1;
        std::vector<TileOp> ops;
        build_segment_ops(plan, seg, cfg, ops);
        if (seg.cls == SEG_LOW6 && (int)ops.size() > kTileMaxOps) return -21;
        c.nops += (int64_t)ops.size();
        if ((int64_t)ops.size() > c.max_ops) c.max_ops = (int64_t)ops.size();
        for (size_t o = 0; o < ops.size(); ++o) {
            if (ops[o].op == kOpMono) {
                c.mono_phases++;
                c.mono_members += ops[o].ph.count;
            } else if (ops[o].op == kOpPhase) {
                c.reg_phases++;
                c.reg_members += ops[o].ph.count;
            }
        }
        ExpandArgs tex, cex;
        fill_expand(tex, seg.tile_mask);
        fill_expand(cex, seg.tile_mask | seg.common_ctrl);
        const int k = tex.m;
        const uint64_t n_tiles = 1ull << (n - cex.m);
        std::vector<cplx> tile(1u << k);
        for (uint64_t t = 0; t < n_tiles; ++t) {
            uint64_t base = t;
            for (int i = 0; i < cex.m; ++i) base = insert_zero(base, cex.pos[i]);
            base |= seg.common_ctrl;
            auto global_of = [&](uint32_t l) {
                uint64_t g = base;
                for (int i = 0; i < k; ++i)
                    if ((l >> i) & 1) g |= 1ull << tex.pos[i];
                return g;
            };
            for (uint32_t l = 0; l < (1u << k); ++l) tile[l] = psi[global_of(l)];
            if (int rc = run_program(tile, ops.data(), (int)ops.size(), k, base, plan, c)) return rc;
            for (uint32_t l = 0; l < (1u << k); ++l) psi[global_of(l)] = tile[l];
        }
    }
    const cplx s(plan.scalar_re, plan.scalar_im);
    for (uint64_t i = 0; i < N; ++i) {
        const cplx v = psi[i] * s;
        amps[2 * i] = v.real();
        amps[2 * i + 1] = v.imag();
    }
    if (info) {
        const int64_t out[12] = {c.tile, c.perm, c.gate1q, c.low6, c.mono_phases, c.mono_members, c.reg_phases, c.reg_members, c.max_ops, c.nops, c.mono_rounds_max, c.mono_gates};
        std::memcpy(info, out, sizeof(out));
    }
    return 0;
}

// Introspection for profiling scripts: the op program of tile segment `seg_index` as (kind, detail) pairs --
// detail = members of a register phase, block qubits, multiplexor index bits, local controls of a plain op.
extern "C" int emul_describe_segment(int n, int dtype, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts, int seg_index,
                                     int32_t* out, int cap) {
    PlanConfig cfg = default_config(n, dtype, opts);
    if (n < 13) cfg.perm_runs = false;
    Plan plan = make_plan(gates, n_gates, cfg, true);
    if (seg_index < 0 || seg_index >= (int)plan.segs.size()) return -1;
    const Segment& seg = plan.segs[seg_index];
    if (seg.cls != SEG_TILE && seg.cls != SEG_LOW6) return 0;
    std::vector<TileOp> ops;
    build_segment_ops(plan, seg, cfg, ops);
    int m = 0;
    for (size_t o = 0; o < ops.size() && 2 * m + 1 < cap; ++o) {
        const TileOp& op = ops[o];
        int detail = 0;
        if (op.op == kOpPhase) detail = op.ph.count;
        else if (op.op == kOpMono) detail = op.mono.fac * 100 + op.mono.ebits;
        else if (op.op == kOpBlock) detail = op.blk.kb;
        else if (op.op == kOpMux) detail = op.mux.nmux;
        else detail = __builtin_popcount(op.lctrl);
        out[2 * m] = op.op;
        out[2 * m + 1] = detail;
        ++m;
        if (op.op == kOpPhase) o += op.ph.count;
    }
    return m;
}

// Plan + build every segment's op program (what b2sv_apply does on the host before the first launch);
// returns {segments, ops, table doubles} -- used to time the host side.
extern "C" int emul_build_all(int n, int dtype, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts, int64_t* out3) {
    PlanConfig cfg = default_config(n, dtype, opts);
    if (n < 13) cfg.perm_runs = false;
    Plan plan = make_plan(gates, n_gates, cfg, true);
    std::vector<TileOp> ops;
    for (const Segment& seg : plan.segs)
        if (seg.cls == SEG_TILE || seg.cls == SEG_LOW6) build_segment_ops(plan, seg, cfg, ops);
    out3[0] = (int64_t)plan.segs.size();
    out3[1] = (int64_t)ops.size();
    out3[2] = (int64_t)plan.mux_tables.size();
    return 0;
}

// Introspection for profiling scripts: every dense block of the plan as {kb, n_controls, real, real_up_to_phase}.
extern "C" int emul_block_stats(int n, int dtype, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts, int32_t* out, int cap) {
    PlanConfig cfg = default_config(n, dtype, opts);
    if (n < 13) cfg.perm_runs = false;
    Plan plan = make_plan(gates, n_gates, cfg, true);
    int m = 0;
    for (const PGate& g : plan.gates) {
        if (g.op != B2SV_OP_BLOCK || 4 * m + 3 >= cap) continue;
        const BlockInfo& bi = plan.blocks[g.blk];
        const int dim = 1 << bi.kb;
        const double* M = plan.mux_tables.data() + 2 * bi.mat_off;
        bool real = true;
        double best = 0, pre = 1, pim = 0;
        for (int e = 0; e < dim * dim; ++e) {
            real = real && M[2 * e + 1] == 0.0;
            const double a = std::hypot(M[2 * e], M[2 * e + 1]);
            if (a > best) {
                best = a;
                pre = M[2 * e] / a;
                pim = M[2 * e + 1] / a;
            }
        }
        bool rup = true;  // real after dividing by the phase of the largest entry
        for (int e = 0; e < dim * dim; ++e) {
            const double im = M[2 * e + 1] * pre - M[2 * e] * pim;
            rup = rup && std::fabs(im) < 1e-14;
        }
        out[4 * m] = bi.kb;
        out[4 * m + 1] = __builtin_popcountll(g.ctrl);
        out[4 * m + 2] = real ? 1 : 0;
        out[4 * m + 3] = rup ? 1 : 0;
        ++m;
    }
    return m;
}

// Like emul_describe_segment, with the qubits each pass works on: {kind, detail, target-qubit mask lo, hi, control-qubit mask lo, hi}.
extern "C" int emul_describe_segment_qubits(int n, int dtype, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts, int seg_index,
                                            int64_t* out, int cap) {
    PlanConfig cfg = default_config(n, dtype, opts);
    if (n < 13) cfg.perm_runs = false;
    Plan plan = make_plan(gates, n_gates, cfg, true);
    if (seg_index < 0 || seg_index >= (int)plan.segs.size()) return -1;
    const Segment& seg = plan.segs[seg_index];
    if (seg.cls != SEG_TILE && seg.cls != SEG_LOW6) return 0;
    std::vector<TileOp> ops;
    build_segment_ops(plan, seg, cfg, ops);
    ExpandArgs ex;
    fill_expand(ex, seg.tile_mask);
    auto qmask_of = [&](uint32_t local) {
        uint64_t q = 0;
        for (int i = 0; i < ex.m; ++i)
            if ((local >> i) & 1u) q |= 1ull << ex.pos[i];
        return q;
    };
    auto targets_of = [&](const TileOp& op) -> uint32_t {
        uint32_t t = 0;
        if (op.op == kOpBlock) {
            for (int i = 0; i < op.blk.kb; ++i) t |= 1u << op.blk.bpos[i];
        } else if (op.op != B2SV_OP_PHASE && op.op <= kOpMux) {
            if (op.tpos != 0xFF) t |= 1u << op.tpos;
            if (op.op == B2SV_OP_SWAP) t |= 1u << op.tpos2;
        }
        return t;
    };
    int m = 0;
    for (size_t o = 0; o < ops.size() && 4 * m + 3 < cap; ++o) {
        const TileOp& op = ops[o];
        int detail = 0;
        uint64_t tq = 0, cq = 0;
        if (op.op == kOpPhase) {
            detail = op.ph.count;
            for (int i = 0; i < 3; ++i) tq |= qmask_of(1u << op.ph.rpos[i]);
            cq = qmask_of(op.lctrl);
        } else if (op.op == kOpMono) {
            detail = op.mono.fac * 100 + op.mono.ebits;
            tq = qmask_of(op.mono.tmask | op.mono.qlmask);
        } else {
            detail = op.op == kOpBlock ? op.blk.kb : (op.op == kOpMux ? op.mux.nmux : __builtin_popcount(op.lctrl));
            tq = qmask_of(targets_of(op));
            cq = qmask_of(op.lctrl) | op.ext_ctrl;
        }
        out[4 * m] = op.op;
        out[4 * m + 1] = detail;
        out[4 * m + 2] = (int64_t)tq;
        out[4 * m + 3] = (int64_t)cq;
        ++m;
        if (op.op == kOpPhase) o += op.ph.count;
    }
    return m;
}
