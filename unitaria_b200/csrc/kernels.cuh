// kernels.cuh -- hand-written sm_100a kernels of the B200 statevector engine.
//
// Every kernel here works on amplitudes in LSB order (bit q of the index = qubit q), the layout
// Circuit.simulate returns (/root/reference/src/unitaria/circuit.py:63-70).  All of them are
// HBM-bound integer/indexing + complex-FMA work (<= 14 flops per amplitude per gate); none is
// GEMM-shaped, so none uses the tensor cores.  The levers are the ones that matter for such
// kernels on B200: 16-byte accesses, long contiguous runs per request (bulk-async / TMA copies of
// whole tile chunks), grids that are multiples of the SM count, and as few sweeps as possible.
#pragma once

#include <cuda.h>  // CUtensorMap (type only: the driver entry point is resolved at run time)
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2sv.h"

// -DB2SV_BOUNDS_CHECK (python -m unitaria_b200.build --bounds -> libunitaria_b200_dbg.so): every shared-memory tile access
// and the global accesses whose index is computed from gate data are range-checked on the device and trap with a message.
// compute-sanitizer is closed on the B200 pool this engine is developed on (profiles/sanitizer_r2.txt); this build plus
// the oracle comparison of scripts/sanitize_cases.py is the memcheck stand-in.
#ifdef B2SV_BOUNDS_CHECK
#include <cstdio>
#define B2SV_CHECK(cond)                                                                      \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            printf("b2sv bounds check failed: %s (kernels.cuh:%d, block %d thread %d)\n", #cond, __LINE__, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                         \
            __trap();                                                                         \
        }                                                                                     \
    } while (0)
#else
#define B2SV_CHECK(cond) \
    do {                 \
    } while (0)
#endif

namespace b2svk {

constexpr int kMaxFixed = 40;
constexpr int kThreads = 256;

struct ExpandArgs {
    int m;
    uint8_t pos[kMaxFixed];  // ascending bit positions at which a 0 is inserted
};

__host__ __device__ __forceinline__ uint64_t insert_zero(uint64_t k, int p) {
    const uint64_t lo = k & ((1ull << p) - 1);
    return ((k >> p) << (p + 1)) | lo;
}

__device__ __forceinline__ uint64_t expand_bits(uint64_t k, const ExpandArgs& e) {
    for (int i = 0; i < e.m; ++i) k = insert_zero(k, e.pos[i]);
    return k;
}

// ---- amplitude access: storage type A (double2 = complex128, float2 = complex64), math in fp64 ----

__device__ __forceinline__ double2 to_c128(double2 v) { return v; }
__device__ __forceinline__ double2 to_c128(float2 v) { return make_double2((double)v.x, (double)v.y); }
template <typename A>
__device__ __forceinline__ A from_c128(double2 v);
template <>
__device__ __forceinline__ double2 from_c128<double2>(double2 v) {
    return v;
}
template <>
__device__ __forceinline__ float2 from_c128<float2>(double2 v) {
    return make_float2((float)v.x, (float)v.y);
}
__host__ __device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {  // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

// Compute type of the tile sweeps: the storage type itself.  complex64 registers are swept in fp32 arithmetic: the fp64
// round trip (F2F conversions run at 16 per clock per SM, 4 per amplitude per shared-memory pass) made the complex64
// sweeps of config 3 conversion-bound at 3.4x their HBM time, and every pass rounds to fp32 when it stores anyway.
// Matrices are fused and multiplied out in fp64 on the host and rounded once.
template <typename A>
struct Num;
template <>
struct Num<double2> {
    using C = double2;
    using R = double;
    static __device__ __forceinline__ C mk(double re, double im) { return make_double2(re, im); }
};
template <>
struct Num<float2> {
    using C = float2;
    using R = float;
    static __device__ __forceinline__ C mk(float re, float im) { return make_float2(re, im); }
};
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 cfma(float2 a, float2 b, float2 c) {
    return make_float2(fmaf(a.x, b.x, fmaf(-a.y, b.y, c.x)), fmaf(a.x, b.y, fmaf(a.y, b.x, c.y)));
}
__device__ __forceinline__ float2 narrow(double2 v, float2*) { return make_float2((float)v.x, (float)v.y); }
__device__ __forceinline__ double2 narrow(double2 v, double2*) { return v; }

// ===================================================================================================
// 1. Stand-alone multi-controlled gate: visits only the 2^(n-c[-1]) amplitudes selected by the
//    control bits (the schedule of one qulacs update, but with all controls handled in the index).
// ===================================================================================================

struct Gate1Args {
    ExpandArgs ex;      // fixed positions = controls + target(s)
    uint64_t ctrl;      // OR-ed into every expanded index
    uint64_t bit0, bit1;
    uint64_t count;     // number of (pairs of) amplitudes to visit
    uint64_t total;     // amplitudes of the register (range checks)
    int op;
    double m[8];
};

template <typename A>
__global__ void __launch_bounds__(kThreads) k_gate1(A* __restrict__ psi, const Gate1Args a) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.count; k += stride) {
        const uint64_t i0 = expand_bits(k, a.ex) | a.ctrl;
        B2SV_CHECK((i0 | a.bit0 | a.bit1) < a.total);
        switch (a.op) {
            case B2SV_OP_X: {
                const uint64_t i1 = i0 | a.bit0;
                const A v0 = psi[i0], v1 = psi[i1];
                psi[i0] = v1;
                psi[i1] = v0;
            } break;
            case B2SV_OP_SWAP: {
                const uint64_t ia = i0 | a.bit0, ib = i0 | a.bit1;
                const A va = psi[ia], vb = psi[ib];
                psi[ia] = vb;
                psi[ib] = va;
            } break;
            case B2SV_OP_MAT1: {
                const uint64_t i1 = i0 | a.bit0;
                const double2 v0 = to_c128(psi[i0]), v1 = to_c128(psi[i1]);
                const double2 m00 = make_double2(a.m[0], a.m[1]), m01 = make_double2(a.m[2], a.m[3]);
                const double2 m10 = make_double2(a.m[4], a.m[5]), m11 = make_double2(a.m[6], a.m[7]);
                psi[i0] = from_c128<A>(cfma(m01, v1, cmul(m00, v0)));
                psi[i1] = from_c128<A>(cfma(m11, v1, cmul(m10, v0)));
            } break;
            case B2SV_OP_DIAG1: {
                const uint64_t i1 = i0 | a.bit0;
                const double2 d0 = make_double2(a.m[0], a.m[1]), d1 = make_double2(a.m[2], a.m[3]);
                if (!(d0.x == 1.0 && d0.y == 0.0)) psi[i0] = from_c128<A>(cmul(d0, to_c128(psi[i0])));
                if (!(d1.x == 1.0 && d1.y == 0.0)) psi[i1] = from_c128<A>(cmul(d1, to_c128(psi[i1])));
            } break;
            default: {  // B2SV_OP_PHASE
                const double2 w = make_double2(a.m[0], a.m[1]);
                psi[i0] = from_c128<A>(cmul(w, to_c128(psi[i0])));
            } break;
        }
    }
}

// ===================================================================================================
// 2. Fused tile sweep.  A CTA owns one tile = all 2^k combinations of the tile qubits T for one
//    setting of the other n-k qubits.  T always contains the low `lc` qubits, so the tile is
//    2^(k-lc) contiguous chunks of 2^lc amplitudes; each chunk moves with ONE bulk-async (TMA engine)
//    copy global->shared, completion tracked by an mbarrier, and one bulk copy back.  In between,
//    the gates of the sweep are applied to the tile in shared memory: targets are tile-local bit
//    positions, controls inside the tile are per-element predicates, controls outside the tile are
//    CTA-uniform (the whole CTA skips the gate).  Algorithmic traffic: 2 * sizeof(A) * 2^n per sweep.
// ===================================================================================================

struct __align__(16) TileOp {
    uint8_t op;
    uint8_t tpos, tpos2;  // tile-local target positions (0xFF: DIAG1 target outside the tile)
    uint8_t nfix;         // number of fixed tile-local positions while enumerating
    uint32_t lctrl;       // tile-local control mask
    uint64_t fixpos;      // nfix ascending tile-local positions, 4 bits each
    uint64_t ext_ctrl;    // controls outside the tile (global index bits)
    uint64_t ext_tbit;    // DIAG1 whose target is outside the tile: that global bit
    union {
        double m[8];
        struct {               // op == kOpMux: table of 2x2 matrices indexed by control bits
            uint64_t table_off;    // first entry (4 double2 per entry) in TileArgs::tables
            uint8_t nmux;
            uint8_t mpos[7];       // per index bit: tile-local position, or 0xFF when outside the tile
            uint8_t mq[8];         // per index bit: global qubit
            // enumeration that keeps the table index uniform over long stretches of consecutive pairs: the index bits
            // that lie inside the tile are enumerated LAST (as the top bits of the pair counter)
            uint64_t mfixpos;      // fixpos plus the tile-local positions of the index bits (ascending, 4 bits each)
            uint8_t mnfix;         // entries of mfixpos
            uint8_t nl;            // index bits inside the tile
            uint8_t pad_[2];
            uint32_t lppack;       // nibble i: tile-local position of the i-th inside index bit
            uint32_t ljpack;       // nibble i: which index bit (j) that is
        } mux;
        struct {               // op == kOpBlock: dense 2^kb x 2^kb matrix on kb <= 3 tile qubits
            uint64_t mat_off;      // first double2 of the row-major matrix in TileArgs::tables
            uint8_t kb;
            uint8_t bpos[3];       // tile-local positions of the block qubits (matrix index bit i)
            uint8_t slot;          // first 256-byte granule of the matrix in the launch's BlkSlots, or 0xFF
            uint8_t real;          // 1: every entry of the matrix is real (half the FP64 work)
        } blk;
        struct {               // op == kOpPhase: header of a register phase, `count` member ops follow
            uint16_t count;
            uint8_t rpos[3];       // the three tile-local positions held in registers (ascending)
        } ph;
        struct {               // op == kOpMono: a monomial phase (no member ops follow: count == 0)
            uint16_t count;        // 0 (same place as ph.count)
            uint8_t lb;            // a thread owns 2^lb amplitudes (tile-local index = thread bits + loop bits)
            uint8_t vbits;         // 2^vbits index walks per thread: 3 (pure permutation) or 2 (with factors)
            uint8_t fac;           // 1: the run has PHASE / DIAG1 members (amplitudes pick up factors)
            uint8_t ebits;         // qubits outside the tile that members test (beyond the sweep's common
            uint8_t pad[2];        //   controls): one table slice per value of those bits
            uint32_t tmask;        // tile-local positions that take the thread number (ascending)
            uint32_t qlmask;       // loop positions that number a thread's walks (all touched loop positions are here)
            uint32_t nmask;        // the other loop positions (not touched by the run)
            uint8_t epos[4];       // global qubit numbers of the `ebits` outside bits
            uint32_t identity;     // bit p set: slice p moves nothing and multiplies nothing -- skip the phase
            uint64_t mask_off;     // first uint16 XOR mask, counted in double2 units from the start of the tables
            uint64_t fac_off;      // first double2 factor (fac == 1), same units
        } mono;
    };
};
constexpr int kOpMux = 5;
constexpr int kOpBlock = 6;
constexpr int kOpPhase = 7;
constexpr int kOpMono = 8;   // header of a monomial phase (index moves + phase factors only)
constexpr int kOpNop = 9;    // padding in front of a batch boundary of a long op list
constexpr int kMonoElems = 8;              // amplitudes a thread carries through one round of a monomial phase
constexpr int kMonoThreadBits = 8;         // log2(kThreads)
constexpr int kMonoV = 8, kMonoVFac = 4;   // index walks per thread (pure permutation / with factors)
constexpr int kMonoMaxEbits = 4;           // table-driven phases: at most 16 slices
static_assert(sizeof(TileOp) == 96, "TileOp layout");
// Entry i of a STAGED plain op's m[]: the complex64 kernels narrow the eight doubles to floats in place while staging
// (tile_narrow_ops), so the sweep itself never converts.
template <typename A>
__device__ __forceinline__ typename Num<A>::R mval(const double* m, int i) {
    if constexpr (sizeof(A) == 16) return m[i];
    else return reinterpret_cast<const float*>(m)[i];
}

struct TileArgs {
    ExpandArgs cta_ex;    // tile qubits + the sweep's common outside controls: block index -> tile base
    uint64_t cta_or;      // the common outside controls (set in every launched tile's base)
    ExpandArgs tile_ex;   // the k tile-qubit positions (ascending)
    int k, lc;
    int n_ops;
    const TileOp* ops;
    const double2* tables;  // multiplexor tables of this apply
    // init_mode 1: the register is the (not yet materialised) basis state |init_index>: the tile is
    // generated in shared memory instead of being read -- the first sweep after b2sv_set_basis costs
    // one write of the state and no read, and the separate initialisation kernel disappears.
    int init_mode;
    uint64_t init_index;
    uint64_t tile_mask;
    // phase trace (profiling aid, B2SV_TRACE): per CTA {smid, t_start, t_loaded, t_computed, t_stored} in ns (%globaltimer)
    unsigned long long* trace;
};

__host__ __device__ __forceinline__ uint32_t expand_local(uint32_t p, uint64_t fixpos, int nfix) {
    for (int i = 0; i < nfix; ++i) {
        const int pos = (int)((fixpos >> (4 * i)) & 15);
        const uint32_t lo = p & ((1u << pos) - 1);
        p = ((p >> pos) << (pos + 1)) | lo;
    }
    return p;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "B2SV_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra B2SV_DONE_%=;\n"
        "bra B2SV_WAIT_%=;\n"
        "B2SV_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}

constexpr int kTileMaxOps = 112;  // ops of one sweep are staged in shared memory (10.5 KiB)

// View of a tile in shared memory.  SWZ = true: the tile was written by TMA tensor copies with the
// hardware 128-byte swizzle (16-byte chunk index XOR row index mod 8), so element l lives at
// l ^ f(l).  The swizzle is what makes the access patterns of gates on the LOW tile qubits -- fixed low
// bits, i.e. every lane in the same 16-byte column of its 128-byte row -- spread over all banks
// (ncu on Pseudoinverse(BPX) sweeps: 69 % of the shared-memory wavefronts were bank conflicts without it).
template <typename A, bool SWZ>
struct TileView {
    A* p;
    uint32_t n;  // elements of the tile (range checks of the -DB2SV_BOUNDS_CHECK build)
    __device__ __forceinline__ A& operator[](uint32_t l) const {
        B2SV_CHECK(l < n);
        if (SWZ) {
            if (sizeof(A) == 16) l ^= (l >> 3) & 7u;
            else l ^= ((l >> 4) & 7u) << 1;
        }
        return p[l];
    }
    // The swizzle is linear over GF(2): swz(l ^ b) = swz(l) ^ swz(b).  An op therefore swizzles an element index once
    // and reaches its partners (l | target bit, the 2^kb corners of a block, the eight amplitudes of a register group)
    // with ONE XOR each against a per-op constant, instead of re-deriving every address (~6 instructions apiece in
    // passes that are instruction-issue bound).
    __device__ __forceinline__ uint32_t swz(uint32_t l) const {
        if (SWZ) {
            if (sizeof(A) == 16) l ^= (l >> 3) & 7u;
            else l ^= ((l >> 4) & 7u) << 1;
        }
        return l;
    }
    __device__ __forceinline__ A& raw(uint32_t s) const {  // s = swz(l)
        B2SV_CHECK(s < n);
        return p[s];
    }
};

// Enumerate the tile-local indices with the fixed positions cleared, UNROLL at a time, so that a
// thread has UNROLL independent shared-memory round trips in flight (the tile phase is latency-
// bound otherwise: 8 warps per CTA, a dependent LDS -> DFMA -> STS chain per element).
// Element p = tid + i * nt of the enumeration (nt = CTA size, a power of two) has the bits of tid in the lowest free
// positions and the bits of i in the free positions above them, so the deposit of tid is done ONCE per op and the
// i part is stepped as "next subset of a mask" (3 instructions) -- depositing every p bit by bit was ~10
// instructions per fixed position per element, a quarter of the instructions of a dense-block pass.
__device__ __forceinline__ uint32_t subset_next(uint32_t cur, uint32_t mask) { return ((cur | ~mask) + 1u) & mask; }

// Bank-conflict-free lane assignment.  The lanes of one shared-memory phase (8 lanes for 16-byte accesses, 16 for 8-byte
// ones) differ in the LOWEST bits of the thread number, which the enumeration deposits into the lowest free tile
// positions.  With the 128-byte swizzle a position p < 6 moves the 16-byte bank group by bit (p mod 3): positions 0 and 3
// (1 and 4, 2 and 5) collide, so free positions {0, 2, 3} under the three lowest lane bits give a 2-way conflict on every
// access of the pass (ncu on the 8x8 block passes of BPX: twice the ideal shared-memory wavefronts, L1 data pipe 85 %
// busy).  The host picks, per op, free positions with DISTINCT bank-group bits for those lane bits (tileops.hpp:
// assign_lane_swaps) and passes the choice as up to four bit swaps of the thread number (6 bits each: a, b).
__device__ __forceinline__ uint32_t lane_swaps(uint32_t t, uint32_t code) {
    if (code == 0u) return t;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const uint32_t a = (code >> (6 * s)) & 7u, b = (code >> (6 * s + 3)) & 7u;
        const uint32_t d = ((t >> a) ^ (t >> b)) & 1u;
        t ^= (d << a) | (d << b);
    }
    return t;
}

// The body gets SWIZZLED element indices (TileView::raw).
template <int UNROLL, typename TV, typename F>
__device__ __forceinline__ void tile_for_each(TV tile, uint32_t cnt, uint64_t fixpos, int nfix, uint32_t lctrl, uint32_t swaps, F&& body) {
    const uint32_t tid = lane_swaps(threadIdx.x, swaps), nt = blockDim.x;
    if (tid >= cnt) return;
    const uint32_t lt = tile.swz(expand_local(tid, fixpos, nfix) | lctrl);
    const uint32_t himask = cnt > nt ? expand_local((cnt - 1u) & ~(nt - 1u), fixpos, nfix) : 0u;
    const uint32_t iters = cnt > nt ? cnt / nt : 1u;  // a power of two
    uint32_t hi = 0, i = 0;
    uint32_t l[UNROLL];
#pragma unroll 1
    for (; i + UNROLL <= iters; i += UNROLL) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            l[u] = lt ^ tile.swz(hi);
            hi = subset_next(hi, himask);
        }
        body(l, UNROLL);
    }
#pragma unroll 1
    for (; i < iters; ++i) {
        l[0] = lt ^ tile.swz(hi);
        hi = subset_next(hi, himask);
        body(l, 1);
    }
}

// The bit-by-bit form (every element deposited on its own): kept for the dense-block passes, whose bodies sit at the
// register limit -- the three extra live values of the stepped form made ptxas spill 1.2 KB per thread there.
template <int UNROLL, typename TV, typename F>
__device__ __forceinline__ void tile_for_each_deposit(TV tile, uint32_t cnt, uint64_t fixpos, int nfix, uint32_t lctrl, uint32_t swaps, F&& body) {
    uint32_t p = lane_swaps(threadIdx.x, swaps);
    const uint32_t nt = blockDim.x;
    for (; p + (UNROLL - 1) * nt < cnt; p += UNROLL * nt) {
        uint32_t l[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) l[u] = tile.swz(expand_local(p + u * nt, fixpos, nfix) | lctrl);
        body(l, UNROLL);
    }
    for (; p < cnt; p += nt) {
        uint32_t l[UNROLL];
        l[0] = tile.swz(expand_local(p, fixpos, nfix) | lctrl);
        body(l, 1);
    }
}

// Dense-block matrices of a sweep travel as KERNEL PARAMETERS (constant bank 0): with the slot number a compile-time
// constant, every matrix entry is an immediate constant-bank operand of the DFMA that uses it -- no load instruction
// at all.  (Round 1 read the matrix with __ldg inside the FMA loop: ncu on the heaviest BPX sweep showed 58 % of the
// stall samples of those DFMAs waiting on that load (long scoreboard), the L1 data pipe at 93 % with a quarter of its
// wavefronts being matrix loads, and the FP64 pipe 25 % busy.)  A granule is 16 double2 = one 4x4 matrix; an 8x8
// matrix takes four consecutive granules.
constexpr int kBlkGranules = 16;
constexpr int kBlkGranuleElems = 16;
struct BlkSlots {
    double2 m[kBlkGranules * kBlkGranuleElems];  // 4 KiB (complex64 sweeps: the host stores float2 entries, same numbering)
};

template <bool REAL, typename C>
__device__ __forceinline__ C bmul(C m, C v) {
    if (REAL) {
        C r;
        r.x = m.x * v.x;
        r.y = m.x * v.y;
        return r;
    }
    return cmul(m, v);
}
template <bool REAL, typename C>
__device__ __forceinline__ C bfma(C m, C v, C c) {
    if (REAL) {
        C r;
        r.x = fma(m.x, v.x, c.x);
        r.y = fma(m.x, v.y, c.y);
        return r;
    }
    return cfma(m, v, c);
}

// One dense block pass; U(i) yields entry i of the row-major matrix.
template <typename A, int KB, bool REAL, typename TV, typename MAT>
__device__ __forceinline__ void tile_block_pass(TV tile, const TileOp& op, int k, MAT U) {
    using C = typename Num<A>::C;
    const uint32_t cnt = 1u << (k - op.nfix);
    if (KB == 3) {
        const uint32_t b0 = tile.swz(1u << op.blk.bpos[0]), b1 = tile.swz(1u << op.blk.bpos[1]), b2 = tile.swz(1u << op.blk.bpos[2]);
        tile_for_each_deposit<1>(tile, cnt, op.fixpos, op.nfix, op.lctrl, (uint32_t)op.ext_tbit, [&](const uint32_t* l, int) {
            uint32_t off[8];
            C v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                off[j] = l[0] ^ ((j & 1) ? b0 : 0u) ^ ((j & 2) ? b1 : 0u) ^ ((j & 4) ? b2 : 0u);
                v[j] = tile.raw(off[j]);
            }
            // four rows at a time: four independent accumulation chains hide the DFMA latency
#pragma unroll
            for (int r0 = 0; r0 < 8; r0 += 4) {
                C acc[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = bmul<REAL>(U((r0 + r) * 8), v[0]);
#pragma unroll
                for (int c = 1; c < 8; ++c)
#pragma unroll
                    for (int r = 0; r < 4; ++r) acc[r] = bfma<REAL>(U((r0 + r) * 8 + c), v[c], acc[r]);
#pragma unroll
                for (int r = 0; r < 4; ++r) tile.raw(off[r0 + r]) = acc[r];
            }
        });
    } else {  // KB == 2
        const uint32_t b0 = tile.swz(1u << op.blk.bpos[0]), b1 = tile.swz(1u << op.blk.bpos[1]);
        tile_for_each_deposit<1>(tile, cnt, op.fixpos, op.nfix, op.lctrl, (uint32_t)op.ext_tbit, [&](const uint32_t* l, int) {
            uint32_t off[4];
            C v[4], acc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                off[j] = l[0] ^ ((j & 1) ? b0 : 0u) ^ ((j & 2) ? b1 : 0u);
                v[j] = tile.raw(off[j]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = bmul<REAL>(U(r * 4), v[0]);
#pragma unroll
            for (int c = 1; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r] = bfma<REAL>(U(r * 4 + c), v[c], acc[r]);
#pragma unroll
            for (int r = 0; r < 4; ++r) tile.raw(off[r]) = acc[r];
        });
    }
}

template <typename A, int KB, int G, typename TV>
__device__ __forceinline__ void tile_block_slot(TV tile, const TileOp& op, int k, const BlkSlots* bs) {
    auto U = [bs](int i) { return reinterpret_cast<const typename Num<A>::C*>(bs->m)[G * kBlkGranuleElems + i]; };  // G and i are compile-time: an immediate operand
    if (op.blk.real) tile_block_pass<A, KB, true>(tile, op, k, U);
    else tile_block_pass<A, KB, false>(tile, op, k, U);
}

template <typename A, typename TV>
__device__ __forceinline__ void tile_apply_block(TV tile, const TileOp& op, int k, const double2* __restrict__ tables, const BlkSlots* bs) {
    if (bs != nullptr && op.blk.slot != 0xFF) {
        if (op.blk.kb == 3) {
            switch (op.blk.slot) {
                case 0: tile_block_slot<A, 3, 0>(tile, op, k, bs); break;
                case 4: tile_block_slot<A, 3, 4>(tile, op, k, bs); break;
                case 8: tile_block_slot<A, 3, 8>(tile, op, k, bs); break;
                default: tile_block_slot<A, 3, 12>(tile, op, k, bs); break;
            }
        } else {
            switch (op.blk.slot) {
#define B2SV_CASE(G) case G: tile_block_slot<A, 2, G>(tile, op, k, bs); break;
                B2SV_CASE(0) B2SV_CASE(1) B2SV_CASE(2) B2SV_CASE(3) B2SV_CASE(4) B2SV_CASE(5) B2SV_CASE(6) B2SV_CASE(7)
                B2SV_CASE(8) B2SV_CASE(9) B2SV_CASE(10) B2SV_CASE(11) B2SV_CASE(12) B2SV_CASE(13) B2SV_CASE(14)
#undef B2SV_CASE
                default: tile_block_slot<A, 2, 15>(tile, op, k, bs); break;
            }
        }
        return;
    }
    // no slot left in this launch (or a kernel without the parameter block): the matrix comes through the read-only path
    const double2* __restrict__ Ug = tables + op.blk.mat_off;
    auto U = [Ug](int i) { return narrow(__ldg(Ug + i), (typename Num<A>::C*)nullptr); };
    if (op.blk.kb == 3) {
        if (op.blk.real) tile_block_pass<A, 3, true>(tile, op, k, U);
        else tile_block_pass<A, 3, false>(tile, op, k, U);
    } else {
        if (op.blk.real) tile_block_pass<A, 2, true>(tile, op, k, U);
        else tile_block_pass<A, 2, false>(tile, op, k, U);
    }
}

// FUSED = false compiles the light variant (no table-driven ops): fewer registers for sweeps that
// only hold plain gates.
template <typename A, bool FUSED, typename TV>
__device__ __forceinline__ void tile_apply_op(TV tile, const TileOp& op, int k, uint64_t base,
                                              const double2* __restrict__ tables, const BlkSlots* bs) {
    using C = typename Num<A>::C;
    using N = Num<A>;
    const uint32_t cnt = 1u << (k - op.nfix);
    const uint64_t fixpos = op.fixpos;
    const int nfix = op.nfix;
    const uint32_t lctrl = op.lctrl;
    // lane assignment of this op (a DIAG1 whose target lies outside the tile keeps that target's bit in ext_tbit instead)
    const uint32_t swaps = (op.op == B2SV_OP_DIAG1 && op.tpos == 0xFF) ? 0u : (uint32_t)op.ext_tbit;
    switch (op.op) {
        case B2SV_OP_X: {
            const uint32_t bt = tile.swz(1u << op.tpos);
            tile_for_each<4>(tile, cnt, fixpos, nfix, lctrl, swaps, [&](const uint32_t* l, int m) {
                A v0[4], v1[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u < m) {
                        v0[u] = tile.raw(l[u]);
                        v1[u] = tile.raw(l[u] ^ bt);
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u < m) {
                        tile.raw(l[u]) = v1[u];
                        tile.raw(l[u] ^ bt) = v0[u];
                    }
            });
        } break;
        case B2SV_OP_SWAP: {
            const uint32_t ba = tile.swz(1u << op.tpos), bb = tile.swz(1u << op.tpos2);
            tile_for_each<4>(tile, cnt, fixpos, nfix, lctrl, swaps, [&](const uint32_t* l, int m) {
                A va[4], vb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u < m) {
                        va[u] = tile.raw(l[u] ^ ba);
                        vb[u] = tile.raw(l[u] ^ bb);
                    }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u < m) {
                        tile.raw(l[u] ^ ba) = vb[u];
                        tile.raw(l[u] ^ bb) = va[u];
                    }
            });
        } break;
        case B2SV_OP_MAT1: {
            const uint32_t bt = tile.swz(1u << op.tpos);
            const C m00 = N::mk(mval<A>(op.m, 0), mval<A>(op.m, 1)), m01 = N::mk(mval<A>(op.m, 2), mval<A>(op.m, 3));
            const C m10 = N::mk(mval<A>(op.m, 4), mval<A>(op.m, 5)), m11 = N::mk(mval<A>(op.m, 6), mval<A>(op.m, 7));
            tile_for_each<2>(tile, cnt, fixpos, nfix, lctrl, swaps, [&](const uint32_t* l, int m) {
                C v0[2], v1[2];
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (u < m) {
                        v0[u] = tile.raw(l[u]);
                        v1[u] = tile.raw(l[u] ^ bt);
                    }
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (u < m) {
                        tile.raw(l[u]) = cfma(m01, v1[u], cmul(m00, v0[u]));
                        tile.raw(l[u] ^ bt) = cfma(m11, v1[u], cmul(m10, v0[u]));
                    }
            });
        } break;
        case B2SV_OP_DIAG1: {
            const C d0 = N::mk(mval<A>(op.m, 0), mval<A>(op.m, 1)), d1 = N::mk(mval<A>(op.m, 2), mval<A>(op.m, 3));
            if (op.tpos == 0xFF) {  // target outside the tile: one CTA-uniform factor
                const C d = (base & op.ext_tbit) ? d1 : d0;
                if (d.x == 1 && d.y == 0) break;
                tile_for_each<4>(tile, cnt, fixpos, nfix, lctrl, swaps, [&](const uint32_t* l, int m) {
                    C v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u < m) v[u] = tile.raw(l[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (u < m) tile.raw(l[u]) = cmul(d, v[u]);
                });
            } else {
                const uint32_t bt = tile.swz(1u << op.tpos);
                tile_for_each<2>(tile, cnt, fixpos, nfix, lctrl, swaps, [&](const uint32_t* l, int m) {
                    C v0[2], v1[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        if (u < m) {
                            v0[u] = tile.raw(l[u]);
                            v1[u] = tile.raw(l[u] ^ bt);
                        }
#pragma unroll
                    for (int u = 0; u < 2; ++u)
                        if (u < m) {
                            tile.raw(l[u]) = cmul(d0, v0[u]);
                            tile.raw(l[u] ^ bt) = cmul(d1, v1[u]);
                        }
                });
            }
        } break;
        case kOpMux: if constexpr (FUSED) {
            // one pass over all pairs of the target; the 2x2 matrix of a pair is looked up by the pair's control bits
            // (outside the tile: CTA-uniform; inside: the top bits of the pair counter, so a thread -- and for all but
            // tiny sub-ranges its whole warp -- keeps one matrix in registers over consecutive pairs.  Fetching the
            // matrix per pair made the complex64 state-preparation sweeps of cfg 3 L1-bound: ncu showed 67 % of the L1
            // data-pipe wavefronts being table reads and the pipe 97 % busy.)
            const uint32_t bt = tile.swz(1u << op.tpos);
            const int nmux = op.mux.nmux;
            uint32_t idx_ext = 0;
            for (int j = 0; j < nmux; ++j)
                if (op.mux.mpos[j] == 0xFF) idx_ext |= (uint32_t)((base >> op.mux.mq[j]) & 1ull) << j;
            const double2* __restrict__ tab = tables + op.mux.table_off * 4;
            const int nl = op.mux.nl, mnfix = op.mux.mnfix;
            const uint64_t mfixpos = op.mux.mfixpos;
            const uint32_t lppack = op.mux.lppack, ljpack = op.mux.ljpack;
            const uint32_t fbits = (uint32_t)(k - mnfix);
            // threads split into nt / T groups of T = min(nt, 2^fbits); a group walks the values of the inside index bits
            // (outer loop: matrix in registers), its threads the pairs of one value (inner loop)
            const uint32_t nt = blockDim.x;
            const uint32_t per = 1u << fbits, T = per < nt ? per : nt;
            const uint32_t t = lane_swaps(threadIdx.x & (T - 1), swaps), grp = threadIdx.x / T, ngrp = nt / T;
            const uint32_t lt = expand_local(t, mfixpos, mnfix);  // see tile_for_each
            const uint32_t himask = expand_local((per - 1u) & ~(T - 1u), mfixpos, mnfix);
            for (uint32_t il = grp; il < (1u << nl); il += ngrp) {
                uint32_t idx = idx_ext, lbits = 0;
                for (int i = 0; i < nl; ++i)
                    if ((il >> i) & 1u) {
                        idx |= 1u << ((ljpack >> (4 * i)) & 15u);
                        lbits |= 1u << ((lppack >> (4 * i)) & 15u);
                    }
                const double2* M = tab + (size_t)idx * 4;
                const C m00 = narrow(__ldg(M), (C*)nullptr), m01 = narrow(__ldg(M + 1), (C*)nullptr);
                const C m10 = narrow(__ldg(M + 2), (C*)nullptr), m11 = narrow(__ldg(M + 3), (C*)nullptr);
                const uint32_t lfix = lctrl | lbits;
                uint32_t p = t, hi = 0;
                for (; p + T < per; p += 2 * T) {  // two independent shared-memory round trips in flight
                    const uint32_t la = tile.swz(lt | hi | lfix);
                    hi = subset_next(hi, himask);
                    const uint32_t lb = tile.swz(lt | hi | lfix);
                    hi = subset_next(hi, himask);
                    const C a0 = tile.raw(la), a1 = tile.raw(la ^ bt);
                    const C b0 = tile.raw(lb), b1 = tile.raw(lb ^ bt);
                    tile.raw(la) = cfma(m01, a1, cmul(m00, a0));
                    tile.raw(la ^ bt) = cfma(m11, a1, cmul(m10, a0));
                    tile.raw(lb) = cfma(m01, b1, cmul(m00, b0));
                    tile.raw(lb ^ bt) = cfma(m11, b1, cmul(m10, b0));
                }
                for (; p < per; p += T) {
                    const uint32_t la = tile.swz(lt | hi | lfix);
                    hi = subset_next(hi, himask);
                    const C a0 = tile.raw(la), a1 = tile.raw(la ^ bt);
                    tile.raw(la) = cfma(m01, a1, cmul(m00, a0));
                    tile.raw(la ^ bt) = cfma(m11, a1, cmul(m10, a0));
                }
            }
        } break;
        case kOpBlock: if constexpr (FUSED) {
            // every group of 2^kb amplitudes (block bits free, predicate bits set) is multiplied by one dense matrix
            tile_apply_block<A>(tile, op, k, tables, bs);
        } break;
        default: {  // PHASE
            const C w = N::mk(mval<A>(op.m, 0), mval<A>(op.m, 1));
            tile_for_each<4>(tile, cnt, fixpos, nfix, lctrl, swaps, [&](const uint32_t* l, int m) {
                C v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u < m) v[u] = tile.raw(l[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (u < m) tile.raw(l[u]) = cmul(w, v[u]);
            });
        } break;
    }
}

// Register phase: consecutive plain gates whose targets fall on at most three tile qubits R are applied
// WITHOUT going back to shared memory in between: a thread loads the 8 amplitudes of one R-group,
// runs every gate of the phase on them in registers (targets are static register pairs, controls are
// predicates: outside the tile CTA-uniform, inside the tile but outside R per group, inside R per
// element) and stores them once.  One shared-memory round trip and one barrier per phase instead of
// one per gate; X and SWAP stay exact register moves.
#define B2SV_PAIRS(RP, F)                                   \
    switch (RP) {                                           \
        case 0: F(0, 1) F(2, 3) F(4, 5) F(6, 7) break;      \
        case 1: F(0, 2) F(1, 3) F(4, 6) F(5, 7) break;      \
        default: F(0, 4) F(1, 5) F(2, 6) F(3, 7) break;     \
    }

template <typename A, typename TV>
__device__ __forceinline__ void tile_apply_phase(TV tile, const TileOp& hdr, const TileOp* ops, int k, uint64_t base) {
    using C = typename Num<A>::C;
    using R = typename Num<A>::R;
    using N = Num<A>;
    const uint32_t r0 = hdr.ph.rpos[0], r1 = hdr.ph.rpos[1], r2 = hdr.ph.rpos[2];
    const uint32_t sb0 = tile.swz(1u << r0), sb1 = tile.swz(1u << r1), sb2 = tile.swz(1u << r2);
    const int count = hdr.ph.count;
    // groups are enumerated over the tile bits that are neither register bits nor controls shared by
    // every gate of the phase (hdr.lctrl): heavily controlled stretches only touch their own corner
    const uint32_t cnt = 1u << (k - hdr.nfix);
    if (threadIdx.x >= cnt) return;
    const uint32_t lt = expand_local(lane_swaps(threadIdx.x, (uint32_t)hdr.ext_tbit), hdr.fixpos, hdr.nfix) | hdr.lctrl;  // see tile_for_each
    const uint32_t himask = cnt > blockDim.x ? expand_local((cnt - 1u) & ~(blockDim.x - 1u), hdr.fixpos, hdr.nfix) : 0u;
    uint32_t hi = 0;
    for (uint32_t g = threadIdx.x; g < cnt; g += blockDim.x) {
        const uint32_t l = lt | hi;
        hi = subset_next(hi, himask);
        uint32_t off[8];
        C v[8];
        const uint32_t sl = tile.swz(l);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            off[j] = sl ^ ((j & 1) ? sb0 : 0u) ^ ((j & 2) ? sb1 : 0u) ^ ((j & 4) ? sb2 : 0u);
            v[j] = tile.raw(off[j]);
        }
        for (int mi = 0; mi < count; ++mi) {
            // the member's descriptor comes in as two 16-byte shared-memory loads issued back to back (field-by-field
            // reads were a chain of dependent loads and branches); the matrix follows once the member is known to act
            struct alignas(16) MemberWords {
                uint4 w0, w1;
            };
            const MemberWords op_w = *reinterpret_cast<const MemberWords*>(&ops[mi]);
            struct {
                uint8_t op, tpos;
                uint32_t nfix;
                uint64_t fixpos, ext_ctrl, ext_tbit;
                const double* m;
            } op;
            op.op = (uint8_t)(op_w.w0.x & 0xFFu);
            op.tpos = (uint8_t)((op_w.w0.x >> 8) & 0xFFu);
            op.nfix = op_w.w0.x >> 24;
            op.fixpos = (uint64_t)op_w.w0.z | ((uint64_t)op_w.w0.w << 32);
            op.ext_ctrl = (uint64_t)op_w.w1.x | ((uint64_t)op_w.w1.y << 32);
            op.ext_tbit = (uint64_t)op_w.w1.z | ((uint64_t)op_w.w1.w << 32);
            op.m = ops[mi].m;
            if ((base & op.ext_ctrl) != op.ext_ctrl) continue;
            // pre-decoded by the host (tileops.hpp): controls outside R, per-element mask of the controls inside R,
            // register position of the target(s), real-matrix flag
            const uint64_t fx = op.fixpos;
            const uint32_t lco = (uint32_t)fx;
            if ((l & lco) != lco) continue;
            const uint32_t ok = op.nfix;  // bit j: element j of the group satisfies the controls that lie in R
            const int rp = (int)((fx >> 32) & 3u);
            switch (op.op) {
                case B2SV_OP_X: {
#define B2SV_F(a, b)                       \
    if ((ok >> a) & 1u) {                  \
        const C t_ = v[a];           \
        v[a] = v[b];                       \
        v[b] = t_;                         \
    }
                    B2SV_PAIRS(rp, B2SV_F)
#undef B2SV_F
                } break;
                case B2SV_OP_MAT1: {
                    if ((fx >> 36) & 1u) {  // real matrix: two FP64 operations per component instead of four
                        const R r00 = mval<A>(op.m, 0), r01 = mval<A>(op.m, 2), r10 = mval<A>(op.m, 4), r11 = mval<A>(op.m, 6);
#define B2SV_F(a, b)                                                                  \
    if ((ok >> a) & 1u) {                                                             \
        const C x_ = v[a], y_ = v[b];                                                 \
        v[a] = N::mk(fma(r01, y_.x, r00 * x_.x), fma(r01, y_.y, r00 * x_.y));         \
        v[b] = N::mk(fma(r11, y_.x, r10 * x_.x), fma(r11, y_.y, r10 * x_.y));         \
    }
                        B2SV_PAIRS(rp, B2SV_F)
#undef B2SV_F
                        break;
                    }
                    const C m00 = N::mk(mval<A>(op.m, 0), mval<A>(op.m, 1)), m01 = N::mk(mval<A>(op.m, 2), mval<A>(op.m, 3));
                    const C m10 = N::mk(mval<A>(op.m, 4), mval<A>(op.m, 5)), m11 = N::mk(mval<A>(op.m, 6), mval<A>(op.m, 7));
#define B2SV_F(a, b)                                   \
    if ((ok >> a) & 1u) {                              \
        const C x_ = v[a], y_ = v[b];                  \
        v[a] = cfma(m01, y_, cmul(m00, x_));           \
        v[b] = cfma(m11, y_, cmul(m10, x_));           \
    }
                    B2SV_PAIRS(rp, B2SV_F)
#undef B2SV_F
                } break;
                case B2SV_OP_DIAG1: {
                    const C d0 = N::mk(mval<A>(op.m, 0), mval<A>(op.m, 1)), d1 = N::mk(mval<A>(op.m, 2), mval<A>(op.m, 3));
                    if (op.tpos == 0xFF) {
                        const C d = (base & op.ext_tbit) ? d1 : d0;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if ((ok >> j) & 1u) v[j] = cmul(d, v[j]);
                    } else {
#define B2SV_F(a, b)                   \
    if ((ok >> a) & 1u) {              \
        v[a] = cmul(d0, v[a]);         \
        v[b] = cmul(d1, v[b]);         \
    }
                        B2SV_PAIRS(rp, B2SV_F)
#undef B2SV_F
                    }
                } break;
                case B2SV_OP_SWAP: {
                    const int rq = (int)((fx >> 34) & 3u);
                    const int pair = rp + rq;  // {0,1} -> 1, {0,2} -> 2, {1,2} -> 3
#define B2SV_S(a, b)                       \
    if ((ok >> a) & 1u) {                  \
        const C t_ = v[a];           \
        v[a] = v[b];                       \
        v[b] = t_;                         \
    }
                    if (pair == 1) { B2SV_S(1, 2) B2SV_S(5, 6) }
                    else if (pair == 2) { B2SV_S(1, 4) B2SV_S(3, 6) }
                    else { B2SV_S(2, 4) B2SV_S(3, 5) }
#undef B2SV_S
                } break;
                default: {  // PHASE
                    const C w = N::mk(mval<A>(op.m, 0), mval<A>(op.m, 1));
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if ((ok >> j) & 1u) v[j] = cmul(w, v[j]);
                } break;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) tile.raw(off[j]) = v[j];
    }
}
#undef B2SV_PAIRS

// Monomial phase: a run of consecutive X / SWAP / PHASE / DIAG1 gates (any controls) maps every basis
// state of the tile to ONE basis state times a factor, so the whole run is "index arithmetic + one
// move" -- and the index arithmetic does not depend on the amplitudes, so the HOST does it once per plan
// (tileops.hpp: mono_build_table, bit-sliced) and the device only moves data.  Only the tile bits the run
// touches (targets and tile-local controls: the Q bits) decide where an amplitude goes, so the thread <->
// amplitude map puts the Q bits into the thread number first: a thread then owns 2^lb amplitudes that
// share at most V different Q values ("walks"), fetches the V results -- the XOR mask from source to
// destination index and, if the run has phase gates, the accumulated factor -- and moves its amplitudes
// in rounds of kMonoElems with one barrier between the reads and the writes of a round (a round holds
// every Q value of a set of non-Q values, so it is closed under the run).
// One shared-memory round trip for a run of ANY length: the ripple-carry arithmetic of
// circuits/arithmetic.py under LCU / BlockDiagonal selector controls is 50-150 such gates in a row, which
// used to be one shared-memory pass each.  X and SWAP stay exact data movement.
//
// scatter the low bits of x to the set positions of mask (ascending)
__host__ __device__ __forceinline__ uint32_t mono_deposit(uint32_t x, uint32_t mask) {
    uint32_t out = 0;
#pragma unroll 1
    while (mask) {
        const uint32_t low = mask & (0u - mask);
        if (x & 1u) out |= low;
        x >>= 1;
        mask ^= low;
    }
    return out;
}
// successor of subset `cur` of `mask` in counting order (wraps to 0 after the full mask)
__host__ __device__ __forceinline__ uint32_t mono_next_subset(uint32_t cur, uint32_t mask) { return ((cur | ~mask) + 1u) & mask; }

template <typename A, int V, bool FAC, typename TV>
__device__ __forceinline__ void tile_apply_mono(TV tile, const TileOp& hdr, uint64_t base, uint32_t slice,
                                                const double2* __restrict__ tables) {
    using C = typename Num<A>::C;
    constexpr int E = kMonoElems;
    static_assert(V == 8 || V == 4, "walks per thread");
    const uint32_t tmask = hdr.mono.tmask, qlmask = hdr.mono.qlmask, nmask = hdr.mono.nmask;
    const uint32_t per_thread = 1u << hdr.mono.lb;
    const int tbits = __popc(tmask);
    // (a CTA with more than 2^tbits threads -- the pipelined kernel -- has the upper half mirror the lower half:
    // same reads, same writes of the same values)
    const uint32_t t = threadIdx.x & ((1u << tbits) - 1u);
    const uint32_t tpart = mono_deposit(t, tmask);
    // this thread's walk results: entry (slice, thread, walk)
    const size_t first = (((size_t)slice << tbits) + t) * V;
    uint32_t xw[V / 2];  // V uint16 XOR masks, two per word
    if (V == 8) {
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(tables + hdr.mono.mask_off) + first));
        xw[0] = w.x;
        xw[1] = w.y;
        xw[V / 2 - 2] = w.z;
        xw[V / 2 - 1] = w.w;
    } else {
        const uint2 w = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(tables + hdr.mono.mask_off) + first));
        xw[0] = w.x;
        xw[1] = w.y;
    }
    C fac[FAC ? V : 1];
    if (FAC) {
        const double2* __restrict__ fm = tables + hdr.mono.fac_off + first;
#pragma unroll
        for (int v = 0; v < V; ++v) fac[v] = narrow(__ldg(fm + v), (C*)nullptr);
    }
    // amplitude e of a round belongs to walk e mod V (E is a multiple of V); `q` runs through the walk bits,
    // `ncur` through the untouched loop bits.  The destinations are recomputed after the barrier rather
    // than kept in registers.
    uint32_t ncur = 0;
#pragma unroll 1
    for (uint32_t j0 = 0; j0 < per_thread; j0 += E) {
        C val[E];
        {
            uint32_t q = 0, nn = ncur;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                constexpr int VM = V - 1;
                const int v = e & VM;
                val[e] = tile[tpart | q | nn];
                if (FAC) val[e] = cmul(fac[v], val[e]);
                q = mono_next_subset(q, qlmask);
                if (v == VM) nn = mono_next_subset(nn, nmask);
            }
        }
        __syncthreads();  // every amplitude of the round is in registers before the first one is overwritten
        {
            uint32_t q = 0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                constexpr int VM = V - 1;
                const int v = e & VM;
                const uint32_t xm = (v & 1) ? (xw[v >> 1] >> 16) : (xw[v >> 1] & 0xFFFFu);
                tile[(tpart | q | ncur) ^ xm] = val[e];
                q = mono_next_subset(q, qlmask);
                if (v == VM) ncur = mono_next_subset(ncur, nmask);
            }
        }
        // the next round works on a disjoint set of amplitudes: no barrier needed in between
    }
}

// Run the staged ops [0, n) of a sweep on the tile in shared memory.
template <typename A, bool FUSED, typename TV>
__device__ __forceinline__ void tile_run_ops(TV tile, const TileOp* s_ops, int n, int k, uint64_t base,
                                             const double2* __restrict__ tables, const BlkSlots* bs) {
    for (int o = 0; o < n; ++o) {
        const TileOp& op = s_ops[o];
        if (op.op == kOpMono) {
            uint32_t slice = 0;  // CTA-uniform: which of the outside bits the run tests are set for this tile
            for (int i = 0; i < op.mono.ebits; ++i) slice |= (uint32_t)((base >> op.mono.epos[i]) & 1ull) << i;
            if (!((op.mono.identity >> slice) & 1u)) {
                if (op.mono.fac) tile_apply_mono<A, kMonoVFac, true>(tile, op, base, slice, tables);
                else tile_apply_mono<A, kMonoV, false>(tile, op, base, slice, tables);
                __syncthreads();
            }
            continue;
        }
        if (op.op == kOpPhase) {
            bool any = false;  // CTA-uniform: is any gate of the phase active on this tile?
            for (int mi = 1; mi <= op.ph.count; ++mi) any |= (base & s_ops[o + mi].ext_ctrl) == s_ops[o + mi].ext_ctrl;
            if (any) {
                tile_apply_phase<A>(tile, op, s_ops + o + 1, k, base);
                __syncthreads();
            }
            o += op.ph.count;
            continue;
        }
        if (op.op == kOpNop) continue;
        if ((base & op.ext_ctrl) != op.ext_ctrl) continue;  // CTA-uniform: gate not active on this tile
        tile_apply_op<A, FUSED>(tile, op, k, base, tables, bs);
        __syncthreads();
    }
}

// Long op lists come in batches of kTileMaxOps slots (the host never lets a phase straddle a batch
// boundary): the first batch is staged while the tile is in flight, later ones between batches.
__device__ __forceinline__ void tile_stage_ops(TileOp* s_ops, const TileOp* g_ops, int n, int tid, int nt) {
    const uint4* src = reinterpret_cast<const uint4*>(g_ops);
    uint4* dst = reinterpret_cast<uint4*>(s_ops);
    const int n16 = n * (int)(sizeof(TileOp) / 16);
    for (int i = tid; i < n16; i += nt) dst[i] = __ldg(src + i);
}

// complex64 sweeps: narrow the matrix entries of the staged plain ops (stand-alone ones and register-phase members) to
// floats in place, once per staged batch (see mval)
template <typename A>
__device__ __forceinline__ void tile_narrow_ops(TileOp* s_ops, int n, int tid, int nt) {
    if constexpr (sizeof(A) == 8) {
        for (int o = tid; o < n; o += nt) {
            TileOp& op = s_ops[o];
            if (op.op > B2SV_OP_PHASE) continue;  // table-driven ops and phase headers carry no matrix here
            double t[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) t[i] = op.m[i];
            float* f = reinterpret_cast<float*>(op.m);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = (float)t[i];
        }
    }
}

template <typename A, bool FUSED, typename TV>
__device__ __forceinline__ void tile_run_batches(TV tile, TileOp* s_ops, const TileOp* g_ops, int n_ops, int k, uint64_t base,
                                                 const double2* __restrict__ tables, const BlkSlots* bs) {
    for (int b0 = 0; b0 < n_ops; b0 += kTileMaxOps) {
        const int nb = n_ops - b0 < kTileMaxOps ? n_ops - b0 : kTileMaxOps;
        if (b0 > 0) {
            __syncthreads();
            tile_stage_ops(s_ops, g_ops + b0, nb, (int)threadIdx.x, (int)blockDim.x);
            __syncthreads();
            tile_narrow_ops<A>(s_ops, nb, (int)threadIdx.x, (int)blockDim.x);
            __syncthreads();
        }
        tile_run_ops<A, FUSED>(tile, s_ops, nb, k, base, tables, bs);
    }
}

// global offset (in amplitudes, relative to the tile base) of chunk j of the tile
__device__ __forceinline__ uint64_t chunk_offset(uint32_t j, const ExpandArgs& ex, int k, int lc) {
    uint64_t off = 0;
    for (int i = lc; i < k; ++i)
        if ((j >> (i - lc)) & 1) off |= 1ull << ex.pos[i];
    return off;
}

template <typename A, bool BULK, bool FUSED>
__global__ void __launch_bounds__(kThreads, 3) k_tile(A* __restrict__ psi, const TileArgs a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    A* tile = reinterpret_cast<A*>(smem_raw);
    const int k = a.k, lc = a.lc;
    // the sweep's ops live behind the tile: fetched once, coalesced, while the tile is in flight
    TileOp* s_ops = reinterpret_cast<TileOp*>(smem_raw + ((size_t)sizeof(A) << k));
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x;
    const uint32_t n_chunks = 1u << (k - lc), chunk_elems = 1u << lc;
    const uint32_t chunk_bytes = chunk_elems * (uint32_t)sizeof(A);
    const uint64_t base = expand_bits((uint64_t)blockIdx.x, a.cta_ex) | a.cta_or;

    const bool load = a.init_mode == 0;
    if (BULK && load) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) mbar_expect_tx(&bar, chunk_bytes * n_chunks);
        __syncthreads();
        for (uint32_t j = tid; j < n_chunks; j += kThreads)
            bulk_g2s(tile + (size_t)j * chunk_elems, psi + (base | chunk_offset(j, a.tile_ex, k, lc)), chunk_bytes, &bar);
    }
    tile_stage_ops(s_ops, a.ops, a.n_ops < kTileMaxOps ? a.n_ops : kTileMaxOps, tid, kThreads);
    if constexpr (sizeof(A) == 8) {
        __syncthreads();
        tile_narrow_ops<A>(s_ops, a.n_ops < kTileMaxOps ? a.n_ops : kTileMaxOps, tid, kThreads);
    }
    if (!load) {
        const uint32_t elems = 1u << k;
        const A zero = from_c128<A>(make_double2(0.0, 0.0));
        for (uint32_t l = tid; l < elems; l += kThreads) tile[l] = zero;
        __syncthreads();
        if (tid == 0 && (a.init_index & ~a.tile_mask) == base) {
            uint32_t l = 0;
            for (int i = 0; i < k; ++i) l |= (uint32_t)((a.init_index >> a.tile_ex.pos[i]) & 1ull) << i;
            tile[l] = from_c128<A>(make_double2(1.0, 0.0));
        }
    } else if (BULK) {
        mbar_wait(&bar, 0);
    } else {
        const uint32_t elems = 1u << k;
        for (uint32_t l = tid; l < elems; l += kThreads) {
            const uint64_t g = base | chunk_offset(l >> lc, a.tile_ex, k, lc) | (l & (chunk_elems - 1));
            tile[l] = psi[g];
        }
    }
    __syncthreads();

    // a generated (lazy-init) tile that does not hold the basis index is all zeros: every gate maps it to
    // itself, so only the one tile that owns the index runs the gates -- the others just write zeros
    const int n_ops_here = (load || (a.init_index & ~a.tile_mask) == base) ? a.n_ops : 0;
    tile_run_batches<A, FUSED>(TileView<A, false>{tile, 1u << k}, s_ops, a.ops, n_ops_here, k, base, a.tables, nullptr);

    if (BULK) {
        // generic-proxy writes to shared memory must be visible to the async proxy before the copy out
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        for (uint32_t j = tid; j < n_chunks; j += kThreads)
            bulk_s2g(psi + (base | chunk_offset(j, a.tile_ex, k, lc)), tile + (size_t)j * chunk_elems, chunk_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    } else {
        const uint32_t elems = 1u << k;
        for (uint32_t l = tid; l < elems; l += kThreads) {
            const uint64_t g = base | chunk_offset(l >> lc, a.tile_ex, k, lc) | (l & (chunk_elems - 1));
            psi[g] = tile[l];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Swizzled form of the tile sweep (the default whenever a tile chunk is at least one 128-byte row): the
// chunks are moved by TMA *tensor* copies (cp.async.bulk.tensor.2d, SASS UTMALDG / UTMASTG) through a
// tensor map that views the state as rows of 128 bytes with the hardware 128-byte swizzle, so the tile
// lands in shared memory already XOR-swizzled and goes back un-swizzled, with no extra pass.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_g2s_2d(void* dst_smem, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_s2g_2d(const CUtensorMap* map, uint32_t c0, uint32_t c1, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(smem_u32(src_smem))
                 : "memory");
}

struct alignas(64) TileArgsSwz {
    CUtensorMap map;  // {16 doubles (128 B), rows}, box {16, rows per chunk}, SWIZZLE_128B
    TileArgs t;
    BlkSlots blk;     // the sweep's dense-block matrices (immediate constant-bank operands, see tile_apply_block)
};

template <typename A, bool FUSED>
__global__ void __launch_bounds__(kThreads, 3) k_tile_swz(const __grid_constant__ TileArgsSwz args) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const TileArgs& a = args.t;
    A* tile_raw = reinterpret_cast<A*>(smem_raw);
    const TileView<A, true> tile{tile_raw, 1u << args.t.k};
    const int k = a.k, lc = a.lc;
    TileOp* s_ops = reinterpret_cast<TileOp*>(smem_raw + ((size_t)sizeof(A) << k));
    // the mbarrier lives in dynamic shared memory too: static shared memory would push the 1 KiB-aligned
    // tile down by a whole kilobyte and cost the third resident CTA
    uint64_t& bar = *reinterpret_cast<uint64_t*>(smem_raw + ((size_t)sizeof(A) << k) + (size_t)kTileMaxOps * sizeof(TileOp));
    const int tid = threadIdx.x;
    const uint32_t n_chunks = 1u << (k - lc), chunk_elems = 1u << lc;
    const uint32_t chunk_bytes = chunk_elems * (uint32_t)sizeof(A);
    constexpr int kRowShift = sizeof(A) == 16 ? 3 : 4;  // amplitudes per 128-byte row
    const uint64_t base = expand_bits((uint64_t)blockIdx.x, a.cta_ex) | a.cta_or;
    const bool load = a.init_mode == 0;

    auto stamp = [&](int slot) {
        if (a.trace != nullptr && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            a.trace[(size_t)blockIdx.x * 5 + slot] = t;
            if (slot == 1) {
                unsigned int smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                a.trace[(size_t)blockIdx.x * 5] = smid;
            }
        }
    };
    stamp(1);
    if (load) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) mbar_expect_tx(&bar, chunk_bytes * n_chunks);
        __syncthreads();
        for (uint32_t j = tid; j < n_chunks; j += kThreads) {
            const uint64_t g = base | chunk_offset(j, a.tile_ex, k, lc);
            tma_g2s_2d(tile_raw + (size_t)j * chunk_elems, &args.map, 0u, (uint32_t)(g >> kRowShift), &bar);
        }
    }
    tile_stage_ops(s_ops, a.ops, a.n_ops < kTileMaxOps ? a.n_ops : kTileMaxOps, tid, kThreads);
    if constexpr (sizeof(A) == 8) {
        __syncthreads();
        tile_narrow_ops<A>(s_ops, a.n_ops < kTileMaxOps ? a.n_ops : kTileMaxOps, tid, kThreads);
    }
    if (!load) {
        const uint32_t elems = 1u << k;
        const A zero = from_c128<A>(make_double2(0.0, 0.0));
        for (uint32_t l = tid; l < elems; l += kThreads) tile_raw[l] = zero;
        __syncthreads();
        if (tid == 0 && (a.init_index & ~a.tile_mask) == base) {
            uint32_t l = 0;
            for (int i = 0; i < k; ++i) l |= (uint32_t)((a.init_index >> a.tile_ex.pos[i]) & 1ull) << i;
            tile[l] = from_c128<A>(make_double2(1.0, 0.0));
        }
    } else {
        mbar_wait(&bar, 0);
    }
    __syncthreads();
    stamp(2);

    // a generated (lazy-init) tile that does not hold the basis index is all zeros: every gate maps it to
    // itself, so only the one tile that owns the index runs the gates -- the others just write zeros
    const int n_ops_here = (load || (a.init_index & ~a.tile_mask) == base) ? a.n_ops : 0;
    tile_run_batches<A, FUSED>(tile, s_ops, a.ops, n_ops_here, k, base, a.tables, FUSED ? &args.blk : nullptr);

    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    stamp(3);
    for (uint32_t j = tid; j < n_chunks; j += kThreads) {
        const uint64_t g = base | chunk_offset(j, a.tile_ex, k, lc);
        tma_s2g_2d(&args.map, 0u, (uint32_t)(g >> kRowShift), tile_raw + (size_t)j * chunk_elems);
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    stamp(4);
}

// ---------------------------------------------------------------------------------------------------
// Low-qubit warp-shuffle sweep.  When every target of a sweep lies in the low six qubits (state
// preparation on the low register: circuits/state_prep.py emits 2*2^k gates per target there), the
// 64 amplitudes that interact are one contiguous, aligned 1 KiB block.  A warp holds such a block in
// registers -- lane i owns amplitudes i and i+32 -- so qubit 5 is an in-thread pair and qubits 0..4
// are butterfly exchanges with lane ^ (1 << q) by warp shuffle.  All gates of the sweep run back to
// back on the registers; global memory is touched once in, once out, fully coalesced, and no shared
// memory is involved.  Controls above the block are warp-uniform predicates.
// ---------------------------------------------------------------------------------------------------
struct Low6Args {
    ExpandArgs blk_ex;   // low 6 bits + the sweep's common outside controls: block number -> base index
    uint64_t blk_or;
    uint64_t n_blocks;
    int n_ops;
    const TileOp* ops;   // built against the pseudo tile {0..5}: tile-local position == qubit
    const double2* tables;
    int init_mode;       // 1: generate the pending basis state |init_index> instead of loading
    uint64_t init_index;
};

__device__ __forceinline__ double2 shfl_xor_c128(double2 v, int lane_mask) {
    return make_double2(__shfl_xor_sync(0xffffffffu, v.x, lane_mask), __shfl_xor_sync(0xffffffffu, v.y, lane_mask));
}

// apply the 2x2 matrix (m00..m11) on `target` to the two amplitudes a lane owns (e0 = lane, e1 = lane|32);
// ok0/ok1: the element satisfies the gate's in-block controls.  All 32 lanes must call this together.
__device__ __forceinline__ void low6_apply_2x2(double2& v0, double2& v1, int lane, int target, bool ok0, bool ok1,
                                               double2 m00, double2 m01, double2 m10, double2 m11) {
    if (target == 5) {
        const double2 n0 = cfma(m01, v1, cmul(m00, v0)), n1 = cfma(m11, v1, cmul(m10, v0));
        if (ok0) {
            v0 = n0;
            v1 = n1;
        }
    } else {
        const double2 p0 = shfl_xor_c128(v0, 1 << target), p1 = shfl_xor_c128(v1, 1 << target);
        const bool hi = (lane >> target) & 1;
        const double2 a = hi ? m11 : m00, b = hi ? m10 : m01;  // new = a * mine + b * partner
        if (ok0) v0 = cfma(b, p0, cmul(a, v0));
        if (ok1) v1 = cfma(b, p1, cmul(a, v1));
    }
}

constexpr int kLow6Blocks = 4;  // 64-amplitude blocks a warp holds at once (amortises the table look-ups)

// coefficients of "new = ca * mine + cb * partner" of a multiplexor for the lane's two elements
__device__ __forceinline__ void low6_mux_coeffs(const TileOp& op, const double2* __restrict__ tables, uint64_t base, uint32_t e0,
                                                uint32_t e1, int tgt, bool hi, double2& ca0, double2& cb0, double2& ca1,
                                                double2& cb1) {
    uint32_t i0 = 0, i1 = 0;
    for (int j = 0; j < op.mux.nmux; ++j) {
        if (op.mux.mpos[j] == 0xFF) {
            const uint32_t bit = (uint32_t)((base >> op.mux.mq[j]) & 1ull) << j;
            i0 |= bit;
            i1 |= bit;
        } else {
            i0 |= ((e0 >> op.mux.mpos[j]) & 1u) << j;
            i1 |= ((e1 >> op.mux.mpos[j]) & 1u) << j;
        }
    }
    const double2* T0 = tables + (op.mux.table_off + i0) * 4;
    const double2* T1 = tables + (op.mux.table_off + i1) * 4;
    if (tgt == 5) {  // in-thread pair: element 0 is the |0> row, element 1 the |1> row
        ca0 = __ldg(T0);
        cb0 = __ldg(T0 + 1);
        ca1 = __ldg(T1 + 3);
        cb1 = __ldg(T1 + 2);
    } else {
        ca0 = __ldg(T0 + (hi ? 3 : 0));
        cb0 = __ldg(T0 + (hi ? 2 : 1));
        ca1 = __ldg(T1 + (hi ? 3 : 0));
        cb1 = __ldg(T1 + (hi ? 2 : 1));
    }
}

template <typename A>
__global__ void __launch_bounds__(kThreads) k_low6(A* __restrict__ psi, const Low6Args a) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    TileOp* s_ops = reinterpret_cast<TileOp*>(smem_raw);
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.ops);
        uint4* dst = reinterpret_cast<uint4*>(s_ops);
        const int n16 = a.n_ops * (int)(sizeof(TileOp) / 16);
        for (int i = threadIdx.x; i < n16; i += kThreads) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    constexpr int NB = kLow6Blocks;
    const int lane = threadIdx.x & 31;
    const uint32_t e0 = lane, e1 = lane | 32;
    const uint64_t warps = (uint64_t)gridDim.x * (kThreads / 32);
    const uint64_t n_groups = (a.n_blocks + NB - 1) / NB;
    for (uint64_t grp = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); grp < n_groups; grp += warps) {
        uint64_t base[NB];
        bool live[NB];
        double2 v0[NB], v1[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const uint64_t b = grp * NB + u;
            live[u] = b < a.n_blocks;
            base[u] = expand_bits(live[u] ? b : 0, a.blk_ex) | a.blk_or;
            if (a.init_mode || !live[u]) {
                v0[u] = make_double2(live[u] && (base[u] | e0) == a.init_index ? 1.0 : 0.0, 0.0);
                v1[u] = make_double2(live[u] && (base[u] | e1) == a.init_index ? 1.0 : 0.0, 0.0);
            } else {
                v0[u] = to_c128(psi[base[u] + e0]);
                v1[u] = to_c128(psi[base[u] + e1]);
            }
        }
        // a generated (lazy-init) group that does not hold the basis index is all zeros, which every gate
        // maps to itself: only the one warp that owns the index runs the gates (warp-uniform decision)
        bool run_ops = true;
        if (a.init_mode) {
            run_ops = false;
#pragma unroll
            for (int u = 0; u < NB; ++u) run_ops |= live[u] && base[u] == (a.init_index & ~63ull);
        }
        const int n_ops_here = run_ops ? a.n_ops : 0;
        for (int o = 0; o < n_ops_here; ++o) {
            // everything about the gate is warp-uniform; the per-block / per-element conditions are plain
            // predicates, so the shuffles below are never inside divergent code
            const TileOp& op = s_ops[o];
            const uint32_t lc = op.lctrl;
            const uint64_t ext = op.ext_ctrl;
            const int kind = op.op;
            const int tgt = op.tpos;
            const bool in_thread = tgt == 5;
            const int xm = (tgt < 5) ? (1 << tgt) : 0;
            const bool hi = tgt < 5 ? ((lane >> tgt) & 1) : false;
            const bool ok0 = (e0 & lc) == lc, ok1 = (e1 & lc) == lc;
            bool a0[NB], a1[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                const bool act = (base[u] & ext) == ext;
                a0[u] = act && ok0;
                a1[u] = act && ok1;
            }
            if (kind == B2SV_OP_X) {  // exact data movement: take the partner's amplitude, no arithmetic
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    const double2 p0 = in_thread ? v1[u] : shfl_xor_c128(v0[u], xm);
                    const double2 p1 = in_thread ? v0[u] : shfl_xor_c128(v1[u], xm);
                    if (a0[u]) v0[u] = p0;
                    if (a1[u]) v1[u] = p1;
                }
            } else if (kind == B2SV_OP_DIAG1) {
                const double2 d0 = make_double2(op.m[0], op.m[1]), d1 = make_double2(op.m[2], op.m[3]);
                const bool ext_t = tgt == 0xFF;
                const uint64_t tb = op.ext_tbit;
                const double2 f0 = ext_t ? d0 : (((e0 >> (tgt & 7)) & 1u) ? d1 : d0);
                const double2 f1 = ext_t ? d0 : (((e1 >> (tgt & 7)) & 1u) ? d1 : d0);
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    const bool up = ext_t && (base[u] & tb);
                    if (a0[u]) v0[u] = cmul(up ? d1 : f0, v0[u]);
                    if (a1[u]) v1[u] = cmul(up ? d1 : f1, v1[u]);
                }
            } else if (kind == B2SV_OP_PHASE) {
                const double2 w = make_double2(op.m[0], op.m[1]);
#pragma unroll
                for (int u = 0; u < NB; ++u) {
                    if (a0[u]) v0[u] = cmul(w, v0[u]);
                    if (a1[u]) v1[u] = cmul(w, v1[u]);
                }
            } else {  // MAT1 or multiplexor: new = ca * mine + cb * partner
                double2 ca0, cb0, ca1, cb1;
                bool per_block = false;
                if (kind == kOpMux) {
                    for (int j = 0; j < op.mux.nmux; ++j) per_block |= op.mux.mpos[j] == 0xFF;
                    if (!per_block) low6_mux_coeffs(op, a.tables, 0ull, e0, e1, tgt, hi, ca0, cb0, ca1, cb1);
                } else if (in_thread) {
                    ca0 = make_double2(op.m[0], op.m[1]);
                    cb0 = make_double2(op.m[2], op.m[3]);
                    ca1 = make_double2(op.m[6], op.m[7]);
                    cb1 = make_double2(op.m[4], op.m[5]);
                } else {
                    ca0 = ca1 = hi ? make_double2(op.m[6], op.m[7]) : make_double2(op.m[0], op.m[1]);
                    cb0 = cb1 = hi ? make_double2(op.m[4], op.m[5]) : make_double2(op.m[2], op.m[3]);
                }
                if (lc == 0 && ext == 0 && !per_block) {
                    // the common state-preparation case: no predicate at all (a multiplexor carries its
                    // controls in the table), so the update is unconditional straight-line FP64 code
                    if (in_thread) {
#pragma unroll
                        for (int u = 0; u < NB; ++u) {
                            const double2 x = v0[u], y = v1[u];
                            v0[u] = cfma(cb0, y, cmul(ca0, x));
                            v1[u] = cfma(cb1, x, cmul(ca1, y));
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < NB; ++u) {
                            const double2 p0 = shfl_xor_c128(v0[u], xm), p1 = shfl_xor_c128(v1[u], xm);
                            v0[u] = cfma(cb0, p0, cmul(ca0, v0[u]));
                            v1[u] = cfma(cb1, p1, cmul(ca1, v1[u]));
                        }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < NB; ++u) {
                        if (per_block) low6_mux_coeffs(op, a.tables, base[u], e0, e1, tgt, hi, ca0, cb0, ca1, cb1);
                        const double2 p0 = in_thread ? v1[u] : shfl_xor_c128(v0[u], xm);
                        const double2 p1 = in_thread ? v0[u] : shfl_xor_c128(v1[u], xm);
                        const double2 n0 = cfma(cb0, p0, cmul(ca0, v0[u])), n1 = cfma(cb1, p1, cmul(ca1, v1[u]));
                        if (a0[u]) v0[u] = n0;
                        if (a1[u]) v1[u] = n1;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < NB; ++u)
            if (live[u]) {
                psi[base[u] + e0] = from_c128<A>(v0[u]);
                psi[base[u] + e1] = from_c128<A>(v1[u]);
            }
    }
}

// Zero-fill of a state buffer: the write-only half of a window-mode permutation sweep (zero-fill + forward scatter) and
// of a lazily requested zero / basis state.  256-bit stores (STG.E.256), four per thread, ONE 32 KiB block per CTA and no
// grid-stride loop: measured on the 8 GiB scratch buffer of cfg 2 (gpurun_out/r2ai_*, r2aj_*) -- cudaMemsetAsync 1.224 ms,
// a persistent grid (148 x 16 CTAs) of 128-bit streaming stores 1.32-1.33 ms, of plain 128-bit stores 1.36 ms, of 256-bit
// stores 1.29 ms, TMA bulk stores from a zeroed shared-memory buffer 1.40 ms, this form 1.193 ms (7.2 TB/s): consecutive
// CTAs write consecutive 32 KiB blocks, so the block scheduler keeps the DRAM pages in order.
constexpr uint32_t kFillBlockBytes = 32768;
__global__ void __launch_bounds__(256) k_fill_zero(unsigned char* __restrict__ p, uint64_t n32) {
    const uint64_t i = (uint64_t)blockIdx.x * 1024u + threadIdx.x;
#pragma unroll
    for (int u = 0; u < 4; ++u)
        if (i + (uint64_t)u * 256u < n32) {
            const unsigned long long z = 0ull;
            asm volatile("st.global.v4.b64 [%0], {%1, %1, %1, %1};" ::"l"(p + (i + (uint64_t)u * 256u) * 32), "l"(z) : "memory");
        }
}

// ===================================================================================================
// 3. Permutation run: a run of X / multi-controlled X / (controlled) SWAP gates is a bijection f on
//    amplitude indices.  One out-of-place sweep computes out[j] = in[f^-1(j)] (writes coalesced).
//    f^-1 is evaluated bit-sliced: a thread owns 32 output indices; word q holds bit q of all 32, so
//    an MCX is "W[t] ^= W[c0] & W[c1] & ..." -- one pass of word ops per gate for 32 amplitudes.
//    Words live in shared memory (a register file cannot be indexed by a run-time qubit number); the
//    gate program is read through the constant cache so it does not compete with shared memory.
// ===================================================================================================

constexpr int kPermMaxOps = 3840;          // 60 KiB of constant memory, 16 bytes per op
constexpr int kPermThreads = 256;
constexpr int kPermRowBytes = kPermThreads * 4;
constexpr int kPermTemps = 2;              // scratch rows for control lists longer than three

// One word op: W[t] ^= W[c0] & W[c1] & W[c2], every field a row BYTE offset (row * kPermRowBytes).
// Row n is all ones (unused control slots), rows n+1.. are zero-initialised temporaries.
//   X(t; c...)        -> one op (<= 3 controls); longer lists AND into a temporary first, use it as a
//                        control and un-compute it afterwards
//   SWAP(a,b; c...)   -> three ops: b ^= a&c, a ^= b&c, b ^= a&c
__constant__ uint4 c_perm_ops[kPermMaxOps];

// 32x32 bit-matrix transpose, LSB convention: out[b] bit q == in[q] bit b.
__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
    uint32_t m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j != 0; j >>= 1) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if ((k & j) == 0) {
                const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
                a[k] ^= t << j;
                a[k + j] ^= t;
            }
        }
        m ^= m << (j >> 1);
    }
}

template <typename A>
__global__ void __launch_bounds__(kPermThreads) k_perm(const A* __restrict__ in, A* __restrict__ out, int n, int n_ops,
                                                       uint64_t win_mask, uint64_t win_value) {
    extern __shared__ __align__(16) uint32_t W[];  // (n + 1 + kPermTemps) rows of kPermThreads words
    const int tid = threadIdx.x, lane = tid & 31;
    const uint64_t warp_global = (uint64_t)blockIdx.x * (kPermThreads / 32) + (tid >> 5);
    const uint64_t warp_base = warp_global << 10;  // 1024 outputs per warp: j = warp_base + b*32 + lane
    // initial words: bit b of word q = bit q of (warp_base + b*32 + lane)
    for (int q = 0; q < n; ++q) {
        uint32_t w;
        if (q < 5) w = ((lane >> q) & 1) ? 0xFFFFFFFFu : 0u;
        else if (q == 5) w = 0xAAAAAAAAu;
        else if (q == 6) w = 0xCCCCCCCCu;
        else if (q == 7) w = 0xF0F0F0F0u;
        else if (q == 8) w = 0xFF00FF00u;
        else if (q == 9) w = 0xFFFF0000u;
        else w = ((warp_global >> (q - 10)) & 1) ? 0xFFFFFFFFu : 0u;
        W[q * kPermThreads + tid] = w;
    }
    W[n * kPermThreads + tid] = 0xFFFFFFFFu;
    for (int r = 1; r <= kPermTemps; ++r) W[(n + r) * kPermThreads + tid] = 0u;
    unsigned char* col = reinterpret_cast<unsigned char*>(W + tid);  // this thread's column
#pragma unroll 4
    for (int o = 0; o < n_ops; ++o) {
        const uint4 op = c_perm_ops[o];
        const uint32_t m = *reinterpret_cast<uint32_t*>(col + op.y) & *reinterpret_cast<uint32_t*>(col + op.z) &
                           *reinterpret_cast<uint32_t*>(col + op.w);
        *reinterpret_cast<uint32_t*>(col + op.x) ^= m;
    }
    // un-slice with a register bit-matrix transpose: src[b] = source index of output b
    uint32_t lo[32];
#pragma unroll
    for (int q = 0; q < 32; ++q) lo[q] = q < n ? W[q * kPermThreads + tid] : 0u;
    transpose32(lo);
#pragma unroll
    for (int b0 = 0; b0 < 32; b0 += 8) {
        uint64_t src[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) src[u] = lo[b0 + u];
        for (int q = 32; q < n; ++q) {  // registers wider than 32 qubits (sharded / 33-34 qubit runs)
            const uint32_t w = W[q * kPermThreads + tid] >> b0;
#pragma unroll
            for (int u = 0; u < 8; ++u) src[u] |= (uint64_t)((w >> u) & 1u) << q;
        }
        // live window (win_mask != 0): the input is known to be zero outside {i : (i & win_mask) == win_value}
        // (a basis state that has only been through one tile so far) and is not even materialised there
        A v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            B2SV_CHECK(src[u] < (1ull << n));
            v[u] = ((src[u] & win_mask) == win_value) ? in[src[u]] : from_c128<A>(make_double2(0.0, 0.0));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) out[warp_base + (uint64_t)(b0 + u) * 32 + lane] = v[u];
    }
}

// ---------------------------------------------------------------------------------------------------
// Distributed permutation sweep (sharded register, one process per GPU): out[j] = s_r * in_r[f^-1(rank:j)],
// where the upper bits of the source index pick the peer shard `r` to read -- peers' shards are mapped into
// this process with CUDA IPC, so the loads below go over NVLink.  The NVRTC-specialised form of this kernel
// is generated by perm_jit.hpp (perm_dist_source); the interpreter below is what runs for shards too small
// for the bit-sliced layout (< 2^13 amplitudes) or when NVRTC is not available.
// ---------------------------------------------------------------------------------------------------
constexpr int kDistMaxRanks = 16;

struct DistPermArgs {
    const void* src[kDistMaxRanks];                        // current state buffer of every rank (own pointer for itself)
    unsigned long long wmask[kDistMaxRanks], wval[kDistMaxRanks];  // live window of every source shard (0,0: all of it)
    double sre[kDistMaxRanks], sim[kDistMaxRanks];         // pending lazy scalar of every source shard
    void* out;
    unsigned rank_logical;  // logical values of the global index bits of this rank's outputs
    unsigned flip_in;       // physical source rank = logical source bits ^ flip_in
    int has_scale;
    unsigned block_rot;     // this rank starts its sweep block_rot blocks in: in a k-qubit swap the ranks then read from
                            // DIFFERENT peers at any moment (unrotated, all of them pull from the same peer's HBM at once:
                            // 172 GB/s per direction at N = 8)
};

struct DistGate {
    uint64_t ctrl;
    uint8_t op, t0, t1, pad[5];
};

template <typename A>
__global__ void __launch_bounds__(kThreads) k_dist_perm_generic(const __grid_constant__ DistPermArgs a, const DistGate* __restrict__ gates,
                                                                int n_gates, int nl, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t lmask = (1ull << nl) - 1ull;
    A* out = reinterpret_cast<A*>(a.out);
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count; j += stride) {
        uint64_t idx = j | ((uint64_t)a.rank_logical << nl);
        for (int gi = n_gates - 1; gi >= 0; --gi) {  // gather: the self-inverse gates in reverse order
            const DistGate g = gates[gi];
            if ((idx & g.ctrl) != g.ctrl) continue;
            if (g.op == B2SV_OP_X) {
                idx ^= 1ull << g.t0;
            } else if (((idx >> g.t0) ^ (idx >> g.t1)) & 1ull) {
                idx ^= (1ull << g.t0) | (1ull << g.t1);
            }
        }
        const unsigned r = (unsigned)(idx >> nl) ^ a.flip_in;
        B2SV_CHECK(r < (unsigned)kDistMaxRanks && a.src[r] != nullptr);
        const uint64_t s = idx & lmask;
        A v = from_c128<A>(make_double2(0.0, 0.0));
        if ((s & a.wmask[r]) == a.wval[r]) v = reinterpret_cast<const A*>(a.src[r])[s];
        if (a.has_scale) v = from_c128<A>(cmul(make_double2(a.sre[r], a.sim[r]), to_c128(v)));
        out[j] = v;
    }
}

// Seeded dense test vector, generated on the device (benchmarks and full-size property tests: no 2^n host array, no
// PCIe upload): psi_i = scale * (u(2i), u(2i+1)) for global index index_base + i < support, 0 above, with
// u(k) = splitmix64(seed + k) mapped to [-1, 1).  tests/helpers.py restates the same hash in numpy.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ double hash_unit(uint64_t seed, uint64_t k) {
    return (double)(splitmix64(seed + k) >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}
template <typename A>
__global__ void __launch_bounds__(kThreads) k_fill_random(A* __restrict__ psi, uint64_t count, uint64_t support, uint64_t seed,
                                                          uint64_t index_base, double scale) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint64_t gi = index_base + i;
        double2 v = make_double2(0.0, 0.0);
        if (gi < support) v = make_double2(scale * hash_unit(seed, 2 * gi), scale * hash_unit(seed, 2 * gi + 1));
        psi[i] = from_c128<A>(v);
    }
}

// Scatter form of a permutation sweep for a source that is a small live window (a basis state that has only been
// through one tile: 64 or 4096 amplitudes, everything else implied zeros).  The gather evaluates f^-1 for every one of
// the 2^n outputs only to find that almost all of them are zero -- index arithmetic that makes the sweep ALU-bound
// (ncu, config 3: SM throughput 97 %, DRAM 69 %).  Here the output is zero-filled at write bandwidth and THIS kernel
// walks the few window amplitudes FORWARD through the run and drops each at its destination: out[f(i)] = s * in[i].
// f is a bijection, so no two threads write the same amplitude.  Sharded form: blockIdx.y is the source rank; every
// rank walks every window amplitude (read over NVLink: a few KiB) and keeps the ones that land in its own shard.
// The walk is one dependent chain per amplitude, so its latency per gate is what counts: the gate list is staged through
// shared memory in chunks (reading it from global memory cost ~140 ns per gate: 0.2 ms for the 1434-gate run of config
// 2 on a 16-CTA grid) and a gate is applied branch-free.
constexpr int kScatterChunk = 512;
template <typename A>
__global__ void __launch_bounds__(kThreads) k_perm_scatter(const __grid_constant__ DistPermArgs a, const DistGate* __restrict__ gates,
                                                           int n_gates, int nl) {
    __shared__ DistGate s_gates[kScatterChunk];
    const unsigned r = blockIdx.y;
    const uint64_t lmask = (1ull << nl) - 1ull;
    const uint64_t free_mask = ~a.wmask[r] & lmask;
    const int free_bits = __popcll(free_mask);
    if ((a.wval[r] & ~a.wmask[r]) != 0) return;  // empty window (a shard that holds no amplitude): uniform over the CTA
    const uint64_t count = 1ull << free_bits;
    A* out = reinterpret_cast<A*>(a.out);
    const A* in = reinterpret_cast<const A*>(a.src[r]);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t k0 = (uint64_t)blockIdx.x * blockDim.x; k0 < count; k0 += stride) {  // uniform trip count per CTA
        const uint64_t k = k0 + threadIdx.x;
        uint64_t src = a.wval[r], m = free_mask, kk = k;  // deposit the bits of k at the free positions
        while (m) {
            const uint64_t low = m & (0ull - m);
            if (kk & 1ull) src |= low;
            kk >>= 1;
            m ^= low;
        }
        uint64_t idx = src | ((uint64_t)(r ^ a.flip_in) << nl);  // logical index of the source amplitude
        for (int g0 = 0; g0 < n_gates; g0 += kScatterChunk) {    // forward: execution order
            const int nb = n_gates - g0 < kScatterChunk ? n_gates - g0 : kScatterChunk;
            __syncthreads();
            for (int i = threadIdx.x; i < nb; i += blockDim.x) {
                const uint4 w = __ldg(reinterpret_cast<const uint4*>(gates + g0 + i));
                *reinterpret_cast<uint4*>(&s_gates[i]) = w;
            }
            __syncthreads();
#pragma unroll 4
            for (int gi = 0; gi < nb; ++gi) {
                const DistGate g = s_gates[gi];
                const uint64_t b0 = 1ull << g.t0, b1 = 1ull << g.t1;
                // X: flip t0; SWAP: flip both when they differ
                const uint64_t flip = g.op == B2SV_OP_X ? b0 : ((((idx >> g.t0) ^ (idx >> g.t1)) & 1ull) ? (b0 | b1) : 0ull);
                idx ^= ((idx & g.ctrl) == g.ctrl) ? flip : 0ull;
            }
        }
        if (k >= count || (unsigned)(idx >> nl) != a.rank_logical) continue;  // beyond the window / lands on another rank
        B2SV_CHECK(src <= lmask);
        A v = in[src];
        if (a.has_scale) v = from_c128<A>(cmul(make_double2(a.sre[r], a.sim[r]), to_c128(v)));
        out[idx & lmask] = v;
    }
}

template <typename A>
__global__ void k_set_one(A* __restrict__ psi, uint64_t index) {
    psi[index] = from_c128<A>(make_double2(1.0, 0.0));
}

// ===================================================================================================
// 4. Projection onto a Subspace with all ancillae 0 (Subspace.test_basis / enumerate_basis / project,
//    /root/reference/src/unitaria/subspace.py:297-335) fused with the norm (nodes/node.py:390-403).
// ===================================================================================================

constexpr int kMaxSubInstr = 65535;  // program lives in global memory (per-thread divergent reads)

// member test + rank (position in ascending-index order) of index `bits`: a stack-free walk of the
// decision program (see include/b2sv.h).  `max_steps` bounds the walk so that a malformed program
// cannot hang the device.
__device__ __forceinline__ bool sub_member_rank(const b2sv_subinstr* __restrict__ prog, int entry, int max_steps,
                                                uint64_t bits, uint64_t& rank_out) {
    uint64_t r = 0;
    int pc = entry;
    for (int step = 0; step < max_steps; ++step) {
        const uint4 h = __ldg(reinterpret_cast<const uint4*>(prog + pc));
        const uint32_t op = h.x & 0xFFu, pos = (h.x >> 8) & 0xFFu, len = (h.x >> 16) & 0xFFu;
        const uint64_t field = (bits >> pos) & (len >= 64 ? ~0ull : ((1ull << len) - 1ull));
        if (op == B2SV_SUB_END) {
            rank_out = r;
            return true;
        }
        if (op == B2SV_SUB_REJECT) return false;
        if (op == B2SV_SUB_ZEROS) {
            if (field) return false;
            pc = (int)(h.y & 0xFFFFu);
        } else if (op == B2SV_SUB_FREE) {
            r += field * ((uint64_t)h.z | ((uint64_t)h.w << 32));
            pc = (int)(h.y & 0xFFFFu);
        } else {  // BRANCH on bit `pos`
            if ((bits >> pos) & 1ull) {
                r += __ldg(&prog[pc].add);
                pc = (int)(h.y >> 16);
            } else {
                pc = (int)(h.y & 0xFFFFu);
            }
        }
    }
    return false;
}

// GROUPED: the program starts with a FREE instruction on the low bits [0, len >= 3), so the 8 amplitudes of
// an aligned group are members together (and their ranks are consecutive multiples of that instruction's
// weight): the decision program is walked once per group instead of once per amplitude -- for the
// "0..0#..#" subspaces of block encodings the walk, not the memory, was the bound (ncu: 17 % DRAM, 68 % SM).
// A warp takes 256 consecutive amplitudes: lane l walks group l, the verdicts are exchanged by ballot /
// shuffle, and the loads stay fully coalesced (lane l reads amplitude j*32 + l in step j).
constexpr int kProjectGroup = 8;        // amplitudes per group
constexpr int kProjectWarpSpan = 256;   // amplitudes a warp handles per iteration (32 groups)

// `scan`: positions below the scanned range that every member has 0 (ZEROS instructions every walk passes): the
// non-grouped kernels enumerate only the indices with those bits clear -- a "000###...###000"-shaped subspace
// (BASELINE config 3: 2^20 members strided by 8 in a 2^26 range) needs 2^20 walks, not 2^26.
template <typename A, bool GROUPED>
__global__ void __launch_bounds__(kThreads) k_project_norm2(const A* __restrict__ psi, const b2sv_subinstr* __restrict__ sub,
                                                            uint64_t count, int entry, int max_steps, uint64_t index_base,
                                                            double* __restrict__ partial, const ExpandArgs scan) {
    double acc = 0.0;
    if (GROUPED) {
        const int lane = threadIdx.x & 31;
        const uint64_t warps = (uint64_t)gridDim.x * (kThreads / 32);
        for (uint64_t w = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); w * kProjectWarpSpan < count; w += warps) {
            const uint64_t base = w * kProjectWarpSpan;  // count is a multiple of kProjectWarpSpan
            uint64_t r;
            const bool mine = sub_member_rank(sub, entry, max_steps, index_base | (base + (uint64_t)lane * kProjectGroup), r);
            const uint32_t members = __ballot_sync(0xffffffffu, mine);
            if (members == 0u) continue;  // warp-uniform
            double2 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int group = j * 4 + (lane >> 3);
                v[j] = ((members >> group) & 1u) ? to_c128(psi[base + (uint64_t)(j * 32 + lane)]) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fma(v[j].x, v[j].x, fma(v[j].y, v[j].y, acc));
        }
    } else {
        // four indices per iteration: their membership walks first, then the four loads together (a per-rank
        // specialised program is often a single test, and one 16-byte load in flight per thread starves HBM:
        // 3.6 ms for a 16 GiB shard on 8 GPUs before, the copy floor is 2.5 ms)
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
        for (uint64_t k0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k0 < count; k0 += 4 * stride) {
            uint64_t idx[4];
            bool mem[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint64_t k = k0 + (uint64_t)u * stride;
                uint64_t r;
                idx[u] = expand_bits(k < count ? k : 0, scan);
                mem[u] = k < count && sub_member_rank(sub, entry, max_steps, index_base | idx[u], r);
            }
            double2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = mem[u] ? to_c128(psi[idx[u]]) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 4; ++u) acc = fma(v[u].x, v[u].x, fma(v[u].y, v[u].y, acc));
        }
    }
    __shared__ double red[kThreads / 32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

// deterministic final reduction of the per-block partial sums (one block)
// Post-selection norm of a BRANCH-FREE subspace ("0" on some qubits, "#" on the rest: every from_dim window, every
// one-ancilla block encoding, every per-rank program of such a subspace): the members are exactly the indices whose
// always-zero bits are clear, so no membership walk is needed at all -- a strided streaming reduction with eight loads in
// flight per thread.  (The general kernel walks the decision program per index or per group of 8 and was walk-bound on
// these: 2.5-3.4 TB/s for a 30-qubit register with one ancilla anywhere but at the top, scripts/microbench_project.py.)
template <typename A>
__global__ void __launch_bounds__(kThreads) k_norm2_scan(const A* __restrict__ psi, uint64_t count, double* __restrict__ partial,
                                                         const ExpandArgs scan) {
    double acc = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; k + 7 * stride < count; k += 8 * stride) {
        A v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = psi[expand_bits(k + (uint64_t)u * stride, scan)];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const double2 w = to_c128(v[u]);
            acc = fma(w.x, w.x, fma(w.y, w.y, acc));
        }
    }
    for (; k < count; k += stride) {
        const double2 w = to_c128(psi[expand_bits(k, scan)]);
        acc = fma(w.x, w.x, fma(w.y, w.y, acc));
    }
    __shared__ double red[kThreads / 32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < kThreads / 32 ? red[threadIdx.x] : 0.0;
        for (int o = 4; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) partial[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(kThreads) k_reduce_partials(const double* __restrict__ partial, int n, double* out) {
    __shared__ double red[kThreads];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += kThreads) acc += partial[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kThreads / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

// SPARSE: instead of filling a dense result, append (rank, value) of every NON-ZERO projected amplitude to a compact list
// (b2sv_project_gather_sparse): a column of a block encoding often has a handful of non-zeros out of 2^26 (three for the
// Laplacian), and hauling the zeros over PCIe was 80 % of an end-to-end Node.simulate call.  Once the list is full the
// threads stop taking slots (the host sees count >= cap and falls back to the dense path).
struct SparseOut {
    unsigned long long* count;
    unsigned long long cap;
    unsigned long long* idx;
};
__device__ __forceinline__ void sparse_emit(const SparseOut& sp, double2* __restrict__ val, uint64_t r, double2 v) {
    if (v.x == 0.0 && v.y == 0.0) return;
    if (*reinterpret_cast<volatile unsigned long long*>(sp.count) >= sp.cap) return;
    const unsigned long long pos = atomicAdd(sp.count, 1ull);
    if (pos < sp.cap) {
        sp.idx[pos] = r;
        val[pos] = v;
    }
}

template <typename A, bool GROUPED, bool SPARSE = false>
__global__ void __launch_bounds__(kThreads) k_project_gather(const A* __restrict__ psi, const b2sv_subinstr* __restrict__ sub,
                                                             uint64_t count, int entry, int max_steps,
                                                             double2 scale, double2* __restrict__ out, uint64_t dim, uint64_t weight,
                                                             const ExpandArgs scan, const SparseOut sp = SparseOut{nullptr, 0, nullptr}) {
    if (GROUPED) {
        const int lane = threadIdx.x & 31;
        const uint64_t warps = (uint64_t)gridDim.x * (kThreads / 32);
        for (uint64_t w = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); w * kProjectWarpSpan < count; w += warps) {
            const uint64_t base = w * kProjectWarpSpan;
            uint64_t r = 0;
            const bool mine = sub_member_rank(sub, entry, max_steps, base + (uint64_t)lane * kProjectGroup, r);
            const uint32_t members = __ballot_sync(0xffffffffu, mine);
            if (members == 0u) continue;  // warp-uniform
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int group = j * 4 + (lane >> 3);
                // rank of the group's first amplitude, from the lane that walked it
                const unsigned long long rg = __shfl_sync(0xffffffffu, (unsigned long long)r, group);
                if ((members >> group) & 1u) {
                    const uint64_t rr = rg + (uint64_t)(lane & 7) * weight;
                    if (rr < dim) {
                        const double2 v = cmul(scale, to_c128(psi[base + (uint64_t)(j * 32 + lane)]));
                        if (SPARSE) sparse_emit(sp, out, rr, v);
                        else out[rr] = v;
                    }
                }
            }
        }
    } else {
        const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
        for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < count; k += stride) {
            const uint64_t i = expand_bits(k, scan);
            uint64_t r;
            if (sub_member_rank(sub, entry, max_steps, i, r) && r < dim) {
                const double2 v = cmul(scale, to_c128(psi[i]));
                if (SPARSE) sparse_emit(sp, out, r, v);
                else out[r] = v;
            }
        }
    }
}

// The inverse of the gather: psi_i = in[rank(i)] for the members of the subspace, 0 everywhere else
// (the embedding BlockEncoding.compute builds on the host, nodes/block_encoding.py:49-52); `count` is
// the index range that can hold members, `total` the whole register.
template <typename A>
__global__ void __launch_bounds__(kThreads) k_embed(A* __restrict__ psi, const b2sv_subinstr* __restrict__ sub, uint64_t total,
                                                    uint64_t count, int entry, int max_steps, const double2* __restrict__ in,
                                                    uint64_t dim) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        uint64_t r;
        double2 v = make_double2(0.0, 0.0);
        if (i < count && sub_member_rank(sub, entry, max_steps, i, r) && r < dim) v = __ldg(in + r);
        psi[i] = from_c128<A>(v);
    }
}

__global__ void __launch_bounds__(kThreads) k_enumerate_basis(const b2sv_subinstr* __restrict__ sub, uint64_t count, int entry, int max_steps,
                                                              long long* __restrict__ out,
                                                              uint64_t dim) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint64_t r;
        if (sub_member_rank(sub, entry, max_steps, i, r) && r < dim) out[r] = (long long)i;
    }
}

// ===================================================================================================
// 5. State initialisation, dtype conversion, scaling, global<->local swap packing.
// ===================================================================================================

template <typename A>
__global__ void __launch_bounds__(kThreads) k_set_basis(A* __restrict__ psi, uint64_t count, uint64_t index) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        psi[i] = from_c128<A>(make_double2(i == index ? 1.0 : 0.0, 0.0));
}

// Zero everything outside the live window {i : (i & mask) == value} (see b2sv.cu: a basis state that was only
// ever touched inside one tile is kept as "that tile + implied zeros" until somebody needs the whole state).
template <typename A>
__global__ void __launch_bounds__(kThreads) k_window_fill(A* __restrict__ psi, uint64_t count, uint64_t mask, uint64_t value) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const A zero = from_c128<A>(make_double2(0.0, 0.0));
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        if ((i & mask) != value) psi[i] = zero;
}

template <typename A>
__global__ void __launch_bounds__(kThreads) k_scale(A* __restrict__ psi, uint64_t count, double2 s) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        psi[i] = from_c128<A>(cmul(s, to_c128(psi[i])));
}

// dst[i] = convert(src[i]); used for complex64 storage <-> complex128 host buffers
template <typename D, typename S>
__global__ void __launch_bounds__(kThreads) k_convert(D* __restrict__ dst, const S* __restrict__ src, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        dst[i] = from_c128<D>(to_c128(src[i]));
}

// staging[k] = s * psi[i] (pack) or psi[i] = s * staging[k] (unpack) for the half of the shard whose
// bit q equals `sel`; k enumerates that half in ascending index order.  `s` carries the shard's lazy
// global scalar across the exchange (pack: s, unpack: conj(s)) so that no separate scaling pass over the
// whole shard is needed; SCALE = false is the exact data-movement form used when s == 1.
template <typename A, bool PACK, bool SCALE>
__global__ void __launch_bounds__(kThreads) k_swap_pack(A* __restrict__ psi, A* __restrict__ staging, uint64_t first,
                                                        uint64_t half, int q, uint64_t sel_bit, double2 s) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 4;
    for (uint64_t k0 = first + ((uint64_t)blockIdx.x * blockDim.x * 4) + threadIdx.x; k0 < half; k0 += stride) {
        A v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t k = k0 + (uint64_t)u * blockDim.x;
            if (k < half) v[u] = PACK ? psi[insert_zero(k, q) | sel_bit] : staging[k];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t k = k0 + (uint64_t)u * blockDim.x;
            if (k < half) {
                const A w = SCALE ? from_c128<A>(cmul(s, to_c128(v[u]))) : v[u];
                if (PACK) staging[k] = w;
                else psi[insert_zero(k, q) | sel_bit] = w;
            }
        }
    }
}

}  // namespace b2svk
