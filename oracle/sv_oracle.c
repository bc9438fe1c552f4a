/* CPU oracle (plain C + OpenMP) for the unitaria simulation hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (unitaria_b200/) never links or calls it.
 *
 * It restates, gate by gate, what the reference's simulate call computes
 * (/root/reference/src/unitaria/circuit.py:62-70 -> tq.simulate(..., backend="qulacs")):
 * a complex128 statevector of 2^n amplitudes in LSB order (index bit q <-> qubit q), updated in
 * place with ONE sweep per gate over the 2^(n-c) amplitudes whose c control bits are all 1 and
 * no gate fusion -- the schedule of qulacs 0.6.13's per-gate update functions (third-party,
 * pinned at /root/reference/uv.lock:1598-1599, not vendored and not installable here, hence a
 * restatement of its published algorithm rather than a build of it).  Gate conventions are the
 * tequila ones; they are pinned against the reference's numpy compute() path
 * through oracle/np_oracle.py (tests/test_oracle.py checks C == numpy on random circuits, and
 * numpy == reference goldens).
 *
 * Parity status: pinned vs the reference's numpy toarray()/compute() goldens; UNPINNED vs qulacs.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex c128;

/* gate kinds of the oracle's own wire format (independent of the product's lowering) */
enum { ORC_I = 0, ORC_X, ORC_Y, ORC_Z, ORC_H, ORC_RX, ORC_RY, ORC_RZ, ORC_PHASE, ORC_GPHASE, ORC_SWAP };

typedef struct {
    int32_t kind;
    int32_t t0, t1;      /* targets (t1 only for SWAP; -1 otherwise) */
    uint64_t ctrl_mask;  /* all of these bits must be 1 */
    double param;        /* angle, or power for X/Y/Z/H */
} orc_gate;

/* bench.py --impl reference runs under torchrun, which exports OMP_NUM_THREADS=1: the CPU arm asks for the box's cores */
void orc_set_threads(int n) {
    if (n > 0) omp_set_num_threads(n);
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* insert a 0 bit at position p */
static inline uint64_t ins0(uint64_t x, int p) {
    uint64_t lo = x & ((1ull << p) - 1);
    return ((x >> p) << (p + 1)) | lo;
}

/* Expand a compact counter k (over the free bits) to a full index with zeros at the sorted
 * positions pos[0..m). */
static inline uint64_t expand(uint64_t k, const int *pos, int m) {
    for (int i = 0; i < m; ++i) k = ins0(k, pos[i]);
    return k;
}

static int sorted_positions(uint64_t mask, int *pos) {
    int m = 0;
    for (int q = 0; q < 64; ++q)
        if ((mask >> q) & 1) pos[m++] = q;
    return m;
}

static void unitary_of(const orc_gate *g, c128 m[4]) {
    double t = g->param / 2.0;
    switch (g->kind) {
        case ORC_RX: m[0] = cos(t); m[1] = -I * sin(t); m[2] = -I * sin(t); m[3] = cos(t); return;
        case ORC_RY: m[0] = cos(t); m[1] = -sin(t); m[2] = sin(t); m[3] = cos(t); return;
        case ORC_RZ: m[0] = cexp(-I * t); m[1] = 0; m[2] = 0; m[3] = cexp(I * t); return;
        case ORC_PHASE: m[0] = 1; m[1] = 0; m[2] = 0; m[3] = cexp(I * g->param); return;
        default: break;
    }
    c128 b[4];
    const double s = 1.0 / sqrt(2.0);
    switch (g->kind) {
        case ORC_X: b[0] = 0; b[1] = 1; b[2] = 1; b[3] = 0; break;
        case ORC_Y: b[0] = 0; b[1] = -I; b[2] = I; b[3] = 0; break;
        case ORC_Z: b[0] = 1; b[1] = 0; b[2] = 0; b[3] = -1; break;
        case ORC_H: b[0] = s; b[1] = s; b[2] = s; b[3] = -s; break;
        default: b[0] = 1; b[1] = 0; b[2] = 0; b[3] = 1; break;
    }
    if (g->param == 1.0) { memcpy(m, b, sizeof(b)); return; }
    /* U^p = P+ + e^{i pi p} P-,  P+- = (1 +- U)/2 */
    c128 w = cexp(I * M_PI * g->param);
    m[0] = 0.5 * (1 + b[0]) + w * 0.5 * (1 - b[0]);
    m[1] = 0.5 * b[1] - w * 0.5 * b[1];
    m[2] = 0.5 * b[2] - w * 0.5 * b[2];
    m[3] = 0.5 * (1 + b[3]) + w * 0.5 * (1 - b[3]);
}

/* One gate, one sweep over the controlled amplitudes. */
void orc_apply_gate(c128 *psi, int n, const orc_gate *g) {
    int pos[66];
    if (g->kind == ORC_I) return;
    if (g->kind == ORC_GPHASE) {
        int m = sorted_positions(g->ctrl_mask, pos);
        uint64_t cnt = 1ull << (n - m);
        c128 w = cexp(I * g->param);
        uint64_t cm = g->ctrl_mask;
#pragma omp parallel for schedule(static)
        for (uint64_t k = 0; k < cnt; ++k) {
            uint64_t i = expand(k, pos, m) | cm;
            psi[i] *= w;
        }
        return;
    }
    if (g->kind == ORC_SWAP) {
        uint64_t ba = 1ull << g->t0, bb = 1ull << g->t1;
        int m = sorted_positions(g->ctrl_mask | ba | bb, pos);
        uint64_t cnt = 1ull << (n - m);
        uint64_t cm = g->ctrl_mask;
#pragma omp parallel for schedule(static)
        for (uint64_t k = 0; k < cnt; ++k) {
            uint64_t i = expand(k, pos, m) | cm;
            c128 tmp = psi[i | ba];
            psi[i | ba] = psi[i | bb];
            psi[i | bb] = tmp;
        }
        return;
    }
    uint64_t bt = 1ull << g->t0;
    int m = sorted_positions(g->ctrl_mask | bt, pos);
    uint64_t cnt = 1ull << (n - m);
    uint64_t cm = g->ctrl_mask;
    if (g->kind == ORC_X && g->param == 1.0) { /* exact data movement */
#pragma omp parallel for schedule(static)
        for (uint64_t k = 0; k < cnt; ++k) {
            uint64_t i = expand(k, pos, m) | cm;
            c128 tmp = psi[i];
            psi[i] = psi[i | bt];
            psi[i | bt] = tmp;
        }
        return;
    }
    c128 u[4];
    unitary_of(g, u);
    if (u[1] == 0 && u[2] == 0) {
        const c128 d0 = u[0], d1 = u[3];
        const int skip0 = (d0 == 1.0);
#pragma omp parallel for schedule(static)
        for (uint64_t k = 0; k < cnt; ++k) {
            uint64_t i = expand(k, pos, m) | cm;
            if (!skip0) psi[i] *= d0;
            psi[i | bt] *= d1;
        }
        return;
    }
#pragma omp parallel for schedule(static)
    for (uint64_t k = 0; k < cnt; ++k) {
        uint64_t i = expand(k, pos, m) | cm;
        c128 a0 = psi[i], a1 = psi[i | bt];
        psi[i] = u[0] * a0 + u[1] * a1;
        psi[i | bt] = u[2] * a0 + u[3] * a1;
    }
}

void orc_apply(c128 *psi, int n, const orc_gate *gates, int64_t n_gates) {
    for (int64_t j = 0; j < n_gates; ++j) orc_apply_gate(psi, n, &gates[j]);
}

void orc_set_basis(c128 *psi, int n, uint64_t index) {
    uint64_t cnt = 1ull << n;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < cnt; ++i) psi[i] = 0;
    psi[index] = 1;
}

/* sum |psi_i|^2 over i < 2^t with member[i] != 0 (Subspace.project + norm, nodes/node.py:379,402) */
double orc_project_norm2(const c128 *psi, const uint8_t *member, int t) {
    uint64_t cnt = 1ull << t;
    double acc = 0;
#pragma omp parallel for schedule(static) reduction(+ : acc)
    for (uint64_t i = 0; i < cnt; ++i)
        if (member[i]) acc += creal(psi[i]) * creal(psi[i]) + cimag(psi[i]) * cimag(psi[i]);
    return acc;
}
