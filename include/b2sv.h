/* b2sv.h -- C ABI of the B200 statevector engine that sits behind unitaria's simulate seam.
 *
 * The reference (tequilahub/unitaria v0.2.0, pure Python) has no FFI of its own; its only seam to a
 * simulator is the Python call
 *     tq.simulate(padded, initial_state=wfn, backend="qulacs").to_array(LSB)
 * made from Circuit.simulate (src/unitaria/circuit.py:62-70), followed by Subspace.project and the
 * normalization factor in Node.simulate / Node.simulate_norm (src/unitaria/nodes/node.py:363-403,
 * src/unitaria/subspace.py:297-335).  Each entry point below names the piece of that path it
 * replaces.  The ctypes binding a maintainer would add is shown in INTEGRATION.md and shipped as
 * unitaria_b200/cabi.py.
 *
 * Conventions
 *   - amplitudes are little-endian in the qubit index ("LSB", circuit.py:63-70): bit q of an
 *     amplitude's index is the state of qubit q.
 *   - host amplitude buffers are always complex128 (interleaved re,im doubles), whatever the
 *     device storage dtype is; they are borrowed for the duration of the call only.
 *   - every function returns 0 on success or a negative B2SV_E* code; b2sv_last_error() gives the
 *     message for the calling thread.  Nothing throws across the ABI, there are no callbacks.
 *   - a handle is not thread-safe; different handles may be used from different threads.
 *   - there is no CPU fallback: without a usable CUDA device every compute entry point fails with
 *     B2SV_ENODEV.
 */
#ifndef B2SV_H
#define B2SV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2SV_ABI_VERSION 2

typedef struct b2sv b2sv_t;

enum {
    B2SV_OK = 0,
    B2SV_EINVAL = -1,  /* bad argument */
    B2SV_ENODEV = -2,  /* no CUDA device / CUDA runtime failure at start-up */
    B2SV_ENOMEM = -3,  /* device or host allocation failed */
    B2SV_ECUDA = -4,   /* a CUDA call or kernel failed */
    B2SV_ESTATE = -5   /* call not valid in the handle's current state */
};

enum { B2SV_C128 = 0, B2SV_C64 = 1 }; /* device storage dtype of the amplitudes */

/* Lowered gate alphabet.  unitaria_b200/lowering.py maps the tequila gates unitaria emits
 * (I X Y Z H Rx Ry Rz Phase GlobalPhase SWAP, any control set) onto these. */
enum {
    B2SV_OP_MAT1 = 0,  /* dense 2x2 on t0: m = {m00,m01,m10,m11} as (re,im) pairs, row-major      */
    B2SV_OP_DIAG1 = 1, /* diag(d0,d1) on t0: m[0..1] = d0, m[2..3] = d1                            */
    B2SV_OP_X = 2,     /* bit flip of t0 -- exact data movement, no arithmetic on the amplitudes    */
    B2SV_OP_SWAP = 3,  /* exchange qubits t0,t1 -- exact data movement                              */
    B2SV_OP_PHASE = 4  /* no target: multiply the controlled block by m[0..1] (GlobalPhase)         */
};

/* One lowered gate; applied iff all bits of ctrl_mask are 1 in the amplitude's index. */
typedef struct b2sv_gate {
    uint8_t op;
    uint8_t n_targets; /* 0 (PHASE), 1, or 2 (SWAP) */
    uint8_t t0, t1;
    uint32_t reserved;
    uint64_t ctrl_mask;
    double m[8];
} b2sv_gate;

/* Planner / execution options for b2sv_apply (NULL = defaults). */
typedef struct b2sv_plan_opts {
    int32_t fuse;         /* 0: one kernel launch per gate (qulacs-like schedule); 1: fused sweeps (default) */
    int32_t tile_qubits;  /* qubits per shared-memory tile of a fused sweep; 0 = default             */
    int32_t perm_runs;    /* 1: runs of X/SWAP gates become one index-map sweep (default); 0: off    */
    int32_t timing;       /* 1: bracket every launch with CUDA events and accumulate per-class times */
    int32_t lookahead;    /* commutation-aware lookahead window in gates; 0 = default                 */
    int32_t tile_io;      /* tile sweeps: 0 = one tile per CTA, TMA tensor copies with the 128-byte swizzle, 3 CTAs/SM
                             (default); 1 = plain ld/st (debug aid); 2 = linear bulk-async copies, no swizzle */
    int32_t jit;          /* permutation runs: 0 = specialise with NVRTC in the background after first use
                             (default), 1 = never (generic kernel only), 2 = compile synchronously       */
    int32_t reserved;     /* A/B switches: bit 0 = no multiplexor / block fusion, bit 1 = no register phases,
                             bit 2 = no warp-shuffle low-qubit sweeps, bit 3 = no monomial phases, bit 4 = no SWAP frame;
                             bit 5 = "a permutation sweep follows this call" (the distributed gather): a pending basis
                             state that has only been through one tile may stay a live window */
} b2sv_plan_opts;

/* Subspace program (restates Subspace.test_basis / enumerate_basis,
 * subspace.py:297-329).  unitaria's Subspace is a tree of tensor factors; because it is a tree, the
 * absolute bit position and the mixed-radix rank weight of every factor are static, so the host
 * (unitaria_b200/subspace_prog.py) compiles it into a stack-free decision program:
 *   B2SV_SUB_END     accept; the accumulated rank is the index's position in enumerate_basis order
 *   B2SV_SUB_ZEROS   reject unless bits [pos, pos+len) are all 0              (ZeroQubitSubspace run)
 *   B2SV_SUB_FREE    rank += field(pos, len) * weight                          (run of free qubits)
 *   B2SV_SUB_BRANCH  bit `pos` (the top qubit of a ControlledSubspace) picks next1 (and rank += add)
 *                    or next0
 *   B2SV_SUB_REJECT  reject (used by per-rank specialised programs of the distributed layer, where a
 *                    bit that lives in the rank number is known on the host)
 * Evaluation starts at instruction `entry`; ZEROS/FREE continue at next0. */
enum { B2SV_SUB_END = 0, B2SV_SUB_ZEROS = 1, B2SV_SUB_FREE = 2, B2SV_SUB_BRANCH = 3, B2SV_SUB_REJECT = 4 };

typedef struct b2sv_subinstr {
    uint8_t op, pos, len, pad;
    uint16_t next0, next1;
    uint64_t weight;
    uint64_t add;
    uint64_t reserved;
} b2sv_subinstr;

/* Counters filled by b2sv_stats (accumulated since creation or the last b2sv_stats_reset). */
typedef struct b2sv_stats_t {
    uint64_t gates_in;        /* lowered gates handed to b2sv_apply                                   */
    uint64_t gates_live;      /* ... of which not exact no-ops                                        */
    uint64_t launches;        /* kernel launches made by b2sv_apply                                   */
    uint64_t launches_by_class[12]; /* 0 gate1q, 1 tile sweep, 2 perm run, 3 scale, 4 project, 5 init/convert, 6 swap pack,
                                      7 perm run (NVRTC-specialised), 8 low-qubit warp-shuffle sweep,
                                      9 distributed permutation sweep (gather over peer-mapped shards) */
    double alg_bytes_by_class[12];  /* algorithmic bytes (2*B*M per sweep)                */
    double ms_by_class[12];        /* CUDA-event time per class (only with opts.timing)                */
    double apply_ms;               /* CUDA-event time of whole b2sv_apply calls (always measured)      */
} b2sv_stats_t;

/* ---- lifetime ---------------------------------------------------------------------------------- */

/* State of `n_qubits` qubits (2^n amplitudes) on CUDA device `device`.  Replaces the
 * QubitWaveFunction + qulacs QuantumState that tq.simulate allocates per call (circuit.py:62-69). */
int b2sv_create(int n_qubits, int dtype, int device, b2sv_t** out);
int b2sv_destroy(b2sv_t* sv);

/* ---- state I/O (circuit.py:62-65: from_basis_state / from_array; :70: to_array) ---------------- */
int b2sv_set_basis(b2sv_t* sv, uint64_t index);
int b2sv_set_zero(b2sv_t* sv); /* all amplitudes 0: a shard that does not own the basis state (lazy: nothing is
                                  written until a reader needs the zeros; gates on a zero state are skipped) */
int b2sv_upload(b2sv_t* sv, const void* host_amps_c128);
/* Seeded dense test vector generated on the device (benchmarks / full-size property tests; no counterpart in the
 * reference, which uploads numpy arrays): psi_i = scale * (u(2k), u(2k+1)), k = index_base + i, for k < 2^support_bits
 * and 0 above, u = splitmix64(seed + .) mapped to [-1,1), scale = sqrt(1.5 / 2^support_bits).  index_base lets the
 * shards of a distributed register hold one global vector. */
int b2sv_fill_random(b2sv_t* sv, uint64_t seed, int support_bits, uint64_t index_base);
int b2sv_download(b2sv_t* sv, void* host_out_c128);

/* ---- the simulate call itself (circuit.py:69, tq.simulate -> qulacs update_quantum_state) ------ */
int b2sv_apply(b2sv_t* sv, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts);

/* ---- fused post-processing (node.py:379 project*normalization, node.py:402 norm) --------------- */
/* sum of |psi_i|^2 over i in the subspace on the low `target_qubits` qubits with every higher
 * (ancilla) qubit 0. */
int b2sv_project_norm2(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                       double* out);
/* Sharded form: this handle holds the amplitudes index_base | i (i < 2^n_qubits, index_base's low
 * n_qubits bits clear) of a wider register; the subspace program spans `target_qubits` (<= 64) qubits
 * of the WIDE index.  The caller sums the per-shard results (unitaria_b200/distributed.py). */
int b2sv_project_norm2_at(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                          uint64_t index_base, double* out);
/* host_out[rank(i)] = (scale_re + i*scale_im) * psi_i for the members i, in ascending-index order
 * (= Subspace.enumerate_basis order, subspace.py:325-329); `dim` must equal the subspace dimension. */
int b2sv_project_gather(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                        double scale_re, double scale_im, void* host_out_c128, uint64_t dim);
/* Sparse form of the same result (replaces the dense `normalization * subspace_out.project(...)` array of
 * nodes/node.py:377-379 when almost all of it is zeros -- a column of the Laplacian block encoding has three non-zeros out
 * of 2^26): up to `cap` pairs (rank, value) of the NON-ZERO projected amplitudes, in no particular order.  *nnz_out is the
 * number of pairs written; *nnz_out >= cap means "too many, the lists are incomplete -- use b2sv_project_gather".  Only
 * 8 + 24 * nnz bytes cross PCIe. */
int b2sv_project_gather_sparse(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                               double scale_re, double scale_im, uint64_t dim, uint64_t cap, uint64_t* idx_out,
                               void* val_out_c128, uint64_t* nnz_out);
/* Pipelined form for callers that run one circuit on many inputs (the matrix case of Node.simulate,
 * nodes/node.py:381-388; verify, verifier.py:119-126): the gather lands in one of two device buffers (`slot` 0/1)
 * and is copied to `host_out` (page-locked memory for a real overlap) by a second stream, so the call returns at
 * once and the next input's sweeps run while the result crosses PCIe.  `host_out` must stay valid until
 * b2sv_gather_wait(slot) (or the next async call on the same slot) returns. */
int b2sv_project_gather_async(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                              double scale_re, double scale_im, void* host_out_c128, uint64_t dim, int slot);
int b2sv_gather_wait(b2sv_t* sv, int slot);
/* Device-resident hand-off (chained nodes, fixed-point / Newton loops of
 * examples/example_amplified_encodings.py:32-63 -- no host array in between): the projected vector stays in a device
 * buffer of the handle (complex128[dim], valid until the next projection on this handle; the call returns after the
 * kernel has finished, so any stream may read it), and a state can be loaded from a complex128 device buffer
 * (2^n amplitudes) of the same GPU. */
int b2sv_project_gather_device(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                               double scale_re, double scale_im, void** dev_out_c128, uint64_t dim);
int b2sv_upload_device(b2sv_t* sv, const void* dev_amps_c128);
/* The inverse hand-off: psi := sum_r host_in[r] |basis_r>, i.e. the members of the subspace (ascending
 * index order) take the `dim` amplitudes and every other amplitude is 0.  Replaces the 2^n-element host
 * array BlockEncoding.compute fills and uploads (nodes/block_encoding.py:45-52: full_input[basis] = input):
 * only `dim` amplitudes cross PCIe.  `dim` must equal the subspace dimension. */
int b2sv_embed(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
               const void* host_in_c128, uint64_t dim);
/* Same predicate on the host-visible side of enumerate_basis: writes the member indices (int64). */
int b2sv_enumerate_basis(b2sv_t* sv, const b2sv_subinstr* prog, int n_instr, int entry, int target_qubits,
                         int64_t* host_out, uint64_t dim);

/* ---- device-resident hand-off and the pieces the distributed layer drives ----------------------- */
int b2sv_device_ptr(b2sv_t* sv, void** amps, uint64_t* bytes);
int b2sv_sync(b2sv_t* sv);
/* Page-locked host memory for results that cross PCIe (cudaHostAlloc / cudaFreeHost): D2H into such a
 * buffer runs at link speed instead of being staged through the driver's bounce buffers. */
int b2sv_host_alloc(size_t bytes, void** out);
int b2sv_host_free(void* ptr);
/* CUDA-event stopwatch on the handle's stream (slots 0..7): b2sv_mark records an event after all work
 * enqueued so far; b2sv_elapsed_ms synchronises and returns the device time between two marks. */
int b2sv_mark(b2sv_t* sv, int slot);
int b2sv_elapsed_ms(b2sv_t* sv, int slot_a, int slot_b, double* ms);
/* Exchange support for global<->local qubit swaps (unitaria_b200/distributed.py): pack the half of
 * the shard whose local qubit `q` differs from `keep_bit` into `staging` (contiguous), and the
 * inverse. */
int b2sv_swap_pack(b2sv_t* sv, int q, int keep_bit, void* staging_dev);
int b2sv_swap_unpack(b2sv_t* sv, int q, int keep_bit, const void* staging_dev);
/* The same on packed amplitudes [first, first+count) only (staging is still indexed from 0), so that the
 * caller can pipeline pack / exchange / unpack in chunks. */
int b2sv_swap_pack_range(b2sv_t* sv, int q, int keep_bit, void* staging_dev, uint64_t first, uint64_t count);
int b2sv_swap_unpack_range(b2sv_t* sv, int q, int keep_bit, const void* staging_dev, uint64_t first, uint64_t count);

/* ---- distributed execution: one process per GPU, the register sharded by its top index bits ------------
 * (unitaria_b200/distributed.py drives these; the reference is single-process)
 * Every rank exports its two state buffers and two interprocess events as an opaque blob
 * (b2sv_ipc_export), the host side all-gathers the blobs, and b2sv_ipc_attach maps every peer's buffers into
 * this process (CUDA IPC; handles of the same process are attached by pointer).  After that a *distributed
 * permutation sweep* -- a run of X / multi-controlled X / (controlled) SWAP gates over the WHOLE register, targets
 * and controls on local or global index positions alike -- is ONE kernel per rank that gathers
 * out[j] = in[f^-1(rank:j)] straight out of the peers' HBM over NVLink: the all-to-all and the index map are
 * fused, nothing is packed or staged.  A global<->local swap of k qubits is the same sweep with SWAP gates.
 * Protocol per sweep (device-side ordering by interprocess events, the host never waits for the GPU):
 *   1. b2sv_dist_ready        records "my shard is final" and returns this rank's state words
 *   2. host all-gather of the state words (also guarantees every rank has recorded its event)
 *   3. b2sv_dist_perm         waits for the peers' events on the stream, launches the gather, records "done"
 *   4. host barrier           (every rank has recorded "done")
 *   5. b2sv_dist_done         the stream waits for the peers' "done" before anything overwrites what they read */
#define B2SV_IPC_BYTES 320
#define B2SV_DIST_STATE_WORDS 8
int b2sv_ipc_export(b2sv_t* sv, void* blob_out /* B2SV_IPC_BYTES */);
int b2sv_ipc_attach(b2sv_t* sv, int rank, int world, const void* blobs /* world * B2SV_IPC_BYTES */);
int b2sv_ipc_detach(b2sv_t* sv);
int b2sv_dist_ready(b2sv_t* sv, uint64_t* state_out /* B2SV_DIST_STATE_WORDS */);
/* gates: X / SWAP only, qubit numbers are index POSITIONS of the n_total-bit register (positions >= the
 * handle's n_qubits are rank bits).  flip_in / flip_out: per global position, whether the physical rank bit
 * is the complement of the logical bit before / after the sweep (an uncontrolled X on a global qubit is a
 * relabel).  states: every rank's words from step 1, in rank order. */
int b2sv_dist_perm(b2sv_t* sv, const b2sv_gate* gates, size_t n_gates, int n_total, uint32_t flip_in, uint32_t flip_out,
                   const uint64_t* states /* world * B2SV_DIST_STATE_WORDS */, const b2sv_plan_opts* opts);
int b2sv_dist_done(b2sv_t* sv);
/* Host-only: the CUDA source of the specialised gather kernel for `gates` (introspection / evidence). */
int b2sv_dist_source(int n_local, int n_global, int dtype, const b2sv_gate* gates, size_t n_gates, char* out, size_t cap,
                     size_t* need);

/* ---- planner introspection (no GPU needed): how would `gates` be executed? ---------------------- */
/* Writes up to `cap` int64 words describing the plan: for each segment
 *   {class, n_gates, tile_mask, common_ctrl, n_ops, gate_index_0, ..., gate_index_{n_gates-1}}
 * (a fused gate lists every caller gate it stands for; common_ctrl = controls shared by all gates
 * of a tile sweep outside its tile; n_ops = device ops after fusion), then the trailer
 * {-1, 2, 0, 0, 0, bits(scalar_re), bits(scalar_im)},
 * then {-2, count, 0, 0, 0, count x {gate id, op, t0, t1, ctrl_mask}}: the gates the planner's SWAP frame executes
 * on renamed qubits (a deferred SWAP renames instead of moving data) and the synthetic SWAPs (ids >= n_gates) that
 * restore the layout -- a gate id listed here is to be replayed with THESE qubits;
 * and returns the number of words needed in *n_words. */
int b2sv_plan_describe(int n_qubits, int dtype, const b2sv_gate* gates, size_t n_gates, const b2sv_plan_opts* opts,
                       int64_t* out_words, size_t cap, size_t* n_words);

/* Host-only check of the run-time specialisation: generates the CUDA source of the permutation-run
 * kernel for `gates` (all X / SWAP), compiles it with NVRTC for sm_100a and returns the cubin size. */
int b2sv_perm_jit_compile_check(int n_qubits, int dtype, const b2sv_gate* gates, size_t n_gates, size_t* cubin_bytes);

/* ---- bookkeeping ------------------------------------------------------------------------------- */
int b2sv_stats(b2sv_t* sv, b2sv_stats_t* out);
int b2sv_stats_reset(b2sv_t* sv);
int b2sv_abi_version(void);
int b2sv_device_count(void);
const char* b2sv_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* B2SV_H */
