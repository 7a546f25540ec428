/* spartan_b200.h -- C ABI of the B200-native r1cs-spartan prover hot path.
 *
 * Drop-in boundary for tsunrise/r1cs-spartan: each entry point below is what the reference's Rust
 * functions would bind through `extern "C"` (see INTEGRATION.md for the Rust-side stubs).  Reference
 * citations are file:line into the reference repository.
 *
 * Data conventions (identical to arkworks' in-memory layout, so Rust slices cross zero-copy):
 *   Fr  : 32 bytes, 4 x u64 little-endian limbs, MONTGOMERY form (R = 2^256), BLS12-381 scalar field.
 *   Fq  : 48 bytes, 6 x u64 little-endian limbs, Montgomery form (R = 2^384).
 *   G1 affine: x, y (96 bytes).  G2 affine: x.c0, x.c1, y.c0, y.c1 (192 bytes).
 *             The point at infinity is encoded as all-zero bytes (arkworks `infinity: true`).
 *   Variable order of every multilinear table is LSB-first: table[b], bit i of b <-> variable i
 *   (src/data_structures/eq.rs:11, src/data_structures/r1cs_reader.rs:22-24, src/commitment/open.rs:43-44).
 *
 * Ownership: the caller owns every host buffer, for the duration of the call only.  Handles own
 * device memory, are immutable after creation (except sb_prover) and are freed by *_destroy.
 * Threading: one call at a time per sb_ctx.  All calls are synchronous.
 * Errors: every function returns an sb_status; outputs are untouched on error; sb_last_error()
 * gives the message.  No C++ exception crosses this boundary.  There is no CPU fallback: without a
 * CUDA device sb_ctx_create fails with SB_ECUDA.
 */
#ifndef SPARTAN_B200_H
#define SPARTAN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    SB_OK = 0,
    SB_EINVAL = 1,    /* reference: Error::InvalidArgument, src/error.rs:5-14 */
    SB_ECUDA = 2,
    SB_ENOMEM = 3,
    SB_ECOMM = 4,     /* collective exchange failed (multi-GPU) */
    SB_EINTERNAL = 5
} sb_status;

typedef struct sb_ctx sb_ctx;
typedef struct sb_index sb_index;     /* reference: IndexPK, src/ahp/indexer.rs:11-17 */
typedef struct sb_pp sb_pp;           /* reference: PublicParameter, src/commitment/data_structures.rs:10-17 */
typedef struct sb_prover sb_prover;   /* reference: Prover*State, src/ahp/prover.rs:25-64 */
typedef struct sb_witness sb_witness; /* z = v || w kept in HBM between proofs (no reference counterpart) */

/* One R1CS matrix in CSR form: row_ptr[n+1], col[nnz] (< n), val[nnz] (Fr).  The Rust shim flattens
 * ark_relations::r1cs::Matrix = Vec<Vec<(F, usize)>> into this (tuple layout is unspecified in Rust). */
typedef struct {
    const uint64_t* row_ptr;
    const uint32_t* col;
    const void* val;
} sb_csr;

/* Exchange hooks for the hypercube-sharded prover (one process per GPU).  allgather must copy
 * `bytes` bytes from every rank's `send` into `recv` (rank-major, world * bytes).  NULL = single GPU. */
typedef struct {
    int rank;
    int world;
    int (*allgather)(void* user, const void* send, void* recv, size_t bytes);
    /* optional (may be NULL): called by every rank at the start of every library call on the context, so that an
     * exchange layer with sequence numbers can re-align ranks after a call that failed on some of them */
    int (*barrier)(void* user);
    void* user;
} sb_comm;

/* A ready-made sb_comm for ranks on one node: allgather through a POSIX shared-memory mailbox (about a
 * microsecond; the payloads are at most a few hundred bytes and already on the host).  `name` is a shm name
 * ("/something") shared by all ranks; exactly one rank passes create != 0 and must do so before the others attach. */
sb_status sb_comm_shm_open(const char* name, int rank, int world, int create, sb_comm* out);
/* The same mailbox in process memory for `world` contexts driven by threads of one process; fills out[0..world). */
sb_status sb_comm_local_open(int world, sb_comm* out);
/* a rank whose library call failed locally releases the peers that wait for it in an exchange (they fail that call at once
 * instead of after SB_COMM_TIMEOUT_S); the multi-GPU context does this for its shards by itself */
void sb_comm_shm_abort(sb_comm* comm);
void sb_comm_shm_close(sb_comm* comm);

/* ---- context ---------------------------------------------------------------------------------- */
sb_status sb_ctx_create(int device, sb_ctx** out);
sb_status sb_ctx_create_sharded(int device, const sb_comm* comm, sb_ctx** out);
/* ONE context over several GPUs of this process (ndev a power of two): the hypercube-sharded prover with one host thread per
 * GPU inside the library and an in-process exchange, so that MLArgumentForR1CS::prove (src/lib.rs:58) stays one call from
 * one process.  Handles made from it (index, parameters, witness, prover state) hold one part per GPU; index / keygen /
 * load / commit / open / prove / the prover rounds work as on a single-GPU context and give the same bytes.  The
 * single-matrix helpers (sb_msm, sb_eq_table, sb_sum_over_y, sb_eval_on_x, sb_pp_export) need a single-GPU context. */
sb_status sb_ctx_create_multi(const int* devices, int ndev, sb_ctx** out);
void sb_ctx_destroy(sb_ctx* ctx);
const char* sb_last_error(const sb_ctx* ctx);   /* ctx may be NULL: last error of a failed sb_ctx_create */
/* number of CUDA kernels this library has launched in this process (bench.py: gpu_launches) */
uint64_t sb_launch_count(void);
/* CUDA devices visible to this process (0 when there is none: every sb_ctx_create* then fails with SB_ECUDA) */
int sb_device_count(void);

/* ---- indexer: MLArgumentForR1CS::index, src/lib.rs:45-51 -> src/ahp/indexer.rs:41-64 ------------ */
/* Same checks as the reference: n = 2^log_n rows per matrix (indexer.rs:49, r1cs_reader.rs:38-52),
 * every column index < n (r1cs_reader.rs:55-62) -> SB_EINVAL otherwise.  Also absorbs the three
 * matrices into the Fiat-Shamir transcript once (src/lib.rs:62-64) and keeps the hash state.
 * Both sparse-product plans are built on the device from these arrays (the caller's buffers are only read during the call).
 * The columns of a row should be distinct, as upstream `to_matrices` produces them; they need not be sorted.  (With a
 * repeated (row, column) the reference's own sum_over_y adds both terms while its eval_on_x keeps one; this library adds
 * them in both.) */
sb_status sb_index_create(sb_ctx* ctx, uint32_t log_n, const sb_csr* a, const sb_csr* b, const sb_csr* c, sb_index** out);
void sb_index_destroy(sb_index* idx);
/* wall time of sb_index_create, in two parts: validation + device-side plan construction (upload included), and what the
 * transcript hash of the matrices (a serial BLAKE2s chain on a host thread, running beside the former) added after it */
void sb_index_timing(const sb_index* idx, double* plan_ms, double* hash_wait_ms);

/* ---- public parameters ------------------------------------------------------------------------ */
/* Load a reference PublicParameter (src/commitment/data_structures.rs:10-17):
 * powers_of_g0 = powers_of_g[0] (2^nv G1 points; the only G1 level the prover reads, commit.rs:25),
 * powers_of_h[i] = 2^(nv-i) G2 points for i < nv (open.rs:49), h (open.rs:54). */
sb_status sb_pp_load(sb_ctx* ctx, uint32_t nv, const void* powers_of_g0, const void* const* powers_of_h, const void* h, sb_pp** out);
/* MLPolyCommit::keygen, src/commitment/setup.rs:27-105, with caller-supplied generators and trapdoor
 * (the reference samples g, h, t from its rng).  g: G1 affine, h: G2 affine, t: nv Fr.
 * keep_all_levels != 0 also keeps every powers_of_g level and powers_of_h[0] for export. */
sb_status sb_pp_keygen(sb_ctx* ctx, uint32_t nv, const void* g, const void* h, const void* t, int keep_all_levels, sb_pp** out);
/* Copy a level back to the host (tests / serialization): group 1 or 2; level < nv; out = 2^(nv-level) points.
 * Requires keep_all_levels for G1 levels > 0 and for G2 level 0. */
sb_status sb_pp_export(sb_ctx* ctx, const sb_pp* pp, int group, uint32_t level, void* out);
/* VerifierParameter.g_mask_random = g^{t_i}, src/commitment/setup.rs:88-101 (only after sb_pp_keygen). */
sb_status sb_pp_export_g_mask(sb_ctx* ctx, const sb_pp* pp, void* out_nv_g1);
void sb_pp_destroy(sb_pp* pp);

/* ---- MLPolyCommit ----------------------------------------------------------------------------- */
/* commit: src/commitment/commit.rs:17-29.  z: 2^nv Fr (host).  out: G1 affine. */
sb_status sb_commit(sb_ctx* ctx, const sb_pp* pp, const void* z, void* out_g1);
/* open: src/commitment/open.rs:19-58.  point: nv Fr.  out_eval: Fr.  out_proofs: nv G2 affine. */
sb_status sb_open(sb_ctx* ctx, const sb_pp* pp, const void* z, const void* point, void* out_eval, void* out_proofs_g2);
/* VariableBaseMSM::multi_scalar_mul over caller bases (host arrays), n >= 1.  group: 1 or 2. */
sb_status sb_msm(sb_ctx* ctx, int group, const void* bases, const void* scalars, size_t n, void* out_affine);

/* ---- data-structure level entry points -------------------------------------------------------- */
/* eq_extension, src/data_structures/eq.rs:5-20, as the single product table
 * out[x] = prod_i eq_i(x) (2^dim Fr); see DESIGN.md D1 for why the dim separate tables are never built. */
sb_status sb_eq_table(sb_ctx* ctx, const void* t, uint32_t dim, void* out);
/* MatrixExtension::sum_over_y x3, src/data_structures/r1cs_reader.rs:75-85: out = Az, Bz, Cz (n Fr each) */
sb_status sb_sum_over_y(sb_ctx* ctx, const sb_index* idx, const void* z, void* az, void* bz, void* cz);
/* MatrixExtension::eval_on_x(r_x).multiply(r_k) summed over k, r1cs_reader.rs:91-117 + prover.rs:239-241:
 * out[y] = sum_k r_abc[k] * M_k(r_x, y).  Pass r_abc = NULL for a single matrix `which` in {0,1,2} unscaled. */
sb_status sb_eval_on_x(sb_ctx* ctx, const sb_index* idx, const void* r_x, const void* r_abc, int which, void* out);

/* ---- AHP prover rounds: MLProofForR1CS::*, src/ahp/prover.rs:109-281 ---------------------------- */
/* prover_init (prover.rs:109-121): |v| power of two, |v| + |w| == n, else SB_EINVAL. */
sb_status sb_prover_init(sb_ctx* ctx, const sb_index* idx, const void* v, size_t nv_len, const void* w, size_t nw_len, sb_prover** out);
void sb_prover_destroy(sb_prover* p);
/* prover_first_round (prover.rs:123-141): commitment to z = v || w. */
sb_status sb_prover_first_round(sb_prover* p, const sb_pp* pp, void* out_commit_g1);
/* prover_second_round (prover.rs:143-160): r_v has log2|v| Fr; opens z at (r_v, 0..0). */
sb_status sb_prover_second_round(sb_prover* p, const sb_pp* pp, const void* r_v, void* out_z_rv_0, void* out_proofs_g2);
/* prover_third_round (prover.rs:163-196): tor has log_n Fr.  Builds eq(tor,.), Az, Bz, Cz. */
sb_status sb_prover_third_round(sb_prover* p, const void* tor);
/* prove_first_sumcheck_round (prover.rs:199-207): v_msg = NULL in the first round, else the previous
 * challenge.  out_evals: log_n + 3 Fr (the ProverMsg.evaluations of the reference). */
sb_status sb_prover_first_sumcheck_round(sb_prover* p, const void* v_msg, void* out_evals);
/* prove_fourth_round (prover.rs:210-228): out_vabc = va, vb, vc. */
sb_status sb_prover_fourth_round(sb_prover* p, const void* last_random_point, void* out_vabc);
/* prove_fifth_round (prover.rs:230-255): r_abc = r_a, r_b, r_c. */
sb_status sb_prover_fifth_round(sb_prover* p, const void* r_abc);
/* prove_second_sumcheck_round (prover.rs:258-266): out_evals: 3 Fr. */
sb_status sb_prover_second_sumcheck_round(sb_prover* p, const void* v_msg, void* out_evals);
/* prove_sixth_round (prover.rs:268-281): out_z_ry Fr, out_proofs nv G2. */
sb_status sb_prover_sixth_round(sb_prover* p, const sb_pp* pp, const void* last_random_point, void* out_z_ry, void* out_proofs_g2);
/* Intermediates for parity tests: copies Az, Bz, Cz (after third round) to the host; any may be NULL. */
sb_status sb_prover_export_abc(sb_prover* p, void* az, void* bz, void* cz);

/* ---- the non-interactive argument: MLArgumentForR1CS::prove, src/lib.rs:58-146 ------------------ */
/* Optional trace of everything the parity tests compare; every pointer may be NULL. */
typedef struct {
    void* az; void* bz; void* cz;          /* n Fr each */
    void* sc1_evals;                       /* log_n * (log_n + 3) Fr */
    void* sc2_evals;                       /* log_n * 3 Fr */
    void* r_v; void* tor; void* r_x; void* r_abc; void* r_y;   /* challenges */
    void* vabc;                            /* 3 Fr */
    void* commitment;                      /* G1 affine */
    void* z_rv_0; void* z_ry;              /* Fr */
    void* open1_proofs; void* open2_proofs; /* log_n G2 affine each */
    double phase_ms[16];                   /* device+host time per phase, see sb_phase_name */
} sb_trace;
const char* sb_phase_name(int i);          /* NULL past the last phase */
/* the reference's own timer span for phase i ("Prove 1" ... "Prove Sumcheck 2", src/lib.rs:71-135); sb_prove wraps each
 * phase in an NVTX range of that name */
const char* sb_phase_span(int i);
/* Serialized Proof (src/data_structures/proof.rs:11-20, CanonicalSerialize, compressed points).
 * *len: in = capacity of `proof`, out = bytes written (or needed, with SB_EINVAL, if too small). */
sb_status sb_prove(sb_ctx* ctx, const sb_index* idx, const sb_pp* pp, const void* v, size_t nv_len, const void* w, size_t nw_len,
                   uint8_t* proof, size_t* len, sb_trace* trace);
size_t sb_proof_size(uint32_t log_n);
/* Same, with the witness already resident in HBM (bench.py's device-resident arm). */
sb_status sb_witness_upload(sb_ctx* ctx, const sb_index* idx, const void* v, size_t nv_len, const void* w, size_t nw_len, sb_witness** out);
void sb_witness_destroy(sb_witness* w);
sb_status sb_prove_resident(sb_ctx* ctx, const sb_index* idx, const sb_pp* pp, const sb_witness* w, uint8_t* proof, size_t* len, sb_trace* trace);

/* ---- self-test / measurement hooks ------------------------------------------------------------- */
/* bytes this library copied host->device / device->host so far in this process */
void sb_copy_counters(uint64_t* h2d_bytes, uint64_t* d2h_bytes);
/* per-kernel CUDA-event timing on the launching stream: enable, run, then read a JSON object
 * {"kernel": {"launches": n, "ms": t}, ...}; the report call synchronizes and clears the records.
 * Returns the size needed for the full report (including the terminating NUL). */
void sb_prof_enable(int on);
/* on != 0: run the MSMs of an opening one after another on one stream (so that the per-kernel event
 * times of sb_prof_report do not include time spent queued behind other streams); default off. */
void sb_set_serial_msm(sb_ctx* ctx, int on);
size_t sb_prof_report(char* buf, size_t cap);
/* same records as a timeline: JSON array [[kernel, start_ms, end_ms], ...] relative to the first launch */
size_t sb_prof_timeline(char* buf, size_t cap);
/* out = a op b elementwise on the device.  field: 0 = Fr, 1 = Fq.  op: 0 add, 1 sub, 2 mul, 3 mul (portable path) */
/* host field arithmetic against itself, no device needed: fast 64-bit product vs the portable loop, binary-GCD inversion vs
 * the Fermat ladder, n operands per field; returns the number of mismatches */
int sb_selftest_host_field(int n, uint64_t seed);
sb_status sb_field_binop(sb_ctx* ctx, int field, int op, const void* a, const void* b, void* out, size_t n);
/* integer-pipe microbenchmark: n_threads threads x 2 chains x iters Montgomery products; returns
 * milliseconds of the kernel (CUDA events).  field: 0 = Fr (8 limbs), 1 = Fq (12 limbs). */
sb_status sb_mul_bench(sb_ctx* ctx, int field, size_t n_threads, int iters, double* out_ms);
/* time one kernel of the hot path alone with CUDA events over `reps` launches (bench.py roofline):
 * which: 0 = sumcheck-1 fused fold+evaluate round on 2^log_m-entry tables, 1 = sumcheck-1 first round,
 * 2 = sumcheck-2 fused round, 3 = opening fold.  flush_l2 != 0 writes a 256 MiB buffer between launches. */
sb_status sb_kernel_bench(sb_ctx* ctx, int which, uint32_t log_m, int reps, int flush_l2, double* out_ms_avg);

#ifdef __cplusplus
}
#endif
#endif /* SPARTAN_B200_H */
