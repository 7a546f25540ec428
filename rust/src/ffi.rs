//! ffi.rs -- the C ABI of libspartan_b200 (include/spartan_b200.h) as Rust declarations, one per exported symbol.
//!
//! Each block names the reference function whose body it replaces (file:line in tsunrise/r1cs-spartan).  Field
//! elements cross zero-copy: arkworks' `Fp256` / `Fp384` are 4 / 6 little-endian u64 Montgomery limbs, the same bytes
//! as the library's 8 / 12 u32 limbs.  Affine points are `(x, y)` with `(0, 0)` for infinity (G1 96 bytes, G2 192).
//! NOT compiled in the authoring environment (no cargo/rustc); kept in step with the header by
//! tests/test_wire_cpu.py::test_rust_shim_lists_every_export.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

pub const SB_OK: c_int = 0;
pub const SB_EINVAL: c_int = 1;     // Error::InvalidArgument (src/error.rs), same preconditions as the reference
pub const SB_ECUDA: c_int = 2;
pub const SB_ENOMEM: c_int = 3;
pub const SB_ECOMM: c_int = 4;
pub const SB_EINTERNAL: c_int = 5;

pub enum sb_ctx {}
pub enum sb_index {}
pub enum sb_pp {}
pub enum sb_prover {}
pub enum sb_witness {}

/// One constraint matrix in CSR form (a `Matrix<F> = Vec<Vec<(F, usize)>>` flattened: a Rust tuple has no defined layout).
#[repr(C)]
pub struct sb_csr {
    pub row_ptr: *const u64,
    pub col: *const u32,
    pub val: *const c_void,
}
/// Exchange hooks of a hypercube-sharded context (one process per GPU); see INTEGRATION.md "Multi-GPU".
#[repr(C)]
pub struct sb_comm {
    pub rank: c_int,
    pub world: c_int,
    pub allgather: Option<unsafe extern "C" fn(user: *mut c_void, send: *const c_void, recv: *mut c_void, bytes: usize) -> c_int>,
    pub barrier: Option<unsafe extern "C" fn(user: *mut c_void) -> c_int>,
    pub user: *mut c_void,
}
/// Optional trace of the non-interactive prover (every pointer may be null).
#[repr(C)]
pub struct sb_trace {
    pub az: *mut c_void, pub bz: *mut c_void, pub cz: *mut c_void,
    pub sc1_evals: *mut c_void, pub sc2_evals: *mut c_void,
    pub r_v: *mut c_void, pub tor: *mut c_void, pub r_x: *mut c_void, pub r_abc: *mut c_void, pub r_y: *mut c_void,
    pub vabc: *mut c_void, pub commitment: *mut c_void, pub z_rv_0: *mut c_void, pub z_ry: *mut c_void,
    pub open1_proofs: *mut c_void, pub open2_proofs: *mut c_void,
    pub phase_ms: [c_double; 16],
}

extern "C" {
    // ---- exchange layers shipped with the library
    pub fn sb_comm_shm_open(name: *const c_char, rank: c_int, world: c_int, create: c_int, out: *mut sb_comm) -> c_int;
    pub fn sb_comm_local_open(world: c_int, out: *mut sb_comm) -> c_int;
    pub fn sb_comm_shm_abort(comm: *mut sb_comm);
    pub fn sb_comm_shm_close(comm: *mut sb_comm);

    // ---- contexts.  sb_ctx_create_multi: ONE context over several GPUs of this process, so that
    // MLArgumentForR1CS::prove (src/lib.rs:58) stays one call from one process
    pub fn sb_ctx_create(device: c_int, out: *mut *mut sb_ctx) -> c_int;
    pub fn sb_ctx_create_sharded(device: c_int, comm: *const sb_comm, out: *mut *mut sb_ctx) -> c_int;
    pub fn sb_ctx_create_multi(devices: *const c_int, ndev: c_int, out: *mut *mut sb_ctx) -> c_int;
    pub fn sb_ctx_destroy(ctx: *mut sb_ctx);
    pub fn sb_last_error(ctx: *const sb_ctx) -> *const c_char;
    pub fn sb_launch_count() -> u64;
    pub fn sb_device_count() -> c_int;

    // ---- MLArgumentForR1CS::index (src/lib.rs:45-51) -> MLProofForR1CS::index (src/ahp/indexer.rs:41-64),
    // MatrixExtension::new (src/data_structures/r1cs_reader.rs:36-70)
    pub fn sb_index_create(ctx: *mut sb_ctx, log_n: u32, a: *const sb_csr, b: *const sb_csr, c: *const sb_csr, out: *mut *mut sb_index) -> c_int;
    pub fn sb_index_destroy(idx: *mut sb_index);
    pub fn sb_index_timing(idx: *const sb_index, plan_ms: *mut c_double, hash_wait_ms: *mut c_double);

    // ---- PublicParameter (src/commitment/data_structures.rs:10-17), MLPolyCommit::keygen (src/commitment/setup.rs:27-105)
    pub fn sb_pp_load(ctx: *mut sb_ctx, nv: u32, powers_of_g0: *const c_void, powers_of_h: *const *const c_void, h: *const c_void, out: *mut *mut sb_pp) -> c_int;
    pub fn sb_pp_keygen(ctx: *mut sb_ctx, nv: u32, g: *const c_void, h: *const c_void, t: *const c_void, keep_all_levels: c_int, out: *mut *mut sb_pp) -> c_int;
    pub fn sb_pp_export(ctx: *mut sb_ctx, pp: *const sb_pp, group: c_int, level: u32, out: *mut c_void) -> c_int;
    pub fn sb_pp_export_g_mask(ctx: *mut sb_ctx, pp: *const sb_pp, out_nv_g1: *mut c_void) -> c_int;
    pub fn sb_pp_destroy(pp: *mut sb_pp);

    // ---- MLPolyCommit::commit (src/commitment/commit.rs:17-29), MLPolyCommit::open (src/commitment/open.rs:19-58),
    // VariableBaseMSM::multi_scalar_mul as called from those
    pub fn sb_commit(ctx: *mut sb_ctx, pp: *const sb_pp, z: *const c_void, out_g1: *mut c_void) -> c_int;
    pub fn sb_open(ctx: *mut sb_ctx, pp: *const sb_pp, z: *const c_void, point: *const c_void, out_eval: *mut c_void, out_proofs_g2: *mut c_void) -> c_int;
    pub fn sb_msm(ctx: *mut sb_ctx, group: c_int, bases: *const c_void, scalars: *const c_void, n: usize, out_affine: *mut c_void) -> c_int;

    // ---- eq_extension (src/data_structures/eq.rs:5-20), MatrixExtension::sum_over_y / eval_on_x (r1cs_reader.rs:75-117)
    pub fn sb_eq_table(ctx: *mut sb_ctx, t: *const c_void, dim: u32, out: *mut c_void) -> c_int;
    pub fn sb_sum_over_y(ctx: *mut sb_ctx, idx: *const sb_index, z: *const c_void, az: *mut c_void, bz: *mut c_void, cz: *mut c_void) -> c_int;
    pub fn sb_eval_on_x(ctx: *mut sb_ctx, idx: *const sb_index, r_x: *const c_void, r_abc: *const c_void, which: c_int, out: *mut c_void) -> c_int;

    // ---- the nine round functions of MLProofForR1CS (src/ahp/prover.rs:109-281), same names, same argument meaning
    pub fn sb_prover_init(ctx: *mut sb_ctx, idx: *const sb_index, v: *const c_void, nv_len: usize, w: *const c_void, nw_len: usize, out: *mut *mut sb_prover) -> c_int;
    pub fn sb_prover_destroy(p: *mut sb_prover);
    pub fn sb_prover_first_round(p: *mut sb_prover, pp: *const sb_pp, out_commit_g1: *mut c_void) -> c_int;
    pub fn sb_prover_second_round(p: *mut sb_prover, pp: *const sb_pp, r_v: *const c_void, out_z_rv_0: *mut c_void, out_proofs_g2: *mut c_void) -> c_int;
    pub fn sb_prover_third_round(p: *mut sb_prover, tor: *const c_void) -> c_int;
    pub fn sb_prover_first_sumcheck_round(p: *mut sb_prover, v_msg: *const c_void, out_evals: *mut c_void) -> c_int;
    pub fn sb_prover_fourth_round(p: *mut sb_prover, last_random_point: *const c_void, out_vabc: *mut c_void) -> c_int;
    pub fn sb_prover_fifth_round(p: *mut sb_prover, r_abc: *const c_void) -> c_int;
    pub fn sb_prover_second_sumcheck_round(p: *mut sb_prover, v_msg: *const c_void, out_evals: *mut c_void) -> c_int;
    pub fn sb_prover_sixth_round(p: *mut sb_prover, pp: *const sb_pp, last_random_point: *const c_void, out_z_ry: *mut c_void, out_proofs_g2: *mut c_void) -> c_int;
    pub fn sb_prover_export_abc(p: *mut sb_prover, az: *mut c_void, bz: *mut c_void, cz: *mut c_void) -> c_int;

    // ---- MLArgumentForR1CS::prove (src/lib.rs:58-146); the bytes are Proof's CanonicalSerialize layout (data_structures/proof.rs:11-20)
    pub fn sb_phase_name(i: c_int) -> *const c_char;
    pub fn sb_phase_span(i: c_int) -> *const c_char;
    pub fn sb_prove(ctx: *mut sb_ctx, idx: *const sb_index, pp: *const sb_pp, v: *const c_void, nv_len: usize, w: *const c_void, nw_len: usize,
                    proof: *mut u8, len: *mut usize, trace: *mut sb_trace) -> c_int;
    pub fn sb_proof_size(log_n: u32) -> usize;
    pub fn sb_witness_upload(ctx: *mut sb_ctx, idx: *const sb_index, v: *const c_void, nv_len: usize, w: *const c_void, nw_len: usize, out: *mut *mut sb_witness) -> c_int;
    pub fn sb_witness_destroy(w: *mut sb_witness);
    pub fn sb_prove_resident(ctx: *mut sb_ctx, idx: *const sb_index, pp: *const sb_pp, w: *const sb_witness, proof: *mut u8, len: *mut usize, trace: *mut sb_trace) -> c_int;

    // ---- measurement hooks (bench.py / tests)
    pub fn sb_copy_counters(h2d_bytes: *mut u64, d2h_bytes: *mut u64);
    pub fn sb_prof_enable(on: c_int);
    pub fn sb_set_serial_msm(ctx: *mut sb_ctx, on: c_int);
    pub fn sb_prof_report(buf: *mut c_char, cap: usize) -> usize;
    pub fn sb_prof_timeline(buf: *mut c_char, cap: usize) -> usize;
    pub fn sb_selftest_host_field(n: c_int, seed: u64) -> c_int;
    pub fn sb_field_binop(ctx: *mut sb_ctx, field: c_int, op: c_int, a: *const c_void, b: *const c_void, out: *mut c_void, n: usize) -> c_int;
    pub fn sb_mul_bench(ctx: *mut sb_ctx, field: c_int, n_threads: usize, iters: c_int, out_ms: *mut c_double) -> c_int;
    pub fn sb_kernel_bench(ctx: *mut sb_ctx, which: c_int, log_m: u32, reps: c_int, flush_l2: c_int, out_ms_avg: *mut c_double) -> c_int;
}

/// `Error::InvalidArgument` for SB_EINVAL (same preconditions as the reference: prover.rs:114-119, indexer.rs:49,
/// r1cs_reader.rs:38-63); a new `Error::Device(String)` variant for SB_ECUDA / SB_ENOMEM / SB_ECOMM / SB_EINTERNAL.
pub fn check(ctx: *const sb_ctx, st: c_int) -> crate::error::SResult<()> {
    if st == SB_OK { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(sb_last_error(ctx)) }.to_string_lossy().into_owned();
    if st == SB_EINVAL { Err(crate::error::invalid_arg(&msg)) } else { Err(crate::Error::Device(msg)) }
}

/// One context for the whole process: every visible GPU when there are several (a power of two), else device 0.
/// `MLArgumentForR1CS::index / prove` and `MLPolyCommit::keygen / commit / open` take it from here.
pub fn process_context() -> *mut sb_ctx {
    use std::sync::Once;
    static INIT: Once = Once::new();
    static mut CTX: *mut sb_ctx = std::ptr::null_mut();
    unsafe {
        INIT.call_once(|| {
            let n = sb_device_count();
            let mut use_n = 1;
            while use_n * 2 <= n { use_n *= 2; }
            let devs: Vec<c_int> = (0..use_n).collect();
            let mut ctx: *mut sb_ctx = std::ptr::null_mut();
            let st = if use_n > 1 { sb_ctx_create_multi(devs.as_ptr(), use_n, &mut ctx) } else { sb_ctx_create(0, &mut ctx) };
            assert!(st == SB_OK, "libspartan_b200: {}", std::ffi::CStr::from_ptr(sb_last_error(std::ptr::null())).to_string_lossy());
            CTX = ctx;
        });
        CTX
    }
}
