//! Rust face of `libsvb200.so` (C ABI: `include/stark_verifier_b200.h`) for the reference crate
//! `semaphore_aggregation` (DoHoonKim8/stark-verifier).
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no cargo/rustc there, and the crate's git
//! dependencies plonky2 / halo2 are not on disk).  It is the binding a maintainer of the reference
//! adds as `src/plonky2_verifier/gpu.rs`; it uses only the reference's own data model
//! (`ProofValues`, `FriProofValues`, `FriParams`, `CommonData`, `VerificationKeyValues`:
//! `src/plonky2_verifier/types/*.rs`).  The Python ctypes binding (`stark-verifier_b200/api.py`)
//! is the same seam, exercised by the test-suite.
//!
//! Seam replaced: `FriVerifierChip::construct(..)` + `verify_fri_proof(initial_merkle_caps,
//! fri_challenges, fri_openings, fri_proof, fri_instance_info)` (`chip/fri_chip.rs:35-46,329-362`),
//! called from `PlonkVerifierChip::verify_proof_with_challenges`
//! (`chip/plonk/plonk_verifier_chip.rs:212-240`).  The Rust side keeps deserialisation and the
//! Fiat-Shamir transcript (`get_challenges`, `plonk_verifier_chip.rs:55-154`); the GPU does the
//! FRI query phase of a whole batch of proofs and returns one accept bit per proof.
#![allow(non_camel_case_types)]

use std::ffi::{c_char, c_int, c_void, CStr};

use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::types::Field;

use crate::plonky2_verifier::types::{
    common_data::{CommonData, FriParams},
    proof::{FriProofValues, OpeningSetValues, ProofValues},
    verification_key::VerificationKeyValues,
    MerkleCapValues,
};

pub const SV_MEM_HOST: c_int = 0;
pub const SV_MEM_DEVICE: c_int = 1;
pub const SV_HASH_POSEIDON_GOLDILOCKS: u32 = 0;
pub const SV_MAX_STEPS: usize = 32;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct sv_fri_shape {
    pub degree_bits: u32,
    pub rate_bits: u32,
    pub cap_height: u32,
    pub num_query_rounds: u32,
    pub proof_of_work_bits: u32,
    pub num_steps: u32,
    pub final_poly_len: u32,
    pub hiding: u32,
    pub oracle_num_polys: [u32; 4],
    pub oracle_blinding: [u32; 4],
    pub num_zs: u32,
    pub hash_kind: u32,
    pub reduction_arity_bits: [u32; SV_MAX_STEPS],
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct sv_fri_layout {
    pub ncap: u32, pub lde_bits: u32, pub n0: u32, pub n1: u32,
    pub off_init_caps: u32, pub off_step_caps: u32, pub off_open0: u32, pub off_open1: u32,
    pub off_final_poly: u32, pub off_pow_witness: u32,
    pub off_alpha: u32, pub off_betas: u32, pub off_pow_response: u32, pub off_indices: u32,
    pub off_zeta: u32, pub off_zeta_next: u32,
    pub header_words: u32,
    pub leaf_len: [u32; 4],
    pub q_off_init_evals: [u32; 4], pub q_off_init_sibs: [u32; 4], pub init_depth: u32,
    pub q_off_step_evals: [u32; SV_MAX_STEPS], pub q_off_step_sibs: [u32; SV_MAX_STEPS],
    pub step_depth: [u32; SV_MAX_STEPS],
    pub step_arity_bits: [u32; SV_MAX_STEPS], pub step_index_shift: [u32; SV_MAX_STEPS],
    pub query_words: u32, pub record_words: u32,
    pub algo_bytes_per_query: u32, pub algo_bytes_shared: u32, pub perms_per_query: u32,
}

/// `sv_plonk_common`: the CommonData / CircuitConfig fields that fix the vector lengths of the wire format.
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct sv_plonk_common {
    pub num_constants: u32,
    pub num_routed_wires: u32,
    pub num_wires: u32,
    pub num_challenges: u32,
    pub num_partial_products: u32,
    pub quotient_degree_factor: u32,
    pub num_public_inputs: u32,
}

pub const SV_FAIL_MALFORMED: u32 = 8;
pub const SV_FAIL_PLONK: u32 = 9;
pub const SV_MAX_GATES: usize = 32;
pub const SV_MAX_SELECTORS: usize = 8;
pub const SV_MAX_ROUTED_WIRES: usize = 128;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct sv_plonk_gate { pub kind: u32, pub param: u32, pub param2: u32, pub param3: u32, pub selector_index: u32 }

/// `sv_common_circuit_data`: plonky2's CommonCircuitData as plain arrays (what CommonData::from reads)
#[repr(C)]
pub struct sv_common_circuit_data {
    pub common: sv_plonk_common,
    pub rate_bits: u32, pub cap_height: u32, pub proof_of_work_bits: u32, pub num_query_rounds: u32,
    pub hiding: u32, pub degree_bits: u32,
    pub num_reduction_steps: u32, pub reduction_arity_bits: *const u32,
    pub num_gate_constraints: u32,
    pub num_gates: u32, pub gate_ids: *const *const c_char,
    pub selector_indices: *const u32,
    pub num_selector_groups: u32, pub group_starts: *const u32, pub group_ends: *const u32,
    pub num_k_is: u32, pub k_is: *const u64,
    pub hash_kind: u32,
}

/// What `eval_vanishing_poly` reads from `CommonData` (types/common_data.rs:56-96).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct sv_plonk_circuit {
    pub common: sv_plonk_common,
    pub degree_bits: u32,
    pub num_gate_constraints: u32,
    pub num_selectors: u32,
    pub group_lo: [u32; SV_MAX_SELECTORS],
    pub group_hi: [u32; SV_MAX_SELECTORS],
    pub num_gates: u32,
    pub gates: [sv_plonk_gate; SV_MAX_GATES],
    pub k_is: [u64; SV_MAX_ROUTED_WIRES],
}

#[repr(C)]
pub struct sv_ctx { _private: [u8; 0] }

#[link(name = "svb200")]
extern "C" {
    pub fn sv_ctx_create(device: c_int, out: *mut *mut sv_ctx) -> c_int;
    pub fn sv_ctx_destroy(ctx: *mut sv_ctx);
    pub fn sv_last_error(ctx: *const sv_ctx) -> *const c_char;
    pub fn sv_ctx_set_stream(ctx: *mut sv_ctx, cuda_stream: *mut c_void) -> c_int;
    pub fn sv_ctx_synchronize(ctx: *mut sv_ctx) -> c_int;
    pub fn sv_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn sv_host_free(p: *mut c_void) -> c_int;
    pub fn sv_fri_layout_make(shape: *const sv_fri_shape, out: *mut sv_fri_layout) -> c_int;
    pub fn sv_poseidon_permute_batch(ctx: *mut sv_ctx, input: *const u64, out: *mut u64, n: usize,
                                     hash_kind: c_int, mem: c_int) -> c_int;
    pub fn sv_poseidon_permute_batch_coop(ctx: *mut sv_ctx, input: *const u64, out: *mut u64, n: usize, mem: c_int) -> c_int;
    pub fn sv_goldilocks_mul_add_batch(ctx: *mut sv_ctx, a: *const u64, b: *const u64, c: *const u64, out: *mut u64,
                                       n: usize, mem: c_int) -> c_int;
    pub fn sv_merkle_verify_batch(ctx: *mut sv_ctx, leaf_len: u32, depth: u32, cap_height: u32, hash_kind: c_int,
                                  paths: *const u64, indices: *const u64, caps: *const u64, ok: *mut u8,
                                  n: usize, mem: c_int) -> c_int;
    pub fn sv_merkle_tree_build(ctx: *mut sv_ctx, hash_kind: c_int, leaf_len: u32, leaves: *const u64, n_leaves: usize,
                                cap_height: u32, layers_out: *mut u64, mem: c_int) -> c_int;
    pub fn sv_fri_verify_batch(ctx: *mut sv_ctx, shape: *const sv_fri_shape, n_proofs: usize, records: *const u64,
                               accept_bitmap: *mut u32, first_fail: *mut u32, mem: c_int) -> c_int;
    pub fn sv_fri_challenges_batch(ctx: *mut sv_ctx, shape: *const sv_fri_shape, n_proofs: usize, records: *mut u64,
                                   circuit_digest: *const u64, public_inputs_hashes: *const u64, num_challenges: u32,
                                   mem: c_int) -> c_int;
    pub fn sv_fri_verify_batch_fs(ctx: *mut sv_ctx, shape: *const sv_fri_shape, n_proofs: usize, records: *mut u64,
                                  circuit_digest: *const u64, public_inputs_hashes: *const u64, num_challenges: u32,
                                  accept_bitmap: *mut u32, first_fail: *mut u32, mem: c_int) -> c_int;
    pub fn sv_fri_challenges(shape: *const sv_fri_shape, record: *mut u64, circuit_digest: *const u64,
                             public_inputs_hash: *const u64, num_challenges: u32) -> c_int;
    // wire format: ProofWithPublicInputs::to_bytes() in, accept bits out
    pub fn sv_wire_proof_bytes(shape: *const sv_fri_shape, common: *const sv_plonk_common) -> usize;
    pub fn sv_wire_unpack_batch(shape: *const sv_fri_shape, common: *const sv_plonk_common, constants_sigmas_cap: *const u64,
                                blob: *const u8, stride_bytes: usize, n_proofs: usize, records_out: *mut u64,
                                pi_hashes_out: *mut u64, public_inputs_out: *mut u64, malformed_out: *mut u8,
                                nthreads: c_int) -> c_int;
    pub fn sv_verify_proofs_wire(ctx: *mut sv_ctx, shape: *const sv_fri_shape, common: *const sv_plonk_common,
                                 constants_sigmas_cap: *const u64, circuit_digest: *const u64, blob: *const u8,
                                 stride_bytes: usize, n_proofs: usize, accept_bitmap: *mut u32, first_fail: *mut u32) -> c_int;
    pub fn sv_public_inputs_hash(public_inputs: *const u64, n: usize, out: *mut u64) -> c_int;
    // plonk-level checks and the complete verifier
    pub fn sv_plonk_gate_from_id(gate_id: *const c_char, out: *mut sv_plonk_gate) -> c_int;
    /// CommonData::from + CustomGateRef::from in one call (types/common_data.rs:224-270, gates/mod.rs:138-196)
    pub fn sv_circuit_from_common_data(cd: *const sv_common_circuit_data, shape_out: *mut sv_fri_shape, circuit_out: *mut sv_plonk_circuit) -> c_int;
    pub fn sv_plonk_circuit_check(circuit: *const sv_plonk_circuit) -> c_int;
    pub fn sv_verify_proofs_full(ctx: *mut sv_ctx, shape: *const sv_fri_shape, circuit: *const sv_plonk_circuit,
                                 constants_sigmas_cap: *const u64, circuit_digest: *const u64, blob: *const u8,
                                 stride_bytes: usize, n_proofs: usize, accept_bitmap: *mut u32, first_fail: *mut u32) -> c_int;
}

#[derive(Debug)]
pub struct GpuError(pub c_int, pub String);

/// `FriParams` + `CommonData::fri_oracles()` + `zs_range()` -> `sv_fri_shape`
/// (types/common_data.rs:43-54,148-150,202-221).
pub fn shape_from<F: halo2_proofs::halo2curves::ff::PrimeField>(cd: &CommonData<F>) -> sv_fri_shape {
    let p: &FriParams = &cd.fri_params;
    // any 2^k-ary reduction, k = 1..4 (the in-circuit verifier stops at arity 2, fri_chip.rs:211; the GPU fold does not)
    assert!(p.reduction_arity_bits.len() <= SV_MAX_STEPS && p.reduction_arity_bits.iter().all(|&a| (1..=4).contains(&a)));
    let oracles = cd.fri_oracles();
    let mut s = sv_fri_shape::default();
    s.degree_bits = p.degree_bits as u32;
    s.rate_bits = p.config.rate_bits as u32;
    s.cap_height = p.config.cap_height as u32;
    s.num_query_rounds = p.config.num_query_rounds as u32;
    s.proof_of_work_bits = p.config.proof_of_work_bits;
    s.num_steps = p.reduction_arity_bits.len() as u32;
    for (i, &a) in p.reduction_arity_bits.iter().enumerate() { s.reduction_arity_bits[i] = a as u32; }
    s.final_poly_len = 1u32 << (p.degree_bits - p.reduction_arity_bits.iter().sum::<usize>());
    s.hiding = p.hiding as u32;
    for k in 0..4 {
        s.oracle_num_polys[k] = oracles[k].num_polys as u32;
        s.oracle_blinding[k] = oracles[k].blinding as u32;
    }
    s.num_zs = cd.config.num_challenges as u32;
    s.hash_kind = SV_HASH_POSEIDON_GOLDILOCKS;
    s
}

fn put_cap<F: halo2_proofs::halo2curves::ff::PrimeField>(dst: &mut [u64], cap: &MerkleCapValues<F>) {
    for (i, h) in cap.0.iter().enumerate() {
        for j in 0..4 { dst[4 * i + j] = h.elements[j].0; }
    }
}

/// Flatten one proof into the record layout of `include/stark_verifier_b200.h`.  The challenge fields
/// (zeta, zeta_next, alpha, betas, pow_response, indices) are filled by `sv_fri_challenges` -- or by
/// the caller's own `get_challenges` -- afterwards.
pub fn flatten<F: halo2_proofs::halo2curves::ff::PrimeField>(
    l: &sv_fri_layout, s: &sv_fri_shape, proof: &ProofValues<F, 2>, vk: &VerificationKeyValues<F>, rec: &mut [u64],
) {
    assert_eq!(rec.len(), l.record_words as usize);
    let ncap4 = 4 * l.ncap as usize;
    // initial caps in FriInstanceInfo oracle order (plonk_verifier_chip.rs:212-217)
    let caps = [&vk.constants_sigmas_cap, &proof.wires_cap, &proof.plonk_zs_partial_products_cap, &proof.quotient_polys_cap];
    for (k, c) in caps.iter().enumerate() {
        let o = l.off_init_caps as usize + k * ncap4;
        put_cap(&mut rec[o..o + ncap4], c);
    }
    let fri: &FriProofValues<F, 2> = &proof.opening_proof;
    for (i, c) in fri.commit_phase_merkle_cap_values.iter().enumerate() {
        let o = l.off_step_caps as usize + i * ncap4;
        put_cap(&mut rec[o..o + ncap4], c);
    }
    // FriOpenings.batches[0] order: constants, sigmas, wires, zs, partial_products, quotient
    // (types/assigned.rs:26-37); batches[1] = plonk_zs_next (:38-40)
    let op: &OpeningSetValues<F, 2> = &proof.openings;
    let mut o = l.off_open0 as usize;
    for v in op.constants.iter().chain(&op.plonk_sigmas).chain(&op.wires).chain(&op.plonk_zs)
        .chain(&op.partial_products).chain(&op.quotient_polys) {
        rec[o] = v.elements[0].0; rec[o + 1] = v.elements[1].0; o += 2;
    }
    let mut o = l.off_open1 as usize;
    for v in op.plonk_zs_next.iter() { rec[o] = v.elements[0].0; rec[o + 1] = v.elements[1].0; o += 2; }
    let mut o = l.off_final_poly as usize;
    for v in fri.final_poly.0.iter() { rec[o] = v.elements[0].0; rec[o + 1] = v.elements[1].0; o += 2; }
    rec[l.off_pow_witness as usize] = fri.pow_witness.0;
    for (q, round) in fri.query_round_proofs.iter().enumerate() {
        let qb = l.header_words as usize + q * l.query_words as usize;
        for (k, (evals, path)) in round.initial_trees_proof.evals_proofs.iter().enumerate() {
            let e = qb + l.q_off_init_evals[k] as usize;
            for (j, v) in evals.iter().enumerate() { rec[e + j] = v.0; }
            let sb = qb + l.q_off_init_sibs[k] as usize;
            for (lv, h) in path.siblings.iter().enumerate() { for j in 0..4 { rec[sb + 4 * lv + j] = h.elements[j].0; } }
        }
        for (i, step) in round.steps.iter().enumerate() {
            let e = qb + l.q_off_step_evals[i] as usize;
            for (j, v) in step.evals.iter().enumerate() { rec[e + 2 * j] = v.elements[0].0; rec[e + 2 * j + 1] = v.elements[1].0; }
            let sb = qb + l.q_off_step_sibs[i] as usize;
            for (lv, h) in step.merkle_proof.siblings.iter().enumerate() { for j in 0..4 { rec[sb + 4 * lv + j] = h.elements[j].0; } }
        }
    }
    let _ = s;
}

/// One GPU context per (host thread, GPU).
pub struct GpuFriVerifier { ctx: *mut sv_ctx }
unsafe impl Send for GpuFriVerifier {}

impl GpuFriVerifier {
    pub fn new(device: i32) -> Result<Self, GpuError> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sv_ctx_create(device, &mut ctx) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sv_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(GpuError(rc, msg));
        }
        Ok(Self { ctx })
    }

    /// Batch replacement of `FriVerifierChip::verify_fri_proof`: one accept flag per proof
    /// (the reference panics / fails MockProver on an invalid proof; here invalidity is data).
    /// All proofs share one circuit shape (`common_data`), like the proofs of one aggregation layer.
    pub fn verify_batch<F: halo2_proofs::halo2curves::ff::PrimeField>(
        &mut self,
        proofs: &[(ProofValues<F, 2>, [GoldilocksField; 4] /* public_inputs_hash */)],
        vk: &VerificationKeyValues<F>,
        common_data: &CommonData<F>,
    ) -> Result<Vec<bool>, GpuError> {
        let shape = shape_from(common_data);
        let mut l = std::mem::MaybeUninit::<sv_fri_layout>::uninit();
        let rc = unsafe { sv_fri_layout_make(&shape, l.as_mut_ptr()) };
        if rc != 0 { return Err(GpuError(rc, "bad FRI shape".into())); }
        let l = unsafe { l.assume_init() };
        let rw = l.record_words as usize;
        let mut records = vec![0u64; rw * proofs.len()];
        let digest: Vec<u64> = vk.circuit_digest.elements.iter().map(|e| e.0).collect();
        for (i, (p, pi_hash)) in proofs.iter().enumerate() {
            let rec = &mut records[i * rw..(i + 1) * rw];
            flatten(&l, &shape, p, vk, rec);
            let ph: Vec<u64> = pi_hash.iter().map(|e| e.0).collect();
            let rc = unsafe { sv_fri_challenges(&shape, rec.as_mut_ptr(), digest.as_ptr(), ph.as_ptr(),
                                                common_data.config.num_challenges as u32) };
            if rc != 0 { return Err(GpuError(rc, "sv_fri_challenges".into())); }
        }
        let mut bitmap = vec![0u32; (proofs.len() + 31) / 32];
        let rc = unsafe { sv_fri_verify_batch(self.ctx, &shape, proofs.len(), records.as_ptr(), bitmap.as_mut_ptr(),
                                              std::ptr::null_mut(), SV_MEM_HOST) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sv_last_error(self.ctx)) }.to_string_lossy().into_owned();
            return Err(GpuError(rc, msg));
        }
        Ok((0..proofs.len()).map(|i| (bitmap[i >> 5] >> (i & 31)) & 1 == 1).collect())
    }
}

/// `CommonData` -> `sv_plonk_common` (types/common_data.rs:23-40,68-96).
pub fn common_from<F: halo2_proofs::halo2curves::ff::PrimeField>(cd: &CommonData<F>) -> sv_plonk_common {
    sv_plonk_common {
        num_constants: cd.num_constants as u32,
        num_routed_wires: cd.config.num_routed_wires as u32,
        num_wires: cd.config.num_wires as u32,
        num_challenges: cd.config.num_challenges as u32,
        num_partial_products: cd.num_partial_products as u32,
        quotient_degree_factor: cd.quotient_degree_factor as u32,
        num_public_inputs: cd.num_public_inputs as u32,
    }
}

/// `CommonData` -> `sv_plonk_circuit`.  `gate_ids[i]` = `common_circuit_data.gates[i].0.id()` of the plonky2 value the
/// `CommonData` was converted from (the reference's own `CustomGateRef::from` matches on the same strings,
/// chip/plonk/gates/mod.rs:138-196); an id it does not know is an error here, `unimplemented!()` there.
pub fn circuit_from<F: halo2_proofs::halo2curves::ff::PrimeField>(cd: &CommonData<F>, gate_ids: &[String]) -> Result<sv_plonk_circuit, GpuError> {
    let mut c: sv_plonk_circuit = unsafe { std::mem::zeroed() };
    c.common = common_from(cd);
    c.degree_bits = cd.fri_params.degree_bits as u32;
    c.num_gate_constraints = cd.num_gate_constraints as u32;
    c.num_selectors = cd.selectors_info.groups.len() as u32;
    for (s, g) in cd.selectors_info.groups.iter().enumerate() { c.group_lo[s] = g.start as u32; c.group_hi[s] = g.end as u32; }
    c.num_gates = gate_ids.len() as u32;
    for (i, id) in gate_ids.iter().enumerate() {
        let cid = std::ffi::CString::new(id.as_str()).unwrap();
        if unsafe { sv_plonk_gate_from_id(cid.as_ptr(), &mut c.gates[i]) } != 0 { return Err(GpuError(-2, format!("unknown gate {id}"))); }
        c.gates[i].selector_index = cd.selectors_info.selector_indices[i] as u32;
    }
    for (j, k) in cd.k_is.iter().enumerate() { c.k_is[j] = k.0; }
    let rc = unsafe { sv_plonk_circuit_check(&c) };
    if rc != 0 { return Err(GpuError(rc, "inconsistent circuit description".into())); }
    Ok(c)
}

impl GpuFriVerifier {
    /// The complete verifier (plonk identity AND FRI) on serialised proofs of one circuit: what
    /// `verify_inside_snark_mock` establishes per proof (verifier_api.rs:34-56), natively and per batch.
    pub fn verify_serialized_full<F: halo2_proofs::halo2curves::ff::PrimeField>(
        &mut self, blob: &[u8], vk: &VerificationKeyValues<F>, common_data: &CommonData<F>, gate_ids: &[String],
    ) -> Result<Vec<bool>, GpuError> {
        let shape = shape_from(common_data);
        let circuit = circuit_from(common_data, gate_ids)?;
        let nb = unsafe { sv_wire_proof_bytes(&shape, &circuit.common) };
        if nb == 0 || blob.len() % nb != 0 { return Err(GpuError(-1, "blob is not a whole number of proofs".into())); }
        let n = blob.len() / nb;
        let mut cap = vec![0u64; 4 << shape.cap_height];
        put_cap(&mut cap, &vk.constants_sigmas_cap);
        let digest: Vec<u64> = vk.circuit_digest.elements.iter().map(|e| e.0).collect();
        let mut bitmap = vec![0u32; (n + 31) / 32];
        let rc = unsafe { sv_verify_proofs_full(self.ctx, &shape, &circuit, cap.as_ptr(), digest.as_ptr(), blob.as_ptr(), nb, n,
                                                bitmap.as_mut_ptr(), std::ptr::null_mut()) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sv_last_error(self.ctx)) }.to_string_lossy().into_owned();
            return Err(GpuError(rc, msg));
        }
        Ok((0..n).map(|i| (bitmap[i >> 5] >> (i & 31)) & 1 == 1).collect())
    }

    /// The whole verifier-side path from serialised proofs: `blob` holds `n` back-to-back
    /// `ProofWithPublicInputs::to_bytes()` strings of ONE circuit (they all have the same length,
    /// `sv_wire_proof_bytes`).  Unpacking, the public-inputs hash, the Fiat-Shamir transcript and the FRI
    /// query phase all run on the device; nothing is repacked on the host.
    pub fn verify_serialized<F: halo2_proofs::halo2curves::ff::PrimeField>(
        &mut self, blob: &[u8], vk: &VerificationKeyValues<F>, common_data: &CommonData<F>,
    ) -> Result<Vec<bool>, GpuError> {
        let (shape, common) = (shape_from(common_data), common_from(common_data));
        let nb = unsafe { sv_wire_proof_bytes(&shape, &common) };
        if nb == 0 || blob.len() % nb != 0 { return Err(GpuError(-1, "blob is not a whole number of proofs".into())); }
        let n = blob.len() / nb;
        let mut cap = vec![0u64; 4 << shape.cap_height];
        put_cap(&mut cap, &vk.constants_sigmas_cap);
        let digest: Vec<u64> = vk.circuit_digest.elements.iter().map(|e| e.0).collect();
        let mut bitmap = vec![0u32; (n + 31) / 32];
        let rc = unsafe { sv_verify_proofs_wire(self.ctx, &shape, &common, cap.as_ptr(), digest.as_ptr(), blob.as_ptr(), nb, n,
                                                bitmap.as_mut_ptr(), std::ptr::null_mut()) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(sv_last_error(self.ctx)) }.to_string_lossy().into_owned();
            return Err(GpuError(rc, msg));
        }
        Ok((0..n).map(|i| (bitmap[i >> 5] >> (i & 31)) & 1 == 1).collect())
    }
}

impl Drop for GpuFriVerifier {
    fn drop(&mut self) { unsafe { sv_ctx_destroy(self.ctx) } }
}

const _: fn() = || { let _ = GoldilocksField::ZERO; };
