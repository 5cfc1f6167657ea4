"""ctypes binding of libsvb200.so and a thin mirror of the reference's verifier interface.

Reference interface mirrored (paths relative to src/plonky2_verifier/):
  FriConfig, FriParams ............ types/common_data.rs:10-54
  FriVerifierChip::construct ....... chip/fri_chip.rs:35-46   (offset = MULTIPLICATIVE_GROUP_GENERATOR = 7)
  FriVerifierChip::verify_fri_proof  chip/fri_chip.rs:329-362 (here: a batch of flat proof records)
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

MEM_HOST = 0
MEM_DEVICE = 1
HASH_POSEIDON_GOLDILOCKS = 0
HASH_POSEIDON_BN254 = 1      # bn245_poseidon/plonky2_config.rs:54-75 (the reference's outermost proof)
SV_MAX_STEPS = 32
SV_MAX_ARITY_BITS = 4

FAIL_NAMES = {0: "ok", 1: "pow", 2: "noncanonical", 3: "init_merkle", 4: "zero_denominator",
              5: "step_eval", 6: "step_merkle", 7: "final_poly", 8: "malformed", 9: "plonk_identity"}
FAIL_MALFORMED = 8
FAIL_PLONK = 9


class SvError(RuntimeError):
    pass


class FriShape(ctypes.Structure):
    """sv_fri_shape (include/stark_verifier_b200.h)."""
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "degree_bits", "rate_bits", "cap_height", "num_query_rounds", "proof_of_work_bits",
        "num_steps", "final_poly_len", "hiding")] + [
        ("oracle_num_polys", ctypes.c_uint32 * 4), ("oracle_blinding", ctypes.c_uint32 * 4),
        ("num_zs", ctypes.c_uint32), ("hash_kind", ctypes.c_uint32), ("reduction_arity_bits", ctypes.c_uint32 * SV_MAX_STEPS)]


class Layout(ctypes.Structure):
    """sv_fri_layout (include/stark_verifier_b200.h)."""
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "ncap", "lde_bits", "n0", "n1", "off_init_caps", "off_step_caps", "off_open0", "off_open1",
        "off_final_poly", "off_pow_witness", "off_alpha", "off_betas", "off_pow_response",
        "off_indices", "off_zeta", "off_zeta_next", "header_words")] + [
        ("leaf_len", ctypes.c_uint32 * 4), ("q_off_init_evals", ctypes.c_uint32 * 4),
        ("q_off_init_sibs", ctypes.c_uint32 * 4), ("init_depth", ctypes.c_uint32),
        ("q_off_step_evals", ctypes.c_uint32 * SV_MAX_STEPS), ("q_off_step_sibs", ctypes.c_uint32 * SV_MAX_STEPS),
        ("step_depth", ctypes.c_uint32 * SV_MAX_STEPS), ("step_arity_bits", ctypes.c_uint32 * SV_MAX_STEPS),
        ("step_index_shift", ctypes.c_uint32 * SV_MAX_STEPS), ("query_words", ctypes.c_uint32),
        ("record_words", ctypes.c_uint32), ("algo_bytes_per_query", ctypes.c_uint32),
        ("algo_bytes_shared", ctypes.c_uint32), ("perms_per_query", ctypes.c_uint32)]


class PlonkCommon(ctypes.Structure):
    """sv_plonk_common: the CommonData / CircuitConfig fields the wire format takes its vector lengths from
    (types/common_data.rs:23-40, 68-96)."""
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "num_constants", "num_routed_wires", "num_wires", "num_challenges", "num_partial_products",
        "quotient_degree_factor", "num_public_inputs")]


GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC = 0, 1, 2, 3
GATE_ARITHMETIC_EXT, GATE_MUL_EXT, GATE_BASE_SUM, GATE_REDUCING, GATE_REDUCING_EXT = 4, 5, 6, 7, 8
GATE_RANDOM_ACCESS, GATE_POSEIDON_MDS, GATE_POSEIDON = 9, 10, 11
SV_MAX_GATES, SV_MAX_SELECTORS, SV_MAX_ROUTED_WIRES = 32, 8, 128


class PlonkGate(ctypes.Structure):
    """sv_plonk_gate"""
    _fields_ = [("kind", ctypes.c_uint32), ("param", ctypes.c_uint32), ("param2", ctypes.c_uint32), ("param3", ctypes.c_uint32),
                ("selector_index", ctypes.c_uint32)]


class PlonkCircuit(ctypes.Structure):
    """sv_plonk_circuit: what eval_vanishing_poly reads from CommonData (types/common_data.rs:68-96)."""
    _fields_ = [("common", PlonkCommon), ("degree_bits", ctypes.c_uint32), ("num_gate_constraints", ctypes.c_uint32),
                ("num_selectors", ctypes.c_uint32), ("group_lo", ctypes.c_uint32 * SV_MAX_SELECTORS),
                ("group_hi", ctypes.c_uint32 * SV_MAX_SELECTORS), ("num_gates", ctypes.c_uint32),
                ("gates", PlonkGate * SV_MAX_GATES), ("k_is", ctypes.c_uint64 * SV_MAX_ROUTED_WIRES)]


@dataclass
class FriConfig:
    """types/common_data.rs:10-21"""
    rate_bits: int
    cap_height: int
    proof_of_work_bits: int
    num_query_rounds: int


@dataclass
class FriParams:
    """types/common_data.rs:43-54 plus the FriInstanceInfo facts the FRI verifier reads
    (types/fri.rs:50-72, types/common_data.rs:153-221)."""
    config: FriConfig
    hiding: bool
    degree_bits: int
    reduction_arity_bits: List[int]
    oracle_num_polys: Sequence[int] = (84, 135, 20, 16)      # constants_sigmas, wires, zs_pp, quotient
    oracle_blinding: Sequence[bool] = (False, True, True, True)  # PlonkOracle::*.blinding (common_data.rs:101-123)
    num_zs: int = 2                                           # = num_challenges (zs_range, common_data.rs:148-150)
    hash_kind: int = HASH_POSEIDON_GOLDILOCKS                 # GenericConfig::Hasher of the proof

    def lde_bits(self) -> int:
        return self.degree_bits + self.config.rate_bits

    def final_poly_len(self) -> int:
        return 1 << (self.degree_bits - sum(self.reduction_arity_bits))

    def to_shape(self) -> FriShape:
        # arity 2^k folds, k = 1 .. 4: the reference's next_eval is arity-2 only (chip/fri_chip.rs:211); its demo's
        # ConstantArityBits(3, 5) proofs (plonky2_semaphore/access_set.rs:124) need the general fold (csrc/fri_fold.cuh)
        if len(self.reduction_arity_bits) > SV_MAX_STEPS or any(not 1 <= a <= SV_MAX_ARITY_BITS for a in self.reduction_arity_bits):
            raise SvError("reduction_arity_bits: at most %d entries, each in 1 .. %d" % (SV_MAX_STEPS, SV_MAX_ARITY_BITS))
        s = FriShape()
        s.degree_bits = self.degree_bits
        s.rate_bits = self.config.rate_bits
        s.cap_height = self.config.cap_height
        s.num_query_rounds = self.config.num_query_rounds
        s.proof_of_work_bits = self.config.proof_of_work_bits
        s.num_steps = len(self.reduction_arity_bits)
        s.final_poly_len = self.final_poly_len()
        s.hiding = int(self.hiding)
        s.oracle_num_polys = (ctypes.c_uint32 * 4)(*self.oracle_num_polys)
        s.oracle_blinding = (ctypes.c_uint32 * 4)(*[int(b) for b in self.oracle_blinding])
        s.num_zs = self.num_zs
        s.hash_kind = self.hash_kind
        for i, a in enumerate(self.reduction_arity_bits):
            s.reduction_arity_bits[i] = a
        return s


def constant_arity_bits(arity_bits: int, final_poly_bits: int, degree_bits: int, rate_bits: int, cap_height: int) -> List[int]:
    """plonky2 FriReductionStrategy::ConstantArityBits(arity_bits, final_poly_bits).reduction_arity_bits: fold by 2^arity_bits
    while the polynomial has more than 2^final_poly_bits coefficients and the next layer's tree would still be at least as tall
    as the cap (so every step tree has a cap of 2^cap_height entries)."""
    out, d = [], degree_bits
    while d > final_poly_bits and d + rate_bits - arity_bits >= cap_height:
        if d < arity_bits:
            raise SvError("ConstantArityBits: degree_bits < arity_bits")     # plonky2: assert!(degree_bits >= arity_bits)
        out.append(arity_bits)
        d -= arity_bits
    return out


def _params(degree_bits, rate_bits, cap_height, pow_bits, queries, hiding=False, final_bits=5, arity_bits=1, **kw) -> FriParams:
    # FriReductionStrategy::ConstantArityBits(1, 5) (bn245_poseidon/plonky2_config.rs:84); ConstantArityBits(3, 5) is the
    # semaphore demo's (plonky2_semaphore/access_set.rs:124)
    if arity_bits == 1:
        steps = [1] * max(0, degree_bits - final_bits)
    else:
        steps = constant_arity_bits(arity_bits, final_bits, degree_bits, rate_bits, cap_height)
    return FriParams(FriConfig(rate_bits, cap_height, pow_bits, queries), hiding, degree_bits, steps, **kw)


#: BASELINE configs[1]/[3]: 2^12 trace, blowup 8, 28 queries, cap 4, PoW 16 (standard_inner_stark_verifier_config)
SHAPE_A = _params(12, 3, 4, 16, 28)
#: BASELINE configs[2]: 2^20 trace, blowup 4, 84 queries
SHAPE_B = _params(20, 2, 4, 16, 84)
#: the reference's OUTER wrapped proof: Bn254PoseidonGoldilocksConfig + standard_stark_verifier_config
#: (bn245_poseidon/plonky2_config.rs:92-104: cap_height 0), hash family B
SHAPE_OUTER_BN254 = _params(12, 3, 0, 16, 28, hash_kind=HASH_POSEIDON_BN254)
#: BASELINE configs[0]: semaphore-shaped (zero_knowledge => salted leaves), plonky2_semaphore/access_set.rs:68-84
SHAPE_SEMAPHORE = _params(12, 3, 4, 16, 28, hiding=True)


@dataclass
class CommonData:
    """The part of the reference's CommonData (types/common_data.rs:68-96, from plonky2's CommonCircuitData :224-270)
    that fixes the proof's wire format and the FRI instance.  Defaults: standard_recursion_config
    (135 wires, 80 routed, 2 constants + 2 selectors, 2 challenges, 9 partial products, quotient degree factor 8)."""
    fri_params: "FriParams"
    num_constants: int = 4
    num_routed_wires: int = 80
    num_wires: int = 135
    num_challenges: int = 2
    num_partial_products: int = 9
    quotient_degree_factor: int = 8
    num_public_inputs: int = 4

    def to_c(self) -> PlonkCommon:
        return PlonkCommon(self.num_constants, self.num_routed_wires, self.num_wires, self.num_challenges,
                           self.num_partial_products, self.quotient_degree_factor, self.num_public_inputs)

    @staticmethod
    def for_params(params: "FriParams", num_public_inputs: int = 4, num_constants: int = 4) -> "CommonData":
        """A CommonData consistent with the oracle widths of `params` (CommonData::fri_oracles, :195-221)."""
        w = list(params.oracle_num_polys)
        nch = params.num_zs
        assert w[2] % nch == 0 and w[3] % nch == 0 and w[0] >= num_constants
        return CommonData(params, num_constants, w[0] - num_constants, w[1], nch, w[2] // nch - 1, w[3] // nch,
                          num_public_inputs)


# ---------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def version() -> str:
    """sv_version(): library version + kernel revision"""
    return lib().sv_version().decode()


def lib_path() -> str:
    """The in-tree library; SVB200_LIB overrides it (e.g. an AddressSanitizer build of the same sources for the CPU tests)."""
    return os.environ.get("SVB200_LIB") or os.path.join(_HERE, "libsvb200.so")


def lib() -> ctypes.CDLL:
    """Load libsvb200.so (built in-tree by build.sh / __graft_entry__.build)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise SvError(f"{p} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                      f"(there is no CPU fallback)")
    L = ctypes.CDLL(p)
    vp, u32p, u64 = ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint64
    L.sv_version.restype = ctypes.c_char_p
    L.sv_last_error.restype = ctypes.c_char_p
    L.sv_last_error.argtypes = [vp]
    L.sv_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    L.sv_ctx_destroy.argtypes = [vp]
    L.sv_ctx_destroy.restype = None
    L.sv_ctx_set_stream.argtypes = [vp, vp]
    L.sv_ctx_synchronize.argtypes = [vp]
    L.sv_ctx_launch_count.argtypes = [vp]
    L.sv_ctx_launch_count.restype = u64
    L.sv_ctx_kernel_timing.argtypes = [vp, ctypes.c_int]
    L.sv_ctx_kernel_time_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]
    L.sv_host_alloc.argtypes = [ctypes.c_size_t, ctypes.POINTER(vp)]
    L.sv_host_free.argtypes = [vp]
    L.sv_fri_layout_make.argtypes = [ctypes.POINTER(FriShape), ctypes.POINTER(Layout)]
    L.sv_poseidon_permute_batch.argtypes = [vp, vp, vp, ctypes.c_size_t, ctypes.c_int, ctypes.c_int]
    L.sv_poseidon_permute_batch_coop.argtypes = [vp, vp, vp, ctypes.c_size_t, ctypes.c_int]
    L.sv_goldilocks_mul_add_batch.argtypes = [vp, vp, vp, vp, vp, ctypes.c_size_t, ctypes.c_int]
    L.sv_merkle_verify_batch.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_int,
                                         vp, vp, vp, vp, ctypes.c_size_t, ctypes.c_int]
    L.sv_merkle_tree_build.argtypes = [vp, ctypes.c_int, ctypes.c_uint32, vp, ctypes.c_size_t, ctypes.c_uint32, vp, ctypes.c_int]
    L.sv_fri_verify_batch.argtypes = [vp, ctypes.POINTER(FriShape), ctypes.c_size_t, vp, vp, vp, ctypes.c_int]
    L.sv_allgather_bitmap.argtypes = [vp, vp, vp, vp, ctypes.c_size_t]
    L.sv_fri_challenges.argtypes = [ctypes.POINTER(FriShape), vp, vp, vp, ctypes.c_uint32]
    L.sv_synth_proofs.argtypes = [ctypes.POINTER(FriShape), u64, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_uint32,
                                  vp, ctypes.c_int]
    L.sv_synth_proofs_pi.argtypes = [ctypes.POINTER(FriShape), u64, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_uint32,
                                     vp, vp, ctypes.c_int]
    L.sv_synth_public_inputs.argtypes = [ctypes.POINTER(FriShape), u64, ctypes.c_uint32, ctypes.c_size_t, vp, vp]
    L.sv_fri_challenges_batch.argtypes = [vp, ctypes.POINTER(FriShape), ctypes.c_size_t, vp, vp, vp, ctypes.c_uint32, ctypes.c_int]
    L.sv_fri_verify_batch_fs.argtypes = [vp, ctypes.POINTER(FriShape), ctypes.c_size_t, vp, vp, vp, ctypes.c_uint32, vp, vp,
                                         ctypes.c_int]
    sp, cp = ctypes.POINTER(FriShape), ctypes.POINTER(PlonkCommon)
    L.sv_fri_shape_from_common.argtypes = [cp] + [ctypes.c_uint32] * 6 + [vp, ctypes.c_uint32, ctypes.c_uint32, sp]
    L.sv_wire_proof_bytes.argtypes = [sp, cp]
    L.sv_wire_proof_bytes.restype = ctypes.c_size_t
    L.sv_wire_pack.argtypes = [sp, cp, vp, vp, vp]
    L.sv_wire_unpack_batch.argtypes = [sp, cp, vp, vp, ctypes.c_size_t, ctypes.c_size_t, vp, vp, vp, vp, ctypes.c_int]
    L.sv_wire_unpack_batch_gpu.argtypes = [vp, sp, cp, vp, vp, ctypes.c_size_t, ctypes.c_size_t, vp, vp, vp, ctypes.c_int]
    L.sv_verify_proofs_wire.argtypes = [vp, sp, cp, vp, vp, vp, ctypes.c_size_t, ctypes.c_size_t, vp, vp]
    L.sv_public_inputs_hash.argtypes = [vp, ctypes.c_size_t, vp]
    pc = ctypes.POINTER(PlonkCircuit)
    L.sv_plonk_circuit_check.argtypes = [pc]
    L.sv_circuit_from_common_data.argtypes = [vp, sp, pc]
    L.sv_plonk_gate_from_id.argtypes = [ctypes.c_char_p, ctypes.POINTER(PlonkGate)]
    L.sv_plonk_challenges.argtypes = [sp, vp, vp, vp, ctypes.c_uint32, vp]
    L.sv_plonk_check_host.argtypes = [sp, pc, ctypes.c_size_t, vp, vp, vp, vp, ctypes.c_int]
    L.sv_plonk_check_batch.argtypes = [vp, sp, pc, ctypes.c_size_t, vp, vp, vp, vp, ctypes.c_int]
    L.sv_ntt_batch.argtypes = [vp, ctypes.c_uint32, ctypes.c_size_t, vp, ctypes.c_int, ctypes.c_int]
    L.sv_lde_batch.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t, vp, u64, vp, ctypes.c_int]
    L.sv_commit_batch.argtypes = [vp, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t, vp, ctypes.c_uint32, ctypes.c_int, vp, vp, ctypes.c_int]
    L.sv_ntt_host.argtypes = [ctypes.c_uint32, ctypes.c_size_t, vp, ctypes.c_int, ctypes.c_int]
    L.sv_lde_host.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_size_t, vp, u64, vp, ctypes.c_int]
    L.sv_verify_proofs_full.argtypes = [vp, sp, pc, vp, vp, vp, ctypes.c_size_t, ctypes.c_size_t, vp, vp]
    _LIB = L
    return L


def make_layout(params_or_shape) -> Layout:
    s = params_or_shape.to_shape() if isinstance(params_or_shape, FriParams) else params_or_shape
    out = Layout()
    if lib().sv_fri_layout_make(ctypes.byref(s), ctypes.byref(out)) != 0:
        raise SvError("bad FRI shape")
    return out


def _ptr(a) -> int:
    """address of a numpy array / int device pointer"""
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    return int(a)


def synth_proofs(params: FriParams, n_proofs: int, seed: int = 0xB2000002, n_circuits: int = 1,
                 num_challenges: int = 2, nthreads: Optional[int] = None, out: Optional[np.ndarray] = None,
                 pi_hashes: Optional[np.ndarray] = None) -> np.ndarray:
    """Synthetic valid proofs as a (n_proofs, record_words) uint64 array (host side, CPU).  pi_hashes (n_proofs, 4):
    bind proof i to that public-inputs hash instead of one drawn from the seed."""
    s = params.to_shape()
    L = make_layout(s)
    if out is None:
        out = np.zeros((n_proofs, L.record_words), dtype=np.uint64)
    assert out.dtype == np.uint64 and out.size == n_proofs * L.record_words
    nthreads = nthreads or os.cpu_count() or 1
    if pi_hashes is not None:
        pi_hashes = np.ascontiguousarray(pi_hashes, dtype=np.uint64)
        assert pi_hashes.size == 4 * n_proofs
    rc = lib().sv_synth_proofs_pi(ctypes.byref(s), ctypes.c_uint64(seed), n_circuits, n_proofs, num_challenges,
                                  _ptr(pi_hashes) if pi_hashes is not None else None, _ptr(out), nthreads)
    if rc != 0:
        raise SvError(f"sv_synth_proofs failed: {rc}")
    return out


def synth_public_inputs(params: FriParams, n_proofs: int, seed: int = 0xB2000002, n_circuits: int = 1):
    """(circuit_digests[n_circuits,4], pi_hashes[n_proofs,4]) that synth_proofs used with the same arguments."""
    s = params.to_shape()
    nc = max(1, min(n_circuits, n_proofs))
    cd = np.zeros((nc, 4), dtype=np.uint64)
    ph = np.zeros((n_proofs, 4), dtype=np.uint64)
    rc = lib().sv_synth_public_inputs(ctypes.byref(s), ctypes.c_uint64(seed), n_circuits, n_proofs, _ptr(cd), _ptr(ph))
    if rc != 0:
        raise SvError(f"sv_synth_public_inputs failed: {rc}")
    return cd, ph


def fri_challenges(params: FriParams, record: np.ndarray, circuit_digest, pi_hash, num_challenges: int = 2) -> None:
    """Host-side Fiat-Shamir (plonk_verifier_chip.rs:55-154): rewrites the challenge fields of `record`."""
    s = params.to_shape()
    cd = np.asarray(circuit_digest, dtype=np.uint64)
    ph = np.asarray(pi_hash, dtype=np.uint64)
    rc = lib().sv_fri_challenges(ctypes.byref(s), _ptr(record), _ptr(cd), _ptr(ph), num_challenges)
    if rc != 0:
        raise SvError(f"sv_fri_challenges failed: {rc}")


# -- wire format ---------------------------------------------------------------------
def shape_from_common(common: CommonData) -> FriShape:
    """sv_fri_shape_from_common: CommonData::fri_oracles / fri_zs_polys (types/common_data.rs:153-221)."""
    p, c, out = common.fri_params, common.to_c(), FriShape()
    ab = (ctypes.c_uint32 * max(1, len(p.reduction_arity_bits)))(*p.reduction_arity_bits)
    rc = lib().sv_fri_shape_from_common(ctypes.byref(c), p.degree_bits, p.config.rate_bits, p.config.cap_height,
                                        p.config.num_query_rounds, p.config.proof_of_work_bits, len(p.reduction_arity_bits),
                                        ctypes.cast(ab, ctypes.c_void_p), int(p.hiding), p.hash_kind, ctypes.byref(out))
    if rc != 0:
        raise SvError(f"sv_fri_shape_from_common failed: {rc}")
    return out


def wire_proof_bytes(common: CommonData) -> int:
    """Length of one serialised ProofWithPublicInputs of this circuit."""
    s, c = common.fri_params.to_shape(), common.to_c()
    n = lib().sv_wire_proof_bytes(ctypes.byref(s), ctypes.byref(c))
    if n == 0:
        raise SvError("wire format: FriParams and CommonData disagree")
    return int(n)


def wire_pack(common: CommonData, records: np.ndarray, public_inputs: np.ndarray) -> np.ndarray:
    """records (n, record_words) + public_inputs (n, num_public_inputs) -> (n, proof_bytes) uint8, plonky2's
    ProofWithPublicInputs::to_bytes layout (host, CPU)."""
    s, c = common.fri_params.to_shape(), common.to_c()
    nb = wire_proof_bytes(common)
    records = np.ascontiguousarray(records, dtype=np.uint64)
    n = records.shape[0]
    public_inputs = np.ascontiguousarray(public_inputs, dtype=np.uint64).reshape(n, common.num_public_inputs)
    out = np.zeros((n, nb), dtype=np.uint8)
    for i in range(n):
        rc = lib().sv_wire_pack(ctypes.byref(s), ctypes.byref(c), _ptr(records[i]), _ptr(public_inputs[i]) if public_inputs.size else None,
                                _ptr(out[i]))
        if rc != 0:
            raise SvError(f"sv_wire_pack failed: {rc}")
    return out


def wire_unpack_batch(common: CommonData, constants_sigmas_cap, blob: np.ndarray, n_proofs: Optional[int] = None,
                      stride: Optional[int] = None, nthreads: int = 1):
    """CPU unpacker: (records, pi_hashes, public_inputs, malformed) of n proofs in `blob` (uint8)."""
    s, c = common.fri_params.to_shape(), common.to_c()
    L = make_layout(s)
    nb = wire_proof_bytes(common)
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    stride = nb if stride is None else stride
    n = (blob.size // stride if n_proofs is None else n_proofs)
    cap = np.ascontiguousarray(constants_sigmas_cap, dtype=np.uint64)
    assert cap.size == 4 << s.cap_height
    recs = np.zeros((n, L.record_words), dtype=np.uint64)
    pih = np.zeros((n, 4), dtype=np.uint64)
    pis = np.zeros((n, common.num_public_inputs), dtype=np.uint64)
    mal = np.zeros(n, dtype=np.uint8)
    rc = lib().sv_wire_unpack_batch(ctypes.byref(s), ctypes.byref(c), _ptr(cap), _ptr(blob), stride, n, _ptr(recs), _ptr(pih),
                                    _ptr(pis) if pis.size else None, _ptr(mal), nthreads)
    if rc != 0:
        raise SvError(f"sv_wire_unpack_batch failed: {rc}")
    return recs, pih, pis, mal


def public_inputs_hash(public_inputs) -> np.ndarray:
    """PlonkVerifierChip::get_public_inputs_hash (plonk_verifier_chip.rs:41-53), host."""
    pi = np.ascontiguousarray(public_inputs, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    rc = lib().sv_public_inputs_hash(_ptr(pi) if pi.size else None, pi.size, _ptr(out))
    if rc != 0:
        raise SvError(f"sv_public_inputs_hash failed: {rc}")
    return out


# -- plonk-level checks ----------------------------------------------------------------
def make_plonk_circuit(common: CommonData, gates, groups, k_is, num_gate_constraints: int) -> PlonkCircuit:
    """gates: [(kind, param)] or [(kind, param, param2, param3)] in CommonData.gates order; groups: [(lo, hi)] = SelectorsInfo.groups; the selector index
    of a gate is the group that contains its position (SelectorsInfo.selector_indices)."""
    c = PlonkCircuit()
    c.common = common.to_c()
    c.degree_bits = common.fri_params.degree_bits
    c.num_gate_constraints = num_gate_constraints
    c.num_selectors = len(groups)
    for s, (lo, hi) in enumerate(groups):
        c.group_lo[s], c.group_hi[s] = lo, hi
    c.num_gates = len(gates)
    for i, gate in enumerate(gates):
        kind, param, param2, param3 = (list(gate) + [0, 0])[:4]
        sel = [s for s, (lo, hi) in enumerate(groups) if lo <= i < hi]
        if len(sel) != 1:
            raise SvError(f"gate {i} is not in exactly one selector group")
        c.gates[i] = PlonkGate(kind, param, param2, param3, sel[0])
    for j, k in enumerate(k_is):
        c.k_is[j] = int(k)
    rc = lib().sv_plonk_circuit_check(ctypes.byref(c))
    if rc != 0:
        raise SvError(f"sv_plonk_circuit_check refused the circuit: {rc}")
    return c


class CommonCircuitDataC(ctypes.Structure):
    """sv_common_circuit_data"""
    _fields_ = [("common", PlonkCommon), ("rate_bits", ctypes.c_uint32), ("cap_height", ctypes.c_uint32), ("proof_of_work_bits", ctypes.c_uint32),
                ("num_query_rounds", ctypes.c_uint32), ("hiding", ctypes.c_uint32), ("degree_bits", ctypes.c_uint32),
                ("num_reduction_steps", ctypes.c_uint32), ("reduction_arity_bits", ctypes.POINTER(ctypes.c_uint32)),
                ("num_gate_constraints", ctypes.c_uint32), ("num_gates", ctypes.c_uint32), ("gate_ids", ctypes.POINTER(ctypes.c_char_p)),
                ("selector_indices", ctypes.POINTER(ctypes.c_uint32)), ("num_selector_groups", ctypes.c_uint32),
                ("group_starts", ctypes.POINTER(ctypes.c_uint32)), ("group_ends", ctypes.POINTER(ctypes.c_uint32)),
                ("num_k_is", ctypes.c_uint32), ("k_is", ctypes.POINTER(ctypes.c_uint64)), ("hash_kind", ctypes.c_uint32)]


def circuit_from_common_data(common: CommonData, gate_ids, selector_indices, groups, k_is, num_gate_constraints: int):
    """sv_circuit_from_common_data: what CommonData::from (types/common_data.rs:224-270) + CustomGateRef::from
    (chip/plonk/gates/mod.rs:138-196) make of a plonky2 CommonCircuitData -> (FriShape, PlonkCircuit).
    gate_ids: `gate.0.id()` strings; selector_indices / groups: SelectorsInfo; common.fri_params carries the FRI side."""
    p = common.fri_params
    u32 = ctypes.c_uint32
    cd = CommonCircuitDataC()
    cd.common = common.to_c()
    cd.rate_bits, cd.cap_height = p.config.rate_bits, p.config.cap_height
    cd.proof_of_work_bits, cd.num_query_rounds = p.config.proof_of_work_bits, p.config.num_query_rounds
    cd.hiding, cd.degree_bits = int(p.hiding), p.degree_bits
    ab = (u32 * max(1, len(p.reduction_arity_bits)))(*p.reduction_arity_bits)
    cd.num_reduction_steps, cd.reduction_arity_bits = len(p.reduction_arity_bits), ab
    cd.num_gate_constraints = num_gate_constraints
    ids = (ctypes.c_char_p * len(gate_ids))(*[g.encode() for g in gate_ids])
    cd.num_gates, cd.gate_ids = len(gate_ids), ids
    si = (u32 * len(selector_indices))(*selector_indices)
    cd.selector_indices = si
    gs, ge = (u32 * len(groups))(*[g[0] for g in groups]), (u32 * len(groups))(*[g[1] for g in groups])
    cd.num_selector_groups, cd.group_starts, cd.group_ends = len(groups), gs, ge
    ks = (ctypes.c_uint64 * len(k_is))(*[int(k) for k in k_is])
    cd.num_k_is, cd.k_is = len(k_is), ks
    cd.hash_kind = p.hash_kind
    shape, circuit = FriShape(), PlonkCircuit()
    rc = lib().sv_circuit_from_common_data(ctypes.byref(cd), ctypes.byref(shape), ctypes.byref(circuit))
    if rc != 0:
        raise SvError(f"sv_circuit_from_common_data refused the circuit data: {rc}")
    return shape, circuit


def plonk_gate_from_id(gate_id: str):
    """plonky2 gate id string -> (kind, param, param2, param3), the mapping of CustomGateRef::from (gates/mod.rs:138-196)."""
    g = PlonkGate()
    rc = lib().sv_plonk_gate_from_id(gate_id.encode(), ctypes.byref(g))
    if rc != 0:
        raise SvError(f"unknown gate id {gate_id!r} (the reference: unimplemented!())")
    return (g.kind, g.param, g.param2, g.param3)


def plonk_challenges(params: FriParams, record: np.ndarray, circuit_digest, pi_hash, num_challenges: int = 2) -> np.ndarray:
    """[betas | gammas | alphas] of one proof (plonk_verifier_chip.rs:65-103), host."""
    s = params.to_shape()
    cd = np.ascontiguousarray(circuit_digest, dtype=np.uint64)
    ph = np.ascontiguousarray(pi_hash, dtype=np.uint64)
    out = np.zeros(3 * num_challenges, dtype=np.uint64)
    rc = lib().sv_plonk_challenges(ctypes.byref(s), _ptr(record), _ptr(cd), _ptr(ph), num_challenges, _ptr(out))
    if rc != 0:
        raise SvError(f"sv_plonk_challenges failed: {rc}")
    return out


def plonk_check_host(params: FriParams, circuit: PlonkCircuit, records, pi_hashes, plonk_chal, nthreads: int = 1) -> np.ndarray:
    """The vanishing-polynomial identity of every record, on CPU threads (the function the device kernel runs)."""
    s = params.to_shape()
    records = np.ascontiguousarray(records, dtype=np.uint64)
    n = records.shape[0]
    pi_hashes = np.ascontiguousarray(pi_hashes, dtype=np.uint64)
    plonk_chal = np.ascontiguousarray(plonk_chal, dtype=np.uint64)
    assert pi_hashes.size == 4 * n and plonk_chal.size == 3 * circuit.common.num_challenges * n
    bm = np.zeros((n + 31) // 32, dtype=np.uint32)
    rc = lib().sv_plonk_check_host(ctypes.byref(s), ctypes.byref(circuit), n, _ptr(records), _ptr(pi_hashes), _ptr(plonk_chal),
                                   _ptr(bm), nthreads)
    if rc != 0:
        raise SvError(f"sv_plonk_check_host failed: {rc}")
    return bm


# -- commit-phase library ---------------------------------------------------------------
def ntt_host(polys: np.ndarray, inverse: bool = False, nthreads: int = 1) -> np.ndarray:
    """(n_polys, 2^k) coefficients -> evaluations at omega^bitrev(i) (or back), on CPU threads (the kernels' function)."""
    a = np.array(polys, dtype=np.uint64, copy=True, order="C").reshape(-1, np.shape(polys)[-1])
    k = int(a.shape[1]).bit_length() - 1
    assert a.shape[1] == 1 << k
    rc = lib().sv_ntt_host(k, a.shape[0], _ptr(a), int(inverse), nthreads)
    if rc != 0:
        raise SvError(f"sv_ntt_host failed: {rc}")
    return a


def lde_host(coeffs: np.ndarray, rate_bits: int, shift: int = 7, nthreads: int = 1) -> np.ndarray:
    """(n_polys, 2^k) coefficients -> (n_polys, 2^(k+rate_bits)) values at shift * omega_N^bitrev(i), on CPU threads."""
    a = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, np.shape(coeffs)[-1])
    k = int(a.shape[1]).bit_length() - 1
    assert a.shape[1] == 1 << k
    out = np.zeros((a.shape[0], a.shape[1] << rate_bits), dtype=np.uint64)
    rc = lib().sv_lde_host(k, rate_bits, a.shape[0], _ptr(a), ctypes.c_uint64(shift), _ptr(out), nthreads)
    if rc != 0:
        raise SvError(f"sv_lde_host failed: {rc}")
    return out


class Context:
    """sv_ctx: one per (host thread, GPU)."""

    def __init__(self, device: int = 0):
        self._lib = lib()
        h = ctypes.c_void_p()
        rc = self._lib.sv_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise SvError(f"sv_ctx_create({device}) failed ({rc}): {self._lib.sv_last_error(None).decode()}")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.sv_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc, what):
        if rc != 0:
            raise SvError(f"{what} failed ({rc}): {self._lib.sv_last_error(self._h).decode()}")

    def set_stream(self, cuda_stream: int):
        self._ck(self._lib.sv_ctx_set_stream(self._h, ctypes.c_void_p(cuda_stream)), "sv_ctx_set_stream")

    def synchronize(self):
        self._ck(self._lib.sv_ctx_synchronize(self._h), "sv_ctx_synchronize")

    def kernel_timing(self, enable: bool = True):
        self._ck(self._lib.sv_ctx_kernel_timing(self._h, int(enable)), "sv_ctx_kernel_timing")

    def kernel_time_ms(self):
        ms, n = ctypes.c_double(), ctypes.c_uint64()
        self._ck(self._lib.sv_ctx_kernel_time_ms(self._h, ctypes.byref(ms), ctypes.byref(n)), "sv_ctx_kernel_time_ms")
        return ms.value, int(n.value)

    @property
    def launch_count(self) -> int:
        return int(self._lib.sv_ctx_launch_count(self._h))

    # -- hot path -------------------------------------------------------------------------------
    def poseidon_permute_batch(self, states, out=None, n: Optional[int] = None, mem: int = MEM_HOST,
                               hash_kind: int = HASH_POSEIDON_GOLDILOCKS):
        if mem == MEM_HOST:
            states = np.ascontiguousarray(states, dtype=np.uint64)
            n = states.size // 12
            out = np.empty_like(states) if out is None else out
        self._ck(self._lib.sv_poseidon_permute_batch(self._h, _ptr(states), _ptr(out), n, hash_kind, mem),
                 "sv_poseidon_permute_batch")
        return out

    def poseidon_permute_batch_coop(self, states):
        """Poseidon-Goldilocks on the lane-cooperative mapping of the device-side transcript (host arrays in / out)."""
        states = np.ascontiguousarray(states, dtype=np.uint64)
        out = np.empty_like(states)
        self._ck(self._lib.sv_poseidon_permute_batch_coop(self._h, _ptr(states), _ptr(out), states.size // 12, MEM_HOST),
                 "sv_poseidon_permute_batch_coop")
        return out

    def goldilocks_mul_add_batch(self, a, b, c):
        """a*b + c mod p on the device (host arrays in, canonical host array out)."""
        a = np.ascontiguousarray(a, dtype=np.uint64); b = np.ascontiguousarray(b, dtype=np.uint64)
        c = np.ascontiguousarray(c, dtype=np.uint64)
        assert a.shape == b.shape == c.shape
        out = np.empty_like(a)
        self._ck(self._lib.sv_goldilocks_mul_add_batch(self._h, _ptr(a), _ptr(b), _ptr(c), _ptr(out), a.size, MEM_HOST),
                 "sv_goldilocks_mul_add_batch")
        return out

    def merkle_verify_batch(self, leaf_len: int, depth: int, cap_height: int, paths, indices, caps, ok=None,
                            n: Optional[int] = None, mem: int = MEM_HOST, hash_kind: int = HASH_POSEIDON_GOLDILOCKS):
        if mem == MEM_HOST:
            n = len(indices)
            ok = np.zeros(n, dtype=np.uint8) if ok is None else ok
        self._ck(self._lib.sv_merkle_verify_batch(self._h, leaf_len, depth, cap_height, hash_kind,
                                                  _ptr(paths), _ptr(indices), _ptr(caps), _ptr(ok), n, mem),
                 "sv_merkle_verify_batch")
        return ok

    def merkle_tree_build(self, leaves, leaf_len: int, cap_height: int, n_leaves: Optional[int] = None, layers_out=None,
                          mem: int = MEM_HOST, hash_kind: int = HASH_POSEIDON_GOLDILOCKS):
        """Digest layers of the Merkle tree over `leaves` (n x leaf_len), bottom-up, as a list of (m, 4) arrays
        (host mode) or the flat device buffer (device mode)."""
        if mem == MEM_HOST:
            leaves = np.ascontiguousarray(leaves, dtype=np.uint64).reshape(-1, leaf_len)
            n_leaves = leaves.shape[0]
            layers_out = np.zeros(4 * (2 * n_leaves - (1 << cap_height)), dtype=np.uint64)
        self._ck(self._lib.sv_merkle_tree_build(self._h, hash_kind, leaf_len, _ptr(leaves), n_leaves, cap_height, _ptr(layers_out), mem),
                 "sv_merkle_tree_build")
        if mem != MEM_HOST:
            return layers_out
        out, off, m = [], 0, n_leaves
        while m >= (1 << cap_height):
            out.append(layers_out[off:off + 4 * m].reshape(m, 4))
            off += 4 * m
            m >>= 1
        return out

    def fri_verify_batch(self, params: FriParams, records, n_proofs: Optional[int] = None, accept_bitmap=None,
                         first_fail=None, want_fail: bool = False, mem: int = MEM_HOST):
        s = params.to_shape()
        if mem == MEM_HOST:
            n_proofs = records.shape[0] if n_proofs is None else n_proofs
            accept_bitmap = np.zeros((n_proofs + 31) // 32, dtype=np.uint32) if accept_bitmap is None else accept_bitmap
            if want_fail and first_fail is None:
                first_fail = np.zeros(n_proofs, dtype=np.uint32)
        self._ck(self._lib.sv_fri_verify_batch(self._h, ctypes.byref(s), n_proofs, _ptr(records), _ptr(accept_bitmap),
                                               _ptr(first_fail) if first_fail is not None else None, mem),
                 "sv_fri_verify_batch")
        return (accept_bitmap, first_fail) if want_fail else accept_bitmap

    def fri_challenges_batch(self, params: FriParams, records, circuit_digest, pi_hashes, num_challenges: int = 2,
                             n_proofs: Optional[int] = None, mem: int = MEM_HOST):
        """Device-side Fiat-Shamir: rewrites the challenge fields of every record (in place)."""
        s = params.to_shape()
        cd = np.ascontiguousarray(circuit_digest, dtype=np.uint64)
        if mem == MEM_HOST:
            n_proofs = records.shape[0]
            pi_hashes = np.ascontiguousarray(pi_hashes, dtype=np.uint64)
        self._ck(self._lib.sv_fri_challenges_batch(self._h, ctypes.byref(s), n_proofs, _ptr(records), _ptr(cd), _ptr(pi_hashes),
                                                   num_challenges, mem), "sv_fri_challenges_batch")
        return records

    def fri_verify_batch_fs(self, params: FriParams, records, circuit_digest, pi_hashes, num_challenges: int = 2,
                            n_proofs: Optional[int] = None, accept_bitmap=None, first_fail=None, want_fail: bool = False,
                            mem: int = MEM_HOST):
        """get_challenges + verify_fri_proof: the challenge fields of `records` are derived on the device."""
        s = params.to_shape()
        cd = np.ascontiguousarray(circuit_digest, dtype=np.uint64)
        if mem == MEM_HOST:
            n_proofs = records.shape[0] if n_proofs is None else n_proofs
            pi_hashes = np.ascontiguousarray(pi_hashes, dtype=np.uint64)
            accept_bitmap = np.zeros((n_proofs + 31) // 32, dtype=np.uint32) if accept_bitmap is None else accept_bitmap
            if want_fail and first_fail is None:
                first_fail = np.zeros(n_proofs, dtype=np.uint32)
        self._ck(self._lib.sv_fri_verify_batch_fs(self._h, ctypes.byref(s), n_proofs, _ptr(records), _ptr(cd), _ptr(pi_hashes),
                                                  num_challenges, _ptr(accept_bitmap),
                                                  _ptr(first_fail) if first_fail is not None else None, mem),
                 "sv_fri_verify_batch_fs")
        return (accept_bitmap, first_fail) if want_fail else accept_bitmap

    def wire_unpack_batch(self, common: CommonData, constants_sigmas_cap, blob, n_proofs: Optional[int] = None,
                          stride: Optional[int] = None, records_out=None, pi_hashes_out=None, malformed_out=None,
                          mem: int = MEM_HOST):
        """GPU unpacker (wire_unpack_kernel + wire_pi_hash_kernel): (records, pi_hashes, malformed)."""
        s, c = common.fri_params.to_shape(), common.to_c()
        nb = wire_proof_bytes(common)
        stride = nb if stride is None else stride
        cap = np.ascontiguousarray(constants_sigmas_cap, dtype=np.uint64)
        if mem == MEM_HOST:
            blob = np.ascontiguousarray(blob, dtype=np.uint8)
            n_proofs = blob.size // stride if n_proofs is None else n_proofs
            L = make_layout(s)
            records_out = np.zeros((n_proofs, L.record_words), dtype=np.uint64)
            pi_hashes_out = np.zeros((n_proofs, 4), dtype=np.uint64)
            malformed_out = np.zeros(n_proofs, dtype=np.uint32)
        self._ck(self._lib.sv_wire_unpack_batch_gpu(self._h, ctypes.byref(s), ctypes.byref(c), _ptr(cap), _ptr(blob), stride, n_proofs,
                                                    _ptr(records_out), _ptr(pi_hashes_out) if pi_hashes_out is not None else None,
                                                    _ptr(malformed_out) if malformed_out is not None else None, mem),
                 "sv_wire_unpack_batch_gpu")
        return records_out, pi_hashes_out, malformed_out

    def verify_proofs_wire(self, common: CommonData, constants_sigmas_cap, circuit_digest, blob, n_proofs: Optional[int] = None,
                           stride: Optional[int] = None, want_fail: bool = False):
        """Serialised proofs (host uint8 array or pinned host pointer) -> accept bitmap: unpack, public-input hashes,
        transcript and FRI query phase all on the device."""
        s, c = common.fri_params.to_shape(), common.to_c()
        nb = wire_proof_bytes(common)
        stride = nb if stride is None else stride
        if isinstance(blob, np.ndarray):
            blob = np.ascontiguousarray(blob, dtype=np.uint8)
            n_proofs = blob.size // stride if n_proofs is None else n_proofs
        cap = np.ascontiguousarray(constants_sigmas_cap, dtype=np.uint64)
        cd = np.ascontiguousarray(circuit_digest, dtype=np.uint64)
        bitmap = np.zeros((n_proofs + 31) // 32, dtype=np.uint32)
        ff = np.zeros(n_proofs, dtype=np.uint32) if want_fail else None
        self._ck(self._lib.sv_verify_proofs_wire(self._h, ctypes.byref(s), ctypes.byref(c), _ptr(cap), _ptr(cd), _ptr(blob), stride,
                                                 n_proofs, _ptr(bitmap), _ptr(ff) if ff is not None else None),
                 "sv_verify_proofs_wire")
        return (bitmap, ff) if want_fail else bitmap

    def verify_proofs_full(self, common: CommonData, circuit: PlonkCircuit, constants_sigmas_cap, circuit_digest, blob,
                           n_proofs: Optional[int] = None, stride: Optional[int] = None, want_fail: bool = False):
        """The complete verifier on serialised proofs: wire format, public-inputs hash, transcript, vanishing-polynomial
        identity and FRI query phase, all on the device (sv_verify_proofs_full)."""
        s = common.fri_params.to_shape()
        nb = wire_proof_bytes(common)
        stride = nb if stride is None else stride
        if isinstance(blob, np.ndarray):
            blob = np.ascontiguousarray(blob, dtype=np.uint8)
            n_proofs = blob.size // stride if n_proofs is None else n_proofs
        cap = np.ascontiguousarray(constants_sigmas_cap, dtype=np.uint64)
        cd = np.ascontiguousarray(circuit_digest, dtype=np.uint64)
        bitmap = np.zeros((n_proofs + 31) // 32, dtype=np.uint32)
        ff = np.zeros(n_proofs, dtype=np.uint32) if want_fail else None
        self._ck(self._lib.sv_verify_proofs_full(self._h, ctypes.byref(s), ctypes.byref(circuit), _ptr(cap), _ptr(cd), _ptr(blob), stride,
                                                 n_proofs, _ptr(bitmap), _ptr(ff) if ff is not None else None),
                 "sv_verify_proofs_full")
        return (bitmap, ff) if want_fail else bitmap

    def plonk_check_batch(self, params: FriParams, circuit: PlonkCircuit, records, pi_hashes, plonk_chal,
                          n_proofs: Optional[int] = None, accept_bitmap=None, mem: int = MEM_HOST):
        """plonk_check_kernel: the vanishing-polynomial identity of every record, one GPU thread per proof."""
        s = params.to_shape()
        if mem == MEM_HOST:
            records = np.ascontiguousarray(records, dtype=np.uint64)
            n_proofs = records.shape[0]
            pi_hashes = np.ascontiguousarray(pi_hashes, dtype=np.uint64)
            plonk_chal = np.ascontiguousarray(plonk_chal, dtype=np.uint64)
            accept_bitmap = np.zeros((n_proofs + 31) // 32, dtype=np.uint32)
        self._ck(self._lib.sv_plonk_check_batch(self._h, ctypes.byref(s), ctypes.byref(circuit), n_proofs, _ptr(records),
                                                _ptr(pi_hashes), _ptr(plonk_chal), _ptr(accept_bitmap), mem),
                 "sv_plonk_check_batch")
        return accept_bitmap

    def ntt_batch(self, data, log_n: Optional[int] = None, n_polys: Optional[int] = None, inverse: bool = False, mem: int = MEM_HOST):
        """In-place batch NTT (ntt_pass_kernel).  Host mode: returns a transformed copy of the (n_polys, 2^k) array."""
        if mem == MEM_HOST:
            data = np.array(data, dtype=np.uint64, copy=True, order="C").reshape(-1, np.shape(data)[-1])
            n_polys, log_n = data.shape[0], int(data.shape[1]).bit_length() - 1
        self._ck(self._lib.sv_ntt_batch(self._h, log_n, n_polys, _ptr(data), int(inverse), mem), "sv_ntt_batch")
        return data

    def lde_batch(self, coeffs, rate_bits: int, shift: int = 7, log_n: Optional[int] = None, n_polys: Optional[int] = None, out=None,
                  mem: int = MEM_HOST):
        """Low-degree extension onto shift * <omega_N> in Merkle-leaf order (lde_scale_pad_kernel + ntt_pass_kernel)."""
        if mem == MEM_HOST:
            coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, np.shape(coeffs)[-1])
            n_polys, log_n = coeffs.shape[0], int(coeffs.shape[1]).bit_length() - 1
            out = np.zeros((n_polys, coeffs.shape[1] << rate_bits), dtype=np.uint64)
        self._ck(self._lib.sv_lde_batch(self._h, log_n, rate_bits, n_polys, _ptr(coeffs), ctypes.c_uint64(shift), _ptr(out), mem),
                 "sv_lde_batch")
        return out

    def commit_batch(self, coeffs, rate_bits: int, cap_height: int, hash_kind: int = HASH_POSEIDON_GOLDILOCKS):
        """(n_polys, 2^k) coefficients -> (leaves (N, n_polys), digest layers bottom-up as in merkle_tree_build), host arrays:
        LDE onto 7 * <omega_N>, leaf-major copy and Merkle tree, all on the device (sv_commit_batch)."""
        coeffs = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, np.shape(coeffs)[-1])
        n_polys, log_n = coeffs.shape[0], int(coeffs.shape[1]).bit_length() - 1
        N = coeffs.shape[1] << rate_bits
        leaves = np.zeros((N, n_polys), dtype=np.uint64)
        flat = np.zeros(4 * (2 * N - (1 << cap_height)), dtype=np.uint64)
        self._ck(self._lib.sv_commit_batch(self._h, log_n, rate_bits, n_polys, _ptr(coeffs), cap_height, hash_kind, _ptr(leaves), _ptr(flat),
                                           MEM_HOST), "sv_commit_batch")
        layers, off, m = [], 0, N
        while m >= (1 << cap_height):
            layers.append(flat[off:off + 4 * m].reshape(m, 4))
            off += 4 * m
            m >>= 1
        return leaves, layers

    def allgather_bitmap(self, nccl_comm: int, local_ptr: int, all_ptr: int, words_per_rank: int):
        self._ck(self._lib.sv_allgather_bitmap(self._h, ctypes.c_void_p(nccl_comm), local_ptr, all_ptr, words_per_rank),
                 "sv_allgather_bitmap")


class PlonkVerifierChip:
    """Mirror of the reference's ``PlonkVerifierChip`` (chip/plonk/plonk_verifier_chip.rs): the same three steps, same
    names, on batches and natively.  ``verify_proof_with_challenges`` returns one bool per proof where the reference
    fails the circuit."""

    def __init__(self, ctx: "Context", common_data: CommonData, circuit: PlonkCircuit):
        self.ctx, self.common_data, self.circuit = ctx, common_data, circuit
        self.fri_params = common_data.fri_params
        self.layout = make_layout(self.fri_params)

    def get_public_inputs_hash(self, public_inputs) -> np.ndarray:
        """plonk_verifier_chip.rs:41-53 (PublicInputsHasherChip: Poseidon-Goldilocks sponge)."""
        return public_inputs_hash(public_inputs)

    def get_challenges(self, public_inputs_hash_, circuit_digest, record: np.ndarray) -> np.ndarray:
        """plonk_verifier_chip.rs:55-154: fills the FRI challenges (zeta, alpha, betas, PoW response, query indices) into
        the record header and returns the plonk challenges [betas | gammas | alphas].  Host side."""
        nch = self.common_data.num_challenges
        fri_challenges(self.fri_params, record, circuit_digest, public_inputs_hash_, nch)
        return plonk_challenges(self.fri_params, record, circuit_digest, public_inputs_hash_, nch)

    def verify_proof_with_challenges(self, records: np.ndarray, public_inputs_hashes, plonk_challenges_) -> List[bool]:
        """plonk_verifier_chip.rs:156-240: vanishing-polynomial identity, then FriVerifierChip::verify_fri_proof."""
        records = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, self.layout.record_words)
        n = records.shape[0]
        pl = self.ctx.plonk_check_batch(self.fri_params, self.circuit, records, public_inputs_hashes, plonk_challenges_)
        fri = self.ctx.fri_verify_batch(self.fri_params, records, n)
        return [bool((int(pl[i >> 5]) & int(fri[i >> 5])) >> (i & 31) & 1) for i in range(n)]


def verify_batch(ctx: "Context", proofs: Sequence[bytes], constants_sigmas_cap, circuit_digest, common: CommonData,
                 gate_ids: Sequence[str], selector_groups, k_is, num_gate_constraints: int) -> List[bool]:
    """The drop-in for a loop over the reference's ``verify_inside_snark_mock(degree, (proof_with_pis, vd, cd))``
    (verifier_api.rs:34-56): `proofs` are ``ProofWithPublicInputs::to_bytes()`` strings of ONE circuit, the verifier key is
    ``VerifierOnlyCircuitData`` (constants_sigmas_cap, circuit_digest), the circuit description comes from
    ``CommonCircuitData`` (gate ids as the reference matches them, gates/mod.rs:138-196; selector groups; k_is).  One bool
    per proof: plonk identity AND FRI proof, verified natively on the GPU -- invalidity is data, where the reference panics."""
    nb = wire_proof_bytes(common)
    for i, p in enumerate(proofs):
        if len(p) != nb:
            raise SvError(f"proof {i} has {len(p)} bytes, the circuit's proofs have {nb}")
    circuit = make_plonk_circuit(common, [plonk_gate_from_id(g) for g in gate_ids], selector_groups, k_is, num_gate_constraints)
    blob = np.frombuffer(b"".join(proofs), dtype=np.uint8) if proofs else np.zeros(0, dtype=np.uint8)
    bm = ctx.verify_proofs_full(common, circuit, constants_sigmas_cap, circuit_digest, blob, n_proofs=len(proofs))
    return [bool((int(bm[i >> 5]) >> (i & 31)) & 1) for i in range(len(proofs))]


class FriVerifierChip:
    """Mirror of the reference's seam: ``FriVerifierChip::construct(config, offset, fri_params)`` then
    ``verify_fri_proof(initial_merkle_caps, fri_challenges, fri_openings, fri_proof, fri_instance_info)``
    (chip/fri_chip.rs:35-46, 329-362).  Here the five arguments of one proof are one flat record
    (layout: include/stark_verifier_b200.h) and the call takes a batch of them; rejection is a bit in the
    returned bitmap instead of a panic."""

    OFFSET = 7  # GoldilocksField::MULTIPLICATIVE_GROUP_GENERATOR (plonk_verifier_chip.rs:225-227)

    def __init__(self, ctx: Context, fri_params: FriParams, offset: int = 7):
        if offset != self.OFFSET:
            raise SvError("the LDE coset shift is fixed to the multiplicative generator 7, like the reference")
        self.ctx = ctx
        self.fri_params = fri_params
        self.layout = make_layout(fri_params)

    def verify_fri_proof(self, records: np.ndarray, want_fail: bool = False):
        """records: (n_proofs, record_words) uint64 host array -> list[bool] (and fail codes)."""
        records = np.ascontiguousarray(records, dtype=np.uint64).reshape(-1, self.layout.record_words)
        n = records.shape[0]
        res = self.ctx.fri_verify_batch(self.fri_params, records, n, want_fail=want_fail)
        bitmap = res[0] if want_fail else res
        accept = [bool((int(bitmap[i >> 5]) >> (i & 31)) & 1) for i in range(n)]
        return (accept, res[1]) if want_fail else accept
