"""ORACLE -- TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_build/liboracle.so.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcShape(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "degree_bits", "rate_bits", "cap_height", "num_query_rounds", "proof_of_work_bits",
        "num_steps", "final_poly_len", "hiding")] + [
        ("oracle_num_polys", ctypes.c_uint32 * 4), ("oracle_blinding", ctypes.c_uint32 * 4),
        ("num_zs", ctypes.c_uint32), ("hash_kind", ctypes.c_uint32), ("reduction_arity_bits", ctypes.c_uint32 * 32)]


class OrcLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "ncap", "lde_bits", "n0", "n1", "off_init_caps", "off_step_caps", "off_open0", "off_open1",
        "off_final_poly", "off_pow_witness", "off_alpha", "off_betas", "off_pow_response",
        "off_indices", "off_zeta", "off_zeta_next", "header_words")] + [
        ("leaf_len", ctypes.c_uint32 * 4), ("q_off_init_evals", ctypes.c_uint32 * 4),
        ("q_off_init_sibs", ctypes.c_uint32 * 4), ("init_depth", ctypes.c_uint32),
        ("q_off_step_evals", ctypes.c_uint32 * 32), ("q_off_step_sibs", ctypes.c_uint32 * 32),
        ("step_depth", ctypes.c_uint32 * 32), ("query_words", ctypes.c_uint32), ("record_words", ctypes.c_uint32)]


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.c", "wire.c", "plonk.c", "oracle.h", "orc_field.h", "poseidon_g_constants.h", "poseidon_b_constants.h",
                                             os.path.join("fast", "fast_avx512.c"))]
    fast = os.path.join(_HERE, "_build", "liboracle_fast.so")
    if force or not os.path.exists(so) or not os.path.exists(fast) or any(os.path.getmtime(s) > min(os.path.getmtime(so), os.path.getmtime(fast)) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        vp = ctypes.c_void_p
        L.orc_make_layout.argtypes = [ctypes.POINTER(OrcShape), ctypes.POINTER(OrcLayout)]
        L.orc_poseidon.argtypes = [vp]
        L.orc_poseidon_naive.argtypes = [vp]
        L.orc_poseidon_batch.argtypes = [vp, ctypes.c_size_t]
        L.orc_poseidon_b_fr.argtypes = [vp]
        L.orc_poseidon_b.argtypes = [vp]
        L.orc_set_hash_kind.argtypes = [ctypes.c_int]
        L.orc_hash_no_pad.argtypes = [vp, ctypes.c_size_t, vp]
        L.orc_two_to_one.argtypes = [vp, vp, vp]
        L.orc_merkle_verify.argtypes = [vp, ctypes.c_size_t, ctypes.c_uint64, vp, ctypes.c_size_t, vp, ctypes.c_size_t]
        L.orc_merkle_verify_batch.argtypes = [vp, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_size_t, vp, vp, ctypes.c_size_t, vp]
        L.orc_fri_verify.argtypes = [ctypes.POINTER(OrcShape), vp, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        L.orc_fri_verify_batch.argtypes = [ctypes.POINTER(OrcShape), vp, ctypes.c_size_t, vp, ctypes.c_int]
        L.orc_fri_challenges.argtypes = [ctypes.POINTER(OrcShape), vp, vp, vp, ctypes.c_uint32]
        u64 = ctypes.c_uint64
        L.orc_f_mul.argtypes = [u64, u64]; L.orc_f_mul.restype = u64
        L.orc_f_pow.argtypes = [u64, u64]; L.orc_f_pow.restype = u64
        L.orc_f_inv.argtypes = [u64]; L.orc_f_inv.restype = u64
        L.orc_f_red128.argtypes = [u64, u64, ctypes.c_int]; L.orc_f_red128.restype = u64
        L.orc_f2_mul.argtypes = [vp, vp, vp]
        L.orc_f2_inv.argtypes = [vp, vp]
        _LIB = L
    return _LIB


def shape_from(sv_shape) -> OrcShape:
    """Field-by-field copy of an sv_fri_shape-like ctypes struct."""
    s = OrcShape()
    for name, _ in OrcShape._fields_:
        v = getattr(sv_shape, name)
        if hasattr(v, "__len__"):
            setattr(s, name, (ctypes.c_uint32 * len(v))(*list(v)))
        else:
            setattr(s, name, v)
    return s


def layout(shape: OrcShape) -> OrcLayout:
    L = OrcLayout()
    assert lib().orc_make_layout(ctypes.byref(shape), ctypes.byref(L)) == 0
    return L


def poseidon(state, naive=False):
    a = np.array(state, dtype=np.uint64)
    (lib().orc_poseidon_naive if naive else lib().orc_poseidon)(a.ctypes.data)
    return a


def poseidon_batch(states):
    a = np.array(states, dtype=np.uint64, copy=True).reshape(-1, 12)
    lib().orc_poseidon_batch(a.ctypes.data, a.shape[0])
    return a


def poseidon_b(state):
    """Hash family B: the Goldilocks-wrapped Poseidon-BN254 permutation (plonky2_config.rs:38-51)."""
    a = np.array(state, dtype=np.uint64)
    lib().orc_poseidon_b(a.ctypes.data)
    return a


def poseidon_b_fr(state5):
    """Raw Poseidon-BN254 permutation on 5 canonical integers < r (native.rs:43-60) -> list of 5 ints."""
    a = np.zeros(20, dtype=np.uint64)
    for i, v in enumerate(state5):
        for k in range(4):
            a[4 * i + k] = (int(v) >> (64 * k)) & (2**64 - 1)
    lib().orc_poseidon_b_fr(a.ctypes.data)
    return [sum(int(a[4 * i + k]) << (64 * k) for k in range(4)) for i in range(5)]


def hash_no_pad(inp, kind=0):
    a = np.ascontiguousarray(inp, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_set_hash_kind(kind)
    lib().orc_hash_no_pad(a.ctypes.data, a.size, out.ctypes.data)
    lib().orc_set_hash_kind(0)
    return out


def two_to_one(l, r, kind=0):
    l = np.ascontiguousarray(l, dtype=np.uint64)
    r = np.ascontiguousarray(r, dtype=np.uint64)
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_set_hash_kind(kind)
    lib().orc_two_to_one(l.ctypes.data, r.ctypes.data, out.ctypes.data)
    lib().orc_set_hash_kind(0)
    return out


def merkle_verify_batch(records, leaf_len, depth, indices, caps, cap_height, kind=0):
    records = np.ascontiguousarray(records, dtype=np.uint64)
    indices = np.ascontiguousarray(indices, dtype=np.uint64)
    caps = np.ascontiguousarray(caps, dtype=np.uint64)
    ok = np.zeros(len(indices), dtype=np.uint8)
    lib().orc_set_hash_kind(kind)
    lib().orc_merkle_verify_batch(records.ctypes.data, len(indices), leaf_len, depth, indices.ctypes.data,
                                  caps.ctypes.data, cap_height, ok.ctypes.data)
    lib().orc_set_hash_kind(0)
    return ok


def fri_verify(shape: OrcShape, record):
    record = np.ascontiguousarray(record, dtype=np.uint64)
    f, fq = ctypes.c_int(), ctypes.c_int()
    ok = lib().orc_fri_verify(ctypes.byref(shape), record.ctypes.data, ctypes.byref(f), ctypes.byref(fq))
    return bool(ok), f.value, fq.value


def fri_verify_batch(shape: OrcShape, records, nthreads=1):
    records = np.ascontiguousarray(records, dtype=np.uint64)
    L = layout(shape)
    n = records.size // L.record_words
    bm = np.zeros((n + 31) // 32, dtype=np.uint32)
    lib().orc_fri_verify_batch(ctypes.byref(shape), records.ctypes.data, n, bm.ctypes.data, nthreads)
    return bm


def fri_challenges(shape: OrcShape, record, circuit_digest, pi_hash, num_challenges=2):
    cd = np.asarray(circuit_digest, dtype=np.uint64)
    ph = np.asarray(pi_hash, dtype=np.uint64)
    lib().orc_fri_challenges(ctypes.byref(shape), record.ctypes.data, cd.ctypes.data, ph.ctypes.data, num_challenges)


# -- wire format (oracle/wire.c) -------------------------------------------------------------------
class OrcCommon(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint32) for n in (
        "num_constants", "num_routed_wires", "num_wires", "num_challenges", "num_partial_products",
        "quotient_degree_factor", "num_public_inputs")]


def common_from(sv_common) -> OrcCommon:
    return OrcCommon(*[getattr(sv_common, n) for n, _ in OrcCommon._fields_])


def _wire_lib():
    L = lib()
    vp, sp, cp = ctypes.c_void_p, ctypes.POINTER(OrcShape), ctypes.POINTER(OrcCommon)
    L.orc_wire_proof_bytes.argtypes = [sp, cp]
    L.orc_wire_proof_bytes.restype = ctypes.c_size_t
    L.orc_wire_read_proof.argtypes = [sp, cp, vp, vp, ctypes.c_size_t, vp, vp, vp]
    L.orc_wire_write_proof.argtypes = [sp, cp, vp, vp, vp]
    return L


def wire_proof_bytes(shape: OrcShape, common: OrcCommon) -> int:
    return int(_wire_lib().orc_wire_proof_bytes(ctypes.byref(shape), ctypes.byref(common)))


def wire_read_proof(shape: OrcShape, common: OrcCommon, vk_cap, data):
    """-> (rc, record, public_inputs, pi_hash); rc 0 parsed, 1 malformed, < 0 cannot be framed."""
    L = layout(shape)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    cap = np.ascontiguousarray(vk_cap, dtype=np.uint64)
    rec = np.zeros(L.record_words, dtype=np.uint64)
    pis = np.zeros(max(1, common.num_public_inputs), dtype=np.uint64)
    pih = np.zeros(4, dtype=np.uint64)
    rc = _wire_lib().orc_wire_read_proof(ctypes.byref(shape), ctypes.byref(common), cap.ctypes.data, data.ctypes.data, data.size,
                                         rec.ctypes.data, pis.ctypes.data, pih.ctypes.data)
    return rc, rec, pis[:common.num_public_inputs], pih


def wire_write_proof(shape: OrcShape, common: OrcCommon, record, public_inputs):
    record = np.ascontiguousarray(record, dtype=np.uint64)
    pis = np.ascontiguousarray(public_inputs, dtype=np.uint64)
    out = np.zeros(wire_proof_bytes(shape, common), dtype=np.uint8)
    pp = pis.ctypes.data if pis.size else None
    rc = _wire_lib().orc_wire_write_proof(ctypes.byref(shape), ctypes.byref(common), record.ctypes.data, pp, out.ctypes.data)
    assert rc == 0, rc
    return out


# -- plonk-level checks (oracle/plonk.c) -----------------------------------------------------------
class OrcPlonkGate(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_uint32), ("param", ctypes.c_uint32), ("param2", ctypes.c_uint32), ("param3", ctypes.c_uint32),
                ("selector_index", ctypes.c_uint32)]


class OrcPlonkCircuit(ctypes.Structure):
    _fields_ = [("common", OrcCommon), ("degree_bits", ctypes.c_uint32), ("num_gate_constraints", ctypes.c_uint32),
                ("num_selectors", ctypes.c_uint32), ("group_lo", ctypes.c_uint32 * 8), ("group_hi", ctypes.c_uint32 * 8),
                ("num_gates", ctypes.c_uint32), ("gates", OrcPlonkGate * 32), ("k_is", ctypes.c_uint64 * 128)]


def plonk_circuit_from(sv_circuit) -> OrcPlonkCircuit:
    """Field-by-field copy of an sv_plonk_circuit-like ctypes struct."""
    c = OrcPlonkCircuit()
    c.common = common_from(sv_circuit.common)
    for name in ("degree_bits", "num_gate_constraints", "num_selectors", "num_gates"):
        setattr(c, name, getattr(sv_circuit, name))
    for i in range(8):
        c.group_lo[i], c.group_hi[i] = sv_circuit.group_lo[i], sv_circuit.group_hi[i]
    for i in range(32):
        g = sv_circuit.gates[i]
        c.gates[i] = OrcPlonkGate(g.kind, g.param, g.param2, g.param3, g.selector_index)
    for i in range(128):
        c.k_is[i] = sv_circuit.k_is[i]
    return c


def plonk_check(circuit: OrcPlonkCircuit, open0, open1, pi_hash, chal, zeta) -> int:
    L = lib()
    L.orc_plonk_check.argtypes = [ctypes.POINTER(OrcPlonkCircuit)] + [ctypes.c_void_p] * 5
    a = [np.ascontiguousarray(v, dtype=np.uint64) for v in (open0, open1, pi_hash, chal, zeta)]
    return int(L.orc_plonk_check(ctypes.byref(circuit), *[v.ctypes.data for v in a]))


def plonk_challenges(shape: OrcShape, record, circuit_digest, pi_hash, num_challenges=2):
    L = lib()
    L.orc_plonk_challenges.argtypes = [ctypes.POINTER(OrcShape), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_uint32, ctypes.c_void_p]
    rec = np.ascontiguousarray(record, dtype=np.uint64)
    cd = np.ascontiguousarray(circuit_digest, dtype=np.uint64)
    ph = np.ascontiguousarray(pi_hash, dtype=np.uint64)
    out = np.zeros(3 * num_challenges, dtype=np.uint64)
    L.orc_plonk_challenges(ctypes.byref(shape), rec.ctypes.data, cd.ctypes.data, ph.ctypes.data, num_challenges, out.ctypes.data)
    return out


# ---- oracle/fast: the AVX-512 arm of the CPU baseline (bench-only; same verdicts as orc_fri_verify_batch) -----------------
_FAST = None


def fast_available() -> bool:
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " avx512f" in flags and " avx512dq" in flags


def fast_lib():
    global _FAST
    if _FAST is None:
        build()
        so = os.path.join(_HERE, "_build", "liboracle_fast.so")
        L = ctypes.CDLL(so)
        L.orc_fast_poseidon8.argtypes = [ctypes.c_void_p]
        L.orc_fast_fri_verify_batch.argtypes = [ctypes.POINTER(OrcShape), ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int]
        _FAST = L
    return _FAST


def fast_poseidon8(states):
    a = np.array(states, dtype=np.uint64, copy=True).reshape(8, 12)
    fast_lib().orc_fast_poseidon8(a.ctypes.data)
    return a


def fast_fri_verify_batch(shape: OrcShape, records, nthreads=1):
    records = np.ascontiguousarray(records, dtype=np.uint64)
    n = records.shape[0]
    bm = np.zeros((n + 31) // 32, dtype=np.uint32)
    fast_lib().orc_fast_fri_verify_batch(ctypes.byref(shape), records.ctypes.data, n, bm.ctypes.data, nthreads)
    return bm
