"""Proof containers, plonky2's wire bytes and the flat record of the C ABI -- TEST INFRASTRUCTURE, pure Python.

Containers mirror the reference's types (types/proof.rs:34-43,112-170,217-220,318-323,380-387; types/common_data.rs:10-54,
68-123).  Wire bytes = plonky2's `ProofWithPublicInputs::to_bytes()` at the revision Cargo.lock pins (0.1.0, DoHoonKim8 fork
72229c4; util/serialization.rs, not on disk -- restated from its published layout): every field element a little-endian
u64, a hash 4 of them, an extension element 2, NO length prefixes except one u8 (number of siblings) in front of every
Merkle proof; order = declaration order of the structs above, public inputs last.
The flat record is the documented input format of sv_fri_verify_batch (include/stark_verifier_b200.h, "Flat per-proof
record"): this module lays it out from that description alone."""
import struct

from . import gl
from .gl import P

OPENING_FIELDS = ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys")


class FriParams:
    """FriParams + FriConfig (types/common_data.rs:10-54) and what FriInstanceInfo adds (types/fri.rs:50-72,
    types/common_data.rs:153-221): oracle widths / blinding, and batch 1 = the first num_zs polynomials of oracle 2."""

    def __init__(self, degree_bits, rate_bits, cap_height, proof_of_work_bits, num_query_rounds, reduction_arity_bits,
                 oracle_num_polys, num_zs, hiding=False, oracle_blinding=(False, True, True, True), hash_kind=0):
        self.degree_bits, self.rate_bits, self.cap_height = degree_bits, rate_bits, cap_height
        self.proof_of_work_bits, self.num_query_rounds = proof_of_work_bits, num_query_rounds
        self.reduction_arity_bits = list(reduction_arity_bits)
        self.oracle_num_polys, self.num_zs = list(oracle_num_polys), num_zs
        self.hiding, self.oracle_blinding, self.hash_kind = bool(hiding), list(oracle_blinding), hash_kind

    def lde_bits(self):
        return self.degree_bits + self.rate_bits

    def final_poly_len(self):
        return 1 << (self.degree_bits - sum(self.reduction_arity_bits))

    def leaf_len(self, k):
        return self.oracle_num_polys[k] + (4 if self.hiding and self.oracle_blinding[k] else 0)


class Common:
    """the fields of CommonData / CircuitConfig that fix the vector lengths of a proof (types/common_data.rs:23-40,68-96)"""

    def __init__(self, num_constants, num_routed_wires, num_wires, num_challenges, num_partial_products,
                 quotient_degree_factor, num_public_inputs):
        self.num_constants, self.num_routed_wires, self.num_wires = num_constants, num_routed_wires, num_wires
        self.num_challenges, self.num_partial_products = num_challenges, num_partial_products
        self.quotient_degree_factor, self.num_public_inputs = quotient_degree_factor, num_public_inputs

    def oracle_num_polys(self):
        """CommonData::fri_oracles (types/common_data.rs:195-221)"""
        return [self.num_constants + self.num_routed_wires, self.num_wires,
                self.num_challenges * (1 + self.num_partial_products), self.num_challenges * self.quotient_degree_factor]

    def opening_lens(self):
        return {"constants": self.num_constants, "plonk_sigmas": self.num_routed_wires, "wires": self.num_wires,
                "plonk_zs": self.num_challenges, "plonk_zs_next": self.num_challenges,
                "partial_products": self.num_challenges * self.num_partial_products,
                "quotient_polys": self.num_challenges * self.quotient_degree_factor}

    @staticmethod
    def for_widths(oracle_num_polys, num_zs, num_public_inputs, num_routed_wires=None):
        """a Common whose oracle widths are the given ones (synthetic FRI-only proofs)"""
        w0, w1, w2, w3 = oracle_num_polys
        nr = num_routed_wires if num_routed_wires is not None else max(0, w0 - 4)
        assert w2 % num_zs == 0 and w3 % num_zs == 0
        return Common(w0 - nr, nr, w1, num_zs, w2 // num_zs - 1, w3 // num_zs, num_public_inputs)


class QueryStep:
    def __init__(self, evals, siblings):
        self.evals, self.siblings = evals, siblings            # [Fp2] (arity of them), [digest]


class QueryRound:
    def __init__(self, initial, steps):
        self.initial, self.steps = initial, steps              # [(evals [Fp], siblings [digest])] x 4, [QueryStep]


class FriProof:
    def __init__(self, commit_phase_merkle_caps, query_round_proofs, final_poly, pow_witness):
        self.commit_phase_merkle_caps, self.query_round_proofs = commit_phase_merkle_caps, query_round_proofs
        self.final_poly, self.pow_witness = final_poly, pow_witness


class Proof:
    def __init__(self, wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap, openings, opening_proof, public_inputs):
        self.wires_cap, self.plonk_zs_partial_products_cap = wires_cap, plonk_zs_partial_products_cap
        self.quotient_polys_cap, self.openings, self.opening_proof = quotient_polys_cap, openings, opening_proof
        self.public_inputs = public_inputs


# ---- wire bytes ---------------------------------------------------------------------------------------------------
def write_proof(proof):
    out = bytearray()
    w = lambda v: out.extend(struct.pack("<Q", int(v)))

    def cap(c):
        for h in c:
            for v in h:
                w(v)

    def merkle(sibs):
        out.append(len(sibs))
        for h in sibs:
            for v in h:
                w(v)
    cap(proof.wires_cap)
    cap(proof.plonk_zs_partial_products_cap)
    cap(proof.quotient_polys_cap)
    for f in OPENING_FIELDS:
        for ext in proof.openings[f]:
            w(ext[0]); w(ext[1])
    fp = proof.opening_proof
    for c in fp.commit_phase_merkle_caps:
        cap(c)
    for qr in fp.query_round_proofs:
        for evals, sibs in qr.initial:
            for v in evals:
                w(v)
            merkle(sibs)
        for st in qr.steps:
            for ext in st.evals:
                w(ext[0]); w(ext[1])
            merkle(st.siblings)
    for ext in fp.final_poly:
        w(ext[0]); w(ext[1])
    w(fp.pow_witness)
    for v in proof.public_inputs:
        w(v)
    return bytes(out)


class Malformed(Exception):
    pass


def read_proof(data, params, common):
    """the reader needs the circuit's CommonCircuitData for every length, like plonky2's from_bytes"""
    pos = [0]

    def r():
        if pos[0] + 8 > len(data):
            raise Malformed("truncated")
        v = struct.unpack_from("<Q", data, pos[0])[0]
        pos[0] += 8
        return v
    ncap = 1 << params.cap_height
    cap = lambda: [[r() for _ in range(4)] for _ in range(ncap)]

    def merkle(depth):
        if pos[0] >= len(data):
            raise Malformed("truncated")
        n = data[pos[0]]
        pos[0] += 1
        if n != depth:
            raise Malformed("Merkle proof of %d siblings where the shape says %d" % (n, depth))
        return [[r() for _ in range(4)] for _ in range(n)]
    wires_cap, zs_cap, q_cap = cap(), cap(), cap()
    lens = common.opening_lens()
    openings = {f: [(r(), r()) for _ in range(lens[f])] for f in OPENING_FIELDS}
    caps = [cap() for _ in params.reduction_arity_bits]
    lde = params.lde_bits()
    rounds = []
    for _ in range(params.num_query_rounds):
        initial = []
        for k in range(4):
            evals = [r() for _ in range(params.leaf_len(k))]
            initial.append((evals, merkle(lde - params.cap_height)))
        steps, bits = [], lde
        for ab in params.reduction_arity_bits:
            bits -= ab
            evals = [(r(), r()) for _ in range(1 << ab)]
            steps.append(QueryStep(evals, merkle(bits - params.cap_height)))
        rounds.append(QueryRound(initial, steps))
    final_poly = [(r(), r()) for _ in range(params.final_poly_len())]
    pow_witness = r()
    pis = [r() for _ in range(common.num_public_inputs)]
    if pos[0] != len(data):
        raise Malformed("trailing bytes")
    return Proof(wires_cap, zs_cap, q_cap, openings, FriProof(caps, rounds, final_poly, pow_witness), pis)


# ---- the flat record of include/stark_verifier_b200.h -----------------------------------------------------------------
def _up4(x):
    return (x + 3) & ~3


class RecordLayout:
    """word offsets; every segment starts on a 4-word boundary.  Header: init_caps 4 x ncap x 4 | step_caps S x ncap x 4 |
    open0 n0 x 2 | open1 n1 x 2 | final_poly len x 2 | pow_witness | alpha 2 | betas S x 2 | pow_response | indices Q |
    zeta 2 | zeta_next 2; then Q query blocks: for k < 4: evals[leaf_len k], siblings[init_depth x 4]; for each step:
    evals[2^arity_bits x 2], siblings[step_depth x 4]."""

    def __init__(self, p):
        S, ncap = len(p.reduction_arity_bits), 1 << p.cap_height
        self.ncap, self.n0, self.n1 = ncap, sum(p.oracle_num_polys), p.num_zs
        o = 0

        def seg(words):
            nonlocal o
            at = o
            o = _up4(o + words)
            return at
        self.off_init_caps = seg(4 * ncap * 4)
        self.off_step_caps = seg(S * ncap * 4)
        self.off_open0 = seg(2 * self.n0)
        self.off_open1 = seg(2 * self.n1)
        self.off_final_poly = seg(2 * p.final_poly_len())
        self.off_pow_witness = seg(1)
        self.off_alpha = seg(2)
        self.off_betas = seg(2 * S)
        self.off_pow_response = seg(1)
        self.off_indices = seg(p.num_query_rounds)
        self.off_zeta = seg(2)
        self.off_zeta_next = seg(2)
        self.header_words = o
        o = 0
        self.init_depth = p.lde_bits() - p.cap_height
        self.q_off_init_evals, self.q_off_init_sibs = [], []
        for k in range(4):
            self.q_off_init_evals.append(seg(p.leaf_len(k)))
            self.q_off_init_sibs.append(seg(4 * self.init_depth))
        self.q_off_step_evals, self.q_off_step_sibs, self.step_depth = [], [], []
        bits = p.lde_bits()
        for ab in p.reduction_arity_bits:
            bits -= ab
            self.step_depth.append(bits - p.cap_height)
            self.q_off_step_evals.append(seg(2 << ab))
            self.q_off_step_sibs.append(seg(4 * (bits - p.cap_height)))
        self.query_words = o
        self.record_words = self.header_words + p.num_query_rounds * o


def to_record(params, proof, constants_sigmas_cap, challenges=None, zeta_next=None):
    """-> list of record_words ints.  challenges: the dict of challenger.get_challenges (None leaves the fields 0)."""
    from .challenger import fri_openings
    L = RecordLayout(params)
    rec = [0] * L.record_words
    flat = lambda xs: [int(v) for x in xs for v in x]

    def put(off, words):
        rec[off:off + len(words)] = words
    caps = [constants_sigmas_cap, proof.wires_cap, proof.plonk_zs_partial_products_cap, proof.quotient_polys_cap]
    for k in range(4):
        put(L.off_init_caps + k * 4 * L.ncap, flat(caps[k]))
    fp = proof.opening_proof
    for i, c in enumerate(fp.commit_phase_merkle_caps):
        put(L.off_step_caps + i * 4 * L.ncap, flat(c))
    b0, b1 = fri_openings(proof.openings)
    assert len(b0) == L.n0 and len(b1) == L.n1
    put(L.off_open0, flat(b0))
    put(L.off_open1, flat(b1))
    put(L.off_final_poly, flat(fp.final_poly))
    rec[L.off_pow_witness] = int(fp.pow_witness)
    if challenges is not None:
        put(L.off_alpha, list(challenges["fri_alpha"]))
        put(L.off_betas, flat(challenges["fri_betas"]))
        rec[L.off_pow_response] = challenges["fri_pow_response"]
        put(L.off_indices, list(challenges["fri_query_indices"]))
        put(L.off_zeta, list(challenges["plonk_zeta"]))
        put(L.off_zeta_next, list(zeta_next))
    for q, qr in enumerate(fp.query_round_proofs):
        qb = L.header_words + q * L.query_words
        for k, (evals, sibs) in enumerate(qr.initial):
            put(qb + L.q_off_init_evals[k], [int(v) for v in evals])
            put(qb + L.q_off_init_sibs[k], flat(sibs))
        for i, st in enumerate(qr.steps):
            put(qb + L.q_off_step_evals[i], flat(st.evals))
            put(qb + L.q_off_step_sibs[i], flat(st.siblings))
    return rec


def from_record(params, common, rec):
    """record -> (Proof without public inputs, constants_sigmas_cap, challenges dict, zeta_next): parameter plumbing so
    that records corrupted word by word can be judged by the structured verifier."""
    L = RecordLayout(params)
    rec = [int(v) for v in rec]
    digests = lambda off, n: [rec[off + 4 * j: off + 4 * j + 4] for j in range(n)]
    exts = lambda off, n: [(rec[off + 2 * j], rec[off + 2 * j + 1]) for j in range(n)]
    caps = [digests(L.off_init_caps + k * 4 * L.ncap, L.ncap) for k in range(4)]
    S = len(params.reduction_arity_bits)
    step_caps = [digests(L.off_step_caps + i * 4 * L.ncap, L.ncap) for i in range(S)]
    b0, b1 = exts(L.off_open0, L.n0), exts(L.off_open1, L.n1)
    lens = common.opening_lens()
    openings, at = {}, 0
    for f in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys"):
        openings[f] = b0[at:at + lens[f]]
        at += lens[f]
    assert at == len(b0)
    openings["plonk_zs_next"] = b1
    rounds = []
    for q in range(params.num_query_rounds):
        qb = L.header_words + q * L.query_words
        initial = [(rec[qb + L.q_off_init_evals[k]: qb + L.q_off_init_evals[k] + params.leaf_len(k)],
                    digests(qb + L.q_off_init_sibs[k], L.init_depth)) for k in range(4)]
        steps = [QueryStep(exts(qb + L.q_off_step_evals[i], 1 << ab), digests(qb + L.q_off_step_sibs[i], L.step_depth[i]))
                 for i, ab in enumerate(params.reduction_arity_bits)]
        rounds.append(QueryRound(initial, steps))
    fp = FriProof(step_caps, rounds, exts(L.off_final_poly, params.final_poly_len()), rec[L.off_pow_witness])
    proof = Proof(caps[1], caps[2], caps[3], openings, fp, [])
    ch = {"fri_alpha": tuple(rec[L.off_alpha:L.off_alpha + 2]), "fri_betas": exts(L.off_betas, S),
          "fri_pow_response": rec[L.off_pow_response], "fri_query_indices": rec[L.off_indices:L.off_indices + params.num_query_rounds],
          "plonk_zeta": tuple(rec[L.off_zeta:L.off_zeta + 2])}
    return proof, caps[0], ch, tuple(rec[L.off_zeta_next:L.off_zeta_next + 2])
