"""TEST INFRASTRUCTURE: a complete plonky2-style prover for the toy circuits of tests/plonk_prover.py -- commitments,
Fiat-Shamir, openings and the FRI opening proof -- so that one proof exercises the WHOLE verifier (wire format ->
public-inputs hash -> transcript -> vanishing-polynomial identity -> FRI query phase) the way a real plonky2 proof would.

Pure Python end to end: the algebra of the plonk argument is tests/plonk_prover.py, and hashing, Merkle trees, the
transcript, the FRI commit phase (any 2^k-ary reduction), the wire bytes and the flat record are tests/pyref/ -- written from
the reference's Rust sources and sharing NO code with oracle/ or stark-verifier_b200/ (the only uses of the product package
below read numbers out of its parameter objects).  So a proof made here and accepted by the C oracle, by the product's host
twins and by the CUDA kernels is evidence from an independent implementation of both sides of the protocol.

The toy traces have 2^4 .. 2^7 rows; with ConstantArityBits(1, 5) (bn245_poseidon/plonky2_config.rs:84) that means 0, 1 or 2
arity-2 reduction steps; other reduction strategies are passed as reduction_arity_bits."""
import numpy as np

import plonk_prover as pp
from pyref import challenger as pch
from pyref import proof as ppf
from pyref import prover as ppr


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def pyref_params(params):
    """the product's FriParams object -> pyref.proof.FriParams (parameter plumbing: numbers only)"""
    if isinstance(params, ppf.FriParams):
        return params
    c = params.config
    return ppf.FriParams(params.degree_bits, c.rate_bits, c.cap_height, c.proof_of_work_bits, c.num_query_rounds,
                         list(params.reduction_arity_bits), list(params.oracle_num_polys), params.num_zs, hiding=params.hiding,
                         oracle_blinding=[bool(b) for b in params.oracle_blinding], hash_kind=params.hash_kind)


def pyref_common(C, num_public_inputs):
    return ppf.Common(C.num_constants, C.num_routed_wires, C.num_wires, C.num_challenges, C.num_partial_products, C.qdf,
                      num_public_inputs)


def prove(C, params, seed, public_inputs, circuit_digest):
    """-> dict(proof = pyref Proof, vk_cap, challenges, zeta_next, record (list of ints), blob (wire bytes), pi_hash, and the
    plonk prover's own outputs: open0/open1/betas/gammas/alphas/zeta/polys)"""
    p = pyref_params(params)
    assert p.final_poly_len() == C.n >> sum(p.reduction_arity_bits)
    common = pyref_common(C, len(public_inputs))
    pr = ppr.Prover(p, circuit_digest, public_inputs, seed ^ 0x5A17)
    nch = C.num_challenges

    def draw_betas_gammas(const_polys, sigma_polys, wire_polys):
        pr.commit(0, const_polys + sigma_polys)            # the verifier key's constants_sigmas_cap
        pr.commit(1, wire_polys)
        return pr.draw_betas_gammas(nch)

    def draw_alphas(z_polys, pp_polys):
        pr.commit(2, z_polys + [q for i in range(nch) for q in pp_polys[i]])
        return pr.draw_alphas(nch)

    def draw_zeta(quotient_chunks):
        pr.commit(3, [q for i in range(nch) for q in quotient_chunks[i]])
        return pr.draw_zeta()

    out = pp.prove(C, seed, pr.pi_hash, draw_betas_gammas, draw_alphas, draw_zeta)
    proof = pr.open_and_prove(common)
    # the plonk prover evaluated its own openings: they are the ones the FRI prover derived from the committed polynomials
    b0, b1 = pch.fri_openings(proof.openings)
    assert [tuple(e) for e in out["open0"]] == [tuple(e) for e in b0] and [tuple(e) for e in out["open1"]] == [tuple(e) for e in b1]
    ch = pch.get_challenges(proof, pr.pi_hash, pr.circuit_digest, nch, p.num_query_rounds, p.hash_kind)
    assert all(ch[k] == v for k, v in pr.challenges.items())
    assert (ch["plonk_betas"], ch["plonk_gammas"], ch["plonk_alphas"]) == (out["betas"], out["gammas"], out["alphas"])
    zn = pch.zeta_next(ch["plonk_zeta"], p.degree_bits)
    vk_cap = pr.trees[0].cap()
    out.update(proof=proof, vk_cap=vk_cap, challenges=ch, zeta_next=zn, pi_hash=pr.pi_hash, params=p, common=common,
               record=ppf.to_record(p, proof, vk_cap, ch, zn), blob=ppf.write_proof(proof))
    return out


def prove_full(C, params, seed, public_inputs, circuit_digest):
    """-> (flat record as a uint64 array with every field filled, dict of prove())"""
    out = prove(C, params, seed, [int(v) for v in public_inputs], [int(v) for v in circuit_digest])
    return np.array(out["record"], dtype=np.uint64), out


def toy_setup(svb, cfg, hiding=False, reduction_arity_bits=None):
    """(Circuit, FriParams of the product) for a plonk_prover configuration, with a FRI configuration that matches the
    reference's shape in miniature: rate 1/8, cap height 1, 2 PoW bits, 5 query rounds, arity-2 reduction down to 32
    coefficients (degree_bits <= 5: no reduction steps; 6: one; 7: two) unless reduction_arity_bits says otherwise."""
    C = pp.Circuit(**cfg)
    widths = (C.num_constants + C.num_routed_wires, C.num_wires, C.num_challenges * (1 + C.num_partial_products),
              C.num_challenges * C.qdf)
    params = svb.api._params(C.degree_bits, 3, 1, 2, 5, hiding=hiding, oracle_num_polys=widths, num_zs=C.num_challenges)
    if reduction_arity_bits is not None:
        params.reduction_arity_bits = list(reduction_arity_bits)
    return C, params
