"""TEST INFRASTRUCTURE: a complete plonky2-style prover for the toy circuits of tests/plonk_prover.py -- commitments,
Fiat-Shamir, openings and the FRI opening proof -- so that one proof exercises the WHOLE verifier (wire format ->
public-inputs hash -> transcript -> vanishing-polynomial identity -> FRI query phase) the way a real plonky2 proof would.

Pure Python for the algebra; hashing and the transcript go through the CPU oracle (oracle/oracle.c), which is pinned to
plonky2's Poseidon test vectors.  The trace has 2^4 rows, which is below the final-polynomial size of the reference's
FRI configuration (ConstantArityBits(1, 5), bn245_poseidon/plonky2_config.rs:84), so there are no reduction steps: the
final polynomial IS the batched DEEP quotient and the query phase checks it against the four oracle openings.  With
degree_bits = 6 or 7 the same prover runs one or two arity-2 reduction steps (commit phase, betas, step trees).

Conventions (the ones the verifier implies; same as the product's synthetic prover, stark-verifier_b200/csrc/host_side.cpp):
leaf i of an oracle tree holds the evaluations at 7 * omega^bitrev(i); batch 0 = every polynomial at zeta in oracle order,
batch 1 = the Z polynomials at g * zeta; final = q_0 * alpha^{|batch 1|} + q_1 with q_b = (r_b(x) - r_b(point_b)) / (x - point_b)."""
import numpy as np

import plonk_prover as pp
from plonk_prover import P, e_add, e_mul, e_scale, inv, poly_eval, poly_eval_ext


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


class Tree:
    def __init__(self, orc, leaves, cap_height, kind=0):
        """leaves: list of rows (lists of ints); digest layers bottom-up down to the cap."""
        dig = []
        for row in leaves:
            dig.append(list(orc.hash_no_pad(row, kind)) if len(row) > 4 else [int(v) for v in row] + [0] * (4 - len(row)))
        self.layers = [[[int(v) for v in d] for d in dig]]
        while len(self.layers[-1]) > (1 << cap_height):
            cur = self.layers[-1]
            self.layers.append([[int(v) for v in orc.two_to_one(cur[2 * j], cur[2 * j + 1], kind)] for j in range(len(cur) // 2)])
        self.leaves = leaves

    def cap(self):
        return [w for d in self.layers[-1] for w in d]

    def path(self, index):
        out = []
        for layer in self.layers[:-1]:
            out += layer[index ^ 1]
            index >>= 1
        return out


def ext_poly_divide_by_linear(coeffs, z):
    """(a(x) - a(z)) / (x - z) for a with Fp2 coefficients: synthetic division, remainder dropped."""
    b = [(0, 0)] * (len(coeffs) - 1)
    carry = (0, 0)
    for k in range(len(coeffs) - 1, 0, -1):
        carry = e_add(coeffs[k], e_mul(carry, z))
        b[k - 1] = carry
    return b


def poly_eval_ext_ext(coeffs, x):
    """polynomial with Fp2 coefficients at a base-field point"""
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = e_add(e_scale(acc, x), c)
    return acc


def prove_full(svb, orc, C, params, seed, public_inputs, circuit_digest):
    """-> (record with every field filled, dict of the plonk prover's outputs).  params: FriParams consistent with C and
    with no reduction steps."""
    L = svb.api.make_layout(params)
    oshape = orc.shape_from(params.to_shape())
    kind = params.hash_kind
    steps = len(params.reduction_arity_bits)
    assert params.final_poly_len() == C.n >> steps
    salt_rng = np.random.default_rng(seed ^ 0x5A17)
    lde_bits, cap_h, nch = L.lde_bits, params.config.cap_height, C.num_challenges
    N = 1 << lde_bits
    omega = pow(7, (P - 1) >> lde_bits, P)
    xs = [7 * pow(omega, bitrev(i, lde_bits), P) % P for i in range(N)]
    rec = np.zeros(L.record_words, dtype=np.uint64)
    pi_hash = [int(v) for v in svb.public_inputs_hash(public_inputs)]
    trees, oracle_polys = [None] * 4, [None] * 4

    def commit(k, polys):
        oracle_polys[k] = polys
        rows = [[poly_eval(p, x) for p in polys] for x in xs]
        if params.hiding and params.oracle_blinding[k]:     # salted leaves: 4 extra limbs at the end (types/assigned.rs:57-71)
            rows = [r + [int(v) for v in salt_rng.integers(0, P, size=4, dtype=np.uint64)] for r in rows]
        trees[k] = Tree(orc, rows, cap_h, kind)
        capw = 4 * L.ncap
        rec[L.off_init_caps + k * capw: L.off_init_caps + (k + 1) * capw] = trees[k].cap()

    def draw_betas_gammas(const_polys, sigma_polys, wire_polys):
        commit(0, const_polys + sigma_polys)            # the verifier key's constants_sigmas_cap
        commit(1, wire_polys)
        ch = orc.plonk_challenges(oshape, rec, circuit_digest, pi_hash, nch)
        return [int(v) for v in ch[:nch]], [int(v) for v in ch[nch:2 * nch]]

    def draw_alphas(z_polys, pp_polys):
        commit(2, z_polys + [p for i in range(nch) for p in pp_polys[i]])
        return [int(v) for v in orc.plonk_challenges(oshape, rec, circuit_digest, pi_hash, nch)[2 * nch:]]

    def draw_zeta(quotient_chunks):
        commit(3, [p for i in range(nch) for p in quotient_chunks[i]])
        orc.fri_challenges(oshape, rec, circuit_digest, pi_hash, nch)
        return (int(rec[L.off_zeta]), int(rec[L.off_zeta + 1]))

    out = pp.prove(C, seed, pi_hash, draw_betas_gammas, draw_alphas, draw_zeta)
    zeta = out["zeta"]
    o0 = [w for e in out["open0"] for w in e]
    o1 = [w for e in out["open1"] for w in e]
    assert len(o0) == 2 * L.n0 and len(o1) == 2 * L.n1
    rec[L.off_open0:L.off_open0 + len(o0)] = o0
    rec[L.off_open1:L.off_open1 + len(o1)] = o1
    orc.fri_challenges(oshape, rec, circuit_digest, pi_hash, nch)
    assert (int(rec[L.off_zeta]), int(rec[L.off_zeta + 1])) == zeta
    alpha = (int(rec[L.off_alpha]), int(rec[L.off_alpha + 1]))
    gz = e_scale(zeta, C.g)
    assert (int(rec[L.off_zeta_next]), int(rec[L.off_zeta_next + 1])) == gz

    # ---- FRI opening proof without reduction steps: the final polynomial is the batched DEEP quotient ----
    def batch_poly(polys):
        acc, ap = [(0, 0)] * C.n, (1, 0)
        for p in polys:
            acc = [e_add(a, e_scale(ap, c)) for a, c in zip(acc, p)]
            ap = e_mul(ap, alpha)
        return acc

    all_polys = [p for k in range(4) for p in oracle_polys[k]]
    z_polys = oracle_polys[2][:nch]
    r0, r1 = batch_poly(all_polys), batch_poly(z_polys)
    q0, q1 = ext_poly_divide_by_linear(r0, zeta), ext_poly_divide_by_linear(r1, gz)
    alpha_n1 = (1, 0)
    for _ in range(len(z_polys)):
        alpha_n1 = e_mul(alpha_n1, alpha)
    final = [e_add(e_mul(a, alpha_n1), b) for a, b in zip(q0, q1)] + [(0, 0)]
    assert len(final) == C.n
    # ---- commit phase (fri_chip.rs:168-226, 275-315 from the prover's side): layer st holds the values of the current
    # polynomial on shift_st * <omega_st>, leaf k of its tree = the coset pair (2k, 2k+1); f(x) = f_E(x^2) + x f_O(x^2)
    # folds to f_E + beta f_O, i.e. coefficient-wise c'_j = c_2j + beta c_2j+1
    step_vals, step_trees = [], []
    shift, bits_cur, w_cur = 7, lde_bits, omega
    for st in range(steps):
        vals = [poly_eval_ext_ext(final, shift * pow(w_cur, bitrev(i, bits_cur), P) % P) for i in range(1 << bits_cur)]
        tree = Tree(orc, [list(vals[2 * k]) + list(vals[2 * k + 1]) for k in range(len(vals) // 2)], cap_h, kind)
        capw = 4 * L.ncap
        rec[L.off_step_caps + st * capw: L.off_step_caps + (st + 1) * capw] = tree.cap()
        orc.fri_challenges(oshape, rec, circuit_digest, pi_hash, nch)
        beta = (int(rec[L.off_betas + 2 * st]), int(rec[L.off_betas + 2 * st + 1]))
        final = [e_add(final[2 * j], e_mul(beta, final[2 * j + 1])) for j in range(len(final) // 2)]
        step_vals.append(vals)
        step_trees.append(tree)
        shift, bits_cur, w_cur = shift * shift % P, bits_cur - 1, w_cur * w_cur % P
    assert len(final) == params.final_poly_len()
    rec[L.off_final_poly:L.off_final_poly + 2 * len(final)] = [w for e in final for w in e]
    # proof of work: the smallest witness whose response has proof_of_work_bits leading zero bits
    bits = params.config.proof_of_work_bits
    w = 0
    while True:
        rec[L.off_pow_witness] = w
        orc.fri_challenges(oshape, rec, circuit_digest, pi_hash, nch)
        if bits == 0 or int(rec[L.off_pow_response]) >> (64 - bits) == 0:
            break
        w += 1
    # ---- query rounds ----
    for q in range(params.config.num_query_rounds):
        idx = int(rec[L.off_indices + q]) & (N - 1)
        qb = L.header_words + q * L.query_words
        for k in range(4):
            row = trees[k].leaves[idx]
            rec[qb + L.q_off_init_evals[k]: qb + L.q_off_init_evals[k] + len(row)] = row
            path = trees[k].path(idx)
            assert len(path) == 4 * L.init_depth
            rec[qb + L.q_off_init_sibs[k]: qb + L.q_off_init_sibs[k] + len(path)] = path
        cur = idx
        for st in range(steps):
            coset = cur >> 1
            rec[qb + L.q_off_step_evals[st]: qb + L.q_off_step_evals[st] + 4] = list(step_vals[st][2 * coset]) + list(step_vals[st][2 * coset + 1])
            path = step_trees[st].path(coset)
            assert len(path) == 4 * L.step_depth[st]
            rec[qb + L.q_off_step_sibs[st]: qb + L.q_off_step_sibs[st] + len(path)] = path
            cur = coset
    out["pi_hash"] = pi_hash
    return rec, out


def toy_setup(svb, cfg, hiding=False):
    """(Circuit, FriParams) for a plonk_prover configuration, with a FRI configuration that matches the reference's shape
    in miniature: rate 1/8, cap height 1, 2 PoW bits, 5 query rounds, arity-2 reduction down to 32 coefficients
    (degree_bits <= 5: no reduction steps; 6: one; 7: two)."""
    C = pp.Circuit(**cfg)
    widths = (C.num_constants + C.num_routed_wires, C.num_wires, C.num_challenges * (1 + C.num_partial_products),
              C.num_challenges * C.qdf)
    params = svb.api._params(C.degree_bits, 3, 1, 2, 5, hiding=hiding, oracle_num_polys=widths, num_zs=C.num_challenges)
    return C, params
