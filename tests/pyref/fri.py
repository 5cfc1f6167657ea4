"""FriVerifierChip::verify_fri_proof (chip/fri_chip.rs:329-362) and everything under it -- TEST INFRASTRUCTURE, pure Python.

Written from fri_chip.rs, with ONE extension: the reference's next_eval folds arity 2 only (:211 "TODO: only 2-arity is
supported"), while its own demo produces ConstantArityBits(3, 5) proofs (plonky2_semaphore/access_set.rs:124) and checks them
with plonky2's native verifier.  fold() below is that verifier's compute_evaluation (plonky2 fri/verifier.rs): interpolate the
2^k coset values and evaluate at beta.  For k = 1 it is the reference's two-point formula (:212-224), which
tests/test_pyref.py checks against the literal expression.

The verdict is (accept, code, query) with the first-failure codes and order of include/stark_verifier_b200.h: the reference
panics at the FIRST failing assert, walking the query rounds in order and inside a round: range checks of the witnesses
(assign_value, native_chip/arithmetic_chip.rs:256-268), the four initial Merkle proofs, the DEEP-quotient divisions, then per
reduction step eval consistency / fold / step Merkle proof, finally the final polynomial."""
from . import gl
from . import merkle
from .challenger import fri_openings
from .gl import P

OK, FAIL_POW, FAIL_NONCANONICAL, FAIL_INIT_MERKLE, FAIL_ZERO_DENOM, FAIL_STEP_EVAL, FAIL_STEP_MERKLE, FAIL_FINAL = range(8)


def reduce_extension(base, terms):
    """goldilocks_extension_chip.rs:331-342: Horner from the LAST term, acc = acc * base + term"""
    acc = (0, 0)
    for t in reversed(terms):
        acc = gl.e_add(gl.e_mul(acc, base), t)
    return acc


def pow_ok(pow_response, bits):
    """fri_verify_proof_of_work (:364-376): the top `bits` bits of the 64-bit decomposition are zero"""
    return all(((pow_response >> (63 - i)) & 1) == 0 for i in range(bits))


def x_from_index(x_index, lde_bits):
    """:262-264 with x_from_subgroup :152-166: offset * omega^(x_index bits reversed), offset = 7 (plonk_verifier_chip.rs:225-227)"""
    return gl.GENERATOR * pow(gl.root_of_unity(lde_bits), gl.bitrev(x_index, lde_bits), P) % P


def batch_initial_polynomials(params, alpha, x, initial, reduced_openings, points):
    """:112-149.  -> (sum, zero_denominator_seen).  Batch 0 = every polynomial of every oracle in order, at zeta; batch 1 = the
    first num_zs polynomials of oracle 2, at g * zeta (types/fri.rs:50-72); a salted leaf carries 4 extra limbs at its END
    (types/assigned.rs:57-71), so polynomial indices never shift."""
    batches = [[initial[k][0][j] for k in range(4) for j in range(params.oracle_num_polys[k])],
               [initial[2][0][j] for j in range(params.num_zs)]]
    total, zero = (0, 0), False
    for evals, ro, point in zip(batches, reduced_openings, points):
        reduced_evals = reduce_extension(alpha, [gl.e(v) for v in evals])       # :139-140
        numerator = gl.e_sub(reduced_evals, ro)                                 # :141-142
        denominator = gl.e_sub(gl.e(x), point)                                  # :143
        total = gl.e_mul(total, gl.e_pow(alpha, len(evals)))                    # :144 shift
        if denominator == (0, 0):
            zero = True                                                         # div_extension of zero cannot be witnessed
            continue
        total = gl.e_add(gl.e_mul(numerator, gl.e_inv(denominator)), total)     # :145-146
    return total, zero


def fold(x, x_index_within_coset, arity_bits, evals, beta):
    """plonky2 compute_evaluation == next_eval (:168-226) for arity 2.  evals[j] is the value at the j-th point of the coset in
    LEAF order; after reverse_index_bits the i-th entry sits at coset_start * g^i, g = primitive 2^arity_bits-th root."""
    arity = 1 << arity_bits
    g = gl.root_of_unity(arity_bits)
    ev = [evals[gl.bitrev(i, arity_bits)] for i in range(arity)]               # reverse_index_bits_in_place (:188-189)
    rev = gl.bitrev(x_index_within_coset, arity_bits)
    coset_start = x * pow(gl.inv(g), rev, P) % P                                # :191-200
    pts = [coset_start * pow(g, i, P) % P for i in range(arity)]
    # Lagrange interpolation through (pts[i], ev[i]), evaluated at beta
    acc = (0, 0)
    for i in range(arity):
        num, den = (1, 0), 1
        for j in range(arity):
            if j != i:
                num = gl.e_mul(num, gl.e_sub(beta, gl.e(pts[j])))
                den = den * (pts[i] - pts[j]) % P
        acc = gl.e_add(acc, gl.e_mul(ev[i], gl.e_scale(num, gl.inv(den))))
    return acc


def fold_arity2_reference(x, bit, evals, beta):
    """the literal two-point formula of next_eval (:212-224): a1 + (beta - a0)(b1 - a1)/(b0 - a0)"""
    g = P - 1
    coset_start = x * pow(gl.inv(g), bit, P) % P
    a0, a1, b0, b1 = gl.e(coset_start), evals[0], gl.e(coset_start * g), evals[1]
    num = gl.e_mul(gl.e_sub(beta, a0), gl.e_sub(b1, a1))
    return gl.e_add(gl.e_mul(num, gl.e_inv(gl.e_sub(b0, a0))), a1)


def _canonical(words):
    return all(0 <= int(w) < P for w in words)


def check_round_algebra(params, ch, points, reduced_openings, fri_proof, q):
    """the non-Merkle part of check_consistency (:228-327) for query round q -> list of (order_key, code) failures.
    order keys: 5 = DEEP quotient, 8 + 3 i (+0 eval, +1 fold division, +2 Merkle) for step i, 8 + 3 S = final."""
    qr = fri_proof.query_round_proofs[q]
    lde_bits, S = params.lde_bits(), len(params.reduction_arity_bits)
    x_index = ch["fri_query_indices"][q] & ((1 << lde_bits) - 1)              # :245-250
    x = x_from_index(x_index, lde_bits)
    fails = []
    prev, zero = batch_initial_polynomials(params, ch["fri_alpha"], x, qr.initial, reduced_openings, points)   # :266-273
    if zero:
        fails.append((5, FAIL_ZERO_DENOM))
    for i, ab in enumerate(params.reduction_arity_bits):                       # :275-316
        evals = qr.steps[i].evals
        within = x_index & ((1 << ab) - 1)                                     # :279-282
        if tuple(evals[within]) != tuple(prev):                                # :285-292
            fails.append((8 + 3 * i, FAIL_STEP_EVAL))
        prev = fold(x, within, ab, evals, ch["fri_betas"][i])                 # :294-301
        x = pow(x, 1 << ab, P)                                                 # :313
        x_index >>= ab                                                         # :315
    final_eval = reduce_extension(gl.e(x), fri_proof.final_poly)               # :317-323
    if tuple(final_eval) != tuple(prev):                                       # :324
        fails.append((8 + 3 * S, FAIL_FINAL))
    return fails


def merkle_items(params, caps, ch, fri_proof, q):
    """the Merkle checks of round q as (order_key, code, verify_to_cap arguments)"""
    qr = fri_proof.query_round_proofs[q]
    lde_bits = params.lde_bits()
    x_index = ch["fri_query_indices"][q] & ((1 << lde_bits) - 1)
    bits = [(x_index >> i) & 1 for i in range(lde_bits)]
    cap_index = x_index >> (lde_bits - params.cap_height)                      # :72-82: the top cap_height bits
    items = []
    for k in range(4):                                                         # :254-260
        evals, sibs = qr.initial[k]
        items.append((1 + k, FAIL_INIT_MERKLE, (evals, bits, cap_index, caps[k], sibs)))
    for i, ab in enumerate(params.reduction_arity_bits):
        bits = bits[ab:]                                                       # coset_index_bits (:279)
        st = qr.steps[i]
        leaf = [v for ext in st.evals for v in ext]                            # :305
        items.append((8 + 3 * i + 2, FAIL_STEP_MERKLE, (leaf, bits, cap_index, fri_proof.commit_phase_merkle_caps[i], st.siblings)))
    return items


def round_words(qr):
    for evals, sibs in qr.initial:
        yield from evals
        for h in sibs:
            yield from h
    for st in qr.steps:
        for ext in st.evals:
            yield from ext
        for h in st.siblings:
            yield from h


def verify_fri_proof(params, initial_merkle_caps, challenges, points, openings_batches, fri_proof, batch_hashing=True):
    """-> (accept, code, query).  initial_merkle_caps: [constants_sigmas (verifier key), wires, zs_partial_products, quotient]
    (plonk_verifier_chip.rs:212-217); points = (zeta, g * zeta); openings_batches = to_fri_openings()."""
    kind = params.hash_kind
    header = [v for c in initial_merkle_caps for h in c for v in h] + [v for c in fri_proof.commit_phase_merkle_caps for h in c for v in h]
    header += [v for b in openings_batches for ext in b for v in ext] + [v for ext in fri_proof.final_poly for v in ext]
    header += [fri_proof.pow_witness, *challenges["fri_alpha"], *[v for b in challenges["fri_betas"] for v in b],
               challenges["fri_pow_response"], *challenges["fri_query_indices"], *points[0], *points[1]]
    if not _canonical(header):
        return False, FAIL_NONCANONICAL, 0
    if not pow_ok(challenges["fri_pow_response"], params.proof_of_work_bits):          # :339-344
        return False, FAIL_POW, 0
    reduced_openings = [reduce_extension(challenges["fri_alpha"], b) for b in openings_batches]   # :347, :58-70
    Q = len(fri_proof.query_round_proofs)
    items = [merkle_items(params, initial_merkle_caps, challenges, fri_proof, q) for q in range(Q)]
    canon = [_canonical(round_words(fri_proof.query_round_proofs[q])) for q in range(Q)]
    ok = {}
    if batch_hashing:
        # all Merkle checks of all canonical rounds in one vectorised pass (same semantics as verify_to_cap)
        todo = [(q, j) for q in range(Q) if canon[q] for j in range(len(items[q]))]
        for key, good in zip(todo, merkle.verify_to_cap_batch([items[q][j][2] for q, j in todo], kind)):
            ok[key] = good
    for q in range(Q):                                                                  # :348-361
        fails = []
        if not canon[q]:
            return False, FAIL_NONCANONICAL, q
        for j, (key, code, args) in enumerate(items[q]):
            if not (ok[(q, j)] if batch_hashing else merkle.verify_to_cap(*args, kind)):
                fails.append((key, code))
        fails += check_round_algebra(params, challenges, points, reduced_openings, fri_proof, q)
        if fails:
            return False, min(fails)[1], q
    return True, OK, 0


def verify_record(params, common, rec, batch_hashing=True):
    """the same verdict for a flat record with its challenge fields filled in (the input of sv_fri_verify_batch)"""
    from . import proof as pf
    proof, cs_cap, ch, zn = pf.from_record(params, common, rec)
    caps = [cs_cap, proof.wires_cap, proof.plonk_zs_partial_products_cap, proof.quotient_polys_cap]
    return verify_fri_proof(params, caps, ch, (ch["plonk_zeta"], zn), fri_openings(proof.openings), proof.opening_proof, batch_hashing)
