"""The prover side of the FRI opening protocol -- TEST INFRASTRUCTURE, pure Python/numpy.

Follows plonky2's prover (PolynomialBatch::from_coeffs, prove_openings, fri_committed_trees, fri_proof_of_work,
fri_prover_query_rounds; the dependency is not on disk, its algorithm is restated) with the conventions the reference's verifier
implies: LDE on the coset 7 * <omega_N>; leaf i of an oracle tree = the values at 7 * omega_N^bitrev(i)
(chip/fri_chip.rs:152-166,262-264), plus 4 salt limbs when blinded (types/assigned.rs:57-71); batch 0 = all polynomials at
zeta, batch 1 = the Z polynomials at g * zeta (types/fri.rs:50-72); final polynomial of the DEEP step =
sum over batches of alpha^(later batch sizes) * (sum_i alpha^i p_i(X) - sum_i alpha^i p_i(z)) / (X - z)  (chip/fri_chip.rs:112-149 read
backwards); commit phase: values in bit-reversed order, chunks of 2^arity_bits per leaf, coefficients folded as
c'_j = sum_i beta^i c_(j arity + i); PoW witness = the smallest w whose response has proof_of_work_bits leading zeros."""
import numpy as np

from . import gl
from . import poseidon as ps
from .challenger import Challenger, zeta_next
from .gl import P
from .merkle import MerkleTree
from .proof import FriProof, Proof, QueryRound, QueryStep


def bitrev_perm(bits):
    n = 1 << bits
    return np.array([gl.bitrev(i, bits) for i in range(n)], dtype=np.int64)


def ntt(a, bits):
    """a: (m, 2^bits) uint64, coefficients -> values at omega^i (natural order), omega = root_of_unity(bits)"""
    a = gl.varr(a)[:, bitrev_perm(bits)].copy()
    for s in range(bits):
        half = 1 << s
        w = gl.root_of_unity(s + 1)
        tw = np.array([pow(w, j, P) for j in range(half)], dtype=np.uint64)
        a = a.reshape(a.shape[0], -1, 2 * half)
        lo, hi = a[:, :, :half], gl.vmul(a[:, :, half:], tw[None, None, :])
        a = np.concatenate([gl.vadd(lo, hi), gl.vsub(lo, hi)], axis=2)
    return a.reshape(a.shape[0], -1)


def coset_lde(coeffs, bits_out, shift):
    """(m, n) coefficients -> (m, 2^bits_out) values at shift * omega^i"""
    coeffs = gl.varr(coeffs)
    m, n = coeffs.shape
    N = 1 << bits_out
    sp = np.array([pow(shift, i, P) for i in range(n)], dtype=np.uint64)
    a = np.zeros((m, N), dtype=np.uint64)
    a[:, :n] = gl.vmul(coeffs, sp[None, :])
    return ntt(a, bits_out)


def poly_eval_ext(coeffs, z):
    """base-field coefficients at an F_p^2 point"""
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = gl.e_add(gl.e_mul(acc, z), (int(c), 0))
    return acc


def divide_by_linear(coeffs, z):
    """(a(X) - a(z)) / (X - z) for F_p^2 coefficients: synthetic division, remainder dropped"""
    b = [(0, 0)] * (len(coeffs) - 1)
    carry = (0, 0)
    for k in range(len(coeffs) - 1, 0, -1):
        carry = gl.e_add(coeffs[k], gl.e_mul(carry, z))
        b[k - 1] = carry
    return b


class Prover:
    def __init__(self, params, circuit_digest, public_inputs, seed):
        self.p = params
        self.kind = params.hash_kind
        self.rng = np.random.default_rng(seed)
        self.circuit_digest = [int(v) for v in circuit_digest]
        self.public_inputs = [int(v) for v in public_inputs]
        self.pi_hash = ps.hash_no_pad(self.public_inputs, ps.HASH_G)    # PublicInputsHasherChip: always Poseidon-Goldilocks
        self.polys, self.trees = [None] * 4, [None] * 4
        self.ch = Challenger(self.kind)
        self.n, self.N = 1 << params.degree_bits, 1 << params.lde_bits()
        self.perm = bitrev_perm(params.lde_bits())

    def rand_field(self, shape):
        v = self.rng.integers(0, P, size=shape, dtype=np.uint64)
        return v

    def commit(self, k, polys):
        """oracle k := the polynomials `polys` ((m, n) coefficients)"""
        polys = gl.varr(polys).reshape(-1, self.n)
        assert polys.shape[0] == self.p.oracle_num_polys[k]
        vals = coset_lde(polys, self.p.lde_bits(), gl.GENERATOR)
        leaves = vals[:, self.perm].T
        if self.p.hiding and self.p.oracle_blinding[k]:
            leaves = np.concatenate([leaves, self.rand_field((self.N, 4))], axis=1)
        self.polys[k] = polys
        self.trees[k] = MerkleTree(leaves, self.p.cap_height, self.kind)
        return self.trees[k].cap()

    # the transcript, in the order of get_challenges (chip/plonk/plonk_verifier_chip.rs:55-154)
    def draw_betas_gammas(self, nch):
        self.ch.observe_many(self.circuit_digest)
        self.ch.observe_many(self.pi_hash)
        self.ch.observe_cap(self.trees[1].cap())
        return self.ch.squeeze(nch), self.ch.squeeze(nch)

    def draw_alphas(self, nch):
        self.ch.observe_cap(self.trees[2].cap())
        return self.ch.squeeze(nch)

    def draw_zeta(self):
        self.ch.observe_cap(self.trees[3].cap())
        self.zeta = tuple(self.ch.squeeze(2))
        return self.zeta

    def _grind(self):
        """smallest pow_witness whose response (observe witness, squeeze 1) has the required leading zero bits"""
        bits = self.p.proof_of_work_bits
        buf = list(self.ch.absorbing)
        base = self.ch.clone()
        base.absorbing = []
        # absorb every chunk that does not contain the witness; the witness is the last word of the last chunk
        m = (len(buf) + 1) % 8 or 8                       # length of the chunk that holds the witness
        head = buf[:len(buf) + 1 - m]
        st = list(base.state)
        for off in range(0, len(head), 8):
            st[:8] = head[off:off + 8]
            st = ps.permute(st, self.kind)
        tail = buf[len(buf) + 1 - m:]
        st[:len(tail)] = tail
        w0, step = 0, 1 << 12
        while True:
            cand = np.arange(w0, w0 + step, dtype=np.uint64)
            states = np.tile(np.array(st, dtype=np.uint64), (step, 1))
            states[:, m - 1] = cand
            resp = ps.permute_batch(states, self.kind)[:, 7]      # squeeze pops the LAST rate word
            good = np.nonzero((resp >> np.uint64(64 - bits)) == 0)[0] if bits else np.array([0])
            if len(good):
                return int(cand[good[0]])
            w0 += step

    def open_and_prove(self, common):
        p, kind = self.p, self.kind
        zeta = self.zeta
        gz = zeta_next(zeta, p.degree_bits)
        all_polys = np.concatenate(self.polys, axis=0)
        z_polys = self.polys[2][:p.num_zs]
        open0 = [poly_eval_ext(c, zeta) for c in all_polys]
        open1 = [poly_eval_ext(c, gz) for c in z_polys]
        lens = common.opening_lens()
        openings, at = {}, 0
        for f in ("constants", "plonk_sigmas", "wires", "plonk_zs", "partial_products", "quotient_polys"):
            openings[f] = open0[at:at + lens[f]]
            at += lens[f]
        assert at == len(open0)
        openings["plonk_zs_next"] = open1
        for ext in open0 + open1:
            self.ch.observe_ext(ext)
        alpha = tuple(self.ch.squeeze(2))

        def batch_poly(polys):
            """sum_i alpha^i polys[i] as F_p^2 coefficients"""
            c0, c1, ap = np.zeros(self.n, dtype=np.uint64), np.zeros(self.n, dtype=np.uint64), (1, 0)
            for row in polys:
                c0 = gl.vadd(c0, gl.vmul(row, np.uint64(ap[0])))
                c1 = gl.vadd(c1, gl.vmul(row, np.uint64(ap[1])))
                ap = gl.e_mul(ap, alpha)
            return [(int(a), int(b)) for a, b in zip(c0, c1)]
        q0 = divide_by_linear(batch_poly(all_polys), zeta)
        q1 = divide_by_linear(batch_poly(z_polys), gz)
        an1 = gl.e_pow(alpha, len(z_polys))
        coeffs = [gl.e_add(gl.e_mul(a, an1), b) for a, b in zip(q0, q1)] + [(0, 0)]
        # ---- commit phase ----
        caps, trees, layers, betas = [], [], [], []
        shift, bits = gl.GENERATOR, p.lde_bits()
        for ab in p.reduction_arity_bits:
            arity = 1 << ab
            c = np.array(coeffs, dtype=np.uint64)
            vals = coset_lde(np.stack([c[:, 0], c[:, 1]]), bits, shift)            # (2, 2^bits): both limbs
            vals = vals[:, bitrev_perm(bits)].T                                    # (2^bits, 2) in leaf order
            leaves = vals.reshape(-1, 2 * arity)                                   # flatten(chunk of arity values)
            tree = MerkleTree(leaves, p.cap_height, kind)
            self.ch.observe_cap(tree.cap())
            beta = tuple(self.ch.squeeze(2))
            folded = []
            for j in range(len(coeffs) // arity):
                acc = (0, 0)
                for ccc in reversed(coeffs[j * arity:(j + 1) * arity]):           # reduce_with_powers(chunk, beta)
                    acc = gl.e_add(gl.e_mul(acc, beta), ccc)
                folded.append(acc)
            coeffs = folded
            caps.append(tree.cap()); trees.append(tree); layers.append(leaves); betas.append(beta)
            shift, bits = pow(shift, arity, P), bits - ab
        final_poly = coeffs[:p.final_poly_len()]
        assert all(c == (0, 0) for c in coeffs[p.final_poly_len():]) and len(final_poly) == p.final_poly_len()
        for ext in final_poly:
            self.ch.observe_ext(ext)
        pow_witness = self._grind()
        self.ch.observe(pow_witness)
        pow_response = self.ch.squeeze(1)[0]
        assert p.proof_of_work_bits == 0 or pow_response >> (64 - p.proof_of_work_bits) == 0
        indices = self.ch.squeeze(p.num_query_rounds)
        # ---- query rounds ----
        rounds = []
        for idx_fe in indices:
            idx = idx_fe & (self.N - 1)
            initial = [(self.trees[k].leaf(idx), self.trees[k].prove(idx)) for k in range(4)]
            steps, cur = [], idx
            for i, ab in enumerate(p.reduction_arity_bits):
                coset = cur >> ab
                row = [int(v) for v in layers[i][coset]]
                steps.append(QueryStep([(row[2 * t], row[2 * t + 1]) for t in range(1 << ab)], trees[i].prove(coset)))
                cur = coset
            rounds.append(QueryRound(initial, steps))
        fri = FriProof(caps, rounds, final_poly, pow_witness)
        proof = Proof(self.trees[1].cap(), self.trees[2].cap(), self.trees[3].cap(), openings, fri, self.public_inputs)
        self.challenges = {"plonk_zeta": zeta, "fri_alpha": alpha, "fri_betas": betas, "fri_pow_response": pow_response,
                           "fri_query_indices": indices}
        return proof


def prove_random(params, common, seed, num_public_inputs=None):
    """a proof whose four oracles commit to uniformly random polynomials (valid for the FRI verifier; it does not satisfy any
    plonk identity) -> (Proof, constants_sigmas_cap of the 'verifier key', circuit_digest)"""
    rng = np.random.default_rng(seed ^ 0xC1C0)
    npi = common.num_public_inputs if num_public_inputs is None else num_public_inputs
    cd = [int(v) for v in rng.integers(0, P, size=4, dtype=np.uint64)]
    pis = [int(v) for v in rng.integers(0, P, size=npi, dtype=np.uint64)]
    pr = Prover(params, cd, pis, seed)
    for k in range(4):
        pr.commit(k, pr.rand_field((params.oracle_num_polys[k], 1 << params.degree_bits)))
        if k == 1:
            pr.draw_betas_gammas(common.num_challenges)
        elif k == 2:
            pr.draw_alphas(common.num_challenges)
    pr.draw_zeta()
    proof = pr.open_and_prove(common)
    return proof, pr.trees[0].cap(), cd, pr
