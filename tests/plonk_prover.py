"""TEST INFRASTRUCTURE: a miniature plonky2-style prover in pure Python (big integers), just enough to produce openings
that satisfy the vanishing-polynomial identity the verifier checks (reference: chip/plonk/plonk_verifier_chip.rs:156-210,
chip/plonk/vanishing_poly.rs) for a small circuit over any subset of the reference's gates (chip/plonk/gates/*.rs).

It follows the protocol, not the verifier's code: witness rows -> copy-constraint permutation -> sigma polynomials ->
grand products Z with partial products -> vanishing polynomial on a coset -> division by Z_H -> quotient chunks ->
openings at a random zeta in F_p^2.  It shares no code with the product (stark-verifier_b200/csrc/plonk_check.hpp) or the
oracle (oracle/plonk.c); those restate the reference's verifier, and accepting this prover's output is the evidence
that the restatements are right.  Sizes are tiny (2^4 rows), so everything is naive O(n^2)."""
import numpy as np

P = 0xFFFFFFFF00000001
UNUSED_SELECTOR = 0xFFFFFFFF
GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_ARITHMETIC = 0, 1, 2, 3
GATE_ARITHMETIC_EXT, GATE_MUL_EXT, GATE_BASE_SUM, GATE_REDUCING, GATE_REDUCING_EXT = 4, 5, 6, 7, 8
GATE_RANDOM_ACCESS, GATE_POSEIDON_MDS, GATE_POSEIDON = 9, 10, 11


def inv(a):
    return pow(a, P - 2, P)


# F_p^2 = F_p[X]/(X^2 - 7), elements as (c0, c1)
def e_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def e_mul(a, b): return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def e_scale(a, s): return (a[0] * s % P, a[1] * s % P)


def poly_eval_ext(coeffs, z):
    acc = (0, 0)
    for c in reversed(coeffs):
        acc = e_add(e_mul(acc, z), (c, 0))
    return acc


def poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % P
    return acc


# D = 2 wire elements over the base field: pairs (x0, x1) = x0 + x1 X, X^2 = 7 -- the witness-side meaning of the
# extension gates; their constraints are these formulas, limb by limb
def x_mul(a, b): return ((a[0] * b[0] + 7 * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def x_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def x_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def x_scale(s, a): return (s * a[0] % P, s * a[1] % P)


def _poseidon_tables():
    """The Poseidon-Goldilocks tables (reference: chip/plonk/gates/poseidon.rs:26-322), from tests/pyref/constants.json
    (parsed from the reference's .rs by tools/gen_pyref_constants.py)."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyref", "constants.json")))
    out = {"ALL_ROUND_CONSTANTS": [int(x, 16) for x in d["g_round_constants"]], "MDS_MATRIX_CIRC": d["g_mds_circ"], "MDS_MATRIX_DIAG": d["g_mds_diag"]}
    for name in ("FAST_PARTIAL_FIRST_ROUND_CONSTANT", "FAST_PARTIAL_ROUND_CONSTANTS", "FAST_PARTIAL_ROUND_VS", "FAST_PARTIAL_ROUND_W_HATS",
                 "FAST_PARTIAL_ROUND_INITIAL_MATRIX"):
        out[name] = [int(x, 16) for x in d["g_" + name.lower()]]
    return out


PT = _poseidon_tables()
PG_SWAP, PG_DELTA, PG_FULL0, PG_PARTIAL, PG_FULL1 = 24, 25, 29, 65, 87


def _mds(st):
    return [(sum(PT["MDS_MATRIX_CIRC"][i] * st[(i + r) % 12] for i in range(12)) + PT["MDS_MATRIX_DIAG"][r] * st[r]) % P for r in range(12)]


def poseidon_gate_trace(w, check):
    """The PoseidonGate's view of the permutation (fast form, every S-box input a wire).  check=False: fill the wires of
    row `w` from its inputs (wires 0..12) and swap bit (wire 24); check=True: return the 123 constraint values."""
    cs = []
    swap = w[PG_SWAP]
    cs.append((swap * swap - swap) % P)
    for i in range(4):
        d = swap * (w[i + 4] - w[i]) % P
        if check:
            cs.append((d - w[PG_DELTA + i]) % P)
        else:
            w[PG_DELTA + i] = d
    st = [0] * 12
    for i in range(4):
        st[i] = (w[i] + w[PG_DELTA + i]) % P
        st[i + 4] = (w[i + 4] - w[PG_DELTA + i]) % P
    st[8:12] = w[8:12]

    def wire(at, value):            # a wire that must equal `value`: constrain it (check) or assign it (fill)
        if check:
            cs.append((value - w[at]) % P)
        else:
            w[at] = value
        return w[at]

    rc = 0
    for r in range(4):
        st = [(st[i] + PT["ALL_ROUND_CONSTANTS"][i + 12 * rc]) % P for i in range(12)]
        if r:
            st = [wire(PG_FULL0 + 12 * (r - 1) + i, st[i]) for i in range(12)]
        st = _mds([pow(x, 7, P) for x in st])
        rc += 1
    st = [(st[i] + PT["FAST_PARTIAL_FIRST_ROUND_CONSTANT"][i]) % P for i in range(12)]
    init = PT["FAST_PARTIAL_ROUND_INITIAL_MATRIX"]
    st = [st[0]] + [sum(init[(r - 1) * 11 + (c - 1)] * st[r] for r in range(1, 12)) % P for c in range(1, 12)]
    for r in range(22):
        s0 = pow(wire(PG_PARTIAL + r, st[0]), 7, P)
        if r != 21:
            s0 = (s0 + PT["FAST_PARTIAL_ROUND_CONSTANTS"][r]) % P
        d = (25 * s0 + sum(PT["FAST_PARTIAL_ROUND_W_HATS"][r * 11 + i - 1] * st[i] for i in range(1, 12))) % P
        st = [d] + [(PT["FAST_PARTIAL_ROUND_VS"][r * 11 + i - 1] * s0 + st[i]) % P for i in range(1, 12)]
    rc += 22
    for r in range(4):
        st = [(st[i] + PT["ALL_ROUND_CONSTANTS"][i + 12 * rc]) % P for i in range(12)]
        st = [wire(PG_FULL1 + 12 * r + i, st[i]) for i in range(12)]
        st = _mds([pow(x, 7, P) for x in st])
        rc += 1
    for i in range(12):
        wire(12 + i, st[i])
    return cs


def gate_id(kind, param):
    """plonky2's `gate.id()` string of a gate, the key CustomGateRef::from matches on (chip/plonk/gates/mod.rs:138-196)"""
    ph = "PhantomData<plonky2_field::goldilocks_field::GoldilocksField>"
    if kind == GATE_NOOP: return "NoopGate"
    if kind == GATE_PUBLIC_INPUT: return "PublicInputGate"
    if kind == GATE_CONSTANT: return "ConstantGate { num_consts: %d }" % param
    if kind == GATE_ARITHMETIC: return "ArithmeticGate { num_ops: %d }" % param
    if kind == GATE_ARITHMETIC_EXT: return "ArithmeticExtensionGate { num_ops: %d }" % param
    if kind == GATE_MUL_EXT: return "MulExtensionGate { num_ops: %d }" % param
    if kind == GATE_BASE_SUM: return "BaseSumGate { num_limbs: %d } + Base: 2" % param
    if kind == GATE_REDUCING: return "ReducingGate { num_coeffs: %d }" % param
    if kind == GATE_REDUCING_EXT: return "ReducingExtensionGate { num_coeffs: %d }" % param
    if kind == GATE_RANDOM_ACCESS:
        return "RandomAccessGate { bits: %d, num_copies: %d, num_extra_constants: %d, _phantom: %s }<D=2>" % (param[0], param[1], param[2], ph)
    if kind == GATE_POSEIDON_MDS: return "PoseidonMdsGate(%s)<WIDTH=12>" % ph
    if kind == GATE_POSEIDON: return "PoseidonGate(%s)<WIDTH=12>" % ph
    raise ValueError(kind)


def _ra(param):
    bits, copies, extra = param
    return bits, copies, extra, 1 << bits, (2 + (1 << bits)) * copies + extra


def gate_dims(kind, param):
    """(wires used, constraints, polynomial degree)"""
    if kind == GATE_RANDOM_ACCESS:
        bits, copies, extra, vec, routed = _ra(param)
        return (routed + bits * copies, copies * (bits + 2) + extra, bits + 1)
    if kind == GATE_POSEIDON_MDS:
        return (48, 24, 1)
    if kind == GATE_POSEIDON:
        return (135, 123, 7)
    return {GATE_NOOP: (0, 0, 0), GATE_CONSTANT: (param, param, 1), GATE_PUBLIC_INPUT: (4, 4, 1),
            GATE_ARITHMETIC: (4 * param, param, 3), GATE_ARITHMETIC_EXT: (8 * param, 2 * param, 3),
            GATE_MUL_EXT: (6 * param, 2 * param, 3), GATE_BASE_SUM: (1 + param, 1 + param, 2),
            GATE_REDUCING: (3 * param + 4, 2 * param, 2), GATE_REDUCING_EXT: (4 * param + 4, 2 * param, 2)}[kind]


class Circuit:
    """Configuration of the toy circuit.  gates: list of (kind, param) in CommonData.gates order; groups: selector
    groups as (lo, hi) ranges over gate indices."""

    def __init__(self, degree_bits=4, num_wires=13, num_routed_wires=12, num_challenges=2, quotient_degree_factor=8,
                 groups=((0, 4),), extra_gates=()):
        self.degree_bits = degree_bits
        self.n = 1 << degree_bits
        self.num_wires, self.num_routed_wires, self.num_challenges = num_wires, num_routed_wires, num_challenges
        self.qdf = quotient_degree_factor
        self.num_ops = num_routed_wires // 4                       # ArithmeticGate::new_from_config
        self.gates = [(GATE_NOOP, 0), (GATE_CONSTANT, 2), (GATE_PUBLIC_INPUT, 0), (GATE_ARITHMETIC, self.num_ops)]
        self.gates += list(extra_gates)
        assert all(gate_dims(k, p)[0] <= num_wires for k, p in self.gates)
        self.groups = list(groups)
        self.num_selectors = len(self.groups)
        self.num_constants = self.num_selectors + 2
        self.num_partial_products = (num_routed_wires + self.qdf - 1) // self.qdf - 1
        self.num_gate_constraints = max(gate_dims(k, p)[1] for k, p in self.gates)
        self.k_is = [pow(7, j, P) for j in range(num_routed_wires)]   # distinct cosets of the trace subgroup
        self.g = pow(7, (P - 1) >> degree_bits, P)

    def selector_index(self, gate):
        return next(s for s, (lo, hi) in enumerate(self.groups) if lo <= gate < hi)


def _constraints(C, consts, wires, pi_hash):
    """Unfiltered-then-filtered gate constraints at one point, base field (the prover's own statement of the gates)."""
    out = [0] * C.num_gate_constraints
    gc = consts[C.num_selectors:]
    for i, (kind, param) in enumerate(C.gates):
        s = C.selector_index(i)
        lo, hi = C.groups[s]
        f = 1
        for k in range(lo, hi):
            if k != i:
                f = f * (k - consts[s]) % P
        if C.num_selectors > 1:
            f = f * (UNUSED_SELECTOR - consts[s]) % P
        if kind == GATE_CONSTANT:
            cs = [(gc[k] - wires[k]) % P for k in range(param)]
        elif kind == GATE_PUBLIC_INPUT:
            cs = [(wires[k] - pi_hash[k]) % P for k in range(4)]
        elif kind == GATE_ARITHMETIC:
            cs = [(wires[4 * k + 3] - (gc[0] * wires[4 * k] * wires[4 * k + 1] + gc[1] * wires[4 * k + 2])) % P for k in range(param)]
        elif kind in (GATE_ARITHMETIC_EXT, GATE_MUL_EXT):
            stride = 8 if kind == GATE_ARITHMETIC_EXT else 6
            cs = []
            for k in range(param):
                el = lambda j: (wires[stride * k + 2 * j], wires[stride * k + 2 * j + 1])
                want = x_scale(gc[0], x_mul(el(0), el(1)))
                if kind == GATE_ARITHMETIC_EXT:
                    want = x_add(want, x_scale(gc[1], el(2)))
                cs.extend(x_sub(el(stride // 2 - 1), want))
        elif kind == GATE_BASE_SUM:
            limbs = wires[1:1 + param]
            cs = [(sum(l << k for k, l in enumerate(limbs)) - wires[0]) % P] + [l * (l - 1) % P for l in limbs]
        elif kind in (GATE_REDUCING, GATE_REDUCING_EXT):
            ext = kind == GATE_REDUCING_EXT
            alpha, acc = (wires[2], wires[3]), (wires[4], wires[5])
            start_accs = 6 + (2 * param if ext else param)
            cs = []
            for k in range(param):
                coeff = (wires[6 + 2 * k], wires[7 + 2 * k]) if ext else (wires[6 + k], 0)
                at = 0 if k == param - 1 else start_accs + 2 * k
                acc_k = (wires[at], wires[at + 1])
                cs.extend(x_sub(x_add(x_mul(acc, alpha), coeff), acc_k))
                acc = acc_k
        elif kind == GATE_RANDOM_ACCESS:
            bits, copies, extra, vec, routed = _ra(param)
            cs = []
            for copy in range(copies):
                base = (2 + vec) * copy
                bv = [wires[routed + copy * bits + i] for i in range(bits)]
                cs += [b * (b - 1) % P for b in bv]
                cs.append((sum(b << i for i, b in enumerate(bv)) - wires[base]) % P)
                items = list(wires[base + 2: base + 2 + vec])
                for b in bv:
                    items = [(x + b * (y - x)) % P for x, y in zip(items[0::2], items[1::2])]
                cs.append((items[0] - wires[base + 1]) % P)
            cs += [(gc[i] - wires[(2 + vec) * copies + i]) % P for i in range(extra)]
        elif kind == GATE_POSEIDON_MDS:
            cs = []
            for limb_pairs in [[(wires[2 * i], wires[2 * i + 1]) for i in range(12)]]:
                lo, hi = _mds([x[0] for x in limb_pairs]), _mds([x[1] for x in limb_pairs])
                for r in range(12):
                    cs += [(wires[2 * (12 + r)] - lo[r]) % P, (wires[2 * (12 + r) + 1] - hi[r]) % P]
        elif kind == GATE_POSEIDON:
            cs = poseidon_gate_trace(list(wires), check=True)
        else:
            cs = []
        assert len(cs) == gate_dims(kind, param)[1]
        for k, c in enumerate(cs):
            out[k] = (out[k] + f * c) % P
    return out


def prove(C, seed, pi_hash, draw_betas_gammas=None, draw_alphas=None, draw_zeta=None):
    """-> dict(open0=[...Fp2], open1=[...], betas, gammas, alphas, zeta, polys=...).  The three optional callbacks let a
    caller derive the challenges from a transcript over commitments (tests/full_prover.py): draw_betas_gammas(wire_polys)
    -> (betas, gammas); draw_alphas(z_polys, pp_polys) -> alphas; draw_zeta(quotient_chunks) -> (c0, c1).  Default: random."""
    rng = np.random.default_rng(seed)
    rnd = lambda: int(rng.integers(0, P, dtype=np.uint64))
    n, nr, nw, nch, qdf, npp = C.n, C.num_routed_wires, C.num_wires, C.num_challenges, C.qdf, C.num_partial_products
    g = C.g
    # ---- witness -------------------------------------------------------------------------------------
    row_gate = [GATE_NOOP] * n                     # index into C.gates (== kind for the first four)
    row_gate[0] = GATE_PUBLIC_INPUT
    row_gate[1] = row_gate[2] = GATE_CONSTANT
    extra = list(range(4, len(C.gates)))
    assert 3 + len(extra) < n - 2, "no row left for the arithmetic gate"
    for r in range(3, n - 2):
        row_gate[r] = extra[r - 3] if r - 3 < len(extra) else GATE_ARITHMETIC
    assert set(row_gate) == set(range(len(C.gates))), "every gate of the circuit must be used by some row"
    wires = [[rnd() for _ in range(nw)] for _ in range(n)]
    consts = [[0] * C.num_constants for _ in range(n)]
    parent = {}

    def find(c):
        while parent.setdefault(c, c) != c:
            parent[c] = parent[parent[c]]
            c = parent[c]
        return c

    filled = []   # routed cells whose value is final and may be copied
    for r in range(n):
        gate = row_gate[r]
        kind, param = C.gates[gate]
        for s, (lo, hi) in enumerate(C.groups):
            consts[r][s] = gate if lo <= gate < hi else UNUSED_SELECTOR
        gc0, gc1 = rnd(), rnd()
        consts[r][C.num_selectors], consts[r][C.num_selectors + 1] = gc0, gc1
        w = wires[r]
        if kind == GATE_PUBLIC_INPUT:
            wires[r][0:4] = list(pi_hash)
        elif kind == GATE_CONSTANT:
            wires[r][0], wires[r][1] = gc0, gc1
        elif kind in (GATE_ARITHMETIC_EXT, GATE_MUL_EXT):
            stride = 8 if kind == GATE_ARITHMETIC_EXT else 6
            for k in range(param):
                el = lambda j: (w[stride * k + 2 * j], w[stride * k + 2 * j + 1])
                out = x_scale(gc0, x_mul(el(0), el(1)))
                if kind == GATE_ARITHMETIC_EXT:
                    out = x_add(out, x_scale(gc1, el(2)))
                w[stride * k + stride - 2], w[stride * k + stride - 1] = out
        elif kind == GATE_RANDOM_ACCESS:
            bits, copies, extra, vec, routed = _ra(param)
            for copy in range(copies):
                base = (2 + vec) * copy
                idx = int(rng.integers(0, vec))
                w[base], w[base + 1] = idx, w[base + 2 + idx]
                for i in range(bits):
                    w[routed + copy * bits + i] = (idx >> i) & 1
            gcs = [gc0, gc1]
            for i in range(extra):
                w[(2 + vec) * copies + i] = gcs[i]
        elif kind == GATE_POSEIDON_MDS:
            lo, hi = _mds([w[2 * i] for i in range(12)]), _mds([w[2 * i + 1] for i in range(12)])
            for r_ in range(12):
                w[2 * (12 + r_)], w[2 * (12 + r_) + 1] = lo[r_], hi[r_]
        elif kind == GATE_POSEIDON:
            w[PG_SWAP] = int(rng.integers(0, 2))
            poseidon_gate_trace(w, check=False)
        elif kind == GATE_BASE_SUM:
            for k in range(param):
                w[1 + k] = int(rng.integers(0, 2))
            w[0] = sum(w[1 + k] << k for k in range(param)) % P
        elif kind in (GATE_REDUCING, GATE_REDUCING_EXT):
            ext = kind == GATE_REDUCING_EXT
            alpha, acc = (w[2], w[3]), (w[4], w[5])
            start_accs = 6 + (2 * param if ext else param)
            for k in range(param):
                coeff = (w[6 + 2 * k], w[7 + 2 * k]) if ext else (w[6 + k], 0)
                acc = x_add(x_mul(acc, alpha), coeff)
                at = 0 if k == param - 1 else start_accs + 2 * k
                w[at], w[at + 1] = acc
        elif kind == GATE_ARITHMETIC:
            for k in range(C.num_ops):
                for col in (4 * k, 4 * k + 1, 4 * k + 2):
                    if filled and rng.random() < 0.6:           # copy constraint to an earlier cell
                        src = filled[int(rng.integers(0, len(filled)))]
                        wires[r][col] = wires[src[0]][src[1]]
                        parent[find((r, col))] = find(src)
                wires[r][4 * k + 3] = (gc0 * wires[r][4 * k] * wires[r][4 * k + 1] + gc1 * wires[r][4 * k + 2]) % P
        filled.extend((r, col) for col in range(nr))
    # ---- permutation ---------------------------------------------------------------------------------
    classes = {}
    for r in range(n):
        for col in range(nr):
            classes.setdefault(find((r, col)), []).append((r, col))
    sigma = {}
    for cells in classes.values():
        for a, b in zip(cells, cells[1:] + cells[:1]):
            assert wires[a[0]][a[1]] == wires[b[0]][b[1]]
            sigma[a] = b
    gpow = [pow(g, r, P) for r in range(n)]
    sig_vals = [[C.k_is[sigma[(r, col)][1]] * gpow[sigma[(r, col)][0]] % P for col in range(nr)] for r in range(n)]
    assert any(sigma[c] != c for c in sigma), "the permutation should not be trivial"

    ninv = inv(n)

    def interpolate(vals):
        return [ninv * sum(v * pow(g, (-j * r) % n, P) for r, v in enumerate(vals)) % P for j in range(n)]

    const_polys = [interpolate([consts[r][k] for r in range(n)]) for k in range(C.num_constants)]
    sigma_polys = [interpolate([sig_vals[r][col] for r in range(n)]) for col in range(nr)]
    wire_polys = [interpolate([wires[r][col] for r in range(n)]) for col in range(nw)]
    # ---- grand products ------------------------------------------------------------------------------
    if draw_betas_gammas:
        betas, gammas = draw_betas_gammas(const_polys, sigma_polys, wire_polys)
    else:
        betas, gammas = [rnd() for _ in range(nch)], [rnd() for _ in range(nch)]
    z_vals, pp_vals = [], []
    for i in range(nch):
        z = [1] * (n + 1)
        pp = [[0] * n for _ in range(npp)]
        for r in range(n):
            acc = z[r]
            for w, c0 in enumerate(range(0, nr, qdf)):
                for col in range(c0, min(nr, c0 + qdf)):
                    num = (betas[i] * C.k_is[col] * gpow[r] + wires[r][col] + gammas[i]) % P
                    den = (betas[i] * sig_vals[r][col] + wires[r][col] + gammas[i]) % P
                    acc = acc * num % P * inv(den) % P
                if w < npp:
                    pp[w][r] = acc
            z[r + 1] = acc
        assert z[n] == 1, "copy constraints violated"
        z_vals.append(z[:n])
        pp_vals.append(pp)

    z_polys = [interpolate(z) for z in z_vals]
    pp_polys = [[interpolate(pp_vals[i][w]) for w in range(npp)] for i in range(nch)]
    alphas = draw_alphas(z_polys, pp_polys) if draw_alphas else [rnd() for _ in range(nch)]

    # ---- vanishing polynomial on a coset, quotient ---------------------------------------------------
    m = 16 * n
    wm = pow(7, (P - 1) // m, P)
    shift = 7
    q_vals = [[0] * m for _ in range(nch)]
    for t in range(m):
        x = shift * pow(wm, t, P) % P
        cv = [poly_eval(p, x) for p in const_polys]
        sv = [poly_eval(p, x) for p in sigma_polys]
        wv = [poly_eval(p, x) for p in wire_polys]
        zh = (pow(x, n, P) - 1) % P
        l0 = zh * inv(n * (x - 1) % P) % P
        gate_terms = _constraints(C, cv, wv, pi_hash)
        terms_z1, terms_pp = [], []
        for i in range(nch):
            zx, zgx = poly_eval(z_polys[i], x), poly_eval(z_polys[i], g * x % P)
            terms_z1.append(l0 * (zx - 1) % P)
            accs = [zx] + [poly_eval(p, x) for p in pp_polys[i]] + [zgx]
            for w, c0 in enumerate(range(0, nr, qdf)):
                num = den = 1
                for col in range(c0, min(nr, c0 + qdf)):
                    num = num * ((betas[i] * C.k_is[col] * x + wv[col] + gammas[i]) % P) % P
                    den = den * ((betas[i] * sv[col] + wv[col] + gammas[i]) % P) % P
                terms_pp.append((accs[w] * num - accs[w + 1] * den) % P)
        terms = terms_z1 + terms_pp + gate_terms
        zh_inv = inv(zh)
        for i in range(nch):
            v, ap = 0, 1
            for term in terms:
                v = (v + term * ap) % P
                ap = ap * alphas[i] % P
            q_vals[i][t] = v * zh_inv % P
    minv = inv(m)
    winv = [pow(wm, (-t) % m, P) for t in range(m)]
    quotient_chunks = []
    for i in range(nch):
        coeffs = []
        for j in range(m):
            c = minv * sum(q_vals[i][t] * winv[(j * t) % m] for t in range(m)) % P
            coeffs.append(c * inv(pow(shift, j, P)) % P)
        assert all(c == 0 for c in coeffs[qdf * n:]), "quotient degree too high: the constraints do not vanish on H"
        quotient_chunks.append([coeffs[j * n:(j + 1) * n] for j in range(qdf)])

    # ---- openings ------------------------------------------------------------------------------------
    zeta = draw_zeta(quotient_chunks) if draw_zeta else (rnd(), rnd())
    gz = e_scale(zeta, g)
    open0 = [poly_eval_ext(p, zeta) for p in const_polys + sigma_polys + wire_polys + z_polys]
    open0 += [poly_eval_ext(pp_polys[i][w], zeta) for i in range(nch) for w in range(npp)]
    open0 += [poly_eval_ext(quotient_chunks[i][j], zeta) for i in range(nch) for j in range(qdf)]
    open1 = [poly_eval_ext(p, gz) for p in z_polys]
    return dict(open0=open0, open1=open1, betas=betas, gammas=gammas, alphas=alphas, zeta=zeta,
                polys=dict(constants=const_polys, sigmas=sigma_polys, wires=wire_polys, zs=z_polys, partial_products=pp_polys,
                           quotient=quotient_chunks))
