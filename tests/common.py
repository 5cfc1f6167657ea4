"""Shared helpers for the parity tests: shapes, seeded corruption, bit access."""
import numpy as np

P = 0xFFFFFFFF00000001


def tiny_params(svb, hiding=False, cap=2, queries=6, pow_bits=4, degree_bits=7, rate_bits=3, **kw):
    return svb.api._params(degree_bits, rate_bits, cap, pow_bits, queries, hiding=hiding, **kw)


def bit(bitmap, i):
    return (int(bitmap[i >> 5]) >> (i & 31)) & 1


def corrupt(recs, L, rng, every=8, num_steps=None):
    """Seeded negative controls, round-robin over the corruption classes of SURVEY 8d config 2:
    one sibling limb / one leaf eval / one step eval / one final-poly coeff / pow response /
    a non-canonical word / a step sibling.  Returns {proof index: class name}."""
    classes = ["sibling", "leaf", "step_eval", "final_poly", "pow", "noncanonical", "step_sibling", "cap", "opening"]
    out = {}
    n = recs.shape[0]
    nq = (recs.shape[1] - L.header_words) // L.query_words
    k = 0
    for i in range(every // 2, n, every):
        c = classes[k % len(classes)]
        k += 1
        if num_steps == 0 and c in ("step_eval", "step_sibling"):
            c = "leaf"   # a shape without reduction steps has no step data to corrupt
        q = int(rng.integers(0, nq))
        qb = L.header_words + q * L.query_words
        delta = np.uint64(1) << np.uint64(int(rng.integers(0, 20)))
        if c == "sibling":
            o = int(rng.integers(0, 4))
            recs[i, qb + L.q_off_init_sibs[o] + int(rng.integers(0, 4 * L.init_depth))] ^= delta
        elif c == "leaf":
            o = int(rng.integers(0, 4))
            recs[i, qb + L.q_off_init_evals[o] + int(rng.integers(0, L.leaf_len[o]))] ^= delta
        elif c == "step_eval":
            ns = num_steps if num_steps is not None else max(1, sum(1 for j in range(32) if L.q_off_step_evals[j] or j == 0))
            st = int(rng.integers(0, ns))
            recs[i, qb + L.q_off_step_evals[st] + int(rng.integers(0, 4))] ^= delta
        elif c == "final_poly":
            recs[i, L.off_final_poly + int(rng.integers(0, 8))] ^= delta
        elif c == "pow":
            recs[i, L.off_pow_response] |= np.uint64(1 << 63)
        elif c == "noncanonical":
            o = int(rng.integers(0, 4))
            recs[i, qb + L.q_off_init_sibs[o] + int(rng.integers(0, 4 * L.init_depth))] = np.uint64(P + int(rng.integers(0, 1000)))
        elif c == "step_sibling":
            recs[i, qb + L.q_off_step_sibs[0] + int(rng.integers(0, 4 * L.step_depth[0]))] ^= delta
        elif c == "cap":
            # the cap entry query q actually lands on (a never-used entry would not reject the proof)
            idx = int(recs[i, L.off_indices + q]) & ((1 << L.lde_bits) - 1)
            ci = idx >> (L.lde_bits - (L.ncap.bit_length() - 1))
            o = int(rng.integers(0, 4))
            recs[i, L.off_init_caps + (o * L.ncap + ci) * 4 + int(rng.integers(0, 4))] ^= delta
        elif c == "opening":
            recs[i, L.off_open0 + int(rng.integers(0, 2 * L.n0))] ^= delta
        out[i] = c
    return out
