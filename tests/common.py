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
        if c == "step_sibling" and L.step_depth[0] == 0:
            c = "step_eval"   # a step tree of depth 0 has no siblings
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


# The fourth vector of upstream plonky2's `test_vectors12` (plonky2/src/hash/poseidon_goldilocks.rs, width-12
# known-answer test; the other three are the all-zero, iota and all-(p-1) states of tests/golden/poseidon_g.json).
# It is published by plonky2 itself, i.e. not derived from any code of this repository.
PLONKY2_TV12_IN = [0x8ccbbbea4fe5d2b7, 0xc2af59ee9ec49970, 0x90f7e1a9e658446a, 0xdcc0630a3ab8b1b8,
                   0x7ff8256bca20588c, 0x5d99a7ca0c44ecfb, 0x48452b17a70fbee3, 0xeb09d654690b6c88,
                   0x4a55d3a39c676a88, 0xc0407a38d2285139, 0xa234bac9356386d1, 0xe1633f2bad98a52f]
PLONKY2_TV12_OUT = [0xa89280105650c4ec, 0xab542d53860d12ed, 0x5704148e9ccab94f, 0xd3a826d4b62da9f5,
                    0x8a7a6ca87892574f, 0xc7017e1cad1a674e, 0x1f06668922318e34, 0xa3b203bc8102676f,
                    0xfcc781b0ce382bf2, 0x934c69ff3ed14ba5, 0x504688a5996e8f13, 0x401f3f2ed524a2ba]
