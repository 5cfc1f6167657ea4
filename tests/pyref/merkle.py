"""Merkle trees with caps -- TEST INFRASTRUCTURE, pure Python.

Prover side: plonky2's MerkleTree::new(leaves, cap_height) / prove(index) (call sites plonky2_semaphore/access_set.rs:25,
circuit.rs:91): digests = hash_or_noop(leaf); parent = two_to_one(left, right); the cap is the layer with 2^cap_height
nodes; a proof is the list of siblings bottom-up below the cap.
Verifier side: MerkleProofChip::verify_merkle_proof_to_cap_with_cap_index (chip/merkle_proof_chip.rs:39-87)."""
import numpy as np

from . import poseidon as ps


class MerkleTree:
    def __init__(self, leaves, cap_height, kind=ps.HASH_G):
        """leaves: (n, leaf_len) array-like of canonical field elements, n a power of two >= 2^cap_height"""
        self.leaves = np.ascontiguousarray(np.asarray(leaves, dtype=np.uint64))
        n = self.leaves.shape[0]
        assert n & (n - 1) == 0 and n >= (1 << cap_height)
        self.kind, self.cap_height = kind, cap_height
        self.layers = [ps.hash_or_noop_batch(self.leaves, kind)]
        while self.layers[-1].shape[0] > (1 << cap_height):
            cur = self.layers[-1]
            self.layers.append(ps.two_to_one_batch(cur[0::2], cur[1::2], kind))

    def cap(self):
        """list of 2^cap_height digests (4 ints each)"""
        return [[int(v) for v in d] for d in self.layers[-1]]

    def prove(self, index):
        sibs = []
        for layer in self.layers[:-1]:
            sibs.append([int(v) for v in layer[index ^ 1]])
            index >>= 1
        return sibs

    def leaf(self, index):
        return [int(v) for v in self.leaves[index]]


def verify_to_cap(leaf, index_bits, cap_index, cap, siblings, kind=ps.HASH_G):
    """merkle_proof_chip.rs:39-87.  index_bits: LSB first, zipped with the siblings (:58); cap_index chosen by the caller
    (the FRI verifier passes the same one for every tree, chip/fri_chip.rs:252,308)."""
    state = ps.hash_or_noop(leaf, kind)                                     # :46-56
    for bit, sib in zip(index_bits, siblings):                              # :58
        sib = [int(v) for v in sib]
        state = ps.two_to_one(sib, state, kind) if bit else ps.two_to_one(state, sib, kind)   # :60-70 select by bit
    return state == [int(v) for v in cap[cap_index]]                        # :73-84


def verify_to_cap_batch(items, kind=ps.HASH_G):
    """Many independent checks at once (same semantics as verify_to_cap), grouped so that the hashing is vectorised.
    items: list of (leaf, index_bits, cap_index, cap, siblings) -> list of bool"""
    out = [None] * len(items)
    groups = {}
    for i, it in enumerate(items):
        groups.setdefault((len(it[0]), min(len(it[4]), len(it[1]))), []).append(i)
    for (ll, depth), idxs in groups.items():
        rows = np.array([[int(v) for v in items[i][0]] for i in idxs], dtype=np.uint64).reshape(len(idxs), ll)
        state = ps.hash_or_noop_batch(rows, kind)
        for lvl in range(depth):
            sib = np.array([[int(v) for v in items[i][4][lvl]] for i in idxs], dtype=np.uint64)
            bit = np.array([items[i][1][lvl] for i in idxs], dtype=bool)[:, None]
            left, right = np.where(bit, sib, state), np.where(bit, state, sib)
            state = ps.two_to_one_batch(left, right, kind)
        for j, i in enumerate(idxs):
            out[i] = [int(v) for v in state[j]] == [int(v) for v in items[i][3][items[i][2]]]
    return out
