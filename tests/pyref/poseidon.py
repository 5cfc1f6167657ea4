"""Both hash families of the reference, from their definitions -- TEST INFRASTRUCTURE, pure Python.

G  Poseidon over Goldilocks, width 12, x^7, 4 + 22 + 4 rounds, in the NAIVE form: 30 x (add round constants, S-box on all
   lanes / on lane 0, MDS) with ALL_ROUND_CONSTANTS (chip/plonk/gates/poseidon.rs:26-124), MDS row r =
   sum_i CIRC[i] * state[(i + r) % 12] + DIAG[r] * state[r] (:455-487, tables :321-322).  The reference's in-circuit
   restatement runs the equivalent "fast" form (:634-686); equality of the two is what the C oracle's KAT tests check.
B  Poseidon over BN254 Fr, T = 5, x^5, 8 + 60 rounds (bn245_poseidon/native.rs:16-60, constants.rs:5-404) wrapped around
   the 12 Goldilocks limbs: 3 limbs -> x0 + x1 p + x2 p^2 (native.rs:62-67), 4 words + a zero, permute, the first 4
   words -> 3 base-p digits each (native.rs:69-77, bn245_poseidon/plonky2_config.rs:38-51).

Sponge and compression: HasherChip::hash / ::permute (chip/hasher_chip.rs:122-171), hash_or_noop
(chip/merkle_proof_chip.rs:52-56)."""
import json
import os

import numpy as np

from . import gl
from .gl import P

HASH_G, HASH_B = 0, 1
_C = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "constants.json")))
RC = [int(x, 16) for x in _C["g_round_constants"]]
CIRC, DIAG = _C["g_mds_circ"], _C["g_mds_diag"]
R_BN = 21888242871839275222246405745257275088548364400416034343698204186575808495617
B_RC = [int(x, 16) for x in _C["b_round_constants"]]
B_MDS = [[int(x, 16) for x in _C["b_mds"][5 * i:5 * i + 5]] for i in range(5)]
N_FULL_HALF, N_PARTIAL = 4, 22


# ---- family G, one state (the definition) -----------------------------------------------------------------
def _mds(s):
    return [(sum(CIRC[i] * s[(i + r) % 12] for i in range(12)) + DIAG[r] * s[r]) % P for r in range(12)]


def permute_g(state):
    s = [x % P for x in state]
    for rnd in range(2 * N_FULL_HALF + N_PARTIAL):
        s = [(x + RC[12 * rnd + i]) % P for i, x in enumerate(s)]
        if rnd < N_FULL_HALF or rnd >= N_FULL_HALF + N_PARTIAL:
            s = [pow(x, 7, P) for x in s]
        else:
            s[0] = pow(s[0], 7, P)
        s = _mds(s)
    return s


# ---- family G, n states at once: numpy uint64 array of shape (12, n) ---------------------------------------
_RCV = np.array(RC, dtype=np.uint64).reshape(30, 12)
# row r of the MDS as a 12-vector of small coefficients
_MDSM = np.zeros((12, 12), dtype=np.uint64)
for _r in range(12):
    for _i in range(12):
        _MDSM[_r, (_i + _r) % 12] += np.uint64(CIRC[_i])
    _MDSM[_r, _r] += np.uint64(DIAG[_r])


def _vmds(s):
    """coefficients < 2^6, 12 terms: the 32-bit halves are summed exactly in uint64 and recombined"""
    lo = _MDSM @ (s & np.uint64(0xFFFFFFFF))            # < 2^42
    hi = _MDSM @ (s >> np.uint64(32))                   # < 2^42;  value = lo + 2^32 hi
    low = lo + (hi << np.uint64(32))                    # low 64 bits (wrapping) ...
    return gl.vreduce128(low, (hi >> np.uint64(32)) + (low < lo).astype(np.uint64))   # ... and the rest


def permute_g_batch(states):
    """states: (n, 12) canonical uint64 -> (n, 12)"""
    s = gl.varr(states).T.copy()
    for rnd in range(30):
        s = gl.vadd(s, _RCV[rnd][:, None])
        if rnd < N_FULL_HALF or rnd >= N_FULL_HALF + N_PARTIAL:
            s = gl.vpow7(s)
        else:
            s[0] = gl.vpow7(s[0])
        s = _vmds(s)
    return np.ascontiguousarray(s.T)


# ---- family B ---------------------------------------------------------------------------------------------
def permute_b_fr(st):
    st = [x % R_BN for x in st]
    c = 0
    for rnd in range(68):
        st = [(x + B_RC[c + i]) % R_BN for i, x in enumerate(st)]
        c += 5
        if rnd < 4 or rnd >= 64:
            st = [pow(x, 5, R_BN) for x in st]
        else:
            st[0] = pow(st[0], 5, R_BN)
        st = [sum(st[j] * B_MDS[i][j] for j in range(5)) % R_BN for i in range(5)]
    return st


def permute_b(state):
    s = [x % P for x in state]
    words = [s[3 * k] + s[3 * k + 1] * P + s[3 * k + 2] * P * P for k in range(4)] + [0]
    out = permute_b_fr(words)
    res = []
    for k in range(4):
        v = out[k]
        for _ in range(3):
            res.append(v % P)
            v //= P
    return res


# ---- the permutation of a family, scalar and batched -------------------------------------------------------
def permute(state, kind=HASH_G):
    return permute_b(state) if kind == HASH_B else permute_g(state)


def permute_batch(states, kind=HASH_G):
    states = gl.varr(states)
    if states.shape[0] == 0:
        return states.copy()
    if kind == HASH_B:
        return np.array([permute_b([int(v) for v in row]) for row in states], dtype=np.uint64)
    return permute_g_batch(states)


# ---- sponge / compression ---------------------------------------------------------------------------------
def hash_no_pad(inputs, kind=HASH_G):
    """HasherChip::hash with 4 outputs: overwrite-mode sponge, rate 8, on a fresh zero state (hasher_chip.rs:122-148)"""
    st = [0] * 12
    for off in range(0, len(inputs), 8):
        chunk = [int(v) for v in inputs[off:off + 8]]
        st[:len(chunk)] = chunk
        st = permute(st, kind)
    return st[:4]


def two_to_one(left, right, kind=HASH_G):
    """HasherChip::permute on a fresh hasher (hasher_chip.rs:150-171, merkle_proof_chip.rs:58-71)"""
    return permute([int(v) for v in left] + [int(v) for v in right] + [0, 0, 0, 0], kind)[:4]


def hash_or_noop(leaf, kind=HASH_G):
    """merkle_proof_chip.rs:52-56: a leaf of at most 4 elements is its own digest (zero-padded)"""
    leaf = [int(v) for v in leaf]
    return leaf + [0] * (4 - len(leaf)) if len(leaf) <= 4 else hash_no_pad(leaf, kind)


def hash_or_noop_batch(rows, kind=HASH_G):
    """rows: (n, leaf_len) -> (n, 4)"""
    rows = gl.varr(rows)
    n, ll = rows.shape
    if ll <= 4:
        out = np.zeros((n, 4), dtype=np.uint64)
        out[:, :ll] = rows
        return out
    st = np.zeros((n, 12), dtype=np.uint64)
    for off in range(0, ll, 8):
        w = min(8, ll - off)
        st[:, :w] = rows[:, off:off + w]
        st = permute_batch(st, kind)
    return np.ascontiguousarray(st[:, :4])


def two_to_one_batch(left, right, kind=HASH_G):
    left, right = gl.varr(left), gl.varr(right)
    st = np.zeros((left.shape[0], 12), dtype=np.uint64)
    st[:, 0:4] = left
    st[:, 4:8] = right
    return np.ascontiguousarray(permute_batch(st, kind)[:, :4])
