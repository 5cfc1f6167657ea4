"""Exactness of the FP64 MDS layers (stark-verifier_b200/csrc/poseidon_g.cuh: mds_freq_half, the frequency-domain circulant on
SUBNORMAL binary64 operands) at the extremes, against big-integer arithmetic -- the claim the kernel's comments argue (row sums
< 2^42 for 32-bit halves, < 2^49 for the second layer of a double-layer partial round) driven with all-0xFFFFFFFF, all-zero and
alternating halves instead of random states.  mds_freq_half is one __host__ __device__ function: the host instantiation runs
here (x86 keeps subnormals unless FTZ/DAZ is switched on, which g++ -O2 does not do)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from common import P

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CIRC = [17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20]          # MDS_MATRIX_CIRC, chip/plonk/gates/poseidon.rs:321
DIAG0 = 8                                                        # MDS_MATRIX_DIAG[0], :322


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("mds") / "mds_probe.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "host_twins", "mds_probe.cpp")])
    lib = ctypes.CDLL(so)
    lib.mds_freq_half_bits.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]

    def run(x, rc=None, mode=0):
        xb = np.array(x, dtype=np.uint64)                       # integer n as the bit pattern of the subnormal n * 2^-1074
        rb = np.array(rc if rc is not None else [0] * 12, dtype=np.uint64)
        y = np.zeros(12, dtype=np.uint64)
        lib.mds_freq_half_bits(xb.ctypes.data, rb.ctypes.data, mode, y.ctypes.data)
        return [int(v) for v in y]
    return run


def mds_int(x, rc=None):
    return [sum(CIRC[k] * x[(i + k) % 12] for k in range(12)) + (DIAG0 * x[0] if i == 0 else 0) + (rc[i] if rc else 0) for i in range(12)]


M32 = 0xFFFFFFFF


def extremal_vectors(limit):
    yield [0] * 12
    yield [limit] * 12
    yield [limit if i % 2 else 0 for i in range(12)]
    yield [0 if i % 2 else limit for i in range(12)]
    yield [limit if i % 3 == 0 else 0 for i in range(12)]
    yield [limit if (i // 3) % 2 else 0 for i in range(12)]       # the 4-point DFT's own pattern: maximises |Ur|, |Ui|
    yield [limit if i in (0, 6) else 0 for i in range(12)]
    yield [limit if i in (3, 9) else 0 for i in range(12)]
    for i in range(12):
        yield [limit if j == i else 0 for j in range(12)]
        yield [0 if j == i else limit for j in range(12)]
    rng = np.random.default_rng(7)
    for _ in range(300):
        yield [int(limit if b else 0) for b in rng.integers(0, 2, size=12)]
    for _ in range(300):
        yield [int(v) for v in rng.integers(0, limit + 1, size=12, dtype=np.uint64)]


def test_single_layer_on_32_bit_halves(probe):
    """operands = 32-bit halves (as the full rounds feed them), round constants = 32-bit halves too: every output is the exact
    integer (< 2^42), including the all-ones state; the all-zero state gives +0 bit patterns (no -0)"""
    rc = [M32 - i for i in range(12)]
    for x in extremal_vectors(M32):
        assert probe(x) == mds_int(x)
        assert probe(x, rc, 12) == mds_int(x, rc)
        assert probe(x, [rc[0]] + [0] * 11, 1) == mds_int(x, [rc[0]] + [0] * 11)
    assert probe([0] * 12) == [0] * 12
    assert max(mds_int([M32] * 12)) < 1 << 42


def test_second_layer_of_a_double_layer_stays_exact(probe):
    """the partial rounds feed the un-recombined half-sums of one MDS (< 2^42 each after the constant) straight into the next
    one: sums < 2^50, still inside the 2^52 window where subnormal binary64 arithmetic is integer arithmetic"""
    limit = max(mds_int([M32] * 12, [M32] * 12))
    assert limit < 1 << 42
    for x in extremal_vectors(limit):
        x = [M32 if i == 0 and x[0] > M32 else x[i] for i in range(12)]     # lane 0 is recombined and reduced: a 32-bit half again
        got = probe(x, [M32] + [0] * 11, 1)
        assert got == mds_int(x, [M32] + [0] * 11)
        assert max(got) < 1 << 52


def test_double_layer_partial_rounds_against_big_integers(probe):
    """two naive partial rounds computed the kernel's way (halves through mds_freq_half twice, lane 0 recombined and passed
    through x^7 in between, lanes 1..11 left as half-sums) == the definition on Python integers mod p"""
    rng = np.random.default_rng(11)
    states = [[P - 1] * 12, [0] * 12, [M32 << 32] * 12, [M32] * 12, [(M32 << 32) if i % 2 else M32 for i in range(12)]]
    states += [[int(v) for v in rng.integers(0, P, size=12, dtype=np.uint64)] for _ in range(50)]
    for s in states:
        c1, c2 = int(rng.integers(0, P, dtype=np.uint64)), int(rng.integers(0, P, dtype=np.uint64))   # lane-0 constants of the two rounds
        # definition: (S-box on lane 0, MDS, + constant on lane 0) twice
        want = list(s)
        for c in (c1, c2):
            want[0] = pow(want[0], 7, P)
            want = [v % P for v in mds_int(want)]
            want[0] = (want[0] + c) % P
        # the kernel's route
        t = list(s)
        t[0] = pow(t[0], 7, P)
        yl = probe([v & M32 for v in t], [c1 & M32] + [0] * 11, 1)
        yh = probe([v >> 32 for v in t], [c1 >> 32] + [0] * 11, 1)
        s0 = pow((yl[0] + (yh[0] << 32)) % P, 7, P)
        zl = probe([s0 & M32] + yl[1:], [c2 & M32] + [0] * 11, 1)
        zh = probe([s0 >> 32] + yh[1:], [c2 >> 32] + [0] * 11, 1)
        assert [(a + (b << 32)) % P for a, b in zip(zl, zh)] == want
