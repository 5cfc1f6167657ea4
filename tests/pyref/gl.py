"""Goldilocks F_p, p = 2^64 - 2^32 + 1 (native_chip/arithmetic_chip.rs:19) and F_p^2 = F_p[X]/(X^2 - 7)
(chip/goldilocks_extension_chip.rs:49) -- TEST INFRASTRUCTURE, pure Python.

Scalars are Python integers reduced with `% P` (the definition).  The v* functions are the same operations on numpy uint64
arrays of canonical values, exact (32-bit limb products, no floating point); tests/test_pyref.py checks them against the
scalar definition on corner values and random inputs."""
import numpy as np

P = 0xFFFFFFFF00000001
EPS = 0xFFFFFFFF            # 2^64 mod p
GENERATOR = 7               # GoldilocksField::MULTIPLICATIVE_GROUP_GENERATOR
W = 7                       # X^2 = 7


def inv(a):
    return pow(a % P, P - 2, P)


def root_of_unity(bits):
    """primitive 2^bits-th root of unity the reference uses: 7^((p-1)/2^bits)  (chip/fri_chip.rs:160-163)"""
    return pow(GENERATOR, (P - 1) >> bits, P)


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


# ---- F_p^2, elements are (c0, c1) tuples -------------------------------------------------------------------
def e(a):
    return (a % P, 0)


def e_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def e_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def e_mul(a, b):
    """(a0 b0 + 7 a1 b1, a0 b1 + a1 b0)   native_chip/arithmetic_chip.rs:109-132"""
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def e_scale(a, s):
    return (a[0] * s % P, a[1] * s % P)


def e_inv(a):
    """1/(a0 + a1 X) = (a0 - a1 X) / (a0^2 - 7 a1^2)"""
    n = inv((a[0] * a[0] - W * a[1] * a[1]) % P)
    return (a[0] * n % P, (P - a[1]) * n % P if a[1] else 0)


def e_pow(a, k):
    r = (1, 0)
    while k:
        if k & 1:
            r = e_mul(r, a)
        a = e_mul(a, a)
        k >>= 1
    return r


# ---- exact uint64 vectors ----------------------------------------------------------------------------------
_M32 = np.uint64(0xFFFFFFFF)
_EPS = np.uint64(EPS)
_P = np.uint64(P)
_S32 = np.uint64(32)


def varr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def _canon(r):
    return np.where(r >= _P, r - _P, r)


def vreduce128(lo, hi):
    """(lo + 2^64 hi) mod p with 2^64 = 2^32 - 1 and 2^96 = -1 (mod p)"""
    h0, h1 = hi & _M32, hi >> _S32
    t = lo - h1
    t = np.where(lo < h1, t - _EPS, t)          # the wrap added 2^64 = EPS (mod p)
    u = h0 * _EPS
    r = t + u
    r = np.where(r < u, r + _EPS, r)            # the wrap dropped 2^64
    return _canon(r)


def vmul(a, b):
    a, b = varr(a), varr(b)
    a0, a1, b0, b1 = a & _M32, a >> _S32, b & _M32, b >> _S32
    p00, p01, p10, p11 = a0 * b0, a0 * b1, a1 * b0, a1 * b1
    mid = p01 + p10
    mid_c = (mid < p01).astype(np.uint64)
    lo = p00 + (mid << _S32)
    lo_c = (lo < p00).astype(np.uint64)
    hi = p11 + (mid >> _S32) + (mid_c << _S32) + lo_c
    return vreduce128(lo, hi)


def vadd(a, b):
    a, b = varr(a), varr(b)
    r = a + b
    r = np.where(r < a, r + _EPS, r)            # canonical inputs: a + b < 2p, one correction is enough
    return _canon(r)


def vsub(a, b):
    a, b = varr(a), varr(b)
    r = a - b
    return np.where(a < b, r + _P, r)


def vneg(a):
    a = varr(a)
    return np.where(a == 0, a, _P - a)


def vpow7(x):
    x2 = vmul(x, x)
    x4 = vmul(x2, x2)
    return vmul(vmul(x, x2), x4)
