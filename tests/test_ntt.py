"""NTT / LDE, CPU side: the host twin runs the kernels' own phase functions (ntt.hpp: ntt_tile_load /
ntt_tile_round / ntt_tile_store), so these tests pin the index arithmetic, the pass plan (1, 2 and 3 passes), the radix-8 / 4 / 2
rounds, the twiddle factorisation, the LDE-as-cosets decomposition and the output order against naive big-integer evaluation
and against the Merkle-leaf convention of the independent Python prover."""
import numpy as np
import pytest

from common import P


def bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def naive_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + int(c)) % P
    return acc


@pytest.mark.parametrize("k", [1, 2, 3, 5, 7])
def test_forward_matches_naive_dft_in_bit_reversed_order(svb, k):
    rng = np.random.default_rng(k)
    a = rng.integers(0, P, size=(3, 1 << k), dtype=np.uint64)
    a[1] = 0
    a[1, 1] = 1                                             # the polynomial x: values are the points themselves
    got = svb.ntt_host(a)
    w = pow(7, (P - 1) >> k, P)
    for p in range(3):
        want = [naive_eval(a[p], pow(w, bitrev(i, k), P)) for i in range(1 << k)]
        assert [int(v) for v in got[p]] == want
    assert (svb.ntt_host(got, inverse=True) == a).all()


@pytest.mark.parametrize("k", [9, 10, 11, 12, 13, 14, 15, 17, 20, 23])
def test_multi_pass_sizes_sampled_points_and_round_trip(svb, k):
    """up to 13: one pass (rounds 3+3+.. with a last round of 1, 2 or 3 stages); 14-22: two passes; 23: three.  Values at sampled
    positions against Horner; the inverse undoes it."""
    rng = np.random.default_rng(100 + k)
    n = 1 << k
    a = rng.integers(0, P, size=(2 if k < 20 else 1, n), dtype=np.uint64)
    if k > 13:
        a[:, 64:] = 0                                       # sparse enough for the big-integer Horner below
        deg = 64
    else:
        deg = n
    got = svb.ntt_host(a, nthreads=8)
    w = pow(7, (P - 1) >> k, P)
    for i in [0, 1, 2, n // 2, n - 1] + [int(x) for x in rng.integers(0, n, size=6)]:
        x = pow(w, bitrev(i, k), P)
        for p in range(a.shape[0]):
            assert int(got[p, i]) == naive_eval(a[p, :deg], x), (k, i)
    assert (svb.ntt_host(got, inverse=True, nthreads=8) == a).all()


@pytest.mark.parametrize("k,rate_bits", [(3, 1), (5, 3), (9, 2), (12, 3), (13, 1), (14, 2), (16, 1), (17, 1)])
def test_lde_sampled_points(svb, k, rate_bits):
    """LDE = 2^rate_bits size-n transforms of the coset-scaled coefficients (ntt.hpp): value i of the output is
    f(shift * omega_N^bitrev(i)), for one- and two-pass n and both levels of the scale table (k > 16: two-level)"""
    rng = np.random.default_rng(k * 7 + rate_bits)
    n, N, lN = 1 << k, 1 << (k + rate_bits), k + rate_bits
    c = rng.integers(0, P, size=(2, n), dtype=np.uint64)
    deg = n
    if k > 12:
        c[:, 48:n - 16] = 0                                 # keeps the big-integer Horner affordable: low and top coefficients
    for shift in (7, 1, 12345):
        got = svb.lde_host(c, rate_bits, shift=shift, nthreads=8)
        wN = pow(7, (P - 1) >> lN, P)
        for i in [0, 1, N // 2, N - 1, n, n + 1] + [int(x) for x in rng.integers(0, N, size=4)]:
            x = shift * pow(wN, bitrev(i, lN), P) % P
            for p in range(2):
                if k > 12:
                    want = (naive_eval(c[p, :48], x) + pow(x, n - 16, P) * naive_eval(c[p, n - 16:], x)) % P
                else:
                    want = naive_eval(c[p, :deg], x)
                assert int(got[p, i]) == want, (k, rate_bits, shift, i)


def test_linearity_and_convolution(svb):
    k = 11
    rng = np.random.default_rng(5)
    n = 1 << k
    a = rng.integers(0, P, size=n, dtype=np.uint64)
    b = rng.integers(0, P, size=n, dtype=np.uint64)
    a[n // 2:] = 0
    b[n // 2:] = 0
    fa, fb = svb.ntt_host(a[None])[0], svb.ntt_host(b[None])[0]
    prod = np.array([int(x) * int(y) % P for x, y in zip(fa, fb)], dtype=np.uint64)
    c = svb.ntt_host(prod[None], inverse=True)[0]
    # coefficient 5 and the top coefficient of the product polynomial, by hand
    assert int(c[5]) == sum(int(a[i]) * int(b[5 - i]) for i in range(6)) % P
    assert int(c[n - 2]) == int(a[n // 2 - 1]) * int(b[n // 2 - 1]) % P and int(c[n - 1]) == 0


def test_lde_is_the_merkle_leaf_order_of_the_python_prover(svb, orc):
    """sv_lde_host of a proof's wire polynomials == the leaves the independent Python prover commits to
    (leaf i = evaluations at 7 * omega_N^bitrev(i)); and the Merkle cap over those leaves is the proof's wires cap."""
    import full_prover as fp
    import plonk_prover as pp
    from test_plonk_check import CONFIGS
    C, params = fp.toy_setup(svb, CONFIGS["one_selector"])
    rng = np.random.default_rng(2)
    cd = rng.integers(0, P, size=4, dtype=np.uint64)
    rec, out = fp.prove_full(C, params, 4, rng.integers(0, P, size=3, dtype=np.uint64), cd)
    wires = np.array(out["polys"]["wires"], dtype=np.uint64)            # (num_wires, n) coefficients
    lde = svb.lde_host(wires, params.config.rate_bits, shift=7)
    L = svb.api.make_layout(params)
    N = 1 << L.lde_bits
    w = pow(7, (P - 1) >> L.lde_bits, P)
    for i in (0, 1, 5, N - 1):
        x = 7 * pow(w, bitrev(i, L.lde_bits), P) % P
        assert [int(v) for v in lde[:, i]] == [naive_eval(p, x) for p in wires]
    from pyref.merkle import MerkleTree
    tree = MerkleTree(np.ascontiguousarray(lde.T), params.config.cap_height)
    capw = 4 * L.ncap
    assert [v for d in tree.cap() for v in d] == [int(v) for v in rec[L.off_init_caps + capw: L.off_init_caps + 2 * capw]]


def test_bad_arguments(svb):
    with pytest.raises(svb.SvError):
        svb.ntt_host(np.full((1, 8), P, dtype=np.uint64))               # non-canonical input
    with pytest.raises(svb.SvError):
        svb.lde_host(np.zeros((1, 8), dtype=np.uint64), 2, shift=0)
