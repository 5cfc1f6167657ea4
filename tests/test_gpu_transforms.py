"""GPU side of the NTT / LDE library: ntt_pass_kernel (plain, inverse, LDE first pass) against the host twin of the
same phase functions (pinned by tests/test_ntt.py), and the whole commit phase on the device -- LDE -> leaves -> Merkle cap --
against the cap a complete proof of the Python prover carries."""
import numpy as np
import pytest

from common import P

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", [1, 2, 5, 9, 10, 12, 13, 14, 15, 19, 22, 23])
def test_ntt_kernel_matches_host_twin(svb, ctx, k):
    rng = np.random.default_rng(k)
    n_polys = 3 if k < 19 else (2 if k < 23 else 1)
    a = rng.integers(0, P, size=(n_polys, 1 << k), dtype=np.uint64)
    a[0, :] = P - 1
    want = svb.ntt_host(a, nthreads=4)
    got = ctx.ntt_batch(a)
    assert (got == want).all()
    back = ctx.ntt_batch(got, inverse=True)
    assert (back == a).all()


def test_ntt_device_memory_and_many_polys(svb, ctx):
    import torch
    rng = np.random.default_rng(3)
    k, n_polys = 12, 300
    a = rng.integers(0, P, size=(n_polys, 1 << k), dtype=np.uint64)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    torch.cuda.synchronize()
    ctx.ntt_batch(d.data_ptr(), log_n=k, n_polys=n_polys, mem=svb.MEM_DEVICE)
    ctx.synchronize()
    assert (d.cpu().numpy().view(np.uint64) == svb.ntt_host(a, nthreads=4)).all()
    ctx.ntt_batch(d.data_ptr(), log_n=k, n_polys=n_polys, inverse=True, mem=svb.MEM_DEVICE)
    ctx.synchronize()
    assert (d.cpu().numpy().view(np.uint64) == a).all()


@pytest.mark.parametrize("k,rate_bits", [(1, 1), (4, 3), (9, 1), (12, 3), (13, 2), (14, 2), (17, 1)])
def test_lde_kernel_matches_host_twin(svb, ctx, k, rate_bits):
    rng = np.random.default_rng(10 * k + rate_bits)
    c = rng.integers(0, P, size=(5, 1 << k), dtype=np.uint64)
    assert (ctx.lde_batch(c, rate_bits) == svb.lde_host(c, rate_bits, nthreads=4)).all()
    assert (ctx.lde_batch(c, rate_bits, shift=1)[:, 0] == np.array([sum(int(v) for v in row) % P for row in c], dtype=np.uint64)).all()


def test_commit_phase_on_the_device_reproduces_a_proofs_cap(svb, orc, ctx):
    """coefficients -> sv_lde_batch -> leaf-major -> sv_merkle_tree_build: the cap equals the wires cap inside a complete
    proof made by the independent Python prover (which hashed the same leaves through the CPU oracle)."""
    import torch
    import full_prover as fp
    from test_plonk_check import CONFIGS
    C, params = fp.toy_setup(svb, CONFIGS["recursion_gate_set"])
    rng = np.random.default_rng(2)
    cd = rng.integers(0, P, size=4, dtype=np.uint64)
    rec, out = fp.prove_full(C, params, 4, rng.integers(0, P, size=3, dtype=np.uint64), cd)
    L = svb.api.make_layout(params)
    capw = 4 * L.ncap
    for oracle_index, polys in ((1, out["polys"]["wires"]), (0, out["polys"]["constants"] + out["polys"]["sigmas"])):
        coeffs = np.array(polys, dtype=np.uint64)
        lde = ctx.lde_batch(coeffs, params.config.rate_bits)                       # (n_polys, N)
        leaves = np.ascontiguousarray(lde.T)                                       # (N, n_polys): one row per leaf
        layers = ctx.merkle_tree_build(leaves, leaves.shape[1], params.config.cap_height)
        assert (layers[-1].reshape(-1) == rec[L.off_init_caps + oracle_index * capw: L.off_init_caps + (oracle_index + 1) * capw]).all()


@pytest.mark.parametrize("kind", [0, 1])
def test_commit_batch_one_call(svb, orc, ctx, kind):
    """sv_commit_batch = LDE + leaf-major copy + Merkle tree; against the host LDE twin and the oracle's hashes."""
    rng = np.random.default_rng(7 + kind)
    k, rate_bits, cap_height, n_polys = 5, 2, 1, 37                    # 37 columns: ragged transposition tiles
    coeffs = rng.integers(0, P, size=(n_polys, 1 << k), dtype=np.uint64)
    leaves, layers = ctx.commit_batch(coeffs, rate_bits, cap_height, hash_kind=kind)
    want_leaves = np.ascontiguousarray(svb.lde_host(coeffs, rate_bits).T)
    assert (leaves == want_leaves).all()
    N = 1 << (k + rate_bits)
    assert [l.shape[0] for l in layers] == [N >> i for i in range(k + rate_bits - cap_height + 1)]
    for i in (0, 1, N - 1):
        assert (layers[0][i] == orc.hash_no_pad(want_leaves[i], kind)).all()
    for lv in range(1, len(layers)):
        for j in (0, layers[lv].shape[0] - 1):
            assert (layers[lv][j] == orc.two_to_one(layers[lv - 1][2 * j], layers[lv - 1][2 * j + 1], kind)).all()
