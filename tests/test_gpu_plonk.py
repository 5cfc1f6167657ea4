"""GPU side of the plonk-level checks: plonk_check_kernel (one thread per proof) against the host twin
of the same function and the oracle, on proofs from the pure-Python prover and on arbitrary inputs."""
import numpy as np
import pytest

import plonk_prover as pp
from common import P, bit
from test_plonk_check import CONFIGS, oracle_bits, plonk_setup, to_records

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["one_selector", "two_selectors", "three_chunks", "all_gates", "recursion_gate_set"])
def test_plonk_kernel_matches_host_and_oracle(svb, orc, ctx, name):
    C, params, circuit, L = plonk_setup(svb, CONFIGS[name])
    ocirc = orc.plonk_circuit_from(circuit)
    rng = np.random.default_rng(3)
    n = 77                                                  # ragged last bitmap word, partial last block
    recs = np.zeros((n, L.record_words), dtype=np.uint64)
    recs[:, :L.header_words] = rng.integers(0, P, size=(n, L.header_words), dtype=np.uint64)
    pih = rng.integers(0, P, size=(n, 4), dtype=np.uint64)
    chal = rng.integers(0, P, size=(n, 3 * C.num_challenges), dtype=np.uint64)
    good = [0, 31, 32, 45, 76]
    proofs = [pp.prove(C, 50 + k, [int(x) for x in pih[i]]) for k, i in enumerate(good)]
    grecs, gchal = to_records(L, proofs)
    for k, i in enumerate(good):
        recs[i], chal[i] = grecs[k], gchal[k]
    recs[45, L.off_open0 + 7] ^= np.uint64(2)               # a valid proof with one flipped opening bit
    recs[5, L.off_zeta:L.off_zeta + 2] = (1, 0)             # division by zero in L_0
    recs[6, L.off_open1] = np.uint64(P + 1)                 # non-canonical
    want = svb.plonk_check_host(params, circuit, recs, pih, chal, nthreads=4)
    assert [i for i in range(n) if bit(want, i)] == [0, 31, 32, 76]
    assert [bit(want, i) for i in range(n)] == oracle_bits(orc, ocirc, L, recs, pih, chal)
    got = ctx.plonk_check_batch(params, circuit, recs, pih, chal)
    assert (got == want).all()


def test_plonk_kernel_device_memory(svb, ctx):
    import torch
    C, params, circuit, L = plonk_setup(svb, CONFIGS["one_selector"])
    rng = np.random.default_rng(4)
    n = 64
    pih = rng.integers(0, P, size=(n, 4), dtype=np.uint64)
    proofs = [pp.prove(C, 70 + i, [int(x) for x in pih[i]]) for i in range(4)]
    grecs, gchal = to_records(L, proofs)
    recs = np.repeat(grecs, n // 4, axis=0)
    chal = np.repeat(gchal, n // 4, axis=0)
    pih = np.repeat(pih[:4], n // 4, axis=0)
    recs[9, L.off_open0 + 2 * L.n0 - 1] ^= np.uint64(1)     # last quotient opening of proof 9
    want = svb.plonk_check_host(params, circuit, recs, pih, chal)
    assert [i for i in range(n) if not bit(want, i)] == [9]
    d = [torch.from_numpy(a.view(np.int64)).cuda() for a in (recs, pih, chal)]
    d_bm = torch.zeros(n // 32, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.plonk_check_batch(params, circuit, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n_proofs=n,
                          accept_bitmap=d_bm.data_ptr(), mem=svb.MEM_DEVICE)
    ctx.synchronize()
    assert (d_bm.cpu().numpy().view(np.uint32) == want).all()
