"""The CUDA path against fixtures that NO code of this repository's product or C oracle produced: tests/golden/pyref_fri.npz
comes from the pure-Python prover + verifier of tests/pyref/ (tools/gen_golden_pyref.py).  Every record -- valid, or with one
corrupted word -- must get the accept bit AND the first-failure code / query round the Python verifier gave it, through
sv_fri_verify_batch, through the device transcript (sv_fri_verify_batch_fs) and from wire bytes (sv_verify_proofs_wire);
2^k-ary reductions (k = 1..4, mixed), salted leaves, both hash families, no reduction steps, BASELINE shape A."""
import numpy as np
import pytest

import golden_pyref as gp
from common import bit

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", gp.names())
def test_cuda_reproduces_the_python_verdicts(svb, ctx, name):
    G = gp.load(svb, name)
    params = G["params"]
    recs = np.ascontiguousarray(G["records"])
    n = recs.shape[0]
    bm, ff = ctx.fri_verify_batch(params, recs, want_fail=True)
    for i, (accept, code, query) in enumerate(G["verdicts"]):
        assert bit(bm, i) == accept, (name, i)
        assert int(ff[i]) == (0 if accept else (query << 8) | code), (name, i, hex(int(ff[i])), code, query)
    assert n > len(G["base"])           # corrupted copies are in the batch


@pytest.mark.parametrize("name", gp.names())
def test_device_transcript_reproduces_the_python_challenges(svb, ctx, name):
    """challenge fields stripped -> sv_fri_challenges_batch == the Python challenger's; verdicts through sv_fri_verify_batch_fs"""
    G = gp.load(svb, name)
    params, L = G["params"], svb.api.make_layout(G["params"])
    nch = G["meta"]["num_zs"]
    for i, rec in enumerate(G["base"]):       # one circuit digest per golden proof
        pih = svb.public_inputs_hash(G["public_inputs"][i])[None, :]
        stripped = rec[None, :].copy()
        stripped[:, L.off_alpha:L.header_words] = 0
        filled = ctx.fri_challenges_batch(params, stripped.copy(), G["circuit_digests"][i], pih, num_challenges=nch)
        assert (filled[0] == rec).all()
        bm = ctx.fri_verify_batch_fs(params, stripped, G["circuit_digests"][i], pih, num_challenges=nch)
        assert bit(bm, 0) == 1


@pytest.mark.parametrize("name", gp.names())
def test_wire_bytes_of_the_python_writer_on_the_device(svb, ctx, name):
    """bytes written by tests/pyref/proof.py -> H2D -> device unpack + public-inputs hash + transcript + query phase"""
    G = gp.load(svb, name)
    common = G["common"]
    for i in range(G["blobs"].shape[0]):
        blob = np.ascontiguousarray(G["blobs"][i:i + 1])
        bm, ff = ctx.verify_proofs_wire(common, G["vk_caps"][i], G["circuit_digests"][i], blob, want_fail=True)
        assert bit(bm, 0) == 1 and int(ff[0]) == 0
        bad = blob.copy()
        bad[0, 40] ^= 1                      # a byte of the wires cap: other challenges, and the wires tree no longer opens
        bm = ctx.verify_proofs_wire(common, G["vk_caps"][i], G["circuit_digests"][i], bad)
        assert bit(bm, 0) == 0
