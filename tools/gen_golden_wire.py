#!/usr/bin/env python3
"""tests/golden/wire_small.npz: a few tiny-shape proofs in plonky2's wire format.

Bytes come from the ORACLE's cursor-style writer (oracle/wire.c), records from the oracle's reader, public-input
hashes from the pure-Python Poseidon of tools/gen_golden.py -- none of it from the product's table-driven packer,
which tests/test_wire_format.py::test_golden_wire_blob checks against this file.  The proofs themselves come from
the product's synthetic prover (the reference ships no serialised proof and its Rust prover cannot run here)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import stark_verifier_b200 as svb
    from oracle import binding as orc
    from common import P, tiny_params
    from gen_golden import hash_no_pad_py

    kw = dict(hiding=1, cap=1, degree_bits=6, rate_bits=2, queries=4, pow_bits=3)
    n, n_pi = 5, 11
    params = tiny_params(svb, **kw)
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi)
    L = svb.api.make_layout(params)
    recs = svb.synth_proofs(params, n, seed=0x601D, n_circuits=1)
    rng = np.random.default_rng(0x601D)
    pis = rng.integers(0, P, size=(n, n_pi), dtype=np.uint64)
    vk_cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    oshape, ocommon = orc.shape_from(params.to_shape()), orc.common_from(common.to_c())
    blob = np.stack([orc.wire_write_proof(oshape, ocommon, recs[i], pis[i]) for i in range(n)])
    # proof 3: the length byte of the second oracle's Merkle proof in query round 1 says one sibling too many
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + len(params.reduction_arity_bits) * 32 * L.ncap
    qbytes = (blob.shape[1] - q0 - 16 * params.final_poly_len() - 8 - 8 * n_pi) // params.config.num_query_rounds
    at = q0 + qbytes + 8 * L.leaf_len[0] + 1 + 32 * L.init_depth + 8 * L.leaf_len[1]
    assert blob[3, at] == L.init_depth
    blob[3, at] += 1
    out = dict(blob=blob, vk_cap=vk_cap, public_inputs=pis, num_public_inputs=np.uint32(n_pi),
               param_names=np.array(list(kw.keys())), param_values=np.array(list(kw.values()), dtype=np.int64))
    records, malformed = [], []
    for i in range(n):
        rc, rec, opis, _ = orc.wire_read_proof(oshape, ocommon, vk_cap, blob[i])
        assert rc in (0, 1) and (opis == pis[i]).all()
        records.append(rec)
        malformed.append(rc)
    out["records"] = np.stack(records)
    out["malformed"] = np.array(malformed, dtype=np.uint8)
    out["pi_hashes"] = np.array([hash_no_pad_py([int(x) for x in pis[i]]) for i in range(n)], dtype=np.uint64)
    assert list(out["malformed"]) == [0, 0, 0, 1, 0]
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "wire_small.npz"), **out)
    print("wrote wire_small.npz:", blob.shape, "bytes per proof", blob.shape[1])


if __name__ == "__main__":
    main()
