"""Loader of tests/golden/pyref_fri.npz (made by tools/gen_golden_pyref.py from tests/pyref/ alone): per configuration the
product's FriParams / CommonData, every valid and corrupted record, and the pure-Python verifier's verdicts."""
import json
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pyref_fri.npz")


def names():
    return list(json.loads(str(np.load(PATH)["meta"])))


def load(svb, name):
    fx = np.load(PATH)
    m = json.loads(str(fx["meta"]))[name]
    params = svb.FriParams(svb.FriConfig(m["rate_bits"], m["cap_height"], m["proof_of_work_bits"], m["num_query_rounds"]), m["hiding"],
                           m["degree_bits"], m["reduction_arity_bits"], oracle_num_polys=tuple(m["oracle_num_polys"]), num_zs=m["num_zs"],
                           hash_kind=m["hash_kind"])
    common = svb.CommonData(params, num_constants=m["num_constants"], num_routed_wires=m["num_routed_wires"],
                            num_wires=m["oracle_num_polys"][1], num_challenges=m["num_zs"], num_partial_products=m["num_partial_products"],
                            quotient_degree_factor=m["quotient_degree_factor"], num_public_inputs=m["num_public_inputs"])
    base = fx[name + "_records"]
    patches, verdicts = fx[name + "_patches"], fx[name + "_verdicts"]
    recs, want = [], []
    for proof, patch, accept, code, query in verdicts:
        r = base[proof].copy()
        if patch >= 0:
            assert int(patches[patch][0]) == proof
            r[int(patches[patch][1])] = patches[patch][2]
        recs.append(r)
        want.append((int(accept), int(code), int(query)))
    return dict(params=params, common=common, meta=m, base=base, records=np.stack(recs), verdicts=want,
                blobs=fx[name + "_blobs"], circuit_digests=fx[name + "_circuit_digests"], public_inputs=fx[name + "_public_inputs"],
                vk_caps=fx[name + "_vk_caps"], proof_of=[int(v[0]) for v in verdicts])
