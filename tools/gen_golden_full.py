#!/usr/bin/env python3
"""tests/golden/full_proof_toy.npz: wire bytes of COMPLETE proofs (plonk identity + FRI) of the toy circuit with the
reference's recursion gate set (tests/full_prover.py: pure-Python algebra, hashing and transcript through the CPU
oracle), plus corrupted copies, labelled by the CPU side of the verifier."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import stark_verifier_b200 as svb
    from oracle import binding as orc
    import test_full_proof as T
    name = "recursion_gate_set"
    B = T.build(svb, orc, name, 2, seed=0x601D, n_pi=5)
    blob = np.concatenate([B["blob"], B["blob"][:1], B["blob"][1:]])
    blob[2, 3 * 32 * B["L"].ncap + 100] ^= 1          # an opening of proof 0
    blob[3, -20] ^= 4                                 # a public input of proof 1
    fri, pl, opl, mal, _ = T.cpu_verdicts(svb, orc, B, blob)
    accept = np.array([int(f and p and not m) for f, p, m in zip(fri, pl, mal)], dtype=np.uint8)
    assert list(accept) == [1, 1, 0, 0] and pl == opl
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "full_proof_toy.npz"), blob=blob, accept=accept, config=name,
                        num_public_inputs=np.uint32(5), circuit_digest=B["cd"], vk_cap=B["vk_cap"])
    print("wrote full_proof_toy.npz", blob.shape)


if __name__ == "__main__":
    main()
