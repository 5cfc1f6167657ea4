#!/usr/bin/env python3
"""tests/golden/pyref_fri.npz -- fixtures made by the INDEPENDENT pure-Python implementation alone (tests/pyref/: prover,
transcript, verifier, wire writer; it imports nothing from oracle/ or stark-verifier_b200/).

For every configuration: a few valid proofs as flat records (include/stark_verifier_b200.h) and as plonky2 wire bytes, the
transcript inputs (circuit digest, public inputs, verifier-key cap), seeded corruptions stored as (proof, word, value) patches,
and for every valid / corrupted record the verdict of tests/pyref/fri.py: (accept, first-failure code, query round).
The C oracle (CPU suite) and the CUDA library (-m gpu suite) must reproduce every verdict bit for bit.

    python tools/gen_golden_pyref.py            # tiny shapes + BASELINE shape A (about 5 minutes)
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pyref import challenger, fri, gl, proof as pf, prover  # noqa: E402

P = gl.P
# name -> (degree_bits, rate_bits, cap_height, pow_bits, queries, reduction_arity_bits, widths, num_zs, hiding, hash_kind, n_valid)
CONFIGS = {
    "arity2": (7, 3, 2, 4, 6, [1, 1], (9, 11, 6, 4), 2, False, 0, 3),
    "arity4_salted": (6, 2, 1, 3, 5, [2], (7, 9, 4, 4), 2, True, 0, 3),
    "arity8": (8, 3, 2, 4, 6, [3], (9, 11, 6, 4), 2, False, 0, 3),          # ConstantArityBits(3, 5) on 2^8
    "arity16": (9, 2, 0, 2, 5, [4], (6, 7, 4, 2), 2, False, 0, 2),          # cap_height 0: the cap is the root
    "arity_mixed": (9, 3, 3, 4, 4, [1, 3, 2], (9, 11, 6, 4), 2, True, 0, 2),
    "arity8_hash_b": (6, 1, 1, 2, 3, [3], (5, 6, 4, 2), 2, False, 1, 1),    # Poseidon-BN254 (outer wrapped-proof hasher)
    "no_steps": (4, 3, 1, 2, 5, [], (6, 5, 4, 4), 2, False, 0, 2),          # trace shorter than the final polynomial
    # BASELINE configs[1] and the semaphore demo's reduction strategy on the same trace
    "shape_a": (12, 3, 4, 16, 28, [1] * 7, (84, 135, 20, 16), 2, False, 0, 1),
    "shape_a_arity8": (12, 3, 4, 16, 28, [3, 3, 3], (84, 135, 20, 16), 2, False, 0, 1),
}
KINDS = ["sibling", "leaf", "step_eval", "final_poly", "pow", "noncanonical", "step_sibling", "cap", "opening", "step_eval_other",
         "alpha", "index"]


def params_of(cfg):
    d, r, c, pw, q, ab, widths, nz, hiding, kind, _ = cfg
    return pf.FriParams(d, r, c, pw, q, ab, list(widths), nz, hiding=hiding, hash_kind=kind)


def corruptions(params, rec, rng):
    """-> list of (kind, word offset, new value): one word changed per corruption, every check of the verifier hit once"""
    L = pf.RecordLayout(params)
    S, Q = len(params.reduction_arity_bits), params.num_query_rounds
    out = []
    for kind in KINDS:
        q = int(rng.integers(0, Q))
        qb = L.header_words + q * L.query_words
        o = int(rng.integers(0, 4))
        flip = 1 << int(rng.integers(0, 20))
        at = None
        if kind == "sibling" and L.init_depth:
            at = qb + L.q_off_init_sibs[o] + int(rng.integers(0, 4 * L.init_depth))
        elif kind == "leaf":
            at = qb + L.q_off_init_evals[o] + int(rng.integers(0, params.leaf_len(o)))
        elif kind == "step_eval" and S:
            # the entry the consistency check reads: evals[x_index_within_coset] of step 0
            idx = rec[L.off_indices + q] & ((1 << params.lde_bits()) - 1)
            within = idx & ((1 << params.reduction_arity_bits[0]) - 1)
            at = qb + L.q_off_step_evals[0] + 2 * within + int(rng.integers(0, 2))
        elif kind == "step_eval_other" and S:
            # another entry of the coset: consistency holds, the fold and the step Merkle proof do not
            st = int(rng.integers(0, S))
            shift = sum(params.reduction_arity_bits[:st])
            idx = (rec[L.off_indices + q] & ((1 << params.lde_bits()) - 1)) >> shift
            within = idx & ((1 << params.reduction_arity_bits[st]) - 1)
            other = within ^ 1
            at = qb + L.q_off_step_evals[st] + 2 * other
        elif kind == "final_poly":
            at = L.off_final_poly + int(rng.integers(0, 2 * params.final_poly_len()))
        elif kind == "pow":
            out.append((kind, L.off_pow_response, rec[L.off_pow_response] | (1 << 63) if rec[L.off_pow_response] | (1 << 63) < P
                        else (1 << 63)))
            continue
        elif kind == "noncanonical":
            at = qb + L.q_off_init_evals[o]
            out.append((kind, at, P + int(rng.integers(0, 1000))))
            continue
        elif kind == "step_sibling" and S and L.step_depth[0]:
            at = qb + L.q_off_step_sibs[0] + int(rng.integers(0, 4 * L.step_depth[0]))
        elif kind == "cap":
            idx = rec[L.off_indices + q] & ((1 << params.lde_bits()) - 1)
            ci = idx >> (params.lde_bits() - params.cap_height)
            at = L.off_init_caps + (o * L.ncap + ci) * 4 + int(rng.integers(0, 4))
        elif kind == "opening":
            at = L.off_open0 + int(rng.integers(0, 2 * L.n0))
        elif kind == "alpha":
            at = L.off_alpha
        elif kind == "index":
            at = L.off_indices + q
        if at is None:
            continue
        v = rec[at] ^ flip
        if v >= P:
            v = rec[at] ^ 1
        out.append((kind, at, v))
    return out


def main():
    only = sys.argv[1:] or list(CONFIGS)
    path = os.path.join(ROOT, "tests", "golden", "pyref_fri.npz")
    fx = dict(np.load(path)) if os.path.exists(path) and sys.argv[1:] else {}
    meta = json.loads(str(fx["meta"])) if "meta" in fx else {}
    for name in only:
        cfg = CONFIGS[name]
        params = params_of(cfg)
        n_valid = cfg[-1]
        common = pf.Common.for_widths(params.oracle_num_polys, params.num_zs, 3)
        rng = np.random.default_rng(sum(map(ord, name)))
        recs, blobs, cds, pis, vks, patches, verdicts = [], [], [], [], [], [], []
        t0 = time.time()
        for i in range(n_valid):
            proof, vk_cap, cd, pr = prover.prove_random(params, common, seed=1000 * sum(map(ord, name)) + i)
            ch = challenger.get_challenges(proof, pr.pi_hash, cd, common.num_challenges, params.num_query_rounds, params.hash_kind)
            assert all(ch[k] == v for k, v in pr.challenges.items())
            zn = challenger.zeta_next(ch["plonk_zeta"], params.degree_bits)
            rec = pf.to_record(params, proof, vk_cap, ch, zn)
            v = fri.verify_record(params, common, rec)
            assert v == (True, 0, 0), (name, i, v)
            # the reader inverts the writer
            blob = pf.write_proof(proof)
            assert pf.write_proof(pf.read_proof(blob, params, common)) == blob
            recs.append(rec); blobs.append(np.frombuffer(blob, dtype=np.uint8)); cds.append(cd); pis.append(proof.public_inputs)
            vks.append([w for h in vk_cap for w in h])
            verdicts.append((i, -1, 1, 0, 0))
            for j, (kind, at, val) in enumerate(corruptions(params, rec, rng)):
                bad = list(rec)
                bad[at] = val
                ok, code, q = fri.verify_record(params, common, bad)
                patches.append((i, at, val))
                verdicts.append((i, len(patches) - 1, int(ok), code, q))
                print(f"  {name}[{i}] {kind:16s} -> accept={int(ok)} code={code} query={q}")
        print(f"{name}: {n_valid} proofs, {len(patches)} corruptions, {time.time() - t0:.1f} s")
        fx[name + "_records"] = np.array(recs, dtype=np.uint64)
        fx[name + "_blobs"] = np.stack(blobs)
        fx[name + "_circuit_digests"] = np.array(cds, dtype=np.uint64)
        fx[name + "_public_inputs"] = np.array(pis, dtype=np.uint64)
        fx[name + "_vk_caps"] = np.array(vks, dtype=np.uint64)
        fx[name + "_patches"] = np.array(patches, dtype=np.uint64).reshape(-1, 3)
        fx[name + "_verdicts"] = np.array(verdicts, dtype=np.int64)      # (proof, patch or -1, accept, code, query)
        meta[name] = dict(degree_bits=cfg[0], rate_bits=cfg[1], cap_height=cfg[2], proof_of_work_bits=cfg[3], num_query_rounds=cfg[4],
                          reduction_arity_bits=cfg[5], oracle_num_polys=list(cfg[6]), num_zs=cfg[7], hiding=cfg[8], hash_kind=cfg[9],
                          num_public_inputs=3, num_constants=common.num_constants, num_routed_wires=common.num_routed_wires,
                          num_partial_products=common.num_partial_products, quotient_degree_factor=common.quotient_degree_factor)
    fx["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(path, **fx)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
