#!/usr/bin/env python3
"""tests/pyref/constants.json: the numeric parameter tables of both hash families, parsed STRAIGHT from the reference's
Rust sources (chip/plonk/gates/poseidon.rs:26-124,321-322 and bn245_poseidon/constants.rs:5-384).  Data only: the
pure-Python reference under tests/pyref/ reads this file and nothing from oracle/ or stark-verifier_b200/."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/plonky2_verifier"


def main():
    t = open(f"{REF}/chip/plonk/gates/poseidon.rs").read()

    def table(name):
        body = re.search(r"const\s+" + name + r"\s*:[^=]*=\s*\[(.*?)\];", t, re.S).group(1)
        body = re.sub(r"//[^\n]*", "", body)
        return [int(x, 16) if x.startswith("0x") else int(x) for x in re.findall(r"0x[0-9a-fA-F]+|\b\d+\b", body)]
    fast = {name: table(name) for name in ("FAST_PARTIAL_FIRST_ROUND_CONSTANT", "FAST_PARTIAL_ROUND_CONSTANTS", "FAST_PARTIAL_ROUND_VS",
                                           "FAST_PARTIAL_ROUND_W_HATS", "FAST_PARTIAL_ROUND_INITIAL_MATRIX")}
    assert [len(v) for v in fast.values()] == [12, 22, 242, 242, 121], [len(v) for v in fast.values()]
    rc = table("ALL_ROUND_CONSTANTS")
    circ, diag = table("MDS_MATRIX_CIRC"), table("MDS_MATRIX_DIAG")
    assert len(rc) == 360 and len(circ) == 12 and len(diag) == 12, (len(rc), len(circ), len(diag))
    b = open(f"{REF}/bn245_poseidon/constants.rs").read()
    brc = re.findall(r'"0x([0-9a-fA-F]+)"', re.search(r"ROUND_CONSTANTS_STR[^=]*=\s*\[(.*?)\];", b, re.S).group(1))
    bmds = re.findall(r'"0x([0-9a-fA-F]+)"', re.search(r"MDS_MATRIX_STR[^=]*=\s*\[(.*?)\];\s*\n\s*fn ", b, re.S).group(1))
    assert len(brc) == 340 and len(bmds) == 25
    out = {"source": "tools/gen_pyref_constants.py from /root/reference (gates/poseidon.rs, bn245_poseidon/constants.rs)",
           "g_round_constants": [f"{x:016x}" for x in rc], "g_mds_circ": circ, "g_mds_diag": diag,
           "b_round_constants": brc, "b_mds": bmds,
           # the "fast partial round" tables (gates/poseidon.rs:127-319): only the PoseidonGate witness of tests/plonk_prover.py reads them
           **{"g_" + k.lower(): [f"{x:016x}" for x in v] for k, v in fast.items()}}
    json.dump(out, open(os.path.join(ROOT, "tests", "pyref", "constants.json"), "w"), indent=0)
    print("wrote tests/pyref/constants.json")


if __name__ == "__main__":
    main()
