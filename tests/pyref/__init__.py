"""TEST INFRASTRUCTURE -- an independent, pure-Python/numpy restatement of the verifier-side path (and of the prover that
feeds it), written from the reference's Rust sources and plonky2's published algorithms:

    gl.py          Goldilocks F_p and F_p^2 (Python integers; exact uint64 numpy vectors for batches)
    poseidon.py    both hash families (naive round definition), sponge, two-to-one, hash_or_noop
    merkle.py      Merkle trees with caps: build, open, verify-to-cap
    challenger.py  the duplex challenger and PlonkVerifierChip::get_challenges
    fri.py         FriVerifierChip::verify_fri_proof with plonky2's general 2^k-ary fold
    proof.py       proof containers, plonky2's wire bytes (writer + reader), the flat record of include/stark_verifier_b200.h
    prover.py      FRI opening prover (commit phase, PoW, queries) for random or given oracle polynomials

It imports NOTHING from oracle/ or stark-verifier_b200/ (tests/test_pyref.py greps for that): the C oracle and the CUDA
library are checked AGAINST it, and the committed fixtures under tests/golden/ are produced by it alone
(tools/gen_golden_pyref.py).  Parameter tables: tests/pyref/constants.json, parsed from the reference's .rs files by
tools/gen_pyref_constants.py.  Citations `file:line` are relative to /root/reference/src/plonky2_verifier/."""
