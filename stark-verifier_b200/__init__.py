"""stark-verifier_b200 -- B200-native batch verifier for the FRI query phase of plonky2 proofs.

Python face of the C ABI in ``include/stark_verifier_b200.h`` (ctypes; no torch types cross the
boundary).  Names mirror the reference's ``plonky2_verifier`` data model
(``types/common_data.rs``: ``FriConfig``/``FriParams``; ``chip/fri_chip.rs``: ``FriVerifierChip``
with ``verify_fri_proof``).  The package directory name contains a hyphen; import it with
``importlib.import_module("stark-verifier_b200")`` or through the ``stark_verifier_b200`` shim at
the repository root.

There is no CPU fallback: creating a :class:`Context` without a CUDA device raises.
"""
from . import shard  # noqa: F401
from .api import (  # noqa: F401
    Context,
    FriConfig,
    FriParams,
    FriShape,
    FriVerifierChip,
    Layout,
    SvError,
    SHAPE_A,
    SHAPE_B,
    SHAPE_SEMAPHORE,
    SHAPE_OUTER_BN254,
    HASH_POSEIDON_GOLDILOCKS,
    HASH_POSEIDON_BN254,
    FAIL_NAMES,
    fri_challenges,
    lib,
    lib_path,
    version,
    circuit_from_common_data,
    constant_arity_bits,
    synth_proofs,
    synth_public_inputs,
    MEM_HOST,
    MEM_DEVICE,
    CommonData,
    PlonkCommon,
    FAIL_MALFORMED,
    FAIL_PLONK,
    shape_from_common,
    wire_proof_bytes,
    wire_pack,
    wire_unpack_batch,
    public_inputs_hash,
    PlonkCircuit,
    PlonkGate,
    GATE_NOOP,
    GATE_CONSTANT,
    GATE_PUBLIC_INPUT,
    GATE_ARITHMETIC,
    GATE_ARITHMETIC_EXT,
    GATE_MUL_EXT,
    GATE_BASE_SUM,
    GATE_REDUCING,
    GATE_REDUCING_EXT,
    GATE_RANDOM_ACCESS,
    GATE_POSEIDON_MDS,
    GATE_POSEIDON,
    make_plonk_circuit,
    plonk_challenges,
    plonk_gate_from_id,
    plonk_check_host,
    ntt_host,
    verify_batch,
    PlonkVerifierChip,
    lde_host,
)
