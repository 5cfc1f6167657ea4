"""The Fiat-Shamir transcript -- TEST INFRASTRUCTURE, pure Python.

Challenger = HasherChip's duplex sponge (chip/hasher_chip.rs:51-120) behind TranscriptChip (chip/transcript_chip.rs):
`observe` buffers an element and clears the output buffer (:51-59); `squeeze` first absorbs the buffered inputs in chunks
of RATE = 8 -- overwrite the first len(chunk) state words, permute, output buffer = state[0..8] (:61-72,:107-120) -- then,
if the output buffer is empty, permutes and refills it (:80-83), and POPS FROM THE END (:84-86).

get_challenges = PlonkVerifierChip::get_challenges (chip/plonk/plonk_verifier_chip.rs:55-154)."""
from . import gl
from . import poseidon as ps

RATE = 8


class Challenger:
    def __init__(self, kind=ps.HASH_G):
        self.kind = kind
        self.state = [0] * 12
        self.absorbing = []
        self.output = []

    def clone(self):
        c = Challenger(self.kind)
        c.state, c.absorbing, c.output = list(self.state), list(self.absorbing), list(self.output)
        return c

    def observe(self, x):
        self.output = []
        self.absorbing.append(int(x))

    def observe_many(self, xs):
        for x in xs:
            self.observe(x)

    def observe_ext(self, ext):
        self.observe_many(ext)

    def observe_cap(self, cap):
        for h in cap:
            self.observe_many(h)

    def _absorb(self):
        buf, self.absorbing = self.absorbing, []
        for off in range(0, len(buf), RATE):
            chunk = buf[off:off + RATE]
            self.state[:len(chunk)] = chunk
            self.state = ps.permute(self.state, self.kind)
            self.output = list(self.state[:RATE])

    def squeeze(self, n=1):
        out = []
        for _ in range(n):
            self._absorb()
            if not self.output:
                self.state = ps.permute(self.state, self.kind)
                self.output = list(self.state[:RATE])
            out.append(self.output.pop())
        return out


def fri_openings(openings):
    """OpeningSetValues -> the two FRI batches (types/assigned.rs:26-40)"""
    zeta_batch = (list(openings["constants"]) + list(openings["plonk_sigmas"]) + list(openings["wires"])
                  + list(openings["plonk_zs"]) + list(openings["partial_products"]) + list(openings["quotient_polys"]))
    return [zeta_batch, list(openings["plonk_zs_next"])]


def get_challenges(proof, public_inputs_hash, circuit_digest, num_challenges, num_query_rounds, kind=ps.HASH_G):
    """plonk_verifier_chip.rs:55-154.  `proof`: pyref.proof.Proof.  Returns a dict of every challenge."""
    ch = Challenger(kind)
    ch.observe_many(circuit_digest)                                         # :64-67
    ch.observe_many(public_inputs_hash)                                     # :69-71
    fp = proof.opening_proof
    ch.observe_cap(proof.wires_cap)                                         # :86-90
    plonk_betas = ch.squeeze(num_challenges)                                # :91
    plonk_gammas = ch.squeeze(num_challenges)                               # :92
    ch.observe_cap(proof.plonk_zs_partial_products_cap)                     # :94-98
    plonk_alphas = ch.squeeze(num_challenges)                               # :99
    ch.observe_cap(proof.quotient_polys_cap)                                # :101-105
    plonk_zeta = tuple(ch.squeeze(2))                                       # :106
    for batch in fri_openings(proof.openings):                              # :108-114
        for ext in batch:
            ch.observe_ext(ext)
    fri_alpha = tuple(ch.squeeze(2))                                        # :117-118
    fri_betas = []
    for cap in fp.commit_phase_merkle_caps:                                 # :121-129
        ch.observe_cap(cap)
        fri_betas.append(tuple(ch.squeeze(2)))
    for ext in fp.final_poly:                                               # :131-135
        ch.observe_ext(ext)
    ch.observe(fp.pow_witness)                                              # :137
    fri_pow_response = ch.squeeze(1)[0]                                     # :138
    fri_query_indices = ch.squeeze(num_query_rounds)                        # :140-141
    return {"plonk_betas": plonk_betas, "plonk_gammas": plonk_gammas, "plonk_alphas": plonk_alphas, "plonk_zeta": plonk_zeta,
            "fri_alpha": fri_alpha, "fri_betas": fri_betas, "fri_pow_response": fri_pow_response,
            "fri_query_indices": fri_query_indices}


def zeta_next(zeta, degree_bits):
    """g * zeta with g = 7^((p-1)/2^degree_bits)  (plonk_verifier_chip.rs:219-222)"""
    return gl.e_scale(zeta, gl.root_of_unity(degree_bits))
