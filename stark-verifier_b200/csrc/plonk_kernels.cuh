// Device side of the plonk-level checks: one thread per proof runs plonk_check_one (plonk_check.hpp);
// a warp ballot packs 32 verdicts into one bitmap word.  O(1) work per proof (a few hundred Fp2 multiplications),
// reading only the record header: latency-bound and negligible beside the query phase (~3 500 permutations per proof).
#pragma once
#include "fri_kernels.cuh"
#include "plonk_check.hpp"

namespace svb {

struct PlonkRecordView {
    u32 record_words, off_open0, off_open1, off_zeta;
};

// The circuit description lives in global memory (1.6 KB, read by every thread: L1-resident).
__global__ void __launch_bounds__(128) plonk_check_kernel(const u64* __restrict__ records, PlonkRecordView V,
                                                          const sv_plonk_circuit* __restrict__ circuit,
                                                          const u64* __restrict__ pi_hashes, const u64* __restrict__ chal, u32 n,
                                                          u32* __restrict__ accept_bitmap) {
    const u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    bool ok = false;
    if (p < n) {
        const u64* rec = records + (size_t)p * V.record_words;
        const u32 nch = circuit->common.num_challenges;
        u64 pih[4];
        for (int k = 0; k < 4; k++) pih[k] = pi_hashes[4 * (size_t)p + k];
        ok = plonk_check_one(*circuit, rec + V.off_open0, rec + V.off_open1, pih, chal + 3 * (size_t)nch * p,
                             mk2(rec[V.off_zeta], rec[V.off_zeta + 1]));
    }
    const u32 word = __ballot_sync(0xFFFFFFFFu, ok);
    if ((threadIdx.x & 31) == 0 && (p >> 5) < (n + 31) / 32) accept_bitmap[p >> 5] = word;
}

// The plonk challenges of every proof: the first part of the transcript of fri_challenges_kernel (same DevChallenger),
// which squeezes and drops them on its way to zeta (plonk_verifier_chip.rs:65-103).  chal_out: n x 3*num_challenges
// words, betas | gammas | alphas.  One thread per proof, 3 permutations.
template <int KIND>
__global__ void __launch_bounds__(SVB_FS_BLOCK) plonk_challenges_kernel(const u64* __restrict__ records, FriKernelParams P, FsParams F,
                                                                        const u64* __restrict__ pi_hashes, u64* __restrict__ chal_out) {
    __shared__ u64 pscratch[PermScratch<KIND>::array_len(SVB_FS_BLOCK)];
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_proofs) return;
    const sv_fri_layout& L = P.L;
    const u64* rec = records + (size_t)p * L.record_words;
    u64* out = chal_out + 3 * (size_t)F.num_challenges * p;
    DevChallenger<KIND> ch;
#pragma unroll
    for (int i = 0; i < 12; i++) ch.st[i] = 0;
    ch.n_in = ch.n_out = 0;
    ch.scratch = pscratch;
    const u32 cap_words = L.ncap * 4;
    for (int i = 0; i < 4; i++) ch.observe(F.circuit_digest[i]);
    ch.observe_n(pi_hashes + 4 * (size_t)p, 4);
    ch.observe_n(rec + L.off_init_caps + 1 * cap_words, cap_words);                   // wires_cap
    for (u32 i = 0; i < 2 * F.num_challenges; i++) out[i] = ch.squeeze();             // plonk betas, gammas
    ch.observe_n(rec + L.off_init_caps + 2 * cap_words, cap_words);                   // zs_partial_products_cap
    for (u32 i = 0; i < F.num_challenges; i++) out[2 * F.num_challenges + i] = ch.squeeze();   // plonk alphas
}

// verdict = FRI verdict AND plonk identity; a proof that fails the identity reports SV_FAIL_PLONK (the reference checks
// it before the FRI proof, plonk_verifier_chip.rs:194-240).
__global__ void plonk_and_kernel(const u32* __restrict__ plonk_bitmap, u32* __restrict__ accept_bitmap, u32* __restrict__ first_fail, u32 n) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || ((plonk_bitmap[p >> 5] >> (p & 31)) & 1u)) return;
    atomicAnd(accept_bitmap + (p >> 5), ~(1u << (p & 31)));
    if (first_fail) first_fail[p] = SV_FAIL_PLONK;
}

}  // namespace svb
