// Device side of the plonk-level checks: one thread per proof runs plonk_check_one (plonk_check.hpp);
// a warp ballot packs 32 verdicts into one bitmap word.  O(1) work per proof (a few hundred Fp2 multiplications),
// reading only the record header: latency-bound and negligible beside the query phase (~3 500 permutations per proof).
#pragma once
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

}  // namespace svb
