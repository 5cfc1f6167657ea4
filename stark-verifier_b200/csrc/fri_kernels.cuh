// sm_100a kernels of the FRI query-phase verifier.
//
// Work decomposition.  A (proof x query) unit of check_consistency (chip/fri_chip.rs:228-327) is
// 4 + num_steps INDEPENDENT Merkle chains (each: optional leaf sponge, then one two-to-one
// permutation per level, then a cap compare -- chip/merkle_proof_chip.rs:39-87) plus one short
// algebra chain (DEEP quotient, folds, final polynomial).  Only the algebra chain is sequential
// across steps, and it never needs a hash output: the step evals it consumes are proof data, and
// their authenticity is exactly what the step chains check.  So every chain is its own work item:
//
//      work item = (class, unit)      class in { init oracle 0..3, step 0..S-1, algebra }
//
// One thread per work item, 32 consecutive units of the SAME class per warp, so a warp runs one
// instruction stream with no divergence (all lanes hash the same number of levels).  Blocks are
// ordered heaviest class first so the tail of the grid is made of the short chains.  A failing item
// clears its proof's accept bit with one atomicAnd (the rare path).
//
// Memory.  Records are read straight from global memory with 128-bit read-only loads: a lane reads
// the 32 B digest of its own query block, i.e. each request touches 32 distinct sectors, all fully
// used (no over-fetch).  The path is bound by the integer multiplier (~1.7e4 instructions per 32 B
// sibling), so the records are not staged through shared memory: there is no reuse to exploit, and the
// next level's sibling is prefetched while the current permutation runs.
#pragma once
#include "layout.hpp"
#include "fri_fold.cuh"
#include "poseidon_g.cuh"
#include "poseidon_b.cuh"
#include "poseidon_g_coop2.cuh"

namespace svb {

#define SVB_BLOCK 128   // threads per block of every kernel; also the stride of the per-thread smem scratch
#ifndef SVB_MINBLOCKS
#define SVB_MINBLOCKS 6  // __launch_bounds__ second argument of the Poseidon-Goldilocks kernels (register cap = 65536 / (128 * N) = 80): measured 4 -> 423.7 k, 5 -> 429.1 k, 6 -> 432.1 k proofs/s
#endif
// Poseidon-BN254 kernels want ~130 registers: capping them at 126 costs 6 % (spills), so they keep 3 blocks per SM
#define SVB_MINBLOCKS_K(KIND) ((KIND) == SV_HASH_POSEIDON_BN254 ? 3 : SVB_MINBLOCKS)

struct FriKernelParams {
    sv_fri_layout L;
    u32 num_queries, num_steps, final_poly_len, pow_bits, hash_kind;
    u32 oracle_num_polys[4];
    u32 num_zs;
    u32 n_proofs;
    u32 n_units;          // n_proofs * num_queries
    u32 blocks_per_class; // ceil(n_units / block)
    u32 group_blocks;     // grid phase A: blocks of one class per unit group (group-major, then class, then block)
    u32 n_groups;         // ceil(blocks_per_class / group_blocks)
    u32 n_classes_a;      // classes of phase A: the 4 oracle trees (heaviest first) + the algebra chain
    u32 n_classes;        // 4 + num_steps + 1
    u32 class_order[SV_MAX_STEPS + 5];  // phase A classes, then the step trees (deepest first)
    u32 n_leaf_classes;   // oracle trees whose leaves are hashed (leaf_len > 4), heaviest first: the grid of fri_leaf_kernel
    u32 leaf_class_order[4];
    u64 omega_pow2[40];   // omega^(2^i), omega = 7^((p-1)/2^lde_bits)
};

// Shared-memory scratch of one permutation, in u64 words per thread: Poseidon-BN254 stages its whole 5 x 8-limb
// state; Poseidon-Goldilocks needs none in the naive-round form (11 words for the outputs of the initial matrix in
// the sparse form, SVB_PARTIAL_NAIVE = 0).
template <int KIND> struct PermScratch {
    static constexpr int words = KIND == SV_HASH_POSEIDON_BN254 ? SVB_B_SMEM_WORDS / 2 : (SVB_PARTIAL_NAIVE ? 0 : 11);
    static constexpr int array_len(int threads) { return words ? words * threads : 1; }   // a zero-length array is not allowed
};
// The width-12 permutation of hash family KIND on `s` (LOOSE in; G leaves LOOSE words, B canonical ones).
// scratch: base of the block's shared array, stride = threads per block.
template <int KIND>
SVB_D void permute_dev(u64 s[12], u64* scratch, u32 stride) {
    if (KIND == SV_HASH_POSEIDON_BN254) poseidon_b_dev(s, reinterpret_cast<u32*>(scratch) + threadIdx.x, stride);
    else poseidon_g_dev(s, scratch + threadIdx.x, stride);
}

SVB_D void ldg4(const u64* p, u64 out[4]) {
    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p);
    ulonglong2 a = __ldg(q), b = __ldg(q + 1);
    out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
}

// Pull the 32-byte sector at p towards the SM while the current permutation runs (no register cost).
// d_PREFETCH_MODE: 1 = into L1 (default), 2 = into L2 only, 0 = off (lab knob, env SVB_PREFETCH at sv_ctx_create).
__constant__ u32 d_PREFETCH_MODE = 1;
SVB_D void prefetch32(const u64* p) {
    const u32 mode = d_PREFETCH_MODE;
    if (mode == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    else if (mode == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

SVB_D void report_fail(u32* accept_bitmap, u32* first_fail, u32 proof, u32 query, u32 order_key, u32 code) {
    atomicAnd(&accept_bitmap[proof >> 5], ~(1u << (proof & 31)));
    // order key: query-major, then the reference's check order inside the round, code in the low byte
    if (first_fail) atomicMin(&first_fail[proof], (query << 20) | (order_key << 8) | code);
}

// One Merkle chain: hash_or_noop(leaf) then `depth` two-to-one levels, compare with cap[cap_index].
// Returns 0 ok, SV_FAIL_NONCANONICAL, or `merkle_code`.
// Written as ONE loop over "absorb a block, permute" so the kernel holds a single inlined copy of
// the permutation: iterations [0, n_sponge) overwrite the rate lanes with the next leaf chunk
// (overwrite-mode sponge, rate 8: hasher_chip.rs:122-148), iterations [n_sponge, n_sponge+depth)
// compress with the sibling of that level on a fresh capacity (merkle_proof_chip.rs:58-71).
template <int KIND>
SVB_D u32 merkle_chain(const u64* __restrict__ leaf, u32 leaf_len, const u64* __restrict__ sibs, u32 depth,
                       u64 index, const u64* __restrict__ cap_entry, u32 merkle_code, u64* scratch) {
    u64 s[12];
    bool canon_ok = true;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
    u32 n_sponge = 0;
    if (leaf_len <= 4) {
        // hash_or_noop: the leaf IS the digest (merkle_proof_chip.rs:52-53); segments are 4-word padded
        u64 w[4];
        ldg4(leaf, w);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            s[i] = (u32)i < leaf_len ? w[i] : 0;
            canon_ok &= is_canonical(s[i]);
        }
    } else {
        n_sponge = (leaf_len + 7) >> 3;
    }
    const u32 n_iter = n_sponge + depth;
#pragma unroll 1
    for (u32 it = 0; it < n_iter; it++) {
        if (it < n_sponge) {
            u32 off = it * 8, rem = leaf_len - off;
            u64 w[8];
            ldg4(leaf + off, w);
            if (rem > 4) ldg4(leaf + off + 4, w + 4);
#pragma unroll
            for (int i = 0; i < 8; i++)
                if ((u32)i < rem) {
                    s[i] = w[i];
                    canon_ok &= is_canonical(w[i]);
                }
        } else {
            u32 lvl = it - n_sponge;
            u64 sib[4];
            ldg4(sibs + 4 * lvl, sib);
            bool bit = (index >> lvl) & 1;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                canon_ok &= is_canonical(sib[i]);
                u64 cur = s[i];
                s[i] = bit ? sib[i] : cur;      // select(sibling, state, bit)  (merkle_proof_chip.rs:61-69)
                s[i + 4] = bit ? cur : sib[i];
            }
#pragma unroll
            for (int i = 8; i < 12; i++) s[i] = 0;  // fresh hasher per level (:59)
        }
        // the next iteration's input (leaf chunk or sibling) is fetched while this permutation runs
        if (it + 1 < n_sponge) {
            prefetch32(leaf + (it + 1) * 8);
            if ((it + 1) * 8 + 4 < leaf_len) prefetch32(leaf + (it + 1) * 8 + 4);
        } else if (it + 1 < n_iter) {
            prefetch32(sibs + 4 * (it + 1 - n_sponge));
        }
        permute_dev<KIND>(s, scratch, SVB_BLOCK);
    }
    u64 c[4];
    ldg4(cap_entry, c);
    bool eq = true;
#pragma unroll
    for (int i = 0; i < 4; i++) eq &= (canon(s[i]) == c[i]);
    if (!canon_ok) return SV_FAIL_NONCANONICAL;
    return eq ? 0u : merkle_code;
}

// reduce_extension with base-field terms: acc = acc*alpha + (e, 0), from the last term
// (goldilocks_extension_chip.rs:331-355)
SVB_D fp2 horner_base(fp2 alpha, const u64* __restrict__ terms, u32 n, fp2 acc) {
    for (u32 i = n; i-- > 0;) {
        fp2 t = mul2(acc, alpha);
        acc = mk2(add(t.c0, __ldg(terms + i)), t.c1);
    }
    return acc;
}

// Per-proof preparation: range-check the header, proof-of-work check, reduced openings.
// One WARP per proof, 32 proofs (one accept-bitmap word) per block.  compute_reduced_openings
// (fri_chip.rs:58-70) is sum_i opening_i * alpha^i: lane j folds its contiguous chunk of c = ceil(n/32)
// openings by Horner, scales it by (alpha^c)^j and the warp adds the 32 partial sums with shuffles --
// exact field arithmetic, so the value is the one the reference's sequential Horner gives, at 1/10 of the
// latency of one thread per proof (this kernel was 1 % of the step).
#define SVB_PREP_BLOCK 1024
SVB_D fp2 warp_sum2(fp2 v) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        fp2 o;
        o.c0 = ((u64)__shfl_xor_sync(0xFFFFFFFFu, (u32)(v.c0 >> 32), off) << 32) | __shfl_xor_sync(0xFFFFFFFFu, (u32)v.c0, off);
        o.c1 = ((u64)__shfl_xor_sync(0xFFFFFFFFu, (u32)(v.c1 >> 32), off) << 32) | __shfl_xor_sync(0xFFFFFFFFu, (u32)v.c1, off);
        v = add2(v, o);
    }
    return v;
}
// sum_{i<n} terms[i] * alpha^i over the warp (terms = n Fp2 values, 2 words each)
SVB_D fp2 warp_reduce_openings(const u64* __restrict__ terms, u32 n, fp2 alpha, u32 lane) {
    const u32 c = (n + 31) / 32;
    const u32 lo = lane * c, hi = lo + c < n ? lo + c : n;
    fp2 acc = mk2(0, 0);
    for (u32 i = hi; i-- > lo && lo < n;) {
        ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(terms + 2 * i));
        acc = add2(mul2(acc, alpha), mk2(v.x, v.y));
    }
    fp2 ac = mk2(1, 0);                       // alpha^c
    for (u32 i = 0; i < c; i++) ac = mul2(ac, alpha);
    fp2 scale = mk2(1, 0);                    // (alpha^c)^lane by square and multiply
    for (u32 bit = 0; bit < 5; bit++) {
        if ((lane >> bit) & 1) scale = mul2(scale, ac);
        ac = mul2(ac, ac);
    }
    return warp_sum2(mul2(acc, scale));
}
__global__ void __launch_bounds__(SVB_PREP_BLOCK) fri_prepare_kernel(const u64* __restrict__ records, FriKernelParams P,
                                                               u64* __restrict__ scratch, u32* __restrict__ accept_bitmap,
                                                               u32* __restrict__ first_fail) {
    __shared__ u32 ok_flags[32];
    const u32 lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 p = blockIdx.x * 32 + warp;
    bool ok = false;
    if (p < P.n_proofs) {          // warp-uniform
        const sv_fri_layout& L = P.L;
        const u64* rec = records + (size_t)p * L.record_words;
        bool canon_lane = true;
        for (u32 w = 2 * lane; w < L.header_words; w += 64) {
            ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(rec + w));
            canon_lane &= is_canonical(v.x) & is_canonical(v.y);
        }
        const bool canon_ok = __all_sync(0xFFFFFFFFu, canon_lane);
        // fri_verify_proof_of_work (fri_chip.rs:364-376)
        u64 resp = rec[L.off_pow_response];
        bool pow_ok = P.pow_bits == 0 || (resp >> (64 - P.pow_bits)) == 0;
        fp2 alpha = mk2(rec[L.off_alpha], rec[L.off_alpha + 1]);
        fp2 ro0 = mk2(0, 0), ro1 = mk2(0, 0);
        if (canon_ok) {            // non-canonical words would break the canonical-input contract of add2/sub2
            ro0 = warp_reduce_openings(rec + L.off_open0, L.n0, alpha, lane);
            ro1 = warp_reduce_openings(rec + L.off_open1, L.n1, alpha, lane);
        }
        ok = canon_ok && pow_ok;
        if (lane == 0) {
            scratch[4 * (size_t)p + 0] = ro0.c0; scratch[4 * (size_t)p + 1] = ro0.c1;
            scratch[4 * (size_t)p + 2] = ro1.c0; scratch[4 * (size_t)p + 3] = ro1.c1;
            if (first_fail) first_fail[p] = ok ? 0xFFFFFFFFu : (canon_ok ? SV_FAIL_POW : SV_FAIL_NONCANONICAL);
        }
    }
    if (lane == 0) ok_flags[warp] = ok;
    __syncthreads();
    if (warp == 0) {
        u32 word = __ballot_sync(0xFFFFFFFFu, ok_flags[lane] != 0);
        if (lane == 0 && blockIdx.x < (P.n_proofs + 31) / 32) accept_bitmap[blockIdx.x] = word;
    }
}

// The challenge-INDEPENDENT part of a query round: the leaf sponges of the four oracle trees (hash_or_noop of the
// opened evaluations, chip/merkle_proof_chip.rs:52-53 -- 34 of the 126 permutations of a shape-A query round).  They need
// neither the query index nor any other challenge, so when the Fiat-Shamir transcript runs on the device (~155
// DEPENDENT permutations per proof, 2-3 ms of latency that nothing else of the query phase can overlap) these digests
// CAN be computed beside it, as soon as the bytes are in HBM (opt-in, SVB_LEAF_SPLIT=1: on B200 the contention with the
// transcript's warps costs more than the overlap gains, see capi.cu); fri_query_kernel then starts its oracle-tree chains from
// leaf_digests instead of hashing the leaf (same words, same canonical-range check: a non-canonical evaluation is
// passed on as an all-ones digest, which the chain reports as SV_FAIL_NONCANONICAL exactly as before).
// One thread per (hashed oracle tree, unit), class-major, heaviest first.  leaf_digests: [unit][4 trees][4 words].
template <int KIND>
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS_K(KIND)) fri_leaf_kernel(const u64* __restrict__ records, FriKernelParams P,
                                                                                   u64* __restrict__ leaf_digests) {
    __shared__ u64 pscratch[PermScratch<KIND>::array_len(SVB_BLOCK)];
    const sv_fri_layout& L = P.L;
    const u32 cls = P.leaf_class_order[blockIdx.x / P.blocks_per_class], blk = blockIdx.x % P.blocks_per_class;
    const u32 unit = blk * blockDim.x + threadIdx.x;
    if (unit >= P.n_units) return;
    const u32 proof = unit / P.num_queries, query = unit - proof * P.num_queries;
    const u64* leaf = records + (size_t)proof * L.record_words + L.header_words + (size_t)query * L.query_words + L.q_off_init_evals[cls];
    const u32 leaf_len = L.leaf_len[cls], n_sponge = (leaf_len + 7) >> 3;
    u64 s[12];
    bool canon_ok = true;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = 0;
#pragma unroll 1
    for (u32 it = 0; it < n_sponge; it++) {
        const u32 off = it * 8, rem = leaf_len - off;
        u64 w[8];
        ldg4(leaf + off, w);
        if (rem > 4) ldg4(leaf + off + 4, w + 4);
#pragma unroll
        for (int i = 0; i < 8; i++)
            if ((u32)i < rem) {
                s[i] = w[i];
                canon_ok &= is_canonical(w[i]);
            }
        if (it + 1 < n_sponge) {
            prefetch32(leaf + (it + 1) * 8);
            if ((it + 1) * 8 + 4 < leaf_len) prefetch32(leaf + (it + 1) * 8 + 4);
        }
        permute_dev<KIND>(s, pscratch, SVB_BLOCK);
    }
    ulonglong2* o = reinterpret_cast<ulonglong2*>(leaf_digests + ((size_t)unit * 4 + cls) * 4);
    const u64 bad = ~0ull;
    o[0] = canon_ok ? make_ulonglong2(canon(s[0]), canon(s[1])) : make_ulonglong2(bad, bad);
    o[1] = canon_ok ? make_ulonglong2(canon(s[2]), canon(s[3])) : make_ulonglong2(bad, bad);
}

// The fused query kernel: one thread per (class, unit).  leaf_digests: nullptr = every oracle-tree chain hashes its own
// leaf (the fused form); otherwise the digests fri_leaf_kernel left for this batch.
template <int KIND>
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS_K(KIND)) fri_query_kernel(const u64* __restrict__ records, FriKernelParams P,
                                                        const u64* __restrict__ scratch, u32* __restrict__ accept_bitmap,
                                                        u32* __restrict__ first_fail, const u64* __restrict__ leaf_digests) {
    __shared__ u64 pscratch[PermScratch<KIND>::array_len(SVB_BLOCK)];
    const sv_fri_layout& L = P.L;
    // Grid order.  Phase A: unit groups of `group_blocks` blocks; inside a group the four oracle-tree classes
    // (heaviest first) and the algebra class over the SAME units, so that the leaf evaluations, which both the
    // Merkle chain of their tree and the algebra chain read, come from DRAM once and from L2 the second time
    // (class-major order over the whole batch fetched them twice: tools/lab/NOTES.md).  Phase B: the step
    // trees, class-major over the whole batch, deepest first -- they share no data with anything, and ending
    // the grid with the shortest chains keeps the drain of the last wave short.
    const u32 per_group = P.group_blocks * P.n_classes_a, n_a = P.n_groups * per_group;
    u32 cls, blk;
    if (blockIdx.x < n_a) {
        const u32 grp = blockIdx.x / per_group, in_grp = blockIdx.x - grp * per_group;
        cls = P.class_order[in_grp / P.group_blocks];
        blk = grp * P.group_blocks + in_grp % P.group_blocks;
    } else {
        const u32 b = blockIdx.x - n_a;
        cls = P.class_order[P.n_classes_a + b / P.blocks_per_class];
        blk = b % P.blocks_per_class;
    }
    const u32 unit = blk * blockDim.x + threadIdx.x;
    if (blk >= P.blocks_per_class || unit >= P.n_units) return;
    u32 proof = unit / P.num_queries, query = unit - proof * P.num_queries;
    const u64* rec = records + (size_t)proof * L.record_words;
    const u64* q = rec + L.header_words + (size_t)query * L.query_words;
    u64 x_index_fe = __ldg(rec + L.off_indices + query);
    u32 lde_bits = L.lde_bits;
    u64 x_index = x_index_fe & ((1ull << lde_bits) - 1);       // to_bits(.., 64).take(lde_bits)  (fri_chip.rs:245-250)
    u32 cap_index = (u32)(x_index >> (lde_bits - __popc(L.ncap - 1)));  // top cap_height bits (:72-82), reused by every tree (:308)

    if (cls < 4 + P.num_steps) {
        // cls < 4: verify_initial_merkle_proof, oracle `cls` (fri_chip.rs:85-110)
        // else:    step Merkle proof (fri_chip.rs:303-311): leaf = the 2^arity_bits Fp2 evals of the coset (their own
        //          digest when they are 4 words, hashed otherwise), index = x_index >> (arity bits consumed so far)
        bool init = cls < 4;
        u32 i = init ? 0 : cls - 4;
        const u64* cap = rec + (init ? L.off_init_caps + ((size_t)cls * L.ncap + cap_index) * 4
                                     : L.off_step_caps + ((size_t)i * L.ncap + cap_index) * 4);
        const u64* leaf = q + (init ? L.q_off_init_evals[cls] : L.q_off_step_evals[i]);
        const u64* sibs = q + (init ? L.q_off_init_sibs[cls] : L.q_off_step_sibs[i]);
        u32 leaf_len = init ? L.leaf_len[cls] : (2u << L.step_arity_bits[i]);
        if (leaf_digests && init && leaf_len > 4) {   // the leaf was hashed by fri_leaf_kernel: its digest enters as a 4-word "leaf"
            leaf = leaf_digests + ((size_t)unit * 4 + cls) * 4;
            leaf_len = 4;
        }
        u32 rc = merkle_chain<KIND>(leaf, leaf_len, sibs, init ? L.init_depth : L.step_depth[i],
                              init ? x_index : x_index >> L.step_index_shift[i], cap, init ? SV_FAIL_INIT_MERKLE : SV_FAIL_STEP_MERKLE,
                              pscratch);
        if (rc) report_fail(accept_bitmap, first_fail, proof, query,
                            rc == SV_FAIL_NONCANONICAL ? 0 : (init ? 1 + cls : 8 + 3 * i + 2), rc);
        return;
    }

    // ---- algebra chain -------------------------------------------------------------------------
    // x = 7 * omega^{bitrev(x_index)} (fri_chip.rs:152-166,262-264): bit (lde_bits-1-i) of x_index selects omega^(2^i)
    u64 x = 7;
    for (u32 i = 0; i < lde_bits; i++)
        if ((x_index >> (lde_bits - 1 - i)) & 1) x = mul(x, P.omega_pow2[i]);
    x = canon(x);

    fp2 alpha = mk2(__ldg(rec + L.off_alpha), __ldg(rec + L.off_alpha + 1));
    fp2 zeta = mk2(__ldg(rec + L.off_zeta), __ldg(rec + L.off_zeta + 1));
    fp2 zeta_next = mk2(__ldg(rec + L.off_zeta_next), __ldg(rec + L.off_zeta_next + 1));
    fp2 ro0 = mk2(scratch[4 * (size_t)proof], scratch[4 * (size_t)proof + 1]);
    fp2 ro1 = mk2(scratch[4 * (size_t)proof + 2], scratch[4 * (size_t)proof + 3]);

    // batch_initial_polynomials (fri_chip.rs:112-149).  The evals are range-checked by the init
    // chains of the same unit; here they are only consumed.
    // batch 0: all polys, oracle order; reduce_extension folds from the last term.
    fp2 r0 = mk2(0, 0);
    for (int k = 3; k >= 0; k--) r0 = horner_base(alpha, q + L.q_off_init_evals[k], P.oracle_num_polys[k], r0);
    fp2 r1 = horner_base(alpha, q + L.q_off_init_evals[2], P.num_zs, mk2(0, 0));
    fp2 d0 = sub2(mk2(x, 0), zeta), d1 = sub2(mk2(x, 0), zeta_next);
    u32 fail_key = 0xFFFFFFFFu, fail_code = 0;
    if (is_zero2(d0) || is_zero2(d1)) { fail_key = 5; fail_code = SV_FAIL_ZERO_DENOM; }
    // sum = ((0 * alpha^n0 + (r0 - ro0)/d0) * alpha^n1) + (r1 - ro1)/d1
    fp2 an1 = mk2(1, 0);
    for (u32 i = 0; i < L.n1; i++) an1 = mul2(an1, alpha);   // exp() as a product loop (goldilocks_extension_chip.rs:264-282)
    fp2 sum = mul2(sub2(r0, ro0), inv2(d0));
    sum = add2(mul2(sum, an1), mul2(sub2(r1, ro1), inv2(d1)));
    fp2 prev = sum;

    u64 idx = x_index;
    u64 inv_x = P.num_steps ? canon(inv(x)) : 0;   // x = 7 * omega^e is never 0; carried along by squaring
    for (u32 i = 0; i < P.num_steps; i++) {
        const u32 ab = L.step_arity_bits[i];
        const u64* ev = q + L.q_off_step_evals[i];
        const u32 within = (u32)idx & ((1u << ab) - 1);       // x_index_within_coset (fri_chip.rs:279-282)
        // evals[x_index_within_coset] == prev_eval (fri_chip.rs:285-292)
        ulonglong2 e = __ldg(reinterpret_cast<const ulonglong2*>(ev + 2 * within));
        if ((e.x != prev.c0 || e.y != prev.c1) && 8 + 3 * i < fail_key) { fail_key = 8 + 3 * i; fail_code = SV_FAIL_STEP_EVAL; }
        // next_eval (fri_chip.rs:168-226; any arity: fri_fold.cuh).  Its division is by differences of distinct coset
        // points, never zero, so order key 8 + 3 i + 1 stays unused.
        fp2 beta = mk2(__ldg(rec + L.off_betas + 2 * i), __ldg(rec + L.off_betas + 2 * i + 1));
        prev = fri_fold(ab, ev, x, inv_x, within, beta);
        for (u32 k = 0; k < ab; k++) { x = mulc(x, x); inv_x = mulc(inv_x, inv_x); }   // exp_power_of_2(x, arity_bits) (:313)
        idx >>= ab;
    }
    // final_poly(x) == prev_eval (fri_chip.rs:317-325)
    fp2 fe = mk2(0, 0);
    const u64* fp = rec + L.off_final_poly;
    for (u32 j = P.final_poly_len; j-- > 0;) {
        ulonglong2 c = __ldg(reinterpret_cast<const ulonglong2*>(fp + 2 * j));
        fe = add2(scale2(fe, x), mk2(c.x, c.y));
    }
    if (!eq2(fe, prev) && 8 + 3 * P.num_steps < fail_key) { fail_key = 8 + 3 * P.num_steps; fail_code = SV_FAIL_FINAL; }
    if (fail_code) report_fail(accept_bitmap, first_fail, proof, query, fail_key, fail_code);
}

// n independent permutations, thread per state (canonical in / out).
template <int KIND>
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS_K(KIND)) poseidon_permute_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n) {
    __shared__ u64 scratch[PermScratch<KIND>::array_len(SVB_BLOCK)];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 s[12];
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(in + 12 * i);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        ulonglong2 v = __ldg(p + k);
        s[2 * k] = v.x;
        s[2 * k + 1] = v.y;
    }
    permute_dev<KIND>(s, scratch, SVB_BLOCK);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(out + 12 * i);
#pragma unroll
    for (int k = 0; k < 6; k++) o[k] = make_ulonglong2(canon(s[2 * k]), canon(s[2 * k + 1]));
}

// n independent Merkle paths, thread per path.  Record = up4(leaf_len) + 4*depth words.
template <int KIND>
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS_K(KIND)) merkle_verify_kernel(const u64* __restrict__ paths, const u64* __restrict__ indices,
                                                            const u64* __restrict__ caps, unsigned char* __restrict__ ok,
                                                            size_t n, u32 leaf_len, u32 depth, u32 cap_height) {
    __shared__ u64 scratch[PermScratch<KIND>::array_len(SVB_BLOCK)];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 leaf_words = up4(leaf_len);
    const u64* rec = paths + i * (size_t)(leaf_words + 4 * depth);
    u64 index = __ldg(indices + i);
    u32 cap_index = (u32)((index >> depth) & ((1ull << cap_height) - 1));
    u32 rc = merkle_chain<KIND>(rec, leaf_len, rec + leaf_words, depth, index, caps + 4 * (size_t)cap_index, 1, scratch);
    ok[i] = rc == 0;
}

// out[i] = a[i] * b[i] + c[i] mod p, canonical; inputs are arbitrary u64 (LOOSE).  Exposes the device field
// arithmetic (GoldilocksChip::mul_add, chip/goldilocks_chip.rs:175) so that its rare carry/borrow paths
// can be driven with crafted operands from the tests.
__global__ void goldilocks_mul_add_kernel(const u64* __restrict__ a, const u64* __restrict__ b, const u64* __restrict__ c,
                                          u64* __restrict__ out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = canon(mul_add(a[i], b[i], c[i]));
}

// ---- device-side Fiat-Shamir -------------------------------------------------------
// One thread per proof replays PlonkVerifierChip::get_challenges (chip/plonk/plonk_verifier_chip.rs:
// 55-154) over the duplex sponge of HasherChip (chip/hasher_chip.rs:51-120: rate 8, overwrite mode,
// squeeze pops from the END of state[0..8]) and writes zeta, zeta_next = g*zeta, fri_alpha, fri_betas,
// fri_pow_response and the query indices into the record header -- the same fields, in the same
// order, as the host-side sv_fri_challenges.  ~85 dependent permutations per proof: latency-bound, so
// it runs on its own small blocks beside the query kernel of the neighbouring chunk.
#define SVB_FS_BLOCK 32
struct FsParams {
    u64 circuit_digest[4];
    u64 g;                 // generator of the trace subgroup, 7^((p-1)/2^degree_bits)
    u32 num_challenges;
};
template <int KIND>
struct DevChallenger {
    u64 st[12];
    u64 in[8], out[8];     // dynamically indexed: local memory, negligible beside the permutations
    int n_in, n_out;
    u64* scratch;
    // one out-of-line copy of the permutation for the ~10 call sites of the transcript
    __device__ __noinline__ void permute() {
        permute_dev<KIND>(st, scratch, SVB_FS_BLOCK);
#pragma unroll
        for (int i = 0; i < 12; i++) st[i] = canon(st[i]);
#pragma unroll
        for (int i = 0; i < 8; i++) out[i] = st[i];
        n_out = 8;
    }
    SVB_D void duplex(int len) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i < len) st[i] = in[i];
        permute();
        n_in = 0;
    }
    SVB_D void observe(u64 v) {
        n_out = 0;                 // update() clears the output buffer (hasher_chip.rs:56)
        in[n_in++] = v;
        if (n_in == 8) duplex(8);
    }
    SVB_D void observe_n(const u64* __restrict__ p, u32 n) {
        for (u32 i = 0; i < n; i++) observe(p[i]);
    }
    SVB_D u64 squeeze() {
        if (n_in) duplex(n_in);
        if (n_out == 0) permute();
        return out[--n_out];
    }
};

template <int KIND>
__global__ void __launch_bounds__(SVB_FS_BLOCK) fri_challenges_kernel(u64* __restrict__ records, FriKernelParams P, FsParams F,
                                                                      const u64* __restrict__ pi_hashes) {
    __shared__ u64 pscratch[PermScratch<KIND>::array_len(SVB_FS_BLOCK)];
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_proofs) return;
    const sv_fri_layout& L = P.L;
    u64* rec = records + (size_t)p * L.record_words;
    DevChallenger<KIND> ch;
#pragma unroll
    for (int i = 0; i < 12; i++) ch.st[i] = 0;
    ch.n_in = ch.n_out = 0;
    ch.scratch = pscratch;
    const u32 cap_words = L.ncap * 4;
    for (int i = 0; i < 4; i++) ch.observe(F.circuit_digest[i]);                      // :65-68
    ch.observe_n(pi_hashes + 4 * (size_t)p, 4);                                       // :69-71
    ch.observe_n(rec + L.off_init_caps + 1 * cap_words, cap_words);                   // wires_cap
    for (u32 i = 0; i < 2 * F.num_challenges; i++) (void)ch.squeeze();                // plonk betas, gammas
    ch.observe_n(rec + L.off_init_caps + 2 * cap_words, cap_words);                   // zs_partial_products_cap
    for (u32 i = 0; i < F.num_challenges; i++) (void)ch.squeeze();                    // plonk alphas
    ch.observe_n(rec + L.off_init_caps + 3 * cap_words, cap_words);                   // quotient_polys_cap
    u64 z0 = ch.squeeze(), z1 = ch.squeeze();                                         // plonk_zeta
    rec[L.off_zeta] = z0; rec[L.off_zeta + 1] = z1;
    rec[L.off_zeta_next] = mulc(z0, F.g); rec[L.off_zeta_next + 1] = mulc(z1, F.g);
    ch.observe_n(rec + L.off_open0, 2 * L.n0);                                        // openings, batch order
    ch.observe_n(rec + L.off_open1, 2 * L.n1);
    u64 a0 = ch.squeeze(), a1 = ch.squeeze();                                         // fri_alpha
    rec[L.off_alpha] = a0; rec[L.off_alpha + 1] = a1;
    for (u32 st = 0; st < P.num_steps; st++) {
        ch.observe_n(rec + L.off_step_caps + (size_t)st * cap_words, cap_words);
        u64 b0 = ch.squeeze(), b1 = ch.squeeze();
        rec[L.off_betas + 2 * st] = b0; rec[L.off_betas + 2 * st + 1] = b1;
    }
    ch.observe_n(rec + L.off_final_poly, 2 * P.final_poly_len);
    ch.observe(rec[L.off_pow_witness]);
    rec[L.off_pow_response] = ch.squeeze();
    for (u32 q = 0; q < P.num_queries; q++) rec[L.off_indices + q] = ch.squeeze();
}

// The same transcript with the lane-cooperative permutation in its latency form (poseidon_g_coop2.cuh): one 16-lane
// group per proof, lane l holds sponge word l.  Absorbing is a coalesced load of up to 8 consecutive words by lanes
// 0..7; a squeeze broadcasts the word of lane (n_out - 1).  Poseidon-Goldilocks only.  (The first cooperative mapping
// lives on in tools/lab/poseidon_g_coop_v1.cuh for tools/lab/coopbench.cu: 10.5 us per permutation against 6.1 here.)
#define SVB_COOP_BLOCK 128
#define SVB_COOP_GROUPS (SVB_COOP_BLOCK / SVB_COOP2_GROUP)
// ONE out-of-line copy of the permutation (~2 000 instructions) for the ~20 call sites of the transcript
__device__ __noinline__ u64 coop2_permute_canonical(u64 s, int l, Coop2Tables<SVB_COOP_GROUPS>* T, int g) {
    return canon(poseidon_g_coop2(s, l, *T, g));
}
struct CoopChallenger {
    u64 s;          // this lane's state word
    int n_out, g, l;
    Coop2Tables<SVB_COOP_GROUPS>* T;
    SVB_D void permute() {
        s = coop2_permute_canonical(s, l, T, g);
        n_out = 8;
    }
    // observe the concatenation of two segments (n1 may be 0), in rate-8 chunks, overwrite mode
    SVB_D void absorb(const u64* __restrict__ p0, u32 n0, const u64* __restrict__ p1, u32 n1) {
        const u32 total = n0 + n1;
        for (u32 off = 0; off < total; off += 8) {
            u32 idx = off + (u32)l;
            if (l < 8 && idx < total) s = idx < n0 ? p0[idx] : p1[idx - n0];
            permute();
        }
    }
    SVB_D u64 squeeze() {
        if (n_out == 0) permute();
        n_out--;
        return coop2_shfl(s, n_out);
    }
};

__global__ void __launch_bounds__(SVB_COOP_BLOCK) fri_challenges_coop_kernel(u64* __restrict__ records, FriKernelParams P, FsParams F,
                                                                             const u64* __restrict__ pi_hashes) {
    __shared__ Coop2Tables<SVB_COOP_GROUPS> T;
    coop2_load_tables(T);
    const int l = threadIdx.x & (SVB_COOP2_GROUP - 1);
    u32 p = (blockIdx.x * blockDim.x + threadIdx.x) / SVB_COOP2_GROUP;
    const bool valid = p < P.n_proofs;
    if (!valid) p = P.n_proofs - 1;            // keep the warp convergent for the shuffles; writes are suppressed
    const bool writer = valid && l == 0;
    const sv_fri_layout& L = P.L;
    u64* rec = records + (size_t)p * L.record_words;
    const u32 cap_words = L.ncap * 4;
    CoopChallenger ch;
    ch.l = l; ch.n_out = 0; ch.T = &T; ch.g = threadIdx.x / SVB_COOP2_GROUP;
    // first chunk: circuit digest (4) + public-input hash (4)   (plonk_verifier_chip.rs:65-71)
    ch.s = l < 4 ? F.circuit_digest[l & 3] : (l < 8 ? pi_hashes[4 * (size_t)p + (l - 4)] : 0);
    ch.permute();
    ch.absorb(rec + L.off_init_caps + 1 * cap_words, cap_words, nullptr, 0);           // wires_cap
    for (u32 i = 0; i < 2 * F.num_challenges; i++) (void)ch.squeeze();                  // plonk betas, gammas
    ch.absorb(rec + L.off_init_caps + 2 * cap_words, cap_words, nullptr, 0);           // zs_partial_products_cap
    for (u32 i = 0; i < F.num_challenges; i++) (void)ch.squeeze();                      // plonk alphas
    ch.absorb(rec + L.off_init_caps + 3 * cap_words, cap_words, nullptr, 0);           // quotient_polys_cap
    u64 z0 = ch.squeeze(), z1 = ch.squeeze();                                           // plonk_zeta
    if (writer) {
        rec[L.off_zeta] = z0; rec[L.off_zeta + 1] = z1;
        rec[L.off_zeta_next] = mulc(z0, F.g); rec[L.off_zeta_next + 1] = mulc(z1, F.g);
    }
    ch.absorb(rec + L.off_open0, 2 * L.n0, rec + L.off_open1, 2 * L.n1);                // openings, batch order
    u64 a0 = ch.squeeze(), a1 = ch.squeeze();                                           // fri_alpha
    if (writer) { rec[L.off_alpha] = a0; rec[L.off_alpha + 1] = a1; }
    for (u32 st = 0; st < P.num_steps; st++) {
        ch.absorb(rec + L.off_step_caps + (size_t)st * cap_words, cap_words, nullptr, 0);
        u64 b0 = ch.squeeze(), b1 = ch.squeeze();
        if (writer) { rec[L.off_betas + 2 * st] = b0; rec[L.off_betas + 2 * st + 1] = b1; }
    }
    ch.absorb(rec + L.off_final_poly, 2 * P.final_poly_len, rec + L.off_pow_witness, 1);
    u64 pw = ch.squeeze();
    if (writer) rec[L.off_pow_response] = pw;
    for (u32 q = 0; q < P.num_queries; q++) {
        u64 v = ch.squeeze();
        if (writer) rec[L.off_indices + q] = v;
    }
}

// n independent permutations on the lane-cooperative mapping of the transcript (16 lanes per state): exposes
// coop2_permute_canonical to the tests so that its carry / borrow paths can be driven with crafted states.
__global__ void __launch_bounds__(SVB_COOP_BLOCK) poseidon_permute_coop_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n) {
    __shared__ Coop2Tables<SVB_COOP_GROUPS> T;
    coop2_load_tables(T);
    const int l = threadIdx.x & (SVB_COOP2_GROUP - 1);
    size_t g = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) / SVB_COOP2_GROUP;
    const bool valid = g < n;
    if (!valid) g = n - 1;                       // keep the warp convergent; writes are suppressed
    u64 s = l < 12 ? in[12 * g + l] : 0;
    s = coop2_permute_canonical(s, l, &T, threadIdx.x / SVB_COOP2_GROUP);
    if (valid && l < 12) out[12 * g + l] = s;
}

// ---- Merkle tree construction (the prover side of the same hash) ---------------------
// Leaf digests: hash_or_noop of each leaf row (plonky2 MerkleTree::new; call sites
// plonky2_semaphore/access_set.rs:25, circuit.rs:91).  One thread per leaf.
template <int KIND>
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS_K(KIND)) merkle_leaf_hash_kernel(const u64* __restrict__ leaves, u32 leaf_len,
                                                                                    size_t n, u64* __restrict__ digests) {
    __shared__ u64 scratch[PermScratch<KIND>::array_len(SVB_BLOCK)];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u64* leaf = leaves + i * (size_t)leaf_len;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    if (leaf_len <= 4) {
        for (u32 k = 0; k < leaf_len; k++) s[k] = leaf[k];
    } else {
#pragma unroll 1
        for (u32 off = 0; off < leaf_len; off += 8) {
            u32 rem = leaf_len - off;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if ((u32)k < rem) s[k] = leaf[off + k];
            permute_dev<KIND>(s, scratch, SVB_BLOCK);
        }
    }
    ulonglong2* o = reinterpret_cast<ulonglong2*>(digests + 4 * i);
    o[0] = make_ulonglong2(canon(s[0]), canon(s[1]));
    o[1] = make_ulonglong2(canon(s[2]), canon(s[3]));
}
// The same digests from POLYNOMIAL-MAJOR values (the layout an LDE produces: word j of leaf i at cols[j * n + i]): thread i
// walks down its column, so every load of a warp is one coalesced 256-byte run -- no transposing copy before the hash.
template <int KIND>
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS_K(KIND)) merkle_leaf_hash_cols_kernel(const u64* __restrict__ cols, u32 leaf_len,
                                                                                         size_t n, u64* __restrict__ digests) {
    __shared__ u64 scratch[PermScratch<KIND>::array_len(SVB_BLOCK)];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    if (leaf_len <= 4) {
        for (u32 k = 0; k < leaf_len; k++) s[k] = __ldg(cols + (size_t)k * n + i);
    } else {
#pragma unroll 1
        for (u32 off = 0; off < leaf_len; off += 8) {
            u32 rem = leaf_len - off;
#pragma unroll
            for (int k = 0; k < 8; k++)
                if ((u32)k < rem) s[k] = __ldg(cols + (size_t)(off + k) * n + i);
            permute_dev<KIND>(s, scratch, SVB_BLOCK);
        }
    }
    ulonglong2* o = reinterpret_cast<ulonglong2*>(digests + 4 * i);
    o[0] = make_ulonglong2(canon(s[0]), canon(s[1]));
    o[1] = make_ulonglong2(canon(s[2]), canon(s[3]));
}
// One tree level: parent j = two_to_one(child 2j, child 2j+1).  One thread per parent.
template <int KIND>
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS_K(KIND)) merkle_level_kernel(const u64* __restrict__ children, u64* __restrict__ parents,
                                                                                size_t n_parents) {
    __shared__ u64 scratch[PermScratch<KIND>::array_len(SVB_BLOCK)];
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (j >= n_parents) return;
    u64 s[12];
    ldg4(children + 8 * j, s);
    ldg4(children + 8 * j + 4, s + 4);
#pragma unroll
    for (int k = 8; k < 12; k++) s[k] = 0;
    permute_dev<KIND>(s, scratch, SVB_BLOCK);
    ulonglong2* o = reinterpret_cast<ulonglong2*>(parents + 4 * j);
    o[0] = make_ulonglong2(canon(s[0]), canon(s[1]));
    o[1] = make_ulonglong2(canon(s[2]), canon(s[3]));
}

// first_fail post-pass: 0xFFFFFFFF (never failed) -> 0, else (query << 8) | code.
__global__ void fri_finalize_kernel(u32* __restrict__ first_fail, u32 n) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    u32 v = first_fail[p];
    first_fail[p] = v == 0xFFFFFFFFu ? 0u : (((v >> 20) << 8) | (v & 0xFFu));
}

}  // namespace svb
