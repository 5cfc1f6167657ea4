// LAB COPY (not part of the library since kernels r2.4: the transcript runs on csrc/poseidon_g_coop2.cuh, the latency form; this first
// mapping is kept for the before / after of tools/lab/coopbench.cu).
// Lane-cooperative Poseidon-Goldilocks: one permutation spread over a 16-lane group (lanes 0..11 hold
// the 12 state words, lanes 12..15 idle), two groups per warp.  Same function as poseidon_g_dev
// (chip/plonk/gates/poseidon.rs:634-686, fast form); a different mapping.
//
// Why it exists.  The thread-per-permutation kernel maximises THROUGHPUT (all 32 lanes busy, no
// shuffles), but one permutation takes ~44 us on a lone warp.  The Fiat-Shamir transcript is ~85
// DEPENDENT permutations per proof, so at a few thousand proofs per call it is latency-bound; here the
// S-boxes of a full round, the 11 multiply-adds of a partial round and the MDS rows run in parallel
// across lanes (warp shuffles gather the circulant MDS inputs and tree-reduce the partial-round dot
// product), which cuts the latency of one permutation ~5x.  The price is throughput: the 22 partial
// rounds keep 11 of 12 lanes idle during the lane-0 S-box -- measured in tools/lab (coop vs thread).
#pragma once
#include "../../stark-verifier_b200/csrc/poseidon_g.cuh"

namespace svb {

#if defined(__CUDACC__)
#define SVB_COOP_GROUP 16

// constants staged in shared memory (per-lane addresses differ, which constant memory serialises)
struct CoopTables {
    u64 rc_full[96];     // ALL_ROUND_CONSTANTS of rounds 0..3 and 26..29
    u64 first[12];       // FAST_PARTIAL_FIRST_ROUND_CONSTANT
    u64 prc[22];         // FAST_PARTIAL_ROUND_CONSTANTS
    u64 init[121];       // FAST_PARTIAL_ROUND_INITIAL_MATRIX
    u64 w[242], v[242];  // FAST_PARTIAL_ROUND_W_HATS / _VS
};
SVB_D void coop_load_tables(CoopTables& T) {
    for (int i = threadIdx.x; i < 96; i += blockDim.x) T.rc_full[i] = d_ALL_ROUND_CONSTANTS[i < 48 ? i : 12 * 26 + (i - 48)];
    for (int i = threadIdx.x; i < 12; i += blockDim.x) T.first[i] = d_FAST_PARTIAL_FIRST_ROUND_CONSTANT[i];
    for (int i = threadIdx.x; i < 22; i += blockDim.x) T.prc[i] = d_FAST_PARTIAL_ROUND_CONSTANTS[i];
    for (int i = threadIdx.x; i < 121; i += blockDim.x) T.init[i] = d_FAST_PARTIAL_ROUND_INITIAL_MATRIX[i];
    for (int i = threadIdx.x; i < 242; i += blockDim.x) { T.w[i] = d_FAST_PARTIAL_ROUND_W_HATS[i]; T.v[i] = d_FAST_PARTIAL_ROUND_VS[i]; }
    __syncthreads();
}

SVB_D constexpr u32 coop_circ(int k) {
    constexpr u32 circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};   // MDS_MATRIX_CIRC (poseidon.rs:321)
    return circ[k];
}
SVB_D u64 coop_shfl(u64 v, int src) {
    u32 lo = __shfl_sync(0xFFFFFFFFu, (u32)v, src, SVB_COOP_GROUP);
    u32 hi = __shfl_sync(0xFFFFFFFFu, (u32)(v >> 32), src, SVB_COOP_GROUP);
    return ((u64)hi << 32) | lo;
}

// full round `slot` (0..7 = rounds 0..3, 26..29): constant layer, S-box on every lane, circulant MDS
SVB_D u64 coop_full_round(u64 s, int l, int slot, const CoopTables& T) {
    const bool act = l < 12;
    s = sbox7(add_lc(s, act ? T.rc_full[12 * slot + l] : 0));
    u64 al = 0, ah = 0;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        int src = l + k;
        src = src >= 12 ? src - 12 : src;
        u64 v = coop_shfl(s, act ? src : l);
        al += (u64)coop_circ(k) * (u32)v;                 // row l: sum_k CIRC[k] * s[(l + k) % 12]  (:450-479)
        ah += (u64)coop_circ(k) * (u32)(v >> 32);
    }
    if (l == 0) { al += 8ull * (u32)s; ah += 8ull * (u32)(s >> 32); }            // MDS_MATRIX_DIAG[0]
    u64 t = al + (ah << 32);
    u32 top = (u32)(ah >> 32) + (t < al ? 1u : 0u);
    return act ? reduce96(t, top) : 0;
}

// One permutation; `s` is this lane's state word (LOOSE in, LOOSE out; lanes 12..15 carry 0).
SVB_D u64 poseidon_g_coop(u64 s, int l, const CoopTables& T) {
    const bool act = l < 12;
#pragma unroll 1
    for (int slot = 0; slot < 4; slot++) s = coop_full_round(s, l, slot, T);
    // partial section: first-round constants, then the dense initial matrix on lanes 1..11 (:504-537)
    s = add_lc(s, act ? T.first[l] : 0);
    {
        dot_acc a;
        dot_init(a);
#pragma unroll 1
        for (int r = 1; r < 12; r++) {
            u64 v = coop_shfl(s, r);
            u64 m = (l >= 1 && act) ? T.init[(r - 1) * 11 + (l - 1)] : 0;
            dot_mac(a, v, m);
        }
        u64 t = dot_reduce(a);
        s = l == 0 ? s : (act ? t : 0);
    }
    // 22 partial rounds, software-pipelined.  The dot product D_r = sum_i w_hat_i(r) * s_i(r) does not depend on
    // this round's S-box output, only on the state left by the previous round, so it is formed (products
    // + shuffle tree) one round ahead and stays off the critical path  S-box -> 25*t + D -> reduce.  The
    // S-box runs on every lane (no divergent branch, so ptxas can interleave it with the tree of the
    // previous iteration); only lane 0's result is used.
    u32 d0, d1, d2, d3, d4;
    auto dot_ahead = [&](int r, u64 sv) {
        u32 p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0;
        u64 w = (l >= 1 && act) ? T.w[r * 11 + l - 1] : 0;
        mulw4(sv, w, p0, p1, p2, p3);                               // w_hat_i * s_i, unreduced (0 on lanes 0, 12..15)
#pragma unroll
        for (int off = 8; off >= 1; off >>= 1) {
            u32 q0 = __shfl_xor_sync(0xFFFFFFFFu, p0, off, SVB_COOP_GROUP);
            u32 q1 = __shfl_xor_sync(0xFFFFFFFFu, p1, off, SVB_COOP_GROUP);
            u32 q2 = __shfl_xor_sync(0xFFFFFFFFu, p2, off, SVB_COOP_GROUP);
            u32 q3 = __shfl_xor_sync(0xFFFFFFFFu, p3, off, SVB_COOP_GROUP);
            u32 q4 = __shfl_xor_sync(0xFFFFFFFFu, p4, off, SVB_COOP_GROUP);
            asm("add.cc.u32 %0, %0, %5;\n\t addc.cc.u32 %1, %1, %6;\n\t addc.cc.u32 %2, %2, %7;\n\t addc.cc.u32 %3, %3, %8;\n\t"
                "addc.u32 %4, %4, %9;"
                : "+r"(p0), "+r"(p1), "+r"(p2), "+r"(p3), "+r"(p4) : "r"(q0), "r"(q1), "r"(q2), "r"(q3), "r"(q4));
        }
        d0 = p0; d1 = p1; d2 = p2; d3 = p3; d4 = p4;
    };
    dot_ahead(0, s);
#pragma unroll 1
    for (int r = 0; r < 22; r++) {
        const u64 t = sbox7_add(s, l == 0 ? T.prc[r] : 0);        // lane 0: the S-box (+ round constant)
        const u64 s0 = coop_shfl(t, 0);
        // lane 0: 25 * s0 + D_r  (MDS_CIRC[0] + MDS_DIAG[0] = 25)
        u64 lo = 25ull * (u32)s0, hi = 25ull * (u32)(s0 >> 32);
        u64 mid = (lo >> 32) + (u32)hi;
        u32 e0 = (u32)lo, e1 = (u32)mid, e2 = (u32)(hi >> 32) + (u32)(mid >> 32), e3 = 0, e4 = 0;
        asm("add.cc.u32 %0, %0, %5;\n\t addc.cc.u32 %1, %1, %6;\n\t addc.cc.u32 %2, %2, %7;\n\t addc.cc.u32 %3, %3, %8;\n\t"
            "addc.u32 %4, %4, %9;"
            : "+r"(e0), "+r"(e1), "+r"(e2), "+r"(e3), "+r"(e4) : "r"(d0), "r"(d1), "r"(d2), "r"(d3), "r"(d4));
        const u64 n0 = red5(e0, e1, e2, e3, e4);
        const u64 ni = (l >= 1 && act) ? mul_add(T.v[r * 11 + l - 1], s0, s) : 0;   // s_i + v_i * s0
        s = l == 0 ? n0 : ni;
        if (r < 21) dot_ahead(r + 1, s);
    }
#pragma unroll 1
    for (int slot = 4; slot < 8; slot++) s = coop_full_round(s, l, slot, T);
    return s;
}
#endif

}  // namespace svb
