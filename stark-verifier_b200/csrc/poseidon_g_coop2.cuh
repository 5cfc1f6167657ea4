// Lane-cooperative Poseidon-Goldilocks, latency form: one permutation per 16-lane group (lane l < 12 holds state
// word l), two groups per warp.  Same function as poseidon_g_dev / poseidon_g_coop (chip/plonk/gates/poseidon.rs:
// 634-686); a different SCHEDULE, built for the Fiat-Shamir transcript, which is ~155 DEPENDENT permutations per
// proof and therefore bound by the latency of one permutation, not by throughput.
//
// What is on the critical path of poseidon_g_coop (measured 10.5 us per permutation, one warp per sub-partition) and is not here:
//   * partial rounds.  Everything in the partial section except the one x^7 per round is linear, so with
//     z = the S-box outputs of full round 3 and q_r = 25 x_r^7 every quantity is a linear form over (1, z, q):
//         x_{r+1} = q_r + B_{r+1}(1, z, q_0 .. q_{r-1})        (tools/gen_poseidon_coop2_constants.py)
//     Every lane computes the S-box chain REDUNDANTLY (no broadcast of its result), as
//         x_{r+1} = x_r^3 * x_r^4 + B_{r+1}   (the factor 25 scaled away)   three dependent multiplications per round,
//     while the rows B_r and the 12 rows F of the state that leaves the section are accumulated one per lane and
//     slot (3 slots x 16 lanes = 22 B rows + 12 F rows) from q_{r-1} = x_r - B_r -- one round behind the chain,
//     off its critical path; B_{r+1} reaches the lanes by one shuffle.  The fast form's dot product + shuffle
//     tree + v-vector update (a fourth multiplication, two reductions and five shuffle levels per round) is gone.
//   * full rounds.  The S-box outputs are exchanged through 96 bytes of shared memory per group (one store, six
//     128-bit broadcast loads) instead of 24 shuffles; the constants of the next round ride in the MDS
//     accumulators; the MDS layer of round 3 and mds_partial_layer_init are folded into the tables.
#pragma once
#include "poseidon_g.cuh"

namespace svb {

#if defined(__CUDACC__)
#define SVB_COOP2_GROUP 16

// plain device memory: the constant bank of the library is full, and the tables are read once per block into shared memory
#define SVB_TABLE(name, n) __device__ const uint64_t d_##name[n]
#include "poseidon_g_coop2_constants.inc"
#undef SVB_TABLE

// tables staged in shared memory (per-lane addresses differ, which constant memory would serialise), plus the
// exchange buffers of the groups of a block
template <int GROUPS>
struct Coop2Tables {
    u64 zc[12 * 3 * 16];          // COOP2_ZC[k][slot][lane]
    u64 qc[23 * 3 * 16];          // COOP2_QC[r][slot][lane]
    u64 c0[3 * 16];               // COOP2_C0[slot][lane]
    u64 rc_in[16];                // ALL_ROUND_CONSTANTS of round 0 (lanes 12..15: 0)
    u64 rc_next[8 * 16];          // FULL_RC_NEXT: constants added by the MDS layer that follows full round f
    alignas(16) u32 mc[16][12];   // row l of the MDS matrix: M[l][k] = CIRC[(k - l) mod 12] + DIAG (lanes 12..15: 0)
    alignas(16) u64 xbuf[2][GROUPS][12];
};
template <int GROUPS>
SVB_D void coop2_load_tables(Coop2Tables<GROUPS>& T) {
    for (int i = threadIdx.x; i < 12 * 3 * 16; i += blockDim.x) T.zc[i] = d_COOP2_ZC[i];
    for (int i = threadIdx.x; i < 23 * 3 * 16; i += blockDim.x) T.qc[i] = d_COOP2_QC[i];
    for (int i = threadIdx.x; i < 3 * 16; i += blockDim.x) T.c0[i] = d_COOP2_C0[i];
    for (int i = threadIdx.x; i < 16; i += blockDim.x) T.rc_in[i] = i < 12 ? d_ALL_ROUND_CONSTANTS[i] : 0;
    for (int i = threadIdx.x; i < 8 * 16; i += blockDim.x) T.rc_next[i] = (i & 15) < 12 ? d_FULL_RC_NEXT[(i >> 4) * 12 + (i & 15)] : 0;
    for (int i = threadIdx.x; i < 16 * 12; i += blockDim.x) {
        const int l = i / 12, k = i % 12;
        u32 v = (u32)d_MDS_MATRIX_CIRC[(k - l + 12) % 12];                       // poseidon.rs:321-322
        if (k == 0 && l == 0) v += (u32)d_MDS_MATRIX_DIAG[0];
        T.mc[l][k] = l < 12 ? v : 0;
    }
    __syncthreads();
}

// what a lane keeps in registers for one permutation: its row of the MDS matrix
struct Coop2Lane {
    int l;              // lane in the group
    u32 mc[12];
};
template <int GROUPS>
SVB_D void coop2_lane_init(Coop2Lane& L, int l, const Coop2Tables<GROUPS>& T) {
    L.l = l;
    const uint4* m = reinterpret_cast<const uint4*>(T.mc[l]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint4 v = m[k];
        L.mc[4 * k] = v.x; L.mc[4 * k + 1] = v.y; L.mc[4 * k + 2] = v.z; L.mc[4 * k + 3] = v.w;
    }
}

SVB_D u64 coop2_shfl(u64 v, int src) {
    u32 lo = __shfl_sync(0xFFFFFFFFu, (u32)v, src, SVB_COOP2_GROUP);
    u32 hi = __shfl_sync(0xFFFFFFFFu, (u32)(v >> 32), src, SVB_COOP2_GROUP);
    return ((u64)hi << 32) | lo;
}

// S-box layer of a full round + exchange + this lane's MDS row + the constants of the next round.
// `s` already carries this round's constant.
SVB_D u64 coop2_full_round(u64 s, const Coop2Lane& L, u64 rc_next, u64* __restrict__ buf) {
    const u64 z = sbox7(s);
    if (L.l < 12) buf[L.l] = z;
    __syncwarp();
    u64 al[3] = {(u32)rc_next, 0, 0}, ah[3] = {rc_next >> 32, 0, 0};
    const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(buf);
#pragma unroll
    for (int k = 0; k < 6; k++) {
        const ulonglong2 v = b2[k];
        al[(2 * k) % 3] += (u64)L.mc[2 * k] * (u32)v.x;
        ah[(2 * k) % 3] += (u64)L.mc[2 * k] * (u32)(v.x >> 32);
        al[(2 * k + 1) % 3] += (u64)L.mc[2 * k + 1] * (u32)v.y;
        ah[(2 * k + 1) % 3] += (u64)L.mc[2 * k + 1] * (u32)(v.y >> 32);
    }
    const u64 lo = al[0] + al[1] + al[2], hi = ah[0] + ah[1] + ah[2];      // < 2^43 each
    const u64 t = lo + (hi << 32);
    const u32 top = (u32)(hi >> 32) + (t < lo ? 1u : 0u);
    return reduce96(t, top);
}

SVB_D void coop2_acc_init(dot_acc& a, u64 c) {
    dot_init(a);
    a.e0 = (u32)c;
    a.e1 = (u32)(c >> 32);
}
// a - b for two LOOSE operands, LOOSE result.  A borrow means the register holds a - b + 2^64 = a - b + EPS (mod p): take EPS
// (the borrow mask) away; if THAT borrows (the value was below EPS) the same correction once more lands >= 2^64 - 2 EPS.
SVB_D u64 coop2_sub_loose(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), r0, r1;
    asm("{\n\t.reg .u32 t0, t1, m, n;\n\t"
        "sub.cc.u32 t0, %2, %4;\n\t subc.cc.u32 t1, %3, %5;\n\t subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 t0, t0, m;\n\t subc.cc.u32 t1, t1, 0;\n\t subc.u32 n, 0, 0;\n\t"
        "sub.cc.u32 %0, t0, n;\n\t subc.u32 %1, t1, 0;\n\t}"
        : "=r"(r0), "=r"(r1) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return ((u64)r1 << 32) | r0;
}

// One partial round in the scaled variable t (x_r = a_r t_r, generator script): t_{r+1} = t_r^3 * t_r^4 + B'_{r+1}.
// In: x = t_r (identical on every lane), pprev = p_{r-1} = t_{r-1}^7; out: t_{r+1}, p_r.
// LAZY: rounds 0..11 also add the z_r terms of slots 1 and 2 (slot 0 is complete before round 0).  SLOT: where B'_{r+1} lives.
template <bool LAZY, int SLOT, int GROUPS>
SVB_D void coop2_partial_round(int r, u64& x, u64& pprev, dot_acc& a0, dot_acc& a1, dot_acc& a2, const Coop2Lane& L,
                               const Coop2Tables<GROUPS>& T, const u64* __restrict__ zbuf) {
    // beside the chain: p_{r-1} into the accumulators, then B'_{r+1} to every lane
    const u64* qc = T.qc + (r * 3) * 16 + L.l;
    if (SLOT == 0) dot_mac(a0, qc[0], pprev);        // rows B'_0..B'_15 are all consumed once SLOT is 1
    else dot_mac(a1, qc[16], pprev);
    const u64 bsrc = dot_reduce(SLOT ? a1 : a0);
    const u64 bn = coop2_shfl(bsrc, (r + 1) & 15);
    // the chain: three dependent multiplications
    const u64 x2 = mul(x, x);
    const u64 x4 = mul(x2, x2);
    const u64 x3 = mul(x, x2);
    if (SLOT == 0) dot_mac(a1, qc[16], pprev);
    dot_mac(a2, qc[32], pprev);
    if (LAZY) {
        const u64 zr = zbuf[r];
        const u64* zc = T.zc + (r * 3) * 16 + L.l;
        dot_mac(a1, zc[16], zr);
        dot_mac(a2, zc[32], zr);
    }
    const u64 xn = mul_add(x3, x4, bn);
    pprev = coop2_sub_loose(xn, bn);
    x = xn;
}

// One permutation; `s` is this lane's state word (LOOSE in, LOOSE out; lanes 12..15 carry 0).  `g` = group index in the block.
template <int GROUPS>
SVB_D u64 poseidon_g_coop2(u64 s, int l, Coop2Tables<GROUPS>& T, int g) {
    Coop2Lane L;
    coop2_lane_init(L, l, T);
    u64* b0 = T.xbuf[0][g];
    u64* b1 = T.xbuf[1][g];
    s = add_lc(s, T.rc_in[l]);
    s = coop2_full_round(s, L, T.rc_next[0 * 16 + l], b0);
    s = coop2_full_round(s, L, T.rc_next[1 * 16 + l], b1);
    s = coop2_full_round(s, L, T.rc_next[2 * 16 + l], b0);
    // full round 3: S-box only; its MDS layer, the first-round constants and mds_partial_layer_init live in the tables
    {
        const u64 z = sbox7(s);
        if (l < 12) b1[l] = z;
        __syncwarp();
    }
    dot_acc a0, a1, a2;
    coop2_acc_init(a0, T.c0[l]);
    coop2_acc_init(a1, T.c0[16 + l]);
    coop2_acc_init(a2, T.c0[32 + l]);
#pragma unroll
    for (int k = 0; k < 12; k++) dot_mac(a0, T.zc[(k * 3) * 16 + l], b1[k]);
    u64 x = coop2_shfl(dot_reduce(a0), 0);       // x_0
    u64 q = 0;
#pragma unroll 2
    for (int r = 0; r < 12; r++) coop2_partial_round<true, 0>(r, x, q, a0, a1, a2, L, T, b1);
#pragma unroll
    for (int r = 12; r < 15; r++) coop2_partial_round<false, 0>(r, x, q, a0, a1, a2, L, T, b1);
#pragma unroll 2
    for (int r = 15; r < 22; r++) coop2_partial_round<false, 1>(r, x, q, a0, a1, a2, L, T, b1);
    dot_mac(a2, T.qc[(22 * 3 + 2) * 16 + l], q);
    s = dot_reduce(a2);                          // state entering round 26, its constants included
    s = coop2_full_round(s, L, T.rc_next[4 * 16 + l], b0);
    s = coop2_full_round(s, L, T.rc_next[5 * 16 + l], b1);
    s = coop2_full_round(s, L, T.rc_next[6 * 16 + l], b0);
    s = coop2_full_round(s, L, T.rc_next[7 * 16 + l], b1);
    return s;
}
#endif

}  // namespace svb
