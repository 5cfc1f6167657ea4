// Hash family "B": Poseidon over the BN254 scalar field Fr (T = 5, S-box x^5, 4 + 60 + 4 rounds) wrapped
// around the 12 Goldilocks limbs of a plonky2 sponge state.  This is the `Hasher` of the reference's
// outermost proof (Bn254PoseidonGoldilocksConfig, bn245_poseidon/plonky2_config.rs:68-75).
//
// Replaces (semantics):
//   Bn254PoseidonPermutation::permute ........ bn245_poseidon/plonky2_config.rs:38-51
//   permute_bn254_poseidon_native ............ bn245_poseidon/native.rs:16-60
//   encode_fe / decode_fe ..................... bn245_poseidon/native.rs:62-77 (+ native_chip/utils.rs:25-36)
//   in-circuit twin AllChip::permute .......... chip/native_chip/all_chip.rs:52-89
//   parameters ................................ bn245_poseidon/constants.rs:5-384,402-404 (generated table
//                                               poseidon_b_constants.inc, Montgomery form)
//
// Fr elements are 8 x 32-bit limbs in Montgomery form (R = 2^256).  Products are computed by product
// scanning with the Montgomery reduction interleaved (FIPS), so that an MDS row -- a 5-term dot
// product -- costs 5*64 + 64 limb products and ONE reduction instead of 5 full Montgomery products:
// 5 r^2 < r 2^256, so the unreduced sum still reduces to < 2r.  Every limb product is one
// IMAD.WIDE.U32 accumulating into a 96-bit column (mad.lo.cc / madc.hi.cc / addc, fused by ptxas).
// On the device the 5-word state is staged in shared memory ([limb][thread], conflict-free) between
// layers, which keeps the loops over lanes / rows rolled and the code small.
#pragma once
#include "goldilocks.cuh"

namespace svb {

#define SVB_TABLE(name, n) static const uint64_t h_##name[n]
#include "poseidon_b_constants.inc"
#undef SVB_TABLE
#if defined(__CUDACC__)
#define SVB_TABLE(name, n) __constant__ uint64_t d_##name[n]
#include "poseidon_b_constants.inc"
#undef SVB_TABLE
#endif
#if defined(__CUDA_ARCH__)
#define SVB_TB(name) d_##name
#else
#define SVB_TB(name) h_##name
#endif

struct fr {
    u32 l[8];
};

// r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
SVB_HD constexpr u32 fr_mod(int i) {
    constexpr u32 m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u, 0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return m[i];
}
SVB_HD constexpr u32 fr_r2(int i) {   // 2^512 mod r
    constexpr u32 m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u, 0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return m[i];
}
static constexpr u32 FR_NINV32 = 0xefffffffu;   // -r^-1 mod 2^32

// 96-bit column accumulator += a * b
SVB_HD void fr_mac(u32& c0, u32& c1, u32& c2, u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t madc.hi.cc.u32 %1, %3, %4, %1;\n\t addc.u32 %2, %2, 0;"
        : "+r"(c0), "+r"(c1), "+r"(c2) : "r"(a), "r"(b));
#else
    u64 p = (u64)a * b;
    u64 s = (u64)c0 + (u32)p;
    c0 = (u32)s;
    s = (u64)c1 + (u32)(p >> 32) + (s >> 32);
    c1 = (u32)s;
    c2 += (u32)(s >> 32);
#endif
}
SVB_HD void fr_acc_add(u32& c0, u32& c1, u32& c2, u32 a) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, 0;\n\t addc.u32 %2, %2, 0;" : "+r"(c0), "+r"(c1), "+r"(c2) : "r"(a));
#else
    u64 s = (u64)c0 + a;
    c0 = (u32)s;
    s = (u64)c1 + (s >> 32);
    c1 = (u32)s;
    c2 += (u32)(s >> 32);
#endif
}

// r = r - mod if r >= mod (r < 2 mod)
SVB_HD void fr_cond_sub(fr& r) {
#if defined(__CUDA_ARCH__)
    // one borrow chain, the final borrow as an all-ones mask, one LOP3 select per limb
    u32 d0, d1, d2, d3, d4, d5, d6, d7, keep;
    asm("sub.cc.u32 %0, %9, 0xf0000001;\n\t subc.cc.u32 %1, %10, 0x43e1f593;\n\t subc.cc.u32 %2, %11, 0x79b97091;\n\t"
        "subc.cc.u32 %3, %12, 0x2833e848;\n\t subc.cc.u32 %4, %13, 0x8181585d;\n\t subc.cc.u32 %5, %14, 0xb85045b6;\n\t"
        "subc.cc.u32 %6, %15, 0xe131a029;\n\t subc.cc.u32 %7, %16, 0x30644e72;\n\t subc.u32 %8, 0, 0;"
        : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3), "=r"(d4), "=r"(d5), "=r"(d6), "=r"(d7), "=r"(keep)
        : "r"(r.l[0]), "r"(r.l[1]), "r"(r.l[2]), "r"(r.l[3]), "r"(r.l[4]), "r"(r.l[5]), "r"(r.l[6]), "r"(r.l[7]));
    const u32 d[8] = {d0, d1, d2, d3, d4, d5, d6, d7};
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = (r.l[i] & keep) | (d[i] & ~keep);   // keep = 0xFFFFFFFF iff r < mod
#else
    u32 d[8];
    u32 borrow = 0;
    for (int i = 0; i < 8; i++) {
        u64 t = (u64)r.l[i] - fr_mod(i) - borrow;
        d[i] = (u32)t;
        borrow = (u32)(t >> 32) & 1u;
    }
    if (!borrow)
        for (int i = 0; i < 8; i++) r.l[i] = d[i];
#endif
}
SVB_HD fr fr_add(const fr& a, const fr& b) {
    fr r;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32 %0, %8, %16;\n\t addc.cc.u32 %1, %9, %17;\n\t addc.cc.u32 %2, %10, %18;\n\t addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t addc.cc.u32 %5, %13, %21;\n\t addc.cc.u32 %6, %14, %22;\n\t addc.u32 %7, %15, %23;"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
#else
    u32 carry = 0;
    for (int i = 0; i < 8; i++) {
        u64 t = (u64)a.l[i] + b.l[i] + carry;
        r.l[i] = (u32)t;
        carry = (u32)(t >> 32);
    }
#endif
    fr_cond_sub(r);   // a + b < 2r < 2^255: no carry out of the top limb
    return r;
}

// Montgomery reduction interleaved with the product scanning of sum_t a[t] * b[t] (N terms):
// returns sum * 2^-256 mod r.  Needs N * r^2 < r * 2^256, i.e. N <= 5.
// SVB_B_ACCS > 1 splits a column over several independent 96-bit accumulators (more ILP); measured on
// B200 (tools/lab, b_acc1/2/4: 60.3 / 55.4 / 53.3 M perms/s) the extra folding adds cost more than the
// shorter dependency chains gain, so the default is one accumulator.
#ifndef SVB_B_ACCS
#define SVB_B_ACCS 1
#endif
template <int N>
SVB_HD fr fr_dot_mont(const fr* a, const fr* b) {
#if defined(__CUDA_ARCH__)
    constexpr int A = (N == 1) ? (SVB_B_ACCS > 2 ? 2 : SVB_B_ACCS) : SVB_B_ACCS;
#else
    constexpr int A = 1;   // the host has out-of-order cores and no use for the extra accumulators
#endif
    u32 m[8];
    fr out;
    u32 c0[A], c1[A], c2[A];
#pragma unroll
    for (int q = 0; q < A; q++) c0[q] = c1[q] = c2[q] = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int n = 0;   // running MAC index inside the column (compile-time after unrolling)
#pragma unroll
        for (int t = 0; t < N; t++)
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (k - i >= 0 && k - i < 8) { fr_mac(c0[n % A], c1[n % A], c2[n % A], a[t].l[i], b[t].l[k - i]); n++; }
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (k - i >= 1 && k - i < 8 && i < k) { fr_mac(c0[n % A], c1[n % A], c2[n % A], m[i], fr_mod(k - i)); n++; }
        // fold the partial accumulators into accumulator 0
#pragma unroll
        for (int q = 1; q < A; q++) {
#if defined(__CUDA_ARCH__)
            asm("add.cc.u32 %0, %0, %3;\n\t addc.cc.u32 %1, %1, %4;\n\t addc.u32 %2, %2, %5;"
                : "+r"(c0[0]), "+r"(c1[0]), "+r"(c2[0]) : "r"(c0[q]), "r"(c1[q]), "r"(c2[q]));
#else
            u64 t0 = (u64)c0[0] + c0[q];
            u64 t1 = (u64)c1[0] + c1[q] + (t0 >> 32);
            c0[0] = (u32)t0; c1[0] = (u32)t1; c2[0] += c2[q] + (u32)(t1 >> 32);
#endif
            c0[q] = c1[q] = c2[q] = 0;
        }
        if (k < 8) {
            m[k] = c0[0] * FR_NINV32;
            fr_mac(c0[0], c1[0], c2[0], m[k], fr_mod(0));   // makes c0 == 0
        } else {
            out.l[k - 8] = c0[0];
        }
        c0[0] = c1[0]; c1[0] = c2[0]; c2[0] = 0;
    }
    // k = 15 left its carry in c0 (< 2): the value is < 2r < 2^255, so it is 0
    fr_cond_sub(out);
    return out;
}
SVB_HD fr fr_mmul(const fr& a, const fr& b) { return fr_dot_mont<1>(&a, &b); }
SVB_HD fr fr_pow5(const fr& x) {
    fr x2 = fr_mmul(x, x);
    fr x4 = fr_mmul(x2, x2);
    return fr_mmul(x4, x);
}
SVB_HD fr fr_const(const u64* tab, int idx) {
    fr r;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        u64 w = tab[4 * idx + k];
        r.l[2 * k] = (u32)w;
        r.l[2 * k + 1] = (u32)(w >> 32);
    }
    return r;
}

// encode_fe (native.rs:62-67): x0 + x1 p + x2 p^2 < 2^192 as an integer, then into Montgomery form
SVB_HD fr fr_encode3(u64 x0, u64 x1, u64 x2) {
    // w = x1 + p * x2 (< 2^128), v = x0 + p * w (< 2^192); p * y = (y << 64) - (y << 32) + y
    unsigned __int128 w = (unsigned __int128)x2 * GL_P + x1;
    u64 w0 = (u64)w, w1 = (u64)(w >> 64);
    unsigned __int128 lo = (unsigned __int128)w0 * GL_P + x0;
    unsigned __int128 hi = (unsigned __int128)w1 * GL_P + (u64)(lo >> 64);
    fr v, r2;
    v.l[0] = (u32)(u64)lo; v.l[1] = (u32)((u64)lo >> 32);
    v.l[2] = (u32)(u64)hi; v.l[3] = (u32)((u64)hi >> 32);
    v.l[4] = (u32)(u64)(hi >> 64); v.l[5] = (u32)((u64)(hi >> 64) >> 32);
    v.l[6] = 0; v.l[7] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = fr_r2(i);
    return fr_mmul(v, r2);
}
// decode_fe (native.rs:69-77): out of Montgomery form, then the first 3 base-p digits
SVB_HD void fr_decode3(const fr& s, u64 out[3]) {
    fr one;
#pragma unroll
    for (int i = 0; i < 8; i++) one.l[i] = i == 0 ? 1u : 0u;
    fr c = fr_mmul(s, one);   // canonical: REDC of a value < r
    u64 a[4];
#pragma unroll
    for (int k = 0; k < 4; k++) a[k] = ((u64)c.l[2 * k + 1] << 32) | c.l[2 * k];
    // long division by p, three times (schoolbook, 128/64 steps)
    for (int d = 0; d < 3; d++) {
        unsigned __int128 rem = 0;
        for (int i = 3; i >= 0; i--) {
            unsigned __int128 cur = (rem << 64) | a[i];
            a[i] = (u64)(cur / GL_P);
            rem = cur % GL_P;
        }
        out[d] = (u64)rem;
    }
}

// ================================================================================================
// Portable path (host: transcript and synthetic prover).  Canonical in, canonical out.
// ================================================================================================
SVB_HD void poseidon_b_fr(fr st[5]) {
    int counter = 0;
    for (int round = 0; round < 68; round++) {
        for (int i = 0; i < 5; i++) st[i] = fr_add(st[i], fr_const(SVB_TB(B_ROUND_CONSTANTS_MONT), counter++));
        int nl = (round < 4 || round >= 64) ? 5 : 1;
        for (int i = 0; i < nl; i++) st[i] = fr_pow5(st[i]);
        fr nw[5];
        for (int i = 0; i < 5; i++) {
            fr row[5];
            for (int j = 0; j < 5; j++) row[j] = fr_const(SVB_TB(B_MDS_MONT), 5 * i + j);
            nw[i] = fr_dot_mont<5>(st, row);
        }
        for (int i = 0; i < 5; i++) st[i] = nw[i];
    }
}
// The same permutation with the 60 partial rounds in sparse form (tools/gen_poseidon_bn254_constants.py
// derives and checks the constants): a matrix [[1,0],[0,B]] commutes with the lane-0 S-box, so every
// partial-round MDS factors into a sparse matrix [[m00, w^T],[v, I]] times such a block matrix that is
// handed to the round before; what remains is one dense 4x4 block in front (D0) and per round one 5-term
// dot product (lane 0) plus four multiply-adds instead of five 5-term dot products.
// Table rows (14 elements per partial round r): m00, w_hat_r[4], v_r[4], c'_r[5].
SVB_HD void poseidon_b_fr_sparse(fr st[5]) {
    const u64* RC = SVB_TB(B_ROUND_CONSTANTS_MONT);
    const u64* SP = SVB_TB(B_SPARSE_ROUNDS_MONT);
    for (int half = 0; half < 2; half++) {
        for (int round = half ? 64 : 0; round < (half ? 68 : 4); round++) {
            for (int i = 0; i < 5; i++) st[i] = fr_pow5(fr_add(st[i], fr_const(RC, 5 * round + i)));
            fr nw[5];
            for (int i = 0; i < 5; i++) {
                fr row[5];
                for (int j = 0; j < 5; j++) row[j] = fr_const(SVB_TB(B_MDS_MONT), 5 * i + j);
                nw[i] = fr_dot_mont<5>(st, row);
            }
            for (int i = 0; i < 5; i++) st[i] = nw[i];
        }
        if (half) break;
        // t_0 = D0 (s + c_0) = D0 s + c'_0
        fr t[5];
        t[0] = fr_add(st[0], fr_const(SP, 9));
        for (int i = 0; i < 4; i++) {
            fr row[4];
            for (int j = 0; j < 4; j++) row[j] = fr_const(SVB_TB(B_SPARSE_D0_MONT), 4 * i + j);
            t[1 + i] = fr_add(fr_dot_mont<4>(st + 1, row), fr_const(SP, 10 + i));
        }
        for (int r = 0; r < 60; r++) {
            fr a[5], b[5];
            a[0] = fr_pow5(t[0]);
            for (int j = 1; j < 5; j++) a[j] = t[j];
            for (int j = 0; j < 5; j++) b[j] = fr_const(SP, 14 * r + j);
            t[0] = fr_dot_mont<5>(a, b);
            for (int j = 1; j < 5; j++) t[j] = fr_add(a[j], fr_mmul(fr_const(SP, 14 * r + 4 + j), a[0]));
            if (r < 59)
                for (int j = 0; j < 5; j++) t[j] = fr_add(t[j], fr_const(SP, 14 * (r + 1) + 9 + j));
        }
        for (int i = 0; i < 5; i++) st[i] = t[i];
    }
}
inline void poseidon_b_canonical(u64 s[12]) {
    fr st[5];
    for (int k = 0; k < 4; k++) st[k] = fr_encode3(canon(s[3 * k]), canon(s[3 * k + 1]), canon(s[3 * k + 2]));
    for (int i = 0; i < 8; i++) st[4].l[i] = 0;
    poseidon_b_fr_sparse(st);
    for (int k = 0; k < 4; k++) fr_decode3(st[k], s + 3 * k);
}

#if defined(__CUDACC__)
// ================================================================================================
// Device path.  `sm` points at this thread's column of a [40][stride] u32 shared-memory array.
// ================================================================================================
#define SVB_B_SMEM_WORDS 40
SVB_D fr b_load(const u32* sm, u32 stride, int lane) {
    fr r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = sm[(lane * 8 + i) * stride];
    return r;
}
SVB_D void b_store(u32* sm, u32 stride, int lane, const fr& v) {
#pragma unroll
    for (int i = 0; i < 8; i++) sm[(lane * 8 + i) * stride] = v.l[i];
}
// canonical (or LOOSE) Goldilocks words in, canonical out
SVB_D void poseidon_b_dev(u64 s[12], u32* sm, u32 stride) {
#pragma unroll
    for (int k = 0; k < 4; k++) b_store(sm, stride, k, fr_encode3(canon(s[3 * k]), canon(s[3 * k + 1]), canon(s[3 * k + 2])));
    {
        fr z;
#pragma unroll
        for (int i = 0; i < 8; i++) z.l[i] = 0;
        b_store(sm, stride, 4, z);
    }
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        // four full rounds (native.rs:45-49 / :55-59): constant layer + x^5 on every lane, then the MDS
#pragma unroll 1
        for (int round = half ? 64 : 0; round < (half ? 68 : 4); round++) {
#pragma unroll 1
            for (int lane = 0; lane < 5; lane++)
                b_store(sm, stride, lane, fr_pow5(fr_add(b_load(sm, stride, lane), fr_const(d_B_ROUND_CONSTANTS_MONT, 5 * round + lane))));
            fr st[5];
#pragma unroll
            for (int j = 0; j < 5; j++) st[j] = b_load(sm, stride, j);
#pragma unroll 1
            for (int i = 0; i < 5; i++) {
                fr row[5];
#pragma unroll
                for (int j = 0; j < 5; j++) row[j] = fr_const(d_B_MDS_MONT, 5 * i + j);
                b_store(sm, stride, i, fr_dot_mont<5>(st, row));
            }
        }
        if (half) break;
        // the 60 partial rounds (native.rs:50-54) in sparse form, see poseidon_b_fr_sparse
        {
            fr st[5];
#pragma unroll
            for (int j = 0; j < 5; j++) st[j] = b_load(sm, stride, j);
            b_store(sm, stride, 0, fr_add(st[0], fr_const(d_B_SPARSE_ROUNDS_MONT, 9)));
#pragma unroll 1
            for (int i = 0; i < 4; i++) {
                fr row[4];
#pragma unroll
                for (int j = 0; j < 4; j++) row[j] = fr_const(d_B_SPARSE_D0_MONT, 4 * i + j);
                b_store(sm, stride, 1 + i, fr_add(fr_dot_mont<4>(st + 1, row), fr_const(d_B_SPARSE_ROUNDS_MONT, 10 + i)));
            }
        }
#pragma unroll 1
        for (int r = 0; r < 60; r++) {
            const u64* sp = d_B_SPARSE_ROUNDS_MONT + 4 * 14 * r;
            fr a[5], b[5];
            a[0] = fr_pow5(b_load(sm, stride, 0));
#pragma unroll
            for (int j = 1; j < 5; j++) a[j] = b_load(sm, stride, j);
#pragma unroll
            for (int j = 0; j < 5; j++) b[j] = fr_const(sp, j);
            fr n0 = fr_dot_mont<5>(a, b);
            const bool more = r < 59;
            b_store(sm, stride, 0, more ? fr_add(n0, fr_const(sp, 14 + 9)) : n0);
#pragma unroll 1
            for (int j = 1; j < 5; j++) {
                fr u = fr_add(b_load(sm, stride, j), fr_mmul(fr_const(sp, 4 + j), a[0]));
                b_store(sm, stride, j, more ? fr_add(u, fr_const(sp, 14 + 9 + j)) : u);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) fr_decode3(b_load(sm, stride, k), s + 3 * k);
}
#endif

}  // namespace svb
