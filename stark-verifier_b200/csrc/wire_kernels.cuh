// Device side of the wire format: serialised plonky2 proofs -> flat records, on the GPU.
//
// The wire bytes cross PCIe as they are; unpacking is a fixed permutation of 8-byte words (wire.hpp) and therefore
// pure HBM-bound byte shuffling: a thread reads the (unaligned) 8 source bytes of a record word as two aligned
// words and a funnel shift, and writes one coalesced word.  Consecutive threads read consecutive source words, so a
// warp request covers 256-264 contiguous bytes; the offset tables (a few KB) stay in L1.  Algorithmic bytes per
// proof: wire bytes read + record bytes written (shape A: 156 812 + 157 728).
#pragma once
#include "fri_kernels.cuh"
#include "wire.hpp"

namespace svb {

#define SVB_WIRE_BLOCK 256

// grid = (n_proofs, 1 + num_queries): block (p, 0) writes the header of proof p, block (p, 1 + q) its query round q; the
// threads stride over the words of that segment (no division anywhere).  blob8: 8-byte aligned base; proof p starts at
// byte first_off + p * stride.  malformed[p] is set to 1 when a Merkle-proof length byte is wrong (zeroed by the caller).
__global__ void __launch_bounds__(SVB_WIRE_BLOCK) wire_unpack_kernel(const u64* __restrict__ blob8, size_t first_off, size_t stride,
                                                                     WireDims d, const u32* __restrict__ hdr_src,
                                                                     const u32* __restrict__ q_src, const u32* __restrict__ chk,
                                                                     const u64* __restrict__ vk_cap, u64* __restrict__ records,
                                                                     u32* __restrict__ malformed, const u64* __restrict__ hdr_packed,
                                                                     u32 y0 /* segment of block row 0: 0 = header, 1 = first query round */) {
    const size_t p = blockIdx.x, proof_off = first_off + p * stride;
    u64* rec = records + p * (size_t)d.record_words;
    const u32 seg = blockIdx.y + y0;
    if (seg == 0) {
        if (hdr_packed) {   // batch-wide transcript: the header was unpacked (and its challenge fields filled) before the chunks
            const u64* h = hdr_packed + p * (size_t)d.header_words;
            for (u32 w = threadIdx.x; w < d.header_words; w += SVB_WIRE_BLOCK) rec[w] = h[w];
            return;
        }
        for (u32 w = threadIdx.x; w < d.header_words; w += SVB_WIRE_BLOCK) rec[w] = wire_header_word(hdr_src, vk_cap, blob8, proof_off, w);
        return;
    }
    const u32 q = seg - 1;
    u64* out = rec + d.header_words + (size_t)q * d.query_words;
    bool bad = false;
    for (u32 r = threadIdx.x; r < d.query_words; r += SVB_WIRE_BLOCK)
        out[r] = wire_query_word(d, q_src, chk, vk_cap, blob8, proof_off, q, r, &bad);
    if (bad) malformed[p] = 1;
}

// The headers of a whole batch from their two byte spans (sv_verify_proofs_wire / _full): a serialised proof keeps its
// header fields at both ends -- caps, openings and commit-phase caps in front of the query rounds, final polynomial, PoW
// witness and public inputs behind them -- so the host sends those two spans of EVERY proof first (two strided copies),
// this kernel gathers them into packed headers (header_words per proof), and the transcript runs once for the batch while
// the query rounds are still crossing PCIe.  front8 / back8: row p at p * pitch bytes (8-byte aligned rows); a source code
// below front_bytes lives in the front span, any other in the back span, which starts at byte back_off of the proof.
__global__ void __launch_bounds__(SVB_WIRE_BLOCK) wire_header_unpack_kernel(const u64* __restrict__ front8, size_t front_pitch,
                                                                            const u64* __restrict__ back8, size_t back_pitch,
                                                                            u32 front_bytes, u32 back_off, u32 header_words,
                                                                            const u32* __restrict__ hdr_src, const u64* __restrict__ vk_cap,
                                                                            u64* __restrict__ hdr_out) {
    const size_t p = blockIdx.x;
    u64* out = hdr_out + p * (size_t)header_words;
    for (u32 w = threadIdx.x; w < header_words; w += SVB_WIRE_BLOCK) {
        const u32 code = hdr_src[w];
        u64 v;
        if (code == WIRE_ZERO) v = 0;
        else if (code & WIRE_VK_FLAG) v = vk_cap[code & ~WIRE_VK_FLAG];
        else if (code < front_bytes) v = wire_gather64(front8, p * front_pitch + code);
        else v = wire_gather64(back8, p * back_pitch + (code - back_off));
        out[w] = v;
    }
}

// Public-inputs hash, one thread per proof: hash_n_to_hash_no_pad over Poseidon-Goldilocks
// (PlonkVerifierChip::get_public_inputs_hash, chip/plonk/plonk_verifier_chip.rs:41-53; PublicInputsHasherChip::hash,
// chip/public_inputs_hasher_chip.rs:315-341).  Zero public inputs hash to the zero digest (no permutation).  A public
// input >= p marks the proof malformed (the reference range-checks every assigned value,
// native_chip/arithmetic_chip.rs:256-268; the record words are checked by the query phase itself).
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS) wire_pi_hash_kernel(const u64* __restrict__ blob8, size_t first_off, size_t stride,
                                                                                WireDims d, size_t n, u64* __restrict__ pi_hashes,
                                                                                u32* __restrict__ malformed) {
    __shared__ u64 scratch[PermScratch<SV_HASH_POSEIDON_GOLDILOCKS>::array_len(SVB_BLOCK)];
    size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    size_t at = first_off + p * stride + d.pi_off;
    u64 s[12];
#pragma unroll
    for (int k = 0; k < 12; k++) s[k] = 0;
    bool bad = false;
#pragma unroll 1
    for (u32 off = 0; off < d.num_public_inputs; off += 8) {
        u32 rem = d.num_public_inputs - off;
#pragma unroll
        for (int k = 0; k < 8; k++)
            if ((u32)k < rem) {
                s[k] = wire_gather64(blob8, at + 8 * (size_t)(off + k));
                bad |= !is_canonical(s[k]);
            }
        permute_dev<SV_HASH_POSEIDON_GOLDILOCKS>(s, scratch, SVB_BLOCK);
    }
    if (bad) malformed[p] = 1;
    ulonglong2* o = reinterpret_cast<ulonglong2*>(pi_hashes + 4 * p);
    o[0] = make_ulonglong2(canon(s[0]), canon(s[1]));
    o[1] = make_ulonglong2(canon(s[2]), canon(s[3]));
}

// After the query phase: a malformed proof is rejected whatever its record happened to verify as.
__global__ void wire_reject_malformed_kernel(const u32* __restrict__ malformed, u32* __restrict__ accept_bitmap,
                                             u32* __restrict__ first_fail, u32 n) {
    u32 p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n || !malformed[p]) return;
    atomicAnd(accept_bitmap + (p >> 5), ~(1u << (p & 31)));
    if (first_fail) first_fail[p] = SV_FAIL_MALFORMED;
}

}  // namespace svb
