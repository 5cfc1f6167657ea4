// Wire format <-> flat record.
//
// plonky2 serialises a ProofWithPublicInputs as a flat byte string with no framing other than one u8 length in
// front of every Merkle proof (util/serialization.rs of the plonky2 revision Cargo.lock pins: write_proof,
// write_opening_set, write_fri_proof, write_fri_query_rounds, write_merkle_proof, write_field = little-endian
// canonical u64).  For a fixed circuit every vector length is known from CommonCircuitData, so a proof of that
// circuit is a FIXED permutation of 8-byte words between the wire bytes and the flat record of layout.hpp.  That
// permutation is built once on the host as two small tables (WireMap) and applied
//   * by the device gather kernel (wire_kernels.cuh: one thread per record word), and
//   * by the host unpacker / packer (wire_host.cpp) -- through the SAME wire_record_word() below, which is what
//     the CPU tests exercise against the oracle's independent cursor-style reader.
// Field order mirrored: ProofValues (types/proof.rs:380-387), OpeningSetValues (:34-43), FriProofValues
// (:143-160 and the query-round types above it), VerificationKeyValues (types/verification_key.rs:9-12).
#pragma once
#include "../../include/stark_verifier_b200.h"
#include "goldilocks.cuh"
#include "layout.hpp"

#include <vector>

namespace svb {

// table codes: source byte offset of a record word (relative to the proof for header words, to the query round's
// first byte for query-block words), or one of
static constexpr u32 WIRE_ZERO = 0xFFFFFFFFu;     // padding / challenge field: written as 0
static constexpr u32 WIRE_VK_FLAG = 0x80000000u;  // | word index into the verifier key's constants_sigmas_cap

struct WireDims {
    u32 header_words, query_words, record_words, num_queries;
    u32 query_base;    // byte offset of the first query round inside a proof
    u32 query_bytes;   // bytes per query round
    u32 pi_off;        // byte offset of the public inputs
    u32 num_public_inputs;
    u32 proof_bytes;
    u32 n_chk;         // Merkle proofs per query round (4 + num_steps), each with one length byte
};

struct WireMap {
    WireDims d;
    std::vector<u32> hdr_src;          // [header_words]
    std::vector<u32> q_src;            // [query_words]
    std::vector<u32> chk;              // [n_chk]: (byte offset of the length byte inside the query round) << 8 | expected length
};

// 8 bytes at byte offset `off` from an 8-byte aligned base, little-endian.  Touches word off/8 and, when off is not
// a multiple of 8, word off/8 + 1 (both hold bytes of the value).
SVB_HD u64 wire_gather64(const u64* base, size_t off) {
    size_t wi = off >> 3;
    u32 sh = (u32)(off & 7) * 8;
    u64 lo = base[wi];
    if (sh == 0) return lo;
    u64 hi = base[wi + 1];
    return (lo >> sh) | (hi << (64 - sh));
}
SVB_HD u32 wire_byte(const u64* base, size_t off) { return (u32)(base[off >> 3] >> ((off & 7) * 8)) & 0xFFu; }

// One source code: offset of a byte run, WIRE_ZERO (padding / challenge field -> 0) or a word of the verifier key's cap.
SVB_HD u64 wire_fetch(u32 code, const u64* vk_cap, const u64* base, size_t from) {
    if (code == WIRE_ZERO) return 0;
    if (code & WIRE_VK_FLAG) return vk_cap[code & ~WIRE_VK_FLAG];
    return wire_gather64(base, from + code);
}
// Header word w (< header_words) of the proof whose bytes start `proof_off` bytes after `base`.
SVB_HD u64 wire_header_word(const u32* hdr_src, const u64* vk_cap, const u64* base, size_t proof_off, u32 w) {
    return wire_fetch(hdr_src[w], vk_cap, base, proof_off);
}
// Word r (< query_words) of query round q.  The first n_chk words of every query block also compare one Merkle-proof
// length byte each with the depth the shape implies and set *malformed when it differs (plonky2 would read a different
// number of siblings and fail further on).
SVB_HD u64 wire_query_word(const WireDims& d, const u32* q_src, const u32* chk, const u64* vk_cap, const u64* base, size_t proof_off,
                           u32 q, u32 r, bool* malformed) {
    const size_t from = proof_off + d.query_base + (size_t)q * d.query_bytes;
    if (r < d.n_chk) {
        const u32 c = chk[r];
        if (wire_byte(base, from + (c >> 8)) != (c & 0xFFu)) *malformed = true;
    }
    return wire_fetch(q_src[r], vk_cap, base, from);
}
// Record word `w` of a proof: the two above behind one index (the host unpacker's loop).
SVB_HD u64 wire_record_word(const WireDims& d, const u32* hdr_src, const u32* q_src, const u32* chk,
                            const u64* vk_cap, const u64* base, size_t proof_off, u32 w, bool* malformed) {
    if (w < d.header_words) return wire_header_word(hdr_src, vk_cap, base, proof_off, w);
    const u32 r = w - d.header_words, q = r / d.query_words;
    return wire_query_word(d, q_src, chk, vk_cap, base, proof_off, q, r - q * d.query_words, malformed);
}

// Host: build the tables.  Returns 0, or < 0 when shape and common disagree.
int make_wire_map(const sv_fri_shape& s, const sv_plonk_common& c, WireMap& M);

}  // namespace svb
