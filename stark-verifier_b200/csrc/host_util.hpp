// Small host-side helpers shared by host_side.cpp (transcript, synthetic prover) and wire_host.cpp (wire format).
#pragma once
#include "../../include/stark_verifier_b200.h"
#include "goldilocks.cuh"
#include "poseidon_g.cuh"
#include "poseidon_b.cuh"

#include <algorithm>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

namespace svb {

// the width-12 permutation of a hash family, canonical in / canonical out
static inline void permute_kind(u32 kind, u64 st[12]) {
    if (kind == SV_HASH_POSEIDON_BN254) poseidon_b_canonical(st);
    else poseidon_g_canonical(st);
}

// hash_n_to_hash_no_pad: overwrite-mode sponge, rate 8 (chip/hasher_chip.rs:122-147)
static inline void hash_no_pad(u32 kind, const u64* in, size_t n, u64 out[4]) {
    u64 st[12] = {0};
    for (size_t off = 0; off < n; off += 8) {
        size_t len = std::min<size_t>(8, n - off);
        for (size_t i = 0; i < len; i++) st[i] = in[off + i];
        permute_kind(kind, st);
    }
    memcpy(out, st, 32);
}

static inline void parallel_for(size_t n, int nthreads, const std::function<void(size_t, size_t)>& body) {
    if (nthreads <= 1 || n < 2) { body(0, n); return; }
    std::vector<std::thread> th;
    size_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        size_t b = std::min(n, (size_t)t * chunk), e = std::min(n, b + chunk);
        if (b < e) th.emplace_back([=, &body] { body(b, e); });
    }
    for (auto& x : th) x.join();
}

}  // namespace svb
