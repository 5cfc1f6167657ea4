// Host twin of the NTT / LDE kernels: the same phase functions (ntt.hpp: ntt_tile_load / ntt_tile_round / ntt_tile_store)
// that ntt_pass_kernel runs between barriers, here with every phase looped over the "threads" of a tile, on CPU
// threads.  For the CPU test-suite (tests/test_ntt.py: against naive big-integer evaluation) and for callers that
// transform a handful of small polynomials.
#include "ntt.hpp"
#include "host_util.hpp"

using namespace svb;

// one pass over every tile of `n_items` (polynomial, coset) items
static void host_pass(const NttPass& P, size_t n_items, const u64* src, size_t src_stride, u64* dst, size_t dst_stride, const u64* tw,
                      const u64* scale_lo, const u64* scale_hi, int nthreads) {
    const u64 tiles = ntt_tiles_per_poly(P);
    const NttRounds Q = ntt_rounds(P);
    const u32 VT = 64;   // virtual threads per tile: any number gives the same result, > 1 exercises the strided loops
    parallel_for(n_items * tiles, nthreads, [&](size_t b, size_t e) {
        std::vector<u64> tile(NTT_SMEM_WORDS);
        for (size_t it = b; it < e; it++) {
            const size_t item = it / tiles;
            const u64 t = it % tiles;
            const NttTileMap M = ntt_tile_map(P, t);
            const u64* s = src + (item >> P.coset_bits) * src_stride;
            u64* d = dst + item * dst_stride;
            const u32 coset = (u32)(item & ((1u << P.coset_bits) - 1));
            for (u32 tid = 0; tid < VT; tid++) ntt_tile_load(P, M, s, scale_lo, scale_hi, coset, tile.data(), tid, VT);
            for (u32 i = 0; i < Q.n; i++)
                for (u32 tid = 0; tid < VT; tid++) ntt_tile_round_any(P, M, tw, tile.data(), Q.b[i], Q.r[i], tid, VT);
            for (u32 tid = 0; tid < VT; tid++) ntt_tile_store(P, M, d, tile.data(), tid, VT);
        }
    });
}

static int ntt_host_impl(u32 k, size_t n_polys, u64* data, bool inverse, int nthreads) {
    NttPass plan[NTT_MAX_PASSES];
    int np = ntt_plan(k, inverse, plan);
    if (np < 0) return -1;
    const u64 n = 1ull << k;
    std::vector<u64> tw(std::max<u64>(1, n / 2));
    ntt_twiddles(k, inverse, tw.data());
    for (int q = 0; q < np; q++) host_pass(plan[q], n_polys, data, n, data, n, tw.data(), nullptr, nullptr, nthreads);
    return 0;
}

extern "C" int sv_ntt_host(uint32_t log_n, size_t n_polys, uint64_t* data, int inverse, int nthreads) {
    if (!data && n_polys) return -1;
    if (log_n == 0 || log_n > 26) return -2;
    for (size_t i = 0; i < (n_polys << log_n); i++)
        if (!is_canonical(data[i])) return -3;
    return ntt_host_impl(log_n, n_polys, data, inverse != 0, nthreads < 1 ? 1 : nthreads);
}

extern "C" int sv_lde_host(uint32_t log_n, uint32_t rate_bits, size_t n_polys, const uint64_t* coeffs, uint64_t shift,
                           uint64_t* out, int nthreads) {
    if ((!coeffs || !out) && n_polys) return -1;
    if (log_n == 0 || log_n + rate_bits > 26 || !is_canonical(shift) || shift == 0) return -2;
    const u64 n = 1ull << log_n;
    for (size_t i = 0; i < n_polys * n; i++)
        if (!is_canonical(coeffs[i])) return -3;
    // 2^rate_bits size-n transforms per polynomial, the first pass scaling on load (ntt.hpp)
    NttPass plan[NTT_MAX_PASSES];
    int np = ntt_plan(log_n, false, plan);
    if (np < 0) return -2;
    const u32 h = lde_scale_h(log_n);
    std::vector<u64> lo((size_t)1 << (rate_bits + h)), hi((size_t)1 << (rate_bits + log_n - h)), tw(std::max<u64>(1, n / 2));
    lde_scale_tables(log_n, rate_bits, shift, lo.data(), hi.data());
    ntt_twiddles(log_n, false, tw.data());
    const size_t items = n_polys << rate_bits;
    if (nthreads < 1) nthreads = 1;
    for (int q = 0; q < np; q++) {
        NttPass P = plan[q];
        P.coset_bits = rate_bits;
        P.scale_h = h;
        P.load_scaled = q == 0;
        if (q == 0) host_pass(P, items, coeffs, n, out, n, tw.data(), lo.data(), hi.data(), nthreads);
        else {
            P.coset_bits = 0;   // in place on the output blocks from now on
            host_pass(P, items, out, n, out, n, tw.data(), nullptr, nullptr, nthreads);
        }
    }
    return 0;
}
