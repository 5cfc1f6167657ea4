// Host twin of the NTT / LDE kernels: the same ntt_pass_block (ntt.hpp) with one "thread" per block, on CPU threads.
// For the CPU test-suite and for callers that transform a handful of small polynomials.
#include "ntt.hpp"
#include "host_util.hpp"

using namespace svb;

struct NoSync { void operator()() const {} };

static int ntt_host_impl(u32 k, size_t n_polys, u64* data, bool inverse, int nthreads) {
    NttPass plan[8];
    int np = ntt_plan(k, inverse, plan);
    if (np < 0) return -1;
    const u64 n = 1ull << k;
    std::vector<u64> tw(std::max<u64>(1, n / 2));
    ntt_twiddles(k, inverse, tw.data());
    const u64 ninv = inverse ? inv(n % GL_P) : 1;
    parallel_for(n_polys, nthreads, [&](size_t b, size_t e) {
        std::vector<u64> tile((size_t)1 << NTT_TILE_LOG);
        for (size_t p = b; p < e; p++) {
            u64* poly = data + p * n;
            for (int q = 0; q < np; q++)
                for (u64 blk = 0; blk < ntt_blocks_per_poly(plan[q]); blk++)
                    ntt_pass_block(plan[q], poly, tw.data(), tile.data(), blk, 0, 1, NoSync());
            if (inverse)
                for (u64 i = 0; i < n; i++) poly[i] = mulc(poly[i], ninv);
        }
    });
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
    const u64 n = 1ull << log_n, N = 1ull << (log_n + rate_bits);
    for (size_t i = 0; i < n_polys * n; i++)
        if (!is_canonical(coeffs[i])) return -3;
    for (size_t p = 0; p < n_polys; p++)
        for (u64 j = 0; j < N; j++) out[p * N + j] = lde_scaled_coeff(coeffs + p * n, n, shift, j);
    return ntt_host_impl(log_n + rate_bits, n_polys, out, false, nthreads < 1 ? 1 : nthreads);
}
