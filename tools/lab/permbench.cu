// Lab harness: time poseidon_permute_kernel / merkle-style chains for one build configuration and check
// the outputs against the host path.  Build with -D switches (see tools/lab/run_variants.sh).
#include "../../stark-verifier_b200/csrc/fri_kernels.cuh"
#include "poseidon_g_coop_v1.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace svb;
#ifndef LAB_KIND
#define LAB_KIND 0   // 0 = Poseidon-Goldilocks, 1 = Poseidon-BN254 wrapped
#endif

// chain of `depth` dependent permutations per thread (Merkle-like: no memory traffic in the loop)
__global__ void __launch_bounds__(SVB_BLOCK, SVB_MINBLOCKS) chain_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n, int depth) {
    __shared__ u64 scratch[PermScratch<LAB_KIND>::array_len(SVB_BLOCK)];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 s[12];
    for (int k = 0; k < 12; k++) s[k] = in[12 * i + k];
    for (int d = 0; d < depth; d++) {
        permute_dev<LAB_KIND>(s, scratch, SVB_BLOCK);
        for (int k = 4; k < 12; k++) s[k] = canon(s[k]) ^ (u64)d;   // keep lanes live, cheap
        for (int k = 4; k < 12; k++) s[k] = s[k] >= GL_P ? s[k] - GL_P : s[k];
    }
    for (int k = 0; k < 12; k++) out[12 * i + k] = canon(s[k]);
}

// lane-cooperative permutation: one 16-lane group per state, `depth` dependent permutations
__global__ void __launch_bounds__(128) coop_chain_kernel(const u64* __restrict__ in, u64* __restrict__ out, size_t n_states, int depth) {
    __shared__ CoopTables T;
    coop_load_tables(T);
    const int l = threadIdx.x & 15;
    size_t g = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 4;
    if (g >= n_states) g = n_states - 1;
    u64 s = l < 12 ? in[12 * g + l] : 0;
    for (int d = 0; d < depth; d++) s = canon(poseidon_g_coop(s, l, T));
    if (l < 12) out[12 * g + l] = s;
}
// one thread per state, launched with ONE warp per SM: the latency of a lone warp
__global__ void __launch_bounds__(32) lone_chain_kernel(const u64* __restrict__ in, u64* __restrict__ out, int depth) {
    __shared__ u64 scratch[11 * 32];
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    u64 s[12];
    for (int k = 0; k < 12; k++) s[k] = in[12 * i + k];
    for (int d = 0; d < depth; d++) poseidon_g_dev(s, scratch + threadIdx.x, 32);
    for (int k = 0; k < 12; k++) out[12 * i + k] = canon(s[k]);
}

__global__ void field_kernel(const u64* a, const u64* b, const u64* c, u64* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[2 * i] = canon(mul(a[i], b[i]));
    out[2 * i + 1] = canon(mul_add(a[i], b[i], c[i]));
}
static size_t field_corner_test() {
    // operands whose products hit the rare limb patterns of red5 (borrow with r2 = 0, carry-out of T, ...)
    const u64 sp[] = {0, 1, 2, 0xFFFFFFFFull, 0x100000000ull, 0x100000001ull, 0xFFFFFFFF00000000ull, GL_P - 1, GL_P, GL_P + 1,
                      0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFEFFFFFFFFull, 1ull << 48, 5ull << 48, 0x8000000000000000ull, 0x00000001FFFFFFFFull,
                      0xFFFFFFFF00000001ull, 0x0000FFFF0000FFFFull, 3ull << 62, 0x7FFFFFFF80000001ull};
    const int m = sizeof sp / sizeof sp[0];
    std::vector<u64> a, b, c;
    for (int i = 0; i < m; i++) for (int j = 0; j < m; j++) for (int k = 0; k < m; k += 3) { a.push_back(sp[i]); b.push_back(sp[j]); c.push_back(sp[k]); }
    int n = (int)a.size();
    u64 *da, *db, *dc, *dout;
    cudaMalloc(&da, n * 8); cudaMalloc(&db, n * 8); cudaMalloc(&dc, n * 8); cudaMalloc(&dout, 2 * n * 8);
    cudaMemcpy(da, a.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(db, b.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dc, c.data(), n * 8, cudaMemcpyHostToDevice);
    field_kernel<<<(n + 127) / 128, 128>>>(da, db, dc, dout, n);
    std::vector<u64> out(2 * n);
    cudaMemcpy(out.data(), dout, 2 * n * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (int i = 0; i < n; i++) {
        unsigned __int128 pr = (unsigned __int128)a[i] * b[i];
        u64 w1 = (u64)(pr % GL_P), w2 = (u64)((pr % GL_P + c[i] % GL_P) % GL_P);
        if (out[2 * i] != w1 || out[2 * i + 1] != w2) {
            if (bad < 5) printf("  field mismatch a=%016llx b=%016llx c=%016llx got %016llx %016llx want %016llx %016llx\n",
                                (unsigned long long)a[i], (unsigned long long)b[i], (unsigned long long)c[i],
                                (unsigned long long)out[2 * i], (unsigned long long)out[2 * i + 1], (unsigned long long)w1, (unsigned long long)w2);
            bad++;
        }
    }
    return bad;
}

int main(int argc, char** argv) {
    size_t n = LAB_KIND ? 148 * 4 * 128 * 2 : 148 * 5 * 128 * 8;   // whole waves of blocks
    int depth = argc > 1 ? atoi(argv[1]) : 20;
    std::vector<u64> h(12 * n), ref(12 * n), got(12 * n);
    u64 x = 0x9E3779B97F4A7C15ull;
    for (auto& v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x % GL_P; }
    // edge states
    for (int k = 0; k < 12; k++) { h[k] = 0; h[12 + k] = GL_P - 1; h[24 + k] = k; h[36 + k] = (k & 1) ? GL_P - 1 : 0xFFFFFFFFull; }
    u64 *din, *dout;
    cudaMalloc(&din, 12 * n * 8); cudaMalloc(&dout, 12 * n * 8);
    cudaMemcpy(din, h.data(), 12 * n * 8, cudaMemcpyHostToDevice);
    // correctness: single permutation kernel vs host on the first 4096 states
    poseidon_permute_kernel<LAB_KIND><<<(unsigned)((n + SVB_BLOCK - 1) / SVB_BLOCK), SVB_BLOCK>>>(din, dout, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
    cudaMemcpy(got.data(), dout, 12 * n * 8, cudaMemcpyDeviceToHost);
    size_t bad = 0;
    for (size_t i = 0; i < (LAB_KIND ? 256 : 4096); i++) {
        u64 s[12];
        for (int k = 0; k < 12; k++) s[k] = h[12 * i + k];
        if (LAB_KIND) poseidon_b_canonical(s); else poseidon_g_canonical(s);
        for (int k = 0; k < 12; k++) bad += s[k] != got[12 * i + k];
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        chain_kernel<<<(unsigned)((n + SVB_BLOCK - 1) / SVB_BLOCK), SVB_BLOCK>>>(din, dout, n, depth);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
#if LAB_KIND == 0
    {   // lane-cooperative vs thread-per-permutation: correctness, lone-warp latency, full-grid throughput
        const int cd = 8;
        size_t ns = 148 * 8 * 16;                      // 16 blocks of 128 threads per SM, 8 states per block
        coop_chain_kernel<<<(unsigned)(ns * 16 / 128), 128>>>(din, dout, ns, 1);
        cudaDeviceSynchronize();
        std::vector<u64> cg(12 * 64);
        cudaMemcpy(cg.data(), dout, cg.size() * 8, cudaMemcpyDeviceToHost);
        size_t cbad = 0;
        for (size_t i = 0; i < 64; i++) {
            u64 s[12];
            for (int k = 0; k < 12; k++) s[k] = h[12 * i + k];
            poseidon_g_canonical(s);
            for (int k = 0; k < 12; k++) cbad += s[k] != cg[12 * i + k];
        }
        float ms;
        auto timeit = [&](auto launch) { float b = 1e30f; for (int r = 0; r < 3; r++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (r && ms < b) b = ms; } return b; };
        float t_coop_full = timeit([&] { coop_chain_kernel<<<(unsigned)(ns * 16 / 128), 128>>>(din, dout, ns, cd); });
        float t_coop_lone = timeit([&] { coop_chain_kernel<<<148, 32>>>(din, dout, 148 * 2, cd); });
        float t_thr_lone = timeit([&] { lone_chain_kernel<<<148, 32>>>(din, dout, cd); });
        printf("  coop: mismatches %zu; full grid %.1f Mperm/s; lone-warp latency coop %.1f us/perm, thread-per-perm %.1f us/perm\n", cbad,
               ns * (double)cd / t_coop_full / 1e3, t_coop_lone * 1e3 / cd, t_thr_lone * 1e3 / cd);
        bad += cbad;
    }
#endif
    size_t fbad = field_corner_test();
    if (fbad) printf("  FIELD CORNER MISMATCHES: %zu\n", fbad);
    bad += fbad;
    int regs = 0; cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, chain_kernel); regs = fa.numRegs;
    printf("%-28s mismatches %zu  chain: %.3f ms  %.1f Mperm/s  regs %d\n", VARIANT, bad, best, n * (double)depth / best / 1e3, regs);
    return bad ? 1 : 0;
}
