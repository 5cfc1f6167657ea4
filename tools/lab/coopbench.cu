// Lab harness: the lane-cooperative permutations (poseidon_g_coop = v1, poseidon_g_coop2 = latency form) against the host
// permutation, their latency on a lone warp per SM and their time for transcript-like batches (155 dependent permutations
// for 416 .. 4096 proofs).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr coopbench.cu
#include "../../stark-verifier_b200/csrc/fri_kernels.cuh"
#include "poseidon_g_coop_v1.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace svb;

__global__ void __launch_bounds__(128) coop1_chain(const u64* __restrict__ in, u64* __restrict__ out, size_t n_states, int depth) {
    __shared__ CoopTables T;
    coop_load_tables(T);
    const int l = threadIdx.x & 15;
    size_t g = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 4;
    const bool valid = g < n_states;
    if (!valid) g = n_states - 1;
    u64 s = l < 12 ? in[12 * g + l] : 0;
    for (int d = 0; d < depth; d++) s = canon(poseidon_g_coop(s, l, T));
    if (l < 12 && valid) out[12 * g + l] = s;
}
__global__ void __launch_bounds__(128) coop2_chain(const u64* __restrict__ in, u64* __restrict__ out, size_t n_states, int depth) {
    __shared__ Coop2Tables<8> T;
    coop2_load_tables(T);
    struct { int l; } L = {(int)(threadIdx.x & 15)};
    size_t g = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 4;
    const bool valid = g < n_states;
    if (!valid) g = n_states - 1;
    u64 s = L.l < 12 ? in[12 * g + L.l] : 0;
    for (int d = 0; d < depth; d++) s = canon(poseidon_g_coop2(s, L.l, T, threadIdx.x >> 4));
    if (L.l < 12 && valid) out[12 * g + L.l] = s;
}

int main() {
    const size_t n = 8192;
    std::vector<u64> h(12 * n), got(12 * n);
    u64 x = 0x9E3779B97F4A7C15ull;
    for (auto& v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = x % GL_P; }
    for (int k = 0; k < 12; k++) { h[k] = 0; h[12 + k] = GL_P - 1; h[24 + k] = k; h[36 + k] = (k & 1) ? GL_P - 1 : 0xFFFFFFFFull; h[48 + k] = 0xFFFFFFFF00000000ull; }
    u64 *din, *dout;
    cudaMalloc(&din, 12 * n * 8); cudaMalloc(&dout, 12 * n * 8);
    cudaMemcpy(din, h.data(), 12 * n * 8, cudaMemcpyHostToDevice);
    size_t bad_total = 0;
    for (int ver = 1; ver <= 2; ver++) {
        for (int depth = 1; depth <= 3; depth += 2) {
            cudaMemset(dout, 0, 12 * n * 8);
            if (ver == 1) coop1_chain<<<(unsigned)(n * 16 / 128), 128>>>(din, dout, n, depth);
            else coop2_chain<<<(unsigned)(n * 16 / 128), 128>>>(din, dout, n, depth);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 2; }
            cudaMemcpy(got.data(), dout, 12 * n * 8, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (size_t i = 0; i < n; i++) {
                u64 s[12];
                for (int k = 0; k < 12; k++) s[k] = h[12 * i + k];
                for (int d = 0; d < depth; d++) poseidon_g_canonical(s);
                for (int k = 0; k < 12; k++) bad += s[k] != got[12 * i + k];
            }
            printf("coop v%d depth %d: mismatching words %zu of %zu\n", ver, depth, bad, 12 * n);
            bad_total += bad;
        }
    }
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](auto launch) {
        float b = 1e30f, ms;
        for (int r = 0; r < 4; r++) { cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); if (r && ms < b) b = ms; }
        return b;
    };
    const int cd = 16;
    for (int ver = 1; ver <= 2; ver++) {
        float lone = timeit([&] { if (ver == 1) coop1_chain<<<148, 32>>>(din, dout, 296, cd); else coop2_chain<<<148, 32>>>(din, dout, 296, cd); });
        printf("coop v%d lone warp per SM: %.2f us per permutation\n", ver, lone * 1e3 / cd);
        const size_t sizes[] = {208, 416, 832, 1664, 2432, 4096, 8192};
        for (size_t np : sizes) {
            float t = timeit([&] { if (ver == 1) coop1_chain<<<(unsigned)((np * 16 + 127) / 128), 128>>>(din, dout, np, 155); else coop2_chain<<<(unsigned)((np * 16 + 127) / 128), 128>>>(din, dout, np, 155); });
            printf("coop v%d transcript-like: %5zu proofs x 155 permutations: %.3f ms  (%.1f M perms/s, %.2f us per permutation)\n", ver, np, t,
                   np * 155.0 / t / 1e3, t * 1e3 / 155);
        }
    }
    return bad_total ? 1 : 0;
}
