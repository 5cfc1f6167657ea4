// C ABI (include/stark_verifier_b200.h): context, launches, host<->device pipeline.
// No CPU fallback: every compute entry point needs a CUDA device and fails loudly otherwise.
#include "../../include/stark_verifier_b200.h"
#include "fri_kernels.cuh"
#include "wire_kernels.cuh"
#include "plonk_kernels.cuh"
#include "ntt_kernels.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace svb;

// SV_MEM_HOST pipeline depth: staging buffers in flight (H2D of chunk i+2.. while chunks i, i+1 compute)
#define SV_NBUF 6
#define SV_NKS 4    // compute streams the chunk kernels rotate over

struct sv_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;   // compute
    cudaStream_t aux_stream[SV_NKS - 1] = {};   // further compute streams of the SV_MEM_HOST pipeline
    cudaStream_t copy_stream = nullptr;  // H2D of the next chunks
    cudaStream_t stream = nullptr;       // stream used for SV_MEM_DEVICE work (own or caller's)
    // device scratch, grown on demand, reused across calls (no hidden allocation after warm-up)
    u64* d_scratch = nullptr; size_t scratch_words = 0;          // reduced openings
    cudaEvent_t ev_scratch = nullptr; cudaStream_t scratch_stream = nullptr; bool scratch_busy = false;   // last SV_MEM_DEVICE user of d_scratch
    u64* d_stage[SV_NBUF] = {}; size_t stage_words[SV_NBUF] = {};   // H2D chunk ring
    u64* d_leaf[SV_NBUF] = {}; size_t leaf_words[SV_NBUF] = {};     // leaf digests of the chunk in d_stage[b] (fri_leaf_kernel)
    u64* d_leaf_dev = nullptr; size_t leaf_dev_words = 0;           // the same for a SV_MEM_DEVICE batch
    u32* d_bitmap = nullptr; size_t bitmap_words = 0;
    u32* d_fail = nullptr; size_t fail_words = 0;
    u64* d_pi = nullptr; size_t pi_words = 0;                          // public-input hashes (device-side transcript)
    u64 *d_hfront = nullptr, *d_hback = nullptr; size_t hfront_words = 0, hback_words = 0;   // header byte spans of a wire batch
    u64* d_hdr = nullptr; size_t hdr_words = 0;                        // record headers of a whole host batch (transcript)
    cudaStream_t fs_stream = nullptr;                                  // the batch-wide transcript of the host pipeline
    cudaEvent_t ev_hdr = nullptr, ev_fs = nullptr;
    cudaStream_t fs_part_stream[2] = {};                               // transcript parts 1 and 2 of a wire batch (part 0: fs_stream)
    cudaEvent_t ev_part[3] = {}, ev_plonk[3] = {}, ev_hdr_part[3] = {}, ev_hdr_ready = nullptr;   // per transcript part: challenges written / plonk identity checked
    cudaEvent_t ev_copied[SV_NBUF] = {}, ev_done[SV_NBUF] = {}, ev_join[SV_NKS] = {};
    cudaStream_t prep_stream = nullptr;                                 // high priority: the short kernels in front of a chunk's query kernel
    cudaEvent_t ev_prep[SV_NBUF] = {};
    // wire format (sv_wire_unpack_batch_gpu, sv_verify_proofs_wire): offset tables of the last (shape, common) seen,
    // the verifier key's cap, wire-byte staging ring, malformed flags
    bool w_valid = false;
    sv_fri_shape w_shape; sv_plonk_common w_common;
    WireMap w_map; std::vector<u64> w_vk;
    u32* d_wtab = nullptr; size_t wtab_words = 0;                      // hdr_src | q_src | chk
    u64* d_wvk = nullptr; size_t wvk_words = 0;
    u64* d_wire[SV_NBUF] = {}; size_t wire_words[SV_NBUF] = {};
    u32* d_mal = nullptr; size_t mal_words = 0;
    // plonk-level checks: the circuit description on the device, staging for the host path
    sv_plonk_circuit* d_circuit = nullptr; sv_plonk_circuit h_circuit; bool circuit_valid = false;
    u64* d_chal = nullptr; size_t chal_words = 0;
    u32* d_pbm = nullptr; size_t pbm_words = 0;
    // NTT twiddles of the last (log_n, direction) used
    u64* d_tw = nullptr; size_t tw_words = 0; u32 tw_k = 0; int tw_inverse = -1;
    u64 *d_lde_lo = nullptr, *d_lde_hi = nullptr; size_t lde_lo_words = 0, lde_hi_words = 0; u32 lde_log_n = 0, lde_rate_bits = 0; u64 lde_shift = 0;   // LDE scale tables (ntt.hpp)
                           // plonk-identity bitmap of sv_verify_proofs_full
    uint64_t launches = 0;
    // optional CUDA-event timing of the dominant kernel (fri_query_kernel / merkle / permute), on
    // the stream it is launched on
    bool timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tev; size_t tev_used = 0;
    std::string err;
    int sm_count = 0;
    // dlopen'ed NCCL
    void* nccl_lib = nullptr;
    int (*ncclAllGather_fn)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
};

static thread_local std::string g_create_err;   // errors raised without a context (sv_ctx_create, sv_host_alloc)

// launch KERNEL<hash kind>(args) for a run-time hash kind
#define SVB_LAUNCH_KIND(kind, KERNEL, grid, block, stream, ...)                                         \
    do {                                                                                                \
        if ((kind) == SV_HASH_POSEIDON_BN254) KERNEL<SV_HASH_POSEIDON_BN254><<<(grid), (block), 0, (stream)>>>(__VA_ARGS__); \
        else KERNEL<SV_HASH_POSEIDON_GOLDILOCKS><<<(grid), (block), 0, (stream)>>>(__VA_ARGS__);    \
    } while (0)
static bool known_mem(int m) { return m == SV_MEM_HOST || m == SV_MEM_DEVICE; }
static bool known_kind(int k) { return k == SV_HASH_POSEIDON_GOLDILOCKS || k == SV_HASH_POSEIDON_BN254; }

static int fail(sv_ctx* c, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}
#define CK(c, call)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) return fail((c), -100 - (int)e_, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// bump SVB_KERNEL_REV whenever a kernel changes: profiles/traffic_r2.json and the ncu summaries are stamped with it, and bench.py
// reports DRAM traffic only from a capture of the same revision
#define SVB_KERNEL_REV "r2.5"
extern "C" const char* sv_version(void) { return "stark-verifier_b200 0.2 (sm_100a, kernels " SVB_KERNEL_REV ")"; }

extern "C" const char* sv_last_error(const sv_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" void sv_ctx_destroy(sv_ctx* c);

extern "C" int sv_ctx_create(int device, sv_ctx** out) {
    if (!out) return -1;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(nullptr, -2, "no CUDA device (%s); this library has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= n) return fail(nullptr, -3, "device %d out of range (%d devices)", device, n);
    CK(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(nullptr, -4, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    sv_ctx* c = new sv_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    // on failure the partially built context is torn down again (destroy tolerates null handles)
    auto init = [&]() -> int {
        CK(nullptr, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        for (int i = 0; i < SV_NKS - 1; i++) CK(nullptr, cudaStreamCreateWithFlags(&c->aux_stream[i], cudaStreamNonBlocking));
        CK(nullptr, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        {   // the transcript is a chain of dependent permutations on a few warps: its blocks must not queue behind the
            // thousands of query-kernel blocks of another chunk or another call, so its stream gets the highest priority
            int lo = 0, hi = 0;
            CK(nullptr, cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CK(nullptr, cudaStreamCreateWithPriority(&c->fs_stream, cudaStreamNonBlocking, hi));
            for (auto& st : c->fs_part_stream) CK(nullptr, cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, hi));
            CK(nullptr, cudaStreamCreateWithPriority(&c->prep_stream, cudaStreamNonBlocking, hi));
        }
        CK(nullptr, cudaEventCreateWithFlags(&c->ev_hdr, cudaEventDisableTiming));
        CK(nullptr, cudaEventCreateWithFlags(&c->ev_fs, cudaEventDisableTiming));
        CK(nullptr, cudaEventCreateWithFlags(&c->ev_hdr_ready, cudaEventDisableTiming));
        for (auto& e : c->ev_part) CK(nullptr, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : c->ev_plonk) CK(nullptr, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : c->ev_hdr_part) CK(nullptr, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (int i = 0; i < SV_NBUF; i++) {
            CK(nullptr, cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
            CK(nullptr, cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
            CK(nullptr, cudaEventCreateWithFlags(&c->ev_prep[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < SV_NKS; i++) CK(nullptr, cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
        return 0;
    };
    if (int rc = init()) {
        sv_ctx_destroy(c);
        return rc;
    }
    // lab knobs (tools/lab/traffic_probe.sh): prefetch flavour of the Merkle chains, L2 fetch granularity
    if (const char* e = getenv("SVB_PREFETCH")) {
        u32 mode = (u32)atoi(e);
        if (mode <= 2) cudaMemcpyToSymbol(d_PREFETCH_MODE, &mode, sizeof mode);
    }
    if (const char* e = getenv("SVB_L2_FETCH")) {
        int g = atoi(e);
        if (g == 32 || g == 64 || g == 128) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)g);
    }
    c->stream = c->own_stream;
    *out = c;
    return 0;
}

extern "C" void sv_ctx_destroy(sv_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    cudaFree(c->d_scratch);
    for (int i = 0; i < SV_NBUF; i++) { cudaFree(c->d_stage[i]); cudaFree(c->d_leaf[i]); }
    cudaFree(c->d_leaf_dev);
    cudaFree(c->d_bitmap);
    cudaFree(c->d_fail);
    cudaFree(c->d_pi);
    cudaFree(c->d_hdr);
    cudaFree(c->d_hfront);
    cudaFree(c->d_hback);
    cudaFree(c->d_wtab);
    cudaFree(c->d_wvk);
    for (int i = 0; i < SV_NBUF; i++) cudaFree(c->d_wire[i]);
    cudaFree(c->d_mal);
    cudaFree(c->d_circuit);
    cudaFree(c->d_chal);
    cudaFree(c->d_pbm);
    if (c->ev_scratch) cudaEventDestroy(c->ev_scratch);
    cudaFree(c->d_tw);
    cudaFree(c->d_lde_lo);
    cudaFree(c->d_lde_hi);
    auto drop_s = [](cudaStream_t s) { if (s) cudaStreamDestroy(s); };
    auto drop_e = [](cudaEvent_t e) { if (e) cudaEventDestroy(e); };
    drop_e(c->ev_hdr); drop_e(c->ev_fs); drop_e(c->ev_hdr_ready);
    for (auto& e : c->ev_part) drop_e(e);
    for (auto& e : c->ev_plonk) drop_e(e);
    for (auto& e : c->ev_hdr_part) drop_e(e);
    for (int i = 0; i < SV_NBUF; i++) { drop_e(c->ev_copied[i]); drop_e(c->ev_done[i]); drop_e(c->ev_prep[i]); }
    for (int i = 0; i < SV_NKS; i++) drop_e(c->ev_join[i]);
    for (auto& pr : c->tev) { drop_e(pr.first); drop_e(pr.second); }
    drop_s(c->fs_stream);
    for (auto& st : c->fs_part_stream) drop_s(st);
    drop_s(c->own_stream);
    for (int i = 0; i < SV_NKS - 1; i++) drop_s(c->aux_stream[i]);
    drop_s(c->prep_stream);
    drop_s(c->copy_stream);
    if (c->nccl_lib) dlclose(c->nccl_lib);
    delete c;
}

extern "C" int sv_ctx_set_stream(sv_ctx* c, void* s) {
    if (!c) return -1;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return 0;
}
extern "C" int sv_ctx_synchronize(sv_ctx* c) {
    if (!c) return -1;
    CK(c, cudaSetDevice(c->device));
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaStreamSynchronize(c->own_stream));
    for (int i = 0; i < SV_NKS - 1; i++) CK(c, cudaStreamSynchronize(c->aux_stream[i]));
    if (c->prep_stream) CK(c, cudaStreamSynchronize(c->prep_stream));
    CK(c, cudaStreamSynchronize(c->copy_stream));
    CK(c, cudaStreamSynchronize(c->fs_stream));
    for (auto& st : c->fs_part_stream) CK(c, cudaStreamSynchronize(st));
    return 0;
}
extern "C" uint64_t sv_ctx_launch_count(const sv_ctx* c) { return c ? c->launches : 0; }

extern "C" int sv_host_alloc(size_t bytes, void** out) {
    if (!out) return -1;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocDefault);
    return e == cudaSuccess ? 0 : fail(nullptr, -100 - (int)e, "cudaHostAlloc: %s", cudaGetErrorString(e));
}
extern "C" int sv_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? 0 : -1; }

static cudaEvent_t time_begin(sv_ctx* c, cudaStream_t s) {
    if (!c->timing) return nullptr;
    if (c->tev_used == c->tev.size()) {
        cudaEvent_t a, b;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return nullptr;
        c->tev.push_back({a, b});
    }
    cudaEventRecord(c->tev[c->tev_used].first, s);
    return c->tev[c->tev_used].second;
}
static void time_end(sv_ctx* c, cudaEvent_t e, cudaStream_t s) {
    if (!e) return;
    cudaEventRecord(e, s);
    c->tev_used++;
}

extern "C" int sv_ctx_kernel_timing(sv_ctx* c, int enable) {
    if (!c) return -1;
    c->timing = enable != 0;
    c->tev_used = 0;
    return 0;
}
// Sum of the recorded durations (ms) of the dominant kernel since the last call; synchronises.
extern "C" int sv_ctx_kernel_time_ms(sv_ctx* c, double* total_ms, uint64_t* n_launches) {
    if (!c || !total_ms || !n_launches) return -1;
    CK(c, cudaSetDevice(c->device));
    double tot = 0;
    for (size_t i = 0; i < c->tev_used; i++) {
        CK(c, cudaEventSynchronize(c->tev[i].second));
        float ms = 0;
        CK(c, cudaEventElapsedTime(&ms, c->tev[i].first, c->tev[i].second));
        tot += ms;
    }
    *total_ms = tot;
    *n_launches = c->tev_used;
    c->tev_used = 0;
    return 0;
}

template <typename T>
static int grow(sv_ctx* c, T*& ptr, size_t& have, size_t need) {
    if (need <= have) return 0;
    if (ptr) CK(c, cudaFree(ptr));
    ptr = nullptr; have = 0;
    CK(c, cudaMalloc((void**)&ptr, need * sizeof(T)));
    have = need;
    return 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" int sv_poseidon_permute_batch(sv_ctx* c, const uint64_t* in, uint64_t* out, size_t n, int hash_kind, int mem) {
    if (!c || !in || !out) return -1;
    if (!known_kind(hash_kind)) return fail(c, -5, "hash_kind %d not implemented", hash_kind);
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (n == 0) return 0;
    CK(c, cudaSetDevice(c->device));
    const int B = SVB_BLOCK;
    if (mem == SV_MEM_DEVICE) {
        cudaEvent_t te = time_begin(c, c->stream);
        SVB_LAUNCH_KIND(hash_kind, poseidon_permute_kernel, (unsigned)((n + B - 1) / B), B, c->stream, in, out, n);
        time_end(c, te, c->stream);
        c->launches++;
        CK(c, cudaGetLastError());
        return 0;
    }
    size_t words = 12 * n;
    if (grow(c, c->d_stage[0], c->stage_words[0], words)) return -6;
    cudaStream_t s = c->own_stream;
    CK(c, cudaMemcpyAsync(c->d_stage[0], in, words * 8, cudaMemcpyHostToDevice, s));
    SVB_LAUNCH_KIND(hash_kind, poseidon_permute_kernel, (unsigned)((n + B - 1) / B), B, s, c->d_stage[0], c->d_stage[0], n);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, c->d_stage[0], words * 8, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
}

extern "C" int sv_poseidon_permute_batch_coop(sv_ctx* c, const uint64_t* in, uint64_t* out, size_t n, int mem) {
    if (!c || !in || !out) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (n == 0) return 0;
    CK(c, cudaSetDevice(c->device));
    const unsigned grid = (unsigned)((n * SVB_COOP2_GROUP + SVB_COOP_BLOCK - 1) / SVB_COOP_BLOCK);
    if (mem == SV_MEM_DEVICE) {
        poseidon_permute_coop_kernel<<<grid, SVB_COOP_BLOCK, 0, c->stream>>>(in, out, n);
        c->launches++;
        CK(c, cudaGetLastError());
        return 0;
    }
    size_t words = 12 * n;
    if (grow(c, c->d_stage[0], c->stage_words[0], 2 * words)) return -6;
    cudaStream_t s = c->own_stream;
    CK(c, cudaMemcpyAsync(c->d_stage[0], in, words * 8, cudaMemcpyHostToDevice, s));
    poseidon_permute_coop_kernel<<<grid, SVB_COOP_BLOCK, 0, s>>>(c->d_stage[0], c->d_stage[0] + words, n);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, c->d_stage[0] + words, words * 8, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
}

extern "C" int sv_goldilocks_mul_add_batch(sv_ctx* c, const uint64_t* a, const uint64_t* b, const uint64_t* cc, uint64_t* out,
                                           size_t n, int mem) {
    if (!c || !a || !b || !cc || !out) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (n == 0) return 0;
    CK(c, cudaSetDevice(c->device));
    const int B = 256;
    if (mem == SV_MEM_DEVICE) {
        goldilocks_mul_add_kernel<<<(unsigned)((n + B - 1) / B), B, 0, c->stream>>>(a, b, cc, out, n);
        c->launches++;
        CK(c, cudaGetLastError());
        return 0;
    }
    if (grow(c, c->d_stage[0], c->stage_words[0], 4 * n)) return -6;
    u64* d = c->d_stage[0];
    cudaStream_t s = c->own_stream;
    CK(c, cudaMemcpyAsync(d, a, n * 8, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(d + n, b, n * 8, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(d + 2 * n, cc, n * 8, cudaMemcpyHostToDevice, s));
    goldilocks_mul_add_kernel<<<(unsigned)((n + B - 1) / B), B, 0, s>>>(d, d + n, d + 2 * n, d + 3 * n, n);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(out, d + 3 * n, n * 8, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
}

extern "C" int sv_merkle_verify_batch(sv_ctx* c, uint32_t leaf_len, uint32_t depth, uint32_t cap_height, int hash_kind,
                                      const uint64_t* paths, const uint64_t* indices, const uint64_t* caps, uint8_t* ok,
                                      size_t n, int mem) {
    if (!c || !paths || !indices || !caps || !ok) return -1;
    if (!known_kind(hash_kind)) return fail(c, -5, "hash_kind %d not implemented", hash_kind);
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (leaf_len == 0 || depth > 63 || cap_height > 16 || depth + cap_height > 63) return fail(c, -7, "bad merkle shape");
    if (n == 0) return 0;
    CK(c, cudaSetDevice(c->device));
    const int B = SVB_BLOCK;
    size_t rec_words = up4(leaf_len) + 4 * (size_t)depth;
    if (mem == SV_MEM_DEVICE) {
        cudaEvent_t te = time_begin(c, c->stream);
        SVB_LAUNCH_KIND(hash_kind, merkle_verify_kernel, (unsigned)((n + B - 1) / B), B, c->stream, paths, indices, caps, ok, n, leaf_len,
                        depth, cap_height);
        time_end(c, te, c->stream);
        c->launches++;
        CK(c, cudaGetLastError());
        return 0;
    }
    // host buffers: one shot (this entry point is the micro-benchmark path; the FRI entry point pipelines)
    size_t cap_words = 4ull << cap_height;
    size_t total = rec_words * n + n + cap_words + (n + 7) / 8 + 4;
    if (grow(c, c->d_stage[0], c->stage_words[0], total)) return -6;
    u64* d_paths = c->d_stage[0];
    u64* d_idx = d_paths + rec_words * n;
    u64* d_caps = d_idx + n;
    unsigned char* d_ok = reinterpret_cast<unsigned char*>(d_caps + cap_words);
    cudaStream_t s = c->own_stream;
    CK(c, cudaMemcpyAsync(d_paths, paths, rec_words * n * 8, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(d_idx, indices, n * 8, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(d_caps, caps, cap_words * 8, cudaMemcpyHostToDevice, s));
    SVB_LAUNCH_KIND(hash_kind, merkle_verify_kernel, (unsigned)((n + B - 1) / B), B, s, d_paths, d_idx, d_caps, d_ok, n, leaf_len, depth,
                    cap_height);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(ok, d_ok, n, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
}

// layers_out: leaf digests (n x 4 words), then every level up to the cap (2^cap_height x 4 words)
extern "C" int sv_merkle_tree_build(sv_ctx* c, int hash_kind, uint32_t leaf_len, const uint64_t* leaves, size_t n_leaves,
                                    uint32_t cap_height, uint64_t* layers_out, int mem) {
    if (!c || !leaves || !layers_out) return -1;
    if (!known_kind(hash_kind)) return fail(c, -5, "hash_kind %d not implemented", hash_kind);
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (leaf_len == 0 || n_leaves == 0 || (n_leaves & (n_leaves - 1)) || cap_height > 30 || ((size_t)1 << cap_height) > n_leaves)
        return fail(c, -7, "bad tree shape (n_leaves must be a power of two >= 2^cap_height)");
    CK(c, cudaSetDevice(c->device));
    const int B = SVB_BLOCK;
    const size_t ncap = (size_t)1 << cap_height, out_words = 4 * (2 * n_leaves - ncap);
    const u64* d_leaves = leaves;
    u64* d_out = layers_out;
    cudaStream_t s = mem == SV_MEM_DEVICE ? c->stream : c->own_stream;
    if (mem != SV_MEM_DEVICE) {
        size_t in_words = n_leaves * (size_t)leaf_len;
        if (grow(c, c->d_stage[0], c->stage_words[0], in_words)) return -6;
        if (grow(c, c->d_stage[1], c->stage_words[1], out_words)) return -6;
        CK(c, cudaMemcpyAsync(c->d_stage[0], leaves, in_words * 8, cudaMemcpyHostToDevice, s));
        d_leaves = c->d_stage[0];
        d_out = c->d_stage[1];
    }
    SVB_LAUNCH_KIND(hash_kind, merkle_leaf_hash_kernel, (unsigned)((n_leaves + B - 1) / B), B, s, d_leaves, leaf_len, n_leaves, d_out);
    c->launches++;
    u64* cur = d_out;
    for (size_t m = n_leaves; m > ncap; m >>= 1) {
        u64* nxt = cur + 4 * m;
        SVB_LAUNCH_KIND(hash_kind, merkle_level_kernel, (unsigned)((m / 2 + B - 1) / B), B, s, cur, nxt, m / 2);
        c->launches++;
        cur = nxt;
    }
    CK(c, cudaGetLastError());
    if (mem != SV_MEM_DEVICE) {
        CK(c, cudaMemcpyAsync(layers_out, d_out, out_words * 8, cudaMemcpyDeviceToHost, s));
        CK(c, cudaStreamSynchronize(s));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
static int make_params(sv_ctx* c, const sv_fri_shape& s, FriKernelParams& P) {
    memset(&P, 0, sizeof P);
    if (make_layout(s, P.L)) return fail(c, -8, "bad FRI shape");
    if (!known_kind((int)s.hash_kind)) return fail(c, -5, "hash_kind %u not implemented", s.hash_kind);
    P.hash_kind = s.hash_kind;
    if (s.proof_of_work_bits > 63) return fail(c, -8, "proof_of_work_bits > 63");
    if (s.num_query_rounds >= (1u << 12)) return fail(c, -8, "num_query_rounds too large");
    P.num_queries = s.num_query_rounds;
    P.num_steps = s.num_steps;
    P.final_poly_len = s.final_poly_len;
    P.pow_bits = s.proof_of_work_bits;
    for (int k = 0; k < 4; k++) P.oracle_num_polys[k] = s.oracle_num_polys[k];
    P.num_zs = s.num_zs;
    P.n_classes = 4 + s.num_steps + 1;
    // grid order of fri_query_kernel: phase A = the four oracle trees, heaviest first (cost in permutations),
    // then the algebra chain; phase B = the step trees, deepest first
    std::vector<std::pair<u32, u32>> cost;
    for (u32 k = 0; k < 4; k++) cost.push_back({(P.L.leaf_len[k] > 4 ? (P.L.leaf_len[k] + 7) / 8 : 0) + P.L.init_depth, k});
    std::stable_sort(cost.begin(), cost.end(), [](const std::pair<u32, u32>& a, const std::pair<u32, u32>& b) { return a.first > b.first; });
    cost.push_back({2, 4 + s.num_steps});
    P.n_classes_a = 5;
    std::vector<std::pair<u32, u32>> steps;
    for (u32 i = 0; i < s.num_steps; i++) {
        const u32 leaf = 2u << P.L.step_arity_bits[i];        // 2^k evaluations: hashed when more than 4 words
        steps.push_back({(leaf > 4 ? (leaf + 7) / 8 : 0) + P.L.step_depth[i], 4 + i});
    }
    std::stable_sort(steps.begin(), steps.end(), [](const std::pair<u32, u32>& a, const std::pair<u32, u32>& b) { return a.first > b.first; });
    cost.insert(cost.end(), steps.begin(), steps.end());
    for (u32 i = 0; i < P.n_classes; i++) P.class_order[i] = cost[i].second;
    for (u32 i = 0; i < 4; i++)                                  // fri_leaf_kernel: the hashed oracle leaves, heaviest first
        if (P.L.leaf_len[cost[i].second] > 4) P.leaf_class_order[P.n_leaf_classes++] = cost[i].second;
    u64 omega = svb::pow(7, (GL_P - 1) >> P.L.lde_bits);
    for (u32 i = 0; i < P.L.lde_bits; i++) { P.omega_pow2[i] = omega; omega = mulc(omega, omega); }
    return 0;
}

// enqueue prepare + query (+ finalize) for `n` proofs whose records are at d_records
// force: 0 = choose by batch size, 1 = lane-cooperative kernel (lowest latency), 2 = one thread per proof (most throughput)
static int enqueue_challenges(sv_ctx* c, FriKernelParams& P, const FsParams& F, size_t n, u64* d_records, const u64* d_pi,
                              cudaStream_t s, int force = 0) {
    P.n_proofs = (u32)n;
    // SVB_FS_COOP: 1 = always lane-cooperative, 0 = never, unset = by batch size.  The cooperative kernel has the
    // lower latency (6.1 vs 24.4 us per permutation) but 4.6x less throughput (252 vs 1 177 M perms/s), so it wins
    // while the batch is latency-bound: below ~8k proofs per call.
    static const int coop_env = [] { const char* e = getenv("SVB_FS_COOP"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const bool coop = force ? force == 1 : (coop_env < 0 ? n < 8192 : coop_env == 1);
    if (P.hash_kind == SV_HASH_POSEIDON_GOLDILOCKS && coop) {
        // lane-cooperative transcript: 16 lanes per proof
        fri_challenges_coop_kernel<<<(unsigned)((n * SVB_COOP2_GROUP + SVB_COOP_BLOCK - 1) / SVB_COOP_BLOCK), SVB_COOP_BLOCK, 0, s>>>(d_records, P, F, d_pi);
        c->launches++;
        CK(c, cudaGetLastError());
        return 0;
    }
    SVB_LAUNCH_KIND(P.hash_kind, fri_challenges_kernel, (unsigned)((n + SVB_FS_BLOCK - 1) / SVB_FS_BLOCK), SVB_FS_BLOCK, s, d_records, P,
                    F, d_pi);
    c->launches++;
    CK(c, cudaGetLastError());
    return 0;
}

// Proofs in the LEAD part of a device-side transcript (the part on the lane-cooperative kernel, whose query kernels then hide
// the thread-per-proof transcript of the rest): ~1 664 by default -- the cooperative kernel is still latency-bound there (1.36 ms
// against 1.0 ms for a lone warp per sub-partition) and their query phase (3.8 ms) outlasts the second part (3.8 ms + its
// start).  Rounded up to whole chunks.  SVB_FS_LEAD_PROOFS overrides (lab knob).
static size_t fs_lead_proofs(size_t n_proofs, size_t chunk) {
    static const long env = [] { const char* e = getenv("SVB_FS_LEAD_PROOFS"); return e ? atol(e) : 1664L; }();
    size_t lead = env > 0 ? (size_t)env : 1664;
    if (chunk) lead = (lead + chunk - 1) / chunk * chunk;
    return std::min(n_proofs, lead);
}

// Leaf digests before the challenges (fri_leaf_kernel), for the paths whose Fiat-Shamir transcript runs on the device: the
// challenge-independent quarter of the query phase can then run beside the transcript's latency.  OFF unless SVB_LEAF_SPLIT=1
// (read per call): measured on B200 (tools/lab/leaf_split.sh, tools/lab/NOTES.md) the leaf warps share the SM sub-partitions
// with the transcript's few latency-bound warps and slow those down by more than the overlap gains -- wire path 291 k against
// 299 k proofs/s fused, resident batch with device transcript 265 k against 334 k.  Kept as a knob and as a parity-tested
// second decomposition of the query phase (tests/test_gpu_leaf_split.py).
static bool leaf_split_enabled(const FriKernelParams& P) {
    const char* e = getenv("SVB_LEAF_SPLIT");
    return P.n_leaf_classes > 0 && e && atoi(e) != 0;
}
static size_t leaf_digest_words(const FriKernelParams& P, size_t n) { return n * P.num_queries * 16; }
static int enqueue_leaf(sv_ctx* c, FriKernelParams& P, size_t n, const u64* d_records, u64* d_leaf, cudaStream_t s) {
    P.n_proofs = (u32)n;
    P.n_units = (u32)(n * P.num_queries);
    P.blocks_per_class = (P.n_units + SVB_BLOCK - 1) / SVB_BLOCK;
    SVB_LAUNCH_KIND(P.hash_kind, fri_leaf_kernel, P.blocks_per_class * P.n_leaf_classes, SVB_BLOCK, s, d_records, P, d_leaf);
    c->launches++;
    CK(c, cudaGetLastError());
    return 0;
}

// sp / ev_sp: when given, fri_prepare_kernel (a dozen blocks) runs on that (high-priority) stream and the query kernel waits for the
// event -- in the host pipelines the SMs are full of the previous chunk's query blocks, and a short kernel queued behind them on
// an ordinary stream would hold this chunk's query kernel back until that grid has drained
// Lab knob SVB_QUERY_BPS: blocks of fri_query_kernel<G> per SM, capped through unused dynamic shared memory (unset / 0 = the launch
// bound's 6).  The idea was to shorten the drain of a chunk in the host pipelines (a chain's permutations take longer the more warps
// share its sub-partition: 29 dependent permutations = 2.1 ms at six warps, 0.7 ms alone).  Measured on B200 (tools/lab/bps_pass.sh):
// resident throughput 438.7 k (6) -> 432.0 k (4) -> 421.0 k (3) -> 389.8 k (2) proofs/s, wire path 302.6 / 306.2 / 305.1 / 292.4 k --
// no gain end to end, a loss resident, so nothing sets it.
static size_t query_dyn_smem(int bps) {
    static const int env = [] { const char* e = getenv("SVB_QUERY_BPS"); return e ? atoi(e) : -1; }();
    if (env >= 0) bps = env;
    if (bps <= 0 || bps >= SVB_MINBLOCKS) return 0;
    if (bps < 2) bps = 2;
    static const bool attr = [] {
        return cudaFuncSetAttribute(fri_query_kernel<SV_HASH_POSEIDON_GOLDILOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) == cudaSuccess;
    }();
    if (!attr) return 0;
    return ((size_t)227 * 1024 / (size_t)bps - 1024) & ~(size_t)1023;
}
static int enqueue_fri(sv_ctx* c, FriKernelParams& P, size_t n, const u64* d_records, u64* d_scratch, u32* d_bitmap,
                       u32* d_fail, cudaStream_t s, const u64* d_leaf = nullptr, cudaStream_t sp = nullptr, cudaEvent_t ev_sp = nullptr,
                       int bps = 0) {
    const int B = SVB_BLOCK;
    P.n_proofs = (u32)n;
    P.n_units = (u32)(n * P.num_queries);
    P.blocks_per_class = (P.n_units + B - 1) / B;
    // phase A unit groups of 32 blocks per class: the 160 blocks of a group are co-resident on the 148 SMs
    static const u32 group_env = [] { const char* e = getenv("SVB_GROUP_BLOCKS"); return e ? (u32)atoi(e) : 32u; }();
    P.group_blocks = group_env == 0 || group_env > P.blocks_per_class ? P.blocks_per_class : group_env;   // 0: class-major over the batch
    P.n_groups = (P.blocks_per_class + P.group_blocks - 1) / P.group_blocks;
    fri_prepare_kernel<<<(unsigned)((n + 31) / 32), SVB_PREP_BLOCK, 0, sp ? sp : s>>>(d_records, P, d_scratch, d_bitmap, d_fail);
    if (sp) {
        CK(c, cudaEventRecord(ev_sp, sp));
        CK(c, cudaStreamWaitEvent(s, ev_sp, 0));
    }
    cudaEvent_t te = time_begin(c, s);
    const u32 grid = P.n_groups * P.group_blocks * P.n_classes_a + P.blocks_per_class * (P.n_classes - P.n_classes_a);
    const size_t dsm = P.hash_kind == SV_HASH_POSEIDON_GOLDILOCKS ? query_dyn_smem(bps) : 0;
    if (dsm) fri_query_kernel<SV_HASH_POSEIDON_GOLDILOCKS><<<grid, B, dsm, s>>>(d_records, P, d_scratch, d_bitmap, d_fail, d_leaf);
    else SVB_LAUNCH_KIND(P.hash_kind, fri_query_kernel, grid, B, s, d_records, P, d_scratch, d_bitmap, d_fail, d_leaf);
    time_end(c, te, s);
    c->launches += 2;
    if (d_fail) {
        fri_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_fail, (u32)n);
        c->launches++;
    }
    CK(c, cudaGetLastError());
    return 0;
}

// SVB_PREP_STREAM=0: the short kernels of a chunk on the chunk's own compute stream, as before (lab knob)
static bool prep_stream_enabled() {
    static const bool on = [] { const char* e = getenv("SVB_PREP_STREAM"); return e ? atoi(e) != 0 : true; }();
    return on;
}

// Chunk schedule of a host batch: full chunks, then a ramp-down (1/2, 1/4, ... of a chunk, whole 32-proof bitmap words, never
// below 64 proofs).  The pipeline is copy-bound (the query kernels of a chunk take less time than its H2D), so a call ends one
// chunk-processing time after its last byte arrived -- and a chunk cannot be processed faster than its longest chain (28
// dependent permutations, 0.67 ms), however small it is.  With the ramp the last chunk carries ~100 proofs instead of ~400: its
// bytes arrive ~0.8 ms before the end of the copies and the work queued behind the copy engine at that point is a quarter.
// Measured on B200 (4 096 shape-A proofs): record path 316.4 -> 321.8 k proofs/s, with device transcript 276.0 -> 279.6 k; the
// wire path LOSES (303.4 -> 295.4 k, complete verifier 294.5 -> 284.9 k: its 64 MiB chunks ramp down into five small ones, each
// with its strided copy and seven launches), so the ramp is on for the record path only.  SVB_RAMP=0 / 1 forces it off / on
// everywhere (lab knob).  Returns the start of every chunk plus the end sentinel.
static std::vector<size_t> chunk_schedule(size_t n_proofs, size_t chunk, bool ramp_default) {
    const char* e = getenv("SVB_RAMP");                     // read per call: the tests toggle it
    const bool ramp = e ? atoi(e) != 0 : ramp_default;
    const char* em = getenv("SVB_RAMP_MIN");                // smallest ramp-down chunk, proofs (lab knob)
    const size_t ramp_min = em && atol(em) >= 32 ? ((size_t)atol(em) + 31) & ~(size_t)31 : 64;
    std::vector<size_t> starts;
    size_t at = 0;
    while (at < n_proofs) {
        starts.push_back(at);
        size_t left = n_proofs - at, sz = chunk;
        if (ramp && left <= 2 * chunk && left > ramp_min) {
            sz = ((left / 2) + 31) & ~(size_t)31;
            if (sz < ramp_min) sz = ramp_min;
        }
        at += std::min(sz, left);
    }
    starts.push_back(n_proofs);
    return starts;
}

// SV_MEM_HOST leg of sv_fri_verify_batch[_fs]: the chunked H2D / compute pipeline.
static int fri_verify_host(sv_ctx* c, FriKernelParams& P, size_t n_proofs, const uint64_t* records, uint32_t* accept_bitmap,
                           uint32_t* first_fail, const FsParams* fs, const uint64_t* pi_hashes) {
    int rc = 0;
    const size_t rw = P.L.record_words;
    if (fs && grow(c, c->d_pi, c->pi_words, 4 * n_proofs)) return -6;

    // Chunks of whole 32-proof bitmap words move through a ring of SV_NBUF staging
    // buffers.  H2D copies run back to back on the copy stream; the kernels of consecutive chunks
    // rotate over SV_NKS compute streams so that the tail wave of chunk i overlaps the head of
    // chunk i+1 (a chunk is only ~1.5 waves of blocks).  Distinct chunks touch distinct bitmap
    // words / first_fail entries / scratch rows, so they need no ordering among themselves.
    size_t n_words = (n_proofs + 31) / 32;
    if (grow(c, c->d_bitmap, c->bitmap_words, n_words)) return -6;
    if (first_fail && grow(c, c->d_fail, c->fail_words, n_proofs)) return -6;
    size_t chunk_mb = 32;                                        // ~32 MiB per chunk (swept on B200: tools/lab/e2e_sweep.sh)
    int n_ks = SV_NKS;                                           // compute streams in rotation
    if (const char* e = getenv("SVB_CHUNK_MB")) { long v = atol(e); if (v >= 1 && v <= 4096) chunk_mb = (size_t)v; }
    if (const char* e = getenv("SVB_KSTREAMS")) { long v = atol(e); if (v >= 1 && v <= SV_NKS) n_ks = (int)v; }
    size_t chunk = ((chunk_mb << 20) / (rw * 8)) & ~(size_t)31;
    if (chunk < 32) chunk = 32;
    if (chunk > n_proofs) chunk = (n_proofs + 31) & ~(size_t)31;
    const bool split = fs && leaf_split_enabled(P);
    for (int b = 0; b < SV_NBUF; b++) {
        if (grow(c, c->d_stage[b], c->stage_words[b], chunk * rw)) return -6;
        if (split && grow(c, c->d_leaf[b], c->leaf_words[b], leaf_digest_words(P, chunk))) return -6;
    }
    cudaStream_t cs = c->copy_stream;
    cudaStream_t ks[SV_NKS] = {c->own_stream};
    for (int i = 1; i < SV_NKS; i++) ks[i] = c->aux_stream[i - 1];
    const std::vector<size_t> sched = chunk_schedule(n_proofs, chunk, true);
    const size_t n_chunks = sched.size() - 1;
    // Device-side transcript of a host batch: the ~85 dependent permutations per proof are latency-bound
    // (one warp per 32 proofs), so the transcript runs ONCE for the whole batch, on the headers alone
    // (they travel first: 7 % of the bytes), beside the H2D of the full records; every chunk then gets
    // its challenge fields patched in with one strided device-to-device copy.
    const size_t hw = P.L.header_words, chal_off = P.L.off_alpha, chal_words = hw - chal_off;
    size_t fs_lead = 0;
    if (fs) {
        if (grow(c, c->d_hdr, c->hdr_words, n_proofs * hw)) return -6;
        CK(c, cudaMemcpy2DAsync(c->d_hdr, hw * 8, records, rw * 8, hw * 8, n_proofs, cudaMemcpyHostToDevice, cs));
        CK(c, cudaMemcpyAsync(c->d_pi, pi_hashes, n_proofs * 32, cudaMemcpyHostToDevice, cs));
        CK(c, cudaEventRecord(c->ev_hdr, cs));
        CK(c, cudaStreamWaitEvent(c->fs_stream, c->ev_hdr, 0));
        FriKernelParams Ph = P;
        Ph.L.record_words = (u32)hw;          // the headers are packed back to back
        // two parts, as in the wire path: the proofs of the first chunks (fs_lead_proofs) on the low-latency kernel, the rest with one
        // thread per proof beside them (their query kernels are not due before the first chunks are through)
        fs_lead = fs_lead_proofs(n_proofs, chunk);
        if (fs_lead < n_proofs) {
            CK(c, cudaStreamWaitEvent(c->fs_part_stream[0], c->ev_hdr, 0));
            FriKernelParams P1 = Ph;
            if ((rc = enqueue_challenges(c, P1, *fs, n_proofs - fs_lead, c->d_hdr + fs_lead * hw, c->d_pi + 4 * fs_lead, c->fs_part_stream[0], 2))) return rc;
            CK(c, cudaEventRecord(c->ev_part[1], c->fs_part_stream[0]));
        }
        if ((rc = enqueue_challenges(c, Ph, *fs, fs_lead, c->d_hdr, c->d_pi, c->fs_stream, fs_lead < n_proofs ? 1 : 0))) return rc;
        CK(c, cudaEventRecord(c->ev_part[0], c->fs_stream));
    }
    for (size_t i = 0; i < n_chunks; i++) {
        int b = (int)(i % SV_NBUF);
        cudaStream_t k = ks[i % n_ks];
        const size_t first = sched[i], cnt = sched[i + 1] - first;
        if (i >= SV_NBUF) CK(c, cudaStreamWaitEvent(cs, c->ev_done[b], 0));   // buffer b free again
        CK(c, cudaMemcpyAsync(c->d_stage[b], records + first * rw, cnt * rw * 8, cudaMemcpyHostToDevice, cs));
        CK(c, cudaEventRecord(c->ev_copied[b], cs));
        // the short work in front of the query kernel (challenge patch, fri_prepare_kernel) goes to the high-priority stream
        cudaStream_t pre = prep_stream_enabled() && !split ? c->prep_stream : k;
        CK(c, cudaStreamWaitEvent(pre, c->ev_copied[b], 0));
        if (pre != k) CK(c, cudaStreamWaitEvent(k, c->ev_copied[b], 0));
        if (fs) {
            if (split && (rc = enqueue_leaf(c, P, cnt, c->d_stage[b], c->d_leaf[b], k))) return rc;   // needs no challenge
            if (first < fs_lead) CK(c, cudaStreamWaitEvent(pre, c->ev_part[0], 0));             // a ramp-down chunk may straddle
            if (first + cnt > fs_lead) CK(c, cudaStreamWaitEvent(pre, c->ev_part[1], 0));       // the two transcript parts
            CK(c, cudaMemcpy2DAsync(c->d_stage[b] + chal_off, rw * 8, c->d_hdr + first * hw + chal_off, hw * 8, chal_words * 8, cnt,
                                    cudaMemcpyDeviceToDevice, pre));
        }
        rc = enqueue_fri(c, P, cnt, c->d_stage[b], c->d_scratch + 4 * first, c->d_bitmap + first / 32,
                         first_fail ? c->d_fail + first : nullptr, k, split ? c->d_leaf[b] : nullptr, pre != k ? pre : nullptr, c->ev_prep[b]);
        if (rc) return rc;
        CK(c, cudaEventRecord(c->ev_done[b], k));
    }
    // join the second compute stream, then read the results back on the first
    for (int j = 1; j < n_ks && (size_t)j < n_chunks; j++) {
        CK(c, cudaEventRecord(c->ev_join[j], ks[j]));
        CK(c, cudaStreamWaitEvent(ks[0], c->ev_join[j], 0));
    }
    CK(c, cudaMemcpyAsync(accept_bitmap, c->d_bitmap, n_words * 4, cudaMemcpyDeviceToHost, ks[0]));
    if (first_fail) CK(c, cudaMemcpyAsync(first_fail, c->d_fail, n_proofs * 4, cudaMemcpyDeviceToHost, ks[0]));
    CK(c, cudaStreamSynchronize(ks[0]));
    return 0;
}

// fs == nullptr: the records carry their challenges.  Otherwise the transcript runs on the device first
// (pi_hashes: n_proofs x 4 words, same memory space as the records).
static int fri_verify_impl(sv_ctx* c, const sv_fri_shape* shape, size_t n_proofs, const uint64_t* records,
                           uint32_t* accept_bitmap, uint32_t* first_fail, int mem, const FsParams* fs, const uint64_t* pi_hashes) {
    if (!c || !shape || !records || !accept_bitmap) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    FriKernelParams P;
    int rc = make_params(c, *shape, P);
    if (rc) return rc;
    if (n_proofs == 0) return 0;
    if (n_proofs * (size_t)P.num_queries >= (1ull << 31)) return fail(c, -8, "batch too large for one call");
    CK(c, cudaSetDevice(c->device));
    if (grow(c, c->d_scratch, c->scratch_words, 4 * n_proofs)) return -6;

    if (mem == SV_MEM_DEVICE) {
        // SV_MEM_DEVICE calls return without synchronising and all share d_scratch (reduced openings): when the caller moved to
        // another stream since the last such call (sv_ctx_set_stream), this call is ordered behind it
        if (c->scratch_busy && c->scratch_stream != c->stream) CK(c, cudaStreamWaitEvent(c->stream, c->ev_scratch, 0));
        const size_t lead = fs && n_proofs >= 4096 ? fs_lead_proofs(n_proofs, 32) : 0;
        if (fs && lead) {
            // The transcript (~155 dependent permutations per proof) in two parts on the high-priority side streams, both started
            // now: the first ~1 664 proofs on the lane-cooperative kernel (lowest latency, ~1.4 ms), the rest with one thread per
            // proof (3.8 ms, almost no GPU time); the query kernel of the first part hides the rest of the second transcript.
            u64* recs = const_cast<u64*>(records);
            const size_t rw = P.L.record_words;
            CK(c, cudaEventRecord(c->ev_hdr_ready, c->stream));
            CK(c, cudaStreamWaitEvent(c->fs_stream, c->ev_hdr_ready, 0));
            CK(c, cudaStreamWaitEvent(c->fs_part_stream[0], c->ev_hdr_ready, 0));
            FriKernelParams P0 = P, P1 = P;
            if ((rc = enqueue_challenges(c, P0, *fs, lead, recs, pi_hashes, c->fs_stream, 1))) return rc;
            CK(c, cudaEventRecord(c->ev_part[0], c->fs_stream));
            if ((rc = enqueue_challenges(c, P1, *fs, n_proofs - lead, recs + lead * rw, pi_hashes + 4 * lead, c->fs_part_stream[0], 2))) return rc;
            CK(c, cudaEventRecord(c->ev_part[1], c->fs_part_stream[0]));
            // the leaf sponges need no challenge: they fill the GPU while the transcripts run (the transcript kernels only
            // WRITE challenge fields of the header, the leaf kernel only reads query rounds)
            const bool split = leaf_split_enabled(P);
            const u64* lf = nullptr;
            if (split) {
                if (grow(c, c->d_leaf_dev, c->leaf_dev_words, leaf_digest_words(P, n_proofs))) return -6;
                if ((rc = enqueue_leaf(c, P, n_proofs, records, c->d_leaf_dev, c->stream))) return rc;
                lf = c->d_leaf_dev;
            }
            CK(c, cudaStreamWaitEvent(c->stream, c->ev_part[0], 0));
            if ((rc = enqueue_fri(c, P, lead, records, c->d_scratch, accept_bitmap, first_fail, c->stream, lf))) return rc;
            CK(c, cudaStreamWaitEvent(c->stream, c->ev_part[1], 0));
            rc = enqueue_fri(c, P, n_proofs - lead, records + lead * rw, c->d_scratch + 4 * lead, accept_bitmap + lead / 32,
                             first_fail ? first_fail + lead : nullptr, c->stream, lf ? lf + leaf_digest_words(P, lead) : nullptr);
        } else {
            if (fs && (rc = enqueue_challenges(c, P, *fs, n_proofs, const_cast<u64*>(records), pi_hashes, c->stream))) return rc;
            rc = enqueue_fri(c, P, n_proofs, records, c->d_scratch, accept_bitmap, first_fail, c->stream);
        }
        if (!c->ev_scratch) CK(c, cudaEventCreateWithFlags(&c->ev_scratch, cudaEventDisableTiming));
        CK(c, cudaEventRecord(c->ev_scratch, c->stream));
        c->scratch_stream = c->stream;
        c->scratch_busy = true;
        return rc;
    }
    // a host-mode call uses the library's own streams: wait for a device-mode call that may still be reading d_scratch
    if (c->scratch_busy) {
        CK(c, cudaEventSynchronize(c->ev_scratch));
        c->scratch_busy = false;
    }
    // host buffers: on an error half-way, drain what was already enqueued before the caller may free its buffers
    rc = fri_verify_host(c, P, n_proofs, records, accept_bitmap, first_fail, fs, pi_hashes);
    if (rc) {
        std::string keep = c->err;
        sv_ctx_synchronize(c);
        c->err = keep;
    }
    return rc;
}

extern "C" int sv_fri_verify_batch(sv_ctx* c, const sv_fri_shape* shape, size_t n_proofs, const uint64_t* records,
                                   uint32_t* accept_bitmap, uint32_t* first_fail, int mem) {
    return fri_verify_impl(c, shape, n_proofs, records, accept_bitmap, first_fail, mem, nullptr, nullptr);
}

static int make_fs(sv_ctx* c, const sv_fri_shape* shape, const uint64_t circuit_digest[4], uint32_t num_challenges, FsParams& F) {
    if (!shape || !circuit_digest) return -1;
    if (num_challenges == 0 || num_challenges > 16) return fail(c, -8, "num_challenges out of range");
    for (int i = 0; i < 4; i++) {
        if (!is_canonical(circuit_digest[i])) return fail(c, -8, "circuit_digest word >= p");
        F.circuit_digest[i] = circuit_digest[i];
    }
    if (shape->degree_bits > 32) return fail(c, -8, "degree_bits > 32 (Goldilocks has 2-adicity 32)");   // before the shift below
    F.g = svb::pow(7, (GL_P - 1) >> shape->degree_bits);
    F.num_challenges = num_challenges;
    return 0;
}

extern "C" int sv_fri_verify_batch_fs(sv_ctx* c, const sv_fri_shape* shape, size_t n_proofs, uint64_t* records,
                                      const uint64_t circuit_digest[4], const uint64_t* public_inputs_hashes,
                                      uint32_t num_challenges, uint32_t* accept_bitmap, uint32_t* first_fail, int mem) {
    if (!c || !public_inputs_hashes) return -1;
    FsParams F;
    int rc = make_fs(c, shape, circuit_digest, num_challenges, F);
    if (rc) return rc;
    return fri_verify_impl(c, shape, n_proofs, records, accept_bitmap, first_fail, mem, &F, public_inputs_hashes);
}

extern "C" int sv_fri_challenges_batch(sv_ctx* c, const sv_fri_shape* shape, size_t n_proofs, uint64_t* records,
                                       const uint64_t circuit_digest[4], const uint64_t* public_inputs_hashes,
                                       uint32_t num_challenges, int mem) {
    if (!c || !records || !public_inputs_hashes) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    FsParams F;
    int rc = make_fs(c, shape, circuit_digest, num_challenges, F);
    if (rc) return rc;
    FriKernelParams P;
    if ((rc = make_params(c, *shape, P))) return rc;
    if (n_proofs == 0) return 0;
    CK(c, cudaSetDevice(c->device));
    if (mem == SV_MEM_DEVICE) return enqueue_challenges(c, P, F, n_proofs, records, public_inputs_hashes, c->stream);
    // host records: only the header (caps, openings, final poly, pow witness) travels, in both directions
    size_t rw = P.L.record_words, hw = P.L.header_words;
    if (grow(c, c->d_stage[0], c->stage_words[0], n_proofs * rw)) return -6;
    if (grow(c, c->d_pi, c->pi_words, 4 * n_proofs)) return -6;
    cudaStream_t s = c->own_stream;
    CK(c, cudaMemcpy2DAsync(c->d_stage[0], rw * 8, records, rw * 8, hw * 8, n_proofs, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(c->d_pi, public_inputs_hashes, n_proofs * 32, cudaMemcpyHostToDevice, s));
    if ((rc = enqueue_challenges(c, P, F, n_proofs, c->d_stage[0], c->d_pi, s))) return rc;
    CK(c, cudaMemcpy2DAsync(records, rw * 8, c->d_stage[0], rw * 8, hw * 8, n_proofs, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
}

// The plonk circuit description on the device (validated by the caller).
static int upload_circuit(sv_ctx* c, const sv_plonk_circuit* circuit, cudaStream_t s) {
    if (!c->d_circuit) CK(c, cudaMalloc((void**)&c->d_circuit, sizeof(sv_plonk_circuit)));
    if (!c->circuit_valid || memcmp(&c->h_circuit, circuit, sizeof *circuit)) {
        if (int rc = sv_ctx_synchronize(c)) return rc;       // an earlier call may still read the old description
        c->h_circuit = *circuit;
        CK(c, cudaMemcpyAsync(c->d_circuit, &c->h_circuit, sizeof(sv_plonk_circuit), cudaMemcpyHostToDevice, s));
        CK(c, cudaStreamSynchronize(s));
        c->circuit_valid = true;
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Wire format: serialised proofs -> records on the device (wire_kernels.cuh).
struct WireDev {
    WireDims d;
    const u32 *hdr_src, *q_src, *chk;
    const u64* vk;
};
// Build (or reuse) the offset tables of (shape, common) and put them and the verifier key's cap on the device;
// the uploads are enqueued on `s`.
static int wire_tables(sv_ctx* c, const sv_fri_shape& shape, const sv_plonk_common& common, const uint64_t* vk_cap, cudaStream_t s,
                       WireDev& W) {
    bool same = c->w_valid && !memcmp(&c->w_shape, &shape, sizeof shape) && !memcmp(&c->w_common, &common, sizeof common);
    if (!same) {
        // kernels of an earlier call may still be reading the old tables
        if (int rc = sv_ctx_synchronize(c)) return rc;
        c->w_valid = false;
        int rc = make_wire_map(shape, common, c->w_map);
        if (rc) return fail(c, -8, "wire format: shape and common data disagree or the proof is too large (%d)", rc);
        c->w_shape = shape;
        c->w_common = common;
        c->w_vk.clear();
    }
    const WireMap& M = c->w_map;
    const size_t nh = M.hdr_src.size(), nq = M.q_src.size(), nc = M.chk.size(), ncap_words = 4ull << shape.cap_height;
    if (grow(c, c->d_wtab, c->wtab_words, nh + nq + nc)) return -6;
    if (grow(c, c->d_wvk, c->wvk_words, ncap_words)) return -6;
    if (!same) {
        CK(c, cudaMemcpyAsync(c->d_wtab, M.hdr_src.data(), nh * 4, cudaMemcpyHostToDevice, s));
        CK(c, cudaMemcpyAsync(c->d_wtab + nh, M.q_src.data(), nq * 4, cudaMemcpyHostToDevice, s));
        CK(c, cudaMemcpyAsync(c->d_wtab + nh + nq, M.chk.data(), nc * 4, cudaMemcpyHostToDevice, s));
        CK(c, cudaStreamSynchronize(s));   // rare (new circuit); later calls may use the tables from other streams
        c->w_valid = true;
    }
    if (c->w_vk.size() != ncap_words || memcmp(c->w_vk.data(), vk_cap, ncap_words * 8)) {
        if (same) {
            if (int rc = sv_ctx_synchronize(c)) return rc;   // same reason: the old cap may still be in use
        }
        c->w_vk.assign(vk_cap, vk_cap + ncap_words);
        CK(c, cudaMemcpyAsync(c->d_wvk, c->w_vk.data(), ncap_words * 8, cudaMemcpyHostToDevice, s));
        CK(c, cudaStreamSynchronize(s));
    }
    W.d = M.d;
    W.hdr_src = c->d_wtab;
    W.q_src = c->d_wtab + nh;
    W.chk = c->d_wtab + nh + nq;
    W.vk = c->d_wvk;
    return 0;
}

// unpack + public-input hashes of n proofs whose bytes start first_off bytes after d_blob8 (8-byte aligned)
static int enqueue_unpack(sv_ctx* c, const WireDev& W, const u64* d_blob8, size_t first_off, size_t stride, size_t n, u64* d_records,
                          u64* d_pi, u32* d_mal, cudaStream_t s) {
    CK(c, cudaMemsetAsync(d_mal, 0, n * 4, s));
    dim3 grid((unsigned)n, 1 + W.d.num_queries);   // per proof: one block for the header, one per query round
    wire_unpack_kernel<<<grid, SVB_WIRE_BLOCK, 0, s>>>(d_blob8, first_off, stride, W.d, W.hdr_src, W.q_src, W.chk, W.vk, d_records, d_mal, nullptr, 0);
    c->launches++;
    if (d_pi) {
        wire_pi_hash_kernel<<<(unsigned)((n + SVB_BLOCK - 1) / SVB_BLOCK), SVB_BLOCK, 0, s>>>(d_blob8, first_off, stride, W.d, n, d_pi, d_mal);
        c->launches++;
    }
    CK(c, cudaGetLastError());
    return 0;
}

extern "C" int sv_wire_unpack_batch_gpu(sv_ctx* c, const sv_fri_shape* shape, const sv_plonk_common* common,
                                        const uint64_t* vk_cap, const uint8_t* blob, size_t stride, size_t n, uint64_t* records_out,
                                        uint64_t* pi_hashes_out, uint32_t* malformed_out, int mem) {
    if (!c || !shape || !common || !vk_cap || !records_out || (n && !blob)) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (n >= (1ull << 31)) return fail(c, -8, "batch too large for one call");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = mem == SV_MEM_DEVICE ? c->stream : c->own_stream;
    WireDev W;
    if (int rc = wire_tables(c, *shape, *common, vk_cap, s, W)) return rc;
    if (n == 0) return 0;
    if (n > 1 && stride < W.d.proof_bytes) return fail(c, -8, "stride %zu < proof bytes %u", stride, W.d.proof_bytes);
    if (mem == SV_MEM_DEVICE) {
        if (reinterpret_cast<uintptr_t>(blob) & 7) return fail(c, -8, "device blob must be 8-byte aligned");
        u32* d_mal = malformed_out;
        if (!d_mal) {
            if (grow(c, c->d_mal, c->mal_words, n)) return -6;
            d_mal = c->d_mal;
        }
        return enqueue_unpack(c, W, reinterpret_cast<const u64*>(blob), 0, stride, n, records_out, pi_hashes_out, d_mal, s);
    }
    const size_t bytes = (n - 1) * stride + W.d.proof_bytes, rw = W.d.record_words;
    if (grow(c, c->d_wire[0], c->wire_words[0], bytes / 8 + 2)) return -6;
    if (grow(c, c->d_stage[0], c->stage_words[0], n * rw)) return -6;
    if (grow(c, c->d_pi, c->pi_words, 4 * n)) return -6;
    if (grow(c, c->d_mal, c->mal_words, n)) return -6;
    CK(c, cudaMemcpyAsync(c->d_wire[0], blob, bytes, cudaMemcpyHostToDevice, s));
    if (int rc = enqueue_unpack(c, W, c->d_wire[0], 0, stride, n, c->d_stage[0], c->d_pi, c->d_mal, s)) return rc;
    CK(c, cudaMemcpyAsync(records_out, c->d_stage[0], n * rw * 8, cudaMemcpyDeviceToHost, s));
    if (pi_hashes_out) CK(c, cudaMemcpyAsync(pi_hashes_out, c->d_pi, n * 32, cudaMemcpyDeviceToHost, s));
    if (malformed_out) CK(c, cudaMemcpyAsync(malformed_out, c->d_mal, n * 4, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
}

// The pipeline of fri_verify_host with wire bytes as its input, headers first:
//   copy stream   the two header spans of EVERY proof (strided copies), then chunk after chunk of query-round bytes
//   fs stream     once per batch, beside those copies: gather the packed headers, hash the public inputs, (plonk
//                 challenges,) the Fiat-Shamir transcript -- ~85 dependent permutations per proof, latency-bound, so it
//                 must not run per chunk -- (and the vanishing-polynomial identity, which only reads headers)
//   compute       per chunk, rotating over SV_NKS streams: gather the query rounds + copy in the finished header ->
//                 prepare + query kernels -> (AND with the identity) -> reject malformed
static inline size_t up8(size_t x) { return (x + 7) & ~(size_t)7; }
static int wire_verify_host(sv_ctx* c, FriKernelParams& P, const FsParams& F, const sv_fri_shape& shape, const sv_plonk_common& common,
                            const uint64_t* vk_cap, const uint8_t* blob, size_t stride, size_t n_proofs, uint32_t* accept_bitmap,
                            uint32_t* first_fail, const sv_plonk_circuit* circuit) {
    int rc = 0;
    const size_t rw = P.L.record_words, hw = P.L.header_words;
    cudaStream_t cs = c->copy_stream, fss = c->fs_stream;
    WireDev W;
    if ((rc = wire_tables(c, shape, common, vk_cap, cs, W))) return rc;
    const u32 nch = common.num_challenges;
    if (circuit) {
        if ((rc = upload_circuit(c, circuit, cs))) return rc;
        if (grow(c, c->d_chal, c->chal_words, 3 * (size_t)nch * n_proofs)) return -6;
        if (grow(c, c->d_pbm, c->pbm_words, (n_proofs + 31) / 32)) return -6;
    }
    if (n_proofs > 1 && stride < W.d.proof_bytes) return fail(c, -8, "stride %zu < proof bytes %u", stride, W.d.proof_bytes);
    size_t n_words = (n_proofs + 31) / 32;
    if (grow(c, c->d_bitmap, c->bitmap_words, n_words)) return -6;
    if (first_fail && grow(c, c->d_fail, c->fail_words, n_proofs)) return -6;
    if (grow(c, c->d_pi, c->pi_words, 4 * n_proofs)) return -6;
    if (grow(c, c->d_mal, c->mal_words, n_proofs)) return -6;
    if (grow(c, c->d_hdr, c->hdr_words, n_proofs * hw)) return -6;
    // the three byte spans of a proof and the pitches of their packed device copies (rows start 8-byte aligned)
    const size_t front_bytes = W.d.query_base, q_bytes = (size_t)W.d.num_queries * W.d.query_bytes;
    const size_t back_off = front_bytes + q_bytes, back_bytes = W.d.proof_bytes - back_off;
    const size_t front_pitch = up8(front_bytes) + 8, back_pitch = up8(back_bytes) + 8, q_pitch = up8(q_bytes) + 8;
    if (grow(c, c->d_hfront, c->hfront_words, n_proofs * front_pitch / 8 + 2)) return -6;
    if (grow(c, c->d_hback, c->hback_words, n_proofs * back_pitch / 8 + 2)) return -6;
    // 64 MiB chunks here (32 on the record path): the query kernels cannot start before the batch's transcript is done
    // (~3 ms), and until then the copy engine must find free buffers -- 6 x 64 MiB hold 7 ms of PCIe traffic (measured with
    // SVB_TRACE: 32 MiB chunks stall the copies for 1.5 ms; tools/lab/wire_trace.sh)
    size_t chunk_mb = 64;
    int n_ks = SV_NKS;
    if (const char* e = getenv("SVB_CHUNK_MB")) { long v = atol(e); if (v >= 1 && v <= 4096) chunk_mb = (size_t)v; }
    if (const char* e = getenv("SVB_KSTREAMS")) { long v = atol(e); if (v >= 1 && v <= SV_NKS) n_ks = (int)v; }
    size_t chunk = ((chunk_mb << 20) / (rw * 8)) & ~(size_t)31;
    if (chunk < 32) chunk = 32;
    if (chunk > n_proofs) chunk = (n_proofs + 31) & ~(size_t)31;
    const bool split = leaf_split_enabled(P);
    for (int b = 0; b < SV_NBUF; b++) {
        if (grow(c, c->d_stage[b], c->stage_words[b], chunk * rw)) return -6;
        if (grow(c, c->d_wire[b], c->wire_words[b], chunk * q_pitch / 8 + 2)) return -6;
        if (split && grow(c, c->d_leaf[b], c->leaf_words[b], leaf_digest_words(P, chunk))) return -6;
    }
    // SVB_TRACE=1: a timeline of this call on stderr (lab knob)
    static const bool trace = getenv("SVB_TRACE") != nullptr;
    cudaEvent_t tv[8] = {};
    std::vector<cudaEvent_t> tc;                            // per chunk: copied, kernels start, query kernel start, done
    if (trace) {
        for (auto& e : tv) cudaEventCreate(&e);
        cudaEventRecord(tv[0], cs);
    }
    // ---- headers first, transcript part by transcript part ---------------------------------------------------
    // A transcript is ~155 DEPENDENT permutations per proof.  The lane-cooperative kernel has the lower latency (6.9 us per
    // permutation while every warp has an SM sub-partition to itself: ~1.1 ms for up to ~1 200 proofs, 2.8 ms for 4 096), the
    // thread-per-proof kernel needs ~3.8 ms and almost no GPU time.  So the batch is cut into parts, each on its own high-priority
    // stream: the header spans of the LEAD chunks cross PCIe first (0.2 ms instead of 0.8 for the whole batch) and their
    // transcript starts at once -- the query kernels of those chunks then keep the GPU busy while the rest of the headers
    // arrive and the other parts finish.  (One transcript for the whole batch left the GPU idle for 3.3 of 14.4 ms:
    // tools/lab/wire_trace.sh, SVB_TRACE.)
    static const int parts_env = [] { const char* e = getenv("SVB_FS_PARTS"); return e ? atoi(e) : 2; }();
    static const int lead_chunks_env = [] { const char* e = getenv("SVB_FS_LEAD"); return e ? atoi(e) : 0; }();   // chunks in the first part (lab knob)
    const size_t lead_env = lead_chunks_env > 0 ? (size_t)lead_chunks_env : (fs_lead_proofs(n_proofs, chunk) + chunk - 1) / chunk;
    static const int mid_env = [] { const char* e = getenv("SVB_FS_MID"); return e ? atoi(e) : 2; }();     // chunks in the second of three parts
    static const int rest_coop_env = [] { const char* e = getenv("SVB_FS_REST_COOP"); return e ? atoi(e) : 0; }();   // last part: cooperative too
    const size_t part_end[3] = {std::min(n_proofs, (size_t)lead_env * chunk), std::min(n_proofs, (size_t)(lead_env + mid_env) * chunk), n_proofs};
    cudaStream_t pst[3] = {fss, c->fs_part_stream[0], c->fs_part_stream[1]};
    size_t part_lo[3] = {0, 0, 0}, part_hi[3] = {0, 0, 0};   // proofs [lo, hi) of each transcript part (hi == lo: part unused)
    CK(c, cudaMemsetAsync(c->d_mal, 0, n_proofs * 4, cs));
    {
        size_t lo = 0;
        for (int pi = 0; pi < 3; pi++) {
            const size_t hi_ = parts_env >= 3 ? part_end[pi] : (parts_env == 2 ? (pi == 0 ? part_end[0] : n_proofs) : n_proofs);
            if (hi_ <= lo) continue;
            const size_t cnt = hi_ - lo;
            cudaStream_t ps = pst[pi];
            // the two header spans of this part's proofs: packed rows on the device
            CK(c, cudaMemcpy2DAsync((uint8_t*)c->d_hfront + lo * front_pitch, front_pitch, blob + lo * stride, stride, front_bytes, cnt,
                                    cudaMemcpyHostToDevice, cs));
            CK(c, cudaMemcpy2DAsync((uint8_t*)c->d_hback + lo * back_pitch, back_pitch, blob + lo * stride + back_off, stride, back_bytes, cnt,
                                    cudaMemcpyHostToDevice, cs));
            CK(c, cudaEventRecord(c->ev_hdr_part[pi], cs));
            if (trace && hi_ == n_proofs) cudaEventRecord(tv[1], cs);
            CK(c, cudaStreamWaitEvent(ps, c->ev_hdr_part[pi], 0));
            wire_header_unpack_kernel<<<(unsigned)cnt, SVB_WIRE_BLOCK, 0, ps>>>(c->d_hfront + lo * (front_pitch / 8), front_pitch,
                                                                                c->d_hback + lo * (back_pitch / 8), back_pitch, (u32)front_bytes,
                                                                                (u32)back_off, (u32)hw, W.hdr_src, W.vk, c->d_hdr + lo * hw);
            {
                WireDims db = W.d;
                db.pi_off = (u32)(W.d.pi_off - back_off);          // the public inputs inside the back span
                wire_pi_hash_kernel<<<(unsigned)((cnt + SVB_BLOCK - 1) / SVB_BLOCK), SVB_BLOCK, 0, ps>>>(c->d_hback + lo * (back_pitch / 8), 0, back_pitch, db,
                                                                                                       cnt, c->d_pi + 4 * lo, c->d_mal + lo);
            }
            c->launches += 2;
            if (trace && pi == 0) cudaEventRecord(tv[6], ps);
            FriKernelParams Ph = P;
            Ph.L.record_words = (u32)hw;                    // the headers are packed back to back
            if (circuit) {   // plonk challenges: the prefix of the transcript below, kept this time
                Ph.n_proofs = (u32)cnt;
                SVB_LAUNCH_KIND(P.hash_kind, plonk_challenges_kernel, (unsigned)((cnt + SVB_FS_BLOCK - 1) / SVB_FS_BLOCK), SVB_FS_BLOCK, ps,
                                c->d_hdr + lo * hw, Ph, F, c->d_pi + 4 * lo, c->d_chal + 3 * (size_t)nch * lo);
                c->launches++;
            }
            // the big last part: thread per proof (it ends long before the chunks in front of it are through their query kernels)
            const int force = hi_ == n_proofs && cnt > 2048 && lo > 0 && !rest_coop_env ? 2 : 0;
            if ((rc = enqueue_challenges(c, Ph, F, cnt, c->d_hdr + lo * hw, c->d_pi + 4 * lo, ps, force))) return rc;
            if (trace && pi == 0) cudaEventRecord(tv[7], ps);
            CK(c, cudaEventRecord(c->ev_part[pi], ps));       // the chunks' query kernels wait for the challenges only
            if (circuit) {   // the vanishing-polynomial identity reads headers only (zeta is in them now); its verdict is not needed
                             // before the AND at the end of a chunk, so it runs beside the first query kernels (ev_plonk)
                PlonkRecordView V = {(u32)hw, P.L.off_open0, P.L.off_open1, P.L.off_zeta};
                plonk_check_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, ps>>>(c->d_hdr + lo * hw, V, c->d_circuit, c->d_pi + 4 * lo,
                                                                                 c->d_chal + 3 * (size_t)nch * lo, (u32)cnt, c->d_pbm + lo / 32);
                c->launches++;
                CK(c, cudaEventRecord(c->ev_plonk[pi], ps));
            }
            CK(c, cudaGetLastError());
            part_lo[pi] = lo; part_hi[pi] = hi_;
            lo = hi_;
        }
    }
    if (trace) cudaEventRecord(tv[2], c->fs_part_stream[1]);
    // ---- query rounds, chunk by chunk ------------------------------------------------------------------------
    WireDims dq = W.d;
    dq.query_base = 0;                                      // a chunk buffer row holds the query rounds only
    cudaStream_t ks[SV_NKS] = {c->own_stream};
    for (int i = 1; i < SV_NKS; i++) ks[i] = c->aux_stream[i - 1];
    const std::vector<size_t> sched = chunk_schedule(n_proofs, chunk, false);
    const size_t n_chunks = sched.size() - 1;
    for (size_t i = 0; i < n_chunks; i++) {
        int b = (int)(i % SV_NBUF);
        cudaStream_t k = ks[i % n_ks];
        const size_t first = sched[i], cnt = sched[i + 1] - first;
        if (i >= SV_NBUF) CK(c, cudaStreamWaitEvent(cs, c->ev_done[b], 0));   // buffers b free again
        CK(c, cudaMemcpy2DAsync(c->d_wire[b], q_pitch, blob + first * stride + front_bytes, stride, q_bytes, cnt, cudaMemcpyHostToDevice, cs));
        CK(c, cudaEventRecord(c->ev_copied[b], cs));
        // unpack + prepare are short; on the chunk's own stream they would queue behind the ~1 000 pending blocks of the previous
        // chunk's query kernel (the SMs are full) and hold this chunk's query kernel back by ~0.6 ms (SVB_TRACE): high-priority stream
        cudaStream_t pre = prep_stream_enabled() && !split ? c->prep_stream : k;
        CK(c, cudaStreamWaitEvent(pre, c->ev_copied[b], 0));
        if (pre != k) CK(c, cudaStreamWaitEvent(k, c->ev_copied[b], 0));
        if (trace) {
            for (int j = 0; j < 4; j++) { cudaEvent_t e; cudaEventCreate(&e); tc.push_back(e); }
            cudaEventRecord(tc[4 * i], cs);
        }
        // the transcript parts that cover this chunk (a ramp-down chunk may straddle two of them)
        auto wait_parts = [&](cudaEvent_t* ev, cudaStream_t st) -> int {
            for (int pi = 0; pi < 3; pi++)
                if (part_hi[pi] > part_lo[pi] && part_lo[pi] < first + cnt && first < part_hi[pi]) CK(c, cudaStreamWaitEvent(st, ev[pi], 0));
            return 0;
        };
        if (split) {
            // what needs no challenge goes first: the query rounds into the records, then their leaf digests -- this work fills
            // the GPU while the transcript of the chunk's part is still running; the finished header follows
            wire_unpack_kernel<<<dim3((unsigned)cnt, W.d.num_queries), SVB_WIRE_BLOCK, 0, k>>>(c->d_wire[b], 0, q_pitch, dq, W.hdr_src, W.q_src, W.chk,
                                                                                            W.vk, c->d_stage[b], c->d_mal + first, nullptr, 1);
            c->launches++;
            if ((rc = enqueue_leaf(c, P, cnt, c->d_stage[b], c->d_leaf[b], k))) return rc;
            if ((rc = wait_parts(c->ev_part, k))) return rc;
            wire_unpack_kernel<<<dim3((unsigned)cnt, 1), SVB_WIRE_BLOCK, 0, k>>>(c->d_wire[b], 0, q_pitch, dq, W.hdr_src, W.q_src, W.chk, W.vk,
                                                                              c->d_stage[b], c->d_mal + first, c->d_hdr + first * hw, 0);
            c->launches++;
        } else {
            if ((rc = wait_parts(c->ev_part, pre))) return rc;
            if (trace) cudaEventRecord(tc[4 * i + 1], pre);
            dim3 grid((unsigned)cnt, 1 + W.d.num_queries);
            wire_unpack_kernel<<<grid, SVB_WIRE_BLOCK, 0, pre>>>(c->d_wire[b], 0, q_pitch, dq, W.hdr_src, W.q_src, W.chk, W.vk, c->d_stage[b],
                                                                 c->d_mal + first, c->d_hdr + first * hw, 0);
            c->launches++;
        }
        u32* d_fail = first_fail ? c->d_fail + first : nullptr;
        if (trace && !split) cudaEventRecord(tc[4 * i + 2], k);
        if ((rc = enqueue_fri(c, P, cnt, c->d_stage[b], c->d_scratch + 4 * first, c->d_bitmap + first / 32, d_fail, k,
                              split ? c->d_leaf[b] : nullptr, pre != k ? pre : nullptr, c->ev_prep[b])))
            return rc;
        if (trace) cudaEventRecord(tc[4 * i + 3], k);
        if (circuit) {
            if ((rc = wait_parts(c->ev_plonk, k))) return rc;
            plonk_and_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, k>>>(c->d_pbm + first / 32, c->d_bitmap + first / 32, d_fail, (u32)cnt);
            c->launches++;
        }
        wire_reject_malformed_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, k>>>(c->d_mal + first, c->d_bitmap + first / 32, d_fail, (u32)cnt);
        c->launches++;
        CK(c, cudaGetLastError());
        CK(c, cudaEventRecord(c->ev_done[b], k));
        if (trace && i == 0) cudaEventRecord(tv[3], k);
    }
    if (trace) cudaEventRecord(tv[4], cs);
    for (int j = 1; j < n_ks && (size_t)j < n_chunks; j++) {
        CK(c, cudaEventRecord(c->ev_join[j], ks[j]));
        CK(c, cudaStreamWaitEvent(ks[0], c->ev_join[j], 0));
    }
    CK(c, cudaMemcpyAsync(accept_bitmap, c->d_bitmap, n_words * 4, cudaMemcpyDeviceToHost, ks[0]));
    if (first_fail) CK(c, cudaMemcpyAsync(first_fail, c->d_fail, n_proofs * 4, cudaMemcpyDeviceToHost, ks[0]));
    if (trace) cudaEventRecord(tv[5], ks[0]);
    CK(c, cudaStreamSynchronize(ks[0]));
    if (trace) {
        float t[8] = {};
        for (int i = 1; i < 8; i++) cudaEventElapsedTime(&t[i], tv[0], tv[i]);
        fprintf(stderr, "[svb trace] wire: headers copied %.2f ms, headers unpacked + pi hashed %.2f, challenges %.2f, transcript stream done %.2f, "
                        "first chunk done %.2f, last chunk copied %.2f, end %.2f (%zu proofs, %zu chunks of %zu)\n",
                t[1], t[6], t[7], t[2], t[3], t[4], t[5], n_proofs, n_chunks, chunk);
        if (!split) {
            fprintf(stderr, "[svb trace] chunks (copied / kernels start / query start / query done, ms):");
            for (size_t i = 0; i < n_chunks; i++) {
                float a = 0, b = 0, d = 0, e = 0;
                cudaEventElapsedTime(&a, tv[0], tc[4 * i]); cudaEventElapsedTime(&b, tv[0], tc[4 * i + 1]);
                cudaEventElapsedTime(&d, tv[0], tc[4 * i + 2]); cudaEventElapsedTime(&e, tv[0], tc[4 * i + 3]);
                fprintf(stderr, " [%zu] %.2f %.2f %.2f %.2f", i, a, b, d, e);
            }
            fprintf(stderr, "\n");
        }
        for (auto& e : tc) cudaEventDestroy(e);
        for (auto& e : tv) cudaEventDestroy(e);
    }
    return 0;
}

extern "C" int sv_verify_proofs_wire(sv_ctx* c, const sv_fri_shape* shape, const sv_plonk_common* common,
                                     const uint64_t* vk_cap, const uint64_t circuit_digest[4], const uint8_t* blob, size_t stride,
                                     size_t n_proofs, uint32_t* accept_bitmap, uint32_t* first_fail) {
    if (!c || !shape || !common || !vk_cap || !circuit_digest || !accept_bitmap || (n_proofs && !blob)) return -1;
    FsParams F;
    int rc = make_fs(c, shape, circuit_digest, common->num_challenges, F);
    if (rc) return rc;
    FriKernelParams P;
    if ((rc = make_params(c, *shape, P))) return rc;
    if (n_proofs == 0) return 0;
    if (n_proofs * (size_t)P.num_queries >= (1ull << 31)) return fail(c, -8, "batch too large for one call");
    CK(c, cudaSetDevice(c->device));
    if (grow(c, c->d_scratch, c->scratch_words, 4 * n_proofs)) return -6;
    if (c->scratch_busy) {   // a SV_MEM_DEVICE call on the caller's stream may still be reading d_scratch
        CK(c, cudaEventSynchronize(c->ev_scratch));
        c->scratch_busy = false;
    }
    rc = wire_verify_host(c, P, F, *shape, *common, vk_cap, blob, stride, n_proofs, accept_bitmap, first_fail, nullptr);
    if (rc) {
        std::string keep = c->err;
        sv_ctx_synchronize(c);
        c->err = keep;
    }
    return rc;
}

namespace svb { int plonk_shape_matches(const sv_fri_shape& s, const sv_plonk_circuit& C); }

extern "C" int sv_verify_proofs_full(sv_ctx* c, const sv_fri_shape* shape, const sv_plonk_circuit* circuit, const uint64_t* vk_cap,
                                     const uint64_t circuit_digest[4], const uint8_t* blob, size_t stride, size_t n_proofs,
                                     uint32_t* accept_bitmap, uint32_t* first_fail) {
    if (!c || !shape || !circuit || !vk_cap || !circuit_digest || !accept_bitmap || (n_proofs && !blob)) return -1;
    if (int rc = plonk_circuit_check(*circuit)) return fail(c, -8, "plonk circuit description refused (%d)", rc);
    if (int rc = plonk_shape_matches(*shape, *circuit)) return fail(c, -8, "FRI shape and plonk circuit disagree (%d)", rc);
    FsParams F;
    int rc = make_fs(c, shape, circuit_digest, circuit->common.num_challenges, F);
    if (rc) return rc;
    FriKernelParams P;
    if ((rc = make_params(c, *shape, P))) return rc;
    if (n_proofs == 0) return 0;
    if (n_proofs * (size_t)P.num_queries >= (1ull << 31)) return fail(c, -8, "batch too large for one call");
    CK(c, cudaSetDevice(c->device));
    if (grow(c, c->d_scratch, c->scratch_words, 4 * n_proofs)) return -6;
    if (c->scratch_busy) {   // a SV_MEM_DEVICE call on the caller's stream may still be reading d_scratch
        CK(c, cudaEventSynchronize(c->ev_scratch));
        c->scratch_busy = false;
    }
    rc = wire_verify_host(c, P, F, *shape, circuit->common, vk_cap, blob, stride, n_proofs, accept_bitmap, first_fail, circuit);
    if (rc) {
        std::string keep = c->err;
        sv_ctx_synchronize(c);
        c->err = keep;
    }
    return rc;
}

// ---------------------------------------------------------------------------------------------
// Plonk-level checks: plonk_check_kernel, one thread per proof.
namespace svb { int plonk_shape_matches(const sv_fri_shape& s, const sv_plonk_circuit& C); }

extern "C" int sv_plonk_check_batch(sv_ctx* c, const sv_fri_shape* shape, const sv_plonk_circuit* circuit, size_t n,
                                    const uint64_t* records, const uint64_t* pi_hashes, const uint64_t* chal,
                                    uint32_t* accept_bitmap, int mem) {
    if (!c || !shape || !circuit || !accept_bitmap || (n && (!records || !pi_hashes || !chal))) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    sv_fri_layout L;
    if (make_layout(*shape, L)) return fail(c, -8, "bad FRI shape");
    if (int rc = plonk_circuit_check(*circuit)) return fail(c, -8, "plonk circuit description refused (%d)", rc);
    if (int rc = plonk_shape_matches(*shape, *circuit)) return fail(c, -8, "FRI shape and plonk circuit disagree (%d)", rc);
    if (n == 0) return 0;
    if (n >= (1ull << 31)) return fail(c, -8, "batch too large for one call");
    CK(c, cudaSetDevice(c->device));
    cudaStream_t s = mem == SV_MEM_DEVICE ? c->stream : c->own_stream;
    if (int rc = upload_circuit(c, circuit, s)) return rc;
    PlonkRecordView V = {L.record_words, L.off_open0, L.off_open1, L.off_zeta};
    const u32 nch = circuit->common.num_challenges;
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (mem == SV_MEM_DEVICE) {
        plonk_check_kernel<<<grid, 128, 0, s>>>(records, V, c->d_circuit, pi_hashes, chal, (u32)n, accept_bitmap);
        c->launches++;
        CK(c, cudaGetLastError());
        return 0;
    }
    // host buffers: only the record headers travel
    const size_t hw = L.header_words, n_words = (n + 31) / 32;
    if (grow(c, c->d_hdr, c->hdr_words, n * hw)) return -6;
    if (grow(c, c->d_pi, c->pi_words, 4 * n)) return -6;
    if (grow(c, c->d_chal, c->chal_words, 3 * (size_t)nch * n)) return -6;
    if (grow(c, c->d_bitmap, c->bitmap_words, n_words)) return -6;
    CK(c, cudaMemcpy2DAsync(c->d_hdr, hw * 8, records, (size_t)L.record_words * 8, hw * 8, n, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(c->d_pi, pi_hashes, n * 32, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(c->d_chal, chal, 3 * (size_t)nch * n * 8, cudaMemcpyHostToDevice, s));
    V.record_words = (u32)hw;                                // the headers are packed back to back
    plonk_check_kernel<<<grid, 128, 0, s>>>(c->d_hdr, V, c->d_circuit, c->d_pi, c->d_chal, (u32)n, c->d_bitmap);
    c->launches++;
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(accept_bitmap, c->d_bitmap, n_words * 4, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// NTT / LDE: ntt_pass_kernel, one launch per pass of ntt_plan.
static int ntt_twiddles_dev(sv_ctx* c, u32 k, bool inverse, cudaStream_t s) {
    if (c->d_tw && c->tw_k == k && c->tw_inverse == (int)inverse) return 0;
    if (int rc = sv_ctx_synchronize(c)) return rc;           // an earlier transform may still read the old table
    const size_t words = std::max<size_t>(1, ((size_t)1 << k) / 2);
    if (grow(c, c->d_tw, c->tw_words, words)) return -6;
    std::vector<u64> tw(words);
    ntt_twiddles(k, inverse, tw.data());
    CK(c, cudaMemcpyAsync(c->d_tw, tw.data(), words * 8, cudaMemcpyHostToDevice, s));
    CK(c, cudaStreamSynchronize(s));                         // tw goes out of scope
    c->tw_k = k;
    c->tw_inverse = (int)inverse;
    return 0;
}
// the LDE scale tables of (log_n, rate_bits, shift), cached like the twiddles
static int lde_tables_dev(sv_ctx* c, u32 log_n, u32 rate_bits, u64 shift, cudaStream_t s) {
    if (c->d_lde_lo && c->lde_log_n == log_n && c->lde_rate_bits == rate_bits && c->lde_shift == shift) return 0;
    if (int rc = sv_ctx_synchronize(c)) return rc;
    const u32 h = lde_scale_h(log_n);
    const size_t lo_words = (size_t)1 << (rate_bits + h), hi_words = (size_t)1 << (rate_bits + log_n - h);
    if (grow(c, c->d_lde_lo, c->lde_lo_words, lo_words) || grow(c, c->d_lde_hi, c->lde_hi_words, hi_words)) return -6;
    std::vector<u64> lo(lo_words), hi(hi_words);
    lde_scale_tables(log_n, rate_bits, shift, lo.data(), hi.data());
    CK(c, cudaMemcpyAsync(c->d_lde_lo, lo.data(), lo_words * 8, cudaMemcpyHostToDevice, s));
    CK(c, cudaMemcpyAsync(c->d_lde_hi, hi.data(), hi_words * 8, cudaMemcpyHostToDevice, s));
    CK(c, cudaStreamSynchronize(s));
    c->lde_log_n = log_n; c->lde_rate_bits = rate_bits; c->lde_shift = shift;
    return 0;
}

static int launch_ntt_pass(sv_ctx* c, const NttPass& P, size_t items, const u64* src, size_t src_stride, u64* dst, size_t dst_stride,
                           cudaStream_t s) {
    static bool attr_set = false;
    const size_t smem = (size_t)(((size_t)1 << P.logT) + ((size_t)1 << P.logT >> 3) + 8) * 8;
    if (!attr_set) {
        CK(c, cudaFuncSetAttribute(ntt_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NTT_SMEM_WORDS * 8 + 64)));
        attr_set = true;
    }
    const u64 blocks = (u64)items << (P.k - P.logT);
    if (blocks >= (1ull << 31)) return fail(c, -8, "transform batch too large for one launch");
    ntt_pass_kernel<<<(unsigned)blocks, NTT_THREADS, smem, s>>>(src, src_stride, dst, dst_stride, c->d_tw, c->d_lde_lo, c->d_lde_hi, P,
                                                                 ntt_rounds(P));
    c->launches++;
    return 0;
}

static int enqueue_ntt(sv_ctx* c, u32 k, size_t n_polys, u64* d_data, bool inverse, cudaStream_t s) {
    NttPass plan[NTT_MAX_PASSES];
    int np = ntt_plan(k, inverse, plan);
    if (np < 0) return fail(c, -8, "bad transform size 2^%u", k);
    if (int rc = ntt_twiddles_dev(c, k, inverse, s)) return rc;
    const size_t n = (size_t)1 << k;
    for (int q = 0; q < np; q++)
        if (int rc = launch_ntt_pass(c, plan[q], n_polys, d_data, n, d_data, n, s)) return rc;
    CK(c, cudaGetLastError());
    return 0;
}
// coefficients (n_polys x 2^log_n, at d_in) -> values on shift * <omega_N> in leaf order (n_polys x N, at d_out): 2^rate_bits
// size-n transforms per polynomial, the first pass reading the coefficients and scaling them on load (ntt.hpp)
static int enqueue_lde(sv_ctx* c, u32 log_n, u32 rate_bits, size_t n_polys, const u64* d_in, u64 shift, u64* d_out, cudaStream_t s) {
    NttPass plan[NTT_MAX_PASSES];
    int np = ntt_plan(log_n, false, plan);
    if (np < 0) return fail(c, -8, "bad transform size 2^%u", log_n);
    if (int rc = ntt_twiddles_dev(c, log_n, false, s)) return rc;
    if (int rc = lde_tables_dev(c, log_n, rate_bits, shift, s)) return rc;
    const size_t n = (size_t)1 << log_n, items = n_polys << rate_bits;
    for (int q = 0; q < np; q++) {
        NttPass P = plan[q];
        if (q == 0) {
            P.coset_bits = rate_bits;
            P.scale_h = lde_scale_h(log_n);
            P.load_scaled = 1;
            if (int rc = launch_ntt_pass(c, P, items, d_in, n, d_out, n, s)) return rc;
        } else if (int rc = launch_ntt_pass(c, P, items, d_out, n, d_out, n, s))
            return rc;
    }
    CK(c, cudaGetLastError());
    return 0;
}

extern "C" int sv_ntt_batch(sv_ctx* c, uint32_t log_n, size_t n_polys, uint64_t* data, int inverse, int mem) {
    if (!c || (!data && n_polys)) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (log_n == 0 || log_n > 26) return fail(c, -8, "log_n out of range (1..26)");
    if (n_polys == 0) return 0;
    CK(c, cudaSetDevice(c->device));
    if (mem == SV_MEM_DEVICE) return enqueue_ntt(c, log_n, n_polys, data, inverse != 0, c->stream);
    const size_t words = n_polys << log_n;
    if (grow(c, c->d_stage[0], c->stage_words[0], words)) return -6;
    cudaStream_t s = c->own_stream;
    CK(c, cudaMemcpyAsync(c->d_stage[0], data, words * 8, cudaMemcpyHostToDevice, s));
    if (int rc = enqueue_ntt(c, log_n, n_polys, c->d_stage[0], inverse != 0, s)) return rc;
    CK(c, cudaMemcpyAsync(data, c->d_stage[0], words * 8, cudaMemcpyDeviceToHost, s));
    CK(c, cudaStreamSynchronize(s));
    return 0;
}

extern "C" int sv_lde_batch(sv_ctx* c, uint32_t log_n, uint32_t rate_bits, size_t n_polys, const uint64_t* coeffs, uint64_t shift,
                            uint64_t* out, int mem) {
    if (!c || ((!coeffs || !out) && n_polys)) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (log_n == 0 || log_n + rate_bits > 26 || !is_canonical(shift) || shift == 0) return fail(c, -8, "bad LDE parameters");
    if (n_polys == 0) return 0;
    CK(c, cudaSetDevice(c->device));
    const u32 log_N = log_n + rate_bits;
    const size_t in_words = n_polys << log_n, out_words = n_polys << log_N;
    cudaStream_t s = mem == SV_MEM_DEVICE ? c->stream : c->own_stream;
    const u64* d_in = coeffs;
    u64* d_out = out;
    if (mem != SV_MEM_DEVICE) {
        if (grow(c, c->d_stage[0], c->stage_words[0], in_words)) return -6;
        if (grow(c, c->d_stage[1], c->stage_words[1], out_words)) return -6;
        CK(c, cudaMemcpyAsync(c->d_stage[0], coeffs, in_words * 8, cudaMemcpyHostToDevice, s));
        d_in = c->d_stage[0];
        d_out = c->d_stage[1];
    }
    if (int rc = enqueue_lde(c, log_n, rate_bits, n_polys, d_in, shift, d_out, s)) return rc;
    if (mem != SV_MEM_DEVICE) {
        CK(c, cudaMemcpyAsync(out, d_out, out_words * 8, cudaMemcpyDeviceToHost, s));
        CK(c, cudaStreamSynchronize(s));
    }
    return 0;
}

extern "C" int sv_commit_batch(sv_ctx* c, uint32_t log_n, uint32_t rate_bits, size_t n_polys, const uint64_t* coeffs, uint32_t cap_height,
                               int hash_kind, uint64_t* leaves_out, uint64_t* layers_out, int mem) {
    if (!c || !coeffs || !layers_out || n_polys == 0) return -1;
    if (!known_mem(mem)) return fail(c, -5, "mem %d is neither SV_MEM_HOST nor SV_MEM_DEVICE", mem);
    if (!known_kind(hash_kind)) return fail(c, -5, "hash_kind %d not implemented", hash_kind);
    if (log_n == 0 || log_n + rate_bits > 26 || cap_height > log_n + rate_bits) return fail(c, -8, "bad commit parameters");
    CK(c, cudaSetDevice(c->device));
    const u32 log_N = log_n + rate_bits;
    const size_t N = (size_t)1 << log_N, ncap = (size_t)1 << cap_height;
    const size_t in_words = n_polys << log_n, lde_words = n_polys << log_N, layer_words = 4 * (2 * N - ncap);
    cudaStream_t s = mem == SV_MEM_DEVICE ? c->stream : c->own_stream;
    // device buffers: [2] = LDE (polynomial-major), leaves (point-major, only when asked for), layers
    if (grow(c, c->d_stage[2], c->stage_words[2], lde_words)) return -6;
    const u64* d_in = coeffs;
    u64 *d_leaves = leaves_out, *d_layers = layers_out;
    if (mem != SV_MEM_DEVICE) {
        if (grow(c, c->d_stage[0], c->stage_words[0], in_words)) return -6;
        if (leaves_out && grow(c, c->d_stage[1], c->stage_words[1], lde_words)) return -6;
        if (grow(c, c->d_stage[3], c->stage_words[3], layer_words)) return -6;
        CK(c, cudaMemcpyAsync(c->d_stage[0], coeffs, in_words * 8, cudaMemcpyHostToDevice, s));
        d_in = c->d_stage[0];
        d_leaves = leaves_out ? c->d_stage[1] : nullptr;
        d_layers = c->d_stage[3];
    }
    u64* d_lde = c->d_stage[2];
    if (int rc = enqueue_lde(c, log_n, rate_bits, n_polys, d_in, 7, d_lde, s)) return rc;
    const int B = SVB_BLOCK;
    // leaf digests straight from the polynomial-major values: thread i reads word i of every polynomial (coalesced)
    SVB_LAUNCH_KIND(hash_kind, merkle_leaf_hash_cols_kernel, (unsigned)((N + B - 1) / B), B, s, d_lde, (u32)n_polys, N, d_layers);
    c->launches++;
    if (d_leaves) {
        dim3 tg((unsigned)((N + 31) / 32), (unsigned)((n_polys + 31) / 32));
        transpose_kernel<<<tg, dim3(32, 8), 0, s>>>(d_lde, d_leaves, n_polys, N);
        c->launches++;
    }
    u64* cur = d_layers;
    for (size_t m = N; m > ncap; m >>= 1) {
        u64* nxt = cur + 4 * m;
        SVB_LAUNCH_KIND(hash_kind, merkle_level_kernel, (unsigned)((m / 2 + B - 1) / B), B, s, cur, nxt, m / 2);
        c->launches++;
        cur = nxt;
    }
    CK(c, cudaGetLastError());
    if (mem != SV_MEM_DEVICE) {
        if (leaves_out) CK(c, cudaMemcpyAsync(leaves_out, d_leaves, lde_words * 8, cudaMemcpyDeviceToHost, s));
        CK(c, cudaMemcpyAsync(layers_out, d_layers, layer_words * 8, cudaMemcpyDeviceToHost, s));
        CK(c, cudaStreamSynchronize(s));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
extern "C" int sv_allgather_bitmap(sv_ctx* c, void* nccl_comm, const uint32_t* local_words, uint32_t* all_words,
                                   size_t words_per_rank) {
    if (!c || !nccl_comm || !local_words || !all_words) return -1;
    if (!c->ncclAllGather_fn) {
        c->nccl_lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!c->nccl_lib) return fail(c, -9, "dlopen libnccl.so.2: %s", dlerror());
        c->ncclAllGather_fn = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(c->nccl_lib, "ncclAllGather");
        if (!c->ncclAllGather_fn) return fail(c, -9, "dlsym ncclAllGather failed");
    }
    CK(c, cudaSetDevice(c->device));
    const int ncclUint32 = 3;  // ncclDataType_t: ncclInt8=0, ncclUint8=1, ncclInt32=2, ncclUint32=3
    int r = c->ncclAllGather_fn(local_words, all_words, words_per_rank, ncclUint32, nccl_comm, c->stream);
    if (r != 0) return fail(c, -10, "ncclAllGather failed with %d", r);
    return 0;
}
