/* Minimal C caller of the drop-in boundary: serialised plonky2 proofs in, accept bits out.
 *
 *   gcc -std=c99 -I include examples/verify_wire.c -L stark-verifier_b200 -lsvb200 -Wl,-rpath,$PWD/stark-verifier_b200 -o verify_wire
 *   ./verify_wire proofs.bin vk.bin            # vk.bin = constants_sigmas_cap (2^cap_height x 4 u64) then circuit_digest (4 u64)
 *
 * What a maintainer's Rust `verify_batch(&[ProofTuple])` does through the same five calls.  The circuit
 * parameters below are the reference's standard recursion configuration (CircuitConfig::standard_recursion_config,
 * bn245_poseidon/plonky2_config.rs:78-90): 2^12 rows, rate 1/8, cap height 4, 28 queries, 16 PoW bits, arity-2 folds
 * down to a 32-coefficient final polynomial. */
#include <stdio.h>
#include <stdlib.h>
#include "stark_verifier_b200.h"

static void* slurp(const char* path, size_t* len) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    void* p = NULL;
    if (n < 0 || sv_host_alloc((size_t)n + 8, &p) != 0) { fclose(f); return NULL; }   /* pinned: H2D at full PCIe rate */
    *len = fread(p, 1, (size_t)n, f);
    fclose(f);
    return p;
}

int main(int argc, char** argv) {
    if (argc != 3) { fprintf(stderr, "usage: %s proofs.bin vk.bin\n", argv[0]); return 2; }
    sv_plonk_common common = {4, 80, 135, 2, 9, 8, 4};
    sv_fri_shape shape;
    if (sv_fri_shape_from_common(&common, 12, 3, 4, 28, 16, 7, NULL /* arity 2 throughout */, 0, SV_HASH_POSEIDON_GOLDILOCKS, &shape)) return 1;
    size_t nb = sv_wire_proof_bytes(&shape, &common), len = 0, vk_len = 0;
    sv_ctx* ctx = NULL;
    if (sv_ctx_create(0, &ctx)) { fprintf(stderr, "%s\n", sv_last_error(NULL)); return 1; }   /* no CPU fallback */
    uint8_t* blob = slurp(argv[1], &len);
    uint64_t* vk = slurp(argv[2], &vk_len);
    size_t cap_words = (size_t)4 << shape.cap_height;
    if (!blob || !vk || len % nb || vk_len != (cap_words + 4) * 8) { fprintf(stderr, "bad input sizes (proof = %zu bytes)\n", nb); return 1; }
    size_t n = len / nb;
    uint32_t* bitmap = calloc((n + 31) / 32, 4);
    uint32_t* why = calloc(n, 4);
    int rc = sv_verify_proofs_wire(ctx, &shape, &common, vk, vk + cap_words, blob, nb, n, bitmap, why);
    if (rc) { fprintf(stderr, "sv_verify_proofs_wire: %d %s\n", rc, sv_last_error(ctx)); return 1; }
    size_t ok = 0;
    for (size_t i = 0; i < n; i++) {
        int a = (bitmap[i >> 5] >> (i & 31)) & 1;
        ok += (size_t)a;
        if (!a) printf("proof %zu rejected: query round %u, check %u\n", i, why[i] >> 8, why[i] & 0xFF);
    }
    printf("%zu / %zu proofs accepted\n", ok, n);
    free(bitmap); free(why);
    sv_host_free(blob); sv_host_free(vk);
    sv_ctx_destroy(ctx);
    return ok == n ? 0 : 3;
}
