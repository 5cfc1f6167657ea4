#!/usr/bin/env python3
"""Dry run of the GPU test files whose kernels were written without a GPU, on the CPU: a mock `Context` whose methods
are the HOST TWINS of the device entry points (the same __host__ __device__ functions) composed with the CPU oracle.

It checks the TEST LOGIC -- indices of the corrupted bytes, expected verdicts and first-failure codes, shapes, the Python
wrappers' argument handling -- not the kernels; the kernels' bodies are checked by the CPU tests of the host twins
(tests/test_wire_format.py, test_plonk_check.py, test_ntt.py, test_full_proof.py) and meet a GPU in `pytest -m gpu`.

    python tools/dryrun_gpu_tests.py
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import stark_verifier_b200 as svb  # noqa: E402
from oracle import binding as orc  # noqa: E402
import full_prover as fp  # noqa: E402


def bitmap(bits):
    bm = np.zeros((len(bits) + 31) // 32, dtype=np.uint32)
    for i, b in enumerate(bits):
        if b:
            bm[i >> 5] |= np.uint32(1 << (i & 31))
    return bm


class MockCtx:
    """CPU stand-in for svb.Context: same signatures, same verdict composition and first-failure order."""

    # -- wire ------------------------------------------------------------------------------------
    def wire_unpack_batch(self, common, cap, blob, n_proofs=None, stride=None, mem=0, **kw):
        assert mem == 0
        r, h, _, m = svb.wire_unpack_batch(common, cap, blob, n_proofs=n_proofs, stride=stride)
        return r, h, m.astype(np.uint32)

    def _unpacked(self, common, cap, cd, blob, n_proofs, stride):
        nb = svb.wire_proof_bytes(common)
        stride = nb if stride is None else stride
        if not isinstance(blob, np.ndarray):
            blob = np.ctypeslib.as_array(ctypes.cast(blob, ctypes.POINTER(ctypes.c_uint8)), shape=(n_proofs * stride,))
        n = blob.size // stride if n_proofs is None else n_proofs
        if n == 0:
            return None
        r, h, _, m = svb.wire_unpack_batch(common, cap, blob, n_proofs=n, stride=stride)
        nch = common.num_challenges
        chal = np.stack([svb.plonk_challenges(common.fri_params, r[i], cd, h[i], nch) for i in range(n)])
        for i in range(n):
            svb.fri_challenges(common.fri_params, r[i], cd, h[i], nch)
        return r, h, m, chal

    def _verify(self, common, circuit, cap, cd, blob, n_proofs, stride, want_fail):
        u = self._unpacked(common, cap, cd, blob, n_proofs, stride)
        if u is None:
            return (np.zeros(0, dtype=np.uint32), np.zeros(0, dtype=np.uint32)) if want_fail else np.zeros(0, dtype=np.uint32)
        r, h, m, chal = u
        osh = orc.shape_from(common.fri_params.to_shape())
        n = r.shape[0]
        pl = svb.plonk_check_host(common.fri_params, circuit, r, h, chal) if circuit is not None else None
        bits, ff = [], np.zeros(n, dtype=np.uint32)
        for i in range(n):
            ok, code, q = orc.fri_verify(osh, r[i])
            if not ok:
                ff[i] = (max(q, 0) << 8) | code
            if pl is not None and not (int(pl[i >> 5]) >> (i & 31)) & 1:
                ok, ff[i] = False, svb.FAIL_PLONK
            if m[i]:
                ok, ff[i] = False, svb.FAIL_MALFORMED
            bits.append(int(ok))
        return (bitmap(bits), ff) if want_fail else bitmap(bits)

    def verify_proofs_wire(self, common, cap, cd, blob, n_proofs=None, stride=None, want_fail=False):
        return self._verify(common, None, cap, cd, blob, n_proofs, stride, want_fail)

    def verify_proofs_full(self, common, circuit, cap, cd, blob, n_proofs=None, stride=None, want_fail=False):
        return self._verify(common, circuit, cap, cd, blob, n_proofs, stride, want_fail)

    # -- records ---------------------------------------------------------------------------------
    def fri_verify_batch(self, params, recs, n=None, want_fail=False, **kw):
        osh = orc.shape_from(params.to_shape())
        bm = orc.fri_verify_batch(osh, recs)
        if not want_fail:
            return bm
        ff = np.zeros(recs.shape[0], dtype=np.uint32)
        for i in range(recs.shape[0]):
            ok, code, q = orc.fri_verify(osh, recs[i])
            ff[i] = 0 if ok else (max(q, 0) << 8) | code
        return bm, ff

    def fri_challenges_batch(self, params, recs, cd, pih, num_challenges=2, **kw):
        for i in range(recs.shape[0]):
            svb.fri_challenges(params, recs[i], cd, pih[i], num_challenges)
        return recs

    def fri_verify_batch_fs(self, params, recs, cd, pih, num_challenges=2, **kw):
        r = recs.copy()
        for i in range(r.shape[0]):
            svb.fri_challenges(params, r[i], cd, pih[i], num_challenges)
        return orc.fri_verify_batch(orc.shape_from(params.to_shape()), r)

    def plonk_check_batch(self, params, circuit, recs, pih, chal, **kw):
        return svb.plonk_check_host(params, circuit, recs, pih, chal)

    # -- transforms ------------------------------------------------------------------------------
    def ntt_batch(self, data, inverse=False, **kw):
        return svb.ntt_host(data, inverse=inverse)

    def lde_batch(self, c, rate_bits, shift=7, **kw):
        return svb.lde_host(c, rate_bits, shift=shift)

    def merkle_tree_build(self, leaves, leaf_len, cap_height, hash_kind=0, **kw):
        from pyref.merkle import MerkleTree
        t = MerkleTree(np.asarray(leaves, dtype=np.uint64).reshape(len(leaves), -1), cap_height, hash_kind)
        return [np.array(layer, dtype=np.uint64) for layer in t.layers]

    def commit_batch(self, coeffs, rate_bits, cap_height, hash_kind=0):
        leaves = np.ascontiguousarray(svb.lde_host(coeffs, rate_bits).T)
        return leaves, self.merkle_tree_build(leaves, leaves.shape[1], cap_height, hash_kind)


class PinnedShim:
    """stands in for the module in tests that allocate pinned memory through sv_host_alloc (needs CUDA)."""

    def __init__(self):
        self._keep = {}

    def __getattr__(self, k):
        return getattr(svb, k)

    def lib(self):
        real, keep = svb.lib(), self._keep

        class L:
            def __getattr__(self, k):
                return getattr(real, k)

            @staticmethod
            def sv_host_alloc(n, pp):
                buf = ctypes.create_string_buffer(n)
                keep[ctypes.addressof(buf)] = buf
                pp._obj.value = ctypes.addressof(buf)
                return 0

            @staticmethod
            def sv_host_free(p):
                return 0

        return L()


def params_of(fn):
    for m in getattr(fn, "pytestmark", []):
        if m.name == "parametrize":
            names = [x.strip() for x in m.args[0].split(",")]
            yield names, m.args[1]


def run(fn, fixtures):
    """call a test function for the cross product of its parametrize marks (device-memory tests are skipped)."""
    import inspect
    import itertools
    want = list(inspect.signature(fn).parameters)
    marks = list(params_of(fn))
    combos = itertools.product(*[[dict(zip(names, v if isinstance(v, (tuple, list)) and len(names) > 1 else (v,))) for v in values]
                                 for names, values in marks]) if marks else [()]
    count = 0
    for combo in combos:
        kw = dict(fixtures)
        for d in combo:
            kw.update(d)
        fn(**{k: kw[k] for k in want})
        count += 1
    return count


def main():
    import test_gpu_arity as t_ar
    import test_gpu_parity_python_prover as t_pp
    import test_gpu_pyref_golden as t_gold
    import test_gpu_plonk as t_plonk
    import test_gpu_transforms as t_tr
    import test_gpu_verify_full as t_full
    import test_gpu_wire as t_wire
    ctx = MockCtx()
    fx = dict(svb=svb, orc=orc, ctx=ctx)
    skip = {"test_unpack_kernel_device_memory", "test_plonk_kernel_device_memory", "test_ntt_device_memory_and_many_polys"}  # need torch.cuda
    if "--quick" in sys.argv:
        # the CPU suite's budget: the four slowest dry runs (~110 s of mock verification) only run without --quick; every one of
        # these test functions has been green on real B200s since round 2 (profiles/, GPUTEST records)
        skip |= {"test_full_verifier_matches_cpu_side", "test_verify_proofs_wire_matches_oracle",
                 "test_fri_kernel_accepts_python_prover_proofs", "test_fri_arity_small_shapes"}
    total = 0
    for mod in (t_gold, t_ar, t_pp, t_plonk, t_tr, t_full, t_wire):
        for name in sorted(n for n in dir(mod) if n.startswith("test_") and n not in skip):
            f = dict(fx)
            if name == "test_verify_proofs_wire_many_chunks":
                f["svb"] = PinnedShim()
            k = run(getattr(mod, name), f)
            total += k
            print(f"{mod.__name__}.{name}: {k} case(s) ok")
    print(f"dry run ok: {total} cases")


if __name__ == "__main__":
    main()
