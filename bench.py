#!/usr/bin/env python3
"""bench.py -- plonky2 proofs verified/s (FRI query phase) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload A|B|merkle]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path (fri_prepare + fused fri_query kernel) over one batch of synthetic
proofs per GPU.  Default workload = BASELINE configs[1]: 4 096 proofs of shape A (2^12 trace, 28 FRI
queries, blowup 8, cap 4, PoW 16) per GPU; proofs shard across ranks with no data-path collective
(weak scaling), and one all-gather of the accept bitmap per step (NCCL) is inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="A", choices=["A", "B", "merkle", "outer"],
                    help="A = configs[1], B = configs[2], merkle = configs[4], outer = the reference's outer wrapped-proof "
                         "configuration (hash family B = Poseidon-BN254, cap_height 0)")
    ap.add_argument("--proofs", type=int, default=0, help="proofs per GPU per step (default 4096 for A, 256 for B)")
    ap.add_argument("--total-proofs", type=int, default=0, help="a fixed batch sharded over the ranks (strong split of e.g. 2^20 proofs, configs[3]) instead of --proofs per GPU")
    ap.add_argument("--distinct", type=int, default=0, help="distinct base proofs generated on the host (default 64 / 4)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="proofs in the cpu_baseline sample (default: sized for ~12 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-wire", action="store_true", help="skip the serialised-proof leg of e2e")
    ap.add_argument("--wire-leg", action="store_true", help="(internal) run only the serialised-proof leg and print its JSON")
    ap.add_argument("--transforms-leg", action="store_true", help="run only the commit-phase kernels (LDE, NTT, commitment) and print their JSON")
    ap.add_argument("--device-transcript", action="store_true",
                    help="derive the Fiat-Shamir challenges on the device too (sv_fri_verify_batch_fs): the records "
                         "enter with their challenge fields zeroed")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2])); power.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index):
    """Pin this process (and therefore its first-touch host allocations, including the pinned record
    buffer) to the CPUs of the NUMA node the GPU hangs off; with 8 ranks copying at once the H2D leg is
    otherwise limited by cross-socket traffic.  Best effort: returns a short description or None."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip().splitlines()[0].strip()
        dom, bus, rest = out.split(":")
        dev = f"{dom[-4:]}:{bus}:{rest}".lower()
        base = f"/sys/bus/pci/devices/{dev}"
        node = int(open(f"{base}/numa_node").read())
        cpulist = open(f"{base}/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node} ({len(cpus)} cpus)"
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


PERM_INSTR, PERM_WIDE, PERM_FP64 = 14579, 2603, 4916


def issue_roofline(perms_per_s, sm_mhz):
    """Ceilings of the permutation-bound kernels on one B200 (148 SMs x 4 sub-partitions x 32 lanes) at the SM
    clock seen during the run: instruction issue, the 32x32->64 multiplier pipe and the FP64 pipe."""
    lanes = 148 * 4 * 32 * sm_mhz * 1e6
    ceil = {"issue": lanes / PERM_INSTR, "int_mul": lanes / 4.24 / PERM_WIDE, "fp64": lanes / 2.18 / PERM_FP64}
    bound = min(ceil, key=ceil.get)
    return {"bound": bound, "achieved": perms_per_s / 1e9, "peak": ceil[bound] / 1e9, "unit": "1e9 permutations/s",
            "frac": perms_per_s / ceil[bound], "frac_int_mul": perms_per_s / ceil["int_mul"], "frac_fp64": perms_per_s / ceil["fp64"],
            "warp_instr_per_perm": PERM_INSTR, "imad_wide_per_perm": PERM_WIDE, "fp64_ops_per_perm": PERM_FP64,
            "kernel": "fri_query_kernel"}


def workload_params(svb, name):
    return {"A": svb.SHAPE_A, "B": svb.SHAPE_B, "outer": svb.SHAPE_OUTER_BN254}[name]


def cpu_baseline(params_shape, record_words, base, sample_proofs, threads):
    """The oracle (CPU restatement of the reference semantics; NOT the Rust binary) timed on the host cores."""
    from oracle import binding as orc
    oshape = orc.shape_from(params_shape)
    n = base.shape[0]
    reps = (sample_proofs + n - 1) // n
    recs = np.ascontiguousarray(np.tile(base, (reps, 1))[:sample_proofs])
    verify = cpu_verify_fn(orc)[0]
    verify(oshape, recs[: min(len(recs), threads)], nthreads=threads)   # warm-up
    t = time.perf_counter()
    bm = verify(oshape, recs, nthreads=threads)
    dt = time.perf_counter() - t
    return sample_proofs / dt, dt, bm, recs


def cpu_verify_fn(orc):
    """(batch verifier of the CPU arm, its description): oracle/fast (AVX-512, 8 Merkle proofs per vector; verdicts bit-identical
    to oracle.c, tests/test_oracle_fast.py) when the CPU has AVX-512, else the scalar oracle.c"""
    if orc.fast_available() and not os.environ.get("SVB_CPU_SCALAR"):
        try:
            orc.fast_lib()
            return orc.fast_fri_verify_batch, "oracle/fast/fast_avx512.c (AVX-512, 8 Merkle proofs per vector) over oracle/oracle.c"
        except OSError:
            pass
    return orc.fri_verify_batch, "oracle/oracle.c (scalar)"


WORKLOAD_TEXT = {"A": "BASELINE configs[1]: {n} proofs/GPU/step, shape A", "B": "BASELINE configs[2]: {n} proofs/GPU/step, shape B",
                 "outer": "outer wrapped-proof configuration (Poseidon-BN254 hash, cap_height 0): {n} proofs/GPU/step"}
# (degree_bits, rate_bits, cap_height, pow_bits, queries, hash_kind) of the bench workloads: the reference arm builds its shape
# from these numbers alone, so that it never has to load the product library
WORKLOAD_SHAPE = {"A": (12, 3, 4, 16, 28, 0), "B": (20, 2, 4, 16, 84, 0), "outer": (12, 3, 0, 16, 28, 1)}


def workload_config(wl, n, distinct, record_bytes, world):
    """The workload definition -- the SAME dict, key for key and value for value, in the b200 arm and in the reference arm
    (what differs between the arms is reported under the top-level "arm" key)."""
    d, r, c, pw, q, kind = WORKLOAD_SHAPE[wl]
    return {"workload": WORKLOAD_TEXT[wl].format(n=n), "hash_kind": kind, "trace_bits": d, "fri_queries": q, "blowup": 1 << r,
            "cap_height": c, "pow_bits": pw, "reduction_arity_bits": [1] * (d - 5), "proofs_per_gpu": n, "distinct_base_proofs": distinct,
            "record_bytes": record_bytes, "l2_policy": f"inputs larger than L2 ({n * record_bytes / 1e6:.0f} MB/GPU resident, physically distinct copies)",
            "sharding": f"proofs sharded over {world} ranks; all-gather of the accept bitmap only",
            "corrupted": "1/64 proofs, five kinds round-robin (" + ", ".join(CORRUPTION_KINDS) + ")"}


def reference_fixture(wl):
    """One valid proof of the workload's shape as a flat record, from a committed fixture (no product code runs):
    A = the pure-Python prover's shape-A proof (tests/golden/pyref_fri.npz, tools/gen_golden_pyref.py), B / outer = the
    round-1 fixtures of the product's host prover (tests/golden/fri_full_shapes.npz, fri_outer_shape.npz)."""
    g = os.path.join(ROOT, "tests", "golden")
    if wl == "A":
        return np.load(os.path.join(g, "pyref_fri.npz"))["shape_a_records"][:1].copy()
    if wl == "B":
        return np.load(os.path.join(g, "fri_full_shapes.npz"))["shape_b_record"][None, :].copy()
    return np.load(os.path.join(g, "fri_outer_shape.npz"))["record"][None, :].copy()


def run_reference(args):
    """--impl reference: the reference's CPU path.  The Rust crate cannot be built in this image (no cargo/rustc,
    dependencies not on disk), so this times the oracle port on all host threads -- on committed fixture proofs: the product
    library (libsvb200.so) is never loaded by this arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding as orc
    wl = "A" if args.workload == "merkle" else args.workload
    d, r, c, pw, q, kind = WORKLOAD_SHAPE[wl]
    oshape = orc.OrcShape()
    oshape.degree_bits, oshape.rate_bits, oshape.cap_height, oshape.num_query_rounds, oshape.proof_of_work_bits = d, r, c, q, pw
    oshape.num_steps, oshape.final_poly_len, oshape.hiding = d - 5, 32, 0
    for i, w in enumerate((84, 135, 20, 16)):
        oshape.oracle_num_polys[i] = w
        oshape.oracle_blinding[i] = int(i > 0)
    oshape.num_zs, oshape.hash_kind = 2, kind
    for i in range(d - 5):
        oshape.reduction_arity_bits[i] = 1
    oL = orc.layout(oshape)
    threads = os.cpu_count() or 1
    base = reference_fixture(wl)
    assert base.shape[1] == oL.record_words
    sample = args.cpu_sample or {"A": 64 * threads, "B": 2 * threads, "outer": max(2, threads // 4)}[wl]
    n_gpu_arm = args.proofs or (4096 if wl == "A" else 256)
    distinct = args.distinct or {"A": 256, "B": 2, "outer": 4}[wl]
    recs = np.ascontiguousarray(np.tile(base, (sample, 1)))
    verify, verify_desc = cpu_verify_fn(orc)
    for _ in range(args.warmup):
        verify(oshape, recs[:threads], nthreads=threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        bm = verify(oshape, recs, nthreads=threads)
    dt = time.perf_counter() - t
    assert all((int(bm[i >> 5]) >> (i & 31)) & 1 for i in range(sample))
    v = sample * args.steps / dt
    # one thread on a small sample: the per-core figure (SURVEY 8d asks for both), and the scalar restatement beside it
    n1 = max(1, min(sample, {"A": 16, "B": 1, "outer": 1}[wl]))
    t = time.perf_counter()
    verify(oshape, recs[:n1], nthreads=1)
    v1 = n1 / (time.perf_counter() - t)
    t = time.perf_counter()
    orc.fri_verify_batch(oshape, recs[:n1], nthreads=1)
    v1_scalar = n1 / (time.perf_counter() - t)
    leaf = [84, 135, 20, 16]
    lde = d + r
    perms_per_query = sum((x + 7) // 8 for x in leaf) + 4 * (lde - c) + sum(lde - (i + 1) - c for i in range(d - 5))
    perms_per_proof = q * perms_per_query
    cpu = open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t") if os.path.exists("/proc/cpuinfo") else "?"
    cfg = workload_config(wl, n_gpu_arm, distinct, oL.record_words * 8, args.gpus)
    arm = {"sample": f"each step verifies a bounded sample of {sample} valid proofs of that workload on the host CPU (no corrupted ones)",
           "transcript": "n/a (the CPU arm verifies records whose challenges are filled in)"}
    print(json.dumps({
        "impl": "reference", "metric": "plonky2_proofs_verified_per_sec", "value": v, "unit": "proofs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": cfg, "arm": arm,
        "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": threads, "kind": "port",
                         "perms_per_sec": v * perms_per_proof, "single_thread_value": v1, "single_thread_perms_per_sec": v1 * perms_per_proof,
                         "perm_ns_per_thread": 1e9 / (v1 * perms_per_proof), "scalar_perm_ns_per_thread": 1e9 / (v1_scalar * perms_per_proof),
                         "implementation": verify_desc,
                         "sample": f"{sample} proofs x {args.steps} steps, {verify_desc} on {threads} threads ({cpu}); "
                                   "CPU restatement of reference semantics, not the Rust binary; inputs: committed fixture proof, "
                                   "the product library is not loaded"},
        "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def wire_leg(svb, torch, ctx, params, L, n_host, distinct, seed, threads, steps):
    """e2e from plonky2 wire bytes: `distinct` proofs bound to hash(public inputs), serialised, tiled into a pinned
    host buffer, 1/64 corrupted (one sibling byte), verified through sv_verify_proofs_wire."""
    import ctypes
    n_pi = 4
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi)
    rng = np.random.default_rng(99)
    pis = rng.integers(0, 0xFFFFFFFF00000001, size=(distinct, n_pi), dtype=np.uint64)
    pih = np.stack([svb.public_inputs_hash(pis[i]) for i in range(distinct)])
    recs = svb.synth_proofs(params, distinct, seed=seed ^ 0x77, n_circuits=1, nthreads=threads, pi_hashes=pih)
    cds, _ = svb.synth_public_inputs(params, distinct, seed=seed ^ 0x77, n_circuits=1)
    vk_cap = recs[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    blob = svb.wire_pack(common, recs, pis)
    nb = blob.shape[1]
    p = ctypes.c_void_p()
    if svb.lib().sv_host_alloc(n_host * nb, ctypes.byref(p)) != 0:
        raise RuntimeError("sv_host_alloc failed")
    try:
        host = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n_host, nb))
        for i in range(0, n_host, distinct):
            c = min(distinct, n_host - i)
            host[i:i + c] = blob[:c]
        q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + len(params.reduction_arity_bits) * 32 * L.ncap
        exp = np.full((n_host + 31) // 32, 0xFFFFFFFF, dtype=np.uint32)
        if n_host & 31:
            exp[-1] = np.uint32((1 << (n_host & 31)) - 1)
        for i in range(32, n_host, 64):
            host[i, q0 + 8 * L.leaf_len[0] + 1 + 7] ^= 1        # a sibling of oracle 0, query round 0
            exp[i >> 5] &= np.uint32(~(1 << (i & 31)) & 0xFFFFFFFF)
        for _ in range(2):
            bm = ctx.verify_proofs_wire(common, vk_cap, cds[0], p.value, n_proofs=n_host)
        if not (bm == exp).all():
            raise RuntimeError("accept bitmap of the wire leg differs from the expected pattern")
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.verify_proofs_wire(common, vk_cap, cds[0], p.value, n_proofs=n_host)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        full = full_leg(svb, torch, ctx, params, common, vk_cap, cds[0], p.value, n_host, steps)
        plonk = plonk_leg(svb, torch, ctx, params, recs, n_host, steps)
        transforms = transforms_leg(svb, torch, ctx, params, steps)
    finally:
        svb.lib().sv_host_free(p)
    return {"plonk_check": plonk, "full_verifier": full, "transforms": transforms, "value": n_host * steps / dt, "unit": "proofs/s", "h2d_bytes_per_step": int(n_host * nb),
            "d2h_bytes_per_step": int(exp.size * 4), "steps": steps, "proof_bytes": int(nb), "public_inputs": n_pi,
            "note": "sv_verify_proofs_wire: pinned plonky2 wire bytes -> H2D -> wire_unpack_kernel + wire_pi_hash_kernel -> "
                    "device transcript -> fri_query_kernel, per 32 MiB chunk"}


def recursion_circuit(svb, params, common):
    """The reference's recursion gate set (chip/plonk/gates/mod.rs:138-196) on the workload's circuit configuration."""
    gates = [(svb.GATE_NOOP, 0), (svb.GATE_CONSTANT, 2), (svb.GATE_PUBLIC_INPUT, 0), (svb.GATE_ARITHMETIC, common.num_routed_wires // 4),
             (svb.GATE_ARITHMETIC_EXT, 10), (svb.GATE_MUL_EXT, 13), (svb.GATE_BASE_SUM, 63), (svb.GATE_REDUCING, 43),
             (svb.GATE_REDUCING_EXT, 32), (svb.GATE_RANDOM_ACCESS, 4, 4, 2), (svb.GATE_POSEIDON_MDS, 0), (svb.GATE_POSEIDON, 0)]
    return svb.make_plonk_circuit(common, gates, [(0, 7), (7, 12)], [pow(7, j, 0xFFFFFFFF00000001) for j in range(common.num_routed_wires)], 123)


def full_leg(svb, torch, ctx, params, common, vk_cap, cd, ptr, n_host, steps):
    """The complete verifier (sv_verify_proofs_full) on the same pinned wire bytes.  The synthetic proofs carry random
    openings, so every proof fails the plonk identity (all bits 0) -- the work done does not depend on that."""
    try:
        circuit = recursion_circuit(svb, params, common)
        for _ in range(2):
            bm = ctx.verify_proofs_full(common, circuit, vk_cap, cd, ptr, n_proofs=n_host)
        if bm.any():
            return {"error": "a synthetic proof satisfied the plonk identity"}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.verify_proofs_full(common, circuit, vk_cap, cd, ptr, n_proofs=n_host)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return {"value": n_host * steps / dt, "unit": "proofs/s", "steps": steps,
                "note": "sv_verify_proofs_full: e2e.wire plus plonk challenges, the vanishing-polynomial identity (recursion gate set) and the AND"}
    except Exception as ex:   # noqa: BLE001
        return {"error": f"{type(ex).__name__}: {ex}"}


def transforms_leg(svb, torch, ctx, params, steps, big_polys=64):
    """Throughput of the commit-phase kernels on device-resident data, timed with CUDA events on the launching stream: the LDE
    of the workload's wires oracle (135 columns of 2^degree_bits coefficients -> 2^lde_bits points, one pass), a batch of
    2^22-point NTTs (two passes of 16 B per element each), the same inverse, and the one-call commitment."""
    import ctypes
    out = {}
    try:
        peak, _ = measured_peak_gbs()
        k, rb = params.degree_bits, params.config.rate_bits
        ncol = params.oracle_num_polys[1]
        n, N = 1 << k, 1 << (k + rb)
        g = torch.Generator(device="cuda").manual_seed(1)
        coeffs = torch.randint(0, 2**62, (ncol, n), dtype=torch.int64, device="cuda", generator=g)
        lde = torch.empty((ncol, N), dtype=torch.int64, device="cuda")
        big = torch.randint(0, 2**62, (big_polys, 1 << 22), dtype=torch.int64, device="cuda", generator=g)
        stream = torch.cuda.Stream()
        torch.cuda.synchronize()            # torch filled the buffers on its own stream
        ctx.set_stream(stream.cuda_stream)

        def timed(fn, reps):
            for _ in range(3):
                fn()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            stream.synchronize()
            ev0.record(stream)
            for _ in range(reps):
                fn()
            ev1.record(stream)
            stream.synchronize()
            return ev0.elapsed_time(ev1) / 1e3 / reps

        t = timed(lambda: ctx.lde_batch(coeffs.data_ptr(), rb, log_n=k, n_polys=ncol, out=lde.data_ptr(), mem=svb.MEM_DEVICE), max(steps, 20))
        lde_bytes = ncol * (n + N) * 8
        out["lde"] = {"ms": 1e3 * t, "polys": ncol, "log_n": k, "log_N": k + rb, "passes": 1, "algorithmic_bytes": lde_bytes,
                      "algorithmic_gbs": lde_bytes / t / 1e9, "frac_of_hbm": lde_bytes / t / 1e9 / peak, "out_gbs": ncol * N * 8 / t / 1e9}
        passes = 2
        ntt_bytes = big_polys * (1 << 22) * 16 * passes
        t = timed(lambda: ctx.ntt_batch(big.data_ptr(), log_n=22, n_polys=big_polys, mem=svb.MEM_DEVICE), steps)
        out["ntt_2^22"] = {"ms": 1e3 * t, "polys": big_polys, "passes": passes, "algorithmic_bytes": ntt_bytes,
                           "algorithmic_gbs": ntt_bytes / t / 1e9, "frac_of_hbm": ntt_bytes / t / 1e9 / peak,
                           "elements_per_s": big_polys * (1 << 22) / t,
                           "note": "round 1 ran this transform in 3 passes (48 B per element): its 578 GB/s were 1.39 ms per 4 polynomials"}
        t = timed(lambda: ctx.ntt_batch(big.data_ptr(), log_n=22, n_polys=big_polys, inverse=True, mem=svb.MEM_DEVICE), steps)
        out["intt_2^22"] = {"ms": 1e3 * t, "polys": big_polys, "passes": passes, "algorithmic_gbs": ntt_bytes / t / 1e9,
                            "frac_of_hbm": ntt_bytes / t / 1e9 / peak}
        del big
        leaves = torch.empty((N, ncol), dtype=torch.int64, device="cuda")
        cap_h = params.config.cap_height
        layers = torch.empty(4 * (2 * N - (1 << cap_h)), dtype=torch.int64, device="cuda")
        vp = ctypes.c_void_p

        def commit(want_leaves):
            rc = ctx._lib.sv_commit_batch(ctx._h, k, rb, ncol, vp(coeffs.data_ptr()), cap_h, params.hash_kind,
                                          vp(leaves.data_ptr()) if want_leaves else None, vp(layers.data_ptr()), svb.MEM_DEVICE)
            if rc != 0:
                raise RuntimeError(f"sv_commit_batch failed: {rc}")
        t = timed(lambda: commit(False), steps)
        t2 = timed(lambda: commit(True), steps)
        perms = N * ((ncol + 7) // 8) + (N - (1 << cap_h))
        out["commit"] = {"ms": 1e3 * t, "ms_with_leaf_major_copy": 1e3 * t2, "leaves": N, "leaf_len": ncol, "permutations": perms,
                         "perms_per_s": perms / t, "note": "LDE + leaf digests straight from the polynomial-major values + Merkle levels of one oracle"}
        ctx.set_stream(0)
    except Exception as ex:   # noqa: BLE001
        out["error"] = f"{type(ex).__name__}: {ex}"
    return out


def plonk_leg(svb, torch, ctx, params, recs, n_host, steps):
    """Throughput of plonk_check_kernel (the vanishing-polynomial identity with the reference's whole recursion gate set)
    on n_host device-resident records.  The synthetic proofs do not satisfy the identity -- the work per proof does not
    depend on that -- so every bit must come back 0."""
    try:
        nch = params.num_zs
        common = svb.CommonData.for_params(params, num_public_inputs=4)
        circuit = recursion_circuit(svb, params, common)
        reps = (n_host + recs.shape[0] - 1) // recs.shape[0]
        d_recs = torch.from_numpy(recs.view(np.int64)).cuda().repeat(reps, 1)[:n_host].contiguous()
        rng = np.random.default_rng(5)
        d_pih = torch.from_numpy(rng.integers(0, 2**63, size=(n_host, 4), dtype=np.int64)).cuda()
        d_chal = torch.from_numpy(rng.integers(0, 2**63, size=(n_host, 3 * nch), dtype=np.int64)).cuda()
        d_bm = torch.ones((n_host + 31) // 32, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        run = lambda: ctx.plonk_check_batch(params, circuit, d_recs.data_ptr(), d_pih.data_ptr(), d_chal.data_ptr(), n_proofs=n_host,
                                            accept_bitmap=d_bm.data_ptr(), mem=svb.MEM_DEVICE)
        for _ in range(2):
            run()
        ctx.synchronize()
        if d_bm.cpu().numpy().any():
            return {"error": "a random opening set satisfied the identity"}
        t0 = time.perf_counter()
        for _ in range(steps):
            run()
        ctx.synchronize()
        dt = time.perf_counter() - t0
        return {"value": n_host * steps / dt, "unit": "proofs/s", "ms_per_call": 1e3 * dt / steps, "proofs_per_call": n_host,
                "note": "plonk_check_kernel, recursion gate set (12 gates incl. PoseidonGate), device-resident records"}
    except Exception as ex:   # noqa: BLE001
        return {"error": f"{type(ex).__name__}: {ex}"}


def run_wire_leg(args):
    import torch
    import stark_verifier_b200 as svb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(0)
    params = workload_params(svb, args.workload)
    L = svb.api.make_layout(params)
    ctx = svb.Context(0)
    threads = len(os.sched_getaffinity(0)) or 1
    out = wire_leg(svb, torch, ctx, params, L, args.proofs, args.distinct, 0xB2000002, threads, args.steps)
    print(json.dumps(out))


CORRUPTION_KINDS = ("sibling limb", "leaf evaluation", "step evaluation", "final-polynomial coefficient", "proof of work")


def record_corruptions(params, L, n, rng):
    """1/64 of the proofs corrupted, the five kinds round-robin -> [(row, word, kind index)].  Kind 4 sets the
    top bit of the squeezed PoW response (the record path carries its challenges; on the wire path it is the PoW witness)."""
    out, S = [], len(params.reduction_arity_bits)
    for j, i in enumerate(range(32, n, 64)):
        kind = j % 5
        q = int(rng.integers(0, params.config.num_query_rounds))
        qb = L.header_words + q * L.query_words
        o = int(rng.integers(0, 4))
        if kind == 2 and S == 0:
            kind = 1
        if kind == 0:
            w = qb + L.q_off_init_sibs[o] + int(rng.integers(0, 4 * L.init_depth))
        elif kind == 1:
            w = qb + L.q_off_init_evals[o] + int(rng.integers(0, L.leaf_len[o]))
        elif kind == 2:
            st = int(rng.integers(0, S))
            w = qb + L.q_off_step_evals[st] + int(rng.integers(0, 2 << params.reduction_arity_bits[st]))
        elif kind == 3:
            w = L.off_final_poly + int(rng.integers(0, 2 * params.final_poly_len()))
        else:
            w = L.off_pow_response
        out.append((i, w, kind))
    return out


def apply_record_corruption(rec, w, kind):
    rec[w] = (int(rec[w]) | (1 << 63)) if kind == 4 else (int(rec[w]) ^ 1)


def wire_corruptions(params, L, common, nb, n, rng):
    """the same five kinds as byte flips inside serialised proofs -> [(proof, byte offset, kind index)]"""
    S = len(params.reduction_arity_bits)
    q0 = 3 * 32 * L.ncap + 16 * (L.n0 + L.n1) + S * 32 * L.ncap
    init = [8 * L.leaf_len[k] + 1 + 32 * L.init_depth for k in range(4)]
    steps = [(16 << params.reduction_arity_bits[i]) + 1 + 32 * L.step_depth[i] for i in range(S)]
    qbytes = sum(init) + sum(steps)
    fin = q0 + params.config.num_query_rounds * qbytes
    out = []
    for j, i in enumerate(range(32, n, 64)):
        kind = j % 5
        q = int(rng.integers(0, params.config.num_query_rounds))
        qb = q0 + q * qbytes
        o = int(rng.integers(0, 4))
        if kind == 2 and S == 0:
            kind = 1
        if kind == 0:
            at = qb + sum(init[:o]) + 8 * L.leaf_len[o] + 1 + int(rng.integers(0, 32 * L.init_depth))
        elif kind == 1:
            at = qb + sum(init[:o]) + int(rng.integers(0, 8 * L.leaf_len[o]))
        elif kind == 2:
            st = int(rng.integers(0, S))
            at = qb + sum(init) + sum(steps[:st]) + int(rng.integers(0, 16 << params.reduction_arity_bits[st]))
        elif kind == 3:
            at = fin + int(rng.integers(0, 16 * params.final_poly_len()))
        else:
            at = fin + 16 * params.final_poly_len()          # the PoW witness
        out.append((i, at, kind))
    assert fin + 16 * params.final_poly_len() + 8 + 8 * common.num_public_inputs == nb
    return out


def oracle_bits(orc, oshape, recs, threads):
    bm = orc.fri_verify_batch(oshape, np.ascontiguousarray(recs), nthreads=threads)
    return [(int(bm[i >> 5]) >> (i & 31)) & 1 for i in range(recs.shape[0])]


def setup_abi_allgather(torch, dist, rank, world, local_rank):
    """An ncclComm_t of our own for sv_allgather_bitmap (the C-ABI collective): unique id from rank 0, broadcast over the
    existing process group.  -> (nccl library handle, comm) or raises."""
    import ctypes
    nccl = ctypes.CDLL("libnccl.so.2", mode=ctypes.RTLD_GLOBAL)      # the library torch already loaded (same soname)

    class UniqueId(ctypes.Structure):
        _fields_ = [("internal", ctypes.c_byte * 128)]
    uid = UniqueId()
    if rank == 0:
        nccl.ncclGetUniqueId.argtypes = [ctypes.POINTER(UniqueId)]
        if nccl.ncclGetUniqueId(ctypes.byref(uid)) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
    t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    ctypes.memmove(ctypes.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, UniqueId, ctypes.c_int]
    if nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) != 0:
        raise RuntimeError("ncclCommInitRank failed")
    return nccl, comm


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.wire_leg:
        return run_wire_leg(args)
    if args.transforms_leg:
        import torch
        import stark_verifier_b200 as svb
        torch.cuda.set_device(0)
        print(json.dumps(transforms_leg(svb, torch, svb.Context(0), workload_params(svb, "A" if args.workload == "merkle" else args.workload),
                                        max(3, min(args.steps, 10)))))
        return

    import ctypes
    import torch
    import torch.distributed as dist
    import stark_verifier_b200 as svb
    from oracle import binding as orc      # the CHECKER of this run (expected bitmaps, cpu_baseline); never on the measured path

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    threads = len(os.sched_getaffinity(0)) or 1
    if args.workload == "merkle":
        return bench_merkle(args, svb, torch, dist, rank, local_rank, world)
    wl = args.workload
    params = workload_params(svb, wl)
    L = svb.api.make_layout(params)
    oshape = orc.shape_from(params.to_shape())
    n = args.proofs or (4096 if wl == "A" else 256)
    n = (n + 31) & ~31
    if args.total_proofs:
        # a FIXED batch cut into contiguous per-rank ranges on bitmap-word boundaries (stark-verifier_b200/shard.py): configs[3]
        first, last = svb.shard.shard_range(args.total_proofs, rank, world)
        n = (last - first + 31) & ~31
        if n == 0:
            raise SystemExit(f"rank {rank}: --total-proofs {args.total_proofs} leaves this rank without work")
    distinct = min(n, args.distinct or {"A": 256, "B": 2, "outer": 4}[wl])
    # ---- synthetic proofs: `distinct` base proofs per rank (own seed) of ONE circuit, each bound to the hash of its own public
    # inputs, so that the same proofs exist as flat records (challenges filled in by the host transcript) and as wire bytes
    t0 = time.perf_counter()
    seed = 0xB2000002 ^ (rank << 20)
    n_pi = 4
    common = svb.CommonData.for_params(params, num_public_inputs=n_pi)
    rng = np.random.default_rng(1234 + rank)
    pis = rng.integers(0, 0xFFFFFFFF00000001, size=(distinct, n_pi), dtype=np.uint64)
    pih = np.stack([svb.public_inputs_hash(pis[i]) for i in range(distinct)])
    gen_threads = max(1, threads // max(1, world))
    base = svb.synth_proofs(params, distinct, seed=seed, n_circuits=1, nthreads=gen_threads, pi_hashes=pih)
    cds, _ = svb.synth_public_inputs(params, distinct, seed=seed, n_circuits=1)
    cd = cds[0]
    vk_cap = base[0, L.off_init_caps:L.off_init_caps + 4 * L.ncap].copy()
    t_gen = time.perf_counter() - t0
    rw = L.record_words
    ctx = svb.Context(local_rank)
    stream = torch.cuda.Stream()          # a real (non-default) stream: handle 0 would mean "ctx's own stream"
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    # ---- device batch: the base proofs tiled ON THE DEVICE into n physically distinct records (config 3 is 54 GB -- it never
    # exists on the host), then the seeded negative controls: 1/64 of the proofs, five corruption kinds round-robin
    d_recs = torch.empty((n, rw), dtype=torch.int64, device="cuda")
    d_recs[:distinct].copy_(torch.from_numpy(base.view(np.int64)))
    k = distinct
    while k < n:
        c = min(k, n - k)
        d_recs[k:k + c].copy_(d_recs[:c])
        k += c
    cor = record_corruptions(params, L, n, rng)
    # expected bitmap FROM THE ORACLE, in this run: every base proof, and every corrupted record rebuilt on the host
    base_bits = oracle_bits(orc, oshape, base, threads)
    bad_recs = np.stack([base[i % distinct] for i, _, _ in cor]) if cor else np.zeros((0, rw), dtype=np.uint64)
    for r, (_, w, kind) in zip(bad_recs, cor):
        apply_record_corruption(r, w, kind)
    bad_bits = oracle_bits(orc, oshape, bad_recs, threads) if cor else []
    if not all(base_bits) or any(bad_bits):
        raise SystemExit(f"rank {rank}: the oracle rejects a base proof or accepts a corrupted one")
    exp = np.zeros(n // 32, dtype=np.uint32)
    for i in range(n):
        if base_bits[i % distinct]:
            exp[i >> 5] |= np.uint32(1 << (i & 31))
    for (i, _, _), b in zip(cor, bad_bits):
        if not b:
            exp[i >> 5] &= np.uint32(~(1 << (i & 31)) & 0xFFFFFFFF)
    if cor:
        r_t = torch.tensor([i for i, _, _ in cor], device="cuda")
        c_t = torch.tensor([w for _, w, _ in cor], device="cuda")
        vals = torch.from_numpy(np.array([r[w] for r, (_, w, _) in zip(bad_recs, cor)], dtype=np.uint64).view(np.int64)).cuda()
        d_recs[r_t, c_t] = vals
    words = n // 32
    d_bm = torch.zeros(words, dtype=torch.int32, device="cuda")
    d_all = torch.zeros(words * world, dtype=torch.int32, device="cuda")
    d_all_torch = torch.zeros(words * world, dtype=torch.int32, device="cuda")
    # the accept-bitmap all-gather goes through the C ABI (sv_allgather_bitmap on a communicator of our own); torch's gather
    # is the cross-check after the timed region, and the fallback if the communicator cannot be made
    abi = None
    gather_via = "n/a (1 rank)"
    if world > 1:
        try:
            abi = setup_abi_allgather(torch, dist, rank, world, local_rank)
            gather_via = "sv_allgather_bitmap (C ABI, ncclAllGather on the ctx stream)"
        except Exception as ex:   # noqa: BLE001
            gather_via = f"torch.distributed.all_gather_into_tensor (C-ABI communicator unavailable: {type(ex).__name__}: {ex})"
    torch.cuda.synchronize()

    def step():
        ctx.fri_verify_batch(params, d_recs.data_ptr(), n_proofs=n, accept_bitmap=d_bm.data_ptr(), mem=svb.MEM_DEVICE)
        if world > 1:
            if abi is not None:
                ctx.allgather_bitmap(abi[1].value, d_bm.data_ptr(), d_all.data_ptr(), words)
            else:
                dist.all_gather_into_tensor(d_all, d_bm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    got = d_bm.cpu().numpy().view(np.uint32)
    if not (got == exp).all():
        raise SystemExit(f"rank {rank}: accept bitmap differs from the oracle's")
    if world > 1:
        dist.all_gather_into_tensor(d_all_torch, d_bm)
        torch.cuda.synchronize()
        if not bool((d_all == d_all_torch).all()):
            raise SystemExit(f"rank {rank}: sv_allgather_bitmap and torch's all-gather disagree")

    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.kernel_timing(True)
    l0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    kernel_ms, kernel_n = ctx.kernel_time_ms()
    ctx.kernel_timing(False)
    launches = ctx.launch_count - l0 + (args.steps if world > 1 else 0)
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * n * args.steps / (ms / 1e3)

    def timed_events(fn, reps):
        """device time of `reps` resident calls on the ctx stream, max over ranks"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        barrier()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item()) / 1e3

    # resident data with the Fiat-Shamir transcript on the device as well (sv_fri_verify_batch_fs): challenge fields zeroed
    resident_fs = None
    try:
        if n * rw * 8 > 40e9:
            raise RuntimeError("skipped: a second copy of the records would not fit beside the first")
        d_fs = d_recs.clone()
        d_fs[:, L.off_alpha:L.header_words] = 0
        ph_dev = torch.from_numpy(np.ascontiguousarray(np.tile(pih, ((n + distinct - 1) // distinct, 1))[:n]).view(np.int64)).cuda()
        d_bm2 = torch.zeros(words, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        run_fs = lambda: ctx.fri_verify_batch_fs(params, d_fs.data_ptr(), cd, ph_dev.data_ptr(), n_proofs=n, accept_bitmap=d_bm2.data_ptr(),
                                                 mem=svb.MEM_DEVICE)
        for _ in range(2):
            run_fs()
        torch.cuda.synchronize()
        # a corrupted PoW response is recomputed by the device transcript: those proofs are valid again
        exp_fs = exp.copy()
        for i, _, kind in cor:
            if kind == 4:
                exp_fs[i >> 5] |= np.uint32(1 << (i & 31))
        if not (d_bm2.cpu().numpy().view(np.uint32) == exp_fs).all():
            raise RuntimeError("bitmap of the device-transcript leg differs from the oracle's")
        dt = timed_events(run_fs, args.steps)
        resident_fs = {"value": world * n * args.steps / dt, "unit": "proofs/s", "steps": args.steps,
                       "note": "sv_fri_verify_batch_fs on resident records: device transcript (one launch per batch) + prepare + query"}
        del d_fs
    except Exception as ex:   # noqa: BLE001
        resident_fs = {"error": f"{type(ex).__name__}: {ex}"}

    # ---- end to end through the public C ABI with HOST buffers (pinned).  The headline leg starts from SERIALISED proofs: plonky2
    # wire bytes -> H2D -> device unpack, public-input hashes, Fiat-Shamir transcript, query phase -> D2H of the bitmap
    # (sv_verify_proofs_wire).  Sub-legs: the same bytes through the complete verifier; flat records with host-side challenges
    # (sv_fri_verify_batch) and with the device transcript; the copy-only ceiling of the same bytes.
    e2e = None
    if not args.no_e2e:
        ctx.set_stream(0)
        n_host = min(n, 4096 if wl == "A" else 512)
        hwords = n_host // 32
        e2e_steps = args.steps
        blob = svb.wire_pack(common, base, pis)
        nb = blob.shape[1]
        p = ctypes.c_void_p()
        if svb.lib().sv_host_alloc(n_host * nb, ctypes.byref(p)) != 0:
            raise SystemExit("sv_host_alloc failed")
        try:
            hostb = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n_host, nb))
            for i in range(0, n_host, distinct):
                c = min(distinct, n_host - i)
                hostb[i:i + c] = blob[:c]
            wcor = wire_corruptions(params, L, common, nb, n_host, np.random.default_rng(77 + rank))
            for i, at, _ in wcor:
                hostb[i, at] ^= 1
            # the oracle's verdict on the corrupted proofs, from their bytes (its own reader and transcript)
            ocommon = orc.common_from(common.to_c())
            exp_w = np.zeros(hwords, dtype=np.uint32)
            for i in range(n_host):
                if base_bits[i % distinct]:
                    exp_w[i >> 5] |= np.uint32(1 << (i & 31))
            for i, _, _ in wcor:
                rc, rec_o, _, pih_o = orc.wire_read_proof(oshape, ocommon, vk_cap, hostb[i])
                ok = False
                if rc == 0:
                    orc.fri_challenges(oshape, rec_o, cd, pih_o, common.num_challenges)
                    ok = bool(orc.fri_verify(oshape, rec_o)[0])
                if not ok:
                    exp_w[i >> 5] &= np.uint32(~(1 << (i & 31)) & 0xFFFFFFFF)

            def time_host(call, reps):
                for _ in range(2):
                    call()
                barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    call()
                barrier()
                tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
                if world > 1:
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                return float(tt.item())

            hb = {}

            def wire_call():
                hb["bm"] = ctx.verify_proofs_wire(common, vk_cap, cd, p.value, n_proofs=n_host)
            dt = time_host(wire_call, e2e_steps)
            if not (hb["bm"] == exp_w).all():
                raise SystemExit(f"rank {rank}: accept bitmap of the wire leg differs from the oracle's")
            e2e = {"value": world * n_host * e2e_steps / dt, "unit": "proofs/s", "h2d_bytes_per_step": int(n_host * nb),
                   "d2h_bytes_per_step": int(hwords * 4), "steps": e2e_steps, "proofs_per_step_per_gpu": n_host, "host_binding": numa,
                   "proof_bytes": int(nb), "public_inputs": n_pi,
                   "entry_point": "sv_verify_proofs_wire, one synchronous call per step (bytes -> verdict: unpack, public-inputs hash, "
                                  "Fiat-Shamir transcript and query phase on the device; headers first, transcript in two parts)",
                   "corrupted": "1/64 proofs, five kinds round-robin as byte flips; bitmap == the oracle's (its own wire reader + transcript)"}
            if world == 1 and wl == "A":
                # the same entry point on a 4x larger batch per call: the ~2.6 ms before the first query kernel can start (header
                # copies + the first transcript part, the one latency a lone call cannot hide) amortise over 4x the bytes
                big = 4 * n_host
                pb = ctypes.c_void_p()
                if svb.lib().sv_host_alloc(big * nb, ctypes.byref(pb)) == 0:
                    try:
                        for r_ in range(4):
                            ctypes.memmove(pb.value + r_ * n_host * nb, p.value, n_host * nb)
                        hb2 = {}

                        def big_call():
                            hb2["bm"] = ctx.verify_proofs_wire(common, vk_cap, cd, pb.value, n_proofs=big)
                        dtb = time_host(big_call, max(2, e2e_steps // 4))
                        if not (hb2["bm"] == np.tile(exp_w, 4)).all():
                            raise SystemExit("accept bitmap of the large-batch wire leg differs from the oracle's")
                        e2e["large_batch"] = {"value": big * max(2, e2e_steps // 4) / dtb, "unit": "proofs/s", "proofs_per_call": big,
                                              "h2d_bytes_per_call": int(big * nb)}
                    finally:
                        svb.lib().sv_host_free(pb)
            if world == 1 and wl == "A":
                # two host threads, each with its own sv_ctx (the header's threading contract: one context per (host thread, GPU),
                # distinct contexts fully concurrent), each making the same synchronous call on the same pinned bytes: the head
                # latency of one call (header copies + first transcript part) hides behind the query kernels of the other
                try:
                    import threading
                    ctx2 = svb.Context(local_rank)
                    res = [None, None]
                    reps2 = max(2, e2e_steps // 2)

                    def worker(k, cx, reps):
                        for _ in range(reps):
                            res[k] = cx.verify_proofs_wire(common, vk_cap, cd, p.value, n_proofs=n_host)
                    worker(1, ctx2, 2)                                  # warm-up of the second context (allocations)
                    torch.cuda.synchronize()
                    ths = [threading.Thread(target=worker, args=(k, cx, reps2)) for k, cx in enumerate((ctx, ctx2))]
                    t0 = time.perf_counter()
                    for th in ths:
                        th.start()
                    for th in ths:
                        th.join()
                    torch.cuda.synchronize()
                    dt2 = time.perf_counter() - t0
                    if not ((res[0] == exp_w).all() and (res[1] == exp_w).all()):
                        raise RuntimeError("accept bitmap of the two-context leg differs from the oracle's")
                    e2e["two_contexts"] = {"value": 2 * reps2 * n_host / dt2, "unit": "proofs/s", "calls": 2 * reps2, "proofs_per_call": n_host,
                                           "note": "two host threads x their own sv_ctx, synchronous sv_verify_proofs_wire calls in flight together"}
                    del ctx2
                except Exception as ex:   # noqa: BLE001
                    e2e["two_contexts"] = {"error": f"{type(ex).__name__}: {ex}"}
            # copy-only ceiling of the same bytes
            stage = torch.empty(n_host * nb, dtype=torch.uint8, device="cuda")
            hview = torch.from_numpy(hostb.reshape(-1))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                stage.copy_(hview, non_blocking=True)
            torch.cuda.synchronize()
            h2d_s = (time.perf_counter() - t0) / 3
            del stage
            e2e["h2d_only_gbs"] = n_host * nb / h2d_s / 1e9
            e2e["h2d_only_proofs_per_s"] = world * n_host / h2d_s
            e2e["frac_of_h2d_only"] = e2e["value"] / e2e["h2d_only_proofs_per_s"]
            # sub-leg: flat records with host-side challenges (the round-1 headline), and with the device transcript
            host = torch.empty((n_host, rw), dtype=torch.int64).pin_memory()
            host.copy_(d_recs[:n_host])
            hbm = np.zeros(hwords, dtype=np.uint32)
            dt = time_host(lambda: ctx.fri_verify_batch(params, host.data_ptr(), n_proofs=n_host, accept_bitmap=hbm, mem=svb.MEM_HOST), e2e_steps)
            assert (hbm == exp[:hwords]).all()
            e2e["record_path"] = {"value": world * n_host * e2e_steps / dt, "unit": "proofs/s", "h2d_bytes_per_step": int(n_host * rw * 8),
                                  "note": "sv_fri_verify_batch(SV_MEM_HOST): flat records whose challenges the host derived"}
            ph_host = np.ascontiguousarray(np.tile(pih, ((n_host + distinct - 1) // distinct, 1))[:n_host])
            dt = time_host(lambda: ctx.fri_verify_batch_fs(params, host.data_ptr(), cd, ph_host, n_proofs=n_host, accept_bitmap=hbm,
                                                           mem=svb.MEM_HOST), e2e_steps)
            e2e["record_path_device_transcript"] = {"value": world * n_host * e2e_steps / dt, "unit": "proofs/s",
                                                    "note": "sv_fri_verify_batch_fs(SV_MEM_HOST)"}
            del host
            if world == 1 and wl != "B":
                e2e["full_verifier"] = full_leg(svb, torch, ctx, params, common, vk_cap, cd, p.value, n_host, e2e_steps)
                e2e["plonk_check"] = plonk_leg(svb, torch, ctx, params, base, n_host, e2e_steps)
                if not args.no_wire:
                    e2e["transforms"] = transforms_leg(svb, torch, ctx, params, min(e2e_steps, 10))
        finally:
            svb.lib().sv_host_free(p)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    # DRAM traffic of one launch of the dominant kernel: from the committed ncu capture of THIS library revision and workload only
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))
        if tj.get("sv_version") == svb.api.version():
            traffic = tj.get("bytes", {}).get(f"{wl}:{n}")
            traffic_src = tj.get("source")
        else:
            traffic_src = f"profiles/traffic_r2.json is for library {tj.get('sv_version')}, this is {svb.api.version()}: not reported"
    except Exception:
        pass
    algo_bytes = n * (params.config.num_query_rounds * L.algo_bytes_per_query + L.algo_bytes_shared)
    k_avg_s = (kernel_ms / max(1, kernel_n)) / 1e3
    achieved = algo_bytes / k_avg_s / 1e9
    perms = n * params.config.num_query_rounds * L.perms_per_query
    cfg = workload_config(wl, n, distinct, rw * 8, world)
    arm = {"transcript": "host (challenges arrive in the records) for `value`; on the device for `e2e`",
           "bitmap_check": "accept bitmap == the oracle's, computed in this run", "bitmap_gather": gather_via}
    if args.total_proofs:
        cfg["total_proofs"] = args.total_proofs
        cfg["workload"] = f"BASELINE configs[3]: {args.total_proofs} shape-A proofs sharded over {world} GPUs ({n} per GPU, shard.py)" if wl == "A" else cfg["workload"]
    out = {
        "metric": "plonky2_proofs_verified_per_sec", "value": value, "unit": "proofs/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.total_proofs else "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": cfg, "arm": arm,
        "e2e": e2e,
        "resident_with_device_transcript": resident_fs,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "kernel": "fri_query_kernel", "kernel_ms": kernel_ms / max(1, kernel_n),
                     "kernel_share_of_step": kernel_ms / ms, "algorithmic_bytes_per_launch": int(algo_bytes),
                     "peak_source": peak_src,
                     "note": "integer-issue bound, not HBM bound (SURVEY 8d): see issue_roofline"},
        "perms_per_sec": world * perms * args.steps / (ms / 1e3),
        # what actually binds: instruction issue.  One Poseidon-Goldilocks permutation executes
        # PERM_INSTR warp-instructions (tools/sass_dyn.py on the shipped library), of which PERM_WIDE IMAD.WIDE.U32
        # (4.24 cycles each on the multiplier pipe, profiles/pipes2_b200_r1.txt) and PERM_FP64 DADD/DFMA (2.18 cycles each);
        # a sub-partition issues at most one warp-instruction per cycle.
        "issue_roofline": None if params.hash_kind else issue_roofline(perms / k_avg_s, clocks.get("sm_mhz") or 1965.0),
        "synth_seconds": t_gen,
        "library": svb.api.version(),
    }
    if not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)      # the CPU baseline gets every host core
        threads = len(all_cpus)
        sample = args.cpu_sample
        shape_c = params.to_shape()
        if not sample:
            # probe on a small batch, then size the sample for ~12 s of CPU work on all host threads
            probe = 8 * threads if wl == "A" else threads
            if wl == "outer":
                probe = max(4, threads // 4)
            v0, _, _, _ = cpu_baseline(shape_c, rw, base, probe, threads)
            sample = max(probe, int(v0 * 12.0) // threads * threads)
        v, dt, bm_cpu, recs_cpu = cpu_baseline(shape_c, rw, base, sample, threads)
        # the same sample through the GPU: bit-for-bit the oracle's bitmap
        bm_gpu = ctx.fri_verify_batch(params, recs_cpu)
        if not (bm_gpu == bm_cpu).all():
            raise SystemExit("GPU and oracle bitmaps differ on the cpu_baseline sample")
        v1, _, _, _ = cpu_baseline(shape_c, rw, base, max(1, min(sample, {"A": 16, "B": 1, "outer": 1}[wl])), 1)
        perms_per_proof = params.config.num_query_rounds * L.perms_per_query
        out["cpu_baseline"] = {"value": v, "unit": "proofs/s", "cores": threads, "kind": "port",
                               "perms_per_sec": v * perms_per_proof, "single_thread_value": v1,
                               "single_thread_perms_per_sec": v1 * perms_per_proof, "perm_ns_per_thread": 1e9 / (v1 * perms_per_proof),
                               "gpu_bitmap_equal_on_sample": True, "implementation": cpu_verify_fn(orc)[1],
                               "sample": f"{sample} proofs of the same workload in {dt:.1f} s on {threads} threads; {cpu_verify_fn(orc)[1]}, "
                                         "CPU restatement of reference semantics (not the Rust binary)"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def bench_merkle(args, svb, torch, dist, rank, local_rank, world):
    """BASELINE configs[4]: 2^24 independent Merkle paths x depth 20, 4-limb leaves, cap_height 0."""
    n = args.proofs or (1 << 24)
    depth, leaf_len = 20, 4
    rec_words = leaf_len + 4 * depth
    g = torch.Generator(device="cuda"); g.manual_seed(0xB2000005 + rank)
    # uniform canonical field elements on the device: 63-bit randoms are < p
    paths = torch.randint(0, 1 << 62, (n, rec_words), dtype=torch.int64, device="cuda", generator=g)
    idx = torch.randint(0, 1 << depth, (n,), dtype=torch.int64, device="cuda", generator=g)
    caps = torch.zeros(4, dtype=torch.int64, device="cuda")
    ok = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ctx = svb.Context(local_rank)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def step():
        ctx.merkle_verify_batch(leaf_len, depth, 0, paths.data_ptr(), idx.data_ptr(), caps.data_ptr(), ok.data_ptr(), n=n, mem=svb.MEM_DEVICE)

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    # parity on a sampled subset against the oracle
    from oracle import binding as orc
    sub = 4096
    hp = paths[:sub].cpu().numpy().view(np.uint64); hi = idx[:sub].cpu().numpy().view(np.uint64)
    # give the first half of the subset a matching root so both outcomes occur
    want = orc.merkle_verify_batch(hp, leaf_len, depth, hi, np.zeros(4, dtype=np.uint64), 0)
    assert (ok[:sub].cpu().numpy() == want).all()
    sampler = ClockSampler(local_rank); sampler.start()
    ctx.kernel_timing(True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    kernel_ms, kernel_n = ctx.kernel_time_ms()
    clocks = sampler.stop()
    peak, peak_src = measured_peak_gbs()
    algo = n * (32 + 32 * depth)
    achieved = algo / (kernel_ms / kernel_n / 1e3) / 1e9
    if rank == 0:
        print(json.dumps({
            "metric": "merkle_paths_verified_per_sec", "value": world * n * args.steps / (ms / 1e3), "unit": "paths/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[4]: {n} Merkle paths x depth {depth}, 4-limb leaves, cap_height 0",
                       "l2_policy": f"inputs larger than L2 ({n * rec_words * 8 / 1e9:.2f} GB resident)"},
            "gpu_launches": args.steps, "clocks": clocks,
            "perms_per_sec": world * n * depth * args.steps / (ms / 1e3),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "merkle_verify_kernel", "kernel_ms": kernel_ms / kernel_n,
                         "algorithmic_bytes_per_launch": int(algo), "peak_source": peak_src},
            "issue_roofline": dict(issue_roofline(n * depth / (kernel_ms / kernel_n / 1e3), clocks.get("sm_mhz") or 1965.0),
                                   kernel="merkle_verify_kernel"),
        }))


if __name__ == "__main__":
    main()
