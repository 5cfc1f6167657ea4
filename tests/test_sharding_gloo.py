"""N>1 path on CPU: world_size-2 gloo.  Each rank takes its shard of a proof batch, produces its accept
words (the oracle stands in for the GPU verifier here -- host-side logic only) and the packed bitmaps
are all-gathered exactly as bench.py / the GPU path do."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from importlib import import_module
    import stark_verifier_b200 as svb
    from oracle import binding as orc
    from common import corrupt, tiny_params
    shard = import_module("stark-verifier_b200.shard")
    params = tiny_params(svb, queries=3, degree_bits=6)
    L = svb.api.make_layout(params)
    recs = svb.synth_proofs(params, n, seed=42, n_circuits=1, nthreads=2)     # same batch on every rank
    corrupt(recs, L, np.random.default_rng(0), every=8, num_steps=1)
    first, last = shard.shard_range(n, rank, world)
    w = shard.words_per_rank(n, world)
    local = np.zeros(w, dtype=np.uint32)
    oshape = orc.shape_from(params.to_shape())
    if last > first:
        bm = orc.fri_verify_batch(oshape, recs[first:last], nthreads=1)
        local[: len(bm)] = bm
    full = shard.gather_bitmap(torch.from_numpy(local.view(np.int32)), world).numpy().view(np.uint32)
    want = orc.fri_verify_batch(oshape, recs, nthreads=1)
    assert (full[: len(want)] == want).all()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [100, 64])
def test_sharded_bitmap_allgather_gloo(tmp_path, n):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert (a == b).all()
