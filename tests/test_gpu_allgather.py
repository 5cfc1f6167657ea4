"""sv_allgather_bitmap over real NCCL communicators (needs >= 2 GPUs; skipped otherwise)."""
import ctypes
import ctypes.util

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_allgather_bitmap_two_gpus(svb):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    # torch bundles its own NCCL; load the same library the C ABI resolves at run time
    try:
        nccl = ctypes.CDLL("libnccl.so.2", mode=ctypes.RTLD_GLOBAL)
    except OSError:
        import glob
        import os
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
        if not cands:
            pytest.skip("libnccl.so.2 not found")
        nccl = ctypes.CDLL(cands[0], mode=ctypes.RTLD_GLOBAL)
    ndev = 2
    comms = (ctypes.c_void_p * ndev)()
    devs = (ctypes.c_int * ndev)(0, 1)
    assert nccl.ncclCommInitAll(comms, ndev, devs) == 0
    ctxs = [svb.Context(i) for i in range(ndev)]
    words = 8
    local, full = [], []
    for i in range(ndev):
        with torch.cuda.device(i):
            local.append(torch.arange(words, dtype=torch.int32, device=f"cuda:{i}") + 1000 * (i + 1))
            full.append(torch.zeros(words * ndev, dtype=torch.int32, device=f"cuda:{i}"))
    for i in range(ndev):
        torch.cuda.synchronize(i)
    assert nccl.ncclGroupStart() == 0
    for i in range(ndev):
        ctxs[i].allgather_bitmap(comms[i], local[i].data_ptr(), full[i].data_ptr(), words)
    assert nccl.ncclGroupEnd() == 0
    want = np.concatenate([np.arange(words) + 1000 * (i + 1) for i in range(ndev)]).astype(np.int32)
    for i in range(ndev):
        ctxs[i].synchronize()
        assert (full[i].cpu().numpy() == want).all()
    for i in range(ndev):
        nccl.ncclCommDestroy(ctypes.c_void_p(comms[i]))
        ctxs[i].close()
