"""Slab-decomposed periodic step on 2 real GPUs (NCCL all-to-all over NVLink) against the oracle.
Skipped when fewer than 2 GPUs are visible (the driver's `pytest -m gpu` box has one)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nx, ny, steps, transport, out):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import parity_cases as pc
        from rustpde_b200 import _ffi
        from rustpde_b200.slab import Navier2DSlab

        lib = _ffi.product_lib(rank)
        n, o = pc.make_navier_pair(lib, True, nx, ny, 1e6, 1.0, 2e-3)
        s = Navier2DSlab(n, transport=transport)
        assert s.transport == transport
        s.update(steps)
        s.gather_state()
        for _ in range(steps):
            o.update()
        err = pc.navier_field_errors(n, o)
        out[rank] = (max(err.values()) <= 1e-9, err)
        s.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["collective", "p2p"])
@pytest.mark.parametrize("nx,ny,steps", [(64, 65, 10), (512, 513, 5), (2048, 129, 3)])
def test_slab_periodic_nccl(nx, ny, steps, transport):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29650 + nx % 89 + (7 if transport == "p2p" else 0), nx, ny, steps, transport, out), nprocs=2, join=True)
    for r in range(2):
        ok, err = out[r]
        assert ok, (r, err)
