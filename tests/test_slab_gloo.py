"""N > 1 path on CPU: the slab-decomposed periodic step (rustpde_b200/slab.py) with world_size 2 and 3 over
gloo, the kernels running in the CUDA emulation (tests/cuemu), against the oracle.  Uneven splits on purpose
(17 Fourier modes and 33 grid columns over 2 or 3 ranks)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nx, ny, steps, out, pw=None):
    if pw is not None:
        os.environ["RUSTPDE_B200_PW"] = pw  # force the row-sweep (1) or tile (0) per-mode kernels on the slabs
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "cuemu")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import build_emu
        import parity_cases as pc
        from rustpde_b200 import _ffi
        from rustpde_b200.slab import Navier2DSlab, split

        lib = _ffi.Lib(build_emu.build())
        n, o = pc.make_navier_pair(lib, True, nx, ny, 1e5, 1.0, 0.01)
        s = Navier2DSlab(n)
        assert s.world == world and s.ksz == split(nx // 2 + 1, world)[0]
        s.update(steps)
        # before the gather only this rank's rows are current
        mine = n.ux.vhat[s.k0:s.k0 + s.mkl].copy()
        s.gather_state()
        assert np.array_equal(n.ux.vhat[s.k0:s.k0 + s.mkl], mine)
        for _ in range(steps):
            o.update()
        err = pc.navier_field_errors(n, o)
        diag = [abs(a - b) / max(abs(b), 1e-300) for a, b in zip(n.eval(), pc.oracle_diag(o))]
        ok = max(err.values()) <= 1e-10 and max(diag) <= 1e-9 and abs(n.time - o.time) < 1e-12
        out[rank] = (ok, err, diag)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nx,ny,pw", [(2, 32, 33, None), (3, 32, 33, None), (2, 128, 129, None), (3, 64, 65, "1")])
def test_slab_periodic_gloo(world, nx, ny, pw):
    sys.path.insert(0, os.path.join(ROOT, "tests", "cuemu"))
    import build_emu  # build once in the parent so the workers do not race on the shared object
    build_emu.build()
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29600 + world * 7 + (nx % 97)
    mp.spawn(_worker, args=(world, port, nx, ny, 3, out, pw), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        ok, err, diag = out[r]
        assert ok, (r, err, diag)


def test_split():
    sys.path.insert(0, ROOT)
    from rustpde_b200.slab import split
    assert split(4097, 8) == ([513] + [512] * 7, [0, 513, 1025, 1537, 2049, 2561, 3073, 3585])
    assert split(8193, 8)[0] == [1025] + [1024] * 7
    assert sum(split(17, 3)[0]) == 17 and split(17, 3)[1] == [0, 6, 12]


@pytest.mark.parametrize("world,nx,ny", [(2, 32, 33), (3, 32, 33), (3, 128, 129)])
def test_fused_transposes_virtual_ranks(emu, world, nx, ny):
    """The fused-transpose entry points (rp_navier_slab_phase{1,2}_p2p) with the peers played by buffers of ONE
    process: every virtual rank scatters into every other rank's buffer exactly like the NVLink stores do."""
    import ctypes as C
    import torch
    import parity_cases as pc
    from rustpde_b200.slab import split

    n, o = pc.make_navier_pair(emu, True, nx, ny, 1e5, 1.0, 0.01)
    mk = nx // 2 + 1
    ksz, koff = split(mk, world)
    jsz, joff = split(ny, world)
    xin = [[torch.zeros(mk * jsz[q] * 2, dtype=torch.float64) for _ in range(6)] for q in range(world)]
    s3 = [[torch.zeros(ksz[q] * ny * 2, dtype=torch.float64) for _ in range(3)] for q in range(world)]
    work = [torch.zeros(8 * nx * jsz[q], dtype=torch.float64) for q in range(world)]
    peers1 = (C.c_void_p * (6 * world))(*[xin[q][a].data_ptr() for a in range(6) for q in range(world)])
    peers2 = (C.c_void_p * (3 * world))(*[s3[q][f].data_ptr() for f in range(3) for q in range(world)])
    cj = (C.c_int * (world + 1))(*(joff + [ny]))
    ck = (C.c_int * (world + 1))(*(koff + [mk]))
    for _ in range(3):
        for p in range(world):
            emu.call("rp_navier_slab_phase1_p2p", n._h, koff[p], ksz[p], world, cj, peers1)
        for p in range(world):
            in6 = (C.c_void_p * 6)(*[t.data_ptr() for t in xin[p]])
            emu.call("rp_navier_slab_phase2_p2p", n._h, joff[p], jsz[p], in6, C.c_void_p(work[p].data_ptr()), world, ck, peers2)
        for p in range(world):
            in3 = (C.c_void_p * 3)(*[t.data_ptr() for t in s3[p]])
            emu.call("rp_navier_slab_phase3", n._h, koff[p], ksz[p], in3)
        o.update()
    err = pc.navier_field_errors(n, o)
    assert max(err.values()) <= 1e-10, err
