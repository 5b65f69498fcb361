"""Slab-decomposed periodic Navier2D over the GPUs of one node (BASELINE config 5).

One process per GPU (torch.distributed, NCCL over NVLink; gloo with the CPU emulation in the tests).
The spectral arrays `[nx/2+1, .]` are split over the Fourier modes kx (row slabs): every per-mode
operation -- y transforms, stencils, rhs assembly, the per-mode Helmholtz / Poisson solves
(hholtz.rs:156-197, poisson.rs:131-149), projection -- is local to a rank.  The x FFT and the
physical-space products (conv_term.rs:41) need full x lines, so they run on column slabs of physical
y; the two layouts are connected by an all-to-all:

    phase 1 (kx slab)   B_y S_y and B_y D_y S_y of ux, uy, temp                   rp_navier_slab_phase1
    all-to-all          6 complex arrays  [mk_loc, ny] -> [mk, ny_loc]
    phase 2 (y slab)    c2r along x, products, r2c along x + dealias              rp_navier_slab_phase2
    all-to-all          3 complex arrays  [mk, ny_loc] -> [mk_loc, ny]
    phase 3 (kx slab)   forward DCT-y + dealias, rhs, solves, projection          rp_navier_slab_phase3

The y transform is applied before the x transform in the backward direction (the reference does x
first, space2.rs:346-356; the operators commute, the results differ by rounding only).

Every rank holds a full `Navier2D` (set-up, initial conditions and diagnostics reuse the single-GPU
code; 21 GB at 8192x8193) but steps only its own rows; `gather_state()` makes the full arrays
consistent again on every rank before diagnostics are evaluated.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def split(n, parts):
    """Sizes and offsets of `parts` nearly equal consecutive chunks of range(n) (remainder to the first ranks)."""
    base, rem = divmod(n, parts)
    sizes = [base + (1 if p < rem else 0) for p in range(parts)]
    offs = [sum(sizes[:p]) for p in range(parts)]
    return sizes, offs


class Navier2DSlab:
    """transport = "collective": dense phase outputs + all_to_all_single (NCCL / gloo).
    transport = "p2p": fused transposes -- the phase-1 / phase-2 kernels store every element straight into the
    buffer of the rank that owns it over NVLink peer memory (buffers shared with CUDA IPC); the phases are
    separated by a stream-ordered one-element all-reduce instead of nine all-to-alls and their pack copies."""

    def __init__(self, nav, group=None, transport="collective"):
        if not nav.periodic:
            raise ValueError("the slab decomposition is defined for Navier2D.new_periodic (Fourier x Chebyshev)")
        self.nav = nav
        self.lib = nav._lib
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.nx, self.ny = nav.nx, nav.ny
        self.mk = self.nx // 2 + 1
        self.ksz, self.koff = split(self.mk, self.world)
        self.jsz, self.joff = split(self.ny, self.world)
        if min(self.ksz) < 1 or min(self.jsz) < 1:
            raise ValueError("more ranks than Fourier modes / grid columns")
        self.k0, self.mkl = self.koff[self.rank], self.ksz[self.rank]
        self.j0, self.nyl = self.joff[self.rank], self.jsz[self.rank]
        dev = "cpu" if self.lib.emulated else torch.device("cuda", torch.cuda.current_device())
        f64 = dict(dtype=torch.float64, device=dev)
        mk, ny, nx, mkl, nyl = self.mk, self.ny, self.nx, self.mkl, self.nyl
        # exchange buffers (interleaved complex = trailing dimension 2)
        self.s1 = [torch.zeros(mkl * ny * 2, **f64) for _ in range(6)]    # phase-1 output   [mkl, ny]
        self.x_in = [torch.zeros(mk * nyl * 2, **f64) for _ in range(6)]  # after exchange   [mk, nyl]
        self.x_out = [torch.zeros(mk * nyl * 2, **f64) for _ in range(3)]  # phase-2 output  [mk, nyl]
        self.s3 = [torch.zeros(mkl * ny * 2, **f64) for _ in range(3)]    # phase-3 input    [mkl, ny]
        self.pack = torch.zeros(mkl * ny * 2, **f64)                      # send / receive staging
        self.work = torch.zeros(8 * nx * nyl, **f64)
        self._p = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        self.transport = transport if self.world > 1 else "collective"
        if self.transport == "p2p":
            self._setup_p2p(f64)
        # our kernels per slab step: phase 1 one launch (3 fields), phase 2 three (c2r value + d/dx, c2r d/dy, products + r2c),
        # phase 3 five (forward DCT-y, Helmholtz x2 + x1, divergence + Poisson, projection)
        # (with the per-mode row sweeps of fast_pw.cu -- rows >= 448 unless RUSTPDE_B200_PW overrides -- the three Helmholtz
        # solves are one launch)
        import os
        pw = os.environ.get("RUSTPDE_B200_PW", "")
        self.launches_per_step = 8 if (pw[:1] == "1" or (pw[:1] != "0" and mkl >= 448)) else 9
        self.fences_per_step = 2 if self.transport == "p2p" else 0
        # NVLink egress of this rank per step: 6 arrays rows->cols, 3 arrays cols->rows (16 B per complex element)
        self.bytes_exchanged_per_step = 16 * (6 * mkl * (ny - nyl) + 3 * nyl * (mk - mkl))

    # -- exchanges ---------------------------------------------------------------------------------
    def _rows_to_cols(self, src, dst):
        """[mkl, ny] complex on every rank -> [mk, nyl]: column blocks are packed, row blocks arrive in rank order."""
        if self.world == 1:
            dst.copy_(src)
            return
        s = src.view(self.mkl, self.ny, 2)
        send_split, off = [], 0
        for q in range(self.world):
            cnt = self.mkl * self.jsz[q] * 2
            self.pack[off:off + cnt].view(self.mkl, self.jsz[q], 2).copy_(s[:, self.joff[q]:self.joff[q] + self.jsz[q], :])
            send_split.append(cnt)
            off += cnt
        recv_split = [self.ksz[q] * self.nyl * 2 for q in range(self.world)]
        dist.all_to_all_single(dst, self.pack, recv_split, send_split, group=self.group)

    def _cols_to_rows(self, src, dst):
        """[mk, nyl] complex -> [mkl, ny]: row blocks leave as they are, column blocks are unpacked on arrival."""
        if self.world == 1:
            dst.copy_(src)
            return
        send_split = [self.ksz[q] * self.nyl * 2 for q in range(self.world)]
        recv_split = [self.mkl * self.jsz[q] * 2 for q in range(self.world)]
        dist.all_to_all_single(self.pack, src, recv_split, send_split, group=self.group)
        d = dst.view(self.mkl, self.ny, 2)
        off = 0
        for q in range(self.world):
            cnt = recv_split[q]
            d[:, self.joff[q]:self.joff[q] + self.jsz[q], :].copy_(self.pack[off:off + cnt].view(self.mkl, self.jsz[q], 2))
            off += cnt

    # -- fused transposes over peer memory ---------------------------------------------------------
    def _setup_p2p(self, f64):
        lib, W = self.lib, self.world
        mk, ny = self.mk, self.ny
        nbytes = lambda q: 16 * (6 * mk * self.jsz[q] + 3 * self.ksz[q] * ny)
        base = C.c_void_p()
        lib.call("rp_dev_alloc", nbytes(self.rank), C.byref(base))
        handle = C.create_string_buffer(64)
        lib.call("rp_ipc_export", base, handle)
        handles = [None] * W
        dist.all_gather_object(handles, handle.raw, group=self.group)
        self._peer_base, self._opened = [], []
        for q in range(W):
            if q == self.rank:
                self._peer_base.append(base.value)
            else:
                ptr = C.c_void_p()
                lib.call("rp_ipc_open", C.create_string_buffer(handles[q], 64), C.byref(ptr))
                self._peer_base.append(ptr.value)
                self._opened.append(ptr)
        self._own = base
        xin = lambda q, a: self._peer_base[q] + 16 * a * mk * self.jsz[q]
        s3 = lambda q, f: self._peer_base[q] + 16 * (6 * mk * self.jsz[q] + f * self.ksz[q] * ny)
        self._peers1 = (C.c_void_p * (6 * W))(*[xin(q, a) for a in range(6) for q in range(W)])
        self._peers2 = (C.c_void_p * (3 * W))(*[s3(q, f) for f in range(3) for q in range(W)])
        self._in6 = (C.c_void_p * 6)(*[xin(self.rank, a) for a in range(6)])
        self._in3 = (C.c_void_p * 3)(*[s3(self.rank, f) for f in range(3)])
        self._joff = (C.c_int * (W + 1))(*(self.joff + [ny]))
        self._koff = (C.c_int * (W + 1))(*(self.koff + [mk]))
        self._flag = torch.zeros(1, **f64)
        dist.barrier(group=self.group)

    def _fence(self):
        # stream-ordered: completes on this rank only after every rank's preceding kernels have finished
        dist.all_reduce(self._flag, group=self.group)

    def close(self):
        if getattr(self, "_own", None) is not None:
            self.sync()
            dist.barrier(group=self.group)
            for p in self._opened:
                self.lib.call("rp_ipc_close", p)
            self.lib.call("rp_dev_free", self._own)
            self._own = None

    # -- time stepping -------------------------------------------------------------------------------
    def update(self, nsteps=1):
        lib, h = self.lib, self.nav._h
        if self.transport == "p2p":
            for _ in range(int(nsteps)):
                lib.call("rp_navier_slab_phase1_p2p", h, self.k0, self.mkl, self.world, self._joff, self._peers1)
                self._fence()
                lib.call("rp_navier_slab_phase2_p2p", h, self.j0, self.nyl, self._in6, C.c_void_p(self.work.data_ptr()), self.world, self._koff, self._peers2)
                self._fence()
                lib.call("rp_navier_slab_phase3", h, self.k0, self.mkl, self._in3)
            return
        for _ in range(int(nsteps)):
            lib.call("rp_navier_slab_phase1", h, self.k0, self.mkl, self._p(self.s1))
            for a in range(6):
                self._rows_to_cols(self.s1[a], self.x_in[a])
            lib.call("rp_navier_slab_phase2", h, self.j0, self.nyl, self._p(self.x_in), C.c_void_p(self.work.data_ptr()), self._p(self.x_out))
            for a in range(3):
                self._cols_to_rows(self.x_out[a], self.s3[a])
            lib.call("rp_navier_slab_phase3", h, self.k0, self.mkl, self._p(self.s3))

    def profile_step(self, reps=5):
        """Device time (ms, CUDA events on the stepping stream) of the parts of one p2p slab step: phase 1, fence,
        phase 2, fence, phase 3 -- averaged over `reps` steps.  Advances the solution like update(reps)."""
        if self.transport != "p2p" or self.lib.emulated:
            return None
        lib, h = self.lib, self.nav._h
        names = ["phase1", "fence1", "phase2", "fence2", "phase3"]
        tot = dict.fromkeys(names, 0.0)
        for _ in range(int(reps)):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
            ev[0].record()
            lib.call("rp_navier_slab_phase1_p2p", h, self.k0, self.mkl, self.world, self._joff, self._peers1)
            ev[1].record()
            self._fence()
            ev[2].record()
            lib.call("rp_navier_slab_phase2_p2p", h, self.j0, self.nyl, self._in6, C.c_void_p(self.work.data_ptr()), self.world, self._koff, self._peers2)
            ev[3].record()
            self._fence()
            ev[4].record()
            lib.call("rp_navier_slab_phase3", h, self.k0, self.mkl, self._in3)
            ev[5].record()
            torch.cuda.synchronize()
            for i, n in enumerate(names):
                tot[n] += ev[i].elapsed_time(ev[i + 1])
        return {n: tot[n] / reps for n in names}

    def sync(self):
        self.nav.sync()
        if not self.lib.emulated:
            torch.cuda.synchronize()

    def gather_state(self):
        """Make temp / ux / uy / pres vhat of the full per-rank model consistent (each rank owns rows [k0, k0+mkl)).
        The row slabs travel as tensors through all_gather (NCCL on GPUs, gloo in the CPU tests), padded to the
        largest slab."""
        self.sync()
        if self.world == 1:
            return
        dev = "cpu" if self.lib.emulated else torch.device("cuda", torch.cuda.current_device())
        kmax = max(self.ksz)
        for f in (self.nav.temp, self.nav.ux, self.nav.uy, self.nav.pres[0], self.nav.pres[1]):
            ncol = f.shape_spectral[1]
            mine = torch.zeros(kmax, ncol, 2, dtype=torch.float64, device=dev)
            rows = f.vhat_rows(self.k0, self.mkl)
            mine[: self.mkl].copy_(torch.from_numpy(rows.view(np.float64).reshape(self.mkl, ncol, 2)))
            parts = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(parts, mine, group=self.group)
            for q in range(self.world):
                if q == self.rank:
                    continue
                blk = parts[q][: self.ksz[q]].cpu().numpy().reshape(self.ksz[q], ncol * 2).view(np.complex128)
                f.set_vhat_rows(self.koff[q], blk)
