"""RPSNAP1 snapshot container (csrc/snapshot.cu): reader, writer and HDF5 converter.

`Navier2D.write()` / `read()` (navier.rs:956-1014) store the reference's datasets under the reference's names --
`temp|ux|uy|pres/{v, vhat | vhat_re, vhat_im}`, `x, dx, y, dy`, `time, ra, pr, nu, kappa` -- in a flat
little-endian container, because neither libhdf5 nor h5py exists in the build image:

    magic "RPSNAP1\\0" | u32 count | count x { u32 name_len, name, u32 ndim, u64 dims[ndim], f64 data[prod(dims)] }

Where h5py is installed, `python -m rustpde_b200.snapshot to-h5 in.rpsnap out.h5` writes the very file the reference's
`write()` produces (hdf5-interface/src/lib.rs:145-277: one f64 dataset per name, groups created on demand), and
`from-h5` turns a reference snapshot into a restart file for `rp_navier_read_snapshot`.  The h5py legs cannot be
exercised in this image (stated in DESIGN.md); the container itself is covered by tests/test_snapshot.py.
"""
import struct
import sys

import numpy as np

MAGIC = b"RPSNAP1\0"


def read_snapshot(path):
    """dict name -> ndarray (0-d for scalars)."""
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s: not an RPSNAP1 file" % path)
        (count,) = struct.unpack("<I", f.read(4))
        for _ in range(count):
            (nl,) = struct.unpack("<I", f.read(4))
            name = f.read(nl).decode()
            (nd,) = struct.unpack("<I", f.read(4))
            dims = struct.unpack("<%dQ" % nd, f.read(8 * nd)) if nd else ()
            n = int(np.prod(dims)) if nd else 1
            out[name] = np.frombuffer(f.read(8 * n), dtype="<f8").reshape(dims).copy()
    return out


def write_snapshot(path, data):
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(data)))
        for name in sorted(data):
            a = np.asarray(data[name], dtype="<f8", order="C")
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)) + nb + struct.pack("<I", a.ndim))
            if a.ndim:
                f.write(struct.pack("<%dQ" % a.ndim, *a.shape))
            f.write(a.tobytes())


def to_hdf5(src, dst):
    import h5py  # not available in the build image
    with h5py.File(dst, "w") as h:
        for name, a in read_snapshot(src).items():
            h.create_dataset(name, data=a)  # intermediate groups are created on demand


def from_hdf5(src, dst):
    import h5py
    out = {}
    with h5py.File(src, "r") as h:
        h.visititems(lambda name, obj: out.__setitem__(name, np.array(obj)) if isinstance(obj, h5py.Dataset) else None)
    write_snapshot(dst, out)


if __name__ == "__main__":
    if len(sys.argv) != 4 or sys.argv[1] not in ("to-h5", "from-h5"):
        raise SystemExit("usage: python -m rustpde_b200.snapshot to-h5|from-h5 SRC DST")
    (to_hdf5 if sys.argv[1] == "to-h5" else from_hdf5)(sys.argv[2], sys.argv[3])
