"""CPU oracle: numpy/scipy restatement of rustpde's Navier2D hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rustpde_b200/`` may import this
package; it is imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` only, and only
as the checker / reported baseline - never as the product path.

Every function cites the reference file:line it restates (paths relative to
the reference checkout).  The reference is pure Rust and cannot be built in
this image (no cargo/rustc), so this is a behavioural restatement in the
reference's op order and memory layout (row-major ``[nx, ny]``).

Pinning: the restatement is checked against every known-answer vector the
reference's own unit tests and doc-tests hold for this path
(``tests/test_oracle_kat.py``; list in SURVEY.md section 8c).  Third-party
arithmetic that is not vendored in the reference tree:
  * ndrustfft 0.2.1 / rustfft 6.0.1 / realfft 2.0.1 / rustdct 0.6.0
    (Cargo.lock:531,800,860,869) -> scipy.fft.dct(type=1) unnormalised and
    numpy.fft.rfft/irfft; conventions pinned by the reference doc-tests
    (ortho.rs:210-220,260-270; r2c.rs:130-147,177-194).
  * ndarray-linalg 0.14.1 / lax 0.2.0 / openblas-src 0.10.4 (dgeev, dgetri)
    -> scipy.linalg.eig / inv (same LAPACK routines, OpenBLAS 0.3.31 here).
    Eigenvector normalisation is unpinned (and irrelevant: only lam, Q,
    Q^-1 C^-1 jointly matter).
End-to-end Navier2D values are NOT pinned by the reference (smoke doc-test
only, navier.rs:139-152): our deterministic-IC runs are the only pin.
"""

from .funspace import (  # noqa: F401
    Chebyshev,
    CompositeChebyshev,
    FourierR2c,
    Space2,
    chebyshev,
    cheb_dirichlet,
    cheb_neumann,
    cheb_dirichlet_bc,
    cheb_neumann_bc,
    fourier_r2c,
)
from .field import Field2  # noqa: F401
from .solver import (  # noqa: F401
    Fdma,
    FdmaTensor,
    Hholtz,
    HholtzAdi,
    MatVecFdma,
    Poisson,
    eig,
    inv,
)
from .navier import Navier2D, integrate  # noqa: F401
from .navier_adjoint import Navier2DAdjoint  # noqa: F401
from .solid_masks import Statistics, solid_cylinder_inner, solid_porosity, solid_roughness_sinusoid  # noqa: F401
