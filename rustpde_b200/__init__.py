"""rustpde_b200 -- B200-native Navier2D time step of preiter93/rustpde.

Hand-written sm_100a CUDA kernels behind a C ABI (include/rustpde_b200.h,
librustpde_b200.so) with a host-side mirror of the reference API (api.py).
Importing the package does not touch the GPU; the first object created loads
the CUDA library and raises if it (or a GPU) is missing -- no CPU fallback.
"""
from .api import (  # noqa: F401
    Base,
    Field2,
    Hholtz,
    HholtzAdi,
    Navier2D,
    Navier2DAdjoint,
    Poisson,
    RustpdeError,
    Space2,
    cheb_dirichlet,
    cheb_dirichlet_bc,
    cheb_neumann,
    cheb_neumann_bc,
    chebyshev,
    fourier_r2c,
    integrate,
)
from .solid_masks import (  # noqa: F401
    Statistics,
    solid_cylinder_inner,
    solid_porosity,
    solid_porosity_interpolate,
    solid_roughness_sinusoid,
)

__all__ = [
    "Base", "Field2", "Hholtz", "HholtzAdi", "Navier2D", "Navier2DAdjoint", "Poisson", "RustpdeError", "Space2",
    "cheb_dirichlet", "cheb_dirichlet_bc", "cheb_neumann", "cheb_neumann_bc", "chebyshev", "fourier_r2c", "integrate",
    "Statistics", "solid_cylinder_inner", "solid_porosity", "solid_porosity_interpolate", "solid_roughness_sinusoid",
]
