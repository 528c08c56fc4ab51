"""fe_project_b200 -- B200-native (sm_100a, FP64 CUDA) dynamics hot path of FE-Project's SCALE-DG.

Only what the hot path needs lives here: `csrc/` (CUDA kernels + the C ABI of include/fedg.h, built
to `libfedg.so`), the host-side mirror of the reference's driver interface (`dyncore`), and the
set-up code that produces the arrays a Fortran caller would hand over (`element`, `mesh`, `initcond`).
"""
from .element import HexElement, LineElement  # noqa: F401
from .mesh import LocalMeshCube  # noqa: F401

__all__ = ["HexElement", "LineElement", "LocalMeshCube"]
