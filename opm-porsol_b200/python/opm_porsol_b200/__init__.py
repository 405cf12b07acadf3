"""B200-native explicit saturation transport (opm-porsol's opm/porsol/euler hot path).

Python is plumbing only: this package loads the C-ABI shared library
(opm-porsol_b200/lib/libeuler_b200.so, hand-written sm_100a CUDA) with ctypes and mirrors the
reference's ``EulerUpstream`` interface (``init`` / ``initObj`` / ``transportSolve``).
There is no CPU fallback: a missing library or device raises.
"""
from .binding import EulerUpstream, EulerB200Error, lib_path, load_library, resolve_boundary, Report  # noqa: F401
from . import synth  # noqa: F401
