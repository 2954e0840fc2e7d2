"""cannoles_b200: B200-native `linsolve` backend for the KKT factor/solve path of CaNNOLeS.jl.

Public surface (mirrors reference/src/solver_types.jl):
    B200Struct(N, rows, cols, vals)      -- the `linsolve=:b200` LinearSolverStruct
    try_to_factorize / solve_ldl / get_vals
and the restated host-side caller (``cannoles``, ``CaNNOLeSSolver``, ``solve``).
The CUDA library (csrc/libcannoles_b200.so) is loaded lazily by ``linsolve``; there is no CPU
fallback: constructing a ``B200Struct`` without the library or without a GPU raises.
"""
from .solver import (CaNNOLeSSolver, ExecutionStats, ParamCaNNOLeS, cannoles, newton_system,  # noqa: F401
                     prepare_newton_system, register_linsolve, solve)
from .linsolve import B200Error, B200Struct  # noqa: E402,F401

register_linsolve("b200", B200Struct)
