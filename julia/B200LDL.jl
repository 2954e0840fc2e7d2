# B200LDL.jl -- Julia glue for `linsolve = :b200` (NOT executed offline: Julia is absent from the
# build image; kept deliberately thin -- every numerical step is a ccall into
# libcannoles_b200.so, whose C ABI is include/cannoles_b200.h).
#
# It mirrors, verb for verb, the two existing adapters of reference/src/solver_types.jl
# (MA57Struct :17-43, LDLFactStruct :45-98) and is `include`d from src/CaNNOLeS.jl next to
# `include("solver_types.jl")`.  The two edits to src/CaNNOLeS.jl are listed in INTEGRATION.md.
#
# Limitation (state it in the docstring at src/CaNNOLeS.jl:121): Float64 only; the reference's
# LDLFactorizations path is generic in T (test/runtests.jl:102-113 exercises Float16..BigFloat).

const libb200 = get(ENV, "CANNOLES_B200_LIB", "libcannoles_b200")

const B2_ORDER_ND = Cint(0)
const B2_ORDER_AMD = Cint(3)

b200_error() = unsafe_string(ccall((:b2_last_error, libb200), Cstring, ()))
b200_check(rc::Cint) = rc == 0 ? nothing : error("libcannoles_b200: " * b200_error())

"""Opaque device factor; what `LDLT.factor` is for this backend (read at src/CaNNOLeS.jl:1049)."""
mutable struct B200Factor
  handle::Ptr{Cvoid}
  refine_steps::Cint
  function B200Factor(handle, refine_steps)
    F = new(handle, refine_steps)
    finalizer(F) do f
      f.handle == C_NULL || ccall((:b2_free, libb200), Cint, (Ptr{Cvoid},), f.handle)
      f.handle = C_NULL
    end
    return F
  end
end

mutable struct B200Struct <: LinearSolverStruct
  rows::Vector{Int64}
  cols::Vector{Int64}
  vals::Vector{Float64}          # THE live buffer (aliased by solver.vals, src/CaNNOLeS.jl:328)
  N::Int
  ordering::Cint
  device::Cint
  factor::B200Factor             # handle == C_NULL until the first try_to_factorize
  nuploads::Int                  # to recognise the rho retries of newton_system!
  inertia::NTuple{3, Int64}
end

# ctor with the reference's signature X(N, rows, cols, vals) (src/solver_types.jl:21, :61).
# The block sizes nvar/nequ/ncon are not known here (the reference passes them only to
# try_to_factorize), so the symbolic analysis is done lazily at the first factorization.
function B200Struct(N, rows::Vector{Int64}, cols::Vector{Int64}, vals::Vector{Float64};
                    ordering = B2_ORDER_ND, device = 0, refine_steps = 1)
  return B200Struct(rows, cols, vals, N, ordering, device, B200Factor(C_NULL, refine_steps), 0, (0, 0, 0))
end

get_vals(LDLT::B200Struct) = LDLT.vals          # src/solver_types.jl:25, :67

function b200_analyze!(LDLT::B200Struct, nvar, nequ, ncon)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  b200_check(ccall((:b2_analyze, libb200), Cint,
    (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Int64, Cint, Ptr{Int64}, Cint, Ref{Ptr{Cvoid}}),
    LDLT.N, length(LDLT.vals), LDLT.rows, LDLT.cols, nvar, nequ, ncon, LDLT.ordering, C_NULL,
    LDLT.device, h))
  LDLT.factor.handle = h[]
  # vals has a stable address for the life of the solver (src/CaNNOLeS.jl:276-279): pin it once
  ccall((:b2_register_host, libb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
        h[], LDLT.vals, sizeof(LDLT.vals))
  return LDLT
end

# try_to_factorize (src/solver_types.jl:32, :79): Bool, never throws on numerical failure.
function try_to_factorize(LDLT::B200Struct, vals::AbstractVector{Float64}, nvar::Integer,
                          nequ::Integer, ncon::Integer, eig_tol::Float64)
  LDLT.factor.handle == C_NULL && b200_analyze!(LDLT, nvar, nequ, ncon)
  h = LDLT.factor.handle
  npos, nzero, nneg, brk = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Cint}(0)
  ρ = nvar > 0 ? vals[end] : 0.0
  # newton_system! (src/CaNNOLeS.jl:1029-1043) re-enters with ONLY the trailing rho segment
  # changed to a non-zero constant: shift the diagonal on the device, upload nothing.
  retry = LDLT.nuploads > 0 && nvar > 0 && ρ != 0 && vals === LDLT.vals && vals[end - nvar + 1] == ρ
  rc = Cint(-1)
  if retry
    rc = ccall((:b2_refactorize_shift, libb200), Cint,
      (Ptr{Cvoid}, Float64, Float64, Float64, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cint}),
      h, ρ, NaN, eig_tol, npos, nzero, nneg, brk)
  end
  if rc != 0
    b200_check(ccall((:b2_factorize, libb200), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Float64, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cint}),
      h, vals, eig_tol, npos, nzero, nneg, brk))
    LDLT.nuploads += 1
  end
  LDLT.inertia = (npos[], nzero[], nneg[])
  return npos[] == nvar && nzero[] == 0           # src/solver_types.jl:96
end

# solve_ldl! (src/solver_types.jl:26, :69): d = -(K \ rhs); returns true like both reference backends.
function solve_ldl!(rhs::AbstractVector{Float64}, factor::B200Factor, d::AbstractVector{Float64})
  b200_check(ccall((:b2_solve, libb200), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Cint, Ptr{Float64}),
    factor.handle, rhs, d, 1, factor.refine_steps, C_NULL))
  return true
end
