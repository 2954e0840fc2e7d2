# B200LDL.jl -- Julia glue for `linsolve = :b200` (NOT executed offline: Julia is absent from the
# build image; kept deliberately thin -- every numerical step is a ccall into
# libcannoles_b200.so, whose C ABI is include/cannoles_b200.h).
#
# It mirrors, verb for verb, the two existing adapters of reference/src/solver_types.jl
# (MA57Struct :17-43, LDLFactStruct :45-98) and is `include`d from src/CaNNOLeS.jl next to
# `include("solver_types.jl")`.  The two edits to src/CaNNOLeS.jl are listed in INTEGRATION.md.
#
# Limitation (state it in the docstring at src/CaNNOLeS.jl:121): the device path is Float64 only.
# The reference's LDLFactorizations path is generic in T (test/runtests.jl:102-113 exercises
# Float16..BigFloat): for any other element type the constructor below warns and hands back an
# `LDLFactStruct`, exactly as src/CaNNOLeS.jl:317-320 downgrades `:ma57` when HSL is missing.

const libb200 = get(ENV, "CANNOLES_B200_LIB", "libcannoles_b200")

const B2_ORDER_ND = Cint(0)
const B2_ORDER_AMD = Cint(3)

b200_error() = unsafe_string(ccall((:b2_last_error, libb200), Cstring, ()))
b200_check(rc::Cint) = rc == 0 ? nothing : error("libcannoles_b200: " * b200_error())

"""
Opaque device factor; what `LDLT.factor` is for this backend (read at src/CaNNOLeS.jl:1049).

It OWNS a reference to every host buffer it has pinned with `b2_register_host` (`vals`, and `rhs` /
`d` from the first `solve_ldl!` on), so that the garbage collector cannot free a still-registered
buffer before the finalizer has unregistered it (`b2_free` unregisters, then releases the device).
"""
mutable struct B200Factor
  handle::Ptr{Cvoid}
  refine_steps::Cint
  pinned::Vector{Vector{Float64}}     # buffers registered with the CUDA driver: kept alive here
  function B200Factor(handle, refine_steps)
    F = new(handle, refine_steps, Vector{Float64}[])
    finalizer(close!, F)
    return F
  end
end

"""Release the device factor and unpin the host buffers now (idempotent; also the finalizer)."""
function close!(F::B200Factor)
  if F.handle != C_NULL
    ccall((:b2_free, libb200), Cint, (Ptr{Cvoid},), F.handle)   # unregisters what it registered
    F.handle = C_NULL
  end
  empty!(F.pinned)
  return nothing
end

# pin a caller-owned buffer once (addresses are stable: src/CaNNOLeS.jl:241-243, 276-279); a failed
# registration only costs PCIe speed (pageable copies), so it is reported, not fatal
function b200_pin!(F::B200Factor, buf::Vector{Float64})
  for b in F.pinned
    b === buf && return true
  end
  rc = ccall((:b2_register_host, libb200), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t),
             F.handle, buf, sizeof(buf))
  if rc == 0
    push!(F.pinned, buf)
    return true
  end
  @warn "libcannoles_b200: could not pin a host buffer, copies will be pageable" error = b200_error()
  return false
end

mutable struct B200Struct <: LinearSolverStruct
  rows::Vector{Int64}
  cols::Vector{Int64}
  vals::Vector{Float64}          # THE live buffer (aliased by solver.vals, src/CaNNOLeS.jl:328)
  N::Int
  ordering::Cint
  device::Cint
  factor::B200Factor             # handle == C_NULL until the first try_to_factorize
  nuploads::Int                  # full uploads so far
  nshifts::Int                   # rho retries answered by the device-side shift
  inertia::NTuple{3, Int64}
  # out-parameters of the ccalls, allocated once (allocation test, test/runtests.jl:28-36)
  npos::Base.RefValue{Int64}
  nzero::Base.RefValue{Int64}
  nneg::Base.RefValue{Int64}
  brk::Base.RefValue{Cint}
  held::Base.RefValue{Cint}
end

# ctor with the reference's signature X(N, rows, cols, vals) (src/solver_types.jl:21, :61).
# The block sizes nvar/nequ/ncon are not known here (the reference passes them only to
# try_to_factorize), so the symbolic analysis is done lazily at the first factorization.
function B200Struct(N, rows::Vector{Int64}, cols::Vector{Int64}, vals::Vector{Float64};
                    ordering = B2_ORDER_ND, device = 0, refine_steps = 1)
  return B200Struct(rows, cols, vals, N, ordering, device, B200Factor(C_NULL, refine_steps), 0, 0,
                    (0, 0, 0), Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Cint}(0), Ref{Cint}(0))
end

# Any other element type: the device path is Float64 only -- warn and fall through to the generic
# LDLFactorizations adapter, like the HSL guard of src/CaNNOLeS.jl:317-320 does for :ma57.
function B200Struct(N, rows::Vector{Ti}, cols::Vector{Ti}, vals::Vector{T}; kwargs...) where {T, Ti <: Integer}
  @warn "linsolve = :b200 is Float64 / Int64 only; using :ldlfactorizations for eltype $T / $Ti"
  return LDLFactStruct(N, rows, cols, vals)
end

get_vals(LDLT::B200Struct) = LDLT.vals          # src/solver_types.jl:25, :67

function b200_analyze!(LDLT::B200Struct, nvar, nequ, ncon)
  h = Ref{Ptr{Cvoid}}(C_NULL)
  b200_check(ccall((:b2_analyze, libb200), Cint,
    (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Int64, Cint, Ptr{Int64}, Cint, Ref{Ptr{Cvoid}}),
    LDLT.N, length(LDLT.vals), LDLT.rows, LDLT.cols, nvar, nequ, ncon, LDLT.ordering, C_NULL,
    LDLT.device, h))
  LDLT.factor.handle = h[]
  # vals has a stable address for the life of the solver (src/CaNNOLeS.jl:276-279): pin it once;
  # the factor keeps it alive until it has been unregistered
  b200_pin!(LDLT.factor, LDLT.vals)
  return LDLT
end

# try_to_factorize (src/solver_types.jl:32, :79): Bool, never throws on numerical failure.
function try_to_factorize(LDLT::B200Struct, vals::AbstractVector{Float64}, nvar::Integer,
                          nequ::Integer, ncon::Integer, eig_tol::Float64)
  LDLT.factor.handle == C_NULL && b200_analyze!(LDLT, nvar, nequ, ncon)
  h = LDLT.factor.handle
  ρ = nvar > 0 ? vals[end] : 0.0
  # newton_system! (src/CaNNOLeS.jl:1029-1043) re-enters with ONLY the trailing rho segment changed
  # to a non-zero constant.  A call that looks like that goes through b2_factorize_retry: the matrix
  # shifted on the device is factorized at once while vals is uploaded on a copy stream and compared
  # with the previous upload on the device; if the caller changed anything else the result is the
  # full factorization of what it passed -- never a stale one.
  looks_like_retry = LDLT.nuploads > 0 && nvar > 0 && ρ != 0 && vals[end - nvar + 1] == ρ
  if looks_like_retry
    b200_check(ccall((:b2_factorize_retry, libb200), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Float64, Float64, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cint}, Ref{Cint}),
      h, vals, ρ, eig_tol, LDLT.npos, LDLT.nzero, LDLT.nneg, LDLT.brk, LDLT.held))
    LDLT.held[] != 0 ? (LDLT.nshifts += 1) : (LDLT.nuploads += 1)
  else
    b200_check(ccall((:b2_factorize, libb200), Cint,
      (Ptr{Cvoid}, Ptr{Float64}, Float64, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cint}),
      h, vals, eig_tol, LDLT.npos, LDLT.nzero, LDLT.nneg, LDLT.brk))
    LDLT.nuploads += 1
  end
  LDLT.inertia = (LDLT.npos[], LDLT.nzero[], LDLT.nneg[])
  return LDLT.npos[] == nvar && LDLT.nzero[] == 0           # src/solver_types.jl:96
end

# solve_ldl! (src/solver_types.jl:26, :69): d = -(K \ rhs); returns true like both reference backends.
# rhs and d are allocated once by the solver (src/CaNNOLeS.jl:241-243): pinned at the first call.
function solve_ldl!(rhs::Vector{Float64}, factor::B200Factor, d::Vector{Float64})
  b200_pin!(factor, rhs)
  b200_pin!(factor, d)
  b200_check(ccall((:b2_solve, libb200), Cint,
    (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Cint, Ptr{Float64}),
    factor.handle, rhs, d, 1, factor.refine_steps, C_NULL))
  return true
end

# ------------------------------------------------------------------------------------------
# Batches of small dense instances that share one KKT pattern (multi-start / per-sample
# estimation; BASELINE.json config 5): the WHOLE iteration of solve! (src/CaNNOLeS.jl:418-864) runs
# on the device, one CTA per instance (csrc/nls_kernels.cuh).  The reference has no batch entry
# point; this is the one a maintainer would put next to `cannoles`.  Never executed in this
# repository's CI (no Julia offline); the C side is exercised through the same symbols from Python
# (cannoles_b200/batched_nls.py, tests/test_gpu_nls.py).
#   F(x) = A x + 0.1 sin(B x) - y,   c(x) = C x + 0.05 (x .* x)[1:ncon] - e = 0
#   A, B :: Array{Float64,3} of size (m, n, batch), C :: (ncon, n, batch)   (column-major = the C side's layout)
struct B2NLSParams                      # b2_nls_params_t
  eig_tol::Float64; delta_min::Float64; kappa_dec::Float64; kappa_inc::Float64; kappa_largeinc::Float64
  rho0::Float64; rho_max::Float64; rho_min::Float64; gamma_A::Float64
  atol::Float64; rtol::Float64; Fatol::Float64; Frtol::Float64; delta_dec::Float64; cgls_tol::Float64
  max_iter::Int32; max_eval::Int32; max_inner::Int32; always_accept_extrapolation::Int32
  use_initial_multiplier::Int32; reserved::Int32
end
struct B2DenseNLS                       # b2_dense_nls_t
  n::Int64; m::Int64; ncon::Int64; shared_model::Int64
  At::Ptr{Float64}; Bt::Ptr{Float64}; Ct::Ptr{Float64}; y::Ptr{Float64}; e::Ptr{Float64}
  x0::Ptr{Float64}; y0::Ptr{Float64}
end

"""
    cannoles_batch(A, B, C, y, e, x0; rows, cols) -> records

`rows`, `cols`: the COO layout of the Newton system of ONE instance (src/CaNNOLeS.jl:281-315, exact
residual Hessian).  `records[:, b]` = status, iter, nfact, nlinsolve, nbk, neval_residual, neval_cons,
objective, ||c||, ||dual||_inf, rho, delta, x (n), lambda (ncon) of instance `b` (status: 1 first_order,
2 small_residual, 3 stalled, 4 exception, 5 max_eval, 7 max_iter; 8-10 = the errors solve! throws).
"""
function cannoles_batch(A::Array{Float64,3}, B::Array{Float64,3}, C::Array{Float64,3}, y::Matrix{Float64},
                        e::Matrix{Float64}, x0::Matrix{Float64}; rows::Vector{Int64}, cols::Vector{Int64},
                        device::Integer = 0)
  m, n, batch = size(A)
  ncon = size(C, 1)
  N = n + m + ncon
  h = Ref{Ptr{Cvoid}}(C_NULL)
  b200_check(ccall((:b2b_analyze, libb200), Cint,
    (Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int64, Int64, Int64, Int64, Cint, Ptr{Int64}, Cint, Ptr{Ptr{Cvoid}}),
    N, length(rows), rows, cols, n, m, ncon, batch, 3, C_NULL, device, h))      # 3 = B2_ORDER_AMD
  prm = Ref{B2NLSParams}()
  ccall((:b2_nls_default_params, libb200), Cvoid, (Ptr{B2NLSParams},), prm)
  reclen = ccall((:b2b_nls_record_len, libb200), Int64, (Ptr{Cvoid},), h[])
  rec = Matrix{Float64}(undef, reclen, batch)
  GC.@preserve A B C y e x0 begin
    md = Ref(B2DenseNLS(n, m, ncon, 0, pointer(A), pointer(B), pointer(C), pointer(y), pointer(e),
                        pointer(x0), C_NULL))
    rc = ccall((:b2b_nls_dense_solve, libb200), Cint,
      (Ptr{Cvoid}, Ptr{B2DenseNLS}, Int64, Ptr{B2NLSParams}, Ptr{Float64}, Int64), h[], md, batch, prm, rec, 0)
  end
  ccall((:b2b_free, libb200), Cint, (Ptr{Cvoid},), h[])
  b200_check(rc)
  return rec
end
