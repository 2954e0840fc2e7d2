# gen_ref_vectors.jl -- TEST INFRASTRUCTURE, to be run by a maintainer on any machine that HAS Julia
# (the build image has none, so the oracle of this repository is "parity unpinned" at the
# factorization boundary until this script has been run once and its outputs committed).
#
#   julia --project=/path/to/CaNNOLeS.jl oracle/gen_ref_vectors.jl [repo root]
#
# It drives the REFERENCE's own `:ldlfactorizations` adapter (reference/src/solver_types.jl:45-98:
# `LDLFactStruct(N, rows, cols, vals)` -> `try_to_factorize` -> `solve_ldl!`) on the committed
# golden inputs `tests/golden/ref_inputs/*.txt` (written by tests/golden/make_golden.py) and
# writes, next to them, `tests/golden/ref_<name>.txt` with what the factorization boundary holds:
#   P        the permutation ldl_analyze chose (SuiteSparse AMD through AMD.jl), 1-based
#   d        factor.d, in pivot order (what the inertia loop :90-95 reads)
#   inertia  #(d > eps), #(|d| <= eps), #(d < -eps)  and the Bool try_to_factorize returned
#   x        the output of solve_ldl!(rhs, factor, d)  (= -K \ rhs)
# and `tests/golden/ref_cannoles.txt` with iter / nfact / nlinsolve / status / solution of
# `cannoles(nls; linsolve = :ldlfactorizations)` on MGH01CON (config 1) and on a 2000-variable slice
# of config 2.  tests/test_ref_vectors.py consumes these files when present (and says "unpinned"
# when they are absent): it runs the C oracle on the SAME permutation P and compares d, inertia, x,
# and checks whether the oracle's own AMD reproduces P.
#
# Only documented API of CaNNOLeS.jl / LDLFactorizations.jl / ADNLPModels.jl is used; nothing in
# this repository is needed besides the input files.

using LinearAlgebra, SparseArrays, Printf
using CaNNOLeS
using ADNLPModels, NLPModels

root = length(ARGS) >= 1 ? ARGS[1] : normpath(joinpath(@__DIR__, ".."))
indir = joinpath(root, "tests", "golden", "ref_inputs")
outdir = joinpath(root, "tests", "golden")

"input format: line 1 `N nnz nvar nequ ncon`, then nnz lines `row col val` (1-based, %.17g), then N lines rhs"
function read_input(path)
  lines = readlines(path)
  N, nnz, nvar, nequ, ncon = parse.(Int, split(lines[1]))
  rows = Vector{Int}(undef, nnz); cols = Vector{Int}(undef, nnz); vals = Vector{Float64}(undef, nnz)
  for t in 1:nnz
    f = split(lines[1 + t])
    rows[t] = parse(Int, f[1]); cols[t] = parse(Int, f[2]); vals[t] = parse(Float64, f[3])
  end
  rhs = [parse(Float64, lines[1 + nnz + i]) for i in 1:N]
  return N, nvar, nequ, ncon, rows, cols, vals, rhs
end

function write_vec(io, name, v)
  println(io, name, " ", length(v))
  for x in v
    x isa Integer ? println(io, x) : @printf(io, "%.17g\n", x)
  end
end

for f in sort(filter(n -> endswith(n, ".txt"), readdir(indir)))
  name = splitext(f)[1]
  N, nvar, nequ, ncon, rows, cols, vals, rhs = read_input(joinpath(indir, f))
  eig_tol = eps(Float64)
  # exactly what CaNNOLeSSolver does at src/CaNNOLeS.jl:327 and newton_system! at :1023, :1049
  LDLT = CaNNOLeS.LDLFactStruct(N, rows, cols, vals)
  ok = CaNNOLeS.try_to_factorize(LDLT, vals, nvar, nequ, ncon, eig_tol)
  dvec = copy(LDLT.factor.d)
  x = zeros(N)
  ok && CaNNOLeS.solve_ldl!(rhs, LDLT.factor, x)
  pos = count(>(eig_tol), dvec); zer = count(v -> abs(v) <= eig_tol, dvec)
  open(joinpath(outdir, "ref_" * name * ".txt"), "w") do io
    println(io, "# written by oracle/gen_ref_vectors.jl; Julia ", VERSION)
    println(io, "ok ", ok ? 1 : 0)
    println(io, "inertia ", pos, " ", zer, " ", N - pos - zer)
    write_vec(io, "P", collect(LDLT.factor.P))
    write_vec(io, "d", dvec)
    write_vec(io, "x", x)
  end
  println("wrote ref_", name, ".txt  ok=", ok)
end

# ---- end-to-end counters of the loop (no reference test pins them: SURVEY 8(c)) --------------
include(joinpath(dirname(pathof(CaNNOLeS)), "..", "test", "mgh01con.jl"))   # MGH01CON (config 1)

function c2_slice(n)
  # config 2 at n variables (SURVEY App. F): F_larger of test/runtests.jl:59 + c_j = x_{2j-1} + x_{2j} - 2
  F(x) = [[10 * (x[i + 1] - x[i]^2) for i = 1:(n - 1)]; [x[i] - 1 for i = 1:(n - 1)]]
  c(x) = [x[2j - 1] + x[2j] - 2 for j = 1:(n ÷ 2)]
  x0 = [0.9 + 0.01 * sin(1.0 * i) for i = 1:n]     # deterministic, no RNG stream to match
  return ADNLSModel(F, x0, 2 * (n - 1), c, zeros(n ÷ 2), zeros(n ÷ 2))
end

open(joinpath(outdir, "ref_cannoles.txt"), "w") do io
  println(io, "# written by oracle/gen_ref_vectors.jl; Julia ", VERSION)
  for (name, nls, method) in (("mgh01con", MGH01CON(), :Newton), ("c2_slice_2000", c2_slice(2000), :Newton_noFHess))
    stats = cannoles(nls, linsolve = :ldlfactorizations, method = method, max_time = 3600.0)
    println(io, "case ", name)
    println(io, "status ", stats.status)
    println(io, "iter ", stats.iter)
    println(io, "nfact ", stats.solver_specific[:nfact])
    println(io, "nlinsolve ", stats.solver_specific[:nlinsolve])
    @printf(io, "objective %.17g\n", stats.objective)
    @printf(io, "primal_feas %.17g\n", stats.primal_feas)
    @printf(io, "dual_feas %.17g\n", stats.dual_feas)
    write_vec(io, "solution", stats.solution)
  end
end
println("wrote ref_cannoles.txt")
