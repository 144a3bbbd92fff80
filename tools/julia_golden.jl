# julia_golden.jl — writes golden vectors of the hot path FROM THE REAL REFERENCE (vasudeva-ram/Julia-NewtonRaphsonHANK).
#
# Julia is not installed in this repository's build image, so this script has never been executed here; it is the
# recipe that PINS the oracle (DESIGN.md "Oracle — parity unpinned").  Run it once on a machine with Julia >= 1.11 and
# the reference's Manifest instantiated:
#
#     cd /path/to/Julia-NewtonRaphsonHANK
#     julia --project=. /path/to/repo/tools/julia_golden.jl /tmp/julia_golden [n_a n_e T]
#     python /path/to/repo/tools/julia_golden_to_npz.py /tmp/julia_golden /path/to/repo/tests/golden/julia_ks.npz
#
# tests/test_julia_golden.py then checks the oracle (CPU) and the CUDA path (GPU) against that file and stops
# reporting "parity unpinned".  Only stdlib I/O is used (raw little-endian arrays + a text manifest), so no package
# outside the reference's own Manifest is needed.
#
# The driver sequence is the corrected one of SURVEY.md Appendix B (RunMain.jl does not run as shipped).
using LinearAlgebra, SparseArrays, Random, Printf
import ForwardDiff, IterativeSolvers

for f in ("GeneralStructures.jl", "ModelParser.jl", "KrusellSmith.jl", "BackwardIteration.jl", "ForwardIteration.jl",
          "Aggregation.jl", "SteadyState.jl", "SteadyStateJacobian.jl", "NewtonRaphson.jl")
    include(joinpath(pwd(), f))                      # same files, same order as test_SteadyState.jl:11-18 (+ NewtonRaphson.jl)
end

outdir = length(ARGS) >= 1 ? ARGS[1] : "julia_golden"
mkpath(outdir)
manifest = IOBuffer()
function put(name::String, a)
    arr = a isa Number ? [a] : collect(a)
    T = eltype(arr) <: Integer ? Int64 : Float64
    open(joinpath(outdir, name * ".bin"), "w") do io
        write(io, convert(Array{T}, arr))            # column-major, little-endian
    end
    println(manifest, name, " ", T == Int64 ? "i8" : "f8", " ", join(size(arr), "x"))
end

yaml = "KrusellSmith.yaml"
if length(ARGS) >= 4                                 # optional other grid: rewrite n / T in a temporary copy of the YAML
    n_a, n_e, Tn = parse.(Int, ARGS[2:4])
    txt = read(yaml, String)
    txt = replace(txt, r"(name:\s*wealth[\s\S]*?\bn:\s*)\d+" => SubstitutionString("\\g<1>$(n_a)"))
    txt = replace(txt, r"(name:\s*productivity[\s\S]*?\bn:\s*)\d+" => SubstitutionString("\\g<1>$(n_e)"))
    txt = replace(txt, r"(name:\s*T\s*\n\s*value:\s*)\d+" => SubstitutionString("\\g<1>$(Tn)"))
    yaml = joinpath(outdir, "model.yaml"); write(yaml, txt)
end

mod = build_model_from_yaml(yaml)
ss, _ = get_SteadyStates(mod)
T = mod.compspec.T; P = T - 1
endog = vars_of_type(mod, :endogenous); allk = var_names(mod)
w = mod.heterogeneity.wealth; pr = mod.heterogeneity.productivity
n_a, n_e = w.n, pr.n
p = mod.params

put("T", T); put("n_a", n_a); put("n_e", n_e)
put("grid", w.grid); put("z", pr.grid); put("Pi", pr.transition)          # Pi[e, e2], column-major
put("beta", p.β); put("gamma", p.γ); put("alpha", p.α); put("delta", p.δ); put("borrow_cons", p.borrow_cons)
put("ss_vars", Float64[ss.vars[k] for k in allk])                          # Y, KS, r, w, KD, Z
put("ss_value", ss.value); put("ss_D", ss.D); put("ss_policy", ss.policies.KD)

# ── one EGM step (KrusellSmith.jl:43-83) from the steady-state value, and its derivative w.r.t. r and w through
#    ForwardDiff: pins Interpolations' gridded-linear / Flat rule, its knot derivative and DiffRules' max tie rule
xss = Float64[ss.vars[k] for k in allk]
step = mod.value_fn(ss.value, xss, mod)
put("egm_value", step.Value); put("egm_policy", step.KD)
jr = findfirst(==(:r), allk); jw = findfirst(==(:w), allk)
for (nm, j) in (("r", jr), ("w", jw))
    f(s) = (x = convert(Vector{typeof(s)}, xss); x[j] += s; r = mod.value_fn(ss.value, x, mod); vcat(vec(r.Value), vec(r.KD)))
    d = ForwardDiff.derivative(f, 0.0)
    put("egm_dvalue_d" * nm, reshape(d[1:n_a*n_e], n_a, n_e)); put("egm_dpolicy_d" * nm, reshape(d[n_a*n_e+1:end], n_a, n_e))
end

# ── lottery of the steady-state policy (ForwardIteration.jl:37-78): CSC pattern and weights of Λ_endog
Λe = make_endogenous_transition(ss.policies.KD, w, n_e)
put("lottery_colptr", Λe.colptr); put("lottery_rowval", Λe.rowval); put("lottery_nzval", Λe.nzval)

# ── the full function on a perturbed path: policies, aggregates, residuals
x0 = repeat(Float64[ss.vars[k] for k in endog], P)
X = reshape(copy(x0), length(endog), P)
ir = findfirst(==(:r), endog); iw = findfirst(==(:w), endog)
for t in 1:P
    X[ir, t] *= 1 + 0.05 * 0.9^t
    X[iw, t] *= 1 + 0.02 * 0.9^t
end
x1 = vec(X)
exog = (Z = 1.0 .+ [0.8^t for t in 1:P],)                                  # RunMain.jl:50-51
pol = BackwardIteration(x1, exog, mod, ss)
agg = ForwardIteration(pol, mod, ss)
put("x1", x1); put("Z", exog.Z)
for t in (1, max(1, P ÷ 2), P)
    put("policy_t$(t)", pol.KD[t])
end
put("KD_path", agg.KD)
fullF(x) = Residuals(assemble_full_xMat(x, ForwardIteration(BackwardIteration(x, exog, mod, ss), mod, ss), exog, mod, ss, ss), mod)
put("F_x1", fullF(x1)); put("F_x0", fullF(x0))

# ── JVP columns of test_SteadyState.jl:186-224 (Z = 1 paths), plus two dense directions at x1
exog_ss = (Z = fill(Float64(ss.vars.Z), P),)
pipe(x) = Residuals(assemble_full_xMat(x, ForwardIteration(BackwardIteration(x, exog_ss, mod, ss), mod, ss), exog_ss, mod, ss, ss), mod)
n = length(x0)
Random.seed!(42)
cols = [1, 2, rand(3:n-2, 3)..., n - 1, n]
put("jvp_cols", cols)
put("jvp_columns", reduce(hcat, [Vector(JVP(pipe, x0, sparsevec([c], [1.0], n))) for c in cols]))
Random.seed!(7)
V = randn(n, 2)
put("jvp_V", V)
put("jvp_dense", reduce(hcat, [Vector(JVP(fullF, x1, V[:, k])) for k in 1:2]))

# ── the steady-state Jacobian the reference builds (only a preconditioner: SURVEY.md Appendix B/C)
J̅ = getSteadyStateJacobian(ss, mod)
put("Jbar", Matrix(J̅))

# ── Newton path with the reference's own functions, instrumented copy of NewtonRaphson.jl:27-114 (same calls, same
#    order; `log = true` only adds the iteration history), then the un-instrumented call as a cross-check
function newton_logged(x_0, J̅, exog, mod, ss0, ssT; ε = 1e-9)
    F(x) = Residuals(assemble_full_xMat(x, ForwardIteration(BackwardIteration(x, exog, mod, ssT), mod, ss0), exog, mod, ss0, ssT), mod)
    x = x_0; y = x_0; i = 1
    inner = Int[]; gm = Int[]
    while (ε < norm(y)) && (i < 100)
        y_old = ones(length(y)); M = ones(length(y)); R = ones(length(y))
        Fx = F(x); k = 0
        while ε < norm(y - y_old)
            Λxy = JVP(F, x, y)
            _, h1 = IterativeSolvers.gmres!(R, J̅, Fx - Λxy; log = true)
            _, h2 = IterativeSolvers.gmres!(M, J̅, Λxy; log = true)
            push!(gm, h1.iters)
            y_old = y; y = y_old + 0.5 * R; k += 1
        end
        push!(inner, k)
        x = x - y; i += 1
        @printf("outer %d: inner %d, ‖y‖ = %.3e\n", i - 1, k, norm(y))
    end
    x, inner, gm
end
xN, inner, gm = newton_logged(x0, J̅, exog, mod, ss, ss)
put("newton_x", xN); put("newton_inner", inner); put("newton_gmres_iters", gm)
xN2 = NewtonRaphsonHANK(x0, J̅, exog, mod, ss, ss)
put("newton_x_uninstrumented", xN2)

write(joinpath(outdir, "manifest.txt"), String(take!(manifest)))
println("golden vectors written to ", outdir)
