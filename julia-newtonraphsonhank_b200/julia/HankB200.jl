# HankB200.jl — Julia host glue over libhankb200.so (include/hankb200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: Julia is not installed in the build image (SURVEY.md §0.7),
# so this file is written against the C header and the reference's call signatures and is checked
# only by review; the same entry points are exercised from Python (hankb200/) by the GPU tests.
#
# Usage inside the reference repository (after its own `include`s, see INTEGRATION.md):
#
#     include("HankB200.jl"); using .HankB200
#     mod = build_model_from_yaml("KrusellSmith.yaml")
#     ss, _ = get_SteadyStates(mod)
#     blk = HankB200.HouseholdBlock(mod, ss, ss)            # one hank_ctx on GPU 0
#     J̅  = HankB200.jacobian(blk, x_ss, exog_ss)           # brute-force JVP columns (config 3)
#     x  = HankB200.NewtonRaphsonHANK(x_0, J̅, exog_paths, blk)
#
# The methods below keep the reference's names and argument meaning:
#   BackwardIteration(xVec_endog, exog_paths, model, ss_end)   BackwardIteration.jl:46-49
#   ForwardIteration(policy_seqs, model, ss_initial)           ForwardIteration.jl:253-255
#   JVP(func, primal, tangent)                                 GeneralStructures.jl:542-544
#   NewtonRaphsonHANK(x_0, J̅, exog_paths, mod, ss0, ssT; ε)    NewtonRaphson.jl:27-33
module HankB200

using LinearAlgebra, SparseArrays
import ForwardDiff

const LIB = get(ENV, "HANKB200_LIB", joinpath(@__DIR__, "..", "lib", "libhankb200.so"))

struct HankError <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::HankError) = print(io, "hankb200 status $(e.code): $(e.msg)")

mutable struct HouseholdBlock
    ctx::Ptr{Cvoid}
    n_a::Int; n_e::Int; T::Int; P::Int
    function HouseholdBlock(ctx, n_a, n_e, T)
        b = new(ctx, n_a, n_e, T, T - 1)
        finalizer(x -> ccall((:hank_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ctx), b)
        b
    end
end

last_error(ctx) = unsafe_string(ccall((:hank_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
check(b::HouseholdBlock, rc) = rc == 0 ? nothing : throw(HankError(rc, last_error(b.ctx)))

"""
    HouseholdBlock(model::SequenceModel, ss_initial, ss_ending; device = 0)

Creates the device context from the fields `ValueFunction` reads on every call
(KrusellSmith.jl:44-52) and uploads the two steady-state records' `.value` / `.D`
(SteadyState.jl:21-27).  `Π` is passed as stored (column-major, row-stochastic).
"""
function HouseholdBlock(model, ss_initial, ss_ending; device::Integer = 0)
    w = model.heterogeneity.wealth; pr = model.heterogeneity.productivity
    p = model.params; T = model.compspec.T
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:hank_ctx_create, LIB), Cint,
               (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64),
               ref, device, w.n, pr.n, T, collect(Float64, w.grid), collect(Float64, pr.grid),
               Matrix{Float64}(pr.transition), p.β, p.γ, p.borrow_cons)
    if rc != 0
        msg = ref[] == C_NULL ? "context allocation failed" : last_error(ref[])
        ref[] == C_NULL || ccall((:hank_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), ref[])
        throw(HankError(rc, msg))
    end
    b = HouseholdBlock(ref[], w.n, pr.n, T)
    check(b, ccall((:hank_set_terminal, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.ctx, Matrix{Float64}(ss_ending.value)))
    check(b, ccall((:hank_set_initial_dist, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.ctx, Vector{Float64}(ss_initial.D)))
    check(b, ccall((:hank_ks_configure, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Float64), b.ctx, p.α, p.δ, ss_initial.vars.KS))
    b
end

# ── Dual packing: Vector{Dual{T,Float64,N}} is bit-compatible with an (N+1) x len Float64 matrix,
#    value first (ForwardDiff.jl/src/dual.jl:14-16, partials.jl:1-3) ────────────────────────────
unpack(v::AbstractVector{Float64}) = (Vector{Float64}(v), Matrix{Float64}(undef, length(v), 0))
function unpack(v::AbstractVector{ForwardDiff.Dual{Tg,Float64,N}}) where {Tg,N}
    raw = reinterpret(reshape, Float64, collect(v))          # (N+1) x len
    (Vector{Float64}(raw[1, :]), Matrix{Float64}(permutedims(raw[2:end, :])))   # len x N
end
repack(::Type{Float64}, val, part) = val
function repack(::Type{ForwardDiff.Dual{Tg,Float64,N}}, val::AbstractVector, part::AbstractMatrix) where {Tg,N}
    [ForwardDiff.Dual{Tg}(val[i], ForwardDiff.Partials(ntuple(k -> part[i, k], Val(N)))) for i in eachindex(val)]
end

"""Device handle returned by `BackwardIteration`; `Array(h)` materialises the `Vector{Matrix}` of
policies for callers that need host matrices (SteadyStateJacobian.jl:226-229)."""
struct DevicePolicies{TF}
    blk::HouseholdBlock
    K::Int
end
function Base.Array(h::DevicePolicies{Float64})
    b = h.blk
    map(1:b.P) do t
        out = Matrix{Float64}(undef, b.n_a, b.n_e)
        check(b, ccall((:hank_get_policy, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}), b.ctx, t, 0, out))
        out
    end
end

"""
    BackwardIteration(xVec_endog, exog_paths, blk::HouseholdBlock, ss_end) -> (KD = DevicePolicies,)

Same meaning as BackwardIteration.jl:46-116.  Only `r` and `w` enter the KS household block
(KrusellSmith.jl:53-54): rows 3 and 4 of `reshape(x, n_endog, T-1)`.
"""
function BackwardIteration(xVec_endog::AbstractVector{TF}, exog_paths::NamedTuple, b::HouseholdBlock, ss_end = nothing) where {TF}
    xv, xp = unpack(xVec_endog)
    K = size(xp, 2)
    X = reshape(xv, 4, b.P)
    r = X[3, :]; w = X[4, :]
    dr = Matrix{Float64}(undef, b.P, K); dw = similar(dr)
    for k in 1:K
        D = reshape(view(xp, :, k), 4, b.P)
        dr[:, k] .= D[3, :]; dw[:, k] .= D[4, :]
    end
    check(b, ccall((:hank_backward, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}),
                   b.ctx, r, w, K, dr, dw))
    (KD = DevicePolicies{TF}(b, K),)
end

"""
    ForwardIteration(policy_seqs, blk, ss_initial) -> (KD = Vector,)

ForwardIteration.jl:253-311 on the device-resident policies left by `BackwardIteration`.
"""
function ForwardIteration(policy_seqs::NamedTuple{(:KD,),Tuple{DevicePolicies{TF}}}, b::HouseholdBlock, ss_initial = nothing) where {TF}
    K = policy_seqs.KD.K
    KD = Vector{Float64}(undef, b.P); dKD = Matrix{Float64}(undef, b.P, K)
    check(b, ccall((:hank_forward, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), b.ctx, KD, dKD))
    (KD = repack(TF, KD, dKD),)
end

"""fullFunction(x) of NewtonRaphson.jl:77-83 on the device (sweeps + residuals); keeps the linearisation."""
function fullFunction(b::HouseholdBlock, x::Vector{Float64}, Z::Vector{Float64})
    F = Vector{Float64}(undef, length(x))
    check(b, ccall((:hank_ks_linearize, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), b.ctx, x, Z, F))
    F
end

"""
    JVP(blk, primal, tangent[s]; Z) -> SparseVector / Matrix

`JVP(func, primal, tangent)` of GeneralStructures.jl:542-550 for `func = fullFunction`; a matrix of
tangents (n x K) rides as K lanes in one pass — the batched entry ForwardDiff's chunking (<= 12
lanes, prelude.jl:1-11) cannot express.
"""
function JVP(b::HouseholdBlock, primal::Vector{Float64}, tangents::AbstractVecOrMat{Float64}; Z::Vector{Float64})
    fullFunction(b, primal, Z)
    V = tangents isa AbstractVector ? reshape(Vector{Float64}(tangents), :, 1) : Matrix{Float64}(tangents)
    JV = similar(V)
    check(b, ccall((:hank_ks_jvp, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}), b.ctx, size(V, 2), V, JV))
    tangents isa AbstractVector ? sparse(vec(JV)) : JV
end

"""`(F(x), J(x)·V)` in one call (hank_ks_fjvp): the seed upload overlaps the primal backward sweep."""
function fjvp(b::HouseholdBlock, x::Vector{Float64}, Z::Vector{Float64}, V::Matrix{Float64})
    F = Vector{Float64}(undef, length(x)); JV = similar(V)
    check(b, ccall((:hank_ks_fjvp, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   b.ctx, x, Z, size(V, 2), V, F, JV))
    (F, JV)
end

"""Columns `cols` (a range) of the sequence-space Jacobian at `x`: directJVPJacobian (SteadyState.jl:296-320)
generalised to any column range; the Y / KS columns skip the household sweeps."""
function jacobian(b::HouseholdBlock, x::Vector{Float64}, Z::Vector{Float64}, cols::UnitRange{Int} = 1:length(x))
    fullFunction(b, x, Z)
    J = Matrix{Float64}(undef, length(x), length(cols))
    check(b, ccall((:hank_ks_jacobian_columns, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}), b.ctx, first(cols), last(cols) + 1, J))
    J
end

"""
    NewtonRaphsonHANK(x_0, J̅, exog_paths, blk; ε = 1e-9, solver = :lu)

NewtonRaphson.jl:27-114 on the device.  `solver = :gmres` reproduces the reference's restarted
GMRES(20) preconditioner solve; `:lu` factorises J̅ once; `:lu_batched` additionally assembles J(x)
once per outer iteration from batched lanes.
"""
function NewtonRaphsonHANK(x_0::Vector{Float64}, J̅::AbstractMatrix, exog_paths::NamedTuple, b::HouseholdBlock;
                           ε = 1e-9, solver::Symbol = :lu)
    n = length(x_0)
    x = Vector{Float64}(undef, n); stats = zeros(8); inner = zeros(Cint, 100)
    code = Dict(:gmres => 0, :lu => 1, :lu_batched => 2)[solver]
    check(b, ccall((:hank_newton_solve, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}),
                   b.ctx, Matrix{Float64}(J̅), x_0, Vector{Float64}(exog_paths.Z), ε, 1e-9, code, x, stats, inner))
    println("Newton: $(Int(stats[1])) outer iterations, $(Int(stats[2])) JVPs, ‖y‖ = $(stats[4])")
    x
end

"""value_fn plug-in: one EGM step on the device (KrusellSmith.jl:43-83), Float64 or Dual value_next."""
function ValueFunction(value_next::AbstractMatrix{TV}, xVals::AbstractVector{TX}, b::HouseholdBlock) where {TV,TX}
    TF = promote_type(TV, TX)
    vv, vp = unpack(vec(value_next)); xv, xp = unpack(collect(xVals))
    K = max(size(vp, 2), size(xp, 2))
    G = b.n_a * b.n_e
    dv = size(vp, 2) == K ? vp : zeros(G, K)
    dr = size(xp, 2) == K ? xp[3, :] : zeros(K); dw = size(xp, 2) == K ? xp[4, :] : zeros(K)
    val = Vector{Float64}(undef, G); pol = similar(val); dval = Matrix{Float64}(undef, G, K); dpol = similar(dval)
    check(b, ccall((:hank_egm_step, LIB), Cint,
                   (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   b.ctx, vv, K == 0 ? C_NULL : dv, xv[3], xv[4], K, dr, dw, val, pol, dval, dpol))
    sh(v, d) = reshape(repack(TF, v, d), b.n_a, b.n_e)
    (Value = sh(val, dval), KD = sh(pol, dpol))
end

end # module
