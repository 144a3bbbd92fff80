# HankB200.jl — Julia host glue over libhankb200.so (include/hankb200.h): a DROP-IN for the hot path of
# vasudeva-ram/Julia-NewtonRaphsonHANK.
#
# NOT EXECUTED IN THIS REPOSITORY: Julia is not installed in the build image (SURVEY.md §0.7), so this file is written
# against the C header and the reference's call signatures and is checked by review only; the same C entry points
# are exercised from Python (hankb200/) by the GPU tests.  INTEGRATION.md §2 lists every deviation from the reference.
#
# Usage: `include` it in `Main` AFTER the reference's own files (RunMain.jl:1-9 / test_SteadyState.jl:11-18):
#
#     include("GeneralStructures.jl"); …; include("NewtonRaphson.jl")
#     include("/path/to/julia-newtonraphsonhank_b200/julia/HankB200.jl")
#
# Nothing else changes in the caller.  The file adds METHODS to the reference's own generic functions, each with the
# reference's positional signature and a first argument (or steady-state argument) typed one step more specifically, so
# that Julia's dispatch picks them for `Float64` / `ForwardDiff.Dual{…,Float64,N}` inputs and the reference's
# untyped methods stay available as the fallback (`invoke`) for anything the device path does not cover:
#
#   BackwardIteration(xVec_endog, exog_paths::NamedTuple, model::SequenceModel, ss_end)      BackwardIteration.jl:46-49
#   ForwardIteration(policy_seqs::NamedTuple, model::SequenceModel, ss_initial)              ForwardIteration.jl:253-255
#   NewtonRaphsonHANK(x_0, J̅::SparseMatrixCSC, exog_paths, mod, ss_initial, ss_ending; ε)    NewtonRaphson.jl:27-33
#   get_xVals(asm::SSAssembler, p_vec)                                                       SteadyState.jl:111-154
#   ValueFunction(value_next, xVals, model)  (the `value_fn` plug-in)                        KrusellSmith.jl:43-83
#
# `JVP(func, primal, tangent)` (GeneralStructures.jl:542-550) and `y_Iteration.fullFunction` (NewtonRaphson.jl:77-83)
# are NOT redefined: they call `BackwardIteration` / `ForwardIteration` on `Vector{Dual{Tag,Float64,1}}` and therefore
# reach the device through the methods below unchanged.  `JVP(func, primal, tangents::AbstractMatrix)` is an
# additional batched method (K lanes per pass; ForwardDiff itself never carries more than 12, prelude.jl:1-11).
#
# The device path covers the Krusell-Smith household block (`model.value_fn === ValueFunction`, heterogeneity
# `(wealth, productivity)`, one heterogeneous variable `KD`); `HankB200.ENABLED[] = false` or any other model
# falls back to the reference's Julia code.

module HankB200Lib
const LIB = get(ENV, "HANKB200_LIB", joinpath(@__DIR__, "..", "lib", "libhankb200.so"))
struct HankError <: Exception
    code::Cint
    msg::String
end
Base.showerror(io::IO, e::HankError) = print(io, "hankb200 status $(e.code): $(e.msg)")
last_error(ctx) = unsafe_string(ccall((:hank_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
# status -> exception, like the reference's error(...) / DomainError / ArgumentError (find_ss's line search
# catches them, SteadyState.jl:199-207)
check(ctx, rc) = rc == 0 ? nothing : throw(HankError(rc, last_error(ctx)))
end # module HankB200Lib

module HankB200
const ENABLED = Ref(true)          # false: every call goes to the reference's Julia code
const DEVICE = Ref(0)              # CUDA device of the contexts created from here on
const NEWTON_SOLVER = Ref(:gmres)  # :gmres = the reference's restarted GMRES(20) preconditioner solve (IterativeSolvers
                                   # defaults); :lu = exact solve with a cached LU of J̅; :lu_batched = as :lu with J(x)
                                   # assembled once per outer iteration from batched unit-seed lanes
end

using LinearAlgebra, SparseArrays
import ForwardDiff
import .HankB200Lib: LIB, HankError, check

# ── one device context per SequenceModel (created on first use, destroyed by the finalizer) ──────────────────────
mutable struct HankBlock
    ctx::Ptr{Cvoid}
    n_a::Int; n_e::Int; T::Int; P::Int
    ir::Int; iw::Int; n_endog::Int          # rows of r and w in reshape(x, n_endog, P)
    terminal_id::UInt; initial_id::UInt     # objectid of the steady-state records last uploaded
    K::Int                                  # lanes of the last hank_backward
end
const _BLOCKS = IdDict{Any,HankBlock}()

_is_ks(model::SequenceModel) = HankB200.ENABLED[] && model.value_fn === ValueFunction &&
    keys(model.heterogeneity) == (:wealth, :productivity) && vars_of_type(model, :heterogeneous) == (:KD,)

function _block(model::SequenceModel)
    get!(_BLOCKS, model) do
        w = model.heterogeneity.wealth; pr = model.heterogeneity.productivity
        p = model.params; T = model.compspec.T
        ref = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:hank_ctx_create, LIB), Cint,
                   (Ref{Ptr{Cvoid}}, Cint, Cint, Cint, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Float64),
                   ref, HankB200.DEVICE[], w.n, pr.n, T, collect(Float64, w.grid), collect(Float64, pr.grid),
                   Matrix{Float64}(pr.transition), p.β, p.γ, p.borrow_cons)
        if rc != 0
            msg = ref[] == C_NULL ? "context allocation failed" : HankB200Lib.last_error(ref[])
            ref[] == C_NULL || ccall((:hank_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), ref[])
            throw(HankError(rc, msg))
        end
        endog = vars_of_type(model, :endogenous)          # r and w are looked up BY NAME (KrusellSmith.jl:53-54)
        b = HankBlock(ref[], w.n, pr.n, T, T - 1, findfirst(==(:r), endog), findfirst(==(:w), endog), length(endog), 0, 0, 0)
        finalizer(x -> ccall((:hank_ctx_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ctx), b)
        b
    end
end
function _set_terminal!(b::HankBlock, ss_end)
    objectid(ss_end) == b.terminal_id && return
    check(b.ctx, ccall((:hank_set_terminal, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.ctx, Matrix{Float64}(ss_end.value)))
    b.terminal_id = objectid(ss_end)
end
function _set_initial!(b::HankBlock, model, ss_initial)
    objectid(ss_initial) == b.initial_id && return
    check(b.ctx, ccall((:hank_set_initial_dist, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), b.ctx, Vector{Float64}(ss_initial.D)))
    p = model.params   # constants of the compiled KS residuals + the lag padding column (GeneralStructures.jl:350-354)
    check(b.ctx, ccall((:hank_ks_configure, LIB), Cint, (Ptr{Cvoid}, Float64, Float64, Float64), b.ctx, p.α, p.δ, ss_initial.vars.KS))
    b.initial_id = objectid(ss_initial)
end

# ── Dual packing: Vector{Dual{T,Float64,N}} is bit-compatible with an (N+1) x len Float64 matrix, value first
#    (ForwardDiff.jl/src/dual.jl:14-16, partials.jl:1-3) ─────────────────────────────────────────────────────────
const HankReal = Union{Float64,ForwardDiff.Dual{<:Any,Float64}}
_unpack(v::AbstractVector{Float64}) = (Vector{Float64}(v), Matrix{Float64}(undef, length(v), 0))
function _unpack(v::AbstractVector{ForwardDiff.Dual{Tg,Float64,N}}) where {Tg,N}
    raw = reinterpret(reshape, Float64, collect(v))                               # (N+1) x len
    (Vector{Float64}(raw[1, :]), Matrix{Float64}(permutedims(raw[2:end, :])))     # len x N
end
_repack(::Type{Float64}, val, part) = val
_repack(::Type{ForwardDiff.Dual{Tg,Float64,N}}, val::AbstractArray, part::AbstractMatrix) where {Tg,N} =
    [ForwardDiff.Dual{Tg}(val[i], ForwardDiff.Partials(ntuple(k -> part[i, k], Val(N)))) for i in eachindex(val)]

"""
`policy_seqs.KD` as returned by the device `BackwardIteration`: an `AbstractVector{Matrix{TF}}` whose matrices stay on
the GPU; `h[t]` (or `collect(h)`) materialises period t for callers that need host matrices
(SteadyStateJacobian.jl:226-229, SteadyState.jl:225), `ForwardIteration` consumes it without a transfer.
"""
struct DevicePolicies{TF} <: AbstractVector{Matrix{TF}}
    blk::HankBlock
    K::Int
end
Base.size(h::DevicePolicies) = (h.blk.P,)
function Base.getindex(h::DevicePolicies{TF}, t::Int) where {TF}
    b = h.blk; G = b.n_a * b.n_e
    val = Vector{Float64}(undef, G); part = Matrix{Float64}(undef, G, h.K)
    check(b.ctx, ccall((:hank_get_policy, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}), b.ctx, t, 0, val))
    for k in 1:h.K
        check(b.ctx, ccall((:hank_get_policy, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}), b.ctx, t, k, view(part, :, k)))
    end
    reshape(_repack(TF, val, part), b.n_a, b.n_e)
end

"""
    BackwardIteration(xVec_endog, exog_paths::NamedTuple, model::SequenceModel, ss_end)

BackwardIteration.jl:46-116 on the device (`hank_backward`): the reference's signature with `xVec_endog` restricted to
`Float64` / `Dual{…,Float64,N}` element types.  Only `r` and `w` enter the KS household block (KrusellSmith.jl:53-54).
Returns `(KD = DevicePolicies,)` — same key, same indexing as the reference's `NamedTuple{het}(Vector{Matrix})`.
"""
function BackwardIteration(xVec_endog::AbstractVector{TF}, exog_paths::NamedTuple, model::SequenceModel, ss_end) where {TF<:HankReal}
    _is_ks(model) || return invoke(BackwardIteration, Tuple{Any,NamedTuple,SequenceModel,Any}, xVec_endog, exog_paths, model, ss_end)
    b = _block(model)
    _set_terminal!(b, ss_end)
    xv, xp = _unpack(xVec_endog)
    K = size(xp, 2)
    X = reshape(xv, b.n_endog, b.P)
    r = X[b.ir, :]; w = X[b.iw, :]
    dr = Matrix{Float64}(undef, b.P, K); dw = similar(dr)
    for k in 1:K
        D = reshape(view(xp, :, k), b.n_endog, b.P)
        dr[:, k] .= D[b.ir, :]; dw[:, k] .= D[b.iw, :]
    end
    check(b.ctx, ccall((:hank_backward, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}),
                       b.ctx, r, w, K, dr, dw))
    b.K = K
    (KD = DevicePolicies{TF}(b, K),)
end

"""
    ForwardIteration(policy_seqs::NamedTuple, model::SequenceModel, ss_initial)

ForwardIteration.jl:253-311 on the device-resident policies left by `BackwardIteration` (`hank_forward`).
Host policy matrices (any other `policy_seqs`) go to the reference's method, or to `hank_forward_policies`
through `ForwardIterationDevice` below.
"""
function ForwardIteration(policy_seqs::NamedTuple{(:KD,),<:Tuple{DevicePolicies{TF}}}, model::SequenceModel, ss_initial) where {TF}
    b = policy_seqs.KD.blk
    _set_initial!(b, model, ss_initial)
    K = policy_seqs.KD.K
    KD = Vector{Float64}(undef, b.P); dKD = Matrix{Float64}(undef, b.P, K)
    check(b.ctx, ccall((:hank_forward, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), b.ctx, KD, dKD))
    (KD = _repack(TF, KD, dKD),)
end

"""ForwardIteration for caller-supplied `Vector{Matrix{Float64}}` policies (hank_forward_policies)."""
function ForwardIterationDevice(policies::AbstractVector{<:AbstractMatrix{Float64}}, model::SequenceModel, ss_initial)
    b = _block(model); _set_initial!(b, model, ss_initial)
    pol = reduce(hcat, vec.(policies))                                   # G x P
    KD = Vector{Float64}(undef, b.P)
    check(b.ctx, ccall((:hank_forward_policies, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       b.ctx, pol, 0, C_NULL, KD, C_NULL))
    (KD = KD,)
end

"""
    JVP(func, primal, tangents::AbstractMatrix) -> Matrix

Batched companion of `JVP(func, primal, tangent)` (GeneralStructures.jl:542-550): the K columns of `tangents` ride as
K lanes of ONE pass (`Dual{Tag,Float64,K}` seeds built like derivative.jl:12-15 builds its single one).
"""
function JVP(func::Function, primal::AbstractVector{Float64}, tangents::AbstractMatrix{Float64})
    K = size(tangents, 2)
    Tg = typeof(ForwardDiff.Tag(func, Float64))
    x = [ForwardDiff.Dual{Tg}(primal[i], ForwardDiff.Partials(ntuple(k -> tangents[i, k], K))) for i in eachindex(primal)]
    y = func(x)
    [ForwardDiff.partials(y[i], k) for i in eachindex(y), k in 1:K]
end

"""
    NewtonRaphsonHANK(x_0, J̅::SparseMatrixCSC, exog_paths, mod, ss_initial::SteadyState, ss_ending::SteadyState; ε = 1e-9)

NewtonRaphson.jl:27-114 with the sweeps, residuals, JVPs and the preconditioner solve on the device
(`hank_newton_solve`): same six positional arguments, same `ε`, same printed progress line per outer iteration count.
`HankB200.NEWTON_SOLVER[]` selects the inner solve (default `:gmres`, the reference's).
"""
function NewtonRaphsonHANK(x_0::Vector{Float64}, J̅::SparseMatrixCSC, exog_paths::NamedTuple, mod::SequenceModel,
                           ss_initial::SteadyState, ss_ending::SteadyState; ε = 1e-9)
    _is_ks(mod) || return invoke(NewtonRaphsonHANK, Tuple{Vector{Float64},SparseMatrixCSC,NamedTuple,SequenceModel,Any,Any},
                                 x_0, J̅, exog_paths, mod, ss_initial, ss_ending; ε = ε)
    b = _block(mod)
    _set_terminal!(b, ss_ending); _set_initial!(b, mod, ss_initial)
    n = length(x_0)
    x = Vector{Float64}(undef, n); stats = zeros(8); inner = zeros(Cint, 100)
    code = Dict(:gmres => 0, :lu => 1, :lu_batched => 2)[HankB200.NEWTON_SOLVER[]]
    check(b.ctx, ccall((:hank_newton_solve, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Cint}),
                       b.ctx, Matrix{Float64}(J̅), x_0, Vector{Float64}(exog_paths.Z), ε, 1e-9, code, x, stats, inner))
    println("Iteration: $(Int(stats[1]) + 1), norm(y): $(stats[4])   [device: $(Int(stats[2])) JVPs, inner $(Int.(inner[1:Int(stats[1])]))]")
    x
end

"""
    ValueFunction(value_next, xVals, model::SequenceModel)

The `value_fn` plug-in (KrusellSmith.jl:43-83) as one device EGM step (`hank_egm_step`) for `Float64` or `Dual`
inputs; `r` and `w` are read from `xVals` by name, as the reference does.
"""
function ValueFunction(value_next::AbstractMatrix{TV}, xVals::AbstractVector{TX}, model::SequenceModel) where {TV<:HankReal,TX<:HankReal}
    HankB200.ENABLED[] || return invoke(ValueFunction, Tuple{Any,Any,SequenceModel}, value_next, xVals, model)
    b = _block(model)
    TF = promote_type(TV, TX)
    names = var_names(model); jr = findfirst(==(:r), names); jw = findfirst(==(:w), names)
    vv, vp = _unpack(vec(value_next)); xv, xp = _unpack(collect(xVals))
    K = max(size(vp, 2), size(xp, 2))
    G = b.n_a * b.n_e
    dv = size(vp, 2) == K ? vp : zeros(G, K)
    dr = size(xp, 2) == K ? xp[jr, :] : zeros(K); dw = size(xp, 2) == K ? xp[jw, :] : zeros(K)
    val = Vector{Float64}(undef, G); pol = similar(val); dval = Matrix{Float64}(undef, G, K); dpol = similar(dval)
    check(b.ctx, ccall((:hank_egm_step, LIB), Cint,
                       (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Float64, Cint, Ptr{Float64}, Ptr{Float64},
                        Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       b.ctx, vv, K == 0 ? C_NULL : dv, xv[jr], xv[jw], K, dr, dw, val, pol, dval, dpol))
    sh(v, d) = reshape(_repack(TF, v, d), b.n_a, b.n_e)
    (Value = sh(val, dval), KD = sh(pol, dpol))
end

"""
    get_xVals(asm::SSAssembler, p_vec)

SteadyState.jl:111-154 with the inner VFI loop (:132-141, ~470 EGM steps per evaluation) run on the device by
`hank_vfi` (value_fn iterated from `ones` until `max|ΔValue| < ε`, carrying the `Dual` lanes of `p_vec`); the
stationary distribution and the aggregation stay the reference's Julia code (`make_endogenous_transition`,
`invariant_dist`, `dot`).
"""
function get_xVals(asm::SSAssembler, p_vec::AbstractVector{T_num}) where {T_num<:HankReal}
    model = asm.model
    _is_ks(model) || return invoke(get_xVals, Tuple{SSAssembler,AbstractVector}, asm, p_vec)
    all_keys, free_keys, endog_dim, n_exog = asm.all_keys, asm.free_keys, asm.endog_dim, asm.n_exog
    xVals = zeros(T_num, model.compspec.n_v)
    for (i, k) in enumerate(free_keys)
        xVals[findfirst(==(k), all_keys)] = p_vec[i]
    end
    for (sym, val) in pairs(asm.ss_spec.fixed)
        xVals[findfirst(==(sym), all_keys)] = val
    end
    b = _block(model)
    jr = findfirst(==(:r), all_keys); jw = findfirst(==(:w), all_keys)
    xv, xp = _unpack(xVals); K = size(xp, 2); G = b.n_a * b.n_e
    val = Vector{Float64}(undef, G); pol = similar(val); dval = Matrix{Float64}(undef, G, K); dpol = similar(dval)
    iters = Ref{Cint}(0)
    check(b.ctx, ccall((:hank_vfi, LIB), Cint,
                       (Ptr{Cvoid}, Float64, Float64, Cint, Ptr{Float64}, Ptr{Float64}, Float64, Cint,
                        Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Cint}),
                       b.ctx, xv[jr], xv[jw], K, xp[jr, :], xp[jw, :], model.compspec.ε, 10_000, val, pol, dval, dpol, iters))
    b.terminal_id = 0   # hank_vfi uses the context's sweep buffers: the next sweep re-uploads its terminal value
    Value = reshape(_repack(T_num, val, dval), b.n_a, n_exog)
    KD = reshape(_repack(T_num, pol, dpol), b.n_a, n_exog)
    Λ_endog = make_endogenous_transition(KD, endog_dim, n_exog)
    D = invariant_dist((asm.Λ_exog * Λ_endog)')
    xVals[findfirst(==(:KD), all_keys)] = dot(vec(KD), D)
    return xVals, Value
end

"""Columns `cols` of the sequence-space Jacobian at `x` as unit-seed JVPs in batched lanes: `directJVPJacobian`
(SteadyState.jl:296-320) generalised to any column range (BASELINE config 3).  Y / KS columns skip the sweeps."""
function directJVPJacobian(x::Vector{Float64}, exog_paths::NamedTuple, model::SequenceModel, ss_initial::SteadyState,
                           ss_ending::SteadyState, cols::UnitRange{Int})
    b = _block(model); _set_terminal!(b, ss_ending); _set_initial!(b, model, ss_initial)
    F = Vector{Float64}(undef, length(x))
    check(b.ctx, ccall((:hank_ks_linearize, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       b.ctx, x, Vector{Float64}(exog_paths.Z), F))
    J = Matrix{Float64}(undef, length(x), length(cols))
    check(b.ctx, ccall((:hank_ks_jacobian_columns, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{Float64}),
                       b.ctx, first(cols), last(cols) + 1, J))
    sparse(J)
end

# ── multi-GPU: one Julia process per GPU (Distributed / MPI.jl); the unique id travels through the host program ──
"""`inv(J̅)` on the device (`hank_dense_inverse`: the blocked Gauss-Jordan inverse `hank_newton_solve` uses for its
preconditioner solve, NewtonRaphson.jl:97) for callers that want the inverse itself."""
function dense_inverse(model::SequenceModel, A::AbstractMatrix{Float64})
    b = _block(model); n = size(A, 1)
    n == size(A, 2) || error("dense_inverse: the matrix must be square")
    Ad = Matrix{Float64}(A); out = Matrix{Float64}(undef, n, n)
    check(b.ctx, ccall((:hank_dense_inverse, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}), b.ctx, n, Ad, out))
    return out
end

# ── the model's equations as device bytecode (hank_eq_configure; the host half of compile_residuals, ModelParser.jl:217-259) ──
const _EQ_OPS = Dict(:+ => 2, :- => 3, :* => 4, :/ => 5, :^ => 6)
const _EQ_FUNS = Dict(:exp => 8, :log => 9, :sqrt => 10)
function _eq_emit!(code::Vector{Cint}, consts::Vector{Float64}, ex, vidx::Dict{Symbol,Int}, params)
    pushc(v) = (i = findfirst(==(Float64(v)), consts); i === nothing && (push!(consts, Float64(v)); i = length(consts)); append!(code, Cint[0, i - 1]))
    if ex isa Number
        pushc(ex)
    elseif ex isa Symbol
        haskey(vidx, ex) ? append!(code, Cint[1, vidx[ex], 0]) : pushc(getfield(params, ex))
    elseif ex isa Expr && ex.head == :call
        f, args = ex.args[1], ex.args[2:end]
        if f isa Symbol && haskey(vidx, f) && length(args) == 1 && args[1] isa Integer      # VAR(±k)
            append!(code, Cint[1, vidx[f], args[1]])
        elseif f == :- && length(args) == 1
            _eq_emit!(code, consts, args[1], vidx, params); push!(code, 7)
        elseif haskey(_EQ_OPS, f)                                                            # n-ary: left fold like transform_expr
            _eq_emit!(code, consts, args[1], vidx, params)
            for a in args[2:end]
                _eq_emit!(code, consts, a, vidx, params); push!(code, _EQ_OPS[f])
            end
        elseif haskey(_EQ_FUNS, f) && length(args) == 1
            _eq_emit!(code, consts, args[1], vidx, params); push!(code, _EQ_FUNS[f])
        else
            error("HankB200: unsupported call $f in an equation")
        end
    else
        error("HankB200: unsupported expression $ex")
    end
end
"""Uploads `model.equations` as bytecode: after this call the device evaluates F, J·V and Jacobian columns of the model's own
aggregate block (any number of endogenous / exogenous variables, lags and leads) instead of the built-in Krusell-Smith one."""
function configure_equations!(model::SequenceModel, equations::Vector{String}, ss_start, ss_end)
    b = _block(model)
    names = var_names(model); endog = vars_of_type(model, :endogenous); exog = vars_of_type(model, :exogenous)
    vidx = Dict(s => i - 1 for (i, s) in enumerate(names))
    code = Cint[]; consts = Float64[]; off = Cint[0]
    for eq in equations
        parts = split(eq, "="; limit = 2)
        length(parts) == 2 || error("Equation must contain exactly one '=': $eq")
        _eq_emit!(code, consts, Meta.parse(strip(String(parts[1]))), vidx, model.params)
        _eq_emit!(code, consts, Meta.parse(strip(String(parts[2]))), vidx, model.params)
        push!(code, 3); push!(off, length(code))
    end
    ss0 = Float64[ss_start.vars[k] for k in names]; ss1 = Float64[ss_end.vars[k] for k in names]
    check(b.ctx, ccall((:hank_eq_configure, LIB), Cint,
                       (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Ptr{Cint}, Ptr{Cint}, Cint, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                       b.ctx, length(endog), length(exog), vidx[:r], vidx[:w], off, code, length(consts), consts, ss0, ss1))
end

comm_unique_id() = (id = zeros(UInt8, 128); ccall((:hank_comm_unique_id, LIB), Cint, (Ptr{UInt8},), id) == 0 || error("ncclGetUniqueId failed"); id)
comm_init(model::SequenceModel, nranks, rank, id::Vector{UInt8}) =
    (b = _block(model); check(b.ctx, ccall((:hank_comm_init, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), b.ctx, nranks, rank, id)))
"""All-gather equal-sized column blocks held in host matrices (`hank_allgather_columns`): every rank passes its
`n x k` block and receives `n x (nranks*k)`."""
function allgather_columns(model::SequenceModel, loc::Matrix{Float64}, nranks::Int)
    b = _block(model)
    all = Matrix{Float64}(undef, size(loc, 1), size(loc, 2) * nranks)
    check(b.ctx, ccall((:hank_allgather_columns, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Csize_t, Ptr{Float64}), b.ctx, loc, length(loc), all))
    all
end
