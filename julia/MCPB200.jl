# MCPB200.jl — thin `ccall` wrapper that keeps MixedComplementarityProblems.jl's API surface and routes the
# Newton hot path to libmcpb200.so (B200, sm_100a).  NOT EXECUTABLE IN THE BUILD IMAGE (no Julia there): the
# boundary is proven through the same C ABI from Python (mcp_b200/capi.py, tests/); this file is the binding a
# maintainer of the reference would add.  See INTEGRATION.md.
#
# Reference surface kept (file:line in TianyuQ/MCP):
#   PrimalDualMCP(G, H; unconstrained_dimension, constrained_dimension, parameter_dimension,
#                 compute_sensitivities)                                   src/mcp.jl:27-52
#   PrimalDualMCP(K, lower_bounds, upper_bounds; parameter_dimension, …)    src/mcp.jl:155-177
#   solve(InteriorPoint(), mcp, θ; x₀, y₀, s₀, tol, max_inner_iters, …)    src/solver.jl:35-51
#   ChainRulesCore.rrule(solve, …) / ForwardDiff.Dual overload             src/AutoDiff.jl:42-117
# New: solve(InteriorPoint(), mcp, Θ::AbstractMatrix; X₀, Y₀, S₀, …) — one solve per column of Θ.
module MCPB200

using Symbolics: Symbolics, Num
using SparseArrays: findnz
using ChainRulesCore: ChainRulesCore
using ForwardDiff: ForwardDiff
using LinearAlgebra: I

const LIB = get(ENV, "MCPB200_LIB", joinpath(@__DIR__, "..", "mcp_b200", "libmcpb200.so"))

abstract type SolverType end
struct InteriorPoint <: SolverType end

# ---- include/mcpb200.h ---------------------------------------------------------------------------------
struct ProblemDesc
    nx::Int32; ny::Int32; ntheta::Int32; n_nodes::Int32
    op::Ptr{Int32}; a::Ptr{Int32}; b::Ptr{Int32}
    n_consts::Int32; consts::Ptr{Float64}; gh_nodes::Ptr{Int32}
    jz_nnz::Int32; jz_rows::Ptr{Int32}; jz_cols::Ptr{Int32}; jz_nodes::Ptr{Int32}
    jt_nnz::Int32; jt_rows::Ptr{Int32}; jt_cols::Ptr{Int32}; jt_nodes::Ptr{Int32}
end

struct SolverOpts
    tol::Float64; max_inner_iters::Int32; max_outer_iters::Int32
    tightening_rate::Float64; loosening_rate::Float64; min_stepsize::Float64
end

const OK = 0
const ERR_NO_SENSITIVITIES = -5

last_error(h) = unsafe_string(ccall((:mcpb200_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
global_error() = unsafe_string(ccall((:mcpb200_global_error, LIB), Cstring, ()))
check(h, rc) = rc == OK || error("libmcpb200 error $rc: $(last_error(h))")

# ---- MCP-IR emitter: Symbolics expression DAG -> flat tape (op-codes of enum mcpb200_op) ----------------
const OP = (CONST = 0, X = 1, Y = 2, THETA = 3, ADD = 4, SUB = 5, MUL = 6, DIV = 7, NEG = 8, SQRT = 9,
            EXP = 10, LOG = 11, SIN = 12, COS = 13, POWI = 14)

mutable struct Tape
    op::Vector{Int32}; a::Vector{Int32}; b::Vector{Int32}; consts::Vector{Float64}
    memo::Dict{Any,Int32}; leaves::Dict{Any,Tuple{Int32,Int32}}
end
Tape() = Tape(Int32[], Int32[], Int32[], Float64[], Dict{Any,Int32}(), Dict{Any,Tuple{Int32,Int32}}())

function push_node!(t::Tape, op, a, b = -1)
    get!(t.memo, (op, a, b)) do
        push!(t.op, op); push!(t.a, a); push!(t.b, b)
        Int32(length(t.op) - 1)
    end
end

function emit!(t::Tape, ex)
    ex = Symbolics.unwrap(ex)
    if ex isa Real
        push!(t.consts, Float64(ex)); return push_node!(t, OP.CONST, Int32(length(t.consts) - 1))
    elseif haskey(t.leaves, ex)
        kind, idx = t.leaves[ex]; return push_node!(t, kind, idx)
    end
    f, args = Symbolics.operation(ex), Symbolics.arguments(ex)
    ids = [emit!(t, a) for a in args]
    fold(op) = foldl((l, r) -> push_node!(t, op, l, r), ids)
    f === (+) && return fold(OP.ADD)
    f === (*) && return fold(OP.MUL)
    f === (-) && return length(ids) == 1 ? push_node!(t, OP.NEG, ids[1]) : push_node!(t, OP.SUB, ids[1], ids[2])
    f === (/) && return push_node!(t, OP.DIV, ids[1], ids[2])
    f === (^) && Symbolics.unwrap(args[2]) isa Integer && return push_node!(t, OP.POWI, ids[1], Int32(Symbolics.unwrap(args[2])))
    f === sqrt && return push_node!(t, OP.SQRT, ids[1])
    f === exp && return push_node!(t, OP.EXP, ids[1])
    f === log && return push_node!(t, OP.LOG, ids[1])
    f === sin && return push_node!(t, OP.SIN, ids[1])
    f === cos && return push_node!(t, OP.COS, ids[1])
    error("MCP-IR has no op for $f")
end

# ---- PrimalDualMCP (src/mcp.jl:13-24 fields kept; the compiled closures are replaced by a device handle) --
mutable struct PrimalDualMCP
    handle::Ptr{Cvoid}
    unconstrained_dimension::Int
    constrained_dimension::Int
    parameter_dimension::Int
    has_sensitivities::Bool
end

"Symbolic constructor, src/mcp.jl:55-150: G, H, x, y, θ are Vector{Num}."
function PrimalDualMCP(G_sym::Vector{Num}, H_sym::Vector{Num}, x_sym::Vector{Num}, y_sym::Vector{Num},
                       θ_sym::Vector{Num}; compute_sensitivities = true, backend_options = (;))
    t = Tape()
    for (i, v) in enumerate(x_sym); t.leaves[Symbolics.unwrap(v)] = (OP.X, Int32(i - 1)); end
    for (i, v) in enumerate(y_sym); t.leaves[Symbolics.unwrap(v)] = (OP.Y, Int32(i - 1)); end
    for (i, v) in enumerate(θ_sym); t.leaves[Symbolics.unwrap(v)] = (OP.THETA, Int32(i - 1)); end
    GH = [G_sym; H_sym]
    gh_nodes = Int32[emit!(t, e) for e in GH]
    Jz = Symbolics.sparsejacobian(GH, [x_sym; y_sym])              # src/mcp.jl:97-99
    rz, cz, vz = findnz(Jz)                                        # CSC order, src/mcp.jl:110
    jz_nodes = Int32[emit!(t, e) for e in vz]
    if compute_sensitivities
        Jt = Symbolics.sparsejacobian(GH, θ_sym)                   # src/mcp.jl:125-126
        rt, ct, vt = findnz(Jt)
        jt_nodes = Int32[emit!(t, e) for e in vt]
    else
        rt, ct, jt_nodes = Int[], Int[], Int32[]
    end
    rz32, cz32 = Int32.(rz .- 1), Int32.(cz .- 1)
    rt32, ct32 = Int32.(rt .- 1), Int32.(ct .- 1)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve t gh_nodes jz_nodes jt_nodes rz32 cz32 rt32 ct32 begin
        desc = ProblemDesc(length(x_sym), length(y_sym), length(θ_sym), length(t.op), pointer(t.op), pointer(t.a),
                           pointer(t.b), length(t.consts), pointer(t.consts), pointer(gh_nodes), length(jz_nodes),
                           pointer(rz32), pointer(cz32), pointer(jz_nodes),
                           compute_sensitivities ? length(jt_nodes) : -1, pointer(rt32), pointer(ct32), pointer(jt_nodes))
        rc = ccall((:mcpb200_create, LIB), Cint, (Ref{ProblemDesc}, UInt32, Ref{Ptr{Cvoid}}), desc, 0, h)
        rc == OK || error("mcpb200_create failed ($rc): $(global_error())")
    end
    mcp = PrimalDualMCP(h[], length(x_sym), length(y_sym), length(θ_sym), compute_sensitivities)
    finalizer(m -> ccall((:mcpb200_destroy, LIB), Cint, (Ptr{Cvoid},), m.handle), mcp)
end

"Callable constructor, src/mcp.jl:27-52."
function PrimalDualMCP(G, H; unconstrained_dimension, constrained_dimension, parameter_dimension,
                       compute_sensitivities = true, backend = nothing, backend_options = (;))
    x = Symbolics.variables(:x, 1:unconstrained_dimension)
    y = Symbolics.variables(:y, 1:constrained_dimension)
    θ = Symbolics.variables(:θ, 1:parameter_dimension)
    PrimalDualMCP(collect(Num, G(x, y; θ)), collect(Num, H(x, y; θ)), x, y, θ; compute_sensitivities)
end

"K(z; θ) with bounds, src/mcp.jl:155-210."
function PrimalDualMCP(K, lower_bounds::Vector, upper_bounds::Vector; parameter_dimension,
                       compute_sensitivities = true, backend = nothing, backend_options = (;))
    @assert all(isinf.(upper_bounds)) && all(isinf.(lower_bounds) .|| lower_bounds .== 0)   # src/mcp.jl:191
    z = Symbolics.variables(:z, 1:length(lower_bounds))
    θ = Symbolics.variables(:θ, 1:parameter_dimension)
    Ksym = collect(Num, K(z; θ))
    unc, con = findall(isinf, lower_bounds), findall(!isinf, lower_bounds)                  # src/mcp.jl:193-194
    PrimalDualMCP(Ksym[unc], Ksym[con], z[unc], z[con], θ; compute_sensitivities)
end

# ---- solve ---------------------------------------------------------------------------------------------------
function _solve_batched(mcp::PrimalDualMCP, Θ::Matrix{Float64}, X₀, Y₀, S₀, opts::SolverOpts)
    B = size(Θ, 2)
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    x, y, s = Matrix{Float64}(undef, nx, B), Matrix{Float64}(undef, ny, B), Matrix{Float64}(undef, ny, B)
    kkt, ϵ = Vector{Float64}(undef, B), Vector{Float64}(undef, B)
    outer, status, steps = Vector{Int32}(undef, B), Vector{Int32}(undef, B), Vector{Int32}(undef, B)
    ptr(A) = A === nothing ? Ptr{Float64}(C_NULL) : pointer(A)
    GC.@preserve Θ X₀ Y₀ S₀ x y s kkt ϵ outer status steps begin
        rc = ccall((:mcpb200_solve_batched, LIB), Cint,
                   (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{SolverOpts},
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}),
                   mcp.handle, B, Θ, ptr(X₀), ptr(Y₀), ptr(S₀), opts, x, y, s, kkt, ϵ, outer, status, steps)
        check(mcp.handle, rc)
    end
    (; status, x, y, s, kkt_error = kkt, ϵ, outer_iters = outer, newton_steps = steps)
end

"Batched solve: one interior-point solve per column of Θ (new); kwargs as src/solver.jl:39-50."
function solve(::InteriorPoint, mcp::PrimalDualMCP, Θ::AbstractMatrix{<:Real}; X₀ = nothing, Y₀ = nothing, S₀ = nothing,
               tol = 1e-4, max_inner_iters = 20, max_outer_iters = 50, tightening_rate = 0.1, loosening_rate = 0.5,
               min_stepsize = 1e-4, verbose = false, linear_solve_algorithm = nothing)
    isnothing(linear_solve_algorithm) || @warn "linear_solve_algorithm is ignored: the B200 path has one fixed KKT solver"
    opts = SolverOpts(tol, max_inner_iters, max_outer_iters, tightening_rate, loosening_rate, min_stepsize)
    dense(A) = A === nothing ? nothing : Matrix{Float64}(A)
    _solve_batched(mcp, Matrix{Float64}(Θ), dense(X₀), dense(Y₀), dense(S₀), opts)
end

"""
`solve(InteriorPoint(), mcp, θ; x₀, y₀, s₀, …)` — src/solver.jl:35-122; returns the same NamedTuple (src/solver.jl:121).

Like the reference, which sets `x = x₀` and updates it in place (src/solver.jl:64-66,103-105), the returned `x`, `y`, `s`
ALIAS the caller's `x₀`, `y₀`, `s₀` whenever those are mutable `Vector{Float64}`s: the solution is written back into
them and they are what the NamedTuple holds.  (Other array types — ranges, views, non-Float64 — are left untouched and
fresh vectors are returned, where the reference would have thrown on the in-place update.)
"""
function solve(ip::InteriorPoint, mcp::PrimalDualMCP, θ::AbstractVector{<:Real}; x₀ = nothing, y₀ = nothing, s₀ = nothing, kwargs...)
    col(v) = v === nothing ? nothing : reshape(collect(Float64, v), :, 1)
    r = solve(ip, mcp, reshape(collect(Float64, θ), :, 1); X₀ = col(x₀), Y₀ = col(y₀), S₀ = col(s₀), kwargs...)
    alias(v₀, v) = v₀ isa Vector{Float64} ? copyto!(v₀, v) : v
    (; status = r.status[1] == 0 ? :solved : :failed, x = alias(x₀, r.x[:, 1]), y = alias(y₀, r.y[:, 1]), s = alias(s₀, r.s[:, 1]),
       kkt_error = r.kkt_error[1], ϵ = r.ϵ[1], outer_iters = Int(r.outer_iters[1]))
end

# ---- sensitivities (src/AutoDiff.jl) ---------------------------------------------------------------------------
"`_solve_jacobian_θ`, src/AutoDiff.jl:18-40."
function _solve_jacobian_θ(mcp::PrimalDualMCP, solution, θ)
    mcp.has_sensitivities || throw(ArgumentError(
        "Missing sensitivities. Set `compute_sensitivities = true` when constructing the PrimalDualMCP."))   # :19-23
    n = mcp.unconstrained_dimension + 2mcp.constrained_dimension
    J = Matrix{Float64}(undef, n, mcp.parameter_dimension)
    θv, x, y, s, ϵ = Float64.(θ), Float64.(solution.x), Float64.(solution.y), Float64.(solution.s), [Float64(solution.ϵ)]
    GC.@preserve θv x y s ϵ J begin
        rc = ccall((:mcpb200_sensitivities, LIB), Cint,
                   (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                    Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                   mcp.handle, 1, θv, x, y, s, ϵ, J, C_NULL, C_NULL, 0, C_NULL, C_NULL, C_NULL)
        check(mcp.handle, rc)
    end
    J
end

function ChainRulesCore.rrule(::typeof(solve), solver_type::SolverType, mcp::PrimalDualMCP, θ; kwargs...)   # src/AutoDiff.jl:42-82
    solution = solve(solver_type, mcp, θ; kwargs...)
    project_to_θ = ChainRulesCore.ProjectTo(θ)
    function solve_pullback(∂solution)
        ∂θ = ChainRulesCore.@thunk let
            nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
            z̄ = zeros(nx + 2ny)
            for (rng, t) in ((1:nx, ∂solution.x), (nx+1:nx+ny, ∂solution.y), (nx+ny+1:nx+2ny, ∂solution.s))
                t isa ChainRulesCore.AbstractZero || (z̄[rng] .= t)
            end
            θ̄ = Vector{Float64}(undef, mcp.parameter_dimension)
            θv, x, y, s, ϵ = Float64.(θ), solution.x, solution.y, solution.s, [solution.ϵ]
            GC.@preserve θv x y s ϵ z̄ θ̄ check(mcp.handle, ccall((:mcpb200_sensitivities, LIB), Cint,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                 Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                mcp.handle, 1, θv, x, y, s, ϵ, C_NULL, z̄, θ̄, 0, C_NULL, C_NULL, C_NULL))
            project_to_θ(θ̄)                                                                              # :65-75
        end
        ChainRulesCore.NoTangent(), ChainRulesCore.NoTangent(), ChainRulesCore.NoTangent(), ∂θ            # :53-57,78
    end
    solution, solve_pullback
end

function solve(solver_type::InteriorPoint, mcp::PrimalDualMCP, θ::AbstractVector{<:ForwardDiff.Dual{T}}; kwargs...) where {T}   # src/AutoDiff.jl:84-117
    θ_v, θ_p = ForwardDiff.value.(θ), ForwardDiff.partials.(θ)
    solution = solve(solver_type, mcp, θ_v; kwargs...)
    z_p = _solve_jacobian_θ(mcp, solution, θ_v) * θ_p                                                      # :96-98
    nx, ny = mcp.unconstrained_dimension, mcp.constrained_dimension
    x_d = ForwardDiff.Dual{T}.(solution.x, @view z_p[1:nx])
    y_d = ForwardDiff.Dual{T}.(solution.y, @view z_p[nx+1:nx+ny])
    s_d = ForwardDiff.Dual{T}.(solution.s, @view z_p[nx+ny+1:end])   # NB the reference wraps solution.y here (:109-114, a bug)
    (; solution.status, solution.kkt_error, solution.ϵ, x = x_d, y = y_d, s = s_d)
end

end # module
