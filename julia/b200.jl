# b200.jl -- Julia side of the drop-in boundary: `include("b200.jl")` at the end of src/Penguin.jl (after the reference's own
# definitions) and the unsteady cut-cell diffusion path runs on libpenguin_b200.so; every other subsystem of Penguin.jl is untouched.
#
# The methods below have the SAME names and argument lists as the reference's (cited per method as file:line under
# /root/reference/src) and are MORE SPECIFIC in one argument -- the body is a `B200Body` (a callable struct, `<: Function`, that the
# GPU can evaluate) -- so that Julia's dispatch selects them for those bodies and keeps the reference's CPU methods for everything else
# (arbitrary closures: `Capacity(body::Function, mesh)` stays the reference's; `b200_import(cap)` uploads such a capacity afterwards).
#
# Julia is not part of the build image of this repository (DESIGN.md section 1): this file is written against include/penguin_b200.h and
# mirrored call for call by penguin.jl_b200/api.py, which the test-suite drives; tests/abi_smoke.c walks the same call sequence from C.
#
# Memory / threading contract: every host buffer is a Julia `Vector{Float64}` owned by Julia and pinned for the duration of the `ccall`
# (`GC.@preserve`); the library never keeps a host pointer; handles are freed by finalizers; calls are blocking; no callbacks into Julia.

using SparseArrays, StaticArrays, Libdl

const libpb = get(ENV, "PENGUIN_B200_LIB", "libpenguin_b200.so")

# ---- status codes, error conversion (the reference throws `error(...)`, e.g. src/solver/diffusion.jl:269-271) ----------------------
const PB200_OK, PB200_ENOTCONV = Cint(0), Cint(5)
b200_last_error(ctx::Ptr{Cvoid} = C_NULL) = unsafe_string(ccall((:pb200_last_error, libpb), Cstring, (Ptr{Cvoid},), ctx))
function pbcheck(rc::Integer; allow = ())
    (rc == PB200_OK || rc in allow) && return rc
    error("libpenguin_b200 error $rc: $(b200_last_error(B200_CTX[]))")
end

# ---- context: one per process (pb200_init); several GPUs of one box from ONE Julia process: pb200_init_multi (below) ----------------
const B200_CTX = Ref{Ptr{Cvoid}}(C_NULL)
function b200_context(device::Integer = parse(Int, get(ENV, "PB200_DEVICE", "0")))
    if B200_CTX[] == C_NULL
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = ccall((:pb200_init, libpb), Cint, (Ref{Ptr{Cvoid}}, Cint), h, device)
        rc == PB200_OK || error("libpenguin_b200: $(b200_last_error()) (there is no CPU fallback)")
        B200_CTX[] = h[]
        atexit(() -> (B200_CTX[] != C_NULL && ccall((:pb200_finalize, libpb), Cint, (Ptr{Cvoid},), B200_CTX[]); B200_CTX[] = C_NULL))
    end
    B200_CTX[]
end
"""One Julia process driving `length(devices)` GPUs (slabs of the slowest dimension): the handles returned by the calls below then stand for
the whole team; per-cell host arrays keep the reference's GLOBAL padded length (the library cuts and reassembles the slabs)."""
function b200_context_multi(devices::Vector{<:Integer})
    B200_CTX[] == C_NULL || error("context already initialised")
    h = Ref{Ptr{Cvoid}}(C_NULL)
    dev = Cint.(devices)
    pbcheck(ccall((:pb200_init_multi, libpb), Cint, (Ref{Ptr{Cvoid}}, Ptr{Cint}, Cint), h, dev, length(dev)))
    B200_CTX[] = h[]
end

# handles of Julia objects (objects of the reference's own types carry no extra field): WeakKeyDict + finalizers
const B200_HANDLES = WeakKeyDict{Any,Ptr{Cvoid}}()
function b200_register!(obj, h::Ptr{Cvoid}, destroy::Symbol)
    B200_HANDLES[obj] = h
    finalizer(o -> (B200_CTX[] != C_NULL && ccall((destroy, libpb), Cint, (Ptr{Cvoid},), h); nothing), obj)
    obj
end
b200_handle(obj) = get(B200_HANDLES, obj) do
    error("this object was not created by the B200 path (use a B200Body, or b200_import(capacity))")
end

# ---- GPU-evaluable bodies: callable like the closures of the reference's scripts, sign convention fluid = {body < 0} ----------------
abstract type B200Body <: Function end
struct B200Balls{N} <: B200Body            # phi = min_k |x - c_k| - r_k (disjoint balls: interval / circle / sphere(s))
    centers::Vector{NTuple{N,Float64}}
    radii::Vector{Float64}
    fluid_inside::Bool
end
struct B200HalfSpace <: B200Body           # phi = x[dim] - c
    dim::Int
    c::Float64
    fluid_below::Bool
end
Circle(center::NTuple{2,<:Real}, r::Real) = B200Balls{2}([Float64.(center)], [Float64(r)], true)
Sphere(center::NTuple{3,<:Real}, r::Real) = B200Balls{3}([Float64.(center)], [Float64(r)], true)
Interval(center::Real, r::Real) = B200Balls{1}([(Float64(center),)], [Float64(r)], true)
Base.:-(b::B200Balls{N}) where {N} = B200Balls{N}(b.centers, b.radii, !b.fluid_inside)     # the reference's `-(...)` bodies
Base.:-(b::B200HalfSpace) = B200HalfSpace(b.dim, b.c, !b.fluid_below)
function (b::B200Balls{N})(x...) where {N}
    phi = minimum(sqrt(sum((x[d] - c[d])^2 for d in 1:N)) - r for (c, r) in zip(b.centers, b.radii))
    b.fluid_inside ? phi : -phi
end
(b::B200HalfSpace)(x...) = b.fluid_below ? x[b.dim] - b.c : -(x[b.dim] - b.c)

struct pb200_levelset                       # field for field as in include/penguin_b200.h
    kind::Cint; nballs::Cint; centers::Ptr{Cdouble}; radii::Ptr{Cdouble}; fluid_inside::Cint; hs_dim::Cint; hs_c::Cdouble
end

b200_mesh_args(mesh::Mesh{N}) where {N} = begin
    n = Cint[length(mesh.centers[d]) for d in 1:N]
    x0 = Float64[mesh.centers[d][1] for d in 1:N]                                            # centers[d][1] = x0 (src/mesh.jl:49)
    h = Float64[N > 0 && length(mesh.centers[d]) > 1 ? mesh.centers[d][2] - mesh.centers[d][1] : 2 * (mesh.nodes[d][1] - mesh.centers[d][1]) for d in 1:N]
    (n, x0, h .* n)                                                                          # domain_size = n h
end

# ---- Capacity(body, mesh; method="VOFI", compute_centroids=true)   replaces src/capacity.jl:51-64, 81-123, 137-197 -----------------
function Capacity(body::B200Body, mesh::Mesh{N}; method::String = "VOFI", compute_centroids::Bool = true) where {N}
    ctx = b200_context()
    n, x0, L = b200_mesh_args(mesh)
    if body isa B200Balls
        cen = Float64[c[d] for c in body.centers for d in 1:N]                               # ball-major
        rad = copy(body.radii)
        ls = pb200_levelset(0, length(rad), pointer(cen), pointer(rad), body.fluid_inside, 0, 0.0)
    else
        cen = Float64[]; rad = Float64[]
        ls = pb200_levelset(1, 0, C_NULL, C_NULL, body.fluid_below, body.dim - 1, body.c)
    end
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve cen rad n x0 L pbcheck(ccall((:pb200_capacity_create, libpb), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{pb200_levelset}, Cint, Ref{Ptr{Cvoid}}), ctx, N, n, x0, L, Ref(ls), compute_centroids, h))
    b200_capacity_from_handle(h[], mesh, body, compute_centroids)
end
# fields of the reference struct (src/capacity.jl:25-36) filled from the device: diagonal sparse matrices from the exported vectors
function b200_capacity_from_handle(h::Ptr{Cvoid}, mesh::Mesh{N}, body, has_cg::Bool) where {N}
    nt = prod(length(mesh.centers[d]) + 1 for d in 1:N)
    V, Γ, ct = zeros(nt), zeros(nt), zeros(nt)
    A, B, W, Cω, Cγ = zeros(N * nt), zeros(N * nt), zeros(N * nt), zeros(N * nt), zeros(N * nt)
    pbcheck(ccall((:pb200_capacity_export, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
                   Ptr{Cdouble}, Ptr{Cdouble}), h, V, Γ, ct, A, B, W, Cω, Cγ))
    dg(v, d) = spdiagm(0 => v[(d-1)*nt+1:d*nt])
    cap = Capacity{N}(ntuple(d -> dg(A, d), N), ntuple(d -> dg(B, d), N), spdiagm(0 => V), ntuple(d -> dg(W, d), N),
                      [SVector{N}(ntuple(d -> Cω[(d-1)*nt+i], N)) for i in 1:nt],
                      has_cg ? [SVector{N}(ntuple(d -> Cγ[(d-1)*nt+i], N)) for i in 1:nt] : SVector{N,Float64}[],
                      spdiagm(0 => Γ), ct, mesh, body)
    b200_register!(cap, h, :pb200_capacity_destroy)
end
"""Upload a capacity computed by the reference's CPU code (arbitrary closure bodies) -- pb200_capacity_import."""
function b200_import(cap::Capacity{N}) where {N}
    haskey(B200_HANDLES, cap) && return cap
    ctx = b200_context()
    n, x0, L = b200_mesh_args(cap.mesh)
    cat(t) = vcat((Vector(diag(m)) for m in t)...)
    V, Γ, ct = Vector(diag(cap.V)), Vector(diag(cap.Γ)), cap.cell_types
    A, B, W = cat(cap.A), cat(cap.B), cat(cap.W)
    Cω = Float64[c[d] for d in 1:N for c in cap.C_ω]
    Cγ = isempty(cap.C_γ) ? Float64[] : Float64[c[d] for d in 1:N for c in cap.C_γ]
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve V Γ ct A B W Cω Cγ pbcheck(ccall((:pb200_capacity_import, libpb), Cint,
        (Ptr{Cvoid}, Cint, Ptr{Cint}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble},
         Ptr{Cdouble}, Ref{Ptr{Cvoid}}), ctx, N, n, x0, L, V, Γ, ct, A, B, W, Cω, isempty(Cγ) ? C_NULL : pointer(Cγ), h))
    b200_register!(cap, h[], :pb200_capacity_destroy)
end

# ---- DiffusionOps(cap)   replaces src/operators.jl:127-178 (G, H are never assembled: empty placeholders; W! and V are real) ----------
function DiffusionOps(cap::Capacity{N}) where {N}
    haskey(B200_HANDLES, cap) || return invoke(DiffusionOps, Tuple{AbstractCapacity}, cap)    # a CPU capacity: the reference's constructor
    h = Ref{Ptr{Cvoid}}(C_NULL)
    pbcheck(ccall((:pb200_ops_create, libpb), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}), b200_handle(cap), h))
    sz = ntuple(d -> length(cap.mesh.centers[d]) + 1, N)
    nt = prod(sz)
    wd = zeros(N * nt)
    pbcheck(ccall((:pb200_ops_export_wdag, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h[], wd))
    op = DiffusionOps{N}(spzeros(Float64, Int, 0, 0), spzeros(Float64, Int, 0, 0), spdiagm(0 => wd), cap.V, sz)
    b200_register!(op, h[], :pb200_ops_destroy)
end
# ∇(op, p) = W!(G pω + H pγ), ∇₋(op, qω, qγ) = -(G'+H') qω + H' qγ   (src/operators.jl:20-34)
function ∇(op::DiffusionOps{N}, p::Vector{Float64}) where {N}
    haskey(B200_HANDLES, op) || return invoke(∇, Tuple{AbstractOperators,Vector{Float64}}, op, p)
    out = zeros(N * prod(op.size))
    pbcheck(ccall((:pb200_ops_grad, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), b200_handle(op), p, out)); out
end
function ∇₋(op::DiffusionOps{N}, qω::Vector{Float64}, qγ::Vector{Float64}) where {N}
    haskey(B200_HANDLES, op) || return invoke(∇₋, Tuple{AbstractOperators,Vector{Float64},Vector{Float64}}, op, qω, qγ)
    out = zeros(prod(op.size))
    pbcheck(ccall((:pb200_ops_div, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}), b200_handle(op), qω, qγ, out)); out
end

# ---- closures evaluated where the reference evaluates them (src/solver.jl:230-323): constants stay scalars, nothing is uploaded for them --
b200_call(f, c, t) = t === nothing ? f(c...) : (applicable(f, c..., t) ? f(c..., t) : f(c...))
function const_or_array(f, coords, t = nothing)
    f isa Number && return (Float64(f), Float64[])
    pad3(c) = (c..., ntuple(_ -> 0.0, 3 - length(c))...)                       # the reference calls f(x, y, z[, t]) with zeros for absent dims
    v = Float64[b200_call(f, pad3(Tuple(c)), t) for c in coords]
    all(==(v[1]), v) ? (v[1], Float64[]) : (0.0, v)
end
ptr_or_null(a::Vector{Float64}) = isempty(a) ? Ptr{Cdouble}(C_NULL) : pointer(a)

# ---- border keys (classify_boundary_cell_fast, src/solver.jl:379-409): side index of the C ABI, dimension, low/high ------------------
const B200_SIDES = Dict(:left => (0, 2, false), :right => (1, 2, true), :bottom => (2, 1, false), :top => (3, 1, true),
                        :backward => (4, 3, false), :forward => (5, 3, true))
b200_bc_kind(::Dirichlet) = Cint(1); b200_bc_kind(::Neumann) = Cint(2); b200_bc_kind(::Robin) = Cint(3); b200_bc_kind(::Periodic) = Cint(4)
b200_bc_kind(::AbstractBoundary) = Cint(0)
"""BC_border_mono! / BC_border_diph! (src/solver.jl:417-580): one pb200_solver_set_border per key; unknown keys never match a cell and are skipped
like in the reference; values are evaluated at `mesh.centers` of the side's real cells (other dims, x fastest)."""
function set_borders!(h::Ptr{Cvoid}, mesh::Mesh{N}, bc_b::BorderConditions, t) where {N}
    for (key, cond) in bc_b.borders
        haskey(B200_SIDES, key) || continue
        side, dim, hi = B200_SIDES[key]
        dim > N && continue
        kind = b200_bc_kind(cond)
        val, arr = 0.0, Float64[]
        if kind == 1 || (kind == 2 && N == 1)
            v = cond.value
            if v isa Number
                val = Float64(v)
            else
                ax = [d == dim ? [hi ? mesh.centers[d][end] : mesh.centers[d][1]] : mesh.centers[d] for d in 1:N]
                arr = Float64[b200_call(v, (p..., ntuple(_ -> 0.0, 3 - N)...), t) for p in Iterators.product(ax...)][:]
                kind == 2 && (val = arr[1]; arr = Float64[])
            end
        end
        GC.@preserve arr pbcheck(ccall((:pb200_solver_set_border, libpb), Cint, (Ptr{Cvoid}, Cint, Cint, Cdouble, Ptr{Cdouble}), h, side, kind, val, ptr_or_null(arr)))
    end
end

# ---- structs of the step call, field for field as in the header ---------------------------------------------------------------------
struct pb200_solver_desc
    phase_type::Cint; time_type::Cint; ops1::Ptr{Cvoid}; ops2::Ptr{Cvoid}; D1::Cdouble; D2::Cdouble; D1_arr::Ptr{Cdouble}; D2_arr::Ptr{Cdouble}
    ifc_kind::Cint; alpha::Cdouble; beta::Cdouble; alpha1::Cdouble; alpha2::Cdouble; beta1::Cdouble; beta2::Cdouble
end
struct pb200_step_in
    scheme::Cint; dt::Cdouble
    f_const::NTuple{4,Cdouble}            # [phase][0: t_n, 1: t_n + dt], row-major
    f_arr::NTuple{4,Ptr{Cdouble}}
    g_const::NTuple{2,Cdouble}
    g_arr::NTuple{2,Ptr{Cdouble}}
end
struct pb200_krylov_opts
    method::Cint; rtol::Cdouble; atol::Cdouble; maxit::Cint; warm_start::Cint; check_every::Cint; path::Cint; precond::Cint
end
struct pb200_step_stats
    iters::Cint; converged::Cint; rnorm::Cdouble; bnorm::Cdouble; solve_ms::Cdouble; setup_ms::Cdouble
    dof_bulk::Int64; dof_ifc::Int64; launches::Int64; apply_ms::Cdouble; apply_launches::Int64
    apply_cells_uniform::Int64; apply_cells_general::Int64
    kernel_ms::NTuple{8,Cdouble}; kernel_launches::NTuple{8,Int64}
    apply_cells_fast::Int64; band_cells::Int64; band_rows::Int64
end
pb200_step_stats() = pb200_step_stats(0, 0, 0.0, 0.0, 0.0, 0.0, 0, 0, 0, 0.0, 0, 0, 0, ntuple(_ -> 0.0, 8), ntuple(_ -> 0, 8), 0, 0, 0)

"""`method` / `algorithm` / kwargs of solve_system! (src/solver.jl:158-188) -> pb200_krylov_opts.  The direct routes (`\\`, an `algorithm`) are
served by the iterative solver at a tight tolerance (same answer to 1e-12); IterativeSolvers' `reltol` / `abstol` / `maxiter` carry over."""
function krylov_opts(method, kwargs)
    kw = Dict{Symbol,Any}(kwargs)
    name = method === nothing ? "auto" : lowercase(string(nameof(method)))
    m = name == "cg" ? 1 : startswith(name, "bicgstab") ? 2 : 0
    direct = method === nothing || name == "\\"
    rtol = Float64(get(kw, :reltol, direct ? 1e-13 : sqrt(eps(Float64))))        # IterativeSolvers default reltol = sqrt(eps) (SURVEY B.3)
    pb200_krylov_opts(m, rtol, Float64(get(kw, :abstol, 0.0)), Int(get(kw, :maxiter, 20000)), Int(get(kw, :warm_start, 0)), 4, 0, get(kw, :precond, :default) == :mg ? 1 : 0)
end

# ---- solver construction shared by the four constructors (src/solver/diffusion.jl:14-28, 88-102, 192-210, 319-332) -------------------
const B200_FIRST = WeakKeyDict{Any,Any}()      # the constructor's step: (scheme, Δt, interface condition)
function b200_make_solver(time_type, phase_type, ph1::Phase, ph2, bc_i, ic)
    D1, D1a = const_or_array(ph1.Diffusion_coeff, ph1.capacity.C_ω)             # build_I_D (src/solver.jl:255-266), evaluated ONCE
    D2, D2a = ph2 === nothing ? (0.0, Float64[]) : const_or_array(ph2.Diffusion_coeff, ph2.capacity.C_ω)
    kind, al, be = Cint(0), 0.0, 0.0
    a1 = a2 = b1 = b2 = 0.0
    if ph2 === nothing                                                           # build_I_bc (src/solver.jl:203-223)
        bc_i isa Dirichlet ? (kind = Cint(1)) : bc_i isa Neumann ? (kind = Cint(2)) :
            bc_i isa Robin ? (kind = Cint(3); al = Float64(bc_i.α); be = Float64(bc_i.β)) : error("interface condition must be Dirichlet, Neumann or Robin")
    else
        a1, a2, b1, b2 = Float64(ic.scalar.α₁), Float64(ic.scalar.α₂), Float64(ic.flux.β₁), Float64(ic.flux.β₂)
    end
    d = pb200_solver_desc(ph2 === nothing ? 0 : 1, time_type, b200_handle(ph1.operator), ph2 === nothing ? C_NULL : b200_handle(ph2.operator),
                          D1, D2, ptr_or_null(D1a), ptr_or_null(D2a), kind, al, be, a1, a2, b1, b2)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve D1a D2a pbcheck(ccall((:pb200_solver_create, libpb), Cint, (Ptr{Cvoid}, Ref{pb200_solver_desc}, Ref{Ptr{Cvoid}}), b200_context(), Ref(d), h))
    h[]
end
b200_on_device(ph::Phase) = haskey(B200_HANDLES, ph.operator)

# one solve_system! of the reference loops: build b from the device state, solve, fetch the state
function b200_step!(s::Solver, ph1::Phase, ph2, bc_i, ic, scheme::String, Δt, t, opts::pb200_krylov_opts, unsteady::Bool)
    h = b200_handle(s)
    keep = Any[]
    fc, fa = zeros(4), fill(Ptr{Cdouble}(C_NULL), 4)
    for (k, ph) in enumerate((ph1, ph2))
        ph === nothing && continue
        for (w, tt) in enumerate(unsteady ? (t, t + Δt) : (nothing,))          # build_source at t_n and t_n + Δt (src/solver.jl:283-286)
            c, a = const_or_array(ph.source, ph.capacity.C_ω, tt)
            fc[2(k-1)+w] = c; isempty(a) || (push!(keep, a); fa[2(k-1)+w] = pointer(a))
        end
    end
    gc, ga = zeros(2), fill(Ptr{Cdouble}(C_NULL), 2)
    if ph2 === nothing                                                          # build_g_g (src/solver.jl:309-323): mono g(t_n), g(t_n + Δt)
        for (w, tt) in enumerate(unsteady ? (t, t + Δt) : (nothing,))
            c, a = const_or_array(bc_i.value, ph1.capacity.C_γ, tt)
            gc[w] = c; isempty(a) || (push!(keep, a); ga[w] = pointer(a))
        end
    else                                                                        # diph: [1] scalar-jump g at C_γ of phase 1, [2] flux-jump h at C_γ of phase 2, no t (diffusion.jl:397)
        for (w, (bc, cap)) in enumerate(((ic.scalar, ph1.capacity), (ic.flux, ph2.capacity)))
            c, a = const_or_array(bc.value, cap.C_γ)
            gc[w] = c; isempty(a) || (push!(keep, a); ga[w] = pointer(a))
        end
    end
    si = pb200_step_in(scheme == "CN" ? 1 : 0, unsteady ? Float64(Δt) : 0.0, Tuple(fc), Tuple(fa), Tuple(gc), Tuple(ga))
    st = Ref(pb200_step_stats())
    GC.@preserve keep pbcheck(ccall((:pb200_solver_step, libpb), Cint, (Ptr{Cvoid}, Ref{pb200_step_in}, Ref{pb200_krylov_opts}, Ref{pb200_step_stats}),
                                    h, Ref(si), Ref(opts), st); allow = (PB200_ENOTCONV,))
    nt = prod(ph1.operator.size)
    s.x = Vector{Float64}(undef, (ph2 === nothing ? 2 : 4) * nt)
    pbcheck(ccall((:pb200_solver_get_state, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), h, s.x))
    push!(s.ch, st[])
    st[]
end

# ---- steady ---------------------------------------------------------------------------------------------------------------------------
function DiffusionSteadyMono(phase::Phase, bc_b::BorderConditions, bc_i::AbstractBoundary)      # src/solver/diffusion.jl:14-28
    b200_on_device(phase) || return invoke(DiffusionSteadyMono, Tuple{Any,Any,Any}, phase, bc_b, bc_i)
    s = Solver(Steady, Monophasic, Diffusion, nothing, nothing, nothing, [], [])
    b200_register!(s, b200_make_solver(0, Monophasic, phase, nothing, bc_i, nothing), :pb200_solver_destroy)
    set_borders!(b200_handle(s), phase.capacity.mesh, bc_b, nothing)
    B200_FIRST[s] = (phase, nothing, bc_i, nothing); s
end
function solve_DiffusionSteadyMono!(s::Solver; method = nothing, algorithm = nothing, kwargs...)  # :60-72
    haskey(B200_HANDLES, s) || return invoke(solve_DiffusionSteadyMono!, Tuple{Any}, s; method = method, algorithm = algorithm, kwargs...)
    ph, _, bc_i, _ = B200_FIRST[s]
    b200_step!(s, ph, nothing, bc_i, nothing, "BE", 0.0, nothing, krylov_opts(method, kwargs), false); s
end
function DiffusionSteadyDiph(ph1::Phase, ph2::Phase, bc_b::BorderConditions, ic::InterfaceConditions)   # :88-102
    b200_on_device(ph1) || return invoke(DiffusionSteadyDiph, Tuple{Any,Any,Any,Any}, ph1, ph2, bc_b, ic)
    s = Solver(Steady, Diphasic, Diffusion, nothing, nothing, nothing, [], [])
    b200_register!(s, b200_make_solver(0, Diphasic, ph1, ph2, nothing, ic), :pb200_solver_destroy)
    set_borders!(b200_handle(s), ph1.capacity.mesh, bc_b, nothing)
    B200_FIRST[s] = (ph1, ph2, nothing, ic); s
end
function solve_DiffusionSteadyDiph!(s::Solver; method = nothing, algorithm = nothing, kwargs...)    # :163-175
    haskey(B200_HANDLES, s) || return invoke(solve_DiffusionSteadyDiph!, Tuple{Any}, s; method = method, algorithm = algorithm, kwargs...)
    ph1, ph2, _, ic = B200_FIRST[s]
    b200_step!(s, ph1, ph2, nothing, ic, "BE", 0.0, nothing, krylov_opts(method, kwargs), false); s
end

# ---- unsteady ---------------------------------------------------------------------------------------------------------------------------
function DiffusionUnsteadyMono(phase::Phase, bc_b::BorderConditions, bc_i::AbstractBoundary, Δt::Float64, Tᵢ::Vector{Float64}, scheme::String)   # :192-210
    b200_on_device(phase) || return invoke(DiffusionUnsteadyMono, Tuple{Any,Any,Any,Any,Any,Any}, phase, bc_b, bc_i, Δt, Tᵢ, scheme)
    s = Solver(Unsteady, Monophasic, Diffusion, nothing, nothing, nothing, [], [])
    b200_register!(s, b200_make_solver(1, Monophasic, phase, nothing, bc_i, nothing), :pb200_solver_destroy)
    pbcheck(ccall((:pb200_solver_set_state, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), b200_handle(s), Tᵢ))
    set_borders!(b200_handle(s), phase.capacity.mesh, bc_b, 0.0)
    B200_FIRST[s] = (scheme == "CN" ? "CN" : "BE", Δt, bc_i); s             # the constructor fixes the system of the FIRST solve (t = 0, ctor scheme)
end
function solve_DiffusionUnsteadyMono!(s::Solver, phase::Phase, Δt::Float64, Tₑ::Float64, bc_b::BorderConditions, bc::AbstractBoundary, scheme::String;
                                      method = nothing, algorithm = nothing, kwargs...)                                                                  # :268-301
    haskey(B200_HANDLES, s) || return invoke(solve_DiffusionUnsteadyMono!, Tuple{Any,Any,Any,Any,Any,Any,Any}, s, phase, Δt, Tₑ, bc_b, bc, scheme;
                                             method = method, algorithm = algorithm, kwargs...)
    opts = krylov_opts(method, kwargs)
    sch0, dt0, bc0 = B200_FIRST[s]
    b200_step!(s, phase, nothing, bc0, nothing, sch0, dt0, 0.0, opts, true)
    push!(s.states, s.x)
    println("Time : 0.0"); println("Max value : $(maximum(abs.(s.x)))")          # as the reference prints (:281-283)
    t = 0.0
    while t < Tₑ                                                                  # same floating-point accumulation => same number of solves
        t += Δt
        set_borders!(b200_handle(s), phase.capacity.mesh, bc_b, t)               # BC_border_mono! is called with t every step (:293): values only, masks stay
        b200_step!(s, phase, nothing, bc, nothing, scheme, Δt, t, opts, true)
        push!(s.states, s.x)
        println("Time : $(t)"); println("Max value : $(maximum(abs.(s.x)))")
    end
    s
end
function DiffusionUnsteadyDiph(ph1::Phase, ph2::Phase, bc_b::BorderConditions, ic::InterfaceConditions, Δt::Float64, Tᵢ::Vector{Float64}, scheme::String)  # :319-332
    b200_on_device(ph1) || return invoke(DiffusionUnsteadyDiph, Tuple{Any,Any,Any,Any,Any,Any,Any}, ph1, ph2, bc_b, ic, Δt, Tᵢ, scheme)
    s = Solver(Unsteady, Diphasic, Diffusion, nothing, nothing, nothing, [], [])
    b200_register!(s, b200_make_solver(1, Diphasic, ph1, ph2, nothing, ic), :pb200_solver_destroy)
    pbcheck(ccall((:pb200_solver_set_state, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), b200_handle(s), Tᵢ))
    set_borders!(b200_handle(s), ph1.capacity.mesh, bc_b, nothing)               # BC_border_diph! is called without t (:330)
    B200_FIRST[s] = (scheme, Δt, ic); s
end
function solve_DiffusionUnsteadyDiph!(s::Solver, ph1::Phase, ph2::Phase, Δt::Float64, Tₑ::Float64, bc_b::BorderConditions, ic::InterfaceConditions, scheme::String;
                                      method = nothing, algorithm = nothing, kwargs...)                                                                   # :422-454
    haskey(B200_HANDLES, s) || return invoke(solve_DiffusionUnsteadyDiph!, Tuple{Any,Any,Any,Any,Any,Any,Any,Any}, s, ph1, ph2, Δt, Tₑ, bc_b, ic, scheme;
                                             method = method, algorithm = algorithm, kwargs...)
    opts = krylov_opts(method, kwargs)
    sch0, dt0, ic0 = B200_FIRST[s]
    b200_step!(s, ph1, ph2, nothing, ic0, sch0, dt0, 0.0, opts, true)            # the constructor's system is solved first (:429-431)
    push!(s.states, s.x)
    println("Time : 0.0"); println("Max value : $(maximum(abs.(s.x)))")
    set_borders!(b200_handle(s), ph1.capacity.mesh, bc_b, nothing)
    t = 0.0
    while t < Tₑ
        t += Δt
        b200_step!(s, ph1, ph2, nothing, ic, scheme, Δt, t, opts, true)
        push!(s.states, s.x)
        println("Time : $(t)"); println("Max value : $(maximum(abs.(s.x)))")   # :448-449
    end
    s
end

# ---- advection-diffusion (src/operators.jl:194-209, src/solver/advectiondiffusion.jl:12-283): ConvectionOps + the monophasic solvers ------
# ConvectionOps(cap, uₒ, uᵧ): the diffusion operator handle gains the advective coefficient arrays on the device; C and K are never assembled
# (the struct keeps empty matrices, `size` and `V` stay as in the reference).
function ConvectionOps(cap::Capacity{N}, uₒ, uᵧ) where {N}
    haskey(B200_HANDLES, cap) || return invoke(ConvectionOps, Tuple{AbstractCapacity,Any,Any}, cap, uₒ, uᵧ)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    pbcheck(ccall((:pb200_ops_create, libpb), Cint, (Ptr{Cvoid}, Ptr{Ptr{Cvoid}}), b200_handle(cap), h))
    uo = Float64.(vcat((vec(u) for u in uₒ)...)); ug = Float64.(vec(uᵧ))
    pbcheck(ccall((:pb200_ops_set_convection, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}), h[], uo, ug))
    n = prod(length.(cap.mesh.nodes)); Z = spzeros(n, n)
    op = ConvectionOps{N}(ntuple(_ -> Z, N), ntuple(_ -> Z, N), spzeros(N * n, n), spzeros(N * n, n), spzeros(N * n, N * n), cap.V, ntuple(i -> length(cap.mesh.nodes[i]), N))
    b200_register!(op, h[], :pb200_ops_destroy); op
end
function AdvectionDiffusionSteadyMono(phase::Phase, bc_b::BorderConditions, bc_i::AbstractBoundary)              # advectiondiffusion.jl:12-28
    b200_on_device(phase) || return invoke(AdvectionDiffusionSteadyMono, Tuple{Any,Any,Any}, phase, bc_b, bc_i)
    s = Solver(Steady, Monophasic, DiffusionAdvection, nothing, nothing, nothing, [], [])
    b200_register!(s, b200_make_solver(0, Monophasic, phase, nothing, bc_i, nothing), :pb200_solver_destroy)
    set_borders!(b200_handle(s), phase.capacity.mesh, bc_b, nothing)
    B200_FIRST[s] = (phase, nothing, bc_i, nothing); s
end
function solve_AdvectionDiffusionSteadyMono!(s::Solver; method = nothing, algorithm = nothing, kwargs...)          # :65-71 (gmres there, BiCGSTAB on the device)
    haskey(B200_HANDLES, s) || return invoke(solve_AdvectionDiffusionSteadyMono!, Tuple{Any}, s; method = method, algorithm = algorithm, kwargs...)
    ph, _, bc_i, _ = B200_FIRST[s]
    b200_step!(s, ph, nothing, bc_i, nothing, "BE", 0.0, nothing, krylov_opts(nothing, kwargs), false); s
end
function AdvectionDiffusionUnsteadyMono(phase::Phase, bc_b::BorderConditions, bc_i::AbstractBoundary, Δt::Float64, Tᵢ::Vector{Float64}, scheme::String)   # :163-176
    b200_on_device(phase) || return invoke(AdvectionDiffusionUnsteadyMono, Tuple{Any,Any,Any,Any,Any,Any}, phase, bc_b, bc_i, Δt, Tᵢ, scheme)
    s = Solver(Unsteady, Monophasic, DiffusionAdvection, nothing, nothing, nothing, [], [])
    b200_register!(s, b200_make_solver(1, Monophasic, phase, nothing, bc_i, nothing), :pb200_solver_destroy)
    pbcheck(ccall((:pb200_solver_set_state, libpb), Cint, (Ptr{Cvoid}, Ptr{Cdouble}), b200_handle(s), Tᵢ))
    B200_FIRST[s] = (scheme == "CN" ? "CN" : "BE", Δt, bc_i); s                   # NO border rows in the constructor's system (the reference applies none, :172-176)
end
function solve_AdvectionDiffusionUnsteadyMono!(s::Solver, phase::Phase, Δt::Float64, Tₑ, bc_b::BorderConditions, bc::AbstractBoundary, scheme::String;
                                               method = nothing, algorithm = nothing, kwargs...)                                                       # :254-283
    haskey(B200_HANDLES, s) || return invoke(solve_AdvectionDiffusionUnsteadyMono!, Tuple{Any,Any,Any,Any,Any,Any,Any}, s, phase, Δt, Tₑ, bc_b, bc, scheme;
                                             method = method, algorithm = algorithm, kwargs...)
    opts = krylov_opts(nothing, kwargs)
    sch0, dt0, bc0 = B200_FIRST[s]
    b200_step!(s, phase, nothing, bc0, nothing, sch0, dt0, 0.0, opts, true)
    push!(s.states, s.x); println("Time: 0.0"); println("Solver Extremum: ", maximum(abs.(s.x)))
    t = 0.0
    while t < Tₑ
        t += Δt
        set_borders!(b200_handle(s), phase.capacity.mesh, bc_b, t)               # BC_border_mono!(…; t = t) (:273)
        b200_step!(s, phase, nothing, bc, nothing, scheme, Δt, t, opts, true)   # (the reference's RHS call at :272 drops the diffusion coefficient: a MethodError as written)
        push!(s.states, s.x); println("Time: ", t); println("Solver Extremum: ", maximum(abs.(s.x)))
    end
    s
end

# ---- Darcy (src/solver/darcy.jl:1-89): the diffusion systems under another name + the velocity u = -∇p -------------------------------
DarcyFlow(phase::Phase, bc_b::BorderConditions, bc_i::AbstractBoundary) = DiffusionSteadyMono(phase, bc_b, bc_i)
solve_DarcyFlow!(s::Solver; kw...) = (solve_DiffusionSteadyMono!(s; kw...); push!(s.states, s.x); s)
DarcyFlowUnsteady(phase::Phase, bc_b, bc_i, Δt, Tᵢ, scheme) = DiffusionUnsteadyMono(phase, bc_b, bc_i, Δt, Tᵢ, scheme)
solve_DarcyFlowUnsteady!(s::Solver, phase, Δt, Tₑ, bc_b, bc_i, scheme; kw...) = solve_DiffusionUnsteadyMono!(s, phase, Δt, Tₑ, bc_b, bc_i, scheme; kw...)
function solve_darcy_velocity(solver::Solver, Fluide::Phase; state_i = 1)          # :26-40: NaN masking on the host, ∇ on the device
    ct = Fluide.capacity.cell_types
    p = copy(solver.states[state_i]); n = length(p) ÷ 2
    pω, pγ = view(p, 1:n), view(p, n+1:2n)
    pω[ct .== 0] .= NaN; pγ[ct .== 0] .= NaN; pγ[ct .== 1] .= NaN
    -∇(Fluide.operator, p)                          # NaN * (stored zero) = NaN as in SparseArrays: the kernel multiplies every stencil entry
end

# ---- check_convergence(u_analytical, solver, capacity, p = 2, relative = false)   replaces src/convergence.jl:59-93 ------------------
function check_convergence(u_analytical::Function, solver::Solver, capacity::Capacity{N}, p::Real = 2, relative::Bool = false) where {N}
    haskey(B200_HANDLES, solver) || return invoke(check_convergence, Tuple{Function,Any,Any,Real,Bool}, u_analytical, solver, capacity, p, relative)
    u_ana = Float64[u_analytical(c...) for c in capacity.C_ω]                       # host closure at C_ω, as the reference (:63-71)
    out = zeros(4)                                                                  # all fluid (full + cut) / full / cut / empty
    pbcheck(ccall((:pb200_solver_error_norms, libpb), Cint, (Ptr{Cvoid}, Cint, Ptr{Cdouble}, Cdouble, Cint, Ptr{Cdouble}),
                  b200_handle(solver), 0, u_ana, Float64(p), relative, out))        # ONE fused reduction over the DEVICE state
    println("All cells L$p norm        = $(out[1])"); println("Full cells L$p norm   = $(out[2])")
    println("Cut cells L$p norm    = $(out[3])");     println("Empty cells L$p norm  = $(out[4])")
    (u_ana, solver.x[1:end÷2], out[1], out[2], out[3], out[4])
end
