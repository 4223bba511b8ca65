# EBMCUDAExt.jl -- package extension of EnergyBalanceModel.jl that runs `integrate` for whole ensembles on an
# NVIDIA B200 through libebm_cuda.so (include/ebm_cuda.h).  Thin `ccall` layer only: no CUDA.jl, no kernels in
# Julia, no CPU fallback (a missing library or device throws).
#
# NOT EXECUTED IN THE BUILD ENVIRONMENT: Julia is not installed there (DESIGN.md section 1).  The ABI it binds is
# exercised by the ctypes mirror under energybalancemodel.jl_b200/ and tests/.
#
# Wiring (same mechanism as ext/CairoExt.jl + Project.toml [weakdeps]/[extensions], see INTEGRATION.md):
#   [weakdeps]    Libdl = "8f399da3-3557-5675-b5ff-fb832c97cbdb"
#   [extensions]  EBMCUDAExt = "Libdl"
# so `using EnergyBalanceModel, Libdl` activates it; the library path comes from ENV["EBM_CUDA_LIB"].
module EBMCUDAExt

import EnergyBalanceModel as EBM
import EnergyBalanceModel.Infrastructure: SpaceTime, Forcing, Collection, Solutions, Vec, integrate
import Libdl

const CLASSIC_PAR = (:D, :A, :B, :cw, :S0, :S1, :S2, :a0, :a2, :ai, :Fb, :k, :Lf, :cg, :tau)   # ebm_classic_params_t
const MIZ_PAR = (:D, :A, :B, :cw, :S0, :S1, :S2, :a0, :a2, :ai, :Fb, :k, :Lf, :Tm, :m1, :m2, :alpha, :rl,
                 :Dmin, :Dmax, :hmin, :kappa)                                                  # ebm_miz_params_t
const CLASSIC_VARS = (:E, :T, :h)
const MIZ_VARS = (:T, :Ei, :Ti, :D, :n, :h, :phi, :E, :Ew, :Tw)                                # EBM_MV_* order

const LIB = Ref{Ptr{Cvoid}}(C_NULL)
function lib()
    if LIB[] == C_NULL
        path = get(ENV, "EBM_CUDA_LIB", "libebm_cuda.so")
        LIB[] = Libdl.dlopen(path)            # throws if the library is missing: there is no CPU fallback
    end
    return LIB[]
end
sym(name::Symbol) = Libdl.dlsym(lib(), name)

# mirrors of the C structs (include/ebm_cuda.h); all isbits
struct CGrid
    nx::Int32; nt::Int32; dur::Int32; grid_kind::Int32; winter_inx::Int32; summer_inx::Int32
    x::Ptr{Float64}; t::Ptr{Float64}
end
struct COptions
    device::Int32; lastonly::Int32; field_stride::Int32; strict::Int32; years_per_launch::Int32
    newton_maxit::Int32; newton_tol::Float64; step_limit::Int32; start_year::Int32; classic_stencil::Int32
end
struct CMulti          # ebm_multi_t
    ndevices::Int32; diag_device::Int32; packet::Int32; reserved::Int32; devices::Ptr{Int32}
end
struct CClassicOutputs
    diag::Ptr{Float64}; seasonal::Ptr{Float64}; raw::Ptr{Float64}; E_final::Ptr{Float64}; Tg_final::Ptr{Float64}
    flags::Ptr{Int32}
end
struct CMizOutputs
    diag::Ptr{Float64}; seasonal::Ptr{Float64}; raw::Ptr{Float64}
    Ei::Ptr{Float64}; Ew::Ptr{Float64}; h::Ptr{Float64}; D::Ptr{Float64}; phi::Ptr{Float64}; T0::Ptr{Float64}
    newton_iters::Ptr{Int64}; nonconv::Ptr{Int64}; flags::Ptr{Int32}
end

check(rc::Int32) = rc == 0 ? nothing :
    (msg = unsafe_string(ccall(sym(:ebm_last_error), Cstring, ()));
     rc == -1 ? throw(ArgumentError(msg)) : error("libebm_cuda error $rc: $msg"))

grid_kind(::SpaceTime{identity}) = Int32(0)      # get_diffop path (infrastructure.jl:495-497)
grid_kind(::SpaceTime) = Int32(1)                # generic flux-form stencil (infrastructure.jl:500-527)

# Forcing -> 10 doubles: base, peak, cool, rate_up, rate_down, domain[1:5]
forcing_row(f::Forcing) = Float64[f.base, f.peak, f.cool, f.rates[1], f.rates[2], Float64.(f.domain)...]
par_rows(pars, names) = permutedims(Float64[getproperty(p, n) for p in pars, n in names])   # [npar × nmem], i.e. member-major rows in C
state_rows(inits, name, nx) = (a = reduce(hcat, [getproperty(i, name) for i in inits]); size(a, 1) == nx ||
                               throw(ArgumentError("init.$name must have length nx=$nx")); a)  # [nx × nmem] column-major == [nmem][nx] in C

"""
    integrate(model, st, forcings::Vector{<:Forcing}, pars::Vector{Collection{Float64}}, inits::Vector{Collection{Vec}};
              lastonly=true, field_stride=1, debug=nothing, verbose=false) -> (Vector{Solutions}, NamedTuple)

Ensemble form of `Infrastructure.integrate` (src/infrastructure.jl:615-636): one member per entry of the three
vectors, integrated on the GPU.  `Solutions` are rebuilt for every `field_stride`-th member; the NamedTuple carries
the per-member-year scalar diagnostics `diag[4, 3, dur, nmem]`, final states, flags and (MIZ) closure statistics.
"""
function integrate(model::Symbol, st::SpaceTime{F}, forcings::AbstractVector{<:Forcing},
                   pars::AbstractVector{Collection{Float64}}, inits::AbstractVector{Collection{Vec}};
                   lastonly::Bool=true, field_stride::Int=1, debug::Union{Expr,Nothing}=nothing,
                   verbose::Bool=false, device::Int=-1, devices::Union{Nothing,AbstractVector{<:Integer}}=nothing,
                   classic_stencil::Int=0) where F   # 1: generic flux-form stencil in kappa (classic on non-uniform grids)
    isnothing(debug) || throw(ArgumentError("`debug::Expr` cannot be evaluated on the device"))
    model in (:Classic, :MIZ) || throw(MethodError(EBM.Infrastructure.step!, (Val(model),)))
    nmem = length(pars)
    (nmem > 0 && length(forcings) == nmem == length(inits)) || throw(ArgumentError("forcings, pars, inits must have equal non-zero length"))
    # devices = [0, 1, ...]: several GPUs behind one call (ebm_*_run_multi: one host thread + stream per GPU inside
    # the library, members dealt in 32-member packets after a sort by cost); every output as in the single-GPU call
    multi = !isnothing(devices)
    devs = multi ? Int32.(collect(devices)) : Int32[]
    nx, nt, dur = st.nx, st.nt, st.dur
    nsel = field_stride > 0 ? cld(nmem, field_stride) : 0
    nraw = lastonly ? nt : nt * dur
    vars = model === :MIZ ? MIZ_VARS : CLASSIC_VARS
    nvar = length(vars)
    forc = reduce(hcat, forcing_row.(forcings))                      # [10 × nmem]
    diag = fill(NaN, 4, 3, dur, nmem)
    seasonal = fill(NaN, nx, nvar, 3, dur, max(nsel, 1))
    raw = fill(NaN, nx, nvar, nraw, max(nsel, 1))
    flags = zeros(Int32, nmem)
    x, t = st.x, st.t
    opt = Ref(COptions(device, lastonly, field_stride, 0, 0, 0, 0.0, 0, 0, classic_stencil))
    local final, stats
    GC.@preserve x t forc diag seasonal raw flags devs begin
        mopt = Ref(CMulti(length(devs), -1, 0, 0, multi ? pointer(devs) : Ptr{Int32}(C_NULL)))
        grid = Ref(CGrid(nx, nt, dur, grid_kind(st), st.winter.inx, st.summer.inx, pointer(x), pointer(t)))
        sp = nsel > 0 ? pointer(seasonal) : Ptr{Float64}(C_NULL)
        rp = nsel > 0 ? pointer(raw) : Ptr{Float64}(C_NULL)
        if model === :Classic
            par = par_rows(pars, CLASSIC_PAR)
            E0, Tg0 = state_rows(inits, :E, nx), state_rows(inits, :Tg, nx)
            Ef, Tgf = similar(E0), similar(Tg0)
            GC.@preserve par E0 Tg0 Ef Tgf begin
                out = Ref(CClassicOutputs(pointer(diag), sp, rp, pointer(Ef), pointer(Tgf), pointer(flags)))
                if multi
                    check(ccall(sym(:ebm_classic_run_multi), Int32,
                                (Ref{CGrid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{COptions}, Ref{CMulti}, Ref{CClassicOutputs}),
                                grid, nmem, par, forc, E0, Tg0, opt, mopt, out))
                else
                    check(ccall(sym(:ebm_classic_run), Int32,
                                (Ref{CGrid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{COptions}, Ref{CClassicOutputs}),
                                grid, nmem, par, forc, E0, Tg0, opt, out))
                end
            end
            final, stats = (E=Ef, Tg=Tgf), (;)
        else
            par = par_rows(pars, MIZ_PAR)
            s0 = [state_rows(inits, k, nx) for k in (:Ei, :Ew, :h, :D, :phi)]
            sf = [similar(s0[1]) for _ in 1:6]
            iters, nonconv = zeros(Int64, nmem), zeros(Int64, nmem)
            GC.@preserve par s0 sf iters nonconv begin
                out = Ref(CMizOutputs(pointer(diag), sp, rp, pointer.(sf)..., pointer(iters), pointer(nonconv), pointer(flags)))
                if multi
                    check(ccall(sym(:ebm_miz_run_multi), Int32,
                                (Ref{CGrid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                                 Ptr{Float64}, Ptr{Float64}, Ref{COptions}, Ref{CMulti}, Ref{CMizOutputs}),
                                grid, nmem, par, forc, s0[1], s0[2], s0[3], s0[4], s0[5], C_NULL, opt, mopt, out))
                else
                    check(ccall(sym(:ebm_miz_run), Int32,
                                (Ref{CGrid}, Int64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
                                 Ptr{Float64}, Ptr{Float64}, Ref{COptions}, Ref{CMizOutputs}),
                                grid, nmem, par, forc, s0[1], s0[2], s0[3], s0[4], s0[5], C_NULL, opt, out))
                end
            end
            verbose && any(>(0), nonconv) && @warn "Solving for T0 failed at $(sum(nonconv)) member-steps."   # miz.jl:61-63
            final = (Ei=sf[1], Ew=sf[2], h=sf[3], D=sf[4], phi=sf[5], T0=sf[6])
            stats = (newton_iters=iters, nonconv=nonconv)
        end
    end
    # rebuild Solutions{F,C} (infrastructure.jl:333-383): raw.E[ti] == column ti of the returned matrices
    sols = Solutions[]
    for k in 1:nsel
        m = (k - 1) * field_stride + 1
        s = Solutions(st, forcings[m], pars[m], inits[m], Set{Symbol}(vars), lastonly)
        for (vi, v) in enumerate(vars)
            getproperty(s.raw, v) .= [raw[:, vi, ti, k] for ti in 1:nraw]
            for (si, season) in enumerate((:winter, :summer, :avg))
                getproperty(getproperty(s.seasonal, season), v) .= [seasonal[:, vi, si, y, k] for y in 1:dur]
            end
        end
        push!(sols, s)
    end
    return sols, (; diag, final, flags, stats...)
end

"""
    integrate(model, sts::Vector{<:SpaceTime}, forcings, pars, inits; kwargs...) -> Vector{Tuple{Vector{Solutions},NamedTuple}}, where

Members on different grids (per-member `nx`, `nt`, `dur`, grid function; SURVEY 8f-4): `sts[m]` is member m's SpaceTime.
Members are grouped by SpaceTime (one kernel launch integrates one grid), each group is one ensemble call; the second
return value maps member m to `(group, row)`.  A member's result is bit-identical to integrating it alone.
"""
function integrate(model::Symbol, sts::AbstractVector{<:SpaceTime}, forcings::AbstractVector{<:Forcing},
                   pars::AbstractVector{Collection{Float64}}, inits::AbstractVector{Collection{Vec}}; kwargs...)
    nmem = length(sts)
    (nmem > 0 && length(forcings) == nmem == length(pars) == length(inits)) ||
        throw(ArgumentError("sts, forcings, pars, inits must have equal non-zero length"))
    keyof(st) = (st.nx, st.nt, st.dur, grid_kind(st), typeof(st))
    order = Vector{Tuple{Any,Vector{Int}}}()
    slot = Dict{Any,Int}()
    for (m, st) in enumerate(sts)
        k = keyof(st)
        haskey(slot, k) || (push!(order, (st, Int[])); slot[k] = length(order))
        push!(order[slot[k]][2], m)
    end
    where = Vector{Tuple{Int,Int}}(undef, nmem)
    results = map(enumerate(order)) do (g, (st, idx))
        for (r, m) in enumerate(idx); where[m] = (g, r); end
        integrate(model, st, forcings[idx], pars[idx], inits[idx]; kwargs...)
    end
    return results, where
end

# single member on the GPU: same signature as the reference plus a trailing Val(:CUDA) so the CPU method stays reachable
function integrate(model::Symbol, st::SpaceTime, forcing::Forcing, par::Collection{Float64}, init::Collection{Vec},
                   ::Val{:CUDA}; kwargs...)
    sols, _ = integrate(model, st, [forcing], [par], [init]; field_stride=1, kwargs...)
    return sols[1]
end

end # module EBMCUDAExt
