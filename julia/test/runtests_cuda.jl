# runtests_cuda.jl -- the reference's own test (test/runtests.jl:20-48) run through the CUDA extension, plus the
# ensemble form.  NOT EXECUTED IN THE BUILD ENVIRONMENT (no Julia there); the same checks run from Python in
# tests/test_miz_gpu.py and tests/test_classic_gpu.py against the C oracle.
#
#   EBM_CUDA_LIB=/path/to/libebm_cuda.so julia --project=. julia/test/runtests_cuda.jl
using EnergyBalanceModel, Libdl, Test
import EnergyBalanceModel: Vec

@testset "CUDA extension vs the CPU path" begin
    # --- the reference fixture's setup (test/runtests.jl:22-32): MIZ, sin grid, zero init, one year
    st = SpaceTime{sin}(180, 2000, 1)
    forcing = Forcing(0.0)
    par = default_parameters(:MIZ)
    init = Collection{Vec}(
        :Ei => zeros(Float64, st.nx), :Ew => zeros(Float64, st.nx), :h => zeros(Float64, st.nx),
        :D => zeros(Float64, st.nx), :phi => zeros(Float64, st.nx)
    )
    cpu = integrate(:MIZ, st, forcing, par, init; lastonly=false)
    gpu = integrate(:MIZ, st, forcing, par, init, Val(:CUDA); lastonly=false)
    for var in (:T, :Ei, :Ti, :D, :n, :h, :phi, :E, :Ew, :Tw)
        a = copy(getproperty(cpu.raw, var)[10]); b = copy(getproperty(gpu.raw, var)[10])
        @test isnan.(a) == isnan.(b)
        a[isnan.(a)] .= 0.0; b[isnan.(b)] .= 0.0            # test/runtests.jl:42-43
        @test all(isapprox.(a, b))                           # rtol = sqrt(eps), the reference's own criterion
    end
    # the MIZ model amplifies rounding differences (DESIGN.md): compare a short horizon only
    for ti in 1:20, var in (:E, :T, :phi)
        a = copy(getproperty(cpu.raw, var)[ti]); b = copy(getproperty(gpu.raw, var)[ti])
        @test all(abs.(a .- b) .<= 1.5e-8 .* max.(abs.(a), 1.0))
    end

    # --- classic: a 30-year spin-up is strongly contracting, so the whole stored solution must agree to 1e-9
    stc = SpaceTime(100, 2000, 30)
    parc = default_parameters(:Classic)
    initc = Collection{Vec}(:E => fill(98.0, 100), :Tg => fill(10.0, 100))
    cpuc = integrate(:Classic, stc, forcing, parc, initc)
    gpuc = integrate(:Classic, stc, forcing, parc, initc, Val(:CUDA))
    for var in (:E, :T, :h), ti in (1, 522, 1548, 2000)
        a = getproperty(cpuc.raw, var)[ti]; b = getproperty(gpuc.raw, var)[ti]
        @test all(abs.(a .- b) .<= 1e-9 .* max.(abs.(a), 1.0))
    end
    @test all(abs.(cpuc.seasonal.avg.T[30] .- gpuc.seasonal.avg.T[30]) .<= 1e-9 .* max.(abs.(cpuc.seasonal.avg.T[30]), 1.0))

    # --- ensemble form: 64 members, forcing sweep; every 16th member comes back as a Solutions
    n = 64
    forcs = [Forcing(-10.0 + 20.0 * (m - 1) / (n - 1)) for m in 1:n]
    sols, ens = integrate(:Classic, SpaceTime(100, 2000, 2), forcs, fill(parc, n), fill(initc, n); field_stride=16)
    @test length(sols) == 4 && size(ens.diag) == (4, 3, 2, n) && all(ens.flags .== 0)
    @test_throws ArgumentError integrate(:Classic, stc, forcs, fill(parc, n), fill(initc, n); debug=:(vars.E))

    # --- members on different grids: grouped by SpaceTime, addressed by the caller's member index
    sts = [SpaceTime(100, 2000, 1), SpaceTime(60, 1000, 2), SpaceTime(100, 2000, 1)]
    inits3 = [Collection{Vec}(:E => fill(98.0, st.nx), :Tg => fill(10.0, st.nx)) for st in sts]
    groups, where = integrate(:Classic, sts, forcs[1:3], fill(parc, 3), inits3; field_stride=1)
    @test length(groups) == 2 && where == [(1, 1), (2, 1), (1, 2)]
    for (m, st) in enumerate(sts)
        g, r = where[m]
        alone = integrate(:Classic, st, forcs[m], parc, inits3[m], Val(:CUDA))
        @test groups[g][1][r].raw.E[end] == alone.raw.E[end]          # bit-identical to integrating the member alone
    end
end
