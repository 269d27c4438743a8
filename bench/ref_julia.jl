# bench/ref_julia.jl — the TRUE reference number for BASELINE.json's metric, to be run wherever Julia ≥ 1.11 and Muscade 0.7.0 are installed
# (there is no Julia in the build image, so bench.py's reference arm times a C++ port of the algorithm instead and says so: cpu_baseline.kind = "port").
#
#   JULIA_NUM_THREADS=1 julia --project bench/ref_julia.jl [N]          (N elements, default 100_000)
#
# Builds the synthetic chain of SURVEY.md §8d exactly as muscade.jl_b200/synthetic.py does — nodes p_k = k·(0.8,0.6,0), orient2 = (0,1,0), material of
# inspect/PerformanceEulerBeam3D.jl:17, state from splitmix64 — and times the reference's own serial `assemble!{:iter}` (src/Assemble.jl:470-487) for
# SweepX{0} and SweepX{2}, one element-assembly = one residual + 12×12 tangent + scatter.  Template: inspect/PerformanceEulerBeam3D.jl:35-37.
using Muscade, Muscade.Toolbox, BenchmarkTools, StaticArrays, Printf

N = length(ARGS) ≥ 1 ? parse(Int, ARGS[1]) : 100_000

function splitmix64(seed::UInt64, counter::UInt64)
    z = seed + counter * 0x9E3779B97F4A7C15
    z = (z ⊻ (z >> 30)) * 0xBF58476D1CE4E5B9
    z = (z ⊻ (z >> 27)) * 0x94D049BB133111EB
    return z ⊻ (z >> 31)
end
uniform_pm1(seed, i) = Float64(splitmix64(UInt64(seed), UInt64(i)) >> 11) * (2.0 / 2^53) - 1.0     # dof index i 0-based, as synthetic.uniform_pm1

function chain(N; dynamic=false)
    model = Model(:chain)
    nod   = [addnode!(model, k .* SVector(0.8, 0.6, 0.0)) for k = 0:N]
    mesh  = hcat(nod[1:N], nod[2:N+1])
    mat   = dynamic ? BeamCrossSection(EA=10., EI₂=3., EI₃=3., GJ=4., μ=1., ι₁=1., Ca₂=169.6, Ca₃=169.6, Cq₂=235.2, Cq₃=235.2) :
                      BeamCrossSection(EA=10., EI₂=3., EI₃=3., GJ=4., μ=1., ι₁=1.)
    addelement!(model, EulerBeam3D, mesh; mat, orient2=SVector(0., 1., 0.))
    return model
end

for (OX, dynamic) ∈ ((0, false), (2, true))
    model  = chain(N; dynamic)
    state  = initialize!(model)
    dis    = state.dis
    ndof   = Muscade.getndof(model, :X)
    state  = Muscade.State{1,OX+1,1}(copy(state))
    for d = 0:OX, i = 0:ndof-1
        u = uniform_pm1(0x5EED + d, i)
        state.X[d+1][i+1] = d == 0 ? u * (i % 6 < 3 ? 0.05 : 0.1) : 0.1u
    end
    out, asm, Xdofgr = Muscade.prepare(Muscade.AssemblySweepX{OX}, model, dis)
    out.c = Muscade.Newmarkβcoefficients{OX}(0.3, 1 / 4, 1 / 2)
    t = @belapsed Muscade.assemble!{:iter}($out, $asm, $dis, $model, $state, 0.3, (;))
    @printf("{\"impl\": \"muscade.jl\", \"metric\": \"EulerBeam3D residual+Jacobian element-assemblies/s\", \"OX\": %d, \"elements\": %d, \"value\": %.6g, \"unit\": \"element-assemblies/s\", \"cores\": 1, \"ms_per_step\": %.6g}\n",
            OX, N, N / t, 1e3t)
end
