# MuscadeB200.jl — the reference-side binding of libmuscade_b200.so (include/muscade_b200.h).
#
# WRITE-ONLY in this repository: there is no Julia in the build image, so this file is not executed by the test-suite.
# It is the thin wrapper a Muscade maintainer adds next to src/SweepX.jl; its logic is mirrored 1:1 by the tested
# Python host (muscade.jl_b200/sweepx.py, engine.py), call for call.
#
# What it does: defines `AssemblySweepXB200{OX} <: Assembly`, a `prepare` method that hands the model's own arrays to the
# shim once, and an `assemble!{mission}` method that replaces the element loop of src/Assemble.jl:470-487 by one ccall.
# `solve(SweepX{OX};…)` (src/SweepX.jl:179-226) then runs unchanged if its `prepare(AssemblySweepX{OX},…)` call is
# pointed at `AssemblySweepXB200{OX}`.
module MuscadeB200

using Muscade, SparseArrays, LinearAlgebra, Printf
using Muscade: Assembly, assemble!, zero!, Newmarkβcoefficients, allXdofs, getndof, getneletyp, muscadeerror, 𝕣, 𝕫

const LIB = get(ENV, "MUSCADE_B200_LIB", "libmuscade_b200")

struct ErrInfo
    kind    :: Int32
    ieletyp :: Int32
    iele    :: Int64
    step    :: Int64
end

check(h, rc, dbg=(;)) = rc == 0 || muscadeerror(dbg, unsafe_string(ccall((:mb_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

mutable struct AssemblySweepXB200{OX,Tλ,Tλx} <: Assembly
    Lλ      :: Tλ                                 # same fields as AssemblySweepX (src/SweepX.jl:17-23) …
    Lλx     :: Tλx
    c       :: Newmarkβcoefficients{OX}
    h       :: Ptr{Cvoid}                         # … plus the device handle
    hosttyp :: Vector{Int}                        # element types evaluated on the host (closures: DofLoad, Hold, costs)
end

"""
    out,asm,Xdofgr = prepare(AssemblySweepXB200{OX},model,dis)

Replaces `prepare(AssemblySweepX{OX},model,dis)` (src/SweepX.jl:24-33).  The sparsity pattern and the index maps are built on the
device, bit-identical to `asmvec!`/`asmmat!` (src/Assemble.jl:340-448); `asm` is only fetched when host-evaluated types need it.
"""
function Muscade.prepare(::Type{AssemblySweepXB200{OX}}, model, dis; device=0) where {OX}
    Xdofgr  = allXdofs(model, dis)
    ndof    = getndof(Xdofgr)
    href    = Ref{Ptr{Cvoid}}()
    check(C_NULL, ccall((:mb_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, href))
    h       = href[]
    hosttyp = Int[]
    for ieletyp = 1:getneletyp(model)
        eleobj = model.eleobj[ieletyp]
        d      = dis.dis[ieletyp]
        E      = eltype(eleobj)
        # dis.dis[ieletyp].index is a Vector of isbits XUA{Int64,nX,nU,nA}: reinterpret gives the nX×nele Int64 matrix the ABI wants
        idxX   = [d.index[iele].X[i] for i = 1:length(d.scale.X), iele = 1:length(eleobj)]
        ityp   = Ref{Int32}()
        if E <: Muscade.Toolbox.EulerBeam3D{Muscade.Toolbox.BeamCrossSection}
            udof = length(d.scale.U) > 0
            idxU = udof ? [d.index[iele].U[i] for i = 1:3, iele = 1:length(eleobj)] : Matrix{Int64}(undef, 0, 0)
            GC.@preserve eleobj idxX idxU check(h, ccall((:mb_add_eulerbeam3d, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
                h, length(eleobj), pointer(reinterpret(Float64, eleobj)), udof, idxX, udof ? pointer(idxU) : C_NULL,
                collect(d.scale.X), udof ? collect(d.scale.U) : C_NULL, ityp))
        elseif E <: Muscade.Toolbox.Bar3D{Muscade.Toolbox.AxisymmetricBarCrossSection}
            udof = length(d.scale.U) > 0
            idxU = udof ? [d.index[iele].U[i] for i = 1:3, iele = 1:length(eleobj)] : Matrix{Int64}(undef, 0, 0)
            GC.@preserve eleobj idxX idxU check(h, ccall((:mb_add_bar3d, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
                h, length(eleobj), pointer(reinterpret(Float64, eleobj)), udof, idxX, udof ? pointer(idxU) : C_NULL,
                collect(d.scale.X), udof ? collect(d.scale.U) : C_NULL, ityp))
        elseif E <: Muscade.Toolbox.SoilContact
            GC.@preserve eleobj idxX check(h, ccall((:mb_add_soilcontact, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ref{Int32}),
                h, length(eleobj), pointer(reinterpret(Float64, eleobj)), idxX, collect(d.scale.X), ityp))
        else   # anything else keeps running through Muscade's own addin! on the host and is merged as dense element contributions
            check(h, ccall((:mb_add_host_elements, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Int64}, Ref{Int32}),
                h, length(eleobj), size(idxX, 1), idxX, ityp))
            push!(hosttyp, ieletyp)
        end
    end
    nnz = Ref{Int64}()
    check(h, ccall((:mb_sweepx_prepare, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}), h, ndof, nnz))
    colptr, rowval = Vector{Int64}(undef, ndof + 1), Vector{Int64}(undef, nnz[])
    check(h, ccall((:mb_sweepx_get_pattern, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}), h, colptr, rowval))
    Lλ   = Vector{𝕣}(undef, ndof)
    Lλx  = SparseMatrixCSC(ndof, ndof, colptr, rowval, ones(𝕣, nnz[]))
    out  = AssemblySweepXB200{OX,typeof(Lλ),typeof(Lλx)}(Lλ, Lλx, Newmarkβcoefficients{OX}(), h, hosttyp)
    finalizer(o -> ccall((:mb_destroy, LIB), Int32, (Ptr{Cvoid},), o.h), out)
    asm  = nothing      # fetched with mb_sweepx_get_asm on demand
    return out, asm, Xdofgr
end

"""
    assemble!{mission}(out::AssemblySweepXB200{OX},asm,dis,model,state,Δt,dbg)

Replaces the loops of src/Assemble.jl:470-487 for this assembly type: one `ccall` per Newton iteration.
"""
function Muscade.assemble!{mission}(out::AssemblySweepXB200{OX}, asm, dis, model, state, Δt, dbg) where {mission,OX}
    host_contributions!(out, dis, model, state, mission, dbg)          # Hold, DofLoad, … → mb_set_host_elements (dense Re, Rp, Ke)
    c       = out.c
    newmark = 𝕣[c.a₁, c.a₂, c.a₃, c.b₁, c.b₂, c.b₃, c.Δt]
    X       = state.X
    U0      = length(state.U) ≥ 1 && !isempty(state.U[1]) ? pointer(state.U[1]) : Ptr{𝕣}(C_NULL)
    where   = Ref(ErrInfo(0, 0, 0, 0))
    rc = GC.@preserve X state ccall((:mb_sweepx_assemble, LIB), Int32,
            (Ptr{Cvoid}, Int32, Int32, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, 𝕣, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ref{ErrInfo}),
            out.h, OX, mission == :step ? 0 : 1, X[1], OX ≥ 1 ? pointer(X[2]) : C_NULL, OX ≥ 2 ? pointer(X[3]) : C_NULL, U0,
            state.time, newmark, out.Lλ, out.Lλx.nzval, where)
    rc == 3 && muscadeerror((dbg..., ieletyp=where[].ieletyp, iele=where[].iele, t=state.time),
                            "residual(...) returned NaN in R, FB or derivatives")     # src/Assemble.jl:630
    check(out.h, rc, dbg)
    return
end

# Host-evaluated element types: run Muscade's own addin! into small dense buffers and hand them over.
function host_contributions!(out::AssemblySweepXB200{OX}, dis, model, state, mission, dbg) where {OX}
    for ieletyp ∈ out.hosttyp
        eleobj, d = model.eleobj[ieletyp], dis.dis[ieletyp]
        nele, nx  = length(eleobj), length(d.scale.X)
        Re, Rp, Ke = zeros(𝕣, nx, nele), zeros(𝕣, nx, nele), zeros(𝕣, nx * nx, nele)
        tmp  = Muscade.AssemblySweepX{OX,Vector{𝕣},Matrix{𝕣}}(zeros(𝕣, nx), zeros(𝕣, nx, nx), out.c)   # element-local "out"
        id1  = [reshape(collect(1:nx), nx, 1), reshape(collect(1:nx*nx), nx * nx, 1)]                    # identity asm for one element
        for iele = 1:nele
            zero!(tmp.Lλ); fill!(tmp.Lλx, 0)
            index = d.index[iele]
            Xe = NTuple{OX + 1}(x[index.X] for x ∈ state.X)
            Ue = NTuple{1}(u[index.U] for u ∈ state.U)
            Muscade.addin!{mission}(tmp, id1, 1, d.scale, eleobj[iele], (), Xe, Ue, state.A[index.A], state.time, 0., state.SP, dbg)
            Re[:, iele] .= tmp.Lλ; Ke[:, iele] .= vec(tmp.Lλx)
        end
        check(out.h, ccall((:mb_set_host_elements, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}), out.h, ieletyp, Re, Rp, Ke))
    end
end

# ---- device-resident Newton update (SURVEY §8f-1): state.X stays in HBM between iterations -----------------------------------------------
"""
    upload_state!(out,state) / download_state!(state,out)
    Newmarkβdecrement!(out::AssemblySweepXB200{OX},Δx,firstiter) -> (Σ Δx², Σ Lλ²)

Device versions of `Newmarkβdecrement!{OX}` (src/SweepX.jl:98-132) with `getdof!`/`decrement!` (src/Assemble.jl:206-233) for
`Xdofgr = allXdofs(model,dis)`; same rounding sequence as the reference.  With the state resident, `assemble!` passes `C_NULL` for `X0`.
"""
function upload_state!(out::AssemblySweepXB200{OX}, state, dis) where {OX}
    X = state.X
    check(out.h, ccall((:mb_sweepx_set_dof_scale, LIB), Int32, (Ptr{Cvoid}, Ptr{𝕣}), out.h, dis.scaleX))
    GC.@preserve X check(out.h, ccall((:mb_sweepx_set_state, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}),
        out.h, OX, X[1], OX ≥ 1 ? pointer(X[2]) : C_NULL, OX ≥ 2 ? pointer(X[3]) : C_NULL, C_NULL))
end
function download_state!(state, out::AssemblySweepXB200{OX}) where {OX}
    X = state.X
    GC.@preserve X check(out.h, ccall((:mb_sweepx_get_state, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}),
        out.h, OX, X[1], OX ≥ 1 ? pointer(X[2]) : C_NULL, OX ≥ 2 ? pointer(X[3]) : C_NULL))
end
function Newmarkβdecrement!(out::AssemblySweepXB200{OX}, Δx::Vector{𝕣}, firstiter::Bool) where {OX}
    c = out.c;  newmark = 𝕣[c.a₁, c.a₂, c.a₃, c.b₁, c.b₂, c.b₃, c.Δt]
    Δx², Lλ² = Ref(0.), Ref(0.)
    check(out.h, ccall((:mb_sweepx_newmark_decrement, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{𝕣}, Ptr{𝕣}, Ref{𝕣}, Ref{𝕣}),
        out.h, OX, firstiter, Δx, newmark, Δx², Lλ²))
    return Δx²[], Lλ²[]
end

# ---- getresult for all EulerBeam3D elements of a type (src/Output.jl:131-181): 77×nele matrix, layout in include/muscade_b200.h ---------------
function beam_results(out::AssemblySweepXB200{OX}, ieletyp::Integer, nele::Integer) where {OX}
    res = Matrix{𝕣}(undef, 77, nele)
    check(out.h, ccall((:mb_beam_results, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}),
        out.h, ieletyp, OX, C_NULL, C_NULL, C_NULL, res))          # C_NULL: at the device-resident state
    return res          # res[1,:] ε, res[2:10,:] rₛₘ, res[11:13,:] κ, then per Gauss point x, κgp, fᵢ, mᵢ, fₑ, mₑ
end

# Sliding window over the time steps of a DirectXUA problem too long to materialise (BASELINE configs[3]): `h` was prepared for an interior window
# [lo,hi) (3 ≤ lo, hi ≤ nstep-3, 0-based); returns the shift to add to the rowval the handle was built with.
function direct_rebase!(h::Ptr{Cvoid}, new_lo::Integer)
    shift = Ref{Int64}(0)
    check(h, ccall((:mb_direct_rebase, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{Int64}), h, new_lo, shift))
    return shift[]
end

# L2[X,X][1,1] entries of host-evaluated element types whose residual is not linear in X (second-order branch of DirectXUA, src/DirectXUA.jl:121-150):
# i, j 1-based model X dofs (one triplet per entry, summed over the elements in element order by the caller), v already scaled by scale.X[i]*scale.X[j].
function direct_set_host_xx!(h::Ptr{Cvoid}, step::Integer, i::Vector{Int64}, j::Vector{Int64}, v::Vector{Float64})
    GC.@preserve i j v check(h, ccall((:mb_direct_set_host_xx, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
                                      h, step, length(i), i, j, v))
end


# =====================================================================================================================================================
# DirectXUA on the device: AssemblyDirectB200{OX,OU,0} replaces what `solve(DirectXUA{OX,OU,0};…)` (src/DirectXUA.jl:438-513) does between its
# `prepare` (:452) and its linear solve (:486-492): prepare(AssemblyDirect) + preparebig (:22-56, :245-315), assemblebig!{:matrices} (:316-356),
# sparser! (src/SparseTools.jl:172-199) and decrementbig! (:357-383).  One handle owns the time steps [lo,hi) (0-based) of ONE experiment; with several
# GPUs each process owns a contiguous range and the halo blocks travel over NCCL inside the shim (mb_direct_halo_exchange).
# Mirrors muscade.jl_b200/directxua.py (`prepare`, `DirectEngine`, `solve`) call for call — that Python host is what the GPU tests execute.
# =====================================================================================================================================================
mutable struct AssemblyDirectB200{OX,OU,IA} <: Assembly
    h        :: Ptr{Cvoid}
    nstep    :: Int
    lo       :: Int                               # owned steps lo:hi-1 (0-based, as the C ABI counts them)
    hi       :: Int
    ncol     :: Int                               # owned columns of Lvv = rows of Lv: (hi-lo)·(2nX+nU)
    nnz      :: Int
    Lv       :: Vector{𝕣}
    hosttyp  :: Vector{Int}                       # X-class types evaluated by Muscade's own addin! on the host (Hold, DofLoad, DofConstraint{:X})
    costtyp  :: Vector{Int}                       # SingleDofCost types (user closures) → mb_direct_set_host_cost
    gaugetyp :: Vector{Tuple{Int,Int32}}          # (ieletyp, device type) of ElementCost{StrainGaugeOnEulerBeam3D} types with a quadratic strain cost (devicebeam below)
end

# finitediff(order,n,s) (src/FiniteDifferences.jl:8-31) is Muscade's own; only the block pattern of the OWNED columns is needed here:
# makepattern(0,[nstep],out) (src/DirectXUA.jl:245-307) restricted to the block columns 3lo:3hi-1, as CSC over blocks, 0-based block numbers 3·step+class
function owned_block_pattern(OX, OU, nstep, lo, hi)
    nder = (1, OX + 1, OU + 1)
    cols = Dict{Int,Set{Int}}()
    for istep = max(1, lo - 1):min(nstep, hi + 2), α = 1:3, β = 1:3
        α == 1 && β == 1 && continue                                                   # Lλλ is always zero (src/DirectXUA.jl:41)
        for αder = 1:nder[α], βder = 1:nder[β]
            for (Δαs, _) ∈ Muscade.finitediff(αder - 1, nstep, istep), (Δβs, _) ∈ Muscade.finitediff(βder - 1, nstep, istep)
                sc = istep + Δβs - 1
                lo ≤ sc < hi && push!(get!(cols, 3sc + β - 1, Set{Int}()), 3 * (istep + Δαs - 1) + α - 1)
            end
        end
    end
    bcolptr, browval = Int32[0], Int32[]
    for bc = 3lo:3hi-1
        append!(browval, sort!(collect(get(cols, bc, Set{Int}()))))
        push!(bcolptr, length(browval))
    end
    return bcolptr, browval
end

"""
    out,asm,dofgr = prepare(AssemblyDirectB200{OX,OU,0},model,dis;nstep,Δt,lo=0,hi=nstep,device=0,t₀=0.)

Replaces `prepare(AssemblyDirect{OX,OU,IA},…)` + `preparebig` (src/DirectXUA.jl:452-455).  Class-pair patterns, element→nz maps and the CSC
structure of the owned columns of `Lvv` are built on the device, bit-identical to `asmmat!` / `SparseTools.prepare` (fetch them with
`mb_direct_class_pattern`, `mb_direct_get_asm`, `mb_direct_big_pattern` when the host needs them).
"""
function Muscade.prepare(::Type{AssemblyDirectB200{OX,OU,0}}, model, dis; nstep, Δt, lo=0, hi=nstep, device=0, t₀=0.) where {OX,OU}
    href = Ref{Ptr{Cvoid}}()
    check(C_NULL, ccall((:mb_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, href))
    h = href[]
    hosttyp, costtyp, gaugetyp = Int[], Int[], Tuple{Int,Int32}[]
    for ieletyp = 1:getneletyp(model)
        eleobj, d, E = model.eleobj[ieletyp], dis.dis[ieletyp], eltype(model.eleobj[ieletyp])
        nx   = length(d.scale.X)
        idxX = [d.index[iele].X[i] for i = 1:nx, iele = 1:length(eleobj)]
        udof = length(d.scale.U) > 0
        idxU = udof ? [d.index[iele].U[i] for i = 1:length(d.scale.U), iele = 1:length(eleobj)] : Matrix{Int64}(undef, 0, 0)
        ityp = Ref{Int32}()
        beam = devicebeam(E)
        if !isnothing(beam) && !isnothing(beam.gauge)             # ElementCost accelerator inside the windowed path (src/DirectXUA.jl:172-198 as it is meant)
            beams = beam.unwrap.(eleobj)
            GC.@preserve beams idxX idxU check(h, ccall((:mb_add_eulerbeam3d, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
                h, length(beams), pointer(reinterpret(Float64, beams)), udof, idxX, udof ? pointer(idxU) : C_NULL,
                collect(d.scale.X), udof ? collect(d.scale.U) : C_NULL, ityp))
            g = beam.gauge(eleobj[1])
            G = permutedims(hcat(g.E, g.K1, g.K2, g.K3))          # [Ngauge][4] row-major
            check(h, ccall((:mb_direct_set_gauge_cost, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{𝕣}, 𝕣), h, ityp[], length(g.E), G, g.σ))
            push!(gaugetyp, (ieletyp, ityp[]))
            continue
        end
        if E <: Muscade.Toolbox.EulerBeam3D{Muscade.Toolbox.BeamCrossSection} || E <: Muscade.Toolbox.Bar3D{Muscade.Toolbox.AxisymmetricBarCrossSection}
            f = E <: Muscade.Toolbox.EulerBeam3D ? :mb_add_eulerbeam3d : :mb_add_bar3d
            GC.@preserve eleobj idxX idxU check(h, ccall((f, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
                h, length(eleobj), pointer(reinterpret(Float64, eleobj)), udof, idxX, udof ? pointer(idxU) : C_NULL,
                collect(d.scale.X), udof ? collect(d.scale.U) : C_NULL, ityp))
        elseif E <: Muscade.Toolbox.SoilContact
            GC.@preserve eleobj idxX check(h, ccall((:mb_add_soilcontact, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Ptr{Int64}, Ptr{Float64}, Ref{Int32}),
                h, length(eleobj), pointer(reinterpret(Float64, eleobj)), idxX, collect(d.scale.X), ityp))
        elseif E <: Muscade.SingleDofCost
            push!(costtyp, ieletyp)                               # no device group: gradient / second derivative per dof, mb_direct_set_host_cost
        elseif !udof && length(d.scale.A) == 0
            check(h, ccall((:mb_add_host_elements, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{Int64}, Ref{Int32}), h, length(eleobj), nx, idxX, ityp))
            push!(hosttyp, ieletyp)
        else
            muscadeerror((;ieletyp), "AssemblyDirectB200: element types with U or A dofs other than EulerBeam3D / Bar3D stay on the reference path")
        end
    end
    nX, nU = getndof(model, :X), getndof(model, :U)
    bcolptr, browval = owned_block_pattern(OX, OU, nstep, lo, hi)
    ncol, nnz = Ref{Int64}(), Ref{Int64}()
    check(h, ccall((:mb_direct_prepare, LIB), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int64, Int64, Int64, Int64, Int64, 𝕣, Ptr{Int32}, Ptr{Int32}, Ref{Int64}, Ref{Int64}),
        h, OX, OU, nX, nU, nstep, lo, hi, Δt, bcolptr, browval, ncol, nnz))
    check(h, ccall((:mb_direct_set_time0, LIB), Int32, (Ptr{Cvoid}, 𝕣), h, t₀))
    check(h, ccall((:mb_direct_set_lambda_scale, LIB), Int32, (Ptr{Cvoid}, 𝕣), h, model.scaleΛ))
    check(h, ccall((:mb_direct_set_dof_scale, LIB), Int32, (Ptr{Cvoid}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}), h, dis.scaleΛ, dis.scaleX, dis.scaleU))
    out = AssemblyDirectB200{OX,OU,0}(h, nstep, lo, hi, ncol[], nnz[], zeros(𝕣, ncol[]), hosttyp, costtyp, gaugetyp)
    finalizer(o -> ccall((:mb_destroy, LIB), Int32, (Ptr{Cvoid},), o.h), out)
    dofgr = (Muscade.allΛdofs(model, dis), allXdofs(model, dis), Muscade.allUdofs(model, dis), Muscade.allAdofs(model, dis))
    return out, nothing, dofgr
end

"state[step] (one experiment) → device; the handle stores the steps lo-2:hi+1 its finite-difference stencils reach"
function upload_states!(out::AssemblyDirectB200{OX,OU}, state::Vector) where {OX,OU}
    for s = max(0, out.lo - 2):min(out.nstep, out.hi + 2)-1
        st = state[s+1]; X = st.X
        GC.@preserve X st begin
            check(out.h, ccall((:mb_direct_set_state, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}),
                out.h, s, X[1], OX ≥ 1 ? pointer(X[2]) : C_NULL, OX ≥ 2 ? pointer(X[3]) : C_NULL, isempty(st.U[1]) ? C_NULL : pointer(st.U[1])))
            check(out.h, ccall((:mb_direct_set_lambda, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{𝕣}), out.h, s, st.Λ[1]))
        end
    end
end
"measured strains εₘ(t) of every stored step for the costed beam types (the functor's captured εₘ, evaluated at the step's time): once per solve, and after a rebase for the new steps"
function upload_gauge_measurements!(out::AssemblyDirectB200, model, state::Vector)
    for (ieletyp, ityp) ∈ out.gaugetyp
        εₘ = devicebeam(eltype(model.eleobj[ieletyp])).gauge(model.eleobj[ieletyp][1]).εₘ
        for s = max(0, out.lo - 2):min(out.nstep, out.hi + 2)-1
            e = collect(𝕣, εₘ(state[s+1].time))
            check(out.h, ccall((:mb_direct_set_gauge_measurements, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{𝕣}, Int32), out.h, s, ityp, e, 0))
        end
    end
end

"""
    assemblebig!{:matrices}(out::AssemblyDirectB200,…)  — replaces src/DirectXUA.jl:316-356 for the owned steps

Evaluates every owned step (and, single GPU, the halo steps) on the device, merges the host-evaluated types, forms the owned columns of `Lvv`
(left in HBM: `mb_direct_sparser` / `mb_direct_get_sparse` hand the compacted CSC to the solver) and the owned rows of `Lv` (returned in `out.Lv`).
`comm=true`: time shards over NCCL — own steps only, then `mb_direct_halo_exchange`.
"""
function assemblebig!(out::AssemblyDirectB200{OX,OU}, model, dis, state::Vector, dbg; comm=false) where {OX,OU}
    host_direct_contributions!(out, model, dis, state, dbg)
    where = Ref(ErrInfo(0, 0, 0, 0))
    if comm
        rc = ccall((:mb_direct_assemble, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{𝕣}, Ptr{𝕣}, Ref{ErrInfo}), out.h, out.lo, out.hi, 0, C_NULL, C_NULL, where)
        rc == 0 && (rc = ccall((:mb_direct_halo_exchange, LIB), Int32, (Ptr{Cvoid},), out.h))
        rc == 0 && (rc = ccall((:mb_direct_assemble, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{𝕣}, Ptr{𝕣}, Ref{ErrInfo}), out.h, out.lo, out.lo, 1, C_NULL, out.Lv, where))
    else
        rc = ccall((:mb_direct_assemble, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Int32, Ptr{𝕣}, Ptr{𝕣}, Ref{ErrInfo}), out.h, -1, -1, 1, C_NULL, out.Lv, where)
    end
    rc == 3 && muscadeerror((dbg..., ieletyp=where[].ieletyp, iele=where[].iele, istep=where[].step), "residual(...) returned NaN in R, FB or derivatives")
    check(out.h, rc, dbg)
    return
end

"""
    K,C,M = step_matrices(out,step)

`out.L2[ind.Λ,ind.X][1,1]`, `[1,2]`, `[1,3]` of one evaluated step as `SparseMatrixCSC` — what `solve(EigX{ℝ})` / `solve(EigX{ℂ})` read after their single
`assemble!{:matrices}(out::AssemblyDirect{2,0,0},…)` (src/EigX.jl:33-40, 106-113): with `AssemblyDirectB200{2,0,0}` (nstep = 6, lo = 0, hi = 1, the state uploaded to step 0,
`mb_direct_assemble(h,0,1,0,…)`) EigX runs on the device assembly.  Mirrors muscade.jl_b200/eigx.py, which the GPU tests run against test/TestEigX.jl's goldens.
"""
function step_matrices(out::AssemblyDirectB200{OX}, step=out.lo) where {OX}
    n = Ref{Int64}()
    check(out.h, ccall((:mb_direct_class_pattern, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Int64}, Ptr{Int64}, Ptr{Int64}), out.h, 0, n, C_NULL, C_NULL))
    nX     = getndofX(out)
    colptr = Vector{Int64}(undef, nX + 1); rowval = Vector{Int64}(undef, n[])
    check(out.h, ccall((:mb_direct_class_pattern, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Int64}, Ptr{Int64}, Ptr{Int64}), out.h, 0, n, colptr, rowval))
    mats = map(0:OX) do der
        nz = Vector{𝕣}(undef, n[])
        check(out.h, ccall((:mb_direct_get_step_block, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Int32, Ptr{𝕣}), out.h, step, 1, der, nz))
        SparseMatrixCSC(nX, nX, colptr, rowval, nz)
    end
    return mats
end
"number of X-dofs of the model behind a handle (mb_direct_get_state writes nX values per derivative; the shim reports it through the size of the L1[Λ] block)"
function getndofX(out::AssemblyDirectB200)
    p, n = Ref{Ptr{𝕣}}(), Ref{Int64}()
    check(out.h, ccall((:mb_direct_step_ptrs, LIB), Int32, (Ptr{Cvoid}, Int64, Ref{Ptr{𝕣}}, Ref{Int64}, Ref{Ptr{𝕣}}, Ref{Int64}, Ref{Ptr{𝕣}}, Ref{Int64}),
                     out.h, out.lo, Ref{Ptr{𝕣}}(), Ref{Int64}(), Ref{Ptr{𝕣}}(), Ref{Int64}(), p, n))
    return Int(n[])
end

"cLvv = sparser!(Lvv,rtol) on the device (src/SparseTools.jl:172-199, called at src/DirectXUA.jl:486): the owned columns, global rows, compacted"
function sparser(out::AssemblyDirectB200, rtol=1e-9)
    n = Ref{Int64}()
    check(out.h, ccall((:mb_direct_sparser, LIB), Int32, (Ptr{Cvoid}, 𝕣, Ref{Int64}), out.h, rtol, n))
    colptr, rowval, nzval = Vector{Int64}(undef, out.ncol + 1), Vector{Int64}(undef, n[]), Vector{𝕣}(undef, n[])
    check(out.h, ccall((:mb_direct_get_sparse, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{𝕣}), out.h, colptr, rowval, nzval))
    return colptr, rowval, nzval          # rows are GLOBAL: with one handle, SparseMatrixCSC(ncol,ncol,colptr,rowval,nzval) is the reference's cLvv
end

"decrementbig!(state,Δ²,Lvdis,dofgr,Δv,nder,Δt,nstep) (src/DirectXUA.jl:357-383) on the device-resident states; Δv rows of steps s0:s1-1 in Lv's layout"
function decrementbig!(out::AssemblyDirectB200, Δv::Vector{𝕣}, s0=0, s1=out.nstep)
    Δ² = zeros(𝕣, 3)
    check(out.h, ccall((:mb_direct_decrement, LIB), Int32, (Ptr{Cvoid}, Int64, Int64, Ptr{𝕣}, Ptr{𝕣}), out.h, s0, s1, Δv, Δ²))
    return Δ²                             # maxₜ ΣΔΛ², ΣΔX², ΣΔU² over the owned steps: the convergence test of src/DirectXUA.jl:493-500
end

"states back to the host after convergence (src/DirectXUA.jl:505-511)"
function download_states!(state::Vector, out::AssemblyDirectB200{OX,OU}) where {OX,OU}
    for s = out.lo:out.hi-1
        st = state[s+1]; X = st.X
        GC.@preserve X st check(out.h, ccall((:mb_direct_get_state, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}),
            out.h, s, X[1], OX ≥ 1 ? pointer(X[2]) : C_NULL, OX ≥ 2 ? pointer(X[3]) : C_NULL, isempty(st.U[1]) ? C_NULL : pointer(st.U[1]), st.Λ[1]))
    end
end

# Host-evaluated types of the all-steps problem: Muscade's own second-order addin! (src/DirectXUA.jl:121-171) on a one-element AssemblyDirect, one
# stored step at a time, handed over as dense per-element arrays (mb_direct_set_host_elements, mb_direct_set_host_xx) or per-dof vectors (SingleDofCost,
# mb_direct_set_host_cost).  Mirrors directxua.host_elements / directxua.host_costs of the Python host.
function host_direct_contributions!(out::AssemblyDirectB200{OX,OU}, model, dis, state::Vector, dbg) where {OX,OU}
    (isempty(out.hosttyp) && isempty(out.costtyp)) && return
    nX, nU, nd = getndof(model, :X), getndof(model, :U), OX + 1
    for s = max(0, out.lo - 2):min(out.nstep, out.hi + 2)-1
        st = state[s+1]
        if !isempty(out.costtyp)
            gX, hX, gU, hU = zeros(𝕣, nX), zeros(𝕣, nX), zeros(𝕣, nU), zeros(𝕣, nU)
            for ieletyp ∈ out.costtyp, (iele, e) ∈ enumerate(model.eleobj[ieletyp])
                d = dis.dis[ieletyp]; index = d.index[iele]
                isX = length(index.X) == 1
                x   = Muscade.variate{2,1}(Muscade.variate{1,1}(isX ? st.X[1][index.X[1]] : st.U[1][index.U[1]]))      # second-order dual of the one dof
                c   = e.cost(x, st.time, e.costargs...)
                sc  = isX ? d.scale.X[1] : d.scale.U[1]
                g, hh = Muscade.∂{2,1}(c)[1], Muscade.∂{1,1}(Muscade.∂{2,1}(c)[1])[1]
                if isX gX[index.X[1]] += Muscade.value{1}(g) * sc; hX[index.X[1]] += hh * sc^2
                else   gU[index.U[1]] += Muscade.value{1}(g) * sc; hU[index.U[1]] += hh * sc^2 end
            end
            check(out.h, ccall((:mb_direct_set_host_cost, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}), out.h, s, gX, hX, gU, hU))
        end
        for ieletyp ∈ out.hosttyp
            eleobj, d = model.eleobj[ieletyp], dis.dis[ieletyp]
            nele, nx  = length(eleobj), length(d.scale.X)
            R, dR, GX = zeros(𝕣, nx, nele), zeros(𝕣, nx, nd * nx, nele), zeros(𝕣, nd, nx, nele)
            tmp, tasm, _ = Muscade.prepare(Muscade.AssemblyDirect{OX,OU,0}, one_element_model(model, ieletyp), one_element_dis(dis, ieletyp))   # element-local out
            ii, jj, vv = Int64[], Int64[], 𝕣[]
            for iele = 1:nele
                zero!(tmp)
                index = d.index[iele]
                Muscade.addin!{:matrices}(tmp, tasm, 1, d.scale, eleobj[iele], (st.Λ[1][index.X],), NTuple{nd}(x[index.X] for x ∈ st.X), (st.U[1][index.U],),
                                          st.A[index.A], st.time, 0., st.SP, dbg)
                R[:, iele] .= tmp.L1[1][1]
                for der = 1:nd
                    GX[der, :, iele] .= tmp.L1[2][der]
                    dR[:, (der-1)*nx+1:der*nx, iele] .= Matrix(tmp.L2[1, 2][1, der])
                end
                H = Matrix(tmp.L2[2, 2][1, 1])
                for a = 1:nx, b = 1:nx
                    H[a, b] != 0 && (push!(ii, index.X[a]); push!(jj, index.X[b]); push!(vv, H[a, b]))
                end
            end
            check(out.h, ccall((:mb_direct_set_host_elements, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}), out.h, s, ieletyp, R, dR, GX))
            isempty(ii) || direct_set_host_xx!(out.h, s, ii, jj, vv)
        end
    end
end
# one-element views of model / dis for the element-local AssemblyDirect above (a maintainer has these as two-line helpers over model.eleobj / dis.dis)
one_element_model(model, ieletyp) = Muscade.submodel(model, ieletyp, 1)
one_element_dis(dis, ieletyp)     = Muscade.subdis(dis, ieletyp, 1)

"""
    solve_direct_b200(OX,OU;initialstate,time,maxiter=50,maxΔλ=1e-5,maxΔx=1e-5,maxΔu=1e-5,device=0)

The body of `solve(DirectXUA{OX,OU,0};…)` (src/DirectXUA.jl:438-513) with the assembly, `sparser!` and `decrementbig!` on the device.
"""
function solve_direct_b200(OX, OU; initialstate, time, maxiter=50, maxΔλ=1e-5, maxΔx=1e-5, maxΔu=1e-5, device=0, dbg=(;))
    model, dis = initialstate.model, initialstate.dis
    nstep, Δt  = length(time), step(time)
    out, _, _  = Muscade.prepare(AssemblyDirectB200{OX,OU,0}, model, dis; nstep, Δt, device, t₀=first(time))
    state      = [Muscade.State{1,OX+1,OU+1}(copy(initialstate, time=t)) for t ∈ time]
    upload_states!(out, state)
    upload_gauge_measurements!(out, model, state)
    for iter = 1:maxiter
        assemblebig!(out, model, dis, state, (dbg..., iter=iter))
        colptr, rowval, nzval = sparser(out, 1e-20)
        Δv = try lu(SparseMatrixCSC(out.ncol, out.ncol, colptr, rowval, nzval)) \ out.Lv
             catch; muscadeerror(@sprintf("Lvv matrix factorization failed at iter=%i", iter)) end
        Δ² = decrementbig!(out, Δv)
        all(Δ² .≤ (maxΔλ^2, maxΔx^2, maxΔu^2)) && break
        iter == maxiter && muscadeerror(@sprintf("no convergence after %3d iterations.", iter))
        isempty(out.hosttyp) && isempty(out.costtyp) || download_states!(state, out)       # the host-evaluated types read the updated states
    end
    download_states!(state, out)
    return state
end

# ---- NCCL inside the shim: one handle per GPU / process ---------------------------------------------------------------------------------------------
comm_unique_id() = (id = zeros(UInt8, 128); ccall((:mb_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id) == 0 || error("ncclGetUniqueId failed"); id)
comm_init!(h::Ptr{Cvoid}, id::Vector{UInt8}, rank, world) = check(h, ccall((:mb_comm_init, LIB), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), h, id, rank, world))
comm_share!(h::Ptr{Cvoid}, owner::Ptr{Cvoid})             = check(h, ccall((:mb_comm_share, LIB), Int32, (Ptr{Cvoid}, Ptr{Cvoid}), h, owner))
comm_allreduce!(h::Ptr{Cvoid}, v::Vector{𝕣}, op=0)        = check(h, ccall((:mb_comm_allreduce, LIB), Int32, (Ptr{Cvoid}, Ptr{𝕣}, Int64, Int32), h, v, length(v), op))
"element-range shard of a large SweepX mesh: after `assemble!` with device-resident outputs, the interface rows go to / come from the neighbours"
iface_exchange!(out::AssemblySweepXB200) = check(out.h, ccall((:mb_iface_exchange, LIB), Int32, (Ptr{Cvoid},), out.h))

# ------------------------------------------------------------------------------------------------------------------ DirectXUA{OX,OU,IA}, general form
"""
    AssemblyXUAB200{OX,OU,IA}

`AssemblyDirect{OX,OU,IA}` (src/DirectXUA.jl:18-56) with its maps, patterns, `Lvv`/`Lv`, `Lvvasm`/`Lvasm` on the device (`mb_xua_*`): any element type, A-dofs,
several experiments.  Muscade differentiates the elements with its own dual numbers exactly as `addin!` does (src/DirectXUA.jl:70-171) and hands the partials over
as packets; the hot loop of `assemblebig!` — scatter into `out`, finite-difference weighted addition into `Lvv`/`Lv` — `sparser!` and `decrementbig!` run on the GPU.
"""
mutable struct AssemblyXUAB200{OX,OU,IA} <: Assembly
    h      :: Ptr{Cvoid}
    nstep  :: Vector{Int64}
    Δt     :: Vector{𝕣}
    nbig   :: Int64
    nnzbig :: Int64
    devtyp :: Dict{Int,Any}
end
"""
The strain cost the device evaluates, declared in the input script with Muscade's own macro so that it is a `Functor` (src/Functors.jl:23-30), e.g.

    @functor with(σ=15e-6, εₘ=measured) QuadraticGaugeCost(eleres,t) = (Δε = eleres.ε - εₘ(t); (Δε ⋅ Δε)/(2σ^2))        # test/TestBeamElementStrainGauge.jl:90-97

`devicebeam` recognises an `ElementCost` whose cost is a `Functor{:QuadraticGaugeCost}` and reads σ and εₘ from what it captured.
"""
const GaugeCostFunctor = Muscade.Functor{:QuadraticGaugeCost}
"which element types the device kernels evaluate inside the general form, and how to reach the wrapped beam"
devicebeam(::Type) = nothing
devicebeam(::Type{<:Muscade.Toolbox.EulerBeam3D{Muscade.Toolbox.BeamCrossSection}}) = (unwrap = identity, gauge = nothing)
devicebeam(::Type{<:Muscade.ElementCost{<:Muscade.Toolbox.StrainGaugeOnEulerBeam3D{N,<:Muscade.Toolbox.EulerBeam3D{Muscade.Toolbox.BeamCrossSection}},R,<:GaugeCostFunctor}}) where {N,R} =
    (unwrap = o -> o.eleobj.eleobj, gauge = o -> (E = o.eleobj.E, K1 = o.eleobj.K1, K2 = o.eleobj.K2, K3 = o.eleobj.K3, σ = o.cost.captured.σ, εₘ = o.cost.captured.εₘ))
function Muscade.prepare(::Type{AssemblyXUAB200{OX,OU,IA}}, model, dis; nstep::Vector{Int64}, Δt::Vector{𝕣}, device=0,
                         Xwhite=false, XUindep=false, UAindep=false, XAindep=false) where {OX,OU,IA}
    href = Ref{Ptr{Cvoid}}()
    check(C_NULL, ccall((:mb_create, LIB), Int32, (Int32, Ref{Ptr{Cvoid}}), device, href))
    h = href[]
    devtyp = Dict{Int,Any}()                       # element types the DEVICE evaluates: ieletyp → nothing | (cost = the QuadraticGaugeCost functor's data)
    for ieletyp = 1:getneletyp(model)
        eleobj, d = model.eleobj[ieletyp], dis.dis[ieletyp]
        idx(f, n) = Int64[getfield(d.index[iele], f)[i] for i = 1:n, iele = 1:length(eleobj)]
        nx, nu, na = length(d.scale.X), length(d.scale.U), length(d.scale.A)
        iX, iU, iA = idx(:X, nx), idx(:U, nu), idx(:A, na)
        E    = eltype(eleobj)
        beam = devicebeam(E)                       # EulerBeam3D{BeamCrossSection}, or ElementCost{StrainGaugeOnEulerBeam3D{…,EulerBeam3D}} with a quadratic strain cost
        if !isnothing(beam)
            beams = beam.unwrap.(eleobj)           # the wrapped Vector{EulerBeam3D}: same memory image as mb_add_eulerbeam3d reads
            idev, ityp = Ref{Int32}(), Ref{Int32}()
            GC.@preserve beams iX iU check(h, ccall((:mb_add_eulerbeam3d, LIB), Int32,
                (Ptr{Cvoid}, Int64, Ptr{Float64}, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
                h, length(beams), pointer(reinterpret(Float64, beams)), nu > 0, iX, nu > 0 ? pointer(iU) : C_NULL, collect(d.scale.X), nu > 0 ? collect(d.scale.U) : C_NULL, idev))
            check(h, ccall((:mb_xua_add_device_eletyp, LIB), Int32, (Ptr{Cvoid}, Int32, Ref{Int32}), h, idev[], ityp))
            if !isnothing(beam.gauge)              # G = [E K1 K2 K3] of the gauges (toolbox/StrainGaugeOnBeamElement.jl:62-65), σ of the cost
                g = beam.gauge(eleobj[1])
                G = permutedims(hcat(g.E, g.K1, g.K2, g.K3))            # 4 × Ngauge column-major = [Ngauge][4] row-major
                check(h, ccall((:mb_xua_set_gauge_cost, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Ptr{𝕣}, 𝕣, 𝕣), h, ityp[], length(g.E), G, g.σ, model.scaleΛ))
            end
            devtyp[ieletyp] = beam
            continue
        end
        check(h, ccall((:mb_xua_add_eletyp, LIB), Int32, (Ptr{Cvoid}, Int64, Int32, Int32, Int32, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Int32, Ref{Int32}),
                       h, length(eleobj), nx, nu, na, iX, iU, iA, E <: Muscade.Acost, Ref{Int32}()))
    end
    flags = Int32(Xwhite) | Int32(XUindep) << 1 | Int32(UAindep) << 2 | Int32(XAindep) << 3
    nbig, nnz = Ref{Int64}(), Ref{Int64}()
    check(h, ccall((:mb_xua_prepare, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32, Int64, Int64, Int64, Int32, Ptr{Int64}, Ptr{𝕣}, Int32, Ref{Int64}, Ref{Int64}),
                   h, OX, OU, IA, getndof(model, :X), getndof(model, :U), getndof(model, :A), length(nstep), nstep, Δt, flags, nbig, nnz))
    check(h, ccall((:mb_xua_set_dof_scale, LIB), Int32, (Ptr{Cvoid}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}), h, dis.scaleΛ, dis.scaleX, dis.scaleU, dis.scaleA))
    out = AssemblyXUAB200{OX,OU,IA}(h, nstep, Δt, nbig[], nnz[], devtyp)
    finalizer(o -> ccall((:mb_destroy, LIB), Int32, (Ptr{Cvoid},), o.h), out)
    return out
end

"""
    g,H = packet(eleobj_vector, d, state, OX, OU, IA, SP, dbg; assembleA=false)

∇L [Np × nele] and ∇²L [Np × Np × nele] (Julia column-major = the C ABI's row-major [nele][Np][Np], ∇²L being symmetric) of one element type at one state: the same
calls `addin!` makes (src/DirectXUA.jl:70-84 for `assembleA`, :85-120 for `no_second_order` types — Λ row/column only —, :152-171 otherwise).
"""
function packet(eleobj::Vector{E}, d, state, OX, OU, IA, SP, dbg; assembleA=false) where {E}
    nx, nu, na = length(d.scale.X), length(d.scale.U), length(d.scale.A)
    Np = assembleA ? na : nx + nx * (OX + 1) + nu * (OU + 1) + na * IA
    g, H = zeros(𝕣, Np, length(eleobj)), zeros(𝕣, Np, Np, length(eleobj))
    for (iele, o) ∈ enumerate(eleobj)
        index = d.index[iele]
        Λe = state.Λ[1][index.X]; Xe = ntuple(i -> state.X[i][index.X], OX + 1); Ue = ntuple(i -> state.U[i][index.U], OU + 1); Ae = state.A[index.A]
        if assembleA
            C, _ = Muscade.lagrangian(o, nothing, nothing, nothing, Muscade.revariate{2}(Ae), nothing, nothing, dbg)
            ∇C   = Muscade.∂{2,na}(C)
            g[:, iele] .= Muscade.value{1}.(∇C); H[:, :, iele] .= Muscade.∂{1,na}(∇C)
        elseif Muscade.no_second_order(E) == Val(true)
            v = IA == 1 ? Muscade.revariate{1}((;X=Xe, U=Ue, A=Ae), (;X=d.scale.X, U=d.scale.U, A=d.scale.A)) : (Muscade.revariate{1}((;X=Xe, U=Ue), (;X=d.scale.X, U=d.scale.U))..., Ae)
            R, _ = Muscade.residual(o, v[1], v[2], v[3], state.time, SP, dbg)
            g[1:nx, iele] .= Muscade.value{1}.(R)
            dR = Muscade.∂{1,Np - nx}(R)
            H[1:nx, nx+1:Np, iele] .= dR; H[nx+1:Np, 1:nx, iele] .= dR'
        else
            v = IA == 1 ? Muscade.revariate{2}((;Λ=Λe, X=Xe, U=Ue, A=Ae), (;Λ=d.scale.Λ, X=d.scale.X, U=d.scale.U, A=d.scale.A)) :
                          (Muscade.revariate{2}((;Λ=Λe, X=Xe, U=Ue), (;Λ=d.scale.Λ, X=d.scale.X, U=d.scale.U))..., Ae)
            L, _ = Muscade.getlagrangian(o, v[1], v[2], v[3], v[4], state.time, SP, dbg)
            ∇L   = Muscade.∂{2,Np}(L)
            g[:, iele] .= Muscade.value{1}.(∇L); H[:, :, iele] .= Muscade.∂{1,Np}(∇L)
        end
    end
    return g, H
end

"assemblebig!{:matrices} (src/DirectXUA.jl:316-356): state[iexp][istep]"
function assemblebig!(out::AssemblyXUAB200{OX,OU,IA}, model, dis, state, SP, dbg) where {OX,OU,IA}
    h = out.h
    setp(ityp, g, H) = check(h, ccall((:mb_xua_set_packet, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{𝕣}, Ptr{𝕣}), h, ityp, g, H))
    check(h, ccall((:mb_xua_zero, LIB), Int32, (Ptr{Cvoid},), h))
    if IA == 1
        for ityp = 1:getneletyp(model)
            eltype(model.eleobj[ityp]) <: Muscade.Acost && setp(ityp, packet(model.eleobj[ityp], dis.dis[ityp], state[1][1], OX, OU, IA, SP, dbg; assembleA=true)...)
        end
        check(h, ccall((:mb_xua_add_A, LIB), Int32, (Ptr{Cvoid},), h))
    end
    for iexp = 1:length(out.nstep), istep = 1:out.nstep[iexp]
        if !isempty(out.devtyp)                 # device element types: measurements of the step, then R, ∂R/∂X, ∂R/∂U (and the strain-gauge terms) → packets, all on the device
            for (ityp, beam) ∈ out.devtyp
                isnothing(beam.gauge) && continue
                εₘ = collect(𝕣, beam.gauge(model.eleobj[ityp][1]).εₘ(state[iexp][istep].time))
                check(h, ccall((:mb_xua_set_gauge_measurements, LIB), Int32, (Ptr{Cvoid}, Int32, Ptr{𝕣}, Int32), h, ityp, εₘ, 0))
            end
            where = Ref(ErrInfo(0, 0, 0, 0))
            rc = ccall((:mb_xua_eval_device, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ref{ErrInfo}), h, iexp, istep, where)   # reads the device-resident state (upload_states!)
            rc == 3 && muscadeerror((dbg..., ieletyp=where[].ieletyp, iele=where[].iele, step=istep), "residual(...) returned NaN in R, FB or derivatives")
            check(h, rc)
        end
        for ityp = 1:getneletyp(model)          # Acost vectors too: `assemble_!(…,eleobj::Acost,…) = nothing` (src/Assemble.jl:477) does not match them
            haskey(out.devtyp, ityp) && continue
            setp(ityp, packet(model.eleobj[ityp], dis.dis[ityp], state[iexp][istep], OX, OU, IA, SP, (dbg..., step=istep))...)
        end
        check(h, ccall((:mb_xua_add_step, LIB), Int32, (Ptr{Cvoid}, Int32, Int64), h, iexp, istep))
    end
end

"sparser!(cLvv,Lvv,rtol) on the device → SparseMatrixCSC for the factorisation, and Lv"
function sparser(out::AssemblyXUAB200, rtol=1e-20)
    n = Ref{Int64}()
    check(out.h, ccall((:mb_xua_sparser, LIB), Int32, (Ptr{Cvoid}, 𝕣, Ref{Int64}), out.h, rtol, n))
    colptr, rowval, nzval = Vector{Int64}(undef, out.nbig + 1), Vector{Int64}(undef, n[]), Vector{𝕣}(undef, n[])
    check(out.h, ccall((:mb_xua_get_sparse, LIB), Int32, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{𝕣}), out.h, colptr, rowval, nzval))
    Lv = Vector{𝕣}(undef, out.nbig)
    check(out.h, ccall((:mb_xua_get_big, LIB), Int32, (Ptr{Cvoid}, Ptr{𝕣}, Ptr{𝕣}), out.h, C_NULL, Lv))
    return SparseMatrixCSC(out.nbig, out.nbig, colptr, rowval, nzval), Lv
end

"decrementbig! (src/DirectXUA.jl:357-383) on the device-resident states → Δ² (Λ,X,U,A); upload_states! / download_states! move state[iexp][istep]"
function decrementbig!(out::AssemblyXUAB200, Δv::Vector{𝕣})
    Δ² = zeros(𝕣, 4)
    check(out.h, ccall((:mb_xua_decrement, LIB), Int32, (Ptr{Cvoid}, Ptr{𝕣}, Ptr{𝕣}), out.h, Δv, Δ²))
    return Δ²
end
function upload_states!(out::AssemblyXUAB200{OX,OU,IA}, state) where {OX,OU,IA}
    for iexp = 1:length(out.nstep), istep = 1:out.nstep[iexp]
        s = state[iexp][istep]
        check(out.h, ccall((:mb_xua_set_state, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}),
                           out.h, iexp, istep, s.Λ[1], reduce(vcat, s.X), reduce(vcat, s.U), s.A))
    end
end
function download_states!(state, out::AssemblyXUAB200{OX,OU,IA}) where {OX,OU,IA}
    for iexp = 1:length(out.nstep), istep = 1:out.nstep[iexp]
        s = state[iexp][istep]
        X, U = Vector{𝕣}(undef, (OX + 1) * length(s.X[1])), Vector{𝕣}(undef, (OU + 1) * length(s.U[1]))
        check(out.h, ccall((:mb_xua_get_state, LIB), Int32, (Ptr{Cvoid}, Int32, Int64, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}, Ptr{𝕣}), out.h, iexp, istep, s.Λ[1], X, U, s.A))
        for i = 1:OX+1 s.X[i] .= view(X, (i-1)*length(s.X[1])+1:i*length(s.X[1])) end
        for i = 1:OU+1 s.U[i] .= view(U, (i-1)*length(s.U[1])+1:i*length(s.U[1])) end
    end
end

end # module
