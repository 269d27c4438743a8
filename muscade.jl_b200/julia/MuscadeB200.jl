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

using Muscade, SparseArrays
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

end # module
