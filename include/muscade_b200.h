/* muscade_b200.h — C ABI of the B200 element-evaluation-and-assembly engine.
 *
 * Drop-in boundary for ONE path of SINTEF/Muscade.jl (v0.7.0): what `assemble!{mission}(out,asm,dis,model,state,Δt,dbg)`
 * (src/Assemble.jl:470-487) does for the solvers SweepX{0,1,2} (src/SweepX.jl:24-96) and DirectXUA (src/DirectXUA.jl:22-120,
 * 316-356): gather element dofs, evaluate every element's residual with first-order forward-mode partials, scatter-add
 * values into the gradient vector and partials into the sparse matrix.  The reference has no FFI; a Julia host binds these
 * entry points with `ccall` from new `Assembly` subtypes.  All arrays crossing the boundary are the
 * reference's own: Float64, Int64, 1-based indices, column-major.
 *
 * Conventions: every function returns int32 status (MB_OK = 0). No exception crosses the ABI. Host pointers unless the name
 * says `_dev`. A handle is single-caller (the reference calls assemble! from one task, State/out are not thread safe).
 * There is no CPU fallback: without a CUDA device mb_create fails with MB_ERR_CUDA.
 */
#ifndef MUSCADE_B200_H
#define MUSCADE_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mb_handle mb_handle;

enum { MB_OK = 0, MB_ERR_CUDA = 1, MB_ERR_ARG = 2, MB_ERR_NAN = 3, MB_ERR_STATE = 4, MB_ERR_TOOBIG = 5, MB_ERR_NCCL = 6 };

/* Where the first NaN was found, in element order (deterministic).  Replaces the MuscadeException thrown at
 * src/Assemble.jl:625,630 ("residual(...) returned NaN in R, FB or derivatives"); the wrapper rethrows with (ieletyp,iele). */
typedef struct { int32_t kind; int32_t ieletyp; int64_t iele; int64_t step; } mb_errinfo;   /* 1-based ieletyp / iele / step, 0 = none */

/* ---- lifetime ------------------------------------------------------------------------------------------------------ */
int32_t     mb_create(int32_t device, mb_handle** out);
int32_t     mb_destroy(mb_handle* h);
const char* mb_last_error(const mb_handle* h);
int32_t     mb_version(void);

/* ---- element groups: one per reference element type, in model.eleobj order (src/ModelDescription.jl:212-219) ------------ */
/* EulerBeam3D{BeamCrossSection,Udof} (toolbox/BeamElement.jl:87-103).
 *   eleobj : nele × 69 Float64 — the reference's isbits Vector{EulerBeam3D{BeamCrossSection,Udof}} as it lies in memory
 *            (cₘ3 rₘ9(col-major) ζgp4 ζnod2 tgₘ3 tgₑ3 yₐ4 yᵤ4 yᵥ4 κₐ4 κᵤ4 κᵥ4 L dL4 mat16).
 *   idxX   : 12 × nele Int64, dis.dis[ityp].index[iele].X (src/Assemble.jl:15,96); idxU: 3 × nele or NULL when !udof.
 *   scaleX : 12, scaleU : 3 — dis.dis[ityp].scale.X/.U (src/Assemble.jl:68). */
int32_t mb_add_eulerbeam3d(mb_handle* h, int64_t nele, const double* eleobj, int32_t udof, const int64_t* idxX, const int64_t* idxU,
                           const double* scaleX, const double* scaleU, int32_t* ieletyp_out);
/* Bar3D{AxisymmetricBarCrossSection,Udof} (toolbox/BarElement.jl:89-101): nele × 38 Float64
 *   (cₘ3 tgₘ3 tgₑ3 L₀ Lₛ mat9 wgp4 ζgp4 ζnod2 ψ₁4 ψ₂4); idxX 6 × nele. */
int32_t mb_add_bar3d(mb_handle* h, int64_t nele, const double* eleobj, int32_t udof, const int64_t* idxX, const int64_t* idxU,
                     const double* scaleX, const double* scaleU, int32_t* ieletyp_out);
/* SoilContact (toolbox/SoilContact.jl:2-8): nele × 5 Float64 (z₀ Kh Kv Ch Cv); idxX 3 × nele. */
int32_t mb_add_soilcontact(mb_handle* h, int64_t nele, const double* eleobj, const int64_t* idxX, const double* scaleX,
                           int32_t* ieletyp_out);
/* Any other element type with X-dofs only (Hold, DofLoad, DofConstraint{:X}, …; src/BasicElements.jl): the host evaluates it
 * (user closures cannot cross a C ABI) and hands dense, already scaled element contributions before each assemble:
 *   Re : nx × nele     (Lλ .* scale.X, SweepX.jl:55)      Rp : nx × nele, the :step predictor partial, or NULL
 *   Ke : nx·nx × nele  (column-major i + nx·(j−1), Lλ[i].dx[j]) */
int32_t mb_add_host_elements(mb_handle* h, int64_t nele, int32_t nx, const int64_t* idxX, int32_t* ieletyp_out);
int32_t mb_set_host_elements(mb_handle* h, int32_t ieletyp, const double* Re, const double* Rp, const double* Ke);

/* ---- SweepX: prepare(AssemblySweepX{OX},model,dis) (src/SweepX.jl:24-33) ----------------------------------------------- */
/* Builds on the device, bit-identical to asmvec! / asmmat! (src/Assemble.jl:340-357, 373-448):
 *   asm[1,ityp], asm[2,ityp], and the CSC pattern (colptr, rowval) of Lλx; plus the nnz-sorted contributor lists used by the
 *   atomics-free segmented reduction. ndofX = getndof(model,:X). */
int32_t mb_sweepx_prepare(mb_handle* h, int64_t ndofX, int64_t* nnz_out);
int32_t mb_sweepx_get_pattern(mb_handle* h, int64_t* colptr /* ndofX+1 */, int64_t* rowval /* nnz */);
int32_t mb_sweepx_get_asm(mb_handle* h, int32_t ieletyp, int64_t* asm1 /* nx × nele */, int64_t* asm2 /* nx² × nele */);
/* same for the elements [e0,e1) (0-based) only — the interface elements of a shard */
int32_t mb_sweepx_get_asm_range(mb_handle* h, int32_t ieletyp, int64_t e0, int64_t e1, int64_t* asm1, int64_t* asm2);

/* assemble!{:step|:iter}(out::AssemblySweepX{OX},…) (src/Assemble.jl:470, src/SweepX.jl:45-96).
 *   mission : 0 = :step, 1 = :iter.   X0,X1,X2 : state.X[1..OX+1] (ndofX each; X1/X2 may be NULL when OX is lower).
 *   U0 : state.U[1] (only read by Udof elements; may be NULL).   newmark : a₁ a₂ a₃ b₁ b₂ b₃ Δt (src/SweepX.jl:3-15).
 *   Llambda (ndofX) and nzval (nnz) receive out.Lλ and out.Lλx.nzval; either may be NULL to leave the result on the device. */
int32_t mb_sweepx_assemble(mb_handle* h, int32_t OX, int32_t mission, const double* X0, const double* X1, const double* X2,
                           const double* U0, double t, const double* newmark, double* Llambda, double* nzval, mb_errinfo* where);
/* Same with the state already resident on the device (see mb_state_ptrs_dev) and results left there. Asynchronous on the
 * handle's stream; mb_sync waits and reports NaN. */
int32_t mb_sweepx_assemble_dev(mb_handle* h, int32_t OX, int32_t mission, double t, const double* newmark);
int32_t mb_sync(mb_handle* h, mb_errinfo* where);

/* Device buffers owned by the handle, so that a device sparse solver (cuDSS) and the state update can consume them without a
 * host round-trip: X0,X1,X2 (ndofX), U0, Llambda (ndofX), nzval (nnz), colptr32/rowval32 (0-based int32 CSC). */
typedef struct { double *X0, *X1, *X2, *U0, *Llambda, *nzval; int32_t *colptr0, *rowval0; int64_t ndofX, ndofU, nnz; } mb_dev_ptrs;
int32_t mb_get_device_ptrs(mb_handle* h, mb_dev_ptrs* out);
int32_t mb_set_ndofU(mb_handle* h, int64_t ndofU);

/* ---- DirectXUA{OX,OU,0}: prepare(AssemblyDirect) + preparebig + assemblebig! (src/DirectXUA.jl:22-56, 245-356) -------------------- */
/* The handle owns the block columns of the time steps [step_lo,step_hi) (0-based) of nstep (one experiment) and stores the per-step
 * blocks of [step_lo-2,step_hi+2)∩[0,nstep), which its finite-difference stencils reach (src/FiniteDifferences.jl).  Built on the device,
 * bit-identical to the reference: the class-pair patterns/maps of asmmat! for Λ,X,U and the CSC structure of the owned columns of Lvv
 * (SparseTools.prepare).  bcolptr/browval: CSC over BLOCKS of the owned block columns (3 per step: Λ,X,U), block row = 3·step+class, as
 * makepattern (src/DirectXUA.jl:245-307) yields — computed by the host wrapper, a few entries per step.
 * Device element types: EulerBeam3D, Bar3D (with or without Udof), SoilContact; host-evaluated X-class types and U/X costs merged (below).  This beam-specialised path is
 * IA = 0, one experiment per handle; IA = 1, A-dofs and experiments coupled through A go through the general form (mb_xua_*, further down). */
int32_t mb_direct_prepare(mb_handle* h, int32_t OX, int32_t OU, int64_t ndofX, int64_t ndofU, int64_t nstep, int64_t step_lo, int64_t step_hi,
                          double dt, const int32_t* bcolptr, const int32_t* browval, int64_t* ncol_out, int64_t* nnz_out);
/* class-pair patterns: which = 0 X×X (also Λ×X, X×Λ), 1 X×U (Λ×U), 2 U×X (U×Λ), 3 U×U ; 1-based CSC like SparseMatrixCSC */
int32_t mb_direct_class_pattern(mb_handle* h, int32_t which, int64_t* nnz, int64_t* colptr, int64_t* rowval);
/* asm[arrnum(α,β), ieletyp] for the pair type `which` (element entry → nz, 1-based; n_i·n_j × nele, column-major) */
int32_t mb_direct_get_asm(mb_handle* h, int32_t ieletyp, int32_t which, int64_t* out);
/* state[iexp][step] → device (X0..X_OX of ndofX, U0 of ndofU); step must be stored on this handle */
int32_t mb_direct_set_state(mb_handle* h, int64_t step, const double* X0, const double* X1, const double* X2, const double* U0);
/* state[step].time = t0 + step·dt (default t0 = 0); only Bar3D's weight ramp reads the time (toolbox/BarElement.jl:144) */
int32_t mb_direct_set_time0(mb_handle* h, double t0);
/* model.scaleΛ (setscale!(model;Λscale), src/ModelDescription.jl:287-299): scale.Λ = scale.X·Λscale (src/Assemble.jl:55). Read by the element types
 * that take DirectXUA's second-order path — SoilContact, which does not declare no_second_order (src/DirectXUA.jl:152-171): L1[Λ] = R·scale.Λ,
 * L1[X][der] = Λᵀ∂R/∂X_der, L2[Λ,X] scaled by scale.Λ·scale.X. state[step].Λ comes from mb_direct_set_lambda. Default 1. */
int32_t mb_direct_set_lambda_scale(mb_handle* h, double lambda_scale);
/* assemblebig!{:matrices}: evaluates the steps [eval_lo,eval_hi) (eval_lo<0: all stored steps, i.e. owned + halo recomputed locally),
 * then, if build_big, forms the owned columns of Lvv (nzval) and rows of Lv; host outputs may be NULL (results stay on the device). */
int32_t mb_direct_assemble(mb_handle* h, int64_t eval_lo, int64_t eval_hi, int32_t build_big, double* Lvv_nzval, double* Lv, mb_errinfo* where);
int32_t mb_direct_big_pattern(mb_handle* h, int64_t* colptr /* ncol+1 */, int64_t* rowval /* nnz, global rows */);
/* out.L1/out.L2 of one stored step: which = 0 L1[Λ][1], 1 L2[Λ,X][1,der+1], 2 L2[X,Λ][der+1,1], 3 L2[Λ,U][1,1], 4 L2[U,Λ][1,1] */
int32_t mb_direct_get_step_block(mb_handle* h, int64_t step, int32_t which, int32_t der, double* out);
/* Newton update of the all-steps problem on the device.
 *   mb_direct_sparser     : sparser!(cLvv,Lvv,rtol) (src/SparseTools.jl:172-199, called at src/DirectXUA.jl:486): compacts the owned columns of
 *                           Lvv to the entries with |v| ≥ rtol·max|Lvv| (the structural zeros of the X-X, U-U blocks go), order preserved.
 *   mb_direct_get_sparse  : the compacted CSC (1-based colptr / global rowval) — 40 % fewer bytes over PCIe than the full nzval.
 *   mb_direct_set_lambda / get_state / set_dof_scale : state[step].Λ, reading states back, dofgr scales (Λ, X, U; default ones).
 *   mb_direct_decrement   : decrementbig!(state,Δ²,Lvdis,dofgr,Δv,nder,Δt,nstep) (src/DirectXUA.jl:357-383) for the steps stored on this handle.
 *                           dv = Δv rows of steps [s0,s1) in Lv's layout (per step Λ(nX) X(nX) U(nU)), host or device; the range must reach
 *                           two steps beyond the stored ones (stencils of finitediff). delta2[3] = maxₜ ΣΔβ² over the OWNED steps (Λ,X,U). */
/* Host-evaluated single-dof costs of one stored step (SingleDofCost on X or U dofs, src/BasicElements.jl:198-208, evaluated by the host
 * because `cost` is a user closure): dense per-dof gradient (→ L1[X][1], L1[U][1]) and second derivative (→ diagonal of L2[X,X][1,1],
 * L2[U,U][1,1]) vectors, already multiplied by scale / scale²; NULL = zeros. Merged into Lv / Lvv by mb_direct_assemble. */
int32_t mb_direct_set_host_cost(mb_handle* h, int64_t step, const double* gX, const double* hX, const double* gU, const double* hU);
/* Host-evaluated X-class element types (Hold, DofLoad, DofConstraint — user closures, added with mb_add_host_elements) in DirectXUA's second-order
 * branch (src/DirectXUA.jl:152-171), for residuals linear in X (so L2[X,X] = 0), one stored step at a time:
 *   R  [nele][nx]          Rᵢ·scale.Λᵢ → L1[Λ];     GX [nele][nx][nd]   Σₖ Λₖ·∂Rₖ/∂X_der,ᵢ·scale.Xᵢ → L1[X][der+1];
 *   dR [nele][nd·nx][nx]   ∂Rᵢ/∂X_der,ⱼ·scale.Λᵢ·scale.Xⱼ at [(nx·der+j)][i] → L2[Λ,X][1,der+1] and, transposed, L2[X,Λ][der+1,1].  NULL = zeros. */
int32_t mb_direct_set_host_elements(mb_handle* h, int64_t step, int32_t ieletyp, const double* R, const double* dR, const double* GX);
/* The X-X part of the same branch for host-evaluated types that are NOT linear in X (DofConstraint in `positive` mode, curved gaps; src/DirectXUA.jl:121-150):
 * L2[X,X][1,1](i,j) += Σₖ Λₖ·∂²Rₖ/∂Xᵢ∂Xⱼ·scale.Xᵢ·scale.Xⱼ as triplets (1-based model X dofs), one per (i,j) — the caller sums its elements in element order.
 * Replaces what was set for this step; n = 0 clears. Merged into the (step,step) X-X block of Lvv by mb_direct_assemble. */
int32_t mb_direct_set_host_xx(mb_handle* h, int64_t step, int64_t n, const int64_t* i, const int64_t* j, const double* v);
/* ElementCost{StrainGaugeOnEulerBeam3D} with cost(eleres,t) = Σ_g (ε_g − εm_g(t))²/(2σ²) on an EulerBeam3D type of this (windowed) path — the accelerator
 * addin!(…,eleobj::ElementCost,…) of src/DirectXUA.jl:172-198 as it is meant (L = Λ∘₁R + cost, first-order R and eleres; see mb_xua_set_gauge_cost for the general form).
 * mb_direct_set_gauge_cost: between mb_add_eulerbeam3d and mb_direct_prepare; G [ngauge][4] with ε_g = G[g]·(εₐₓ,κ₁,κ₂,κ₃) (toolbox/StrainGaugeOnBeamElement.jl:70-76).
 * mb_direct_set_gauge_measurements: measured strains of one stored step, [ngauge] for all elements or [nele][ngauge] (per_element, fixed by the first call); they are
 * kept per stored step and move with mb_direct_rebase. */
int32_t mb_direct_set_gauge_cost(mb_handle* h, int32_t ieletyp, int32_t ngauge, const double* G, double sigma);
int32_t mb_direct_set_gauge_measurements(mb_handle* h, int64_t step, int32_t ieletyp, const double* epsm, int32_t per_element);
int32_t mb_direct_sparser(mb_handle* h, double rtol, int64_t* nnz_out);
int32_t mb_direct_get_sparse(mb_handle* h, int64_t* colptr, int64_t* rowval, double* nzval);
int32_t mb_direct_set_lambda(mb_handle* h, int64_t step, const double* Lambda);
int32_t mb_direct_get_state(mb_handle* h, int64_t step, double* X0, double* X1, double* X2, double* U0, double* Lambda);
int32_t mb_direct_set_dof_scale(mb_handle* h, const double* scaleL, const double* scaleX, const double* scaleU);
int32_t mb_direct_decrement(mb_handle* h, int64_t s0, int64_t s1, const double* dv, double* delta2);
/* Sliding window over the time steps (BASELINE.json configs[3]: 2000 steps do not fit as one materialised Lvv).  Away from the first and last step
 * finitediff (src/FiniteDifferences.jl:8-31) is the central stencil everywhere, so makepattern's block pattern (src/DirectXUA.jl:245-307) and the CSC
 * structure of the owned columns (SparseTools.prepare, src/SparseTools.jl:32-94) of a window [lo,hi) with 3 ≤ lo, hi ≤ nstep−3 repeat for every such window
 * up to a shift of the global row numbers.  mb_direct_rebase moves the handle to [new_lo, new_lo+hi−lo) without rebuilding anything: row_shift (added to
 * the rowval the handle was built with; mb_direct_big_pattern / mb_direct_get_sparse export shifted rows) is returned, states and per-step blocks of steps
 * stored before and after the move are kept, so a window advancing by its own length evaluates only its new steps
 * (mb_direct_set_state + mb_direct_assemble(eval_lo = old stored end, eval_hi = new stored end)).  mb_direct_set_state also accepts device pointers. */
int32_t mb_direct_rebase(mb_handle* h, int64_t new_lo, int64_t* row_shift_out);
/* device pointers of the per-step blocks a neighbouring time-shard needs (halo exchange over NCCL): L2[Λ,X][1,:], L2[Λ,U][1,1], L1[Λ] */
int32_t mb_direct_step_ptrs(mb_handle* h, int64_t step, double** LX, int64_t* nLX, double** LU, int64_t* nLU, double** L1L, int64_t* nL1);
/* CUDA-event timing of the owned steps: ms[0] element kernels + per-step reductions, ms[1] Lvv/Lv build, ms[2] the element kernels alone (float ms[3]) */
int32_t mb_direct_time_dev(mb_handle* h, int32_t reps, float* ms);

/* ---- DirectXUA{OX,OU,IA}, GENERAL form: any dof classes (Λ,X,U,A), IA = 0 or 1, several experiments (csrc/mb_xua.cu) ------------------------------
 * The reference's whole assembly for the all-steps optimisation problem, for element types the beam-specialised path above does not cover (A-dofs, U-costs,
 * user Lagrangians): prepare(AssemblyDirect{OX,OU,IA}) (src/DirectXUA.jl:22-56: asmvec!/asmmat! for all class pairs), makepattern / preparebig (:245-315, with the
 * A block row / column and `for iexp`), SparseTools.prepare (src/SparseTools.jl:32-94: Lvv, Lvvasm, Lvasm — bit-identical structures), assembleA! (:320-326),
 * assemblebig!{:matrices} (:316-356), sparser! and decrementbig! (:357-383).  Classes are numbered 1 Λ, 2 X, 3 U, 4 A as `ind` (:14); experiments and steps 1-based.
 * Element derivatives come in as PACKETS per element type and step: ∇L [nele][Np] and ∇²L [nele][Np][Np] (symmetric; the device reads entry (row,col) at [col][row]) in the order of partials of
 * DirectXUA_lagrangian_addition! (:121-150): Λ (nx), X₀…X_OX (nx each), U₀…U_OU (nu each), A (na, IA = 1 only), scaled as revariate(…,scale) scales them.  Which parts are
 * filled follows the reference's addin! method for the element type: no_second_order types (:85-120) fill ∇L[Λ] = R and the Λ-row / Λ-column of ∇²L only.
 * Acost types (acost = 1; only A dofs) are assembled by mb_xua_add_A from packets with Np = na (unscaled, :70-84) AND, like any type, by mb_xua_add_step from their
 * general packets — src/Assemble.jl:477 (`assemble_!(…,eleobj::Acost,…) = nothing`) does not match the Vector{<:Acost} it is called with; test/TestDirectXUA.jl:107 pins it.
 *   mb_xua_add_eletyp  : dof numbers [nele][n] per class, 1-based within the class (dis.dis[ieletyp].index[iele].X|U|A).
 *   mb_xua_prepare     : flags = 1 Xwhite | 2 XUindep | 4 UAindep | 8 XAindep (:22).  nstep[nexp], dt[nexp] = length.(time), step.(time) (:443).
 *   mb_xua_class_pattern / get_asm / big_pattern / big_asm : the structures, for checking against the reference (test/TestDirectXUA.jl:93-136).
 *   mb_xua_zero; [mb_xua_set_packet… mb_xua_add_A]; for every step [mb_xua_set_packet… mb_xua_add_step] : one assemblebig!.
 *   mb_xua_get_out     : out.L1[α][αder] (beta = 0) / out.L2[α,β][αder,βder].nzval of the last added step. */
int32_t mb_xua_add_eletyp(mb_handle* h, int64_t nele, int32_t nx, int32_t nu, int32_t na, const int64_t* idxX, const int64_t* idxU, const int64_t* idxA,
                          int32_t acost, int32_t* ieletyp_out);
int32_t mb_xua_prepare(mb_handle* h, int32_t OX, int32_t OU, int32_t IA, int64_t ndofX, int64_t ndofU, int64_t ndofA, int32_t nexp, const int64_t* nstep,
                       const double* dt, int32_t flags, int64_t* nbig_out, int64_t* nnzbig_out);
int32_t mb_xua_class_pattern(mb_handle* h, int32_t alpha, int32_t beta, int64_t* nnz, int64_t* colptr, int64_t* rowval);
int32_t mb_xua_get_asm(mb_handle* h, int32_t ieletyp, int32_t alpha, int32_t beta, int64_t* out);
int32_t mb_xua_big_pattern(mb_handle* h, int64_t* colptr, int64_t* rowval);
int32_t mb_xua_big_asm(mb_handle* h, int64_t* nb, int64_t* nblock, int64_t* bcolptr, int64_t* browval, int64_t* boff, int64_t* basm, int64_t* pgr);
int32_t mb_xua_zero(mb_handle* h);
int32_t mb_xua_set_packet(mb_handle* h, int32_t ieletyp, const double* gradL, const double* hessL);
int32_t mb_xua_add_A(mb_handle* h);
int32_t mb_xua_add_step(mb_handle* h, int32_t iexp, int64_t istep);
int32_t mb_xua_get_out(mb_handle* h, int32_t alpha, int32_t beta, int32_t ader, int32_t bder, double* out);
int32_t mb_xua_out_shape(mb_handle* h, int32_t alpha, int32_t beta, int32_t* na, int32_t* nb);
int32_t mb_xua_get_big(mb_handle* h, double* Lvv_nzval, double* Lv);
/* time shards of the general form: each rank adds its own steps into its copy of Lvv / Lv, then this sums them over the ranks (ncclAllReduce, mb_comm_init first) —
 * with them the sums over steps of L1[A], L2[A,A] (src/DirectXUA.jl:321-326,332-352) */
int32_t mb_xua_allreduce_big(mb_handle* h);
int32_t mb_xua_sparser(mb_handle* h, double rtol, int64_t* nnz_out);
int32_t mb_xua_get_sparse(mb_handle* h, int64_t* colptr, int64_t* rowval, double* nzval);
/* DEVICE element types in the general form, and the ElementCost accelerator (src/DirectXUA.jl:172-198) for strain gauges on beams.
 *   mb_xua_add_device_eletyp : an EulerBeam3D, Bar3D or SoilContact type of this handle (mb_add_*; 1-based number among the handle's device types) becomes the next element type;
 *                              its packets come from the device kernels of the beam-specialised path (R, ∂R/∂X_der, ∂R/∂U: no_second_order = Val(true) for beams and bars,
 *                              toolbox/BeamElement.jl:106, BarElement.jl:104; SoilContact on the second-order branch in closed form) — no host work.
 *   mb_xua_set_gauge_cost    : wraps the type in ElementCost{StrainGaugeOnEulerBeam3D} (toolbox/StrainGaugeOnBeamElement.jl) with the cost Σ_g (ε_g − εm_g)²/(2σ²),
 *                              ε_g = E_g·εₐₓ + K1_g·κ₁ + K2_g·κ₂ + K3_g·κ₃ (:62-65,73): G [ngauge][4] = (E,K1,K2,K3).  As the accelerator is meant (its code hands
 *                              DirectXUA_lagrangian_addition! a Lagrangian without Λ-partials, so the reference has no behaviour to copy there): L = Λ∘₁R + cost with first-order R
 *                              and eleres, second-order cost ⇒ ∇²L[X,X] = Jᵀ·∇²cost·J, no Λ·∂²R/∂X² term.
 *   mb_xua_set_gauge_measurements : εm of the step about to be evaluated, [ngauge] or [nele][ngauge].
 *   mb_xua_eval_device       : packets of all device types at the device-resident state[iexp][istep]; NaN → MB_ERR_NAN with the element (where = NULL: asynchronous, no check).
 *   mb_xua_get_packet / mb_xua_get_gauge : the packet as it stands; (εₐₓ,κ) [nele][4], ∂(εₐₓ,κ)/∂X₀ [nele][4][12] (scaled), cost [nele] of a costed type. */
int32_t mb_xua_add_device_eletyp(mb_handle* h, int32_t ieletyp_dev, int32_t* ieletyp_out);
/* model.scaleΛ for the device types on the second-order branch (SoilContact, costed beams); state[iexp][istep].time = t0 + (istep−1)·Δt[iexp] for Bar3D's weight ramp */
int32_t mb_xua_set_lambda_scale(mb_handle* h, double lambda_scale);
int32_t mb_xua_set_time0(mb_handle* h, int32_t iexp, double t0);
int32_t mb_xua_set_gauge_cost(mb_handle* h, int32_t ieletyp, int32_t ngauge, const double* G, double sigma, double lambda_scale);
int32_t mb_xua_set_gauge_measurements(mb_handle* h, int32_t ieletyp, const double* epsm, int32_t per_element);
int32_t mb_xua_eval_device(mb_handle* h, int32_t iexp, int64_t istep, mb_errinfo* where);
int32_t mb_xua_get_packet(mb_handle* h, int32_t ieletyp, double* gradL, double* hessL);
/* CUDA-event time of whole assemblebig! passes over the device element types (eval_device + add_step for every step), ms per pass */
int32_t mb_xua_time_device_pass(mb_handle* h, int32_t reps, float* ms);
int32_t mb_xua_get_gauge(mb_handle* h, int32_t ieletyp, double* e4, double* J, double* cost);
/* state[iexp][istep]: Λ (nX), X (OX+1)·nX, U (OU+1)·nU contiguous by derivative, A (nA) shared by all states (:452); NULL leaves a part untouched.
 * mb_xua_decrement: decrementbig! with dv = Δv in Lv's layout (host or device) → Δ² for Λ, X, U (max over steps of ΣΔβ²) and A. */
int32_t mb_xua_set_state(mb_handle* h, int32_t iexp, int64_t istep, const double* Lambda, const double* X, const double* U, const double* A);
int32_t mb_xua_get_state(mb_handle* h, int32_t iexp, int64_t istep, double* Lambda, double* X, double* U, double* A);
int32_t mb_xua_set_dof_scale(mb_handle* h, const double* sL, const double* sX, const double* sU, const double* sA);
int32_t mb_xua_decrement(mb_handle* h, const double* dv, double* delta2 /* 4 */);

/* ---- sharding over the GPUs of one box (one handle per GPU / process) ------------------------------------------------------------------------- */
/* Device-resident state: the Newton update runs where the state lives, so X only crosses PCIe when the caller asks.
 *   mb_sweepx_set_state / get_state : state.X[1..OX+1] (and state.U[1]) host↔device; pointers may be host or device memory.
 *   mb_sweepx_set_dof_scale         : dis.scaleX per model dof (Xdofgr = allXdofs, src/SweepX.jl:196); default all ones.
 *   mb_sweepx_newmark_decrement     : Newmarkβdecrement!{OX}(state,Δx,Xdofgr,c,firstiter,…) (src/SweepX.jl:98-132) with getdof!/decrement!
 *                                     (src/Assemble.jl:206-233), same operation order (no FMA contraction) ⇒ bit-identical to the reference
 *                                     arithmetic. dx: the solver's Δx (ndofX, host or device). Optional outputs Σ Δx² and Σ Lλ² of the last
 *                                     assembled gradient (the convergence test of src/SweepX.jl:209,214), reduced on the device.
 * After mb_sweepx_set_state, mb_sweepx_assemble may be called with X0 = NULL: it then assembles at the device-resident state. */
int32_t mb_sweepx_set_state(mb_handle* h, int32_t OX, const double* X0, const double* X1, const double* X2, const double* U0);
int32_t mb_sweepx_get_state(mb_handle* h, int32_t OX, double* X0, double* X1, double* X2);
int32_t mb_sweepx_set_dof_scale(mb_handle* h, const double* scaleX);
int32_t mb_sweepx_newmark_decrement(mb_handle* h, int32_t OX, int32_t firstiter, const double* dx, const double* newmark, double* dx2, double* Ll2);

/* getresult(state,req,els) (src/Output.jl:131-181) for one EulerBeam3D element type, all elements at once, values only: the ☼/♢ requestables of
 * residual (toolbox/BeamElement.jl:151-174) and resultants (:28-64). out: nele × 77 doubles,
 *   [0] ε   [1..9] rₛₘ (column-major)   [10..12] ♢κ   then for igp = 1..4 at 13+16(igp−1):  x(3) κgp(3) fᵢ mᵢ(3) fₑ(3) mₑ(3).
 * X0 = NULL reads the device-resident state (mb_sweepx_set_state); out may be host or device memory. */
int32_t mb_beam_results(mb_handle* h, int32_t ieletyp, int32_t OX, const double* X0, const double* X1, const double* X2, double* out);

/* Run this handle's kernels and copies on a caller-owned CUDA stream (e.g. the stream NCCL work is ordered against). */
int32_t mb_set_stream(mb_handle* h, void* cuda_stream);
/* SweepX element-range sharding: entries of the local Lλ / nzval that belong to nodes shared with a neighbouring shard.
 *   send_nz / send_v : 1-based positions in nzval / Lλ packed into the send buffer (nz first, then v)
 *   recv_nz / recv_v : 1-based positions the first n_recv_nz / next n_recv_v received values are ADDED to; a position 0 means
 *                      "ghost coupling, keep in the receive buffer" (rows owned here, columns of the neighbour's interior node). */
int32_t mb_iface_setup(mb_handle* h, int64_t n_send_nz, const int64_t* send_nz, int64_t n_send_v, const int64_t* send_v,
                       int64_t n_recv_nz, const int64_t* recv_nz, int64_t n_recv_v, const int64_t* recv_v);
int32_t mb_iface_pack_dev(mb_handle* h, double* sendbuf_dev);
int32_t mb_iface_unpack_add_dev(mb_handle* h, const double* recvbuf_dev);

/* The interface exchange itself, inside the shim (NCCL on the handle's stream; see "multi-GPU" below): pack → send to rank+1 / receive from rank−1 →
 * unpack-add.  Replaces nothing in the reference (its assemble! is single-process, src/Assemble.jl:470-487); it is what makes the sharded nzval / Lλ
 * equal the unsharded ones on the owned rows.  mb_iface_buffers: the device buffers it uses (the first n_recv_nz received values whose position is 0
 * are the ghost couplings); mb_iface_get_recvbuf copies the receive buffer to the host. */
int32_t mb_iface_exchange(mb_handle* h);
int32_t mb_iface_buffers(mb_handle* h, double** sendbuf_dev, int64_t* nsend, double** recvbuf_dev, int64_t* nrecv);
int32_t mb_iface_get_recvbuf(mb_handle* h, double* out);

/* ---- multi-GPU: one handle per GPU / process, NCCL inside the shim (north_star: Julia → thin shim, no PyTorch) -------------------------------------
 * mb_comm_unique_id  : ncclGetUniqueId — one process creates the 128-byte id and hands it to the others by its own means (file, socket, MPI).
 * mb_comm_init       : ncclCommInitRank on the handle's device; rank r of `world` consecutive shards (element ranges or time-step ranges).
 * mb_comm_allreduce  : small HOST vectors in place (op 0 sum, 1 max, 2 min), blocking — timings, norms of the convergence test (src/SweepX.jl:209,
 *                      src/DirectXUA.jl:493-500), and the A-class sums Σ_step L1[A], L2[A,A] of time shards (src/DirectXUA.jl:321-326).
 * mb_comm_allreduce_dev : the same on device memory, asynchronous on the handle's stream.
 * mb_direct_halo_exchange : time shards (mb_direct_prepare with step_lo/step_hi): the per-step blocks L2[Λ,X] (and L1[X] of second-order element types) of the
 *                      two steps at each end of the owned range go to the neighbours, the halo steps' blocks come in (src/FiniteDifferences.jl:2-4 reach ±2).
 * libnccl.so.2 is opened with dlopen at the first call: MB_ERR_NCCL if it is absent or an NCCL call fails (mb_last_error has the text). */
int32_t mb_comm_unique_id(uint8_t* id128);
int32_t mb_comm_init(mb_handle* h, const uint8_t* id128, int32_t rank, int32_t world);
int32_t mb_comm_share(mb_handle* h, mb_handle* owner);     /* a second handle on the same GPU borrows the owner's communicator (owner must outlive it) */
int32_t mb_comm_destroy(mb_handle* h);
int32_t mb_comm_info(const mb_handle* h, int32_t* rank, int32_t* world, int32_t* nccl_version);
int32_t mb_comm_allreduce(mb_handle* h, double* buf, int64_t n, int32_t op);
int32_t mb_comm_allreduce_dev(mb_handle* h, double* dev, int64_t n, int32_t op);
int32_t mb_comm_barrier(mb_handle* h);
int32_t mb_direct_halo_exchange(mb_handle* h);

/* Page-lock / unlock a host array the caller owns (Julia: the Vector behind out.Lλx.nzval), so that the copies inside
 * mb_sweepx_assemble run at PCIe speed. */
int32_t mb_host_register(mb_handle* h, void* p, int64_t bytes);
int32_t mb_host_unregister(mb_handle* h, void* p);

/* ---- measurement helpers (bench.py) ------------------------------------------------------------------------------------ */
/* Times one launch set of the last assemble configuration with CUDA events on the handle's stream:
 *   ms[0] = element kernels, ms[1] = segmented reduction into nzval/Lλ. */
int32_t mb_sweepx_time_dev(mb_handle* h, int32_t OX, int32_t mission, double t, const double* newmark, int32_t reps, float* ms);
/* On a handle with a communicator and interface lists (element-range shard) every step is mb_sweepx_assemble_dev + mb_iface_exchange. */
/* Times `reps` whole device-resident steps (mb_sweepx_assemble_dev exactly as a solver calls it: for large models the segmented reduction of
 * element chunk j overlaps the element kernels of chunk j+1 on a second, high-priority stream): *ms = average per step. */
int32_t mb_sweepx_time_step_dev(mb_handle* h, int32_t OX, int32_t mission, double t, const double* newmark, int32_t reps, float* ms);
/* Sustained FP64 FMA throughput of this device (dependent-chain-free DFMA loop), TFLOP/s — the FP64 roofline denominator. */
int32_t mb_measure_fp64_tflops(mb_handle* h, double* tflops);
/* Device copy bandwidth (read+write bytes / time), GB/s. */
int32_t mb_measure_copy_gbs(mb_handle* h, double* gbs);
/* nbytes host→device and nbytes device→host at once from / to pinned host buffers (what mb_sweepx_assemble's host-state pipeline moves): ms per round, the e2e ceiling */
int32_t mb_measure_host_copy_ms(mb_handle* h, const void* host_in, void* host_out, int64_t nbytes, int32_t reps, double* ms);
int64_t mb_launch_count(const mb_handle* h);      /* kernels of this library launched so far on this handle */

#ifdef __cplusplus
}
#endif
#endif
