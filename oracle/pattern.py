"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of the reference's integer index structures.

All indices are kept 1-based int64 exactly as the reference stores them, so arrays compare with ``==`` against the
goldens of test/TestAssemble.jl, test/TestDirectXUA.jl and test/TestSparseTools.jl.

Restates (paths relative to the reference root):
  src/Assemble.jl:167-204   DofGroup / allXdofs …            → dofgroup()
  src/Assemble.jl:236-246   indexedstate                      → indexedstate()
  src/Assemble.jl:329-335   gradientpartition                 → gradientpartition()
  src/Assemble.jl:340-357   asmvec! / asmvec_kernel!          → asmvec()
  src/Assemble.jl:373-448   asmmat!                           → asmmat()
  src/SweepX.jl:24-33       prepare(AssemblySweepX{OX})       → prepare_sweepx()
  src/DirectXUA.jl:22-56    prepare(AssemblyDirect{OX,OU,IA}) → prepare_direct()
  src/DirectXUA.jl:245-315  makepattern / preparebig          → makepattern(), preparebig()
  src/SparseTools.jl:32-137 prepare / addin!                  → sparsetools_prepare(), addin_block(), addin_vec()
  src/FiniteDifferences.jl  finitediff                        → finitediff()
"""
import numpy as np

I64 = np.int64


# ------------------------------------------------------------------------------------------------ dof groups
def dofgroup(nX, nU, nA, iL=(), iX=(), iU=(), iA=()):
    """DofGroup(dis,iΛ,iX,iU,iA)  src/Assemble.jl:195-204 (integer part)."""
    iL, iX, iU, iA = (np.asarray(list(v), I64) for v in (iL, iX, iU, iA))
    nl, nx, nu, na = len(iL), len(iX), len(iU), len(iA)
    jL, jX, jU, jA = gradientpartition(nl, nx, nu, na)
    return dict(nX=nX, nU=nU, nA=nA, iL=iL, iX=iX, iU=iU, iA=iA, jL=jL, jX=jX, jU=jU, jA=jA)


def allLdofs(nX, nU, nA): return dofgroup(nX, nU, nA, iL=range(1, nX + 1))
def allXdofs(nX, nU, nA): return dofgroup(nX, nU, nA, iX=range(1, nX + 1))
def allUdofs(nX, nU, nA): return dofgroup(nX, nU, nA, iU=range(1, nU + 1))
def allAdofs(nX, nU, nA): return dofgroup(nX, nU, nA, iA=range(1, nA + 1))


def getndof(gr):
    return len(gr["iL"]) + len(gr["iX"]) + len(gr["iU"]) + len(gr["iA"])


def gradientpartition(nL, nX, nU, nA, nXder=1, nUder=None):
    """src/Assemble.jl:329-335"""
    if nUder is None:
        nUder = nXder
    iL = np.arange(1, nL + 1, dtype=I64)
    iX = nL + np.arange(1, nX + 1, dtype=I64)
    iU = nL + nX * nXder + np.arange(1, nU + 1, dtype=I64)
    iA = nL + nX * nXder + nU * nUder + np.arange(1, nA + 1, dtype=I64)
    return iL, iX, iU, iA


def indexedstate(gr):
    """src/Assemble.jl:236-246"""
    L = np.zeros(gr["nX"], I64); X = np.zeros(gr["nX"], I64); U = np.zeros(gr["nU"], I64); A = np.zeros(gr["nA"], I64)
    L[gr["iL"] - 1] = gr["jL"]; X[gr["iX"] - 1] = gr["jX"]; U[gr["iU"] - 1] = gr["jU"]; A[gr["iA"] - 1] = gr["jA"]
    return L, X, U, A


# ------------------------------------------------------------------------------------------------ asmvec! / asmmat!
def asmvec(dofgr, dis):
    """asmvec!(asm,dofgr,dis)  src/Assemble.jl:340-357.

    dis: list (one per element type) of dicts with 'X','U','A' = (nele,n) 1-based int64 index arrays
    (dis.dis[ieletyp].index[iele].X …).  Returns list asm[ieletyp] of (ndof_in_gradient, nele) arrays."""
    L, X, U, A = indexedstate(dofgr)
    out = []
    for d in dis:
        nele = d["X"].shape[0]
        nL = 0 if len(dofgr["iL"]) == 0 else d["X"].shape[1]      # gradientstructure, src/Assemble.jl:321-327
        nX = 0 if len(dofgr["iX"]) == 0 else d["X"].shape[1]
        nU = 0 if len(dofgr["iU"]) == 0 else d["U"].shape[1]
        nA = 0 if len(dofgr["iA"]) == 0 else d["A"].shape[1]
        a = np.zeros((nL + nX + nU + nA, nele), I64)
        if nL: a[0:nL, :] = L[d["X"] - 1].T
        if nX: a[nL:nL + nX, :] = X[d["X"] - 1].T
        if nU: a[nL + nX:nL + nX + nU, :] = U[d["U"] - 1].T
        if nA: a[nL + nX + nU:, :] = A[d["A"] - 1].T
        out.append(a)
    return out


def asmmat(iasm, jasm, nimoddof, njmoddof):
    """asmmat!(asm,iasm,jasm,nimoddof,njmoddof)  src/Assemble.jl:373-448.

    Returns (asm list, colptr, rowval) with asm[ieletyp][ieledof+nieledof*(jeledof-1), iele] = inz (1-based, 0 = none)."""
    # 2) list all (jmoddof,imoddof) pairs: element types outer, then iele, jeledof, ieledof (ieledof fastest)
    Js, Is, shapes = [], [], []
    for ia, ja in zip(iasm, jasm):
        ni, nele = ia.shape
        nj = ja.shape[0]
        I = np.broadcast_to(ia.T[:, None, :], (nele, nj, ni))      # [iele,jeledof,ieledof] = iasm[ieledof,iele]
        J = np.broadcast_to(ja.T[:, :, None], (nele, nj, ni))
        Is.append(I.reshape(-1)); Js.append(J.reshape(-1)); shapes.append((ni, nj, nele))
    I = np.concatenate(Is) if Is else np.zeros(0, I64)
    J = np.concatenate(Js) if Js else np.zeros(0, I64)
    keep = (I != 0) & (J != 0)
    Ik, Jk = I[keep], J[keep]
    npair = Ik.size
    # 3) sortperm of (j,i) pairs (lexicographic); ties are order-insensitive (K[I[ipair]] = inz)
    order = np.lexsort((Ik, Jk))
    Js_, Is_ = Jk[order], Ik[order]
    # 4) unique entries → nnz, rowval, colptr, K
    new = np.ones(npair, bool)
    if npair > 1:
        new[1:] = (Js_[1:] != Js_[:-1]) | (Is_[1:] != Is_[:-1])
    inz_sorted = np.cumsum(new).astype(I64)
    nnz = int(inz_sorted[-1]) if npair else 0
    rowval = Is_[new].astype(I64)
    colj = Js_[new]
    colptr = np.ones(njmoddof + 1, I64)
    # colptr[icol] = first inz of the first column ≥ icol ; columns after the last used one get nnz+1
    counts = np.bincount(colj - 1, minlength=njmoddof)
    colptr[1:] = 1 + np.cumsum(counts)
    K = np.zeros(npair, I64)
    K[order] = inz_sorted
    # 5) distribute K into asm
    full = np.zeros(I.size, I64)
    full[keep] = K
    out, p = [], 0
    for (ni, nj, nele) in shapes:
        n = ni * nj * nele
        out.append(np.ascontiguousarray(full[p:p + n].reshape(nele, nj * ni).T))   # [ieledof+ni*(jeledof-1), iele]
        p += n
    return out, colptr, rowval


def prepare_sweepx(dis, nX, nU, nA):
    """prepare(AssemblySweepX{OX},model,dis)  src/SweepX.jl:24-33 → asm1, asm2, colptr, rowval."""
    gr = allXdofs(nX, nU, nA)
    asm1 = asmvec(gr, dis)
    asm2, colptr, rowval = asmmat(asm1, asm1, nX, nX)
    return asm1, asm2, colptr, rowval


# ------------------------------------------------------------------------------------------------ DirectXUA
IND = dict(L=1, X=2, U=3, A=4)
NCLASS = 4


def arrnum(a, b=None):
    """src/DirectXUA.jl:16-17"""
    return a if b is None else NCLASS + b + NCLASS * (a - 1)


def prepare_direct(dis, nX, nU, nA, OX, OU, IA, Xwhite=False, XUindep=False, UAindep=False, XAindep=False):
    """prepare(AssemblyDirect{OX,OU,IA},model,dis;…)  src/DirectXUA.jl:22-56.

    Returns dict: asm[k] (k = arrnum, 1-based keys) → list per eletyp; pat[(α,β)] = (m,n,colptr,rowval);
    nL1[α] = number of derivative vectors; nL2[(α,β)] = (nα,nβ) shape of the matrix-of-matrices."""
    grs = (allLdofs(nX, nU, nA), allXdofs(nX, nU, nA), allUdofs(nX, nU, nA), allAdofs(nX, nU, nA))
    ndof = [getndof(g) for g in grs]
    nder = (1, OX + 1, OU + 1, IA)
    asm, pat, nL2 = {}, {}, {}
    for a in range(1, 5):
        asm[arrnum(a)] = asmvec(grs[a - 1], dis)
    for a in range(1, 5):
        for b in range(1, 5):
            am, colptr, rowval = asmmat(asm[arrnum(a)], asm[arrnum(b)], ndof[a - 1], ndof[b - 1])
            asm[arrnum(a, b)] = am
            pat[(a, b)] = (ndof[a - 1], ndof[b - 1], colptr, rowval)
            na, nb = nder[a - 1], nder[b - 1]
            if a == b == IND["L"]: na, nb = 0, 0
            if Xwhite and a == b == IND["X"]: na, nb = 1, 1
            if XUindep and {a, b} == {IND["X"], IND["U"]}: na, nb = 0, 0
            if XAindep and {a, b} == {IND["X"], IND["A"]}: na, nb = 0, 0
            if UAindep and {a, b} == {IND["U"], IND["A"]}: na, nb = 0, 0
            nL2[(a, b)] = (na, nb)
    return dict(asm=asm, pat=pat, nL1={a: nder[a - 1] for a in range(1, 5)}, nL2=nL2, ndof=ndof)


_FD = [[[(0, 1.)]],
       [[(0, -1.), (1, 1.)], [(-1, -1.), (0, 1.)], [(-1, -.5), (1, .5)]],
       [[(0, 1.), (1, -2.), (2, 1.)], [(-2, 1.), (-1, -2.), (0, 1.)], [(-1, 1.), (0, -2.), (1, 1.)]]]


def finitediff(order, n, s):
    """finitediff(order,n,s)  src/FiniteDifferences.jl:8-31 → list of (Δs,w); s is 1-based."""
    if order > 0 and n < 6:
        raise ValueError("Number of steps must be ≥6")
    if order == 0:
        return _FD[0][0]
    k = 0 if s == 1 else (1 if s == n else 2)
    return _FD[order][k]


def makepattern(IA, nstep, nL2, pat):
    """makepattern(IA,nstep,out)  src/DirectXUA.jl:245-307 → sorted unique list of (αblk,βblk,(α,β)) block coordinates
    (column-major order, as `sparse(αblk[u],βblk[u],nz[u])` stores them) — the block stored is Lαβ[1,1]'s pattern."""
    blocks = {}
    cumblk = 0
    for ns in nstep:
        for istep in range(1, ns + 1):
            for a in (1, 2, 3):
                for b in (1, 2, 3):
                    na, nb = nL2[(a, b)]
                    for ad in range(1, na + 1):
                        for bd in range(1, nb + 1):
                            for (das, _) in finitediff(ad - 1, ns, istep):
                                for (dbs, _) in finitediff(bd - 1, ns, istep):
                                    key = (cumblk + 3 * (istep + das - 1) + a, cumblk + 3 * (istep + dbs - 1) + b)
                                    blocks.setdefault(key, (a, b))
        cumblk += 3 * ns
    if IA == 1:
        Ablk = 3 * sum(nstep) + 1
        blocks.setdefault((Ablk, Ablk), (4, 4))
        cumblk = 0
        for ns in nstep:
            for istep in range(1, ns + 1):
                for a in (1, 2, 3):
                    if nL2[(4, a)][0] > 0:
                        blocks.setdefault((Ablk, cumblk + 3 * (istep - 1) + a), (4, a))
                        blocks.setdefault((cumblk + 3 * (istep - 1) + a, Ablk), (a, 4))
            cumblk += 3 * ns
    keys = sorted(blocks.keys(), key=lambda k: (k[1], k[0]))     # CSC order: column, then row
    return [(k[0], k[1], blocks[k]) for k in keys]


def sparsetools_prepare(nbr, nbc, blocks):
    """prepare(pattern::SparseMatrixCSC{SparseMatrixCSC})  src/SparseTools.jl:32-94.

    blocks: list of (ibr, ibc, (m, n, colptr, rowval)) in CSC (column-major) order of the pattern, 1-based.
    Returns dict(m,n,colptr,rowval) of bigmat, bigmatasm = dict(colptr,rowval,nzval=list of int64 arrays), pgr, pgc."""
    nlr = -np.ones(nbr + 1, I64); nlc = -np.ones(nbc + 1, I64)
    pcol = np.ones(nbc + 1, I64)
    for (ibr, ibc, (m, n, cp, rv)) in blocks:
        assert nlr[ibr] in (-1, m) and nlc[ibc] in (-1, n)
        nlr[ibr] = m; nlc[ibc] = n
        pcol[ibc] += 1
    pcol = np.concatenate([[1], 1 + np.cumsum(pcol[1:] - 1)]).astype(I64)
    nlr[0] = 1; nlc[0] = 1
    if (nlr == -1).any() or (nlc == -1).any():
        raise ValueError("invalid sparse-of-sparse pattern")
    pgr = np.cumsum(nlr); pgc = np.cumsum(nlc)
    ngr, ngc = int(pgr[-1] - 1), int(pgc[-1] - 1)
    ngv = sum(len(b[2][3]) for b in blocks)
    pigr = np.zeros(ngc + 1, I64); igr = np.zeros(ngv, I64)
    asm_igv = [np.zeros(len(b[2][3]), I64) for b in blocks]
    pigr[0] = 1
    igv = 1
    for ibc in range(1, nbc + 1):
        for ilc in range(1, int(nlc[ibc]) + 1):
            igc = pgc[ibc - 1] - 1 + ilc
            for ibv in range(int(pcol[ibc - 1]), int(pcol[ibc])):
                ibr, _, (m, n, cp, rv) = blocks[ibv - 1]
                lo, hi = int(cp[ilc - 1]), int(cp[ilc])
                cnt = hi - lo
                igr[igv - 1:igv - 1 + cnt] = pgr[ibr - 1] - 1 + rv[lo - 1:hi - 1]
                asm_igv[ibv - 1][lo - 1:hi - 1] = np.arange(igv, igv + cnt)
                igv += cnt
            pigr[igc] = igv
    big = dict(m=ngr, n=ngc, colptr=pigr, rowval=igr)
    bigasm = dict(colptr=pcol, rowval=np.array([b[0] for b in blocks], I64), nzval=asm_igv)
    return big, bigasm, pgr, pgc


def find_block(bigasm, ibr, ibc):
    """dichotomy of addin!  src/SparseTools.jl:101-117 → 0-based index into bigasm.nzval"""
    lo = int(bigasm["colptr"][ibc - 1]); hi = int(bigasm["colptr"][ibc]) - 1
    while True:
        ibv = (lo + hi) // 2
        aibr = int(bigasm["rowval"][ibv - 1])
        if ibr == aibr:
            return ibv - 1
        if hi <= lo:
            raise KeyError("BlockSparseAssembler pattern has no block [%d,%d]" % (ibr, ibc))
        if ibr > aibr: lo = ibv + 1
        else: hi = ibv - 1


def addin_block(bigasm, out_nzval, block_nzval, ibr, ibc, factor=1.0):
    """addin!(asm,out,block,ibr,ibc,factor)  src/SparseTools.jl:101-122"""
    aigv = bigasm["nzval"][find_block(bigasm, ibr, ibc)]
    np.add.at(out_nzval, aigv - 1, block_nzval * factor)


def addin_vec(pgr, out, block, ibr, factor=1.0):
    """addin!(pgr,out,block,ibr,factor)  src/SparseTools.jl:133-137"""
    out[pgr[ibr - 1] - 1:pgr[ibr] - 1] += block * factor


def preparebig(IA, nstep, nL2, pat):
    """preparebig(IA,nstep,out)  src/DirectXUA.jl:308-315"""
    blk = makepattern(IA, nstep, nL2, pat)
    nb = 3 * sum(nstep) + (1 if IA == 1 else 0)
    blocks = [(r, c, pat[ab]) for (r, c, ab) in blk]
    return sparsetools_prepare(nb, nb, blocks)


def assemblebig(IA, nstep, dt, P, big, bigasm, pgr, outs):
    """assemblebig!{:matrices}(Lvv,Lv,Lvvasm,Lvasm,…)  src/DirectXUA.jl:316-356 for one experiment and IA = 0.

    outs[istep-1] = per-step `out`: dict(L1={α: (nder, ndofα) or (ndofα,)}, L2={(α,β): (nαder·nβder?, nnz)}) with only the blocks that carry
    values; L2[(α,β)] is indexed [βder-1] for α==Λ, [αder-1] for β==Λ, and [(αder-1, βder-1)] tuples otherwise (host costs).
    Returns (Lvv.nzval, Lv)."""
    assert IA == 0
    nz = np.zeros(len(big["rowval"])); Lv = np.zeros(big["m"])
    nL2 = P["nL2"]; nL1 = P["nL1"]
    for istep in range(1, nstep + 1):
        out = outs[istep - 1]
        for b in (1, 2, 3):
            for bder in range(1, nL1[b] + 1):
                v = out["L1"].get(b)
                if v is None:
                    continue
                v = np.atleast_2d(v)
                if bder > v.shape[0]:
                    continue
                s = dt ** (1 - bder)
                for (ds, w) in finitediff(bder - 1, nstep, istep):
                    addin_vec(pgr, Lv, v[bder - 1], 3 * (istep + ds - 1) + b, w * s)
        for a in (1, 2, 3):
            for b in (1, 2, 3):
                na, nb = nL2[(a, b)]
                blk = out["L2"].get((a, b))
                for ad in range(1, na + 1):
                    for bd in range(1, nb + 1):
                        if blk is None:
                            continue
                        if a == 1: val = blk[bd - 1]
                        elif b == 1: val = blk[ad - 1]
                        else: val = blk.get((ad, bd)) if isinstance(blk, dict) else None
                        if val is None:
                            continue
                        s = dt ** (2 - ad - bd)
                        for (das, wa) in finitediff(ad - 1, nstep, istep):
                            for (dbs, wb) in finitediff(bd - 1, nstep, istep):
                                addin_block(bigasm, nz, val, 3 * (istep + das - 1) + a, 3 * (istep + dbs - 1) + b, wa * wb * s)
    return nz, Lv


# ------------------------------------------------------------------------------------------------ sparser! / decrementbig!
def sparser(colptr, rowval, nzval, rtol=1e-9):
    """sparser!(T,S,rtol)  src/SparseTools.jl:172-199 ; 1-based colptr/rowval in and out"""
    atol = rtol * np.abs(nzval).max()
    keep = np.abs(nzval) >= atol
    ndrop_before = np.concatenate([[0], np.cumsum(~keep)])
    return colptr - ndrop_before[colptr - 1], rowval[keep], nzval[keep]


def decrementbig(states, dv, OX, OU, dt, nstep, nX, nU, scaleL, scaleX, scaleU):
    """decrementbig!(state,Δ²,Lvdis,dofgr,Δv,nder,Δt,nstep)  src/DirectXUA.jl:357-383, one experiment, IA=0.
    states[istep] = dict(L=[Λ], X=[X0..], U=[U0..]) mutated in place; dv in the block order of Lv (per step Λ, X, U). Returns Δ²[3]."""
    W = 2 * nX + nU
    off = {1: 0, 2: nX, 3: 2 * nX}; n = {1: nX, 2: nX, 3: nU}
    key = {1: "L", 2: "X", 3: "U"}; sc = {1: scaleL, 2: scaleX, 3: scaleU}; nder = {1: 1, 2: OX + 1, 3: OU + 1}
    d2 = np.zeros(3)
    inv = 1.0 / dt
    dtp = [1.0, inv, inv * inv]                      # Δt^(1−βder)
    for istep in range(1, nstep + 1):
        for b in (1, 2, 3):
            for bder in range(1, nder[b] + 1):
                for (ds, w) in finitediff(bder - 1, nstep, istep):
                    blk = (istep + ds - 1) * W + off[b]
                    db = dv[blk: blk + n[b]]
                    states[istep - 1][key[b]][bder - 1] -= ((db * w) * dtp[bder - 1]) * sc[b]
                    if bder == 1:
                        d2[b - 1] = max(d2[b - 1], float((db * db).sum()))
    return d2


# ------------------------------------------------------------------------------------------------ DirectXUA, general form (IA = 0/1, several experiments)
def out_zeros(P):
    """zero!(out::AssemblyDirect) with the shapes of prepare (src/DirectXUA.jl:22-56): L1[α] (nder α, ndof α); L2[(α,β)] (nα, nβ, nnz)"""
    nnz = {ab: len(P["pat"][ab][3]) for ab in P["pat"]}
    return dict(L1={a: np.zeros((P["nL1"][a], P["ndof"][a - 1])) for a in range(1, 5)},
                L2={ab: np.zeros((P["nL2"][ab][0], P["nL2"][ab][1], nnz[ab])) for ab in P["pat"]})


def lagrangian_addition(P, OX, OU, IA, ityp, gradL, hessL, out):
    """DirectXUA_lagrangian_addition!{:matrices}  src/DirectXUA.jl:121-150 for every element of type `ityp` (0-based):
    gradL (nele,Np), hessL (nele,Np,Np) = ∂{2,Np}(L) in the order of partials Λ, X₀…, U₀…, A.  Element loop outermost, as assemble_! runs it."""
    asm = P["asm"]
    nele = gradL.shape[0]
    ndofe = [asm[arrnum(a)][ityp].shape[0] for a in range(1, 5)]          # Nx, Nx, Nu, Na
    nder = (1, OX + 1, OU + 1, IA)
    for e in range(nele):
        pa = 0
        for a in range(1, 5):
            for i in range(1, nder[a - 1] + 1):
                ia = pa + np.arange(ndofe[a - 1]); pa += ndofe[a - 1]
                La = out["L1"][a]
                if i <= La.shape[0]:
                    av = asm[arrnum(a)][ityp][:, e]
                    for k in range(ndofe[a - 1]):
                        if av[k]: La[i - 1, av[k] - 1] += gradL[e, ia[k]]
                pb = 0
                for b in range(1, 5):
                    for j in range(1, nder[b - 1] + 1):
                        ib = pb + np.arange(ndofe[b - 1]); pb += ndofe[b - 1]
                        Lab = out["L2"][(a, b)]
                        if i <= Lab.shape[0] and j <= Lab.shape[1]:
                            am = asm[arrnum(a, b)][ityp][:, e]
                            ni = ndofe[a - 1]
                            for kb in range(ndofe[b - 1]):
                                for ka in range(ni):
                                    inz = am[ka + ni * kb]
                                    if inz: Lab[i - 1, j - 1, inz - 1] += hessL[e, ia[ka], ib[kb]]
    return out


def assemblebig_general(IA, nstep, dt, P, big, bigasm, pgr, outA, outs):
    """assemblebig!{:matrices}  src/DirectXUA.jl:316-356, all experiments, IA = 0 or 1.
    outA: the out of assembleA! (None when IA = 0); outs[iexp][istep-1]: the out of assemble! at that step (shapes of out_zeros).  → (Lvv.nzval, Lv)"""
    nz = np.zeros(len(big["rowval"])); Lv = np.zeros(big["m"])
    nL2 = P["nL2"]
    Ablk = 3 * sum(nstep) + 1
    if IA == 1:
        addin_vec(pgr, Lv, outA["L1"][4][0], Ablk)
        addin_block(bigasm, nz, outA["L2"][(4, 4)][0, 0], Ablk, Ablk)
    classes = (1, 2, 3, 4) if IA == 1 else (1, 2, 3)
    cumblk = 0
    for iexp, ns in enumerate(nstep):
        for istep in range(1, ns + 1):
            out = outs[iexp][istep - 1]
            for b in classes:
                Lb = out["L1"][b]
                for bder in range(1, Lb.shape[0] + 1):
                    s = dt[iexp] ** (1 - bder)
                    for (ds, w) in finitediff(bder - 1, ns, istep):
                        blk = Ablk if b == 4 else cumblk + 3 * (istep + ds - 1) + b
                        addin_vec(pgr, Lv, Lb[bder - 1], blk, w * s)
            for a in classes:
                for b in classes:
                    Lab = out["L2"][(a, b)]
                    for ad in range(1, Lab.shape[0] + 1):
                        for bd in range(1, Lab.shape[1] + 1):
                            s = dt[iexp] ** (2 - ad - bd)
                            for (das, wa) in finitediff(ad - 1, ns, istep):
                                for (dbs, wb) in finitediff(bd - 1, ns, istep):
                                    ablk = Ablk if a == 4 else cumblk + 3 * (istep + das - 1) + a
                                    bblk = Ablk if b == 4 else cumblk + 3 * (istep + dbs - 1) + b
                                    addin_block(bigasm, nz, Lab[ad - 1, bd - 1], ablk, bblk, wa * wb * s)
        cumblk += 3 * ns
    return nz, Lv


def decrementbig_general(states, A, dv, OX, OU, IA, dt, nstep, nX, nU, nA, scaleL, scaleX, scaleU, scaleA):
    """decrementbig!  src/DirectXUA.jl:357-383, all experiments.  states[iexp][istep-1] = dict(L=[Λ], X=[X0..], U=[U0..]) mutated in place, A mutated in place
    (all states share it); dv in the block order of Lv.  Returns Δ² (4 entries; the A entry is 0 when IA = 0)."""
    W = 2 * nX + nU
    off = {1: 0, 2: nX, 3: 2 * nX}; n = {1: nX, 2: nX, 3: nU}
    key = {1: "L", 2: "X", 3: "U"}; sc = {1: scaleL, 2: scaleX, 3: scaleU}; nder = {1: 1, 2: OX + 1, 3: OU + 1}
    d2 = np.zeros(4)
    cum = 0
    for iexp, ns in enumerate(nstep):
        inv = 1.0 / dt[iexp]
        dtp = [1.0, inv, inv * inv]
        for istep in range(1, ns + 1):
            for b in (1, 2, 3):
                for bder in range(1, nder[b] + 1):
                    for (ds, w) in finitediff(bder - 1, ns, istep):
                        blk = (cum + istep + ds - 1) * W + off[b]
                        db = dv[blk: blk + n[b]]
                        states[iexp][istep - 1][key[b]][bder - 1] -= ((db * w) * dtp[bder - 1]) * sc[b]
                        if bder == 1:
                            d2[b - 1] = max(d2[b - 1], float((db * db).sum()))
        cum += ns
    if IA == 1:
        da = dv[cum * W: cum * W + nA]
        d2[3] = float((da * da).sum())
        A -= da * scaleA
    return d2
