"""The product's Riemann solvers (idefix_b200/csrc/riemann_*.cuh -- the functions the sweep kernels inline)
compiled for the HOST by nvcc and compared BIT FOR BIT with the oracle's restatement on random and on
deliberately degenerate left / right states.  Runs without a GPU: it pins the solver source itself,
including branches the golden runs rarely visit (supersonic sides, HLLD degeneracies and its HLLC-type
fallback, the Roe entropy fix), for every solver x EOS x direction x number of components."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle.godunov import Oracle, lib as oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOLVERS = {"tvdlf": 0, "hll": 1, "hllc": 2, "hlld": 3, "roe": 4}


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("riemann_host") / "libriemann_host.so"
    cmd = ["nvcc", "-std=c++17", "-O2", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-I", os.path.join(ROOT, "idefix_b200", "csrc"),
           "-o", str(out), os.path.join(ROOT, "tests", "host_cpp", "riemann_host.cu")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    lib = C.CDLL(str(out))
    lib.idfx_host_riemann.argtypes = ([C.c_int] * 6 + [C.c_double, C.c_double, C.c_long] + [C.c_void_p] * 4
                                      + [C.c_int, C.c_void_p])
    lib.idfx_host_reconstruct.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    lib.idfx_host_convert.argtypes = [C.c_int] * 3 + [C.c_double, C.c_double, C.c_long] + [C.c_void_p] * 3
    lib.idfx_host_reconstruct_irregular.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_int] + [C.c_void_p] * 4
    lib.idfx_host_geometric_source.argtypes = ([C.c_int] * 5 + [C.c_double] * 3 + [C.c_long] + [C.c_void_p] * 8)
    lib.idfx_host_curvilinear_rhs.argtypes = ([C.c_int] * 6 + [C.c_double] * 3 + [C.c_int] + [C.c_void_p] * 10)
    return lib


def states(rng, n, ncomp, mhd, iso, dir):
    """(n, nvar) left / right primitive states in the reference's variable order: a random bulk plus blocks of
    special pairs (identical states, supersonic to either side, vanishing normal / transverse field, strong jumps)"""
    nvar = 1 + ncomp + (0 if iso else 1) + (ncomp if mhd else 0)
    PRS = None if iso else ncomp + 1
    BX1 = ncomp + (1 if iso else 2)

    def draw(m):
        v = np.empty((m, nvar))
        v[:, 0] = 10.0 ** rng.uniform(-1.5, 1.0, m)
        v[:, 1:1 + ncomp] = rng.uniform(-1.5, 1.5, (m, ncomp))
        if PRS is not None:
            v[:, PRS] = 10.0 ** rng.uniform(-2.0, 1.0, m)
        if mhd:
            v[:, BX1:BX1 + ncomp] = rng.uniform(-1.5, 1.5, (m, ncomp))
        return v

    m = n // 8
    vL, vR = draw(n), draw(n)
    vR[0:m] = vL[0:m]                                            # no jump at all
    vL[m:2 * m, 1 + dir] += 8.0; vR[m:2 * m, 1 + dir] += 8.0     # everything moves right: sl > 0
    vL[2 * m:3 * m, 1 + dir] -= 8.0; vR[2 * m:3 * m, 1 + dir] -= 8.0   # sr < 0
    vR[3 * m:4 * m] = vL[3 * m:4 * m] * (1.0 + 1e-9 * rng.standard_normal((m, nvar)))   # nearly equal states
    if mhd:
        vL[4 * m:5 * m, BX1:BX1 + ncomp] = 0.0                   # field-free left state (and zero normal field)
        vR[4 * m:5 * m, BX1 + dir] = 0.0
        for c in range(ncomp):
            if c != dir:                                          # purely normal field: HLLD degeneracy
                vL[5 * m:6 * m, BX1 + c] = 0.0
                vR[5 * m:6 * m, BX1 + c] = 0.0
        vL[6 * m:7 * m, BX1:BX1 + ncomp] *= 1e-7                  # weak field
        vR[6 * m:7 * m, BX1:BX1 + ncomp] *= 1e-7
        vR[:, BX1 + dir] = vL[:, BX1 + dir]                       # the sweeps impose the face value on both sides
    if PRS is not None:
        vL[7 * m:8 * m, PRS] *= 1e3                               # strong shock
    vL[7 * m:8 * m, 1 + dir] = np.abs(vL[7 * m:8 * m, 1 + dir]) + 1.0   # converging flow
    vR[7 * m:8 * m, 1 + dir] = -np.abs(vR[7 * m:8 * m, 1 + dir])
    return vL, vR


CASES = [(mhd, iso, solver, ncomp)
         for mhd in (False, True) for iso in (False, True)
         for solver in (("tvdlf", "hll", "hlld", "roe") if mhd else ("tvdlf", "hll", "hllc", "roe"))
         for ncomp in (1, 2, 3)]


@pytest.mark.parametrize("mhd,iso,solver,ncomp", CASES,
                         ids=[f"{'mhd' if m else 'hd'}-{'iso' if i else 'adi'}-{s}-{c}c" for m, i, s, c in CASES])
def test_product_solver_source_equals_oracle_bitwise(host_lib, mhd, iso, solver, ncomp):
    gamma, csiso, n = 5.0 / 3.0, 0.7, 40000
    rng = np.random.default_rng(1234 + 17 * ncomp + SOLVERS[solver])
    o = Oracle(ncomp, ncomp, mhd, solver, "uct_contact", 2, [4] * ncomp + [1] * (3 - ncomp), ["periodic"] * ncomp,
               gamma, 0.5, 1e-4, [0.1, 0.1, 0.1], csiso=csiso if iso else None)
    for dir in range(ncomp):
        vL, vR = states(rng, n, ncomp, mhd, iso, dir)
        Fo, auxo = o.riemann(dir, vL, vR)
        Fp, auxp = np.zeros_like(vL), np.zeros((n, 4))
        rc = host_lib.idfx_host_riemann(ncomp, ncomp, int(mhd), int(iso), SOLVERS[solver], dir, gamma, csiso, n,
                                        vL.ctypes.data, vR.ctypes.data, Fp.ctypes.data, auxp.ctypes.data, 0, None)
        assert rc == 0
        assert np.isfinite(Fo).all(), "the state generator produced an unphysical pair"
        bad = np.flatnonzero((Fo.view(np.int64) != Fp.view(np.int64)).any(axis=1))
        assert bad.size == 0, (f"dir {dir}: {bad.size} of {n} fluxes differ, first at {bad[0]}: vL={vL[bad[0]]} "
                               f"vR={vR[bad[0]]} oracle={Fo[bad[0]]} product={Fp[bad[0]]}")
        assert (auxo[:, 0].view(np.int64) == auxp[:, 0].view(np.int64)).all(), "cmax differs"
        if not mhd and solver == "roe":
            # (riemann_roe_hd.cuh, not dispatched by the sweeps yet) strong-shock switch: HLL flux in 2-D / 3-D; in 1-D
            # the reference keeps the previous flux, which both sides report through a flag
            assert (auxo[:, 3] == auxp[:, 3]).all()
            assert (auxo[:, 3].sum() > 0) == (ncomp == 1), "the strong-shock switch was not exercised"
        if mhd:   # outermost speeds feed the UCT-HLL / UCT-HLLD coefficients
            assert (auxo[:, 1:3].view(np.int64) == auxp[:, 1:3].view(np.int64)).all(), "sl / sr differ"
    o.close()


EMFS = {"arithmetic": 0, "uct0": 1, "uct_contact": 2, "uct_hll": 3, "uct_hlld": 4}
EMF_CASES = [(nd, nc, iso, solver, emf)
             for nd, nc in ((2, 2), (3, 3), (2, 3)) for iso in (False, True)
             for solver, emfs in (("hlld", ("arithmetic", "uct_contact", "uct_hll", "uct_hlld")),
                                  ("roe", ("uct0", "uct_contact", "uct_hlld")), ("hll", ("uct0", "uct_hll")),
                                  ("tvdlf", ("arithmetic", "uct_contact")))
             for emf in emfs]


@pytest.mark.parametrize("nd,nc,iso,solver,emf", EMF_CASES,
                         ids=[f"{d}d{c}c-{'iso' if i else 'adi'}-{s}-{e}" for d, c, i, s, e in EMF_CASES])
def test_emf_storage_source_equals_oracle_bitwise(host_lib, nd, nc, iso, solver, emf):
    """what each interface hands to constrained transport (face EMFs, contact weights, UCT-HLL / UCT-HLLD
    coefficients; emf_store.cuh) for 2-D, 3-D and 2.5-D"""
    gamma, csiso, n = 1.4, 0.9, 20000
    rng = np.random.default_rng(99 + 7 * nd + nc + SOLVERS[solver] + 31 * EMFS[emf])
    o = Oracle(nd, nc, True, solver, emf, 2, [4] * nd + [1] * (3 - nd), ["periodic"] * nd, gamma, 0.5, 1e-4,
               [0.1, 0.1, 0.1], csiso=csiso if iso else None)
    for dir in range(nd):
        vL, vR = states(rng, n, nc, True, iso, dir)
        Fo, auxo, eo = o.riemann(dir, vL, vR, emf=True)
        Fp, auxp, ep = np.zeros_like(vL), np.zeros((n, 4)), np.zeros((n, 7))
        rc = host_lib.idfx_host_riemann(nd, nc, 1, int(iso), SOLVERS[solver], dir, gamma, csiso, n, vL.ctypes.data,
                                        vR.ctypes.data, Fp.ctypes.data, auxp.ctypes.data, EMFS[emf], ep.ctypes.data)
        assert rc == 0
        assert (Fo.view(np.int64) == Fp.view(np.int64)).all()
        # NaN coefficients (0/0 for two identical states at rest) must be NaN on both sides; everything else bitwise
        same = (eo.view(np.int64) == ep.view(np.int64)) | (np.isnan(eo) & np.isnan(ep))
        bad = np.flatnonzero(~same.all(axis=1))
        assert bad.size == 0, (f"dir {dir}: {bad.size} of {n} differ, first at {bad[0]}: vL={vL[bad[0]]} vR={vR[bad[0]]} "
                               f"oracle={eo[bad[0]]} product={ep[bad[0]]}")
        assert np.nansum(np.abs(eo)) > 0
    o.close()


@pytest.mark.parametrize("order", [2, 3, 4])
@pytest.mark.parametrize("positive", [0, 1])
def test_reconstruction_source_equals_oracle_bitwise(host_lib, order, positive):
    """recon.cuh evaluates the limited profile once per CELL (the pencil sweeps reuse it for both faces); the
    reference and the oracle evaluate it per INTERFACE -- same bits on smooth data, extrema, jumps, plateaus
    and (positive = 1: density / pressure) on profiles that would reconstruct to negative values"""
    rng = np.random.default_rng(5 + order)
    n = 200000
    q = np.cumsum(rng.standard_normal(n)) * 0.1 + np.sin(np.arange(n) * 0.05)
    q[n // 4:n // 4 + 2000] = 1.0                                        # plateau
    q[n // 2:n // 2 + 5000] = 10.0 ** rng.uniform(-6, 2, 5000)           # wild positive jumps
    q[3 * n // 4:3 * n // 4 + 5000] = np.abs(q[3 * n // 4:3 * n // 4 + 5000]) * rng.choice([1e-8, 1.0], 5000)
    if positive:
        q = np.abs(q) + 1e-12
    q = np.ascontiguousarray(q)
    out = [np.zeros(n) for _ in range(4)]
    oracle_lib().oracle_reconstruct(order, n, q.ctypes.data, positive, 0.01, out[0].ctypes.data, out[1].ctypes.data)
    assert host_lib.idfx_host_reconstruct(order, n, q.ctypes.data, positive, 0.01, out[2].ctypes.data,
                                          out[3].ctypes.data) == 0
    assert np.abs(out[0]).sum() > 0
    assert (out[0].view(np.int64) == out[2].view(np.int64)).all(), "left states differ"
    assert (out[1].view(np.int64) == out[3].view(np.int64)).all(), "right states differ"


@pytest.mark.parametrize("ncomp,mhd,iso", [(1, 0, 0), (2, 0, 0), (3, 0, 0), (1, 0, 1), (3, 0, 1), (1, 1, 0), (2, 1, 0),
                                           (3, 1, 0), (2, 1, 1), (3, 1, 1)])
def test_conversion_source_equals_oracle_bitwise(host_lib, ncomp, mhd, iso):
    """K_ConsToPrim (with the pressure floor rewriting the energy) followed by K_PrimToCons (phys.cuh)"""
    rng = np.random.default_rng(77 + ncomp + 2 * mhd + 4 * iso)
    n = 100000
    nvar = 1 + ncomp + (0 if iso else 1) + (ncomp if mhd else 0)
    u = rng.uniform(-2.0, 2.0, (n, nvar))
    u[:, 0] = 10.0 ** rng.uniform(-3, 2, n)
    if not iso:
        u[:, ncomp + 1] = rng.uniform(-0.5, 20.0, n)      # some energies below kinetic + magnetic: the floor fires
    o = Oracle(ncomp, ncomp, bool(mhd), "hll", "uct_contact", 2, [4] * ncomp + [1] * (3 - ncomp), ["periodic"] * ncomp,
               1.4, 0.5, 1e-4, [0.1, 0.1, 0.1], csiso=0.8 if iso else None)
    uo, vo, u2o = u.copy(), np.zeros_like(u), np.zeros_like(u)
    up, vp, u2p = u.copy(), np.zeros_like(u), np.zeros_like(u)
    oracle_lib().oracle_convert(o.h, n, uo.ctypes.data, vo.ctypes.data, u2o.ctypes.data)
    assert host_lib.idfx_host_convert(ncomp, mhd, iso, 1.4, 0.8, n, up.ctypes.data, vp.ctypes.data, u2p.ctypes.data) == 0
    if not iso:
        assert (uo != u).any(), "the pressure floor never fired"
    for a, b in ((uo, up), (vo, vp), (u2o, u2p)):
        assert (a.view(np.int64) == b.view(np.int64)).all()
    o.close()


# ---- curvilinear geometry: the oracle now restates it (pinned bitwise on the reference builds by tests/test_oracle.py),
# so the product's irregular-grid reconstruction and geometric source terms can be checked on the CPU as well ----------
def _spherical_oracle(n1, n2, ncomp=3, mhd=True, order=2, geometry=4, csiso=None, rng=None):
    """a 2-D (r, theta) block on a randomly stretched grid; returns (oracle, coordinates with ghosts)"""
    import math
    ng = 3 if order == 4 else 2
    grid = []
    for a, (n, lo, hi) in enumerate(((n1, 1.0, 4.0), (n2, 0.3, 2.8), (1, 0.0, 1.0))):
        g = ng if a < 2 else 0
        w = rng.uniform(0.5, 1.5, n + 2 * g) if a < 2 else np.ones(1)
        edges = np.concatenate([[0.0], np.cumsum(w)])
        edges = lo + (hi - lo) * (edges - edges[g]) / (edges[g + n] - edges[g])
        xl, xr = edges[:-1].copy(), edges[1:].copy()
        grid.append((0.5 * (xl + xr), xl, xr, xr - xl))
    o = Oracle(2, ncomp, mhd, "hll", "uct_contact", 2, [n1, n2, 1], ["outflow", "outflow"], 1.4, 0.5, 1e-4, [0.1, 0.1, 0.1],
               order=order, geometry=geometry, grid=grid, csiso=csiso)
    return o, grid, ng


@pytest.mark.parametrize("order", [2, 3, 4])
@pytest.mark.parametrize("positive", [0, 1])
def test_irregular_reconstruction_source_equals_oracle_bitwise(host_lib, order, positive):
    """recon.cuh on a non-uniform grid (the PLM weights of extrapolateToFaces.hpp:43-71, LimO3 with the local cell width):
    product source per cell vs the oracle per interface, on a randomly stretched spherical radial grid"""
    rng = np.random.default_rng(11 + order)
    n1 = 4000
    o, grid, ng = _spherical_oracle(n1, 8, order=order, rng=rng)
    x, xl, xr, dx = grid[0]
    n = n1 + 2 * ng
    # the weights as the product's host code forms them (block.cu::buildGeometry, spherical radial direction)
    gc = x + 2.0 * x * dx * dx / (12.0 * x * x + dx * dx)
    w6 = np.zeros((n, 6))
    i = np.arange(1, n - 1)
    w6[i, 0] = dx[i] / (gc[i + 1] - gc[i])
    w6[i, 1] = dx[i] / (gc[i] - gc[i - 1])
    w6[i, 2] = (gc[i + 1] - gc[i]) / (xr[i] - gc[i])
    w6[i, 3] = (gc[i] - gc[i - 1]) / (gc[i] - xr[i - 1])
    w6[i, 4] = (xr[i] - gc[i]) / dx[i]
    w6[i, 5] = (gc[i] - xr[i - 1]) / dx[i]
    q = np.cumsum(rng.standard_normal(n)) * 0.1 + np.sin(np.arange(n) * 0.05)
    q[n // 2:n // 2 + 400] = 10.0 ** rng.uniform(-6, 2, 400)
    if positive:
        q = np.abs(q) + 1e-12
    q = np.ascontiguousarray(q)
    out = [np.zeros(n) for _ in range(4)]
    oracle_lib().oracle_reconstruct_curv(o.h, 0, q.ctypes.data, positive, out[0].ctypes.data, out[1].ctypes.data)
    assert host_lib.idfx_host_reconstruct_irregular(order, n, q.ctypes.data, positive, w6.ctypes.data,
                                                    np.ascontiguousarray(dx).ctypes.data, out[2].ctypes.data,
                                                    out[3].ctypes.data) == 0
    assert np.abs(out[0]).sum() > 0
    assert (out[0].view(np.int64) == out[2].view(np.int64)).all(), "left states differ"
    assert (out[1].view(np.int64) == out[3].view(np.int64)).all(), "right states differ"
    o.close()


@pytest.mark.parametrize("geometry,ncomp,mhd,iso", [(4, 3, 1, 0), (4, 3, 1, 1), (3, 2, 0, 0), (3, 2, 1, 0), (4, 3, 0, 1)])
def test_geometric_source_terms_equal_oracle_bitwise(host_lib, geometry, ncomp, mhd, iso):
    """sweep.cuh::geometricSource (centrifugal / pressure / hoop-stress terms of polar and spherical grids) against the
    oracle's restatement of addSourceTerms.hpp:96-192 on random states of a randomly stretched grid"""
    import math
    rng = np.random.default_rng(31 + geometry + ncomp + 2 * mhd + 4 * iso)
    n1, n2 = 96, 64
    o, grid, ng = _spherical_oracle(n1, n2, ncomp=ncomp, mhd=bool(mhd), geometry=geometry, csiso=0.7 if iso else None, rng=rng)
    nvar = o.nvar
    o.Vc[...] = rng.uniform(-1.5, 1.5, o.Vc.shape)
    o.Vc[0] = 10.0 ** rng.uniform(-1.5, 1.0, o.Vc.shape[1:])
    if not iso:
        o.Vc[ncomp + 1] = 10.0 ** rng.uniform(-2.0, 1.0, o.Vc.shape[1:])
    nt = o.Vc.shape
    Uc = np.ctypeslib.as_array(oracle_lib().oracle_uc(o.h), shape=nt)
    Uc[...] = 0.0
    oracle_lib().oracle_add_source_terms(o.h)
    first_dt = 1e-4   # the oracle's dt before its first cycle
    jj, ii = np.meshgrid(np.arange(ng, ng + n2, dtype=np.int32), np.arange(ng, ng + n1, dtype=np.int32), indexing="ij")
    jj, ii = np.ascontiguousarray(jj.ravel()), np.ascontiguousarray(ii.ravel())
    v = np.ascontiguousarray(o.Vc[:, 0][:, jj, ii].T)          # (n, nvar)
    x1, xl1, xr1, _ = grid[0]
    rt = np.array([(p * p * p - m * m * m) / (p * p - m * m) / 1.5 for p, m in zip(xr1, xl1)])
    tanx2 = np.array([math.tan(t) for t in grid[1][0]])
    s1, s2 = np.zeros(len(jj)), np.zeros(len(jj))
    assert host_lib.idfx_host_geometric_source(geometry, 2, ncomp, mhd, iso, 1.4, 0.7, first_dt, len(jj), v.ctypes.data,
                                               jj.ctypes.data, ii.ctypes.data, np.ascontiguousarray(x1).ctypes.data,
                                               rt.ctypes.data, tanx2.ctypes.data, s1.ctypes.data, s2.ctypes.data) == 0
    ref1 = Uc[1, 0][jj, ii]
    assert np.abs(ref1).sum() > 0
    assert (ref1.view(np.int64) == s1.view(np.int64)).all(), "radial source differs"
    if geometry == 4 and ncomp >= 2:
        ref2 = Uc[2, 0][jj, ii]
        assert np.abs(ref2).sum() > 0
        assert (np.ascontiguousarray(ref2).view(np.int64) == s2.view(np.int64)).all(), "latitudinal source differs"
    o.close()


def _stretched_block(rng, geometry, ndim, ncomp, mhd, n=(10, 9, 8)):
    """an oracle block of `ndim` dimensions in polar (3) / spherical (4) geometry on randomly stretched grids"""
    ng = 2
    grid, np_int = [], []
    ranges = ((1.0, 3.0), (0.4, 2.6) if geometry == 4 else (0.0, 1.5), (0.0, 1.2))
    for a in range(3):
        na = n[a] if a < ndim else 1
        g = ng if a < ndim else 0
        w = rng.uniform(0.6, 1.4, na + 2 * g)
        edges = np.concatenate([[0.0], np.cumsum(w)])
        lo, hi = ranges[a]
        edges = lo + (hi - lo) * (edges - edges[g]) / (edges[g + na] - edges[g])
        xl, xr = edges[:-1].copy(), edges[1:].copy()
        grid.append((0.5 * (xl + xr), xl, xr, xr - xl))
        np_int.append(na)
    o = Oracle(ndim, ncomp, mhd, "hll", "uct_contact", 2, np_int, ["outflow"] * ndim, 1.4, 0.5, 1e-4, [0.1, 0.1, 0.1],
               geometry=geometry, grid=grid)
    return o, grid, ng


@pytest.mark.parametrize("geometry,ndim,ncomp,mhd", [(4, 3, 3, True), (3, 3, 3, True), (4, 2, 3, True), (3, 2, 2, False),
                                                     (4, 3, 3, False)])
def test_curvilinear_update_source_equals_oracle_bitwise(host_lib, geometry, ndim, ncomp, mhd):
    """sweep.cuh::areaWeightFlux + updateCell (area weighting in angular-momentum form, the B_theta / B_phi flux factors,
    the curvilinear right-hand side, the cell crossing time of the CFL estimate, geometric sources on the last sweep)
    against the oracle's Fluid_CorrectFluxFunctor / Fluid_CalcRHSFunctor / AddSourceTerms on random data"""
    import math
    rng = np.random.default_rng(101 + 7 * geometry + ndim + ncomp + int(mhd))
    o, grid, ng = _stretched_block(rng, geometry, ndim, ncomp, mhd)
    L = oracle_lib()
    shape = o.Vc.shape
    ncell = int(np.prod(shape[1:]))
    view = lambda ptr, shp: np.ctypeslib.as_array(ptr, shape=shp)   # noqa: E731
    Uc, Flux = view(L.oracle_uc(o.h), shape), view(L.oracle_flux(o.h), shape)
    cMax, InvDt = view(L.oracle_cmax(o.h), shape[1:]), view(L.oracle_invdt(o.h), shape[1:])
    o.Vc[...] = rng.uniform(-1.5, 1.5, shape)
    o.Vc[0] = 10.0 ** rng.uniform(-1.5, 1.0, shape[1:])
    o.Vc[ncomp + 1] = 10.0 ** rng.uniform(-2.0, 1.0, shape[1:])
    Uc[...] = rng.uniform(-2.0, 2.0, shape)
    InvDt[...] = rng.uniform(0.0, 5.0, shape[1:])
    mine_Uc, mine_InvDt = Uc.copy(), InvDt.copy()
    start_Uc = Uc.copy()
    x1, x1l, x1r, dx1 = grid[0]
    x2, x2l, x2r, dx2 = grid[1]
    coords = [x1, x1l, dx1, np.array([(p * p * p - m * m * m) / (p * p - m * m) / 1.5 for p, m in zip(x1r, x1l)]),
              dx2, np.array([math.sin(t) for t in x2]), np.array([math.sin(t) for t in x2l]),
              np.array([abs(math.cos(a) - math.cos(b)) for a, b in zip(x2l, x2r)]), np.array([math.tan(t) for t in x2]),
              grid[2][3]]
    coords = [np.ascontiguousarray(c, float) for c in coords]
    cptr = (C.c_void_p * len(coords))(*[c.ctypes.data for c in coords])
    nx = (C.c_int * 3)(shape[3], shape[2], shape[1])
    ngs = (C.c_int * 3)(*[ng if a < ndim else 0 for a in range(3)])
    area, vol = np.zeros(shape[1:]), np.zeros(shape[1:])
    for dir in range(ndim):
        raw = rng.uniform(-1.0, 1.0, shape)
        cm = rng.uniform(0.1, 3.0, shape[1:])
        Flux[...] = raw
        cMax[...] = cm
        L.oracle_geometry(o.h, dir, area.ctypes.data, vol.ctypes.data)
        L.oracle_calc_rhs(o.h, dir)
        assert host_lib.idfx_host_curvilinear_rhs(geometry, ndim, ncomp, int(mhd), 0, dir, 1.4, 0.0, 1e-4,
                                                  int(dir == ndim - 1), nx, ngs, cptr, o.Vc.ctypes.data,
                                                  mine_Uc.ctypes.data, mine_InvDt.ctypes.data, raw.ctypes.data,
                                                  cm.ctypes.data, area.ctypes.data, vol.ctypes.data) == 0
    L.oracle_add_source_terms(o.h)
    active = tuple(slice(ng, -ng) if a < ndim else slice(None) for a in (2, 1, 0))
    assert (Uc[(slice(None),) + active] != start_Uc[(slice(None),) + active]).mean() > 0.6   # the update did something
    assert ncell > 0 and (Uc != mine_Uc).sum() == 0, np.argwhere(Uc != mine_Uc)[:5]
    assert (InvDt.view(np.int64) == mine_InvDt.view(np.int64)).all()
    assert (Uc.view(np.int64) == mine_Uc.view(np.int64)).all()
    o.close()
