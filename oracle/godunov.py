"""ctypes front end of the CPU oracle (oracle/godunov_oracle.c).

TEST INFRASTRUCTURE ONLY: used by tests/ and __graft_entry__.smoke() as the checker.  The product
(idefix_b200/) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgodunov_oracle.so")

SOLVERS = {"tvdlf": 0, "hll": 1, "hllc": 2, "hlld": 3, "roe": 4}
EMFS = {"arithmetic": 0, "uct0": 1, "uct_contact": 2, "uct_hll": 3, "uct_hlld": 4}
BCS = {"periodic": 1, "reflective": 2, "outflow": 3, "axis": 5}


class OracleDesc(C.Structure):
    _fields_ = [("ndim", C.c_int), ("ncomp", C.c_int), ("mhd", C.c_int), ("solver", C.c_int), ("emf", C.c_int),
                ("nstages", C.c_int), ("np_int", C.c_int * 3), ("nghost", C.c_int * 3), ("bc", C.c_int * 3),
                ("gamma", C.c_double), ("cfl", C.c_double), ("cfl_max_var", C.c_double), ("first_dt", C.c_double),
                ("dx", C.c_double * 3), ("ntracer", C.c_int), ("order", C.c_int), ("isothermal", C.c_int),
                ("csiso", C.c_double), ("shock_flattening", C.c_double), ("evolve_vector_potential", C.c_int),
                ("geometry", C.c_int), ("bc_right", C.c_int * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "godunov_oracle.c")):
            subprocess.run(["make", "-C", HERE], check=True, capture_output=True)
        L = C.CDLL(LIB)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(OracleDesc)]
        for name in ("oracle_vc", "oracle_vs", "oracle_ve"):
            getattr(L, name).restype = C.POINTER(C.c_double)
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ("oracle_time", "oracle_dt"):
            getattr(L, name).restype = C.c_double
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_nvar.argtypes = [C.c_void_p]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_boundaries.argtypes = [C.c_void_p]
        L.oracle_derive_vector_potential.argtypes = [C.c_void_p]
        L.oracle_set_grid.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4
        L.oracle_set_grid.restype = None
        L.oracle_add_source_terms.argtypes = [C.c_void_p]
        L.oracle_add_source_terms.restype = None
        L.oracle_uc.argtypes = [C.c_void_p]
        L.oracle_uc.restype = C.POINTER(C.c_double)
        L.oracle_reconstruct_curv.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_reconstruct_curv.restype = None
        L.oracle_calc_rhs.argtypes = [C.c_void_p, C.c_int]
        L.oracle_calc_rhs.restype = None
        for name in ("oracle_flux", "oracle_cmax", "oracle_invdt"):
            getattr(L, name).restype = C.POINTER(C.c_double)
            getattr(L, name).argtypes = [C.c_void_p]
        L.oracle_geometry.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_geometry.restype = None
        L.oracle_cycle.argtypes = [C.c_void_p]
        L.oracle_riemann.argtypes = [C.c_void_p, C.c_int, C.c_long] + [C.c_void_p] * 5
        L.oracle_riemann.restype = None
        L.oracle_reconstruct.argtypes = [C.c_int, C.c_long, C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_reconstruct.restype = None
        L.oracle_convert.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_convert.restype = None
        L.oracle_run.argtypes = [C.c_void_p, C.c_double, C.c_long]
        _lib = L
    return _lib


class Oracle:
    """One block evolved by the CPU restatement.  Fields are numpy views of the C arrays, indexed
    (n, k, j, i) like the reference's IdefixArray4D."""

    def __init__(self, ndim, ncomp, mhd, solver, emf, nstages, np_int, bc, gamma, cfl, first_dt, dx,
                 cfl_max_var=1.1, ntracer=0, order=2, csiso=None, shock_flattening=0.0, vector_potential=False, geometry=1, grid=None,
                 bc_right=None):
        """geometry 1 cartesian (uniform spacing dx), 3 polar, 4 spherical: `grid` = per direction the arrays (x, xl, xr, dx)
        with ghosts, as GridHost::MakeGrid builds them; bc / bc_right: boundary type names of the left / right sides"""
        if mhd and solver == "tvdlf" and emf in ("uct_hll", "uct_hlld"):
            raise ValueError("TVDLF-MHD stores no upwind coefficients for uct_hlld and extends its sweeps by one "
                             "cell only (tvdlfMHD.hpp:30-41, 226-229): not pinned by any reference run")
        d = OracleDesc()
        d.isothermal = int(csiso is not None)
        d.csiso = csiso or 0.0
        d.ntracer = ntracer
        d.shock_flattening = shock_flattening
        d.evolve_vector_potential = int(bool(vector_potential) and mhd and ndim >= 2)
        d.order = order
        d.ndim, d.ncomp, d.mhd = ndim, ncomp, int(mhd)
        d.solver, d.emf, d.nstages = SOLVERS[solver], EMFS[emf or "uct_contact"], nstages
        for a in range(3):
            d.np_int[a] = np_int[a] if a < ndim else 1
            d.nghost[a] = (3 if order == 4 else 2) if a < ndim else 0
            d.bc[a] = BCS[bc[a]] if a < ndim else 0
            d.bc_right[a] = BCS[bc_right[a]] if (bc_right is not None and a < ndim) else 0
            d.dx[a] = dx[a]
        d.geometry = geometry
        d.gamma, d.cfl, d.cfl_max_var, d.first_dt = gamma, cfl, cfl_max_var, first_dt
        self.desc = d
        self.h = lib().oracle_create(C.byref(d))
        nt = [d.np_int[a] + 2 * d.nghost[a] for a in range(3)]
        self.nvar = lib().oracle_nvar(self.h)
        self.Vc = np.ctypeslib.as_array(lib().oracle_vc(self.h), shape=(self.nvar, nt[2], nt[1], nt[0]))
        self.Vs = None
        if mhd:
            off = [1 if a < ndim else 0 for a in range(3)]
            self.Vs = np.ctypeslib.as_array(lib().oracle_vs(self.h),
                                            shape=(ndim, nt[2] + off[2], nt[1] + off[1], nt[0] + off[0]))
        if geometry != 1:
            if geometry not in (3, 4) or ntracer:
                raise ValueError("curvilinear oracle: polar / spherical, no tracers")
            for a in range(3):
                arrs = [np.ascontiguousarray(q, float) for q in grid[a]]
                assert all(len(q) == nt[a] for q in arrs)
                lib().oracle_set_grid(self.h, a, *[q.ctypes.data for q in arrs])
        self.Ve = None
        if d.evolve_vector_potential:   # AX1e, AX2e, AX3e in 3-D; AX3e alone in 2-D
            self.Ve = np.ctypeslib.as_array(lib().oracle_ve(self.h),
                                            shape=(3 if ndim == 3 else 1, nt[2] + off[2], nt[1] + off[1], nt[0] + off[0]))

    @classmethod
    def from_setup(cls, setup):
        """same configuration as an idefix_b200.setups.Setup built with device=None (host only)"""
        defs, inp, grid = setup.definitions, setup.input, setup.grid
        if defs.order not in (2, 3, 4) or defs.geometry not in (1, 3, 4):
            raise ValueError("the C oracle restates PLM / LimO3 / PPM in cartesian, polar and spherical geometry")
        names = {1: "periodic", 2: "reflective", 3: "outflow", 5: "axis"}
        nd = defs.dimensions
        emf = inp.Get(str, "Hydro", "emf", 0) if inp.CheckEntry("Hydro", "emf") >= 0 else "uct_contact"
        if defs.isothermal and defs.mhd and emf == "uct_hlld":
            raise ValueError("isothermal uct_hlld is restated but pinned by no reference run (the reference crashes "
                             "on it in 2.5-D): usable for solver-level cross-checks only, not as a cycle-level checker")
        first_dt = inp.GetOrSet(float, "TimeIntegrator", "first_dt", 0, 1e-10)
        o = cls(nd, defs.components, defs.mhd, inp.Get(str, "Hydro", "solver", 0), emf,
                inp.Get(int, "TimeIntegrator", "nstages", 0), grid.np_int, [names[grid.lbound[a]] for a in range(nd)],
                setup.data.hydro.gamma, inp.Get(float, "TimeIntegrator", "CFL", 0), first_dt,
                [float(grid.dx[a][0]) for a in range(3)], ntracer=setup.data.hydro.nTracer,
                order=defs.order, csiso=setup.data.hydro.csiso if defs.isothermal else None,
                shock_flattening=setup.data.hydro.shockFlattening,
                vector_potential=getattr(defs, "evolve_vector_potential", False), geometry=defs.geometry,
                grid=[(grid.x[a], grid.xl[a], grid.xr[a], grid.dx[a]) for a in range(3)],
                bc_right=[names[grid.rbound[a]] for a in range(nd)])
        return o

    def set_state(self, Vc, Vs=None, Ve=None):
        """initial condition (InitFlow), then main.cpp:172-176: DeriveVectorPotential, SetBoundaries"""
        self.Vc[...] = Vc
        if self.Ve is not None:
            self.Ve[...] = Ve
            lib().oracle_derive_vector_potential(self.h)
        elif self.Vs is not None:
            self.Vs[...] = Vs
        lib().oracle_set_boundaries(self.h)

    def riemann(self, dir, vL, vR, emf=False):
        """the oracle's interface solver on packed (n, nvar) left / right states -> F, (cmax, sl, sr, keep)
        [, (Et, Eb, SV, aL, aR, dL, dR) as stored for constrained transport]"""
        vL, vR = np.ascontiguousarray(vL, float), np.ascontiguousarray(vR, float)
        F, out = np.zeros_like(vL), np.zeros((vL.shape[0], 4))
        e7 = np.zeros((vL.shape[0], 7)) if emf else None
        lib().oracle_riemann(self.h, dir, vL.shape[0], vL.ctypes.data, vR.ctypes.data, F.ctypes.data, out.ctypes.data,
                             e7.ctypes.data if emf else None)
        return (F, out, e7) if emf else (F, out)

    def cycle(self, n=1):
        for _ in range(n):
            rc = lib().oracle_cycle(self.h)
            if rc:
                raise RuntimeError(f"oracle_cycle failed with status {rc}")

    def run(self, tstop, maxcycles=-1):
        rc = lib().oracle_run(self.h, tstop, maxcycles)
        if rc:
            raise RuntimeError(f"oracle_run failed with status {rc}")

    @property
    def t(self):
        return lib().oracle_time(self.h)

    @property
    def dt(self):
        return lib().oracle_dt(self.h)

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None
