"""DataBlock / Fluid / DataBlockHost mirrors on top of the C ABI.

Names and call meaning follow the reference's interface for the hot path:
``DataBlock`` (src/dataBlock/dataBlock.hpp:54-183) owns one ``Fluid`` (``data.hydro``,
src/fluid/fluid.hpp:53-247) whose methods forward 1:1 to the idfx_* entry points;
``DataBlockHost`` (src/dataBlock/dataBlockHost.hpp:25-102) is the numpy host mirror a
set-up's ``InitFlow`` fills before ``SyncToDevice()``.
All grid work happens in libidefix_b200.so; nothing here computes on field data
except the initial-condition helpers of DataBlockHost (host code in the reference too).
"""
import ctypes as C
import math

import numpy as np

from . import _abi
from ._abi import IdefixError, check, lib

IDIR, JDIR, KDIR = 0, 1, 2


class Fluid:
    """Fluid<DefaultPhysics>: option parsing of the [Hydro] block (fluid.hpp:264-400,
    riemannSolver.hpp:91-140, constrainedTransport.hpp:157-178, eos_*.hpp) + the ABI calls."""

    def __init__(self, inp, data):
        defs = data.defs
        self.data = data
        solver = inp.Get(str, "Hydro", "solver", 0)
        if solver not in _abi.SOLVER_NAMES:
            raise IdefixError(f"Unknown {'MHD' if defs.mhd else 'HD'} solver type {solver}")
        if solver == "hlld" and not defs.mhd:
            raise IdefixError("hlld Riemann solver requires a MHD fluid")
        if solver == "hllc" and defs.mhd:
            raise IdefixError("hllc Riemann solver requires a HD fluid")
        self.solver = _abi.SOLVER_NAMES[solver]
        self.emf = _abi.EMF_UCT_CONTACT
        if defs.mhd and inp.CheckEntry("Hydro", "emf") >= 0:
            name = inp.Get(str, "Hydro", "emf", 0)
            if name not in _abi.EMF_NAMES:
                raise IdefixError("Unknown EMF averaging scheme")
            self.emf = _abi.EMF_NAMES[name]
        self.gamma = 5.0 / 3.0
        self.csiso = 0.0
        if defs.isothermal:
            kind = inp.Get(str, "Hydro", "csiso", 0)
            if kind != "constant":
                raise IdefixError("csiso admits only constant entries in this build")
            self.csiso = inp.Get(float, "Hydro", "csiso", 1)
        else:
            self.gamma = inp.GetOrSet(float, "Hydro", "gamma", 0, 5.0 / 3.0)
        # passive tracers (fluid.hpp:315-323): stored behind the fluid variables, first index TRG = nvar
        self.nTracer = 0
        if inp.CheckEntry("Hydro", "tracer") >= 0:
            self.nTracer = inp.Get(int, "Hydro", "tracer", 0)
            if self.nTracer < 1:
                raise IdefixError("The number of passive tracers should be >= 1")
        self.haveTracer = self.nTracer > 0
        # [Hydro] shockFlattening <smoothing> (riemannSolver.hpp:154-159)
        self.shockFlattening = 0.0
        if inp.CheckEntry("Hydro", "shockFlattening") >= 0:
            self.shockFlattening = inp.Get(float, "Hydro", "shockFlattening", 0)
        for opt in ("viscosity", "resistivity", "ambipolar", "hall", "rotation",
                    "gravity", "fargo", "TDiffusion", "bragViscosity", "bragTDiffusion", "drag"):
            if inp.CheckEntry("Hydro", opt) >= 0:
                raise IdefixError(f"[Hydro] {opt} is outside the scope of this build (explicit Godunov path only)")
        # variable indices (src/idefix.hpp:74-94)
        nc = defs.components
        energy = 0 if defs.isothermal else 1
        self.RHO, self.VX1, self.VX2, self.VX3 = 0, 1, 2, 3
        self.PRS = nc + 1 if energy else None
        self.BX1 = nc + 1 + energy if defs.mhd else None
        self.nvar = 1 + nc + energy + (nc if defs.mhd else 0)
        self.TRG = self.nvar

    # ---- forwards ------------------------------------------------------------------
    def _h(self):
        return self.data._handle

    def ResetStage(self):
        check(lib().idfx_reset_stage(self._h()))

    def ConvertConsToPrim(self):
        check(lib().idfx_cons_to_prim(self._h()))

    def ConvertPrimToCons(self):
        check(lib().idfx_prim_to_cons(self._h()))

    def EvolveStage(self, t, dt):
        check(lib().idfx_evolve_stage(self._h(), t, dt))

    def CheckDivB(self):
        v = C.c_double()
        check(lib().idfx_check_divb(self._h(), C.byref(v)))
        return v.value

    def CheckNan(self):
        n = C.c_int()
        check(lib().idfx_check_nan(self._h(), C.byref(n)))
        return n.value


class DataBlock:
    """One sub-domain on one GPU (dataBlock.cpp:22-160)."""

    def __init__(self, grid, inp, device=0):
        self.mygrid = grid
        self.defs = defs = grid.defs
        sub = grid.subdomain()
        self.np_int = sub["np_int"]
        self.nghost = list(grid.nghost)
        self.np_tot = [self.np_int[d] + 2 * self.nghost[d] for d in range(3)]
        self.beg = list(self.nghost)
        self.end = [self.nghost[d] + self.np_int[d] for d in range(3)]
        self.gbeg = sub["gbeg"]
        self.lbound, self.rbound = sub["lbound"], sub["rbound"]
        self.x, self.xl, self.xr, self.dx = sub["x"], sub["xl"], sub["xr"], sub["dx"]
        self.hydro = Fluid(inp, self)
        # TimeIntegrator options that the device-side dt controller needs (timeIntegrator.cpp:23-85)
        self.nstages = inp.Get(int, "TimeIntegrator", "nstages", 0)
        self.fixed_dt = -1.0
        self.cfl, self.cfl_max_var, first_dt = 0.0, 1.1, 1.0e-10
        if inp.CheckEntry("TimeIntegrator", "fixed_dt") > 0:
            self.fixed_dt = inp.Get(float, "TimeIntegrator", "fixed_dt", 0)
        else:
            self.cfl = inp.Get(float, "TimeIntegrator", "CFL", 0)
            self.cfl_max_var = inp.GetOrSet(float, "TimeIntegrator", "CFL_max_var", 0, 1.1)
            first_dt = inp.GetOrSet(float, "TimeIntegrator", "first_dt", 0, 1.0e-10)

        d = _abi.IdfxBlockDesc()
        d.ndim, d.ncomp = defs.dimensions, defs.components
        d.mhd, d.isothermal = int(defs.mhd), int(defs.isothermal)
        d.geometry, d.order = defs.geometry, defs.order
        d.solver, d.emf, d.nstages = self.hydro.solver, self.hydro.emf, self.nstages
        for k in range(3):
            d.np_int[k], d.nghost[k] = self.np_int[k], self.nghost[k]
            d.lbound[k], d.rbound[k] = self.lbound[k], self.rbound[k]
            d.x[k] = self.x[k].ctypes.data_as(C.POINTER(C.c_double))
            d.xl[k] = self.xl[k].ctypes.data_as(C.POINTER(C.c_double))
            d.xr[k] = self.xr[k].ctypes.data_as(C.POINTER(C.c_double))
            d.dx[k] = self.dx[k].ctypes.data_as(C.POINTER(C.c_double))
        d.regular_grid = int(grid.isRegularCartesian)
        d.device = device or 0
        d.gamma, d.csiso = self.hydro.gamma, self.hydro.csiso
        d.cfl, d.cfl_max_var, d.first_dt, d.fixed_dt = self.cfl, self.cfl_max_var, first_dt, self.fixed_dt
        d.ntracer = self.hydro.nTracer
        d.shock_flattening = self.hydro.shockFlattening
        d.evolve_vector_potential = int(bool(getattr(defs, "evolve_vector_potential", False)))
        self._desc = d
        self._callbacks = []
        self._handle = None
        self.host_mirror = None
        self._fresh = True   # device arrays still hold the zeros of idfx_block_create
        if device is None:
            # host-only block: set-up code (InitFlow) can run, every device call raises.  Used by
            # the CPU tests of the host logic; there is no CPU implementation of the update.
            return
        h = C.c_void_p()
        check(lib().idfx_block_create(C.byref(d), C.byref(h)))
        self._handle = h
        nproc = (C.c_int * 3)(*grid.nproc)
        check(lib().idfx_set_nproc(h, nproc))

    def close(self):
        if getattr(self, "_handle", None):
            lib().idfx_block_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- time -------------------------------------------------------------------------
    @property
    def t(self):
        return self._time()[0]

    @property
    def dt(self):
        return self._time()[1]

    def _time(self):
        if self._handle is None:
            raise IdefixError("host-only DataBlock: no CUDA device block behind it (no CPU fallback)")
        t, dt = C.c_double(), C.c_double()
        check(lib().idfx_get_time(self._handle, C.byref(t), C.byref(dt)))
        return t.value, dt.value

    def set_time(self, t, dt):
        check(lib().idfx_set_time(self._handle, t, dt))

    # ---- forwards (dataBlock.hpp:120-160) ------------------------------------------------
    def SetBoundaries(self):
        check(lib().idfx_set_boundaries(self._handle, self.t))

    def DeriveVectorPotential(self):
        """dataBlock.cpp:366-373: Vs = curl Ve when the vector potential is the evolved variable (no-op otherwise)"""
        if self._handle is not None:
            check(lib().idfx_derive_vector_potential(self._handle))

    def ResetStage(self):
        self.hydro.ResetStage()

    def PrimToCons(self):
        self.hydro.ConvertPrimToCons()

    def ConsToPrim(self):
        self.hydro.ConvertConsToPrim()

    def EvolveStage(self):
        t, dt = self._time()
        self.hydro.EvolveStage(t, dt)

    def ComputeTimestep(self):
        v = C.c_double()
        check(lib().idfx_compute_timestep(self._handle, C.byref(v)))
        return v.value

    def CheckNan(self):
        return self.hydro.CheckNan()

    def state_copy(self):
        """states["begin"].CopyFrom(states["current"])"""
        check(lib().idfx_state_copy(self._handle))

    def state_add_and_store(self, wc, w0):
        """states["current"].AddAndStore(wc, w0, states["begin"])"""
        check(lib().idfx_state_add_and_store(self._handle, wc, w0))

    def sync(self):
        check(lib().idfx_sync(self._handle))

    # ---- fields -------------------------------------------------------------------------
    def field_dims(self, field):
        dims = (C.c_int * 4)()
        p = C.POINTER(C.c_double)()
        check(lib().idfx_field_info(self._handle, field, C.byref(p), C.byref(dims)))
        return tuple(dims), C.cast(p, C.c_void_p).value

    def download(self, field):
        dims, ptr = self.field_dims(field)
        if not ptr or dims[0] == 0:
            raise IdefixError("field is not allocated in this configuration")
        out = np.empty(dims, dtype=np.float64)
        check(lib().idfx_download(self._handle, field, out.ctypes.data))
        return out

    def upload(self, field, arr):
        dims, ptr = self.field_dims(field)
        a = np.ascontiguousarray(arr, dtype=np.float64)
        if a.size != int(np.prod(dims)):
            raise IdefixError(f"upload: array of size {a.shape} does not match field dims {dims}")
        check(lib().idfx_upload(self._handle, field, a.ctypes.data))

    def math_exceptions(self):
        """threads whose branch-free division / sqrt failed a range check (0 = IEEE-exact arithmetic)"""
        n = C.c_uint()
        check(lib().idfx_math_exceptions(self._handle, C.byref(n)))
        return n.value

    def launch_count(self):
        n = C.c_longlong()
        check(lib().idfx_launch_count(self._handle, C.byref(n)))
        return n.value

    # ---- callbacks ------------------------------------------------------------------------
    def EnrollUserDefBoundary(self, fn):
        """fn(data, dir, side, t) with side 0 = left, 1 = right (boundary.hpp:28-35)"""
        cb = _abi.USER_BOUNDARY_FN(lambda h, d, s, t, u: fn(self, d, s, t))
        self._callbacks.append(cb)
        check(lib().idfx_enroll_userdef_boundary(self._handle, cb, None))

    def enroll_halo_exchange(self, fn):
        cb = _abi.HALO_FN(lambda h, d, u: fn(self, d))
        self._callbacks.append(cb)
        check(lib().idfx_enroll_halo_exchange(self._handle, cb, None))


class DataBlockHost:
    """Host mirror filled by InitFlow (dataBlockHost.hpp:25-102)."""

    def __init__(self, data):
        self.data = data
        self.np_tot, self.np_int = data.np_tot, data.np_int
        self.beg, self.end, self.nghost = data.beg, data.end, data.nghost
        self.x, self.xl, self.xr, self.dx = data.x, data.xl, data.xr, data.dx
        n3, n2, n1 = data.np_tot[2], data.np_tot[1], data.np_tot[0]
        self.Vc = np.zeros((data.hydro.nvar + data.hydro.nTracer, n3, n2, n1))
        self._Uc = None      # allocated on first access; an untouched mirror is all zeros
        self._InvDt = None
        self.Vs = None
        self.Ve = None
        if data.defs.mhd:
            nd = data.defs.dimensions
            off = [1 if d < nd else 0 for d in range(3)]
            self.Vs = np.zeros((nd, n3 + off[2], n2 + off[1], n1 + off[0]))
            if getattr(data.defs, "evolve_vector_potential", False):
                # edge-centred potential, laid out like Vs: AX1e, AX2e, AX3e in 3-D, AX3e alone in 2-D (idefix.hpp:116-125)
                self.Ve = np.zeros((3 if nd == 3 else 1, n3 + off[2], n2 + off[1], n1 + off[0]))

    @property
    def Uc(self):
        if self._Uc is None:
            self._Uc = np.zeros_like(self.Vc)
        return self._Uc

    @Uc.setter
    def Uc(self, v):
        self._Uc = v

    @property
    def InvDt(self):
        if self._InvDt is None:
            self._InvDt = np.zeros(self.Vc.shape[1:])
        return self._InvDt

    @InvDt.setter
    def InvDt(self, v):
        self._InvDt = v

    def SyncToDevice(self):
        """pushes everything the mirror holds (dataBlockHost.cpp:108-145): Vc, InvDt, Vs, corner EMFs, Uc"""
        d = self.data
        d.host_mirror = self
        if d._handle is None:
            return
        d.upload(_abi.FIELD_VC, self.Vc)
        # mirrors never touched by the set-up are zero; a freshly created block's device arrays are
        # zero too (idfx_block_create), so the push is skipped only in that case
        fresh = d._fresh
        if self._Uc is not None or not fresh:
            d.upload(_abi.FIELD_UC, self.Uc)
        if self._InvDt is not None or not fresh:
            d.upload(_abi.FIELD_INVDT, self.InvDt)
        if self.Ve is not None:
            d.upload(_abi.FIELD_VE, self.Ve)
        if self.Vs is not None:
            d.upload(_abi.FIELD_VS, self.Vs)
            if not fresh:
                zeros = np.zeros(self.Vc.shape[1:])
                for f in (_abi.FIELD_EX, _abi.FIELD_EY, _abi.FIELD_EZ):
                    if d.field_dims(f)[1]:
                        d.upload(f, zeros)
        d._fresh = False

    def SyncFromDevice(self):
        d = self.data
        self.Vc = d.download(_abi.FIELD_VC)
        self.Uc = d.download(_abi.FIELD_UC)
        self.InvDt = d.download(_abi.FIELD_INVDT)
        if self.Vs is not None:
            self.Vs = d.download(_abi.FIELD_VS)
        if self.Ve is not None:
            self.Ve = d.download(_abi.FIELD_VE)

    def MakeVsFromAmag(self, A):
        """Vs = curl(A) on the staggered mesh, cartesian / polar / spherical
        (dataBlockHost.cpp:180-262).  A is indexed (dir, k, j, i)."""
        data = self.data
        nd = data.defs.dimensions
        geo = data.defs.geometry
        kb, ke = data.beg[2], data.end[2] + (1 if nd == 3 else 0)
        jb, je = data.beg[1], data.end[1] + (1 if nd >= 2 else 0)
        ib, ie = data.beg[0], data.end[0] + 1
        K, J, I = slice(kb, ke), slice(jb, je), slice(ib, ie)
        Kp, Jp, Ip = slice(kb + 1, ke + 1), slice(jb + 1, je + 1), slice(ib + 1, ie + 1)
        dx1 = self.dx[0][I][None, None, :]
        dx2 = self.dx[1][J][None, :, None]
        dx3 = self.dx[2][K][:, None, None]
        x1 = self.x[0][I][None, None, :]
        x1m = self.xl[0][I][None, None, :]
        x1mp = self.xl[0][Ip][None, None, :]   # end+2 <= np_tot because nghost >= 2
        Vs = self.Vs
        if geo == _abi.CARTESIAN:
            b1 = 0.0
            if nd >= 2:
                b1 = b1 + 1 / dx2 * (A[2][K, Jp, I] - A[2][K, J, I])
            if nd == 3:
                b1 = b1 - 1 / dx3 * (A[1][Kp, J, I] - A[1][K, J, I])
            Vs[0][K, J, I] = b1
            if nd >= 2:
                b2 = -1 / dx1 * (A[2][K, J, Ip] - A[2][K, J, I])
                if nd == 3:
                    b2 = b2 + 1 / dx3 * (A[0][Kp, J, I] - A[0][K, J, I])
                Vs[1][K, J, I] = b2
            if nd == 3:
                Vs[2][K, J, I] = 1 / dx1 * (A[1][K, J, Ip] - A[1][K, J, I]) - 1 / dx2 * (A[0][K, Jp, I] - A[0][K, J, I])
        elif geo == _abi.POLAR:
            b1 = 0.0
            if nd >= 2:
                b1 = b1 + 1 / (x1m * dx2) * (A[2][K, Jp, I] - A[2][K, J, I])
            if nd == 3:
                b1 = b1 - 1 / dx3 * (A[1][Kp, J, I] - A[1][K, J, I])
            Vs[0][K, J, I] = b1
            b2 = -1 / dx1 * (A[2][K, J, Ip] - A[2][K, J, I])
            if nd == 3:
                b2 = b2 + 1 / dx3 * (A[0][Kp, J, I] - A[0][K, J, I])
            Vs[1][K, J, I] = b2
            if nd == 3:
                Vs[2][K, J, I] = (1 / (x1 * dx1) * (x1mp * A[1][K, J, Ip] - x1m * A[1][K, J, I])
                                  - 1 / (x1 * dx2) * (A[0][K, Jp, I] - A[0][K, J, I]))
        elif geo == _abi.SPHERICAL:
            x2m_full = np.append(self.xl[1], self.xr[1][-1])
            cosm = np.array([math.cos(v) for v in x2m_full])
            sinm = np.array([math.sin(v) for v in x2m_full])
            cj, cjp = cosm[J][None, :, None], cosm[Jp][None, :, None]
            sj, sjp = sinm[J][None, :, None], sinm[Jp][None, :, None]
            b1 = 0.0
            if nd >= 2:
                b1 = b1 + 1 / (x1m * (cj - cjp)) * (sjp * A[2][K, Jp, I] - sj * A[2][K, J, I])
            if nd == 3:
                b1 = b1 - dx2 / (x1m * (cj - cjp) * dx3) * (A[1][Kp, J, I] - A[1][K, J, I])
            Vs[0][K, J, I] = b1
            Ax2m = np.abs(sj)
            Ax2m = np.where(np.abs(Ax2m) < 1e-12, 1.0, Ax2m)
            b2 = -1 / (x1 * dx1) * (x1mp * A[2][K, J, Ip] - x1m * A[2][K, J, I])
            if nd == 3:
                b2 = b2 + 1 / (x1 * Ax2m * dx3) * (A[0][Kp, J, I] - A[0][K, J, I])
            Vs[1][K, J, I] = b2
            if nd == 3:
                Vs[2][K, J, I] = (1 / (x1 * dx1) * (x1mp * A[1][K, J, Ip] - x1m * A[1][K, J, I])
                                  - 1 / (x1 * dx2) * (A[0][K, Jp, I] - A[0][K, J, I]))
        else:
            raise IdefixError("Not yet defined")
