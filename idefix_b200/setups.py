"""Problem set-ups of the BASELINE configurations, written against the DataBlockHost mirror
the way a reference ``setup.cpp`` is (``Setup::InitFlow(DataBlock&)``, src/setup.hpp:23-28).

Initial conditions are the closed-form ones of the reference's tests (cited per class); they
are evaluated with the C library's sin/cos/sqrt (``math``), per coordinate or per point, so
that the initial state is bit-identical with what the reference's set-up produces.
"""
import math

import numpy as np

from . import _abi
from .datablock import DataBlock, DataBlockHost
from .grid import Definitions, Grid
from .input import Input
from .timeintegrator import TimeIntegrator

_sin = np.frompyfunc(math.sin, 1, 1)
_cos = np.frompyfunc(math.cos, 1, 1)


def fsin(a):
    return np.asarray(_sin(a), dtype=np.float64)


def fcos(a):
    return np.asarray(_cos(a), dtype=np.float64)


class Setup:
    """Base class: subclasses give ``defs`` (definitions.hpp), ``ini`` (idefix.ini text) and InitFlow."""
    defs = None
    ini = ""

    def __init__(self, order=2, overrides=None, nproc=(1, 1, 1), xproc=(0, 0, 0), device=0, ini=None,
                 vector_potential=False):
        d = self.defs
        self.definitions = Definitions(d.dimensions, d.components, d.geometry, d.mhd, d.isothermal, order,
                                       bool(vector_potential) or d.evolve_vector_potential)
        self.input = Input.from_string(ini if ini is not None else self.ini)
        for (block, name), vals in (overrides or {}).items():
            if vals is None:
                self.input.remove(block, name)
            else:
                self.input.set(block, name, *(vals if isinstance(vals, (list, tuple)) else [vals]))
        self.grid = Grid(self.input, self.definitions, nproc, xproc)
        self.data = DataBlock(self.grid, self.input, device)
        self.tint = TimeIntegrator(self.input, self.data)
        self.tstop = self.input.Get(float, "TimeIntegrator", "tstop", 0)

    def initialise(self):
        """main.cpp:159-177: InitFlow, DeriveVectorPotential, then SetBoundaries"""
        self.InitFlow(self.data)
        self.data.DeriveVectorPotential()
        self.data.SetBoundaries()
        return self

    def InitFlow(self, data):
        raise NotImplementedError


class SodHD(Setup):
    """test/HD/sod/setup.cpp:33-41, idefix-hllc.ini (C1)"""
    defs = Definitions(1, 1, _abi.CARTESIAN, mhd=False)
    ini = """
[Grid]
X1-grid    1  0.0  500  u  1.0
[TimeIntegrator]
CFL         0.8
tstop       0.2
first_dt    1.e-4
nstages     2
[Hydro]
solver    hllc
gamma     1.4
[Boundary]
X1-beg    outflow
X1-end    outflow
"""

    def InitFlow(self, data):
        d = DataBlockHost(data)
        right = d.x[0] > 0.5
        d.Vc[0] = np.where(right, 0.125, 1.0)[None, None, :]
        d.Vc[1] = 0.0
        if not data.defs.isothermal:
            d.Vc[2] = np.where(right, 0.1, 1.0)[None, None, :]
        d.SyncToDevice()


class SodMHD(Setup):
    """test/MHD/sod/setup.cpp:26-46, idefix.ini (1-D Brio-Wu tube, 2 components)"""
    defs = Definitions(1, 2, _abi.CARTESIAN, mhd=True)
    ini = """
[Grid]
X1-grid    1  0.0  800  u  100.0
[TimeIntegrator]
CFL         0.8
tstop       10.0
first_dt    1.e-4
nstages     2
[Hydro]
solver    roe
gamma     1.66666666666667
[Boundary]
X1-beg    outflow
X1-end    outflow
"""

    def InitFlow(self, data):
        d = DataBlockHost(data)
        h = data.hydro
        left = (d.x[0] < 50.0)[None, None, :]
        d.Vc[h.RHO] = np.where(left, 1.0, 0.125)
        d.Vc[h.VX1] = 0.0
        if h.PRS is not None:
            d.Vc[h.PRS] = np.where(left, 1.0, 0.1)
        d.Vc[h.BX1] = 0.75
        d.Vc[h.BX1 + 1] = np.where(left, 1.0, -1.0)
        # the reference's loop stops at np_tot: the last face keeps its initial 0
        d.Vs[0][:, :, :d.np_tot[0]] = 0.75
        d.SyncToDevice()


class OrszagTang2D(Setup):
    """test/MHD/OrszagTang/setup.cpp:26-62, idefix-hlld-hlld.ini (C2 at 1024^2)"""
    defs = Definitions(2, 2, _abi.CARTESIAN, mhd=True)
    ini = """
[Grid]
X1-grid    1  0.0  128  u  1.0
X2-grid    1  0.0  128  u  1.0
[TimeIntegrator]
CFL         0.6
tstop       0.5
first_dt    1.e-4
nstages     2
[Hydro]
solver    hlld
emf       uct_hlld
[Boundary]
X1-beg    periodic
X1-end    periodic
X2-beg    periodic
X2-end    periodic
"""

    def InitFlow(self, data):
        d = DataBlockHost(data)
        n3, n2, n1 = d.np_tot[2], d.np_tot[1], d.np_tot[0]
        B0 = 1.0 / math.sqrt(4.0 * math.pi)
        x = d.x[0][None, None, :]
        y = d.x[1][None, :, None]
        d.Vc[0] = 25.0 / (36.0 * math.pi)
        d.Vc[3] = 5.0 / (12.0 * math.pi)
        d.Vc[1] = np.broadcast_to(-fsin(2.0 * math.pi * y), (n3, n2, n1))
        d.Vc[2] = np.broadcast_to(fsin(2.0 * math.pi * x), (n3, n2, n1))
        if data.defs.evolve_vector_potential:   # setup.cpp:41-45: A_z on the cell corners
            xl = d.xl[0][None, None, :]
            yl = d.xl[1][None, :, None]
            d.Ve[0][:n3, :n2, :n1] = B0 / (2.0 * math.pi) * (fcos(2.0 * math.pi * yl) + fcos(4.0 * math.pi * xl) / 2.0)
        else:
            # the loops of the reference stop at np_tot: the last face of each staggered array stays 0
            d.Vs[0][:n3, :n2, :n1] = np.broadcast_to(-B0 * fsin(2.0 * math.pi * y), (n3, n2, n1))
            d.Vs[1][:n3, :n2, :n1] = np.broadcast_to(B0 * fsin(4.0 * math.pi * x), (n3, n2, n1))
        d.SyncToDevice()


class OrszagTang3D(Setup):
    """test/MHD/OrszagTang3D/setup.cpp:213-262 without tracers (C4 at 512^3; the published benchmark)"""
    defs = Definitions(3, 3, _abi.CARTESIAN, mhd=True)
    ini = """
[Grid]
X1-grid    1  0.0  32  u  1.0
X2-grid    1  0.0  64  u  1.0
X3-grid    1  0.0  32  u  1.0
[TimeIntegrator]
CFL         0.9
tstop       0.2
first_dt    1.e-4
nstages     2
[Hydro]
solver    hlld
[Boundary]
X1-beg    periodic
X1-end    periodic
X2-beg    periodic
X2-end    periodic
X3-beg    periodic
X3-end    periodic
"""

    def InitFlow(self, data):
        d = DataBlockHost(data)
        n3, n2, n1 = d.np_tot[2], d.np_tot[1], d.np_tot[0]
        shape = (n3, n2, n1)
        B0 = 1.0 / math.sqrt(4.0 * math.pi)
        x = d.x[0][None, None, :]
        y = d.x[1][None, :, None]
        z = d.x[2][:, None, None]
        d.Vc[0] = 25.0 / (36.0 * math.pi)
        d.Vc[4] = 5.0 / (12.0 * math.pi)
        d.Vc[1] = np.broadcast_to(-fsin(2.0 * math.pi * y), shape)
        d.Vc[2] = fsin(2.0 * math.pi * x) + fcos(2.0 * math.pi * z)
        d.Vc[3] = np.broadcast_to(fcos(2.0 * math.pi * x), shape)
        xl = d.xl[0][None, None, :]
        yl = d.xl[1][None, :, None]
        Ve = np.zeros((3, n3 + 1, n2 + 1, n1 + 1))
        Ve[0][:n3, :n2, :n1] = np.broadcast_to(B0 / (2.0 * math.pi) * (fcos(2.0 * math.pi * yl)), shape)
        Ve[1][:n3, :n2, :n1] = np.broadcast_to(B0 / (2.0 * math.pi) * fsin(2.0 * math.pi * xl), shape)
        Ve[2][:n3, :n2, :n1] = B0 / (2.0 * math.pi) * (fcos(2.0 * math.pi * yl) + fcos(4.0 * math.pi * xl) / 2.0)
        if data.hydro.haveTracer:   # test/MHD/OrszagTang3D/setup.cpp:247-250: two step functions
            trg = data.hydro.TRG
            d.Vc[trg] = np.broadcast_to(np.where(x > 0.5, 1.0, 0.0), shape)
            if data.hydro.nTracer > 1:
                d.Vc[trg + 1] = np.broadcast_to(np.where(z > 0.5, 1.0, 0.0), shape)
        if data.defs.evolve_vector_potential:   # setup.cpp:217-223: the potential itself is the evolved variable
            d.Ve[...] = Ve
        else:
            d.MakeVsFromAmag(Ve)
        d.SyncToDevice()


class LinearWave3D(Setup):
    """test/MHD/LinearWaveTest/setup.cpp:22-215 (Gardiner & Stone 2008 oblique waves; C3 at 128^3)"""
    defs = Definitions(3, 3, _abi.CARTESIAN, mhd=True)
    ini = """
[Grid]
X1-grid    1  0.0  64  u  3.0
X2-grid    1  0.0  32  u  1.5
X3-grid    1  0.0  32  u  1.5
[TimeIntegrator]
CFL         0.9
tstop       0.5
first_dt    1.e-4
nstages     2
[Hydro]
solver    roe
[Boundary]
X1-beg    periodic
X1-end    periodic
X2-beg    periodic
X2-end    periodic
X3-beg    periodic
X3-end    periodic
[Setup]
mode      1
epsilon   1e-6
"""

    def InitFlow(self, data):
        inp = self.input
        mode = inp.GetOrSet(int, "Setup", "mode", 0, 1)
        epsilon = inp.GetOrSet(float, "Setup", "epsilon", 0, 1e-6)
        d = DataBlockHost(data)
        gamma = data.hydro.gamma
        RHO, VX1, VX2, VX3, PRS, BX1, BX2, BX3 = range(8)
        B = [0.0] * 8
        R = [0.0] * 8
        B[RHO], B[BX1], B[BX2], B[PRS] = 1, 1.0, 3.0 / 2.0, 1 / gamma
        amode = abs(mode)
        sign = mode // amode
        if amode == 1:
            R0 = 1.0 / (2 * math.sqrt(5))
            R[RHO], R[VX1], R[VX2], R[BX2] = 2, sign * 4.0, -sign * 2.0, 4.0
            R[PRS] = (gamma - 1) * (9.0 - B[BX1] * R[BX1] - B[BX2] * R[BX2] - B[BX3] * R[BX3])
        elif amode == 2:
            R0 = 1.0
            R[VX3], R[BX3] = -sign, 1.0
        elif amode == 3:
            R0 = 1.0 / (2 * math.sqrt(5))
            R[RHO], R[VX1], R[VX2], R[BX2] = 4, 2 * sign, 4 * sign, -2.0
            R[PRS] = (gamma - 1) * (3.0 - B[BX1] * R[BX1] - B[BX2] * R[BX2] - B[BX3] * R[BX3])
        elif amode == 4:
            R0 = 0.5
            B[VX1] = 1
            R[RHO] = 2
            R[PRS] = (gamma - 1) * (1.0 - B[BX1] * R[BX1] - B[BX2] * R[BX2] - B[BX3] * R[BX3] - B[VX1] * R[RHO] / 2.0)
        else:
            raise _abi.IdefixError("unknown mode")
        sinA = 2.0 / 3.0
        sinB = 2.0 / math.sqrt(5.0)
        cosA = math.sqrt(1 - sinA * sinA)
        cosB = math.sqrt(1 - sinB * sinB)
        e1 = (cosA * cosB, cosA * sinB, sinA)
        e2 = (-sinB, cosB, 0)
        e3 = (-sinA * cosB, -sinA * sinB, cosA)

        def g(a, dirn):
            return [a[None, None, :], a[None, :, None], a[:, None, None]][dirn]

        x, y, z = g(d.x[0], 0), g(d.x[1], 1), g(d.x[2], 2)
        xl, yl, zl = g(d.xl[0], 0), g(d.xl[1], 1), g(d.xl[2], 2)
        x1 = cosA * cosB * x + cosA * sinB * y + sinA * z
        phase = epsilon * fcos(2.0 * math.pi * x1)
        d.Vc[RHO] = B[RHO] + R0 * R[RHO] * phase
        d.Vc[PRS] = B[PRS] + R0 * R[PRS] * phase
        for c, n in enumerate((VX1, VX2, VX3)):
            d.Vc[n] = (B[VX1] * e1[c] + B[VX2] * e2[c] + B[VX3] * e3[c]
                       + R0 * (R[VX1] * e1[c] + R[VX2] * e2[c] + R[VX3] * e3[c]) * phase)
        A = np.zeros((3,) + d.Vc[0].shape)
        for c, (xx, yy, zz) in enumerate(((x, yl, zl), (xl, y, zl), (xl, yl, z))):
            x1 = cosA * cosB * xx + cosA * sinB * yy + sinA * zz
            x2 = -sinB * xx + cosB * yy
            ph = epsilon * fsin(2.0 * math.pi * x1) / (2.0 * math.pi)
            A1 = 0.0
            A2 = B[BX3] * x1 + R0 * R[BX3] * ph
            A3 = -(B[BX2] * x1 + R0 * R[BX2] * ph) + B[BX1] * x2
            A[c] = A1 * e1[c] + A2 * e2[c] + A3 * e3[c]
        d.MakeVsFromAmag(A)
        d.SyncToDevice()


class SphericalMHD(Setup):
    """C5: spherical 3-D MHD on a logarithmic radial grid away from the axis:
    rigidly rotating gas threaded by a uniform vertical field, A_phi = B0 r sin(theta) / 2."""
    defs = Definitions(3, 3, _abi.SPHERICAL, mhd=True)
    ini = """
[Grid]
X1-grid    1  1.0  48  l  10.0
X2-grid    1  0.7853981633974483  24  u  2.356194490192345
X3-grid    1  0.0  48  u  6.283185307179586
[TimeIntegrator]
CFL         0.5
tstop       0.1
first_dt    1.e-5
nstages     2
[Hydro]
solver    roe
gamma     1.6666666666666667
[Boundary]
X1-beg    outflow
X1-end    outflow
X2-beg    reflective
X2-end    reflective
X3-beg    periodic
X3-end    periodic
"""

    def InitFlow(self, data):
        d = DataBlockHost(data)
        n3, n2, n1 = d.np_tot[2], d.np_tot[1], d.np_tot[0]
        B0, omega = 0.1, 0.1
        r = d.x[0][None, None, :]
        th = d.x[1][None, :, None]
        d.Vc[0] = 1.0
        d.Vc[1] = 0.0
        d.Vc[2] = 0.0
        d.Vc[3] = np.broadcast_to(omega * r * fsin(th), (n3, n2, n1))
        d.Vc[4] = 1.0
        A = np.zeros((3, n3, n2, n1))
        rl = d.xl[0][None, None, :]
        thl = d.xl[1][None, :, None]
        A[2] = np.broadcast_to(0.5 * B0 * rl * fsin(thl), (n3, n2, n1))
        d.MakeVsFromAmag(A)
        d.SyncToDevice()


REGISTRY = {"sod": SodHD, "mhdsod": SodMHD, "ot2d": OrszagTang2D, "ot3d": OrszagTang3D,
            "linearwave": LinearWave3D, "spherical": SphericalMHD}
