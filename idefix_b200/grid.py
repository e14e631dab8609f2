"""Global grid: the host-side restatement of Grid / GridHost::MakeGrid.

Follows src/grid.cpp:45-269 (sizes, ghost depth, boundary parsing, decomposition) and
src/gridHost.cpp:44-266 (uniform ``u``, logarithmic ``l`` and stretched ``s+`` / ``s-``
patches) expression by expression, using the C library's ``pow``/``log`` through
``math`` so that coordinates are bit-identical with the reference's.
"""
import math
from dataclasses import dataclass

import numpy as np

from . import _abi
from ._abi import IdefixError


@dataclass
class Definitions:
    """The reference's compile-time configuration: definitions.hpp macros
    (src/idefix.hpp:47-108) and CMake options (CMakeLists.txt:12-33)."""
    dimensions: int = 1
    components: int = 1
    geometry: int = _abi.CARTESIAN
    mhd: bool = False
    isothermal: bool = False
    order: int = 2          # Idefix_RECONSTRUCTION: 1 Constant, 2 Linear, 3 LimO3, 4 Parabolic
    evolve_vector_potential: bool = False   # Idefix_EVOLVE_VECTOR_POTENTIAL (idefix.hpp:116-125)

    def __post_init__(self):
        if isinstance(self.geometry, str):
            self.geometry = _abi.GEOMETRY_NAMES[self.geometry.upper()]


class Grid:
    def __init__(self, inp, defs, nproc=(1, 1, 1), xproc=(0, 0, 0)):
        self.defs = defs
        nd = defs.dimensions
        self.isRegularCartesian = defs.geometry == _abi.CARTESIAN   # grid.cpp:49-51
        self.haveAxis = False
        self.nghost = [0, 0, 0]
        self.np_int = [1, 1, 1]
        for d in range(nd):
            self.nghost[d] = 2 if defs.order < 4 else 3              # grid.cpp:61-65
            label = f"X{d + 1}-grid"
            npatch = inp.Get(int, "Grid", label, 0)
            self.np_int[d] = sum(inp.Get(int, "Grid", label, 2 + 3 * p) for p in range(npatch))
        self.np_tot = [self.np_int[d] + 2 * self.nghost[d] for d in range(3)]
        self.lbound = [_abi.BC_UNDEFINED] * 3
        self.rbound = [_abi.BC_UNDEFINED] * 3
        for d in range(nd):
            for side, arr in (("beg", self.lbound), ("end", self.rbound)):
                name = inp.Get(str, "Boundary", f"X{d + 1}-{side}", 0)
                if name not in _abi.BC_NAMES:
                    raise IdefixError(f"Unknown boundary type {name}")
                if name == "axis":
                    if d != 1:
                        raise IdefixError("Axis Boundaries are only applicable to X2")
                    self.haveAxis = True
                arr[d] = _abi.BC_NAMES[name]
        self.nproc = list(nproc)
        self.xproc = list(xproc)
        for d in range(3):
            if self.np_int[d] % self.nproc[d]:
                raise IdefixError("The number of grid points must be a multiple of the domain decomposition")  # grid.cpp:197
        self.x = [None] * 3
        self.xl = [None] * 3
        self.xr = [None] * 3
        self.dx = [None] * 3
        self.xbeg = [0.0] * 3
        self.xend = [0.0] * 3
        self._make_grid(inp)

    # GridHost::MakeGrid (gridHost.cpp:44-266)
    def _make_grid(self, inp):
        defs = self.defs
        for d in range(3):
            n = self.np_tot[d]
            x, xl, xr, dx = (np.zeros(n) for _ in range(4))
            if d < defs.dimensions:
                label = f"X{d + 1}-grid"
                npatch = inp.Get(int, "Grid", label, 0)
                self.xbeg[d] = inp.Get(float, "Grid", label, 1)
                self.xend[d] = inp.Get(float, "Grid", label, 4 + (npatch - 1) * 3)
                ng = self.nghost[d]

                def patch(p):
                    return (inp.Get(str, "Grid", label, 3 + p * 3), inp.Get(float, "Grid", label, 1 + p * 3),
                            inp.Get(float, "Grid", label, 4 + p * 3), inp.Get(int, "Grid", label, 2 + p * 3))

                idx = ng
                for p in range(npatch):
                    ptype, ps, pe, psize = patch(p)
                    gs = ng if p == 0 else 0
                    ge = ng if p == npatch - 1 else 0
                    if ptype == "u":
                        delta = (pe - ps) / psize
                        for i in range(-gs, psize + ge):
                            dx[i + idx] = delta
                            x[i + idx] = ps + (i + 0.5) * delta
                            xl[i + idx] = ps + i * delta
                            xr[i + idx] = ps + (i + 1) * delta
                    elif ptype == "l":
                        self.isRegularCartesian = False
                        alpha = (pe + abs(ps) - ps) / abs(ps)
                        for i in range(-gs, psize + ge):
                            xl[i + idx] = ps * math.pow(alpha, float(i) / float(psize))
                            xr[i + idx] = ps * math.pow(alpha, float(i + 1) / float(psize))
                            dx[i + idx] = xr[i + idx] - xl[i + idx]
                            x[i + idx] = 0.5 * (xr[i + idx] + xl[i + idx])
                    elif ptype not in ("s+", "s-"):
                        raise IdefixError(f"GridHost::MakeGrid: Unknown grid type :{ptype}")
                    idx += psize
                idx = ng
                for p in range(npatch):
                    ptype, ps, pe, psize = patch(p)
                    gs = ng if p == 0 else 0
                    ge = ng if p == npatch - 1 else 0
                    if ptype in ("s+", "s-"):
                        self.isRegularCartesian = False
                        ref = p + 1 if ptype == "s+" else p - 1
                        if ref < 0 or ref >= npatch:
                            raise IdefixError("You're attempting to construct a stretched patch from a non-existent patch")
                        rtype = patch(ref)[0]
                        if rtype != "u" and (rtype != "l" or ref == p + 1):
                            raise IdefixError("You can only construct a stretched patch from a uniform grid "
                                              "or AFTER a logarithmic grid")
                        offset = psize if ref == p + 1 else -1
                        delta = float(dx[idx + offset])
                        logdelta = math.log((pe - ps) / delta)
                        if abs((pe - ps) / psize - delta) < 1e-10:
                            raise IdefixError("A Stretch grid can be defined only if the stretched domain has a mean\n"
                                              "spacing different from the reference uniform grid.")
                        q = 1.05
                        for it in range(51):   # Newton iterations, gridHost.cpp:165-175
                            f = math.log((math.pow(q, psize + 1) - q) / (q - 1)) - logdelta
                            fp = ((psize + 1) * math.pow(q, psize) - 1) / (math.pow(q, psize + 1) - q) - 1 / (q - 1)
                            dq = f / fp
                            q = q - dq
                            if abs(dq) < 1e-14 * q:
                                break
                            if it == 50:
                                raise IdefixError("Failed to create the stretched grid")
                        for i in range(-gs, psize + ge):
                            if ptype == "s-":
                                xl[i + idx] = ps + q * (math.pow(q, i) - 1) / (q - 1) * delta
                                xr[i + idx] = ps + q * (math.pow(q, i + 1) - 1) / (q - 1) * delta
                                dx[i + idx] = math.pow(q, i + 1) * delta
                            else:
                                xl[i + idx] = pe - q * (math.pow(q, psize - i) - 1) / (q - 1) * delta
                                xr[i + idx] = pe - q * (math.pow(q, psize - i - 1) - 1) / (q - 1) * delta
                                dx[i + idx] = math.pow(q, psize - i) * delta
                            x[i + idx] = 0.5 * (xr[i + idx] + xl[i + idx])
                    idx += psize
            else:
                xs, xe = -0.5, 0.5
                if defs.geometry == _abi.SPHERICAL and d == 1:
                    xs, xe = 0.0, math.pi
                self.xbeg[d], self.xend[d] = xs, xe
                dx[:] = xe - xs
                x[:] = 0.5 * (xs + xe)
                xl[:] = xs
                xr[:] = xe
            self.x[d], self.xl[d], self.xr[d], self.dx[d] = x, xl, xr, dx
        # Axis treatment (gridHost.cpp:227-263): the domain must touch the axis, and the theta spacing is mirrored
        # across it (ghost faces rebuilt outwards from the axis face; the cell centres keep their values)
        if self.haveAxis:
            if defs.geometry != _abi.SPHERICAL:
                raise IdefixError("Axis boundaries only compatible with Spherical boundary conditions")
            if defs.dimensions < 2:
                raise IdefixError("Axis Boundaries requires at least two dimenions")
            x, xl, xr, dx = self.x[1], self.xl[1], self.xr[1], self.dx[1]
            if abs(self.xbeg[1]) > 1e-10 and self.lbound[1] == _abi.BC_AXIS:
                raise IdefixError("Axis Boundaries requires your X2 domain to start at exactly x2=0.0")
            if abs(self.xend[1] - math.pi) > 1e-10 and self.rbound[1] == _abi.BC_AXIS:
                raise IdefixError("Axis Boundaries requires your X2 domain to end at exactly x2=Pi")
            if self.lbound[1] == _abi.BC_AXIS:
                jref = self.nghost[1]
                for j in range(jref - 1, -1, -1):
                    dx[j] = dx[2 * jref - j - 1]
                    xl[j] = xl[j + 1] - dx[j]
                    xr[j] = xl[j + 1]
            if self.rbound[1] == _abi.BC_AXIS:
                jref = self.nghost[1] + self.np_int[1] - 1
                for j in range(jref + 1, self.np_tot[1]):
                    dx[j] = dx[2 * jref - j + 1]
                    xr[j] = xr[j - 1] + dx[j]
                    xl[j] = xr[j - 1]

    # DataBlock::ExtractSubdomain (dataBlock.cpp:33-66, 162-190)
    def subdomain(self):
        """coordinates and boundary types of the block (nproc, xproc) selects"""
        out = {"x": [], "xl": [], "xr": [], "dx": [], "np_int": [], "lbound": [], "rbound": [], "gbeg": []}
        for d in range(3):
            npi = self.np_int[d] // self.nproc[d]
            g = self.nghost[d]
            gbeg = g + self.xproc[d] * npi
            sl = slice(gbeg - g, gbeg + npi + g)
            for key, arr in (("x", self.x), ("xl", self.xl), ("xr", self.xr), ("dx", self.dx)):
                out[key].append(np.ascontiguousarray(arr[d][sl]))
            out["np_int"].append(npi)
            out["gbeg"].append(gbeg)
            out["lbound"].append(self.lbound[d] if self.xproc[d] == 0 else _abi.BC_INTERNAL)
            out["rbound"].append(self.rbound[d] if self.xproc[d] == self.nproc[d] - 1 else _abi.BC_INTERNAL)
        return out
