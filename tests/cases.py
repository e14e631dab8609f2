"""Parity cases shared by tests/golden/make_golden.py (which runs the reference binaries of
oracle/_ref to produce the fixtures) and the GPU parity tests.

SHORT cases: reduced grids, a few cycles, one dump per cycle ([Output] dmp 1e-7), fixtures
``tests/golden/short/<name>.c<NNNN>.dmp`` generated HERE from the reference's own build.
FULL cases: the reference's regression runs (testme.py of each test) compared against copies of
its golden dumps ``tests/golden/reference/<dir>/<file>`` with the reference's own tolerance.
"""
from collections import namedtuple

Short = namedtuple("Short", "name setup order ref overrides cycles")
Full = namedtuple("Full", "name setup order overrides golden tol")

_G = ("Grid",)


def _grid(*dims):
    return {("Grid", f"X{d + 1}-grid"): v for d, v in enumerate(dims)}


def _ot2d(n, solver, emf, nstages=2):
    o = _grid(["1", "0.0", str(n), "u", "1.0"], ["1", "0.0", str(n), "u", "1.0"])
    o[("Hydro", "solver")] = solver
    o[("Hydro", "emf")] = emf
    o[("TimeIntegrator", "nstages")] = nstages
    return o


def _ot3d(n1, n2, n3, solver="hlld", emf=None, nstages=2):
    o = _grid(["1", "0.0", str(n1), "u", "1.0"], ["1", "0.0", str(n2), "u", "1.0"], ["1", "0.0", str(n3), "u", "1.0"])
    o[("Hydro", "solver")] = solver
    o[("Hydro", "emf")] = emf
    o[("TimeIntegrator", "nstages")] = nstages
    return o


def _lw(n1, n2, n3, solver, nstages, mode):
    o = _grid(["1", "0.0", str(n1), "u", "3.0"], ["1", "0.0", str(n2), "u", "1.5"], ["1", "0.0", str(n3), "u", "1.5"])
    o[("Hydro", "solver")] = solver
    o[("TimeIntegrator", "nstages")] = nstages
    o[("Setup", "mode")] = mode
    return o


def _sod(solver, nstages=2):
    return {("Hydro", "solver"): solver, ("TimeIntegrator", "nstages"): nstages}


SHORT = [
    # C1: HD sod, every HD solver x reconstruction
    Short("sod_hllc_plm", "sod", 2, "sod_plm", _sod("hllc"), (1, 5, 20)),
    Short("sod_hll_plm", "sod", 2, "sod_plm", _sod("hll"), (1, 20)),
    Short("sod_tvdlf_plm", "sod", 2, "sod_plm", _sod("tvdlf"), (1, 20)),
    Short("sod_hllc_limo3", "sod", 3, "sod_limo3", _sod("hllc"), (1, 20)),
    Short("sod_hllc_ppm_rk3", "sod", 4, "sod_ppm", _sod("hllc", 3), (1, 20)),
    # 1-D MHD shock tube
    Short("mhdsod_roe_plm", "mhdsod", 2, "mhdsod_plm", _sod("roe"), (1, 20)),
    Short("mhdsod_hlld_plm", "mhdsod", 2, "mhdsod_plm", _sod("hlld"), (1, 20)),
    Short("mhdsod_hll_plm", "mhdsod", 2, "mhdsod_plm", _sod("hll"), (1, 20)),
    Short("mhdsod_tvdlf_plm", "mhdsod", 2, "mhdsod_plm", _sod("tvdlf"), (1, 20)),
    # C2 physics: 2-D Orszag-Tang, every solver x EMF averaging
    Short("ot2d_hlld_ucthlld_plm", "ot2d", 2, "ot2d_plm", _ot2d(64, "hlld", "uct_hlld"), (1, 2, 10)),
    Short("ot2d_hlld_ucthll_plm", "ot2d", 2, "ot2d_plm", _ot2d(64, "hlld", "uct_hll"), (1, 10)),
    Short("ot2d_hlld_contact_plm", "ot2d", 2, "ot2d_plm", _ot2d(64, "hlld", "uct_contact"), (1, 10)),
    Short("ot2d_hlld_uct0_plm", "ot2d", 2, "ot2d_plm", _ot2d(64, "hlld", "uct0"), (1, 10)),
    Short("ot2d_hlld_arith_plm", "ot2d", 2, "ot2d_plm", _ot2d(64, "hlld", "arithmetic"), (1, 10)),
    Short("ot2d_hll_ucthll_plm", "ot2d", 2, "ot2d_plm", _ot2d(64, "hll", "uct_hll"), (1, 10)),
    Short("ot2d_roe_contact_plm", "ot2d", 2, "ot2d_plm", _ot2d(64, "roe", "uct_contact"), (1, 10)),
    Short("ot2d_tvdlf_contact_plm", "ot2d", 2, "ot2d_plm", _ot2d(64, "tvdlf", "uct_contact"), (1, 10)),
    Short("ot2d_hlld_ucthlld_limo3", "ot2d", 3, "ot2d_limo3", _ot2d(64, "hlld", "uct_hlld"), (1, 10)),
    Short("ot2d_hlld_ucthlld_ppm_rk3", "ot2d", 4, "ot2d_ppm", _ot2d(64, "hlld", "uct_hlld", 3), (1, 10)),
    # C4 physics: 3-D Orszag-Tang
    Short("ot3d_hlld_contact_plm", "ot3d", 2, "ot3d_plm", _ot3d(16, 32, 16), (1, 2, 10)),
    Short("ot3d_hlld_ucthlld_plm", "ot3d", 2, "ot3d_plm", _ot3d(16, 32, 16, emf="uct_hlld"), (1, 10)),
    # passive tracers: the stock test/MHD/OrszagTang3D/idefix.ini carries two
    Short("ot3d_hlld_contact_plm_tr2", "ot3d", 2, "ot3d_plm", {**_ot3d(16, 32, 16), ("Hydro", "tracer"): "2"}, (1, 10)),
    Short("ot3d_hll_uct0_plm_rk3_tr2", "ot3d", 2, "ot3d_plm",
          {**_ot3d(16, 32, 16, solver="hll", emf="uct0", nstages=3), ("Hydro", "tracer"): "2"}, (1, 10)),
    Short("ot3d_roe_contact_plm", "ot3d", 2, "ot3d_plm", _ot3d(16, 32, 16, solver="roe"), (1, 10)),
    Short("ot3d_hll_uct0_plm_rk3", "ot3d", 2, "ot3d_plm", _ot3d(16, 32, 16, solver="hll", emf="uct0", nstages=3), (1, 10)),
    # shock flattening ([Hydro] shockFlattening, shockFlattening.hpp): flagged cells fall back on the reference's MinMod
    Short("sod_hllc_plm_flat", "sod", 2, "sod_plm", {**_sod("hllc"), ("Hydro", "shockFlattening"): "0.5"}, (1, 20, 60)),
    Short("sod_hll_limo3_flat", "sod", 3, "sod_limo3", {**_sod("hll"), ("Hydro", "shockFlattening"): "0.2"}, (1, 40)),
    Short("ot2d_hlld_ucthlld_plm_flat", "ot2d", 2, "ot2d_plm",
          {**_ot2d(64, "hlld", "uct_hlld"), ("Hydro", "shockFlattening"): "0.0005"}, (1, 10)),
    Short("ot3d_hlld_contact_plm_flat_rk3", "ot3d", 2, "ot3d_plm",
          {**_ot3d(16, 32, 16, nstages=3), ("Hydro", "shockFlattening"): "0.001"}, (1, 10)),
    # EVOLVE_VECTOR_POTENTIAL (evolveVectorPotential.hpp): reference binaries built with -DEVOLVE_VECTOR_POTENTIAL
    # (ref name ending in _vp <=> the set-up is created with vector_potential=True)
    Short("ot2d_hlld_ucthlld_plm_vp", "ot2d", 2, "ot2d_plm_vp", _ot2d(64, "hlld", "uct_hlld"), (1, 10)),
    Short("ot2d_roe_contact_plm_rk3_vp", "ot2d", 2, "ot2d_plm_vp", _ot2d(64, "roe", "uct_contact", 3), (1, 10)),
    Short("ot3d_hlld_contact_plm_vp", "ot3d", 2, "ot3d_plm_vp", _ot3d(16, 32, 16), (1, 10)),
    Short("ot3d_hll_uct0_plm_rk3_vp", "ot3d", 2, "ot3d_plm_vp", _ot3d(16, 32, 16, solver="hll", emf="uct0", nstages=3), (1, 10)),
    # C3 physics: linear waves, HLLD + PPM + RK3 has no shipped golden -> reference binary
    Short("lw_fast_hlld_ppm_rk3", "linearwave", 4, "lw_ppm", _lw(32, 16, 16, "hlld", 3, 1), (1, 10)),
    Short("lw_alfven_roe_limo3", "linearwave", 3, "lw_limo3", _lw(32, 16, 16, "roe", 2, 2), (1, 10)),
    Short("lw_slow_hlld_plm_rk3", "linearwave", 2, "lw_plm", _lw(32, 16, 16, "hlld", 3, 3), (1, 10)),
]

# Curvilinear cases (BASELINE config 5 physics and polar geometry).  Their problem directories
# (idefix_b200/host/problems/<dir>: definitions.hpp + setup.cpp + idefix.ini, written against the
# reference's user API) are compiled twice -- by oracle/ref_build against the REFERENCE sources (the
# fixtures) and by idefix_b200/host against our host layer (the run under test, tests/test_host_cpp.py).
PROBLEM_DIRS = {"axis": "AxisMHD", "axis25d": "Axis25D", "othooks": "otHooks", "sodstrong": "sodStrong", "sph": "SphericalMHD", "polar": "PolarMHD", "sodiso": "sodIso", "iso25d": "IsoMHD25D", "sph25d": "Spherical25D", "polarhd": "PolarHD2D"}


def _sph(solver, emf=None, nstages=2):
    o = _grid(["1", "1.0", "24", "l", "4.0"], ["1", "0.785398163397448", "12", "u", "2.356194490192345"],
              ["1", "0.0", "24", "u", "6.283185307179586"])
    o[("Hydro", "solver")] = solver
    o[("Hydro", "emf")] = emf
    o[("TimeIntegrator", "nstages")] = nstages
    return o


def _axis(solver, emf=None, nstages=2, x2=("0.0", "3.141592653589793"), x3=("0.0", "6.283185307179586"), beg="axis",
          end="axis", n=(16, 16, 16)):
    """problems/AxisMHD: a spherical shell with the polar axis inside the domain (axis.cpp): theta from 0 and / or to pi"""
    o = _grid(["1", "1.0", str(n[0]), "l", "4.0"], ["1", x2[0], str(n[1]), "u", x2[1]], ["1", x3[0], str(n[2]), "u", x3[1]])
    o[("Hydro", "solver")] = solver
    o[("Hydro", "emf")] = emf
    o[("TimeIntegrator", "nstages")] = nstages
    o[("Boundary", "X2-beg")] = beg
    o[("Boundary", "X2-end")] = end
    return o


def _polar(solver, emf=None, nstages=2):
    o = _grid(["2", "1.0", "12", "u", "2.0", "4", "s-", "3.0"], ["1", "0.0", "16", "u", "6.283185307179586"],
              ["1", "-0.5", "8", "u", "0.5"])
    o[("Hydro", "solver")] = solver
    o[("Hydro", "emf")] = emf
    o[("TimeIntegrator", "nstages")] = nstages
    return o


HOST_SHORT = [
    # C5 physics: spherical, log-r grid, Roe + LimO3 + uct_contact (Roe goes with uct_contact)
    Short("sph_roe_contact_limo3", "sph", 3, "sph_limo3", _sph("roe"), (1, 10)),
    Short("sph_hlld_ucthlld_plm_rk3", "sph", 2, "sph_plm", _sph("hlld", "uct_hlld", 3), (1, 10)),
    Short("sph_hll_uct0_plm", "sph", 2, "sph_plm", _sph("hll", "uct0"), (1, 10)),
    Short("sph_tvdlf_arith_limo3", "sph", 3, "sph_limo3", _sph("tvdlf", "arithmetic"), (1, 10)),
    Short("polar_hlld_contact_plm", "polar", 2, "polar_plm", _polar("hlld"), (1, 10)),
    # the polar axis inside the domain (axis.cpp): full 2 pi (ghosts from the opposite meridian, Ex1
    # averaged around the pole, B_theta on the axis from its neighbours), a quarter of the circle (Ex1 = 0 and
    # BX2s = 0 on the axis), an axis on one side only, and the vector-potential variant (Ex3 = 0 on the axis matters)
    Short("axis_hlld_contact_plm", "axis", 2, "axis_plm", _axis("hlld"), (1, 10)),
    Short("axis_hlld_ucthlld_plm_rk3", "axis", 2, "axis_plm", _axis("hlld", "uct_hlld", 3, n=(12, 20, 24)), (1, 10)),
    Short("axis_quarter_hll_uct0_plm_rk3", "axis", 2, "axis_plm", _axis("hll", "uct0", 3, x3=("0.0", "1.5707963267948966")), (1, 10)),
    Short("axis_north_roe_contact_limo3", "axis", 3, "axis_limo3",
          _axis("roe", x2=("0.0", "1.5707963267948966"), end="outflow"), (1, 10)),
    Short("axis_south_tvdlf_arith_plm", "axis", 2, "axis_plm",
          _axis("tvdlf", "arithmetic", x2=("1.5707963267948966", "3.141592653589793"), beg="outflow"), (1, 10)),
    Short("axis_hlld_contact_plm_vp", "axis", 2, "axis_plm_vp", _axis("hlld"), (1, 10)),
    # 2.5-D spherical block between two axis boundaries (problems/Axis25D): no average around the pole, B_theta = 0 and
    # E_phi = 0 on the axis
    Short("axis25d_hlld_contact_plm", "axis25d", 2, "axis25d_plm", {("Hydro", "solver"): "hlld"}, (1, 10)),
    Short("axis25d_roe_uct0_plm_rk3", "axis25d", 2, "axis25d_plm",
          {("Hydro", "solver"): "roe", ("Hydro", "emf"): "uct0", ("TimeIntegrator", "nstages"): 3}, (1, 10)),
    Short("axis_hlld_ucthlld_plm", "axis", 2, "axis_plm", _axis("hlld", "uct_hlld"), (1, 10)),
    Short("axis_hll_contact_plm_rk3", "axis", 2, "axis_plm", _axis("hll", None, 3, n=(12, 20, 24)), (1, 10)),
    Short("axis_south_hlld_contact_plm", "axis", 2, "axis_plm",
          _axis("hlld", x2=("1.5707963267948966", "3.141592653589793"), beg="outflow"), (1, 10)),
    # shock flattening in curvilinear geometry (area-weighted divergence, shockFlattening.hpp:78-87)
    Short("sph_hlld_contact_plm_flat", "sph", 2, "sph_plm", {**_sph("hlld"), ("Hydro", "shockFlattening"): "0.0002"}, (1, 10)),
    Short("polar_hll_uct0_plm_flat", "polar", 2, "polar_plm",
          {**_polar("hll", "uct0"), ("Hydro", "shockFlattening"): "0.0002"}, (1, 10)),
    Short("polar_hll_ucthll_plm_rk3", "polar", 2, "polar_plm", _polar("hll", "uct_hll", 3), (1, 10)),
    # isothermal EOS: HD tube, and 2.5-D MHD (COMPONENTS 3 > DIMENSIONS 2: B_z through Uc)
    Short("sodiso_hllc_plm", "sodiso", 2, "sodiso_plm", _sod("hllc"), (1, 20)),
    Short("sodiso_hll_plm_rk3", "sodiso", 2, "sodiso_plm", _sod("hll", 3), (1, 20)),
    Short("sodiso_tvdlf_plm", "sodiso", 2, "sodiso_plm", _sod("tvdlf"), (1, 20)),
    Short("iso25d_hlld_contact_plm", "iso25d", 2, "iso25d_plm", {("Hydro", "solver"): "hlld"}, (1, 10)),
    Short("iso25d_roe_contact_limo3", "iso25d", 3, "iso25d_limo3", {("Hydro", "solver"): "roe", ("Hydro", "emf"): None}, (1, 10)),
    Short("iso25d_hll_uct0_plm_rk3", "iso25d", 2, "iso25d_plm",
          {("Hydro", "solver"): "hll", ("Hydro", "emf"): "uct0", ("TimeIntegrator", "nstages"): 3}, (1, 10)),
    # 2.5-D spherical (r, theta) with three components: disc wedge, outflow / reflective sides
    Short("sph25d_hlld_contact_plm", "sph25d", 2, "sph25d_plm", {("Hydro", "solver"): "hlld"}, (1, 10)),
    Short("sph25d_roe_uct0_plm_rk3", "sph25d", 2, "sph25d_plm",
          {("Hydro", "solver"): "roe", ("Hydro", "emf"): "uct0", ("TimeIntegrator", "nstages"): 3}, (1, 10)),
    # hydrodynamics in 2-D polar geometry: HLLC + PPM + RK3, HLL + LimO3
    Short("polarhd_hllc_ppm_rk3", "polarhd", 4, "polarhd_ppm", {("Hydro", "solver"): "hllc"}, (1, 10)),
    Short("polarhd_hll_limo3", "polarhd", 3, "polarhd_limo3",
          {("Hydro", "solver"): "hll", ("TimeIntegrator", "nstages"): 2}, (1, 10)),
    Short("iso25d_tvdlf_arith_limo3", "iso25d", 3, "iso25d_limo3",
          {("Hydro", "solver"): "tvdlf", ("Hydro", "emf"): "arithmetic"}, (1, 10)),
    # user hooks of the stage (enroll.hpp:33,71): internal boundary (density floor) and EMF boundary (scaled ez)
    Short("othooks_identity", "othooks", 2, "othooks_plm", _ot2d(64, "hlld", "uct_hlld"), (1, 10)),
    Short("othooks_floor_scaled_emf", "othooks", 2, "othooks_plm",
          {**_ot2d(64, "hlld", "uct_contact", 3), ("Setup", "rhofloor"): "0.2215", ("Setup", "emfscale"): "0.75"}, (1, 10)),
]

FULL = [
    Full("sod_hllc", "sod", 2, _sod("hllc"), "HD/sod/dump.ref.double.plm.idefix-hllc.ini.dmp", 0.0),
    Full("sod_hll", "sod", 2, _sod("hll"), "HD/sod/dump.ref.double.plm.idefix-hll.ini.dmp", 0.0),
    Full("sod_tvdlf", "sod", 2, _sod("tvdlf"), "HD/sod/dump.ref.double.plm.idefix-tvdlf.ini.dmp", 0.0),
    Full("sod_hllc_limo3", "sod", 3, _sod("hllc"), "HD/sod/dump.ref.double.limo3.idefix-hllc.ini.dmp", 0.0),
    Full("sod_hllc_ppm", "sod", 4, _sod("hllc"), "HD/sod/dump.ref.double.ppm.idefix-hllc.ini.dmp", 0.0),
    Full("sod_hllc_ppm_rk3", "sod", 4, _sod("hllc", 3), "HD/sod/dump.ref.double.ppm.idefix-hllc-rk3.ini.dmp", 0.0),
    Full("mhdsod_roe", "mhdsod", 2, _sod("roe"), "MHD/sod/dump.ref.double.plm.idefix.ini.dmp", 1e-14),
    Full("mhdsod_hlld", "mhdsod", 2, _sod("hlld"), "MHD/sod/dump.ref.double.plm.idefix-hlld.ini.dmp", 1e-14),
    Full("mhdsod_hll", "mhdsod", 2, _sod("hll"), "MHD/sod/dump.ref.double.plm.idefix-hll.ini.dmp", 1e-14),
    Full("mhdsod_tvdlf", "mhdsod", 2, _sod("tvdlf"), "MHD/sod/dump.ref.double.plm.idefix-tvdlf.ini.dmp", 1e-14),
    Full("ot2d_hlld_hlld", "ot2d", 2, _ot2d(128, "hlld", "uct_hlld"),
         "MHD/OrszagTang/dump.ref.double.plm.idefix-hlld-hlld.ini.dmp", 1e-12),
    Full("ot2d_hlld_hll", "ot2d", 2, _ot2d(128, "hlld", "uct_hll"),
         "MHD/OrszagTang/dump.ref.double.plm.idefix-hlld-hll.ini.dmp", 1e-12),
    Full("ot2d_hlld_contact", "ot2d", 2, _ot2d(128, "hlld", None),
         "MHD/OrszagTang/dump.ref.double.plm.idefix-hlld.ini.dmp", 1e-12),
    Full("ot2d_hlld_uct0", "ot2d", 2, _ot2d(128, "hlld", "uct0"),
         "MHD/OrszagTang/dump.ref.double.plm.idefix-hlld-uct0.ini.dmp", 1e-12),
    Full("ot2d_hlld_arithmetic", "ot2d", 2, _ot2d(128, "hlld", "arithmetic"),
         "MHD/OrszagTang/dump.ref.double.plm.idefix-hlld-arithmetic.ini.dmp", 1e-12),
    Full("ot2d_hll", "ot2d", 2, _ot2d(128, "hll", None), "MHD/OrszagTang/dump.ref.double.plm.idefix-hll.ini.dmp", 1e-12),
    Full("ot2d_tvdlf", "ot2d", 2, _ot2d(128, "tvdlf", None),
         "MHD/OrszagTang/dump.ref.double.plm.idefix-tvdlf.ini.dmp", 1e-12),
    Full("ot2d_roe", "ot2d", 2, _ot2d(128, "roe", None), "MHD/OrszagTang/dump.ref.double.plm.idefix.ini.dmp", 1e-12),
    Full("ot2d_hlld_hlld_limo3", "ot2d", 3, _ot2d(128, "hlld", "uct_hlld"),
         "MHD/OrszagTang/dump.ref.double.limo3.idefix-hlld-hlld.ini.dmp", 1e-12),
    Full("ot2d_hlld_hlld_ppm", "ot2d", 4, _ot2d(128, "hlld", "uct_hlld"),
         "MHD/OrszagTang/dump.ref.double.ppm.idefix-hlld-hlld.ini.dmp", 1e-12),
    Full("ot3d_hlld_contact", "ot3d", 2, _ot3d(32, 64, 32), "MHD/OrszagTang3D/dump.ref.double.plm.idefix.ini.dmp", 1e-13),
    # the same shipped golden with the two tracers of the stock ini switched on: Vc-TR0 / Vc-TR1 are compared too
    Full("ot3d_hlld_contact_tracers", "ot3d", 2, {**_ot3d(32, 64, 32), ("Hydro", "tracer"): "2"},
         "MHD/OrszagTang3D/dump.ref.double.plm.idefix.ini.dmp", 1e-13),
    # the reference's shipped vector-potential goldens (testme.py of test/MHD/OrszagTang with -vectPot)
    Full("ot2d_hlld_hlld_vp", "ot2d_vp", 2, _ot2d(128, "hlld", "uct_hlld"),
         "MHD/OrszagTang/dump.ref.double.plm.idefix-hlld-hlld.ini.vectPot.dmp", 1e-12),
    Full("ot2d_hlld_contact_vp", "ot2d_vp", 2, _ot2d(128, "hlld", None),
         "MHD/OrszagTang/dump.ref.double.plm.idefix-hlld.ini.vectPot.dmp", 1e-12),
    Full("ot2d_roe_vp", "ot2d_vp", 2, _ot2d(128, "roe", None), "MHD/OrszagTang/dump.ref.double.plm.idefix.ini.vectPot.dmp", 1e-12),
    Full("lw_fast_roe_plm", "linearwave", 2, _lw(64, 32, 32, "roe", 2, 1),
         "MHD/LinearWaveTest/dump.ref.double.plm.idefix-fast.ini.dmp", 1e-14),
    Full("lw_alfven_roe_ppm", "linearwave", 4, {**_lw(64, 32, 32, "roe", 3, 2), ("TimeIntegrator", "tstop"): "1.0"},
         # the reference's own build in this image scores 3.408e-14 against this shipped golden
         # (different libm/compiler than the machine that wrote it); our GPU result is that same number
         "MHD/LinearWaveTest/dump.ref.double.ppm.idefix-alfven.ini.dmp", 5e-14),
]


# The HD Roe solver of the reference's default sod test (tests/test_roe_hd_iso_gpu.py on the GPU, tests/test_oracle.py
# for the oracle).  sodstrong_*: a 1e5 pressure ratio drives interfaces through the solver's strong-shock switch,
# which in one dimension leaves the flux of the previous call in place (roeHD.hpp:312-328) -- reproduced.
ORACLE_SHORT = [
    Short("sod_roe_plm", "sod", 2, "sod_plm", _sod("roe"), (1, 20, 60)),
    Short("sod_roe_ppm_rk3", "sod", 4, "sod_ppm", _sod("roe", 3), (1, 40)),
    Short("sodiso_roe_plm", "sodiso", 2, "sodiso_plm", _sod("roe"), (1, 40)),
    Short("sodstrong_roe_plm", "sodstrong", 2, "sodstrong_plm", _sod("roe"), (1, 10, 60)),
    Short("sodstrong_roe_plm_rk3", "sodstrong", 2, "sodstrong_plm", _sod("roe", 3), (1, 40)),
]
ORACLE_FULL = [
    Full("sod_roe", "sod", 2, _sod("roe"), "HD/sod/dump.ref.double.plm.idefix.ini.dmp", 0.0),
    Full("sod_roe_ppm_rk3", "sod", 4, _sod("roe", 3), "HD/sod/dump.ref.double.ppm.idefix-rk3.ini.dmp", 0.0),
]


# BASELINE.json configurations at (or, for C4 / C5, as close as one reference CPU run in this container allows to) the
# sizes they are quoted on.  Fixtures: tests/golden/config_size/<name>.json = SHA-256 of every dumped field after N
# cycles of the reference's own build (tests/golden/make_config_size.py); GPU test: tests/test_config_size_gpu.py.
CONFIG_SIZE = [
    Short("c1_sod_hllc_500", "sod", 2, "sod_plm", _sod("hllc"), (5, 100)),
    Short("c2_ot2d_1024", "ot2d", 2, "ot2d_plm", _ot2d(1024, "hlld", "uct_hlld"), (3,)),
    Short("c3_lw_128_hlld_ppm_rk3", "linearwave", 4, "lw_ppm", _lw(128, 128, 128, "hlld", 3, 1), (3,)),
    Short("c4_ot3d_256", "ot3d", 2, "ot3d_plm", _ot3d(256, 256, 256), (3,)),
]


def _sph_sized(n1, n2, n3, solver):
    o = _grid(["1", "1.0", str(n1), "l", "10.0"], ["1", "0.785398163397448", str(n2), "u", "2.356194490192345"],
              ["1", "0.0", str(n3), "u", "6.283185307179586"])
    o[("Hydro", "solver")] = solver
    return o


# C5 physics (spherical, log-r, Roe + LimO3 + uct_contact) at the reduced size SURVEY 8(d) names, through the C++ host
CONFIG_SIZE_HOST = [
    Short("c5_sph_96x48x96_roe_limo3", "sph", 3, "sph_limo3", _sph_sized(96, 48, 96, "roe"), (3,)),
]
CONFIG_SIZE = CONFIG_SIZE + CONFIG_SIZE_HOST


def ini_for(case, dump_every_cycle=False):
    """ini text of a case = the set-up's default ini with the case's overrides applied"""
    import os
    from idefix_b200.input import Input
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if case.setup in PROBLEM_DIRS:
        inp = Input(os.path.join(here, "idefix_b200", "host", "problems", PROBLEM_DIRS[case.setup], "idefix.ini"))
        for name in ("dmp", "vtk"):
            inp.remove("Output", name)
    else:
        from idefix_b200.setups import REGISTRY
        name = case.setup[:-3] if case.setup.endswith("_vp") else case.setup   # *_vp: the same set-up, vector potential
        inp = Input.from_string(REGISTRY[name].ini)
    for (block, name), vals in case.overrides.items():
        if vals is None:
            inp.remove(block, name)
        else:
            inp.set(block, name, *(vals if isinstance(vals, (list, tuple)) else [vals]))
    if dump_every_cycle:
        inp.set("Output", "dmp", "1e-7")   # smaller than any dt: one dump per cycle
    inp.set("Output", "log", "1000000")
    return inp.dumps()
