"""ctypes binding of the C ABI declared in include/idefix_b200.h.

This is the only place where Python touches the shared library.  There is NO
fallback: if ``libidefix_b200.so`` is missing or a call fails, an exception is
raised (IdefixError mirrors the reference's IDEFIX_ERROR, src/error.cpp:14-56).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# IDEFIX_B200_LIB selects another build of the same library (e.g. the DEV experiment build)
LIB_PATH = os.environ.get("IDEFIX_B200_LIB") or os.path.join(_HERE, "libidefix_b200.so")


class IdefixError(RuntimeError):
    """Raised where the reference would call IDEFIX_ERROR or throw std::runtime_error."""


# enumerations (include/idefix_b200.h)
CARTESIAN, CYLINDRICAL, POLAR, SPHERICAL = 1, 2, 3, 4
TVDLF, HLL, HLLC, HLLD, ROE = 0, 1, 2, 3, 4
EMF_ARITHMETIC, EMF_UCT0, EMF_UCT_CONTACT, EMF_UCT_HLL, EMF_UCT_HLLD = 0, 1, 2, 3, 4
BC_INTERNAL, BC_PERIODIC, BC_REFLECTIVE, BC_OUTFLOW, BC_SHEARINGBOX, BC_AXIS, BC_USERDEF, BC_UNDEFINED = range(8)
FIELD_VC, FIELD_UC, FIELD_VS, FIELD_INVDT, FIELD_EX, FIELD_EY, FIELD_EZ, FIELD_UC0, FIELD_VS0, FIELD_VE, FIELD_VE0 = range(11)

GEOMETRY_NAMES = {"CARTESIAN": CARTESIAN, "CYLINDRICAL": CYLINDRICAL, "POLAR": POLAR, "SPHERICAL": SPHERICAL}
SOLVER_NAMES = {"tvdlf": TVDLF, "hll": HLL, "hllc": HLLC, "hlld": HLLD, "roe": ROE}
EMF_NAMES = {"arithmetic": EMF_ARITHMETIC, "uct0": EMF_UCT0, "uct_contact": EMF_UCT_CONTACT,
             "uct_hll": EMF_UCT_HLL, "uct_hlld": EMF_UCT_HLLD}
BC_NAMES = {"internal": BC_INTERNAL, "periodic": BC_PERIODIC, "reflective": BC_REFLECTIVE,
            "outflow": BC_OUTFLOW, "shearingbox": BC_SHEARINGBOX, "axis": BC_AXIS, "userdef": BC_USERDEF}

_pd = C.POINTER(C.c_double)


class IdfxBlockDesc(C.Structure):
    _fields_ = [
        ("ndim", C.c_int), ("ncomp", C.c_int), ("mhd", C.c_int), ("isothermal", C.c_int),
        ("geometry", C.c_int), ("order", C.c_int), ("solver", C.c_int), ("emf", C.c_int),
        ("nstages", C.c_int),
        ("np_int", C.c_int * 3), ("nghost", C.c_int * 3), ("lbound", C.c_int * 3), ("rbound", C.c_int * 3),
        ("regular_grid", C.c_int), ("device", C.c_int),
        ("gamma", C.c_double), ("csiso", C.c_double), ("cfl", C.c_double), ("cfl_max_var", C.c_double),
        ("first_dt", C.c_double), ("fixed_dt", C.c_double),
        ("x", _pd * 3), ("xl", _pd * 3), ("xr", _pd * 3), ("dx", _pd * 3),
        ("ntracer", C.c_int),
        ("shock_flattening", C.c_double),
        ("evolve_vector_potential", C.c_int),
    ]


USER_BOUNDARY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p)
HALO_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p)
REDUCE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p)
STAGE_HOOK_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_double, C.c_void_p)
HOOK_INTERNAL_BOUNDARY, HOOK_EMF_BOUNDARY = 0, 1
REDUCE_MIN, REDUCE_MAX = 0, 1

# every symbol include/idefix_b200.h declares: name -> (restype, argtypes)
_h = C.c_void_p
SYMBOLS = {
    "idfx_last_error": (C.c_char_p, []),
    "idfx_version": (C.c_char_p, []),
    "idfx_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "idfx_block_create": (C.c_int, [C.POINTER(IdfxBlockDesc), C.POINTER(_h)]),
    "idfx_block_destroy": (C.c_int, [_h]),
    "idfx_sync": (C.c_int, [_h]),
    "idfx_field_info": (C.c_int, [_h, C.c_int, C.POINTER(_pd), C.POINTER(C.c_int * 4)]),
    "idfx_upload": (C.c_int, [_h, C.c_int, C.c_void_p]),
    "idfx_download": (C.c_int, [_h, C.c_int, C.c_void_p]),
    "idfx_set_time": (C.c_int, [_h, C.c_double, C.c_double]),
    "idfx_get_time": (C.c_int, [_h, _pd, _pd]),
    "idfx_reset_stage": (C.c_int, [_h]),
    "idfx_set_boundaries": (C.c_int, [_h, C.c_double]),
    "idfx_enforce_boundary_dir": (C.c_int, [_h, C.c_double, C.c_int]),
    "idfx_reconstruct_vc_field": (C.c_int, [_h, C.c_int]),
    "idfx_prim_to_cons": (C.c_int, [_h]),
    "idfx_cons_to_prim": (C.c_int, [_h]),
    "idfx_state_copy": (C.c_int, [_h]),
    "idfx_state_add_and_store": (C.c_int, [_h, C.c_double, C.c_double]),
    "idfx_evolve_stage": (C.c_int, [_h, C.c_double, C.c_double]),
    "idfx_compute_timestep": (C.c_int, [_h, _pd]),
    "idfx_check_divb": (C.c_int, [_h, _pd]),
    "idfx_derive_vector_potential": (C.c_int, [_h]),
    "idfx_check_nan": (C.c_int, [_h, C.POINTER(C.c_int)]),
    "idfx_cycles": (C.c_int, [_h, C.c_int]),
    "idfx_cycle": (C.c_int, [_h, _pd, _pd, C.POINTER(C.c_int)]),
    "idfx_set_tstop": (C.c_int, [_h, C.c_double]),
    "idfx_set_fused": (C.c_int, [_h, C.c_int]),
    "idfx_get_ncycles": (C.c_int, [_h, C.POINTER(C.c_longlong)]),
    "idfx_launch_count": (C.c_int, [_h, C.POINTER(C.c_longlong)]),
    "idfx_math_exceptions": (C.c_int, [_h, C.POINTER(C.c_uint)]),
    "idfx_enroll_userdef_boundary": (C.c_int, [_h, USER_BOUNDARY_FN, C.c_void_p]),
    "idfx_enroll_halo_exchange": (C.c_int, [_h, HALO_FN, C.c_void_p]),
    "idfx_enroll_reduce": (C.c_int, [_h, REDUCE_FN, C.c_void_p]),
    "idfx_enroll_stage_hook": (C.c_int, [_h, C.c_int, STAGE_HOOK_FN, C.c_void_p]),
    "idfx_halo_buffer_size": (C.c_int, [_h, C.c_int, C.POINTER(C.c_size_t)]),
    "idfx_halo_buffers": (C.c_int, [_h, C.c_int, C.c_int, C.POINTER(_pd), C.POINTER(_pd)]),
    "idfx_halo_pack": (C.c_int, [_h, C.c_int]),
    "idfx_halo_unpack": (C.c_int, [_h, C.c_int]),
    "idfx_halo_bind_peer": (C.c_int, [_h, C.c_int, C.c_int, C.c_void_p]),
    "idfx_halo_ipc_handle": (C.c_int, [_h, C.c_int, C.c_int, C.c_void_p]),
    "idfx_halo_open_ipc": (C.c_int, [_h, C.c_void_p, C.POINTER(_pd)]),
    "idfx_set_nproc": (C.c_int, [_h, C.c_int * 3]),
    "idfx_set_stream": (C.c_int, [_h, C.c_void_p]),
    "idfx_timing": (C.c_int, [_h, C.c_int, _pd]),
    "idfx_selftest_tmem": (C.c_int, [C.c_int, C.POINTER(C.c_longlong)]),
    "idfx_selftest_tma": (C.c_int, [C.c_int, C.POINTER(C.c_longlong)]),
    "idfx_p2p_mailbox": (C.c_int, [_h, C.POINTER(C.c_void_p), C.c_void_p]),
    "idfx_p2p_connect": (C.c_int, [_h, C.c_int, C.c_int, (C.c_void_p * 2) * 3, (C.c_void_p * 2) * 3,
                                   C.POINTER(C.c_void_p)]),
    "idfx_p2p_enable_peer_access": (C.c_int, [C.c_int, C.c_int]),
    "idfx_p2p_timeouts": (C.c_int, [_h, C.POINTER(C.c_ulonglong)]),
}

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IdefixError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise IdefixError(lib().idfx_last_error().decode())
