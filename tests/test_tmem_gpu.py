"""The tensor-memory scratchpad (csrc/tmem.cuh) and the TMA box loads (csrc/tma.cuh) of the sweep kernels: device
self-tests through the C ABI.  (The file sorts first among the GPU tests that launch sweeps only by accident; the
sweeps' own parity tests cover the same paths end to end.)"""
import ctypes as C

import pytest

from idefix_b200 import _abi

pytestmark = pytest.mark.gpu


def test_tmem_scratchpad_roundtrip():
    bad = C.c_longlong(-1)
    assert _abi.lib().idfx_selftest_tmem(0, C.byref(bad)) == 0
    assert bad.value == 0


@pytest.mark.timeout(120)
def test_tma_box_loads_roundtrip():
    bad = C.c_longlong(-1)
    assert _abi.lib().idfx_selftest_tma(0, C.byref(bad)) == 0
    assert bad.value == 0
